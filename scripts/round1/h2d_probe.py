import torch, time
dev = torch.device("cuda", 0)
for mb in (12, 33, 84):
    h = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device=dev)
    s = torch.cuda.Stream()
    for name, stream in (("main", torch.cuda.current_stream()), ("side", s)):
        with torch.cuda.stream(stream):
            for _ in range(2):
                d.copy_(h, non_blocking=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            for _ in range(5):
                d.copy_(h, non_blocking=True)
            e1.record(stream)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            print("%3d MB %s: %.3f ms per copy (%.1f GB/s), host issue %.3f ms, pinned %s" % (
                mb, name, e0.elapsed_time(e1) / 5, mb * 1.048576 / (e0.elapsed_time(e1) / 5), (t1 - t0) * 1e3 / 5, h.is_pinned()))
