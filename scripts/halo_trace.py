"""Where the fixed per-launch time of conv_halo_kernel goes: per-CTA %globaltimer stamps (MMR_HALO_DBG=16) of one
launch of a layer, printed as (min, median, max over CTAs) relative to the first CTA's entry.
usage: MMR_HALO_DBG=16 python scripts/halo_trace.py [layer-substring ...]"""
import ctypes as C
import os
import sys

os.environ["MMR_HALO_DBG"] = str(int(os.environ.get("MMR_HALO_DBG", "0")) | 16)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from mmrseg_b200 import _lib, convplan  # noqa: E402
from scripts.bench_halo import LAYERS  # noqa: E402

N = int(os.environ.get("BATCH", "16"))
names = sys.argv[1:] or ["layer1", "layer4", "x_3_3.conv2", "x_1_3"]
lib = _lib.lib()
gen = torch.Generator(device="cuda").manual_seed(0)
flush = None
labels = ["entry", "dep-wait", "setup", "1st operands", "last MMA", "last store", "finalised", "exit"]
for name, hw, srcs, cout in LAYERS:
    if not any(n in name for n in names):
        continue
    sources = [((torch.randn((N, hw // up, hw // up, c), generator=gen, device="cuda")).to(torch.bfloat16), up)
               for c, up in srcs]
    cin = sum(c for c, _ in srcs)
    w = torch.randn((cout, cin, 3, 3), generator=gen, device="cuda") / (9 * cin) ** 0.5
    out = torch.empty((N, hw, hw, cout), device="cuda", dtype=torch.bfloat16)
    for kind in ("fprop", "dgrad"):
        if kind == "fprop":
            stats = torch.zeros((8 * 2 * cout,), device="cuda", dtype=torch.float64) if os.environ.get("STATS", "1") != "0" else None
            plan = convplan.build_fprop_halo(sources, w, out, stats=stats, stats_ld=cout)
        else:
            dz = torch.randn((N, hw, hw, cout), generator=gen, device="cuda").to(torch.bfloat16)
            grads = [torch.empty((N, hw, hw, c), device="cuda", dtype=torch.bfloat16) for c, _ in srcs]
            plan = convplan.build_dgrad_halo(dz, w, grads)
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if os.environ.get("FLUSH"):
            # cold operands, as inside a step: 256 MB written between the launches, ONE traced launch
            if flush is None:
                flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
            flush.zero_()
            e0.record()
            plan.run()
            e1.record()
            torch.cuda.synchronize()
            e0.elapsed_time(e1)
            e1 = e0   # the per-launch figure below is meaningless for a single cold launch
        else:
            e0.record()
            for _ in range(5):
                plan.run()
            e1.record()
            torch.cuda.synchronize()
        n_ctas = min(148, plan.cfg.get("total_items", 148)) if isinstance(plan.cfg, dict) else 148
        buf = np.zeros((256, 8), dtype=np.uint64)
        _lib.check(lib.mmr_debug_halo_trace(buf.ctypes.data_as(C.c_void_p), 256))
        t = buf.astype(np.int64)
        live = t[:, 0] > 0
        t = t[live][:148]
        t0 = t[:, 0].min()
        rel = (t - t0) / 1e3
        print("%-14s %-5s %.1f us/launch (back to back), %d CTAs, cfg %s" % (
            name, kind, e0.elapsed_time(e1) / 5 * 1e3, t.shape[0],
            {k: plan.cfg[k] for k in ("bn", "tx", "rph", "tps") if k in plan.cfg}))
        for i, lab in enumerate(labels):
            col = rel[:, i]
            print("    %-13s min %7.2f  median %7.2f  max %7.2f us" % (lab, col.min(), np.median(col), col.max()))
        del plan
