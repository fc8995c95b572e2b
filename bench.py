"""Headline benchmark: train images/sec of U-Net++ (resnet18 encoder, 2 classes) at 512x512,
full train step (forward + Dice/CE loss + backward + Adam), batch 16 per GPU (BASELINE.json
configs[1]), on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One JSON line on rank 0 (see DESIGN.md "Measurement" for every key).  `value` is measured with the
batch already resident in HBM; `e2e` through the public API with pinned host batches copied in and
the loss read back every step; `roofline` is the dominant kernel (the tcgen05 implicit-GEMM
convolution: every fprop and dgrad launch of the step) timed with CUDA events on its stream.
`--impl reference` times the reference's CPU path (the oracle restatement of smp U-Net++ in fp32
PyTorch eager + torch.optim.Adam, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
import torch  # noqa: E402

METRIC = "train images/sec (512x512, U-Net++)"
WORKLOAD = ("BASELINE configs[1]: U-Net++ (resnet18 encoder, random init) binary train step "
            "fwd + 0.5*Dice+0.5*CE + bwd + Adam(lr 1e-3, wd 1e-5), batch %d per GPU @ %dx%d")
H = W = 512
CLASSES = 2
BATCH_PER_GPU = 16
CPU_SAMPLE_BATCH = 2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"tflops": p.get("bf16_tflops_sustained", p.get("bf16_tflops")), "hbm": p.get("hbm_gbs"),
                "src": "MEASURED_PEAKS.json (bf16_tflops_sustained: kernel timed inside a long step)"}
    return {"tflops": 1400.0, "hbm": 6650.0, "src": "fallback of B200_PROFILING.md (sustained 1.4 PFLOP/s)"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                 getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                 getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                 getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def synthetic(n, gen_seed=6210, pinned=False):
    g = torch.Generator().manual_seed(gen_seed)
    u = torch.rand((n, 3, H, W), generator=g)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    x = ((u - mean) / std).contiguous()
    y = torch.randint(0, CLASSES, (n, H, W), generator=g)
    if pinned:
        x, y = x.pin_memory(), y.pin_memory()
    return x, y


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_train_rate(steps, warmup, batch=CPU_SAMPLE_BATCH):
    """The reference's CPU path for this workload: fp32 eager PyTorch U-Net++ (oracle restatement of
    smp.UnetPlusPlus) + 0.5*Dice + 0.5*CE + torch.optim.Adam(lr 1e-3, wd 1e-5), all host threads."""
    from oracle.losses import mixed_loss
    from oracle.unetpp import UnetPlusPlus
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(6210)
    model = UnetPlusPlus("resnet18", None, 3, CLASSES).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    x, y = synthetic(batch)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for p in model.parameters():
            p.grad = None
        loss = mixed_loss(model(x), y, 0.5)
        loss.backward()
        opt.step()
        float(loss)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * len(times) / total, total / len(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 6)), max(1, min(args.warmup, 2))
    rate, sec, cores = cpu_train_rate(steps, warmup)
    sample = "%d timed steps of batch %d @ %dx%d (of the batch-%d workload), fp32 eager, %d threads" % (
        steps, CPU_SAMPLE_BATCH, H, W, BATCH_PER_GPU, cores)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD % (BATCH_PER_GPU, H, W), "global_batch": BATCH_PER_GPU * args.gpus,
                       "classes": CLASSES, "parallelism": "cpu, %d threads" % cores,
                       "sample_batch_per_step": CPU_SAMPLE_BATCH},
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
def kernels_per_call(lib, fn):
    two = (lib.mmr_wgrad_plan_run, lib.mmr_wgrad_halo_plan_run, lib.mmr_head_grad_prep)
    if not hasattr(fn, "restype"):      # scheduling marker of the engine's launch lists, not a kernel
        return 0
    return 2 if any(fn is f for f in two) else 1


def conv_kernel_time(eng, n_iter=3):
    """CUDA-event time of every implicit-GEMM conv launch (fprop + dgrad: conv_halo_kernel for the 3x3
    stride-1 layers and the stem, conv_gemm_tc_kernel for the strided / 1x1 ones) of one step, on the
    launching stream; returns (ms per step spent in those kernels, launches per step, ms and launches of the
    dgrad launches whose epilogue also takes a BatchNorm-backward reduction)."""
    lib = eng.lib
    stream = torch.cuda.current_stream()
    s = stream.cuda_stream
    import ctypes as C
    sp = C.c_void_p(s)
    fused = getattr(eng, "fused_dgrad_handles", set())
    total_ms, launches, fused_ms, fused_n = 0.0, 0, 0.0, 0
    for it in range(n_iter):
        evs = []
        for calls in (eng.repack_calls, eng.fwd_calls, eng.bwd_calls[False]):
            for fn, a in calls:
                if fn is lib.mmr_conv_plan_run or fn is lib.mmr_halo_conv_plan_run:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    fn(*a, sp)
                    e1.record(stream)
                    evs.append((e0, e1, getattr(a[0], "value", None) in fused))
                else:
                    fn(*a, sp)
        torch.cuda.synchronize()
        if it > 0:
            total_ms += sum(a.elapsed_time(b) for a, b, _ in evs)
            fused_ms += sum(a.elapsed_time(b) for a, b, f in evs if f)
            launches, fused_n = len(evs), sum(1 for _, _, f in evs if f)
    return total_ms / (n_iter - 1), launches, fused_ms / (n_iter - 1), fused_n


def run_ours(args):
    import torch.distributed as dist
    from mmrseg_b200 import _lib
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    from mmrseg_b200.models import UnetPlusPlus
    from mmrseg_b200.optim import FusedAdam
    from mmrseg_b200.parallel import DistributedDataParallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: whatever libraries print (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a B200: the CUDA path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    torch.manual_seed(6210)
    model = UnetPlusPlus("resnet18", classes=CLASSES).to(dev).train()
    crit = DiceCrossEntropyLoss(0.5)
    opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    model._ensure_flat(dev)
    ddp = DistributedDataParallel(model) if world > 1 else None
    if ddp is not None:
        ddp.sync_parameters()                         # rank 0's weights everywhere before step 1
    n = BATCH_PER_GPU
    xh, yh = synthetic(n, 6210 + rank, pinned=True)
    xd, yd = xh.to(dev), yh.to(dev)

    def step_resident():
        for p in model.parameters():
            p.grad = None
        loss = crit(model(xd), yd)
        loss.backward()
        opt.step()
        return loss

    # the same step fed with uint8 HWC frames (what the loaders hold): /255 and utils.normalize on the device
    fh = (torch.rand((n, H, W, 3), generator=torch.Generator().manual_seed(6210 + rank)) * 255).to(torch.uint8).pin_memory()
    model.set_input_normalization((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))

    # the loop a user writes: batches come through the package's DevicePrefetcher (one batch in flight on a
    # copy stream); every timed step still moves its own inputs host -> device and reads its loss back
    from mmrseg_b200.data import DevicePrefetcher

    feed = DevicePrefetcher(None, dev)      # one prefetcher (copy stream + two device buffer sets) for the run

    def run_e2e(host_x, steps):
        last = None
        feed.loader = [(host_x, yh)] * steps
        for x, y in feed:
            for p in model.parameters():
                p.grad = None
            loss = crit(model(x), y)
            loss.backward()
            opt.step()
            last = loss.item()      # device -> host read of the step's result
        return last

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    step_resident()                                   # builds the plan (tensor maps, buffers)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(step_resident, args.steps)
    clocks = sampler.result()
    def timed_loop(host_x):
        run_e2e(host_x, 3)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(host_x, args.steps)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    ms_e2e = timed_loop(xh)
    ms_e2e_u8 = timed_loop(fh)

    eng = model._engine_for(xd, training=True)
    line = None
    if rank == 0:
        pk = peaks()
        conv_ms, conv_launches, fused_ms, fused_n = conv_kernel_time(eng)
        conv_flops = sum(u["fplan"].flops for u in eng.units if "fplan" in u) + \
            sum(u["dplan"].flops for u in eng.units if "dplan" in u)
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12
        # the same without the dgrad launches that also carry a BatchNorm-backward reduction (separate kernel
        # instantiation conv_halo_kernel<3>, extra HBM reads and ALU work that are not conv FLOPs)
        fused_flops = sum(u["dplan"].flops for u in eng.units
                          if "dplan" in u and getattr(u["dplan"].handle, "value", None) in eng.fused_dgrad_handles)
        conv_only = (conv_flops - fused_flops) / ((conv_ms - fused_ms) * 1e-3) / 1e12 if conv_ms > fused_ms else None
        lib = eng.lib
        per_step = sum(kernels_per_call(lib, fn) for calls in (eng.repack_calls, eng.fwd_calls, eng.bwd_calls[False])
                       for fn, _ in calls) + 3 + 1   # + loss fwd (2 kernels) + loss bwd + Adam
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        step_flops = eng.conv_flops_fwd + eng.conv_flops_bwd
        line = {
            "metric": METRIC, "value": world * n * args.steps / (ms * 1e-3), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": WORKLOAD % (n, H, W),
                       "global_batch": world * n, "classes": CLASSES, "parallelism": "dp%d" % world,
                       "l2": "per-step working set %.1f GB >> 126 MB L2 (no flush needed)" % (
                           (eng.arena_bytes + sum(a.buf.numel() * a.buf.element_size() for a in eng.acts.values() if a.buf is not None)) / 1e9),
                       "conv_tflop_per_step": step_flops / 1e12},
            "clocks": clocks,
            "e2e": {"value": world * n * args.steps / (ms_e2e * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": xh.numel() * 4 + yh.numel() * 8, "d2h_bytes_per_step": 4},
            "e2e_uint8_frames": {"value": world * n * args.steps / (ms_e2e_u8 * 1e-3), "unit": "images/s",
                                 "h2d_bytes_per_step": fh.numel() + yh.numel() * 8, "d2h_bytes_per_step": 4,
                                 "note": "same step through model(frames_u8): ToTensor + utils.normalize fused into the stem loader"},
            "gpu_launches": per_step * args.steps,
            "roofline": {"bound": "tensor", "kernel": "conv_halo_kernel + conv_gemm_tc_kernel (all fprop + dgrad launches of a step; their epilogues also take the BatchNorm batch statistics (fprop) and the BatchNorm-backward sums of single-reader units (dgrad), which is not counted as work)",
                         "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                         "traffic": traffic, "launches_per_step": conv_launches, "kernel_ms_per_step": conv_ms,
                         "algorithmic_tflop_per_step": conv_flops / 1e12, "peak_source": pk["src"],
                         "conv_only_launches": {"launches": conv_launches - fused_n, "achieved": conv_only,
                                                "frac": conv_only / pk["tflops"] if conv_only else None,
                                                "note": "excluding the %d dgrad launches (conv_halo_kernel<3>) whose epilogue "
                                                        "also reduces a BatchNorm backward" % fused_n}},
            "step_tensor_frac": step_flops / (ms / args.steps * 1e-3) / 1e12 / pk["tflops"],
        }
    if world > 1:
        dist.barrier()
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            rate, sec, cores = cpu_train_rate(2, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": "2 timed steps (1 warm-up) of batch %d @ %dx%d of the same train step, fp32 eager "
                                              "PyTorch oracle (restated smp U-Net++) + torch.optim.Adam" % (CPU_SAMPLE_BATCH, H, W)}
        else:
            line["cpu_baseline"] = None
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
