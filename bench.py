"""Headline benchmark: train images/sec of U-Net++ (resnet18 encoder, 2 classes) at 512x512,
full train step (forward + Dice/CE loss + backward + Adam), batch 16 per GPU (BASELINE.json
configs[1]), on N GPUs of one node -- plus the other BASELINE configs behind `--config`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

  c2 (default)  U-Net++-R18 binary train step, batch 16 per GPU @ 512x512 (weak scaling)      BASELINE configs[1]
  c3            ResNetUNet-34 10-class train step, batch 32 per GPU @ 512x512                 configs[2]
  c4            U-Net++-R18 + deep supervision, DDP, GLOBAL batch 256 @ 512x512 sharded 256/N (strong scaling)
                                                                                                configs[3]
  c5            U-Net++-R18 10-class inference, batch 64 per GPU @ 1024x1280, argmax + confusion matrix
                                                                                                configs[4]

One JSON line on rank 0 (see DESIGN.md "Measurement" for every key).  `value` is measured with the
batch already resident in HBM; `e2e` through the public API with pinned host batches copied in and
the loss / confusion matrix read back every step; `roofline` is the dominant kernel family (the tcgen05
implicit-GEMM convolutions: every fprop, dgrad and wgrad launch of the step) timed with CUDA events on the
launching stream, with the per-pass split and the aggregate over SURVEY 8(d)'s >= 60 % target set (3x3
stride-1 layers with Cout >= 32).  The default run appends short lines for the other configs under
`other_configs` (`--no-extra` skips them).  `--impl reference` times the reference's CPU path (the oracle
restatement in fp32 PyTorch eager + torch.optim.Adam, all host threads) on a bounded sample of the same
workload.
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

CONFIGS = {
    "c2": dict(kind="train", model="unetpp", classes=2, batch=16, H=512, W=512, ds=False, scaling="weak",
               metric=METRIC, cpu_batch=2,
               workload="BASELINE configs[1]: U-Net++ (resnet18 encoder, random init) binary train step fwd + "
                        "0.5*Dice+0.5*CE + bwd + Adam(lr 1e-3, wd 1e-5), batch %(batch)d per GPU @ %(H)dx%(W)d"),
    "c3": dict(kind="train", model="resnet_unet34", classes=10, batch=32, H=512, W=512, ds=False, scaling="weak",
               metric="train images/sec (512x512, ResNetUNet-34)", cpu_batch=2,
               workload="BASELINE configs[2]: in-tree ResNetUNet(n_class=10, resnet_model=34) (random-init encoder) "
                        "multiclass train step fwd + 0.5*Dice+0.5*CE + bwd + Adam, batch %(batch)d per GPU @ %(H)dx%(W)d"),
    "c4": dict(kind="train", model="unetpp", classes=2, global_batch=256, H=512, W=512, ds=True, scaling="strong",
               metric="train images/sec (512x512, U-Net++, deep supervision, global batch 256)", cpu_batch=2,
               workload="BASELINE configs[3]: U-Net++ (resnet18) with deep supervision (3 auxiliary heads, loss = mean "
                        "over 4 outputs), batch-sharded DDP, GLOBAL batch 256 @ %(H)dx%(W)d = %(batch)d per GPU"),
    "c5": dict(kind="infer", model="unetpp", classes=10, batch=64, H=1024, W=1280, ds=False, scaling="weak",
               metric="inference images/sec (1024x1280, U-Net++, argmax + confusion matrix)", cpu_batch=1,
               workload="BASELINE configs[4]: U-Net++ (resnet18) 10-class inference (BN folded) + argmax + "
                        "confusion-matrix metric, batch %(batch)d per GPU @ %(H)dx%(W)d"),
}


def resolve(name, world):
    cfg = dict(CONFIGS[name], name=name)
    if "global_batch" in cfg:
        if cfg["global_batch"] % world:
            raise SystemExit("config %s: global batch %d is not divisible by %d ranks" % (name, cfg["global_batch"], world))
        cfg["batch"] = cfg["global_batch"] // world
    cfg["workload"] = cfg["workload"] % cfg
    return cfg


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


def synthetic(cfg, n, gen_seed=6210, pinned=False):
    """frames rand in [0,1] -> ImageNet mean/std normalisation (SU/ModelTraining.py:300-301);
    labels randint(0, C); seed 6210 (the reference's, SU/ModelTraining.py:150)."""
    H, W = cfg["H"], cfg["W"]
    g = torch.Generator().manual_seed(gen_seed)
    u = torch.rand((n, 3, H, W), generator=g)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    x = ((u - mean) / std).contiguous()
    y = torch.randint(0, cfg["classes"], (n, H, W), generator=g)
    if pinned:
        x, y = x.pin_memory(), y.pin_memory()
    return x, y


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_rate(cfg, steps, warmup):
    """The reference's CPU path for the config: fp32 eager PyTorch (oracle restatement of smp.UnetPlusPlus, or
    the oracle restatement of the in-tree ResNetUNet) + the reference's loss / metric + torch.optim.Adam, all
    host threads, on a bounded sample (cfg['cpu_batch'] images per step) of the workload."""
    from oracle.losses import mixed_loss
    from oracle import metrics as OM
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(6210)
    batch = cfg["cpu_batch"]
    if cfg["model"] == "resnet_unet34":
        from oracle.resnet_unet import ResNetUNet
        model = ResNetUNet(cfg["classes"], 34)
    elif cfg["ds"]:
        from oracle.unetpp import DeepSupervisionUnetPlusPlus
        model = DeepSupervisionUnetPlusPlus("resnet18", None, 3, cfg["classes"])
    else:
        from oracle.unetpp import UnetPlusPlus
        model = UnetPlusPlus("resnet18", None, 3, cfg["classes"])
    x, y = synthetic(cfg, batch)
    times = []
    if cfg["kind"] == "train":
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            for p in model.parameters():
                p.grad = None
            out = model(x)
            outs = out if isinstance(out, (list, tuple)) else [out]
            loss = sum(mixed_loss(o, y, 0.5) for o in outs) / len(outs)
            loss.backward()
            opt.step()
            float(loss)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    else:
        model.eval()
        with torch.no_grad():
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                logits = model(x).numpy()
                OM.confusion_matrix(OM.argmax_first(logits), y.numpy(), cfg["classes"])
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
    total = sum(times)
    return batch * len(times) / total, total / len(times), torch.get_num_threads()


def cpu_sample_text(cfg, steps, cores):
    what = "train step (fwd + loss + bwd + Adam)" if cfg["kind"] == "train" else "eval forward + argmax + confusion matrix"
    return "%d timed steps of batch %d @ %dx%d (of the batch-%d workload): %s, fp32 eager PyTorch oracle, %d threads" % (
        steps, cfg["cpu_batch"], cfg["H"], cfg["W"], cfg["batch"], what, cores)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = resolve(args.config, max(1, args.gpus))
    steps, warmup = max(1, min(args.steps, 6)), max(1, min(args.warmup, 2))
    rate, sec, cores = cpu_rate(cfg, steps, warmup)
    sample = cpu_sample_text(cfg, steps, cores)
    line = {"impl": "reference", "metric": cfg["metric"], "value": rate, "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "name": cfg["name"], "global_batch": cfg["batch"] * args.gpus,
                       "classes": cfg["classes"], "parallelism": "cpu, %d threads" % cores,
                       "sample_batch_per_step": cfg["cpu_batch"]},
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
def kernels_per_call(lib, fn):
    two = (lib.mmr_wgrad_plan_run, lib.mmr_wgrad_halo_plan_run, lib.mmr_head_grad_prep, lib.mmr_pointwise_head_bwd)
    if not hasattr(fn, "restype"):      # scheduling marker of the engine's launch lists, not a kernel
        return 0
    return 2 if any(fn is f for f in two) else 1


def conv_kernel_times(eng, n_iter=8):
    """CUDA-event time of every tcgen05 conv launch of one step (fprop, dgrad and -- training -- wgrad: the
    halo kernels for the 3x3 stride-1 layers and the stem, the first-generation kernels for the strided / 1x1
    ones), on the launching stream, replaying the step's launch lists serially.  Returns {plan handle: ms}: the
    median over n_iter - 1 passes (the SM clock moves under the power cap from pass to pass)."""
    import ctypes as C
    lib = eng.lib
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    timed = (lib.mmr_conv_plan_run, lib.mmr_halo_conv_plan_run, lib.mmr_wgrad_plan_run, lib.mmr_wgrad_halo_plan_run)
    lists = [eng.repack_calls, eng.fwd_calls] + ([eng.bwd_calls[False]] if eng.training else [])
    acc = {}
    for it in range(n_iter):
        evs = []
        for calls in lists:
            for fn, a in calls:
                if not hasattr(fn, "restype"):
                    continue
                if any(fn is t for t in timed):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    fn(*a, sp)
                    e1.record(stream)
                    evs.append((getattr(a[0], "value", a[0]), e0, e1))
                else:
                    fn(*a, sp)
        torch.cuda.synchronize()
        if it > 0:
            for h, e0, e1 in evs:
                acc.setdefault(h, []).append(e0.elapsed_time(e1))
    return {h: sorted(v)[len(v) // 2] for h, v in acc.items()}


def conv_roofline(eng, pk):
    """Per-pass and target-set aggregates of the conv launches (algorithmic FLOPs = 2 x dense MACs)."""
    ms_of = conv_kernel_times(eng)
    fused = getattr(eng, "fused_dgrad_handles", set())
    rows = []
    for u in eng.units:
        if "fplan" not in u:
            continue
        op = u.get("op", {})
        target = u["kind"] == "conv" and u.get("k") == 3 and u.get("s") == 1 and u["cout"] >= 32
        for kind, key in (("fprop", "fplan"), ("dgrad", "dplan"), ("wgrad", "wplan")):
            plan = u.get(key)
            if plan is None:
                continue
            h = getattr(plan.handle, "value", plan.handle)
            if h not in ms_of:
                continue
            rows.append({"layer": op.get("conv", "?"), "pass": kind, "flops": plan.flops, "ms": ms_of[h],
                         "target": target, "fused_bn_bwd": kind == "dgrad" and h in fused})

    def agg(sel):
        fl, ms = sum(r["flops"] for r in sel), sum(r["ms"] for r in sel)
        tf = fl / (ms * 1e-3) / 1e12 if ms > 0 else None
        return {"launches": len(sel), "tflop": fl / 1e12, "ms": ms, "achieved": tf,
                "frac": tf / pk["tflops"] if tf else None}

    total = agg(rows)
    out = {"bound": "tensor",
           "kernel": "conv_halo_kernel / conv_wgrad_halo_kernel (+ wgrad_halo_reduce_kernel) / conv_gemm_tc_kernel / "
                     "conv_wgrad_tc_kernel: every fprop, dgrad and wgrad launch of a step; the epilogues also take the "
                     "BatchNorm batch statistics (fprop) and the BatchNorm-backward sums of single-reader units (dgrad), "
                     "which is not counted as work",
           "achieved": total["achieved"], "peak": pk["tflops"], "unit": "TFLOP/s", "frac": total["frac"],
           "launches_per_step": total["launches"], "kernel_ms_per_step": total["ms"],
           "algorithmic_tflop_per_step": total["tflop"], "peak_source": pk["src"],
           "by_pass": {k: agg([r for r in rows if r["pass"] == k]) for k in ("fprop", "dgrad", "wgrad")
                       if any(r["pass"] == k for r in rows)},
           "target_set": dict(agg([r for r in rows if r["target"]]),
                              note="SURVEY 8(d) >= 60 % target set: all 3x3 stride-1 layers with Cout >= 32, "
                                   "fprop + dgrad + wgrad"),
           "target_set_by_pass": {k: agg([r for r in rows if r["target"] and r["pass"] == k])
                                  for k in ("fprop", "dgrad", "wgrad") if any(r["pass"] == k for r in rows)}}
    return out, rows


def build_model(cfg, dev):
    from mmrseg_b200.models import ResNetUNet, UnetPlusPlus
    torch.manual_seed(6210)
    if cfg["model"] == "resnet_unet34":
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")       # BASELINE config 3: random-init encoder
            model = ResNetUNet(cfg["classes"], 34)
    else:
        model = UnetPlusPlus("resnet18", classes=cfg["classes"], deep_supervision=cfg["ds"])
    return model.to(dev)


def measure(cfg, args, world, rank, local, dev, full=True):
    """Runs one config; returns the JSON-able line (rank 0) or None.  full=False: the short form used for
    `other_configs` (device-resident value + conv aggregate only)."""
    import torch.distributed as dist
    from mmrseg_b200.data import DevicePrefetcher, ResultReader
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    from mmrseg_b200.metrics import confusion_matrix
    from mmrseg_b200.optim import FusedAdam
    from mmrseg_b200.parallel import DistributedDataParallel

    n, H, W, classes = cfg["batch"], cfg["H"], cfg["W"], cfg["classes"]
    train = cfg["kind"] == "train"
    torch.cuda.reset_peak_memory_stats(dev)
    model = build_model(cfg, dev)
    model.train() if train else model.eval()
    crit = DiceCrossEntropyLoss(0.5)
    opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-5) if train else None
    model._ensure_flat(dev)
    ddp = DistributedDataParallel(model) if (world > 1 and train) else None      # broadcasts rank 0's weights
    xh, yh = synthetic(cfg, n, 6210 + rank, pinned=True)
    xd, yd = xh.to(dev), yh.to(dev)
    cm = torch.zeros((n, classes, classes), device=dev, dtype=torch.int64)

    def loss_of(out, y):
        if isinstance(out, list):       # deep supervision: mean over the four outputs (oracle/unetpp.py)
            return sum(crit(o, y) for o in out) / len(out)
        return crit(out, y)

    def step_on(x, y):
        if train:
            for p in model.parameters():
                p.grad = None
            loss = loss_of(model(x), y)
            loss.backward()
            opt.step()
            return loss
        # eval step: argmax + confusion matrix in the head's epilogue (model.segment), no logits in HBM
        pred, c = model.segment(x, y)
        cm.add_(c)
        return cm

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

    steps = args.steps if full else max(3, min(args.steps, 5))
    warm = max(args.warmup, 3)
    step_on(xd, yd)                                   # builds the plan (tensor maps, buffers)
    for _ in range(warm):
        step_on(xd, yd)
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(lambda: step_on(xd, yd), steps)
    clocks = sampler.result()
    eng = model._engine_for(xd, training=train)
    step_flops = eng.conv_flops_fwd + (eng.conv_flops_bwd if train else 0)
    pk = peaks()
    line = {"metric": cfg["metric"], "value": world * n * steps / (ms * 1e-3), "unit": "images/s", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": cfg["workload"], "name": cfg["name"], "global_batch": world * n,
                       "classes": classes, "parallelism": "dp%d" % world,
                       "conv_tflop_per_step": step_flops / 1e12},
            "clocks": clocks,
            "step_tensor_frac": step_flops / (ms / steps * 1e-3) / 1e12 / pk["tflops"]}
    if not full:
        if rank == 0:
            roof, _ = conv_roofline(eng, pk)
            line["roofline"] = {k: roof[k] for k in ("bound", "achieved", "peak", "unit", "frac", "by_pass", "target_set")}
            line["mem_gb"] = torch.cuda.max_memory_allocated(dev) / 1e9
        if world > 1:
            dist.barrier()
        del model, opt, eng, ddp
        torch.cuda.empty_cache()
        return line if rank == 0 else None

    # ---- end to end through the public API: pinned host batches -> DevicePrefetcher -> step -> result read back
    feed = DevicePrefetcher(None, dev)

    reader = ResultReader(lag=1)

    def run_e2e(host_x, k):
        # every step's result (the loss / the confusion matrix) is read back to the host: an asynchronous copy into
        # pinned memory behind the step's kernels, collected one step later (ResultReader), so the read does not
        # drain the GPU queue the way `loss.item()` after every step does
        got = []
        feed.loader = [(host_x, yh)] * k
        for x, y in feed:
            r = step_on(x, y)
            got += reader.push(r if train else r.sum())
        got += reader.flush()
        assert len(got) == k
        return float(got[-1])

    def timed_loop(host_x):
        run_e2e(host_x, 3)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(host_x, steps)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    ms_e2e = timed_loop(xh)
    # the same step fed with uint8 HWC frames (what the loaders hold): /255 and utils.normalize on the device
    fh = (torch.rand((n, H, W, 3), generator=torch.Generator().manual_seed(6210 + rank)) * 255).to(torch.uint8).pin_memory()
    model.set_input_normalization((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    ms_e2e_u8 = timed_loop(fh)
    d2h = 4 if train else 8
    line["e2e"] = {"value": world * n * steps / (ms_e2e * 1e-3), "unit": "images/s",
                   "h2d_bytes_per_step": xh.numel() * 4 + yh.numel() * 8, "d2h_bytes_per_step": d2h}
    line["e2e_uint8_frames"] = {"value": world * n * steps / (ms_e2e_u8 * 1e-3), "unit": "images/s",
                                "h2d_bytes_per_step": fh.numel() + yh.numel() * 8, "d2h_bytes_per_step": d2h,
                                "note": "same step through model(frames_u8): ToTensor + utils.normalize fused into the stem loader"}
    if rank == 0:
        roof, rows = conv_roofline(eng, pk)
        lib = eng.lib
        lists = (eng.repack_calls, eng.fwd_calls) + ((eng.bwd_calls[False],) if train else ())
        per_step = sum(kernels_per_call(lib, fn) for calls in lists for fn, _ in calls)
        per_step += (3 * (4 if cfg["ds"] else 1) + 1) if train else 1   # loss fwd (2) + bwd per output, Adam | zeroing of the confusion counters (the metric itself is the head's epilogue)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")
        if os.path.exists(tpath) and cfg["name"] == "c2":
            tj = json.load(open(tpath))
            # DRAM bytes of every conv kernel of one step (ncu) over this line's own launch count, like `achieved`
            traffic = (tj["dram_read_bytes_per_step"] + tj["dram_write_bytes_per_step"]) / max(roof["launches_per_step"], 1)
        roof["traffic"] = traffic
        line["roofline"] = roof
        line["gpu_launches"] = per_step * steps
        acts = sum(a.buf.numel() * a.buf.element_size() for a in eng.acts.values() if a.buf is not None)
        line["config"]["l2"] = "per-step working set %.1f GB >> 126 MB L2 (no flush needed)" % (
            (getattr(eng, "arena_bytes", 0) + acts) / 1e9)
        if os.environ.get("MMR_BENCH_LAYERS"):
            with open(os.environ["MMR_BENCH_LAYERS"], "w") as fh_:
                fh_.write("layer,pass,gflop,ms,tflops,target_set,fused_bn_bwd\n")
                for r in rows:
                    fh_.write("%s,%s,%.3f,%.4f,%.1f,%d,%d\n" % (r["layer"], r["pass"], r["flops"] / 1e9, r["ms"],
                                                              r["flops"] / r["ms"] / 1e9, r["target"], r["fused_bn_bwd"]))
    if world > 1:
        dist.barrier()
    del model, opt, eng, ddp, feed
    torch.cuda.empty_cache()
    return line if rank == 0 else None


def run_ours(args):
    import torch.distributed as dist
    from mmrseg_b200 import _lib

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
    cfg = resolve(args.config, world)
    line = measure(cfg, args, world, rank, local, dev, full=True)
    others = {}
    if not args.no_extra and args.config == "c2":
        # the other BASELINE configs, short form, in the same process (config 4 is the DDP one: it is what the
        # driver's 1/2/4/8-GPU scaling run reads its deep-supervision / global-batch-256 numbers from)
        names = ["c4"] if world > 1 else ["c3", "c4", "c5"]
        for name in names:
            try:
                o = measure(resolve(name, world), args, world, rank, local, dev, full=False)
            except Exception as e:      # never lose the headline line to an auxiliary config
                o = {"error": "%s: %s" % (type(e).__name__, e)}
                torch.cuda.empty_cache()
            if rank == 0:
                others[name] = o
    if rank == 0:
        if others:
            line["other_configs"] = others
        if not args.no_cpu_baseline and world == 1:
            rate, sec, cores = cpu_rate(cfg, 2, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": cpu_sample_text(cfg, 2, cores)}
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
    ap.add_argument("--config", default=os.environ.get("MMR_BENCH_CONFIG", "c2"), choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short other_configs lines")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
