"""Per-layer timing of the halo-tile conv kernel on the U-Net++ (resnet18) 3x3 s1 layer shapes at
batch 16 @ 512x512.  usage: python scripts/bench_halo.py [tag] [key=value ...forced config]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmrseg_b200 import convplan  # noqa: E402

N = int(os.environ.get("BATCH", "16"))
# name, H (=W), [(C, up)], Cout
LAYERS = [
    ("layer1.conv", 128, [(64, 1)], 64),
    ("layer2.conv", 64, [(128, 1)], 128),
    ("layer3.conv", 32, [(256, 1)], 256),
    ("layer4.conv", 16, [(512, 1)], 512),
    ("x_0_0.conv1", 32, [(512, 2), (256, 1)], 256),
    ("x_1_1.conv1", 64, [(256, 2), (128, 1)], 128),
    ("x_2_2.conv1", 128, [(128, 2), (64, 1)], 64),
    ("x_3_3.conv1", 256, [(64, 2), (64, 1)], 64),
    ("x_3_3.conv2", 256, [(64, 1)], 64),
    ("x_0_1.conv1", 64, [(256, 2), (128, 1), (128, 1)], 128),
    ("x_1_2.conv1", 128, [(128, 2), (64, 1), (64, 1)], 64),
    ("x_2_3.conv1", 256, [(64, 2), (64, 1), (64, 1)], 64),
    ("x_0_2.conv1", 128, [(128, 2), (64, 1), (64, 1), (64, 1)], 64),
    ("x_1_3.conv1", 256, [(64, 2), (64, 1), (64, 1), (64, 1)], 64),
    ("x_0_3.conv1", 256, [(64, 2), (64, 1), (64, 1), (64, 1), (64, 1)], 32),
    ("x_0_3.conv2", 256, [(32, 1)], 32),
    ("x_0_4.conv1", 512, [(32, 2)], 16),
    ("x_0_4.conv2", 512, [(16, 1)], 16),
    ("head10", 512, [(16, 1)], 10),      # segmentation head: fp32 NCHW logits (fprop only)
]


def timeit(plan, iters=10):
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        plan.run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 and "=" not in sys.argv[1] else "halo"
    force = {k: int(v) for k, v in (a.split("=") for a in sys.argv[1:] if "=" in a)} or None
    only = os.environ.get("ONLY")
    gen = torch.Generator(device="cuda").manual_seed(0)
    rows = []
    kinds = os.environ.get("KINDS", "fprop,dgrad,wgrad").split(",")
    tot = {"fprop": [0.0, 0.0], "dgrad": [0.0, 0.0], "wgrad": [0.0, 0.0]}
    for name, hw, srcs, cout in LAYERS:
        if only and only not in name:
            continue
        sources = [((torch.randn((N, hw // up, hw // up, c), generator=gen, device="cuda")).to(torch.bfloat16), up)
                   for c, up in srcs]
        cin = sum(c for c, _ in srcs)
        w = torch.randn((cout, cin, 3, 3), generator=gen, device="cuda") / (9 * cin) ** 0.5
        out = torch.empty((N, hw, hw, cout), device="cuda", dtype=torch.bfloat16)
        for kind in kinds:
            if name.startswith("head") and kind != "fprop":
                continue
            try:
                if kind == "wgrad":
                    cz = -(-cout // 16) * 16
                    dz = torch.randn((N, hw, hw, cz), generator=gen, device="cuda").to(torch.bfloat16)
                    dst = torch.empty((cout, cin, 3, 3), device="cuda")
                    plan = convplan.build_wgrad_halo(dz, sources, dst, force=force)
                elif kind == "fprop" and name.startswith("head"):
                    logits = torch.empty((N, cout, hw, hw), device="cuda")
                    plan = convplan.build_fprop_halo(sources, w, None, bias=torch.zeros(cout, device="cuda"),
                                                     out_f32=logits, force=force)
                elif kind == "fprop":
                    stats = None
                    if os.environ.get("STATS"):   # BatchNorm batch statistics in the epilogue, as the training step runs it
                        stats = torch.zeros((8 * 2 * cout,), device="cuda", dtype=torch.float64)
                    plan = convplan.build_fprop_halo(sources, w, out, force=force, stats=stats, stats_ld=cout)
                else:
                    cz = -(-cout // 16) * 16
                    dz = torch.randn((N, hw, hw, cz), generator=gen, device="cuda").to(torch.bfloat16)
                    grads = [torch.empty((N, hw, hw, c), device="cuda", dtype=torch.bfloat16) for c, _ in srcs]
                    plan = convplan.build_dgrad_halo(dz, w, grads, force=force)
            except Exception as e:  # forced configuration impossible for this layer
                print("%-14s %-5s skipped: %s" % (name, kind, str(e)[:80]))
                continue
            ms = timeit(plan)
            tf = plan.flops / ms / 1e9
            c = plan.cfg
            if kind == "wgrad":
                print("%-14s %-5s %7.3f ms %7.1f TFLOP/s  bn=%d tx=%d split=%d stages=%d" % (
                    name, kind, ms, tf, c["bn"], c["tx"], c["n_split"], c["stages"]), flush=True)
            else:
                print("%-14s %-5s %7.3f ms %7.1f TFLOP/s  bn=%d tx=%d rph=%d tps=%d acc=%d hs=%d ws=%d os=%d" % (
                    name, kind, ms, tf, c["bn"], c["tx"], c.get("rph", 1), c["tps"], c["acc_bufs"], c["halo_stages"],
                    c["w_slots"], c["out_stages"]), flush=True)
            rows.append((name, kind, plan.flops / 1e9, ms, tf))
            tot[kind][0] += plan.flops / 1e9
            tot[kind][1] += ms
            del plan
    for kind, (gf, ms) in tot.items():
        if ms:
            print("TOTAL %-5s %8.1f GFLOP %7.3f ms %7.1f TFLOP/s" % (kind, gf, ms, gf / ms))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "halo_layers_%s.csv" % tag), "w") as fh:
        fh.write("layer,kind,gflop,ms,tflops\n")
        for r in rows:
            fh.write("%s,%s,%.2f,%.4f,%.1f\n" % r)


if __name__ == "__main__":
    main()
