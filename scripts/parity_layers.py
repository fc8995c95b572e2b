"""Layer-by-layer comparison of the engine's stored tensors (conv outputs z, activations) with the fp32 oracle
run with bf16 rounding at the engine's storage points (oracle/bf16_points.py): where does the first
difference beyond summation order appear?  Diagnostic; test infrastructure."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import model_pair, rel, synthetic_batch  # noqa: E402
from oracle import bf16_points  # noqa: E402

n, hw = int(sys.argv[1]) if len(sys.argv) > 1 else 2, int(sys.argv[2]) if len(sys.argv) > 2 else 128
ref, net = model_pair(2)
x, y = synthetic_batch(n, 2, hw, hw)
ref.train()
net.train()
with torch.no_grad():
    got = net(x.cuda())
eng = list(net._engines.values())[0]
trace = {}
with torch.no_grad():
    want = bf16_points.unetpp_forward(ref, x, trace)
nchw = lambda t: t.float().permute(0, 3, 1, 2).cpu()
units = {u["out"].name: u for u in eng.units if "out" in u and hasattr(u["out"], "name")}
for name, (z, a) in trace.items():
    u = units[name]
    ez, ea = nchw(u["z"]), nchw(u["out"].buf)
    dz, da = (ez - z), (ea - a)
    ulp_frac = ((dz.abs() > 0).float().mean().item())
    print("%-28s z rel %.2e (differing %.4f, max %.3g) | act rel %.2e | mean %.2e invstd %.2e" % (
        name, rel(ez, z), ulp_frac, dz.abs().max().item(), rel(ea, a),
        rel(u["mean"].cpu(), z.mean((0, 2, 3))), rel(u["invstd"].cpu(), (z.var((0, 2, 3), unbiased=False) + 1e-5).rsqrt())))
print("logits rel %.2e" % rel(got.cpu(), want))
