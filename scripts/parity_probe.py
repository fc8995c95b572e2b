"""Measure (not assert) whole-model parity of the engine against (a) the fp32 oracle with bf16 rounding at the
engine's storage points (oracle/bf16_points.py) and (b) the plain fp32 oracle, at several sizes up to BASELINE's
512x512.  Prints one line per case; the numbers feed the tolerances written in tests/test_parity_gpu.py and the
table in DESIGN.md.  Test infrastructure: runs the oracle on the host CPU."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import model_pair, rel, synthetic_batch  # noqa: E402


def run(encoder, classes, n, h, w, matched=True, fp32=True, randomize_bn=True):
    from oracle import bf16_points
    from oracle.losses import mixed_loss
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    torch.set_num_threads(os.cpu_count())
    ref, net = model_pair(classes, encoder, randomize_bn=randomize_bn)
    x, y = synthetic_batch(n, classes, h, w)
    ref.train()
    net.train()
    got = net(x.cuda())
    loss = DiceCrossEntropyLoss(0.5)(got, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().cpu().clone() for k, p in net.named_parameters()}
    out = {}
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    for name, fwd in (("bf16pts", lambda: bf16_points.unetpp_forward(ref, x)), ("fp32", lambda: ref(x))):
        if name == "fp32" and not fp32:
            continue
        ref.load_state_dict(sd0)
        ref.zero_grad(set_to_none=True)
        t0 = time.time()
        want = fwd()
        l = mixed_loss(want, y, 0.5)
        l.backward()
        dt = time.time() - t0
        rp = dict(ref.named_parameters())
        errs = sorted(((rel(grads[k], rp[k].grad), k) for k in grads), reverse=True)
        cos = min(torch.nn.functional.cosine_similarity(grads[k].flatten(), rp[k].grad.flatten(), dim=0).item()
                  for k in grads)
        med = errs[len(errs) // 2][0]
        print("%-8s %s C=%d %dx%dx%d: logits %.2e loss %.2e | grad worst %.2e (%s) 2nd %.2e median %.2e min-cos %.5f"
              " | oracle %.1fs" % (name, encoder, classes, n, h, w, rel(got.detach().cpu(), want.detach()),
                                   abs(loss.item() - l.item()) / abs(l.item()), errs[0][0], errs[0][1], errs[1][0],
                                   med, cos, dt), flush=True)
        out[name] = (rel(got.detach().cpu(), want.detach()), errs[0][0], med)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    a = ap.parse_args()
    cases = [("resnet18", 2, 4, 64, 64), ("resnet18", 10, 2, 128, 128), ("resnet34", 10, 4, 64, 64),
             ("resnet18", 2, 2, 256, 256)]
    if a.full:
        cases += [("resnet18", 2, 2, 512, 512), ("resnet18", 2, 4, 512, 512)]
    for c in cases:
        run(*c)
