"""Data- and weight-gradient kernels (tcgen05) against torch autograd of F.conv2d in fp32 on the
same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mk(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device="cuda") * scale).to(torch.bfloat16)


def _w_dgrad(w):  # OIHW -> [I][taps*O] bf16, column = (ky*k+kx)*O + co
    O, I, kh, kw = w.shape
    return w.permute(1, 2, 3, 0).reshape(I, kh * kw * O).contiguous().to(torch.bfloat16)


def _autograd(sources, w, dz, stride, pad):
    leaves, xs = [], []
    for t, up in sources:
        x = t.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
        leaves.append(x)
        xs.append(F.interpolate(x, scale_factor=2, mode="nearest") if up == 2 else x)
    full = torch.cat(xs, 1)
    full.retain_grad()
    wl = w.to(torch.bfloat16).float().requires_grad_(True)
    y = F.conv2d(full, wl, None, stride, pad)
    y.backward(dz.float().permute(0, 3, 1, 2))
    return full.grad, wl.grad


def _close(got, ref, tol):
    err = (got - ref).abs().max().item()
    rms = ref.pow(2).mean().sqrt().item()
    assert err <= tol * max(rms, 1e-9), (err, rms)


DGRAD_CASES = [
    # N, H, W, [C...] (concat of same-res tensors), Cout, k, stride
    (2, 16, 16, [64], 64, 3, 1),
    (2, 16, 16, [128, 64], 64, 3, 1),
    (1, 32, 32, [32], 32, 3, 1),
    (2, 16, 16, [64], 128, 3, 2),
    (2, 16, 16, [64], 128, 1, 2),
    (1, 32, 32, [16], 16, 3, 1),
    (3, 24, 40, [64, 64, 64], 64, 3, 1),
]


@pytest.mark.parametrize("case", DGRAD_CASES)
@pytest.mark.parametrize("impl", [1, 0])
def test_dgrad(case, impl):
    from mmrseg_b200 import convplan
    N, H, W, Cs, Cout, k, stride = case
    gen = torch.Generator(device="cuda").manual_seed(6210)
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    sources = [(_mk((N, H, W, c), gen), 1) for c in Cs]
    w = torch.randn((Cout, sum(Cs), k, k), generator=gen, device="cuda") / (sum(Cs) * k * k) ** 0.5
    dz = _mk((N, Ho, Wo, Cout), gen)
    grads = [torch.full((N, H, W, c), float("nan"), device="cuda", dtype=torch.bfloat16) for c in Cs]
    plan = convplan.build_dgrad(dz, _w_dgrad(w), k, stride, pad, (H, W), grads)
    plan.run(impl=impl)
    torch.cuda.synchronize()
    gfull, _ = _autograd(sources, w, dz, stride, pad)
    got = torch.cat([g.float().permute(0, 3, 1, 2) for g in grads], 1)
    _close(got, gfull, 0.04)


WGRAD_CASES = [
    # N, H, W, [(C, up)], Cout, k, stride
    (2, 16, 16, [(64, 1)], 64, 3, 1),
    (2, 16, 16, [(64, 1)], 128, 3, 1),
    (2, 16, 16, [(128, 1)], 256, 3, 1),
    (1, 8, 8, [(512, 1)], 512, 3, 1),
    (1, 32, 32, [(32, 1)], 32, 3, 1),
    (1, 32, 32, [(16, 1)], 16, 3, 1),
    (2, 16, 16, [(64, 1)], 128, 3, 2),
    (2, 16, 16, [(64, 1)], 128, 1, 2),
    (2, 16, 16, [(128, 2), (64, 1)], 64, 3, 1),
    (1, 32, 32, [(64, 2), (64, 1), (64, 1)], 64, 3, 1),
    (3, 24, 40, [(64, 1)], 64, 3, 1),
    (1, 64, 64, [(32, 2)], 16, 3, 1),
]


@pytest.mark.parametrize("case", WGRAD_CASES)
@pytest.mark.parametrize("impl", [1, 0])
def test_wgrad(case, impl):
    from mmrseg_b200 import convplan
    N, H, W, srcs, Cout, k, stride = case
    gen = torch.Generator(device="cuda").manual_seed(6210)
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    sources = [(_mk((N, H // up, W // up, c), gen), up) for c, up in srcs]
    cin = sum(c for c, _ in srcs)
    w = torch.randn((Cout, cin, k, k), generator=gen, device="cuda")
    dz = _mk((N, Ho, Wo, Cout), gen)
    dst = torch.full((Cout, cin, k, k), float("nan"), device="cuda")
    plan = convplan.build_wgrad(dz, sources, k, stride, pad, dst)
    plan.run(impl=impl)
    torch.cuda.synchronize()
    _, gw = _autograd(sources, w, dz, stride, pad)
    _close(dst, gw, 2e-3)
    # accumulation doubles it
    plan.run(impl=impl, accumulate=True)
    torch.cuda.synchronize()
    _close(dst, 2 * gw, 2e-3)


def test_wgrad_split_k_deterministic():
    from mmrseg_b200 import convplan
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = _mk((4, 32, 32, 64), gen)
    dz = _mk((4, 32, 32, 64), gen)
    outs = []
    for _ in range(2):
        dst = torch.empty((64, 64, 3, 3), device="cuda")
        plan = convplan.build_wgrad(dz, [(x, 1)], 3, 1, 1, dst, n_split=16)
        plan.run()
        torch.cuda.synchronize()
        outs.append(dst.clone())
    assert torch.equal(outs[0], outs[1])
