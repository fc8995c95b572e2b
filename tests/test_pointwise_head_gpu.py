"""The CUDA-core 1x1 head kernels (csrc/pointwise_head.cu: `ResNetUNet.conv_last`, SU/UArchModel/resnet_unet.py:204,298)
against fp32 PyTorch on the same bf16 activation and fp32 weights.

Tolerances: logits and the fp32 gradients (head weight / bias, producer bias) 1e-5 relative Frobenius (fp32 sums
in a different order), the bf16 data gradient exactly the bf16 rounding of the fp32 reference up to one-ulp flips
(<= 4e-3 relative, as for every bf16 tensor the kernels store)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("n,h,w,cout", [(2, 16, 24, 10), (1, 7, 9, 2), (3, 32, 32, 16), (2, 20, 12, 5), (1, 64, 64, 3)])
@pytest.mark.parametrize("relu_mask", [0, 1])
def test_pointwise_head_forward_backward(n, h, w, cout, relu_mask):
    from mmrseg_b200 import _lib as L
    lib = L.lib()
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    gen = torch.Generator(device="cuda").manual_seed(1000 * n + 10 * cout + relu_mask)
    x = torch.randn((n, h, w, 64), generator=gen, device="cuda").to(torch.bfloat16)
    wt = torch.randn((cout, 64, 1, 1), generator=gen, device="cuda") * 0.2
    b = torch.randn((cout,), generator=gen, device="cuda")
    logits = torch.full((n, cout, h, w), 9.0, device="cuda")
    L.check(lib.mmr_pointwise_head_fwd(_p(x), _p(wt), _p(b), n, h, w, 64, cout, _p(logits), s))
    # the reference in float64 on the CPU (cuDNN's fp32 conv may run in TF32)
    xr = x.double().cpu().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wr = wt.double().cpu().requires_grad_(True)
    ref = F.conv2d(xr, wr, b.double().cpu())
    assert _rel(logits.cpu(), ref.detach()) <= 1e-5

    dl = torch.randn((n, cout, h, w), generator=gen, device="cuda")
    ref.backward(dl.double().cpu())
    want_dx = xr.grad * (xr.detach() > 0).float() if relu_mask else xr.grad
    dx = torch.full((n, h, w, 64), 5.0, device="cuda", dtype=torch.bfloat16)
    dw = torch.full((cout, 64, 1, 1), 2.0, device="cuda")
    db = torch.full((cout,), 2.0, device="cuda")
    dbl = torch.full((64,), 2.0, device="cuda")
    ws = torch.empty(int(lib.mmr_pointwise_head_bwd_workspace_bytes(cout)) // 4, device="cuda")
    for accumulate in (0, 1):
        L.check(lib.mmr_pointwise_head_bwd(_p(dl), _p(x), _p(wt), n, h, w, 64, cout, relu_mask, _p(dx), _p(dw), _p(db),
                                           _p(dbl), accumulate, _p(ws), s))
        torch.cuda.synchronize()
        k = 1 + accumulate
        got_dx = dx.float().cpu().permute(0, 3, 1, 2)
        assert _rel(got_dx, want_dx.float().to(torch.bfloat16).float()) <= 4e-3
        assert _rel(dw.cpu(), k * wr.grad) <= 1e-5
        assert _rel(db.cpu(), k * dl.double().cpu().sum((0, 2, 3))) <= 1e-5
        assert _rel(dbl.cpu(), k * got_dx.double().sum((0, 2, 3))) <= 1e-5      # column sums of the STORED gradient
    # no producer bias: the pointer may be NULL
    L.check(lib.mmr_pointwise_head_bwd(_p(dl), _p(x), _p(wt), n, h, w, 64, cout, relu_mask, _p(dx), _p(dw), _p(db), None, 0,
                                       _p(ws), s))
    torch.cuda.synchronize()
    assert _rel(dw.cpu(), wr.grad) <= 1e-5


def test_pointwise_head_rejects_other_shapes():
    from mmrseg_b200 import _lib as L
    lib = L.lib()
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    x = torch.zeros((1, 4, 4, 32), device="cuda", dtype=torch.bfloat16)
    wt = torch.zeros((2, 32), device="cuda")
    out = torch.zeros((1, 2, 4, 4), device="cuda")
    assert lib.mmr_pointwise_head_fwd(_p(x), _p(wt), None, 1, 4, 4, 32, 2, _p(out), s) != 0
    assert lib.mmr_pointwise_head_fwd(_p(x), _p(wt), None, 1, 4, 4, 64, 17, _p(out), s) != 0
