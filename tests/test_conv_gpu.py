"""tcgen05 implicit-GEMM convolution (csrc/conv_gemm.cu) against torch.nn.functional.conv2d.

The comparison runs F.conv2d in fp32 on the same bf16-rounded operands, so the only
difference left is accumulation order: tolerance 2e-3 relative to the output RMS
(bf16 output rounding is 2^-9 = 2e-3 relative per element).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mk(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device="cuda") * scale).to(torch.bfloat16)


def _weights_fwd(w):  # OIHW fp32 -> [O][taps*I] bf16, column = (ky*k+kx)*I + ci
    O, I, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(O, kh * kw * I).contiguous().to(torch.bfloat16)


def _ref_conv(sources, w, stride, pad):
    xs = []
    for t, up in sources:
        x = t.float().permute(0, 3, 1, 2)
        if up == 2:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        xs.append(x)
    x = torch.cat(xs, 1)
    return F.conv2d(x, w.to(torch.bfloat16).float(), None, stride, pad)


def _check(out_nhwc, ref_nchw, tol=4e-3):
    got = out_nhwc.float().permute(0, 3, 1, 2)
    err = (got - ref_nchw).abs().max().item()
    rms = ref_nchw.pow(2).mean().sqrt().item()
    assert err <= tol * max(rms, 1e-6) * 8, (err, rms)


CASES = [
    # (N, H, W, [(C, up)], Cout, k, stride)
    (2, 16, 16, [(64, 1)], 64, 3, 1),
    (2, 16, 16, [(64, 1)], 128, 3, 1),
    (1, 32, 32, [(32, 1)], 32, 3, 1),
    (1, 32, 32, [(16, 1)], 16, 3, 1),
    (2, 16, 16, [(64, 1)], 128, 3, 2),
    (2, 16, 16, [(64, 1)], 128, 1, 2),
    (2, 16, 16, [(128, 2), (64, 1)], 64, 3, 1),
    (1, 32, 32, [(64, 2), (64, 1), (64, 1)], 64, 3, 1),
    (3, 24, 40, [(64, 1)], 64, 3, 1),       # ragged tiles
    (1, 8, 8, [(512, 1)], 512, 3, 1),
    (1, 64, 64, [(32, 2)], 16, 3, 1),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("impl", [1, 0])
def test_fprop_matches_conv2d(case, impl):
    from mmrseg_b200 import convplan
    N, H, W, srcs, Cout, k, stride = case
    gen = torch.Generator(device="cuda").manual_seed(6210)
    sources = []
    for C, up in srcs:
        sources.append((_mk((N, H // up, W // up, C), gen), up))
    cin = sum(c for c, _ in srcs)
    w = torch.randn((Cout, cin, k, k), generator=gen, device="cuda") * (1.0 / (cin * k * k) ** 0.5)
    pad = k // 2
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    out = torch.full((N, Ho, Wo, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    plan = convplan.build_fprop(sources, _weights_fwd(w), k, stride, pad, out)
    plan.run(impl=impl)
    torch.cuda.synchronize()
    ref = _ref_conv(sources, w, stride, pad)
    _check(out, ref)


def test_fprop_epilogue_scale_bias_residual_relu():
    from mmrseg_b200 import convplan
    gen = torch.Generator(device="cuda").manual_seed(1)
    N, H, W, C, Cout = 2, 16, 16, 64, 64
    x = _mk((N, H, W, C), gen)
    w = torch.randn((Cout, C, 3, 3), generator=gen, device="cuda") / 24
    scale = torch.rand(Cout, generator=gen, device="cuda") + 0.5
    bias = torch.randn(Cout, generator=gen, device="cuda")
    res = _mk((N, H, W, Cout), gen)
    out = torch.empty((N, H, W, Cout), device="cuda", dtype=torch.bfloat16)
    plan = convplan.build_fprop([(x, 1)], _weights_fwd(w), 3, 1, 1, out, scale=scale, bias=bias,
                                residual=res, relu=True)
    plan.run()
    torch.cuda.synchronize()
    ref = _ref_conv([(x, 1)], w, 1, 1) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    ref = torch.relu(ref + res.float().permute(0, 3, 1, 2))
    _check(out, ref)


def test_head_f32_nchw_output():
    from mmrseg_b200 import convplan
    from mmrseg_b200._lib import MMR_OUT_F32_NCHW
    gen = torch.Generator(device="cuda").manual_seed(2)
    N, H, W, C, classes = 2, 32, 32, 16, 2
    x = _mk((N, H, W, C), gen)
    w = torch.randn((classes, C, 3, 3), generator=gen, device="cuda") / 12
    b = torch.randn(classes, generator=gen, device="cuda")
    wf = torch.zeros((16, 9 * C), device="cuda", dtype=torch.bfloat16)
    wf[:classes] = _weights_fwd(w)
    out = torch.full((N, classes, H, W), float("nan"), device="cuda")
    plan = convplan.build_fprop([(x, 1)], wf, 3, 1, 1, out, bias=b, out_mode=MMR_OUT_F32_NCHW,
                                cout=classes, bn=16)
    plan.run()
    torch.cuda.synchronize()
    ref = _ref_conv([(x, 1)], w, 1, 1) + b.view(1, -1, 1, 1)
    err = (out - ref).abs().max().item()
    assert err < 1e-3, err
