"""Halo-tile tcgen05 convolution (csrc/conv_halo.cu) against torch.nn.functional.conv2d.

Same comparison as tests/test_conv_gpu.py: F.conv2d in fp32 on the same bf16-rounded operands, so
what is left is accumulation order and the bf16 rounding of the stored output: tolerance 4e-3 x 8 of
the output RMS for forward / data-gradient tensors.  BatchNorm statistics taken in the epilogue are
compared with the sums of the stored bf16 tensor in float64: 1e-5 relative (fp32 partial sums over at
most a few thousand values, then double atomics).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mk(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device="cuda") * scale).to(torch.bfloat16)


def _ref_conv(sources, w):
    xs = []
    for t, up in sources:
        x = t.float().permute(0, 3, 1, 2)
        if up == 2:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        xs.append(x)
    return F.conv2d(torch.cat(xs, 1), w.to(torch.bfloat16).float(), None, 1, 1)


def _check(out_nhwc, ref_nchw, tol=4e-3):
    got = out_nhwc.float().permute(0, 3, 1, 2)
    assert torch.isfinite(got).all()
    err = (got - ref_nchw).abs().max().item()
    rms = ref_nchw.pow(2).mean().sqrt().item()
    assert err <= tol * max(rms, 1e-6) * 8, (err, rms)


FPROP_CASES = [
    # (N, H, W, [(C, up)], Cout, force)
    (2, 16, 16, [(64, 1)], 64, None),
    (2, 32, 32, [(64, 1)], 64, dict(tx=4)),
    (2, 32, 32, [(64, 1)], 64, dict(tx=2, acc_bufs=1)),
    (2, 32, 32, [(64, 1)], 64, dict(tx=1, out_stages=1)),
    (2, 32, 32, [(64, 1)], 64, dict(tx=2, tps=3)),
    (2, 16, 16, [(64, 1)], 128, None),
    (1, 32, 32, [(128, 1)], 256, None),
    (1, 16, 16, [(512, 1)], 512, None),
    (1, 32, 32, [(32, 1)], 32, None),
    (1, 32, 32, [(16, 1)], 16, None),
    (1, 32, 64, [(16, 1)], 16, dict(tx=4)),
    (2, 32, 32, [(128, 2), (64, 1)], 64, None),
    (1, 32, 32, [(64, 2), (64, 1), (64, 1)], 64, dict(tx=4)),
    (1, 32, 32, [(64, 2), (64, 1), (64, 1)], 64, dict(tx=1)),
    (2, 64, 64, [(64, 2), (64, 1), (64, 1), (64, 1), (64, 1)], 32, None),
    (3, 24, 40, [(64, 1)], 64, None),        # ragged tiles in both directions
    (1, 40, 24, [(64, 1)], 64, dict(tx=2)),
    (2, 8, 8, [(512, 1)], 512, None),        # image smaller than one tile
    (1, 64, 64, [(32, 2)], 16, None),
    (1, 32, 48, [(64, 2)], 64, dict(tx=4)),  # upsampled source, ragged in x
    # row-phase stacking: one MMA of N = 2-3 x Cout serves vertically adjacent output rows
    (2, 64, 32, [(64, 1)], 64, dict(rph=2, tx=2)),
    (2, 64, 32, [(64, 1)], 64, dict(rph=2, tx=1, acc_bufs=1)),
    (1, 64, 32, [(64, 1)], 64, dict(rph=4, tx=1)),
    (1, 64, 32, [(32, 1)], 32, dict(rph=4, tx=2, acc_bufs=1)),
    (1, 64, 64, [(128, 2), (64, 1)], 64, dict(rph=2, tx=2)),       # nearest-x2 source: 8*rph low-res body rows
    (1, 64, 32, [(64, 2), (64, 1), (64, 1)], 64, dict(rph=4, tx=1)),
    (1, 64, 64, [(32, 1)], 32, dict(rph=4, tx=2)),
    (1, 64, 64, [(64, 1), (64, 1)], 32, dict(rph=2, tx=2)),
    (1, 64, 64, [(16, 1)], 16, dict(rph=4, tx=2)),
    (1, 128, 64, [(32, 2)], 16, dict(rph=2, tx=4)),
    (3, 40, 24, [(64, 1)], 64, dict(rph=2, tx=2)),                 # ragged: 40 rows = one full + one partial item
    (2, 72, 40, [(64, 1)], 64, dict(rph=4, tx=1)),
    (2, 64, 64, [(64, 1)], 64, None),                              # whatever the config model picks
    # narrow rows (32- / 64-byte pixels), ragged tiles, nearest-x2 sources; loader 1 = cp.async gather instead of TMA
    (1, 32, 32, [(16, 1)], 16, dict(loader=1)),
    (1, 64, 64, [(32, 2)], 16, dict(loader=1)),
    (3, 24, 40, [(16, 1)], 16, dict(tx=4, loader=1)),              # ragged, zero fill on every side
    (3, 24, 40, [(16, 1)], 16, dict(tx=1, halo_stages=2, loader=1)),
    (2, 40, 24, [(32, 1)], 32, dict(tx=2, loader=1)),
    (2, 32, 24, [(32, 2)], 16, dict(tx=2, loader=1)),
    (2, 64, 48, [(32, 2), (16, 1)], 16, dict(rph=2, tx=4, loader=1)),
    (1, 128, 64, [(16, 1)], 16, dict(rph=4, tx=2, loader=1)),
    (3, 24, 40, [(16, 1)], 16, dict(tx=4)),                        # ragged, zero fill on every side
    (3, 24, 40, [(16, 1)], 16, dict(tx=1, halo_stages=2)),
    (2, 40, 24, [(32, 1)], 32, dict(tx=2)),
    (2, 32, 24, [(32, 2)], 16, dict(tx=2)),                        # nearest-x2 source, ragged in x
    (2, 64, 48, [(32, 2), (16, 1)], 16, dict(rph=2, tx=4)),        # 16-channel chunks of a 32-channel source
    (1, 128, 64, [(16, 1)], 16, dict(rph=4, tx=2)),
]


@pytest.mark.parametrize("case", FPROP_CASES)
def test_fprop_matches_conv2d(case):
    from mmrseg_b200 import convplan
    N, H, W, srcs, Cout, force = case
    gen = torch.Generator(device="cuda").manual_seed(6210)
    sources = [(_mk((N, H // up, W // up, C), gen), up) for C, up in srcs]
    cin = sum(c for c, _ in srcs)
    w = torch.randn((Cout, cin, 3, 3), generator=gen, device="cuda") * (1.0 / (cin * 9) ** 0.5)
    out = torch.full((N, H, W, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    plan = convplan.build_fprop_halo(sources, w, out, force=force)
    plan.run()
    torch.cuda.synchronize()
    _check(out, _ref_conv(sources, w))
    # a second launch on the same plan (persistent state must be clean)
    out.fill_(float("nan"))
    plan.run()
    torch.cuda.synchronize()
    _check(out, _ref_conv(sources, w))


@pytest.mark.parametrize("Cout,force", [(128, None), (64, dict(rph=2)), (64, dict(rph=1)), (32, dict(rph=4))])
def test_fprop_epilogue_scale_bias_residual_relu(Cout, force):
    from mmrseg_b200 import convplan
    gen = torch.Generator(device="cuda").manual_seed(1)
    N, H, W, C = 2, 64, 32, 64
    x = _mk((N, H, W, C), gen)
    w = torch.randn((Cout, C, 3, 3), generator=gen, device="cuda") / 24
    scale = torch.rand(Cout, generator=gen, device="cuda") + 0.5
    bias = torch.randn(Cout, generator=gen, device="cuda")
    res = _mk((N, H, W, Cout), gen)
    out = torch.empty((N, H, W, Cout), device="cuda", dtype=torch.bfloat16)
    plan = convplan.build_fprop_halo([(x, 1)], w, out, scale=scale, bias=bias, residual=res, relu=True, force=force)
    plan.run()
    torch.cuda.synchronize()
    ref = _ref_conv([(x, 1)], w) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    ref = torch.relu(ref + res.float().permute(0, 3, 1, 2))
    _check(out, ref)


def test_head_f32_nchw_output():
    from mmrseg_b200 import convplan
    gen = torch.Generator(device="cuda").manual_seed(2)
    N, H, W, C, classes = 2, 32, 40, 16, 2
    x = _mk((N, H, W, C), gen)
    w = torch.randn((classes, C, 3, 3), generator=gen, device="cuda") / 12
    b = torch.randn(classes, generator=gen, device="cuda")
    out = torch.full((N, classes, H, W), float("nan"), device="cuda")
    plan = convplan.build_fprop_halo([(x, 1)], w, None, bias=b, out_f32=out)
    plan.run()
    torch.cuda.synchronize()
    ref = _ref_conv([(x, 1)], w) + b.view(1, -1, 1, 1)
    err = (out - ref).abs().max().item()
    assert err < 1e-3, err


@pytest.mark.parametrize("shape", [(2, 32, 32, 64, 64), (3, 24, 40, 64, 128), (2, 32, 32, 16, 16),
                                   (2, 32, 32, 32, 32), (2, 64, 32, 64, 64, dict(rph=2)),
                                   (2, 64, 32, 64, 64, dict(rph=2, tx=1)), (1, 72, 40, 32, 32, dict(rph=4)),
                                   (2, 64, 64, 16, 16, dict(rph=2))])
def test_batchnorm_statistics_from_epilogue(shape):
    from mmrseg_b200 import convplan
    force = shape[5] if len(shape) > 5 else None
    N, H, W, C, Cout = shape[:5]
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = _mk((N, H, W, C), gen)
    w = torch.randn((Cout, C, 3, 3), generator=gen, device="cuda") / (9 * C) ** 0.5
    out = torch.empty((N, H, W, Cout), device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros((8, 2, Cout), device="cuda", dtype=torch.float64)
    plan = convplan.build_fprop_halo([(x, 1)], w, out, stats=stats, stats_ld=Cout, force=force)
    if force:
        assert all(plan.cfg[k] == v for k, v in force.items())
    plan.run()
    torch.cuda.synchronize()
    _check(out, _ref_conv([(x, 1)], w))
    z = out.double().reshape(-1, Cout)
    got = stats.sum(0)
    assert torch.allclose(got[0], z.sum(0), rtol=1e-5, atol=1e-5 * z.abs().sum(0).max().item())
    assert torch.allclose(got[1], (z * z).sum(0), rtol=1e-5)


DGRAD_CASES = [
    # (N, H, W, [Cs per source], Cout, force)
    (2, 32, 32, [64], 64, None),
    (2, 32, 32, [64, 64, 128], 64, None),
    (1, 32, 32, [64, 64, 64, 64, 64], 32, None),
    (2, 16, 16, [128], 256, None),
    (1, 64, 64, [32], 16, None),
    (1, 32, 32, [16], 16, None),
    (3, 24, 40, [64, 64], 64, None),
    (2, 32, 32, [64, 64], 64, dict(bn=128)),
    (2, 32, 32, [64, 64], 64, dict(bn=64, tx=4)),
    (2, 64, 32, [64, 64, 64], 64, dict(bn=64, rph=2, tx=2)),       # three N tiles, stacked row phases
    (1, 64, 32, [64], 64, dict(rph=4, tx=1)),
    (1, 64, 64, [32], 16, dict(rph=2)),
    (2, 40, 24, [64, 64], 32, dict(bn=64, rph=2)),
]


@pytest.mark.parametrize("case", DGRAD_CASES)
def test_dgrad_matches_autograd(case):
    from mmrseg_b200 import convplan
    N, H, W, sizes, Cout, force = case
    gen = torch.Generator(device="cuda").manual_seed(11)
    cin = sum(sizes)
    cz = -(-Cout // 16) * 16
    dz = torch.zeros((N, H, W, cz), device="cuda", dtype=torch.bfloat16)
    dz[..., :Cout] = _mk((N, H, W, Cout), gen)
    w = torch.randn((Cout, cin, 3, 3), generator=gen, device="cuda") * (1.0 / (Cout * 9) ** 0.5)
    grads = [torch.full((N, H, W, c), float("nan"), device="cuda", dtype=torch.bfloat16) for c in sizes]
    plan = convplan.build_dgrad_halo(dz, w, grads, force=force)
    plan.run()
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_input((N, cin, H, W), w.to(torch.bfloat16).float(),
                                     dz[..., :Cout].float().permute(0, 3, 1, 2), stride=1, padding=1)
    got = torch.cat(grads, 3)
    _check(got, ref)


@pytest.mark.parametrize("case", [
    # (N, H, W, channel counts per source, pooled flags, Cout, force)
    (2, 32, 32, [128, 64], [True, False], 64, None),
    (2, 32, 32, [128, 64], [True, False], 64, dict(bn=64, rph=1, tx=2)),      # rph 1: window rows are lanes 8 apart
    (2, 64, 32, [64, 64, 64], [True, False, False], 64, dict(bn=64, rph=2, tx=2)),
    (1, 64, 32, [64], [True], 64, dict(rph=4, tx=1)),
    (1, 64, 64, [32], [True], 16, None),                                      # x_0_4.conv1: direct-store group
    (3, 24, 40, [64, 64], [True, False], 64, None),                           # ragged tiles, even image
    (2, 32, 32, [64, 64], [False, True], 128, dict(bn=128)),                  # the pooled group is the LAST of its item
])
def test_dgrad_pooled_groups(case):
    """Data gradient with respect to a nearest-x2 source: the epilogue stores the 2x2 sum of the conv-resolution
    gradient (fp32 sum, one bf16 rounding) into a half-resolution tensor."""
    from mmrseg_b200 import convplan
    N, H, W, sizes, pooled, Cout, force = case
    gen = torch.Generator(device="cuda").manual_seed(12)
    cin = sum(sizes)
    cz = -(-Cout // 16) * 16
    dz = torch.zeros((N, H, W, cz), device="cuda", dtype=torch.bfloat16)
    dz[..., :Cout] = _mk((N, H, W, Cout), gen)
    w = torch.randn((Cout, cin, 3, 3), generator=gen, device="cuda") * (1.0 / (Cout * 9) ** 0.5)
    grads = [torch.full((N, H // 2, W // 2, c) if pl else (N, H, W, c), float("nan"), device="cuda", dtype=torch.bfloat16)
             for c, pl in zip(sizes, pooled)]
    plan = convplan.build_dgrad_halo(dz, w, grads, force=force, pooled=pooled)
    plan.run()
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_input((N, cin, H, W), w.to(torch.bfloat16).float(),
                                     dz[..., :Cout].float().permute(0, 3, 1, 2), stride=1, padding=1)
    c0 = 0
    for g, c, pl in zip(grads, sizes, pooled):
        want = ref[:, c0:c0 + c]
        if pl:
            want = F.avg_pool2d(want, 2) * 4.0
        _check(g, want)
        c0 += c


WGRAD_CASES = [
    # (N, H, W, [(C, up)], Cout, force)
    (2, 32, 32, [(64, 1)], 64, None),
    (2, 32, 32, [(64, 1)], 64, dict(tx=1, n_split=1)),
    (2, 32, 32, [(64, 1)], 64, dict(tx=2, n_split=3)),
    (2, 32, 64, [(16, 1)], 16, dict(tx=4)),
    (2, 16, 16, [(128, 1)], 128, None),
    (1, 32, 32, [(128, 2), (64, 1)], 64, None),
    (2, 32, 32, [(64, 2), (64, 1), (64, 1)], 32, dict(tx=2)),
    (1, 32, 32, [(32, 1)], 32, None),
    (1, 64, 64, [(32, 2)], 16, None),
    (2, 32, 32, [(16, 1)], 16, None),
    (2, 32, 32, [(16, 1)], 2, None),          # the head: dz padded to 16 channels
    (3, 24, 40, [(64, 1)], 64, None),         # ragged
    (1, 8, 8, [(256, 1)], 256, None),
    # third generation (filter column on the dz side) against the second, both forced
    (2, 32, 32, [(64, 1)], 64, dict(mode=0)),
    (2, 32, 32, [(64, 1)], 64, dict(mode=1, tx=1, n_split=1)),
    (2, 32, 32, [(64, 1)], 64, dict(mode=1, tx=2, n_split=3)),
    (3, 24, 40, [(64, 1)], 64, dict(mode=1, tx=2)),              # ragged in x and y
    (3, 24, 40, [(64, 1)], 64, dict(mode=1, tx=1)),
    (1, 32, 32, [(128, 2), (64, 1)], 64, dict(mode=1, tx=1)),    # nearest-x2 source
    (1, 32, 48, [(128, 2), (64, 1)], 128, dict(mode=1, tx=2)),
    (2, 32, 32, [(64, 2), (64, 1), (64, 1)], 32, dict(mode=1)),  # bn = 32: N = 96
    (1, 16, 16, [(64, 1)], 96, dict(mode=1)),                    # Cout = 96: bn = 32, three N tiles
    (1, 40, 24, [(64, 1)], 40, None),                            # Cout padded to 48: bn = 16, second generation
    # narrow layers (mode 2: one MMA of N = 3 bn per 16 pixels, cp.async gather) against the second generation
    (2, 32, 64, [(16, 1)], 16, dict(mode=2, tx=4)),
    (2, 32, 64, [(16, 1)], 16, dict(mode=2, tx=2, n_split=5)),
    (2, 32, 64, [(16, 1)], 16, dict(mode=0)),
    (3, 24, 40, [(16, 1)], 16, dict(mode=2, tx=4)),              # ragged in x and y
    (3, 24, 40, [(32, 1)], 32, dict(mode=2, tx=2)),
    (1, 64, 64, [(32, 2)], 16, dict(mode=2)),                    # x_0_4.conv1: nearest-x2 32-channel source
    (2, 32, 48, [(32, 2), (16, 1)], 16, dict(mode=2, tx=4)),     # 16-channel chunks of a 32-channel source
    (2, 32, 32, [(16, 1)], 2, dict(mode=2)),                     # the head: dz padded to 16 channels
    (1, 32, 32, [(32, 1)], 32, dict(mode=2)),                    # x_0_3.conv2
    (1, 32, 32, [(16, 1)], 48, dict(mode=2)),                    # three N tiles of 16
]


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_wgrad_matches_autograd(case):
    from mmrseg_b200 import convplan
    N, H, W, srcs, Cout, force = case
    gen = torch.Generator(device="cuda").manual_seed(21)
    sources = [(_mk((N, H // up, W // up, C), gen), up) for C, up in srcs]
    cin = sum(c for c, _ in srcs)
    cz = -(-Cout // 16) * 16
    dz = torch.zeros((N, H, W, cz), device="cuda", dtype=torch.bfloat16)
    dz[..., :Cout] = _mk((N, H, W, Cout), gen, 0.1)
    dst = torch.full((Cout, cin, 3, 3), float("nan"), device="cuda")
    plan = convplan.build_wgrad_halo(dz, sources, dst, force=force)
    plan.run()
    torch.cuda.synchronize()
    xs = []
    for t, up in sources:
        x = t.float().permute(0, 3, 1, 2)
        xs.append(F.interpolate(x, scale_factor=2, mode="nearest") if up == 2 else x)
    ref = torch.nn.grad.conv2d_weight(torch.cat(xs, 1), (Cout, cin, 3, 3),
                                      dz[..., :Cout].float().permute(0, 3, 1, 2), stride=1, padding=1)
    assert torch.isfinite(dst).all()
    err = (dst - ref).abs().max().item()
    rms = ref.pow(2).mean().sqrt().item()
    assert err <= 2e-3 * max(rms, 1e-6) * 8, (err, rms)
    # accumulate flag
    plan.run(accumulate=True)
    torch.cuda.synchronize()
    assert (dst - 2 * ref).abs().max().item() <= 4e-3 * max(rms, 1e-6) * 8


# (O, I, cb, bn, n_ntiles, nchunks): full tiles, a ragged N tile, ragged K (I = 3 in a 16-channel chunk, I = 40 in 64-channel
# chunks: the scalar path of the batch kernel), narrow rows
PACK_CASES = [(64, 64, 64, 64, 1, 1), (128, 192, 64, 128, 1, 3), (96, 64, 64, 64, 2, 1), (64, 3, 16, 64, 1, 1),
              (32, 40, 64, 32, 1, 1), (16, 32, 32, 16, 1, 1), (2, 16, 16, 16, 1, 1)]


def test_pack_weights_batch_matches_reference_layout():
    """mmr_pack_weights_halo_batch (eight k per thread, 16-byte stores) against the packed layout written out in
    PyTorch -- out[((nt*nchunks + c)*9 + slot)*bn + r][k], fprop and dgrad (mirrored taps, O and I swapped), both tap
    orders -- and against the one-job kernel mmr_pack_weights_halo: bit-exact (one fp32 -> bf16 rounding either way)."""
    import ctypes as C
    from mmrseg_b200 import _lib
    lib = _lib.lib()
    gen = torch.Generator(device="cuda").manual_seed(7)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    jobs, outs, want, singles = [], [], [], []
    keep = []
    for (O, I, cb, bn, nnt, nch) in PACK_CASES:
        w = torch.randn((O, I, 3, 3), generator=gen, device="cuda")
        keep.append(w)
        for mode in (0, 1):
            for layout in (0, 1):
                n_rows, n_k = (O, I) if mode == 0 else (I, O)
                # a dgrad job tiles the INPUT channels over N and the OUTPUT channels over K
                nnt_m = nnt if mode == 0 else -(-n_rows // bn)
                nch_m = nch if mode == 0 else -(-n_k // cb)
                out = torch.full((nnt_m * nch_m * 9 * bn * cb,), -1.0, device="cuda", dtype=torch.bfloat16)
                jobs.append(_lib.MmrPackJob(w.data_ptr(), out.data_ptr(), O, I, mode, cb, bn, nnt_m, nch_m, layout))
                outs.append(out)
                # reference layout in PyTorch
                ww = w if mode == 0 else w.flip(2, 3).permute(1, 0, 2, 3)      # [rows][k][ky][kx], taps mirrored for dgrad
                pad = torch.zeros((nnt_m * bn, nch_m * cb, 3, 3), device="cuda")
                pad[:n_rows, :n_k] = ww
                taps = pad.reshape(nnt_m, bn, nch_m, cb, 9)
                order = [(2 - t % 3) * 3 + t // 3 for t in range(9)] if layout else list(range(9))
                ref = taps[..., order].permute(0, 2, 4, 1, 3).contiguous().to(torch.bfloat16).reshape(-1)
                want.append(ref)
                single = torch.full_like(out, -1.0)
                _lib.check(lib.mmr_pack_weights_halo(C.c_void_p(w.data_ptr()), O, I, mode, cb, bn, nnt_m, nch_m, layout,
                                                     C.c_void_p(single.data_ptr()), stream))
                singles.append(single)
    arr = (_lib.MmrPackJob * len(jobs))(*jobs)
    dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone().cuda()
    per = lib.mmr_pack_items_per_block()
    blocks = sum(-(-(j.n_ntiles * j.nchunks * j.bn * j.cb) // per) for j in jobs)
    _lib.check(lib.mmr_pack_weights_halo_batch(C.c_void_p(dev.data_ptr()), len(jobs), C.c_int64(blocks), stream))
    torch.cuda.synchronize()
    for i, (got, ref, single) in enumerate(zip(outs, want, singles)):
        assert torch.equal(got.view(torch.int16), ref.view(torch.int16)), ("batch vs layout", i, PACK_CASES[i // 4])
        assert torch.equal(single.view(torch.int16), ref.view(torch.int16)), ("single vs layout", i, PACK_CASES[i // 4])
