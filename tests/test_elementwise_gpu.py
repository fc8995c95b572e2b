"""HBM-bound kernels (csrc/elementwise.cu) against torch fp32 references on the same bf16 inputs:
BatchNorm forward/backward, gradient gathering with the 2x2 sum-pool of nearest-x2 upsampling,
max-pool, packing."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from mmrseg_b200 import _lib
    return _lib


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _nchw(t):
    return t.float().permute(0, 3, 1, 2)


def _contribs(items):
    from mmrseg_b200._lib import MmrContrib
    arr = (MmrContrib * len(items))()
    for i, (t, pool2) in enumerate(items):
        arr[i].ptr = t.data_ptr()
        arr[i].pool2 = pool2
    return arr


# C % 16 == 0 takes the 32-byte apply kernels, C = 8 the 8-channel ones
@pytest.mark.parametrize("shape", [(4, 16, 16, 64), (2, 32, 32, 16), (3, 8, 8, 512), (2, 24, 40, 128), (2, 10, 6, 8)])
@pytest.mark.parametrize("with_res", [False, True])
def test_bn_forward_backward(shape, with_res):
    L = _lib()
    lib = L.lib()
    gen = torch.Generator(device="cuda").manual_seed(5)
    n, h, w, c = shape
    P = n * h * w
    z = (torch.randn(shape, generator=gen, device="cuda") * 1.7 + 0.3).to(torch.bfloat16)
    res = torch.randn(shape, generator=gen, device="cuda").to(torch.bfloat16) if with_res else None
    gamma = torch.rand(c, generator=gen, device="cuda") + 0.5
    beta = torch.randn(c, generator=gen, device="cuda") * 0.2
    rm = torch.randn(c, generator=gen, device="cuda") * 0.1
    rv = torch.rand(c, generator=gen, device="cuda") + 0.5
    nbt = torch.zeros((), device="cuda", dtype=torch.int64)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    nblk = 37
    partial = torch.empty((nblk * 2 * c,), device="cuda", dtype=torch.float64)
    st = torch.empty((7, c), device="cuda")
    out = torch.empty(shape, device="cuda", dtype=torch.bfloat16)
    L.check(lib.mmr_bn_stats(_p(z), P, c, _p(partial), nblk, _s()))
    L.check(lib.mmr_bn_finalize(_p(partial), nblk, P, c, _p(gamma), _p(beta), 1e-5, 0.1, _p(rm), _p(rv),
                                _p(nbt), _p(st[0]), _p(st[1]), _p(st[2]), _p(st[3]), _s()))
    L.check(lib.mmr_bn_apply(_p(z), P, c, _p(st[2]), _p(st[3]), _p(res), 1, _p(out), _s()))
    # reference
    zf = _nchw(z).clone().requires_grad_(True)
    gl = gamma.clone().requires_grad_(True)
    bl = beta.clone().requires_grad_(True)
    y = F.batch_norm(zf, rm_ref, rv_ref, gl, bl, True, 0.1, 1e-5)
    resf = _nchw(res).clone().requires_grad_(True) if with_res else None
    yr = torch.relu(y + resf) if with_res else torch.relu(y)
    torch.cuda.synchronize()
    assert (_nchw(out) - yr).abs().max().item() <= 0.03
    assert torch.allclose(rm, rm_ref, atol=1e-5) and torch.allclose(rv, rv_ref, rtol=1e-4, atol=1e-5)
    assert int(nbt) == 1
    # backward: two contributions
    g1 = torch.randn(shape, generator=gen, device="cuda").to(torch.bfloat16)
    g2 = torch.randn(shape, generator=gen, device="cuda").to(torch.bfloat16)
    yr.backward(_nchw(g1) + _nchw(g2))
    g = torch.empty(shape, device="cuda", dtype=torch.bfloat16)
    dz = torch.empty(shape, device="cuda", dtype=torch.bfloat16)
    dgamma = torch.empty(c, device="cuda")
    dbeta = torch.empty(c, device="cuda")
    arr = _contribs([(g1, 0), (g2, 0)])
    L.check(lib.mmr_bn_bwd_reduce(arr, 2, _p(out), _p(z), _p(st[0]), _p(st[1]), n, h, w, c, _p(g),
                                  _p(partial), nblk, _s()))
    L.check(lib.mmr_bn_bwd_finalize(_p(partial), nblk, P, c, _p(gamma), _p(st[1]), _p(dgamma), _p(dbeta), 0,
                                    _p(st[4]), _s()))
    L.check(lib.mmr_bn_bwd_apply(_p(g), _p(z), _p(st[0]), _p(st[1]), _p(st[4]), P, c, _p(dz), _s()))
    torch.cuda.synchronize()
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    assert rel(dgamma, gl.grad) < 5e-3, rel(dgamma, gl.grad)
    assert rel(dbeta, bl.grad) < 5e-3
    assert rel(_nchw(dz), zf.grad) < 1e-2, rel(_nchw(dz), zf.grad)
    if with_res:
        assert rel(_nchw(g), resf.grad) < 1e-2


@pytest.mark.parametrize("c,shift", [(16, 0), (48, 0), (80, 0), (24, 0), (256, 0), (64, 8)])
def test_bn_apply_passes_any_channel_count(c, shift):
    """The three streaming apply passes against their formulas (fp32 on the same bf16 inputs, bf16 output rounding)
    for channel counts whose 16-channel groups are not a power of two (48, 80), for the 8-channel kernels (24)
    and with a grid that wraps the channel groups (256), with and without residual / ReLU; shift = 8: tensors that
    are 16- but not 32-byte aligned must take the 8-channel kernels and give the same results."""
    L = _lib()
    lib = L.lib()
    gen = torch.Generator(device="cuda").manual_seed(c)
    P = 3 * 37 * 29
    def buf(fill):
        t = torch.empty((P * c + shift,), device="cuda", dtype=torch.bfloat16)[shift:].view(P, c)
        if fill is not None:
            t.copy_((torch.randn((P, c), generator=gen, device="cuda") * fill).to(torch.bfloat16))
        return t

    z, g, res, out = buf(1.5), buf(1.0), buf(1.0), buf(None)
    assert z.data_ptr() % 32 == (16 if shift else 0)
    st = torch.randn((7, c), generator=gen, device="cuda")    # mean, invstd, scale, shift, cA, cB, cC
    st[1] = st[1].abs() + 0.5
    zf, gf = z.float(), g.float()
    close = lambda a, b: torch.allclose(a.float(), b, rtol=1e-2, atol=1e-2)
    for r, relu in ((None, 1), (res, 1), (res, 0), (None, 0)):
        L.check(lib.mmr_bn_apply(_p(z), P, c, _p(st[2]), _p(st[3]), _p(r), relu, _p(out), _s()))
        want = zf * st[2] + st[3] + (r.float() if r is not None else 0)
        assert close(out, torch.relu(want) if relu else want)
    xhat = (zf - st[0]) * st[1]
    L.check(lib.mmr_bn_bwd_apply(_p(g), _p(z), _p(st[0]), _p(st[1]), _p(st[4]), P, c, _p(out), _s()))
    assert close(out, st[4] * gf + st[5] * xhat + st[6])
    L.check(lib.mmr_bn_bwd_apply_masked(_p(g), _p(z), _p(st[0]), _p(st[1]), _p(st[4]), _p(st[2]), _p(st[3]), P, c,
                                        _p(out), _s()))
    gm = torch.where(torch.addcmul(st[3], zf, st[2]) > 0, gf, torch.zeros_like(gf))
    assert close(out, st[4] * gm + st[5] * xhat + st[6])


def test_bn_bwd_reduce_fused_mask_from_z():
    """mmr_bn_bwd_reduce_fused: the ReLU mask recomputed from z (scale / shift of the forward pass) is the
    mask read from the stored activation, bit for bit in g, and the fused finalisation matches
    mmr_bn_bwd_reduce + mmr_bn_bwd_finalize."""
    L = _lib()
    lib = L.lib()
    gen = torch.Generator(device="cuda").manual_seed(11)
    n, h, w, c = 3, 20, 12, 64
    shape = (n, h, w, c)
    P = n * h * w
    z = (torch.randn(shape, generator=gen, device="cuda") * 2 + 0.3).to(torch.bfloat16)
    gamma = torch.rand(c, generator=gen, device="cuda") + 0.5
    beta = torch.randn(c, generator=gen, device="cuda") * 0.2
    nblk = 23
    partial = torch.empty((nblk * 2 * c,), device="cuda", dtype=torch.float64)
    st = torch.empty((7, c), device="cuda")
    act = torch.empty(shape, device="cuda", dtype=torch.bfloat16)
    L.check(lib.mmr_bn_stats(_p(z), P, c, _p(partial), nblk, _s()))
    L.check(lib.mmr_bn_finalize(_p(partial), nblk, P, c, _p(gamma), _p(beta), 1e-5, 0.1, None, None, None,
                                _p(st[0]), _p(st[1]), _p(st[2]), _p(st[3]), _s()))
    L.check(lib.mmr_bn_apply(_p(z), P, c, _p(st[2]), _p(st[3]), None, 1, _p(act), _s()))
    g1 = torch.randn(shape, generator=gen, device="cuda").to(torch.bfloat16)
    arr = _contribs([(g1, 0)])
    # unfused reference path
    g_ref = torch.empty(shape, device="cuda", dtype=torch.bfloat16)
    dgamma_ref, dbeta_ref = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    coef_ref = torch.empty((3, c), device="cuda")
    L.check(lib.mmr_bn_bwd_reduce(arr, 1, _p(act), _p(z), _p(st[0]), _p(st[1]), n, h, w, c, _p(g_ref),
                                  _p(partial), nblk, _s()))
    L.check(lib.mmr_bn_bwd_finalize(_p(partial), nblk, P, c, _p(gamma), _p(st[1]), _p(dgamma_ref), _p(dbeta_ref),
                                    0, _p(coef_ref), _s()))
    slots = torch.zeros((8 * 2 * c,), device="cuda", dtype=torch.float64)
    ticket = torch.zeros((1,), device="cuda", dtype=torch.int32)
    for use_z in (False, True):
        g = torch.empty(shape, device="cuda", dtype=torch.bfloat16)
        dgamma, dbeta = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
        coef = torch.empty((3, c), device="cuda")
        L.check(lib.mmr_bn_bwd_reduce_fused(arr, 1, None if use_z else _p(act), _p(z), _p(st[0]), _p(st[1]), n, h,
                                            w, c, _p(g), _p(slots), nblk, _p(gamma), _p(dgamma), _p(dbeta), 0,
                                            _p(coef), _p(ticket), _p(st[2]) if use_z else None,
                                            _p(st[3]) if use_z else None, _s()))
        torch.cuda.synchronize()
        assert torch.equal(g, g_ref), use_z
        assert int(ticket) == 0 and float(slots.abs().max()) == 0.0      # re-armed by the last CTA
        assert torch.allclose(dgamma, dgamma_ref, rtol=1e-5, atol=1e-5)
        assert torch.allclose(dbeta, dbeta_ref, rtol=1e-5, atol=1e-5)
        assert torch.allclose(coef, coef_ref, rtol=1e-5, atol=1e-7)
    # no-g variant: the reduction writes only the sums, the apply pass recomputes the masked gradient from dx
    dz_ref = torch.empty(shape, device="cuda", dtype=torch.bfloat16)
    dz = torch.empty(shape, device="cuda", dtype=torch.bfloat16)
    L.check(lib.mmr_bn_bwd_apply(_p(g_ref), _p(z), _p(st[0]), _p(st[1]), _p(coef_ref), P, c, _p(dz_ref), _s()))
    coef = torch.empty((3, c), device="cuda")
    L.check(lib.mmr_bn_bwd_reduce_fused(arr, 1, None, _p(z), _p(st[0]), _p(st[1]), n, h, w, c, None, _p(slots), nblk,
                                        _p(gamma), _p(dgamma), _p(dbeta), 0, _p(coef), _p(ticket), _p(st[2]),
                                        _p(st[3]), _s()))
    L.check(lib.mmr_bn_bwd_apply_masked(_p(g1), _p(z), _p(st[0]), _p(st[1]), _p(coef), _p(st[2]), _p(st[3]), P, c,
                                        _p(dz), _s()))
    torch.cuda.synchronize()
    assert torch.allclose(coef, coef_ref, rtol=1e-5, atol=1e-7)
    assert (dz.float() - dz_ref.float()).abs().max().item() <= 2.0 ** -7 * dz_ref.float().abs().max().item()


def test_grad_gather_pool2_and_mask():
    L = _lib()
    lib = L.lib()
    gen = torch.Generator(device="cuda").manual_seed(6)
    n, h, w, c = 2, 8, 12, 64
    a = torch.randn((n, h, w, c), generator=gen, device="cuda").to(torch.bfloat16)
    g_same = torch.randn((n, h, w, c), generator=gen, device="cuda").to(torch.bfloat16)
    g_up = torch.randn((n, 2 * h, 2 * w, c), generator=gen, device="cuda").to(torch.bfloat16)
    out = torch.empty((n, h, w, c), device="cuda", dtype=torch.bfloat16)
    nblk = 11
    partial = torch.empty((nblk * 2 * c,), device="cuda", dtype=torch.float64)
    arr = _contribs([(g_same, 0), (g_up, 1)])
    L.check(lib.mmr_grad_gather(arr, 2, _p(a), n, h, w, c, _p(out), _p(partial), nblk, _s()))
    torch.cuda.synchronize()
    want = _nchw(g_same) + F.avg_pool2d(_nchw(g_up), 2) * 4
    want = want * (_nchw(a) > 0)
    assert (_nchw(out) - want).abs().max().item() <= 0.05
    sums = partial.view(nblk, 2, c)[:, 0].sum(0)
    assert torch.allclose(sums.float(), want.sum((0, 2, 3)), rtol=1e-3, atol=1e-2)


# even H and W: the 2x2-block backward kernel; odd sizes and three contributions: the per-pixel one
@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (1, 14, 10, 64), (2, 13, 9, 16), (1, 12, 15, 8), (3, 2, 2, 8)])
def test_maxpool(shape):
    L = _lib()
    lib = L.lib()
    gen = torch.Generator(device="cuda").manual_seed(7)
    n, h, w, c = shape
    x = torch.randn(shape, generator=gen, device="cuda").to(torch.bfloat16)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    out = torch.empty((n, ho, wo, c), device="cuda", dtype=torch.bfloat16)
    idx = torch.empty((n, ho, wo, c), device="cuda", dtype=torch.uint8)
    L.check(lib.mmr_maxpool3x3s2_fwd(_p(x), n, h, w, c, _p(out), _p(idx), _s()))
    xf = _nchw(x).clone().requires_grad_(True)
    y = F.max_pool2d(xf, 3, 2, 1)
    torch.cuda.synchronize()
    assert torch.equal(_nchw(out), y.detach())
    g = torch.randn((n, ho, wo, c), generator=gen, device="cuda").to(torch.bfloat16)
    y.backward(_nchw(g))
    gin = torch.empty(shape, device="cuda", dtype=torch.bfloat16)
    arr = _contribs([(g, 0)])
    L.check(lib.mmr_maxpool3x3s2_bwd(arr, 1, _p(idx), n, h, w, c, _p(gin), _s()))
    torch.cuda.synchronize()
    assert (_nchw(gin) - xf.grad).abs().max().item() <= 0.03
    # several contributions are summed in front of the routing
    extra = [torch.randn((n, ho, wo, c), generator=gen, device="cuda").to(torch.bfloat16) for _ in range(2)]
    for k in (2, 3):
        gs = [g] + extra[:k - 1]
        xf.grad = None
        F.max_pool2d(xf, 3, 2, 1).backward(sum(_nchw(t) for t in gs))
        L.check(lib.mmr_maxpool3x3s2_bwd(_contribs([(t, 0) for t in gs]), k, _p(idx), n, h, w, c, _p(gin), _s()))
        torch.cuda.synchronize()
        assert (_nchw(gin) - xf.grad).abs().max().item() <= 0.06


def test_pack_unpack_im2col_headprep():
    L = _lib()
    lib = L.lib()
    gen = torch.Generator(device="cuda").manual_seed(8)
    n, h, w = 2, 32, 64
    x = torch.randn((n, 3, h, w), generator=gen, device="cuda")
    packed = torch.empty((n, h, w, 8), device="cuda", dtype=torch.bfloat16)
    L.check(lib.mmr_pack_nchw_f32_to_nhwc_bf16(_p(x), n, 3, h, w, _p(packed), 8, _s()))
    back = torch.empty((n, 3, h, w), device="cuda")
    L.check(lib.mmr_unpack_nhwc_bf16_to_nchw_f32(_p(packed), n, 3, 8, h, w, _p(back), _s()))
    torch.cuda.synchronize()
    assert torch.equal(back, x.to(torch.bfloat16).float())
    assert packed[..., 3:].abs().max().item() == 0
    # im2col of the 7x7 s2 p3 stem == unfold
    ho, wo = h // 2, w // 2
    mat = torch.empty((n * ho * wo, 160), device="cuda", dtype=torch.bfloat16)
    L.check(lib.mmr_stem_im2col(_p(x), n, h, w, _p(mat), 160, None, None, _s()))
    torch.cuda.synchronize()
    cols = F.unfold(x, 7, padding=3, stride=2)  # [n, 147, ho*wo], row = c*49 + ky*7 + kx
    want = cols.permute(0, 2, 1).reshape(n * ho * wo, 147).to(torch.bfloat16)
    assert torch.equal(mat[:, :147], want)
    assert mat[:, 147:].abs().max().item() == 0
    # head gradient prep
    dl = torch.randn((n, 2, h, w), generator=gen, device="cuda")
    g = torch.empty((n, h, w, 16), device="cuda", dtype=torch.bfloat16)
    db = torch.empty(2, device="cuda")
    L.check(lib.mmr_head_grad_prep(_p(dl), n, 2, h, w, _p(g), 16, _p(db), 0, _s()))
    torch.cuda.synchronize()
    assert torch.equal(g[..., :2].float().permute(0, 3, 1, 2), dl.to(torch.bfloat16).float())
    assert g[..., 2:].abs().max().item() == 0
    assert torch.allclose(db, dl.sum((0, 2, 3)), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("n,classes,h,w", [(2, 2, 24, 20), (1, 3, 16, 16), (3, 10, 12, 8), (2, 16, 8, 8), (2, 3, 7, 9)])
def test_head_grad_prep_layouts(n, classes, h, w):
    """fp32 NCHW dlogits -> bf16 NHWC padded to 16 channels + per-class sums: bit-exact conversion for the
    four-pixel kernel (<= 4 classes, <= 16 classes) and the scalar one (H*W not a multiple of 4), bias gradient
    with and without accumulation."""
    L = _lib()
    lib = L.lib()
    gen = torch.Generator(device="cuda").manual_seed(n * 100 + classes)
    dl = torch.randn((n, classes, h, w), generator=gen, device="cuda")
    g = torch.full((n, h, w, 16), 7.0, device="cuda", dtype=torch.bfloat16)
    db = torch.full((classes,), 3.0, device="cuda")
    L.check(lib.mmr_head_grad_prep(_p(dl), n, classes, h, w, _p(g), 16, _p(db), 0, _s()))
    torch.cuda.synchronize()
    assert torch.equal(g[..., :classes].permute(0, 3, 1, 2), dl.to(torch.bfloat16))
    assert classes == 16 or g[..., classes:].abs().max().item() == 0
    want = dl.double().sum((0, 2, 3))
    assert torch.allclose(db.double(), want, rtol=1e-5, atol=1e-5)
    L.check(lib.mmr_head_grad_prep(_p(dl), n, classes, h, w, _p(g), 16, _p(db), 1, _s()))
    torch.cuda.synchronize()
    assert torch.allclose(db.double(), 2 * want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("planes,h,w,f", [(6, 16, 24, 2), (4, 8, 8, 4), (2, 4, 6, 8), (3, 5, 7, 2), (2, 6, 5, 3), (5, 7, 9, 1)])
def test_nearest_f32_upsample_and_sum_pool(planes, h, w, f):
    """Auxiliary (deep-supervision) logits: fp32 NCHW nearest upsampling by f (vector kernel for f = 2 / f % 4 == 0,
    scalar one otherwise) is an exact copy; its adjoint sums f x f windows."""
    L = _lib()
    lib = L.lib()
    gen = torch.Generator(device="cuda").manual_seed(planes * 100 + f)
    x = torch.randn((1, planes, h, w), generator=gen, device="cuda")
    up = torch.full((1, planes, h * f, w * f), 3.0, device="cuda")
    L.check(lib.mmr_upsample_nearest_f32_nchw(_p(x), C.c_int64(planes), h, w, f, _p(up), _s()))
    torch.cuda.synchronize()
    assert torch.equal(up, F.interpolate(x, scale_factor=f, mode="nearest"))
    g = torch.randn((1, planes, h * f, w * f), generator=gen, device="cuda")
    pooled = torch.empty_like(x)
    L.check(lib.mmr_sumpool_f32_nchw(_p(g), C.c_int64(planes), h, w, f, _p(pooled), _s()))
    torch.cuda.synchronize()
    assert torch.allclose(pooled, F.avg_pool2d(g, f) * float(f * f), rtol=1e-5, atol=1e-5)
