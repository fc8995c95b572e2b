"""Host-side geometry for the tcgen05 implicit-GEMM convolution (csrc/conv_gemm.cu).

Turns "conv k x k / stride s / pad p over the channel-concatenation of sources, some of
them nearest-x2 upsampled" (smp DecoderBlock.forward: F.interpolate(scale_factor=2,
mode='nearest') -> torch.cat -> Conv2dReLU; torchvision BasicBlock convs) into the K-step
tables of include/mmrseg.h.  Nothing is replicated or concatenated in memory: an
upsampled source is read at (y>>1, x>>1) by splitting the output pixels into their four
parities, for which the gather is a plain shifted box.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import (MMR_OUT_BF16_NHWC, MMR_OUT_F32_NCHW, MmrConvClass, MmrConvDesc, MmrKStep,
                   MmrOutSeg, MmrSrc)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def choose_box(gx, gy, n):
    """Pick (box_w, box_h, box_n) with product 128 minimising padded pixels."""
    best = None
    for bw in (16, 8, 32, 4, 64, 2, 128, 1):
        for bh in (8, 16, 4, 32, 2, 64, 1, 128):
            if 128 % (bw * bh) or bw * bh > 128:
                continue
            bn = 128 // (bw * bh)
            waste = (-(-gx // bw) * bw) * (-(-gy // bh) * bh) * (-(-n // bn) * bn)
            # prefer wider rows (contiguous NHWC runs) then squarer boxes on ties
            key = (waste, 0 if bw >= 8 else 1, abs(bw - 2 * bh))
            if best is None or key < best[0]:
                best = (key, (bw, bh, bn))
    return best[1]


def pick_bk(channels):
    """Largest K-step width (64/32/16) dividing every segment's channel count."""
    for bk in (64, 32, 16):
        if all(c % bk == 0 for c in channels):
            return bk
    raise ValueError("channel counts %r need a common divisor of 16" % (channels,))


def pick_bn(sizes, cap=128):
    for bn in (256, 128, 64, 32, 16):
        if bn <= cap and all(s % bn == 0 for s in sizes):
            return bn
    return 16


class ConvPlan:
    """Owns one mmr_conv_plan (tensor maps + device tables) and the tensors it points at."""

    def __init__(self, desc, keep):
        self._keep = keep  # python refs that must outlive the plan (tensors, tables)
        self.desc = desc
        h = C.c_void_p()
        _lib.check(_lib.lib().mmr_conv_plan_create(C.byref(desc), C.byref(h)))
        self.handle = h
        self.flops = 0

    def run(self, stream=None, impl=0):
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _lib.check(_lib.lib().mmr_conv_plan_run(self.handle, impl, C.c_void_p(s)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().mmr_conv_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def _make_desc(srcs, ksteps_per_cls, cls_meta, weights, bk, bn, out_segs, grid, n_img, o_mul,
               Hout, Wout, cout_total, scale, bias, residual, relu, out_mode, box=None):
    d = MmrConvDesc()
    d.nsrc = len(srcs)
    for i, (t, es) in enumerate(srcs):
        N, H, W, Cc = t.shape
        d.src[i] = MmrSrc(t.data_ptr(), Cc, W, H, N, es)
    d.weights = weights.data_ptr()
    d.w_rows, d.w_cols = weights.shape
    d.bk, d.bn = bk, bn
    gx, gy = grid
    bw, bh, bnimg = box if box is not None else choose_box(gx, gy, n_img)
    d.box_w, d.box_h, d.box_n = bw, bh, bnimg
    flat = []
    d.ncls = len(ksteps_per_cls)
    for ci, ks in enumerate(ksteps_per_cls):
        oy_add, ox_add = cls_meta[ci]
        d.cls[ci] = MmrConvClass(len(flat), len(ks), oy_add, ox_add)
        flat.extend(ks)
    arr = (MmrKStep * max(1, len(flat)))(*flat)
    d.nksteps = len(flat)
    d.ksteps = C.cast(arr, C.POINTER(MmrKStep))
    segs = (MmrOutSeg * len(out_segs))(*[MmrOutSeg(t.data_ptr(), ldc, coff, 1, 0, 0)
                                         for (t, ldc, coff) in out_segs])
    d.n_tiles_n = len(out_segs)
    d.outsegs = C.cast(segs, C.POINTER(MmrOutSeg))
    d.gx_count, d.gy_count, d.n_img = gx, gy, n_img
    d.oy_mul, d.ox_mul = o_mul
    d.Hout, d.Wout = Hout, Wout
    d.cout_total = cout_total
    d.scale = scale.data_ptr() if scale is not None else None
    d.bias = bias.data_ptr() if bias is not None else None
    d.residual = residual.data_ptr() if residual is not None else None
    d.res_ldc = residual.shape[-1] if residual is not None else 0
    d.relu = int(bool(relu))
    d.out_mode = out_mode
    keep = [arr, segs, [t for t, _ in srcs], weights, [t for t, _, _ in out_segs], scale, bias,
            residual]
    return d, keep


def build_fprop(sources, weights, ksize, stride, pad, out, *, scale=None, bias=None, residual=None,
                relu=False, out_mode=MMR_OUT_BF16_NHWC, cout=None, bn=None, box=None):
    """Forward conv.  sources: list of (tensor [N,Hs,Ws,Cs] bf16, up in {1,2}) in concat order.
    weights: bf16 [Cout_rows][taps*Cin_total], column = (ky*k+kx)*Cin_total + ci.
    out: bf16 [N,Ho,Wo,Cout] (or fp32 [N,Cout,Ho,Wo] when out_mode is F32_NCHW)."""
    N = sources[0][0].shape[0]
    Hin = sources[0][0].shape[1] * sources[0][1]
    Win = sources[0][0].shape[2] * sources[0][1]
    for t, up in sources:
        assert t.dtype == torch.bfloat16 and t.is_contiguous()
        assert (t.shape[1] * up, t.shape[2] * up) == (Hin, Win), "sources disagree on resolution"
    Ho = (Hin + 2 * pad - ksize) // stride + 1
    Wo = (Win + 2 * pad - ksize) // stride + 1
    cin_total = sum(t.shape[3] for t, _ in sources)
    assert weights.shape[1] == ksize * ksize * cin_total, (weights.shape, ksize, cin_total)
    if out_mode == MMR_OUT_F32_NCHW:
        cout = out.shape[1] if cout is None else cout
        assert tuple(out.shape) == (N, out.shape[1], Ho, Wo)
        ldc = out.shape[1]
    else:
        cout = out.shape[3] if cout is None else cout
        assert tuple(out.shape[:3]) == (N, Ho, Wo)
        ldc = out.shape[3]
    bk = pick_bk([t.shape[3] for t, _ in sources])
    if bn is None:
        bn = min(128, -(-cout // 16) * 16)
    n_tiles = -(-cout // bn)
    out_segs = [(out, ldc, i * bn) for i in range(n_tiles)]
    any_up = any(up == 2 for _, up in sources)
    assert stride in (1, 2)
    assert not (any_up and stride != 1), "upsampled sources only with stride 1"

    def ksteps_for(py, px):
        ks = []
        seg_off = 0
        for si, (t, up) in enumerate(sources):
            Cs = t.shape[3]
            for c0 in range(0, Cs, bk):
                for ky in range(ksize):
                    for kx in range(ksize):
                        wk = (ky * ksize + kx) * cin_total + seg_off + c0
                        if any_up:
                            if up == 2:
                                ax = ay = 1
                                by = (py + ky - pad) >> 1
                                bx = (px + kx - pad) >> 1
                            else:
                                ax = ay = 2
                                by = py + ky - pad
                                bx = px + kx - pad
                        else:
                            ax = ay = stride
                            by = ky - pad
                            bx = kx - pad
                        ks.append(MmrKStep(si, c0, ax, bx, ay, by, wk, 0))
            seg_off += Cs
        return ks

    if any_up:
        assert Ho % 2 == 0 and Wo % 2 == 0
        classes = [(py, px) for py in (0, 1) for px in (0, 1)]
        ksteps = [ksteps_for(py, px) for py, px in classes]
        srcs = [(t, 1 if up == 2 else 2) for t, up in sources]
        grid, o_mul = (Wo // 2, Ho // 2), (2, 2)
    else:
        classes = [(0, 0)]
        ksteps = [ksteps_for(0, 0)]
        srcs = [(t, stride) for t, _ in sources]
        grid, o_mul = (Wo, Ho), (1, 1)
    d, keep = _make_desc(srcs, ksteps, classes, weights, bk, bn, out_segs, grid, N, o_mul, Ho, Wo,
                         cout, scale, bias, residual, relu, out_mode, box)
    plan = ConvPlan(d, keep)
    plan.flops = 2 * N * Ho * Wo * cout * ksize * ksize * cin_total
    return plan


def build_dgrad(dz, weights_d, ksize, stride, pad, in_hw, grads, *, bn=None, box=None, bias=None):
    """Data gradient.  dz: [N,Ho,Wo,Cout] bf16.  weights_d: bf16 [Cin_total][taps*Cout],
    column = (ky*k+kx)*Cout + co.  in_hw: (Hin, Win) of the (virtually concatenated) conv
    input.  grads: list of (tensor [N,Hin,Win,Cs] bf16) in concat order — one per source;
    for an upsampled source the tensor is still at (Hin, Win) and its consumer 2x2-pools it."""
    N, Ho, Wo, Cout = dz.shape
    Hin, Win = in_hw
    assert dz.dtype == torch.bfloat16 and dz.is_contiguous()
    cin_total = sum(g.shape[3] for g in grads)
    assert tuple(weights_d.shape) == (cin_total, ksize * ksize * Cout), (weights_d.shape,)
    bk = pick_bk([Cout])
    sizes = [g.shape[3] for g in grads]
    if bn is None:
        bn = pick_bn(sizes)
    out_segs = []
    for g in grads:
        assert tuple(g.shape[:3]) == (N, Hin, Win)
        for c in range(0, g.shape[3], bn):
            out_segs.append((g, g.shape[3], c))

    def ksteps_for(py, px):
        ks = []
        for c0 in range(0, Cout, bk):
            for ky in range(ksize):
                for kx in range(ksize):
                    wk = (ky * ksize + kx) * Cout + c0
                    if stride == 1:
                        ks.append(MmrKStep(0, c0, 1, pad - kx, 1, pad - ky, wk, 0))
                    else:
                        ty, tx = py + pad - ky, px + pad - kx
                        if ty % 2 or tx % 2:
                            continue
                        ks.append(MmrKStep(0, c0, 1, tx // 2, 1, ty // 2, wk, 0))
        return ks

    if stride == 1:
        classes, ksteps = [(0, 0)], [ksteps_for(0, 0)]
        grid, o_mul = (Win, Hin), (1, 1)
    else:
        assert stride == 2 and Hin % 2 == 0 and Win % 2 == 0
        classes = [(py, px) for py in (0, 1) for px in (0, 1)]
        ksteps = [ksteps_for(py, px) for py, px in classes]
        grid, o_mul = (Win // 2, Hin // 2), (2, 2)
    # bias: a per-output-channel constant added in the epilogue -- what makes this "data gradient" the FORWARD
    # pass of a ConvTranspose2d(k, stride) (whose arithmetic is exactly the data gradient of Conv2d(k, stride))
    d, keep = _make_desc([(dz, 1)], ksteps, classes, weights_d, bk, bn, out_segs, grid, N, o_mul,
                         Hin, Win, cin_total, None, bias, None, False, MMR_OUT_BF16_NHWC, box)
    plan = ConvPlan(d, keep)
    plan.flops = 2 * N * Ho * Wo * Cout * ksize * ksize * cin_total
    return plan


def choose_pixbox(gx, gy, n, total=32):
    """(w, h, n) box with product `total` minimising padded pixels (wgrad K-step)."""
    best = None
    for bw in (8, 16, 4, 32, 2, 1):
        for bh in (4, 8, 2, 16, 1, 32):
            if bw * bh > total or total % (bw * bh):
                continue
            bn = total // (bw * bh)
            waste = (-(-gx // bw) * bw) * (-(-gy // bh) * bh) * (-(-n // bn) * bn)
            key = (waste, 0 if bw >= 8 else 1, bn)
            if best is None or key < best[0]:
                best = (key, (bw, bh, bn))
    return best[1]


class WgradPlan:
    def __init__(self, desc, keep):
        self._keep = keep
        self.desc = desc
        h = C.c_void_p()
        _lib.check(_lib.lib().mmr_wgrad_plan_create(C.byref(desc), C.byref(h)))
        self.handle = h
        self.flops = 0

    def run(self, stream=None, impl=0, accumulate=False):
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _lib.check(_lib.lib().mmr_wgrad_plan_run(self.handle, impl, int(accumulate), C.c_void_p(s)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().mmr_wgrad_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def build_wgrad(dz, sources, ksize, stride, pad, dst, *, cout_gemm=None, dst_cin=None, n_sms=148,
                n_split=None, partial=None, dz_c0=0):
    """Weight gradient.  dz: [N,Ho,Wo,Cz] bf16 (Cz >= cout_gemm).  sources: as in build_fprop.
    dst: fp32 OIHW gradient [Cout][Cin_total][k][k] (written or accumulated at run time).
    dz_c0: first channel of dz this plan reads (a plan covers at most 512 output channels; wider layers run one
    plan per slice: dz channels [dz_c0, dz_c0 + cout_gemm) -> dst rows of the same range, passed as dst)."""
    from ._lib import MmrWgChunk, MmrWgradDesc
    N, Ho, Wo, Cz = dz.shape
    assert dz.dtype == torch.bfloat16 and dz.is_contiguous()
    assert dst.dtype == torch.float32 and dst.is_contiguous()
    cout = dst.shape[0]
    taps = ksize * ksize
    cin_total = sum(t.shape[3] for t, _ in sources)
    if dst_cin is None:
        dst_cin = dst.shape[1]
    assert dst.numel() == cout * dst_cin * taps
    if cout_gemm is None:
        cout_gemm = Cz - dz_c0
    assert 0 <= dz_c0 and dz_c0 + cout_gemm <= Cz and dz_c0 % 8 == 0
    dz_ptr = dz.data_ptr() + 2 * dz_c0
    chunk_ch = pick_bk([t.shape[3] for t, _ in sources])
    cpm = 128 // chunk_ch
    any_up = any(up == 2 for _, up in sources)
    assert not (any_up and stride != 1)
    classes = [(py, px) for py in (0, 1) for px in (0, 1)] if any_up else [(0, 0)]
    chunks = []
    seg_off = 0
    for si, (t, up) in enumerate(sources):
        assert t.dtype == torch.bfloat16 and t.is_contiguous()
        for c0 in range(0, t.shape[3], chunk_ch):
            for ky in range(ksize):
                for kx in range(ksize):
                    ch = MmrWgChunk()
                    ch.src, ch.c0 = si, c0
                    for ci, (py, px) in enumerate(classes):
                        if any_up:
                            if up == 2:
                                ch.a = 1
                                ch.by[ci] = (py + ky - pad) >> 1
                                ch.bx[ci] = (px + kx - pad) >> 1
                            else:
                                ch.a = 2
                                ch.by[ci] = py + ky - pad
                                ch.bx[ci] = px + kx - pad
                        else:
                            ch.a = stride
                            ch.by[ci] = ky - pad
                            ch.bx[ci] = kx - pad
                    ch.dst_ci = seg_off + c0
                    ch.dst_tap = ky * ksize + kx
                    chunks.append(ch)
        seg_off += t.shape[3]
    while len(chunks) % cpm:
        pad_ch = MmrWgChunk()
        C.memmove(C.byref(pad_ch), C.byref(chunks[0]), C.sizeof(MmrWgChunk))
        pad_ch.dst_tap = -1
        chunks.append(pad_ch)
    arr = (MmrWgChunk * len(chunks))(*chunks)

    d = MmrWgradDesc()
    if any_up:
        gx, gy = Wo // 2, Ho // 2
        d.dz = MmrSrc(dz_ptr, Cz, Wo, Ho, N, 2)
        d.dz_a = 2
        for ci, (py, px) in enumerate(classes):
            d.dz_bx[ci], d.dz_by[ci] = px, py
        srcs = [(t, 1 if up == 2 else 2) for t, up in sources]
    else:
        gx, gy = Wo, Ho
        d.dz = MmrSrc(dz_ptr, Cz, Wo, Ho, N, 1)
        d.dz_a = 1
        srcs = [(t, stride) for t, _ in sources]
    d.nsrc = len(srcs)
    for i, (t, es) in enumerate(srcs):
        n_, h_, w_, c_ = t.shape
        d.src[i] = MmrSrc(t.data_ptr(), c_, w_, h_, n_, es)
    d.ncls = len(classes)
    d.nchunks = len(chunks)
    d.chunks = C.cast(arr, C.POINTER(MmrWgChunk))
    d.chunk_ch = chunk_ch
    d.cout_gemm = cout_gemm
    d.kp_w, d.kp_h, d.kp_n = choose_pixbox(gx, gy, N)
    d.gx_count, d.gy_count, d.n_img = gx, gy, N
    ksteps = (-(-gx // d.kp_w)) * (-(-gy // d.kp_h)) * (-(-N // d.kp_n)) * len(classes)
    nt_cols = min(256, cout_gemm)
    n_ntiles = cout_gemm // nt_cols
    n_mtiles = len(chunks) // cpm
    max_mt = min(8, 512 // nt_cols)
    n_groups = -(-n_mtiles // max_mt)
    if n_split is None:
        n_split = max(1, n_sms // (n_groups * n_ntiles))
        n_split = max(1, min(n_split, ksteps // 4))
    n_split = max(1, min(n_split, ksteps))
    per_split = n_mtiles * 128 * cout_gemm
    if partial is None:
        partial = torch.empty((n_split * per_split,), device=dz.device, dtype=torch.float32)
    else:  # shared scratch: shrink the split factor until the partials fit
        assert partial.dtype == torch.float32 and partial.numel() >= per_split
        n_split = min(n_split, partial.numel() // per_split)
    d.n_split = n_split
    d.partial = partial.data_ptr()
    d.dst = dst.data_ptr()
    d.dst_cout, d.dst_cin, d.dst_taps = cout, dst_cin, taps
    d.chunk_valid_ch = chunk_ch
    plan = WgradPlan(d, [arr, partial, dz, dst, [t for t, _ in sources]])
    plan.flops = 2 * N * Ho * Wo * cout * taps * cin_total
    plan.n_ctas = n_groups * n_ntiles * n_split
    return plan


# ------------------------------------------------------------------------------------------
# Second kernel generation (csrc/conv_halo.cu): 3x3 / stride 1 / pad 1 convolutions.
# ------------------------------------------------------------------------------------------
HALO_SMEM = 227 * 1024 - 1024


def _mma_clk(bn):
    """Cycles of one M=128 x bn x K=16 SS-mode tcgen05.mma: tensor-pipe floor bn/2, or the shared
    memory operand read (4 KB of A + bn*32 B of B at 128 B/clk) — measured, scripts/probe/probe3.cu."""
    return max(bn / 2.0, (4096 + bn * 32) / 128.0)


def halo_config(H, W, N, cb, nchunks, cout, any_up, bf16_out=True, n_sms=148, seg_sizes=None, force=None,
                stats=False):
    """Pick (bn, tx, rph, tps, pipeline depths) for a halo-kernel plan: minimise the modelled time of the
    whole layer (MMA issue under the shared-memory operand limit, L2 -> shared traffic, exposed
    epilogue, wave quantisation over the SMs) subject to 227 KB of shared memory and 512 TMEM columns.
    seg_sizes: channel counts of the destination tensors in concat order (dgrad); an N tile's store
    groups must not straddle them.  stats: the launch takes BatchNorm statistics in its epilogue.
    rph > 1 (row-phase stacking, csrc/conv_halo.cu) reads the activation operand once for up to three
    vertically adjacent output rows: the remedy for N tiles of 64 channels and fewer."""
    cpad = -(-cout // 16) * 16
    seg_sizes = seg_sizes or [cpad]
    best = None
    no_r = bool(__import__("os").environ.get("MMR_NO_RPH"))
    # halo loader of the narrow-row layers: 0 = TMA, 1 = cp.async gather by two warps.  Default: the gather for
    # 16-channel sources at their own resolution (measured on x_0_4.conv2 / the head: 0.075 -> 0.064 ms; two warps
    # sustain ~19 B/clk/SM of LDGSTS, TMA 10 B/clk/SM on 32-byte rows), TMA for 64-byte rows and nearest-x2 sources
    # (x_0_4.conv1 fprop 0.079 vs 0.097 ms with the gather).  force['loader'] / MMR_HALO_LOADER override.
    want = (force or {}).get("loader", os.environ.get("MMR_HALO_LOADER"))
    cpl = cb <= 32 and (int(want) == 1 if want is not None else (cb == 16 and not any_up))
    # N tiles wider than 128 (or 96/160/192) are legal for the kernel but measured slower than 64/128 with
    # a wider macro tile (their weight slots crowd the shared-memory port): scripts/sweep_halo.sh
    for bn in ((force or {}).get("bn"),) if force and "bn" in force else (128, 64, 32, 16):
        if bn > cpad or cpad % bn:
            continue
        sg = 64 if bn % 64 == 0 else (32 if bn % 32 == 0 else 16)
        if any(s % sg for s in seg_sizes):
            continue
        n_nt = cpad // bn
        gpn = bn // sg
        if n_nt * gpn > 16:
            continue
        for rph in (4, 2, 1):
            if rph > 1:
                # the fast epilogue only; statistics need one channel set per CTA; fp32 logits stay unstacked
                if no_r or not bf16_out or 3 * bn > 256 or H % rph or (any_up and H % (16 * rph)):
                    continue
                if stats and not (n_nt == 1 and gpn == 1):
                    continue
                if H < 16 * rph:
                    continue
            for tx in (4, 2, 1):
                if tx > 1 and 8 * tx > -(-W // 8) * 8:
                    continue
                if rph == 4 and tx == 4:
                    continue
                pitch = 8 * tx + (4 if any_up else 2)
                rows = 16 * rph + 2
                halo_stage = -(-(rows * pitch * cb * 2) // 1024) * 1024
                for acc_bufs in (2, 1):
                    if acc_bufs * tx * rph * bn > 512:
                        continue
                    for tps in ((3,) if rph > 1 else (9, 3, 1)):
                        if tps * bn > 256:
                            continue
                        w_slot = -(-(tps * bn * cb * 2) // 1024) * 1024
                        for out_stages in ((2, 1) if bf16_out else (0,)):
                            out_stage = 128 * sg * 2
                            for halo_stages in (3, 2):
                                rest = HALO_SMEM - halo_stages * halo_stage - out_stages * out_stage
                                w_slots = min(8, rest // w_slot)
                                if w_slots < 2:
                                    continue
                                cfg = dict(bn=bn, sg=sg, n_ntiles=n_nt, tx=tx, rph=rph, tps=tps, acc_bufs=acc_bufs,
                                           halo_stages=halo_stages, w_slots=w_slots, out_stages=out_stages)
                                if force and any(cfg[k] != v for k, v in force.items() if k in cfg):
                                    continue
                                items = N * (-(-H // (16 * rph))) * (-(-W // (8 * tx))) * n_nt
                                # One item (tx * rph M tiles x bn channels).  Measured (scripts/probe/probe3.cu, ncu
                                # of x_1_3.conv1): an SS-mode MMA costs max(N/2 clk of tensor pipe, its operand
                                # bytes at 128 B/clk), and the same 128 B/clk shared-memory port also takes the
                                # TMA fills and the epilogue staging, which is what bounds the bn = 64 layers.
                                ksteps = cb // 16
                                if rph == 1:
                                    mma_list = [bn] * 9
                                else:   # per filter column: row shifts ty = -1 .. rph, 1-3 phases each
                                    mma_list = [bn * (min(rph - 1, ty + 1) - max(0, ty - 1) + 1)
                                                for ty in range(-1, rph + 1)] * 3
                                # Measured in the kernel (scripts/gpu_rph_sweep.sh, stage-skip diagnostics): an
                                # SS-mode M = 128 x N x K = 16 MMA costs about 28 + 0.625 N clk -- the tensor
                                # time N/2 plus most of its operand fetch, which overlaps the math poorly -- i.e.
                                # 68 clk at N = 64, 108 at N = 128, 148 at N = 192.
                                tensor = tx * nchunks * ksteps * sum(28.0 + 0.625 * n for n in mma_list)
                                fill = nchunks * (rows * pitch * cb * 2 + 9 * bn * cb * 2)
                                epi_bytes = tx * rph * 128 * bn * 2 * (2 if sg == 64 else 0)
                                smem = (tx * nchunks * ksteps * sum(4096 + 32 * n for n in mma_list) + fill
                                        + epi_bytes) / 128.0
                                # one elected lane issues every TMA of a ring: ~520 clk per operation
                                wprod = nchunks * (9 // tps) * 520
                                hprod = nchunks * (3 if any_up else 1) * 520
                                traffic = fill / 75.0
                                epi = tx * rph * gpn * (300.0 if sg == 64 else 200.0)
                                per_item = max(tensor, smem, wprod, hprod, traffic) + 1200 + nchunks * 150
                                if acc_bufs == 1:
                                    per_item += 1.5 * epi
                                per_item = max(per_item, epi)
                                if w_slots * tps < 3:
                                    per_item *= 1.15
                                if out_stages == 1:
                                    per_item *= 1.01
                                total = -(-items // n_sms) * per_item
                                key = (total, -tx, -w_slots)
                                if best is None or key < best[0]:
                                    best = (key, cfg)
    if best is None:
        raise ValueError("no halo-kernel configuration for cout=%d cb=%d" % (cout, cb))
    cfg = best[1]
    cfg.update(cb=cb, nchunks=nchunks, cpad=cpad, loader=1 if cpl else 0)
    if force and "direct" in force:
        cfg["direct"] = force["direct"]
    return cfg


def halo_supported(sources, ksize, stride, pad):
    """3x3 / stride 1 / pad 1; an upsampled source additionally needs H % 16 == 0 (its 16 interior
    halo rows come from a tensor map that merges batch and rows)."""
    if not (ksize == 3 and stride == 1 and pad == 1):
        return False
    H = sources[0][0].shape[1] * sources[0][1]
    if any(up == 2 for _, up in sources) and H % 16:
        return False
    return all(t.shape[3] % 16 == 0 for t, _ in sources)


def halo_packed_weights_numel(cfg):
    return cfg["n_ntiles"] * cfg["nchunks"] * 9 * cfg["bn"] * cfg["cb"]


class HaloPlan:
    def __init__(self, desc, keep):
        self._keep = keep
        self.desc = desc
        h = C.c_void_p()
        _lib.check(_lib.lib().mmr_halo_conv_plan_create(C.byref(desc), C.byref(h)))
        self.handle = h
        self.flops = 0

    def run(self, stream=None, impl=0):
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _lib.check(_lib.lib().mmr_halo_conv_plan_run(self.handle, C.c_void_p(s)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().mmr_halo_conv_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def pack_weights_halo(w_oihw, cfg, mode, out=None, stream=None):
    """fp32 OIHW 3x3 weights -> the packed bf16 layout of cfg (mode 0 fprop, 1 dgrad)."""
    O, I = w_oihw.shape[0], w_oihw.shape[1]
    assert w_oihw.dtype == torch.float32 and w_oihw.is_contiguous() and tuple(w_oihw.shape[2:]) == (3, 3)
    if out is None:
        out = torch.empty((halo_packed_weights_numel(cfg),), device=w_oihw.device, dtype=torch.bfloat16)
    s = torch.cuda.current_stream().cuda_stream if stream is None else stream
    _lib.check(_lib.lib().mmr_pack_weights_halo(C.c_void_p(w_oihw.data_ptr()), O, I, mode, cfg["cb"], cfg["bn"],
                                                cfg["n_ntiles"], cfg["nchunks"], int(cfg.get("rph", 1) > 1),
                                                C.c_void_p(out.data_ptr()), C.c_void_p(s)))
    return out


def build_halo(cfg, sources, packed, groups, N, H, W, cout, *, scale=None, bias=None, residual=None,
               relu=False, out_f32=None, stats=None, stats_ld=0, bn_finalize=None, bn_bwd=None, head_metric=None):
    """sources: [(tensor [N,Hs,Ws,Cs] bf16, up)], packed: bf16 weights from pack_weights_halo,
    groups: [(dst tensor [N,H,W,ldc] bf16, coff)] one per store group of cfg['sg'] channels (bf16 NHWC
    mode), or out_f32 = fp32 [N,C,H,W] (head logits)."""
    from ._lib import MmrHaloConvDesc, MmrHaloSrc
    d = MmrHaloConvDesc()
    d.nsrc = len(sources)
    for i, (t, up) in enumerate(sources):
        assert t.dtype == torch.bfloat16 and t.is_contiguous()
        n_, h_, w_, c_ = t.shape
        assert (n_, h_ * up, w_ * up) == (N, H, W), (t.shape, up, N, H, W)
        d.src[i] = MmrHaloSrc(t.data_ptr(), c_, w_, h_, n_, up)
    d.N, d.H, d.W = N, H, W
    assert packed.numel() == halo_packed_weights_numel(cfg)
    d.weights = packed.data_ptr()
    d.cb, d.bn, d.sg, d.n_ntiles = cfg["cb"], cfg["bn"], cfg["sg"], cfg["n_ntiles"]
    d.tx, d.tps = cfg["tx"], cfg["tps"]
    d.rph = cfg.get("rph", 1)
    d.loader = cfg.get("loader", 0)
    d.halo_stages, d.w_slots, d.acc_bufs = cfg["halo_stages"], cfg["w_slots"], cfg["acc_bufs"]
    d.out_stages = max(1, cfg["out_stages"])
    d.direct_store = int(cfg.get("direct", cfg["sg"] < 64))
    keep = [sources, packed, scale, bias, residual, stats]
    if out_f32 is None and head_metric is None:
        # (tensor, coff) or (tensor, coff, step, oy, ox): pixel (y, x) -> (step*y + oy, step*x + ox) of the tensor
        groups = [g if len(g) == 5 else (g[0], g[1], 1, 0, 0) for g in groups]
        segs = (MmrOutSeg * len(groups))(*[MmrOutSeg(t.data_ptr(), t.shape[3], coff, step, oy, ox)
                                           for t, coff, step, oy, ox in groups])
        for t, _, step, _, _ in groups:    # step -2: 2x2 sum-pooled store into a half-resolution tensor
            want = (N, H // 2, W // 2) if step == -2 else (N, step * H, step * W)
            assert t.dtype == torch.bfloat16 and t.is_contiguous() and tuple(t.shape[:3]) == want
        groups = [(t, coff) for t, coff, _, _, _ in groups]
        d.ngroups = len(groups)
        d.groups = C.cast(segs, C.POINTER(MmrOutSeg))
        d.out_mode = MMR_OUT_BF16_NHWC
        keep += [segs, [t for t, _ in groups]]
    else:
        d.out_mode = MMR_OUT_F32_NCHW
        if out_f32 is not None:
            assert out_f32.dtype == torch.float32 and tuple(out_f32.shape) == (N, out_f32.shape[1], H, W)
            d.out_f32 = out_f32.data_ptr()
            d.out_ldc = out_f32.shape[1]
            keep.append(out_f32)
        if head_metric is not None:     # (labels int64 / uint8 [N,H,W] or None, pred uint8 [N,H,W] or None, cm int64 [N,C,C] or None)
            from ._lib import MmrHeadMetric
            labels, pred, cm = head_metric
            hm = MmrHeadMetric()
            if labels is not None:
                assert labels.dtype in (torch.int64, torch.uint8) and tuple(labels.shape) == (N, H, W) and labels.is_contiguous()
                hm.labels, hm.labels_u8 = labels.data_ptr(), int(labels.dtype == torch.uint8)
            if pred is not None:
                assert pred.dtype == torch.uint8 and tuple(pred.shape) == (N, H, W) and pred.is_contiguous()
                hm.pred_out = pred.data_ptr()
            if cm is not None:
                assert cm.dtype == torch.int64 and tuple(cm.shape) == (N, cout, cout) and cm.is_contiguous()
                hm.confusion = cm.data_ptr()
            d.head_metric = C.pointer(hm)
            keep += [hm, labels, pred, cm]
    d.cout_total = cout
    d.scale = scale.data_ptr() if scale is not None else None
    d.bias = bias.data_ptr() if bias is not None else None
    d.residual = residual.data_ptr() if residual is not None else None
    d.res_ldc = residual.shape[-1] if residual is not None else 0
    d.relu = int(bool(relu))
    d.stats = stats.data_ptr() if stats is not None else None
    d.stats_ld = stats_ld
    if bn_finalize is not None:
        d.bn_finalize = C.pointer(bn_finalize)
        keep.append(bn_finalize)
    if bn_bwd is not None:
        d.bn_bwd = C.pointer(bn_bwd)
        keep.append(bn_bwd)
    plan = HaloPlan(d, keep)
    plan.cfg = cfg
    return plan


def fprop_halo_cfg(sources, cout, bf16_out=True, force=None, stats=False):
    N = sources[0][0].shape[0]
    H = sources[0][0].shape[1] * sources[0][1]
    W = sources[0][0].shape[2] * sources[0][1]
    cin = sum(t.shape[3] for t, _ in sources)
    cb = pick_bk([t.shape[3] for t, _ in sources])
    any_up = any(up == 2 for _, up in sources)
    return halo_config(H, W, N, cb, cin // cb, cout, any_up, bf16_out=bf16_out, force=force, stats=stats)


def build_fprop_halo(sources, w_oihw, out, *, scale=None, bias=None, residual=None, relu=False,
                     out_f32=None, stats=None, stats_ld=0, force=None, packed=None, cfg=None, bn_finalize=None,
                     head_metric=None):
    """Forward 3x3 s1 p1 conv over the concatenation of `sources` (see build_fprop)."""
    N = sources[0][0].shape[0]
    H = sources[0][0].shape[1] * sources[0][1]
    W = sources[0][0].shape[2] * sources[0][1]
    cout, cin = w_oihw.shape[0], w_oihw.shape[1]
    # cin below the stored channel count: a zero-padded source (the 3-channel image stored as 16)
    assert cin <= sum(t.shape[3] for t, _ in sources)
    head = out_f32 is not None or head_metric is not None
    if cfg is None:
        cfg = fprop_halo_cfg(sources, cout, not head, force, stats=stats is not None)
    if packed is None:
        packed = pack_weights_halo(w_oihw, cfg, 0)
    groups = None
    if not head:
        assert out.shape[3] >= cfg["cpad"]
        groups = [(out, g * cfg["sg"]) for g in range(cfg["cpad"] // cfg["sg"])]
    plan = build_halo(cfg, sources, packed, groups, N, H, W, cout, scale=scale, bias=bias, residual=residual,
                      relu=relu, out_f32=out_f32, stats=stats, stats_ld=stats_ld, bn_finalize=bn_finalize,
                      head_metric=head_metric)
    plan.flops = 2 * N * H * W * cout * 9 * cin
    plan.packed = packed
    return plan


def dgrad_halo_cfg(dz_shape, sizes, force=None):
    N, H, W, Cz = dz_shape
    cb = pick_bk([Cz])
    return halo_config(H, W, N, cb, Cz // cb, sum(sizes), False, seg_sizes=list(sizes), force=force)


def build_dgrad_halo(dz, w_oihw, grads, *, force=None, packed=None, cfg=None, bn_bwd=None, pooled=None):
    """Data gradient of a 3x3 s1 p1 conv: dz [N,H,W,Cz] bf16 (Cz = Cout padded to 16), grads = one bf16
    tensor [N,H,W,Cs] per source in concat order.  pooled[i]: source i was read through nearest x2, its
    gradient tensor is [N,H/2,W/2,Cs] and receives the 2x2 sum of the conv-resolution gradient."""
    N, H, W, Cz = dz.shape
    cout, cin = w_oihw.shape[0], w_oihw.shape[1]
    assert Cz >= cout and Cz % 16 == 0
    sizes = [g.shape[3] for g in grads]
    assert sum(sizes) == cin
    pooled = pooled or [False] * len(grads)
    if cfg is None:
        cfg = dgrad_halo_cfg(dz.shape, sizes, force)
    if packed is None:
        packed = pack_weights_halo(w_oihw, cfg, 1)
    groups = []
    for g, pl in zip(grads, pooled):
        assert tuple(g.shape[1:3]) == ((H // 2, W // 2) if pl else (H, W))
        for c in range(0, g.shape[3], cfg["sg"]):
            groups.append((g, c, -2, 0, 0) if pl else (g, c))
    plan = build_halo(cfg, [(dz, 1)], packed, groups, N, H, W, cin, bn_bwd=bn_bwd)
    plan.flops = 2 * N * H * W * cout * 9 * cin
    plan.packed = packed
    return plan


class WgradHaloPlan:
    def __init__(self, desc, keep):
        self._keep = keep
        self.desc = desc
        h = C.c_void_p()
        _lib.check(_lib.lib().mmr_wgrad_halo_plan_create(C.byref(desc), C.byref(h)))
        self.handle = h
        self.flops = 0

    def run(self, stream=None, accumulate=False):
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _lib.check(_lib.lib().mmr_wgrad_halo_plan_run(self.handle, int(accumulate), C.c_void_p(s)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().mmr_wgrad_halo_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def wgrad_halo_config(H, W, N, cb, nchunks, cout_gemm, any_up, n_sms=148, force=None, max_partial=None,
                      dz_phased=False):
    """(mode, bn, tx, n_split) of a halo weight-gradient plan: slices = nchunks * cout_gemm/bn CTAs wide,
    split-K so that one wave of CTAs covers the SMs, tx as large as shared memory allows while
    keeping at least 3 pipeline stages.  mode 1 (64-channel chunks, bn 64 / 32): the filter column is a
    one-pixel shift of the dz tile, 2 MMAs of N = 3 bn per 16 pixels instead of 5 of N = bn."""
    bn = 64 if cout_gemm % 64 == 0 else (32 if cout_gemm % 32 == 0 else 16)
    n_nt = cout_gemm // bn
    A = 5 if cb == 64 else 3
    if os.environ.get("MMR_WGRAD_SMS"):   # A/B: CTAs of one launch (the side stream shares the SMs with the main one)
        n_sms = int(os.environ["MMR_WGRAD_SMS"])
    if cb == 64 and bn in (64, 32) and not dz_phased:
        modes = (1, 0)
    elif cb in (16, 32) and bn in (16, 32) and not dz_phased and W >= 16:
        modes = (2, 0)    # narrow layers: one MMA of N = 3 bn per 16 pixels, tiles gathered with cp.async
    else:
        modes = (0,)
    if os.environ.get("MMR_WGRAD_MODE", "") != "":   # A/B switch for measurements
        modes = tuple(m for m in modes if m == int(os.environ["MMR_WGRAD_MODE"])) or (0,)
    if force and "mode" in force:
        modes = tuple(m for m in modes if m == force["mode"])
    best = None
    for mode in modes:
        for tx in (4, 2, 1):
            if tx > 1 and 8 * tx > -(-W // 8) * 8:
                continue
            if mode == 2:
                if tx < 2:
                    continue
                stage = (-(-((16 + 128 // cb) * 8 * tx * cb * 2) // 1024) * 1024
                         + -(-(16 * (8 * tx + 2) * bn * 2) // 1024) * 1024)
            elif mode == 1:
                if tx > 2:
                    continue
                stage = -(-(19 * 8 * tx * 128) // 1024) * 1024 + -(-(16 * (8 * tx + 2) * bn * 2) // 1024) * 1024
            else:
                pitch = 8 * tx + (4 if any_up else 2)
                stage = -(-(18 * pitch * cb * 2 + 1024) // 1024) * 1024 + -(-(16 * 8 * tx * bn * 2) // 1024) * 1024
            stages = min(6, (224 * 1024) // stage)
            if stages < 2:
                continue
            tiles = N * (-(-H // 16)) * (-(-W // (8 * tx)))
            slices = nchunks * n_nt
            n_split = max(1, min(tiles, n_sms // slices)) if slices < n_sms else 1
            cfg = dict(mode=mode, bn=bn, tx=tx, n_split=n_split, stages=stages)
            if force and any(cfg.get(k, v) != v for k, v in force.items() if k != "n_split"):
                continue
            if force and "n_split" in force:
                cfg["n_split"] = n_split = max(1, min(tiles, force["n_split"]))
            # per-stage issue time vs. the ~2-4 TMA operations one lane issues per stage
            mma = tx * 8 * (_mma_clk(3 * bn) if mode == 2 else (2 * _mma_clk(3 * bn) if mode == 1 else A * _mma_clk(bn)))
            prod = (4 if any_up else 2) * 380
            if cb <= 32:   # narrow rows: TMA ~4 clk per pixel row (modes 0 / 1); the gather streams at the HBM share
                rows_px = (18 * (8 * tx + (0 if mode else 2)) + 16 * (8 * tx + (2 if mode else 0)))
                prod = rows_px * (cb + bn) / 24.0 if mode == 2 else rows_px * 4
            per_tile = max(mma, prod) / tx
            waves = -(-(slices * n_split) // n_sms)
            total = waves * (-(-tiles // n_split)) * tx * per_tile * (1.0 if stages >= 3 else 1.1)
            key = (total, -tx)
            if best is None or key < best[0]:
                best = (key, cfg)
    if best is None:
        raise ValueError("no halo wgrad configuration")
    cfg = best[1]
    cfg.update(cb=cb, nchunks=nchunks, n_ntiles=n_nt, A=A)
    # fp32 partials of one CTA
    cfg["per_cta"] = {2: 3 * cb * 3 * bn, 1: 192 * 3 * bn, 0: A * 128 * bn}[cfg["mode"]]
    if max_partial is not None:
        per_split = nchunks * n_nt * cfg["per_cta"]
        cfg["n_split"] = max(1, min(cfg["n_split"], max_partial // per_split))
    return cfg


def build_wgrad_halo(dz, sources, dst, *, cout_gemm=None, force=None, partial=None, n_sms=148, dz_phased=False):
    """Weight gradient of a 3x3 s1 p1 conv.  dz: [N,H,W,Cz] bf16; sources as in build_fprop_halo;
    dst: fp32 OIHW [Cout][Cin_total][3][3].  dz_phased: dz is [N,2H,2W,C] and the GEMM's output channel (q, c)
    is channel c of pixel (2y + qy, 2x + qx) (the space-to-depth stem): Cout = 4 C."""
    from ._lib import MmrHaloSrc, MmrWgradHaloDesc
    N, H, W, Cz = dz.shape
    if dz_phased:
        H, W, Cz = H // 2, W // 2, 4 * Cz
        force = dict(force or {}, bn=dz.shape[3])
    assert dz.dtype == torch.bfloat16 and dz.is_contiguous()
    assert dst.dtype == torch.float32 and dst.is_contiguous() and tuple(dst.shape[2:]) == (3, 3)
    cout, cin = dst.shape[0], dst.shape[1]
    cin_stored = sum(t.shape[3] for t, _ in sources)
    assert cin <= cin_stored
    if cout_gemm is None:
        cout_gemm = Cz
    cb = pick_bk([t.shape[3] for t, _ in sources])
    any_up = any(up == 2 for _, up in sources)
    cfg = wgrad_halo_config(H, W, N, cb, cin_stored // cb, cout_gemm, any_up, n_sms=n_sms, force=force,
                            max_partial=partial.numel() if partial is not None else None, dz_phased=dz_phased)
    need = cfg["nchunks"] * cfg["n_ntiles"] * cfg["n_split"] * cfg["per_cta"]
    if partial is None:
        partial = torch.empty((need,), device=dz.device, dtype=torch.float32)
    assert partial.dtype == torch.float32 and partial.numel() >= need
    d = MmrWgradHaloDesc()
    d.dz = MmrHaloSrc(dz.data_ptr(), dz.shape[3], W, H, N, 2 if dz_phased else 1)
    d.nsrc = len(sources)
    for i, (t, up) in enumerate(sources):
        assert t.dtype == torch.bfloat16 and t.is_contiguous()
        n_, h_, w_, c_ = t.shape
        assert (n_, h_ * up, w_ * up) == (N, H, W)
        d.src[i] = MmrHaloSrc(t.data_ptr(), c_, w_, h_, n_, up)
    d.N, d.H, d.W = N, H, W
    d.cb, d.bn, d.cout_gemm, d.tx, d.n_split = cb, cfg["bn"], cout_gemm, cfg["tx"], cfg["n_split"]
    d.partial = partial.data_ptr()
    d.dst = dst.data_ptr()
    d.dst_cout, d.dst_cin = cout, cin
    d.mode = cfg["mode"]
    plan = WgradHaloPlan(d, [dz, sources, dst, partial])
    plan.cfg = cfg
    plan.flops = 2 * N * H * W * cout * 9 * cin
    return plan
