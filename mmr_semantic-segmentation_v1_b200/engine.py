"""Static execution plan of one network: every buffer, tensor map and launch argument is fixed
at build time for a given (batch, height, width); running a step is replaying lists of C-ABI
launches on one CUDA stream (capturable in a CUDA graph, no allocation, no host sync).

Forward (training):  conv (tcgen05, raw bf16 output z) -> batch statistics -> y = relu(z*scale +
shift [+ residual]).  Forward (eval): BatchNorm folded into the conv epilogue, one launch per conv.
Backward: per unit  g = relu-mask * sum(consumer contributions)  (+ BN reductions) -> dz -> wgrad
(fp32 OIHW into the flat gradient buffer) and dgrad (one bf16 tensor per source; a nearest-x2
source receives it at the conv's resolution and its own gather 2x2-pools it).

Replaces the autograd graph PyTorch builds for `seg = model(img)` / `loss.backward()`
(SU/ModelTraining.py:589,614; ED/Main_MMR_SegModel.py:697,715).
"""
import ctypes as C
import os

import torch

from . import _lib, convplan
from ._lib import MMR_OUT_BF16_NHWC, MMR_OUT_F32_NCHW, MmrContrib

STEM_KPAD = 160  # 7*7*3 = 147 im2col columns padded to a multiple of 32
BN_BLOCKS = 296  # CTAs of the per-channel reductions: 2 per SM, all co-resident (one wave)
HALO_STAT_SLOTS = 8  # statistics slots the halo conv kernel accumulates into (csrc/conv_halo.cu)
WG_PARTIAL_FLOATS = 48 * 1024 * 1024  # 192 MB of fp32 split-K partials


class _Act:
    def __init__(self, name, shape, needs_grad=True):
        self.name, self.shape, self.needs_grad = name, shape, needs_grad
        self.buf = None
        self.contribs = []   # (buffer key, pool2) filled while building the backward plan
        self.producer = None


class _PointwisePlan:
    """Stands where a conv plan would for the launches of csrc/pointwise_head.cu (no handle: CUDA-core kernels)."""
    handle = None
    flops = 0


class _Arena:
    """Liveness-planned scratch for backward temporaries (g, dz, per-source data gradients)."""

    def __init__(self):
        self.req = []  # (key, nbytes, t_alloc, t_free)

    def request(self, key, nbytes, t_alloc, t_free):
        if os.environ.get("MMR_NO_ARENA_REUSE"):
            t_free = 1 << 30
        self.req.append((key, (nbytes + 1023) // 1024 * 1024, t_alloc, t_free))

    def plan(self):
        free, live, offsets, top = [], [], {}, 0
        for key, nbytes, t0, t1 in sorted(self.req, key=lambda r: r[2]):
            for item in [l for l in live if l[0] < t0]:
                live.remove(item)
                free.append((item[1], item[2]))
            free.sort()
            merged = []
            for off, sz in free:
                if merged and merged[-1][0] + merged[-1][1] == off:
                    merged[-1] = (merged[-1][0], merged[-1][1] + sz)
                else:
                    merged.append((off, sz))
            free = merged
            if free and free[-1][0] + free[-1][1] == top:   # trailing hole: give it back
                top = free[-1][0]
                free.pop()
            got = None
            for i, (off, sz) in enumerate(free):
                if sz >= nbytes:
                    got = off
                    if sz > nbytes:
                        free[i] = (off + nbytes, sz - nbytes)
                    else:
                        free.pop(i)
                    break
            if got is None:
                got = top
                top += nbytes
            offsets[key] = got
            live.append((t1, got, nbytes))
            self.peak = max(getattr(self, "peak", 0), top)
        return offsets, getattr(self, "peak", 0)


class Engine:
    def __init__(self, ops, params, grads, N, H, W, device, training=True, n_sms=None):
        """params: name -> fp32 tensor (weights, biases, BN affine and running buffers);
        grads: name -> fp32 tensor receiving the gradient of the parameter of that name."""
        self.lib = _lib.lib()
        if not _lib.device_ok():
            raise _lib.MmrError("mmrseg_b200 needs a CUDA device of compute capability 10.x "
                                "(B200); there is no CPU or PyTorch fallback")
        self.ops, self.P, self.G = ops, params, grads
        self.N, self.H, self.W = N, H, W
        self.dev = device
        self.training = training
        self.n_sms = n_sms or torch.cuda.get_device_properties(device).multi_processor_count
        self.keep = []
        self.acts = {}
        self.units = []
        self.fwd_calls = []
        self.repack_calls = []
        self.bwd_calls = {False: [], True: []}
        self.conv_flops_fwd = 0
        self.conv_flops_bwd = 0
        self.n_launch_fwd = 0
        self.n_launch_bwd = 0
        self.param_ready_hooks = []  # (index in bwd call list, [param names]) for DDP overlap
        self.generation = 0          # forwards run so far: the saved activations belong to the latest one
        self.use_halo = not os.environ.get("MMR_NO_HALO")
        self.use_graphs = not os.environ.get("MMR_NO_GRAPH")
        self.use_lanes = not os.environ.get("MMR_NO_LANES")
        self._side = None
        self._graphs = {}
        self._train_calls = {}
        self._fwd_u8 = None
        # BatchNorm statistics taken in the conv epilogues: [unit][slot][2][C] doubles, bump-allocated
        self.halo_stats = torch.zeros((HALO_STAT_SLOTS * 2 * 16384,), device=device, dtype=torch.float64)
        self.halo_stats_used = 0
        self.pack_jobs = []
        # tickets of the fused BatchNorm finalisations (one per launch site) and the shared slots of
        # the backward reductions; both are re-armed by the kernels themselves
        self.tickets = torch.zeros((4096,), device=device, dtype=torch.int32)
        self.tickets_used = 0
        self.bwd_slots = torch.zeros((8 * 2 * 2048,), device=device, dtype=torch.float64)
        self.fuse_finalize = not os.environ.get("MMR_NO_FUSED_FINALIZE")
        self.x_in = torch.empty((N, 3, H, W), device=device, dtype=torch.float32)
        # uint8 HWC frames straight from the loader: /255 and the optional
        # (x - mean) / std of utils.normalize happen in the kernels that read the image
        self.x_u8 = torch.empty((N, H, W, 3), device=device, dtype=torch.uint8)
        self.in_norm = torch.tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]], device=device, dtype=torch.float32)
        self.u8_swaps = []           # (index in fwd_calls, replacement call) of the image readers
        self.bn_partial = torch.empty((BN_BLOCKS * 2 * 512,), device=device, dtype=torch.float64)
        # split-K partials of the weight-gradient GEMMs: one buffer, used by one layer at a time
        self.wg_partial = torch.empty((WG_PARTIAL_FLOATS,), device=device, dtype=torch.float32) \
            if training else None
        self._build_forward()
        if training:
            self._build_backward()

    # ------------------------------------------------------------------ helpers
    def _bf16(self, *shape, zero=False):
        f = torch.zeros if zero else torch.empty
        return f(shape, device=self.dev, dtype=torch.bfloat16)

    def _f32(self, *shape, zero=False):
        f = torch.zeros if zero else torch.empty
        return f(shape, device=self.dev, dtype=torch.float32)

    def _rec(self, calls, fname, *args):
        conv = []
        for a in args:
            if isinstance(a, torch.Tensor):
                self.keep.append(a)
                conv.append(C.c_void_p(a.data_ptr()))
            else:
                conv.append(a)
        calls.append((getattr(self.lib, fname), tuple(conv)))

    def _pack_job(self, w, out, O, I, mode, cfg):
        """Queue one layer's fp32 OIHW -> packed bf16 conversion; all jobs run as ONE launch."""
        self.pack_jobs.append(_lib.MmrPackJob(w.data_ptr(), out.data_ptr(), O, I, mode, cfg["cb"], cfg["bn"],
                                              cfg["n_ntiles"], cfg["nchunks"], int(cfg.get("rph", 1) > 1)))
        self.keep += [w, out]

    def _ticket(self):
        t = self.tickets[self.tickets_used:self.tickets_used + 1]
        self.tickets_used += 1
        assert self.tickets_used <= self.tickets.numel()
        return t

    def _nblk(self, P, Cc):
        rows_per_iter = 256 // (Cc // 8)
        k = int(os.environ.get("MMR_BN_ITERS_PER_CTA", "4"))   # measurement knob: row iterations per CTA at least
        return int(max(1, min(BN_BLOCKS, -(-P // (rows_per_iter * k)))))

    # ------------------------------------------------------------------ forward plan
    def _build_forward(self):
        N, H, W = self.N, self.H, self.W
        fc = self.fwd_calls
        self.acts["image"] = _Act("image", (N, H, W, 3), needs_grad=False)
        for op in self.ops:
            kind = op["op"]
            if kind == "stem":
                self._fwd_stem(op)
            elif kind == "maxpool":
                src = self.acts[op["in"]]
                n, h, w, c = src.shape
                k = op.get("k", 3)   # 3: MaxPool2d(3, 2, 1) of the ResNet stem; 2: MaxPool2d(2) of the in-tree UNet
                oshape = (n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c) if k == 3 else (n, h // 2, w // 2, c)
                out = _Act(op["out"], oshape)
                out.buf = self._bf16(*out.shape)
                # the argmax positions are only read by the backward pass
                out.idx = torch.empty(out.shape, device=self.dev, dtype=torch.uint8) if self.training else None
                out.producer = {"kind": "maxpool", "src": src, "out": out, "k": k}
                self.acts[op["out"]] = out
                self.units.append(out.producer)
                self._rec(fc, "mmr_maxpool3x3s2_fwd" if k == 3 else "mmr_maxpool2x2s2_fwd", src.buf, n, h, w, c,
                          out.buf, out.idx)
            elif kind in ("conv", "head"):
                self._fwd_conv(op)
            elif kind == "convt":
                self._fwd_convt(op)
            elif kind == "pack16":   # image -> NHWC bf16, 3 channels zero-padded to 16
                act = _Act(op["out"], (N, H, W, 16), needs_grad=False)
                act.buf = self._bf16(N, H, W, 16)
                self.acts[op["out"]] = act
                self._rec(fc, "mmr_pack_nchw_f32_to_nhwc_bf16", self.x_in, N, 3, H, W, act.buf, 16)
                alt = []
                self._rec(alt, "mmr_pack_nhwc_u8_to_nhwc_bf16", self.x_u8, N, H, W, act.buf, 16, self.in_norm[0],
                          self.in_norm[1])
                self.u8_swaps.append((len(fc) - 1, alt[0]))
            elif kind == "upsample":
                src = self.acts[op["in"]]
                n, h, w, c = src.shape
                out = _Act(op["out"], (n, 2 * h, 2 * w, c))
                out.buf = self._bf16(*out.shape)
                out.producer = {"kind": "upsample", "src": src, "out": out}
                self.acts[op["out"]] = out
                self.units.append(out.producer)
                self._rec(fc, "mmr_upsample_bilinear2x_fwd", src.buf, n, h, w, c, out.buf)
            elif kind == "input":    # test seam: an activation fed from outside (bf16 NHWC)
                h, w, c = op["shape"]
                act = _Act(op["out"], (N, h, w, c))
                act.buf = self._bf16(N, h, w, c, zero=True)
                act.producer = {"kind": "input", "out": act, "grad": self._bf16(N, h, w, c, zero=True)}
                self.acts[op["out"]] = act
                self.units.append(act.producer)
            elif kind == "output":   # test seam: an activation whose gradient is seeded from outside
                act = self.acts[op["in"]]
                self.out_seeds = getattr(self, "out_seeds", {})
                self.out_seeds[op["in"]] = self._bf16(*act.shape, zero=True)
            else:
                raise ValueError("unknown op %r" % kind)
        if self.pack_jobs:
            arr = (_lib.MmrPackJob * len(self.pack_jobs))(*self.pack_jobs)
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
            self.pack_jobs_dev = raw.to(self.dev)
            per = self.lib.mmr_pack_items_per_block()
            blocks = sum(-(-(j.n_ntiles * j.nchunks * j.bn * j.cb) // per) for j in self.pack_jobs)
            self.repack_calls.append((self.lib.mmr_pack_weights_halo_batch,
                                      (C.c_void_p(self.pack_jobs_dev.data_ptr()), len(self.pack_jobs),
                                       C.c_int64(blocks))))
        if getattr(self, "folded", None):
            jobs = []
            for u in self.folded:
                bn = u["bn"]
                cb = self.P[u["fold_bias"]].data_ptr() if u.get("fold_bias") else None
                jobs.append(_lib.MmrBnFoldJob(
                    self.P[bn + ".weight"].data_ptr(), self.P[bn + ".bias"].data_ptr(),
                    self.P[bn + ".running_mean"].data_ptr(), self.P[bn + ".running_var"].data_ptr(), cb,
                    u["scale"].data_ptr(), u["shift"].data_ptr(), self.P[bn + ".weight"].numel(),
                    u.get("fold_rep", 1)))
            arr = (_lib.MmrBnFoldJob * len(jobs))(*jobs)
            self.fold_jobs_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone().to(self.dev)
            self.repack_calls.append((self.lib.mmr_bn_fold_batch,
                                      (C.c_void_p(self.fold_jobs_dev.data_ptr()), len(jobs), C.c_float(1e-5))))
        if self.halo_stats_used:
            # one memset per forward re-arms every statistics slot the conv epilogues accumulate into
            fc.insert(0, (self.lib.mmr_zero_async, (C.c_void_p(self.halo_stats.data_ptr()),
                                                     C.c_int64(self.halo_stats_used * 8))))
            self.u8_swaps = [(i + 1, call) for i, call in self.u8_swaps]
            if getattr(self, "metric_swap", None):
                self.metric_swap = (self.metric_swap[0] + 1, self.metric_swap[1])
        self.n_launch_fwd = len(fc) + len(self.repack_calls)

    def _bn_state(self, unit, bn_name, Cc):
        st = self._f32(7, Cc)  # mean, invstd, scale, shift, coefA, coefB, coefC
        unit.update(bn=bn_name, mean=st[0], invstd=st[1], scale=st[2], shift=st[3], coef=st[4:7])
        self.keep.append(st)

    def _fwd_bn_train(self, unit, z, out, residual, relu, stats=None, finalized=False):
        """stats: per-channel sums already left by the conv epilogue ([8][2][C] doubles), else a
        separate reduction pass over z."""
        fc = self.fwd_calls
        bn = unit["bn"]
        Cc = z.shape[-1]
        Pn = z.numel() // Cc
        nblk = self._nblk(Pn, Cc)
        partial = self.bn_partial
        if stats is not None:
            partial, nblk = stats, HALO_STAT_SLOTS
        else:
            self._rec(fc, "mmr_bn_stats", z, Pn, Cc, self.bn_partial, nblk)
        if not finalized:   # else the conv kernel's last CTA already wrote mean / invstd / scale / shift
            self._rec(fc, "mmr_bn_finalize", partial, nblk, Pn, Cc, self.P[bn + ".weight"],
                      self.P[bn + ".bias"], C.c_float(1e-5), C.c_float(0.1), self.P[bn + ".running_mean"],
                      self.P[bn + ".running_var"], self.P[bn + ".num_batches_tracked"], unit["mean"],
                      unit["invstd"], unit["scale"], unit["shift"])
        self._rec(fc, "mmr_bn_apply", z, Pn, Cc, unit["scale"], unit["shift"], residual, int(relu), out)

    def _fwd_stem(self, op):
        N, H, W = self.N, self.H, self.W
        Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        Pn = N * Ho * Wo
        fc = self.fwd_calls
        w = self.P[op["conv"] + ".weight"]
        cout = w.shape[0]
        unit = {"kind": "stem", "op": op, "cout": cout}
        out = _Act(op["out"], (N, Ho, Wo, cout))
        out.buf = self._bf16(*out.shape)
        out.producer = unit
        unit["out"] = out
        if cout == 64 and H % 4 == 0 and W % 4 == 0 and not os.environ.get("MMR_NO_S2D"):
            self._fwd_stem_s2d(op, unit, out, w)
        else:
            unit["mat"] = self._bf16(1, 1, Pn, STEM_KPAD)
            unit["wf"] = self._bf16(cout, STEM_KPAD, zero=True)
            self._rec(fc, "mmr_stem_im2col", self.x_in, N, H, W, unit["mat"], STEM_KPAD, None, None)
            alt = []
            self._rec(alt, "mmr_stem_im2col_u8", self.x_u8, N, H, W, unit["mat"], STEM_KPAD, self.in_norm[0],
                      self.in_norm[1])
            self.u8_swaps.append((len(fc) - 1, alt[0]))
            self._rec(self.repack_calls, "mmr_repack_weights", w, cout, 147, 1, unit["wf"], STEM_KPAD, None, 0, 0)
            if self.training:
                unit["z"] = self._bf16(N, Ho, Wo, cout)
                self._bn_state(unit, op["bn"], cout)
                plan = convplan.build_fprop([(unit["mat"], 1)], unit["wf"], 1, 1, 0, unit["z"].view(1, 1, Pn, cout))
                fc.append((self.lib.mmr_conv_plan_run, (plan.handle, 0)))
                self._fwd_bn_train(unit, unit["z"], out.buf, None, True)
            else:
                self._fold(unit, op["bn"], cout)
                plan = convplan.build_fprop([(unit["mat"], 1)], unit["wf"], 1, 1, 0, out.buf.view(1, 1, Pn, cout),
                                            scale=unit["scale"], bias=unit["shift"], relu=True)
                fc.append((self.lib.mmr_conv_plan_run, (plan.handle, 0)))
            unit["fplan"] = plan
        unit["fplan"].flops = 2 * Pn * cout * 147
        self.conv_flops_fwd += unit["fplan"].flops
        self.acts[op["out"]] = out
        self.units.append(unit)

    def _fwd_stem_s2d(self, op, unit, out, w):
        """The 7x7 stride-2 stem as a 3x3 convolution over 4x4 pixel blocks (include/mmrseg.h,
        mmr_stem_s2d_*) on the halo kernel: 48 block channels -> 4 output phases x 64 channels, each phase
        stored by its own strided TMA map (depth-to-space for free); no im2col matrix."""
        N, H, W = self.N, self.H, self.W
        fc = self.fwd_calls
        cout = unit["cout"]
        s2d = self._bf16(N, H // 4, W // 4, 64)
        self._rec(fc, "mmr_stem_s2d_pack", self.x_in, 0, N, H, W, s2d, None, None)
        alt = []
        self._rec(alt, "mmr_stem_s2d_pack", self.x_u8, 1, N, H, W, s2d, self.in_norm[0], self.in_norm[1])
        self.u8_swaps.append((len(fc) - 1, alt[0]))
        w3 = self._f32(4 * cout, 64, 3, 3)
        self._rec(self.repack_calls, "mmr_stem_s2d_weights", w, cout, w3)
        sources = [(s2d, 1)]
        hcfg = convplan.fprop_halo_cfg(sources, 4 * cout, bf16_out=True, force={"rph": 1, "sg": 64})
        assert hcfg["sg"] == cout == 64
        unit["wf_h"] = self._bf16(convplan.halo_packed_weights_numel(hcfg))
        self._pack_job(w3, unit["wf_h"], 4 * cout, 64, 0, hcfg)
        unit["s2d"] = s2d
        if self.training:    # raw z through the strided groups, then the usual statistics / apply passes
            unit["z"] = self._bf16(*out.shape)
            self._bn_state(unit, op["bn"], cout)
            groups = [(unit["z"], 0, 2, q >> 1, q & 1) for q in range(4)]
            plan = convplan.build_halo(hcfg, sources, unit["wf_h"], groups, N, H // 4, W // 4, 4 * cout)
        else:
            self._fold(unit, op["bn"], cout, rep=4)
            groups = [(out.buf, 0, 2, q >> 1, q & 1) for q in range(4)]
            plan = convplan.build_halo(hcfg, sources, unit["wf_h"], groups, N, H // 4, W // 4, 4 * cout,
                                       scale=unit["scale"], bias=unit["shift"], relu=True)
        fc.append((self.lib.mmr_halo_conv_plan_run, (plan.handle,)))
        unit["fplan"] = plan
        if self.training:
            self._fwd_bn_train(unit, unit["z"], out.buf, None, True)

    def _fwd_convt(self, op):
        """nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2) with bias (SU/UArchModel/unet_parts.py:269):
        out[n, 2y+ky, 2x+kx, co] = b[co] + sum_ci x[n, y, x, ci] * W[ci, co, ky, kx], i.e. exactly the data
        gradient of V = Conv2d(cout -> cin, k 2, stride 2) whose OIHW weight IS the ConvTranspose2d weight
        ([cin][cout][2][2]).  So the forward runs as V's dgrad plan (four output parities, one 1x1 GEMM each,
        bias in the epilogue), the data gradient as V's fprop plan, the weight gradient as V's wgrad plan with
        the roles of activation and gradient swapped -- all on the first-generation tcgen05 kernels."""
        fc = self.fwd_calls
        src = self.acts[op["in"]]
        n, h, w_, cin = src.shape
        wt = self.P[op["conv"] + ".weight"]        # [cin][cout][2][2]
        assert tuple(wt.shape) == (cin, op["cout"], 2, 2), (op["conv"], tuple(wt.shape), cin)
        cout = op["cout"]
        assert cin % 16 == 0 and cout % 16 == 0
        out = _Act(op["out"], (n, 2 * h, 2 * w_, cout))
        out.buf = self._bf16(*out.shape)
        unit = {"kind": "convt", "op": dict(op, relu=False, bias=True), "cout": cout, "cpad": cout, "k": 2, "s": 2,
                "pad": 0, "srcs": [(src, 1)], "in_hw": (h, w_), "out": out, "res": None, "halo": False}
        # V's GEMM layouts: fprop rows = V's outputs (cin of the ConvT), dgrad rows = V's inputs (cout of the ConvT)
        unit["wf"] = self._bf16(cin, 4 * cout, zero=True)
        unit["wd"] = self._bf16(cout, 4 * cin, zero=True)
        self._rec(self.repack_calls, "mmr_repack_weights", wt, cin, cout, 4, unit["wf"], 4 * cout, unit["wd"],
                  4 * cin, cin)
        plan = convplan.build_dgrad(src.buf, unit["wd"], 2, 2, 0, (2 * h, 2 * w_), [out.buf],
                                    bias=self.P[op["conv"] + ".bias"])
        plan.flops = 2 * n * h * w_ * cin * cout * 4
        fc.append((self.lib.mmr_conv_plan_run, (plan.handle, 0)))
        self.conv_flops_fwd += plan.flops
        unit["fplan"] = plan
        out.producer = unit
        self.acts[op["out"]] = out
        self.units.append(unit)

    def _fold(self, unit, bn_name, Cc, conv_bias=None, rep=1):
        """Eval mode: scale/shift from the running statistics, recomputed by ONE `mmr_bn_fold_batch` launch at the
        head of every eval forward (so they always follow the current parameters / running statistics: an
        optimiser step, a training forward or load_state_dict in between needs no notification); a conv bias
        in front of the BatchNorm (the in-tree UNet's DoubleConv) folds into the shift.  rep: the GEMM's
        output channels are `rep` copies of the layer's (the space-to-depth stem's four output phases)."""
        st = self._f32(2, Cc * rep)
        unit.update(bn=bn_name, scale=st[0], shift=st[1], fold_bias=conv_bias, fold_rep=rep)
        self.keep.append(st)
        self.folded = getattr(self, "folded", [])
        self.folded.append(unit)

    def _fwd_conv(self, op):
        fc = self.fwd_calls
        head = op["op"] == "head"
        k = op["k"]
        s = 1 if head else op["s"]
        pad = k // 2
        srcs = [(self.acts[name], up) for name, up in op["src"]]
        n, h0, w0, _ = srcs[0][0].shape
        Hin, Win = h0 * srcs[0][1], w0 * srcs[0][1]
        Ho, Wo = (Hin + 2 * pad - k) // s + 1, (Win + 2 * pad - k) // s + 1
        w = self.P[op["conv"] + ".weight"]
        cout, cin = w.shape[0], w.shape[1]
        if not head and op.get("cin_pad"):   # image padded to 16 channels: the extra ones meet zero weights
            assert len(srcs) == 1 and cin <= srcs[0][0].shape[3], (op["conv"], cin)
        else:
            assert cin == sum(a.shape[3] for a, _ in srcs), (op["conv"], cin)
        cpad = -(-cout // 16) * 16
        taps = k * k
        unit = {"kind": "head" if head else "conv", "op": op, "cout": cout, "cpad": cpad, "k": k, "s": s,
                "pad": pad, "srcs": srcs, "in_hw": (Hin, Win)}
        sources = [(a.buf, up) for a, up in srcs]
        # 1x1 head over a 64-channel map at its own resolution (ResNetUNet.conv_last): the CUDA-core fp32 kernels of
        # csrc/pointwise_head.cu (HBM-bound; one backward pass also applies the producer's ReLU mask)
        pw = bool(head and k == 1 and len(srcs) == 1 and srcs[0][1] == 1 and cin == 64 and cout <= 16
                  and op.get("up", 1) == 1 and not os.environ.get("MMR_NO_POINTWISE_HEAD"))
        unit["pw"] = pw
        halo = self.use_halo and not pw and convplan.halo_supported(sources, k, s, pad)
        unit["halo"] = halo
        res = self.acts[op["res"]] if (not head and op.get("res")) else None
        unit["res"] = res
        bn_train = bool(not head and op.get("bn") and self.training)
        if halo:
            # statistics of a biased conv go through the generic epilogue, which has no row-phase stacking
            hcfg = convplan.fprop_halo_cfg(sources, cout, bf16_out=not head, stats=bn_train,
                                           force={"rph": 1} if (bn_train and op.get("bias")) else None)
            unit["wf_h"] = self._bf16(convplan.halo_packed_weights_numel(hcfg))
            self._pack_job(w, unit["wf_h"], cout, cin, 0, hcfg)
            if self.training:
                dsizes = [a.shape[3] for a, _ in srcs]
                dcfg = convplan.dgrad_halo_cfg((n, Hin, Win, cpad), dsizes)
                unit["dcfg"] = dcfg
                unit["wd_h"] = self._bf16(convplan.halo_packed_weights_numel(dcfg))
                self._pack_job(w, unit["wd_h"], cout, cin, 1, dcfg)
        elif not pw:
            unit["wf"] = self._bf16(cpad, taps * cin, zero=True)
            unit["wd"] = self._bf16(cin, taps * cpad, zero=True) if self.training else None
            self._rec(self.repack_calls, "mmr_repack_weights", w, cout, cin, taps, unit["wf"], taps * cin,
                      unit["wd"], taps * cpad, cpad)

        def fprop(dst, **kw):
            if pw:
                self._rec(fc, "mmr_pointwise_head_fwd", srcs[0][0].buf, w, kw["bias"], n, Ho, Wo, cin, cout, dst)
                return _PointwisePlan()
            if halo:
                plan = convplan.build_fprop_halo(sources, w, dst if not head else None, cfg=hcfg,
                                                 packed=unit["wf_h"], out_f32=dst if head else None, **kw)
                fc.append((self.lib.mmr_halo_conv_plan_run, (plan.handle,)))
            else:
                kw.pop("stats", None)
                kw.pop("stats_ld", None)
                kw.pop("bn_finalize", None)
                if head:
                    plan = convplan.build_fprop(sources, unit["wf"], k, 1, pad, dst, out_mode=MMR_OUT_F32_NCHW,
                                                cout=cout, bn=cpad, **kw)
                else:
                    plan = convplan.build_fprop(sources, unit["wf"], k, s, pad, dst, **kw)
                fc.append((self.lib.mmr_conv_plan_run, (plan.handle, 0)))
            return plan

        if head:
            out = _Act(op["out"], (n, Ho, Wo, cout))
            out.buf = self._f32(n, cout, Ho, Wo)  # NCHW fp32 logits
            plan = fprop(out.buf, bias=self.P[op["conv"] + ".bias"])
            if halo and not self.training and op["out"] == "logits" and cout <= 16 and hcfg["bn"] == 16:
                # eval: the same head with argmax + confusion matrix in its epilogue and NO logits store
                # (`forward(..., metric=True)` swaps this launch in)
                self.metric_labels = torch.zeros((n, Ho, Wo), device=self.dev, dtype=torch.int64)
                self.metric_pred = torch.empty((n, Ho, Wo), device=self.dev, dtype=torch.uint8)
                self.metric_cm = torch.zeros((n, cout, cout), device=self.dev, dtype=torch.int64)
                try:
                    mplan = convplan.build_fprop_halo(sources, w, None, cfg=hcfg, packed=unit["wf_h"],
                                                      bias=self.P[op["conv"] + ".bias"],
                                                      head_metric=(self.metric_labels, self.metric_pred, self.metric_cm))
                    unit["mplan"] = mplan
                    self.metric_swap = (len(fc) - 1, (self.lib.mmr_halo_conv_plan_run, (mplan.handle,)))
                except _lib.MmrError:      # 14-16 classes need 1 KB more shared memory than this plan has left:
                    self.metric_swap = None  # model.segment() then takes forward() + the metric kernel
            up = op.get("up", 1)
            unit["up"] = up
            unit["result"] = out.buf
            if up > 1 and self.training:   # auxiliary (deep-supervision) head: logits at full resolution
                unit["result"] = self._f32(n, cout, Ho * up, Wo * up)
                self._rec(fc, "mmr_upsample_nearest_f32_nchw", out.buf, C.c_int64(n * cout), Ho, Wo, up,
                          unit["result"])
        else:
            out = _Act(op["out"], (n, Ho, Wo, cout))
            out.buf = self._bf16(*out.shape)
            bias = self.P[op["conv"] + ".bias"] if op.get("bias") else None
            if bn_train:
                unit["z"] = self._bf16(*out.shape)
                self._bn_state(unit, op["bn"], cout)
                stats = None
                if halo:
                    nst = HALO_STAT_SLOTS * 2 * cout
                    assert self.halo_stats_used + nst <= self.halo_stats.numel()
                    stats = self.halo_stats[self.halo_stats_used:self.halo_stats_used + nst]
                    self.halo_stats_used += nst
                bnf = None
                if stats is not None and self.fuse_finalize:
                    bn = op["bn"]
                    bnf = _lib.MmrBnFinalize(
                        self.P[bn + ".weight"].data_ptr(), self.P[bn + ".bias"].data_ptr(), 1e-5, 0.1,
                        self.P[bn + ".running_mean"].data_ptr(), self.P[bn + ".running_var"].data_ptr(),
                        self.P[bn + ".num_batches_tracked"].data_ptr(), unit["mean"].data_ptr(),
                        unit["invstd"].data_ptr(), unit["scale"].data_ptr(), unit["shift"].data_ptr(),
                        n * Ho * Wo, self._ticket().data_ptr())
                plan = fprop(unit["z"], bias=bias, stats=stats, stats_ld=cout, bn_finalize=bnf)
                self._fwd_bn_train(unit, unit["z"], out.buf, res.buf if res else None, op["relu"], stats,
                                   finalized=bnf is not None)
            else:
                scale = shift = None
                if op.get("bn"):
                    self._fold(unit, op["bn"], cout, op["conv"] + ".bias" if bias is not None else None)
                    scale, shift = unit["scale"], unit["shift"]
                else:
                    shift = bias
                plan = fprop(out.buf, scale=scale, bias=shift, residual=res.buf if res else None,
                             relu=op["relu"])
        plan.flops = 2 * n * Ho * Wo * cout * taps * cin
        self.conv_flops_fwd += plan.flops
        if not pw:      # "fplan" marks the tensor-core launches (bench roofline, per-layer tables)
            unit["fplan"] = plan
        unit["out"] = out
        out.producer = unit
        self.acts[op["out"]] = out
        self.units.append(unit)

    # ------------------------------------------------------------------ backward plan
    def _build_backward(self):
        # how many ops read each activation (sources, residuals, pools, test seeds): a unit whose output has
        # exactly one reader can have its BatchNorm-backward sums taken by that reader's data-gradient launch
        self.n_readers = {}
        for op in self.ops:
            for name in [n for n, _ in op.get("src", [])] + [op.get("res"), op.get("in")]:
                if name is not None:
                    self.n_readers[name] = self.n_readers.get(name, 0) + 1
        for name in getattr(self, "out_seeds", {}):
            self.n_readers[name] = self.n_readers.get(name, 0) + 1
        self.fuse_bwd_reduce = self.fuse_finalize and not os.environ.get("MMR_NO_FUSED_BWD_REDUCE")
        # channel counts of the units whose reduction rides on the consumer's data gradient (MMR_FUSED_BWD_CH: A/B)
        self.fuse_bwd_channels = tuple(int(v) for v in os.environ.get("MMR_FUSED_BWD_CH", "64").split(",") if v)
        self.wgrad_late = os.environ.get("MMR_WGRAD_LATE", "0") == "1"
        # the main stream waits for the weight gradient of unit t when it reaches unit t + K (and recycles that
        # unit's dz only then): a deeper K lets the side stream fall further behind instead of stalling the chain
        K = self.wgrad_join = max(2, int(os.environ.get("MMR_WGRAD_JOIN", "2")))
        self.pool_dgrad = os.environ.get("MMR_POOL_DGRAD", "1") == "1"   # A/B switch
        self.fused_dgrad_handles = set()   # data-gradient plans that also take a BatchNorm-backward reduction
        order = list(reversed(self.units))
        t_of = {id(u): t for t, u in enumerate(order)}
        arena = _Arena()
        bpe = 2

        def nbytes(shape):
            n = bpe
            for d in shape:
                n *= d
            return n

        # 1. liveness of every temporary
        for t, u in enumerate(order):
            kind = u["kind"]
            if kind == "input":
                continue
            if kind in ("maxpool", "upsample"):
                src = u["src"]
                if src.needs_grad:
                    arena.request(("gin", id(u)), nbytes(src.shape), t, t_of[id(src.producer)])
                continue
            out = u["out"]
            oshape = out.shape[:3] + (u.get("cpad", u["cout"]),)
            t_g = t
            if u.get("res") is not None and u["res"].needs_grad and u["res"].producer is not None:
                t_g = t_of[id(u["res"].producer)]
            if u.get("pw"):
                # pointwise head: no bf16 copy of dlogits; when its source is a single-reader conv + bias + ReLU the
                # backward kernel writes that producer's dz itself (ReLU mask applied), which therefore lives from here
                L = self._pw_fuse_target(u)
                if L is not None:
                    L["dz_by_reader"] = t
                    continue
            elif u.get("dz_by_reader") is not None:
                arena.request(("g", id(u)), nbytes(oshape), u["dz_by_reader"], max(t_g, t + K - 1))
            else:
                # the weight gradient runs on the side stream and is joined K units later: its dz operand
                # (g itself when there is no BatchNorm) must not be recycled before that
                arena.request(("g", id(u)), nbytes(oshape), t, t_g if u.get("bn") else max(t_g, t + K - 1))
            if u.get("bn"):
                arena.request(("dz", id(u)), nbytes(oshape), t, t + K - 1)
            if kind == "stem":
                continue
            Hin, Win = u["in_hw"]
            for si, (a, up) in enumerate(u["srcs"]):
                if a.needs_grad and a.producer is not None:
                    arena.request(("dx", id(u), si), nbytes((a.shape[0], Hin, Win, a.shape[3])), t,
                                  t_of[id(a.producer)])
        offsets, peak = arena.plan()
        self.arena_bytes = peak
        self.arena = torch.empty((max(peak, 1024),), device=self.dev, dtype=torch.uint8)

        def view(key, shape):
            off = offsets[key]
            return self.arena[off:off + nbytes(shape)].view(torch.bfloat16).view(shape)

        self.arena_view = view      # tests read the backward temporaries through it (MMR_NO_ARENA_REUSE=1)
        self.arena_keys = set(offsets)

        # 2. emit launches in backward order; contributions are registered as we go
        for acc in (False, True):
            for a in self.acts.values():
                a.contribs = []
            for name, seed in getattr(self, "out_seeds", {}).items():
                self.acts[name].contribs.append((seed, 0))
            calls = self.bwd_calls[acc]
            for t, u in enumerate(order):
                calls.append((Engine._mark, ("join", t - self.wgrad_join)))
                self._bwd_t = t
                self._bwd_unit(u, calls, view, int(acc), record_hooks=not acc)
        self.n_launch_bwd = len(self.bwd_calls[False])

    def _pw_fuse_target(self, u):
        """The conv + bias + ReLU unit (no BatchNorm, no residual) whose 64-channel output only the pointwise head
        `u` reads: the head's backward kernel then writes that unit's dz and bias gradient.  None otherwise."""
        src = u["srcs"][0][0]
        L = src.producer
        if (L is not None and L.get("kind") == "conv" and not L["op"].get("bn") and L["op"]["relu"]
                and L.get("res") is None and self.n_readers.get(src.name, 0) == 1
                and L["cout"] == src.shape[3] == L.get("cpad", L["cout"]) and src.needs_grad):
            return L
        return None

    def _contrib_array(self, act):
        arr = (MmrContrib * max(1, len(act.contribs)))()
        for i, (buf, pool2) in enumerate(act.contribs):
            arr[i].ptr = buf.data_ptr()
            arr[i].pool2 = pool2
            self.keep.append(buf)
        self.keep.append(arr)
        return arr, len(act.contribs)

    def _bwd_unit(self, u, calls, view, acc, record_hooks):
        kind = u["kind"]
        if kind == "input":
            out = u["out"]
            n, h, w, c = out.shape
            arr, cnt = self._contrib_array(out)
            self._rec(calls, "mmr_grad_gather", arr, cnt, None, n, h, w, c, u["grad"], None, self._nblk(n * h * w, c))
            return
        if kind == "upsample":
            src, out = u["src"], u["out"]
            if not src.needs_grad:
                return
            n, h, w, c = src.shape
            gin = view(("gin", id(u)), src.shape)
            arr, cnt = self._contrib_array(out)
            self._rec(calls, "mmr_upsample_bilinear2x_bwd", arr, cnt, n, h, w, c, gin)
            src.contribs.append((gin, 0))
            return
        if kind == "maxpool":
            src, out = u["src"], u["out"]
            if not src.needs_grad:
                return
            n, h, w, c = src.shape
            gin = view(("gin", id(u)), src.shape)
            arr, cnt = self._contrib_array(out)
            self._rec(calls, "mmr_maxpool3x3s2_bwd" if u.get("k", 3) == 3 else "mmr_maxpool2x2s2_bwd", arr, cnt,
                      out.idx, n, h, w, c, gin)
            src.contribs.append((gin, 0))
            return
        out = u["out"]
        n, ho, wo, _ = out.shape
        cpad = u.get("cpad", u["cout"])
        oshape = (n, ho, wo, cpad)
        Pn = n * ho * wo
        conv = u["op"]["conv"]
        if kind == "head" and u.get("pw"):
            # the loss leaves fp32 NCHW dlogits in u["dlogits"]: one pass gives the head's weight / bias gradients and
            # the source's data gradient (csrc/pointwise_head.cu)
            if "dlogits" not in u:
                u["dlogits"] = self._f32(n, u["cout"], ho, wo, zero=True)
            src = u["srcs"][0][0]
            L = self._pw_fuse_target(u)
            if "pw_ws" not in u:
                u["pw_ws"] = self._f32(int(self.lib.mmr_pointwise_head_bwd_workspace_bytes(u["cout"])) // 4)
            names = [conv + ".weight", conv + ".bias"]
            if L is not None:
                Lconv = L["op"]["conv"]
                dx = view(("g", id(L)), L["out"].shape)
                dbl = self.G[Lconv + ".bias"] if L["op"].get("bias") else None
                if dbl is not None:
                    names.append(Lconv + ".bias")
            else:
                dx, dbl = view(("dx", id(u), 0), src.shape), None
                src.contribs.append((dx, 0))
            self._rec(calls, "mmr_pointwise_head_bwd", u["dlogits"], src.buf, self.P[conv + ".weight"], n, ho, wo,
                      src.shape[3], u["cout"], int(L is not None), dx, self.G[conv + ".weight"],
                      self.G[conv + ".bias"], dbl, acc, u["pw_ws"])
            if record_hooks:
                self.conv_flops_bwd += 2 * (2 * Pn * u["cout"] * src.shape[3])
                self.param_ready_hooks.append((len(calls), names))
            return
        g = view(("g", id(u)), oshape)
        if kind == "head":
            # the loss leaves fp32 NCHW dlogits in u["dlogits"]; convert + bias gradient
            if "dlogits" not in u:
                u["dlogits"] = self._f32(n, u["cout"], ho, wo, zero=True)
                if u.get("up", 1) > 1:
                    u["dlogits_full"] = self._f32(n, u["cout"], ho * u["up"], wo * u["up"], zero=True)
            if u.get("up", 1) > 1:
                self._rec(calls, "mmr_sumpool_f32_nchw", u["dlogits_full"], C.c_int64(n * u["cout"]), ho, wo,
                          u["up"], u["dlogits"])
            self._rec(calls, "mmr_head_grad_prep", u["dlogits"], n, u["cout"], ho, wo, g, cpad,
                      self.G[conv + ".bias"], acc)
            dz = g
        else:
            arr, cnt = self._contrib_array(out)
            relu_act = out.buf if u["op"]["relu"] else None
            Cc = u["cout"]
            nblk = self._nblk(Pn, Cc)
            if u.get("bn"):
                bn = u["bn"]
                dz = view(("dz", id(u)), oshape)
                if self.fuse_finalize and Cc <= 2048:
                    if "bwd_ticket" not in u:
                        u["bwd_ticket"] = self._ticket()
                    # without a residual the ReLU mask is a function of z alone: do not re-read the activation
                    zmask = relu_act is not None and u.get("res") is None
                    # one full-resolution contribution: the apply pass recomputes the masked gradient from it,
                    # the reduction writes no g at all
                    nog = zmask and cnt == 1 and not out.contribs[0][1] and not os.environ.get("MMR_NO_NOG")
                    if u.get("reduce_done_by_dgrad"):
                        assert nog    # the reader's data-gradient launch already left dgamma / dbeta / coef
                    else:
                        self._rec(calls, "mmr_bn_bwd_reduce_fused", arr, cnt, None if zmask else relu_act, u["z"],
                                  u["mean"], u["invstd"], n, ho, wo, Cc, None if nog else g, self.bwd_slots, nblk,
                                  self.P[bn + ".weight"],
                                  self.G[bn + ".weight"], self.G[bn + ".bias"], acc, u["coef"], u["bwd_ticket"],
                                  u["scale"] if zmask else None, u["shift"] if zmask else None)
                else:
                    nog = False
                    self._rec(calls, "mmr_bn_bwd_reduce", arr, cnt, relu_act, u["z"], u["mean"], u["invstd"],
                              n, ho, wo, Cc, g, self.bn_partial, nblk)
                    self._rec(calls, "mmr_bn_bwd_finalize", self.bn_partial, nblk, Pn, Cc,
                              self.P[bn + ".weight"], u["invstd"], self.G[bn + ".weight"],
                              self.G[bn + ".bias"], acc, u["coef"])
                if nog:
                    self._rec(calls, "mmr_bn_bwd_apply_masked", out.contribs[0][0], u["z"], u["mean"], u["invstd"],
                              u["coef"], u["scale"], u["shift"], Pn, Cc, dz)
                else:
                    self._rec(calls, "mmr_bn_bwd_apply", g, u["z"], u["mean"], u["invstd"], u["coef"], Pn, Cc, dz)
            elif u.get("dz_by_reader") is not None:
                dz = g      # written, masked, by the pointwise head's backward pass (bias gradient included)
            else:
                has_bias = u["op"].get("bias")
                self._rec(calls, "mmr_grad_gather", arr, cnt, relu_act, n, ho, wo, Cc, g,
                          self.bn_partial if has_bias else None, nblk)
                if has_bias:
                    self._rec(calls, "mmr_bias_grad_finalize", self.bn_partial, nblk, Cc,
                              self.G[conv + ".bias"], acc)
                dz = g
            if u.get("res") is not None and u["res"].needs_grad:
                u["res"].contribs.append((g, 0))
        # weight gradient
        gw = self.G[conv + ".weight"]
        wplan = u.get("wplan")
        if wplan is None:
            if kind == "stem" and "s2d" in u:
                # weight gradient of the 3x3 space-to-depth form (dz read phase by phase through strided maps),
                # folded back onto the 7x7 filter afterwards
                u["dw3"] = self._f32(4 * u["cout"], 64, 3, 3)
                wplan = convplan.build_wgrad_halo(dz, [(u["s2d"], 1)], u["dw3"], n_sms=self.n_sms,
                                                  partial=self.wg_partial, dz_phased=True)
                wplan.flops = 2 * Pn * u["cout"] * 147
            elif kind == "stem":
                wplan = convplan.build_wgrad(dz.view(1, 1, Pn, cpad), [(u["mat"], 1)], 1, 1, 0,
                                             gw.view(gw.shape[0], 147, 1, 1), n_sms=self.n_sms,
                                             partial=self.wg_partial)
                wplan.flops = 2 * Pn * u["cout"] * 147
            elif kind == "convt":
                # dW[ci][co][ky][kx] = sum x[ci] * g[co, 2y+ky, 2x+kx]: V's weight gradient with V's "dz" = the
                # ConvT input and V's source = the ConvT output gradient
                src_a = u["srcs"][0][0]
                cin_t = src_a.shape[3]
                u["wplans"] = [convplan.build_wgrad(src_a.buf, [(dz, 1)], 2, 2, 0, gw[c0:c0 + 512], dz_c0=c0,
                                                    cout_gemm=min(512, cin_t - c0), n_sms=self.n_sms,
                                                    partial=self.wg_partial)
                               for c0 in range(0, cin_t, 512)]      # a plan covers <= 512 GEMM columns
                wplan = u["wplans"][0]
                wplan.flops = u["fplan"].flops
            else:
                sources = [(a.buf, up) for a, up in u["srcs"]]
                if u.get("halo"):
                    wplan = convplan.build_wgrad_halo(dz, sources, gw, cout_gemm=cpad, n_sms=self.n_sms,
                                                      partial=self.wg_partial)
                else:
                    wplan = convplan.build_wgrad(dz, sources, u["k"], u["s"], u["pad"], gw, cout_gemm=cpad,
                                                 n_sms=self.n_sms, partial=self.wg_partial)
            u["wplan"] = wplan
        wcalls = [(Engine._mark, ("side_begin", self._bwd_t))]
        if kind == "stem" and "s2d" in u:
            wcalls.append((self.lib.mmr_wgrad_halo_plan_run, (wplan.handle, 0)))
            self._rec(wcalls, "mmr_stem_s2d_wgrad_fold", u["dw3"], u["cout"], gw, acc)
        elif isinstance(wplan, convplan.WgradHaloPlan):
            wcalls.append((self.lib.mmr_wgrad_halo_plan_run, (wplan.handle, acc)))
        else:
            for wp in u.get("wplans", [wplan]):
                wcalls.append((self.lib.mmr_wgrad_plan_run, (wp.handle, 0, acc)))
        wcalls.append((Engine._mark, ("side_end", self._bwd_t)))

        def emit_wgrad():
            calls.extend(wcalls)
            if record_hooks:
                self.conv_flops_bwd += wplan.flops
                names = [conv + ".weight"]
                if u.get("bn"):
                    names += [u["bn"] + ".weight", u["bn"] + ".bias"]
                if kind == "head" or u["op"].get("bias"):
                    names.append(conv + ".bias")
                self.param_ready_hooks.append((len(calls), names))

        # data gradient, one tensor per source that needs it
        need = [] if kind == "stem" else [a.needs_grad and a.producer is not None for a, _ in u["srcs"]]
        # MMR_WGRAD_LATE: fork the weight gradient AFTER the unit's data gradient is enqueued, so that it runs
        # beside the next unit's BatchNorm-backward passes (HBM-bound) instead of beside its own data gradient
        # (both tensor-bound, one CTA per SM each)
        late = self.wgrad_late and any(need)
        if not late:
            emit_wgrad()
        if not any(need):
            return
        Hin, Win = u["in_hw"]
        grads = []
        # a nearest-x2 source gets its gradient already 2x2 sum-pooled by the data-gradient epilogue (halo kernel,
        # plain epilogue): a quarter of the bytes written here and read by the source's BatchNorm-backward passes
        plain_epi = u.get("halo") and int(u["dcfg"].get("direct", u["dcfg"]["sg"] < 64)) == int(u["dcfg"]["sg"] < 64)
        pool_dx = [bool(up == 2 and plain_epi and self.pool_dgrad and Hin % 2 == 0 and Win % 2 == 0)
                   for _, up in u["srcs"]]
        for si, (a, up) in enumerate(u["srcs"]):
            shape = (a.shape[0], Hin // 2, Win // 2, a.shape[3]) if pool_dx[si] else (a.shape[0], Hin, Win, a.shape[3])
            if need[si]:
                grads.append(view(("dx", id(u), si), shape))
            else:
                grads.append(None)
        if not all(need):
            raise NotImplementedError("mixed grad / no-grad sources in one conv")
        # the only reader of a conv -> BN -> ReLU unit's output takes that unit's BatchNorm-backward sums in its
        # own epilogue (MmrBnBwdFused): one store group that is the whole tensor, same resolution, no residual
        bb = None
        src_a, src_up = u["srcs"][0]
        L = src_a.producer
        if (self.fuse_bwd_reduce and u.get("halo") and len(u["srcs"]) == 1 and src_up == 1 and L is not None
                and L.get("kind") == "conv" and L.get("bn") and L["op"]["relu"] and L.get("res") is None
                and self.n_readers.get(src_a.name, 0) == 1 and L["cout"] == src_a.shape[3]
                and L["cout"] in self.fuse_bwd_channels and u["dcfg"]["n_ntiles"] == 1 and u["dcfg"]["bn"] == u["dcfg"]["sg"]
                and int(u["dcfg"].get("direct", u["dcfg"]["sg"] < 64)) == int(u["dcfg"]["sg"] < 64)):
            bnL = L["bn"]
            if "bwd_ticket" not in L:
                L["bwd_ticket"] = self._ticket()
            bb = _lib.MmrBnBwdFused(
                L["z"].data_ptr(), L["scale"].data_ptr(), L["shift"].data_ptr(), L["mean"].data_ptr(),
                L["invstd"].data_ptr(), self.P[bnL + ".weight"].data_ptr(), self.G[bnL + ".weight"].data_ptr(),
                self.G[bnL + ".bias"].data_ptr(), L["coef"].data_ptr(), self.bwd_slots.data_ptr(),
                L["bwd_ticket"].data_ptr(), src_a.shape[0] * Hin * Win, acc)
            L["reduce_done_by_dgrad"] = True
        key = ("dplan", acc if bb is not None else None)
        dplan = u.get(key)
        if dplan is None:
            if u.get("halo"):
                dplan = convplan.build_dgrad_halo(dz, self.P[conv + ".weight"], grads, cfg=u["dcfg"],
                                                  packed=u["wd_h"], bn_bwd=bb, pooled=pool_dx)
            elif kind == "convt":   # dx = V(g): a plain 2x2 stride-2 conv of the output gradient
                dplan = convplan.build_fprop([(dz, 1)], u["wf"], 2, 2, 0, grads[0])
                dplan.flops = u["fplan"].flops
            else:
                dplan = convplan.build_dgrad(dz, u["wd"], u["k"], u["s"], u["pad"], (Hin, Win), grads)
            u[key] = dplan
            u["dplan"] = dplan
            if bb is not None:
                self.fused_dgrad_handles.add(dplan.handle.value)
        if u.get("halo"):
            calls.append((self.lib.mmr_halo_conv_plan_run, (dplan.handle,)))
        else:
            calls.append((self.lib.mmr_conv_plan_run, (dplan.handle, 0)))
        if record_hooks:
            self.conv_flops_bwd += dplan.flops
        for si, (a, up) in enumerate(u["srcs"]):
            a.contribs.append((grads[si], 1 if (up == 2 and not pool_dx[si]) else 0))
        if late:
            emit_wgrad()

    # ------------------------------------------------------------------ execution
    def _run(self, calls, stream, lo=0, hi=None):
        """Replay calls[lo:hi].  On torch's current stream (stream None) a segment is launched eagerly
        the first time (kernel attributes, lazily allocated workspaces), captured into a CUDA graph the
        second time and replayed from then on: every argument of a plan is fixed at build time, so the
        graph stays valid for the life of the engine and the ~300 launches of a step cost one."""
        hi = len(calls) if hi is None else hi
        if hi <= lo:
            return
        if stream is None and self.use_graphs:
            key = (id(calls), lo, hi)
            g = self._graphs.get(key)
            if g is None:
                self._graphs[key] = False          # seen once: capture on the next use
            elif g is False:
                g = torch.cuda.CUDAGraph()
                # MMR_MAIN_PRIO=1 (A/B): capture on a high-priority stream, so the kernel nodes of the main chain carry a
                # higher launch priority than the weight-gradient branch (side stream, default priority)
                cap = None
                if os.environ.get("MMR_MAIN_PRIO"):
                    cap = self._cap_stream = getattr(self, "_cap_stream", None) or torch.cuda.Stream(device=self.dev, priority=-1)
                with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
                    self._launch(calls, torch.cuda.current_stream().cuda_stream, lo, hi, lanes=True)
                self._graphs[key] = g
                g.replay()
                return
            else:
                g.replay()
                return
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        self._launch(calls, st, lo, hi, lanes=stream is None)

    @staticmethod
    def _mark(kind, key, stream=None):
        """Scheduling marker inside a launch list (a no-op when a list is replayed serially):
        ("side_begin", k) .. ("side_end", k) bracket launches that may run on the side stream once the
        main stream has reached the marker; ("join", k) makes the main stream wait for bracket k."""
        return 0

    def _launch(self, calls, stream, lo, hi, lanes=False):
        """lanes: `stream` is torch's current stream; bracketed launches (the weight-gradient GEMMs, which
        nothing on the backward critical path waits for) go to a second stream, forked and joined with
        events -- inside a capture this becomes a parallel branch of the CUDA graph."""
        err = 0
        if not (lanes and self.use_lanes):
            s = C.c_void_p(stream)
            for fn, args in calls[lo:hi]:
                err |= fn(*args, s)
        else:
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.dev)
            side = self._side
            cur, done, pending = main, {}, None
            for fn, args in calls[lo:hi]:
                if fn is Engine._mark:
                    kind, key = args
                    if kind == "side_begin":
                        ev = torch.cuda.Event()
                        ev.record(main)
                        side.wait_event(ev)
                        cur = side
                    elif kind == "side_end":
                        ev = torch.cuda.Event()
                        ev.record(side)
                        done[key] = pending = ev
                        cur = main
                    elif kind == "join" and key in done:
                        main.wait_event(done.pop(key))
                    continue
                err |= fn(*args, C.c_void_p(cur.cuda_stream))
            if pending is not None:      # the segment ends with everything back on the main stream
                main.wait_event(pending)
        if err:
            raise _lib.MmrError(self.lib.mmr_last_error().decode(errors="replace"))

    def forward(self, x=None, stream=None, metric=False):
        """x: fp32 NCHW [N,3,H,W] on the device (copied into the plan's input buffer), or uint8 NHWC
        [N,H,W,3] frames (normalised by `set_input_norm` constants inside the first kernels).
        metric (eval engines with a fused head metric): run the head with argmax + confusion matrix in its
        epilogue instead of storing logits; the caller fills `metric_labels` first and reads `metric_pred` /
        `metric_cm` (re-zeroed by this call) afterwards; returns None."""
        u8 = x is not None and x.dtype == torch.uint8
        if u8:
            self.x_u8.copy_(x, non_blocking=True)
            if self._fwd_u8 is None:
                self._fwd_u8 = list(self.fwd_calls)
                for i, call in self.u8_swaps:
                    self._fwd_u8[i] = call
        elif x is not None:
            self.x_in.copy_(x, non_blocking=True)
        fwd = self._fwd_u8 if u8 else self.fwd_calls
        if metric:
            key = ("u8" if u8 else "f32") + "+metric"
            if key not in self._train_calls:
                calls = list(fwd)
                i, call = self.metric_swap
                calls[i] = call
                zero = (self.lib.mmr_zero_async, (C.c_void_p(self.metric_cm.data_ptr()),
                                                  C.c_int64(self.metric_cm.numel() * 8)))
                self._train_calls[key] = self.repack_calls + [zero] + calls
            self._run(self._train_calls[key], stream)
            self.generation += 1
            return None
        # Both modes re-derive the bf16 GEMM weights (and, in eval mode, the folded BatchNorm constants) from
        # the fp32 masters at the head of every forward: the optimiser kernels, the training engine's running
        # statistics and load_state_dict all write those buffers without telling this engine (two launches,
        # ~40 us, inside the same CUDA graph as the forward).
        key = "u8" if u8 else "f32"
        if key not in self._train_calls:
            self._train_calls[key] = self.repack_calls + fwd
        self._run(self._train_calls[key], stream)
        self.generation += 1
        heads = self.head_units()
        if self.training and len(heads) > 1:
            return [u["result"] for u in heads]
        return self.acts["logits"].buf if "logits" in self.acts else None

    def set_input_norm(self, mean, std):
        """(x/255 - mean) / std for uint8 frames (utils.normalize, SU/utils.py:480-519)."""
        self.in_norm[0].copy_(torch.as_tensor(mean, dtype=torch.float32))
        self.in_norm[1].copy_(torch.as_tensor(std, dtype=torch.float32))

    def head_units(self):
        return [u for u in self.units if u["kind"] == "head"]

    def backward(self, dlogits=None, accumulate=False, stream=None, on_ready=None, cuts=None):
        """dlogits: fp32 NCHW gradient of the main head (copied in), or None when a loss kernel
        already wrote into `dlogits_buffer()`.  on_ready(param_names) is called (host side, in
        launch order) after the launches that complete those parameters' gradients."""
        if dlogits is not None:
            heads = self.head_units()
            grads = list(dlogits) if isinstance(dlogits, (list, tuple)) else [dlogits]
            assert len(grads) == len(heads), "one gradient per head output"
            for u, g in zip(heads, grads):
                dst = u["dlogits_full"] if u.get("up", 1) > 1 else u["dlogits"]
                if g is None:
                    dst.zero_()
                else:
                    dst.copy_(g, non_blocking=True)
        calls = self.bwd_calls[bool(accumulate)]
        if on_ready is None:
            self._run(calls, stream)
            return
        # cuts: the hook positions after which the caller wants control (DDP: where a gradient bucket
        # is complete); the hooks in between are coalesced so that a segment is one graph launch
        lo, names = 0, []
        for hi, ns in self.param_ready_hooks:
            names += ns
            if cuts is not None and hi not in cuts:
                continue
            self._run(calls, stream, lo, hi)
            on_ready(names)
            lo, names = hi, []
        self._run(calls, stream, lo, len(calls))
        if names:
            on_ready(names)

    def dlogits_buffer(self, which=0):
        return self.head_units()[which]["dlogits"]
