"""Optimiser seam: torch.optim.Adam(lr, weight_decay) (SU/ModelTraining.py:366), AdamW
(:369, ED/Main_MMR_SegModel.py:878-880), SGD(momentum=0.9) (SU/ModelTraining.py:372,381),
clip_grad_norm_ (ED/...:718-727) on the flat fp32 parameter / gradient buffers of a plan model:
HBM-bound launches over contiguous memory (Adam: 28 B per parameter).

State layout is torch's: `state[p] = {"step", "exp_avg", "exp_avg_sq"}` (Adam) / `{"momentum_buffer"}` (SGD),
keyed per parameter, so `optimizer.state_dict()` / `load_state_dict()` round-trip and a checkpoint written
by the reference's torch.optim.Adam resumes here (ED/Main_MMR_SegModel.py:991).  The moment tensors are views
of one flat buffer laid out like the model's flat parameter buffer, which is what lets a param group be
stepped with ONE launch per contiguous run of its parameters (the whole model: one launch).
"""
import torch

from . import _lib, ops  # noqa: F401  (ops registers the mmrseg:: custom ops)
from .losses import _stream


def _runs(params, slots):
    """Split `params` (each with a gradient) into maximal runs that are contiguous, in the same order and at the
    same relative offsets, in ALL of: parameter storage, gradient storage and every state slot in `slots`
    (name -> {param: tensor}).  Returns [(indices into params, [start pointers], numel including alignment
    padding between members, all-members-launchable-as-one)].
    Only the parameters handed in are ever covered by a run: a group holding a subset of a model's flat
    buffer steps that subset and nothing else."""
    order = sorted(range(len(params)), key=lambda i: params[i].data_ptr())
    runs, cur = [], None
    for i in order:
        p = params[i]
        ptrs = [p.data_ptr(), p.grad.data_ptr()] + [slots[k][p].data_ptr() for k in sorted(slots)]
        ok = p.is_contiguous() and p.grad.is_contiguous() and p.dtype == torch.float32 and \
            p.grad.dtype == torch.float32 and all(q % 16 == 0 for q in ptrs)
        if cur is not None and ok and cur["ok"]:
            gap = ptrs[0] - cur["end"][0]
            # the flat layout pads every tensor to a multiple of 4 floats; pad elements are zero in p, g, m, v
            if 0 <= gap < 16 and all(q - e == gap for q, e in zip(ptrs, cur["end"])):
                cur["idx"].append(i)
                cur["end"] = [q + p.numel() * 4 for q in ptrs]
                continue
        cur = {"idx": [i], "start": ptrs, "end": [q + p.numel() * 4 for q in ptrs], "ok": ok}
        runs.append(cur)
    return [(r["idx"], r["start"], (r["end"][0] - r["start"][0]) // 4, r["ok"]) for r in runs]


def _span(t, numel):
    """fp32 tensor over `numel` consecutive elements of t's storage starting at t's first element (a run of
    parameters that are views of one flat buffer, including the alignment padding between them)."""
    return torch.empty(0, dtype=t.dtype, device=t.device).set_(t.untyped_storage(), t.storage_offset(), (numel,))


class _FlatStateOptimizer(torch.optim.Optimizer):
    """Shared machinery: per-parameter torch-layout state whose tensors are views of flat buffers that mirror
    the parameters' own storage layout."""

    _slots = ()

    def __init__(self, params, defaults):
        super().__init__(params, defaults)
        self.grad_scale = 1.0
        self._flat_state = {}     # (storage ptr, slot) -> flat fp32 tensor spanning that storage
        self._plans = {}

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._plans = {}          # loaded tensors are fresh allocations: re-home them on the next step
        for g in self.param_groups:
            for k, v in self.defaults.items():
                g.setdefault(k, v)

    def _slot_view(self, p, slot):
        """View for parameter p inside the flat `slot` buffer that mirrors p's storage."""
        st = p.untyped_storage()
        key = (st.data_ptr(), slot)
        flat = self._flat_state.get(key)
        if flat is None or flat.numel() * 4 != st.nbytes() or flat.device != p.device:
            flat = torch.zeros(st.nbytes() // 4, dtype=torch.float32, device=p.device)
            self._flat_state[key] = flat
        off = p.storage_offset()
        return flat[off:off + p.numel()].view(p.shape)

    def _home(self, p, slot, init=None):
        """state[p][slot] as a view of the flat buffer; a tensor that came from load_state_dict (or an
        existing one after the parameters moved) is copied in."""
        st = self.state[p]
        cur = st.get(slot)
        if p.dtype != torch.float32 or not p.is_contiguous():
            if cur is None:
                st[slot] = cur = torch.zeros_like(p, memory_format=torch.preserve_format)
            return cur
        view = self._slot_view(p, slot)
        if cur is None:
            if init is not None:
                view.copy_(init)
            else:
                view.zero_()
        elif cur.data_ptr() != view.data_ptr():
            view.copy_(cur.to(view.dtype))
        st[slot] = view
        return view

    def _plan(self, gi, params):
        """Cached run decomposition of group gi, keyed by what it depends on (parameter, gradient and state
        addresses), so the per-step host cost is one tuple comparison."""
        sig = tuple((p.data_ptr(), p.grad.data_ptr()) for p in params)
        plan = self._plans.get(gi)
        if plan is not None and plan[0] == sig:
            return plan[1]
        slots = {s: {p: self._home(p, s) for p in params} for s in self._slots}
        runs = []
        for idx, ptrs, numel, ok in _runs(params, slots):
            first = min((params[i] for i in idx), key=lambda p: p.data_ptr())
            # the tensors the custom op is launched on: one span per run for (p, g, state slots...)
            spans = [_span(first.data, numel), _span(first.grad, numel)] + \
                    [_span(slots[s][first], numel) for s in sorted(slots)] if ok else None
            runs.append((idx, ptrs, numel, ok, spans))
        self._plans[gi] = (sig, runs)
        return runs


class FusedAdam(_FlatStateOptimizer):
    """Adam (L2 folded into the gradient, like torch.optim.Adam) or AdamW (decoupled=True)."""

    _slots = ("exp_avg", "exp_avg_sq")

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                      decoupled=decoupled))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = _lib.lib()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            b1, b2 = group["betas"]
            runs = self._plan(gi, params)
            mode = 1 if group.get("decoupled", self.defaults["decoupled"]) else 0
            for idx, ptrs, numel, ok, spans in runs:
                # torch keeps one step counter per parameter; the parameters of a run advance together
                steps = set()
                for i in idx:
                    st = self.state[params[i]]
                    if "step" in st:
                        st["step"] += 1
                    else:
                        st["step"] = torch.tensor(1.0, dtype=torch.float32)
                    steps.add(float(st["step"]))
                if len(steps) == 1 and ok:
                    jobs = [(spans, steps.pop())]
                else:
                    jobs = []
                    for i in idx:
                        p = params[i]
                        st = self.state[p]
                        if not (p.grad.is_contiguous() and p.grad.dtype == torch.float32 and p.is_contiguous()
                                and p.dtype == torch.float32):
                            raise _lib.MmrError("FusedAdam needs contiguous fp32 parameters and gradients")
                        jobs.append(([p.data, p.grad, st["exp_avg"], st["exp_avg_sq"]], float(st["step"])))
                for (pt, gt, mt, vt), t in jobs:
                    if any(q.data_ptr() % 16 for q in (pt, gt, mt, vt)):
                        raise _lib.MmrError("FusedAdam: a parameter, gradient or moment buffer is not 16-byte "
                                            "aligned (parameters of mmrseg_b200 models always are)")
                    torch.ops.mmrseg.adam_step(pt, gt, mt, vt, group["lr"], b1, b2, group["eps"],
                                               group["weight_decay"], t, bool(mode), self.grad_scale)
        return loss


class FusedSGD(_FlatStateOptimizer):
    """torch.optim.SGD(lr, momentum, weight_decay) (dampening 0, no Nesterov), the reference's
    `optim.SGD(model.parameters(), lr=args.lr, momentum=0.9)`; per-group learning rates (the differential
    encoder / decoder rates of SU/ModelTraining.py:375-383) run one launch per contiguous run of a group."""

    _slots = ("momentum_buffer",)

    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = _lib.lib()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            # torch initialises the momentum buffer with the first gradient; a zero-initialised buffer gives
            # the same first step (momentum * 0 + g), so the kernel's first_step flag stays 0
            for idx, ptrs, numel, ok, spans in self._plan(gi, params):
                if ok:
                    jobs = [spans]
                else:
                    jobs = [[params[i].data, params[i].grad, self.state[params[i]]["momentum_buffer"]] for i in idx]
                for pt, gt, bt in jobs:
                    torch.ops.mmrseg.sgd_step(pt, gt, bt, group["lr"], group["momentum"], group["weight_decay"],
                                              self.grad_scale)
        return loss


def clip_grad_norm_(parameters, max_norm, norm_type=2.0, error_if_nonfinite=False, foreach=None):
    """`torch.nn.utils.clip_grad_norm_(parameters, max_norm)` (ED/Main_MMR_SegModel.py:722:
    `clip_grad_norm_(self.model.parameters(), 12)`), same signature and return value (the total norm, a
    0-dim tensor on the gradients' device).  Two launches when the gradients are the views of a plan
    model's flat gradient buffer (sum of squares, then an in-place scale that reads the norm on the device:
    no host synchronisation), one pair per tensor otherwise.  `parameters` may also be a plan model or its
    flat gradient tensor."""
    if float(norm_type) != 2.0:
        raise NotImplementedError("only the 2-norm (the reference's call) is built")
    if hasattr(parameters, "flat_parameters"):
        grads = [parameters.flat_parameters()[1]]
    elif isinstance(parameters, torch.Tensor):   # a single parameter (torch's API), or a gradient buffer itself
        grads = [parameters.grad if (parameters.requires_grad or parameters.grad is not None) else parameters]
    else:
        grads = [p.grad for p in parameters if p.grad is not None]
    grads = [g for g in grads if g is not None]
    if not grads:
        return torch.tensor(0.0)
    dev = grads[0].device
    if not grads[0].is_cuda:
        raise _lib.MmrError("clip_grad_norm_ runs on a B200 only (gradients on %s); there is no CPU fallback" % dev)
    # gradients that tile one flat buffer (views of the model's gflat, in order, alignment padding only)
    jobs = []
    order = sorted(grads, key=lambda g: g.data_ptr())
    start, end = order[0].data_ptr(), order[0].data_ptr() + order[0].numel() * 4
    if any(g.dtype != torch.float32 or not g.is_contiguous() for g in order):
        raise _lib.MmrError("clip_grad_norm_ needs contiguous fp32 gradients")
    for g in order[1:]:
        if 0 <= g.data_ptr() - end < 16 and g.untyped_storage().data_ptr() == order[0].untyped_storage().data_ptr():
            end = g.data_ptr() + g.numel() * 4
        else:
            jobs.append((start, (end - start) // 4))
            start, end = g.data_ptr(), g.data_ptr() + g.numel() * 4
    jobs.append((start, (end - start) // 4))
    lib = _lib.lib()
    acc = torch.zeros((1,), device=dev, dtype=torch.float64)
    norm = torch.empty((1,), device=dev, dtype=torch.float32)
    for ptr, n in jobs:
        _lib.check(lib.mmr_sumsq(ptr, n, acc.data_ptr(), _stream()))
    for ptr, n in jobs:
        if ptr % 16:
            raise _lib.MmrError("clip_grad_norm_: gradient buffer is not 16-byte aligned")
        _lib.check(lib.mmr_clip_scale(ptr, n, acc.data_ptr(), float(max_norm), norm.data_ptr(), _stream()))
    if error_if_nonfinite and not torch.isfinite(norm).all():
        raise RuntimeError("The total norm of order 2.0 for gradients from `parameters` is non-finite, so it "
                           "cannot be clipped.")
    return norm[0]
