"""Optimiser seam: torch.optim.Adam(lr, weight_decay) (SU/ModelTraining.py:366), AdamW
(:369, ED/Main_MMR_SegModel.py:878-880), SGD(momentum=0.9) (SU/ModelTraining.py:372,381),
clip_grad_norm_ (ED/...:718-727) on the flat fp32
parameter / gradient buffers of a plan model: one HBM-bound launch (28 B per parameter)."""
import ctypes as C
import math

import torch

from . import _lib
from .losses import _stream


class FusedAdam(torch.optim.Optimizer):
    """Adam (L2 folded into the gradient, like torch.optim.Adam) or AdamW (decoupled=True)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                      decoupled=decoupled))
        self.grad_scale = 1.0

    def _flat_of(self, tensors):
        """Flat fp32 views spanning the storages of `tensors` (they are views of one buffer)."""
        st = tensors[0].untyped_storage()
        if any(t.untyped_storage().data_ptr() != st.data_ptr() for t in tensors):
            return None
        flat = torch.empty(0, dtype=torch.float32, device=tensors[0].device).set_(st)
        return flat

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = _lib.lib()
        for group in self.param_groups:
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            b1, b2 = group["betas"]
            st = self.state.setdefault("group%d" % id(group), {})
            st["step"] = st.get("step", 0) + 1
            bc1, bc2 = 1.0 - b1 ** st["step"], 1.0 - b2 ** st["step"]
            flat = self._flat_of([p.data for p in params])
            gflat = self._flat_of([p.grad for p in params])
            if flat is not None and gflat is not None and flat.numel() == gflat.numel() and \
                    len(params) == len(group["params"]):
                jobs = [("flat", flat, gflat)]
            else:
                jobs = [(id(p), p.data, p.grad) for p in params]
            for key, pt, gt in jobs:
                if key not in st:
                    st[key] = (torch.zeros_like(pt), torch.zeros_like(pt))
                m, v = st[key]
                _lib.check(lib.mmr_adam_step(pt.data_ptr(), gt.data_ptr(), m.data_ptr(), v.data_ptr(),
                                             pt.numel(), group["lr"], b1, b2, group["eps"],
                                             group["weight_decay"], bc1, bc2,
                                             1 if group["decoupled"] else 0, self.grad_scale, _stream()))
        return loss


def _flat_view(tensors):
    """One fp32 tensor spanning the common storage of `tensors` (views of one buffer), else None."""
    st = tensors[0].untyped_storage()
    if any(t.untyped_storage().data_ptr() != st.data_ptr() for t in tensors):
        return None
    return torch.empty(0, dtype=torch.float32, device=tensors[0].device).set_(st)


class FusedSGD(torch.optim.Optimizer):
    """torch.optim.SGD(lr, momentum, weight_decay) (dampening 0, no Nesterov), the reference's
    `optim.SGD(model.parameters(), lr=args.lr, momentum=0.9)`; per-group learning rates (the differential
    encoder / decoder rates of SU/ModelTraining.py:375-383) run one launch per parameter tensor."""

    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self.grad_scale = 1.0

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = _lib.lib()
        for group in self.param_groups:
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            st = self.state.setdefault("group%d" % id(group), {})
            flat = _flat_view([p.data for p in params]) if len(self.param_groups) == 1 else None
            gflat = _flat_view([p.grad for p in params]) if flat is not None else None
            if flat is not None and gflat is not None and flat.numel() == gflat.numel() and \
                    len(params) == len(group["params"]):
                jobs = [("flat", flat, gflat)]
            else:
                jobs = [(id(p), p.data, p.grad) for p in params]
            for key, pt, gt in jobs:
                first = key not in st
                if first:
                    st[key] = torch.zeros_like(pt) if group["momentum"] != 0 else None
                buf = st[key]
                _lib.check(lib.mmr_sgd_step(pt.data_ptr(), gt.data_ptr(), buf.data_ptr() if buf is not None else None,
                                            pt.numel(), group["lr"], group["momentum"], group["weight_decay"],
                                            int(first), self.grad_scale, _stream()))
        return loss


def clip_grad_norm_(model_or_gflat, max_norm):
    """torch.nn.utils.clip_grad_norm_ on the flat gradient buffer; returns the total norm as a
    device tensor (no host sync) and scales the gradients in place when it exceeds max_norm."""
    g = model_or_gflat.flat_parameters()[1] if hasattr(model_or_gflat, "flat_parameters") else model_or_gflat
    acc = torch.zeros((1,), device=g.device, dtype=torch.float64)
    _lib.check(_lib.lib().mmr_sumsq(g.data_ptr(), g.numel(), acc.data_ptr(), _stream()))
    norm = acc.sqrt().float()
    coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
    g.mul_(coef)
    return norm[0]
