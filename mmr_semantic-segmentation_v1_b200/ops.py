"""The thin torch custom-op layer over the C-ABI ("Python side: torch.library.custom_op("mmrseg::...",
mutates_args=...) + register_fake + register_autograd").

Every op below is one C-ABI entry point (or one replayed plan of them) of include/mmrseg.h: the op body turns
tensors into raw device pointers and launches on torch's current stream; `register_fake` gives the shape / dtype
rule (so the ops trace under FakeTensor / torch.export without a GPU); `register_autograd` wires the backward op.
The nn.Modules of this package (models / losses / metrics / optim) are argument plumbing around these ops:

    mmrseg::plan_forward / plan_backward        seg = model(img); loss.backward()       SU/ModelTraining.py:589,614
    mmrseg::dice_ce_fwd / dice_ce_bwd           w*dice_loss + (1-w)*CrossEntropy, DiceCELoss  SU/...:600-603, ED/...:709
    mmrseg::confusion_from_logits / _preds      Evaluate.addBatch, get_stats            SU/utils.py:109-133, ED/...:634
    mmrseg::adam_step / sgd_step                optimizer.step()                        SU/...:617, ED/...:718-727
"""
import ctypes as C
import weakref
from typing import List, Tuple

import torch

from . import _lib
from ._lib import MmrLossParams

_LOSS_BLOCKS_CAP = 592


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t, what):
    if not t.is_cuda:
        raise _lib.MmrError("%s runs on a B200 only (input on %s); there is no CPU fallback" % (what, t.device))


# ------------------------------------------------------------------------------------------ loss
def loss_blocks(hw, n):
    return int(max(1, min(-(-hw // 2048), -(-_LOSS_BLOCKS_CAP // n))))


def _prm(dice_eps_nr, dice_eps_dr, onehot_eps, w_dice, w_ce, dice_channels, ce_ignore_index):
    return MmrLossParams(dice_eps_nr, dice_eps_dr, onehot_eps, w_dice, w_ce, dice_channels, ce_ignore_index)


@torch.library.custom_op("mmrseg::dice_ce_fwd", mutates_args=())
def dice_ce_fwd(logits: torch.Tensor, labels: torch.Tensor, dice_eps_nr: float, dice_eps_dr: float,
                onehot_eps: float, w_dice: float, w_ce: float, dice_channels: int,
                ce_ignore_index: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (out[3] = total, dice, ce  fp32;  workspace of per-(image, class) sums, float64, read by the backward)."""
    _need_cuda(logits, "the loss kernel")
    logits, labels = logits.contiguous(), labels.contiguous()
    n, c, h, w = logits.shape
    nblk = loss_blocks(h * w, n)
    lib = _lib.lib()
    ws = torch.empty((lib.mmr_dice_ce_workspace_doubles(n, c, nblk),), device=logits.device, dtype=torch.float64)
    out = torch.empty((3,), device=logits.device, dtype=torch.float32)
    prm = _prm(dice_eps_nr, dice_eps_dr, onehot_eps, w_dice, w_ce, dice_channels, ce_ignore_index)
    _lib.check(lib.mmr_dice_ce_fwd(logits.data_ptr(), labels.data_ptr(), n, c, h, w, C.byref(prm), ws.data_ptr(),
                                   nblk, out.data_ptr(), _stream()))
    return out, ws


@dice_ce_fwd.register_fake
def _(logits, labels, dice_eps_nr, dice_eps_dr, onehot_eps, w_dice, w_ce, dice_channels, ce_ignore_index):
    n, c, h, w = logits.shape
    need = _lib.lib().mmr_dice_ce_workspace_doubles(n, c, loss_blocks(h * w, n))
    return logits.new_empty((3,), dtype=torch.float32), logits.new_empty((need,), dtype=torch.float64)


@torch.library.custom_op("mmrseg::dice_ce_bwd", mutates_args=())
def dice_ce_bwd(logits: torch.Tensor, labels: torch.Tensor, ws: torch.Tensor, grad_out: torch.Tensor,
                dice_eps_nr: float, dice_eps_dr: float, onehot_eps: float, w_dice: float, w_ce: float,
                dice_channels: int, ce_ignore_index: int) -> torch.Tensor:
    """dlogits of (w_dice * dice + w_ce * ce) scaled by the 1-element fp32 tensor grad_out (read on the device)."""
    logits, labels = logits.contiguous(), labels.contiguous()
    n, c, h, w = logits.shape
    dlogits = torch.empty_like(logits)
    g = grad_out.to(torch.float32).contiguous()
    prm = _prm(dice_eps_nr, dice_eps_dr, onehot_eps, w_dice, w_ce, dice_channels, ce_ignore_index)
    _lib.check(_lib.lib().mmr_dice_ce_bwd(logits.data_ptr(), labels.data_ptr(), n, c, h, w, C.byref(prm),
                                          ws.data_ptr(), C.c_float(1.0), g.data_ptr(), dlogits.data_ptr(), _stream()))
    return dlogits


@dice_ce_bwd.register_fake
def _(logits, labels, ws, grad_out, *args):
    return torch.empty_like(logits)


def _dice_ce_setup(ctx, inputs, output):
    logits, labels = inputs[0], inputs[1]
    ctx.prm = tuple(inputs[2:])
    ctx.save_for_backward(logits, labels, output[1])


def _dice_ce_backward(ctx, g_out, g_ws):
    logits, labels, ws = ctx.saved_tensors
    # out = (total, dice, ce): the modules expose `total` = w_dice * dice + w_ce * ce, which is what the backward
    # kernel differentiates (the dice and ce entries are reported values, not differentiable outputs)
    dlogits = torch.ops.mmrseg.dice_ce_bwd(logits, labels, ws, g_out[0:1], *ctx.prm)
    return (dlogits, None) + (None,) * len(ctx.prm)


dice_ce_fwd.register_autograd(_dice_ce_backward, setup_context=_dice_ce_setup)


# ------------------------------------------------------------------------------------------ metric
@torch.library.custom_op("mmrseg::confusion_from_logits", mutates_args=("cm",))
def confusion_from_logits(logits: torch.Tensor, labels: torch.Tensor, cm: torch.Tensor, want_pred: bool) -> torch.Tensor:
    """cm[n][g][p] += #{label == g and argmax(logits) == p} (int64, accumulated in place); returns the argmax
    (int64 [N,H,W]) when want_pred, else an empty tensor."""
    _need_cuda(logits, "the metric kernel")
    logits, labels = logits.contiguous(), labels.contiguous()
    n, c, h, w = logits.shape
    pred = torch.empty((n, h, w) if want_pred else (0,), device=logits.device, dtype=torch.int64)
    _lib.check(_lib.lib().mmr_confusion_from_logits(logits.data_ptr(), labels.data_ptr(), n, c, h, w, cm.data_ptr(),
                                                    pred.data_ptr() if want_pred else None, _stream()))
    return pred


@confusion_from_logits.register_fake
def _(logits, labels, cm, want_pred):
    n, c, h, w = logits.shape
    return logits.new_empty((n, h, w) if want_pred else (0,), dtype=torch.int64)


@torch.library.custom_op("mmrseg::confusion_from_preds", mutates_args=("cm",))
def confusion_from_preds(preds: torch.Tensor, labels: torch.Tensor, cm: torch.Tensor, num_classes: int,
                         ignore_index: int, overflow_bin: bool) -> None:
    _need_cuda(preds, "the metric kernel")
    n = preds.shape[0]
    _lib.check(_lib.lib().mmr_confusion_from_preds(preds.data_ptr(), labels.data_ptr(), n, num_classes,
                                                   preds.numel() // n, ignore_index, int(overflow_bin), cm.data_ptr(),
                                                   _stream()))


@confusion_from_preds.register_fake
def _(preds, labels, cm, num_classes, ignore_index, overflow_bin):
    return None


# ------------------------------------------------------------------------------------------ optimiser
@torch.library.custom_op("mmrseg::adam_step", mutates_args=("p", "m", "v"))
def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, lr: float, beta1: float,
              beta2: float, eps: float, weight_decay: float, step: float, decoupled: bool, grad_scale: float) -> None:
    """torch.optim.Adam / AdamW single-tensor update over contiguous fp32 buffers (28 B per parameter)."""
    _need_cuda(p, "the optimiser kernel")
    _lib.check(_lib.lib().mmr_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1,
                                        beta2, eps, weight_decay, 1.0 - beta1 ** step, 1.0 - beta2 ** step,
                                        1 if decoupled else 0, grad_scale, _stream()))


@adam_step.register_fake
def _(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, decoupled, grad_scale):
    return None


@torch.library.custom_op("mmrseg::sgd_step", mutates_args=("p", "buf"))
def sgd_step(p: torch.Tensor, g: torch.Tensor, buf: torch.Tensor, lr: float, momentum: float, weight_decay: float,
             grad_scale: float) -> None:
    _need_cuda(p, "the optimiser kernel")
    _lib.check(_lib.lib().mmr_sgd_step(p.data_ptr(), g.data_ptr(), buf.data_ptr() if momentum != 0 else None,
                                       p.numel(), lr, momentum, weight_decay, 0, grad_scale, _stream()))


@sgd_step.register_fake
def _(p, g, buf, lr, momentum, weight_decay, grad_scale):
    return None


# ------------------------------------------------------------------------------------------ model plans
_PLANS = {}          # handle -> weakref to the plan model (custom ops take tensors and scalars, not modules)
_next_handle = [1]


def register_model(model):
    h = _next_handle[0]
    _next_handle[0] += 1
    _PLANS[h] = weakref.ref(model)
    return h


def _model(handle):
    ref = _PLANS.get(handle)
    m = ref() if ref is not None else None
    if m is None:
        raise _lib.MmrError("plan handle %d does not name a live model" % handle)
    return m


@torch.library.custom_op("mmrseg::plan_forward", mutates_args=())
def plan_forward(x: torch.Tensor, handle: int, training: bool, params: List[torch.Tensor]) -> torch.Tensor:
    """Replay the forward plan of model `handle` for x (fp32 [N,3,H,W] or uint8 [N,H,W,3] frames).  Returns the
    logits of every head stacked: fp32 [K, N, classes, H, W] (K = 1; 4 with deep supervision in training mode) in
    fresh memory.  `params` are the model's parameters, passed so that autograd sees the dependency; their
    gradients are left in the model's flat gradient buffer by mmrseg::plan_backward."""
    model = _model(handle)
    eng = model._engine_for(x, training=training)
    out = eng.forward(x.contiguous())
    outs = out if isinstance(out, list) else [out]
    return torch.stack(outs, 0)       # the copy out of the plan's buffers: results never alias the next forward


@plan_forward.register_fake
def _(x, handle, training, params):
    model = _model(handle)
    if x.dtype == torch.uint8:
        n, h, w, _ = x.shape
    else:
        n, _, h, w = x.shape
    k = model._n_heads(training)
    return x.new_empty((k, n, model._n_classes(), h, w), dtype=torch.float32)


@torch.library.custom_op("mmrseg::plan_backward", mutates_args=())
def plan_backward(dlogits: torch.Tensor, handle: int, key: List[int], generation: int) -> torch.Tensor:
    """Replay the backward plan: dlogits fp32 [K, N, classes, H, W].  Parameter gradients land in the model's flat
    gradient buffer (published as each parameter's .grad); returns that buffer (no data gradient exists: the
    input is an image)."""
    model = _model(handle)
    eng = model._engines[tuple(key[:3]) + (bool(key[3]),)]
    if eng.generation != generation:
        raise _lib.MmrError(
            "backward() of a forward pass that is no longer the latest one at this input shape: the plan "
            "keeps ONE set of saved activations per (batch, H, W), and another model(x) call in training "
            "mode has overwritten them.  Call backward() before the next training-mode forward (gradient "
            "accumulation: forward, backward, forward, backward), or run the extra forward under "
            "model.eval()")
    accumulate = model._grads_live()
    grads = [g.contiguous() for g in dlogits.unbind(0)]
    cuts = model._grad_cuts(eng.param_ready_hooks) if model._grad_cuts is not None else None
    eng.backward(grads if len(grads) > 1 else grads[0], accumulate=accumulate, on_ready=model._on_grads_ready, cuts=cuts)
    model._publish_grads()
    if model._after_backward is not None:
        model._after_backward()
    return model._gflat


@plan_backward.register_fake
def _(dlogits, handle, key, generation):
    return dlogits.new_empty((_model(handle)._gflat.numel(),), dtype=torch.float32)


def _plan_setup(ctx, inputs, output):
    x, handle, training, params = inputs
    model = _model(handle)
    if x.dtype == torch.uint8:
        n, h, w, _ = x.shape
    else:
        n, _, h, w = x.shape
    ctx.handle, ctx.n_params = handle, len(params)
    ctx.key = [n, h, w, int(bool(training))]
    ctx.generation = model._engines[(n, h, w, bool(training))].generation


def _plan_backward(ctx, g):
    torch.ops.mmrseg.plan_backward(g.contiguous(), ctx.handle, ctx.key, ctx.generation)
    # parameter gradients were published into .grad (views of the flat buffer) by the op: nothing flows back here
    return None, None, None, [None] * ctx.n_params


plan_forward.register_autograd(_plan_backward, setup_context=_plan_setup)
