"""Loss seam of the reference, on one pair of sm_100a kernels (csrc/loss_metric.cu).

    dice_loss / DiceLoss           SU/dice_loss.py:37-161, 165-259   (softmax soft-Dice, eps 1.0,
                                   kornia one_hot: +1e-6 on every element)
    w*dice + (1-w)*CrossEntropy    SU/ModelTraining.py:342-360, 600-603 (train), 747-750 (val)
    DiceCELoss(softmax=True)       monai, ED/Main_MMR_SegModel.py:578, 709, 822

All of them are one call of the custom op mmrseg::dice_ce_fwd (mmr_dice_ce_fwd: per-(image, class) sums,
warp-shuffle reduction) and, under autograd, one call of mmrseg::dice_ce_bwd (mmrseg_b200/ops.py).  Inputs: fp32 NCHW logits (what the models
return) and int64 labels; no one-hot tensor and no softmax tensor is materialised.
"""
import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _lib

def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _loss(logits, labels, prm):
    """One mmrseg::dice_ce_fwd (autograd: mmrseg::dice_ce_bwd); prm = (dice_eps_nr, dice_eps_dr, onehot_eps, w_dice,
    w_ce, dice_channels, ce_ignore_index).  Returns (total, dice, ce) as a 3-element fp32 tensor."""
    from . import ops
    if not logits.is_cuda:
        raise _lib.MmrError("loss kernels run on a B200 only (input on %s); there is no CPU fallback"
                            % logits.device)
    if logits.dtype != torch.float32:
        logits = logits.float()
    out, _ = torch.ops.mmrseg.dice_ce_fwd(logits, labels, *prm)
    return out


def _validate(input, target):
    # same checks and messages as SU/dice_loss.py:100-113
    if not isinstance(input, torch.Tensor):
        raise TypeError(f"Input type is not a torch.Tensor. Got {type(input)}")
    if not len(input.shape) == 4:
        raise ValueError(f"Invalid input shape, we expect BxCxHxW. Got: {input.shape}")
    if not input.shape[-2:] == target.shape[-2:]:
        raise ValueError(f"input and target shapes must be the same. Got: {input.shape} and {target.shape}")
    if not input.device == target.device:
        raise ValueError(f"input and target must be in the same device. Got: {input.device} and {target.device}")


def dice_loss(input: torch.Tensor, target: torch.Tensor, eps: float = 1.0,
              ignore_index: Optional[int] = None) -> torch.Tensor:
    """`dice_loss(input, target, eps, ignore_index)` of SU/dice_loss.py:37-161."""
    _validate(input, target)
    c = input.shape[1]
    dc = c if ignore_index is None else int(ignore_index)
    if dc < 0:
        dc += c
    if not 1 <= dc <= c:
        raise ValueError("ignore_index=%r leaves no channel in the Dice term" % (ignore_index,))
    return _loss(input, target, (float(eps), float(eps), 1e-6, 1.0, 0.0, dc, -100))[0]


class DiceLoss(nn.Module):
    """`DiceLoss(eps=1.0, ignore_index=None)` of SU/dice_loss.py:165-259."""

    def __init__(self, eps: float = 1.0, ignore_index=None) -> None:
        super().__init__()
        self.eps: float = eps
        self.ignore_index = ignore_index

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return dice_loss(input, target, self.eps, self.ignore_index)


class DiceCrossEntropyLoss(nn.Module):
    """The reference's training objective as one kernel: `w*dice_loss(seg,label) +
    (1-w)*CrossEntropyLoss(seg,label)` (SU/ModelTraining.py:600-603); `dice_weight=-1` means
    cross-entropy only (the `--dice_loss_factor -1` convention, :342-360)."""

    def __init__(self, dice_weight=0.5, eps=1.0, dice_ignore_index=None, ce_ignore_index=-100):
        super().__init__()
        self.dice_weight, self.eps = dice_weight, eps
        self.dice_ignore_index, self.ce_ignore_index = dice_ignore_index, ce_ignore_index

    def params(self, c):
        w = self.dice_weight
        wd, wc = (0.0, 1.0) if w == -1 else (float(w), 1.0 - float(w))
        dc = c if self.dice_ignore_index is None else int(self.dice_ignore_index)
        return (float(self.eps), float(self.eps), 1e-6, wd, wc, dc, int(self.ce_ignore_index))

    def forward(self, input, target):
        _validate(input, target)
        return _loss(input, target, self.params(input.shape[1]))[0]


def onehot_to_labels(onehot):
    """[N,C,H,W] one-hot (float or int64) -> [N,H,W] int64, first maximal channel."""
    n, c, h, w = onehot.shape
    if onehot.dtype == torch.int64:
        src, is_float = onehot.contiguous(), 0
    else:
        src, is_float = onehot.float().contiguous(), 1
    labels = torch.empty((n, h, w), device=onehot.device, dtype=torch.int64)
    _lib.check(_lib.lib().mmr_onehot_to_labels(src.data_ptr(), is_float, n, c, h, w, labels.data_ptr(),
                                               _stream()))
    return labels


class DiceCELoss(nn.Module):
    """monai `DiceCELoss(softmax=True)` as the reference uses it (ED/Main_MMR_SegModel.py:578,709):
    input fp32 logits [N,C,H,W], target one-hot float [N,C,H,W];
    mean_{n,c}[1 - (2I + 1e-5)/(sum p + sum y + 1e-5)] + CrossEntropy."""

    def __init__(self, softmax=True, smooth_nr=1e-5, smooth_dr=1e-5, lambda_dice=1.0, lambda_ce=1.0):
        super().__init__()
        if not softmax:
            raise NotImplementedError("only softmax=True (the reference's setting) is built")
        self.smooth_nr, self.smooth_dr = smooth_nr, smooth_dr
        self.lambda_dice, self.lambda_ce = lambda_dice, lambda_ce

    def forward(self, input, target):
        if target.dim() == 4 and target.shape[1] == input.shape[1]:
            labels = onehot_to_labels(target)
        elif target.dim() == 3:
            labels = target
        else:
            raise ValueError("target must be one-hot [N,C,H,W] or labels [N,H,W], got %s" % (tuple(target.shape),))
        prm = (float(self.smooth_nr), float(self.smooth_dr), 0.0, float(self.lambda_dice), float(self.lambda_ce),
               input.shape[1], -100)
        return _loss(input, labels, prm)[0]
