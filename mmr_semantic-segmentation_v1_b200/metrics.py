"""Metric seam of the reference on the integer confusion-matrix kernel (csrc/loss_metric.cu).

    Evaluate(key, use_gpu).addBatch / getIoU / getPRF1 / reset      SU/utils.py:31-181
    dice(im1, im2, empty_score)                                      SU/utils.py:523-576
    get_stats(..., mode='multiclass') / iou_score                    smp.metrics, as called at
                                             ED/Main_MMR_SegModel.py:634-639, 1323-1325

Counting is int64 on the device (argmax with torch's first-maximum rule, shared-memory integer
atomics, bit-exact); the reference counts in float32 and is exact only below 2^24 per class and
call.  Ratios are float64 as in the reference.
"""
import ctypes as C

import torch

from . import _lib, ops  # noqa: F401  (ops registers the mmrseg:: custom ops)
from .losses import onehot_to_labels, _stream


def confusion_matrix(logits, labels, cm=None, return_pred=False):
    """cm[n][g][p] += #{label == g and argmax(logits) == p}; logits fp32 [N,C,H,W]."""
    if not logits.is_cuda:
        raise _lib.MmrError("metric kernels run on a B200 only (input on %s); there is no CPU fallback"
                            % logits.device)
    logits = logits.contiguous()
    if logits.dtype != torch.float32:
        logits = logits.float()
    n, c, h, w = logits.shape
    if cm is None:
        cm = torch.zeros((n, c, c), device=logits.device, dtype=torch.int64)
    pred = torch.ops.mmrseg.confusion_from_logits(logits, labels.contiguous(), cm, bool(return_pred))
    return (cm, pred) if return_pred else cm


def confusion_matrix_from_preds(preds, labels, num_classes, ignore_index=None, cm=None, overflow_bin=False):
    if not preds.is_cuda:
        raise _lib.MmrError("metric kernels run on a B200 only (input on %s); there is no CPU fallback"
                            % preds.device)
    preds = preds.contiguous().long()
    labels = labels.contiguous().long()
    n = preds.shape[0]
    npix = preds.numel() // n
    cb = num_classes + (1 if overflow_bin else 0)
    if cm is None:
        cm = torch.zeros((n, cb, cb), device=preds.device, dtype=torch.int64)
    ign = -(2 ** 62) if ignore_index is None else int(ignore_index)
    torch.ops.mmrseg.confusion_from_preds(preds, labels, cm, num_classes, ign, bool(overflow_bin))
    return cm


class Evaluate:
    """`utils.Evaluate`: accumulates tp / fp / fn per class over batches."""

    def __init__(self, key, use_gpu):
        self.num_classes = len(key)
        self.key = key
        self.use_gpu = use_gpu
        self.reset()

    def reset(self):
        self._cm = None

    def addBatch(self, seg, gt, args):
        if getattr(args, "dataset", None) == "synapse":
            seg = seg[:, 0:21, :, :]
            gt = gt[:, 0:21, :, :]
        labels = onehot_to_labels(gt)
        n, c = seg.shape[0], seg.shape[1]
        cm = confusion_matrix(seg, labels)
        cm = cm.sum(0)
        self._cm = cm if self._cm is None else self._cm + cm

    def addBatchFromModel(self, model, img, labels):
        """`seg = model(img); evaluator.addBatch(seg, oneHotGT, args); seg = torch.argmax(seg, 1)`
        (SU/ModelTraining.py:736-760) as ONE replay of the eval plan: argmax and confusion counts are taken in
        the segmentation head's epilogue (`model.segment`), the logits never reach HBM.  labels: int64 [N,H,W]
        class indices (what the one-hot ground truth encodes).  Returns the argmax'ed prediction (uint8 [N,H,W])."""
        pred, cm = model.segment(img, labels)
        cm = cm.sum(0)
        self._cm = cm if self._cm is None else self._cm + cm
        return pred

    def confusion(self):
        """int64 [C,C] on the device: rows ground truth, columns prediction."""
        return self._cm

    def _tpfpfn(self):
        if self._cm is None:
            return 0, 0, 0
        cm = self._cm.cpu()
        tp = cm.diagonal().double()
        return tp, cm.sum(0).double() - tp, cm.sum(1).double() - tp

    @property
    def tp(self):
        return self._tpfpfn()[0]

    @property
    def fp(self):
        return self._tpfpfn()[1]

    @property
    def fn(self):
        return self._tpfpfn()[2]

    def getIoU(self):
        tp, fp, fn = self._tpfpfn()
        return tp / (tp + fp + fn + 1e-15)

    def getPRF1(self):
        tp, fp, fn = self._tpfpfn()
        epsilon = 1e-15
        precision = tp / (tp + fp + epsilon)
        recall = tp / (tp + fn + epsilon)
        f1 = (2 * precision * recall) / (precision + recall + epsilon)
        return precision, recall, f1


def dice(im1, im2, empty_score=1.0):
    """`utils.dice` on device tensors: 2|A and B| / (|A| + |B|) of the boolean masks."""
    a = torch.as_tensor(im1)
    b = torch.as_tensor(im2)
    if a.shape != b.shape:
        raise ValueError("Shape mismatch: im1 and im2 must have the same shape.")
    a = (a != 0).reshape(1, -1).long()
    b = (b != 0).reshape(1, -1).long()
    cm = confusion_matrix_from_preds(a, b, 2).cpu()[0]
    im_sum = int(cm[1].sum() + cm[:, 1].sum())
    if im_sum == 0:
        return empty_score
    return 2.0 * int(cm[1, 1]) / im_sum


def dice_per_image(preds, labels, num_classes, empty_score=1.0):
    """The reference's per-image Dice of the validation loop (SU/ModelTraining.py:625-634, 765-774):
    `dice(one_hot(seg_im), one_hot(label_im))` for every image of the batch, i.e. 2 |A and B| / (|A| + |B|) over
    the C x H x W one-hot volumes, without the per-image `.cpu()` round trips: one confusion-matrix launch
    for the whole batch, |A and B| = trace(cm[n]), |A| + |B| = row sums + column sums of cm[n].
    preds / labels: int64 [N, H, W] on the device; returns float64 [N] on the device."""
    cm = confusion_matrix_from_preds(preds, labels, num_classes).double()
    inter = torch.diagonal(cm, dim1=1, dim2=2).sum(1)
    total = cm.sum((1, 2)) * 2.0            # every counted pixel is one element of A and one of B
    out = 2.0 * inter / total.clamp_min(1.0)
    return torch.where(total > 0, out, torch.full_like(out, float(empty_score)))


def hausdorff_distance(preds, labels, num_classes, workspace=None):
    """skimage.metrics.hausdorff_distance(seg_slice, label_slice) for every image and class of a batch, on the
    device (the reference's every-25-epochs loop, SU/ModelTraining.py:625-649, 765-789, moves every image to the
    CPU and runs two cKDTree queries per class).  preds: uint8 or int64 [N,H,W] class indices (torch.argmax's
    output or model.segment()'s), labels: int64 [N,H,W].  Returns float64 [N, num_classes] on the device: the
    exact Euclidean Hausdorff distance, 0 where both masks are empty, inf where exactly one is."""
    if not preds.is_cuda:
        raise _lib.MmrError("metric kernels run on a B200 only (input on %s); there is no CPU fallback" % preds.device)
    if preds.shape != labels.shape or preds.dim() != 3:
        raise ValueError("preds and labels must both be [N, H, W], got %s and %s" % (tuple(preds.shape), tuple(labels.shape)))
    if preds.dtype not in (torch.uint8, torch.int64):
        preds = preds.long()
    preds, labels = preds.contiguous(), labels.contiguous().long()
    n, h, w = preds.shape
    lib = _lib.lib()
    per_image = lib.mmr_hausdorff_workspace_bytes(num_classes, h, w)
    if workspace is None or workspace.numel() < per_image:
        workspace = torch.empty((min(n, max(1, (256 << 20) // per_image)) * per_image,), device=preds.device,
                                dtype=torch.uint8)
    hd2 = torch.empty((n, num_classes), device=preds.device, dtype=torch.int64)
    _lib.check(lib.mmr_hausdorff_sq(preds.data_ptr(), int(preds.dtype == torch.uint8), labels.data_ptr(), n,
                                    num_classes, h, w, workspace.data_ptr(), workspace.numel(), hd2.data_ptr(),
                                    _stream()))
    out = hd2.double().sqrt()           # exact integers below 2^53: sqrt is the correctly rounded distance
    return torch.where(hd2 < 0, torch.full_like(out, float("inf")), out)     # all-ones (-1 as int64) marks "one mask empty"


def detailed_metrics(preds, labels, num_classes, inf_value=1000.0, empty_score=1.0):
    """The reference's detailed per-image metrics block in two launches per batch instead of N * (1 + C) host round
    trips (SU/ModelTraining.py:625-649): for each image the boolean Dice of the one-hot volumes (`utils.dice`) and,
    per class, the Hausdorff distance with infinite distances capped at `inf_value` (the reference's 1000).
    Returns (dice float64 [N], hausdorff float64 [N, C]) on the device; `dice.sum()` and `hausdorff.sum()` are what
    the reference adds to total_dice_coeff / total_haus_dist."""
    d = dice_per_image(preds.long() if preds.dtype != torch.int64 else preds, labels, num_classes, empty_score)
    hd = hausdorff_distance(preds, labels, num_classes)
    return d, torch.where(torch.isinf(hd), torch.full_like(hd, float(inf_value)), hd)


def get_stats(output, target, mode="multiclass", ignore_index=None, threshold=None, num_classes=None):
    """smp.metrics.get_stats for mode='multiclass': (tp, fp, fn, tn), each int64 [N, C]."""
    if mode != "multiclass":
        raise NotImplementedError("only mode='multiclass' (the reference's) is built")
    if num_classes is None:
        raise ValueError("``num_classes`` attribute should be not ``None`` for 'multiclass' mode.")
    if output.shape != target.shape:
        raise ValueError("Dimensions should match, but ``output`` shape is not equal to ``target`` "
                         "shape, %s != %s" % (tuple(output.shape), tuple(target.shape)))
    n = output.shape[0]
    npix = output.numel() // n
    cm = confusion_matrix_from_preds(output.reshape(n, -1), target.reshape(n, -1), num_classes,
                                     ignore_index, overflow_bin=True)
    c = num_classes
    tp = cm[:, :c, :c].diagonal(dim1=1, dim2=2)
    fp = cm.sum(1)[:, :c] - tp          # histogram of the predictions - tp
    fn = cm.sum(2)[:, :c] - tp          # histogram of the labels - tp
    # smp: tn = numel - tp - fp - fn - ignored;  numel - ignored = every pixel the kernel counted
    tn = cm.sum((1, 2))[:, None] - tp - fp - fn
    return tp.contiguous(), fp.contiguous(), fn.contiguous(), tn.contiguous()


def iou_score(tp, fp, fn, tn, reduction=None, class_weights=None, zero_division=1.0):
    """smp.metrics.iou_score: tp/(tp+fp+fn), NaN -> zero_division; reductions none / macro / micro."""
    def score(a, b, c):
        s = a / (a + b + c)
        return torch.where(torch.isnan(s), torch.full_like(s, float(zero_division)), s)
    tp, fp, fn = tp.float(), fp.float(), fn.float()
    if reduction in (None, "none"):
        return score(tp, fp, fn)
    if reduction == "macro":
        return score(tp.sum(0), fp.sum(0), fn.sum(0)).mean()
    if reduction == "micro":
        return score(tp.sum(), fp.sum(), fn.sum())
    raise NotImplementedError("reduction %r" % (reduction,))
