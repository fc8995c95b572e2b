"""Oracle restatement of the reference's metrics (TEST INFRASTRUCTURE): numpy int64 counting,
float64 ratios.  `Evaluate` and `dice` are pinned by tests/golden/metrics_reference.npz, which
oracle/make_golden.py produced by running the reference's own SU/utils.py."""
import numpy as np


def argmax_first(logits):
    """torch.argmax over dim 1: index of the first maximal value (SU/utils.py:109)."""
    return np.argmax(np.asarray(logits), axis=1)


def confusion_matrix(pred, label, num_classes, ignore_index=None):
    """cm[n, g, p] = #{label == g and pred == p} per image, int64."""
    pred = np.asarray(pred).reshape(pred.shape[0], -1).astype(np.int64)
    label = np.asarray(label).reshape(label.shape[0], -1).astype(np.int64)
    cm = np.zeros((pred.shape[0], num_classes, num_classes), dtype=np.int64)
    for n in range(pred.shape[0]):
        ok = (label[n] >= 0) & (label[n] < num_classes) & (pred[n] >= 0) & (pred[n] < num_classes)
        if ignore_index is not None:
            ok &= label[n] != ignore_index
        idx = label[n][ok] * num_classes + pred[n][ok]
        cm[n] = np.bincount(idx, minlength=num_classes * num_classes).reshape(num_classes, num_classes)
    return cm


class Evaluate:
    """SU/utils.py:31-181 as a confusion matrix: tp = diag, fp = column sum - diag,
    fn = row sum - diag; int64 counts (the reference sums in float32, exact below 2^24)."""

    def __init__(self, key, use_gpu=True):
        self.num_classes = len(key)
        self.reset()

    def reset(self):
        self.tp = np.zeros(self.num_classes)
        self.fp = np.zeros(self.num_classes)
        self.fn = np.zeros(self.num_classes)

    def addBatch(self, seg, gt_onehot, args=None):
        seg, gt = np.asarray(seg), np.asarray(gt_onehot)
        if args is not None and getattr(args, "dataset", None) == "synapse":   # :103-105
            seg, gt = seg[:, 0:21], gt[:, 0:21]
        pred = argmax_first(seg)                                               # :109
        label = np.argmax(gt, axis=1)
        cm = confusion_matrix(pred, label, self.num_classes).sum(0)
        tp = np.diag(cm).astype(np.float64)
        self.tp += tp                                                          # :131-133
        self.fp += cm.sum(0) - tp
        self.fn += cm.sum(1) - tp

    def getIoU(self):                                                          # :140-157
        return self.tp / (self.tp + self.fp + self.fn + 1e-15)

    def getPRF1(self):                                                         # :159-181
        eps = 1e-15
        p = self.tp / (self.tp + self.fp + eps)
        r = self.tp / (self.tp + self.fn + eps)
        return p, r, (2 * p * r) / (p + r + eps)


def dice(im1, im2, empty_score=1.0):
    """SU/utils.py:523-576."""
    im1 = np.asarray(im1).astype(bool)
    im2 = np.asarray(im2).astype(bool)
    if im1.shape != im2.shape:
        raise ValueError("Shape mismatch: im1 and im2 must have the same shape.")
    im_sum = im1.sum() + im2.sum()
    if im_sum == 0:
        return empty_score
    return 2.0 * np.logical_and(im1, im2).sum() / im_sum


def get_stats(output, target, num_classes, ignore_index=None):
    """smp.metrics.get_stats(mode='multiclass') (un-vendored):
    per image tp/fp/fn/tn int64 [N, C]."""
    output = np.asarray(output).astype(np.int64)
    target = np.asarray(target).astype(np.int64)
    n = output.shape[0]
    npix = output[0].size
    ignored = np.zeros(n, dtype=np.int64)
    if ignore_index is not None:
        ign = target == ignore_index
        output = np.where(ign, -1, output)
        target = np.where(ign, -1, target)
        ignored = ign.reshape(n, -1).sum(1)
    tp = np.zeros((n, num_classes), dtype=np.int64)
    fp, fn, tn = tp.copy(), tp.copy(), tp.copy()
    for i in range(n):
        o, t = output[i].ravel(), target[i].ravel()
        matched = np.where(o == t, t, -1)
        hist = lambda v: np.bincount(v[(v >= 0) & (v < num_classes)], minlength=num_classes)
        tp[i] = hist(matched)
        fp[i] = hist(o) - tp[i]
        fn[i] = hist(t) - tp[i]
        tn[i] = npix - tp[i] - fp[i] - fn[i] - ignored[i]
    return tp, fp, fn, tn


def iou_score(tp, fp, fn, tn, reduction=None, zero_division=1.0):
    def score(a, b, c):
        with np.errstate(divide="ignore", invalid="ignore"):
            s = a.astype(np.float32) / (a + b + c).astype(np.float32)
        return np.where(np.isnan(s), np.float32(zero_division), s)
    if reduction in (None, "none"):
        return score(tp, fp, fn)
    if reduction == "macro":
        return score(tp.sum(0), fp.sum(0), fn.sum(0)).mean()
    if reduction == "micro":
        return score(tp.sum(), fp.sum(), fn.sum())
    raise NotImplementedError(reduction)


def hausdorff_distance(image0, image1):
    """skimage.metrics.hausdorff_distance(image0, image1) (method='standard'), as called per class slice at
    SU/ModelTraining.py:644, 784.  scikit-image is an un-vendored dependency that is not installed here: restated
    from its published source (nonzero pixel coordinates of both images, two scipy.spatial.cKDTree nearest-neighbour
    queries, max of the maxima; 0 for two empty images, inf when exactly one is empty) -- parity unpinned at that
    boundary, although the quantity is a closed-form definition."""
    from scipy.spatial import cKDTree
    a = np.transpose(np.nonzero(np.asarray(image0)))
    b = np.transpose(np.nonzero(np.asarray(image1)))
    if len(a) == 0:
        return 0.0 if len(b) == 0 else np.inf
    if len(b) == 0:
        return np.inf
    fwd = cKDTree(a).query(b, k=1)[0]
    bwd = cKDTree(b).query(a, k=1)[0]
    return max(max(fwd), max(bwd))
