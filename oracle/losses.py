"""Oracle restatement of the reference's losses (TEST INFRASTRUCTURE; PyTorch fp32 on CPU)."""
import torch
import torch.nn.functional as F


def kornia_one_hot(labels, num_classes, device=None, dtype=torch.float32, eps=1e-6):
    """kornia.utils.one_hot.one_hot (imported at SU/dice_loss.py:33; not vendored, not
    installed: restated from its published source).  Note the eps added to EVERY element."""
    n, h, w = labels.shape
    oh = torch.zeros((n, num_classes, h, w), device=device or labels.device, dtype=dtype)
    return oh.scatter_(1, labels.unsqueeze(1), 1.0) + eps


def dice_loss(input, target, eps=1.0, ignore_index=None):
    """SU/dice_loss.py:118-159, line for line."""
    if not isinstance(input, torch.Tensor):
        raise TypeError(f"Input type is not a torch.Tensor. Got {type(input)}")
    if not len(input.shape) == 4:
        raise ValueError(f"Invalid input shape, we expect BxCxHxW. Got: {input.shape}")
    if not input.shape[-2:] == target.shape[-2:]:
        raise ValueError(f"input and target shapes must be the same. Got: {input.shape} and {target.shape}")
    if not input.device == target.device:
        raise ValueError(f"input and target must be in the same device. Got: {input.device} and {target.device}")
    input_soft = F.softmax(input, dim=1)                                   # :118
    target_one_hot = kornia_one_hot(target, input.shape[1], input.device, input.dtype)  # :124-129
    if ignore_index is not None:                                           # :134-136
        input_soft = input_soft[:, :ignore_index]
        target_one_hot = target_one_hot[:, :ignore_index]
    dims = (2, 3)
    intersection = torch.sum(input_soft * target_one_hot, dims)            # :145
    cardinality = torch.sum(input_soft + target_one_hot, dims)             # :149
    dice_score = (2.0 * intersection + eps) / (cardinality + eps)          # :153
    return torch.mean(-dice_score + 1.0)                                   # :159


def mixed_loss(seg, label, dice_loss_factor=0.5, ce_ignore_index=-100):
    """SU/ModelTraining.py:600-603: w*dice + (1-w)*CE; w = -1 means CE only (:342-360)."""
    ce = F.cross_entropy(seg, label, ignore_index=ce_ignore_index)
    if dice_loss_factor == -1:
        return ce
    return dice_loss_factor * dice_loss(seg, label) + (1 - dice_loss_factor) * ce


def monai_dice_ce(input, target_onehot, smooth_nr=1e-5, smooth_dr=1e-5, lambda_dice=1.0, lambda_ce=1.0):
    """monai.losses.DiceCELoss(softmax=True) with the defaults the reference leaves in place
    (ED/Main_MMR_SegModel.py:578): include_background, no squared_pred, reduction mean, batch=False.
    Un-vendored dependency, restated from the published algorithm."""
    p = F.softmax(input, 1)
    inter = torch.sum(p * target_onehot, (2, 3))
    den = torch.sum(target_onehot, (2, 3)) + torch.sum(p, (2, 3))
    dice = torch.mean(1.0 - (2.0 * inter + smooth_nr) / (den + smooth_dr))
    ce = F.cross_entropy(input, target_onehot)  # probabilities target == one-hot float
    return lambda_dice * dice + lambda_ce * ce
