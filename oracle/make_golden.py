"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE from /root/reference
(read-only, only present in the build container).  Commit the outputs; the GPU box never
reads /root/reference.

    python oracle/make_golden.py

What is pinned:
  metrics_reference.npz   SU/utils.py Evaluate.addBatch/getIoU/getPRF1 and utils.dice on seeded
                          logits / labels (incl. ties, an absent class, an all-background image)
  normalize_reference.npz SU/utils.py normalize()
  resnet_unet_reference.npz  SU/UArchModel/resnet_unet.py ResNetUNet(n_class=3, resnet_model=18)
                          eval-mode logits on a seeded 1x3x64x64 input, weights from
                          torch.manual_seed(6210) with torchvision `pretrained` patched off
                          (no network); stores the input, the logits and a per-tensor checksum
                          of the state_dict so the weights can be regenerated and verified.
  unet_convt_reference.npz  the same file's UNet(3, 3, bilinear=False): ConvTranspose2d 2x2 stride-2 upsampling
  unet_reference.npz      SU/UArchModel/unet.py UNet(n_channels=3, n_classes=3, bilinear=True) (the call
                          of SU/ModelTraining.py:242): eval-mode logits with randomised BatchNorm
                          statistics and train-mode logits on a seeded 2x3x32x48 input, same checksums.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/MMR_Seg_Unet/MMR_Core_ModelData"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_reference_utils():
    sys.path.insert(0, REF)
    import utils as ref_utils  # noqa: the reference's SU/utils.py
    return ref_utils


def metric_cases():
    g = torch.Generator().manual_seed(6210)
    cases = []
    # (name, N, C, H, W)
    for name, n, c, h, w in [("c2", 2, 2, 16, 16), ("c10", 3, 10, 12, 20), ("c5", 1, 5, 8, 8)]:
        logits = torch.randn((n, c, h, w), generator=g)
        labels = torch.randint(0, c, (n, h, w), generator=g)
        cases.append((name, logits, labels))
    # ties (bf16-like coarse logits) and an absent class / all-background image
    logits = (torch.randint(-2, 3, (2, 4, 16, 16), generator=g)).float() * 0.5
    labels = torch.randint(0, 3, (2, 16, 16), generator=g)  # class 3 never appears
    labels[1] = 0
    cases.append(("ties", logits, labels))
    return cases


def unet_golden(bilinear=True):
    """Run the reference's own UNet (package import: unet.py does `from .unet_parts import *`); bilinear=False is
    the ConvTranspose2d variant of unet_parts.py:269 (-> unet_convt_reference.npz)."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from UArchModel import unet as ref_unet  # the reference's SU/UArchModel/unet.py
    torch.manual_seed(6210)
    model = ref_unet.UNet(n_channels=3, n_classes=3, bilinear=bilinear)
    g = torch.Generator().manual_seed(6211)
    for m in model.modules():   # non-trivial BatchNorm state, as tests/helpers.model_pair does
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.uniform_(0.5, 1.5, generator=g)
            m.bias.data.normal_(0, 0.2, generator=g)
            m.running_mean.normal_(0, 0.2, generator=g)
            m.running_var.uniform_(0.5, 1.5, generator=g)
    x = torch.randn((2, 3, 32, 48), generator=torch.Generator().manual_seed(1))
    model.eval()
    with torch.no_grad():
        y_eval = model(x)
    sums = {k: float(v.double().sum()) for k, v in model.state_dict().items() if v.dtype.is_floating_point}
    model.train()
    with torch.no_grad():
        y_train = model(x)
    name = "unet_reference.npz" if bilinear else "unet_convt_reference.npz"
    np.savez_compressed(os.path.join(OUT, name), x=x.numpy(), logits_eval=y_eval.numpy(),
                        logits_train=y_train.numpy(), keys=np.array(list(sums.keys())),
                        sums=np.array(list(sums.values())))
    print("wrote", name)


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--unet-only" in sys.argv:
        unet_golden()
        unet_golden(bilinear=False)
        return
    U = load_reference_utils()
    args = types.SimpleNamespace(dataset="sarrarp50")
    out = {}
    for name, logits, labels in metric_cases():
        c = logits.shape[1]
        key = {i: np.array([i, i, i]) for i in range(c)}
        ev = U.Evaluate(key, use_gpu=True)  # use_gpu=True dodges the inverted .cuda() branch
        onehot = torch.nn.functional.one_hot(labels, c).permute(0, 3, 1, 2)
        ev.addBatch(logits, onehot, args)
        ev.addBatch(logits.flip(0), onehot.flip(0), args)  # accumulation over two batches
        p, r, f1 = ev.getPRF1()
        out[name + "_logits"] = logits.numpy()
        out[name + "_labels"] = labels.numpy()
        out[name + "_tp"] = ev.tp.numpy()
        out[name + "_fp"] = ev.fp.numpy()
        out[name + "_fn"] = ev.fn.numpy()
        out[name + "_iou"] = ev.getIoU().numpy()
        out[name + "_p"], out[name + "_r"], out[name + "_f1"] = p.numpy(), r.numpy(), f1.numpy()
        pred = torch.argmax(logits, 1)
        d = []
        for i in range(logits.shape[0]):
            a = torch.nn.functional.one_hot(pred[i], c).permute(2, 0, 1).numpy()
            b = onehot[i].numpy()
            d.append(U.dice(a, b))
        d.append(U.dice(np.zeros((4, 4)), np.zeros((4, 4))))       # both empty -> empty_score
        d.append(U.dice(np.zeros((4, 4)), np.zeros((4, 4)), 0.5))
        out[name + "_dice"] = np.array(d, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "metrics_reference.npz"), **out)

    g = torch.Generator().manual_seed(6210)
    batch = torch.rand((3, 3, 8, 8), generator=g)
    mean, std = torch.tensor([0.485, 0.456, 0.406]), torch.tensor([0.229, 0.224, 0.225])
    normed = U.normalize(batch.clone(), mean.clone(), std.clone())
    np.savez_compressed(os.path.join(OUT, "normalize_reference.npz"), batch=batch.numpy(),
                        normed=normed.numpy())

    # ResNetUNet from the reference, random-init (pretrained patched off: no network)
    import torchvision
    orig18, orig34 = torchvision.models.resnet18, torchvision.models.resnet34
    torchvision.models.resnet18 = lambda pretrained=False, **k: orig18(weights=None)
    torchvision.models.resnet34 = lambda pretrained=False, **k: orig34(weights=None)
    sys.path.insert(0, os.path.join(REF, "UArchModel"))
    import resnet_unet as ref_ru  # the reference's SU/UArchModel/resnet_unet.py
    torch.manual_seed(6210)
    model = ref_ru.ResNetUNet(3, 18).eval()
    x = torch.randn((1, 3, 64, 64), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        y = model(x)
    sums = {k: float(v.double().sum()) for k, v in model.state_dict().items() if v.dtype.is_floating_point}
    np.savez_compressed(os.path.join(OUT, "resnet_unet_reference.npz"), x=x.numpy(), logits=y.numpy(),
                        keys=np.array(list(sums.keys())), sums=np.array(list(sums.values())))
    torchvision.models.resnet18, torchvision.models.resnet34 = orig18, orig34
    unet_golden()
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
