"""Native checkpoint format (mmrseg_b200/checkpoint.py): lossless round trip to the reference's
state_dict checkpoints, one flat blob for the parameters, optimiser state in torch's layout, and the bf16 OHWI
inference flavour.  Pure host logic: runs on CPU."""
import os
import warnings

import numpy as np
import pytest
import torch

from mmrseg_b200 import checkpoint


def _models():
    from mmrseg_b200.models import ResNetUNet, UNet, UnetPlusPlus
    torch.manual_seed(3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return [UnetPlusPlus("resnet18", classes=10), UnetPlusPlus("resnet34", classes=2, deep_supervision=True),
                ResNetUNet(4, 18), UNet(3, 3, bilinear=True), UNet(3, 2, bilinear=False)]


def _randomize(model, seed):
    g = torch.Generator().manual_seed(seed)
    for b in model.buffers():
        if b.dtype.is_floating_point:
            b.copy_(torch.rand(b.shape, generator=g) + 0.5)
        else:
            b.fill_(7)


@pytest.mark.parametrize("idx", range(5))
def test_round_trip_is_lossless(idx, tmp_path):
    model = _models()[idx]
    _randomize(model, idx)
    want = {k: v.clone() for k, v in model.state_dict().items()}
    path = str(tmp_path / "m.mmrseg")
    checkpoint.save(path, model, extra={"epoch": 12, "best_f1": 0.449})
    # (a) to the reference's state_dict: same keys, same order, same bits
    sd = checkpoint.to_state_dict(path)
    assert list(sd.keys()) == list(want.keys())
    for k in want:
        assert sd[k].dtype == want[k].dtype and torch.equal(sd[k], want[k]), k
    # (b) into a fresh model built from the header alone
    fresh, extra = checkpoint.load(path)
    assert type(fresh) is type(model) and extra == {"epoch": 12, "best_f1": 0.449}
    for k, v in fresh.state_dict().items():
        assert torch.equal(v, want[k]), k
    # (c) the parameters are ONE blob: file size = flat buffer + buffers + header
    n_params = sum(p.numel() for p in model.parameters())
    assert os.path.getsize(path) < 4 * n_params * 1.02 + 64 * 1024 * 4
    # (d) from a reference state_dict (what a smp / in-tree checkpoint holds) and back
    path2 = str(tmp_path / "m2.mmrseg")
    checkpoint.from_state_dict(path2, want, checkpoint._arch_of(model))
    sd2 = checkpoint.to_state_dict(path2)
    assert all(torch.equal(sd2[k], want[k]) for k in want)


def test_wrong_file_and_wrong_architecture(tmp_path):
    from mmrseg_b200.models import UnetPlusPlus
    bad = tmp_path / "x.bin"
    bad.write_bytes(b"not a checkpoint at all")
    with pytest.raises(ValueError):
        checkpoint.to_state_dict(str(bad))
    path = str(tmp_path / "m.mmrseg")
    checkpoint.save(path, UnetPlusPlus("resnet18", classes=2))
    with pytest.raises(ValueError):
        checkpoint.load(path, UnetPlusPlus("resnet18", classes=3))


def test_bf16_inference_checkpoint(tmp_path):
    from mmrseg_b200.models import UnetPlusPlus
    torch.manual_seed(5)
    model = UnetPlusPlus("resnet18", classes=10)
    p32, p16 = str(tmp_path / "a.mmrseg"), str(tmp_path / "b.mmrseg")
    checkpoint.save(p32, model)
    checkpoint.save(p16, model, weights="bf16")
    assert os.path.getsize(p16) < 0.52 * os.path.getsize(p32)
    sd = checkpoint.to_state_dict(p16)
    for k, v in model.state_dict().items():
        if v.dim() == 4:      # conv weights: the bf16 rounding the kernels apply anyway, OIHW again after loading
            assert torch.equal(sd[k], v.to(torch.bfloat16).float()), k
        else:
            assert torch.equal(sd[k], v), k
    loaded, _ = checkpoint.load(p16)
    assert all(p.dtype == torch.float32 for p in loaded.parameters())


def test_optimizer_state_round_trip(tmp_path):
    """FusedAdam state (torch layout: step / exp_avg / exp_avg_sq per parameter) through the native file into a
    fresh FusedAdam and into torch.optim.Adam."""
    from mmrseg_b200.models import UnetPlusPlus
    from mmrseg_b200.optim import FusedAdam
    torch.manual_seed(7)
    model = UnetPlusPlus("resnet18", classes=2)
    model._ensure_flat(torch.device("cpu"))
    opt = FusedAdam(model.parameters(), lr=3e-4, weight_decay=1e-5)
    for i, p in enumerate(model.parameters()):      # what a few steps leave behind (no kernel launch on CPU)
        p.grad = model._gviews[model._flat_names[i]]
    opt._plan(0, list(model.parameters()))
    g = torch.Generator().manual_seed(1)
    for p in model.parameters():
        opt.state[p]["step"] = torch.tensor(5.0)
        opt.state[p]["exp_avg"].copy_(torch.randn(p.shape, generator=g))
        opt.state[p]["exp_avg_sq"].copy_(torch.rand(p.shape, generator=g))
    path = str(tmp_path / "t.mmrseg")
    checkpoint.save(path, model, optimizer=opt, extra={"epoch": 3})
    for make in (lambda ps: FusedAdam(ps, lr=1.0), lambda ps: torch.optim.Adam(ps, lr=1.0)):
        fresh = UnetPlusPlus("resnet18", classes=2)
        opt2 = make(fresh.parameters())
        checkpoint.load(path, fresh, optimizer=opt2)
        assert opt2.param_groups[0]["lr"] == 3e-4 and opt2.param_groups[0]["weight_decay"] == 1e-5
        for p_old, p_new in zip(model.parameters(), fresh.parameters()):
            assert torch.equal(p_old, p_new)
            st = opt2.state[p_new]
            assert float(st["step"]) == 5.0
            assert torch.equal(st["exp_avg"], opt.state[p_old]["exp_avg"])
            assert torch.equal(st["exp_avg_sq"], opt.state[p_old]["exp_avg_sq"])
