"""The torch custom-op layer (mmrseg_b200/ops.py) on a box without a GPU: every op of the hot path is
registered under the `mmrseg::` namespace with a schema, a fake (meta) implementation that gives the output
shapes / dtypes without launching anything, and -- for the differentiable ones -- an autograd formula."""
import torch

import mmrseg_b200.ops  # noqa: F401


def test_ops_are_registered_with_schemas():
    names = ["plan_forward", "plan_backward", "dice_ce_fwd", "dice_ce_bwd", "confusion_from_logits",
             "confusion_from_preds", "adam_step", "sgd_step"]
    for n in names:
        op = getattr(torch.ops.mmrseg, n)
        assert op.default._schema.name == "mmrseg::" + n
    s = str(torch.ops.mmrseg.adam_step.default._schema)
    assert "!) p" in s and "!) m" in s and "!) v" in s and "Tensor g," in s      # declared mutations: p, m, v
    assert "!) cm" in str(torch.ops.mmrseg.confusion_from_logits.default._schema)


def test_fake_implementations_give_shapes_without_a_gpu():
    logits = torch.empty((4, 10, 64, 96), device="meta")
    labels = torch.empty((4, 64, 96), dtype=torch.int64, device="meta")
    out, ws = torch.ops.mmrseg.dice_ce_fwd(logits, labels, 1.0, 1.0, 1e-6, 0.5, 0.5, 10, -100)
    assert out.shape == (3,) and out.dtype == torch.float32 and ws.dtype == torch.float64 and ws.numel() > 0
    d = torch.ops.mmrseg.dice_ce_bwd(logits, labels, ws, out[0:1], 1.0, 1.0, 1e-6, 0.5, 0.5, 10, -100)
    assert d.shape == logits.shape and d.dtype == torch.float32
    cm = torch.empty((4, 10, 10), dtype=torch.int64, device="meta")
    pred = torch.ops.mmrseg.confusion_from_logits(logits, labels, cm, True)
    assert pred.shape == (4, 64, 96) and pred.dtype == torch.int64
    # a model plan: logits of every head, stacked
    from mmrseg_b200.models import UnetPlusPlus
    net = UnetPlusPlus("resnet18", classes=10, deep_supervision=True)
    x = torch.empty((2, 3, 64, 96), device="meta")
    assert torch.ops.mmrseg.plan_forward(x, net._handle, True, []).shape == (4, 2, 10, 64, 96)
    assert torch.ops.mmrseg.plan_forward(x, net._handle, False, []).shape == (1, 2, 10, 64, 96)
    frames = torch.empty((2, 64, 96, 3), dtype=torch.uint8, device="meta")
    assert torch.ops.mmrseg.plan_forward(frames, net._handle, False, []).shape == (1, 2, 10, 64, 96)


def test_real_tensors_on_cpu_fail_loudly():
    import pytest
    from mmrseg_b200._lib import MmrError
    with pytest.raises(MmrError):
        torch.ops.mmrseg.dice_ce_fwd(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long), 1.0, 1.0, 1e-6,
                                     0.5, 0.5, 2, -100)
