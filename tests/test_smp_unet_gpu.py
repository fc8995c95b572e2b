"""`smp.Unet` on the ResNet encoders (the reference's `--model smp_unet18`, SU/ModelTraining.py:255-262)
against the oracle restatement (oracle/unetpp.py Unet; smp itself is not on disk: numerically unpinned).

Tolerances as for U-Net++ (tests/test_model_gpu.py, tests/test_parity_gpu.py): eval logits <= 2e-2; train-mode logits
<= 8e-2, loss <= 2e-3, parameter gradients <= 1.2e-1 and cosine >= 0.99 against the fp32 oracle back-propagating through
the engine's ReLU sign pattern; teacher-forced layer-local parity bf16 <= 5e-4 / fp32 <= 1e-4."""
import pytest
import torch

from tests.helpers import ReLUWithMasks, rel, synthetic_batch

pytestmark = pytest.mark.gpu


def _pair(classes, encoder="resnet18", seed=6210):
    from oracle.unetpp import Unet as OracleNet
    from mmrseg_b200.models import Unet
    torch.manual_seed(seed)
    ref = OracleNet(encoder, None, 3, classes)
    g = torch.Generator().manual_seed(seed + 1)
    for m in ref.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.uniform_(0.5, 1.5, generator=g)
            m.bias.data.normal_(0, 0.2, generator=g)
            m.running_mean.normal_(0, 0.2, generator=g)
            m.running_var.uniform_(0.5, 1.5, generator=g)
    net = Unet(encoder, classes=classes)
    net.load_state_dict(ref.state_dict(), strict=True)
    return ref, net.cuda()


def _install_masks(ref, eng):
    m = lambda name: (eng.acts[name].buf.float().permute(0, 3, 1, 2).cpu() > 0).float()
    ref.encoder.relu = ReLUWithMasks([m("f_stem")])
    for li in range(1, 5):
        for bi, blk in enumerate(getattr(ref.encoder, "layer%d" % li)):
            base = "encoder.layer%d.%d." % (li, bi)
            blk.relu = ReLUWithMasks([m(base + "t1"), m(base + "out")])
    for i, blk in enumerate(ref.decoder.blocks):
        blk.conv1[2] = ReLUWithMasks([m("d%d.mid" % i)])
        blk.conv2[2] = ReLUWithMasks([m("d%d" % i)])


def test_constructor_and_state_dict_keys():
    import mmrseg_b200.models as smp
    net = smp.Unet(encoder_name="resnet18", encoder_weights="imagenet", in_channels=3, classes=4)   # the stock call
    keys = list(net.state_dict().keys())
    assert "decoder.blocks.0.conv1.0.weight" in keys and "decoder.blocks.4.conv2.1.running_var" in keys
    assert net.state_dict()["decoder.blocks.3.conv1.0.weight"].shape == (32, 128, 3, 3)
    assert net.state_dict()["segmentation_head.0.weight"].shape == (4, 16, 3, 3)
    assert abs(sum(p.numel() for p in net.parameters()) - 14_328_644) < 10        # smp's 14.3 M at 4 classes
    assert type(smp.create_model("Unet", "resnet34", None, 3, 2)).__name__ == "Unet"


@pytest.mark.parametrize("encoder,classes,shape", [("resnet18", 3, (2, 64, 96)), ("resnet34", 10, (1, 128, 160))])
def test_eval_forward_matches_oracle(encoder, classes, shape):
    ref, net = _pair(classes, encoder)
    x, _ = synthetic_batch(shape[0], classes, shape[1], shape[2])
    ref.eval()
    net.eval()
    with torch.no_grad():
        want = ref(x)
        got = net(x.cuda()).cpu()
    assert got.shape == want.shape
    assert rel(got, want) <= 2e-2, rel(got, want)


def test_train_step_matches_oracle():
    from oracle.losses import mixed_loss
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    ref, net = _pair(3)
    x, y = synthetic_batch(4, 3, 64, 64)
    ref.train()
    net.train()
    got = net(x.cuda())
    loss = DiceCrossEntropyLoss(0.5)(got, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    _install_masks(ref, list(net._engines.values())[0])
    want = ref(x)
    loss_ref = mixed_loss(want, y, 0.5)
    loss_ref.backward()
    assert rel(got.detach().cpu(), want.detach()) <= 8e-2, rel(got.detach().cpu(), want.detach())
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item())
    ref_params = dict(ref.named_parameters())
    for name, p in net.named_parameters():
        assert p.grad is not None, name
        g, r = p.grad.cpu(), ref_params[name].grad
        cos = torch.nn.functional.cosine_similarity(g.flatten(), r.flatten(), dim=0).item()
        assert rel(g, r) <= 1.2e-1 and cos >= 0.99, (name, rel(g, r), cos)


def test_teacher_forced_at_512(monkeypatch):
    from tests.test_parity_gpu import _teacher_forced
    _, net = _pair(2)
    x, y = synthetic_batch(1, 2, 512, 512)
    worst = _teacher_forced(net, x, y, monkeypatch)
    print("teacher-forced smp.Unet-R18 (1, 512, 512): worst %s" % (worst,))
