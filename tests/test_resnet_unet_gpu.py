"""The reference's in-tree ResNet encoder/decoder (SU/UArchModel/resnet_unet.py) through the plan
engine, against the oracle restatement (oracle/resnet_unet.py, itself pinned bit-exactly to the
reference's own file by tests/golden/resnet_unet_reference.npz).

Tolerances (relative Frobenius error): bilinear x2 kernels vs F.interpolate on the same bf16 inputs
4e-3 (output rounding) forward and adjoint; eval logits 2e-2; train-mode logits 8e-2, loss 2e-3, every
parameter gradient 1.5e-1 and cosine >= 0.985 against the fp32 oracle back-propagating through the
engine's ReLU sign pattern (see tests/helpers.py; the decoder has no BatchNorm, so bf16 rounding is
not re-normalised between its layers).
"""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import install_masks_by_call_order, rel, resnet_unet_relu_order, synthetic_batch


def _pair(n_class, resnet_model, seed=6210):
    from oracle.resnet_unet import ResNetUNet as OracleNet
    from mmrseg_b200.models import ResNetUNet
    torch.manual_seed(seed)
    ref = OracleNet(n_class, resnet_model)
    g = torch.Generator().manual_seed(seed + 1)
    for m in ref.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.uniform_(0.5, 1.5, generator=g)
            m.bias.data.normal_(0, 0.2, generator=g)
            m.running_mean.normal_(0, 0.2, generator=g)
            m.running_var.uniform_(0.5, 1.5, generator=g)
    net = ResNetUNet(n_class, resnet_model)
    net.load_state_dict(ref.state_dict(), strict=True)
    return ref, net


def test_state_dict_keys_match_reference_layout():
    ref, net = _pair(10, 18)
    assert list(net.state_dict().keys()) == list(ref.state_dict().keys())
    assert hasattr(net, "base_model") and hasattr(net, "conv_last")
    assert net.conv_last.weight.shape == (10, 64, 1, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 16, 24, 64), (1, 8, 8, 256), (3, 5, 7, 16)])
def test_bilinear_x2_forward_and_adjoint(shape):
    import ctypes as C
    from mmrseg_b200 import _lib
    from mmrseg_b200._lib import MmrContrib
    lib = _lib.lib()
    N, H, W, Cc = shape
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((N, H, W, Cc), generator=g, device="cuda").to(torch.bfloat16)
    out = torch.empty((N, 2 * H, 2 * W, Cc), device="cuda", dtype=torch.bfloat16)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.mmr_upsample_bilinear2x_fwd(C.c_void_p(x.data_ptr()), N, H, W, Cc, C.c_void_p(out.data_ptr()), s))
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True)
    assert rel(out.float().permute(0, 3, 1, 2), ref.detach()) <= 4e-3
    go = torch.randn((N, 2 * H, 2 * W, Cc), generator=g, device="cuda").to(torch.bfloat16)
    go2 = torch.randn((N, 4 * H, 4 * W, Cc), generator=g, device="cuda").to(torch.bfloat16)
    arr = (MmrContrib * 2)()
    arr[0].ptr, arr[0].pool2 = go.data_ptr(), 0
    arr[1].ptr, arr[1].pool2 = go2.data_ptr(), 1
    gin = torch.empty_like(x)
    _lib.check(lib.mmr_upsample_bilinear2x_bwd(arr, 2, N, H, W, Cc, C.c_void_p(gin.data_ptr()), s))
    total = go.float().permute(0, 3, 1, 2) + F.avg_pool2d(go2.float().permute(0, 3, 1, 2), 2) * 4
    ref.backward(total)
    assert rel(gin.float().permute(0, 3, 1, 2), xr.grad) <= 4e-3


@pytest.mark.gpu
@pytest.mark.parametrize("resnet_model,n_class,shape", [(18, 10, (2, 64, 96)), (34, 3, (1, 64, 64))])
def test_eval_forward_matches_oracle(resnet_model, n_class, shape):
    ref, net = _pair(n_class, resnet_model)
    net = net.cuda()
    x, _ = synthetic_batch(shape[0], n_class, shape[1], shape[2])
    ref.eval()
    net.eval()
    with torch.no_grad():
        want = ref(x)
        got = net(x.cuda()).cpu()
    assert got.shape == want.shape
    assert rel(got, want) <= 2e-2, rel(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("resnet_model,n_class,n,hw", [(18, 10, 4, 64), (34, 10, 2, 128)])
def test_train_step_matches_oracle(resnet_model, n_class, n, hw):
    from oracle.losses import mixed_loss
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    ref, net = _pair(n_class, resnet_model)
    net = net.cuda()
    x, y = synthetic_batch(n, n_class, hw, hw)
    ref.train()
    net.train()
    got = net(x.cuda())
    loss = DiceCrossEntropyLoss(0.5)(got, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    eng = list(net._engines.values())[0]
    left = install_masks_by_call_order(ref, eng, resnet_unet_relu_order(resnet_model))
    want = ref(x)
    assert not left, "ReLU call order of the oracle and the mask list disagree"
    loss_ref = mixed_loss(want, y, 0.5)
    loss_ref.backward()
    assert rel(got.detach().cpu(), want.detach()) <= 8e-2, rel(got.detach().cpu(), want.detach())
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item())
    ref_params = dict(ref.named_parameters())
    worst = []
    for name, p in net.named_parameters():
        r = ref_params[name].grad
        if r is None:            # base_model.fc: unused by the forward in the reference as well
            assert p.grad is None, name
            continue
        assert p.grad is not None, name
        g = p.grad.cpu()
        cos = torch.nn.functional.cosine_similarity(g.flatten(), r.flatten(), dim=0).item()
        worst.append((rel(g, r), cos, name))
    worst.sort(reverse=True)
    assert worst[0][0] <= 1.5e-1 and min(w[1] for w in worst) >= 0.985, worst[:5]
