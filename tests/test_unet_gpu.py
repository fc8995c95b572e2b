"""The reference's in-tree UNet (SU/UArchModel/unet.py, constructed `UNet(3, C, bilinear=True)` at
SU/ModelTraining.py:242) through the plan engine, against the oracle restatement (oracle/unet.py, itself
pinned to the reference's own files by tests/golden/unet_reference.npz).

New pieces on this path: MaxPool2d(2) (bit-exact forward, exact routing backward), 3x3 convs WITH bias in
front of BatchNorm (bias folded into the BN shift in eval mode; in training the batch mean absorbs it and
its gradient is identically zero), skip-first concat order [skip, nearest-x2(x)], a 1x1 head.
Tolerances as for U-Net++ (tests/test_model_gpu.py): eval logits 2e-2, train logits 8e-2, loss 2e-3,
gradients 1.5e-1 / cosine 0.985 against the fp32 oracle back-propagating through the engine's ReLU masks
and max-pool argmax positions (near-ties inside a window resolve differently in bf16);
conv biases in front of BatchNorm: |grad| <= 1e-3 of the weight-gradient scale on both sides.
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import install_masks_by_call_order, install_pool_routes, rel, synthetic_batch


def _pair(n_classes, seed=6210, bilinear=True):
    from oracle.unet import UNet as OracleNet
    from mmrseg_b200.models import UNet
    torch.manual_seed(seed)
    ref = OracleNet(3, n_classes, bilinear=bilinear)
    g = torch.Generator().manual_seed(seed + 1)
    for m in ref.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.uniform_(0.5, 1.5, generator=g)
            m.bias.data.normal_(0, 0.2, generator=g)
            m.running_mean.normal_(0, 0.2, generator=g)
            m.running_var.uniform_(0.5, 1.5, generator=g)
    net = UNet(3, n_classes, bilinear=bilinear)
    net.load_state_dict(ref.state_dict(), strict=True)
    return ref, net


RELU_ORDER = ["x1.mid", "x1", "x2.mid", "x2", "x3.mid", "x3", "x4.mid", "x4", "x5.mid", "x5",
              "u1.mid", "u1", "u2.mid", "u2", "u3.mid", "u3", "u4.mid", "u4"]


def test_state_dict_keys_and_errors():
    from mmrseg_b200.models import UNet
    ref, net = _pair(3)
    assert list(net.state_dict().keys()) == list(ref.state_dict().keys())
    assert net.outc.conv.weight.shape == (3, 64, 1, 1)
    with pytest.raises(NotImplementedError):
        UNet(1, 2, bilinear=True)
    ref_t, net_t = _pair(3, bilinear=False)      # ConvTranspose2d variant (unet_parts.py:269)
    assert list(net_t.state_dict().keys()) == list(ref_t.state_dict().keys())
    assert net_t.up1.up.weight.shape == (1024, 512, 2, 2) and net_t.down4.maxpool_conv[1].double_conv[0].out_channels == 1024


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 16, 24, 64), (1, 9, 7, 16), (3, 8, 8, 128)])
def test_maxpool2x2_forward_and_backward(shape):
    from mmrseg_b200 import _lib
    from mmrseg_b200._lib import MmrContrib
    lib = _lib.lib()
    N, H, W, Cc = shape
    g = torch.Generator(device="cuda").manual_seed(9)
    # coarse values: many exact ties inside a window, the first one in scan order must win
    x = (torch.randint(-3, 4, (N, H, W, Cc), generator=g, device="cuda").float() * 0.5).to(torch.bfloat16)
    Ho, Wo = H // 2, W // 2
    out = torch.empty((N, Ho, Wo, Cc), device="cuda", dtype=torch.bfloat16)
    idx = torch.empty((N, Ho, Wo, Cc), device="cuda", dtype=torch.uint8)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(lib.mmr_maxpool2x2s2_fwd(p(x), N, H, W, Cc, p(out), p(idx), s))
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref, ref_idx = F.max_pool2d(xr, 2, return_indices=True)
    assert torch.equal(out.float().permute(0, 3, 1, 2), ref.detach())
    # torch's flat index -> window position
    pos = ((ref_idx // W) % 2) * 2 + (ref_idx % W) % 2
    assert torch.equal(idx.permute(0, 3, 1, 2).long(), pos)
    g1 = torch.randn((N, Ho, Wo, Cc), generator=g, device="cuda").to(torch.bfloat16)
    g2 = torch.randn((N, Ho, Wo, Cc), generator=g, device="cuda").to(torch.bfloat16)
    arr = (MmrContrib * 2)()
    arr[0].ptr, arr[0].pool2, arr[1].ptr, arr[1].pool2 = g1.data_ptr(), 0, g2.data_ptr(), 0
    gin = torch.full((N, H, W, Cc), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.mmr_maxpool2x2s2_bwd(arr, 2, p(idx), N, H, W, Cc, p(gin), s))
    torch.cuda.synchronize()
    gsum = (g1.float() + g2.float()).to(torch.bfloat16).float()      # the kernel adds in fp32, stores bf16
    ref.backward(gsum.permute(0, 3, 1, 2))
    assert torch.equal(gin.float().permute(0, 3, 1, 2), xr.grad)


@pytest.mark.gpu
def test_eval_forward_matches_oracle_and_reference_golden():
    import os
    ref, net = _pair(3)
    net = net.cuda().eval()
    ref.eval()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "unet_reference.npz"))
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        got = net(x.cuda()).cpu()
        want = ref(x)
    assert np.allclose(want.numpy(), g["logits_eval"], atol=1e-5)    # the pair IS the golden model
    assert rel(got, want) <= 2e-2, rel(got, want)
    assert rel(got, torch.from_numpy(g["logits_eval"])) <= 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("n_classes,n,h,w", [(3, 2, 32, 48), (10, 2, 128, 128)])
def test_train_step_matches_oracle(n_classes, n, h, w):
    from oracle.losses import mixed_loss
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    ref, net = _pair(n_classes)
    net = net.cuda()
    x, y = synthetic_batch(n, n_classes, h, w)
    ref.train()
    net.train()
    got = net(x.cuda())
    loss = DiceCrossEntropyLoss(0.5)(got, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    eng = list(net._engines.values())[0]
    left = install_masks_by_call_order(ref, eng, RELU_ORDER)
    left_p = install_pool_routes(ref, eng, ["p1", "p2", "p3", "p4"])
    want = ref(x)
    assert not left and not left_p, "ReLU / max-pool call order of the oracle and the engine lists disagree"
    loss_ref = mixed_loss(want, y, 0.5)
    loss_ref.backward()
    assert rel(got.detach().cpu(), want.detach()) <= 8e-2, rel(got.detach().cpu(), want.detach())
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item())
    ref_params = dict(ref.named_parameters())
    worst = []
    for name, p in net.named_parameters():
        r = ref_params[name].grad
        gr = p.grad.cpu()
        if name.endswith("double_conv.0.bias") or name.endswith("double_conv.3.bias"):
            # a bias in front of training-mode BatchNorm: zero gradient analytically; the oracle's is rounding noise
            wname = name[:-4] + "weight"
            scale = ref_params[wname].grad.abs().max().item()
            assert gr.abs().max().item() <= 1e-3 * scale and r.abs().max().item() <= 1e-3 * scale, name
            continue
        cos = torch.nn.functional.cosine_similarity(gr.flatten(), r.flatten(), dim=0).item()
        worst.append((rel(gr, r), cos, name))
    worst.sort(reverse=True)
    assert worst[0][0] <= 1.5e-1 and min(w_[1] for w_ in worst) >= 0.985, worst[:5]
    # running statistics include the conv bias, as torch's do
    rb = dict(ref.named_buffers())
    for name, b in net.named_buffers():
        if name.endswith("running_mean"):
            assert rel(b.cpu(), rb[name]) <= 2e-2, name


@pytest.mark.gpu
def test_convtranspose_variant_matches_oracle_and_reference_golden():
    """UNet(3, C, bilinear=False): ConvTranspose2d(in, in // 2, 2, stride=2) upsampling (unet_parts.py:269) as the
    data gradient of a 2x2 stride-2 conv on the tcgen05 kernels.  Eval logits against the oracle and against the
    golden logits produced by RUNNING the reference's own UNet(3, 3, bilinear=False) (oracle/make_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "unet_convt_reference.npz"))
    from tests.test_oracle_cpu import _oracle_unet_like_golden
    from mmrseg_b200.models import UNet
    ref = _oracle_unet_like_golden(bilinear=False)
    net = UNet(3, 3, bilinear=False)
    net.load_state_dict(ref.state_dict(), strict=True)
    net = net.cuda().eval()
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        got = net(x.cuda()).cpu()
    assert rel(got, torch.from_numpy(g["logits_eval"])) <= 2e-2
    net.train()
    with torch.no_grad():
        got_t = net(x.cuda()).cpu()
    assert rel(got_t, torch.from_numpy(g["logits_train"])) <= 8e-2


@pytest.mark.gpu
def test_convtranspose_variant_teacher_forced(monkeypatch):
    """Every unit of the ConvTranspose2d U-Net, forward and backward, recomputed in fp32 from the engine's own
    operands (tests/teacher.py): transposed-conv output / data gradient to one bf16 ulp flip, weight and bias
    gradients to 1e-4."""
    from tests.test_parity_gpu import _teacher_forced
    _, net = _pair(10, bilinear=False)
    x, y = synthetic_batch(2, 10, 96, 64)
    worst = _teacher_forced(net.cuda(), x, y, monkeypatch)
    print("teacher-forced UNet(bilinear=False): worst", worst)
