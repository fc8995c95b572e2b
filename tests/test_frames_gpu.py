"""uint8 HWC frames straight into the model: `model(frames_u8)` with
`set_input_normalization(mean, std)` replaces the reference's CPU path ToTensor (/255) ->
utils.normalize (SU/utils.py:480-519, pinned by tests/golden/normalize_reference.npz) -> fp32 H2D copy.
The oracle is the reference's own arithmetic on the CPU (oracle.resnet_unet.normalize) fed to the fp32
oracle network; the two device paths (uint8 frames vs the pre-normalised fp32 tensor) must agree to bf16
rounding of the first layer's operands."""
import ctypes as C

import pytest
import torch

from tests.helpers import model_pair, rel

pytestmark = pytest.mark.gpu

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def _frames(n, h, w, seed=6210):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8)


def _reference_preprocess(frames):
    from oracle.resnet_unet import normalize
    x = frames.permute(0, 3, 1, 2).float() / 255.0          # torchvision ToTensor
    return normalize(x, torch.tensor(MEAN), torch.tensor(STD))


def test_stem_im2col_u8_matches_float_path():
    from mmrseg_b200 import _lib as L
    lib = L.lib()
    frames = _frames(2, 64, 96)
    x = _reference_preprocess(frames).cuda().contiguous()
    fu = frames.cuda()
    n, h, w = 2, 64, 96
    ho, wo, kpad = h // 2, w // 2, 160
    a = torch.empty((n * ho * wo, kpad), device="cuda", dtype=torch.bfloat16)
    b = torch.empty_like(a)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    mean, std = torch.tensor(MEAN, device="cuda"), torch.tensor(STD, device="cuda")
    L.check(lib.mmr_stem_im2col(C.c_void_p(x.data_ptr()), n, h, w, C.c_void_p(a.data_ptr()), kpad, None, None, s))
    L.check(lib.mmr_stem_im2col_u8(C.c_void_p(fu.data_ptr()), n, h, w, C.c_void_p(b.data_ptr()), kpad,
                                   C.c_void_p(mean.data_ptr()), C.c_void_p(std.data_ptr()), s))
    torch.cuda.synchronize()
    # reference im2col of the normalised image: column = c*49 + ky*7 + kx
    cols = torch.nn.functional.unfold(x.cpu(), 7, padding=3, stride=2)           # [n, 147, ho*wo]
    want = cols.permute(0, 2, 1).reshape(n * ho * wo, 147)
    assert torch.equal(a[:, 147:].float().cpu(), torch.zeros(n * ho * wo, kpad - 147))
    assert (a[:, :147].float().cpu() - want).abs().max().item() <= 0.02           # bf16 rounding of |x| <= 2.7
    # the two device paths differ by at most one bf16 ulp (x * (1/255) vs x / 255 before rounding)
    assert (a.float() - b.float()).abs().max().item() <= 2.0 ** -6
    assert (a != b).float().mean().item() < 0.02


@pytest.mark.parametrize("arch", ["unetpp", "resnet_unet"])
def test_uint8_frames_equal_normalised_float_input(arch):
    frames = _frames(2, 64, 96)
    x = _reference_preprocess(frames)
    if arch == "unetpp":
        ref, net = model_pair(3)
    else:
        from mmrseg_b200.models import ResNetUNet
        torch.manual_seed(6210)
        net = ResNetUNet(4, 18).cuda()
        ref = None
    net.set_input_normalization(MEAN, STD)
    net.eval()
    with torch.no_grad():
        got_f = net(x.cuda()).float().cpu()
        got_u = net(frames.cuda()).float().cpu()
    assert rel(got_u, got_f) <= 5e-3, rel(got_u, got_f)
    if ref is not None:
        ref.eval()
        with torch.no_grad():
            want = ref(x)
        assert rel(got_u, want) <= 2e-2, rel(got_u, want)


def test_uint8_frames_train_step():
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    _, net = model_pair(2)
    frames = _frames(2, 64, 64)
    y = torch.randint(0, 2, (2, 64, 64), generator=torch.Generator().manual_seed(1))
    net.set_input_normalization(MEAN, STD)
    net.train()
    grads = []
    for inp in (_reference_preprocess(frames).cuda(), frames.cuda()):
        for p in net.parameters():
            p.grad = None
        loss = DiceCrossEntropyLoss(0.5)(net(inp), y.cuda())
        loss.backward()
        torch.cuda.synchronize()
        grads.append({k: p.grad.clone() for k, p in net.named_parameters()})
    worst = max(rel(grads[1][k], grads[0][k]) for k in grads[0])
    assert worst <= 5e-2, worst


def test_device_prefetcher_order_and_values():
    from mmrseg_b200.data import DevicePrefetcher
    g = torch.Generator().manual_seed(3)
    host = [(torch.randint(0, 256, (2, 8, 8, 3), generator=g, dtype=torch.uint8).pin_memory(),
             torch.randint(0, 5, (2, 8, 8), generator=g).pin_memory(), i) for i in range(5)]
    seen = 0
    for (x, y, i), (hx, hy, hi) in zip(DevicePrefetcher(host), host):
        assert x.is_cuda and y.is_cuda and i == hi
        assert torch.equal(x.cpu(), hx) and torch.equal(y.cpu(), hy)
        seen += 1
    assert seen == 5 and len(DevicePrefetcher(host)) == 5
    assert list(DevicePrefetcher([])) == []


@pytest.mark.parametrize("u8", [False, True])
def test_stem_space_to_depth_matches_conv2d(u8):
    """The 7x7 stride-2 pad-3 stem as a 3x3 conv over 4x4 pixel blocks (mmr_stem_s2d_pack / _weights + the halo
    kernel with strided store groups) against F.conv2d on the same bf16-rounded operands: 4e-3 x 8 of the output
    RMS, as for every conv kernel; fp32 NCHW and uint8 HWC inputs."""
    import torch.nn.functional as F
    from mmrseg_b200 import _lib as L
    from mmrseg_b200 import convplan
    lib = L.lib()
    n, h, w, cout = 2, 64, 96, 64
    g = torch.Generator().manual_seed(5)
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    s2d = torch.empty((n, h // 4, w // 4, 64), device="cuda", dtype=torch.bfloat16)
    if u8:
        frames = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8)
        x = _reference_preprocess(frames)
        mean, std = torch.tensor(MEAN, device="cuda"), torch.tensor(STD, device="cuda")
        L.check(lib.mmr_stem_s2d_pack(p(frames.cuda()), 1, n, h, w, p(s2d), p(mean), p(std), s))
    else:
        x = torch.randn((n, 3, h, w), generator=g)
        L.check(lib.mmr_stem_s2d_pack(p(x.cuda()), 0, n, h, w, p(s2d), None, None, s))
    w7 = (torch.randn((cout, 3, 7, 7), generator=g) / 12).cuda()
    w3 = torch.empty((4 * cout, 64, 3, 3), device="cuda")
    L.check(lib.mmr_stem_s2d_weights(p(w7), cout, p(w3), s))
    # every original tap appears exactly once per output phase
    assert torch.allclose(w3.view(4, cout, -1).sum(2), w7.view(cout, -1).sum(1).expand(4, cout), atol=1e-4)
    out = torch.full((n, h // 2, w // 2, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    cfg = convplan.fprop_halo_cfg([(s2d, 1)], 4 * cout, force={"rph": 1, "sg": 64})
    packed = convplan.pack_weights_halo(w3, cfg, 0)
    groups = [(out, 0, 2, q >> 1, q & 1) for q in range(4)]
    plan = convplan.build_halo(cfg, [(s2d, 1)], packed, groups, n, h // 4, w // 4, 4 * cout)
    plan.run()
    torch.cuda.synchronize()
    ref = F.conv2d(x.to(torch.bfloat16).float(), w7.cpu().to(torch.bfloat16).float(), None, 2, 3)
    got = out.float().permute(0, 3, 1, 2).cpu()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err <= 4e-3 * 8 * ref.pow(2).mean().sqrt().item() + (0.02 if u8 else 0.0), err
