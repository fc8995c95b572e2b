"""Shared test helpers: synthetic inputs, oracle/candidate model pairs, and the
mask-matched ReLU that lets an fp32 oracle back-propagate through the same ReLU sign pattern as
the bf16 plan engine."""
import torch


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def synthetic_batch(n, classes, h, w, seed=6210):
    """frames: rand in [0,1] normalised with the ImageNet mean/std the reference uses
    (SU/ModelTraining.py:300-301); labels: randint(0, classes)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand((n, 3, h, w), generator=g)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    return (u - mean) / std, torch.randint(0, classes, (n, h, w), generator=g)


def model_pair(classes, encoder="resnet18", seed=6210, randomize_bn=True):
    """(oracle fp32 CPU model, plan model on cuda) with identical weights."""
    from oracle.unetpp import UnetPlusPlus as OracleNet
    from mmrseg_b200.models import UnetPlusPlus
    torch.manual_seed(seed)
    ref = OracleNet(encoder, None, 3, classes)
    if randomize_bn:
        g = torch.Generator().manual_seed(seed + 1)
        for m in ref.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.data.uniform_(0.5, 1.5, generator=g)
                m.bias.data.normal_(0, 0.2, generator=g)
                m.running_mean.normal_(0, 0.2, generator=g)
                m.running_var.uniform_(0.5, 1.5, generator=g)
    net = UnetPlusPlus(encoder, classes=classes)
    net.load_state_dict(ref.state_dict(), strict=True)
    return ref, net.cuda()


class _MaskedReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask):
        ctx.save_for_backward(mask)
        return torch.relu(x)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * mask, None


class ReLUWithMasks(torch.nn.Module):
    """ReLU whose backward uses externally supplied masks, consumed in call order."""

    def __init__(self, masks):
        super().__init__()
        self.masks = list(masks)

    def forward(self, x):
        return _MaskedReLU.apply(x, self.masks.pop(0))


def install_engine_masks(ref, eng):
    """Make the oracle U-Net++ back-propagate through the ReLU sign pattern of the engine's stored
    activations (a forward error eps flips ~0.8*eps of the masks, which alone moves fp32 gradients
    by sqrt(0.8*eps) per layer; matching the masks isolates the arithmetic of the kernels)."""
    m = lambda name: (eng.acts[name].buf.float().permute(0, 3, 1, 2).cpu() > 0).float()
    ref.encoder.relu = ReLUWithMasks([m("f_stem")])
    for li in range(1, 5):
        for bi, blk in enumerate(getattr(ref.encoder, "layer%d" % li)):
            base = "encoder.layer%d.%d." % (li, bi)
            blk.relu = ReLUWithMasks([m(base + "t1"), m(base + "out")])
    for name, blk in ref.decoder.blocks.items():
        blk.conv1[2] = ReLUWithMasks([m(name + ".mid")])
        blk.conv2[2] = ReLUWithMasks([m(name)])


def install_masks_by_call_order(ref, eng, act_names):
    """Generic version of install_engine_masks: every nn.ReLU of `ref` (shared module objects hooked
    once) back-propagates through the engine's ReLU sign pattern; `act_names` lists the engine
    activations in the order the oracle's forward calls its ReLUs."""
    masks = [(eng.acts[n].buf.float().permute(0, 3, 1, 2).cpu() > 0).float() for n in act_names]
    queue = list(masks)

    def hook(mod, inp, out):
        return _MaskedReLU.apply(inp[0], queue.pop(0))

    seen = set()
    for m in ref.modules():
        if isinstance(m, torch.nn.ReLU) and id(m) not in seen:
            seen.add(id(m))
            m.inplace = False
            m.register_forward_hook(hook)
    return queue


def resnet_unet_relu_order(resnet_model):
    """Engine activation names in the order oracle.resnet_unet.ResNetUNet.forward applies ReLU."""
    layers = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3)}[resnet_model]
    names = ["xo0", "xo1", "f_stem"]
    for li, n in enumerate(layers, start=1):
        for bi in range(n):
            base = "base_model.layer%d.%d." % (li, bi)
            names += [base + "t1", base + "out"]
    names += ["l4p", "l3p", "d3", "l2p", "d2", "l1p", "d1", "l0p", "d0", "dfull"]
    return names


class _RoutedMaxPool2(torch.autograd.Function):
    """MaxPool2d(2) whose backward routes through externally supplied window positions (0..3)."""

    @staticmethod
    def forward(ctx, x, pos):
        ctx.save_for_backward(pos)
        ctx.shape = x.shape
        return torch.nn.functional.max_pool2d(x, 2)

    @staticmethod
    def backward(ctx, g):
        (pos,) = ctx.saved_tensors
        n, c, h, w = ctx.shape
        gin = torch.zeros((n, c, h, w), dtype=g.dtype)
        ho, wo = g.shape[2], g.shape[3]
        for k in range(4):
            dy, dx = k // 2, k % 2
            gin[:, :, dy:2 * ho:2, dx:2 * wo:2] = g * (pos == k)
        return gin, None


def install_pool_routes(ref, eng, pool_names):
    """Every nn.MaxPool2d of `ref`, in call order, back-propagates through the engine's recorded argmax
    positions: with bf16 activations near-ties inside a window resolve differently than in fp32, which
    moves gradients the same way a flipped ReLU mask does (see install_engine_masks)."""
    queue = [eng.acts[n].idx.permute(0, 3, 1, 2).cpu().long() for n in pool_names]

    def hook(mod, inp, out):
        return _RoutedMaxPool2.apply(inp[0], queue.pop(0))

    for m in ref.modules():
        if isinstance(m, torch.nn.MaxPool2d):
            m.register_forward_hook(hook)
    return queue
