"""Oracle restatement of `segmentation_models_pytorch.UnetPlusPlus` (TEST INFRASTRUCTURE).

The reference builds its default model from the third-party package
`segmentation_models_pytorch` (unpinned: MMR_EN:DE_CODER/pyproject.toml:16), which is neither
vendored under /root/reference nor installed here:
    SU/ModelTraining.py:247-254   smp.UnetPlusPlus(encoder_name="resnet18",
                                    encoder_weights="imagenet", in_channels=3, classes=C)
    SU/ModelEval.py:331-337       same
    ED/Main_MMR_SegModel.py:589   smp.create_model(**config['model'])  (arch UnetPlusPlus)
This file restates that published architecture in plain PyTorch (fp32, CPU) with the same
module tree, so `state_dict()` keys and shapes are the ones smp checkpoints carry
(SURVEY.md section 8b).  Structure is pinned against the reference's own torchinfo dump
(MMR_EN:DE_CODER/README.md:149-188) in tests/test_oracle_structure.py; numerical parity with
smp itself is unpinned.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision

DECODER_CHANNELS = (256, 128, 64, 32, 16)


class Conv2dReLU(nn.Sequential):
    """smp base/modules.py Conv2dReLU: conv(bias = not use_batchnorm) -> BN -> ReLU."""

    def __init__(self, cin, cout, kernel_size=3, padding=1, use_batchnorm=True):
        conv = nn.Conv2d(cin, cout, kernel_size, padding=padding, bias=not use_batchnorm)
        bn = nn.BatchNorm2d(cout) if use_batchnorm else nn.Identity()
        super().__init__(conv, bn, nn.ReLU(inplace=True))


class DecoderBlock(nn.Module):
    """smp decoders/unetplusplus/decoder.py DecoderBlock (attention_type=None):
    nearest x2 -> cat([x, skip]) -> Conv2dReLU -> Conv2dReLU."""

    def __init__(self, cin, cskip, cout, use_batchnorm=True):
        super().__init__()
        self.conv1 = Conv2dReLU(cin + cskip, cout, 3, 1, use_batchnorm)
        self.attention1 = nn.Identity()
        self.conv2 = Conv2dReLU(cout, cout, 3, 1, use_batchnorm)
        self.attention2 = nn.Identity()

    def forward(self, x, skip=None):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)
        x = self.conv1(x)
        x = self.conv2(x)
        return x


def decoder_block_specs(encoder_channels, decoder_channels=DECODER_CHANNELS):
    """Block name -> (in, skip, out) channels, in construction order (smp
    UnetPlusPlusDecoder.__init__)."""
    enc = list(encoder_channels[1:])[::-1]
    in_ch = [enc[0]] + list(decoder_channels[:-1])
    skip_ch = list(enc[1:]) + [0]
    out_ch = list(decoder_channels)
    specs = {}
    for layer_idx in range(len(in_ch) - 1):
        for depth_idx in range(layer_idx + 1):
            if depth_idx == 0:
                i, s, o = in_ch[layer_idx], skip_ch[layer_idx] * (layer_idx + 1), out_ch[layer_idx]
            else:
                o = skip_ch[layer_idx]
                s = skip_ch[layer_idx] * (layer_idx + 1 - depth_idx)
                i = skip_ch[layer_idx - 1]
            specs["x_%d_%d" % (depth_idx, layer_idx)] = (i, s, o)
    specs["x_0_%d" % (len(in_ch) - 1)] = (in_ch[-1], 0, out_ch[-1])
    return specs


def decoder_schedule(depth=4):
    """Execution order of smp UnetPlusPlusDecoder.forward as a list of
    (block, x_source, [skip sources]) where sources are 'f<k>' (reversed encoder feature k)
    or a block name."""
    sched = []
    for layer_idx in range(depth):
        for depth_idx in range(depth - layer_idx):
            if layer_idx == 0:
                sched.append(("x_%d_%d" % (depth_idx, depth_idx), "f%d" % depth_idx,
                              ["f%d" % (depth_idx + 1)]))
            else:
                L = depth_idx + layer_idx
                skips = ["x_%d_%d" % (i, L) for i in range(depth_idx + 1, L + 1)] + ["f%d" % (L + 1)]
                sched.append(("x_%d_%d" % (depth_idx, L), "x_%d_%d" % (depth_idx, L - 1), skips))
    sched.append(("x_0_%d" % depth, "x_0_%d" % (depth - 1), []))
    return sched


class UnetPlusPlusDecoder(nn.Module):
    def __init__(self, encoder_channels, decoder_channels=DECODER_CHANNELS, use_batchnorm=True):
        super().__init__()
        self.depth = len(decoder_channels) - 1
        self.blocks = nn.ModuleDict({
            name: DecoderBlock(i, s, o, use_batchnorm)
            for name, (i, s, o) in decoder_block_specs(encoder_channels, decoder_channels).items()})

    def forward(self, *features):
        feats = list(features[1:])[::-1]
        dense = {}
        for name, xsrc, skips in decoder_schedule(self.depth):
            get = lambda s: feats[int(s[1:])] if s[0] == "f" else dense[s]
            if skips:
                skip = torch.cat([get(s) for s in skips], dim=1)
                dense[name] = self.blocks[name](get(xsrc), skip)
            else:
                dense[name] = self.blocks[name](get(xsrc))
        return dense["x_0_%d" % self.depth]


class ResNetEncoder(torchvision.models.ResNet):
    """smp encoders/resnet.py ResNetEncoder: torchvision ResNet without fc, returning the six
    features [x, relu(bn1(conv1)), layer1(maxpool), layer2, layer3, layer4]."""

    def __init__(self, name="resnet18"):
        layers = {"resnet18": [2, 2, 2, 2], "resnet34": [3, 4, 6, 3]}[name]
        super().__init__(torchvision.models.resnet.BasicBlock, layers)
        del self.fc
        self.out_channels = (3, 64, 64, 128, 256, 512)

    def forward(self, x):
        feats = [x]
        x = self.relu(self.bn1(self.conv1(x)))
        feats.append(x)
        x = self.layer1(self.maxpool(x))
        feats.append(x)
        for layer in (self.layer2, self.layer3, self.layer4):
            x = layer(x)
            feats.append(x)
        return feats


class UnetPlusPlus(nn.Module):
    """`smp.UnetPlusPlus(encoder_name, encoder_weights, in_channels, classes)` with the
    defaults the reference relies on: encoder_depth 5, decoder_channels (256,128,64,32,16),
    batch-norm decoder, no attention, 3x3 head with bias, identity activation.
    `encoder_weights` must be None here (no network for ImageNet weights; BASELINE configs
    say random-init)."""

    def __init__(self, encoder_name="resnet18", encoder_weights=None, in_channels=3, classes=1):
        super().__init__()
        if encoder_weights is not None:
            raise ValueError("oracle: pretrained encoder weights are unavailable offline")
        if in_channels != 3:
            raise ValueError("oracle: in_channels must be 3")
        self.encoder = ResNetEncoder(encoder_name)
        self.decoder = UnetPlusPlusDecoder(self.encoder.out_channels)
        self.segmentation_head = nn.Sequential(
            nn.Conv2d(DECODER_CHANNELS[-1], classes, 3, padding=1), nn.Identity(), nn.Identity())
        self._init()

    def _init(self):
        # smp base/initialization.py: decoder convs kaiming_uniform(fan_in, relu), BN 1/0;
        # head xavier_uniform, bias 0.  The encoder keeps torchvision's default init.
        for m in self.decoder.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        for m in self.segmentation_head.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        h, w = x.shape[-2:]
        if h % 32 or w % 32:
            raise RuntimeError("Wrong input shape height=%d, width=%d. Expected image height and "
                               "width divisible by 32." % (h, w))
        return self.segmentation_head(self.decoder(*self.encoder(x)))


class UnetDecoder(nn.Module):
    """smp decoders/unet/decoder.py UnetDecoder (ResNet encoder: `center` = Identity, no attention):
    features[1:] reversed; x = head; for every block x = block(x, skip_i or None)."""

    def __init__(self, encoder_channels, decoder_channels=DECODER_CHANNELS, use_batchnorm=True):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]
        in_ch = [enc[0]] + list(decoder_channels[:-1])
        skip_ch = list(enc[1:]) + [0]
        self.center = nn.Identity()
        self.blocks = nn.ModuleList([DecoderBlock(i, s, o, use_batchnorm)
                                     for i, s, o in zip(in_ch, skip_ch, decoder_channels)])

    def forward(self, *features):
        feats = list(features[1:])[::-1]
        x = self.center(feats[0])
        skips = feats[1:]
        for i, block in enumerate(self.blocks):
            x = block(x, skips[i] if i < len(skips) else None)
        return x


class Unet(UnetPlusPlus):
    """`smp.Unet(encoder_name, encoder_weights, in_channels, classes)` (SU/ModelTraining.py:255-262, the
    `--model smp_unet18` branch) with smp's defaults; same initialisation scheme as UnetPlusPlus.  smp is not on
    disk: restated from the published architecture, numerically unpinned."""

    def __init__(self, encoder_name="resnet18", encoder_weights=None, in_channels=3, classes=1):
        super().__init__(encoder_name, encoder_weights, in_channels, classes)
        self.decoder = UnetDecoder(self.encoder.out_channels)
        self._init()


def create_model(arch="UnetPlusPlus", encoder_name="resnet18", encoder_weights=None, in_channels=3,
                 classes=1, **kwargs):
    """smp.create_model as called at ED/Main_MMR_SegModel.py:589."""
    if arch.lower() != "unetplusplus":
        raise KeyError("oracle restates only UnetPlusPlus, got %r" % arch)
    return UnetPlusPlus(encoder_name, encoder_weights, in_channels, classes)


class DeepSupervisionUnetPlusPlus(UnetPlusPlus):
    """BASELINE config 4 'U-Net++ with deep supervision'.  The reference has NO such code
    (README.md:26 is prose only): this definition is ours and parity is
    unpinned.  Extra 3x3 heads (16->classes would need equal widths, so each head takes its
    node's own width) on x_0_1, x_0_2, x_0_3, nearest-upsampled to full resolution; forward
    returns [main, ds3, ds2, ds1]; the loss is the mean over heads."""

    def __init__(self, encoder_name="resnet18", encoder_weights=None, in_channels=3, classes=1):
        super().__init__(encoder_name, encoder_weights, in_channels, classes)
        self.ds_heads = nn.ModuleDict({
            "x_0_1": nn.Conv2d(128, classes, 3, padding=1),
            "x_0_2": nn.Conv2d(64, classes, 3, padding=1),
            "x_0_3": nn.Conv2d(32, classes, 3, padding=1)})
        for m in self.ds_heads.values():
            nn.init.xavier_uniform_(m.weight)
            nn.init.constant_(m.bias, 0)

    def forward(self, x):
        feats = list(self.encoder(x)[1:])[::-1]
        dense = {}
        dec = self.decoder
        for name, xsrc, skips in decoder_schedule(dec.depth):
            get = lambda s: feats[int(s[1:])] if s[0] == "f" else dense[s]
            if skips:
                dense[name] = dec.blocks[name](get(xsrc), torch.cat([get(s) for s in skips], 1))
            else:
                dense[name] = dec.blocks[name](get(xsrc))
        outs = [self.segmentation_head(dense["x_0_4"])]
        for name, scale in (("x_0_3", 2), ("x_0_2", 4), ("x_0_1", 8)):
            outs.append(F.interpolate(self.ds_heads[name](dense[name]), scale_factor=scale,
                                      mode="nearest"))
        return outs
