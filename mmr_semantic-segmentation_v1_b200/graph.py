"""Static dataflow graphs of the networks on the hot path.

A graph is a list of op dicts in forward execution order.  Tensor names are strings; every
op writes exactly one tensor.  Parameter names are `state_dict` prefixes of the reference
models, so the engine can bind the reference's own parameter tensors.

  {'op': 'stem',    'out', 'conv', 'bn'}                         7x7 s2 p3 conv + BN + ReLU on the image
  {'op': 'stem3',   'out', 'conv', 'cout', 'relu'}               3x3 s1 p1 conv(+bias)+ReLU on the image
  {'op': 'pack16',  'out'}                                        image -> NHWC bf16 padded to 16 channels
  {'op': 'upsample','in', 'out'}                                  bilinear x2, align_corners=True
  {'op': 'maxpool', 'in', 'out', 'k'}                             3x3 s2 p1 (k absent or 3) or 2x2 s2 (k = 2)
  {'op': 'conv',    'out', 'conv', 'src': [(tensor, up)], 'k', 's', 'cout',
                    'bn': name|None, 'bias': bool, 'relu': bool, 'res': tensor|None}
  {'op': 'head',    'out', 'conv', 'src': [(tensor, 1)], 'k', 'cout'}   conv + bias -> fp32 NCHW logits

U-Net++ follows smp's UnetPlusPlusDecoder.forward order (call sites
SU/ModelTraining.py:247-254, ED/Main_MMR_SegModel.py:589); the encoder is the torchvision
BasicBlock ResNet that smp wraps.
"""

RESNET_LAYERS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3)}
DECODER_CHANNELS = (256, 128, 64, 32, 16)


def resnet_encoder_ops(prefix, layers, ops):
    """Appends stem, maxpool and the four BasicBlock stages; returns the feature names
    [stem, layer1, layer2, layer3, layer4] and their channel counts."""
    ops.append({"op": "stem", "out": "f_stem", "conv": prefix + "conv1", "bn": prefix + "bn1", "relu": True})
    ops.append({"op": "maxpool", "in": "f_stem", "out": "pool"})
    x, cin = "pool", 64
    feats, chans = ["f_stem"], [64]
    for li, (nblk, cout) in enumerate(zip(layers, (64, 128, 256, 512)), start=1):
        for bi in range(nblk):
            s = 2 if (bi == 0 and li > 1) else 1
            base = "%slayer%d.%d." % (prefix, li, bi)
            t1 = base + "t1"
            ops.append({"op": "conv", "out": t1, "conv": base + "conv1", "src": [(x, 1)], "k": 3, "s": s,
                        "cout": cout, "bn": base + "bn1", "bias": False, "relu": True, "res": None})
            idn = x
            if s != 1 or cin != cout:
                idn = base + "idn"
                ops.append({"op": "conv", "out": idn, "conv": base + "downsample.0", "src": [(x, 1)],
                            "k": 1, "s": s, "cout": cout, "bn": base + "downsample.1", "bias": False,
                            "relu": False, "res": None})
            out = base + "out"
            ops.append({"op": "conv", "out": out, "conv": base + "conv2", "src": [(t1, 1)], "k": 3,
                        "s": 1, "cout": cout, "bn": base + "bn2", "bias": False, "relu": True,
                        "res": idn})
            x, cin = out, cout
        feats.append(x)
        chans.append(cout)
    return feats, chans


def decoder_block_specs(encoder_channels, decoder_channels=DECODER_CHANNELS):
    """Block name -> (in, skip, out) channels (smp UnetPlusPlusDecoder.__init__)."""
    enc = list(encoder_channels[1:])[::-1]
    in_ch = [enc[0]] + list(decoder_channels[:-1])
    skip_ch = list(enc[1:]) + [0]
    out_ch = list(decoder_channels)
    specs = {}
    for layer_idx in range(len(in_ch) - 1):
        for depth_idx in range(layer_idx + 1):
            if depth_idx == 0:
                spec = (in_ch[layer_idx], skip_ch[layer_idx] * (layer_idx + 1), out_ch[layer_idx])
            else:
                spec = (skip_ch[layer_idx - 1], skip_ch[layer_idx] * (layer_idx + 1 - depth_idx),
                        skip_ch[layer_idx])
            specs["x_%d_%d" % (depth_idx, layer_idx)] = spec
    specs["x_0_%d" % (len(in_ch) - 1)] = (in_ch[-1], 0, out_ch[-1])
    return specs


def decoder_schedule(depth=4):
    """(block, x source, [skip sources]) in smp's execution order; 'f<k>' is the k-th feature of
    the reversed encoder pyramid (f0 deepest)."""
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


def unetpp_graph(encoder_name="resnet18", classes=2, deep_supervision=False):
    ops = []
    feats, chans = resnet_encoder_ops("encoder.", RESNET_LAYERS[encoder_name], ops)
    rev = feats[::-1]  # f0 = layer4 ... f4 = stem
    specs = decoder_block_specs([3] + chans)
    name_of = lambda s: rev[int(s[1:])] if s[0] == "f" else s
    for blk, xsrc, skips in decoder_schedule(4):
        cin, cskip, cout = specs[blk]
        src = [(name_of(xsrc), 2)] + [(name_of(s), 1) for s in skips]
        base = "decoder.blocks.%s." % blk
        ops.append({"op": "conv", "out": blk + ".mid", "conv": base + "conv1.0", "src": src, "k": 3,
                    "s": 1, "cout": cout, "bn": base + "conv1.1", "bias": False, "relu": True,
                    "res": None})
        ops.append({"op": "conv", "out": blk, "conv": base + "conv2.0", "src": [(blk + ".mid", 1)],
                    "k": 3, "s": 1, "cout": cout, "bn": base + "conv2.1", "bias": False, "relu": True,
                    "res": None})
    ops.append({"op": "head", "out": "logits", "conv": "segmentation_head.0", "src": [("x_0_4", 1)],
                "k": 3, "cout": classes})
    if deep_supervision:
        for node, up in (("x_0_3", 2), ("x_0_2", 4), ("x_0_1", 8)):
            ops.append({"op": "head", "out": "logits." + node, "conv": "ds_heads." + node,
                        "src": [(node, 1)], "k": 3, "cout": classes, "up": up})
    return ops


def smp_unet_graph(encoder_name="resnet18", classes=2):
    """`smp.Unet` (the reference's `--model smp_unet18`, SU/ModelTraining.py:255-262): the plain U-Net decoder on the
    same ResNet encoder -- five DecoderBlocks in a chain, block i = nearest x2 of the previous output ->
    cat([x, skip_i]) -> Conv2dReLU -> Conv2dReLU with (in, skip, out) = (512, 256, 256), (256, 128, 128),
    (128, 64, 64), (64, 64, 32), (32, 0, 16); `center` is the identity for ResNet encoders; 3x3 head with bias.
    Parameter names are smp's (`decoder.blocks.<i>.conv1.0.weight`, ...)."""
    ops = []
    feats, chans = resnet_encoder_ops("encoder.", RESNET_LAYERS[encoder_name], ops)
    rev = feats[::-1]                 # layer4, layer3, layer2, layer1, stem
    x = rev[0]
    for i, cout in enumerate(DECODER_CHANNELS):
        src = [(x, 2)] + ([(rev[i + 1], 1)] if i + 1 < len(rev) else [])
        base = "decoder.blocks.%d." % i
        ops.append({"op": "conv", "out": "d%d.mid" % i, "conv": base + "conv1.0", "src": src, "k": 3, "s": 1,
                    "cout": cout, "bn": base + "conv1.1", "bias": False, "relu": True, "res": None})
        ops.append({"op": "conv", "out": "d%d" % i, "conv": base + "conv2.0", "src": [("d%d.mid" % i, 1)], "k": 3,
                    "s": 1, "cout": cout, "bn": base + "conv2.1", "bias": False, "relu": True, "res": None})
        x = "d%d" % i
    ops.append({"op": "head", "out": "logits", "conv": "segmentation_head.0", "src": [(x, 1)], "k": 3,
                "cout": classes})
    return ops


def resnet_unet_graph(resnet_model=18, n_class=10):
    """The reference's in-tree ResNet encoder/decoder (SU/UArchModel/resnet_unet.py:134-300):
    torchvision BasicBlock encoder, 1x1 conv+ReLU laterals, bilinear x2 (align_corners=True)
    -> cat([upsampled, lateral]) -> 3x3 conv + bias + ReLU decoder without BatchNorm, a full-resolution
    side path of two 3x3 convs, and a 1x1 `conv_last`.  Parameter names are the reference's."""
    ops = [{"op": "pack16", "out": "image16"}]

    def conv(out, name, src, k, cout, cin_pad=False):
        ops.append({"op": "conv", "out": out, "conv": name, "src": src, "k": k, "s": 1, "cout": cout,
                    "bn": None, "bias": True, "relu": True, "res": None, "cin_pad": cin_pad})

    conv("xo0", "conv_original_size0.0", [("image16", 1)], 3, 64, cin_pad=True)
    conv("xo1", "conv_original_size1.0", [("xo0", 1)], 3, 64)
    feats, chans = resnet_encoder_ops("base_model.", RESNET_LAYERS["resnet%d" % resnet_model], ops)
    l0, l1, l2, l3, l4 = feats
    conv("l4p", "layer4_1x1.0", [(l4, 1)], 1, 512)
    ops.append({"op": "upsample", "in": "l4p", "out": "u4"})
    conv("l3p", "layer3_1x1.0", [(l3, 1)], 1, 256)
    conv("d3", "conv_up3.0", [("u4", 1), ("l3p", 1)], 3, 512)
    ops.append({"op": "upsample", "in": "d3", "out": "u3"})
    conv("l2p", "layer2_1x1.0", [(l2, 1)], 1, 128)
    conv("d2", "conv_up2.0", [("u3", 1), ("l2p", 1)], 3, 256)
    ops.append({"op": "upsample", "in": "d2", "out": "u2"})
    conv("l1p", "layer1_1x1.0", [(l1, 1)], 1, 64)
    conv("d1", "conv_up1.0", [("u2", 1), ("l1p", 1)], 3, 256)
    ops.append({"op": "upsample", "in": "d1", "out": "u1"})
    conv("l0p", "layer0_1x1.0", [(l0, 1)], 1, 64)
    conv("d0", "conv_up0.0", [("u1", 1), ("l0p", 1)], 3, 128)
    ops.append({"op": "upsample", "in": "d0", "out": "u0"})
    conv("dfull", "conv_original_size2.0", [("u0", 1), ("xo1", 1)], 3, 64)
    ops.append({"op": "head", "out": "logits", "conv": "conv_last", "src": [("dfull", 1)], "k": 1,
                "cout": n_class})
    return ops


def unet_graph(n_classes=2, bilinear=True):
    """The reference's in-tree `UNet(n_channels=3, n_classes, bilinear=True)` (SU/UArchModel/unet.py:104-245,
    unet_parts.py): DoubleConv = (3x3 conv WITH bias -> BatchNorm -> ReLU) x 2, Down = MaxPool2d(2) +
    DoubleConv, Up = nearest x2 (the `bilinear=True` branch builds nn.Upsample(mode='nearest')) ->
    cat([skip, upsampled]) -> DoubleConv(in, out, mid = in // 2), OutConv = 1x1 conv with bias.
    `F.pad` to the skip's size is the identity when H and W are multiples of 16.
    bilinear=False (unet_parts.py:269; unet.py:153-163): Up = ConvTranspose2d(in, in // 2, kernel_size=2, stride=2)
    -> cat([skip, upsampled]) -> DoubleConv(in, out), and down4 widens to 1024 channels
      {'op': 'convt', 'in', 'out', 'conv', 'cout'}          2x2 stride-2 transposed conv with bias"""
    ops = [{"op": "pack16", "out": "image16"}]

    def double_conv(prefix, src, mid, cout, out, cin_pad=False):
        ops.append({"op": "conv", "out": out + ".mid", "conv": prefix + "double_conv.0", "src": src, "k": 3, "s": 1,
                    "cout": mid, "bn": prefix + "double_conv.1", "bias": True, "relu": True, "res": None,
                    "cin_pad": cin_pad})
        ops.append({"op": "conv", "out": out, "conv": prefix + "double_conv.3", "src": [(out + ".mid", 1)], "k": 3,
                    "s": 1, "cout": cout, "bn": prefix + "double_conv.4", "bias": True, "relu": True, "res": None})

    double_conv("inc.", [("image16", 1)], 64, 64, "x1", cin_pad=True)
    prev = "x1"
    factor = 2 if bilinear else 1
    for i, cout in enumerate((128, 256, 512, 1024 // factor), start=1):
        ops.append({"op": "maxpool", "in": prev, "out": "p%d" % i, "k": 2})
        double_conv("down%d.maxpool_conv.1." % i, [("p%d" % i, 1)], cout, cout, "x%d" % (i + 1))
        prev = "x%d" % (i + 1)
    for i, (skip, cin, cout) in enumerate((("x4", 1024, 512 // factor), ("x3", 512, 256 // factor),
                                           ("x2", 256, 128 // factor), ("x1", 128, 64)), start=1):
        if bilinear:
            double_conv("up%d.conv." % i, [(skip, 1), (prev, 2)], cin // 2, cout, "u%d" % i)
        else:
            ops.append({"op": "convt", "in": prev, "out": "t%d" % i, "conv": "up%d.up" % i, "cout": cin // 2})
            double_conv("up%d.conv." % i, [(skip, 1), ("t%d" % i, 1)], cout, cout, "u%d" % i)
        prev = "u%d" % i
    ops.append({"op": "head", "out": "logits", "conv": "outc.conv", "src": [(prev, 1)], "k": 1,
                "cout": n_classes})
    return ops


def graph_param_names(ops):
    """state_dict names of the parameters a graph reads (their gradients are produced by the plan)."""
    names = []
    for op in ops:
        if op["op"] in ("conv", "head", "stem", "convt"):
            names.append(op["conv"] + ".weight")
            if op["op"] in ("head", "convt") or op.get("bias"):
                names.append(op["conv"] + ".bias")
            if op.get("bn"):
                names += [op["bn"] + ".weight", op["bn"] + ".bias"]
    return names
