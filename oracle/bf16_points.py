"""Oracle run "with the same bf16 rounding points" (TEST INFRASTRUCTURE).

The product computes in bf16 with fp32 accumulation.  Against the all-fp32 oracle a 40-layer
ReLU network at random init shows a few per cent of forward error and, because a forward error
eps flips about 0.8*eps of the ReLU masks, tens of per cent of gradient error — for any bf16
implementation.  To separate that arithmetic from bugs, this module re-runs the oracle's own
modules in fp32 while rounding to bf16 at exactly the places where the plan engine stores bf16:

  forward : image, conv weights, every conv output z, every activation after BN(+residual)+ReLU
  backward: every per-consumer data gradient, every gathered gradient g, every dz

Everything else (convolution sums, BatchNorm statistics, the softmax / loss) stays fp32, so the
remaining differences are summation order and rare one-ulp mask flips (logits
<= 5e-3, gradients <= 1e-2 relative).
"""
import torch
import torch.nn.functional as F

from .unetpp import decoder_schedule


def _r(t):
    return t.to(torch.bfloat16).to(torch.float32)


class _RoundBoth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _r(x)

    @staticmethod
    def backward(ctx, g):
        return _r(g)


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _r(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        return _r(g)


rb, rf, rbw = _RoundBoth.apply, _RoundFwd.apply, _RoundBwd.apply


def _conv(conv, x):
    """x is already bf16-valued; its gradient (this consumer's dgrad output) is rounded; the conv
    output z is stored as bf16 and receives a bf16 dz."""
    x = rbw(x)
    z = F.conv2d(x, rf(conv.weight), None, conv.stride, conv.padding)
    return rb(z)


def _bn(bn, z):
    return F.batch_norm(z, bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.training,
                        bn.momentum, bn.eps)


def _conv_bn_relu(conv, bn, x, res=None, relu=True, trace=None, name=None):
    z = _conv(conv, x)
    y = _bn(bn, z)
    if res is not None:
        y = y + res
    if relu:
        y = torch.relu(y)
    y = rb(y)
    if trace is not None:
        trace[name] = (z.detach(), y.detach())
    return y


def _basic_block(blk, x, trace=None, base=None):
    t = _conv_bn_relu(blk.conv1, blk.bn1, x, trace=trace, name=base and base + "t1")
    idn = x
    if blk.downsample is not None:
        idn = _conv_bn_relu(blk.downsample[0], blk.downsample[1], x, relu=False, trace=trace, name=base and base + "idn")
    return _conv_bn_relu(blk.conv2, blk.bn2, t, res=idn, trace=trace, name=base and base + "out")


def unetpp_forward(model, x, trace=None):
    """model: oracle.unetpp.UnetPlusPlus.  Returns fp32 logits with bf16 rounding points.  trace (a dict)
    receives {engine activation name: (conv output z, activation)} for layer-by-layer comparisons."""
    enc, dec = model.encoder, model.decoder
    x = rf(x)
    f_stem = _conv_bn_relu(enc.conv1, enc.bn1, x, trace=trace, name="f_stem")
    # the pooled tensor is stored in bf16 (exact: max of bf16 values); its input gradient is rounded
    t = F.max_pool2d(rbw(f_stem), 3, 2, 1)
    feats = [f_stem]
    for li, layer in enumerate((enc.layer1, enc.layer2, enc.layer3, enc.layer4), start=1):
        for bi, blk in enumerate(layer):
            t = _basic_block(blk, t, trace, "encoder.layer%d.%d." % (li, bi))
        feats.append(t)
    feats = feats[::-1]
    dense = {}
    get = lambda s: feats[int(s[1:])] if s[0] == "f" else dense[s]
    for name, xsrc, skips in decoder_schedule(dec.depth):
        blk = dec.blocks[name]
        up = F.interpolate(get(xsrc), scale_factor=2, mode="nearest")
        cat = torch.cat([up] + [get(s) for s in skips], 1) if skips else up
        mid = _conv_bn_relu(blk.conv1[0], blk.conv1[1], cat, trace=trace, name=name + ".mid")
        dense[name] = _conv_bn_relu(blk.conv2[0], blk.conv2[1], mid, trace=trace, name=name)
    head = model.segmentation_head[0]
    xin = rbw(dense["x_0_%d" % dec.depth])
    logits = F.conv2d(xin, rf(head.weight), head.bias, 1, head.padding)
    return rbw(logits)
