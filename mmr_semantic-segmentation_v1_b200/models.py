"""Drop-in model classes for the reference's model-construction seam.

    smp.UnetPlusPlus(encoder_name=..., encoder_weights=..., in_channels=3, classes=C)
        SU/ModelTraining.py:247-254, SU/ModelEval.py:331-337
    smp.create_model(arch="UnetPlusPlus", encoder_name=..., ...)
        ED/Main_MMR_SegModel.py:589 (config['model'], ED/common_utils.py:235-241)

The modules below are parameter containers with smp's module tree, so `state_dict()` keys and
shapes are byte-compatible with the reference's checkpoints; `forward` never calls a torch.nn
layer: it replays the static plan of `engine.Engine` (hand-written sm_100a kernels through the
C-ABI of include/mmrseg.h).  There is no CPU or eager-PyTorch fallback: calling the model with a
CPU tensor raises.
"""
import glob
import math
import os
import warnings

import torch
import torch.nn as nn

from . import _lib, graph, ops
from .engine import Engine


# ------------------------------------------------------------------ parameter containers
def _conv(cin, cout, k, stride=1, bias=False):
    return nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, bias=bias)


class _BasicBlock(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = _conv(cin, cout, 3, stride)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = _conv(cout, cout, 3)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(_conv(cin, cout, 1, stride), nn.BatchNorm2d(cout))


class _ResNetEncoder(nn.Module):
    """torchvision-layout ResNet (BasicBlock) without avgpool/fc, as smp's ResNetEncoder."""

    def __init__(self, name):
        super().__init__()
        layers = graph.RESNET_LAYERS[name]
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        cin = 64
        for li, (n, cout) in enumerate(zip(layers, (64, 128, 256, 512)), start=1):
            blocks = []
            for bi in range(n):
                blocks.append(_BasicBlock(cin, cout, 2 if (bi == 0 and li > 1) else 1))
                cin = cout
            setattr(self, "layer%d" % li, nn.Sequential(*blocks))
        for m in self.modules():  # torchvision's default initialisation
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


class _Conv2dReLU(nn.Sequential):
    def __init__(self, cin, cout):
        super().__init__(_conv(cin, cout, 3), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _DecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = _Conv2dReLU(cin + cskip, cout)
        self.attention1 = nn.Identity()
        self.conv2 = _Conv2dReLU(cout, cout)
        self.attention2 = nn.Identity()


class _Decoder(nn.Module):
    def __init__(self, encoder_channels):
        super().__init__()
        self.blocks = nn.ModuleDict({name: _DecoderBlock(i, s, o) for name, (i, s, o) in
                                     graph.decoder_block_specs(encoder_channels).items()})
        for m in self.modules():  # smp initialize_decoder
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


def _load_pretrained_resnet(module, name, source):
    """`encoder_weights="imagenet"` (SU/ModelTraining.py:248-253) / `models.resnet18(pretrained=True)`
    (SU/UArchModel/resnet_unet.py:139-145) without a network: take the torchvision checkpoint from the local
    torch-hub cache (or $MMRSEG_PRETRAINED_DIR) when it is there, otherwise keep the random initialisation and
    say so -- the stock constructor call must not raise on a box without egress."""
    dirs = [os.environ.get("MMRSEG_PRETRAINED_DIR"), os.path.join(torch.hub.get_dir(), "checkpoints")]
    for d in [d for d in dirs if d]:
        files = sorted(glob.glob(os.path.join(d, name + "-*.pth")) + glob.glob(os.path.join(d, name + ".pth")))
        for f in files:
            sd = torch.load(f, map_location="cpu", weights_only=True)
            own = module.state_dict()
            sd = {k: v for k, v in sd.items() if k in own}      # smp deletes fc; torchvision keeps it
            missing = [k for k in own if k not in sd and not k.endswith("num_batches_tracked")]
            if missing:
                continue
            module.load_state_dict(sd, strict=False)
            return f
    warnings.warn("%s: no %s checkpoint in the local torch-hub cache (%s) and no network: the encoder keeps its "
                  "random initialisation; load a checkpoint with load_state_dict" % (source, name, dirs[-1]))
    return None


# ------------------------------------------------------------------ autograd bridge
# `seg = model(img)` / `loss.backward()` go through the custom ops mmrseg::plan_forward / mmrseg::plan_backward
# (mmrseg_b200/ops.py): plan_forward replays the static plan and returns the stacked logits in fresh memory, its
# registered autograd formula replays the backward plan, which leaves the parameter gradients in the model's flat
# gradient buffer (exposed as each parameter's .grad).


class _PlanModel(nn.Module):
    """Shared machinery: flat fp32 parameter / gradient buffers, plan cache, autograd bridge."""

    _size_multiple = 32   # five stride-2 stages

    def _graph(self):
        raise NotImplementedError

    def __init__(self):
        super().__init__()
        self._engines = {}
        self._flat = None
        self._on_grads_ready = None  # DDP installs callbacks here
        self._after_backward = None
        self._grad_cuts = None       # DDP: hook positions where a gradient bucket completes
        self._input_norm = None      # (mean, std) applied on the device to uint8 frames
        self._io_dtype = None        # .half() / .bfloat16(): dtype of the returned logits (masters stay fp32)
        self._io_channels_last = False
        self._handle = ops.register_model(self)     # what the custom ops know this model by

    def _n_classes(self):
        return getattr(self, "classes", None) or getattr(self, "n_class", None) or getattr(self, "n_classes")

    def _n_heads(self, training):
        return 4 if (training and getattr(self, "deep_supervision", False)) else 1

    # ---- dtype / memory-format requests of the reference's inference path -------------------
    # `self.model.to(memory_format=torch.channels_last); self.model.half()` (ED/Main_MMR_SegModel.py:1243-1244)
    # ask PyTorch for what this engine does on its own: NHWC activations and 16-bit tensor-core math.  The
    # fp32 master parameters are therefore left alone (converting them would only lose the checkpoint's
    # precision); the request is honoured at the boundary: inputs of any float dtype / memory format are
    # accepted and the logits come back in the requested dtype (and memory format).
    def half(self):
        self._io_dtype = torch.float16
        return self

    def bfloat16(self):
        self._io_dtype = torch.bfloat16
        return self

    def float(self):
        self._io_dtype = None
        return super().float()

    def to(self, *args, **kwargs):
        device, dtype, non_blocking, memory_format = torch._C._nn._parse_to(*args, **kwargs)
        if dtype is not None:
            if not dtype.is_floating_point:
                raise TypeError("nn.Module.to only accepts floating point dtypes, but got desired dtype=%s" % dtype)
            self._io_dtype = None if dtype == torch.float32 else dtype
        if memory_format is not None:
            self._io_channels_last = memory_format == torch.channels_last
        if device is not None:
            return super().to(device=device, non_blocking=non_blocking)
        return self

    # ---- flat parameter storage ------------------------------------------------------------
    def _named_params(self):
        """(name, Parameter) pairs, cached: walking the module tree five times per step cost more host time than
        launching the step's two CUDA graphs.  Parameter objects survive .to() / .cuda() / load_state_dict (only
        their .data moves); assigning a new submodule or Parameter invalidates the cache through __setattr__."""
        cached = self.__dict__.get("_named_params_cache")
        if cached is None:
            cached = list(super().named_parameters())
            self.__dict__["_named_params_cache"] = cached
        return cached

    def named_parameters(self, prefix="", recurse=True, remove_duplicate=True):
        """The reference's loops (`for param in model.parameters(): param.grad = None`, SU/ModelTraining.py:610)
        hit the cached list as well."""
        if prefix == "" and recurse and remove_duplicate:
            return iter(self._named_params())
        return super().named_parameters(prefix, recurse, remove_duplicate)

    def __setattr__(self, name, value):
        if isinstance(value, (nn.Module, nn.Parameter)):
            self.__dict__["_named_params_cache"] = None
        super().__setattr__(name, value)

    def _flatten(self, device):
        """Re-home every parameter into one fp32 buffer (so Adam and the gradient all-reduce are
        single launches) and create the matching flat gradient buffer."""
        self.__dict__["_named_params_cache"] = None
        named = self._named_params()
        offs, total = [], 0
        for _, p in named:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        flat = torch.zeros(total, device=device, dtype=torch.float32)
        gflat = torch.zeros(total, device=device, dtype=torch.float32)
        views = {}
        for (name, p), off in zip(named, offs):
            v = flat[off:off + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
            views[name] = gflat[off:off + p.numel()].view(p.shape)
        self._flat, self._gflat, self._gviews = flat, gflat, views
        self._flat_names = [n for n, _ in named]
        self._flat_offsets = dict(zip(self._flat_names, offs))
        self._flat_ptrs = [p.data_ptr() for _, p in named]
        self._engines = {}

    def _ensure_flat(self, device):
        ptrs = [p.data_ptr() for _, p in self._named_params()]
        if self._flat is None or ptrs != self._flat_ptrs or self._flat.device != device:
            for _, p in self._named_params():
                if p.dtype != torch.float32:
                    raise _lib.MmrError("parameters must stay fp32 masters (the kernels compute in "
                                        "bf16 with fp32 accumulate on their own); got %s" % p.dtype)
            self._flatten(device)

    def flat_parameters(self):
        return self._flat, self._gflat

    def _grads_live(self):
        """True when every .grad is still our view (gradient accumulation step)."""
        used = getattr(self, "_plan_params", None)
        live = [p.grad is not None for n, p in self._named_params() if used is None or n in used]
        if not any(live):
            return False
        if all(live) and all(p.grad.data_ptr() == self._gviews[n].data_ptr()
                             for n, p in self._named_params() if p.grad is not None):
            return True
        raise _lib.MmrError("parameter .grad tensors were replaced; call zero_grad(set_to_none=True) "
                            "(or leave them untouched) between backward passes")

    def _publish_grads(self):
        """Parameters the plan never reads (e.g. the torchvision `fc` the reference's ResNetUNet
        keeps) stay without a gradient, as under autograd."""
        used = getattr(self, "_plan_params", None)
        if used is None:
            used = self._plan_params = set(graph.graph_param_names(self._graph()))
        for n, p in self._named_params():
            if p.grad is None and n in used:
                p.grad = self._gviews[n]

    # ---- plan cache ------------------------------------------------------------------------
    def _engine_for(self, x, training):
        if not x.is_cuda:
            raise _lib.MmrError("mmrseg_b200 models run on a B200 only: input is on %s and there is "
                                "no CPU fallback" % x.device)
        if x.dtype == torch.uint8:       # frames as the loaders hold them: [N, H, W, 3]
            if x.dim() != 4 or x.shape[3] != 3:
                raise ValueError("expected uint8 frames of shape [N, H, W, 3], got %s" % (tuple(x.shape),))
            n, h, w, _ = x.shape
        else:
            if x.dim() != 4 or x.shape[1] != 3:
                raise ValueError("expected input of shape [N, 3, H, W], got %s" % (tuple(x.shape),))
            n, _, h, w = x.shape
        m = self._size_multiple
        if h % m or w % m:
            raise RuntimeError("Wrong input shape height=%d, width=%d. Expected image height and width "
                               "divisible by %d." % (h, w, m))
        self._ensure_flat(x.device)
        key = (n, h, w, bool(training))
        eng = self._engines.get(key)
        if eng is None:
            tensors = dict(self.named_parameters())
            tensors.update(dict(self.named_buffers()))
            params = {k: v.data for k, v in tensors.items()}
            eng = Engine(self._graph(), params, self._gviews, n, h, w, x.device, training=training)
            if self._input_norm is not None:
                eng.set_input_norm(*self._input_norm)
            self._engines[key] = eng
        return eng

    def set_input_normalization(self, mean, std):
        """Constants of `utils.normalize(batch, mean, std)` (SU/utils.py:480-519; SU/ModelTraining.py:300-301,
        579) applied on the device to uint8 HWC frames: `model(frames_u8)` then equals
        `model(normalize(ToTensor(frames)))` without the CPU loop and the fp32 host-to-device copy."""
        self._input_norm = (tuple(float(v) for v in mean), tuple(float(v) for v in std))
        for eng in self._engines.values():
            eng.set_input_norm(*self._input_norm)

    def _finish(self, t):
        if self._io_dtype is not None:
            t = t.to(self._io_dtype)
        if self._io_channels_last:
            t = t.contiguous(memory_format=torch.channels_last)
        return t

    def forward(self, x):
        if x.dtype != torch.uint8 and x.dtype != torch.float32:
            x = x.float()
        if not x.is_cuda:
            raise _lib.MmrError("mmrseg_b200 models run on a B200 only: input is on %s and there is "
                                "no CPU fallback" % x.device)
        self._ensure_flat(x.device)
        params = [p for _, p in self._named_params()] if (self.training and torch.is_grad_enabled()) else []
        out = torch.ops.mmrseg.plan_forward(x.contiguous(), self._handle, self.training, params)
        if out.shape[0] > 1:        # deep supervision (training): [main, aux...] at full resolution
            return [self._finish(t) for t in out.unbind(0)]
        # a view, not out[0]: the backward of a select zero-fills a [1, N, C, H, W] tensor and copies the gradient into
        # it (two more passes over the logits: 0.67 GB at BASELINE config 3) before plan_backward sees it
        return self._finish(out.view(out.shape[1:]))


    @torch.no_grad()
    def segment(self, x, labels=None):
        """The eval step of the reference in one replay -- `seg = model(img); evaluator.addBatch(seg, oneHotGT,
        args); seg = torch.argmax(seg, 1)` (SU/ModelTraining.py:736-760, SU/ModelEval.py:363-458) -- with the
        argmax and the confusion matrix taken in the segmentation head's epilogue: the logits never reach HBM
        (at BASELINE config 5 that is 3.4 GB per batch written and read back).
        x as for forward(); labels: int64 [N,H,W] class indices (or None).  Returns (pred uint8 [N,H,W] --
        torch.argmax's first-maximum rule, bit-identical to argmax of forward()'s logits --, confusion int64
        [N,C,C] with rows = label, columns = prediction, or None without labels); both are fresh tensors.
        Architectures whose head is not on the halo kernel (1x1 heads) take forward() + the metric kernel."""
        if self.training:
            raise _lib.MmrError("segment() is the eval-mode step: call model.eval() first")
        if x.dtype != torch.uint8 and x.dtype != torch.float32:
            x = x.float()
        eng = self._engine_for(x, training=False)
        if getattr(eng, "metric_swap", None) is None:
            from .metrics import confusion_matrix
            logits = eng.forward(x.contiguous())
            if labels is None:
                cm, pred = confusion_matrix(logits, torch.zeros(logits.shape[:1] + logits.shape[2:], dtype=torch.int64,
                                                                device=logits.device), return_pred=True)
                return pred.to(torch.uint8), None
            cm, pred = confusion_matrix(logits, labels, return_pred=True)
            return pred.to(torch.uint8), cm
        if labels is not None:
            if labels.shape != eng.metric_labels.shape:
                raise ValueError("labels must be [N, H, W] = %s, got %s" % (tuple(eng.metric_labels.shape), tuple(labels.shape)))
            eng.metric_labels.copy_(labels, non_blocking=True)
        else:
            eng.metric_labels.fill_(-1)      # out of range: nothing is counted
        eng.forward(x.contiguous(), metric=True)
        return eng.metric_pred.clone(), (eng.metric_cm.clone() if labels is not None else None)


class UnetPlusPlus(_PlanModel):
    """`smp.UnetPlusPlus` with the reference's arguments.  `encoder_weights="imagenet"` loads the torchvision
    checkpoint from the local torch-hub cache when present and otherwise warns and keeps the random
    initialisation (no network here; BASELINE configs are random-init); checkpoints load through
    `load_state_dict`."""

    def __init__(self, encoder_name="resnet18", encoder_depth=5, encoder_weights=None,
                 decoder_use_batchnorm=True, decoder_channels=(256, 128, 64, 32, 16),
                 decoder_attention_type=None, in_channels=3, classes=1, activation=None,
                 aux_params=None, deep_supervision=False):
        super().__init__()
        if encoder_name not in graph.RESNET_LAYERS:
            raise KeyError("encoder %r is not built; available: %s" % (encoder_name, list(graph.RESNET_LAYERS)))
        if (encoder_depth != 5 or tuple(decoder_channels) != graph.DECODER_CHANNELS or in_channels != 3
                or decoder_attention_type is not None or not decoder_use_batchnorm
                or activation is not None or aux_params is not None):
            raise NotImplementedError("only the configuration the reference uses is built: depth 5, "
                                      "decoder (256,128,64,32,16) with batch-norm, 3 input channels")
        if encoder_weights not in (None, "imagenet"):
            raise KeyError("Wrong pretrained weights `%s` for encoder `%s`. Available options are: ['imagenet']"
                           % (encoder_weights, encoder_name))
        self.encoder_name, self.classes, self.deep_supervision = encoder_name, classes, deep_supervision
        self._init_kwargs = {"encoder_name": encoder_name, "classes": classes, "deep_supervision": bool(deep_supervision)}
        self.encoder = _ResNetEncoder(encoder_name)
        if encoder_weights == "imagenet":
            _load_pretrained_resnet(self.encoder, encoder_name, "UnetPlusPlus(encoder_weights='imagenet')")
        self.decoder = _Decoder((3, 64, 64, 128, 256, 512))
        head = _conv(16, classes, 3, bias=True)
        nn.init.xavier_uniform_(head.weight)
        nn.init.constant_(head.bias, 0)
        self.segmentation_head = nn.Sequential(head, nn.Identity(), nn.Identity())
        if deep_supervision:
            self.ds_heads = nn.ModuleDict({"x_0_1": _conv(128, classes, 3, bias=True),
                                           "x_0_2": _conv(64, classes, 3, bias=True),
                                           "x_0_3": _conv(32, classes, 3, bias=True)})
            for m in self.ds_heads.values():
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)

    def _graph(self):
        return graph.unetpp_graph(self.encoder_name, self.classes, self.deep_supervision)


class _UnetDecoder(nn.Module):
    """smp decoders/unet/decoder.py UnetDecoder for a ResNet encoder: `center` identity, five blocks in a list."""

    def __init__(self):
        super().__init__()
        self.center = nn.Identity()
        specs = [(512, 256, 256), (256, 128, 128), (128, 64, 64), (64, 64, 32), (32, 0, 16)]
        self.blocks = nn.ModuleList([_DecoderBlock(i, s, o) for i, s, o in specs])
        for m in self.modules():  # smp initialize_decoder
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, mode="fan_in", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


class Unet(_PlanModel):
    """`smp.Unet(encoder_name="resnet18", encoder_weights="imagenet", in_channels=3, classes=C)` -- the reference's
    `--model smp_unet18` (SU/ModelTraining.py:255-262): same constructor arguments and `state_dict` keys as smp's;
    the decoder is a chain of the DecoderBlocks U-Net++ uses, on the same kernels (nearest x2 + concat in the
    consumer's loader)."""

    def __init__(self, encoder_name="resnet18", encoder_depth=5, encoder_weights=None,
                 decoder_use_batchnorm=True, decoder_channels=(256, 128, 64, 32, 16),
                 decoder_attention_type=None, in_channels=3, classes=1, activation=None, aux_params=None):
        super().__init__()
        if encoder_name not in graph.RESNET_LAYERS:
            raise KeyError("encoder %r is not built; available: %s" % (encoder_name, list(graph.RESNET_LAYERS)))
        if (encoder_depth != 5 or tuple(decoder_channels) != graph.DECODER_CHANNELS or in_channels != 3
                or decoder_attention_type is not None or not decoder_use_batchnorm
                or activation is not None or aux_params is not None):
            raise NotImplementedError("only the configuration the reference uses is built: depth 5, "
                                      "decoder (256,128,64,32,16) with batch-norm, 3 input channels")
        if encoder_weights not in (None, "imagenet"):
            raise KeyError("Wrong pretrained weights `%s` for encoder `%s`. Available options are: ['imagenet']"
                           % (encoder_weights, encoder_name))
        self.encoder_name, self.classes = encoder_name, classes
        self._init_kwargs = {"encoder_name": encoder_name, "classes": classes}
        self.encoder = _ResNetEncoder(encoder_name)
        if encoder_weights == "imagenet":
            _load_pretrained_resnet(self.encoder, encoder_name, "Unet(encoder_weights='imagenet')")
        self.decoder = _UnetDecoder()
        head = _conv(16, classes, 3, bias=True)
        nn.init.xavier_uniform_(head.weight)
        nn.init.constant_(head.bias, 0)
        self.segmentation_head = nn.Sequential(head, nn.Identity(), nn.Identity())

    def _graph(self):
        return graph.smp_unet_graph(self.encoder_name, self.classes)


def _convrelu(cin, cout, kernel, padding):  # SU/UArchModel/resnet_unet.py:36-44
    return nn.Sequential(nn.Conv2d(cin, cout, kernel, padding=padding), nn.ReLU(inplace=True))


class ResNetUNet(_PlanModel):
    """The reference's in-tree `ResNetUNet(n_class, resnet_model)` (SU/UArchModel/resnet_unet.py:134-300;
    constructed at SU/ModelTraining.py:244-246): same module tree, so `state_dict()` keys (including the
    `layerN.*` aliases of `base_model.*`), `.base_model` (differential learning rate) and `conv_last`
    (skipped on resume) are the reference's.  The torchvision backbone is a parameter container only;
    `forward` replays the static plan.  `pretrained=True` of the reference: the checkpoint is taken from the
    local torch-hub cache when present, else the encoder stays random-init with a warning (BASELINE config 3
    is random-init); checkpoints load through `load_state_dict`."""

    def __init__(self, n_class, resnet_model, pretrained=True):
        super().__init__()
        import torchvision
        self.n_class, self.resnet_model = n_class, resnet_model
        self._init_kwargs = {"n_class": n_class, "resnet_model": resnet_model, "pretrained": False}
        if resnet_model == 18:
            self.base_model = torchvision.models.resnet18(weights=None)
        elif resnet_model == 34:
            self.base_model = torchvision.models.resnet34(weights=None)
        else:
            raise ValueError("Only ResNet-18 and ResNet-34 are supported")
        if pretrained:   # the reference's `models.resnet18(pretrained=True)`: local cache or random init + warning
            _load_pretrained_resnet(self.base_model, "resnet%d" % resnet_model, "ResNetUNet(pretrained=True)")
        self.base_layers = list(self.base_model.children())
        self.layer0 = nn.Sequential(*self.base_layers[:3])
        self.layer0_1x1 = _convrelu(64, 64, 1, 0)
        self.layer1 = nn.Sequential(*self.base_layers[3:5])
        self.layer1_1x1 = _convrelu(64, 64, 1, 0)
        self.layer2 = self.base_layers[5]
        self.layer2_1x1 = _convrelu(128, 128, 1, 0)
        self.layer3 = self.base_layers[6]
        self.layer3_1x1 = _convrelu(256, 256, 1, 0)
        self.layer4 = self.base_layers[7]
        self.layer4_1x1 = _convrelu(512, 512, 1, 0)
        self.upsample = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv_up3 = _convrelu(256 + 512, 512, 3, 1)
        self.conv_up2 = _convrelu(128 + 512, 256, 3, 1)
        self.conv_up1 = _convrelu(64 + 256, 256, 3, 1)
        self.conv_up0 = _convrelu(64 + 256, 128, 3, 1)
        self.conv_original_size0 = _convrelu(3, 64, 3, 1)
        self.conv_original_size1 = _convrelu(64, 64, 3, 1)
        self.conv_original_size2 = _convrelu(64 + 128, 64, 3, 1)
        self.conv_last = nn.Conv2d(64, n_class, 1)

    def _graph(self):
        return graph.resnet_unet_graph(self.resnet_model, self.n_class)


class _DoubleConv(nn.Module):      # SU/UArchModel/unet_parts.py DoubleConv: indices 0,1,3,4 carry parameters
    def __init__(self, cin, cout, mid=None):
        super().__init__()
        mid = mid or cout
        self.double_conv = nn.Sequential(nn.Conv2d(cin, mid, 3, padding=1), nn.BatchNorm2d(mid), nn.ReLU(inplace=True),
                                         nn.Conv2d(mid, cout, 3, padding=1), nn.BatchNorm2d(cout),
                                         nn.ReLU(inplace=True))


class _Down(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), _DoubleConv(cin, cout))


class _Up(nn.Module):      # SU/UArchModel/unet_parts.py:248-271
    def __init__(self, cin, cout, bilinear=True):
        super().__init__()
        if bilinear:
            self.up = nn.Upsample(scale_factor=2, mode="nearest")
            self.conv = _DoubleConv(cin, cout, cin // 2)
        else:
            self.up = nn.ConvTranspose2d(cin, cin // 2, kernel_size=2, stride=2)
            self.conv = _DoubleConv(cin, cout)


class _OutConv(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1)


class UNet(_PlanModel):
    """The reference's in-tree `UNet(n_channels, n_classes, bilinear)` (SU/UArchModel/unet.py:104-245;
    constructed as `UNet(n_channels=3, n_classes=num_classes, bilinear=True)` at SU/ModelTraining.py:242 and
    SU/ModelEval.py:328): same module tree and `state_dict()` keys, default PyTorch initialisation in the
    same construction order.  `bilinear=True` (what the scripts construct) has a nearest-x2 Up block that is
    folded into the consumer conv's loader; `bilinear=False` upsamples with ConvTranspose2d(in, in // 2, 2, 2)
    (unet_parts.py:269), whose forward is the data gradient of a 2x2 stride-2 conv on the tcgen05 kernels
    (engine._fwd_convt).  Only n_channels=3 (the reference's call sites) is built."""

    _size_multiple = 16   # four MaxPool2d(2) stages; F.pad to the skip size is then the identity

    def __init__(self, n_channels=3, n_classes=2, bilinear=False):
        super().__init__()
        if n_channels != 3:
            raise NotImplementedError("only n_channels=3 is built (the reference's call sites)")
        self.n_channels, self.n_classes, self.bilinear = n_channels, n_classes, bilinear
        self._init_kwargs = {"n_channels": n_channels, "n_classes": n_classes, "bilinear": bool(bilinear)}
        self.inc = _DoubleConv(n_channels, 64)
        self.down1 = _Down(64, 128)
        self.down2 = _Down(128, 256)
        self.down3 = _Down(256, 512)
        factor = 2 if bilinear else 1                      # unet.py:153-163
        self.down4 = _Down(512, 1024 // factor)
        self.up1 = _Up(1024, 512 // factor, bilinear)
        self.up2 = _Up(512, 256 // factor, bilinear)
        self.up3 = _Up(256, 128 // factor, bilinear)
        self.up4 = _Up(128, 64, bilinear)
        self.outc = _OutConv(64, n_classes)

    def _graph(self):
        return graph.unet_graph(self.n_classes, self.bilinear)


def create_model(arch="UnetPlusPlus", encoder_name="resnet18", encoder_weights=None, in_channels=3,
                 classes=1, **kwargs):
    """`smp.create_model(**config['model'])` (ED/Main_MMR_SegModel.py:589)."""
    archs = {"unetplusplus": UnetPlusPlus, "unet": Unet}
    if arch.lower() not in archs:
        raise KeyError("architecture %r is not built; available: ['UnetPlusPlus', 'Unet']" % arch)
    return archs[arch.lower()](encoder_name=encoder_name, encoder_weights=encoder_weights,
                               in_channels=in_channels, classes=classes, **kwargs)
