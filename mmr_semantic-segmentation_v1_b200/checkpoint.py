"""Native checkpoint format of the plan models, with a lossless round trip to the
reference's `state_dict` checkpoints.

The reference saves `torch.save({'epoch', 'state_dict', 'optimizer'}, path)` (SU/ModelTraining.py:477-485,
825-845) and `{'epoch', 'network_state_dict', 'optimizer_state_dict', ...}` (ED/Main_MMR_SegModel.py:892-1001): a
pickle of ~120 separate tensors, each moved device -> host on save and host -> device on load.  The models here
keep every parameter in ONE flat fp32 buffer (what makes Adam and the gradient all-reduce single launches), so
the native file is that buffer verbatim:

    "MMRSEGCK" | u32 version | u32 header bytes | JSON header | zero padding to 64 B | blobs (64-B aligned)

    header = {arch: {class, kwargs}, params: [{name, shape, offset (floats into the flat blob)}],
              buffers: [{name, shape, dtype, blob}], blobs: {name: {offset, nbytes, dtype}}, optimizer: {...}, extra}
    blobs  = params_flat (fp32, the model's flat buffer incl. its 16-byte alignment padding)
             one blob per BatchNorm buffer group (running_mean / running_var fp32, num_batches_tracked int64)
             adam_exp_avg / adam_exp_avg_sq (fp32, same layout as params_flat) + per-parameter step counts

Save = one device -> host copy of the flat buffer (+ the small buffers); load = one host -> device copy into the
model's flat buffer -- no per-tensor traffic, no pickle.  `to_state_dict(path)` rebuilds, bit for bit, the
`state_dict()` the reference's scripts (and smp) expect; `from_state_dict` goes the other way.

Inference variant (`weights="bf16"`): convolution weights stored as bf16 in OHWI ("NHWC") order -- the precision
and layout the tensor-core kernels consume -- at half the size.  Loading it writes fp32 masters whose bf16
rounding is the stored value itself, so logits are bit-identical to those of the model loaded from the
checkpoint's bf16-rounded weights (BatchNorm affine parameters, statistics and biases stay fp32).
"""
import json
import struct
from collections import OrderedDict

import numpy as np
import torch

MAGIC = b"MMRSEGCK"
VERSION = 1
_ALIGN = 64
_NP = {"float32": np.float32, "int64": np.int64, "bfloat16": np.uint16, "uint8": np.uint8}
_TORCH = {"float32": torch.float32, "int64": torch.int64, "bfloat16": torch.bfloat16, "uint8": torch.uint8}


def _arch_of(model):
    kw = getattr(model, "_init_kwargs", None)
    if kw is None:
        raise TypeError("%s is not a mmrseg_b200 plan model" % type(model).__name__)
    return {"class": type(model).__name__, "kwargs": kw}


def build_model(arch):
    """Instantiate the architecture a native checkpoint names (random init; `load` fills it)."""
    import warnings
    from . import models
    cls = getattr(models, arch["class"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")     # pretrained-cache warnings: the checkpoint overwrites the weights anyway
        return cls(**arch["kwargs"])


class _Writer:
    def __init__(self):
        self.blobs, self.payload, self.size = {}, [], 0

    def add(self, name, array, dtype):
        raw = np.ascontiguousarray(array).tobytes()
        pad = (-self.size) % _ALIGN
        self.payload.append(b"\0" * pad)
        self.size += pad
        self.blobs[name] = {"offset": self.size, "nbytes": len(raw), "dtype": dtype}
        self.payload.append(raw)
        self.size += len(raw)


def _cpu_np(t):
    t = t.detach()
    if t.dtype == torch.bfloat16:
        return t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)
    return t.contiguous().cpu().numpy()


def save(path, model, optimizer=None, extra=None, weights="fp32"):
    """Write a native checkpoint.  weights: "fp32" (training checkpoint, lossless) or "bf16" (inference checkpoint:
    conv weights as bf16 OHWI, no optimiser state).  extra: a JSON-able dict (epoch, metrics ...)."""
    if weights not in ("fp32", "bf16"):
        raise ValueError("weights must be 'fp32' or 'bf16'")
    named = list(model._named_params())
    dev = named[0][1].device
    model._ensure_flat(dev)
    w = _Writer()
    header = {"version": VERSION, "arch": _arch_of(model), "weights": weights, "extra": extra or {}}
    if weights == "fp32":
        w.add("params_flat", _cpu_np(model._flat), "float32")        # ONE device -> host copy
        header["params"] = [{"name": n, "shape": list(p.shape), "offset": model._flat_offsets[n]} for n, p in named]
    else:
        params = []
        for n, p in named:
            if p.dim() == 4:      # conv / transposed-conv weight: [O][I][kh][kw] fp32 -> [O][kh][kw][I] bf16
                arr = _cpu_np(p.detach().permute(0, 2, 3, 1).to(torch.bfloat16))
                w.add("w:" + n, arr, "bfloat16")
                params.append({"name": n, "shape": list(p.shape), "blob": "w:" + n, "layout": "OHWI"})
            else:
                w.add("p:" + n, _cpu_np(p), "float32")
                params.append({"name": n, "shape": list(p.shape), "blob": "p:" + n, "layout": "native"})
        header["params"] = params
    bufs = []
    for n, b in model.named_buffers():
        dt = str(b.dtype).replace("torch.", "")
        w.add("b:" + n, _cpu_np(b), dt)
        bufs.append({"name": n, "shape": list(b.shape), "dtype": dt, "blob": "b:" + n})
    header["buffers"] = bufs
    if optimizer is not None and weights == "fp32":
        header["optimizer"] = _save_optimizer(w, model, optimizer)
    header["blobs"] = w.blobs
    hj = json.dumps(header).encode()
    head = MAGIC + struct.pack("<II", VERSION, len(hj)) + hj
    head += b"\0" * ((-len(head)) % _ALIGN)
    with open(path, "wb") as fh:
        fh.write(head)
        for chunk in w.payload:
            fh.write(chunk)
    return path


def _save_optimizer(w, model, optimizer):
    groups = []
    index = {id(p): n for n, p in model._named_params()}
    for g in optimizer.param_groups:
        groups.append({"hyper": {k: v for k, v in g.items() if k != "params" and isinstance(v, (int, float, bool, str, list, tuple, type(None)))},
                       "params": [index[id(p)] for p in g["params"]]})
    out = {"class": type(optimizer).__name__, "groups": groups, "state": {}}
    slots = sorted({k for st in optimizer.state.values() for k, v in st.items() if torch.is_tensor(v) and v.dim() > 0})
    total = model._flat.numel()
    for slot in slots:      # moment tensors laid out like params_flat (zeros where a parameter has no state)
        flat = np.zeros((total,), dtype=np.float32)
        for n, p in model._named_params():
            st = optimizer.state.get(p, {})
            if slot in st:
                off = model._flat_offsets[n]
                flat[off:off + p.numel()] = _cpu_np(st[slot].float()).reshape(-1)
        w.add("o:" + slot, flat, "float32")
        out["state"][slot] = "o:" + slot
    out["steps"] = {n: float(optimizer.state[p]["step"]) for n, p in model._named_params()
                    if p in optimizer.state and "step" in optimizer.state[p]}
    return out


def _read(path):
    with open(path, "rb") as fh:
        if fh.read(8) != MAGIC:
            raise ValueError("%s is not a mmrseg_b200 native checkpoint" % path)
        version, hlen = struct.unpack("<II", fh.read(8))
        if version != VERSION:
            raise ValueError("native checkpoint version %d is not supported (this build reads %d)" % (version, VERSION))
        header = json.loads(fh.read(hlen).decode())
        base = 16 + hlen
        base += (-base) % _ALIGN
    data = np.memmap(path, dtype=np.uint8, mode="r", offset=base) if header["blobs"] else np.zeros((0,), np.uint8)
    return header, data


def _blob(header, data, name):
    b = header["blobs"][name]
    raw = np.frombuffer(data, dtype=np.uint8, count=b["nbytes"], offset=b["offset"])
    return raw.view(_NP[b["dtype"]])


def _param_tensor(header, data, entry, flat):
    """fp32 torch tensor of one parameter, from either checkpoint flavour."""
    shape = entry["shape"]
    numel = int(np.prod(shape)) if shape else 1
    if "offset" in entry:
        return torch.from_numpy(np.array(flat[entry["offset"]:entry["offset"] + numel])).view(shape)
    raw = _blob(header, data, entry["blob"])
    if entry["layout"] == "OHWI":
        o, i, kh, kw = shape
        t = torch.from_numpy(np.array(raw)).view(torch.bfloat16).view(o, kh, kw, i).float()
        return t.permute(0, 3, 1, 2).contiguous()
    return torch.from_numpy(np.array(raw)).view(shape)


def to_state_dict(path):
    """The checkpoint as the `state_dict()` of the reference's model (smp / in-tree key names, OIHW fp32 weights):
    for an fp32 checkpoint bit-identical to the state_dict the model had when it was saved."""
    header, data = _read(path)
    flat = _blob(header, data, "params_flat") if header["weights"] == "fp32" else None
    params = {e["name"]: _param_tensor(header, data, e, flat) for e in header["params"]}
    bufs = {}
    for e in header["buffers"]:
        bufs[e["name"]] = torch.from_numpy(np.array(_blob(header, data, e["blob"]))).view(e["shape"]).to(_TORCH[e["dtype"]])
    # nn.Module.state_dict order (per module: parameters, then buffers) and alias keys (the reference's ResNetUNet
    # registers the backbone's layers twice: base_model.* and layerN.*) come from the architecture itself
    model = build_model(header["arch"])
    canon = {id(p): n for n, p in model._named_params()}
    canon.update({id(b): n for n, b in model.named_buffers()})
    sd = OrderedDict()
    for k, v in model.state_dict(keep_vars=True).items():
        name = canon[id(v)]
        sd[k] = params[name] if name in params else bufs[name]
    return sd


def from_state_dict(path, state_dict, arch, extra=None, weights="fp32"):
    """Write a native checkpoint from a reference `state_dict` (e.g. a smp checkpoint from Drive).
    arch: {"class": "UnetPlusPlus" | "ResNetUNet" | "UNet", "kwargs": {...constructor arguments...}}."""
    model = build_model(arch)
    model.load_state_dict(state_dict, strict=True)
    return save(path, model, extra=extra, weights=weights)


def load(path, model=None, optimizer=None, device=None):
    """Load a native checkpoint.  model None: the architecture is built from the header.  With an fp32 checkpoint
    the parameters arrive as ONE host -> device copy into the model's flat buffer.  Returns (model, extra)."""
    header, data = _read(path)
    if model is None:
        model = build_model(header["arch"])
        if device is not None:
            model = model.to(device)
    elif _arch_of(model) != header["arch"]:
        raise ValueError("checkpoint holds %s, the model is %s" % (header["arch"], _arch_of(model)))
    named = list(model._named_params())
    dev = named[0][1].device
    model._ensure_flat(dev)
    with torch.no_grad():
        if header["weights"] == "fp32":
            flat = _blob(header, data, "params_flat")
            want = [(e["name"], e["offset"]) for e in header["params"]]
            have = [(n, model._flat_offsets[n]) for n, _ in named]
            if want == have and flat.size == model._flat.numel():
                model._flat.copy_(torch.from_numpy(np.array(flat)), non_blocking=False)
            else:   # another layout version: per tensor
                params = dict(named)
                for e in header["params"]:
                    params[e["name"]].copy_(_param_tensor(header, data, e, flat))
        else:
            params = dict(named)
            for e in header["params"]:
                params[e["name"]].copy_(_param_tensor(header, data, e, None))
        bufs = dict(model.named_buffers())
        for e in header["buffers"]:
            t = torch.from_numpy(np.array(_blob(header, data, e["blob"]))).view(e["shape"])
            bufs[e["name"]].copy_(t.to(bufs[e["name"]].dtype))
    if optimizer is not None and "optimizer" in header:
        _load_optimizer(header, data, model, optimizer)
    return model, header.get("extra", {})


def _load_optimizer(header, data, model, optimizer):
    """Rebuild torch-layout optimiser state (state[p] = {step, exp_avg, exp_avg_sq} / {momentum_buffer}) and hand
    it to optimizer.load_state_dict, which works for torch.optim.Adam as well as FusedAdam."""
    o = header["optimizer"]
    names = [n for n, _ in model._named_params()]
    params = dict(model._named_params())
    groups, index, state = [], {}, {}
    for g in o["groups"]:
        ids = []
        for n in g["params"]:
            index[n] = len(index)
            ids.append(index[n])
        groups.append(dict(g["hyper"], params=ids))
    for slot, blob in o["state"].items():
        flat = _blob(header, data, blob)
        for n in names:
            if n not in index:
                continue
            p = params[n]
            off = model._flat_offsets[n]
            if n in o["steps"] or slot == "momentum_buffer":
                state.setdefault(index[n], {})[slot] = torch.from_numpy(np.array(flat[off:off + p.numel()])).view(p.shape)
    for n, s in o["steps"].items():
        if n in index:
            state.setdefault(index[n], {})["step"] = torch.tensor(s, dtype=torch.float32)
    optimizer.load_state_dict({"state": state, "param_groups": groups})
