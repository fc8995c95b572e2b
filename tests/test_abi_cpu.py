"""The C-ABI library loads without a GPU and exports every symbol include/mmrseg.h declares;
the host-side logic (graph construction, bucket planning, error paths) runs on CPU."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "mmrseg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mmr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from mmrseg_b200 import _lib
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, "ctypes signature missing for %s" % n
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.mmr_abi_version() == 1
    assert isinstance(lib.mmr_last_error(), bytes)


def test_no_gpu_means_loud_failure_not_fallback():
    from mmrseg_b200 import _lib
    from mmrseg_b200.models import UnetPlusPlus
    from mmrseg_b200.losses import dice_loss
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert _lib.device_ok() is False
    net = UnetPlusPlus("resnet18", classes=2)
    with pytest.raises(_lib.MmrError):
        net(torch.zeros(1, 3, 32, 32))
    with pytest.raises(_lib.MmrError):
        dice_loss(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mmr_semantic-segmentation_v1_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "/root/reference" not in text, f


def test_state_dict_is_smp_compatible():
    from oracle.unetpp import UnetPlusPlus as OracleNet
    from mmrseg_b200.models import UnetPlusPlus, create_model
    for enc, classes in (("resnet18", 2), ("resnet34", 10)):
        ref = OracleNet(enc, None, 3, classes).state_dict()
        net = create_model("UnetPlusPlus", encoder_name=enc, encoder_weights=None, in_channels=3, classes=classes)
        sd = net.state_dict()
        assert list(sd.keys()) == list(ref.keys())
        assert all(sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype for k in sd)
        net.load_state_dict(ref, strict=True)
    with pytest.raises(KeyError):
        create_model("DeepLabV3Plus")
    with pytest.raises(KeyError):
        UnetPlusPlus("tu-mobilenetv3_small_100")


def test_graph_matches_reference_dataflow():
    """concat order and channel sums of every decoder conv1."""
    from mmrseg_b200 import graph
    ops = graph.unetpp_graph("resnet18", 2)
    convs = {o["conv"]: o for o in ops if o["op"] in ("conv", "head")}
    assert len(convs) == 19 + 22 + 1  # 16 BasicBlock convs + 3 downsamples, 22 decoder convs, head (stem is its own op)
    enc = [o for o in ops if o["op"] == "conv" and o["conv"].startswith("encoder.")]
    assert len(enc) == 16 + 3
    want = {"x_0_0": ["encoder.layer4.1.out", "encoder.layer3.1.out"],
            "x_0_1": ["x_0_0", "x_1_1", "encoder.layer2.1.out"],
            "x_0_3": ["x_0_2", "x_1_3", "x_2_3", "x_3_3", "f_stem"],
            "x_0_4": ["x_0_3"]}
    for blk, srcs in want.items():
        op = convs["decoder.blocks.%s.conv1.0" % blk]
        assert [s for s, _ in op["src"]] == srcs
        assert [u for _, u in op["src"]] == [2] + [1] * (len(srcs) - 1)
    order = [o["conv"].split(".")[2] for o in ops if o["op"] == "conv" and o["conv"].endswith("conv1.0")]
    assert order == ["x_0_0", "x_1_1", "x_2_2", "x_3_3", "x_0_1", "x_1_2", "x_2_3", "x_0_2", "x_1_3", "x_0_3", "x_0_4"]


def test_arena_never_overlaps_live_buffers():
    import random
    from mmrseg_b200.engine import _Arena
    rnd = random.Random(0)
    ar = _Arena()
    reqs = []
    for i in range(300):
        t0 = rnd.randrange(0, 60)
        t1 = t0 + rnd.randrange(0, 12)
        size = rnd.randrange(1, 50) * 1024
        ar.request(i, size, t0, t1)
        reqs.append((i, size, t0, t1))
    off, peak = ar.plan()
    for i, s, a0, a1 in reqs:
        assert off[i] + s <= peak
        for j, s2, b0, b1 in reqs:
            if i < j and not (a1 < b0 or b1 < a0):            # lifetimes intersect
                assert off[i] + s <= off[j] or off[j] + s2 <= off[i], (i, j)
    assert peak < sum(s for _, s, _, _ in reqs)


def test_cached_parameter_list_matches_module_tree():
    """Plan models answer named_parameters() / parameters() from a cached list (the reference's per-step loops walk
    it): same names and order as nn.Module's traversal, shared modules de-duplicated, invalidated by assignment."""
    import torch
    from mmrseg_b200.models import ResNetUNet, UNet, UnetPlusPlus
    for m in (UnetPlusPlus("resnet18", classes=2), ResNetUNet(3, 18), UNet(3, 2, bilinear=True)):
        want = [n for n, _ in torch.nn.Module.named_parameters(m)]
        assert [n for n, _ in m.named_parameters()] == want
        assert [id(p) for p in m.parameters()] == [id(p) for _, p in torch.nn.Module.named_parameters(m)]
        assert [n for n, _ in m.named_parameters(prefix="net")] == ["net." + n for n in want]
        m.extra = torch.nn.Linear(2, 2)
        assert len(list(m.parameters())) == len(want) + 2
