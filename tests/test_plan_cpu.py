"""Host-side planners of the conv kernels (no GPU): which weight-gradient generation / halo loader a layer gets, and
that the Python side and the C-ABI agree on the size of the split-K partial buffers."""
import itertools

from mmrseg_b200 import _lib, convplan, graph


def test_wgrad_generation_by_layer_shape():
    # 64-channel chunks, 64 / 32 output channels: the filter column moves to the dz side (mode 1)
    c = convplan.wgrad_halo_config(256, 256, 16, 64, 4, 64, True)
    assert c["mode"] == 1 and c["bn"] == 64 and c["tx"] in (1, 2) and c["nchunks"] * c["n_ntiles"] * c["n_split"] <= 148
    assert convplan.wgrad_halo_config(256, 256, 16, 64, 5, 32, True)["mode"] == 1
    # narrow layers: one MMA per 16 pixels, cp.async gathers (mode 2)
    for cb, cout in ((16, 16), (32, 16), (32, 32)):
        c = convplan.wgrad_halo_config(512, 512, 16, cb, 1, cout, cb == 32)
        assert c["mode"] == 2 and c["tx"] in (2, 4), c
    # the space-to-depth stem (phased dz) and 16-channel dz slices of wide chunks stay on the second generation
    assert convplan.wgrad_halo_config(128, 128, 16, 64, 1, 256, False, dz_phased=True)["mode"] == 0
    assert convplan.wgrad_halo_config(64, 64, 2, 64, 1, 48, False)["mode"] == 0
    # forcing a generation
    assert convplan.wgrad_halo_config(64, 64, 2, 64, 1, 64, False, force=dict(mode=0))["mode"] == 0


def test_partial_buffer_sizes_match_the_library():
    lib = _lib.lib()
    for cb, cout, nchunks in itertools.product((16, 32, 64), (16, 32, 64, 128), (1, 3)):
        c = convplan.wgrad_halo_config(128, 128, 4, cb, nchunks, cout, False)
        need = c["nchunks"] * c["n_ntiles"] * c["n_split"] * c["per_cta"]
        if c["mode"] == 2:
            want = lib.mmr_wgrad_thin_partial_floats(c["nchunks"], cb, c["bn"], c["n_ntiles"], c["n_split"])
        elif c["mode"] == 1:
            want = lib.mmr_wgrad_kx_partial_floats(c["nchunks"], c["bn"], c["n_ntiles"], c["n_split"])
        else:
            want = lib.mmr_wgrad_halo_partial_floats(c["nchunks"], cb, c["bn"], c["n_ntiles"], c["n_split"])
        assert need == want, (cb, cout, nchunks, c)


def test_halo_loader_choice():
    # 16-channel sources at their own resolution: cp.async gather; 64-byte rows and nearest-x2 sources: TMA
    assert convplan.halo_config(512, 512, 16, 16, 1, 16, False)["loader"] == 1
    assert convplan.halo_config(512, 512, 16, 32, 1, 16, True)["loader"] == 0
    assert convplan.halo_config(256, 256, 16, 64, 4, 64, True)["loader"] == 0
    assert convplan.halo_config(512, 512, 16, 16, 1, 16, False, force=dict(loader=0))["loader"] == 0
    assert convplan.halo_config(512, 512, 16, 32, 1, 16, True, force=dict(loader=1))["loader"] == 1


def test_smp_unet_graph_dataflow():
    ops = graph.smp_unet_graph("resnet18", 3)
    convs = {o["conv"]: o for o in ops if o["op"] == "conv" and o["conv"].startswith("decoder")}
    assert [s for s in convs["decoder.blocks.0.conv1.0"]["src"]] == [("encoder.layer4.1.out", 2), ("encoder.layer3.1.out", 1)]
    assert convs["decoder.blocks.3.conv1.0"]["src"] == [("d2", 2), ("f_stem", 1)] and convs["decoder.blocks.3.conv1.0"]["cout"] == 32
    assert convs["decoder.blocks.4.conv1.0"]["src"] == [("d3", 2)] and convs["decoder.blocks.4.conv2.0"]["cout"] == 16
    assert ops[-1]["op"] == "head" and ops[-1]["src"] == [("d4", 1)] and ops[-1]["cout"] == 3
