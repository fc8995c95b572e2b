"""Host logic of the optimiser seam on CPU (no kernel launches): which memory a param group's step covers,
and that optimiser state keeps torch's per-parameter layout across state_dict round trips (the reference resumes
with `optimizer.load_state_dict`, ED/Main_MMR_SegModel.py:991)."""
import torch

from mmrseg_b200.optim import FusedAdam, FusedSGD, _runs


def _flat_model(sizes=(10, 7, 64, 5)):
    """Parameters that are views of one flat buffer, each padded to 4 floats (models._PlanModel._flatten)."""
    offs, tot = [], 0
    for n in sizes:
        offs.append(tot)
        tot += (n + 3) // 4 * 4
    flat, gflat = torch.zeros(tot), torch.zeros(tot)
    params = []
    for n, o in zip(sizes, offs):
        p = torch.nn.Parameter(torch.empty(0))
        p.data = flat[o:o + n]
        p.grad = gflat[o:o + n]
        params.append(p)
    return flat, gflat, params, offs


def test_runs_cover_exactly_the_group():
    flat, gflat, params, offs = _flat_model()
    opt = FusedAdam(params)
    runs = opt._plan(0, params)
    assert len(runs) == 1                                   # the whole model: one launch
    idx, ptrs, numel, ok, spans = runs[0]
    assert spans[0].data_ptr() == flat.data_ptr() and spans[0].numel() == numel and len(spans) == 4
    assert sorted(idx) == [0, 1, 2, 3] and ok
    assert ptrs[0] == flat.data_ptr() and ptrs[1] == gflat.data_ptr()
    assert numel == offs[-1] + params[-1].numel()           # padding between members, none after the last
    # state tensors are views of one buffer with the parameters' layout
    m0 = opt.state[params[0]]["exp_avg"]
    assert all(opt.state[p]["exp_avg"].data_ptr() - m0.data_ptr() == p.data_ptr() - params[0].data_ptr()
               for p in params)


def test_subset_group_does_not_touch_other_parameters():
    flat, gflat, params, offs = _flat_model()
    opt = FusedAdam([params[1], params[2]])                 # e.g. FusedAdam(model.encoder.parameters())
    runs = opt._plan(0, [params[1], params[2]])
    assert len(runs) == 1
    _, ptrs, numel, _, _ = runs[0]
    lo = (ptrs[0] - flat.data_ptr()) // 4
    assert lo == offs[1] and lo + numel == offs[2] + params[2].numel()      # nothing of params[0] / params[3]
    # non-adjacent members: separate launches
    opt2 = FusedAdam([params[0], params[2]])
    assert len(opt2._plan(0, [params[0], params[2]])) == 2
    # two groups (differential learning rates): each parameter is covered by exactly one group's runs
    opt3 = FusedSGD([{"params": params[:2], "lr": 0.1}, {"params": params[2:], "lr": 0.01}], momentum=0.9)
    covered = []
    for gi, g in enumerate(opt3.param_groups):
        for idx, ptrs, numel, ok, spans in opt3._plan(gi, g["params"]):
            lo = (ptrs[0] - flat.data_ptr()) // 4
            covered.append((lo, lo + numel))
    assert sorted(covered) == [(0, offs[1] + 7), (offs[2], offs[3] + 5)]


def test_gradient_elsewhere_breaks_the_run():
    flat, gflat, params, offs = _flat_model()
    params[1].grad = torch.zeros(7)                         # a gradient that is not the flat view
    opt = FusedAdam(params)
    runs = opt._plan(0, params)
    assert len(runs) == 3 and sorted(len(r[0]) for r in runs) == [1, 1, 2]


def test_state_dict_round_trip_keeps_step_and_moments():
    flat, gflat, params, offs = _flat_model()
    opt = FusedAdam(params, lr=1e-3)
    opt._plan(0, params)
    for i, p in enumerate(params):                          # what three steps would have left
        opt.state[p]["step"] = torch.tensor(3.0)
        opt.state[p]["exp_avg"].fill_(i + 1.0)
        opt.state[p]["exp_avg_sq"].fill_(10.0 * (i + 1))
    sd = opt.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}          # torch.optim.Adam's keys
    # same optimiser
    opt.load_state_dict(sd)
    opt._plan(0, params)
    for i, p in enumerate(params):
        st = opt.state[p]
        assert float(st["step"]) == 3.0 and torch.all(st["exp_avg"] == i + 1.0) and torch.all(st["exp_avg_sq"] == 10.0 * (i + 1))
    # a fresh optimiser over a fresh model (resume), and a checkpoint written by torch.optim.Adam itself
    flat2, gflat2, params2, _ = _flat_model()
    ref = torch.optim.Adam([torch.nn.Parameter(torch.zeros(n)) for n in (10, 7, 64, 5)], lr=1e-3)
    for i, p in enumerate(ref.param_groups[0]["params"]):
        p.grad = torch.full_like(p, 0.5 * (i + 1))
    ref.step()
    for loaded in (sd, ref.state_dict()):
        opt2 = FusedAdam(params2, lr=1e-3)
        opt2.load_state_dict(loaded)
        runs = opt2._plan(0, params2)
        assert len(runs) == 1                                # re-homed into one flat buffer again
        for i, p in enumerate(params2):
            src = loaded["state"][i]
            st = opt2.state[p]
            assert float(st["step"]) == float(src["step"])
            assert torch.equal(st["exp_avg"], src["exp_avg"]) and torch.equal(st["exp_avg_sq"], src["exp_avg_sq"])
        assert "decoupled" in opt2.param_groups[0]           # torch's checkpoint has no such key: default restored
