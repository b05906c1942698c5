"""GPU tests (B200 box) of the device FuseLoss (forward + gradients w.r.t. the network outputs)
against ``oracle/loss_oracle.py`` -- the CPU restatement pinned on the unmodified ``libs/loss.py:FuseLoss``
(tests/test_oracle.py::test_fuse_loss_restatement_matches_reference_when_present).

The comparison isolates the loss: both sides consume the SAME predictions (the CUDA forward's, copied to the CPU).  Floats are
held to the protocol's bar (oracle/protocol.py:_floor_cmp): 1e-4 of the tensor's scale against the float32 run, or twice the
float32 rounding floor measured with a float64 run of the same restatement; integer counters are exact."""
import numpy as np
import pytest
import torch

from oracle.protocol import _floor_cmp

pytestmark = pytest.mark.gpu


def _to(d, fn):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            out[k] = fn(v)
        elif isinstance(v, list):
            out[k] = [fn(x) if isinstance(x, torch.Tensor) else x for x in v]
        elif isinstance(v, dict):
            out[k] = {kk: {n: (fn(x) if isinstance(x, torch.Tensor) else x) for n, x in dict(vv).items()} for kk, vv in v.items()}
        else:
            out[k] = v
    return out


def _case(fixture_weights, B=1, mode="val"):
    from oracle import oracle
    from pcaccumulation_b200 import config, synth
    from pcaccumulation_b200.motionnet import MotionNet

    cfg = config.workload_config("C1", mode=mode)
    sd = fixture_weights(cfg)
    vg = cfg["voxel_generator"]
    scenes = []
    for i in range(B):
        s = dict(synth.make_workload_scene("C1", 5 + i))
        p4 = np.concatenate((s["input_points"], s["time_indice"]), 1).astype(np.float32)
        s.update(oracle.voxelize(p4, vg["voxel_size"], vg["range"], vg["n_sweeps"]))
        scenes.append(s)
    inp = synth.collate(scenes)
    model = MotionNet(cfg).cuda().eval()
    model.load_state_dict(sd)
    return cfg, inp, model


def _cpu64(t):
    t = t.detach().cpu()
    return t.double() if t.is_floating_point() else t


@pytest.mark.parametrize("B", [1, 2])
def test_fuse_loss_forward_matches_oracle(fixture_weights, B):
    from oracle import loss_oracle
    from pcaccumulation_b200.loss import FuseLoss

    cfg, inp, model = _case(fixture_weights, B)
    inp_c = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in inp.items()}
    torch.manual_seed(3)
    pred = model(inp_c)
    stats = FuseLoss(loss_oracle.DEFAULT_WEIGHTS)(pred, inp_c)
    pred32 = _to(pred, lambda t: t.detach().cpu())
    pred64 = _to(pred, _cpu64)
    inp64 = {k: (_cpu64(v) if isinstance(v, torch.Tensor) else [m.double() for m in v]) for k, v in inp.items()}
    want32 = loss_oracle.fuse_loss(pred32, inp)
    want64 = loss_oracle.fuse_loss(pred64, inp64)
    rec = {}
    assert set(want32.keys()) - {"offset_gt"} == set(stats.keys()), (sorted(want32.keys()), sorted(stats.keys()))
    for k, v in want32.items():
        if k.endswith("_metric"):
            for name in v:
                assert np.array_equal(v[name], stats[k][name]), (k, name, v[name], stats[k][name])
        elif k == "offset_gt":
            _floor_cmp(k, pred["offset_gt"], v, want64[k], rec)
        else:
            _floor_cmp(k, torch.as_tensor(float(stats[k])).reshape(1), torch.as_tensor(float(v)).reshape(1),
                       torch.as_tensor(float(want64[k])).reshape(1), rec)
    for k in ("fb_loss", "mos_loss", "offset_loss", "obj_loss", "perm_loss"):
        assert float(want32[k]) > 0, k
    assert isinstance(stats["loss"], torch.Tensor) and stats["loss"].is_cuda and isinstance(stats["offset_l2_error"], float)


def test_fuse_loss_gradients_match_autograd_of_the_oracle(fixture_weights):
    """d loss / d {fb_seg_est, mos_est, offset_est, perm_matrix}: the analytic device gradients against autograd through the
    restatement (float32 = the reference's arithmetic; float64 = yardstick for the Jaccard differences, which the float32
    reference evaluates with ~1e-2 relative rounding noise per element)."""
    from oracle import loss_oracle
    from pcaccumulation_b200.loss import FuseLoss

    cfg, inp, model = _case(fixture_weights, 1)
    inp_c = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in inp.items()}
    torch.manual_seed(3)
    pred = model(inp_c)
    keys = ("fb_seg_est", "mos_est", "offset_est")

    def run(fn, p, i, cast):
        p = dict(p)
        leaves = {k: cast(p[k]).clone().requires_grad_(True) for k in keys}
        perm = [cast(q).clone().requires_grad_(True) for q in p["perm_matrix"]]
        p.update(leaves)
        p["perm_matrix"] = perm
        fn(p, i)["loss"].backward()
        return {k: v.grad for k, v in leaves.items()}, [q.grad for q in perm]

    g_ours, gp_ours = run(FuseLoss(loss_oracle.DEFAULT_WEIGHTS), pred, inp_c, lambda t: t.detach())
    pred32 = _to(pred, lambda t: t.detach().cpu())
    g32, gp32 = run(loss_oracle.fuse_loss, pred32, inp, lambda t: t)
    pred64 = _to(pred, _cpu64)
    inp64 = {k: (_cpu64(v) if isinstance(v, torch.Tensor) else [m.double() for m in v]) for k, v in inp.items()}
    g64, gp64 = run(loss_oracle.fuse_loss, pred64, inp64, lambda t: t)
    rec = {}
    for k in keys:
        assert float(g32[k].abs().max()) > 0
        _floor_cmp("grad_" + k, g_ours[k], g32[k], g64[k], rec)
    for a, b, c in zip(gp_ours, gp32, gp64):
        _floor_cmp("grad_perm", a, b, c, rec)


def test_validation_step_has_the_trainer_contract(fixture_weights):
    """libs/trainer.py:165-196: inference_one_batch(input_dict, 'val') -> stats with python floats for every *loss* key; the
    'train' phase and a differentiated forward fail loudly (no backward through the CUDA stages)."""
    from oracle import loss_oracle
    from pcaccumulation_b200.loss import FuseLoss
    from pcaccumulation_b200.trainer import inference_one_batch

    cfg, inp, model = _case(fixture_weights, 1)
    loss = FuseLoss(loss_oracle.DEFAULT_WEIGHTS)
    torch.manual_seed(3)
    stats = inference_one_batch(model, loss, dict(inp), "val")
    for k, v in stats.items():
        if "loss" in k:
            assert isinstance(v, float) and np.isfinite(v), k
    assert stats["loss"] > 0 and set(stats["mos_metric"]) == {"intersection", "union", "pred_positives", "gt_positives"}
    with pytest.raises(NotImplementedError):
        inference_one_batch(model, loss, dict(inp), "train")
    model.train()
    with pytest.raises(NotImplementedError):
        model({k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in inp.items()})
