"""GPU tests of the guard / failure paths of the forward (run on the B200 box): the branches the reference takes when a stage
has nothing to work on.  Each case drives the CUDA path and the CPU oracle into the same branch by injecting the same
FG/BG map (``MotionNet.inject`` / ``OracleMotionNet(inject=...)``) and asserts the same outcome -- results or exception.

  * STPN skipped when the foreground has <= MIN_POINTS points            (models/motionnet.py:217-226)
  * TubeNet skipped when <= MIN_POINTS points carry an instance           (models/motionnet.py:242-260)
  * Q4: STPN skipped but instances exist -> NameError                     (models/motionnet.py:222-245)
  * no background pillar in a frame -> IndexError                         (models/egomotion.py:155-169)
  * non-finite Kabsch covariance (all-zero geometric features, 0/0 in the L2 normalisation) -> torch.svd raises on the CPU
    and the reference falls back to R = I, t = 0                          (toolbox/register_utils.py:295-304)
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def cuda_dict(d):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def _setup(fixture_weights, mode="test", scene=5):
    from oracle import oracle
    from pcaccumulation_b200 import config, synth
    from pcaccumulation_b200.motionnet import MotionNet

    cfg = config.workload_config("C1", mode=mode)
    sd = fixture_weights(cfg)
    s = synth.make_workload_scene("C1", scene)
    p4 = np.concatenate((s["input_points"], s["time_indice"]), 1).astype(np.float32)
    vg = cfg["voxel_generator"]
    s = dict(s)
    s.update(oracle.voxelize(p4, vg["voxel_size"], vg["range"], vg["n_sweeps"]))
    inp = synth.collate([s])
    model = MotionNet(cfg).cuda().eval()
    model.load_state_dict(sd)
    return cfg, sd, inp, model


def _both(cfg, sd, inp, model, inject, seed=11):
    from oracle import oracle

    torch.manual_seed(seed)
    ref = oracle.OracleMotionNet(cfg, sd, inject=inject).forward(inp)
    model.inject = dict(inject)
    torch.manual_seed(seed)
    out = model(cuda_dict(inp))
    model.inject = {}
    return out, ref


def _grid(cfg, inp):
    T = cfg["voxel_generator"]["n_sweeps"]
    nx, ny = int(inp["shape"][0][0]), int(inp["shape"][0][1])
    return T, ny, nx


@pytest.mark.parametrize("n_fg_pillars", [0, 1])
def test_stpn_and_tubenet_skipped_when_foreground_is_tiny(fixture_weights, n_fg_pillars):
    cfg, sd, inp, model = _setup(fixture_weights)
    T, ny, nx = _grid(cfg, inp)
    fb_map = torch.zeros(1, T, 1, ny, nx, dtype=torch.int64)
    if n_fg_pillars:  # one foreground pillar holding at most MIN_POINTS points
        p2v = inp["point_to_voxel_map"][:, 0]
        counts = torch.bincount(p2v[p2v >= 0])
        pid = int(torch.nonzero((counts > 0) & (counts <= 15))[0])
        z, y, x, t = [int(v) for v in inp["coordinates"][pid][-4:]]
        fb_map[0, t, 0, y, x] = 1
    out, ref = _both(cfg, sd, inp, model, {"fb_est_map": fb_map})
    n_fg = int(ref["fb_est_per_points"].sum())
    assert n_fg <= 15 and (n_fg > 0) == bool(n_fg_pillars)
    assert torch.equal(out["fb_est_per_points"].cpu(), ref["fb_est_per_points"])
    # STPN skipped: the initial values survive (motionnet.py:217-219), nothing is dynamic, no instance, no TubeNet
    assert torch.equal(out["mos_est"].cpu(), ref["mos_est"]) and torch.equal(out["offset_est"].cpu(), ref["offset_est"])
    assert float(out["mos_est"][:, 0].min()) == 1.0 and float(out["mos_est"][:, 1].abs().max()) == 0.0
    assert torch.equal(out["inst_labels_est"].cpu(), ref["inst_labels_est"]) and int(ref["inst_labels_est"].abs().max()) == 0
    assert torch.equal(out["rec_est"], out["transformed_points"])
    for k in ("tpointnet_loss_terms", "inst_pose_est", "sub_rec_est", "inst_l2_error"):
        assert k not in ref and k not in out, k
    assert set(out.keys()) == set(ref.keys())


def test_q4_name_error_when_stpn_was_skipped_but_instances_exist(fixture_weights):
    """Upstream dies with NameError('mos_feats') when the foreground guard failed and the clustering still finds instances
    (only reachable with externally supplied motion logits); the drop-in raises the same exception type."""
    cfg, sd, inp, model = _setup(fixture_weights)
    T, ny, nx = _grid(cfg, inp)
    from oracle import oracle

    torch.manual_seed(11)
    free = oracle.OracleMotionNet(cfg, sd).forward(inp)
    assert int((free["inst_labels_est"] != 0).sum()) > 15
    inject = {"fb_est_map": torch.zeros(1, T, 1, ny, nx, dtype=torch.int64), "mos_est": free["mos_est"], "offset_est": free["offset_est"]}
    with pytest.raises(NameError):
        torch.manual_seed(11)
        oracle.OracleMotionNet(cfg, sd, inject=inject).forward(inp)
    model.inject = dict(inject)
    with pytest.raises(NameError):
        torch.manual_seed(11)
        model(cuda_dict(inp))
    model.inject = {}


def test_index_error_when_a_frame_has_no_background_pillar(fixture_weights):
    cfg, sd, inp, model = _setup(fixture_weights)
    T, ny, nx = _grid(cfg, inp)
    from oracle import oracle

    inject = {"fb_est_map": torch.ones(1, T, 1, ny, nx, dtype=torch.int64)}
    with pytest.raises(IndexError):
        torch.manual_seed(11)
        oracle.OracleMotionNet(cfg, sd, inject=inject).forward(inp)
    model.inject = dict(inject)
    with pytest.raises(IndexError):
        torch.manual_seed(11)
        model(cuda_dict(inp))
    model.inject = {}
    torch.cuda.synchronize()
    # the model is usable afterwards
    torch.manual_seed(11)
    out = model(cuda_dict(inp))
    assert torch.isfinite(out["rec_est"]).all()


def test_degenerate_kabsch_falls_back_to_identity(fixture_weights):
    """All-zero geometric features: 0/0 in the L2 normalisation -> NaN affinities -> NaN covariance.  torch.svd raises on the
    CPU and the reference returns R = I, t = 0 for the pair (register_utils.py:295-304); so must the kernel."""
    from oracle import oracle

    cfg, sd, inp, model = _setup(fixture_weights)
    sd = dict(sd)
    last = sorted(k for k in sd if k.startswith("ego_feats_head.") and k.endswith(".weight") and sd[k].dim() == 4)[-1]
    sd[last] = torch.zeros_like(sd[last])
    sd[last.replace(".weight", ".bias")] = torch.zeros_like(sd[last.replace(".weight", ".bias")])
    model.load_state_dict(sd)
    torch.manual_seed(3)
    ref = oracle.OracleMotionNet(cfg, sd).forward(inp)
    model.inject = {"fb_est_map": ref["fb_seg_est"].max(dim=2, keepdim=True)[1]}  # identical background sets (protocol.py)
    torch.manual_seed(3)
    out = model(cuda_dict(inp))
    model.inject = {}
    eye = torch.eye(4).expand_as(ref["ego_motion_est"])
    assert torch.equal(ref["ego_motion_est"].float(), eye), "the oracle (= the reference) takes the identity branch here"
    assert torch.equal(out["ego_motion_est"].float().cpu(), eye)
    assert all(bool(torch.isnan(p).all()) for p in ref["perm_matrix"]) and all(bool(torch.isnan(p).all()) for p in out["perm_matrix"])
    # downstream of the identity pose everything is well defined again
    assert torch.equal(out["transformed_points"].cpu(), ref["transformed_points"])
    assert torch.equal(out["fb_est_per_points"].cpu(), ref["fb_est_per_points"])
