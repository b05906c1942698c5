"""TEST INFRASTRUCTURE ONLY -- the staged parity protocol shared by tests/, __graft_entry__.smoke() and bench.py's parity leg.

Bars (BASELINE.json north_star): integer outputs bit-exact, floats within REL = 1e-4 of the tensor's magnitude.
FP32 arithmetic does not meet those bars against ITSELF through the ~60 layers of this network: the reference run in
float32 (its own arithmetic, "ref32") and the same algorithm in float64 ("ref64": ``OracleMotionNet(dtype=float64)``
following ref32's discrete decisions) differ by up to 0.4e-4 in the motion logits and flip labels whose two logits tie
below that.  So every comparison measures that floor and derives its tolerance from it -- nothing is loosened by hand:
  float tensor:  |ours - ref32| <= 1e-4 * scale,  OR  |ours - ref64| <= 2 * |ref32 - ref64|   (the CUDA path may be at most
                 twice as far from the exact result as the reference's own float32 rounding puts the reference),
  label flips :  every point where ours != ref32 must have an exact-arithmetic margin |logit1 - logit0| (ref64) no larger
                 than twice the tolerance that was applied to the logits; anything else is a bug, not a rounding tie.
Stage-wise injection (one flipped FG/BG pillar changes the background count n, torch.randperm(n) then
draws other keypoints): "free" = nothing injected (checks everything up to and including the ego pose); "staged" =
ref32's FG/BG map, ego pose, motion logits and offsets injected, so each downstream stage starts from identical inputs.
"""
import json
import os

import torch

REL = 1e-4
REPORT = {}


def cuda_dict(d):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def _seeded(model, inp, seed, inject=None):
    model.inject = inject or {}
    torch.manual_seed(seed)
    out = model(inp)
    model.inject = {}
    return out


def _floor_cmp(name, ours, r32, r64, rec, rel=REL):
    o = torch.as_tensor(ours).detach().cpu().double()
    a = torch.as_tensor(r32).detach().cpu().double()
    b = torch.as_tensor(r64).detach().cpu().double()
    assert o.shape == a.shape == b.shape, (name, o.shape, a.shape, b.shape)
    scale = max(float(a.abs().max()), 1e-6)
    e32, e64, floor = float((o - a).abs().max()), float((o - b).abs().max()), float((a - b).abs().max())
    rec[name] = {"scale": scale, "err_vs_ref32": e32 / scale, "err_vs_ref64": e64 / scale, "fp32_floor": floor / scale}
    ok = e32 <= rel * scale or e64 <= 2.0 * floor
    rec[name]["ok"] = bool(ok)
    assert ok, (f"{name}: |ours-ref32| {e32 / scale:.2e} > {rel:.0e} and |ours-ref64| {e64 / scale:.2e} > 2 x the float32 floor "
                f"{floor / scale:.2e} (relative to {scale:.3e})")
    return max(rel * scale, 2.0 * floor)  # the absolute tolerance that was in force


def _flip_check(name, ours_lab, ref_lab, margin64, tol_abs, rec):
    flipped = (ours_lab != ref_lab)
    n = int(flipped.sum())
    worst = float(margin64[flipped].abs().max()) if n else 0.0
    rec[name] = {"flips": n, "of": int(ref_lab.numel()), "largest_exact_margin_at_a_flip": worst, "allowed_margin": 2.0 * tol_abs}
    assert worst <= 2.0 * tol_abs, f"{name}: {n} flips, one with exact margin {worst:.3e} > {2.0 * tol_abs:.3e}: not a rounding tie"
    return n


def _nchw(x):  # ours NHWC -> reference NCHW
    return x.permute(0, 3, 1, 2)


_ORACLE_CACHE = {}


def oracle_runs(cfg, sd, inp, seed, key=None):
    """(o32, r32, o64, r64): the float32 oracle run (free) and the float64 run that follows its discrete decisions.
    ``key``: when given, the last result is kept and reused for the same key (the two conv paths of one test case)."""
    from oracle import oracle

    if key is not None and _ORACLE_CACHE.get("key") == key:
        return _ORACLE_CACHE["val"]
    _ORACLE_CACHE.clear()
    test_mode = cfg["misc"]["mode"] == "test"
    o32 = oracle.OracleMotionNet(cfg, sd)
    torch.manual_seed(seed)
    r32 = o32.forward(inp)
    fb_map = r32["fb_seg_est"].max(dim=2, keepdim=True)[1]
    inj64 = {"fb_est_map": fb_map, "ego_motion_est": r32["ego_motion_est"], "mos_est": r32["mos_est"], "offset_est": r32["offset_est"],
             "transformed_points": r32["transformed_points"]}
    if test_mode:
        inj64["inst_labels_est"] = r32["inst_labels_est"]
    o64 = oracle.OracleMotionNet(cfg, sd, dtype=torch.float64, inject=inj64)
    torch.manual_seed(seed)
    r64 = o64.forward(inp)
    if key is not None:
        _ORACLE_CACHE.update(key=key, val=(o32, r32, o64, r64))
    return o32, r32, o64, r64


def run_protocol(model, cfg, sd, inp, seed, tag, cache_key=None):
    """Free-running + staged comparison of the CUDA path with the oracle (float32 = the reference's arithmetic, float64 =
    exact-arithmetic yardstick).  Returns (free-running results of the CUDA path, float32 oracle results)."""
    rec = REPORT.setdefault(tag, {})
    test_mode = cfg["misc"]["mode"] == "test"
    o32, r32, o64, r64 = oracle_runs(cfg, sd, inp, seed, cache_key)
    fb_map = r32["fb_seg_est"].max(dim=2, keepdim=True)[1]
    inj = {"fb_est_map": fb_map, "ego_motion_est": r32["ego_motion_est"], "mos_est": r32["mos_est"], "offset_est": r32["offset_est"]}
    inp_c = cuda_dict(inp)
    B, T = r32["ego_motion_est"].shape[:2]

    # ---------------- free-running: everything up to and including the ego pose ----------------
    a = _seeded(model, inp_c, seed)
    st = dict(model.stages)
    assert torch.equal(a["fb_seg_gt"].cpu(), r32["fb_seg_gt"]) and torch.equal(a["occ_map"].cpu(), r32["occ_map"])
    _floor_cmp("pillar_mean", st["pillar_mean"], o32.stages["pillar_mean"], o64.stages["pillar_mean"], rec)
    _floor_cmp("pillar_feats", st["pillar_feats"], o32.stages["pillar_feats"], o64.stages["pillar_feats"], rec)
    _floor_cmp("bev_feats", _nchw(st["bev_feats"]), o32.stages["bev_feats"], o64.stages["bev_feats"], rec)
    tol_fb = _floor_cmp("fb_seg_est", a["fb_seg_est"], r32["fb_seg_est"], r64["fb_seg_est"], rec)
    geo = _nchw(st["geo"]).cpu()
    _floor_cmp("geo_feats(normalised)", geo / geo.norm(dim=1, keepdim=True), o32.stages["geo_feats"], o64.stages["geo_feats"], rec)
    del geo
    # FG/BG decision per occupied cell and per point
    occ = r32["occ_map"] > 0
    m64 = (r64["fb_seg_est"][:, :, 1:2] - r64["fb_seg_est"][:, :, 0:1])
    ours_map = a["fb_seg_est"].cpu().max(dim=2, keepdim=True)[1]
    cells = _flip_check("fb_cells", ours_map[occ], fb_map[occ], m64[occ], tol_fb, rec)
    pt_flips = int((a["fb_est_per_points"].cpu() != r32["fb_est_per_points"]).sum())
    rec["fb_points"] = {"flips": pt_flips, "of": int(r32["fb_est_per_points"].numel())}
    assert (pt_flips == 0) == (cells == 0)
    rec["ref32_vs_ref64_fb_cells"] = int(((r64["fb_seg_est"].max(dim=2, keepdim=True)[1] != fb_map) & occ).sum())
    if cells:
        # continue from the reference's label map (identical background sets -> identical keypoint draws)
        a = _seeded(model, inp_c, seed, {"fb_est_map": fb_map})
        assert torch.equal(a["fb_est_per_points"].cpu(), r32["fb_est_per_points"])
    _floor_cmp("ego_motion_gt", a["ego_motion_gt"], r32["ego_motion_gt"], r64["ego_motion_gt"], rec)
    _floor_cmp("ego_motion_est", a["ego_motion_est"], r32["ego_motion_est"], r64["ego_motion_est"], rec)
    assert len(a["perm_matrix"]) == len(r32["perm_matrix"]) == B * (T - 1)
    for i, (x, y, z) in enumerate(zip(a["perm_matrix"], r32["perm_matrix"], r64["perm_matrix"])):
        _floor_cmp(f"perm_matrix[{i}]", x, y, z, rec)
    # (the rotation error is acos() of a value within 1e-7 of 1, ill-conditioned by construction: the float32 floor measured
    # on it is what bounds it, no hand-set tolerance)
    for k in ("ego_l1_loss", "ego_l2_loss", "ego_trans_error", "ego_rot_error"):
        _floor_cmp(k, torch.as_tensor(float(a[k])).reshape(1), torch.as_tensor(float(r32[k])).reshape(1),
                   torch.as_tensor(float(r64[k])).reshape(1), rec)
    # un-injected end-to-end agreement (reported): per-point labels and accumulated points of the free run
    if test_mode:
        rec["free_running"] = {
            "mos_label_mismatches": int((a["mos_est"].cpu().argmax(1) != r32["mos_est"].argmax(1)).sum()),
            "inst_label_mismatches": int((a["inst_labels_est"].cpu() != r32["inst_labels_est"]).sum()),
            "rec_est_median_err_m": float((a["rec_est"].cpu() - r32["rec_est"]).norm(dim=1).median()),
            "ref32_vs_ref64_pose_err": float((r32["ego_motion_est"].double() - r64["ego_motion_est"]).abs().max())}
        assert rec["free_running"]["rec_est_median_err_m"] < 1e-4

    # ---------------- staged: reference FG/BG map, pose, motion logits, offsets injected ----------------
    s = _seeded(model, inp_c, seed, inj)
    st = dict(model.stages)
    tp_exact = torch.equal(s["transformed_points"].cpu(), r32["transformed_points"])
    rec["transformed_points_bit_equal"] = tp_exact
    _floor_cmp("transformed_points", s["transformed_points"], r32["transformed_points"], r64["transformed_points"], rec, rel=1e-5)
    w32, w64 = o32.stages["warped_feats"], o64.stages["warped_feats"]  # [B,C,T,H,W]
    ours_w = st["warped"].view(B, T, *st["warped"].shape[1:]).permute(0, 4, 1, 2, 3)
    _floor_cmp("warped_feats", ours_w, w32, w64, rec)
    if "mos_feats" in o32.stages:
        _floor_cmp("mos_feats", _nchw(st["mos_feats"]), o32.stages["mos_feats"], o64.stages["mos_feats"], rec)
    tol_mos = _floor_cmp("mos_est", s["mos_est"], r32["mos_est"], r64["mos_est"], rec)
    _floor_cmp("offset_est", s["offset_est"], r32["offset_est"], r64["offset_est"], rec)
    mm64 = r64["mos_est"][:, 1] - r64["mos_est"][:, 0]
    _flip_check("mos_points", s["mos_est"].cpu().argmax(1), r32["mos_est"].argmax(1), mm64, tol_mos, rec)
    rec["ref32_vs_ref64_mos_points"] = int((r64["mos_est"].argmax(1) != r32["mos_est"].argmax(1)).sum())
    if not tp_exact:
        # the 5 cm dedupe hash of the clustering quantises (transformed point + offset): to compare the INTEGER pipeline
        # bit for bit it has to start from bit-identical coordinates
        s = _seeded(model, inp_c, seed, dict(inj, transformed_points=r32["transformed_points"]))
        st = dict(model.stages)
    if test_mode:
        assert torch.equal(s["inst_labels_est"].cpu(), r32["inst_labels_est"]), "instance labels (clustering) must be bit-exact"
    assert "inst_pose_est" in r32, "the test scene must exercise the TubeNet branch"
    assert torch.equal(s["inst_labels_adjusted"].cpu(), r32["inst_labels_adjusted"])
    _floor_cmp("backbone_feats", st["backbone_feats"], o32.stages["backbone_feats"], o64.stages["backbone_feats"], rec)
    _floor_cmp("motion_feats", st["motion_feats"], o32.stages["motion_feats"], o64.stages["motion_feats"], rec)
    for k in ("inst_pose_est", "sub_rec_est", "rec_est"):
        _floor_cmp(k, s[k], r32[k], r64[k], rec)
    for k in ("inst_l2_error", "dynamic_inst_l2_error"):
        _floor_cmp(k, torch.as_tensor(float(s[k])).reshape(1), torch.as_tensor(float(r32[k])).reshape(1),
                   torch.as_tensor(float(r64[k])).reshape(1), rec)
    for it, terms in r32["tpointnet_loss_terms"].items():
        t64 = r64["tpointnet_loss_terms"][it]
        _floor_cmp(f"inst_est_motion[{it}]", s["tpointnet_loss_terms"][it]["inst_est_motion"], terms["inst_est_motion"], t64["inst_est_motion"], rec)
        for name in ("l1_loss", "l2_loss", "rot_loss", "trans_loss"):
            x, y, z = float(s["tpointnet_loss_terms"][it][name]), float(terms[name]), float(t64[name])
            _floor_cmp(f"{name}[{it}]", torch.tensor([x]), torch.tensor([y]), torch.tensor([z]), rec)
    write_report()
    return a, r32


def write_report():
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.json"), "w") as f:
            json.dump(REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass


