"""TEST INFRASTRUCTURE ONLY -- generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
Writes
  tests/golden/state_dict_manifest.json   names/shapes/dtypes of the reference MotionNet.state_dict()
  tests/golden/forward_<name>.npz         inputs + every output of reference MotionNet.forward (test mode) on a
                                          small synthetic scene with the fixture weights (seed 42), plus float64
                                          checksums of the large stage tensors
  tests/golden/chamfer.npz                reference chamfer_distance CPU extension outputs on random clouds
The reference ships no golden vectors of its own; these are outputs of the reference itself.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from pcaccumulation_b200 import fixture, synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def forward_golden(ns, name, dataset, T, ppf, overrides, seed):
    cfg = ref_loader.reference_config(dataset, "test", overrides)
    model = ns["MotionNet"](cfg).eval()
    sd = fixture.fixture_state_dict(model.state_dict(), 42)
    model.load_state_dict(sd)
    rng = cfg["voxel_generator"]["range"]
    crop = cfg["voxel_generator"]["crop_range"]
    ground = cfg["data"]["ground_height"] + cfg["data"]["ground_slack"]
    scene = synth.make_scene(T=T, pts_per_frame=ppf, seed=seed, freq=cfg["data"]["freq"], z_lo=max(ground, rng[2]) + 0.06,
                             z_hi=min(crop[2], rng[5]) - 0.05, crop_xy=float(crop[0]), max_range=1.4 * float(crop[0]), n_boxes=24)
    pts4 = np.concatenate((scene["input_points"], scene["time_indice"]), 1).astype(np.float32)
    v = ns["Voxelization"](cfg["voxel_generator"])(pts4)
    sample = dict(scene)
    sample.update(v)
    inp = ns["collate_fn"]([sample])
    stages = {}
    model.unet.register_forward_hook(lambda m, i, o: stages.__setitem__("bev_feats", o))
    model.pillar_encoder.register_forward_hook(lambda m, i, o: stages.__setitem__("pillar_feats", o))
    torch.manual_seed(42)
    with torch.no_grad():
        res = model(inp)
    out = {
        "in_points4": pts4,
        "in_fb_labels": scene["fb_labels"][:, 0].astype(np.int8), "in_sd_labels": scene["sd_labels"][:, 0].astype(np.int8),
        "in_inst_labels": scene["inst_labels"][:, 0].astype(np.int16), "in_ego_motion_gt": scene["ego_motion_gt"],
        "in_inst_motion_gt": scene["inst_motion_gt"],
        # reference voxeliser outputs (libs/voxel_generator.py, numba): pins the voxeliser restatement
        "vox_coordinates": v["coordinates"].astype(np.int32), "vox_point_to_voxel_map": v["point_to_voxel_map"][:, 0].astype(np.int32),
        "vox_num_voxels": v["num_voxels"], "vox_shape": v["shape"],
    }
    for k in ("fb_est_per_points", "ego_motion_est", "ego_motion_gt", "transformed_points", "mos_est", "offset_est", "rec_est",
              "inst_labels_est", "inst_pose_est", "inst_labels_adjusted", "sub_rec_est"):
        if k in res:
            out["out_" + k] = res[k].numpy()
    out["out_scalars"] = np.array([float(res["ego_l1_loss"]), float(res["ego_l2_loss"]), res["ego_rot_error"],
                                   res["ego_trans_error"], res.get("inst_l2_error", np.nan), res.get("dynamic_inst_l2_error", np.nan)])
    out["out_fb_seg_est_sum"] = np.array([res["fb_seg_est"].double().sum().item(), res["fb_seg_est"].double().abs().sum().item()])
    out["out_perm_rowsum"] = np.stack([p[0].sum(1).numpy() for p in res["perm_matrix"]])
    out["stage_pillar_feats_sub8"] = stages["pillar_feats"][::8].numpy()
    out["stage_bev_feats_sum"] = np.array([stages["bev_feats"].double().sum().item(), stages["bev_feats"].double().abs().sum().item()])
    out["stage_bev_feats_sample"] = stages["bev_feats"][:, :, ::16, ::16].numpy()
    # float32 rounding floor of THIS machine's run: distance of the reference's outputs from the float64 evaluation of the same
    # algorithm (oracle dtype=float64 following the reference's discrete decisions); the cross-CPU checks are derived from it
    from oracle.protocol import oracle_runs
    from pcaccumulation_b200 import synth as _synth

    v2 = dict(sample)
    _, r32, _, r64 = oracle_runs(cfg, sd, _synth.collate([v2]), 42)
    for k in ("ego_motion_est", "transformed_points", "mos_est", "offset_est", "rec_est", "inst_pose_est"):
        assert torch.equal(res[k], r32[k]), k  # the oracle restatement is bit-identical to the reference on the same CPU
        out["floor_" + k] = np.array([float((res[k].double() - r64[k]).abs().max())])
    np.savez_compressed(os.path.join(GOLD, f"forward_{name}.npz"), **out)
    print(name, "N", pts4.shape[0], "M", int(v["num_voxels"][0]), "inst", int(res["inst_labels_est"].max()),
          "FG", float((res["fb_est_per_points"] == 1).float().mean()))
    return model


def main():
    os.makedirs(GOLD, exist_ok=True)
    ns = ref_loader.load()
    model = forward_golden(ns, "waymo_small", "waymo", 5, 24000, None, 7)
    forward_golden(ns, "nuscene_small", "nuscene", 10, 12000, {"voxel_generator": {"n_sweeps": 10}, "data": {"n_frames": 10}}, 11)
    manifest = {k: [list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()}
    with open(os.path.join(GOLD, "state_dict_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=0)
    # chamfer: reference CPU extension
    import importlib
    cdm = importlib.import_module("chamfer_distance.chamfer_distance")
    g = np.random.default_rng(3)
    a = torch.tensor(g.normal(size=(2, 700, 3)).astype(np.float32))
    b = torch.tensor(g.normal(size=(2, 900, 3)).astype(np.float32))
    b[0, 5] = b[0, 3]  # exact tie -> lowest index must win
    d1, d2 = torch.zeros(2, 700), torch.zeros(2, 900)
    i1, i2 = torch.zeros(2, 700, dtype=torch.int), torch.zeros(2, 900, dtype=torch.int)
    cdm.cd.forward(a, b, d1, d2, i1, i2)
    g1, g2 = torch.tensor(g.normal(size=(2, 700)).astype(np.float32)), torch.tensor(g.normal(size=(2, 900)).astype(np.float32))
    ga, gb = torch.zeros_like(a), torch.zeros_like(b)
    cdm.cd.backward(a, b, ga, gb, g1, g2, i1, i2)
    np.savez_compressed(os.path.join(GOLD, "chamfer.npz"), xyz1=a.numpy(), xyz2=b.numpy(), dist1=d1.numpy(), dist2=d2.numpy(),
                        idx1=i1.numpy(), idx2=i2.numpy(), g1=g1.numpy(), g2=g2.numpy(), grad1=ga.numpy(), grad2=gb.numpy())
    print("chamfer golden written")
    alignment_golden(ns)


def alignment_golden(ns):
    """tests/golden/alignment.npz: the reference's BaseModel.align_frames / get_alignment_errors (models/tpointnet.py:95-163)."""
    import importlib

    tp = importlib.import_module("models.tpointnet")
    bm = tp.BaseModel(ref_loader.reference_config("waymo", "test", None))
    g = np.random.default_rng(12)
    n = 6000
    pts = torch.tensor(g.uniform(-30, 30, (n, 3)).astype(np.float32))
    t = torch.tensor(g.integers(0, 5, n))

    def pose(a, tx, ty):
        P = np.eye(4, dtype=np.float32)
        c, s = np.cos(a), np.sin(a)
        P[:2, :2] = [[c, -s], [s, c]]
        P[0, 3], P[1, 3] = tx, ty
        return P

    est = torch.tensor(np.stack([pose(0.01 * i, 0.3 * i, 0.1 * i) for i in range(5)]))
    gt = torch.tensor(np.stack([pose(0.011 * i, 0.31 * i, 0.09 * i) for i in range(5)]))
    cd, l2 = bm.get_alignment_errors(pts, t, est, gt)
    al = bm.align_frames(pts, t, est)
    np.savez_compressed(os.path.join(GOLD, "alignment.npz"), points=pts.numpy(), time=t.numpy(), est=est.numpy(), gt=gt.numpy(),
                        chamfer=np.array([float(cd)]), l2=np.array([float(l2)]), aligned=al.numpy())
    print("alignment golden written", float(cd), float(l2))


if __name__ == "__main__":
    main()
