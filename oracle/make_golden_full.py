"""TEST INFRASTRUCTURE ONLY -- full-size goldens of the BASELINE.json configurations from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python oracle/make_golden_full.py [C2 C3 C5]
Writes tests/golden/full_<workload>.npz: outputs of reference ``MotionNet.forward`` (test mode, fixture weights seed 42,
``torch.manual_seed(42)``) on ``synth.make_workload_scene(<workload>, 0)`` -- the scene itself is regenerated from its
seed by the tests, so only outputs are stored:
  * integer outputs in full, bit-packed / compressed (FG/BG label per point, motion label per point, instance labels),
  * sha256 of the voxeliser outputs (pure IEEE float32 / integer arithmetic: identical on every CPU),
  * float outputs sampled every ``stride`` rows (ego poses and scalars in full).
The reference ships no golden vectors of its own; these are outputs of the reference itself.
"""
import hashlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from pcaccumulation_b200 import config, fixture, synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
N_SAMPLE_ROWS = 6000


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def full_golden(ns, name):
    w = config.WORKLOADS[name]
    cfg = ref_loader.reference_config(w["dataset"], "test", w["overrides"] or None)
    model = ns["MotionNet"](cfg).eval()
    sd = fixture.fixture_state_dict(model.state_dict(), 42)
    model.load_state_dict(sd)
    scene = synth.make_workload_scene(name, 0)
    pts4 = np.concatenate((scene["input_points"], scene["time_indice"]), 1).astype(np.float32)
    v = ns["Voxelization"](cfg["voxel_generator"])(pts4)
    sample = dict(scene)
    sample.update(v)
    inp = ns["collate_fn"]([sample])
    stages = {}
    model.unet.register_forward_hook(lambda m, i, o: stages.__setitem__("bev_feats", o))
    torch.manual_seed(42)
    t0 = time.time()
    with torch.no_grad():
        res = model(inp)
    dt = time.time() - t0
    n = pts4.shape[0]
    stride = max(1, n // N_SAMPLE_ROWS)
    out = {
        "n_points": np.array([n]), "stride": np.array([stride]), "points_sha256": np.array(sha(pts4)),
        "vox_coordinates_sha256": np.array(sha(v["coordinates"].astype(np.int32))),
        "vox_p2v_sha256": np.array(sha(v["point_to_voxel_map"].astype(np.int64))), "vox_num_voxels": v["num_voxels"],
        "fb_bits": np.packbits(res["fb_est_per_points"][:, 0].numpy().astype(np.uint8)),
        "mos_bits": np.packbits(res["mos_est"].argmax(1).numpy().astype(np.uint8)),
        "inst_labels_est": res["inst_labels_est"].numpy().astype(np.int16),
        "inst_labels_adjusted": res["inst_labels_adjusted"].numpy().astype(np.int16),
        "ego_motion_est": res["ego_motion_est"].numpy(), "ego_motion_gt": res["ego_motion_gt"].numpy(),
        "inst_pose_est": res["inst_pose_est"].numpy(),
        "scalars": np.array([float(res["ego_l1_loss"]), float(res["ego_l2_loss"]), res["ego_rot_error"], res["ego_trans_error"],
                             res["inst_l2_error"], res["dynamic_inst_l2_error"]]),
        "perm_rowsum": np.stack([p[0].sum(1).numpy() for p in res["perm_matrix"]]),
        "fb_seg_est_sample": res["fb_seg_est"][:, :, :, ::8, ::8].numpy(),
        "bev_feats_sample": stages["bev_feats"][:, :, ::24, ::24].numpy(),
    }
    for k in ("transformed_points", "mos_est", "offset_est", "rec_est"):
        out[k + "_sample"] = res[k][::stride].numpy()
    out["sub_rec_est_sample"] = res["sub_rec_est"][::max(1, stride // 8)].numpy()
    path = os.path.join(GOLD, f"full_{name}.npz")
    np.savez_compressed(path, **out)
    print(name, "N", n, "M", int(v["num_voxels"][0]), "inst", int(res["inst_labels_est"].max()),
          "FG", float((res["fb_est_per_points"] == 1).float().mean()), "dyn", float(res["mos_est"].argmax(1).float().mean()),
          f"reference forward {dt:.1f} s on {torch.get_num_threads()} threads;", os.path.getsize(path) // 1024, "KiB")


def main():
    ns = ref_loader.load()
    for name in (sys.argv[1:] or ["C2", "C3", "C5"]):
        full_golden(ns, name)


if __name__ == "__main__":
    main()
