"""Runs ON THE GPU BOX: one small test-mode scene through the CUDA forward, the device FuseLoss and the device evaluation
tail; writes the results dict (the API boundary of SURVEY.md section 8b) and the device consumers' numbers to
gpurun_out/gpu_results.npz.  tests/test_host_logic.py::test_reference_consumers_accept_gpu_results feeds that dict to the
UNMODIFIED reference consumers (libs/loss.py:FuseLoss, the metric code of libs/tester.py:58-93) in the build container.

The grid is 128 x 128 (range +-16 m) and only the first soft-assignment matrix is kept (one is 4 MB) so that the fixture
stays small; both consumers see the same truncated list.   usage: python tools/make_gpu_results_fixture.py [out.npz]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pcaccumulation_b200 import config, fixture, synth  # noqa: E402
from pcaccumulation_b200.evaluation import ClusterEvaluator, FlowEvaluator  # noqa: E402
from pcaccumulation_b200.loss import FuseLoss  # noqa: E402
from pcaccumulation_b200.motionnet import MotionNet  # noqa: E402
from pcaccumulation_b200.voxel_generator import Voxelization  # noqa: E402

LOSS_W = {"w_pose_l1_loss": 1.0, "w_perm_loss": 0.005, "w_mos_bce_loss": 1.0, "w_mos_lovasz_loss": 1.0, "w_fb_bce_loss": 1.0,
          "w_fb_lovasz_loss": 1.0, "w_offset_norm_loss": 0.5, "w_offset_dir_loss": 0.5, "w_obj_l1_loss": 1.0, "w_obj_pose_loss": 1.0,
          "w_obj_loss": 0.3, "w_obj_rot_loss": 50, "w_obj_trans_loss": 1.0, "obj_gamma": 0.7}  # configs/default.yaml:99-113


def small_config():
    return config.get_config("waymo", mode="test", voxel_generator={"range": [-16, -16, -2, 16, 16, 6], "crop_range": [14, -2, 6]})


def main(out_path):
    cfg = small_config()
    vg = cfg["voxel_generator"]
    T = vg["n_sweeps"]
    s = dict(synth.make_scene(T=T, pts_per_frame=6000, seed=77, freq=cfg["data"]["freq"], z_lo=0.45, z_hi=5.9, crop_xy=14.0,
                              max_range=20.0, n_boxes=14))
    p4 = np.concatenate((s["input_points"], s["time_indice"]), 1).astype(np.float32)
    v = Voxelization(vg)(torch.tensor(p4).cuda())
    s.update({k: v[k].cpu().numpy() for k in ("coordinates", "num_voxels", "shape", "point_to_voxel_map")})
    inp = synth.collate([s])
    inp_c = {k: (t.cuda() if isinstance(t, torch.Tensor) else t) for k, t in inp.items()}
    model = MotionNet(cfg).cuda().eval()
    model.load_state_dict(fixture.fixture_state_dict(model.state_dict(), 42))
    torch.manual_seed(5)
    pred = model(inp_c)
    assert "inst_pose_est" in pred and int((pred["inst_labels_est"] != 0).sum()) > 15, "scene must reach TubeNet"
    pred["perm_matrix"] = pred["perm_matrix"][:1]
    loss = FuseLoss(LOSS_W)
    stats = loss(pred, inp_c)
    fe = FlowEvaluator(T)
    epe, rel = fe.update(inp_c, pred)
    ce = ClusterEvaluator()
    ce.update(pred["inst_labels_est"], inp_c["inst_labels"][:, 0], inp_c["sd_labels"][:, 0])
    out = {}
    for k, t in inp.items():
        if isinstance(t, torch.Tensor):
            out["in_" + k] = t.numpy()
    out["in_inst_motion_gt"] = inp["inst_motion_gt"][0].numpy()
    for k, t in pred.items():
        if isinstance(t, torch.Tensor):
            out["out_" + k] = t.detach().cpu().numpy()
        elif isinstance(t, float):
            out["out_" + k] = np.array([t])
    p0 = pred["perm_matrix"][0].cpu().numpy()
    nz = np.flatnonzero(p0)
    out["perm0_shape"], out["perm0_idx"], out["perm0_val"] = np.array(p0.shape), nz.astype(np.int32), p0.reshape(-1)[nz]
    for it, terms in pred["tpointnet_loss_terms"].items():
        for name in ("l1_loss", "l2_loss", "rot_loss", "trans_loss", "inst_est_motion"):
            out[f"tpn_{it}_{name}"] = np.asarray(terms[name].detach().cpu().numpy() if isinstance(terms[name], torch.Tensor) else terms[name])
    for k, t in stats.items():
        if k.endswith("_metric"):
            for name, arr in t.items():
                out[f"stat_{k}_{name}"] = np.asarray(arr)
        else:
            out["stat_" + k] = np.array([float(t)])
    out["eval_epe"], out["eval_rel"] = epe.cpu().numpy(), rel.cpu().numpy()
    out["eval_sf"], out["eval_mos"] = fe.sf.cpu().numpy(), fe.mos.cpu().numpy()
    out["eval_cluster"] = ce.counters.cpu().numpy()
    # compact storage of the big, highly redundant maps
    out["out_fb_seg_gt"] = out["out_fb_seg_gt"].astype(np.int8)
    out["out_occ_map"] = out["out_occ_map"].astype(np.uint8)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, {k: v.shape for k, v in out.items() if v.size > 10000})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/gpu_results.npz")
