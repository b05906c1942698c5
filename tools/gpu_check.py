"""Developer script: run the CUDA forward and the CPU oracle on one synthetic scene and print every difference."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import oracle  # noqa: E402
from pcaccumulation_b200 import config, fixture, synth  # noqa: E402
from pcaccumulation_b200.motionnet import MotionNet  # noqa: E402
from pcaccumulation_b200.voxel_generator import Voxelization  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C1"
ppf = int(sys.argv[2]) if len(sys.argv) > 2 else None
cfg = config.workload_config(name)
scene = synth.make_workload_scene(name, 0, pts_per_frame=ppf)
vg = cfg["voxel_generator"]
pts4 = np.concatenate((scene["input_points"], scene["time_indice"]), 1).astype(np.float32)
v = oracle.voxelize(pts4, vg["voxel_size"], vg["range"], vg["n_sweeps"])
vox = Voxelization(vg)
g = vox(torch.tensor(pts4).cuda())
print("voxelize: coords", np.array_equal(v["coordinates"], g["coordinates"].cpu().numpy()), "p2v",
      np.array_equal(v["point_to_voxel_map"], g["point_to_voxel_map"].cpu().numpy()), "M", int(g["num_voxels"][0]), int(v["num_voxels"][0]))
sample = dict(scene)
sample.update(v)
inp = synth.collate([sample])
model = MotionNet(cfg).cuda().eval()
sd = fixture.fixture_state_dict(model.state_dict(), 42)
model.load_state_dict(sd)
model.keep_stages = True
orc = oracle.OracleMotionNet(cfg, sd)
torch.manual_seed(42)
t = time.time()
ref = orc.forward(inp)
print("oracle s", time.time() - t)
for kk, vv in ref.get("tpointnet_loss_terms", {}).items():
    print(kk, {n: float(x) for n, x in vv.items() if n != "inst_est_motion"})
inp_g = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in inp.items()}
torch.manual_seed(42)
t = time.time()
out = model(inp_g)
torch.cuda.synchronize()
print("gpu s (cold)", time.time() - t)
torch.manual_seed(42)
t = time.time()
out = model(inp_g)
torch.cuda.synchronize()
print("gpu s (warm)", time.time() - t)


def cmp(a, b, name):
    if isinstance(a, torch.Tensor):
        b = b.cpu() if isinstance(b, torch.Tensor) else torch.tensor(b)
        if a.shape != b.shape:
            print(f"{name:40s} SHAPE {tuple(a.shape)} vs {tuple(b.shape)}")
            return
        if a.is_floating_point():
            d = (a - b).abs()
            rel = d.max().item() / (a.abs().max().item() + 1e-30)
            print(f"{name:40s} maxabs {d.max().item():.3e} rel-to-max {rel:.3e} mean {d.mean().item():.3e}")
        else:
            print(f"{name:40s} mismatches {(a != b).sum().item()} / {a.numel()}")
    elif isinstance(a, (float, int)):
        print(f"{name:40s} {a} vs {b}")
    elif isinstance(a, list):
        for i, (x, y) in enumerate(zip(a, b)):
            cmp(x, y, f"{name}[{i}]")
    elif isinstance(a, dict):
        for k in a:
            cmp(a[k], b[k], f"{name}.{k}")


for k in ref:
    if k in out:
        cmp(ref[k], out[k], k)
    else:
        print("MISSING", k)
# stages (oracle NCHW -> NHWC)
so, sg = orc.stages, model.stages
cmp(so["pillar_mean"], sg["pillar_mean"], "stage.pillar_mean")
cmp(so["pillar_feats"], sg["pillar_feats"], "stage.pillar_feats")
cmp(so["bev_feats"].permute(0, 2, 3, 1), sg["bev_feats"], "stage.bev_feats")
cmp(so["warped_feats"].permute(0, 2, 3, 4, 1).reshape(sg["warped"].shape), sg["warped"], "stage.warped")
if "mos_feats" in so and sg.get("mos_feats") is not None:
    cmp(so["mos_feats"].permute(0, 2, 3, 1), sg["mos_feats"], "stage.mos_feats")
if "backbone_feats" in so and "backbone_feats" in sg:
    cmp(so["backbone_feats"], sg["backbone_feats"], "stage.backbone_feats")
    cmp(so["motion_feats"], sg["motion_feats"], "stage.motion_feats")
