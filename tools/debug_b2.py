import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import oracle
from pcaccumulation_b200 import config, fixture, synth
from pcaccumulation_b200.motionnet import MotionNet
cfg = config.workload_config("C1")
vg = cfg["voxel_generator"]
samples = []
for i in (11, 12):
    s = synth.make_workload_scene("C1", i, pts_per_frame=9000)
    p4 = np.concatenate((s["input_points"], s["time_indice"]), 1).astype(np.float32)
    s.update(oracle.voxelize(p4, vg["voxel_size"], vg["range"], vg["n_sweeps"]))
    samples.append(s)
inp = synth.collate(samples)
model = MotionNet(cfg).cuda().eval()
sd = fixture.fixture_state_dict(model.state_dict(), 42)
model.load_state_dict(sd)
model.keep_stages = True
orc = oracle.OracleMotionNet(cfg, sd)
torch.manual_seed(3)
ref = orc.forward(inp)
torch.manual_seed(3)
res = model({k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in inp.items()})
print("fb mism", (res["fb_est_per_points"].cpu() != ref["fb_est_per_points"]).sum().item())
print("pose diff per (b,t):", (res["ego_motion_est"].cpu() - ref["ego_motion_est"]).abs().amax(dim=(2, 3)))
print("counts", model.stages["counts"])
bev_o = orc.stages["bev_feats"].permute(0, 2, 3, 1)
print("bev diff per frame", (model.stages["bev_feats"].cpu() - bev_o).abs().amax(dim=(1, 2, 3)))
geo_o = orc.stages["geo_feats"].permute(0, 2, 3, 1)
g = model.stages["geo"].cpu()
g = g / g.norm(dim=3, keepdim=True)
print("geo diff per frame", (g - geo_o).abs().amax(dim=(1, 2, 3)))
for a, b in zip(res["perm_matrix"], ref["perm_matrix"]):
    print("perm diff", (a.cpu() - b).abs().max().item(), "rowsum", a.sum().item(), b.sum().item())
