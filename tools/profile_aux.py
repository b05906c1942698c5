"""Run, inside a cudaProfilerStart/Stop range, one Chamfer call at the C3 size and one FuseLoss forward on a C2 scene
(for ncu --profile-from-start off: the nearest-neighbour grid and the loss kernels)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pcaccumulation_b200 import config, fixture, synth  # noqa: E402
from pcaccumulation_b200.chamfer_distance import chamfer_with_indices  # noqa: E402
from pcaccumulation_b200.loss import FuseLoss  # noqa: E402
from pcaccumulation_b200.motionnet import MotionNet  # noqa: E402
from pcaccumulation_b200.voxel_generator import Voxelization  # noqa: E402

s3 = synth.make_workload_scene("C3", 0)
p = torch.tensor(s3["input_points"]).cuda()
moved = (p + torch.tensor([0.05, -0.02, 0.01]).cuda()).contiguous()
cfg = config.workload_config("C2", mode="val")
s = dict(synth.make_workload_scene("C2", 0))
p4 = np.concatenate((s["input_points"], s["time_indice"]), 1).astype(np.float32)
v = Voxelization(cfg["voxel_generator"])(torch.tensor(p4).cuda())
s.update({k: v[k].cpu().numpy() for k in ("coordinates", "num_voxels", "shape", "point_to_voxel_map")})
inp = {k: (t.cuda() if isinstance(t, torch.Tensor) else t) for k, t in synth.collate([s]).items()}
model = MotionNet(cfg).cuda().eval()
model.load_state_dict(fixture.fixture_state_dict(model.state_dict(), 42))
torch.manual_seed(0)
pred = model(inp)
loss = FuseLoss(cfg["loss"])
chamfer_with_indices(p[None], moved[None])
loss(pred, inp)
torch.cuda.synchronize()
torch.cuda.profiler.start()
chamfer_with_indices(p[None], moved[None])
loss(pred, inp)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled: chamfer n = m =", p.shape[0], "; FuseLoss on", inp["input_points"].shape[0], "points")
