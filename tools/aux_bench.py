"""Developer script: device time of the rows widened after the forward on a C2 scene --
FuseLoss forward (+ backward to the network outputs), the ICP refinement branches, the augmented data front-end.
usage: python tools/aux_bench.py"""
import copy
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pcaccumulation_b200 import config, fixture, synth  # noqa: E402
from pcaccumulation_b200.loss import FuseLoss  # noqa: E402
from pcaccumulation_b200.motionnet import MotionNet  # noqa: E402
from pcaccumulation_b200.runner import SceneRunner  # noqa: E402
from pcaccumulation_b200.voxel_generator import Voxelization  # noqa: E402


def timed(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def scene_input(cfg, name, idx):
    s = dict(synth.make_workload_scene(name, idx))
    vg = cfg["voxel_generator"]
    p4 = np.concatenate((s["input_points"], s["time_indice"]), 1).astype(np.float32)
    v = Voxelization(vg)(torch.tensor(p4).cuda())
    s.update({k: v[k].cpu().numpy() for k in ("coordinates", "num_voxels", "shape", "point_to_voxel_map")})
    inp = synth.collate([s])
    return s, {k: (t.cuda() if isinstance(t, torch.Tensor) else t) for k, t in inp.items()}


name = "C2"
cfg = config.workload_config(name, mode="val")
sd = fixture.fixture_state_dict(MotionNet(cfg).state_dict(), 42)
scene, inp = scene_input(cfg, name, 0)
n = inp["input_points"].shape[0]
model = MotionNet(cfg).cuda().eval()
model.load_state_dict(sd)
model.warmup()
torch.manual_seed(0)
pred = model(inp)
loss = FuseLoss(cfg["loss"])
print(f"{name}: {n} points")
print(f"forward (val mode, one scene at a time)      {timed(lambda: model(inp), 5):8.3f} ms")
print(f"FuseLoss forward                             {timed(lambda: loss(pred, inp)):8.3f} ms")


def fwd_bwd():
    p = dict(pred)
    for k in ("fb_seg_est", "mos_est", "offset_est"):
        p[k] = pred[k].detach().requires_grad_(True)
    loss(p, inp)["loss"].backward()


print(f"FuseLoss forward + backward to the outputs   {timed(fwd_bwd):8.3f} ms")

for branch in ("ego_icp", "tpointnet_icp"):
    c2 = copy.deepcopy(config.workload_config(name, mode="test"))
    base = MotionNet(c2).cuda().eval()
    base.load_state_dict(sd)
    base.warmup()
    inp_t = inp
    t_plain = timed(lambda: base(inp_t), 5)
    c2["model"][branch] = True
    m2 = MotionNet(c2).cuda().eval()
    m2.load_state_dict(sd)
    m2.warmup()
    t_icp = timed(lambda: m2(inp_t), 5)
    print(f"forward with {branch:14s} {t_icp:8.3f} ms  (without: {t_plain:.3f} ms -> the refinement costs {t_icp - t_plain:.3f} ms)")

runner = SceneRunner(config.workload_config(name))
sample = {"raw_points": torch.tensor(scene["input_points"]).cuda(), "time_indice": torch.tensor(scene["time_indice"][:, 0]).cuda(),
          **{k: torch.tensor(scene[k][:, 0]).cuda() for k in ("sd_labels", "fb_labels", "inst_labels")},
          "ego_motion_gt": scene["ego_motion_gt"], "inst_motion_gt": scene["inst_motion_gt"]}
print(f"front-end: crop + ground removal             {timed(lambda: runner.prep_raw(sample)):8.3f} ms (incl. the count readback)")
np.random.seed(0)
from pcaccumulation_b200 import dataset as ds  # noqa: E402

aug = ds.sample_augmentation(cfg["data_aug"], n, exact_noise=False)
print(f"front-end: + augmentation (device jitter)    {timed(lambda: runner.prep_raw(sample, aug)):8.3f} ms")
aug = ds.sample_augmentation(cfg["data_aug"], n, exact_noise=True)
noise_d = torch.as_tensor(aug["noise"]).cuda()
aug["noise"] = noise_d
print(f"front-end: + augmentation (host jitter, resident) {timed(lambda: runner.prep_raw(sample, aug)):8.3f} ms")
