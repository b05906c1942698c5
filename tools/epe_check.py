"""Developer script: free-running GPU forward vs the oracle on bench scenes: FG flips, pose difference, EPE (both conv paths)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import fixture_weights, oracle_forward_fn  # noqa: E402
from pcaccumulation_b200 import config, synth  # noqa: E402
from pcaccumulation_b200.runner import SceneRunner, scene_to_points4  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = config.workload_config(name)
sd = fixture_weights(cfg)
runner = SceneRunner(cfg)
runner.model.load_state_dict(sd)
torch.set_num_threads(32)
run = oracle_forward_fn(cfg, sd)
for i in range(n):
    s = synth.make_workload_scene(name, i)
    ref = run(s)
    p4 = torch.tensor(scene_to_points4(s)).cuda()
    for tc in (False, True):
        runner.model.use_tensor_cores = tc
        torch.manual_seed(42)
        res = runner.run_device(p4, [p4.shape[0]], ego_motion_gt=torch.tensor(s["ego_motion_gt"])[None].cuda())
        fb = (res["fb_est_per_points"].cpu() != ref["fb_est_per_points"]).sum().item()
        pose = (res["ego_motion_est"].cpu() - ref["ego_motion_est"]).abs().max().item()
        epe = (res["rec_est"].cpu() - ref["rec_est"]).norm(dim=1)
        inst = (res["inst_labels_est"].cpu() != ref["inst_labels_est"]).sum().item() if "inst_labels_est" in ref else -1
        mos = (res["mos_est"].cpu().argmax(1) != ref["mos_est"].argmax(1)).sum().item()
        print(f"scene {i} tc={tc}: fb flips {fb} pose maxabs {pose:.3e} mos flips {mos} inst mism {inst} EPE mean {epe.mean():.3e} max {epe.max():.3e}",
              flush=True)
