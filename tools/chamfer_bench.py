"""Developer script: Chamfer forward at the C3 size (n = m = all points of a nuScenes-shaped scene), exact grid search vs the
every-pair kernel.  usage: python tools/chamfer_bench.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pcaccumulation_b200 import synth  # noqa: E402
from pcaccumulation_b200.chamfer_distance import chamfer_with_indices  # noqa: E402

s = synth.make_workload_scene("C3", 0)
p = torch.tensor(s["input_points"]).cuda()
ang = 0.01
R = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], dtype=torch.float32).cuda()
est = (p @ R.T + torch.tensor([0.05, -0.02, 0.01]).cuda())[None].contiguous()
gt = p[None].contiguous()
n = p.shape[0]
for brute in (False, True):
    for _ in range(2):
        chamfer_with_indices(gt, est, brute=brute)
    torch.cuda.synchronize()
    reps = 3 if brute else 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = chamfer_with_indices(gt, est, brute=brute)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{'every pair' if brute else 'grid      '}: n = m = {n}  {ms:8.3f} ms   pair-evals/s (algorithmic 2nm) {2.0 * n * n / ms * 1e3:.3e}")
