"""Developer script: time ClusterEvaluator.update at the C2 size against the reference-style mask loop on the same GPU."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from pcaccumulation_b200 import synth
from pcaccumulation_b200.evaluation import ClusterEvaluator
from oracle import oracle
s = synth.make_workload_scene("C2", 0)
gt = torch.tensor(s["inst_labels"][:, 0].astype(np.int64)).cuda()
mos = torch.tensor(s["sd_labels"][:, 0].astype(np.int64)).cuda()
est = gt.clone()
est[torch.rand(gt.shape, device="cuda") < 0.1] = 0
ev = ClusterEvaluator()
for _ in range(3): ev.update(est, gt, mos)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): ev.update(est, gt, mos)
torch.cuda.synchronize()
print("ClusterEvaluator.update: %.3f ms/scene (%d points, %d gt instances)" % ((time.perf_counter() - t0) * 50, gt.shape[0], int(gt.max())))
t0 = time.perf_counter()
oracle.cluster_eval(est, gt, mos)
torch.cuda.synchronize()
print("reference-style mask loop on the same GPU: %.1f ms/scene" % ((time.perf_counter() - t0) * 1e3))
