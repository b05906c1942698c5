"""Developer script: time pcab_flow_eval at the C2 size against the torch formulation of libs/tester.py:58-88 on the GPU."""
import sys
import torch
sys.path.insert(0, ".")
from pcaccumulation_b200 import config, synth
from pcaccumulation_b200.evaluation import FlowEvaluator
cfg = config.workload_config("C2")
T = cfg["voxel_generator"]["n_sweeps"]
s = synth.make_workload_scene("C2", 0)
inp = synth.collate([dict(s, coordinates=__import__("numpy").zeros((1, 4), "int32"), num_voxels=__import__("numpy").array([1]),
                          shape=__import__("numpy").array([288, 288, 1, T]), point_to_voxel_map=__import__("numpy").zeros((s["input_points"].shape[0], 1), "int64"))])
inp = {k: (v.cuda() if isinstance(v, torch.Tensor) else [x.cuda() for x in v]) for k, v in inp.items()}
n = inp["input_points"].shape[0]
pred = {"rec_est": inp["input_points"].float() + 0.01 * torch.randn(n, 3, device="cuda"), "mos_est": torch.randn(n, 2, device="cuda"),
        "fb_est_per_points": (torch.rand(n, 1, device="cuda") < 0.3).long()}
ev = FlowEvaluator(T, keep_per_point=False)
for _ in range(3): ev.update(inp, pred)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ev.update(inp, pred)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("FlowEvaluator.update (wrapper + kernel): %.3f ms/scene for %d points" % (ms, n))
# kernel alone, inputs prepared once
from pcaccumulation_b200._lib import I, P, call, stream
pts = inp["input_points"].float().contiguous(); tidx = inp["time_indice"][:, 1].to(torch.int32).contiguous()
ego = inp["ego_motion_gt"].float()[0].contiguous(); ig = inp["inst_motion_gt"][0].float().contiguous()
i64 = lambda t: t.reshape(-1).to(torch.int64).contiguous()
inst, fb, sd, fbe = i64(inp["inst_labels"]), i64(inp["fb_labels"]), i64(inp["sd_labels"]), i64(pred["fb_est_per_points"])
epe = torch.empty(n, device="cuda"); rel = torch.empty(n, device="cuda")
def k():
    call("pcab_flow_eval", P(pts), P(tidx), P(pred["rec_est"]), P(ego), P(inst), P(ig), I(ig.shape[0]), P(fb), P(sd), P(pred["mos_est"]), P(fbe),
         I(n), I(T), P(epe), P(rel), P(ev.sf), P(ev.mos), stream())
for _ in range(3): k()
e0.record()
for _ in range(50): k()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
nbytes = n * (12 + 4 + 12 + 8 * 4 + 8 + 8)
print("pcab_flow_eval kernel: %.1f us, %.0f GB/s algorithmic (%.1f MB per scene: 68 B read + 8 B written per point)" % (ms * 1e3, nbytes / ms / 1e6, nbytes / 1e6))
# torch formulation of the reference (same ops as libs/tester.py:58-88 + loss.get_mos_loss metric), on the GPU
from oracle import oracle
cpu_like = {"input_points": inp["input_points"], "time_indice": inp["time_indice"], "ego_motion_gt": inp["ego_motion_gt"], "inst_motion_gt": inp["inst_motion_gt"],
            "inst_labels": inp["inst_labels"], "fb_labels": inp["fb_labels"], "sd_labels": inp["sd_labels"]}
for _ in range(2): oracle.flow_eval(cpu_like, pred, T)
torch.cuda.synchronize()
e0.record()
for _ in range(5): oracle.flow_eval(cpu_like, pred, T)
e1.record(); torch.cuda.synchronize()
print("torch formulation of the reference tail on the same GPU: %.3f ms/scene" % (e0.elapsed_time(e1) / 5))
