"""Per-stage GPU-timeline durations of one warm forward (events at stage boundaries) + host wall time."""
import sys, time
import torch
sys.path.insert(0, ".")
from pcaccumulation_b200 import config, fixture, synth
from pcaccumulation_b200.runner import SceneRunner, scene_to_points4
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
torch.set_num_threads(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
cfg = config.workload_config(name)
runner = SceneRunner(cfg)
runner.model.load_state_dict(fixture.fixture_state_dict(runner.model.state_dict(), 42))
s = synth.make_workload_scene(name, 0)
p4 = torch.tensor(scene_to_points4(s)).cuda()
for i in range(3):
    torch.manual_seed(i); runner.run_device(p4, [p4.shape[0]])
torch.cuda.synchronize()
acc = {}
walls = []
for it in range(5):
    runner.model.stage_marks = []
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
    t0 = time.perf_counter()
    torch.manual_seed(5); runner.run_device(p4, [p4.shape[0]])
    torch.cuda.synchronize()
    walls.append((time.perf_counter() - t0) * 1e3)
    prev = e0
    for nm, e in runner.model.stage_marks:
        acc.setdefault(nm, []).append(prev.elapsed_time(e)); prev = e
print("threads", torch.get_num_threads(), "wall ms", [round(w, 2) for w in walls])
tot = 0
order = ["voxelize+build_input", "start", "index+stats", "pillar_encoder", "unet", "fb_head", "ego", "warp+stpn", "cluster"]
labels = {"start": "voxelize + build_input (runner)", "index+stats": "casts + pillar index/stats/canvases", "pillar_encoder": "pillar encoder",
          "unet": "unet", "fb_head": "fb head", "ego": "ego (2 head convs + pairs)", "warp+stpn": "warp + stpn + point head", "cluster": "cluster", "tubenet": "tubenet"}
for k, v in acc.items():
    m = sum(v) / len(v); tot += m
    print(f"{labels.get(k, k):32s} {m:7.3f} ms")
print("sum", round(tot, 3))
# host cost of the randperm protocol
import torch as T
t0 = time.perf_counter()
for _ in range(8): T.randperm(48000)[:1024]
print("8x randperm(48000) host ms", (time.perf_counter() - t0) * 1e3)
