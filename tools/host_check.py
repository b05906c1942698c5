"""Developer script: warm per-kernel GPU time of ONE serial forward (torch profiler, not cold-cache ncu) vs wall, + host CPU time."""
import sys, time, collections
import torch
sys.path.insert(0, ".")
from pcaccumulation_b200 import config, fixture, synth
from pcaccumulation_b200.runner import SceneRunner, scene_to_points4
from torch.profiler import ProfilerActivity, profile
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
torch.set_num_threads(1)
cfg = config.workload_config(name)
runner = SceneRunner(cfg)
runner.model.load_state_dict(fixture.fixture_state_dict(runner.model.state_dict(), 42))
runner.warmup()
s = synth.make_workload_scene(name, 0)
p4 = torch.tensor(scene_to_points4(s)).cuda()
for i in range(4):
    torch.manual_seed(i); runner.run_device(p4, [p4.shape[0]])
torch.cuda.synchronize()
walls, cpus = [], []
for i in range(5):
    t0, c0 = time.perf_counter(), time.thread_time()
    torch.manual_seed(5); runner.run_device(p4, [p4.shape[0]]); torch.cuda.synchronize()
    walls.append((time.perf_counter() - t0) * 1e3); cpus.append((time.thread_time() - c0) * 1e3)
print(name, "wall ms", [round(w, 2) for w in walls], "thread cpu ms", [round(c, 2) for c in cpus])
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        torch.manual_seed(5); runner.run_device(p4, [p4.shape[0]])
    torch.cuda.synchronize()
agg = collections.OrderedDict(); tot = 0.0; n = 0
for e in prof.key_averages():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        t = e.device_time_total / 3e3
        agg[e.key[:70]] = (e.count / 3, t); tot += t; n += e.count / 3
print("warm GPU busy ms/scene %.3f  launches %.0f" % (tot, n))
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
    print(f"{t:8.3f} ms x{c:5.1f}  {k}")
