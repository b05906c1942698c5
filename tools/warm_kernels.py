"""Developer script: warm per-kernel GPU times of one scene (torch.profiler / CUPTI; graphs are replayed as in production).

usage: python tools/warm_kernels.py [C2] [--no-tc] [--per-launch PATTERN]
Prints, per kernel name, the time summed over one scene (mean of 5 scenes) and, with --per-launch, every launch of the
kernels matching PATTERN in order (the per-layer view of the convolution stacks).
"""
import collections
import sys
import time

import torch

sys.path.insert(0, ".")
from pcaccumulation_b200 import config, fixture, synth  # noqa: E402
from pcaccumulation_b200.runner import SceneRunner, scene_to_points4  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
name = args[0] if args else "C2"
pattern = None
if "--per-launch" in sys.argv:
    pattern = sys.argv[sys.argv.index("--per-launch") + 1]
    args = [a for a in args if a != pattern]
    name = args[0] if args else "C2"
cfg = config.workload_config(name)
runner = SceneRunner(cfg)
runner.model.load_state_dict(fixture.fixture_state_dict(runner.model.state_dict(), 42))
runner.model.use_tensor_cores = "--no-tc" not in sys.argv
if "--tf32" in sys.argv:
    runner.model.conv_operands = "tf32"
runner.warmup()
s = synth.make_workload_scene(name, 0)
p4 = torch.tensor(scene_to_points4(s)).cuda()
for i in range(4):
    torch.manual_seed(i)
    runner.run_device(p4, [p4.shape[0]])
torch.cuda.synchronize()
walls = []
for i in range(5):
    t0 = time.perf_counter()
    torch.manual_seed(5)
    runner.run_device(p4, [p4.shape[0]])
    torch.cuda.synchronize()
    walls.append((time.perf_counter() - t0) * 1e3)
N = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(N):
        torch.manual_seed(5)
        runner.run_device(p4, [p4.shape[0]])
    torch.cuda.synchronize()
agg = collections.OrderedDict()
launches = []
for e in prof.events():
    if e.device_type != torch.autograd.DeviceType.CUDA:
        continue
    t = e.device_time if hasattr(e, "device_time") else e.cuda_time
    agg.setdefault(e.name, [0, 0.0])
    agg[e.name][0] += 1
    agg[e.name][1] += t
    launches.append((e.time_range.start, e.name, t))
tot = sum(v[1] for v in agg.values())
print(f"{name} wall ms {[round(w, 2) for w in walls]}")
print(f"warm GPU busy ms/scene {tot / N / 1e3:.3f}  launches {sum(v[0] for v in agg.values()) / N:.0f}")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:48]:
    print(f"{t / N / 1e3:8.3f} ms x{c / N:5.1f}  {k[:100]}")
if pattern:
    launches.sort()
    sel = [(n_, t) for _, n_, t in launches if pattern in n_]
    per = len(sel) // N
    print(f"--- launches matching {pattern!r} of the last scene ({per}):")
    for i, (n_, t) in enumerate(sel[-per:]):
        print(f"  {i:3d} {t:8.1f} us  {n_[:90]}")
