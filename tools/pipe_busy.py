"""Developer script: is ScenePipeline GPU-bound or host-bound?  Runs `steps` scenes through the pipeline under CUPTI and
reports (a) the union of all kernel intervals / the span (fraction of the time at least one kernel is running) and (b) the sum of
kernel durations / span (average concurrency).   usage: python tools/pipe_busy.py [depth] [steps]"""
import sys

import torch

sys.path.insert(0, ".")
from pcaccumulation_b200 import config, fixture, synth  # noqa: E402
from pcaccumulation_b200.runner import ScenePipeline, SceneRunner, scene_to_points4  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg = config.workload_config("C2")
sd = fixture.fixture_state_dict(SceneRunner(cfg).model.state_dict(), 42)
pipe = ScenePipeline(cfg, sd, depth=depth)
scenes = [synth.make_workload_scene("C2", i) for i in range(4)]
dev_pts = [torch.from_numpy(scene_to_points4(s)).cuda() for s in scenes]
nums = [[p.shape[0]] for p in dev_pts]


def run(n):
    futs = [pipe.submit(dev_pts[i % 4], nums[i % 4], seed=i, host=False) for i in range(n)]
    cur = torch.cuda.current_stream()
    for f in futs:
        _, done = f.result()
        cur.wait_event(done)
    torch.cuda.synchronize()


run(16)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(steps)
iv = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start:
        iv.append((e.time_range.start, e.time_range.end))
iv.sort()
span = iv[-1][1] - iv[0][0]
total = sum(b - a for a, b in iv)
union, cur_a, cur_b = 0, iv[0][0], iv[0][1]
for a, b in iv[1:]:
    if a > cur_b:
        union += cur_b - cur_a
        cur_a, cur_b = a, b
    else:
        cur_b = max(cur_b, b)
union += cur_b - cur_a
print(f"depth {depth}: {steps} scenes in {span / 1e3:.1f} ms = {span / 1e3 / steps:.3f} ms/scene; "
      f"GPU busy (union of kernels) {union / span:.3f} of the span; sum of kernel time {total / 1e3 / steps:.3f} ms/scene "
      f"(average concurrency {total / union:.2f})")
