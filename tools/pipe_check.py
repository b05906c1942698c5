"""Developer script: throughput of ScenePipeline (device-resident vs host inputs) for several depths, repeated."""
import sys, time
import torch
sys.path.insert(0, ".")
from pcaccumulation_b200 import config, fixture, synth
from pcaccumulation_b200.runner import ScenePipeline, SceneRunner, scene_to_points4
name = "C2"
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = config.workload_config(name)
r0 = SceneRunner(cfg)
sd = fixture.fixture_state_dict(r0.model.state_dict(), 42)
pipe = ScenePipeline(cfg, sd, depth=depth)
scenes = [synth.make_workload_scene(name, i) for i in range(4)]
host_pts = [torch.from_numpy(scene_to_points4(s)).pin_memory() for s in scenes]
dev_pts = [p.cuda() for p in host_pts]
nums = [[p.shape[0]] for p in host_pts]
host_ego = [torch.from_numpy(s["ego_motion_gt"])[None].contiguous().pin_memory() for s in scenes]
dev_ego = [e.cuda() for e in host_ego]
use_ego = "--ego" in sys.argv
if "--threads1" in sys.argv:
    torch.set_num_threads(1)
if "--serial-runner" in sys.argv:
    r0.model.load_state_dict(sd)
    for i in range(3):
        r0.run_device(dev_pts[i], nums[i], ego_motion_gt=dev_ego[i])
torch.cuda.synchronize()
def run(steps, host):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    futs = [pipe.submit((host_pts if host else dev_pts)[i % 4], nums[i % 4], ego=((host_ego if host else dev_ego)[i % 4] if use_ego else None),
                        seed=i, host=host) for i in range(steps)]
    cur = torch.cuda.current_stream()
    for f in futs:
        _, done = f.result()
        cur.wait_event(done)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, (time.perf_counter() - t0) * 1e3 / steps
for rep in range(6):
    for host in (False, True):
        ms, wall = run(24, host)
        print(f"depth {depth} rep {rep} host={host}: {ms:.2f} ms/scene (device events) {wall:.2f} ms/scene (wall)", flush=True)
