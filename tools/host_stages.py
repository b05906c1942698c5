"""Developer script: host CPU time of one forward PER STAGE (thread CPU time between the stage marks, blocking-sync waits so
that waiting for the GPU does not count) + a cProfile of the same.   usage: python tools/host_stages.py [C2]"""
import ctypes
import sys
import time

cudart = ctypes.CDLL("libcudart.so.12")
cudart.cudaSetDeviceFlags(4)  # cudaDeviceScheduleBlockingSync
import torch  # noqa: E402

sys.path.insert(0, ".")
from pcaccumulation_b200 import config, fixture, synth  # noqa: E402
from pcaccumulation_b200.runner import SceneRunner, scene_to_points4  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
torch.set_num_threads(1)
cfg = config.workload_config(name)
runner = SceneRunner(cfg)
runner.model.load_state_dict(fixture.fixture_state_dict(runner.model.state_dict(), 42))
runner.warmup()
s = synth.make_workload_scene(name, 0)
p4 = torch.tensor(scene_to_points4(s)).cuda()
for i in range(4):
    torch.manual_seed(i)
    runner.run_device(p4, [p4.shape[0]])
torch.cuda.synchronize()
marks = []
model = runner.model
model._mark = lambda nm: marks.append((nm, time.thread_time(), time.perf_counter()))
acc = {}
for it in range(10):
    marks.clear()
    c0, w0 = time.thread_time(), time.perf_counter()
    torch.manual_seed(5)
    runner.run_device(p4, [p4.shape[0]])
    marks.append(("end(deferred floats)", time.thread_time(), time.perf_counter()))
    torch.cuda.synchronize()
    pc, pw = c0, w0
    for nm, c, w in marks:
        a = acc.setdefault(nm, [0.0, 0.0])
        a[0] += (c - pc) * 1e3 / 10
        a[1] += (w - pw) * 1e3 / 10
        pc, pw = c, w
print(f"{'stage (time until this mark)':36s} {'cpu ms':>8s} {'wall ms':>8s}")
for nm, (c, w) in acc.items():
    print(f"{nm:36s} {c:8.3f} {w:8.3f}")
print(f"{'total':36s} {sum(v[0] for v in acc.values()):8.3f} {sum(v[1] for v in acc.values()):8.3f}")
import cProfile  # noqa: E402
import pstats  # noqa: E402

del model._mark
pr = cProfile.Profile()
pr.enable()
for i in range(5):
    torch.manual_seed(5)
    runner.run_device(p4, [p4.shape[0]])
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(40)
