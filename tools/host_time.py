"""Developer script: host (CPU) time of one forward, with blocking sync so that waiting for the GPU does not count."""
import ctypes, sys, time
cudart = ctypes.CDLL("libcudart.so.12")
print("cudaSetDeviceFlags(blocking sync) ->", cudart.cudaSetDeviceFlags(4))
import torch
sys.path.insert(0, ".")
from pcaccumulation_b200 import config, fixture, synth
from pcaccumulation_b200.runner import SceneRunner, scene_to_points4
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
for i in range(8):
    t0, c0 = time.perf_counter(), time.thread_time()
    torch.manual_seed(5); runner.run_device(p4, [p4.shape[0]]); torch.cuda.synchronize()
    walls.append((time.perf_counter() - t0) * 1e3); cpus.append((time.thread_time() - c0) * 1e3)
print(name, "wall ms", [round(w, 2) for w in walls])
print(name, "thread cpu ms", [round(c, 2) for c in cpus])
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for i in range(5):
    torch.manual_seed(5); runner.run_device(p4, [p4.shape[0]])
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
