"""Run warm-up steps, then ONE step of the hot path inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off)."""
import sys

import torch

sys.path.insert(0, ".")
from pcaccumulation_b200 import config, fixture, synth  # noqa: E402
from pcaccumulation_b200.runner import SceneRunner, scene_to_points4  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
no_tc = "--no-tc" in sys.argv
cfg = config.workload_config(name)
runner = SceneRunner(cfg)
runner.model.load_state_dict(fixture.fixture_state_dict(runner.model.state_dict(), 42))
runner.model.use_tensor_cores = not no_tc
runner.warmup()
s = synth.make_workload_scene(name, 0)
p4 = torch.tensor(scene_to_points4(s)).cuda()
for i in range(2):
    torch.manual_seed(i)
    runner.run_device(p4, [p4.shape[0]])
torch.cuda.synchronize()
torch.cuda.profiler.start()
torch.manual_seed(5)
runner.run_device(p4, [p4.shape[0]])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step of", name)
