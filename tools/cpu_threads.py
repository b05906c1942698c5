import sys, time, os, torch, numpy as np
sys.path.insert(0, ".")
sys.argv = ["x"]
import bench
from pcaccumulation_b200 import config, synth
cfg = config.workload_config("C2")
sd = bench.fixture_weights(cfg)
run = bench.oracle_forward_fn(cfg, sd)
sc = synth.make_workload_scene("C2", 0)
print("cpu_count", os.cpu_count())
for nt in (16, 32, 64):
    torch.set_num_threads(nt)
    t = time.time(); run(sc); print(nt, "threads:", time.time() - t, "s", flush=True)
