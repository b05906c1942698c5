import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import load_golden_forward
from pcaccumulation_b200 import fixture
from pcaccumulation_b200.motionnet import MotionNet
name = sys.argv[1] if len(sys.argv) > 1 else "nuscene_small"
cfg, g, v, inp = load_golden_forward(name)
model = MotionNet(cfg).cuda().eval()
model.load_state_dict(fixture.fixture_state_dict(model.state_dict(), 42))
model.use_tensor_cores = False
inp_c = {k: (t.cuda() if isinstance(t, torch.Tensor) else t) for k, t in inp.items()}
bad_runs = 0
for rep in range(30):
    model.inject = {"ego_motion_est": torch.tensor(g["out_ego_motion_est"]), "mos_est": torch.tensor(g["out_mos_est"]), "offset_est": torch.tensor(g["out_offset_est"])}
    torch.manual_seed(42)
    c = model(inp_c)
    lab = c["inst_labels_est"].cpu().numpy()
    mis = int((lab != g["out_inst_labels_est"]).sum())
    if mis:
        bad_runs += 1
        bad = np.nonzero(lab != g["out_inst_labels_est"])[0]
        print("rep", rep, "label mismatches", mis, "n_inst", lab.max(), g["out_inst_labels_est"].max(), "first", bad[:6], lab[bad[:6]], g["out_inst_labels_est"][bad[:6]])
print(name, "bad runs", bad_runs, "of 30")
