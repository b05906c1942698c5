import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import load_golden_forward
from pcaccumulation_b200 import fixture
from pcaccumulation_b200.motionnet import MotionNet
cfg, g, v, inp = load_golden_forward("waymo_small")
model = MotionNet(cfg).cuda().eval()
model.load_state_dict(fixture.fixture_state_dict(model.state_dict(), 42))
model.use_tensor_cores = False
inp_c = {k: (t.cuda() if isinstance(t, torch.Tensor) else t) for k, t in inp.items()}
for rep in range(3):
    model.inject = {"ego_motion_est": torch.tensor(g["out_ego_motion_est"]), "mos_est": torch.tensor(g["out_mos_est"]), "offset_est": torch.tensor(g["out_offset_est"])}
    torch.manual_seed(42)
    c = model(inp_c)
    lab = c["inst_labels_est"].cpu().numpy()
    tp = c["transformed_points"].cpu().numpy()
    print("rep", rep, "label mismatches", int((lab != g["out_inst_labels_est"]).sum()), "n_inst", lab.max(), g["out_inst_labels_est"].max(),
          "tp bit-equal", np.array_equal(tp, g["out_transformed_points"]), "tp maxdiff", np.abs(tp - g["out_transformed_points"]).max())
    bad = np.nonzero(lab != g["out_inst_labels_est"])[0]
    print("   first mismatches", bad[:10], lab[bad[:10]], g["out_inst_labels_est"][bad[:10]])
