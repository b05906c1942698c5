import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden_forward(name):
    """Rebuild the reference input_dict from a golden fixture (voxelised by the oracle restatement)."""
    from oracle import oracle
    from pcaccumulation_b200 import config as cfgmod
    from pcaccumulation_b200 import synth

    g = np.load(os.path.join(GOLDEN, f"forward_{name}.npz"))
    if name.startswith("waymo"):
        cfg = cfgmod.get_config("waymo")
    else:
        cfg = cfgmod.get_config("nuscene", voxel_generator={"n_sweeps": 10}, data={"n_frames": 10})
    vg = cfg["voxel_generator"]
    pts4 = g["in_points4"]
    v = oracle.voxelize(pts4, vg["voxel_size"], vg["range"], vg["n_sweeps"])
    sample = {
        "input_points": pts4[:, :3].copy(), "num_points": np.array([pts4.shape[0]], dtype=np.int64),
        "time_indice": pts4[:, 3:4].astype(np.int64), "sd_labels": g["in_sd_labels"].astype(np.int64)[:, None],
        "inst_labels": g["in_inst_labels"].astype(np.int64)[:, None], "fb_labels": g["in_fb_labels"].astype(np.int64)[:, None],
        "ego_motion_gt": g["in_ego_motion_gt"], "inst_motion_gt": g["in_inst_motion_gt"],
    }
    sample.update(v)
    return cfg, g, v, synth.collate([sample])


@pytest.fixture(scope="session")
def fixture_weights():
    from pcaccumulation_b200 import fixture
    from pcaccumulation_b200.motionnet import MotionNet

    cache = {}

    def get(cfg):
        key = cfg["voxel_generator"]["n_sweeps"]
        if key not in cache:
            cache[key] = fixture.fixture_state_dict(MotionNet(cfg).state_dict(), 42)
        return cache[key]

    return get
