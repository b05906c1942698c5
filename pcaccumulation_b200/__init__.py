"""pcaccumulation_b200: B200 (sm_100a) implementation of the per-scene hot path of prs-eth/PCAccumulation.

Public surface mirrors the reference's: ``MotionNet(cfg)`` / ``model(input_dict) -> results``
(models/motionnet.py), ``Voxelization`` (libs/voxel_generator.py), ``ChamferDistance``
(chamfer_distance/chamfer_distance.py), config dicts with the reference's keys.
"""
from .config import get_config, workload_config  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not require CUDA
    if name == "MotionNet":
        from .motionnet import MotionNet
        return MotionNet
    if name == "Voxelization":
        from .voxel_generator import Voxelization
        return Voxelization
    if name == "ChamferDistance":
        from .chamfer_distance import ChamferDistance
        return ChamferDistance
    raise AttributeError(name)
