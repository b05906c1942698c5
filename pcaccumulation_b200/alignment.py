"""``BaseModel`` of the reference (``models/tpointnet.py:95-163``): frame alignment + Chamfer / L2 alignment errors.

Same method names, argument meaning and weighting; ``align_frames`` runs as ``pcab_transform_points`` (one launch instead
of a boolean mask per frame), the Chamfer distance as the sm_100a kernels behind ``ChamferDistance``."""
import torch

from ._lib import I, P, call, stream
from .chamfer_distance import ChamferDistance

_EPS = 1e-20  # toolbox/utils.py:13


class BaseModel(torch.nn.Module):
    def __init__(self, config):
        super().__init__()
        self.n_frames = config["voxel_generator"]["n_sweeps"]
        self.chamfer_dist = ChamferDistance()

    def align_frames(self, points, time_indice, poses):
        """points [N,3], time_indice [N], poses [n_frames,4,4] -> points of frame t moved by poses[t] (models/tpointnet.py:105-119;
        like upstream, points whose frame index is outside [0, n_frames) stay where they are)."""
        if not points.is_cuda:
            raise RuntimeError("pcaccumulation_b200.alignment.BaseModel is CUDA-only (no CPU fallback)")
        pts = points.float().contiguous()
        t = time_indice.reshape(-1).to(torch.int32)
        eye = torch.eye(4, device=pts.device)
        # frames outside the pose list keep their coordinates: give them an identity slot
        tab = torch.cat((poses.float().reshape(-1, 4, 4)[: self.n_frames], eye[None])).contiguous()
        slot = torch.where((t >= 0) & (t < self.n_frames), t, torch.full_like(t, tab.shape[0] - 1)).contiguous()
        out = torch.empty_like(pts)
        call("pcab_transform_points", P(pts), P(slot), P(tab), I(pts.shape[0]), P(out), stream())
        return out

    def get_chamfer_distance(self, est_points, gt_points, weights):
        """models/tpointnet.py:121-131: (sum(w * d(gt -> est)) + sum(w * d(est -> gt))) / 2."""
        dist1, dist2 = self.chamfer_dist(gt_points[None], est_points[None])
        dist1, dist2 = dist1 * weights, dist2 * weights
        return (dist1.sum() + dist2.sum()) / 2

    def get_l2_distance(self, est_points, gt_points, weights):
        """models/tpointnet.py:133-143."""
        return (torch.norm(est_points - gt_points, dim=1) * weights).sum()

    def get_alignment_errors(self, points, time_indice, est_poses, gt_poses):
        """models/tpointnet.py:145-163: Chamfer and L2 error of the points of frame 1 (weights 1/n_1, 0 elsewhere)."""
        est_points = self.align_frames(points, time_indice, est_poses)
        gt_points = self.align_frames(points, time_indice, gt_poses)
        weights = torch.zeros(est_points.size(0), device=est_points.device)
        weights[time_indice.reshape(-1) == 1] = 1.0
        weights = weights / (weights.sum() + _EPS)
        return self.get_chamfer_distance(est_points, gt_points, weights), self.get_l2_distance(est_points, gt_points, weights)
