"""Chamfer distance with the reference's module contract (``chamfer_distance/chamfer_distance.py:9-57``):
``ChamferDistance()(xyz1[B,n,3], xyz2[B,m,3]) -> (dist1[B,n], dist2[B,m])``, autograd-capable.
Forward and backward run as sm_100a kernels behind ``pcab_chamfer_forward`` / ``pcab_chamfer_backward``."""
import torch

from ._lib import F, I, P, Z, call, scratch, size, stream


class ChamferDistanceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        if not xyz1.is_cuda:
            raise RuntimeError("pcaccumulation_b200.ChamferDistance is CUDA-only (no CPU fallback)")
        xyz1 = xyz1.contiguous().float()
        xyz2 = xyz2.contiguous().float()
        B, n, _ = xyz1.shape
        m = xyz2.shape[1]
        dev = xyz1.device
        dist1 = torch.empty(B, n, device=dev)
        dist2 = torch.empty(B, m, device=dev)
        idx1 = torch.empty(B, n, dtype=torch.int32, device=dev)
        idx2 = torch.empty(B, m, dtype=torch.int32, device=dev)
        ws = scratch(size("pcab_chamfer_workspace", I(B), I(n), I(m)), dev)
        call("pcab_chamfer_forward", P(xyz1), P(xyz2), I(B), I(n), I(m), P(dist1), P(dist2), P(idx1), P(idx2), P(ws),
             Z(ws.numel()), stream())
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, g1, g2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        B, n, _ = xyz1.shape
        m = xyz2.shape[1]
        g1 = g1.contiguous().float()
        g2 = g2.contiguous().float()
        gx1 = torch.empty_like(xyz1)
        gx2 = torch.empty_like(xyz2)
        call("pcab_chamfer_backward", P(xyz1), P(xyz2), I(B), I(n), I(m), P(g1), P(g2), P(idx1), P(idx2), P(gx1), P(gx2),
             stream())
        return gx1, gx2


class ChamferDistance(torch.nn.Module):
    def forward(self, xyz1, xyz2):
        return ChamferDistanceFunction.apply(xyz1, xyz2)


def chamfer_with_indices(xyz1, xyz2, brute=False):
    """(dist1, dist2, idx1, idx2) -- the indices the reference keeps in ctx.saved_tensors.  ``brute`` evaluates every pair
    (the reference kernel's formulation) instead of the exact grid search; the results are bit-identical."""
    xyz1 = xyz1.contiguous().float()
    xyz2 = xyz2.contiguous().float()
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    dist1 = torch.empty(B, n, device=dev)
    dist2 = torch.empty(B, m, device=dev)
    idx1 = torch.empty(B, n, dtype=torch.int32, device=dev)
    idx2 = torch.empty(B, m, dtype=torch.int32, device=dev)
    ws = scratch(size("pcab_chamfer_workspace", I(B), I(n), I(m)), dev)
    call("pcab_chamfer_forward_brute" if brute else "pcab_chamfer_forward", P(xyz1), P(xyz2), I(B), I(n), I(m), P(dist1), P(dist2),
         P(idx1), P(idx2), P(ws), Z(ws.numel()), stream())
    return dist1, dist2, idx1, idx2


def nearest_neighbours(queries, targets, max_dist=0.0):
    """(dist[n], idx[n]) of the nearest target of every query through the exact grid search; ``max_dist`` > 0 bounds the
    search (dist = NaN, idx = -1 where nothing lies strictly within it)."""
    q = queries.contiguous().float()
    t = targets.contiguous().float()
    n, m = q.shape[0], t.shape[0]
    dist = torch.empty(n, device=q.device)
    idx = torch.empty(n, dtype=torch.int32, device=q.device)
    ws = scratch(size("pcab_nn_workspace", I(n), I(m)), q.device)
    call("pcab_nn_search", P(q), I(n), P(t), I(m), F(max_dist), P(dist), P(idx), P(ws), Z(ws.numel()), stream())
    return dist, idx
