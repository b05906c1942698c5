"""GPU voxeliser with the reference's ``Voxelization`` call contract (``libs/voxel_generator.py:117-154``).

``Voxelization(cfg['voxel_generator'])(points[N,4])`` returns ``{'coordinates' i32[M,4] (z,y,x,t),
'num_voxels' i64[1], 'shape' i64[4] = (nx,ny,nz,nt), 'point_to_voxel_map' i64[N,1]}`` -- device tensors
when given a CUDA tensor, numpy arrays when given numpy (one H2D / D2H round trip, for drop-in use under
``libs/dataset.py:187``).  Pillar ids are bit-exact with the numba kernel (first-touch order).
``voxelize_batch`` is the batched device entry the runner uses: it also applies the running pillar offset
that ``libs/dataloader.py:33-38`` adds at collate time.
"""
import numpy as np
import torch

from ._lib import I, L, P, Z, call, host_floats, scratch, size, stream


class Voxelization:
    def __init__(self, cfg):
        self.voxel_size = np.array(cfg["voxel_size"], dtype=np.float32)
        self.point_cloud_range = np.array(cfg["range"], dtype=np.float32)
        self.n_sweeps = cfg["n_sweeps"]
        grid = (self.point_cloud_range[3:] - self.point_cloud_range[:3]) / self.voxel_size
        self.grid_size = np.round(grid).astype(np.int64)
        self.max_voxels = int(self.grid_size[0] * self.grid_size[1] * self.grid_size[2] * self.n_sweeps)

    def voxelize_batch(self, points4, point_batch=None, batch_size=1):
        """points4: CUDA f32 [N,4]; point_batch: CUDA i32 [N] (scene of each point, ascending) or None."""
        assert points4.is_cuda and points4.dtype == torch.float32
        points4 = points4.contiguous()
        dev = points4.device
        n = points4.shape[0]
        if n == 0:  # nothing to voxelise: no launch, no readback of unwritten counters
            z = lambda *shape: torch.zeros(*shape, dtype=torch.int32, device=dev)
            return {"coordinates": z(0, 4), "pillar_batch": z(0), "point_to_voxel_map": z(0), "num_voxels": z(batch_size),
                    "total_voxels": 0, "n_rejected": 0}
        cells = int(batch_size) * self.max_voxels
        m_cap = min(n, cells)
        coords = torch.empty(m_cap, 4, dtype=torch.int32, device=dev)
        pillar_batch = torch.empty(m_cap, dtype=torch.int32, device=dev)
        p2v = torch.empty(n, dtype=torch.int32, device=dev)
        num_voxels = torch.empty(batch_size, dtype=torch.int32, device=dev)
        total = torch.empty(1, dtype=torch.int32, device=dev)
        ws = scratch(size("pcab_voxelize_workspace", I(n), L(cells)), dev)
        call("pcab_voxelize", P(points4), P(point_batch), I(n), I(batch_size), host_floats(self.point_cloud_range),
             host_floats(self.voxel_size), I(self.n_sweeps), P(coords), P(pillar_batch), P(p2v), P(num_voxels), P(total),
             P(ws), Z(ws.numel()), stream())
        m, n_rejected = torch.stack((total[0], (p2v < 0).sum().to(torch.int32))).tolist()  # one readback for both
        return {"coordinates": coords[:m], "pillar_batch": pillar_batch[:m], "point_to_voxel_map": p2v,
                "num_voxels": num_voxels, "total_voxels": m, "n_rejected": n_rejected}

    def __call__(self, points):
        is_np = isinstance(points, np.ndarray)
        pts = torch.as_tensor(points, dtype=torch.float32)
        if not pts.is_cuda:
            pts = pts.cuda()
        out = self.voxelize_batch(pts)
        res = {
            "coordinates": out["coordinates"],
            "num_voxels": out["num_voxels"].to(torch.int64),
            "shape": torch.tensor(np.hstack((self.grid_size, np.array([self.n_sweeps]))), dtype=torch.int64),
            "point_to_voxel_map": out["point_to_voxel_map"].to(torch.int64)[:, None],
        }
        if is_np:
            res = {k: v.cpu().numpy() for k, v in res.items()}
        return res
