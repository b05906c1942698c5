"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's per-scene forward.

This file is the parity oracle: a plain torch-CPU / numpy restatement of the algorithm of
prs-eth/PCAccumulation's hot path, written from the reference's behaviour with the file:line each
function follows.  It is imported ONLY by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``; the product path
(``pcaccumulation_b200``) never imports it and has no CPU fallback.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned
against the UNMODIFIED reference executed in the build container (``oracle/ref_loader.py``):
``tests/test_oracle.py::test_forward_bit_identical_to_reference_when_present`` checks every output of ``forward`` against
``models.motionnet.MotionNet.forward`` on identical inputs and weights, and
``oracle/make_golden.py`` commits reference-generated fixtures under ``tests/golden/`` that the
oracle (and the CUDA path) are checked against where /root/reference does not exist.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

_EPS = 1e-20  # toolbox/utils.py:13
MIN_POINTS = 15  # models/motionnet.py:11


# ------------------------------------------------------------------------------------------------
# voxeliser (libs/voxel_generator.py:4-61, 117-154)
# ------------------------------------------------------------------------------------------------
def voxelize(points, voxel_size, pc_range, n_sweeps):
    """First-touch 4-D pillar assignment.  points f32[N,4] = (x,y,z,t).

    Pillar id = rank of the cell (z,y,x,t) by the stream index of its first point; all coordinate
    arithmetic in float32 like the numba kernel (floor((p - lo) / vs), reject if outside the grid).
    """
    points = np.asarray(points, dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    rng = np.asarray(pc_range, dtype=np.float32)
    grid = np.round((rng[3:] - rng[:3]) / vs).astype(np.int64)  # nx, ny, nz
    c = np.floor((points[:, :3] - rng[:3]) / vs)  # float32
    ok = np.all((c >= 0) & (c < grid.astype(np.float32)), axis=1)
    ci = c.astype(np.int64)
    t = points[:, 3].astype(np.int64)
    nx, ny, nz = int(grid[0]), int(grid[1]), int(grid[2])
    cell = ((ci[:, 2] * ny + ci[:, 1]) * nx + ci[:, 0]) * n_sweeps + t
    N = points.shape[0]
    p2v = -np.ones((N, 1), dtype=np.int64)
    idx_ok = np.nonzero(ok)[0]
    cells_ok = cell[idx_ok]
    uniq, first = np.unique(cells_ok, return_index=True)  # first occurrence (in stream order) per cell
    order = np.argsort(first, kind="stable")  # rank cells by first-touch index
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    pos = np.searchsorted(uniq, cells_ok)
    p2v[idx_ok, 0] = rank[pos]
    first_pt = idx_ok[first[order]]
    coords = np.stack([ci[first_pt, 2], ci[first_pt, 1], ci[first_pt, 0], t[first_pt]], 1).astype(np.int32)
    return {
        "coordinates": coords,
        "num_voxels": np.array([coords.shape[0]], dtype=np.int64),
        "shape": np.hstack((grid, np.array([n_sweeps]))).astype(np.int64),
        "point_to_voxel_map": p2v,
    }


_numba_kernel = None


def voxelize_sequential(points, voxel_size, pc_range, n_sweeps):
    """The same assignment as ``voxelize`` written the way the reference runs it (libs/voxel_generator.py:4-61): ONE sequential
    pass in stream order over a dense cell -> pillar table, jit-compiled with numba like upstream.  Used as the timed CPU
    baseline (the vectorised numpy statement above is ~10x slower than the reference's numba loop); falls back to
    ``voxelize`` when numba is not importable.  tests/test_oracle.py checks that both statements agree exactly."""
    global _numba_kernel
    try:
        import numba
    except Exception:
        return voxelize(points, voxel_size, pc_range, n_sweeps)
    if _numba_kernel is None:
        @numba.njit(cache=False)
        def kernel(points, vs, lo, grid, table, coords, p2v):
            n = points.shape[0]
            m = 0
            for i in range(n):
                ok = True
                c0 = np.int64(0)
                c1 = np.int64(0)
                c2 = np.int64(0)
                for j in range(3):
                    c = np.floor((points[i, j] - lo[j]) / vs[j])  # float32 arithmetic, like the reference
                    if c < 0 or c >= grid[j]:
                        ok = False
                        break
                    if j == 0:
                        c0 = np.int64(c)
                    elif j == 1:
                        c1 = np.int64(c)
                    else:
                        c2 = np.int64(c)
                if not ok:
                    continue
                t = np.int64(points[i, 3])
                idx = table[c2, c1, c0, t]
                if idx == -1:
                    idx = m
                    table[c2, c1, c0, t] = m
                    coords[m, 0] = c2
                    coords[m, 1] = c1
                    coords[m, 2] = c0
                    coords[m, 3] = t
                    m += 1
                p2v[i, 0] = idx
            return m
        _numba_kernel = kernel
    points = np.ascontiguousarray(points, dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    rng = np.asarray(pc_range, dtype=np.float32)
    grid = np.round((rng[3:] - rng[:3]) / vs).astype(np.int64)
    table = -np.ones((int(grid[2]), int(grid[1]), int(grid[0]), n_sweeps), dtype=np.int32)
    coords = np.zeros((points.shape[0], 4), dtype=np.int32)
    p2v = -np.ones((points.shape[0], 1), dtype=np.int64)
    m = _numba_kernel(points, vs, rng[:3].copy(), grid.astype(np.float32), table, coords, p2v)
    return {"coordinates": coords[:m].copy(), "num_voxels": np.array([m], dtype=np.int64),
            "shape": np.hstack((grid, np.array([n_sweeps]))).astype(np.int64), "point_to_voxel_map": p2v}


# ------------------------------------------------------------------------------------------------
# segment reductions (third-party torch_scatter; semantics in oracle/shims/torch_scatter)
# ------------------------------------------------------------------------------------------------
def seg_sum(src, index, n):
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    return torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype).scatter_add_(0, idx, src)


def seg_mean(src, index, n):
    s = seg_sum(src, index, n)
    cnt = torch.zeros(n, dtype=torch.long).scatter_add_(0, index, torch.ones_like(index)).clamp(min=1)
    cnt = cnt.view((-1,) + (1,) * (src.dim() - 1))
    return s / cnt.to(src.dtype) if src.is_floating_point() else torch.div(s, cnt, rounding_mode="floor")


def seg_max(src, index, n):
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype)
    out.scatter_reduce_(0, idx, src, reduce="amax", include_self=False)
    return out


# ------------------------------------------------------------------------------------------------
# small geometry helpers
# ------------------------------------------------------------------------------------------------
def square_distance(src, dst, normalised=False):
    """toolbox/utils.py:125-144."""
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    if normalised:
        dist += 2
    else:
        dist += torch.sum(src ** 2, dim=-1)[:, :, None]
        dist += torch.sum(dst ** 2, dim=-1)[:, None, :]
    return torch.clamp(dist, min=1e-12, max=None)


def kabsch(x1, x2, weights, eps=1e-7):
    """toolbox/register_utils.py:247-318 (weighted Kabsch; returns R[b,3,3], t[b,3,1])."""
    weights = weights / (torch.sum(weights, dim=1, keepdim=True) + eps)
    w = weights.unsqueeze(2)
    x1_mean = torch.matmul(w.transpose(1, 2), x1) / (torch.sum(w, dim=1).unsqueeze(1) + eps)
    x2_mean = torch.matmul(w.transpose(1, 2), x2) / (torch.sum(w, dim=1).unsqueeze(1) + eps)
    x1c, x2c = x1 - x1_mean, x2 - x2_mean
    cov = torch.matmul(x1c.transpose(1, 2), torch.matmul(torch.diag_embed(w.squeeze(2)), x2c))
    try:
        u, s, v = torch.svd(cov)
    except Exception:  # SVD failure -> identity (register_utils.py:295-304)
        b = x1.shape[0]
        return torch.eye(3).repeat(b, 1, 1), torch.zeros(b, 3, 1)
    det = torch.det(torch.matmul(v.transpose(1, 2), u.transpose(1, 2)))
    dm = torch.diag_embed(torch.cat((torch.ones(det.shape[0], 2), det.unsqueeze(1)), 1))
    rot = torch.matmul(v, torch.matmul(dm, u.transpose(1, 2)))
    trans = x2_mean.transpose(1, 2) - torch.matmul(rot, x1_mean.transpose(1, 2))
    return rot, trans


def icp_point_to_point(src, tgt, max_dist, init=None, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6):
    """open3d.pipelines.registration.registration_icp(src, tgt, max_dist, init, TransformationEstimationPointToPoint(),
    ICPConvergenceCriteria(max_iteration=max_iter)) as the reference calls it (models/egomotion.py:9-28, models/alignnet.py:
    79-83).  Open3D (unpinned by the reference, README "open3d") is NOT installed here and not vendored: PARITY UNPINNED for
    this function -- it restates the published algorithm of Open3D's RegistrationICP (cpp/open3d/pipelines/registration/
    Registration.cpp): correspondences = nearest target strictly within max_dist (KD-tree hybrid search, max_nn = 1);
    fitness = #correspondences / #source, inlier_rmse = sqrt(mean squared distance); per iteration the update is
    Eigen::umeyama(source, target, with_scaling=false) over the correspondence set (identity when it is empty), applied to
    the source in float64; stop when both |d fitness| < 1e-6 and |d rmse| < 1e-6.  Returns (T[4,4] float64, fitness, rmse)."""
    from scipy.spatial import cKDTree

    src = np.asarray(src, dtype=np.float64).reshape(-1, 3)
    tgt = np.asarray(tgt, dtype=np.float64).reshape(-1, 3)
    T = np.eye(4) if init is None else np.asarray(init, dtype=np.float64).copy()
    if len(src) == 0 or len(tgt) == 0:
        return T, 0.0, 0.0
    tree = cKDTree(tgt)
    pts = src @ T[:3, :3].T + T[:3, 3]

    def correspondences(p):
        d, j = tree.query(p, k=1)
        ok = d < max_dist
        n = int(ok.sum())
        return ok, j, (n / len(p), float(np.sqrt((d[ok] ** 2).sum() / n)) if n else 0.0)

    def umeyama(a, b):
        ma, mb = a.mean(0), b.mean(0)
        cov = (b - mb).T @ (a - ma) / len(a)
        U, _, Vt = np.linalg.svd(cov)
        S = np.eye(3)
        if np.linalg.det(U) * np.linalg.det(Vt) < 0:
            S[2, 2] = -1
        R = U @ S @ Vt
        out = np.eye(4)
        out[:3, :3], out[:3, 3] = R, mb - R @ ma
        return out

    ok, j, (fit, rmse) = correspondences(pts)
    for _ in range(max_iter):
        upd = umeyama(pts[ok], tgt[j[ok]]) if ok.any() else np.eye(4)
        T = upd @ T
        pts = pts @ upd[:3, :3].T + upd[:3, 3]
        ok, j, (fit2, rmse2) = correspondences(pts)
        done = abs(fit - fit2) < rel_fitness and abs(rmse - rmse2) < rel_rmse
        fit, rmse = fit2, rmse2
        if done:
            break
    return T, fit, rmse


def relative_pose(tsfm_src, tsfm_tgt):
    """toolbox/register_utils.py:184-197 (waymo / nuscene branch): inv(T_tgt) @ T_src."""
    return torch.linalg.solve(tsfm_tgt, tsfm_src)


def rotation_error(R1, R2):
    """toolbox/register_utils.py:19-42 (degrees)."""
    R_ = torch.matmul(R1.transpose(1, 2), R2)
    e = torch.stack([(torch.trace(R_[i]) - 1) / 2 for i in range(R_.shape[0])], dim=0).unsqueeze(1)
    e = torch.clamp(e, -1, 1)
    return 180.0 * torch.acos(e) / torch.tensor([math.pi]).type(e.dtype)


def reconstruct_sequence(points, time_indice, inst_labels, tsfm, n_frames):
    """toolbox/register_utils.py:72-93."""
    tsfm = tsfm.reshape(-1, 4, 4)
    indice = (inst_labels.long() * n_frames + time_indice).long()
    pt = tsfm[indice]
    return (torch.matmul(pt[:, :3, :3], points[:, :, None]) + pt[:, :3, 3][:, :, None]).squeeze(-1)


def quat2mat(q):
    """toolbox/se3_utils.py:44-64, quaternion layout (x, y, z, w)."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz = w * x, w * y, w * z
    xy, xz, yz = x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).reshape(-1, 3, 3)


def ravel_hash(x):
    """dataset_toolbox/prep_nuscene_waymo_sf/libs/spv_utils.py:7-20 (torchsparse 1.4 ravel hash)."""
    x = x - np.min(x, axis=0)
    x = x.astype(np.uint64, copy=False)
    xmax = np.max(x, axis=0).astype(np.uint64) + 1
    h = np.zeros(x.shape[0], dtype=np.uint64)
    for k in range(x.shape[1] - 1):
        h += x[:, k]
        h *= xmax[k + 1]
    h += x[:, -1]
    return h


def chamfer(xyz1, xyz2, chunk=2048):
    """chamfer_distance/chamfer_distance.cpp:59-111: squared NN distance + lowest-index argmin, both ways.

    float32 arithmetic in the reference's order ((dx*dx + dy*dy) + dz*dz), strict '<' so the lowest
    index wins ties.
    """
    def one_way(a, b):
        n = a.shape[0]
        dist = np.empty(n, np.float32)
        idx = np.empty(n, np.int32)
        for s in range(0, n, chunk):
            q = a[s:s + chunk]
            dx = b[None, :, 0] - q[:, None, 0]
            dy = b[None, :, 1] - q[:, None, 1]
            dz = b[None, :, 2] - q[:, None, 2]
            d = (dx * dx + dy * dy) + dz * dz
            k = np.argmin(d, axis=1)  # first minimum
            idx[s:s + chunk] = k
            dist[s:s + chunk] = d[np.arange(q.shape[0]), k]
        return dist, idx

    out1, out2, i1, i2 = [], [], [], []
    for b in range(xyz1.shape[0]):
        a, c = np.asarray(xyz1[b], np.float32), np.asarray(xyz2[b], np.float32)
        d1, k1 = one_way(a, c)
        d2, k2 = one_way(c, a)
        out1.append(d1), out2.append(d2), i1.append(k1), i2.append(k2)
    return np.stack(out1), np.stack(out2), np.stack(i1), np.stack(i2)


# ------------------------------------------------------------------------------------------------
# the model
# ------------------------------------------------------------------------------------------------
class OracleMotionNet:
    """Functional restatement of models/motionnet.py:MotionNet with weights from a state_dict."""

    def __init__(self, cfg, state_dict, dtype=torch.float32, inject=None):
        """dtype=torch.float32 is the reference's arithmetic (the parity target).  dtype=torch.float64 runs the SAME
        restatement in double precision: the "truth" against which the tests measure the FP32 rounding noise of the
        reference itself, stage by stage (tests/test_gpu_parity.py::_check_protocol).  ``inject`` (same keys as
        ``pcaccumulation_b200.MotionNet.inject``: fb_est_map, ego_motion_est, mos_est, offset_est, transformed_points,
        inst_labels_est)
        replaces a computed tensor for the stages DOWNSTREAM of it, so that a float64 run follows the discrete decisions
        (labels, keypoint draws) of the float32 run it is compared with."""
        self.cfg = cfg
        self.dt = dtype
        self.inject = dict(inject or {})
        self.w = {k: v.detach().clone().to(dtype) if v.is_floating_point() else v.clone() for k, v in state_dict.items()}
        vg = cfg["voxel_generator"]
        self.pc_range = vg["range"]
        self.resolution = vg["voxel_size"]
        self.n_sweeps = vg["n_sweeps"]
        self.mode = cfg["misc"]["mode"]
        self.stages = {}

    # --- building blocks --------------------------------------------------------------------
    def lin(self, x, name):
        return F.linear(x, self.w[name + ".weight"], self.w.get(name + ".bias"))

    def conv(self, x, name):
        return F.conv2d(x, self.w[name + ".weight"], self.w[name + ".bias"], padding=1)

    def bn(self, x, name):
        return F.batch_norm(x, self.w[name + ".running_mean"], self.w[name + ".running_var"],
                            self.w[name + ".weight"], self.w[name + ".bias"], False, 0.0, 1e-5)

    def resblock(self, x, name):
        """models/pillar_encoder.py:46-55 (pre-activation, bias-free shortcut)."""
        net = self.lin(F.relu(x), name + ".fc_0")
        dx = self.lin(F.relu(net), name + ".fc_1")
        return F.linear(x, self.w[name + ".shortcut.weight"]) + dx

    def down(self, x, name, pool):
        """models/unet.py:64-71."""
        x = F.relu(self.conv(x, name + ".conv1"))
        x = F.relu(self.conv(x, name + ".conv2"))
        return (F.max_pool2d(x, 2, 2) if pool else x), x

    def up(self, skip, x, name):
        """models/unet.py:100-113."""
        x = F.conv_transpose2d(x, self.w[name + ".upconv.weight"], self.w[name + ".upconv.bias"], stride=2)
        x = torch.cat((x, skip), 1)
        x = F.relu(self.conv(x, name + ".conv1"))
        return F.relu(self.conv(x, name + ".conv2"))

    def unet(self, x, prefix, depth, final):
        """models/unet.py:222-233 / models/stpn.py:83-89."""
        enc = []
        for i in range(depth):
            x, before = self.down(x, f"{prefix}down_convs.{i}", i < depth - 1)
            enc.append(before)
        for i in range(depth - 1):
            x = self.up(enc[-(i + 2)], x, f"{prefix}up_convs.{i}")
        return self.conv(x, prefix + "conv_final") if final else x

    def seghead2d(self, x, name):
        """models/unet.py:264-277."""
        x = self.conv(x, name + ".seg_head.0")
        x = F.relu(self.bn(x, name + ".seg_head.1"))
        return self.conv(x, name + ".seg_head.3")

    def seghead1d(self, x, name):
        """models/unet.py:240-256."""
        x = self.lin(x, name + ".seg_head.0")
        x = F.relu(self.bn(x, name + ".seg_head.1"))
        return self.lin(x, name + ".seg_head.3")

    # --- canvas scatter / gather (models/pillar_encoder.py:125-204) ------------------------------
    def flat_index(self, coords):
        nx, ny = self.Nx, self.Ny
        return (coords[:, 4] * nx * ny + coords[:, 2] * nx + coords[:, 3]).long()

    def scatter_canvas(self, feats, coords, B):
        C = feats.size(1)
        canvas = torch.zeros(B, C, self.nt * self.Nx * self.Ny, dtype=feats.dtype)
        flat = self.flat_index(coords)
        b = coords[:, 0].long()
        canvas[b, :, flat] = feats
        return canvas.view(B, C, self.nt, self.Ny, self.Nx)

    def gather_canvas(self, canvas, coords):
        B, C = canvas.shape[:2]
        flat = self.flat_index(coords)
        return canvas.reshape(B, C, -1)[coords[:, 0].long(), :, flat]

    # --- bilinear point sampling (models/pillar_encoder.py:206-267) -----------------------------
    def ungrid(self, feats, points, time_indice):
        """feats [B,C,H,W]; border padding, align_corners=False; one sample per point."""
        out = torch.zeros(points.size(0), feats.size(1))
        u = points[:, 0] / abs(self.pc_range[0])
        v = points[:, 1] / abs(self.pc_range[1])
        for b in range(feats.size(0)):
            sel = time_indice[:, 0] == b
            if sel.sum():
                grid = torch.stack((u[sel], v[sel]), 1)[None, None]  # [1,1,K,2]
                s = F.grid_sample(feats[b:b + 1], grid, mode="bilinear", padding_mode="border", align_corners=False)
                out[sel] = s[0, :, 0].t()
        return out

    def temporal_ungrid(self, feats, points, time_indice):
        out = torch.zeros(points.size(0), feats.size(2))
        for t in range(feats.size(1)):
            sel = time_indice[:, 1] == t
            if sel.sum():
                out[sel] = self.ungrid(feats[:, t], points[sel], time_indice[sel])
        return out

    # --- pillar encoder (models/pillar_encoder.py:97-122) ---------------------------------------
    def pillar_encoder(self, pts, p2v, coords, pillar_mean, time_indice, M):
        scale = abs(self.pc_range[0])
        vx, vy = self.resolution[0], self.resolution[1]
        x_off, y_off = vx / 2 + self.pc_range[0], vy / 2 + self.pc_range[1]
        d_mean = pts - pillar_mean[p2v]
        mc = coords[p2v]
        f_center = torch.zeros_like(pts[:, :2])
        f_center[:, 0] = pts[:, 0] - (mc[:, 3] * vx + x_off)
        f_center[:, 1] = pts[:, 1] - (mc[:, 2] * vy + y_off)
        feats = torch.cat([pts, d_mean, f_center, time_indice[:, 1:2]], dim=-1).to(self.dt)
        feats[:, :-1] /= scale
        feats[:, -1] /= self.n_sweeps
        net = self.lin(feats, "pillar_encoder.fc_pos")
        net = self.resblock(net, "pillar_encoder.blocks.0")
        depth = self.cfg["pillar_encoder"]["depth"]
        for i in range(1, depth):
            pooled = seg_max(net, p2v, M)[p2v]
            net = self.resblock(torch.cat([net, pooled], 1), f"pillar_encoder.blocks.{i}")
        return seg_max(self.lin(net, "pillar_encoder.fc_c"), p2v, M)

    # --- ego motion (models/egomotion.py) -------------------------------------------------------
    def sinkhorn(self, log_alpha, n_iters):
        """models/egomotion.py:100-137 (slack row/column, log domain)."""
        la = F.pad(log_alpha, (0, 1, 0, 1))
        for _ in range(n_iters):
            la = torch.cat((la[:, :-1, :] - torch.logsumexp(la[:, :-1, :], dim=2, keepdim=True), la[:, -1, None, :]), dim=1)
            la = torch.cat((la[:, :, :-1] - torch.logsumexp(la[:, :, :-1], dim=1, keepdim=True), la[:, :, -1, None]), dim=2)
        return la[:, :-1, :-1]

    def sample_indices(self, n):
        """models/egomotion.py:155-166: host randperm when n > n_kpts, else pad with the last index."""
        k = self.cfg["pose_estimation"]["n_kpts"]
        if n > k:
            return torch.randperm(n)[:k]
        c = torch.arange(k)
        c[n:] = n - 1
        return c

    def pairwise(self, feats_s, feats_t, coor_s, coor_t, duration):
        """models/egomotion.py:140-192."""
        pe = self.cfg["pose_estimation"]
        cs = self.sample_indices(feats_s.size(0))
        ct = self.sample_indices(feats_t.size(0))
        fs, xs = feats_s[cs][None], coor_s[cs][None]
        ft, xt = feats_t[ct][None], coor_t[ct][None]
        thr = duration * self.cfg["data"]["max_speed"]
        support = (square_distance(xs, xt) < thr ** 2).to(self.dt)
        feat_dist = square_distance(fs, ft, normalised=True)
        alpha, beta = self.w["ego_motion_head.alpha"], self.w["ego_motion_head.beta"]
        affinity = -(feat_dist - F.softplus(alpha)) / (torch.exp(beta) + 0.02)
        perm = torch.exp(self.sinkhorn(affinity, pe["sinkhorn_iter"])) * support
        weighted_t = perm @ xt / (torch.sum(perm, dim=2, keepdim=True) + _EPS)
        R, t = kabsch(xs, weighted_t, torch.sum(perm, dim=2))
        pose = torch.eye(4)
        pose[:3, :3] = R[0]
        pose[:3, 3] = t[0][:, 0]
        return pose, perm

    def ego_motion(self, geo, fb_est, occ_map, pts_mean_map, ego_gt, results, raw=None):
        """models/egomotion.py:387-469 with the sequence strategies of :195-357."""
        B, T, C, Ny, Nx = geo.shape
        freq = self.cfg["data"]["freq"]
        mode = self.cfg["pose_estimation"]["seq_pose"]
        perm_list, chained_est, chained_gt = [], [], []
        tot_l1 = tot_l2 = 0
        count = 0
        eye = torch.eye(4)
        for b in range(B):
            gt = ego_gt[b]
            pts, feats, bg = [], [], []
            for t in range(T):
                occ = occ_map[b, t, 0].reshape(-1) > 0
                pts.append(pts_mean_map[b, t].permute(1, 2, 0).reshape(Ny * Nx, 3)[occ])
                feats.append(geo[b, t].permute(1, 2, 0).reshape(Ny * Nx, C)[occ])
                bg.append((fb_est[b, t, 0].reshape(-1) == 0)[occ])
            chained_est.append(eye)
            chained_gt.append(eye)
            if mode == "skip":
                pairs = [(0, t, t / freq) for t in range(1, T)]
            elif mode == "chain":
                pairs = [(t - 1, t, 1.0 / freq) for t in range(1, T)]
            else:
                pairs = [(a, a + gap, gap / freq) for gap in range(1, T) for a in range(T - 1) if a + gap < T]
            chain = eye
            for anchor, ref, duration in pairs:
                pose, perm = self.pairwise(feats[ref][bg[ref]], feats[anchor][bg[anchor]],
                                           pts[ref][bg[ref]], pts[anchor][bg[anchor]], duration)
                pose_gt = relative_pose(gt[ref], gt[anchor])
                ph = torch.cat([pts[ref], torch.ones(pts[ref].size(0), 1)], dim=1)
                pc_est, pc_gt = (pose @ ph.T).T[:, :3], (pose_gt @ ph.T).T[:, :3]
                tot_l1 = tot_l1 + torch.norm(pc_est - pc_gt, p=1, dim=1).mean()
                tot_l2 = tot_l2 + torch.norm(pc_est - pc_gt, p=2, dim=1).mean()
                count += 1
                if mode == "chain":
                    chain = chain @ pose
                    perm_list.append(perm)
                    chained_est.append(chain)
                    chained_gt.append(relative_pose(gt[ref], gt[0]))
                elif anchor == 0:
                    perm_list.append(perm)
                    chained_est.append(pose)
                    chained_gt.append(pose_gt)
            if self.cfg["model"]["ego_icp"]:  # models/egomotion.py:360-384,439-441: raw background points, frame 0 = anchor
                points, time_indice, fb_pp = raw
                pe = self.cfg["pose_estimation"]
                sel0 = (time_indice[:, 0] == b) & (time_indice[:, 1] == 0) & (fb_pp[:, 0] == 0)
                anchor_pts = points[sel0].double().numpy()
                refined = [eye]
                for t in range(1, T):
                    sel = (time_indice[:, 0] == b) & (time_indice[:, 1] == t) & (fb_pp[:, 0] == 0)
                    init = chained_est[-T:][t].float().numpy()
                    Tm, _, _ = icp_point_to_point(points[sel].double().numpy(), anchor_pts, pe["icp_threshold"], init, pe["icp_max_iter"])
                    refined.append(torch.tensor(Tm).float())
                chained_est[-T:] = refined
        est, gtp = torch.stack(chained_est), torch.stack(chained_gt)
        rot_err = rotation_error(est[:, :3, :3], gtp[:, :3, :3]).mean().item()
        trans_err = torch.norm(est[:, :3, 3].unsqueeze(-1) - gtp[:, :3, 3].unsqueeze(-1), dim=(1, 2)).mean().item()
        n = self.n_sweeps
        results["ego_l1_loss"] = tot_l1 / count
        results["ego_l2_loss"] = tot_l2 / count
        results["ego_rot_error"] = rot_err * n / (n - 1)
        results["ego_trans_error"] = trans_err * n / (n - 1)
        results["perm_matrix"] = perm_list
        results["ego_motion_est"] = est.view(B, T, 4, 4)
        results["ego_motion_gt"] = gtp.view(B, T, 4, 4)

    # --- feature warp (models/motionnet.py:45-114) ----------------------------------------------
    def warp_feats(self, bev, pose):
        B, T, C, H, W = bev.shape
        x_min, y_min = self.pc_range[0], self.pc_range[1]
        out = []
        for b in range(B):
            grids = []
            for t in range(1, T):
                inv = torch.linalg.inv(pose[b, t])
                xx = (torch.arange(0, W).view(1, -1).repeat(H, 1) + 0.5).to(self.dt) * self.resolution[0] + x_min
                yy = (torch.arange(0, H).view(-1, 1).repeat(1, W) + 0.5).to(self.dt) * self.resolution[1] + y_min
                g = torch.stack((xx.reshape(-1), yy.reshape(-1)), 0)
                tg = inv[:2, :2] @ g + inv[:2, 3:4]
                tg[0] = tg[0] / abs(x_min)
                tg[1] = tg[1] / abs(y_min)
                grids.append(tg.view(2, H, W))
            grids = torch.stack(grids).permute(0, 2, 3, 1)
            sampled = F.grid_sample(bev[b, 1:], grids, mode="bilinear", padding_mode="zeros", align_corners=False)
            # quirk Q1: slot 0 is the LAST frame, unwarped (leaked loop variable, motionnet.py:111)
            out.append(torch.cat((bev[b, T - 1:T], sampled), dim=0))
        return torch.stack(out)

    def transform_points(self, points, time_indice, tsfm):
        """models/motionnet.py:117-135."""
        B, T = tsfm.shape[:2]
        out = torch.ones_like(points)
        for b in range(B):
            for t in range(T):
                sel = (time_indice[:, 0] == b) & (time_indice[:, 1] == t)
                out[sel] = ((tsfm[b, t, :3, :3] @ points[sel].T) + tsfm[b, t, :3, 3:4]).T
        return out

    # --- STPN (models/stpn.py:67-104) -----------------------------------------------------------
    def stpn(self, x, points, time_indice):
        for i in (0, 2, 4, 6):
            x = F.relu(F.conv3d(x, self.w[f"motionhead.init_conv.{i}.weight"], self.w[f"motionhead.init_conv.{i}.bias"], padding=1))
        self.stages["stpn_conv3d"] = x
        x = torch.max(x, dim=2)[0]
        x = self.unet(x, "motionhead.", 5, final=False)
        ung = self.ungrid(x, points, time_indice)
        pos = points / abs(self.pc_range[0])
        pos = F.relu(self.lin(F.relu(self.lin(pos, "motionhead.positional_encoding.0")), "motionhead.positional_encoding.2"))
        enc = F.relu(self.lin(torch.cat([pos, ung], dim=-1), "motionhead.final_proj.0"))
        mos = self.seghead1d(enc, "motionhead.mos_seg")
        off = self.seghead1d(enc, "motionhead.offset_head")
        off = torch.where(torch.isnan(off), torch.zeros_like(off), off)
        off = torch.where(torch.isinf(off), torch.zeros_like(off), off)
        return mos, torch.clamp(off, -20, 20), x

    # --- clustering (models/cluster.py) ---------------------------------------------------------
    def cluster(self, tp, mos, offset, time_indice):
        from sklearn.cluster import DBSCAN

        cc = self.cfg["cluster"]
        min_p = cc["min_p_cluster"]
        est = DBSCAN(min_samples=cc["min_samples_dbscan"], metric=cc["cluster_metric"], eps=cc["eps_dbscan"])
        B = int(time_indice[:, 0].max() + 1)
        outs = []
        for b in range(B):
            selb = time_indice[:, 0] == b
            if not selb.sum():
                continue
            m, o, p = mos[selb], offset[selb], tp[selb].clone()
            full = torch.zeros(m.size(0)).long()
            sel = m == 1
            if sel.sum() > min_p:
                q = p.clone()
                q[:, :2] += o
                q = q[sel].numpy()
                coords = np.floor(np.round(q / 0.05) / 1).astype(np.int32)  # cluster.py:9-11 + sparse_quantize
                _, sub, inv = np.unique(ravel_hash(coords), return_index=True, return_inverse=True)
                q[:, -1] = 0
                lab = est.fit_predict(q[sub])
                for u in np.unique(lab).tolist():  # cluster.py:36-41
                    if (lab == u).sum() < min_p:
                        lab[lab == u] = -1
                uniq = sorted(set(lab.tolist()))  # toolbox/utils.py:237-250
                remap = {e: i for i, e in enumerate(uniq)}
                canon = np.array([remap[e] for e in lab.tolist()])
                if lab.min() != -1:
                    canon = canon + 1
                full[sel] = torch.from_numpy(canon[inv]).long()
            outs.append(full)
        return torch.cat(outs).long()

    # --- TubeNet (models/alignnet.py:166-285, models/tpointnet.py:211-305) ----------------------
    def mlp3(self, x, name):
        x = F.relu(self.lin(x, name + ".0"))
        x = F.relu(self.lin(x, name + ".2"))
        return self.lin(x, name + ".4")

    def tpointnet(self, mos_feat, frame_feats, points, time_indice, inst_indice, mos_labels, inst_motion_gt):
        p = "reconstructor.alignment."
        K, T = inst_motion_gt.shape[:2]
        frame_indice = (inst_indice * T + time_indice).long()
        count = torch.ones(frame_indice.size(0))
        frame_count = seg_sum(count, frame_indice, K * T)
        frame_weights = (frame_count > self.cfg["tpointnet"]["min_points"]).to(self.dt)
        inst_mos = seg_max(mos_labels, frame_indice, K * T)
        mos_w = torch.ones_like(inst_mos)
        mos_w[inst_mos == 0] = 0.2
        temporal_w = (torch.arange(self.n_sweeps) + 1).repeat(K) / self.n_sweeps
        frame_weights = frame_weights * mos_w * temporal_w

        mos_emb = seg_max(self.mlp3(mos_feat, p + "motion_embed"), inst_indice, K)
        geo_emb = seg_max(self.mlp3(frame_feats, p + "geo_embed"), inst_indice, K)
        frame_centroid = seg_mean(points, frame_indice, K * T)
        inst_centroid = frame_centroid[::T]
        centered = points - inst_centroid[inst_indice]
        frame_in = torch.cat((centered, time_indice.unsqueeze(-1) / T), dim=1).to(self.dt)
        frame_emb = seg_max(self.mlp3(frame_in, p + "pos_embed"), frame_indice, K * T)
        anchor = frame_emb[::T].repeat_interleave(T, 0)
        reg_in = torch.cat((geo_emb.repeat_interleave(T, 0), mos_emb.repeat_interleave(T, 0), frame_emb, anchor), dim=1)
        x = F.relu(self.bn(self.lin(reg_in, p + "regressor.0"), p + "regressor.1"))
        x = F.relu(self.bn(self.lin(x, p + "regressor.3"), p + "regressor.4"))
        rep = self.lin(x, p + "regressor.6")  # [K*T, 7] = (quat xyzw, trans)
        quat = F.normalize(rep[:, :4], p=2, dim=1)
        tsfm = torch.eye(4)[None].repeat(rep.size(0), 1, 1)
        tsfm[:, :3, :3] = quat2mat(quat)
        tsfm[:, :3, 3] = rep[:, 4:]

        # losses against the (centred) GT poses; tpointnet.py:43-73,275-288
        gt = inst_motion_gt.clone().view(-1, 4, 4)
        cen = inst_centroid.repeat_interleave(T, 0).unsqueeze(2)
        gt[:, :3, 3] += torch.matmul(gt[:, :3, :3] - torch.eye(3)[None], cen).squeeze(2)
        from scipy.spatial.transform import Rotation

        gt_quat = torch.from_numpy(Rotation.from_matrix(gt[:, :3, :3].numpy()).as_quat()).to(self.dt)
        gt_rep = torch.cat((gt_quat, gt[:, :3, 3]), 1)
        rec_est = reconstruct_sequence(centered, time_indice, inst_indice, tsfm.view(K, T, 4, 4), T)
        rec_gt = reconstruct_sequence(centered, time_indice, inst_indice, gt.view(K, T, 4, 4), T)
        diff = rec_est - rec_gt
        f_l1 = seg_mean(torch.norm(diff, p=2, dim=1), frame_indice, K * T)  # names swapped upstream (Q6)
        f_l2 = seg_mean(torch.norm(diff, p=1, dim=1), frame_indice, K * T)
        wsum = frame_weights.sum() + _EPS
        l1 = (f_l1 * frame_weights).sum() / wsum
        l2 = (f_l2 * frame_weights).sum() / wsum
        rot_loss = (torch.norm(gt_rep[:, :4] - quat, p=2, dim=1) * frame_weights).sum() / wsum
        trans_loss = (torch.norm(gt_rep[:, 4:] - rep[:, 4:], p=2, dim=1) * frame_weights).sum() / wsum

        tsfm[:, :3, 3] += torch.matmul(torch.eye(3)[None] - tsfm[:, :3, :3], cen).squeeze(2)
        tsfm = tsfm.view(K, T, 4, 4)
        tsfm[:, 0] = torch.eye(4)[None]
        return {"l1_loss": l1, "l2_loss": l2, "rot_loss": rot_loss, "trans_loss": trans_loss, "inst_est_motion": tsfm}

    def alignnet(self, inp, results):
        T = self.n_sweeps
        mos_labels = inp["mos_labels"]
        inst_labels = inp["inst_labels"].clone()
        time_indice = inp["time_indice"]
        tp = inp["transformed_points"].clone()
        n_points = inst_labels.size(0)
        ego_est, ego_gt = inp["ego_motion_est"], inp["ego_motion_gt"]
        if self.mode == "test":
            n_inst = int(inst_labels.max()) + 1
            inst_motion_gt = [torch.eye(4)[None, None].repeat(n_inst, T, 1, 1)]
        else:
            inst_motion_gt = [m.to(self.dt) for m in inp["inst_motion_gt"]]
        # alignnet.py:9-38: compensate the GT by the ego-pose error
        upd = []
        for b, m in enumerate(inst_motion_gt):
            K = m.size(0)
            g = ego_gt[b][None].repeat(K, 1, 1, 1).view(-1, 4, 4)
            e = ego_est[b][None].repeat(K, 1, 1, 1).view(-1, 4, 4)
            upd.append((m.view(-1, 4, 4) @ g @ torch.linalg.inv(e)).view(K, -1, 4, 4))
        run = 0
        for b in range(len(upd)):
            sel = time_indice[:, 0] == b
            if sel.sum():
                inst_labels[sel] += run
                run += upd[b].size(0)
        motion = torch.cat(upd)
        # alignnet.py:115-163: padding + relabel
        K = motion.size(0)
        t_idx = time_indice[:, 1]
        frame_indice = (inst_labels * T + t_idx).long()
        ones = torch.ones(frame_indice.size(0))
        frame_count = seg_sum(ones, frame_indice, K * T)
        inst_count = seg_sum(ones, inst_labels, K)
        anchor_count = frame_count[::T]
        pad = []
        for k in torch.where((anchor_count == 0) & (inst_count > 0))[0].tolist():
            c = frame_count[k * T:(k + 1) * T]
            f = k * T + torch.where(c > 0)[0][0]
            pad.append(torch.where(frame_indice == f)[0])
        keep = inst_count > 0
        motion = motion[keep]
        mapping = -torch.ones(K).long()
        mapping[keep] = torch.arange(int(keep.sum()))
        inst_labels = mapping[inst_labels]
        inst_motion_gt = motion.clone()
        K = motion.size(0)
        if pad:
            pad = torch.cat(pad)
            p_time = torch.cat((t_idx, torch.zeros_like(pad).long()))
            p_idx = torch.cat((torch.arange(n_points).long(), pad))
        else:
            p_time, p_idx = t_idx, torch.arange(n_points).long()
        p_bb, p_mf = inp["backbone_feats"][p_idx], inp["motion_feats"][p_idx]
        p_inst, p_mos, p_pts = inst_labels[p_idx], mos_labels[p_idx], tp[p_idx]
        p_pts0 = p_pts.clone()
        results["tpointnet_loss_terms"] = {}
        final = None
        for it in range(self.cfg["tpointnet"]["n_iterations"]):
            pred = self.tpointnet(p_mf, p_bb, p_pts, p_time, p_inst, p_mos, motion)
            results["tpointnet_loss_terms"][f"{it}_th"] = pred
            c = pred["inst_est_motion"]
            p_pts = reconstruct_sequence(p_pts, p_time, p_inst, c, T)
            motion = motion.view(-1, 4, 4)
            c = c.reshape(-1, 4, 4)
            motion[:, :3, :3] = torch.matmul(motion[:, :3, :3], c[:, :3, :3].transpose(1, 2))
            motion[:, :3, 3] = motion[:, :3, 3] - torch.matmul(motion[:, :3, :3], c[:, :3, 3].unsqueeze(-1)).squeeze(-1)
            motion = motion.view(K, T, 4, 4)
            final = c if final is None else torch.matmul(c, final)
        final = final.view(K, T, 4, 4)
        if self.cfg["model"]["tpointnet_icp"]:  # models/alignnet.py:54-112,264-266
            rec = reconstruct_sequence(p_pts0, p_time, p_inst, final, T)
            thr = self.cfg["tpointnet"]["icp_threshold"]
            final = final.clone()
            for k in range(K):
                sel = p_inst == k
                pk, tk = rec[sel].double().numpy(), p_time[sel].numpy()
                assert tk.min() == 0
                anchor_pts = pk[tk == 0]
                ref = []
                for t in range(T):
                    cur = pk[tk == t]
                    ref.append(icp_point_to_point(cur, anchor_pts, thr, None, 50)[0] if (t != 0 and len(cur)) else np.eye(4))
                final[k] = torch.matmul(torch.tensor(np.array(ref)).to(final.dtype), final[k])
        rec_est = reconstruct_sequence(inp["transformed_points"], t_idx, inst_labels, final, T)
        rec_gt = reconstruct_sequence(inp["transformed_points"], t_idx, inst_labels, inst_motion_gt, T)
        l2 = torch.norm(rec_est - rec_gt, p=2, dim=1)
        w = t_idx > 0
        wm = (mos_labels == 1) & (t_idx > 0)
        results["inst_l2_error"] = ((l2 * w).sum() / (w.sum() + _EPS)).item()
        results["dynamic_inst_l2_error"] = ((l2 * wm).sum() / (wm.sum() + _EPS)).item()
        results["inst_labels_adjusted"] = inst_labels
        results["inst_pose_est"] = final
        results["sub_rec_est"] = rec_est

    # --- forward (models/motionnet.py:137-262) --------------------------------------------------
    @torch.no_grad()
    def forward(self, input_dict):
        prev = torch.get_default_dtype()
        torch.set_default_dtype(self.dt)  # the factory calls below (zeros / eye / ones) follow the run's precision
        try:
            return self._forward(input_dict)
        finally:
            torch.set_default_dtype(prev)

    def _forward(self, input_dict):
        st = self.stages = {}
        pts = input_dict["input_points"].to(self.dt)
        time_indice = input_dict["time_indice"]
        fb_labels = input_dict["fb_labels"]
        p2v = input_dict["point_to_voxel_map"].long()[:, 0]
        ego_gt = input_dict["ego_motion_gt"].to(self.dt)
        coords = input_dict["coordinates"]
        num_voxels = input_dict["num_voxels"]
        shape = input_dict["shape"][0]
        self.Nx, self.Ny, self.nt = int(shape[0]), int(shape[1]), int(shape[3])
        M = coords.size(0)
        B = num_voxels.size(0)
        pillar_mean = seg_mean(pts, p2v, M)
        fb_sub = seg_max(fb_labels, p2v, M)
        results = {}
        occ_map = self.scatter_canvas(torch.ones(M, 1), coords, B).permute(0, 2, 1, 3, 4)
        fb_map = self.scatter_canvas(fb_sub, coords, B).permute(0, 2, 1, 3, 4)
        mean_map = self.scatter_canvas(pillar_mean, coords, B).permute(0, 2, 1, 3, 4)
        results["fb_seg_gt"], results["occ_map"] = fb_map, occ_map
        st["pillar_mean"] = pillar_mean

        feats = self.pillar_encoder(pts, p2v, coords, pillar_mean, time_indice, M)
        st["pillar_feats"] = feats
        bev = self.scatter_canvas(feats, coords, B)
        _, C, T, Ny, Nx = bev.shape
        bev = bev.permute(0, 2, 1, 3, 4).contiguous().view(B * T, C, Ny, Nx)
        bev_feats = self.unet(bev, "unet.", self.cfg["unet"]["depth"], final=True)
        st["bev_feats"] = bev_feats

        fb_seg = self.seghead2d(bev_feats, "semseg_head").view(B, T, 2, Ny, Nx)
        fb_est = fb_seg.max(dim=2, keepdim=True)[1]
        results["fb_seg_est"] = fb_seg
        if "fb_est_map" in self.inject:
            fb_est = self.inject["fb_est_map"].long().view(B, T, 1, Ny, Nx)
        fb_pillar = self.gather_canvas(fb_est.permute(0, 2, 1, 3, 4).contiguous(), coords)
        fb_pp = fb_pillar[p2v]
        results["fb_est_per_points"] = fb_pp

        geo = self.seghead2d(bev_feats, "ego_feats_head")
        geo = geo / torch.norm(geo, p=2, dim=1, keepdim=True)
        st["geo_feats"] = geo
        self.ego_motion(geo.view(B, T, -1, Ny, Nx), fb_est, occ_map, mean_map, ego_gt, results, raw=(pts, time_indice, fb_pp))

        pose_est = results["ego_motion_est"].to(self.dt)
        if "ego_motion_est" in self.inject:
            pose_est = self.inject["ego_motion_est"].to(self.dt)
        bev_feats = bev_feats.view(B, T, -1, Ny, Nx)
        warped = self.warp_feats(bev_feats, pose_est).permute(0, 2, 1, 3, 4)
        st["warped_feats"] = warped
        tp = self.transform_points(pts.clone(), time_indice, pose_est)
        results["transformed_points"] = tp

        if self.mode in ("train", "val"):
            fb_mask = torch.logical_or(fb_labels[:, 0] == 1, fb_pp[:, 0] == 1)
        else:
            fb_mask = fb_pp[:, 0] == 1
        full_mos = torch.zeros(tp.size(0), 2)
        full_off = torch.zeros(tp.size(0), 2)
        full_mos[:, 0] = 1
        mos_feats = None
        if fb_mask.sum() > MIN_POINTS:
            mos, off, mos_feats = self.stpn(warped, tp.clone()[fb_mask], time_indice[fb_mask])
            full_mos[fb_mask] = mos
            full_off[fb_mask] = off
            st["mos_feats"] = mos_feats
        results["mos_est"], results["offset_est"] = full_mos, full_off
        if "mos_est" in self.inject:
            full_mos = self.inject["mos_est"].to(self.dt)
        if "offset_est" in self.inject:
            full_off = self.inject["offset_est"].to(self.dt)
        if "transformed_points" in self.inject:
            tp = self.inject["transformed_points"].to(self.dt)
        results["rec_est"] = tp.clone()

        if self.mode in ("train", "val"):
            inst_labels = input_dict["inst_labels"][:, 0].long()
            rec_mask = input_dict["fb_labels"][:, 0] == 1
        else:
            if "inst_labels_est" in self.inject:  # (a float64 run would quantise the 5 cm dedupe hash differently)
                inst_labels = self.inject["inst_labels_est"].long()
            else:
                inst_labels = self.cluster(tp, full_mos.argmax(1), full_off, time_indice)
            results["inst_labels_est"] = inst_labels
            rec_mask = inst_labels != 0
        if rec_mask.sum() > MIN_POINTS:
            if mos_feats is None:  # quirk Q4: upstream raises NameError here
                raise NameError("mos_feats is undefined: STPN was skipped but instances exist (motionnet.py:222-245)")
            bb = self.temporal_ungrid(bev_feats, pts[rec_mask].clone(), time_indice[rec_mask])
            mf = self.ungrid(mos_feats, tp[rec_mask].clone(), time_indice[rec_mask])
            st["backbone_feats"], st["motion_feats"] = bb, mf
            self.alignnet({
                "inst_labels": inst_labels[rec_mask], "time_indice": time_indice[rec_mask],
                "transformed_points": tp[rec_mask], "backbone_feats": bb, "motion_feats": mf,
                "inst_motion_gt": input_dict["inst_motion_gt"], "mos_labels": input_dict["sd_labels"][rec_mask, 0].long(),
                "ego_motion_est": results["ego_motion_est"], "ego_motion_gt": results["ego_motion_gt"],
            }, results)
            results["rec_est"][rec_mask] = results["sub_rec_est"]
        return results


# ------------------------------------------------------------------------------------------------
# evaluation tail of the test loop.  TEST INFRASTRUCTURE like the rest of this file.
# ------------------------------------------------------------------------------------------------
def ego_motion_compensation(points, time_indice, tsfm):
    """toolbox/register_utils.py:59-69."""
    pt = tsfm[time_indice.long()]
    return (torch.matmul(pt[:, :3, :3], points[:, :, None]) + pt[:, :3, 3][:, :, None]).squeeze(-1)


def reconstruct_sequence(points, time_indice, inst_labels, tsfm, n_frames):
    """toolbox/register_utils.py:72-93."""
    pt = tsfm.view(-1, 4, 4)[(inst_labels.long() * n_frames + time_indice).long()]
    return (torch.matmul(pt[:, :3, :3], points[:, :, None]) + pt[:, :3, 3][:, :, None]).squeeze(-1)


def sf_counts(epe, rel):
    """Counts behind toolbox/sf_eval_utils.py:71-88 (compute_sf_metrics_torch): n, sum EPE, Acc3DS, Acc3DR, Outlier, ROutlier."""
    return [int(epe.numel()), float(epe.double().sum()),
            int(torch.logical_or(epe < 0.05, rel < 0.05).sum()), int(torch.logical_or(epe < 0.1, rel < 0.1).sum()),
            int(torch.logical_or(epe > 0.3, rel > 0.1).sum()), int(torch.logical_and(epe > 0.3, rel > 0.3).sum())]


def flow_eval(input_dict, predictions, n_frames):
    """libs/tester.py:58-88 for one scene (B = 1) + the category split of toolbox/sf_eval_utils.py:90-102 (restricted to the
    points the tester keeps, ``time_indice > 0``) + the motion-segmentation IoU counters of libs/loss.py:17-48,139-149."""
    pts = input_dict["input_points"].float()
    t = input_dict["time_indice"][:, 1].long()
    ego_gt = input_dict["ego_motion_gt"].float()[0]
    inst_gt = input_dict["inst_motion_gt"][0].float()
    inst = input_dict["inst_labels"][:, 0]
    fb, sd = input_dict["fb_labels"][:, 0], input_dict["sd_labels"][:, 0]
    rec_gt = reconstruct_sequence(ego_motion_compensation(pts, t, ego_gt), t, inst, inst_gt, n_frames)
    err = (predictions["rec_est"] - pts) - (rec_gt - pts)
    epe = torch.norm(err, p=2, dim=1)
    rel = epe / (torch.norm(rec_gt - pts, p=2, dim=1) + 1e-20)
    sel = t > 0
    out = {"epe_per_point": epe, "relative_error": rel, "sel": sel,
           "sf": {"all": sf_counts(epe[sel], rel[sel]), "dynamic": sf_counts(epe[sel & (sd == 1)], rel[sel & (sd == 1)]),
                  "static": sf_counts(epe[sel & (fb == 1)], rel[sel & (fb == 1)])}}
    mask = torch.logical_or(fb == 1, predictions["fb_est_per_points"][:, 0] == 1)
    pred, gt = predictions["mos_est"].argmax(1)[mask], sd.long()[mask]
    out["mos"] = {"masked": int(mask.sum()),
                  "intersection": [int(((pred == c) & (gt == c)).sum()) for c in (0, 1)],
                  "pred_positives": [int((pred == c).sum()) for c in (0, 1)],
                  "gt_positives": [int((gt == c).sum()) for c in (0, 1)]}
    return out


def prep_input_augmented(raw_points, time_indice, sd_labels, fb_labels, inst_labels, ego_motion_gt, inst_motion_gt, cfg):
    """libs/dataset.py:147-207 WITH step 1 (``self.augmentation``): random rigid transform (:101-111; toolbox/register_utils.py:
    199-206), jitter + scale (:90-98), conjugated ground-truth motions (:113-133), then steps 2-4.  Random numbers come from
    numpy's GLOBAL stream in the reference's order, so ``np.random.seed(s)`` before the call replays a reference run."""
    from scipy.spatial.transform import Rotation

    da, T = cfg["data_aug"], cfg["data"]["n_frames"]
    euler = [0, 0, np.random.uniform(0, np.pi * da["rot_aug"])]
    rot = Rotation.from_euler("xyz", euler).as_matrix()
    shift = [np.random.uniform(-da["augment_shift_range"], da["augment_shift_range"]),
             np.random.uniform(-da["augment_shift_range"], da["augment_shift_range"]), 0]
    tsfm = np.eye(4)
    tsfm[:3, :3], tsfm[:3, 3] = rot, np.array(shift)
    raw_points = (tsfm[:3, :3] @ raw_points.T + tsfm[:3, 3][:, None]).T
    raw_points += (np.random.rand(raw_points.shape[0], 3) - 0.5) * da["augment_noise"]
    raw_points = raw_points * np.random.uniform(da["augment_scale_min"], da["augment_scale_max"])
    c = tsfm[None].repeat(T, 0)
    ego_motion_gt = c @ ego_motion_gt @ np.linalg.inv(c)
    im = inst_motion_gt.reshape(-1, 4, 4)
    c = tsfm[None].repeat(im.shape[0], 0)
    inst_motion_gt = (c @ im @ np.linalg.inv(c)).reshape(-1, T, 4, 4)
    data = prep_input_test_mode(raw_points, time_indice, sd_labels, fb_labels, inst_labels, cfg)
    data["ego_motion_gt"], data["inst_motion_gt"] = ego_motion_gt, inst_motion_gt
    return data


def prep_input_test_mode(raw_points, time_indice, sd_labels, fb_labels, inst_labels, cfg):
    """libs/dataset.py:163-207, steps 2-4 of ``BaseDataset.prep_input`` (no augmentation): crop, ground removal, voxelise."""
    vg, dc = cfg["voxel_generator"], cfg["data"]
    crop_xy, z_min, z_max = vg["crop_range"]
    ground = dc["ground_height"] + dc["ground_slack"]
    sel_xy = np.logical_and(np.abs(raw_points[:, 0]) < crop_xy, np.abs(raw_points[:, 1]) < crop_xy)
    sel_z = np.logical_and(raw_points[:, 2] < z_max, raw_points[:, 2] > z_min)
    sel = np.logical_and(sel_xy, sel_z)
    raw_points, time_indice = raw_points[sel], time_indice[sel]
    sd_labels, fb_labels, inst_labels = sd_labels[sel], fb_labels[sel], inst_labels[sel]
    if dc["remove_ground"]:
        ng = raw_points[:, 2] > ground
        raw_points, time_indice = raw_points[ng], time_indice[ng]
        sd_labels, fb_labels, inst_labels = sd_labels[ng], fb_labels[ng], inst_labels[ng]
    points = np.concatenate((raw_points, time_indice[:, None]), axis=1).astype(np.float32)
    data = {"input_points": raw_points, "num_points": np.array([raw_points.shape[0]], dtype=np.int64),
            "time_indice": time_indice[:, None], "sd_labels": sd_labels[:, None], "inst_labels": inst_labels[:, None],
            "fb_labels": fb_labels[:, None]}
    data.update(voxelize(points, vg["voxel_size"], vg["range"], vg["n_sweeps"]))
    return data


def cluster_eval(inst_est, inst_gt, mos_label):
    """toolbox/cluster_eval.py:71-152 (ClusterEvaluation.forward) for one scene, returning the per-scene quantities it appends:
    per class [mean_cov or None, mean_weighted_cov or None, n_gt_inst] and tp / fp lists per IoU threshold."""
    thresholds = [0.5, 0.6, 0.7, 0.8, 0.9]
    mos_label = mos_label.float()

    def instances(labels):
        out = [[], []]
        for uid in torch.unique(labels):
            if uid == 0:
                continue
            m = labels == uid
            out[round(mos_label[m].mean().item())].append(m)
        return out

    est, gt = instances(inst_est), instances(inst_gt)
    res = {"cov": [], "tp": {t: [None, None] for t in thresholds}, "fp": {t: [None, None] for t in thresholds}}
    for c in range(2):
        sum_cov, wcov, npts = 0.0, 0.0, 0
        for g in gt[c]:
            ovmax, ng = 0.0, g.sum().item()
            npts += ng
            for e in est[c]:
                iou = float((g & e).sum() / (g | e).sum())
                ovmax = max(ovmax, iou)
            sum_cov += ovmax
            wcov += ovmax * ng
        n_inst = len(gt[c])
        res["cov"].append([sum_cov / n_inst if n_inst else None, wcov / npts if n_inst else None, n_inst])
    for c in range(2):
        tp = {t: 0 for t in thresholds}
        fp = {t: 0 for t in thresholds}
        for e in est[c]:
            ovmax = -1.0
            for g in gt[c]:
                iou = float((e & g).sum() / (e | g).sum())
                ovmax = max(ovmax, iou)
            for t in thresholds:
                if ovmax > t:
                    tp[t] += 1
                else:
                    fp[t] += 1
        for t in thresholds:
            res["tp"][t][c], res["fp"][t][c] = tp[t], fp[t]
    return res
