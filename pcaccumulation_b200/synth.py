"""Deterministic synthetic multi-sweep LiDAR scenes and the reference's ``input_dict`` schema.

There is no dataset in the build or bench environment, so every test and benchmark runs on scenes
generated here: a static world of vertical planar patches, boxes (half of them
moving) and clutter, observed from a moving ego vehicle, already ground-removed and cropped the way
``libs/dataset.py:170-180`` does it.  ``collate`` restates ``libs/dataloader.py:7-40`` so the
resulting ``input_dict`` has exactly the dtypes the reference's DataLoader hands to the model
(float64 ``coordinates``/``time_indice`` included).
"""
import numpy as np

from .config import WORKLOADS, workload_config


def _yaw(a):
    c, s = np.cos(a), np.sin(a)
    m = np.eye(4)
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
    return m


def _pose(yaw, xy):
    m = _yaw(yaw)
    m[0, 3], m[1, 3] = xy[0], xy[1]
    return m


def make_scene(T=5, pts_per_frame=150_000, seed=0, freq=10.0, z_lo=0.4, z_hi=5.0, crop_xy=32.0,
               max_range=45.0, n_boxes=None):
    """One scene in the per-sample format ``libs/dataset.py:prep_input`` returns (before voxelising).

    Returns a dict with ``input_points f32[N,3]`` (sensor coordinates of each point's own sweep),
    ``time_indice i64[N,1]``, ``fb_labels/sd_labels/inst_labels i64[N,1]``, ``ego_motion_gt f32[T,4,4]``
    (sweep t -> sweep 0, frame 0 = I), ``inst_motion_gt f32[K,T,4,4]`` (instance 0 = background = I)
    and ``num_points i64[1]``.
    """
    rng = np.random.default_rng(seed)
    dt = 1.0 / freq
    # ego: constant speed / yaw-rate
    speed = rng.uniform(0.0, 15.0)
    yaw_rate = rng.uniform(-0.1, 0.1)
    ego = np.zeros((T, 4, 4))
    x = y = th = 0.0
    sub = 20
    for t in range(T):
        ego[t] = _pose(th, (x, y))
        for _ in range(sub):
            x += speed * np.cos(th) * dt / sub
            y += speed * np.sin(th) * dt / sub
            th += yaw_rate * dt / sub
    ego_inv = np.linalg.inv(ego)

    # static world: vertical planar patches (walls/fences)
    n_walls = int(rng.integers(120, 200))
    r = rng.uniform(2.0, max_range, n_walls)
    ang = rng.uniform(0, 2 * np.pi, n_walls)
    wc = np.stack([r * np.cos(ang), r * np.sin(ang)], 1)
    wlen = rng.uniform(5.0, 30.0, n_walls)
    wdir = rng.uniform(0, np.pi, n_walls)
    wh = rng.uniform(0.4, z_hi - z_lo, n_walls)
    # boxes
    if n_boxes is None:
        n_boxes = int(rng.integers(20, 61))
    br = rng.uniform(3.0, 0.9 * crop_xy, n_boxes)
    ba = rng.uniform(0, 2 * np.pi, n_boxes)
    bc = np.stack([br * np.cos(ba), br * np.sin(ba)], 1)
    is_ped = rng.random(n_boxes) < 0.3
    bdim = np.where(is_ped[:, None], np.array([[0.6, 0.6, 1.7]]), np.array([[4.5, 1.9, 1.6]]))
    bdim = bdim * rng.uniform(0.9, 1.1, (n_boxes, 3))
    bdim[:, 2] = np.minimum(bdim[:, 2], z_hi - z_lo - 0.1)
    bhead = rng.uniform(0, 2 * np.pi, n_boxes)
    moving = rng.random(n_boxes) < 0.5
    bspeed = np.where(moving, rng.uniform(1.0, 15.0, n_boxes), 0.0)
    bspeed = np.where(is_ped & moving, rng.uniform(1.0, 2.5, n_boxes), bspeed)

    # box pose (box -> world) per time
    box_pose = np.zeros((n_boxes, T, 4, 4))
    for k in range(n_boxes):
        for t in range(T):
            p = bc[k] + bspeed[k] * t * dt * np.array([np.cos(bhead[k]), np.sin(bhead[k])])
            box_pose[k, t] = _pose(bhead[k], p)
    inst_motion = np.tile(np.eye(4), (n_boxes + 1, T, 1, 1))
    for k in range(n_boxes):
        for t in range(T):
            inst_motion[k + 1, t] = box_pose[k, 0] @ np.linalg.inv(box_pose[k, t])

    pts_all, t_all, fb_all, sd_all, inst_all = [], [], [], [], []
    for t in range(T):
        target = int(round(pts_per_frame * (1.0 + rng.uniform(-0.05, 0.05))))
        over = int(target * 2.2) + 1000
        n_w, n_b = int(over * 0.60), int(over * 0.25)
        n_c = over - n_w - n_b
        # walls
        wi = rng.integers(0, n_walls, n_w)
        s = rng.uniform(-0.5, 0.5, n_w) * wlen[wi]
        pw = np.empty((n_w, 3))
        pw[:, 0] = wc[wi, 0] + s * np.cos(wdir[wi]) + rng.normal(0, 0.01, n_w)
        pw[:, 1] = wc[wi, 1] + s * np.sin(wdir[wi]) + rng.normal(0, 0.01, n_w)
        pw[:, 2] = z_lo + rng.uniform(0, 1, n_w) * wh[wi]
        # boxes: points on the four vertical faces and the top
        bi = rng.integers(0, n_boxes, n_b)
        u = rng.uniform(-0.5, 0.5, (n_b, 3))
        face = rng.integers(0, 5, n_b)
        u[face == 0, 0] = 0.5
        u[face == 1, 0] = -0.5
        u[face == 2, 1] = 0.5
        u[face == 3, 1] = -0.5
        u[face == 4, 2] = 0.5
        local = u * bdim[bi]
        local[:, 2] += bdim[bi, 2] / 2
        lh = np.concatenate([local[:, :2], np.zeros((n_b, 1)), np.ones((n_b, 1))], 1)
        wxy = np.einsum("nij,nj->ni", box_pose[bi, t], lh)[:, :2]
        pb = np.concatenate([wxy, (z_lo + 0.02 + local[:, 2:3])], 1)
        # clutter
        rc = rng.uniform(2.0, max_range, n_c)
        ac = rng.uniform(0, 2 * np.pi, n_c)
        pc = np.stack([rc * np.cos(ac), rc * np.sin(ac), rng.uniform(z_lo, z_hi, n_c)], 1)

        world = np.concatenate([pw, pb, pc], 0)
        inst = np.concatenate([np.zeros(n_w, np.int64), bi + 1, np.zeros(n_c, np.int64)])
        fb = (inst > 0).astype(np.int64)
        sd = np.concatenate([np.zeros(n_w, np.int64), moving[bi].astype(np.int64), np.zeros(n_c, np.int64)])
        # to the sweep's sensor frame
        wh4 = np.concatenate([world, np.ones((world.shape[0], 1))], 1)
        sens = (wh4 @ ego_inv[t].T)[:, :3]
        sens = sens.astype(np.float32)
        keep = (np.abs(sens[:, 0]) < crop_xy) & (np.abs(sens[:, 1]) < crop_xy) & (sens[:, 2] > z_lo - 0.05) & (
            sens[:, 2] < z_hi)
        idx = np.nonzero(keep)[0]
        if idx.shape[0] > target:
            idx = np.sort(rng.choice(idx, target, replace=False))
        perm = rng.permutation(idx.shape[0])  # LiDAR returns are not spatially sorted
        idx = idx[perm]
        pts_all.append(sens[idx])
        t_all.append(np.full(idx.shape[0], t, np.int64))
        fb_all.append(fb[idx])
        sd_all.append(sd[idx])
        inst_all.append(inst[idx])

    pts = np.concatenate(pts_all, 0).astype(np.float32)
    return {
        "input_points": pts,
        "num_points": np.array([pts.shape[0]], dtype=np.int64),
        "time_indice": np.concatenate(t_all)[:, None],
        "sd_labels": np.concatenate(sd_all)[:, None],
        "inst_labels": np.concatenate(inst_all)[:, None],
        "fb_labels": np.concatenate(fb_all)[:, None],
        "ego_motion_gt": ego.astype(np.float32),
        "inst_motion_gt": inst_motion.astype(np.float32),
    }


def make_workload_scene(name, scene_idx=0, pts_per_frame=None):
    """Scene for a BASELINE.json config (C1/C2/C3/C5); seed = 1000*config + scene_idx."""
    w = WORKLOADS[name]
    cfg = workload_config(name)
    rng_lo, rng_hi = cfg["voxel_generator"]["range"][2], cfg["voxel_generator"]["range"][5]
    crop = cfg["voxel_generator"]["crop_range"]
    ground = cfg["data"]["ground_height"] + cfg["data"]["ground_slack"]
    seed = 1000 * int(name[1:]) + scene_idx
    return make_scene(
        T=w["T"],
        pts_per_frame=pts_per_frame or w["pts_per_frame"],
        seed=seed,
        freq=cfg["data"]["freq"],
        z_lo=max(ground, rng_lo) + 0.06,
        z_hi=min(crop[2], rng_hi) - 0.05,
        crop_xy=float(crop[0]),
        max_range=1.4 * float(crop[0]),
    )


def collate(samples):
    """Restatement of ``libs/dataloader.py:7-40`` returning torch CPU tensors.

    ``samples``: list of per-scene dicts (``make_scene`` output merged with the voxeliser's
    ``coordinates i32[M,4]``, ``num_voxels``, ``shape``, ``point_to_voxel_map``).
    """
    import torch

    out = {}
    keys = samples[0].keys()
    for key in keys:
        elems = [s[key] for s in samples]
        if key in ("coordinates", "time_indice"):
            rows = []
            for i, e in enumerate(elems):
                rows.append(np.concatenate((np.ones((e.shape[0], 1)) * i, e), axis=1))  # float64, as upstream
            out[key] = torch.tensor(np.concatenate(rows, axis=0))
        elif key in ("ego_motion_gt", "shape"):
            out[key] = torch.tensor(np.stack(elems, axis=0))
        elif key == "inst_motion_gt":
            out[key] = [torch.tensor(e) for e in elems]
        else:
            out[key] = torch.tensor(np.concatenate(elems, axis=0))
    run, acc = 0, 0
    for b in range(len(samples)):
        n = int(out["num_points"][b])
        out["point_to_voxel_map"][run:run + n] += acc
        run += n
        acc += int(out["num_voxels"][b])
    return out
