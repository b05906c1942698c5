"""Configuration dictionaries with the reference's keys and default values.

The reference builds one nested ``config`` dict from ``configs/default.yaml`` merged with a dataset
YAML (``toolbox/config.py:119-138``) and then copies three voxeliser fields into the pillar-encoder
section (``main.py:10-14``).  ``MotionNet(cfg)`` reads exactly these keys
(``models/motionnet.py:16-42`` and the sub-module constructors), so they are kept verbatim here and a
dict produced by the reference's own ``get_config`` + ``update_config`` can be passed instead.
"""
import copy

_DEFAULT = {
    "misc": {"use_gpu": True, "seed": 42, "mode": "test", "exp_name": "synthetic", "pretrain": ""},
    "data": {"speed_threshold": 0.5, "max_speed": 20, "ground_slack": 0.3, "remove_ground": True},
    "cluster": {
        "cluster_metric": "euclidean",
        "min_p_cluster": 15,
        "min_samples_dbscan": 5,
        "eps_dbscan": 0.4,
        "voxel_size": 0.15,
    },
    "pillar_encoder": {"depth": 3, "num_input_features": 9, "num_filters": 32},
    "unet": {"start_filts": 32, "in_channels": 32, "depth": 5, "merge_mode": "concat"},
    "pose_estimation": {
        "n_kpts": 1024,
        "add_slack": True,
        "sinkhorn_iter": 3,
        "feats_dim": 64,
        "icp_threshold": 0.15,
        "icp_max_iter": 50,
        "seq_pose": "skip",
    },
    "stpn": {"feat_dim": 32},
    "tpointnet": {"n_iterations": 1, "min_points": 10, "icp_threshold": 0.25},
    "model": {"ego_icp": False, "tpointnet_icp": False},
    # configs/default.yaml:44-49 (training-time augmentation) and :99-113 (loss weights)
    "data_aug": {"augment_noise": 0.01, "augment_shift_range": 0.25, "augment_scale_min": 0.995, "augment_scale_max": 1.005,
                 "rot_aug": 0.5},
    "loss": {"w_pose_l1_loss": 1.0, "w_perm_loss": 0.005, "w_mos_bce_loss": 1.0, "w_mos_lovasz_loss": 1.0, "w_fb_bce_loss": 1.0,
             "w_fb_lovasz_loss": 1.0, "w_offset_norm_loss": 0.5, "w_offset_dir_loss": 0.5, "w_obj_l1_loss": 1.0,
             "w_obj_pose_loss": 1.0, "w_obj_loss": 0.3, "w_obj_rot_loss": 50, "w_obj_trans_loss": 1.0, "obj_gamma": 0.7},
}

# configs/waymo/waymo.yaml
_WAYMO = {
    "voxel_generator": {
        "range": [-36, -36, -2, 36, 36, 6],
        "voxel_size": [0.25, 0.25, 8],
        "n_sweeps": 5,
        "crop_range": [32, -2, 6],
    },
    "data": {"dataset": "waymo", "n_frames": 5, "interval": 1, "freq": 10.0, "ground_height": 0.04, "max_speed": 30},
    "pose_estimation": {"icp_threshold": 0.1},
    "tpointnet": {"n_iterations": 2, "min_points": 50, "icp_threshold": 0.15},
}

# configs/nuscene/nuscene.yaml
_NUSCENE = {
    "voxel_generator": {
        "range": [-36, -36, -5, 36, 36, 3],
        "voxel_size": [0.25, 0.25, 8],
        "n_sweeps": 11,
        "crop_range": [32, -5, 3],
    },
    "data": {
        "dataset": "nuscene",
        "n_frames": 11,
        "interval": 1,
        "freq": 20.0,
        "ground_height": -1.84,
        "radius": 1.0,
        "max_speed": 10,
    },
    "pose_estimation": {"icp_threshold": 0.2},
    "tpointnet": {"n_iterations": 2, "min_points": 50, "icp_threshold": 0.25},
}


def update_recursive(dst, src):
    """Same merge rule as the reference (``toolbox/config.py:141-155``)."""
    for k, v in src.items():
        if isinstance(v, dict):
            dst.setdefault(k, {})
            update_recursive(dst[k], v)
        else:
            dst[k] = v
    return dst


def get_config(dataset="waymo", mode="test", **overrides):
    """Return the merged config for ``dataset`` in {'waymo','nuscene'}.

    ``overrides`` are nested dicts merged last, e.g. ``voxel_generator={'n_sweeps': 10}``; when
    ``n_sweeps`` is overridden ``data.n_frames`` should be overridden consistently by the caller
    (the reference does the same through ``--voxel_generator.n_sweeps=10 --data.n_frames=10``).
    """
    cfg = copy.deepcopy(_DEFAULT)
    update_recursive(cfg, copy.deepcopy({"waymo": _WAYMO, "nuscene": _NUSCENE}[dataset]))
    update_recursive(cfg, copy.deepcopy(overrides))
    cfg["misc"]["mode"] = mode
    finalize_config(cfg)
    return cfg


def finalize_config(cfg):
    """``main.py:10-14``: the pillar encoder reads the voxeliser geometry from its own section."""
    cfg["pillar_encoder"]["voxel_size"] = cfg["voxel_generator"]["voxel_size"]
    cfg["pillar_encoder"]["pc_range"] = cfg["voxel_generator"]["range"]
    cfg["pillar_encoder"]["n_sweeps"] = cfg["voxel_generator"]["n_sweeps"]
    return cfg


# BASELINE.json configs: name -> (dataset, T, points per frame, overrides)
WORKLOADS = {
    "C1": dict(dataset="waymo", T=5, pts_per_frame=20_000, overrides={}),
    "C2": dict(dataset="waymo", T=5, pts_per_frame=150_000, overrides={}),
    "C3": dict(
        dataset="nuscene",
        T=10,
        pts_per_frame=35_000,
        overrides={"voxel_generator": {"n_sweeps": 10}, "data": {"n_frames": 10}},
    ),
    "C5": dict(
        dataset="waymo",
        T=5,
        pts_per_frame=400_000,
        overrides={"voxel_generator": {"range": [-64, -64, -2, 64, 64, 6], "crop_range": [60, -2, 6]}},
    ),
}


def workload_config(name, mode="test"):
    w = WORKLOADS[name]
    return get_config(w["dataset"], mode=mode, **w["overrides"])
