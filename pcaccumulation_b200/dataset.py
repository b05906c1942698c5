"""Host side of the GPU data front-end: what ``libs/dataset.py:BaseDataset`` does around the
per-point work, with the per-point work itself (augmentation, crop, ground removal, voxelisation, collate) on the device
(``csrc/dataprep.cu``, ``csrc/voxelize.cu``, ``SceneRunner.build_input``).

  * ``load_sample``            ``__getitem__`` (:209-224): the arrays of one ``.npz`` sample file (writer:
                               dataset_toolbox/prep_nuscene_waymo_sf/waymo.py:206-220), reference key names
  * ``sample_augmentation``    the random draws of ``_sample_random_tsfm`` (:101-111) and ``apply_data_augmentation`` (:90-98) in
                               the reference's ORDER on numpy's global stream (the H3-style protocol: the host owns the random
                               numbers, the device applies them), so a seeded run replays the reference's augmentation exactly
  * ``update_transformation_after_data_augmentation``   (:113-133) on the [T,4,4] / [K,T,4,4] ground-truth motions (a few
                               4x4 products: stays in numpy float64 like the reference)
"""
import numpy as np
from scipy.spatial.transform import Rotation

SAMPLE_KEYS = ("raw_points", "time_indice", "sd_labels", "fb_labels", "inst_labels", "sem_labels", "ego_motion_gt", "bbox_tsfm")


def load_sample(path):
    """One ``np.savez_compressed`` sample -> dict with the names ``prep_input`` uses (``bbox_tsfm`` -> ``inst_motion_gt``)."""
    data = np.load(path, allow_pickle=True)
    out = {k: data[k] for k in SAMPLE_KEYS if k in data.files}
    out["inst_motion_gt"] = out.pop("bbox_tsfm")
    out["data_path"] = path
    return out


def sample_augmentation(data_aug, n_points, exact_noise=True, rng=np.random):
    """Random numbers of one augmented sample.  ``exact_noise`` draws the [n,3] jitter on the host exactly where the reference
    does (bit-identical replay of a seeded reference run); otherwise only a seed for the device generator is drawn."""
    euler = [0, 0, rng.uniform(0, np.pi * data_aug["rot_aug"])]
    rot = Rotation.from_euler("xyz", euler).as_matrix()
    r = data_aug["augment_shift_range"]
    shift = [rng.uniform(-r, r), rng.uniform(-r, r), 0]
    tsfm = np.eye(4)
    tsfm[:3, :3] = rot
    tsfm[:3, 3] = np.array(shift)
    if exact_noise:
        noise, seed = rng.rand(n_points, 3), 0
    else:
        noise, seed = None, int(rng.randint(0, 2**31 - 1))
    scale = rng.uniform(data_aug["augment_scale_min"], data_aug["augment_scale_max"])
    return {"tsfm": tsfm, "noise": noise, "seed": seed, "noise_amp": float(data_aug["augment_noise"]), "scale": float(scale)}


def update_transformation_after_data_augmentation(aug_tsfm, ego_motion, inst_motion, n_frames):
    """T'_ego = T' T_ego T'^-1 (and the same for every instance motion)."""
    c = aug_tsfm[None].repeat(n_frames, 0)
    ego_motion = c @ ego_motion @ np.linalg.inv(c)
    inst_motion = inst_motion.reshape(-1, 4, 4)
    c = aug_tsfm[None].repeat(inst_motion.shape[0], 0)
    inst_motion = c @ inst_motion @ np.linalg.inv(c)
    return ego_motion, inst_motion.reshape(-1, n_frames, 4, 4)
