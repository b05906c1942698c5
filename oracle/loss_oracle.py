"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch, differentiable) of the reference's training loss ``FuseLoss``
(``libs/loss.py:52-320``, ``libs/lovasz_softmax.py:56-94``, ``libs/outlier_loss.py:15-29``), the checker of
``pcaccumulation_b200/loss.py``.  Pinned: ``tests/test_oracle.py::test_fuse_loss_restatement_matches_reference_when_present``
runs the UNMODIFIED ``libs.loss.FuseLoss`` (imported through ``oracle/ref_loader.py``) on the same predictions and compares
every stat and the autograd gradients.  Only ``tests/`` imports this file.
"""
import numpy as np
import torch

_EPS = 1e-20  # toolbox/utils.py:13

DEFAULT_WEIGHTS = {  # configs/default.yaml:99-113
    "w_pose_l1_loss": 1.0, "w_perm_loss": 0.005, "w_mos_bce_loss": 1.0, "w_mos_lovasz_loss": 1.0, "w_fb_bce_loss": 1.0,
    "w_fb_lovasz_loss": 1.0, "w_offset_norm_loss": 0.5, "w_offset_dir_loss": 0.5, "w_obj_l1_loss": 1.0, "w_obj_pose_loss": 1.0,
    "w_obj_loss": 0.3, "w_obj_rot_loss": 50, "w_obj_trans_loss": 1.0, "obj_gamma": 0.7,
}


def compute_iou(pred, gt, n_class=2, ignore_index=-1):
    """libs/loss.py:17-48 (counts in thousands)."""
    inter, union, pp, gp = [], [], [], []
    for c in range(n_class):
        if c == ignore_index:
            continue
        sg, sp = gt == c, pred == c
        pp.append(sp.sum().item() / 1e3), gp.append(sg.sum().item() / 1e3)
        i = (pred[sg] == c).sum().item() / 1e3
        inter.append(i), union.append(sp.sum().item() / 1e3 + sg.sum().item() / 1e3 - i)
    return {"intersection": np.array(inter), "union": np.array(union), "pred_positives": np.array(pp), "gt_positives": np.array(gp)}


def ce_weights(gt, n_classes=2, max_weights=50):
    """libs/loss.py:93-111, weights_mode 'sqrt_inv_freq'."""
    counts = torch.tensor([(gt == c).sum().item() + _EPS for c in range(n_classes)])
    return torch.clamp(torch.sqrt(counts.sum() / counts), 0, max_weights)


def lovasz_grad(gt_sorted):
    """libs/lovasz_softmax.py:56-69 (the reference's .float() casts are the identity on its float32 inputs; the dtype of the
    input is kept so that a float64 run of this file is the exact-arithmetic yardstick of the tests)."""
    p = len(gt_sorted)
    gts = gt_sorted.sum()
    inter = gts - gt_sorted.cumsum(0)
    union = gts + (1 - gt_sorted).cumsum(0)
    jac = 1.0 - inter / union
    if p > 1:
        jac[1:p] = jac[1:p] - jac[0:-1]
    return jac


def lovasz_softmax_flat(probas, labels):
    """libs/lovasz_softmax.py:72-94."""
    if probas.numel() == 0:
        return probas * 0.0
    losses = []
    for c in range(probas.size(1)):
        fg = (labels == c).to(probas.dtype)
        if fg.sum() == 0:
            continue
        errors = (fg - probas[:, c]).abs()
        errors_sorted, perm = torch.sort(errors, 0, descending=True)
        losses.append(torch.dot(errors_sorted, lovasz_grad(fg[perm.data])))
    return sum(losses) / len(losses)


def seg_loss(gt, est):
    """libs/loss.py:113-136."""
    ce = torch.nn.CrossEntropyLoss(weight=ce_weights(gt).to(est.dtype), ignore_index=-1)(est, gt)
    lov = lovasz_softmax_flat(torch.softmax(est, dim=1), gt)
    return {"bce_loss": ce, "lovasz_loss": lov, "metric": compute_iou(est.argmax(1), gt)}


def fb_loss(pred):
    """libs/loss.py:165-186."""
    est = pred["fb_seg_est"].permute(0, 1, 3, 4, 2).contiguous().view(-1, 2)
    gt = pred["fb_seg_gt"].permute(0, 1, 3, 4, 2).contiguous().view(-1)
    mask = pred["occ_map"].permute(0, 1, 3, 4, 2).contiguous().view(-1) == 1
    return seg_loss(gt[mask], est[mask])


def mos_loss(pred, inp):
    """libs/loss.py:139-163."""
    gt, est = inp["sd_labels"][:, 0].long(), pred["mos_est"]
    mask = torch.logical_or(inp["fb_labels"][:, 0] == 1, pred["fb_est_per_points"][:, 0] == 1)
    if mask.sum():
        return seg_loss(gt[mask], est[mask])
    z = np.zeros(2)
    return {"metric": {"intersection": z, "union": z, "pred_positives": z, "gt_positives": z}, "bce_loss": torch.tensor(0.0),
            "lovasz_loss": torch.tensor(0.0)}


def _apply(tsfm, pts):
    return (torch.matmul(tsfm[:, :3, :3], pts[:, :, None]) + tsfm[:, :3, 3][:, :, None]).squeeze(-1)


def offset_loss(inp, pred):
    """libs/loss.py:189-245 (GT reconstruction: toolbox/register_utils.py:59-93; scatter mean = sum / count)."""
    pts, ti = inp["input_points"], inp["time_indice"]
    inst = inp["inst_labels"][:, 0].long()
    fb_mask = inp["fb_labels"][:, 0] == 1
    if not fb_mask.sum():
        return torch.tensor(0.0), torch.tensor(0.0), 0, None
    T = inp["ego_motion_gt"].size(1)
    centres = []
    for b in range(len(inp["inst_motion_gt"])):
        sel = ti[:, 0] == b
        lab, t = inst[sel], ti[sel, 1].long()
        comp = _apply(inp["ego_motion_gt"][b][t], pts[sel])
        rec = _apply(inp["inst_motion_gt"][b].view(-1, 4, 4)[lab * T + t], comp)
        K = int(lab.max()) + 1
        s = torch.zeros(K, 3, dtype=rec.dtype).index_add_(0, lab, rec)
        cnt = torch.zeros(K, dtype=rec.dtype).index_add_(0, lab, torch.ones(lab.numel(), dtype=rec.dtype)).clamp(min=1)
        centres.append((s / cnt[:, None])[lab])
    centres = torch.cat(centres)[:, :2]
    gt_off = (centres - pred["transformed_points"][:, :2])[fb_mask]
    est = pred["offset_est"][fb_mask]
    norm_loss = torch.abs(gt_off - est).mean(dim=0).sum()
    l2 = torch.norm(gt_off - est, p=2, dim=1).mean().item()
    ng = gt_off / (torch.norm(gt_off, dim=1, p=2).unsqueeze(-1) + _EPS)
    ne = est / (torch.norm(est, dim=1, p=2).unsqueeze(-1) + _EPS)
    return norm_loss, (1 - (ng * ne).sum(-1)).mean(), l2, gt_off


def outlier_loss(perm):
    """libs/outlier_loss.py:15-29, reduction 'mean'."""
    ref = torch.cat([1.0 - torch.sum(p, dim=1) for p in perm], 1)
    src = torch.cat([1.0 - torch.sum(p, dim=2) for p in perm], 0)
    return torch.mean(ref) + torch.mean(src)


def fuse_loss(pred, inp, w=None):
    """libs/loss.py:273-320."""
    w = dict(DEFAULT_WEIGHTS, **(w or {}))
    stats = {}
    total = w["w_pose_l1_loss"] * pred["ego_l1_loss"]
    stats["ego_l1_loss"] = total
    for k in ("ego_l2_loss", "ego_rot_error", "ego_trans_error"):
        stats[k] = pred[k]
    stats["perm_loss"] = outlier_loss(pred["perm_matrix"]) * w["w_perm_loss"]
    total = total + stats["perm_loss"]
    fb = fb_loss(pred)
    stats["fb_loss"] = w["w_fb_bce_loss"] * fb["bce_loss"] + w["w_fb_lovasz_loss"] * fb["lovasz_loss"]
    stats["fb_metric"] = fb["metric"]
    total = total + stats["fb_loss"]
    mos = mos_loss(pred, inp)
    stats["mos_loss"] = w["w_mos_bce_loss"] * mos["bce_loss"] + w["w_mos_lovasz_loss"] * mos["lovasz_loss"]
    stats["mos_metric"] = mos["metric"]
    total = total + stats["mos_loss"]
    norm_l, dir_l, l2, gt_off = offset_loss(inp, pred)
    stats["offset_loss"] = dir_l * w["w_offset_dir_loss"] + norm_l * w["w_offset_norm_loss"]
    stats["offset_l1_loss"], stats["offset_dir_loss"], stats["offset_l2_error"] = norm_l, dir_l, l2
    stats["offset_gt"] = gt_off
    total = total + stats["offset_loss"]
    if "tpointnet_loss_terms" in pred:
        obj, n_it = 0, len(pred["tpointnet_loss_terms"])
        for n_th, v in enumerate(pred["tpointnet_loss_terms"].values(), 1):
            pose = w["w_obj_trans_loss"] * v["trans_loss"] + w["w_obj_rot_loss"] * v["rot_loss"]
            obj = obj + (w["w_obj_l1_loss"] * v["l1_loss"] + w["w_obj_pose_loss"] * pose) * w["obj_gamma"] ** (n_it - n_th)
        stats["obj_loss"] = obj * w["w_obj_loss"]
        total = total + stats["obj_loss"]
        stats["inst_l2_error"], stats["dynamic_inst_l2_error"] = pred["inst_l2_error"], pred["dynamic_inst_l2_error"]
    stats["loss"] = total
    return stats
