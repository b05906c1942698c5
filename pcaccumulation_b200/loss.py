"""``FuseLoss`` of the reference (``libs/loss.py:52-320``) on the device: same constructor keys (``config['loss']``), same
``forward(predictions, input_dict) -> stats`` contract and stat names, consumed by ``libs/trainer.py:165-196`` /
``libs/tester.py:67-93``.

Every term is evaluated by sm_100a kernels behind the C ABI (``csrc/loss.cu``): the two segmentation losses (online class
weights + weighted cross entropy + Lovasz-softmax + IoU counters) as one pass + two device sorts each, the offset loss as two
passes over the points, the outlier loss as one reduction.  Nothing is read back: ``stats`` holds 0-dim device tensors for the
losses (the trainer calls ``float()`` on them) -- except the four values the reference itself returns as python numbers
(``offset_l2_error`` and the IoU counter arrays), which are fetched with ONE synchronisation at the end.

Gradients: the terms are ``torch.autograd.Function``s whose backward returns the analytic gradient with respect to the network
outputs (``fb_seg_est``, ``mos_est``, ``offset_est``, ``perm_matrix``), i.e. the first link of the training step's backward
pass.  ``MotionNet.forward`` itself is inference-only (no autograd through the CUDA stages yet), so ``stats['loss'].backward()``
reaches the predictions and stops there.
"""
import numpy as np
import torch

from ._lib import F, I, L, P, Z, call, scratch, size, stream


def _i64(t):
    return t.reshape(-1).to(torch.int64).contiguous()


class _SegLoss(torch.autograd.Function):
    """w_ce * CE + w_lovasz * Lovasz of libs/loss.py:113-136; returns (combined loss, out13)."""

    @staticmethod
    def forward(ctx, logits, hw, gt, sel_float, sel_a, sel_b, w_ce, w_lov):
        dev = logits.device
        lg = logits.detach().float().contiguous()
        n = lg.numel() // 2
        out = torch.empty(13, device=dev)
        ws = scratch(size("pcab_seg_loss_workspace", L(n)), dev)
        call("pcab_seg_loss", P(lg), I(hw), P(gt), P(sel_float), P(sel_a), P(sel_b), L(n), P(out), P(ws), Z(ws.numel()), stream())
        ctx.saved = (lg, hw, gt, sel_float, sel_a, sel_b, out, ws, n, w_ce, w_lov, logits.shape)
        ctx.mark_non_differentiable(out)
        return w_ce * out[0] + w_lov * out[1], out

    @staticmethod
    def backward(ctx, g_loss, _g_out):
        lg, hw, gt, sel_float, sel_a, sel_b, out, ws, n, w_ce, w_lov, shape = ctx.saved
        grad = torch.empty_like(lg)
        call("pcab_seg_loss_grad", P(lg), I(hw), P(gt), P(sel_float), P(sel_a), P(sel_b), L(n), P(out), F(w_ce), F(w_lov), P(grad),
             P(ws), Z(ws.numel()), stream())
        return (grad * g_loss).view(shape), None, None, None, None, None, None, None


class _OffsetLoss(torch.autograd.Function):
    """w_norm * offset_norm_loss + w_dir * offset_dir_loss of libs/loss.py:189-245; returns (combined, out4, gt_offset)."""

    @staticmethod
    def forward(ctx, offset_est, args, w_norm, w_dir):
        pts, pbatch, ptime, inst, fb, ego_gt, motion, koff, k_total, T, tp = args
        dev = offset_est.device
        off = offset_est.detach().float().contiguous()
        n = off.shape[0]
        out = torch.empty(4, device=dev)
        gt_offset = torch.empty(n, 2, device=dev)
        ws = scratch(size("pcab_offset_loss_workspace", I(k_total)), dev)
        call("pcab_offset_loss", P(pts), P(pbatch), P(ptime), P(inst), P(fb), P(ego_gt), P(motion), P(koff), I(k_total), I(T), P(tp),
             P(off), L(n), P(gt_offset), P(out), P(ws), Z(ws.numel()), stream())
        ctx.saved = (fb, gt_offset, off, out, n, w_norm, w_dir)
        ctx.mark_non_differentiable(out, gt_offset)
        return w_dir * out[1] + w_norm * out[0], out, gt_offset

    @staticmethod
    def backward(ctx, g_loss, _g1, _g2):
        fb, gt_offset, off, out, n, w_norm, w_dir = ctx.saved
        grad = torch.empty_like(off)
        call("pcab_offset_loss_grad", P(fb), P(gt_offset), P(off), L(n), P(out), F(w_norm), F(w_dir), P(grad), stream())
        return grad * g_loss, None, None, None


class _PermLoss(torch.autograd.Function):
    """OutlierLoss (libs/outlier_loss.py:15-29): mean(1 - column sums) + mean(1 - row sums) over all matrices."""

    @staticmethod
    def forward(ctx, stacked):
        dev = stacked.device
        x = stacked.detach().float().contiguous()
        n_mats, m = x.shape[0], x.shape[-1]
        out = torch.empty(1, device=dev)
        acc = torch.empty(1, dtype=torch.float64, device=dev)
        call("pcab_perm_loss", P(x), I(n_mats), I(m), P(acc), P(out), stream())
        ctx.meta = (n_mats, m, stacked.shape)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        n_mats, m, shape = ctx.meta
        return (g * (-2.0 / (n_mats * m))).expand(shape)


def _stack_perm(perm_list):
    """The forward hands out views of ONE [pairs,1,m,m] tensor: recognise that and skip the copy."""
    first = perm_list[0]
    step = first.numel() * first.element_size()
    if all(p.is_contiguous() and p.data_ptr() == first.data_ptr() + i * step for i, p in enumerate(perm_list)) and \
            not any(p.requires_grad for p in perm_list) and first._base is not None and first._base.is_contiguous():
        base = first._base
        i0 = (first.data_ptr() - base.data_ptr()) // step
        if (first.data_ptr() - base.data_ptr()) % step == 0 and i0 + len(perm_list) <= base.numel() // first.numel():
            return base.reshape(-1, *first.shape)[i0:i0 + len(perm_list)]
    return torch.stack(list(perm_list))


class FuseLoss(torch.nn.Module):
    def __init__(self, config):
        super().__init__()
        self.n_classes = 2
        self.ignore_index = -1
        self.weights_mode = "sqrt_inv_freq"
        for k in ("w_pose_l1_loss", "w_perm_loss", "w_mos_bce_loss", "w_mos_lovasz_loss", "w_fb_bce_loss", "w_fb_lovasz_loss",
                  "w_offset_norm_loss", "w_offset_dir_loss", "w_obj_loss", "w_obj_rot_loss", "w_obj_trans_loss", "w_obj_l1_loss",
                  "w_obj_pose_loss", "obj_gamma"):
            setattr(self, k, config[k])

    # -- libs/loss.py:165-186 / :139-163 / :189-245: device terms -> (combined loss, raw outputs) ---------------------------
    def _fb_term(self, predictions):
        est = predictions["fb_seg_est"]  # [B,T,2,Ny,Nx]
        hw = est.shape[-1] * est.shape[-2]
        occ = predictions["occ_map"].reshape(-1).float().contiguous()
        return _SegLoss.apply(est, hw, _i64(predictions["fb_seg_gt"]), occ, None, None, float(self.w_fb_bce_loss),
                              float(self.w_fb_lovasz_loss))

    def _mos_term(self, predictions, input_dict):
        return _SegLoss.apply(predictions["mos_est"], 1, _i64(input_dict["sd_labels"][:, 0]), None, _i64(input_dict["fb_labels"][:, 0]),
                              _i64(predictions["fb_est_per_points"][:, 0]), float(self.w_mos_bce_loss), float(self.w_mos_lovasz_loss))

    def _offset_term(self, input_dict, predictions):
        dev = predictions["offset_est"].device
        ti = input_dict["time_indice"]
        motions = [m.to(dev).float().reshape(m.shape[0], -1, 4, 4) for m in input_dict["inst_motion_gt"]]
        ks = [m.shape[0] for m in motions]
        koff = torch.tensor(np.concatenate(([0], np.cumsum(ks)[:-1])).astype(np.int32), device=dev)
        ego_gt = input_dict["ego_motion_gt"].to(dev).float().contiguous()
        args = (input_dict["input_points"].float().contiguous(), ti[:, 0].to(torch.int32).contiguous(), ti[:, 1].to(torch.int32).contiguous(),
                _i64(input_dict["inst_labels"][:, 0]), _i64(input_dict["fb_labels"][:, 0]), ego_gt, torch.cat(motions).contiguous(), koff,
                int(sum(ks)), int(ego_gt.shape[1]), predictions["transformed_points"].detach().float().contiguous())
        return _OffsetLoss.apply(predictions["offset_est"], args, float(self.w_offset_norm_loss), float(self.w_offset_dir_loss))

    # -- the reference's public methods, same return shapes (libs/tester.py:87-93 calls get_mos_loss / evaluate_cluster) ------
    def _seg_stats(self, out):
        return {"bce_loss": out[0], "lovasz_loss": out[1], "metric": self._metric(out.tolist())}

    def get_fb_loss(self, predictions):
        return self._seg_stats(self._fb_term(predictions)[1])

    def get_mos_loss(self, predictions, input_dict):
        return self._seg_stats(self._mos_term(predictions, input_dict)[1])

    def get_offset_loss(self, input_dict, predictions):
        _, out, gt_offset = self._offset_term(input_dict, predictions)
        host = out.tolist()
        if host[3] > 0:
            predictions["offset_gt"] = gt_offset[_i64(input_dict["fb_labels"][:, 0]) == 1]
        return out[0], out[1], (host[2] if host[3] > 0 else 0)

    def evaluate_cluster(self, predictions, input_dict):
        """libs/loss.py:261-270: per scene, the estimated instances against the labels (scores accumulate in
        ``self.cluster_eval_offset``, an ``evaluation.ClusterEvaluator``)."""
        from .evaluation import ClusterEvaluator

        if getattr(self, "cluster_eval_offset", None) is None:
            self.cluster_eval_offset = ClusterEvaluator(predictions["inst_labels_est"].device)
        ti = input_dict["time_indice"]
        for b in range(int(ti[:, 0].max() + 1)):
            sel = ti[:, 0] == b
            self.cluster_eval_offset.update(predictions["inst_labels_est"][sel], input_dict["inst_labels"][sel, 0],
                                            input_dict["sd_labels"][sel, 0])

    # -- libs/loss.py:248-258 -----------------------------------------------------------------------------------------
    def get_tpointnet_loss(self, predictions):
        total = 0
        n_iter = len(predictions["tpointnet_loss_terms"])
        for n_th, (_, v) in enumerate(predictions["tpointnet_loss_terms"].items(), 1):
            pose = self.w_obj_trans_loss * v["trans_loss"] + self.w_obj_rot_loss * v["rot_loss"]
            total = total + (self.w_obj_l1_loss * v["l1_loss"] + self.w_obj_pose_loss * pose) * self.obj_gamma ** (n_iter - n_th)
        return total

    @staticmethod
    def _metric(o):
        """compute_iou (libs/loss.py:17-48) from the device counters: arrays over the two classes, in thousands."""
        o = np.asarray(o, dtype=np.float64)
        return {"intersection": o[2:4] / 1e3, "union": o[6:8] / 1e3 + o[8:10] / 1e3 - o[2:4] / 1e3, "pred_positives": o[6:8] / 1e3,
                "gt_positives": o[8:10] / 1e3}

    # -- libs/loss.py:273-320 -----------------------------------------------------------------------------------------
    def forward(self, predictions, input_dict):
        if not predictions["fb_seg_est"].is_cuda:
            raise RuntimeError("pcaccumulation_b200.FuseLoss is CUDA-only (no CPU fallback)")
        stats = {}
        ego_l1 = self.w_pose_l1_loss * predictions["ego_l1_loss"]
        total = ego_l1
        stats["ego_l1_loss"] = ego_l1
        for k in ("ego_l2_loss", "ego_rot_error", "ego_trans_error"):
            stats[k] = predictions[k]
        perm_loss = _PermLoss.apply(_stack_perm(predictions["perm_matrix"])) * self.w_perm_loss
        total = total + perm_loss
        stats["perm_loss"] = perm_loss

        fb_loss, fb_out = self._fb_term(predictions)
        total = total + fb_loss
        stats["fb_loss"] = fb_loss
        mos_loss, mos_out = self._mos_term(predictions, input_dict)
        total = total + mos_loss
        stats["mos_loss"] = mos_loss

        offset_loss, off_out, gt_offset = self._offset_term(input_dict, predictions)
        total = total + offset_loss
        stats["offset_loss"] = offset_loss
        stats["offset_l1_loss"], stats["offset_dir_loss"] = off_out[0], off_out[1]
        # the reference stores the FG rows of the GT offsets as a side effect (libs/loss.py:241); kept lazy: the rows of all
        # points + the mask, compacted on access
        predictions["offset_gt_all"] = gt_offset

        if "tpointnet_loss_terms" in predictions:
            obj_loss = self.get_tpointnet_loss(predictions) * self.w_obj_loss
            total = total + obj_loss
            stats["obj_loss"] = obj_loss
            stats["inst_l2_error"] = predictions["inst_l2_error"]
            stats["dynamic_inst_l2_error"] = predictions["dynamic_inst_l2_error"]
        stats["loss"] = total
        # the python-number outputs of the reference, with one synchronisation
        host = torch.cat((fb_out, mos_out, off_out)).tolist()
        stats["fb_metric"] = self._metric(host[0:13])
        stats["mos_metric"] = self._metric(host[13:26])
        stats["offset_l2_error"] = host[26 + 2] if host[26 + 3] > 0 else 0
        if host[26 + 3] > 0:
            predictions["offset_gt"] = gt_offset[_i64(input_dict["fb_labels"][:, 0]) == 1]
        return stats
