"""Evaluation tail of the reference's test loop on the GPU.

Mirrors ``libs/tester.py:58-107``: per-point end-point error and relative error of the accumulated cloud against the
ground-truth accumulation (``ego_motion_compensation`` + ``reconstruct_sequence`` with the GT instance motions,
``toolbox/register_utils.py:59-93``), restricted to ``time_indice > 0``, plus the motion-segmentation IoU counters
(``libs/loss.py:17-48,139-149``) and the scene-flow threshold metrics of ``toolbox/sf_eval_utils.py:46-52,71-100``.
One fused kernel per scene (``pcab_flow_eval``) instead of a dozen full-size torch ops and five host copies; the
counters stay on the device until ``summary()``.
"""
import numpy as np
import torch

from . import _lib as L
from ._lib import I, P, call, stream

CATEGORIES = ("all", "dynamic", "static")
SF_KEYS = ("EPE3D", "Acc3DS", "Acc3DR", "Outlier", "ROutlier")


class FlowEvaluator:
    def __init__(self, n_frames, device="cuda", keep_per_point=True):
        self.n_frames = int(n_frames)
        self.device = torch.device(device)
        self.sf = torch.zeros(3, 6, dtype=torch.float64, device=self.device)
        self.mos = torch.zeros(8, dtype=torch.int64, device=self.device)
        self.keep_per_point = keep_per_point
        self.dump = {k: [] for k in ("epe_per_point", "relative_error", "time_indice", "fb_label", "sd_label")}

    @torch.no_grad()
    def update(self, input_dict, predictions):
        """One scene (batch size 1, as ``libs/tester.py``).  Returns the per-point (epe, relative_error) device tensors."""
        dev = self.device
        pts = input_dict["input_points"].to(dev).float().contiguous()
        n = pts.shape[0]
        fast = input_dict.get("_pcab")
        tidx = fast["ptime"] if fast is not None else input_dict["time_indice"][:, 1].to(dev).to(torch.int32).contiguous()
        ego_gt = input_dict["ego_motion_gt"].to(dev).float()[0].contiguous()
        inst_gt = input_dict["inst_motion_gt"][0].to(dev).float().contiguous()
        i64 = lambda t: t.to(dev).reshape(-1).to(torch.int64).contiguous()
        inst, fb, sd = i64(input_dict["inst_labels"]), i64(input_dict["fb_labels"]), i64(input_dict["sd_labels"])
        rec = predictions["rec_est"].float().contiguous()
        mos_est = predictions["mos_est"].float().contiguous()
        fb_est = i64(predictions["fb_est_per_points"])
        epe = torch.empty(n, device=dev)
        rel = torch.empty(n, device=dev)
        call("pcab_flow_eval", P(pts), P(tidx), P(rec), P(ego_gt), P(inst), P(inst_gt), I(inst_gt.shape[0]), P(fb), P(sd),
             P(mos_est), P(fb_est), I(n), I(self.n_frames), P(epe), P(rel), P(self.sf), P(self.mos), stream())
        if self.keep_per_point:  # what libs/tester.py:81-85 appends (fp16 errors, int8 / bool labels), gathered on the device
            sel = tidx > 0
            self.dump["epe_per_point"].append(epe[sel].to(torch.float16))
            self.dump["relative_error"].append(rel[sel].to(torch.float16))
            self.dump["time_indice"].append(tidx[sel].to(torch.int8))
            self.dump["fb_label"].append(fb[sel] != 0)
            self.dump["sd_label"].append(sd[sel] != 0)
        return epe, rel

    def all_reduce(self):
        """Sum the counters over the ranks of a scene-sharded run (the only collective of the path: a 26-number all-reduce,
        NCCL on GPUs / gloo in the CPU tests).  The per-point dumps stay rank-local, as the reference's per-scene files do."""
        from .dist_utils import reduce_metrics
        if getattr(self, "_reduced", False):
            raise RuntimeError("FlowEvaluator.all_reduce() was already called: the counters hold the global sums")
        self.sf = reduce_metrics(self.sf, op="sum")
        self.mos = reduce_metrics(self.mos, op="sum")
        self._reduced = True
        return self

    def per_point_arrays(self):
        """The dict ``libs/tester.py:99-107`` saves with ``np.savez_compressed`` (one host copy per array, at the end)."""
        return {k: (torch.cat(v).cpu().numpy() if v else np.zeros(0)) for k, v in self.dump.items()}

    def summary(self):
        sf = self.sf.cpu().numpy()
        mos = self.mos.cpu().numpy()
        if int(mos[7]):
            raise IndexError(f"{int(mos[7])} points carried a frame index outside [0, n_frames) or an instance label without a "
                             "ground-truth motion (the reference's gathers raise at toolbox/register_utils.py:66,85)")
        out = {}
        for c, name in enumerate(CATEGORIES):
            n = max(sf[c, 0], 1.0)
            out[name] = {"count": int(sf[c, 0]), "EPE3D": sf[c, 1] / n, "Acc3DS": sf[c, 2] / n, "Acc3DR": sf[c, 3] / n,
                         "Outlier": sf[c, 4] / n, "ROutlier": sf[c, 5] / n}
        inter = np.array([mos[0], mos[3]], dtype=np.float64)
        pred = np.array([mos[1], mos[4]], dtype=np.float64)
        gt = np.array([mos[2], mos[5]], dtype=np.float64)
        union = pred + gt - inter
        out["mos"] = {"intersection": inter, "union": union, "pred_positives": pred, "gt_positives": gt,
                      "iou": inter / np.maximum(union, 1.0), "recall": inter / np.maximum(gt, 1.0),
                      "precision": inter / np.maximum(pred, 1.0), "masked_points": int(mos[6])}
        return out


class ClusterEvaluator:
    """``toolbox/cluster_eval.py:ClusterEvaluation`` (coverage + precision / recall of the predicted instances at IoU
    0.5 .. 0.9, per motion class) with the per-scene work on the device: one pass over the points builds the (est, gt)
    contingency table (``pcab_cluster_eval``); the reference loops over all instance pairs in Python with boolean masks."""

    THRESHOLDS = (0.5, 0.6, 0.7, 0.8, 0.9)

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self.counters = torch.zeros(28, dtype=torch.float64, device=self.device)

    @torch.no_grad()
    def update(self, inst_est, inst_gt, mos_label):
        """One scene: inst_est / inst_gt [N] (0 = background), mos_label [N] (0 static, 1 dynamic), as
        ``libs/loss.py:261-270`` passes them per batch element."""
        from ._lib import Z, scratch, size
        dev = self.device
        i64 = lambda t: t.to(dev).reshape(-1).to(torch.int64).contiguous()
        est, gt, mos = i64(inst_est), i64(inst_gt), i64(mos_label)
        n = est.shape[0]
        if n == 0:
            return
        max_est, max_gt = [int(v) for v in torch.stack((est.max(), gt.max())).tolist()]
        ws = scratch(size("pcab_cluster_eval_workspace", I(max_est), I(max_gt)), dev)
        call("pcab_cluster_eval", P(est), P(gt), P(mos), I(n), I(max_est), I(max_gt), P(self.counters), P(ws), Z(ws.numel()),
             stream())

    def all_reduce(self):
        from .dist_utils import reduce_metrics
        if getattr(self, "_reduced", False):
            raise RuntimeError("ClusterEvaluator.all_reduce() was already called: the counters hold the global sums")
        self.counters = reduce_metrics(self.counters, op="sum")
        self._reduced = True
        return self

    def summary(self):
        """The numbers ``ClusterEvaluation.final_eval`` logs (toolbox/cluster_eval.py:33-69), per class [static, dynamic]."""
        c = self.counters.cpu().numpy()
        scenes = np.maximum(c[[2, 6]], 1.0)
        out = {"MUCov": c[[0, 4]] / scenes, "MWCov": c[[1, 5]] / scenes, "total_gt_inst": c[[3, 7]]}
        for k, thr in enumerate(self.THRESHOLDS):
            tp = np.array([c[8 + 2 * (2 * k + cls)] for cls in (0, 1)])
            fp = np.array([c[8 + 2 * (2 * k + cls) + 1] for cls in (0, 1)])
            out[f"@{thr}"] = {"precision": tp / np.maximum(tp + fp, 1.0), "recall": tp / np.maximum(out["total_gt_inst"], 1.0),
                              "tp": tp, "fp": fp}
        return out
