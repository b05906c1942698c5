"""Scene runner: raw multi-sweep points -> GPU voxelise -> ``MotionNet.forward`` (the tester's inner loop).

Mirrors what ``libs/tester.py:52-56`` does per scene (DataLoader sample -> ``.to(device)`` -> ``model(input_dict)``)
except that voxelisation and collation run on the device (``libs/voxel_generator.py`` and
``libs/dataloader.py:7-40`` are CPU code in the reference).  ``run_host`` is the end-to-end entry used by the
benchmark: pinned host buffers in, host results out, with the H2D/D2H copies on the caller's stream.
"""
import queue
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from .motionnet import MotionNet
from .voxel_generator import Voxelization


def set_host_sync_mode(mode="blocking"):
    """How host threads wait for the GPU (readbacks, event waits): ``"blocking"`` sleeps on an interrupt
    (cudaDeviceScheduleBlockingSync), ``"spin"`` busy-waits (cudaDeviceScheduleSpin), ``"yield"`` yields its time slice.
    With several scenes in flight per GPU and several ranks per node, spinning waiters compete with the threads that have
    launches to issue: a rank with fewer host cores than (scenes in flight + 2) should block.  Must be called before the
    process creates its CUDA context (i.e. before the first CUDA call of torch).  Returns the cudart status (0 = applied)."""
    import ctypes

    flags = {"spin": 1, "yield": 2, "blocking": 4}[mode]
    try:
        cudart = ctypes.CDLL("libcudart.so.12")
    except OSError:
        cudart = ctypes.CDLL("libcudart.so")
    return int(cudart.cudaSetDeviceFlags(ctypes.c_uint(flags)))


class SceneRunner:
    def __init__(self, cfg, model=None, device="cuda"):
        self.cfg = cfg
        self.device = torch.device(device)
        self.model = (model or MotionNet(cfg)).to(self.device).eval()
        self.vox = Voxelization(cfg["voxel_generator"])
        vg = cfg["voxel_generator"]
        self.T = vg["n_sweeps"]
        g = self.vox.grid_size
        self.shape = torch.tensor([[int(g[0]), int(g[1]), int(g[2]), self.T]], dtype=torch.int64)

    def build_input(self, points4, num_points, labels=None, ego_motion_gt=None, inst_motion_gt=None, reference_schema=False):
        """points4: CUDA f32 [N,4] (x,y,z,t) with the scenes of a batch concatenated; num_points: list[int]."""
        dev = points4.device
        B = len(num_points)
        N = points4.shape[0]
        if B > 1:
            pbatch = torch.repeat_interleave(torch.arange(B, device=dev, dtype=torch.int32),
                                             torch.tensor(num_points, device=dev))
        else:
            if getattr(self, "_zeros32", None) is None or self._zeros32.shape[0] < N:
                self._zeros32 = torch.zeros(N, dtype=torch.int32, device=dev)
            pbatch = self._zeros32[:N]
        v = self.vox.voxelize_batch(points4, pbatch if B > 1 else None, B)
        if v["n_rejected"]:  # libs/dataset.py:218 rejects such samples
            raise ValueError("points outside the voxel range")
        ptime = points4[:, 3].to(torch.int32)
        labels = labels or {}
        T = self.T
        if getattr(self, "_zeros", None) is None or self._zeros.shape[0] < N:
            self._zeros = torch.zeros(N, 1, dtype=torch.int64, device=dev)
        zeros = self._zeros[:N]
        d = {
            "input_points": points4[:, :3].contiguous(),
            "num_points": torch.tensor(num_points, dtype=torch.int64),
            "sd_labels": labels.get("sd_labels", zeros),
            "inst_labels": labels.get("inst_labels", zeros),
            "fb_labels": labels.get("fb_labels", zeros),
            "ego_motion_gt": ego_motion_gt if ego_motion_gt is not None else torch.eye(4, device=dev).repeat(B, T, 1, 1),
            "inst_motion_gt": inst_motion_gt if inst_motion_gt is not None else [torch.eye(4).repeat(1, T, 1, 1)] * B,
            "num_voxels": v["num_voxels"],
            "shape": self.shape.repeat(B, 1),
            # the int32 device arrays the kernels consume (what the reference-schema f64/i64 entries encode)
            "_pcab": {"p2v": v["point_to_voxel_map"], "pbatch": pbatch, "ptime": ptime,
                      "coords_zyxt": v["coordinates"].contiguous(), "pillar_batch": v["pillar_batch"].contiguous()},
        }
        if reference_schema:  # exactly what libs/dataloader.py:collate_fn would have produced
            d["coordinates"] = torch.cat((v["pillar_batch"][:, None], v["coordinates"]), 1).double()
            d["time_indice"] = torch.stack((pbatch.double(), points4[:, 3].double()), 1)
            d["num_voxels"] = v["num_voxels"].to(torch.int64)
            d["point_to_voxel_map"] = v["point_to_voxel_map"].to(torch.int64)[:, None]
        return d

    def prep_raw(self, sample, aug=None):
        """Steps (1-)2-4 of ``libs/dataset.py:prep_input`` on the device.

        ``sample``: the arrays of one ``.npz`` file (numpy or torch) - ``raw_points`` f32 [N,3] and ``time_indice``,
        ``sd_labels``, ``fb_labels``, ``inst_labels`` [N].  ``aug``: ``dataset.sample_augmentation(...)`` for the training-time
        augmentation (step 1), None for the validation / test path.  Returns (points4 [M,4] f32, labels dict with [M,1] i64
        tensors, M): augmentation + scene crop + ground removal as one stable compaction.
        """
        import ctypes

        from ._lib import D, F, I, P, Z, call, scratch, size, stream
        dev = self.device
        raw = torch.as_tensor(sample["raw_points"]).to(dev).float().contiguous()
        n = raw.shape[0]
        # (named tensors: a temporary handed to P() would be freed -- and its block reused -- before the kernel reads it)
        ins = {k: torch.as_tensor(sample[k]).to(dev).reshape(-1).to(torch.int64).contiguous()
               for k in ("time_indice", "sd_labels", "fb_labels", "inst_labels")}
        i64 = ins.__getitem__
        vg, dc = self.cfg["voxel_generator"], self.cfg["data"]
        p4 = torch.empty(n, 4, device=dev)
        t32 = torch.empty(n, dtype=torch.int32, device=dev)
        outs = {k: torch.empty(n, dtype=torch.int64, device=dev) for k in ("sd_labels", "fb_labels", "inst_labels")}
        count = torch.empty(1, dtype=torch.int32, device=dev)
        ws = scratch(size("pcab_prep_points_workspace", I(n)), dev)
        ground = dc["ground_height"] + dc["ground_slack"]
        if aug is None:
            call("pcab_prep_points", P(raw), P(i64("time_indice")), P(i64("sd_labels")), P(i64("fb_labels")), P(i64("inst_labels")),
                 I(n), F(vg["crop_range"][0]), F(vg["crop_range"][1]), F(vg["crop_range"][2]), I(int(dc["remove_ground"])),
                 F(ground), P(p4), P(t32), P(outs["sd_labels"]), P(outs["fb_labels"]), P(outs["inst_labels"]), P(count), P(ws),
                 Z(ws.numel()), stream())
        else:
            tsfm = (ctypes.c_double * 16)(*[float(v) for v in np.asarray(aug["tsfm"], dtype=np.float64).reshape(-1)])
            noise = None if aug.get("noise") is None else torch.as_tensor(aug["noise"], dtype=torch.float64).to(dev).contiguous()
            assert noise is None or noise.shape == (n, 3)
            call("pcab_prep_points_augmented", P(raw), P(i64("time_indice")), P(i64("sd_labels")), P(i64("fb_labels")),
                 P(i64("inst_labels")), I(n), P(tsfm), P(noise), ctypes.c_ulonglong(int(aug.get("seed", 0))), D(aug["noise_amp"]),
                 D(aug["scale"]), D(vg["crop_range"][0]), D(vg["crop_range"][1]), D(vg["crop_range"][2]), I(int(dc["remove_ground"])),
                 D(ground), P(p4), P(t32), P(outs["sd_labels"]), P(outs["fb_labels"]), P(outs["inst_labels"]), P(count), P(ws),
                 Z(ws.numel()), stream())
        m = int(count.item())
        return p4[:m], {k: v[:m, None] for k, v in outs.items()}, m

    def prep_sample(self, sample, augment=False, exact_noise=True):
        """``BaseDataset.prep_input`` (libs/dataset.py:147-207) for one raw sample -> (points4, labels, M, ego_motion_gt
        [T,4,4], inst_motion_gt [K,T,4,4]); with ``augment`` the random numbers come from numpy's global stream in the
        reference's order (``dataset.sample_augmentation``) and the ground-truth motions are conjugated accordingly."""
        from . import dataset as ds

        ego, inst = np.asarray(sample["ego_motion_gt"]), np.asarray(sample["inst_motion_gt"])
        aug = None
        if augment:
            aug = ds.sample_augmentation(self.cfg["data_aug"], int(np.asarray(sample["raw_points"]).shape[0]), exact_noise)
            ego, inst = ds.update_transformation_after_data_augmentation(aug["tsfm"], ego, inst, self.T)
        p4, labels, m = self.prep_raw(sample, aug)
        return p4, labels, m, ego, inst

    @torch.no_grad()
    def run_raw(self, sample, augment=False):
        """Raw sample arrays (as stored by the reference's dataset writers) -> [augmentation ->] crop / ground removal ->
        voxelise -> forward."""
        p4, labels, m, ego, inst = self.prep_sample(sample, augment)
        ego = torch.as_tensor(ego).to(self.device).float()[None].contiguous()
        inst_motion = [torch.as_tensor(inst).to(self.device).float()]
        return self.model(self.build_input(p4, [m], labels=labels, ego_motion_gt=ego, inst_motion_gt=inst_motion))

    @torch.no_grad()
    def run_npz(self, paths, augment=False):
        """``.npz`` sample files (one per scene of the batch) -> device collate (libs/dataloader.py:7-40) -> forward."""
        from . import dataset as ds

        parts = [self.prep_sample(ds.load_sample(p), augment) for p in ([paths] if isinstance(paths, str) else paths)]
        p4 = torch.cat([q[0] for q in parts])
        labels = {k: torch.cat([q[1][k] for q in parts]) for k in parts[0][1]}
        ego = torch.stack([torch.as_tensor(q[3]).float() for q in parts]).to(self.device).contiguous()
        inst_motion = [torch.as_tensor(q[4]).to(self.device).float() for q in parts]
        return self.model(self.build_input(p4, [q[2] for q in parts], labels=labels, ego_motion_gt=ego, inst_motion_gt=inst_motion))

    def warmup(self, batch_size=1):
        """Capture the model's CUDA graphs (see ``MotionNet.warmup``); call after loading weights, before serving."""
        self.model.warmup(batch_size)

    @torch.no_grad()
    def run_device(self, points4, num_points, **kw):
        return self.model(self.build_input(points4, num_points, **kw))

    @torch.no_grad()
    def run_host(self, points4_pinned, num_points, ego_motion_gt_host=None, out=None, copy_stream=None):
        """End to end with HOST buffers: H2D of the raw points, forward, D2H of the per-point results.

        ``out``: optional dict of pinned host tensors (rec_est f32[N,3], fb i64[N], mos f32[N,2], inst i64[N],
        ego f32[B,T,4,4]) to receive the results.  ``copy_stream``: when given, the D2H copies are queued there (after an
        event on the current stream), so the next forward on the current stream does not wait for them; the caller then
        synchronises with ``copy_stream`` before reading ``out``.  Returns (results_on_device, bytes_h2d, bytes_d2h).
        """
        dev = self.device
        pts = points4_pinned.to(dev, non_blocking=True)
        h2d = points4_pinned.numel() * 4
        ego = None
        if ego_motion_gt_host is not None:
            ego = ego_motion_gt_host.to(dev, non_blocking=True)
            h2d += ego_motion_gt_host.numel() * 4
        res = self.run_device(pts, num_points, ego_motion_gt=ego)
        d2h = 0
        if out is not None:
            srcs = [(key, src) for key, src in (("rec_est", res["rec_est"]), ("fb", res["fb_est_per_points"][:, 0]),
                                                ("mos", res["mos_est"]), ("inst", res.get("inst_labels_est")),
                                                ("ego", res["ego_motion_est"])) if src is not None and key in out]
            cur = torch.cuda.current_stream()
            if copy_stream is not None:
                copy_stream.wait_stream(cur)
            with torch.cuda.stream(copy_stream if copy_stream is not None else cur):
                for key, src in srcs:
                    out[key].copy_(src, non_blocking=True)
                    if copy_stream is not None:
                        src.record_stream(copy_stream)
                    d2h += src.numel() * src.element_size()
        return res, h2d, d2h


class ScenePipeline:
    """Keeps ``depth`` independent scenes in flight on ONE GPU.

    Scenes never interact, but one forward has ~7 host round trips (pillar count, background
    counts for the keypoint draw, FG / instance counts) and many small latency-bound kernels (ego pairs, DBSCAN
    union-find, the 18x18 / 36x36 maps).  Each slot owns a host thread, a CUDA stream, a ``SceneRunner`` (its own
    activations, pinned staging and keypoint generator; weights are loaded per slot, 44 MB) so that the bubbles of
    one scene are filled with the kernels of another.  ``depth=1`` is the plain serial runner on a side stream.
    """

    def __init__(self, cfg, state_dict=None, depth=2, device="cuda"):
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.depth = int(depth)
        self.runners = [SceneRunner(cfg, device=self.device) for _ in range(self.depth)]
        self._free = queue.SimpleQueue()
        self._slots = []
        for r in self.runners:
            r.model.rng = torch.Generator()  # per-slot CPU generator: the global one would interleave between threads
            self._slots.append((r, torch.cuda.Stream(device=self.device)))
            self._free.put(self._slots[-1])
            r.copy_stream = torch.cuda.Stream(device=self.device)  # D2H of a finished scene overlaps the slot's next forward
        if state_dict is not None:
            self.load_state_dict(state_dict)
        # torch initialises its linear-algebra backend lazily and not thread-safely: touch it once from this thread
        eye = torch.eye(4, device=self.device).repeat(2, 1, 1)
        torch.linalg.inv(eye)
        torch.linalg.svd(eye[:, :3, :3])
        torch.cuda.synchronize(self.device)
        self._pool = ThreadPoolExecutor(max_workers=self.depth, thread_name_prefix="pcab-scene")

    def load_state_dict(self, state_dict, batch_size=1):
        """Load the weights into every slot and capture its CUDA graphs (single-threaded: the workers are idle)."""
        for r, stream in self._slots:
            r.model.load_state_dict(state_dict)
            with torch.cuda.stream(stream):
                r.warmup(batch_size)

    def _work(self, points4, num_points, ego, seed, out, host, post=None, keep=True):
        runner, stream = self._free.get()
        try:
            torch.cuda.set_device(self.device)
            with torch.cuda.stream(stream):
                if seed is not None:
                    runner.model.rng.manual_seed(int(seed))
                done = torch.cuda.Event()
                if host:
                    res = runner.run_host(points4, num_points, ego_motion_gt_host=ego, out=out, copy_stream=runner.copy_stream)[0]
                    if post is not None:  # the completion event then covers the post-processing on this stream too
                        post(res)
                        runner.copy_stream.wait_stream(stream)
                    done.record(runner.copy_stream)
                else:
                    res = runner.run_device(points4, num_points, ego_motion_gt=ego)
                    if post is not None:
                        post(res)
                    done.record(stream)
            return (res if keep else None), done
        finally:
            self._free.put((runner, stream))

    def submit(self, points4, num_points, ego=None, seed=None, out=None, host=False, post=None, keep_results=True):
        """Queue one scene; returns a future of ``(results, cuda_event)``.  The results live on the slot's stream:
        wait for the event (``event.synchronize()`` or ``stream.wait_event``) before reading them.  ``post(results)``: optional
        callable run on the slot's thread and stream right after the forward (e.g. the Chamfer alignment errors).
        ``keep_results=False`` returns ``(None, event)``: the device tensors of the scene go back to the allocator at once
        (a long queue of futures otherwise pins every scene's outputs)."""
        return self._pool.submit(self._work, points4, num_points, ego, seed, out, host, post, keep_results)

    def close(self):
        self._pool.shutdown(wait=True)


def scene_to_points4(scene):
    return np.concatenate((scene["input_points"], scene["time_indice"].astype(np.float32)), 1).astype(np.float32)
