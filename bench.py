#!/usr/bin/env python
"""Benchmark: scenes/sec of the per-scene hot path (voxelise -> MotionNet.forward, test mode) on B200.

Contract (one JSON line on rank 0):  python bench.py --gpus N --steps K --warmup W [--impl reference]
  * a "step" is one scene (BASELINE.json configs[1]: Waymo-shaped 5 x ~150k points, grid 288x288) through the
    full pipeline; at N>1 each rank processes its own scenes (no data-path collective, weak scaling) and the
    time is the max over ranks (NCCL all-reduce of the CUDA-event time);
  * ``value``: raw points already resident in HBM when the timed region starts;
  * ``e2e``: the same through ``SceneRunner.run_host`` with pinned HOST buffers (H2D of the points and D2H of
    the per-point results inside the timed region);
  * ``roofline``: all conv3x3 launches (the dominant kernels), algorithmic FLOPs / CUDA-event time measured
    inside the timed region, against the measured dense tensor peak in MEASURED_PEAKS.json;
  * ``cpu_baseline`` / ``--impl reference``: the oracle restatement of the reference's CPU path
    (``oracle/oracle.py``: voxelise + collate + forward), timed on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pcaccumulation_b200 import config, fixture, synth  # noqa: E402

N_SCENES = 4  # distinct synthetic scenes cycled through the steps
CONV_DRAM_BYTES_PER_LAUNCH = {"C2": 32.53e6}  # measured with ncu (profiles/r1_conv_dram.csv), see the roofline block below


def env_int(name, default):
    return int(os.environ.get(name, default))


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region, sampled in-process through NVML every 100 ms
    (forking nvidia-smi five times a second perturbs a multi-threaded timed region; it stays as the fallback)."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []  # (sm_mhz, sm_max_mhz, [active reasons])
        self.stop_flag = False
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            idx = index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[index])
                except ValueError:
                    idx = index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksEventReasonSwPowerCap}
        self.rows.append((sm, self.sm_max, [k for k, b in bits.items() if mask & b]))

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            c = [x.strip() for x in out.split(",")]
            self.rows.append((float(c[0]), float(c[1]), [n for i, n in enumerate(self.NAMES) if c[2 + i].lower().startswith("active")]))

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.1 if self.nvml is not None else 0.5)

    def summary(self):
        rows = list(self.rows)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(r[0] for r in rows)
        reasons = sorted(set(x for r in rows for x in r[2]))
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": rows[0][1], "reasons": reasons, "samples": len(rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_scenes(workload, rank, n):
    return [synth.make_workload_scene(workload, scene_idx=rank * 100 + i) for i in range(n)]


def oracle_forward_fn(cfg, sd):
    from oracle import oracle

    orc = oracle.OracleMotionNet(cfg, sd)
    vg = cfg["voxel_generator"]

    def run(scene):
        pts4 = np.concatenate((scene["input_points"], scene["time_indice"]), 1).astype(np.float32)
        v = oracle.voxelize(pts4, vg["voxel_size"], vg["range"], vg["n_sweeps"])
        sample = dict(scene)
        sample.update(v)
        inp = synth.collate([sample])
        torch.manual_seed(42)
        return orc.forward(inp)

    return run


def fixture_weights(cfg):
    """Fixture state_dict from the package's own parameter template (names/shapes == reference)."""
    from pcaccumulation_b200.motionnet import MotionNet

    tmpl = MotionNet(cfg).state_dict()
    return fixture.fixture_state_dict(tmpl, 42)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port) on the host cores; rank 0 only."""
    if rank != 0:
        return
    cfg = config.workload_config(args.workload)
    sd = fixture_weights(cfg)
    torch.set_num_threads(min(32, os.cpu_count()))  # 32 threads is the fastest setting measured for this path on the 128-core box
    run = oracle_forward_fn(cfg, sd)
    scenes = make_scenes(args.workload, 0, min(N_SCENES, 2))
    steps = min(args.steps, 10)
    warm = min(args.warmup, 1)
    for i in range(warm):
        run(scenes[i % len(scenes)])
    t0 = time.perf_counter()
    for i in range(steps):
        run(scenes[i % len(scenes)])
    dt = time.perf_counter() - t0
    val = steps / dt
    line = {
        "impl": "reference", "metric": "scenes/sec", "value": val, "unit": "scenes/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {workload_desc(args.workload)}", "batch": 1, "mode": "test"},
        "cpu_baseline": {"value": val, "unit": "scenes/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{steps} scenes of {args.workload} (voxelise+collate+forward), oracle/oracle.py"},
        "e2e": {"value": val, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_desc(name):
    w = config.WORKLOADS[name]
    cfg = config.workload_config(name)
    r = cfg["voxel_generator"]["range"]
    g = int(round((r[3] - r[0]) / cfg["voxel_generator"]["voxel_size"][0]))
    return f"{w['dataset']}-shaped {w['T']}x~{w['pts_per_frame'] // 1000}k pts, grid {g}x{g}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tc", action="store_true", help="force the FP32 CUDA-core convolution path")
    ap.add_argument("--in-flight", type=int, default=4, help="independent scenes kept in flight per GPU (1 = serial)")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        return run_reference(args, rank, world)
    args.warmup = max(args.warmup, 3)

    torch.set_num_threads(1)  # the GPU arm's host logic is single-threaded (OpenMP fan-out only slows torch.randperm)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from pcaccumulation_b200.runner import ScenePipeline, SceneRunner, scene_to_points4

    cfg = config.workload_config(args.workload)
    runner = SceneRunner(cfg, device=dev)
    model = runner.model
    sd = fixture.fixture_state_dict(model.state_dict(), 42)
    model.load_state_dict(sd)
    model.use_tensor_cores = not args.no_tc
    runner.warmup()
    pipe = ScenePipeline(cfg, depth=max(1, args.in_flight), device=dev)
    for r in pipe.runners:
        r.model.use_tensor_cores = not args.no_tc
    pipe.load_state_dict(sd)
    scenes = make_scenes(args.workload, rank, N_SCENES)
    host_pts = [torch.from_numpy(scene_to_points4(s)).pin_memory() for s in scenes]
    host_ego = [torch.from_numpy(s["ego_motion_gt"])[None].contiguous().pin_memory() for s in scenes]
    dev_pts = [p.to(dev) for p in host_pts]
    dev_ego = [e.to(dev) for e in host_ego]
    nums = [[p.shape[0]] for p in host_pts]
    out_bufs = [{"rec_est": torch.empty(p.shape[0], 3).pin_memory(), "fb": torch.empty(p.shape[0], dtype=torch.int64).pin_memory(),
                 "mos": torch.empty(p.shape[0], 2).pin_memory(), "inst": torch.empty(p.shape[0], dtype=torch.int64).pin_memory(),
                 "ego": torch.empty(1, cfg["voxel_generator"]["n_sweeps"], 4, 4).pin_memory()} for p in host_pts]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev(i):
        torch.manual_seed(1000 + i)
        return runner.run_device(dev_pts[i % N_SCENES], nums[i % N_SCENES], ego_motion_gt=dev_ego[i % N_SCENES])

    def step_host(i):
        torch.manual_seed(1000 + i)
        k = i % N_SCENES
        return runner.run_host(host_pts[k], nums[k], ego_motion_gt_host=host_ego[k], out=out_bufs[k])

    def timed(fn, steps):
        """K serial steps on the current stream (used for the per-kernel roofline events)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def timed_pipeline(steps, host):
        """EXACTLY K scenes through the pipeline (``in_flight`` of them concurrently), device time start -> last result."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        futs = []
        for i in range(steps):
            k = i % N_SCENES
            if host:
                futs.append(pipe.submit(host_pts[k], nums[k], ego=host_ego[k], seed=1000 + i, out=out_bufs[i % len(out_bufs)], host=True))
            else:
                futs.append(pipe.submit(dev_pts[k], nums[k], ego=dev_ego[k], seed=1000 + i))
        cur = torch.cuda.current_stream()
        for f in futs:
            _, done = f.result()
            cur.wait_event(done)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # one set of pinned result buffers per in-flight scene and per distinct scene size
    out_bufs = out_bufs * max(1, args.in_flight)
    out_bufs = [{k: torch.empty_like(v).pin_memory() for k, v in o.items()} for o in out_bufs]

    sampler = ClockSampler(local)
    sampler.start()  # started before the warm-up: its first fork/exec of nvidia-smi stays outside the timed regions
    for i in range(args.warmup):
        step_dev(i)
        step_host(i)
    # every slot (stream + caching-allocator pool) has to see every scene size before its pool stops growing
    # (cudaMalloc synchronises the device): warm up with two full cycles of the scenes through the slots
    n_warm = max(args.warmup, 2 * N_SCENES * max(1, args.in_flight))
    timed_pipeline(n_warm, host=False)
    timed_pipeline(n_warm, host=True)
    sampler.rows.clear()
    ms_dev = timed_pipeline(args.steps, host=False)
    ms_host = timed_pipeline(args.steps, host=True)
    # serial pass on one stream: latency of one scene and CUDA-event brackets around every conv launch (roofline)
    model.conv_events = []
    ms_serial = timed(step_dev, args.steps)
    events = model.conv_events
    model.conv_events = None
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # kernel launch census of one step (our kernels only: everything that is not an ATen kernel)
    launches_per_step = None
    try:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step_dev(0)
            torch.cuda.synchronize()
        names = [e.key for e in prof.key_averages() for _ in range(e.count) if e.device_type == torch.autograd.DeviceType.CUDA]
        ours = [n for n in names if "at::" not in n and "Memcpy" not in n and "Memset" not in n]
        launches_per_step = len(ours)
    except Exception:
        pass


    # roofline of the convolution kernels (dominant): algorithmic FLOPs / event time
    tot_flops = sum(e[2] for e in events)
    tot_ms = sum(e[0].elapsed_time(e[1]) for e in events)
    tot_bytes = sum(e[4] for e in events)
    paths = sorted(set(e[3] for e in events))
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
    achieved = tot_flops / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
    n_conv = len(events) / max(args.steps, 1)

    value = world * args.steps / (ms_dev * 1e-3)
    e2e = world * args.steps / (ms_host * 1e-3)
    h2d = int(np.mean([p.numel() * 4 + e.numel() * 4 for p, e in zip(host_pts, host_ego)]))
    d2h = int(np.mean([sum(t.numel() * t.element_size() for t in o.values()) for o in out_bufs]))

    cpu_base = None
    epe = None
    epe_fp32 = None
    if rank == 0 and not args.no_cpu_baseline:
        torch.set_num_threads(min(32, os.cpu_count()))  # 32 threads is the fastest setting measured for this path on the 128-core box
        run = oracle_forward_fn(cfg, sd)
        run(scenes[0])  # warm-up
        n_cpu = 2
        t0 = time.perf_counter()
        refs = [run(scenes[i]) for i in range(n_cpu)]
        dt = time.perf_counter() - t0
        cpu_base = {"value": n_cpu / dt, "unit": "scenes/s", "cores": torch.get_num_threads(), "kind": "port",
                    "sample": f"{n_cpu} scenes of {args.workload} (voxelise+collate+forward) after 1 warm-up, oracle/oracle.py"}
        # parity half of the metric: the free-running CUDA forward against the oracle on the same scenes and seed.
        # Labels are integer decisions (bit-exact target); a point whose two logits tie to ~1e-6 can flip between FP32
        # summation orders, and the TubeNet poses of the instance it joins then differ (random-weight fixture), so the
        # EPE is reported as median / fraction of points within 1 mm next to the mean.
        def compare(use_tc):
            fb_mis, mos_mis, inst_mis, pose_err, epe_mean, epe_med, within = 0, 0, 0, 0.0, [], [], []
            n_pts = 0
            model.use_tensor_cores = use_tc
            for i in range(n_cpu):
                torch.manual_seed(42)
                res = runner.run_device(dev_pts[i], nums[i], ego_motion_gt=dev_ego[i])
                ref = refs[i]
                n_pts += dev_pts[i].shape[0]
                fb_mis += int((res["fb_est_per_points"].cpu() != ref["fb_est_per_points"]).sum())
                mos_mis += int((res["mos_est"].cpu().argmax(1) != ref["mos_est"].argmax(1)).sum())
                if "inst_labels_est" in ref:
                    inst_mis += int((res["inst_labels_est"].cpu() != ref["inst_labels_est"]).sum())
                pose_err = max(pose_err, float((res["ego_motion_est"].cpu() - ref["ego_motion_est"]).abs().max()))
                d = (res["rec_est"].cpu() - ref["rec_est"]).norm(dim=1)
                epe_mean.append(float(d.mean()))
                epe_med.append(float(d.median()))
                within.append(float((d < 1e-3).float().mean()))
            model.use_tensor_cores = not args.no_tc
            return {"scenes": n_cpu, "points": n_pts, "fb_label_mismatches": fb_mis, "mos_label_mismatches": mos_mis,
                    "inst_label_mismatches": inst_mis, "ego_pose_max_abs_err": pose_err, "epe_mean_m": float(np.mean(epe_mean)),
                    "epe_median_m": float(np.mean(epe_med)), "frac_points_within_1mm": float(np.mean(within))}

        # free-running (nothing injected) on the timed path, and on the FP32 CUDA-core path: the latter isolates what the
        # 3xTF32 tensor-core rounding (~3e-5 per conv stack) contributes through label ties
        epe = compare(not args.no_tc)
        epe_fp32 = compare(False) if not args.no_tc else None

    if rank == 0:
        line = {
            "metric": "scenes/sec", "value": value, "unit": "scenes/s", "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": ("f32 (tensor cores: fp16-pair split operands, 22 significant bits, FP32 accumulate)" if ("tc-f16pair" in paths or "tc-p16" in paths)
                                        else "f32 (tensor cores: 3xTF32 split, FP32 accumulate)" if "tc" in paths else "f32"),
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {workload_desc(args.workload)}", "batch": 1, "mode": "test",
                       "scenes_per_rank": N_SCENES, "scenes_in_flight": max(1, args.in_flight),
                       "serial_ms_per_scene": ms_serial / args.steps, "parallelism": f"dp{world} (scene sharding, no data-path collective)",
                       "l2": "per-step working set (>1 GB of activations) exceeds the 126 MB L2; no explicit flush",
                       "conv_path": paths},
            "e2e": {"value": e2e, "unit": "scenes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": (launches_per_step or 0) * args.steps,
            "clocks": sampler.summary(),
            "roofline": {"bound": "tensor", "kernel": "conv3x3 (all launches, %.0f per step)" % n_conv, "achieved": achieved,
                         "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         # DRAM bytes per conv launch from profiles/r1_conv_dram.csv (ncu dram__bytes_read+write summed over
                         # the 44 conv launches of one C2 step / 44; cold-cache replay, so an upper bound on the warm traffic)
                         "traffic": CONV_DRAM_BYTES_PER_LAUNCH.get(args.workload),
                         "algorithmic_bytes_per_launch": tot_bytes / max(len(events), 1),
                         "algorithmic_flops_per_launch": tot_flops / max(len(events), 1),
                         "peak_source": peak_src,
                         "conv_ms_per_step": tot_ms / max(args.steps, 1), "conv_share_of_step": tot_ms / ms_serial},
            "cpu_baseline": cpu_base,
            "parity_vs_oracle": epe,
            "parity_vs_oracle_fp32_path": epe_fp32,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
