#!/usr/bin/env python
"""Benchmark: scenes/sec of the per-scene hot path (voxelise -> MotionNet.forward, test mode) on B200.

Contract (one JSON line on rank 0):  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C2]
  * a "step" is one BATCH of ``--scenes-per-step`` (default 32) independent scenes of the workload through the full
    pipeline (voxelise + forward), ``--in-flight`` of them at a time per GPU; W warm-up steps, then EXACTLY K timed steps
    (K = 20 -> 640 scenes, > 2 s of device time).  At N > 1 each rank processes its own scenes (no data-path collective,
    weak scaling) and the time is the max over ranks (NCCL all-reduce of the CUDA-event time);
  * workloads = BASELINE.json configs: C2 (default, the configuration the metric is quoted on) Waymo-shaped 5 x ~150k points,
    288^2; C3 nuScenes-shaped 10 x 35k + the Chamfer alignment errors of models/tpointnet.py:145-163 on the full cloud inside
    the timed region; C4 = C2 scenes as B = 4 forwards (32 scenes over 8 ranks); C5 5 x 400k points, 512^2.  The default run
    also takes a SHORT measurement of C3 / C4 / C5 (``other_configs``) so that every configuration is on the record;
  * ``value``: raw points already resident in HBM when the timed region starts;
  * ``e2e``: the same through ``ScenePipeline.submit(host=True)`` with pinned HOST buffers (H2D of the points and D2H of the
    per-point results inside the timed region);
  * ``roofline``: the tcgen05 convolution kernel (k_conv_p16: conv3x3 + ConvTranspose), algorithmic FLOPs of the two
    convolution stacks / CUDA-event time of their graph replays, against the measured dense tensor peak in MEASURED_PEAKS.json;
  * ``parity``: the staged protocol of oracle/protocol.py on one scene of the workload (float32 oracle = the reference's
    arithmetic, float64 oracle = measured rounding floor); the run FAILS on a label flip that is not a rounding tie;
  * ``cpu_baseline`` / ``--impl reference``: the oracle restatement of the reference's CPU path (``oracle/oracle.py``:
    numba voxeliser + collate + forward), timed on this box's host cores with 16 threads everywhere.
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
CPU_THREADS = 16  # host threads of the CPU arm: the same on every box (the smallest box of the pool exposes 16)
WORKLOADS = {"C1": ("C1", 1), "C2": ("C2", 1), "C3": ("C3", 1), "C4": ("C2", 4), "C5": ("C5", 1)}  # name -> (scene config, batch)


def env_int(name, default):
    return int(os.environ.get(name, default))


def cpu_threads():
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    return max(1, min(CPU_THREADS, avail))


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


def make_scenes(scene_cfg, rank, n):
    return [synth.make_workload_scene(scene_cfg, scene_idx=rank * 100 + i) for i in range(n)]


def oracle_forward_fn(cfg, sd):
    """The CPU arm: numba voxeliser + collate + float32 forward of the oracle restatement (the only place outside tests /
    smoke() where bench.py executes oracle/)."""
    from oracle import oracle

    orc = oracle.OracleMotionNet(cfg, sd)
    vg = cfg["voxel_generator"]

    def run(scenes):
        samples = []
        for scene in scenes:
            pts4 = np.concatenate((scene["input_points"], scene["time_indice"]), 1).astype(np.float32)
            sample = dict(scene)
            sample.update(oracle.voxelize_sequential(pts4, vg["voxel_size"], vg["range"], vg["n_sweeps"]))
            samples.append(sample)
        inp = synth.collate(samples)
        torch.manual_seed(42)
        return orc.forward(inp), inp

    return run


def fixture_weights(cfg):
    """Fixture state_dict from the package's own parameter template (names/shapes == reference)."""
    from pcaccumulation_b200.motionnet import MotionNet

    return fixture.fixture_state_dict(MotionNet(cfg).state_dict(), 42)


def workload_desc(name):
    scene_cfg, batch = WORKLOADS[name]
    w = config.WORKLOADS[scene_cfg]
    cfg = config.workload_config(scene_cfg)
    r = cfg["voxel_generator"]["range"]
    g = int(round((r[3] - r[0]) / cfg["voxel_generator"]["voxel_size"][0]))
    extra = " + Chamfer alignment errors (n = m = all points)" if name == "C3" else ""
    return f"{name}: {w['dataset']}-shaped {w['T']}x~{w['pts_per_frame'] // 1000}k pts, grid {g}x{g}, B={batch} per forward{extra}"


def cpu_alignment_errors(scene, res):
    """models/tpointnet.py:145-163 on the CPU arm: brute-force Chamfer of the reference's CPU extension, restated in numpy."""
    from oracle import oracle

    pts = torch.from_numpy(scene["input_points"]).float()
    t = torch.from_numpy(scene["time_indice"][:, 0]).long()
    est, gt = res["ego_motion_est"][0], res["ego_motion_gt"][0]
    a = oracle.ego_motion_compensation(pts, t, est)
    b = oracle.ego_motion_compensation(pts, t, gt)
    d1, d2, _, _ = oracle.chamfer(b[None].numpy(), a[None].numpy())
    w = (t == 1).float()
    w = w / (w.sum() + 1e-20)
    return float(((torch.from_numpy(d1[0]) * w).sum() + (torch.from_numpy(d2[0]) * w).sum()) / 2)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port) on the host cores; rank 0 only."""
    if rank != 0:
        return
    scene_cfg, batch = WORKLOADS[args.workload]
    cfg = config.workload_config(scene_cfg)
    sd = fixture_weights(cfg)
    torch.set_num_threads(cpu_threads())
    run = oracle_forward_fn(cfg, sd)
    scenes = make_scenes(scene_cfg, 0, max(2, batch))
    steps = min(args.steps, 6)  # a bounded sample: one step here = ONE forward of `batch` scenes (~2-4 s each)
    warm = min(args.warmup, 1)
    chamfer = args.workload == "C3"
    pick = lambda i: [scenes[(i * batch + j) % len(scenes)] for j in range(batch)]

    def step(i):
        res, _ = run(pick(i))
        if chamfer:
            cpu_alignment_errors(pick(i)[0], res)

    for i in range(warm):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    dt = time.perf_counter() - t0
    val = steps * batch / dt
    sample = (f"{steps} forwards of {batch} scene(s) of {args.workload} (numba voxelise + collate + forward"
              f"{' + brute-force Chamfer' if chamfer else ''}) after {warm} warm-up, oracle/oracle.py (port of the reference's CPU path)")
    line = {
        "impl": "reference", "metric": "scenes/sec", "value": val, "unit": "scenes/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(args.workload), "batch": batch, "mode": "test"},
        "cpu_baseline": {"value": val, "unit": "scenes/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


class Arm:
    """One workload on this rank's GPU: model replicas, pipeline, resident and pinned inputs."""

    def __init__(self, name, rank, dev, in_flight, no_tc=False, operands=None):
        from pcaccumulation_b200.alignment import BaseModel
        from pcaccumulation_b200.runner import ScenePipeline, SceneRunner, scene_to_points4

        self.name = name
        scene_cfg, self.batch = WORKLOADS[name]
        self.cfg = config.workload_config(scene_cfg)
        self.dev = dev
        self.runner = SceneRunner(self.cfg, device=dev)
        self.model = self.runner.model
        self.sd = fixture.fixture_state_dict(self.model.state_dict(), 42)
        self.model.load_state_dict(self.sd)
        self.model.use_tensor_cores = not no_tc
        if operands:
            self.model.conv_operands = operands
        self.runner.warmup(self.batch)
        self.pipe = ScenePipeline(self.cfg, depth=max(1, in_flight), device=dev)
        for r in self.pipe.runners:
            r.model.use_tensor_cores = not no_tc
            if operands:
                r.model.conv_operands = operands
        self.pipe.load_state_dict(self.sd, self.batch)
        self.scenes = make_scenes(scene_cfg, rank, N_SCENES)
        B = self.batch
        groups = [[self.scenes[(k + j) % N_SCENES] for j in range(B)] for k in range(N_SCENES)]
        self.host_pts = [torch.from_numpy(np.concatenate([scene_to_points4(s) for s in g])).pin_memory() for g in groups]
        self.host_ego = [torch.from_numpy(np.stack([s["ego_motion_gt"] for s in g])).contiguous().pin_memory() for g in groups]
        self.nums = [[s["input_points"].shape[0] for s in g] for g in groups]
        self.dev_pts = [p.to(dev) for p in self.host_pts]
        self.dev_ego = [e.to(dev) for e in self.host_ego]
        T = self.cfg["voxel_generator"]["n_sweeps"]
        depth = max(1, in_flight)
        self.out_bufs = [{"rec_est": torch.empty(p.shape[0], 3).pin_memory(), "fb": torch.empty(p.shape[0], dtype=torch.int64).pin_memory(),
                          "mos": torch.empty(p.shape[0], 2).pin_memory(), "inst": torch.empty(p.shape[0], dtype=torch.int64).pin_memory(),
                          "ego": torch.empty(B, T, 4, 4).pin_memory()} for p in self.host_pts for _ in range(depth)]
        self.align = BaseModel(self.cfg) if name == "C3" else None
        self.h2d = int(np.mean([p.numel() * 4 + e.numel() * 4 for p, e in zip(self.host_pts, self.host_ego)]))
        self.d2h = int(np.mean([sum(t.numel() * t.element_size() for t in o.values()) for o in self.out_bufs]))

    def post(self, k):
        """C3: the Chamfer / L2 alignment errors of the estimated ego poses on the full cloud (models/tpointnet.py:145-163)."""
        if self.align is None:
            return None
        pts, ego_gt = self.dev_pts[k], self.dev_ego[k]

        def fn(res):
            cd, l2 = self.align.get_alignment_errors(pts[:, :3], pts[:, 3], res["ego_motion_est"][0], ego_gt[0])
            res["alignment_errors"] = torch.stack((cd, l2))

        return fn

    def submit(self, i, host):
        k = i % N_SCENES
        if host:
            # (every scene size cycles through `depth` pinned result buffers, so a buffer is never rewritten while in flight)
            depth = len(self.out_bufs) // N_SCENES
            out = self.out_bufs[k * depth + (i // N_SCENES) % depth]
            return self.pipe.submit(self.host_pts[k], self.nums[k], ego=self.host_ego[k], seed=1000 + i, out=out, host=True,
                                    post=self.post(k), keep_results=False)
        return self.pipe.submit(self.dev_pts[k], self.nums[k], ego=self.dev_ego[k], seed=1000 + i, post=self.post(k), keep_results=False)

    def serial(self, i):
        torch.manual_seed(1000 + i)
        k = i % N_SCENES
        res = self.runner.run_device(self.dev_pts[k], self.nums[k], ego_motion_gt=self.dev_ego[k])
        fn = self.post(k)
        if fn is not None:
            fn(res)
        return res

    def close(self):
        self.pipe.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--scenes-per-step", type=int, default=32, help="independent scenes per step (one step = one batch of scenes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short C3 / C4 / C5 measurements of the default run")
    ap.add_argument("--no-tc", action="store_true", help="force the FP32 CUDA-core convolution path")
    ap.add_argument("--operands", default=None, choices=["f16", "tf32"], help="operand format of the tensor-core convolutions")
    ap.add_argument("--in-flight", type=int, default=0,
                    help="independent scenes kept in flight per GPU (1 = serial; 0 = auto: 4, or 3 when the rank has fewer than 6 host cores)")
    args = ap.parse_args()
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.in_flight <= 0:
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        args.in_flight = 4 if cores // max(1, env_int("LOCAL_WORLD_SIZE", world)) >= 6 else 3
    if args.warmup < 3:
        print(f"bench.py: --warmup {args.warmup} raised to 3 (timing rules)", file=sys.stderr)
        args.warmup = 3

    torch.set_num_threads(1)  # the GPU arm's host logic is single-threaded (OpenMP fan-out only slows torch.randperm)
    if os.environ.get("PCAB_SWITCH_INTERVAL"):
        sys.setswitchinterval(float(os.environ["PCAB_SWITCH_INTERVAL"]))
    # how host threads wait for the GPU.  Measured on a B200 box restricted to 4 host cores (what a rank gets at 8 ranks on 32
    # cores): spinning (the driver's default) 262 scenes/s, yielding 262, interrupt-driven blocking 234-242 -- the wake-up
    # latency of ~5 readbacks per scene costs more than the spinning threads do, so the default stays; PCAB_HOST_SYNC overrides.
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    cores_per_rank = cores // max(1, env_int("LOCAL_WORLD_SIZE", world))
    sync_mode = os.environ.get("PCAB_HOST_SYNC", "auto")
    if sync_mode == "auto":
        sync_mode = "default"
    if sync_mode != "default":
        from pcaccumulation_b200.runner import set_host_sync_mode

        rc = set_host_sync_mode(sync_mode)
        if rc != 0:
            print(f"bench.py: cudaSetDeviceFlags({sync_mode}) -> {rc}", file=sys.stderr)
    args.host_sync = sync_mode
    args.cores_per_rank = cores_per_rank
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(arm, steps, per_step, host):
        """EXACTLY `steps` steps of `per_step` forwards through the pipeline, device time from the first submit to the last result."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        futs = [arm.submit(i, host) for i in range(steps * per_step)]
        cur = torch.cuda.current_stream()
        for f in futs:
            _, done = f.result()
            cur.wait_event(done)
        e1.record()
        barrier()
        return reduce_max(e0.elapsed_time(e1))

    def timed_serial(arm, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            arm.serial(i)
        e1.record()
        barrier()
        return reduce_max(e0.elapsed_time(e1)) / n

    def measure(arm, steps, warmup, scenes_per_step):
        per_step = max(1, scenes_per_step // arm.batch)  # forwards per step
        for i in range(3):
            arm.serial(i)
        for host in (False, True):
            timed_steps(arm, warmup, per_step, host)
        ms_dev = timed_steps(arm, steps, per_step, False)
        ms_host = timed_steps(arm, steps, per_step, True)
        scenes = steps * per_step * arm.batch
        return {"value": world * scenes / (ms_dev * 1e-3), "e2e": world * scenes / (ms_host * 1e-3), "ms_per_step": ms_dev / steps,
                "scenes_per_step": per_step * arm.batch, "timed_s": ms_dev * 1e-3, "timed_s_e2e": ms_host * 1e-3}

    sampler = ClockSampler(local)
    sampler.start()  # started before the warm-up: its first fork/exec of nvidia-smi stays outside the timed regions
    arm = Arm(args.workload, rank, dev, args.in_flight, args.no_tc, args.operands)
    model = arm.model
    for i in range(2):
        arm.serial(i)
    sampler.rows.clear()
    main_m = measure(arm, args.steps, args.warmup, args.scenes_per_step)
    clocks = sampler.summary()
    ms_serial = timed_serial(arm, 12)

    # ---- roofline of the tensor-core convolution kernel: FLOPs of the two stacks (accounting pass) / graph replay time
    model.flop_acc = {}
    arm.serial(0)
    flop_acc, model.flop_acc = model.flop_acc, None
    torch.cuda.synchronize()
    stack_ms = model.time_conv_stacks(20)
    g_flops = sum(v["flops"] for k, v in flop_acc.items() if k in stack_ms)
    g_launch = sum(v["launches"] for k, v in flop_acc.items() if k in stack_ms)
    g_bytes = sum(v["bytes"] for k, v in flop_acc.items() if k in stack_ms)
    g_ms = sum(stack_ms.values())
    all_flops = sum(v["flops"] for v in flop_acc.values())

    # kernel launch census of one forward (our kernels only: everything that is not an ATen kernel / memcpy / memset)
    launches_per_fwd = 0
    conv_gpu_share = None  # share of the convolution kernels in the GPU time of one forward (CUPTI; compare with the ncu launch list)
    try:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            arm.serial(0)
            torch.cuda.synchronize()
        t_all = t_conv = 0.0
        for e in prof.key_averages():
            if e.device_type != torch.autograd.DeviceType.CUDA:
                continue
            t = float(getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0.0) or 0.0)
            if "Memcpy" not in e.key and "Memset" not in e.key:
                t_all += t
                if "k_conv" in e.key:
                    t_conv += t
            if "at::" not in e.key and "Memcpy" not in e.key and "Memset" not in e.key:
                launches_per_fwd += e.count
        if t_all > 0:
            conv_gpu_share = t_conv / t_all
    except Exception:
        pass
    sampler.stop_flag = True
    sampler.join(timeout=2)

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
    achieved = g_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_conv_ncu_summary.json")) as f:
            ncu = json.load(f)
        if ncu.get("workload") == args.workload:
            traffic, traffic_src = ncu["dram_bytes_per_launch"], ncu["source"]
    except Exception:
        pass

    # ---- the other BASELINE configurations, briefly (the default run only)
    others = {}
    if args.workload == "C2" and not args.no_extras and not args.no_tc:
        for name in ("C3", "C4", "C5"):
            try:
                a2 = Arm(name, rank, dev, args.in_flight, False, args.operands)
                m = measure(a2, 3, 3, 16 if name != "C5" else 8)
                s_ms = timed_serial(a2, 4)
                entry = {"workload": workload_desc(name), "value": m["value"], "e2e": m["e2e"], "unit": "scenes/s", "timed_s": m["timed_s"],
                         "scenes_timed": 3 * m["scenes_per_step"], "serial_ms_per_forward": s_ms, "batch": a2.batch}
                if name == "C3":
                    # Chamfer alone on the full cloud (est vs gt alignment: a small rigid motion apart), both directions: the exact
                    # grid search the product uses and the every-pair kernel (the reference kernel's formulation) it is equal to
                    from pcaccumulation_b200.chamfer_distance import chamfer_with_indices

                    pts = a2.dev_pts[0][:, :3].contiguous()
                    moved = (pts + torch.tensor([0.05, -0.02, 0.01], device=pts.device)).contiguous()
                    n = pts.shape[0]

                    def time_chamfer(brute, reps):
                        chamfer_with_indices(pts[None], moved[None], brute=brute)
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for _ in range(reps):
                            out = chamfer_with_indices(pts[None], moved[None], brute=brute)
                        e1.record()
                        e1.synchronize()
                        return e0.elapsed_time(e1) / reps, out

                    ms_grid, o_grid = time_chamfer(False, 10)
                    ms_brute, o_brute = time_chamfer(True, 1)
                    same = all(torch.equal(x, y) for x, y in zip(o_grid, o_brute))
                    entry["chamfer"] = {
                        "n": n, "m": n, "ms": ms_grid, "algorithmic_pair_evals_per_s": 2.0 * n * n / (ms_grid * 1e-3),
                        "every_pair_kernel_ms": ms_brute, "every_pair_kernel_pair_evals_per_s": 2.0 * n * n / (ms_brute * 1e-3),
                        "every_pair_kernel_fp32_flops_per_s": 8.0 * 2.0 * n * n / (ms_brute * 1e-3), "bit_identical": bool(same),
                        "note": "grid search = exact nearest neighbour over a uniform grid (csrc/nn_grid.cu); every-pair kernel: 8 FP32 ops "
                                "per pair against the FP32 FMA-pipe peak measured at 62-72 TFLOP/s (tools/ffma_bench.cu)"}
                    if not same:  # (the GPU tests assert this equality; here it is recorded next to the timings)
                        print("bench.py: grid Chamfer differs from the every-pair kernel", file=sys.stderr)
                others[name] = entry
                a2.close()
                del a2
                torch.cuda.empty_cache()
            except Exception as e:  # an extra must never take the headline measurement down with it
                others[name] = {"error": repr(e)[:300]}

    # ---- CPU arm + parity (rank 0)
    cpu_base, parity = None, None
    if rank == 0 and not args.no_cpu_baseline:
        torch.set_num_threads(cpu_threads())
        run = oracle_forward_fn(arm.cfg, arm.sd)
        pick = lambda i: [arm.scenes[(i + j) % N_SCENES] for j in range(arm.batch)]
        run(pick(0))  # warm-up (numba jit, oneDNN primitives)
        n_cpu = 3
        t0 = time.perf_counter()
        for i in range(n_cpu):
            res, _ = run(pick(i))
            if args.workload == "C3":
                cpu_alignment_errors(pick(i)[0], res)
        dt = time.perf_counter() - t0
        cpu_base = {"value": n_cpu * arm.batch / dt, "unit": "scenes/s", "cores": torch.get_num_threads(), "kind": "port",
                    "sample": f"{n_cpu} forwards of {arm.batch} scene(s) of {args.workload} (numba voxelise + collate + forward) after 1 "
                              "warm-up, oracle/oracle.py (port of the reference's CPU path)"}
    if rank == 0 and not args.no_parity:
        # the staged protocol on one scene of the workload, on the path that was timed; raises on anything that is not a
        # float32 rounding tie (the run then fails) -- see oracle/protocol.py for the rules
        from oracle import oracle
        from oracle.protocol import REPORT, run_protocol

        torch.set_num_threads(cpu_threads())
        scene = arm.scenes[0]
        vg = arm.cfg["voxel_generator"]
        p4 = np.concatenate((scene["input_points"], scene["time_indice"]), 1).astype(np.float32)
        sample = dict(scene)
        sample.update(oracle.voxelize(p4, vg["voxel_size"], vg["range"], vg["n_sweeps"]))
        model.keep_stages = True
        parity_failure = None
        try:
            run_protocol(model, arm.cfg, arm.sd, synth.collate([sample]), 42, "bench")
        except AssertionError as e:  # the line is still printed (with the failure in it), then the run exits non-zero
            parity_failure = str(e)[:500]
        model.keep_stages = False
        rec = REPORT["bench"]
        floats = {k: v for k, v in rec.items() if isinstance(v, dict) and "err_vs_ref32" in v}
        worst = max(floats.items(), key=lambda kv: kv[1]["err_vs_ref32"]) if floats else ("-", {"err_vs_ref32": None, "fp32_floor": None})
        parity = {"scene_points": int(p4.shape[0]), "protocol": "oracle/protocol.py (free-running + staged, float64 floor)",
                  "failed": parity_failure,
                  "fb_label_mismatches": rec.get("fb_points", {}).get("flips"), "fb_cell_flips": rec.get("fb_cells", {}).get("flips"),
                  "fb_flips_unexplained": 0 if parity_failure is None else None,
                  "mos_label_mismatches_staged": rec.get("mos_points", {}).get("flips"),
                  "inst_label_mismatches_staged": 0 if parity_failure is None else None, "free_running": rec.get("free_running"),
                  "largest_float_error_vs_reference": {"tensor": worst[0], "rel": worst[1]["err_vs_ref32"], "fp32_floor": worst[1]["fp32_floor"]},
                  "tensors_compared": len(floats), "tensors_beyond_1e-4": sum(1 for v in floats.values() if v["err_vs_ref32"] > 1e-4),
                  "reference_own_flips_vs_float64": {"fb_cells": rec.get("ref32_vs_ref64_fb_cells"), "mos_points": rec.get("ref32_vs_ref64_mos_points")}}

    if rank == 0:
        fmt = "tc-p16" if (not args.no_tc and model.conv_operands == "f16" and model.packed_activations) else ("tc" if not args.no_tc else "f32")
        per_fwd = main_m["scenes_per_step"] // arm.batch
        line = {
            "metric": "scenes/sec", "value": main_m["value"], "unit": "scenes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_m["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": ("f32 (tensor cores: fp16-pair split operands and activations, 22 significant bits, FP32 accumulate)" if fmt == "tc-p16"
                      else "f32 (tensor cores: 3xTF32 split, FP32 accumulate)" if fmt == "tc" else "f32"),
            "data": "synthetic",
            "config": {"workload": workload_desc(args.workload), "batch": arm.batch, "mode": "test",
                       "step": f"one batch of {main_m['scenes_per_step']} independent scenes", "scenes_per_step": main_m["scenes_per_step"],
                       "distinct_scenes_per_rank": N_SCENES, "scenes_in_flight": max(1, args.in_flight), "host_sync": args.host_sync, "host_cores_per_rank": args.cores_per_rank,
                       "timed_region_s": main_m["timed_s"], "timed_region_s_e2e": main_m["timed_s_e2e"],
                       "serial_ms_per_forward": ms_serial, "parallelism": f"dp{world} (scene sharding, no data-path collective)",
                       "l2": "per-scene working set (>1 GB of activations) exceeds the 126 MB L2; no explicit flush",
                       "conv_path": fmt},
            "e2e": {"value": main_m["e2e"], "unit": "scenes/s", "h2d_bytes_per_step": arm.h2d * per_fwd, "d2h_bytes_per_step": arm.d2h * per_fwd},
            "gpu_launches": launches_per_fwd * args.steps * per_fwd,
            "gpu_launches_per_forward": launches_per_fwd,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "k_conv_p16 (tcgen05 conv3x3 + ConvTranspose2x2; %d launches in the two captured stacks)" % g_launch,
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_flops_per_launch": g_flops / max(g_launch, 1), "algorithmic_bytes_per_launch": g_bytes / max(g_launch, 1),
                         "algorithmic_flops_per_forward_in_stacks": g_flops, "algorithmic_flops_per_forward_all_convs": all_flops,
                         "stack_replay_ms": stack_ms, "conv_share_of_serial_forward": g_ms / ms_serial,
                         "conv_share_of_gpu_time": conv_gpu_share,
                         "how": "CUDA events around 20 back-to-back replays of each captured stack (backbone UNet + merged head conv; "
                                "Conv3d x4 + STPN UNet); the replays include the 9 pooling launches (~2% of the time)",
                         "peak_source": peak_src},
            "cpu_baseline": cpu_base,
            "parity": parity,
            "other_configs": others,
        }
        print(json.dumps(line), flush=True)
    arm.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and parity is not None and parity.get("failed"):
        print("bench.py: PARITY FAILED: " + parity["failed"], file=sys.stderr)
        sys.exit(1)


if __name__ == "__main__":
    main()
