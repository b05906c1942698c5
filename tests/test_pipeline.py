"""ScenePipeline (several scenes in flight on one GPU) must return exactly what the serial runner returns."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pipeline_matches_serial_runner(fixture_weights):
    from pcaccumulation_b200 import config, synth
    from pcaccumulation_b200.runner import ScenePipeline, SceneRunner, scene_to_points4

    cfg = config.workload_config("C1")
    sd = fixture_weights(cfg)
    serial = SceneRunner(cfg)
    serial.model.load_state_dict(sd)
    scenes = [synth.make_workload_scene("C1", i) for i in range(3)]
    pts = [torch.tensor(scene_to_points4(s)).cuda() for s in scenes]
    ego = [torch.tensor(s["ego_motion_gt"])[None].cuda() for s in scenes]
    ref = []
    for i, p in enumerate(pts):
        torch.manual_seed(100 + i)  # the serial runner draws keypoints from the global CPU generator, as upstream
        ref.append(serial.run_device(p, [p.shape[0]], ego_motion_gt=ego[i]))
    torch.cuda.synchronize()
    pipe = ScenePipeline(cfg, sd, depth=3)
    try:
        # twice as many jobs as slots, submitted at once: scenes overlap on different streams
        futs = [pipe.submit(pts[i % 3], [pts[i % 3].shape[0]], ego=ego[i % 3], seed=100 + i % 3) for i in range(6)]
        outs = []
        for f in futs:
            res, done = f.result()
            done.synchronize()
            outs.append(res)
    finally:
        pipe.close()
    for i, res in enumerate(outs):
        r = ref[i % 3]
        for key in ("fb_est_per_points", "inst_labels_est"):
            assert torch.equal(res[key], r[key]), key
        for key in ("ego_motion_est", "mos_est", "offset_est", "rec_est", "fb_seg_est"):
            assert torch.allclose(res[key], r[key], rtol=0, atol=1e-6 * max(1.0, float(r[key].abs().max()))), key


def test_pipeline_host_buffers(fixture_weights):
    from pcaccumulation_b200 import config, synth
    from pcaccumulation_b200.runner import ScenePipeline, scene_to_points4

    cfg = config.workload_config("C1")
    pipe = ScenePipeline(cfg, fixture_weights(cfg), depth=2)
    try:
        s = synth.make_workload_scene("C1", 0)
        p = torch.tensor(scene_to_points4(s)).pin_memory()
        n = p.shape[0]
        outs = [{"rec_est": torch.empty(n, 3).pin_memory(), "fb": torch.empty(n, dtype=torch.int64).pin_memory()} for _ in range(2)]
        futs = [pipe.submit(p, [n], seed=5, out=outs[i], host=True) for i in range(2)]
        res = []
        for f in futs:
            r, done = f.result()
            done.synchronize()
            res.append(r)
    finally:
        pipe.close()
    for o, r in zip(outs, res):
        assert torch.equal(o["fb"], r["fb_est_per_points"][:, 0].cpu())
        assert torch.equal(o["rec_est"], r["rec_est"].cpu())
    assert torch.equal(outs[0]["fb"], outs[1]["fb"])
