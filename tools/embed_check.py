"""Developer script: phase counters of the tensor-core TubeNet embed kernel on synthetic rows."""
import sys
import torch
sys.path.insert(0, ".")
from pcaccumulation_b200 import _lib as L, config, fixture
from pcaccumulation_b200._lib import I, P, call, stream
from pcaccumulation_b200.motionnet import MotionNet
cfg = config.workload_config("C2")
model = MotionNet(cfg).cuda().eval()
model.load_state_dict(fixture.fixture_state_dict(model.state_dict(), 42))
W = model._weights()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96000
K = 100
feat = torch.randn(n, 64, device="cuda")
seg = torch.sort(torch.randint(0, K, (n,), device="cuda")).values.to(torch.int32)
src = torch.randperm(n, device="cuda").to(torch.int32)
out = torch.empty(K, 128, device="cuda")
ws, bias = W["tpn_motion_tc"]
def run():
    call("pcab_embed_segmax_tc", I(0), P(feat), P(src), P(seg), I(n), I(K), P(ws[0]), P(ws[1]), P(ws[2]), P(bias), P(out), stream())
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
print("motion embed tc %.3f ms" % (e0.elapsed_time(e1) / 5))
st = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
L.lib().pcab_stpn_head_tc_set_stats(P(st))
run(); torch.cuda.synchronize()
L.lib().pcab_stpn_head_tc_set_stats(P(None))
tiles = (n + 127) // 128
per = st.view(-1, 8)[:148].double().sum(0) / tiles
print("cycles per tile:", " ".join("%s %.0f" % (nm, v) for nm, v in zip(["load+put", "wait L0", "epi L0", "wait L1", "epi L1", "wait L2", "epi L2(segmax)", "-"], per.tolist())), "| total %.0f" % per.sum().item())
