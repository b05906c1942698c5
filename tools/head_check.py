"""Developer script: the tensor-core STPN head (pcab_stpn_head_tc) against the FP32 CUDA-core head on random inputs + timing."""
import sys
import torch
sys.path.insert(0, ".")
from pcaccumulation_b200 import _lib as L, config, fixture
from pcaccumulation_b200._lib import F, I, P, call, stream
from pcaccumulation_b200.motionnet import MotionNet

torch.manual_seed(0)
cfg = config.workload_config("C2")
model = MotionNet(cfg).cuda().eval()
model.load_state_dict(fixture.fixture_state_dict(model.state_dict(), 42))
W = model._weights()
H = Wd = 288
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
feats = torch.randn(1, H, Wd, 64, device="cuda")
tp = (torch.rand(N, 3, device="cuda") * 2 - 1) * torch.tensor([36.0, 36.0, 3.0], device="cuda")
pbatch = torch.zeros(N, dtype=torch.int32, device="cuda")
n_fg = int(sys.argv[2]) if len(sys.argv) > 2 else 150001
fg_idx = torch.randperm(N, device="cuda")[:n_fg].sort().values.to(torch.int32)
outs = []
for name in ("pcab_stpn_head", "pcab_stpn_head_tc"):
    mos = torch.zeros(N, 2, device="cuda"); off = torch.zeros(N, 2, device="cuda")
    def run():
        if name.endswith("_tc"):
            call(name, P(feats), I(H), I(Wd), P(tp), P(pbatch), P(fg_idx), I(n_fg), P(W["stpn_head_host"]), P(W["stpn_head_tc1"]),
                 P(W["stpn_head_tc"]), F(36.0), F(36.0), P(mos), P(off), stream())
        else:
            call(name, P(feats), I(H), I(Wd), P(tp), P(pbatch), P(fg_idx), I(n_fg), P(W["stpn_head"]), F(36.0), F(36.0), P(mos), P(off), stream())
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    print(name, "%.3f ms" % (e0.elapsed_time(e1) / 5))
    outs.append((mos.clone(), off.clone()))
(m0, o0), (m1, o1) = outs
sel = fg_idx.long()
for nm, a, b in (("mos", m0[sel], m1[sel]), ("off", o0[sel], o1[sel])):
    err = (a - b).abs().max().item(); scale = a.abs().max().item()
    print(nm, "max abs err %.3e  scale %.3e  rel %.2e  mean abs err %.2e" % (err, scale, err / scale, (a - b).abs().mean().item()))
print("argmax flips", int((m0[sel].argmax(1) != m1[sel].argmax(1)).sum()), "of", n_fg)
untouched = torch.ones(N, dtype=torch.bool, device="cuda"); untouched[sel] = False
print("untouched rows stay zero:", bool((m1[untouched] == 0).all()) and bool((o1[untouched] == 0).all()))

# phase breakdown of the tensor-core head (cycles per tile, averaged over CTAs)
st = torch.zeros(148 * 2 * 8, dtype=torch.int64, device="cuda")
L.lib().pcab_stpn_head_tc_set_stats(P(st))
mos = torch.zeros(N, 2, device="cuda"); off = torch.zeros(N, 2, device="cuda")
call("pcab_stpn_head_tc", P(feats), I(H), I(Wd), P(tp), P(pbatch), P(fg_idx), I(n_fg), P(W["stpn_head_host"]), P(W["stpn_head_tc1"]),
     P(W["stpn_head_tc"]), F(36.0), F(36.0), P(mos), P(off), stream())
torch.cuda.synchronize()
L.lib().pcab_stpn_head_tc_set_stats(P(None))
s = st.view(148, 2, 8).double()
tiles = (n_fg + 127) // 128
per = s.sum(0) / tiles
names = ["gather+pe0", "wait pe2", "epi pe2", "wait fp", "epi fp", "wait head", "epi head", "wait last"]
for hh in range(2):
    print("half", hh, " ".join("%s %.0f" % (n, v) for n, v in zip(names, per[hh].tolist())), "| total %.0f" % per[hh].sum().item())
