"""Developer script: tcgen05 conv vs FP32 CUDA-core conv on the layer shapes of the model (accuracy + time)."""
import sys
import time

import torch

sys.path.insert(0, ".")
from pcaccumulation_b200 import _lib, motionnet as mn, tc_pack  # noqa: E402
from pcaccumulation_b200._lib import I, P, call, stream  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
_lib.lib().pcab_conv3x3_tc_set_base_offset_mode(I(mode))
torch.manual_seed(0)
dev = "cuda"
# (n_img, [src channels], Cout, H, W, temporal_T)
shapes = [(5, [32], 32, 288, 288, 1), (5, [32], 64, 288, 288, 1), (5, [64], 64, 144, 144, 1), (5, [32, 32], 32, 288, 288, 1),
          (5, [64, 64], 64, 144, 144, 1), (5, [128], 128, 72, 72, 1), (5, [128, 128], 128, 72, 72, 1), (5, [64], 128, 72, 72, 1),
          (5, [32], 32, 288, 288, 5), (1, [64], 64, 288, 288, 1), (2, [32], 32, 100, 76, 1), (1, [256], 256, 72, 72, 1),
          (5, [256], 512, 18, 18, 1), (5, [512], 512, 18, 18, 1), (1, [128], 256, 18, 18, 1), (1, [256], 256, 18, 18, 1), (5, [256, 256], 256, 36, 36, 1)]
for n, cs, cout, H, W, T in shapes:
    cin = sum(cs)
    xs = [torch.randn(n, H, W, c, device=dev) for c in cs]
    if T > 1:
        w = torch.randn(cout, cs[0], 3, 3, 3, device=dev) * 0.1

        class Lyr:
            pass
        conv = type("C", (), {"weight": w, "bias": torch.randn(cout, device=dev)})()
        layer = mn._ConvLayer(conv, temporal=True)
        srcs = [xs[0]] * 3
        c3 = [cs[0]] * 3
    else:
        w = torch.randn(cout, cin, 3, 3, device=dev) * 0.1
        conv = type("C", (), {"weight": w, "bias": torch.randn(cout, device=dev)})()
        layer = mn._ConvLayer(conv, splits=cs)
        srcs = xs + [None] * (3 - len(xs))
        c3 = cs + [0] * (3 - len(cs))
    ok = _lib.lib().pcab_conv3x3_tc_supported(I(3 if T > 1 else len(cs)), I(c3[0]), I(c3[1]), I(c3[2]), I(cout), I(H), I(W))
    ref = torch.empty(n, H, W, cout, device=dev)
    out = torch.full((n, H, W, cout), float("nan"), device=dev)
    sc, sh = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev)
    args = lambda pack, o: (P(srcs[0]), I(c3[0]), P(srcs[1]), I(c3[1]), P(srcs[2]), I(c3[2]), I(T), P(pack), P(layer.bias), P(sc), P(sh),
                            I(1), P(o), I(n), I(H), I(W), I(cout), I(cout), I(0), stream())
    call("pcab_conv3x3_f32", *args(layer.pack, ref))
    torch.cuda.synchronize()
    if not ok:
        print(f"shape n={n} cs={cs} cout={cout} {H}x{W} T={T}: TC unsupported")
        continue
    tcw = tc_pack.pack_conv_tc(layer)
    try:
        call("pcab_conv3x3_tc", *args(tcw, out))
        torch.cuda.synchronize()
    except Exception as e:
        print("TC FAILED", n, cs, cout, H, W, T, e)
        sys.exit(1)
    err = (out - ref).abs().max().item()
    if T == 1 and n * H * W <= 5 * 144 * 144:
        import torch.nn.functional as F
        x64 = torch.cat(xs, 3).permute(0, 3, 1, 2).double()
        y64 = F.relu(F.conv2d(x64, w.double(), layer.bias.double(), padding=1) * sc.double()[None, :, None, None] + sh.double()[None, :, None, None])
        y64 = y64.permute(0, 2, 3, 1)
        print("   vs float64: f32 path err %.3e | tc path err %.3e" % ((ref.double() - y64).abs().max().item(), (out.double() - y64).abs().max().item()))
    # fp16-pair operand variant
    from pcaccumulation_b200._lib import F
    out16 = torch.full((n, H, W, cout), float("nan"), device=dev)
    w16 = tc_pack.pack_conv_tc_f16(layer)
    args16 = lambda o: (P(srcs[0]), I(c3[0]), P(srcs[1]), I(c3[1]), P(srcs[2]), I(c3[2]), I(T), P(w16), F(1.0 / tc_pack.F16_WEIGHT_SCALE),
                        P(layer.bias), P(sc), P(sh), I(1), P(o), I(n), I(H), I(W), I(cout), I(cout), I(0), stream())
    call("pcab_conv3x3_tc_f16", *args16(out16))
    torch.cuda.synchronize()
    err16 = (out16 - ref).abs().max().item()
    for _ in range(2):
        call("pcab_conv3x3_tc_f16", *args16(out16))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call("pcab_conv3x3_tc_f16", *args16(out16))
    e1.record()
    torch.cuda.synchronize()
    t16 = e0.elapsed_time(e1) / 10
    nan = torch.isnan(out).sum().item() + torch.isnan(out16).sum().item()
    flops = 2.0 * 9 * cin * cout * H * W * n * (1 if T == 1 else (3 * T - 2) / T)
    ts = {}
    for name, fn, pack, o in (("f32", "pcab_conv3x3_f32", layer.pack, ref), ("tc", "pcab_conv3x3_tc", tcw, out)):
        for _ in range(2):
            call(fn, *args(pack, o))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call(fn, *args(pack, o))
        e1.record()
        torch.cuda.synchronize()
        ts[name] = e0.elapsed_time(e1) / 10
    import ctypes
    plan = (ctypes.c_int * 10)()
    _lib.lib().pcab_conv3x3_tc_plan(I(n), I(H), I(W), I(cout), plan)
    print("   plan c=%d mt=%d strip=%d mtx=%d R=%d Wt=%d tiles=%dx%d ctiles=%d items=%d (%.2f rounds)" % (*list(plan), plan[9] / 148))
    if "--stats" in sys.argv:
        st = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
        _lib.lib().pcab_conv3x3_tc_set_stats(P(st))
        if "--f16" in sys.argv:
            call("pcab_conv3x3_tc_f16", *args16(out16))
        else:
            call("pcab_conv3x3_tc", *args(tcw, out))
        torch.cuda.synchronize()
        _lib.lib().pcab_conv3x3_tc_set_stats(P(None))
        s2 = st.view(148, 16).double()
        act = s2[:, 8] > 0
        m = s2[act].mean(0)
        names = ["mma:wait_lo", "mma:wait_acc", "mma:wait_w", "split:wait_free", "split:wait_hi", "split:work", "drain:wait_acc", "mma:total", "chunks", "epilogue", "mma:issue"]
        print("   stats (mean cycles per CTA, %d CTAs): " % int(act.sum()) + ", ".join(f"{nm}={m[i].item():.0f}" for i, nm in enumerate(names)),
              "| per chunk: total %.0f" % (m[7] / m[8]).item())
    print(f"n={n} cs={cs} cout={cout} {H}x{W} T={T}: maxerr {err:.3e} (ref max {ref.abs().max().item():.1f}) nan {nan} | "
          f"f32 {ts['f32']:.3f} ms {flops / ts['f32'] / 1e9:.1f} TF/s | tc {ts['tc']:.3f} ms {flops / ts['tc'] / 1e9:.1f} TF/s | "
          f"f16-pair maxerr {err16:.3e} {t16:.3f} ms {flops / t16 / 1e9:.1f} TF/s")
