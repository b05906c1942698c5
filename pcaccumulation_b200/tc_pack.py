"""Weight pack of the tcgen05 convolution path: [2 (hi, lo)][Cout][K] with K = the row order of the FP32 pack
(per source: tap-major, channel-minor).  hi = w rounded to the nearest tf32 (sign, exponent, 10 mantissa bits; zero-mean split error),
lo = w - hi (exact in FP32); see csrc/conv_tc.cu."""
import torch


def pack_conv_tc(layer):
    k = layer.pack.numel() // layer.cout
    w = layer.pack.view(k, layer.cout).t().contiguous()  # [Cout][K]
    hi = ((w.view(torch.int32) + 0x1000) & -8192).view(torch.float32)  # round the magnitude to the nearest tf32
    lo = w - hi
    return torch.cat((hi, lo), 0).contiguous()


def split_tf32(w):
    """w [rows, K] (K-major) -> [hi rows; lo rows]: hi = w rounded to the nearest tf32, lo = w - hi (exact in FP32)."""
    w = w.detach().float().contiguous()
    hi = ((w.view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    return torch.cat((hi, w - hi), 0).contiguous()


def pack_stpn_head_tc(mh):
    """Tensor-core packs of the STPN point head (csrc/mlp_tc.cu): nn.Linear weights are [out, in] = K-major rows already.
    Returns (pe2 [128, 32], [final_proj; mos0; off0] [768, 128])."""
    w1 = split_tf32(mh.positional_encoding[2].weight)
    w = torch.cat([split_tf32(mh.final_proj[0].weight), split_tf32(mh.mos_seg.seg_head[0].weight),
                   split_tf32(mh.offset_head.seg_head[0].weight)], 0).contiguous()
    return w1, w


def pack_embed_tc(seq, first=0):
    """Tensor-core packs of a TubeNet embedding MLP (nn.Sequential of Linear / ReLU): ([split weight per Linear from
    ``first``], host tensor of their biases back to back)."""
    lin = [m for m in seq if hasattr(m, "weight")][first:]
    return [split_tf32(m.weight) for m in lin], torch.cat([m.bias.detach().float().reshape(-1) for m in lin]).cpu().contiguous()


F16_WEIGHT_SCALE = 256.0  # power of two: keeps the l halves of O(0.01) weights out of the fp16 subnormals


def pack_conv_tc_f16(layer):
    """fp16-pair pack of a 3x3 convolution for ``pcab_conv3x3_tc_f16``: [2 (h, l)][Cout][K/32][64] fp16, the first 32 of every 64
    = one 32-channel group of the tf32 pack's K order, the rest zero.  h = fp16(s*w), l = fp16(s*w - h)."""
    k = layer.pack.numel() // layer.cout
    w = layer.pack.view(k, layer.cout).t().contiguous() * F16_WEIGHT_SCALE  # [Cout][K]
    h = w.half()
    l = (w - h.float()).half()
    out = torch.zeros(2, layer.cout, k // 32, 64, dtype=torch.float16, device=w.device)
    out[0, :, :, :32] = h.view(layer.cout, k // 32, 32)
    out[1, :, :, :32] = l.view(layer.cout, k // 32, 32)
    return out.contiguous()
