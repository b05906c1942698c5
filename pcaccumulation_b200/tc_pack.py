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
