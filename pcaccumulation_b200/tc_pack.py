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


def f16_weight_scale(w):
    """Power-of-two scale for the fp16-pair split of a weight tensor: 256 unless that would push max|w| beyond 2^14 (fp16
    overflows at 65504; a checkpoint with |w| >= 256 must not turn into inf), then the largest power of two that fits."""
    m = float(w.detach().abs().max()) if w.numel() else 0.0
    s = F16_WEIGHT_SCALE
    while m * s > 16384.0 and s > 2.0 ** -24:
        s *= 0.5
    return s


def pack_rows_f16(rows, scale):
    """rows [R][K] float32 (K-major, K % 32 == 0) -> fp16 [2 (h, l)][R][K/32][64]: the first 32 of every 64 = one 32-channel
    group, the rest zero (so a group is one 128-byte swizzle row of the MMA B operand).  h = fp16(s*w), l = fp16(s*w - h)."""
    r, k = rows.shape
    w = rows.float() * scale
    h = w.half()
    l = (w - h.float()).half()
    out = torch.zeros(2, r, k // 32, 64, dtype=torch.float16, device=rows.device)
    out[0, :, :, :32] = h.view(r, k // 32, 32)
    out[1, :, :, :32] = l.view(r, k // 32, 32)
    return out.contiguous()


def pack_conv_tc_f16(layer, scale=F16_WEIGHT_SCALE):
    """fp16-pair pack of a 3x3 convolution for ``pcab_conv3x3_tc_f16`` / ``pcab_conv3x3_p16``: rows = output channels, K = the
    tf32 pack's order (per source: tap-major, channel-minor)."""
    k = layer.pack.numel() // layer.cout
    return pack_rows_f16(layer.pack.view(k, layer.cout).t().contiguous(), scale)


def pack_rows_f16_dense(rows, scale):
    """rows [R][K] float32 (K % 32 == 0, already in the kernel's consumption order) -> fp16 [2 (h, l)][R][Kpad], Kpad = K rounded
    up to a multiple of 64 (two 32-channel K groups per 128-byte row of the MMA B operand).  h = fp16(s*w), l = fp16(s*w - h)."""
    r, k = rows.shape
    kp = (k + 63) // 64 * 64
    w = rows.float() * scale
    h = w.half()
    l = (w - h.float()).half()
    out = torch.zeros(2, r, kp, dtype=torch.float16, device=rows.device)
    out[0, :, :k] = h
    out[1, :, :k] = l
    return out.contiguous()


def pack_conv_p16(layer, scale):
    """Weight pack of ``pcab_conv3x3_p16``: rows = output channels; K in consumption order: per source, per 32-channel chunk,
    per tap (ky*3+kx), 32 channels.  ``layer.pack`` holds per source [9][C_s][Cout]."""
    cols, off = [], 0
    for cs in layer.splits:
        blk = layer.pack[off:off + 9 * cs * layer.cout].view(9, cs // 32, 32, layer.cout)
        cols.append(blk.permute(1, 0, 2, 3).reshape(9 * cs, layer.cout))  # [chunk][tap][32] x Cout
        off += 9 * cs * layer.cout
    return pack_rows_f16_dense(torch.cat(cols, 0).t().contiguous(), scale)


def pack_conv3d_fused_p16(weight, scale):
    """Conv3d weight [32, 32, 3 (kt), 3, 3] for ``pcab_conv3d_p16`` (temporal taps fused into the MMA N dimension): rows =
    (kt = 2, 1, 0) x output channel -- input frame f feeds output frames f-1, f, f+1 through kt = 2, 1, 0 --, K = tap (ky*3+kx)
    x input channel, dense, padded to 320."""
    cout, cin = weight.shape[:2]
    assert cout == 32 and cin == 32
    rows = torch.cat([weight.detach().float()[:, :, kt].permute(0, 2, 3, 1).reshape(cout, 9 * cin) for kt in (2, 1, 0)], 0)
    return pack_rows_f16_dense(rows.contiguous(), scale)


def pack_convT_p16(weight, scale):
    """ConvTranspose2d(2, stride 2) weight [Cin, Cout, 2, 2] for ``pcab_convT2x2_p16``: GEMM rows (output columns) =
    (dy*2+dx)*Cout + co, K = Cin."""
    cin, cout = weight.shape[:2]
    rows = weight.detach().float().permute(2, 3, 1, 0).reshape(4 * cout, cin).contiguous()
    return pack_rows_f16_dense(rows, scale)


def unpack_p16(t):
    """Decode a P16 activation tensor (float32-typed storage [..., C], csrc/pair16.cuh) to float32 values h + l."""
    c = t.shape[-1]
    h = t.contiguous().view(torch.float16).view(*t.shape[:-1], c // 32, 2, 32)
    return (h[..., 0, :].float() + h[..., 1, :].float()).reshape(*t.shape[:-1], c)


def pack_p16(x):
    """Encode float32 values [..., C] (C % 32 == 0) as a P16 tensor (float32-typed storage of the same shape)."""
    c = x.shape[-1]
    x = x.float().contiguous().view(*x.shape[:-1], c // 32, 32)
    h = x.clamp(-65504.0, 65504.0).half()
    l = (x - h.float()).clamp(-65504.0, 65504.0).half()
    return torch.stack((h, l), -2).contiguous().view(*x.shape[:-2], 2 * c).view(torch.float32)


def pack_pfn_tc(pe):
    """Tensor-core packs of the pillar encoder (csrc/pillar.cu:pfn_tc) from the reference's ``PillarFeatureNet`` parameters
    (models/pillar_encoder.py:59-95).  Returns (fp16 blob [3 stages][20480], float32 bias [3][128], 9 host floats 1/scale)."""
    def hl(w, scale, rows, width=64):
        """[out, in] float32 -> [h rows (rows); l rows (rows)] x width fp16 (zero padded)."""
        out = torch.zeros(2, rows, width, dtype=torch.float16, device=w.device)
        x = w.detach().float() * scale
        h = x.half()
        out[0, : w.shape[0], : w.shape[1]] = h
        out[1, : w.shape[0], : w.shape[1]] = (x - h.float()).half()
        return out.reshape(-1)

    dev = pe.fc_pos.weight.device
    blobs, biases, inv = [], [], []
    for s, blk in enumerate(pe.blocks):
        s0 = f16_weight_scale(blk.fc_0.weight)
        s1 = min(f16_weight_scale(blk.shortcut.weight), f16_weight_scale(blk.fc_1.weight))
        w0 = hl(blk.fc_0.weight, s0, 32)                                   # K groups 0, 1 of the block input
        w1a = hl(blk.shortcut.weight, s1, 32)                              # stage 0 of [Ws | W1]: K groups 0, 1
        w1b = hl(blk.fc_1.weight, s1, 32)                                  # stage 1: K group 2 (+ 32 zeros)
        bx = torch.zeros(64, device=dev)
        if s == 0:
            sx = f16_weight_scale(pe.fc_pos.weight)
            wx = hl(pe.fc_pos.weight, sx, 64)
            bx = pe.fc_pos.bias.detach().float()
        elif s == 2:
            sx = f16_weight_scale(pe.fc_c.weight)
            wx = torch.cat((hl(pe.fc_c.weight, sx, 32), torch.zeros(64 * 64, dtype=torch.float16, device=dev)))
            bx[:32] = pe.fc_c.bias.detach().float()
        else:
            sx, wx = 1.0, torch.zeros(128 * 64, dtype=torch.float16, device=dev)
        blobs.append(torch.cat((w0, w1a, w1b, wx)))
        biases.append(torch.cat((blk.fc_0.bias.detach().float(), blk.fc_1.bias.detach().float(), bx)))
        inv += [1.0 / s0, 1.0 / s1, 1.0 / sx]
    blob = torch.stack(blobs).contiguous()
    assert blob.shape == (3, 20480)
    return blob, torch.stack(biases).contiguous(), inv
