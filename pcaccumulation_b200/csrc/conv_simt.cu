// FP32 CUDA-core convolutions on NHWC activations: exact-order-free but full-precision path.
//
// conv3x3 (pad 1) with up to three input sources accumulated into one output.  The multi-source form
// covers, without materialising anything:
//   * plain Conv2d                                  (models/unet.py:11-20)           1 source
//   * Conv2d over torch.cat((up, skip), 1)          (models/unet.py:106-111)         2 sources
//   * Conv3d 3x3x3 over [B,C,T,H,W]                 (models/stpn.py:13-22)           3 sources = frames t-1,t,t+1
// plus ConvTranspose2d(k=2,s=2) (models/unet.py:22-28), MaxPool2d(2) and the temporal max of
// models/stpn.py:80.  Epilogue: bias, optional BatchNorm(eval) scale/shift (models/unet.py:264-269), ReLU.
//
// These kernels are the full-precision reference path inside the product (used for layers whose shape
// does not suit the tcgen05 tiles and as the on-device cross-check of the tensor-core path).
#include "common.cuh"
#include "pair16.cuh"
#include "pcab200.h"

namespace {

constexpr int CK = 16;  // input channels staged per chunk

struct ConvArgs {
  const float* src[3];
  int src_c[3];      // channels of each source
  int frame_shift[3];  // conv3d: source image index = n + shift, valid iff 0 <= (n % T) + shift < T
  int nsrc;
  int T;             // frames per scene for the temporal validity test (1 = plain 2-D)
  const float* w;    // per source block [9][C_s][Cout], blocks concatenated
  const float* bias;
  const float* bn_scale;  // optional: y = (acc + bias) * scale + shift
  const float* bn_shift;
  float* out;
  int N, H, W, Cout;
  int out_cstride, out_coff;
  int relu;
};

template <int TH, int TW, int COT>
__global__ void __launch_bounds__((TH * TW / 4) * (COT / 16)) k_conv3x3(ConvArgs a) {
  constexpr int NPG = TH * TW / 4;   // pixel groups (4 consecutive x)
  constexpr int NT = NPG * (COT / 16);
  constexpr int PH = TH + 2, PW = TW + 4;  // patch pitch padded to a multiple of 4
  extern __shared__ float smem[];
  float* s_in = smem;                      // [CK][PH][PW]
  float* s_w = smem + CK * PH * PW;        // [9][CK][COT]

  const int tiles_x = (a.W + TW - 1) / TW;
  const int tile = blockIdx.x;
  const int ty0 = (tile / tiles_x) * TH, tx0 = (tile % tiles_x) * TW;
  const int n = blockIdx.z;
  const int co0 = blockIdx.y * COT;
  const int tid = threadIdx.x;
  const int pg = tid % NPG, cg = tid / NPG;
  const int py = pg / (TW / 4), px = (pg % (TW / 4)) * 4;

  float acc[4][16];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[p][o] = 0.f;

  int w_row0 = 0;  // running row offset (in units of [C_s] rows) into the packed weights
  size_t w_base = 0;
  for (int s = 0; s < a.nsrc; ++s) {
    const int Cs = a.src_c[s];
    const int tt = (a.T > 1) ? (n % a.T) + a.frame_shift[s] : 0;
    const bool valid = (a.T <= 1) || (tt >= 0 && tt < a.T);
    if (valid) {
      const float* img = a.src[s] + (size_t)(n + a.frame_shift[s]) * a.H * a.W * Cs;
      for (int c0 = 0; c0 < Cs; c0 += CK) {
        __syncthreads();
        // stage the input patch: (TH+2) x (TW+2) pixels x CK channels, channel-planar in smem
        for (int e = tid; e < PH * (TW + 2) * (CK / 4); e += NT) {
          int q = e % (CK / 4);
          int pix = e / (CK / 4);
          int r = pix / (TW + 2), cx = pix % (TW + 2);
          int gy = ty0 + r - 1, gx = tx0 + cx - 1;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W && c0 + 4 * q < Cs)
            v = *reinterpret_cast<const float4*>(img + ((size_t)gy * a.W + gx) * Cs + c0 + 4 * q);
          float* d = s_in + (4 * q) * PH * PW + r * PW + cx;
          d[0] = v.x, d[PH * PW] = v.y, d[2 * PH * PW] = v.z, d[3 * PH * PW] = v.w;
        }
        // stage weights [9][CK][COT]
        for (int e = tid; e < 9 * CK * (COT / 4); e += NT) {
          int o4 = e % (COT / 4);
          int k = (e / (COT / 4)) % CK;
          int tap = e / ((COT / 4) * CK);
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c0 + k < Cs && co0 + 4 * o4 < a.Cout)
            v = *reinterpret_cast<const float4*>(a.w + w_base + ((size_t)tap * Cs + c0 + k) * a.Cout + co0 + 4 * o4);
          *reinterpret_cast<float4*>(s_w + (tap * CK + k) * COT + 4 * o4) = v;
        }
        __syncthreads();
#pragma unroll 2
        for (int k = 0; k < CK; ++k) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const float* row = s_in + k * PH * PW + (py + ky) * PW + px;
            float4 i0 = *reinterpret_cast<const float4*>(row);
            float2 i1 = *reinterpret_cast<const float2*>(row + 4);
            float in[6] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y};
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const float4* w4 = reinterpret_cast<const float4*>(s_w + ((ky * 3 + kx) * CK + k) * COT + cg * 16);
#pragma unroll
              for (int o4 = 0; o4 < 4; ++o4) {
                float4 w = w4[o4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                  float x = in[p + kx];
                  acc[p][4 * o4 + 0] = fmaf(x, w.x, acc[p][4 * o4 + 0]);
                  acc[p][4 * o4 + 1] = fmaf(x, w.y, acc[p][4 * o4 + 1]);
                  acc[p][4 * o4 + 2] = fmaf(x, w.z, acc[p][4 * o4 + 2]);
                  acc[p][4 * o4 + 3] = fmaf(x, w.w, acc[p][4 * o4 + 3]);
                }
              }
            }
          }
        }
      }
    }
    w_base += (size_t)9 * Cs * a.Cout;
    w_row0 += Cs;
  }
  (void)w_row0;
  // epilogue
  const int gy = ty0 + py;
  if (gy >= a.H) return;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    int gx = tx0 + px + p;
    if (gx >= a.W) continue;
    float* o = a.out + (((size_t)n * a.H + gy) * a.W + gx) * a.out_cstride + a.out_coff + co0 + cg * 16;
#pragma unroll
    for (int o4 = 0; o4 < 4; ++o4) {
      int co = co0 + cg * 16 + 4 * o4;
      if (co >= a.Cout) continue;
      float r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v = acc[p][4 * o4 + u] + a.bias[co + u];
        if (a.bn_scale) v = fmaf(v, a.bn_scale[co + u], a.bn_shift[co + u]);
        r[u] = a.relu ? fmaxf(v, 0.f) : v;
      }
      if (co + 3 < a.Cout) {
        *reinterpret_cast<float4*>(o + 4 * o4) = make_float4(r[0], r[1], r[2], r[3]);
      } else {
        for (int u = 0; u < 4 && co + u < a.Cout; ++u) o[4 * o4 + u] = r[u];
      }
    }
  }
}

// ConvTranspose2d(kernel 2, stride 2): out[n, 2y+dy, 2x+dx, co] = b[co] + sum_ci in[n,y,x,ci] * W[dy*2+dx][ci][co]
// = a GEMM [pixels x Cin] . [Cin x (4 taps x Cout)] with a scattered store.  CTA: 128 threads own 64 input pixels x
// (4 taps x 32 couts); thread tile 8 pixels x 8 outputs (two float4 groups = two taps x 4 couts), 16 FMAs per 128-bit
// shared-memory load.  The activation chunk is transposed to channel-major on its way into shared memory (through
// registers, loaded one chunk ahead); the weight chunk streams in with cp.async, double-buffered.  Each output is one
// fmaf chain over ascending ci, then + bias.
constexpr int TK = 32;              // input channels per chunk
constexpr int TLD = 64 + 4;         // row stride of the transposed activation chunk
__global__ void __launch_bounds__(128, 3) k_convT2x2(const float* __restrict__ in, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ out, int npix_in,
                                                     int H, int W, int Cin, int Cout, int out_cstride, int out_coff) {
  extern __shared__ __align__(16) float s_dyn[];
  float(*s_a)[TK * TLD] = reinterpret_cast<float(*)[TK * TLD]>(s_dyn);                  // [2][TK*TLD]
  float(*s_b)[TK * 128] = reinterpret_cast<float(*)[TK * 128]>(s_dyn + 2 * TK * TLD);  // [2][TK*128]
  const int tid = threadIdx.x;
  const int tr = tid >> 4, tc = tid & 15;
  const int p0 = blockIdx.x * 64;
  const int co0 = blockIdx.y * 32;
  const int nch = Cin / TK;
  float acc[8][8];
#pragma unroll
  for (int o = 0; o < 8; ++o)
#pragma unroll
    for (int p = 0; p < 8; ++p) acc[o][p] = 0.f;

  // activation chunk: 64 px x 32 ch = 512 float4, four per thread (px = e / 8, k4 = e % 8)
  float4 areg[4];
  auto load_a = [&](int ch) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 128 * i, px = e >> 3, k4 = e & 7;
      areg[i] = (p0 + px < npix_in) ? *reinterpret_cast<const float4*>(in + (size_t)(p0 + px) * Cin + ch * TK + 4 * k4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_a = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 128 * i, px = e >> 3, k4 = e & 7;
      float* d = s_a[buf] + (4 * k4) * TLD + px;
      d[0] = areg[i].x, d[TLD] = areg[i].y, d[2 * TLD] = areg[i].z, d[3 * TLD] = areg[i].w;
    }
  };
  // weight chunk: [k][tap*32 + co] <- w[(tap*Cin + ch*TK + k)*Cout + co0 + co]; 32 k x 4 taps x 8 float4
  auto load_b = [&](int ch, int buf) {
    for (int e = tid; e < TK * 32; e += 128) {
      const int q = e & 7, t = (e >> 3) & 3, k = e >> 5;
      const float* src = w + ((size_t)t * Cin + ch * TK + k) * Cout + co0 + 4 * q;
      uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_b[buf] + k * 128 + t * 32 + 4 * q);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  load_a(0);
  load_b(0, 0);
  for (int ch = 0; ch < nch; ++ch) {
    const int buf = ch & 1;
    store_a(buf);
    if (ch + 1 < nch) load_a(ch + 1);  // global loads of the next chunk fly during this chunk's FMAs
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // chunk ch is in shared memory; every thread has finished chunk ch-1 (its buffers are reused next)
    if (ch + 1 < nch) load_b(ch + 1, buf ^ 1);
    const float* A = s_a[buf];
    const float* B = s_b[buf];
#pragma unroll 4
    for (int k = 0; k < TK; ++k) {
      const float4 xa = *reinterpret_cast<const float4*>(A + k * TLD + tr * 8);
      const float4 xb = *reinterpret_cast<const float4*>(A + k * TLD + tr * 8 + 4);
      const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const float4 wq = *reinterpret_cast<const float4*>(B + k * 128 + g * 64 + tc * 4);
        const float wv[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int p = 0; p < 8; ++p) acc[4 * g + u][p] = fmaf(xv[p], wv[u], acc[4 * g + u][p]);
      }
    }
  }
  // thread's outputs: pixels p0 + tr*8 .. +7; group g -> tap = 2g + tc/8, couts co0 + (tc%8)*4 .. +3
  const int co = co0 + (tc & 7) * 4;
  const float4 bq = *reinterpret_cast<const float4*>(bias + co);
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int pix = p0 + tr * 8 + p;
    if (pix >= npix_in) continue;
    const int n = pix / (H * W), rem = pix % (H * W), y = rem / W, x = rem % W;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int t = 2 * g + (tc >> 3);
      const int oy = 2 * y + (t >> 1), ox = 2 * x + (t & 1);
      float* o = out + (((size_t)n * 2 * H + oy) * (2 * W) + ox) * out_cstride + out_coff + co;
      *reinterpret_cast<float4*>(o) = make_float4(acc[4 * g][p] + bq.x, acc[4 * g + 1][p] + bq.y, acc[4 * g + 2][p] + bq.z,
                                                  acc[4 * g + 3][p] + bq.w);
    }
  }
}

template <bool P16>
__global__ void k_maxpool2(const float* __restrict__ in, float* __restrict__ out, int N, int H, int W, int C) {
  int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  long long total = (long long)N * Ho * Wo * C4;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += stride) {
    int c = (int)(e % C4);
    long long r = e / C4;
    int x = (int)(r % Wo);
    r /= Wo;
    int y = (int)(r % Ho);
    int n = (int)(r / Ho);
    const size_t p00 = ((size_t)n * H + 2 * y) * W + 2 * x;
    // (P16: h + l is exact in float32 and splits back to the same pair, so pooling pairs equals pooling values)
    const float4 a = p16::ld4<P16>(in, p00, C, 4 * c), b = p16::ld4<P16>(in, p00 + 1, C, 4 * c),
                 d = p16::ld4<P16>(in, p00 + W, C, 4 * c), f = p16::ld4<P16>(in, p00 + W + 1, C, 4 * c);
    float4 m = make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, f.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, f.y)),
                           fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, f.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, f.w)));
    p16::st4<P16>(out, ((size_t)n * Ho + y) * Wo + x, C, 4 * c, m);
  }
}

// max over the T frames of each scene: in [B*T, HW, C] -> out [B, HW, C]
template <bool P16>
__global__ void k_temporal_max(const float* __restrict__ in, float* __restrict__ out, int B, int T, long long hw, int C) {
  const int C4 = C / 4;
  long long total = (long long)B * hw * C4;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += stride) {
    const int c = (int)(e % C4);
    const long long pix = e / C4, b = pix / hw, r = pix % hw;
    float4 m = p16::ld4<P16>(in, (size_t)(b * T) * hw + r, C, 4 * c);
    for (int t = 1; t < T; ++t) {
      const float4 v = p16::ld4<P16>(in, (size_t)(b * T + t) * hw + r, C, 4 * c);
      m.x = fmaxf(m.x, v.x), m.y = fmaxf(m.y, v.y), m.z = fmaxf(m.z, v.z), m.w = fmaxf(m.w, v.w);
    }
    p16::st4<P16>(out, (size_t)pix, C, 4 * c, m);
  }
}

template <int TH, int TW, int COT>
int launch_conv(const ConvArgs& a, cudaStream_t stream) {
  constexpr int NT = (TH * TW / 4) * (COT / 16);
  size_t smem = (size_t)(CK * (TH + 2) * (TW + 4) + 9 * CK * COT) * sizeof(float);
  static PcabSmemOnce once;
  if (pcab_set_max_smem(k_conv3x3<TH, TW, COT>, (int)smem, once) != cudaSuccess) return -1;
  dim3 grid(cdiv(a.H, TH) * cdiv(a.W, TW), cdiv(a.Cout, COT), a.N);
  k_conv3x3<TH, TW, COT><<<grid, NT, smem, stream>>>(a);
  return 0;
}

}  // namespace

extern "C" int pcab_conv3x3_f32(const float* src0, int c0, const float* src1, int c1, const float* src2, int c2,
                                int temporal_T, const float* weight_packed, const float* bias, const float* bn_scale,
                                const float* bn_shift, int relu, float* out, int n_images, int H, int W, int Cout,
                                int out_cstride, int out_coff, cudaStream_t stream) {
  ConvArgs a;
  a.src[0] = src0, a.src[1] = src1, a.src[2] = src2;
  a.src_c[0] = c0, a.src_c[1] = c1, a.src_c[2] = c2;
  a.nsrc = src2 ? 3 : (src1 ? 2 : 1);
  a.T = temporal_T > 1 ? temporal_T : 1;
  if (a.T > 1) {
    PCAB_REQUIRE(a.nsrc == 3 && src0 == src1 && src1 == src2, "temporal mode takes the same tensor three times");
    a.frame_shift[0] = -1, a.frame_shift[1] = 0, a.frame_shift[2] = 1;
  } else {
    a.frame_shift[0] = a.frame_shift[1] = a.frame_shift[2] = 0;
  }
  PCAB_REQUIRE(c0 % 4 == 0 && c1 % 4 == 0 && c2 % 4 == 0, "channel counts must be multiples of 4");
  PCAB_REQUIRE(out_cstride % 4 == 0 && out_coff % 4 == 0 || Cout < 4, "output channel layout must be 16B aligned");
  a.w = weight_packed, a.bias = bias, a.bn_scale = bn_scale, a.bn_shift = bn_shift, a.relu = relu;
  a.out = out, a.N = n_images, a.H = H, a.W = W, a.Cout = Cout, a.out_cstride = out_cstride, a.out_coff = out_coff;
  if (H * W >= 96 * 96 && Cout <= 32)
    launch_conv<16, 32, 32>(a, stream);
  else if (H * W >= 64 * 64)
    launch_conv<8, 32, 64>(a, stream);
  else
    launch_conv<8, 16, 64>(a, stream);
  PCAB_CHECK_LAUNCH("pcab_conv3x3_f32");
  return PCAB_OK;
}

extern "C" int pcab_convT2x2_f32(const float* in, const float* weight_packed, const float* bias, float* out,
                                 int n_images, int H, int W, int Cin, int Cout, int out_cstride, int out_coff,
                                 cudaStream_t stream) {
  PCAB_REQUIRE(Cin % 32 == 0 && Cout % 32 == 0, "Cin%32, Cout%32");
  PCAB_REQUIRE(out_cstride % 4 == 0 && out_coff % 4 == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)in & 15) == 0 &&
                   ((uintptr_t)weight_packed & 15) == 0 && ((uintptr_t)bias & 15) == 0,
               "16B alignment of in / weights / bias / output channel layout");
  int npix = n_images * H * W;
  dim3 grid(cdiv(npix, 64), Cout / 32);
  const size_t smem = (size_t)(2 * TK * TLD + 2 * TK * 128) * sizeof(float);
  static PcabSmemOnce once;
  PCAB_CUDA(pcab_set_max_smem(k_convT2x2, (int)smem, once));
  k_convT2x2<<<grid, 128, smem, stream>>>(in, weight_packed, bias, out, npix, H, W, Cin, Cout, out_cstride, out_coff);
  PCAB_CHECK_LAUNCH("pcab_convT2x2_f32");
  return PCAB_OK;
}

extern "C" int pcab_maxpool2x2(const float* in, float* out, int n_images, int H, int W, int C, int fmt, cudaStream_t stream) {
  PCAB_REQUIRE(C % 4 == 0 && H % 2 == 0 && W % 2 == 0 && (!fmt || C % 32 == 0), "C%4 (P16: C%32), even H/W");
  long long total = (long long)n_images * (H / 2) * (W / 2) * (C / 4);
  if (fmt)
    k_maxpool2<true><<<grid_for(total, 256), 256, 0, stream>>>(in, out, n_images, H, W, C);
  else
    k_maxpool2<false><<<grid_for(total, 256), 256, 0, stream>>>(in, out, n_images, H, W, C);
  PCAB_CHECK_LAUNCH("pcab_maxpool2x2");
  return PCAB_OK;
}

extern "C" int pcab_temporal_max(const float* in, float* out, int B, int T, int H, int W, int C, int fmt, cudaStream_t stream) {
  PCAB_REQUIRE(C % 4 == 0 && (!fmt || C % 32 == 0), "C%4 (P16: C%32)");
  long long hw = (long long)H * W, total = (long long)B * hw * (C / 4);
  if (fmt)
    k_temporal_max<true><<<grid_for(total, 256), 256, 0, stream>>>(in, out, B, T, hw, C);
  else
    k_temporal_max<false><<<grid_for(total, 256), 256, 0, stream>>>(in, out, B, T, hw, C);
  PCAB_CHECK_LAUNCH("pcab_temporal_max");
  return PCAB_OK;
}
