// Block-level dense layers for per-point MLPs: a CTA of NT = 128 threads owns PTS = 64 rows whose
// activations live in shared memory CHANNEL-MAJOR ([C][LDP], LDP = PTS + 4); weights stream through a double-buffered
// shared-memory k-chunk filled with cp.async (the next chunk lands while the current one is consumed).  Thread tile is
// 8 rows x (OUT/16) outputs: per k two 128-bit activation loads (8 rows, warp-broadcast) and OUT/64 128-bit weight
// loads (a quarter warp reads 128 contiguous bytes: conflict-free) feed 8*OUT/16 FMAs - 16 FMAs per shared-memory
// load at OUT = 128, so the FP32 pipe, not the LSU, is the limiter.  Every output is one fmaf chain over ascending k
// (then + bias), i.e. the same rounding sequence whatever the tiling.
#pragma once
#include "common.cuh"

namespace mlp {

constexpr int NT = 128;   // threads per CTA of every kernel built on block_dense
constexpr int PTS = 64;
constexpr int LDP = PTS + 4;  // row stride of the channel-major activation buffers (keeps float4 alignment, spreads banks)
constexpr int KC = 32;  // weight rows staged per chunk
constexpr int SW_FLOATS = 2 * KC * 128;  // shared scratch block_dense needs (two chunks of the widest layer)

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Y[OUT][PTS] = epilogue(W[IN][OUT]^T applied to X[IN][PTS] + b).  X, Y: shared memory, channel-major.
// W is [in][out] row-major in global memory (16 B aligned).  epilogue: optional BatchNorm(eval) scale/shift, optional
// ReLU.  s_w: shared scratch of SW_FLOATS floats.  All NT threads must call.  X and Y must not alias.
template <int IN, int OUT>
__device__ __forceinline__ void block_dense(const float* X, const float* __restrict__ W, const float* __restrict__ b,
                                            const float* __restrict__ scale, const float* __restrict__ shift, bool relu,
                                            float* Y, float* s_w) {
  constexpr int OPT = OUT / 16;         // outputs per thread: 8 (two float4 groups 64 apart), 4 (one float4) or 2 (float2)
  constexpr int NG = OPT >= 4 ? OPT / 4 : 1;
  constexpr int NCH = (IN + KC - 1) / KC;
  static_assert(OUT == 32 || OUT == 64 || OUT == 128, "OUT must be 32, 64 or 128");
  const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;
  float acc[OPT][8];
#pragma unroll
  for (int o = 0; o < OPT; ++o)
#pragma unroll
    for (int p = 0; p < 8; ++p) acc[o][p] = 0.f;
  auto stage = [&](int ch) {
    const int k0 = ch * KC;
    const int kc = (IN - k0) < KC ? (IN - k0) : KC;
    float* dst = s_w + (ch & 1) * KC * OUT;
    const float* src = W + (size_t)k0 * OUT;
    for (int e = threadIdx.x; e < kc * OUT / 4; e += NT) cp_async16(dst + 4 * e, src + 4 * e);
    cp_async_commit();
  };
  stage(0);  // callers synchronise after writing X, and every block_* helper ends with a barrier: s_w is free here
#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    cp_async_wait_all();
    __syncthreads();  // chunk ch visible to everyone; everyone is done with chunk ch-1, whose buffer chunk ch+1 reuses
    if (ch + 1 < NCH) stage(ch + 1);
    const int k0 = ch * KC;
    const int kc = (IN - k0) < KC ? (IN - k0) : KC;
    const float* wbuf = s_w + (ch & 1) * KC * OUT;
#pragma unroll 4
    for (int k = 0; k < kc; ++k) {
      const float4 xa = *reinterpret_cast<const float4*>(X + (k0 + k) * LDP + tr * 8);
      const float4 xb = *reinterpret_cast<const float4*>(X + (k0 + k) * LDP + tr * 8 + 4);
      const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
      if constexpr (OPT >= 4) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const float4 w = *reinterpret_cast<const float4*>(wbuf + k * OUT + g * 64 + tc * 4);
          const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int p = 0; p < 8; ++p) acc[4 * g + u][p] = fmaf(xv[p], wv[u], acc[4 * g + u][p]);
        }
      } else {
        const float2 w = *reinterpret_cast<const float2*>(wbuf + k * OUT + tc * 2);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          acc[0][p] = fmaf(xv[p], w.x, acc[0][p]);
          acc[1][p] = fmaf(xv[p], w.y, acc[1][p]);
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < OPT; ++o) {
    const int c = OPT >= 4 ? (o >> 2) * 64 + tc * 4 + (o & 3) : tc * 2 + o;
    float bb = b ? b[c] : 0.f;
    float sc = scale ? scale[c] : 1.f, sh = scale ? shift[c] : 0.f;
    float v[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      float t = acc[o][p] + bb;
      if (scale) t = fmaf(t, sc, sh);
      v[p] = relu ? fmaxf(t, 0.f) : t;
    }
    *reinterpret_cast<float4*>(Y + c * LDP + tr * 8) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(Y + c * LDP + tr * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  __syncthreads();
}

// small output width (OUT <= 8): one thread per (row, out); Y is [OUT][PTS] too
template <int IN, int OUT>
__device__ __forceinline__ void block_dense_small(const float* X, const float* __restrict__ W,
                                                  const float* __restrict__ b, float* Y) {
  for (int e = threadIdx.x; e < PTS * OUT; e += NT) {
    int p = e % PTS, o = e / PTS;
    float a = 0.f;
    for (int k = 0; k < IN; ++k) a = fmaf(X[k * LDP + p], W[(size_t)k * OUT + o], a);
    Y[o * LDP + p] = a + b[o];
  }
  __syncthreads();
}

// float atomic max through the integer ordering trick (destination initialised to -inf)
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// border-padded bilinear sample of an NHWC map (grid_sample, align_corners=False) -> 4 taps + weights
struct Bilinear {
  int o00, o01, o10, o11;  // pixel offsets (y*W+x)
  float w00, w01, w10, w11;
};
__device__ __forceinline__ Bilinear bilinear_border(float px, float py, float x_abs, float y_abs, int H, int W) {
  float u = px / x_abs, v = py / y_abs;
  float ix = ((u + 1.f) * W - 1.f) / 2.f;
  float iy = ((v + 1.f) * H - 1.f) / 2.f;
  ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
  iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
  float fx0 = floorf(ix), fy0 = floorf(iy);
  int x0 = (int)fx0, y0 = (int)fy0;
  float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = (fx0 + 1.f) - ix, wy0 = (fy0 + 1.f) - iy;
  bool vx1 = x0 + 1 < W, vy1 = y0 + 1 < H;
  int x1 = vx1 ? x0 + 1 : x0, y1 = vy1 ? y0 + 1 : y0;
  Bilinear r;
  r.o00 = y0 * W + x0, r.o01 = y0 * W + x1, r.o10 = y1 * W + x0, r.o11 = y1 * W + x1;
  r.w00 = wx0 * wy0;
  r.w01 = vx1 ? wx1 * wy0 : 0.f;
  r.w10 = vy1 ? wx0 * wy1 : 0.f;
  r.w11 = (vx1 && vy1) ? wx1 * wy1 : 0.f;
  return r;
}

// STPN point-head weight pack ([in][out] matrices), floats:
//   pe0 W[3][32] b[32] | pe2 W[32][64] b[64] | fp W[128][128] b[128] |
//   mos0 W[128][128] b[128] s[128] t[128] | mos3 W[128][2] b[2] | off0 W,b,s,t | off3 W[128][2] b[2]
namespace stpn_pack {
constexpr int S_PE0W = 0, S_PE0B = S_PE0W + 3 * 32, S_PE2W = S_PE0B + 32, S_PE2B = S_PE2W + 32 * 64;
constexpr int S_FPW = S_PE2B + 64, S_FPB = S_FPW + 128 * 128;
constexpr int S_M0W = S_FPB + 128, S_M0B = S_M0W + 128 * 128, S_M0S = S_M0B + 128, S_M0T = S_M0S + 128;
constexpr int S_M3W = S_M0T + 128, S_M3B = S_M3W + 128 * 2;
constexpr int S_O0W = S_M3B + 2 + 2 /*pad to 4*/, S_O0B = S_O0W + 128 * 128, S_O0S = S_O0B + 128, S_O0T = S_O0S + 128;
constexpr int S_O3W = S_O0T + 128, S_O3B = S_O3W + 128 * 2;
constexpr int S_PACK = S_O3B + 2 + 2;
}  // namespace stpn_pack

}  // namespace mlp
