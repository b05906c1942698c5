// Block-level dense layers for per-point MLPs: a CTA of 256 threads owns PTS = 64 rows whose
// activations live in shared memory CHANNEL-MAJOR ([C][LDP], LDP = PTS + 4); weights stream through a shared-memory
// k-chunk.  Thread tile is 4 rows x (OUT/16) contiguous outputs: per k one 128-bit activation load
// (4 rows) and OUT/64 128-bit weight loads feed 4*OUT/16 FMAs.
#pragma once
#include "common.cuh"

namespace mlp {

constexpr int PTS = 64;
constexpr int LDP = PTS + 4;  // row stride of the channel-major activation buffers (keeps float4 alignment, spreads banks)
constexpr int KC = 32;  // weight rows staged per chunk

// Y[OUT][PTS] = epilogue(W[IN][OUT]^T applied to X[IN][PTS] + b).  X, Y: shared memory, channel-major.
// W is [in][out] row-major in global memory.  epilogue: optional BatchNorm(eval) scale/shift, optional ReLU.
// s_w: shared scratch of KC*OUT floats.  All 256 threads must call.  X and Y must not alias.
template <int IN, int OUT>
__device__ __forceinline__ void block_dense(const float* X, const float* __restrict__ W, const float* __restrict__ b,
                                            const float* __restrict__ scale, const float* __restrict__ shift, bool relu,
                                            float* Y, float* s_w) {
  constexpr int OPT = OUT / 16;
  static_assert(OUT % 32 == 0, "OUT must be a multiple of 32");
  const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;
  float acc[OPT][4];
#pragma unroll
  for (int o = 0; o < OPT; ++o)
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[o][p] = 0.f;
  for (int k0 = 0; k0 < IN; k0 += KC) {
    const int kc = (IN - k0) < KC ? (IN - k0) : KC;
    __syncthreads();
    for (int e = threadIdx.x; e < kc * OUT / 4; e += 256)
      reinterpret_cast<float4*>(s_w)[e] = reinterpret_cast<const float4*>(W + (size_t)k0 * OUT)[e];
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < kc; ++k) {
      float4 x = *reinterpret_cast<const float4*>(X + (k0 + k) * LDP + tr * 4);
      const float* wrow = s_w + k * OUT + tc * OPT;
      if constexpr (OPT % 4 == 0) {
#pragma unroll
        for (int o4 = 0; o4 < OPT / 4; ++o4) {
          float4 w = *reinterpret_cast<const float4*>(wrow + 4 * o4);
          const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc[4 * o4 + u][0] = fmaf(x.x, wv[u], acc[4 * o4 + u][0]);
            acc[4 * o4 + u][1] = fmaf(x.y, wv[u], acc[4 * o4 + u][1]);
            acc[4 * o4 + u][2] = fmaf(x.z, wv[u], acc[4 * o4 + u][2]);
            acc[4 * o4 + u][3] = fmaf(x.w, wv[u], acc[4 * o4 + u][3]);
          }
        }
      } else {
#pragma unroll
        for (int o2 = 0; o2 < OPT / 2; ++o2) {
          float2 w = *reinterpret_cast<const float2*>(wrow + 2 * o2);
          acc[2 * o2][0] = fmaf(x.x, w.x, acc[2 * o2][0]);
          acc[2 * o2][1] = fmaf(x.y, w.x, acc[2 * o2][1]);
          acc[2 * o2][2] = fmaf(x.z, w.x, acc[2 * o2][2]);
          acc[2 * o2][3] = fmaf(x.w, w.x, acc[2 * o2][3]);
          acc[2 * o2 + 1][0] = fmaf(x.x, w.y, acc[2 * o2 + 1][0]);
          acc[2 * o2 + 1][1] = fmaf(x.y, w.y, acc[2 * o2 + 1][1]);
          acc[2 * o2 + 1][2] = fmaf(x.z, w.y, acc[2 * o2 + 1][2]);
          acc[2 * o2 + 1][3] = fmaf(x.w, w.y, acc[2 * o2 + 1][3]);
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < OPT; ++o) {
    int c = tc * OPT + o;
    float bb = b ? b[c] : 0.f;
    float sc = scale ? scale[c] : 1.f, sh = scale ? shift[c] : 0.f;
    float v[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float t = acc[o][p] + bb;
      if (scale) t = fmaf(t, sc, sh);
      v[p] = relu ? fmaxf(t, 0.f) : t;
    }
    *reinterpret_cast<float4*>(Y + c * LDP + tr * 4) = make_float4(v[0], v[1], v[2], v[3]);
  }
  __syncthreads();
}

// small output width (OUT <= 8): one thread per (row, out); Y is [OUT][PTS] too
template <int IN, int OUT>
__device__ __forceinline__ void block_dense_small(const float* X, const float* __restrict__ W,
                                                  const float* __restrict__ b, float* Y) {
  for (int e = threadIdx.x; e < PTS * OUT; e += 256) {
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

}  // namespace mlp
