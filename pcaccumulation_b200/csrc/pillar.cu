// Per-pillar PointNet (PillarFeatureNet) over points sorted by pillar + dense canvas scatter.
//
// Replaces models/pillar_encoder.py:97-122 (PillarFeatureNet.forward, ResnetBlockFC) and
// models/pillar_encoder.py:125-174 (scatter_point_pillar).  Points are processed in pillar-sorted
// order (pcab_pillar_index) so every pillar's points are contiguous: the three segment-max poolings
// become contiguous-range reductions (no atomics) and all per-point activations stay in registers
// inside a stage; only the 32-wide block outputs cross HBM between stages (coalesced, sorted order).
//
// Weight pack (floats, all matrices stored [in][out] so a warp reads one broadcast row per k):
//   fc_pos W[9][64] b[64] | for blk in 0..2: fc_0 W[64][32] b[32], fc_1 W[32][32] b[32], shortcut W[64][32] |
//   fc_c W[32][32] b[32]
#include "common.cuh"
#include "pcab200.h"

namespace {

constexpr int kPosW = 0;
constexpr int kPosB = kPosW + 9 * 64;
constexpr int kBlk0 = kPosB + 64;
constexpr int kBlkSize = 64 * 32 + 32 + 32 * 32 + 32 + 64 * 32;
constexpr int kFcC = kBlk0 + 3 * kBlkSize;
constexpr int kPackSize = kFcC + 32 * 32 + 32;

template <int IN, int OUT, bool RELU_IN>
__device__ __forceinline__ void dense(const float* __restrict__ W, const float* __restrict__ b, const float (&x)[IN],
                                      float (&y)[OUT]) {
#pragma unroll
  for (int o = 0; o < OUT; ++o) y[o] = b ? b[o] : 0.f;
#pragma unroll
  for (int k = 0; k < IN; ++k) {
    float xv = RELU_IN ? fmaxf(x[k], 0.f) : x[k];
    const float4* w4 = reinterpret_cast<const float4*>(W + k * OUT);
#pragma unroll
    for (int o4 = 0; o4 < OUT / 4; ++o4) {
      float4 w = w4[o4];
      y[4 * o4 + 0] = fmaf(xv, w.x, y[4 * o4 + 0]);
      y[4 * o4 + 1] = fmaf(xv, w.y, y[4 * o4 + 1]);
      y[4 * o4 + 2] = fmaf(xv, w.z, y[4 * o4 + 2]);
      y[4 * o4 + 3] = fmaf(xv, w.w, y[4 * o4 + 3]);
    }
  }
}

// ResnetBlockFC 64 -> 32 (pre-activation; models/pillar_encoder.py:46-55)
__device__ __forceinline__ void resblock(const float* __restrict__ Wb, const float (&x)[64], float (&out)[32]) {
  const float* W0 = Wb;
  const float* b0 = W0 + 64 * 32;
  const float* W1 = b0 + 32;
  const float* b1 = W1 + 32 * 32;
  const float* Ws = b1 + 32;
  float net[32];
  dense<64, 32, true>(W0, b0, x, net);
  float dx[32];
  dense<32, 32, true>(W1, b1, net, dx);
  dense<64, 32, false>(Ws, nullptr, x, out);
#pragma unroll
  for (int o = 0; o < 32; ++o) out[o] += dx[o];
}

struct PfnGeom {
  double vx, vy, x_off, y_off;
  float scale, n_frames;
};

__global__ void __launch_bounds__(128) k_pfn_stage0(const float* __restrict__ xyz, const int* __restrict__ ptime,
                                                    const int* __restrict__ order, const int* __restrict__ p2v,
                                                    const int* __restrict__ coords, const float* __restrict__ pmean,
                                                    const float* __restrict__ pack, int n, PfnGeom g,
                                                    float* __restrict__ net_out) {
  extern __shared__ float sw[];
  for (int i = threadIdx.x; i < kBlk0 + kBlkSize; i += blockDim.x) sw[i] = pack[i];
  __syncthreads();
  int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    int i = order[j];
    int m = p2v[i];
    float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
    float f[9];
    f[0] = px, f[1] = py, f[2] = pz;
    f[3] = __fsub_rn(px, pmean[3 * m]);
    f[4] = __fsub_rn(py, pmean[3 * m + 1]);
    f[5] = __fsub_rn(pz, pmean[3 * m + 2]);
    int4 c = reinterpret_cast<const int4*>(coords)[m];  // z, y, x, t
    f[6] = (float)((double)px - ((double)c.z * g.vx + g.x_off));
    f[7] = (float)((double)py - ((double)c.y * g.vy + g.y_off));
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = __fdiv_rn(f[k], g.scale);
    f[8] = __fdiv_rn((float)ptime[i], g.n_frames);
    float x[64];
    dense<9, 64, false>(sw + kPosW, sw + kPosB, f, x);
    float out[32];
    resblock(sw + kBlk0, x, out);
    float4* dst = reinterpret_cast<float4*>(net_out + (size_t)j * 32);
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
  }
}

// pooled[m][c] = max over the pillar's points.  Eight lanes per pillar (one float4 of channels each), four pillars per
// warp, rows of a pillar loaded two at a time: a pillar holds ~3 points, so the kernel is bound by the number of
// independent loads in flight, not by arithmetic.
__global__ void k_segmax32(const float* __restrict__ net, const int* __restrict__ pstart, int m,
                           float* __restrict__ pooled, const int* __restrict__ cell_of_pillar,
                           float* __restrict__ canvas) {
  const int sub = threadIdx.x & 7;
  const int group = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int ngroup = (gridDim.x * blockDim.x) >> 3;
  const float4* net4 = reinterpret_cast<const float4*>(net);
  for (int p = group; p < m; p += ngroup) {
    const int s = pstart[p], e = pstart[p + 1];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // torch_scatter leaves empty segments at 0
    if (s < e) {
      v = net4[(size_t)s * 8 + sub];
      int j = s + 1;
      for (; j + 1 < e; j += 2) {
        const float4 a = net4[(size_t)j * 8 + sub], b = net4[(size_t)(j + 1) * 8 + sub];
        v.x = fmaxf(v.x, fmaxf(a.x, b.x)), v.y = fmaxf(v.y, fmaxf(a.y, b.y));
        v.z = fmaxf(v.z, fmaxf(a.z, b.z)), v.w = fmaxf(v.w, fmaxf(a.w, b.w));
      }
      if (j < e) {
        const float4 a = net4[(size_t)j * 8 + sub];
        v.x = fmaxf(v.x, a.x), v.y = fmaxf(v.y, a.y), v.z = fmaxf(v.z, a.z), v.w = fmaxf(v.w, a.w);
      }
    }
    reinterpret_cast<float4*>(pooled)[(size_t)p * 8 + sub] = v;
    if (canvas) reinterpret_cast<float4*>(canvas)[(size_t)cell_of_pillar[p] * 8 + sub] = v;
  }
}

template <bool FINAL>
__global__ void __launch_bounds__(128) k_pfn_block(const float* __restrict__ net_in, const float* __restrict__ pooled,
                                                   const int* __restrict__ order, const int* __restrict__ p2v,
                                                   const float* __restrict__ pack, int blk, int n,
                                                   float* __restrict__ net_out) {
  extern __shared__ float sw[];
  const float* src = pack + kBlk0 + blk * kBlkSize;
  for (int i = threadIdx.x; i < kBlkSize; i += blockDim.x) sw[i] = src[i];
  if (FINAL)
    for (int i = threadIdx.x; i < 32 * 32 + 32; i += blockDim.x) sw[kBlkSize + i] = pack[kFcC + i];
  __syncthreads();
  int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    int m = p2v[order[j]];
    float x[64];
    const float4* a = reinterpret_cast<const float4*>(net_in + (size_t)j * 32);
    const float4* b = reinterpret_cast<const float4*>(pooled + (size_t)m * 32);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 v = a[q];
      x[4 * q] = v.x, x[4 * q + 1] = v.y, x[4 * q + 2] = v.z, x[4 * q + 3] = v.w;
      float4 u = b[q];
      x[32 + 4 * q] = u.x, x[32 + 4 * q + 1] = u.y, x[32 + 4 * q + 2] = u.z, x[32 + 4 * q + 3] = u.w;
    }
    float out[32];
    resblock(sw, x, out);
    if (FINAL) {
      float y[32];
      dense<32, 32, false>(sw + kBlkSize, sw + kBlkSize + 32 * 32, out, y);
#pragma unroll
      for (int o = 0; o < 32; ++o) out[o] = y[o];
    }
    float4* dst = reinterpret_cast<float4*>(net_out + (size_t)j * 32);
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
  }
}

// cell index of each pillar in the [B*T, Ny, Nx] canvas (models/pillar_encoder.py:158)
__global__ void k_pillar_cells(const int* __restrict__ coords, const int* __restrict__ pbatch, int m, int T, int ny,
                               int nx, int* __restrict__ cell, int* __restrict__ pframe,
                               int* __restrict__ cell2pillar) {
  int stride = gridDim.x * blockDim.x;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < m; p += stride) {
    int4 c = reinterpret_cast<const int4*>(coords)[p];
    int idx = ((pbatch[p] * T + c.w) * ny + c.y) * nx + c.z;
    cell[p] = idx;
    if (pframe) pframe[p] = pbatch[p] * T + c.w;
    if (cell2pillar) cell2pillar[idx] = p;
  }
}

}  // namespace

extern "C" int pcab_pfn_pack_size(void) { return kPackSize; }

extern "C" int pcab_pillar_cells(const int* coords_zyxt, const int* pillar_batch, int n_pillars, int n_sweeps, int ny,
                                 int nx, int* pillar_cell, int* pillar_frame, int* cell_to_pillar,
                                 cudaStream_t stream) {
  k_pillar_cells<<<grid_for(n_pillars, 256), 256, 0, stream>>>(coords_zyxt, pillar_batch, n_pillars, n_sweeps, ny, nx,
                                                               pillar_cell, pillar_frame, cell_to_pillar);
  PCAB_CHECK_LAUNCH("pcab_pillar_cells");
  return PCAB_OK;
}

// scratch: 2 * n_points * 32 floats (net ping/pong) + n_pillars * 32 floats (pooled)
extern "C" size_t pcab_pillar_encode_workspace(int n_points, int n_pillars) {
  return ((size_t)n_points * 64 + (size_t)n_pillars * 32) * sizeof(float) + 512;
}

extern "C" int pcab_pillar_encode(const float* xyz, const int* point_time, const int* order, const int* p2v,
                                  const int* pstart, const int* coords_zyxt, const int* pillar_cell,
                                  const float* pillar_mean, const float* weight_pack, int n_points, int n_pillars,
                                  const float* range6, const float* voxel_size3, int n_sweeps, float* pillar_feats,
                                  float* canvas_nhwc, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PCAB_REQUIRE(workspace_bytes >= pcab_pillar_encode_workspace(n_points, n_pillars), "workspace too small");
  float* net_a = (float*)workspace;
  float* net_b = net_a + (size_t)n_points * 32;
  float* pooled = net_b + (size_t)n_points * 32;
  PfnGeom g;
  g.vx = voxel_size3[0];
  g.vy = voxel_size3[1];
  g.x_off = g.vx / 2 + (double)range6[0];
  g.y_off = g.vy / 2 + (double)range6[1];
  g.scale = fabsf(range6[0]);
  g.n_frames = (float)n_sweeps;
  const int B = 128;
  int gp = grid_for(n_points, B, 8);
  int gw = grid_for((long long)n_pillars * 8, 256, 8);
  size_t sm0 = (size_t)(kBlk0 + kBlkSize) * 4, sm1 = (size_t)kBlkSize * 4, sm2 = (size_t)(kBlkSize + 32 * 32 + 32) * 4;
  k_pfn_stage0<<<gp, B, sm0, stream>>>(xyz, point_time, order, p2v, coords_zyxt, pillar_mean, weight_pack, n_points, g,
                                       net_a);
  k_segmax32<<<gw, 256, 0, stream>>>(net_a, pstart, n_pillars, pooled, nullptr, nullptr);
  k_pfn_block<false><<<gp, B, sm1, stream>>>(net_a, pooled, order, p2v, weight_pack, 1, n_points, net_b);
  k_segmax32<<<gw, 256, 0, stream>>>(net_b, pstart, n_pillars, pooled, nullptr, nullptr);
  k_pfn_block<true><<<gp, B, sm2, stream>>>(net_b, pooled, order, p2v, weight_pack, 2, n_points, net_a);
  k_segmax32<<<gw, 256, 0, stream>>>(net_a, pstart, n_pillars, pillar_feats, pillar_cell, canvas_nhwc);
  PCAB_CHECK_LAUNCH("pcab_pillar_encode");
  return PCAB_OK;
}
