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
#include "pair16.cuh"
#include "tc_common.cuh"
#include "pcab200.h"

namespace {

constexpr int kPosW = 0;
constexpr int kPosB = kPosW + 9 * 64;
constexpr int kBlk0 = kPosB + 64;
constexpr int kBlkSize = 64 * 32 + 32 + 32 * 32 + 32 + 64 * 32;
constexpr int kFcC = kBlk0 + 3 * kBlkSize;
constexpr int kPackSize = kFcC + 32 * 32 + 32;

// ---------------------------------------------------------------------------------------------------------------
// Tiled formulation.  A CTA of 128 threads owns tiles of PT = 128 pillar-sorted points (persistent: the stage's weights
// are loaded into shared memory once per CTA).  Activations sit in shared memory channel-major ([C][LD]); a thread
// computes 8 points x (OUT/8) outputs, so one 128-bit weight load feeds 32 FMAs and the FP32 pipe is the limiter (the
// earlier thread-per-point kernels issued one broadcast weight load per 4 FMAs and were bound by the LSU).  Every
// output is the same rounding sequence as before: accumulator initialised with the bias, one fmaf per input channel in
// ascending order.
// The segment max over a pillar's points is fused: points are sorted by pillar, so a tile holds a few runs; a run that
// lies strictly inside one thread's scan range is written with a plain store, runs touching a range edge go through
// atomic max (exact and order-independent for floats) into a buffer pre-filled with -inf.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PT = 128;       // points per tile
constexpr int LD = PT + 4;    // channel-major row stride
constexpr int NTP = 256;      // threads per CTA
constexpr int RPT = PT * 8 / NTP;  // points per thread tile (4): 16 resident warps per SM hide the shared-memory latency

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// acc[o][p]: thread (tr = tid>>3, tc = tid&7) owns points tr*RPT+p and outputs g*32 + 4*tc + u  (o = 4*g + u)
template <int IN, int OUT, bool RELU_IN>
__device__ __forceinline__ void tile_dense(const float* __restrict__ X, const float* __restrict__ W,
                                           const float* __restrict__ b, float (&acc)[OUT / 8][RPT], int tr, int tc) {
  constexpr int NG = OUT / 32;
  static_assert(RPT == 4, "one 128-bit activation load per k");
#pragma unroll
  for (int o = 0; o < OUT / 8; ++o) {
    const float bb = b ? b[(o >> 2) * 32 + 4 * tc + (o & 3)] : 0.f;
#pragma unroll
    for (int p = 0; p < RPT; ++p) acc[o][p] = bb;
  }
#pragma unroll 8
  for (int k = 0; k < IN; ++k) {
    const float4 xa = *reinterpret_cast<const float4*>(X + k * LD + tr * RPT);
    float xv[RPT] = {xa.x, xa.y, xa.z, xa.w};
    if (RELU_IN) {
#pragma unroll
      for (int p = 0; p < RPT; ++p) xv[p] = fmaxf(xv[p], 0.f);
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const float4 w = *reinterpret_cast<const float4*>(W + k * OUT + g * 32 + 4 * tc);
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int p = 0; p < RPT; ++p) acc[4 * g + u][p] = fmaf(xv[p], wv[u], acc[4 * g + u][p]);
    }
  }
}

template <int OUT>
__device__ __forceinline__ void tile_store(float* __restrict__ Y, const float (&acc)[OUT / 8][RPT], int tr, int tc) {
#pragma unroll
  for (int o = 0; o < OUT / 8; ++o) {
    float* y = Y + ((o >> 2) * 32 + 4 * tc + (o & 3)) * LD + tr * RPT;
    *reinterpret_cast<float4*>(y) = make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
  }
}

struct PfnGeom {
  double vx, vy, x_off, y_off;
  float scale, n_frames;
};

// STAGE 0: features -> fc_pos -> block 0;  STAGE 1: [net, pooled] -> block 1;  STAGE 2: [net, pooled] -> block 2 -> fc_c
template <int STAGE>
__global__ void __launch_bounds__(NTP, 2)
k_pfn_tile(const float* __restrict__ xyz, const int* __restrict__ ptime, const int* __restrict__ order,
           const int* __restrict__ p2v, const int* __restrict__ coords, const float* __restrict__ pmean,
           const float* __restrict__ net_in, const float* __restrict__ pooled_in, const float* __restrict__ pack, int n,
           PfnGeom g, float* __restrict__ net_out, float* __restrict__ pooled_out) {
  extern __shared__ __align__(16) float sm[];
  float* X = sm;                 // [64][LD]  block input; reused for the block output tile
  float* NETs = X + 64 * LD;     // [32][LD]  relu(fc_0) ; stage 0: the 9 input features ; stage 2: fc_c output tile
  float* Wb = NETs + 32 * LD;    // block weights (kBlkSize floats)
  float* Wx = Wb + kBlkSize;     // stage 0: fc_pos W,b (640) ; stage 2: fc_c W,b (1056)
  __shared__ int s_pil[PT];
  const int tid = threadIdx.x, tr = tid >> 3, tc = tid & 7;
  {
    const float* src = pack + kBlk0 + STAGE * kBlkSize;
    for (int i = tid; i < kBlkSize / 4; i += NTP) reinterpret_cast<float4*>(Wb)[i] = reinterpret_cast<const float4*>(src)[i];
    if (STAGE == 0)
      for (int i = tid; i < (9 * 64 + 64) / 4; i += NTP) reinterpret_cast<float4*>(Wx)[i] = reinterpret_cast<const float4*>(pack + kPosW)[i];
    if (STAGE == 2)
      for (int i = tid; i < (32 * 32 + 32) / 4; i += NTP) reinterpret_cast<float4*>(Wx)[i] = reinterpret_cast<const float4*>(pack + kFcC)[i];
  }
  const float* W0 = Wb;
  const float* b0 = W0 + 64 * 32;
  const float* W1 = b0 + 32;
  const float* b1 = W1 + 32 * 32;
  const float* Ws = b1 + 32;
  const int ntiles = (n + PT - 1) / PT;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int j0 = tile * PT;
    __syncthreads();  // previous tile fully consumed (and the weights are in place on the first pass)
    {
      // two threads per row: half 0 / 1 load the first / second 32 input channels of row `row`
      const int row = tid & (PT - 1), half = tid >> 7;
      const int j = j0 + row;
      const int i = j < n ? order[j] : -1;
      const int m = i >= 0 ? p2v[i] : -1;
      if (half == 0) s_pil[row] = m;
      if (STAGE == 0) {
        if (half == 0) {
          float f[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) f[k] = 0.f;
          if (i >= 0) {
            const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
            f[0] = px, f[1] = py, f[2] = pz;
            f[3] = __fsub_rn(px, pmean[3 * m]);
            f[4] = __fsub_rn(py, pmean[3 * m + 1]);
            f[5] = __fsub_rn(pz, pmean[3 * m + 2]);
            const int4 c = reinterpret_cast<const int4*>(coords)[m];  // z, y, x, t
            f[6] = (float)((double)px - ((double)c.z * g.vx + g.x_off));
            f[7] = (float)((double)py - ((double)c.y * g.vy + g.y_off));
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = __fdiv_rn(f[k], g.scale);
            f[8] = __fdiv_rn((float)ptime[i], g.n_frames);
          }
#pragma unroll
          for (int k = 0; k < 9; ++k) NETs[k * LD + row] = f[k];
        }
      } else {
        // half 0: the 32 channels of the previous block output; half 1: the 32 channels of the pillar's pooled vector
        const float4* a = half == 0 ? reinterpret_cast<const float4*>(net_in + (size_t)(j < n ? j : 0) * 32)
                                    : reinterpret_cast<const float4*>(pooled_in + (size_t)(m >= 0 ? m : 0) * 32);
        float* dst = X + half * 32 * LD + row;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = j < n ? a[q] : make_float4(0.f, 0.f, 0.f, 0.f);
          dst[(4 * q + 0) * LD] = v.x, dst[(4 * q + 1) * LD] = v.y, dst[(4 * q + 2) * LD] = v.z, dst[(4 * q + 3) * LD] = v.w;
        }
      }
    }
    __syncthreads();
    if (STAGE == 0) {
      float a64[8][RPT];
      tile_dense<9, 64, false>(NETs, Wx, Wx + 9 * 64, a64, tr, tc);
      tile_store<64>(X, a64, tr, tc);
      __syncthreads();
    }
    float out[4][RPT];
    {
      float net[4][RPT];
      tile_dense<64, 32, true>(X, W0, b0, net, tr, tc);
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int p = 0; p < RPT; ++p) net[o][p] = fmaxf(net[o][p], 0.f);  // fc_1 consumes relu(net)
      tile_store<32>(NETs, net, tr, tc);
    }
    tile_dense<64, 32, false>(X, Ws, nullptr, out, tr, tc);
    __syncthreads();  // NETs complete; everyone is done reading X
    {
      float dx[4][RPT];
      tile_dense<32, 32, false>(NETs, W1, b1, dx, tr, tc);
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int p = 0; p < RPT; ++p) out[o][p] += dx[o][p];
    }
    float* R = X;  // result tile [32][LD]
    if (STAGE == 2) {
      tile_store<32>(X, out, tr, tc);
      __syncthreads();
      float y[4][RPT];
      tile_dense<32, 32, false>(X, Wx, Wx + 32 * 32, y, tr, tc);
      tile_store<32>(NETs, y, tr, tc);  // every thread finished reading NETs before the barrier above
      R = NETs;
    } else {
      tile_store<32>(X, out, tr, tc);
      // the next stage reads the block output row-major from HBM: 8 points x 16 B per thread, a 128 B row per 8 lanes
#pragma unroll
      for (int p = 0; p < RPT; ++p) {
        const int j = j0 + tr * RPT + p;
        if (j < n) *reinterpret_cast<float4*>(net_out + (size_t)j * 32 + 4 * tc) = make_float4(out[0][p], out[1][p], out[2][p], out[3][p]);
      }
    }
    __syncthreads();
    // fused segment max: thread = channel c x one slice of the tile's rows
    {
      const int c = tid & 31, r0 = (tid >> 5) * (PT * 32 / NTP), r1 = r0 + PT * 32 / NTP;
      const float* row = R + c * LD;
      int cur = s_pil[r0], start = r0;
      float v = -INFINITY;
      for (int r = r0; r <= r1; ++r) {
        const int pil = r < r1 ? s_pil[r] : -2;
        if (pil != cur) {
          if (cur >= 0) {
            float* dst = pooled_out + (size_t)cur * 32 + c;
            if (start > r0 && r < r1) *dst = v; else atomic_max_float(dst, v);
          }
          cur = pil, start = r, v = -INFINITY;
        }
        if (r < r1) v = fmaxf(v, row[r]);
      }
    }
  }
}

// =====================================================================================================================
// The same three stages on the 5th-generation tensor cores (tcgen05, fp16-pair operands like csrc/conv_p16.cu):
// a tile = 128 pillar-sorted points = the M of every MMA.  Per tile the CTA
//   * builds the A operand in shared memory: one 128-byte swizzled row per point and 32-channel group, [32 ch h | 32 ch l]
//     (thread (row, half) owns channels 32*half .. +31 of its point for the whole tile),
//   * one thread issues  D[:, 0:O] = a_h.w_h ,  D[:, O:2O] = a_h.w_l + a_l.w_h  against weights that stay resident in shared
//     memory (K-dense rows: two 32-channel K groups per 128-byte row),
//   * the epilogue (TMEM -> registers -> bias / ReLU) writes the next layer's A rows in place:
//        RX = relu(x) -> fc_0 -> relu(net) ;  [x | relu(net)] . [shortcut | fc_1] -> block output   (one accumulator, K = 96)
//     so the hidden layers never leave the SM; only the 32-wide block outputs cross HBM between the stages (the global
//     segment max over a pillar's points separates them), and the segment max itself is the fused scan of the FP32 kernel.
// Two CTAs per SM (96 KB of shared memory, 256 TMEM columns each): one tile's epilogue overlaps the other's MMAs.
// =====================================================================================================================
namespace pfn_tc {
using namespace pcab_tc;

constexpr uint32_t kAX = 0;                 // A rows of the block input x (2 groups x 16 KB); later the FP32 result tile
constexpr uint32_t kAN = 32768;             // A rows of relu(net) / the 9 input features / the fc_c input (16 KB)
constexpr uint32_t kW0 = kAN + 16384;       // fc_0      : 64 rows  x 128 B          ( 8 KB)
constexpr uint32_t kW1 = kW0 + 8192;        // [Ws | W1] : 2 stages x 64 rows x 128 B (16 KB)
constexpr uint32_t kWX = kW1 + 16384;       // stage 0: fc_pos 128 rows x 128 B (16 KB) ; stage 2: fc_c 64 rows x 128 B (8 KB)
constexpr uint32_t kBias = kWX + 16384;     // b0[32] b1[32] bx[64] floats
constexpr uint32_t kBar = kBias + 512;
constexpr uint32_t kSmem = kBar + 64 + 1024;
constexpr int kBlobHalves = (8192 + 16384 + 16384) / 2;  // fp16 elements of one stage's weight blob

struct Scales {
  float inv0, inv1, invx;  // 1 / (power-of-two scale) of fc_0, [Ws | W1], fc_pos or fc_c
};

__device__ __forceinline__ uint32_t row_chunk(uint32_t base, int row, uint32_t chunk) {
  return base + (uint32_t)row * 128u + ((chunk ^ ((uint32_t)row & 7u)) << 4);
}
// eight consecutive channels (one 16-byte chunk of h halves, one of l halves) of row `row`
__device__ __forceinline__ void put8(uint32_t base, int row, uint32_t chunk, const float* v, bool relu) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float a = relu ? fmaxf(v[2 * u], 0.f) : v[2 * u], b = relu ? fmaxf(v[2 * u + 1], 0.f) : v[2 * u + 1];
    p16::split2(a, b, h[u], l[u]);
  }
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_chunk(base, row, chunk)), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_chunk(base, row, 4u + chunk)), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
}

// D[:, 0:2*OUT] (+)= A . B over `ngroups` 32-channel groups; A groups are 16 KB apart from a_base; the weights of K group kg sit
// in stage kg / 2 (2*OUT rows x 128 B) at byte (kg % 2) * 64 of the rows.  kk_max: k-steps of 16 channels per group (1 or 2).
template <int OUT>
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_base, int ngroups, uint32_t w_base, int kg0, int kk_max, bool first) {
  const uint32_t idesc_base = (1u << 4) | ((128u >> 4) << 24);
  const uint32_t idesc2 = idesc_base | ((uint32_t)((2 * OUT) >> 3) << 17), idesc1 = idesc_base | ((uint32_t)(OUT >> 3) << 17);
  const uint64_t desc_hi = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;
  const uint32_t lbo = 1u << 16;
  for (int g = 0; g < ngroups; ++g) {
    const int kg = kg0 + g;
    const uint32_t a16 = lbo | (((a_base + (uint32_t)g * 16384u) & 0x3FFFF) >> 4);
    const uint32_t b16 = lbo | (((w_base + (uint32_t)(kg >> 1) * (uint32_t)(2 * OUT) * 128u) & 0x3FFFF) >> 4);
    for (int kk = 0; kk < kk_max; ++kk) {
      const uint64_t dah = desc_hi | (a16 + 2u * kk), dal = desc_hi | (a16 + 4u + 2u * kk);
      const uint64_t db = desc_hi | (b16 + (uint32_t)(kg & 1) * 4u + 2u * kk);
      umma_f16(tmem_d, dah, db, idesc2, (first && g == 0 && kk == 0) ? 0u : 1u);
      umma_f16(tmem_d + (uint32_t)OUT, dal, db, idesc1, 1u);
    }
  }
}

template <int STAGE>
__global__ void __launch_bounds__(256, 2)
k_pfn_tc(const float* __restrict__ xyz, const int* __restrict__ ptime, const int* __restrict__ order, const int* __restrict__ p2v,
         const int* __restrict__ coords, const float* __restrict__ pmean, const float* __restrict__ net_in,
         const float* __restrict__ pooled_in, const __half* __restrict__ wblob, const float* __restrict__ bias /* b0 b1 bx */, int n,
         PfnGeom g, Scales sc, float* __restrict__ net_out, float* __restrict__ pooled_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sp = smem_raw + (sb - smem_u32(smem_raw));
  float* s_bias = reinterpret_cast<float*>(sp + kBias);
  float* R = reinterpret_cast<float*>(sp + kAX);  // [32][LD] FP32 result tile (aliases the x rows, dead by then)
  __shared__ int s_pil[PT];
  const uint32_t bar = sb + kBar, tmem_slot = sb + kBar + 16;
  const int tid = threadIdx.x, row = tid & 127, half = tid >> 7;
  const int warp = tid >> 5;
  // weights -> shared memory, swizzled like a TMA 128B-swizzle load (16-byte chunk index ^ row % 8)
  {
    const uint4* src = reinterpret_cast<const uint4*>(wblob + (size_t)STAGE * kBlobHalves);
    for (int e = tid; e < (8192 + 16384 + 16384) / 16; e += 256) {
      const int r = e >> 3, c = e & 7;  // rows of 128 B across the three matrices back to back (all 1024-aligned)
      const uint4 v = src[e];
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + kW0 + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4)), "r"(v.x),
                   "r"(v.y), "r"(v.z), "r"(v.w)
                   : "memory");
    }
    for (int e = tid; e < 128; e += 256) s_bias[e] = bias[STAGE * 128 + e];
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t D0 = tmem_base;          // fc_pos (columns 0..127) / fc_0 / fc_c (columns 0..63)
  const uint32_t DB = tmem_base + 192u;   // block output (columns 192..255)
  uint32_t phase = 0;
  auto mma_done = [&]() {  // every thread waits for the MMAs committed by thread 0
    mbar_wait(bar, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };
  auto publish = [&]() {  // A rows written with st.shared -> visible to the tensor core of the issuing thread
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };
  // 16 accumulator columns [col, col+16) of this thread's row: main + correction halves (OUT columns apart), scaled back
  auto drain16 = [&](uint32_t d, uint32_t col, uint32_t out, float inv, float* v) {
    uint32_t vm[16], vc[16];
    tmem_ld16(d + lane_addr + col, vm);
    tmem_ld16(d + lane_addr + out + col, vc);
    tmem_ld_wait16(vm, vc);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (__uint_as_float(vm[j]) + __uint_as_float(vc[j])) * inv;
  };

  const int ntiles = (n + PT - 1) / PT;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int j0 = tile * PT, j = j0 + row;
    const int i = j < n ? order[j] : -1;
    const int m = i >= 0 ? p2v[i] : -1;
    __syncthreads();  // the previous tile's scan is done with s_pil / R
    if (half == 0) s_pil[row] = m;
    float xv[32];
    if (STAGE == 0) {
      if (half == 0) {
        float f[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) f[k] = 0.f;
        if (i >= 0) {
          const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
          f[0] = px, f[1] = py, f[2] = pz;
          f[3] = __fsub_rn(px, pmean[3 * m]);
          f[4] = __fsub_rn(py, pmean[3 * m + 1]);
          f[5] = __fsub_rn(pz, pmean[3 * m + 2]);
          const int4 c = reinterpret_cast<const int4*>(coords)[m];  // z, y, x, t
          f[6] = (float)((double)px - ((double)c.z * g.vx + g.x_off));
          f[7] = (float)((double)py - ((double)c.y * g.vy + g.y_off));
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = __fdiv_rn(f[k], g.scale);
          f[8] = __fdiv_rn((float)ptime[i], g.n_frames);
        }
        put8(sb + kAN, row, 0, f, false);
        put8(sb + kAN, row, 1, f + 8, false);
      }
      publish();
      if (tid == 0) {  // fc_pos: 9 (16) -> 64
        issue_layer<64>(D0, sb + kAN, 1, sb + kWX, 0, 1, true);
        umma_commit(bar);
      }
      mma_done();
      drain16(D0, 32u * half, 64u, sc.invx, xv);
      drain16(D0, 32u * half + 16u, 64u, sc.invx, xv + 16);
#pragma unroll
      for (int k = 0; k < 32; ++k) xv[k] += s_bias[64 + 32 * half + k];
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {
      // half 0: the 32 channels of the previous block output; half 1: the 32 channels of the pillar's pooled vector
      const float4* a = half == 0 ? reinterpret_cast<const float4*>(net_in + (size_t)(j < n ? j : 0) * 32)
                                  : reinterpret_cast<const float4*>(pooled_in + (size_t)(m >= 0 ? m : 0) * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = j < n ? a[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        xv[4 * q] = v.x, xv[4 * q + 1] = v.y, xv[4 * q + 2] = v.z, xv[4 * q + 3] = v.w;
      }
    }
    // ---- fc_0 on relu(x)
#pragma unroll
    for (int q = 0; q < 4; ++q) put8(sb + kAX + 16384u * half, row, q, xv + 8 * q, true);
    publish();
    if (tid == 0) {
      issue_layer<32>(D0, sb + kAX, 2, sb + kW0, 0, 2, true);
      umma_commit(bar);
    }
    mma_done();
    {
      float nv[16];
      drain16(D0, 16u * half, 32u, sc.inv0, nv);
#pragma unroll
      for (int k = 0; k < 16; ++k) nv[k] += s_bias[16 * half + k];
      put8(sb + kAN, row, 2 * half, nv, true);  // fc_1 consumes relu(net)
      put8(sb + kAN, row, 2 * half + 1, nv + 8, true);
    }
    // the x rows themselves (the shortcut reads x, not relu(x)); fc_0 has finished reading this buffer
#pragma unroll
    for (int q = 0; q < 4; ++q) put8(sb + kAX + 16384u * half, row, q, xv + 8 * q, false);
    publish();
    if (tid == 0) {  // block output = shortcut(x) + fc_1(relu(net)): one accumulator over K = 64 + 32
      issue_layer<32>(DB, sb + kAX, 2, sb + kW1, 0, 2, true);
      issue_layer<32>(DB, sb + kAN, 1, sb + kW1, 2, 2, false);
      umma_commit(bar);
    }
    mma_done();
    float ov[16];
    drain16(DB, 16u * half, 32u, sc.inv1, ov);
#pragma unroll
    for (int k = 0; k < 16; ++k) ov[k] += s_bias[32 + 16 * half + k];
    if (STAGE == 2) {
      put8(sb + kAN, row, 2 * half, ov, false);
      put8(sb + kAN, row, 2 * half + 1, ov + 8, false);
      publish();
      if (tid == 0) {  // fc_c: 32 -> 32
        issue_layer<32>(D0, sb + kAN, 1, sb + kWX, 0, 2, true);
        umma_commit(bar);
      }
      mma_done();
      drain16(D0, 16u * half, 32u, sc.invx, ov);
#pragma unroll
      for (int k = 0; k < 16; ++k) ov[k] += s_bias[64 + 16 * half + k];
    } else if (j < n) {
      float4* dst = reinterpret_cast<float4*>(net_out + (size_t)j * 32 + 16 * half);
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[q] = make_float4(ov[4 * q], ov[4 * q + 1], ov[4 * q + 2], ov[4 * q + 3]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // every MMA that read the x rows has completed (mma_done above): the result tile may overwrite them
#pragma unroll
    for (int k = 0; k < 16; ++k) R[(16 * half + k) * LD + row] = ov[k];
    __syncthreads();
    // fused segment max: thread = channel c x one slice of the tile's rows (same scan as the FP32 kernel)
    {
      const int c = tid & 31, r0 = (tid >> 5) * (PT * 32 / 256), r1 = r0 + PT * 32 / 256;
      const float* rr = R + c * LD;
      int cur = s_pil[r0], start = r0;
      float v = -INFINITY;
      for (int r = r0; r <= r1; ++r) {
        const int pil = r < r1 ? s_pil[r] : -2;
        if (pil != cur) {
          if (cur >= 0) {
            float* dst = pooled_out + (size_t)cur * 32 + c;
            if (start > r0 && r < r1) *dst = v; else atomic_max_float(dst, v);
          }
          cur = pil, start = r, v = -INFINITY;
        }
        if (r < r1) v = fmaxf(v, rr[r]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

}  // namespace pfn_tc

__global__ void k_fill_neg_inf(float4* __restrict__ a, long long n4) {
  long long stride = (long long)gridDim.x * blockDim.x;
  const float4 v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) a[i] = v;
}

// pillar features -> canvas cells (models/pillar_encoder.py:158-172); a pillar without points keeps 0 like torch_scatter
template <bool P16>
__global__ void k_canvas_scatter(float4* __restrict__ feats, const int* __restrict__ cell_of_pillar, int m,
                                 float* __restrict__ canvas) {
  long long total = (long long)m * 8, stride = (long long)gridDim.x * blockDim.x;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += stride) {
    const int p = (int)(e >> 3), q = (int)(e & 7);
    float4 v = feats[e];
    if (v.x == -INFINITY) {
      v = make_float4(0.f, 0.f, 0.f, 0.f);
      feats[e] = v;
    }
    p16::st4<P16>(canvas, (size_t)cell_of_pillar[p], 32, 4 * q, v);
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

// The tensor-core variant: same contract; `w_tc` = three stage blobs of fp16 [fc_0 64 x 64 | [Ws|W1] 2 x 64 x 64 | fc_pos 128 x 64
// or fc_c 64 x 64 (+ padding to 128 rows)] (tc_pack.pack_pfn_tc), `bias_tc` = per stage b0[32] b1[32] bx[64], `scales_inv9` (HOST)
// = per stage 1 / scale of (fc_0, [Ws|W1], extra).
extern "C" int pcab_pillar_encode_tc(const float* xyz, const int* point_time, const int* order, const int* p2v, const int* coords_zyxt,
                                     const int* pillar_cell, const float* pillar_mean, const void* w_tc, const float* bias_tc,
                                     const float* scales_inv9, int n_points, int n_pillars, const float* range6,
                                     const float* voxel_size3, int n_sweeps, float* pillar_feats, float* canvas_nhwc, int canvas_fmt,
                                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PCAB_REQUIRE(workspace_bytes >= pcab_pillar_encode_workspace(n_points, n_pillars), "workspace too small");
  PCAB_REQUIRE(((uintptr_t)w_tc & 15) == 0, "weight blob must be 16 B aligned");
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
  static PcabSmemOnce once0, once1, once2;
  PCAB_CUDA(pcab_set_max_smem(pfn_tc::k_pfn_tc<0>, (int)pfn_tc::kSmem, once0));
  PCAB_CUDA(pcab_set_max_smem(pfn_tc::k_pfn_tc<1>, (int)pfn_tc::kSmem, once1));
  PCAB_CUDA(pcab_set_max_smem(pfn_tc::k_pfn_tc<2>, (int)pfn_tc::kSmem, once2));
  const int ntiles = (n_points + PT - 1) / PT;
  const int gp = ntiles < 2 * pcab_sm_count() ? ntiles : 2 * pcab_sm_count();
  const long long pool4 = (long long)n_pillars * 8;
  const int gf = grid_for(pool4, 256, 8);
  const __half* wb = (const __half*)w_tc;
  pfn_tc::Scales s0 = {scales_inv9[0], scales_inv9[1], scales_inv9[2]}, s1 = {scales_inv9[3], scales_inv9[4], scales_inv9[5]},
                 s2 = {scales_inv9[6], scales_inv9[7], scales_inv9[8]};
  k_fill_neg_inf<<<gf, 256, 0, stream>>>((float4*)pillar_feats, pool4);
  pfn_tc::k_pfn_tc<0><<<gp, 256, pfn_tc::kSmem, stream>>>(xyz, point_time, order, p2v, coords_zyxt, pillar_mean, nullptr, nullptr, wb, bias_tc,
                                                        n_points, g, s0, net_a, pillar_feats);
  k_fill_neg_inf<<<gf, 256, 0, stream>>>((float4*)pooled, pool4);
  pfn_tc::k_pfn_tc<1><<<gp, 256, pfn_tc::kSmem, stream>>>(xyz, point_time, order, p2v, coords_zyxt, pillar_mean, net_a, pillar_feats, wb, bias_tc,
                                                        n_points, g, s1, net_b, pooled);
  k_fill_neg_inf<<<gf, 256, 0, stream>>>((float4*)pillar_feats, pool4);
  pfn_tc::k_pfn_tc<2><<<gp, 256, pfn_tc::kSmem, stream>>>(xyz, point_time, order, p2v, coords_zyxt, pillar_mean, net_b, pooled, wb, bias_tc,
                                                        n_points, g, s2, nullptr, pillar_feats);
  if (canvas_fmt)
    k_canvas_scatter<true><<<gf, 256, 0, stream>>>((float4*)pillar_feats, pillar_cell, n_pillars, canvas_nhwc);
  else
    k_canvas_scatter<false><<<gf, 256, 0, stream>>>((float4*)pillar_feats, pillar_cell, n_pillars, canvas_nhwc);
  PCAB_CHECK_LAUNCH("pcab_pillar_encode_tc");
  return PCAB_OK;
}

extern "C" int pcab_pillar_encode(const float* xyz, const int* point_time, const int* order, const int* p2v,
                                  const int* pstart, const int* coords_zyxt, const int* pillar_cell,
                                  const float* pillar_mean, const float* weight_pack, int n_points, int n_pillars,
                                  const float* range6, const float* voxel_size3, int n_sweeps, float* pillar_feats,
                                  float* canvas_nhwc, int canvas_fmt, void* workspace, size_t workspace_bytes,
                                  cudaStream_t stream) {
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
  (void)pstart;  // the segment boundaries are recovered from the sorted pillar ids inside the tiles
  const size_t smem = (size_t)(64 * LD + 32 * LD + kBlkSize + 1056) * sizeof(float);
  static PcabSmemOnce once0, once1, once2;
  PCAB_CUDA(pcab_set_max_smem(k_pfn_tile<0>, (int)smem, once0));
  PCAB_CUDA(pcab_set_max_smem(k_pfn_tile<1>, (int)smem, once1));
  PCAB_CUDA(pcab_set_max_smem(k_pfn_tile<2>, (int)smem, once2));
  const int ntiles = (n_points + PT - 1) / PT;
  const int gp = ntiles < 2 * pcab_sm_count() ? ntiles : 2 * pcab_sm_count();  // persistent: two CTAs (16 warps) per SM
  const long long pool4 = (long long)n_pillars * 8;
  const int gf = grid_for(pool4, 256, 8);
  // pooled vectors ping-pong between `pillar_feats` and the scratch array (each is re-filled with -inf before reuse)
  k_fill_neg_inf<<<gf, 256, 0, stream>>>((float4*)pillar_feats, pool4);
  k_pfn_tile<0><<<gp, NTP, smem, stream>>>(xyz, point_time, order, p2v, coords_zyxt, pillar_mean, nullptr, nullptr,
                                           weight_pack, n_points, g, net_a, pillar_feats);
  k_fill_neg_inf<<<gf, 256, 0, stream>>>((float4*)pooled, pool4);
  k_pfn_tile<1><<<gp, NTP, smem, stream>>>(xyz, point_time, order, p2v, coords_zyxt, pillar_mean, net_a, pillar_feats,
                                           weight_pack, n_points, g, net_b, pooled);
  k_fill_neg_inf<<<gf, 256, 0, stream>>>((float4*)pillar_feats, pool4);
  k_pfn_tile<2><<<gp, NTP, smem, stream>>>(xyz, point_time, order, p2v, coords_zyxt, pillar_mean, net_b, pooled,
                                           weight_pack, n_points, g, nullptr, pillar_feats);
  if (canvas_fmt)
    k_canvas_scatter<true><<<gf, 256, 0, stream>>>((float4*)pillar_feats, pillar_cell, n_pillars, canvas_nhwc);
  else
    k_canvas_scatter<false><<<gf, 256, 0, stream>>>((float4*)pillar_feats, pillar_cell, n_pillars, canvas_nhwc);
  PCAB_CHECK_LAUNCH("pcab_pillar_encode");
  return PCAB_OK;
}
