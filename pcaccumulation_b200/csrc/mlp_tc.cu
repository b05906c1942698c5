// Per-point MLPs on the 5th-generation tensor cores (tcgen05.mma kind::tf32, FP32 accumulators in TMEM, 3xTF32 error
// compensation as in conv_tc.cu): the STPN point head of models/stpn.py:91-103 (positional encoding, bilinear pickup,
// final_proj, mos_seg and offset_head) and, further down, the TubeNet embeddings of models/tpointnet.py:171-262
// (k_embed_tc: MLP + max over the rows of each segment).
//
// A CTA (one per SM, persistent) owns tiles of 128 foreground points = the 128 rows (M) of every MMA.  The activations
// of a tile live in shared memory as the K-major, 128B-swizzled A operand: 4 "atoms" of [128 rows x 32 channels] per
// plane, one plane with the FP32 values (the tensor core reads their upper 19 bits = a_hi) and one with the residuals
// a_lo.  A layer y = act(W x + b) is
//     D[:, 0:2N] = a_hi . [w_hi | w_lo]^T        (one N = 2*out MMA per 8 input channels)
//     D[:, N:2N] += a_lo . w_hi^T                 (one N = out MMA)
// and y = D[:, 0:N] + D[:, N:2N] (+ bias, BN, ReLU) is formed by the epilogue warps, which write it back IN PLACE as the
// A operand of the next layer (the MMAs of the producing layer have completed by then).  Weights ([w_hi; w_lo] rows,
// K-major, pre-split on the host) stream through a 3-stage TMA ring in 32-input-channel chunks.  mos_seg and offset_head
// share their input, use the two halves of TMEM and are drained by different warps, so their 128 -> 2 projections are
// register dot products in the epilogue and their hidden layers never exist in memory.
//
// Warp roles: 0-7 gather / positional encoding / epilogues (thread = one point x one half of the channels; TMEM lane
// quadrant = warp % 4), 8 = weight producer (TMA), 9 = TMEM allocator + MMA issuer.
#include <cuda.h>
#include <cstring>
#include <type_traits>
#include "common.cuh"
#include "pair16.cuh"
#include "mlp.cuh"
#include "pcab200.h"
#include "tc_common.cuh"

namespace {

using namespace pcab_tc;
using namespace mlp::stpn_pack;

constexpr int kNT = 320;
constexpr uint32_t kAtom = 128 * 128;     // [128 rows][32 ch] fp32 = 16 KB
constexpr uint32_t kPlane = 4 * kAtom;    // 128 channels
constexpr uint32_t kWStage = 256 * 128;   // [w_hi 128 rows | w_lo 128 rows] x 32 input channels
constexpr int kStages = 3;
constexpr int kChunksPerTile = 13;        // pe2: 1, final_proj / mos0 / off0: 4 each
constexpr size_t kSmemBytes = 1024 + 2 * (size_t)kPlane + kStages * (size_t)kWStage + 128 + 3 * 128 * 4;


struct HeadArgs {
  const float* mos_feats;
  int fmt;  // activation format of mos_feats: 0 = float32 NHWC, 1 = P16 (pair16.cuh)
  int H, W;
  const float* tp;
  const int* pbatch;
  const int* fg_idx;
  int n_fg;
  float x_abs, y_abs;
  float* mos_out;
  float* off_out;
  int n_tiles;
  long long* stats;  // debug: per-CTA phase cycle counters (null = off)
};

// Small per-layer vectors (biases, BN affine, pe0, the two 128 -> 2 projections) travel as a kernel parameter: parameters
// live in the constant bank, so with the fully unrolled epilogues they become c[0x0][..] operands of the FMAs - no loads at
// all.  (The CTA's shared memory is full and with a 227 KB carve-out there is no L1 left to cache global loads.)
struct HeadConsts {
  float pe0w[96], pe0b[32], pe2b[64], fpb[128];
  float hb[2][128], hs[2][128], ht[2][128], h3w[2][256], h3b[2][2];  // [0] = mos_seg, [1] = offset_head
};

__device__ __forceinline__ void mbar_arrive_cnt(uint32_t bar, uint32_t cnt) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cnt) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
// channels c..c+3 (c % 4 == 0) of row r, as the A operand expects them: value plane and residual plane
__device__ __forceinline__ void put4(uint32_t a_hi, uint32_t a_lo, int r, int c, float x, float y, float z, float w) {
  const uint32_t off = (uint32_t)(c >> 5) * kAtom + (uint32_t)r * 128u + (uint32_t)((((c & 31) >> 2) ^ (r & 7)) << 4);
  sts128(a_hi + off, x, y, z, w);
  sts128(a_lo + off, split_lo(x), split_lo(y), split_lo(z), split_lo(w));
}

__global__ void __launch_bounds__(kNT, 1)
k_stpn_head_tc(const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w, const HeadArgs a,
               const __grid_constant__ HeadConsts k) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_hi = sbase, a_lo = sbase + kPlane, w0 = sbase + 2 * kPlane;
  const uint32_t bars = w0 + kStages * kWStage;
  const uint32_t bar_w_full = bars, bar_w_free = bars + 24, bar_a_ready = bars + 48, bar_acc_full = bars + 56,
                 bar_acc_empty = bars + 72, tmem_slot = bars + 88;
  const uint32_t s_px = bars + 128, s_py = s_px + 512, s_pb = s_py + 512;  // per-point x, y, batch of the tile being gathered
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(bar_w_full + 8 * i, 1), mbar_init(bar_w_free + 8 * i, 1);
    mbar_init(bar_a_ready, 256);
    for (int i = 0; i < 2; ++i) mbar_init(bar_acc_full + 8 * i, 1), mbar_init(bar_acc_empty + 8 * i, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 8 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w1)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  if (warp == 8) {
    // ===================== weight producer =====================
    if (lane == 0) {
      int wg = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        for (int c = 0; c < kChunksPerTile; ++c, ++wg) {
          const int ws = wg % kStages;
          if (wg >= kStages) mbar_wait(bar_w_free + 8 * ws, ((wg / kStages) - 1) & 1);
          const uint32_t dst = w0 + ws * kWStage, bar = bar_w_full + 8 * ws;
          if (c == 0) {
            mbar_expect_tx(bar, 128u * 128u);
            tma_load_2d(&map_w1, dst, bar, 0, 0);  // pe2: [w_hi 64 rows; w_lo 64 rows] x 32
          } else {
            mbar_expect_tx(bar, kWStage);
            tma_load_2d(&map_w, dst, bar, ((c - 1) & 3) * 32, ((c - 1) >> 2) * 256);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24);
    const uint64_t desc_hi = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;  // SBO 1024 B, version 1, SWIZZLE_128B
    const uint32_t lbo = 1u << 16;
    const uint32_t ah_base = lbo | ((a_hi & 0x3FFFF) >> 4), al_base = lbo | ((a_lo & 0x3FFFF) >> 4);
    const uint32_t b_base = lbo | ((w0 & 0x3FFFF) >> 4);
    int wg = 0, job = 0, ar = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      for (int j = 0; j < 4; ++j, ++job) {
        const int buf = job & 1;
        const int n = j == 0 ? 64 : 128, nchunks = j == 0 ? 1 : 4;
        if (j < 3) {
          mbar_wait(bar_a_ready, ar & 1);
          ++ar;
        }
        if (job >= 2) mbar_wait(bar_acc_empty + 8 * buf, ((job >> 1) - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t tmem_d = tmem_base + (uint32_t)(buf * 256);
          const uint32_t idesc2 = idesc_base | ((uint32_t)((2 * n) >> 3) << 17), idesc1 = idesc_base | ((uint32_t)(n >> 3) << 17);
          for (int c = 0; c < nchunks; ++c) {
            const int ws = (wg + c) % kStages;
            mbar_wait(bar_w_full + 8 * ws, ((wg + c) / kStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ah = ah_base + (uint32_t)c * (kAtom >> 4), al = al_base + (uint32_t)c * (kAtom >> 4);
            const uint32_t b16 = b_base + (uint32_t)ws * (kWStage >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t dah = desc_hi | (ah + 2u * kk), dal = desc_hi | (al + 2u * kk), db = desc_hi | (b16 + 2u * kk);
              umma_tf32(tmem_d, dah, db, idesc2, (c | kk) ? 1u : 0u);  // [a_hi*w_hi | a_hi*w_lo]
              umma_tf32(tmem_d + (uint32_t)n, dal, db, idesc1, 1u);    // += a_lo*w_hi into the second half
            }
            umma_commit(bar_w_free + 8 * ws);
          }
          umma_commit(bar_acc_full + 8 * buf);
        }
        __syncwarp();
        wg += nchunks;
      }
    }
  } else {
    // ===================== gather, positional encoding, epilogues =====================
    const int r = (warp & 3) * 32 + lane, h = warp >> 2;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const bool st_on = a.stats != nullptr && (threadIdx.x == 0 || threadIdx.x == 128);
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = st_on ? clock64() : 0;
#define PCAB_PHASE(k)                      \
  if (st_on) {                             \
    const long long tnow = clock64();      \
    ph[k] += tnow - tprev, tprev = tnow;   \
  }
    // Per-tile inputs are fetched one tile ahead: the point metadata goes through a small shared table, the bilinear
    // pickup (work item = (point, 16-byte chunk): the 16 lanes of a point read whole 256 B pixel rows of the four taps,
    // coalesced - a thread-per-point gather would send 16 B requests to 32 different lines per instruction and there is no
    // L1 left to merge them) is combined into registers while the head MMAs of the previous tile run, and is written into
    // the A planes once those MMAs have released them.
    int i_cur = 0;
    float px = 0.f, py = 0.f, pz = 0.f;
    bool valid = false;
    float4 g[8];
    auto prefetch = [&](int tile) {
      const int base = tile * 128;
      valid = base + r < a.n_fg;
      i_cur = a.fg_idx[valid ? base + r : base];  // padding rows recompute the tile's first point (never stored)
      px = a.tp[3 * i_cur], py = a.tp[3 * i_cur + 1], pz = a.tp[3 * i_cur + 2];
      if (h == 0) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_px + 4u * r), "f"(px) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_py + 4u * r), "f"(py) : "memory");
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(s_pb + 4u * r), "r"(a.pbatch[i_cur]) : "memory");
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int ctid = (int)threadIdx.x;  // 0..255
      if (a.fmt) {
        // P16 feature map: work item = (point, 8-channel block): one 16-byte load of the h halves and one of the l halves per
        // tap (the 8 lanes of a point still cover whole 256-byte pixels), g[2*it], g[2*it+1] = channels 8*q8 .. 8*q8+7
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int e = it * 256 + ctid, p = e >> 3, q8 = e & 7;
          float qx, qy;
          int qb;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(qx) : "r"(s_px + 4u * p));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(qy) : "r"(s_py + 4u * p));
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(qb) : "r"(s_pb + 4u * p));
          const mlp::Bilinear bl = mlp::bilinear_border(qx, qy, a.x_abs, a.y_abs, a.H, a.W);
          const char* base = reinterpret_cast<const char*>(a.mos_feats) + (size_t)qb * a.H * a.W * 256 + (q8 >> 2) * 128 + (q8 & 3) * 16;
          const int offs[4] = {bl.o00, bl.o01, bl.o10, bl.o11};
          const float wts[4] = {bl.w00, bl.w01, bl.w10, bl.w11};
          uint4 hh[4], ll[4];
#pragma unroll
          for (int tp_ = 0; tp_ < 4; ++tp_) {
            hh[tp_] = *reinterpret_cast<const uint4*>(base + (size_t)offs[tp_] * 256);
            ll[tp_] = *reinterpret_cast<const uint4*>(base + (size_t)offs[tp_] * 256 + 64);
          }
          float r8[8];
#pragma unroll
          for (int tp_ = 0; tp_ < 4; ++tp_) {
            const uint32_t hw_[4] = {hh[tp_].x, hh[tp_].y, hh[tp_].z, hh[tp_].w}, lw_[4] = {ll[tp_].x, ll[tp_].y, ll[tp_].z, ll[tp_].w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 v = p16::join2(hw_[u], lw_[u]);
              r8[2 * u] = tp_ == 0 ? v.x * wts[0] : fmaf(v.x, wts[tp_], r8[2 * u]);
              r8[2 * u + 1] = tp_ == 0 ? v.y * wts[0] : fmaf(v.y, wts[tp_], r8[2 * u + 1]);
            }
          }
          g[2 * it] = make_float4(r8[0], r8[1], r8[2], r8[3]);
          g[2 * it + 1] = make_float4(r8[4], r8[5], r8[6], r8[7]);
        }
        return;
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int e = it * 256 + ctid, p = e >> 4, q = e & 15;
        float qx, qy;
        int qb;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(qx) : "r"(s_px + 4u * p));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(qy) : "r"(s_py + 4u * p));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(qb) : "r"(s_pb + 4u * p));
        const mlp::Bilinear bl = mlp::bilinear_border(qx, qy, a.x_abs, a.y_abs, a.H, a.W);
        const float4* b4 = reinterpret_cast<const float4*>(a.mos_feats + (size_t)qb * a.H * a.W * 64) + q;
        const float4 t00 = b4[(size_t)bl.o00 * 16], t01 = b4[(size_t)bl.o01 * 16], t10 = b4[(size_t)bl.o10 * 16], t11 = b4[(size_t)bl.o11 * 16];
        g[it].x = fmaf(t11.x, bl.w11, fmaf(t10.x, bl.w10, fmaf(t01.x, bl.w01, t00.x * bl.w00)));
        g[it].y = fmaf(t11.y, bl.w11, fmaf(t10.y, bl.w10, fmaf(t01.y, bl.w01, t00.y * bl.w00)));
        g[it].z = fmaf(t11.z, bl.w11, fmaf(t10.z, bl.w10, fmaf(t01.z, bl.w01, t00.z * bl.w00)));
        g[it].w = fmaf(t11.w, bl.w11, fmaf(t10.w, bl.w10, fmaf(t01.w, bl.w01, t00.w * bl.w00)));
      }
    };
    if ((int)blockIdx.x < a.n_tiles) prefetch(blockIdx.x);
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const int i = i_cur;
      const bool valid_cur = valid;
      // ---- gathered motion features of this tile -> A channels 64..127
      {
        const int ctid = (int)threadIdx.x;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          // float32 map: item (point e >> 4, channels 4 * (e & 15)); P16 map: g[2j], g[2j+1] = item (point, channels 8 * q8 ..)
          const int e = a.fmt ? (it >> 1) * 256 + ctid : it * 256 + ctid;
          const int p = a.fmt ? e >> 3 : e >> 4, c = a.fmt ? 8 * (e & 7) + 4 * (it & 1) : 4 * (e & 15);
          put4(a_hi, a_lo, p, 64 + c, g[it].x, g[it].y, g[it].z, g[it].w);
        }
      }
      // ---- positional encoding layer 0 (3 -> 32, ReLU) on the CUDA cores: this thread's 16 hidden channels -> A channels 16h ..
      {
        const float p0 = px / a.x_abs, p1 = py / a.x_abs, p2 = pz / a.x_abs;  // all three by the x scale (models/stpn.py:94)
        auto pe0 = [&](auto hc) {
          constexpr int HH = decltype(hc)::value;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int c = HH * 16 + 4 * q + u;
              float t = fmaf(p0, k.pe0w[c], 0.f);
              t = fmaf(p1, k.pe0w[32 + c], t);
              t = fmaf(p2, k.pe0w[64 + c], t);
              v[u] = fmaxf(t + k.pe0b[c], 0.f);
            }
            put4(a_hi, a_lo, r, HH * 16 + 4 * q, v[0], v[1], v[2], v[3]);
          }
        };
        if (h == 0) pe0(std::integral_constant<int, 0>{}); else pe0(std::integral_constant<int, 1>{});
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(bar_a_ready);
      PCAB_PHASE(0)

      // ---- epilogue of pe2 (32 -> 64, ReLU): this thread's 32 outputs -> A channels 32h ..   (job 0: buffer 0, phase 0)
      mbar_wait(bar_acc_full + 0, 0);
      PCAB_PHASE(1)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      {
        const uint32_t tm = tmem_base + lane_addr;
        auto epi = [&](auto hc) {
          constexpr int HH = decltype(hc)::value;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            uint32_t vm[16], vc[16];
            const int c0 = HH * 32 + b * 16;
            tmem_ld16(tm + (uint32_t)c0, vm);
            tmem_ld16(tm + (uint32_t)(64 + c0), vc);
            tmem_ld_wait16(vm, vc);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float v[4];
#pragma unroll
              for (int u = 0; u < 4; ++u)
                v[u] = fmaxf((__uint_as_float(vm[4 * q + u]) + __uint_as_float(vc[4 * q + u])) + k.pe2b[c0 + 4 * q + u], 0.f);
              put4(a_hi, a_lo, r, c0 + 4 * q, v[0], v[1], v[2], v[3]);
            }
          }
        };
        if (h == 0) epi(std::integral_constant<int, 0>{}); else epi(std::integral_constant<int, 1>{});
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 0);
      mbar_arrive(bar_a_ready);
      PCAB_PHASE(2)

      // ---- epilogue of final_proj (128 -> 128, ReLU): this thread's 64 outputs -> A channels 64h ..  (job 1: buffer 1, phase 0)
      mbar_wait(bar_acc_full + 8, 0);
      PCAB_PHASE(3)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      {
        const uint32_t tm = tmem_base + lane_addr + 256u;
        auto epi = [&](auto hc) {
          constexpr int HH = decltype(hc)::value;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            uint32_t vm[16], vc[16];
            const int c0 = HH * 64 + b * 16;
            tmem_ld16(tm + (uint32_t)c0, vm);
            tmem_ld16(tm + (uint32_t)(128 + c0), vc);
            tmem_ld_wait16(vm, vc);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float v[4];
#pragma unroll
              for (int u = 0; u < 4; ++u)
                v[u] = fmaxf((__uint_as_float(vm[4 * q + u]) + __uint_as_float(vc[4 * q + u])) + k.fpb[c0 + 4 * q + u], 0.f);
              put4(a_hi, a_lo, r, c0 + 4 * q, v[0], v[1], v[2], v[3]);
            }
          }
        };
        if (h == 0) epi(std::integral_constant<int, 0>{}); else epi(std::integral_constant<int, 1>{});
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8);
      mbar_arrive(bar_a_ready);
      PCAB_PHASE(4)
      if (tile + (int)gridDim.x < a.n_tiles) prefetch(tile + gridDim.x);  // flies while the head MMAs run

      // ---- heads: warps 0-3 drain mos0 (job 2: buffer 0, phase 1), warps 4-7 drain off0 (job 3: buffer 1, phase 1).
      // hidden = ReLU(BN(W x + b)); out = W3 hidden + b3 as a register dot product over the 128 hidden channels.
      {
        mbar_wait(bar_acc_full + 8 * h, 1);
        PCAB_PHASE(5)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tm = tmem_base + lane_addr + (uint32_t)(h * 256);
        float o0 = 0.f, o1 = 0.f;
        auto epi = [&](auto hc) {
          constexpr int HH = decltype(hc)::value;
#pragma unroll
          for (int b = 0; b < 8; ++b) {
            uint32_t vm[16], vc[16];
            const int c0 = b * 16;
            tmem_ld16(tm + (uint32_t)c0, vm);
            tmem_ld16(tm + (uint32_t)(128 + c0), vc);
            tmem_ld_wait16(vm, vc);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const int c = c0 + u;
              const float g = fmaxf(fmaf((__uint_as_float(vm[u]) + __uint_as_float(vc[u])) + k.hb[HH][c], k.hs[HH][c], k.ht[HH][c]), 0.f);
              o0 = fmaf(g, k.h3w[HH][2 * c], o0), o1 = fmaf(g, k.h3w[HH][2 * c + 1], o1);
            }
          }
          o0 += k.h3b[HH][0], o1 += k.h3b[HH][1];
        };
        if (h == 0) epi(std::integral_constant<int, 0>{}); else epi(std::integral_constant<int, 1>{});
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cnt(bar_acc_empty + 8 * h, 2);
        if (valid_cur) {
          if (h == 0) {
            *reinterpret_cast<float2*>(a.mos_out + 2 * (size_t)i) = make_float2(o0, o1);
          } else {
            if (isnan(o0) || isinf(o0)) o0 = 0.f;  // safe_guard_offset (models/stpn.py:61-65)
            if (isnan(o1) || isinf(o1)) o1 = 0.f;
            *reinterpret_cast<float2*>(a.off_out + 2 * (size_t)i) =
                make_float2(fminf(fmaxf(o0, -20.f), 20.f), fminf(fmaxf(o1, -20.f), 20.f));
          }
        }
        // the next tile overwrites the A planes: off0's MMAs (job 3) must have finished reading them
        PCAB_PHASE(6)
        if (h == 0) mbar_wait(bar_acc_full + 8, 1);
        PCAB_PHASE(7)
      }
    }
    if (st_on) {
      long long* sp = a.stats + (blockIdx.x * 2 + h) * 8;
      for (int k = 0; k < 8; ++k) sp[k] += ph[k];
    }
#undef PCAB_PHASE
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 9) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// TubeNet embeddings (models/tpointnet.py:171-205, 240-262): a 2- or 3-layer point MLP (ReLU between layers, none after
// the last) followed by a max over the rows of each segment (instance, or instance x frame).  Same machinery as the head
// above: 128 rows per tile, activations as in-place A planes, weights through the TMA ring, two TMEM buffers.  Rows arrive
// sorted by segment, so the last epilogue reduces the runs inside a warp with a segmented shuffle scan and issues one
// atomic max per (run, channel); the destination is pre-filled with -inf.
// ---------------------------------------------------------------------------------------------------------------------
struct EmbedArgs {
  const float* feat;   // [n_src][K0] input rows
  const int* src_idx;  // [n] row of `feat` for each point, or null = identity
  const int* seg;      // [n] segment of each point (ascending)
  int n;
  float* out;          // [n_seg][128]
  int n_tiles;
  long long* stats;    // debug: per-CTA phase cycle counters (null = off)
};
struct EmbedConsts {
  float b[3][128];
};

template <int K0, int N1, int N2>
__global__ void __launch_bounds__(kNT, 1)
k_embed_tc(const __grid_constant__ CUtensorMap map_w0, const __grid_constant__ CUtensorMap map_w1,
           const __grid_constant__ CUtensorMap map_w2, const EmbedArgs a, const __grid_constant__ EmbedConsts k) {
  constexpr int NL = N2 ? 3 : 2;
  constexpr int KL[3] = {K0, N1, N2};
  constexpr int NO[3] = {N1, N2 ? N2 : 128, 128};
  constexpr int QPR = K0 / 4;            // 16-byte chunks per input row
  constexpr int ITEMS = 128 * QPR / 256;  // (row, chunk) items per thread
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_hi = sbase, a_lo = sbase + kPlane, w0 = sbase + 2 * kPlane;
  const uint32_t bars = w0 + kStages * kWStage;
  const uint32_t bar_w_full = bars, bar_w_free = bars + 24, bar_a_ready = bars + 48, bar_acc_full = bars + 56,
                 bar_acc_empty = bars + 72, tmem_slot = bars + 88;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(bar_w_full + 8 * i, 1), mbar_init(bar_w_free + 8 * i, 1);
    mbar_init(bar_a_ready, 256);
    for (int i = 0; i < 2; ++i) mbar_init(bar_acc_full + 8 * i, 1), mbar_init(bar_acc_empty + 8 * i, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  if (warp == 8) {
    // ===================== weight producer =====================
    if (lane == 0) {
      int wg = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
#pragma unroll
        for (int l = 0; l < NL; ++l) {
          const CUtensorMap* m = l == 0 ? &map_w0 : (l == 1 ? &map_w1 : &map_w2);
          for (int c = 0; c < KL[l] / 32; ++c, ++wg) {
            const int ws = wg % kStages;
            if (wg >= kStages) mbar_wait(bar_w_free + 8 * ws, ((wg / kStages) - 1) & 1);
            mbar_expect_tx(bar_w_full + 8 * ws, 2u * NO[l] * 128u);
            tma_load_2d(m, w0 + ws * kWStage, bar_w_full + 8 * ws, c * 32, 0);  // [w_hi NO rows; w_lo NO rows] x 32
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24);
    const uint64_t desc_hi = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;
    const uint32_t lbo = 1u << 16;
    const uint32_t ah_base = lbo | ((a_hi & 0x3FFFF) >> 4), al_base = lbo | ((a_lo & 0x3FFFF) >> 4);
    const uint32_t b_base = lbo | ((w0 & 0x3FFFF) >> 4);
    int wg = 0, job = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
#pragma unroll
      for (int l = 0; l < NL; ++l, ++job) {
        const int buf = job & 1, n = NO[l], nchunks = KL[l] / 32;
        mbar_wait(bar_a_ready, job & 1);
        if (job >= 2) mbar_wait(bar_acc_empty + 8 * buf, ((job >> 1) - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t tmem_d = tmem_base + (uint32_t)(buf * 256);
          const uint32_t idesc2 = idesc_base | ((uint32_t)((2 * n) >> 3) << 17), idesc1 = idesc_base | ((uint32_t)(n >> 3) << 17);
          for (int c = 0; c < nchunks; ++c) {
            const int ws = (wg + c) % kStages;
            mbar_wait(bar_w_full + 8 * ws, ((wg + c) / kStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ah = ah_base + (uint32_t)c * (kAtom >> 4), al = al_base + (uint32_t)c * (kAtom >> 4);
            const uint32_t b16 = b_base + (uint32_t)ws * (kWStage >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t dah = desc_hi | (ah + 2u * kk), dal = desc_hi | (al + 2u * kk), db = desc_hi | (b16 + 2u * kk);
              umma_tf32(tmem_d, dah, db, idesc2, (c | kk) ? 1u : 0u);
              umma_tf32(tmem_d + (uint32_t)n, dal, db, idesc1, 1u);
            }
            umma_commit(bar_w_free + 8 * ws);
          }
          umma_commit(bar_acc_full + 8 * buf);
        }
        __syncwarp();
        wg += nchunks;
      }
    }
  } else {
    // ===================== input rows, epilogues, segment max =====================
    const int r = (warp & 3) * 32 + lane, h = warp >> 2;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const int ctid = (int)threadIdx.x;
    float4 g[ITEMS];
    int seg_next = -1;
    auto prefetch = [&](int tile) {
      const int base = tile * 128;
      seg_next = base + r < a.n ? a.seg[base + r] : -1;
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const int e = it * 256 + ctid, p = e / QPR, q = e % QPR;
        const int j = base + p < a.n ? base + p : base;  // padding rows copy the tile's first row (never reduced)
        const int src = a.src_idx ? a.src_idx[j] : j;
        g[it] = reinterpret_cast<const float4*>(a.feat + (size_t)src * K0)[q];
      }
    };
    if ((int)blockIdx.x < a.n_tiles) prefetch(blockIdx.x);
    int job = 0;
    const bool st_on = a.stats != nullptr && threadIdx.x == 0;
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = st_on ? clock64() : 0;
#define PCAB_PHASE(kk)                     \
  if (st_on) {                             \
    const long long tnow = clock64();      \
    ph[kk] += tnow - tprev, tprev = tnow;  \
  }
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const int seg = seg_next;
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const int e = it * 256 + ctid, p = e / QPR, q = e % QPR;
        put4(a_hi, a_lo, p, 4 * q, g[it].x, g[it].y, g[it].z, g[it].w);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(bar_a_ready);
      PCAB_PHASE(0)
#pragma unroll
      for (int l = 0; l < NL; ++l, ++job) {
        const int buf = job & 1;
        if (l == NL - 1 && tile + (int)gridDim.x < a.n_tiles) prefetch(tile + gridDim.x);  // flies while the last layer's MMAs run
        mbar_wait(bar_acc_full + 8 * buf, (job >> 1) & 1);
        PCAB_PHASE(1 + 2 * l)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tm = tmem_base + lane_addr + (uint32_t)(buf * 256);
        const int n = NO[l], per = n / 2;
        if (l < NL - 1) {
          auto epi = [&](auto hc) {
            constexpr int HH = decltype(hc)::value;
#pragma unroll
            for (int b = 0; b < NO[l] / 32; ++b) {
              uint32_t vm[16], vc[16];
              const int c0 = HH * per + b * 16;
              tmem_ld16(tm + (uint32_t)c0, vm);
              tmem_ld16(tm + (uint32_t)(n + c0), vc);
              tmem_ld_wait16(vm, vc);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  v[u] = fmaxf((__uint_as_float(vm[4 * q + u]) + __uint_as_float(vc[4 * q + u])) + k.b[l][c0 + 4 * q + u], 0.f);
                put4(a_hi, a_lo, r, c0 + 4 * q, v[0], v[1], v[2], v[3]);
              }
            }
          };
          if (h == 0) epi(std::integral_constant<int, 0>{}); else epi(std::integral_constant<int, 1>{});
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);
          mbar_arrive(bar_a_ready);
          PCAB_PHASE(2 + 2 * l)
        } else {
          // last layer (no activation).  Its MMAs are done, so the A planes are free: the tile's outputs go there
          // channel-major ([128 ch][132]), and after a barrier thread (channel, half of the rows) scans its 64 rows: a run
          // strictly inside the scan range is a whole segment (plain store), runs touching an edge go through one atomic
          // max each.  (A shuffle-based segmented scan in registers costs 5 dependent SHFLs per value: 4x slower.)
          constexpr uint32_t LDT = 132;
          const uint32_t s_seg = sbase + 128u * LDT * 4u;
          auto epi = [&](auto hc) {
            constexpr int HH = decltype(hc)::value;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              uint32_t vm[16], vc[16];
              const int c0 = HH * 64 + b * 16;
              tmem_ld16(tm + (uint32_t)c0, vm);
              tmem_ld16(tm + (uint32_t)(128 + c0), vc);
              tmem_ld_wait16(vm, vc);
#pragma unroll
              for (int u = 0; u < 16; ++u) {
                const float v = (__uint_as_float(vm[u]) + __uint_as_float(vc[u])) + k.b[l][c0 + u];
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbase + ((uint32_t)(c0 + u) * LDT + (uint32_t)r) * 4u), "f"(v) : "memory");
              }
            }
          };
          if (h == 0) epi(std::integral_constant<int, 0>{}); else epi(std::integral_constant<int, 1>{});
          if (h == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(s_seg + 4u * r), "r"(seg) : "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);
          asm volatile("bar.sync 2, 256;" ::: "memory");
          {
            const int c = ctid & 127, r0 = (ctid >> 7) * 64, r1 = r0 + 64;
            int cur = -1, start = r0;
            float v = -INFINITY;
            for (int rr = r0; rr <= r1; ++rr) {
              int sg = -2;
              float x = 0.f;
              if (rr < r1) {
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(sg) : "r"(s_seg + 4u * rr));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(sbase + ((uint32_t)c * LDT + (uint32_t)rr) * 4u));
              }
              if (sg != cur) {
                if (cur >= 0) {
                  float* dst = a.out + (size_t)cur * 128 + c;
                  if (start > r0 && rr < r1) *dst = v; else mlp::atomic_max_float(dst, v);
                }
                cur = sg, start = rr, v = -INFINITY;
              }
              if (rr < r1) v = fmaxf(v, x);
            }
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");  // the next tile's input rows overwrite the planes
          PCAB_PHASE(2 + 2 * l)
        }
      }
    }
    if (st_on) {
      long long* sp = a.stats + blockIdx.x * 8;
      for (int q = 0; q < 8; ++q) sp[q] += ph[q];
    }
#undef PCAB_PHASE
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 9) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

__global__ void k_fill_f(float* __restrict__ a, long long n, float v) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) a[i] = v;
}
__global__ void k_neg_inf_to_zero(float* __restrict__ a, long long n) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
    if (a[i] == -INFINITY) a[i] = 0.f;  // torch_scatter leaves empty segments at 0
}

// TubeNet positional embedding, layer 0 on the CUDA cores: [p - anchor centroid(inst), t/T] (4) -> 32, ReLU; rows [n][32]
__global__ void k_tpn_pos_l0(const float* __restrict__ pts, const int* __restrict__ inst, const int* __restrict__ tidx, int n,
                             int T, const double* __restrict__ sums, const float* __restrict__ pk /* W[4][32] b[32] */,
                             float* __restrict__ rows, int* __restrict__ seg) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, q = threadIdx.x & 7;  // 8 threads per row, 4 outputs each
  if (j >= n) return;
  const int ki = inst[j], t = tidx[j];
  const double* s = sums + (size_t)ki * T * 4;  // anchor frame (t = 0) of the instance
  const double cnt = s[3] > 0 ? s[3] : 1.0;
  const float x[4] = {pts[3 * j] - (float)(s[0] / cnt), pts[3 * j + 1] - (float)(s[1] / cnt), pts[3 * j + 2] - (float)(s[2] / cnt),
                      (float)((double)t / (double)T)};
  float v[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int c = 4 * q + u;
    float acc = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) acc = fmaf(x[kk], pk[kk * 32 + c], acc);
    v[u] = fmaxf(acc + pk[128 + c], 0.f);
  }
  reinterpret_cast<float4*>(rows + (size_t)j * 32)[q] = make_float4(v[0], v[1], v[2], v[3]);
  if (q == 0) seg[j] = ki * T + t;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int encode_weights(EncodeTiledFn enc, CUtensorMap* map, const float* w, int rows, int K, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    pcab_set_error("mlp_tc: cuTensorMapEncodeTiled failed: %d", (int)r);
    return PCAB_ERR_CUDA;
  }
  return PCAB_OK;
}

}  // namespace

// floats in the tensor-core weight packs of the STPN head: pe2 [hi 64 rows; lo 64 rows][32] and
// [final_proj, mos0, off0] x [hi 128 rows; lo 128 rows][128]   (rows = output channels, K-major)
extern "C" size_t pcab_stpn_head_tc_pack_floats(int which) { return which == 0 ? (size_t)128 * 32 : (size_t)3 * 256 * 128; }

// Same contract as pcab_stpn_head (points.cu), except that the FP32 pack of that entry point is passed as a HOST pointer
// (biases, BN, pe0 and the two 128 -> 2 projections are copied from it into the kernel's parameter block); `w1_tc` / `w_tc`
// are the pre-split K-major device matrices described above.
extern "C" int pcab_stpn_head_tc(const float* mos_feats_nhwc, int feats_fmt, int H, int W, const float* transformed_points,
                                 const int* point_batch, const int* fg_idx, int n_fg, const float* weight_pack_host,
                                 const float* w1_tc, const float* w_tc, float x_abs, float y_abs, float* mos_out,
                                 float* offset_out, cudaStream_t stream) {
  if (n_fg <= 0) return PCAB_OK;
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    pcab_set_error("pcab_stpn_head_tc: cuTensorMapEncodeTiled unavailable");
    return PCAB_ERR_CUDA;
  }
  PCAB_REQUIRE(weight_pack_host != nullptr && ((uintptr_t)w1_tc & 15) == 0 && ((uintptr_t)w_tc & 15) == 0 &&
                   ((uintptr_t)mos_feats_nhwc & 15) == 0 && ((uintptr_t)mos_out & 7) == 0 && ((uintptr_t)offset_out & 7) == 0,
               "alignment of the weight packs / feature map / outputs");
  CUtensorMap m1, m2;
  int rc = encode_weights(enc, &m1, w1_tc, 128, 32, 128);
  if (rc != PCAB_OK) return rc;
  rc = encode_weights(enc, &m2, w_tc, 3 * 256, 128, 256);
  if (rc != PCAB_OK) return rc;
  HeadArgs a;
  a.mos_feats = mos_feats_nhwc, a.fmt = feats_fmt, a.H = H, a.W = W, a.tp = transformed_points, a.pbatch = point_batch, a.fg_idx = fg_idx;
  a.n_fg = n_fg, a.x_abs = x_abs, a.y_abs = y_abs, a.mos_out = mos_out, a.off_out = offset_out;
  a.n_tiles = cdiv(n_fg, 128);
  a.stats = nullptr;  // per-phase cycle counters (debug builds pass a device buffer here)
  static PcabSmemOnce once;
  PCAB_CUDA(pcab_set_max_smem(k_stpn_head_tc, (int)kSmemBytes, once));
  HeadConsts k;
  const float* pk = weight_pack_host;
  memcpy(k.pe0w, pk + S_PE0W, sizeof(k.pe0w)), memcpy(k.pe0b, pk + S_PE0B, sizeof(k.pe0b));
  memcpy(k.pe2b, pk + S_PE2B, sizeof(k.pe2b)), memcpy(k.fpb, pk + S_FPB, sizeof(k.fpb));
  const int hb[2] = {S_M0B, S_O0B}, hs[2] = {S_M0S, S_O0S}, ht[2] = {S_M0T, S_O0T}, h3w[2] = {S_M3W, S_O3W}, h3b[2] = {S_M3B, S_O3B};
  for (int i = 0; i < 2; ++i) {
    memcpy(k.hb[i], pk + hb[i], sizeof(k.hb[i])), memcpy(k.hs[i], pk + hs[i], sizeof(k.hs[i]));
    memcpy(k.ht[i], pk + ht[i], sizeof(k.ht[i])), memcpy(k.h3w[i], pk + h3w[i], sizeof(k.h3w[i]));
    memcpy(k.h3b[i], pk + h3b[i], sizeof(k.h3b[i]));
  }
  const int grid = a.n_tiles < pcab_sm_count() ? a.n_tiles : pcab_sm_count();
  k_stpn_head_tc<<<grid, kNT, kSmemBytes, stream>>>(m1, m2, a, k);
  PCAB_CHECK_LAUNCH("pcab_stpn_head_tc");
  return PCAB_OK;
}

namespace {
template <int K0, int N1, int N2>
int launch_embed(EncodeTiledFn enc, const float* feat, const int* src_idx, const int* seg, int n, int n_seg, const float* w0,
                 const float* w1, const float* w2, const float* b0, const float* b1, const float* b2, float* out,
                 cudaStream_t stream) {
  constexpr int NOUT1 = N2 ? N2 : 128;
  CUtensorMap m0, m1, m2;
  int rc = encode_weights(enc, &m0, w0, 2 * N1, K0, 2 * N1);
  if (rc != PCAB_OK) return rc;
  rc = encode_weights(enc, &m1, w1, 2 * NOUT1, N1, 2 * NOUT1);
  if (rc != PCAB_OK) return rc;
  if (N2) {
    rc = encode_weights(enc, &m2, w2, 256, N2, 256);
    if (rc != PCAB_OK) return rc;
  } else {
    m2 = m1;
  }
  EmbedConsts k;
  memset(&k, 0, sizeof(k));
  memcpy(k.b[0], b0, N1 * sizeof(float));
  memcpy(k.b[1], b1, NOUT1 * sizeof(float));
  if (N2) memcpy(k.b[2], b2, 128 * sizeof(float));
  EmbedArgs a;
  a.feat = feat, a.src_idx = src_idx, a.seg = seg, a.n = n, a.out = out, a.n_tiles = cdiv(n, 128);
  a.stats = nullptr;  // per-phase cycle counters (debug builds pass a device buffer here)
  static PcabSmemOnce once;
  PCAB_CUDA(pcab_set_max_smem(k_embed_tc<K0, N1, N2>, (int)kSmemBytes, once));
  const long long total = (long long)n_seg * 128;
  k_fill_f<<<grid_for(total, 256), 256, 0, stream>>>(out, total, -INFINITY);
  const int grid = a.n_tiles < pcab_sm_count() ? a.n_tiles : pcab_sm_count();
  k_embed_tc<K0, N1, N2><<<grid, kNT, kSmemBytes, stream>>>(m0, m1, m2, a, k);
  k_neg_inf_to_zero<<<grid_for(total, 256), 256, 0, stream>>>(out, total);
  return PCAB_OK;
}
}  // namespace

// TubeNet embeddings on the tensor cores.  which: 0 = motion_embed (64 -> 64 -> 128 -> 128), 1 = geo_embed (32 -> 32 -> 64 ->
// 128), 2 = layers 1-2 of pos_embed (32 -> 64 -> 128; `feat` = the rows written by pcab_tpn_pos_l0).  w*_tc: [hi rows; lo rows]
// K-major per layer (tc_pack.split_tf32); bias_host: the layers' biases back to back, on the HOST.  `seg` ascending.
extern "C" int pcab_embed_segmax_tc(int which, const float* feat, const int* src_idx, const int* seg, int n, int n_seg,
                                    const float* w0_tc, const float* w1_tc, const float* w2_tc, const float* bias_host,
                                    float* out /* [n_seg,128] */, cudaStream_t stream) {
  if (n <= 0 || n_seg <= 0) return PCAB_OK;
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    pcab_set_error("pcab_embed_segmax_tc: cuTensorMapEncodeTiled unavailable");
    return PCAB_ERR_CUDA;
  }
  PCAB_REQUIRE(bias_host != nullptr && ((uintptr_t)feat & 15) == 0 && ((uintptr_t)w0_tc & 15) == 0 && ((uintptr_t)w1_tc & 15) == 0,
               "alignment / null arguments");
  int rc;
  if (which == 0)
    rc = launch_embed<64, 64, 128>(enc, feat, src_idx, seg, n, n_seg, w0_tc, w1_tc, w2_tc, bias_host, bias_host + 64, bias_host + 192, out, stream);
  else if (which == 1)
    rc = launch_embed<32, 32, 64>(enc, feat, src_idx, seg, n, n_seg, w0_tc, w1_tc, w2_tc, bias_host, bias_host + 32, bias_host + 96, out, stream);
  else if (which == 2)
    rc = launch_embed<32, 64, 0>(enc, feat, src_idx, seg, n, n_seg, w0_tc, w1_tc, nullptr, bias_host, bias_host + 64, nullptr, out, stream);
  else {
    pcab_set_error("pcab_embed_segmax_tc: which must be 0, 1 or 2");
    return PCAB_ERR_ARG;
  }
  if (rc != PCAB_OK) return rc;
  PCAB_CHECK_LAUNCH("pcab_embed_segmax_tc");
  return PCAB_OK;
}

// layer 0 of pos_embed for every (padded) TubeNet row + the (instance, frame) segment id of the row
extern "C" int pcab_tpn_pos_l0(const float* points, const int* inst, const int* tidx, int n, int T, const double* frame_sums,
                               const float* pack_pos /* device: W0[4][32] b0[32] .. */, float* rows /* [n,32] */,
                               int* seg /* [n] */, cudaStream_t stream) {
  if (n <= 0) return PCAB_OK;
  k_tpn_pos_l0<<<cdiv(n * 8, 256), 256, 0, stream>>>(points, inst, tidx, n, T, frame_sums, pack_pos, rows, seg);
  PCAB_CHECK_LAUNCH("pcab_tpn_pos_l0");
  return PCAB_OK;
}
