// conv3x3 (pad 1) and ConvTranspose 2x2 on the 5th-generation tensor cores over PAIR-PACKED activations (pair16.cuh).
//
// Replaces the cuDNN calls behind models/unet.py:11-20,57-62,88-97 and models/stpn.py:13-22 (same contract as
// pcab_conv3x3_f32: multi-source accumulate = concat / temporal 3x3x3, bias, BN(eval), ReLU).
//
// Arithmetic: every operand is an fp16 pair (x = h + l, 22 significant bits), three products per algorithmic MAC
//   D[:, 0:C]  = a_h . w_h                      (main half of one N = 2C MMA:  a_h x [w_h | w_l])
//   D[:, C:2C] = a_h . w_l  +  a_l . w_h        (second half of that MMA + one N = C MMA)
// tcgen05.mma kind::f16, FP32 accumulators in TMEM, drained per 32-input-channel chunk into FP32 registers (the tensor
// core truncates when it adds into an accumulator; chunk-wise draining bounds that chain to 18 steps whatever the layer).
//
// What is new against the float32-activation kernel of conv_tc.cu (kept as the `tf32` / float32-activation path):
//   * activations arrive ALREADY split: the producing layer's epilogue wrote [32 ch h | 32 ch l] per pixel, so the TMA
//     halo plane [(R+2) x (Wt+2) pixels][128 B] IS the swizzled K-major A operand (bytes 0-63 = a_h, 64-127 = a_l; the
//     nine taps are nine shifted descriptor views of it).  The four operand-split warps and their shared-memory pass
//     (the bottleneck of the 32-channel layers) are gone; their planes became pipeline stages (up to 4 planes in flight);
//   * the epilogue packs its outputs to the same format, stages them in shared memory in the TMA box layout and ONE
//     thread issues cp.async.bulk.tensor stores: full 128-byte lines, image-edge clipping by the TMA unit;
//   * output-channel tiles of 96 and 128 (main MMA N = 192 / 256): the A plane is read once per 2C + C columns, which is
//     what moves the MMAs from shared-memory-operand bound (N = 64: 4 KB of A per 2 KB of B) to math bound;
//   * ConvTranspose2d(2, stride 2) runs on the same pipeline as a 1-tap convolution with 4 x Cout output columns whose
//     channel groups are scattered to the four (dy, dx) positions through strided output tensor maps.
// Warp roles (512 threads): 0-11 drain + epilogue (4 TMEM lane quarters x up to 3 channel groups) | 12 plane TMA |
// 13 weight TMA | 14, 15 MMA issue (one M tile each; 14 owns TMEM).
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "pair16.cuh"
#include "tc_common.cuh"
#include "pcab200.h"

namespace {

using namespace pcab_tc;

constexpr int kThreads = 512;
constexpr int kWStages = 3;
constexpr int kMaxSmem = 227 * 1024;

struct Args {
  int nsrc;
  int src_c[3];
  int T;
  int N, H, W, Cout;  // Cout: output columns of the GEMM (ConvT: 4 x the layer's output channels)
  int mt;             // M tiles of 128 rows per work item (1 or 2)
  int strip;          // 1: strip tiles (8 px wide groups), 0: flattened padded grid
  int mtx;            // strip mode: M tiles side by side
  int R, Wt, Wp;      // tile rows / columns, plane pitch in pixels (Wt + 2 halo columns, Wt for 1-tap)
  int tiles_x, tiles_y, n_ctile, total_items;
  int np;             // plane pipeline stages
  uint32_t plane_bytes;
  int ntaps;          // 9 (conv3x3) or 1 (ConvTranspose positions)
  int resident;       // 1: all weights of the (single) column tile stay in shared memory for the life of the CTA
  int wslots;         // weight stages in shared memory: all of them (resident) or a ring of kWStages
  int kg_src[3];      // first K group (32 input channels of one tap, consumption order) of each source
  int kg_total;
  int relu;
  int cout_real;      // ConvT: output channels of the layer (bias index = column % cout_real); else == Cout
  float wscale_inv;
  const float* bias;
  const float* bn_scale;
  const float* bn_shift;
  unsigned int* sat_counter;  // incremented when an output beyond the fp16 range was clamped (may be null)
};

struct Item {
  int n, x0, y0, co0, tframe;
};
__device__ __forceinline__ Item decode_item(const Args& a, int C, int it) {
  Item r;
  const int ct = it % a.n_ctile, sp = it / a.n_ctile;
  const int tx = sp % a.tiles_x, ty = (sp / a.tiles_x) % a.tiles_y;
  r.n = sp / (a.tiles_x * a.tiles_y);
  r.x0 = tx * a.Wt, r.y0 = ty * a.R, r.co0 = ct * C;
  r.tframe = a.T > 1 ? r.n % a.T : 0;
  return r;
}
__device__ __forceinline__ bool src_valid(const Args& a, int s, int tframe) {
  return a.T <= 1 || (tframe + s - 1 >= 0 && tframe + s - 1 < a.T);
}
__device__ __forceinline__ int item_chunks(const Args& a, int tframe) {
  int n = 0;
  for (int s = 0; s < a.nsrc; ++s)
    if (src_valid(a, s, tframe)) n += a.src_c[s] / 32;
  return n;
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// C: output columns per work item.  NT: taps (9 = conv3x3, 1 = ConvTranspose).  Up to four output maps: conv uses
// map_o0 only; ConvT picks map_o[(dy, dx) position of the channel group].
template <int C, int NT>
__global__ void __launch_bounds__(kThreads, 1)
k_conv_p16(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
           const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b,
           const __grid_constant__ CUtensorMap map_o0, const __grid_constant__ CUtensorMap map_o1,
           const __grid_constant__ CUtensorMap map_o2, const __grid_constant__ CUtensorMap map_o3, Args a) {
  constexpr int CH = (C == 32) ? 16 : (C == 128 ? 64 : 32);  // channels per epilogue thread
  constexpr int NG = C / CH;                                  // epilogue warp groups (4 warps each)
  constexpr int MT_MAX = (C <= 64) ? 2 : 1;
  constexpr int NEPI = 128 * NG;
  constexpr uint32_t kWBytes = 2u * C * 128u;  // one tap stage: [w_h rows | w_l rows]
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w0 = sbase + (uint32_t)a.np * a.plane_bytes;
  const uint32_t stg0 = w0 + (uint32_t)a.wslots * kWBytes;
  const uint32_t stg_bytes = (uint32_t)a.mt * 16384u;  // one 32-channel group of the item's pixels, 128 B per pixel
  const uint32_t bars = stg0 + (uint32_t)(C / 32) * stg_bytes;
  const uint32_t bar_plane_full = bars, bar_plane_free = bars + 32, bar_w_full = bars + 64, bar_w_free = bars + 88,
                 bar_acc_full = bars + 112, bar_acc_empty = bars + 128, tmem_slot = bars + 144;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while (tmem_cols < 4u * a.mt * C) tmem_cols <<= 1;  // 2 stages x mt tiles x [main C | corr C]
  // Programmatic dependent launch: the next kernel of the stream may be scheduled now (its barrier / TMEM / descriptor
  // set-up then overlaps this kernel's tail on SMs that are already idle); it waits for THIS grid to complete before it
  // touches global memory (griddepcontrol.wait below does the same for us against our predecessor).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bar_plane_full + 8 * i, 1), mbar_init(bar_plane_free + 8 * i, a.mt);
    for (int i = 0; i < kWStages; ++i) mbar_init(bar_w_full + 8 * i, 1), mbar_init(bar_w_free + 8 * i, a.mt);
    for (int i = 0; i < 2; ++i) mbar_init(bar_acc_full + 8 * i, a.mt), mbar_init(bar_acc_empty + 8 * i, NEPI);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 14) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 12 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_o0)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the producer of our inputs (and of the buffers we overwrite) is done

  constexpr int HALO = (NT == 9) ? 1 : 0;
  const uint32_t box_bytes = (uint32_t)(a.R + 2 * HALO) * a.Wp * 128u;

  if (warp == 12) {
    // ===================== plane producer =====================
    if (lane == 0) {
      int g = 0;
      for (int it = blockIdx.x; it < a.total_items; it += gridDim.x) {
        const Item t = decode_item(a, C, it);
        for (int s = 0; s < a.nsrc; ++s) {
          if (!src_valid(a, s, t.tframe)) continue;
          const CUtensorMap* am = a.T > 1 ? &map_a0 : (s == 0 ? &map_a0 : (s == 1 ? &map_a1 : &map_a2));
          const int img = a.T > 1 ? t.n + s - 1 : t.n;
          for (int c0 = 0; c0 < a.src_c[s]; c0 += 32, ++g) {
            const int ps = g % a.np;
            if (g >= a.np) mbar_wait(bar_plane_free + 8 * ps, (uint32_t)((g / a.np) - 1) & 1u);
            mbar_expect_tx(bar_plane_full + 8 * ps, box_bytes);
            tma_load_4d(am, sbase + (uint32_t)ps * a.plane_bytes, bar_plane_full + 8 * ps, 2 * c0, t.x0 - HALO, t.y0 - HALO, img);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 13) {
    // ===================== weight producer =====================
    // The weight matrix is K-dense in CONSUMPTION order (source, 32-channel chunk, tap): one 128-byte row = two K groups, one
    // stage = [w_h rows | w_l rows] of a pair of groups.  Resident mode loads every stage once; otherwise the stages an item
    // touches stream through a ring of kWStages slots.
    if (lane == 0) {
      if (a.resident) {
        const int nst = (a.kg_total + 1) >> 1;
        mbar_expect_tx(bar_w_full, (uint32_t)nst * kWBytes);
        for (int j = 0; j < nst; ++j) {
          tma_load_2d(&map_b, w0 + (uint32_t)j * kWBytes, bar_w_full, 64 * j, 0);
          tma_load_2d(&map_b, w0 + (uint32_t)j * kWBytes + C * 128u, bar_w_full, 64 * j, a.Cout);
        }
      } else {
        int wc = 0;
        for (int it = blockIdx.x; it < a.total_items; it += gridDim.x) {
          const Item t = decode_item(a, C, it);
          int prev_j = -1;
          for (int s = 0; s < a.nsrc; ++s) {
            if (!src_valid(a, s, t.tframe)) continue;
            const int ng = (a.src_c[s] / 32) * NT;
            for (int kg = a.kg_src[s]; kg < a.kg_src[s] + ng; ++kg) {
              const int j = kg >> 1;
              if (j == prev_j) continue;
              prev_j = j;
              const int ws = wc % kWStages;
              if (wc >= kWStages) mbar_wait(bar_w_free + 8 * ws, (uint32_t)((wc / kWStages) - 1) & 1u);
              mbar_expect_tx(bar_w_full + 8 * ws, kWBytes);
              tma_load_2d(&map_b, w0 + ws * kWBytes, bar_w_full + 8 * ws, 64 * j, t.co0);
              tma_load_2d(&map_b, w0 + ws * kWBytes + C * 128u, bar_w_full + 8 * ws, 64 * j, a.Cout + t.co0);
              ++wc;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 14) {
    // ===================== MMA issuers =====================
    // warp 14 owns M tile 0, warp 15 M tile 1 (disjoint accumulators: no ordering between them).  The whole warp runs the
    // loop with warp-uniform state; the tcgen05 instructions themselves are issued by one fixed lane.
    const int mi = warp - 14;
    if (mi < a.mt) {
      const bool leader = elect_one();
      const uint32_t idesc_base = (1u << 4) | ((128u >> 4) << 24);  // D = F32, A = B = F16, K-major both, M = 128
      const uint32_t idesc2 = idesc_base | ((uint32_t)((2 * C) >> 3) << 17);
      const uint32_t idesc1 = idesc_base | ((uint32_t)(C >> 3) << 17);
      const uint32_t sbo_a = a.strip ? (uint32_t)a.Wp * 8u : 64u;  // stride between 8-row groups, in 16 B units
      const uint64_t desc_hi_a = (uint64_t)(sbo_a | (1u << 14) | (2u << 29)) << 32;
      const uint64_t desc_hi_b = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;
      const uint32_t lbo = 1u << 16;
      const uint32_t tile_off16 =
          a.strip ? ((uint32_t)(mi % a.mtx) * 8u + (uint32_t)(mi / a.mtx) * 16u * (uint32_t)a.Wp) * 8u : (uint32_t)mi * 1024u;
      const uint32_t wp8 = (uint32_t)a.Wp * 8u;
      const uint32_t a_base = lbo | (((sbase & 0x3FFFF) >> 4) + tile_off16);
      const uint32_t b_base = lbo | ((w0 & 0x3FFFF) >> 4);
      int g = 0, wc = 0;
      if (a.resident) {
        mbar_wait(bar_w_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      for (int it = blockIdx.x; it < a.total_items; it += gridDim.x) {
        const Item t = decode_item(a, C, it);
        int prev_j = -1, slot = 0;
        for (int s = 0; s < a.nsrc; ++s) {
          if (!src_valid(a, s, t.tframe)) continue;
          for (int c0 = 0, kg0 = a.kg_src[s]; c0 < a.src_c[s]; c0 += 32, kg0 += NT, ++g) {
            const int st = g & 1, ps = g % a.np;
            mbar_wait(bar_plane_full + 8 * ps, (uint32_t)(g / a.np) & 1u);
            if (g >= 2) mbar_wait(bar_acc_empty + 8 * st, (uint32_t)((g >> 1) - 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ap = a_base + (uint32_t)ps * (a.plane_bytes >> 4);
            const uint32_t tmem_d = tmem_base + (uint32_t)(st * a.mt * 2 * C + mi * 2 * C);
#pragma unroll
            for (int tap = 0; tap < NT; ++tap) {
              const int kg = kg0 + tap;
              uint32_t bst;
              if (a.resident) {
                bst = b_base + (uint32_t)(kg >> 1) * (kWBytes >> 4);
              } else {
                const int j = kg >> 1;
                if (j != prev_j) {  // first group of a new weight stage: release the previous slot, wait for the next one
                  if (prev_j >= 0 && leader) umma_commit(bar_w_free + 8 * slot);
                  slot = wc % kWStages;
                  mbar_wait(bar_w_full + 8 * slot, (uint32_t)(wc / kWStages) & 1u);
                  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                  ++wc, prev_j = j;
                }
                bst = b_base + (uint32_t)slot * (kWBytes >> 4);
              }
              const uint32_t b16 = bst + (uint32_t)(kg & 1) * 4u;  // second group of the pair: bytes 64-127 of the rows
              const uint32_t shift16 = NT == 9 ? (uint32_t)(tap / 3) * wp8 + (uint32_t)(tap % 3) * 8u : 0u;
              if (leader) {
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                  const uint64_t dah = desc_hi_a | (ap + shift16 + 2u * kk);       // bytes 0-63 of a row: a_h
                  const uint64_t dal = desc_hi_a | (ap + shift16 + 4u + 2u * kk);  // bytes 64-127: a_l
                  const uint64_t db = desc_hi_b | (b16 + 2u * kk);
                  umma_f16(tmem_d, dah, db, idesc2, (tap == 0 && kk == 0) ? 0u : 1u);  // [a_h*w_h | a_h*w_l]
                  umma_f16(tmem_d + (uint32_t)C, dal, db, idesc1, 1u);                // += a_l*w_h into the second half
                }
              }
            }
            if (leader) {
              umma_commit(bar_plane_free + 8 * ps);
              umma_commit(bar_acc_full + 8 * st);
            }
            __syncwarp();
          }
        }
        if (!a.resident && prev_j >= 0 && leader) umma_commit(bar_w_free + 8 * slot);
        __syncwarp();
      }
    }
  } else if (warp < 4 * NG) {
    // ===================== drain + epilogue =====================
    // thread = one TMEM lane (pixel of an M tile) x CH output columns of the tile
    const int quarter = warp & 3, grp = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    float acc[MT_MAX][CH];
    int g = 0;
    bool stored = false;
    for (int it = blockIdx.x; it < a.total_items; it += gridDim.x) {
      const Item t = decode_item(a, C, it);
      const int nchunks = item_chunks(a, t.tframe);
      for (int ci = 0; ci < nchunks; ++ci, ++g) {
        const int st = g & 1;
        mbar_wait(bar_acc_full + 8 * st, (uint32_t)(g >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tm_stage = tmem_base + lane_addr + (uint32_t)(st * a.mt * 2 * C);
#pragma unroll
        for (int mi = 0; mi < MT_MAX; ++mi) {
          if (mi < a.mt) {
#pragma unroll
            for (int b16 = 0; b16 < CH / 16; ++b16) {
              uint32_t vm[16], vc[16];
              const uint32_t col = (uint32_t)(mi * 2 * C + grp * CH + b16 * 16);
              tmem_ld16(tm_stage + col, vm);
              tmem_ld16(tm_stage + col + (uint32_t)C, vc);
              tmem_ld_wait16(vm, vc);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float v = __uint_as_float(vm[j]) + __uint_as_float(vc[j]);
                acc[mi][b16 * 16 + j] = ci == 0 ? v : acc[mi][b16 * 16 + j] + v;
              }
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(bar_acc_empty + 8 * st);
      }
      // ---- epilogue of this work item: bias / BN / ReLU -> (h, l) fp16 pairs -> staging tile -> TMA store
      if (threadIdx.x == 0 && stored) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging is free again
      asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory");
      bool sat = false;
#pragma unroll
      for (int mi = 0; mi < MT_MAX; ++mi) {
        if (mi < a.mt) {
          const int m = quarter * 32 + lane;
          int r, xc;
          bool valid = true;
          if (a.strip) {
            r = (mi / a.mtx) * 16 + (m >> 3), xc = (mi % a.mtx) * 8 + (m & 7);
          } else {
            const int mm = mi * 128 + m;
            r = mm / a.Wp, xc = mm % a.Wp;
            valid = r < a.R && xc < a.Wt;
          }
          if (valid) {
            const uint32_t srow = (uint32_t)(r * a.Wt + xc), sw = srow & 7u;
#pragma unroll
            for (int q = 0; q < CH / 8; ++q) {
              const int cl = grp * CH + 8 * q;                 // column within the item's C columns
              const int cb = (t.co0 + cl) % a.cout_real;       // bias / BN channel (ConvT: column % layer channels)
              float o[8];
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + cb));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.bias + cb + 4));
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int u = 0; u < 8; ++u) o[u] = fmaf(acc[mi][8 * q + u], a.wscale_inv, bb[u]);
              if (a.bn_scale) {
                const float4 s0 = __ldg(reinterpret_cast<const float4*>(a.bn_scale + cb));
                const float4 s1 = __ldg(reinterpret_cast<const float4*>(a.bn_scale + cb + 4));
                const float4 h0 = __ldg(reinterpret_cast<const float4*>(a.bn_shift + cb));
                const float4 h1 = __ldg(reinterpret_cast<const float4*>(a.bn_shift + cb + 4));
                const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                const float hh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                for (int u = 0; u < 8; ++u) o[u] = fmaf(o[u], ss[u], hh[u]);
              }
              if (a.relu) {
#pragma unroll
                for (int u = 0; u < 8; ++u) o[u] = fmaxf(o[u], 0.f);
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) sat |= !(fabsf(o[u]) <= 65504.f);  // also catches NaN
              uint32_t h[4], l[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) p16::split2(o[2 * u], o[2 * u + 1], h[u], l[u]);
              const uint32_t row_addr = stg0 + (uint32_t)(cl >> 5) * stg_bytes + srow * 128u;
              const uint32_t ch = (uint32_t)(cl & 31) >> 3;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((ch ^ sw) << 4)), "r"(h[0]), "r"(h[1]),
                           "r"(h[2]), "r"(h[3])
                           : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + (((4u + ch) ^ sw) << 4)), "r"(l[0]),
                           "r"(l[1]), "r"(l[2]), "r"(l[3])
                           : "memory");
            }
          }
        }
      }
      if (sat && a.sat_counter) atomicAdd(a.sat_counter, 1u);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA unit
      asm volatile("bar.sync 2, %0;" ::"n"(NEPI) : "memory");
      if (threadIdx.x == 0) {
#pragma unroll
        for (int gi = 0; gi < C / 32; ++gi) {
          const int cg = t.co0 + 32 * gi;  // first output column of this 32-channel group
          if (NT == 9) {
            tma_store_4d(&map_o0, stg0 + (uint32_t)gi * stg_bytes, 2 * cg, t.x0, t.y0, t.n);
          } else {
            const int pos = cg / a.cout_real, cc = cg % a.cout_real;  // (dy, dx) position, channel inside the layer's output
            const CUtensorMap* om = pos == 0 ? &map_o0 : (pos == 1 ? &map_o1 : (pos == 2 ? &map_o2 : &map_o3));
            tma_store_4d(om, stg0 + (uint32_t)gi * stg_bytes, 2 * cc, t.x0, t.y0, t.n);
          }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        stored = true;
      }
    }
    if (threadIdx.x == 0 && stored) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 14) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// =====================================================================================================================
// Conv3d 3x3x3 (32 -> 32 channels, models/stpn.py:13-22) with the TEMPORAL taps fused into the MMA N dimension.
// The generic kernel runs a Conv3d as three 2-D K-slices per output frame: every input plane is staged and streamed
// through the tensor core three times with N = 64 + 32, where the MMAs are bound by the shared-memory reads of A.
// Here a CTA owns a 128-pixel tile of ONE scene for all T frames: input frame f is staged once and multiplied against
// [W(kt=2); W(kt=1); W(kt=0)] (N = 96), i.e. it contributes to output frames f-1, f, f+1 at once.  The accumulators of the
// output frames form two rings of eight 32-column blocks in TMEM (main sums at columns 0..255, correction sums at
// 256..511; output frame g lives in block (g+1) % 8), the three products are three N = 96 MMAs on a 96-column window
// of those rings (split in two at the ring wrap), and an output frame is drained, stored and its block zeroed again as soon
// as the input frame after it has been consumed.  A read per algorithmic MAC drops 3x, the MMAs come close to math bound.
// Warp roles (512 threads): 0-7 drain + epilogue | 12 plane TMA | 13 weights (once) | 14 MMA issue.
// =====================================================================================================================
struct Args3d {
  int B, T, H, W;
  int tiles_x, tiles_y, total_items;
  int np;
  uint32_t plane_bytes;
  int relu;
  float wscale_inv;
  const float* bias;
  unsigned int* sat_counter;
};

constexpr uint32_t kW3dStage = 192u * 128u;  // one pair of taps: [w_h rows of kt = 2, 1, 0 | w_l rows of kt = 2, 1, 0] x 128 B
constexpr int kW3dStages = 5;                // nine taps, two per 128-byte row

__global__ void __launch_bounds__(kThreads, 1)
k_conv3d_p16(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
             const __grid_constant__ CUtensorMap map_o, Args3d a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w0 = sbase + (uint32_t)a.np * a.plane_bytes;
  const uint32_t stg0 = w0 + kW3dStages * kW3dStage;
  const uint32_t bars = stg0 + 16384u;
  const uint32_t bar_plane_full = bars, bar_plane_free = bars + 32, bar_w_full = bars + 64, bar_fd = bars + 72, bar_fc = bars + 88,
                 tmem_slot = bars + 104;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bar_plane_full + 8 * i, 1), mbar_init(bar_plane_free + 8 * i, 1);
    mbar_init(bar_w_full, 1);
    for (int i = 0; i < 2; ++i) mbar_init(bar_fd + 8 * i, 1), mbar_init(bar_fc + 8 * i, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 14) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 12 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_o)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
  // every accumulator block starts at zero (all MMAs accumulate): the epilogue warps clear the 512 columns once
  if (warp < 8) {
    const uint32_t t0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 256u;
#pragma unroll 1
    for (int c = 0; c < 256; c += 16)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(t0 + c), "r"(0u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  constexpr int Wp = 10, R = 16;  // strip tile: 8 x 16 output pixels, halo plane 10 x 18
  const uint32_t box_bytes = 18u * Wp * 128u;
  auto item_of = [&](int it, int& b, int& x0, int& y0) {
    const int tx = it % a.tiles_x, ty = (it / a.tiles_x) % a.tiles_y;
    b = it / (a.tiles_x * a.tiles_y), x0 = tx * 8, y0 = ty * R;
  };

  if (warp == 12) {
    // ===================== plane producer: one halo plane per input frame =====================
    if (lane == 0) {
      int g = 0;
      for (int it = blockIdx.x; it < a.total_items; it += gridDim.x) {
        int b, x0, y0;
        item_of(it, b, x0, y0);
        for (int f = 0; f < a.T; ++f, ++g) {
          const int ps = g % a.np;
          if (g >= a.np) mbar_wait(bar_plane_free + 8 * ps, (uint32_t)((g / a.np) - 1) & 1u);
          mbar_expect_tx(bar_plane_full + 8 * ps, box_bytes);
          tma_load_4d(&map_a, sbase + (uint32_t)ps * a.plane_bytes, bar_plane_full + 8 * ps, 0, x0 - 1, y0 - 1, b * a.T + f);
        }
      }
    }
    __syncwarp();
  } else if (warp == 13) {
    // ===================== weights: resident for the life of the CTA =====================
    if (lane == 0) {
      mbar_expect_tx(bar_w_full, kW3dStages * kW3dStage);
      for (int j = 0; j < kW3dStages; ++j) tma_load_2d(&map_b, w0 + (uint32_t)j * kW3dStage, bar_w_full, 64 * j, 0);
    }
    __syncwarp();
  } else if (warp == 14) {
    // ===================== MMA issue =====================
    const bool leader = elect_one();
    const uint32_t idesc_base = (1u << 4) | ((128u >> 4) << 24);  // D = F32, A = B = F16, K-major both, M = 128
    auto idesc = [&](uint32_t n) { return idesc_base | ((n >> 3) << 17); };
    const uint64_t desc_hi_a = (uint64_t)((uint32_t)Wp * 8u | (1u << 14) | (2u << 29)) << 32;  // strip: 8-row groups are image rows
    const uint64_t desc_hi_b = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;
    const uint32_t lbo = 1u << 16;
    const uint32_t a_base = lbo | ((sbase & 0x3FFFF) >> 4);
    const uint32_t b_base = lbo | ((w0 & 0x3FFFF) >> 4);
    mbar_wait(bar_w_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    int F = 0;  // global frame counter of this CTA
    for (int it = blockIdx.x; it < a.total_items; it += gridDim.x) {
      for (int f = 0; f < a.T; ++f, ++F) {
        const int ps = F % a.np, sl = F & 1;
        mbar_wait(bar_plane_full + 8 * ps, (uint32_t)(F / a.np) & 1u);
        // back-pressure: the epilogue has finished the duties of frame F-2 (of F-1 at the start of an item: the previous
        // item's last output blocks are drained and cleared before any of them is accumulated into again)
        if (F >= 2) mbar_wait(bar_fc + 8 * sl, (uint32_t)((F >> 1) - 1) & 1u);
        if (f == 0 && F >= 1) mbar_wait(bar_fc + 8 * (sl ^ 1), (uint32_t)((F - 1) >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ap = a_base + (uint32_t)ps * (a.plane_bytes >> 4);
        const int j0 = f & 7;  // ring block of output frame f-1; the window is blocks j0, j0+1, j0+2 (mod 8)
        if (leader) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t shift16 = (uint32_t)(tap / 3) * (Wp * 8u) + (uint32_t)(tap % 3) * 8u;
            const uint32_t bst = b_base + (uint32_t)(tap >> 1) * (kW3dStage >> 4) + (uint32_t)(tap & 1) * 4u;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t dah = desc_hi_a | (ap + shift16 + 2u * kk), dal = desc_hi_a | (ap + shift16 + 4u + 2u * kk);
              // B rows: [0, 96) = w_h of kt = 2, 1, 0 ; [96, 192) = w_l of kt = 2, 1, 0  (128 B per row: 8 x 16 B)
              const uint32_t bh = bst + 2u * kk, bl = bst + 96u * 8u + 2u * kk;
              auto three = [&](uint32_t col, uint32_t row_off, uint32_t n) {
                const uint32_t ro = row_off * 8u;
                umma_f16(tmem_base + col, dah, desc_hi_b | (bh + ro), idesc(n), 1u);          // main ring  += a_h . w_h
                umma_f16(tmem_base + 256u + col, dah, desc_hi_b | (bl + ro), idesc(n), 1u);   // corr ring  += a_h . w_l
                umma_f16(tmem_base + 256u + col, dal, desc_hi_b | (bh + ro), idesc(n), 1u);   // corr ring  += a_l . w_h
              };
              if (j0 <= 5) {
                three(32u * j0, 0u, 96u);
              } else if (j0 == 6) {
                three(192u, 0u, 64u);
                three(0u, 64u, 32u);
              } else {
                three(224u, 0u, 32u);
                three(0u, 32u, 64u);
              }
            }
          }
          umma_commit(bar_plane_free + 8 * ps);
          umma_commit(bar_fd + 8 * sl);
        }
        __syncwarp();
      }
    }
  } else if (warp < 8) {
    // ===================== drain + epilogue =====================
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int m = quarter * 32 + lane;
    const uint32_t srow = (uint32_t)m, sw = srow & 7u;  // strip tile: staging row = TMEM lane ((m >> 3) * 8 + (m & 7))
    bool stored = false, sat = false;
    int F = 0;
    auto clear_block = [&](int j) {
      const uint32_t t = tmem_base + lane_addr + 32u * j + 16u * half;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(t), "r"(0u) : "memory");
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(t + 256u), "r"(0u) : "memory");
    };
    for (int it = blockIdx.x; it < a.total_items; it += gridDim.x) {
      int b, x0, y0;
      item_of(it, b, x0, y0);
      auto finalize = [&](int g) {  // output frame g is complete: bias / ReLU -> pairs -> staging -> TMA store; clear its block
        const int j = (g + 1) & 7;
        uint32_t vm[16], vc[16];
        tmem_ld16(tmem_base + lane_addr + 32u * j + 16u * half, vm);
        tmem_ld16(tmem_base + lane_addr + 256u + 32u * j + 16u * half, vc);
        tmem_ld_wait16(vm, vc);
        clear_block(j);
        if (threadIdx.x == 0 && stored) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int cb = 16 * half + 8 * q;
          float o[8];
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + cb)), b1 = __ldg(reinterpret_cast<const float4*>(a.bias + cb + 4));
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            o[u] = fmaf(__uint_as_float(vm[8 * q + u]) + __uint_as_float(vc[8 * q + u]), a.wscale_inv, bb[u]);
            if (a.relu) o[u] = fmaxf(o[u], 0.f);
            sat |= !(fabsf(o[u]) <= 65504.f);
          }
          uint32_t h[4], l[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) p16::split2(o[2 * u], o[2 * u + 1], h[u], l[u]);
          const uint32_t row_addr = stg0 + srow * 128u, ch = (uint32_t)cb >> 3;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((ch ^ sw) << 4)), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + (((4u + ch) ^ sw) << 4)), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (threadIdx.x == 0) {
          tma_store_4d(&map_o, stg0, 0, x0, y0, b * a.T + g);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          stored = true;
        }
      };
      for (int f = 0; f < a.T; ++f, ++F) {
        const int sl = F & 1;
        mbar_wait(bar_fd + 8 * sl, (uint32_t)(F >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (f == 0)
          clear_block(0);  // the block of "output frame -1" only collected the kt = 2 products of frame 0
        else
          finalize(f - 1);
        if (f == a.T - 1) {
          finalize(a.T - 1);
          clear_block((a.T + 1) & 7);  // "output frame T"
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(bar_fc + 8 * sl);
      }
    }
    if (sat && a.sat_counter) atomicAdd(a.sat_counter, 1u);
    if (threadIdx.x == 0 && stored) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 14) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// --------------------------------------------------------------------------------------------------------------------
// host side: tile plan, tensor maps, launch
// --------------------------------------------------------------------------------------------------------------------
struct Cfg {
  int c, mt, strip, mtx, R, Wt, Wp, tiles_x, tiles_y, n_ctile, total, np, plane_rows, resident, wslots;
  size_t smem;
};

// clocks of one (tap, k-step) MMA pair at M = 128: the larger of the math time (N/2 per MMA) and the shared-memory operand
// time at 128 B/clk (A 4 KB per MMA, B 32 N bytes)
int mma_clk(int c) { return c == 32 ? 88 : (c == 64 ? 112 : (c == 96 ? 144 : 192)); }

size_t smem_bytes(int c, int mt, int np, int plane_rows, int wslots) {
  return 1024 + (size_t)np * plane_rows * 128 + (size_t)wslots * 2 * c * 128 + (size_t)(c / 32) * mt * 16384 + 256;
}

// pick (columns per item, tile shape, resident / streamed weights) with the lowest modelled time over the SMs of the device.
// nchunks: 32-channel input chunks per item (all sources); the model charges every chunk its MMA time (shared-memory
// operand bound below N = 128), the shared-memory write time of what TMA brings in, and the L2 -> SM bandwidth of that
// traffic summed over the SMs (~6 KB/clk chip-wide): re-streaming the weights for every 128-pixel item is what bounds the
// wide column tiles unless the weights stay resident.
bool choose(int n_img, int H, int W, int Cout, int nchunks, int ntaps, Cfg* best) {
  if (Cout % 32) return false;
  const int halo = ntaps == 9 ? 1 : 0;
  const int nsm = pcab_sm_count();
  const int kg2 = (nchunks * ntaps + 1) / 2;  // weight stages (pairs of K groups) of one column tile
  double best_cost = 1e30;
  bool found = false;
  const int cands[4] = {128, 96, 64, 32};
  for (int ic = 0; ic < 4; ++ic) {
    const int c = cands[ic];
    if (Cout % c) continue;
    const int mt_max = c <= 64 ? 2 : 1;
    for (int strip = 0; strip <= 1; ++strip) {
      for (int mt = 1; mt <= mt_max; ++mt) {
        for (int mtx = 1; mtx <= (strip ? mt : 1); ++mtx) {
          for (int Wt = (strip ? 8 * mtx : 6); Wt <= (strip ? 8 * mtx : 41); ++Wt) {
            for (int resident = 0; resident <= 1; ++resident) {
              Cfg k;
              k.c = c, k.mt = mt, k.strip = strip, k.mtx = mtx, k.Wt = Wt, k.Wp = Wt + 2 * halo;
              k.n_ctile = Cout / c;
              if (resident && k.n_ctile != 1) continue;
              if (strip) {
                k.R = 16 * (mt / mtx);
              } else {
                k.R = (mt * 128) / k.Wp;
                if (k.R > H) k.R = H;
                if (Wt > W) continue;
              }
              if (k.R < 1 || k.R + 2 * halo > 256) continue;
              // last plane row a (shifted) view can touch, and the rows the TMA box fills
              const int last = strip ? ((k.R - 1 + 2 * halo) * k.Wp + (Wt - 8) + 2 * halo + 7) : (mt * 128 - 1 + 2 * halo * k.Wp + 2 * halo);
              int rows = (k.R + 2 * halo) * k.Wp;
              if (last + 1 > rows) rows = last + 1;
              k.plane_rows = (rows + 7) & ~7;
              k.resident = resident, k.wslots = resident ? kg2 : kWStages;
              k.np = 4;
              while (k.np >= 2 && smem_bytes(c, mt, k.np, k.plane_rows, k.wslots) > (size_t)kMaxSmem) --k.np;
              if (k.np < 2) continue;
              k.smem = smem_bytes(c, mt, k.np, k.plane_rows, k.wslots);
              k.tiles_x = cdiv(W, Wt), k.tiles_y = cdiv(H, k.R);
              k.total = n_img * k.tiles_x * k.tiles_y * k.n_ctile;
              const double rounds = (double)cdiv(k.total, nsm);
              const double mma = (double)mt * 2 * ntaps * mma_clk(c);
              const double tma_bytes = (double)(k.R + 2 * halo) * k.Wp * 128 + (resident ? 0.0 : (double)ntaps * 2 * c * 64);
              const int busy = k.total < nsm ? k.total : nsm;
              double chunk = mma + tma_bytes / 128.0;
              const double l2 = tma_bytes * busy / 6000.0;
              if (l2 > chunk) chunk = l2;
              const double item = (double)nchunks * (chunk + 250.0) + 1200.0 + 400.0 * mt * (c / 32);
              const double cost = rounds * item * (k.np >= 3 ? 1.0 : 1.03) + (resident ? (double)kg2 * 2 * c * 128 / 64.0 : 0.0);
              if (cost < best_cost - 1e-9) best_cost = cost, *best = k, found = true;
            }
          }
        }
      }
    }
  }
  return found;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// P16 activation tensor [n][H][W][C] viewed as fp16 {2C, W, H, n} with an optional (dy, dx) stride-2 scatter view
// (Cuse: channels addressed through the map, starting at `base`; C: channels per pixel of the tensor = the pixel pitch)
bool encode_act(EncodeTiledFn enc, CUtensorMap* m, const void* base, int Cuse, int C, int W, int H, int n, int box_w, int box_h,
                int dy = -1, int dx = 0) {
  cuuint64_t dims[4] = {(cuuint64_t)2 * Cuse, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  const char* p = reinterpret_cast<const char*>(base);
  if (dy >= 0) {  // ConvT output position: pixels (2y + dy, 2x + dx) of a [2H][2W] image
    p += ((size_t)dy * 2 * W + dx) * (size_t)C * 4;
    strides[0] = (cuuint64_t)2 * C * 4, strides[1] = (cuuint64_t)2 * (2 * W) * C * 4, strides[2] = (cuuint64_t)(2 * H) * (2 * W) * C * 4;
  }
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int C, int NT>
int launch(const Cfg& cfg, const CUtensorMap* maps, const Args& a, cudaStream_t stream) {
  static PcabSmemOnce once;
  PCAB_CUDA(pcab_set_max_smem(k_conv_p16<C, NT>, kMaxSmem, once));
  const int nsm = pcab_sm_count();
  const int grid = cfg.total < nsm ? cfg.total : nsm;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(grid), lc.blockDim = dim3(kThreads), lc.dynamicSmemBytes = cfg.smem, lc.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr, lc.numAttrs = 1;
  PCAB_CUDA(cudaLaunchKernelEx(&lc, k_conv_p16<C, NT>, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7], a));
  return PCAB_OK;
}

int run(const void* src0, int c0, int src0_cstride, const void* src1, int c1, const void* src2, int c2, int temporal_T, const void* weight_f16_packed,
        float weight_scale_inv, const float* bias, const float* bn_scale, const float* bn_shift, int relu, void* out, int n_images,
        int H, int W, int cout_layer, int ntaps, unsigned int* sat_counter, cudaStream_t stream) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    pcab_set_error("pcab_conv_p16: cuTensorMapEncodeTiled unavailable");
    return PCAB_ERR_CUDA;
  }
  const void* srcs[3] = {src0, src1, src2};
  int cs[3] = {c0, c1, c2};
  const int nsrc = src2 ? 3 : (src1 ? 2 : 1);
  const int T = temporal_T > 1 ? temporal_T : 1;
  if (T > 1) PCAB_REQUIRE(nsrc == 3 && src0 == src1 && src1 == src2, "temporal mode takes the same tensor three times");
  PCAB_REQUIRE(src0_cstride == 0 || (src0_cstride % 32 == 0 && src0_cstride >= c0 && T == 1), "src0_cstride: 0 (dense) or a multiple of 32 >= c0");
  int cin_total = 0;
  for (int s = 0; s < nsrc; ++s) {
    PCAB_REQUIRE(cs[s] > 0 && cs[s] % 32 == 0, "source channels must be a multiple of 32");
    PCAB_REQUIRE(((uintptr_t)srcs[s] & 127) == 0, "P16 tensors must be 128 B aligned");
    cin_total += cs[s];
  }
  const int cols = ntaps == 9 ? cout_layer : 4 * cout_layer;  // GEMM output columns
  PCAB_REQUIRE(cout_layer % 32 == 0 && ((uintptr_t)out & 127) == 0, "Cout % 32, 128 B aligned output");
  PCAB_REQUIRE(((uintptr_t)bias & 15) == 0 && ((uintptr_t)bn_scale & 15) == 0 && ((uintptr_t)bn_shift & 15) == 0,
               "bias / BN vectors must be 16B aligned");
  Cfg cfg;
  PCAB_REQUIRE(choose(n_images, H, W, cols, T > 1 ? (3 * cs[0]) / 32 : cin_total / 32, ntaps, &cfg), "unsupported shape");
  const int halo = ntaps == 9 ? 1 : 0;
  CUtensorMap maps[8];
  for (int s = 0; s < 3; ++s) {
    const int ss = s < nsrc ? s : 0;
    if (!encode_act(enc, &maps[s], srcs[ss], cs[ss], (ss == 0 && src0_cstride > 0) ? src0_cstride : cs[ss], W, H, n_images, cfg.Wp,
                    cfg.R + 2 * halo)) {
      pcab_set_error("pcab_conv_p16: cuTensorMapEncodeTiled(A%d) failed", s);
      return PCAB_ERR_CUDA;
    }
  }
  {
    // fp16 [2 * cols rows][Kpad]: K dense in consumption order (source, 32-channel chunk, tap), padded to a multiple of 64;
    // a box = C rows x 64 elements = the (h or l) half of one weight stage
    cuuint64_t K = (cuuint64_t)((ntaps * (cin_total / 32) + 1) / 2) * 64;
    cuuint64_t dims[2] = {K, (cuuint64_t)2 * cols};
    cuuint64_t strides[1] = {K * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)cfg.c};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&maps[3], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)weight_f16_packed, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      pcab_set_error("pcab_conv_p16: cuTensorMapEncodeTiled(B) failed");
      return PCAB_ERR_CUDA;
    }
  }
  for (int p = 0; p < 4; ++p) {
    const bool ok = ntaps == 9 ? encode_act(enc, &maps[4 + p], out, cout_layer, cout_layer, W, H, n_images, cfg.Wt, cfg.R)
                               : encode_act(enc, &maps[4 + p], out, cout_layer, cout_layer, W, H, n_images, cfg.Wt, cfg.R, p >> 1, p & 1);
    if (!ok) {
      pcab_set_error("pcab_conv_p16: cuTensorMapEncodeTiled(out%d) failed", p);
      return PCAB_ERR_CUDA;
    }
  }
  Args a;
  a.nsrc = nsrc;
  for (int s = 0; s < 3; ++s) a.src_c[s] = cs[s];
  a.T = T;
  a.N = n_images, a.H = H, a.W = W, a.Cout = cols;
  a.mt = cfg.mt, a.strip = cfg.strip, a.mtx = cfg.mtx, a.R = cfg.R, a.Wt = cfg.Wt, a.Wp = cfg.Wp;
  a.tiles_x = cfg.tiles_x, a.tiles_y = cfg.tiles_y, a.n_ctile = cfg.n_ctile, a.total_items = cfg.total;
  a.np = cfg.np, a.plane_bytes = (uint32_t)cfg.plane_rows * 128u;
  a.ntaps = ntaps, a.relu = relu, a.cout_real = cout_layer, a.wscale_inv = weight_scale_inv;
  a.resident = cfg.resident, a.wslots = cfg.wslots;
  a.kg_total = 0;
  for (int s = 0; s < 3; ++s) a.kg_src[s] = a.kg_total, a.kg_total += s < nsrc ? (cs[s] / 32) * ntaps : 0;
  a.bias = bias, a.bn_scale = bn_scale, a.bn_shift = bn_shift, a.sat_counter = sat_counter;
  int rc;
  if (ntaps == 9) {
    rc = cfg.c == 128 ? launch<128, 9>(cfg, maps, a, stream)
         : cfg.c == 96 ? launch<96, 9>(cfg, maps, a, stream)
         : cfg.c == 64 ? launch<64, 9>(cfg, maps, a, stream)
                       : launch<32, 9>(cfg, maps, a, stream);
  } else {
    rc = cfg.c == 128 ? launch<128, 1>(cfg, maps, a, stream)
         : cfg.c == 96 ? launch<96, 1>(cfg, maps, a, stream)
         : cfg.c == 64 ? launch<64, 1>(cfg, maps, a, stream)
                       : launch<32, 1>(cfg, maps, a, stream);
  }
  if (rc != PCAB_OK) return rc;
  PCAB_CHECK_LAUNCH("pcab_conv_p16");
  return PCAB_OK;
}

}  // namespace

extern "C" int pcab_conv3x3_p16_supported(int n_sources, int c0, int c1, int c2, int Cout, int H, int W) {
  if (n_sources < 1 || n_sources > 3) return 0;
  int cs[3] = {c0, c1, c2};
  for (int s = 0; s < n_sources; ++s)
    if (cs[s] <= 0 || cs[s] % 32) return 0;
  Cfg k;
  return (H >= 8 && W >= 8 && choose(1, H, W, Cout, 1, 9, &k)) ? 1 : 0;
}

// the tile plan: out[0..12] = columns per item, mt, strip, mtx, R, Wt, tiles_x, tiles_y, column tiles, work items, plane stages,
// dynamic shared memory bytes, weights resident
extern "C" int pcab_conv_p16_plan(int n_images, int H, int W, int Cout, int cin_total, int ntaps, int* out13) {
  Cfg k;
  PCAB_REQUIRE(ntaps == 9 || ntaps == 1, "ntaps is 9 (conv3x3) or 1 (ConvTranspose2x2)");
  PCAB_REQUIRE(choose(n_images, H, W, ntaps == 9 ? Cout : 4 * Cout, cin_total / 32, ntaps, &k), "unsupported shape");
  int v[13] = {k.c, k.mt, k.strip, k.mtx, k.R, k.Wt, k.tiles_x, k.tiles_y, k.n_ctile, k.total, k.np, (int)k.smem, k.resident};
  for (int i = 0; i < 13; ++i) out13[i] = v[i];
  return PCAB_OK;
}

extern "C" int pcab_conv3x3_p16(const void* src0, int c0, int src0_cstride, const void* src1, int c1, const void* src2, int c2, int temporal_T,
                                const void* weight_f16_packed, float weight_scale_inv, const float* bias, const float* bn_scale,
                                const float* bn_shift, int relu, void* out, int n_images, int H, int W, int Cout,
                                unsigned int* sat_counter, cudaStream_t stream) {
  return run(src0, c0, src0_cstride, src1, c1, src2, c2, temporal_T, weight_f16_packed, weight_scale_inv, bias, bn_scale, bn_shift, relu,
             out, n_images, H, W, Cout, 9, sat_counter, stream);
}

// Conv3d 3x3x3, 32 -> 32 channels, temporal taps fused into the MMA N dimension (k_conv3d_p16).  weight: fp16 [2 (h, l)][96 rows =
// (kt = 2, 1, 0) x 32 output channels][320] (nine taps x 32 input channels, K dense, padded to 320; tc_pack.pack_conv3d_fused_p16).
extern "C" int pcab_conv3d_p16(const void* src, int T, const void* weight_f16_packed, float weight_scale_inv, const float* bias, int relu,
                               void* out, int n_images, int H, int W, unsigned int* sat_counter, cudaStream_t stream) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    pcab_set_error("pcab_conv3d_p16: cuTensorMapEncodeTiled unavailable");
    return PCAB_ERR_CUDA;
  }
  PCAB_REQUIRE(T >= 2 && n_images % T == 0 && H >= 8 && W >= 8, "n_images = B * T, T >= 2, maps of at least 8 x 8");
  PCAB_REQUIRE(((uintptr_t)src & 127) == 0 && ((uintptr_t)out & 127) == 0 && ((uintptr_t)bias & 15) == 0, "alignment");
  CUtensorMap ma, mb, mo;
  if (!encode_act(enc, &ma, src, 32, 32, W, H, n_images, 10, 18) || !encode_act(enc, &mo, out, 32, 32, W, H, n_images, 8, 16)) {
    pcab_set_error("pcab_conv3d_p16: cuTensorMapEncodeTiled(activations) failed");
    return PCAB_ERR_CUDA;
  }
  {
    cuuint64_t dims[2] = {320, 192};
    cuuint64_t strides[1] = {640};
    cuuint32_t box[2] = {64u, 192u};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)weight_f16_packed, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      pcab_set_error("pcab_conv3d_p16: cuTensorMapEncodeTiled(B) failed");
      return PCAB_ERR_CUDA;
    }
  }
  Args3d a;
  a.B = n_images / T, a.T = T, a.H = H, a.W = W;
  a.tiles_x = cdiv(W, 8), a.tiles_y = cdiv(H, 16), a.total_items = a.B * a.tiles_x * a.tiles_y;
  a.np = 3, a.plane_bytes = 184u * 128u;
  a.relu = relu, a.wscale_inv = weight_scale_inv, a.bias = bias, a.sat_counter = sat_counter;
  const size_t smem = 1024 + (size_t)a.np * a.plane_bytes + (size_t)kW3dStages * kW3dStage + 16384 + 256;
  static PcabSmemOnce once;
  PCAB_CUDA(pcab_set_max_smem(k_conv3d_p16, (int)smem, once));
  const int nsm = pcab_sm_count();
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(a.total_items < nsm ? a.total_items : nsm), lc.blockDim = dim3(kThreads), lc.dynamicSmemBytes = smem, lc.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr, lc.numAttrs = 1;
  PCAB_CUDA(cudaLaunchKernelEx(&lc, k_conv3d_p16, ma, mb, mo, a));
  PCAB_CHECK_LAUNCH("pcab_conv3d_p16");
  return PCAB_OK;
}

extern "C" int pcab_convT2x2_p16(const void* in, int Cin, const void* weight_f16_packed, float weight_scale_inv, const float* bias,
                                 void* out /* [n, 2H, 2W, Cout] P16 */, int n_images, int H, int W, int Cout,
                                 unsigned int* sat_counter, cudaStream_t stream) {
  return run(in, Cin, 0, nullptr, 0, nullptr, 0, 1, weight_f16_packed, weight_scale_inv, bias, nullptr, nullptr, 0, out, n_images, H, W,
             Cout, 1, sat_counter, stream);
}
