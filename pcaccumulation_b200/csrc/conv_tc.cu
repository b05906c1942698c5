// conv3x3 (pad 1) on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with FP32 accumulators in TMEM,
// operands staged by TMA (cp.async.bulk.tensor, 128B swizzle), 3xTF32 error compensation so the result keeps
// FP32 accuracy (needed for bit-exact segmentation labels downstream).
//
// Same contract as pcab_conv3x3_f32 (multi-source accumulate = concat / temporal 3x3x3, bias, BN(eval), ReLU).
// Replaces the cuDNN calls behind models/unet.py:11-20,57-62,88-97 and models/stpn.py:13-22.
//
// Formulation: implicit GEMM with pixels as the MMA M dimension and output channels as N.  One CTA owns an
// output tile of R rows x Wt columns of one image.  For each 32-channel slice of the input it TMA-loads ONE halo
// plane  [(R+2) x (Wt+2) pixels][32 ch]  (zero-filled outside the image = the conv padding) into shared memory,
// 128 bytes per pixel = one swizzle row.  Output pixels are indexed in the FLATTENED padded grid m = r*(Wt+2)+x,
// so the A operand of tap (ky,kx) is the same plane viewed from row  m + ky*(Wt+2) + kx : nine shifted views of
// one staged plane instead of nine loads (the two extra columns per row compute garbage that is never stored).
// 3xTF32:  a = a_hi + a_lo, w = w_hi + w_lo; a_hi = the 19 bits the tensor core reads of a, w_hi = w rounded to tf32;
//          acc += a_hi*[w_hi|w_lo] (one N=2*Cout MMA) + a_lo*w_hi   (a_lo*w_lo ~ 2^-22 is dropped).
// w_hi / w_lo are pre-split on the host; a_lo is produced in shared memory by the epilogue warps right after the
// plane lands.  Roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = a_lo split and
// epilogue (tcgen05.ld -> bias/BN/ReLU -> NHWC global stores).
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "pcab200.h"

namespace {

constexpr int kThreads = 192;
constexpr int kMaxSmem = 227 * 1024;

struct TcArgs {
  int nsrc;
  int src_c[3];
  int T;  // temporal frames per scene (1 = plain)
  int N, H, W, Cout;
  int cout_t;   // output channels per CTA (MMA N)
  int mt;       // M tiles (of 128 flattened pixels) per CTA
  int R, Wt, Wp;
  int plane_rows;  // allocated rows (128 B each) per plane
  int tiles_x, tiles_y;
  int relu;
  int out_cstride, out_coff;
  int base_offset_mode;
  int nst;  // weight pipeline stages (2..4)
  const float* bias;
  const float* bn_scale;
  const float* bn_shift;
  float* out;
};

using namespace pcab_tc;

__global__ void __launch_bounds__(kThreads, 2)
k_conv3x3_tc(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
             const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b, TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [plane_hi][plane_lo][b stage 0: hi, lo][b stage 1: hi, lo][barriers]
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t plane_bytes = (uint32_t)a.plane_rows * 128u;
  const uint32_t b_bytes = (uint32_t)a.cout_t * 128u;
  uint8_t* plane_hi = base;
  uint8_t* plane_lo = base + plane_bytes;
  uint8_t* b_stage = base + 2 * plane_bytes;  // stage s: hi at s*2*b_bytes, lo at +b_bytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_stage + 2 * a.nst * b_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const uint32_t bar_a_full = smem_u32(bars + 0), bar_lo_done = smem_u32(bars + 1), bar_a_free = smem_u32(bars + 2);
  const uint32_t bar_acc = smem_u32(bars + 3), bar_b_full0 = smem_u32(bars + 4), bar_b_empty0 = smem_u32(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y, n = tile / (a.tiles_x * a.tiles_y);
  const int x0 = tx * a.Wt, y0 = ty * a.R;
  const int co0 = blockIdx.y * a.cout_t;
  const int tframe = a.T > 1 ? n % a.T : 0;

  // chunk list: (source, channel offset), identical for every role
  int nchunks = 0;
  for (int s = 0; s < a.nsrc; ++s) {
    bool valid = a.T <= 1 || (tframe + s - 1 >= 0 && tframe + s - 1 < a.T);
    if (valid) nchunks += a.src_c[s] / 32;
  }

  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * a.mt * a.cout_t) tmem_cols <<= 1;  // main + correction accumulators

  if (threadIdx.x == 0) {
    mbar_init(bar_a_full, 1);
    mbar_init(bar_lo_done, 128);
    mbar_init(bar_a_free, 1);
    for (int i = 0; i < 4; ++i) mbar_init(bar_b_full0 + 8 * i, 1), mbar_init(bar_b_empty0 + 8 * i, 1);
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int ci = 0, bs = 0, bph = 0;
      int kbase_src = 0;
      for (int s = 0; s < a.nsrc; ++s) {
        const int Cs = a.src_c[s];
        bool valid = a.T <= 1 || (tframe + s - 1 >= 0 && tframe + s - 1 < a.T);
        if (valid) {
          const CUtensorMap* am = a.T > 1 ? &map_a0 : (s == 0 ? &map_a0 : (s == 1 ? &map_a1 : &map_a2));
          const int nsrc_img = a.T > 1 ? n + s - 1 : n;
          for (int c0 = 0; c0 < Cs; c0 += 32, ++ci) {
            if (ci > 0) mbar_wait(bar_a_free, (ci - 1) & 1);
            mbar_expect_tx(bar_a_full, (uint32_t)(a.R + 2) * a.Wp * 128u);
            tma_load_4d(am, smem_u32(plane_hi), bar_a_full, c0, x0 - 1, y0 - 1, nsrc_img);
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(bar_b_empty0 + 8 * bs, bph ^ 1);
              mbar_expect_tx(bar_b_full0 + 8 * bs, 2 * b_bytes);
              const int k0 = kbase_src + tap * Cs + c0;
              tma_load_2d(&map_b, smem_u32(b_stage + (2 * bs) * b_bytes), bar_b_full0 + 8 * bs, k0, co0);
              tma_load_2d(&map_b, smem_u32(b_stage + (2 * bs + 1) * b_bytes), bar_b_full0 + 8 * bs, k0, a.Cout + co0);
              if (++bs == a.nst) bs = 0, bph ^= 1;
            }
          }
        }
        kbase_src += 9 * Cs;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this loop with warp-uniform values (so descriptors live in uniform registers); only the
    // tcgen05 instructions themselves are issued by one elected lane.  A descriptor differs from tile to tile only in
    // its 14-bit start-address field, so it is built once and advanced with a 32-bit add per MMA.
    {
      // instruction descriptors: D=f32, A=B=tf32, K-major both, M = 128; N = 2*cout_t for the fused main MMA
      // (weights staged as [w_hi rows | w_lo rows] = one 2*cout_t-row operand), N = cout_t for the a_lo correction
      const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24);
      const uint32_t idesc2 = idesc_base | ((uint32_t)((2 * a.cout_t) >> 3) << 17);
      const uint32_t idesc1 = idesc_base | ((uint32_t)(a.cout_t >> 3) << 17);
      const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);  // SBO, version, SWIZZLE_128B (upper word)
      const uint32_t desc_lo0 = 1u << 16;                                          // LBO field (lower word)
      const uint32_t ah16 = (smem_u32(plane_hi) & 0x3FFFF) >> 4, al16 = (smem_u32(plane_lo) & 0x3FFFF) >> 4;
      const uint32_t b16 = (smem_u32(b_stage) & 0x3FFFF) >> 4, bstep16 = b_bytes >> 4;
      int bs = 0, bph = 0;
      for (int ci = 0; ci < nchunks; ++ci) {
        mbar_wait(bar_lo_done, ci & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(bar_b_full0 + 8 * bs, bph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t shift16 = (uint32_t)((tap / 3) * a.Wp + (tap % 3)) * 8u;  // rows * 128 B / 16
          const uint32_t bh16 = b16 + (uint32_t)(2 * bs) * bstep16;
          for (int mt = 0; mt < a.mt; ++mt) {
            // TMEM columns of tile mt: [0, c) = sum a_hi*w_hi (main), [c, 2c) = sum of the ~2^-11 smaller corrections
            // a_hi*w_lo + a_lo*w_hi.  One N=2c MMA  a_hi x [w_hi | w_lo]  fills both halves reading a_hi ONCE (the kernel is
            // bound by shared-memory operand bandwidth), one N=c MMA  a_lo x w_hi  adds into the correction half.  The
            // tensor core truncates when it adds into an accumulator, so keeping the small terms out of the main half also
            // cuts its number of (biased) roundings from 3K/8 to K/8.
            const uint32_t tmem_d = tmem_base + (uint32_t)(mt * 2 * a.cout_t);
            const uint32_t tmem_c = tmem_d + (uint32_t)a.cout_t;
            const uint32_t arow16 = (uint32_t)(mt * 128) * 8u + shift16;
            const uint32_t first = (ci == 0 && tap == 0) ? 0u : 1u;
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t dah = ((uint64_t)desc_hi << 32) | (desc_lo0 | (ah16 + arow16 + 2u * kk));
                const uint64_t dal = ((uint64_t)desc_hi << 32) | (desc_lo0 | (al16 + arow16 + 2u * kk));
                const uint64_t dbh = ((uint64_t)desc_hi << 32) | (desc_lo0 | (bh16 + 2u * kk));
                umma_tf32(tmem_d, dah, dbh, idesc2, (kk == 0) ? first : 1u);
                umma_tf32(tmem_c, dal, dbh, idesc1, 1u);
              }
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit(bar_b_empty0 + 8 * bs);  // frees this weight stage once the MMAs above retire
          __syncwarp();
          if (++bs == a.nst) bs = 0, bph ^= 1;
        }
        if (elect_one()) umma_commit(bar_a_free);  // planes may be overwritten
        __syncwarp();
      }
      if (elect_one()) umma_commit(bar_acc);
      __syncwarp();
    }
  } else {
    // ===================== a_lo split, then epilogue =====================
    const int et = threadIdx.x - 64;  // 0..127
    const int n_f4 = (a.R + 2) * a.Wp * 8;  // float4 per plane (the rows the TMA box wrote)
    for (int ci = 0; ci < nchunks; ++ci) {
      mbar_wait(bar_a_full, ci & 1);
      const float4* hi = reinterpret_cast<const float4*>(plane_hi);
      float4* lo = reinterpret_cast<float4*>(plane_lo);
      // a_hi is what the tensor core reads of the raw FP32 value (it ignores the low 13 mantissa bits), so the hi plane
      // needs no rewrite; a_lo = a - trunc_tf32(a) is exact in FP32.  Four independent 128-bit loads in flight per thread.
      auto low = [](float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); };
      int i = et;
      for (; i + 384 < n_f4; i += 512) {
        float4 v0 = hi[i], v1 = hi[i + 128], v2 = hi[i + 256], v3 = hi[i + 384];
        lo[i] = make_float4(low(v0.x), low(v0.y), low(v0.z), low(v0.w));
        lo[i + 128] = make_float4(low(v1.x), low(v1.y), low(v1.z), low(v1.w));
        lo[i + 256] = make_float4(low(v2.x), low(v2.y), low(v2.z), low(v2.w));
        lo[i + 384] = make_float4(low(v3.x), low(v3.y), low(v3.z), low(v3.w));
      }
      for (; i < n_f4; i += 128) {
        float4 v = hi[i];
        lo[i] = make_float4(low(v.x), low(v.y), low(v.z), low(v.w));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA (async proxy)
      mbar_arrive(bar_lo_done);
    }
    mbar_wait(bar_acc, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int quarter = warp & 3;  // TMEM lane quarter this warp may read
    for (int mt = 0; mt < a.mt; ++mt) {
      const int m = mt * 128 + quarter * 32 + lane;
      const int r = m / a.Wp, xc = m % a.Wp;
      const int gy = y0 + r, gx = x0 + xc;
      const bool valid = r < a.R && xc < a.Wt && gy < a.H && gx < a.W;
      float* orow = a.out + (((size_t)n * a.H + gy) * a.W + gx) * a.out_cstride + a.out_coff + co0;
      for (int c32 = 0; c32 < a.cout_t / 32; ++c32) {
        uint32_t v[32], vc[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * 2 * a.cout_t + c32 * 32), v);
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * 2 * a.cout_t + a.cout_t + c32 * 32), vc);
        tmem_ld_wait(v, vc);
        if (valid) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              int co = co0 + c32 * 32 + 4 * q + u;
              float f = (__uint_as_float(v[4 * q + u]) + __uint_as_float(vc[4 * q + u])) + a.bias[co];
              if (a.bn_scale) f = fmaf(f, a.bn_scale[co], a.bn_shift[co]);
              o[u] = a.relu ? fmaxf(f, 0.f) : f;
            }
            *reinterpret_cast<float4*>(orow + c32 * 32 + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}


// =====================================================================================================================
// v2: persistent, fully pipelined kernel (one CTA per SM, 16 warps).
//   * each CTA walks a static list of (output tile, cout tile) work items;
//   * hi planes are double buffered (TMA of chunk g+1 overlaps the MMAs of chunk g), so are the a_lo planes written by
//     the splitter warps, the weight tap stages run through a 3-deep ring fed by their own producer warp;
//   * the TMEM accumulators are per 32-channel CHUNK and double buffered: after the 36 K-steps of a chunk the drain
//     warps pull the partial sums into FP32 registers (round-to-nearest adds) while the tensor core already works on the
//     next chunk.  The tensor core drops the addend bits below the accumulator's ulp on every accumulation, so long
//     in-TMEM chains (K up to 4608 here) were the dominant error of v1; with chunk-wise draining the chain is 36 steps
//     whatever the layer.
//   * tiles: "strip" = 8 pixels wide x 16*mt rows, the 8-row core-matrix groups of the UMMA descriptor are image rows
//     (stride-byte-offset = plane pitch), every M row is a real output pixel; "flat" = flattened padded grid as in v1
//     (used for narrow maps).
// Warp roles (the warp scheduler favours high warp ids, so the latency-critical roles sit at the top):
//   0-7 drain + epilogue | 8-11 a_lo split | 12 plane TMA | 13 weight TMA | 14,15 MMA issue (one M tile each; 14 owns TMEM)
// Warp roles: 0 plane TMA | 1 weight TMA | 2,3 MMA issue (one M tile each; 2 also owns TMEM) | 4-11 drain + epilogue | 12-15 a_lo split
// =====================================================================================================================
namespace v2 {

constexpr int kThreads2 = 512;
constexpr int kPlaneRows = 344;                 // rows of 128 B per plane buffer
constexpr uint32_t kPlaneBytes = kPlaneRows * 128;  // 44032 = 43 KiB (keeps every buffer 1024-aligned)
constexpr int kWStages = 3;

struct Args {
  int nsrc;
  int src_c[3];
  int T;
  int N, H, W, Cout;
  int c;       // output channels per work item (MMA N = 2c main, c correction)
  int mt;      // M tiles of 128 rows per work item (1 or 2)
  int strip;   // 1: strip tiles (8 px wide groups), 0: flattened padded grid
  int mtx;     // strip mode: M tiles side by side (1 or 2); tile mi sits at column block mi % mtx, row block mi / mtx
  int R, Wt, Wp;
  int tiles_x, tiles_y, n_ctile, total_items;
  int relu;
  int out_cstride, out_coff;
  int store32;  // output rows are 32 B aligned: 256-bit stores
  float wscale_inv;  // fp16-pair operands: the weights were scaled by a power of two on the host; 1 for the tf32 operands
  const float* bias;
  const float* bn_scale;
  const float* bn_shift;
  float* out;
  long long* stats;  // debug: per-CTA wait-cycle counters (null = off)
  int dbg;           // debug: bit 0 = splitter does no work, bit 1 = no epilogue stores, bit 2 = producer loads the weights of the first item only
};

// barrier wait that adds the waited cycles to a register counter when the debug counters are on
__device__ __forceinline__ void timed_wait(uint32_t bar, uint32_t parity, bool on, long long& acc) {
  if (on) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
  } else {
    mbar_wait(bar, parity);
  }
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
struct Item {
  int n, x0, y0, co0, tframe;
};
__device__ __forceinline__ Item decode_item(const Args& a, int it) {
  Item r;
  const int ct = it % a.n_ctile, sp = it / a.n_ctile;
  const int tx = sp % a.tiles_x, ty = (sp / a.tiles_x) % a.tiles_y;
  r.n = sp / (a.tiles_x * a.tiles_y);
  r.x0 = tx * a.Wt, r.y0 = ty * a.R, r.co0 = ct * a.c;
  r.tframe = a.T > 1 ? r.n % a.T : 0;
  return r;
}
__device__ __forceinline__ bool src_valid(const Args& a, int s, int tframe) {
  return a.T <= 1 || (tframe + s - 1 >= 0 && tframe + s - 1 < a.T);
}

// F16 = false: 3xTF32 (A = the raw FP32 plane + a residual plane, kind::tf32, K = 8 per MMA).
// F16 = true : the same three products with both operands split into fp16 PAIRS (x = h + l, h = fp16(x), l = fp16(x - h); 22
//   significant bits, products exact in FP32): the splitter packs [32 ch h | 32 ch l] into ONE 128-byte row per pixel, the
//   MMAs run kind::f16 (K = 16 per MMA: half as many MMAs for the same bytes per MMA, i.e. half the tensor-pipe and
//   shared-memory-operand time per chunk).  Weights arrive pre-split as fp16 rows of 64 (32 used) per 32-channel group.
template <int C, bool F16>
__global__ void __launch_bounds__(kThreads2, 1)
k_conv3x3_tc2(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
              const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b, Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t hi0 = sbase, lo0 = sbase + 2 * kPlaneBytes, w0 = sbase + 4 * kPlaneBytes;
  constexpr uint32_t kWBytes = 2u * C * 128u;  // one tap stage: [w_hi rows | w_lo rows]
  const uint32_t bars = w0 + kWStages * kWBytes;
  // barrier slots (8 B each)
  const uint32_t bar_hi_full = bars, bar_plane_free = bars + 16, bar_lo_full = bars + 32, bar_w_full = bars + 48,
                 bar_w_free = bars + 72, bar_acc_full = bars + 96, bar_acc_empty = bars + 112, tmem_slot = bars + 128,
                 bar_raw_free = bars + 136;  // fp16 pairs only: the TMA plane has been read by the split warps

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // warp-uniform for the compiler
  const uint32_t tmem_cols_needed = 4u * a.mt * C;  // 2 stages x mt tiles x [main c | corr c]
  uint32_t tmem_cols = 32;
  while (tmem_cols < tmem_cols_needed) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_hi_full + 8 * i, 1);
      mbar_init(bar_plane_free + 8 * i, a.mt);  // one tcgen05.commit per issuing warp
      mbar_init(bar_lo_full + 8 * i, 128);
      mbar_init(bar_raw_free + 8 * i, 128);
      mbar_init(bar_acc_full + 8 * i, a.mt);
      mbar_init(bar_acc_empty + 8 * i, 256);
    }
    for (int i = 0; i < kWStages; ++i) mbar_init(bar_w_full + 8 * i, 1), mbar_init(bar_w_free + 8 * i, a.mt);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 14) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 12 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  const uint32_t box_bytes = (uint32_t)(a.R + 2) * a.Wp * 128u;

  if (warp == 12) {
    // ===================== plane producer =====================
    if (lane == 0) {
      int g = 0;
      for (int it = blockIdx.x; it < a.total_items; it += gridDim.x) {
        const Item t = decode_item(a, it);
        for (int s = 0; s < a.nsrc; ++s) {
          if (!src_valid(a, s, t.tframe)) continue;
          const CUtensorMap* am = a.T > 1 ? &map_a0 : (s == 0 ? &map_a0 : (s == 1 ? &map_a1 : &map_a2));
          const int img = a.T > 1 ? t.n + s - 1 : t.n;
          for (int c0 = 0; c0 < a.src_c[s]; c0 += 32, ++g) {
            const int st = g & 1;
            if (a.dbg & 8) continue;
            // tf32: the MMAs read this buffer (a_hi), it is free when they are done.  fp16 pairs: only the split warps read it,
            // so the next plane can land while the MMAs of the previous chunk still run on the packed buffer.
            if (g >= 2) mbar_wait((F16 ? bar_raw_free : bar_plane_free) + 8 * st, ((g >> 1) - 1) & 1);
            mbar_expect_tx(bar_hi_full + 8 * st, box_bytes);
            tma_load_4d(am, hi0 + st * kPlaneBytes, bar_hi_full + 8 * st, c0, t.x0 - 1, t.y0 - 1, img);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 13) {
    // ===================== weight producer =====================
    if (lane == 0) {
      int wg = 0;
      for (int it = blockIdx.x; it < a.total_items; it += gridDim.x) {
        const Item t = decode_item(a, it);
        int kbase_src = 0;
        for (int s = 0; s < a.nsrc; ++s) {
          const int Cs = a.src_c[s];
          if (src_valid(a, s, t.tframe)) {
            for (int c0 = 0; c0 < Cs; c0 += 32) {
              for (int tap = 0; tap < 9; ++tap, ++wg) {
                const int ws = wg % kWStages;
                if (a.dbg & 4) continue;
                if (wg >= kWStages) mbar_wait(bar_w_free + 8 * ws, ((wg / kWStages) - 1) & 1);
                mbar_expect_tx(bar_w_full + 8 * ws, kWBytes);
                const int k0 = (kbase_src + tap * Cs + c0) * (F16 ? 2 : 1);  // fp16 rows: 64 elements per 32-channel group
                tma_load_2d(&map_b, w0 + ws * kWBytes, bar_w_full + 8 * ws, k0, t.co0);
                tma_load_2d(&map_b, w0 + ws * kWBytes + C * 128u, bar_w_full + 8 * ws, k0, a.Cout + t.co0);
              }
            }
          }
          kbase_src += 9 * Cs;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 14) {
    // ===================== MMA issuers =====================
    // Two issuing warps: warp 14 owns M tile 0, warp 15 owns M tile 1 (disjoint accumulators, so no ordering between
    // them is needed).  The tensor-core queue is shallow: every cycle an issuer spends on barrier waits, fences and
    // descriptor arithmetic is a cycle its MMAs are not queued, and a second issuer fills those gaps.  One elected
    // lane per warp runs the whole chunk (9 taps, fully unrolled: tap offsets, weight stage and parity are literals).
    const int mi = warp - 14;
    if (mi < a.mt) {
      const uint32_t idesc_base = F16 ? ((1u << 4) | ((128u >> 4) << 24))  // D = F32, A = B = F16
                                      : ((1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24));
      const uint32_t idesc2 = idesc_base | ((uint32_t)((2 * C) >> 3) << 17);
      const uint32_t idesc1 = idesc_base | ((uint32_t)(C >> 3) << 17);
      const uint32_t sbo_a = a.strip ? (uint32_t)a.Wp * 8u : 64u;  // stride between 8-row groups, in 16 B units
      const uint64_t desc_hi_a = (uint64_t)(sbo_a | (1u << 14) | (2u << 29)) << 32;
      const uint64_t desc_hi_b = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;
      const uint32_t lbo = 1u << 16;
      // this warp's M tile in the plane (16 B units): strip tiles are 8 px x 16 rows blocks, flat tiles 128 consecutive rows
      const uint32_t tile_off16 = a.strip ? ((uint32_t)(mi % a.mtx) * 8u + (uint32_t)(mi / a.mtx) * 16u * (uint32_t)a.Wp) * 8u : (uint32_t)mi * 1024u;
      const uint32_t wp8 = (uint32_t)a.Wp * 8u;
      const uint32_t ah_base = lbo | ((((hi0)&0x3FFFF) >> 4) + tile_off16), al_base = lbo | ((((lo0)&0x3FFFF) >> 4) + tile_off16);
      const uint32_t b_base = lbo | ((w0 & 0x3FFFF) >> 4);
      const bool st_on = a.stats != nullptr && mi == 0;
      long long c_lo = 0, c_acc = 0, c_issue = 0;
      int g = 0;
      const long long tm0 = a.stats ? clock64() : 0;
      for (int it = blockIdx.x; it < a.total_items; it += gridDim.x) {
        const Item t = decode_item(a, it);
        int nchunks = 0;
        for (int s = 0; s < a.nsrc; ++s)
          if (src_valid(a, s, t.tframe)) nchunks += a.src_c[s] / 32;
        for (int ci = 0; ci < nchunks; ++ci, ++g) {
          const int st = g & 1;
          timed_wait(bar_lo_full + 8 * st, (g >> 1) & 1, st_on, c_lo);
          if (g >= 2) timed_wait(bar_acc_empty + 8 * st, ((g >> 1) - 1) & 1, st_on, c_acc);
          if (elect_one()) {
            const long long ti0 = st_on ? clock64() : 0;
            const uint32_t ah = ah_base + (uint32_t)st * (kPlaneBytes >> 4), al = al_base + (uint32_t)st * (kPlaneBytes >> 4);
            const uint32_t tmem_d = tmem_base + (uint32_t)(st * a.mt * 2 * C + mi * 2 * C);
            const uint32_t wpar = (uint32_t)g;  // weight stage `tap % 3` is in its (3g + tap/3)-th use
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              constexpr int kStagesLit = 3;
              static_assert(kWStages == kStagesLit, "the unrolled tap loop assumes a 3-deep weight ring");
              const int ws = tap % 3;
              if (!(a.dbg & 4)) mbar_wait(bar_w_full + 8 * ws, (wpar + (uint32_t)(tap / 3)) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t shift16 = (uint32_t)(tap / 3) * wp8 + (uint32_t)(tap % 3) * 8u;
              const uint32_t b16 = b_base + (uint32_t)ws * (kWBytes >> 4);
              if (F16) {
                // packed plane (the "lo" buffers): bytes 0-63 of a row = h halves (k-steps 0, 1), bytes 64-127 = l halves
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                  const uint64_t dah = desc_hi_a | (al + shift16 + 2u * kk);
                  const uint64_t dal = desc_hi_a | (al + shift16 + 4u + 2u * kk);
                  const uint64_t db = desc_hi_b | (b16 + 2u * kk);
                  umma_f16(tmem_d, dah, db, idesc2, (tap == 0 && kk == 0) ? 0u : 1u);  // [a_h*w_h | a_h*w_l]
                  umma_f16(tmem_d + (uint32_t)C, dal, db, idesc1, 1u);                // += a_l*w_h into the second half
                }
              } else {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  const uint64_t dah = desc_hi_a | (ah + shift16 + 2u * kk);
                  const uint64_t dal = desc_hi_a | (al + shift16 + 2u * kk);
                  const uint64_t db = desc_hi_b | (b16 + 2u * kk);
                  umma_tf32(tmem_d, dah, db, idesc2, (tap == 0 && kk == 0) ? 0u : 1u);  // [a_hi*w_hi | a_hi*w_lo]
                  umma_tf32(tmem_d + (uint32_t)C, dal, db, idesc1, 1u);                // += a_lo*w_hi into the second half
                }
              }
              umma_commit(bar_w_free + 8 * ws);
            }
            umma_commit(bar_plane_free + 8 * st);
            umma_commit(bar_acc_full + 8 * st);
            if (st_on) c_issue += clock64() - ti0;
          }
          __syncwarp();
        }
      }
      if (st_on) {
        // the elected lane may differ from call to call in principle; reduce the per-lane issue counters over the warp
        for (int o = 16; o > 0; o >>= 1) c_issue += __shfl_xor_sync(0xffffffffu, c_issue, o);
        if (lane == 0) {
          long long* sp = a.stats + blockIdx.x * 16;
          sp[0] += c_lo, sp[1] += c_acc, sp[10] += c_issue;
          sp[7] += clock64() - tm0;
          sp[8] += g;
        }
      }
    }
  } else if (warp >= 8) {
    // ===================== a_lo splitter =====================
    const int et = threadIdx.x - 256;  // 0..127
    const int n_f4 = (a.R + 2) * a.Wp * 8;
    const bool st_on = a.stats != nullptr && threadIdx.x == 256;
    long long c_free = 0, c_hi = 0, c_work = 0;
    int g = 0;
    for (int it = blockIdx.x; it < a.total_items; it += gridDim.x) {
      const Item t = decode_item(a, it);
      int nchunks = 0;
      for (int s = 0; s < a.nsrc; ++s)
        if (src_valid(a, s, t.tframe)) nchunks += a.src_c[s] / 32;
      for (int ci = 0; ci < nchunks; ++ci, ++g) {
        const int st = g & 1;
        if (g >= 2) timed_wait(bar_plane_free + 8 * st, ((g >> 1) - 1) & 1, st_on, c_free);  // a_lo[st] free
        if (!(a.dbg & 8)) timed_wait(bar_hi_full + 8 * st, (g >> 1) & 1, st_on, c_hi);
        const long long ts0 = st_on ? clock64() : 0;
        const uint32_t hi = hi0 + st * kPlaneBytes, lo = lo0 + st * kPlaneBytes;
        int i = (a.dbg & 1) ? n_f4 : et;
        if (F16) {
          // 16-byte chunk i of the TMA plane = row i>>3, physical chunk i&7 = channels 4q..4q+3 with q = (i&7) ^ (row&7)
          // (128B swizzle).  Its h halves go to logical chunk q>>1 (byte (q&1)*8), its l halves to logical chunk 4 + (q>>1).
          auto pack = [&](int ii, const float4& v) {
            const uint32_t row = (uint32_t)ii >> 3, sw = row & 7u, q = ((uint32_t)ii & 7u) ^ sw;
            // saturating conversions: beyond +-65504 h (and then l) clamp instead of becoming inf - inf
            uint32_t uh01, uh23, ul01, ul23;
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(uh01) : "f"(v.y), "f"(v.x));
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(uh23) : "f"(v.w), "f"(v.z));
            const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&uh01));
            const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&uh23));
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(ul01) : "f"(v.y - f01.y), "f"(v.x - f01.x));
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(ul23) : "f"(v.w - f23.y), "f"(v.z - f23.x));
            const uint32_t base = lo + row * 128u + (q & 1u) * 8u;
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(base + (((q >> 1) ^ sw) << 4)),
                         "r"(uh01), "r"(uh23) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(base + (((4u + (q >> 1)) ^ sw) << 4)),
                         "r"(ul01), "r"(ul23) : "memory");
          };
          for (; i + 7 * 128 < n_f4; i += 1024) {  // eight independent 128-bit loads in flight per thread
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = lds128(hi + 16u * (i + 128 * u));
#pragma unroll
            for (int u = 0; u < 8; ++u) pack(i + 128 * u, v[u]);
          }
          for (; i < n_f4; i += 128) pack(i, lds128(hi + 16u * i));
        }
        for (; i + 7 * 128 < n_f4; i += 1024) {  // eight independent 128-bit loads in flight per thread
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = lds128(hi + 16u * (i + 128 * u));
#pragma unroll
          for (int u = 0; u < 8; ++u)
            sts128(lo + 16u * (i + 128 * u), make_float4(split_lo(v[u].x), split_lo(v[u].y), split_lo(v[u].z), split_lo(v[u].w)));
        }
        for (; i < n_f4; i += 128) {
          float4 v = lds128(hi + 16u * i);
          sts128(lo + 16u * i, make_float4(split_lo(v.x), split_lo(v.y), split_lo(v.z), split_lo(v.w)));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(bar_lo_full + 8 * st);
        if (F16) mbar_arrive(bar_raw_free + 8 * st);
        if (st_on) c_work += clock64() - ts0;
      }
    }
    if (st_on) {
      long long* sp = a.stats + blockIdx.x * 16;
      sp[3] += c_free, sp[4] += c_hi, sp[5] += c_work;
    }
  } else {
    // ===================== drain + epilogue =====================
    // thread = one TMEM lane (pixel row of an M tile) x one half of the tile's output channels
    constexpr int CH = C / 2;
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    float acc[2][CH];
    const bool st_on = a.stats != nullptr && threadIdx.x == 0;
    long long c_full = 0, c_epi = 0;
    int g = 0;
    for (int it = blockIdx.x; it < a.total_items; it += gridDim.x) {
      const Item t = decode_item(a, it);
      int nchunks = 0;
      for (int s = 0; s < a.nsrc; ++s)
        if (src_valid(a, s, t.tframe)) nchunks += a.src_c[s] / 32;
      for (int ci = 0; ci < nchunks; ++ci, ++g) {
        const int st = g & 1;
        timed_wait(bar_acc_full + 8 * st, (g >> 1) & 1, st_on, c_full);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tm_stage = tmem_base + lane_addr + (uint32_t)(st * a.mt * 2 * C);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          if (mi < a.mt) {
#pragma unroll
            for (int b16 = 0; b16 < CH / 16; ++b16) {
              uint32_t vm[16], vc[16];
              const uint32_t col = (uint32_t)(mi * 2 * C + half * CH + b16 * 16);
              tmem_ld16(tm_stage + col, vm);
              tmem_ld16(tm_stage + col + (uint32_t)C, vc);
              tmem_ld_wait16(vm, vc);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float v = __uint_as_float(vm[j]) + __uint_as_float(vc[j]);
                acc[mi][b16 * 16 + j] = ci == 0 ? v : acc[mi][b16 * 16 + j] + v;
              }
              if (a.dbg & 16) break;
            }
          }
          if (a.dbg & 16) break;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(bar_acc_empty + 8 * st);
      }
      // epilogue of this work item
      const long long te0 = st_on ? clock64() : 0;
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        if (mi < a.mt) {
          const int m = quarter * 32 + lane;
          int r, xc;
          bool valid;
          if (a.strip) {
            r = (mi / a.mtx) * 16 + (m >> 3), xc = (mi % a.mtx) * 8 + (m & 7);
            valid = true;
          } else {
            const int mm = mi * 128 + m;
            r = mm / a.Wp, xc = mm % a.Wp;
            valid = r < a.R && xc < a.Wt;
          }
          const int gy = t.y0 + r, gx = t.x0 + xc;
          valid = valid && gy < a.H && gx < a.W && !(a.dbg & 2);
          if (valid) {
            const int cbase = t.co0 + half * CH;
            float* orow = a.out + (((size_t)t.n * a.H + gy) * a.W + gx) * a.out_cstride + a.out_coff + cbase;
            // 32 B per store instruction and thread = whole sectors (16 B pieces would reach L2 as partial-sector writes)
#pragma unroll
            for (int q = 0; q < CH / 8; ++q) {
              float o[8];
              // per-channel epilogue constants: 128-bit read-only loads (same address in every lane -> one broadcast each)
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + cbase + 8 * q));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.bias + cbase + 8 * q + 4));
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int u = 0; u < 8; ++u) o[u] = fmaf(acc[mi][8 * q + u], a.wscale_inv, bb[u]);  // wscale_inv = 1: plain add
              if (a.bn_scale) {
                const float4 s0 = __ldg(reinterpret_cast<const float4*>(a.bn_scale + cbase + 8 * q));
                const float4 s1 = __ldg(reinterpret_cast<const float4*>(a.bn_scale + cbase + 8 * q + 4));
                const float4 h0 = __ldg(reinterpret_cast<const float4*>(a.bn_shift + cbase + 8 * q));
                const float4 h1 = __ldg(reinterpret_cast<const float4*>(a.bn_shift + cbase + 8 * q + 4));
                const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                const float hh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                for (int u = 0; u < 8; ++u) o[u] = fmaf(o[u], ss[u], hh[u]);
              }
              if (a.relu) {
#pragma unroll
                for (int u = 0; u < 8; ++u) o[u] = fmaxf(o[u], 0.f);
              }
              if (a.store32) {
                asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + 8 * q), "f"(o[0]), "f"(o[1]),
                             "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
                             : "memory");
              } else {
                *reinterpret_cast<float4*>(orow + 8 * q) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(orow + 8 * q + 4) = make_float4(o[4], o[5], o[6], o[7]);
              }
            }
          }
        }
      }
      if (st_on) c_epi += clock64() - te0;
    }
    if (st_on) {
      long long* sp = a.stats + blockIdx.x * 16;
      sp[6] += c_full, sp[9] += c_epi;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 14) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

struct Cfg {
  int c, mt, strip, mtx, R, Wt, Wp, tiles_x, tiles_y, n_ctile, total;
};

// pick the tile shape with the fewest tensor-core rounds over 148 SMs (ties: fewer wasted M rows, then larger tiles)
bool choose(int n_img, int H, int W, int Cout, Cfg* best) {
  if (Cout % 32) return false;
  const int c = (Cout % 64 == 0) ? 64 : 32;
  double best_cost = 1e30;
  bool found = false;
  for (int strip = 0; strip <= 1; ++strip) {
    for (int mt = 1; mt <= 2; ++mt) {
      for (int mtx = 1; mtx <= (strip ? mt : 1); ++mtx) {
        for (int Wt = (strip ? 8 * mtx : 6); Wt <= (strip ? 8 * mtx : 41); ++Wt) {
          Cfg k;
          k.c = c, k.mt = mt, k.strip = strip, k.mtx = mtx, k.Wt = Wt, k.Wp = Wt + 2;
          if (strip) {
            k.R = 16 * (mt / mtx);
          } else {
            k.R = (mt * 128) / k.Wp;
            if (k.R > H) k.R = H;
            if (Wt > W) continue;
          }
          if (k.R < 1 || (k.R + 2) * k.Wp > kPlaneRows || k.R + 2 > 256) continue;
          // last row a shifted view can touch must stay inside the plane buffer
          const int last = strip ? ((k.R + 1) * k.Wp + (Wt - 8) + 2 + 7) : (mt * 128 + 2 * k.Wp + 1);
          if (last >= kPlaneRows) continue;
          k.tiles_x = cdiv(W, Wt), k.tiles_y = cdiv(H, k.R), k.n_ctile = Cout / c;
          k.total = n_img * k.tiles_x * k.tiles_y * k.n_ctile;
          const double rounds = (double)cdiv(k.total, 148);
          // cost ~ rounds x (MMA time of one item + fixed per-item overhead); one M tile leaves the second issuer idle
          const double cost = rounds * (mt == 2 ? 2.15 : 1.35) + 1e-3 * (double)k.total * mt * 128 / ((double)n_img * H * W * k.n_ctile);
          if (cost < best_cost - 1e-9) best_cost = cost, *best = k, found = true;
        }
      }
    }
  }
  return found;
}

size_t smem_bytes(int c) { return (size_t)4 * kPlaneBytes + (size_t)kWStages * 2 * c * 128 + 256 + 1024; }

}  // namespace v2

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

struct TileCfg {
  int cout_t, mt, Wt, Wp, R, plane_rows, nst;
  size_t smem;
};

int g_max_stages = 4;
int g_tile_mode = 1;  // 0: 4 M-tiles, 1 CTA/SM; 1: 2 M-tiles sized for 2 co-resident CTAs/SM (phases of one hide the other's)

bool choose_cfg(int H, int W, int Cout, TileCfg* c) {
  if (Cout % 32) return false;
  c->cout_t = Cout <= 64 ? Cout : 64;  // 2 accumulators x 4 M-tiles x 64 columns = the 512 TMEM columns
  if (Cout % c->cout_t) return false;
  int Wt;
  if (g_tile_mode == 0) {
    c->mt = 4;
    // widest column tile with Wt + 2 <= 98 that divides W when possible
    Wt = W <= 96 ? W : 96;
    for (int cand = 96; cand >= 48; --cand)
      if (W % cand == 0) {
        Wt = cand;
        break;
      }
  } else {
    c->mt = 2;
    Wt = W <= 24 ? W : 24;
    for (int cand = 24; cand >= 12; --cand)
      if (W % cand == 0) {
        Wt = cand;
        break;
      }
  }
  c->Wt = Wt;
  c->Wp = Wt + 2;
  c->R = (c->mt * 128) / c->Wp;
  if (c->R > H) c->R = H;
  if (c->R < 1) return false;
  int rows = c->mt * 128 + 2 * c->Wp + 2;
  int box = (c->R + 2) * c->Wp;
  if (box > rows) rows = box;
  c->plane_rows = (rows + 7) & ~7;
  size_t cap = g_tile_mode == 0 ? (size_t)kMaxSmem : (size_t)113 * 1024;
  for (c->nst = g_max_stages; c->nst >= 2; --c->nst) {  // deepest weight pipeline that fits
    c->smem = (size_t)2 * c->plane_rows * 128 + (size_t)2 * c->nst * c->cout_t * 128 + 128 + 1024;
    if (c->smem <= cap) break;
  }
  return c->nst >= 2 && c->smem <= cap && c->Wp <= 256 && c->R + 2 <= 256;
}

int g_base_offset_mode = 0;
int g_min_hw = 16;
long long* g_stats = nullptr;
int g_dbg = 0;
int g_impl = 2;  // 2: persistent pipelined kernel (v2); 1: the first-generation kernel

}  // namespace

extern "C" int pcab_conv3x3_tc_set_base_offset_mode(int mode) {
  g_base_offset_mode = mode & 1;
  g_tile_mode = ((mode >> 1) & 1) ^ 1;  // bit 1 set selects the large 1-CTA/SM tile shape (tuning knob)
  g_min_hw = (mode & 4) ? 32 : 16;        // bit 2 set: leave maps below 32x32 to the FP32 path
  g_max_stages = (mode & 8) ? 2 : 4;      // bit 3 set: 2-stage weight pipeline
  g_impl = (mode & 16) ? 1 : 2;           // bit 4 set: first-generation kernel
  g_dbg = (mode >> 5) & 31;               // bits 5-7: v2 timing experiments (results are wrong when set)
  return 0;
}

// debug: device buffer of 148*16 int64 wait-cycle counters filled by the v2 kernel (null = off)
extern "C" int pcab_conv3x3_tc_set_stats(long long* device_counters) {
  g_stats = device_counters;
  return 0;
}

// the tile plan the v2 kernel would use: out[0..9] = c, mt, strip, mtx, R, Wt, tiles_x, tiles_y, cout tiles, work items
extern "C" int pcab_conv3x3_tc_plan(int n_images, int H, int W, int Cout, int* out10) {
  v2::Cfg k;
  PCAB_REQUIRE(v2::choose(n_images, H, W, Cout, &k), "unsupported shape");
  int v[10] = {k.c, k.mt, k.strip, k.mtx, k.R, k.Wt, k.tiles_x, k.tiles_y, k.n_ctile, k.total};
  for (int i = 0; i < 10; ++i) out10[i] = v[i];
  return PCAB_OK;
}

extern "C" int pcab_conv3x3_tc_supported(int n_sources, int c0, int c1, int c2, int Cout, int H, int W) {
  if (n_sources < 1 || n_sources > 3) return 0;
  int cs[3] = {c0, c1, c2};
  for (int s = 0; s < n_sources; ++s)
    if (cs[s] <= 0 || cs[s] % 32) return 0;
  if (g_impl == 2) {
    v2::Cfg k;
    return (H >= 8 && W >= 8 && v2::choose(1, H, W, Cout, &k)) ? 1 : 0;
  }
  if (H < g_min_hw || W < g_min_hw) return 0;  // the smallest maps keep the FP32 CUDA-core path (too few tiles for 148 SMs)
  TileCfg c;
  return choose_cfg(H, W, Cout, &c) ? 1 : 0;
}

// floats in the tensor-core weight pack: [2 (hi, lo)][Cout][9 * cin_total]
extern "C" size_t pcab_conv3x3_tc_pack_floats(int cin_total, int Cout) { return (size_t)2 * Cout * 9 * cin_total; }

namespace {
int conv3x3_tc_v2(EncodeTiledFn enc, const float* src0, int c0, const float* src1, int c1, const float* src2, int c2,
                  int temporal_T, const void* weight_tc_packed, const float* bias, const float* bn_scale,
                  const float* bn_shift, int relu, float* out, int n_images, int H, int W, int Cout, int out_cstride,
                  int out_coff, cudaStream_t stream, bool f16 = false, float wscale_inv = 1.f) {
  v2::Cfg cfg;
  PCAB_REQUIRE(v2::choose(n_images, H, W, Cout, &cfg), "unsupported shape");
  PCAB_REQUIRE(out_cstride % 4 == 0 && out_coff % 4 == 0 && ((uintptr_t)out & 15) == 0, "output channel layout must be 16B aligned");
  PCAB_REQUIRE(((uintptr_t)bias & 15) == 0 && ((uintptr_t)bn_scale & 15) == 0 && ((uintptr_t)bn_shift & 15) == 0,
               "bias / BN vectors must be 16B aligned");
  const float* srcs[3] = {src0, src1, src2};
  int cs[3] = {c0, c1, c2};
  int nsrc = src2 ? 3 : (src1 ? 2 : 1);
  int T = temporal_T > 1 ? temporal_T : 1;
  if (T > 1) PCAB_REQUIRE(nsrc == 3 && src0 == src1 && src1 == src2, "temporal mode takes the same tensor three times");
  int cin_total = 0;
  for (int s = 0; s < nsrc; ++s) {
    PCAB_REQUIRE(cs[s] > 0 && cs[s] % 32 == 0, "source channels must be a multiple of 32");
    cin_total += cs[s];
  }
  CUtensorMap maps[4];
  for (int s = 0; s < 3; ++s) {
    int ss = s < nsrc ? s : 0;
    cuuint64_t dims[4] = {(cuuint64_t)cs[ss], (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_images};
    cuuint64_t strides[3] = {(cuuint64_t)cs[ss] * 4, (cuuint64_t)W * cs[ss] * 4, (cuuint64_t)H * W * cs[ss] * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)cfg.Wp, (cuuint32_t)(cfg.R + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&maps[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)srcs[ss], dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      pcab_set_error("pcab_conv3x3_tc: cuTensorMapEncodeTiled(A%d) failed: %d", s, (int)r);
      return PCAB_ERR_CUDA;
    }
  }
  {
    // tf32: [2*Cout rows][K] fp32, box 32 x c (128 B rows).  fp16 pairs: [2*Cout rows][2K] fp16 (64 per 32-channel group, 32
    // used), box 64 x c (128 B rows again)
    cuuint64_t K = (cuuint64_t)9 * cin_total * (f16 ? 2 : 1);
    cuuint64_t dims[2] = {K, (cuuint64_t)2 * Cout};
    cuuint64_t strides[1] = {K * (f16 ? 2 : 4)};
    cuuint32_t box[2] = {f16 ? 64u : 32u, (cuuint32_t)cfg.c};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&maps[3], f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)weight_tc_packed,
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      pcab_set_error("pcab_conv3x3_tc: cuTensorMapEncodeTiled(B) failed: %d", (int)r);
      return PCAB_ERR_CUDA;
    }
  }
  v2::Args a;
  a.nsrc = nsrc;
  for (int s = 0; s < 3; ++s) a.src_c[s] = cs[s];
  a.T = T;
  a.N = n_images, a.H = H, a.W = W, a.Cout = Cout;
  a.c = cfg.c, a.mt = cfg.mt, a.strip = cfg.strip, a.mtx = cfg.mtx, a.R = cfg.R, a.Wt = cfg.Wt, a.Wp = cfg.Wp;
  a.store32 = (out_cstride % 8 == 0 && out_coff % 8 == 0 && ((uintptr_t)out & 31) == 0) ? 1 : 0;
  a.tiles_x = cfg.tiles_x, a.tiles_y = cfg.tiles_y, a.n_ctile = cfg.n_ctile, a.total_items = cfg.total;
  a.relu = relu, a.out_cstride = out_cstride, a.out_coff = out_coff;
  a.bias = bias, a.bn_scale = bn_scale, a.bn_shift = bn_shift, a.out = out;
  a.wscale_inv = wscale_inv;
  a.stats = g_stats;
  a.dbg = g_dbg;
  static bool configured = false;
  if (!configured) {
    PCAB_CUDA(cudaFuncSetAttribute(v2::k_conv3x3_tc2<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v2::smem_bytes(32)));
    PCAB_CUDA(cudaFuncSetAttribute(v2::k_conv3x3_tc2<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v2::smem_bytes(64)));
    PCAB_CUDA(cudaFuncSetAttribute(v2::k_conv3x3_tc2<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v2::smem_bytes(32)));
    PCAB_CUDA(cudaFuncSetAttribute(v2::k_conv3x3_tc2<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v2::smem_bytes(64)));
    configured = true;
  }
  const int grid = cfg.total < 148 ? cfg.total : 148;
  if (cfg.c == 64 && f16)
    v2::k_conv3x3_tc2<64, true><<<grid, v2::kThreads2, v2::smem_bytes(64), stream>>>(maps[0], maps[1], maps[2], maps[3], a);
  else if (cfg.c == 64)
    v2::k_conv3x3_tc2<64, false><<<grid, v2::kThreads2, v2::smem_bytes(64), stream>>>(maps[0], maps[1], maps[2], maps[3], a);
  else if (f16)
    v2::k_conv3x3_tc2<32, true><<<grid, v2::kThreads2, v2::smem_bytes(32), stream>>>(maps[0], maps[1], maps[2], maps[3], a);
  else
    v2::k_conv3x3_tc2<32, false><<<grid, v2::kThreads2, v2::smem_bytes(32), stream>>>(maps[0], maps[1], maps[2], maps[3], a);
  PCAB_CHECK_LAUNCH("pcab_conv3x3_tc(v2)");
  return PCAB_OK;
}
}  // namespace

// The same convolution with fp16-pair operands (kind::f16, twice the MMA rate of the tf32 formulation at the same accuracy
// class).  weight_f16_packed: fp16 [2 (h, l)][Cout][9*cin_total/32][64] - per 32-channel group of the tf32 pack's K order 32 values
// + 32 zeros - of the weights multiplied by 1/weight_scale_inv (a power of two that keeps the l halves out of the fp16
// subnormals).  Activations beyond +-65504 saturate.
extern "C" int pcab_conv3x3_tc_f16(const float* src0, int c0, const float* src1, int c1, const float* src2, int c2,
                                   int temporal_T, const void* weight_f16_packed, float weight_scale_inv, const float* bias,
                                   const float* bn_scale, const float* bn_shift, int relu, float* out, int n_images, int H,
                                   int W, int Cout, int out_cstride, int out_coff, cudaStream_t stream) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    pcab_set_error("pcab_conv3x3_tc_f16: cuTensorMapEncodeTiled unavailable");
    return PCAB_ERR_CUDA;
  }
  return conv3x3_tc_v2(enc, src0, c0, src1, c1, src2, c2, temporal_T, weight_f16_packed, bias, bn_scale, bn_shift, relu, out,
                       n_images, H, W, Cout, out_cstride, out_coff, stream, true, weight_scale_inv);
}

extern "C" int pcab_conv3x3_tc(const float* src0, int c0, const float* src1, int c1, const float* src2, int c2,
                               int temporal_T, const float* weight_tc_packed, const float* bias, const float* bn_scale,
                               const float* bn_shift, int relu, float* out, int n_images, int H, int W, int Cout,
                               int out_cstride, int out_coff, cudaStream_t stream) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    pcab_set_error("pcab_conv3x3_tc: cuTensorMapEncodeTiled unavailable");
    return PCAB_ERR_CUDA;
  }
  if (g_impl == 2)
    return conv3x3_tc_v2(enc, src0, c0, src1, c1, src2, c2, temporal_T, weight_tc_packed, bias, bn_scale, bn_shift, relu, out,
                         n_images, H, W, Cout, out_cstride, out_coff, stream);
  TileCfg cfg;
  PCAB_REQUIRE(choose_cfg(H, W, Cout, &cfg), "unsupported shape");
  PCAB_REQUIRE(out_cstride % 4 == 0 && out_coff % 4 == 0, "output channel layout must be 16B aligned");
  const float* srcs[3] = {src0, src1, src2};
  int cs[3] = {c0, c1, c2};
  int nsrc = src2 ? 3 : (src1 ? 2 : 1);
  int T = temporal_T > 1 ? temporal_T : 1;
  if (T > 1) PCAB_REQUIRE(nsrc == 3 && src0 == src1 && src1 == src2, "temporal mode takes the same tensor three times");
  int cin_total = 0;
  for (int s = 0; s < nsrc; ++s) cin_total += cs[s];

  CUtensorMap maps[4];
  for (int s = 0; s < 3; ++s) {
    int ss = s < nsrc ? s : 0;
    cuuint64_t dims[4] = {(cuuint64_t)cs[ss], (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_images};
    cuuint64_t strides[3] = {(cuuint64_t)cs[ss] * 4, (cuuint64_t)W * cs[ss] * 4, (cuuint64_t)H * W * cs[ss] * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)cfg.Wp, (cuuint32_t)(cfg.R + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&maps[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)srcs[ss], dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      pcab_set_error("pcab_conv3x3_tc: cuTensorMapEncodeTiled(A%d) failed: %d", s, (int)r);
      return PCAB_ERR_CUDA;
    }
  }
  {
    cuuint64_t K = (cuuint64_t)9 * cin_total;
    cuuint64_t dims[2] = {K, (cuuint64_t)2 * Cout};
    cuuint64_t strides[1] = {K * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)cfg.cout_t};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&maps[3], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)weight_tc_packed, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      pcab_set_error("pcab_conv3x3_tc: cuTensorMapEncodeTiled(B) failed: %d", (int)r);
      return PCAB_ERR_CUDA;
    }
  }
  TcArgs a;
  a.nsrc = nsrc;
  for (int s = 0; s < 3; ++s) a.src_c[s] = cs[s];
  a.T = T;
  a.N = n_images, a.H = H, a.W = W, a.Cout = Cout;
  a.cout_t = cfg.cout_t, a.mt = cfg.mt, a.R = cfg.R, a.Wt = cfg.Wt, a.Wp = cfg.Wp, a.plane_rows = cfg.plane_rows;
  a.tiles_x = cdiv(W, cfg.Wt), a.tiles_y = cdiv(H, cfg.R);
  a.relu = relu, a.out_cstride = out_cstride, a.out_coff = out_coff, a.base_offset_mode = g_base_offset_mode;
  a.nst = cfg.nst;
  a.bias = bias, a.bn_scale = bn_scale, a.bn_shift = bn_shift, a.out = out;
  static bool configured = false;
  if (!configured) {
    PCAB_CUDA(cudaFuncSetAttribute(k_conv3x3_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    configured = true;
  }
  dim3 grid(a.tiles_x * a.tiles_y * n_images, Cout / cfg.cout_t);
  k_conv3x3_tc<<<grid, kThreads, cfg.smem, stream>>>(maps[0], maps[1], maps[2], maps[3], a);
  PCAB_CHECK_LAUNCH("pcab_conv3x3_tc");
  return PCAB_OK;
}
