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
#include "common.cuh"
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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spins = 0;; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (spins > (1u << 26)) __trap();  // never hang the GPU: a protocol bug becomes a launch error instead
  }
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t dst, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// K-major, 128B-swizzled shared-memory operand descriptor (8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, int base_offset_mode) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  if (base_offset_mode) d |= (uint64_t)((saddr >> 7) & 7) << 49;
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
// wait for the outstanding tcgen05.ld's; the "+r" operands tie the loaded registers to the wait so that no use of
// them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32], uint32_t (&w)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]) : : "memory");
  asm volatile("" : "+r"(w[0]), "+r"(w[1]), "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8]), "+r"(w[9]), "+r"(w[10]), "+r"(w[11]), "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15]), "+r"(w[16]), "+r"(w[17]), "+r"(w[18]), "+r"(w[19]), "+r"(w[20]), "+r"(w[21]), "+r"(w[22]), "+r"(w[23]), "+r"(w[24]), "+r"(w[25]), "+r"(w[26]), "+r"(w[27]), "+r"(w[28]), "+r"(w[29]), "+r"(w[30]), "+r"(w[31]) : : "memory");
}

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

}  // namespace

extern "C" int pcab_conv3x3_tc_set_base_offset_mode(int mode) {
  g_base_offset_mode = mode & 1;
  g_tile_mode = ((mode >> 1) & 1) ^ 1;  // bit 1 set selects the large 1-CTA/SM tile shape (tuning knob)
  g_min_hw = (mode & 4) ? 32 : 16;        // bit 2 set: leave maps below 32x32 to the FP32 path
  g_max_stages = (mode & 8) ? 2 : 4;      // bit 3 set: 2-stage weight pipeline
  return 0;
}

extern "C" int pcab_conv3x3_tc_supported(int n_sources, int c0, int c1, int c2, int Cout, int H, int W) {
  if (n_sources < 1 || n_sources > 3) return 0;
  int cs[3] = {c0, c1, c2};
  for (int s = 0; s < n_sources; ++s)
    if (cs[s] <= 0 || cs[s] % 32) return 0;
  if (H < g_min_hw || W < g_min_hw) return 0;  // the smallest maps keep the FP32 CUDA-core path (too few tiles for 148 SMs)
  TileCfg c;
  return choose_cfg(H, W, Cout, &c) ? 1 : 0;
}

// floats in the tensor-core weight pack: [2 (hi, lo)][Cout][9 * cin_total]
extern "C" size_t pcab_conv3x3_tc_pack_floats(int cin_total, int Cout) { return (size_t)2 * Cout * 9 * cin_total; }

extern "C" int pcab_conv3x3_tc(const float* src0, int c0, const float* src1, int c1, const float* src2, int c2,
                               int temporal_T, const float* weight_tc_packed, const float* bias, const float* bn_scale,
                               const float* bn_shift, int relu, float* out, int n_images, int H, int W, int Cout,
                               int out_cstride, int out_coff, cudaStream_t stream) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    pcab_set_error("pcab_conv3x3_tc: cuTensorMapEncodeTiled unavailable");
    return PCAB_ERR_CUDA;
  }
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
