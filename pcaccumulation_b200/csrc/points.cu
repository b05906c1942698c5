// Per-point stages: index compaction, bilinear feature pickup (ungrid), the STPN point head and the
// TubeNet (TPointNet) embeddings / pose regression / rigid reconstruction.
//
// Replaces models/pillar_encoder.py:206-267 (temporal_ungrid / ungrid), models/stpn.py:91-103 (point
// decoder), models/tpointnet.py:211-305 (TPointNet.forward, batch_quat2mat :20-40),
// toolbox/se3_utils.py:44-64 (quat2mat), toolbox/register_utils.py:72-93 (reconstruct_sequence) and the
// torch_scatter segment max / mean calls inside them.
#include <cub/cub.cuh>
#include "mlp.cuh"
#include "pair16.cuh"
#include "pcab200.h"

namespace {

using mlp::PTS;
using mlp::LDP;

__global__ void k_select_write(const int* __restrict__ flag, const int* __restrict__ pos, int n, int* __restrict__ idx,
                               int* __restrict__ count) {
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (flag[i]) idx[pos[i]] = i;
    if (i == n - 1) *count = pos[i] + (flag[i] ? 1 : 0);
  }
}

__global__ void k_flag_eq(const long long* __restrict__ v, int n, long long value, int* __restrict__ flag) {
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) flag[i] = v[i] == value;
}

// out[j][0..C) = bilinear(feats[frame_of(point)], xy(point)), border padding.  One warp per point.
template <bool P16>
__global__ void k_ungrid(const float* __restrict__ feats, int C, int H, int W, const float* __restrict__ xyz,
                         const int* __restrict__ frame_of_point, const int* __restrict__ idx, int k, float x_abs,
                         float y_abs, float* __restrict__ out) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int j = warp; j < k; j += nwarp) {
    int i = idx ? idx[j] : j;
    mlp::Bilinear bl = mlp::bilinear_border(xyz[3 * i], xyz[3 * i + 1], x_abs, y_abs, H, W);
    const size_t fb = (size_t)frame_of_point[i] * H * W;
    const float* base = feats + fb * C;
    for (int c = lane; c < C; c += 32) {
      float t00, t01, t10, t11;
      if (P16) {
        t00 = p16::load1(feats, fb + bl.o00, C, c), t01 = p16::load1(feats, fb + bl.o01, C, c);
        t10 = p16::load1(feats, fb + bl.o10, C, c), t11 = p16::load1(feats, fb + bl.o11, C, c);
      } else {
        t00 = base[(size_t)bl.o00 * C + c], t01 = base[(size_t)bl.o01 * C + c];
        t10 = base[(size_t)bl.o10 * C + c], t11 = base[(size_t)bl.o11 * C + c];
      }
      float v = t00 * bl.w00;
      v = fmaf(t01, bl.w01, v);
      v = fmaf(t10, bl.w10, v);
      v = fmaf(t11, bl.w11, v);
      out[(size_t)j * C + c] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// STPN point head.  Weight pack ([in][out] matrices):
//   pe0 W[3][32] b[32] | pe2 W[32][64] b[64] | fp W[128][128] b[128] |
//   mos0 W[128][128] b[128] s[128] t[128] | mos3 W[128][2] b[2] | off0 W,b,s,t | off3 W[128][2] b[2]
// ---------------------------------------------------------------------------------------------
using namespace mlp::stpn_pack;

__global__ void __launch_bounds__(mlp::NT, 2) k_stpn_head(const float* __restrict__ mos_feats /* [B,H,W,64] */, int H, int W,
                                                   const float* __restrict__ tp, const int* __restrict__ pbatch,
                                                   const int* __restrict__ fg_idx, int k, const float* __restrict__ pk,
                                                   float x_abs, float y_abs, float* __restrict__ mos_out,
                                                   float* __restrict__ off_out) {
  extern __shared__ __align__(16) float sm[];
  float* E = sm;                  // [128][LDP]
  float* F = E + 128 * LDP;       // [128][LDP]
  float* G = E;                   // the hidden layer of the motion head reuses E (dead after final_proj)
  float* s_w = F + 128 * LDP;     // [SW_FLOATS]
  float* s_o = s_w + mlp::SW_FLOATS;  // [4][LDP]
  __shared__ int s_idx[PTS];
  const int base = blockIdx.x * PTS;
  if (threadIdx.x < PTS) s_idx[threadIdx.x] = (base + threadIdx.x < k) ? fg_idx[base + threadIdx.x] : -1;
  __syncthreads();
  // bilinear pickup of the 64 motion-feature channels into E[64..127] - first, with four iterations (16 independent
  // 128-bit loads per thread) in flight: these are scattered L2 reads and nothing else in the kernel hides their latency
#pragma unroll 4
  for (int e = threadIdx.x; e < PTS * 16; e += mlp::NT) {
    int p = e / 16, q = e % 16;
    const bool valid = s_idx[p] >= 0;
    const int i = valid ? s_idx[p] : s_idx[0];  // padding rows read a real point so the loads stay branch-free
    mlp::Bilinear bl = mlp::bilinear_border(tp[3 * i], tp[3 * i + 1], x_abs, y_abs, H, W);
    const float4* b4 = reinterpret_cast<const float4*>(mos_feats + (size_t)pbatch[i] * H * W * 64) + q;
    float4 a = b4[(size_t)bl.o00 * 16], b = b4[(size_t)bl.o01 * 16], c = b4[(size_t)bl.o10 * 16], d = b4[(size_t)bl.o11 * 16];
    float4 v;
    v.x = fmaf(d.x, bl.w11, fmaf(c.x, bl.w10, fmaf(b.x, bl.w01, a.x * bl.w00)));
    v.y = fmaf(d.y, bl.w11, fmaf(c.y, bl.w10, fmaf(b.y, bl.w01, a.y * bl.w00)));
    v.z = fmaf(d.z, bl.w11, fmaf(c.z, bl.w10, fmaf(b.z, bl.w01, a.z * bl.w00)));
    v.w = fmaf(d.w, bl.w11, fmaf(c.w, bl.w10, fmaf(b.w, bl.w01, a.w * bl.w00)));
    if (!valid) v = make_float4(0.f, 0.f, 0.f, 0.f);
    E[(64 + 4 * q + 0) * LDP + p] = v.x;
    E[(64 + 4 * q + 1) * LDP + p] = v.y;
    E[(64 + 4 * q + 2) * LDP + p] = v.z;
    E[(64 + 4 * q + 3) * LDP + p] = v.w;
  }
  // inputs: pos = p / |x_min| (all three by x scale, models/stpn.py:94), channel-major into G[0..2]
  for (int e = threadIdx.x; e < 3 * PTS; e += mlp::NT) {
    int c = e / PTS, p = e % PTS;
    int i = s_idx[p];
    G[c * LDP + p] = i >= 0 ? tp[3 * i + c] / x_abs : 0.f;
  }
  __syncthreads();
  mlp::block_dense<3, 32>(G, pk + S_PE0W, pk + S_PE0B, nullptr, nullptr, true, F, s_w);
  mlp::block_dense<32, 64>(F, pk + S_PE2W, pk + S_PE2B, nullptr, nullptr, true, E, s_w);
  mlp::block_dense<128, 128>(E, pk + S_FPW, pk + S_FPB, nullptr, nullptr, true, F, s_w);
  mlp::block_dense<128, 128>(F, pk + S_M0W, pk + S_M0B, pk + S_M0S, pk + S_M0T, true, G, s_w);
  mlp::block_dense_small<128, 2>(G, pk + S_M3W, pk + S_M3B, s_o);
  mlp::block_dense<128, 128>(F, pk + S_O0W, pk + S_O0B, pk + S_O0S, pk + S_O0T, true, E, s_w);
  mlp::block_dense_small<128, 2>(E, pk + S_O3W, pk + S_O3B, s_o + 2 * LDP);
  for (int e = threadIdx.x; e < PTS * 2; e += mlp::NT) {
    int p = e % PTS, o = e / PTS;
    int i = s_idx[p];
    if (i < 0) continue;
    mos_out[2 * i + o] = s_o[o * LDP + p];
    float v = s_o[(2 + o) * LDP + p];
    if (isnan(v) || isinf(v)) v = 0.f;  // safe_guard_offset (models/stpn.py:61-65)
    off_out[2 * i + o] = fminf(fmaxf(v, -20.f), 20.f);
  }
}

__global__ void k_init_point_outputs(int n, float* __restrict__ mos, float* __restrict__ off) {
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    mos[2 * i] = 1.f, mos[2 * i + 1] = 0.f;
    off[2 * i] = 0.f, off[2 * i + 1] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
// TubeNet
// ---------------------------------------------------------------------------------------------
// three-layer embed: IN -> H1 relu -> H2 relu -> 128, then atomic max into dst[seg[p]][128]
template <int IN, int H1, int H2>
__device__ void embed_and_max(float* A, float* Bf, float* s_w, const float* __restrict__ pk, const int* s_seg,
                              float* __restrict__ dst) {
  const float* W0 = pk;
  const float* b0 = W0 + IN * H1;
  const float* W1 = b0 + H1;
  const float* b1 = W1 + H1 * H2;
  const float* W2 = b1 + H2;
  const float* b2 = W2 + H2 * 128;
  mlp::block_dense<IN, H1>(A, W0, b0, nullptr, nullptr, true, Bf, s_w);
  mlp::block_dense<H1, H2>(Bf, W1, b1, nullptr, nullptr, true, A, s_w);
  mlp::block_dense<H2, 128>(A, W2, b2, nullptr, nullptr, false, Bf, s_w);
  // points arrive sorted by segment, so a CTA sees a few runs: reduce each run in shared memory and issue ONE
  // atomic per (run, channel).  One thread per channel scans the rows.
  {
    const int c = threadIdx.x;
    static_assert(mlp::NT == 128, "one thread per output channel");
    const float* row = Bf + c * LDP;
    int cur_seg = -1;
    float cur = -INFINITY;
    for (int p = 0; p < PTS; ++p) {
      int s = s_seg[p];
      if (s != cur_seg) {
        if (cur_seg >= 0) mlp::atomic_max_float(dst + (size_t)cur_seg * 128 + c, cur);
        cur_seg = s;
        cur = -INFINITY;
      }
      if (s >= 0) cur = fmaxf(cur, row[p]);
    }
    if (cur_seg >= 0) mlp::atomic_max_float(dst + (size_t)cur_seg * 128 + c, cur);
  }
  __syncthreads();
}

// motion (64->64->128->128) and geometry (32->32->64->128) embeddings, max-pooled per instance
__global__ void __launch_bounds__(mlp::NT, 2) k_tpn_static_embed(const float* __restrict__ mos_feat,
                                                          const float* __restrict__ geo_feat,
                                                          const int* __restrict__ src_idx, const int* __restrict__ inst,
                                                          int n, const float* __restrict__ pk_motion,
                                                          const float* __restrict__ pk_geo, float* __restrict__ mos_emb,
                                                          float* __restrict__ geo_emb) {
  extern __shared__ __align__(16) float sm[];
  float* A = sm;
  float* Bf = A + 128 * LDP;
  float* s_w = Bf + 128 * LDP;
  __shared__ int s_seg[PTS], s_src[PTS];
  int base = blockIdx.x * PTS;
  if (threadIdx.x < PTS) {
    int j = base + threadIdx.x;
    s_seg[threadIdx.x] = j < n ? inst[j] : -1;
    s_src[threadIdx.x] = j < n ? src_idx[j] : -1;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < PTS * 64; e += mlp::NT) {
    int p = e / 64, c = e % 64;
    A[c * LDP + p] = s_src[p] >= 0 ? mos_feat[(size_t)s_src[p] * 64 + c] : 0.f;
  }
  __syncthreads();
  embed_and_max<64, 64, 128>(A, Bf, s_w, pk_motion, s_seg, mos_emb);
  for (int e = threadIdx.x; e < PTS * 32; e += mlp::NT) {
    int p = e / 32, c = e % 32;
    A[c * LDP + p] = s_src[p] >= 0 ? geo_feat[(size_t)s_src[p] * 32 + c] : 0.f;
  }
  __syncthreads();
  embed_and_max<32, 32, 64>(A, Bf, s_w, pk_geo, s_seg, geo_emb);
}

// per (instance, frame) sums of the points (double) and counts.  Rows arrive sorted by segment, so each warp reduces
// its 32 consecutive rows with a segmented shuffle scan and issues one atomic per run instead of one per row.
__global__ void k_tpn_frame_sums(const float* __restrict__ pts, const int* __restrict__ inst,
                                 const int* __restrict__ tidx, int T, int n, double* __restrict__ sums /* [KT][4] */) {
  int stride = gridDim.x * blockDim.x;
  int lane = threadIdx.x & 31;
  for (int j0 = (blockIdx.x * blockDim.x + threadIdx.x) - lane; j0 < n; j0 += stride) {
    int j = j0 + lane;
    bool ok = j < n;
    int seg = ok ? inst[j] * T + tidx[j] : -1;
    double x = ok ? (double)pts[3 * j] : 0.0, y = ok ? (double)pts[3 * j + 1] : 0.0, z = ok ? (double)pts[3 * j + 2] : 0.0;
    double c = ok ? 1.0 : 0.0;
    // inclusive segmented suffix sums: lane i accumulates lanes i.. of its run; the run head then holds the run total
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int s2 = __shfl_down_sync(0xffffffffu, seg, o);
      double x2 = __shfl_down_sync(0xffffffffu, x, o), y2 = __shfl_down_sync(0xffffffffu, y, o);
      double z2 = __shfl_down_sync(0xffffffffu, z, o), c2 = __shfl_down_sync(0xffffffffu, c, o);
      if (lane + o < 32 && s2 == seg) x += x2, y += y2, z += z2, c += c2;
    }
    int prev = __shfl_up_sync(0xffffffffu, seg, 1);
    bool head = ok && (lane == 0 || prev != seg);
    if (head) {
      double* s = sums + (size_t)seg * 4;
      atomicAdd(s, x), atomicAdd(s + 1, y), atomicAdd(s + 2, z), atomicAdd(s + 3, c);
    }
  }
}

// positional embedding of [p - anchor_centroid(inst), t/T] (4->32->64->128), max-pooled per (inst, frame)
__global__ void __launch_bounds__(mlp::NT, 2) k_tpn_pos_embed(const float* __restrict__ pts, const int* __restrict__ inst,
                                                       const int* __restrict__ tidx, int n, int T,
                                                       const double* __restrict__ sums, const float* __restrict__ pk_pos,
                                                       float* __restrict__ frame_emb) {
  extern __shared__ __align__(16) float sm[];
  float* A = sm;
  float* Bf = A + 128 * LDP;
  float* s_w = Bf + 128 * LDP;
  __shared__ int s_seg[PTS];
  int base = blockIdx.x * PTS;
  if (threadIdx.x < PTS) {
    int j = base + threadIdx.x;
    int p = threadIdx.x;
    if (j < n) {
      int k = inst[j], t = tidx[j];
      s_seg[p] = k * T + t;
      const double* s = sums + (size_t)k * T * 4;  // anchor frame (t = 0) of the instance
      double cnt = s[3] > 0 ? s[3] : 1.0;
      A[0 * LDP + p] = pts[3 * j] - (float)(s[0] / cnt);
      A[1 * LDP + p] = pts[3 * j + 1] - (float)(s[1] / cnt);
      A[2 * LDP + p] = pts[3 * j + 2] - (float)(s[2] / cnt);
      A[3 * LDP + p] = (float)((double)t / (double)T);
    } else {
      s_seg[p] = -1;
      A[p] = A[LDP + p] = A[2 * LDP + p] = A[3 * LDP + p] = 0.f;
    }
  }
  __syncthreads();
  embed_and_max<4, 32, 64>(A, Bf, s_w, pk_pos, s_seg, frame_emb);
}

__global__ void k_fix_neg_inf(float* __restrict__ a, long long n) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
    if (a[i] == -INFINITY) a[i] = 0.f;  // torch_scatter leaves empty segments at 0
}

// regressor input rows [K*T][512] = [geo(k), mos(k), frame(k,t), frame(k,0)]
__global__ void k_tpn_regressor_input(const float* __restrict__ geo_emb, const float* __restrict__ mos_emb,
                                      const float* __restrict__ frame_emb, int KT, int T, float* __restrict__ X) {
  long long total = (long long)KT * 512;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += stride) {
    int row = (int)(e / 512), c = (int)(e % 512);
    int k = row / T;
    float v;
    if (c < 128) v = geo_emb[(size_t)k * 128 + c];
    else if (c < 256) v = mos_emb[(size_t)k * 128 + c - 128];
    else if (c < 384) v = frame_emb[(size_t)row * 128 + c - 256];
    else v = frame_emb[(size_t)k * T * 128 + c - 384];
    X[e] = v;
  }
}

// The three layers of the TubeNet regressor (512 -> 256 -> BN -> ReLU -> 128 -> BN -> ReLU -> 7) for RR rows per CTA in one
// launch: the hidden rows stay in shared memory and every weight element is read once per CTA (coalesced over the output
// index).  Each output is one rounding sequence: fmaf over ascending k from 0, + bias, BN fmaf, ReLU.
constexpr int RR = 4;
__global__ void __launch_bounds__(256) k_regressor(const float* __restrict__ X, int R, const float* __restrict__ W0,
                                                   const float* __restrict__ b0, const float* __restrict__ s0,
                                                   const float* __restrict__ t0, const float* __restrict__ W1,
                                                   const float* __restrict__ b1, const float* __restrict__ s1,
                                                   const float* __restrict__ t1, const float* __restrict__ W2,
                                                   const float* __restrict__ b2, float* __restrict__ Y) {
  __shared__ __align__(16) float sx[RR][512];
  __shared__ __align__(16) float h0[RR][256];
  __shared__ __align__(16) float h1[RR][128];
  const int r0 = blockIdx.x * RR, tid = threadIdx.x;
  for (int e = tid; e < RR * 128; e += 256) {
    const int r = e >> 7, q = e & 127;
    reinterpret_cast<float4*>(sx[r])[q] = r0 + r < R ? reinterpret_cast<const float4*>(X + (size_t)(r0 + r) * 512)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  {
    float a[RR] = {};
    const float* w = W0 + tid;
#pragma unroll 1
    for (int k0 = 0; k0 < 512; k0 += 16) {
      float wv[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) wv[u] = w[(size_t)(k0 + u) * 256];
#pragma unroll
      for (int u = 0; u < 16; ++u)
#pragma unroll
        for (int r = 0; r < RR; ++r) a[r] = fmaf(sx[r][k0 + u], wv[u], a[r]);
    }
#pragma unroll
    for (int r = 0; r < RR; ++r) h0[r][tid] = fmaxf(fmaf(a[r] + b0[tid], s0[tid], t0[tid]), 0.f);
  }
  __syncthreads();
  if (tid < 128) {
    float a[RR] = {};
    const float* w = W1 + tid;
#pragma unroll 1
    for (int k0 = 0; k0 < 256; k0 += 16) {
      float wv[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) wv[u] = w[(size_t)(k0 + u) * 128];
#pragma unroll
      for (int u = 0; u < 16; ++u)
#pragma unroll
        for (int r = 0; r < RR; ++r) a[r] = fmaf(h0[r][k0 + u], wv[u], a[r]);
    }
#pragma unroll
    for (int r = 0; r < RR; ++r) h1[r][tid] = fmaxf(fmaf(a[r] + b1[tid], s1[tid], t1[tid]), 0.f);
  }
  __syncthreads();
  if (tid < RR * 7) {
    const int r = tid / 7, o = tid % 7;
    float a = 0.f;
    for (int k = 0; k < 128; ++k) a = fmaf(h1[r][k], W2[k * 7 + o], a);
    if (r0 + r < R) Y[(size_t)(r0 + r) * 7 + o] = a + b2[o];
  }
}

// rep[K*T][7] = (quat xyzw, trans) -> pose [K*T][4][4] with the centring undone and frame 0 := I
__global__ void k_tpn_pose(const float* __restrict__ rep, const double* __restrict__ sums, int KT, int T,
                           float* __restrict__ pose_centered, float* __restrict__ pose) {
  int stride = gridDim.x * blockDim.x;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < KT; r += stride) {
    const float* q = rep + (size_t)r * 7;
    float nrm = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);  // F.normalize
    float x = q[0] / nrm, y = q[1] / nrm, z = q[2] / nrm, w = q[3] / nrm;
    float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
    float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
    float R[9] = {w2 + x2 - y2 - z2, 2 * xy - 2 * wz,   2 * wy + 2 * xz,    //
                  2 * wz + 2 * xy,   w2 - x2 + y2 - z2, 2 * yz - 2 * wx,    //
                  2 * xz - 2 * wy,   2 * wx + 2 * yz,   w2 - x2 - y2 + z2};
    float t[3] = {q[4], q[5], q[6]};
    if (pose_centered) {
      float* P = pose_centered + (size_t)r * 16;
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) P[4 * i + j] = R[3 * i + j];
        P[4 * i + 3] = t[i];
      }
      P[12] = P[13] = P[14] = 0.f, P[15] = 1.f;
    }
    int k = r / T;
    const double* s = sums + (size_t)k * T * 4;
    double cnt = s[3] > 0 ? s[3] : 1.0;
    float c[3] = {(float)(s[0] / cnt), (float)(s[1] / cnt), (float)(s[2] / cnt)};
    float* P = pose + (size_t)r * 16;
    bool anchor = (r % T) == 0;
    for (int i = 0; i < 3; ++i) {
      float d = 0.f;
      for (int j = 0; j < 3; ++j) {
        float e = ((i == j) ? 1.f : 0.f) - R[3 * i + j];
        d = fmaf(e, c[j], d);
        P[4 * i + j] = anchor ? (i == j ? 1.f : 0.f) : R[3 * i + j];
      }
      P[4 * i + 3] = anchor ? 0.f : t[i] + d;
    }
    P[12] = P[13] = P[14] = 0.f, P[15] = 1.f;
  }
}

// out[j] = R[seg[j]] p_j + t[seg[j]]  (reconstruct_sequence / ego_motion_compensation)
__global__ void k_apply_seg_pose(const float* __restrict__ pts, const int* __restrict__ seg,
                                 const float* __restrict__ pose, int n, float* __restrict__ out) {
  int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const float* P = pose + (size_t)seg[j] * 16;
    float x = pts[3 * j], y = pts[3 * j + 1], z = pts[3 * j + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float d = P[4 * r] * x;
      d = fmaf(P[4 * r + 1], y, d);
      d = fmaf(P[4 * r + 2], z, d);
      out[3 * j + r] = d + P[4 * r + 3];
    }
  }
}

__global__ void k_scatter_rows3(const float* __restrict__ src, const int* __restrict__ idx, int k,
                                float* __restrict__ dst) {
  int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += stride) {
    int i = idx[j];
    dst[3 * i] = src[3 * j], dst[3 * i + 1] = src[3 * j + 1], dst[3 * i + 2] = src[3 * j + 2];
  }
}


// ---------------------------------------------------------------------------------------------
// TubeNet bookkeeping (models/alignnet.py:115-163 `padding`, :201-225): drop empty instances and relabel, find the
// instances without anchor-frame (t = 0) points and duplicate the rows of their first non-empty frame as t = 0, order
// all rows by (instance, frame).
// ---------------------------------------------------------------------------------------------
__global__ void k_tpn_hist(const long long* __restrict__ inst, const long long* __restrict__ tidx, int n, int T,
                           int* __restrict__ frame_count) {
  // points arrive in stream order, i.e. long runs of one (instance, frame): lanes with the same bin combine before the atomic
  const int stride = gridDim.x * blockDim.x;
  for (int j0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); j0 < n; j0 += stride) {
    const int j = j0 + (threadIdx.x & 31);
    const int bin = j < n ? (int)inst[j] * T + (int)tidx[j] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(frame_count + bin, __popc(peers));
  }
}

// single block: mapping (old id -> new id or -1), first non-empty frame, pad source frame (or -1), totals {K, P}
__global__ void __launch_bounds__(1024) k_tpn_relabel(const int* __restrict__ frame_count, int K0, int T,
                                                      int* __restrict__ mapping, int* __restrict__ pad_frame,
                                                      int* __restrict__ totals) {
  __shared__ int s_scan[1024];
  __shared__ int s_base, s_pad;
  if (threadIdx.x == 0) s_base = 0, s_pad = 0;
  __syncthreads();
  for (int k0 = 0; k0 < K0; k0 += 1024) {
    int k = k0 + threadIdx.x;
    int cnt = 0, first = -1;
    if (k < K0)
      for (int t = 0; t < T; ++t) {
        int c = frame_count[k * T + t];
        cnt += c;
        if (first < 0 && c > 0) first = t;
      }
    int keep = cnt > 0;
    s_scan[threadIdx.x] = keep;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // inclusive Hillis-Steele scan
      int v = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0;
      __syncthreads();
      s_scan[threadIdx.x] += v;
      __syncthreads();
    }
    if (k < K0) {
      mapping[k] = keep ? s_base + s_scan[threadIdx.x] - 1 : -1;
      bool need = keep && first > 0;  // no anchor-frame points
      pad_frame[k] = need ? first : -1;
      if (need) atomicAdd(&s_pad, frame_count[k * T + first]);
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base += s_scan[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[0] = s_base, totals[1] = s_pad;
}

__global__ void k_tpn_keys(const long long* __restrict__ inst, const long long* __restrict__ tidx, int n, int T,
                           const int* __restrict__ mapping, const int* __restrict__ pad_frame, int* __restrict__ keys,
                           int* __restrict__ vals, int* __restrict__ seg_rows, long long* __restrict__ inst_new,
                           int* __restrict__ pad_cursor) {
  int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    int k = (int)inst[j], t = (int)tidx[j];
    int m = mapping[k];
    keys[j] = m * T + t;
    vals[j] = j;
    seg_rows[j] = m * T + t;
    inst_new[j] = m;
    if (pad_frame[k] == t) {  // duplicate as an anchor-frame row
      int q = n + atomicAdd(pad_cursor, 1);
      keys[q] = m * T;
      vals[q] = j;
    }
  }
}

__global__ void k_tpn_rows(const int* __restrict__ keys_sorted, const int* __restrict__ vals_sorted, int n_pad, int T,
                           const float* __restrict__ pts_rec, int* __restrict__ row_src, int* __restrict__ row_inst,
                           int* __restrict__ row_time, int* __restrict__ row_seg, float* __restrict__ row_pts) {
  int stride = gridDim.x * blockDim.x;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_pad; q += stride) {
    int key = keys_sorted[q], src = vals_sorted[q];
    row_src[q] = src, row_inst[q] = key / T, row_time[q] = key % T, row_seg[q] = key;
    row_pts[3 * q] = pts_rec[3 * src], row_pts[3 * q + 1] = pts_rec[3 * src + 1], row_pts[3 * q + 2] = pts_rec[3 * src + 2];
  }
}

__global__ void k_fill(float* a, long long n, float v) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) a[i] = v;
}

size_t al256p(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t pcab_select_workspace(int n) {
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int*)nullptr, (int*)nullptr, n);
  return 2 * al256p((size_t)n * 4) + al256p(scan_bytes) + 256;
}

// idx = ascending indices i with flags[i] != 0 (or values[i] == value when flags is null); count on device
extern "C" int pcab_select_indices(const int* flags, const long long* values, long long value, int n, int* idx,
                                   int* count, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PCAB_REQUIRE(workspace_bytes >= pcab_select_workspace(n), "workspace too small");
  char* w = (char*)workspace;
  int* flag = (int*)w;
  w += al256p((size_t)n * 4);
  int* pos = (int*)w;
  w += al256p((size_t)n * 4);
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flag, pos, n);
  const int* f = flags;
  if (!f) {
    k_flag_eq<<<grid_for(n, 256), 256, 0, stream>>>(values, n, value, flag);
    f = flag;
  }
  PCAB_CUDA(cub::DeviceScan::ExclusiveSum(w, scan_bytes, f, pos, n, stream));
  k_select_write<<<grid_for(n, 256), 256, 0, stream>>>(f, pos, n, idx, count);
  PCAB_CHECK_LAUNCH("pcab_select_indices");
  return PCAB_OK;
}

extern "C" int pcab_ungrid(const float* feats_nhwc, int C, int fmt, int H, int W, const float* xyz, const int* frame_of_point,
                           const int* idx, int k, float x_abs, float y_abs, float* out, cudaStream_t stream) {
  if (k <= 0) return PCAB_OK;
  PCAB_REQUIRE(!fmt || C % 32 == 0, "P16 tensors have C % 32 == 0");
  if (fmt)
    k_ungrid<true><<<grid_for((long long)k * 32, 256, 8), 256, 0, stream>>>(feats_nhwc, C, H, W, xyz, frame_of_point, idx, k, x_abs,
                                                                           y_abs, out);
  else
    k_ungrid<false><<<grid_for((long long)k * 32, 256, 8), 256, 0, stream>>>(feats_nhwc, C, H, W, xyz, frame_of_point, idx, k, x_abs,
                                                                            y_abs, out);
  PCAB_CHECK_LAUNCH("pcab_ungrid");
  return PCAB_OK;
}

extern "C" int pcab_stpn_head_pack_size(void) { return S_PACK; }

extern "C" int pcab_init_point_outputs(int n_points, float* mos, float* offset, cudaStream_t stream) {
  k_init_point_outputs<<<grid_for(n_points, 256), 256, 0, stream>>>(n_points, mos, offset);
  PCAB_CHECK_LAUNCH("pcab_init_point_outputs");
  return PCAB_OK;
}

extern "C" int pcab_stpn_head(const float* mos_feats_nhwc, int H, int W, const float* transformed_points,
                              const int* point_batch, const int* fg_idx, int n_fg, const float* weight_pack,
                              float x_abs, float y_abs, float* mos_out, float* offset_out, cudaStream_t stream) {
  if (n_fg <= 0) return PCAB_OK;
  size_t smem = (size_t)(2 * 128 * LDP + mlp::SW_FLOATS + 4 * LDP) * sizeof(float);
  static PcabSmemOnce once;
  PCAB_CUDA(pcab_set_max_smem(k_stpn_head, (int)smem, once));
  k_stpn_head<<<cdiv(n_fg, PTS), mlp::NT, smem, stream>>>(mos_feats_nhwc, H, W, transformed_points, point_batch, fg_idx,
                                                      n_fg, weight_pack, x_abs, y_abs, mos_out, offset_out);
  PCAB_CHECK_LAUNCH("pcab_stpn_head");
  return PCAB_OK;
}


// step 1: histogram + relabel.  frame_count [K0*T] (zero-filled here), mapping [K0], pad_frame [K0], totals {K, P} (device)
extern "C" int pcab_tpn_relabel(const long long* inst, const long long* tidx, int n, int K0, int T, int* frame_count,
                                int* mapping, int* pad_frame, int* totals, cudaStream_t stream) {
  PCAB_CUDA(cudaMemsetAsync(frame_count, 0, (size_t)K0 * T * 4, stream));
  k_tpn_hist<<<grid_for(n, 256), 256, 0, stream>>>(inst, tidx, n, T, frame_count);
  k_tpn_relabel<<<1, 1024, 0, stream>>>(frame_count, K0, T, mapping, pad_frame, totals);
  PCAB_CHECK_LAUNCH("pcab_tpn_relabel");
  return PCAB_OK;
}

extern "C" size_t pcab_tpn_rows_workspace(int n_rows) {
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, n_rows);
  return al256p(sort_bytes) + 4 * al256p((size_t)n_rows * 4) + 512;
}

// step 2 (after the host read K and P): the n + P rows ordered by (instance, frame) with their source row, instance,
// frame, segment id and point; also the relabelled instance / segment id of the n original rows
extern "C" int pcab_tpn_rows(const long long* inst, const long long* tidx, int n, int n_pad_rows, int T, const int* mapping,
                             const int* pad_frame, const float* pts_rec, int* seg_rows, long long* inst_new, int* row_src,
                             int* row_inst, int* row_time, int* row_seg, float* row_pts, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream) {
  int total = n + n_pad_rows;
  PCAB_REQUIRE(workspace_bytes >= pcab_tpn_rows_workspace(total), "workspace too small");
  char* w = (char*)workspace;
  int* keys = (int*)w;
  w += al256p((size_t)total * 4);
  int* vals = (int*)w;
  w += al256p((size_t)total * 4);
  int* keys_s = (int*)w;
  w += al256p((size_t)total * 4);
  int* vals_s = (int*)w;
  w += al256p((size_t)total * 4);
  int* cursor = (int*)w;
  w += 256;
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys, keys_s, vals, vals_s, total);
  PCAB_CUDA(cudaMemsetAsync(cursor, 0, 4, stream));
  k_tpn_keys<<<grid_for(n, 256), 256, 0, stream>>>(inst, tidx, n, T, mapping, pad_frame, keys, vals, seg_rows, inst_new, cursor);
  PCAB_CUDA(cub::DeviceRadixSort::SortPairs(w, sort_bytes, keys, keys_s, vals, vals_s, total, 0, 32, stream));
  k_tpn_rows<<<grid_for(total, 256), 256, 0, stream>>>(keys_s, vals_s, total, T, pts_rec, row_src, row_inst, row_time, row_seg,
                                                       row_pts);
  PCAB_CHECK_LAUNCH("pcab_tpn_rows");
  return PCAB_OK;
}

// embeds: each pack = W0[IN][H1] b0 W1[H1][H2] b1 W2[H2][128] b2
extern "C" int pcab_tpn_static_embed(const float* mos_feat, const float* geo_feat, const int* src_idx, const int* inst,
                                     int n, int K, const float* pack_motion, const float* pack_geo, float* mos_emb,
                                     float* geo_emb, cudaStream_t stream) {
  size_t smem = (size_t)(2 * 128 * LDP + mlp::SW_FLOATS) * sizeof(float);
  static PcabSmemOnce once;
  PCAB_CUDA(pcab_set_max_smem(k_tpn_static_embed, (int)smem, once));
  k_fill<<<grid_for((long long)K * 128, 256), 256, 0, stream>>>(mos_emb, (long long)K * 128, -INFINITY);
  k_fill<<<grid_for((long long)K * 128, 256), 256, 0, stream>>>(geo_emb, (long long)K * 128, -INFINITY);
  k_tpn_static_embed<<<cdiv(n, PTS), mlp::NT, smem, stream>>>(mos_feat, geo_feat, src_idx, inst, n, pack_motion, pack_geo,
                                                          mos_emb, geo_emb);
  k_fix_neg_inf<<<grid_for((long long)K * 128, 256), 256, 0, stream>>>(mos_emb, (long long)K * 128);
  k_fix_neg_inf<<<grid_for((long long)K * 128, 256), 256, 0, stream>>>(geo_emb, (long long)K * 128);
  PCAB_CHECK_LAUNCH("pcab_tpn_static_embed");
  return PCAB_OK;
}

// one TPointNet iteration: frame centroids, positional embedding, regressor, pose assembly.
// regressor pack: W0[512][256] b0 s0 t0 | W1[256][128] b1 s1 t1 | W2[128][7] b2
// scratch floats: frame_emb KT*128 | X KT*512 | H0 KT*256 | H1 KT*128 | rep KT*7 ; doubles sums KT*4
extern "C" size_t pcab_tpn_iteration_workspace(int K, int T) {
  size_t kt = (size_t)K * T;
  return al256p(kt * (128 + 512 + 256 + 128 + 8) * 4) + al256p(kt * 4 * 8) + 256;
}

extern "C" int pcab_tpn_iteration(const float* points, const int* inst, const int* tidx, int n, int K, int T,
                                  const float* mos_emb, const float* geo_emb, const float* pack_pos,
                                  const float* pack_regressor, float* pose_out, float* pose_centered_out,
                                  float* rep_out, void* workspace, size_t workspace_bytes, const float* pos_w0_tc,
                                  const float* pos_w1_tc, const float* pos_bias_host, void* pos_scratch,
                                  cudaStream_t stream) {
  PCAB_REQUIRE(workspace_bytes >= pcab_tpn_iteration_workspace(K, T), "workspace too small");
  size_t kt = (size_t)K * T;
  float* f = (float*)workspace;
  float* frame_emb = f;
  f += kt * 128;
  float* X = f;
  f += kt * 512;
  float* H0 = f;
  f += kt * 256;
  float* H1 = f;
  f += kt * 128;
  float* rep = f;
  f += kt * 8;
  double* sums = (double*)((char*)workspace + al256p(kt * (128 + 512 + 256 + 128 + 8) * 4));
  PCAB_CUDA(cudaMemsetAsync(sums, 0, kt * 4 * 8, stream));
  size_t smem = (size_t)(2 * 128 * LDP + mlp::SW_FLOATS) * sizeof(float);
  static PcabSmemOnce once;
  PCAB_CUDA(pcab_set_max_smem(k_tpn_pos_embed, (int)smem, once));
  k_tpn_frame_sums<<<grid_for(n, 256), 256, 0, stream>>>(points, inst, tidx, T, n, sums);
  if (pos_w0_tc) {
    // tensor-core path (csrc/mlp_tc.cu): layer 0 on the CUDA cores, layers 1-2 + the (instance, frame) max on tcgen05
    PCAB_REQUIRE(pos_w1_tc && pos_bias_host && pos_scratch, "tensor-core positional embedding needs its packs and scratch");
    float* rows = (float*)pos_scratch;
    int* seg = (int*)(rows + (size_t)n * 32);
    int rc = pcab_tpn_pos_l0(points, inst, tidx, n, T, sums, pack_pos, rows, seg, stream);
    if (rc != PCAB_OK) return rc;
    rc = pcab_embed_segmax_tc(2, rows, nullptr, seg, n, (int)kt, pos_w0_tc, pos_w1_tc, nullptr, pos_bias_host, frame_emb, stream);
    if (rc != PCAB_OK) return rc;
  } else {
    k_fill<<<grid_for((long long)kt * 128, 256), 256, 0, stream>>>(frame_emb, (long long)kt * 128, -INFINITY);
    k_tpn_pos_embed<<<cdiv(n, PTS), mlp::NT, smem, stream>>>(points, inst, tidx, n, T, sums, pack_pos, frame_emb);
    k_fix_neg_inf<<<grid_for((long long)kt * 128, 256), 256, 0, stream>>>(frame_emb, (long long)kt * 128);
  }
  k_tpn_regressor_input<<<grid_for((long long)kt * 512, 256), 256, 0, stream>>>(geo_emb, mos_emb, frame_emb, (int)kt, T, X);
  const float* W0 = pack_regressor;
  const float* b0 = W0 + 512 * 256;
  const float* s0 = b0 + 256;
  const float* t0 = s0 + 256;
  const float* W1 = t0 + 256;
  const float* b1 = W1 + 256 * 128;
  const float* s1 = b1 + 128;
  const float* t1 = s1 + 128;
  const float* W2 = t1 + 128;
  const float* b2 = W2 + 128 * 7;
  (void)H0, (void)H1;  // (hidden rows of the regressor now stay in shared memory)
  k_regressor<<<cdiv(kt, RR), 256, 0, stream>>>(X, (int)kt, W0, b0, s0, t0, W1, b1, s1, t1, W2, b2, rep_out ? rep_out : rep);
  k_tpn_pose<<<grid_for((long long)kt, 128), 128, 0, stream>>>(rep_out ? rep_out : rep, sums, (int)kt, T,
                                                               pose_centered_out, pose_out);
  PCAB_CHECK_LAUNCH("pcab_tpn_iteration");
  return PCAB_OK;
}

// out[j] = R[seg[j]] p_j + t[seg[j]] with seg = a per-point index into pose[*][4][4]
extern "C" int pcab_apply_seg_pose(const float* points, const int* seg, const float* pose, int n, float* out,
                                   cudaStream_t stream) {
  if (n <= 0) return PCAB_OK;
  k_apply_seg_pose<<<grid_for(n, 256), 256, 0, stream>>>(points, seg, pose, n, out);
  PCAB_CHECK_LAUNCH("pcab_apply_seg_pose");
  return PCAB_OK;
}

namespace {
// TubeNet inputs of the selected points in one pass (replaces six boolean-mask gathers of models/motionnet.py:246-258)
__global__ void k_tpn_gather(const int* __restrict__ idx, int k, const long long* __restrict__ inst, const int* __restrict__ pbatch,
                             const int* __restrict__ ptime, const float* __restrict__ tp, const long long* __restrict__ sd,
                             long long* __restrict__ o_inst, long long* __restrict__ o_batch, long long* __restrict__ o_time,
                             float* __restrict__ o_tp, long long* __restrict__ o_mos) {
  int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += stride) {
    const int i = idx[j];
    o_inst[j] = inst[i], o_batch[j] = pbatch[i], o_time[j] = ptime[i], o_mos[j] = sd[i];
    o_tp[3 * j] = tp[3 * i], o_tp[3 * j + 1] = tp[3 * i + 1], o_tp[3 * j + 2] = tp[3 * i + 2];
  }
}

// general 4x4 inverse in double (Gauss-Jordan, partial pivoting) and G[t] = gt[t] @ inv(est[t])  (models/alignnet.py:9-38)
__global__ void k_pose_error(const float* __restrict__ gt, const float* __restrict__ est, int T, float* __restrict__ out) {
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    double a[4][8];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) a[i][j] = est[16 * t + 4 * i + j], a[i][4 + j] = (i == j);
    for (int c = 0; c < 4; ++c) {
      int piv = c;
      for (int r = c + 1; r < 4; ++r)
        if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
      for (int j = 0; j < 8; ++j) {
        const double tmp = a[c][j];
        a[c][j] = a[piv][j], a[piv][j] = tmp;
      }
      const double d = a[c][c];
      for (int j = 0; j < 8; ++j) a[c][j] /= d;
      for (int r = 0; r < 4; ++r)
        if (r != c) {
          const double f = a[r][c];
          for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double v = 0;
        for (int k = 0; k < 4; ++k) v += (double)gt[16 * t + 4 * i + k] * a[k][4 + j];
        out[16 * t + 4 * i + j] = (float)v;
      }
  }
}

// sums for inst_l2_error / dynamic_inst_l2_error (models/alignnet.py:271-279): acc = {sum l2 w, sum w, sum l2 wm, sum wm}
__global__ void k_inst_errors(const float* __restrict__ a, const float* __restrict__ b, const long long* __restrict__ tidx,
                              const long long* __restrict__ mos, int n, double* __restrict__ acc) {
  double s[4] = {0, 0, 0, 0};
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float dx = a[3 * i] - b[3 * i], dy = a[3 * i + 1] - b[3 * i + 1], dz = a[3 * i + 2] - b[3 * i + 2];
    const float l2 = sqrtf(dx * dx + dy * dy + dz * dz);
    const bool w = tidx[i] > 0, wm = w && mos[i] == 1;
    if (w) s[0] += l2, s[1] += 1.0;
    if (wm) s[2] += l2, s[3] += 1.0;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    s[k] = warp_sum_d(s[k]);
    if ((threadIdx.x & 31) == 0 && s[k] != 0.0) atomicAdd(acc + k, s[k]);
  }
}
__global__ void k_inst_errors_final(const double* __restrict__ acc, float* __restrict__ out2) {
  out2[0] = (float)(acc[0] / (acc[1] + 1e-20));
  out2[1] = (float)(acc[2] / (acc[3] + 1e-20));
}
}  // namespace

extern "C" int pcab_tpn_gather(const int* idx, int k, const long long* inst, const int* point_batch, const int* point_time,
                               const float* points, const long long* sd_labels, long long* inst_out, long long* batch_out,
                               long long* time_out, float* points_out, long long* mos_out, cudaStream_t stream) {
  if (k <= 0) return PCAB_OK;
  k_tpn_gather<<<grid_for(k, 256), 256, 0, stream>>>(idx, k, inst, point_batch, point_time, points, sd_labels, inst_out, batch_out,
                                                     time_out, points_out, mos_out);
  PCAB_CHECK_LAUNCH("pcab_tpn_gather");
  return PCAB_OK;
}

extern "C" int pcab_pose_error(const float* pose_gt, const float* pose_est, int T, float* out, cudaStream_t stream) {
  k_pose_error<<<1, 32, 0, stream>>>(pose_gt, pose_est, T, out);
  PCAB_CHECK_LAUNCH("pcab_pose_error");
  return PCAB_OK;
}

extern "C" int pcab_inst_errors(const float* rec_est, const float* rec_gt, const long long* time_idx, const long long* mos_labels, int n,
                                double* scratch4, float* out2, cudaStream_t stream) {
  PCAB_CUDA(cudaMemsetAsync(scratch4, 0, 4 * sizeof(double), stream));
  if (n > 0) k_inst_errors<<<grid_for(n, 256), 256, 0, stream>>>(rec_est, rec_gt, time_idx, mos_labels, n, scratch4);
  k_inst_errors_final<<<1, 1, 0, stream>>>(scratch4, out2);
  PCAB_CHECK_LAUNCH("pcab_inst_errors");
  return PCAB_OK;
}

extern "C" int pcab_scatter_rows3(const float* src, const int* idx, int k, float* dst, cudaStream_t stream) {
  if (k <= 0) return PCAB_OK;
  k_scatter_rows3<<<grid_for(k, 256), 256, 0, stream>>>(src, idx, k, dst);
  PCAB_CHECK_LAUNCH("pcab_scatter_rows3");
  return PCAB_OK;
}
