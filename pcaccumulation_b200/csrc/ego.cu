// Ego-motion head: background-pillar compaction, keypoint gather (+L2 normalisation), feature
// affinity, log-domain Sinkhorn with slack, soft correspondences, weighted Kabsch (3x3 Jacobi SVD),
// sequence pose assembly, point losses and pose errors.
//
// Replaces models/egomotion.py:100-137 (sinkhorn), :140-192 (pairwise_ego_motion_estimation),
// :195-357 (sequence strategies), :387-469 (forward), toolbox/utils.py:125-144 (square_distance),
// toolbox/register_utils.py:247-318 (kabsch), :184-197 (relative pose), :19-56 (errors) and
// models/motionnet.py:199 (feature normalisation, applied at gather time to the sampled rows only).
//
// All pairs of all scenes are processed by the same launches (batched over pairs).
// Sinkhorn never rewrites the 1024x1024 matrix: after any number of row/column normalisations the
// padded matrix equals  A_pad[i][j] - r_i - c_j  (r = 0 on the slack row, c = 0 on the slack column), so
// one iteration is  r_i = logsumexp_j(A_pad[i][j] - c_j)  followed by  c_j = logsumexp_i(A_pad[i][j] - r_i).
#include <cub/cub.cuh>
#include "common.cuh"
#include "svd3.cuh"
#include "pair16.cuh"
#include "pcab200.h"

namespace {

constexpr int KP = 1024;  // keypoints per side (pose_estimation.n_kpts)
constexpr int FD = 64;    // feature dim (pose_estimation.feats_dim)

__global__ void k_bg_flag(const int* __restrict__ cell2pillar, const int* __restrict__ fb_est, long long ncell,
                          int* __restrict__ flag) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ncell; i += stride)
    flag[i] = (cell2pillar[i] >= 0 && fb_est[i] == 0) ? 1 : 0;
}

__global__ void k_bg_compact(const int* __restrict__ flag, const int* __restrict__ pos, long long ncell, int hw,
                             int* __restrict__ bg_cells, int* __restrict__ frame_off, int nframes) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ncell; i += stride) {
    if (flag[i]) bg_cells[pos[i]] = (int)i;
    if (i % hw == 0) frame_off[i / hw] = pos[i];
    if (i == ncell - 1) frame_off[nframes] = pos[i] + flag[i];
  }
}

// gather 1024 keypoints per (pair, side): coordinates = pillar mean, features = L2-normalised head output
template <bool P16>
__global__ void k_ego_gather(const float* __restrict__ geo, const int* __restrict__ cell2pillar,
                             const float* __restrict__ pillar_mean, const int* __restrict__ bg_cells,
                             const int* __restrict__ frame_off, const int* __restrict__ pair_frames,  // [P][2] src,tgt
                             const int* __restrict__ choice,                                       // [P][2][KP]
                             int npairs, float* __restrict__ feats, float* __restrict__ xyz) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < npairs * 2 * KP; r += nwarp) {
    int ps = r / KP;
    int frame = pair_frames[ps];
    int cell = bg_cells[frame_off[frame] + choice[r]];
    float a, b;
    if (P16) {
      a = p16::load1(geo, (size_t)cell, FD, lane), b = p16::load1(geo, (size_t)cell, FD, lane + 32);
    } else {
      const float* g = geo + (size_t)cell * FD;
      a = g[lane], b = g[lane + 32];
    }
    float ss = warp_sum(a * a + b * b);
    float nrm = sqrtf(ss);
    feats[(size_t)r * FD + lane] = a / nrm;
    feats[(size_t)r * FD + lane + 32] = b / nrm;
    if (lane < 3) xyz[(size_t)r * 3 + lane] = pillar_mean[(size_t)cell2pillar[cell] * 3 + lane];
  }
}

// A[p][i][j] = -(clamp(2 - 2 <fs_i, ft_j>, 1e-12) - softplus(alpha)) / (exp(beta) + 0.02)
__global__ void __launch_bounds__(256) k_ego_affinity(const float* __restrict__ feats, const float* __restrict__ alpha,
                                                      const float* __restrict__ beta, float* __restrict__ A) {
  __shared__ float fs[64][FD + 1], ft[64][FD + 1];
  int p = blockIdx.z;
  const float* S = feats + (size_t)(p * 2 + 0) * KP * FD + (size_t)blockIdx.y * 64 * FD;
  const float* T = feats + (size_t)(p * 2 + 1) * KP * FD + (size_t)blockIdx.x * 64 * FD;
  for (int e = threadIdx.x; e < 64 * FD; e += 256) {
    fs[e / FD][e % FD] = S[e];
    ft[e / FD][e % FD] = T[e];
  }
  __syncthreads();
  int ti = threadIdx.x / 16, tj = threadIdx.x % 16;
  float acc[4][4] = {};
#pragma unroll 8
  for (int k = 0; k < FD; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = fs[ti + 16 * u][k], b[u] = ft[tj + 16 * u][k];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
  }
  float al = *alpha, be = *beta;
  float sp = al > 20.f ? al : log1pf(expf(al));  // torch Softplus (beta=1, threshold=20)
  float den = expf(be) + 0.02f;
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      int i = blockIdx.y * 64 + ti + 16 * u, j = blockIdx.x * 64 + tj + 16 * v;
      const float raw = -2.f * acc[u][v] + 2.f;
      const float d = raw != raw ? raw : fmaxf(raw, 1e-12f);  // torch.clamp keeps NaN (all-zero feature rows: 0/0), fmaxf would drop it
      A[((size_t)p * KP + i) * KP + j] = -(d - sp) / den;
    }
}

// r_i = logsumexp over j in [0,K] of (A_pad[i][j] - c_j); one warp per row
__global__ void k_sinkhorn_rows(const float* __restrict__ A, const float* __restrict__ c, float* __restrict__ r) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;  // row in [0, P*KP)
  int p = row / KP;
  const float* a = A + (size_t)row * KP;
  const float* cc = c + (size_t)p * KP;
  float v[KP / 32];
  float mx = 0.f;  // slack column entry: 0 - 0
#pragma unroll
  for (int q = 0; q < KP / 32; ++q) {
    v[q] = a[q * 32 + lane] - cc[q * 32 + lane];
    mx = fmaxf(mx, v[q]);
  }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < KP / 32; ++q) s += expf(v[q] - mx);
  s = warp_sum(s) + expf(0.f - mx);
  if (lane == 0) r[row] = mx + logf(s);
}

// c_j = logsumexp over i in [0,J] of (A_pad[i][j] - r_i); block = 32 columns x 32 row slices
__global__ void __launch_bounds__(1024) k_sinkhorn_cols(const float* __restrict__ A, const float* __restrict__ r,
                                                        float* __restrict__ c) {
  __shared__ float sm[32][33], ss[32][33];
  int p = blockIdx.y;
  int j = blockIdx.x * 32 + threadIdx.x;
  int slice = threadIdx.y;
  const float* a = A + (size_t)p * KP * KP;
  const float* rr = r + (size_t)p * KP;
  float mx = -INFINITY, s = 0.f;
  for (int i = slice; i < KP; i += 32) {
    float v = a[(size_t)i * KP + j] - rr[i];
    if (v > mx) {
      s = s * expf(mx - v) + 1.f;
      mx = v;
    } else {
      s += expf(v - mx);
    }
  }
  sm[slice][threadIdx.x] = mx;
  ss[slice][threadIdx.x] = s;
  __syncthreads();
  if (slice == 0) {
    float M = 0.f;  // slack row entry
    for (int q = 0; q < 32; ++q) M = fmaxf(M, sm[q][threadIdx.x]);
    float S = expf(0.f - M);
    for (int q = 0; q < 32; ++q) S += ss[q][threadIdx.x] * expf(sm[q][threadIdx.x] - M);
    c[(size_t)p * KP + j] = M + logf(S);
  }
}

// perm = exp(A - r_i - c_j) * support ; w_i = sum_j perm ; xhat_i = perm @ xt / (w_i + 1e-20). One warp per row.
__global__ void k_ego_perm(const float* __restrict__ A, const float* __restrict__ r, const float* __restrict__ c,
                           const float* __restrict__ xyz, const float* __restrict__ thr2, float* __restrict__ perm,
                           float* __restrict__ w, float* __restrict__ xhat) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int p = row / KP, i = row % KP;
  const float* xs = xyz + ((size_t)(p * 2 + 0) * KP + i) * 3;
  const float* xt = xyz + (size_t)(p * 2 + 1) * KP * 3;
  float sx = xs[0], sy = xs[1], sz = xs[2];
  float ss = sx * sx + sy * sy + sz * sz;
  float ri = r[row], t2 = thr2[p];
  float ws = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
  for (int j = lane; j < KP; j += 32) {
    float tx = xt[3 * j], ty = xt[3 * j + 1], tz = xt[3 * j + 2];
    float dot = fmaf(sz, tz, fmaf(sy, ty, sx * tx));
    float d = -2.f * dot;
    d += ss;
    d += tx * tx + ty * ty + tz * tz;
    d = fmaxf(d, 1e-12f);
    const float lp = A[(size_t)row * KP + j] - ri - c[(size_t)p * KP + j];
    float pv = (d < t2) ? expf(lp) : (lp != lp ? lp : 0.f);  // exp(.) * support: a NaN stays a NaN outside the support too
    perm[(size_t)row * KP + j] = pv;
    ws += pv, ax = fmaf(pv, tx, ax), ay = fmaf(pv, ty, ay), az = fmaf(pv, tz, az);
  }
  ws = warp_sum(ws), ax = warp_sum(ax), ay = warp_sum(ay), az = warp_sum(az);
  if (lane == 0) {
    w[row] = ws;
    float den = ws + 1e-20f;
    xhat[3 * row] = ax / den, xhat[3 * row + 1] = ay / den, xhat[3 * row + 2] = az / den;
  }
}

__device__ double block_sum(double v, double* sh) {
  v = warp_sum_d(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double t = 0;
  for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += sh[q];
  return t;
}

// weighted Kabsch per pair; one block of 1024 threads per pair (toolbox/register_utils.py:268-313)
__global__ void __launch_bounds__(1024) k_ego_kabsch(const float* __restrict__ xyz, const float* __restrict__ xhat,
                                                     const float* __restrict__ w, float* __restrict__ pose) {
  __shared__ double sh[32];
  __shared__ double mean[6];
  int p = blockIdx.x, i = threadIdx.x;
  const float* x1 = xyz + ((size_t)(p * 2 + 0) * KP + i) * 3;
  const float* x2 = xhat + ((size_t)p * KP + i) * 3;
  double wi = w[(size_t)p * KP + i];
  double wsum = block_sum(wi, sh);
  double wn = (double)((float)wi / ((float)wsum + 1e-7f));
  double sn = block_sum(wn, sh);
  double den = (double)((float)sn + 1e-7f);
  double v[6] = {x1[0], x1[1], x1[2], x2[0], x2[1], x2[2]};
  for (int k = 0; k < 6; ++k) {
    double t = block_sum(wn * v[k], sh);
    if (i == 0) mean[k] = (double)(float)(t / den);
  }
  __syncthreads();
  double a[3] = {(double)(float)(v[0] - mean[0]), (double)(float)(v[1] - mean[1]), (double)(float)(v[2] - mean[2])};
  double b[3] = {(double)(float)(v[3] - mean[3]), (double)(float)(v[4] - mean[4]), (double)(float)(v[5] - mean[5])};
  double C[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) C[r][c] = block_sum(a[r] * wn * b[c], sh);
  if (i == 0) {
    double R[3][3];
    float* P = pose + (size_t)p * 16;
    bool finite = true;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) finite &= isfinite(C[r][c]);
    if (!finite) {  // torch.svd raises on a non-finite covariance: R = I, t = 0 (toolbox/register_utils.py:295-304)
      for (int k = 0; k < 16; ++k) P[k] = (k % 5 == 0) ? 1.f : 0.f;
      return;
    }
    kabsch_rotation(C, R);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) P[4 * r + c] = (float)R[r][c];
      double t = mean[3 + r] - ((double)(float)R[r][0] * mean[0] + (double)(float)R[r][1] * mean[1] +
                                (double)(float)R[r][2] * mean[2]);
      P[4 * r + 3] = (float)t;
    }
    P[12] = P[13] = P[14] = 0.f, P[15] = 1.f;
  }
}

__device__ void inv4(const double* M, double* out) {  // general 4x4 inverse (Gauss-Jordan, partial pivoting)
  double a[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) a[i][j] = M[4 * i + j], a[i][4 + j] = (i == j);
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    for (int r = c + 1; r < 4; ++r)
      if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
    for (int j = 0; j < 8; ++j) {
      double t = a[c][j];
      a[c][j] = a[piv][j], a[piv][j] = t;
    }
    double d = a[c][c];
    for (int j = 0; j < 8; ++j) a[c][j] /= d;
    for (int r = 0; r < 4; ++r)
      if (r != c) {
        double f = a[r][c];
        for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
      }
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) out[4 * i + j] = a[i][4 + j];
}

__device__ void mm4(const double* A, const double* B, double* C) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
      C[4 * i + j] = s;
    }
}

// GT relative pose per pair: inv(gt[anchor]) @ gt[ref]  (float32 output like torch.linalg.solve on f32 inputs)
__global__ void k_ego_pose_gt(const float* __restrict__ ego_gt, const int* __restrict__ pair_frames, int npairs,
                              float* __restrict__ pose_gt) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  double S[16], Tg[16], Ti[16], M[16];
  for (int k = 0; k < 16; ++k) S[k] = ego_gt[(size_t)pair_frames[2 * p] * 16 + k], Tg[k] = ego_gt[(size_t)pair_frames[2 * p + 1] * 16 + k];
  inv4(Tg, Ti);
  mm4(Ti, S, M);
  for (int k = 0; k < 16; ++k) pose_gt[(size_t)p * 16 + k] = (float)M[k];
}

// l1 / l2 point losses over ALL occupied pillars of the pair's source frame (models/egomotion.py:342-348)
__global__ void k_ego_losses(const float* __restrict__ pillar_mean, const int* __restrict__ pillar_frame, int m,
                             const int* __restrict__ pair_frames, int npairs, const float* __restrict__ pose,
                             const float* __restrict__ pose_gt, double* __restrict__ acc /* [P][3] l1,l2,count */) {
  int p = blockIdx.y;
  int frame = pair_frames[2 * p];
  const float* E = pose + (size_t)p * 16;
  const float* G = pose_gt + (size_t)p * 16;
  double l1 = 0, l2 = 0, cnt = 0;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < m; q += gridDim.x * blockDim.x) {
    if (pillar_frame[q] != frame) continue;
    float x = pillar_mean[3 * q], y = pillar_mean[3 * q + 1], z = pillar_mean[3 * q + 2];
    float d[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float e = E[4 * r] * x + E[4 * r + 1] * y + E[4 * r + 2] * z + E[4 * r + 3];
      float g = G[4 * r] * x + G[4 * r + 1] * y + G[4 * r + 2] * z + G[4 * r + 3];
      d[r] = e - g;
    }
    l1 += fabsf(d[0]) + fabsf(d[1]) + fabsf(d[2]);
    l2 += sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    cnt += 1;
  }
  l1 = warp_sum_d(l1), l2 = warp_sum_d(l2), cnt = warp_sum_d(cnt);
  if ((threadIdx.x & 31) == 0 && cnt > 0) {
    atomicAdd(acc + 3 * p, l1), atomicAdd(acc + 3 * p + 1, l2), atomicAdd(acc + 3 * p + 2, cnt);
  }
}

// rotation error (toolbox/register_utils.py:19-42): the angle of R_est^T R_gt in degrees.  The reference takes
// acos((trace - 1) / 2) in float32, which loses half of its digits for the sub-degree angles that occur here; the same
// angle is evaluated as atan2(|axis part|, (trace - 1) / 2) in double, i.e. at least as close to the exact value.
// Translation error (:45-56): |t_est - t_gt|.
__device__ void pose_errors(const float* E, const float* G, double& rot_sum, double& trans_sum) {
  double Rr[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += (double)E[4 * k + i] * (double)G[4 * k + j];
      Rr[3 * i + j] = s;
    }
  const double cs = fmin(fmax((Rr[0] + Rr[4] + Rr[8] - 1.0) * 0.5, -1.0), 1.0);
  const double ax = Rr[7] - Rr[5], ay = Rr[2] - Rr[6], az = Rr[3] - Rr[1];
  rot_sum += 180.0 * atan2(0.5 * sqrt(ax * ax + ay * ay + az * az), cs) / 3.14159265358979323846;
  float dx = E[3] - G[3], dy = E[7] - G[7], dz = E[11] - G[11];
  trans_sum += (double)sqrtf(dx * dx + dy * dy + dz * dz);
}

// the two error scalars of models/egomotion.py:451-455 recomputed for refined poses (ICP branch, :439-441)
__global__ void k_ego_errors(const float* __restrict__ est, const float* __restrict__ gt, int B, int T, float* __restrict__ out2) {
  if (threadIdx.x || blockIdx.x) return;
  double rot_sum = 0, trans_sum = 0;
  for (int f = 0; f < B * T; ++f) pose_errors(est + (size_t)f * 16, gt + (size_t)f * 16, rot_sum, trans_sum);
  const double n = T;
  out2[0] = (float)(rot_sum / (B * T) * n / (n - 1));
  out2[1] = (float)(trans_sum / (B * T) * n / (n - 1));
}

// sequence assembly + errors; single thread (B*T tiny 4x4 products)
__global__ void k_ego_finalize(const float* __restrict__ pose, const float* __restrict__ ego_gt,
                               const int* __restrict__ chain_pair /* [B*T], -1 for t=0 */, int B, int T, int chain_mode,
                               const double* __restrict__ acc, int npairs, float* __restrict__ est_out,
                               float* __restrict__ gt_out, float* __restrict__ scalars /* l1,l2,rot,trans */) {
  if (threadIdx.x || blockIdx.x) return;
  double rot_sum = 0, trans_sum = 0;
  for (int b = 0; b < B; ++b) {
    double chain[16];
    for (int t = 0; t < T; ++t) {
      float* E = est_out + (size_t)(b * T + t) * 16;
      float* G = gt_out + (size_t)(b * T + t) * 16;
      if (t == 0) {
        for (int k = 0; k < 16; ++k) E[k] = G[k] = (k % 5 == 0) ? 1.f : 0.f, chain[k] = (k % 5 == 0);
      } else {
        const float* P = pose + (size_t)chain_pair[b * T + t] * 16;
        if (chain_mode) {
          float Pf[16], Cf[16];
          for (int k = 0; k < 16; ++k) Cf[k] = (float)chain[k];
          for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
              float s = 0.f;
              for (int k = 0; k < 4; ++k) s = fmaf(Cf[4 * i + k], P[4 * k + j], s);
              Pf[4 * i + j] = s;
            }
          for (int k = 0; k < 16; ++k) chain[k] = Pf[k], E[k] = Pf[k];
        } else {
          for (int k = 0; k < 16; ++k) E[k] = P[k];
        }
        double S[16], Tg[16], Ti[16], M[16];
        for (int k = 0; k < 16; ++k) S[k] = ego_gt[(size_t)(b * T + t) * 16 + k], Tg[k] = ego_gt[(size_t)(b * T) * 16 + k];
        inv4(Tg, Ti);
        mm4(Ti, S, M);
        for (int k = 0; k < 16; ++k) G[k] = (float)M[k];
      }
      pose_errors(E, G, rot_sum, trans_sum);
    }
  }
  double l1 = 0, l2 = 0;
  for (int p = 0; p < npairs; ++p) {
    l1 += (double)(float)(acc[3 * p] / acc[3 * p + 2]);
    l2 += (double)(float)(acc[3 * p + 1] / acc[3 * p + 2]);
  }
  scalars[0] = (float)(l1 / npairs);
  scalars[1] = (float)(l2 / npairs);
  double n = T;
  scalars[2] = (float)(rot_sum / (B * T) * n / (n - 1));
  scalars[3] = (float)(trans_sum / (B * T) * n / (n - 1));
}

size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" int pcab_ego_pose_errors(const float* ego_est, const float* ego_gt, int B, int T, float* out2, cudaStream_t stream) {
  PCAB_REQUIRE(B > 0 && T > 1, "bad sizes");
  k_ego_errors<<<1, 32, 0, stream>>>(ego_est, ego_gt, B, T, out2);
  PCAB_CHECK_LAUNCH("pcab_ego_pose_errors");
  return PCAB_OK;
}

extern "C" size_t pcab_bg_compact_workspace(long long n_cells) {
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int*)nullptr, (int*)nullptr, (int)n_cells);
  return 2 * al256((size_t)n_cells * 4) + al256(scan_bytes) + 256;
}

// bg_cells: compacted (cell-order) list of occupied background cells; frame_off[f..f+1] = its range for frame f
extern "C" int pcab_bg_compact(const int* cell_to_pillar, const int* fb_est, int n_frames, int H, int W, int* bg_cells,
                               int* frame_off, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  long long ncell = (long long)n_frames * H * W;
  PCAB_REQUIRE(workspace_bytes >= pcab_bg_compact_workspace(ncell), "workspace too small");
  char* w = (char*)workspace;
  int* flag = (int*)w;
  w += al256((size_t)ncell * 4);
  int* pos = (int*)w;
  w += al256((size_t)ncell * 4);
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flag, pos, (int)ncell);
  k_bg_flag<<<grid_for(ncell, 256), 256, 0, stream>>>(cell_to_pillar, fb_est, ncell, flag);
  PCAB_CUDA(cub::DeviceScan::ExclusiveSum(w, scan_bytes, flag, pos, (int)ncell, stream));
  k_bg_compact<<<grid_for(ncell, 256), 256, 0, stream>>>(flag, pos, ncell, H * W, bg_cells, frame_off, n_frames);
  PCAB_CHECK_LAUNCH("pcab_bg_compact");
  return PCAB_OK;
}

// floats: feats P*2*KP*64 | xyz P*2*KP*3 | A P*KP*KP | r P*KP | c P*KP | w P*KP | xhat P*KP*3 | pose_gt P*16 ; doubles acc P*3
extern "C" size_t pcab_ego_pairs_workspace(int npairs) {
  size_t f = (size_t)npairs * (2 * KP * FD + 2 * KP * 3 + (size_t)KP * KP + 3 * KP + 3 * KP + 16);
  return al256(f * 4) + al256((size_t)npairs * 3 * 8) + 256;
}

extern "C" int pcab_ego_pairs(const float* geo_nhwc, int geo_fmt, const int* cell_to_pillar, const float* pillar_mean,
                              const int* pillar_frame, int n_pillars, const int* bg_cells, const int* frame_off,
                              const int* pair_frames, const int* choice, const float* thr2, int npairs,
                              const float* alpha, const float* beta, int sinkhorn_iters, const float* ego_gt,
                              const int* chain_pair, int B, int T, int chain_mode, float* perm_out, float* pose_pairs,
                              float* ego_est, float* ego_gt_out, float* scalars, void* workspace, size_t workspace_bytes,
                              cudaStream_t stream) {
  PCAB_REQUIRE(workspace_bytes >= pcab_ego_pairs_workspace(npairs), "workspace too small");
  PCAB_REQUIRE(npairs > 0, "no pairs");
  float* f = (float*)workspace;
  float* feats = f;
  f += (size_t)npairs * 2 * KP * FD;
  float* xyz = f;
  f += (size_t)npairs * 2 * KP * 3;
  float* A = f;
  f += (size_t)npairs * KP * KP;
  float* r = f;
  f += (size_t)npairs * KP;
  float* c = f;
  f += (size_t)npairs * KP;
  float* w = f;
  f += (size_t)npairs * KP;
  float* xhat = f;
  f += (size_t)npairs * KP * 3;
  float* pose_gt = f;
  f += (size_t)npairs * 16;
  size_t fbytes = al256((size_t)((char*)f - (char*)workspace));
  double* acc = (double*)((char*)workspace + fbytes);

  int rows = npairs * 2 * KP;
  if (geo_fmt)
    k_ego_gather<true><<<cdiv((long long)rows * 32, 256), 256, 0, stream>>>(geo_nhwc, cell_to_pillar, pillar_mean, bg_cells, frame_off,
                                                                           pair_frames, choice, npairs, feats, xyz);
  else
    k_ego_gather<false><<<cdiv((long long)rows * 32, 256), 256, 0, stream>>>(geo_nhwc, cell_to_pillar, pillar_mean, bg_cells, frame_off,
                                                                            pair_frames, choice, npairs, feats, xyz);
  k_ego_affinity<<<dim3(KP / 64, KP / 64, npairs), 256, 0, stream>>>(feats, alpha, beta, A);
  PCAB_CUDA(cudaMemsetAsync(c, 0, (size_t)npairs * KP * 4, stream));
  for (int it = 0; it < sinkhorn_iters; ++it) {
    k_sinkhorn_rows<<<npairs * KP / 8, 256, 0, stream>>>(A, c, r);
    k_sinkhorn_cols<<<dim3(KP / 32, npairs), dim3(32, 32), 0, stream>>>(A, r, c);
  }
  if (sinkhorn_iters == 0) PCAB_CUDA(cudaMemsetAsync(r, 0, (size_t)npairs * KP * 4, stream));
  k_ego_perm<<<npairs * KP / 8, 256, 0, stream>>>(A, r, c, xyz, thr2, perm_out, w, xhat);
  k_ego_kabsch<<<npairs, 1024, 0, stream>>>(xyz, xhat, w, pose_pairs);
  k_ego_pose_gt<<<cdiv(npairs, 32), 32, 0, stream>>>(ego_gt, pair_frames, npairs, pose_gt);
  PCAB_CUDA(cudaMemsetAsync(acc, 0, (size_t)npairs * 3 * 8, stream));
  k_ego_losses<<<dim3(pcab_sm_count(), npairs), 256, 0, stream>>>(pillar_mean, pillar_frame, n_pillars, pair_frames, npairs,
                                                      pose_pairs, pose_gt, acc);
  k_ego_finalize<<<1, 32, 0, stream>>>(pose_pairs, ego_gt, chain_pair, B, T, chain_mode, acc, npairs, ego_est,
                                       ego_gt_out, scalars);
  PCAB_CHECK_LAUNCH("pcab_ego_pairs");
  return PCAB_OK;
}
