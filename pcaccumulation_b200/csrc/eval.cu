// Evaluation tail of the test loop on the device: per-point scene-flow errors against the
// ground-truth accumulation, the scene-flow accuracy counters and the motion-segmentation IoU counters, in ONE pass over
// the points.
//
// Replaces libs/tester.py:58-88 (ego_motion_compensation + reconstruct_sequence with the GT instance motions, EPE /
// relative error per point, `time_indice > 0` selection), toolbox/register_utils.py:59-93, the threshold counters of
// toolbox/sf_eval_utils.py:46-52,71-100 (Acc3DS / Acc3DR / Outlier / ROutlier, per category) and libs/loss.py:17-48,
// 139-149 (compute_iou on the FG-masked motion labels).  The reference does this with a dozen full-size torch ops, five
// host copies per scene and Python loops; here it is 52 B read + 8 B written per point.
#include "common.cuh"
#include "pcab200.h"

namespace {

constexpr int kCat = 3;   // 0: all points with t > 0, 1: dynamic (sd == 1), 2: "static" (fb == 1), as collect_scene_stats does
constexpr int kSf = 6;    // count, sum epe, Acc3DS, Acc3DR, Outlier, ROutlier
constexpr int kMos = 8;   // class 0/1: intersection, pred positives, gt positives ; [6] = masked points, [7] = rows with an out-of-range frame / instance index

struct EvalArgs {
  const float* pts;
  const int* tidx;
  const float* rec;
  const float* ego_gt;        // [T,4,4]
  const long long* inst;      // [N]
  const float* inst_gt;       // [K,T,4,4]
  const long long* fb_gt;
  const long long* sd_gt;
  const float* mos_est;       // [N,2]
  const long long* fb_est;    // [N]
  int n, T, K;
  float* epe;
  float* rel;
  double* sf;                 // [kCat][kSf]
  long long* mos;             // [kMos]
};

__device__ __forceinline__ void apply(const float* __restrict__ m, float x, float y, float z, float& ox, float& oy, float& oz) {
  // (R p) + t with the products summed in index order, as a [3x3]x[3x1] matmul does
  ox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z)), m[3]);
  oy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[4], x), __fmul_rn(m[5], y)), __fmul_rn(m[6], z)), m[7]);
  oz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[8], x), __fmul_rn(m[9], y)), __fmul_rn(m[10], z)), m[11]);
}

__global__ void __launch_bounds__(256) k_flow_eval(EvalArgs a) {
  __shared__ double s_sf[kCat * kSf];
  __shared__ unsigned long long s_mos[kMos];
  for (int i = threadIdx.x; i < kCat * kSf; i += blockDim.x) s_sf[i] = 0.0;
  for (int i = threadIdx.x; i < kMos; i += blockDim.x) s_mos[i] = 0ull;
  __syncthreads();
  float sf[kCat][kSf];
  int mos[kMos];
#pragma unroll
  for (int c = 0; c < kCat; ++c)
#pragma unroll
    for (int q = 0; q < kSf; ++q) sf[c][q] = 0.f;
#pragma unroll
  for (int q = 0; q < kMos; ++q) mos[q] = 0;
  double epe_sum[kCat] = {0.0, 0.0, 0.0};
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    const float x = a.pts[3 * i], y = a.pts[3 * i + 1], z = a.pts[3 * i + 2];
    const int t = a.tidx[i];
    float ex, ey, ez, gx, gy, gz;
    const long long k = a.inst[i];
    if (t < 0 || t >= a.T || k < 0 || k >= a.K) {
      // the reference's gathers (toolbox/register_utils.py:66,85) raise an IndexError here: count the row (the host side
      // raises from summary()) instead of reading out of bounds or silently evaluating against the wrong motion
      mos[7] += 1;
      a.epe[i] = a.rel[i] = __int_as_float(0x7fc00000);
      continue;
    }
    apply(a.ego_gt + (size_t)t * 16, x, y, z, ex, ey, ez);
    apply(a.inst_gt + ((size_t)k * a.T + t) * 16, ex, ey, ez, gx, gy, gz);
    const float fx = __fsub_rn(__fsub_rn(a.rec[3 * i], x), __fsub_rn(gx, x));
    const float fy = __fsub_rn(__fsub_rn(a.rec[3 * i + 1], y), __fsub_rn(gy, y));
    const float fz = __fsub_rn(__fsub_rn(a.rec[3 * i + 2], z), __fsub_rn(gz, z));
    const float epe = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)), __fmul_rn(fz, fz)));
    const float mx = __fsub_rn(gx, x), my = __fsub_rn(gy, y), mz = __fsub_rn(gz, z);
    const float mag = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my)), __fmul_rn(mz, mz)));
    const float rel = __fdiv_rn(epe, __fadd_rn(mag, 1e-20f));
    a.epe[i] = epe, a.rel[i] = rel;
    const bool sd = a.sd_gt[i] == 1, fb = a.fb_gt[i] == 1;
    if (t > 0) {
      const bool in_cat[kCat] = {true, sd, fb};
      const float s = (epe < 0.05f || rel < 0.05f) ? 1.f : 0.f, r = (epe < 0.1f || rel < 0.1f) ? 1.f : 0.f;
      const float o = (epe > 0.3f || rel > 0.1f) ? 1.f : 0.f, ro = (epe > 0.3f && rel > 0.3f) ? 1.f : 0.f;
#pragma unroll
      for (int c = 0; c < kCat; ++c)
        if (in_cat[c]) {
          sf[c][0] += 1.f, sf[c][2] += s, sf[c][3] += r, sf[c][4] += o, sf[c][5] += ro;
          epe_sum[c] += (double)epe;
        }
    }
    // motion segmentation IoU on the points that are foreground in the labels or in the prediction (libs/loss.py:144-149)
    if (fb || a.fb_est[i] == 1) {
      const int pred = a.mos_est[2 * i + 1] > a.mos_est[2 * i] ? 1 : 0;  // argmax, first maximum wins
      const int gt = (int)a.sd_gt[i];
      mos[6] += 1;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        mos[3 * c + 0] += (pred == c && gt == c);
        mos[3 * c + 1] += (pred == c);
        mos[3 * c + 2] += (gt == c);
      }
    }
  }
  // per-thread counts are small integers held exactly in float; reduce per warp, then per block, then one atomic per block
#pragma unroll
  for (int c = 0; c < kCat; ++c) {
#pragma unroll
    for (int q = 0; q < kSf; ++q) {
      double v = q == 1 ? epe_sum[c] : (double)sf[c][q];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(&s_sf[c * kSf + q], v);
    }
  }
#pragma unroll
  for (int q = 0; q < kMos; ++q) {
    int v = mos[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_mos[q], (unsigned long long)v);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kCat * kSf; i += blockDim.x)
    if (s_sf[i] != 0.0) atomicAdd(a.sf + i, s_sf[i]);
  for (int i = threadIdx.x; i < kMos; i += blockDim.x)
    if (s_mos[i]) atomicAdd(reinterpret_cast<unsigned long long*>(a.mos) + i, s_mos[i]);
}

}  // namespace

// sf_counters [3 categories][6] doubles (count, sum EPE, Acc3DS, Acc3DR, Outlier, ROutlier counts) and mos_counters [8] int64
// (class 0: intersection, predicted, labelled; class 1: the same; masked points; unused) are ACCUMULATED: zero them before
// the first scene.  epe / rel: per point.  One scene per call (instance labels index inst_motion_gt [K,T,4,4]).
extern "C" int pcab_flow_eval(const float* input_points, const int* time_idx, const float* rec_est, const float* ego_motion_gt,
                              const long long* inst_labels, const float* inst_motion_gt, int n_instances,
                              const long long* fb_labels, const long long* sd_labels, const float* mos_est,
                              const long long* fb_est_per_point, int n_points, int n_frames, float* epe_out, float* rel_out,
                              double* sf_counters, long long* mos_counters, cudaStream_t stream) {
  if (n_points <= 0) return PCAB_OK;
  PCAB_REQUIRE(n_instances >= 1 && n_frames >= 1, "at least one (background) instance and one frame");
  EvalArgs a;
  a.pts = input_points, a.tidx = time_idx, a.rec = rec_est, a.ego_gt = ego_motion_gt, a.inst = inst_labels;
  a.inst_gt = inst_motion_gt, a.fb_gt = fb_labels, a.sd_gt = sd_labels, a.mos_est = mos_est, a.fb_est = fb_est_per_point;
  a.n = n_points, a.T = n_frames, a.K = n_instances, a.epe = epe_out, a.rel = rel_out, a.sf = sf_counters, a.mos = mos_counters;
  k_flow_eval<<<grid_for(n_points, 256, 8), 256, 0, stream>>>(a);
  PCAB_CHECK_LAUNCH("pcab_flow_eval");
  return PCAB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Instance-segmentation part of the evaluation tail: toolbox/cluster_eval.py:71-152 (ClusterEvaluation.forward, adopted
// there from ASIS).  The reference builds one boolean mask per instance and loops over all (gt, est) pairs in Python with a
// device sync per pair; here one pass over the points fills the (est, gt) contingency table, and one small block turns it
// into the coverage and precision / recall counters.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

__global__ void k_contingency(const long long* __restrict__ est, const long long* __restrict__ gt,
                              const long long* __restrict__ mos, int n, int E, int G, int* __restrict__ table,
                              int* __restrict__ est_size, int* __restrict__ est_mos, int* __restrict__ gt_size,
                              int* __restrict__ gt_mos) {
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int e = (int)est[i], g = (int)gt[i], m = mos[i] != 0;
    if (e > 0 && e <= E) atomicAdd(est_size + e, 1), atomicAdd(est_mos + e, m);
    if (g > 0 && g <= G) atomicAdd(gt_size + g, 1), atomicAdd(gt_mos + g, m);
    if (e > 0 && e <= E && g > 0 && g <= G) atomicAdd(table + (size_t)e * (G + 1) + g, 1);
  }
}

// out (doubles, ACCUMULATED): per class c in {0, 1}: [c*4 + 0] sum of per-scene mean coverage, [c*4 + 1] sum of per-scene
// weighted coverage, [c*4 + 2] scenes that had a gt instance of the class, [c*4 + 3] gt instances; then for threshold k in
// 0..4 and class c: tp at [8 + (k*2 + c)*2], fp at [8 + (k*2 + c)*2 + 1].
__global__ void __launch_bounds__(256) k_cluster_scores(const int* __restrict__ table, const int* __restrict__ est_size,
                                                        const int* __restrict__ est_mos, const int* __restrict__ gt_size,
                                                        const int* __restrict__ gt_mos, int E, int G,
                                                        double* __restrict__ out) {
  __shared__ double s_cov[2], s_wcov[2];
  __shared__ int s_ngt[2], s_npts[2], s_tp[10], s_fp[10];
  if (threadIdx.x < 2) s_cov[threadIdx.x] = s_wcov[threadIdx.x] = 0.0, s_ngt[threadIdx.x] = s_npts[threadIdx.x] = 0;
  if (threadIdx.x < 10) s_tp[threadIdx.x] = s_fp[threadIdx.x] = 0;
  __syncthreads();
  // an instance's class = round(mean(mos_label)) with Python's round-half-to-even: 1 iff 2*sum > count
  // coverage: every gt instance looks for its best-overlapping predicted instance of the same class
  for (int g = 1 + threadIdx.x; g <= G; g += blockDim.x) {
    const int ng = gt_size[g];
    if (ng == 0) continue;
    const int cls = 2 * gt_mos[g] > ng;
    float ovmax = 0.f;
    for (int e = 1; e <= E; ++e) {
      const int ne = est_size[e];
      if (ne == 0 || (2 * est_mos[e] > ne) != cls) continue;
      const int inter = table[(size_t)e * (G + 1) + g];
      const float iou = (float)inter / (float)(ne + ng - inter);  // int64 / int64 -> float32 true division, as torch
      if (iou > ovmax) ovmax = iou;
    }
    atomicAdd(&s_cov[cls], (double)ovmax);
    atomicAdd(&s_wcov[cls], (double)ovmax * ng);
    atomicAdd(&s_ngt[cls], 1);
    atomicAdd(&s_npts[cls], ng);
  }
  // precision / recall: every predicted instance looks for its best gt instance of the same class
  const double thr[5] = {0.5, 0.6, 0.7, 0.8, 0.9};
  for (int e = 1 + threadIdx.x; e <= E; e += blockDim.x) {
    const int ne = est_size[e];
    if (ne == 0) continue;
    const int cls = 2 * est_mos[e] > ne;
    float ovmax = -1.f;
    for (int g = 1; g <= G; ++g) {
      const int ng = gt_size[g];
      if (ng == 0 || (2 * gt_mos[g] > ng) != cls) continue;
      const int inter = table[(size_t)e * (G + 1) + g];
      const float iou = (float)inter / (float)(ne + ng - inter);
      if (iou > ovmax) ovmax = iou;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if ((double)ovmax > thr[k]) atomicAdd(&s_tp[k * 2 + cls], 1);
      else atomicAdd(&s_fp[k * 2 + cls], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    const int c = threadIdx.x;
    if (s_ngt[c] > 0) {
      out[c * 4 + 0] += s_cov[c] / s_ngt[c];
      out[c * 4 + 1] += s_wcov[c] / s_npts[c];
      out[c * 4 + 2] += 1.0;
    }
    out[c * 4 + 3] += s_ngt[c];
  }
  if (threadIdx.x < 10) {
    out[8 + threadIdx.x * 2] += s_tp[threadIdx.x];
    out[8 + threadIdx.x * 2 + 1] += s_fp[threadIdx.x];
  }
}

size_t al256e(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t pcab_cluster_eval_workspace(int max_est, int max_gt) {
  return al256e(((size_t)(max_est + 1) * (max_gt + 1) + 2 * (size_t)(max_est + 1) + 2 * (size_t)(max_gt + 1)) * 4) + 256;
}

// One scene.  inst_est / inst_gt: 0 = background, instances 1..max_est / 1..max_gt; mos_label: 0 static, 1 dynamic.
// counters: 28 doubles, ACCUMULATED over scenes (layout above k_cluster_scores): zero them before the first scene.
extern "C" int pcab_cluster_eval(const long long* inst_est, const long long* inst_gt, const long long* mos_label, int n_points,
                                 int max_est, int max_gt, double* counters, void* workspace, size_t workspace_bytes,
                                 cudaStream_t stream) {
  PCAB_REQUIRE(max_est >= 0 && max_gt >= 0, "negative instance count");
  PCAB_REQUIRE(workspace_bytes >= pcab_cluster_eval_workspace(max_est, max_gt), "workspace too small");
  const size_t ints = (size_t)(max_est + 1) * (max_gt + 1) + 2 * (size_t)(max_est + 1) + 2 * (size_t)(max_gt + 1);
  int* table = (int*)workspace;
  int* est_size = table + (size_t)(max_est + 1) * (max_gt + 1);
  int* est_mos = est_size + (max_est + 1);
  int* gt_size = est_mos + (max_est + 1);
  int* gt_mos = gt_size + (max_gt + 1);
  PCAB_CUDA(cudaMemsetAsync(workspace, 0, ints * 4, stream));
  if (n_points > 0)
    k_contingency<<<grid_for(n_points, 256), 256, 0, stream>>>(inst_est, inst_gt, mos_label, n_points, max_est, max_gt, table,
                                                               est_size, est_mos, gt_size, gt_mos);
  k_cluster_scores<<<1, 256, 0, stream>>>(table, est_size, est_mos, gt_size, gt_mos, max_est, max_gt, counters);
  PCAB_CHECK_LAUNCH("pcab_cluster_eval");
  return PCAB_OK;
}
