// FuseLoss forward (first half): the loss terms of libs/loss.py:52-320 on the device.
//
//   pcab_seg_loss     get_seg_loss (libs/loss.py:113-136): online class weights (get_ce_weights :93-111, sqrt inverse
//                     frequency clamped at 50), weighted cross entropy with ignore_index -1, Lovasz-softmax
//                     (libs/lovasz_softmax.py:56-94: errors sorted descending, Jaccard gradient, dot product, mean over
//                     the classes that are present) and the IoU counters of compute_iou (:17-48) -- for the FG/BG map
//                     (get_fb_loss :165-186: occupied pillars of fb_seg_est) and the motion logits (get_mos_loss :139-163:
//                     points that are foreground in the labels or in the estimate).
//   pcab_offset_loss  get_offset_loss (:189-245): GT reconstruction (ego compensation + instance motion,
//                     toolbox/register_utils.py:59-93), instance centres (scatter mean), L1 / direction losses, L2 error.
//   pcab_perm_loss    OutlierLoss (libs/outlier_loss.py:15-29) over the soft-assignment matrices.
// and the gradients of those terms with respect to the network outputs (the first step of the backward pass):
//   pcab_seg_loss_grad / pcab_offset_loss_grad.
//
// The selection (occupied pillars / candidate points) is never compacted: unselected items get the sort key -1 (errors are
// >= 0), so after the descending sort the selected ones are the first n_sel rows; n_sel stays on the device.
#include <cub/cub.cuh>
#include "common.cuh"
#include "pcab200.h"

namespace {

constexpr double EPS = 1e-20;  // toolbox/utils.py:13

// counters of one segmentation loss (doubles): n_sel, n[2], S[2] (sum of -log p_y per class), pred[2], inter[2], lov[2]
enum { C_NSEL = 0, C_N0 = 1, C_S0 = 3, C_PRED0 = 5, C_INT0 = 7, C_LOV0 = 9, C_TOTAL = 12 };

__device__ __forceinline__ size_t logit_addr(long long item, int c, int hw) {
  // [.., 2, hw] blocks (hw = Ny*Nx for the BEV map, 1 for [N,2] point logits)
  return (size_t)(item / hw) * (2 * (size_t)hw) + (size_t)(item % hw) + (size_t)c * hw;
}

// sum over the block, then ONE atomic per block and counter (a few hundred thousand warp-level double atomics on a handful of
// addresses would cost more than the pass itself).  All threads of the block must call it; blockDim.x <= 1024.
template <int NV>
__device__ __forceinline__ void block_atomic_add(double* dst, const double (&v)[NV]) {
  __shared__ double s_part[32][NV];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = warp_sum_d(v[k]);
    if (lane == 0) s_part[warp][k] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = lane < nwarp ? s_part[lane][k] : 0.0;
      s = warp_sum_d(s);
      if (lane == 0 && s != 0.0) atomicAdd(dst + k, s);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void softmax2(float z0, float z1, float& p0, float& p1, float& l0, float& l1) {
  const float m = fmaxf(z0, z1);
  const float e0 = expf(z0 - m), e1 = expf(z1 - m);
  const float s = e0 + e1;
  p0 = e0 / s, p1 = e1 / s;
  const float ls = logf(s);
  l0 = (z0 - m) - ls, l1 = (z1 - m) - ls;
}

// selected(i): sel_a[i] == 1 (float map, may be NULL) or sel_b[i] == 1 or sel_c[i] == 1 (int64 label arrays, may be NULL)
__device__ __forceinline__ bool is_selected(long long i, const float* sel_a, const long long* sel_b, const long long* sel_c) {
  bool s = false;
  if (sel_a) s |= sel_a[i] == 1.f;
  if (sel_b) s |= sel_b[i] == 1;
  if (sel_c) s |= sel_c[i] == 1;
  return s;
}

__global__ void k_seg_prepare(const float* __restrict__ logits, int hw, const long long* __restrict__ gt,
                              const float* __restrict__ sel_a, const long long* __restrict__ sel_b,
                              const long long* __restrict__ sel_c, long long n, float* __restrict__ err0,
                              float* __restrict__ err1, unsigned char* __restrict__ fg0, unsigned char* __restrict__ fg1,
                              double* __restrict__ C) {
  double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // n_sel, n0, n1, S0, S1, pred0, pred1, int0, int1
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    if (!is_selected(i, sel_a, sel_b, sel_c)) {
      err0[i] = err1[i] = -1.f;
      fg0[i] = fg1[i] = 0;
      continue;
    }
    const float z0 = logits[logit_addr(i, 0, hw)], z1 = logits[logit_addr(i, 1, hw)];
    float p0, p1, l0, l1;
    softmax2(z0, z1, p0, p1, l0, l1);
    const long long y = gt[i];
    const float f0 = y == 0 ? 1.f : 0.f, f1 = y == 1 ? 1.f : 0.f;
    err0[i] = fabsf(f0 - p0), err1[i] = fabsf(f1 - p1);
    fg0[i] = y == 0, fg1[i] = y == 1;
    acc[0] += 1.0;
    if (y == 0) acc[1] += 1.0, acc[3] += (double)(-l0);
    if (y == 1) acc[2] += 1.0, acc[4] += (double)(-l1);
    const int pred = z1 > z0 ? 1 : 0;  // argmax, first maximum on ties
    acc[5 + pred] += 1.0;
    if (y == pred && (y == 0 || y == 1)) acc[7 + pred] += 1.0;
  }
  block_atomic_add<9>(C, acc);
}

// Lovasz extension of one class over the sorted errors: sum_i e_i * (J_i - J_{i-1}),  J_i = 1 - (G - c_i) / (G + (i+1) - c_i)
__global__ void k_lovasz(const float* __restrict__ err_sorted, const int* __restrict__ cum, const double* __restrict__ C, int cls,
                         long long n, double* __restrict__ out) {
  const long long n_sel = (long long)C[C_NSEL];
  const double G = C[C_N0 + cls];
  double acc = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_sel && i < n; i += stride) {
    const double c = cum[i];
    const double J = 1.0 - (G - c) / (G + (double)(i + 1) - c);
    double Jp = 0.0;
    if (i > 0) {
      const double cp = cum[i - 1];
      Jp = 1.0 - (G - cp) / (G + (double)i - cp);
    }
    acc += (double)err_sorted[i] * (J - Jp);
  }
  const double a1[1] = {acc};
  block_atomic_add<1>(out, a1);
}

// d(lovasz_c)/d(p_c,i) = g_rank(i) * sign(p_c,i - fg_i): scattered back through the sort permutation
__global__ void k_lovasz_grad(const float* __restrict__ err_sorted, const int* __restrict__ cum, const int* __restrict__ perm,
                              const double* __restrict__ C, int cls, long long n, float* __restrict__ g_out) {
  const long long n_sel = (long long)C[C_NSEL];
  const double G = C[C_N0 + cls];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_sel && i < n; i += stride) {
    const double c = cum[i];
    const double J = 1.0 - (G - c) / (G + (double)(i + 1) - c);
    double Jp = 0.0;
    if (i > 0) {
      const double cp = cum[i - 1];
      Jp = 1.0 - (G - cp) / (G + (double)i - cp);
    }
    g_out[perm[i]] = (float)(J - Jp);
  }
}

// out[0] = weighted CE, out[1] = Lovasz, out[2..9] = intersection, union, pred_positives, gt_positives for classes 0, 1
// (raw counts; compute_iou divides them by 1e3), out[10] = n_sel, out[11..12] = class weights
__global__ void k_seg_final(const double* __restrict__ C, float* __restrict__ out) {
  if (threadIdx.x || blockIdx.x) return;
  const double c0 = C[C_N0] + EPS, c1 = C[C_N0 + 1] + EPS;
  // get_ce_weights works in float32 (torch.tensor of python floats): counts.sum() / counts, sqrt, clamp(0, 50)
  const float f0 = (float)c0, f1 = (float)c1;
  const float tot = f0 + f1;
  const float w0 = fminf(fmaxf(sqrtf(tot / f0), 0.f), 50.f), w1 = fminf(fmaxf(sqrtf(tot / f1), 0.f), 50.f);
  const double den = (double)w0 * C[C_N0] + (double)w1 * C[C_N0 + 1];
  out[0] = den > 0 ? (float)(((double)w0 * C[C_S0] + (double)w1 * C[C_S0 + 1]) / den) : 0.f;  // nothing selected: 0 (libs/loss.py:152-162)
  double lov = 0;
  int present = 0;
  for (int c = 0; c < 2; ++c)
    if (C[C_N0 + c] > 0) lov += C[C_LOV0 + c], ++present;
  out[1] = present ? (float)(lov / present) : 0.f;
  for (int c = 0; c < 2; ++c) {
    const double inter = C[C_INT0 + c], pred = C[C_PRED0 + c], gtp = C[C_N0 + c];
    out[2 + c] = (float)inter, out[4 + c] = (float)(pred + gtp - inter), out[6 + c] = (float)pred, out[8 + c] = (float)gtp;
  }
  out[10] = (float)C[C_NSEL];
  out[11] = w0, out[12] = w1;
}

// gradient of  a * CE + b * Lovasz  with respect to the two logits of every item (0 where unselected)
__global__ void k_seg_grad(const float* __restrict__ logits, int hw, const long long* __restrict__ gt,
                           const float* __restrict__ sel_a, const long long* __restrict__ sel_b,
                           const long long* __restrict__ sel_c, long long n, const double* __restrict__ C,
                           const float* __restrict__ out, const float* __restrict__ g0, const float* __restrict__ g1, float w_ce,
                           float w_lov, float* __restrict__ grad) {
  const float w0 = out[11], w1 = out[12];
  const double den = (double)w0 * C[C_N0] + (double)w1 * C[C_N0 + 1];
  const int present = (C[C_N0] > 0) + (C[C_N0 + 1] > 0);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    float d0 = 0.f, d1 = 0.f;
    if (is_selected(i, sel_a, sel_b, sel_c)) {
      const float z0 = logits[logit_addr(i, 0, hw)], z1 = logits[logit_addr(i, 1, hw)];
      float p0, p1, l0, l1;
      softmax2(z0, z1, p0, p1, l0, l1);
      const long long y = gt[i];
      if (y == 0 || y == 1) {  // ignore_index and out-of-range labels carry no CE gradient
        const float wy = (float)((double)(y == 0 ? w0 : w1) / den);
        d0 += w_ce * wy * (p0 - (y == 0 ? 1.f : 0.f));
        d1 += w_ce * wy * (p1 - (y == 1 ? 1.f : 0.f));
      }
      if (present) {
        // dL/dp_c = g_c * sign(p_c - fg_c) / present; through the softmax: dz_k = p_k * (dL/dp_k - sum_j p_j dL/dp_j)
        const float s0 = (p0 - (y == 0 ? 1.f : 0.f)) > 0.f ? 1.f : ((p0 - (y == 0 ? 1.f : 0.f)) < 0.f ? -1.f : 0.f);
        const float s1 = (p1 - (y == 1 ? 1.f : 0.f)) > 0.f ? 1.f : ((p1 - (y == 1 ? 1.f : 0.f)) < 0.f ? -1.f : 0.f);
        const float q0 = C[C_N0] > 0 ? g0[i] * s0 / present : 0.f;
        const float q1 = C[C_N0 + 1] > 0 ? g1[i] * s1 / present : 0.f;
        const float dot = p0 * q0 + p1 * q1;
        d0 += w_lov * p0 * (q0 - dot);
        d1 += w_lov * p1 * (q1 - dot);
      }
    }
    grad[logit_addr(i, 0, hw)] = d0;
    grad[logit_addr(i, 1, hw)] = d1;
  }
}

size_t al(size_t v) { return (v + 255) & ~(size_t)255; }

struct SegWs {
  size_t C, err0, err1, errs, fg0, fg1, fgs, cum, idx, perm, g0, g1, tmp, tmp_bytes, total;
};
SegWs seg_ws(long long n) {
  SegWs W;
  size_t sort_b = 0, sort_i = 0, scan_b = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_b, (float*)nullptr, (float*)nullptr, (unsigned char*)nullptr,
                                            (unsigned char*)nullptr, (int)n);
  cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_i, (float*)nullptr, (float*)nullptr, (int*)nullptr, (int*)nullptr, (int)n);
  cub::DeviceScan::InclusiveSum(nullptr, scan_b, (unsigned char*)nullptr, (int*)nullptr, (int)n);
  size_t off = 0;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  W.C = off, off += al(C_TOTAL * 8);
  W.err0 = off, off += al(nn * 4);
  W.err1 = off, off += al(nn * 4);
  W.errs = off, off += al(nn * 4);
  W.fg0 = off, off += al(nn);
  W.fg1 = off, off += al(nn);
  W.fgs = off, off += al(nn);
  W.cum = off, off += al(nn * 4);
  W.idx = off, off += al(nn * 4);
  W.perm = off, off += al(nn * 4);
  W.g0 = off, off += al(nn * 4);
  W.g1 = off, off += al(nn * 4);
  W.tmp_bytes = sort_b > sort_i ? sort_b : sort_i;
  if (scan_b > W.tmp_bytes) W.tmp_bytes = scan_b;
  W.tmp = off, off += al(W.tmp_bytes);
  W.total = off;
  return W;
}

__global__ void k_iota(int* __restrict__ p, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) p[i] = (int)i;
}

// ---------------------------------------------------------------------------------------------------------------------
// offset loss
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void apply34(const float* M, float x, float y, float z, float& ox, float& oy, float& oz) {
  ox = M[0] * x + M[1] * y + M[2] * z + M[3];
  oy = M[4] * x + M[5] * y + M[6] * z + M[7];
  oz = M[8] * x + M[9] * y + M[10] * z + M[11];
}

constexpr int kRep = 64;  // replicas of the per-instance sums (contention of the double atomics)

// rec = bbox_tsfm[inst, t] (ego_gt[b, t] p): sums per (scene, instance)
__global__ void k_offset_centres(const float* __restrict__ pts, const int* __restrict__ pbatch, const int* __restrict__ ptime,
                                 const long long* __restrict__ inst, const float* __restrict__ ego_gt,
                                 const float* __restrict__ motion /* [sum K_b, T, 4, 4] */, const int* __restrict__ koff, int T,
                                 long long n, double* __restrict__ sums /* [sum K_b][4] */) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const int b = pbatch[i], t = ptime[i];
    const long long k = koff[b] + inst[i];
    float x, y, z, rx, ry, rz;
    apply34(ego_gt + ((size_t)b * T + t) * 16, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], x, y, z);
    apply34(motion + ((size_t)k * T + t) * 16, x, y, z, rx, ry, rz);
    // most points of a warp share an instance (background = one hot address): lanes with equal k combine first
    const unsigned peers = __match_any_sync(__activemask(), k);
    const int leader = __ffs(peers) - 1;
    double ax = 0, ay = 0, az = 0, an = 0;
    for (unsigned mm = peers; mm; mm &= mm - 1) {
      const int l = __ffs(mm) - 1;
      ax += (double)__shfl_sync(peers, rx, l), ay += (double)__shfl_sync(peers, ry, l), az += (double)__shfl_sync(peers, rz, l);
      an += 1.0;
    }
    if ((int)(threadIdx.x & 31) == leader) {
      // ... and the warps spread over kRep replicas of the table (the background instance alone collects ~10^5 warp sums)
      double* s = sums + 4 * ((size_t)k * kRep + ((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) & (kRep - 1)));
      atomicAdd(s, ax), atomicAdd(s + 1, ay), atomicAdd(s + 2, az), atomicAdd(s + 3, an);
    }
  }
}

__global__ void k_offset_fold(const double* __restrict__ rep, int k_total, double* __restrict__ sums) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 4 * k_total) return;
  const int k = t >> 2, c = t & 3;
  double a = 0;
  for (int r = 0; r < kRep; ++r) a += rep[4 * ((size_t)k * kRep + r) + c];  // fixed order: the centres do not depend on timing
  sums[t] = a;
}

// acc: sum |dx|, sum |dy|, sum |d|_2, sum (1 - cos), count
__global__ void k_offset_terms(const int* __restrict__ pbatch, const long long* __restrict__ inst, const long long* __restrict__ fb,
                               const int* __restrict__ koff, const double* __restrict__ sums, const float* __restrict__ tp,
                               const float* __restrict__ off_est, long long n, float* __restrict__ gt_offset /* [n,2] or NULL */,
                               double* __restrict__ acc) {
  double a[5] = {0, 0, 0, 0, 0};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const double* s = sums + 4 * (koff[pbatch[i]] + inst[i]);
    const float cx = (float)(s[0] / s[3]), cy = (float)(s[1] / s[3]);
    const float gx = cx - tp[3 * i], gy = cy - tp[3 * i + 1];
    if (gt_offset) gt_offset[2 * i] = gx, gt_offset[2 * i + 1] = gy;
    if (fb[i] != 1) continue;
    const float ex = off_est[2 * i], ey = off_est[2 * i + 1];
    const float dx = gx - ex, dy = gy - ey;
    a[0] += fabsf(dx), a[1] += fabsf(dy), a[2] += sqrtf(dx * dx + dy * dy);
    const float gn = sqrtf(gx * gx + gy * gy) + 1e-20f, en = sqrtf(ex * ex + ey * ey) + 1e-20f;
    a[3] += 1.f - ((gx / gn) * (ex / en) + (gy / gn) * (ey / en));
    a[4] += 1.0;
  }
  block_atomic_add<5>(acc, a);
}

__global__ void k_offset_final(const double* __restrict__ acc, float* __restrict__ out) {
  const double n = acc[4];
  if (n > 0) {
    out[0] = (float)(acc[0] / n + acc[1] / n);  // offset_norm_loss: |.|.mean(dim=0).sum()
    out[1] = (float)(acc[3] / n);               // offset_dir_loss
    out[2] = (float)(acc[2] / n);               // offset_l2_error
  } else {
    out[0] = out[1] = out[2] = 0.f;
  }
  out[3] = (float)n;
}

// d(w_norm * norm_loss + w_dir * dir_loss) / d offset_est
__global__ void k_offset_grad(const long long* __restrict__ fb, const float* __restrict__ gt_offset, const float* __restrict__ off_est,
                              long long n, const float* __restrict__ out, float w_norm, float w_dir, float* __restrict__ grad) {
  const float cnt = out[3];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    float g0 = 0.f, g1 = 0.f;
    if (fb[i] == 1 && cnt > 0.f) {
      const float gx = gt_offset[2 * i], gy = gt_offset[2 * i + 1], ex = off_est[2 * i], ey = off_est[2 * i + 1];
      const float dx = gx - ex, dy = gy - ey;
      g0 = w_norm * (dx > 0.f ? -1.f : (dx < 0.f ? 1.f : 0.f)) / cnt;
      g1 = w_norm * (dy > 0.f ? -1.f : (dy < 0.f ? 1.f : 0.f)) / cnt;
      // 1 - <g^, e / (|e| + eps)>: d/de = -(g^ / s - e <g^, e> / (|e| s^2)),  s = |e| + eps
      const float gn = sqrtf(gx * gx + gy * gy) + 1e-20f;
      const float ux = gx / gn, uy = gy / gn;
      const float en = sqrtf(ex * ex + ey * ey), s = en + 1e-20f;
      if (en > 0.f) {
        const float dot = ux * ex + uy * ey;
        g0 += -w_dir * (ux / s - ex * dot / (en * s * s)) / cnt;
        g1 += -w_dir * (uy / s - ey * dot / (en * s * s)) / cnt;
      }
    }
    grad[2 * i] = g0, grad[2 * i + 1] = g1;
  }
}

__global__ void k_sum_all(const float* __restrict__ x, long long n, double* __restrict__ acc) {
  double a = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) a += (double)x[i];
  a = warp_sum_d(a);
  if ((threadIdx.x & 31) == 0) atomicAdd(acc, a);
}
__global__ void k_perm_final(const double* __restrict__ acc, double rows_total, float* __restrict__ out) {
  // mean(1 - column sums) + mean(1 - row sums) over all matrices: both equal 1 - sum(all) / (P * m)
  out[0] = (float)(2.0 * (1.0 - acc[0] / rows_total));
}

}  // namespace

extern "C" size_t pcab_seg_loss_workspace(long long n_items) { return seg_ws(n_items).total + 256; }

// logits: [.., 2, hw] (hw = Ny*Nx for fb_seg_est [B,T,2,Ny,Nx], 1 for mos_est [N,2]); gt: int64 per item; an item takes part
// when sel_float[i] == 1 or sel_a[i] == 1 or sel_b[i] == 1 (NULL arrays are skipped).  out13: see k_seg_final.
// The workspace keeps what pcab_seg_loss_grad needs (call it with the same workspace, before anything else reuses it).
extern "C" int pcab_seg_loss(const float* logits, int hw, const long long* gt, const float* sel_float, const long long* sel_a,
                             const long long* sel_b, long long n_items, float* out13, void* workspace, size_t workspace_bytes,
                             cudaStream_t stream) {
  PCAB_REQUIRE(n_items > 0 && n_items < (1LL << 31) && hw > 0, "bad sizes");
  PCAB_REQUIRE(workspace_bytes >= pcab_seg_loss_workspace(n_items), "workspace too small");
  const SegWs W = seg_ws(n_items);
  char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  double* C = (double*)(base + W.C);
  float *err0 = (float*)(base + W.err0), *err1 = (float*)(base + W.err1), *errs = (float*)(base + W.errs);
  unsigned char *fg0 = (unsigned char*)(base + W.fg0), *fg1 = (unsigned char*)(base + W.fg1), *fgs = (unsigned char*)(base + W.fgs);
  int* cum = (int*)(base + W.cum);
  const int n = (int)n_items;
  PCAB_CUDA(cudaMemsetAsync(C, 0, C_TOTAL * 8, stream));
  k_seg_prepare<<<grid_for(n, 256), 256, 0, stream>>>(logits, hw, gt, sel_float, sel_a, sel_b, n_items, err0, err1, fg0, fg1, C);
  for (int c = 0; c < 2; ++c) {
    size_t tb = W.tmp_bytes;
    PCAB_CUDA(cub::DeviceRadixSort::SortPairsDescending(base + W.tmp, tb, c ? err1 : err0, errs, c ? fg1 : fg0, fgs, n, 0, 32, stream));
    tb = W.tmp_bytes;
    PCAB_CUDA(cub::DeviceScan::InclusiveSum(base + W.tmp, tb, fgs, cum, n, stream));
    k_lovasz<<<grid_for(n, 256), 256, 0, stream>>>(errs, cum, C, c, n_items, C + C_LOV0 + c);
  }
  k_seg_final<<<1, 32, 0, stream>>>(C, out13);
  PCAB_CHECK_LAUNCH("pcab_seg_loss");
  return PCAB_OK;
}

// grad [same layout as logits] = d(w_ce * CE + w_lovasz * Lovasz) / d logits (0 at unselected items); same arguments and
// workspace as the pcab_seg_loss call it follows
extern "C" int pcab_seg_loss_grad(const float* logits, int hw, const long long* gt, const float* sel_float, const long long* sel_a,
                                  const long long* sel_b, long long n_items, const float* out13, float w_ce, float w_lovasz,
                                  float* grad, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PCAB_REQUIRE(n_items > 0 && n_items < (1LL << 31) && hw > 0, "bad sizes");
  PCAB_REQUIRE(workspace_bytes >= pcab_seg_loss_workspace(n_items), "workspace too small");
  const SegWs W = seg_ws(n_items);
  char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  double* C = (double*)(base + W.C);
  float *err0 = (float*)(base + W.err0), *err1 = (float*)(base + W.err1), *errs = (float*)(base + W.errs);
  unsigned char *fg0 = (unsigned char*)(base + W.fg0), *fg1 = (unsigned char*)(base + W.fg1), *fgs = (unsigned char*)(base + W.fgs);
  int *cum = (int*)(base + W.cum), *idx = (int*)(base + W.idx), *perm = (int*)(base + W.perm);
  float *g0 = (float*)(base + W.g0), *g1 = (float*)(base + W.g1);
  const int n = (int)n_items;
  PCAB_CUDA(cudaMemsetAsync(g0, 0, (size_t)n * 4, stream));
  PCAB_CUDA(cudaMemsetAsync(g1, 0, (size_t)n * 4, stream));
  k_iota<<<grid_for(n, 256), 256, 0, stream>>>(idx, n_items);
  for (int c = 0; c < 2; ++c) {
    // (the flags are sorted along once more to rebuild the cumulative counts: err / fg of the forward call are still in place)
    size_t tb = W.tmp_bytes;
    PCAB_CUDA(cub::DeviceRadixSort::SortPairsDescending(base + W.tmp, tb, c ? err1 : err0, errs, c ? fg1 : fg0, fgs, n, 0, 32, stream));
    tb = W.tmp_bytes;
    PCAB_CUDA(cub::DeviceScan::InclusiveSum(base + W.tmp, tb, fgs, cum, n, stream));
    tb = W.tmp_bytes;
    PCAB_CUDA(cub::DeviceRadixSort::SortPairsDescending(base + W.tmp, tb, c ? err1 : err0, errs, idx, perm, n, 0, 32, stream));
    k_lovasz_grad<<<grid_for(n, 256), 256, 0, stream>>>(errs, cum, perm, C, c, n_items, c ? g1 : g0);
  }
  k_seg_grad<<<grid_for(n, 256), 256, 0, stream>>>(logits, hw, gt, sel_float, sel_a, sel_b, n_items, C, out13, g0, g1, w_ce, w_lovasz,
                                                   grad);
  PCAB_CHECK_LAUNCH("pcab_seg_loss_grad");
  return PCAB_OK;
}

extern "C" size_t pcab_offset_loss_workspace(int n_instances_total) {
  const size_t k = (size_t)(n_instances_total > 0 ? n_instances_total : 1);
  return al(k * 32) + al(k * kRep * 32) + al(64) + 256;
}

// out4 = {offset_norm_loss, offset_dir_loss, offset_l2_error, n_foreground}; gt_offset [N,2] (every point; the reference keeps
// the rows of the foreground points as predictions['offset_gt']) may be NULL
extern "C" int pcab_offset_loss(const float* points, const int* point_batch, const int* point_time, const long long* inst_labels,
                                const long long* fb_labels, const float* ego_motion_gt, const float* inst_motion_gt,
                                const int* inst_offset, int n_instances_total, int T, const float* transformed_points,
                                const float* offset_est, long long n_points, float* gt_offset, float* out4, void* workspace,
                                size_t workspace_bytes, cudaStream_t stream) {
  PCAB_REQUIRE(n_points > 0 && T > 0 && n_instances_total > 0, "bad sizes");
  PCAB_REQUIRE(workspace_bytes >= pcab_offset_loss_workspace(n_instances_total), "workspace too small");
  char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  double* sums = (double*)base;
  double* rep = (double*)(base + al((size_t)n_instances_total * 32));
  double* acc = (double*)(base + al((size_t)n_instances_total * 32) + al((size_t)n_instances_total * kRep * 32));
  PCAB_CUDA(cudaMemsetAsync(rep, 0, (size_t)n_instances_total * kRep * 32, stream));
  PCAB_CUDA(cudaMemsetAsync(acc, 0, 64, stream));
  k_offset_centres<<<grid_for(n_points, 256), 256, 0, stream>>>(points, point_batch, point_time, inst_labels, ego_motion_gt,
                                                               inst_motion_gt, inst_offset, T, n_points, rep);
  k_offset_fold<<<cdiv(4LL * n_instances_total, 128), 128, 0, stream>>>(rep, n_instances_total, sums);
  k_offset_terms<<<grid_for(n_points, 256), 256, 0, stream>>>(point_batch, inst_labels, fb_labels, inst_offset, sums, transformed_points,
                                                             offset_est, n_points, gt_offset, acc);
  k_offset_final<<<1, 1, 0, stream>>>(acc, out4);
  PCAB_CHECK_LAUNCH("pcab_offset_loss");
  return PCAB_OK;
}

extern "C" int pcab_offset_loss_grad(const long long* fb_labels, const float* gt_offset, const float* offset_est, long long n_points,
                                     const float* out4, float w_norm, float w_dir, float* grad, cudaStream_t stream) {
  PCAB_REQUIRE(n_points > 0, "bad sizes");
  k_offset_grad<<<grid_for(n_points, 256), 256, 0, stream>>>(fb_labels, gt_offset, offset_est, n_points, out4, w_norm, w_dir, grad);
  PCAB_CHECK_LAUNCH("pcab_offset_loss_grad");
  return PCAB_OK;
}

// OutlierLoss over n_mats contiguous [m, m] matrices; scratch1 = one double
extern "C" int pcab_perm_loss(const float* perm, int n_mats, int m, double* scratch1, float* out1, cudaStream_t stream) {
  PCAB_REQUIRE(n_mats > 0 && m > 0, "bad sizes");
  PCAB_CUDA(cudaMemsetAsync(scratch1, 0, 8, stream));
  const long long n = (long long)n_mats * m * m;
  k_sum_all<<<grid_for(n, 256), 256, 0, stream>>>(perm, n, scratch1);
  k_perm_final<<<1, 1, 0, stream>>>(scratch1, (double)n_mats * m, out1);
  PCAB_CHECK_LAUNCH("pcab_perm_loss");
  return PCAB_OK;
}
