// Uniform-grid exact nearest-neighbour search (see nn_grid.cuh) and the batched point-to-point ICP built on it.
//
// ICP replaces the Open3D calls of the reference's optional refinement branches (models/egomotion.py:9-28,360-384 --
// model.ego_icp; models/alignnet.py:54-112 -- model.tpointnet_icp): registration_icp(source, target, max_dist, init,
// TransformationEstimationPointToPoint(), ICPConvergenceCriteria(max_iteration)).  Open3D is not vendored by the
// reference (shim 3); the published algorithm of RegistrationICP is restated:
//   result = correspondences(T src);  repeat max_iteration times { update = umeyama(no scale) over the correspondence
//   set; T = update T; result' = correspondences(T src); stop when |fitness - fitness'| < 1e-6 and |rmse - rmse'| < 1e-6 }
// with a correspondence = nearest target within max_dist.  All problems of a call (frames of a scene, (instance, frame)
// pairs of TubeNet) run in the same launches; a problem only matches targets of its own group.
#include <cub/cub.cuh>
#include "common.cuh"
#include "nn_grid.cuh"
#include "pcab200.h"
#include "svd3.cuh"

namespace nngrid {

namespace {

__device__ __forceinline__ unsigned enc(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void k_init(Header* H) {
  H->bbox[0] = H->bbox[1] = H->bbox[2] = 0xffffffffu;
  H->bbox[3] = H->bbox[4] = H->bbox[5] = 0u;
}

__global__ void k_bbox(const float* __restrict__ p, int m, Header* H) {
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  bool bad = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = p[3 * (size_t)i + a];
      bad |= !isfinite(v);
      lo[a] = fminf(lo[a], v), hi[a] = fmaxf(hi[a], v);
    }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (bad) lo[a] = -INFINITY, hi[a] = INFINITY;  // a non-finite coordinate invalidates the grid (brute force answers)
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&H->bbox[a], enc(lo[a]));
      atomicMax(&H->bbox[3 + a], enc(hi[a]));
    }
  }
}

__global__ void k_params(Header* H, int m, float cell_hint, int ncell_max) {
  float lo[3], ext[3], big = 0.f;
  bool ok = m > 0;
  for (int a = 0; a < 3; ++a) {
    lo[a] = dec(H->bbox[a]);
    const float hi = dec(H->bbox[3 + a]);
    ok = ok && isfinite(lo[a]) && isfinite(hi) && hi >= lo[a];
    ext[a] = hi - lo[a];
    ok = ok && isfinite(ext[a]);
    big = fmaxf(big, ext[a]);
  }
  H->valid = ok ? 1 : 0;
  H->dx = H->dy = H->dz = 1;
  H->ox = H->oy = H->oz = 0.f, H->h = H->inv_h = 1.f;
  if (!ok) return;
  float h = cell_hint;
  if (!(h > 0.f)) {
    float vol = 1.f;
    for (int a = 0; a < 3; ++a) vol *= fmaxf(ext[a], 1e-3f * big);
    h = big > 0.f ? cbrtf(vol / (float)m) : 1.f;
    if (!(h > 0.f) || !isfinite(h)) h = 1.f;
  }
  int d[3];
  for (int it = 0; it < 400; ++it) {
    bool fits = true;
    for (int a = 0; a < 3; ++a) {
      const float c = floorf(ext[a] / h) + 1.f;
      fits = fits && c <= (float)MAX_DIM;
      d[a] = (int)fminf(c, (float)MAX_DIM);
    }
    if (fits && (long long)d[0] * d[1] * d[2] <= ncell_max) break;
    h *= 1.2f;
  }
  if ((long long)d[0] * d[1] * d[2] > ncell_max) {  // (cannot happen after 400 enlargements; keep the table in bounds anyway)
    H->valid = 0;
    return;
  }
  H->ox = lo[0], H->oy = lo[1], H->oz = lo[2];
  H->h = h, H->inv_h = 1.f / h;
  H->dx = d[0], H->dy = d[1], H->dz = d[2];
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ int cell_of(const Header& H, float x, float y, float z) {
  const int cx = clampi((int)floorf((x - H.ox) * H.inv_h), 0, H.dx - 1);
  const int cy = clampi((int)floorf((y - H.oy) * H.inv_h), 0, H.dy - 1);
  const int cz = clampi((int)floorf((z - H.oz) * H.inv_h), 0, H.dz - 1);
  return (cz * H.dy + cy) * H.dx + cx;
}

__global__ void k_hist(const float* __restrict__ p, int m, const Header* __restrict__ Hp, int* __restrict__ count) {
  const Header H = *Hp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    atomicAdd(count + cell_of(H, p[3 * (size_t)i], p[3 * (size_t)i + 1], p[3 * (size_t)i + 2]), 1);
}

__global__ void k_scatter(const float* __restrict__ p, int m, const Header* __restrict__ Hp, const int* __restrict__ start,
                          int* __restrict__ count, float4* __restrict__ sorted) {
  const Header H = *Hp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const float x = p[3 * (size_t)i], y = p[3 * (size_t)i + 1], z = p[3 * (size_t)i + 2];
    const int c = cell_of(H, x, y, z);
    const int pos = start[c] + atomicSub(count + c, 1) - 1;  // order inside a cell is irrelevant: (d, index) is minimised
    sorted[pos] = make_float4(x, y, z, __int_as_float(i));
  }
}

// nearest target of (px,py,pz): packed (d bits << 32 | index), ~0 when nothing qualifies.  `settled` = the result is final.
template <bool GROUPED>
__device__ __forceinline__ unsigned long long search(const Header& H, const int* __restrict__ start,
                                                     const float4* __restrict__ pts, float px, float py, float pz,
                                                     float md2 /* INFINITY = unbounded */, const int* __restrict__ tgt_group,
                                                     int group, bool& settled) {
  const float ax = px - H.ox, ay = py - H.oy, az = pz - H.oz;
  const int cx = clampi((int)floorf(ax * H.inv_h), 0, H.dx - 1);
  const int cy = clampi((int)floorf(ay * H.inv_h), 0, H.dy - 1);
  const int cz = clampi((int)floorf(az * H.inv_h), 0, H.dz - 1);
  const float h = H.h, marg = 4e-3f * h;  // covers the float32 rounding of the cell assignment (<= 1024 cells per axis)
  unsigned long long best = ~0ull;
  float bd = INFINITY;
  settled = false;
  auto scan = [&](int a, int b) {
    for (int i = a; i < b; ++i) {
      const float4 t = __ldg(pts + i);
      const float ex = __fsub_rn(t.x, px), ey = __fsub_rn(t.y, py), ez = __fsub_rn(t.z, pz);
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
      if (d <= bd) {
        const int idx = __float_as_int(t.w);
        if (GROUPED && tgt_group[idx] != group) continue;
        const unsigned long long pk = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)idx;
        if (pk < best) best = pk, bd = d;
      }
    }
  };
  for (int r = 0; r <= MAX_RINGS; ++r) {
    const int z0 = max(cz - r, 0), z1 = min(cz + r, H.dz - 1), y0 = max(cy - r, 0), y1 = min(cy + r, H.dy - 1);
    const int xl = cx - r, xr = cx + r, xa = max(xl, 0), xb = min(xr, H.dx - 1);
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) {
        const int row = (z * H.dy + y) * H.dx;
        if (r == 0 || z == cz - r || z == cz + r || y == cy - r || y == cy + r) {
          scan(start[row + xa], start[row + xb + 1]);  // the cells of a row are contiguous in the sorted array
        } else {
          if (xl >= 0) scan(start[row + xl], start[row + xl + 1]);
          if (xr < H.dx) scan(start[row + xr], start[row + xr + 1]);
        }
      }
    // everything unvisited lies beyond a face of the (2r+1)^3 block of cells: distance of the query to the nearest such face
    float lb = INFINITY;
    if (cx - r > 0) lb = fminf(lb, ax - (float)(cx - r) * h - marg);
    if (cx + r < H.dx - 1) lb = fminf(lb, (float)(cx + r + 1) * h - ax - marg);
    if (cy - r > 0) lb = fminf(lb, ay - (float)(cy - r) * h - marg);
    if (cy + r < H.dy - 1) lb = fminf(lb, (float)(cy + r + 1) * h - ay - marg);
    if (cz - r > 0) lb = fminf(lb, az - (float)(cz - r) * h - marg);
    if (cz + r < H.dz - 1) lb = fminf(lb, (float)(cz + r + 1) * h - az - marg);
    if (lb == INFINITY) {  // the whole grid has been visited
      settled = true;
      break;
    }
    lb = fmaxf(lb, 0.f);
    const float lb2 = lb * lb * 0.99999f;
    if (bd < lb2 || lb2 >= md2) {
      settled = true;
      break;
    }
  }
  if (!(bd < md2)) best = ~0ull;
  return best;
}

__global__ void __launch_bounds__(128) k_query(const Header* __restrict__ Hp, const int* __restrict__ start,
                                               const float4* __restrict__ pts, const float* __restrict__ q,
                                               const float4* __restrict__ qsorted, int n, float max_dist,
                                               const float* __restrict__ tsfm, unsigned long long* __restrict__ best,
                                               int* __restrict__ fb_list, int* __restrict__ fb_count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Header H = *Hp;
  float px, py, pz;
  int out = i;
  if (qsorted) {
    const float4 v = qsorted[i];
    px = v.x, py = v.y, pz = v.z, out = __float_as_int(v.w);
  } else {
    px = q[3 * (size_t)i], py = q[3 * (size_t)i + 1], pz = q[3 * (size_t)i + 2];
  }
  if (tsfm) {
    const float x = px, y = py, z = pz;
    px = tsfm[0] * x + tsfm[1] * y + tsfm[2] * z + tsfm[3];
    py = tsfm[4] * x + tsfm[5] * y + tsfm[6] * z + tsfm[7];
    pz = tsfm[8] * x + tsfm[9] * y + tsfm[10] * z + tsfm[11];
  }
  const bool bounded = max_dist > 0.f;
  bool settled = false;
  unsigned long long b = ~0ull;
  if (H.valid) b = search<false>(H, start, pts, px, py, pz, bounded ? max_dist * max_dist : INFINITY, nullptr, 0, settled);
  best[out] = b;
  if (!bounded && !settled) fb_list[atomicAdd(fb_count, 1)] = out;
}

// ---------------------------------------------------------------------------------------------------------------------
// batched point-to-point ICP
// ---------------------------------------------------------------------------------------------------------------------
constexpr int NS = 18;  // per problem: n_corr, sum p[3], sum t[3], sum p_a t_b [9], sum d^2, n_src

template <bool GROUPED>
__global__ void __launch_bounds__(128) k_icp_query(const Header* __restrict__ Hp, const int* __restrict__ start,
                                                    const float4* __restrict__ pts, const float* __restrict__ tgt,
                                                    const float* __restrict__ src, const int* __restrict__ src_problem, int n,
                                                    const int* __restrict__ tgt_group, const int* __restrict__ problem_group,
                                                    int P, const double* __restrict__ T, const int* __restrict__ done,
                                                    float max_dist, double* __restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int pid = -1;
  if (i < n) {
    pid = src_problem ? src_problem[i] : 0;
    if (pid < 0 || pid >= P || done[pid]) pid = -1;
  }
  double v[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) v[k] = 0.0;
  if (pid >= 0) {
    const Header H = *Hp;
    const double* M = T + (size_t)pid * 12;
    const double x = src[3 * (size_t)i], y = src[3 * (size_t)i + 1], z = src[3 * (size_t)i + 2];
    const double qx = M[0] * x + M[1] * y + M[2] * z + M[3];
    const double qy = M[4] * x + M[5] * y + M[6] * z + M[7];
    const double qz = M[8] * x + M[9] * y + M[10] * z + M[11];
    v[17] = 1.0;
    bool settled;
    unsigned long long b = ~0ull;
    if (H.valid)
      b = search<GROUPED>(H, start, pts, (float)qx, (float)qy, (float)qz, max_dist * max_dist, tgt_group,
                          problem_group ? problem_group[pid] : 0, settled);
    if (b != ~0ull) {
      const int j = (int)(unsigned)(b & 0xffffffffu);
      const double tx = tgt[3 * (size_t)j], ty = tgt[3 * (size_t)j + 1], tz = tgt[3 * (size_t)j + 2];
      const double ex = tx - qx, ey = ty - qy, ez = tz - qz;
      v[0] = 1.0;
      v[1] = qx, v[2] = qy, v[3] = qz, v[4] = tx, v[5] = ty, v[6] = tz;
      v[7] = qx * tx, v[8] = qx * ty, v[9] = qx * tz;
      v[10] = qy * tx, v[11] = qy * ty, v[12] = qy * tz;
      v[13] = qz * tx, v[14] = qz * ty, v[15] = qz * tz;
      v[16] = ex * ex + ey * ey + ez * ez;
    }
  }
  // warps whose lanes all belong to one problem (the common case: rows are grouped by problem) reduce before the atomics
  const int lead = __shfl_sync(0xffffffffu, pid, 0);
  const bool uniform = __all_sync(0xffffffffu, pid == lead);
  if (uniform) {
    if (lead < 0) return;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const double s = warp_sum_d(v[k]);
      if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(sums + (size_t)lead * NS + k, s);
    }
  } else if (pid >= 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k)
      if (v[k] != 0.0) atomicAdd(sums + (size_t)pid * NS + k, v[k]);
  }
}

__global__ void k_icp_init(const float* __restrict__ init, int P, double* __restrict__ T, double* __restrict__ sums,
                           double* __restrict__ prev, int* __restrict__ done) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  for (int k = 0; k < 12; ++k) T[(size_t)p * 12 + k] = init ? (double)init[(size_t)p * 16 + k] : ((k % 5 == 0) ? 1.0 : 0.0);
  for (int k = 0; k < NS; ++k) sums[(size_t)p * NS + k] = 0.0;
  prev[2 * p] = prev[2 * p + 1] = 0.0;
  done[p] = 0;
}

__global__ void k_icp_update(int P, int iter, int last, double rel_fitness, double rel_rmse, double* __restrict__ T,
                             double* __restrict__ sums, double* __restrict__ prev, int* __restrict__ done,
                             float* __restrict__ pose_out, float* __restrict__ stats) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P || done[p]) return;
  double S[NS];
  for (int k = 0; k < NS; ++k) S[k] = sums[(size_t)p * NS + k], sums[(size_t)p * NS + k] = 0.0;
  const double n = S[0];
  const double fitness = S[17] > 0 ? n / S[17] : 0.0, rmse = n > 0 ? sqrt(S[16] / n) : 0.0;
  double* M = T + (size_t)p * 12;
  bool stop = last != 0;
  if (iter > 0 && fabs(prev[2 * p] - fitness) < rel_fitness && fabs(prev[2 * p + 1] - rmse) < rel_rmse) stop = true;
  prev[2 * p] = fitness, prev[2 * p + 1] = rmse;
  if (!stop && n > 0) {
    double pm[3] = {S[1] / n, S[2] / n, S[3] / n}, tm[3] = {S[4] / n, S[5] / n, S[6] / n};
    double C[3][3], R[3][3];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) C[a][b] = S[7 + 3 * a + b] / n - pm[a] * tm[b];
    kabsch_rotation(C, R);
    double U[12], N2[12];
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3; ++b) U[4 * a + b] = R[a][b];
      U[4 * a + 3] = tm[a] - (R[a][0] * pm[0] + R[a][1] * pm[1] + R[a][2] * pm[2]);
    }
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 4; ++b)
        N2[4 * a + b] = U[4 * a] * M[b] + U[4 * a + 1] * M[4 + b] + U[4 * a + 2] * M[8 + b] + (b == 3 ? U[4 * a + 3] : 0.0);
    for (int k = 0; k < 12; ++k) M[k] = N2[k];
  }
  if (stop) done[p] = 1;
  for (int k = 0; k < 12; ++k) pose_out[(size_t)p * 16 + k] = (float)M[k];
  pose_out[(size_t)p * 16 + 12] = pose_out[(size_t)p * 16 + 13] = pose_out[(size_t)p * 16 + 14] = 0.f;
  pose_out[(size_t)p * 16 + 15] = 1.f;
  if (stats) stats[3 * p] = (float)fitness, stats[3 * p + 1] = (float)rmse, stats[3 * p + 2] = (float)iter;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

Layout layout(int m) {
  Layout L;
  L.ncell_max = max_cells(m);
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int*)nullptr, (int*)nullptr, L.ncell_max + 1);
  size_t off = 0;
  L.header = off, off += align256(sizeof(Header));
  L.count = off, off += align256((size_t)(L.ncell_max + 1) * 4);
  L.start = off, off += align256((size_t)(L.ncell_max + 1) * 4);
  L.sorted = off, off += align256((size_t)(m > 0 ? m : 1) * 16);
  L.scan_tmp = off, off += align256(scan_bytes);
  L.total = off;
  return L;
}

int build(const float* targets, int m, float cell_hint, void* ws, cudaStream_t stream) {
  const Layout L = layout(m);
  char* base = (char*)ws;
  Header* H = (Header*)(base + L.header);
  int* count = (int*)(base + L.count);
  int* start = (int*)(base + L.start);
  float4* sorted = (float4*)(base + L.sorted);
  k_init<<<1, 1, 0, stream>>>(H);
  if (m > 0) k_bbox<<<grid_for(m, 256, 4), 256, 0, stream>>>(targets, m, H);
  k_params<<<1, 1, 0, stream>>>(H, m, cell_hint, L.ncell_max);
  PCAB_CUDA(cudaMemsetAsync(count, 0, (size_t)(L.ncell_max + 1) * 4, stream));
  if (m > 0) k_hist<<<grid_for(m, 256), 256, 0, stream>>>(targets, m, H, count);
  size_t scan_bytes = L.total - L.scan_tmp;
  PCAB_CUDA(cub::DeviceScan::ExclusiveSum(base + L.scan_tmp, scan_bytes, count, start, L.ncell_max + 1, stream));
  if (m > 0) k_scatter<<<grid_for(m, 256), 256, 0, stream>>>(targets, m, H, start, count, sorted);
  PCAB_CHECK_LAUNCH("nngrid::build");
  return PCAB_OK;
}

int query(const void* grid_ws, int m, const float* queries, const float4* qsorted, int n, float max_dist, const float* tsfm,
          unsigned long long* best, int* fallback_list, int* fallback_count, cudaStream_t stream) {
  if (n <= 0) return PCAB_OK;
  const Layout L = layout(m);
  const char* base = (const char*)grid_ws;
  k_query<<<cdiv(n, 128), 128, 0, stream>>>((const Header*)(base + L.header), (const int*)(base + L.start),
                                            (const float4*)(base + L.sorted), queries, qsorted, n, max_dist, tsfm, best,
                                            fallback_list, fallback_count);
  PCAB_CHECK_LAUNCH("nngrid::query");
  return PCAB_OK;
}

}  // namespace nngrid

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct IcpLayout {
  size_t grid, T, sums, prev, done, total;
};
IcpLayout icp_layout(int n_tgt, int P) {
  IcpLayout L;
  size_t off = 0;
  L.grid = off, off += nngrid::layout(n_tgt).total;
  L.T = off, off += nngrid::align256((size_t)P * 12 * 8);
  L.sums = off, off += nngrid::align256((size_t)P * nngrid::NS * 8);
  L.prev = off, off += nngrid::align256((size_t)P * 2 * 8);
  L.done = off, off += nngrid::align256((size_t)P * 4);
  L.total = off;
  return L;
}
}  // namespace

extern "C" size_t pcab_icp_workspace(int n_targets, int n_problems) { return icp_layout(n_targets, n_problems > 0 ? n_problems : 1).total + 256; }

extern "C" int pcab_icp_point_to_point(const float* src, const int* src_problem, int n_src, const float* tgt, const int* tgt_group,
                                       int n_tgt, const int* problem_group, int n_problems, const float* init_pose, float max_dist,
                                       int max_iter, float rel_fitness, float rel_rmse, float* pose_out, float* stats,
                                       void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PCAB_REQUIRE(n_problems > 0 && n_src >= 0 && n_tgt >= 0 && max_iter >= 0, "bad sizes");
  PCAB_REQUIRE(max_dist > 0.f, "ICP needs a positive correspondence distance");
  PCAB_REQUIRE((tgt_group == nullptr) == (problem_group == nullptr), "tgt_group and problem_group go together");
  PCAB_REQUIRE(workspace_bytes >= pcab_icp_workspace(n_tgt, n_problems), "workspace too small");
  const IcpLayout L = icp_layout(n_tgt, n_problems);
  char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  const int P = n_problems;
  double* T = (double*)(base + L.T);
  double* sums = (double*)(base + L.sums);
  double* prev = (double*)(base + L.prev);
  int* done = (int*)(base + L.done);
  int rc = nngrid::build(tgt, n_tgt, max_dist, base + L.grid, stream);
  if (rc != PCAB_OK) return rc;
  const nngrid::Layout G = nngrid::layout(n_tgt);
  const nngrid::Header* H = (const nngrid::Header*)(base + L.grid + G.header);
  const int* start = (const int*)(base + L.grid + G.start);
  const float4* pts = (const float4*)(base + L.grid + G.sorted);
  nngrid::k_icp_init<<<cdiv(P, 64), 64, 0, stream>>>(init_pose, P, T, sums, prev, done);
  auto q = [&]() {
    if (n_src <= 0) return;
    if (tgt_group)
      nngrid::k_icp_query<true><<<cdiv(n_src, 128), 128, 0, stream>>>(H, start, pts, tgt, src, src_problem, n_src, tgt_group,
                                                                      problem_group, P, T, done, max_dist, sums);
    else
      nngrid::k_icp_query<false><<<cdiv(n_src, 128), 128, 0, stream>>>(H, start, pts, tgt, src, src_problem, n_src, nullptr,
                                                                       nullptr, P, T, done, max_dist, sums);
  };
  q();
  for (int it = 0; it < max_iter; ++it) {
    nngrid::k_icp_update<<<cdiv(P, 64), 64, 0, stream>>>(P, it, 0, rel_fitness, rel_rmse, T, sums, prev, done, pose_out, stats);
    q();
  }
  nngrid::k_icp_update<<<cdiv(P, 64), 64, 0, stream>>>(P, max_iter, 1, rel_fitness, rel_rmse, T, sums, prev, done, pose_out, stats);
  PCAB_CHECK_LAUNCH("pcab_icp_point_to_point");
  return PCAB_OK;
}
