// Hash-voxelise (first-touch pillar numbering) + pillar index build + pillar statistics.
//
// Replaces libs/voxel_generator.py:4-61 (sequential numba loop), the pillar-offset part of
// libs/dataloader.py:33-38 and models/motionnet.py:159-160 (torch_scatter mean / max).
//
// Exact parallel restatement of the sequential first-touch rule:
//   first[cell] = min over points of the stream index  (atomicMin)
//   flag[i]     = first[cell(i)] == i
//   id[cell]    = exclusive_scan(flag)[first[cell]]
// All coordinate arithmetic is IEEE float32 (true division then floor) so that pillar ids are
// bit-exact with the reference.  Points of scene b precede those of scene b+1 in the stream, so the
// global rank already carries the running pillar offset the reference's collate_fn adds.
#include <cub/cub.cuh>
#include "common.cuh"
#include "pcab200.h"

namespace {

struct VoxGeom {
  float lo[3];
  float vs[3];
  int grid[3];  // nx, ny, nz
  int nt;
};

__global__ void k_cell(const float* __restrict__ pts4, const int* __restrict__ pbatch, int n, VoxGeom g,
                       int* __restrict__ cell_of, int* __restrict__ first) {
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float4 p = ld_stream_f4(reinterpret_cast<const float4*>(pts4) + i);
    float c0 = floorf(__fdiv_rn(__fsub_rn(p.x, g.lo[0]), g.vs[0]));
    float c1 = floorf(__fdiv_rn(__fsub_rn(p.y, g.lo[1]), g.vs[1]));
    float c2 = floorf(__fdiv_rn(__fsub_rn(p.z, g.lo[2]), g.vs[2]));
    int t = (int)p.w;
    bool ok = c0 >= 0.f && c0 < (float)g.grid[0] && c1 >= 0.f && c1 < (float)g.grid[1] && c2 >= 0.f &&
              c2 < (float)g.grid[2] && t >= 0 && t < g.nt;
    int cell = -1;
    if (ok) {
      int b = pbatch ? pbatch[i] : 0;
      cell = ((((b * g.grid[2] + (int)c2) * g.grid[1] + (int)c1) * g.grid[0] + (int)c0) * g.nt) + t;
      atomicMin(first + cell, i);
    }
    cell_of[i] = cell;
  }
}

__global__ void k_flag(const int* __restrict__ cell_of, const int* __restrict__ first, int n, int* __restrict__ flag) {
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int c = cell_of[i];
    flag[i] = (c >= 0 && first[c] == i) ? 1 : 0;
  }
}

__global__ void k_assign(const int* __restrict__ cell_of, const int* __restrict__ flag, const int* __restrict__ rank,
                         int n, VoxGeom g, int* __restrict__ first, int* __restrict__ coords,
                         int* __restrict__ pillar_batch, int* __restrict__ num_voxels, int* __restrict__ total) {
  const int stride = gridDim.x * blockDim.x;
  for (int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {  // warp-uniform trip count
    const int i = i0 + (threadIdx.x & 31);
    int b = -1;
    if (i < n && flag[i]) {
      int id = rank[i];
      int c = cell_of[i];
      int t = c % g.nt;
      int r = c / g.nt;
      int x = r % g.grid[0];
      r /= g.grid[0];
      int y = r % g.grid[1];
      r /= g.grid[1];
      int z = r % g.grid[2];
      b = r / g.grid[2];
      reinterpret_cast<int4*>(coords)[id] = make_int4(z, y, x, t);
      pillar_batch[id] = b;
      first[c] = -id - 2;  // reuse the table as cell -> pillar id (encoded negative)
    }
    // pillars per scene: the new pillars of a warp that belong to the same scene add up before the atomic (with one scene
    // per call every new pillar would otherwise hit the same counter)
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    if (b >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(num_voxels + b, __popc(peers));
    if (i == n - 1) *total = rank[i] + flag[i];
  }
}

__global__ void k_map(const int* __restrict__ cell_of, const int* __restrict__ first, int n, int* __restrict__ p2v) {
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int c = cell_of[i];
    p2v[i] = c >= 0 ? (-first[c] - 2) : -1;
  }
}

__global__ void k_iota(int* a, int n) {
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] = i;
}

// pstart[m] = first sorted position whose pillar id is >= m  (keys are sorted ascending)
__global__ void k_segment_starts(const int* __restrict__ keys_sorted, int n, int m, int* __restrict__ pstart) {
  int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= n; j += stride) {
    int cur = j < n ? keys_sorted[j] : m;
    int prev = j > 0 ? keys_sorted[j - 1] : -1;
    if (cur < 0) cur = m;                // rejected points (-1) sort LAST (partial-bit sort of pcab_pillar_index): they
    if (j > 0 && prev < 0) prev = m;     // form one trailing run behind the last pillar
    for (int q = prev + 1; q <= cur && q <= m; ++q) pstart[q] = j;
  }
}

// one thread per pillar, sequential in stream order -> bit-identical to a CPU scatter_add
// 8 lanes per pillar: the lanes fetch eight points of the pillar at once (the gathers through `order` are what the time goes
// into), lane 0 of the group then adds them in stream order from registers -- the same sequential rounding sequence as
// a CPU scatter_add, so the means stay bit-identical to the reference's.
__global__ void __launch_bounds__(256) k_pillar_stats(const float* __restrict__ xyz, const long long* __restrict__ fb_labels,
                                                      const int* __restrict__ order, const int* __restrict__ pstart, int m,
                                                      float* __restrict__ pillar_mean, int* __restrict__ fb_sub) {
  const int sub = threadIdx.x & 7;
  const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
  const int groups = (gridDim.x * blockDim.x) >> 3;
  const int m8 = (m + 7) & ~7;  // (every lane of a warp runs the same number of iterations: the shuffles below are warp-wide)
  for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; p < m8; p += groups) {
    const bool live = p < m;
    const int s = live ? pstart[p] : 0, e = live ? pstart[p + 1] : 0;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    long long mx = 0;
    bool any = false;
    const int iters = (e - s + 7) >> 3;
    int max_iters = iters;
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) max_iters = max(max_iters, __shfl_xor_sync(0xffffffffu, max_iters, o));
    for (int it = 0; it < max_iters; ++it) {
      const int j = s + 8 * it + sub;
      float x = 0.f, y = 0.f, z = 0.f;
      long long v = 0;
      if (j < e) {
        const int i = order[j];
        x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        if (fb_labels) v = fb_labels[i];
      }
      const int cnt = min(8, e - (s + 8 * it));
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xk = __shfl_sync(0xffffffffu, x, (threadIdx.x & 24) + k), yk = __shfl_sync(0xffffffffu, y, (threadIdx.x & 24) + k),
                    zk = __shfl_sync(0xffffffffu, z, (threadIdx.x & 24) + k);
        const long long vk = __shfl_sync(0xffffffffu, v, (threadIdx.x & 24) + k);
        if (k < cnt) {
          sx = __fadd_rn(sx, xk), sy = __fadd_rn(sy, yk), sz = __fadd_rn(sz, zk);
          mx = any ? (vk > mx ? vk : mx) : vk;
          any = true;
        }
      }
    }
    (void)gmask;
    if (live && sub == 0) {
      const int cnt = e - s;
      const float c = (float)(cnt > 0 ? cnt : 1);
      pillar_mean[3 * p + 0] = __fdiv_rn(sx, c);
      pillar_mean[3 * p + 1] = __fdiv_rn(sy, c);
      pillar_mean[3 * p + 2] = __fdiv_rn(sz, c);
      if (fb_sub) fb_sub[p] = (int)mx;
    }
  }
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t pcab_voxelize_workspace(int n_points, long long n_cells) {
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int*)nullptr, (int*)nullptr, n_points);
  return align256((size_t)n_cells * 4) + 3 * align256((size_t)n_points * 4) + align256(scan_bytes) + 256;
}

extern "C" int pcab_voxelize(const float* points4, const int* point_batch, int n_points, int batch_size,
                             const float* range6, const float* voxel_size3, int n_sweeps, int* coords_zyxt,
                             int* pillar_batch, int* p2v, int* num_voxels, int* total_voxels, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream) {
  PCAB_REQUIRE(n_points > 0 && batch_size > 0, "empty input");
  VoxGeom g;
  for (int j = 0; j < 3; ++j) {
    g.lo[j] = range6[j];
    g.vs[j] = voxel_size3[j];
    // grid = round((hi - lo) / vs) in float32 (libs/voxel_generator.py:27-28)
    g.grid[j] = (int)nearbyintf((range6[3 + j] - range6[j]) / voxel_size3[j]);
  }
  g.nt = n_sweeps;
  long long n_cells = (long long)batch_size * g.grid[0] * g.grid[1] * g.grid[2] * n_sweeps;
  PCAB_REQUIRE(n_cells < (1LL << 31), "cell table too large");
  PCAB_REQUIRE(workspace_bytes >= pcab_voxelize_workspace(n_points, n_cells), "workspace too small");
  char* w = (char*)workspace;
  int* first = (int*)w;
  w += align256((size_t)n_cells * 4);
  int* cell_of = (int*)w;
  w += align256((size_t)n_points * 4);
  int* flag = (int*)w;
  w += align256((size_t)n_points * 4);
  int* rank = (int*)w;
  w += align256((size_t)n_points * 4);
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flag, rank, n_points);
  void* scan_tmp = w;

  PCAB_CUDA(cudaMemsetAsync(first, 0x7f, (size_t)n_cells * 4, stream));
  PCAB_CUDA(cudaMemsetAsync(num_voxels, 0, (size_t)batch_size * 4, stream));
  const int B = 256;
  int gsz = grid_for(n_points, B);
  k_cell<<<gsz, B, 0, stream>>>(points4, point_batch, n_points, g, cell_of, first);
  k_flag<<<gsz, B, 0, stream>>>(cell_of, first, n_points, flag);
  PCAB_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, flag, rank, n_points, stream));
  k_assign<<<gsz, B, 0, stream>>>(cell_of, flag, rank, n_points, g, first, coords_zyxt, pillar_batch, num_voxels,
                                  total_voxels);
  k_map<<<gsz, B, 0, stream>>>(cell_of, first, n_points, p2v);
  PCAB_CHECK_LAUNCH("pcab_voxelize");
  return PCAB_OK;
}

extern "C" size_t pcab_pillar_index_workspace(int n_points) {
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr,
                                  n_points);
  return align256(sort_bytes) + 2 * align256((size_t)n_points * 4) + 256;
}

// order[j] = original index of the j-th point after a STABLE sort by pillar id; pstart[m..m+1] = its segment.
extern "C" int pcab_pillar_index(const int* p2v, int n_points, int n_pillars, int* order, int* pstart, void* workspace,
                                 size_t workspace_bytes, cudaStream_t stream) {
  PCAB_REQUIRE(workspace_bytes >= pcab_pillar_index_workspace(n_points), "workspace too small");
  char* w = (char*)workspace;
  int* iota = (int*)w;
  w += align256((size_t)n_points * 4);
  int* keys_sorted = (int*)w;
  w += align256((size_t)n_points * 4);
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, p2v, keys_sorted, iota, order, n_points);
  const int B = 256;
  k_iota<<<grid_for(n_points, B), B, 0, stream>>>(iota, n_points);
  // pillar ids are < n_pillars: only their significant bits (+1) take part in the sort -- 3 radix passes instead of 4 at C2.
  // A rejected point carries id -1 (all bits set): in the low (bits + 1) bits it is larger than every valid id, so such
  // points still end up together behind the last pillar, as with the full-width signed sort.
  int end_bit = 1;
  while ((1LL << end_bit) <= (long long)n_pillars && end_bit < 31) ++end_bit;
  end_bit = end_bit + 1 > 32 ? 32 : end_bit + 1;
  PCAB_CUDA(cub::DeviceRadixSort::SortPairs(w, sort_bytes, p2v, keys_sorted, iota, order, n_points, 0, end_bit, stream));
  k_segment_starts<<<grid_for(n_points + 1, B), B, 0, stream>>>(keys_sorted, n_points, n_pillars, pstart);
  PCAB_CHECK_LAUNCH("pcab_pillar_index");
  return PCAB_OK;
}

extern "C" int pcab_pillar_stats(const float* xyz, const long long* fb_labels, const int* order, const int* pstart,
                                 int n_pillars, float* pillar_mean, int* fb_sub, cudaStream_t stream) {
  const int B = 128;
  k_pillar_stats<<<grid_for((long long)n_pillars * 8, 256, 16), 256, 0, stream>>>(xyz, fb_labels, order, pstart, n_pillars,
                                                                                 pillar_mean, fb_sub);
  PCAB_CHECK_LAUNCH("pcab_pillar_stats");
  return PCAB_OK;
}
