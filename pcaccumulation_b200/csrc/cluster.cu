// Instance proposal on the GPU: offset-shifted dynamic points -> 5 cm hash de-duplication -> DBSCAN
// (eps / min_samples) with sklearn's label semantics -> small-cluster rejection -> canonical labels.
//
// Replaces models/cluster.py:9-13 (voxel_downsample), :23-49 (cluster), :52-84 (cluster_per_batch),
// torchsparse.utils.quantize.sparse_quantize (spec: dataset_toolbox/prep_nuscene_waymo_sf/libs/
// spv_utils.py:7-20,57-60,81), sklearn.cluster.DBSCAN and toolbox/utils.py:237-250
// (canonicalise_random_indice).  The reference round-trips through host numpy/sklearn here; this
// version stays on the device.
//
// Semantics reproduced exactly:
//   * de-dup key = ravel hash of floor(round(q / 0.05)) over x,y,z; survivor = FIRST occurrence of each key,
//     survivors ordered by ascending key (np.unique);
//   * DBSCAN on the survivors with z := 0: core = >= min_samples neighbours within eps INCLUDING self
//     (float64 distances); clusters numbered by ascending index of their lowest-index core point; a border
//     point joins the lowest-numbered cluster that has a core neighbour of it; noise = -1;
//   * clusters with < min_p_cluster survivors -> noise; kept clusters renumbered 1..L in ascending order,
//     noise/background = 0.
#include <cub/cub.cuh>
#include "common.cuh"
#include "pcab200.h"

namespace {

struct Counts {  // device-side bookkeeping
  int mn[3], mx[3];
  int n_unique;
  int n_clusters;
  int n_kept;
};

__global__ void k_dyn_flag(const float* __restrict__ mos, int n0, int n, int* __restrict__ flag) {
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    flag[i] = mos[2 * (n0 + i) + 1] > mos[2 * (n0 + i)] ? 1 : 0;  // argmax == 1 (first max wins ties)
}

__global__ void k_quant(const float* __restrict__ tp, const float* __restrict__ off, const int* __restrict__ sel, int n0,
                        int s, float voxel, float* __restrict__ q, int* __restrict__ c, Counts* cnt) {
  int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < s; j += stride) {
    int i = n0 + sel[j];
    float x = __fadd_rn(tp[3 * i], off[2 * i]), y = __fadd_rn(tp[3 * i + 1], off[2 * i + 1]), z = tp[3 * i + 2];
    q[3 * j] = x, q[3 * j + 1] = y, q[3 * j + 2] = z;
    int cx = (int)rintf(__fdiv_rn(x, voxel)), cy = (int)rintf(__fdiv_rn(y, voxel)), cz = (int)rintf(__fdiv_rn(z, voxel));
    c[3 * j] = cx, c[3 * j + 1] = cy, c[3 * j + 2] = cz;
    atomicMin(&cnt->mn[0], cx), atomicMin(&cnt->mn[1], cy), atomicMin(&cnt->mn[2], cz);
    atomicMax(&cnt->mx[0], cx), atomicMax(&cnt->mx[1], cy), atomicMax(&cnt->mx[2], cz);
  }
}

__global__ void k_init_counts(Counts* c) {
  for (int k = 0; k < 3; ++k) c->mn[k] = INT_MAX, c->mx[k] = INT_MIN;
  c->n_unique = c->n_clusters = c->n_kept = 0;
}

__global__ void k_hash(const int* __restrict__ c, int s, const Counts* __restrict__ cnt,
                       unsigned long long* __restrict__ key, int* __restrict__ val) {
  int stride = gridDim.x * blockDim.x;
  unsigned long long ry = (unsigned long long)(cnt->mx[1] - cnt->mn[1]) + 1ull;
  unsigned long long rz = (unsigned long long)(cnt->mx[2] - cnt->mn[2]) + 1ull;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < s; j += stride) {
    unsigned long long h = (unsigned long long)(c[3 * j] - cnt->mn[0]);
    h *= ry;
    h += (unsigned long long)(c[3 * j + 1] - cnt->mn[1]);
    h *= rz;
    h += (unsigned long long)(c[3 * j + 2] - cnt->mn[2]);
    key[j] = h;
    val[j] = j;
  }
}

__global__ void k_heads(const unsigned long long* __restrict__ key_sorted, int s, int* __restrict__ head) {
  int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < s; j += stride)
    head[j] = (j == 0 || key_sorted[j] != key_sorted[j - 1]) ? 1 : 0;
}

// uid[j] (inclusive-scan(head) - 1) -> survivors (xy of the run head) and inverse map
__global__ void k_unique(const int* __restrict__ head, const int* __restrict__ rank_excl,
                         const int* __restrict__ val_sorted, const float* __restrict__ q, int s, float* __restrict__ pxy,
                         int* __restrict__ inverse, Counts* cnt) {
  int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < s; j += stride) {
    int u = rank_excl[j] + head[j] - 1;
    int orig = val_sorted[j];
    inverse[orig] = u;
    if (head[j]) pxy[2 * u] = q[3 * orig], pxy[2 * u + 1] = q[3 * orig + 1];
    if (j == s - 1) cnt->n_unique = u + 1;
  }
}

__device__ __forceinline__ unsigned int cell_key(int cx, int cy) { return ((unsigned int)cy << 16) | (unsigned int)cx; }

__global__ void k_cellkeys(const float* __restrict__ pxy, const Counts* __restrict__ cnt, float voxel, float eps,
                           int cap, unsigned int* __restrict__ ckey, int* __restrict__ cval) {
  int U = cnt->n_unique;
  float x0 = (float)cnt->mn[0] * voxel - 1.f, y0 = (float)cnt->mn[1] * voxel - 1.f;
  int stride = gridDim.x * blockDim.x;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < cap; u += stride) {
    if (u < U) {
      int cx = (int)floorf((pxy[2 * u] - x0) / eps) + 1, cy = (int)floorf((pxy[2 * u + 1] - y0) / eps) + 1;
      ckey[u] = cell_key(cx, cy);
    } else {
      ckey[u] = 0xffffffffu;  // padding sorts last
    }
    cval[u] = u;
  }
}

__device__ __forceinline__ int lower_bound(const unsigned int* a, int n, unsigned int v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Neighbour search is shared by EIGHT lanes per point (a scene has only ~10^4..10^5 de-duplicated dynamic points: one thread
// per point leaves the SMs nearly empty).  The three cells (cx-1 .. cx+1) of a grid row are consecutive keys, so the 3x3
// block is THREE contiguous ranges of the cell-sorted order: k_ranges finds them once per point (6 binary searches) and every
// pass re-reads the 6 ints instead of searching 9 cells again; the eight lanes stride through each range together, reading the
// cell-sorted copy of the coordinates (coalesced).  The float64 distance test of sklearn is only evaluated for pairs within
// 1e-5 (relative) of the radius; everything else is decided by a float32 squared distance whose rounding error is < 1e-6.
constexpr int kParts = 8;

struct NbrCtx {
  const float* pxy;         // coordinates by point id
  const float2* pxy_s;      // the same in cell-sorted order
  const int* cval_sorted;   // point id of a sorted slot
  const int* rng;           // [U][6]: (begin, end) of the three cell rows around a point
  double eps;
  float eps2_lo, eps2_hi;
};

template <typename F>
__device__ __forceinline__ void for_neighbors(int u, const NbrCtx& c, int part, F&& fn) {
  const float xf = c.pxy[2 * u], yf = c.pxy[2 * u + 1];
  const double x = xf, y = yf;
  const int2* r = reinterpret_cast<const int2*>(c.rng + 6 * (size_t)u);
#pragma unroll
  for (int row = 0; row < 3; ++row) {
    const int2 ab = r[row];
    for (int j = ab.x + part; j < ab.y; j += kParts) {
      const float2 q = c.pxy_s[j];
      const float ex = q.x - xf, ey = q.y - yf;
      const float d2 = ex * ex + ey * ey;
      bool in = d2 < c.eps2_lo;
      if (!in && d2 <= c.eps2_hi) {
        const double dx = (double)q.x - x, dy = (double)q.y - y;
        in = sqrt(dx * dx + dy * dy) <= c.eps;
      }
      if (in) fn(c.cval_sorted[j]);
    }
  }
}

// cell-sorted coordinates + the three row ranges of every point (slot j of the sorted order = point cval_sorted[j])
__global__ void k_ranges(const float* __restrict__ pxy, const unsigned int* __restrict__ ckey_sorted,
                         const int* __restrict__ cval_sorted, const Counts* __restrict__ cnt, float2* __restrict__ pxy_s,
                         int* __restrict__ rng, int* __restrict__ key_of) {
  const int U = cnt->n_unique;
  const int stride = gridDim.x * blockDim.x;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 3 * U; t += stride) {
    const int j = t / 3, row = t % 3;
    const int u = cval_sorted[j];
    const unsigned int key = ckey_sorted[j];
    const int cx = key & 0xffff, cy = key >> 16;
    rng[6 * (size_t)u + 2 * row] = lower_bound(ckey_sorted, U, cell_key(cx - 1, cy + row - 1));
    rng[6 * (size_t)u + 2 * row + 1] = lower_bound(ckey_sorted, U, cell_key(cx + 2, cy + row - 1));
    if (row == 0) {
      pxy_s[j] = make_float2(pxy[2 * u], pxy[2 * u + 1]);
      key_of[u] = (int)key;
    }
  }
}

// warp-uniform loop over groups of kParts lanes: j0 is the first point of this warp's current batch of 32/kParts points
#define PCAB_GROUP_LOOP(j, ok, part, U)                                                                   \
  const int part = threadIdx.x & (kParts - 1), grp_in_warp = (threadIdx.x & 31) / kParts;                 \
  const int grp_stride = (gridDim.x * blockDim.x) / kParts;                                               \
  for (int j0 = (blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / kParts, j = j0 + grp_in_warp, ok = j < (U); j0 < (U); \
       j0 += grp_stride, j = j0 + grp_in_warp, ok = j < (U))

__device__ __forceinline__ int group_sum(int v) {
#pragma unroll
  for (int o = 1; o < kParts; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int group_min(int v) {
#pragma unroll
  for (int o = 1; o < kParts; o <<= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void k_core(NbrCtx nc, const Counts* __restrict__ cnt, int min_samples, int* __restrict__ core,
                       int* __restrict__ parent) {
  const int U = cnt->n_unique;
  PCAB_GROUP_LOOP(u, ok, part, U) {
    int c = 0;
    if (ok) for_neighbors(u, nc, part, [&](int) { ++c; });
    c = group_sum(c);
    if (ok && part == 0) {
      core[u] = c >= min_samples;
      parent[u] = u;
    }
  }
}

// Union-find over the core points.  Roots only ever move to smaller indices (the larger root is linked under the
// smaller one), so the root of a finished component is its lowest-index core point whatever the order of the unions.
// Reads go to L2 (ld.global.cg): a stale L1 line could show a node as its own root long after it was linked and make
// the CAS loop spin until the line happens to be evicted.
__device__ __forceinline__ int uf_find(int* parent, int i) {
  while (true) {
    int p = __ldcg(parent + i);
    if (p == i) return i;
    int gp = __ldcg(parent + p);
    if (gp != p) parent[i] = gp;  // path halving: gp is an ancestor of i at every moment, so the store is always valid
    i = p;
  }
}

// links the roots of a and b (larger index under the smaller one) and returns the surviving root
__device__ __forceinline__ int uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a), b = uf_find(parent, b);
    if (a == b) return a;
    if (a < b) { int t = a; a = b; b = t; }
    int old = atomicCAS(parent + a, a, b);
    if (old == a) return b;
  }
}

// Hooking pass: every core point points at its lowest-index core neighbour (itself included).  Pointers only go to
// smaller indices, so this is a forest; after k_roots has flattened it, a dense cluster consists of a handful of trees
// (one per local index minimum) and k_union only has to stitch those together.
__global__ void k_hook_min(NbrCtx nc, const Counts* __restrict__ cnt, const int* __restrict__ core, int* __restrict__ parent) {
  const int U = cnt->n_unique;
  PCAB_GROUP_LOOP(u, ok, part, U) {
    int m = INT_MAX;
    const bool is_core = ok && core[u];
    if (is_core)
      for_neighbors(u, nc, part, [&](int v) {
        if (core[v]) m = min(m, v);
      });
    m = group_min(m);
    if (is_core && part == 0) parent[u] = min(m, u);
  }
}

// Stitching pass over the flattened forest: a lane remembers the root its point started under (r0) and its current root
// (ru) and skips every neighbour whose parent pointer equals either - one L2 load per pair.  Only pairs that straddle two
// trees reach the union-find proper (two dependent pointer chases and a CAS attempt).
__global__ void k_union(NbrCtx nc, const Counts* __restrict__ cnt, const int* __restrict__ core, int* parent) {
  const int U = cnt->n_unique;
  PCAB_GROUP_LOOP(u, ok, part, U) {
    if (!ok || !core[u]) continue;
    const int r0 = __ldcg(parent + u);
    int ru = r0;
    for_neighbors(u, nc, part, [&](int v) {
      if (v < u && core[v]) {
        // pre-check through L1: a stale line can only show an OLDER parent of v, and "parent[v] was ru (or r0) at some moment"
        // already proves that v is in u's component (components only ever merge), so skipping on it is safe; a mismatch goes
        // to the union-find proper, which reads through L2
        int pv = parent[v];
        if (pv != ru && pv != r0) ru = uf_union(parent, ru, pv);
      }
    });
  }
}

// parent[u] := root(u) for every core point.  Read-only walk (no path halving here: a halving store could land AFTER
// another thread has finalised that entry and replace the root by an intermediate ancestor, which k_labels would then
// read as a cluster id); concurrent finalising stores only ever shorten the walk.
__global__ void k_roots(const Counts* __restrict__ cnt, const int* __restrict__ core, int* parent,
                        int* __restrict__ is_root, int cap) {
  int U = cnt->n_unique;
  int stride = gridDim.x * blockDim.x;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < cap; u += stride) {
    int r = 0;
    if (u < U && core[u]) {
      int root = u;
      for (int p = __ldcg(parent + root); p != root; p = __ldcg(parent + root)) root = p;
      parent[u] = root;
      r = root == u;
    }
    is_root[u] = r;
  }
}

// label[u]: core -> cluster number of its root; border -> min cluster number among core neighbours; noise -> -1
__global__ void k_labels(NbrCtx nc, Counts* cnt, const int* __restrict__ core, const int* __restrict__ parent,
                         const int* __restrict__ root_rank, const int* __restrict__ is_root, int cap, int* __restrict__ label,
                         int* __restrict__ size) {
  const int U = cnt->n_unique;
  PCAB_GROUP_LOOP(u, ok, part, U) {
    int best = INT_MAX;
    const bool is_core = ok && core[u];
    if (ok && !is_core)
      for_neighbors(u, nc, part, [&](int v) {
        if (core[v]) best = min(best, root_rank[parent[v]]);
      });
    best = group_min(best);
    if (ok && part == 0) {
      int l = is_core ? root_rank[parent[u]] : (best != INT_MAX ? best : -1);
      label[u] = l;  // (cluster sizes: k_sizes)
      if (u == 0) cnt->n_clusters = root_rank[cap - 1] + is_root[cap - 1];
    }
  }
}

// size[l] = number of points labelled l.  One thread per point, lanes with equal labels combine before the atomic (all
// points of a big cluster would otherwise hit one counter).
__global__ void k_sizes(const int* __restrict__ label, const Counts* __restrict__ cnt, int cap, int* __restrict__ size) {
  const int U = cnt->n_unique;
  const int stride = gridDim.x * blockDim.x;
  for (int u0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); u0 < U; u0 += stride) {
    const int u = u0 + (threadIdx.x & 31);
    const int l = u < U ? label[u] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, l);
    if (l >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(size + l, __popc(peers));
  }
}

__global__ void k_keep(const int* __restrict__ size, const Counts* __restrict__ cnt, int min_p, int cap,
                       int* __restrict__ keep) {
  int stride = gridDim.x * blockDim.x;
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < cap; l += stride)
    keep[l] = (l < cnt->n_clusters && size[l] >= min_p) ? 1 : 0;
}

__global__ void k_final(const int* __restrict__ label, const int* __restrict__ keep, const int* __restrict__ keep_rank,
                        const int* __restrict__ inverse, const int* __restrict__ sel, int n0, int s, int cap, Counts* cnt,
                        long long* __restrict__ inst_out) {
  int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < s; j += stride) {
    int l = label[inverse[j]];
    inst_out[n0 + sel[j]] = (l >= 0 && keep[l]) ? (long long)(keep_rank[l] + 1) : 0;
    if (j == 0) cnt->n_kept = keep_rank[cap - 1] + keep[cap - 1];
  }
}

size_t alc(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t pcab_cluster_workspace(int s) {
  size_t sort64 = 0, sort32 = 0, scan = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort64, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (int*)nullptr, (int*)nullptr, s);
  cub::DeviceRadixSort::SortPairs(nullptr, sort32, (unsigned int*)nullptr, (unsigned int*)nullptr, (int*)nullptr,
                                  (int*)nullptr, s);
  cub::DeviceScan::ExclusiveSum(nullptr, scan, (int*)nullptr, (int*)nullptr, s);
  size_t tmp = sort64 > sort32 ? sort64 : sort32;
  if (scan > tmp) tmp = scan;
  // must mirror the take() sequence of pcab_cluster_scene: q, c (12 B/row) | key, key_s (8) | val, val_s, head, rank (4) |
  // pxy (8) | spare, inverse, ckey, ckey_s, cval, cval_s, key_of, core, parent, is_root, root_rank, label, size, keep,
  // keep_rank (15 x 4) | pxy_s (8) | rng (24) | counts | sort/scan temp
  return 2 * alc((size_t)s * 12) + 2 * alc((size_t)s * 8) + 4 * alc((size_t)s * 4) + alc((size_t)s * 8) +
         15 * alc((size_t)s * 4) + alc((size_t)s * 8) + alc((size_t)s * 24) + alc(sizeof(Counts)) + alc(tmp) + 1024;
}

// flags[i] = argmax(mos[n0+i]) == 1 for i in [0, n)
extern "C" int pcab_dynamic_flags(const float* mos, int n0, int n, int* flags, cudaStream_t stream) {
  k_dyn_flag<<<grid_for(n, 256), 256, 0, stream>>>(mos, n0, n, flags);
  PCAB_CHECK_LAUNCH("pcab_dynamic_flags");
  return PCAB_OK;
}

// sel[0..s): ascending indices (relative to n0) of the selected points of one scene.  Writes inst_out[n0+sel[j]]
// (the caller zero-fills inst_out) and counts_out[0] = number of kept clusters (device int).
extern "C" int pcab_cluster_scene(const float* transformed_points, const float* offset, const int* sel, int n0, int s,
                                  float dedupe_voxel, double eps, int min_samples, int min_p_cluster,
                                  long long* inst_out, int* n_instances_out, void* workspace, size_t workspace_bytes,
                                  cudaStream_t stream) {
  PCAB_REQUIRE(s > 0, "empty selection");
  PCAB_REQUIRE(workspace_bytes >= pcab_cluster_workspace(s), "workspace too small");
  char* w = (char*)workspace;
  auto take = [&](size_t bytes) {
    char* p = w;
    w += alc(bytes);
    return (void*)p;
  };
  float* q = (float*)take((size_t)s * 12);
  int* c = (int*)take((size_t)s * 12);
  unsigned long long* key = (unsigned long long*)take((size_t)s * 8);
  unsigned long long* key_s = (unsigned long long*)take((size_t)s * 8);
  int* val = (int*)take((size_t)s * 4);
  int* val_s = (int*)take((size_t)s * 4);
  int* head = (int*)take((size_t)s * 4);
  int* rank = (int*)take((size_t)s * 4);
  float* pxy = (float*)take((size_t)s * 8);
  (void)take((size_t)s * 4);
  int* inverse = (int*)take((size_t)s * 4);
  unsigned int* ckey = (unsigned int*)take((size_t)s * 4);
  unsigned int* ckey_s = (unsigned int*)take((size_t)s * 4);
  int* cval = (int*)take((size_t)s * 4);
  int* cval_s = (int*)take((size_t)s * 4);
  int* key_of = (int*)take((size_t)s * 4);
  int* core = (int*)take((size_t)s * 4);
  int* parent = (int*)take((size_t)s * 4);
  int* is_root = (int*)take((size_t)s * 4);
  int* root_rank = (int*)take((size_t)s * 4);
  int* label = (int*)take((size_t)s * 4);
  int* size = (int*)take((size_t)s * 4);
  int* keep = (int*)take((size_t)s * 4);
  int* keep_rank = (int*)take((size_t)s * 4);
  float2* pxy_s = (float2*)take((size_t)s * 8);
  int* rng = (int*)take((size_t)s * 24);
  Counts* cnt = (Counts*)take(sizeof(Counts));
  size_t sort64 = 0, sort32 = 0, scan = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort64, key, key_s, val, val_s, s);
  cub::DeviceRadixSort::SortPairs(nullptr, sort32, ckey, ckey_s, cval, cval_s, s);
  cub::DeviceScan::ExclusiveSum(nullptr, scan, head, rank, s);
  void* tmp = w;
  size_t tmp_need = sort64 > sort32 ? sort64 : sort32;
  if (scan > tmp_need) tmp_need = scan;
  PCAB_REQUIRE((size_t)(w - (char*)workspace) + tmp_need <= workspace_bytes, "workspace carve exceeds the buffer");

  const int B = 256;
  int g = grid_for(s, B);
  int g4 = grid_for((long long)s * kParts, B);
  k_init_counts<<<1, 1, 0, stream>>>(cnt);
  k_quant<<<g, B, 0, stream>>>(transformed_points, offset, sel, n0, s, dedupe_voxel, q, c, cnt);
  k_hash<<<g, B, 0, stream>>>(c, s, cnt, key, val);
  PCAB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, sort64, key, key_s, val, val_s, s, 0, 64, stream));
  k_heads<<<g, B, 0, stream>>>(key_s, s, head);
  PCAB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, scan, head, rank, s, stream));
  k_unique<<<g, B, 0, stream>>>(head, rank, val_s, q, s, pxy, inverse, cnt);
  k_cellkeys<<<g, B, 0, stream>>>(pxy, cnt, dedupe_voxel, (float)eps * 1.001f, s, ckey, cval);
  PCAB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, sort32, ckey, ckey_s, cval, cval_s, s, 0, 32, stream));
  NbrCtx nc;
  nc.pxy = pxy, nc.pxy_s = pxy_s, nc.cval_sorted = cval_s, nc.rng = rng, nc.eps = eps;
  nc.eps2_lo = (float)(eps * eps * (1.0 - 1e-5)), nc.eps2_hi = (float)(eps * eps * (1.0 + 1e-5));
  k_ranges<<<grid_for(3LL * s, B), B, 0, stream>>>(pxy, ckey_s, cval_s, cnt, pxy_s, rng, key_of);
  k_core<<<g4, B, 0, stream>>>(nc, cnt, min_samples, core, parent);
  k_hook_min<<<g4, B, 0, stream>>>(nc, cnt, core, parent);
  k_roots<<<g, B, 0, stream>>>(cnt, core, parent, is_root, s);  // flatten the hooking forest
  k_union<<<g4, B, 0, stream>>>(nc, cnt, core, parent);
  k_roots<<<g, B, 0, stream>>>(cnt, core, parent, is_root, s);
  PCAB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, scan, is_root, root_rank, s, stream));
  PCAB_CUDA(cudaMemsetAsync(size, 0, (size_t)s * 4, stream));
  k_labels<<<g4, B, 0, stream>>>(nc, cnt, core, parent, root_rank, is_root, s, label, size);
  k_sizes<<<g, B, 0, stream>>>(label, cnt, s, size);
  k_keep<<<g, B, 0, stream>>>(size, cnt, min_p_cluster, s, keep);
  PCAB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, scan, keep, keep_rank, s, stream));
  k_final<<<g, B, 0, stream>>>(label, keep, keep_rank, inverse, sel, n0, s, s, cnt, inst_out);
  PCAB_CUDA(cudaMemcpyAsync(n_instances_out, &cnt->n_kept, sizeof(int), cudaMemcpyDeviceToDevice, stream));
  PCAB_CHECK_LAUNCH("pcab_cluster_scene");
  return PCAB_OK;
}
