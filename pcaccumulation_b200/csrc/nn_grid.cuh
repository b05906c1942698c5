// Exact nearest-neighbour search over a uniform grid (internal interface shared by chamfer.cu and nn_grid.cu).
//
// The brute-force nearest neighbour of the reference (chamfer_distance.cpp:59-84, chamfer_distance.cu:6-136) evaluates
// all n*m pairs (2.4e11 for the nuScenes-sized alignment error of models/tpointnet.py:145-163).  The RESULT is defined
// without reference to the search order: min over targets of (d, index) with d = (dx*dx + dy*dy) + dz*dz in float32 and
// the lowest index winning ties.  Any search that provably visits every target that can attain that minimum returns the
// same bits, so the targets are binned into a uniform grid (counting sort, cells of a row contiguous in memory) and
// each query walks Chebyshev rings of cells around its own cell until the lower bound of everything unvisited exceeds
// the best distance found (with a margin that covers the float32 rounding of the cell assignment).
#pragma once
#include "common.cuh"

namespace nngrid {

struct Header {     // device-resident, first bytes of a grid workspace; written by k_params
  float ox, oy, oz;  // origin = lower corner of the targets' bounding box
  float h, inv_h;    // cubic cell
  int dx, dy, dz;    // cells per axis (<= 1024 each)
  int valid;         // 0: empty / non-finite bounding box -> every query is sent to the brute-force list
  unsigned bbox[6];  // order-preserving encodings of min x,y,z and max x,y,z
};

constexpr int MAX_DIM = 1024;
constexpr int MAX_RINGS = 24;  // queries not settled after this many rings go to the brute-force list (unbounded search only)

static inline int max_cells(int m) {
  long long c = 2LL * (m > 0 ? m : 1);
  if (c < 4096) c = 4096;
  if (c > (1LL << 22)) c = 1LL << 22;
  return (int)c;
}

struct Layout {
  size_t header, count, start, sorted, scan_tmp, total;
  int ncell_max;
};
Layout layout(int m);

// Build the grid of `m` target points (xyz rows) inside `ws` (>= layout(m).total bytes, 256-byte aligned).
// cell_hint > 0 fixes the cell edge (ICP: the correspondence threshold), otherwise it is chosen for ~1 point per cell of the
// bounding volume.
int build(const float* targets, int m, float cell_hint, void* ws, cudaStream_t stream);

// sorted target records of a built grid: float4 (x, y, z, original index as int bits), grouped by cell
static inline const float4* sorted_points(const void* ws, int m) { return (const float4*)((const char*)ws + layout(m).sorted); }

// Nearest target of every query.  Queries come either as xyz rows (`queries`, thread i = query i) or as the sorted
// records of another grid (`qsorted`: spatially coherent warps; results are written to the record's original index).
// best[i] = (float bits of d) << 32 | target index.  max_dist > 0 bounds the search (no match: best[i] = ~0);
// max_dist <= 0 is the exact unbounded search: unsettled queries are appended to fallback_list / fallback_count.
// tsfm (optional, 12 floats R|t row-major on the device) is applied to the queries first.
int query(const void* grid_ws, int m, const float* queries, const float4* qsorted, int n, float max_dist, const float* tsfm,
          unsigned long long* best, int* fallback_list, int* fallback_count, cudaStream_t stream);

}  // namespace nngrid
