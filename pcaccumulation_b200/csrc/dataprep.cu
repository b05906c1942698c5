// GPU data front-end: scene crop + ground removal as ONE stable stream compaction over the
// raw sample arrays, feeding the voxeliser directly.
//
// Replaces steps 2-4 of libs/dataset.py:163-207 (BaseDataset.prep_input without augmentation: |x|,|y| < crop_xy,
// crop_z_min < z < crop_z_max, then z > ground_height + ground_slack), which the reference runs with six boolean-mask
// gathers per step in the DataLoader workers.  Comparisons are in float32 like numpy's (float32 array vs Python float).
// The kept points leave as (x, y, z, t) float32 rows - the voxeliser's input - with their labels, in the original order.
#include <cub/cub.cuh>
#include "common.cuh"
#include "pcab200.h"

namespace {

struct PrepArgs {
  const float* pts;
  const long long* tidx;
  const long long* sd;
  const long long* fb;
  const long long* inst;
  int n;
  float crop_xy, z_min, z_max, ground;
  int remove_ground;
};

__device__ __forceinline__ bool keep(const PrepArgs& a, int i) {
  const float x = a.pts[3 * i], y = a.pts[3 * i + 1], z = a.pts[3 * i + 2];
  bool k = fabsf(x) < a.crop_xy && fabsf(y) < a.crop_xy && z < a.z_max && z > a.z_min;
  if (a.remove_ground) k = k && z > a.ground;
  return k;
}

__global__ void k_prep_flags(PrepArgs a, int* __restrict__ flag) {
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) flag[i] = keep(a, i) ? 1 : 0;
}

__global__ void k_prep_scatter(PrepArgs a, const int* __restrict__ flag, const int* __restrict__ pos, float4* __restrict__ points4,
                               int* __restrict__ t32, long long* __restrict__ sd, long long* __restrict__ fb,
                               long long* __restrict__ inst, int* __restrict__ count) {
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    if (flag[i]) {
      const int j = pos[i];
      const long long t = a.tidx[i];
      points4[j] = make_float4(a.pts[3 * i], a.pts[3 * i + 1], a.pts[3 * i + 2], (float)t);
      t32[j] = (int)t;
      sd[j] = a.sd[i], fb[j] = a.fb[i], inst[j] = a.inst[i];
    }
    if (i == a.n - 1) *count = pos[i] + flag[i];
  }
}

size_t al256d(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t pcab_prep_points_workspace(int n) {
  size_t scan = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan, (int*)nullptr, (int*)nullptr, n);
  return 2 * al256d((size_t)n * 4) + al256d(scan) + 256;
}

// Outputs hold up to n rows; count_out (device int) receives the number of kept points.
extern "C" int pcab_prep_points(const float* raw_points, const long long* time_idx, const long long* sd_labels,
                                const long long* fb_labels, const long long* inst_labels, int n, float crop_xy, float crop_z_min,
                                float crop_z_max, int remove_ground, float ground_height, float* points4_out, int* time_out,
                                long long* sd_out, long long* fb_out, long long* inst_out, int* count_out, void* workspace,
                                size_t workspace_bytes, cudaStream_t stream) {
  if (n <= 0) {
    PCAB_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int), stream));
    return PCAB_OK;
  }
  PCAB_REQUIRE(workspace_bytes >= pcab_prep_points_workspace(n), "workspace too small");
  PCAB_REQUIRE(((uintptr_t)points4_out & 15) == 0, "points4_out must be 16B aligned");
  char* w = (char*)workspace;
  int* flag = (int*)w;
  w += al256d((size_t)n * 4);
  int* pos = (int*)w;
  w += al256d((size_t)n * 4);
  size_t scan = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan, flag, pos, n);
  PrepArgs a;
  a.pts = raw_points, a.tidx = time_idx, a.sd = sd_labels, a.fb = fb_labels, a.inst = inst_labels, a.n = n;
  a.crop_xy = crop_xy, a.z_min = crop_z_min, a.z_max = crop_z_max, a.ground = ground_height, a.remove_ground = remove_ground;
  k_prep_flags<<<grid_for(n, 256), 256, 0, stream>>>(a, flag);
  PCAB_CUDA(cub::DeviceScan::ExclusiveSum(w, scan, flag, pos, n, stream));
  k_prep_scatter<<<grid_for(n, 256), 256, 0, stream>>>(a, flag, pos, (float4*)points4_out, time_out, sd_out, fb_out, inst_out,
                                                       count_out);
  PCAB_CHECK_LAUNCH("pcab_prep_points");
  return PCAB_OK;
}
