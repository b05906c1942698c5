// GPU data front-end: scene crop + ground removal as ONE stable stream compaction over the
// raw sample arrays, feeding the voxeliser directly.
//
// Replaces steps 2-4 of libs/dataset.py:163-207 (BaseDataset.prep_input without augmentation: |x|,|y| < crop_xy,
// crop_z_min < z < crop_z_max, then z > ground_height + ground_slack), which the reference runs with six boolean-mask
// gathers per step in the DataLoader workers.  Comparisons are in float32 like numpy's (float32 array vs Python float).
// The kept points leave as (x, y, z, t) float32 rows - the voxeliser's input - with their labels, in the original order.
//
// Training-time augmentation (step 1, libs/dataset.py:90-113,167-171) is applied inside the same pass when requested: random
// rigid transform (float64, toolbox/register_utils.py:199-206), uniform jitter, global scale -- all in float64 like numpy does
// once the float64 transform has touched the float32 points; the crop / ground comparisons then see float64 values and the
// rows are rounded to float32 at the end (the reference's .astype(np.float32) / model-side .float()).  The random numbers are
// the HOST's (numpy global stream, drawn in the reference's order: libs/dataset.py:103-106, 96, 99) when an exact replay of
// the reference's stream is wanted, or a counter-based generator on the device (jitter only) when it is not.
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
  // augmentation (augment = 0: raw float32 points, float32 comparisons)
  int augment;
  double tsfm[12];      // rows of [R | t]
  const double* noise;  // [n,3] uniform [0,1) drawn by the host, or NULL: device generator with `seed`
  unsigned long long seed;
  double noise_amp, scale;
  double crop_xy_d, z_min_d, z_max_d, ground_d;
};

__device__ __forceinline__ double uniform01(unsigned long long seed, unsigned long long idx) {  // splitmix64 of (seed, idx)
  unsigned long long z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// augmented coordinates of point i in float64: ((R p + t) + (u - 0.5) * noise_amp) * scale
__device__ __forceinline__ void augmented(const PrepArgs& a, int i, double& x, double& y, double& z) {
  const double px = a.pts[3 * i], py = a.pts[3 * i + 1], pz = a.pts[3 * i + 2];
  double v[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    // R @ src.T + t: the three products summed left to right, then the translation (no contraction: dgemm's own
    // FMA use is not reproducible either; differences stay at 1e-16 relative and vanish in the float32 rounding)
    const double acc = __dadd_rn(__dadd_rn(__dmul_rn(a.tsfm[4 * r], px), __dmul_rn(a.tsfm[4 * r + 1], py)), __dmul_rn(a.tsfm[4 * r + 2], pz));
    const double u = a.noise ? a.noise[3 * (size_t)i + r] : uniform01(a.seed, 3ull * i + r);
    v[r] = __dmul_rn(__dadd_rn(__dadd_rn(acc, a.tsfm[4 * r + 3]), __dmul_rn(u - 0.5, a.noise_amp)), a.scale);
  }
  x = v[0], y = v[1], z = v[2];
}

__device__ __forceinline__ bool keep(const PrepArgs& a, int i) {
  if (a.augment) {
    double x, y, z;
    augmented(a, i, x, y, z);
    bool k = fabs(x) < a.crop_xy_d && fabs(y) < a.crop_xy_d && z < a.z_max_d && z > a.z_min_d;
    if (a.remove_ground) k = k && z > a.ground_d;
    return k;
  }
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
      if (a.augment) {
        double x, y, z;
        augmented(a, i, x, y, z);
        points4[j] = make_float4((float)x, (float)y, (float)z, (float)t);
      } else {
        points4[j] = make_float4(a.pts[3 * i], a.pts[3 * i + 1], a.pts[3 * i + 2], (float)t);
      }
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
static int prep_core(PrepArgs a, const float* raw_points, const long long* time_idx, const long long* sd_labels,
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
  a.pts = raw_points, a.tidx = time_idx, a.sd = sd_labels, a.fb = fb_labels, a.inst = inst_labels, a.n = n;
  a.crop_xy = crop_xy, a.z_min = crop_z_min, a.z_max = crop_z_max, a.ground = ground_height, a.remove_ground = remove_ground;
  k_prep_flags<<<grid_for(n, 256), 256, 0, stream>>>(a, flag);
  PCAB_CUDA(cub::DeviceScan::ExclusiveSum(w, scan, flag, pos, n, stream));
  k_prep_scatter<<<grid_for(n, 256), 256, 0, stream>>>(a, flag, pos, (float4*)points4_out, time_out, sd_out, fb_out, inst_out,
                                                       count_out);
  PCAB_CHECK_LAUNCH("pcab_prep_points");
  return PCAB_OK;
}

extern "C" int pcab_prep_points(const float* raw_points, const long long* time_idx, const long long* sd_labels,
                                const long long* fb_labels, const long long* inst_labels, int n, float crop_xy, float crop_z_min,
                                float crop_z_max, int remove_ground, float ground_height, float* points4_out, int* time_out,
                                long long* sd_out, long long* fb_out, long long* inst_out, int* count_out, void* workspace,
                                size_t workspace_bytes, cudaStream_t stream) {
  PrepArgs a = {};
  a.augment = 0;
  return prep_core(a, raw_points, time_idx, sd_labels, fb_labels, inst_labels, n, crop_xy, crop_z_min, crop_z_max, remove_ground,
                   ground_height, points4_out, time_out, sd_out, fb_out, inst_out, count_out, workspace, workspace_bytes, stream);
}

// The same with the training-time augmentation of libs/dataset.py:90-113,167-171 applied first.  tsfm16: HOST pointer to the
// random rigid transform [4,4] float64 (row major); noise: DEVICE [n,3] float64 uniforms in [0,1) drawn by the host in the
// reference's order, or NULL for the device generator seeded with `seed`; thresholds as float64 (the comparisons are float64).
extern "C" int pcab_prep_points_augmented(const float* raw_points, const long long* time_idx, const long long* sd_labels,
                                          const long long* fb_labels, const long long* inst_labels, int n, const double* tsfm16,
                                          const double* noise, unsigned long long seed, double noise_amp, double scale,
                                          double crop_xy, double crop_z_min, double crop_z_max, int remove_ground,
                                          double ground_height, float* points4_out, int* time_out, long long* sd_out,
                                          long long* fb_out, long long* inst_out, int* count_out, void* workspace,
                                          size_t workspace_bytes, cudaStream_t stream) {
  PCAB_REQUIRE(tsfm16 != nullptr, "tsfm16 is required");
  PrepArgs a = {};
  a.augment = 1;
  for (int k = 0; k < 12; ++k) a.tsfm[k] = tsfm16[k];
  a.noise = noise, a.seed = seed, a.noise_amp = noise_amp, a.scale = scale;
  a.crop_xy_d = crop_xy, a.z_min_d = crop_z_min, a.z_max_d = crop_z_max, a.ground_d = ground_height;
  return prep_core(a, raw_points, time_idx, sd_labels, fb_labels, inst_labels, n, (float)crop_xy, (float)crop_z_min,
                   (float)crop_z_max, remove_ground, (float)ground_height, points4_out, time_out, sd_out, fb_out, inst_out, count_out,
                   workspace, workspace_bytes, stream);
}
