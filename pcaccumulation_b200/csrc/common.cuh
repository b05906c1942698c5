// Shared helpers for the pcab200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define PCAB_OK 0
#define PCAB_ERR_CUDA -1
#define PCAB_ERR_ARG -2
#define PCAB_ERR_WORKSPACE -3

void pcab_set_error(const char* fmt, ...);

#define PCAB_CHECK_LAUNCH(name)                                                  \
  do {                                                                           \
    cudaError_t e_ = cudaGetLastError();                                         \
    if (e_ != cudaSuccess) {                                                     \
      pcab_set_error("%s: %s", name, cudaGetErrorString(e_));                    \
      return PCAB_ERR_CUDA;                                                      \
    }                                                                            \
  } while (0)

#define PCAB_CUDA(call)                                                          \
  do {                                                                           \
    cudaError_t e_ = (call);                                                     \
    if (e_ != cudaSuccess) {                                                     \
      pcab_set_error("%s:%d %s", __FILE__, __LINE__, cudaGetErrorString(e_));    \
      return PCAB_ERR_CUDA;                                                      \
    }                                                                            \
  } while (0)

#define PCAB_REQUIRE(cond, msg)                                                  \
  do {                                                                           \
    if (!(cond)) {                                                               \
      pcab_set_error("%s: requirement failed: %s", __func__, msg);               \
      return PCAB_ERR_ARG;                                                       \
    }                                                                            \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// SM count of the CURRENT device (148 on B200), cached per device ordinal; thread-safe (racing writers store the same value)
int pcab_sm_count();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: remember per device ordinal what
// has been set instead of a process-wide flag.  Racing threads at worst set the same value twice.
struct PcabSmemOnce {
  int done[64] = {};
};
template <typename K>
static inline cudaError_t pcab_set_max_smem(K kernel, int bytes, PcabSmemOnce& once) {
  int d = 0;
  cudaError_t e = cudaGetDevice(&d);
  if (e != cudaSuccess) return e;
  const bool tracked = d >= 0 && d < 64;
  if (tracked && __atomic_load_n(&once.done[d], __ATOMIC_ACQUIRE) >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && tracked) __atomic_store_n(&once.done[d], bytes, __ATOMIC_RELEASE);
  return e;
}

// grid sized as a multiple of the SM count for grid-stride kernels
static inline int grid_for(long long n, int block, int per_sm = 8) {
  long long want = (n + block - 1) / block;
  long long cap = (long long)pcab_sm_count() * per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit load that does not pollute L1
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
