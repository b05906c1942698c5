// Microbenchmark: FP32 FMA throughput on sm_100a with scalar FFMA vs packed fma.rn.f32x2 (FFMA2).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma_bench tools/ffma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int NACC>
__global__ void k_scalar(float* out, int iters, float a0, float b0) {
  float acc[NACC], w[8];
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int i = 0; i < 8; ++i) w[i] = b0 + i * 1e-3f;
  float x = a0 + threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fmaf(x, w[i & 7], acc[i]);
    x += 1e-9f;
  }
  float s = 0;
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>  // NACC packed accumulators = 2*NACC floats
__global__ void k_packed(float* out, int iters, float a0, float b0) {
  unsigned long long acc[NACC], w[8];
  for (int i = 0; i < NACC; ++i) acc[i] = pack2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  for (int i = 0; i < 8; ++i) w[i] = pack2(b0 + i * 1e-3f, b0 - i * 1e-3f);
  float xs = a0 + threadIdx.x * 1e-6f;
  for (int it = 0; it < iters; ++it) {
    unsigned long long x = pack2(xs, xs);
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma2(x, w[i & 7], acc[i]);
    xs += 1e-9f;
  }
  float s = 0;
  for (int i = 0; i < NACC; ++i) {
    float a, b;
    unpack2(acc[i], a, b);
    s += a + b;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000, grid = 148 * 4, block = 256;
  {
    float ms = time_ms([&] { k_scalar<32><<<grid, block>>>(out, iters, 1.0001f, 0.5f); });
    double fl = 2.0 * 32 * iters * (double)grid * block;
    printf("scalar FFMA  (32 acc/thread, 8 warps x 4 CTAs/SM): %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
  }
  {
    float ms = time_ms([&] { k_packed<16><<<grid, block>>>(out, iters, 1.0001f, 0.5f); });
    double fl = 2.0 * 32 * iters * (double)grid * block;
    printf("packed FFMA2 (16x2 acc/thread)                    : %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
  }
  {
    float ms = time_ms([&] { k_packed<32><<<grid, block>>>(out, iters, 1.0001f, 0.5f); });
    double fl = 2.0 * 64 * iters * (double)grid * block;
    printf("packed FFMA2 (32x2 acc/thread)                    : %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
