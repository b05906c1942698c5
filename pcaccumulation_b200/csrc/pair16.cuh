// Pair-packed activation format ("P16") of the BEV tensors on the tensor-core path.
//
// NHWC like the float32 layout and the same 4 bytes per element, but every 32-channel group of a pixel is stored as
//   [32 x fp16 h | 32 x fp16 l]   (128 bytes),   x ~= h + l,   h = fp16(x),   l = fp16(x - h)
// i.e. each value is already split into the fp16 pair the tcgen05 convolutions multiply (22 significant bits; absolute
// floor 2^-25 once l becomes subnormal, saturation at +-65504).  The producing kernel's epilogue writes it, TMA lands a
// pixel's 128-byte group directly as one swizzled row of the MMA A operand: [h | l] are the two K halves of the row.
// Channel c of a pixel with C channels: group g = c / 32, j = c % 32 -> h at fp16 index g*64 + j, l at g*64 + 32 + j.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace p16 {

// (h, l) halves of two values, packed as f16x2 words (low half = first value); saturating like the tensor-core operand
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& h, uint32_t& l) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(x1 - f.y), "f"(x0 - f.x));
}
__device__ __forceinline__ float2 join2(uint32_t h, uint32_t l) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&l));
  return make_float2(a.x + b.x, a.y + b.y);
}

// byte offset of channel c's h half inside a pixel of a P16 tensor (the l half is 64 bytes further)
__device__ __forceinline__ size_t chan_off(int c) { return (size_t)(c >> 5) * 128 + (size_t)(c & 31) * 2; }

// four consecutive channels c..c+3 (c % 4 == 0) of pixel `pix` of a tensor with C channels
__device__ __forceinline__ float4 load4(const void* t, size_t pix, int C, int c) {
  const char* p = reinterpret_cast<const char*>(t) + pix * (size_t)C * 4 + chan_off(c);
  const uint2 h = *reinterpret_cast<const uint2*>(p);
  const uint2 l = *reinterpret_cast<const uint2*>(p + 64);
  const float2 a = join2(h.x, l.x), b = join2(h.y, l.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store4(void* t, size_t pix, int C, int c, float4 v) {
  char* p = reinterpret_cast<char*>(t) + pix * (size_t)C * 4 + chan_off(c);
  uint2 h, l;
  split2(v.x, v.y, h.x, l.x);
  split2(v.z, v.w, h.y, l.y);
  *reinterpret_cast<uint2*>(p) = h;
  *reinterpret_cast<uint2*>(p + 64) = l;
}
__device__ __forceinline__ float load1(const void* t, size_t pix, int C, int c) {
  const __half* p = reinterpret_cast<const __half*>(reinterpret_cast<const char*>(t) + pix * (size_t)C * 4 + chan_off(c));
  return __half2float(p[0]) + __half2float(p[32]);
}

// generic accessors used by kernels that serve both layouts (fmt: 0 = float32 NHWC, 1 = P16)
template <bool P16>
__device__ __forceinline__ float4 ld4(const float* t, size_t pix, int C, int c) {
  if (P16) return load4(t, pix, C, c);
  return *reinterpret_cast<const float4*>(t + pix * (size_t)C + c);
}
template <bool P16>
__device__ __forceinline__ void st4(float* t, size_t pix, int C, int c, float4 v) {
  if (P16)
    store4(t, pix, C, c, v);
  else
    *reinterpret_cast<float4*>(t + pix * (size_t)C + c) = v;
}

}  // namespace p16
