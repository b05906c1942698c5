// BEV-map utilities: 2-class head conv + argmax, label scatter-back, pose warp, point transform,
// occupancy / label / mean canvases.
//
// Replaces models/motionnet.py:45-114 (warp_feats), :117-135 (transform_points), :167-170 (canvases),
// :188-194 (FB decision + inverse scatter), models/pillar_encoder.py:177-204 and the second conv of
// SegHead2D (models/unet.py:268) for the 2-class semseg head.
#include "common.cuh"
#include "pair16.cuh"
#include "pcab200.h"

namespace {

// conv3x3 Cin -> 2 (pad 1), NHWC input (float32 or P16), planar NCHW logits out [n][2][H][W] + argmax (first max wins).
// CTA = 8 x 32 output pixels: the 10 x 34 halo tile is staged channel-planar in shared memory with coalesced 16-byte
// loads (a thread-per-pixel gather would fetch nine whole 128-byte pixels per thread), then one thread per pixel runs
// the 9 x CIN x 2 FMAs in the reference's (ky, kx, c) order from conflict-free shared-memory columns.
constexpr int H2_TH = 8, H2_TW = 32, H2_P = 345;  // P: halo-tile pixel pitch of a channel plane, == 1 (mod 8): conflict-free transposing stores
template <int CIN, bool P16>
__global__ void __launch_bounds__(256) k_head2(const float* __restrict__ in, const float* __restrict__ w /* [9][CIN][2] */,
                                               const float* __restrict__ bias, int N, int H, int W, int CS /* channels of `in` */,
                                               float* __restrict__ logits, int* __restrict__ argmax) {
  extern __shared__ __align__(16) float sm_h2[];
  float* tile = sm_h2;                 // [CIN][H2_P]
  float* sw = sm_h2 + CIN * H2_P;      // [9][CIN][2]
  for (int i = threadIdx.x; i < 9 * CIN * 2; i += blockDim.x) sw[i] = w[i];
  const int tiles_x = (W + H2_TW - 1) / H2_TW, tiles_y = (H + H2_TH - 1) / H2_TH;
  const int tx0 = (blockIdx.x % tiles_x) * H2_TW, ty0 = ((blockIdx.x / tiles_x) % tiles_y) * H2_TH, n = blockIdx.x / (tiles_x * tiles_y);
  constexpr int HW_ = H2_TW + 2, HH_ = H2_TH + 2, C4 = CIN / 4;
  for (int e = threadIdx.x; e < HH_ * HW_ * C4; e += blockDim.x) {
    const int q = e % C4, hp = e / C4, hy = hp / HW_, hx = hp % HW_;
    const int gy = ty0 + hy - 1, gx = tx0 + hx - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = p16::ld4<P16>(in, ((size_t)n * H + gy) * W + gx, CS, 4 * q);
    float* d = tile + (4 * q) * H2_P + hp;
    d[0] = v.x, d[H2_P] = v.y, d[2 * H2_P] = v.z, d[3 * H2_P] = v.w;
  }
  __syncthreads();
  const int ty = threadIdx.x / H2_TW, tx = threadIdx.x % H2_TW;
  const int y = ty0 + ty, x = tx0 + tx;
  if (y >= H || x >= W) return;
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    if (y + ky - 1 < 0 || y + ky - 1 >= H) continue;  // (zero padding: skipped like the taps outside the image before)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      if (x + kx - 1 < 0 || x + kx - 1 >= W) continue;
      const float* t = tile + (ty + ky) * HW_ + tx + kx;
      const float* ww = sw + (ky * 3 + kx) * CIN * 2;
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        const float v = t[c * H2_P];
        a0 = fmaf(v, ww[2 * c], a0), a1 = fmaf(v, ww[2 * c + 1], a1);
      }
    }
  }
  a0 += bias[0], a1 += bias[1];
  const size_t hw = (size_t)H * W;
  logits[(size_t)n * 2 * hw + (size_t)y * W + x] = a0;
  logits[(size_t)n * 2 * hw + hw + (size_t)y * W + x] = a1;
  argmax[(size_t)n * hw + (size_t)y * W + x] = a1 > a0 ? 1 : 0;
}

__global__ void k_fb_per_point(const int* __restrict__ fb_map, const int* __restrict__ pillar_cell,
                               const int* __restrict__ p2v, int n, long long* __restrict__ fb_pp) {
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) fb_pp[i] = fb_map[pillar_cell[p2v[i]]];
}

// occupancy f32, GT label i64 and pillar-mean f32 canvases ([B*T,H,W] planar; mean is [B*T,3,H,W])
__global__ void k_canvases(const int* __restrict__ pillar_cell, const int* __restrict__ fb_sub,
                           const float* __restrict__ pmean, int m, int hw, float* __restrict__ occ,
                           long long* __restrict__ fb_map, float* __restrict__ mean_map) {
  int stride = gridDim.x * blockDim.x;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < m; p += stride) {
    int cell = pillar_cell[p];
    occ[cell] = 1.f;
    fb_map[cell] = fb_sub[p];
    int f = cell / hw, r = cell % hw;
    mean_map[((size_t)f * 3 + 0) * hw + r] = pmean[3 * p];
    mean_map[((size_t)f * 3 + 1) * hw + r] = pmean[3 * p + 1];
    mean_map[((size_t)f * 3 + 2) * hw + r] = pmean[3 * p + 2];
  }
}

// rows 0,1 of inv(M) for a 4x4 in double (torch.linalg.inv), out = {i00,i01,i03,i10,i11,i13}
__device__ void inv4_rows01(const float* P, float* out6) {
  double a[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) a[i][j] = P[4 * i + j], a[i][4 + j] = (i == j);
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    for (int r = c + 1; r < 4; ++r)
      if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
    for (int j = 0; j < 8; ++j) {
      double tmp = a[c][j];
      a[c][j] = a[piv][j], a[piv][j] = tmp;
    }
    double d = a[c][c];
    for (int j = 0; j < 8; ++j) a[c][j] /= d;
    for (int r = 0; r < 4; ++r)
      if (r != c) {
        double fct = a[r][c];
        for (int j = 0; j < 8; ++j) a[r][j] -= fct * a[c][j];
      }
  }
  out6[0] = (float)a[0][4], out6[1] = (float)a[0][5], out6[2] = (float)a[0][7];
  out6[3] = (float)a[1][4], out6[4] = (float)a[1][5], out6[5] = (float)a[1][7];
}

constexpr int kMaxWarpFrames = 256;

// bilinear warp of frames 1..T-1 by inv(pose) (zeros padding, align_corners=False); slot 0 = frame T-1 (quirk Q1).
// Input and output share the activation format (float32 NHWC or P16).
template <bool P16>
__global__ void __launch_bounds__(256) k_warp(const float* __restrict__ bev, const float* __restrict__ pose, int B,
                                              int T, int H, int W, int C4, float vx, float vy, float x_min, float y_min,
                                              float* __restrict__ out) {
  __shared__ float s_inv[kMaxWarpFrames][6];
  for (int f = threadIdx.x; f < B * T; f += blockDim.x) inv4_rows01(pose + (size_t)f * 16, s_inv[f]);
  __syncthreads();
  long long total = (long long)B * T * H * W;
  const int C = 4 * C4;
  int lane_c = threadIdx.x % C4;  // threads of a pixel cover its channels (four each)
  int pix_per_block = blockDim.x / C4;
  // One pixel: the four taps are fetched unconditionally from clamped coordinates with the weight of an out-of-image tap set
  // to zero (fma(s, 0, r) == r for the finite s that is read instead), so all loads of a pixel -- and of the second pixel
  // handled in the same iteration -- are in flight together; the accumulation order (and the result bits) is unchanged.
  auto one = [&](long long pix) {
    int x = (int)(pix % W);
    int y = (int)((pix / W) % H);
    int f = (int)(pix / ((long long)W * H));
    int b = f / T, t = f % T;
    if (t == 0) return p16::ld4<P16>(bev, ((size_t)(b * T + T - 1) * H + y) * W + x, C, 4 * lane_c);
    const float* iv = s_inv[f];
    float gx = ((float)x + 0.5f) * vx + x_min;
    float gy = ((float)y + 0.5f) * vy + y_min;
    float u = (iv[0] * gx + iv[1] * gy + iv[2]) / fabsf(x_min);
    float v = (iv[3] * gx + iv[4] * gy + iv[5]) / fabsf(y_min);
    float ix = ((u + 1.f) * W - 1.f) / 2.f;
    float iy = ((v + 1.f) * H - 1.f) / 2.f;
    float fx0 = floorf(ix), fy0 = floorf(iy);
    float wx0 = (fx0 + 1.f) - ix, wx1 = ix - fx0, wy0 = (fy0 + 1.f) - iy, wy1 = iy - fy0;
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (fx0 >= -1.f && fx0 <= (float)W && fy0 >= -1.f && fy0 <= (float)H) {
      const int x0 = (int)fx0, y0 = (int)fy0;
      const size_t fbase = (size_t)f * H * W;
      const bool xa = x0 >= 0 && x0 < W, xb = x0 + 1 >= 0 && x0 + 1 < W, ya = y0 >= 0 && y0 < H, yb = y0 + 1 >= 0 && y0 + 1 < H;
      const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x0 + 1, 0), W - 1), cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y0 + 1, 0), H - 1);
      const float4 s00 = p16::ld4<P16>(bev, fbase + (size_t)cy0 * W + cx0, C, 4 * lane_c);
      const float4 s01 = p16::ld4<P16>(bev, fbase + (size_t)cy0 * W + cx1, C, 4 * lane_c);
      const float4 s10 = p16::ld4<P16>(bev, fbase + (size_t)cy1 * W + cx0, C, 4 * lane_c);
      const float4 s11 = p16::ld4<P16>(bev, fbase + (size_t)cy1 * W + cx1, C, 4 * lane_c);
      auto acc = [&](const float4& s, float wgt, bool ok) {
        if (ok) r.x = fmaf(s.x, wgt, r.x), r.y = fmaf(s.y, wgt, r.y), r.z = fmaf(s.z, wgt, r.z), r.w = fmaf(s.w, wgt, r.w);
      };
      acc(s00, wx0 * wy0, xa && ya);
      acc(s01, wx1 * wy0, xb && ya);
      acc(s10, wx0 * wy1, xa && yb);
      acc(s11, wx1 * wy1, xb && yb);
    }
    return r;
  };
  const long long step = (long long)gridDim.x * pix_per_block;
  for (long long pix = (long long)blockIdx.x * pix_per_block + threadIdx.x / C4; pix < total; pix += 2 * step) {
    const long long pix2 = pix + step;
    const float4 r0 = one(pix);
    float4 r1 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pix2 < total) r1 = one(pix2);
    p16::st4<P16>(out, (size_t)pix, C, 4 * lane_c, r0);
    if (pix2 < total) p16::st4<P16>(out, (size_t)pix2, C, 4 * lane_c, r1);
  }
}

__global__ void k_transform_points(const float* __restrict__ xyz, const int* __restrict__ pframe,
                                   const float* __restrict__ pose, int n, float* __restrict__ out) {
  int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float* P = pose + (size_t)pframe[i] * 16;
    float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      // (R @ p^T) + t : dot product of three terms, then the translation
      float d = P[4 * r] * x;
      d = fmaf(P[4 * r + 1], y, d);
      d = fmaf(P[4 * r + 2], z, d);
      out[3 * i + r] = d + P[4 * r + 3];
    }
  }
}

}  // namespace

extern "C" int pcab_head2_conv(const float* in_nhwc, int cin, int in_cstride, int fmt, const float* weight_packed,
                               const float* bias, int n_images, int H, int W, float* logits_nchw, int* argmax_map,
                               cudaStream_t stream) {
  PCAB_REQUIRE(cin == 32 && in_cstride >= cin && in_cstride % 32 == 0, "cin == 32, read from the first 32 of in_cstride channels");
  const int blocks = n_images * cdiv(H, H2_TH) * cdiv(W, H2_TW);
  const size_t smem = (size_t)(32 * H2_P + 9 * 32 * 2) * sizeof(float);
  static PcabSmemOnce once0, once1;
  if (fmt) {
    PCAB_CUDA(pcab_set_max_smem(k_head2<32, true>, (int)smem, once1));
    k_head2<32, true><<<blocks, 256, smem, stream>>>(in_nhwc, weight_packed, bias, n_images, H, W, in_cstride, logits_nchw, argmax_map);
  } else {
    PCAB_CUDA(pcab_set_max_smem(k_head2<32, false>, (int)smem, once0));
    k_head2<32, false><<<blocks, 256, smem, stream>>>(in_nhwc, weight_packed, bias, n_images, H, W, in_cstride, logits_nchw, argmax_map);
  }
  PCAB_CHECK_LAUNCH("pcab_head2_conv");
  return PCAB_OK;
}

extern "C" int pcab_fb_per_point(const int* fb_map, const int* pillar_cell, const int* p2v, int n_points,
                                 long long* fb_per_point, cudaStream_t stream) {
  k_fb_per_point<<<grid_for(n_points, 256), 256, 0, stream>>>(fb_map, pillar_cell, p2v, n_points, fb_per_point);
  PCAB_CHECK_LAUNCH("pcab_fb_per_point");
  return PCAB_OK;
}

extern "C" int pcab_canvases(const int* pillar_cell, const int* fb_sub, const float* pillar_mean, int n_pillars, int H,
                             int W, float* occ_map, long long* fb_map, float* mean_map, cudaStream_t stream) {
  k_canvases<<<grid_for(n_pillars, 256), 256, 0, stream>>>(pillar_cell, fb_sub, pillar_mean, n_pillars, H * W, occ_map,
                                                          fb_map, mean_map);
  PCAB_CHECK_LAUNCH("pcab_canvases");
  return PCAB_OK;
}

extern "C" int pcab_warp_bev(const float* bev_nhwc, const float* pose, int B, int T, int H, int W, int C, float vx,
                             float vy, float x_min, float y_min, float* out_nhwc, int fmt, cudaStream_t stream) {
  PCAB_REQUIRE(C % 4 == 0 && 256 % (C / 4) == 0 && (!fmt || C % 32 == 0), "C/4 must divide 256 (P16: C%32)");
  PCAB_REQUIRE(B * T <= kMaxWarpFrames, "too many frames per call");
  long long total = (long long)B * T * H * W;
  int pix_per_block = 256 / (C / 4);
  if (fmt)
    k_warp<true><<<grid_for(total, pix_per_block, 16), 256, 0, stream>>>(bev_nhwc, pose, B, T, H, W, C / 4, vx, vy, x_min, y_min, out_nhwc);
  else
    k_warp<false><<<grid_for(total, pix_per_block, 16), 256, 0, stream>>>(bev_nhwc, pose, B, T, H, W, C / 4, vx, vy, x_min, y_min, out_nhwc);
  PCAB_CHECK_LAUNCH("pcab_warp_bev");
  return PCAB_OK;
}

extern "C" int pcab_transform_points(const float* xyz, const int* point_frame, const float* pose, int n_points,
                                     float* out, cudaStream_t stream) {
  k_transform_points<<<grid_for(n_points, 256), 256, 0, stream>>>(xyz, point_frame, pose, n_points, out);
  PCAB_CHECK_LAUNCH("pcab_transform_points");
  return PCAB_OK;
}
