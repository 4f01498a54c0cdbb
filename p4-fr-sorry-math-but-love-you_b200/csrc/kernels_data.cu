// kernels_data.cu -- the image input pipeline on the device.
//
// Replaces, for a batch of decoded uint8 images already in HBM, what LoadDataset.__getitem__ (data/dataset.py:62-83) and
// get_valid_transforms / get_test_transforms (data/augmentations.py:27-46) do per image on the CPU:
//   90-degree rotation of tall images (h / w > 2, dataset.py:77-78) -> A.Resize(height, width) = cv2.resize INTER_LINEAR
//   on the uint8 image -> A.Normalize(mean, std) -> ToTensorV2 (HWC -> CHW), collated to [B, C, height, width] fp32.
// The resize reproduces OpenCV's 8-bit fixed-point arithmetic bit for bit (11-bit coefficients, its two-stage shifted
// vertical pass, and the 2x2 box average it switches to when both scale factors are exactly 2), so the tensor is
// identical to the reference's, not merely close.  One thread per output pixel; HBM-bound: every output pixel reads
// 4 source pixels that neighbouring threads share through L1.
#include <cstdarg>
#include <cstdio>
#include <vector>

#include "../../include/frx.h"
#include "common.cuh"
#include "runtime.h"

namespace frx {

struct ImgMeta {
  long long offset;        // byte offset of the image in the packed buffer (HWC, tightly packed)
  int h, w;                // stored size
  int rotate;              // 1: the pipeline rotates this image by 90 degrees (counter-clockwise) first
  int area2;               // 1: both scale factors are exactly 2 -> OpenCV's INTER_AREA fast path
  double scale_x, scale_y; // source / destination size ratios of the (rotated) image, as OpenCV computes them
};

// pixel (y, x) of the image AFTER the optional rotation: rot90(img, 1)[y][x] = img[x][w - 1 - y]
__device__ __forceinline__ int src_px(const unsigned char* p, const ImgMeta& m, int y, int x, int c, int C) {
  const long long idx = m.rotate ? ((long long)x * m.w + (m.w - 1 - y)) : ((long long)y * m.w + x);
  return p[idx * C + c];
}

__device__ __forceinline__ void lin_coef(int d, double scale, int n, int clamp_frac, int& s, int& c0, int& c1) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int i = (int)floorf(f);
  f -= (float)i;
  if (clamp_frac) {            // horizontal pass: resize.cpp zeroes the fraction at the borders
    if (i < 0) { f = 0.f; i = 0; }
    if (i >= n - 1) { f = 0.f; i = n - 1; }
  }
  s = i;
  c0 = __float2int_rn((1.f - f) * 2048.f);
  c1 = __float2int_rn(f * 2048.f);
}

template <int C>
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const unsigned char* __restrict__ packed, const ImgMeta* __restrict__ meta,
                                                            float* __restrict__ out, int out_h, int out_w, float3 mean, float3 inv_std) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= out_h * out_w) return;
  const int dy = pix / out_w, dx = pix - dy * out_w;
  const ImgMeta m = meta[b];
  const unsigned char* p = packed + m.offset;
  const int H = m.rotate ? m.w : m.h, W = m.rotate ? m.h : m.w;   // size after the rotation
  int v[C];
  if (m.area2) {
#pragma unroll
    for (int c = 0; c < C; ++c)
      v[c] = (src_px(p, m, 2 * dy, 2 * dx, c, C) + src_px(p, m, 2 * dy, 2 * dx + 1, c, C) + src_px(p, m, 2 * dy + 1, 2 * dx, c, C) +
              src_px(p, m, 2 * dy + 1, 2 * dx + 1, c, C) + 2) >> 2;
  } else {
    int sx, a0, a1, sy, b0, b1;
    lin_coef(dx, m.scale_x, W, 1, sx, a0, a1);
    lin_coef(dy, m.scale_y, H, 0, sy, b0, b1);
    const int x1 = min(sx + 1, W - 1);
    const int y0 = min(max(sy, 0), H - 1), y1 = min(max(sy + 1, 0), H - 1);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int s0 = src_px(p, m, y0, sx, c, C) * a0 + src_px(p, m, y0, x1, c, C) * a1;
      const int s1 = src_px(p, m, y1, sx, c, C) * a0 + src_px(p, m, y1, x1, c, C) * a1;
      const int r = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
      v[c] = min(max(r, 0), 255);
    }
  }
  const float mu[3] = {mean.x, mean.y, mean.z}, is[3] = {inv_std.x, inv_std.y, inv_std.z};
#pragma unroll
  for (int c = 0; c < C; ++c) out[(((long long)b * C + c) * out_h + dy) * out_w + dx] = ((float)v[c] - mu[c]) * is[c];
}

}  // namespace frx

using namespace frx;

extern "C" int frx_preprocess_u8(const uint8_t* packed, const int64_t* offsets, const int32_t* heights, const int32_t* widths,
                                 int32_t batch, int32_t channels, int32_t out_h, int32_t out_w, const float* mean, const float* stddev,
                                 int32_t rotate_tall, float* out, void* stream) {
  if (!packed || !offsets || !heights || !widths || !out || !mean || !stddev) return 1;
  if (batch <= 0 || (channels != 1 && channels != 3) || out_h <= 0 || out_w <= 0) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<ImgMeta> host(batch);
  for (int i = 0; i < batch; ++i) {
    ImgMeta& m = host[i];
    m.offset = offsets[i]; m.h = heights[i]; m.w = widths[i];
    if (m.h <= 0 || m.w <= 0) return 1;
    m.rotate = rotate_tall && ((double)m.h / (double)m.w > 2.0) ? 1 : 0;      // dataset.py:77  h / w > 2
    const int H = m.rotate ? m.w : m.h, W = m.rotate ? m.h : m.w;
    const double inv_x = (double)out_w / (double)W, inv_y = (double)out_h / (double)H;   // resize.cpp: scale = 1 / inv_scale
    m.scale_x = 1.0 / inv_x; m.scale_y = 1.0 / inv_y;
    m.area2 = (W == 2 * out_w && H == 2 * out_h) ? 1 : 0;
  }
  ImgMeta* dev = nullptr;
  if (cudaMallocAsync((void**)&dev, sizeof(ImgMeta) * batch, st) != cudaSuccess) return 2;
  // the vector dies when this call returns: synchronous copy semantics for pageable memory make that safe
  if (cudaMemcpyAsync(dev, host.data(), sizeof(ImgMeta) * batch, cudaMemcpyHostToDevice, st) != cudaSuccess) { cudaFreeAsync(dev, st); return 2; }
  const float3 mu = channels == 3 ? make_float3(mean[0] * 255.f, mean[1] * 255.f, mean[2] * 255.f) : make_float3(mean[0] * 255.f, 0.f, 0.f);
  const float3 is = channels == 3 ? make_float3(1.f / (stddev[0] * 255.f), 1.f / (stddev[1] * 255.f), 1.f / (stddev[2] * 255.f))
                                  : make_float3(1.f / (stddev[0] * 255.f), 0.f, 0.f);
  const dim3 grid((out_h * out_w + 255) / 256, batch);
  if (channels == 3) preprocess_u8_kernel<3><<<grid, 256, 0, st>>>(packed, dev, out, out_h, out_w, mu, is);
  else preprocess_u8_kernel<1><<<grid, 256, 0, st>>>(packed, dev, out, out_h, out_w, mu, is);
  const cudaError_t e = cudaGetLastError();
  cudaFreeAsync(dev, st);
  return e == cudaSuccess ? 0 : 3;
}
