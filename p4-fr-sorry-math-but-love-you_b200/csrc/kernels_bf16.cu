// kernels_bf16.cu -- bandwidth-bound kernels of the bf16 trunk (NHWC bf16
// activations, fp32 arithmetic): stem conv, depthwise 3x3, squeeze-excite.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"

#ifndef FRX_DW_ACC32
#define FRX_DW_ACC32 2
#endif

namespace frx {

namespace {
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = eh2_unpack(w[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = eh2_pack(f[2 * i], f[2 * i + 1]);
  return make_uint4(w[0], w[1], w[2], w[3]);
}
}  // namespace

// Stem: conv3x3 s2 p0 + BN + SiLU (EfficientSATRN.py:67-73,:82-83); NCHW fp32 in, NHWC bf16 out.
// One thread per output pixel x 8 channels (16-byte store).
__global__ void __launch_bounds__(256) stem_conv_bf16_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             eh_t* __restrict__ out, int B, int Cin, int H, int W,
                                                             int OH, int OW, int Cout) {
  extern __shared__ float ws[];
  for (int i = threadIdx.x; i < Cout * Cin * 9; i += blockDim.x) ws[i] = w[i];
  float* ssc = ws + Cout * Cin * 9;
  float* ssh = ssc + Cout;
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) { ssc[i] = scale[i]; ssh[i] = shift[i]; }
  __syncthreads();
  const int C8 = Cout / 8;
  long long total = (long long)B * OH * OW * C8;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co0 = (int)(idx % C8) * 8;
  long long pix = idx / C8;
  const int ow = (int)(pix % OW), oh = (int)((pix / OW) % OH), n = (int)(pix / ((long long)OW * OH));
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int ci = 0; ci < Cin; ++ci) {
    const float* ip = in + (((long long)n * Cin + ci) * H + oh * 2) * W + ow * 2;
    float x[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) x[kh * 3 + kw] = __ldg(ip + kh * W + kw);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* wp = ws + ((co0 + j) * Cin + ci) * 9;
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[j] = fmaf(x[t], wp[t], acc[j]);
    }
  }
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = act_apply(acc[j] * ssc[co0 + j] + ssh[co0 + j], ACT_SILU);
  *reinterpret_cast<uint4*>(out + pix * Cout + co0) = pack8(o);
}

// Stem, one thread per output PIXEL (all COUT channels): the per-(pixel, 8 channels) form above re-reads the 9 inputs
// three times and spends one shared load per FMA on the weights (130 us against 20 us of HBM time at B = 256).  Here
// a tap's COUT weights are contiguous in shared memory and read as broadcast 16-byte loads (one per 4 FMAs); the
// accumulation order per output (ci outer, tap inner) is the same as above, so the results are bit-identical.
template <int COUT>
__global__ void __launch_bounds__(128) stem_conv_px_bf16_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                                eh_t* __restrict__ out, int B, int Cin, int H, int W,
                                                                int OH, int OW) {
  extern __shared__ __align__(16) float ws[];   // [Cin * 9][COUT], then scale[COUT], shift[COUT]
  const int taps = Cin * 9;
  for (int i = threadIdx.x; i < COUT * taps; i += blockDim.x) {
    const int co = i / taps, r = i - co * taps;
    ws[r * COUT + co] = w[i];
  }
  float* ssc = ws + COUT * taps;
  float* ssh = ssc + COUT;
  uint4* stage = reinterpret_cast<uint4*>(ws + ((COUT * taps + 2 * COUT + 3) & ~3));   // [128][COUT / 8] x 16 B
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) { ssc[i] = scale[i]; ssh[i] = shift[i]; }
  __syncthreads();
  const long long npix = (long long)B * OH * OW, pix0 = (long long)blockIdx.x * blockDim.x;
  const long long pix = min(pix0 + threadIdx.x, npix - 1);   // the tail threads recompute the last pixel (not stored)
  const int ow = (int)(pix % OW), oh = (int)((pix / OW) % OH), n = (int)(pix / ((long long)OW * OH));
  float acc[COUT];
#pragma unroll
  for (int j = 0; j < COUT; ++j) acc[j] = 0.f;
  for (int ci = 0; ci < Cin; ++ci) {
    const float* ip = in + (((long long)n * Cin + ci) * H + oh * 2) * W + ow * 2;
    float x[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) x[kh * 3 + kw] = __ldg(ip + kh * W + kw);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4* wp = reinterpret_cast<const float4*>(ws + (ci * 9 + t) * COUT);
#pragma unroll
      for (int j = 0; j < COUT / 4; ++j) {
        const float4 wv = wp[j];
        acc[4 * j] = fmaf(x[t], wv.x, acc[4 * j]);
        acc[4 * j + 1] = fmaf(x[t], wv.y, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(x[t], wv.z, acc[4 * j + 2]);
        acc[4 * j + 3] = fmaf(x[t], wv.w, acc[4 * j + 3]);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < COUT / 8; ++g) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = act_apply(acc[g * 8 + j] * ssc[g * 8 + j] + ssh[g * 8 + j], ACT_SILU);
    stage[threadIdx.x * (COUT / 8) + g] = pack8(o);
  }
  // the CTA's 128 pixels are one contiguous run of the NHWC output: store it as full lines
  __syncthreads();
  const long long cta_vecs = min((long long)blockDim.x, npix - pix0) * (COUT / 8);
  uint4* op = reinterpret_cast<uint4*>(out + pix0 * COUT);
  for (int i = threadIdx.x; i < cta_vecs; i += blockDim.x) op[i] = stage[i];
}

void launch_stem_conv_bf16(const float* in, const float* w, const float* scale, const float* shift,
                           eh_t* out, int B, int Cin, int H, int W, int OH, int OW, int Cout, cudaStream_t st) {
  if (Cout == 24) {
    const long long pixels = (long long)B * OH * OW;
    const int smem24 = ((24 * Cin * 9 + 2 * 24 + 3) & ~3) * sizeof(float) + 128 * 3 * 16;
    stem_conv_px_bf16_kernel<24><<<(unsigned)((pixels + 127) / 128), 128, smem24, st>>>(in, w, scale, shift, out, B, Cin, H, W, OH, OW);
    return;
  }
  long long total = (long long)B * OH * OW * (Cout / 8);
  int smem = (Cout * Cin * 9 + 2 * Cout) * sizeof(float);
  stem_conv_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, smem, st>>>(in, w, scale, shift, out, B, Cin, H, W, OH, OW, Cout);
}

// Depthwise 3x3 + folded BN (+bias) + activation; 8 channels (16 bytes) per thread.
__global__ void __launch_bounds__(256) dwconv3x3_bf16_kernel(const eh_t* __restrict__ in, const float* __restrict__ w,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             eh_t* __restrict__ out, int B, int H, int W, int C, int OH,
                                                             int OW, int stride, int pad_t, int pad_l, int act) {
  const int C8 = C >> 3;
  long long total = (long long)B * OH * OW * C8;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C8) * 8;
  long long pix = idx / C8;
  const int ow = (int)(pix % OW), oh = (int)((pix / OW) % OH), n = (int)(pix / ((long long)OW * OH));
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int ih = oh * stride - pad_t + kh;
    if (ih < 0 || ih >= H) continue;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int iw = ow * stride - pad_l + kw;
      if (iw < 0 || iw >= W) continue;
      float x[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(in + (((long long)n * H + ih) * W + iw) * C + c)), x);
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (kh * 3 + kw) * C + c));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (kh * 3 + kw) * C + c + 4));
      acc[0] = fmaf(x[0], w0.x, acc[0]); acc[1] = fmaf(x[1], w0.y, acc[1]);
      acc[2] = fmaf(x[2], w0.z, acc[2]); acc[3] = fmaf(x[3], w0.w, acc[3]);
      acc[4] = fmaf(x[4], w1.x, acc[4]); acc[5] = fmaf(x[5], w1.y, acc[5]);
      acc[6] = fmaf(x[6], w1.z, acc[6]); acc[7] = fmaf(x[7], w1.w, acc[7]);
    }
  }
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = act_apply(acc[j] * __ldg(scale + c + j) + __ldg(shift + c + j), act);
  *reinterpret_cast<uint4*>(out + pix * C + c) = pack8(o);
}

void launch_dwconv_bf16(const eh_t* in, const float* w, const float* scale, const float* shift,
                        eh_t* out, int B, int H, int W, int C, int OH, int OW, int stride, int pad_t,
                        int pad_l, int act, cudaStream_t st) {
  long long total = (long long)B * OH * OW * (C / 8);
  dwconv3x3_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, w, scale, shift, out, B, H, W, C, OH, OW,
                                                                        stride, pad_t, pad_l, act);
}

// Squeeze-excite, one CTA per image: gate = sigmoid(W2 silu(W1 mean_hw(x) + b1) + b2), then x *= gate
// in place (so the following 1x1 projection is a plain tcgen05 GEMM).  Deterministic (no atomics).
__global__ void __launch_bounds__(256) se_scale_bf16_kernel(eh_t* __restrict__ x, const float* __restrict__ w1,
                                                            const float* __restrict__ b1, const float* __restrict__ w2,
                                                            const float* __restrict__ b2, int HW, int C, int R) {
  extern __shared__ float sm[];
  float* mean = sm;       // C
  float* red = sm + C;    // R
  float* gate = red + R;  // C
  const int n = blockIdx.x;
  eh_t* xp = x + (long long)n * HW * C;
  const float inv = 1.f / (float)HW;
  for (int c8 = threadIdx.x; c8 < C / 8; c8 += blockDim.x) {
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    for (int q = 0; q < HW; ++q) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(xp + (long long)q * C + c8 * 8), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) mean[c8 * 8 + j] = s[j] * inv;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < R; r += 8) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(__ldg(w1 + (long long)r * C + c), mean[c], s);
    s = warp_sum(s);
    if (lane == 0) red[r] = act_apply(s + __ldg(b1 + r), ACT_SILU);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = __ldg(b2 + c);
    const float* wr = w2 + (long long)c * R;
    for (int r = 0; r < R; ++r) s = fmaf(__ldg(wr + r), red[r], s);
    gate[c] = 1.f / (1.f + expf(-s));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < HW * (C / 8); i += blockDim.x) {
    const int c8 = i % (C / 8);
    uint4* ptr = reinterpret_cast<uint4*>(xp + (long long)(i / (C / 8)) * C + c8 * 8);
    float v[8];
    unpack8(*ptr, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= gate[c8 * 8 + j];
    *ptr = pack8(v);
  }
}

void launch_se_scale_bf16(eh_t* x, const float* w1, const float* b1, const float* w2, const float* b2,
                          int B, int HW, int C, int R, cudaStream_t st) {
  se_scale_bf16_kernel<<<B, 256, (2 * C + R) * sizeof(float), st>>>(x, w1, b1, w2, b2, HW, C, R);
}


// ---------------------------------------------------------------------------
// MBConv middle section, restructured for parallelism (the one-CTA-per-image SE
// kernel above was latency-bound):
//   1. dwconv_se_mean: depthwise 3x3 + BN + SiLU for one (image, 64-channel chunk)
//      per CTA, which also owns that chunk's spatial mean -> deterministic, no atomics;
//   2. se_fc: the two tiny FCs per image -> gate[B][C];
//   3. se_apply: x *= gate, fully parallel 16-byte elementwise pass.
// ---------------------------------------------------------------------------
// One CTA = one (image, 64-channel chunk).  The whole input tile of the chunk is staged in shared memory with
// coalesced 16-byte loads issued up front (all in flight at once), then every output pixel reads its 9 taps
// from shared memory; the CTA also owns the chunk's spatial mean (deterministic, no atomics).
__global__ void __launch_bounds__(256) dwconv_se_mean_kernel(const eh_t* __restrict__ in, const float* __restrict__ w,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             eh_t* __restrict__ out, float* __restrict__ mean,
                                                             int H, int W, int C, int OH, int OW, int stride, int pad_t,
                                                             int pad_l, const float* __restrict__ se_w1,
                                                             const float* __restrict__ se_w2, int se_floats) {
  extern __shared__ __align__(16) unsigned char dw_smem[];
  {  // the SE FC weights are read once per forward pass: pull them into L2 now so that se_fc does not wait on DRAM
    const long long line = ((long long)(blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x) * 32;  // 128 B
    if (line < se_floats) {
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(se_w1 + line));
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(se_w2 + line));
    }
  }
  uint4* tile = reinterpret_cast<uint4*>(dw_smem);                                   // [H*W][8 channel groups] x 16 B
  float (*red)[64 + 1] = reinterpret_cast<float (*)[64 + 1]>(dw_smem + (size_t)H * W * 8 * 16);  // [32][65]
  float* wsm = reinterpret_cast<float*>(dw_smem + (size_t)H * W * 8 * 16 + 32 * 65 * sizeof(float));  // [11][64]: 9 taps, scale, shift
  const int n = blockIdx.x, c0 = blockIdx.y * 64;
  const int cg = threadIdx.x & 7, pl = threadIdx.x >> 3;  // 8 channels per thread, 32 pixel lanes
  const int c = c0 + cg * 8;
  const bool c_ok = c < C;
  const eh_t* ip = in + (long long)n * H * W * C;
  for (int i = threadIdx.x; i < H * W * 8; i += 256) {
    const int px = i >> 3, g = i & 7;
    tile[i] = (c0 + g * 8 < C) ? __ldg(reinterpret_cast<const uint4*>(ip + (long long)px * C + c0 + g * 8)) : make_uint4(0u, 0u, 0u, 0u);
  }
  float sum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sum[j] = 0.f;
  for (int i = threadIdx.x; i < 11 * 64; i += 256) {
    const int t = i >> 6, cc = c0 + (i & 63);
    const float* src = t < 9 ? w + t * C : (t == 9 ? scale : shift);
    wsm[i] = cc < C ? __ldg(src + cc) : 0.f;
  }
  __syncthreads();
  auto ld8s = [&](int t, float (&dst)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(wsm + t * 64 + cg * 8);
    const float4 b4 = *reinterpret_cast<const float4*>(wsm + t * 64 + cg * 8 + 4);
    dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w; dst[4] = b4.x; dst[5] = b4.y; dst[6] = b4.z; dst[7] = b4.w;
  };
  float sc[8], sh[8];
  ld8s(9, sc);
  ld8s(10, sh);
  eh_t* op = out + (long long)n * OH * OW * C;
  if (c_ok) {
    for (int px = pl; px < OH * OW; px += 32) {
      const int oh = px / OW, ow = px - oh * OW;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = oh * stride - pad_t + kh;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int iw = ow * stride - pad_l + kw;
          if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
            float x[8], wv[8];
            unpack8(tile[(ih * W + iw) * 8 + cg], x);
            ld8s(kh * 3 + kw, wv);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(x[j], wv[j], acc[j]);
          }
        }
      }
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = fmaf(acc[j], sc[j], sh[j]);
        o[j] = __fdividef(v, 1.f + __expf(-v));  // SiLU
      }
      const uint4 packed = pack8(o);
      *reinterpret_cast<uint4*>(op + (long long)px * C + c) = packed;
      float back[8];
      unpack8(packed, back);  // the mean is taken over the stored (bf16-rounded) activations
#pragma unroll
      for (int j = 0; j < 8; ++j) sum[j] += back[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[pl][cg * 8 + j] = sum[j];
  __syncthreads();
  if (threadIdx.x < 64 && c0 + threadIdx.x < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += red[i][threadIdx.x];
    mean[(long long)n * C + c0 + threadIdx.x] = s / (float)(OH * OW);
  }
}

// Default MBConv depthwise kernel: 4 channels per thread, 32-channel chunks per CTA (8 channel groups x 16 pixel lanes =
// 128 threads, 8 KB tile).  The generic kernel above is bound by shared-memory load wavefronts (78 % of the L1/TEX
// peak, 40 % of them bank-conflict replays of the per-pixel tap re-reads) and its 256-thread CTAs keep too few tile
// loads in flight; here
//   * the 9 taps of the thread's channels live in registers, a tap costs one 8-byte shared load;
//   * the small CTAs keep 6-8 CTAs per SM resident, so the tile loads of some overlap the arithmetic of others;
//   * the 9-tap sum, the folded BN and the SiLU run on PACKED HALF PAIRS: the tile is converted bf16 -> fp16 once while
//     it is staged, SiLU(v) = v (0.5 tanh(v/2) + 0.5) costs one tanh.approx.f16x2 per pair.  fp16 carries 11 significand
//     bits through the 9-term sum -- more than the 8 bits the bf16 store keeps; activations and folded-BN outputs are
//     O(1..100), far from the fp16 range limit.
template <int CH>
__global__ void __launch_bounds__(CH * 4) dwconv4_se_mean_kernel(const eh_t* __restrict__ in, const float* __restrict__ w,
                                                                 const float* __restrict__ scale, const float* __restrict__ shift,
                                                                 eh_t* __restrict__ out, float* __restrict__ mean,
                                                                 int H, int W, int C, int OH, int OW, int stride, int pad_t,
                                                                 int pad_l, const float* __restrict__ se_w1,
                                                                 const float* __restrict__ se_w2, int se_floats) {
  constexpr int NT = CH * 4, G8 = CH / 8, G4 = CH / 4;
  extern __shared__ __align__(16) unsigned char dw_smem[];
  {  // SE FC weights -> L2 (see dwconv_se_mean_kernel)
    const long long line = ((long long)(blockIdx.y * gridDim.x + blockIdx.x) * NT + threadIdx.x) * 32;
    if (line < se_floats) {
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(se_w1 + line));
      asm volatile("prefetch.global.L2 [%0];\n" ::"l"(se_w2 + line));
    }
  }
  uint4* tile = reinterpret_cast<uint4*>(dw_smem);                                               // [H*W][CH/8] x 16 B
  float (*red)[CH + 1] = reinterpret_cast<float (*)[CH + 1]>(dw_smem + (size_t)H * W * G8 * 16);  // [16][CH + 1]
  const int n = blockIdx.x, c0 = blockIdx.y * CH;
  const eh_t* ip = in + (long long)n * H * W * C;
  for (int i = threadIdx.x; i < H * W * G8; i += NT) {
    const int px = i / G8, g = i - px * G8;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (c0 + g * 8 < C) {
#if FRX_ENC_FP16
      v = __ldg(reinterpret_cast<const uint4*>(ip + (long long)px * C + c0 + g * 8));   // already packed halves
#else
      float x[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(ip + (long long)px * C + c0 + g * 8)), x);
      const __half2 h0 = __floats2half2_rn(x[0], x[1]), h1 = __floats2half2_rn(x[2], x[3]);
      const __half2 h2 = __floats2half2_rn(x[4], x[5]), h3 = __floats2half2_rn(x[6], x[7]);
      v = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                     *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
#endif
    }
    tile[i] = v;
  }
  const int cg = threadIdx.x % G4, pl = threadIdx.x / G4;  // 4 channels per thread, 16 pixel lanes
  const int c = c0 + cg * 4;
  const bool c_ok = c < C;
#if FRX_DW_ACC32
  // fp32 taps / accumulation / folded BN / SiLU (operands stay packed halves in shared memory): the 9-term sum and the
  // activation then add no rounding of their own to the fp16-stored trunk (FRX_DW_ACC32=0: packed-half arithmetic)
  float4 wt[9], sc, sh;
#if FRX_DW_ACC32 == 2
  __half2 wh[9][2];
#endif
  {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[t] = c_ok ? __ldg(reinterpret_cast<const float4*>(w + t * C + c)) : z;
#if FRX_DW_ACC32 == 2
#pragma unroll
    for (int t = 0; t < 9; ++t) { wh[t][0] = __floats2half2_rn(wt[t].x, wt[t].y); wh[t][1] = __floats2half2_rn(wt[t].z, wt[t].w); }
#endif
    sc = c_ok ? __ldg(reinterpret_cast<const float4*>(scale + c)) : z;
    sh = c_ok ? __ldg(reinterpret_cast<const float4*>(shift + c)) : z;
  }
  __syncthreads();
  const uint2* tile2 = reinterpret_cast<const uint2*>(tile);  // [H*W][CH/4] x 8 B
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
  eh_t* op = out + (long long)n * OH * OW * C;
  if (c_ok) {
    for (int px = pl; px < OH * OW; px += 16) {
      const int oh = px / OW, ow = px - oh * OW;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = oh * stride - pad_t + kh;
#if FRX_DW_ACC32 == 2
        // one filter row (3 taps) in packed halves, rows combined in fp32: a third of the conversions of the all-fp32 form
        __half2 r0 = __floats2half2_rn(0.f, 0.f), r1 = r0;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int iw = ow * stride - pad_l + kw;
          if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
            const uint2 u = tile2[(ih * W + iw) * G4 + cg];
            r0 = __hfma2(*reinterpret_cast<const __half2*>(&u.x), wh[kh * 3 + kw][0], r0);
            r1 = __hfma2(*reinterpret_cast<const __half2*>(&u.y), wh[kh * 3 + kw][1], r1);
          }
        }
        const float2 a = __half22float2(r0), b = __half22float2(r1);
        acc.x += a.x; acc.y += a.y; acc.z += b.x; acc.w += b.y;
#else
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int iw = ow * stride - pad_l + kw;
          if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
            const uint2 u = tile2[(ih * W + iw) * G4 + cg];
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
            const float4 ww = wt[kh * 3 + kw];
            acc.x = fmaf(a.x, ww.x, acc.x); acc.y = fmaf(a.y, ww.y, acc.y); acc.z = fmaf(b.x, ww.z, acc.z); acc.w = fmaf(b.y, ww.w, acc.w);
          }
        }
#endif
      }
      float2 f0 = make_float2(fmaf(acc.x, sc.x, sh.x), fmaf(acc.y, sc.y, sh.y)), f1 = make_float2(fmaf(acc.z, sc.z, sh.z), fmaf(acc.w, sc.w, sh.w));
      f0.x = __fdividef(f0.x, 1.f + __expf(-f0.x)); f0.y = __fdividef(f0.y, 1.f + __expf(-f0.y));
      f1.x = __fdividef(f1.x, 1.f + __expf(-f1.x)); f1.y = __fdividef(f1.y, 1.f + __expf(-f1.y));
#else
  __half2 wt[9][2], sc[2], sh[2];
  {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 f = c_ok ? __ldg(reinterpret_cast<const float4*>(w + t * C + c)) : z;
      wt[t][0] = __floats2half2_rn(f.x, f.y); wt[t][1] = __floats2half2_rn(f.z, f.w);
    }
    const float4 a = c_ok ? __ldg(reinterpret_cast<const float4*>(scale + c)) : z;
    const float4 d = c_ok ? __ldg(reinterpret_cast<const float4*>(shift + c)) : z;
    sc[0] = __floats2half2_rn(a.x, a.y); sc[1] = __floats2half2_rn(a.z, a.w);
    sh[0] = __floats2half2_rn(d.x, d.y); sh[1] = __floats2half2_rn(d.z, d.w);
  }
  const __half2 half_ = __floats2half2_rn(0.5f, 0.5f);
  __syncthreads();
  const uint2* tile2 = reinterpret_cast<const uint2*>(tile);  // [H*W][CH/4] x 8 B
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
  eh_t* op = out + (long long)n * OH * OW * C;
  if (c_ok) {
    for (int px = pl; px < OH * OW; px += 16) {
      const int oh = px / OW, ow = px - oh * OW;
      __half2 acc0 = __floats2half2_rn(0.f, 0.f), acc1 = acc0;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = oh * stride - pad_t + kh;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int iw = ow * stride - pad_l + kw;
          if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
            const uint2 u = tile2[(ih * W + iw) * G4 + cg];
            acc0 = __hfma2(*reinterpret_cast<const __half2*>(&u.x), wt[kh * 3 + kw][0], acc0);
            acc1 = __hfma2(*reinterpret_cast<const __half2*>(&u.y), wt[kh * 3 + kw][1], acc1);
          }
        }
      }
      const __half2 v0 = __hfma2(acc0, sc[0], sh[0]), v1 = __hfma2(acc1, sc[1], sh[1]);       // folded BN
      const __half2 o0 = __hmul2(v0, __hfma2(h2tanh_approx(__hmul2(v0, half_)), half_, half_));  // SiLU = v (0.5 tanh(v/2) + 0.5)
      const __half2 o1 = __hmul2(v1, __hfma2(h2tanh_approx(__hmul2(v1, half_)), half_, half_));
      const float2 f0 = __half22float2(o0), f1 = __half22float2(o1);
#endif
      uint2 packed;
      packed.x = eh2_pack(f0.x, f0.y);
      packed.y = eh2_pack(f1.x, f1.y);
      *reinterpret_cast<uint2*>(op + (long long)px * C + c) = packed;
      // the mean is taken over the stored (rounded) activations
      const float2 b0 = eh2_unpack(packed.x), b1 = eh2_unpack(packed.y);
      sum[0] += b0.x; sum[1] += b0.y; sum[2] += b1.x; sum[3] += b1.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[pl][cg * 4 + j] = sum[j];
  __syncthreads();
  if (threadIdx.x < CH && c0 + threadIdx.x < C) {
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s2 += red[i][threadIdx.x];
    mean[(long long)n * C + c0 + threadIdx.x] = s2 / (float)(OH * OW);
  }
}

constexpr int SE_IMGS = 4;  // images per CTA
constexpr int SE_ROWS = 4;  // FC1 rows per warp: each 16-byte weight load meets 4 images, each 16-byte mean load 4 rows
// One CTA = 4 images.  FC1 (C -> R, SiLU): warp w owns rows [4w, 4w+4), lanes stride over C in float4 steps, so the
// shared-memory traffic for the means (the previous version's limit) is amortised over 4 rows; FC2 (R -> C,
// sigmoid): one thread per output channel, outputs split over gridDim.y CTAs.
__global__ void __launch_bounds__(512, 1) se_fc_kernel(const float* __restrict__ mean, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, float* __restrict__ gate, int B, int C,
                                                       int R) {
  extern __shared__ float sm[];
  float* mv = sm;                  // [SE_IMGS][C]
  float* red = sm + SE_IMGS * C;   // [SE_IMGS][R]
  const int n0 = blockIdx.x * SE_IMGS;
  for (int i = threadIdx.x; i < SE_IMGS * C; i += blockDim.x) {
    const int im = i / C;
    mv[i] = n0 + im < B ? mean[(long long)n0 * C + i] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  for (int r0 = warp * SE_ROWS; r0 < R; r0 += nwarps * SE_ROWS) {
    float s[SE_ROWS][SE_IMGS];
#pragma unroll
    for (int j = 0; j < SE_ROWS; ++j)
#pragma unroll
      for (int im = 0; im < SE_IMGS; ++im) s[j][im] = 0.f;
#pragma unroll 3
    for (int c4 = lane; c4 < C / 4; c4 += 32) {  // C % 4 == 0
      float4 wv[SE_ROWS], m4[SE_IMGS];
#pragma unroll
      for (int j = 0; j < SE_ROWS; ++j)
        wv[j] = r0 + j < R ? __ldg(reinterpret_cast<const float4*>(w1 + (long long)(r0 + j) * C) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int im = 0; im < SE_IMGS; ++im) m4[im] = *reinterpret_cast<const float4*>(mv + im * C + c4 * 4);
#pragma unroll
      for (int j = 0; j < SE_ROWS; ++j)
#pragma unroll
        for (int im = 0; im < SE_IMGS; ++im)
          s[j][im] = fmaf(wv[j].x, m4[im].x, fmaf(wv[j].y, m4[im].y, fmaf(wv[j].z, m4[im].z, fmaf(wv[j].w, m4[im].w, s[j][im]))));
    }
#pragma unroll
    for (int j = 0; j < SE_ROWS; ++j)
#pragma unroll
      for (int im = 0; im < SE_IMGS; ++im) {
        const float t = warp_sum(s[j][im]);
        if (lane == 0 && r0 + j < R) red[im * R + r0 + j] = act_apply(t + __ldg(b1 + r0 + j), ACT_SILU);
      }
  }
  __syncthreads();
  for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < C; c += gridDim.y * blockDim.x) {  // expand outputs split over gridDim.y
    float s[SE_IMGS];
    const float bb = __ldg(b2 + c);
#pragma unroll
    for (int im = 0; im < SE_IMGS; ++im) s[im] = bb;
#pragma unroll 32
    for (int r = 0; r < R; ++r) {
      const float wv = __ldg(w2 + (long long)r * C + c);  // w2 is [R][C] (transposed at pack time)
#pragma unroll
      for (int im = 0; im < SE_IMGS; ++im) s[im] = fmaf(wv, red[im * R + r], s[im]);
    }
#pragma unroll
    for (int im = 0; im < SE_IMGS; ++im)
      if (n0 + im < B) gate[(long long)(n0 + im) * C + c] = 1.f / (1.f + expf(-s[im]));
  }
}

__global__ void __launch_bounds__(256) se_apply_kernel(eh_t* __restrict__ x, const float* __restrict__ gate,
                                                       long long total8, int HW, int C) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int C8 = C >> 3;
  const int c = (int)(i % C8) * 8;
  const long long n = i / ((long long)C8 * HW);
  uint4* ptr = reinterpret_cast<uint4*>(x) + i;
  float v[8];
  unpack8(*ptr, v);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + n * C + c));
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + n * C + c + 4));
  v[0] *= g0.x; v[1] *= g0.y; v[2] *= g0.z; v[3] *= g0.w;
  v[4] *= g1.x; v[5] *= g1.y; v[6] *= g1.z; v[7] *= g1.w;
  *ptr = pack8(v);
}

void launch_mbconv_dw_se_bf16(const eh_t* in, const float* w, const float* scale, const float* shift,
                              eh_t* out, float* mean, float* gate, const float* w1, const float* b1,
                              const float* w2, const float* b2, int B, int H, int W, int C, int OH, int OW, int stride,
                              int pad_t, int pad_l, int R, cudaStream_t st) {
  dim3 g(B, (C + 63) / 64);
  const size_t smem = (size_t)H * W * 8 * 16 + 32 * 65 * sizeof(float) + 11 * 64 * sizeof(float);
  static SmemOptIn opt;
  opt.ensure(dwconv_se_mean_kernel, smem);
  if (C % 8 == 0) {  // 4 channels per thread, 32-channel chunks: small CTAs, taps in registers, packed-half arithmetic
    constexpr int CH = 32;
    const size_t sm4 = (size_t)H * W * (CH / 8) * 16 + 16 * (CH + 1) * sizeof(float);
    static SmemOptIn opt4;
    opt4.ensure(dwconv4_se_mean_kernel<CH>, sm4);
    dwconv4_se_mean_kernel<CH><<<dim3(B, (C + CH - 1) / CH), CH * 4, sm4, st>>>(in, w, scale, shift, out, mean, H, W, C, OH, OW,
                                                                             stride, pad_t, pad_l, w1, w2, R * C);
  } else {
    dwconv_se_mean_kernel<<<g, 256, smem, st>>>(in, w, scale, shift, out, mean, H, W, C, OH, OW, stride, pad_t, pad_l, w1, w2, R * C);
  }
  se_fc_kernel<<<dim3((B + SE_IMGS - 1) / SE_IMGS, 2), 512, SE_IMGS * (C + R) * sizeof(float), st>>>(mean, w1, b1, w2, b2, gate, B, C, R);
  long long total8 = (long long)B * OH * OW * (C / 8);
  se_apply_kernel<<<(unsigned)((total8 + 255) / 256), 256, 0, st>>>(out, gate, total8, OH * OW, C);
}

// ---------------------------------------------------------------------------
// Encoder self-attention for the bf16 path (EfficientSATRN.py:164-172,:198-228): S = 32 tokens, head dim 64,
// scores divided by sqrt(d_model).  One warp = 16 query rows of one (image, head) on mma.sync.m16n8k16:
// S = Q K^T (4 key tiles x 4 k-steps), softmax in the accumulator layout, O = P V (8 dim tiles x 2 k-steps) with
// the probabilities re-used as the A fragment.  Operands are converted fp32 -> bf16 while the fragments are loaded
// (the qkv rows are the fp32 output of the projection GEMM); 64 HMMA per warp instead of ~16k scalar instructions.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  return eh2_pack(lo, hi);
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32." FRX_EH_PTX "." FRX_EH_PTX ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int S>   // tokens per image: 32 (EfficientSATRN, 4 x 8 map) or 128 (LiteSATRN, 8 x 16 map); one warp per 16 query rows
__global__ void __launch_bounds__(S * 2) enc_attn_mma_kernel(const float* __restrict__ qkv,       // [B*S, 3*D]
                                                             eh_t* __restrict__ out,              // [B*S, D]
                                                             int D, int heads, float inv_temp) {
  constexpr int HD = 64, NT = S / 8;
  const int b = blockIdx.x / heads, hh = blockIdx.x % heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  const int ld = 3 * D;
  const float* qb = qkv + (long long)b * S * ld + hh * HD;  // q of token 0; k at +D, v at +2D
  const int i0 = warp * 16;
  float sacc[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) sacc[nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const float* q0 = qb + (long long)(i0 + gid) * ld + 16 * ks + 2 * tig;
    const float* q1 = q0 + 8 * ld;
    const float2 a00 = __ldg(reinterpret_cast<const float2*>(q0)), a10 = __ldg(reinterpret_cast<const float2*>(q1));
    const float2 a01 = __ldg(reinterpret_cast<const float2*>(q0 + 8)), a11 = __ldg(reinterpret_cast<const float2*>(q1 + 8));
    const uint32_t a[4] = {pack2(a00.x, a00.y), pack2(a10.x, a10.y), pack2(a01.x, a01.y), pack2(a11.x, a11.y)};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float* kp = qb + D + (long long)(8 * nt + gid) * ld + 16 * ks + 2 * tig;
      const float2 k0 = __ldg(reinterpret_cast<const float2*>(kp)), k1 = __ldg(reinterpret_cast<const float2*>(kp + 8));
      mma16816(sacc[nt], a, pack2(k0.x, k0.y), pack2(k1.x, k1.y));
    }
  }
  // softmax over the S keys of rows gid (c0, c1) and gid + 8 (c2, c3)
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) sacc[nt][e] *= inv_temp;
    mx0 = fmaxf(mx0, fmaxf(sacc[nt][0], sacc[nt][1]));
    mx1 = fmaxf(mx1, fmaxf(sacc[nt][2], sacc[nt][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    sacc[nt][0] = __expf(sacc[nt][0] - mx0); sacc[nt][1] = __expf(sacc[nt][1] - mx0);
    sacc[nt][2] = __expf(sacc[nt][2] - mx1); sacc[nt][3] = __expf(sacc[nt][3] - mx1);
    sum0 += sacc[nt][0] + sacc[nt][1];
    sum1 += sacc[nt][2] + sacc[nt][3];
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  float oacc[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[n][e] = 0.f;
#pragma unroll
  for (int kk = 0; kk < S / 16; ++kk) {  // 16 keys per k-step = score tiles 2kk (k lo) and 2kk + 1 (k hi)
    const uint32_t a[4] = {pack2(sacc[2 * kk][0], sacc[2 * kk][1]), pack2(sacc[2 * kk][2], sacc[2 * kk][3]),
                           pack2(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]), pack2(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3])};
    const float* vp = qb + 2 * D + (long long)(16 * kk + 2 * tig) * ld + gid;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const float v00 = __ldg(vp + 8 * n), v01 = __ldg(vp + ld + 8 * n);
      const float v10 = __ldg(vp + 8 * ld + 8 * n), v11 = __ldg(vp + 9 * ld + 8 * n);
      mma16816(oacc[n], a, pack2(v00, v01), pack2(v10, v11));
    }
  }
  const float r0 = __fdividef(1.f, sum0), r1 = __fdividef(1.f, sum1);
  eh_t* o0 = out + ((long long)b * S + i0 + gid) * D + hh * HD + 2 * tig;
  eh_t* o1 = o0 + (long long)8 * D;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    *reinterpret_cast<uint32_t*>(o0 + 8 * n) = pack2(oacc[n][0] * r0, oacc[n][1] * r0);
    *reinterpret_cast<uint32_t*>(o1 + 8 * n) = pack2(oacc[n][2] * r1, oacc[n][3] * r1);
  }
}

// returns false when the shape is not one the kernel is compiled for (the caller then uses the generic kernel)
bool launch_enc_attn_mma_bf16(const float* qkv, eh_t* out, int B, int S, int D, int heads, cudaStream_t st) {
  if (D / heads != 64 || D % heads != 0) return false;
  const float inv_temp = 1.f / sqrtf((float)D);
  if (S == 32) enc_attn_mma_kernel<32><<<B * heads, 64, 0, st>>>(qkv, out, D, heads, inv_temp);
  else if (S == 128) enc_attn_mma_kernel<128><<<B * heads, 256, 0, st>>>(qkv, out, D, heads, inv_temp);
  else return false;
  return true;
}

// ---------------------------------------------------------------------------
// Narrow 3x3 convolution (the two 24 -> 24 blocks of trunk stage 0, 63 x 127 maps): out = silu(bn(conv(x))) (+ x).
// With 24 channels a per-tap im2col feed re-reads every input pixel 9x from L2 in 48-byte pieces and pads K
// from 216 to 576 and N from 24 to 32; this kernel stages an 8 x 32 output tile's halo ((8+2) x (32+2) pixels x
// 24 channels, 16 KB) in shared memory ONCE, and runs the implicit GEMM [256 pixels x 216] x [216 x 24] on
// mma.sync.m16n8k16 with A fragments gathered straight from the halo tile (k = tap * 24 + channel; an 8-aligned
// k group never straddles a tap, so every fragment register is one 4-byte shared load).  Warp w owns tile row w.
// Weights: B fragments pre-packed on the host, [14 k-steps][3 n-tiles][32 lanes] x {b0, b1}, staged in smem.
// ---------------------------------------------------------------------------
constexpr int C24 = 24, C24_TH = 8, C24_TW = 32, C24_KS = 14;  // K = 216 padded to 224

__global__ void __launch_bounds__(256) conv3x3_c24_mma_kernel(const eh_t* __restrict__ in, const uint2* __restrict__ wfrag,
                                                              const float* __restrict__ scale, const float* __restrict__ shift,
                                                              eh_t* __restrict__ out, int H, int W, int add_res) {
  constexpr int PW = C24_TW + 2, PH = C24_TH + 2;
  __shared__ __align__(16) eh_t halo[PH * PW * C24];   // 16320 B
  __shared__ uint2 wsm[C24_KS * 3 * 32];                        // 10752 B
  const int tiles_x = (W + C24_TW - 1) / C24_TW, tiles_y = (H + C24_TH - 1) / C24_TH;
  const int n = blockIdx.x / (tiles_x * tiles_y), t = blockIdx.x % (tiles_x * tiles_y);
  const int y0 = (t / tiles_x) * C24_TH, x0 = (t % tiles_x) * C24_TW;
  const eh_t* ip = in + (long long)n * H * W * C24;
  // halo: 3 x 16-byte chunks per pixel; a halo row is contiguous in global memory (NHWC, 48 B per pixel)
  for (int i = threadIdx.x; i < PH * PW * 3; i += 256) {
    const int pix = i / 3, ch = i - pix * 3;
    const int py = pix / PW, px = pix - py * PW;
    const int y = y0 + py - 1, x = x0 + px - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(reinterpret_cast<const uint4*>(ip + ((long long)y * W + x) * C24) + ch);
    reinterpret_cast<uint4*>(halo)[i] = v;
  }
  for (int i = threadIdx.x; i < C24_KS * 3 * 32; i += 256) wsm[i] = __ldg(wfrag + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  float acc[2][3][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
  const uint32_t* hw = reinterpret_cast<const uint32_t*>(halo);  // 12 words per pixel
#pragma unroll
  for (int s = 0; s < C24_KS; ++s) {
    const int klo = 16 * s, khi = 16 * s + 8;
    const int tap_lo = klo / C24, ch_lo = klo % C24, tap_hi = khi / C24, ch_hi = khi % C24;  // compile-time after unrolling
    uint32_t a[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int px = mt * 16 + gid;  // tile column of fragment rows gid (a0, a2); +8 for a1, a3
      const int p_lo = ((warp + tap_lo / 3) * PW + px + tap_lo % 3) * 12 + (ch_lo >> 1) + tig;
      a[mt][0] = hw[p_lo];
      a[mt][1] = hw[p_lo + 8 * 12];
      if (tap_hi < 9) {
        const int p_hi = ((warp + tap_hi / 3) * PW + px + tap_hi % 3) * 12 + (ch_hi >> 1) + tig;
        a[mt][2] = hw[p_hi];
        a[mt][3] = hw[p_hi + 8 * 12];
      } else {
        a[mt][2] = 0u; a[mt][3] = 0u;
      }
    }
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      const uint2 b = wsm[(s * 3 + nt) * 32 + lane];
      mma16816(acc[0][nt], a[0], b.x, b.y);
      mma16816(acc[1][nt], a[1], b.x, b.y);
    }
  }
  // epilogue: folded BN, SiLU, residual (the centre tap of the halo), bf16 pairs
  const int y = y0 + warp;
  if (y >= H) return;
  eh_t* op = out + ((long long)n * H + y) * W * C24;
#pragma unroll
  for (int nt = 0; nt < 3; ++nt) {
    const int ch = nt * 8 + 2 * tig;
    const float s0 = __ldg(scale + ch), s1 = __ldg(scale + ch + 1), h0 = __ldg(shift + ch), h1 = __ldg(shift + ch + 1);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        const int px = mt * 16 + gid + hr * 8, x = x0 + px;
        if (x >= W) continue;
        float v0 = fmaf(acc[mt][nt][2 * hr], s0, h0), v1 = fmaf(acc[mt][nt][2 * hr + 1], s1, h1);
        v0 = __fdividef(v0, 1.f + __expf(-v0));
        v1 = __fdividef(v1, 1.f + __expf(-v1));
        if (add_res) {
          const uint32_t r = hw[((warp + 1) * PW + px + 1) * 12 + (ch >> 1)];
          const float2 rf = eh2_unpack(r);
          v0 += rf.x;
          v1 += rf.y;
        }
        *reinterpret_cast<uint32_t*>(op + (long long)x * C24 + ch) = pack2(v0, v1);
      }
  }
}

void launch_conv3x3_c24_bf16(const eh_t* in, const void* wfrag, const float* scale, const float* shift,
                             eh_t* out, int B, int H, int W, int add_res, cudaStream_t st) {
  const int tiles = ((W + C24_TW - 1) / C24_TW) * ((H + C24_TH - 1) / C24_TH);
  conv3x3_c24_mma_kernel<<<B * tiles, 256, 0, st>>>(in, reinterpret_cast<const uint2*>(wfrag), scale, shift, out, H, W, add_res);
}

// ---------------------------------------------------------------------------
// LiteSATRN ShallowCNN in bf16 mode (networks/LiteSATRN.py:21-70).
// Layer 0 fused: conv3x3 p1 (Cin input channels, NCHW fp32) + folded BN + ReLU + maxpool 2x2 -> NHWC bf16 at half
// resolution.  The full-resolution activation (128 channels x 128 x 256 per image) never exists in memory.
// One thread = one pooled pixel x 8 output channels; the 4x4 input patch per input channel is read once.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lite_conv0_pool_bf16_kernel(const float* __restrict__ in, const float* __restrict__ w,  // [O][I][3][3]
                                                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                                                   eh_t* __restrict__ out, int B, int Cin, int H, int W, int Cout) {
  extern __shared__ float ws0[];
  for (int i = threadIdx.x; i < Cout * Cin * 9; i += blockDim.x) ws0[i] = w[i];
  float* ssc = ws0 + Cout * Cin * 9;
  float* ssh = ssc + Cout;
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) { ssc[i] = scale[i]; ssh[i] = shift[i]; }
  __syncthreads();
  const int OH = H / 2, OW = W / 2, C8 = Cout / 8;
  const long long total = (long long)B * OH * OW * C8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co0 = (int)(idx % C8) * 8;
  const long long pix = idx / C8;
  const int ow = (int)(pix % OW), oh = (int)((pix / OW) % OH), n = (int)(pix / ((long long)OW * OH));
  float acc[4][8];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
  for (int ci = 0; ci < Cin; ++ci) {
    float x[4][4];  // input rows 2oh-1 .. 2oh+2, columns 2ow-1 .. 2ow+2 (zero padding 1)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int ih = 2 * oh - 1 + a, iw = 2 * ow - 1 + b;
        x[a][b] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(in + (((long long)n * Cin + ci) * H + ih) * W + iw) : 0.f;
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* wp = ws0 + ((co0 + j) * Cin + ci) * 9;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float wv = wp[kh * 3 + kw];
          acc[0][j] = fmaf(x[kh][kw], wv, acc[0][j]);
          acc[1][j] = fmaf(x[kh][kw + 1], wv, acc[1][j]);
          acc[2][j] = fmaf(x[kh + 1][kw], wv, acc[2][j]);
          acc[3][j] = fmaf(x[kh + 1][kw + 1], wv, acc[3][j]);
        }
    }
  }
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float s = ssc[co0 + j], t = ssh[co0 + j];
    const float m = fmaxf(fmaxf(fmaf(acc[0][j], s, t), fmaf(acc[1][j], s, t)), fmaxf(fmaf(acc[2][j], s, t), fmaf(acc[3][j], s, t)));
    o[j] = fmaxf(m, 0.f);  // relu(max(.)) == max(relu(.))
  }
  *reinterpret_cast<uint4*>(out + pix * Cout + co0) = pack8(o);
}

// The same layer on mma.sync (Cout = 64 or 128): one warp = 4 pooled pixels = 16 conv pixels x Cout channels, K = 9 Cin padded to 16 KS.
// MMA row r <-> (position r / 4 of the 2 x 2 pool window, pooled pixel r % 4), so a thread's two rows (gid, gid + 8) are two
// window positions of ONE pooled pixel and lane ^ 16 holds the other two: max-pool = one fmax + one shuffle.  The patches come
// straight from the fp32 image (L1-resident: 3 loads per thread and k-step), weights as B fragments in registers.
template <int KS, int Cout>
__global__ void __launch_bounds__(256) lite_conv0_pool_mma_kernel(const float* __restrict__ in, const float* __restrict__ w,  // [Cout][Cin][3][3]
                                                                  const float* __restrict__ scale, const float* __restrict__ shift,
                                                                  eh_t* __restrict__ out, int B, int Cin, int H, int W) {
  constexpr int NT = Cout / 8;
  const int lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  const int OH = H / 2, OW = W / 2, K = 9 * Cin;
  const long long tiles = (long long)B * OH * (OW / 4);
  const long long tile = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tile >= tiles) return;
  uint32_t bf[KS][NT][2];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int k = 16 * ks + 8 * h2 + 2 * tig, n = 8 * nt + gid;
        const float lo = k < K ? __ldg(w + (long long)n * K + k) : 0.f, hi = k + 1 < K ? __ldg(w + (long long)n * K + k + 1) : 0.f;
        bf[ks][nt][h2] = pack2(lo, hi);
      }
  const int owq = (int)(tile % (OW / 4));
  const long long t2 = tile / (OW / 4);
  const int oh = (int)(t2 % OH), n = (int)(t2 / OH);
  // this thread's rows: gid (position gid / 4, pooled pixel gid % 4) and gid + 8 (position gid / 4 + 2, same pooled pixel)
  const int pp = gid & 3, pos0 = gid >> 2;
  const int ow = owq * 4 + pp;
  const float* img = in + (long long)n * Cin * H * W;
  auto patch = [&](int pos, int k) -> float {   // input value of conv pixel `pos` of the pool window, im2col column k
    if (k >= K) return 0.f;
    const int ci = k / 9, tap = k - ci * 9, kh = tap / 3, kw = tap - kh * 3;
    const int ih = 2 * oh + (pos >> 1) + kh - 1, iw = 2 * ow + (pos & 1) + kw - 1;
    return (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(img + ((long long)ci * H + ih) * W + iw) : 0.f;
  };
  float acc[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int k0 = 16 * ks + 2 * tig;
    const uint32_t a[4] = {pack2(patch(pos0, k0), patch(pos0, k0 + 1)), pack2(patch(pos0 + 2, k0), patch(pos0 + 2, k0 + 1)),
                           pack2(patch(pos0, k0 + 8), patch(pos0, k0 + 9)), pack2(patch(pos0 + 2, k0 + 8), patch(pos0 + 2, k0 + 9))};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) mma16816(acc[nt], a, bf[ks][nt][0], bf[ks][nt][1]);
  }
  eh_t* op = out + (((long long)n * OH + oh) * OW + ow) * Cout;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int ch = 8 * nt + 2 * tig;
    const float s0 = __ldg(scale + ch), s1 = __ldg(scale + ch + 1), t0 = __ldg(shift + ch), t1 = __ldg(shift + ch + 1);
    float m0 = fmaxf(fmaf(acc[nt][0], s0, t0), fmaf(acc[nt][2], s0, t0));   // rows gid and gid + 8: window positions pos0, pos0 + 2
    float m1 = fmaxf(fmaf(acc[nt][1], s1, t1), fmaf(acc[nt][3], s1, t1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 16));                   // lane ^ 16 = gid ^ 4: the other two positions
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 16));
    if (gid < 4) *reinterpret_cast<uint32_t*>(op + ch) = pack2(fmaxf(m0, 0.f), fmaxf(m1, 0.f));   // relu(max(.)) == max(relu(.))
  }
}

void launch_lite_conv0_pool_bf16(const float* in, const float* w, const float* scale, const float* shift, eh_t* out,
                                 int B, int Cin, int H, int W, int Cout, cudaStream_t st) {
  if ((Cout == 64 || Cout == 128) && Cin <= 3 && (W / 2) % 4 == 0 && H % 2 == 0) {
    const long long tiles = (long long)B * (H / 2) * (W / 2 / 4);
    const unsigned grid = (unsigned)((tiles + 7) / 8);
    if (Cout == 64) {
      if (Cin == 1) lite_conv0_pool_mma_kernel<1, 64><<<grid, 256, 0, st>>>(in, w, scale, shift, out, B, Cin, H, W);
      else lite_conv0_pool_mma_kernel<2, 64><<<grid, 256, 0, st>>>(in, w, scale, shift, out, B, Cin, H, W);
    } else {
      if (Cin == 1) lite_conv0_pool_mma_kernel<1, 128><<<grid, 256, 0, st>>>(in, w, scale, shift, out, B, Cin, H, W);
      else lite_conv0_pool_mma_kernel<2, 128><<<grid, 256, 0, st>>>(in, w, scale, shift, out, B, Cin, H, W);
    }
    return;
  }
  const long long total = (long long)B * (H / 2) * (W / 2) * (Cout / 8);
  const int smem = (Cout * Cin * 9 + 2 * Cout) * sizeof(float);
  lite_conv0_pool_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, smem, st>>>(in, w, scale, shift, out, B, Cin, H, W, Cout);
}

// Max-pool 2x2 stride 2, NHWC bf16 (8 channels per thread); out_f32 != nullptr writes fp32 instead (the last pool feeds the
// fp32 positional-encoding / encoder-layer kernels).
__global__ void __launch_bounds__(256) maxpool2_bf16_kernel(const eh_t* __restrict__ in, eh_t* __restrict__ out,
                                                            float* __restrict__ out_f32, int B, int H, int W, int C) {
  const int OH = H / 2, OW = W / 2, C8 = C / 8;
  const long long total = (long long)B * OH * OW * C8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % C8) * 8;
  const long long pix = idx / C8;
  const int ow = (int)(pix % OW), oh = (int)((pix / OW) % OH), n = (int)(pix / ((long long)OW * OH));
  const eh_t* base = in + (((long long)n * H + oh * 2) * W + ow * 2) * C + c;
  float a[8], b[8], d[8], e[8], o[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(base)), a);
  unpack8(__ldg(reinterpret_cast<const uint4*>(base + C)), b);
  unpack8(__ldg(reinterpret_cast<const uint4*>(base + (long long)W * C)), d);
  unpack8(__ldg(reinterpret_cast<const uint4*>(base + (long long)W * C + C)), e);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaxf(a[j], b[j]), fmaxf(d[j], e[j]));
  if (out_f32) {
    float4* op = reinterpret_cast<float4*>(out_f32 + pix * C + c);
    op[0] = make_float4(o[0], o[1], o[2], o[3]);
    op[1] = make_float4(o[4], o[5], o[6], o[7]);
  } else {
    *reinterpret_cast<uint4*>(out + pix * C + c) = pack8(o);
  }
}

void launch_maxpool2_bf16(const eh_t* in, eh_t* out, float* out_f32, int B, int H, int W, int C, cudaStream_t st) {
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / 8);
  maxpool2_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, out_f32, B, H, W, C);
}

}  // namespace frx
