// kernels_swin.cu -- SwinTRN encoder kernels (networks/SWIN.py), fp32 mode.
// Linear layers go through the fp32 implicit-GEMM kernel; this file holds the pieces that
// are specific to Swin: patch embedding, shifted-window attention, patch merging.
#include "common.cuh"
#include "kernels.h"

namespace frx {

// PatchEmbed (:531-573) + patch LayerNorm + absolute position embedding (:725-729):
// conv 4x4 stride 4 (3 -> E channels) on NCHW fp32, one block (E threads) per token.
__global__ void __launch_bounds__(128) swin_patch_embed_kernel(const float* __restrict__ img, const float* __restrict__ w,  // [E][3][4][4]
                                                               const float* __restrict__ bias, const float* __restrict__ g,
                                                               const float* __restrict__ b, const float* __restrict__ ape,   // [R*R][E]
                                                               float* __restrict__ out, int IMGS, int R, int E) {
  __shared__ float patch[48];
  __shared__ float red[2][4];
  const int tok = blockIdx.x % (R * R), n = blockIdx.x / (R * R);
  const int ph = tok / R, pw = tok % R;
  if (threadIdx.x < 48) {
    const int ci = threadIdx.x / 16, kh = (threadIdx.x / 4) & 3, kw = threadIdx.x & 3;
    patch[threadIdx.x] = __ldg(img + (((long long)n * 3 + ci) * IMGS + ph * 4 + kh) * IMGS + pw * 4 + kw);
  }
  __syncthreads();
  const int e = threadIdx.x;
  float acc = __ldg(bias + e);
  const float* wr = w + (long long)e * 48;
#pragma unroll
  for (int k = 0; k < 48; ++k) acc = fmaf(patch[k], __ldg(wr + k), acc);
  // LayerNorm over the E = 128 channels of this token
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float s = warp_sum(acc);
  if (lane == 0) red[0][warp] = s;
  __syncthreads();
  const float mean = (red[0][0] + red[0][1] + red[0][2] + red[0][3]) / (float)E;
  const float d = acc - mean;
  float q = warp_sum(d * d);
  if (lane == 0) red[1][warp] = q;
  __syncthreads();
  const float rstd = rsqrtf((red[1][0] + red[1][1] + red[1][2] + red[1][3]) / (float)E + 1e-5f);
  out[((long long)n * R * R + tok) * E + e] = d * rstd * __ldg(g + e) + __ldg(b + e) + __ldg(ape + (long long)tok * E + e);
}

void launch_swin_patch_embed(const float* img, const float* w, const float* bias, const float* g, const float* b,
                             const float* ape, float* out, int B, int IMGS, int R, int E, cudaStream_t st) {
  swin_patch_embed_kernel<<<B * R * R, 128, 0, st>>>(img, w, bias, g, b, ape, out, IMGS, R, E);
}

// (Shifted-)window multi-head self-attention (:149-193, :314-362): one block per (image, window, head).
// qkv [B*R*R, 3C] (q | k | v, head h at columns h*32) -> out [B*R*R, C] in the ORIGINAL token order
// (cyclic shift, window partition, window reverse and the reverse shift are index arithmetic here).
// scores = (q * hd^-0.5) . k + relative_position_bias[idx] + (-100 across shifted regions), softmax, . v
__global__ void __launch_bounds__(128) swin_window_attn_kernel(const float* __restrict__ qkv, const float* __restrict__ bias_table,  // [(2ws-1)^2][heads]
                                                               float* __restrict__ out, int R, int C, int heads, int ws,
                                                               int shift) {
  extern __shared__ float sm[];
  const int N = ws * ws, LD = 33;
  float* Q = sm;
  float* K = Q + N * LD;
  float* V = K + N * LD;
  float* P = V + N * LD;         // [4 warps][N]
  int* tokidx = reinterpret_cast<int*>(P + 4 * N);  // [N] original token index of window position i
  int* region = tokidx + N;      // [N] shifted-window region id (attention allowed only within a region)
  const int nWr = R / ws;
  const int hh = blockIdx.x % heads;
  const int win = (blockIdx.x / heads) % (nWr * nWr);
  const int n = blockIdx.x / (heads * nWr * nWr);
  const int wh = win / nWr, ww = win % nWr;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int hs = wh * ws + i / ws, wsft = ww * ws + i % ws;     // coordinates in the shifted image
    const int h = (hs + shift) % R, w = (wsft + shift) % R;       // roll(x, -shift): shifted[hs] = x[hs + shift]
    tokidx[i] = h * R + w;
    int rid = 0;
    if (shift > 0) {
      const int hr = hs < R - ws ? 0 : (hs < R - shift ? 1 : 2);
      const int wr = wsft < R - ws ? 0 : (wsft < R - shift ? 1 : 2);
      rid = hr * 3 + wr;
    }
    region[i] = rid;
  }
  __syncthreads();
  const float scale = rsqrtf(32.f);
  const float* base = qkv + (long long)n * R * R * 3 * C + hh * 32;
  for (int i = threadIdx.x; i < N * 32; i += blockDim.x) {
    const int r = i >> 5, c = i & 31;
    const float* rp = base + (long long)tokidx[r] * 3 * C + c;
    Q[r * LD + c] = __ldg(rp) * scale;
    K[r * LD + c] = __ldg(rp + C);
    V[r * LD + c] = __ldg(rp + 2 * C);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Pw = P + warp * N;
  for (int i = warp; i < N; i += 4) {
    const int ri = i / ws, ci = i % ws, regi = region[i];
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) s = fmaf(Q[i * LD + c], K[j * LD + c], s);
      const int rel = (ri - j / ws + ws - 1) * (2 * ws - 1) + (ci - j % ws + ws - 1);
      s += __ldg(bias_table + (long long)rel * heads + hh);
      if (region[j] != regi) s += -100.f;
      Pw[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float e = expf(Pw[j] - mx);
      Pw[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(Pw[j] / sum, V[j * LD + lane], acc);
    out[((long long)n * R * R + tokidx[i]) * C + hh * 32 + lane] = acc;
    __syncwarp();
  }
}

void launch_swin_window_attn(const float* qkv, const float* bias_table, float* out, int B, int R, int C, int heads, int ws,
                             int shift, cudaStream_t st) {
  const int N = ws * ws;
  const size_t smem = (size_t)(3 * N * 33 + 4 * N) * sizeof(float) + 2 * N * sizeof(int);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(swin_window_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  const int nW = (R / ws) * (R / ws);
  swin_window_attn_kernel<<<B * nW * heads, 128, smem, st>>>(qkv, bias_table, out, R, C, heads, ws, shift);
}

// PatchMerging gather (:404-416): out[b, h2*R2 + w2, q*C + c] = x[b, (2*h2 + q%2)*R + 2*w2 + q/2, c], q = 0..3
// (x0 = even/even, x1 = odd row/even col, x2 = even row/odd col, x3 = odd/odd).
__global__ void __launch_bounds__(256) swin_patch_merge_kernel(const float* __restrict__ x, float* __restrict__ out, int B,
                                                               int R, int C) {
  const int R2 = R / 2, C4 = C / 4;
  long long total = (long long)B * R2 * R2 * 4 * C4;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = (int)(idx % C4);
  const int q = (int)((idx / C4) % 4);
  long long t = idx / (4 * C4);
  const int w2 = (int)(t % R2), h2 = (int)((t / R2) % R2), n = (int)(t / ((long long)R2 * R2));
  const float4 v = __ldg(reinterpret_cast<const float4*>(x + (((long long)n * R + 2 * h2 + (q & 1)) * R + 2 * w2 + (q >> 1)) * C) + c4);
  reinterpret_cast<float4*>(out + (((long long)n * R2 + h2) * R2 + w2) * 4 * C + (long long)q * C)[c4] = v;
}

void launch_swin_patch_merge(const float* x, float* out, int B, int R, int C, cudaStream_t st) {
  long long total = (long long)B * (R / 2) * (R / 2) * C;
  swin_patch_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, out, B, R, C);
}

}  // namespace frx
