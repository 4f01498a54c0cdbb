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
  static SmemOptIn opt;
  opt.ensure(swin_window_attn_kernel, smem);
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

// ---------------------------------------------------------------------------
// bf16-mode (shifted-)window attention on mma.sync.m16n8k16: same indexing as swin_window_attn_kernel (one block per
// (image, window, head), cyclic shift / window partition / reverse as index arithmetic), window 12 x 12 = 144 tokens =
// 9 query tiles x 18 key tiles, head dim 32.  q*hd^-0.5, k (row-major) and v (transposed) are staged in shared memory
// as bf16; S = Q K^T (36 HMMA per query tile), + relative-position bias + (-100 across shifted regions), softmax in
// the accumulator layout, O = P V (36 HMMA) with P re-used as the A fragment; output bf16 (the A operand of the
// projection GEMM) in the ORIGINAL token order.
// ---------------------------------------------------------------------------
namespace {
__device__ __forceinline__ uint32_t sw_pack2(float lo, float hi) {
  return eh2_pack(lo, hi);
}
__device__ __forceinline__ void sw_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32." FRX_EH_PTX "." FRX_EH_PTX ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
}  // namespace

constexpr int SWA_N = 144, SWA_WS = 12, SWA_LDK = 40, SWA_LDV = 152, SWA_NB = (2 * SWA_WS - 1) * (2 * SWA_WS - 1);

__global__ void __launch_bounds__(128) swin_window_attn_mma_kernel(const float* __restrict__ qkv, const float* __restrict__ bias_table,
                                                                   eh_t* __restrict__ out, int R, int C, int heads, int shift) {
  __shared__ __align__(16) eh_t Qb[SWA_N * SWA_LDK];
  __shared__ __align__(16) eh_t Kb[SWA_N * SWA_LDK];
  __shared__ __align__(16) eh_t Vt[32 * SWA_LDV];
  __shared__ float bias_h[SWA_NB];
  __shared__ int tokidx[SWA_N];
  __shared__ int colinfo[SWA_N];  // y | x << 8 | region << 16 of window position j
  constexpr int ws = SWA_WS, N = SWA_N;
  const int nWr = R / ws;
  const int hh = blockIdx.x % heads;
  const int win = (blockIdx.x / heads) % (nWr * nWr);
  const int n = blockIdx.x / (heads * nWr * nWr);
  const int wh = win / nWr, ww = win % nWr;
  for (int i = threadIdx.x; i < N; i += 128) {
    const int hs = wh * ws + i / ws, wsft = ww * ws + i % ws;     // coordinates in the shifted image
    const int h = (hs + shift) % R, w = (wsft + shift) % R;       // roll(x, -shift): shifted[hs] = x[hs + shift]
    tokidx[i] = h * R + w;
    int rid = 0;
    if (shift > 0) {
      const int hr = hs < R - ws ? 0 : (hs < R - shift ? 1 : 2);
      const int wr = wsft < R - ws ? 0 : (wsft < R - shift ? 1 : 2);
      rid = hr * 3 + wr;
    }
    colinfo[i] = (i / ws) | ((i % ws) << 8) | (rid << 16);
  }
  for (int i = threadIdx.x; i < SWA_NB; i += 128) bias_h[i] = __ldg(bias_table + (long long)i * heads + hh);
  __syncthreads();
  const float scale = rsqrtf(32.f);
  const float* base = qkv + (long long)n * R * R * 3 * C + hh * 32;
  for (int i = threadIdx.x; i < N * 32; i += 128) {
    const int r = i >> 5, c = i & 31;
    const float* rp = base + (long long)tokidx[r] * 3 * C + c;
    Qb[r * SWA_LDK + c] = eh_from_float(__ldg(rp) * scale);
    Kb[r * SWA_LDK + c] = eh_from_float(__ldg(rp + C));
    Vt[c * SWA_LDV + r] = eh_from_float(__ldg(rp + 2 * C));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
  const uint32_t* Qw = reinterpret_cast<const uint32_t*>(Qb);
  const uint32_t* Kw = reinterpret_cast<const uint32_t*>(Kb);
  const uint32_t* Vw = reinterpret_cast<const uint32_t*>(Vt);
  for (int mt = warp; mt < N / 16; mt += 4) {
    const int r0 = mt * 16 + gid, r1 = r0 + 8;
    uint32_t a[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      a[ks][0] = Qw[(r0 * SWA_LDK + 16 * ks + 2 * tig) >> 1];
      a[ks][1] = Qw[(r1 * SWA_LDK + 16 * ks + 2 * tig) >> 1];
      a[ks][2] = Qw[(r0 * SWA_LDK + 16 * ks + 8 + 2 * tig) >> 1];
      a[ks][3] = Qw[(r1 * SWA_LDK + 16 * ks + 8 + 2 * tig) >> 1];
    }
    float sacc[18][4];
#pragma unroll
    for (int nt = 0; nt < 18; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) sacc[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int kb = ((8 * nt + gid) * SWA_LDK + 16 * ks + 2 * tig) >> 1;
        sw_mma(sacc[nt], a[ks], Kw[kb], Kw[kb + 4]);
      }
    }
    // relative-position bias + shifted-window mask, then softmax over the 144 keys of rows r0 (c0, c1) and r1 (c2, c3)
    const int i0 = colinfo[r0], i1 = colinfo[r1];
    const int y0 = i0 & 255, x0 = (i0 >> 8) & 255, g0 = i0 >> 16, y1 = i1 & 255, x1 = (i1 >> 8) & 255, g1 = i1 >> 16;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 18; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cj = colinfo[8 * nt + 2 * tig + e];
        const int yj = cj & 255, xj = (cj >> 8) & 255, gj = cj >> 16;
        float s0 = sacc[nt][e] + bias_h[(y0 - yj + ws - 1) * (2 * ws - 1) + (x0 - xj + ws - 1)];
        float s1 = sacc[nt][2 + e] + bias_h[(y1 - yj + ws - 1) * (2 * ws - 1) + (x1 - xj + ws - 1)];
        if (gj != g0) s0 += -100.f;
        if (gj != g1) s1 += -100.f;
        sacc[nt][e] = s0; sacc[nt][2 + e] = s1;
        mx0 = fmaxf(mx0, s0); mx1 = fmaxf(mx1, s1);
      }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 18; ++nt) {
      sacc[nt][0] = __expf(sacc[nt][0] - mx0); sacc[nt][1] = __expf(sacc[nt][1] - mx0);
      sacc[nt][2] = __expf(sacc[nt][2] - mx1); sacc[nt][3] = __expf(sacc[nt][3] - mx1);
      sum0 += sacc[nt][0] + sacc[nt][1];
      sum1 += sacc[nt][2] + sacc[nt][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    float oacc[4][4];
#pragma unroll
    for (int nd = 0; nd < 4; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) oacc[nd][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 9; ++kk) {  // 16 keys per k-step = score tiles 2kk (k lo) and 2kk + 1 (k hi)
      const uint32_t pa[4] = {sw_pack2(sacc[2 * kk][0], sacc[2 * kk][1]), sw_pack2(sacc[2 * kk][2], sacc[2 * kk][3]),
                              sw_pack2(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]), sw_pack2(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3])};
#pragma unroll
      for (int nd = 0; nd < 4; ++nd) {
        const int vb = ((8 * nd + gid) * SWA_LDV + 16 * kk + 2 * tig) >> 1;
        sw_mma(oacc[nd], pa, Vw[vb], Vw[vb + 4]);
      }
    }
    const float q0 = __fdividef(1.f, sum0), q1 = __fdividef(1.f, sum1);
    eh_t* o0 = out + ((long long)n * R * R + tokidx[r0]) * C + hh * 32 + 2 * tig;
    eh_t* o1 = out + ((long long)n * R * R + tokidx[r1]) * C + hh * 32 + 2 * tig;
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) {
      *reinterpret_cast<uint32_t*>(o0 + 8 * nd) = sw_pack2(oacc[nd][0] * q0, oacc[nd][1] * q0);
      *reinterpret_cast<uint32_t*>(o1 + 8 * nd) = sw_pack2(oacc[nd][2] * q1, oacc[nd][3] * q1);
    }
  }
}

// returns false when the window is not 12 x 12 with 32-wide heads (the caller then uses the fp32 kernel)
bool launch_swin_window_attn_mma(const float* qkv, const float* bias_table, eh_t* out, int B, int R, int C, int heads,
                                 int ws, int shift, cudaStream_t st) {
  if (ws != SWA_WS || C != heads * 32 || R % ws != 0) return false;
  const int nW = (R / ws) * (R / ws);
  swin_window_attn_mma_kernel<<<B * nW * heads, 128, 0, st>>>(qkv, bias_table, out, R, C, heads, shift);
  return true;
}

}  // namespace frx
