// kernels_decode_bf16.cu -- the whole greedy decode loop as ONE persistent kernel.
//
// Design (DESIGN.md "decode"): every operation of the decoder step is local to
// an image, so there is no reason for grid-wide synchronisation.  A thread-block
// CLUSTER of 8 CTAs owns 16 images (the M of one mma.m16n8k16) for all steps:
//   * each CTA computes 1/8 of the output columns of every linear layer, with
//     the bf16 weights streamed from L2 straight into tensor-core B fragments
//     (pre-packed on the host in fragment order -> coalesced 16-byte loads);
//   * the partial rows are all-gathered through distributed shared memory with
//     st.async: each CTA stores its slice into all 8 CTAs' buffers and every store
//     completes transaction bytes on the DESTINATION CTA's mbarrier, so a stage is
//     over for a CTA exactly when all bytes it needs have landed -- no cluster
//     barrier (whose release semantics cost a MEMBAR.ALL.GPU each), no fence, no
//     global-memory round trips, no grid sync;
//   * attention over the bf16 KV cache (head-major [L][B][H][T][32], 64 B rows)
//     is done by one warp per (image, head) with 16-byte coalesced loads and
//     warp-shuffle reductions; the step's K/V rows are routed through DSMEM to
//     the CTA that owns the image, which writes (and later reads) them itself, so
//     no cross-CTA global-memory visibility is ever required.
// Tensor cores are used through mma.sync (M=16 images per cluster); tcgen05's
// minimum tile (M=64..128 rows) does not fit a 16-row per-step problem, and the
// step is bandwidth/latency-bound, not MMA-bound.
//
// Implements the recurrence of SURVEY App. A.4 (networks/EfficientSATRN.py:374-397,
// :539-558): cached layer OUTPUTS, current layer INPUT as the last key, scores
// divided by sqrt(d_model), post-LN, ReLU after both FFN linears.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace frx {

namespace {

constexpr int CL = DEC_CLUSTER;   // CTAs per cluster
constexpr int IMG = DEC_IMG;      // smem rows per cluster (= M of the MMA); NIMG <= IMG of them hold images
constexpr int D = 256;            // decoder width this kernel is specialised for
constexpr int HD = 32;
constexpr int H = D / HD;         // 8 heads
constexpr int NTHR = 256;
constexpr int APAD = 8;           // bf16 padding of the A-operand rows (bank-conflict-free fragments)

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint64_t make_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}

__device__ __forceinline__ uint4 ldg_weights(const uint4* p, uint64_t pol) {
  uint4 v;  // weights are re-read every step by every cluster: ask L2 to keep them
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;\n"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
  return v;
}

// KV-cache rows are written by other SMs earlier in this kernel: coherent (L2) load, L1 bypassed.
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

struct Smem {
  float xres[IMG][D];              // residual stream (fp32)
  float pre[IMG][D];               // gathered pre-LayerNorm rows
  float q[IMG][D];                 // q of the current layer input (also q2, logits)
  __nv_bfloat16 kv[IMG][2 * D];    // k | v of the current layer input (the extra "current" key/value)
  __nv_bfloat16 abf[IMG][D + APAD];        // A operand (bf16) of the next GEMM, K = D
  __nv_bfloat16 abf2[IMG][DEC_FMAX + APAD];  // A operand of FFN linear1, K = F
  float red[4][32][4];             // K-split partial accumulators
  long long prof[16];
  __nv_bfloat16 kvrow[IMG / 8][2][D];  // this step's K|V rows of the image(s) this CTA owns (sent by the 8 column owners)
  unsigned long long bar[2];           // stage mbarriers (alternate by stage parity)
  DecClusterLayer lw[4];           // per-layer pointers (dynamic indexing of kernel params would spill them)
};

// ---------------------------------------------------------------------------
// One GEMM stage: out[16, NT*8 columns of this CTA] = A[16, K] * W^T, A in smem
// (bf16), W pre-packed for this CTA: [NT tiles][K/32][32 lanes] uint4.
// NT >= 8: warp w owns tiles w, w+8, ...; NT == 4: two warps split K per tile.
// Epi(tile, acc, lane) is called by the warp that holds the final accumulator.
// ---------------------------------------------------------------------------
// Weight prefetch: the first PF k-pairs of every tile a warp owns are loaded into registers BEFORE the
// previous stage's barrier wait / LayerNorm, so their L2 latency overlaps that wait.  (Weights are static;
// nothing orders these loads against the exchange.)
template <int K, int NT>
struct WPre {
  static constexpr int TPW = NT >= 8 ? (NT + 7) / 8 : 1;
  static constexpr int KP = K / 32;
  static constexpr int KPW = NT >= 8 ? KP : KP / 2;  // k-pairs per warp task
  static constexpr int PF = NT >= 8 ? 2 : 4;         // prefetched k-pairs
  uint4 w[TPW][PF];
};

template <int K, int NT>
__device__ __forceinline__ WPre<K, NT> prefetch_weights(const uint4* __restrict__ Wp, uint64_t pol) {
  using P = WPre<K, NT>;
  P pre;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if constexpr (NT >= 8) {
#pragma unroll
    for (int i = 0; i < P::TPW; ++i) {
      const int tile = warp + 8 * i;
#pragma unroll
      for (int j = 0; j < P::PF; ++j)
        pre.w[i][j] = tile < NT ? ldg_weights(Wp + ((size_t)tile * P::KP + j) * 32 + lane, pol) : make_uint4(0u, 0u, 0u, 0u);
    }
  } else {
    const int tile = warp & 3, half = warp >> 2;
#pragma unroll
    for (int j = 0; j < P::PF; ++j)
      pre.w[0][j] = ldg_weights(Wp + ((size_t)tile * P::KP + half * P::KPW + j) * 32 + lane, pol);
  }
  return pre;
}

template <int K, int NT, typename Epi>
__device__ __forceinline__ void gemm_stage(Smem& s, uint64_t pol, const __nv_bfloat16* A, int lda,
                                           const uint4* __restrict__ Wp, const WPre<K, NT>& pre, Epi epi) {
  using P = WPre<K, NT>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  constexpr int KP = K / 32;
  // A fragments of two k16 steps with two ldmatrix.x4 (instead of eight 32-bit LDS): lane -> row address of
  // matrix (lane >> 3): {rows 0-7, k lo}, {rows 8-15, k lo}, {rows 0-7, k hi}, {rows 8-15, k hi} = a0..a3.
  const uint32_t a_lane = smem_addr_u32(A + ((lane & 7) + ((lane >> 3) & 1) * 8) * lda + ((lane >> 4) & 1) * 8);
  auto a_frags = [&](int kp, uint32_t (&a0)[4], uint32_t (&a1)[4]) {
    const uint32_t addr = a_lane + (uint32_t)kp * 64u;  // 32 bf16 per k-pair
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(a0[0]), "=r"(a0[1]), "=r"(a0[2]), "=r"(a0[3]) : "r"(addr));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(a1[0]), "=r"(a1[1]), "=r"(a1[2]), "=r"(a1[3]) : "r"(addr + 32u));
  };
  (void)gid; (void)tig;
  if constexpr (NT >= 8) {
    constexpr int TPW = P::TPW;
    float acc[TPW][2][4];
#pragma unroll
    for (int i = 0; i < TPW; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
#pragma unroll
    for (int kp = 0; kp < P::PF; ++kp) {  // prefetched k-pairs
      uint32_t a0[4], a1[4];
      a_frags(kp, a0, a1);
#pragma unroll
      for (int i = 0; i < TPW; ++i) {
        if (warp + 8 * i < NT) {
          mma_bf16(acc[i][0], a0, pre.w[i][kp].x, pre.w[i][kp].y);
          mma_bf16(acc[i][1], a1, pre.w[i][kp].z, pre.w[i][kp].w);
        }
      }
    }
#pragma unroll 3
    for (int kp = P::PF; kp < KP; ++kp) {
      uint4 w[TPW];
#pragma unroll
      for (int i = 0; i < TPW; ++i) {
        int tile = warp + 8 * i;
        if (tile < NT) w[i] = ldg_weights(Wp + ((size_t)tile * KP + kp) * 32 + lane, pol);
      }
      uint32_t a0[4], a1[4];
      a_frags(kp, a0, a1);
#pragma unroll
      for (int i = 0; i < TPW; ++i) {
        if (warp + 8 * i < NT) {
          mma_bf16(acc[i][0], a0, w[i].x, w[i].y);
          mma_bf16(acc[i][1], a1, w[i].z, w[i].w);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
      int tile = warp + 8 * i;
      if (tile < NT) {
        float c[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) c[e] = acc[i][0][e] + acc[i][1][e];
        epi(tile, c, lane);
      }
    }
  } else {
    static_assert(NT == 4, "NT must be 4 or >= 8");
    const int tile = warp & 3, half = warp >> 2;
    float acc[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
    constexpr int KH = P::KPW;
#pragma unroll
    for (int kk = 0; kk < P::PF; ++kk) {
      uint32_t a0[4], a1[4];
      a_frags(half * KH + kk, a0, a1);
      mma_bf16(acc[0], a0, pre.w[0][kk].x, pre.w[0][kk].y);
      mma_bf16(acc[1], a1, pre.w[0][kk].z, pre.w[0][kk].w);
    }
#pragma unroll 4
    for (int kk = P::PF; kk < KH; ++kk) {
      const int kp = half * KH + kk;
      uint4 w = ldg_weights(Wp + ((size_t)tile * KP + kp) * 32 + lane, pol);
      uint32_t a0[4], a1[4];
      a_frags(kp, a0, a1);
      mma_bf16(acc[0], a0, w.x, w.y);
      mma_bf16(acc[1], a1, w.z, w.w);
    }
    float c[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) c[e] = acc[0][e] + acc[1][e];
    if (half == 1) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s.red[tile][lane][e] = c[e];
    }
    __syncthreads();
    if (half == 0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) c[e] += s.red[tile][lane][e];
      epi(tile, c, lane);
    }
  }
}

// ---- DSMEM all-gather with st.async + mbarrier transaction counting ------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];\n" ::"r"(raddr), "r"(v), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_async_v2(uint32_t raddr, uint32_t a, uint32_t b, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];\n" ::"r"(raddr), "r"(a), "r"(b), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint4 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a byte-accounting bug traps (error reaches the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (int spins = 0; spins < (1 << 26); ++spins) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}

// Store the same value at the same smem offset of every CTA of the cluster; each store
// completes its byte count on the destination CTA's mbarrier `bar` (same offset everywhere).
__device__ __forceinline__ void ag_store_f2(uint32_t bar, float* local, float a, float b) {
  const uint32_t la = smem_u32(local);
#pragma unroll
  for (int r = 0; r < CL; ++r) st_async_v2(mapa_u32(la, r), __float_as_uint(a), __float_as_uint(b), mapa_u32(bar, r));
}
__device__ __forceinline__ void ag_store_u32(uint32_t bar, void* local, uint32_t v) {
  const uint32_t la = smem_u32(local);
#pragma unroll
  for (int r = 0; r < CL; ++r) st_async_b32(mapa_u32(la, r), v, mapa_u32(bar, r));
}
__device__ __forceinline__ void ag_store_u4(uint32_t bar, void* local, uint4 v) {
  const uint32_t la = smem_u32(local);
#pragma unroll
  for (int r = 0; r < CL; ++r) st_async_v4(mapa_u32(la, r), v, mapa_u32(bar, r));
}

// LayerNorm of the 16 gathered rows (every CTA does all rows: the result is
// needed everywhere and recomputing is cheaper than another exchange).
struct LnParams { float g[D / 32], b[D / 32]; };
__device__ __forceinline__ LnParams load_ln(const float* __restrict__ g, const float* __restrict__ b) {
  LnParams o;  // issued BEFORE the cluster barrier so the L2 latency overlaps the wait
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < D / 32; ++i) { o.g[i] = __ldg(g + i * 32 + lane); o.b[i] = __ldg(b + i * 32 + lane); }
  return o;
}

template <int NIMG>
__device__ __forceinline__ void layernorm_rows(Smem& s, const LnParams& P) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int rr = 0; rr < NIMG / 8; ++rr) {
    const int row = warp * (NIMG / 8) + rr;
    float v[D / 32];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < D / 32; ++i) { v[i] = s.pre[row][i * 32 + lane]; sum += v[i]; }
    const float mean = warp_sum(sum) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < D / 32; ++i) { float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < D / 32; ++i) {
      const int c = i * 32 + lane;
      float o = (v[i] - mean) * rstd * P.g[i] + P.b[i];
      s.xres[row][c] = o;
      s.abf[row][c] = __float2bfloat16_rn(o);
    }
  }
}

// Attention of one (image, head) pair by one warp, single pass (online softmax).
// Lane (g = lane>>2, c = lane&3) handles keys j = g (mod 8) and head dims [8c, 8c+8).
// K/V rows are 32 bf16 (64 B), contiguous over keys -> each warp load covers 512
// contiguous bytes; 4 K rows + 4 V rows are in flight per lane before any math.
// Every lane group keeps its own running (max, sum, acc) and the 8 groups are
// merged with shuffles at the end.  All lanes return the 8 output dims [8c, 8c+8).
__device__ __forceinline__ void attend_pair(const float* __restrict__ q, const __nv_bfloat16* __restrict__ Kc,
                                            const __nv_bfloat16* __restrict__ Vc, int n_hist,
                                            const __nv_bfloat16* __restrict__ kx,
                                            const __nv_bfloat16* __restrict__ vx, float inv_temp,
                                            float (&o)[8]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  float qv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qv[i] = q[c * 8 + i];
  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (int j0 = 0; j0 < n_hist; j0 += 32) {
    uint4 kk[4], vv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * 8 + g;
      kk[u] = j < n_hist ? ldg_stream(Kc + (size_t)j * HD + c * 8) : zero;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * 8 + g;
      vv[u] = j < n_hist ? ldg_stream(Vc + (size_t)j * HD + c * 8) : zero;
    }
    float sv[4];
    float cm = m;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float kf[8];
      unpack8(kk[u], kf);
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) part = fmaf(qv[i], kf[i], part);
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      sv[u] = (j0 + u * 8 + g < n_hist) ? part * inv_temp : -INFINITY;
      cm = fmaxf(cm, sv[u]);
    }
    const float scale = (m == -INFINITY) ? 0.f : __expf(m - cm);
    l *= scale;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= scale;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float pj = (sv[u] == -INFINITY) ? 0.f : __expf(sv[u] - cm);
      float vf[8];
      unpack8(vv[u], vf);
      l += pj;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(pj, vf[i], acc[i]);
    }
    m = cm;
  }
  // ---- merge the 8 lane groups (+ the current key) --------------------------------
  float M = m;
  M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 4));
  M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 8));
  M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 16));
  float s_cur = -INFINITY;
  float vcur[8];
  if (kx) {
    float kf[8];
    unpack8(*reinterpret_cast<const uint4*>(kx + c * 8), kf);
    unpack8(*reinterpret_cast<const uint4*>(vx + c * 8), vcur);
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) part = fmaf(qv[i], kf[i], part);
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    s_cur = part * inv_temp;
    M = fmaxf(M, s_cur);
  }
  const float gs = (m == -INFINITY) ? 0.f : __expf(m - M);
  l *= gs;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] *= gs;
  l += __shfl_xor_sync(0xffffffffu, l, 4);
  l += __shfl_xor_sync(0xffffffffu, l, 8);
  l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 4);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  if (kx) {
    const float pc = __expf(s_cur - M);
    l += pc;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(pc, vcur[i], acc[i]);
  }
  const float inv = __fdividef(1.f, l);
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = acc[i] * inv;
}

}  // namespace

// ===========================================================================
// The persistent decode kernel
// ===========================================================================
template <int NIMG>
__global__ void __cluster_dims__(DEC_CLUSTER, 1, 1) __launch_bounds__(256, 2)
dec_cluster_bf16_kernel(const DecClusterP p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  cg::cluster_group cl = cg::this_cluster();
  const int r = (int)cl.block_rank();
  const int img0 = (blockIdx.x / CL) * NIMG;  // first image of this cluster
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint64_t pol = make_evict_last_policy();
  const float inv_temp = 1.f / sqrtf((float)D);  // D = 256: exact reciprocal of 16
  const float emb_scale = sqrtf((float)D);
  const int L = p.L, T = p.T, V = p.V, B = p.B;

  if (tid == 0) {
#pragma unroll
    for (int l = 0; l < 4; ++l) s.lw[l] = p.layer[l];
  }
  // ---- step 0 input: <SOS> embedding + position 0 --------------------------------
  for (int i = tid; i < IMG * D; i += NTHR) {
    const int row = i / D, c = i % D;
    float v = __ldg(p.emb + (size_t)p.sos * D + c) * emb_scale + __ldg(p.pe + c);
    s.xres[row][c] = v;
    s.abf[row][c] = __float2bfloat16_rn(v);
  }
  const uint32_t bar0 = smem_u32(&s.bar[0]), bar1 = smem_u32(&s.bar[1]);
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  cl.sync();  // every CTA of the cluster is resident and its barriers initialised before the first remote store
  // Stage protocol: thread 0 arms the stage's barrier with the bytes THIS CTA will receive, everybody
  // computes and st.async-stores into all 8 CTAs, everybody waits for the local barrier phase.
  int gstage = 0;
  constexpr uint32_t R8 = NIMG / 8;  // row halves of the MMA tile that hold images
  auto stage_bar = [&]() -> uint32_t { return (gstage & 1) ? bar1 : bar0; };
  auto stage_begin = [&](uint32_t bytes) { if (tid == 0) mbar_expect_tx(stage_bar(), bytes); };
  auto stage_end = [&]() { mbar_wait(stage_bar(), (uint32_t)((gstage >> 1) & 1)); ++gstage; };
  float* const logit_s = reinterpret_cast<float*>(&s.abf2[0][0]);  // [IMG][256] fp32 view (abf2 is idle between S8 and S7)

  const bool profiling = p.prof != nullptr && blockIdx.x == 0 && tid == 0;
  if (profiling) for (int i = 0; i < 16; ++i) s.prof[i] = 0;
  long long tprev = clock64();
  auto mark = [&](int id) {
    if (profiling) {
      long long now = clock64();
      s.prof[id] += now - tprev;
      tprev = now;
    }
  };
  for (int t = 0; t < p.steps; ++t) {
    for (int l = 0; l < L; ++l) {
      const DecClusterLayer& W = s.lw[l];
      // ---- S1 (layer 0 only): q | k | v of the embedded input -------------------------
      if (l == 0) {
        stage_begin(NIMG * 2048u);
        const uint32_t sb = stage_bar();
        const uint4* wp = p.w_first + (size_t)r * 12 * (D / 32) * 32;
        const auto pre_s1 = prefetch_weights<D, 12>(wp, pol);
        gemm_stage<D, 12>(s, pol, &s.abf[0][0], D + APAD, wp, pre_s1, [&](int tile, float (&c)[4], int ln) {
          const int seg = tile >> 2, col = seg * D + r * 32 + (tile & 3) * 8 + (ln & 3) * 2, row = ln >> 2;
          const float b0 = __ldg(p.b_first + col), b1 = __ldg(p.b_first + col + 1);
          if (seg == 0) {
            ag_store_f2(sb, &s.q[row][col], c[0] + b0, c[1] + b1);
            if (NIMG == 16) {
            ag_store_f2(sb, &s.q[row + 8][col], c[2] + b0, c[3] + b1);
            }
          } else {
            ag_store_u32(sb, &s.kv[row][col - D], pack_bf16(c[0] + b0, c[1] + b1));
            if (NIMG == 16) {
            ag_store_u32(sb, &s.kv[row + 8][col - D], pack_bf16(c[2] + b0, c[3] + b1));
            }
          }
        });
        mark(0);
        stage_end();
        mark(1);
      }
      // ---- S2: self attention over t cached rows + the current input row ------------------
      {
        stage_begin(NIMG * 512u);
        const uint32_t sb = stage_bar();
#pragma unroll
        for (int pp = 0; pp < NIMG / 8; ++pp) {
          const int pair = warp * (NIMG / 8) + pp;
          const int li = r * (NIMG / 8) + (pair >> 3), hh = pair & 7;  // cluster-local image, head
          const int b = img0 + li;
          float o[8];
          if (b < B) {
            const size_t base = ((((size_t)l * B + b) * H + hh) * T) * HD;
            attend_pair(&s.q[li][hh * HD], p.kself + base, p.vself + base, t, &s.kv[li][hh * HD],
                        &s.kv[li][D + hh * HD], inv_temp, o);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = 0.f;
          }
          if ((lane >> 2) == 0) {
            uint4 v = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
            ag_store_u4(sb, &s.abf[li][hh * HD + (lane & 3) * 8], v);
          }
        }
        mark(2);
      }
      const auto pre_s3 = prefetch_weights<D, 4>(W.w_o + (size_t)r * 4 * (D / 32) * 32, pol);
      stage_end();
      mark(3);
      // ---- S3: out_linear(a) + x -> pre ; LN -> u -----------------------------------------
      stage_begin(NIMG * 1024u);
      uint32_t sb = stage_bar();
      gemm_stage<D, 4>(s, pol, &s.abf[0][0], D + APAD, W.w_o + (size_t)r * 4 * (D / 32) * 32, pre_s3,
                       [&](int tile, float (&c)[4], int ln) {
                         const int col = r * 32 + tile * 8 + (ln & 3) * 2, row = ln >> 2;
                         const float b0 = __ldg(W.b_o + col), b1 = __ldg(W.b_o + col + 1);
                         ag_store_f2(sb, &s.pre[row][col], c[0] + b0 + s.xres[row][col], c[1] + b1 + s.xres[row][col + 1]);
                         if (NIMG == 16) {
                         ag_store_f2(sb, &s.pre[row + 8][col], c[2] + b0 + s.xres[row + 8][col],
                                     c[3] + b1 + s.xres[row + 8][col + 1]);
                         }
                       });
      mark(4);
      const LnParams lnp1 = load_ln(W.ln1_g, W.ln1_b);
      const auto pre_s4 = prefetch_weights<D, 4>(W.w_q2 + (size_t)r * 4 * (D / 32) * 32, pol);
      stage_end();
      mark(5);
      layernorm_rows<NIMG>(s, lnp1);
      __syncthreads();
      mark(6);
      // ---- S4: q2 = q_linear(u) ---------------------------------------------------------------
      stage_begin(NIMG * 1024u);
      sb = stage_bar();
      gemm_stage<D, 4>(s, pol, &s.abf[0][0], D + APAD, W.w_q2 + (size_t)r * 4 * (D / 32) * 32, pre_s4,
                       [&](int tile, float (&c)[4], int ln) {
                         const int col = r * 32 + tile * 8 + (ln & 3) * 2, row = ln >> 2;
                         const float b0 = __ldg(W.b_q2 + col), b1 = __ldg(W.b_q2 + col + 1);
                         ag_store_f2(sb, &s.q[row][col], c[0] + b0, c[1] + b1);
                         if (NIMG == 16) {
                         ag_store_f2(sb, &s.q[row + 8][col], c[2] + b0, c[3] + b1);
                         }
                       });
      mark(4);
      stage_end();
      mark(5);
      // ---- S5: cross attention over the S memory tokens --------------------------------------------
      {
        stage_begin(NIMG * 512u);
        sb = stage_bar();
#pragma unroll
        for (int pp = 0; pp < NIMG / 8; ++pp) {
          const int pair = warp * (NIMG / 8) + pp;
          const int li = r * (NIMG / 8) + (pair >> 3), hh = pair & 7;
          const int b = img0 + li;
          float o[8];
          if (b < B) {
            const size_t base = ((((size_t)l * B + b) * H + hh) * p.S) * HD;
            attend_pair(&s.q[li][hh * HD], p.kcross + base, p.vcross + base, p.S, nullptr, nullptr, inv_temp, o);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = 0.f;
          }
          if ((lane >> 2) == 0) {
            uint4 v = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
            ag_store_u4(sb, &s.abf[li][hh * HD + (lane & 3) * 8], v);
          }
        }
        mark(7);
      }
      const auto pre_s6 = prefetch_weights<D, 4>(W.w_o2 + (size_t)r * 4 * (D / 32) * 32, pol);
      stage_end();
      mark(3);
      // ---- S6: out_linear(c) + u -> pre ; LN -> w ------------------------------------------------
      stage_begin(NIMG * 1024u);
      sb = stage_bar();
      gemm_stage<D, 4>(s, pol, &s.abf[0][0], D + APAD, W.w_o2 + (size_t)r * 4 * (D / 32) * 32, pre_s6,
                       [&](int tile, float (&c)[4], int ln) {
                         const int col = r * 32 + tile * 8 + (ln & 3) * 2, row = ln >> 2;
                         const float b0 = __ldg(W.b_o2 + col), b1 = __ldg(W.b_o2 + col + 1);
                         ag_store_f2(sb, &s.pre[row][col], c[0] + b0 + s.xres[row][col], c[1] + b1 + s.xres[row][col + 1]);
                         if (NIMG == 16) {
                         ag_store_f2(sb, &s.pre[row + 8][col], c[2] + b0 + s.xres[row + 8][col],
                                     c[3] + b1 + s.xres[row + 8][col + 1]);
                         }
                       });
      mark(4);
      const LnParams lnp2 = load_ln(W.ln2_g, W.ln2_b);
      const auto pre_s7 = prefetch_weights<D, 16>(W.w_f0 + (size_t)r * 16 * (D / 32) * 32, pol);
      stage_end();
      mark(5);
      layernorm_rows<NIMG>(s, lnp2);
      __syncthreads();
      mark(6);
      // L2 prefetch of the K/V history the NEXT self-attention of this warp will stream (next layer, or layer 0
      // of the next step): late in the decode the 181 MB cache no longer fits L2, and these hints turn the
      // attention's DRAM round trips into L2 hits without holding any registers.
      if (t > 0) {
        const int ln_next = (l + 1 < L) ? l + 1 : 0;
        const int t_next = (l + 1 < L) ? t : t + 1;
#pragma unroll
        for (int pp = 0; pp < NIMG / 8; ++pp) {
          const int pair = warp * (NIMG / 8) + pp;
          const int li = r * (NIMG / 8) + (pair >> 3), hh = pair & 7;
          const int b = img0 + li;
          if (b < B) {
            const size_t base = ((((size_t)ln_next * B + b) * H + hh) * T) * HD;
            const int bytes = t_next * HD * 2;  // rows [0, t_next) of this (image, head)
            for (int off = lane * 128; off < bytes; off += 32 * 128) {
              asm volatile("prefetch.global.L2 [%0];\n" ::"l"(reinterpret_cast<const char*>(p.kself + base) + off));
              asm volatile("prefetch.global.L2 [%0];\n" ::"l"(reinterpret_cast<const char*>(p.vself + base) + off));
            }
          }
        }
      }
      // ---- S7: ff = relu(linear0(w))  (F = 4 segments of D columns) --------------------------------
      stage_begin(NIMG * 2048u);
      sb = stage_bar();
      gemm_stage<D, 16>(s, pol, &s.abf[0][0], D + APAD, W.w_f0 + (size_t)r * 16 * (D / 32) * 32, pre_s7,
                        [&](int tile, float (&c)[4], int ln) {
                          const int seg = tile >> 2, col = seg * D + r * 32 + (tile & 3) * 8 + (ln & 3) * 2, row = ln >> 2;
                          const float b0 = __ldg(W.b_f0 + col), b1 = __ldg(W.b_f0 + col + 1);
                          ag_store_u32(sb, &s.abf2[row][col], pack_bf16(fmaxf(c[0] + b0, 0.f), fmaxf(c[1] + b1, 0.f)));
                          if (NIMG == 16) {
                          ag_store_u32(sb, &s.abf2[row + 8][col], pack_bf16(fmaxf(c[2] + b0, 0.f), fmaxf(c[3] + b1, 0.f)));
                          }
                        });
      mark(8);
      const auto pre_s8 = prefetch_weights<DEC_FMAX, 4>(W.w_f1 + (size_t)r * 4 * (DEC_FMAX / 32) * 32, pol);
      stage_end();
      mark(5);
      // ---- S8: relu(linear1(ff)) + w -> pre ; LN -> y -----------------------------------------------
      stage_begin(NIMG * 1024u);
      sb = stage_bar();
      gemm_stage<DEC_FMAX, 4>(s, pol, &s.abf2[0][0], DEC_FMAX + APAD, W.w_f1 + (size_t)r * 4 * (DEC_FMAX / 32) * 32, pre_s8,
                              [&](int tile, float (&c)[4], int ln) {
                                const int col = r * 32 + tile * 8 + (ln & 3) * 2, row = ln >> 2;
                                const float b0 = __ldg(W.b_f1 + col), b1 = __ldg(W.b_f1 + col + 1);
                                ag_store_f2(sb, &s.pre[row][col], fmaxf(c[0] + b0, 0.f) + s.xres[row][col],
                                            fmaxf(c[1] + b1, 0.f) + s.xres[row][col + 1]);
                                if (NIMG == 16) {
                                ag_store_f2(sb, &s.pre[row + 8][col], fmaxf(c[2] + b0, 0.f) + s.xres[row + 8][col],
                                            fmaxf(c[3] + b1, 0.f) + s.xres[row + 8][col + 1]);
                                }
                              });
      mark(9);
      const LnParams lnp3 = load_ln(W.ln3_g, W.ln3_b);
      const bool last_layer = l + 1 >= L;
      // prefetch of S9's weights (two shapes: next-layer q|k|v or the generator) overlaps the S8 wait + LayerNorm
      WPre<D, 20> pre_s9a;
      WPre<D, 12> pre_s9b;
      if (!last_layer) pre_s9a = prefetch_weights<D, 20>(W.w_next + (size_t)r * 20 * (D / 32) * 32, pol);
      else pre_s9b = prefetch_weights<D, 12>(W.w_next + (size_t)r * 12 * (D / 32) * 32, pol);
      stage_end();
      mark(5);
      layernorm_rows<NIMG>(s, lnp3);
      __syncthreads();
      mark(6);
      // ---- S9: K/V rows of y -> cache; next layer's q|k|v, or the vocabulary logits -------------------
      stage_begin(NIMG * (l + 1 < L ? 2048u : 1024u) + NIMG * 128u);
      sb = stage_bar();
      auto kv_store = [&](int tile, float (&c)[4], int ln) {
        // tiles 0..3: K columns of head r; tiles 4..7: V columns of head r.  The values go to the CTA that
        // OWNS the image (it attends over that image's cache), which writes them to the global cache itself.
        const int seg = tile >> 2, dcol = (tile & 3) * 8 + (ln & 3) * 2, row = ln >> 2;
        const float* bias = W.b_next + seg * D + r * 32 + dcol;
        const float b0 = __ldg(bias), b1 = __ldg(bias + 1);
#pragma unroll
        for (int hrow = 0; hrow < NIMG / 8; ++hrow) {
          const int li = row + hrow * 8;                 // cluster-local image
          const uint32_t owner = (uint32_t)(li / (NIMG / 8));
          const uint32_t la = smem_u32(&s.kvrow[li % (NIMG / 8)][seg][r * 32 + dcol]);
          st_async_b32(mapa_u32(la, owner), pack_bf16(c[hrow * 2] + b0, c[hrow * 2 + 1] + b1), mapa_u32(sb, owner));
        }
      };
      if (l + 1 < L) {
        gemm_stage<D, 20>(s, pol, &s.abf[0][0], D + APAD, W.w_next + (size_t)r * 20 * (D / 32) * 32, pre_s9a,
                          [&](int tile, float (&c)[4], int ln) {
                            if (tile < 8) { kv_store(tile, c, ln); return; }
                            const int seg = tile >> 2;  // 2,3,4 -> q,k,v of layer l+1
                            const int col = (seg - 2) * D + r * 32 + (tile & 3) * 8 + (ln & 3) * 2, row = ln >> 2;
                            const float b0 = __ldg(W.b_next + 2 * D + col), b1 = __ldg(W.b_next + 2 * D + col + 1);
                            if (seg == 2) {
                              ag_store_f2(sb, &s.q[row][col], c[0] + b0, c[1] + b1);
                              if (NIMG == 16) {
                              ag_store_f2(sb, &s.q[row + 8][col], c[2] + b0, c[3] + b1);
                              }
                            } else {
                              ag_store_u32(sb, &s.kv[row][col - D], pack_bf16(c[0] + b0, c[1] + b1));
                              if (NIMG == 16) {
                              ag_store_u32(sb, &s.kv[row + 8][col - D], pack_bf16(c[2] + b0, c[3] + b1));
                              }
                            }
                          });
      } else {
        // generator: V columns padded to 256; CTA r owns columns [32r, 32r+32) (tiles 8..11)
        gemm_stage<D, 12>(s, pol, &s.abf[0][0], D + APAD, W.w_next + (size_t)r * 12 * (D / 32) * 32, pre_s9b,
                          [&](int tile, float (&c)[4], int ln) {
                            if (tile < 8) { kv_store(tile, c, ln); return; }
                            const int col = r * 32 + (tile - 8) * 8 + (ln & 3) * 2, row = ln >> 2;
                            const float b0 = col < V ? __ldg(W.b_next + 2 * D + col) : 0.f;
                            const float b1 = col + 1 < V ? __ldg(W.b_next + 2 * D + col + 1) : 0.f;
                            const float v00 = c[0] + b0, v01 = c[1] + b1, v10 = c[2] + b0, v11 = c[3] + b1;
                            ag_store_f2(sb, logit_s + row * 256 + col, v00, v01);
                            if (NIMG == 16) {
                            ag_store_f2(sb, logit_s + (row + 8) * 256 + col, v10, v11);
                            }
                            if (p.logits) {
                              const int bA = img0 + row, bB = img0 + row + 8;
                              if (bA < B) {
                                float* lp = p.logits + ((size_t)bA * p.steps + t) * V;
                                if (col < V) lp[col] = v00;
                                if (col + 1 < V) lp[col + 1] = v01;
                              }
                              if (NIMG == 16 && bB < B) {
                                float* lp = p.logits + ((size_t)bB * p.steps + t) * V;
                                if (col < V) lp[col] = v10;
                                if (col + 1 < V) lp[col + 1] = v11;
                              }
                            }
                          });
      }
      mark(10);
      stage_end();
      // write-out of the K/V rows this CTA owns: 16-byte stores, 64 contiguous bytes per (head, K|V)
      if (tid < 64 * (NIMG / 8)) {
        const int j = tid >> 6, rem = tid & 63, seg = rem >> 5, hh = (rem >> 2) & 7, ch = rem & 3;
        const int b = img0 + r * (NIMG / 8) + j;
        if (b < B) {
          const uint4 v = *reinterpret_cast<const uint4*>(&s.kvrow[j][seg][hh * HD + ch * 8]);
          __nv_bfloat16* dst = seg == 0 ? p.kself : p.vself;
          *reinterpret_cast<uint4*>(dst + (((((size_t)l * B + b) * H + hh) * T) + t) * HD + ch * 8) = v;
        }
      }
      mark(5);
    }  // layers
    // ---- greedy pick (first max index) + next input: every CTA does all 16 rows -------------------------
#pragma unroll
    for (int rr = 0; rr < NIMG / 8; ++rr) {
      const int row = warp * (NIMG / 8) + rr;
      float best = -INFINITY;
      int bi = 0x7fffffff;
      for (int i = lane; i < V; i += 32) {
        float v = logit_s[row * 256 + i];
        if (v > best) { best = v; bi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (bi < 0 || bi >= V) bi = 0;
      const int b = img0 + row;
      int nxt = bi;
      if (b < B) {
        if (p.forced) nxt = (int)p.forced[(size_t)b * p.steps + t];
        if (r == 0 && lane == 0 && p.tokens) p.tokens[(size_t)b * p.steps + t] = bi;
      }
      if (t + 1 < p.steps) {
        const float* pe = p.pe + (size_t)(t + 1) * D;
#pragma unroll
        for (int i = 0; i < D / 32; ++i) {
          const int c = i * 32 + lane;
          float v = __ldg(p.emb + (size_t)nxt * D + c) * emb_scale + __ldg(pe + c);
          s.xres[row][c] = v;
          s.abf[row][c] = __float2bfloat16_rn(v);
        }
      }
    }
    __syncthreads();
    mark(11);
  }
  cl.sync();  // nobody exits while a peer's stores to it may still be in flight
  if (profiling) for (int i = 0; i < 16; ++i) p.prof[i] = s.prof[i];
}

size_t dec_cluster_smem_bytes() { return sizeof(Smem); }

template <int NIMG>
static int launch_variant(const DecClusterP& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dec_cluster_bf16_kernel<NIMG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(Smem));
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int clusters = (p.B + NIMG - 1) / NIMG;
  if (getenv("FRX_DEBUG")) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(clusters * CL); cfg.blockDim = dim3(NTHR); cfg.dynamicSmemBytes = sizeof(Smem);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dec_cluster_bf16_kernel<NIMG>, &cfg);
    fprintf(stderr, "[frx] decode kernel: %d clusters of %d CTAs requested, max co-resident clusters = %d (%s), smem %zu B\n",
            clusters, CL, n, cudaGetErrorString(e), sizeof(Smem));
  }
  dec_cluster_bf16_kernel<NIMG><<<clusters * CL, NTHR, sizeof(Smem), st>>>(p);
  return 0;
}

// 8 images per cluster while all clusters can be co-resident (two CTAs per SM: the
// second CTA's latency chains fill the first one's stalls); 16 per cluster beyond.
int launch_dec_cluster_bf16(const DecClusterP& p, int images_per_cluster, cudaStream_t st) {
  int n = images_per_cluster;
  if (n == 0) n = (p.B + 7) / 8 <= DEC_MAX_CLUSTERS_8 ? 8 : 16;
  return n == 8 ? launch_variant<8>(p, st) : launch_variant<16>(p, st);
}

// ===========================================================================
// cross K/V: fp32 [B*S][L*2*D] (k_l | v_l per layer) -> bf16 head-major
// [L][B][H][S][32] caches
// ===========================================================================
__global__ void __launch_bounds__(256) cross_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ kc,
                                                            __nv_bfloat16* __restrict__ vc, int B, int S, int L, int Dm) {
  const int Hh = Dm / 32;
  long long total = (long long)B * S * L * 2 * Dm;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int col = (int)(idx % (L * 2 * Dm));
  long long row = idx / (L * 2 * Dm);
  int sidx = (int)(row % S), b = (int)(row / S);
  int l = col / (2 * Dm), which = (col / Dm) & 1, d = col % Dm;
  int hh = d / 32, dd = d % 32;
  size_t off = (((((size_t)l * B + b) * Hh + hh) * S) + sidx) * 32 + dd;
  (which ? vc : kc)[off] = __float2bfloat16_rn(src[idx]);
}

void launch_cross_to_bf16(const float* src, __nv_bfloat16* kc, __nv_bfloat16* vc, int B, int S, int L, int Dm,
                          cudaStream_t st) {
  long long total = (long long)B * S * L * 2 * Dm;
  cross_to_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, kc, vc, B, S, L, Dm);
}

}  // namespace frx
