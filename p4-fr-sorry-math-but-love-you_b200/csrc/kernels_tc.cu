// kernels_tc.cu -- tcgen05 (5th-gen tensor core) implicit-GEMM for the dense
// contractions: the EfficientNetV2 trunk convolutions (3x3 as implicit GEMM, 1x1
// as plain GEMM), the SATRN encoder projections / 1x1 convs, the cross-attention
// K/V projection.
//
//   C[M, N] = epilogue( A[M, K] * W[N, K]^T )        bf16 operands, fp32 accumulate
//
// * A is either a dense row-major bf16 matrix or an NHWC bf16 activation that is
//   gathered on the fly (k index = (kh*KW + kw)*Cin + ci, TF-"same"/explicit
//   padding zero-filled) -- no im2col buffer ever touches HBM.
// * Operand tiles (A: 128 x 64, W: BN x 64 bf16) are staged in shared memory in
//   the canonical K-major SWIZZLE_128B UMMA layout by all 256 threads with 16-byte
//   cp.async (zero-fill for padding / ragged edges), 3-stage ring.
// * One elected thread issues tcgen05.mma (M=128, N=BN, K=16 x4 per stage) with
//   the accumulator in TMEM; tcgen05.commit arrives on an mbarrier per stage so
//   the ring slot can be refilled while later MMAs run.
// * Epilogue: tcgen05.ld (32 lanes x 16 columns per warp instruction) -> folded
//   BatchNorm / bias, activation, residual -> bf16 or fp32 store.
//
// Descriptor encodings follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor /
// InstrDescriptor) of the CUTLASS tree vendored in this image.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace frx {

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_STAGES = 3;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;  // 16 KiB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA: 2-D tiled bulk tensor copy global -> shared (128B-swizzled box), completion on an mbarrier.
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
// TMA im2col mode: `pixelsPerColumn` output pixels starting at base pixel (w, h, n), filter tap (w_off, h_off),
// `channelsPerPixel` channels from c -- one instruction stages a whole 128 x 64 implicit-GEMM A tile.
__device__ __forceinline__ void tma_load_im2col(uint32_t smem_dst, const CUtensorMap* tmap, int c, int w, int h, int n,
                                                uint16_t w_off, uint16_t h_off, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};\n"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n),
        "h"(w_off), "h"(h_off) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded spin: a protocol bug traps (error returned to the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  for (int spins = 0; spins < (1 << 24); ++spins) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (8-row groups of 1024 B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                  // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                  // layout type SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: D=F32, A=B=eh_t (format 0 = F16, 1 = BF16), both K-major, M=128, N=BN.
__device__ __forceinline__ uint32_t make_idesc(int bn) {
  constexpr uint32_t fmt = FRX_ENC_FP16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// bf16-mode activations: fast intrinsics (2-ulp exp / division are far below bf16 output rounding)
template <int ACT>
__device__ __forceinline__ float act_fast(float v) {
  if (ACT == ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == ACT_SILU) return __fdividef(v, 1.f + __expf(-v));
  if (ACT == ACT_SIGMOID) return __fdividef(1.f, 1.f + __expf(-v));
  if (ACT == ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));  // exact GELU (SWIN.py:30)
  return v;
}

// SiLU of two values with ONE special-function op: x sigmoid(x) = h + h tanh(h), h = x/2, tanh as tanh.approx.f16x2
// (ex2 + rcp per value made the 16-op/clk MUFU pipe the floor of every small-K SiLU layer: 4096 clk per 128 x 256 tile
// against 2168 clk of MMA at K = 256). The product stays in fp32; the result is stored as bf16.
#ifndef FRX_SILU_EXACT
#define FRX_SILU_EXACT 1
#endif
__device__ __forceinline__ void silu_pair(float& a, float& b) {
#if FRX_SILU_EXACT
  a = __fdividef(a, 1.f + __expf(-a));
  b = __fdividef(b, 1.f + __expf(-b));
  return;
#endif
  const float ha = 0.5f * a, hb = 0.5f * b;
  const __half2 h = __floats2half2_rn(ha, hb);
  uint32_t t;
  asm("tanh.approx.f16x2 %0, %1;\n" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t*>(&h)));
  const float2 tf = __half22float2(*reinterpret_cast<const __half2*>(&t));
  a = fmaf(ha, tf.x, ha);
  b = fmaf(hb, tf.y, hb);
}

template <int ACT>
__device__ __forceinline__ void epi_affine_act(float (&v)[16], const float* __restrict__ scale,
                                               const float* __restrict__ shift, int nb) {
  if (scale) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 s = __ldg(reinterpret_cast<const float4*>(scale + nb + i));
      const float4 h = __ldg(reinterpret_cast<const float4*>(shift + nb + i));
      if (ACT == ACT_SILU) {
        v[i] = fmaf(v[i], s.x, h.x); v[i + 1] = fmaf(v[i + 1], s.y, h.y);
        v[i + 2] = fmaf(v[i + 2], s.z, h.z); v[i + 3] = fmaf(v[i + 3], s.w, h.w);
        silu_pair(v[i], v[i + 1]);
        silu_pair(v[i + 2], v[i + 3]);
      } else {
        v[i] = act_fast<ACT>(fmaf(v[i], s.x, h.x));
        v[i + 1] = act_fast<ACT>(fmaf(v[i + 1], s.y, h.y));
        v[i + 2] = act_fast<ACT>(fmaf(v[i + 2], s.z, h.z));
        v[i + 3] = act_fast<ACT>(fmaf(v[i + 3], s.w, h.w));
      }
    }
  } else if (shift) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 h = __ldg(reinterpret_cast<const float4*>(shift + nb + i));
      v[i] = act_fast<ACT>(v[i] + h.x);
      v[i + 1] = act_fast<ACT>(v[i + 1] + h.y);
      v[i + 2] = act_fast<ACT>(v[i + 2] + h.z);
      v[i + 3] = act_fast<ACT>(v[i + 3] + h.w);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = act_fast<ACT>(v[i]);
  }
}

}  // namespace

// Epilogue of one 16-column chunk of one row: folded BN / bias, activation, residual, store.
// GELU (erff) is compiled only into the instantiation SwinTRN uses: it costs registers every other layer would pay for.
template <bool GELU>
__device__ __forceinline__ void epilogue_store16(const TcGemmP& p, const uint32_t (&r)[16], int m, int nb) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
      const bool full = nb + 16 <= p.N;
      if (full) {  // vectorised scale/shift, activation resolved outside the element loop
        if (p.act == ACT_SILU) epi_affine_act<ACT_SILU>(v, p.scale, p.shift, nb);
        else if (p.act == ACT_RELU) epi_affine_act<ACT_RELU>(v, p.scale, p.shift, nb);
        else if (GELU && p.act == ACT_GELU) epi_affine_act<ACT_GELU>(v, p.scale, p.shift, nb);
        else epi_affine_act<ACT_NONE>(v, p.scale, p.shift, nb);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = nb + i;
          if (n < p.N) {
            if (p.scale) v[i] = v[i] * __ldg(p.scale + n) + __ldg(p.shift + n);
            else if (p.shift) v[i] += __ldg(p.shift + n);
            v[i] = act_apply(v[i], p.act);
          }
        }
      }
      if (p.res) {
        if (p.res_f32) {
          const float* rp = reinterpret_cast<const float*>(p.res) + (size_t)m * p.ldr + nb;
          if (full) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 rr = __ldg(reinterpret_cast<const float4*>(rp + i));
              v[i] += rr.x; v[i + 1] += rr.y; v[i + 2] += rr.z; v[i + 3] += rr.w;
            }
          } else {
            for (int i = 0; i < 16; ++i)
              if (nb + i < p.N) v[i] += __ldg(rp + i);
          }
        } else {
          const eh_t* rp = reinterpret_cast<const eh_t*>(p.res) + (size_t)m * p.ldr + nb;
          if (full) {
            const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(rp)), r1 = __ldg(reinterpret_cast<const uint4*>(rp + 8));
            const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 rf = eh2_unpack(rw[i]);
              v[2 * i] += rf.x;
              v[2 * i + 1] += rf.y;
            }
          } else {
            for (int i = 0; i < 16; ++i)
              if (nb + i < p.N) v[i] += eh_to_float(rp[i]);
          }
        }
      }
      if (p.out_f32) {
        float* cp = reinterpret_cast<float*>(p.C) + (size_t)m * p.ldc + nb;
        if (full && (p.ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 31) == 0) {  // two full 32-byte sectors per thread
#pragma unroll
          for (int i = 0; i < 16; i += 8)
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(cp + i), "f"(v[i]), "f"(v[i + 1]),
                         "f"(v[i + 2]), "f"(v[i + 3]), "f"(v[i + 4]), "f"(v[i + 5]), "f"(v[i + 6]), "f"(v[i + 7]) : "memory");
        } else if (full && (p.ldc & 3) == 0) {  // 16-byte stores need a 16-byte-aligned row stride (the logits' is 245 floats)
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(cp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
          for (int i = 0; i < 16; ++i)
            if (nb + i < p.N) cp[i] = v[i];
        }
      } else {
        eh_t* cp = reinterpret_cast<eh_t*>(p.C) + (size_t)m * p.ldc + nb;
        if (full && (p.ldc & 7) == 0) {
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = eh2_pack(v[2 * i], v[2 * i + 1]);
          if ((p.ldc & 15) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 31) == 0) {  // one full 32-byte sector per thread (256-bit store, sm_100+)
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(cp), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                         "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
          } else {
            *reinterpret_cast<uint4*>(cp) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(cp + 8) = make_uint4(w[4], w[5], w[6], w[7]);
          }
        } else {
          for (int i = 0; i < 16; ++i)
            if (nb + i < p.N) cp[i] = eh_from_float(v[i]);
        }
      }
}

template <int NCOLS>
__global__ void __launch_bounds__(256) tc_igemm_kernel(const TcGemmP p) {
  extern __shared__ unsigned char dyn_smem[];
  __shared__ __align__(8) uint64_t mma_done[TC_STAGES];
  __shared__ __align__(8) uint64_t acc_done;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.BN;
  const uint32_t b_bytes = (uint32_t)BN * (TC_BK * 2);
  const uint32_t stage_bytes = TC_A_BYTES + b_bytes;
  const int NS = p.stages;  // ring depth (2 or 3), chosen by the launcher from the K extent
  const uint32_t smem0 = (smem_u32(dyn_smem) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * BN;
  const int KB = (p.K + TC_BK - 1) / TC_BK;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 32) {
#pragma unroll
    for (int s = 0; s < TC_STAGES; ++s) mbar_init(&mma_done[s], 1);
    mbar_init(&acc_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  // ---- loader coordinates: thread -> 16-byte chunk c of rows rbase + 32*i ----------------
  const int c = tid & 7, rbase = tid >> 3;
  const eh_t* a_ptr[4];
  int a_ih0[4], a_iw0[4];
  bool a_ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + rbase + 32 * i;
    a_ok[i] = m < p.M;
    a_ih0[i] = a_iw0[i] = 0;
    a_ptr[i] = p.A;
    if (a_ok[i]) {
      if (p.conv) {
        const int ow = m % p.OW;
        const int t = m / p.OW;
        const int oh = t % p.OH;
        const int n = t / p.OH;
        a_ptr[i] = p.A + (size_t)n * p.H * p.Wd * p.Cin;
        a_ih0[i] = oh * p.stride - p.pad_t;
        a_iw0[i] = ow * p.stride - p.pad_l;
      } else {
        a_ptr[i] = p.A + (size_t)m * p.lda;
      }
    }
  }
  const int nb_rows = (BN + 31) / 32;  // B rows handled per thread (<= 8)

  auto load_stage = [&](int kb, int stage) {
    const uint32_t sa = smem0 + (uint32_t)stage * stage_bytes;
    const uint32_t sb = sa + TC_A_BYTES;
    const int k = kb * TC_BK + c * 8;
    // A: 128 rows
    int tap = 0, ci = k, kh = 0, kw = 0;
    if (p.conv) {
      tap = k / p.Cin;
      ci = k - tap * p.Cin;
      kh = tap / p.KW;
      kw = tap - kh * p.KW;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = rbase + 32 * i;
      const uint32_t dst = sa + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((c ^ (row & 7)) << 4);
      const eh_t* src = p.A;
      uint32_t bytes = 0;
      if (a_ok[i] && k < p.K) {
        if (p.conv) {
          const int ih = a_ih0[i] + kh, iw = a_iw0[i] + kw;
          if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.Wd) {
            src = a_ptr[i] + ((size_t)ih * p.Wd + iw) * p.Cin + ci;
            bytes = 16;
          }
        } else {
          src = a_ptr[i] + k;
          bytes = 16;
        }
      }
      cp_async16(dst, src, bytes);
    }
    // B: BN rows of the weight matrix [N][ldw]
    for (int j = 0; j < nb_rows; ++j) {
      const int row = rbase + 32 * j;
      if (row < BN) {
        const uint32_t dst = sb + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((c ^ (row & 7)) << 4);
        const int n = n0 + row;
        const bool ok = n < p.N && k < p.K;
        cp_async16(dst, ok ? p.W + (size_t)n * p.ldw + k : p.W, ok ? 16u : 0u);
      }
    }
  };

  // ---- main loop --------------------------------------------------------------------------
  for (int s = 0; s < NS - 1; ++s) {
    if (s < KB) load_stage(s, s);
    cp_async_commit();
  }
  const uint32_t idesc = make_idesc(BN);
  for (int kb = 0; kb < KB; ++kb) {
    if (NS == 3) cp_async_wait<1>(); else cp_async_wait<0>();  // this thread's chunks of k-block kb have landed
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    const int stage = kb % NS;
    if (tid == 0) {
      tc_fence_after();
      const uint32_t sa = smem0 + (uint32_t)stage * stage_bytes;
      const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + TC_A_BYTES);
#pragma unroll
      for (int j = 0; j < TC_BK / 16; ++j)  // +32 bytes (2 x 16-byte units) per K=16 step inside the swizzle atom
        umma_f16(tmem, adesc + (uint64_t)(2 * j), bdesc + (uint64_t)(2 * j), idesc, (kb | j) ? 1u : 0u);
      umma_commit(&mma_done[stage]);
      if (kb == KB - 1) umma_commit(&acc_done);
    }
    const int nk = kb + NS - 1;
    if (nk < KB) {
      if (kb >= 1) mbar_wait(&mma_done[(kb - 1) % NS], (uint32_t)(((kb - 1) / NS) & 1));
      load_stage(nk, nk % NS);
    }
    cp_async_commit();
  }
  cp_async_wait<0>();
  mbar_wait(&acc_done, 0);
  tc_fence_after();

  // ---- epilogue: TMEM -> registers -> global ---------------------------------------------------
  {
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int m = m0 + q * 32 + lane;
    const bool m_ok = m < p.M;
    for (int cc = (warp >> 2); cc * 16 < BN; cc += 2) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 16), r);
      const int nb = n0 + cc * 16;
      if (m_ok && nb < p.N) epilogue_store16<false>(p, r, m, nb);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(NCOLS) : "memory");
  }
}


// ===========================================================================
// Persistent, warp-specialised version (the default):
//   warps 0-3  producers : cp.async gather of A (128 x 64) and W (BN x 64) into a 4-stage
//                          swizzled ring; per stage: wait group -> fence.proxy.async -> arrive(full)
//   warps 4-15 epilogue  : tcgen05.ld of the finished accumulator (TMEM buffer a) -> BN/bias, act,
//                          residual -> global; arrive(acc_empty[a])
//   warp  16   MMA       : one lane waits full[s], issues 4 x tcgen05.mma, commits to empty[s];
//                          after the last k-block commits to acc_full[a]
// Each CTA loops over output tiles (n-tile fastest, so concurrent CTAs share A tiles in L2); the
// TMEM accumulator is double-buffered, so tile i's epilogue overlaps tile i+1's loads and MMAs.
// ===========================================================================
constexpr int WS_STAGES = 4;
constexpr int WS_LAG = 2;            // cp.async groups kept in flight per producer thread
constexpr int WS_PROD_WARPS = 8;     // address generation for the gather is instruction-latency bound: spread it
constexpr int WS_EPI_WARPS = 8;      // 2 warps per TMEM lane quarter (4 when the feed is all-TMA: the gather warps join)
constexpr int WS_MMA_WARP = WS_PROD_WARPS + WS_EPI_WARPS;
constexpr int WS_TMA_WARP = WS_MMA_WARP + 1;   // issues the TMA loads when nothing is gathered
constexpr int WS_THREADS = (WS_TMA_WARP + 1) * 32;
constexpr int WS_PROD_THREADS = WS_PROD_WARPS * 32;
constexpr int WS_ROWS_PER_PASS = WS_PROD_THREADS / 8;   // rows covered by one pass of the producer threads
constexpr int WS_A_PASSES = TC_BM / WS_ROWS_PER_PASS;

template <int NCOLS, bool GELU>  // TMEM columns allocated = 2 accumulators of NCOLS/2 columns
__global__ void __launch_bounds__(WS_THREADS, NCOLS >= 256 ? 1 : 2) tc_igemm_ws_kernel(const TcGemmP p, const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmW) {
  extern __shared__ unsigned char dyn_smem[];
  __shared__ __align__(8) uint64_t full_bar[WS_STAGES], empty_bar[WS_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.BN;
  const uint32_t stage_bytes = TC_A_BYTES + (uint32_t)BN * (TC_BK * 2);
  const uint32_t smem0 = (smem_u32(dyn_smem) + 1023u) & ~1023u;
  const int KB = p.im2col ? p.KW * p.KW : (p.K + TC_BK - 1) / TC_BK;  // im2col: one k-block per filter tap
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = (p.M + TC_BM - 1) / TC_BM;
  const int num_tiles = tiles_m * tiles_n;
  const bool gather = p.conv && !p.im2col;  // A tiles gathered with cp.async by all producer threads

  if (warp == WS_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < WS_STAGES; ++s) { mbar_init(&full_bar[s], (gather ? WS_PROD_THREADS : 0) + 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    const int epi_threads = (WS_EPI_WARPS + (gather ? 0 : WS_PROD_WARPS)) * 32;
    mbar_init(&acc_empty[0], epi_threads); mbar_init(&acc_empty[1], epi_threads);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (gather ? warp < WS_PROD_WARPS : warp == WS_TMA_WARP) {
    // ------------------------------ producers ------------------------------------------------
    // W tiles (and A tiles of dense GEMMs / im2col convolutions) come by TMA from one elected thread. The fallback
    // implicit-GEMM gather of A is done by all producer threads with 16-byte cp.async; when nothing is gathered the
    // TMA thread lives in its own warp and the producer warps work as epilogue warps instead.
    const bool tma_thread = gather ? tid == 0 : lane == 0;
    const uint32_t tma_bytes = (uint32_t)BN * (TC_BK * 2) + (gather ? 0u : (uint32_t)TC_A_BYTES);
    if (!gather && !tma_thread) {
      // lanes 1..31 of the TMA warp: nothing to do
    } else {
    const int c = tid & 7, rbase = tid >> 3;  // 16-byte chunk, rows rbase + WS_ROWS_PER_PASS*i
    int it = 0;   // flat k-block counter over all tiles of this CTA
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * TC_BM, n0 = (tile % tiles_n) * BN;
      const eh_t* a_ptr[WS_A_PASSES];
      int a_ih0[WS_A_PASSES], a_iw0[WS_A_PASSES];
      bool a_ok[WS_A_PASSES];
      int bw = 0, bh = 0, bn = 0;  // im2col: base pixel of the tile's first output row
      if (p.im2col) {
        const int ow = m0 % p.OW, tq = m0 / p.OW;
        bw = ow * p.stride - p.pad_l;
        bh = (tq % p.OH) * p.stride - p.pad_t;
        bn = tq / p.OH;
      }
      if (gather) {
#pragma unroll
        for (int i = 0; i < WS_A_PASSES; ++i) {
          const int m = m0 + rbase + WS_ROWS_PER_PASS * i;
          a_ok[i] = m < p.M;
          a_ih0[i] = a_iw0[i] = 0;
          a_ptr[i] = p.A;
          if (a_ok[i]) {
            const int ow = m % p.OW;
            const int tq = m / p.OW;
            const int oh = tq % p.OH;
            const int n = tq / p.OH;
            a_ptr[i] = p.A + (size_t)n * p.H * p.Wd * p.Cin;
            a_ih0[i] = oh * p.stride - p.pad_t;
            a_iw0[i] = ow * p.stride - p.pad_l;
          }
        }
      }
      for (int kb = 0; kb < KB; ++kb, ++it) {
        const int stage = it % WS_STAGES;
        mbar_wait(&empty_bar[stage], (uint32_t)(((it / WS_STAGES) & 1) ^ 1));  // slot released by the MMA warp
        const uint32_t sa = smem0 + (uint32_t)stage * stage_bytes, sb = sa + TC_A_BYTES;
        if (tma_thread) {
          mbar_arrive_expect_tx(&full_bar[stage], tma_bytes);
          tma_load_2d(sb, &tmW, kb * TC_BK, n0, &full_bar[stage]);
          if (p.im2col) tma_load_im2col(sa, &tmA, 0, bw, bh, bn, (uint16_t)(kb % p.KW), (uint16_t)(kb / p.KW), &full_bar[stage]);
          else if (!p.conv) tma_load_2d(sa, &tmA, kb * TC_BK, m0, &full_bar[stage]);
        }
        if (gather) {
          const int k = kb * TC_BK + c * 8;
          const int tap = k / p.Cin;
          const int ci = k - tap * p.Cin;
          const int kh = tap / p.KW;
          const int kw = tap - kh * p.KW;
#pragma unroll
          for (int i = 0; i < WS_A_PASSES; ++i) {
            const int row = rbase + WS_ROWS_PER_PASS * i;
            const uint32_t dst = sa + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((c ^ (row & 7)) << 4);
            const eh_t* src = p.A;
            uint32_t bytes = 0;
            if (a_ok[i] && k < p.K) {
              const int ih = a_ih0[i] + kh, iw = a_iw0[i] + kw;
              if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.Wd) {
                src = a_ptr[i] + ((size_t)ih * p.Wd + iw) * p.Cin + ci;
                bytes = 16;
              }
            }
            cp_async16(dst, src, bytes);
          }
          cp_async_commit();
          if (it >= WS_LAG) {  // the group issued WS_LAG iterations ago has landed: publish that stage
            cp_async_wait<WS_LAG>();
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            mbar_arrive(&full_bar[(it - WS_LAG) % WS_STAGES]);
          }
        }
      }
    }
    if (gather) {  // drain: publish the last WS_LAG stages
      cp_async_wait<0>();
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      for (int j = (it >= WS_LAG ? it - WS_LAG : 0); j < it; ++j) mbar_arrive(&full_bar[j % WS_STAGES]);
    }
    }
  } else if (warp == WS_MMA_WARP) {
    // ------------------------------ MMA issuer ------------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BN);
      int it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
        const int a = ti & 1;
        mbar_wait(&acc_empty[a], (uint32_t)(((ti >> 1) & 1) ^ 1));  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem + (uint32_t)(a * (NCOLS / 2));
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int stage = it % WS_STAGES;
          mbar_wait(&full_bar[stage], (uint32_t)((it / WS_STAGES) & 1));
          tc_fence_after();
          const uint32_t sa = smem0 + (uint32_t)stage * stage_bytes;
          const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + TC_A_BYTES);
#pragma unroll
          for (int j = 0; j < TC_BK / 16; ++j)
            umma_f16(tacc, adesc + (uint64_t)(2 * j), bdesc + (uint64_t)(2 * j), idesc, (kb | j) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);           // frees the smem slot when these MMAs retire
          if (kb == KB - 1) umma_commit(&acc_full[a]);  // accumulator complete
        }
      }
    }
  } else if (warp < WS_MMA_WARP) {
    // ------------------------------ epilogue warps ----------------------------------------------
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int nsub = (WS_EPI_WARPS + (gather ? 0 : WS_PROD_WARPS)) / 4;    // warps sharing a quarter's column chunks
    const int sub = (gather ? warp - WS_PROD_WARPS : warp) >> 2;
    int ti = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
      const int a = ti & 1;
      const int m0 = (tile / tiles_n) * TC_BM, n0 = (tile % tiles_n) * BN;
      mbar_wait(&acc_full[a], (uint32_t)((ti >> 1) & 1));
      tc_fence_after();
      const int m = m0 + q * 32 + lane;
      const bool m_ok = m < p.M;
      const uint32_t tacc = tmem + (uint32_t)(a * (NCOLS / 2)) + ((uint32_t)(q * 32) << 16);
      for (int cc = sub; cc * 16 < BN; cc += nsub) {
        uint32_t r[16];
        tmem_ld16(tacc + (uint32_t)(cc * 16), r);
        const int nb = n0 + cc * 16;
        if (m_ok && nb < p.N) epilogue_store16<GELU>(p, r, m, nb);
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[a]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WS_MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(NCOLS) : "memory");
  }
}

static int tc_pick_bn(int N) {
  if (N <= 256) return (N + 15) / 16 * 16;
  // widest tile that divides N evenly into 16-column multiples, else 256 with a ragged last tile
  for (int bn = 256; bn >= 128; bn -= 16)
    if (N % bn == 0) return bn;
  return 256;
}

template <int NCOLS>
static int tc_launch(const TcGemmP& p, cudaStream_t st) {
  const size_t smem = (size_t)p.stages * (TC_A_BYTES + (size_t)p.BN * TC_BK * 2) + 1024;
  static SmemOptIn opt;
  {
    cudaError_t e = opt.ensure(tc_igemm_kernel<NCOLS>, smem, true);
    if (e != cudaSuccess) return (int)e;
  }
  dim3 grid((p.M + TC_BM - 1) / TC_BM, (p.N + p.BN - 1) / p.BN);
  tc_igemm_kernel<NCOLS><<<grid, 256, smem, st>>>(p);
  return 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency).
static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// bf16 row-major [rows, cols] matrix (row pitch ld elements), box = box_rows x 64 columns, 128-byte swizzle,
// out-of-bounds elements read as zero (handles the K tail and ragged M / N edges).
static int make_tmap_2d(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = tensor_map_encoder();
  if (!enc) return (int)cudaErrorNotSupported;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, (FRX_ENC_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

static PFN_cuTensorMapEncodeIm2col_v12000 im2col_map_encoder() {
  static PFN_cuTensorMapEncodeIm2col_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(ptr);
  }
  return fn;
}

// NHWC bf16 activation [B, H, W, C] as an im2col tensor map: 128 output pixels x 64 channels per load, the filter
// tap is given per instruction.  Base-pixel bounding box per CUTLASS' convention (W, H order):
// lower = -pad_before, upper = pad_after - (k - 1); traversal stride = convolution stride.
static int make_tmap_im2col(CUtensorMap* tm, const TcGemmP& p) {
  PFN_cuTensorMapEncodeIm2col_v12000 enc = im2col_map_encoder();
  if (!enc) return (int)cudaErrorNotSupported;
  const int B = p.M / (p.OH * p.OW), k = p.KW;
  const int pad_r = (p.OW - 1) * p.stride + k - p.Wd - p.pad_l, pad_b = (p.OH - 1) * p.stride + k - p.H - p.pad_t;
  cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Wd, (cuuint64_t)p.H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)p.Cin * 2, (cuuint64_t)p.Wd * p.Cin * 2, (cuuint64_t)p.H * p.Wd * p.Cin * 2};
  int lower[2] = {-p.pad_l, -p.pad_t};
  int upper[2] = {pad_r - (k - 1), pad_b - (k - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)p.stride, (cuuint32_t)p.stride, 1};
  CUresult r = enc(tm, (FRX_ENC_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 4, const_cast<eh_t*>(p.A), dims, strides, lower, upper,
                   (cuuint32_t)TC_BK, (cuuint32_t)TC_BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

template <int NCOLS>
static int tc_launch_ws(const TcGemmP& p_in, int num_sms, cudaStream_t st) {
  TcGemmP p = p_in;
  const size_t smem = (size_t)WS_STAGES * (TC_A_BYTES + (size_t)p.BN * TC_BK * 2) + 1024;
  static SmemOptIn opt, opt_gelu;
  {
    cudaError_t e = opt.ensure(tc_igemm_ws_kernel<NCOLS, false>, smem, true);
    if (e == cudaSuccess) e = opt_gelu.ensure(tc_igemm_ws_kernel<NCOLS, true>, smem, true);
    if (e != cudaSuccess) return (int)e;
  }
  const int tiles = ((p.M + TC_BM - 1) / TC_BM) * ((p.N + p.BN - 1) / p.BN);
  int per_sm = (int)((227 * 1024) / (smem + 2048));     // shared memory
  if (per_sm > 512 / NCOLS) per_sm = 512 / NCOLS;       // tensor memory
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  const int grid = tiles < num_sms * per_sm ? tiles : num_sms * per_sm;
  CUtensorMap tmA, tmW;
  int rc;
  p.im2col = 0;
  if (p.conv && p.Wpad && p.Cin <= TC_BK && p.M % (p.OH * p.OW) == 0 && make_tmap_im2col(&tmA, p) == 0) {
    p.im2col = 1;  // conv A tiles by TMA im2col; weights in the tap-major, 64-channel-padded layout
    rc = make_tmap_2d(&tmW, p.Wpad, p.N, (long long)p.KW * p.KW * TC_BK, (long long)p.KW * p.KW * TC_BK, p.BN);
  } else {
    rc = make_tmap_2d(&tmW, p.W, p.N, p.K, p.ldw, p.BN);
    if (rc) return rc;
    if (!p.conv) rc = make_tmap_2d(&tmA, p.A, p.M, p.K, p.lda, TC_BM);
    else tmA = tmW;
  }
  if (rc) return rc;
  if (getenv("FRX_DEBUG"))
    fprintf(stderr, "[frx] tc gemm M=%d N=%d K=%d BN=%d conv=%d Cin=%d stride=%d -> %s, grid %d\n", p.M, p.N, p.K, p.BN, p.conv,
            p.Cin, p.stride, p.im2col ? "TMA im2col" : (p.conv ? "cp.async gather" : "TMA dense"), grid);
  if (p.act == ACT_GELU) tc_igemm_ws_kernel<NCOLS, true><<<grid, WS_THREADS, smem, st>>>(p, tmA, tmW);
  else tc_igemm_ws_kernel<NCOLS, false><<<grid, WS_THREADS, smem, st>>>(p, tmA, tmW);
  return 0;
}

// Persistent warp-specialised kernel; NCOLS = TMEM columns for the two accumulators.
int launch_tc_igemm_ws(TcGemmP p, int num_sms, cudaStream_t st) {
  if (p.BN == 0) {
    p.BN = tc_pick_bn(p.N);
    // few M tiles (the 4x8 maps of the last trunk stage, the encoder tokens): narrower N tiles until every SM has one
    const int tiles_m = (p.M + TC_BM - 1) / TC_BM;
    while (p.BN > 64 && p.BN % 32 == 0 && p.N % (p.BN / 2) == 0 && tiles_m * ((p.N + p.BN - 1) / p.BN) < num_sms) p.BN /= 2;
  }
  if (p.BN <= 32) { p.BN = 32; return tc_launch_ws<64>(p, num_sms, st); }
  if (p.BN <= 64) return tc_launch_ws<128>(p, num_sms, st);
  if (p.BN <= 128) return tc_launch_ws<256>(p, num_sms, st);
  return tc_launch_ws<512>(p, num_sms, st);
}

int launch_tc_igemm(TcGemmP p, cudaStream_t st) {
  if (p.act == ACT_GELU) {  // only the persistent kernel carries the GELU epilogue
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return launch_tc_igemm_ws(p, sms, st);
  }
  if (p.BN == 0) p.BN = tc_pick_bn(p.N);
  {  // ring depth: 3 when three stages still leave room for >= 2 (3 for narrow tiles) CTAs per SM, else 2
    const int kb = (p.K + TC_BK - 1) / TC_BK;
    const size_t stage = TC_A_BYTES + (size_t)p.BN * TC_BK * 2;
    p.stages = (kb > 3 && 3 * stage <= 72 * 1024) ? 3 : 2;
  }
  if (p.BN <= 32) { p.BN = 32; return tc_launch<32>(p, st); }
  if (p.BN <= 64) return tc_launch<64>(p, st);
  if (p.BN <= 128) return tc_launch<128>(p, st);
  return tc_launch<256>(p, st);
}

// ===========================================================================
// small helpers of the 16-bit path (eh_t = the encoder operand type, common.cuh)
// ===========================================================================
__global__ void __launch_bounds__(256) f32_to_h16_kernel(const float* __restrict__ in, eh_t* __restrict__ out, long long n) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v = __ldg(reinterpret_cast<const float4*>(in + i));
    *reinterpret_cast<uint2*>(out + i) = make_uint2(eh2_pack(v.x, v.y), eh2_pack(v.z, v.w));
  } else {
    for (; i < n; ++i) out[i] = eh_from_float(in[i]);
  }
}

void launch_f32_to_h16(const float* in, eh_t* out, long long n, cudaStream_t st) {
  f32_to_h16_kernel<<<(unsigned)((n / 4 + 256) / 256), 256, 0, st>>>(in, out, n);
}


__global__ void __launch_bounds__(256) h16_to_f32_kernel(const eh_t* __restrict__ in, float* __restrict__ out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = eh_to_float(in[i]);
}
void launch_h16_to_f32(const eh_t* in, float* out, long long n, cudaStream_t st) {
  h16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n);
}


// [N][taps][Cin] -> [N][taps][64] (zero padded) bf16, for the im2col path
__global__ void __launch_bounds__(256) pad_conv_weights_kernel(const eh_t* __restrict__ in, eh_t* __restrict__ out,
                                                               long long rows, int Cin) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 64) return;
  const int ci = (int)(i & 63);
  out[i] = ci < Cin ? in[(i >> 6) * Cin + ci] : eh_from_float(0.f);
}
void launch_pad_conv_weights(const eh_t* in, eh_t* out, long long rows, int Cin, cudaStream_t st) {
  pad_conv_weights_kernel<<<(unsigned)((rows * 64 + 255) / 256), 256, 0, st>>>(in, out, rows, Cin);
}

}  // namespace frx
