// swapab_bench.cu -- micro-benchmark: the decoder step's linear stages as "swap-AB" tcgen05 GEMMs.
//
// Question: the persistent decode kernel spends ~49 k of its ~118 k cycles per step in 19 small linear
// stages whose weights (1.57 MB per CTA and step) are loaded L2 -> registers by every warp and fed to mma.sync with only 8
// of 16 MMA rows used.  Alternative: let the WEIGHTS be the M operand of tcgen05.mma (M = 128 output features per
// instruction), the 8 images the N operand (N = 16, rows 8..15 zero), stream the weights L2 -> shared memory with TMA
// through a deep ring that runs ahead across stage boundaries (the weight sequence of a step is static), accumulate in
// TMEM.  This program runs that chain for one CTA per SM with the real per-stage shapes and measures cycles per step
//   (a) back to back (the floor of the linear stages: TMA streaming vs the L2 throughput cap), and
//   (b) with a spin between stages that stands for attention / LayerNorm / exchanges, to see whether the ring hides the
//       weight stream behind them (stage latency = MMA issue + TMEM read only);
// and validates the first tile against a CPU reference (descriptor / swizzle correctness).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/swapab_bench tools/swapab_bench.cu
//   tools/swapab_bench [ctas=128] [delay_cycles=0] [ring=6] [steps=40]
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int MAXRING = 10;
constexpr int SLAB_BYTES = 128 * 64 * 2;   // 128 rows x 64 K bf16
constexpr int NB = 16;                      // N of the MMA: 8 images + 8 zero rows
constexpr int KMAX = 1024;
constexpr int NACC = 8;                     // independent accumulators per tile (k-step % NACC): consecutive tcgen05.mma into ONE accumulator serialise on its ~200-cycle latency

struct Job { int rows, kslabs, first_row, stage_end, stage_begin; };   // one accumulator tile
constexpr int MAXJOBS = 64;
struct Tape { int njobs; Job job[MAXJOBS]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  for (long long spins = 0; spins < (1ll << 28); ++spins) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {   // K-major SWIZZLE_128B, 8-row groups of 1024 B
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// activations [8 images x K] -> B operand tile in shared memory: K-major SWIZZLE_128B, per 64-wide K slab 16 rows x 128 B
// (rows 8..15 zero): element (row r, k) of slab s at  s*2048 + r*128 + (((k%64)/8) ^ (r&7))*16 + (k%8)*2
__device__ __forceinline__ uint32_t b_off(int r, int k) { return (uint32_t)((k >> 6) * 2048 + r * 128 + ((((k & 63) >> 3) ^ (r & 7)) << 4) + (k & 7) * 2); }

__global__ void __launch_bounds__(192, 1) swapab_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm64,
                                                       const Tape tape, int rows_per_rank, int steps, int ring, int delay,
                                                       const __nv_bfloat16* __restrict__ act, float* __restrict__ dbg, long long* __restrict__ cyc) {
  extern __shared__ unsigned char dyn[];
  __shared__ __align__(8) uint64_t full_bar[MAXRING], empty_bar[MAXRING], acc_full[2], acc_empty[2], act_ready;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem0 = (smem_u32(dyn) + 1023u) & ~1023u;
  const uint32_t act_s = smem0 + (uint32_t)ring * SLAB_BYTES;           // B operand: KMAX/64 slabs x 2 KB
  unsigned char* act_g = dyn + (smem0 - smem_u32(dyn)) + (size_t)ring * SLAB_BYTES;
  if (tid == 0) {
    for (int i = 0; i < ring; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], 4); mbar_init(&acc_empty[1], 4);   // one elected lane per epilogue warp
    mbar_init(&act_ready, 4);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // B operand: the 8 activation rows (same for every stage here), rows 8..15 zero
  for (int i = tid; i < NB * KMAX; i += blockDim.x) {
    const int r = i / KMAX, k = i % KMAX;
    *reinterpret_cast<__nv_bfloat16*>(act_g + b_off(r, k)) = r < 8 ? act[r * KMAX + k] : __float2bfloat16_rn(0.f);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int rank = blockIdx.x & 3;                  // CTA rank in its cluster: which quarter of every layer's columns
  const int row0 = rank * rows_per_rank;
  const long long t_begin = clock64();

  if (warp == 0) {
    if (lane == 0) {                                // ---- TMA producer: the step's weight tape, as far ahead as the ring allows
      int it = 0;
      for (int s = 0; s < steps; ++s)
        for (int j = 0; j < tape.njobs; ++j) {
          const Job jb = tape.job[j];
          for (int k = 0; k < jb.kslabs; ++k, ++it) {
            const int slot = it % ring;
            mbar_wait(&empty_bar[slot], (uint32_t)(((it / ring) & 1) ^ 1));
            mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)jb.rows * 128u);
            tma_load_2d(smem0 + (uint32_t)slot * SLAB_BYTES, jb.rows == 128 ? &tm128 : &tm64, 0, row0 + jb.first_row + k * jb.rows, &full_bar[slot]);
          }
        }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                // ---- MMA issuer
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int it = 0, ti = 0, st = 0;
      for (int s = 0; s < steps; ++s)
        for (int j = 0; j < tape.njobs; ++j, ++ti) {
          const Job jb = tape.job[j];
          if (jb.stage_begin) { mbar_wait(&act_ready, (uint32_t)(st & 1)); ++st; }   // the stage's input rows are in shared memory
          const int a = ti & 1;
          mbar_wait(&acc_empty[a], (uint32_t)(((ti >> 1) & 1) ^ 1));
          tc_fence_after();
          const uint32_t tacc = tmem + (uint32_t)(a * 256);
          for (int k = 0; k < jb.kslabs; ++k, ++it) {
            const int slot = it % ring;
            mbar_wait(&full_bar[slot], (uint32_t)((it / ring) & 1));
            tc_fence_after();
            const uint64_t adesc = make_smem_desc(smem0 + (uint32_t)slot * SLAB_BYTES), bdesc = make_smem_desc(act_s + (uint32_t)k * 2048u);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int ks = k * 4 + q;                       // k-step; accumulator ks % NACC, first touch overwrites
              umma(tacc + (uint32_t)((ks % NACC) * 32), adesc + (uint64_t)(2 * q), bdesc + (uint64_t)(2 * q), idesc, ks >= NACC ? 1u : 0u);
            }
            umma_commit(&empty_bar[slot]);
            if (k == jb.kslabs - 1) umma_commit(&acc_full[a]);
          }
        }
    }
  } else {                                           // ---- epilogue warps 2..5: TMEM lane quarter (warp & 3)
    const int q = warp & 3;
    int ti = 0, st = 0;
    // the first stage's input is ready
    if (lane == 0) mbar_arrive(&act_ready);
    for (int s = 0; s < steps; ++s)
      for (int j = 0; j < tape.njobs; ++j, ++ti) {
        const Job jb = tape.job[j];
        const int a = ti & 1;
        mbar_wait(&acc_full[a], (uint32_t)((ti >> 1) & 1));
        tc_fence_after();
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        const int used = jb.kslabs * 4 < NACC ? jb.kslabs * 4 : NACC;
        for (int u = 0; u < used; ++u) {
          uint32_t r[8];
          tmem_ld8(tmem + (uint32_t)(a * 256 + u * 32) + ((uint32_t)(q * 32) << 16), r);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += __uint_as_float(r[i]);
        }
        if (s == 0 && j == 0 && blockIdx.x == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) dbg[(q * 32 + lane) * 8 + i] = acc[i];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[a]);
        if (jb.stage_end) {                          // everything else of the step between two linear stages
          if (delay > 0) { const long long t0 = clock64(); while (clock64() - t0 < delay) {} }
          // (the real kernel writes the next stage's input rows here) -> visible to the tensor core, then signal
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
          const bool last = (s == steps - 1) && (j == tape.njobs - 1);
          __syncwarp();
          if (!last && lane == 0) mbar_arrive(&act_ready);
          ++st;
        }
      }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) cyc[blockIdx.x] = clock64() - t_begin;
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int ctas = argc > 1 ? atoi(argv[1]) : 128, delay = argc > 2 ? atoi(argv[2]) : 0, ring = argc > 3 ? atoi(argv[3]) : 6, steps = argc > 4 ? atoi(argv[4]) : 40;
  if (ring > MAXRING) { printf("ring <= %d\n", MAXRING); return 1; }
  // one decoder step of EfficientSATRN for one CTA of a 4-CTA cluster (two heads per CTA): per layer
  //   A q|k|v 192 columns (128 + 64), B/C/D 64 columns, cache rows 128 columns, E ffn0 256 columns (K = 256), F ffn1 64 columns (K = 1024)
  Tape tape{};
  int row = 0;
  auto add = [&](int rows, int ks, int begin, int end) { Job j{rows, ks, row, end, begin}; tape.job[tape.njobs++] = j; row += rows * ks; };
  for (int l = 0; l < 3; ++l) {
    add(128, 4, 1, 0); add(64, 4, 0, 1);          // A
    add(64, 4, 1, 1);                              // B
    add(64, 4, 1, 1);                              // C
    add(64, 4, 1, 1);                              // D
    add(128, 4, 1, 0); add(128, 4, 0, 1);          // E
    add(64, 16, 1, 1);                             // F
    add(128, 4, 1, 1);                             // cache rows
  }
  add(64, 4, 1, 1);                                // G
  const int mode = argc > 5 ? atoi(argv[5]) : 0;
  if (mode == 1) { tape.njobs = 0; row = 0; for (int i = 0; i < 28; ++i) add(128, 1, 1, 1); }        // 28 one-slab stages: cost per stage
  if (mode == 2) { tape.njobs = 0; row = 0; add(128, 56, 1, 1); }                                  // one 56-slab tile: cost per slab
  if (mode == 3) { tape.njobs = 0; row = 0; for (int i = 0; i < 28; ++i) add(128, 1, i == 0, i == 27); }  // 28 one-slab tiles in ONE stage
  const int rows_per_rank = row;
  const size_t total_rows = (size_t)rows_per_rank * 4;
  printf("tape: %d tiles, %d stages-with-delay, %.1f KB of weights per CTA and step\n", tape.njobs, 25, rows_per_rank * 128 / 1024.0);
  std::vector<__nv_bfloat16> hw(total_rows * 64), hact(8 * KMAX);
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
  for (auto& v : hw) v = __float2bfloat16(rnd());
  for (auto& v : hact) v = __float2bfloat16(rnd());
  __nv_bfloat16 *dw, *dact;
  float* ddbg;
  long long* dcyc;
  CK(cudaMalloc(&dw, hw.size() * 2)); CK(cudaMalloc(&dact, hact.size() * 2)); CK(cudaMalloc(&ddbg, 128 * 8 * 4)); CK(cudaMalloc(&dcyc, ctas * 8));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dact, hact.data(), hact.size() * 2, cudaMemcpyHostToDevice));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  EncodeTiled enc = (EncodeTiled)fn;
  CUtensorMap tm128, tm64;
  cuuint64_t dims[2] = {64, (cuuint64_t)total_rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t estr[2] = {1, 1};
  for (int v = 0; v < 2; ++v) {
    cuuint32_t box[2] = {64, (cuuint32_t)(v ? 64 : 128)};
    CUresult r = enc(v ? &tm64 : &tm128, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dw, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("tensor map %d failed: %d\n", v, (int)r); return 1; }
  }
  const size_t smem = (size_t)ring * SLAB_BYTES + (KMAX / 64) * 2048 + 2048;
  CK(cudaFuncSetAttribute(swapab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0));
    swapab_kernel<<<ctas, 192, smem>>>(tm128, tm64, tape, rows_per_rank, steps, ring, delay, dact, ddbg, dcyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> hc(ctas);
  std::vector<float> hd(128 * 8);
  CK(cudaMemcpy(hc.data(), dcyc, ctas * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hd.data(), ddbg, hd.size() * 4, cudaMemcpyDeviceToHost));
  long long mx = 0, sum = 0;
  for (long long c : hc) { mx = c > mx ? c : mx; sum += c; }
  // validate tile 0 of CTA 0 (rank 0): out[f][i] = sum_k W[f][k] x[i][k], K = 256 as 4 slabs of [128 rows x 64]
  double worst = 0;
  for (int f = 0; f < 128; ++f)
    for (int i = 0; i < 8; ++i) {
      double acc = 0;
      for (int k = 0; k < 256; ++k) acc += (double)__bfloat162float(hw[((size_t)(k / 64) * 128 + f) * 64 + (k % 64)]) * (double)__bfloat162float(hact[i * KMAX + k]);
      worst = fmax(worst, fabs(acc - hd[f * 8 + i]));
    }
  printf("ctas %d ring %d (%d KB) delay %d: %.1f us per step (kernel %.3f ms / %d steps), cycles per step: mean %.0f, max %.0f; tile-0 max |err| %.2e %s\n",
         ctas, ring, ring * 16, delay, ms * 1e3 / steps, ms, steps, (double)sum / ctas / steps, (double)mx / steps, worst, worst < 2e-3 ? "OK" : "MISMATCH");
  return 0;
}
