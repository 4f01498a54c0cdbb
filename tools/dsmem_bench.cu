// dsmem_bench.cu -- micro-benchmark of the cluster all-gather primitives considered for the decode kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dsmem_bench tools/dsmem_bench.cu
// Every CTA of an 8-CTA cluster sends SLICE bytes to all 8 CTAs (itself included) and waits until the
// 8*SLICE bytes it expects have landed; reports cycles per exchange (average over REPS, CTA 0 thread 0).
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (int spins = 0; spins < (1 << 24); ++spins) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}

constexpr int CL = 8;
// mode 0: st.async b32, 1: v2, 2: v4, 3: cp.async.bulk (one per destination), 4: v4 issued as dest-per-warp
template <int MODE>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(256, 2) bench(int slice, int reps, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  // layout: [2 barriers (16 B)] pad to 128 | staging [slice] | recv [2][8*slice]
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm);
  unsigned char* stg = sm + 128;
  unsigned char* recv = stg + ((slice + 127) / 128) * 128;
  cg::cluster_group cl = cg::this_cluster();
  const int r = cl.block_rank(), tid = threadIdx.x;
  const uint32_t b0 = smem_u32(&bar[0]), b1 = smem_u32(&bar[1]);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(b0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(b1));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int i = tid; i < slice / 4; i += 256) reinterpret_cast<uint32_t*>(stg)[i] = i * 7 + r;
  __syncthreads();
  cl.sync();
  long long t0 = clock64();
  for (int it = 0; it < reps; ++it) {
    const uint32_t sb = (it & 1) ? b1 : b0;
    unsigned char* dstbuf = recv + (size_t)(it & 1) * 8 * slice + (size_t)r * slice;
    if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(sb), "r"(8 * slice) : "memory");
    if (MODE == 3) {
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncthreads();
      if (tid < CL) {
        const uint32_t dst = mapa_u32(smem_u32(dstbuf), tid), rb = mapa_u32(sb, tid);
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     ::"r"(dst), "r"(smem_u32(stg)), "r"(slice), "r"(rb) : "memory");
      }
    } else if (MODE == 4) {
      // warp w sends the whole slice to destination w with 16-byte stores
      const int w = tid >> 5, lane = tid & 31;
      const uint32_t rb = mapa_u32(sb, w);
      for (int c = lane; c < slice / 16; c += 32) {
        const uint4 v = reinterpret_cast<const uint4*>(stg)[c];
        const uint32_t dst = mapa_u32(smem_u32(dstbuf + c * 16), w);
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n"
                     ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rb) : "memory");
      }
    } else {
      constexpr int W = MODE == 0 ? 4 : MODE == 1 ? 8 : 16;
      for (int c = tid; c < slice / W; c += 256) {
        const uint32_t la = smem_u32(dstbuf + c * W);
        if (MODE == 0) {
          const uint32_t v = reinterpret_cast<const uint32_t*>(stg)[c];
#pragma unroll
          for (int d = 0; d < CL; ++d)
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];\n"
                         ::"r"(mapa_u32(la, d)), "r"(v), "r"(mapa_u32(sb, d)) : "memory");
        } else if (MODE == 1) {
          const uint2 v = reinterpret_cast<const uint2*>(stg)[c];
#pragma unroll
          for (int d = 0; d < CL; ++d)
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];\n"
                         ::"r"(mapa_u32(la, d)), "r"(v.x), "r"(v.y), "r"(mapa_u32(sb, d)) : "memory");
        } else {
          const uint4 v = reinterpret_cast<const uint4*>(stg)[c];
#pragma unroll
          for (int d = 0; d < CL; ++d)
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];\n"
                         ::"r"(mapa_u32(la, d)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mapa_u32(sb, d)) : "memory");
        }
      }
    }
    mbar_wait(sb, (it >> 1) & 1);
    __syncthreads();
  }
  long long t1 = clock64();
  cl.sync();
  if (blockIdx.x == 0 && tid == 0) out[0] = (t1 - t0) / reps;
  if (blockIdx.x == 0 && tid == 1) out[1] = reinterpret_cast<uint32_t*>(recv)[slice / 4 + 1];  // keep the data live
}

template <int MODE>
static void run(int slice, int clusters, long long* dout) {
  size_t smem = 128 + ((slice + 127) / 128) * 128 + 2 * 8 * (size_t)slice;
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bench<MODE><<<clusters * CL, 256, smem>>>(slice, 2000, dout);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  printf("mode %d slice %5d B clusters %2d : %6lld cycles/exchange  (%.1f B/cycle sent per CTA) %s\n", MODE, slice, clusters,
         h[0], 8.0 * slice / (double)h[0], e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 16);
  const int slices[] = {512, 1024, 2048, 4096};
  for (int clusters : {1, 32}) {
    for (int s : slices) {
      run<0>(s, clusters, dout);
      run<1>(s, clusters, dout);
      run<2>(s, clusters, dout);
      run<4>(s, clusters, dout);
      run<3>(s, clusters, dout);
    }
  }
  return 0;
}
