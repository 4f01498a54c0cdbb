// phase_bench.cu -- micro-benchmark behind the grid-phase decode design: what does one grid-wide dependent phase
// cost on a B200?  A persistent co-resident grid runs REPS phases; every phase each CTA (a) waits on a grid barrier,
// (b) pulls an A tile (rows written by OTHER CTAs in the previous phase) from L2 with cp.async.cg, (c) does a token
// amount of arithmetic, (d) writes its slice of the next buffer, (e) arrives on the barrier.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/phase_bench tools/phase_bench.cu
// Run:   tools/phase_bench [ctas_per_sm] [threads] [tile_bytes] [reps]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(p) : "memory");
}
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    red_release(ctr);
    int spins = 0;
    while (ld_acquire(ctr) < target) {
      if (++spins > (1 << 22)) __trap();
    }
  }
  __syncthreads();
}

// mode 0: barrier only; 1: barrier + A-tile load (cp.async.cg) + store of a slice; 2: same with plain ld.global.cg
__global__ void bench(int mode, int reps, int tile_bytes, unsigned* ctr, float* buf0, float* buf1, long long* out, int total_rows_bytes) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int tid = threadIdx.x, G = gridDim.x;
  float* bufs[2] = {buf0, buf1};
  long long t0 = clock64();
  float acc = 0.f;
  for (int it = 0; it < reps; ++it) {
    const float* src = bufs[it & 1];
    float* dst = bufs[(it & 1) ^ 1];
    if (mode >= 1) {
      // every CTA reads a tile starting at a CTA-dependent offset (wraps), i.e. data written by other CTAs
      const size_t off = ((size_t)blockIdx.x * 4096) % (size_t)total_rows_bytes;
      const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
      for (int i = tid * 16; i < tile_bytes; i += blockDim.x * 16) {
        const char* g = reinterpret_cast<const char*>(src) + (off + i) % (size_t)total_rows_bytes;
        if (mode == 1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sbase + i), "l"(g) : "memory");
        else {
          float4 v;
          asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(g));
          *reinterpret_cast<float4*>(sm + i) = v;
        }
      }
      if (mode == 1) asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
      __syncthreads();
      // token arithmetic + slice store: CTA b writes total/G bytes
      const int slice = total_rows_bytes / G / 4;  // floats
      for (int i = tid; i < slice; i += blockDim.x) {
        float v = reinterpret_cast<float*>(sm)[i % (tile_bytes / 4)] * 1.0001f + 1.f;
        acc += v;
        dst[(size_t)blockIdx.x * slice + i] = v;
      }
    }
    grid_barrier(ctr, (unsigned)(it + 1) * G);
  }
  long long t1 = clock64();
  if (blockIdx.x == 0 && tid == 0) { out[0] = t1 - t0; out[1] = (long long)acc; }
}

int main(int argc, char** argv) {
  int cps = argc > 1 ? atoi(argv[1]) : 1, threads = argc > 2 ? atoi(argv[2]) : 512;
  int tile = argc > 3 ? atoi(argv[3]) : 32768, reps = argc > 4 ? atoi(argv[4]) : 2000;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int G = prop.multiProcessorCount * cps;
  const int total = 256 * 1024;  // one [256 x 256] fp32 activation
  unsigned* ctr; float *b0, *b1; long long* out;
  cudaMalloc(&ctr, 4); cudaMalloc(&b0, total); cudaMalloc(&b1, total); cudaMalloc(&out, 16);
  cudaMemset(b0, 0, total); cudaMemset(b1, 0, total);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, tile);
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(ctr, 0, 4);
    void* args[] = {&mode, &reps, &tile, &ctr, &b0, &b1, &out, (void*)&total};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchCooperativeKernel((void*)bench, dim3(G), dim3(threads), args, tile, 0);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("grid %d x %d thr, tile %d B, mode %d: %s/%s  %.1f cycles/phase, %.3f us/phase\n", G, threads, tile, mode,
           cudaGetErrorString(e), cudaGetErrorString(e2), (double)h[0] / reps, ms * 1e3 / reps);
  }
  return 0;
}
