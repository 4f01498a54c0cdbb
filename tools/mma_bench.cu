// mma_bench.cu -- latency / issue rate of the legacy warp-level mma.sync.m16n8k16 (bf16 -> fp32) on sm_100a, the
// instruction the decode kernel's attention and linear stages run on.  One CTA per SM, W warps, every warp runs CH
// independent accumulation chains of ITER dependent MMAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench.bin tools/mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int CH>
__global__ void k(int iters, long long* out, float* sink) {
  float c[CH][4];
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b0 = 0x3f803f80u + threadIdx.x, b1 = 0x3f803f80u;
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.f) sink[0] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}

template <int CH>
void run(int warps, long long* d, float* sink) {
  const int iters = 4096;
  k<CH><<<148, warps * 32>>>(iters, d, sink);
  k<CH><<<148, warps * 32>>>(iters, d, sink);
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / iters;   // cycles per iteration of CH MMAs per warp
  printf("warps/SM %2d  chains/warp %d : %.1f cycles per MMA per warp-chain step, %.2f cycles per MMA per SM sub-partition, %.0f dense TFLOP/s at 1.965 GHz\n",
         warps, CH, per, per / CH / ((warps + 3) / 4), 148.0 * warps * CH * 4096 * 2 / per * 1.965e9 / 1e12);
}

int main() {
  long long* d; float* sink;
  cudaMalloc(&d, 8); cudaMalloc(&sink, 4);
  for (int w : {1, 4, 8, 16, 32}) { run<1>(w, d, sink); run<2>(w, d, sink); run<4>(w, d, sink); run<8>(w, d, sink); }
  return 0;
}
