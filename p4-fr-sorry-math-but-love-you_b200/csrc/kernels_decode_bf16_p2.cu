// EfficientSATRN geometry with TWO heads per CTA: clusters of 4 CTAs x 16 warps, so that the 32 clusters of a
// 256-image batch are 128 CTAs, each alone on its SM (the one-head version needs 256 CTAs, two per SM, and two CTAs
// sharing an SM's L1/TEX pipe step in 67 us instead of 42).  Same source as kernels_decode_bf16.cu.
#define FRX_DEC_D 256
#define FRX_DEC_FF 1024
#define FRX_DEC_HPC 2
#define FRX_DEC_KVDEPTH 2     // 207 KB of shared memory per CTA (one CTA per SM): two K/V blocks in flight per warp
#define FRX_DEC_NAME(x) x##_p2
#define FRX_DEC_VARIANT 1
#include "kernels_decode_bf16.cu"
