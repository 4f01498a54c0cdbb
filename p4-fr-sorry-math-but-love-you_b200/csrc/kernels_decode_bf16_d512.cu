// SwinTRN decoder geometry of the persistent decode kernel (networks/SWIN.py:922-1021: hidden 512, 8 heads x 64, filter 512,
// 4 layers, 144 memory tokens): clusters of 8 CTAs x 8 warps, one head per CTA, one CTA per SM (the 64-wide heads need ~200
// registers per thread).  Same source as kernels_decode_bf16.cu; the q|k|v and cache-row stages (48 / 32 weight fragments
// per lane) take the streaming register ring of gemm2.
#define FRX_DEC_D 512
#define FRX_DEC_FF 512
#define FRX_DEC_HD 64
#define FRX_DEC_MINBLOCKS 1
#define FRX_DEC_NAME(x) x##_d512
#define FRX_DEC_VARIANT 1
#include "kernels_decode_bf16.cu"
