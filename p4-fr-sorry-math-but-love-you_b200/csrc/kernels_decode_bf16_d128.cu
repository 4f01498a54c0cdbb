// LiteSATRN geometry of the persistent decode kernel (hidden 128, 4 heads x 32, filter 512 -> clusters of 4 CTAs):
// the same source as kernels_decode_bf16.cu, compiled a second time with these constants.
#define FRX_DEC_D 128
#define FRX_DEC_FF 512
#define FRX_DEC_NAME(x) x##_d128
#define FRX_DEC_VARIANT 1
#include "kernels_decode_bf16.cu"
