// Inference / no-grad instantiations of the fused field kernel (field_fwd_kernel.cuh).
#define B200R_SAVE false
#include "field_fwd_kernel.cuh"
