// Training-forward instantiations of the fused field kernel: every epilogue also writes the tape
// (operand chunks + ReLU sign words, program.h TapeLayout) that b200r_field_bwd and the weight-gradient kernel read.
#define B200R_SAVE true
#include "field_fwd_kernel.cuh"
