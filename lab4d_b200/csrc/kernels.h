// Internal launch interfaces between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>

#include "program.h"

namespace b200r {

constexpr int kMaxPackSlices = 80;

struct PackParams {
  PackSlice slices[kMaxPackSlices];
  const float* weights[B200R_MAX_LAYERS];  // device pointers, per canonical layer
  int n_slices;
  uint32_t total_groups;  // packed_bytes / 16
  uint8_t* packed;
  float alpha;  // < 0: no window
  int L_base, L_color;
};

cudaError_t launch_pack(const PackParams& p, int operand_dtype, cudaStream_t stream);
cudaError_t launch_composite_fwd(const b200r_composite_args& a, cudaStream_t stream);
cudaError_t launch_composite_bwd(const b200r_composite_bwd_args& b, cudaStream_t stream);
cudaError_t launch_field_fwd_desc(const b200r_field_desc& desc, const Program& prog, const b200r_field_args& args,
                                  const void* packed, int n_sm, cudaStream_t stream);

}  // namespace b200r
