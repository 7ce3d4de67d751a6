// Internal launch interfaces between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>

#include "program.h"

namespace b200r {

struct PackParams {
  const PackSlice* slices;                 // device copy of BuiltProgram::slices (cached in the handle)
  const float* weights[B200R_MAX_LAYERS];  // device pointers, per canonical layer
  int n_slices;
  uint32_t total_groups;  // packed_bytes / 16
  uint8_t* packed;
  float alpha;  // < 0: no window
  int L_base, L_color;
};

struct PrologueParams {
  ConstLayout cl;
  FrameLayout fl;
  b200r_field_desc desc;
  b200r_field_params par;
  b200r_frame_tables fr;
  float* workspace;  // [const block][M frame blocks]
  int32_t skip_cams;    // point entries: no cameras; bias rows whose codes are absent are skipped
  int32_t skip_bones;   // b200r_points_fwd: no bone tables
  int32_t n_layers;
  int32_t rgb0_layer;
  int16_t layer_out[B200R_MAX_LAYERS];
  int16_t layer_in[B200R_MAX_LAYERS];
};

struct FieldKernelParams {
  Program prog;
  b200r_field_desc desc;
  b200r_ray_batch rays;
  b200r_field_outputs out;
  const uint8_t* packed;
  const float* workspace;
  const float* points;     // points mode (b200r_points_fwd): (M, ND, 3) canonical points, else NULL
  const float* point_dirs; // points mode: (M, ND, 3) directions in field space or NULL
  uint4* scratch;          // per-CTA scratch for the packed base features (program.h kScratchPerCta)
  int32_t M;
  int32_t ND;              // N * D samples per frame
  int32_t tiles_per_frame;
  int32_t n_tiles;
  int32_t Lmax;            // frequencies of the shared embedding chunk(s): 10 or 12
  int32_t warp_mode;       // 0, MODE_WARP_BWD or MODE_WARP_FWD: b200r_warp_fwd evaluates one warp on `points`
};

cudaError_t launch_pack(const PackParams& p, int operand_dtype, cudaStream_t stream);
cudaError_t launch_prologue(const PrologueParams& p, cudaStream_t stream);
cudaError_t launch_composite_fwd(const b200r_composite_args& a, cudaStream_t stream);
cudaError_t launch_composite_bwd(const b200r_composite_bwd_args& b, cudaStream_t stream);
cudaError_t launch_importance_fwd(const b200r_importance_args& a, cudaStream_t stream);
cudaError_t launch_compose_fwd(const b200r_compose_args& a, cudaStream_t stream);
cudaError_t launch_field_fwd(const FieldKernelParams& p, int n_sm, cudaStream_t stream);

}  // namespace b200r
