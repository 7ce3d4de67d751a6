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
  // training forward (b200r_field_fwd_train): tape buffers, else NULL
  uint8_t* tape_a;
  uint32_t* tape_mask;
  TapeLayout tape;
};

// ---- eikonal modes of the field backward kernel (NeRF.compute_eikonal, lab4d/nnutils/nerf.py:416-453): masked LINEAR chains
// over the basefield on the samples of a list of rays, with the ReLU signs the training forward put on the tape.
//   mode 1 (reverse chain, b200r_eikonal_fwd): a_F = relu'(linear_final) * w_sdf, a_{l-1} = relu'_{l-1} * (W_l^T a_l), g = E(x)^T u
//   mode 2 / 3 (forward chains A / B, b200r_eikonal_bwd): v_0 = E(x) gbar, v_l = relu'_l * (W_l v_{l-1}); chain A walks every
//   layer (the skip layer takes its hidden columns only), chain B starts at the skip layer's embedding columns.
struct EikParams {
  int32_t mode;             // 0 = off (the kernel is the field backward)
  int32_t n_points;         // n_rays * D
  const int32_t* rays_sel;  // (n_rays) flat ray indices f * N + n of the training forward's batch
  const float* gbar;        // (n_points, 3) dL/dg            (modes 2, 3)
  float* g_out;             // (n_points, 3) g = d sdf / d x   (mode 1)
  const float* sdf_w;       // sdf.weight (W)                  (mode 1)
  uint8_t* tape;            // output chunks of this pass: n_tiles + kMaxCtas tiles of n_chunks chunks
  int32_t n_chunks;
  float scale_a;            // mode 1: power-of-two scale of the unit cotangent (16-bit operands)
  int16_t n_layers;         // modes 2, 3: wide layers of the chain
  int16_t v0_chunk;         // mode 2: chunk of v_0 (-1: not stored)
  int16_t head_chunk;       // mode 1: chunk whose column 3 carries the scaled unit cotangent (operand of d sdf.weight)
  int16_t pad_;
  int16_t mask_slot[12];    // modes 2, 3: sign-word slot of every layer, in chain order
  int16_t out_chunk[12];    // mode 1: chunk of a_i, i = 0..D (basefield layer order); modes 2, 3: chunk of every layer's v, in chain order
};

// ---- backward of the field kernel (csrc/field_bwd.cu)
struct BwdKernelParams {
  Program prog;                // build_bwd_program: transposed-weight blocks, same constant / frame block layouts
  TapeLayout tape;
  b200r_field_desc desc;
  b200r_ray_batch rays;
  b200r_field_outputs saved;   // per-sample outputs of the training forward: xyz, rgb, sdf, feature, feat_norm
  b200r_field_grads g;         // cotangents
  const uint8_t* packed_t;
  const float* workspace;      // constant block + frame blocks (rebuilt by the prologue kernel)
  const uint8_t* tape_a;
  uint8_t* tape_g;             // n_tiles + kMaxCtas tiles: dead tiles of a pair write to a scratch tile
  const uint32_t* tape_mask;
  float* g_cblk;               // gradient of the constant block (zeroed by the caller)
  float* g_fblk;               // gradient of the frame blocks (zeroed by the caller)
  const float* scale;          // device scalar: power-of-two gradient scale
  const float* dense_w3[2];    // ComposedWarp: post_warp.{forward_map, backward_map}.linear_final.weight (3, 256)
  const float* g_points;       // warp entry (b200r_warp_bwd): cotangent of the warped points (M*P, 3); the points themselves ride in saved.xyz
  float* g_points_out;         // warp entry: gradient w.r.t. the given points (M*P, 3); non-NULL selects the warp-only instantiation
                               // normals entry: d sdf / d xyz_cam of every sample (S, 3)
  int32_t normals;             // 1: the normals instantiation (b200r_field_normals); eik.scale_a = scale of the unit cotangent
  int32_t M, ND, tiles_per_frame, n_tiles;  // eikonal modes: ND / tiles_per_frame describe the training forward's batch, n_tiles the point tiles
  EikParams eik;
};
cudaError_t launch_field_bwd(const BwdKernelParams& p, int n_sm, cudaStream_t stream);

// ---- backward of the per-frame prologue (csrc/chain.cu)
struct ChainParams {
  ConstLayout cl;
  FrameLayout fl;
  b200r_field_desc desc;
  b200r_field_params par;
  b200r_frame_tables fr;
  b200r_frame_grads gf;
  b200r_param_grads off;   // offsets (the pointers in it are not used by the kernel)
  const float* g_cblk;
  const float* g_fblk;
  float* grad;             // flat gradient buffer
  int32_t n_layers;
  int16_t layer_out[B200R_MAX_LAYERS];
};
cudaError_t launch_chain(const ChainParams& p, cudaStream_t stream);

// ---- weight-gradient kernel (csrc/wgrad.cu): D[row0 + r][col0 + c] of a job's accumulator is added to dst[r * ld + c]
struct WgradView { int32_t row0, col0, rows, cols, ld, per_frame /* destination: 0 weights, 1 frame block, 2 constant block */; int64_t dst_off; };
struct WgradJob {
  int16_t g_chunk, n_g, a_chunk, n_a;  // operand chunk ranges of a tile: G (M = 64 n_g features) and A (N = 64 n_a features)
  uint8_t g_src, a_src;                // which tape holds the operand: 0 = forward-written, 1 = backward-written
  uint8_t n_views, colsum;             // colsum of the G operand: 0 none, 1 into the constant-block gradient, 2 per frame
  int32_t per_frame;                   // flush the accumulators at every frame boundary
  int32_t colsum_off, colsum_n;
  WgradView v[2];
};
struct WgradWork { int32_t job, tile0, tile1; };
struct WgradParams {
  const uint8_t* tape_a;
  const uint8_t* tape_g;
  int32_t n_a, n_g;            // chunks per tile in each tape
  const WgradJob* jobs;        // device
  const WgradWork* work;       // device
  const int32_t* cta_first;    // device, grid + 1 entries: work items of CTA b are [cta_first[b], cta_first[b + 1])
  float* grad;                 // base of the flat weight-gradient buffer (views with per_frame == 0)
  float* g_cblk;               // gradient of the constant block
  float* g_fblk;               // gradient of the M frame blocks
  int32_t frame_floats, tiles_per_frame;
  const float* inv_scale;      // device scalar: 1 / grad_scale of the gradient chunks
};
cudaError_t launch_wgrad(const WgradParams& p, int grid, int operand_dtype, cudaStream_t stream);

cudaError_t launch_pack(const PackParams& p, int operand_dtype, cudaStream_t stream);
cudaError_t launch_prologue(const PrologueParams& p, cudaStream_t stream);
cudaError_t launch_composite_fwd(const b200r_composite_args& a, cudaStream_t stream);
cudaError_t launch_composite_bwd(const b200r_composite_bwd_args& b, cudaStream_t stream);
cudaError_t launch_importance_fwd(const b200r_importance_args& a, cudaStream_t stream);
cudaError_t launch_compose_fwd(const b200r_compose_args& a, cudaStream_t stream);
cudaError_t launch_compose_bwd(const b200r_compose_bwd_args& b, cudaStream_t stream);
cudaError_t launch_quat_mul_fwd(const float* a, const float* b, float* out, long long B, int D1, int D2, cudaStream_t s);
cudaError_t launch_quat_mul_bwd(const float* g, const float* a, const float* b, float* ga, float* gb, long long B, int D1, int D2, cudaStream_t s);
cudaError_t launch_quat_mul_bwd_bwd(const float* u1, const float* u2, const float* g, const float* a, const float* b, float* gg, float* gga, float* ggb,
                                    long long B, int D1, int D2, cudaStream_t s);
cudaError_t launch_quat_conj(const float* q, float* out, long long B, cudaStream_t s);
cudaError_t launch_loss_fwd(const b200r_loss_args& a, cudaStream_t stream);
cudaError_t launch_loss_bwd(const b200r_loss_bwd_args& b, cudaStream_t stream);
size_t match_partial_floats(int R, int K);
cudaError_t launch_match_fwd(const b200r_match_args& a, int n_sm, cudaStream_t stream);
cudaError_t launch_match_bwd(const b200r_match_bwd_args& b, float* partial, cudaStream_t stream);
cudaError_t launch_field_fwd(const FieldKernelParams& p, int n_sm, cudaStream_t stream);
cudaError_t launch_field_fwd_train(const FieldKernelParams& p, int n_sm, cudaStream_t stream);

}  // namespace b200r
