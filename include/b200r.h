/* b200r — C ABI of the Blackwell-native Lab4D ray-sample renderer (libb200render.so).
 *
 * Drop-in boundary.  The reference has no C ABI for this path: its only native
 * interface is the pybind11 module of dqtorch (lab4d/third_party/quaternion/src/bindings.cpp:7-16,
 * eight functions taking at::Tensor) and everything else is torch ops glued in Python.  The entry
 * points below are what a ctypes binding of the reference's hot functions binds instead:
 *
 *   b200r_field_fwd       <- {NeRF,FeatureNeRF,Deformable}.query_field, training-mode body
 *                            (lab4d/nnutils/nerf.py:580-684, feature.py:89-133, deformable.py:300-356):
 *                            sample_cam_rays (utils/render_utils.py:8-56), cam_to_field (nerf.py:821-844),
 *                            SkinningWarp.forward x3 (nnutils/warping.py:277-336, skinning.py:89-153,
 *                            utils/geom_utils.py:45-83), PosEmbedding (nnutils/embedding.py:69-125),
 *                            NeRF.forward (nerf.py:167-215), VisField.forward (visibility.py:52-63),
 *                            compute_feat (feature.py:136-150), compute_flow (nerf.py:948-997),
 *                            cycle_loss (deformable.py:173-198), compute_gauss_density (deformable.py:329-356)
 *   b200r_pack_weights    <- no reference counterpart: converts the nn.Linear weights (row-major (out,in)
 *                            fp32, the reference state_dict layout) into 16-bit swizzled UMMA operand tiles
 *   b200r_composite_fwd   <- render_pixel / compute_weights / integrate (utils/render_utils.py:59-184)
 *   b200r_composite_bwd   <- autograd of the above (hand-derived)
 *
 * Conventions: every pointer is a DEVICE pointer to contiguous fp32 unless stated; sizes are element
 * counts; every call takes the CUDA stream to launch on (pass torch.cuda.current_stream().cuda_stream);
 * returns 0 on success or a negative B200R_E_* code, never throws; b200r_last_error() gives the text.
 * A handle is per-device and not thread-safe (one process per GPU, one driving thread).
 */
#ifndef B200R_H
#define B200R_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200R_OK 0
#define B200R_E_INVALID -1 /* bad argument / unsupported shape */
#define B200R_E_CUDA -2    /* CUDA runtime error (text in b200r_last_error) */
#define B200R_E_ARCH -3    /* device is not sm_100 */

#define B200R_MAX_LAYERS 32 /* distinct dense layers of one field */

typedef struct b200r_handle b200r_handle;
typedef void* b200r_stream; /* cudaStream_t */

/* Field architecture (MultiFields.define_field, lab4d/nnutils/multifields.py:60-100). */
typedef struct {
  int32_t category;      /* 0 = fg (Deformable), 1 = bg (NeRF) */
  int32_t D;             /* basefield depth: linear_1..linear_D + linear_final */
  int32_t W;             /* hidden width: 256 or 128 */
  int32_t L_xyz;         /* position frequencies of the density branch (colour uses L_xyz+2) */
  int32_t L_dir;         /* -1: no direction input, 0: raw direction (3 ch) */
  int32_t appr_channels; /* per-frame appearance code width (folded into the rgb.0 bias) */
  int32_t skip;          /* skip-concat layer index (4) */
  int32_t n_bones;       /* 0 = rigid field, else SkinningWarp with B bones (18 or 25) */
  int32_t has_feature;   /* FeatureNeRF feature field present */
  int32_t operand_dtype; /* tensor-core operand type (fp32 accumulate): 0 = fp16, 1 = bf16, 2 = fp16 head + tail
                            (every operand split as x = fp16(x) + fp16(x - fp16(x)), three MMAs per k-step:
                            ~22-bit operands, the mode that meets the 1e-4 rendered-RGB parity contract) */
  int32_t dense;         /* 1: ComposedWarp = SkinningWarp + DenseWarp(D=2, W=256, 6 xyz frequencies)
                            (nnutils/warping.py:104-170, 417-483); needs n_bones > 0 */
  int32_t pad_;
} b200r_field_desc;

/* Dense layers, in this canonical order (absent groups are skipped):
 *   [delta_field.linear_1, linear_2, linear_final]        if n_bones > 0
 *   [vis_mlp.basefield.linear_1, linear_2]
 *   [basefield.linear_1 .. linear_D, linear_final]
 *   [rgb.0]
 *   [colorfield.linear_1, linear_2, linear_final]
 *   [feature_field.linear_1 .. linear_5, linear_final]     if has_feature
 *   [post_warp.forward_map.linear_1, linear_2, linear_final,
 *    post_warp.backward_map.linear_1, linear_2, linear_final] if dense
 * b200r_layer_count() returns how many that is for a descriptor. */
int b200r_layer_count(const b200r_field_desc* desc);

/* Bytes of the packed operand buffer for a descriptor. */
size_t b200r_packed_bytes(const b200r_field_desc* desc);

int b200r_create(int device, b200r_handle** out);
void b200r_destroy(b200r_handle* h);
const char* b200r_last_error(const b200r_handle* h);

/* Parameters of one field: device pointers into the reference modules' own storage (state_dict
 * layout).  Nothing is copied or owned by the library. */
typedef struct {
  const float* weight[B200R_MAX_LAYERS]; /* (out_i, in_i) row-major nn.Linear weight of canonical layer i */
  const float* bias[B200R_MAX_LAYERS];   /* (out_i) */
  const float* sdf_w;         /* sdf.weight (W) */
  const float* sdf_b;         /* (1) */
  const float* rgb2_w;        /* rgb.2.weight (3, W/2) */
  const float* rgb2_b;        /* (3) */
  const float* vis_final_w;   /* vis_mlp.basefield.linear_final.weight (64) */
  const float* vis_final_b;   /* (1) */
  const float* logibeta;      /* (1) */
  const float* logscale;      /* (1) */
  const float* warp_logibeta; /* warp.logibeta (1), n_bones > 0 */
  const float* log_gauss;     /* warp.skinning_model.log_gauss (B,3), n_bones > 0 */
  const int32_t* symm_idx;    /* (B) left/right bone pairing or NULL (nnutils/skinning.py:150-153) */
} b200r_field_params;

/* Per-frame inputs (M rows each): what get_samples() hands to query_field
 * (nnutils/nerf.py:530-578, deformable.py:254-289) plus the per-frame codes the MLPs are conditioned on. */
typedef struct {
  int32_t M;
  int32_t pad_;
  const float* Kinv;             /* (M,3,3) */
  const float* near_far;         /* (M,2) */
  const float* field2cam_q;      /* (M,4) real-first quaternion */
  const float* field2cam_t;      /* (M,3) already scaled by exp(logscale) */
  const float* inst_base;        /* (M,32) basefield.inst_embedding rows */
  const float* inst_color;       /* (M,32) */
  const float* inst_vis;         /* (M,32) */
  const float* appr_code;        /* (M,appr_channels) or NULL */
  const float* inst_skin;        /* (M,32)  n_bones > 0 */
  const float* skin_t_embed;     /* (M,128) skinning time embedding of each frame */
  const float* skin_t_embed_mean;/* (128)   mean time embedding (forward warps, warping.py:313-314) */
  const float* dense_t_embed;    /* (M,128) post_warp.time_embedding rows            (dense) */
  const float* inst_dense_fwd;   /* (M,32)  post_warp.forward_map.inst_embedding rows (dense) */
  const float* inst_dense_bwd;   /* (M,32)  post_warp.backward_map.inst_embedding rows (dense) */
  const float* t_art_qr;         /* (M,B,4) t_articulation real part */
  const float* t_art_qd;         /* (M,B,4) dual part */
  const float* rest_art_qr;      /* (M,B,4) */
  const float* rest_art_qd;      /* (M,B,4) */
} b200r_frame_tables;

typedef struct {
  int32_t N, D;        /* rays per frame, samples per ray (D >= 2) */
  float flow_thresh;   /* < 0: None */
  int32_t pad_;
  const float* hxy;    /* (M,N,3) homogeneous pixel coordinates */
  const float* depth;  /* (M,N,D) sample depths along every ray, ascending (sample_cam_rays(depth=...), e.g. from
                          b200r_importance_fwd), or NULL: uniform placement between the frame's near and far planes */
} b200r_ray_batch;

/* Per-sample outputs, (M*N*D, c) row-major fp32; any pointer may be NULL. */
typedef struct {
  float* rgb;           /* 3 */
  float* density;       /* 1 */
  float* vis;           /* 1 */
  float* xyz;           /* 3 canonical */
  float* xyz_cam;       /* 3 */
  float* xyz_t;         /* 3 time-t object space */
  float* dir;           /* 3 field-space ray direction */
  float* depth;         /* 1 (already divided by exp(logscale)) */
  float* deltas;        /* 1 */
  float* feature;       /* 16 */
  float* flow;          /* 3 */
  float* cyc_dist;      /* 1 */
  float* delta_skin;    /* 1 */
  float* skin_entropy;  /* 1 */
  float* gauss_density; /* 1 */
  float* sdf;           /* 1 */
  float* feat_norm;     /* 1: 1 / |feature| before normalisation (kept by the training forward for the backward) */
  float* warp_pts;      /* 9: ComposedWarp fields, training forward: the skinned point before the soft deformation and the two
                           softly deformed points that enter the forward skinning warps (flow partner, cycle) */
} b200r_field_outputs;

/* Cotangents of the per-sample outputs of one query_field call, (M*N*D, c) row-major fp32 like b200r_field_outputs;
 * NULL = zero.  What b200r_composite_bwd (or autograd of the caller's own reductions) hands to b200r_field_bwd. */
typedef struct {
  const float* rgb;           /* 3 */
  const float* density;       /* 1 */
  const float* vis;           /* 1 */
  const float* feature;       /* 16 */
  const float* xyz;           /* 3 */
  const float* xyz_cam;       /* 3 */
  const float* depth;         /* 1 */
  const float* flow;          /* 3 (the validity flag's entry is ignored) */
  const float* cyc_dist;      /* 1 */
  const float* delta_skin;    /* 1 */
  const float* skin_entropy;  /* 1 */
  const float* gauss_density; /* 1 */
} b200r_field_grads;

/* alpha: PosEmbedding annealing window (nnutils/embedding.py:112-125) folded into the packed weights
 * of basefield / colorfield; negative = None.  Call after every optimiser step / set_alpha. */
int b200r_pack_weights(b200r_handle* h, const b200r_field_desc* desc, const b200r_field_params* params, float alpha,
                       void* packed, size_t packed_bytes, b200r_stream stream);

/* Bytes of scratch the forward needs for M frames (per-frame blocks built by the prologue kernel). */
size_t b200r_workspace_bytes(const b200r_field_desc* desc, int32_t M);

/* One training-mode query_field call on M frames x N rays x D samples: a per-frame prologue kernel
 * (bias rows with the per-frame codes folded in, bone transforms) then the fused per-sample kernel. */
int b200r_field_fwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed, const b200r_field_params* params,
                    const b200r_frame_tables* frames, const b200r_ray_batch* rays, const b200r_field_outputs* out,
                    void* workspace, size_t workspace_bytes, b200r_stream stream);

/* ------------------------------------------------------------------ training: forward with a tape, backward
 * Replaces `total_loss.mean().backward()` through query_field (lab4d/engine/trainer.py:344-345; autograd of
 * nnutils/nerf.py:580-684, warping.py:277-336, skinning.py:89-153, third_party/quaternion/src/quaternion.cu:67-199).
 * The training forward records, per 128-sample tile, every tensor-core operand it produced (16-bit chunks) and the
 * ReLU signs; b200r_field_bwd turns the cotangents of the per-sample outputs into
 *   - the gradient of every layer's weight (columns fed by per-sample operands) in one flat fp32 buffer,
 *   - the gradient of the constant block (plain bias rows, head weights, rest bone centres, scalars),
 *   - the gradient of the M per-frame blocks (cameras, bias rows that carry per-frame codes, bone tables),
 * whose layouts b200r_get_block_layout describes; the host side chains the two blocks to the per-frame inputs
 * (M rows of quaternion algebra and code mat-vecs, lab4d_b200/prologue_grad.py). */
typedef struct {
  void* a;      /* forward-written operand chunks, a_bytes (1024-B aligned) */
  void* g;      /* backward-written gradient chunks, g_bytes (1024-B aligned) */
  void* mask;   /* ReLU sign words, mask_bytes */
  size_t a_bytes, g_bytes, mask_bytes;
} b200r_tape;

/* Bytes of the three tape buffers for M frames x N rays x D samples. */
int b200r_tape_sizes(const b200r_field_desc* desc, int32_t M, int32_t N, int32_t D, size_t* a_bytes, size_t* g_bytes,
                     size_t* mask_bytes);

/* b200r_field_fwd that also fills tape->a and tape->mask.  `out` must keep xyz, rgb, sdf (and feature, feat_norm). */
int b200r_field_fwd_train(b200r_handle* h, const b200r_field_desc* desc, const void* packed, const b200r_field_params* params,
                          const b200r_frame_tables* frames, const b200r_ray_batch* rays, const b200r_field_outputs* out,
                          const b200r_tape* tape, void* workspace, size_t workspace_bytes, b200r_stream stream);

/* Transposed operand tiles (W^T) for the backward's data-gradient GEMMs; same conventions as b200r_pack_weights. */
size_t b200r_packed_t_bytes(const b200r_field_desc* desc);
int b200r_pack_weights_t(b200r_handle* h, const b200r_field_desc* desc, const b200r_field_params* params, float alpha,
                         void* packed_t, size_t packed_bytes, b200r_stream stream);

/* Host-side introspection (no device work): number of tensor-core steps (weight-ring slots) of one per-tile program of a field, or
 * B200R_E_INVALID when the configuration does not build.  kind: 0 forward (query_field), 1 backward (data gradient), 2 its density
 * chain alone (eikonal reverse chain), 3 the forward warp w = 2 alone (b200r_warp_bwd), 4 density chain + backward warp (normals),
 * 5 / 6 eikonal forward chains A / B, 7 the backward warp w = 0 alone. */
int b200r_program_steps(const b200r_field_desc* desc, int32_t kind);

/* Float offsets inside the constant block and inside one frame block (-1 = absent). */
#define B200R_MAX_COND 12
typedef struct {
  int32_t const_floats, frame_floats;
  /* constant block */
  int32_t c_plain_bias[B200R_MAX_LAYERS]; /* bias row of layer i when it carries no per-frame code */
  int32_t c_sdf_w, c_rgb2_w, c_vis_w, c_dir_w; /* sdf.weight (W), rgb.2.weight (3, W/2), vis final weight (64), rgb.0 direction columns (W/2, 3) */
  int32_t c_center;                       /* rest bone centres (B, 4) */
  int32_t c_scalars;                      /* 8 scalars; gradients: [0] d/d logibeta, [2] d/d warp.logibeta, [1] d/d logscale (rendered depth only),
                                             [3] sdf.bias, [4..6] rgb.2.bias, [7] vis final bias */
  /* frame block */
  int32_t f_cam, f_cam_partner;           /* 24 floats: Kinv[9], near, far, q[4], t[3]; GRADIENT slots 11..17 are w.r.t. the inverse
                                             camera (q^-1, -q^-1 t q) for f_cam and w.r.t. (q, t) for f_cam_partner */
  int32_t f_binv_t, f_se3_bwd, f_binv_rest, f_se3_fwd, f_binv_rest_partner, f_se3_fwd_partner; /* (B,12) / (B,8) */
  int32_t n_cond;
  struct {
    int32_t layer, n, in_dim, frame_off, n_seg;
    int32_t col0[2], width[2], code[2];   /* bias row = b + sum_seg W[:, col0:col0+width] @ code[frame]; code ids: 0 inst_base,
                                             1 inst_color, 2 inst_vis, 3 appr, 4 inst_skin, 5 skin_t_embed, 6 skin_t_embed_mean, 7 dense_t,
                                             8 dense_t (partner frame), 9 inst_dense_fwd, 10 inst_dense_bwd */
  } cond[B200R_MAX_COND];
} b200r_block_layout;
int b200r_get_block_layout(const b200r_field_desc* desc, b200r_block_layout* out);

typedef struct {
  float* flat;                              /* flat fp32 gradient buffer of the field's hot-path parameters, ACCUMULATED into
                                               (zero it for a fresh gradient); all offsets below are float offsets into it, -1 = absent */
  int64_t weight_off[B200R_MAX_LAYERS];     /* layer i's (out_i, in_i) weight */
  int64_t bias_off[B200R_MAX_LAYERS];       /* layer i's bias */
  int64_t sdf_w, sdf_b, rgb2_w, rgb2_b, vis_final_w, vis_final_b, logibeta, logscale, warp_logibeta, log_gauss;
  float* const_block;                       /* (const_floats)     scratch: gradient of the constant block, overwritten */
  float* frame_block;                       /* (M, frame_floats)  scratch: gradient of the frame blocks, overwritten */
} b200r_param_grads;

/* Gradients of the per-frame inputs, shapes of b200r_frame_tables; OVERWRITTEN; any pointer may be NULL. */
typedef struct {
  float* Kinv;              /* (M,3,3) */
  float* field2cam_q;       /* (M,4) */
  float* field2cam_t;       /* (M,3) */
  float* inst_base;         /* (M,32) */
  float* inst_color;
  float* inst_vis;
  float* appr_code;         /* (M,appr_channels) */
  float* inst_skin;
  float* skin_t_embed;      /* (M,128) */
  float* skin_t_embed_mean; /* (128) */
  float* dense_t_embed;     /* (M,128) */
  float* inst_dense_fwd;
  float* inst_dense_bwd;
  float* t_art_qr;          /* (M,B,4) */
  float* t_art_qd;
  float* rest_art_qr;
  float* rest_art_qd;
} b200r_frame_grads;

/* Backward of one b200r_field_fwd_train call (same desc, params, frames, rays; `saved` = its outputs; tape->g is scratch):
 * data-gradient kernel, weight-gradient kernel, then the backward of the per-frame prologue. */
int b200r_field_bwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed_t, const b200r_field_params* params,
                    const b200r_frame_tables* frames, const b200r_ray_batch* rays, const b200r_field_outputs* saved,
                    const b200r_field_grads* grads, const b200r_tape* tape, const b200r_param_grads* out,
                    const b200r_frame_grads* frame_grads, void* workspace, size_t workspace_bytes, b200r_stream stream);

/* ------------------------------------------------------------------ eikonal term (NeRF.compute_eikonal)
 * Replaces NeRF.compute_eikonal (lab4d/nnutils/nerf.py:416-453) and the second-order autograd it rests on
 * (compute_gradient, lab4d/utils/torch_utils.py:4-28: autograd.grad(sdf, xyz, create_graph=True), differentiated again by
 * total_loss.backward(), engine/trainer.py:344-345; through dqtorch for nothing - the points are detached).
 * The reference evaluates g = d sdf / d xyz on all D samples of a random subset of the batch's rays and returns
 * (|g| - 1)^2.  Here g is a REVERSE chain through the basefield with the ReLU signs the training forward left on the tape
 * (b200r_eikonal_fwd, run after b200r_field_fwd_train of the same batch), and the gradient of a loss of g w.r.t. the
 * basefield weights and sdf.weight is two FORWARD chains from dL/dg plus weight-gradient GEMMs of the chains' operands
 * (b200r_eikonal_bwd) - both on tcgen05 with single 16-bit operands (fp32 accumulate).  Biases and instance codes get no
 * gradient (they only move the ReLU signs), exactly as in the reference. */
typedef struct {
  int32_t n_rays;        /* selected rays (the reference: M*N / 16, torch.multinomial) */
  int32_t pad_;
  const int32_t* rays;   /* (n_rays) device: flat ray index f * N + n into the batch of the training forward */
  void* a;               /* reverse-chain tape: written by b200r_eikonal_fwd, read by b200r_eikonal_bwd (1024-B aligned) */
  void* v;               /* forward-chain tape: scratch of b200r_eikonal_bwd (1024-B aligned) */
  size_t a_bytes, v_bytes;
} b200r_eik_batch;

/* Bytes of the two tapes for n_rays rays of D samples. */
int b200r_eikonal_sizes(const b200r_field_desc* desc, int32_t n_rays, int32_t D, size_t* a_bytes, size_t* v_bytes);

/* g_out (n_rays*D, 3) = d sdf / d xyz at the selected rays' samples.  `rays` is the training forward's ray batch (N, D),
 * M its frame count, saved_xyz its per-sample canonical points (M*N*D, 3), `tape` its tape (sign words), packed_t the
 * transposed operand tiles of b200r_pack_weights_t. */
int b200r_eikonal_fwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed_t, const b200r_field_params* params,
                      const b200r_ray_batch* rays, int32_t M, const float* saved_xyz, const b200r_tape* tape,
                      const b200r_eik_batch* eik, float* g_out, b200r_stream stream);

/* Given g_g (n_rays*D, 3) = dL/dg: ACCUMULATES dL/dW of basefield.linear_1..D, linear_final and sdf.weight into out->flat
 * (weight_off of those layers and sdf_w; nothing else of `out` is used).  `packed` = the forward operand tiles
 * (b200r_pack_weights; the heads of the split mode). */
int b200r_eikonal_bwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed, const b200r_field_params* params,
                      const b200r_ray_batch* rays, int32_t M, const float* saved_xyz, const b200r_tape* tape,
                      const b200r_eik_batch* eik, const float* g_g, const b200r_param_grads* out, b200r_stream stream);

/* Eval-mode normals (NeRF.compute_normal, lab4d/nnutils/nerf.py:455-493; lab4d/render.py -> dvr_model.evaluate): g_cam (M*N*D, 3)
 * = d sdf / d xyz_cam at every sample - the gradient of the sdf through the basefield AND the backward warp w.r.t. the
 * camera-space point, which the reference takes with autograd.grad over the whole batch.  Runs after a b200r_field_fwd_train
 * of the same batch (its tape holds the ReLU signs and the warp's operands; tape->g is scratch): reverse chain of the density
 * branch from a unit sdf cotangent, the backward warp's backward, the camera rotation.  No parameter gradients.
 * The caller forms eikonal = (|g| - 1)^2 and normal = g / |g| * (1, -1, -1). */
int b200r_field_normals(b200r_handle* h, const b200r_field_desc* desc, const void* packed_t, const b200r_field_params* params,
                        const b200r_frame_tables* frames, const b200r_ray_batch* rays, const b200r_field_outputs* saved,
                        const b200r_tape* tape, float* g_cam, void* workspace, size_t workspace_bytes, b200r_stream stream);

/* NeRF.forward on given points (lab4d/nnutils/nerf.py:167-215), the boundary the reference's flat-point callers use
 * (geometry_init nerf.py:277, extract_canonical_mesh :328, eval-mode query_nerf :794-805): canonical points in, rgb /
 * density / sdf out.  Only the basefield, colorfield, sdf and rgb heads run (no ray placement, warps, visibility or
 * feature field); `frames` needs M and the instance / appearance code rows only.  Outputs other than rgb, density,
 * sdf and xyz (echo) must be NULL. */
typedef struct {
  int32_t P;           /* points per frame row */
  int32_t pad_;
  const float* xyz;    /* (M,P,3) points in the field's canonical space */
  const float* dir;    /* (M,P,3) view directions in field space, or NULL (L_dir = -1 or density only) */
} b200r_point_batch;

int b200r_points_fwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed, const b200r_field_params* params,
                     const b200r_frame_tables* frames, const b200r_point_batch* points, const b200r_field_outputs* out,
                     void* workspace, size_t workspace_bytes, b200r_stream stream);

/* SkinningWarp.forward / ComposedWarp.forward on given points (lab4d/nnutils/warping.py:277-336, 445-483), the boundary
 * of export.extract_deformation (lab4d/export.py:94-130) and soft_deform_loss (deformable.py:238-252).
 * backward != 0: time-t space -> canonical (bone coordinates from t_articulation, the frame's time code);
 * backward == 0: canonical -> time-t space with the frame's own articulation (mean time code, warping.py:313-314).
 * `frames` needs M, the skinning (and dense-warp) code rows and the articulations; outputs: xyz (warped points),
 * skin_entropy, delta_skin; everything else must be NULL.  Needs n_bones > 0. */
int b200r_warp_fwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed, const b200r_field_params* params,
                   const b200r_frame_tables* frames, const b200r_point_batch* points, int32_t backward,
                   const b200r_field_outputs* out, void* workspace, size_t workspace_bytes, b200r_stream stream);

/* The FORWARD warp (canonical -> time-t space, the frame's own articulation) on given points as a differentiable pair: what
 * FeatureNeRF.forward_project (lab4d/nnutils/feature.py:207-226) and NeRF.forward_warp (nerf.py:846-870) run on the matched
 * points in a training step, and autograd through them (engine/trainer.py:344-345).  b200r_warp_fwd_train = b200r_warp_fwd
 * (backward = 0) that also records the warp's slice of the tape (sizes: b200r_tape_sizes(desc, M, P, 1)); `out` keeps xyz (the
 * warped points; warp_pts too for ComposedWarp fields).  b200r_warp_bwd takes g_xyz (M*P,3), the cotangent of the warped points,
 * and returns g_points (M*P,3) w.r.t. the given points, ACCUMULATES the gradients of the delta MLP (and soft-deformation map)
 * weights / biases, log_gauss into out->flat and OVERWRITES the per-frame gradients (articulations, skinning / dense codes) of
 * frame_grads - same conventions and scratch blocks as b200r_field_bwd.  `saved`: the forward's warp_pts (ComposedWarp), else unused. */
int b200r_warp_fwd_train(b200r_handle* h, const b200r_field_desc* desc, const void* packed, const b200r_field_params* params,
                         const b200r_frame_tables* frames, const b200r_point_batch* points, const b200r_field_outputs* out,
                         const b200r_tape* tape, void* workspace, size_t workspace_bytes, b200r_stream stream);
int b200r_warp_bwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed_t, const b200r_field_params* params,
                   const b200r_frame_tables* frames, const b200r_point_batch* points, const b200r_field_outputs* saved, const float* g_xyz,
                   const b200r_tape* tape, const b200r_param_grads* out, const b200r_frame_grads* frame_grads, float* g_points,
                   void* workspace, size_t workspace_bytes, b200r_stream stream);

/* ------------------------------------------------------------------ compositing (render_pixel) */
#define B200R_MAX_CHANNELS 16
/* how a per-sample array (R*D, nch) is reduced along the ray */
#define B200R_CH_NORM 0       /* sum_k w_k/(mask+1e-6) v_k                 (rgb, depth, xyz, feature, ...) */
#define B200R_CH_NORM_FROZEN 1 /* same, weights detached in backward         (cyc_dist, xyz_cam, skin_entropy) */
#define B200R_CH_MEAN 2       /* plain mean over samples and channels        (eikonal, delta_skin) */
#define B200R_CH_FLOW 3       /* (R*D,3): w*valid renormalised, 2 outputs   (flow) */
#define B200R_CH_WEIGHTSUM 4  /* v is a density: sum of ITS OWN weights     (gauss_density -> gauss_mask) */
#define B200R_CH_VIS 5        /* out[0] = sum_k logsigmoid(v_k) T_k, out[1] = sum_k T_k   (vis, host normalises) */

typedef struct {
  int32_t R, D;               /* rays, samples per ray */
  const float* density;       /* (R*D) */
  const float* deltas;        /* (R*D) */
  float* mask;                /* (R) sum of weights */
  float* weights;             /* (R*D) optional */
  float* transmit;            /* (R*D) optional */
  int32_t n_channels;
  const float* src[B200R_MAX_CHANNELS];
  float* dst[B200R_MAX_CHANNELS];
  int32_t nch[B200R_MAX_CHANNELS];
  int32_t mode[B200R_MAX_CHANNELS];
} b200r_composite_args;

int b200r_composite_fwd(b200r_handle* h, const b200r_composite_args* args, b200r_stream stream);

typedef struct {
  b200r_composite_args fwd;                  /* same tensors as the forward call */
  const float* g_mask;                       /* (R) or NULL */
  const float* g_dst[B200R_MAX_CHANNELS];    /* gradient of each rendered output, or NULL */
  float* g_density;                          /* (R*D) out */
  float* g_src[B200R_MAX_CHANNELS];          /* gradient of each per-sample array, or NULL */
} b200r_composite_bwd_args;

int b200r_composite_bwd(b200r_handle* h, const b200r_composite_bwd_args* args, b200r_stream stream);

/* ------------------------------------------------------------------ importance sampling (eval-mode sample placement)
 * NeRF.importance_sampling (lab4d/nnutils/nerf.py:686-738) after the coarse pass: from the Dc coarse depths of every ray
 * and their compositing weights (b200r_composite_fwd's `weights`), draw Dc deterministic inverse-CDF samples on the
 * mid-points with weights[1:-1] (sample_pdf, lab4d/utils/render_utils.py:187-233, det=True) and merge them with the
 * coarse depths: depth_out (R, 2*Dc) ascending, ready for b200r_ray_batch.depth. */
typedef struct {
  int32_t R, Dc;            /* rays, coarse samples per ray (Dc >= 4) */
  const float* depth_c;     /* (R*Dc) coarse depths, ascending */
  const float* weights;     /* (R*Dc) compositing weights of the coarse samples */
  float* depth_out;         /* (R*2*Dc) */
} b200r_importance_args;

int b200r_importance_fwd(b200r_handle* h, const b200r_importance_args* args, b200r_stream stream);

/* ------------------------------------------------------------------ compose_fields (two fields -> one sample list)
 * MultiFields.compose_fields (lab4d/nnutils/multifields.py:339-398): the samples of field A (first in field order) and
 * field B along every ray are merged by depth (A first on ties); every per-sample array is gathered into the merged
 * order.  A NULL source means the field lacks that key and contributes zeros.  More than two fields: merge pairwise. */
typedef struct {
  int32_t R, Da, Db;          /* rays, samples per ray of field A / B; inputs sorted by depth along the ray */
  int32_t n_channels;
  const float* depth_a;       /* (R*Da) */
  const float* depth_b;       /* (R*Db) */
  int32_t* perm;              /* (R*(Da+Db)) optional: index into the concatenation [A; B] of every output sample */
  const float* src_a[B200R_MAX_CHANNELS];  /* (R*Da, nch) or NULL */
  const float* src_b[B200R_MAX_CHANNELS];  /* (R*Db, nch) or NULL */
  float* dst[B200R_MAX_CHANNELS];          /* (R*(Da+Db), nch) */
  int32_t nch[B200R_MAX_CHANNELS];
} b200r_compose_args;

int b200r_compose_fwd(b200r_handle* h, const b200r_compose_args* args, b200r_stream stream);

/* Backward of b200r_compose_fwd (autograd through multifields.py:393-397): the gradient of every merged per-sample array goes
 * back to the two fields through the permutation the forward wrote (`perm`).  g_a / g_b are OVERWRITTEN (every sample of a
 * field appears exactly once in the merge); NULL = that field lacks the key. */
typedef struct {
  int32_t R, Da, Db;
  int32_t n_channels;
  const int32_t* perm;                       /* (R*(Da+Db)) from the forward */
  const float* g_dst[B200R_MAX_CHANNELS];    /* (R*(Da+Db), nch) gradient of the merged array */
  float* g_a[B200R_MAX_CHANNELS];            /* (R*Da, nch) or NULL */
  float* g_b[B200R_MAX_CHANNELS];            /* (R*Db, nch) or NULL */
  int32_t nch[B200R_MAX_CHANNELS];
} b200r_compose_bwd_args;

int b200r_compose_bwd(b200r_handle* h, const b200r_compose_bwd_args* args, b200r_stream stream);

/* ------------------------------------------------------------------ per-ray feature matching (FeatureNeRF.global_match)
 * lab4d/nnutils/feature.py:152-205: every ray's pixel feature is matched against K candidate samples of the batch,
 *   score[r,k] = exp(logsigma) <feat_px[r], feat_can[idx[k]]>, prob = softmax_k, xyz_matched[r] = sum_k prob[r,k] xyz_can[idx[k]].
 * The caller draws idx (the reference's torch.randperm(S)[:K], no duplicates), so the random stream stays the reference's.
 * b200r_match_bwd is the hand-derived backward of autograd through it (engine/trainer.py:344-345): dense gradients of the
 * per-sample features / points (rows idx[k] are ADDED to; zero the arrays first) and of logsigma (ADDED). */
#define B200R_MATCH_CHANNELS 16
#define B200R_MATCH_MAX_K 2048
typedef struct {
  int32_t R, K;             /* rays, candidates (K <= B200R_MATCH_MAX_K) */
  const float* feat_px;     /* (R,16) pixel features */
  const float* feat_can;    /* (S,16) canonical features of every sample of the batch (feat_dict["feature"]) */
  const float* xyz_can;     /* (S,3)  canonical points of every sample (feat_dict["xyz"]) */
  const int64_t* idx;       /* (K) device: candidate sample indices */
  const float* logsigma;    /* (1) FeatureNeRF.logsigma */
  float* xyz_matched;       /* (R,3) out */
  float* lse;               /* (R) out: log-sum-exp of every ray's scores (kept for the backward), or NULL */
} b200r_match_args;

int b200r_match_fwd(b200r_handle* h, const b200r_match_args* args, b200r_stream stream);

typedef struct {
  b200r_match_args fwd;     /* the forward call's tensors (xyz_matched and lse as it wrote them) */
  const float* g_out;       /* (R,3) gradient of xyz_matched */
  float* g_feat_can;        /* (S,16) or NULL */
  float* g_xyz_can;         /* (S,3) or NULL */
  float* g_logsigma;        /* (1) or NULL */
  float* scratch;           /* b200r_match_scratch_floats(R, K) floats */
} b200r_match_bwd_args;

size_t b200r_match_scratch_floats(int32_t R, int32_t K);
int b200r_match_bwd(b200r_handle* h, const b200r_match_bwd_args* args, b200r_stream stream);

/* ------------------------------------------------------------------ per-pixel reconstruction losses (dvr_model.compute_loss)
 * lab4d/engine/model.py: get_mask_balance_wt (:386-412), compute_recon_loss (:415-498), mask_losses (:520-574) and
 * apply_loss_weights (:576-611) for the terms compute_recon_loss creates, fused: every term k ends as
 *   loss[k] = mean over the entries with value > 0 [/ train_res for flow, feat_reproj] * wt[k].
 * Term order (the reference's loss_dict): 0 mask, 1 feature, 2 feat_reproj, 3 rgb, 4 depth, 5 flow, 6 vis, 7 reg_gauss_mask.
 * All arrays are (M*N, c) row-major fp32 (booleans of the batch as 0 / 1).  Terms whose inputs are absent (bg fields have no
 * feature / feat_reproj / reg_gauss_mask) come back as NaN and must be ignored. */
#define B200R_LOSS_TERMS 8
#define B200R_LOSS_STATS 24
typedef struct {
  int32_t M, N;               /* frames, rays per frame */
  int32_t field_type;         /* 0 fg, 1 bg, 2 comp (config["field_type"]) */
  float train_res;            /* config["train_res"] */
  const float* r_mask;        /* (R)    rendered["mask"] */
  const float* r_mask_fg;     /* (R)    rendered["mask_fg"] (comp) */
  const float* r_rgb;         /* (R,3)  rendered["rgb"] */
  const float* r_depth;       /* (R)    rendered["depth"] */
  const float* r_flow;        /* (R,2)  rendered["flow"] */
  const float* vis_fg;        /* (R)    aux_dict["fg"]["vis"] or NULL */
  const float* vis_bg;        /* (R)    aux_dict["bg"]["vis"] or NULL (enters with 0.01) */
  const float* a_feature;     /* (R,16) aux_dict["fg"]["feature"]   (fg / comp) */
  const float* a_xy_reproj;   /* (R,2)  aux_dict["fg"]["xy_reproj"] (fg / comp) */
  const float* a_gauss_mask;  /* (R)    aux_dict["fg"]["gauss_mask"] or NULL */
  const float* b_mask;        /* (R)    batch["mask"] */
  const float* b_vis2d;       /* (R)    batch["vis2d"] */
  const float* b_is_detected; /* (M)    batch["is_detected"] */
  const float* b_rgb;         /* (R,3) */
  const float* b_depth;       /* (R) */
  const float* b_flow;        /* (R,2) */
  const float* b_flow_uct;    /* (R) */
  const float* b_feature;     /* (R,16) */
  const float* b_hxy;         /* (R,3) */
  float wt[B200R_LOSS_TERMS]; /* config["<term>_wt"] (1 when absent) */
  float* loss;                /* (8) out */
  float* stats;               /* (B200R_LOSS_STATS) out: sums / counts of the positive entries, mask-balance weights (read by the backward) */
} b200r_loss_args;

int b200r_loss_fwd(b200r_handle* h, const b200r_loss_args* args, b200r_stream stream);

typedef struct {
  b200r_loss_args fwd;        /* the forward call's tensors (stats as it wrote them) */
  const float* g_loss;        /* (8) gradient of every term (0 for absent terms) */
  float* g_mask;              /* (R)   any of the outputs may be NULL; all are OVERWRITTEN */
  float* g_mask_fg;           /* (R)   comp */
  float* g_rgb;               /* (R,3) */
  float* g_depth;             /* (R) */
  float* g_flow;              /* (R,2) */
  float* g_vis_fg;            /* (R) */
  float* g_vis_bg;            /* (R) */
  float* g_feature;           /* (R,16) */
  float* g_xy_reproj;         /* (R,2) */
  float* g_gauss_mask;        /* (R) */
} b200r_loss_bwd_args;

int b200r_loss_bwd(b200r_handle* h, const b200r_loss_bwd_args* args, b200r_stream stream);

/* ------------------------------------------------------------------ quaternion operators (the dqtorch extension)
 * Stand-alone replacements of the reference's only native code, lab4d/third_party/quaternion/src/quaternion.cu:29-217
 * (bindings.cpp:7-16: quaternion_mul_forward / _backward / _backward_backward, quaternion_conjugate), which
 * lab4d/utils/quat_transform.py:36-113 calls on flattened (B, 3|4) operands.  a (B,D1), b (B,D2), D in {3, 4}: a 3-vector is a pure
 * quaternion (w = 0).  fp32, contiguous, 16-B aligned for 4-wide operands.  Outputs are OVERWRITTEN.
 *   fwd:      out (B,4) = a b
 *   bwd:      g_a (B,D1) = [G b*],  g_b (B,D2) = [a* G]                     (G = gradient of out; [.] drops w for 3-vectors)
 *   bwd_bwd:  for cotangents u1 (B,D1), u2 (B,D2) of (g_a, g_b):  g_G (B,4) = u1 b + a u2,  g_a' = [G u2*],  g_b' = [u1* G] */
int b200r_quat_mul_fwd(b200r_handle* h, const float* a, const float* b, float* out, int64_t B, int32_t D1, int32_t D2, b200r_stream stream);
int b200r_quat_mul_bwd(b200r_handle* h, const float* grad, const float* a, const float* b, float* g_a, float* g_b, int64_t B, int32_t D1,
                       int32_t D2, b200r_stream stream);
int b200r_quat_mul_bwd_bwd(b200r_handle* h, const float* u1, const float* u2, const float* grad, const float* a, const float* b, float* g_grad,
                           float* g_a, float* g_b, int64_t B, int32_t D1, int32_t D2, b200r_stream stream);
int b200r_quat_conj(b200r_handle* h, const float* q, float* out, int64_t B, b200r_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* B200R_H */
