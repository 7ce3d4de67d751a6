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

#define B200R_MAX_LAYERS 24 /* distinct dense layers of one field */

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
  int32_t operand_dtype; /* tensor-core operand type: 0 = fp16, 1 = bf16 (fp32 accumulate) */
} b200r_field_desc;

/* Dense layers, in this canonical order (absent groups are skipped):
 *   [delta_field.linear_1, linear_2, linear_final]        if n_bones > 0
 *   [vis_mlp.basefield.linear_1, linear_2]
 *   [basefield.linear_1 .. linear_D, linear_final]
 *   [rgb.0]
 *   [colorfield.linear_1, linear_2, linear_final]
 *   [feature_field.linear_1 .. linear_5, linear_final]     if has_feature
 * b200r_layer_count() returns how many that is for a descriptor. */
int b200r_layer_count(const b200r_field_desc* desc);

/* Bytes of the packed operand buffer for a descriptor. */
size_t b200r_packed_bytes(const b200r_field_desc* desc);

int b200r_create(int device, b200r_handle** out);
void b200r_destroy(b200r_handle* h);
const char* b200r_last_error(const b200r_handle* h);

/* weights[i]: (out_i, in_i) row-major fp32 nn.Linear weight of canonical layer i.
 * alpha: PosEmbedding annealing window (nnutils/embedding.py:112-125) folded into the packed
 * weights of basefield / colorfield; pass a negative value for "None". */
int b200r_pack_weights(b200r_handle* h, const b200r_field_desc* desc, const float* const* weights, int n_weights,
                       float alpha, void* packed, size_t packed_bytes, b200r_stream stream);

/* Inputs and outputs of one training-mode query_field call on M frames x N rays x D samples. */
typedef struct {
  int32_t M, N, D;
  float flow_thresh;          /* < 0: None */
  const float* hxy;           /* (M,N,3) homogeneous pixel coordinates */
  const float* Kinv;          /* (M,3,3) */
  const float* near_far;      /* (M,2) */
  const float* field2cam;     /* (M,8): quaternion w,x,y,z ; translation x,y,z ; 0 */
  const float* logibeta;      /* (1) */
  const float* logscale;      /* (1) */
  /* per-layer bias rows: bias[i] + frame * bias_stride[i]; stride 0 = shared by all frames.
   * Per-frame rows carry b + W[:, code columns] @ code (instance / time / appearance codes are
   * constant per frame: nnutils/base.py:140-146, nerf.py:200-204, skinning.py:113-116). */
  const float* bias[B200R_MAX_LAYERS];
  int32_t bias_stride[B200R_MAX_LAYERS];
  const float* delta1_bias_fwd; /* (M,64) delta_field.linear_1 rows for FORWARD warps (mean time code) */
  /* heads evaluated on CUDA cores in the epilogues */
  const float* sdf_w;         /* (W) */
  const float* sdf_b;         /* (1) */
  const float* rgb2_w;        /* (3, W/2) */
  const float* rgb2_b;        /* (3) */
  const float* rgb0_dir_w;    /* (W/2, 3) columns of rgb.0 that multiply the raw direction, L_dir == 0 */
  const float* vis_final_w;   /* (64) */
  const float* vis_final_b;   /* (1) */
  /* skinning tables, n_bones > 0 (per frame, computed from the articulation dual quaternions) */
  const float* bone_inv_t;    /* (M,B,8): inverse of t_articulation as rotation q(4), translation(3), 0 */
  const float* bone_inv_rest; /* (M,B,8): inverse of rest_articulation */
  const float* se3_bwd;       /* (M,B,8): rest (x) t^-1 as dual quaternion real(4), dual(4) */
  const float* se3_fwd;       /* (M,B,8): t (x) rest^-1 */
  const float* inv_gauss;     /* (B,4): 1/exp(log_gauss) xyz, 0 */
  const float* bone_center;   /* (B,4): rest bone centres of frame 0, 0 */
  const float* warp_logibeta; /* (1) */
  /* outputs, (M*N*D, c) row-major; any may be NULL */
  float* rgb;                 /* 3 */
  float* density;             /* 1 */
  float* vis;                 /* 1 */
  float* xyz;                 /* 3 canonical */
  float* xyz_cam;             /* 3 */
  float* xyz_t;               /* 3 time-t object space */
  float* dir;                 /* 3 field-space ray direction */
  float* depth;               /* 1 (already divided by exp(logscale)) */
  float* deltas;              /* 1 */
  float* feature;             /* 16 */
  float* flow;                /* 3 */
  float* cyc_dist;            /* 1 */
  float* delta_skin;          /* 1 */
  float* skin_entropy;        /* 1 */
  float* gauss_density;       /* 1 */
  float* sdf;                 /* 1 */
} b200r_field_args;

int b200r_field_fwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed, const b200r_field_args* args,
                    b200r_stream stream);

/* ------------------------------------------------------------------ compositing (render_pixel) */
#define B200R_MAX_CHANNELS 12
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

#ifdef __cplusplus
}
#endif
#endif /* B200R_H */
