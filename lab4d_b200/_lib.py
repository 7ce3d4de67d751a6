"""ctypes binding of libb200render.so (C ABI: include/b200r.h).  No torch types cross this boundary:
only integers, floats and raw device pointers.  Import fails loudly if the library is missing or the
device is not sm_100 - there is no CPU or PyTorch fallback for the hot path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200R_LIB") or os.path.join(HERE, "libb200render.so")  # B200R_LIB: A/B of a build variant (tools/gpu_ab2.sh)

MAX_LAYERS = 32
MAX_CHANNELS = 16
CH_NORM, CH_NORM_FROZEN, CH_MEAN, CH_FLOW, CH_WEIGHTSUM, CH_VIS = range(6)

f32p = C.c_void_p  # device pointers are passed as integers


class FieldDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("category", "D", "W", "L_xyz", "L_dir", "appr_channels", "skip", "n_bones",
                                         "has_feature", "operand_dtype", "dense", "pad_")]


FIELD_OUTPUTS = [("rgb", 3), ("density", 1), ("vis", 1), ("xyz", 3), ("xyz_cam", 3), ("xyz_t", 3), ("dir", 3), ("depth", 1),
                 ("deltas", 1), ("feature", 16), ("flow", 3), ("cyc_dist", 1), ("delta_skin", 1), ("skin_entropy", 1),
                 ("gauss_density", 1), ("sdf", 1), ("feat_norm", 1), ("warp_pts", 9)]


class FieldParams(C.Structure):
    _fields_ = ([("weight", f32p * MAX_LAYERS), ("bias", f32p * MAX_LAYERS)]
                + [(n, f32p) for n in ("sdf_w", "sdf_b", "rgb2_w", "rgb2_b", "vis_final_w", "vis_final_b", "logibeta",
                                       "logscale", "warp_logibeta", "log_gauss", "symm_idx")])


FRAME_PTRS = ["Kinv", "near_far", "field2cam_q", "field2cam_t", "inst_base", "inst_color", "inst_vis", "appr_code",
              "inst_skin", "skin_t_embed", "skin_t_embed_mean", "dense_t_embed", "inst_dense_fwd", "inst_dense_bwd", "t_art_qr", "t_art_qd", "rest_art_qr", "rest_art_qd"]


class FrameTables(C.Structure):
    _fields_ = [("M", C.c_int32), ("pad_", C.c_int32)] + [(n, f32p) for n in FRAME_PTRS]


class RayBatch(C.Structure):
    _fields_ = [("N", C.c_int32), ("D", C.c_int32), ("flow_thresh", C.c_float), ("pad_", C.c_int32), ("hxy", f32p), ("depth", f32p)]


class FieldOutputs(C.Structure):
    _fields_ = [(n, f32p) for n, _ in FIELD_OUTPUTS]


class CompositeArgs(C.Structure):
    _fields_ = [("R", C.c_int32), ("D", C.c_int32), ("density", f32p), ("deltas", f32p), ("mask", f32p),
                ("weights", f32p), ("transmit", f32p), ("n_channels", C.c_int32), ("src", f32p * MAX_CHANNELS),
                ("dst", f32p * MAX_CHANNELS), ("nch", C.c_int32 * MAX_CHANNELS), ("mode", C.c_int32 * MAX_CHANNELS)]


class CompositeBwdArgs(C.Structure):
    _fields_ = [("fwd", CompositeArgs), ("g_mask", f32p), ("g_dst", f32p * MAX_CHANNELS), ("g_density", f32p),
                ("g_src", f32p * MAX_CHANNELS)]


class PointBatch(C.Structure):
    _fields_ = [("P", C.c_int32), ("pad_", C.c_int32), ("xyz", f32p), ("dir", f32p)]


class ImportanceArgs(C.Structure):
    _fields_ = [("R", C.c_int32), ("Dc", C.c_int32), ("depth_c", f32p), ("weights", f32p), ("depth_out", f32p)]


class ComposeArgs(C.Structure):
    _fields_ = [("R", C.c_int32), ("Da", C.c_int32), ("Db", C.c_int32), ("n_channels", C.c_int32), ("depth_a", f32p),
                ("depth_b", f32p), ("perm", C.c_void_p), ("src_a", f32p * MAX_CHANNELS), ("src_b", f32p * MAX_CHANNELS),
                ("dst", f32p * MAX_CHANNELS), ("nch", C.c_int32 * MAX_CHANNELS)]


class ComposeBwdArgs(C.Structure):
    _fields_ = [("R", C.c_int32), ("Da", C.c_int32), ("Db", C.c_int32), ("n_channels", C.c_int32), ("perm", C.c_void_p),
                ("g_dst", f32p * MAX_CHANNELS), ("g_a", f32p * MAX_CHANNELS), ("g_b", f32p * MAX_CHANNELS), ("nch", C.c_int32 * MAX_CHANNELS)]


GRAD_KEYS = ["rgb", "density", "vis", "feature", "xyz", "xyz_cam", "depth", "flow", "cyc_dist", "delta_skin", "skin_entropy",
             "gauss_density"]


class FieldGrads(C.Structure):
    _fields_ = [(n, f32p) for n in GRAD_KEYS]


class Tape(C.Structure):
    _fields_ = [("a", C.c_void_p), ("g", C.c_void_p), ("mask", C.c_void_p), ("a_bytes", C.c_size_t), ("g_bytes", C.c_size_t),
                ("mask_bytes", C.c_size_t)]


MAX_COND = 12


class CondRow(C.Structure):
    _fields_ = [("layer", C.c_int32), ("n", C.c_int32), ("in_dim", C.c_int32), ("frame_off", C.c_int32), ("n_seg", C.c_int32),
                ("col0", C.c_int32 * 2), ("width", C.c_int32 * 2), ("code", C.c_int32 * 2)]


class BlockLayout(C.Structure):
    _fields_ = ([("const_floats", C.c_int32), ("frame_floats", C.c_int32), ("c_plain_bias", C.c_int32 * MAX_LAYERS)]
                + [(n, C.c_int32) for n in ("c_sdf_w", "c_rgb2_w", "c_vis_w", "c_dir_w", "c_center", "c_scalars", "f_cam", "f_cam_partner",
                                            "f_binv_t", "f_se3_bwd", "f_binv_rest", "f_se3_fwd", "f_binv_rest_partner",
                                            "f_se3_fwd_partner", "n_cond")]
                + [("cond", CondRow * MAX_COND)])


HEAD_GRADS = ["sdf_w", "sdf_b", "rgb2_w", "rgb2_b", "vis_final_w", "vis_final_b", "logibeta", "logscale", "warp_logibeta", "log_gauss"]


class ParamGrads(C.Structure):
    _fields_ = ([("flat", f32p), ("weight_off", C.c_int64 * MAX_LAYERS), ("bias_off", C.c_int64 * MAX_LAYERS)]
                + [(n, C.c_int64) for n in HEAD_GRADS] + [("const_block", f32p), ("frame_block", f32p)])


FRAME_GRADS = ["Kinv", "field2cam_q", "field2cam_t", "inst_base", "inst_color", "inst_vis", "appr_code", "inst_skin", "skin_t_embed",
               "skin_t_embed_mean", "dense_t_embed", "inst_dense_fwd", "inst_dense_bwd", "t_art_qr", "t_art_qd", "rest_art_qr", "rest_art_qd"]


class FrameGrads(C.Structure):
    _fields_ = [(n, f32p) for n in FRAME_GRADS]


class EikBatch(C.Structure):
    _fields_ = [("n_rays", C.c_int32), ("pad_", C.c_int32), ("rays", C.c_void_p), ("a", C.c_void_p), ("v", C.c_void_p),
                ("a_bytes", C.c_size_t), ("v_bytes", C.c_size_t)]


LOSS_TERMS = ["mask", "feature", "feat_reproj", "rgb", "depth", "flow", "vis", "reg_gauss_mask"]
LOSS_INPUTS = ["r_mask", "r_mask_fg", "r_rgb", "r_depth", "r_flow", "vis_fg", "vis_bg", "a_feature", "a_xy_reproj", "a_gauss_mask", "b_mask",
               "b_vis2d", "b_is_detected", "b_rgb", "b_depth", "b_flow", "b_flow_uct", "b_feature", "b_hxy"]
LOSS_GRADS = ["g_mask", "g_mask_fg", "g_rgb", "g_depth", "g_flow", "g_vis_fg", "g_vis_bg", "g_feature", "g_xy_reproj", "g_gauss_mask"]


class LossArgs(C.Structure):
    _fields_ = ([("M", C.c_int32), ("N", C.c_int32), ("field_type", C.c_int32), ("train_res", C.c_float)] + [(n, f32p) for n in LOSS_INPUTS]
                + [("wt", C.c_float * 8), ("loss", f32p), ("stats", f32p)])


class LossBwdArgs(C.Structure):
    _fields_ = [("fwd", LossArgs), ("g_loss", f32p)] + [(n, f32p) for n in LOSS_GRADS]


class MatchArgs(C.Structure):
    _fields_ = [("R", C.c_int32), ("K", C.c_int32), ("feat_px", f32p), ("feat_can", f32p), ("xyz_can", f32p), ("idx", C.c_void_p),
                ("logsigma", f32p), ("xyz_matched", f32p), ("lse", f32p)]


class MatchBwdArgs(C.Structure):
    _fields_ = [("fwd", MatchArgs), ("g_out", f32p), ("g_feat_can", f32p), ("g_xyz_can", f32p), ("g_logsigma", f32p), ("scratch", f32p)]


EXPORTS = ["b200r_compose_bwd", "b200r_program_steps", "b200r_field_normals", "b200r_warp_fwd_train", "b200r_warp_bwd", "b200r_quat_mul_fwd", "b200r_quat_mul_bwd", "b200r_quat_mul_bwd_bwd", "b200r_quat_conj", "b200r_loss_fwd", "b200r_loss_bwd", "b200r_match_fwd", "b200r_match_bwd", "b200r_match_scratch_floats", "b200r_eikonal_sizes", "b200r_eikonal_fwd", "b200r_eikonal_bwd", "b200r_tape_sizes", "b200r_field_fwd_train", "b200r_packed_t_bytes", "b200r_pack_weights_t", "b200r_get_block_layout",
           "b200r_field_bwd", "b200r_layer_count", "b200r_packed_bytes", "b200r_create", "b200r_destroy", "b200r_last_error",
           "b200r_pack_weights", "b200r_workspace_bytes", "b200r_field_fwd", "b200r_composite_fwd", "b200r_composite_bwd",
           "b200r_compose_fwd", "b200r_points_fwd", "b200r_warp_fwd", "b200r_importance_fwd"]

_lib = None


def load():
    """dlopen the library (works without a GPU: used by the CPU test that checks the exported symbols)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m lab4d_b200.build` (or __graft_entry__.build()); "
                           "the renderer has no non-CUDA fallback")
    lib = C.CDLL(LIB_PATH)
    lib.b200r_layer_count.argtypes = [C.POINTER(FieldDesc)]
    lib.b200r_layer_count.restype = C.c_int
    lib.b200r_packed_bytes.argtypes = [C.POINTER(FieldDesc)]
    lib.b200r_packed_bytes.restype = C.c_size_t
    lib.b200r_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.b200r_create.restype = C.c_int
    lib.b200r_destroy.argtypes = [C.c_void_p]
    lib.b200r_destroy.restype = None
    lib.b200r_last_error.argtypes = [C.c_void_p]
    lib.b200r_last_error.restype = C.c_char_p
    lib.b200r_pack_weights.argtypes = [C.c_void_p, C.POINTER(FieldDesc), C.POINTER(FieldParams), C.c_float, C.c_void_p,
                                       C.c_size_t, C.c_void_p]
    lib.b200r_pack_weights.restype = C.c_int
    lib.b200r_workspace_bytes.argtypes = [C.POINTER(FieldDesc), C.c_int32]
    lib.b200r_workspace_bytes.restype = C.c_size_t
    lib.b200r_field_fwd.argtypes = [C.c_void_p, C.POINTER(FieldDesc), C.c_void_p, C.POINTER(FieldParams),
                                    C.POINTER(FrameTables), C.POINTER(RayBatch), C.POINTER(FieldOutputs), C.c_void_p,
                                    C.c_size_t, C.c_void_p]
    lib.b200r_field_fwd.restype = C.c_int
    lib.b200r_composite_fwd.argtypes = [C.c_void_p, C.POINTER(CompositeArgs), C.c_void_p]
    lib.b200r_composite_fwd.restype = C.c_int
    lib.b200r_composite_bwd.argtypes = [C.c_void_p, C.POINTER(CompositeBwdArgs), C.c_void_p]
    lib.b200r_composite_bwd.restype = C.c_int
    lib.b200r_points_fwd.argtypes = [C.c_void_p, C.POINTER(FieldDesc), C.c_void_p, C.POINTER(FieldParams),
                                     C.POINTER(FrameTables), C.POINTER(PointBatch), C.POINTER(FieldOutputs), C.c_void_p,
                                     C.c_size_t, C.c_void_p]
    lib.b200r_points_fwd.restype = C.c_int
    lib.b200r_warp_fwd.argtypes = [C.c_void_p, C.POINTER(FieldDesc), C.c_void_p, C.POINTER(FieldParams), C.POINTER(FrameTables),
                                   C.POINTER(PointBatch), C.c_int32, C.POINTER(FieldOutputs), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.b200r_warp_fwd.restype = C.c_int
    lib.b200r_importance_fwd.argtypes = [C.c_void_p, C.POINTER(ImportanceArgs), C.c_void_p]
    lib.b200r_importance_fwd.restype = C.c_int
    lib.b200r_compose_fwd.argtypes = [C.c_void_p, C.POINTER(ComposeArgs), C.c_void_p]
    lib.b200r_compose_fwd.restype = C.c_int
    lib.b200r_tape_sizes.argtypes = [C.POINTER(FieldDesc), C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                     C.POINTER(C.c_size_t)]
    lib.b200r_tape_sizes.restype = C.c_int
    lib.b200r_field_fwd_train.argtypes = [C.c_void_p, C.POINTER(FieldDesc), C.c_void_p, C.POINTER(FieldParams), C.POINTER(FrameTables),
                                          C.POINTER(RayBatch), C.POINTER(FieldOutputs), C.POINTER(Tape), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.b200r_field_fwd_train.restype = C.c_int
    lib.b200r_packed_t_bytes.argtypes = [C.POINTER(FieldDesc)]
    lib.b200r_packed_t_bytes.restype = C.c_size_t
    lib.b200r_pack_weights_t.argtypes = lib.b200r_pack_weights.argtypes
    lib.b200r_pack_weights_t.restype = C.c_int
    lib.b200r_get_block_layout.argtypes = [C.POINTER(FieldDesc), C.POINTER(BlockLayout)]
    lib.b200r_get_block_layout.restype = C.c_int
    lib.b200r_field_bwd.argtypes = [C.c_void_p, C.POINTER(FieldDesc), C.c_void_p, C.POINTER(FieldParams), C.POINTER(FrameTables),
                                    C.POINTER(RayBatch), C.POINTER(FieldOutputs), C.POINTER(FieldGrads), C.POINTER(Tape),
                                    C.POINTER(ParamGrads), C.POINTER(FrameGrads), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.b200r_field_bwd.restype = C.c_int
    lib.b200r_compose_bwd.argtypes = [C.c_void_p, C.POINTER(ComposeBwdArgs), C.c_void_p]
    lib.b200r_compose_bwd.restype = C.c_int
    lib.b200r_program_steps.argtypes = [C.POINTER(FieldDesc), C.c_int32]
    lib.b200r_program_steps.restype = C.c_int
    lib.b200r_field_normals.argtypes = [C.c_void_p, C.POINTER(FieldDesc), C.c_void_p, C.POINTER(FieldParams), C.POINTER(FrameTables),
                                        C.POINTER(RayBatch), C.POINTER(FieldOutputs), C.POINTER(Tape), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.b200r_field_normals.restype = C.c_int
    lib.b200r_warp_fwd_train.argtypes = [C.c_void_p, C.POINTER(FieldDesc), C.c_void_p, C.POINTER(FieldParams), C.POINTER(FrameTables),
                                         C.POINTER(PointBatch), C.POINTER(FieldOutputs), C.POINTER(Tape), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.b200r_warp_fwd_train.restype = C.c_int
    lib.b200r_warp_bwd.argtypes = [C.c_void_p, C.POINTER(FieldDesc), C.c_void_p, C.POINTER(FieldParams), C.POINTER(FrameTables),
                                   C.POINTER(PointBatch), C.POINTER(FieldOutputs), C.c_void_p, C.POINTER(Tape), C.POINTER(ParamGrads),
                                   C.POINTER(FrameGrads), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.b200r_warp_bwd.restype = C.c_int
    P, I64, I32 = C.c_void_p, C.c_int64, C.c_int32
    lib.b200r_quat_mul_fwd.argtypes = [P, P, P, P, I64, I32, I32, P]
    lib.b200r_quat_mul_bwd.argtypes = [P, P, P, P, P, P, I64, I32, I32, P]
    lib.b200r_quat_mul_bwd_bwd.argtypes = [P, P, P, P, P, P, P, P, P, I64, I32, I32, P]
    lib.b200r_quat_conj.argtypes = [P, P, P, I64, P]
    for fn in (lib.b200r_quat_mul_fwd, lib.b200r_quat_mul_bwd, lib.b200r_quat_mul_bwd_bwd, lib.b200r_quat_conj):
        fn.restype = C.c_int
    lib.b200r_loss_fwd.argtypes = [C.c_void_p, C.POINTER(LossArgs), C.c_void_p]
    lib.b200r_loss_fwd.restype = C.c_int
    lib.b200r_loss_bwd.argtypes = [C.c_void_p, C.POINTER(LossBwdArgs), C.c_void_p]
    lib.b200r_loss_bwd.restype = C.c_int
    lib.b200r_match_fwd.argtypes = [C.c_void_p, C.POINTER(MatchArgs), C.c_void_p]
    lib.b200r_match_fwd.restype = C.c_int
    lib.b200r_match_bwd.argtypes = [C.c_void_p, C.POINTER(MatchBwdArgs), C.c_void_p]
    lib.b200r_match_bwd.restype = C.c_int
    lib.b200r_match_scratch_floats.argtypes = [C.c_int32, C.c_int32]
    lib.b200r_match_scratch_floats.restype = C.c_size_t
    lib.b200r_eikonal_sizes.argtypes = [C.POINTER(FieldDesc), C.c_int32, C.c_int32, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.b200r_eikonal_sizes.restype = C.c_int
    lib.b200r_eikonal_fwd.argtypes = [C.c_void_p, C.POINTER(FieldDesc), C.c_void_p, C.POINTER(FieldParams), C.POINTER(RayBatch), C.c_int32,
                                      C.c_void_p, C.POINTER(Tape), C.POINTER(EikBatch), C.c_void_p, C.c_void_p]
    lib.b200r_eikonal_fwd.restype = C.c_int
    lib.b200r_eikonal_bwd.argtypes = [C.c_void_p, C.POINTER(FieldDesc), C.c_void_p, C.POINTER(FieldParams), C.POINTER(RayBatch), C.c_int32,
                                      C.c_void_p, C.POINTER(Tape), C.POINTER(EikBatch), C.c_void_p, C.POINTER(ParamGrads), C.c_void_p]
    lib.b200r_eikonal_bwd.restype = C.c_int
    _lib = lib
    return lib


class Handle:
    """Per-device library handle; raises RuntimeError(b200r_last_error) on any failure."""

    def __init__(self, device_index):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.b200r_create(int(device_index), C.byref(h))
        if rc != 0:
            raise RuntimeError({-3: "b200r: device is not sm_100 (B200); no fallback path exists",
                                -2: "b200r: CUDA error while opening the device"}.get(rc, f"b200r_create failed ({rc})"))
        self.h = h
        self.device_index = int(device_index)

    def check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {self.lib.b200r_last_error(self.h).decode()} (code {rc})")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.b200r_destroy(self.h)
                self.h = None
        except Exception:
            pass


_handles = {}


def handle_for(device):
    import torch

    idx = torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _handles:
        _handles[idx] = Handle(idx)
    return _handles[idx]
