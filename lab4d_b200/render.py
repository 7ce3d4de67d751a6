"""Host side of the B200 renderer: per-frame prologue (tiny, M rows), pointer marshalling into the
C ABI, output assembly.  The per-sample work is entirely inside libb200render.so.

Functional mirror of the reference entry points:
  FieldRenderer.query_field  <-> {NeRF,FeatureNeRF,Deformable}.query_field (training mode)
  render_pixel               <-> lab4d.utils.render_utils.render_pixel
  compose_fields             <-> MultiFields.compose_fields
Parameters are read from a dict keyed by the reference's state_dict names (lab4d_b200/spec.py).
"""
import ctypes as C

import torch

from . import _lib, quat
from .spec import INST_CH, T_EMBED_CH, FieldConfig, pe_dim


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class FieldRenderer:
    """One field (fg or bg).  Holds the packed tensor-core operands; everything else is per call."""

    def __init__(self, cfg: FieldConfig, device="cuda", operand_dtype="fp16"):
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("lab4d_b200 renders on CUDA (sm_100a) only; there is no CPU path")
        self.handle = _lib.handle_for(self.device)
        self.desc = _lib.FieldDesc(category=0 if cfg.category == "fg" else 1, D=cfg.D, W=cfg.W, L_xyz=cfg.L_xyz,
                                   L_dir=cfg.L_dir, appr_channels=cfg.appr_channels, skip=cfg.skip,
                                   n_bones=cfg.B if cfg.motion != "rigid" else 0, has_feature=int(cfg.has_feature),
                                   operand_dtype={"fp16": 0, "bf16": 1}[operand_dtype])
        self.n_layers = self.handle.lib.b200r_layer_count(C.byref(self.desc))
        nbytes = self.handle.lib.b200r_packed_bytes(C.byref(self.desc))
        if self.n_layers <= 0 or nbytes == 0:
            raise RuntimeError("b200r: unsupported field configuration")
        self.packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._layers = self._layer_table()
        assert len(self._layers) == self.n_layers

    # canonical layer order of include/b200r.h: (weight name, bias name, conditioning spec)
    def _layer_table(self):
        c = self.cfg
        L = []
        pe_b, pe_c = pe_dim(c.L_xyz), pe_dim(c.L_xyz + 2)
        if c.motion != "rigid":
            p = "warp.skinning_model.delta_field."
            L.append((p + "linear_1.0", "delta1"))
            L.append((p + "linear_2.0", None))
            L.append((p + "linear_final", None))
        L.append(("vis_mlp.basefield.linear_1.0", ("inst_vis", pe_dim(10))))
        L.append(("vis_mlp.basefield.linear_2.0", None))
        for i in range(c.D):
            cond = ("inst_base", pe_b) if i in (0, c.skip) else None
            L.append((f"basefield.linear_{i+1}.0", cond))
        L.append(("basefield.linear_final.0", None))
        L.append(("rgb.0", ("appr_code", c.W + pe_dim(c.L_dir)) if c.appr_channels > 0 else None))
        L.append(("colorfield.linear_1.0", ("inst_color", pe_c)))
        L.append(("colorfield.linear_2.0", None))
        L.append(("colorfield.linear_final.0", None))
        if c.has_feature:
            for i in range(5):
                L.append((f"feature_field.linear_{i+1}.0", None))
            L.append(("feature_field.linear_final", None))
        return L

    def pack(self, P, alpha=None):
        """Convert the nn.Linear weights to 16-bit swizzled UMMA operand tiles (call after every
        optimiser step / set_alpha)."""
        ws = [_f32c(P[name + ".weight"]) for name, _ in self._layers]
        arr = (C.c_void_p * len(ws))(*[w.data_ptr() for w in ws])
        rc = self.handle.lib.b200r_pack_weights(self.handle.h, C.byref(self.desc), arr, len(ws),
                                                C.c_float(-1.0 if alpha is None else float(alpha)),
                                                _ptr(self.packed), self.packed.numel(), _stream(self.device))
        self.handle.check(rc, "b200r_pack_weights")
        self._keep = ws

    # ------------------------------------------------------------------ per-frame prologue (M rows)
    def _bias_rows(self, P, tab, M):
        """b + W[:, code columns] @ code for layers that see a per-frame code
        (nnutils/base.py:140-146, nerf.py:200-204, skinning.py:109-116)."""
        rows, extra = [], {}
        for name, cond in self._layers:
            W, b = P[name + ".weight"], P[name + ".bias"]
            if cond is None:
                rows.append((_f32c(b), 0))
            elif cond == "delta1":
                xb = 3 * self.cfg.B
                Wt, Wi = W[:, xb:xb + T_EMBED_CH], W[:, xb + T_EMBED_CH:xb + T_EMBED_CH + INST_CH]
                inst = tab["inst_skin"] @ Wi.t()
                rows.append((_f32c(b + tab["skin_t_embed"] @ Wt.t() + inst), W.shape[0]))
                extra["delta1_bias_fwd"] = _f32c(b + tab["skin_t_embed_mean"].expand(M, -1) @ Wt.t() + inst)
            else:
                key, col = cond
                code = tab[key]
                rows.append((_f32c(b + code @ W[:, col:col + code.shape[1]].t()), W.shape[0]))
        return rows, extra

    def _bone_tables(self, P, tab):
        """Inverse bone transforms, per-bone blend transforms and Gaussian scales
        (nnutils/warping.py:304-314, utils/transforms.py:9-25, nnutils/skinning.py:141-153)."""
        t_art = (tab["t_articulation_qr"], tab["t_articulation_qd"])
        r_art = (tab["rest_articulation_qr"], tab["rest_articulation_qd"])
        z = torch.zeros_like(t_art[0][..., :1])

        def inv_table(art):
            inv = quat.dq_inv(art)
            return _f32c(torch.cat([inv[0], quat.dq_translation(inv), z], -1))

        def se3_table(dq):
            return _f32c(torch.cat([dq[0], dq[1]], -1))

        lg = P["warp.skinning_model.log_gauss"]
        if self.cfg.symm_idx is not None:
            lg = (lg[list(self.cfg.symm_idx)] + lg) / 2
        inv_gauss = torch.cat([(-lg).exp(), torch.zeros_like(lg[:, :1])], -1)
        center = quat.dq_translation((r_art[0][:1], r_art[1][:1]))[0]
        return dict(bone_inv_t=inv_table(t_art), bone_inv_rest=inv_table(r_art),
                    se3_bwd=se3_table(quat.dq_mul(r_art, quat.dq_inv(t_art))),
                    se3_fwd=se3_table(quat.dq_mul(t_art, quat.dq_inv(r_art))),
                    inv_gauss=_f32c(inv_gauss), bone_center=_f32c(torch.cat([center, torch.zeros_like(center[:, :1])], -1)))

    # ------------------------------------------------------------------ query_field
    @torch.no_grad()
    def query_field(self, P, rays, tab, D, flow_thresh=None, want=None):
        """Training-mode query_field.  rays: hxy (M,N,3), Kinv (M,3,3), near_far (M,2);
        tab: per-frame tables (field2cam_q/t, codes, articulations).  Returns (feat_dict, deltas)
        with the reference's keys and (M,N,D,c) shapes.  `eikonal` is returned as zeros: its
        second-order term stays on PyTorch autograd."""
        c = self.cfg
        hxy = _f32c(rays["hxy"])
        M, N = hxy.shape[:2]
        S = M * N * D
        a = _lib.FieldArgs()
        a.M, a.N, a.D = M, N, int(D)
        a.flow_thresh = -1.0 if flow_thresh is None else float(flow_thresh)
        keep = [hxy]

        def put(name, t):
            t = _f32c(t)
            keep.append(t)
            setattr(a, name, t.data_ptr())

        put("hxy", hxy)
        put("Kinv", rays["Kinv"])
        put("near_far", rays["near_far"])
        q, t = tab["field2cam_q"], tab["field2cam_t"]
        put("field2cam", torch.cat([q, t, torch.zeros_like(t[:, :1])], -1))
        put("logibeta", P["logibeta"])
        put("logscale", P["logscale"])
        rows, extra = self._bias_rows(P, tab, M)
        for i, (row, stride) in enumerate(rows):
            keep.append(row)
            a.bias[i] = row.data_ptr()
            a.bias_stride[i] = stride
        put("sdf_w", P["sdf.weight"].reshape(-1))
        put("sdf_b", P["sdf.bias"])
        put("rgb2_w", P["rgb.2.weight"])
        put("rgb2_b", P["rgb.2.bias"])
        if c.L_dir == 0:
            put("rgb0_dir_w", P["rgb.0.weight"][:, c.W:c.W + 3])
        put("vis_final_w", P["vis_mlp.basefield.linear_final.weight"].reshape(-1))
        put("vis_final_b", P["vis_mlp.basefield.linear_final.bias"])
        if c.motion != "rigid":
            put("delta1_bias_fwd", extra["delta1_bias_fwd"])
            for k, v in self._bone_tables(P, tab).items():
                put(k, v)
            put("warp_logibeta", P["warp.logibeta"])
        out = {}
        for name, nch in _lib.FIELD_OUTPUTS:
            if want is not None and name not in want:
                continue
            if name == "feature" and not c.has_feature:
                continue
            if name == "gauss_density" and c.motion == "rigid":
                continue
            out[name] = torch.empty(S, nch, dtype=torch.float32, device=self.device)
            setattr(a, name, out[name].data_ptr())
        timing = getattr(self, "time_next_launch", False)
        if timing:  # CUDA events on the launching stream, around the C-ABI call only (bench.py roofline)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(self.device))
        rc = self.handle.lib.b200r_field_fwd(self.handle.h, C.byref(self.desc), _ptr(self.packed), C.byref(a),
                                             _stream(self.device))
        self.handle.check(rc, "b200r_field_fwd")
        if timing:
            e1.record(torch.cuda.current_stream(self.device))
            e1.synchronize()
            self.last_kernel_ms = e0.elapsed_time(e1)
            self.time_next_launch = False
        self._keep_call = keep
        feat = {k: v.view(M, N, D, -1) for k, v in out.items()}
        deltas = feat.pop("deltas", None)
        for k in ("xyz_t", "dir", "sdf"):
            feat.pop(k, None)
        self.last_aux = {k: out[k].view(M, N, D, -1) for k in ("xyz_t", "dir", "sdf") if k in out}
        if "density" in feat:
            feat["density_" + c.category] = feat["density"]
        if want is None or "eikonal" in want:
            feat["eikonal"] = torch.zeros(M, N, D, 1, device=self.device)
        return feat, deltas


# ---------------------------------------------------------------------------------------- compositing
KEY_SKIP = ("density", "vis", "flow", "eikonal", "xy_reproj", "xyz_reproj", "gauss_density")
KEY_FREEZE = ("cyc_dist", "xyz_cam", "skin_entropy")


def _channel_plan(field_dict):
    plan = []
    for k, v in field_dict.items():
        if k in ("density",):
            continue
        if k == "vis":
            plan.append((k, _lib.CH_VIS, 2))
        elif k == "flow":
            plan.append((k, _lib.CH_FLOW, 2))
        elif k in ("eikonal", "delta_skin"):
            plan.append((k, _lib.CH_MEAN, 1))
        elif k == "gauss_density":
            plan.append((k, _lib.CH_WEIGHTSUM, 1))
        elif k in ("xy_reproj", "xyz_reproj"):
            continue
        elif k in KEY_FREEZE:
            plan.append((k, _lib.CH_NORM_FROZEN, v.shape[-1]))
        else:
            plan.append((k, _lib.CH_NORM, v.shape[-1]))
    return plan


@torch.no_grad()
def render_pixel(field_dict, deltas):
    """Volume-render per-sample field outputs along rays (utils/render_utils.py:59-184).
    field_dict: key -> (M,N,D,c); deltas (M,N,D,1).  Returns key -> (M,N,c)."""
    dens = _f32c(field_dict["density"])
    device = dens.device
    h = _lib.handle_for(device)
    M, N, D = dens.shape[:3]
    R = M * N
    plan = _channel_plan(field_dict)
    out = {"mask": torch.empty(M, N, 1, device=device)}
    keep = [dens, _f32c(deltas)]
    # the ABI takes at most MAX_CHANNELS arrays per launch
    for c0 in range(0, max(len(plan), 1), _lib.MAX_CHANNELS):
        a = _lib.CompositeArgs()
        a.R, a.D = R, D
        a.density, a.deltas = keep[0].data_ptr(), keep[1].data_ptr()
        a.mask = out["mask"].data_ptr()
        part = plan[c0:c0 + _lib.MAX_CHANNELS]
        a.n_channels = len(part)
        for i, (k, mode, nout) in enumerate(part):
            src = _f32c(field_dict[k])
            keep.append(src)
            dst = torch.empty(M, N, nout, device=device)
            out[k] = dst
            a.src[i], a.dst[i] = src.data_ptr(), dst.data_ptr()
            a.nch[i], a.mode[i] = src.shape[-1], mode
        rc = h.lib.b200r_composite_fwd(h.h, C.byref(a), _stream(device))
        h.check(rc, "b200r_composite_fwd")
    # per-batch normalisers (tiny (M,N) tensors)
    if "vis" in out:
        v = out["vis"]
        out["vis"] = (-(v[..., :1] / D) / (v[..., 1].sum() / (R * D)))
    for k in ("eikonal", "delta_skin"):
        if k in out:
            out[k] = out[k][..., 0]
    if "gauss_density" in out:
        out["gauss_mask"] = out.pop("gauss_density")
    dkeys = [k for k in out if k.startswith("density_")]
    if dkeys:
        dsum = torch.cat([out[k] for k in dkeys], -1).sum(-1, keepdim=True) + 1e-6
        for k in dkeys:
            out[k.replace("density_", "mask_")] = out.pop(k) / dsum
    if "normal" in out:
        out["normal"] = torch.nn.functional.normalize(out["normal"], 2, -1)
    return out


@torch.no_grad()
def compose_fields(feats, deltas_list):
    """MultiFields.compose_fields (nnutils/multifields.py:339-398): concatenate the fields' samples
    along the ray and depth-sort every key.  feats: list of dicts in field order."""
    keys = []
    for f in feats:
        for k in f:
            if k not in keys:
                keys.append(k)
    out = {}
    for k in keys:
        ref = next(f[k] for f in feats if k in f)
        out[k] = torch.cat([f[k] if k in f else torch.zeros_like(ref) for f in feats], 2)
    deltas = torch.cat(deltas_list, 2)
    if len(feats) > 1:
        idx = out["depth"].argsort(2)
        out = {k: torch.gather(v, 2, idx.expand_as(v)) for k, v in out.items()}
        deltas = torch.gather(deltas, 2, idx.expand_as(deltas))
    return out, deltas
