"""Host side of the B200 renderer: pointer marshalling into the C ABI and output assembly.  All
arithmetic (per-frame prologue and per-sample work) is inside libb200render.so.

Functional mirror of the reference entry points:
  FieldRenderer.query_field  <-> {NeRF,FeatureNeRF,Deformable}.query_field (training mode)
  render_pixel               <-> lab4d.utils.render_utils.render_pixel
  compose_fields             <-> MultiFields.compose_fields
Parameters are read from a dict keyed by the reference's state_dict names (lab4d_b200/spec.py).
"""
import ctypes as C

import torch

from . import _lib
from .spec import FieldConfig, field_param_shapes, pe_dim


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


class HostStage:
    """Host-resident inputs of a step (ray batch + per-frame tables) in ONE pinned arena with a device mirror: a step's
    host->device transfer is a single asynchronous copy instead of one per tensor.  `host` / `dev` are dicts of views
    into the two arenas with the caller's keys, shapes and dtypes."""

    def __init__(self, tensors, device):
        offs, total = {}, 0
        for k, v in tensors.items():
            offs[k] = total
            total += (v.numel() * v.element_size() + 255) // 256 * 256
        self.host_arena = torch.empty(max(total, 256), dtype=torch.uint8).pin_memory()
        self.dev_arena = torch.empty(max(total, 256), dtype=torch.uint8, device=device)
        self.nbytes = sum(v.numel() * v.element_size() for v in tensors.values())

        def views(arena):
            return {k: arena[offs[k]:offs[k] + v.numel() * v.element_size()].view(v.dtype).view(v.shape) for k, v in tensors.items()}

        self.host, self.dev = views(self.host_arena), views(self.dev_arena)
        for k, v in tensors.items():
            self.host[k].copy_(v)

    def upload(self):
        """Enqueue the host->device copy on the current stream; returns the device views."""
        self.dev_arena.copy_(self.host_arena, non_blocking=True)
        return self.dev


def importance_merge(handle, device, depth_c, weights):
    """b200r_importance_fwd: coarse depths (R,Dc) + their compositing weights (R,Dc) -> merged ascending depths (R,2Dc)."""
    R, Dc = depth_c.shape
    out = torch.empty(R, 2 * Dc, device=device)
    a = _lib.ImportanceArgs()
    a.R, a.Dc = R, Dc
    a.depth_c, a.weights, a.depth_out = depth_c.data_ptr(), weights.data_ptr(), out.data_ptr()
    handle.check(handle.lib.b200r_importance_fwd(handle.h, C.byref(a), _stream(device)), "b200r_importance_fwd")
    return out


class FieldRenderer:
    """One field (fg or bg).  Holds the packed tensor-core operands; everything else is per call.
    operand_dtype: "fp16x3" = split operands (fp16 head + tail, three MMAs per k-step, ~fp32 results: the parity mode),
    "fp16" / "bf16" = single 16-bit operands (the fast modes)."""

    def __init__(self, cfg: FieldConfig, device="cuda", operand_dtype="fp16"):
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("lab4d_b200 renders on CUDA (sm_100a) only; there is no CPU path")
        self.handle = _lib.handle_for(self.device)
        self.desc = _lib.FieldDesc(category=0 if cfg.category == "fg" else 1, D=cfg.D, W=cfg.W, L_xyz=cfg.L_xyz,
                                   L_dir=cfg.L_dir, appr_channels=cfg.appr_channels, skip=cfg.skip,
                                   n_bones=cfg.B if cfg.motion != "rigid" else 0, has_feature=int(cfg.has_feature),
                                   operand_dtype={"fp16": 0, "bf16": 1, "fp16x3": 2}[operand_dtype],
                                   dense=int(cfg.dense and cfg.motion != "rigid"))
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
        if c.dense and c.motion != "rigid":
            for m in ("forward_map", "backward_map"):
                L.append((f"warp.post_warp.{m}.linear_1.0", None))
                L.append((f"warp.post_warp.{m}.linear_2.0", None))
                L.append((f"warp.post_warp.{m}.linear_final", None))
        return L

    def _params(self, P):
        """Pointer table into the caller's parameter storage (no copies unless a tensor is non-contiguous)."""
        par = _lib.FieldParams()
        keep = []

        def ptr(t):
            t = _f32c(t)
            keep.append(t)
            return t.data_ptr()

        for i, (name, _) in enumerate(self._layers):
            par.weight[i] = ptr(P[name + ".weight"])
            par.bias[i] = ptr(P[name + ".bias"])
        par.sdf_w, par.sdf_b = ptr(P["sdf.weight"]), ptr(P["sdf.bias"])
        par.rgb2_w, par.rgb2_b = ptr(P["rgb.2.weight"]), ptr(P["rgb.2.bias"])
        par.vis_final_w = ptr(P["vis_mlp.basefield.linear_final.weight"])
        par.vis_final_b = ptr(P["vis_mlp.basefield.linear_final.bias"])
        par.logibeta, par.logscale = ptr(P["logibeta"]), ptr(P["logscale"])
        if self.cfg.motion != "rigid":
            par.warp_logibeta = ptr(P["warp.logibeta"])
            par.log_gauss = ptr(P["warp.skinning_model.log_gauss"])
            if self.cfg.symm_idx is not None:
                if getattr(self, "_symm", None) is None:
                    self._symm = torch.tensor(list(self.cfg.symm_idx), dtype=torch.int32, device=self.device)
                par.symm_idx = self._symm.data_ptr()
        return par, keep

    def pack(self, P, alpha=None):
        """Convert the nn.Linear weights to 16-bit swizzled UMMA operand tiles (call after every
        optimiser step / set_alpha)."""
        par, keep = self._params(P)
        rc = self.handle.lib.b200r_pack_weights(self.handle.h, C.byref(self.desc), C.byref(par),
                                                C.c_float(-1.0 if alpha is None else float(alpha)),
                                                _ptr(self.packed), self.packed.numel(), _stream(self.device))
        self.handle.check(rc, "b200r_pack_weights")
        self._keep = keep

    # names of the per-frame tables (keys of `tab`) in the order of b200r_frame_tables
    _TAB_KEYS = {"inst_base": "inst_base", "inst_color": "inst_color", "inst_vis": "inst_vis", "appr_code": "appr_code",
                 "inst_skin": "inst_skin", "skin_t_embed": "skin_t_embed", "skin_t_embed_mean": "skin_t_embed_mean",
                 "dense_t_embed": "dense_t_embed", "inst_dense_fwd": "inst_dense_fwd", "inst_dense_bwd": "inst_dense_bwd",
                 "t_art_qr": "t_articulation_qr", "t_art_qd": "t_articulation_qd", "rest_art_qr": "rest_articulation_qr",
                 "rest_art_qd": "rest_articulation_qd", "field2cam_q": "field2cam_q", "field2cam_t": "field2cam_t"}

    # ------------------------------------------------------------------ query_field
    @torch.no_grad()
    def query_field(self, P, rays, tab, D, flow_thresh=None, want=None, depth=None):
        """Training-mode query_field.  rays: hxy (M,N,3), Kinv (M,3,3), near_far (M,2);
        tab: per-frame tables (field2cam_q/t, codes, articulations).  Returns (feat_dict, deltas)
        with the reference's keys and (M,N,D,c) shapes.  `eikonal` is returned as zeros: its
        second-order term stays on PyTorch autograd.  depth (M,N,D[,1]): given ascending sample
        depths (e.g. from `importance_depths`) instead of the uniform placement."""
        c = self.cfg
        hxy = _f32c(rays["hxy"])
        M, N = hxy.shape[:2]
        S = M * N * D
        keep = [hxy]
        par, kp = self._params(P)
        keep += kp
        fr = _lib.FrameTables()
        fr.M = M

        def put(obj, name, t):
            t = _f32c(t)
            keep.append(t)
            setattr(obj, name, t.data_ptr())

        put(fr, "Kinv", rays["Kinv"])
        put(fr, "near_far", rays["near_far"])
        for field, key in self._TAB_KEYS.items():
            if key in tab and tab[key] is not None:
                put(fr, field, tab[key])
        rb = _lib.RayBatch()
        rb.N, rb.D = N, int(D)
        rb.flow_thresh = -1.0 if flow_thresh is None else float(flow_thresh)
        rb.hxy = hxy.data_ptr()
        if depth is not None:
            dep = _f32c(depth.reshape(M, N, int(D)))
            keep.append(dep)
            rb.depth = dep.data_ptr()
        out, oa = {}, _lib.FieldOutputs()
        for name, nch in _lib.FIELD_OUTPUTS:
            if want is not None and name not in want:
                continue
            if name in ("feature", "feat_norm") and not c.has_feature:
                continue
            if name == "gauss_density" and c.motion == "rigid":
                continue
            if name == "warp_pts":
                continue
            out[name] = torch.empty(S, nch, dtype=torch.float32, device=self.device)
            setattr(oa, name, out[name].data_ptr())
        wbytes = self.handle.lib.b200r_workspace_bytes(C.byref(self.desc), M)
        if getattr(self, "_ws", None) is None or self._ws.numel() < wbytes:
            self._ws = torch.empty(wbytes, dtype=torch.uint8, device=self.device)
        timing = getattr(self, "time_next_launch", False)
        if timing:  # CUDA events on the launching stream, around the C-ABI call only (bench.py roofline)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(self.device))
        rc = self.handle.lib.b200r_field_fwd(self.handle.h, C.byref(self.desc), _ptr(self.packed), C.byref(par), C.byref(fr),
                                             C.byref(rb), C.byref(oa), _ptr(self._ws), self._ws.numel(), _stream(self.device))
        self.handle.check(rc, "b200r_field_fwd")
        if timing:
            e1.record(torch.cuda.current_stream(self.device))
            e1.synchronize()
            self.last_kernel_ms = e0.elapsed_time(e1)
            self.time_next_launch = False
        self._keep_call = keep
        feat = {k: v.view(M, N, D, -1) for k, v in out.items()}
        deltas = feat.pop("deltas", None)
        for k in ("xyz_t", "dir", "sdf", "feat_norm", "warp_pts"):
            feat.pop(k, None)
        self.last_aux = {k: out[k].view(M, N, D, -1) for k in ("xyz_t", "dir", "sdf", "feat_norm") if k in out}
        if "density" in feat:
            feat["density_" + c.category] = feat["density"]
        if want is None or "eikonal" in want:
            feat["eikonal"] = torch.zeros(M, N, D, 1, device=self.device)
        return feat, deltas


    # ------------------------------------------------------------------ training: forward with a tape, backward
    # state_dict name of the parameter behind every head offset of b200r_param_grads
    _HEAD_NAMES = {"sdf_w": "sdf.weight", "sdf_b": "sdf.bias", "rgb2_w": "rgb.2.weight", "rgb2_b": "rgb.2.bias",
                   "vis_final_w": "vis_mlp.basefield.linear_final.weight", "vis_final_b": "vis_mlp.basefield.linear_final.bias",
                   "logibeta": "logibeta", "logscale": "logscale", "warp_logibeta": "warp.logibeta",
                   "log_gauss": "warp.skinning_model.log_gauss"}

    def _train_state(self):
        """Layout of the flat gradient buffer (every hot-path parameter, in the order of spec.field_param_shapes; 16-B
        aligned slots), block layouts and the transposed operand buffer (built once)."""
        st = getattr(self, "_train", None)
        if st is not None:
            return st
        lib, h = self.handle.lib, self.handle
        layout = _lib.BlockLayout()
        h.check(lib.b200r_get_block_layout(C.byref(self.desc), C.byref(layout)), "b200r_get_block_layout")
        shapes = field_param_shapes(self.cfg)
        slots, total = {}, 0
        for name, shp in shapes.items():
            n = 1
            for d in shp:
                n *= d
            slots[name] = (total, n, tuple(shp))
            total += (n + 3) // 4 * 4
        nbytes = lib.b200r_packed_t_bytes(C.byref(self.desc))
        st = dict(layout=layout, slots=slots, total=total, shapes=shapes,
                  packed_t=torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.device))
        self._train = st
        return st

    def grad_buffer(self):
        """The renderer's persistent flat gradient buffer and name -> view (param-shaped) into it.  The backward ACCUMULATES
        into it (DDP's gradient_as_bucket_view arrangement: parameters' .grad can be these views, the step's one all-reduce
        runs on the flat buffer)."""
        st = self._train_state()
        if "flat" not in st:
            st["flat"] = torch.zeros(st["total"], device=self.device)
            st["views"] = {k: st["flat"][o:o + n].view(shp) for k, (o, n, shp) in st["slots"].items()}
        return st["flat"], st["views"]

    def pack_train(self, P, alpha=None):
        """pack() plus the transposed (W^T) operand tiles the backward's data-gradient GEMMs read."""
        self.pack(P, alpha)
        st = self._train_state()
        par, keep = self._params(P)
        rc = self.handle.lib.b200r_pack_weights_t(self.handle.h, C.byref(self.desc), C.byref(par),
                                                  C.c_float(-1.0 if alpha is None else float(alpha)), _ptr(st["packed_t"]),
                                                  st["packed_t"].numel(), _stream(self.device))
        self.handle.check(rc, "b200r_pack_weights_t")
        self._keep_t = keep
        self._alpha = alpha

    def _tape(self, M, N, D, slot="field"):
        """Tape buffers of one training forward; `slot` keeps the field's tape and the point-warp's tape apart."""
        a, g, m = C.c_size_t(), C.c_size_t(), C.c_size_t()
        self.handle.check(self.handle.lib.b200r_tape_sizes(C.byref(self.desc), M, N, D, C.byref(a), C.byref(g), C.byref(m)), "b200r_tape_sizes")
        need = (a.value, g.value, m.value)
        store = self.__dict__.setdefault("_tape_slots", {})
        bufs = store.get(slot)
        if bufs is None or any(b.numel() < n + 1024 for b, n in zip(bufs, need)):
            bufs = tuple(torch.empty(n + 1024, dtype=torch.uint8, device=self.device) for n in need)
            store[slot] = bufs
        t = _lib.Tape()
        al = lambda b: (b.data_ptr() + 1023) // 1024 * 1024
        t.a, t.g, t.mask = al(bufs[0]), al(bufs[1]), al(bufs[2])
        t.a_bytes, t.g_bytes, t.mask_bytes = need
        return t

    def _frame_tables(self, rays, tab, keep):
        fr = _lib.FrameTables()
        fr.M = rays["hxy"].shape[0]

        def put(name, t):
            t = _f32c(t)
            keep.append(t)
            setattr(fr, name, t.data_ptr())

        put("Kinv", rays["Kinv"])
        put("near_far", rays["near_far"])
        for field, key in self._TAB_KEYS.items():
            if key in tab and tab[key] is not None:
                put(field, tab[key])
        return fr

    @torch.no_grad()
    def query_field_train(self, P, rays, tab, D, flow_thresh=None, depth=None):
        """Training forward: query_field that also records the tape.  Returns (feat_dict, deltas, ctx); `ctx` goes to
        `backward`.  One call may be in flight per renderer (the tape buffers belong to the renderer)."""
        c = self.cfg
        hxy = _f32c(rays["hxy"])
        M, N = hxy.shape[:2]
        S = M * N * int(D)
        keep = [hxy]
        par, kp = self._params(P)
        keep += kp
        fr = self._frame_tables(rays, tab, keep)
        rb = _lib.RayBatch()
        rb.N, rb.D = N, int(D)
        rb.flow_thresh = -1.0 if flow_thresh is None else float(flow_thresh)
        rb.hxy = hxy.data_ptr()
        if depth is not None:
            dep = _f32c(depth.reshape(M, N, int(D)))
            keep.append(dep)
            rb.depth = dep.data_ptr()
        out, oa = {}, _lib.FieldOutputs()
        for name, nch in _lib.FIELD_OUTPUTS:
            if name in ("feature", "feat_norm") and not c.has_feature:
                continue
            if name == "gauss_density" and c.motion == "rigid":
                continue
            if name == "warp_pts" and not (c.dense and c.motion != "rigid"):
                continue
            out[name] = torch.empty(S, nch, dtype=torch.float32, device=self.device)
            setattr(oa, name, out[name].data_ptr())
        wbytes = self.handle.lib.b200r_workspace_bytes(C.byref(self.desc), M)
        if getattr(self, "_ws", None) is None or self._ws.numel() < wbytes:
            self._ws = torch.empty(wbytes, dtype=torch.uint8, device=self.device)
        tape = self._tape(M, N, int(D))
        rc = self.handle.lib.b200r_field_fwd_train(self.handle.h, C.byref(self.desc), _ptr(self.packed), C.byref(par), C.byref(fr),
                                                   C.byref(rb), C.byref(oa), C.byref(tape), _ptr(self._ws), self._ws.numel(),
                                                   _stream(self.device))
        self.handle.check(rc, "b200r_field_fwd_train")
        feat = {k: v.view(M, N, int(D), -1) for k, v in out.items()}
        deltas = feat.pop("deltas")
        for k in ("xyz_t", "dir", "sdf", "feat_norm", "warp_pts"):
            feat.pop(k, None)
        feat["density_" + c.category] = feat["density"]
        feat["eikonal"] = torch.zeros(M, N, int(D), 1, device=self.device)
        ctx = dict(out=out, par=par, fr=fr, rb=rb, tape=tape, keep=keep, M=M, N=N, D=int(D), P=P, tab=tab, rays=rays)
        return feat, deltas, ctx

    @torch.no_grad()
    def backward(self, ctx, grads, accumulate=False):
        """grads: key -> cotangent of the per-sample output (M,N,D,c) (keys of _lib.GRAD_KEYS; missing = zero; `density_fg` /
        `density_bg` are added to `density`).  One C-ABI call: data-gradient kernel, weight-gradient kernel, backward of the
        per-frame prologue.  Parameter gradients land in the flat buffer (`grad_buffer()`; zeroed first unless
        `accumulate`).  Returns (param_grads, table_grads): name -> view / tensor."""
        st = self._train_state()
        flat, views = self.grad_buffer()
        if not accumulate:
            flat.zero_()
        M = ctx["M"]
        keep = []
        fg = _lib.FieldGrads()
        gd = dict(grads)
        for k in ("density_fg", "density_bg"):
            if gd.get(k) is not None:
                gd["density"] = gd[k] if gd.get("density") is None else gd["density"] + gd[k]
        for k in _lib.GRAD_KEYS:
            if gd.get(k) is not None:
                t = _f32c(gd[k])
                keep.append(t)
                setattr(fg, k, t.data_ptr())
        saved, out = _lib.FieldOutputs(), ctx["out"]
        for k in ("xyz", "rgb", "sdf", "feature", "feat_norm", "warp_pts"):
            if k in out:
                setattr(saved, k, out[k].data_ptr())
        layout, slots = st["layout"], st["slots"]
        g_const = torch.empty(layout.const_floats, device=self.device)
        g_frame = torch.empty(M, layout.frame_floats, device=self.device)
        pgs = _lib.ParamGrads()
        pgs.flat, pgs.const_block, pgs.frame_block = flat.data_ptr(), g_const.data_ptr(), g_frame.data_ptr()
        for i in range(_lib.MAX_LAYERS):
            pgs.weight_off[i], pgs.bias_off[i] = -1, -1
        for i, (name, _) in enumerate(self._layers):
            pgs.weight_off[i], pgs.bias_off[i] = slots[name + ".weight"][0], slots[name + ".bias"][0]
        for fld, name in self._HEAD_NAMES.items():
            setattr(pgs, fld, slots[name][0] if name in slots else -1)
        # gradients of the per-frame inputs: one tensor per table the call received
        tg, fgr, tab, rays = {}, _lib.FrameGrads(), ctx["tab"], ctx["rays"]
        tg["Kinv"] = torch.empty(M, 3, 3, device=self.device)
        fgr.Kinv = tg["Kinv"].data_ptr()
        for field, key in self._TAB_KEYS.items():
            if tab.get(key) is not None and field in _lib.FRAME_GRADS:
                tg[key] = torch.empty(tab[key].shape, device=self.device)
                setattr(fgr, field, tg[key].data_ptr())
        alpha = getattr(self, "_alpha", None)
        wnames = self._window_names() if alpha is not None else []
        before = {nl: views[nl[0]].clone() for nl in wnames}
        rc = self.handle.lib.b200r_field_bwd(self.handle.h, C.byref(self.desc), _ptr(st["packed_t"]), C.byref(ctx["par"]), C.byref(ctx["fr"]),
                                             C.byref(ctx["rb"]), C.byref(saved), C.byref(fg), C.byref(ctx["tape"]), C.byref(pgs), C.byref(fgr),
                                             _ptr(self._ws), self._ws.numel(), _stream(self.device))
        self.handle.check(rc, "b200r_field_bwd")
        if alpha is not None:  # the annealing window is folded into the packed weights: this call's dW = dW_eff * window
            self._apply_window(views, before, alpha)
        self.last_blocks = (g_const, g_frame)
        self._keep_bwd = keep
        return views, tg

    # ------------------------------------------------------------------ eval-mode normals (NeRF.compute_normal)
    @torch.no_grad()
    def sdf_gradient_cam(self, ctx):
        """d sdf / d xyz_cam (M,N,D,3) at every sample of the training-form forward that produced `ctx` (query_field_train, e.g.
        with importance-sampled depths): the gradient NeRF.compute_normal (nnutils/nerf.py:455-493) takes with autograd through
        the backward warp and the basefield - b200r_field_normals.  eikonal = (|g| - 1)^2, normal = g / |g| * (1, -1, -1)."""
        st = self._train_state()
        M, N, D = ctx["M"], ctx["N"], ctx["D"]
        g = torch.empty(M * N * D, 3, device=self.device)
        saved, out = _lib.FieldOutputs(), ctx["out"]
        for k in ("xyz", "warp_pts"):
            if k in out:
                setattr(saved, k, out[k].data_ptr())
        rc = self.handle.lib.b200r_field_normals(self.handle.h, C.byref(self.desc), _ptr(st["packed_t"]), C.byref(ctx["par"]), C.byref(ctx["fr"]),
                                                 C.byref(ctx["rb"]), C.byref(saved), C.byref(ctx["tape"]), g.data_ptr(), _ptr(self._ws),
                                                 self._ws.numel(), _stream(self.device))
        self.handle.check(rc, "b200r_field_normals")
        return g.view(M, N, D, 3)

    # ------------------------------------------------------------------ forward warp of points, differentiable (forward_project)
    def warp_weight_names(self):
        """Parameters the forward warp of points reaches: the skinning delta MLP, the Gaussian bone scales and, for a
        ComposedWarp, the forward soft-deformation map."""
        names = ["warp.skinning_model.log_gauss"]
        for lyr in ("linear_1.0", "linear_2.0", "linear_final"):
            names += [f"warp.skinning_model.delta_field.{lyr}.weight", f"warp.skinning_model.delta_field.{lyr}.bias"]
        if self.cfg.dense:
            for lyr in ("linear_1.0", "linear_2.0", "linear_final"):
                names += [f"warp.post_warp.forward_map.{lyr}.weight", f"warp.post_warp.forward_map.{lyr}.bias"]
        return names

    @torch.no_grad()
    def warp_points_train(self, P, xyz, tab):
        """SkinningWarp / ComposedWarp forward (canonical -> time-t space, the frame's own articulation; nnutils/warping.py:277-336,
        445-483 with backward=False) on points xyz (M,P,3) with a tape: returns (xyz' (M,P,3), ctx) for `warp_backward`.
        Call `pack_train` first.  What FeatureNeRF.forward_project (feature.py:207-226) runs on the matched points.  The warp's
        tape belongs to the renderer (slot "warp", apart from the field's): one point warp may be in flight per renderer."""
        if self.cfg.motion == "rigid":
            raise RuntimeError("warp_points_train: the field has no skinning warp")
        xyz = _f32c(xyz)
        M, Pn = xyz.shape[:2]
        par, keep = self._params(P)
        keep.append(xyz)
        fr = _lib.FrameTables()
        fr.M = M
        for field, key in self._TAB_KEYS.items():
            if key in tab and tab[key] is not None and not field.startswith("field2cam"):
                t = _f32c(tab[key])
                keep.append(t)
                setattr(fr, field, t.data_ptr())
        pb = _lib.PointBatch()
        pb.P, pb.xyz = Pn, xyz.data_ptr()
        out, oa = {}, _lib.FieldOutputs()
        for name, nch in (("xyz", 3),) + ((("warp_pts", 9),) if self.cfg.dense else ()):
            out[name] = torch.empty(M * Pn, nch, dtype=torch.float32, device=self.device)
            setattr(oa, name, out[name].data_ptr())
        wbytes = self.handle.lib.b200r_workspace_bytes(C.byref(self.desc), M)
        if getattr(self, "_ws", None) is None or self._ws.numel() < wbytes:
            self._ws = torch.empty(wbytes, dtype=torch.uint8, device=self.device)
        tape = self._tape(M, Pn, 1, slot="warp")
        rc = self.handle.lib.b200r_warp_fwd_train(self.handle.h, C.byref(self.desc), _ptr(self.packed), C.byref(par), C.byref(fr), C.byref(pb),
                                                  C.byref(oa), C.byref(tape), _ptr(self._ws), self._ws.numel(), _stream(self.device))
        self.handle.check(rc, "b200r_warp_fwd_train")
        ctx = dict(out=out, par=par, fr=fr, pb=pb, tape=tape, keep=keep, M=M, Pn=Pn, tab=tab)
        return out["xyz"].view(M, Pn, 3), ctx

    @torch.no_grad()
    def warp_backward(self, ctx, g_xyz, flat=None):
        """Backward of `warp_points_train` for the cotangent g_xyz (M,P,3) of the warped points (b200r_warp_bwd): returns
        (g_points (M,P,3), name -> view of the flat gradient buffer `flat` (ACCUMULATED; default: a fresh zero buffer),
        table gradients name -> tensor)."""
        st = self._train_state()
        if flat is None:
            flat = torch.zeros(st["total"], device=self.device)
        M, Pn = ctx["M"], ctx["Pn"]
        layout, slots = st["layout"], st["slots"]
        gx = _f32c(g_xyz.reshape(M * Pn, 3))
        g_pts = torch.empty(M * Pn, 3, device=self.device)
        g_const = torch.empty(layout.const_floats, device=self.device)
        g_frame = torch.empty(M, layout.frame_floats, device=self.device)
        pgs = _lib.ParamGrads()
        pgs.flat, pgs.const_block, pgs.frame_block = flat.data_ptr(), g_const.data_ptr(), g_frame.data_ptr()
        for i in range(_lib.MAX_LAYERS):
            pgs.weight_off[i], pgs.bias_off[i] = -1, -1
        for i, (name, _) in enumerate(self._layers):
            pgs.weight_off[i], pgs.bias_off[i] = slots[name + ".weight"][0], slots[name + ".bias"][0]
        for fld, name in self._HEAD_NAMES.items():
            setattr(pgs, fld, slots[name][0] if name in slots else -1)
        tg, fgr, tab = {}, _lib.FrameGrads(), ctx["tab"]
        for field, key in self._TAB_KEYS.items():
            if tab.get(key) is not None and field in _lib.FRAME_GRADS and not field.startswith("field2cam"):
                tg[key] = torch.empty(tab[key].shape, device=self.device)
                setattr(fgr, field, tg[key].data_ptr())
        saved = _lib.FieldOutputs()
        if "warp_pts" in ctx["out"]:
            saved.warp_pts = ctx["out"]["warp_pts"].data_ptr()
        rc = self.handle.lib.b200r_warp_bwd(self.handle.h, C.byref(self.desc), _ptr(st["packed_t"]), C.byref(ctx["par"]), C.byref(ctx["fr"]),
                                            C.byref(ctx["pb"]), C.byref(saved), gx.data_ptr(), C.byref(ctx["tape"]), C.byref(pgs), C.byref(fgr),
                                            g_pts.data_ptr(), _ptr(self._ws), self._ws.numel(), _stream(self.device))
        self.handle.check(rc, "b200r_warp_bwd")
        views = {k: flat[slots[k][0]:slots[k][0] + slots[k][1]].view(slots[k][2]) for k in self.warp_weight_names()}
        self._keep_warp = (gx, g_const, g_frame, flat)
        return g_pts.view(M, Pn, 3), views, tg

    # ------------------------------------------------------------------ eikonal term (NeRF.compute_eikonal)
    def eikonal_weight_names(self):
        """Parameters the eikonal term reaches: the basefield's weights and sdf.weight (biases and codes only move the masks)."""
        c = self.cfg
        return [f"basefield.linear_{i+1}.0.weight" for i in range(c.D)] + ["basefield.linear_final.0.weight", "sdf.weight"]

    @torch.no_grad()
    def eikonal_forward(self, ctx, ray_ids):
        """g = d sdf / d xyz (n_rays, D, 3) at all D samples of the rays `ray_ids` (flat indices f * N + n) of the training
        forward that produced `ctx` (nnutils/nerf.py:416-453, utils/torch_utils.py:4-28): the reverse chain of
        b200r_eikonal_fwd with the tape's ReLU signs.  Returns (g, ectx); ectx goes to `eikonal_backward`.  The chain tapes
        belong to the renderer: one eikonal term may be in flight per renderer (run its backward before the next forward)."""
        st = self._train_state()
        ids = ray_ids.to(device=self.device, dtype=torch.int32).contiguous()
        n, D = int(ids.numel()), ctx["D"]
        a, v = C.c_size_t(), C.c_size_t()
        self.handle.check(self.handle.lib.b200r_eikonal_sizes(C.byref(self.desc), n, D, C.byref(a), C.byref(v)), "b200r_eikonal_sizes")
        bufs = getattr(self, "_eik_bufs", None)
        if bufs is None or bufs[0].numel() < a.value + 1024 or bufs[1].numel() < v.value + 1024:
            bufs = (torch.empty(a.value + 1024, dtype=torch.uint8, device=self.device), torch.empty(v.value + 1024, dtype=torch.uint8, device=self.device))
            self._eik_bufs = bufs
        eb = _lib.EikBatch()
        al = lambda b: (b.data_ptr() + 1023) // 1024 * 1024
        eb.n_rays, eb.rays, eb.a, eb.v, eb.a_bytes, eb.v_bytes = n, ids.data_ptr(), al(bufs[0]), al(bufs[1]), a.value, v.value
        g = torch.empty(n, D, 3, device=self.device)
        rc = self.handle.lib.b200r_eikonal_fwd(self.handle.h, C.byref(self.desc), _ptr(st["packed_t"]), C.byref(ctx["par"]), C.byref(ctx["rb"]),
                                               ctx["M"], ctx["out"]["xyz"].data_ptr(), C.byref(ctx["tape"]), C.byref(eb), g.data_ptr(),
                                               _stream(self.device))
        self.handle.check(rc, "b200r_eikonal_fwd")
        return g, dict(eb=eb, ids=ids, n=n)

    @torch.no_grad()
    def eikonal_backward(self, ctx, ectx, g_g, flat=None):
        """dL/dW of the basefield weights and sdf.weight for the cotangent g_g (n_rays, D, 3) of `eikonal_forward`'s g: two
        forward chains + weight-gradient GEMMs (b200r_eikonal_bwd), ACCUMULATED into `flat` (default: a fresh zero buffer with
        the layout of `grad_buffer()`).  Returns name -> view."""
        st = self._train_state()
        if flat is None:
            flat = torch.zeros(st["total"], device=self.device)
        slots = st["slots"]
        views = {k: flat[slots[k][0]:slots[k][0] + slots[k][1]].view(slots[k][2]) for k in self.eikonal_weight_names()}
        pgs = _lib.ParamGrads()
        pgs.flat = flat.data_ptr()
        for i in range(_lib.MAX_LAYERS):
            pgs.weight_off[i], pgs.bias_off[i] = -1, -1
        for i, (name, _) in enumerate(self._layers):
            pgs.weight_off[i] = slots[name + ".weight"][0]
        for fld in _lib.HEAD_GRADS:
            setattr(pgs, fld, -1)
        pgs.sdf_w = slots["sdf.weight"][0]
        gg = _f32c(g_g.reshape(-1, 3))
        alpha = getattr(self, "_alpha", None)
        wn = [nl for nl in self._window_names() if nl[0].startswith("basefield.")] if alpha is not None else []
        before = {nl: views[nl[0]].clone() for nl in wn}
        rc = self.handle.lib.b200r_eikonal_bwd(self.handle.h, C.byref(self.desc), _ptr(self.packed), C.byref(ctx["par"]), C.byref(ctx["rb"]),
                                               ctx["M"], ctx["out"]["xyz"].data_ptr(), C.byref(ctx["tape"]), C.byref(ectx["eb"]), gg.data_ptr(),
                                               C.byref(pgs), _stream(self.device))
        self.handle.check(rc, "b200r_eikonal_bwd")
        if alpha is not None:  # the annealing window is folded into the packed weights: dW = dW_eff * window (as in `backward`)
            self._apply_window(views, before, alpha, names=wn)
        self._keep_eik = (gg, flat)
        return views

    def _window_names(self):
        c = self.cfg
        return [("basefield.linear_1.0.weight", c.L_xyz), (f"basefield.linear_{c.skip + 1}.0.weight", c.L_xyz), ("colorfield.linear_1.0.weight", c.L_xyz + 2)]

    def _apply_window(self, views, before, alpha, names=None):
        import math

        for name, L in (self._window_names() if names is None else names):
            k = torch.arange(L, device=self.device, dtype=torch.float32)
            wdw = 0.5 * (1 + torch.cos(math.pi * torch.clamp(alpha * L - k, 0.0, 1.0) + math.pi))
            cols = slice(3, 3 + 6 * L)
            delta = views[(name, L)[0]][:, cols] - before[(name, L)][:, cols]
            views[name][:, cols] = before[(name, L)][:, cols] + delta * wdw.repeat_interleave(6)[None]

    # ------------------------------------------------------------------ eval-mode importance sampling
    @torch.no_grad()
    def importance_depths(self, P, rays, tab, D):
        """NeRF.importance_sampling (nnutils/nerf.py:686-738): D/2 uniform samples -> density -> compositing weights ->
        D/2 deterministic inverse-CDF samples, merged: ascending depths (M,N,D,1) for `query_field(..., depth=...)`."""
        Dc = int(D) // 2
        feat, deltas = self.query_field(P, rays, tab, Dc, want=("density", "deltas"))
        M, N = feat["density"].shape[:2]
        R = M * N
        nf = _f32c(rays["near_far"])
        z = torch.linspace(0, 1, Dc, device=self.device)[None]
        depth_c = (nf[:, 0:1] * (1 - z) + nf[:, 1:2] * z)[:, None, :].expand(M, N, Dc).reshape(R, Dc).contiguous()  # sample_cam_rays
        dens, dl = _f32c(feat["density"]), _f32c(deltas)
        w = torch.empty(R, Dc, device=self.device)
        a = _lib.CompositeArgs()
        a.R, a.D, a.n_channels = R, Dc, 0
        a.density, a.deltas, a.weights = dens.data_ptr(), dl.data_ptr(), w.data_ptr()
        self.handle.check(self.handle.lib.b200r_composite_fwd(self.handle.h, C.byref(a), _stream(self.device)), "b200r_composite_fwd")
        return importance_merge(self.handle, self.device, depth_c, w).view(M, N, 2 * Dc, 1)

    # ------------------------------------------------------------------ NeRF.forward on points
    @torch.no_grad()
    def query_points(self, P, xyz, tab, dir=None, want=("rgb", "density", "sdf")):
        """NeRF.forward (nnutils/nerf.py:167-215) on canonical points: xyz (M,P,3) [dir (M,P,3) in field space];
        tab holds the per-frame code rows (inst_base, inst_color, appr_code).  Returns a dict with rgb (M,P,3),
        density (M,P,1), sdf (M,P,1) as requested.  Only the basefield / colorfield / heads run."""
        xyz = _f32c(xyz)
        M, Pn = xyz.shape[:2]
        par, keep = self._params(P)
        keep.append(xyz)
        fr = _lib.FrameTables()
        fr.M = M
        for field in ("inst_base", "inst_color", "appr_code"):
            if tab.get(field) is not None:
                t = _f32c(tab[field])
                keep.append(t)
                setattr(fr, field, t.data_ptr())
        pb = _lib.PointBatch()
        pb.P = Pn
        pb.xyz = xyz.data_ptr()
        if dir is not None:
            d = _f32c(dir)
            keep.append(d)
            pb.dir = d.data_ptr()
        out, oa = {}, _lib.FieldOutputs()
        widths = dict(_lib.FIELD_OUTPUTS)
        for name in want:
            out[name] = torch.empty(M * Pn, widths[name], dtype=torch.float32, device=self.device)
            setattr(oa, name, out[name].data_ptr())
        wbytes = self.handle.lib.b200r_workspace_bytes(C.byref(self.desc), M)
        if getattr(self, "_ws", None) is None or self._ws.numel() < wbytes:
            self._ws = torch.empty(wbytes, dtype=torch.uint8, device=self.device)
        rc = self.handle.lib.b200r_points_fwd(self.handle.h, C.byref(self.desc), _ptr(self.packed), C.byref(par), C.byref(fr),
                                              C.byref(pb), C.byref(oa), _ptr(self._ws), self._ws.numel(), _stream(self.device))
        self.handle.check(rc, "b200r_points_fwd")
        self._keep_call = keep
        return {k: v.view(M, Pn, -1) for k, v in out.items()}


    # ------------------------------------------------------------------ SkinningWarp / ComposedWarp on points
    @torch.no_grad()
    def warp_points(self, P, xyz, tab, backward):
        """SkinningWarp.forward / ComposedWarp.forward (nnutils/warping.py:277-336, 445-483) on points xyz (M,P,3):
        backward=True maps time-t space to canonical space, False canonical to time-t space with the frame's own
        articulation.  tab: the skinning (and dense-warp) code rows and articulations of `query_field`.  Returns
        (xyz' (M,P,3), {"skin_entropy", "delta_skin"} (M,P,1))."""
        if self.cfg.motion == "rigid":
            raise RuntimeError("warp_points: the field has no skinning warp")
        xyz = _f32c(xyz)
        M, Pn = xyz.shape[:2]
        par, keep = self._params(P)
        keep.append(xyz)
        fr = _lib.FrameTables()
        fr.M = M
        for field, key in self._TAB_KEYS.items():
            if key in tab and tab[key] is not None and not field.startswith("field2cam"):
                t = _f32c(tab[key])
                keep.append(t)
                setattr(fr, field, t.data_ptr())
        pb = _lib.PointBatch()
        pb.P = Pn
        pb.xyz = xyz.data_ptr()
        out, oa = {}, _lib.FieldOutputs()
        for name, nch in (("xyz", 3), ("skin_entropy", 1), ("delta_skin", 1)):
            out[name] = torch.empty(M, Pn, nch, dtype=torch.float32, device=self.device)
            setattr(oa, name, out[name].data_ptr())
        wbytes = self.handle.lib.b200r_workspace_bytes(C.byref(self.desc), M)
        if getattr(self, "_ws", None) is None or self._ws.numel() < wbytes:
            self._ws = torch.empty(wbytes, dtype=torch.uint8, device=self.device)
        rc = self.handle.lib.b200r_warp_fwd(self.handle.h, C.byref(self.desc), _ptr(self.packed), C.byref(par), C.byref(fr),
                                            C.byref(pb), int(bool(backward)), C.byref(oa), _ptr(self._ws), self._ws.numel(),
                                            _stream(self.device))
        self.handle.check(rc, "b200r_warp_fwd")
        self._keep_call = keep
        return out["xyz"], {"skin_entropy": out["skin_entropy"], "delta_skin": out["delta_skin"]}


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


def _composite_call(h, device, dens, dl, plan_part, srcs, M, N, D, bwd=None):
    """One b200r_composite_fwd (or _bwd) launch over <= MAX_CHANNELS arrays."""
    R = M * N
    a = _lib.CompositeArgs()
    a.R, a.D = R, D
    a.density, a.deltas = dens.data_ptr(), dl.data_ptr()
    a.n_channels = len(plan_part)
    outs = []
    mask = torch.empty(M, N, 1, device=device)
    a.mask = mask.data_ptr()
    for i, ((k, mode, nout), src) in enumerate(zip(plan_part, srcs)):
        a.src[i] = src.data_ptr()
        a.nch[i], a.mode[i] = src.shape[-1], mode
        if bwd is None:
            dst = torch.empty(M, N, nout, device=device)
            outs.append(dst)
            a.dst[i] = dst.data_ptr()
    if bwd is None:
        h.check(h.lib.b200r_composite_fwd(h.h, C.byref(a), _stream(device)), "b200r_composite_fwd")
        return mask, outs
    g_mask, g_outs, need = bwd
    b = _lib.CompositeBwdArgs()
    b.fwd = a
    b.g_mask = g_mask.data_ptr() if g_mask is not None else None
    g_dens = torch.empty_like(dens)
    b.g_density = g_dens.data_ptr()
    g_srcs = []
    for i, src in enumerate(srcs):
        if g_outs[i] is not None:
            b.g_dst[i] = g_outs[i].data_ptr()
        if need[i]:
            g = torch.empty_like(src)
            b.g_src[i] = g.data_ptr()
            g_srcs.append(g)
        else:
            g_srcs.append(None)
    h.check(h.lib.b200r_composite_bwd(h.h, C.byref(b), _stream(device)), "b200r_composite_bwd")
    return g_dens, g_srcs


class _Composite(torch.autograd.Function):
    """render_pixel's per-ray reductions with the hand-derived backward kernel (csrc/composite.cu)."""

    @staticmethod
    def forward(ctx, plan, density, deltas, *values):
        device = density.device
        h = _lib.handle_for(device)
        M, N, D = density.shape[:3]
        dens, dl = _f32c(density.detach()), _f32c(deltas.detach())
        srcs = [_f32c(v.detach()) for v in values]
        outs, mask = [], None
        for c0 in range(0, max(len(plan), 1), _lib.MAX_CHANNELS):
            m, o = _composite_call(h, device, dens, dl, plan[c0:c0 + _lib.MAX_CHANNELS], srcs[c0:c0 + _lib.MAX_CHANNELS], M, N, D)
            mask = m
            outs += o
        ctx.plan, ctx.shape = plan, (M, N, D)
        ctx.save_for_backward(dens, dl, *srcs)
        return (mask, *outs)

    @staticmethod
    def backward(ctx, g_mask, *g_outs):
        dens, dl, *srcs = ctx.saved_tensors
        plan = ctx.plan
        M, N, D = ctx.shape
        device = dens.device
        h = _lib.handle_for(device)
        g_dens_total, g_vals = None, []
        for c0 in range(0, max(len(plan), 1), _lib.MAX_CHANNELS):
            part = plan[c0:c0 + _lib.MAX_CHANNELS]
            go = [None if g is None else _f32c(g) for g in g_outs[c0:c0 + _lib.MAX_CHANNELS]]
            need = [ctx.needs_input_grad[3 + c0 + i] for i in range(len(part))]
            gm = _f32c(g_mask) if (g_mask is not None and c0 == 0) else None
            g_dens, g_srcs = _composite_call(h, device, dens, dl, part, srcs[c0:c0 + _lib.MAX_CHANNELS], M, N, D, bwd=(gm, go, need))
            g_dens_total = g_dens if g_dens_total is None else g_dens_total + g_dens
            g_vals += g_srcs
        return (None, g_dens_total.view(M, N, D, 1), None, *g_vals)


def render_pixel(field_dict, deltas):
    """Volume-render per-sample field outputs along rays (utils/render_utils.py:59-184).
    field_dict: key -> (M,N,D,c); deltas (M,N,D,1).  Returns key -> (M,N,c).  Differentiable w.r.t.
    every per-sample array (hand-derived backward kernel); deltas are treated as constants."""
    dens = field_dict["density"]
    device = dens.device
    M, N, D = dens.shape[:3]
    R = M * N
    plan = _channel_plan(field_dict)
    res = _Composite.apply(plan, dens, deltas, *[field_dict[k] for k, _, _ in plan])
    out = {"mask": res[0]}
    for (k, mode, nout), v in zip(plan, res[1:]):
        out[k] = v
    # per-batch normalisers (tiny (M,N) tensors)
    if "vis" in out:
        v = out["vis"]  # -(v0 / D) / (sum(v1) / (R D)) with the scalar folded first: three small kernels instead of five
        out["vis"] = v[..., :1] * (-float(R) / v[..., 1].detach().sum())
    for k in ("eikonal", "delta_skin"):
        if k in out:
            out[k] = out[k][..., 0]
    if "gauss_density" in out:
        out["gauss_mask"] = out.pop("gauss_density")
    dkeys = [k for k in out if k.startswith("density_")]
    if dkeys:
        dsum = (out[dkeys[0]] if len(dkeys) == 1 else torch.cat([out[k] for k in dkeys], -1).sum(-1, keepdim=True)) + 1e-6
        for k in dkeys:
            out[k.replace("density_", "mask_")] = out.pop(k) / dsum
    if "normal" in out:
        out["normal"] = torch.nn.functional.normalize(out["normal"], 2, -1)
    return out


def _merge_pair(h, device, fa, fb, da, db, want_perm=False):
    """b200r_compose_fwd over the union of keys of two fields (dicts key -> (M,N,D,c)); deltas travel as one more key.
    want_perm: also return the (M*N, Da+Db) int32 index of every merged sample inside the concatenation [A; B]."""
    depth_a, depth_b = _f32c(fa["depth"]), _f32c(fb["depth"])
    M, N, Da = depth_a.shape[:3]
    Db = depth_b.shape[2]
    R = M * N
    fa, fb = dict(fa, __deltas=da), dict(fb, __deltas=db)
    keys = [k for k in fa] + [k for k in fb if k not in fa]
    out, keep = {}, [depth_a, depth_b]
    perm = torch.empty(R, Da + Db, dtype=torch.int32, device=device) if want_perm else None
    for c0 in range(0, len(keys), _lib.MAX_CHANNELS):
        part = keys[c0:c0 + _lib.MAX_CHANNELS]
        a = _lib.ComposeArgs()
        a.R, a.Da, a.Db, a.n_channels = R, Da, Db, len(part)
        a.depth_a, a.depth_b = depth_a.data_ptr(), depth_b.data_ptr()
        if perm is not None and c0 == 0:
            a.perm = perm.data_ptr()
        for i, k in enumerate(part):
            ta = _f32c(fa[k]) if k in fa else None
            tb = _f32c(fb[k]) if k in fb else None
            nch = (ta if ta is not None else tb).shape[-1]
            keep += [ta, tb]
            a.src_a[i] = ta.data_ptr() if ta is not None else None
            a.src_b[i] = tb.data_ptr() if tb is not None else None
            a.nch[i] = nch
            out[k] = torch.empty(M, N, Da + Db, nch, device=device)
            a.dst[i] = out[k].data_ptr()
        h.check(h.lib.b200r_compose_fwd(h.h, C.byref(a), _stream(device)), "b200r_compose_fwd")
    deltas = out.pop("__deltas")
    return (out, deltas, perm) if want_perm else (out, deltas)


class _Compose(torch.autograd.Function):
    """compose_fields of two fields with the hand-derived backward: the merge is a permutation of the concatenated samples
    (multifields.py:393-397), so every key's gradient goes back through it (b200r_compose_bwd: one launch per 16 keys)."""

    @staticmethod
    def forward(ctx, keys_a, keys_b, da, db, *vals):
        fa = dict(zip(keys_a, vals[:len(keys_a)]))
        fb = dict(zip(keys_b, vals[len(keys_a):]))
        device = da.device
        out, deltas, perm = _merge_pair(_lib.handle_for(device), device, {k: v.detach() for k, v in fa.items()},
                                        {k: v.detach() for k, v in fb.items()}, da.detach(), db.detach(), want_perm=True)
        ctx.keys_a, ctx.keys_b, ctx.out_keys = keys_a, keys_b, list(out.keys())
        ctx.Da, ctx.Db = fa["depth"].shape[2], fb["depth"].shape[2]
        ctx.save_for_backward(perm)
        ctx.mark_non_differentiable(deltas)
        return (deltas, *[out[k] for k in ctx.out_keys])

    @staticmethod
    def backward(ctx, g_deltas, *g_outs):
        (perm,) = ctx.saved_tensors
        R, Dt = perm.shape
        device = perm.device
        h = _lib.handle_for(device)
        g = dict(zip(ctx.out_keys, g_outs))
        keys = [k for k in ctx.out_keys if g.get(k) is not None]
        ga, gb, keep = {}, {}, []
        for c0 in range(0, len(keys), _lib.MAX_CHANNELS):  # b200r_compose_bwd: one launch per 16 keys, both fields
            part = keys[c0:c0 + _lib.MAX_CHANNELS]
            b = _lib.ComposeBwdArgs()
            b.R, b.Da, b.Db, b.n_channels, b.perm = R, ctx.Da, ctx.Db, len(part), perm.data_ptr()
            for i, k in enumerate(part):
                go = _f32c(g[k])
                keep.append(go)
                M, N, _, c = go.shape
                b.g_dst[i], b.nch[i] = go.data_ptr(), c
                if k in ctx.keys_a:
                    ga[k] = torch.empty(M, N, ctx.Da, c, device=device)
                    b.g_a[i] = ga[k].data_ptr()
                if k in ctx.keys_b:
                    gb[k] = torch.empty(M, N, ctx.Db, c, device=device)
                    b.g_b[i] = gb[k].data_ptr()
            h.check(h.lib.b200r_compose_bwd(h.h, C.byref(b), _stream(device)), "b200r_compose_bwd")
        return (None, None, None, None, *[ga.get(k) for k in ctx.keys_a], *[gb.get(k) for k in ctx.keys_b])


def compose_fields(feats, deltas_list):
    """MultiFields.compose_fields (nnutils/multifields.py:339-398): merge the fields' samples along every ray by
    depth and carry every key (zeros where a field lacks it).  feats: list of dicts in field order; every field's
    samples are sorted by depth (uniform placement), so the reference's concatenate + argsort + gather is a merge,
    done by the depth-merge kernel (csrc/compose.cu); more than two fields are merged pairwise in field order.
    Differentiable w.r.t. every per-sample array (the backward is the inverse gather of the merge permutation)."""
    if len(feats) == 1:
        return dict(feats[0]), deltas_list[0]
    device = feats[0]["depth"].device
    h = _lib.handle_for(device)
    out, deltas = feats[0], deltas_list[0]
    for f, d in zip(feats[1:], deltas_list[1:]):
        needs_grad = torch.is_grad_enabled() and any(v.requires_grad for v in list(out.values()) + list(f.values()))
        if needs_grad:
            ka, kb = list(out.keys()), list(f.keys())
            res = _Compose.apply(ka, kb, deltas, d, *[out[k] for k in ka], *[f[k] for k in kb])
            deltas = res[0]
            merged_keys = ka + [k for k in kb if k not in ka]
            out = dict(zip(merged_keys, res[1:]))
        else:
            with torch.no_grad():
                out, deltas = _merge_pair(h, device, out, f, deltas, d)
    return out, deltas


# ---------------------------------------------------------------------------------------- per-ray feature matching
class _Match(torch.autograd.Function):
    """b200r_match_fwd / b200r_match_bwd: softmax matching of every ray's pixel feature against K candidate samples."""

    @staticmethod
    def forward(ctx, feat_px, feat_can, xyz_can, logsigma, idx):
        dev = feat_can.device
        h = _lib.handle_for(dev)
        fp, fc, xc, ls = _f32c(feat_px.detach()), _f32c(feat_can.detach()), _f32c(xyz_can.detach()), _f32c(logsigma.detach())
        R, K = fp.shape[0], int(idx.numel())
        out, lse = torch.empty(R, 3, device=dev), torch.empty(R, device=dev)
        a = _lib.MatchArgs()
        a.R, a.K = R, K
        a.feat_px, a.feat_can, a.xyz_can, a.idx, a.logsigma = fp.data_ptr(), fc.data_ptr(), xc.data_ptr(), idx.data_ptr(), ls.data_ptr()
        a.xyz_matched, a.lse = out.data_ptr(), lse.data_ptr()
        h.check(h.lib.b200r_match_fwd(h.h, C.byref(a), _stream(dev)), "b200r_match_fwd")
        ctx.save_for_backward(fp, fc, xc, ls, idx, out, lse)
        ctx.args = a
        return out

    @staticmethod
    def backward(ctx, g):
        fp, fc, xc, ls, idx, out, lse = ctx.saved_tensors
        dev = fc.device
        h = _lib.handle_for(dev)
        b = _lib.MatchBwdArgs()
        b.fwd = ctx.args
        gg = _f32c(g)
        g_fc, g_xc, g_ls = torch.zeros_like(fc), torch.zeros_like(xc), torch.zeros_like(ls)
        scratch = torch.empty(h.lib.b200r_match_scratch_floats(b.fwd.R, b.fwd.K), device=dev)
        b.g_out, b.g_feat_can, b.g_xyz_can, b.g_logsigma, b.scratch = gg.data_ptr(), g_fc.data_ptr(), g_xc.data_ptr(), g_ls.data_ptr(), scratch.data_ptr()
        h.check(h.lib.b200r_match_bwd(h.h, C.byref(b), _stream(dev)), "b200r_match_bwd")
        return None, g_fc, g_xc, g_ls, None


def global_match(feat_px, feat_canonical, xyz_canonical, logsigma, num_candidates=1024, rng="reference"):
    """FeatureNeRF.global_match (nnutils/feature.py:152-205): feat_px (M,N,16) pixel features, feat_canonical (M,N,D,16) and
    xyz_canonical (M,N,D,3) of the batch's samples, logsigma (1) -> matched canonical points (M,N,3).  Differentiable w.r.t.
    feat_canonical, xyz_canonical and logsigma.  rng="reference": the candidates are drawn like the reference -
    torch.randperm(S) on the default CPU generator (same random stream; ~2 ms of host time at S = 262 144 plus a blocking
    copy, measured on the B200 box); rng="device": the same draw on the GPU's generator (no host work, another stream)."""
    if feat_px.shape[-1] != 16 or feat_canonical.shape[-1] != 16:
        raise NotImplementedError("global_match: built for 16 feature channels")
    shape = feat_px.shape
    fc, xc = feat_canonical.reshape(-1, 16), xyz_canonical.reshape(-1, 3)
    K = min(int(num_candidates), fc.shape[0], 2048)
    idx = torch.randperm(fc.shape[0], device=fc.device)[:K] if rng == "device" else torch.randperm(fc.shape[0])[:K].to(fc.device)
    out = _Match.apply(feat_px.reshape(-1, 16), fc, xc, logsigma.reshape(1), idx)
    return out.view(shape[:-1] + (3,))


# ---------------------------------------------------------------------------------------- per-pixel reconstruction losses
_FIELD_TYPES = {"fg": 0, "bg": 1, "comp": 2}
# (struct field of the rendered input, its gradient field, channels)
_LOSS_DIFF = [("r_mask", "g_mask", 1), ("r_mask_fg", "g_mask_fg", 1), ("r_rgb", "g_rgb", 3), ("r_depth", "g_depth", 1), ("r_flow", "g_flow", 2),
              ("vis_fg", "g_vis_fg", 1), ("vis_bg", "g_vis_bg", 1), ("a_feature", "g_feature", 16), ("a_xy_reproj", "g_xy_reproj", 2),
              ("a_gauss_mask", "g_gauss_mask", 1)]


class _Losses(torch.autograd.Function):
    """b200r_loss_fwd / b200r_loss_bwd over the rendered tensors `diff` (name -> tensor, differentiable) and the batch
    tensors `data` (name -> tensor): returns the 8 weighted loss terms as one vector."""

    @staticmethod
    def forward(ctx, meta, *diff_vals):
        names, data = meta["names"], meta["data"]
        dev = diff_vals[0].device
        h = _lib.handle_for(dev)
        a = _lib.LossArgs()
        a.M, a.N, a.field_type, a.train_res = meta["M"], meta["N"], meta["field_type"], float(meta["train_res"])
        keep = []
        for n, t in list(zip(names, diff_vals)) + list(data.items()):
            t = _f32c(t.detach())
            keep.append(t)
            setattr(a, n, t.data_ptr())
        for k in range(8):
            a.wt[k] = float(meta["wt"][k])
        loss, stats = torch.empty(8, device=dev), torch.empty(24, device=dev)
        a.loss, a.stats = loss.data_ptr(), stats.data_ptr()
        h.check(h.lib.b200r_loss_fwd(h.h, C.byref(a), _stream(dev)), "b200r_loss_fwd")
        ctx.a, ctx.keep, ctx.names, ctx.dev = a, keep + [loss, stats], names, dev
        ctx.shapes = [t.shape for t in diff_vals]
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        h = _lib.handle_for(ctx.dev)
        b = _lib.LossBwdArgs()
        b.fwd = ctx.a
        gl = _f32c(torch.nan_to_num(g_loss, nan=0.0))
        b.g_loss = gl.data_ptr()
        R = ctx.a.M * ctx.a.N
        gfield = {f: (g, c) for f, g, c in _LOSS_DIFF}
        outs = []
        for n, shp in zip(ctx.names, ctx.shapes):
            gname, c = gfield[n]
            t = torch.empty(R, c, device=ctx.dev)
            setattr(b, gname, t.data_ptr())
            outs.append(t.view(shp))
        h.check(h.lib.b200r_loss_bwd(h.h, C.byref(b), _stream(ctx.dev)), "b200r_loss_bwd")
        return (None, *outs)


def recon_losses(rendered, aux_dict, batch, config):
    """The reconstruction part of dvr_model.compute_loss (lab4d/engine/model.py:374-611: get_mask_balance_wt,
    compute_recon_loss, mask_losses and apply_loss_weights for the keys compute_recon_loss creates) as two kernels.
    rendered / aux_dict[cate]: (M,N,c) tensors as render_samples returns them; batch: the reshaped dataloader batch
    (mask, vis2d, is_detected, rgb, depth, flow, flow_uct, feature, hxy); config: field_type, train_res, `<key>_wt`.
    Returns the loss_dict entries (0-d tensors, the reference's key order)."""
    ft = config["field_type"]
    M, N = rendered["rgb"].shape[:2]
    diff = {"r_mask": rendered["mask"], "r_rgb": rendered["rgb"], "r_depth": rendered["depth"], "r_flow": rendered["flow"]}
    if ft == "comp":
        diff["r_mask_fg"] = rendered["mask_fg"]
    for cate in aux_dict:
        diff["vis_" + cate] = aux_dict[cate]["vis"]
    keys = ["mask", "rgb", "depth", "flow", "vis"]
    if ft in ("fg", "comp"):
        diff["a_feature"], diff["a_xy_reproj"] = aux_dict["fg"]["feature"], aux_dict["fg"]["xy_reproj"]
        keys = ["mask", "feature", "feat_reproj", "rgb", "depth", "flow", "vis"]
        if "gauss_mask" in rendered:
            diff["a_gauss_mask"] = aux_dict["fg"]["gauss_mask"]
            keys.append("reg_gauss_mask")
    data = {"b_mask": batch["mask"].float(), "b_vis2d": batch["vis2d"].float(), "b_is_detected": batch["is_detected"].float(),
            "b_rgb": batch["rgb"], "b_depth": batch["depth"], "b_flow": batch["flow"], "b_flow_uct": batch["flow_uct"]}
    if ft in ("fg", "comp"):
        data["b_feature"], data["b_hxy"] = batch["feature"], batch["hxy"]
    names = list(diff)
    meta = dict(names=names, data=data, M=M, N=N, field_type=_FIELD_TYPES[ft], train_res=config["train_res"],
                wt=[config.get(k + "_wt", 1.0) for k in _lib.LOSS_TERMS])
    vec = _Losses.apply(meta, *[diff[n] for n in names])
    return {k: vec[_lib.LOSS_TERMS.index(k)] for k in keys}
