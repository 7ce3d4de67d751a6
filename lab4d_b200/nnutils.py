"""Drop-in adapter for the reference's module surface.

`install()` rebinds, inside an imported `lab4d` package,
    lab4d.nnutils.nerf.NeRF.query_field            (also reached by FeatureNeRF / Deformable via super())
    lab4d.nnutils.deformable.Deformable.query_field
    lab4d.utils.render_utils.render_pixel and the by-name import lab4d.engine.model.render_pixel
    lab4d.nnutils.multifields.MultiFields.compose_fields, lab4d.engine.model.dvr_model.compute_loss (reconstruction terms)
so that lab4d.engine.model.dvr_model.render_samples (engine/model.py:328-361) and lab4d/render.py call the
B200 renderer unchanged.  Parameters stay nn.Parameters of the reference modules (checkpoints, optimiser
param groups and DDP are untouched); per-frame codes / cameras / articulations are still produced by the
reference's small per-frame MLPs and handed over as tables.

Training (autograd recording): query_field goes through lab4d_b200.autograd.FieldFunction - training forward with the
tape, hand-derived backward kernels, gradients delivered to the reference modules' own Parameters and, through the
per-frame tables, to the camera / articulation / embedding modules.  Eval mode: importance sampling + bounding-box
masking + the reference's normals.  The eikonal term of a training step runs on the eikonal kernels (compute_eikonal below);
global_match (feature.py:152-205) runs on the match kernels; forward_project (feature.py:207-226: one warp of N points per frame +
the pinhole projection) stays the reference's own torch code on the matched points.
"""
import functools

import torch

from . import render as _render
from .spec import FieldConfig


def config_from_module(field) -> FieldConfig:
    """Read the architecture of a reference NeRF / Deformable module."""
    base = field.basefield
    W, D = base.W, base.D
    motion, B, symm, dense = "rigid", 0, None, False
    warp = getattr(field, "warp", None)
    if warp is not None and hasattr(warp, "skinning_model"):
        if hasattr(warp, "post_warp"):  # ComposedWarp (nnutils/warping.py:417-443)
            pw = warp.post_warp
            if type(pw).__name__ != "DenseWarp" or pw.forward_map.D != 2 or pw.forward_map.W != 256 or pw.pos_embedding.N_freqs != 6:
                raise NotImplementedError("ComposedWarp: only DenseWarp(D=2, W=256, 6 frequencies) post-warps are accelerated")
            dense = True
        B = warp.skinning_model.num_coords
        motion = "bob" if type(warp.articulation).__name__ == "ArticulationFlatMLP" else "skel"
        if warp.skinning_model.symm_idx is not None:
            symm = tuple(int(i) for i in warp.skinning_model.symm_idx)
    elif warp is not None and type(warp).__name__ != "IdentityWarp":
        raise NotImplementedError(f"warp type {type(warp).__name__} is not accelerated yet")
    return FieldConfig(category=field.category, D=D, W=W, L_xyz=field.pos_embedding.N_freqs,
                       L_dir=field.dir_embedding.N_freqs, appr_channels=field.appr_channels, skip=base.skips[0],
                       motion=motion, B=B, has_feature=hasattr(field, "feature_field"), symm_idx=symm, dense=dense)


def tables_from_module(field, samples_dict):
    """Per-frame tables of the hot path, computed with the reference's own per-frame modules
    (nnutils/{appearance,embedding,pose}.py - M rows, out of scope as kernels)."""
    frame_id, inst_id = samples_dict["frame_id"], samples_dict["inst_id"]
    tab = {"field2cam_q": samples_dict["field2cam"][0], "field2cam_t": samples_dict["field2cam"][1],
           "inst_base": field.basefield.inst_embedding(inst_id), "inst_color": field.colorfield.inst_embedding(inst_id),
           "inst_vis": field.vis_mlp.basefield.inst_embedding(inst_id)}
    if field.appr_channels > 0:
        tab["appr_code"] = field.appr_embedding.get_vals(frame_id)
    warp = getattr(field, "warp", None)
    if warp is not None and hasattr(warp, "skinning_model"):
        sk = warp.skinning_model
        tab["inst_skin"] = sk.delta_field.inst_embedding(inst_id)
        tab["skin_t_embed"] = sk.time_embedding(frame_id)
        tab["skin_t_embed_mean"] = sk.time_embedding.get_mean_embedding(frame_id.device)
        if hasattr(warp, "post_warp"):
            pw = warp.post_warp
            tab["dense_t_embed"] = pw.time_embedding(frame_id)
            tab["inst_dense_fwd"] = pw.forward_map.inst_embedding(inst_id)
            tab["inst_dense_bwd"] = pw.backward_map.inst_embedding(inst_id)
        if "t_articulation" in samples_dict:
            t_art, r_art = samples_dict["t_articulation"], samples_dict["rest_articulation"]
        else:
            t_art, r_art = warp.articulation.get_vals_and_mean(frame_id)
        tab["t_articulation_qr"], tab["t_articulation_qd"] = t_art
        tab["rest_articulation_qr"], tab["rest_articulation_qd"] = r_art
    return tab


def _renderer_for(field, device, operand_dtype):
    cfg = config_from_module(field)
    cache = field.__dict__.setdefault("_b200_renderer", {})
    key = (cfg, str(device), operand_dtype)
    if key not in cache:
        cache[key] = _render.FieldRenderer(cfg, device, operand_dtype=operand_dtype)
    return cfg, cache[key]


def _hot_params(field, cfg):
    """name -> nn.Parameter of everything the per-sample kernels read (state_dict names, lab4d_b200/spec.py)."""
    from .spec import field_param_shapes

    named = dict(field.named_parameters())
    return {k: named[k] for k in field_param_shapes(cfg)}


def compute_eikonal(field, renderer, ctx, P, xyz, sample_ratio=16, bind_grads=False):
    """Replacement body of NeRF.compute_eikonal (nnutils/nerf.py:416-453) for the samples of a training forward: the same
    random subset of rays (one torch.multinomial draw from the default CPU generator, like the reference), the sdf gradient
    from the eikonal kernels, (|g| - 1)^2 scattered into zeros (M,N,D,1).  Gradients reach the basefield weights and
    sdf.weight through lab4d_b200.autograd.EikonalFunction."""
    from . import autograd as _ag

    M, N, D, _ = xyz.shape
    R = M * N
    sample_size = max(R // sample_ratio, 1)
    rand_inds = torch.multinomial(torch.ones(R), sample_size, replacement=False) if R > sample_size else torch.arange(R)
    g = _ag.eikonal(renderer, ctx, P, rand_inds, bind_grads=bind_grads)
    eik = torch.zeros(R, D, device=xyz.device, dtype=xyz.dtype)
    eik[rand_inds.to(xyz.device)] = (g.norm(2, dim=-1) - 1) ** 2
    return eik.view(M, N, D, 1)


def forward_project(field, renderer, P, tab, xyz, samples_dict, bind_grads=False):
    """Replacement body of FeatureNeRF.forward_project (nnutils/feature.py:207-226) for skinned fields in a training step:
    Deformable.forward_warp (deformable.py:154-171) = the forward skinning warp of the matched points - on the kernels, with its
    hand-derived backward (lab4d_b200.autograd.warp_points) - then the reference's own field_to_cam and pinhole projection
    (a handful of (M,N,3) ops).  xyz (M,N,3) -> (xy (M,N,2), xyz_cam (M,N,3))."""
    from lab4d.utils.geom_utils import Kmatinv, pinhole_projection

    from . import autograd as _ag

    xyz_next = _ag.warp_points(renderer, P, xyz, tab, bind_grads=bind_grads)
    xyz_cam = field.field_to_cam(xyz_next[:, :, None], samples_dict["field2cam"])[:, :, 0]
    xy = pinhole_projection(Kmatinv(samples_dict["Kinv"]), xyz_cam)[..., :2]
    return xy, xyz_cam


def query_field(field, samples_dict, flow_thresh=None, n_depth=64, operand_dtype="fp16x3", bind_grads=False, match_rng="reference"):
    """Replacement body of NeRF.query_field (nnutils/nerf.py:580-684) for NeRF / FeatureNeRF / Deformable modules.
    Training mode: the fused kernels (with the tape and the hand-derived backward when autograd is recording); the
    eikonal term runs on the eikonal kernels (`compute_eikonal` above: reverse chain with the tape's ReLU signs, hand-derived
    second-order backward) - on the reference's own `compute_eikonal` only when no gradient is recorded.
    Eval mode (`lab4d/render.py` -> dvr_model.evaluate): importance sampling (nerf.py:686-738), samples outside the
    bounding boxes zeroed like the reference's masked query_nerf (nerf.py:495-528, 769-819), normals and the eval eikonal
    from the normals kernel (`FieldRenderer.sdf_gradient_cam`, nerf.py:455-493).  Returns (feat_dict, deltas, aux_dict) like
    the reference."""
    if field.pos_embedding.alpha is not None and field.pos_embedding_color.alpha != field.pos_embedding.alpha:
        raise NotImplementedError("different annealing windows for density and colour embeddings")
    dev = samples_dict["hxy"].device
    cfg, r = _renderer_for(field, dev, operand_dtype)
    P = _hot_params(field, cfg)
    alpha = field.pos_embedding.alpha
    rays = {"hxy": samples_dict["hxy"], "Kinv": samples_dict["Kinv"], "near_far": samples_dict["near_far"]}
    inst_id = samples_dict["inst_id"]
    if field.training:
        needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in P.values()) or samples_dict["Kinv"].requires_grad)
        tab = tables_from_module(field, samples_dict)
        if needs_grad:
            from . import autograd as _ag

            r.pack_train({k: v.detach() for k, v in P.items()}, alpha=alpha)
            feat, deltas, ctx = _ag.query_field(r, P, rays, tab, n_depth, flow_thresh=flow_thresh, bind_grads=bind_grads, return_ctx=True)
            feat["eikonal"] = compute_eikonal(field, r, ctx, P, feat["xyz"], bind_grads=bind_grads)  # eikonal kernels on the tape's masks
        else:
            with torch.no_grad():
                r.pack(P, alpha=alpha)
                feat, deltas = r.query_field(P, rays, tab, n_depth, flow_thresh=flow_thresh)
            feat["eikonal"] = field.compute_eikonal(feat["xyz"], inst_id=inst_id)  # no tape: the reference's own (nerf.py:416-453)
    else:
        with torch.no_grad():
            tab = tables_from_module(field, samples_dict)
            Pd = {k: v.detach() for k, v in P.items()}
            r.pack_train(Pd, alpha=alpha)
            depth = r.importance_depths(P, rays, tab, n_depth)
            # the training-form forward (with a tape) on the importance-sampled depths: the normals kernel reads its ReLU signs
            feat, deltas, nctx = r.query_field_train(Pd, rays, tab, n_depth, depth=depth)
            for k in ("cyc_dist", "delta_skin", "skin_entropy", "flow", "feature", "eikonal"):  # train-only outputs (nerf.py:590-684)
                feat.pop(k, None)
            xyz_t = nctx["out"]["xyz_t"].view(feat["xyz"].shape)
            valid = field.get_valid_idx(feat["xyz"], xyz_t, feat["vis"], samples_dict)
            if valid is not None:  # the reference evaluates only these samples and leaves zeros elsewhere
                m = valid[..., None].to(feat["rgb"].dtype)
                feat["rgb"], feat["density"] = feat["rgb"] * m, feat["density"] * m
                feat["density_" + cfg.category] = feat["density"]
            # normals (nerf.py:455-493): the gradient of the sdf through the basefield and the backward warp w.r.t. the camera-space
            # points - the normals kernel instead of the reference's autograd.grad over the whole batch
            g = r.sdf_gradient_cam(nctx)
            feat["eikonal"] = (g.norm(2, dim=-1, keepdim=True) - 1) ** 2
            feat["normal"] = torch.nn.functional.normalize(g, dim=-1) * torch.tensor([1.0, -1.0, -1.0], device=g.device)
    aux = {}
    if hasattr(field, "global_match") and "feature" in samples_dict and "feature" in feat:  # FeatureNeRF.query_field, feature.py:119-131
        xyz_matches = _render.global_match(samples_dict["feature"], feat["feature"], feat["xyz"], field.logsigma, rng=match_rng)  # match kernels
        if field.training and cfg.motion != "rigid" and torch.is_grad_enabled() and xyz_matches.requires_grad:
            # forward_project (feature.py:207-226) with the warp of the matched points on the kernels (autograd.WarpFunction)
            xy_reproj, xyz_reproj = forward_project(field, r, P, tab, xyz_matches, samples_dict, bind_grads)
        else:
            xy_reproj, xyz_reproj = field.forward_project(xyz_matches, samples_dict["field2cam"], samples_dict["Kinv"], samples_dict["frame_id"],
                                                           inst_id, samples_dict=samples_dict)
        aux.update(xyz_matches=xyz_matches, xyz_reproj=xyz_reproj, xy_reproj=xy_reproj)
    return feat, deltas, aux


def nerf_forward(field, xyz, dir=None, frame_id=None, inst_id=None, get_density=True):
    """Replacement body of NeRF.forward (nnutils/nerf.py:167-215) for inference callers (mesh extraction, geometry
    queries, eval-mode query_nerf): xyz (M,N,D,3) with per-frame ids (M,), or flat (P,3) with ids None (mean instance
    code) -> (rgb, density-or-sdf) when dir is given, else density-or-sdf, shaped like xyz[..., :1]."""
    if torch.is_grad_enabled() and any(p.requires_grad for p in field.parameters()):
        raise NotImplementedError("lab4d_b200: NeRF.forward is accelerated for inference only (call under torch.no_grad())")
    cfg = config_from_module(field)
    cache = field.__dict__.setdefault("_b200_renderer", {})
    key = (cfg, str(xyz.device))
    if key not in cache:
        cache[key] = _render.FieldRenderer(cfg, xyz.device)
    r = cache[key]
    P = {k: v for k, v in field.named_parameters()}
    r.pack(P, alpha=field.pos_embedding.alpha)
    flat = xyz.dim() == 2
    pts = xyz.reshape(1, -1, 3) if flat else xyz.reshape(xyz.shape[0], -1, 3)
    M = pts.shape[0]
    def code(mlp):  # CondMLP.forward: mean instance code when inst_id is None (nnutils/base.py:131-135)
        c = mlp.inst_embedding.get_mean_embedding() if inst_id is None else mlp.inst_embedding(inst_id)
        return c.reshape(-1, 32).expand(M, -1)

    tab = {"inst_base": code(field.basefield), "inst_color": code(field.colorfield)}
    if field.appr_channels > 0 and dir is not None:
        tab["appr_code"] = field.appr_embedding.get_vals(frame_id).reshape(-1, field.appr_channels).expand(M, -1)
    want = ("density" if get_density else "sdf",) + (("rgb",) if dir is not None else ())
    d = None if dir is None or cfg.L_dir != 0 else dir.reshape(M, -1, 3)
    out = r.query_points(P, pts, tab, dir=d, want=want)
    shape = xyz.shape[:-1]
    val = out[want[0]].reshape(shape + (1,))
    return (out["rgb"].reshape(shape + (3,)), val) if dir is not None else val


def compose_fields(multifields_dict, deltas_dict):
    """Replacement body of MultiFields.compose_fields (nnutils/multifields.py:339-398): the depth-merge kernel instead of
    cat + argsort + 15 gathers, differentiable (inverse gather of the merge permutation).  Same arguments (dicts keyed by
    field category, in field order) and return value."""
    cats = list(multifields_dict.keys())
    return _render.compose_fields([multifields_dict[c] for c in cats], [deltas_dict[c] for c in cats])


def compute_loss(model, batch, results):
    """Replacement body of dvr_model.compute_loss (lab4d/engine/model.py:374-398): the per-pixel reconstruction terms
    (compute_recon_loss + mask_losses + apply_loss_weights, ~40 small launches and their backward) on the loss kernels; the
    regularisers (compute_reg_loss: a few scalars and four per-pixel maps) keep the reference's code."""
    config = model.config
    loss_dict = _render.recon_losses(results["rendered"], results["aux_dict"], batch, config)
    reg = {}
    model.compute_reg_loss(reg, results)
    type(model).apply_loss_weights(reg, config)
    loss_dict.update(reg)
    return loss_dict


def _dq_mul(a, b):
    """lab4d.utils.quat_transform.quaternion_mul (quat_transform.py:106-113) on the quaternion kernels; operands that the
    reference would have to broadcast by hand are broadcast here (3-vectors stay 3-wide: pure quaternions).  CUDA tensors only:
    `install(dqtorch=True)` leaves the reference's own function in charge of CPU tensors (model set-up, data loading)."""
    from . import quaternion as _q

    lead = torch.broadcast_shapes(a.shape[:-1], b.shape[:-1])
    a2 = a.expand(lead + a.shape[-1:]).reshape(-1, a.shape[-1])
    b2 = b.expand(lead + b.shape[-1:]).reshape(-1, b.shape[-1])
    return _q.quaternion_mul(a2, b2).view(lead + (4,))


def _dq_conj(q):
    from . import quaternion as _q

    return _q.quaternion_conjugate(q.reshape(-1, 4)).view(q.shape)


def install(lab4d=None, n_depth=64, operand_dtype="fp16x3", bind_grads=False, match_rng="reference", dqtorch=False):
    """Patch an imported reference package in place; returns a function that undoes the patch.
    operand_dtype: "fp16x3" (parity mode, default), "fp16" or "bf16" (fast modes).
    bind_grads: make the hot-path parameters' .grad views of the renderer's flat gradient buffer (no per-tensor gradient
    copies; all-reduce the buffer yourself) - leave False under DistributedDataParallel, whose reducer waits for autograd's
    per-parameter hooks (engine/trainer.py:110-115).
    dqtorch: also rebind lab4d.utils.quat_transform.quaternion_mul / quaternion_conjugate (the dqtorch extension,
    third_party/quaternion) to the quaternion kernels (lab4d_b200/quaternion.py) in every loaded lab4d module.
    match_rng: "reference" draws global_match's candidates with the reference's own torch.randperm call on the CPU generator
    (same random stream, ~2 ms of host time per step); "device" draws them on the GPU."""
    if lab4d is None:
        import lab4d  # noqa: F401
    import lab4d.engine.model as rmodel
    import lab4d.nnutils.deformable as rdef
    import lab4d.nnutils.feature as rfeat
    import lab4d.nnutils.multifields as rmf
    import lab4d.nnutils.nerf as rnerf
    import lab4d.utils.render_utils as rru

    saved = [(rnerf.NeRF, "query_field", rnerf.NeRF.query_field), (rfeat.FeatureNeRF, "query_field", rfeat.FeatureNeRF.query_field),
             (rdef.Deformable, "query_field", rdef.Deformable.query_field), (rru, "render_pixel", rru.render_pixel),
             (rmodel, "render_pixel", rmodel.render_pixel), (rmf.MultiFields, "compose_fields", rmf.MultiFields.__dict__["compose_fields"]),
             (rmodel.dvr_model, "compute_loss", rmodel.dvr_model.compute_loss)]

    def _qf(self, samples_dict, flow_thresh=None):
        return query_field(self, samples_dict, flow_thresh=flow_thresh, n_depth=n_depth, operand_dtype=operand_dtype, bind_grads=bind_grads,
                           match_rng=match_rng)

    # one body for the three classes: the kernels already produce what FeatureNeRF / Deformable add on top of NeRF
    # (feature field, Gaussian bone density); the per-ray matching of FeatureNeRF runs inside query_field above
    for cls in (rnerf.NeRF, rfeat.FeatureNeRF, rdef.Deformable):
        cls.query_field = _qf
    rru.render_pixel = _render.render_pixel
    rmodel.render_pixel = _render.render_pixel
    rmf.MultiFields.compose_fields = staticmethod(compose_fields)
    rmodel.dvr_model.compute_loss = compute_loss
    if dqtorch:
        # the dqtorch operators behind lab4d.utils.quat_transform (third_party/quaternion): every loaded lab4d module that holds
        # the reference's quaternion_mul / quaternion_conjugate by name gets the kernels' versions
        import sys

        import lab4d.utils.quat_transform as qt

        old = {"quaternion_mul": qt.quaternion_mul, "quaternion_conjugate": qt.quaternion_conjugate}
        old_mul, old_conj = old["quaternion_mul"], old["quaternion_conjugate"]
        # the kernels take CUDA tensors; CPU tensors (the reference's set-up and data-loading code) keep the reference's own functions
        new = {"quaternion_mul": lambda a, b: _dq_mul(a, b) if a.is_cuda else old_mul(a, b),
               "quaternion_conjugate": lambda q: _dq_conj(q) if q.is_cuda else old_conj(q)}
        for mname, mod in list(sys.modules.items()):
            if mod is None or not mname.startswith("lab4d"):
                continue
            for fname, fold in old.items():
                if getattr(mod, fname, None) is fold:
                    saved.append((mod, fname, fold))
                    setattr(mod, fname, new[fname])

    def undo():
        for obj, name, val in saved:
            setattr(obj, name, val)

    return undo
