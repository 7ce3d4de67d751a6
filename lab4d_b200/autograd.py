"""torch.autograd glue of the training path: one Function whose forward is b200r_field_fwd_train and whose backward is
b200r_field_bwd + the per-frame chain (lab4d_b200/prologue_grad.py).  Differentiable inputs: every hot-path parameter
(state_dict names of lab4d_b200/spec.py), every per-frame table of `tab` and rays["Kinv"]; hxy / near_far are data."""
import torch

from . import _lib

_NO_GRAD_OUT = ("deltas", "eikonal", "xyz_t", "dir")


def _bind_views(renderer, params):
    """DDP's gradient_as_bucket_view arrangement: make the leaf parameters' .grad the views of the renderer's flat gradient
    buffer (zeroed when no gradient has been written in this backward yet); the kernels then accumulate straight into it."""
    flat, views = renderer.grad_buffer()
    if all(p.grad is None for p in params.values()):
        flat.zero_()
    for n, p in params.items():
        if p.grad is None:
            p.grad = views[n]
        elif p.grad.data_ptr() != views[n].data_ptr():  # another autograd path got there first: fold it in
            views[n].copy_(p.grad)
            p.grad = views[n]
    return flat


class FieldFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, renderer, meta, *tensors):
        p_names, t_names = meta["p_names"], meta["t_names"]
        P = dict(zip(p_names, tensors[:len(p_names)]))
        tab = dict(zip(t_names, tensors[len(p_names):len(p_names) + len(t_names)]))
        rays = dict(meta["rays"], Kinv=tensors[-1])
        feat, deltas, c = renderer.query_field_train(P, rays, tab, meta["D"], flow_thresh=meta.get("flow_thresh"), depth=meta.get("depth"))
        keys = [k for k in feat if not k.startswith("density_")]
        meta["out_keys"] = keys
        meta["ctx"] = c  # the eikonal term of the same batch reads the tape through it (`eikonal` below)
        ctx.renderer, ctx.c, ctx.keys, ctx.meta = renderer, c, keys, meta
        outs = [feat[k] for k in keys]
        ctx.mark_non_differentiable(deltas, *[o for k, o in zip(keys, outs) if k in _NO_GRAD_OUT])
        return (deltas, *outs)

    @staticmethod
    def backward(ctx, g_deltas, *g_outs):
        grads = {k: g for k, g in zip(ctx.keys, g_outs) if g is not None and k in _lib.GRAD_KEYS}
        meta = ctx.meta
        params = meta.get("bind")  # name -> leaf Parameter whose .grad should BE the flat buffer's view (no per-tensor copies)
        if params is not None:
            _bind_views(ctx.renderer, params)
            _, tg = ctx.renderer.backward(ctx.c, grads, accumulate=True)
            pg = {}
        else:
            pg, tg = ctx.renderer.backward(ctx.c, grads)
            pg = {k: v.clone() for k, v in pg.items()}  # the flat buffer is reused by the next backward
        out = [None, None]
        for n in meta["p_names"]:
            out.append(pg.get(n))
        for n in meta["t_names"]:
            g = tg.get(n)
            out.append(g if g is None else g.reshape(ctx.c["tab"][n].shape))
        out.append(tg.get("Kinv"))
        return tuple(out)


class WarpFunction(torch.autograd.Function):
    """The forward skinning warp (+ soft deformation) of given points, b200r_warp_fwd_train / b200r_warp_bwd: differentiable
    w.r.t. the points, the warp's parameters and the per-frame tables it reads (articulations, skinning / dense codes)."""

    @staticmethod
    def forward(ctx, renderer, meta, xyz, *tensors):
        p_names, t_names = meta["p_names"], meta["t_names"]
        tab = dict(zip(t_names, tensors[len(p_names):]))
        out, c = renderer.warp_points_train(meta["P_all"], xyz, tab)  # the kernels read the field's whole parameter table by pointer
        ctx.renderer, ctx.c, ctx.meta = renderer, c, meta
        return out

    @staticmethod
    def backward(ctx, g_out):
        meta, r = ctx.meta, ctx.renderer
        bind = meta.get("bind")
        if bind is not None:
            flat = _bind_views(r, bind)
            g_pts, _, tg = r.warp_backward(ctx.c, g_out, flat=flat)
            pg = {}
        else:
            g_pts, pg, tg = r.warp_backward(ctx.c, g_out)
        out = [None, None, g_pts]
        for n in meta["p_names"]:
            out.append(pg.get(n))
        for n in meta["t_names"]:
            g = tg.get(n)
            out.append(g if g is None else g.reshape(ctx.c["tab"][n].shape))
        return tuple(out)


def warp_points(renderer, P, xyz, tab, bind_grads=False):
    """Differentiable forward warp (canonical -> time-t space) of points xyz (M,P,3) with the frames' own articulations:
    autograd edges to xyz, to the warp's parameters in P and to tab's tensors.  Call renderer.pack_train first."""
    p_names = [k for k in renderer.warp_weight_names() if k in P]
    warp_tabs = ("inst_skin", "skin_t_embed", "skin_t_embed_mean", "dense_t_embed", "inst_dense_fwd", "inst_dense_bwd", "t_articulation_qr",
                 "t_articulation_qd", "rest_articulation_qr", "rest_articulation_qd")
    t_names = [k for k in warp_tabs if torch.is_tensor(tab.get(k))]
    meta = dict(p_names=p_names, t_names=t_names)
    if bind_grads:
        meta["bind"] = {k: v for k, v in P.items() if v.requires_grad and v.is_leaf}
    meta["P_all"] = {k: v.detach() for k, v in P.items()}
    return WarpFunction.apply(renderer, meta, xyz, *[P[k] for k in p_names], *[tab[k] for k in t_names])


class EikonalFunction(torch.autograd.Function):
    """g = d sdf / d xyz on the samples of a subset of rays (NeRF.compute_eikonal, nnutils/nerf.py:416-453) with a
    hand-derived backward to the basefield weights and sdf.weight - replaces the reference's autograd.grad(create_graph=True)
    and the second-order pass through it (utils/torch_utils.py:4-28)."""

    @staticmethod
    def forward(ctx, renderer, c, ray_ids, bind, names, *weights):
        g, ectx = renderer.eikonal_forward(c, ray_ids)
        ctx.renderer, ctx.c, ctx.ectx, ctx.bind, ctx.names = renderer, c, ectx, bind, names
        return g

    @staticmethod
    def backward(ctx, g_g):
        if ctx.bind is not None:
            flat = _bind_views(ctx.renderer, ctx.bind)
            ctx.renderer.eikonal_backward(ctx.c, ctx.ectx, g_g, flat=flat)
            return (None,) * (5 + len(ctx.names))
        views = ctx.renderer.eikonal_backward(ctx.c, ctx.ectx, g_g)
        return (None,) * 5 + tuple(views[n] for n in ctx.names)


def eikonal(renderer, c, P, ray_ids, bind_grads=False):
    """Differentiable sdf gradient on all samples of the rays `ray_ids` (flat indices into the (M, N) batch of the training
    forward whose context is `c`, `query_field(..., return_ctx=True)`): (n_rays, D, 3), with autograd edges to the
    basefield weights and sdf.weight of P.  bind_grads as in `query_field`."""
    names = renderer.eikonal_weight_names()
    bind = {k: v for k, v in P.items() if v.requires_grad and v.is_leaf} if bind_grads else None
    return EikonalFunction.apply(renderer, c, ray_ids, bind, names, *[P[n] for n in names])


def query_field(renderer, P, rays, tab, D, flow_thresh=None, depth=None, bind_grads=False, return_ctx=False):
    """Differentiable training-mode query_field: (feat_dict, deltas) like FieldRenderer.query_field, with autograd edges
    to P's tensors, tab's tensors and rays['Kinv'].  Call renderer.pack_train(P, alpha) first (every optimiser step).
    bind_grads: P's tensors are leaf parameters; their .grad become views of the renderer's flat gradient buffer and the
    backward kernels accumulate into it directly (no per-parameter gradient tensors, one buffer to all-reduce)."""
    p_names = [k for k in P]
    t_names = [k for k, v in tab.items() if torch.is_tensor(v)]
    meta = dict(p_names=p_names, t_names=t_names, rays={k: v for k, v in rays.items() if k != "Kinv"}, D=int(D), flow_thresh=flow_thresh, depth=depth)
    if bind_grads:
        meta["bind"] = {k: v for k, v in P.items() if v.requires_grad and v.is_leaf}
    res = FieldFunction.apply(renderer, meta, *[P[k] for k in p_names], *[tab[k] for k in t_names], rays["Kinv"])
    deltas, outs = res[0], res[1:]
    feat = dict(zip(meta["out_keys"], outs))
    feat["density_" + renderer.cfg.category] = feat["density"]  # the reference hands out the same tensor (nerf.py:809-812)
    if return_ctx:
        return feat, deltas, meta["ctx"]
    return feat, deltas
