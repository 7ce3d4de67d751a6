"""torch.autograd glue of the training path: one Function whose forward is b200r_field_fwd_train and whose backward is
b200r_field_bwd + the per-frame chain (lab4d_b200/prologue_grad.py).  Differentiable inputs: every hot-path parameter
(state_dict names of lab4d_b200/spec.py), every per-frame table of `tab` and rays["Kinv"]; hxy / near_far are data."""
import torch

from . import _lib

_NO_GRAD_OUT = ("deltas", "eikonal", "xyz_t", "dir")


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
        ctx.renderer, ctx.c, ctx.keys, ctx.meta = renderer, c, keys, meta
        outs = [feat[k] for k in keys]
        ctx.mark_non_differentiable(deltas, *[o for k, o in zip(keys, outs) if k in _NO_GRAD_OUT])
        return (deltas, *outs)

    @staticmethod
    def backward(ctx, g_deltas, *g_outs):
        grads = {k: g for k, g in zip(ctx.keys, g_outs) if g is not None and k in _lib.GRAD_KEYS}
        pg, tg = ctx.renderer.backward(ctx.c, grads)
        meta = ctx.meta
        out = [None, None]
        for n in meta["p_names"]:
            out.append(pg.get(n))
        for n in meta["t_names"]:
            g = tg.get(n)
            out.append(g if g is None else g.reshape(ctx.c["tab"][n].shape))
        out.append(tg.get("Kinv"))
        return tuple(out)


def query_field(renderer, P, rays, tab, D, flow_thresh=None, depth=None):
    """Differentiable training-mode query_field: (feat_dict, deltas) like FieldRenderer.query_field, with autograd edges
    to P's tensors, tab's tensors and rays['Kinv'].  Call renderer.pack_train(P, alpha) first (every optimiser step)."""
    p_names = [k for k in P]
    t_names = [k for k, v in tab.items() if torch.is_tensor(v)]
    meta = dict(p_names=p_names, t_names=t_names, rays={k: v for k, v in rays.items() if k != "Kinv"}, D=int(D), flow_thresh=flow_thresh, depth=depth)
    res = FieldFunction.apply(renderer, meta, *[P[k] for k in p_names], *[tab[k] for k in t_names], rays["Kinv"])
    deltas, outs = res[0], res[1:]
    feat = dict(zip(meta["out_keys"], outs))
    feat["density_" + renderer.cfg.category] = feat["density"]  # the reference hands out the same tensor (nerf.py:809-812)
    return feat, deltas
