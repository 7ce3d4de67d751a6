"""CPU: the composed hand-derived backward of a skinned fg field's query_field (oracle/field_backward.py) against autograd
through the reference-pinned oracle, fp64.  The per-sample part is by hand; the per-frame part (tables <- articulations /
cameras / codes) is chained with autograd on the small per-frame functions, as planned for the product."""
import numpy as np
import torch

import field_backward as FB
import lab4d_oracle as O
import synth
from lab4d_b200 import spec
from test_skin_backward_cpu import _problem
from util import synth_params

KEYS = ("rgb", "density", "vis", "feature", "flow", "cyc_dist", "xyz", "xyz_cam", "delta_skin", "skin_entropy", "gauss_density")


def _setup(dtype=torch.float64, dense=False):
    cfg = spec.FieldConfig(motion="bob", B=25, dense=True) if dense else spec.FG_BOB
    M, N, D = 2, 3, 5
    P = {k: v.requires_grad_(True) for k, v in synth_params(cfg, 1, dtype).items()}
    rays_np = synth.synth_rays(M, N, seed=3)
    rays = {k: torch.from_numpy(v).to(dtype) for k, v in rays_np.items() if k in ("hxy", "Kinv", "near_far")}
    rays["Kinv"].requires_grad_(True)
    _, _, _, t_art, rest, t_embed, t_mean, inst = _problem(25, True, dtype=dtype, seed=2)
    rs = np.random.RandomState(11)
    f = lambda *s, sc=1.0: torch.from_numpy(sc * rs.standard_normal(s)).to(dtype)
    fc = torch.from_numpy(rays_np["field2cam"]).to(dtype)
    tab = {"field2cam_q": fc[:, :4].contiguous(), "field2cam_t": (fc[:, 4:] * 0.2).contiguous(), "inst_base": f(M, 32, sc=0.5),
           "inst_color": f(M, 32, sc=0.5), "inst_vis": f(M, 32, sc=0.5), "appr_code": f(M, 32), "inst_skin": inst[:M].contiguous(),
           "skin_t_embed": t_embed[:M].contiguous(), "skin_t_embed_mean": t_mean, "t_articulation_qr": t_art[0][:M].contiguous(),
           "t_articulation_qd": t_art[1][:M].contiguous(), "rest_articulation_qr": rest[0][:M].contiguous(),
           "rest_articulation_qd": rest[1][:M].contiguous()}
    if dense:
        tab.update({"dense_t_embed": f(M, 128), "inst_dense_fwd": f(M, 32, sc=0.5), "inst_dense_bwd": f(M, 32, sc=0.5)})
    for v in tab.values():
        v.requires_grad_(True)
    return cfg, P, rays, tab, M, N, D


import pytest


@pytest.mark.parametrize("dense", [False, True], ids=["skinning", "composed"])
def test_forward_formulation_and_full_backward(dense):
    cfg, P, rays, tab, M, N, D = _setup(dense=dense)
    ocfg = cfg.as_oracle_cfg()
    S = N * D
    feat, _ = O.query_field(P, ocfg, rays, tab, D)
    g = torch.Generator().manual_seed(8)
    cot = {k: torch.randn(M, S, feat[k].shape[-1] if k != "flow" else 2, generator=g, dtype=torch.float64) for k in KEYS}
    loss = sum((cot[k] * feat[k].reshape(M, S, -1)[..., :cot[k].shape[-1]]).sum() for k in KEYS)
    leaves = list(P.values()) + list(tab.values()) + [rays["Kinv"]]
    names = list(P.keys()) + ["tab/" + k for k in tab] + ["Kinv"]
    auto = dict(zip(names, torch.autograd.grad(loss, leaves, allow_unused=True)))

    Pd = {k: v.detach() for k, v in P.items()}
    tabd = {k: v.detach() for k, v in tab.items()}
    raysd = {k: v.detach() for k, v in rays.items()}
    with torch.no_grad():
        out, saved = FB.forward_saved(Pd, ocfg, raysd, tabd, D)
        for k in KEYS:  # the kernel's formulation reproduces the reference-pinned oracle
            ref = feat[k].reshape(M, S, -1)[..., :out[k].shape[-1]]
            assert float((out[k] - ref).abs().max()) <= 1e-10 * max(1.0, float(ref.abs().max())), k
        grads, tb = FB.backward(Pd, ocfg, raysd, tabd, saved, cot)

    # per-frame chain with autograd on the small per-frame functions
    outs, gouts = [], []
    for w in range(3):
        T = FB._skin_tables(P, ocfg, tab, M, w)
        for k in ("Rp", "tp", "se3_r", "se3_d", "W1x", "b1row", "W2", "b2", "W3", "b3"):
            outs.append(T[k]); gouts.append(tb["skin"][w][k])
    _, ctr = O.dq_to_qt((tab["rest_articulation_qr"][:1], tab["rest_articulation_qd"][:1]))
    outs.append(ctr); gouts.append(tb["g_ctr"])
    q, t = tab["field2cam_q"], tab["field2cam_t"]
    qi = O.qconj(q)
    outs += [qi, O.qrot(qi, -t), O.flip_pair(q), O.flip_pair(t), O.kmat_from_kinv(O.flip_pair(rays["Kinv"]))]
    gouts += [tb["g_qi"], tb["g_ti"], tb["g_qn"], tb["g_tn"], tb["g_Kmat"]]
    chained = dict(zip(names, torch.autograd.grad(outs, leaves, grad_outputs=gouts, allow_unused=True)))
    hand = {}
    for n in names:
        tot = chained[n]
        key = n
        if key in grads:
            tot = grads[key] if tot is None else tot + grads[key]
        hand[n] = tot
    for n, gname in (("tab/inst_vis", "g_inst_vis"), ("tab/inst_base", "g_inst_base"), ("tab/inst_color", "g_inst_color"), ("tab/appr_code", "g_appr")):
        hand[n] = tb[gname] if hand[n] is None else hand[n] + tb[gname]
    hand["Kinv"] = tb["g_Kinv"] if hand["Kinv"] is None else hand["Kinv"] + tb["g_Kinv"]
    if dense:
        for k, gk in tb["g_dense"].items():
            hand["tab/" + k] = gk if hand["tab/" + k] is None else hand["tab/" + k] + gk

    checked = 0
    for n in names:
        a = auto[n]
        if a is None or float(a.abs().max()) == 0.0:
            assert hand[n] is None or float(hand[n].abs().max()) < 1e-12, n
            continue
        assert hand[n] is not None, n
        err = float((hand[n] - a).abs().max() / a.abs().max())
        assert err < 1e-8, (n, err)
        checked += 1
    assert checked >= 60


def test_rigid_bg_field_with_view_directions():
    """bg field (IdentityWarp, raw view direction into rgb.0, no appearance code, no feature field)."""
    dtype = torch.float64
    cfg = spec.BG
    ocfg = cfg.as_oracle_cfg()
    M, N, D = 2, 3, 5
    S = N * D
    P = {k: v.requires_grad_(True) for k, v in synth_params(cfg, 1, dtype).items()}
    rays_np = synth.synth_rays(M, N, seed=3)
    rays = {k: torch.from_numpy(v).to(dtype) for k, v in rays_np.items() if k in ("hxy", "Kinv", "near_far")}
    rays["Kinv"].requires_grad_(True)
    rs = np.random.RandomState(12)
    f = lambda *s, sc=1.0: torch.from_numpy(sc * rs.standard_normal(s)).to(dtype)
    fc = torch.from_numpy(rays_np["field2cam"]).to(dtype)
    tab = {"field2cam_q": fc[:, :4].contiguous(), "field2cam_t": (fc[:, 4:] * 0.2).contiguous(), "inst_base": f(M, 32, sc=0.5),
           "inst_color": f(M, 32, sc=0.5), "inst_vis": f(M, 32, sc=0.5)}
    for v in tab.values():
        v.requires_grad_(True)
    keys = ("rgb", "density", "vis", "flow", "xyz", "xyz_cam")
    feat, _ = O.query_field(P, ocfg, rays, tab, D)
    g = torch.Generator().manual_seed(9)
    cot = {k: torch.randn(M, S, feat[k].shape[-1] if k != "flow" else 2, generator=g, dtype=dtype) for k in keys}
    loss = sum((cot[k] * feat[k].reshape(M, S, -1)[..., :cot[k].shape[-1]]).sum() for k in keys)
    leaves = list(P.values()) + list(tab.values()) + [rays["Kinv"]]
    names = list(P.keys()) + ["tab/" + k for k in tab] + ["Kinv"]
    auto = dict(zip(names, torch.autograd.grad(loss, leaves, allow_unused=True)))
    Pd, tabd, raysd = ({k: v.detach() for k, v in d.items()} for d in (P, tab, rays))
    with torch.no_grad():
        out, saved = FB.forward_saved(Pd, ocfg, raysd, tabd, D)
        for k in keys:
            ref = feat[k].reshape(M, S, -1)[..., :out[k].shape[-1]]
            assert float((out[k] - ref).abs().max()) <= 1e-10 * max(1.0, float(ref.abs().max())), k
        grads, tb = FB.backward(Pd, ocfg, raysd, tabd, saved, cot)
    q, t = tab["field2cam_q"], tab["field2cam_t"]
    qi = O.qconj(q)
    outs = [qi, O.qrot(qi, -t), O.flip_pair(q), O.flip_pair(t), O.kmat_from_kinv(O.flip_pair(rays["Kinv"]))]
    gouts = [tb["g_qi"], tb["g_ti"], tb["g_qn"], tb["g_tn"], tb["g_Kmat"]]
    chained = dict(zip(names, torch.autograd.grad(outs, leaves, grad_outputs=gouts, allow_unused=True)))
    hand = {n: (grads.get(n) if chained[n] is None else (chained[n] + grads[n] if n in grads else chained[n])) for n in names}
    for n, gname in (("tab/inst_vis", "g_inst_vis"), ("tab/inst_base", "g_inst_base"), ("tab/inst_color", "g_inst_color")):
        hand[n] = tb[gname] if hand[n] is None else hand[n] + tb[gname]
    hand["Kinv"] = tb["g_Kinv"] if hand["Kinv"] is None else hand["Kinv"] + tb["g_Kinv"]
    checked = 0
    for n in names:
        a = auto[n]
        if a is None or float(a.abs().max()) == 0.0:
            continue
        assert hand[n] is not None, n
        err = float((hand[n] - a).abs().max() / a.abs().max())
        assert err < 1e-8, (n, err)
        checked += 1
    assert checked >= 25
