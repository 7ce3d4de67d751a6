"""CPU: the hand-derived backward of the blend-skinning warp (oracle/skin_backward.py, table formulation of the CUDA
kernel) against autograd in fp64, and its forward against the reference-pinned oracle."""
import numpy as np
import torch

import lab4d_oracle as O
import skin_backward as SB
from lab4d_b200 import spec
from util import synth_params


def _problem(B, backward, dtype=torch.float64, seed=0, symm=None):
    cfg = spec.FieldConfig(motion="bob" if B == 25 else "skel", B=B, symm_idx=symm)
    P = synth_params(cfg, seed, dtype)
    rs = np.random.RandomState(5 + seed)
    M, S = 3, 17
    f = lambda *s, sc=1.0: torch.from_numpy(sc * rs.standard_normal(s)).to(dtype)

    def art(scale_r, scale_t, rows):
        aa = scale_r * rs.standard_normal((rows, B, 3))
        ang = np.linalg.norm(aa, axis=-1, keepdims=True)
        qr = torch.from_numpy(np.concatenate([np.cos(ang / 2), np.sin(ang / 2) * aa / np.maximum(ang, 1e-9)], -1)).to(dtype)
        t = torch.from_numpy(scale_t * rs.standard_normal((rows, B, 3))).to(dtype)
        return qr, 0.5 * O.qmul(t, qr)

    rest = tuple(a.expand(M, -1, -1).contiguous() for a in art(0.2, 0.08, 1))
    t_art = art(0.3, 0.08, M)
    x = f(M, S, 3, sc=0.15)
    t_embed, t_mean, inst = f(M, 128), f(1, 128, sc=0.5), f(1, 32, sc=0.5).expand(M, -1)
    return cfg, P, x, t_art, rest, t_embed, t_mean, inst


def _tables(cfg, P, t_art, rest, t_embed, t_mean, inst, backward):
    Rp, tp, se3_r, se3_d = SB.tables_from_articulation(P, t_art, rest, backward, cfg.symm_idx and list(cfg.symm_idx))
    pre = "warp.skinning_model.delta_field."
    W1, b1 = P[pre + "linear_1.0.weight"], P[pre + "linear_1.0.bias"]
    B = cfg.B
    te = t_embed if backward else t_mean.expand(t_embed.shape[0], -1)
    b1row = b1 + te @ W1[:, 3 * B:3 * B + 128].T + inst @ W1[:, 3 * B + 128:].T
    return dict(Rp=Rp, tp=tp, se3_r=se3_r, se3_d=se3_d, W1x=W1[:, :3 * B].contiguous(), b1row=b1row, W2=P[pre + "linear_2.0.weight"],
                b2=P[pre + "linear_2.0.bias"], W3=P[pre + "linear_final.weight"], b3=P[pre + "linear_final.bias"])


def test_table_forward_equals_reference_pinned_oracle():
    for B, backward, symm in ((25, True, None), (25, False, None), (18, True, spec.HUMAN_SYMM)):
        cfg, P, x, t_art, rest, t_embed, t_mean, inst = _problem(B, backward, symm=symm)
        T = _tables(cfg, P, t_art, rest, t_embed, t_mean, inst, backward)
        xo, ent, dsk, _ = SB.skin_forward_tables(x, **T)
        ref, aux = O.skinning_warp(P, x[:, :, None, :], t_art, rest, t_embed, t_mean, inst, backward,
                                   symm_idx=cfg.symm_idx and list(cfg.symm_idx))
        assert (xo - ref[:, :, 0]).abs().max() < 1e-12
        assert (ent - aux["skin_entropy"][:, :, 0, 0]).abs().max() < 1e-12
        assert (dsk - aux["delta_skin"][:, :, 0, 0]).abs().max() < 1e-12


def test_hand_derived_backward_equals_autograd():
    for B, backward in ((25, True), (25, False), (18, False)):
        cfg, P, x, t_art, rest, t_embed, t_mean, inst = _problem(B, backward, seed=1)
        T = {k: v.detach().clone().requires_grad_(True) for k, v in _tables(cfg, P, t_art, rest, t_embed, t_mean, inst, backward).items()}
        xg = x.clone().requires_grad_(True)
        xo, ent, dsk, saved = SB.skin_forward_tables(xg, **T)
        g = torch.Generator().manual_seed(3)
        g_xo, g_ent, g_dsk = (torch.randn(xo.shape, generator=g, dtype=xo.dtype), torch.randn(ent.shape, generator=g, dtype=xo.dtype),
                              torch.randn(dsk.shape, generator=g, dtype=xo.dtype))
        ((g_xo * xo).sum() + (g_ent * ent).sum() + (g_dsk * dsk).sum()).backward()
        with torch.no_grad():
            hand = SB.skin_backward_tables(xg.detach(), **{k: v.detach() for k, v in T.items()},
                                           saved={k: v.detach() for k, v in saved.items()}, g_xo=g_xo, g_ent=g_ent, g_dsk=g_dsk)
        auto = dict(x=xg.grad, **{k: v.grad for k, v in T.items()})
        for k, a in auto.items():
            err = float((hand[k] - a).abs().max() / (a.abs().max() + 1e-30))
            assert err < 1e-9, (B, backward, k, err)
