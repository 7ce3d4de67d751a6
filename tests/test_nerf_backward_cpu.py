"""CPU: hand-derived backward of NeRF.forward (oracle/nerf_backward.py) against autograd through the reference-pinned
oracle, fp64: gradients w.r.t. the points, the per-frame codes and every parameter of basefield / colorfield / heads."""
import torch

import lab4d_oracle as O
import nerf_backward as NB
from lab4d_b200 import spec
from util import synth_params


def _run(cfg, with_dir):
    P = {k: v.requires_grad_(True) for k, v in synth_params(cfg, 2, torch.float64).items()}
    g = torch.Generator().manual_seed(4)
    M, S = 2, 19
    rnd = lambda *s, sc=1.0: (sc * torch.randn(*s, generator=g, dtype=torch.float64))
    x = rnd(M, S, 3, sc=0.2).requires_grad_(True)
    dirs = torch.nn.functional.normalize(rnd(M, S, 3), dim=-1) if with_dir else None
    inst_b, inst_c = rnd(M, 32, sc=0.5).requires_grad_(True), rnd(M, 32, sc=0.5).requires_grad_(True)
    appr = rnd(M, cfg.appr_channels).requires_grad_(True) if cfg.appr_channels else None
    ocfg = cfg.as_oracle_cfg()
    rgb_o, dens_o = O.nerf_forward(P, ocfg, x[:, :, None], inst_b, inst_c, dir=(dirs if with_dir else x.detach())[:, :, None], appr=appr)
    rgb, dens, sdf, saved = NB.nerf_forward_saved({k: v.detach() for k, v in P.items()}, ocfg, x.detach(), inst_b.detach(), inst_c.detach(),
                                                  dirs, None if appr is None else appr.detach())
    assert (rgb - rgb_o[:, :, 0]).abs().max() < 1e-12 and (dens - dens_o[:, :, 0]).abs().max() < 1e-12
    g_rgb, g_dens = rnd(M, S, 3), rnd(M, S, 1)
    ((g_rgb * rgb_o[:, :, 0]).sum() + (g_dens * dens_o[:, :, 0]).sum()).backward()
    with torch.no_grad():
        gin, gpar = NB.nerf_backward({k: v.detach() for k, v in P.items()}, ocfg, x.detach(), saved, g_rgb, g_dens)
    rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-30))
    assert rel(gin["x"], x.grad) < 1e-9
    assert rel(gin["inst_base"], inst_b.grad) < 1e-9 and rel(gin["inst_color"], inst_c.grad) < 1e-9
    if appr is not None:
        assert rel(gin["appr"], appr.grad) < 1e-9
    checked = 0
    for k, gk in gpar.items():
        assert P[k].grad is not None, k
        assert rel(gk, P[k].grad) < 1e-9, k
        checked += 1
    assert checked >= 2 * (cfg.D + 1) + 2 * 3 + 4 + 1


def test_fg_field_backward_matches_autograd():
    _run(spec.FG_RIGID, with_dir=False)


def test_bg_field_with_directions_backward_matches_autograd():
    _run(spec.BG, with_dir=True)
