"""CPU: the eikonal term as three masked linear chains (oracle/eikonal_backward.py - the formulation of the eikonal
kernels) against the reference's way (autograd.grad with create_graph, then backward through it) in fp64: the sdf gradient
g and the gradient of a loss of g w.r.t. every basefield weight and sdf.weight; biases and codes receive nothing."""
import pytest
import torch

import eikonal_backward as EB
from lab4d_b200 import spec
from util import synth_params


@pytest.mark.parametrize("cfg,alpha", [(spec.FG_RIGID, None), (spec.BG, None), (spec.FG_RIGID, 0.55)], ids=["fg", "bg", "fg_window"])
def test_eikonal_chains_match_double_backward(cfg, alpha):
    P = {k: v.requires_grad_(True) for k, v in synth_params(cfg, 3, torch.float64).items()}
    Pd = {k: v.detach() for k, v in P.items()}
    gen = torch.Generator().manual_seed(7)
    R, D = 5, 7
    x = 0.2 * torch.randn(R, D, 3, generator=gen, dtype=torch.float64)
    inst = (0.5 * torch.randn(R, 32, generator=gen, dtype=torch.float64)).requires_grad_(True)
    ocfg = dict(cfg.as_oracle_cfg(), skip=cfg.skip)
    coeff = torch.rand(R, D, generator=gen, dtype=torch.float64)
    loss_of = lambda g: (coeff * (g.norm(2, dim=-1) - 1) ** 2).sum()  # (|g| - 1)^2 like compute_eikonal, weighted

    g_ref = EB.sdf_gradient_autograd(P, ocfg, x, inst, alpha)
    loss_of(g_ref).backward()

    def gbar_fn(g):
        gg = g.detach().requires_grad_(True)
        with torch.enable_grad():
            (gb,) = torch.autograd.grad(loss_of(gg), gg)
        return gb

    with torch.no_grad():
        g, grads = EB.eikonal_hand(Pd, ocfg, x, inst.detach(), gbar_fn, alpha)
    rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-30))
    assert rel(g, g_ref.detach()) < 1e-10
    n = 0
    for k, gk in grads.items():
        assert P[k].grad is not None, k
        assert rel(gk, P[k].grad) < 1e-9, (k, rel(gk, P[k].grad))
        n += 1
    assert n == cfg.D + 2
    # nothing else moves: biases and the instance code only enter through the (piecewise constant) masks
    for k, v in P.items():
        if k not in grads and v.grad is not None:
            assert float(v.grad.abs().max()) == 0.0, k
    assert inst.grad is None or float(inst.grad.abs().max()) == 0.0
