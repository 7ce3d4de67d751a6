"""GPU parity of the differentiable forward warp of points (b200r_warp_fwd_train / b200r_warp_bwd, lab4d_b200.autograd.warp_points:
Deformable.forward_warp's skinning warp as FeatureNeRF.forward_project runs it on the matched points, lab4d/nnutils/
deformable.py:154-171, feature.py:207-226) against autograd through the reference-pinned oracle (skinning_warp / dense_warp,
fp64): warped points, and the gradients w.r.t. the points, the warp's parameters and the per-frame tables.

Tolerances as for the field backward (tests/test_gpu_backward.py): the delta MLP runs on scaled fp16 gradient rows."""
import pytest
import torch

import lab4d_oracle as O
import synth
from test_gpu_parity import synth_tables
from util import rel_l2, synth_params

pytestmark = pytest.mark.gpu
DEV = "cuda"
TABS = ("inst_skin", "skin_t_embed", "skin_t_embed_mean", "dense_t_embed", "inst_dense_fwd", "t_articulation_qr", "t_articulation_qd",
        "rest_articulation_qr", "rest_articulation_qd")


def _oracle(cfg, P, tab, xyz, coeff, dtype):
    cv = lambda v: v.detach().to(dtype).clone().requires_grad_(True) if v.dtype.is_floating_point else v
    Pg, tg, x = {k: cv(v) for k, v in P.items()}, {k: cv(v) for k, v in tab.items()}, cv(xyz)
    oc = cfg.as_oracle_cfg()
    x_in = O.dense_warp(Pg, x[:, :, None], tg["dense_t_embed"], tg["inst_dense_fwd"], backward=False) if cfg.dense else x[:, :, None]
    out, _ = O.skinning_warp(Pg, x_in, (tg["t_articulation_qr"], tg["t_articulation_qd"]), (tg["rest_articulation_qr"], tg["rest_articulation_qd"]),
                             tg["skin_t_embed"], tg["skin_t_embed_mean"], tg["inst_skin"], backward=False, symm_idx=oc.get("symm_idx"))
    out = out[:, :, 0]
    (coeff.to(dtype) * out).sum().backward()
    grads = {"xyz": x.grad}
    grads.update({k: v.grad for k, v in Pg.items() if v.grad is not None and float(v.grad.abs().max()) > 0})
    grads.update({"tab:" + k: v.grad for k, v in tg.items() if torch.is_tensor(v) and v.dtype.is_floating_point and v.grad is not None and float(v.grad.abs().max()) > 0})
    return out.detach(), grads


@pytest.mark.parametrize("name,M,N,prec", [("fg_bob", 4, 16, "fp16x3"), ("fg_skelhuman", 4, 40, "fp16x3"), ("fg_compquad", 2, 96, "fp16x3"), ("fg_bob", 128, 16, "fp16")])
def test_point_warp_and_its_backward(name, M, N, prec):
    from lab4d_b200 import autograd as ag
    from lab4d_b200 import spec
    from lab4d_b200.render import FieldRenderer

    cfg = {"fg_bob": spec.FG_BOB, "fg_skelhuman": spec.FG_SKEL_HUMAN, "fg_compquad": spec.FG_COMP_QUAD}[name]
    P = {k: v.requires_grad_(True) for k, v in synth_params(cfg, 3, device=DEV).items()}
    rays = {k: torch.from_numpy(v).to(DEV) for k, v in synth.synth_rays(M, 4, seed=5).items()}
    tab = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in synth_tables(cfg, M, DEV, seed=5, rays=rays, P=P).items()}
    g = torch.Generator().manual_seed(M + N)
    xyz = (0.12 * torch.randn(M, N, 3, generator=g)).to(DEV).requires_grad_(True)
    coeff = torch.randn(M, N, 3, generator=g).to(DEV) / (M * N)
    r = FieldRenderer(cfg, DEV, operand_dtype=prec)
    r.pack_train({k: v.detach() for k, v in P.items()})
    out = ag.warp_points(r, P, xyz, tab)
    (coeff * out).sum().backward()
    torch.cuda.synchronize()
    # same result as the inference entry
    inf, _ = r.warp_points({k: v.detach() for k, v in P.items()}, xyz.detach(), {k: (v.detach() if torch.is_tensor(v) else v) for k, v in tab.items()}, backward=False)
    assert rel_l2(out.detach().cpu(), inf.view(M, N, 3).cpu()) <= 1e-6
    ref, g64 = _oracle(cfg, P, tab, xyz, coeff, torch.float64)
    _, g32 = _oracle(cfg, P, tab, xyz, coeff, torch.float32)
    e_out = rel_l2(out.detach().cpu(), ref.cpu())
    ours = {"xyz": xyz.grad}
    ours.update({k: v.grad for k, v in P.items() if v.grad is not None})
    ours.update({"tab:" + k: v.grad for k, v in tab.items() if torch.is_tensor(v) and v.grad is not None})
    rows, bad = [], []
    scale = 1.0 if prec == "fp16x3" else 20.0
    for k, gr in sorted(g64.items()):
        assert k in ours and ours[k] is not None, f"missing gradient {k}"
        e, floor = rel_l2(ours[k].reshape(gr.shape).cpu(), gr.cpu()), rel_l2(g32[k].cpu(), gr.cpu())
        rows.append(f"{k.replace('warp.skinning_model.', '').replace('warp.post_warp.', '')}={e:.1e}/{floor:.1e}")
        if not e <= max(2e-2 * scale, 4 * floor):
            bad.append((k, e, floor))
    print(f"[warp-points] {name} {M}x{N} {prec}: warped points {e_out:.1e}; gradients (ours vs fp64 / reference fp32 vs fp64): " + " ".join(rows))
    assert e_out <= (1e-4 if prec == "fp16x3" else 5e-3), e_out
    assert not bad, bad
    assert len(rows) >= 8
