"""GPU parity of the normals kernel (b200r_field_normals, NeRF.compute_normal nnutils/nerf.py:455-493): d sdf / d xyz_cam at every
sample - through the basefield and the backward warp - against autograd through the reference-pinned oracle in fp64.
Single 16-bit operands with the tape's ReLU signs: rel-L2 <= 1e-2 with the split-operand forward (cf. tests/test_gpu_eikonal.py:
per-point median ~1e-3, a few samples with a flipped sign)."""
import pytest
import torch

import lab4d_oracle as O
import synth
from test_gpu_parity import synth_tables
from util import rel_l2, synth_params

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _oracle_g(cfg, P, tab, xyz_cam, dtype=torch.float64):
    cv = lambda v: v.detach().to(dtype) if v.dtype.is_floating_point else v
    Pd, td = {k: cv(v) for k, v in P.items()}, {k: cv(v) for k, v in tab.items()}
    oc = cfg.as_oracle_cfg()
    with torch.enable_grad():
        xc = xyz_cam.detach().to(dtype).requires_grad_(True)
        xyz_t, _ = O.cam_to_field(xc, torch.zeros_like(xc), td["field2cam_q"], td["field2cam_t"])
        xyz = xyz_t
        if cfg.motion != "rigid":
            xyz, _ = O.skinning_warp(Pd, xyz_t, (td["t_articulation_qr"], td["t_articulation_qd"]), (td["rest_articulation_qr"], td["rest_articulation_qd"]),
                                     td["skin_t_embed"], td["skin_t_embed_mean"], td["inst_skin"], backward=True, symm_idx=oc.get("symm_idx"))
            if cfg.dense:
                xyz = O.dense_warp(Pd, xyz, td["dense_t_embed"], td["inst_dense_bwd"], backward=True)
        sdf = O.nerf_forward(Pd, oc, xyz, td["inst_base"], None, get_density=False)
        (g,) = torch.autograd.grad(sdf.sum(), xc)
    return g


@pytest.mark.parametrize("name,M,N,D", [("bg", 4, 24, 33), ("fg_rigid", 2, 16, 40), ("fg_bob", 4, 16, 48), ("fg_compquad", 4, 8, 32), ("fg_bob", 8, 16, 128)])
def test_normals_kernel_matches_autograd_through_the_warp(name, M, N, D):
    from lab4d_b200 import spec
    from lab4d_b200.render import FieldRenderer

    cfg = {"bg": spec.BG, "fg_rigid": spec.FG_RIGID, "fg_bob": spec.FG_BOB, "fg_compquad": spec.FG_COMP_QUAD}[name]
    P = synth_params(cfg, 3, device=DEV)
    rays = {k: torch.from_numpy(v).to(DEV) for k, v in synth.synth_rays(M, N, seed=13).items()}
    tab = {k: v.clone() for k, v in synth_tables(cfg, M, DEV, seed=13, rays=rays, P=P).items()}
    r = FieldRenderer(cfg, DEV, operand_dtype="fp16x3")
    r.pack_train(P)
    feat, deltas, ctx = r.query_field_train(P, rays, tab, D)
    g = r.sdf_gradient_cam(ctx)
    torch.cuda.synchronize()
    assert torch.isfinite(g).all()
    g64 = _oracle_g(cfg, P, tab, feat["xyz_cam"])
    e = rel_l2(g.cpu(), g64.cpu())
    pp = ((g.double() - g64).norm(dim=-1) / g64.norm(dim=-1)).flatten().sort().values
    n1, n2 = torch.nn.functional.normalize(g.double(), dim=-1), torch.nn.functional.normalize(g64, dim=-1)
    ang = torch.rad2deg(torch.acos((n1 * n2).sum(-1).clamp(-1, 1))).flatten().sort().values
    print(f"[normals] {name} {M}x{N}x{D}: g rel-L2 {e:.2e}, per sample median {float(pp[len(pp) // 2]):.1e} p99 {float(pp[int(0.99 * len(pp))]):.1e}; "
          f"normal angle median {float(ang[len(ang) // 2]):.3f} deg, p99 {float(ang[int(0.99 * len(ang))]):.3f} deg")
    assert e <= 1e-2, e
    assert float(ang[len(ang) // 2]) < 0.2
