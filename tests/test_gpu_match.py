"""GPU parity of the match kernels (b200r_match_fwd / b200r_match_bwd, FeatureNeRF.global_match nnutils/feature.py:152-205)
against the reference-pinned oracle (oracle/lab4d_oracle.global_match) in fp32 / fp64 on the same candidate draw: matched
points, and the gradients w.r.t. the canonical features, the canonical points and logsigma.  fp32 SIMT: tolerance 1e-5."""
import pytest
import torch

import lab4d_oracle as O
from util import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("M,N,D,K", [(4, 8, 16, 1024), (16, 16, 32, 1024), (128, 16, 128, 1024), (3, 5, 7, 64)])
def test_match_kernels_match_the_oracle(M, N, D, K):
    from lab4d_b200.render import global_match

    g = torch.Generator().manual_seed(M + D)
    feat_px = torch.nn.functional.normalize(torch.randn(M, N, 16, generator=g), dim=-1).to(DEV)
    fc = torch.nn.functional.normalize(torch.randn(M, N, D, 16, generator=g), dim=-1).to(DEV).requires_grad_(True)
    xyz = (0.2 * torch.randn(M, N, D, 3, generator=g)).to(DEV).requires_grad_(True)
    ls = torch.tensor([1.3], device=DEV, requires_grad=True)
    coeff = torch.randn(M, N, 3, generator=g).to(DEV)
    torch.manual_seed(17)
    out = global_match(feat_px, fc, xyz, ls, num_candidates=K)
    (coeff * out).sum().backward()
    torch.cuda.synchronize()
    torch.manual_seed(17)
    idx = torch.randperm(M * N * D)[:min(K, M * N * D)].to(DEV)
    fc2, xyz2, ls2 = (t.detach().double().requires_grad_(True) for t in (fc, xyz, ls))
    ref = O.global_match(feat_px.double(), fc2, xyz2, ls2, idx)
    (coeff.double() * ref).sum().backward()
    errs = {"xyz_matched": rel_l2(out.cpu(), ref.cpu()), "g_feature": rel_l2(fc.grad.cpu(), fc2.grad.cpu()),
            "g_xyz": rel_l2(xyz.grad.cpu(), xyz2.grad.cpu()), "g_logsigma": rel_l2(ls.grad.cpu(), ls2.grad.cpu())}
    print(f"[match] {M}x{N}x{D} K={min(K, M * N * D)}: " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()))
    assert all(v <= 1e-5 for v in errs.values()), errs
    # rows outside the candidate set get exactly zero
    mask = torch.ones(M * N * D, dtype=torch.bool, device=DEV)
    mask[idx] = False
    assert float(fc.grad.reshape(-1, 16)[mask].abs().max()) == 0.0 if mask.any() else True
