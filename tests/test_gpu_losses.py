"""GPU parity of the loss kernels (b200r_loss_fwd / b200r_loss_bwd: the reconstruction part of dvr_model.compute_loss,
lab4d/engine/model.py:374-611) against the reference-pinned oracle (oracle/loss_oracle.recon_losses, fp64) on the same
synthetic batch: every weighted term and the gradient of their sum w.r.t. every rendered input.  fp32 sums: 1e-5."""
import pytest
import torch

import loss_oracle as LO
from test_loss_oracle_cpu import CONFIG, synth_loss_inputs
from util import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _leaves(d):
    return {k: (_leaves(v) if isinstance(v, dict) else (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v)) for k, v in d.items()}


@pytest.mark.parametrize("field_type,M,N", [("fg", 6, 16), ("bg", 4, 64), ("comp", 8, 32), ("fg", 128, 16), ("comp", 256, 16)])
def test_loss_kernels_match_the_oracle(field_type, M, N):
    from lab4d_b200.render import recon_losses

    config = dict(CONFIG, field_type=field_type)
    rendered, aux, batch = synth_loss_inputs(field_type, M, N, seed=M, device=DEV)
    if "fg" in aux:
        rendered["gauss_mask"] = aux["fg"]["gauss_mask"]
    r1, a1 = _leaves(rendered), _leaves(aux)
    if "fg" in a1:
        r1["gauss_mask"] = a1["fg"]["gauss_mask"]
    ours = recon_losses(r1, a1, batch, config)
    sum(ours.values()).backward()
    torch.cuda.synchronize()
    to64 = lambda d: {k: (to64(v) if isinstance(v, dict) else (v.detach().double().requires_grad_(True) if v.dtype.is_floating_point else v)) for k, v in d.items()}
    r2, a2 = to64(rendered), to64(aux)
    if "fg" in a2:
        r2["gauss_mask"] = a2["fg"]["gauss_mask"]
    b2 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in batch.items()}
    ref = LO.recon_losses(r2, a2, b2, field_type, config)
    sum(ref.values()).backward()
    assert list(ours) == list(ref)
    rows = []
    for k in ref:
        e = abs(float(ours[k]) - float(ref[k])) / abs(float(ref[k]))
        rows.append(f"{k}={e:.1e}")
        assert e <= 1e-5, (k, float(ours[k]), float(ref[k]))
    grows = []
    for tag, mine, theirs in [("rendered", r1, r2)] + [("aux." + c, a1[c], a2[c]) for c in a1]:
        for k, v in theirs.items():
            if not torch.is_tensor(v) or v.grad is None or float(v.grad.abs().max()) == 0.0:
                continue
            assert mine[k].grad is not None, (tag, k)
            e = rel_l2(mine[k].grad.cpu(), v.grad.cpu())
            grows.append(f"{tag}.{k}={e:.1e}")
            assert e <= 1e-5, (tag, k, e)
    print(f"[losses] {field_type} {M}x{N}: " + " ".join(rows) + " | grads " + " ".join(grows))
    assert len(grows) >= 5
