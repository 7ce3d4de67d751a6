"""Build container only: what 16-bit GEMM operands do to the gradients of NeRF.forward (fg field, synthetic weights,
e^{logibeta} = 20).  Uses the hand-derived backward (oracle/nerf_backward.py); reference = the same in fp64.
  A. exact forward (masks, activations), backward GEMM operands rounded           -> the backward kernel's own rounding
  B. forward GEMM operands rounded (what the forward kernel does), exact backward  -> what the saved forward state costs
Measured (rel-L2 of the gradient): A fp16 6e-4..1e-3, bf16 5e-3..8e-3;  B fp16 3.4e-2 (median over weight tensors),
6e-2 for basefield.linear_1.  B is entirely the 1.8e-4 of ReLU units whose mask flips (pre-activation within the forward's
rounding of zero): a flipped unit contributes ~0 to the forward but its full path to the backward, so the error is
~sqrt(flip fraction x layers).  With exact masks, rounded activations and sdf cost only 4.8e-4.  The gradient of a ReLU
network is discontinuous at those points, so this is not an accuracy defect of the kernel, but a parity test against the
fp32 reference's gradients must either use the kernel's own masks in the checker or accept ~3 % rel-L2."""
import sys

import torch

sys.path.insert(0, "oracle"); sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import nerf_backward as NB  # noqa: E402
from lab4d_b200 import spec  # noqa: E402
from util import synth_params  # noqa: E402


def main():
    cfg = spec.FG_RIGID
    ocfg = cfg.as_oracle_cfg()
    g = torch.Generator().manual_seed(4)
    M, S = 4, 512
    rnd = lambda *s, sc=1.0: sc * torch.randn(*s, generator=g, dtype=torch.float64)
    x, ib, ic, ap = rnd(M, S, 3, sc=0.2), rnd(M, 32, sc=0.5), rnd(M, 32, sc=0.5), rnd(M, 32)
    g_rgb, g_dens = rnd(M, S, 3), rnd(M, S, 1, sc=0.01)
    P = synth_params(cfg, 2, torch.float64)
    _, _, _, sv = NB.nerf_forward_saved(P, ocfg, x, ib, ic, None, ap)
    gin64, gp64 = NB.nerf_backward(P, ocfg, x, sv, g_rgb, g_dens)
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
    oe, om = torch.einsum, torch.Tensor.__matmul__

    def report(tag, gin, gp):
        errs = {k: rel(gp[k], gp64[k]) for k in gp64}
        print(f"{tag:46s} g_x {rel(gin['x'], gin64['x']):.2e}  weight grads median {sorted(errs.values())[len(errs) // 2]:.2e}  "
              f"basefield.linear_1 {errs['basefield.linear_1.0.weight']:.2e}  sdf.weight {errs['sdf.weight']:.2e}")

    for dt in (torch.float16, torch.bfloat16):
        rd = lambda t: t.to(dt).to(torch.float64)
        torch.einsum = lambda eq, a, b: oe(eq, rd(a), rd(b))
        torch.Tensor.__matmul__ = lambda a, b: om(rd(a), rd(b))
        try:
            gin, gp = NB.nerf_backward(P, ocfg, x, sv, g_rgb, g_dens)
        finally:
            torch.einsum, torch.Tensor.__matmul__ = oe, om
        report(f"A {dt}: rounded backward GEMMs", gin, gp)
        torch.Tensor.__matmul__ = lambda a, b: om(rd(a), rd(b))
        try:
            _, _, _, sv2 = NB.nerf_forward_saved(P, ocfg, x, ib, ic, None, ap)
        finally:
            torch.Tensor.__matmul__ = om
        gin, gp = NB.nerf_backward(P, ocfg, x, sv2, g_rgb, g_dens)
        flips = sum(float((a[1] != b[1]).double().mean()) for a, b in zip(sv["sb"]["layers"][:-1], sv2["sb"]["layers"][:-1])) / cfg.D
        report(f"B {dt}: rounded forward (mask flips {flips:.1e})", gin, gp)


if __name__ == "__main__":
    main()
