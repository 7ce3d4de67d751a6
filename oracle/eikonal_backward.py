"""Hand-derived eikonal term of NeRF.compute_eikonal (lab4d/nnutils/nerf.py:416-453, lab4d/utils/torch_utils.py:4-28):
g = d sdf / d x through the Fourier embedding and the basefield, and the gradient of a loss L(g) w.r.t. the weights - the
reference's second-order autograd (create_graph=True) written as three masked LINEAR chains, the formulation the
eikonal kernels run (csrc/field_bwd.cu eikonal modes + the weight-gradient kernel).  TEST INFRASTRUCTURE: only tests/
import this; tests/test_eikonal_cpu.py checks it against autograd (double backward) in fp64.

With h_l = relu(W_l h_{l-1} + b_l), masks m_l = [z_l > 0], sdf = w . h_F + b:
  a-pass (reverse):  a_F = m_F * w,  a_{l-1} = m_{l-1} * (W_l^T a_l)  [the skip layer hands its embedding columns to u],
                     u = embedding columns of W_1^T a_1 + W_skip^T a_skip,   g = E(x)^T u        (E = d embed / d x, 3 columns)
  given gbar = dL/dg:
  v-pass (forward):  v_0 = E(x) gbar,  v_l = m_l * (W_l v_{l-1})  - no biases, the masks are GIVEN, so the chain is linear and
                     the skip concat splits into chain A (from v_0 through every layer, the skip layer taking only its hidden
                     columns) plus chain B (the skip layer's embedding columns applied to v_0, then the layers after it)
  weight gradients:  dW_l = a_l (vA_{l-1} + vB_{l-1})^T,  dW_1[:, :n_pe] = a_1 v_0^T,  dW_skip[:, :n_pe] = a_skip v_0^T,
                     dw_sdf = vA_F + vB_F;   biases and instance codes receive nothing (they only move the masks).
"""
import torch

import lab4d_oracle as O


def pe_jacobian_apply(x, L, gbar, alpha=None):
    """E(x) gbar: directional derivative of pos_embed (nnutils/embedding.py:69-125) along gbar, (..., 3(2L+1))."""
    cols = [gbar]
    w = None
    if alpha is not None:
        import math

        w = torch.clamp(alpha * L - torch.arange(L, dtype=x.dtype), 0.0, 1.0)
        w = 0.5 * (1 + torch.cos(math.pi * w + math.pi))
    for k in range(L):
        f = 2.0 ** k
        wk = 1.0 if w is None else w[k]
        cols.append(wk * f * torch.cos(f * x) * gbar)
        cols.append(-wk * f * torch.sin(f * x) * gbar)
    return torch.cat(cols, -1)


def pe_jacobian_t_apply(x, L, u, alpha=None):
    """E(x)^T u (the embedding backward of oracle/nerf_backward.pe_backward, with the annealing window)."""
    g = u[..., :3].clone()
    w = None
    if alpha is not None:
        import math

        w = torch.clamp(alpha * L - torch.arange(L, dtype=x.dtype), 0.0, 1.0)
        w = 0.5 * (1 + torch.cos(math.pi * w + math.pi))
    for k in range(L):
        f = 2.0 ** k
        wk = 1.0 if w is None else w[k]
        g = g + wk * f * (u[..., 3 + 6 * k:6 + 6 * k] * torch.cos(f * x) - u[..., 6 + 6 * k:9 + 6 * k] * torch.sin(f * x))
    return g


def masks(P, cfg, x, inst, alpha=None):
    """ReLU masks of the basefield on points x (R,D,3) with per-ray instance code rows inst (R,32): list of D+1 bool tensors."""
    e = torch.cat([O.pos_embed(x, cfg["L_xyz"], alpha), O._per_frame(inst, x)], -1)
    h, out = e, []
    for i in range(cfg["D"]):
        if i == cfg.get("skip", 4):
            h = torch.cat([e, h], -1)
        z = h @ P[f"basefield.linear_{i+1}.0.weight"].T + P[f"basefield.linear_{i+1}.0.bias"]
        out.append(z > 0)
        h = z.clamp(min=0)
    z = h @ P["basefield.linear_final.0.weight"].T + P["basefield.linear_final.0.bias"]
    out.append(z > 0)
    return out


def sdf_gradient_autograd(P, cfg, x, inst, alpha=None, create_graph=True):
    """The reference's way: autograd.grad of the summed sdf w.r.t. the points (compute_gradient, torch_utils.py:4-28)."""
    with torch.enable_grad():
        xs = x.detach().requires_grad_(True)
        s = O.nerf_forward(P, cfg, xs, inst, None, get_density=False, alpha=alpha)
        (g,) = torch.autograd.grad(s.sum(), xs, create_graph=create_graph)
    return g


def a_pass(P, cfg, x, m, alpha=None):
    """Reverse chain with a unit cotangent on the sdf.  Returns g (R,D,3) and the list a[0..D] (a[i] = gradient of the
    pre-activation of basefield layer i; i = D is linear_final)."""
    Dn, skip, L = cfg["D"], cfg.get("skip", 4), cfg["L_xyz"]
    n_pe = 3 * (2 * L + 1)
    n_in = n_pe + 32
    a = [None] * (Dn + 1)
    a[Dn] = m[Dn] * P["sdf.weight"][0]
    g_h = a[Dn] @ P["basefield.linear_final.0.weight"]
    u = torch.zeros(x.shape[:-1] + (n_pe,), dtype=x.dtype)
    for i in reversed(range(Dn)):
        a[i] = g_h * m[i]
        g_h = a[i] @ P[f"basefield.linear_{i+1}.0.weight"]
        if i == skip:
            u = u + g_h[..., :n_pe]
            g_h = g_h[..., n_in:]
    u = u + g_h[..., :n_pe]
    return pe_jacobian_t_apply(x, L, u, alpha), a


def v_pass(P, cfg, x, m, gbar, alpha=None):
    """Forward chains A and B (module docstring).  Returns v0 (R,D,n_pe), vA[0..D], vB[0..D] (None before the skip layer)."""
    Dn, skip, L = cfg["D"], cfg.get("skip", 4), cfg["L_xyz"]
    n_pe = 3 * (2 * L + 1)
    n_in = n_pe + 32
    v0 = pe_jacobian_apply(x, L, gbar, alpha)
    vA, vB = [None] * (Dn + 1), [None] * (Dn + 1)
    hA, hB = v0, None
    for i in range(Dn + 1):
        W = P[f"basefield.linear_{i+1}.0.weight"] if i < Dn else P["basefield.linear_final.0.weight"]
        if i == 0:
            vA[i] = m[i] * (hA @ W[:, :n_pe].T)
        elif i == skip:
            vA[i] = m[i] * (hA @ W[:, n_in:].T)
            vB[i] = m[i] * (v0 @ W[:, :n_pe].T)
        else:
            vA[i] = m[i] * (hA @ W.T)
            if hB is not None:
                vB[i] = m[i] * (hB @ W.T)
        hA, hB = vA[i], vB[i]
    return v0, vA, vB


def weight_grads(P, cfg, a, v0, vA, vB):
    """dL/dW of every basefield layer and of sdf.weight from the three chains (sums over all points)."""
    Dn, skip, L = cfg["D"], cfg.get("skip", 4), cfg["L_xyz"]
    n_pe = 3 * (2 * L + 1)
    n_in = n_pe + 32
    grads = {}
    ein = lambda g, h: torch.einsum("...i,...j->ij", g, h)
    for i in range(Dn + 1):
        name = f"basefield.linear_{i+1}.0.weight" if i < Dn else "basefield.linear_final.0.weight"
        dW = torch.zeros_like(P[name])
        if i == 0:
            dW[:, :n_pe] = ein(a[i], v0)
        else:
            prev = vA[i - 1] if vB[i - 1] is None else vA[i - 1] + vB[i - 1]
            if i == skip:
                dW[:, :n_pe] = ein(a[i], v0)
                dW[:, n_in:] = ein(a[i], prev)
            else:
                dW = ein(a[i], prev)
        grads[name] = dW
    last = vA[Dn] if vB[Dn] is None else vA[Dn] + vB[Dn]
    grads["sdf.weight"] = last.reshape(-1, last.shape[-1]).sum(0, keepdim=True)
    return grads


def eikonal_hand(P, cfg, x, inst, gbar_fn, alpha=None):
    """g and the weight gradients of L = gbar_fn's loss: gbar_fn(g) -> dL/dg (same shape as g)."""
    m = masks(P, cfg, x, inst, alpha)
    g, a = a_pass(P, cfg, x, m, alpha)
    gbar = gbar_fn(g)
    v0, vA, vB = v_pass(P, cfg, x, m, gbar, alpha)
    return g, weight_grads(P, cfg, a, v0, vA, vB)
