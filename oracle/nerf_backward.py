"""Hand-derived backward of NeRF.forward (embedding -> basefield -> sdf / VolSDF density, colorfield, rgb head) — the dense
part of the field kernel's backward.  TEST INFRASTRUCTURE: only tests/ import this.

Everything is written as explicit matrix products and masks (no autograd), in the order a dgrad / wgrad tile program
would run them; tests/test_nerf_backward_cpu.py checks it against autograd through lab4d_oracle.nerf_forward in fp64.
Shapes: x (M,S,3) canonical points, dir (M,S,3) or None, codes per frame (M,C)."""
import math

import torch

import lab4d_oracle as O


def pe_forward(x, L):
    return O.pos_embed(x, L)


def pe_backward(x, L, g_e):
    """e = [x, sin(2^k x), cos(2^k x)]_k  ->  g_x = g_e[:3] + sum_k 2^k (g_sin_k cos(2^k x) - g_cos_k sin(2^k x))."""
    if L <= 0:
        return g_e[..., :3] if L == 0 else torch.zeros_like(x)
    g_x = g_e[..., :3].clone()
    for k in range(L):
        f = 2.0 ** k
        gs, gc = g_e[..., 3 + 6 * k:6 + 6 * k], g_e[..., 6 + 6 * k:9 + 6 * k]
        g_x = g_x + f * (gs * torch.cos(f * x) - gc * torch.sin(f * x))
    return g_x


def mlp_forward_saved(P, prefix, e, code, n_layers, skip=4, final_act=True):
    """BaseMLP / CondMLP forward on [e, code]; returns the output and what the backward needs (inputs of every layer and
    the pre-activation signs)."""
    x0 = torch.cat([e, code[:, None].expand(e.shape[:-1] + (code.shape[-1],))], -1)
    h, saved = x0, []
    for i in range(n_layers):
        if i == skip:
            h = torch.cat([x0, h], -1)
        z = h @ P[f"{prefix}linear_{i+1}.0.weight"].T + P[f"{prefix}linear_{i+1}.0.bias"]
        saved.append((h, z > 0))
        h = z.clamp(min=0)
    fin = f"{prefix}linear_final.0." if final_act else f"{prefix}linear_final."
    z = h @ P[fin + "weight"].T + P[fin + "bias"]
    saved.append((h, (z > 0) if final_act else None))
    return (z.clamp(min=0) if final_act else z), dict(x0=x0, layers=saved, n_e=e.shape[-1])


def mlp_backward(P, prefix, saved, g_out, n_layers, skip=4, final_act=True):
    """Returns (g_e, g_code (per frame), {param name: grad}).  dgrad: G_{l-1} = (G_l * mask_l) W_l;
    wgrad: dW_l = (G_l * mask_l)^T A_{l-1}; the skip layer sends the first columns of its dgrad back to the input."""
    grads = {}
    fin = f"{prefix}linear_final.0." if final_act else f"{prefix}linear_final."
    h, mask = saved["layers"][n_layers]
    gz = g_out * mask if mask is not None else g_out
    grads[fin + "weight"] = torch.einsum("msi,msj->ij", gz, h)
    grads[fin + "bias"] = gz.sum((0, 1))
    g_h = gz @ P[fin + "weight"]
    n_in = saved["x0"].shape[-1]
    g_x0 = torch.zeros_like(saved["x0"])
    for i in reversed(range(n_layers)):
        h, mask = saved["layers"][i]
        gz = g_h * mask
        W = P[f"{prefix}linear_{i+1}.0.weight"]
        grads[f"{prefix}linear_{i+1}.0.weight"] = torch.einsum("msi,msj->ij", gz, h)
        grads[f"{prefix}linear_{i+1}.0.bias"] = gz.sum((0, 1))
        g_h = gz @ W
        if i == skip:
            g_x0 = g_x0 + g_h[..., :n_in]
            g_h = g_h[..., n_in:]
    g_x0 = g_x0 + g_h
    n_e = saved["n_e"]
    return g_x0[..., :n_e], g_x0[..., n_e:].sum(1), grads


def nerf_forward_saved(P, cfg, x, inst_base, inst_color, dirs=None, appr=None):
    """NeRF.forward (nnutils/nerf.py:167-215): returns rgb, density, sdf and the saved tensors."""
    Lb, Lc = cfg["L_xyz"], cfg["L_xyz"] + 2
    feat, sb = mlp_forward_saved(P, "basefield.", pe_forward(x, Lb), inst_base, cfg["D"])
    sdf = feat @ P["sdf.weight"].T + P["sdf.bias"]
    ibeta = P["logibeta"].exp()
    density = (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() * ibeta)) * ibeta
    cfeat, sc = mlp_forward_saved(P, "colorfield.", pe_forward(x, Lc), inst_color, 2)
    f2 = feat + cfeat
    de = O.pos_embed(dirs, cfg["L_dir"]) if dirs is not None else x[..., :0]
    if cfg["appr_channels"] > 0:
        de = torch.cat([de, appr[:, None].expand(de.shape[:-1] + (appr.shape[-1],))], -1)
    r_in = torch.cat([f2, de], -1)
    z0 = r_in @ P["rgb.0.weight"].T + P["rgb.0.bias"]
    h0 = z0.clamp(min=0)
    rgb = (h0 @ P["rgb.2.weight"].T + P["rgb.2.bias"]).sigmoid()
    return rgb, density, sdf, dict(sb=sb, sc=sc, feat=feat, sdf=sdf, ibeta=ibeta, r_in=r_in, z0=z0, h0=h0, rgb=rgb, n_dir=pe_n(cfg["L_dir"]))


def pe_n(L):
    return 0 if L < 0 else 3 * (2 * L + 1)


def nerf_backward(P, cfg, x, saved, g_rgb, g_density):
    """Gradients of L = <g_rgb, rgb> + <g_density, density> w.r.t. x, the per-frame codes and every parameter."""
    v, grads = saved, {}
    W = cfg["W"]
    # rgb = sigmoid(W2 relu(W0 [feat + cfeat, dir, appr] + b0) + b2)
    g_o = g_rgb * v["rgb"] * (1 - v["rgb"])
    grads["rgb.2.weight"] = torch.einsum("msi,msj->ij", g_o, v["h0"])
    grads["rgb.2.bias"] = g_o.sum((0, 1))
    g_z0 = (g_o @ P["rgb.2.weight"]) * (v["z0"] > 0)
    grads["rgb.0.weight"] = torch.einsum("msi,msj->ij", g_z0, v["r_in"])
    grads["rgb.0.bias"] = g_z0.sum((0, 1))
    g_rin = g_z0 @ P["rgb.0.weight"]
    g_f2 = g_rin[..., :W]
    g_appr = g_rin[..., W + v["n_dir"]:].sum(1) if cfg["appr_channels"] > 0 else None
    g_dir = g_rin[..., W:W + 3] if cfg["L_dir"] == 0 else None  # raw view direction (bg fields)
    # colorfield
    g_ec, g_inst_color, gc = mlp_backward(P, "colorfield.", v["sc"], g_f2, 2)
    grads.update(gc)
    # density = (0.5 + 0.5 sign(s) expm1(-|s| ibeta)) ibeta  ->  d/ds = -0.5 ibeta^2 exp(-|s| ibeta)
    s, ibeta = v["sdf"], v["ibeta"]
    ex = torch.exp(-s.abs() * ibeta)
    g_sdf = g_density * (-0.5 * ibeta * ibeta * ex)
    # d density / d logibeta = ibeta * d density / d ibeta
    dd_dib = (0.5 + 0.5 * s.sign() * torch.expm1(-s.abs() * ibeta)) + ibeta * (0.5 * s.sign() * ex * (-s.abs()))
    grads["logibeta"] = (g_density * dd_dib).sum().reshape(1) * ibeta
    grads["sdf.weight"] = torch.einsum("msi,msj->ij", g_sdf, v["feat"])
    grads["sdf.bias"] = g_sdf.sum((0, 1))
    g_feat = g_f2 + g_sdf @ P["sdf.weight"]
    g_eb, g_inst_base, gb = mlp_backward(P, "basefield.", v["sb"], g_feat, cfg["D"])
    grads.update(gb)
    g_x = pe_backward(x, cfg["L_xyz"], g_eb) + pe_backward(x, cfg["L_xyz"] + 2, g_ec)
    return dict(x=g_x, inst_base=g_inst_base, inst_color=g_inst_color, appr=g_appr, dir=g_dir), grads
