"""Hand-derived backward of the blend-skinning warp in the TABLE formulation the CUDA kernel uses
(groundwork for the field kernel's backward).  TEST INFRASTRUCTURE: only tests/ import this.

Forward (per sample x of frame f, B bones; all tables are per frame):
    xb_b   = Rp_b x + tp_b                      Rp/tp: inverse bone transform pre-scaled by 1/gauss (prologue.cu write_binv)
    d2_b   = |xb_b|^2
    h1     = relu(W1x vec(xb) + b1row_f)        b1row: delta_field.linear_1 bias with the time/instance codes folded in
    h2     = relu(W2 h1 + b2);  z = W3 h2 + b3;  dlt_b = 0.1 relu(z_b)
    logit  = -(d2 + dlt);  w = softmax(logit)
    a      = argmax w;  s_b = +1 if <qr_a, qr_b> > 0 else -1         (geom_utils.py:67-69)
    qhr    = sum_b w_b s_b qr_b;  qhd = sum_b w_b s_b qd_b;  n = |qhr|;  Qr = qhr / n;  Qd = qhd / n
    x'     = vec(Qr (0,x) Qr*) + 2 vec(Qd Qr*)
    ent    = logsumexp(logit) - max(logit);   dsk = mean_b dlt_b^2
`skin_forward_tables` is that forward in differentiable torch; `skin_backward_tables` is the same chain rule written out by
hand (quaternion products: for r = a (x) b, g_a = g_r (x) b*, g_b = a* (x) g_r).  tests/test_skin_backward_cpu.py checks the
two against each other in fp64 and the forward against the reference-pinned oracle (lab4d_oracle.skinning_warp).
"""
import torch

import lab4d_oracle as O


def tables_from_articulation(P, t_art, rest_art, backward, symm_idx=None):
    """Per-frame tables of one warp: (Rp (M,B,3,3), tp (M,B,3), se3_r (M,B,4), se3_d (M,B,4)) like prologue.cu."""
    if backward:
        se3 = O.dq_mul(rest_art, O.dq_inv(t_art))
        art = t_art
    else:
        se3 = O.dq_mul(t_art, O.dq_inv(rest_art))
        art = rest_art
    inv_gauss = 1.0 / O.gauss_scale(P, symm_idx)  # (B,3)
    q, t = O.dq_to_qt(O.dq_inv(art))              # x_b = R(q) x + t
    eye = torch.eye(3, dtype=q.dtype).expand(q.shape[:-1] + (3, 3))
    R = torch.stack([O.qrot(q, eye[..., :, i]) for i in range(3)], -1)  # columns = rotated basis vectors
    return R * inv_gauss[None, :, :, None], t * inv_gauss[None], se3[0], se3[1]


def _sign(qr, a_idx):
    qa = torch.gather(qr, -2, a_idx[..., None, None].expand(a_idx.shape + (1, 4)))
    return torch.where((qa * qr).sum(-1) > 0, 1.0, -1.0).to(qr.dtype)


def skin_forward_tables(x, Rp, tp, se3_r, se3_d, W1x, b1row, W2, b2, W3, b3):
    """x (M,S,3); tables (M,B,...); W1x (64,3B); b1row (M,64).  Returns x' (M,S,3), ent (M,S), dsk (M,S), saved."""
    M, S, _ = x.shape
    B = Rp.shape[1]
    xb = torch.einsum("mbij,msj->msbi", Rp, x) + tp[:, None]          # (M,S,B,3)
    d2 = xb.pow(2).sum(-1)
    z1 = xb.reshape(M, S, 3 * B) @ W1x.T + b1row[:, None]
    h1 = z1.clamp(min=0)
    z2 = h1 @ W2.T + b2
    h2 = z2.clamp(min=0)
    z = h2 @ W3.T + b3
    dlt = 0.1 * z.clamp(min=0)
    logit = -(d2 + dlt)
    w = logit.softmax(-1)
    a_idx = w.argmax(-1)
    qr = se3_r[:, None].expand(M, S, B, 4)
    qd = se3_d[:, None].expand(M, S, B, 4)
    s = _sign(qr, a_idx)
    ws = w * s
    qhr = (ws[..., None] * qr).sum(-2)
    qhd = (ws[..., None] * qd).sum(-2)
    n = qhr.norm(dim=-1, keepdim=True)
    Qr, Qd = qhr / n, qhd / n
    xo = O.qrot(Qr, x) + 2 * O.qmul(Qd, O.qconj(Qr))[..., 1:]
    ent = torch.logsumexp(logit, -1) - logit.max(-1)[0]
    dsk = dlt.pow(2).mean(-1)
    saved = dict(xb=xb, z1=z1, h1=h1, z2=z2, h2=h2, z=z, dlt=dlt, logit=logit, w=w, s=s, qhr=qhr, qhd=qhd, n=n, Qr=Qr, Qd=Qd)
    return xo, ent, dsk, saved


def skin_backward_tables(x, Rp, tp, se3_r, se3_d, W1x, b1row, W2, b2, W3, b3, saved, g_xo, g_ent, g_dsk):
    """Gradients of L = <g_xo, x'> + <g_ent, ent> + <g_dsk, dsk> w.r.t. every input, by hand."""
    M, S, _ = x.shape
    B = Rp.shape[1]
    v = saved
    zero = torch.zeros_like(g_xo[..., :1])
    G = torch.cat([zero, g_xo], -1)                    # gradient w.r.t. the quaternion whose vector part is x'
    Pq = torch.cat([zero, x], -1)
    Qr, Qd, n = v["Qr"], v["Qd"], v["n"]
    # x' = vec((Qr P) Qr*) + 2 vec(Qd Qr*)
    u = O.qmul(Qr, Pq)
    g_u = O.qmul(G, Qr)                                # r = u Qr*  ->  g_u = G (Qr*)* = G Qr
    g_Qr = O.qmul(O.qconj(G), u)                       # through the conjugate: conj(u* G) = G* u
    g_Qr = g_Qr + O.qmul(g_u, O.qconj(Pq))             # u = Qr P
    g_x = O.qmul(O.qconj(Qr), g_u)[..., 1:]
    g_Qd = 2 * O.qmul(G, Qr)                           # 2 Qd Qr*
    g_Qr = g_Qr + 2 * O.qmul(O.qconj(G), Qd)
    # Qr = qhr / n, Qd = qhd / n, n = |qhr|
    g_qhd = g_Qd / n
    radial = (g_Qr * v["qhr"]).sum(-1, keepdim=True) + (g_Qd * v["qhd"]).sum(-1, keepdim=True)
    g_qhr = g_Qr / n - v["qhr"] * radial / n.pow(3)
    # qhr = sum_b w_b s_b qr_b
    qr = se3_r[:, None].expand(M, S, B, 4)
    qd = se3_d[:, None].expand(M, S, B, 4)
    w, s = v["w"], v["s"]
    g_w = s * ((g_qhr[..., None, :] * qr).sum(-1) + (g_qhd[..., None, :] * qd).sum(-1))
    g_se3_r = ((w * s)[..., None] * g_qhr[..., None, :]).sum(1)       # (M,B,4): reduction over the frame's samples
    g_se3_d = ((w * s)[..., None] * g_qhd[..., None, :]).sum(1)
    # softmax + entropy (lse - max)
    g_logit = w * (g_w - (w * g_w).sum(-1, keepdim=True))
    onehot = torch.nn.functional.one_hot(v["logit"].argmax(-1), B).to(w.dtype)
    g_logit = g_logit + g_ent[..., None] * (w - onehot)
    # logit = -(d2 + dlt); dsk = mean dlt^2
    g_d2 = -g_logit
    g_dlt = -g_logit + g_dsk[..., None] * 2 * v["dlt"] / B
    g_z = 0.1 * (v["z"] > 0).to(w.dtype) * g_dlt
    # delta MLP
    g_W3 = torch.einsum("msb,msh->bh", g_z, v["h2"])
    g_b3 = g_z.sum((0, 1))
    g_z2 = (g_z @ W3) * (v["z2"] > 0).to(w.dtype)
    g_W2 = torch.einsum("msi,msj->ij", g_z2, v["h1"])
    g_b2 = g_z2.sum((0, 1))
    g_z1 = (g_z2 @ W2) * (v["z1"] > 0).to(w.dtype)
    xb = v["xb"]
    g_W1x = torch.einsum("msi,msj->ij", g_z1, xb.reshape(M, S, 3 * B))
    g_b1row = g_z1.sum(1)                                              # per frame
    g_xb = (g_z1 @ W1x).reshape(M, S, B, 3) + 2 * xb * g_d2[..., None]
    # xb_b = Rp_b x + tp_b
    g_x = g_x + torch.einsum("mbij,msbi->msj", Rp, g_xb)
    g_Rp = torch.einsum("msbi,msj->mbij", g_xb, x)
    g_tp = g_xb.sum(1)
    return dict(x=g_x, Rp=g_Rp, tp=g_tp, se3_r=g_se3_r, se3_d=g_se3_d, W1x=g_W1x, b1row=g_b1row, W2=g_W2, b2=g_b2, W3=g_W3, b3=g_b3)
