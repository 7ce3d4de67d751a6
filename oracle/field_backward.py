"""Hand-derived backward of the whole training-mode query_field of a skinned fg field (with or without the dense warp of
a ComposedWarp; no eikonal), composed from oracle/skin_backward.py and oracle/nerf_backward.py plus the geometry / flow / cycle / visibility / feature /
Gaussian-density stages.  TEST INFRASTRUCTURE: the per-sample part is written out by hand in the order the CUDA backward
will run it; the per-frame part (tables <- articulations, folded bias rows <- codes) is left to autograd exactly as planned
for the product.  tests/test_field_backward_cpu.py checks every parameter and per-frame input gradient
against autograd through the reference-pinned oracle (lab4d_oracle.query_field) in fp64."""
import torch

import lab4d_oracle as O
import nerf_backward as NB
import skin_backward as SB


def _qrot_bwd(q, p, g):
    """r = vec(q (0,p) q*): returns (g_q, g_p) for the cotangent g of r (q need not be unit)."""
    zero = torch.zeros_like(g[..., :1])
    G, Pq = torch.cat([zero, g], -1), torch.cat([zero, p], -1)
    u = O.qmul(q, Pq)
    g_u = O.qmul(G, q)
    g_q = O.qmul(O.qconj(G), u) + O.qmul(g_u, O.qconj(Pq))
    return g_q, O.qmul(O.qconj(q), g_u)[..., 1:]


def _skin_tables(P, cfg, tab, M, which):
    """Tables of warp `which` (0 backward, 1 forward with the partner frame, 2 forward own) as differentiable functions
    of the per-frame inputs (the prologue's job)."""
    B = cfg["B"]
    t_art = (tab["t_articulation_qr"], tab["t_articulation_qd"])
    r_art = (tab["rest_articulation_qr"], tab["rest_articulation_qd"])
    if which == 1:
        t_art, r_art = tuple(O.flip_pair(a) for a in t_art), tuple(O.flip_pair(a) for a in r_art)
    Rp, tp, se3_r, se3_d = SB.tables_from_articulation(P, t_art, r_art, which == 0, cfg.get("symm_idx"))
    pre = "warp.skinning_model.delta_field."
    W1, b1 = P[pre + "linear_1.0.weight"], P[pre + "linear_1.0.bias"]
    te = tab["skin_t_embed"] if which == 0 else tab["skin_t_embed_mean"].expand(M, -1)
    b1row = b1 + te @ W1[:, 3 * B:3 * B + 128].T + tab["inst_skin"] @ W1[:, 3 * B + 128:].T
    return dict(Rp=Rp, tp=tp, se3_r=se3_r, se3_d=se3_d, W1x=W1[:, :3 * B], b1row=b1row, W2=P[pre + "linear_2.0.weight"],
                b2=P[pre + "linear_2.0.bias"], W3=P[pre + "linear_final.weight"], b3=P[pre + "linear_final.bias"])


def _dense_fwd(P, x, t_embed, inst, backward):
    """DenseWarp.forward (warping.py:143-170): x + 0.1 * CondMLP([PE6(x), t, inst]); returns (x', saved)."""
    prefix = "warp.post_warp.backward_map." if backward else "warp.post_warp.forward_map."
    m, sv = NB.mlp_forward_saved(P, prefix, NB.pe_forward(x, 6), torch.cat([t_embed, inst], -1), 2, final_act=False)
    return x + 0.1 * m, dict(sv=sv, x=x, prefix=prefix)


def _dense_bwd(P, saved, g_out, acc):
    """Returns (g_x, g_t_embed (M,128), g_inst (M,32)); parameter gradients go to acc()."""
    g_e, g_code, gp = NB.mlp_backward(P, saved["prefix"], saved["sv"], 0.1 * g_out, 2, final_act=False)
    acc(gp)
    return g_out + NB.pe_backward(saved["x"], 6, g_e), g_code[:, :128], g_code[:, 128:]


def forward_saved(P, cfg, rays, tab, D):
    """Forward in the kernel's formulation; returns the outputs (M,S,c) with S = N*D and what the backward needs."""
    hxy, Kinv, near_far = rays["hxy"], rays["Kinv"], rays["near_far"]
    M, N = hxy.shape[:2]
    S = N * D
    d = torch.einsum("mni,mji->mnj", hxy, Kinv)
    z = torch.linspace(0, 1, D, dtype=hxy.dtype)
    depth = near_far[:, 0:1] * (1 - z) + near_far[:, 1:2] * z                      # (M,D)
    xyz_cam = (d[:, :, None, :] * depth[:, None, :, None]).reshape(M, S, 3)
    q, t = tab["field2cam_q"], tab["field2cam_t"]
    qi = O.qconj(q)
    ti = O.qrot(qi, -t)
    xyz_t = O.qrot(qi[:, None].expand(M, S, 4), xyz_cam) + ti[:, None]
    dir_cam = (d / d.norm(dim=-1, keepdim=True))[:, :, None, :].expand(M, N, D, 3).reshape(M, S, 3)
    dir_f = O.qrot(qi[:, None].expand(M, S, 4), dir_cam)
    if cfg["motion"] == "rigid":
        return _forward_rigid(P, cfg, rays, tab, dict(M=M, N=N, D=D, S=S, d=d, depth=depth, xyz_cam=xyz_cam, qi=qi, xyz_t=xyz_t,
                                                     dir_cam=dir_cam, dir_f=dir_f))
    T = [_skin_tables(P, cfg, tab, M, w) for w in range(3)]
    dense = bool(cfg.get("dense", False))
    xs, ent0, dsk0, sv0 = SB.skin_forward_tables(xyz_t, **T[0])
    dn = [None, None, None]
    if dense:  # ComposedWarp (warping.py:445-483): backward = skin then dense, forward = dense then skin
        xyz, dn[0] = _dense_fwd(P, xs, tab["dense_t_embed"], tab["inst_dense_bwd"], True)
        x1, dn[1] = _dense_fwd(P, xyz, O.flip_pair(tab["dense_t_embed"]), tab["inst_dense_fwd"], False)
        x2, dn[2] = _dense_fwd(P, xyz, tab["dense_t_embed"], tab["inst_dense_fwd"], False)
    else:
        xyz = x1 = x2 = xs
    x_next, _, _, sv1 = SB.skin_forward_tables(x1, **T[1])
    x_cyc, ent2, dsk2, sv2 = SB.skin_forward_tables(x2, **T[2])
    qn, tn, Kn = O.flip_pair(q), O.flip_pair(t), O.flip_pair(Kinv)
    xc = O.qrot(qn[:, None].expand(M, S, 4), x_next) + tn[:, None]
    Kmat = O.kmat_from_kinv(Kn)
    hn = torch.einsum("mij,msj->msi", Kmat, xc)
    flow = hn[..., :2] / (hn[..., 2:] + 1e-6) - hxy[:, :, None, :2].expand(M, N, D, 2).reshape(M, S, 2)
    diff = x_cyc - xyz_t
    cyc = diff.norm(dim=-1, keepdim=True)
    rgb, density, sdf, snerf = NB.nerf_forward_saved(P, cfg, xyz, tab["inst_base"], tab["inst_color"], None, tab.get("appr_code"))
    vis, svis = NB.mlp_forward_saved(P, "vis_mlp.basefield.", NB.pe_forward(xyz, 10), tab["inst_vis"], 2, final_act=False)
    fraw, sfeat = NB.mlp_forward_saved(P, "feature_field.", NB.pe_forward(xyz, 6), xyz[:, 0, :0], 5, final_act=False)
    fnorm = fraw.norm(dim=-1, keepdim=True)
    _, ctr = O.dq_to_qt((tab["rest_articulation_qr"][:1], tab["rest_articulation_qd"][:1]))   # (1,B,3) rest bone centres
    d2 = (xyz[..., None, :] - ctr[:, None]).pow(2).sum(-1)                                     # (M,S,B)
    dmin, amin = d2.min(-1)
    wib = P["warp.logibeta"].exp()
    gd = torch.exp(-0.5 * dmin / 0.01 ** 2)[..., None] * wib
    out = dict(rgb=rgb, density=density, vis=vis, feature=fraw / fnorm, flow=flow, cyc_dist=cyc, xyz=xyz, xyz_cam=xyz_cam,
               delta_skin=0.5 * (dsk0 + dsk2)[..., None], skin_entropy=0.5 * (ent0 + ent2)[..., None], gauss_density=gd)
    saved = dict(dn=dn, x1=x1, x2=x2, M=M, N=N, D=D, S=S, d=d, depth=depth, xyz_cam=xyz_cam, qi=qi, xyz_t=xyz_t, T=T, sv=(sv0, sv1, sv2), xyz=xyz,
                 x_next=x_next, x_cyc=x_cyc, qn=qn, xc=xc, Kmat=Kmat, hn=hn, diff=diff, cyc=cyc, snerf=snerf, svis=svis, sfeat=sfeat,
                 fraw=fraw, fnorm=fnorm, ctr=ctr, amin=amin, gd=gd, wib=wib)
    return out, saved


def _forward_rigid(P, cfg, rays, tab, v):
    """Rigid field (IdentityWarp): canonical point = time-t point; flow through the partner camera only."""
    M, S, N, D = v["M"], v["S"], v["N"], v["D"]
    xyz = v["xyz_t"]
    q, t = tab["field2cam_q"], tab["field2cam_t"]
    qn, tn, Kn = O.flip_pair(q), O.flip_pair(t), O.flip_pair(rays["Kinv"])
    xc = O.qrot(qn[:, None].expand(M, S, 4), xyz) + tn[:, None]
    Kmat = O.kmat_from_kinv(Kn)
    hn = torch.einsum("mij,msj->msi", Kmat, xc)
    flow = hn[..., :2] / (hn[..., 2:] + 1e-6) - rays["hxy"][:, :, None, :2].expand(M, N, D, 2).reshape(M, S, 2)
    dirs = v["dir_f"] if cfg["L_dir"] == 0 else None
    rgb, density, sdf, snerf = NB.nerf_forward_saved(P, cfg, xyz, tab["inst_base"], tab["inst_color"], dirs, tab.get("appr_code"))
    vis, svis = NB.mlp_forward_saved(P, "vis_mlp.basefield.", NB.pe_forward(xyz, 10), tab["inst_vis"], 2, final_act=False)
    out = dict(rgb=rgb, density=density, vis=vis, flow=flow, xyz=xyz, xyz_cam=v["xyz_cam"])
    v.update(rigid=True, xyz=xyz, qn=qn, xc=xc, Kmat=Kmat, hn=hn, snerf=snerf, svis=svis)
    return out, v


def _backward_rigid(P, cfg, rays, tab, v, g):
    M, S, D, N = v["M"], v["S"], v["D"], v["N"]
    grads = {}
    xyz = v["xyz"]
    g_xyz = g["xyz"].clone()
    g_e, g_inst_vis, gp = NB.mlp_backward(P, "vis_mlp.basefield.", v["svis"], g["vis"], 2, final_act=False)
    grads.update(gp)
    g_xyz = g_xyz + NB.pe_backward(xyz, 10, g_e)
    gin, gp = NB.nerf_backward(P, cfg, xyz, v["snerf"], g["rgb"], g["density"])
    grads.update(gp)
    g_xyz = g_xyz + gin["x"]
    hz = v["hn"][..., 2:] + 1e-6
    g_h = torch.cat([g["flow"] / hz, -(g["flow"] * v["hn"][..., :2]).sum(-1, keepdim=True) / hz.pow(2)], -1)
    g_Kmat = torch.einsum("msi,msj->mij", g_h, v["xc"])
    g_xc = torch.einsum("mij,msi->msj", v["Kmat"], g_h)
    g_qn_s, g_x = _qrot_bwd(v["qn"][:, None].expand(M, S, 4), xyz, g_xc)
    g_xyz = g_xyz + g_x
    # camera -> field for the point and, when the field sees it, the view direction
    qi_s = v["qi"][:, None].expand(M, S, 4)
    g_qi_s, g_xyz_cam = _qrot_bwd(qi_s, v["xyz_cam"], g_xyz)
    g_d = ((g_xyz_cam + g["xyz_cam"]).reshape(M, N, D, 3) * v["depth"][:, None, :, None]).sum(2)
    if gin["dir"] is not None:
        g_q2, g_dir_cam = _qrot_bwd(qi_s, v["dir_cam"], gin["dir"])
        g_qi_s = g_qi_s + g_q2
        dvec = v["d"]                                   # dir_cam = d / |d|
        nrm = dvec.norm(dim=-1, keepdim=True)
        u = dvec / nrm
        gdc = g_dir_cam.reshape(M, N, D, 3).sum(2)
        g_d = g_d + (gdc - u * (gdc * u).sum(-1, keepdim=True)) / nrm
    tables = dict(g_qi=g_qi_s.sum(1), g_ti=g_xyz.sum(1), g_qn=g_qn_s.sum(1), g_tn=g_xc.sum(1), g_Kmat=g_Kmat,
                  g_Kinv=torch.einsum("mnj,mni->mji", g_d, rays["hxy"]), g_inst_vis=g_inst_vis, g_inst_base=gin["inst_base"],
                  g_inst_color=gin["inst_color"], g_appr=gin["appr"])
    return grads, tables


def backward(P, cfg, rays, tab, saved, g):
    """g: cotangent per output key (M,S,c).  Returns (param grads by name, table grads of the kernel-level tables)."""
    v = saved
    if v.get("rigid"):
        return _backward_rigid(P, cfg, rays, tab, v, g)
    M, S, D, N = v["M"], v["S"], v["D"], v["N"]
    grads = {}

    def acc(d):
        for k, val in d.items():
            grads[k] = grads[k] + val if k in grads else val

    xyz = v["xyz"]
    # ---- Gaussian bone density: gd = exp(-dmin / (2 * 0.01^2)) * exp(warp.logibeta)
    g_xyz = g["xyz"].clone()
    ggd = g["gauss_density"] * v["gd"]
    c_sel = torch.gather(v["ctr"].expand(M, -1, -1), 1, v["amin"][..., None].expand(M, S, 3))
    g_xyz = g_xyz + ggd * (-(xyz - c_sel) / 0.01 ** 2)
    grads["warp.logibeta"] = ggd.sum().reshape(1)
    g_ctr = torch.zeros_like(v["ctr"]).expand(M, -1, -1).clone()
    g_ctr.scatter_add_(1, v["amin"][..., None].expand(M, S, 3), ggd * ((xyz - c_sel) / 0.01 ** 2))
    g_ctr = g_ctr.sum(0, keepdim=True)
    # ---- feature field: out = f / |f|
    o = v["fraw"] / v["fnorm"]
    g_f = (g["feature"] - o * (g["feature"] * o).sum(-1, keepdim=True)) / v["fnorm"]
    g_e, _, gp = NB.mlp_backward(P, "feature_field.", v["sfeat"], g_f, 5, final_act=False)
    acc(gp)
    g_xyz = g_xyz + NB.pe_backward(xyz, 6, g_e)
    # ---- visibility MLP
    g_e, g_inst_vis, gp = NB.mlp_backward(P, "vis_mlp.basefield.", v["svis"], g["vis"], 2, final_act=False)
    acc(gp)
    g_xyz = g_xyz + NB.pe_backward(xyz, 10, g_e)
    # ---- density + colour
    gin, gp = NB.nerf_backward(P, cfg, xyz, v["snerf"], g["rgb"], g["density"])
    acc(gp)
    g_xyz = g_xyz + gin["x"]
    # ---- cycle: |x_cyc - xyz_t|
    g_diff = g["cyc_dist"] * v["diff"] / v["cyc"]
    g_xyz_t = -g_diff
    T, (sv0, sv1, sv2) = v["T"], v["sv"]
    zero1 = torch.zeros(M, S, dtype=xyz.dtype)
    dn = v["dn"]
    g_dense = {"dense_t_embed": torch.zeros(M, 128, dtype=xyz.dtype), "inst_dense_fwd": torch.zeros(M, 32, dtype=xyz.dtype),
               "inst_dense_bwd": torch.zeros(M, 32, dtype=xyz.dtype)}
    b2 = SB.skin_backward_tables(v["x2"], **T[2], saved=sv2, g_xo=g_diff, g_ent=0.5 * g["skin_entropy"][..., 0], g_dsk=0.5 * g["delta_skin"][..., 0])
    if dn[2] is not None:
        gx, gt, gi = _dense_bwd(P, dn[2], b2["x"], acc)
        g_xyz = g_xyz + gx
        g_dense["dense_t_embed"] += gt
        g_dense["inst_dense_fwd"] += gi
    else:
        g_xyz = g_xyz + b2["x"]
    # ---- flow: h = K xc, flow = h_xy / (h_z + 1e-6) - hxy;  xc = R(qn) x_next + tn
    hz = v["hn"][..., 2:] + 1e-6
    g_h = torch.cat([g["flow"] / hz, -(g["flow"] * v["hn"][..., :2]).sum(-1, keepdim=True) / hz.pow(2)], -1)
    g_Kmat = torch.einsum("msi,msj->mij", g_h, v["xc"])
    g_xc = torch.einsum("mij,msi->msj", v["Kmat"], g_h)
    g_tn = g_xc.sum(1)
    g_qn_s, g_xnext = _qrot_bwd(v["qn"][:, None].expand(M, S, 4), v["x_next"], g_xc)
    g_qn = g_qn_s.sum(1)
    b1 = SB.skin_backward_tables(v["x1"], **T[1], saved=sv1, g_xo=g_xnext, g_ent=zero1, g_dsk=zero1)
    if dn[1] is not None:
        gx, gt, gi = _dense_bwd(P, dn[1], b1["x"], acc)
        g_xyz = g_xyz + gx
        g_dense["dense_t_embed"] += O.flip_pair(gt)  # the partner frame's time code
        g_dense["inst_dense_fwd"] += gi
    else:
        g_xyz = g_xyz + b1["x"]
    # ---- backward warp
    if dn[0] is not None:
        g_xyz, gt, gi = _dense_bwd(P, dn[0], g_xyz, acc)
        g_dense["dense_t_embed"] += gt
        g_dense["inst_dense_bwd"] += gi
    b0 = SB.skin_backward_tables(v["xyz_t"], **T[0], saved=sv0, g_xo=g_xyz, g_ent=0.5 * g["skin_entropy"][..., 0], g_dsk=0.5 * g["delta_skin"][..., 0])
    g_xyz_t = g_xyz_t + b0["x"]
    # ---- camera -> field: xyz_t = R(qi) xyz_cam + ti
    g_ti = g_xyz_t.sum(1)
    g_qi_s, g_xyz_cam = _qrot_bwd(v["qi"][:, None].expand(M, S, 4), v["xyz_cam"], g_xyz_t)
    g_qi = g_qi_s.sum(1)
    g_xyz_cam = g_xyz_cam + g["xyz_cam"]
    # ---- sample placement: xyz_cam = (Kinv hxy) * depth
    g_d = (g_xyz_cam.reshape(M, N, D, 3) * v["depth"][:, None, :, None]).sum(2)
    g_Kinv = torch.einsum("mnj,mni->mji", g_d, rays["hxy"])
    tables = dict(g_dense=g_dense, skin=(b0, b1, b2), g_ctr=g_ctr, g_qi=g_qi, g_ti=g_ti, g_qn=g_qn, g_tn=g_tn, g_Kmat=g_Kmat, g_Kinv=g_Kinv,
                  g_inst_vis=g_inst_vis, g_inst_base=gin["inst_base"], g_inst_color=gin["inst_color"], g_appr=gin["appr"])
    return grads, tables
