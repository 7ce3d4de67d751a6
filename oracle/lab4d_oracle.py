"""CPU restatement of Lab4D's per-ray-sample renderer (the hot path of SURVEY.md §8a).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this file; the product (lab4d_b200/) never does and fails
loudly when its CUDA library is missing.

Parity status: PINNED against the live reference.  oracle/gen_golden.py imports the unmodified
reference from /root/reference (CPU, oracle/ref_shims) and writes tests/golden/*.npz;
tests/test_oracle_golden.py checks every function below against those vectors.  The reference's
own test-suite pins only PosEmbedding (lab4d/tests/test_ops.py:64-133); that case is restated in
tests/test_oracle_golden.py::test_pos_embedding_annealing.

Written as plain functional torch (works in fp32 and fp64) from the formulas, not from the
reference's module code; each function cites the reference lines it restates.
Parameter names are the reference's state_dict keys.
"""
import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# quaternion helpers (lab4d/utils/quat_transform.py:62-81,255-287,337-344,431-465;
# dqtorch kernel semantics third_party/quaternion/src/quaternion.cu:46-62)
# --------------------------------------------------------------------------------------


def qmul(a, b):
    """Hamilton product, real part first; 3-vectors are pure quaternions."""
    if a.shape[-1] == 3:
        a = torch.cat([torch.zeros_like(a[..., :1]), a], -1)
    if b.shape[-1] == 3:
        b = torch.cat([torch.zeros_like(b[..., :1]), b], -1)
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack(
        [
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by - ax * bz + ay * bw + az * bx,
            aw * bz + ax * by - ay * bx + az * bw,
        ],
        -1,
    )


def qconj(q):
    return torch.cat([q[..., :1], -q[..., 1:]], -1)


def qrot(q, p):
    """quaternion_apply: (q (0,p) q*)_{xyz} (quat_transform.py:255-272)."""
    return qmul(qmul(q, p), qconj(q))[..., 1:]


def qt_inverse(q, t):
    """quaternion_translation_inverse (quat_transform.py:282-287)."""
    qi = qconj(q)
    return qi, qrot(qi, -t)


def dq_mul(a, b):
    """dual_quaternion_mul (quat_transform.py:431-438)."""
    return qmul(a[0], b[0]), qmul(a[0], b[1]) + qmul(a[1], b[0])


def dq_inv(a):
    """dual_quaternion_inverse = q-conjugate of both parts (quat_transform.py:441-465)."""
    return qconj(a[0]), qconj(a[1])


def dq_to_qt(dq):
    """dual_quaternion_to_quaternion_translation (quat_transform.py:337-344)."""
    return dq[0], 2 * qmul(dq[1], qconj(dq[0]))[..., 1:]


def dq_apply(dq, p):
    q, t = dq_to_qt(dq)
    return qrot(q, p) + t


# --------------------------------------------------------------------------------------
# embeddings and MLP blocks
# --------------------------------------------------------------------------------------


def pos_embed(x, L, alpha=None):
    """PosEmbedding.forward (nnutils/embedding.py:69-125): [x, sin(2^k x)(C), cos(2^k x)(C)]_k
    with the optional coarse-to-fine window.  L=-1 -> zero channels, L=0 -> identity."""
    if L == -1:
        return x[..., :0]
    if L == 0:
        return x
    freqs = 2.0 ** torch.arange(L, dtype=x.dtype, device=x.device)
    arg = x[..., None, :] * freqs[:, None]  # (..., L, C)
    bands = torch.stack([torch.sin(arg), torch.cos(arg)], -2)  # (..., L, 2, C)
    if alpha is not None:
        w = torch.clamp(alpha * L - torch.arange(L, dtype=x.dtype, device=x.device), 0.0, 1.0)
        w = 0.5 * (1 + torch.cos(math.pi * w + math.pi))
        bands = bands * w[:, None, None]
    return torch.cat([x, bands.reshape(x.shape[:-1] + (-1,))], -1)


def mlp(P, prefix, x, n_layers, skips=(4,), final_act=False, final_seq=None):
    """BaseMLP.forward (nnutils/base.py:65-78): D x (Linear+ReLU) with skip-concat [x, h] before
    layer index i in skips, then linear_final (+ReLU when final_act).  `final_seq`: linear_final is
    nn.Sequential (key 'linear_final.0.*') iff the module was built with final_act=True."""
    if final_seq is None:
        final_seq = final_act
    h = x
    for i in range(n_layers):
        if i in skips:
            h = torch.cat([x, h], -1)
        h = F.relu(F.linear(h, P[f"{prefix}linear_{i+1}.0.weight"], P[f"{prefix}linear_{i+1}.0.bias"]))
    fin = f"{prefix}linear_final.0." if final_seq else f"{prefix}linear_final."
    h = F.linear(h, P[fin + "weight"], P[fin + "bias"])
    return F.relu(h) if final_act else h


def _per_frame(code, ref):
    """Broadcast a per-frame row (M,C) over the sample dims of ref (M,...,c) (base.py:140-146)."""
    M = ref.shape[0]
    c = code.reshape((M,) + (1,) * (ref.dim() - 2) + (code.shape[-1],))
    return c.expand(ref.shape[:-1] + (code.shape[-1],))


# --------------------------------------------------------------------------------------
# stage 1: sample placement and camera -> field transform
# --------------------------------------------------------------------------------------


def sample_cam_rays(hxy, Kinv, near_far, D, depth=None):
    """utils/render_utils.py:8-56 with perturb=False; depth (M,N,D,1) overrides the uniform placement."""
    d = torch.einsum("mni,mji->mnj", hxy, Kinv)
    dn = d.norm(dim=-1)
    M, N = hxy.shape[:2]
    if depth is None:
        z = torch.linspace(0, 1, D, dtype=hxy.dtype, device=hxy.device)[None]
        depth = near_far[:, 0:1] * (1 - z) + near_far[:, 1:2] * z  # (M,D)
        depth = depth[:, None, :, None].expand(M, N, D, 1)
    else:
        D = depth.shape[2]
    xyz = d[:, :, None, :] * depth
    dl = depth[:, :, 1:] - depth[:, :, :-1]
    dl = torch.cat([dl, dl[:, :, -1:]], 2) * dn[..., None, None]
    dirn = (d / dn[..., None])[:, :, None, :].expand(M, N, D, 3)
    return xyz, dirn, dl, depth


def cam_to_field(xyz_cam, dir_cam, q, t):
    """NeRF.cam_to_field (nnutils/nerf.py:821-844)."""
    qi, ti = qt_inverse(q, t)
    qi_, ti_ = _per_frame(qi, xyz_cam), _per_frame(ti, xyz_cam)
    return qrot(qi_, xyz_cam) + ti_, qrot(qi_, dir_cam)


def field_to_cam(xyz, q, t):
    """NeRF.field_to_cam (nnutils/nerf.py:846-863)."""
    return qrot(_per_frame(q, xyz), xyz) + _per_frame(t, xyz)


# --------------------------------------------------------------------------------------
# stage 2: dual-quaternion blend skinning
# --------------------------------------------------------------------------------------


def gauss_scale(P, symm_idx=None):
    """SkinningField.get_gauss (nnutils/skinning.py:141-153)."""
    lg = P["warp.skinning_model.log_gauss"]
    if symm_idx is not None:
        lg = (lg[symm_idx] + lg) / 2
    return lg.exp()


def bone_coords(xyz, art, gauss):
    """get_bone_coords + Gaussian scaling (utils/transforms.py:9-25, skinning.py:122-139).
    xyz (M,...,3), art ((M,B,4),(M,B,4)) -> (M,...,B,3)."""
    M, B = art[0].shape[:2]
    nd = xyz.dim() - 2
    shp = (M,) + (1,) * nd + (B, 4)
    inv = (qconj(art[0]).reshape(shp), qconj(art[1]).reshape(shp))
    x = xyz[..., None, :].expand(xyz.shape[:-1] + (B, 3))
    return dq_apply((inv[0].expand(x.shape[:-1] + (4,)), inv[1].expand(x.shape[:-1] + (4,))), x) / gauss


def skin_logits(P, xyz, art, t_embed, inst_code, symm_idx=None):
    """SkinningField.forward (nnutils/skinning.py:89-120): -(|x_b|^2 + 0.1 relu(delta_mlp))."""
    xb = bone_coords(xyz, art, gauss_scale(P, symm_idx))
    dist2 = xb.pow(2).sum(-1)
    feat = torch.cat(
        [xb.reshape(xyz.shape[:-1] + (-1,)), _per_frame(t_embed, xyz), _per_frame(inst_code, xyz)], -1
    )
    delta = mlp(P, "warp.skinning_model.delta_field.", feat, 2, final_act=False)
    delta = F.relu(delta) * 0.1
    return -(dist2 + delta), delta


def dq_blend_apply(se3, pts, w):
    """dual_quaternion_skinning (utils/geom_utils.py:45-83)."""
    M, B = se3[0].shape[:2]
    nd = pts.dim() - 2
    shp = (M,) + (1,) * nd + (B, 4)
    qr = se3[0].reshape(shp).expand(pts.shape[:-1] + (B, 4))
    qd = se3[1].reshape(shp).expand(pts.shape[:-1] + (B, 4))
    anchor = w.argmax(-1)[..., None, None].expand(pts.shape[:-1] + (1, 4))
    sign = ((torch.gather(qr, -2, anchor) * qr).sum(-1) > 0).to(pts.dtype) * 2 - 1
    qr_w = (w[..., None] * sign[..., None] * qr).sum(-2)
    qd_w = (w[..., None] * sign[..., None] * qd).sum(-2)
    inv = qr_w.norm(dim=-1, keepdim=True).reciprocal()
    return dq_apply((qr_w * inv, qd_w * inv), pts)


def skinning_warp(P, xyz, t_art, rest_art, t_embed, t_embed_mean, inst_code, backward, symm_idx=None):
    """SkinningWarp.forward (nnutils/warping.py:277-336).  Forward warps evaluate the delta MLP
    with the MEAN time embedding (warping.py:313-314, skinning.py:110-111)."""
    if backward:
        se3 = dq_mul(rest_art, dq_inv(t_art))
        art, te = t_art, t_embed
    else:
        se3 = dq_mul(t_art, dq_inv(rest_art))
        art, te = rest_art, t_embed_mean.expand(xyz.shape[0], -1)
    skin, delta = skin_logits(P, xyz, art, te, inst_code, symm_idx)
    out = dq_blend_apply(se3, xyz, skin.softmax(-1))
    # cross_entropy_skin_loss (utils/loss_utils.py:22-43) = logsumexp - max
    ent = (torch.logsumexp(skin, -1) - skin.max(-1)[0])[..., None]
    return out, {"skin_entropy": ent, "delta_skin": delta.pow(2).mean(-1, keepdim=True)}


def dense_warp(P, x, t_embed, inst_code, backward):
    """DenseWarp.forward inside ComposedWarp (nnutils/warping.py:143-170, 445-483): x + 0.1 * CondMLP([PE6(x), t, inst])."""
    e = torch.cat([pos_embed(x, 6), _per_frame(t_embed, x), _per_frame(inst_code, x)], -1)
    prefix = "warp.post_warp.backward_map." if backward else "warp.post_warp.forward_map."
    return x + 0.1 * mlp(P, prefix, e, 2, final_act=False)


def gauss_density(P, xyz, rest_art):
    """Deformable.compute_gauss_density / SkinningWarp.get_gauss_density
    (nnutils/deformable.py:329-356, warping.py:355-387, utils/transforms.py:28-40)."""
    _, c = dq_to_qt((rest_art[0][:1], rest_art[1][:1]))  # (1,B,3)
    d2 = (xyz[..., None, :] - c.reshape((1,) * (xyz.dim() - 1) + c.shape[1:])).pow(2).sum(-1) / 0.01**2
    return (-0.5 * d2).exp().max(-1)[0][..., None] * P["warp.logibeta"].exp()


# --------------------------------------------------------------------------------------
# stage 3: the field MLPs
# --------------------------------------------------------------------------------------


def nerf_forward(P, cfg, xyz, inst_base, inst_color, dir=None, appr=None, get_density=True, alpha=None):
    """NeRF.forward (nnutils/nerf.py:167-215)."""
    e = torch.cat([pos_embed(xyz, cfg["L_xyz"], alpha), _per_frame(inst_base, xyz)], -1)
    feat = mlp(P, "basefield.", e, cfg["D"], final_act=True)
    sdf = F.linear(feat, P["sdf.weight"], P["sdf.bias"])
    if get_density:
        ibeta = P["logibeta"].exp()
        out = (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() * ibeta)) * ibeta
    else:
        out = sdf
    if dir is None:
        return out
    de = pos_embed(dir, cfg["L_dir"])
    if cfg["appr_channels"] > 0:
        de = torch.cat([de, _per_frame(appr, xyz)], -1)
    ec = torch.cat([pos_embed(xyz, cfg["L_xyz"] + 2, alpha), _per_frame(inst_color, xyz)], -1)
    feat = feat + mlp(P, "colorfield.", ec, 2, final_act=True)
    h = F.relu(F.linear(torch.cat([feat, de], -1), P["rgb.0.weight"], P["rgb.0.bias"]))
    rgb = F.linear(h, P["rgb.2.weight"], P["rgb.2.bias"]).sigmoid()
    return rgb, out


def vis_forward(P, xyz, inst_vis):
    """VisField.forward (nnutils/visibility.py:52-63)."""
    e = torch.cat([pos_embed(xyz, 10), _per_frame(inst_vis, xyz)], -1)
    return mlp(P, "vis_mlp.basefield.", e, 2, final_act=False)


def feature_forward(P, xyz):
    """FeatureNeRF.compute_feat (nnutils/feature.py:136-150)."""
    f = mlp(P, "feature_field.", pos_embed(xyz, 6), 5, final_act=False)
    return f / f.norm(dim=-1, keepdim=True)


def flip_pair(x):
    """NeRF.flip_pair (nnutils/nerf.py:929-946)."""
    if x.shape[0] < 2:
        return x
    return x.reshape((x.shape[0] // 2, 2) + x.shape[1:]).flip(1).reshape(x.shape)


def kmat_from_kinv(Kinv):
    """Kmatinv (utils/geom_utils.py:308-341): invert a pinhole (fx,fy,cx,cy) matrix."""
    K = torch.zeros_like(Kinv)
    K[..., 0, 0] = 1.0 / Kinv[..., 0, 0]
    K[..., 1, 1] = 1.0 / Kinv[..., 1, 1]
    K[..., 0, 2] = -Kinv[..., 0, 2] / Kinv[..., 0, 0]
    K[..., 1, 2] = -Kinv[..., 1, 2] / Kinv[..., 1, 1]
    K[..., 2, 2] = 1
    return K


# --------------------------------------------------------------------------------------
# query_field (training-mode path)
# --------------------------------------------------------------------------------------


def query_field(P, cfg, rays, tab, D, flow_thresh=None, alpha=None, eikonal_rays=None, depth=None):
    """{NeRF,FeatureNeRF,Deformable}.query_field in training mode
    (nnutils/nerf.py:580-684, feature.py:89-133, deformable.py:300-356).

    rays: hxy (M,N,3), Kinv (M,3,3), near_far (M,2).  tab: per-frame tables (see
    oracle/ref_harness.frame_tables).  Returns (feat_dict, deltas)."""
    hxy, Kinv, near_far = rays["hxy"], rays["Kinv"], rays["near_far"]
    q, t = tab["field2cam_q"], tab["field2cam_t"]
    xyz_cam, dir_cam, deltas, depth = sample_cam_rays(hxy, Kinv, near_far, D, depth=depth)
    xyz_t, dirf = cam_to_field(xyz_cam, dir_cam, q, t)
    skel = cfg["motion"] != "rigid"
    feat = {}
    if skel:
        t_art = (tab["t_articulation_qr"], tab["t_articulation_qd"])
        r_art = (tab["rest_articulation_qr"], tab["rest_articulation_qd"])
        wargs = (tab["skin_t_embed"], tab["skin_t_embed_mean"], tab["inst_skin"])
        xyz, aux_b = skinning_warp(P, xyz_t, t_art, r_art, *wargs, backward=True, symm_idx=cfg.get("symm_idx"))
        if cfg.get("dense", False):  # ComposedWarp: soft deformation after un-articulating (warping.py:472-476)
            xyz = dense_warp(P, xyz, tab["dense_t_embed"], tab["inst_dense_bwd"], backward=True)
    else:
        xyz = xyz_t
    feat["vis"] = vis_forward(P, xyz, tab["inst_vis"])
    rgb, density = nerf_forward(
        P, cfg, xyz, tab["inst_base"], tab["inst_color"], dir=dirf, appr=tab.get("appr_code"), alpha=alpha
    )
    feat["rgb"], feat["density"], feat["density_" + cfg["category"]] = rgb, density, density

    # flow (nnutils/nerf.py:948-997): warp with the pair-flipped frame's articulation and camera
    qn, tn, Kn = flip_pair(q), flip_pair(t), flip_pair(Kinv)
    if skel:
        t_art_n = (flip_pair(t_art[0]), flip_pair(t_art[1]))
        r_art_n = (flip_pair(r_art[0]), flip_pair(r_art[1]))
        x_in = xyz
        if cfg.get("dense", False):  # soft deformation of the partner frame first (warping.py:459-463)
            x_in = dense_warp(P, xyz, flip_pair(tab["dense_t_embed"]), tab["inst_dense_fwd"], backward=False)
        x_next, _ = skinning_warp(P, x_in, t_art_n, r_art_n, *wargs, backward=False, symm_idx=cfg.get("symm_idx"))
    else:
        x_next = xyz
    xc_next = field_to_cam(x_next, qn, tn)
    Kmat = kmat_from_kinv(Kn)
    h_next = torch.einsum("mij,m...j->m...i", Kmat, xc_next)
    h_next = h_next / (h_next[..., -1:] + 1e-6)
    flow = (h_next - hxy[:, :, None, :])[..., :2]
    valid = xc_next[..., -1:] > 1e-6
    if flow_thresh is not None:
        valid = valid & (flow.norm(dim=-1, keepdim=True) < float(flow_thresh))
    feat["flow"] = torch.cat([flow, valid.to(flow.dtype)], -1)

    # cycle consistency (nnutils/deformable.py:173-198, nerf.py:657-667)
    if skel:
        x_in = dense_warp(P, xyz, tab["dense_t_embed"], tab["inst_dense_fwd"], backward=False) if cfg.get("dense", False) else xyz
        x_cyc, aux_f = skinning_warp(P, x_in, t_art, r_art, *wargs, backward=False, symm_idx=cfg.get("symm_idx"))
        feat["cyc_dist"] = (x_cyc - xyz_t).norm(2, -1, keepdim=True)
        feat["delta_skin"] = (aux_f["delta_skin"] + aux_b["delta_skin"]) / 2
        feat["skin_entropy"] = (aux_f["skin_entropy"] + aux_b["skin_entropy"]) / 2
    else:
        z = torch.zeros_like(xyz[..., :1])
        feat["cyc_dist"], feat["delta_skin"], feat["skin_entropy"] = z, z.clone(), z.clone()

    # eikonal on a given subset of rays (nnutils/nerf.py:416-453); zeros elsewhere
    M, N = hxy.shape[:2]
    eik = torch.zeros(M * N, D, dtype=xyz.dtype, device=xyz.device)
    if eikonal_rays is not None:
        with torch.enable_grad():
            xs = xyz.reshape(M * N, D, 3)[eikonal_rays].detach().requires_grad_(True)
            ib = tab["inst_base"][:, None].expand(M, N, -1).reshape(M * N, -1)[eikonal_rays]
            s = nerf_forward(P, cfg, xs, ib, None, get_density=False, alpha=alpha)
            (g,) = torch.autograd.grad(s.sum(), xs)
        eik[eikonal_rays] = (g.norm(2, dim=-1) - 1) ** 2
    feat["eikonal"] = eik.reshape(M, N, D, 1)

    feat["xyz"], feat["xyz_cam"] = xyz, xyz_cam
    feat["depth"] = depth / P["logscale"].exp()
    if cfg.get("has_feature", False):
        feat["feature"] = feature_forward(P, xyz)
    if skel:
        feat["gauss_density"] = gauss_density(P, xyz, r_art)
    return feat, deltas


# --------------------------------------------------------------------------------------
# stage 4: compositing
# --------------------------------------------------------------------------------------


def compute_weights(density, deltas):
    """utils/render_utils.py:99-126: w_k = (1-e^{-tau_k}) exp(-sum_{j<k} tau_j); T_k = exp(-sum_{j<=k} tau_j)."""
    tau = (deltas * density)[..., 0]
    cs = torch.cumsum(tau, -1)
    T = torch.exp(-cs)
    Tprev = torch.cat([torch.ones_like(T[..., :1]), T[..., :-1]], -1)
    return (1 - torch.exp(-tau)) * Tprev, T


KEY_SKIP = ("density", "vis", "flow", "eikonal", "xy_reproj", "xyz_reproj", "gauss_density")
KEY_FREEZE = ("cyc_dist", "xyz_cam", "skin_entropy")


def render_pixel(feat, deltas):
    """render_pixel + integrate (utils/render_utils.py:59-96,129-184)."""
    w, T = compute_weights(feat["density"], deltas)
    out = {"mask": w.sum(-1, keepdim=True)}
    wn = w / (out["mask"] + 1e-6)
    for k, v in feat.items():
        if k in KEY_SKIP:
            continue
        ww = wn.detach() if k in KEY_FREEZE else wn
        out[k] = (ww[..., None] * v).sum(-2)
    if "flow" in feat:
        wf = w * feat["flow"][..., 2]
        wf = wf / (wf.sum(-1, keepdim=True) + 1e-6)
        out["flow"] = (wf[..., None] * feat["flow"][..., :2]).sum(-2)
    if "normal" in feat:
        out["normal"] = F.normalize(out["normal"], 2, -1)
    dkeys = [k for k in out if "density_" in k]
    dsum = torch.cat([out[k] for k in dkeys], -1).sum(-1, keepdim=True) + 1e-6
    for k in dkeys:
        out[k.replace("density_", "mask_")] = out.pop(k) / dsum
    if "eikonal" in feat:
        out["eikonal"] = feat["eikonal"].mean(dim=(-1, -2))
    if "delta_skin" in feat:
        out["delta_skin"] = feat["delta_skin"].mean(dim=(-1, -2))
    Td = T[..., None].detach()
    out["vis"] = -(F.logsigmoid(feat["vis"]) * Td).mean(-2) / Td.mean()
    if "gauss_density" in feat:
        gw, _ = compute_weights(feat["gauss_density"], deltas)
        out["gauss_mask"] = gw.sum(-1, keepdim=True)
    return out


def sample_pdf(bins, weights, n_importance, eps=1e-5):
    """sample_pdf (utils/render_utils.py:187-233), deterministic branch (u = linspace): inverse-CDF samples of the
    piecewise-constant pdf `weights` (R, n) over the bin edges `bins` (R, n+1)."""
    R, n = weights.shape
    w = weights + eps
    pdf = w / w.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1)  # (R, n+1)
    u = torch.linspace(0, 1, n_importance, dtype=bins.dtype, device=bins.device).expand(R, n_importance).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below, above = (inds - 1).clamp(min=0), inds.clamp(max=n)
    cdf_lo, cdf_hi = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b_lo, b_hi = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)
    return b_lo + (u - cdf_lo) / denom * (b_hi - b_lo)


def importance_sampling(P, cfg, rays, tab, D):
    """NeRF.importance_sampling in eval mode (nnutils/nerf.py:686-738): D/2 uniform samples -> density -> weights ->
    D/2 inverse-CDF samples on the mid-points (weights[1:-1]) -> merged and sorted depths (M,N,D,1); returns
    sample_cam_rays at those depths (xyz_cam, dir, deltas, depth)."""
    hxy, Kinv, near_far = rays["hxy"], rays["Kinv"], rays["near_far"]
    M, N = hxy.shape[:2]
    Dc = D // 2
    xyz_cam, dir_cam, deltas, depth = sample_cam_rays(hxy, Kinv, near_far, Dc)
    xyz_t, _ = cam_to_field(xyz_cam, dir_cam, tab["field2cam_q"], tab["field2cam_t"])
    if cfg["motion"] != "rigid":
        t_art = (tab["t_articulation_qr"], tab["t_articulation_qd"])
        r_art = (tab["rest_articulation_qr"], tab["rest_articulation_qd"])
        xyz, _ = skinning_warp(P, xyz_t, t_art, r_art, tab["skin_t_embed"], tab["skin_t_embed_mean"], tab["inst_skin"],
                               backward=True, symm_idx=cfg.get("symm_idx"))
        if cfg.get("dense", False):
            xyz = dense_warp(P, xyz, tab["dense_t_embed"], tab["inst_dense_bwd"], backward=True)
    else:
        xyz = xyz_t
    density = nerf_forward(P, cfg, xyz, tab["inst_base"], tab["inst_color"])
    weights, _ = compute_weights(density, deltas)
    depth_mid = 0.5 * (depth[:, :, :-1] + depth[:, :, 1:]).reshape(-1, Dc - 1)
    fine = sample_pdf(depth_mid, weights.reshape(-1, Dc)[:, 1:-1], Dc).reshape(M, N, Dc, 1)
    merged, _ = torch.sort(torch.cat([depth, fine], -2), -2)
    return sample_cam_rays(hxy, Kinv, near_far, D, depth=merged)


def compose_fields(feats, deltas_list):
    """MultiFields.compose_fields (nnutils/multifields.py:339-398): concatenate along the sample
    dim (zeros for keys a field lacks), then depth-sort every key and the deltas."""
    keys = []
    for f in feats:
        for k in f:
            if k not in keys:
                keys.append(k)
    out = {}
    for k in keys:
        ref = next(f[k] for f in feats if k in f)
        out[k] = torch.cat([f[k] if k in f else torch.zeros_like(ref) for f in feats], 2)
    deltas = torch.cat(deltas_list, 2)
    if len(feats) > 1:
        idx = out["depth"].argsort(2)
        out = {k: torch.gather(v, 2, idx.expand_as(v)) for k, v in out.items()}
        deltas = torch.gather(deltas, 2, idx.expand_as(deltas))
    return out, deltas


# field configurations of MultiFields.define_field (nnutils/multifields.py:60-100)
CFG_FG_BOB = dict(category="fg", D=8, W=256, L_xyz=10, L_dir=-1, appr_channels=32, motion="bob", B=25, has_feature=True)
CFG_FG_RIGID = dict(CFG_FG_BOB, motion="rigid", B=0)
CFG_BG = dict(category="bg", D=5, W=128, L_xyz=6, L_dir=0, appr_channels=0, motion="rigid", B=0, has_feature=False)


# --------------------------------------------------------------------------------------
# stage 5: per-ray feature matching
# --------------------------------------------------------------------------------------


def global_match(feat_px, feat_canonical, xyz_canonical, logsigma, idx):
    """FeatureNeRF.global_match (nnutils/feature.py:152-205) with the candidate draw `idx` given: softmax of the scaled
    feature similarities over the candidates, expected canonical point.  feat_px (...,C) -> (...,3)."""
    shape = feat_px.shape
    fc = feat_canonical.reshape(-1, shape[-1])[idx]
    xc = xyz_canonical.reshape(-1, 3)[idx]
    score = (feat_px.reshape(-1, shape[-1]) @ fc.t()) * logsigma.exp()
    prob = torch.softmax(score, dim=1)
    return (prob[..., None] * xc).sum(1).view(shape[:-1] + (3,))
