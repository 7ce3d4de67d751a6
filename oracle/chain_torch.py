"""TEST INFRASTRUCTURE (checker of csrc/chain.cu): torch restatement of the backward of the per-frame prologue. Chains the gradients of the kernel-level blocks (constant block, M per-frame blocks:
include/b200r.h b200r_block_layout) to the reference's parameters and to the per-frame inputs of query_field.

The blocks are what csrc/prologue.cu builds every call from M rows of per-frame data - cameras, bias rows with the
per-frame codes folded in (nnutils/base.py:140-146, nerf.py:200-204, skinning.py:109-116), inverse bone transforms
(utils/transforms.py:9-25) and blend transforms (nnutils/warping.py:304-314).  Their backward is M x B rows of quaternion
algebra and a few (M x 32) mat-muls, <0.1 % of the step: plain torch on the GPU (autograd of a restatement of the prologue's
table formulas), like the per-frame MLPs that produce the inputs."""
import torch

CODE_KEYS = ["inst_base", "inst_color", "inst_vis", "appr_code", "inst_skin", "skin_t_embed", "skin_t_embed_mean", "dense_t_embed",
             "dense_t_embed", "inst_dense_fwd", "inst_dense_bwd"]  # b200r_block_layout code ids (8 = partner frame's dense_t)


def flip_pair(x):
    """Swap the two frames of every adjacent pair (nnutils/nerf.py:929-946); a single frame is its own partner."""
    if x.shape[0] < 2:
        return x
    return x.reshape(x.shape[0] // 2, 2, *x.shape[1:]).flip(1).reshape(x.shape)


def _qmul(a, b):
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw), -1)


def _qconj(q):
    return torch.cat((q[..., :1], -q[..., 1:]), -1)


def _qrot(q, p):
    pq = torch.cat((torch.zeros_like(p[..., :1]), p), -1)
    return _qmul(_qmul(q, pq), _qconj(q))[..., 1:]


def _binv(qr, qd, inv_gauss):
    """prologue.cu write_binv: (..., B, 12) rows (R'_i0 R'_i1 R'_i2 t'_i) of the scaled inverse bone transform."""
    q, qdc = _qconj(qr), _qconj(qd)
    t = 2.0 * _qmul(qdc, _qconj(q))[..., 1:]
    w, x, y, z = q.unbind(-1)
    ww, xx, yy, zz, xy, xz, yz, wx, wy, wz = w * w, x * x, y * y, z * z, x * y, x * z, y * z, w * x, w * y, w * z
    R = torch.stack((torch.stack((ww + xx - yy - zz, 2 * (xy - wz), 2 * (xz + wy)), -1),
                     torch.stack((2 * (xy + wz), ww - xx + yy - zz, 2 * (yz - wx)), -1),
                     torch.stack((2 * (xz - wy), 2 * (yz + wx), ww - xx - yy + zz), -1)), -2)  # (..., B, 3, 3)
    rows = torch.cat((R, t[..., :, None]), -1) * inv_gauss[..., :, None]
    return rows.reshape(*rows.shape[:-2], 12)


def _se3(ar, ad, br, bd):
    """prologue.cu write_se3: a (x) b^-1 as (real, dual), (..., B, 8)."""
    bir, bid = _qconj(br), _qconj(bd)
    return torch.cat((_qmul(ar, bir), _qmul(ar, bid) + _qmul(ad, bir)), -1)


_SYMM = {}


def _symm_index(symm_idx, device):
    """Device index tensor of the left/right bone pairing, built once (a list index would copy from the host every call)."""
    key = (tuple(symm_idx), str(device))
    if key not in _SYMM:
        _SYMM[key] = torch.tensor(list(symm_idx), dtype=torch.long, device=device)
    return _SYMM[key]


def chain(layout, layer_names, cfg, P, tab, rays, g_const, g_frame, weight_grads):
    """g_const (C,), g_frame (M, F): gradients of the blocks.  weight_grads: name -> (out, in) gradient views (the kernel
    filled the columns fed by per-sample operands; the code columns are added here).  Returns (param_grads, table_grads):
    name -> tensor for every hot-path parameter / per-frame input that receives a gradient."""
    M = g_frame.shape[0]
    pg = dict(weight_grads)
    tg = {}

    def acc(d, k, v):
        d[k] = d[k] + v if k in d else v

    W, HN = cfg.W, cfg.W // 2
    # ---------------------------------------------------------------- constant block
    for i, name in enumerate(layer_names):
        off = layout.c_plain_bias[i]
        if off >= 0:
            n = P[name + ".bias"].shape[0]
            acc(pg, name + ".bias", g_const[off:off + n])
    acc(pg, "sdf.weight", g_const[layout.c_sdf_w:layout.c_sdf_w + W].reshape(1, W))
    acc(pg, "rgb.2.weight", g_const[layout.c_rgb2_w:layout.c_rgb2_w + 3 * HN].reshape(3, HN))
    acc(pg, "vis_mlp.basefield.linear_final.weight", g_const[layout.c_vis_w:layout.c_vis_w + 64].reshape(1, 64))
    sc = g_const[layout.c_scalars:layout.c_scalars + 8]
    acc(pg, "logibeta", sc[0:1])
    acc(pg, "logscale", sc[1:2])
    acc(pg, "sdf.bias", sc[3:4])
    acc(pg, "rgb.2.bias", sc[4:7])
    acc(pg, "vis_mlp.basefield.linear_final.bias", sc[7:8])
    skinned = cfg.motion != "rigid"
    if skinned:
        acc(pg, "warp.logibeta", sc[2:3])
    # ---------------------------------------------------------------- bias rows that carry per-frame codes
    codes = {}
    for cid, key in enumerate(CODE_KEYS):
        if tab.get(key) is not None:
            c = tab[key]
            if cid == 6:
                c = c.reshape(1, -1).expand(M, -1)
            elif cid == 8:
                c = flip_pair(c)
            codes[cid] = c
    for ci in range(layout.n_cond):
        c = layout.cond[ci]
        name = layer_names[c.layer]
        G = g_frame[:, c.frame_off:c.frame_off + c.n]
        acc(pg, name + ".bias", G.sum(0))
        Wl = P[name + ".weight"]
        for sgi in range(c.n_seg):
            col0, width, cid = c.col0[sgi], c.width[sgi], c.code[sgi]
            if cid not in codes:
                continue
            code = codes[cid]
            pg[name + ".weight"][:, col0:col0 + width] += G.t() @ code
            gc = G @ Wl[:, col0:col0 + width]
            if cid == 6:
                gc = gc.sum(0)
            elif cid == 8:
                gc = flip_pair(gc)
            acc(tg, CODE_KEYS[cid], gc)
    # ---------------------------------------------------------------- cameras
    gc = g_frame[:, layout.f_cam:layout.f_cam + 24]
    gp = g_frame[:, layout.f_cam_partner:layout.f_cam_partner + 24]
    tg["Kinv"] = (gc[:, 0:9] + flip_pair(gp[:, 0:9])).reshape(M, 3, 3)
    q = tab["field2cam_q"].detach().requires_grad_(True)
    t = tab["field2cam_t"].detach().requires_grad_(True)
    with torch.enable_grad():
        qi = _qconj(q)
        ti = _qrot(qi, -t)
        gq, gt = torch.autograd.grad([qi, ti], [q, t], [gc[:, 11:15], gc[:, 15:18]])
    tg["field2cam_q"] = gq + flip_pair(gp[:, 11:15])
    tg["field2cam_t"] = gt + flip_pair(gp[:, 15:18])
    # ---------------------------------------------------------------- bone tables
    if skinned:
        B = cfg.B
        ins = [tab[k].detach().requires_grad_(True) for k in ("t_articulation_qr", "t_articulation_qd", "rest_articulation_qr", "rest_articulation_qd")]
        lg = P["warp.skinning_model.log_gauss"].detach().requires_grad_(True)
        with torch.enable_grad():
            tqr, tqd, rqr, rqd = ins
            lgs = lg
            if cfg.symm_idx is not None:
                lgs = 0.5 * (lg[_symm_index(cfg.symm_idx, lg.device)] + lg)
            ig = torch.exp(-lgs)
            tables = [(_binv(tqr, tqd, ig), layout.f_binv_t, 12), (_se3(rqr, rqd, tqr, tqd), layout.f_se3_bwd, 8),
                      (_binv(rqr, rqd, ig), layout.f_binv_rest, 12), (_se3(tqr, tqd, rqr, rqd), layout.f_se3_fwd, 8),
                      (flip_pair(_binv(rqr, rqd, ig)), layout.f_binv_rest_partner, 12),
                      (flip_pair(_se3(tqr, tqd, rqr, rqd)), layout.f_se3_fwd_partner, 8)]
            # rest bone centres (frame 0) of the Gaussian bone density
            ctr = 2.0 * _qmul(rqd[:1], _qconj(rqr[:1]))[..., 1:]
            outs = [tbl for tbl, _, _ in tables] + [ctr]
            gouts = [g_frame[:, off:off + B * per].reshape(M, B, per) for _, off, per in tables]
            gouts.append(g_const[layout.c_center:layout.c_center + 4 * B].reshape(1, B, 4)[..., :3])
            grads = torch.autograd.grad(outs, ins + [lg], gouts, allow_unused=True)
        for k, gv in zip(("t_articulation_qr", "t_articulation_qd", "rest_articulation_qr", "rest_articulation_qd"), grads[:4]):
            if gv is not None:
                tg[k] = gv
        if grads[4] is not None:
            acc(pg, "warp.skinning_model.log_gauss", grads[4])
    return pg, tg
