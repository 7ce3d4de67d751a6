"""CPU restatement of the per-pixel reconstruction losses of dvr_model (lab4d/engine/model.py): compute_recon_loss (:415-498),
mask_losses (:520-574), get_mask_balance_wt (:386-412) and apply_loss_weights (:576-611) for the keys compute_recon_loss
creates, as ONE pure function - the checker of the loss kernels (csrc/losses.cu).  TEST INFRASTRUCTURE: only tests/ import
this; tests/test_loss_oracle_cpu.py pins it to the unmodified reference's static methods."""
import torch

KEYS = ("mask", "feature", "feat_reproj", "rgb", "depth", "flow", "vis", "reg_gauss_mask")
PX_UNIT = ("flow", "feat_reproj")


def mask_balance_wt(mask, vis2d, is_detected):
    """model.py:386-412."""
    mask = mask.float()
    vis2d = vis2d.float() * is_detected.float()[:, None, None]
    if mask.sum() > 0 and (1 - mask).sum() > 0:
        pos_wt = vis2d.sum() / mask[vis2d > 0].sum()
        neg_wt = vis2d.sum() / (1 - mask[vis2d > 0]).sum()
        return 0.5 * pos_wt * mask + 0.5 * neg_wt * (1 - mask)
    return 1


def recon_losses(rendered, aux, batch, field_type, config):
    """Final weighted scalars of the reconstruction terms, in the reference's key order.  rendered / aux[cate]: (M,N,c)
    tensors; batch: mask, vis2d (M,N,1), is_detected (M,), rgb, depth, flow, flow_uct, feature, hxy; config: `<key>_wt`, train_res."""
    L = {}
    maskfg, vis2d = batch["mask"].float(), batch["vis2d"].float()
    det = batch["is_detected"].float()[:, None, None]
    rfg = {"fg": rendered.get("mask"), "comp": rendered.get("mask_fg"), "bg": None}[field_type]
    wbal = mask_balance_wt(batch["mask"], batch["vis2d"], batch["is_detected"])
    if field_type == "bg":
        L["mask"] = (rendered["mask"] - 1).pow(2)
    elif field_type == "fg":
        L["mask"] = (rfg - maskfg).pow(2) * wbal
    else:
        L["mask"] = (rfg - maskfg).pow(2) * wbal + (rendered["mask"] - 1).pow(2)
    if field_type in ("fg", "comp"):
        L["feature"] = (aux["fg"]["feature"] - batch["feature"]).norm(2, -1, keepdim=True)
        L["feat_reproj"] = (aux["fg"]["xy_reproj"] - batch["hxy"][..., :2]).norm(2, -1, keepdim=True)
    L["rgb"] = (rendered["rgb"] - batch["rgb"]).pow(2)
    L["depth"] = (rendered["depth"] - batch["depth"]).norm(2, -1, keepdim=True)
    L["flow"] = (rendered["flow"] - batch["flow"]).norm(2, -1, keepdim=True) * (batch["flow_uct"] > 0).float()
    L["vis"] = sum((0.01 if cate == "bg" else 1.0) * a["vis"] for cate, a in aux.items())
    if "gauss_mask" in rendered:
        L["reg_gauss_mask"] = (aux["fg"]["gauss_mask"] - rfg.detach()).pow(2)
    # mask_losses
    mtype = {"bg": (1 - maskfg) * vis2d, "fg": maskfg * vis2d, "comp": vis2d}[field_type]
    for k in L:
        if k == "reg_gauss_mask":
            continue
        L[k] = L[k] * (vis2d if k == "mask" else (maskfg if k in ("feature", "feat_reproj") else mtype))
        if k in ("mask", "feature", "feat_reproj"):
            L[k] = L[k] * det
    # apply_loss_weights
    out = {}
    for k, v in L.items():
        s = v[v > 0].mean()
        if k in PX_UNIT:
            s = s / config["train_res"]
        if k + "_wt" in config:
            s = s * config[k + "_wt"]
        out[k] = s
    return out
