"""Drive the UNMODIFIED reference (/root/reference, CPU, via oracle/ref_shims) on synthetic batches.
Runs ONLY in the build container (the reference does not exist on the GPU box); used by
oracle/gen_golden.py to write tests/golden/*.npz.  TEST INFRASTRUCTURE."""
import functools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _install  # noqa: E402

_install.install()

import lab4d.nnutils.nerf as rnerf  # noqa: E402
from lab4d.nnutils.multifields import MultiFields  # noqa: E402
from lab4d.utils.render_utils import render_pixel, sample_cam_rays  # noqa: E402

import synth  # noqa: E402


def make_data_info(T=40, nvid=1):
    rt = np.tile(np.eye(4, dtype=np.float32), (2, T * nvid, 1, 1))
    rt[..., 2, 3] = 3.0
    off = np.arange(nvid + 1) * T
    return dict(
        rtmat=rt,
        geom_path=["", ""],
        vis_info={"bg": 0, "fg": 1},
        frame_info=dict(frame_offset=off, frame_offset_raw=off, frame_mapping=list(range(T * nvid))),
        intrinsics=np.tile(np.array([200, 200, 128, 128], np.float32), (T * nvid, 1)),
    )


def build_field(field_type="fg", fg_motion="bob", T=40, nvid=1, seed=0, train=True):
    """Reference field with the synthetic parameter rule loaded by name."""
    num_inst = nvid
    mf = MultiFields(make_data_info(T, nvid), field_type=field_type, fg_motion=fg_motion, num_inst=num_inst)
    for cat, field in mf.field_params.items():
        sd = field.state_dict()
        new = {}
        for k, v in sd.items():
            if k == "aabb":
                new[k] = v
            elif v.dtype.is_floating_point:
                new[k] = torch.from_numpy(synth.synth_tensor(k, v.shape, seed, cat))
            else:
                new[k] = v
        field.load_state_dict(new)
    mf.train(train)
    mf.set_beta_prob(0.0)
    return mf


def set_n_depth(D):
    """The reference hard-wires 64 samples/ray (utils/render_utils.py:8, nnutils/nerf.py:617-619)."""
    rnerf.sample_cam_rays = functools.partial(sample_cam_rays, n_depth=D)


def frame_tables(field, rays):
    """Per-frame inputs of the hot path that the reference computes with small (out-of-scope) MLPs."""
    fid = torch.from_numpy(rays["frame_id"])
    iid = torch.from_numpy(rays["inst_id"])
    tabs = {}
    with torch.no_grad():
        if field.appr_channels > 0:
            tabs["appr_code"] = field.appr_embedding.get_vals(fid)
        tabs["inst_base"] = field.basefield.inst_embedding(iid)
        tabs["inst_color"] = field.colorfield.inst_embedding(iid)
        tabs["inst_vis"] = field.vis_mlp.basefield.inst_embedding(iid)
        if hasattr(field, "warp") and hasattr(field.warp, "skinning_model"):
            sk = field.warp.skinning_model
            tabs["skin_t_embed"] = sk.time_embedding(fid)
            tabs["skin_t_embed_mean"] = sk.time_embedding.get_mean_embedding(fid.device)
            tabs["inst_skin"] = sk.delta_field.inst_embedding(iid)
            if hasattr(field.warp, "post_warp"):
                pw = field.warp.post_warp
                tabs["dense_t_embed"] = pw.time_embedding(fid)
                tabs["inst_dense_fwd"] = pw.forward_map.inst_embedding(iid)
                tabs["inst_dense_bwd"] = pw.backward_map.inst_embedding(iid)
    return {k: v.detach().numpy() for k, v in tabs.items()}


def make_batch(field, rays, device="cpu"):
    """(Kinv, batch) of one field from a synthetic ray dict, on `device` (the module must live there too)."""
    batch = {
        "hxy": torch.from_numpy(rays["hxy"]).to(device),
        "frameid": torch.from_numpy(rays["frame_id"]).to(device),
        "dataid": torch.from_numpy(rays["inst_id"]).to(device),
        "field2cam": torch.from_numpy(rays["field2cam"]).to(device),
    }
    field.near_far.data = torch.from_numpy(rays["near_far"])[:1].repeat(field.near_far.shape[0], 1).to(device)
    return torch.from_numpy(rays["Kinv"]).to(device), batch


def run_field(mf, category, rays, D, flow_thresh=None, record_eikonal=True):
    """get_samples -> query_field -> render_pixel on one field; returns numpy dicts."""
    set_n_depth(D)
    field = mf.field_params[category]
    M = rays["hxy"].shape[0]
    Kinv, batch = make_batch(field, rays)
    picked = {}
    orig_multinomial = torch.multinomial

    def rec_multinomial(probs, n, replacement=False):
        out = orig_multinomial(probs, n, replacement=replacement)
        picked["eikonal_rays"] = out.clone()
        return out

    torch.multinomial = rec_multinomial
    try:
        samples = field.get_samples(Kinv, batch)
        feat, deltas, aux = field.query_field(samples, flow_thresh=flow_thresh)
    finally:
        torch.multinomial = orig_multinomial
    rendered = render_pixel(feat, deltas)
    tables = frame_tables(field, rays)
    tables["field2cam_q"] = samples["field2cam"][0].detach().numpy()
    tables["field2cam_t"] = samples["field2cam"][1].detach().numpy()
    if "t_articulation" in samples:
        for nm in ("t_articulation", "rest_articulation"):
            tables[nm + "_qr"] = samples[nm][0].detach().numpy()
            tables[nm + "_qd"] = samples[nm][1].detach().numpy()
    if "eikonal_rays" in picked:
        tables["eikonal_rays"] = picked["eikonal_rays"].numpy()
    return (
        {k: v.detach() for k, v in feat.items()},
        deltas.detach(),
        {k: v.detach() for k, v in rendered.items()},
        tables,
        (feat, deltas, rendered, samples),
    )
