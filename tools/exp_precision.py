"""Experiment: which stage dominates the rendered-RGB error on the trained-like field?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")): sys.path.insert(0, p)
import numpy as np, torch
import lab4d_oracle as O, synth
from lab4d_b200 import spec
from lab4d_b200.render import FieldRenderer, render_pixel
from test_gpu_parity import synth_tables
from util import rel_l2
DEV = "cuda"
cfg = spec.FG_BOB
g = torch.Generator(device="cpu").manual_seed(0)
P = {}
shapes = spec.field_param_shapes(cfg)
for k, shp in shapes.items():
    if k.endswith(".weight") and len(shp) == 2: P[k] = ((torch.rand(shp, generator=g) * 2 - 1) / np.sqrt(shp[1])).to(DEV)
    elif k.endswith(".bias"): P[k] = ((torch.rand(shp, generator=g) * 2 - 1) / np.sqrt(shapes[k[:-4] + "weight"][1])).to(DEV)
P["logibeta"] = torch.tensor([-np.log(0.1)], dtype=torch.float32, device=DEV); P["logscale"] = torch.tensor([np.log(0.2)], dtype=torch.float32, device=DEV)
P["warp.logibeta"] = torch.tensor([-np.log(0.01)], dtype=torch.float32, device=DEV); P["warp.skinning_model.log_gauss"] = torch.full((25, 3), float(np.log(0.03)), device=DEV)
fit = [k for k in P if k.startswith("basefield.") or k.startswith("sdf.")]
for k in fit: P[k].requires_grad_(True)
opt = torch.optim.Adam([P[k] for k in fit], lr=1e-3); inst = torch.zeros(1, 32, device=DEV)
for _ in range(500):
    opt.zero_grad(); pts = (torch.rand(256, 3, device=DEV) * 2 - 1) * 0.18
    sdf = O.nerf_forward(P, cfg.as_oracle_cfg(), pts[None], inst, None, get_density=False)[0]; gt = pts.norm(dim=-1, keepdim=True) - 0.1
    scale = ((sdf * gt).sum() / (sdf * sdf).sum()).detach(); ((sdf * scale - gt) ** 2).mean().backward(); opt.step()
P = {k: v.detach() for k, v in P.items()}
M, N, D = 8, 32, 128
rays = {k: torch.from_numpy(v).to(DEV) for k, v in synth.synth_rays(M, N, seed=11, spread=30.0).items()}
rays["near_far"] = torch.tensor([[0.4, 0.8]], device=DEV).repeat(M, 1)
tab = synth_tables(cfg, M, DEV, seed=11, rays=rays, P=P)
for k in ("inst_base", "inst_color", "inst_vis", "inst_skin"): tab[k] = torch.zeros(M, 32, device=DEV)
tab["field2cam_t"] = torch.tensor([[0.0, 0.0, 0.6]], device=DEV).repeat(M, 1)
for k in ("t_articulation_qd", "rest_articulation_qd"): tab[k] = tab[k] * 0.05
def run(P, tab, tag):
    ofeat, odel = O.query_field(P, cfg.as_oracle_cfg(), rays, tab, D); orend = O.render_pixel(ofeat, odel)
    r = FieldRenderer(cfg, DEV); r.pack(P); feat, deltas = r.query_field(P, rays, tab, D); rend = render_pixel(feat, deltas)
    ps = {k: rel_l2(feat[k].cpu(), ofeat[k].cpu()) for k in ("xyz", "density", "rgb", "vis", "feature")}
    pr = {k: rel_l2(rend[k].cpu(), orend[k].cpu()) for k in ("rgb", "depth", "mask")}
    print(tag, "per-sample", {k: f"{v:.1e}" for k, v in ps.items()}, "rendered", {k: f"{v:.1e}" for k, v in pr.items()})
run(P, tab, "full        ")
for sd in (1, 2, 3):
    torch.manual_seed(sd)
    Q = {k: v.clone() for k, v in P.items()}
    for k in fit: Q[k] = ((torch.rand(Q[k].shape, device=DEV) * 2 - 1) / np.sqrt(shapes[k][1] if k.endswith("weight") else shapes[k[:-4] + "weight"][1])).requires_grad_(True)
    opt = torch.optim.Adam([Q[k] for k in fit], lr=1e-3)
    for _ in range(500):
        opt.zero_grad(); pts = (torch.rand(256, 3, device=DEV) * 2 - 1) * 0.18
        sdf = O.nerf_forward(Q, cfg.as_oracle_cfg(), pts[None], inst, None, get_density=False)[0]; gt = pts.norm(dim=-1, keepdim=True) - 0.1
        scale = ((sdf * gt).sum() / (sdf * sdf).sum()).detach(); ((sdf * scale - gt) ** 2).mean().backward(); opt.step()
    Q = {k: v.detach() for k, v in Q.items()}
    run(Q, tab, f"refit seed {sd}")
P2 = dict(P); P2["warp.skinning_model.delta_field.linear_final.weight"] = torch.zeros_like(P["warp.skinning_model.delta_field.linear_final.weight"])
run(P2, tab, "delta W3 = 0")
tab3 = dict(tab)
for nm in ("t_articulation", "rest_articulation"):
    qr = torch.zeros_like(tab[nm + "_qr"]); qr[..., 0] = 1; tab3[nm + "_qr"] = qr; tab3[nm + "_qd"] = torch.zeros_like(tab[nm + "_qd"])
run(P, tab3, "identity art")
