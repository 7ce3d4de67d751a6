"""GPU box: where the time of the PATCHED reference step goes (eager launches through the reference's own entry points):
wall time per phase with a synchronize after each, and the launch count of one step (torch.profiler)."""
import collections
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import torch

import ref_harness as H
import synth


def main():
    dev = "cuda"
    M, N, D = 128, 16, 128
    mf = H.build_field("fg", "bob", seed=0).to(dev)
    field = mf.field_params["fg"]
    H.set_n_depth(D)
    Kinv, batch = H.make_batch(field, synth.synth_rays(M, N, seed=10), dev)
    g = torch.Generator().manual_seed(3)
    batch["feature"] = torch.nn.functional.normalize(torch.randn(M, N, 16, generator=g), dim=-1).to(dev)
    from lab4d_b200 import autograd as ag
    from lab4d_b200 import nnutils, render

    dq = "--dq" in sys.argv  # quaternion kernels for the reference's remaining torch code + candidate draw on the device
    undo = nnutils.install(n_depth=D, dqtorch=dq, match_rng="device" if dq else "reference")
    print(f"install(dqtorch={dq})")
    import lab4d.utils.render_utils as rru

    T = collections.OrderedDict()

    def timed(name, fn):
        def w(*a, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn(*a, **k)
            torch.cuda.synchronize()
            T[name] = T.get(name, 0.0) + time.perf_counter() - t0
            return out
        return w

    field.get_samples = timed("get_samples (camera / articulation modules)", field.get_samples)
    nnutils.tables_from_module = timed("tables_from_module (embeddings)", nnutils.tables_from_module)
    ag.query_field = timed("FieldFunction forward (pack + kernels)", ag.query_field)
    nnutils.compute_eikonal = timed("compute_eikonal (kernels)", nnutils.compute_eikonal)
    render.global_match = timed("global_match (kernels)", render.global_match)
    field.forward_project = timed("forward_project (reference torch warp)", field.forward_project)
    rp = timed("render_pixel", rru.render_pixel)

    def step():
        field.zero_grad()
        s = field.get_samples(Kinv, batch)
        feat, deltas, aux = field.query_field(s, flow_thresh=None)
        r = rp(feat, deltas)
        loss = r["rgb"].sum() + r["mask"].sum() + r["flow"].sum() * 1e-3 + r["eikonal"].mean() + 1e-3 * aux["xy_reproj"].mean()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss.backward()
        torch.cuda.synchronize()
        T["backward (all)"] = T.get("backward (all)", 0.0) + time.perf_counter() - t0

    for _ in range(3):
        step()
    T.clear()
    n = 5
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) / n
    print(f"patched step (with a synchronize after every phase): {tot*1e3:.1f} ms")
    for k, v in T.items():
        print(f"  {k}: {v / n * 1e3:.2f} ms")
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    ev = prof.key_averages()
    kern = [e for e in ev if e.device_type == torch.autograd.DeviceType.CUDA]
    nk = sum(e.count for e in kern)
    tk = sum(e.device_time_total for e in kern) if hasattr(kern[0], "device_time_total") else sum(e.cuda_time_total for e in kern)
    print(f"one step: {nk} device kernels / copies, {tk / 1e3:.2f} ms of device time")
    top = sorted(kern, key=lambda e: -(getattr(e, "device_time_total", None) or e.cuda_time_total))[:12]
    for e in top:
        print(f"  {e.count:5d} x {e.key[:90]}: {(getattr(e, 'device_time_total', None) or e.cuda_time_total) / 1e3:.3f} ms")
    undo()


if __name__ == "__main__":
    main()
