"""CPU, world_size 2, gloo: ray sharding keeps frame pairs together and partitions the batch; the
gradient all-reduce averages like DDP; sharded oracle renders concatenate to the unsharded render."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lab4d_oracle as O
    import synth
    from lab4d_b200 import parallel, spec
    from util import synth_params

    cfg = spec.BG
    P = synth_params(cfg, 0)
    M, N, D = 4, 6, 8
    rays = {k: torch.from_numpy(v) for k, v in synth.synth_rays(M, N, seed=3).items()}
    tab = {"field2cam_q": rays["field2cam"][:, :4].contiguous(), "field2cam_t": (rays["field2cam"][:, 4:] * 0.2).contiguous(),
           "inst_base": torch.zeros(M, 32), "inst_color": torch.zeros(M, 32), "inst_vis": torch.zeros(M, 32)}
    r_loc, t_loc = parallel.shard_batch(rays, tab, rank, world)
    feat, dl = O.query_field(P, cfg.as_oracle_cfg(), r_loc, t_loc, D)
    rgb = O.render_pixel(feat, dl)["rgb"]
    g = torch.full((5,), float(rank + 1))
    parallel.allreduce_mean_(g)
    q.put((rank, parallel.shard_frames(M, rank, world), rgb.numpy(), g.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_allreduce():
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import lab4d_oracle as O
    import synth
    from lab4d_b200 import spec
    from util import synth_params

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == (0, 2) and res[1][1] == (2, 4)
    for r in res:
        assert np.allclose(r[3], 1.5)  # mean of 1 and 2
    cfg = spec.BG
    P = synth_params(cfg, 0)
    M, N, D = 4, 6, 8
    rays = {k: torch.from_numpy(v) for k, v in synth.synth_rays(M, N, seed=3).items()}
    tab = {"field2cam_q": rays["field2cam"][:, :4].contiguous(), "field2cam_t": (rays["field2cam"][:, 4:] * 0.2).contiguous(),
           "inst_base": torch.zeros(M, 32), "inst_color": torch.zeros(M, 32), "inst_vis": torch.zeros(M, 32)}
    feat, dl = O.query_field(P, cfg.as_oracle_cfg(), rays, tab, D)
    full = O.render_pixel(feat, dl)["rgb"].numpy()
    assert np.allclose(np.concatenate([res[0][2], res[1][2]], 0), full, atol=1e-6)


def test_shard_frames_properties():
    from lab4d_b200.parallel import shard_frames

    for M in (2, 8, 10, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_frames(M, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == M
            for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
                assert b == c
            assert all(a % 2 == 0 and b % 2 == 0 for a, b in spans)
    with pytest.raises(ValueError):
        shard_frames(3, 0, 1)


def _grad_worker(rank, world, port, q):
    """One rank of the training step's collective: per-rank gradients of a parameter dict -> flat buffer -> mean all-reduce
    -> written back into .grad (what bench.py / the adapter do with the renderer's flat gradient buffer over NCCL)."""
    import torch.distributed as dist

    from lab4d_b200 import parallel

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    P = {"a.weight": torch.randn(5, 7, requires_grad=True), "a.bias": torch.randn(5, requires_grad=True), "unused": torch.randn(3, requires_grad=True)}
    x = torch.full((7,), float(rank + 1))
    ((P["a.weight"] @ x + P["a.bias"]) ** 2).sum().backward()
    local = {k: v.grad.clone() for k, v in P.items() if v.grad is not None}
    flat = parallel.flat_grads([P])
    assert flat.numel() == 5 * 7 + 5  # parameters without a gradient do not enter the buffer
    parallel.allreduce_mean_(flat)
    assert parallel.unflatten_grads_(flat, [P]) == flat.numel()
    q.put((rank, {k: v.numpy() for k, v in local.items()}, {k: v.grad.numpy().copy() for k, v in P.items() if v.grad is not None}))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k in ("a.weight", "a.bias"):
        mean = 0.5 * (res[0][1][k] + res[1][1][k])
        assert not np.allclose(res[0][1][k], res[1][1][k])
        for r in res:
            assert np.allclose(r[2][k], mean, rtol=1e-6), k
