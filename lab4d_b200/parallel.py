"""Data-parallel plumbing: rays shard over ranks, one all-reduce of the flat gradient buffer per step
(the reference's only collective is DDP's bucketed gradient all-reduce, engine/trainer.py:110-115,345).
Frames stay in adjacent pairs on every rank (flip_pair, nnutils/nerf.py:929-946)."""
import torch
import torch.distributed as dist


def shard_frames(M, rank, world):
    """Contiguous block of frame PAIRS for `rank`: returns (start, stop) frame indices, both even."""
    if M % 2:
        raise ValueError("frames come in pairs")
    pairs = M // 2
    base, rem = divmod(pairs, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return 2 * lo, 2 * hi


def shard_batch(rays, tab, rank, world):
    """Slice every per-frame tensor of a ray batch / frame-table dict to this rank's frames."""
    M = rays["hxy"].shape[0]
    lo, hi = shard_frames(M, rank, world)

    def cut(d):
        out = {}
        for k, v in d.items():
            out[k] = v[lo:hi] if (torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == M) else v
        return out

    return cut(rays), cut(tab)


def allreduce_mean_(flat_grad, group=None):
    """DDP semantics: sum over ranks, divide by world size, in place on a flat fp32 buffer."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        flat_grad.div_(dist.get_world_size(group))
    return flat_grad


def flat_grads(param_dicts):
    """One flat fp32 buffer with the gradient of every parameter of the given name -> tensor dicts (dict order), what
    DDP's single bucket holds for this model (8.6 - 12.6 MB): the operand of the step's one all-reduce."""
    gs = [v.grad.reshape(-1) for P in param_dicts for v in P.values() if v.grad is not None]
    return torch.cat(gs) if gs else None


def unflatten_grads_(flat, param_dicts):
    """Write the (reduced) flat buffer back into the parameters' .grad, in place."""
    o = 0
    for P in param_dicts:
        for v in P.values():
            if v.grad is not None:
                n = v.grad.numel()
                v.grad.copy_(flat[o:o + n].view_as(v.grad))
                o += n
    return o
