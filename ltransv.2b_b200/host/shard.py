"""Multi-GPU host logic: one process per GPU, contiguous particle-id slices against
replicated fields.  Particles never interact, so the data path has no
collective; the only exchanges are the 8-counter statistics reduction and the output
gather.  Works with any torch.distributed backend (nccl on GPUs, gloo in CPU tests)."""
import numpy as np


def slice_for_rank(n_total, rank, world_size):
    """Contiguous slice [lo, hi) of 0-based particle indices owned by `rank`;
    ceil(n/world) per rank so that output order = concatenation of ranks."""
    per = -(-n_total // world_size)
    lo = min(rank * per, n_total)
    hi = min(lo + per, n_total)
    return lo, hi


def first_id(lo):
    """Global 1-based id of local particle 0 (keys the Philox stream)."""
    return lo + 1


def allreduce_stats(stats, dist=None, device=None):
    """Sum the 8 int64 counters of ltgpu_stats over ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(stats, dtype=np.int64)
    import torch
    t = torch.as_tensor(np.asarray(stats, dtype=np.int64), device=device)
    dist.all_reduce(t)
    return t.cpu().numpy()


def gather_output(local, n_total, dist=None, device=None):
    """All-gather one per-particle float64/int32 column into global particle order."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local)
    import torch
    ws = dist.get_world_size()
    per = -(-n_total // ws)
    pad = np.zeros(per, dtype=local.dtype)
    pad[:len(local)] = local
    t = torch.as_tensor(pad, device=device)
    outs = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(outs, t)
    return torch.cat(outs).cpu().numpy()[:n_total]
