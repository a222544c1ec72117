"""Sharding of sample points across ranks and the single exchange step of the path.

One process per GPU (``torch.distributed``; NCCL on the GPU box, gloo in the CPU tests).  Trajectories
are independent, so a rank owns a contiguous index range (or a contiguous range of sample blocks) and
nothing crosses GPUs until the end, when the 2*nout partial sums and the 4 integer counters are
all-reduced (reference: the ``mean(sol.u)`` of src/expectation.jl:293 over all trajectories)."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, world: int, rank: int):
    """Contiguous index range [lo, hi) of rank ``rank``; sizes differ by at most one.  Identical to
    ``shard_range`` in csrc/b200_expect.cu (the single-process multi-device split)."""
    q, r = divmod(int(n), int(world))
    lo = q * rank + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def block_range(nblocks: int, world: int, rank: int):
    """Blocks of the seeded sample stream owned by ``rank`` (block b -> pure function of (seed, b))."""
    return shard_range(nblocks, world, rank)


def allreduce_sums(sums: np.ndarray, sumsq: np.ndarray, counters: dict, *, device=None, group=None):
    """All-reduce (sum) of the per-rank partial results.  Counters travel as int64, so they stay
    exact; sums as float64.  Returns (sums, sumsq, counters) of the whole job."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return sums, sumsq, counters
    nout = len(sums)
    f = torch.empty(2 * nout, dtype=torch.float64, device=device)
    f[:nout] = torch.from_numpy(np.asarray(sums, dtype=np.float64))
    f[nout:] = torch.from_numpy(np.asarray(sumsq, dtype=np.float64))
    c = torch.tensor([counters["nf"], counters["naccept"], counters["nreject"], counters["nfail"]],
                     dtype=torch.int64, device=device)
    dist.all_reduce(f, group=group)
    dist.all_reduce(c, group=group)
    f = f.cpu().numpy()
    c = c.cpu().numpy()
    return f[:nout].copy(), f[nout:].copy(), {"nf": int(c[0]), "naccept": int(c[1]), "nreject": int(c[2]),
                                              "nfail": int(c[3])}
