"""Strong scaling of BASELINE config 4 (lorenz63, MonteCarlo(10^9)) and config 2 at 10^9: the TOTAL sample count is
fixed, rank r of R draws and solves indices [r n/R, (r+1) n/R) of the counter-based stream in its kernel, one
all-reduce of the 2 nout + 4 numbers finishes the solve.  Run under torchrun (one process per GPU) or alone (R = 1).
Prints one JSON line per workload: wall time of the whole solve (barrier + synchronize on both sides, max over ranks)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import scimlexpectations_jl_b200 as sme

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

for name, n in (("lorenz63", 10**9), ("linear2", 10**9), ("oscillator_noise", 10**8)):
    lo, hi = n * rank // world, n * (rank + 1) // world
    with sme.Handle(sme.models.builtin(name, devices=[local])) as h:
        nout = h.spec.nout
        red = torch.zeros(2 * nout + 4, dtype=torch.float64, device="cuda")
        def solve():
            s, s2, c = h.reduce_generated(20261019, hi - lo, first_index=lo)
            red.copy_(torch.from_numpy(np.concatenate([s, s2, [c["nf"], c["naccept"], c["nreject"], c["nfail"]]])))
            if world > 1:
                dist.all_reduce(red)
            return red.cpu().numpy()
        h.reduce_generated(1, 10**6); solve() if n <= 10**8 else None   # warm-up
        best = None
        for rep in range(3):
            barrier()
            t0 = time.perf_counter()
            r = solve()
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            best = float(dt.item()) if best is None else min(best, float(dt.item()))
    if rank == 0:
        print(json.dumps({"workload": f"{name} MonteCarlo({n:.0e})", "n_gpus": world, "wall_s": best, "evals_per_s": n / best,
                          "mean": (r[:nout] / n).tolist(), "nf": int(r[2 * nout])}), flush=True)
if world > 1:
    dist.destroy_process_group()
