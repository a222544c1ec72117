"""Small pass over every kernel instantiation x schedule (reduce, nodes, sampled reduce, in-library Koopman solve)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scimlexpectations_jl_b200 as sme
from tests.sampling_util import sample_points
for name in ("linear2", "lotka_volterra", "oscillator_noise", "linear2_saveat"):
    for sched in (0, 1, 2):
        if sched == 2 and name == "linear2_saveat":
            continue
        n = 20011 if name != "oscillator_noise" else 3001
        x = sample_points(name, n, seed=1)
        with sme.Handle(sme.models.builtin(name, schedule=sched, block_size=64, blocks_per_sm=1)) as h:
            s, s2, c = h.reduce_samples(x)
            dx = h.eval_nodes(x)
            g = h.reduce_generated(3, n)
            if not len(h.spec.save_times) and h.spec.solver == 0:
                h.koopman_solve(*[np.array(v) for v in zip(*[(d[1], d[2]) for d in h.spec.pdf])], reltol=1e-2, abstol=1e-2)
        print(name, sched, s[:2], c["nf"], g[2]["nf"], flush=True)
print("done")
