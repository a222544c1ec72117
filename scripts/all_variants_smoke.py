"""Small pass over every kernel instantiation x schedule x precision (reduce, nodes, sampled reduce, in-library
Koopman solve), on one device and -- when the box has several -- through the in-library multi-device path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scimlexpectations_jl_b200 as sme
from tests.sampling_util import sample_points
nd = sme.device_count()
for devices in ([0],) + (([0, 1],) if nd >= 2 else ()):
    for name in ("linear2", "lotka_volterra", "oscillator_noise", "linear2_saveat", "gbm4_lamba"):
        for fp in (0, 1):
            ref = None
            for sched in (0, 1, 2):
                if sched == 2 and name == "linear2_saveat":
                    continue
                n = 20011 if name not in ("oscillator_noise", "gbm4_lamba") else 3001
                x = sample_points(name, n, seed=1)
                with sme.Handle(sme.models.builtin(name, schedule=sched, fp_mode=fp, block_size=64, blocks_per_sm=1, devices=devices,
                                                   chunk_points=7000)) as h:
                    s, s2, c = h.reduce_samples(x)
                    dx = h.eval_nodes(x)
                    g = h.reduce_generated(3, n)
                    gs = h.reduce_generated(3, 500)
                    if not len(h.spec.save_times) and h.spec.solver == 0 and len(devices) == 1:
                        h.koopman_solve(*[np.array(v) for v in zip(*[(d[1], d[2]) for d in h.spec.pdf])], reltol=1e-2, abstol=1e-2)
                if ref is None:
                    ref = (s, c, g, dx)
                assert c == ref[1] and g[2] == ref[2][2], (name, fp, sched, devices)
                assert np.allclose(s, ref[0], rtol=1e-10) and np.allclose(g[0], ref[2][0], rtol=1e-10) and np.array_equal(dx, ref[3])
                print(len(devices), name, "fp32" if fp else "fp64", sched, s[:2], c["nf"], g[2]["nf"], flush=True)
print("done")
