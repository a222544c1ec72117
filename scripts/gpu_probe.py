"""First-contact GPU probe: parity vs oracle on all builtin models + rough timings."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scimlexpectations_jl_b200 as sme
import oracle
from tests.sampling_util import sample_points

print("devices", sme.device_count())
print("fp64 peak TF", sme.measure_fma_peak(0, False), "fp32 peak TF", sme.measure_fma_peak(0, True))
for name in sme.models.BUILTIN:
    spec = sme.models.builtin(name)
    orc = oracle.OracleProblem.builtin(name)
    n = 20000
    x = sample_points(name, n, seed=1)
    with sme.Handle(spec) as h:
        dx, steps = h.eval_nodes(x, want_steps=True)
        cg = h.last_counters()
        dxo, co, stepso = orc.eval_nodes(x, want_steps=True)
        rel = np.max(np.abs(dx - dxo) / (np.abs(dxo) + 1e-300))
        print(name, "nodes maxrel", rel, "counters", cg, co, "steps equal", bool((steps == stepso).all()))
        s, s2, c = h.reduce_samples(x, weight_by_pdf=False)
        so, s2o, co = orc.reduce_samples(x)
        print(name, "reduce rel", np.abs(s - so) / np.abs(so), np.abs(s2 - s2o) / np.abs(s2o), c == co, h.last_stats())
        # timing with device-resident input
        N = 4_000_000 if spec.solver == 0 else 400_000
        X = sample_points(name, N, seed=2)
        d = h.device_array(X.shape).upload(X)
        for rep in range(3):
            s, s2, c = h.reduce_samples(d, npts=N)
            st = h.last_stats()
            print(name, "N", N, "kernel_ms", st["kernel_ms"], "evals/s", N / st["kernel_ms"] * 1e3, "regs", st["regs_per_thread"], "grid", st["grid"], "occ", st["blocks_per_sm_occ"])
        t0 = time.perf_counter(); s, s2, c = h.reduce_samples(X); t1 = time.perf_counter()
        print(name, "host pageable e2e evals/s", N / (t1 - t0), h.last_stats()["total_ms"])
        d.free()
