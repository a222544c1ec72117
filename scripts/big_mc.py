"""BASELINE configs 4 and 5 at full size on ONE GPU through the public MonteCarlo call with the in-kernel sampler:
MonteCarlo(10^9) of lorenz63 and MonteCarlo(10^8) of the Euler-Maruyama oscillator (no sample set exists anywhere)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scimlexpectations_jl_b200 as sme

out = {}
for name, n in (("lorenz63", 10**9), ("oscillator_noise", 10**8), ("linear2", 10**9)):
    with sme.Handle(sme.models.builtin(name)) as h:
        h.reduce_generated(1, 10**6)                      # warm-up (module load, buffers)
        t0 = time.perf_counter()
        s, s2, c = h.reduce_generated(20261019, n)
        dt = time.perf_counter() - t0
        st = h.last_stats()
    mean = s / n
    sem = np.sqrt(np.maximum(s2 / n - mean ** 2, 0.0) / n)
    out[name] = {"n": n, "wall_s": dt, "kernel_ms": st["kernel_ms"], "evals_per_s": n / dt, "mean": mean.tolist(),
                 "standard_error": sem.tolist(), "counters": c}
    print(name, json.dumps(out[name]), flush=True)
json.dump(out, open(os.path.join("gpurun_out", "big_mc.json"), "w"), indent=1)
