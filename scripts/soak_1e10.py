import sys, time, json; sys.path.insert(0, '/root/repo')
import numpy as np, scimlexpectations_jl_b200 as sme
with sme.Handle(sme.models.builtin("linear2")) as h:
    h.reduce_generated(1, 10**6)
    t0 = time.perf_counter(); s, s2, c = h.reduce_generated(20261019, 10**10); dt = time.perf_counter() - t0
    print("linear2 MonteCarlo(1e10):", dt, "s", (s / 1e10).tolist(), c, "nf per sample", c["nf"] / 1e10)
    # additivity of the stream: [0, 6e9) + [6e9, 1e10) == [0, 1e10) (counters exactly, sums to rounding)
    a = h.reduce_generated(20261019, 6 * 10**9); b = h.reduce_generated(20261019, 4 * 10**9, first_index=6 * 10**9)
    print("split counters equal:", {k: a[2][k] + b[2][k] for k in c} == c, "sum rel diff", np.max(np.abs(a[0] + b[0] - s) / np.abs(s)))
