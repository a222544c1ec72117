"""reference test/processnoise.jl:9-26 on a B200: Koopman (8-dimensional h-adaptive Genz-Malik in the library)
against MonteCarlo(1e6), both with LambaEM(abstol = reltol = 1e-3); prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scimlexpectations_jl_b200 as sme

m = sme.models
prob = sme.SDEProblem(m._GBM, "", [1.0, 2.0, 3.0, 4.0], (0.0, 1.0))
sm = sme.ProcessNoiseSystemMap(prob, 8, sme.LambaEM(), abstol=1e-3, reltol=1e-3)
ex = sme.ExpectationProblem(sm, m.observable_components([0, 1, 2, 3]), m.hmap_scatter(z_from_x=[(k, k) for k in range(8)]), nout=4)
sme.solve(ex, sme.MonteCarlo(1000))  # compile / load
out = {}
for name, alg, kw in (("koopman", sme.Koopman(), dict(ireltol=1e-3, iabstol=1e-3, batch=64, quadalg=sme.B200Cubature(), maxiters=40_000_000)),
                      ("montecarlo_1e6", sme.MonteCarlo(1_000_000), {})):
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        sol = sme.solve(ex, alg, **kw)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out[name] = {"wall_ms": best * 1e3, "u": np.asarray(sol.u).tolist()}
    if name == "koopman":
        st = sol.original
        out[name].update({k: st[k] for k in ("nevals", "nrounds", "nregions", "max_batch", "converged", "kernel_ms")})
        out[name]["resid"] = np.asarray(sol.resid).tolist()
out["max_rel_diff"] = float(np.max(np.abs(np.array(out["koopman"]["u"]) / np.array(out["montecarlo_1e6"]["u"]) - 1)))
print(json.dumps(out))
ex.close()
