"""Wall time of the whole Lotka-Volterra Koopman solve (BASELINE configs[2], bench.py's koopman_lotka_volterra case)
through b200e_koopman_solve for several launch shapes: block:blocks_per_sm ..."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlexpectations_jl_b200 as sme

shapes = [tuple(int(v) for v in s.split(":")) for s in (sys.argv[1] if len(sys.argv) > 1 else "0:0,128:0,64:0,32:0").split(",")]
lb, ub = [1.0, 0.5, 2.0, 0.5], [2.0, 1.5, 4.0, 1.5]
for blk, per in shapes:
    spec = sme.models.builtin("lotka_volterra", abstol=1e-8, reltol=1e-8, block_size=blk, blocks_per_sm=per)
    with sme.Handle(spec) as h:
        best = None
        for rep in range(5):
            t0 = time.perf_counter()
            u, resid, st = h.koopman_solve(lb, ub, reltol=2e-5, abstol=2e-5, maxevals=3_000_000)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        stats = h.last_stats()
        print(json.dumps({"block": blk, "per_sm": per, "wall_ms": round(best * 1e3, 3), "kernel_ms": round(st["kernel_ms"], 3),
                          "nevals": st["nevals"], "rounds": st["nrounds"], "u": float(u[0]), "resid": float(resid[0])}), flush=True)
