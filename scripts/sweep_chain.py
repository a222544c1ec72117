import sys, os, json, numpy as np
sys.path.insert(0, '/root/repo')
import scimlexpectations_jl_b200 as sme
from scimlexpectations_jl_b200.models import hmap_scatter, observable_components
from scimlexpectations_jl_b200._capi import ModelSpec, PDF_UNIFORM
def chain(n, **kw):
    lines = ["    du[0] = -p[0] * u[0];"] + [f"    du[{i}] = p[{i-1}] * u[{i-1}] - p[{i}] * u[{i}];" for i in range(1, n)]
    rhs = "B200E_FN void rhs(real* du, const real* u, const real* p, real t) {\n" + "\n".join(lines) + "\n}\n"
    src = rhs + hmap_scatter(p_from_x=[(i, i) for i in range(n)]) + observable_components(list(range(n)))
    return ModelSpec(source=src, nstate=n, nparam=n, nx=n, nout=n, u0=[1.0] + [0.0]*(n-1), p=[1.0]*n, tspan=(0.0, 2.0),
                     pdf=[(PDF_UNIFORM, 0.5 + 0.05*i, 1.5 + 0.05*i, 0.0, 1.0) for i in range(n)]).replace(**kw)
N = 1000000
for n in (12, 16):
    rng = np.random.default_rng(1)
    X = np.stack([rng.uniform(0.5 + 0.05*i, 1.5 + 0.05*i, N) for i in range(n)], axis=1)
    for blk, per in ((256, 2), (256, 1), (128, 3), (128, 2), (128, 1)):
        with sme.Handle(chain(n, block_size=blk, blocks_per_sm=per)) as h:
            d = h.device_array(X.shape).upload(X)
            best = 1e9
            for rep in range(3):
                s, s2, c = h.reduce_samples(d, npts=N); st = h.last_stats(); best = min(best, st["kernel_ms"])
            d.free()
            print(n, blk, per, "kernel_ms %.3f" % best, "regs", st["regs_per_thread"], "local", st["local_bytes_per_thread"], flush=True)
