"""CPU fuzz of the emulated kernels (tests/cuda_emul) against the oracle on generated models: no GPU needed.
usage: emul_fuzz.py ode|noise|large|logpoly <first seed> <end seed>"""
import os
import sys
import tempfile

import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlexpectations_jl_b200 as sme
from tests.test_kernel_emulation import _build, _run, _check_nodes
from tests.fuzz_util import oracle_of, random_model, random_noise_model, with_logpoly_factors
fam=sys.argv[1]; a=int(sys.argv[2]); b=int(sys.argv[3])
bad=0
for seed in range(a,b):
    try:
        spec, draw = random_noise_model(seed) if fam=="noise" else random_model(seed, nmax=16 if fam=="large" else 5)
        if fam=="logpoly":
            spec, draw = with_logpoly_factors(seed, spec, draw)
            if not any(int(d[0])==5 for d in spec.pdf): continue
        tmp=tempfile.mkdtemp()
        exe=_build(sme, spec, tmp, repro=1 if spec.schedule==2 else 0)
        orc=oracle_of(spec)
        x=draw(200 if fam!="ode" else 400, 9)
        ref,cref,sref=orc.eval_nodes(x,weight_by_pdf=True,want_steps=True)
        dx,steps,_,_,c,_=_run(exe,spec,tmp,x,"nodes_dyn",grid=2)
        flips=int((steps!=sref).sum())
        err=np.abs(dx-ref)/(np.abs(ref).max(axis=0,keepdims=True)+1e-300)
        ok = c["nfail"]==cref["nfail"] and flips==0 and c==cref and err.max()<=2e-6
        print(fam, seed, "ok" if ok else "MISMATCH", "n", spec.nstate, "solver", spec.solver, "nk", spec.n_kkl, "flips", flips, "maxerr %.2e"%err.max(), c, cref if not ok else "", flush=True)
        bad += (not ok)
    except Exception as e:
        bad+=1; print(fam, seed, "EXC", repr(e)[:300], flush=True)
print("bad", bad)
