"""Which trajectories take a different number of attempted steps on the GPU than in the oracle?"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import scimlexpectations_jl_b200 as gpu
from tests.sampling_util import sample_points
name = sys.argv[1] if len(sys.argv) > 1 else "linear2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-10
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 5
x = sample_points(name, n, seed=seed)
orc = oracle.OracleProblem.builtin(name, abstol=tol, reltol=tol)
dxo, co, so = orc.eval_nodes(x, weight_by_pdf=False, want_steps=True)
with gpu.Handle(gpu.models.builtin(name, abstol=tol, reltol=tol)) as h:
    dx, sg = h.eval_nodes(x, weight_by_pdf=False, want_steps=True)
    print("gpu", h.last_counters(), "oracle", co)
bad = np.nonzero(sg != so)[0]
print("differing trajectories:", len(bad), "of", n, " max |diff|", int(np.max(np.abs(sg - so))))
for i in bad[:5]:
    print(i, repr(x[i].tolist()), "gpu steps", sg[i], "oracle", so[i])
print("max rel node err", np.max(np.abs(dx - dxo) / np.abs(dxo).max(axis=0)))
