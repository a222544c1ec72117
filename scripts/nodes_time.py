"""Kernel time of b200e_nodes (per-node outputs, device-resident in and out) for a few launch shapes.
usage: nodes_time.py <model> <npts> <block:per_sm,...>"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scimlexpectations_jl_b200 as sme
from tests.sampling_util import sample_points

name = sys.argv[1]
N = int(float(sys.argv[2]))
configs = [tuple(int(v) for v in c.split(":")) for c in sys.argv[3].split(",")]
X = sample_points(name, N, seed=2)
ref = None
for blk, per in configs:
    spec = sme.models.builtin(name, block_size=blk, blocks_per_sm=per)
    with sme.Handle(spec) as h:
        d = h.device_array(X.shape).upload(X)
        out = h.device_array((N, spec.nout))
        best = 1e9
        for rep in range(4):
            h.eval_nodes(d, npts=N, out=out, weight_by_pdf=True)
            st = h.last_stats()
            best = min(best, st["kernel_ms"])
        got = out.download()
        if ref is None:
            ref = got
        print(json.dumps({"model": name, "mode": "nodes", "block": blk, "per_sm": per, "kernel_ms": round(best, 4),
                          "evals_per_s": round(N / best * 1e3 / 1e9, 4), "regs": st["regs_per_thread"], "grid": st["grid"],
                          "same": bool(np.array_equal(got, ref))}), flush=True)
        d.free(); out.free()
