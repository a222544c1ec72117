"""Launch-shape sweep on a B200: kernel time of b200e_reduce with device-resident input."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scimlexpectations_jl_b200 as sme
from tests.sampling_util import sample_points

models = sys.argv[1].split(",") if len(sys.argv) > 1 else ["linear2"]
N = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
configs = [(128, 4, 32), (128, 5, 32), (128, 6, 32), (128, 8, 32), (64, 10, 32), (64, 8, 32), (256, 2, 32), (96, 6, 32),
           (128, 5, 16), (128, 5, 24), (128, 5, 8), (128, 4, 16), (128, 4, 24)]
if len(sys.argv) > 3:
    configs = [tuple(int(v) for v in c.split(":")) for c in sys.argv[3].split(",")]
for name in models:
    X = sample_points(name, N, seed=2)
    ref = None
    for cfg in configs:
        blk, per, refill = cfg[:3]
        sched = cfg[3] if len(cfg) > 3 else 0
        spec = sme.models.builtin(name, block_size=blk, blocks_per_sm=per, refill_threshold=refill, schedule=sched)
        try:
            with sme.Handle(spec) as h:
                d = h.device_array(X.shape).upload(X)
                best = 1e9
                for rep in range(4):
                    s, s2, c = h.reduce_samples(d, npts=N)
                    st = h.last_stats()
                    best = min(best, st["kernel_ms"])
                d.free()
                if ref is None:
                    ref = (s, c)
                ok = c == ref[1] and np.allclose(s, ref[0], rtol=1e-11)
                print(json.dumps({"model": name, "block": blk, "per_sm": per, "refill": refill, "sched": sched, "kernel_ms": round(best, 4),
                                  "evals_per_s": round(N / best * 1e3 / 1e9, 3), "regs": st["regs_per_thread"], "grid": st["grid"],
                                  "occ": st["blocks_per_sm_occ"], "local": st["local_bytes_per_thread"], "same": bool(ok)}), flush=True)
        except Exception as e:
            print(name, blk, per, refill, "ERR", e, flush=True)
