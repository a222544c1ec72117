"""Latency of ONE b200e_eval_nodes call (the f!(dx, x, p) of a cubature refinement round) against batch size, host
(pageable) in / host out -- what Integrals / Cubature see per round -- and device in / device out for comparison."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scimlexpectations_jl_b200 as sme
from tests.sampling_util import sample_points

for name, kw in (("linear2", {}), ("lotka_volterra", dict(abstol=1e-8, reltol=1e-8))):
    with sme.Handle(sme.models.builtin(name, **kw)) as h:
        for n in (17, 57, 570, 5700, 57000):
            x = sample_points(name, n, seed=1)
            out = np.empty((n, h.spec.nout))
            for _ in range(5):
                h.eval_nodes(x, out=out)
            reps = 200 if n <= 5700 else 50
            t0 = time.perf_counter()
            for _ in range(reps):
                h.eval_nodes(x, out=out)
            host_us = (time.perf_counter() - t0) / reps * 1e6
            st = h.last_stats()
            xd = h.device_array(x.shape).upload(x)
            od = h.device_array(out.shape)
            for _ in range(5):
                h.eval_nodes(xd, npts=n, out=od)
            t0 = time.perf_counter()
            for _ in range(reps):
                h.eval_nodes(xd, npts=n, out=od)
            dev_us = (time.perf_counter() - t0) / reps * 1e6
            st2 = h.last_stats()
            xd.free(); od.free()
            print(json.dumps({"model": name, "npts": n, "host_call_us": round(host_us, 1), "host_kernel_us": round(st["kernel_ms"] * 1e3, 1),
                              "device_call_us": round(dev_us, 1), "device_kernel_us": round(st2["kernel_ms"] * 1e3, 1)}), flush=True)
