"""e2e (host pinned / pageable samples -> host sums) of the default workload against the chunk size of the host pipeline."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scimlexpectations_jl_b200 as sme
from tests.sampling_util import sample_points

n = 10_000_000
x = sample_points("linear2", n, seed=2)
for chunk in (1 << 21, 1 << 20, 1 << 19, 1 << 18, 1 << 17):
    with sme.Handle(sme.models.builtin("linear2", chunk_points=chunk)) as h:
        xp = h.pinned(x.shape)
        xp.array[...] = x
        res = {}
        for name, buf in (("pinned", xp), ("pageable", x)):
            for _ in range(3):
                h.reduce_samples(buf)
            t0 = time.perf_counter()
            for _ in range(8):
                s, s2, c = h.reduce_samples(buf)
            res[name] = (time.perf_counter() - t0) / 8 * 1e3
        xp.free()
        print(json.dumps({"chunk_points": chunk, "pinned_ms": round(res["pinned"], 3), "pageable_ms": round(res["pageable"], 3), "launches": h.last_stats()["launches"]}), flush=True)
