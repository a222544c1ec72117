"""Workload for compute-sanitizer (one tool per gpurun call: memcheck | racecheck | synccheck | initcheck).

Small on purpose -- a kernel under the sanitizer runs 10-100x slower -- but it walks every synchronisation pattern
of the path once:
  * the static instantiations (batch smaller than the grid) and the dynamic ones (atomic work counter, more points
    than lanes) of the Tsit5 kernel in nodes and reduce mode: per-lane shared-memory accumulators, the shuffle /
    shared-memory block reduction, the last-block fold (__threadfence + ticket + __ldcg);
  * a chunked host-pointer reduce_samples with more chunks than ring slots (4-slot ring, copy stream, two compute
    streams, cross-stream events) from pageable memory, and the same in nodes mode (copy-back stream);
  * two pending reductions: begin / begin / end / end (second reduction scratch, one stream);
  * the reproducible dynamic schedule (super-chunk rows + two fold kernels);
  * the in-kernel sampler (b200e_reduce_gen), Euler-Maruyama (three samples per thread) and LambaEM;
  * b200e_eval_regions and b200e_koopman_solve (rule nodes -> model kernel -> Genz-Malik apply, region heap).
Every result is also compared with the oracle, so a sanitizer-clean run is a correct one.

    compute-sanitizer --tool racecheck python scripts/sanitize_run.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
import scimlexpectations_jl_b200 as sme
from tests.sampling_util import sample_points

NTH = max(1, len(os.sched_getaffinity(0)))


def close(a, b, tol=1e-9):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / (np.abs(b).max() + 1e-300))) <= tol


def main():
    scale = float(os.environ.get("SANITIZE_SCALE", "1"))
    name = "linear2"
    orc = oracle.OracleProblem.builtin(name)
    spec = sme.models.builtin(name, devices=[0], chunk_points=8192)
    with sme.Handle(spec) as h:
        x = sample_points(name, 3000, seed=1)
        dx = h.eval_nodes(x, weight_by_pdf=True)                       # static nodes
        s, s2, c = h.reduce_samples(x)                                 # static reduce + last-block fold
        ro, co = orc.eval_nodes(x, weight_by_pdf=True)
        so, _, cs = orc.reduce_samples(x)
        assert c == cs and close(dx, ro, 2e-6) and close(s, so)
        lanes = h.last_stats()["grid"] * h.last_stats()["block"]
        n = int(lanes * 1.3 * scale) + 77
        xb = sample_points(name, n, seed=2)
        so, _, cs = orc.reduce_samples(xb, nthreads=NTH)
        # chunked, pageable host memory: n / 8192 chunks through the 4-slot ring, several launches + finalize kernels
        s, s2, c = h.reduce_samples(xb)
        assert c == cs and close(s, so), ("chunked", c, cs)
        dx = h.eval_nodes(xb[:50_000], weight_by_pdf=True)            # chunked nodes mode: copy-back stream
        ro, _ = orc.eval_nodes(xb[:50_000], weight_by_pdf=True, nthreads=NTH)
        assert close(dx, ro, 2e-6)
        # device-resident: one dynamic launch (atomic work counter), then two pending reductions
        xd = h.device_array(xb.shape).upload(xb)
        s, s2, c = h.reduce_samples(xd, npts=n)
        assert c == cs and close(s, so), ("resident", c, cs)
        dxd = h.device_array((n, 2))
        h.eval_nodes(xd, npts=n, out=dxd, weight_by_pdf=False)         # nodes_dyn, device in / device out
        h.reduce_samples_begin(xd, npts=n)
        h.reduce_samples_begin(xd, npts=n)
        r1 = h.reduce_samples_end()
        r2 = h.reduce_samples_end()
        assert r1[2] == cs and r2[2] == cs and close(r1[0], so) and close(r2[0], so)
        # the sampler in the kernel
        sg, _, cg = h.reduce_generated(5, n, first_index=17)
        xg = sme.sample_host(spec, 5, n, first_index=17)
        sgo, _, cgo = orc.reduce_samples(xg, nthreads=NTH)
        assert cg == cgo and close(sg, sgo)
        # the cubature rule on the device and the whole Koopman solve in the library
        u, resid, st = h.koopman_solve([1.0, 0.0], [10.0, 6.0], reltol=1e-3, abstol=1e-3)
        assert st["nevals"] == 119 and abs(u[0] - 1.5433860531082695) < 2e-9
        cen = np.array([[5.5, 3.0], [3.25, 1.5]])
        hw = np.array([[4.5, 3.0], [2.25, 1.5]])
        val, err, split = h.eval_regions(cen, hw)
        assert np.all(np.isfinite(val)) and np.all(err >= 0)
        xd.free()
        dxd.free()
    print("linear2: static / dynamic / chunked / pending / sampled / cubature ok", flush=True)

    # reproducible dynamic schedule: super-chunk rows + fold kernels
    with sme.Handle(sme.models.builtin(name, devices=[0], schedule=sme._capi.SCHED_DYNAMIC_REPRODUCIBLE)) as h:
        xd = h.device_array(xb.shape).upload(xb)
        a = h.reduce_samples(xd, npts=n)
        b = h.reduce_samples(xd, npts=n)
        assert np.array_equal(a[0], b[0]) and a[2] == cs and close(a[0], so)
        xd.free()
    print("reproducible schedule ok", flush=True)

    # Lotka-Volterra (divergent step counts, partial refills), Euler-Maruyama, LambaEM
    for name2, lamba, m in (("lotka_volterra", False, int(60_000 * scale)), ("oscillator_noise", False, int(20_000 * scale)),
                            ("gbm4", True, int(20_000 * scale))):
        x2 = sample_points(name2, m, seed=3)
        orc2 = oracle.OracleProblem.builtin(name2, **({"solver": oracle.SOLVER_LAMBA_EM, "abstol": 1e-3, "reltol": 1e-3} if lamba else {}))
        with sme.Handle(sme.models.builtin("gbm4_lamba" if lamba else name2, devices=[0])) as h:
            xd = h.device_array(x2.shape).upload(x2)
            s, s2, c = h.reduce_samples(xd, npts=m)
            dx = h.eval_nodes(x2[:5000], weight_by_pdf=True)
            xd.free()
        so2, _, cs2 = orc2.reduce_samples(x2, nthreads=NTH)
        ro2, _ = orc2.eval_nodes(x2[:5000], weight_by_pdf=True, nthreads=NTH)
        assert c == cs2 and close(s, so2, 1e-8) and close(dx, ro2, 2e-6), (name2, c, cs2)
        print(f"{name2} ok", flush=True)
    print("SANITIZE_RUN_OK")


if __name__ == "__main__":
    main()
