"""CPU: host-side logic around the hot path -- cubature driver, sample streams, distributions, the
reference-API mirror's lowering, sharding and the world_size-2 exchange step over gloo."""
import math
import os
import socket
import sys

import numpy as np
import pytest
from scipy import integrate, stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_genz_malik_rule_exact_for_degree7_polynomials(sme):
    from scimlexpectations_jl_b200.cubature import GenzMalikRule, genz_malik_num_points
    assert [genz_malik_num_points(d) for d in (2, 4, 6)] == [17, 57, 149]   # SURVEY.md hard part 5
    rng = np.random.default_rng(0)
    for d in (2, 3, 4):
        rule = GenzMalikRule(d)
        c, h = rng.uniform(-1, 1, (1, d)), rng.uniform(0.2, 1.0, (1, d))
        x = rule.nodes(c, h)
        # monomial x0^3 * x1^2 (total degree 5) and x0^6 * x1 (degree 7): I7 exact; I5 exact only for the first
        def exact(pows):
            out = 1.0
            for j, pw in enumerate(pows):
                a, b = c[0, j] - h[0, j], c[0, j] + h[0, j]
                out *= (b ** (pw + 1) - a ** (pw + 1)) / (pw + 1)
            return out
        for pows in ([3, 2] + [0] * (d - 2), [6, 1] + [0] * (d - 2), [1] * d):
            f = np.prod(x ** np.array(pows), axis=1)[None, :, None]
            val, err, split = rule.apply(f, h)
            assert abs(val[0, 0] - exact(pows)) < 1e-12 * max(1.0, abs(exact(pows)))
        f = (x[:, 0] ** 3 * x[:, 1] ** 2)[None, :, None]
        assert rule.apply(f, h)[1][0, 0] < 1e-12   # degree-5 rule is exact too -> zero error estimate


def test_hcubature_v_against_scipy(sme):
    hc = sme.hcubature_v
    r = hc(lambda x: np.exp(-x[:, 0] ** 2 - 2 * x[:, 1] ** 2), [-1, 0], [2, 1.5], fdim=1, reltol=1e-9)
    ex = integrate.dblquad(lambda y, x: math.exp(-x * x - 2 * y * y), -1, 2, 0, 1.5, epsabs=1e-13)[0]
    assert r.converged and abs(r.u[0] - ex) < 1e-9 and r.resid[0] >= abs(r.u[0] - ex)
    r = hc(lambda x: np.stack([np.cos(3 * x[:, 0]), x[:, 0] ** 2], axis=1), [0], [2], fdim=2, reltol=1e-10)
    assert np.allclose(r.u, [math.sin(6) / 3, 8 / 3], atol=1e-10)
    # every round is ONE integrand call whose size is a multiple of the rule size; maxevals stops it
    sizes = []
    r = hc(lambda x: (sizes.append(len(x)) or np.sin(5 * x.sum(axis=1)) ** 2), [0, 0, 0], [1, 2, 1], fdim=1,
           reltol=1e-7, maxevals=20000)
    assert all(s % 33 == 0 for s in sizes) and r.nrounds == len(sizes) and sum(sizes) == r.nevals
    r2 = hc(lambda x: np.sin(5 * x.sum(axis=1)) ** 2, [0, 0, 0], [1, 2, 1], fdim=1, reltol=1e-12, maxevals=500)
    assert not r2.converged and r2.nevals <= 500 + 2 * 33 * 16
    # max_batch splits a round without changing the result
    r3 = hc(lambda x: np.sin(5 * x.sum(axis=1)) ** 2, [0, 0, 0], [1, 2, 1], fdim=1, reltol=1e-7, maxevals=20000, max_batch=100)
    assert np.array_equal(r3.u, r.u) and max(r3.batch_sizes) <= 99


def test_sample_stream_is_block_keyed_and_distributed_correctly():
    from tests.sampling_util import BLOCK, sample_block, sample_points
    a = sample_points("linear2", 5000, seed=4)
    b = sample_points("linear2", 300, seed=4)
    assert np.array_equal(a[:300], b)                       # prefix property
    assert np.array_equal(sample_block("lorenz63", 1, 3), sample_block("lorenz63", 1, 3))
    assert not np.array_equal(sample_block("linear2", 1, 0, 100), sample_block("linear2", 1, 1, 100))
    x = sample_block("linear2", 2, 0, 200000)
    assert x[:, 0].min() >= 1 and x[:, 0].max() <= 10 and x[:, 1].min() >= 0 and x[:, 1].max() <= 6
    assert stats.kstest(x[:, 0], stats.uniform(1, 9).cdf).pvalue > 1e-3
    assert stats.kstest(x[:, 1], stats.truncnorm(-3, 3, loc=3, scale=1).cdf).pvalue > 1e-3
    assert BLOCK == 1 << 20


def test_generic_distribution_mirror(sme):
    """test/interface.jl:8-40: GenericDistribution vs the product of its factors."""
    d = sme.GenericDistribution(sme.Uniform(1, 10), sme.truncated(sme.Normal(3, 1), 0, 6), sme.Normal(0, 2))
    assert np.array_equal(sme.minimum(d), [1, 0, -math.inf]) and np.array_equal(sme.maximum(d), [10, 6, math.inf])
    lb, ub = sme.extrema(d)
    assert np.array_equal(lb, d.lb) and np.array_equal(ub, d.ub)
    x = [2.0, 4.0, -1.0]
    want = stats.uniform(1, 9).pdf(2.0) * stats.truncnorm(-3, 3, loc=3, scale=1).pdf(4.0) * stats.norm(0, 2).pdf(-1.0)
    assert abs(sme.pdf(d, x) - want) < 1e-15
    assert sme.pdf(d, [0.5, 4.0, 0.0]) == 0.0
    r = sme.rand(d)
    assert r.shape == (3,) and 1 <= r[0] <= 10 and 0 <= r[1] <= 6
    blk = d.rand_block(7, 2, 1000)
    assert np.array_equal(blk, d.rand_block(7, 2, 1000)) and blk.shape == (1000, 3)
    assert d.table()[1] == (3, 0.0, 6.0, 3, 1)
    with pytest.raises(TypeError):
        sme.truncated(sme.Uniform(0, 1), 0.2, 0.8)


def test_api_lowering_to_model_spec(sme):
    """ExpectationProblem(SystemMap(prob, Tsit5(), EnsembleB200(); kwargs...), g, h, d) lowers to the
    C-ABI model exactly as the builtin config does (README.md:31-61)."""
    m = sme.models
    prob = sme.ODEProblem(m._LINEAR2_RHS, [1.0, 1.0], (0.0, 3.0), [1.0, 2.0])
    smap = sme.SystemMap(prob, sme.Tsit5(), sme.EnsembleB200(), save_everystep=False)
    gd = sme.GenericDistribution(sme.Uniform(1, 10), sme.truncated(sme.Normal(3, 1), 0, 6))
    ex = sme.ExpectationProblem(smap, m.observable_components([0, 1]), m.hmap_scatter(u0_from_x=[(0, 0), (1, 1)]), gd, nout=2)
    spec, ref = ex.model_spec(), m.linear2()
    assert (spec.nstate, spec.nparam, spec.nx, spec.nout) == (2, 2, 2, 2)
    assert spec.abstol == 1e-6 and spec.reltol == 1e-3 and [tuple(map(float, t)) for t in spec.pdf] == [tuple(map(float, t)) for t in ref.pdf]
    assert sme.compile_check(spec)[0] > 0
    assert isinstance(ex.parameters(), tuple) and ex.distribution() is gd and ex.mapping() is smap
    # solver kwargs are forwarded (system_utils.jl:42-53 -> expectation.jl:191)
    smap2 = sme.SystemMap(prob, sme.Tsit5(), sme.EnsembleB200(fp32=True), abstol=1e-9, reltol=1e-7, maxiters=77)
    s2 = sme.ExpectationProblem(smap2, ex.g, ex.h, gd, nout=2).model_spec()
    assert (s2.abstol, s2.reltol, s2.maxiters, s2.fp_mode) == (1e-9, 1e-7, 77, 1)
    # process noise: d = n x Truncated(Normal(), -4, 4)  (problem_types.jl:81-85)
    sde = sme.SDEProblem(m._GBM.split("B200E_FN void diffusion")[0], "B200E_FN void diffusion" + m._GBM.split("B200E_FN void diffusion")[1],
                         [1.0, 2.0, 3.0, 4.0], (0.0, 1.0), [])
    pn = sme.ProcessNoiseSystemMap(sde, 8, sme.EM(), dt=1 / 1024)
    exn = sme.ExpectationProblem(pn, m.observable_components([0, 1, 2, 3]), m.hmap_scatter(z_from_x=[(k, k) for k in range(8)]), nout=4)
    sn = exn.model_spec()
    assert sn.solver == 1 and sn.n_kkl == 8 and len(exn.d) == 8 and all(t == (3, -4.0, 4.0, 0.0, 1.0) for t in sn.pdf)
    assert sme.compile_check(sn)[0] > 0
    with pytest.raises(TypeError):
        sme.ExpectationProblem(pn, ex.g, ex.h, gd)
    with pytest.raises(NotImplementedError):
        sme.build_integrand(ex, sme.Koopman(), [5.5, 3.0], ex.params, None)   # unbatched path is not on the device


def test_shard_ranges_partition_exactly(sme):
    from scimlexpectations_jl_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import oracle
    from scimlexpectations_jl_b200.distributed import allreduce_sums, shard_range
    from tests.sampling_util import sample_points
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    x = sample_points("lorenz63", n, seed=21)
    lo, hi = shard_range(n, world, rank)
    # the per-rank compute is the GPU kernel on the box; here the oracle stands in for it (test only)
    s, s2, c = oracle.OracleProblem.builtin("lorenz63").reduce_samples(x[lo:hi], nthreads=2)
    S, S2, Cn = allreduce_sums(s, s2, c)
    if rank == 0:
        q.put((S, S2, Cn))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_exchange_step_over_gloo(sme):
    """N>1 path: contiguous shards, one all-reduce of 2*nout sums + 4 integer counters."""
    import torch.multiprocessing as mp
    import oracle
    from tests.sampling_util import sample_points
    n, world = 30001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    S, S2, Cn = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = sample_points("lorenz63", n, seed=21)
    s, s2, c = oracle.OracleProblem.builtin("lorenz63").reduce_samples(x)
    assert Cn == c                                   # integer bookkeeping is exact across ranks
    assert np.allclose(S, s, rtol=1e-13) and np.allclose(S2, s2, rtol=1e-13)


def test_bench_reference_arm_runs_on_the_cpu_box():
    """`bench.py --impl reference` (the driver's reference arm: the CPU restatement of the path on all host
    cores) needs no GPU and prints the contract's JSON line."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                          "--no-solves"], check=True, capture_output=True, text=True, timeout=600).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "ODE-solve integrand evals/sec" and line["value"] > 1e4
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["config"]["workload"] == "linear2_mc1e7" and line["higher_is_better"] is True


def test_output_buffers_are_never_silently_copied():
    """A strided / float32 / read-only `out` would receive the results in a temporary copy (round-1 advisor finding):
    output buffers must be writeable C-contiguous float64 of the expected size, inputs may still be converted."""
    import scimlexpectations_jl_b200 as sme
    ap = sme._capi._as_pointer
    big = np.zeros((8, 4))
    with pytest.raises(ValueError):
        ap(big[:, 1], 8, output=True)                     # a column view: strided
    with pytest.raises(ValueError):
        ap(np.zeros(8, dtype=np.float32), 8, output=True)
    ro = np.zeros(8)
    ro.flags.writeable = False
    with pytest.raises(ValueError):
        ap(ro, 8, output=True)
    with pytest.raises(ValueError):
        ap(np.zeros(7), 8, output=True)                   # wrong element count
    good = np.zeros((4, 2))
    addr, keep = ap(good, 8, output=True)
    assert addr == good.ctypes.data and keep is good      # in place
    addr, keep = ap(big[:, 1])                            # an input is converted instead
    assert keep.flags.c_contiguous and keep is not big


def test_handle_cache_key_digests_tabulated_factors():
    """Two KDE tables with equal grid ends and equal head / tail values must not share a compiled handle."""
    import scimlexpectations_jl_b200 as sme
    from scimlexpectations_jl_b200.api import _spec_key
    base = sme.models.builtin("lotka_volterra")
    n = 2048
    t1 = np.zeros(n); t2 = np.zeros(n)
    t1[1000] = 1.0; t2[1001] = 1.0                        # differ only where repr() elides
    assert repr(t1) == repr(t2)
    mk = lambda tab: base.replace(pdf=[(sme._capi.PDF_TABLE, 0.0, 1.0, 0.0, 1.0, tab)] + list(base.pdf[1:]))
    assert _spec_key(mk(t1)) != _spec_key(mk(t2))
    assert _spec_key(mk(t1)) == _spec_key(mk(t1.copy()))
    assert _spec_key(base) != _spec_key(base.replace(reltol=1e-4))
