"""CPU: the kernel SOURCE the B200 runs (csrc/kernels/expect_kernels.cuh with expect_math.cuh and expect_sampling.cuh),
compiled for the host against tests/cuda_emul/cuda_emul.h -- one OS thread per CUDA thread, warp intrinsics as
32-party collectives, blocks one after the other -- and checked against the oracle.

What this adds to the "not gpu" suite: the logic of the kernels themselves (start of a trajectory, Tsit5 / LambaEM /
Euler-Maruyama step functions, refill phases, static and dynamic work distribution, super-chunk rows, the in-kernel
sampler, block reduction and last-block fold) is exercised on every CPU run, not only on the GPU box.  It is test
infrastructure: the product has no CPU path, and timing / occupancy / the PTX-level pieces (MUFU seeds, inline asm)
are outside of what an emulation can say -- the `-m gpu` tests remain the parity tests proper.
"""
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle
from tests.sampling_util import oracle_for, sample_points

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "scimlexpectations.jl_b200", "csrc")
EMUL = os.path.join(ROOT, "tests", "cuda_emul")
BLOCK = 64


def _build(sme, spec, tmp, *, repro=0, em_spt=2, fp32=0):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    model = os.path.join(tmp, "model.inc")
    open(model, "w").write(spec.source)
    exe = os.path.join(tmp, "driver")
    clip = 1 if (spec.dtmax > 0 and spec.dtmax < spec.tspan[1] - spec.tspan[0]) else 0
    defs = dict(B200E_NSTATE=spec.nstate, B200E_NPARAM=spec.nparam, B200E_NX=spec.nx, B200E_NOUT=spec.nout,
                B200E_NKKL=max(1, spec.n_kkl), B200E_SOLVER=spec.solver, B200E_FP32=fp32, B200E_BLOCK=BLOCK, B200E_MINBLOCKS=1,
                B200E_DYNAMIC=1, B200E_REPRO=repro, B200E_NSAVE=len(spec.save_times), B200E_TABLE_PDF=0,
                B200E_LOGPOLY_PDF=int(any(int(d[0]) == 5 for d in spec.pdf)),
                B200E_CLIP_DTMAX=clip, B200E_EM_SPT=em_spt)
    cmd = [gxx, "-std=c++20", "-O1", "-pthread", "-Wno-unknown-pragmas", "-I", EMUL, "-I", CSRC, "-I", os.path.join(CSRC, "kernels")]
    cmd += [f"-D{k}={v}" for k, v in defs.items()] + [f'-DMODEL_FILE="{model}"', os.path.join(EMUL, "driver.cpp"), "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def _run(exe, spec, tmp, x, mode, *, grid=3, weight=1, layout=0, refill=0, gen=(0, 0), n=None):
    n = len(x) if n is None else n
    xfile, out = os.path.join(tmp, "x.bin"), os.path.join(tmp, "out.bin")
    xin = x if layout == 0 else np.ascontiguousarray(x.T)
    np.ascontiguousarray(xin, dtype=np.float64).tofile(xfile)
    lines = [f"npts {n}", f"grid {grid}", f"layout {layout}", f"weight {weight}", f"tspan {spec.tspan[0]!r} {spec.tspan[1]!r}",
             f"tol {spec.abstol!r} {spec.reltol!r}", f"dtmax {spec.dtmax!r}", f"dt_fixed {spec.dt_fixed!r}",
             f"maxiters {spec.maxiters}", f"refill {refill}", "u0 " + " ".join(repr(float(v)) for v in spec.u0),
             "p " + " ".join(repr(float(v)) for v in spec.p), f"gen {gen[0]} {gen[1]}"]
    lines += ["pdf " + " ".join(repr(float(v)) for v in (list(d[:5]) + (list(d[5]) if int(d[0]) == 5 else []))) for d in spec.pdf]
    lines += ["save " + " ".join(repr(float(t)) for t in spec.save_times)] if len(spec.save_times) else []
    lines += [f"mode {mode}", f"xfile {xfile}", f"outfile {out}"]
    cfg = os.path.join(tmp, "cfg.txt")
    open(cfg, "w").write("\n".join(lines) + "\n")
    subprocess.run([exe, cfg], check=True, timeout=300)
    raw = open(out, "rb").read()
    nout = spec.nout
    dx = np.frombuffer(raw[:n * nout * 8]).reshape((n, nout) if layout == 0 else (nout, n))
    off = n * nout * 8
    steps = np.frombuffer(raw[off:off + 4 * n], dtype=np.int32)
    off += 4 * n
    sums = np.frombuffer(raw[off:off + 16 * nout])
    off += 16 * nout
    cnt = np.frombuffer(raw[off:off + 32], dtype=np.uint64)
    work = np.frombuffer(raw[off + 32:off + 48], dtype=np.uint64)
    c = dict(zip(("nf", "naccept", "nreject", "nfail"), (int(v) for v in cnt)))
    return (dx if layout == 0 else dx.T), steps, sums[:nout], sums[nout:], c, work


def _check_nodes(dx, steps, c, ref, cref, sref, tol=1e-10):
    assert c == cref and np.array_equal(steps, sref)
    err = np.abs(dx - ref) / (np.abs(ref).max(axis=0, keepdims=True) + 1e-300)
    assert np.median(err) <= 1e-12 and err.max() <= tol, (float(np.median(err)), float(err.max()))


def _check_sums(s, s2, c, so, s2o, co):
    assert c == co
    assert np.max(np.abs(s - so) / np.abs(so)) < 1e-8 and np.max(np.abs(s2 - s2o) / np.abs(s2o)) < 1e-8   # north_star's bar


@pytest.mark.parametrize("name", ["linear2", "lorenz63"])
def test_tsit5_kernels_on_the_cpu(sme, tmp_path, name):
    """All five instantiations of the adaptive Tsit5 kernel: per-node outputs, per-point steps and counters identical
    to the oracle under static ownership and dynamic distribution, in both layouts and with an explicit refill
    threshold; sums from the block reduction + last-block fold; the launch's counters re-armed to zero."""
    spec = sme.models.builtin(name)
    tmp = str(tmp_path)
    exe = _build(sme, spec, tmp)
    n = 700
    x = sample_points(name, n, seed=3)
    orc = oracle.OracleProblem.builtin(name)
    ref, cref, sref = orc.eval_nodes(x, weight_by_pdf=True, want_steps=True)
    for mode, kw in (("nodes", {}), ("nodes_dyn", {}), ("nodes", dict(layout=1, refill=8)), ("nodes_dyn", dict(grid=1, refill=32))):
        dx, steps, _, _, c, work = _run(exe, spec, tmp, x, mode, **kw)
        _check_nodes(dx, steps, c, ref, cref, sref, tol=2e-6)
        assert not work.any()
    so, s2o, co = orc.reduce_samples(x)
    for mode in ("reduce", "reduce_dyn"):
        _, _, s, s2, c, work = _run(exe, spec, tmp, x, mode, weight=0)
        _check_sums(s, s2, c, so, s2o, co)
        assert not work.any()
    # ragged and tiny launches: fewer points than lanes, one point
    for m in (1, 37):
        dx, steps, _, _, c, _ = _run(exe, spec, tmp, x[:m], "nodes_dyn")
        assert np.array_equal(steps, sref[:m]) and np.max(np.abs(dx - ref[:m]) / np.abs(ref).max(axis=0)) <= 2e-6
    # the in-kernel sampler: same sums as the oracle on the host twin's samples
    xg = sme.sample_host(spec, seed=5, first_index=11, n=500)
    sg, s2g, cg = orc.reduce_samples(xg)
    _, _, s, s2, c, _ = _run(exe, spec, tmp, xg, "gen", weight=0, gen=(5, 11))
    _check_sums(s, s2, c, sg, s2g, cg)


def test_reproducible_rows_on_the_cpu(sme, tmp_path):
    """B200E_SCHED_DYNAMIC_REPRODUCIBLE: one row of sums per super-chunk of 384 consecutive points, whatever the grid."""
    spec = sme.models.builtin("linear2", schedule=2)
    tmp = str(tmp_path)
    exe = _build(sme, spec, tmp, repro=1)
    x = sample_points("linear2", 1000, seed=4)
    so, s2o, co = oracle.OracleProblem.builtin("linear2").reduce_samples(x)
    got = []
    for grid in (1, 3):
        _, _, s, s2, c, _ = _run(exe, spec, tmp, x, "reduce_dyn", weight=0, grid=grid)
        _check_sums(s, s2, c, so, s2o, co)
        got.append((s.copy(), s2.copy()))
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])   # bit-identical across launch shapes


def test_lamba_em_kernel_on_the_cpu(sme, tmp_path):
    """K8: the adaptive Euler-Maruyama step function under the same driver."""
    spec = sme.models.builtin("gbm4_lamba")
    tmp = str(tmp_path)
    exe = _build(sme, spec, tmp)
    x = sample_points("gbm4", 200, seed=11)
    orc = oracle.OracleProblem.builtin("gbm4", solver=oracle.SOLVER_LAMBA_EM, abstol=1e-3, reltol=1e-3)
    ref, cref, sref = orc.eval_nodes(x, weight_by_pdf=True, want_steps=True)
    dx, steps, _, _, c, _ = _run(exe, spec, tmp, x, "nodes_dyn", grid=2)
    _check_nodes(dx, steps, c, ref, cref, sref, tol=1e-8)
    so, s2o, co = orc.reduce_samples(x)
    _, _, s, s2, c, _ = _run(exe, spec, tmp, x, "reduce_dyn", weight=0, grid=2)
    _check_sums(s, s2, c, so, s2o, co)


@pytest.mark.parametrize("spt", [1, 2, 3])
def test_euler_maruyama_kernel_on_the_cpu(sme, tmp_path, spt):
    """K3 with one, two and three samples per thread (pools of 32 x SPT points, ragged last pool)."""
    spec = sme.models.builtin("oscillator_noise", dt_fixed=1.0 / 64.0)
    tmp = str(tmp_path)
    exe = _build(sme, spec, tmp, em_spt=spt)
    x = sample_points("oscillator_noise", 333, seed=12)
    orc = oracle.OracleProblem.builtin("oscillator_noise", dt_fixed=1.0 / 64.0)
    ref, cref, sref = orc.eval_nodes(x, weight_by_pdf=True, want_steps=True)
    for mode in ("nodes", "nodes_dyn"):
        dx, steps, _, _, c, _ = _run(exe, spec, tmp, x, mode, grid=2)
        _check_nodes(dx, steps, c, ref, cref, sref, tol=1e-10)
    so, s2o, co = orc.reduce_samples(x)
    _, _, s, s2, c, _ = _run(exe, spec, tmp, x, "reduce_dyn", weight=0, grid=2)
    _check_sums(s, s2, c, so, s2o, co)


def test_time_sampled_observable_kernel_on_the_cpu(sme, tmp_path):
    """K6: save slots emitted from inside the accept branch (Tsit5's free interpolant)."""
    spec = sme.models.builtin("linear2_saveat")
    tmp = str(tmp_path)
    exe = _build(sme, spec, tmp)
    x = sample_points("linear2_saveat", 300, seed=13)
    orc = oracle_for("linear2_saveat")
    ref, cref, sref = orc.eval_nodes(x, weight_by_pdf=True, want_steps=True)
    dx, steps, _, _, c, _ = _run(exe, spec, tmp, x, "nodes_dyn", grid=2)
    _check_nodes(dx, steps, c, ref, cref, sref, tol=1e-8)
    so, s2o, co = orc.reduce_samples(x)
    _, _, s, s2, c, _ = _run(exe, spec, tmp, x, "reduce", weight=0, grid=2)
    _check_sums(s, s2, c, so, s2o, co)
    # the reproducible dynamic schedule with time-sampled observables: the lane's slot values of a chunk are summed
    # over the lanes in a fixed order at the end of the chunk; bit-identical rows for one and three blocks
    exe_r = _build(sme, spec.replace(schedule=2), tmp, repro=1)
    got = []
    for grid in (1, 3):
        _, _, s, s2, c, _ = _run(exe_r, spec, tmp, x, "reduce_dyn", weight=0, grid=grid)
        _check_sums(s, s2, c, so, s2o, co)
        got.append((s.copy(), s2.copy()))
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])


@pytest.mark.parametrize("family,seed", [("ode", 1), ("ode", 6), ("ode", 16), ("noise", 0), ("noise", 4), ("noise", 7)])
def test_generated_models_on_the_cpu(sme, tmp_path, family, seed):
    """The generated models of tests/fuzz_util.py (the same text goes to NVRTC on the GPU box) through the emulated
    kernels: random states / h / densities / observables / tolerances, process-noise models with odd and even numbers
    of KKL modes under both SDE solvers."""
    from tests.fuzz_util import oracle_of, random_model, random_noise_model
    spec, draw = random_noise_model(seed) if family == "noise" else random_model(seed)
    tmp = str(tmp_path)
    exe = _build(sme, spec, tmp, repro=1 if spec.schedule == 2 else 0)
    orc = oracle_of(spec)
    x = draw(150 if family == "noise" else 400, 3)
    ref, cref, sref = orc.eval_nodes(x, weight_by_pdf=True, want_steps=True)
    dx, steps, _, _, c, _ = _run(exe, spec, tmp, x, "nodes_dyn", grid=2)
    _check_nodes(dx, steps, c, ref, cref, sref, tol=2e-6)
    so, s2o, co = orc.reduce_samples(x, weight_by_pdf=True)
    _, _, s, s2, c, _ = _run(exe, spec, tmp, x, "reduce_dyn" if len(spec.save_times) == 0 else "reduce", grid=2)
    assert c == co
    bound = 1e-8 * np.sqrt(len(x) * s2o) + 1e-300
    assert np.all(np.abs(s - so) <= bound) and np.all(np.abs(s2 - s2o) <= 1e-8 * s2o + 1e-300)


@pytest.mark.parametrize("name", ["linear2", "gbm4_lamba"])
def test_fp32_mode_kernels_on_the_cpu(sme, tmp_path, name):
    """fp_mode = B200E_FP32 (docs/src/tutorials/introduction.md:419-433): `real = float` state arithmetic, FP64 sums.
    Against the FP64 oracle the values agree to single precision and a trajectory takes at most one step more or less."""
    spec = sme.models.builtin(name, fp_mode=1)
    tmp = str(tmp_path)
    exe = _build(sme, spec, tmp, fp32=1)
    x = sample_points("linear2" if name == "linear2" else "gbm4", 300, seed=5)
    orc = oracle_for(name) if name != "linear2" else oracle.OracleProblem.builtin(name)
    ref, cref, sref = orc.eval_nodes(x, weight_by_pdf=True, want_steps=True)
    dx, steps, _, _, c, _ = _run(exe, spec, tmp, x, "nodes_dyn", grid=2)
    assert c["nfail"] == 0 and np.abs(steps - sref).max() <= (1 if name == "linear2" else 3)
    assert np.max(np.abs(dx - ref) / np.abs(ref).max(axis=0)) <= (1e-4 if name == "linear2" else 2e-3)


def test_wide_observable_register_accumulators_on_the_cpu(sme, tmp_path):
    """64 outputs (32 save times x 2 components): at this block size the per-lane accumulators do not fit the
    shared-memory columns and fall back to the Accum register arrays; block reduction one column at a time."""
    times = tuple(np.linspace(0.0, 3.0, 32).tolist())
    spec = sme.models.builtin("linear2_saveat", save_times=times, nout=64)
    assert (2 * spec.nout * 8 + 32) * BLOCK > 40960
    tmp = str(tmp_path)
    exe = _build(sme, spec, tmp)
    x = sample_points("linear2_saveat", 150, seed=14)
    orc = oracle.OracleProblem.from_source(spec.source, nstate=spec.nstate, nparam=spec.nparam, nx=spec.nx, nout=spec.nout,
                                           u0=spec.u0, p=spec.p, tspan=spec.tspan, pdf=spec.pdf, save_times=spec.save_times)
    ref, cref, sref = orc.eval_nodes(x, weight_by_pdf=True, want_steps=True)
    dx, steps, _, _, c, _ = _run(exe, spec, tmp, x, "nodes", grid=2)
    _check_nodes(dx, steps, c, ref, cref, sref, tol=1e-8)
    so, s2o, co = orc.reduce_samples(x)
    for mode in ("reduce", "reduce_dyn"):
        _, _, s, s2, c, _ = _run(exe, spec, tmp, x, mode, weight=0, grid=2)
        assert c == co and np.allclose(s, so, rtol=1e-9, atol=1e-12) and np.allclose(s2, s2o, rtol=1e-9)


def test_logpoly_density_factors_on_the_cpu(sme, tmp_path):
    """B200E_PDF_LOGPOLY in the kernel source: Gamma / LogNormal / Exponential / Beta-shaped factors, including
    points on and outside the support's ends."""
    pdf = [sme.truncated(sme.Gamma(9.0, 1.5 / 9.0), 1.0, 2.0).table(), sme.truncated(sme.LogNormal(0.0, 0.4), 0.5, 1.5).table(),
           sme.truncated(sme.Exponential(3.0), 2.0, 4.0).table(),
           (sme._capi.PDF_LOGPOLY, 0.0, 1.5, 0.0, 1.0, np.array([0.3, 0.0, 0.0, 1.5, 0.0, 0.0]))]   # x^1.5 on [0, 1.5]
    spec = sme.models.builtin("lotka_volterra").replace(pdf=pdf)
    exe = _build(sme, spec, str(tmp_path))
    x = sample_points("lotka_volterra", 300, seed=17)
    x[0, 3] = 0.0        # log(0) * c3 > 0: weight 0, not NaN
    x[1, 0] = 2.5        # outside
    x[2, 2] = 4.0        # on the upper end
    orc = oracle.OracleProblem.builtin("lotka_volterra")
    orc.set_pdf(pdf)
    ref, cref, sref = orc.eval_nodes(x, weight_by_pdf=True, want_steps=True)
    dx, steps, _, _, c, _ = _run(exe, spec, str(tmp_path), x, "nodes_dyn")
    _check_nodes(dx, steps, c, ref, cref, sref, tol=2e-6)
    assert np.all(dx[0] == 0.0) and np.all(dx[1] == 0.0) and np.all(dx[2] != 0.0)
    _, _, s, s2, c, _ = _run(exe, spec, str(tmp_path), x, "reduce_dyn")
    so, s2o, co = orc.reduce_samples(x, weight_by_pdf=True)
    _check_sums(s, s2, c, so, s2o, co)


def test_controller_limits_and_failures_on_the_cpu(sme, tmp_path):
    """dtmax below the span (the clip is compiled in on demand), maxiters aborts, and NaN / Inf sample points: counted
    per trajectory exactly as the oracle counts them, every clean column untouched."""
    name = "lorenz63"
    x = sample_points(name, 300, seed=15)
    tmp = str(tmp_path)
    for kw in (dict(dtmax=0.02), dict(maxiters=9)):
        spec = sme.models.builtin(name, **kw)
        exe = _build(sme, spec, tmp)
        orc = oracle.OracleProblem.builtin(name, **kw)
        ref, cref, sref = orc.eval_nodes(x, weight_by_pdf=False, want_steps=True)
        dx, steps, _, _, c, _ = _run(exe, spec, tmp, x, "nodes_dyn", weight=0, grid=2)
        assert c == cref and np.array_equal(steps, sref), kw
        assert np.max(np.abs(dx - ref) / np.abs(ref).max(axis=0)) <= 2e-6
        assert (c["nfail"] == len(x)) == ("maxiters" in kw)
    spec = sme.models.builtin(name)
    exe = _build(sme, spec, tmp)
    orc = oracle.OracleProblem.builtin(name)
    clean, _, _, _, c_clean, _ = _run(exe, spec, tmp, x, "nodes", weight=0, grid=2)
    bad = x.copy()
    bad[7, 3] = np.nan
    bad[100, 0] = np.inf
    bad[211, 4] = -np.inf
    dirty, steps, _, _, c_dirty, _ = _run(exe, spec, tmp, bad, "nodes", weight=0, grid=2)
    _, cref, sref = orc.eval_nodes(bad, weight_by_pdf=False, want_steps=True)
    mask = np.ones(len(x), dtype=bool)
    mask[[7, 100, 211]] = False
    assert np.array_equal(dirty[mask], clean[mask]) and not np.all(np.isfinite(dirty[~mask]))
    assert c_clean["nfail"] == 0 and c_dirty["nfail"] == 3 and np.array_equal(steps, sref)
    # every counter, RHS evaluations of the FAILED trajectories included: an infinite parameter makes the initial-step
    # heuristic see d1 = Inf, the reference's d0 / d1 = 0 takes the "singular" exit after two evaluations -- and so
    # does the kernel (round 1 evaluated a third time there)
    assert c_dirty == cref, (c_dirty, cref)
