"""Sanitizer runs that do not need a GPU ("compute-sanitizer ... ; host side ASan/UBSan").

`compute-sanitizer` is CLOSED on this GPU pool (gpurun answers: "compute-sanitizer is closed on this pool and stays
closed", profiles/r2_compute_sanitizer_closed.log), so the device-side race / memory checks are made where they can be
made: the kernel SOURCE the B200 runs, compiled for the host against tests/cuda_emul (one OS thread per CUDA thread,
warp collectives over barriers, shared memory as statics, atomics as atomics), runs under

  * AddressSanitizer + UndefinedBehaviorSanitizer: out-of-bounds accesses to x / dx / partials / chunk rows / shared
    arrays, signed overflow and invalid shifts in the index arithmetic;
  * ThreadSanitizer: unsynchronised accesses between the threads of a block -- the per-lane shared-memory
    accumulators, the warp / block reduction buffers, `is_last`, the per-warp super-chunk rows, the log / exp tables
    -- i.e. what `racecheck` would look at (blocks run one after the other in the emulation, so the cross-block part
    of the last-block fold, ticket + fence, is not concurrent here; it is covered by the result checks on the GPU).

The oracle and the host side of libb200expect (model validation, the KDE prefilter, NVRTC assembly, the host sampler)
run under ASan + UBSan as well.  scripts/cpu_sanitizers.sh runs this file and keeps the log under profiles/."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from tests.sampling_util import sample_points
from tests.test_kernel_emulation import CSRC, EMUL, ROOT, _run

GXX = shutil.which("g++")
pytestmark = pytest.mark.skipif(GXX is None, reason="g++ not available")


def _build_sanitized(spec, tmp, flags, *, repro=0, em_spt=2, block=64):
    model = os.path.join(tmp, "model.inc")
    open(model, "w").write(spec.source)
    exe = os.path.join(tmp, "driver_san")
    clip = 1 if (spec.dtmax > 0 and spec.dtmax < spec.tspan[1] - spec.tspan[0]) else 0
    defs = dict(B200E_NSTATE=spec.nstate, B200E_NPARAM=spec.nparam, B200E_NX=spec.nx, B200E_NOUT=spec.nout,
                B200E_NKKL=max(1, spec.n_kkl), B200E_SOLVER=spec.solver, B200E_FP32=0, B200E_BLOCK=block, B200E_MINBLOCKS=1,
                B200E_DYNAMIC=1, B200E_REPRO=repro, B200E_NSAVE=len(spec.save_times), B200E_TABLE_PDF=0,
                B200E_CLIP_DTMAX=clip, B200E_EM_SPT=em_spt)
    cmd = [GXX, "-std=c++20", "-O1", "-g", "-pthread", "-Wno-unknown-pragmas", *flags, "-I", EMUL, "-I", CSRC,
           "-I", os.path.join(CSRC, "kernels")]
    cmd += [f"-D{k}={v}" for k, v in defs.items()] + [f'-DMODEL_FILE="{model}"', os.path.join(EMUL, "driver.cpp"), "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def _run_sanitized(monkeypatch, capfd, exe, spec, tmp, x, mode, **kw):
    monkeypatch.setenv("ASAN_OPTIONS", "detect_leaks=1:halt_on_error=1:abort_on_error=0")
    monkeypatch.setenv("UBSAN_OPTIONS", "halt_on_error=1:print_stacktrace=1")
    monkeypatch.setenv("TSAN_OPTIONS", "halt_on_error=1:exitcode=66:second_deadlock_stack=1")
    out = _run(exe, spec, tmp, x, mode, **kw)      # check=True: a sanitizer finding exits non-zero
    err = capfd.readouterr().err
    assert "Sanitizer" not in err and "runtime error" not in err, err[-2000:]
    return out


CASES = [
    # model, instantiation, launch shape, what it walks
    ("linear2", "reduce_dyn", dict(grid=2), {}),                       # atomic work counter, per-lane accumulators, fold
    ("linear2", "nodes", dict(grid=3, layout=1), {}),                  # static ownership, SoA stores
    ("lotka_volterra", "reduce_dyn", dict(grid=2, refill=8), {}),      # partial refills under divergent step counts
    ("linear2", "reduce_dyn", dict(grid=2), dict(repro=1)),            # super-chunk rows in shared memory
    ("linear2", "gen", dict(grid=2, gen=(7, 100)), {}),                # sampler in the kernel
    ("oscillator_noise", "reduce_dyn", dict(grid=2), dict(em_spt=3)),  # Euler-Maruyama, three samples per thread
    ("gbm4_lamba", "reduce_dyn", dict(grid=2), {}),                    # LambaEM step function
]


@pytest.mark.parametrize("name,mode,shape,build", CASES)
def test_kernels_under_asan_ubsan(sme, tmp_path, monkeypatch, capfd, name, mode, shape, build):
    spec = sme.models.builtin(name)
    tmp = str(tmp_path)
    exe = _build_sanitized(spec, tmp, ["-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer"], **build)
    n = 777 if spec.solver == 0 else 300
    x = sample_points(name.replace("_lamba", ""), n, seed=41)
    dx, steps, s, s2, c, work = _run_sanitized(monkeypatch, capfd, exe, spec, tmp, x, mode, **shape)
    assert c["nfail"] == 0 and c["naccept"] > 0
    assert np.all(work == 0)                       # the launch re-armed its counters


@pytest.mark.parametrize("name,mode,shape,build", [CASES[0], CASES[2], CASES[3], CASES[5]])
def test_kernels_under_tsan(sme, tmp_path, monkeypatch, capfd, name, mode, shape, build):
    probe = subprocess.run([GXX, "-fsanitize=thread", "-x", "c++", "-", "-o", str(tmp_path / "probe")], input="int main(){}",
                           capture_output=True, text=True)
    if probe.returncode != 0:
        pytest.skip("libtsan not available")
    spec = sme.models.builtin(name)
    tmp = str(tmp_path)
    exe = _build_sanitized(spec, tmp, ["-fsanitize=thread", "-fno-omit-frame-pointer"], **build)
    n = 400 if spec.solver == 0 else 200
    x = sample_points(name, n, seed=43)
    dx, steps, s, s2, c, work = _run_sanitized(monkeypatch, capfd, exe, spec, tmp, x, mode, **shape)
    assert c["nfail"] == 0 and c["naccept"] > 0


def test_tsan_run_has_teeth(sme, tmp_path, monkeypatch, capfd):
    """Negative control: the same kernel with __syncthreads() compiled out races on the block-reduction buffers, and
    ThreadSanitizer says so -- a clean run above therefore means something."""
    probe = subprocess.run([GXX, "-fsanitize=thread", "-x", "c++", "-", "-o", str(tmp_path / "probe")], input="int main(){}",
                           capture_output=True, text=True)
    if probe.returncode != 0:
        pytest.skip("libtsan not available")
    spec = sme.models.builtin("linear2")
    tmp = str(tmp_path)
    exe = _build_sanitized(spec, tmp, ["-fsanitize=thread", "-fno-omit-frame-pointer", "-DEMU_DROP_BLOCK_BARRIER"])
    x = sample_points("linear2", 400, seed=43)
    monkeypatch.setenv("TSAN_OPTIONS", "halt_on_error=1:exitcode=66")
    with pytest.raises(subprocess.CalledProcessError) as e:
        _run(exe, spec, tmp, x, "reduce_dyn", grid=2)
    assert e.value.returncode == 66 and "ThreadSanitizer: data race" in capfd.readouterr().err


@pytest.mark.parametrize("san", ["thread", "address,undefined"])
def test_copy_pool_under_tsan_and_asan(tmp_path, san):
    """The pageable -> pinned staging pool of the host-fed path (csrc/host_copy_pool.h: worker threads, generation
    hand-over, non-temporal copies) on its own: many sizes and alignments, two callers at once, sleeping workers."""
    exe = str(tmp_path / "copy_pool_test")
    res = subprocess.run([GXX, "-std=c++17", "-O1", "-g", "-pthread", f"-fsanitize={san}", "-fno-omit-frame-pointer", "-I", CSRC,
                          os.path.join(EMUL, "copy_pool_test.cpp"), "-o", exe], capture_output=True, text=True)
    if res.returncode != 0 and "cannot find" in res.stderr:
        pytest.skip(f"-fsanitize={san} runtime not available")
    assert res.returncode == 0, res.stderr[-2000:]
    env = dict(os.environ, B200E_COPY_THREADS="6", TSAN_OPTIONS="halt_on_error=1:exitcode=66", ASAN_OPTIONS="halt_on_error=1",
               UBSAN_OPTIONS="halt_on_error=1")
    run = subprocess.run([exe], env=env, capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "0 mismatches" in run.stdout, (run.stdout, run.stderr[-3000:])
    assert "Sanitizer" not in run.stderr and "runtime error" not in run.stderr, run.stderr[-3000:]


_HOST_SCRIPT = r'''
import os, sys
sys.path.insert(0, ROOT)
import numpy as np
import oracle
from tests.sampling_util import sample_points
# ---- the oracle under ASan + UBSan: every solver, the KDE prefilter and spline, reductions with OpenMP threads ----
for name, kw in (("linear2", {}), ("lotka_volterra", {}), ("lorenz63", {}), ("oscillator_noise", {}),
                 ("gbm4", {"solver": oracle.SOLVER_LAMBA_EM, "abstol": 1e-3, "reltol": 1e-3}), ("decay1", {})):
    orc = oracle.OracleProblem.builtin(name, **kw)
    x = sample_points(name, 257, seed=5)
    dx, c = orc.eval_nodes(x, weight_by_pdf=True, nthreads=3)
    s, s2, c2 = orc.reduce_samples(x, nthreads=3)
    assert c == c2 and np.all(np.isfinite(dx))
f = np.exp(-np.linspace(-2, 2, 65) ** 2)
coef = oracle.kde_prefilter(f)
one = oracle.OracleProblem.builtin("decay1")
one.set_pdf([(4, -2.0, 2.0, 0.0, 1.0, f)])
assert abs(one.pdf(np.array([0.0])) - 1.0) < 1e-12 and one.pdf(np.array([2.5])) == 0.0
for n in (2, 3):
    assert len(oracle.kde_prefilter(f[:n])) == n + 2
print("oracle ok")
# ---- the host side of libb200expect under ASan + UBSan (no GPU: every compute entry point must refuse) ----
import scimlexpectations_jl_b200 as sme
L = sme.load_library()
spec = sme.models.builtin("lotka_volterra")
tab = [(sme._capi.PDF_TABLE, lo, hi, 0.0, 1.0, np.exp(-np.linspace(-1, 1, 129) ** 2)) for lo, hi in ((1, 2), (.5, 1.5), (2, 4), (.5, 1.5))]
nbytes, log = sme.compile_check(spec.replace(pdf=tab))           # validate_and_copy + KDE prefilter + NVRTC assembly
assert nbytes > 0
for bad in (dict(nstate=0), dict(nout=65), dict(tspan=(1.0, 1.0)), dict(block_size=48), dict(abstol=0.0, reltol=0.0),
            dict(pdf=[(sme._capi.PDF_TABLE, 0.0, 1.0, 0.0, 1.0, np.array([1.0]))] + list(spec.pdf[1:]))):
    try:
        sme.compile_check(spec.replace(**bad))
    except sme.B200Error:
        pass
    else:
        raise AssertionError(bad)
x = sme.sample_host(sme.models.builtin("lorenz63"), 3, 200_000)  # threaded host twin of the sampler
xs = sme.sample_host(sme.models.builtin("lorenz63"), 3, 1000, first_index=5, layout=1)
assert np.array_equal(xs.T, x[5:1005])
if sme.device_count() == 0:
    try:
        sme.Handle(spec)
    except sme.B200Error as e:
        assert e.code == sme._capi.ERR_NO_DEVICE
print("host side ok")
'''


def test_oracle_and_library_host_side_under_asan_ubsan(tmp_path):
    """Both shared objects rebuilt with -fsanitize=address,undefined into a scratch directory and driven from a Python
    subprocess with libasan preloaded (the interpreter itself is not instrumented; leak detection is therefore off)."""
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    asan = subprocess.run([gcc, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    ubsan = subprocess.run([gcc, "-print-file-name=libubsan.so"], capture_output=True, text=True).stdout.strip()
    if not (os.path.isabs(asan) and os.path.exists(asan)):
        pytest.skip("libasan not available")
    tmp = str(tmp_path)
    san = ["-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer", "-g"]
    orc_so = os.path.join(tmp, "libexpect_oracle_asan.so")
    subprocess.run([gcc, "-O1", "-mavx2", "-mfma", "-ffp-contract=fast", "-fopenmp", "-fPIC", "-std=gnu11", *san, "-shared", "-o", orc_so,
                    os.path.join(ROOT, "oracle", "expect_oracle.c"), os.path.join(ROOT, "oracle", "models_builtin.c"), "-lm"],
                   check=True, capture_output=True, text=True)
    nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else shutil.which("nvcc")
    if nvcc is None:
        pytest.skip("nvcc not available")
    import __graft_entry__ as ge
    ge._load_build_module().generate_embedded()
    lib_so = os.path.join(tmp, "libb200expect_asan.so")
    cmd = [nvcc, "-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-Xcompiler",
           "-fPIC,-Wno-unused-function,-Wno-unknown-pragmas,-ffp-contract=off,-fsanitize=address,-fsanitize=undefined,-fno-sanitize-recover=all,-fno-omit-frame-pointer",
           "-shared", "-cudart", "static", "-o", lib_so, os.path.join(CSRC, "b200_expect.cu"), "-lnvrtc", "-ldl", "-lpthread"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    env = dict(os.environ, LD_PRELOAD=asan + (":" + ubsan if os.path.exists(ubsan) else ""),
               ASAN_OPTIONS="detect_leaks=0:halt_on_error=1:protect_shadow_gap=0", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1",
               B200E_ORACLE_LIB=orc_so, B200E_LIB=lib_so, B200E_CACHE_DIR="")
    res = subprocess.run([sys.executable, "-c", f"ROOT = {ROOT!r}\n" + _HOST_SCRIPT], env=env, capture_output=True, text=True, timeout=600)
    sys.stdout.write(res.stdout[-2000:])
    assert res.returncode == 0 and "oracle ok" in res.stdout and "host side ok" in res.stdout, res.stderr[-4000:]
    assert "Sanitizer" not in res.stderr and "runtime error" not in res.stderr, res.stderr[-4000:]
