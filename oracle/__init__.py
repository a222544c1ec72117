"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/_build/libexpect_oracle.so`` (a plain-C restatement of the reference
hot path, see ``expect_oracle.h``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package; the product
package ``scimlexpectations.jl_b200`` never does.

The reference (SciMLExpectations.jl) is Julia and Julia is absent from this image, so this is a
"port" oracle.  Parity pin: reference README.md:65-69 (see tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
import hashlib

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libexpect_oracle.so")

ORC_MAX_STATE, ORC_MAX_PARAM, ORC_MAX_X, ORC_MAX_OUT, ORC_MAX_KKL, ORC_MAX_SAVE = 16, 32, 32, 64, 32, 64
PDF_NONE, PDF_UNIFORM, PDF_NORMAL, PDF_TRUNCNORMAL = 0, 1, 2, 3
SOLVER_TSIT5, SOLVER_EM, SOLVER_LAMBA_EM = 0, 1, 2

RHS_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double)
MAP_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                     C.POINTER(C.c_double), C.POINTER(C.c_double))
OBS_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double))


class PdfDim(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("lo", C.c_double), ("hi", C.c_double),
                ("mu", C.c_double), ("sigma", C.c_double), ("table", C.POINTER(C.c_double)), ("ntable", C.c_int64)]


class Problem(C.Structure):
    _fields_ = [
        ("nstate", C.c_int32), ("nparam", C.c_int32), ("nx", C.c_int32), ("nout", C.c_int32),
        ("solver", C.c_int32), ("n_kkl", C.c_int32),
        ("rhs", C.c_void_p), ("diffusion", C.c_void_p), ("hmap", C.c_void_p), ("obs", C.c_void_p),
        ("u0_nom", C.POINTER(C.c_double)), ("p_nom", C.POINTER(C.c_double)),
        ("t0", C.c_double), ("t1", C.c_double),
        ("abstol", C.c_double), ("reltol", C.c_double),
        ("dtmax", C.c_double), ("dt_fixed", C.c_double),
        ("maxiters", C.c_int64),
        ("pdf", C.POINTER(PdfDim)),
        ("nsave", C.c_int32), ("_pad", C.c_int32), ("save_t", C.POINTER(C.c_double)), ("obs_at", C.c_void_p),
    ]


class Counters(C.Structure):
    _fields_ = [("nf", C.c_uint64), ("naccept", C.c_uint64), ("nreject", C.c_uint64), ("nfail", C.c_uint64)]

    def as_dict(self):
        return {"nf": int(self.nf), "naccept": int(self.naccept), "nreject": int(self.nreject),
                "nfail": int(self.nfail)}


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc).  Building the checker is not using it."""
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("expect_oracle.c", "models_builtin.c", "expect_oracle.h", "Makefile")):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        override = os.environ.get("B200E_ORACLE_LIB")   # an instrumented build (tests/test_sanitizers.py)
        if not override:
            build()
        L = C.CDLL(override or _LIB_PATH)
        L.orc_logpdf.restype = C.c_double
        L.orc_logpdf.argtypes = [C.POINTER(PdfDim), C.c_int, C.POINTER(C.c_double)]
        L.orc_pdf.restype = C.c_double
        L.orc_pdf.argtypes = [C.POINTER(PdfDim), C.c_int, C.POINTER(C.c_double)]
        L.orc_kde_prefilter.restype = None
        L.orc_kde_prefilter.argtypes = [C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double)]
        L.orc_kkl_w.restype = C.c_double
        L.orc_kkl_w.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double]
        L.orc_builtin_problem.argtypes = [C.c_char_p, C.POINTER(Problem)]
        L.orc_solve_point.argtypes = [C.POINTER(Problem), C.POINTER(C.c_double), C.c_int,
                                      C.POINTER(C.c_double), C.POINTER(Counters)]
        L.orc_eval_nodes.argtypes = [C.POINTER(Problem), C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                     C.c_void_p, C.POINTER(Counters), C.c_void_p, C.c_int]
        L.orc_reduce_samples.argtypes = [C.POINTER(Problem), C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.POINTER(Counters), C.c_int]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().orc_max_threads())


def kde_prefilter(values) -> np.ndarray:
    """Grid values of a KDE -> the n + 2 coefficients of the quadratic B-spline ``pdf(kde, x)`` evaluates."""
    f = np.ascontiguousarray(values, dtype=np.float64)
    c = np.empty(len(f) + 2)
    lib().orc_kde_prefilter(f.ctypes.data_as(C.POINTER(C.c_double)), len(f), c.ctypes.data_as(C.POINTER(C.c_double)))
    return c


class OracleProblem:
    """A problem the oracle can solve.  Holds the ctypes struct and everything it points to."""

    def __init__(self, st: Problem, keep=()):
        self.st = st
        self._keep = list(keep)

    # --- construction -------------------------------------------------------------------
    @classmethod
    def builtin(cls, name: str, **overrides) -> "OracleProblem":
        st = Problem()
        if lib().orc_builtin_problem(name.encode(), C.byref(st)) != 0:
            raise KeyError(name)
        self = cls(st)
        self.configure(**overrides)
        return self

    @classmethod
    def from_source(cls, src: str, *, nstate, nparam, nx, nout, u0, p, tspan, solver=SOLVER_TSIT5,
                    n_kkl=0, pdf=None, save_times=(), **overrides) -> "OracleProblem":
        """Compile a model written in the shared C/CUDA dialect (see models.py) with gcc and wrap it.

        The source defines ``rhs``, ``hmap``, ``observable`` (and ``diffusion`` for EM) with
        ``B200E_FN`` linkage and ``real`` scalars; here ``real`` is double."""
        save_times = [float(t) for t in save_times]
        so = _compile_model(src, has_diffusion=(solver != SOLVER_TSIT5), has_obs_at=bool(save_times))
        m = C.CDLL(so)
        st = Problem()
        st.nstate, st.nparam, st.nx, st.nout = nstate, nparam, nx, nout
        st.solver, st.n_kkl = solver, n_kkl
        st.rhs = C.cast(m.orc_model_rhs, C.c_void_p)
        st.hmap = C.cast(m.orc_model_hmap, C.c_void_p)
        keep_extra = []
        if save_times:
            assert nout % len(save_times) == 0 and len(save_times) <= ORC_MAX_SAVE and nout <= ORC_MAX_OUT
            st.obs_at = C.cast(m.orc_model_obs_at, C.c_void_p)
            ta = (C.c_double * len(save_times))(*save_times)
            st.nsave, st.save_t = len(save_times), C.cast(ta, C.POINTER(C.c_double))
            keep_extra.append(ta)
        else:
            st.obs = C.cast(m.orc_model_obs, C.c_void_p)
        if solver != SOLVER_TSIT5:
            st.diffusion = C.cast(m.orc_model_diffusion, C.c_void_p)
        u0a = (C.c_double * max(1, nstate))(*[float(v) for v in u0])
        pa = (C.c_double * max(1, nparam))(*[float(v) for v in p])
        st.u0_nom, st.p_nom = C.cast(u0a, C.POINTER(C.c_double)), C.cast(pa, C.POINTER(C.c_double))
        st.t0, st.t1 = float(tspan[0]), float(tspan[1])
        st.abstol, st.reltol, st.maxiters = 1e-6, 1e-3, 100000
        self = cls(st, keep=[m, u0a, pa] + keep_extra)
        if pdf is not None:
            self.set_pdf(pdf)
        self.configure(**overrides)
        return self

    def set_pdf(self, dims):
        """dims: list of (kind, lo, hi, mu, sigma[, table values])"""
        arr = (PdfDim * len(dims))()
        for a, d in zip(arr, dims):
            kind, lo, hi, mu, sigma = d[:5]
            a.kind, a.lo, a.hi, a.mu, a.sigma = int(kind), float(lo), float(hi), float(mu), float(sigma)
            if int(kind) == 5:  # log-polynomial density: (5, lo, hi, 0, 1, [c0 .. c5])
                tab = np.ascontiguousarray(d[5], dtype=np.float64)
                assert tab.shape == (6,)
                self._keep.append(tab)
                a.table, a.ntable = tab.ctypes.data_as(C.POINTER(C.c_double)), 6
            if int(kind) == 4:  # tabulated density: (4, lo, hi, 0, 1, grid values) -> spline coefficients
                vals = np.ascontiguousarray(d[5], dtype=np.float64)
                tab = kde_prefilter(vals)
                self._keep.append(tab)
                a.table, a.ntable = tab.ctypes.data_as(C.POINTER(C.c_double)), len(vals)
        self._keep.append(arr)
        self.st.pdf = C.cast(arr, C.POINTER(PdfDim))

    def configure(self, **kw):
        for k, v in kw.items():
            if v is None:
                continue
            if k == "tspan":
                self.st.t0, self.st.t1 = float(v[0]), float(v[1])
            elif k in ("abstol", "reltol", "dtmax", "dt_fixed", "t0", "t1"):
                setattr(self.st, k, float(v))
            elif k in ("maxiters", "solver"):
                setattr(self.st, k, int(v))
            else:
                raise TypeError(f"unknown oracle option {k}")
        return self

    # --- shapes -------------------------------------------------------------------------
    @property
    def nx(self):
        return int(self.st.nx)

    @property
    def nout(self):
        return int(self.st.nout)

    def _check_x(self, x, layout):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if layout == 0:
            assert x.ndim == 2 and x.shape[1] == self.nx, (x.shape, self.nx)
            n = x.shape[0]
        else:
            assert x.ndim == 2 and x.shape[0] == self.nx, (x.shape, self.nx)
            n = x.shape[1]
        return x, n

    # --- evaluation ---------------------------------------------------------------------
    def logpdf(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        return float(lib().orc_logpdf(self.st.pdf, self.nx, x.ctypes.data_as(C.POINTER(C.c_double))))

    def pdf(self, x):
        """pdf(d, x) as the integrand weights with it (tabulated factors multiplied in directly)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        return float(lib().orc_pdf(self.st.pdf, self.nx, x.ctypes.data_as(C.POINTER(C.c_double))))

    def eval_nodes(self, x, *, layout=0, weight_by_pdf=True, nthreads=0, want_steps=False):
        """Koopman batch integrand (reference expectation.jl:182-196).
        layout 0: x is (npts, nx) C-order == Julia's dim x npts column-major."""
        x, n = self._check_x(x, layout)
        dx = np.empty((n, self.nout) if layout == 0 else (self.nout, n), dtype=np.float64)
        c = Counters()
        steps = np.zeros(n, dtype=np.int32) if want_steps else None
        rc = lib().orc_eval_nodes(C.byref(self.st), x.ctypes.data, n, layout, int(bool(weight_by_pdf)),
                                  dx.ctypes.data, C.byref(c),
                                  steps.ctypes.data if want_steps else None, nthreads)
        assert rc == 0
        return (dx, c.as_dict(), steps) if want_steps else (dx, c.as_dict())

    def reduce_samples(self, x, *, layout=0, weight_by_pdf=False, nthreads=0):
        """MonteCarlo sums (reference expectation.jl:277-293); returns (sum, sumsq, counters)."""
        x, n = self._check_x(x, layout)
        s = np.zeros(self.nout)
        s2 = np.zeros(self.nout)
        c = Counters()
        rc = lib().orc_reduce_samples(C.byref(self.st), x.ctypes.data, n, layout,
                                      int(bool(weight_by_pdf)), s.ctypes.data, s2.ctypes.data,
                                      C.byref(c), nthreads)
        assert rc == 0
        return s, s2, c.as_dict()


_PRELUDE = r"""
#include <math.h>
typedef double real;
#define B200E_FN static inline
#define B200E_REAL(x) x
"""

_EPILOGUE = r"""
void orc_model_rhs(double *du, const double *u, const double *p, double t) { rhs(du, u, p, t); }
void orc_model_hmap(const double *x, const double *u0n, const double *pn, double *u0, double *p) { hmap(x, u0n, pn, u0, p); }
"""
_EPILOGUE_OBS = r"""
void orc_model_obs(const double *u, const double *p, double t, double *out) { observable(u, p, t, out); }
"""
_EPILOGUE_OBS_AT = r"""
void orc_model_obs_at(int k, double t, const double *u, const double *p, double *out) { observable_at(k, t, u, p, out); }
"""
_EPILOGUE_DIFF = r"""
void orc_model_diffusion(double *g, const double *u, const double *p, double t) { diffusion(g, u, p, t); }
"""


def _compile_model(src: str, has_diffusion: bool, has_obs_at: bool = False) -> str:
    full = (_PRELUDE + src + _EPILOGUE + (_EPILOGUE_OBS_AT if has_obs_at else _EPILOGUE_OBS)
            + (_EPILOGUE_DIFF if has_diffusion else ""))
    h = hashlib.sha256(full.encode()).hexdigest()[:16]
    d = os.path.join(tempfile.gettempdir(), "b200e_oracle_models")
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, f"model_{h}.so")
    if not os.path.exists(so):
        cfile = os.path.join(d, f"model_{h}.c")
        with open(cfile, "w") as f:
            f.write(full)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.run([cc, "-O2", "-mavx2", "-mfma", "-ffp-contract=fast", "-fPIC", "-shared", "-std=gnu11",
                        "-o", so + ".tmp", cfile, "-lm"], check=True, capture_output=True)
        os.replace(so + ".tmp", so)
    return so
