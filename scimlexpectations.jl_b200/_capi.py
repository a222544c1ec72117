"""ctypes binding of ``lib/libb200expect.so`` -- a literal mirror of ``include/b200_expect.h``.

Every call the Julia shim (``julia/B200Expectations.jl``) makes with ``ccall`` is made here with
ctypes, argument for argument, so the Python tests exercise exactly the boundary the reference
would bind.  There is no CPU fallback: if the library is missing, or no B200 is visible, the
compute entry points raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libb200expect.so")

ABI_VERSION = 4
MAX_STATE, MAX_PARAM, MAX_X, MAX_OUT, MAX_KKL, MAX_SAVE = 16, 32, 32, 64, 32, 64
OK, ERR_INVALID, ERR_COMPILE, ERR_CUDA, ERR_NO_DEVICE, ERR_NOMEM, ERR_NCCL = 0, -1, -2, -3, -4, -5, -6
SOLVER_TSIT5, SOLVER_EM, SOLVER_LAMBA_EM = 0, 1, 2
FP64, FP32 = 0, 1
LAYOUT_POINT_MAJOR, LAYOUT_SOA = 0, 1
PDF_NONE, PDF_UNIFORM, PDF_NORMAL, PDF_TRUNCNORMAL, PDF_TABLE, PDF_LOGPOLY = 0, 1, 2, 3, 4, 5
SCHED_DYNAMIC, SCHED_STATIC, SCHED_DYNAMIC_REPRODUCIBLE = 0, 1, 2

_ERR_NAMES = {ERR_INVALID: "B200E_ERR_INVALID", ERR_COMPILE: "B200E_ERR_COMPILE", ERR_CUDA: "B200E_ERR_CUDA",
              ERR_NO_DEVICE: "B200E_ERR_NO_DEVICE", ERR_NOMEM: "B200E_ERR_NOMEM", ERR_NCCL: "B200E_ERR_NCCL"}

# every symbol include/b200_expect.h declares (tests check the built library exports each one)
EXPORTS = [
    "b200e_abi_version", "b200e_create", "b200e_destroy", "b200e_compile_check", "b200e_last_error",
    "b200e_compile_log", "b200e_eval_nodes", "b200e_reduce_samples", "b200e_last_counters",
    "b200e_last_stats", "b200e_set_steps_buffer", "b200e_host_alloc", "b200e_host_free",
    "b200e_device_alloc", "b200e_device_free", "b200e_memcpy_h2d", "b200e_memcpy_d2h",
    "b200e_measure_fma_peak", "b200e_device_count", "b200e_eval_regions", "b200e_koopman_solve",
    "b200e_reduce_generated", "b200e_sample_host", "b200e_sample_device",
    "b200e_reduce_samples_begin", "b200e_reduce_samples_end",
]


class B200Error(RuntimeError):
    """A negative status from libb200expect (the Julia shim throws on the same condition)."""

    def __init__(self, code: int, message: str):
        self.code = code
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {message}")


class PdfDim(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("lo", C.c_double), ("hi", C.c_double),
                ("mu", C.c_double), ("sigma", C.c_double), ("table", C.POINTER(C.c_double)), ("ntable", C.c_int64)]


class Model(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("nstate", C.c_int32), ("nparam", C.c_int32), ("nx", C.c_int32), ("nout", C.c_int32),
        ("solver", C.c_int32), ("n_kkl", C.c_int32), ("fp_mode", C.c_int32),
        ("source", C.c_char_p),
        ("u0", C.POINTER(C.c_double)), ("p", C.POINTER(C.c_double)),
        ("t0", C.c_double), ("t1", C.c_double),
        ("abstol", C.c_double), ("reltol", C.c_double),
        ("dtmax", C.c_double), ("dt_fixed", C.c_double),
        ("maxiters", C.c_int64),
        ("pdf", C.POINTER(PdfDim)),
        ("n_devices", C.c_int32),
        ("devices", C.POINTER(C.c_int32)),
        ("block_size", C.c_int32), ("blocks_per_sm", C.c_int32),
        ("refill_threshold", C.c_int32), ("chunk_points", C.c_int32),
        ("schedule", C.c_int32), ("nsave", C.c_int32), ("save_times", C.POINTER(C.c_double)),
        ("cache_dir", C.c_char_p),
    ]


class Counters(C.Structure):
    _fields_ = [("nf", C.c_uint64), ("naccept", C.c_uint64), ("nreject", C.c_uint64), ("nfail", C.c_uint64)]

    def as_dict(self):
        return {"nf": int(self.nf), "naccept": int(self.naccept), "nreject": int(self.nreject),
                "nfail": int(self.nfail)}


class CallStats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("total_ms", C.c_double), ("launches", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("grid", C.c_int32), ("block", C.c_int32), ("regs_per_thread", C.c_int32),
                ("blocks_per_sm_occ", C.c_int32), ("local_bytes_per_thread", C.c_int32),
                ("smem_bytes", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CubatureOpts(C.Structure):
    _fields_ = [("reltol", C.c_double), ("abstol", C.c_double), ("maxevals", C.c_int64)]


class CubatureStats(C.Structure):
    _fields_ = [("nevals", C.c_int64), ("nrounds", C.c_int64), ("nregions", C.c_int64), ("max_batch", C.c_int64),
                ("converged", C.c_int32), ("_pad", C.c_int32), ("kernel_ms", C.c_double), ("total_ms", C.c_double),
                ("counters", Counters)]


_lib = None


def load_library():
    """Load ``libb200expect.so``.  Fails loudly when it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("B200E_LIB") or LIB_PATH     # B200E_LIB: the same variable the Julia shim reads
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python scimlexpectations.jl_b200/build.py` "
            "(or __graft_entry__.build()).  This package has no CPU fallback.")
    L = C.CDLL(path)
    H = C.c_void_p
    L.b200e_abi_version.restype = C.c_int
    L.b200e_device_count.restype = C.c_int
    L.b200e_create.argtypes = [C.POINTER(Model), C.POINTER(H)]
    L.b200e_destroy.argtypes = [H]
    L.b200e_compile_check.argtypes = [C.POINTER(Model), C.POINTER(C.c_int64)]
    L.b200e_last_error.argtypes = [H]
    L.b200e_last_error.restype = C.c_char_p
    L.b200e_compile_log.argtypes = [H]
    L.b200e_compile_log.restype = C.c_char_p
    L.b200e_eval_nodes.argtypes = [H, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    L.b200e_reduce_samples.argtypes = [H, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.POINTER(Counters)]
    L.b200e_last_counters.argtypes = [H, C.POINTER(Counters)]
    L.b200e_last_stats.argtypes = [H, C.POINTER(CallStats)]
    L.b200e_set_steps_buffer.argtypes = [H, C.c_void_p]
    L.b200e_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_int64]
    L.b200e_host_free.argtypes = [C.c_void_p]
    L.b200e_device_alloc.argtypes = [H, C.POINTER(C.c_void_p), C.c_int64]
    L.b200e_device_free.argtypes = [H, C.c_void_p]
    L.b200e_memcpy_h2d.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int64]
    L.b200e_memcpy_d2h.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int64]
    L.b200e_measure_fma_peak.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_double)]
    L.b200e_eval_regions.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200e_koopman_solve.argtypes = [H, C.c_void_p, C.c_void_p, C.POINTER(CubatureOpts), C.c_void_p, C.c_void_p,
                                      C.POINTER(CubatureStats)]
    L.b200e_reduce_samples_begin.argtypes = [H, C.c_void_p, C.c_int64, C.c_int32, C.c_int32]
    L.b200e_reduce_samples_end.argtypes = [H, C.c_void_p, C.c_void_p, C.POINTER(Counters)]
    L.b200e_reduce_generated.argtypes = [H, C.c_uint64, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(Counters)]
    L.b200e_sample_host.argtypes = [C.POINTER(Model), C.c_uint64, C.c_uint64, C.c_int64, C.c_int32, C.c_void_p]
    L.b200e_sample_device.argtypes = [H, C.c_uint64, C.c_uint64, C.c_int64, C.c_int32, C.c_void_p]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("b200e_abi_version", "b200e_device_count"):
            fn.restype = C.c_int
    if L.b200e_abi_version() != ABI_VERSION:
        raise ImportError(f"libb200expect ABI {L.b200e_abi_version()} != binding ABI {ABI_VERSION}")
    _lib = L
    return L


def sample_host(spec: "ModelSpec", seed: int, n: int, *, first_index: int = 0, layout=LAYOUT_POINT_MAJOR):
    """``b200e_sample_host``: samples ``first_index .. first_index+n-1`` of stream ``seed`` of the model's
    product density, bit-identical to what the kernels draw.  Needs no GPU."""
    L = load_library()
    m, keep = spec.to_c()
    x = np.empty((n, spec.nx) if layout == LAYOUT_POINT_MAJOR else (spec.nx, n), dtype=np.float64)
    rc = L.b200e_sample_host(C.byref(m), int(seed), int(first_index), int(n), int(layout), x.ctypes.data)
    del keep
    if rc != OK:
        raise B200Error(rc, _tls_error())
    return x


def device_count() -> int:
    return int(load_library().b200e_device_count())


def _tls_error() -> str:
    return load_library().b200e_last_error(None).decode(errors="replace")


@dataclass
class ModelSpec:
    """Everything the hot path closes over (mirror of ``b200e_model``)."""
    source: str
    nstate: int
    nparam: int
    nx: int
    nout: int
    u0: Sequence[float]
    p: Sequence[float]
    tspan: Sequence[float]
    solver: int = SOLVER_TSIT5
    n_kkl: int = 0
    fp_mode: int = FP64
    abstol: float = 1e-6
    reltol: float = 1e-3
    dtmax: float = 0.0
    dt_fixed: float = 0.0
    maxiters: int = 100000
    pdf: Optional[Sequence[tuple]] = None  # [(kind, lo, hi, mu, sigma[, table values for PDF_TABLE])] * nx
    devices: Optional[Sequence[int]] = None
    block_size: int = 0
    blocks_per_sm: int = 0
    refill_threshold: int = 0
    chunk_points: int = 0
    schedule: int = 0            # SCHED_DYNAMIC (default) / SCHED_STATIC (bit-reproducible sums)
    save_times: Sequence[float] = ()  # time-sampled observable: saveat / the arguments of sol(t); nout = len * outputs per time
    cache_dir: Optional[str] = None

    def replace(self, **kw) -> "ModelSpec":
        import dataclasses
        return dataclasses.replace(self, **kw)

    def to_c(self):
        """Build the C struct; returns (struct, keepalive list)."""
        m = Model()
        keep = []
        m.struct_size = C.sizeof(Model)
        m.nstate, m.nparam, m.nx, m.nout = int(self.nstate), int(self.nparam), int(self.nx), int(self.nout)
        m.solver, m.n_kkl, m.fp_mode = int(self.solver), int(self.n_kkl), int(self.fp_mode)
        src = self.source.encode()
        keep.append(src)
        m.source = src
        u0 = (C.c_double * max(1, len(self.u0)))(*[float(v) for v in self.u0])
        p = (C.c_double * max(1, len(self.p)))(*[float(v) for v in self.p])
        keep += [u0, p]
        m.u0 = C.cast(u0, C.POINTER(C.c_double))
        m.p = C.cast(p, C.POINTER(C.c_double))
        m.t0, m.t1 = float(self.tspan[0]), float(self.tspan[1])
        m.abstol, m.reltol = float(self.abstol), float(self.reltol)
        m.dtmax, m.dt_fixed = float(self.dtmax), float(self.dt_fixed)
        m.maxiters = int(self.maxiters)
        if self.pdf is not None:
            arr = (PdfDim * len(self.pdf))()
            for a, d in zip(arr, self.pdf):
                kind, lo, hi, mu, sigma = d[:5]
                a.kind, a.lo, a.hi, a.mu, a.sigma = int(kind), float(lo), float(hi), float(mu), float(sigma)
                if int(kind) == PDF_LOGPOLY:  # (PDF_LOGPOLY, lo, hi, 0, 1, [c0 .. c5]): log-polynomial density
                    tab = np.ascontiguousarray(d[5], dtype=np.float64)
                    if tab.shape != (6,):
                        raise ValueError("a PDF_LOGPOLY factor takes exactly six coefficients c0 .. c5")
                    keep.append(tab)
                    a.table, a.ntable = tab.ctypes.data_as(C.POINTER(C.c_double)), 6
                if int(kind) == PDF_TABLE:  # (PDF_TABLE, lo, hi, 0, 1, values): tabulated density on a uniform grid
                    tab = np.ascontiguousarray(d[5], dtype=np.float64)
                    keep.append(tab)
                    a.table, a.ntable = tab.ctypes.data_as(C.POINTER(C.c_double)), len(tab)
            m.pdf = C.cast(arr, C.POINTER(PdfDim))
        if self.devices is not None:
            devs = (C.c_int32 * len(self.devices))(*[int(d) for d in self.devices])
            keep.append(devs)
            m.devices = C.cast(devs, C.POINTER(C.c_int32))
            m.n_devices = len(self.devices)
        m.block_size, m.blocks_per_sm = int(self.block_size), int(self.blocks_per_sm)
        m.refill_threshold, m.chunk_points = int(self.refill_threshold), int(self.chunk_points)
        m.schedule = int(self.schedule)
        if len(self.save_times):
            ta = (C.c_double * len(self.save_times))(*[float(t) for t in self.save_times])
            keep.append(ta)
            m.nsave, m.save_times = len(self.save_times), C.cast(ta, C.POINTER(C.c_double))
        cache = self.cache_dir if self.cache_dir is not None else default_cache_dir()
        if cache:
            cb = cache.encode()
            keep.append(cb)
            m.cache_dir = cb
        return m, keep


def default_cache_dir() -> str:
    """In-tree cubin cache (travels to the GPU box with the snapshot).  ``B200E_CACHE_DIR`` wins."""
    env = os.environ.get("B200E_CACHE_DIR")
    if env is not None:
        return env
    return os.path.join(HERE, "_cache")


def compile_check(spec: ModelSpec):
    """NVRTC-compile a model for sm_100a without a GPU.  Returns (cubin_bytes, log)."""
    L = load_library()
    m, keep = spec.to_c()
    n = C.c_int64(0)
    rc = L.b200e_compile_check(C.byref(m), C.byref(n))
    if rc != OK:
        raise B200Error(rc, _tls_error())
    return int(n.value), _tls_error()


def measure_fma_peak(device: int = -1, fp32: bool = False) -> float:
    """FMA-chain microbenchmark (TFLOP/s): the measured roofline denominator of this path."""
    L = load_library()
    v = C.c_double(0.0)
    rc = L.b200e_measure_fma_peak(int(device), int(bool(fp32)), C.byref(v))
    if rc != OK:
        raise B200Error(rc, _tls_error())
    return float(v.value)


class PinnedArray:
    """float64 array in page-locked host memory (``b200e_host_alloc``); ``.array`` is a numpy view."""

    def __init__(self, shape):
        L = load_library()
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        n = int(np.prod(self.shape))
        self._ptr = C.c_void_p()
        rc = L.b200e_host_alloc(C.byref(self._ptr), n * 8)
        if rc != OK:
            raise B200Error(rc, _tls_error())
        buf = (C.c_double * max(1, n)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=np.float64, count=n).reshape(self.shape)

    def free(self):
        if self._ptr is not None and self._ptr.value:
            self.array = None
            load_library().b200e_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceArray:
    """float64 buffer in the handle's HBM (``b200e_device_alloc``) for resident sample sets."""

    def __init__(self, handle: "Handle", shape):
        self.handle = handle
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.nbytes = int(np.prod(self.shape)) * 8
        self._ptr = C.c_void_p()
        handle._check(handle.L.b200e_device_alloc(handle._h, C.byref(self._ptr), self.nbytes))

    @property
    def ptr(self) -> int:
        return int(self._ptr.value)

    def data_ptr(self) -> int:
        return self.ptr

    def upload(self, a: np.ndarray):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.nbytes == self.nbytes, (a.shape, self.shape)
        self.handle._check(self.handle.L.b200e_memcpy_h2d(self.handle._h, self._ptr, a.ctypes.data, a.nbytes))
        return self

    def download(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=np.float64)
        self.handle._check(self.handle.L.b200e_memcpy_d2h(self.handle._h, out.ctypes.data, self._ptr, out.nbytes))
        return out

    def free(self):
        if self._ptr is not None and self._ptr.value and self.handle._h:
            self.handle.L.b200e_device_free(self.handle._h, self._ptr)
        self._ptr = None


def _as_pointer(x, expect_elems: Optional[int] = None, output: bool = False):
    """(address, keepalive) for a numpy array, PinnedArray, DeviceArray, or anything with data_ptr().

    An INPUT array of another dtype / stride is passed as a contiguous float64 copy.  An OUTPUT buffer must be
    written in place, so it has to be float64, C-contiguous and writeable already (a copy would swallow the
    results): anything else raises."""
    if isinstance(x, PinnedArray):
        if expect_elems is not None and int(np.prod(x.shape)) != expect_elems:
            raise ValueError(f"buffer of shape {x.shape} where {expect_elems} elements are expected")
        return x._ptr.value, x
    if isinstance(x, DeviceArray):
        if expect_elems is not None and int(np.prod(x.shape)) != expect_elems:
            raise ValueError(f"buffer of shape {x.shape} where {expect_elems} elements are expected")
        return x.ptr, x
    if isinstance(x, np.ndarray):
        if output:
            if x.dtype != np.float64 or not x.flags.c_contiguous or not x.flags.writeable:
                raise ValueError("an output buffer must be a writeable C-contiguous float64 array "
                                 f"(got dtype {x.dtype}, c_contiguous={x.flags.c_contiguous}, writeable={x.flags.writeable})")
            if expect_elems is not None and x.size != expect_elems:
                raise ValueError(f"output buffer of shape {x.shape} where {expect_elems} elements are expected")
        elif x.dtype != np.float64 or not x.flags.c_contiguous:
            x = np.ascontiguousarray(x, dtype=np.float64)
        if expect_elems is not None:
            assert x.size == expect_elems, (x.shape, expect_elems)
        return x.ctypes.data, x
    if hasattr(x, "data_ptr"):  # torch tensor (host pinned or cuda), float64, contiguous
        if hasattr(x, "is_contiguous"):
            assert x.is_contiguous(), "tensor must be contiguous"
        return int(x.data_ptr()), x
    raise TypeError(f"cannot pass {type(x)} across the C ABI")


class Handle:
    """One compiled model on one (or several) B200s: ``b200e_create`` ... ``b200e_destroy``."""

    def __init__(self, spec: ModelSpec):
        self.L = load_library()
        self.spec = spec
        self._h = C.c_void_p()
        m, keep = spec.to_c()
        rc = self.L.b200e_create(C.byref(m), C.byref(self._h))
        if rc != OK:
            self._h = None
            raise B200Error(rc, _tls_error())
        self._steps_keep = None

    # -- plumbing --
    def _check(self, rc: int):
        if rc != OK:
            raise B200Error(rc, self.L.b200e_last_error(self._h).decode(errors="replace"))

    def close(self):
        if getattr(self, "_h", None):
            self.L.b200e_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def compile_log(self) -> str:
        return self.L.b200e_compile_log(self._h).decode(errors="replace")

    def last_counters(self) -> dict:
        c = Counters()
        self._check(self.L.b200e_last_counters(self._h, C.byref(c)))
        return c.as_dict()

    def last_stats(self) -> dict:
        s = CallStats()
        self._check(self.L.b200e_last_stats(self._h, C.byref(s)))
        return s.as_dict()

    def pinned(self, shape) -> PinnedArray:
        return PinnedArray(shape)

    def device_array(self, shape) -> DeviceArray:
        return DeviceArray(self, shape)

    def _npts(self, x, npts, layout):
        if npts is not None:
            return int(npts)
        shape = x.shape if hasattr(x, "shape") else None
        if shape is None:
            raise ValueError("npts is required for raw pointers")
        shape = tuple(shape)
        nx = self.spec.nx
        if len(shape) == 1:
            assert shape[0] % nx == 0
            return shape[0] // nx
        if layout == LAYOUT_SOA:
            assert shape[0] == nx, (shape, nx)
            return shape[1]
        assert shape[1] == nx, (shape, nx)   # point-major; an unknown layout code is rejected by the library
        return shape[0]

    # -- the two hot entry points --
    def eval_nodes(self, x, *, layout=LAYOUT_POINT_MAJOR, weight_by_pdf=True, out=None, npts=None,
                   want_steps=False):
        """``integrand_koopman_systemmap_batch!(dx, x, p)`` (reference src/expectation.jl:182-196).

        layout 0: ``x`` is ``(npts, nx)`` C-order == Julia's ``dim x npts`` column-major array."""
        n = self._npts(x, npts, layout)
        nout = self.spec.nout
        xp, xk = _as_pointer(x, None)
        if out is None:
            out = np.empty((n, nout) if layout == LAYOUT_POINT_MAJOR else (nout, n), dtype=np.float64)
        op, ok = _as_pointer(out, n * nout, output=True)
        steps = None
        if want_steps:
            steps = np.zeros(n, dtype=np.int32)
            self._check(self.L.b200e_set_steps_buffer(self._h, steps.ctypes.data))
        try:
            self._check(self.L.b200e_eval_nodes(self._h, xp, n, int(layout), int(bool(weight_by_pdf)), op))
        finally:
            if want_steps:
                self.L.b200e_set_steps_buffer(self._h, None)
        del xk, ok
        return (out, steps) if want_steps else out

    def reduce_samples(self, x, *, layout=LAYOUT_POINT_MAJOR, weight_by_pdf=False, npts=None):
        """``solve(exprob, MonteCarlo(n))``'s ensemble solve + sum (reference src/expectation.jl:277-293).

        Returns ``(sum, sumsq, counters)``; the caller divides by n (``mean(sol.u)``)."""
        n = self._npts(x, npts, layout)
        xp, xk = _as_pointer(x, None)
        s = np.zeros(self.spec.nout)
        s2 = np.zeros(self.spec.nout)
        c = Counters()
        self._check(self.L.b200e_reduce_samples(self._h, xp, n, int(layout), int(bool(weight_by_pdf)),
                                                s.ctypes.data, s2.ctypes.data, C.byref(c)))
        del xk
        return s, s2, c.as_dict()

    # -- the caller above the path, on the device --
    def eval_regions(self, centers, halfwidths):
        """One cubature refinement round on the device: regions in, ``(val, err, splitdim)`` per region out."""
        c = np.ascontiguousarray(centers, dtype=np.float64).reshape(-1, self.spec.nx)
        w = np.ascontiguousarray(halfwidths, dtype=np.float64).reshape(-1, self.spec.nx)
        assert c.shape == w.shape
        n = len(c)
        val = np.empty((n, self.spec.nout))
        err = np.empty((n, self.spec.nout))
        split = np.empty(n, dtype=np.int32)
        self._check(self.L.b200e_eval_regions(self._h, c.ctypes.data, w.ctypes.data, n, val.ctypes.data,
                                              err.ctypes.data, split.ctypes.data))
        return val, err, split

    def reduce_samples_begin(self, x, *, npts, layout=LAYOUT_POINT_MAJOR, weight_by_pdf=False):
        """Enqueue the reduction of device-resident points and return at once (``reduce_samples_end`` waits)."""
        xp, xk = _as_pointer(x, None)
        self._check(self.L.b200e_reduce_samples_begin(self._h, xp, int(npts), int(layout), int(bool(weight_by_pdf))))
        self._pending_keep = getattr(self, "_pending_keep", None) or []
        self._pending_keep.append(xk)     # up to two in flight, first in first out

    def reduce_samples_end(self):
        """Wait for the pending reduction: ``(sum, sumsq, counters)``."""
        s = np.zeros(self.spec.nout)
        s2 = np.zeros(self.spec.nout)
        c = Counters()
        self._check(self.L.b200e_reduce_samples_end(self._h, s.ctypes.data, s2.ctypes.data, C.byref(c)))
        if getattr(self, "_pending_keep", None):
            self._pending_keep.pop(0)
        return s, s2, c.as_dict()

    def reduce_generated(self, seed: int, n: int, *, first_index: int = 0):
        """``solve(exprob, MonteCarlo(n))`` with ``rand(d)`` drawn inside the kernel (reference
        src/expectation.jl:277-293): samples ``first_index .. first_index+n-1`` of stream ``seed``.
        Returns ``(sum, sumsq, counters)``; ``sample_host`` yields the identical points for a CPU arm."""
        s = np.zeros(self.spec.nout)
        s2 = np.zeros(self.spec.nout)
        c = Counters()
        self._check(self.L.b200e_reduce_generated(self._h, int(seed), int(first_index), int(n), s.ctypes.data,
                                                  s2.ctypes.data, C.byref(c)))
        return s, s2, c.as_dict()

    def sample_host(self, seed: int, n: int, *, first_index: int = 0, layout=LAYOUT_POINT_MAJOR):
        """Host twin of the in-kernel sampler: bit-identical to what ``reduce_generated`` draws."""
        return sample_host(self.spec, seed, n, first_index=first_index, layout=layout)

    def sample_device(self, seed: int, n: int, *, first_index: int = 0, layout=LAYOUT_POINT_MAJOR, out=None):
        """The same samples written by the device (into ``out``: a DeviceArray, or a new host array)."""
        nx = self.spec.nx
        if out is None:
            out = np.empty((n, nx) if layout == LAYOUT_POINT_MAJOR else (nx, n), dtype=np.float64)
        op, ok = _as_pointer(out, n * nx, output=True)
        self._check(self.L.b200e_sample_device(self._h, int(seed), int(first_index), int(n), int(layout), op))
        del ok
        return out

    def koopman_solve(self, lb, ub, *, reltol=1e-2, abstol=1e-2, maxevals=1000000):
        """``solve(exprob, Koopman(); quadalg=CubatureJLh(), batch=...)`` as ONE C call: hcubature_v's region
        heap in the library, Genz-Malik rule applied on the device.  Returns ``(u, resid, stats)``."""
        lb = np.ascontiguousarray(lb, dtype=np.float64)
        ub = np.ascontiguousarray(ub, dtype=np.float64)
        assert lb.shape == ub.shape == (self.spec.nx,)
        u = np.empty(self.spec.nout)
        resid = np.empty(self.spec.nout)
        opts = CubatureOpts(float(reltol), float(abstol), int(maxevals))
        st = CubatureStats()
        self._check(self.L.b200e_koopman_solve(self._h, lb.ctypes.data, ub.ctypes.data, C.byref(opts), u.ctypes.data,
                                               resid.ctypes.data, C.byref(st)))
        stats = {k: getattr(st, k) for k in ("nevals", "nrounds", "nregions", "max_batch", "converged", "kernel_ms", "total_ms")}
        stats["counters"] = st.counters.as_dict()
        return u, resid, stats
