"""Host-side h-adaptive cubature driver: the caller ABOVE the hot path.

In the reference, ``solve(exprob, Koopman(); quadalg=CubatureJLh(), batch=...)`` hands the batched
integrand to Integrals.jl -> Cubature.jl -> the C library ``cubature`` (S. G. Johnson), function
``hcubature_v`` (reference call site: src/expectation.jl:347-351 -> :357-363; the integrand it calls
back is src/expectation.jl:193-196).  Cubature.jl is a third-party dependency that is not vendored
under the reference tree (Project.toml compat "1.5"), and Julia is not in this image, so this module
restates the published algorithm of ``hcubature_v``:

  * Genz-Malik degree-7 rule with embedded degree-5 rule for dim >= 2 (2^d + 2d^2 + 2d + 1 nodes),
    Gauss-Kronrod 7/15 for dim == 1 with QUADPACK's error heuristic;
  * per-region error = |I7 - I5| per output, split dimension = largest summed fourth difference;
  * a max-heap of regions keyed by the largest per-output error; in the vectorised variant every
    round pops the minimal set of largest-error regions whose removal brings the remaining error
    under the tolerance, bisects them all, and evaluates ALL their nodes in ONE integrand call --
    that call is the refinement batch shipped to the GPU;
  * convergence (error_norm = INDIVIDUAL, Integrals.jl's default for CubatureJLh): every output j has
    err_j <= abstol or err_j <= |val_j| * reltol.

The integrand sees ``x`` as ``(npts, dim)`` C-order, which is byte-for-byte Julia's ``dim x npts``
column-major array, and returns ``(npts, fdim)``.
"""
from __future__ import annotations

import heapq
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

# ---- Genz-Malik constants -----------------------------------------------------------------------
_LAMBDA2 = np.sqrt(9.0 / 70.0)
_LAMBDA4 = np.sqrt(9.0 / 10.0)
_LAMBDA5 = np.sqrt(9.0 / 19.0)
_RATIO = (9.0 / 70.0) / (9.0 / 10.0)  # lambda2^2 / lambda4^2


def genz_malik_num_points(dim: int) -> int:
    return 1 + 4 * dim + 2 * dim * (dim - 1) + (1 << dim)


@dataclass
class GenzMalikRule:
    """Node offsets (in units of the half-widths) and the two weight sets for one dimension count."""
    dim: int
    offsets: np.ndarray = field(init=False)   # (npoints, dim)
    w7: np.ndarray = field(init=False)        # (npoints,) degree-7 weights (times volume)
    w5: np.ndarray = field(init=False)        # (npoints,) embedded degree-5 weights

    def __post_init__(self):
        d = self.dim
        assert d >= 2
        w1 = (12824.0 - 9120.0 * d + 400.0 * d * d) / 19683.0
        w2 = 980.0 / 6561.0
        w3 = (1820.0 - 400.0 * d) / 19683.0
        w4 = 200.0 / 19683.0
        w5 = 6859.0 / 19683.0 / (1 << d)
        e1 = (729.0 - 950.0 * d + 50.0 * d * d) / 729.0
        e2 = 245.0 / 486.0
        e3 = (265.0 - 100.0 * d) / 1458.0
        e4 = 25.0 / 729.0
        offs, w7, w5e = [np.zeros(d)], [w1], [e1]
        for lam, wa, wb in ((_LAMBDA2, w2, e2), (_LAMBDA4, w3, e3)):
            for i in range(d):
                for s in (1.0, -1.0):
                    o = np.zeros(d)
                    o[i] = s * lam
                    offs.append(o); w7.append(wa); w5e.append(wb)
        for i in range(d - 1):
            for j in range(i + 1, d):
                for si in (1.0, -1.0):
                    for sj in (1.0, -1.0):
                        o = np.zeros(d)
                        o[i] = si * _LAMBDA4
                        o[j] = sj * _LAMBDA4
                        offs.append(o); w7.append(w4); w5e.append(e4)
        for mask in range(1 << d):
            o = np.array([(-_LAMBDA5 if (mask >> i) & 1 else _LAMBDA5) for i in range(d)])
            offs.append(o); w7.append(w5); w5e.append(0.0)
        self.offsets = np.array(offs)
        self.w7 = np.array(w7)
        self.w5 = np.array(w5e)
        assert len(self.offsets) == genz_malik_num_points(d)

    @property
    def num_points(self) -> int:
        return len(self.offsets)

    def nodes(self, centers: np.ndarray, halfwidths: np.ndarray) -> np.ndarray:
        """(nregions, dim) x2 -> (nregions * npoints, dim), region-major."""
        x = centers[:, None, :] + self.offsets[None, :, :] * halfwidths[:, None, :]
        return x.reshape(-1, self.dim)

    def apply(self, fvals: np.ndarray, halfwidths: np.ndarray):
        """fvals (nregions, npoints, fdim) -> (val, err, splitdim)."""
        d = self.dim
        vol = np.prod(2.0 * halfwidths, axis=1)                    # (nregions,)
        i7 = np.einsum("k,rkf->rf", self.w7, fvals)
        i5 = np.einsum("k,rkf->rf", self.w5, fvals)
        val = vol[:, None] * i7
        err = np.abs(val - vol[:, None] * i5)
        f0 = fvals[:, 0, :]
        f2 = fvals[:, 1:1 + 2 * d, :].reshape(len(fvals), d, 2, -1).sum(axis=2)          # (r, d, fdim)
        f4 = fvals[:, 1 + 2 * d:1 + 4 * d, :].reshape(len(fvals), d, 2, -1).sum(axis=2)
        diff = np.abs(f2 - 2.0 * f0[:, None, :] - _RATIO * (f4 - 2.0 * f0[:, None, :])).sum(axis=2)  # (r, d)
        split = np.argmax(diff, axis=1)  # first maximum, as the C loop with a strict '>' does
        return val, err, split


# ---- Gauss-Kronrod 7/15 (dim == 1) ---------------------------------------------------------------
_XGK = np.array([0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                 0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                 0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                 0.207784955007898467600689403773245, 0.000000000000000000000000000000000])
_WG = np.array([0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                0.381830050505118944950369775488975, 0.417959183673469387755102040816327])
_WGK = np.array([0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                 0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                 0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                 0.204432940075298892414161999234649, 0.209482141084727828012999174891714])


class GaussKronrod15Rule:
    dim = 1
    num_points = 15

    def __init__(self):
        # node order: centre, then +-xgk[j] for j = 0..6
        offs = [0.0]
        for j in range(7):
            offs += [-_XGK[j], _XGK[j]]
        self.offsets = np.array(offs)[:, None]

    def nodes(self, centers, halfwidths):
        x = centers[:, None, :] + self.offsets[None, :, :] * halfwidths[:, None, :]
        return x.reshape(-1, 1)

    def apply(self, fvals, halfwidths):
        h = halfwidths[:, 0][:, None]                      # (r, 1)
        fc = fvals[:, 0, :]
        fm = fvals[:, 1::2, :]                             # f(c - h xgk[j]), j = 0..6
        fp = fvals[:, 2::2, :]
        v = fm + fp
        rk = fc * _WGK[7] + np.einsum("j,rjf->rf", _WGK[:7], v)
        rg = fc * _WG[3] + np.einsum("j,rjf->rf", _WG[:3], v[:, 1::2, :])   # xgk[1], xgk[3], xgk[5]
        rabs = np.abs(fc) * _WGK[7] + np.einsum("j,rjf->rf", _WGK[:7], np.abs(fm) + np.abs(fp))
        mean = 0.5 * rk
        rasc = _WGK[7] * np.abs(fc - mean) + np.einsum(
            "j,rjf->rf", _WGK[:7], np.abs(fm - mean[:, None, :]) + np.abs(fp - mean[:, None, :]))
        val = rk * h
        err = np.abs(rk - rg) * h
        rabs = rabs * h
        rasc = rasc * h
        with np.errstate(divide="ignore", invalid="ignore"):
            scale = np.power(200.0 * err / rasc, 1.5)
        resc = np.where(scale < 1.0, rasc * scale, rasc)
        err = np.where((rasc != 0.0) & (err != 0.0), resc, err)
        eps, tiny = np.finfo(float).eps, np.finfo(float).tiny
        min_err = 50.0 * eps * rabs
        err = np.where((rabs > tiny / (50.0 * eps)) & (min_err > err), min_err, err)
        return val, err, np.zeros(len(fvals), dtype=np.int64)


def make_rule(dim: int):
    return GaussKronrod15Rule() if dim == 1 else GenzMalikRule(dim)


# ---- the adaptive driver ---------------------------------------------------------------------------
@dataclass
class CubatureResult:
    u: np.ndarray            # (fdim,) integral estimate
    resid: np.ndarray        # (fdim,) error estimate (sum of per-region |I7 - I5|)
    nevals: int              # integrand evaluations (nodes)
    nrounds: int             # integrand CALLS (= refinement batches shipped)
    nregions: int
    converged: bool
    batch_sizes: List[int]   # nodes per round

    retcode = property(lambda self: "Success" if self.converged else "MaxIters")


def _converged(val, err, abstol, reltol) -> bool:
    # ERROR_INDIVIDUAL
    return bool(np.all((err <= abstol) | (err <= np.abs(val) * reltol)))


def hcubature_v(f: Callable[[np.ndarray], np.ndarray], lb, ub, *, fdim: int, reltol: float = 1e-8,
                abstol: float = 0.0, maxevals: int = 0, max_batch: Optional[int] = None,
                record: Optional[list] = None) -> CubatureResult:
    """Vectorised h-adaptive cubature of ``f`` over the box [lb, ub].

    ``f(x)`` receives ``x`` of shape ``(npts, dim)`` and returns ``(npts, fdim)`` (or ``(npts,)`` when
    fdim == 1).  ``max_batch`` splits a round into several integrand calls of at most that many nodes
    (Cubature itself ignores Integrals.jl's ``max_batch``; None reproduces that).  ``record``, if a
    list, receives every node batch ``x`` (the "same Cubature node batches" both arms consume)."""
    lb = np.atleast_1d(np.asarray(lb, dtype=np.float64))
    ub = np.atleast_1d(np.asarray(ub, dtype=np.float64))
    dim = len(lb)
    rule = make_rule(dim)
    npts_rule = rule.num_points

    nrounds = 0
    batch_sizes: List[int] = []

    def eval_regions(centers, halfwidths):
        nonlocal nrounds
        x = np.ascontiguousarray(rule.nodes(centers, halfwidths))
        if record is not None:
            record.append(x.copy())
        outs = []
        step = len(x) if not max_batch else max(npts_rule, (max_batch // npts_rule) * npts_rule)
        for i in range(0, len(x), step):
            xi = x[i:i + step]
            vi = np.asarray(f(xi), dtype=np.float64)
            outs.append(vi.reshape(len(xi), -1))
            nrounds += 1
            batch_sizes.append(len(xi))
        v = np.concatenate(outs, axis=0) if len(outs) > 1 else outs[0]
        assert v.shape == (len(x), fdim), (v.shape, len(x), fdim)
        return rule.apply(v.reshape(len(centers), npts_rule, fdim), halfwidths)

    # region store (struct of arrays) + heap of (-errmax, serial, index)
    centers = [(lb + ub) * 0.5]
    halfw = [(ub - lb) * 0.5]
    val0, err0, split0 = eval_regions(np.array(centers), np.array(halfw))
    vals, errs, splits = [val0[0]], [err0[0]], [int(split0[0])]
    heap = [(-float(err0[0].max()), 0, 0)]
    serial = 1
    tot_val, tot_err = val0[0].copy(), err0[0].copy()
    numeval = npts_rule
    alive = {0}
    converged = False

    while numeval < maxevals or not maxevals:
        if _converged(tot_val, tot_err, abstol, reltol):
            converged = True
            break
        # pop the minimal set of largest-error regions whose removal leaves a converged remainder
        rem_err = tot_err.copy()
        round_val = tot_val.copy()   # the C code tests convergence with the round's starting value
        popped = []
        while True:
            _, _, idx = heapq.heappop(heap)
            alive.discard(idx)
            tot_val -= vals[idx]
            tot_err -= errs[idx]
            rem_err -= errs[idx]
            popped.append(idx)
            numeval += 2 * npts_rule
            if _converged(round_val, rem_err, abstol, reltol):
                break
            if not heap or not (numeval < maxevals or not maxevals):
                break
        # bisect each popped region along its split dimension
        nc, nh = [], []
        for idx in popped:
            d = splits[idx]
            h = halfw[idx].copy()
            h[d] *= 0.5
            c1, c2 = centers[idx].copy(), centers[idx].copy()
            c1[d] -= h[d]
            c2[d] += h[d]
            nc += [c1, c2]
            nh += [h, h.copy()]
        v, e, s = eval_regions(np.array(nc), np.array(nh))
        for k in range(len(nc)):
            idx = len(centers)
            centers.append(nc[k]); halfw.append(nh[k])
            vals.append(v[k]); errs.append(e[k]); splits.append(int(s[k]))
            heapq.heappush(heap, (-float(e[k].max()), serial, idx))
            serial += 1
            alive.add(idx)
            tot_val += v[k]
            tot_err += e[k]
    else:
        converged = _converged(tot_val, tot_err, abstol, reltol)

    # re-sum integral and errors over the final region set
    ids = sorted(alive)
    u = np.sum([vals[i] for i in ids], axis=0)
    resid = np.sum([errs[i] for i in ids], axis=0)
    return CubatureResult(u=np.asarray(u, dtype=np.float64), resid=np.asarray(resid, dtype=np.float64),
                          nevals=numeval, nrounds=nrounds, nregions=len(ids), converged=converged,
                          batch_sizes=batch_sizes)
