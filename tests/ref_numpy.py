"""A second, independent (and slow) restatement of the ODE leg of the hot path, in plain Python.

It exists to catch a bug shared by the C oracle and the CUDA kernel: it is written from the published
algorithm descriptions (Tsitouras 2011 tableau, Hairer-Wanner initial step, PI controller with the
OrdinaryDiffEq defaults listed in SURVEY.md 8c) with a different code structure (Butcher matrix loops
instead of unrolled stages).  Only for small cases."""
import math

import numpy as np

C = [0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0]
A = [
    [],
    [0.161],
    [-0.008480655492356989, 0.335480655492357],
    [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
    [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
    [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383],
    [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774],
]
BTILDE = [-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
          0.5823571654525552, -0.45808210592918697, 0.015151515151515152]


def _rms(v):
    return math.sqrt(sum(x * x for x in v) / len(v))


def tsit5(f, u0, p, t0, t1, abstol=1e-6, reltol=1e-3, maxiters=100000):
    """Returns (u(t1), nf, naccept, nreject)."""
    u = [float(x) for x in u0]
    n = len(u)
    t = t0
    dtmax = t1 - t0
    nf = 0
    f0 = f(u, p, t); nf += 1
    sk = [abstol + abs(u[i]) * reltol for i in range(n)]
    d0 = _rms([u[i] / sk[i] for i in range(n)])
    d1 = _rms([f0[i] / sk[i] for i in range(n)])
    dt0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * (d0 / d1)
    dt0 = min(dt0, dtmax)
    f1 = f([u[i] + dt0 * f0[i] for i in range(n)], p, t + dt0); nf += 1
    d2 = _rms([(f1[i] - f0[i]) / sk[i] for i in range(n)]) / dt0
    dm = max(d1, d2)
    dt1 = max(1e-6, dt0 * 1e-3) if dm <= 1e-15 else 10.0 ** (-(2.0 + math.log10(dm)) / 5.0)
    dt = min(100.0 * dt0, dt1, dtmax)
    k = [None] * 7
    k[0] = f(u, p, t); nf += 1
    qold = 1e-4
    nacc = nrej = 0
    it = 0
    while t < t1:
        it += 1
        if it > maxiters:
            break
        dt = min(dt, dtmax, t1 - t)
        for s in range(1, 7):
            us = [u[i] + dt * sum(A[s][j] * k[j][i] for j in range(s)) for i in range(n)]
            k[s] = f(us, p, t + C[s] * dt)
            nf += 1
        unew = us  # stage 7 is evaluated at the new solution (FSAL)
        est = 0.0
        for i in range(n):
            ut = dt * sum(BTILDE[j] * k[j][i] for j in range(7))
            est += (ut / (abstol + max(abs(u[i]), abs(unew[i])) * reltol)) ** 2
        est = math.sqrt(est / n)
        if est == 0.0:
            q = 0.1
            q11 = 0.0
        else:
            q11 = est ** (7.0 / 50.0)
            q = q11 / qold ** (2.0 / 25.0)
            q = max(0.1, min(5.0, q / 0.9))
        if est <= 1.0:
            qold = max(est, 1e-4)
            tn = t + dt
            if abs(tn - t1) < 100.0 * (np.nextafter(max(t, t1), np.inf) - max(t, t1)):
                tn = t1
            t = tn
            u = unew
            k[0] = k[6]
            dt = min(dtmax, dt / q)
            nacc += 1
        else:
            dt = dt / min(5.0, q11 / 0.9)
            nrej += 1
    return u, nf, nacc, nrej


def linear2_rhs(u, p, t):
    return [u[1], -u[0] - 2.0 * u[1]]


def lv_rhs(u, p, t):
    return [(p[0] - p[1] * u[1]) * u[0], (p[3] * u[0] - p[2]) * u[1]]


def lorenz_rhs(u, p, t):
    return [p[0] * (u[1] - u[0]), u[0] * (p[1] - u[2]) - u[1], u[0] * u[1] - p[2] * u[2]]
