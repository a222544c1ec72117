"""Host twin of the device math (csrc/kernels/expect_math.cuh compiled with g++ -mfma).

The Tsit5 kernel evaluates the PI controller's power function as one table-driven log and one
table-driven exp (reference: OrdinaryDiffEq's PIController behind src/expectation.jl:191 / :290 uses
two `pow`).  The same header compiles for the host, so accuracy and the special-value behaviour the
kernel relies on (finite result for 0 / inf / NaN, so no branch is needed) are checked here on the
CPU box against libm.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KDIR = os.path.join(ROOT, "scimlexpectations.jl_b200", "csrc", "kernels")

SRC = r"""
#define B200E_NSTATE 2
#define B200E_FP32 0
typedef double real;
#include "expect_math.cuh"
#include <cstdio>
#include <random>
using namespace b200e;
int main() {
    static b200e_tabs T;
    for (int j = 0; j < TAB_N; j++) tab_fill(T, j);
    std::mt19937_64 g(1);
    double maxl = 0, maxe = 0, maxr = 0, maxs = 0, maxpow = 0, maxrs = 0;
    for (int i = 0; i < 2000000; i++) {
        const double x = std::exp(std::uniform_real_distribution<double>(-80, 40)(g));
        maxl = std::fmax(maxl, std::fabs(log_pos(T, x) - std::log(x)));
        const double y = std::uniform_real_distribution<double>(-52, 52)(g);
        maxe = std::fmax(maxe, std::fabs(exp_small(T, y) / std::exp(y) - 1));
        maxr = std::fmax(maxr, std::fabs(rcp_pos(x) * x - 1));
        maxs = std::fmax(maxs, std::fabs(sqrt_pos(x) / std::sqrt(x) - 1));
        double s2_, rs_;
        sqrt_rsqrt_pos(x, s2_, rs_);
        maxs = std::fmax(maxs, std::fabs(s2_ / std::sqrt(x) - 1));
        maxrs = std::fmax(maxrs, std::fabs(rs_ * std::sqrt(x) - 1));
        // the controller's use: gamma * e2old^hb2 / e2^hb1 on the range it can reach
        const double e2 = std::exp(std::uniform_real_distribution<double>(-40, 25)(g));
        const double e2o = std::exp(std::uniform_real_distribution<double>(-18.4, 0)(g));
        const double q = exp_small(T, fma(lit::hb2, log_pos(T, e2o), fma(-lit::hb1, log_pos(T, e2), lit::ln_gamma)));
        const double qr = 0.9 * std::pow(e2o, lit::hb2) / std::pow(e2, lit::hb1);
        maxpow = std::fmax(maxpow, std::fabs(q / qr - 1));
    }
    printf("%.3e %.3e %.3e %.3e %.3e %.3e\n", maxl, maxe, maxr, maxs, maxpow, maxrs);
    const double sp[] = {0.0, 4.9e-324, 1e-310, 1e308, INFINITY, NAN, -NAN, -1.0};
    for (double x : sp) printf("%.6f\n", log_pos(T, x));
    // sin(pi t / 2) of the LambaEM noise path against libm, and the recurrence s_{k+1} = 2 cos(pi t) s_k - s_{k-1}
    // for sin((k - 1/2) pi t), k <= 32, built on it (expect_kernels.cuh kkl_w)
    double maxsin = 0, maxrec = 0;
    for (int i = 0; i < 2000000; i++) {
        const double t = std::uniform_real_distribution<double>(-3, 40)(g);
        const double s = sin_halfpi(t);
        maxsin = std::fmax(maxsin, std::fabs(s - std::sin(1.5707963267948966192 * t)));
        if (i % 16 == 0) {
            const double c2 = fma(-4.0 * s, s, 2.0);
            double sp_ = -s, sk = s;
            for (int k = 1; k <= 32; k++) {
                maxrec = std::fmax(maxrec, std::fabs(sk - std::sin((k - 0.5) * 3.14159265358979323846 * t)));
                const double sn = fma(c2, sk, -sp_);
                sp_ = sk;
                sk = sn;
            }
        }
    }
    printf("%.3e %.3e\n", maxsin, maxrec);
    const double ex[] = {0.0, 1.0, 2.0, 3.0, -1.0, 0.5, 1e15};
    for (double t : ex) printf("%.17g\n", sin_halfpi(t));
    return 0;
}
"""


@pytest.fixture(scope="module")
def twin(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    d = tmp_path_factory.mktemp("math_twin")
    src = d / "t.cpp"
    src.write_text(SRC)
    exe = d / "t"
    subprocess.run([gxx, "-O2", "-mfma", "-std=c++17", "-I", KDIR, str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    return out


def test_table_log_exp_rcp_sqrt_accuracy(twin):
    maxl, maxe, maxr, maxs, maxpow, maxrs = (float(v) for v in twin[0].split())
    # accuracy matched to the use (step-size control only, expect_math.cuh rcp_pos): series and Newton steps stop at
    # ~1e-12; the test pins those levels from above AND from below (a silent return to the long forms, or a further cut,
    # both show)
    assert 1e-14 < maxl < 2.5e-13    # absolute: the dropped term r^5 / 5, |r| <= 2^-8
    assert 1e-13 < maxe < 3e-12      # relative: the dropped term r^4 / 24, |r| <= ln2 / 256
    assert 1e-14 < maxr < 1.2e-12    # (2^-19.9)^2 after one Newton step on the 20-bit seed
    assert maxs < 4e-16              # sqrt: one coupled step + the correction is already full precision
    assert maxrs < 2e-12             # 1/sqrt out of the same step: ~(3/2) (2^-20)^2, used by the LambaEM error estimate only
    assert maxpow < 3.5e-12          # the controller factor against two libm pow calls


def test_log_is_finite_and_saturates_for_special_values(twin):
    vals = [float(v) for v in twin[1:9]]
    # 0 / subnormals -> about -709 (controller clamps to q = 1/qmax), huge / inf / NaN / negative
    # patterns -> +709 and above (clamps to the reject factor qmin): the kernel needs no special case
    assert all(v < -700 for v in vals[:3])
    assert all(v > 700 for v in vals[3:])


def test_sin_halfpi_and_the_kkl_recurrence(twin):
    """The one sine of the adaptive process-noise path (K8) and the three-term recurrence that turns it into all
    KKL modes: absolute errors against libm (whose own argument pi t / 2 is rounded: ~4e-15 at t = 40)."""
    maxsin, maxrec = (float(v) for v in twin[9].split())
    assert maxsin < 8e-15
    assert maxrec < 2e-11      # 32 modes; the recurrence amplifies by ~k^2 near t = even integers
    exact = [float(v) for v in twin[10:17]]
    want = [0.0, 1.0, 0.0, -1.0, -1.0, 2 ** -0.5]     # t = 0, 1, 2, 3, -1, 1/2: zeros are exact
    assert exact[0] == 0.0 and exact[2] == 0.0 and all(abs(a - b) < 2.3e-16 for a, b in zip(exact, want))
    assert abs(exact[6]) <= 1.0                        # t = 1e15: still a sine
