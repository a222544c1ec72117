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
    double maxl = 0, maxe = 0, maxr = 0, maxs = 0, maxpow = 0;
    for (int i = 0; i < 2000000; i++) {
        const double x = std::exp(std::uniform_real_distribution<double>(-80, 40)(g));
        maxl = std::fmax(maxl, std::fabs(log_pos(T, x) - std::log(x)));
        const double y = std::uniform_real_distribution<double>(-52, 52)(g);
        maxe = std::fmax(maxe, std::fabs(exp_small(T, y) / std::exp(y) - 1));
        maxr = std::fmax(maxr, std::fabs(rcp_pos(x) * x - 1));
        maxs = std::fmax(maxs, std::fabs(sqrt_pos(x) / std::sqrt(x) - 1));
        // the controller's use: gamma * e2old^hb2 / e2^hb1 on the range it can reach
        const double e2 = std::exp(std::uniform_real_distribution<double>(-40, 25)(g));
        const double e2o = std::exp(std::uniform_real_distribution<double>(-18.4, 0)(g));
        const double q = exp_small(T, fma(lit::hb2, log_pos(T, e2o), fma(-lit::hb1, log_pos(T, e2), lit::ln_gamma)));
        const double qr = 0.9 * std::pow(e2o, lit::hb2) / std::pow(e2, lit::hb1);
        maxpow = std::fmax(maxpow, std::fabs(q / qr - 1));
    }
    printf("%.3e %.3e %.3e %.3e %.3e\n", maxl, maxe, maxr, maxs, maxpow);
    const double sp[] = {0.0, 4.9e-324, 1e-310, 1e308, INFINITY, NAN, -NAN, -1.0};
    for (double x : sp) printf("%.6f\n", log_pos(T, x));
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
    maxl, maxe, maxr, maxs, maxpow = (float(v) for v in twin[0].split())
    assert maxl < 3e-14      # absolute; 1.4e-14 is one ulp of |log x| = 80
    assert maxe < 1e-14      # relative
    assert maxr < 4e-16
    assert maxs < 4e-16
    assert maxpow < 1e-14    # the controller factor against two libm pow calls


def test_log_is_finite_and_saturates_for_special_values(twin):
    vals = [float(v) for v in twin[1:9]]
    # 0 / subnormals -> about -711 (controller clamps to q = 1/qmax), huge / inf / NaN / negative
    # patterns -> about +707 (clamps to the reject factor qmin): the kernel needs no special case
    assert all(v < -700 for v in vals[:3])
    assert all(v > 700 for v in vals[3:])
