"""Host twin of the counter-based sampler (csrc/kernels/expect_sampling.cuh), CPU side.

The sampler replaces ``rand(d)`` of the MonteCarlo prob_func (reference src/expectation.jl:277-280,
src/distribution_utils.jl:43).  Here: the Philox4x32-10 known-answer vectors of the Random123
distribution, the restated log and Wichura's AS241 inverse normal CDF against libm / scipy, and the
statistics + stream indexing of ``b200e_sample_host`` (which needs no GPU).  Bit-identity with the
device is a ``-m gpu`` test (tests/test_gpu_parity.py).
"""
import os
import shutil
import subprocess

import numpy as np
import pytest
import scipy.stats as st
from scipy.special import ndtri

import scimlexpectations_jl_b200 as sme

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KDIR = os.path.join(ROOT, "scimlexpectations.jl_b200", "csrc", "kernels")

SRC = r"""
#include "expect_sampling.cuh"
#include <cstdio>
using namespace b200e_smp;
int main() {
    unsigned int w[4];
    philox4x32_10(0u, 0u, 0u, 0u, 0u, 0u, w);
    printf("%08x %08x %08x %08x\n", w[0], w[1], w[2], w[3]);
    philox4x32_10(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, w);
    printf("%08x %08x %08x %08x\n", w[0], w[1], w[2], w[3]);
    philox4x32_10(0x243f6a88u, 0x85a308d3u, 0x13198a2eu, 0x03707344u, 0xa4093822u, 0x299f31d0u, w);
    printf("%08x %08x %08x %08x\n", w[0], w[1], w[2], w[3]);
    printf("%.17g %.17g\n", uniform52(0u, 0u), uniform52(0xffffffffu, 0xffffffffu));
    for (int i = 1; i < 4000; i++) {
        const double p = i / 4000.0;
        printf("%.17g %.17g\n", p, norm_ppf(p));
    }
    for (int e = -300; e <= 0; e += 3) {
        const double p = std::pow(10.0, (double)e) * 0.731;
        printf("%.17g %.17g %.17g\n", p, norm_ppf(p), det_log(p));
    }
    return 0;
}
"""


@pytest.fixture(scope="module")
def twin_lines(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    d = tmp_path_factory.mktemp("smp_twin")
    (d / "t.cpp").write_text(SRC)
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", "-I", KDIR, str(d / "t.cpp"),
                    "-o", str(d / "t")], check=True)
    return subprocess.run([str(d / "t")], check=True, capture_output=True, text=True).stdout.strip().split("\n")


def test_philox_known_answers(twin_lines):
    # Random123 kat_vectors: philox4x32 10 rounds
    assert twin_lines[0] == "6627e8d5 e169c58d bc57ac4c 9b00dbd8"
    assert twin_lines[1] == "408f276d 41c83b0e a20bc7c6 6d5451fd"
    assert twin_lines[2] == "d16cfe09 94fdcceb 5001e420 24126ea1"


def test_uniform_is_strictly_inside_the_unit_interval(twin_lines):
    lo, hi = (float(v) for v in twin_lines[3].split())
    assert lo == 2.0 ** -53 and hi == 1.0 - 2.0 ** -53


def test_inverse_normal_cdf_and_log(twin_lines):
    grid = np.array([[float(v) for v in l.split()] for l in twin_lines[4:4 + 3999]])
    ref = ndtri(grid[:, 0])
    assert np.max(np.abs(grid[:, 1] - ref) / np.maximum(np.abs(ref), 1e-3)) < 5e-15
    tail = np.array([[float(v) for v in l.split()] for l in twin_lines[4 + 3999:]])
    ref = ndtri(tail[:, 0])
    assert np.max(np.abs(tail[:, 1] - ref) / np.abs(ref)) < 5e-15       # down to p = 7e-301 (third branch of AS241)
    assert np.max(np.abs(tail[:, 2] - np.log(tail[:, 0])) / np.abs(np.log(tail[:, 0]))) < 4e-16


@pytest.mark.parametrize("name", ["linear2", "lotka_volterra", "lorenz63", "oscillator_noise"])
def test_sample_host_statistics_and_indexing(name):
    spec = sme.models.builtin(name)
    n = 200000
    x = sme.sample_host(spec, 20261019, n)
    assert x.shape == (n, spec.nx) and np.all(np.isfinite(x))
    for j, d in enumerate(spec.pdf):
        kind, lo, hi, mu, sigma = d
        col = x[:, j]
        assert col.min() >= lo and col.max() <= hi
        if kind == sme._capi.PDF_UNIFORM:
            cdf = st.uniform(lo, hi - lo).cdf
        elif kind == sme._capi.PDF_NORMAL:
            cdf = st.norm(mu, sigma).cdf
        else:
            cdf = st.truncnorm((lo - mu) / sigma, (hi - mu) / sigma, loc=mu, scale=sigma).cdf
        assert st.kstest(col, cdf).pvalue > 1e-4, (name, j)
    # coordinates are independent draws (different Philox words / counters)
    c = np.corrcoef(x.T)
    assert np.max(np.abs(c - np.eye(spec.nx))) < 0.02
    # sample i is a pure function of (seed, i): any window of the stream, in either layout
    w = sme.sample_host(spec, 20261019, 1000, first_index=12345)
    assert np.array_equal(w, x[12345:13345])
    assert np.array_equal(sme.sample_host(spec, 20261019, 1000, first_index=12345, layout=1), w.T)
    assert not np.array_equal(sme.sample_host(spec, 20261020, 1000), x[:1000])
    # indices beyond 2^32 use the high counter word
    big = sme.sample_host(spec, 1, 4, first_index=(1 << 32) - 2)
    assert len({tuple(r) for r in big}) == 4


def test_sample_host_rejects_models_without_a_density():
    spec = sme.models.builtin("linear2")
    spec.pdf = None
    with pytest.raises(sme.B200Error):
        sme.sample_host(spec, 1, 10)
