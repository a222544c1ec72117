// driver.cpp -- TEST INFRASTRUCTURE (see cuda_emul.h): compiles the kernel source of one model for the host and runs
// one launch of one of its five instantiations.  The launch parameters are filled as the library fills them
// (b200_expect.cu: b200e_create / fill_kpdf / reduce_resident_begin), restated here in a few lines because the
// library keeps that code next to its CUDA calls.
//   usage: driver <config file>      (written by tests/test_kernel_emulation.py)
#include "cuda_emul.h"

#include <fstream>
#include <map>
#include <sstream>
#include <string>

#include "b200e_kparams.h"
#include "expect_prelude.cuh"
#include MODEL_FILE
#include "expect_sampling.cuh"
#include "expect_math.cuh"

namespace b200e {
// the integer-pipe comparisons of expect_math.cuh (device-only there), same bit-pattern semantics
inline double abs_bits(double x) { return __longlong_as_double(__double_as_longlong(x) & 0x7fffffffffffffffLL); }
inline double pmin(double a, double b) { return __longlong_as_double(std::min(__double_as_longlong(a), __double_as_longlong(b))); }
inline double pmax(double a, double b) { return __longlong_as_double(std::max(__double_as_longlong(a), __double_as_longlong(b))); }
inline bool same_value(double a, double b) { return __double_as_longlong(a) == __double_as_longlong(b); }
inline bool le_nonneg(double a, double b) { return (unsigned long long)__double_as_longlong(a) <= (unsigned long long)__double_as_longlong(b); }
inline bool lt_nonneg(double a, double b) { return (unsigned long long)__double_as_longlong(a) < (unsigned long long)__double_as_longlong(b); }
inline bool is_negative(double a) { return __double_as_longlong(a) < 0; }
inline bool lt_any_pos(double a, double b) { return __double_as_longlong(a) < __double_as_longlong(b); }
inline double max_with_negative(double a, double q) {
    return __longlong_as_double((long long)std::min((unsigned long long)__double_as_longlong(a), (unsigned long long)__double_as_longlong(q)));
}
}  // namespace b200e

#include "expect_kernels.cuh"

static double std_normal_cdf(double z) { return 0.5 * std::erfc(-z * M_SQRT1_2); }

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    std::ifstream in(argv[1]);
    std::map<std::string, std::vector<double>> cfg;
    std::map<std::string, std::string> str;
    std::vector<std::vector<double>> pdf;
    for (std::string line; std::getline(in, line);) {
        std::istringstream ls(line);
        std::string key;
        ls >> key;
        if (key == "xfile" || key == "outfile" || key == "mode") { ls >> str[key]; continue; }
        std::vector<double> v;
        for (std::string tok; ls >> tok;) v.push_back(tok == "inf" ? INFINITY : tok == "-inf" ? -INFINITY : std::stod(tok));
        if (key == "pdf") pdf.push_back(v); else cfg[key] = v;
    }
    const long long npts = (long long)cfg["npts"][0];
    const unsigned grid = (unsigned)cfg["grid"][0];
    const std::string mode = str["mode"];
    constexpr int NX = B200E_NX, NOUT = B200E_NOUT, NACC = 2 * B200E_NOUT;

    std::vector<double> x((size_t)npts * NX), dx((size_t)npts * NOUT, 0.0);
    if (mode != "gen") {
        FILE *f = fopen(str["xfile"].c_str(), "rb");
        if (!f || fread(x.data(), sizeof(double), x.size(), f) != x.size()) { fprintf(stderr, "cannot read x\n"); return 3; }
        fclose(f);
    }
    std::vector<int> steps((size_t)npts, -1);
    std::vector<double> partials((size_t)grid * NACC, 0.0), final_sums(NACC, 0.0);
    std::vector<unsigned long long> cpartials((size_t)grid * 4, 0ull), final_counts(4, 0ull), work(2, 0ull);

    b200e_kparams P;
    memset(&P, 0, sizeof P);
    P.x = x.data(); P.dx = dx.data(); P.steps = steps.data();
    P.partials = partials.data(); P.cpartials = cpartials.data(); P.work_counter = work.data();
    P.final_sums = final_sums.data(); P.final_counts = final_counts.data();
    P.npts = npts; P.ld = npts; P.layout = (int)cfg["layout"][0]; P.weight_by_pdf = (int)cfg["weight"][0];
    P.t0 = cfg["tspan"][0]; P.t1 = cfg["tspan"][1];
    P.abstol = cfg["tol"][0]; P.reltol = cfg["tol"][1];
    P.dtmax = cfg["dtmax"][0] > 0 ? cfg["dtmax"][0] : P.t1 - P.t0;
    P.dt_fixed = cfg["dt_fixed"][0];
    P.maxiters = (long long)cfg["maxiters"][0];
    P.snap_tol = 100.0 * (std::nextafter(std::fabs(P.t1), INFINITY) - std::fabs(P.t1));
    P.refill_threshold = (int)cfg["refill"][0];
    P.want_sumsq = 1;
    for (size_t i = 0; i < cfg["u0"].size(); i++) P.u0_nom[i] = cfg["u0"][i];
    for (size_t i = 0; i < cfg["p"].size(); i++) P.p_nom[i] = cfg["p"][i];
    for (size_t i = 0; i < cfg["save"].size(); i++) P.save_t[i] = cfg["save"][i];
    // fill_kpdf (b200_expect.cu): per-dimension constants of logpdf and of the sampler
    for (int j = 0; j < B200E_KP_MAX_X; j++) {
        b200e_kpdf &k = P.pdf[j];
        k.lo = -INFINITY; k.hi = INFINITY; k.smp_d = 1.0;
        if (j >= (int)pdf.size()) continue;
        const double kind = pdf[j][0], lo = pdf[j][1], hi = pdf[j][2], mu = pdf[j][3], sg = pdf[j][4];
        k.kind = (int)kind;
        if (k.kind == 1) { k.lo = lo; k.hi = hi; k.logc = -std::log(hi - lo); k.smp_a = lo; k.smp_b = hi - lo; }
        if (k.kind == 2) { k.mu = mu; k.inv_sigma = 1.0 / sg; k.logc = -std::log(sg) - 0.91893853320467274178; k.smp_a = mu; k.smp_b = sg; k.smp_c = 0.0; k.smp_d = 1.0; }
        if (k.kind == 3) {
            const double tp = std_normal_cdf((hi - mu) / sg) - std_normal_cdf((lo - mu) / sg);
            k.lo = lo; k.hi = hi; k.mu = mu; k.inv_sigma = 1.0 / sg;
            k.logc = -std::log(sg) - 0.91893853320467274178 - std::log(tp);
            k.smp_a = mu; k.smp_b = sg; k.smp_c = std_normal_cdf((lo - mu) / sg); k.smp_d = tp;
        }
        if (k.kind == 5) {  // log-polynomial: the six coefficients follow the five common fields
            k.lo = lo; k.hi = hi; k.logc = pdf[j][5]; k.mu = pdf[j][6]; k.inv_sigma = pdf[j][7];
            k.smp_a = pdf[j][8]; k.smp_b = pdf[j][9]; k.smp_c = pdf[j][10];
        }
    }
    std::vector<double> kkl;
#if B200E_SOLVER == 1
    {
        const long long ns = (long long)std::ceil((P.t1 - P.t0) / P.dt_fixed - 1e-9);
        P.em_nsteps = ns;
        kkl.resize((size_t)(ns + 1) * B200E_NKKL);
        for (long long s = 0; s <= ns; s++) {
            const double t = (s == ns) ? P.t1 : P.t0 + (double)s * P.dt_fixed;
            for (int k = 1; k <= B200E_NKKL; k++)
                kkl[(size_t)s * B200E_NKKL + (k - 1)] = std::sqrt(2.0) * std::sin((k - 0.5) * M_PI * t) / ((k - 0.5) * M_PI);
        }
        P.kkl_table = kkl.data();
    }
#endif
    std::vector<double> rows;
    long long nrows = 0;
    const bool reduce = mode.rfind("reduce", 0) == 0 || mode == "gen";
    if (mode == "gen") { P.gen = 1; P.gen_seed = (unsigned long long)cfg["gen"][0]; P.gen_first = (long long)cfg["gen"][1]; P.x = nullptr; }
#if B200E_REPRO
    if (mode == "reduce_dyn" || mode == "gen") {  // reproducible dynamic schedule: one row of sums per super-chunk
        P.super_chunks = 12;
        nrows = (npts + 32 * 12 - 1) / (32 * 12);
        rows.assign((size_t)nrows * NACC, 0.0);
        P.chunk_rows = rows.data();
        P.final_sums = nullptr;
    }
#endif
    if (mode == "nodes") emu_launch(b200e::b200e_nodes, P, grid, B200E_BLOCK);
    else if (mode == "nodes_dyn") emu_launch(b200e::b200e_nodes_dyn, P, grid, B200E_BLOCK);
    else if (mode == "reduce") emu_launch(b200e::b200e_reduce, P, grid, B200E_BLOCK);
    else if (mode == "reduce_dyn") emu_launch(b200e::b200e_reduce_dyn, P, grid, B200E_BLOCK);
    else if (mode == "gen") emu_launch(b200e::b200e_reduce_gen, P, grid, B200E_BLOCK);
    else return 4;
    if (nrows) {  // fold the rows in index order (the library does it with two small kernels)
        for (int c = 0; c < NACC; c++) {
            double s = 0.0;
            for (long long r = 0; r < nrows; r++) s += rows[(size_t)r * NACC + c];
            final_sums[c] = s;
        }
    }
    FILE *o = fopen(str["outfile"].c_str(), "wb");
    fwrite(dx.data(), sizeof(double), dx.size(), o);
    fwrite(steps.data(), sizeof(int), steps.size(), o);
    fwrite(final_sums.data(), sizeof(double), NACC, o);
    fwrite(final_counts.data(), sizeof(unsigned long long), 4, o);
    fwrite(work.data(), sizeof(unsigned long long), 2, o);
    fclose(o);
    (void)reduce;
    return 0;
}
