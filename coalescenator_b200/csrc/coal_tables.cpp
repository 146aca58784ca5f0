#include "coal_tables.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <thread>

namespace coalgpu {
namespace {

// Abramowitz & Stegun table 25.9 nodes as the reference lists them (ConditionalSamplingHMM.cpp:41,46,51)
const double kNodes4[5] = {0, 0.415774556783, 2.294280360279, 6.289945082937, 1E+300};
const double kNodes8[9] = {0, 0.193043676560, 1.026664895339, 2.567876744951, 4.900353084526, 8.182153444563, 12.7344180291798, 19.3957278622, 1E+300};
const double kNodes16[kQ + 1] = {0, 0.093307812017, 0.492691740302, 1.215595412071, 2.269949526204,
    3.667622721751, 5.425336627414, 7.565916226613, 10.120228568019, 13.130282482176, 16.6544077083,
    20.7764788994, 25.6238942267, 31.407519169754, 38.530683306486, 48.026085572686, 1E+300};

// nearest-integer binomial, as boost::math::binomial_coefficient<double> returns it
// (ConditionalSamplingHMM.cpp:79)
double choose(unsigned n, unsigned k) {
    if (k == 0 || k == n) return 1.0;
    if (k == 1 || k == n - 1) return (double)n;
    unsigned m = std::min(k, n - k);
    long double r = 1.0L;
    for (unsigned i = 1; i <= m; i++) r = r * (long double)(n - m + i) / (long double)i;
    return std::ceil((double)r - 0.5);
}

bool nearly_equal(double a, double b) {   // Utils.hpp:206-209
    return (a == b) || (std::fabs(a - b) < std::fabs(std::min(a, b)) * std::numeric_limits<double>::epsilon());
}

// y[i], and z in structured form, ConditionalSamplingHMM.cpp:85-182
void transition_row(int Q, const double* t, const double* w, double r, unsigned n, double* y, double* zlow, double* g, double* zdiag) {
    const double c1 = n / (r + n);
    if (nearly_equal(r, n)) {
        for (int i = 0; i < Q; i++) {
            y[i] = c1 / w[i] * (std::exp(-t[i] / c1) - std::exp(-t[i + 1] / c1));
            const double core = (w[i] + (t[i] * std::exp(-t[i]) - t[i + 1] * std::exp(-t[i + 1])));
            zlow[i] = core;                 // z(i',i) for i' > i  (:114)
            g[i] = core / w[i];             // z(i,j) = w[j]/w[i]*core for j > i  (:118)
            const double term1 = w[i] * core;
            // integer 1/2 == 0 in the reference (:123): the second half of term2 vanishes
            const double term2 = -(t[i] - t[i + 1]) * std::exp(-(t[i] + t[i + 1]));
            zdiag[i] = 1 / w[i] * (term1 + term2);
        }
    } else {
        const double c2 = r / (r - n);
        const double c3 = n / r;
        for (int i = 0; i < Q; i++) {
            y[i] = c1 / w[i] * (std::exp(-t[i] / c1) - std::exp(-t[i + 1] / c1));
            const double inner = (w[i] - c3 * (std::exp(-t[i] / c3) - std::exp(-t[i + 1] / c3)));
            zlow[i] = c2 * inner;           // :156
            g[i] = c2 / w[i] * inner;       // :160 without the w[j] factor
            const double term1 = w[i] * inner;
            const double term2 = c1 / c2 * (std::exp(-t[i] / c1) - std::exp(-t[i + 1] / c1));
            const double term3 = c3 * (std::exp(-t[i]) * std::exp(-t[i + 1] / c3) - std::exp(-t[i + 1]) * std::exp(-t[i] / c3));
            zdiag[i] = c2 / w[i] * (term1 - term2 - term3);   // :164-168
        }
    }
}

// E[i][d], ConditionalSamplingHMM.cpp:184-226. The exponentials depend on (k+l) only, so they are
// evaluated once per (k+l, i) — same arguments, hence the same values, as the reference's inner loop.
void emission_table(int Q, const double* t, const double* w, const std::vector<double>& bin, int bw, double theta,
                    unsigned n, int L, double* E /*[L+1][kEStride]*/) {
    std::vector<double> rate(L + 1), ex((size_t)(L + 1) * (kQ + 1));
    for (int m = 0; m <= L; m++) {
        rate[m] = (2 * theta * (unsigned)m / n) + 1;
        for (int i = 0; i <= Q; i++) ex[(size_t)m * (kQ + 1) + i] = std::exp(-rate[m] * t[i]);
    }
    const double two_L = std::pow(2, (double)L);
    for (int i = 1; i <= Q; i++) {
        for (int j = 0; j <= L; j++) {
            double e = 0;
            for (int l = 0; l <= j; l++) {
                double inner = 0;
                for (int k = 0; k <= L - j; k++) {
                    const int m = k + l;
                    inner += bin[(size_t)(L - j) * bw + k] / rate[m] * (ex[(size_t)m * (kQ + 1) + i - 1] - ex[(size_t)m * (kQ + 1) + i]);
                }
                e += inner * bin[(size_t)j * bw + l] * ((l & 1) ? -1.0 : 1.0);
            }
            e /= (w[i - 1] * two_L);
            e = std::max(e, std::numeric_limits<double>::epsilon());
            E[(size_t)j * kEStride + (i - 1)] = e;   // distance-major: the intervals of one distance are contiguous
        }
    }
}

}  // namespace

bool build_tables(int Q, int LA, int LB, int n_max, const std::vector<double>& thetas,
                  const std::vector<double>& r_rates, HostTables& out) {
    if ((Q != 4 && Q != 8 && Q != kQ) || LA < 1 || LB < 1 || n_max < 1 || thetas.size() != r_rates.size()) return false;
    out.lay.init(LA, LB);
    out.n_max = n_max;
    // Q = 4 or 8 (ConditionalSamplingHMM.cpp:39-47): the kernels always sweep 16 intervals; intervals beyond Q carry
    // weight 0 and all-zero table entries, so they contribute nothing.
    out.Q = Q;
    const double* nodes = Q == 4 ? kNodes4 : (Q == 8 ? kNodes8 : kNodes16);
    out.quad.assign(kQ + 1, 1E+300);
    for (int i = 0; i <= Q; i++) out.quad[i] = nodes[i];
    out.w.assign(kQ, 0.0);
    for (int i = 1; i <= Q; i++) out.w[i - 1] = std::exp(-out.quad[i - 1]) - std::exp(-out.quad[i]);   // :63-71

    // epochs with identical scaled rates share one table set
    const int T = (int)thetas.size();
    out.set_of_epoch.assign(T, -1);
    std::vector<int> rep;
    for (int c = 0; c < T; c++) {
        for (size_t s = 0; s < rep.size(); s++)
            if (thetas[rep[s]] == thetas[c] && r_rates[rep[s]] == r_rates[c]) { out.set_of_epoch[c] = (int)s; break; }
        if (out.set_of_epoch[c] < 0) { out.set_of_epoch[c] = (int)rep.size(); rep.push_back(c); }
    }
    out.n_sets = (int)rep.size();

    const int maxL = std::max(LA, LB), bw = maxL + 1;
    std::vector<double> bin((size_t)bw * bw, 0.0);   // :73-82
    for (int i = 0; i < bw; i++) for (int j = i; j < bw; j++) bin[(size_t)j * bw + i] = choose(j, i);

    const TableLayout& lay = out.lay;
    out.rows.assign((size_t)out.n_sets * n_max * lay.row_doubles, 0.0);
    out.y.assign((size_t)out.n_sets * n_max * kQ, 0.0);
    const double* t = out.quad.data(); const double* w = out.w.data();

    auto build_range = [&](int lo, int hi) {
        for (int idx = lo; idx < hi; idx++) {
            const int s = idx / n_max, n = idx % n_max + 1;
            double* row = out.rows.data() + (size_t)idx * lay.row_doubles;
            double* y = out.y.data() + (size_t)idx * kQ;
            transition_row(Q, t, w, r_rates[rep[s]], (unsigned)n, y, row + lay.off_zlow, row + lay.off_g, row + lay.off_zdiag);
            for (int i = 0; i < kQ; i++) row[lay.off_yw + i] = y[i] * w[i];
            emission_table(Q, t, w, bin, bw, thetas[rep[s]], (unsigned)n, LA, row + lay.off_EA);
            emission_table(Q, t, w, bin, bw, thetas[rep[s]], (unsigned)n, LB, row + lay.off_EB);
            for (int d = 0; d <= LA; d++) { double a = 0; for (int i = 0; i < kQ; i++) a += w[i] * row[lay.off_EA + d * kEStride + i]; row[lay.off_wEA + d] = a; }
            for (int d = 0; d <= LB; d++) { double a = 0; for (int i = 0; i < kQ; i++) a += w[i] * row[lay.off_EB + d * kEStride + i]; row[lay.off_wEB + d] = a; }
        }
    };
    const int total = out.n_sets * n_max;
    unsigned hw = std::thread::hardware_concurrency();
    int nthr = (int)std::min<unsigned>(hw ? hw : 1, 16);
    if (total < 64 || nthr <= 1) {
        build_range(0, total);
    } else {
        std::vector<std::thread> th;
        for (int k = 0; k < nthr; k++) th.emplace_back(build_range, (int)((long)total * k / nthr), (int)((long)total * (k + 1) / nthr));
        for (auto& x : th) x.join();
    }
    return true;
}

}  // namespace coalgpu
