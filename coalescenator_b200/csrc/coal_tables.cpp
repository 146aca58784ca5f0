#include "coal_tables.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>

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

// y[i], and z in structured form, ConditionalSamplingHMM.cpp:85-182. The reference evaluates exp(-t[i]/c1),
// exp(-t[i]/c3) and exp(-t[i]) at both ends of every interval; each distinct argument is evaluated once here
// (e1, e3, e0) -- same arguments, same values, a quarter of the calls.
void transition_row(int Q, const double* t, const double* w, const double* e0 /*exp(-t[i])*/, double r, unsigned n,
                    double* y, double* zlow, double* g, double* zdiag) {
    const double c1 = n / (r + n);
    double e1[kQ + 1];
    for (int i = 0; i <= Q; i++) e1[i] = std::exp(-t[i] / c1);
    if (nearly_equal(r, n)) {
        for (int i = 0; i < Q; i++) {
            y[i] = c1 / w[i] * (e1[i] - e1[i + 1]);
            const double core = (w[i] + (t[i] * e0[i] - t[i + 1] * e0[i + 1]));
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
        double e3[kQ + 1];
        for (int i = 0; i <= Q; i++) e3[i] = std::exp(-t[i] / c3);
        for (int i = 0; i < Q; i++) {
            y[i] = c1 / w[i] * (e1[i] - e1[i + 1]);
            const double inner = (w[i] - c3 * (e3[i] - e3[i + 1]));
            zlow[i] = c2 * inner;           // :156
            g[i] = c2 / w[i] * inner;       // :160 without the w[j] factor
            const double term1 = w[i] * inner;
            const double term2 = c1 / c2 * (e1[i] - e1[i + 1]);
            const double term3 = c3 * (e0[i] * e3[i + 1] - e0[i + 1] * e3[i]);
            zdiag[i] = c2 / w[i] * (term1 - term2 - term3);   // :164-168
        }
    }
}

// E[i][d], ConditionalSamplingHMM.cpp:184-226. The reference evaluates, per interval i and distance j,
//   e = sum_l ( sum_k bin(L-j,k) / rate(k+l) * (ex(k+l,i-1) - ex(k+l,i)) ) * bin(j,l) * (-1)^l.
// The exponentials depend on (k+l, i) and the quotient bin/rate on (j,k,l) only, so they are hoisted out of the loops
// they do not depend on; every sum still adds the same values in the same order (k inner, l outer, per interval), hence
// the table is bit-identical to the reference's while the interval loop is the innermost, vectorisable one.
void emission_table(int Q, const double* t, const double* w, const std::vector<double>& bin, int bw, double theta,
                    unsigned n, int L, double* E /*[L+1][kEStride]*/) {
    std::vector<double> rate(L + 1), diff((size_t)(L + 1) * kQ, 0.0);
    double ex[kQ + 1];
    for (int m = 0; m <= L; m++) {
        rate[m] = (2 * theta * (unsigned)m / n) + 1;
        for (int i = 0; i <= Q; i++) ex[i] = std::exp(-rate[m] * t[i]);
        for (int i = 0; i < Q; i++) diff[(size_t)m * kQ + i] = ex[i] - ex[i + 1];
    }
    const double two_L = std::pow(2, (double)L);
    for (int j = 0; j <= L; j++) {
        double e[kQ], inner[kQ];
        for (int i = 0; i < kQ; i++) e[i] = 0;
        for (int l = 0; l <= j; l++) {
            for (int i = 0; i < kQ; i++) inner[i] = 0;
            for (int k = 0; k <= L - j; k++) {
                const double c = bin[(size_t)(L - j) * bw + k] / rate[k + l];
                const double* d = &diff[(size_t)(k + l) * kQ];
                for (int i = 0; i < kQ; i++) inner[i] += c * d[i];
            }
            const double b = bin[(size_t)j * bw + l], sgn = (l & 1) ? -1.0 : 1.0;
            for (int i = 0; i < kQ; i++) e[i] += inner[i] * b * sgn;
        }
        for (int i = 0; i < Q; i++) {
            double v = e[i] / (w[i] * two_L);
            v = std::max(v, std::numeric_limits<double>::epsilon());
            E[(size_t)j * kEStride + i] = v;   // distance-major: the intervals of one distance are contiguous
        }
    }
}

// ---- stable emission mode ----------------------------------------------------------------
// The double sum above is an alternating binomial sum. In double precision its absolute error is ~1e-17: entries above
// ~1e-6 are good to 1e-10 even at 30 sites per locus, entries near and below 1e-13 are rounding noise or the clamp.
// Summed in closed form it is
//   sum_l sum_k C(L-d,k) C(d,l) (-1)^l e^{-(a(k+l)+1)t} = e^{-t} (1 + e^{-at})^{L-d} (1 - e^{-at})^d,   a = 2 theta / n,
// a positive integrand, so E[i][d] = 1/(w_i 2^L) * integral over interval i of it -- evaluated here with 32-point
// Gauss-Legendre on the finite intervals and 32-point Gauss-Laguerre on the last, [t_{Q-1}, inf): relative error
// <= 1e-14 on every entry against 300-digit arithmetic for L up to 40 (tests/test_tables_and_math_host.py).
// Opt-in (coalgpu_set_emission_mode): the default stays bit-identical to the reference.
constexpr double kGLx[32] = {-0.9972638618494816, -0.9856115115452684, -0.9647622555875064, -0.9349060759377397, -0.8963211557660521, -0.84936761373257, -0.7944837959679424, -0.7321821187402897, -0.6630442669302152, -0.5877157572407623, -0.5068999089322294, -0.42135127613063533, -0.33186860228212767, -0.23928736225213706, -0.1444719615827965, -0.048307665687738324, 0.048307665687738324, 0.1444719615827965, 0.23928736225213706, 0.33186860228212767, 0.42135127613063533, 0.5068999089322294, 0.5877157572407623, 0.6630442669302152, 0.7321821187402897, 0.7944837959679424, 0.84936761373257, 0.8963211557660521, 0.9349060759377397, 0.9647622555875064, 0.9856115115452684, 0.9972638618494816};
constexpr double kGLw[32] = {0.007018610009470506, 0.016274394730905743, 0.025392065309262024, 0.034273862913021765, 0.042835898022226836, 0.05099805926237609, 0.058684093478535565, 0.06582222277636168, 0.07234579410884834, 0.07819389578707023, 0.08331192422694671, 0.08765209300440378, 0.09117387869576378, 0.09384439908080451, 0.09563872007927471, 0.09654008851472766, 0.09654008851472766, 0.09563872007927471, 0.09384439908080451, 0.09117387869576378, 0.08765209300440378, 0.08331192422694671, 0.07819389578707023, 0.07234579410884834, 0.06582222277636168, 0.058684093478535565, 0.05099805926237609, 0.042835898022226836, 0.034273862913021765, 0.025392065309262024, 0.016274394730905743, 0.007018610009470506};
constexpr double kLagx[32] = {0.04448936583326739, 0.23452610951961825, 0.5768846293018863, 1.072448753817818, 1.7224087764446456, 2.5283367064257942, 3.492213273021994, 4.616456769749767, 5.903958504174244, 7.358126733186241, 8.982940924212597, 10.78301863253997, 12.763697986742725, 14.931139755522558, 17.292454336715316, 19.855860940336054, 22.630889013196775, 25.628636022459247, 28.862101816323474, 32.346629153964734, 36.10049480575197, 40.14571977153944, 44.50920799575494, 49.22439498730864, 54.33372133339691, 59.89250916213402, 65.97537728793505, 72.68762809066271, 80.18744697791352, 88.7353404178924, 98.82954286828398, 111.7513980979377};
constexpr double kLagw[32] = {0.10921834195242515, 0.2104431079387924, 0.23521322966984298, 0.19590333597287354, 0.1299837862860684, 0.07057862386571556, 0.03176091250917397, 0.011918214834838204, 0.003738816294611417, 0.0009808033066149246, 0.0002148649188013578, 3.920341967987801e-05, 5.934541612868448e-06, 7.41640457866734e-07, 7.604567879120552e-08, 6.350602226625618e-09, 4.281382971040738e-10, 2.305899491891237e-11, 9.799379288726772e-13, 3.2378016577291567e-14, 8.171823443420549e-16, 1.542133833393811e-17, 2.119792290163547e-19, 2.0544296737880036e-21, 1.3469825866373617e-23, 5.661294130397462e-26, 1.4185605454629684e-28, 1.9133754944539943e-31, 1.1922487600982012e-34, 2.6715112192396625e-38, 1.3386169421064243e-42, 4.5105361938987784e-48};

inline double stable_integrand(double a, int L, int d, double t) {
    const double x = std::exp(-a * t);
    return std::exp((double)(L - d) * std::log1p(x) + (d > 0 ? (double)d * std::log1p(-x) : 0.0));
}

void emission_table_stable(int Q, const double* t, const double* w, double theta, unsigned n, int L, double* E /*[L+1][kEStride]*/) {
    const double a = 2 * theta / n;
    const double two_L = std::pow(2, (double)L);
    for (int d = 0; d <= L; d++) {
        for (int i = 0; i < Q; i++) {
            double v = 0;
            if (i + 1 < Q || t[Q] < 1e299) {
                const double lo = t[i], hi = t[i + 1], h = 0.5 * (hi - lo), c = 0.5 * (hi + lo);
                for (int q = 0; q < 32; q++) { const double tt = c + h * kGLx[q]; v += kGLw[q] * std::exp(-tt) * stable_integrand(a, L, d, tt); }
                v *= h;
            } else {
                const double lo = t[i];
                for (int q = 0; q < 32; q++) v += kLagw[q] * stable_integrand(a, L, d, lo + kLagx[q]);
                v *= std::exp(-lo);
            }
            v /= (w[i] * two_L);
            E[(size_t)d * kEStride + i] = std::max(v, std::numeric_limits<double>::epsilon());   // the reference's clamp (:221)
        }
    }
}

// ---- emission cache ---------------------------------------------------------------------------------------------
// E depends on (quadrature, L, theta, n) only -- not on the locus pair: theta = 4 lambda V mu is the same for every pair
// of a data set and the pairs of main.cpp:242-311 have a handful of distinct L. The blocks are therefore kept in a
// process-wide cache, so that a batch of pairs builds each of them once (the host table
// builder is the serial section that shows once sampling is fast). Values are the same doubles either way.
// Readers never look at a buffer that can move: a block that has to grow is rebuilt into a NEW vector (old rows copied,
// new rows computed) which is then published under the block's mutex; a reader holds the shared_ptr of the snapshot it
// was given for as long as it copies rows out of it (build_tables), whatever later callers with a larger n_max do.
struct EmissionBlock {
    std::mutex m;
    int stride = 0;                  // doubles per n: (L+1)*kEStride table + (L+1) weighted sums
    std::shared_ptr<const std::vector<double>> data;   // rows for n = 1..data->size()/stride
};
struct EmissionRows {                // an immutable snapshot of a block that covers n = 1..n_max of the request
    std::shared_ptr<const std::vector<double>> data;
    int stride = 0;
    const double* row(int n) const { return data->data() + (size_t)(n - 1) * stride; }
};
struct EmissionKey {
    int Q, L, mode; uint64_t theta_bits;
    bool operator<(const EmissionKey& o) const { return std::tie(Q, L, mode, theta_bits) < std::tie(o.Q, o.L, o.mode, o.theta_bits); }
};
std::mutex g_cache_mutex;
std::map<EmissionKey, std::shared_ptr<EmissionBlock>> g_cache;
size_t g_cache_doubles = 0;
constexpr size_t kCacheMaxDoubles = (size_t)1 << 26;   // 512 MiB; the cache is dropped when it grows past this

EmissionRows emission_block(int Q, int L, int mode, double theta, int n_max, const double* t, const double* w,
                            const std::vector<double>& bin, int bw, int threads) {
    EmissionKey key{Q, L, mode, 0};
    std::memcpy(&key.theta_bits, &theta, sizeof(double));
    std::shared_ptr<EmissionBlock> b;
    {
        std::lock_guard<std::mutex> lk(g_cache_mutex);
        auto it = g_cache.find(key);
        if (it == g_cache.end()) {
            if (g_cache_doubles > kCacheMaxDoubles) { g_cache.clear(); g_cache_doubles = 0; }
            b = std::make_shared<EmissionBlock>();
            b->stride = (L + 1) * kEStride + (L + 1);
            g_cache[key] = b;
        } else {
            b = it->second;
        }
    }
    std::lock_guard<std::mutex> lk(b->m);
    const int n_built = b->data ? (int)(b->data->size() / (size_t)b->stride) : 0;
    if (n_built < n_max) {
        const int lo = n_built;
        auto grown = std::make_shared<std::vector<double>>((size_t)n_max * b->stride, 0.0);
        if (lo) std::memcpy(grown->data(), b->data->data(), sizeof(double) * (size_t)lo * b->stride);
        double* const base = grown->data();
        const int stride = b->stride;
        auto fill = [&](int a, int e) {
            for (int n = a + 1; n <= e; n++) {
                double* E = base + (size_t)(n - 1) * stride;
                if (mode == kEmissionStable) emission_table_stable(Q, t, w, theta, (unsigned)n, L, E);
                else emission_table(Q, t, w, bin, bw, theta, (unsigned)n, L, E);
                double* wE = E + (L + 1) * kEStride;
                for (int d = 0; d <= L; d++) { double acc = 0; for (int i = 0; i < kQ; i++) acc += w[i] * E[d * kEStride + i]; wE[d] = acc; }
            }
        };
        const int total = n_max - lo;
        const int nthr = std::max(1, std::min(threads, total / 16));
        if (nthr <= 1) {
            fill(lo, n_max);
        } else {
            std::vector<std::thread> th;
            for (int k = 0; k < nthr; k++) th.emplace_back(fill, lo + (int)((long)total * k / nthr), lo + (int)((long)total * (k + 1) / nthr));
            for (auto& x : th) x.join();
        }
        {
            std::lock_guard<std::mutex> lk2(g_cache_mutex);
            g_cache_doubles += (size_t)total * b->stride;
        }
        b->data = std::move(grown);
    }
    return EmissionRows{b->data, b->stride};
}

}  // namespace

void clear_table_cache() {
    std::lock_guard<std::mutex> lk(g_cache_mutex);
    g_cache.clear(); g_cache_doubles = 0;
}

bool build_tables(int Q, int LA, int LB, int n_max, const std::vector<double>& thetas,
                  const std::vector<double>& r_rates, HostTables& out, int threads, int emission_mode) {
    if ((Q != 4 && Q != 8 && Q != kQ) || LA < 1 || LB < 1 || n_max < 1 || thetas.size() != r_rates.size()) return false;
    if (threads <= 0) { const unsigned hw = std::thread::hardware_concurrency(); threads = (int)std::min<unsigned>(hw ? hw : 1, 16); }
    out.lay.init(LA, LB);
    out.n_max = n_max;
    // Q = 4 or 8 (ConditionalSamplingHMM.cpp:39-47): the kernels always sweep 16 intervals; intervals beyond Q carry
    // weight 0 and all-zero table entries, so they contribute nothing.
    out.Q = Q;
    const double* nodes = Q == 4 ? kNodes4 : (Q == 8 ? kNodes8 : kNodes16);
    out.quad.assign(kQ + 1, 1E+300);
    for (int i = 0; i <= Q; i++) out.quad[i] = nodes[i];
    out.w.assign(kQ, 0.0);
    double e0[kQ + 1];
    for (int i = 0; i <= kQ; i++) e0[i] = std::exp(-out.quad[i]);
    for (int i = 1; i <= Q; i++) out.w[i - 1] = e0[i - 1] - e0[i];   // :63-71

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

    for (int s = 0; s < out.n_sets; s++) {
        const double theta = thetas[rep[s]], r = r_rates[rep[s]];
        const bool emit = emission_mode != kEmissionNone;
        const EmissionRows bA = emit ? emission_block(Q, LA, emission_mode, theta, n_max, t, w, bin, bw, threads) : EmissionRows();
        const EmissionRows bB = !emit ? EmissionRows() : (LB == LA) ? bA : emission_block(Q, LB, emission_mode, theta, n_max, t, w, bin, bw, threads);
        for (int n = 1; n <= n_max; n++) {
            const size_t idx = (size_t)s * n_max + (n - 1);
            double* row = out.rows.data() + idx * lay.row_doubles;
            double* y = out.y.data() + idx * kQ;
            transition_row(Q, t, w, e0, r, (unsigned)n, y, row + lay.off_zlow, row + lay.off_g, row + lay.off_zdiag);
            for (int i = 0; i < kQ; i++) row[lay.off_yw + i] = y[i] * w[i];
            if (!emit) continue;
            const double* a = bA.row(n);
            const double* b = bB.row(n);
            std::memcpy(row + lay.off_EA, a, sizeof(double) * (LA + 1) * kEStride);
            std::memcpy(row + lay.off_EB, b, sizeof(double) * (LB + 1) * kEStride);
            std::memcpy(row + lay.off_wEA, a + (LA + 1) * kEStride, sizeof(double) * (LA + 1));
            std::memcpy(row + lay.off_wEB, b + (LB + 1) * kEStride, sizeof(double) * (LB + 1));
        }
    }
    // device rows: the bilinear form of the two-locus forward pass (coal_math.h) + the one-locus weight vectors
    out.dev_rows.clear();
    if (emission_mode != kEmissionNone) {
        out.dev_rows.assign((size_t)out.n_sets * n_max * lay.dev_row_doubles, 0.0);
        for (size_t idx = 0; idx < (size_t)out.n_sets * n_max; idx++) {
            const double* row = out.rows.data() + idx * lay.row_doubles;
            double* drow = out.dev_rows.data() + idx * lay.dev_row_doubles;
            for (int dA = 0; dA <= LA + 1; dA++)
                bilinear_row(row, w, lay.off_yw, lay.off_zlow, lay.off_g, lay.off_zdiag, lay.off_EA, lay.off_EB, LA, LB, dA, drow + (size_t)2 * dA * lay.dev_sb);
            std::memcpy(drow + lay.dev_off_wEA, row + lay.off_wEA, sizeof(double) * (LA + 1));
            std::memcpy(drow + lay.dev_off_wEB, row + lay.off_wEB, sizeof(double) * (LB + 1));
        }
    }
    return true;
}

}  // namespace coalgpu
