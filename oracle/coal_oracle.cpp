// TEST INFRASTRUCTURE — CPU oracle (see coal_oracle.h for scope and parity status).
//
// A from-scratch restatement of the reference's importance-sampling likelihood path. Every function
// names the reference lines it follows. STRICT mode keeps the reference's floating-point operation
// order, its std::mt19937 draw order and (through std::unordered_map with the same hash values and
// the same insert/erase sequence) its container iteration order, so whole runs reproduce the
// compiled reference bit for bit. CANONICAL mode swaps in the two conventions of the CUDA path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.

#include "coal_oracle.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

typedef unsigned int uint;

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------

// boost::math::binomial_coefficient<double> semantics (value rounded to the nearest integer), as
// used at ConditionalSamplingHMM.cpp:79 and ProposalEngine.cpp:73-82. Same evaluation as the shim
// the compiled reference is built against (oracle/ref_shim/boost/math/special_functions/binomial.hpp).
double binom(uint n, uint k) {
    if (k > n) return std::nan("");
    if (k == 0 || k == n) return 1.0;
    if (k == 1 || k == n - 1) return (double)n;
    uint kk = (k < n - k) ? k : n - k;
    long double r = 1.0L;
    for (uint i = 1; i <= kk; i++) r = r * (long double)(n - kk + i) / (long double)i;
    return std::ceil((double)r - 0.5);
}

double lgamma_safe(double x) { int s; return ::lgamma_r(x, &s); }

// Utils.hpp:88-107
double log_addition(double a, double b) {
    if (a < b) return b + std::log(1 + std::exp(a - b));
    return a + std::log(1 + std::exp(b - a));
}

// Utils.hpp:206-209
bool double_compare(double a, double b) {
    return (a == b) || (std::fabs(a - b) < std::fabs(std::min(a, b)) * std::numeric_limits<double>::epsilon());
}

// Philox4x32-10 (Salmon et al. 2011), the counter-based generator of the CUDA path.
struct Philox {
    static void round(uint32_t c[4], const uint32_t k[2]) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    static void block(uint64_t seed, uint64_t hist, uint64_t blk, uint32_t out[4]) {
        uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        uint32_t c[4] = {(uint32_t)hist, (uint32_t)(hist >> 32), (uint32_t)blk, (uint32_t)(blk >> 32)};
        for (int r = 0; r < 10; r++) {
            round(c, k);
            k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
        }
        for (int i = 0; i < 4; i++) out[i] = c[i];
    }
    // draw k of history hist: 52 bits -> (0,1). (bits + 0.5) needs 53 significant bits, so every value is exact,
    // the smallest is 2^-53 and the largest 1 - 2^-53: never 0 and never 1 (the reference asserts a positive waiting time).
    static double uniform(uint64_t seed, uint64_t hist, uint64_t k) {
        uint32_t o[4];
        block(seed, hist, k >> 1, o);
        const uint32_t lo = o[(k & 1) * 2], hi = o[(k & 1) * 2 + 1];
        const uint64_t bits = (((uint64_t)hi << 32) | lo) >> 12;
        return ((double)bits + 0.5) * (1.0 / 4503599627370496.0);
    }
};

struct Rng {
    virtual ~Rng() {}
    virtual double uniform() = 0;
    virtual void begin_history(uint64_t hist) = 0;
};
// ImportanceSampler.cpp:67 + std::uniform_real_distribution<double>(0,1) (ProposalEngine.cpp:49,
// TypeContainer.cpp:45): one shared engine, stateless distributions.
struct MtRng : Rng {
    std::mt19937 eng;
    std::uniform_real_distribution<double> dist{0.0, 1.0};
    explicit MtRng(uint seed) : eng(seed) {}
    double uniform() override { return dist(eng); }
    void begin_history(uint64_t) override {}
};
struct PhiloxRng : Rng {
    uint64_t seed, hist = 0, k = 0;
    explicit PhiloxRng(uint64_t s) : seed(s) {}
    double uniform() override { return Philox::uniform(seed, hist, k++); }
    void begin_history(uint64_t h) override { hist = h; k = 0; }
};

// ------------------------------------------------------------------------------------------------
// problem
// ------------------------------------------------------------------------------------------------

struct TypeRec { int gamma; uint64_t word; uint count; };

struct Ctx {
    int T, LA, LB, Q, G;
    std::vector<double> times, viral, mu, rho, lambda;
    double tau, dmu, dlambda, drho;
    std::vector<std::vector<TypeRec>> types;   // per time point, insertion order
    uint64_t maskA, maskB;
    // HMM constants (ConditionalSamplingHMM.cpp:36-83)
    std::vector<double> quad, w, r_rates, thetas;
    std::vector<double> ebin;                  // (maxL+1)^2, ebin[j*(maxL+1)+i] = C(j,i)
    int maxL;
    struct Trans { std::vector<double> y, z; };
    std::map<std::pair<uint, uint>, Trans> trans;                 // (n, c)
    std::map<std::vector<uint>, std::vector<double>> emis;        // (n, c, L) -> [Q][L+1]
};

const double QUAD4[] = {0, 0.415774556783, 2.294280360279, 6.289945082937, 1E+300};
const double QUAD8[] = {0, 0.193043676560, 1.026664895339, 2.567876744951, 4.900353084526, 8.182153444563, 12.7344180291798, 19.3957278622, 1E+300};
const double QUAD16[] = {0, 0.093307812017, 0.492691740302, 1.215595412071, 2.269949526204, 3.667622721751, 5.425336627414, 7.565916226613, 10.120228568019, 13.130282482176, 16.6544077083, 20.7764788994, 25.6238942267, 31.407519169754, 38.530683306486, 48.026085572686, 1E+300};

bool init_ctx(Ctx& c, const coalo_problem* p, int Q) {
    c.T = p->T; c.LA = p->LA; c.LB = p->LB; c.Q = Q;
    if (c.LA + c.LB > 64 || c.LA < 1 || c.LB < 1 || c.T < 1) return false;
    c.times.assign(p->times, p->times + p->T);
    c.viral.assign(p->viral, p->viral + p->T);
    c.mu.assign(p->mu, p->mu + p->n_mu);
    c.rho.assign(p->rho, p->rho + p->n_rho);
    c.lambda.assign(p->lambda, p->lambda + p->n_lambda);
    c.G = p->n_mu * p->n_rho * p->n_lambda;
    c.tau = p->tau; c.dmu = p->driving_mu; c.dlambda = p->driving_lambda; c.drho = p->driving_rho;
    c.maskA = (c.LA == 64) ? ~0ULL : ((1ULL << c.LA) - 1);
    c.maskB = ((c.LB == 64) ? ~0ULL : ((1ULL << c.LB) - 1)) << c.LA;
    c.types.resize(p->T);
    for (int t = 0; t < p->T; t++)
        for (int k = p->type_offsets[t]; k < p->type_offsets[t + 1]; k++)
            c.types[t].push_back(TypeRec{p->type_gamma[k], p->type_words[k], p->type_counts[k]});
    // quadrature (ConditionalSamplingHMM.cpp:38-58)
    const double* q = nullptr;
    if (Q == 4) q = QUAD4; else if (Q == 8) q = QUAD8; else if (Q == 16) q = QUAD16; else return false;
    c.quad.assign(q, q + Q + 1);
    // weights (ConditionalSamplingHMM.cpp:63-71)
    c.w.resize(Q);
    for (int i = 1; i <= Q; i++) c.w[i - 1] = std::exp(-c.quad[i - 1]) - std::exp(-c.quad[i]);
    // scaled driving rates per epoch (ImportanceSampler.cpp:69-76)
    c.thetas.resize(c.T); c.r_rates.resize(c.T);
    for (int i = 0; i < c.T; i++) {
        c.thetas[i] = 4 * c.dlambda * c.viral[i] * c.dmu;
        c.r_rates[i] = 4 * c.dlambda * c.viral[i] * c.drho;
    }
    // e_binomial (ConditionalSamplingHMM.cpp:73-82)
    c.maxL = std::max(c.LA, c.LB);
    const int m = c.maxL + 1;
    c.ebin.assign((size_t)m * m, 0.0);
    for (int i = 0; i < m; i++) for (int j = i; j < m; j++) c.ebin[(size_t)j * m + i] = binom(j, i);
    return true;
}

// ConditionalSamplingHMM::calcConstantsTransition, ConditionalSamplingHMM.cpp:85-182
const Ctx::Trans& transition(Ctx& c, uint n_all, uint time_idx) {
    auto key = std::make_pair(n_all, time_idx);
    auto it = c.trans.find(key);
    if (it != c.trans.end()) return it->second;
    const int Q = c.Q;
    const std::vector<double>& qd = c.quad; const std::vector<double>& w = c.w;
    Ctx::Trans tr; tr.y.assign(Q, 0.0); tr.z.assign((size_t)Q * Q, 0.0);
    const double r = c.r_rates[time_idx];
    const double c1 = n_all / (r + n_all);
    if (double_compare(r, n_all)) {
        for (int i = 0; i < Q; i++) {
            tr.y[i] = c1 / w[i] * (std::exp(-qd[i] / c1) - std::exp(-qd[i + 1] / c1));
            for (int j = 0; j < Q; j++) {
                double v;
                if (i > j) {
                    v = (w[j] + (qd[j] * std::exp(-qd[j]) - qd[j + 1] * std::exp(-qd[j + 1])));
                } else if (i < j) {
                    v = w[j] / w[i] * (w[i] + (qd[i] * std::exp(-qd[i]) - qd[i + 1] * std::exp(-qd[i + 1])));
                } else {
                    double term1 = w[i] * (w[i] + (qd[i] * std::exp(-qd[i]) - qd[i + 1] * std::exp(-qd[i + 1])));
                    // the reference writes 1/2 in integer arithmetic (== 0), ConditionalSamplingHMM.cpp:123
                    double term2 = -(qd[i] - qd[i + 1]) * std::exp(-(qd[i] + qd[i + 1])) - (1 / 2) * (std::exp(-2 * qd[i]) - std::exp(-2 * qd[i + 1]));
                    v = 1 / w[i] * (term1 + term2);
                }
                tr.z[(size_t)i * Q + j] = v;
            }
        }
    } else {
        const double c2 = r / (r - n_all);
        const double c3 = n_all / r;
        for (int i = 0; i < Q; i++) {
            tr.y[i] = c1 / w[i] * (std::exp(-qd[i] / c1) - std::exp(-qd[i + 1] / c1));
            for (int j = 0; j < Q; j++) {
                double v;
                if (i > j) {
                    v = c2 * (w[j] - c3 * (std::exp(-qd[j] / c3) - std::exp(-qd[j + 1] / c3)));
                } else if (i < j) {
                    v = c2 * w[j] / w[i] * (w[i] - c3 * (std::exp(-qd[i] / c3) - std::exp(-qd[i + 1] / c3)));
                } else {
                    double term1 = w[i] * (w[i] - c3 * (std::exp(-qd[i] / c3) - std::exp(-qd[i + 1] / c3)));
                    double term2 = c1 / c2 * (std::exp(-qd[i] / c1) - std::exp(-qd[i + 1] / c1));
                    double term3 = c3 * (std::exp(-qd[i]) * std::exp(-qd[i + 1] / c3) - std::exp(-qd[i + 1]) * std::exp(-qd[i] / c3));
                    v = c2 / w[i] * (term1 - term2 - term3);
                }
                tr.z[(size_t)i * Q + j] = v;
            }
        }
    }
    return c.trans.emplace(key, std::move(tr)).first->second;
}

// ConditionalSamplingHMM::calcConstantsEmission, ConditionalSamplingHMM.cpp:184-226
const std::vector<double>& emission(Ctx& c, uint n_all, uint time_idx, uint L) {
    std::vector<uint> key = {n_all, time_idx, L};
    auto it = c.emis.find(key);
    if (it != c.emis.end()) return it->second;
    const int Q = c.Q; const int m = c.maxL + 1;
    std::vector<double> E((size_t)Q * (L + 1), 0.0);
    const double theta = c.thetas[time_idx];
    for (int i = 1; i <= Q; i++) {
        for (uint j = 0; j <= L; j++) {
            double e_prob = 0;
            for (uint l = 0; l <= j; l++) {
                double inner = 0;
                for (uint k = 0; k <= (L - j); k++) {
                    double rate = (2 * theta * (k + l) / n_all) + 1;
                    inner += c.ebin[(size_t)(L - j) * m + k] / rate * (std::exp(-rate * c.quad[i - 1]) - std::exp(-rate * c.quad[i]));
                }
                e_prob += inner * c.ebin[(size_t)j * m + l] * std::pow(-1, (double)l);
            }
            e_prob /= (c.w[i - 1] * std::pow(2, (double)L));
            e_prob = std::max(e_prob, std::numeric_limits<double>::epsilon());
            E[(size_t)(i - 1) * (L + 1) + j] = e_prob;
        }
    }
    return c.emis.emplace(key, std::move(E)).first->second;
}

// ------------------------------------------------------------------------------------------------
// pi_SMC from class vectors — ConditionalSamplingHMM::piSMC, ConditionalSamplingHMM.cpp:265-448
// ------------------------------------------------------------------------------------------------

struct Classes { std::vector<int> cnt, dA, dB; };

// ConditionalSamplingHMM.cpp:234-263
double partial_emission(const std::vector<double>& E, uint L, int i, const std::vector<int>& cnt, const std::vector<int>& d) {
    uint complete = 0; double imp = 0;
    for (size_t k = 0; k < cnt.size(); k++) {
        if (d[k] >= 0) { imp += cnt[k] * E[(size_t)i * (L + 1) + d[k]]; complete += cnt[k]; }
    }
    if (complete == 0) imp = 1 / std::pow(2, (double)L); else imp /= complete;
    return imp;
}

double pismc_classes(Ctx& c, int gamma, uint n, uint tidx, const Classes& k) {
    const int Q = c.Q; const uint LA = c.LA, LB = c.LB;
    if (k.cnt.empty()) {   // ConditionalSamplingHMM.cpp:277-298
        if (gamma == COALO_GAMMA_A) return 1 / std::pow(2, (double)LA);
        if (gamma == COALO_GAMMA_B) return 1 / std::pow(2, (double)LB);
        return 1 / std::pow(2, (double)(LA + LB));
    }
    const Ctx::Trans& tr = transition(c, n, tidx);
    const std::vector<double>& EA = emission(c, n, tidx, LA);
    const std::vector<double>& EB = emission(c, n, tidx, LB);
    const std::vector<double>& w = c.w;
    double pi = 0;
    if (gamma == COALO_GAMMA_A) {          // :332-350
        for (int i = 0; i < Q; i++) for (size_t a = 0; a < k.dA.size(); a++) {
            double init = k.cnt[a] / (double)n * w[i];
            if (k.dA[a] == -1) pi += init * partial_emission(EA, LA, i, k.cnt, k.dA);
            else pi += init * EA[(size_t)i * (LA + 1) + k.dA[a]];
        }
    } else if (gamma == COALO_GAMMA_B) {   // :352-370
        for (size_t b = 0; b < k.dB.size(); b++) for (int j = 0; j < Q; j++) {
            double init = k.cnt[b] / (double)n * w[j];
            if (k.dB[b] == -1) pi += init * partial_emission(EB, LB, j, k.cnt, k.dB);
            else pi += init * EB[(size_t)j * (LB + 1) + k.dB[b]];
        }
    } else {                               // :372-424
        const size_t nH = k.cnt.size();
        std::vector<double> base((size_t)Q * nH, 0.0), bsum(Q, 0.0);
        for (int i = 0; i < Q; i++) for (size_t a = 0; a < nH; a++) {
            double init = k.cnt[a] / (double)n * w[i];
            double v = (k.dA[a] == -1) ? partial_emission(EA, LA, i, k.cnt, k.dA) * init
                                       : EA[(size_t)i * (LA + 1) + k.dA[a]] * init;
            base[(size_t)i * nH + a] = v; bsum[i] += v;
        }
        for (size_t b = 0; b < nH; b++) for (int j = 0; j < Q; j++) {
            double F2 = 0;
            for (int i = 0; i < Q; i++) F2 += bsum[i] * tr.z[(size_t)i * Q + j];
            F2 *= k.cnt[b] / (double)n;
            F2 += tr.y[j] * base[(size_t)j * nH + b];
            if (k.dB[b] == -1) pi += F2 * partial_emission(EB, LB, j, k.cnt, k.dB);
            else pi += F2 * EB[(size_t)j * (LB + 1) + k.dB[b]];
        }
    }
    return pi;
}

// ------------------------------------------------------------------------------------------------
// type store (TypeContainer, TypeContainer.cpp:42-916) behind one interface with two orderings
// ------------------------------------------------------------------------------------------------

struct Hap { int gamma; uint64_t word; };
struct TempOp { bool add; Hap h; };

// Utils.hpp:60-76 applied to the vector<bool> the reference would hold for this packed word
struct BitHash {
    int off, len;
    size_t operator()(uint64_t wd) const {
        size_t seed = 0;
        for (int s = 0; s < len; s++) {
            size_t b = (wd >> (off + s)) & 1ULL;
            seed ^= b + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        }
        return seed;
    }
};
typedef std::unordered_map<uint64_t, uint, BitHash> HMap;

struct Tracer;

struct Store {
    const Ctx& cx;
    uint nA = 0, nB = 0, nC = 0;
    Tracer* tr = nullptr;
    explicit Store(const Ctx& c) : cx(c) {}
    virtual ~Store() {}
    uint n() const { return nA + nB + nC; }
    uint64_t partA(uint64_t w) const { return w & cx.maskA; }
    uint64_t partB(uint64_t w) const { return w & cx.maskB; }

    virtual void reset(const std::vector<TypeRec>& init) = 0;
    virtual void add(const Hap& h) = 0;       // TypeContainer::addType   :409-456
    virtual void remove(const Hap& h) = 0;    // TypeContainer::removeType :459-522
    virtual void commit() = 0;                // end of event (canonical: drop empty slots)
    virtual void push_temp(const TempOp& op) = 0;   // temporary add/remove around a piSMC call
    virtual void pop_temp(const TempOp& op) = 0;
    // counts of the literal maps (getTypeOccurrence*, :261-406); words already restricted to the locus
    virtual uint cntA(uint64_t wa) const = 0;
    virtual uint cntB(uint64_t wb) const = 0;
    virtual uint cntC(uint64_t wc) const = 0;
    virtual uint cntCa(uint64_t wa) const = 0;
    virtual uint cntCb(uint64_t wb) const = 0;
    // addTypeMap :525-591; returns (old+new, new) pairs in insertion order
    virtual std::vector<std::pair<uint, uint>> add_map(int t) = 0;
    // sampleType iteration order (:648-721) and getHaplotypesA/B order (:604-638)
    virtual void sampling_order(std::vector<TypeRec>& out) const = 0;
    virtual void list(int gamma, std::vector<uint64_t>& out) const = 0;
    virtual Classes differences(const Hap& q) const = 0;   // getDifferences :748-894
    // probabilitySamplePermutation (ProposalEngine.cpp:88-118) iterates the INPUT maps
    virtual double permutation_constant(int t) const = 0;
    virtual uint unique_count() const = 0;
};

struct Tracer {
    FILE* f = nullptr; long hist = -1; int epoch = 0; bool on = false;
    std::map<std::string, int> delta;
    const Ctx* cx = nullptr;
    std::string key(const Hap& h) const {
        std::string s(1, "ABC"[h.gamma]); s.push_back(':');
        int off = (h.gamma == COALO_GAMMA_B) ? cx->LA : 0;
        int len = (h.gamma == COALO_GAMMA_A) ? cx->LA : (h.gamma == COALO_GAMMA_B ? cx->LB : cx->LA + cx->LB);
        for (int i = 0; i < len; i++) s.push_back(((h.word >> (off + i)) & 1) ? '1' : '0');
        return s;
    }
    void note(const Hap& h, int d) { if (on) delta[key(h)] += d; }
    void flush() {
        if (!on) return;
        bool any = false;
        for (auto& kv : delta) if (kv.second) any = true;
        if (any) {
            fprintf(f, "D %ld", hist);
            for (auto& kv : delta) if (kv.second) fprintf(f, " %+d%s", kv.second, kv.first.c_str());
            fprintf(f, "\n");
        }
        delta.clear();
    }
};

// ---- STRICT: five std::unordered_map with the reference's hash values and op sequence ------------
struct StrictStore : Store {
    HMap A, B, C, Ca, Cb;
    std::vector<HMap> inA, inB, inC;   // the Sample's input maps, built by insertion (ref_driver.cpp / Sample.cpp:632-662)
    explicit StrictStore(const Ctx& c)
        : Store(c), A(0, BitHash{0, c.LA}), B(0, BitHash{c.LA, c.LB}), C(0, BitHash{0, c.LA + c.LB}),
          Ca(0, BitHash{0, c.LA}), Cb(0, BitHash{c.LA, c.LB}) {
        for (int t = 0; t < c.T; t++) {
            HMap a(0, BitHash{0, c.LA}), b(0, BitHash{c.LA, c.LB}), cc(0, BitHash{0, c.LA + c.LB});
            for (const TypeRec& r : c.types[t]) {
                HMap& m = r.gamma == COALO_GAMMA_A ? a : (r.gamma == COALO_GAMMA_B ? b : cc);
                if (!m.insert({r.word, r.count}).second) m.at(r.word) += r.count;
            }
            inA.push_back(a); inB.push_back(b); inC.push_back(cc);
        }
    }
    static void bump(HMap& m, uint64_t k, uint by) { if (!m.insert({k, by}).second) m.at(k) += by; }
    static void drop(HMap& m, uint64_t k) { if (m.at(k) > 1) m.at(k)--; else m.erase(k); }
    // TypeContainer ctor :42-99 (copies of the input maps keep their node order)
    void reset(const std::vector<TypeRec>&) override {
        A = inA[0]; B = inB[0]; C = inC[0];
        Ca = HMap(0, BitHash{0, cx.LA}); Cb = HMap(0, BitHash{cx.LA, cx.LB});
        nA = nB = nC = 0;
        for (auto& kv : C) { nC += kv.second; bump(Ca, partA(kv.first), kv.second); bump(Cb, partB(kv.first), kv.second); }
        for (auto& kv : A) nA += kv.second;
        for (auto& kv : B) nB += kv.second;
    }
    void add(const Hap& h) override {
        if (tr) tr->note(h, +1);
        if (h.gamma == COALO_GAMMA_A) { nA++; bump(A, h.word, 1); }
        else if (h.gamma == COALO_GAMMA_B) { nB++; bump(B, h.word, 1); }
        else { nC++; bump(C, h.word, 1); bump(Ca, partA(h.word), 1); bump(Cb, partB(h.word), 1); }
    }
    void remove(const Hap& h) override {
        if (tr) tr->note(h, -1);
        if (h.gamma == COALO_GAMMA_A) { nA--; drop(A, h.word); }
        else if (h.gamma == COALO_GAMMA_B) { nB--; drop(B, h.word); }
        else { nC--; drop(C, h.word); drop(Ca, partA(h.word)); drop(Cb, partB(h.word)); }
    }
    void commit() override {}
    void push_temp(const TempOp& op) override { if (op.add) add(op.h); else remove(op.h); }
    void pop_temp(const TempOp& op) override { if (op.add) remove(op.h); else add(op.h); }
    static uint get(const HMap& m, uint64_t k) { auto it = m.find(k); return it == m.end() ? 0 : it->second; }
    uint cntA(uint64_t wa) const override { return get(A, wa); }
    uint cntB(uint64_t wb) const override { return get(B, wb); }
    uint cntC(uint64_t wc) const override { return get(C, wc); }
    uint cntCa(uint64_t wa) const override { return get(Ca, wa); }
    uint cntCb(uint64_t wb) const override { return get(Cb, wb); }
    std::vector<std::pair<uint, uint>> add_map(int t) override {
        std::vector<std::pair<uint, uint>> ov;
        for (auto& kv : inA[t]) { nA += kv.second; bump(A, kv.first, kv.second); ov.push_back({A[kv.first], kv.second}); }
        for (auto& kv : inB[t]) { nB += kv.second; bump(B, kv.first, kv.second); ov.push_back({B[kv.first], kv.second}); }
        for (auto& kv : inC[t]) {
            nC += kv.second; bump(C, kv.first, kv.second); ov.push_back({C[kv.first], kv.second});
            bump(Ca, partA(kv.first), kv.second); bump(Cb, partB(kv.first), kv.second);
        }
        return ov;
    }
    void sampling_order(std::vector<TypeRec>& out) const override {
        out.clear();
        for (auto& kv : C) out.push_back(TypeRec{COALO_GAMMA_C, kv.first, kv.second});
        for (auto& kv : A) out.push_back(TypeRec{COALO_GAMMA_A, kv.first, kv.second});
        for (auto& kv : B) out.push_back(TypeRec{COALO_GAMMA_B, kv.first, kv.second});
    }
    void list(int gamma, std::vector<uint64_t>& out) const override {
        out.clear();
        const HMap& m = gamma == COALO_GAMMA_A ? A : (gamma == COALO_GAMMA_B ? B : C);
        for (auto& kv : m) out.push_back(kv.first);
    }
    Classes differences(const Hap& q) const override {
        Classes k;
        if (q.gamma == COALO_GAMMA_A || q.gamma == COALO_GAMMA_B) {   // :754-800, :897-916
            const bool isA = q.gamma == COALO_GAMMA_A;
            std::map<uint, uint> red;
            auto scan = [&](const HMap& m) {
                for (auto& kv : m) {
                    uint d = (uint)__builtin_popcountll(kv.first ^ q.word);
                    if (!red.insert({d, kv.second}).second) red.at(d) += kv.second;
                }
            };
            scan(isA ? A : B); scan(isA ? Ca : Cb);
            for (auto it = red.rbegin(); it != red.rend(); ++it) {
                k.cnt.push_back(it->second);
                k.dA.push_back(isA ? (int)it->first : -1);
                k.dB.push_back(isA ? -1 : (int)it->first);
            }
            uint other = isA ? nB : nA;
            if (other > 0) { k.cnt.push_back(other); k.dA.push_back(-1); k.dB.push_back(-1); }
        } else {   // :802-886 — std::map ordered by first+second only (TypeContainer.hpp:79-86)
            struct SumLess { bool operator()(const std::pair<int, int>& a, const std::pair<int, int>& b) const { return a.first + a.second < b.first + b.second; } };
            std::map<std::pair<int, int>, uint, SumLess> red;
            auto put = [&](int da, int db, uint c) { std::pair<int, int> key(da, db); if (!red.insert({key, c}).second) red.at(key) += c; };
            for (auto& kv : A) put(__builtin_popcountll((kv.first ^ q.word) & cx.maskA), -1, kv.second);
            for (auto& kv : B) put(-1, __builtin_popcountll((kv.first ^ q.word) & cx.maskB), kv.second);
            for (auto& kv : C) put(__builtin_popcountll((kv.first ^ q.word) & cx.maskA), __builtin_popcountll((kv.first ^ q.word) & cx.maskB), kv.second);
            for (auto it = red.rbegin(); it != red.rend(); ++it) { k.cnt.push_back(it->second); k.dA.push_back(it->first.first); k.dB.push_back(it->first.second); }
        }
        return k;
    }
    double permutation_constant(int t) const override {
        double p = 0; uint tot = 0;
        for (auto& kv : inA[t]) { tot += kv.second; p += lgamma_safe(kv.second + 1); }
        p -= lgamma_safe(tot + 1); tot = 0;
        for (auto& kv : inB[t]) { tot += kv.second; p += lgamma_safe(kv.second + 1); }
        p -= lgamma_safe(tot + 1); tot = 0;
        for (auto& kv : inC[t]) { tot += kv.second; p += lgamma_safe(kv.second + 1); }
        p -= lgamma_safe(tot + 1);
        return p;
    }
    uint unique_count() const override { return A.size() + B.size() + C.size(); }
};

// ---- CANONICAL: one slot array; the conventions of the CUDA path ------------------
//  * iteration = slot order; a new type is appended; counts change in place
//  * a slot whose count reached 0 is dropped at the END of the event: scanning slots from the
//    highest index down, an empty slot is overwritten by the last slot and the array shrinks
//  * temporary add/remove around a piSMC call never touch the array (they are count deltas)
//  * getDifferences for a C query: classes keyed by dA+dB; the representative (dA,dB) is that of an
//    A type if one has this sum, else a B type, else the C type with the smallest dA
struct CanonStore : Store {
    std::vector<TypeRec> slots;
    std::vector<TempOp> temps;
    explicit CanonStore(const Ctx& c) : Store(c) {}
    int find(int gamma, uint64_t w) const {
        for (size_t i = 0; i < slots.size(); i++) if (slots[i].gamma == gamma && slots[i].word == w) return (int)i;
        return -1;
    }
    void bump(int gamma, uint64_t w, uint by) {
        int i = find(gamma, w);
        if (i >= 0) slots[i].count += by; else slots.push_back(TypeRec{gamma, w, by});
        (gamma == COALO_GAMMA_A ? nA : gamma == COALO_GAMMA_B ? nB : nC) += by;
    }
    void reset(const std::vector<TypeRec>& init) override {
        slots.clear(); temps.clear(); nA = nB = nC = 0;
        for (const TypeRec& r : init) bump(r.gamma, r.word, r.count);
    }
    void add(const Hap& h) override { if (tr) tr->note(h, +1); bump(h.gamma, h.word, 1); }
    void remove(const Hap& h) override {
        if (tr) tr->note(h, -1);
        int i = find(h.gamma, h.word); assert(i >= 0 && slots[i].count > 0);
        slots[i].count--;
        (h.gamma == COALO_GAMMA_A ? nA : h.gamma == COALO_GAMMA_B ? nB : nC)--;
    }
    void commit() override {
        for (int i = (int)slots.size() - 1; i >= 0; i--)
            if (slots[i].count == 0) { slots[i] = slots.back(); slots.pop_back(); }
    }
    void push_temp(const TempOp& op) override {
        temps.push_back(op);
        (op.h.gamma == COALO_GAMMA_A ? nA : op.h.gamma == COALO_GAMMA_B ? nB : nC) += op.add ? 1 : -1;
    }
    void pop_temp(const TempOp& op) override {
        assert(!temps.empty()); temps.pop_back();
        (op.h.gamma == COALO_GAMMA_A ? nA : op.h.gamma == COALO_GAMMA_B ? nB : nC) -= op.add ? 1 : -1;
    }
    // effective multiset = slots with temp deltas folded in
    void effective(std::vector<TypeRec>& out) const {
        out = slots;
        for (const TempOp& op : temps) {
            bool hit = false;
            for (TypeRec& r : out) if (r.gamma == op.h.gamma && r.word == op.h.word) { r.count += op.add ? 1 : -1; hit = true; break; }
            if (!hit) { assert(op.add); out.push_back(TypeRec{op.h.gamma, op.h.word, 1}); }
        }
    }
    uint sum_if(int gamma, uint64_t mask, uint64_t w) const {
        uint s = 0;
        for (const TypeRec& r : slots) if (r.gamma == gamma && (r.word & mask) == w) s += r.count;
        return s;
    }
    uint cntA(uint64_t wa) const override { return sum_if(COALO_GAMMA_A, ~0ULL, wa); }
    uint cntB(uint64_t wb) const override { return sum_if(COALO_GAMMA_B, ~0ULL, wb); }
    uint cntC(uint64_t wc) const override { return sum_if(COALO_GAMMA_C, ~0ULL, wc); }
    uint cntCa(uint64_t wa) const override { return sum_if(COALO_GAMMA_C, cx.maskA, wa); }
    uint cntCb(uint64_t wb) const override { return sum_if(COALO_GAMMA_C, cx.maskB, wb); }
    std::vector<std::pair<uint, uint>> add_map(int t) override {
        std::vector<std::pair<uint, uint>> ov;
        // merge duplicates of the input block first (the reference's input maps are maps)
        std::vector<TypeRec> in;
        for (const TypeRec& r : cx.types[t]) {
            bool hit = false;
            for (TypeRec& q : in) if (q.gamma == r.gamma && q.word == r.word) { q.count += r.count; hit = true; break; }
            if (!hit) in.push_back(r);
        }
        for (const TypeRec& r : in) { bump(r.gamma, r.word, r.count); ov.push_back({slots[find(r.gamma, r.word)].count, r.count}); }
        return ov;
    }
    void sampling_order(std::vector<TypeRec>& out) const override {
        out.clear();
        for (const TypeRec& r : slots) if (r.count > 0) out.push_back(r);
    }
    void list(int gamma, std::vector<uint64_t>& out) const override {
        out.clear();
        for (const TypeRec& r : slots) if (r.gamma == gamma && r.count > 0) out.push_back(r.word);
    }
    Classes differences(const Hap& q) const override {
        std::vector<TypeRec> eff; effective(eff);
        Classes k;
        if (q.gamma == COALO_GAMMA_A || q.gamma == COALO_GAMMA_B) {
            const bool isA = q.gamma == COALO_GAMMA_A;
            const uint64_t mask = isA ? cx.maskA : cx.maskB;
            std::map<uint, uint> red;
            uint other = 0;
            for (const TypeRec& r : eff) {
                if (r.count == 0) continue;
                if (r.gamma == q.gamma || r.gamma == COALO_GAMMA_C) red[(uint)__builtin_popcountll((r.word ^ q.word) & mask)] += r.count;
                else other += r.count;
            }
            for (auto it = red.rbegin(); it != red.rend(); ++it) {
                k.cnt.push_back(it->second); k.dA.push_back(isA ? (int)it->first : -1); k.dB.push_back(isA ? -1 : (int)it->first);
            }
            if (other > 0) { k.cnt.push_back(other); k.dA.push_back(-1); k.dB.push_back(-1); }
        } else {
            struct Cls { uint cnt = 0; int prio = 1 << 30; };
            std::map<int, Cls> red;   // keyed by dA+dB
            for (const TypeRec& r : eff) {
                if (r.count == 0) continue;
                int da = -1, db = -1, prio;
                if (r.gamma != COALO_GAMMA_B) da = __builtin_popcountll((r.word ^ q.word) & cx.maskA);
                if (r.gamma != COALO_GAMMA_A) db = __builtin_popcountll((r.word ^ q.word) & cx.maskB);
                prio = r.gamma == COALO_GAMMA_A ? 0 : (r.gamma == COALO_GAMMA_B ? 1 : 2 + da);
                Cls& c = red[da + db]; c.cnt += r.count; c.prio = std::min(c.prio, prio);
            }
            for (auto it = red.rbegin(); it != red.rend(); ++it) {
                const int sum = it->first, prio = it->second.prio;
                int da, db;
                if (prio == 0) { da = sum + 1; db = -1; } else if (prio == 1) { da = -1; db = sum + 1; } else { da = prio - 2; db = sum - da; }
                k.cnt.push_back(it->second.cnt); k.dA.push_back(da); k.dB.push_back(db);
            }
        }
        return k;
    }
    double permutation_constant(int t) const override {
        // same terms as the strict store; input order, duplicates merged
        std::vector<TypeRec> in;
        for (const TypeRec& r : cx.types[t]) {
            bool hit = false;
            for (TypeRec& q : in) if (q.gamma == r.gamma && q.word == r.word) { q.count += r.count; hit = true; break; }
            if (!hit) in.push_back(r);
        }
        double p = 0;
        for (int g = 0; g < 3; g++) {
            uint tot = 0;
            for (const TypeRec& r : in) if (r.gamma == g) { tot += r.count; p += lgamma_safe(r.count + 1); }
            p -= lgamma_safe(tot + 1);
        }
        return p;
    }
    uint unique_count() const override { uint u = 0; for (const TypeRec& r : slots) if (r.count) u++; return u; }
};

// ------------------------------------------------------------------------------------------------
// one history — ProposalEngine::calculateLikelihood (ProposalEngine.cpp:643-861) and
// sample_backwardMove_forwardTransition (:134-566)
// ------------------------------------------------------------------------------------------------

struct Engine {
    Ctx& cx; Store& H; Rng& rng; Tracer* tr;
    int64_t n_pismc = 0, n_events = 0;
    std::vector<coalo_event>* evlog = nullptr;
    // Candidate-vector dump (tests of the device event code, coalo_set_score_dump): per event one line
    //   V <epoch> <chosen gamma> <chosen word> <ntypes> (<gamma> <word> <count>)* <ncand> (<label> <value>)*
    // with the configuration BEFORE the removal of the chosen lineage, in the store's sampling order, and the unnormalised
    // proposal probabilities of ProposalEngine.cpp:171-238 / 298-338 / 375-414 labelled by move: m<site>, cc, ca, cb, rec
    // (C chosen); m<site>, p<partner word>, same (A / B chosen).
    FILE* vdump = nullptr;
    std::vector<std::string> vlabels;
    std::vector<double>* vcapture = nullptr;      // coalo_score_event: keep the vector, stop before the draw
    Engine(Ctx& c, Store& s, Rng& r, Tracer* t) : cx(c), H(s), rng(r), tr(t) {}

    double pismc(const Hap& q, uint tidx) {
        n_pismc++;
        Classes k = H.differences(q);
        double pi = pismc_classes(cx, q.gamma, H.n(), tidx, k);
        if (tr && tr->on) {
            fprintf(tr->f, "P %c %u %u %d %d %zu", "ABC"[q.gamma], H.n(), tidx, cx.LA, cx.LB, k.cnt.size());
            for (size_t i = 0; i < k.cnt.size(); i++) fprintf(tr->f, " %d %d %d", k.cnt[i], k.dA[i], k.dB[i]);
            fprintf(tr->f, " %.17g\n", pi);
        }
        return pi;
    }
    // ProposalEngine::piSMCJointly :120-132
    double pismc_jointly(const Hap& a, const Hap& b, uint tidx) {
        double t = pismc(a, tidx);
        TempOp addA{true, a}; H.push_temp(addA);
        double first = 0.5 * t * pismc(b, tidx);
        H.pop_temp(addA);
        t = pismc(b, tidx);
        TempOp addB{true, b}; H.push_temp(addB);
        double second = 0.5 * t * pismc(a, tidx);
        H.pop_temp(addB);
        return first + second;
    }
    static double vsum(const std::vector<double>& v) { double s = 0; for (double x : v) s += x; return s; }
    // ProposalEngine::sampleFromVector :568-583
    uint sample_from(const std::vector<double>& v) {
        double r = rng.uniform() * vsum(v);
        uint k = 0; double cum = v[0];
        while (cum < r && k + 1 < v.size()) { k++; cum += v[k]; }
        return k;
    }
    // TypeContainer::sampleType :641-745
    Hap sample_type(double thA, double thB, double rho, double& prob, double& norm_out) {
        std::vector<TypeRec> ord; H.sampling_order(ord);
        std::vector<double> pr(ord.size());
        double norm = 0;
        for (size_t i = 0; i < ord.size(); i++) {
            const TypeRec& r = ord[i]; double p;
            if (r.gamma == COALO_GAMMA_C) {
                uint cA = H.cntA(H.partA(r.word)), cB = H.cntB(H.partB(r.word));
                p = (r.count - 1 + (thA + thB) + rho + cA + cB) * r.count;
            } else if (r.gamma == COALO_GAMMA_A) {
                uint cCa = H.cntCa(r.word);
                p = (r.count - 1 + H.nB + cCa + thA) * r.count;
            } else {
                uint cCb = H.cntCb(r.word);
                p = (r.count - 1 + H.nA + cCb + thB) * r.count;
            }
            pr[i] = p; norm += p;
        }
        double thr = rng.uniform() * norm, cum = 0;
        size_t pick = ord.size() - 1;
        for (size_t i = 0; i < ord.size(); i++) { cum += pr[i]; if (cum > thr) { pick = i; break; } }
        prob = pr[pick] / norm; norm_out = norm;
        Hap h{ord[pick].gamma, ord[pick].word};
        if (tr && tr->on) { tr->flush(); fprintf(tr->f, "S %ld %d %s %.17g %.17g\n", tr->hist, tr->epoch, tr->key(h).c_str(), prob, norm); }
        return h;
    }

    // candidate-vector dump / capture; returns true when the caller must stop before the draw (coalo_score_event)
    bool score_hook(const Hap& x, uint c, const std::vector<TypeRec>& cfg, const std::vector<double>& pv) {
        if (vdump) {
            fprintf(vdump, "V %u %d %llx %zu", c, x.gamma, (unsigned long long)x.word, cfg.size());
            for (const TypeRec& r : cfg) fprintf(vdump, " %d %llx %u", r.gamma, (unsigned long long)r.word, r.count);
            fprintf(vdump, " %zu", pv.size());
            for (size_t k = 0; k < pv.size(); k++) fprintf(vdump, " %s %.17g", vlabels[k].c_str(), pv[k]);
            fprintf(vdump, "\n");
        }
        if (vcapture) { *vcapture = pv; return true; }
        return false;
    }

    void log_event(const Hap& x, int kind, uint64_t aux, int site, uint c, uint n_before) {
        n_events++;
        if (!evlog) return;
        coalo_event e; e.chosen_word = x.word; e.aux_word = aux; e.chosen_gamma = (uint8_t)x.gamma;
        e.kind = (uint8_t)kind; e.epoch = (uint8_t)c; e.site = (uint8_t)site; e.n_before = n_before;
        evlog->push_back(e);
    }

    // :134-566. Returns log of the normalised proposal probability of the sampled move.
    double event(double dN, double dtheta, double drho, std::vector<double>& L, const Hap& x, double wait,
                 uint c, double V, bool& postCollection, bool afterLast) {
        const uint LA = cx.LA, LB = cx.LB, Ltot = LA + LB;
        (void)dN;
        // counts before removal :149-158
        uint Cij = 0, Ci = 0, Cj = 0, Ai = 0, Bj = 0;
        if (x.gamma == COALO_GAMMA_C) { Cij = H.cntC(x.word); Ci = H.cntCa(H.partA(x.word)); Cj = H.cntCb(H.partB(x.word)); Ai = H.cntA(H.partA(x.word)); Bj = H.cntB(H.partB(x.word)); }
        else if (x.gamma == COALO_GAMMA_A) { Ci = H.cntCa(x.word); Ai = H.cntA(x.word); }
        else { Cj = H.cntCb(x.word); Bj = H.cntB(x.word); }
        const uint n = H.n(), C = H.nC, A = H.nA, B = H.nB;
        double ec = 0; int kind = -1;   // 0 coalescence, 1 mutation, 2 recombination
        std::vector<double> pv; uint idx = 0;
        std::vector<TypeRec> vcfg;
        if (vdump || vcapture) { H.sampling_order(vcfg); vlabels.clear(); }
        auto label = [&](const std::string& l) { if (vdump || vcapture) vlabels.push_back(l); };
        char lb[40];

        H.remove(x);                                   // :168
        const double pi0 = pismc(x, c);                // :169

        if (x.gamma == COALO_GAMMA_C) {                // :171-295
            pv.assign(Ltot + 4, 0.0);
            for (uint s = 0; s < Ltot; s++) {          // :179-187
                Hap m{COALO_GAMMA_C, x.word ^ (1ULL << s)};
                double pm = pismc(m, c);
                pv[s] = dtheta * Cij * pm / pi0;
                snprintf(lb, sizeof lb, "m%u", s); label(lb);
            }
            label("cc"); label("ca"); label("cb"); label("rec");
            pv[Ltot] = (Cij > 1) ? Cij * (Cij - 1) / pi0 : 0;      // :190-195
            const Hap xa{COALO_GAMMA_A, H.partA(x.word)}, xb{COALO_GAMMA_B, H.partB(x.word)};
            if (Ai > 0) {                              // :198-212
                TempOp o1{true, x}, o2{false, xa};
                H.push_temp(o1); H.push_temp(o2);
                pv[Ltot + 1] = Ai * Cij / pismc(xa, c);
                H.pop_temp(o2); H.pop_temp(o1);
            }
            if (Bj > 0) {                              // :215-229
                TempOp o1{true, x}, o2{false, xb};
                H.push_temp(o1); H.push_temp(o2);
                pv[Ltot + 2] = Bj * Cij / pismc(xb, c);
                H.pop_temp(o2); H.pop_temp(o1);
            }
            double pj = pismc_jointly(xa, xb, c);      // :232-238
            pv[Ltot + 3] = drho * Cij * pj / pi0;
            if (score_hook(x, c, vcfg, pv)) return 0.0;
            idx = sample_from(pv);                     // :241
            if (idx < Ltot) {                          // :242-249
                Hap m{COALO_GAMMA_C, x.word ^ (1ULL << idx)};
                H.add(m);
                double Ckj = H.cntC(m.word);
                ec = Ckj; kind = 1;
                log_event(x, COALO_EV_MUT, m.word, idx, c, n);
            } else if (idx == Ltot) {                  // :250-253
                ec = C * (Cij - 1); kind = 0;
                log_event(x, COALO_EV_COAL_SAME, 0, 0xff, c, n);
            } else if (idx == Ltot + 1) {              // :254-264
                H.add(x); H.remove(xa);
                ec = A * Cij; kind = 0;
                log_event(x, COALO_EV_COAL_CA, xa.word, 0xff, c, n);
            } else if (idx == Ltot + 2) {              // :265-276
                H.add(x); H.remove(xb);
                ec = B * Cij; kind = 0;
                log_event(x, COALO_EV_COAL_CB, xb.word, 0xff, c, n);
            } else {                                   // :277-291
                H.add(xa); H.add(xb);
                double Ai2 = H.cntA(xa.word), Bj2 = H.cntB(xb.word);
                ec = C * Ai2 * Bj2 / ((A + 1.0) * (B + 1.0)); kind = 2;
                log_event(x, COALO_EV_REC, 0, 0xff, c, n);
            }
        } else {                                       // :298-372 (A) and :375-449 (B), mirror images
            const bool isA = x.gamma == COALO_GAMMA_A;
            const uint Lx = isA ? LA : LB, off = isA ? 0 : LA;
            const uint own = isA ? Ai : Bj, ownC = isA ? Ci : Cj;
            std::vector<uint64_t> partners; H.list(isA ? COALO_GAMMA_B : COALO_GAMMA_A, partners);
            pv.assign(Lx + partners.size() + 1, 0.0);
            for (uint s = 0; s < Lx; s++) {            // :304-312 / :381-389
                Hap m{x.gamma, x.word ^ (1ULL << (off + s))};
                double pm = pismc(m, c);
                pv[s] = dtheta * own * pm / pi0;
                snprintf(lb, sizeof lb, "m%u", s); label(lb);
            }
            for (size_t k = 0; k < partners.size(); k++) { snprintf(lb, sizeof lb, "p%llx", (unsigned long long)partners[k]); label(lb); }
            label("same");
            for (size_t k = 0; k < partners.size(); k++) {   // :315-330 / :392-406
                Hap pt{isA ? COALO_GAMMA_B : COALO_GAMMA_A, partners[k]};
                Hap joined{COALO_GAMMA_C, x.word | pt.word};
                uint pc = isA ? H.cntB(pt.word) : H.cntA(pt.word);
                TempOp o{false, pt};
                H.push_temp(o);
                double pu = pismc(joined, c);
                // the reference passes (chosen, partner) in both branches (:326, :403)
                double pj = pismc_jointly(x, pt, c);
                pv[Lx + k] = isA ? (own * pc * pu / pj) : (pc * own * pu / pj);
                H.pop_temp(o);
            }
            pv[Lx + partners.size()] = (own > 1 || ownC > 0) ? own * (own - 1 + ownC) / pi0 : 0;   // :333-338 / :409-414
            if (score_hook(x, c, vcfg, pv)) return 0.0;
            idx = sample_from(pv);                     // :341 / :417
            if (idx < Lx) {                            // :342-350 / :418-426
                Hap m{x.gamma, x.word ^ (1ULL << (off + idx))};
                H.add(m);
                double cnt = isA ? H.cntA(m.word) : H.cntB(m.word);
                ec = cnt; kind = 1;
                log_event(x, COALO_EV_MUT, m.word, idx, c, n);
            } else if (idx < Lx + partners.size()) {   // :351-363 / :427-440
                Hap pt{isA ? COALO_GAMMA_B : COALO_GAMMA_A, partners[idx - Lx]};
                Hap joined{COALO_GAMMA_C, x.word | pt.word};
                H.remove(pt); H.add(joined);
                double Cnew = H.cntC(joined.word);
                ec = isA ? (A * B * Cnew / (C + 1.0)) : (A * B * Cnew / (C + 1));
                kind = 0;
                log_event(x, COALO_EV_COAL_AB, pt.word, 0xff, c, n);
            } else {                                   // :364-368 / :441-445
                ec = isA ? A * (Ai + Ci - 1) : B * (Bj + Cj - 1); kind = 0;
                log_event(x, COALO_EV_COAL_SAME, 0, 0xff, c, n);
            }
        }
        H.commit();

        const double qmove = std::log(pv[idx]) - std::log(vsum(pv));   // :455

        // :457-546 — target density of the move on the whole (mu, rho, lambda) grid
        const bool post = postCollection;
        if (post) postCollection = false;
        size_t g = 0;
        for (size_t i = 0; i < cx.mu.size(); i++) for (size_t j = 0; j < cx.rho.size(); j++) for (size_t z = 0; z < cx.lambda.size(); z++, g++) {
            double N = cx.lambda[z] * V;
            double theta = 4 * cx.mu[i] * cx.lambda[z] * V;
            double rho = 4 * cx.rho[j] * cx.lambda[z] * V;
            double D = n * (n - 1) + (A + C) * LA * theta + (B + C) * LB * theta + rho * C;
            double rate = D / (2 * N * cx.tau);
            double head = kind == 0 ? std::log(ec / D) : (kind == 1 ? std::log(theta * ec / D) : std::log(rho * ec / D));
            double tail = post ? (-rate * wait) : (std::log(rate) - rate * wait);
            L[g] += head + tail;
            if (afterLast && H.n() == 1) L[g] += -std::log(ec / D);   // :549-562 (unreachable: loop stops at 5)
        }
        return qmove;
    }

    // ProposalEngine::probabilityCombinatorics :66-86
    double combinatorics(uint a0, uint b0, uint c0, const std::vector<std::pair<uint, uint>>& ov) {
        double f = -std::log(binom(H.nA, H.nA - a0));
        f -= std::log(binom(H.nB, H.nB - b0));
        f -= std::log(binom(H.nC, H.nC - c0));
        for (auto& pr : ov) f += std::log(binom(pr.first, pr.second));
        return f;
    }

    // :643-861
    void history(std::vector<double>& L, double& tmrca, std::vector<double>& left) {
        const int T = cx.T; const uint LA = cx.LA, LB = cx.LB;
        L.assign(cx.G, 0.0); left.assign(T, 0.0);
        double V = cx.viral[0];
        double dN = cx.dlambda * V, dtheta = 4 * cx.dmu * dN, drho = 4 * cx.drho * dN;
        double dthA = dtheta * LA, dthB = dtheta * LB;
        const double tau = cx.tau;
        H.reset(cx.types[0]);
        double q = 0, pr, nrm;
        Hap x = sample_type(dthA, dthB, drho, pr, nrm);
        uint nt = H.n(), Ct = H.nC, At = H.nA, Bt = H.nB;
        double D = nt * (nt - 1) + (At + Ct) * dthA + (Bt + Ct) * dthB + drho * Ct;
        double rate = D / (2 * dN * tau);
        double wait = -std::log(rng.uniform()) / rate;     // sampleEventTime :52-56
        double tsum = wait;
        bool post = true;
        auto redraw = [&]() {
            x = sample_type(dthA, dthB, drho, pr, nrm);
            nt = H.n(); Ct = H.nC; At = H.nA; Bt = H.nB;
            D = nt * (nt - 1) + (At + Ct) * dthA + (Bt + Ct) * dthB + drho * Ct;
            rate = D / (2 * dN * tau);
            wait = -std::log(rng.uniform()) / rate;
            tsum += wait;
        };
        for (int c = 1; c < T; c++) {
            while (tsum < cx.times[c]) {                   // :695-718
                q += std::log(rate) - rate * wait;
                q += std::log(pr);
                q += event(dN, dtheta, drho, L, x, wait, c - 1, V, post, false);
                redraw();
            }
            double excess = cx.times[c] - (tsum - wait);   // :721-726
            tsum = cx.times[c];
            q += -rate * excess;
            const uint n = H.n(), C = H.nC, A = H.nA, B = H.nB;   // :729-747
            size_t g = 0;
            for (size_t i = 0; i < cx.mu.size(); i++) for (size_t j = 0; j < cx.rho.size(); j++) for (size_t z = 0; z < cx.lambda.size(); z++, g++) {
                double N = cx.lambda[z] * V;
                double theta = 4 * cx.mu[i] * cx.lambda[z] * V;
                double rho = 4 * cx.rho[j] * cx.lambda[z] * V;
                double Dg = n * (n - 1) + (A + C) * theta * LA + (B + C) * theta * LB + rho * C;
                double fr = Dg / (2 * N * tau);
                L[g] += std::log(fr) - fr * excess;
            }
            left[c - 1] = n;                               // :750-756
            const uint a0 = H.nA, b0 = H.nB, c0 = H.nC;
            if (tr && tr->on) { tr->flush(); tr->epoch++; fprintf(tr->f, "M %ld %d\n", tr->hist, tr->epoch); }
            Tracer* keep = H.tr; H.tr = nullptr;           // the injection itself is not an event delta
            std::vector<std::pair<uint, uint>> ov = H.add_map(c);
            H.tr = keep;
            post = true;
            double comb = combinatorics(a0, b0, c0, ov);
            for (double& v : L) v += comb;
            V = cx.viral[c];                               // :759-765
            dN = cx.dlambda * V; dtheta = 4 * cx.dmu * dN; drho = 4 * cx.drho * dN;
            dthA = dtheta * LA; dthB = dtheta * LB;
            redraw();                                      // :768-782
        }
        nt = H.n(); Ct = H.nC; At = H.nA; Bt = H.nB;       // :786-790
        while (nt > 5) {                                   // :792-815
            q += std::log(rate) - rate * wait;
            q += std::log(pr);
            q += event(dN, dtheta, drho, L, x, wait, T - 1, V, post, true);
            redraw();
        }
        left[T - 1] = nt;                                  // :817-818
        tsum -= wait;
        double perm = 0;                                   // :822-828
        for (int i = 0; i < T; i++) perm += H.permutation_constant(i);
        const double c1 = perm - q;                        // :835
        for (double& v : L) v += c1;
        const double c2 = -(double)(LA + LB) * std::log(2) * Ct - (double)(LA) * std::log(2) * At - (double)(LB) * std::log(2) * Bt;   // :838
        for (double& v : L) v += c2;
        tmrca = tsum;
        if (tr && tr->on) {
            tr->flush();
            fprintf(tr->f, "W %ld %.17g %zu", tr->hist, tmrca, left.size());
            for (double v : left) fprintf(tr->f, " %.17g", v);
            fprintf(tr->f, " %d", cx.G);
            for (double v : L) fprintf(tr->f, " %.17g", v);
            fprintf(tr->f, "\n");
        }
    }
};

}  // namespace

// ------------------------------------------------------------------------------------------------
// C API
// ------------------------------------------------------------------------------------------------

extern "C" {

int coalo_tables(const coalo_problem* p, int Q, int n_all, int time_idx, double* w, double* y, double* z, double* EA, double* EB) {
    Ctx c; if (!init_ctx(c, p, Q)) return -1;
    if (time_idx < 0 || time_idx >= c.T || n_all < 1) return -2;
    if (w) std::copy(c.w.begin(), c.w.end(), w);
    const Ctx::Trans& tr = transition(c, n_all, time_idx);
    if (y) std::copy(tr.y.begin(), tr.y.end(), y);
    if (z) std::copy(tr.z.begin(), tr.z.end(), z);
    if (EA) { const std::vector<double>& e = emission(c, n_all, time_idx, c.LA); std::copy(e.begin(), e.end(), EA); }
    if (EB) { const std::vector<double>& e = emission(c, n_all, time_idx, c.LB); std::copy(e.begin(), e.end(), EB); }
    return 0;
}

double coalo_pismc_classes(const coalo_problem* p, int Q, int gamma, int n_all, int time_idx, int n_classes,
                           const int32_t* counts, const int32_t* dA, const int32_t* dB) {
    // The context (lazily built tables) is reused between calls for the same problem AND quadrature. The key is a
    // fingerprint of what the tables depend on, not the pointer: a freed problem's address can be handed out again.
    static thread_local std::unique_ptr<Ctx> cache; static thread_local uint64_t cache_key = 0;
    uint64_t key = 1469598103934665603ull;
    auto mix = [&key](const void* data, size_t bytes) { const unsigned char* b = (const unsigned char*)data; for (size_t i = 0; i < bytes; i++) { key ^= b[i]; key *= 1099511628211ull; } };
    mix(&Q, sizeof Q); mix(&p->T, sizeof p->T); mix(&p->LA, sizeof p->LA); mix(&p->LB, sizeof p->LB);
    mix(p->times, sizeof(double) * p->T); mix(p->viral, sizeof(double) * p->T);
    mix(&p->tau, sizeof(double)); mix(&p->driving_mu, sizeof(double)); mix(&p->driving_lambda, sizeof(double)); mix(&p->driving_rho, sizeof(double));
    mix(p->type_offsets, sizeof(int32_t) * (p->T + 1));
    mix(p->type_counts, sizeof(uint32_t) * p->type_offsets[p->T]);
    if (!cache || cache_key != key) { cache.reset(new Ctx); if (!init_ctx(*cache, p, Q)) { cache.reset(); return std::nan(""); } cache_key = key; }
    Classes k;
    k.cnt.assign(counts, counts + n_classes); k.dA.assign(dA, dA + n_classes); k.dB.assign(dB, dB + n_classes);
    return pismc_classes(*cache, gamma, n_all, time_idx, k);
}

static std::string g_score_dump_path;
static int64_t g_score_dump_histories = 0;
void coalo_set_score_dump(const char* path, int64_t n_histories) { g_score_dump_path = path ? path : ""; g_score_dump_histories = n_histories; }

// The candidate vector of ONE event in CANONICAL mode: the configuration is rebuilt from the given type list (slot order =
// list order, the chosen lineage still in it), the chosen lineage is removed and every backward move is scored exactly as
// Engine::event does before its draw. out_vals[k] / labels as in the V records; returns the number of candidates or < 0.
int coalo_score_event(const coalo_problem* p, int Q, int epoch, int n_types, const uint8_t* gamma, const uint64_t* words,
                      const uint32_t* counts, int chosen_gamma, uint64_t chosen_word, int max_cand, double* out_vals,
                      char* out_labels /* max_cand * 24 bytes */) {
    Ctx cx; if (!init_ctx(cx, p, Q)) return -1;
    if (epoch < 0 || epoch >= cx.T) return -2;
    CanonStore store(cx);
    PhiloxRng rng(0);
    Engine eng(cx, store, rng, nullptr);
    std::vector<TypeRec> init;
    for (int k = 0; k < n_types; k++) init.push_back(TypeRec{(int)gamma[k], words[k], counts[k]});
    store.reset(init);
    std::vector<double> pv, L(cx.G, 0.0);
    eng.vcapture = &pv;
    const double V = cx.viral[epoch];
    const double dN = cx.dlambda * V, dtheta = 4 * cx.dmu * dN, drho = 4 * cx.drho * dN;
    bool post = true;
    Hap x{chosen_gamma, chosen_word};
    eng.event(dN, dtheta, drho, L, x, 1.0, (uint)epoch, V, post, false);
    if ((int)pv.size() > max_cand) return -3;
    for (size_t k = 0; k < pv.size(); k++) {
        out_vals[k] = pv[k];
        if (out_labels) { snprintf(out_labels + 24 * k, 24, "%s", eng.vlabels[k].c_str()); }
    }
    return (int)pv.size();
}

int coalo_run(const coalo_problem* p, int Q, int mode, uint64_t seed, uint64_t first, uint64_t n,
              double* weights, double* tmrca, double* lineages, int64_t* n_events, int64_t* n_pismc,
              const char* trace_path, int64_t trace_max,
              coalo_event* events, int64_t n_event_hist, int64_t max_events, int64_t* event_counts) {
    Ctx cx; if (!init_ctx(cx, p, Q)) return -1;
    if (mode == COALO_MODE_STRICT && first != 0) return -3;
    std::unique_ptr<Store> store;
    std::unique_ptr<Rng> rng;
    if (mode == COALO_MODE_STRICT) { store.reset(new StrictStore(cx)); rng.reset(new MtRng((uint)seed)); }
    else if (mode == COALO_MODE_CANONICAL) {
        // canonical input order (shared with the CUDA path, coal_api.cu:prepare_host): the types of a sampling time sorted by
        // (class, packed haplotype), so that the histories do not depend on the order in which a caller lists them (the
        // reference's own order is the iteration order of its unordered_maps; STRICT mode keeps the given order for that)
        for (auto& block : cx.types)
            std::stable_sort(block.begin(), block.end(), [](const TypeRec& a, const TypeRec& b) { return a.gamma != b.gamma ? a.gamma < b.gamma : a.word < b.word; });
        store.reset(new CanonStore(cx)); rng.reset(new PhiloxRng(seed));
    }
    else return -4;
    Tracer tr; tr.cx = &cx;
    if (trace_path) { tr.f = fopen(trace_path, "w"); if (!tr.f) return -5; store->tr = &tr; }
    Engine eng(cx, *store, *rng, trace_path ? &tr : nullptr);
    FILE* vfile = g_score_dump_path.empty() ? nullptr : fopen(g_score_dump_path.c_str(), "w");
    std::vector<double> L, left; double t;
    std::vector<coalo_event> evlog;
    for (uint64_t h = 0; h < n; h++) {
        eng.vdump = (vfile && (int64_t)h < g_score_dump_histories) ? vfile : nullptr;
        rng->begin_history(first + h);
        if (trace_path) { tr.hist = (long)h; tr.on = (int64_t)h < trace_max; tr.epoch = 0; tr.delta.clear(); if (tr.on) fprintf(tr.f, "B %ld\n", tr.hist); }
        const bool want_ev = events && (int64_t)h < n_event_hist;
        evlog.clear(); eng.evlog = want_ev ? &evlog : nullptr;
        const int64_t e0 = eng.n_events, p0 = eng.n_pismc;
        eng.history(L, t, left);
        if (weights) std::copy(L.begin(), L.end(), weights + h * cx.G);
        if (tmrca) tmrca[h] = t;
        if (lineages) std::copy(left.begin(), left.end(), lineages + h * cx.T);
        if (n_events) n_events[h] = eng.n_events - e0;
        if (n_pismc) n_pismc[h] = eng.n_pismc - p0;
        if (want_ev) {
            int64_t m = std::min<int64_t>((int64_t)evlog.size(), max_events);
            std::copy(evlog.begin(), evlog.begin() + m, events + h * max_events);
            if (event_counts) event_counts[h] = (int64_t)evlog.size();
        }
    }
    if (tr.f) fclose(tr.f);
    if (vfile) fclose(vfile);
    return 0;
}

int coalo_reduce(int64_t n, int32_t G, int32_t T, const double* weights, const double* tmrca, const double* lineages,
                 double* out_like, double* out_likesq, double* out_tmrca, double* out_lineages) {
    if (n < 1) return -1;
    std::vector<double> v(n);
    auto lse_desc = [&](std::vector<double>& a) {   // LikelihoodSamples.cpp:56-70 (multiset, largest first)
        std::sort(a.begin(), a.end());
        double s = a[n - 1];
        for (int64_t i = n - 2; i >= 0; i--) s = log_addition(s, a[i]);
        return s;
    };
    for (int g = 0; g < G; g++) {
        for (int64_t h = 0; h < n; h++) v[h] = weights[h * G + g];
        const double like = lse_desc(v);
        for (int64_t h = 0; h < n; h++) v[h] = weights[h * G + g] * 2;             // SamplerOutput.cpp:79
        const double likesq = lse_desc(v);
        for (int64_t h = 0; h < n; h++) v[h] = weights[h * G + g] + std::log(tmrca[h]);   // :85
        out_tmrca[g] = lse_desc(v) - like;                                           // :120-121
        for (int l = 0; l < T; l++) {
            for (int64_t h = 0; h < n; h++) v[h] = weights[h * G + g] + std::log(lineages[h * T + l]);   // :91
            out_lineages[(size_t)l * G + g] = lse_desc(v) - like;                    // :126-129
        }
        // :131-132 subtract num_samples, which the result object holds as 0 (ImportanceSampler.cpp:141)
        out_like[g] = like - 0; out_likesq[g] = likesq - 0;
    }
    return 0;
}

double coalo_uniform(uint64_t seed, uint64_t hist, uint64_t k) { return Philox::uniform(seed, hist, k); }

}  // extern "C"
