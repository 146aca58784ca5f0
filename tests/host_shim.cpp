// CPU-side harness around the product's host code and the arithmetic cores its kernels use
// (coalescenator_b200/csrc/coal_tables.cpp, coal_math.h), so they can be checked against the oracle
// where no GPU exists: the table builder (reference-layout rows and the bilinear device rows) and the
// bilinear evaluation of pi_SMC that every lane of the history kernel runs.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../coalescenator_b200/csrc/coal_math.h"
#include "../coalescenator_b200/csrc/coal_tables.h"

using namespace coalgpu;

extern "C" {

struct hs_tables_t { HostTables tab; };

void* hs_build_q(int Q, int LA, int LB, int n_max, int T, const double* thetas, const double* rrates) {
    hs_tables_t* h = new hs_tables_t;
    std::vector<double> th(thetas, thetas + T), rr(rrates, rrates + T);
    if (!build_tables(Q, LA, LB, n_max, th, rr, h->tab)) { delete h; return nullptr; }
    return h;
}
void* hs_build_mode(int Q, int mode, int LA, int LB, int n_max, int T, const double* thetas, const double* rrates) {
    hs_tables_t* h = new hs_tables_t;
    std::vector<double> th(thetas, thetas + T), rr(rrates, rrates + T);
    if (!build_tables(Q, LA, LB, n_max, th, rr, h->tab, 0, mode)) { delete h; return nullptr; }
    return h;
}
void* hs_build(int LA, int LB, int n_max, int T, const double* thetas, const double* rrates) {
    hs_tables_t* h = new hs_tables_t;
    std::vector<double> th(thetas, thetas + T), rr(rrates, rrates + T);
    if (!build_tables(16, LA, LB, n_max, th, rr, h->tab)) { delete h; return nullptr; }
    return h;
}
void hs_free(void* p) { delete (hs_tables_t*)p; }

int hs_tables(void* hp, int n_all, int time_idx, double* w, double* y, double* z, double* EA, double* EB) {
    const HostTables& t = ((hs_tables_t*)hp)->tab;
    const TableLayout& lay = t.lay;
    const int set = t.set_of_epoch[time_idx];
    const double* row = t.row(set, n_all);
    std::memcpy(w, t.w.data(), sizeof(double) * kQ);
    std::memcpy(y, t.y.data() + ((size_t)set * t.n_max + (n_all - 1)) * kQ, sizeof(double) * kQ);
    for (int i = 0; i < kQ; i++) for (int j = 0; j < kQ; j++)
        z[i * kQ + j] = (i > j) ? row[lay.off_zlow + j] : (i < j ? t.w[j] * row[lay.off_g + i] : row[lay.off_zdiag + i]);
    for (int i = 0; i < kQ; i++) for (int d = 0; d <= lay.LA; d++) EA[i * (lay.LA + 1) + d] = row[lay.off_EA + d * kEStride + i];
    for (int i = 0; i < kQ; i++) for (int d = 0; d <= lay.LB; d++) EB[i * (lay.LB + 1) + d] = row[lay.off_EB + d * kEStride + i];
    return t.n_sets;
}

// pi_SMC from class vectors with the forward pass split into p chunks (p in {1,2,4,8,16})
double hs_pismc(void* hp, int gamma, int n_all, int time_idx, int nH, const int32_t* cnt, const int32_t* dA, const int32_t* dB, int p) {
    const HostTables& t = ((hs_tables_t*)hp)->tab;
    const TableLayout& lay = t.lay;
    (void)lay;
    const double p2A = 1.0 / (double)(1ull << lay.LA), p2B = 1.0 / (double)(1ull << lay.LB);
    double wsum = 0; for (int i = 0; i < kQ; i++) wsum += t.w[i];
    if (gamma != 2) {
        double num = 0; uint32_t ncomp = 0;
        const double* dr = t.dev_row(t.set_of_epoch[time_idx], n_all);   // the one-locus weight vectors as the kernels read them
        for (int k = 0; k < nH; k++) {
            const int d = gamma == 0 ? dA[k] : dB[k];
            if (d >= 0) { num += (double)cnt[k] * dr[(gamma == 0 ? lay.dev_off_wEA : lay.dev_off_wEB) + d]; ncomp += cnt[k]; }
        }
        return marginal_finish(num, ncomp, nH > 0 ? (uint32_t)n_all : 0u, gamma == 0 ? p2A : p2B, wsum);
    }
    if (nH == 0) return p2A * p2B;
    // the interval axis is summed out of the bilinear form: nothing left to split over lanes; p selects the class-list form
    // (p != 2) or the dense form of the lockstep kernel (p == 2)
    std::vector<uint32_t> cls(nH);
    ClassTotals tot{0, 0, 0};
    for (int k = 0; k < nH; k++) {
        cls[k] = pack_class((uint32_t)cnt[k], dA[k], dB[k]);
        tot.n += cnt[k]; if (dA[k] >= 0) tot.ncompA += cnt[k]; if (dB[k] >= 0) tot.ncompB += cnt[k];
    }
    const double* drow = t.dev_row(t.set_of_epoch[time_idx], n_all);
    auto ldmy = [drow](int k) { MYPair e; e.m = drow[2 * k]; e.y = drow[2 * k + 1]; return e; };
    if (p == 2) {
        std::vector<uint32_t> vecA(lay.LA + 1, 0), vecB(lay.LB + 1, 0);
        double t2 = 0;
        for (int k = 0; k < nH; k++) {
            if (dA[k] >= 0) vecA[dA[k]] += (uint32_t)cnt[k] | (dB[k] < 0 ? (uint32_t)cnt[k] << 16 : 0u);
            if (dB[k] >= 0) vecB[dB[k]] += (uint32_t)cnt[k] | (dA[k] < 0 ? (uint32_t)cnt[k] << 16 : 0u);
            if (dA[k] >= 0 && dB[k] >= 0) t2 += (double)cnt[k] * ldmy(dA[k] * lay.dev_sb + dB[k]).y;
        }
        return pismc_bilinear_dense(ldmy, lay.dev_sb, lay.LA, lay.LB, vecA.data(), vecB.data(), 1, t2, tot, tot.ncompA ? 1.0 / (double)tot.ncompA : 0.0,
                                    tot.ncompB ? 1.0 / (double)tot.ncompB : 0.0, 1.0 / (double)tot.n, p2A, p2B);
    }
    return pismc_bilinear(ldmy, lay.dev_sb, lay.LA, lay.LB, cls.data(), 1, nH, tot, tot.ncompA ? 1.0 / (double)tot.ncompA : 0.0,
                          tot.ncompB ? 1.0 / (double)tot.ncompB : 0.0, 1.0 / (double)tot.n, p2A, p2B);
}

}  // extern "C"
