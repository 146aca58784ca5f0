// Arithmetic cores shared by the kernels (device) and the host unit-test shim (tests/ compile this
// header with g++ and compare against the oracle on the CPU box, where no GPU exists).
//
// pi_SMC(h | H) for a two-locus query, ConditionalSamplingHMM.cpp:372-424 of the reference, restated so
// that one lane group can split the Q = 16 absorption-time intervals:
//   A_i  = sum_k c_k eA(i,k)   B_i = sum_k c_k eB(i,k)   X_i = sum_k c_k eA(i,k) eB(i,k)
//   eA(i,k) = EA[i][dA_k] if class k is ancestral at locus A, else the count-weighted mean PA_i of the
//             complete classes (ConditionalSamplingHMM.cpp:234-263), or 2^-LA if there is none
//   b_i  = w_i A_i                                  (n * base_sum_vector(i), :385-391)
//   S_j  = zlow_j sum_{i>j} b_i + w_j sum_{i<j} b_i g_i + zdiag_j b_j     (n * the i-loop at :404-407)
//   pi   = ( sum_j S_j B_j / n + sum_j y_j w_j X_j ) / n                  (:409-419)
// sum_j S_j B_j is accumulated in ONE forward sweep with two running prefixes,
//   Pre_j = sum_{i<j} b_i g_i,   PreB_j = sum_{i<j} B_i zlow_i,
//   sum_j S_j B_j = sum_j [ B_j (w_j Pre_j + zdiag_j b_j) + b_j PreB_j ],
// so a chunk of intervals only needs the two prefixes of the chunks before it.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define COAL_HD __host__ __device__ __forceinline__
#else
#define COAL_HD inline
#endif

namespace coalgpu {

// Stride (in doubles) between the emission-table rows of consecutive distances (must be even: rows are copied with
// 16-byte granularity). Padding to 18 removes shared-memory bank conflicts but was 0.7 % slower (larger copies).
#ifndef COAL_ESTRIDE
#define COAL_ESTRIDE 16
#endif
constexpr int kEStride = COAL_ESTRIDE;

// packed class: count | (dA+1) << 16 | (dB+1) << 24   (dA/dB = -1 when the locus is not ancestral)
COAL_HD uint32_t pack_class(uint32_t count, int dA, int dB) { return count | ((uint32_t)(dA + 1) << 16) | ((uint32_t)(dB + 1) << 24); }

struct ChunkPartial {
    double acc;      // sum over the chunk of B_j (w_j Pre'_j + zdiag_j b_j) + b_j PreB'_j with chunk-local prefixes
    double acc2;     // sum over the chunk of yw_j X_j
    double pre;      // sum over the chunk of b_i g_i
    double preB;     // sum over the chunk of B_i zlow_i
    double sBw;      // sum over the chunk of B_j w_j
    double sb;       // sum over the chunk of b_j
};

struct ClassTotals {
    uint32_t n;       // all lineages in the conditioning configuration
    uint32_t ncompA;  // lineages in classes ancestral at A
    uint32_t ncompB;
};

// Forward sweep over intervals [i0, i1) for one query, in blocks of PER intervals (PER divides i1 - i0).
// The loop nest is class-outer / interval-inner so that the class word is decoded once per block and the
// PER table reads of a class are independent loads. Emission tables are stored distance-major
// (E[d*kEStride + i], see coal_tables.h), i.e. the intervals of one class are contiguous.
// `cls` is read with stride `cstride` (32 on the device: one shared-memory column per query).
template <int PER, class LoadF>
COAL_HD ChunkPartial chunk_forward(LoadF ld, const double* w, int off_yw, int off_zlow, int off_g, int off_zdiag,
                                   int off_EA, int off_EB, const uint32_t* cls, int cstride, int nH,
                                   ClassTotals tot, double invA, double invB, double pow2negLA, double pow2negLB, int i0, int i1) {
    // invA = 1/ncompA, invB = 1/ncompB (0 when the count is 0): the device reads them from a table
    ChunkPartial p = {0, 0, 0, 0, 0, 0};
    const double npartA = (double)(tot.n - tot.ncompA), npartB = (double)(tot.n - tot.ncompB);
    for (int ib = i0; ib < i1; ib += PER) {
        double aAll[PER], bAll[PER], x[PER], aAo[PER], bBo[PER];
#pragma unroll
        for (int ii = 0; ii < PER; ii++) { aAll[ii] = 0; bAll[ii] = 0; x[ii] = 0; aAo[ii] = 0; bBo[ii] = 0; }
        for (int k = 0; k < nH; k++) {
            const uint32_t v = cls[k * cstride];
            const double c = (double)(v & 0xffffu);
            const int dA = (int)((v >> 16) & 0xffu) - 1, dB = (int)(v >> 24) - 1;
            const int ea_base = off_EA + (dA < 0 ? 0 : dA) * kEStride + ib, eb_base = off_EB + (dB < 0 ? 0 : dB) * kEStride + ib;
            const double ca0 = dA >= 0 ? c : 0.0, cb0 = dB >= 0 ? c : 0.0;
#pragma unroll
            for (int ii = 0; ii < PER; ii++) {
                const double eb = ld(eb_base + ii);
                const double ca = ca0 * ld(ea_base + ii), cb = cb0 * eb;
                aAll[ii] += ca; bAll[ii] += cb; x[ii] += ca * (dB >= 0 ? eb : 0.0);
                if (dB < 0) aAo[ii] += ca;
                if (dA < 0) bBo[ii] += cb;
            }
        }
#pragma unroll
        for (int ii = 0; ii < PER; ii++) {
            const int i = ib + ii;
            const double PA = tot.ncompA ? aAll[ii] * invA : pow2negLA;
            const double PB = tot.ncompB ? bAll[ii] * invB : pow2negLB;
            const double Ai = aAll[ii] + npartA * PA, Bi = bAll[ii] + npartB * PB;
            const double Xi = x[ii] + PB * aAo[ii] + PA * bBo[ii];
            const double wi = w[i];
            const double bi = wi * Ai;
            p.acc += Bi * (wi * p.pre + ld(off_zdiag + i) * bi) + bi * p.preB;
            p.acc2 += ld(off_yw + i) * Xi;
            p.sBw += Bi * wi;
            p.sb += bi;
            p.pre += bi * ld(off_g + i);
            p.preB += Bi * ld(off_zlow + i);
        }
    }
    return p;
}

// Contribution of a chunk once the prefixes of all earlier chunks are known.
COAL_HD double chunk_finish(const ChunkPartial& p, double pre_before, double preB_before) {
    return p.acc + pre_before * p.sBw + preB_before * p.sb;
}

// pi from the sums over all chunks
COAL_HD double pismc_finish(double sum_acc, double sum_acc2, uint32_t n) {
    const double inv_n = 1.0 / (double)n;
    return (sum_acc * inv_n + sum_acc2) * inv_n;
}

// One-locus query (ConditionalSamplingHMM.cpp:332-370): with hist[d] = lineages at distance d among the
// classes ancestral at that locus, pi = sum_d hist[d] wE[d] / ncomp, wE[d] = sum_i w_i E[i][d]
// (the non-ancestral lineages enter through the mean of the complete classes, which cancels).
COAL_HD double marginal_finish_rcp(double num, double inv_ncomp, uint32_t ncomp, uint32_t n_cond, double pow2negL, double wsum) {
    if (n_cond == 0) return pow2negL;
    if (ncomp == 0) return pow2negL * wsum;
    return num * inv_ncomp;
}
COAL_HD double marginal_finish(double num, uint32_t ncomp, uint32_t n_cond, double pow2negL, double wsum) {
    if (n_cond == 0) return pow2negL;               // empty configuration, :277-298
    if (ncomp == 0) return pow2negL * wsum;         // nothing ancestral at this locus, :249-252
    return num / (double)ncomp;
}

// ---- target density of one history from per-epoch sufficient statistics --------------------------------
// ProposalEngine.cpp:457-546 adds, per event and grid point, log(kappa*ev/D) + log(rate) - rate*t with
// rate = D/(2 N tau): D cancels except for the first event after a collection (postCollection, :500-546)
// and at the collection boundary itself (:735-747). Everything else is linear in (theta, rho, 1/N).
struct EpochStats {
    double slog;        // sum over events of log(event constant)
    double s_nn;        // sum of n(n-1) * t over events and the boundary
    double s_a;         // sum of ((A+C) LA + (B+C) LB) * t
    double s_c;         // sum of C * t
    double post_nn, post_a, post_c;   // counts of the first event after the collection
    double bnd_nn, bnd_a, bnd_c;      // counts at the collection boundary that ends the epoch
    int n_mut, n_rec;   // kappa = theta / rho events
    int n_rate;         // terms carrying -log(2 N tau): non-post events + boundary
    int has_post, has_bnd;
};

COAL_HD void epoch_stats_clear(EpochStats& s) {
    s.slog = s.s_nn = s.s_a = s.s_c = 0; s.post_nn = s.post_a = s.post_c = 0; s.bnd_nn = s.bnd_a = s.bnd_c = 0;
    s.n_mut = s.n_rec = s.n_rate = s.has_post = s.has_bnd = 0;
}

template <class LogF>
COAL_HD double epoch_log_density(const EpochStats& s, double theta, double rho, double inv2Ntau, double log_theta,
                                 double log_rho, double log_2Ntau, LogF lg) {
    double v = s.slog + s.n_mut * log_theta + s.n_rec * log_rho - s.n_rate * log_2Ntau
             - (s.s_nn + theta * s.s_a + rho * s.s_c) * inv2Ntau;
    if (s.has_post) v -= lg(s.post_nn + s.post_a * theta + s.post_c * rho);
    if (s.has_bnd) v += lg(s.bnd_nn + s.bnd_a * theta + s.bnd_c * rho);
    return v;
}

}  // namespace coalgpu
