// Arithmetic cores shared by the kernels (device) and the host unit-test shim (tests/ compile this
// header with g++ and compare against the oracle on the CPU box, where no GPU exists).
//
// pi_SMC(h | H) for a two-locus query, ConditionalSamplingHMM.cpp:372-424 of the reference. With the classes k of
// TypeContainer::getDifferences (count c_k, distances dA_k, dB_k; -1 = locus not ancestral) the reference computes, per
// absorption-time interval i,
//   A_i  = sum_k c_k eA(i,k)   B_i = sum_k c_k eB(i,k)   X_i = sum_k c_k eA(i,k) eB(i,k)
//   eA(i,k) = EA[i][dA_k] if class k is ancestral at locus A, else the count-weighted mean PA_i of the
//             complete classes (ConditionalSamplingHMM.cpp:234-263), or 2^-LA if there is none
//   b_i  = w_i A_i                                  (n * base_sum_vector(i), :385-391)
//   S_j  = sum_i b_i z(i,j)                         (n * the i-loop at :404-407)
//   pi   = ( sum_j S_j B_j / n + sum_j y_j w_j X_j ) / n                  (:409-419)
// Every term is linear in the emission vectors of locus A and of locus B, and the interval axis is summed out. So the
// 16-interval forward pass collapses into a BILINEAR FORM over distances with two small matrices that depend on the
// table row (epoch, lineage count) only and are built once per sampler:
//   M[dA][dB] = sum_ij EA[dA][i] w_i z(i,j) EB[dB][j]        Y[dA][dB] = sum_j y_j w_j EA[dA][j] EB[dB][j]
// (index LA+1 / LB+1 = the all-ones emission vector, standing for "no class ancestral at this locus"). With
// CA / CB = the classes ancestral at A / at B, ncompA / ncompB their lineage totals:
//   pi = T1 / (ncompA ncompB) + ( T2 + T3 / ncompB + T4 / ncompA ) / n
//   T1 = sum_{k in CA, k' in CB} c_k c_k' M[dA_k][dB_k']         (b^T Z B:   A_i = n/ncompA * sum_{CA} c_k EA[dA_k][i])
//   T2 = sum_{k in CA and CB}    c_k Y[dA_k][dB_k]               (classes ancestral at both loci)
//   T3 = sum_{k in CA \ CB, k' in CB} c_k c_k' Y[dA_k][dB_k']   (A-only classes take the mean B emission of CB)
//   T4 = sum_{k' in CB \ CA, k in CA} c_k c_k' Y[dA_k][dB_k']   (B-only classes likewise)
// about nH^2 fused multiply-adds per query instead of 16 intervals x nH classes x 7 + the interval chain. All terms
// are non-negative, so the regrouping changes the result at the 1e-15 level only (tests: <= 1e-13 relative against
// the values the compiled reference recorded).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define COAL_HD __host__ __device__ __forceinline__
#else
#define COAL_HD inline
#endif

namespace coalgpu {

// Stride (in doubles) between the emission-table rows of consecutive distances in the reference-layout rows
// (coal_tables.h; the kernels read the bilinear rows, not these).
#ifndef COAL_ESTRIDE
#define COAL_ESTRIDE 16
#endif
constexpr int kEStride = COAL_ESTRIDE;

// packed class: count | (dA+1) << 16 | (dB+1) << 24   (dA/dB = -1 when the locus is not ancestral)
COAL_HD uint32_t pack_class(uint32_t count, int dA, int dB) { return count | ((uint32_t)(dA + 1) << 16) | ((uint32_t)(dB + 1) << 24); }

struct ClassTotals {
    uint32_t n;       // all lineages in the conditioning configuration
    uint32_t ncompA;  // lineages in classes ancestral at A
    uint32_t ncompB;
};

struct MYPair { double m, y; };

// One dA row of the bilinear tables from a reference-layout row (yw | zlow | g | zdiag | EA | EB, coal_tables.h):
// out[2*dB] = M[dA][dB], out[2*dB+1] = Y[dA][dB] for dB = 0..LB+1. z(i,j) = zlow[j] (i > j), w[j] g[i] (i < j),
// zdiag[i] (i == j) (ConditionalSamplingHMM.cpp:110-173), so U_j = sum_i EA[dA][i] w_i z(i,j) needs one prefix and
// one suffix sum. Shared by the host table builder and the device table builder (stable emission mode).
COAL_HD void bilinear_row(const double* row, const double* w, int off_yw, int off_zlow, int off_g, int off_zdiag,
                          int off_EA, int off_EB, int LA, int LB, int dA, double* out) {
    constexpr int Q = 16;
    double aw[Q], U[Q];
    for (int i = 0; i < Q; i++) aw[i] = (dA <= LA ? row[off_EA + dA * kEStride + i] : 1.0) * w[i];
    double suf[Q + 1];
    suf[Q] = 0.0;
    for (int i = Q - 1; i >= 0; i--) suf[i] = suf[i + 1] + aw[i];       // suf[i] = sum_{i' >= i} aw
    double pre = 0.0;                                                    // sum_{i < j} aw_i g_i
    for (int j = 0; j < Q; j++) {
        U[j] = w[j] * pre + aw[j] * row[off_zdiag + j] + row[off_zlow + j] * suf[j + 1];
        pre += aw[j] * row[off_g + j];
    }
    for (int dB = 0; dB <= LB + 1; dB++) {
        double m = 0.0, y = 0.0;
        for (int j = 0; j < Q; j++) {
            const double eb = dB <= LB ? row[off_EB + dB * kEStride + j] : 1.0;
            const double ea = dA <= LA ? row[off_EA + dA * kEStride + j] : 1.0;
            m += U[j] * eb;
            y += row[off_yw + j] * ea * eb;
        }
        out[2 * dB] = m; out[2 * dB + 1] = y;
    }
}

// pi_SMC of one two-locus query from its class list (see the header comment). `cls` is read with stride `cstride`;
// ldmy(k) returns the (M, Y) pair number k of the row (k = dA * SB + dB, SB = LB + 2). invA = 1/ncompA, invB = 1/ncompB
// (0 when the count is 0: the device reads reciprocals from a table), inv_n = 1/tot.n. nH >= 1.
template <class LoadMY>
COAL_HD double pismc_bilinear(LoadMY ldmy, int SB, int LA, int LB, const uint32_t* cls, int cstride, int nH,
                              ClassTotals tot, double invA, double invB, double inv_n, double pow2negLA, double pow2negLB) {
    if (tot.ncompA != 0 && tot.ncompB != 0) {
        // Branch-free in the class kinds: a class that is not ancestral at A (B) enters with weight 0 and distance 0, so
        // the lanes of a warp -- one query each -- run the same nH x nH instruction sequence (adding 0 * M changes no sum).
        double t1 = 0.0, ty = 0.0;
        for (int k = 0; k < nH; k++) {
            const uint32_t v = cls[k * cstride];
            const uint32_t a1 = (v >> 16) & 0xffu, b1 = v >> 24;
            const double c = a1 ? (double)(v & 0xffffu) : 0.0;      // B-only class: enters through the k2 loop of the others
            const double fk = b1 == 0 ? invB : 0.0;                  // A-only class: mean B emission of CB (T3)
            const int rowoff = (a1 ? (int)a1 - 1 : 0) * SB;
            double sm = 0.0, sy = 0.0;
            for (int k2 = 0; k2 < nH; k2++) {
                const uint32_t v2 = cls[k2 * cstride];
                const uint32_t a2 = (v2 >> 16) & 0xffu, b2 = v2 >> 24;
                const double c2 = b2 ? (double)(v2 & 0xffffu) : 0.0;
                const MYPair e = ldmy(rowoff + (b2 ? (int)b2 - 1 : 0));
                sm += c2 * e.m;
                const double f = fk + (a2 == 0 ? invA : 0.0);        // B-only partner: mean A emission of CA (T4)
                sy += (k2 == k ? (b2 ? 1.0 : 0.0) : c2 * f) * e.y;   // k2 == k: the class itself, ancestral at both loci (T2)
            }
            t1 += c * sm; ty += c * sy;
        }
        return t1 * invA * invB + ty * inv_n;
    }
    // nothing ancestral at one of the loci: its emission is the constant 2^-L (ConditionalSamplingHMM.cpp:249-252)
    double sm = 0.0, sy = 0.0;
    if (tot.ncompA == 0) {
        for (int k2 = 0; k2 < nH; k2++) {
            const uint32_t v2 = cls[k2 * cstride];
            const int dB2 = (int)(v2 >> 24) - 1;
            if (dB2 < 0) continue;
            const double c2 = (double)(v2 & 0xffffu);
            const MYPair e = ldmy((LA + 1) * SB + dB2);
            sm += c2 * e.m; sy += c2 * e.y;
        }
        return pow2negLA * (sm * invB + sy * inv_n);
    }
    for (int k = 0; k < nH; k++) {
        const uint32_t v = cls[k * cstride];
        const int dA = (int)((v >> 16) & 0xffu) - 1;
        if (dA < 0) continue;
        const double c = (double)(v & 0xffffu);
        const MYPair e = ldmy(dA * SB + LB + 1);
        sm += c * e.m; sy += c * e.y;
    }
    return pow2negLB * (sm * invA + sy * inv_n);
}

// Dense form of the same expression, for short loci: the classes are first scattered into two count vectors over distances,
//   vecA[dA] = cA | aAo << 16   lineages in classes ancestral at A with distance dA | of these, in A-only classes
//   vecB[dB] = cB | bBo << 16   likewise for B
// and t2 = sum over the classes ancestral at both loci of c_k Y[dA_k][dB_k] (T2) is collected on the way. Then
//   pi = invA invB sum_{dA,dB} cA cB M + ( t2 + invB sum_{dA,dB} aAo cB Y + invA sum_{dA,dB} cA bBo Y ) / n
// in (LA+1)(LB+1) iterations whatever the class list looks like: the lanes of a warp -- one query each, all of the same
// locus pair -- run the same trip count, where the class-list form makes every lane wait for the longest list squared.
template <class LoadMY>
COAL_HD double pismc_bilinear_dense(LoadMY ldmy, int SB, int LA, int LB, const uint32_t* vecA, const uint32_t* vecB, int vstride,
                                    double t2, ClassTotals tot, double invA, double invB, double inv_n, double pow2negLA, double pow2negLB) {
    if (tot.ncompA != 0 && tot.ncompB != 0) {
        double t1 = 0.0, ty = 0.0;
        for (int dA = 0; dA <= LA; dA++) {
            const uint32_t va = vecA[dA * vstride];
            const double ca = (double)(va & 0xffffu), ao = (double)(va >> 16);
            double sm = 0.0, s3 = 0.0, s4 = 0.0;
            for (int dB = 0; dB <= LB; dB++) {
                const uint32_t vb = vecB[dB * vstride];
                const double cb = (double)(vb & 0xffffu), bo = (double)(vb >> 16);
                const MYPair e = ldmy(dA * SB + dB);
                sm += cb * e.m; s3 += cb * e.y; s4 += bo * e.y;
            }
            t1 += ca * sm;
            ty += ao * s3 * invB + ca * s4 * invA;
        }
        return t1 * invA * invB + (t2 + ty) * inv_n;
    }
    double sm = 0.0, sy = 0.0;
    if (tot.ncompA == 0) {
        for (int dB = 0; dB <= LB; dB++) {
            const double cb = (double)(vecB[dB * vstride] & 0xffffu);
            const MYPair e = ldmy((LA + 1) * SB + dB);
            sm += cb * e.m; sy += cb * e.y;
        }
        return pow2negLA * (sm * invB + sy * inv_n);
    }
    for (int dA = 0; dA <= LA; dA++) {
        const double ca = (double)(vecA[dA * vstride] & 0xffffu);
        const MYPair e = ldmy(dA * SB + LB + 1);
        sm += ca * e.m; sy += ca * e.y;
    }
    return pow2negLB * (sm * invA + sy * inv_n);
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
    // A target value mu = 0 or rho = 0 (main.cpp:169,181 accept them) has log = -inf: a history without such an event keeps
    // a finite weight there (0 * -inf must not turn into NaN), one with such an event gets weight -inf = probability 0.
    double v = s.slog + (s.n_mut ? s.n_mut * log_theta : 0.0) + (s.n_rec ? s.n_rec * log_rho : 0.0) - s.n_rate * log_2Ntau
             - (s.s_nn + theta * s.s_a + rho * s.s_c) * inv2Ntau;
    if (s.has_post) v -= lg(s.post_nn + s.post_a * theta + s.post_c * rho);
    if (s.has_bnd) v += lg(s.bnd_nn + s.bnd_a * theta + s.bnd_c * rho);
    return v;
}

}  // namespace coalgpu
