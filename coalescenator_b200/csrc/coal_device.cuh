// Device-side data layout and kernels of the importance-sampling likelihood path (sm_100a).
//
// K1 draws coalescent histories (ProposalEngine::calculateLikelihood, /root/reference/src/ProposalEngine.cpp:643-861). Its
// shipped form is the lockstep kernel of coal_lockstep.cuh: a CTA advances 16 histories in lockstep -- serial sections with
// one thread per history, cooperative passes with 4 lanes per history over its type store, and query phases with one lane
// per conditional-sampling query of ANY of the CTA's histories. coal_kernels.cu keeps the group form (one group of W = 8, 16
// or 32 lanes per history) as the fallback. Common to both: the ancestral configuration (TypeContainer,
// TypeContainer.cpp:42-591) lives in shared memory as a slot array of bit-packed haplotypes; the candidate backward moves
// of one event (ProposalEngine.cpp:134-566) are scored across lanes:
//  * marginal (one-locus) conditional sampling probabilities collapse to a dot product of a distance
//    histogram with a precomputed weight vector (ConditionalSamplingHMM.cpp:332-370),
//  * two-locus ones (the 16-interval forward pass, ConditionalSamplingHMM.cpp:372-424) are bilinear forms over the
//    class distances with two small per-row matrices in which the interval axis is already summed out (coal_math.h),
//    one query per lane,
//  * table rows are read through L1 (ld.global.nc); the -DCOAL_TMA build stages the rows an event reads into shared
//    memory by TMA bulk copies (cp.async.bulk + mbarrier) instead -- measured slower for the lockstep form (coal_kernels.cu).
// fp64 throughout; no tensor cores (per query the contraction is (LA+1) x (LB+1) with per-lane count vectors).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/coalgpu.h"

namespace coalgpu {

constexpr int kQd = 16;
constexpr unsigned kFull = 0xffffffffu;

// per-epoch driving constants (ProposalEngine.cpp:657-665, 759-765)
struct EpochConst {
    double dN, dtheta, drho, dthA, dthB, inv2Ntau;   // inv2Ntau = 1/(2*dN*tau)
    double pad0, pad1;
};

// per (epoch, grid point) target constants (ProposalEngine.cpp:462-465)
struct GridConst {
    double theta, rho, inv2Ntau, log_theta, log_rho, log_2Ntau;
};

struct DevProblem {
    int T, LA, LB, Ltot, G;
    int n_max, kmax, row_doubles, pvmax, warp_bytes;
    int SB, off_wEA, off_wEB;     // device row (coal_tables.h): MY[(LA+2)][SB = LB+2][2] | wEA[LA+1] | wEB[LB+1], row_doubles in all
    int stage_doubles;            // -DCOAL_TMA build: doubles of a row staged into shared memory per event by TMA; 0: rows are read through L1
    unsigned long long maskA, maskB;
    double perm_const;        // sum_t probabilitySamplePermutation (ProposalEngine.cpp:88-118,822-828)
    double pow2negLA, pow2negLB, pow2negL, wsum;
    const double* rows;       // [n_sets][n_max][row_doubles]
    const int* set_of_epoch;  // [T]
    const double* times;      // [T]
    const EpochConst* ep;     // [T]
    const GridConst* grid;    // [T][G]
    const double* lfact;      // [n_max+2] log factorials
    const double* rcp;        // [n_max+2] 1/k (rcp[0] = 0): fp64 division is a long instruction sequence
    const int* type_off;      // [T+1]
    const unsigned long long* type_w;
    const unsigned int* type_cg;   // count | gamma << 30, duplicates merged, insertion order
};

struct SampleArgs {
    unsigned long long seed, first, n;
    double* weights;          // [n][G]
    double* tmrca;            // [n]
    double* lineages;         // [n][T]
    unsigned int* nevents;    // [n]
    coalgpu_event* events;    // [n_trace][max_events]
    unsigned int* event_counts;
    unsigned int n_trace, max_events;
    unsigned long long* counters;   // [0] next history, [1] events, [2] pismc, [3] overflow, [4] two-locus queries, [5] their classes,
                                    // [6] histories deferred by the reduced-capacity pass, [7] next entry of `list`,
                                    // [8] most distinct types any completed history held (max over the sampler's life)
    // Two-pass capacity scheme (coal_api.cu): the first pass runs with a type store smaller than the worst case and
    // appends the indices of the histories that outgrow it to defer_list (count in counters[6]); the second pass
    // draws exactly those histories (list != nullptr, count read from list_count) with the full capacity.
    unsigned long long* defer_list;
    const unsigned long long* list;
    const unsigned long long* list_count;
    unsigned char* gstore;    // capacity tier of the lockstep form: the launch's arena of per-CTA type stores in global memory (coal_lockstep.cuh: ls_layout)
    // Score mode (coalgpu_score_events, the parity seam of the device event code): "history" h is the given configuration
    // h -- its type list is loaded instead of the sampled lineages, the given lineage is chosen instead of drawn, ONE event is
    // scored through the same phases as a drawn event, and the unnormalised candidate probabilities are written out instead
    // of a move being drawn. score_off == nullptr everywhere else.
    const int* score_off;                 // [n+1] ranges into the three arrays below
    const unsigned char* score_g; const unsigned long long* score_w; const unsigned int* score_c;
    const unsigned char* score_xg; const unsigned long long* score_xw; const int* score_epoch;   // [n] chosen lineage, epoch
    double* score_pv; unsigned long long* score_pw; int* score_ncand; int score_stride;          // [n*stride], [n*stride], [n]
};

#ifndef COAL_PHILOX_UNROLL
#define COAL_PHILOX_UNROLL 2
#endif
constexpr int kPhiloxUnroll = COAL_PHILOX_UNROLL;

// ---- Philox4x32-10 ------------------------------------------------------------------------------
// One copy in the binary (the history kernel is instruction-cache sensitive, profiles/r1_v0_*).
__device__ __noinline__ uint4 philox_block(unsigned long long seed, unsigned long long hist, unsigned long long blk) {
    unsigned int k0 = (unsigned int)seed, k1 = (unsigned int)(seed >> 32);
    unsigned int c0 = (unsigned int)hist, c1 = (unsigned int)(hist >> 32), c2 = (unsigned int)blk, c3 = (unsigned int)(blk >> 32);
#pragma unroll kPhiloxUnroll
    for (int r = 0; r < 10; r++) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned int n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// 52 random bits -> (0,1): (bits + 0.5) * 2^-52 has at most 53 significant bits, so it is exact, at least 2^-53 and at
// most 1 - 2^-53 (with 53 bits the +0.5 would round and the largest draw would be exactly 1: a zero waiting time).
__device__ __forceinline__ double u52(unsigned int lo, unsigned int hi) {
    const unsigned long long bits = (((unsigned long long)hi << 32) | lo) >> 12;
    return ((double)bits + 0.5) * (1.0 / 4503599627370496.0);
}
// Sequential draws of one history: draw k lives in block k >> 1, half k & 1; -> (0,1), 52 bits.
struct PhiloxStream {
    unsigned long long seed, hist, k;
    uint4 blk;
    __device__ __forceinline__ void init(unsigned long long s, unsigned long long h) { seed = s; hist = h; k = 0; blk = make_uint4(0, 0, 0, 0); }
    __device__ __forceinline__ double next() {
        double u;
        if ((k & 1) == 0) { blk = philox_block(seed, hist, k >> 1); u = u52(blk.x, blk.y); }
        else u = u52(blk.z, blk.w);
        k++;
        return u;
    }
};

// ---- warp helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ unsigned int warp_sum_u(unsigned int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_incl_scan(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

}  // namespace coalgpu
