// Kernels of the importance-sampling likelihood path. See coal_device.cuh for the design summary and
// DESIGN.md §4 for the data layout and roofline of each kernel.
//
// v1 layout of the code (profiles/r1_v0_*: the first version was instruction-fetch bound, 254 KB of SASS):
// every helper with more than one call site is __noinline__ (adjust_core, dlog, philox_block, pick), every
// query kind has ONE call site, and per-event logarithms are folded into running products.
#include "coal_device.cuh"
#include "coal_math.h"

#include <cstdio>      // device printf of the -DCOAL_DEBUG_BOUNDS build

namespace coalgpu {

namespace {

constexpr int GA = COALGPU_GAMMA_A, GB = COALGPU_GAMMA_B, GC = COALGPU_GAMMA_C;
constexpr unsigned int kCntMask = 0x3fffffffu;
constexpr int kAppended = 0x40000000;
// Table rows: read through L1 (ld.global.nc) by default. -DCOAL_TMA builds the variant that stages the rows an event reads
// into shared memory with cp.async.bulk + mbarrier (round 1's default: +5 % for the group form at equal occupancy). For the
// lockstep form the staging buffers are what limits the CTAs per SM and what takes the shared-memory carve-out away from
// L1: measured on one B200, 24-pair cfg2 mix 927 k (L1) vs 769 k (TMA) histories/s, cfg3 644 k vs 558 k, cfg4 179 k vs
// 141 k (profiles/r2_kernel_iterations.md).
#if defined(COAL_TMA) && !defined(COAL_TMA_ROWS)
#define COAL_TMA_ROWS 1
#endif
#ifndef COAL_SLOT_UNROLL
#define COAL_SLOT_UNROLL 1
#endif
constexpr int kSlotUnroll = COAL_SLOT_UNROLL;   // unroll factor of the loops over stored types

// problem + launch arguments, copied once per CTA from the kernel parameters so that the out-of-line
// helpers can read them without carrying them in registers
__shared__ DevProblem sP;
__shared__ SampleArgs sA;
extern __shared__ __align__(16) unsigned char smem_dyn[];

// per-warp shared-memory state
struct WarpMem {
    unsigned long long* words;   // [kmax]  packed haplotype of each stored type
    double* pv;                  // [pvmax] unnormalised candidate probabilities / lineage-chooser weights
    double* aux;                 // [kmax]  joint conditional probability per coalescence partner
    unsigned int* cg;            // [kmax]  count | gamma << 30
    unsigned int* xc;            // [kmax]  cross counts: C slot: cntA(A part) | cntB(B part) << 16;
                                 //         A slot: C lineages with this A part; B slot: likewise
    unsigned int* plist;         // [kmax]  slot index of each partner
    unsigned int* hist;          // [32][hstride] class histogram / class list of each query of the batch (hstride odd)
    double* sd;                  // [4]     scalar results of the out-of-line helpers
    struct HistScalars* hs;      //         per-history scalars that are touched once per event (kept out of registers)
#ifdef COAL_TMA_ROWS
    double* rowbuf;              // [stage_doubles] table row of the event's two-locus queries, staged by cp.async.bulk
    double* wbuf;                // [2][wlen]     wEA|wEB tails of the two rows the one-locus queries read
    unsigned long long* mbar;    // mbarrier the bulk copies complete on
#endif
};

// running product with the exponent split off: a sum of logs without a log per factor
struct LogProd {
    double mant; int expo;
    __host__ __device__ __forceinline__ void clear() { mant = 1.0; expo = 0; }
    __device__ __forceinline__ void mul(double x) {
        int e;
        mant = frexp(mant * x, &e);
        expo += e;
    }
};

// Scalars of the history a group is drawing. Written by lane 0 of the group once per event / boundary and
// read by all lanes after a group sync; living in shared memory they cost no registers in the hot loops.
struct HistScalars {
    EpochStats st;            // target-density statistics of the current epoch
    LogProd qprod, evprod;    // proposal density (product part), event constants of the epoch
    double qsum, consts, tsum;
    int peak;                 // most slots the type store held during this history (capacity statistics)
};

#ifdef COAL_TMA_ROWS
__host__ __device__ inline int tma_wlen(int Ltot) { return (Ltot + 2 + 1) & ~1; }           // doubles, 16-byte multiple
#endif
__host__ __device__ inline size_t group_mem_bytes(int kmax, int pvmax, int Ltot, int W, int stage_doubles) {
    size_t b = 0;
#ifdef COAL_TMA_ROWS
    b += (size_t)stage_doubles * 8 + 2 * (size_t)tma_wlen(Ltot) * 8 + 16;   // rowbuf, wbuf, mbarrier
#endif
    b += (size_t)kmax * 8;                 // words
    b += (size_t)pvmax * 8;                // pv
    b += (size_t)kmax * 8;                 // aux
    b += (size_t)kmax * 4 * 3;             // cg, xc, plist
    b += (size_t)((Ltot + 2) | 1) * W * 4; // hist: one row per query of a batch (a batch = W queries)
    b += 4 * 8;                            // sd
    b += (sizeof(HistScalars) + 15) & ~(size_t)15;
    return (b + 15) & ~(size_t)15;
}

// A GROUP of W lanes (W = 8, 16 or 32) draws one history; a warp carries 32/W independent groups. Every
// collective below names only the lanes of its own group (mask), so the groups of a warp never wait for each
// other; where their control flow coincides the hardware issues one instruction for all of them.
template <int W>
struct Grp {
    static __device__ __forceinline__ int lane() { return threadIdx.x & 31; }
    static __device__ __forceinline__ int sub() { return threadIdx.x & (W - 1); }
    static __device__ __forceinline__ int base() { return (threadIdx.x & 31) & ~(W - 1); }
    static __device__ __forceinline__ unsigned int mask() { return W == 32 ? 0xffffffffu : (((1u << W) - 1u) << base()); }
    static __device__ __forceinline__ unsigned int ballot(bool p) { return __ballot_sync(mask(), p) >> base(); }
    template <class T> static __device__ __forceinline__ T shfl(T v, int src) { return __shfl_sync(mask(), v, src, W); }
    static __device__ __forceinline__ void sync() { __syncwarp(mask()); }
    static __device__ __forceinline__ unsigned int sum_u(unsigned int v) {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask(), v, o, W);
        return v;
    }
    static __device__ __forceinline__ double incl_scan(double v) {
        const int s = sub();
#pragma unroll
        for (int o = 1; o < W; o <<= 1) {
            const double t = __shfl_up_sync(mask(), v, o, W);
            if (s >= o) v += t;
        }
        return v;
    }
};

template <int W>
__device__ __forceinline__ WarpMem wm() {
    unsigned char* base = smem_dyn + (size_t)(threadIdx.x / W) * sP.warp_bytes;
    const int kmax = sP.kmax, pvmax = sP.pvmax;
    WarpMem m;
#ifdef COAL_TMA_ROWS
    m.rowbuf = reinterpret_cast<double*>(base); base += (size_t)sP.stage_doubles * 8;        // 16-byte aligned: first in the block
    m.wbuf = reinterpret_cast<double*>(base); base += 2 * (size_t)tma_wlen(sP.Ltot) * 8;
    m.mbar = reinterpret_cast<unsigned long long*>(base); base += 16;
#endif
    m.words = reinterpret_cast<unsigned long long*>(base); base += (size_t)kmax * 8;
    m.pv = reinterpret_cast<double*>(base); base += (size_t)pvmax * 8;
    m.aux = reinterpret_cast<double*>(base); base += (size_t)kmax * 8;
    m.cg = reinterpret_cast<unsigned int*>(base); base += (size_t)kmax * 4;
    m.xc = reinterpret_cast<unsigned int*>(base); base += (size_t)kmax * 4;
    m.plist = reinterpret_cast<unsigned int*>(base); base += (size_t)kmax * 4;
    m.hist = reinterpret_cast<unsigned int*>(base); base += (size_t)((sP.Ltot + 2) | 1) * W * 4;
    m.sd = reinterpret_cast<double*>(base); base += 4 * 8;
    m.hs = reinterpret_cast<HistScalars*>(base);
    return m;
}

struct HState {
    int nslots, nA, nB, nC;
    bool overflow;
};

__device__ __noinline__ double dlog(double x) { return log(x); }

#ifdef COAL_TMA_ROWS
// Table rows staged into shared memory by the TMA unit (1-D bulk copies completing on an mbarrier).
__device__ __forceinline__ unsigned int smem_addr(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned int bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
    unsigned int done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ double tab_ld(const double* p) { return *p; }
#else
__device__ __forceinline__ double tab_ld(const double* p) { return __ldg(p); }
#endif

// Haplotype words in the hot loops over stored types: when LA + LB <= 32 only the low half of the 64-bit
// word is populated and the xor / mask / popcount chain runs on 32-bit registers (half the ALU and XU work).
template <bool N32> struct WordOps;
template <> struct WordOps<true> {
    typedef unsigned int T;
    static __device__ __forceinline__ T load(const unsigned long long* words, int j) { return reinterpret_cast<const unsigned int*>(words)[2 * j]; }
    static __device__ __forceinline__ T narrow(unsigned long long w) { return (unsigned int)w; }
    static __device__ __forceinline__ int popc(T x) { return __popc(x); }
};
template <> struct WordOps<false> {
    typedef unsigned long long T;
    static __device__ __forceinline__ T load(const unsigned long long* words, int j) { return words[j]; }
    static __device__ __forceinline__ T narrow(unsigned long long w) { return w; }
    static __device__ __forceinline__ int popc(T x) { return __popcll(x); }
};

// ---- type store ------------------------------------------------------------------------------------

// Change the multiplicity of type (g, w) by delta (TypeContainer::addType / removeType,
// TypeContainer.cpp:409-522) keeping the cross counts of every related slot current. A type that is
// not stored yet is appended; a slot whose count reaches 0 stays until commit().
// Returns the slot (| kAppended when a slot was appended) or -1 when the store is full.
template <int W>
__device__ __noinline__ int adjust_core(int nslots, int g, unsigned long long w, int delta) {
    typedef Grp<W> G;
    const WarpMem m = wm<W>();
    const int sub = G::sub();
    const unsigned long long maskA = sP.maskA, maskB = sP.maskB;
    int found = -1;
    unsigned int newxc = 0;
    const unsigned long long wa = w & maskA, wb = w & maskB;
#pragma unroll 1
    for (int base = 0; base < nslots; base += W) {
        const int j = base + sub;
        const bool in = j < nslots;
        const unsigned long long wj = in ? m.words[j] : 0ull;
        const unsigned int cgj = in ? m.cg[j] : 0u;
        const int gj = (int)(cgj >> 30);
        const unsigned int cj = cgj & kCntMask;
        const unsigned int b = G::ballot(in && gj == g && wj == w);
        if (b) found = base + __ffs(b) - 1;
        if (in) {
            if (g == GA) {
                if (gj == GC && (wj & maskA) == w) { m.xc[j] += (unsigned int)delta; newxc += cj; }
            } else if (g == GB) {
                if (gj == GC && (wj & maskB) == w) { m.xc[j] += (unsigned int)(delta * 65536); newxc += cj; }
            } else {
                if (gj == GA && wj == wa) { m.xc[j] += (unsigned int)delta; newxc += cj; }
                if (gj == GB && wj == wb) { m.xc[j] += (unsigned int)delta; newxc += cj << 16; }
            }
        }
    }
    int ret;
    if (found < 0) {
        newxc = G::sum_u(newxc);
        if (nslots >= sP.kmax) return -1;
        found = nslots;
        if (sub == 0) { m.words[found] = w; m.cg[found] = (unsigned int)g << 30; m.xc[found] = newxc; }
        ret = found | kAppended;
    } else {
        ret = found;
    }
    G::sync();
    if (sub == 0) m.cg[found] += (unsigned int)delta;
    G::sync();
    return ret;
}

template <int W>
__device__ __forceinline__ int adjust(HState& s, int g, unsigned long long w, int delta) {
    int r = adjust_core<W>(s.nslots, g, w, delta);
    if (r < 0) { s.overflow = true; return -1; }
    if (r & kAppended) { s.nslots++; r &= ~kAppended; }
    if (g == GA) s.nA += delta; else if (g == GB) s.nB += delta; else s.nC += delta;
    return r;
}

// Drop empty slots at the end of an event: highest empty slot first, overwritten by the last slot
// (the canonical order rule shared with the oracle's CANONICAL mode).
template <int W>
__device__ __forceinline__ int commit(const WarpMem& m, int nslots) {
    typedef Grp<W> G;
    const int sub = G::sub();
#pragma unroll 1
    for (;;) {
        int z = -1;
#pragma unroll 1
        for (int base = ((nslots - 1) / W) * W; base >= 0 && z < 0; base -= W) {
            const int j = base + sub;
            const bool empty = j < nslots && (m.cg[j] & kCntMask) == 0;
            const unsigned int b = G::ballot(empty);
            if (b) z = base + 31 - __clz(b);
        }
        if (z < 0) break;
        const int last = nslots - 1;
        if (sub == 0 && z != last) { m.words[z] = m.words[last]; m.cg[z] = m.cg[last]; m.xc[z] = m.xc[last]; }
        nslots = last;
        G::sync();
        if (nslots == 0) break;
    }
    return nslots;
}

// Categorical draw over pv[0..n): first k with cum > u*total (strict, TypeContainer.cpp:730) or with
// cum >= u*total (ProposalEngine.cpp:574). Partial sums run in index order (inclusive group scans).
// Returns k; sd[0] = pv[k], sd[1] = total.
template <int W>
__device__ __noinline__ int pick(int n, double u, bool strict) {
    typedef Grp<W> G;
    const WarpMem m = wm<W>();
    const int sub = G::sub();
    double total = 0.0;
#pragma unroll 1
    for (int base = 0; base < n; base += W) {
        const int k = base + sub;
        const double v = k < n ? m.pv[k] : 0.0;
        total = G::shfl(G::incl_scan(v) + total, W - 1);
    }
    const double target = u * total;
    int idx = -1; double pidx = 0.0, carry = 0.0;
#pragma unroll 1
    for (int base = 0; base < n && idx < 0; base += W) {
        const int k = base + sub;
        const double v = k < n ? m.pv[k] : 0.0;
        const double cum = G::incl_scan(v) + carry;
        const bool hit = k < n && (strict ? (cum > target && v > 0.0) : !(cum < target));
        const unsigned int b = G::ballot(hit);
        if (b) { const int l = __ffs(b) - 1; idx = base + l; pidx = G::shfl(v, l); }
        carry = G::shfl(cum, W - 1);
    }
    if (idx < 0) { idx = n - 1; pidx = m.pv[idx]; }   // u*total rounded past the last partial sum
    G::sync();
    if (sub == 0) { m.sd[0] = pidx; m.sd[1] = total; }
    G::sync();
    return idx;
}

// Lineage-chooser weights (TypeContainer::sampleType, TypeContainer.cpp:641-721) into pv[]
template <int W>
__device__ __forceinline__ void type_weights(const WarpMem& m, const HState& s, double thA, double thB, double rho) {
#pragma unroll 1
    for (int j = Grp<W>::sub(); j < s.nslots; j += W) {
        const unsigned int cgj = m.cg[j], xcj = m.xc[j];
        const unsigned int c = cgj & kCntMask;
        const int gj = (int)(cgj >> 30);
        double wgt;
        if (gj == GC) wgt = ((((double)(c - 1) + (thA + thB)) + rho) + (double)(xcj & 0xffffu) + (double)(xcj >> 16)) * (double)c;
        else if (gj == GA) wgt = ((double)(c - 1 + (unsigned int)s.nB + xcj) + thA) * (double)c;
        else wgt = ((double)(c - 1 + (unsigned int)s.nA + xcj) + thB) * (double)c;
        m.pv[j] = wgt;
    }
    Grp<W>::sync();
}

// ---- conditional sampling probabilities ----------------------------------------------------------------

// One-locus query per lane: query (qg in {A,B}, qw), optionally leaving out one copy of the query's own
// type (the H - b_k configurations of ProposalEngine.cpp:323-329). Evaluated against two table rows at once
// (the same distance histogram serves pi(.|H) and pi(.|H + other-locus lineage), ProposalEngine.cpp:120-132).
template <bool N32>
__device__ __forceinline__ void marginal_query(const WarpMem& m, int nslots, int qg, unsigned long long qw64, bool excl,
                                               const double* __restrict__ wE1, const double* __restrict__ wE2,
                                               double& num1, double& num2, unsigned int& ncomp) {
    typedef WordOps<N32> WO;
    num1 = 0.0; num2 = 0.0; ncomp = 0;
    const typename WO::T mask = WO::narrow((qg == GA) ? sP.maskA : sP.maskB), qw = WO::narrow(qw64);
#pragma unroll kSlotUnroll
    for (int j = 0; j < nslots; j++) {
        const typename WO::T wj = WO::load(m.words, j);
        const unsigned int cgj = m.cg[j];
        const int gj = (int)(cgj >> 30);
        unsigned int c = cgj & kCntMask;
        if (gj == qg || gj == GC) {
            if (excl && gj == qg && wj == qw) c -= 1;
            const int d = WO::popc((wj ^ qw) & mask);
            const double cf = (double)c;
            num1 += cf * tab_ld(wE1 + d);
            num2 += cf * tab_ld(wE2 + d);
            ncomp += c;
        }
    }
}

// Two-locus queries of one event, phase 1: class lists (TypeContainer::getDifferences for a C query,
// TypeContainer.cpp:802-886, incl. the dA+dB collapse of SURVEY.md Q1) of up to W queries at once, one
// query per lane: every lane walks the stored types (broadcast reads) and accumulates its own histogram row
// (rows are query-major with an odd stride: conflict-free here, in the compaction and in the forward pass).
//  C chosen:   query qb+q = x (q = 0) or x with site q-1 flipped, against H'
//  A/B chosen: query qb+q = (x, b_k) with b_k = partner plist[qb+q], against H' - b_k
// On return lane q < nqb holds (nH, totals) of query q and its class list sits in hist[q*hstride ...].
// (A lanes-over-types variant with match.any + group reductions was 1.8x slower: profiles/r1_v2_*.)
template <int W, bool N32>
__device__ __forceinline__ void build_classes(const WarpMem& m, int ns, int nqb, int qb, bool isC, unsigned long long xw, int pg,
                                              int& nH_out, ClassTotals& tot_out) {
    const int sub = Grp<W>::sub();
    const int nsum = sP.Ltot + 2, hstride = nsum | 1;
    typedef WordOps<N32> WO;
    const typename WO::T maskA = WO::narrow(sP.maskA), maskB = WO::narrow(sP.maskB);
    unsigned int* mine = m.hist + sub * hstride;
    const bool active = sub < nqb;
    unsigned long long qw64 = xw, delw64 = 0ull; int delg = -1;
    int nH = 0;
    ClassTotals tot; tot.n = 0; tot.ncompA = 0; tot.ncompB = 0;
    if (active) {
        const int qi = qb + sub;
        if (isC) { if (qi > 0) qw64 = xw ^ (1ull << (qi - 1)); }
        else { delw64 = m.words[m.plist[qi]]; delg = pg; qw64 = xw | delw64; }
        const typename WO::T qw = WO::narrow(qw64), delw = WO::narrow(delw64);
#pragma unroll 1
        for (int sidx = 0; sidx < nsum; sidx++) mine[sidx] = 0xffff0000u;
#pragma unroll kSlotUnroll
        for (int j = 0; j < ns; j++) {
            const typename WO::T wj = WO::load(m.words, j);
            const unsigned int cgj = m.cg[j];
            const int gj = (int)(cgj >> 30);
            unsigned int c = cgj & kCntMask;
            if (gj == delg && wj == delw) c -= 1;
            if (c > 0) {
                const typename WO::T y = wj ^ qw;
                const int da = WO::popc(y & maskA), db = WO::popc(y & maskB);
                int sidx; unsigned int rep;
                if (gj == GA) { sidx = da; rep = 0; } else if (gj == GB) { sidx = db; rep = 1; } else { sidx = da + db + 1; rep = 2 + da; }
                const unsigned int v = mine[sidx];
                mine[sidx] = ((v & 0xffffu) + c) | (min(v >> 16, rep) << 16);
            }
        }
        // compact the classes in place (ascending dA+dB) and collect the totals
#pragma unroll 1
        for (int sidx = 0; sidx < nsum; sidx++) {
            const unsigned int v = mine[sidx];
            const unsigned int cnt = v & 0xffffu;
            if (cnt) {
                const unsigned int rep = v >> 16;
                int dA, dB;
                if (rep == 0) { dA = sidx; dB = -1; } else if (rep == 1) { dA = -1; dB = sidx; } else { dA = (int)rep - 2; dB = sidx - 1 - dA; }
                mine[nH] = pack_class(cnt, dA, dB);
                nH++;
                tot.n += cnt;
                if (dA >= 0) tot.ncompA += cnt;
                if (dB >= 0) tot.ncompB += cnt;
            }
        }
    }
    Grp<W>::sync();
    nH_out = nH; tot_out = tot;
}

// Two-locus queries, phase 2: pi_SMC of the query whose class list the lane has just built (the forward pass of
// ConditionalSamplingHMM.cpp:372-424 as a bilinear form, coal_math.h). STAGED: the row sits in the group's shared
// memory (one 16-byte load per (M, Y) pair); otherwise it is read through L1 from the sampler's table in HBM / L2.
template <bool STAGED>
__device__ __forceinline__ double two_locus_pi(const unsigned int* cls, int nH, ClassTotals tot, const double* __restrict__ row, unsigned int n_cond) {
    if (n_cond == 0 || nH == 0) return sP.pow2negL;      // empty configuration, ConditionalSamplingHMM.cpp:277-298
    auto ldmy = [row](int k) {
        const double2 v = STAGED ? *reinterpret_cast<const double2*>(row + 2 * k) : __ldg(reinterpret_cast<const double2*>(row + 2 * k));
        MYPair e; e.m = v.x; e.y = v.y; return e;
    };
    return pismc_bilinear(ldmy, sP.SB, sP.LA, sP.LB, cls, 1, nH, tot, __ldg(sP.rcp + tot.ncompA), __ldg(sP.rcp + tot.ncompB),
                          __ldg(sP.rcp + tot.n), sP.pow2negLA, sP.pow2negLB);
}

__device__ __forceinline__ const double* row_ptr(int set, int n_all) {
    const int n = n_all < 1 ? 1 : n_all;   // n_all == 0 is never dereferenced for a result
    return sP.rows + ((size_t)set * sP.n_max + (n - 1)) * sP.row_doubles;
}

__device__ __forceinline__ double logprod_value(const LogProd& p) { return dlog(p.mant) + (double)p.expo * 0.69314718055994530942; }

// ---- the history loop --------------------------------------------------------------------------------------
// One flat loop per group: fetch a history index, draw the history event by event, write its outputs, fetch
// the next. Keeping everything in ONE loop body lets the groups that share a warp reconverge at every
// iteration (structured control flow) instead of drifting into different call frames.
template <int W, bool N32>
__device__ __forceinline__ void run_group() {
    typedef Grp<W> G;
    const WarpMem m = wm<W>();
    const int sub = G::sub();
    const int T = sP.T, LA = sP.LA, LB = sP.LB, Ltot = sP.Ltot, NG = sP.G;
    unsigned long long ev_total = 0, pi_total = 0, cq_total = 0;
    unsigned int nh_lane = 0;

    // per-history state
    bool have = false;
    unsigned long long h = 0;
    PhiloxStream rng; rng.init(0, 0);
#ifdef COAL_TMA_ROWS
    unsigned int tma_phase = 0;
    if (sub == 0) mbar_init(m.mbar);
    G::sync();
#endif
    HState s; s.nslots = 0; s.nA = s.nB = s.nC = 0; s.overflow = false;
    bool tracing = false, injecting = false, post = true;
    unsigned int nev = 0, npi = 0, ncq = 0;
    int inj_lo = 0, inj_hi = 0, e = 0, tset = 0;
    HistScalars* const hs = m.hs;
    double* wout = nullptr;
    const unsigned long long n_todo = sA.list ? *sA.list_count : sA.n;

#pragma unroll 1
    for (;;) {
        if (!have) {
            unsigned long long h0 = 0;
            if (sub == 0) h0 = atomicAdd(sA.counters + (sA.list ? 7 : 0), 1ull);
            h = G::shfl(h0, 0);
            if (h >= n_todo) break;
            if (sA.list) h = sA.list[h];
            have = true;
            rng.init(sA.seed, sA.first + h);
            s.nslots = 0; s.nA = s.nB = s.nC = 0; s.overflow = false;
            tracing = sA.events != nullptr && h < sA.n_trace;
            nev = 0; npi = 0; ncq = 0;
            // initial configuration (TypeContainer ctor, TypeContainer.cpp:42-99), then injections at boundaries
            inj_lo = sP.type_off[0]; inj_hi = sP.type_off[1];
            injecting = true; post = true;
            e = 0; tset = sP.set_of_epoch[0];
            if (sub == 0) {
                hs->qsum = 0.0; hs->consts = 0.0; hs->tsum = 0.0; hs->peak = 0;
                hs->qprod.clear(); hs->evprod.clear(); epoch_stats_clear(hs->st);
            }
            G::sync();
            wout = sA.weights ? sA.weights + h * (unsigned long long)NG : nullptr;
            if (wout) for (int g = sub; g < NG; g += W) wout[g] = 0.0;
        }
        bool end_now = false;

        if (injecting) {
            // TypeContainer::addTypeMap :525-591 + probabilityCombinatorics ProposalEngine.cpp:66-86
            const int a0 = s.nA, b0 = s.nB, c0 = s.nC;
            double comb = 0.0;
#pragma unroll 1
            for (int k = inj_lo; k < inj_hi; k++) {
                const unsigned int cg = sP.type_cg[k];
                const int add = (int)(cg & kCntMask);
                const int slot = adjust<W>(s, (int)(cg >> 30), sP.type_w[k], add);
                if (slot < 0) break;
                const int now = (int)(m.cg[slot] & kCntMask);
                comb += sP.lfact[now] - sP.lfact[add] - sP.lfact[now - add];
            }
            if (e > 0 && !s.overflow) {
                comb -= sP.lfact[s.nA] - sP.lfact[s.nA - a0] - sP.lfact[a0];
                comb -= sP.lfact[s.nB] - sP.lfact[s.nB - b0] - sP.lfact[b0];
                comb -= sP.lfact[s.nC] - sP.lfact[s.nC - c0] - sP.lfact[c0];
                if (sub == 0) hs->consts += comb;
            }
            if (sub == 0 && s.nslots > hs->peak) hs->peak = s.nslots;
            injecting = false;
            end_now = s.overflow;
        }

        if (!end_now) {
            // ---- choose a lineage and a waiting time (ProposalEngine.cpp:672-687 / 703-716) ----
            const EpochConst* __restrict__ ecp = sP.ep + e;
            const double dthA = __ldg(&ecp->dthA), dthB = __ldg(&ecp->dthB), drho_e = __ldg(&ecp->drho);
            type_weights<W>(m, s, dthA, dthB, drho_e);
            const int sx = pick<W>(s.nslots, rng.next(), true);
            const double hap_num = m.sd[0], hap_den = m.sd[1];      // p_hap = hap_num / hap_den, folded into qfac
            const int n = s.nA + s.nB + s.nC, Ac = s.nA, Bc = s.nB, Cc = s.nC;
            const double nn = (double)((unsigned int)n * (unsigned int)(n - 1));
            const double Dd = nn + (double)(Ac + Cc) * dthA + (double)(Bc + Cc) * dthB + drho_e * (double)Cc;
            const double rate = Dd * __ldg(&ecp->inv2Ntau);
            const double tsum = hs->tsum;
            const double wait = -dlog(rng.next()) / rate;
            const double aa = (double)((Ac + Cc) * LA + (Bc + Cc) * LB);

            // ---- collection boundary (ProposalEngine.cpp:721-783) or end of the history (:792, :817-858) ----
            const bool boundary = (e < T - 1) && !(tsum + wait < sP.times[e + 1]);
            const bool finish = (e == T - 1) && (n <= 5);
            if (boundary || finish) {
                if (sub == 0) {
                    EpochStats& st = hs->st;
                    if (boundary) {
                        const double excess = sP.times[e + 1] - tsum;
                        hs->tsum = sP.times[e + 1];
                        hs->qsum += -rate * excess;
                        st.s_nn += nn * excess; st.s_a += aa * excess; st.s_c += (double)Cc * excess;
                        st.bnd_nn = nn; st.bnd_a = aa; st.bnd_c = (double)Cc; st.has_bnd = 1; st.n_rate += 1;
                    }
                    st.slog = logprod_value(hs->evprod);
                }
                G::sync();
                if (wout) {
                    const EpochStats st = hs->st;
                    const GridConst* gc = sP.grid + (size_t)e * NG;
#pragma unroll 1
                    for (int g = sub; g < NG; g += W) {
                        const GridConst c = gc[g];
                        wout[g] += epoch_log_density(st, c.theta, c.rho, c.inv2Ntau, c.log_theta, c.log_rho, c.log_2Ntau,
                                                     [](double x) { return dlog(x); });
                    }
                }
                G::sync();
                if (sub == 0) { epoch_stats_clear(hs->st); hs->evprod.clear(); }
                if (sA.lineages && sub == 0) sA.lineages[h * (unsigned long long)T + e] = (double)n;
                if (finish) {
                    end_now = true;
                } else {
                    e++;
                    inj_lo = sP.type_off[e]; inj_hi = sP.type_off[e + 1];
                    injecting = true;
                    post = true;
                    tset = sP.set_of_epoch[e];
                }
            } else {
                // ---- one event (ProposalEngine.cpp:134-566) ----
                const unsigned long long xw = m.words[sx];
                const unsigned int xcg = m.cg[sx], xxc = m.xc[sx];
                const int xg = (int)(xcg >> 30);
                const unsigned int xcount = xcg & kCntMask;
#ifdef COAL_TMA_ROWS
                // Stage what this event reads from the tables — the full row of its two-locus queries (n-1 lineages
                // for a C lineage, n-2 for a one-locus lineage) and the wEA|wEB tails of the two rows of its
                // one-locus queries — with three TMA bulk copies; they overlap the count / remove / partner-list work.
                if (sub == 0) {
                    const bool c_chosen = (xg == GC);
                    const double* g_m1 = row_ptr(tset, n - 1);
                    const double* g_alt = row_ptr(tset, c_chosen ? n : n - 2);
                    const unsigned int rb = (unsigned int)sP.stage_doubles * 8u, wb = (unsigned int)(sP.row_doubles - sP.off_wEA) * 8u;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect(m.mbar, rb + 2u * wb);
                    if (rb) bulk_g2s(m.rowbuf, c_chosen ? g_m1 : g_alt, rb, m.mbar);
                    bulk_g2s(m.wbuf, g_m1 + sP.off_wEA, wb, m.mbar);
                    bulk_g2s(m.wbuf + tma_wlen(Ltot), g_alt + sP.off_wEA, wb, m.mbar);
                }
#endif
                const unsigned long long xa = xw & sP.maskA, xb = xw & sP.maskB;
                unsigned int Cij = 0, Ci = 0, Cj = 0, Ai = 0, Bj = 0;
                if (xg == GC) {
                    Cij = xcount; Ai = xxc & 0xffffu; Bj = xxc >> 16;
                    unsigned int ci = 0, cj = 0;
#pragma unroll 1
                    for (int j = sub; j < s.nslots; j += W) {
                        const unsigned int cgj = m.cg[j];
                        if ((int)(cgj >> 30) == GC) {
                            const unsigned long long wj = m.words[j];
                            if ((wj & sP.maskA) == xa) ci += cgj & kCntMask;
                            if ((wj & sP.maskB) == xb) cj += cgj & kCntMask;
                        }
                    }
                    Ci = G::sum_u(ci); Cj = G::sum_u(cj);
                } else if (xg == GA) { Ai = xcount; Ci = xxc; } else { Bj = xcount; Cj = xxc; }

                adjust<W>(s, xg, xw, -1);               // H' = H - x  (:168)
                const double dtheta_e = __ldg(&ecp->dtheta);
                const int ns = s.nslots;
                const bool isC = (xg == GC), isA = (xg == GA);
                const int Lx = isC ? Ltot : (isA ? LA : LB), offb = (xg == GB) ? LA : 0;
                const int pg = isA ? GB : GA;
                const unsigned int own = isA ? Ai : Bj, ownC = isA ? Ci : Cj;
#ifdef COAL_TMA_ROWS
                // pointers laid out so that (ptr + off_wEA/off_wEB) lands in the staged tails
                const int wlen = tma_wlen(Ltot);
                const double* row_m1 = m.wbuf - sP.off_wEA;
                const double* row_alt = m.wbuf + wlen - sP.off_wEA;
#else
                const double* row_m1 = row_ptr(tset, n - 1);
                const double* row_alt = row_ptr(tset, isC ? n : n - 2);      // second row of the one-locus queries
#endif
                const unsigned int nc_alt = isC ? (unsigned int)n : (unsigned int)(n - 2 < 0 ? 0 : n - 2);

                // partner list = stored types of the other one-locus class, slot order (:300 / :377)
                int np = 0;
                if (!isC) {
#pragma unroll 1
                    for (int base = 0; base < ns; base += W) {
                        const int j = base + sub;
                        const bool is = j < ns && (int)(m.cg[j] >> 30) == pg && (m.cg[j] & kCntMask) > 0;
                        const unsigned int b = G::ballot(is);
                        if (is) m.plist[np + __popc(b & ((1u << sub) - 1))] = (unsigned int)j;
                        np += __popc(b);
                    }
                    G::sync();
                }

                // ---- one-locus queries, one per lane ---------------------------------------------------------
                //  C chosen: lane 0 = A part, lane 1 = B part, rows n-1 and n (:198-238 with :120-132)
                //  A/B chosen: 0 = x (rows n-1, n-2), 1..Lx = mutants (row n-1), then one per partner b_k:
                //              pi(b_k | H'-b_k+x) on row n-1 and pi(b_k | H'-b_k) on row n-2 (:304-330 / :381-406)
#ifdef COAL_TMA_ROWS
                mbar_wait(m.mbar, tma_phase); tma_phase ^= 1u;
#endif
                const int nm = isC ? 2 : 1 + Lx + np;
                double pi0 = 0.0, mA1 = 0.0, mA2 = 0.0, mB1 = 0.0, mB2 = 0.0;
#pragma unroll 1
                for (int q0 = 0; q0 < nm; q0 += W) {
                    const int qi = q0 + sub;
                    const bool valid = qi < nm;
                    const bool partner = !isC && valid && qi > Lx;
                    int qg; unsigned long long qw; bool excl = false;
                    if (isC) { qg = (qi == 1) ? GB : GA; qw = (qi == 1) ? xb : xa; }
                    else if (partner) { qg = pg; qw = m.words[m.plist[qi - Lx - 1]]; excl = true; }
                    else { qg = xg; qw = (valid && qi >= 1) ? (xw ^ (1ull << (offb + qi - 1))) : xw; }
                    const int offw = (qg == GA) ? sP.off_wEA : sP.off_wEB;
                    const double powl = (qg == GA) ? sP.pow2negLA : sP.pow2negLB;
                    double num1 = 0.0, num2 = 0.0; unsigned int ncomp = 0;
                    if (valid) marginal_query<N32>(m, ns, qg, qw, excl, row_m1 + offw, row_alt + offw, num1, num2, ncomp);
                    // fp64 division is a ~25-instruction sequence: reciprocals of lineage counts come from a table
                    const double inv_nc = __ldg(sP.rcp + ncomp);
                    const double v1 = marginal_finish_rcp(num1, inv_nc, ncomp, (unsigned int)(n - 1), powl, sP.wsum);
                    const double v2 = marginal_finish_rcp(num2, inv_nc, ncomp, nc_alt, powl, sP.wsum);
                    if (q0 == 0) {
                        mA1 = G::shfl(v1, 0); mA2 = G::shfl(v2, 0);
                        mB1 = G::shfl(v1, 1); mB2 = G::shfl(v2, 1);
                        if (!isC) pi0 = mA1;
                    }
                    if (!isC && valid) {
                        if (qi >= 1 && qi <= Lx) m.pv[qi - 1] = dtheta_e * (double)own * v1;   // x 1/pi0 below
                        // joint: 0.5 [pi(x|H'-b) pi(b|H'-b+x) + pi(b|H'-b) pi(x|H')]   (:326 / :403 with :120-132)
                        if (partner) m.aux[qi - Lx - 1] = 0.5 * mA2 * v1 + 0.5 * v2 * pi0;
                    }
                }
                G::sync();

                // ---- two-locus queries -----------------------------------------------------------------------
                //  C chosen: x and its Ltot one-site mutants against H' (:169, :179-187)
                //  A/B chosen: the joined haplotype (x, b_k) against H' - b_k (:325 / :402)
                const int nq = isC ? Ltot + 1 : np;
                constexpr bool kStaged =
#ifdef COAL_TMA_ROWS
                    N32;        // rows of up to 32 sites fit the staging buffer; longer loci read their (quadratically larger) rows through L1
#else
                    false;
#endif
#ifdef COAL_TMA_ROWS
                const double* crow = kStaged ? m.rowbuf : (isC ? row_ptr(tset, n - 1) : row_ptr(tset, n - 2));
#else
                const double* crow = isC ? row_ptr(tset, n - 1) : row_ptr(tset, n - 2);
#endif
                const unsigned int cn = isC ? (unsigned int)(n - 1) : nc_alt;
                const int hstride = (Ltot + 2) | 1;
#pragma unroll 1
                for (int qb = 0; qb < nq; qb += W) {
                    const int nqb = min(W, nq - qb);
                    int my_nH; ClassTotals my_tot;
                    build_classes<W, N32>(m, ns, nqb, qb, isC, xw, pg, my_nH, my_tot);
                    nh_lane += (unsigned int)my_nH; ncq += (unsigned int)nqb;
                    if (sub < nqb) {       // one query per lane: the lane that built the class list evaluates it
                        const double pi = two_locus_pi<kStaged>(m.hist + sub * hstride, my_nH, my_tot, crow, cn);
                        const int qi = qb + sub;
                        if (isC) { if (qi == 0) m.sd[2] = pi; else m.pv[qi - 1] = pi; }
                        else {
                            const int pslot = (int)m.plist[qi];
                            m.pv[Lx + qi] = (double)own * (double)(m.cg[pslot] & kCntMask) * pi / m.aux[qi];
                        }
                    }
                    G::sync();
                }

                int ncand;
                if (isC) pi0 = m.sd[2];
                const double inv_pi0 = 1.0 / pi0;
                if (isC) {
                    // 0.5 [pi(a|H') pi(b|H'+a) + pi(b|H') pi(a|H'+b)]   (piSMCJointly :120-132)
                    const double pj = 0.5 * mA1 * mB2 + 0.5 * mB1 * mA2;
#pragma unroll 1
                    for (int k = sub; k < Ltot; k += W) m.pv[k] = dtheta_e * (double)Cij * m.pv[k] * inv_pi0;
                    if (sub == 0) {
                        m.pv[Ltot] = (Cij > 1) ? (double)(Cij * (Cij - 1)) * inv_pi0 : 0.0;
                        m.pv[Ltot + 1] = (Ai > 0) ? (double)(Ai * Cij) / mA1 : 0.0;   // pi(xa | H'+x-xa) == pi(xa | H')
                        m.pv[Ltot + 2] = (Bj > 0) ? (double)(Bj * Cij) / mB1 : 0.0;
                        m.pv[Ltot + 3] = drho_e * (double)Cij * pj * inv_pi0;
                    }
                    ncand = Ltot + 4;
                    npi += (unsigned int)(Ltot + 1) + 4u + (Ai > 0 ? 1u : 0u) + (Bj > 0 ? 1u : 0u);
                } else {
#pragma unroll 1
                    for (int k = sub; k < Lx; k += W) m.pv[k] *= inv_pi0;
                    if (sub == 0) m.pv[Lx + np] = (own > 1 || ownC > 0) ? (double)(own * (own - 1 + ownC)) * inv_pi0 : 0.0;
                    ncand = Lx + np + 1;
                    npi += (unsigned int)(1 + Lx + 5 * np);
                }
                G::sync();

                // ---- sample the move (sampleFromVector :568-583) ----
                const int idx = pick<W>(ncand, rng.next(), false);
                // proposal density of the event: waiting time, lineage choice, move (:696-698, :455)
                const double qfac = (rate * hap_num * m.sd[0]) / (hap_den * m.sd[1]);

                // ---- apply the move, event constant (:242-294, :342-371, :418-448) ----
                double evc = 1.0; int kind = 0;          // kind: 0 coalescence, 1 mutation, 2 recombination
                int ev_kind = 1; unsigned long long ev_aux = 0ull; int ev_site = 0xff;
                int og0 = 0, og1 = 0, od0 = 0, od1 = 0, nops = 0; unsigned long long ow0 = 0ull, ow1 = 0ull;
                if (idx < Lx) {                                                        // mutation of site idx
                    ow0 = xw ^ (1ull << (offb + idx)); og0 = xg; od0 = +1; nops = 1;
                    kind = 1; ev_kind = 0; ev_aux = ow0; ev_site = idx;
                } else if (isC) {
                    if (idx == Ltot) { evc = (double)((unsigned int)Cc * (Cij - 1)); }
                    else if (idx == Ltot + 1) { og0 = GC; ow0 = xw; od0 = +1; og1 = GA; ow1 = xa; od1 = -1; nops = 2;
                                                evc = (double)((unsigned int)Ac * Cij); ev_kind = 2; ev_aux = xa; }
                    else if (idx == Ltot + 2) { og0 = GC; ow0 = xw; od0 = +1; og1 = GB; ow1 = xb; od1 = -1; nops = 2;
                                                evc = (double)((unsigned int)Bc * Cij); ev_kind = 3; ev_aux = xb; }
                    else { og0 = GA; ow0 = xa; od0 = +1; og1 = GB; ow1 = xb; od1 = +1; nops = 2; kind = 2; ev_kind = 4; }
                } else if (idx < Lx + np) {                                            // A^B_k coalescence
                    const unsigned long long pw = m.words[m.plist[idx - Lx]];
                    og0 = pg; ow0 = pw; od0 = -1; og1 = GC; ow1 = xw | pw; od1 = +1; nops = 2;
                    ev_kind = 5; ev_aux = pw;
                } else {
                    evc = isA ? (double)((unsigned int)Ac * (Ai + Ci - 1)) : (double)((unsigned int)Bc * (Bj + Cj - 1));
                }
                double after0 = 1.0, after1 = 1.0;
                if (nops > 0) { const int slot = adjust<W>(s, og0, ow0, od0); if (slot >= 0) after0 = (double)(m.cg[slot] & kCntMask); }
                if (nops > 1) { const int slot = adjust<W>(s, og1, ow1, od1); if (slot >= 0) after1 = (double)(m.cg[slot] & kCntMask); }
                if (ev_kind == 0) evc = after0;
                else if (ev_kind == 4) evc = (double)Cc * after0 * after1 / (((double)Ac + 1.0) * ((double)Bc + 1.0));
                else if (ev_kind == 5) evc = (double)((unsigned int)Ac * (unsigned int)Bc) * after1 / ((double)Cc + 1.0);
                const int ns_event = s.nslots;      // slots in use before the empty ones are dropped: the event's peak
                s.nslots = commit<W>(m, s.nslots);
                if (tracing && sub == 0 && nev < sA.max_events) {
                    coalgpu_event& r = sA.events[(unsigned long long)h * sA.max_events + nev];
                    r.chosen_word = xw; r.aux_word = ev_aux; r.chosen_gamma = (uint8_t)xg; r.kind = (uint8_t)ev_kind;
                    r.epoch = (uint8_t)e; r.site = (uint8_t)ev_site; r.n_before = (uint32_t)n;
                }
                nev++;

                // ---- target density bookkeeping (:457-546) ----
                if (sub == 0) {
                    EpochStats& st = hs->st;
                    hs->tsum = tsum + wait;
                    if (ns_event > hs->peak) hs->peak = ns_event;
                    hs->qsum += -rate * wait;
                    hs->qprod.mul(qfac);
                    hs->evprod.mul(evc);
                    st.s_nn += nn * wait; st.s_a += aa * wait; st.s_c += (double)Cc * wait;
                    if (kind == 1) st.n_mut++; else if (kind == 2) st.n_rec++;
                    if (post) { st.has_post = 1; st.post_nn = nn; st.post_a = aa; st.post_c = (double)Cc; }
                    else st.n_rate++;
                }
                post = false;
                G::sync();
                if (nev >= (1u << 18)) s.overflow = true;   // runaway guard: histories have O(10^2-10^3) events
                end_now = s.overflow;
            }
        }

        if (end_now) {
            // ---- final weight (ProposalEngine.cpp:817-858) ----
            const double ln2 = 0.69314718055994530942;
            G::sync();
            const double q = hs->qsum + logprod_value(hs->qprod);
            const double fin = hs->consts + sP.perm_const - q
                             - ((double)(LA + LB) * ln2 * (double)s.nC + (double)LA * ln2 * (double)s.nA + (double)LB * ln2 * (double)s.nB);
            if (wout) {
                G::sync();
                const double bad = __longlong_as_double(0x7ff8000000000000ll);
                for (int g = sub; g < NG; g += W) wout[g] = s.overflow ? bad : wout[g] + fin;
            }
            if (sub == 0) {
                if (sA.tmrca) sA.tmrca[h] = hs->tsum;
                if (sA.nevents) sA.nevents[h] = nev;
                if (tracing && sA.event_counts) sA.event_counts[h] = nev;
                if (!s.overflow) atomicMax(sA.counters + 8, (unsigned long long)hs->peak);
                if (s.overflow) {
                    if (sA.defer_list) sA.defer_list[atomicAdd(sA.counters + 6, 1ull)] = h;   // drawn again with the full capacity
                    else atomicAdd(sA.counters + 3, 1ull);
                }
            }
            if (!(s.overflow && sA.defer_list)) { ev_total += nev; pi_total += npi; cq_total += ncq; }
            have = false;
        }
    }
    const unsigned long long nh_total = (unsigned long long)G::sum_u(nh_lane);   // < 2^32 per group and launch
    if (sub == 0) {
        atomicAdd(sA.counters + 1, ev_total); atomicAdd(sA.counters + 2, pi_total);
        atomicAdd(sA.counters + 4, cq_total); atomicAdd(sA.counters + 5, nh_total);
    }
}

}  // namespace

// K1: persistent groups of W lanes pull history indices from a global counter until the batch is drawn.
#ifndef COAL_MIN_CTAS
#define COAL_MIN_CTAS 3
#endif
#ifndef COAL_CTA_THREADS
#define COAL_CTA_THREADS 128
#endif
constexpr int kCtaThreads = COAL_CTA_THREADS;   // threads per CTA of the history kernel (W = 16, 32)
// W = 8 packs four histories into a warp; its CTAs are smaller so that shared memory, not the CTA granularity,
// decides how many histories fit an SM.
#ifndef COAL_CTA_THREADS_W8
#define COAL_CTA_THREADS_W8 64
#endif
#ifndef COAL_MIN_CTAS_W8
#define COAL_MIN_CTAS_W8 5
#endif
__host__ __device__ constexpr int cta_threads(int W) { return W == 8 ? COAL_CTA_THREADS_W8 : COAL_CTA_THREADS; }
template <int W, bool N32>
__global__ void __launch_bounds__(W == 8 ? COAL_CTA_THREADS_W8 : COAL_CTA_THREADS, W == 8 ? COAL_MIN_CTAS_W8 : COAL_MIN_CTAS)
history_kernel(const DevProblem P, const SampleArgs A) {
    if (threadIdx.x == 0) { sP = P; sA = A; }
    __syncthreads();
    run_group<W, N32>();
}

// host-side selection of the instantiation: W lanes per history, narrow words when LA + LB <= 32
typedef void (*history_kernel_fn)(const DevProblem, const SampleArgs);
history_kernel_fn history_kernel_select(int W, bool narrow) {
    if (narrow) return W == 8 ? history_kernel<8, true> : W == 16 ? history_kernel<16, true> : history_kernel<32, true>;
    return W == 8 ? history_kernel<8, false> : W == 16 ? history_kernel<16, false> : history_kernel<32, false>;
}

#include "coal_lockstep.cuh"

// K2: batched pi_SMC from class vectors (parity seam for ConditionalSamplingHMM::piSMC): one query per thread, through
// the same table rows and the same arithmetic (coal_math.h) as the history kernel. The class list of a thread sits in
// its own shared-memory column.
__global__ void __launch_bounds__(128) pismc_classes_kernel(DevProblem P, long long nq, const unsigned char* __restrict__ gamma,
                                                            const int* __restrict__ n_all, const int* __restrict__ tidx,
                                                            const long long* __restrict__ offs, const int* __restrict__ cnt,
                                                            const int* __restrict__ dA, const int* __restrict__ dB,
                                                            double* __restrict__ out) {
    unsigned int* col = reinterpret_cast<unsigned int*>(smem_dyn) + threadIdx.x;
    const int cs = blockDim.x;
    for (long long qi = (long long)blockIdx.x * blockDim.x + threadIdx.x; qi < nq; qi += (long long)gridDim.x * blockDim.x) {
        const long long lo = offs[qi], hi = offs[qi + 1];
        const int nH = (int)(hi - lo);
        const int g = gamma[qi];
        const unsigned int n = (unsigned int)n_all[qi];
        const double* row = P.rows + ((size_t)P.set_of_epoch[tidx[qi]] * P.n_max + (n - 1)) * P.row_doubles;
        ClassTotals tot; tot.n = 0; tot.ncompA = 0; tot.ncompB = 0;
        double num = 0.0; unsigned int ncomp = 0;
        for (int k = 0; k < nH; k++) {
            const int c = cnt[lo + k], a = dA[lo + k], b = dB[lo + k];
            col[k * cs] = pack_class((unsigned int)c, a, b);
            tot.n += (unsigned int)c;
            if (a >= 0) tot.ncompA += (unsigned int)c;
            if (b >= 0) tot.ncompB += (unsigned int)c;
            if (g == GA && a >= 0) { num += (double)c * __ldg(row + P.off_wEA + a); ncomp += (unsigned int)c; }
            if (g == GB && b >= 0) { num += (double)c * __ldg(row + P.off_wEB + b); ncomp += (unsigned int)c; }
        }
        double pi;
        if (g == GC) {
            auto ldmy = [row](int k) { const double2 v = __ldg(reinterpret_cast<const double2*>(row + 2 * k)); MYPair e; e.m = v.x; e.y = v.y; return e; };
            pi = nH > 0 ? pismc_bilinear(ldmy, P.SB, P.LA, P.LB, col, cs, nH, tot, tot.ncompA ? 1.0 / (double)tot.ncompA : 0.0,
                                         tot.ncompB ? 1.0 / (double)tot.ncompB : 0.0, 1.0 / (double)tot.n, P.pow2negLA, P.pow2negLB)
                        : P.pow2negL;
        } else {
            pi = marginal_finish(num, ncomp, nH > 0 ? n : 0u, g == GA ? P.pow2negLA : P.pow2negLB, P.wsum);
        }
        out[qi] = pi;
    }
}

// K3: per-block partial log-sum-exp of the (3+T) weighted statistics of SamplerOutput::addSample (SamplerOutput.cpp:76-93)
// in ONE pass over the weights: with m the running maximum of the log-weights w of a grid point and e_h = exp(w_h - m),
//   log sum w = m + log sum e,   log sum w^2 = 2m + log sum e^2,   log sum w t = m + log sum t e,   log sum w n_l = m + log sum n_l e,
// so every history costs one exp per grid point for all statistics (the reference inserts 3+T values per history and
// grid point into multisets and log-adds them in sorted order). Thread (r, c) of a block owns grid column c of its
// column tile and every kLseRows-th history; rows are combined with warp shuffles (columns narrower than a warp) and
// through shared memory. Weights of -inf (a target value of 0) count as 0; NaN weights are the histories that overflowed
// the type store (counted by K1) or a numerical failure: they are skipped and counted in counters[9].
// grid = (blocks over histories, column tiles, ceil(T / kLseT)); block partials [stat][g][block][2].
constexpr int kLseT = 8;          // lineage statistics per pass (blockIdx.z selects l0 = z * kLseT; z > 0 skips the first three)
struct LseAcc {
    double m, s[3 + kLseT];
    __device__ __forceinline__ void clear() { m = -INFINITY; for (int i = 0; i < 3 + kLseT; i++) s[i] = 0.0; }
    // this <- this (+) o
    __device__ __forceinline__ void merge(const LseAcc& o) {
        const double mm = fmax(m, o.m);
        if (mm == -INFINITY) return;
        const double f1 = exp(m - mm), f2 = exp(o.m - mm);          // exp(-inf) = 0: an empty side vanishes
        s[1] = s[1] * (f1 * f1) + o.s[1] * (f2 * f2);
        s[0] = s[0] * f1 + o.s[0] * f2; s[2] = s[2] * f1 + o.s[2] * f2;
#pragma unroll
        for (int i = 3; i < 3 + kLseT; i++) s[i] = s[i] * f1 + o.s[i] * f2;
        m = mm;
    }
};
__global__ void __launch_bounds__(256) lse_partial_kernel(unsigned long long n, int G, int T, int cols /* power of two <= 32 */,
                                                          const double* __restrict__ weights, const double* __restrict__ tmrca,
                                                          const double* __restrict__ lineages, double* __restrict__ block_partials,
                                                          unsigned long long* __restrict__ counters) {
    const int c = threadIdx.x & (cols - 1), r = threadIdx.x / cols, rows = blockDim.x / cols;
    const int g = blockIdx.y * cols + c;
    const int l0 = blockIdx.z * kLseT, nl = min(kLseT, T - l0);
    LseAcc a; a.clear();
    unsigned int n_nan = 0;
    if (g < G) {
        for (unsigned long long h = (unsigned long long)blockIdx.x * rows + r; h < n; h += (unsigned long long)gridDim.x * rows) {
            const double w = weights[h * (unsigned long long)G + g];
            if (w != w) { n_nan++; continue; }
            if (w == -INFINITY) continue;
            double e = 1.0;
            if (w > a.m) {
                const double f = exp(a.m - w);                       // a.m = -inf: 0
                a.s[0] *= f; a.s[1] *= f * f; a.s[2] *= f;
#pragma unroll
                for (int i = 3; i < 3 + kLseT; i++) a.s[i] *= f;
                a.m = w;
            } else {
                e = exp(w - a.m);
            }
            a.s[0] += e; a.s[1] += e * e; a.s[2] += tmrca[h] * e;
            const double* ln = lineages + h * (unsigned long long)T + l0;
#pragma unroll
            for (int i = 0; i < kLseT; i++) if (i < nl) a.s[3 + i] += ln[i] * e;
        }
    }
    // rows of one warp that share a column: xor shuffles over the row bits of the lane index
    for (int o = cols; o < 32; o <<= 1) {
        LseAcc b;
        b.m = __shfl_xor_sync(kFull, a.m, o);
#pragma unroll
        for (int i = 0; i < 3 + kLseT; i++) b.s[i] = __shfl_xor_sync(kFull, a.s[i], o);
        a.merge(b);
    }
    n_nan = warp_sum_u(n_nan);
    // one accumulator per (warp, column) through shared memory, combined by the threads of warp 0
    __shared__ double sh[8][32][4 + kLseT];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < cols) {
        sh[warp][lane][0] = a.m;
#pragma unroll
        for (int i = 0; i < 3 + kLseT; i++) sh[warp][lane][1 + i] = a.s[i];
    }
    __syncthreads();
    if (warp == 0 && lane < cols) {
        for (int w2 = 1; w2 < (int)(blockDim.x >> 5); w2++) {
            LseAcc b; b.m = sh[w2][lane][0];
#pragma unroll
            for (int i = 0; i < 3 + kLseT; i++) b.s[i] = sh[w2][lane][1 + i];
            a.merge(b);
        }
        if (g < G) {
            auto put = [&](int stat, double mx, double sm) {
                double* o = block_partials + (((size_t)stat * G + g) * gridDim.x + blockIdx.x) * 2;
                o[0] = mx; o[1] = sm;
            };
            if (blockIdx.z == 0) { put(0, a.m, a.s[0]); put(1, 2.0 * a.m, a.s[1]); put(2, a.m, a.s[2]); }
            for (int i = 0; i < nl; i++) put(3 + l0 + i, a.m, a.s[3 + i]);
        }
    }
    if (lane == 0 && n_nan && blockIdx.z == 0) atomicAdd(counters + 9, (unsigned long long)n_nan);
}

// K3b: combine block partials -> partials[stat][g][2]
__global__ void lse_combine_kernel(int nblocks, int total, const double* __restrict__ block_partials, double* __restrict__ partials) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double* p = block_partials + (size_t)i * nblocks * 2;
    double mx = -INFINITY;
    for (int b = 0; b < nblocks; b++) if (p[2 * b + 1] > 0.0) mx = fmax(mx, p[2 * b]);
    double sm = 0.0;
    if (mx > -INFINITY) for (int b = 0; b < nblocks; b++) if (p[2 * b + 1] > 0.0) sm += p[2 * b + 1] * exp(p[2 * b] - mx);
    partials[2 * i] = mx; partials[2 * i + 1] = sm;
}

// FP64 FMA throughput probe (the roofline denominator of an fp64-ALU-bound path; MEASURED_PEAKS.json has none):
// 8 independent dependent-FMA chains per thread.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

size_t history_group_bytes(int kmax, int pvmax, int Ltot, int W, int stage_doubles) { return group_mem_bytes(kmax, pvmax, Ltot, W, stage_doubles); }

// K4: emission tables of the stable mode on the device (the host version and the derivation are in
// coal_tables.cpp: E[i][d] = 1/(w_i 2^L) * integral over interval i of e^-t (1+e^-at)^(L-d) (1-e^-at)^d dt, a = 2 theta/n,
// 32-point Gauss-Legendre on the finite intervals, 32-point Gauss-Laguerre on the last). One thread per
// (table set, lineage count n, locus, distance d): 16 intervals x 32 nodes, then the weighted sums wE[d].
// On the host this costs seconds per data set (4 transcendentals per node); here it is a millisecond.
__constant__ double c_GLx[32] = {-0.9972638618494816, -0.9856115115452684, -0.9647622555875064, -0.9349060759377397, -0.8963211557660521, -0.84936761373257, -0.7944837959679424, -0.7321821187402897, -0.6630442669302152, -0.5877157572407623, -0.5068999089322294, -0.42135127613063533, -0.33186860228212767, -0.23928736225213706, -0.1444719615827965, -0.048307665687738324, 0.048307665687738324, 0.1444719615827965, 0.23928736225213706, 0.33186860228212767, 0.42135127613063533, 0.5068999089322294, 0.5877157572407623, 0.6630442669302152, 0.7321821187402897, 0.7944837959679424, 0.84936761373257, 0.8963211557660521, 0.9349060759377397, 0.9647622555875064, 0.9856115115452684, 0.9972638618494816};
__constant__ double c_GLw[32] = {0.007018610009470506, 0.016274394730905743, 0.025392065309262024, 0.034273862913021765, 0.042835898022226836, 0.05099805926237609, 0.058684093478535565, 0.06582222277636168, 0.07234579410884834, 0.07819389578707023, 0.08331192422694671, 0.08765209300440378, 0.09117387869576378, 0.09384439908080451, 0.09563872007927471, 0.09654008851472766, 0.09654008851472766, 0.09563872007927471, 0.09384439908080451, 0.09117387869576378, 0.08765209300440378, 0.08331192422694671, 0.07819389578707023, 0.07234579410884834, 0.06582222277636168, 0.058684093478535565, 0.05099805926237609, 0.042835898022226836, 0.034273862913021765, 0.025392065309262024, 0.016274394730905743, 0.007018610009470506};
__constant__ double c_Lagx[32] = {0.04448936583326739, 0.23452610951961825, 0.5768846293018863, 1.072448753817818, 1.7224087764446456, 2.5283367064257942, 3.492213273021994, 4.616456769749767, 5.903958504174244, 7.358126733186241, 8.982940924212597, 10.78301863253997, 12.763697986742725, 14.931139755522558, 17.292454336715316, 19.855860940336054, 22.630889013196775, 25.628636022459247, 28.862101816323474, 32.346629153964734, 36.10049480575197, 40.14571977153944, 44.50920799575494, 49.22439498730864, 54.33372133339691, 59.89250916213402, 65.97537728793505, 72.68762809066271, 80.18744697791352, 88.7353404178924, 98.82954286828398, 111.7513980979377};
__constant__ double c_Lagw[32] = {0.10921834195242515, 0.2104431079387924, 0.23521322966984298, 0.19590333597287354, 0.1299837862860684, 0.07057862386571556, 0.03176091250917397, 0.011918214834838204, 0.003738816294611417, 0.0009808033066149246, 0.0002148649188013578, 3.920341967987801e-05, 5.934541612868448e-06, 7.41640457866734e-07, 7.604567879120552e-08, 6.350602226625618e-09, 4.281382971040738e-10, 2.305899491891237e-11, 9.799379288726772e-13, 3.2378016577291567e-14, 8.171823443420549e-16, 1.542133833393811e-17, 2.119792290163547e-19, 2.0544296737880036e-21, 1.3469825866373617e-23, 5.661294130397462e-26, 1.4185605454629684e-28, 1.9133754944539943e-31, 1.1922487600982012e-34, 2.6715112192396625e-38, 1.3386169421064243e-42, 4.5105361938987784e-48};
struct EmissionArgs {
    double* rows;               // [n_sets][n_max][row_doubles], reference layout (coal_tables.h): the kernel fills EA | EB | wEA | wEB
    const double* thetas;       // [n_sets]
    int n_sets, n_max, row_doubles, LA, LB, Q;
    int off_EA, off_EB, off_wEA, off_wEB;
    double t[kQd + 1], w[kQd];
};
__global__ void __launch_bounds__(128) emission_stable_kernel(const EmissionArgs A) {
    const int per_row = A.LA + 1 + A.LB + 1;
    const long long total = (long long)A.n_sets * A.n_max * per_row;
    for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(id % per_row);
        const long long rn = id / per_row;
        const int n = (int)(rn % A.n_max) + 1, set = (int)(rn / A.n_max);
        const bool locA = k <= A.LA;
        const int L = locA ? A.LA : A.LB, d = locA ? k : k - A.LA - 1;
        double* row = A.rows + (size_t)rn * A.row_doubles;
        double* E = row + (locA ? A.off_EA : A.off_EB) + (size_t)d * kEStride;
        const double a = 2 * A.thetas[set] / (double)n;
        const double two_L = exp2((double)L);
        double wE = 0.0;
        for (int i = 0; i < kQd; i++) {
            double e = 0.0;
            if (i < A.Q) {
                double v = 0.0;
                const double lo = A.t[i];
                if (i + 1 < A.Q) {
                    const double hi = A.t[i + 1], h = 0.5 * (hi - lo), c = 0.5 * (hi + lo);
                    for (int q = 0; q < 32; q++) {
                        const double tt = c + h * c_GLx[q], x = exp(-a * tt);
                        v += c_GLw[q] * exp(-tt + (double)(L - d) * log1p(x) + (d > 0 ? (double)d * log1p(-x) : 0.0));
                    }
                    v *= h;
                } else {
                    for (int q = 0; q < 32; q++) {
                        const double x = exp(-a * (lo + c_Lagx[q]));
                        v += c_Lagw[q] * exp((double)(L - d) * log1p(x) + (d > 0 ? (double)d * log1p(-x) : 0.0));
                    }
                    v *= exp(-lo);
                }
                e = fmax(v / (A.w[i] * two_L), 2.220446049250313e-16);      // the reference's clamp (ConditionalSamplingHMM.cpp:221)
            }
            E[i] = e;
            wE += A.w[i] * e;
        }
        row[(locA ? A.off_wEA : A.off_wEB) + d] = wE;
    }
}

// K5: device rows (bilinear tables M, Y + one-locus weight vectors, coal_tables.h) from reference-layout rows that live
// on the device -- the second half of the device table builder: the stable-mode emission tables never visit the host.
// One thread per (row, dA) for the MY part, the first LA + LB + 2 threads of a row's range copy the weight vectors.
struct BilinearArgs {
    const double* full_rows;    // [n_rows][row_doubles]
    double* dev_rows;           // [n_rows][dev_row_doubles]
    long long n_rows;
    int row_doubles, dev_row_doubles, LA, LB;
    int off_yw, off_zlow, off_g, off_zdiag, off_EA, off_EB, off_wEA, off_wEB, dev_sb, dev_off_wEA, dev_off_wEB;
    double w[kQd];
};
__global__ void __launch_bounds__(128) bilinear_rows_kernel(const BilinearArgs A) {
    const int per_row = A.LA + 2;
    const long long total = A.n_rows * per_row;
    for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
        const int dA = (int)(id % per_row);
        const long long rn = id / per_row;
        const double* row = A.full_rows + (size_t)rn * A.row_doubles;
        double* drow = A.dev_rows + (size_t)rn * A.dev_row_doubles;
        bilinear_row(row, A.w, A.off_yw, A.off_zlow, A.off_g, A.off_zdiag, A.off_EA, A.off_EB, A.LA, A.LB, dA, drow + (size_t)2 * dA * A.dev_sb);
        if (dA <= A.LA) drow[A.dev_off_wEA + dA] = row[A.off_wEA + dA];
        for (int d = dA; d <= A.LB; d += per_row) drow[A.dev_off_wEB + d] = row[A.off_wEB + d];
    }
}

// Multi-GPU combine (coal_multi.cu), between the two NCCL allreduces: sums of one device to the common maximum: sm[i] *= exp(mx[i] - gmx[i]) (0 for a device without samples)
__global__ void lse_rescale_kernel(int n, const double* __restrict__ mx, const double* __restrict__ gmx, double* __restrict__ sm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double m = mx[i], g = gmx[i];
    sm[i] = (m == -INFINITY || g == -INFINITY) ? 0.0 : sm[i] * exp(m - g);
}

}  // namespace coalgpu
