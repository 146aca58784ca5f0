// K1, lockstep form: a CTA of NT threads advances NH coalescent histories (NH <= 32 <= NT) in lockstep, one "step" (an injection,
// a collection boundary, or one backward event of ProposalEngine.cpp:134-566) per history and iteration. The histories of the
// CTA are dealt to its warps (HPW = NH / warps each); within a warp a history owns LPH = 32 / HPW consecutive lanes, the first
// of which is its SERIAL THREAD. A step is, per warp and without CTA barriers in between (only __syncwarp / shuffles: a warp
// touches the state of its own histories only),
//
//   S1  (serial threads) finish the event of the previous step: assemble the candidate probabilities, draw the move
//       (sampleFromVector, ProposalEngine.cpp:568-583), book the target-density statistics;
//   W1  (all lanes, LPH per history) the move's type changes on the store (TypeContainer::addType / removeType) as ONE
//       cooperative walk, then (serial) append what was not stored, drop the slots the event emptied, event constant;
//   S2  (serial) fetch the next history when the slot is free;
//   W0  (all lanes, only when a history injects) the sampled types of a collection enter the store, one walk per type;
//   W2  (all lanes) lineage-chooser weights of every stored type (TypeContainer::sampleType, TypeContainer.cpp:641-745);
//   S3  (serial) sum the weights in slot order, draw the lineage and the waiting time, handle a collection boundary or the end
//       of the history, start the event descriptor;
//   W3  (all lanes) remove the chosen lineage: cross counts, C-type counts, partner list in slot order (ballot rank);
//   S4/S5 (serial) publish the descriptor / write out a history that ended;
//
// and then, across the CTA and separated by CTA barriers,
//
//   G   (all threads, only in steps where some history closed an epoch or ended) the per-grid-point target density of the
//       epoch from its sufficient statistics (coal_math.h: epoch_log_density), lanes over grid points;
//   Q1  (all threads, ONE LANE PER (history, one-locus query)) the flattened list of all one-locus conditional sampling
//       probabilities of the CTA's events (ConditionalSamplingHMM.cpp:332-370 as a dot product);
//   Q2  (all threads, ONE LANE PER (history, two-locus query)) class list of the query (TypeContainer::getDifferences with
//       the dA+dB collapse, TypeContainer.cpp:802-886) + its pi_SMC as a bilinear form (coal_math.h).
//
// Why: the group form of the kernel (coal_kernels.cu: a group of 8-32 lanes per history, every group its own control
// flow) keeps 40-64 independent instruction streams per SM wandering through ~70 KB of code -- ncu: 3.9 stalled-for-
// instruction-fetch warps per issue at 16 warps/SM, 15 of 32 lanes active (profiles/r2_v7_*). Here the serial sections are
// short scalar code shared by the histories of a warp, every walk over a type store uses all lanes of the warp, and the
// query phases run one query per lane over the queries of ALL histories of the CTA, so lanes are full whatever the number
// of candidate moves of one event is. The type store is laid out slot-major / history-minor ([slot][NH]). Results are those
// of the group form: same canonical slot order, same draws, partial sums of the categorical draws in plain index order
// (exactly the oracle's order). History of the form: profiles/r2_kernel_iterations.md.
#pragma once

namespace {

#ifndef COAL_LS_DENSE_MAX
#define COAL_LS_DENSE_MAX 64
#endif
constexpr int kLsDenseMax = COAL_LS_DENSE_MAX;   // dense bilinear form up to this many (dA, dB) cells, class-list form beyond
#ifndef COAL_LS_UNROLL_S
#define COAL_LS_UNROLL_S 2
#endif
#ifndef COAL_LS_UNROLL_Q
#define COAL_LS_UNROLL_Q 2
#endif
constexpr int kLsUnrollS = COAL_LS_UNROLL_S, kLsUnrollQ = COAL_LS_UNROLL_Q;   // unroll factors of the slot loops (serial / query phases)
constexpr int kLsFlagAccum = 1;    // snapshot: add the epoch's target density to the history's weights
constexpr int kLsFlagFinal = 2;    //           ... and the final constant: the history ends
constexpr int kLsFlagAbort = 4;    //           the history outgrew the type store: weights = NaN
constexpr int kLsFlagFirst = 8;    //           first write of this history's weights (no read of the old value)

struct LsDesc {                    // event descriptor of one history and step (written in S, read in Q1 / Q2 / next S)
    unsigned long long xw;         // chosen haplotype
    const double* crow;            // table row of the two-locus queries (shared-memory staging buffer or the table in HBM)
    const double* row_m1;          // rows of the one-locus queries, biased so that (+ off_wEA / off_wEB) lands on the weight vectors
    const double* row_alt;
    double dtheta_own;             // A/B chosen: driving theta * multiplicity of the chosen type
    double mA1, mA2, mB1, mB2;     // results of one-locus queries 0 and 1 on the two rows (A/B chosen: mA1 = pi0)
    double pi0c;                   // C chosen: pi(x | H')
    double rate, wait, hap_num, hap_den, nn, aa, dtheta_e, drho_e;
    int valid, xg, ns, n, np, nm, nq, Lx, offb, pg, sx;
    int Ac, Bc, Cc;
    unsigned int own, ownC, Cij, Ci, Cj, Ai, Bj, cn, nc_alt, tma_parity;
};

struct LsSnap {                    // what the G phase needs of a history that closed an epoch / ended in this step
    EpochStats st;
    double fin;
    unsigned long long hidx;
    int e, flags;
};

// Shared memory of a CTA. Capacity tier (gstore): the arrays that grow with the type store -- pv, aux, words, cg, xc, plist,
// 34 bytes per slot and history -- live in the CTA's slice of a global-memory arena instead (gtotal bytes per CTA,
// offsets relative to the slice): histories with thousands of lineages (TypeContainer has no size limit,
// TypeContainer.cpp:409-522) then run at L2 speed instead of not at all. Everything per event stays in shared memory.
struct LsLayout { unsigned int words, cg, xc, plist, pv, aux, rowbuf, wbuf, mbar, scratch, desc, snap, pre1, pre2, rbuf, total; size_t gtotal; };

__host__ __device__ inline LsLayout ls_layout(int NH, int NT, int kmax, int pvmax, int Ltot, int stage_doubles, bool narrow, bool gstore = false) {
    LsLayout L;
    unsigned int o = 0, og = 0;
    auto place = [&o](size_t bytes) { const unsigned int at = o; o += (unsigned int)((bytes + 15) & ~(size_t)15); return at; };
    auto place_store = [&o, &og, gstore](size_t bytes) { unsigned int& c = gstore ? og : o; const unsigned int at = c; c += (unsigned int)((bytes + 15) & ~(size_t)15); return at; };
    const int nsum = Ltot + 2;
    L.rowbuf = place((size_t)NH * stage_doubles * 8);
#ifdef COAL_TMA_ROWS
    L.wbuf = place((size_t)NH * 2 * (((Ltot + 2) + 1) & ~1) * 8);
    L.mbar = place((size_t)NH * 8);
#else
    L.wbuf = L.mbar = 0;                                  // wbuf: see below (aliases the query scratch)
#endif
    L.pv = place_store((size_t)NH * pvmax * 8);
    L.aux = place_store((size_t)NH * kmax * 8);
    L.words = place_store((size_t)NH * kmax * (narrow ? 4 : 8));
    L.cg = place_store((size_t)NH * kmax * 4);
    L.xc = place_store((size_t)NH * kmax * 4);
    L.plist = place_store((size_t)NH * kmax * 2);
    // per thread: class histogram / class list [nsum] + the distance vectors of the dense form [nsum] (Q2). Without TMA staging
    // the same bytes first hold the one-locus weight vectors of the step: the wEA | wEB tails of the two table rows an event's
    // one-locus queries read (2 * wlen doubles per history), copied here by the history's lanes in the removal pass and read by
    // Q1 with 32-bit shared-memory addressing; Q2 overwrites them after Q1's barrier.
    {
        size_t sb = (size_t)NT * 2 * nsum * 4;
#ifndef COAL_TMA_ROWS
        const size_t wb = (size_t)NH * 2 * (((Ltot + 2) + 1) & ~1) * 8;
        if (wb > sb) sb = wb;
#endif
        L.scratch = place(sb);
    }
#ifndef COAL_TMA_ROWS
    L.wbuf = L.scratch;
#endif
    L.desc = place((size_t)NH * sizeof(LsDesc));
    L.snap = place((size_t)NH * sizeof(LsSnap));
    L.pre1 = place((size_t)(NT / 32) * (NH + 1) * 4);     // one copy of the work-list prefixes per warp
    L.pre2 = place((size_t)(NT / 32) * (NH + 1) * 4);
    L.rbuf = place((size_t)NT * 16);                      // Philox blocks drawn ahead: NT / NH per history
    L.total = o;
    L.gtotal = ((size_t)og + 255) & ~(size_t)255;
    return L;
}

#ifdef COAL_DEBUG_BOUNDS
#define LS_ASSERT(cond) do { if (!(cond)) { printf("coalgpu lockstep: %s failed at line %d (block %d thread %d)\n", #cond, __LINE__, (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define LS_ASSERT(cond) do { } while (0)
#endif

template <int NH, int NT_, int MD>
struct LsCta {
    static constexpr bool N32 = (MD & 1) != 0;     // 32-bit haplotype words (LA + LB <= 32)
    static constexpr bool GS = (MD & 2) != 0;      // type store in global memory (capacity tier, see ls_layout)
    typedef WordOps<N32> WO;
    typedef typename WO::T WordT;
    static constexpr int NT = NT_;
    WordT* words; unsigned int* cg; unsigned int* xc; unsigned short* plist; double* pv; double* aux;
    double* rowbuf; double* wbuf; unsigned long long* mbar; unsigned int* scratch; LsDesc* desc; LsSnap* snap; int* pre1; int* pre2; uint4* rbuf;

    __device__ __forceinline__ void bind() {
        const LsLayout L = ls_layout(NH, NT, sP.kmax, sP.pvmax, sP.Ltot, sP.stage_doubles, N32, GS);
        unsigned char* b = smem_dyn;
        unsigned char* sb = GS ? sA.gstore + (size_t)blockIdx.x * L.gtotal : b;     // the CTA's slice of the launch's arena
        words = reinterpret_cast<WordT*>(sb + L.words); cg = reinterpret_cast<unsigned int*>(sb + L.cg); xc = reinterpret_cast<unsigned int*>(sb + L.xc);
        plist = reinterpret_cast<unsigned short*>(sb + L.plist); pv = reinterpret_cast<double*>(sb + L.pv); aux = reinterpret_cast<double*>(sb + L.aux);
        rowbuf = reinterpret_cast<double*>(b + L.rowbuf); wbuf = reinterpret_cast<double*>(b + L.wbuf); mbar = reinterpret_cast<unsigned long long*>(b + L.mbar);
        scratch = reinterpret_cast<unsigned int*>(b + L.scratch); desc = reinterpret_cast<LsDesc*>(b + L.desc); snap = reinterpret_cast<LsSnap*>(b + L.snap);
        pre1 = reinterpret_cast<int*>(b + L.pre1); pre2 = reinterpret_cast<int*>(b + L.pre2); rbuf = reinterpret_cast<uint4*>(b + L.rbuf);
    }
    // slot-major / history-minor: element j of history h
    static __device__ __forceinline__ int at(int j, int h) { return j * NH + h; }
};

// Per-history state of the serial phase: lives in the registers of the history's thread across steps.
struct LsHist {
    unsigned long long h;
    unsigned long long rk, rb0;      // Philox4x32-10 stream of the history: next draw k (block k >> 1, half k & 1); first block held in rbuf
    EpochStats st;
    LogProd qprod, evprod;
    double qsum, consts, tsum;
    int nslots, nA, nB, nC, inj_lo, inj_hi, e, tset, peak;
    unsigned int nev, npi, ncq, tma_phase;
    bool have, done, injecting, post, tracing, overflow, wfirst, pending;
};

// TypeContainer::addType / removeType (TypeContainer.cpp:409-522) on the slot array of history h, serial form of
// adjust_core: the multiplicity of type (g, w) changes by delta, the cross counts of every related slot follow; a type
// that is not stored yet is appended; a slot whose count reaches 0 stays until ls_commit. Returns the slot (| kAppended
// when a slot was appended) or -1 when the store is full. Out of line (three call sites), so it takes scalars only:
// a reference to the caller's register state would force that state into local memory.
template <int NH, int NT, int MD>
__device__ __noinline__ int ls_adjust_core(int h, int ns, int g, unsigned long long w64, int delta, bool active) {
    typedef LsCta<NH, NT, MD> C;
    typedef typename C::WO WO;
    C m; m.bind();
    // Every history of the warp calls this at the same program points (a history without the operation passes
    // active = false and idles through the call), so the histories that do have it walk their slots TOGETHER instead of
    // one sub-set of lanes after the other.
    if (!active) ns = 0;
    const typename WO::T w = WO::narrow(w64), maskA = WO::narrow(sP.maskA), maskB = WO::narrow(sP.maskB);
    const typename WO::T wa = w & maskA, wb = w & maskB;
    int found = -1;
    unsigned int newxc = 0;
    // The related slots as two match rules (class, mask, key) -> (cross-count change, contribution to a new slot's cross
    // counts), so that the loop body is the same instruction sequence whatever g is: the histories of a warp change types of
    // different classes at the same time, and a branch per class would run one class after the other.
    //   g = A: C slots with this A part (count of their A part += delta)        g = B: C slots with this B part (<< 16)
    //   g = C: the A slot of its A part and the B slot of its B part (C lineages carrying it += delta)
    const int t1 = (g == GC) ? GA : GC, t2 = (g == GC) ? GB : 3;
    const typename WO::T m1 = (g == GA) ? maskA : (g == GB) ? maskB : (typename WO::T)~(typename WO::T)0;
    const typename WO::T k1 = (g == GC) ? wa : w, k2 = wb;
    const unsigned int d1 = (g == GB) ? (unsigned int)(delta * 65536) : (unsigned int)delta, d2 = (unsigned int)delta;
#pragma unroll kLsUnrollS
    for (int j = 0; j < ns; j++) {
        const typename WO::T wj = m.words[C::at(j, h)];
        const unsigned int cgj = m.cg[C::at(j, h)];
        const int gj = (int)(cgj >> 30);
        const unsigned int cj = cgj & kCntMask;
        if (gj == g && wj == w) found = j;
        const bool h1 = (gj == t1) && ((wj & m1) == k1), h2 = (gj == t2) && (wj == k2);
        const unsigned int xadd = (h1 ? d1 : 0u) + (h2 ? d2 : 0u);
        if (xadd) m.xc[C::at(j, h)] += xadd;
        newxc += (h1 ? cj : 0u) + (h2 ? (cj << 16) : 0u);
    }
    if (!active) return -2;
    int ret = found;
    if (found < 0) {
        if (ns >= sP.kmax) return -1;
        found = ns;
        m.words[C::at(found, h)] = w; m.cg[C::at(found, h)] = (unsigned int)g << 30; m.xc[C::at(found, h)] = newxc;
        ret = found | kAppended;
    }
    m.cg[C::at(found, h)] += (unsigned int)delta;
    return ret;
}
template <int NH, int NT, int MD>
__device__ __forceinline__ int ls_adjust(const LsCta<NH, NT, MD>&, int h, LsHist& s, int g, unsigned long long w64, int delta, bool active = true) {
    int r = ls_adjust_core<NH, NT, MD>(h, s.nslots, g, w64, delta, active);
    if (r == -2) return -1;
    if (r < 0) { s.overflow = true; return -1; }
    if (r & kAppended) { s.nslots++; r &= ~kAppended; }
    if (g == GA) s.nA += delta; else if (g == GB) s.nB += delta; else s.nC += delta;
    return r;
}

// The two type changes of a backward move (ProposalEngine.cpp:242-294, 342-371, 418-448: at most two addType / removeType
// calls) in ONE cooperative pass over the stored types: exactly ls_adjust_core(op 0) followed by ls_adjust_core(op 1) -- op 1
// sees the count op 0 left in its slot and the slot op 0 appended. Called by ALL lanes of a warp at a convergent point: the
// LPH lanes of a history (its serial thread is the first of them) walk the store LPH slots at a time and combine what they
// found with shuffles; the serial thread then appends what is not stored yet. Arguments are those of the lane's history
// (already broadcast from its serial thread); op: bit 0 / 1 = op 0 / 1 has work, bits 2-3 / 4-5 their classes.
// Returns slot 0 | slot 1 << 32 (each as ls_adjust_core returns it; valid on the serial thread).
template <int NH, int NT, int MD, int LPH>
__device__ __noinline__ unsigned long long ls_adjust2_coop(int h, int sub, int ns, int op, unsigned long long w0_64, int delta0,
                                                            unsigned long long w1_64, int delta1) {
    typedef LsCta<NH, NT, MD> C;
    typedef typename C::WO WO;
    typedef typename WO::T W;
    C m; m.bind();
    const bool act0 = (op & 1) != 0, act1 = (op & 2) != 0;
    const int g0 = (op >> 2) & 3, g1 = (op >> 4) & 3;
    const W maskA = WO::narrow(sP.maskA), maskB = WO::narrow(sP.maskB), ones = (W)~(W)0;
    const W w0 = WO::narrow(w0_64), w1 = WO::narrow(w1_64);
    // match rules as in ls_adjust_core; an inactive op gets class 7, which no slot has
    const int G0 = act0 ? g0 : 7, G1 = act1 ? g1 : 7;
    const int t1a = !act0 ? 7 : (g0 == GC ? GA : GC), t2a = !act0 ? 7 : (g0 == GC ? GB : 7);
    const int t1b = !act1 ? 7 : (g1 == GC ? GA : GC), t2b = !act1 ? 7 : (g1 == GC ? GB : 7);
    const W m1a = (g0 == GA) ? maskA : (g0 == GB) ? maskB : ones, m1b = (g1 == GA) ? maskA : (g1 == GB) ? maskB : ones;
    const W k1a = (g0 == GC) ? (w0 & maskA) : w0, k2a = w0 & maskB, k1b = (g1 == GC) ? (w1 & maskA) : w1, k2b = w1 & maskB;
    const unsigned int d1a = (g0 == GB) ? (unsigned int)(delta0 * 65536) : (unsigned int)delta0, d2a = (unsigned int)delta0;
    const unsigned int d1b = (g1 == GB) ? (unsigned int)(delta1 * 65536) : (unsigned int)delta1, d2b = (unsigned int)delta1;
    int found0 = -1, found1 = -1;
    unsigned int newxc0 = 0, newxc1 = 0;
    const int nloop = (act0 || act1) ? ns : 0;
#pragma unroll 1
    for (int j = sub; j < nloop; j += LPH) {
        const W wj = m.words[C::at(j, h)];
        const unsigned int cgj = m.cg[C::at(j, h)];
        const int gj = (int)(cgj >> 30);
        const unsigned int cj = cgj & kCntMask;
        const bool f0 = (gj == G0) && (wj == w0);
        if (f0) found0 = j;
        const bool h1a = (gj == t1a) && ((wj & m1a) == k1a), h2a = (gj == t2a) && (wj == k2a);
        newxc0 += (h1a ? cj : 0u) + (h2a ? (cj << 16) : 0u);
        const unsigned int cj1 = cj + (f0 ? (unsigned int)delta0 : 0u);          // the count op 1 sees
        if ((gj == G1) && (wj == w1)) found1 = j;
        const bool h1b = (gj == t1b) && ((wj & m1b) == k1b), h2b = (gj == t2b) && (wj == k2b);
        newxc1 += (h1b ? cj1 : 0u) + (h2b ? (cj1 << 16) : 0u);
        const unsigned int xadd = (h1a ? d1a : 0u) + (h2a ? d2a : 0u) + (h1b ? d1b : 0u) + (h2b ? d2b : 0u);
        if (xadd) m.xc[C::at(j, h)] += xadd;
    }
    // combine over the lanes of the history (a type is stored at most once: max picks the slot)
#pragma unroll
    for (int o = 1; o < LPH; o <<= 1) {
        found0 = max(found0, __shfl_xor_sync(kFull, found0, o)); found1 = max(found1, __shfl_xor_sync(kFull, found1, o));
        newxc0 += __shfl_xor_sync(kFull, newxc0, o); newxc1 += __shfl_xor_sync(kFull, newxc1, o);
    }
    __syncwarp();                                           // the cross-count updates above are visible to the serial thread
    if (sub != 0) return 0ull;
    unsigned int r0 = 0xfffffffeu, r1 = 0xfffffffeu;       // -2: no op
    if (act0) {
        r0 = (unsigned int)found0;
        if (found0 < 0) {
            if (ns >= sP.kmax) return 0xffffffffffffffffull;          // store full
            found0 = ns++;
            // the appended slot as op 1 sees it: class g0, haplotype w0, count delta0
            const unsigned int c1 = (unsigned int)delta0;
            if (g0 == G1 && w0 == w1) found1 = found0;
            const bool h1b = (g0 == t1b) && ((w0 & m1b) == k1b), h2b = (g0 == t2b) && (w0 == k2b);
            newxc1 += (h1b ? c1 : 0u) + (h2b ? (c1 << 16) : 0u);
            m.words[C::at(found0, h)] = w0; m.cg[C::at(found0, h)] = (unsigned int)g0 << 30;
            m.xc[C::at(found0, h)] = newxc0 + (h1b ? d1b : 0u) + (h2b ? d2b : 0u);
            r0 = (unsigned int)(found0 | kAppended);
        }
        m.cg[C::at(found0, h)] += (unsigned int)delta0;
    }
    if (act1) {
        r1 = (unsigned int)found1;
        if (found1 < 0) {
            if (ns >= sP.kmax) return 0xffffffffffffffffull;
            found1 = ns++;
            m.words[C::at(found1, h)] = w1; m.cg[C::at(found1, h)] = (unsigned int)g1 << 30; m.xc[C::at(found1, h)] = newxc1;
            r1 = (unsigned int)(found1 | kAppended);
        }
        m.cg[C::at(found1, h)] += (unsigned int)delta1;
    }
    return (unsigned long long)r0 | ((unsigned long long)r1 << 32);
}

// Drop empty slots at the end of an event: highest empty slot first, overwritten by the last slot (the canonical order
// rule shared with the oracle's CANONICAL mode). A slot can only have run empty where the event took a
// lineage away: the chosen lineage's slot and the slots of the move's (at most two) type changes -- three candidates,
// visited in descending order, which is the order a downward scan over all slots would meet them in (everything above the
// candidate being visited is non-empty by then).
template <int NH, int NT, int MD>
__device__ __forceinline__ void ls_commit(const LsCta<NH, NT, MD>& m, int h, LsHist& s, int c0, int c1, int c2) {
    typedef LsCta<NH, NT, MD> C;
    // sort descending (3-element network); duplicates and -1 (no candidate) are skipped below
    int t;
    if (c0 < c1) { t = c0; c0 = c1; c1 = t; }
    if (c1 < c2) { t = c1; c1 = c2; c2 = t; }
    if (c0 < c1) { t = c0; c0 = c1; c1 = t; }
    int ns = s.nslots;
    int prev = -1;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int j = (i == 0) ? c0 : (i == 1) ? c1 : c2;
        if (j >= 0 && j != prev && (m.cg[C::at(j, h)] & kCntMask) == 0) {
            LS_ASSERT(j < ns);
            const int last = ns - 1;
            if (j != last) {
                m.words[C::at(j, h)] = m.words[C::at(last, h)]; m.cg[C::at(j, h)] = m.cg[C::at(last, h)]; m.xc[C::at(j, h)] = m.xc[C::at(last, h)];
            }
            ns = last;
        }
        prev = j;
    }
    s.nslots = ns;
}

// Categorical draw over pv[0..n) of history h given total = the sum of pv in index order: first k with cum > u*total
// (strict, TypeContainer.cpp:730) or with cum >= u*total (ProposalEngine.cpp:574). Returns k; pidx = pv[k].
template <int NH, int NT, int MD>
__device__ __forceinline__ int ls_scan(const LsCta<NH, NT, MD>& m, int h, int n, double target, bool strict, double& pidx) {
    typedef LsCta<NH, NT, MD> C;
    double cum = 0.0;
    int idx = n - 1;
#pragma unroll 1
    for (int k = 0; k < n; k++) {
        const double v = m.pv[C::at(k, h)];
        cum += v;
        if (strict ? (cum > target && v > 0.0) : !(cum < target)) { idx = k; break; }
    }
    pidx = m.pv[C::at(idx, h)];
    return idx;
}

// item i of a flattened work list -> (history, local index): pre[] = exclusive prefix of the per-history item counts
template <int NH>
__device__ __forceinline__ int ls_find(const int* pre, int i) {
    int lo = 0;
#pragma unroll
    for (int step = NH / 2; step > 0; step >>= 1) if (pre[lo + step] <= i) lo += step;
    return lo;
}

template <int NH, int NT, int MD>
__device__ __forceinline__ void run_lockstep() {
    typedef LsCta<NH, NT, MD> C;
    typedef typename C::WO WO;
    C m; m.bind();
    const int tid = threadIdx.x;
    // The histories of the CTA are dealt to its warps (HPW = NH / warps each), so the warps run their serial phases side by
    // side and fewer histories share -- and serialise -- one warp's branches. Within a warp a history owns LPH = 32 / HPW
    // consecutive lanes: the first is its serial thread, all of them walk its type store in the cooperative passes.
    constexpr int kWarps = NT / 32;
    constexpr int HPW = NH / kWarps;                                 // histories per warp
    constexpr int LPH = 32 / HPW;                                    // lanes per history in the cooperative passes
    static_assert(HPW * kWarps == NH && LPH * HPW == 32 && LPH >= 1, "lockstep shape");
    const int lane = tid & 31, warp = tid >> 5;
    const int sub = lane & (LPH - 1), leader = lane & ~(LPH - 1);
    const bool serial = (sub == 0);
    const int h = warp * HPW + lane / LPH;                           // history slot of this lane (its serial thread: lane `leader`)
    int* const pre1 = m.pre1 + warp * (NH + 1);
    int* const pre2 = m.pre2 + warp * (NH + 1);
    const int T = sP.T, LA = sP.LA, LB = sP.LB, Ltot = sP.Ltot, NG = sP.G;
    const int nsum = Ltot + 2;
    const int wlen = ((Ltot + 2) + 1) & ~1;            // doubles per staged weight tail (wEA | wEB), even
    constexpr bool kStaged =
#ifdef COAL_TMA_ROWS
        C::N32;
#else
        false;
#endif
    unsigned long long ev_total = 0, pi_total = 0, cq_total = 0;
    unsigned int nh_lane = 0;
    const unsigned long long n_todo = sA.list ? *sA.list_count : sA.n;

    LsHist s;
    s.h = 0; s.rk = 0; s.rb0 = ~0ull; epoch_stats_clear(s.st); s.qprod.clear(); s.evprod.clear(); s.qsum = s.consts = s.tsum = 0.0;
    s.nslots = s.nA = s.nB = s.nC = 0; s.inj_lo = s.inj_hi = 0; s.e = 0; s.tset = 0; s.peak = 0;
    s.nev = s.npi = s.ncq = 0; s.tma_phase = 0;
    s.have = false; s.done = !serial; s.injecting = false; s.post = true; s.tracing = false; s.overflow = false; s.wfirst = true; s.pending = false;
    // Draw k of a history = half k & 1 of Philox block k >> 1 (as PhiloxStream, coal_device.cuh). The blocks are computed AHEAD,
    // LPH at a time by the LPH lanes of the history at a convergent point of the step (before S3: the two draws of S3 and the
    // move draw of the next S1 are then in the buffer), instead of by the serial thread inside its divergent sections.
    auto next_u = [&]() {
        const uint4 blk = m.rbuf[h * LPH + (int)((s.rk >> 1) - s.rb0)];
        const double u = (s.rk & 1) ? u52(blk.z, blk.w) : u52(blk.x, blk.y);
        s.rk++;
        return u;
    };
    if (serial) {
        m.desc[h].valid = 0; m.desc[h].nm = 0; m.desc[h].nq = 0;
        m.snap[h].flags = 0;
#ifdef COAL_TMA_ROWS
        mbar_init(m.mbar + h);
#endif
    }
    __syncthreads();

#ifdef COAL_LS_TIMING
    long long tS = 0, tG = 0, tQ1 = 0, tQ2 = 0, tsteps = 0, t_mark = clock64();
#define LS_TICK(acc) do { const long long _t = clock64(); acc += _t - t_mark; t_mark = _t; } while (0)
#else
#define LS_TICK(acc) do { } while (0)
#endif
#pragma unroll 1
    for (;;) {
        int snap_flags = 0;
        // =================================== S: one thread per history ===================================
        // values that cross from the first serial section (S1) over the cooperative store pass into the second (S2)
        bool end_now = false;
        int mv_op = 0, mv_d0 = 0, mv_d1 = 0, mv_kind = 1;          // op bits as ls_adjust2_coop takes them; event kind
        unsigned long long mv_w0 = 0ull, mv_w1 = 0ull;
        double mv_evc = 1.0;
        if (serial && !s.done) {
            LsDesc& D = m.desc[h];
            // ---- finish the event of the previous step (ProposalEngine.cpp:241-294, 341-371, 417-448, 455-546) ----
            if (s.pending) {
                s.pending = false;
                const unsigned long long xw = D.xw;
                const int xg = D.xg, np = D.np, Lx = D.Lx, offb = D.offb, pg = D.pg, n = D.n;
                const bool isC = (xg == GC), isA = (xg == GA);
                const unsigned int Cij = D.Cij, Ci = D.Ci, Cj = D.Cj, Ai = D.Ai, Bj = D.Bj, own = D.own, ownC = D.ownC;
                const int Ac = D.Ac, Bc = D.Bc, Cc = D.Cc;
                const unsigned long long xa = xw & sP.maskA, xb = xw & sP.maskB;
                const double pi0 = isC ? D.pi0c : D.mA1;
                const double inv_pi0 = 1.0 / pi0;
                int ncand;
                double ptot = 0.0;               // sum of the candidate probabilities, index order (vectorSum, ProposalEngine.cpp:585-594)
                // One instruction sequence for the three classes of the chosen lineage (the histories of a warp finish events of
                // different classes in the same step): the mutation candidates scaled by f / pi0 (f = 1 for a one-locus lineage:
                // 1.0 * x is exact), the A^B_k coalescences of a one-locus lineage (none for C), the coalescence within the class,
                // and the three further candidates of a two-locus lineage.
                {
                    const double f = isC ? D.dtheta_e * (double)Cij : 1.0;
#pragma unroll kLsUnrollS
                    for (int k = 0; k < Lx; k++) { const double v = f * m.pv[C::at(k, h)] * inv_pi0; m.pv[C::at(k, h)] = v; ptot += v; }
                    const int npp = isC ? 0 : np;
#pragma unroll kLsUnrollS
                    for (int k = 0; k < npp; k++) ptot += m.pv[C::at(Lx + k, h)];
                    const unsigned int cw = isC ? Cij * (Cij - 1) : own * (own - 1 + ownC);
                    const double v0 = (isC ? (Cij > 1) : (own > 1 || ownC > 0)) ? (double)cw * inv_pi0 : 0.0;
                    m.pv[C::at(Lx + npp, h)] = v0; ptot += v0;
                    ncand = Lx + npp + 1;
                    if (isC) {
                        // 0.5 [pi(a|H') pi(b|H'+a) + pi(b|H') pi(a|H'+b)]   (piSMCJointly :120-132)
                        const double pj = 0.5 * D.mA1 * D.mB2 + 0.5 * D.mB1 * D.mA2;
                        const double v1 = (Ai > 0) ? (double)(Ai * Cij) / D.mA1 : 0.0;   // pi(xa | H'+x-xa) == pi(xa | H')
                        const double v2 = (Bj > 0) ? (double)(Bj * Cij) / D.mB1 : 0.0;
                        const double v3 = D.drho_e * (double)Cij * pj * inv_pi0;
                        m.pv[C::at(Ltot + 1, h)] = v1; m.pv[C::at(Ltot + 2, h)] = v2; m.pv[C::at(Ltot + 3, h)] = v3;
                        ptot += v1; ptot += v2; ptot += v3;
                        ncand = Ltot + 4;
                        s.npi += (unsigned int)(Ltot + 1) + 4u + (Ai > 0 ? 1u : 0u) + (Bj > 0 ? 1u : 0u);
                    } else {
                        s.npi += (unsigned int)(1 + Lx + 5 * np);
                    }
                }
                LS_ASSERT(ncand <= sP.pvmax);
                if (sA.score_off) {
                    // score mode: the candidate vector (and the partner haplotypes that label its A^B_k entries) goes out, nothing
                    // is drawn or applied, the slot takes the next configuration
                    const size_t ob = (size_t)s.h * sA.score_stride;
#pragma unroll 1
                    for (int k = 0; k < ncand && k < sA.score_stride; k++) {
                        sA.score_pv[ob + k] = m.pv[C::at(k, h)];
                        sA.score_pw[ob + k] = (!isC && k >= Lx && k < Lx + np) ? (unsigned long long)m.words[C::at((int)m.plist[C::at(k - Lx, h)], h)] : 0ull;
                    }
                    sA.score_ncand[s.h] = ncand;
                    s.have = false;
                } else {
                // ---- sample the move (sampleFromVector :568-583) ----
                double pidx;
                const int idx = ls_scan<NH, NT, MD>(m, h, ncand, next_u() * ptot, false, pidx);
                // proposal density of the event: waiting time, lineage choice, move (:696-698, :455)
                const double qfac = (D.rate * D.hap_num * pidx) / (D.hap_den * ptot);
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
                    const int pslot = (int)m.plist[C::at(idx - Lx, h)];
                    LS_ASSERT(pslot < s.nslots);
                    const unsigned long long pw = (unsigned long long)m.words[C::at(pslot, h)];
                    og0 = pg; ow0 = pw; od0 = -1; og1 = GC; ow1 = xw | pw; od1 = +1; nops = 2;
                    ev_kind = 5; ev_aux = pw;
                } else {
                    evc = isA ? (double)((unsigned int)Ac * (Ai + Ci - 1)) : (double)((unsigned int)Bc * (Bj + Cj - 1));
                }
                // bookkeeping that does not depend on the store (:457-546); the type changes follow in the cooperative pass
                if (s.tracing && s.nev < sA.max_events) {
                    coalgpu_event& r = sA.events[(unsigned long long)s.h * sA.max_events + s.nev];
                    r.chosen_word = xw; r.aux_word = ev_aux; r.chosen_gamma = (uint8_t)xg; r.kind = (uint8_t)ev_kind;
                    r.epoch = (uint8_t)s.e; r.site = (uint8_t)ev_site; r.n_before = (uint32_t)n;
                }
                s.nev++;
                s.tsum += D.wait;
                s.qsum += -D.rate * D.wait;
                s.qprod.mul(qfac);
                s.st.s_nn += D.nn * D.wait; s.st.s_a += D.aa * D.wait; s.st.s_c += (double)Cc * D.wait;
                if (kind == 1) s.st.n_mut++; else if (kind == 2) s.st.n_rec++;
                if (s.post) { s.st.has_post = 1; s.st.post_nn = D.nn; s.st.post_a = D.aa; s.st.post_c = (double)Cc; }
                else s.st.n_rate++;
                s.post = false;
                mv_op = (nops > 0 ? 1 : 0) | (nops > 1 ? 2 : 0) | (og0 << 2) | (og1 << 4) | 64;      // bit 6: a move was drawn
                mv_w0 = ow0; mv_w1 = ow1; mv_d0 = od0; mv_d1 = od1; mv_kind = ev_kind; mv_evc = evc;
                }
            }
        }
        // =================================== W1: the move's type changes, LPH lanes per history ===================================
        {
            const int opq = __shfl_sync(kFull, mv_op, leader);
            const int nsq = __shfl_sync(kFull, s.nslots, leader);
            const unsigned long long w0q = __shfl_sync(kFull, mv_w0, leader), w1q = __shfl_sync(kFull, mv_w1, leader);
            const int d0q = __shfl_sync(kFull, mv_d0, leader), d1q = __shfl_sync(kFull, mv_d1, leader);
            unsigned long long r = 0ull;
            if (__any_sync(kFull, (opq & 3) != 0)) r = ls_adjust2_coop<NH, NT, MD, LPH>(h, sub, nsq, opq, w0q, d0q, w1q, d1q);
            if (serial && (mv_op & 64)) {
                LsDesc& D = m.desc[h];
                const int og0 = (mv_op >> 2) & 3, og1 = (mv_op >> 4) & 3;
                const int Ac = D.Ac, Bc = D.Bc, Cc = D.Cc;
                double after0 = 1.0, after1 = 1.0;
                int slot0 = -1, slot1 = -1;
                if ((mv_op & 3) != 0) {
                    if (r == 0xffffffffffffffffull) s.overflow = true;
                    else {
                        const int r0 = (int)(unsigned int)r, r1 = (int)(unsigned int)(r >> 32);
                        if (mv_op & 1) {
                            slot0 = r0 & ~kAppended; if (r0 & kAppended) s.nslots++;
                            if (og0 == GA) s.nA += mv_d0; else if (og0 == GB) s.nB += mv_d0; else s.nC += mv_d0;
                            after0 = (double)(m.cg[C::at(slot0, h)] & kCntMask);
                        }
                        if (mv_op & 2) {
                            slot1 = r1 & ~kAppended; if (r1 & kAppended) s.nslots++;
                            if (og1 == GA) s.nA += mv_d1; else if (og1 == GB) s.nB += mv_d1; else s.nC += mv_d1;
                            after1 = (double)(m.cg[C::at(slot1, h)] & kCntMask);
                        }
                    }
                }
                double evc = mv_evc;
                if (mv_kind == 0) evc = after0;
                else if (mv_kind == 4) evc = (double)Cc * after0 * after1 / (((double)Ac + 1.0) * ((double)Bc + 1.0));
                else if (mv_kind == 5) evc = (double)((unsigned int)Ac * (unsigned int)Bc) * after1 / ((double)Cc + 1.0);
                const int ns_event = s.nslots;      // slots in use before the empty ones are dropped: the event's peak
                ls_commit<NH, NT, MD>(m, h, s, D.sx, slot0, slot1);
                if (ns_event > s.peak) s.peak = ns_event;
                s.evprod.mul(evc);
                if (s.nev >= (1u << 18)) s.overflow = true;   // runaway guard: histories have O(10^2-10^3) events
                end_now = s.overflow;
            }
        }
        // =================================== S2: one thread per history ===================================
        if (serial && !s.done) {
            LsDesc& D = m.desc[h];
            D.valid = 0; D.nm = 0; D.nq = 0;

            // ---- fetch the next history ----
            if (!end_now && !s.have) {
                unsigned long long hx = atomicAdd(sA.counters + (sA.list ? 7 : 0), 1ull);
                if (hx >= n_todo) {
                    s.done = true;
                } else {
                    if (sA.list) hx = sA.list[hx];
                    s.h = hx; s.have = true;
                    s.rk = 0; s.rb0 = ~0ull;
                    s.nslots = 0; s.nA = s.nB = s.nC = 0; s.overflow = false;
                    s.tracing = sA.events != nullptr && hx < sA.n_trace;
                    s.nev = 0; s.npi = 0; s.ncq = 0;
                    // initial configuration (TypeContainer ctor, TypeContainer.cpp:42-99), then injections at boundaries
                    s.inj_lo = sP.type_off[0]; s.inj_hi = sP.type_off[1];
                    s.injecting = true; s.post = true; s.wfirst = true;
                    s.e = 0; s.tset = sP.set_of_epoch[0];
                    s.qsum = 0.0; s.consts = 0.0; s.tsum = 0.0; s.peak = 0;
                    s.qprod.clear(); s.evprod.clear(); epoch_stats_clear(s.st);
                    if (sA.score_off) {      // score mode: the given configuration instead of the sampled lineages
#pragma unroll 1
                        for (int k = sA.score_off[hx]; k < sA.score_off[hx + 1]; k++)
                            if (ls_adjust<NH, NT, MD>(m, h, s, (int)sA.score_g[k], sA.score_w[k], (int)sA.score_c[k]) < 0) break;
                        s.injecting = false;
                        s.e = sA.score_epoch[hx]; s.tset = sP.set_of_epoch[s.e];
                        end_now = s.overflow;
                    }
                }
            }

        }
        // =================================== W0: injections, LPH lanes per history ===================================
        // TypeContainer::addTypeMap :525-591 + probabilityCombinatorics ProposalEngine.cpp:66-86. The sampled types of a
        // collection enter the store one after the other (the order fixes the slot order); each is one cooperative pass of the
        // history's lanes, and the histories of a warp that inject in the same step share the rounds.
        {
            const bool inj = serial && !s.done && !end_now && s.injecting;
            const int injq = __shfl_sync(kFull, inj ? 1 : 0, leader);
            if (__any_sync(kFull, injq != 0)) {
                const int loq = __shfl_sync(kFull, s.inj_lo, leader), hiq = __shfl_sync(kFull, s.inj_hi, leader);
                const int nrounds = __reduce_max_sync(kFull, injq ? hiq - loq : 0);
                const int a0 = s.nA, b0 = s.nB, c0 = s.nC;
                double comb = 0.0;
#pragma unroll 1
                for (int r = 0; r < nrounds; r++) {
                    const int k = loq + r;
                    const int nsq = __shfl_sync(kFull, s.nslots, leader);
                    const int stopq = __shfl_sync(kFull, s.overflow ? 1 : 0, leader);
                    const bool act = injq && k < hiq && !stopq;
                    const unsigned int cgk = act ? __ldg(sP.type_cg + k) : 0u;
                    const unsigned long long wk = act ? __ldg(sP.type_w + k) : 0ull;
                    const int add = (int)(cgk & kCntMask), gk = (int)(cgk >> 30);
                    const unsigned long long rr = ls_adjust2_coop<NH, NT, MD, LPH>(h, sub, nsq, act ? (1 | (gk << 2)) : 0, wk, add, 0ull, 0);
                    if (serial && act) {
                        if (rr == 0xffffffffffffffffull) s.overflow = true;
                        else {
                            const int r0 = (int)(unsigned int)rr;
                            const int slot = r0 & ~kAppended; if (r0 & kAppended) s.nslots++;
                            if (gk == GA) s.nA += add; else if (gk == GB) s.nB += add; else s.nC += add;
                            const int now = (int)(m.cg[C::at(slot, h)] & kCntMask);
                            comb += sP.lfact[now] - sP.lfact[add] - sP.lfact[now - add];
                        }
                    }
                }
                if (inj) {
                    if (s.e > 0 && !s.overflow) {
                        comb -= sP.lfact[s.nA] - sP.lfact[s.nA - a0] - sP.lfact[a0];
                        comb -= sP.lfact[s.nB] - sP.lfact[s.nB - b0] - sP.lfact[b0];
                        comb -= sP.lfact[s.nC] - sP.lfact[s.nC - c0] - sP.lfact[c0];
                        s.consts += comb;
                    }
                    if (s.nslots > s.peak) s.peak = s.nslots;
                    s.injecting = false;
                    end_now = s.overflow;
                }
            }
        }
        // =================================== W2: lineage-chooser weights, LPH lanes per history ===================================
        // (TypeContainer::sampleType, TypeContainer.cpp:641-721): pv[j] = weight of slot j; summed in slot order by
        // the serial thread below (the order of the sum is part of the canonical arithmetic)
        {
            __syncwarp();                  // the store as the serial threads left it (appended slots, injected types)
            const int actq = __shfl_sync(kFull, (serial && !s.done && !end_now) ? 1 : 0, leader);
            const int nsq = __shfl_sync(kFull, s.nslots, leader), eq = __shfl_sync(kFull, s.e, leader);
            const unsigned int nAq = (unsigned int)__shfl_sync(kFull, s.nA, leader), nBq = (unsigned int)__shfl_sync(kFull, s.nB, leader);
            const EpochConst* __restrict__ ecq = sP.ep + eq;
            const double dthA = __ldg(&ecq->dthA), dthB = __ldg(&ecq->dthB), drho_e = __ldg(&ecq->drho);
            const int nl = actq ? nsq : 0;
#pragma unroll 1
            for (int j = sub; j < nl; j += LPH) {
                const unsigned int cgj = m.cg[C::at(j, h)], xcj = m.xc[C::at(j, h)];
                const unsigned int c = cgj & kCntMask;
                const int gj = (int)(cgj >> 30);
                // one instruction sequence for the three classes (adding 0.0 to a non-negative sum is exact):
                //   C: (c-1 + thA+thB + rho + cA + cB) c     A: (c-1 + nB + cCa + thA) c     B: (c-1 + nA + cCb + thB) c
                const bool jC = (gj == GC);
                const unsigned int ti = c - 1 + (jC ? 0u : ((gj == GA ? nBq : nAq) + xcj));
                double v = (double)ti + (jC ? (dthA + dthB) : (gj == GA ? dthA : dthB));
                v = v + (jC ? drho_e : 0.0);
                v = v + (jC ? (double)(xcj & 0xffffu) : 0.0);
                v = v + (jC ? (double)(xcj >> 16) : 0.0);
                m.pv[C::at(j, h)] = v * (double)c;
            }
            __syncwarp();
        }
        // =================================== R: Philox blocks ahead, LPH lanes per history ===================================
        {
            const bool act = serial && !s.done && !end_now && !sA.score_off;
            const bool need = act && (s.rb0 == ~0ull || ((s.rk + 2) >> 1) >= s.rb0 + LPH);
            if (__any_sync(kFull, need)) {      // every history of the warp draws ahead from where it stands: fewer refill rounds
                const unsigned long long kq = __shfl_sync(kFull, s.rk, leader), hq = __shfl_sync(kFull, s.h, leader);
                m.rbuf[h * LPH + sub] = philox_block(sA.seed, sA.first + hq, (kq >> 1) + (unsigned long long)sub);
                if (act) s.rb0 = s.rk >> 1;
                __syncwarp();
            }
        }
        // =================================== S3: one thread per history ===================================
        // values that cross from S3 over the cooperative removal pass into S4
        int rm_op = 0, rm_sx = 0;                  // bit 0: an event is being set up; bits 1-2: class of the chosen lineage
        unsigned long long rm_xw = 0ull;
        if (serial && !s.done) {
            LsDesc& D = m.desc[h];
            bool finish = false;
            if (!end_now) {
                // ---- choose a lineage and a waiting time (ProposalEngine.cpp:672-687 / 703-716) ----
                const EpochConst* __restrict__ ecp = sP.ep + s.e;
                const double dthA = __ldg(&ecp->dthA), dthB = __ldg(&ecp->dthB), drho_e = __ldg(&ecp->drho);
                const int ns0 = s.nslots;
                double hap_den = 0.0;
#pragma unroll kLsUnrollS
                for (int j = 0; j < ns0; j++) hap_den += m.pv[C::at(j, h)];
                double hap_num;
                int sx;
                if (sA.score_off) {          // score mode: the given lineage
                    sx = 0; hap_num = 1.0; hap_den = 1.0;
                    const typename WO::T cw = WO::narrow(sA.score_xw[s.h]); const unsigned int cgx = (unsigned int)sA.score_xg[s.h];
#pragma unroll 1
                    for (int j = 0; j < ns0; j++) if (m.words[C::at(j, h)] == cw && (m.cg[C::at(j, h)] >> 30) == cgx) sx = j;
                } else {
                    sx = ls_scan<NH, NT, MD>(m, h, ns0, next_u() * hap_den, true, hap_num);
                }
                const int n = s.nA + s.nB + s.nC, Ac = s.nA, Bc = s.nB, Cc = s.nC;
                const double nn = (double)((unsigned int)n * (unsigned int)(n - 1));
                const double Dd = nn + (double)(Ac + Cc) * dthA + (double)(Bc + Cc) * dthB + drho_e * (double)Cc;
                const double rate = Dd * __ldg(&ecp->inv2Ntau);
                const double tsum = s.tsum;
                const double wait = sA.score_off ? 1.0 : -dlog(next_u()) / rate;
                const double aa = (double)((Ac + Cc) * LA + (Bc + Cc) * LB);

                // ---- collection boundary (ProposalEngine.cpp:721-783) or end of the history (:792, :817-858) ----
                const bool boundary = !sA.score_off && (s.e < T - 1) && !(tsum + wait < sP.times[s.e + 1]);
                finish = !sA.score_off && (s.e == T - 1) && (n <= 5);
                if (boundary || finish) {
                    if (boundary) {
                        const double excess = sP.times[s.e + 1] - tsum;
                        s.tsum = sP.times[s.e + 1];
                        s.qsum += -rate * excess;
                        s.st.s_nn += nn * excess; s.st.s_a += aa * excess; s.st.s_c += (double)Cc * excess;
                        s.st.bnd_nn = nn; s.st.bnd_a = aa; s.st.bnd_c = (double)Cc; s.st.has_bnd = 1; s.st.n_rate += 1;
                    }
                    s.st.slog = logprod_value(s.evprod);
                    LsSnap& sn = m.snap[h];
                    sn.st = s.st; sn.e = s.e; sn.hidx = s.h; sn.fin = 0.0;
                    snap_flags = kLsFlagAccum | (s.wfirst ? kLsFlagFirst : 0);
                    s.wfirst = false;
                    epoch_stats_clear(s.st); s.evprod.clear();
                    if (sA.lineages) sA.lineages[s.h * (unsigned long long)T + s.e] = (double)n;
                    if (finish) {
                        end_now = true;
                    } else {
                        s.e++;
                        s.inj_lo = sP.type_off[s.e]; s.inj_hi = sP.type_off[s.e + 1];
                        s.injecting = true;
                        s.post = true;
                        s.tset = sP.set_of_epoch[s.e];
                    }
                } else {
                    // ---- set up one event (ProposalEngine.cpp:134-238, 298-338, 375-414) ----
                    const unsigned long long xw = (unsigned long long)m.words[C::at(sx, h)];
                    const unsigned int xcg = m.cg[C::at(sx, h)], xxc = m.xc[C::at(sx, h)];
                    const int xg = (int)(xcg >> 30);
                    const unsigned int xcount = xcg & kCntMask;
                    const bool isC = (xg == GC), isA = (xg == GA);
                    unsigned int Cij = 0, Ci = 0, Cj = 0, Ai = 0, Bj = 0;
                    if (isC) { Cij = xcount; Ai = xxc & 0xffffu; Bj = xxc >> 16; }
                    else if (isA) { Ai = xcount; Ci = xxc; } else { Bj = xcount; Cj = xxc; }

#ifdef COAL_TMA_ROWS
                    // Stage what this event reads from the tables -- the row of its two-locus queries (n-1 lineages for a
                    // C lineage, n-2 for a one-locus lineage) and the wEA|wEB tails of the two rows of its one-locus
                    // queries -- with TMA bulk copies; they overlap the rest of the serial phase.
                    {
                        const double* g_m1 = row_ptr(s.tset, n - 1);
                        const double* g_alt = row_ptr(s.tset, isC ? n : n - 2);
                        const unsigned int rb = (unsigned int)sP.stage_doubles * 8u, wb = (unsigned int)(sP.row_doubles - sP.off_wEA) * 8u;
                        unsigned long long* bar = m.mbar + h;
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_expect(bar, rb + 2u * wb);
                        if (rb) bulk_g2s(m.rowbuf + (size_t)h * sP.stage_doubles, isC ? g_m1 : g_alt, rb, bar);
                        bulk_g2s(m.wbuf + (size_t)h * 2 * wlen, g_m1 + sP.off_wEA, wb, bar);
                        bulk_g2s(m.wbuf + (size_t)h * 2 * wlen + wlen, g_alt + sP.off_wEA, wb, bar);
                        D.tma_parity = s.tma_phase; s.tma_phase ^= 1u;
                        D.row_m1 = m.wbuf + (size_t)h * 2 * wlen - sP.off_wEA;
                        D.row_alt = m.wbuf + (size_t)h * 2 * wlen + wlen - sP.off_wEA;
                        D.crow = kStaged ? (m.rowbuf + (size_t)h * sP.stage_doubles) : (isC ? g_m1 : g_alt);
                    }
#else
                    D.row_m1 = row_ptr(s.tset, n - 1);
                    D.row_alt = row_ptr(s.tset, isC ? n : n - 2);
                    D.crow = isC ? D.row_m1 : D.row_alt;
                    D.tma_parity = 0;
#endif
                    // what of the descriptor does not depend on the removal pass
                    const int Lx = isC ? Ltot : (isA ? LA : LB);
                    D.xw = xw; D.xg = xg; D.sx = sx; D.ns = ns0; D.n = n; D.Lx = Lx; D.offb = (xg == GB) ? LA : 0; D.pg = isA ? GB : GA;
                    D.Cij = Cij; D.Ai = Ai; D.Bj = Bj; D.Ac = Ac; D.Bc = Bc; D.Cc = Cc;
                    D.Ci = Ci; D.Cj = Cj;          // one-locus lineage: from its cross count; two-locus: counted by the pass
                    D.cn = isC ? (unsigned int)(n - 1) : (unsigned int)(n - 2 < 0 ? 0 : n - 2);
                    D.nc_alt = isC ? (unsigned int)n : (unsigned int)(n - 2 < 0 ? 0 : n - 2);
                    D.dtheta_e = __ldg(&ecp->dtheta); D.drho_e = drho_e;
                    D.dtheta_own = D.dtheta_e * (double)(isA ? Ai : Bj);
                    D.rate = rate; D.wait = wait; D.hap_num = hap_num; D.hap_den = hap_den; D.nn = nn; D.aa = aa;
                    rm_op = 1 | (xg << 1); rm_sx = sx; rm_xw = xw;
                }
            }
        }
        // =================================== W3: remove the chosen lineage, LPH lanes per history ===================================
        // ONE pass over the stored types: H' = H - x (:168; the cross counts of the related slots follow, as in ls_adjust_core
        // with delta = -1 -- x is stored, nothing is appended), the counts of the C types that share x's A part / B part BEFORE
        // the removal (:149-158), and the partner list = stored types of the other one-locus class in slot order (:300 / :377;
        // the removal does not touch that class): the lanes of a history take LPH consecutive slots per round and rank their
        // partners with a ballot.
        {
            __syncwarp();                  // the descriptors as the serial threads started them
            const int opq = __shfl_sync(kFull, rm_op, leader);
            if (__any_sync(kFull, opq != 0)) {
                const unsigned long long xwq = __shfl_sync(kFull, rm_xw, leader);
                const int nsq = __shfl_sync(kFull, s.nslots, leader);
                const int nl = (opq & 1) ? nsq : 0;
                const int xg = opq >> 1;
                const bool isC = (xg == GC), isA = (xg == GA);
                const int pg = isA ? GB : GA;
                const typename WO::T maskA = WO::narrow(sP.maskA), maskB = WO::narrow(sP.maskB);
                const typename WO::T xwn = WO::narrow(xwq), xa = xwn & maskA, xb = xwn & maskB;
                // match rules of the removal (see ls_adjust_core): which slots' cross counts lose one
                const int t1 = isC ? GA : GC, t2 = isC ? GB : 3;
                const typename WO::T m1 = isA ? maskA : (isC ? (typename WO::T)~(typename WO::T)0 : maskB);
                const typename WO::T k1 = isC ? xa : xwn;
                const unsigned int d1 = (!isC && !isA) ? 65536u : 1u;
                unsigned int Ci = 0, Cj = 0;
                int np = 0;
                const int nrounds = (__reduce_max_sync(kFull, nl) + LPH - 1) / LPH;
                const unsigned int below = (1u << sub) - 1u, mine_mask = (LPH == 32) ? 0xffffffffu : ((1u << LPH) - 1u);
#pragma unroll 1
                for (int i = 0; i < nrounds; i++) {
                    const int j = i * LPH + sub;
                    bool partner = false;
                    if (j < nl) {
                        const typename WO::T wj = m.words[C::at(j, h)];
                        const unsigned int cgj = m.cg[C::at(j, h)];
                        const int gj = (int)(cgj >> 30);
                        const unsigned int cj = cgj & kCntMask;
                        const bool jC = (gj == GC);
                        Ci += (isC && jC && (wj & maskA) == xa) ? cj : 0u;
                        Cj += (isC && jC && (wj & maskB) == xb) ? cj : 0u;
                        const bool h1 = (gj == t1) && ((wj & m1) == k1), h2 = (gj == t2) && (wj == xb);
                        const unsigned int xsub = (h1 ? d1 : 0u) + (h2 ? 1u : 0u);
                        if (xsub) m.xc[C::at(j, h)] -= xsub;
                        partner = !isC && gj == pg && cj > 0;
                    }
                    const unsigned int bits = (__ballot_sync(kFull, partner) >> leader) & mine_mask;
                    if (partner) m.plist[C::at(np + __popc(bits & below), h)] = (unsigned short)j;
                    np += __popc(bits);
                }
#pragma unroll
                for (int o = 1; o < LPH; o <<= 1) { Ci += __shfl_xor_sync(kFull, Ci, o); Cj += __shfl_xor_sync(kFull, Cj, o); }
#ifndef COAL_TMA_ROWS
                // stage the one-locus weight vectors of the event's two table rows for Q1 (the rows were chosen in S3: the
                // descriptor is visible after the __syncwarp at the head of this pass)
                if (opq & 1) {
                    const double* __restrict__ g1 = m.desc[h].row_m1 + sP.off_wEA;
                    const double* __restrict__ g2 = m.desc[h].row_alt + sP.off_wEA;
                    double* dst = m.wbuf + (size_t)h * 2 * wlen;
                    const int tl = sP.row_doubles - sP.off_wEA;
#pragma unroll 1
                    for (int i = sub; i < tl; i += LPH) { dst[i] = __ldg(g1 + i); dst[wlen + i] = __ldg(g2 + i); }
                }
#endif
                __syncwarp();
                // ---- S4: publish the event (ProposalEngine.cpp:134-238, 298-338, 375-414) ----
                if (serial && (rm_op & 1)) {
                    LsDesc& D = m.desc[h];
                    m.cg[C::at(rm_sx, h)] -= 1u;
                    if (isC) s.nC--; else if (isA) s.nA--; else s.nB--;
                    if (isC) { D.Ci = Ci; D.Cj = Cj; }
                    D.np = np;
                    D.nm = isC ? 2 : 1 + D.Lx + np;
                    D.nq = isC ? Ltot + 1 : np;
                    D.own = isA ? D.Ai : D.Bj; D.ownC = isA ? D.Ci : D.Cj;
                    D.valid = 1;
                    s.pending = true;
                    s.ncq += (unsigned int)D.nq;
                    LS_ASSERT(D.Lx + np + 1 <= sP.pvmax && np <= sP.kmax);
                }
            }
        }
        // =================================== S5: histories that end in this step ===================================
        if (serial && !s.done) {
            LsDesc& D = m.desc[h];

            if (end_now) {
                // ---- final weight (ProposalEngine.cpp:817-858) ----
                const double ln2 = 0.69314718055994530942;
                const double q = s.qsum + logprod_value(s.qprod);
                const double fin = s.consts + sP.perm_const - q
                                 - ((double)(LA + LB) * ln2 * (double)s.nC + (double)LA * ln2 * (double)s.nA + (double)LB * ln2 * (double)s.nB);
                LsSnap& sn = m.snap[h];
                sn.hidx = s.h; sn.fin = fin;
                if (s.overflow) snap_flags = kLsFlagAbort;
                else snap_flags |= kLsFlagFinal;       // the final epoch's statistics were snapshotted above (finish)
                if (sA.tmrca) sA.tmrca[s.h] = s.tsum;
                if (sA.nevents) sA.nevents[s.h] = s.nev;
                if (s.tracing && sA.event_counts) sA.event_counts[s.h] = s.nev;
                if (!s.overflow) atomicMax(sA.counters + 8, (unsigned long long)s.peak);
                if (s.overflow) {
                    if (sA.defer_list) sA.defer_list[atomicAdd(sA.counters + 6, 1ull)] = s.h;   // drawn again with the full capacity
                    else atomicAdd(sA.counters + 3, 1ull);
                }
                if (!(s.overflow && sA.defer_list)) { ev_total += s.nev; pi_total += s.npi; cq_total += s.ncq; }
                s.have = false; s.pending = false;
                D.valid = 0; D.nm = 0; D.nq = 0;
            }
        }
        if (serial) m.snap[h].flags = snap_flags;      // also by a thread that just ran out of histories: a stale flag would be scored again
        const int alive = __syncthreads_or((serial && !s.done) ? 1 : 0);
        LS_TICK(tS);
        if (!alive) break;
        // work lists of the query phases: exclusive prefixes of the per-history query counts; every warp scans the NH counts
        // itself (NH <= 32) into its own copy, so no further CTA barrier is needed
        {
            int c1 = 0, c2 = 0;
            if (lane < NH) { c1 = m.desc[lane].nm; c2 = m.desc[lane].nq; }
            int i1 = c1, i2 = c2;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t1 = __shfl_up_sync(kFull, i1, o), t2 = __shfl_up_sync(kFull, i2, o);
                if (lane >= o) { i1 += t1; i2 += t2; }
            }
            if (lane < NH) { pre1[lane + 1] = i1; pre2[lane + 1] = i2; }
            if (lane == 0) { pre1[0] = 0; pre2[0] = 0; }
            __syncwarp();
        }

        // =================================== G: per-grid-point target density of closed epochs ===================================
        if (__syncthreads_or(snap_flags)) {
            if (sA.weights) {
#pragma unroll 1
                for (int hh = 0; hh < NH; hh++) {
                    const int fl = m.snap[hh].flags;
                    if (!fl) continue;
                    double* wout = sA.weights + m.snap[hh].hidx * (unsigned long long)NG;
                    if (fl & kLsFlagAbort) {
                        const double bad = __longlong_as_double(0x7ff8000000000000ll);
                        for (int g = tid; g < NG; g += NT) wout[g] = bad;
                        continue;
                    }
                    const EpochStats st = m.snap[hh].st;
                    const GridConst* gc = sP.grid + (size_t)m.snap[hh].e * NG;
                    const double fin = (fl & kLsFlagFinal) ? m.snap[hh].fin : 0.0;
#pragma unroll 1
                    for (int g = tid; g < NG; g += NT) {
                        const GridConst c = gc[g];
                        double v = (fl & kLsFlagFirst) ? 0.0 : wout[g];
                        v += epoch_log_density(st, c.theta, c.rho, c.inv2Ntau, c.log_theta, c.log_rho, c.log_2Ntau, [](double x) { return dlog(x); });
                        if (fl & kLsFlagFinal) v += fin;
                        wout[g] = v;
                    }
                }
            }
            // the snapshots are rewritten in the next S phase, after the barriers of the query phases
        }

        LS_TICK(tG);
        // =================================== Q1: one-locus queries, one per lane ===================================
        //  C chosen: 0 = A part, 1 = B part, rows n-1 and n (:198-238 with :120-132)
        //  A/B chosen: 0 = x (rows n-1, n-2), 1..Lx = mutants (row n-1), then one per partner b_k:
        //              pi(b_k | H'-b_k+x) on row n-1 and pi(b_k | H'-b_k) on row n-2 (:304-330 / :381-406)
        {
            const int M1 = pre1[NH];
#pragma unroll 1
            for (int it = tid; it < M1; it += NT) {
                const int hh = ls_find<NH>(pre1, it);
                const int qi = it - pre1[hh];
                const LsDesc& D = m.desc[hh];
                LS_ASSERT(hh < NH && qi >= 0 && qi < D.nm && D.valid);
#ifdef COAL_TMA_ROWS
                mbar_wait(m.mbar + hh, D.tma_parity);
#endif
                const int xg = D.xg, Lx = D.Lx, ns = D.ns;
                const bool isC = (xg == GC);
                const bool partner = !isC && qi > Lx;
                int qg; unsigned long long qw64; bool excl = false;
                if (isC) { qg = (qi == 1) ? GB : GA; qw64 = D.xw & ((qi == 1) ? sP.maskB : sP.maskA); }
                else if (partner) { qg = D.pg; qw64 = (unsigned long long)m.words[C::at((int)m.plist[C::at(qi - Lx - 1, hh)], hh)]; excl = true; }
                else { qg = xg; qw64 = (qi >= 1) ? (D.xw ^ (1ull << (D.offb + qi - 1))) : D.xw; }
                const int offw = (qg == GA) ? sP.off_wEA : sP.off_wEB;
                const double powl = (qg == GA) ? sP.pow2negLA : sP.pow2negLB;
                const double* __restrict__ wE1 = m.wbuf + (size_t)hh * 2 * wlen + (offw - sP.off_wEA);   // staged tails of rows n-1 and n / n-2
                const double* __restrict__ wE2 = wE1 + wlen;
                const typename WO::T mask = WO::narrow((qg == GA) ? sP.maskA : sP.maskB), qw = WO::narrow(qw64);
                double num1 = 0.0, num2 = 0.0; unsigned int ncomp = 0;
                // branch-free: the lanes of a warp walk the slots of different histories (different classes at the same j); a
                // slot that is not ancestral at the query's locus enters with count 0 (adding 0 * wE changes no sum)
#pragma unroll kLsUnrollQ
                for (int j = 0; j < ns; j++) {
                    const typename WO::T wj = m.words[C::at(j, hh)];
                    const unsigned int cgj = m.cg[C::at(j, hh)];
                    const int gj = (int)(cgj >> 30);
                    const bool rel = (gj == qg) || (gj == GC);
                    const unsigned int c = rel ? (cgj & kCntMask) - ((excl && gj == qg && wj == qw) ? 1u : 0u) : 0u;
                    const int d = WO::popc((wj ^ qw) & mask);
                    LS_ASSERT(d <= ((qg == GA) ? LA : LB));
                    const double cf = (double)c;
                    num1 += cf * wE1[d];
                    num2 += cf * wE2[d];
                    ncomp += c;
                }
                // fp64 division is a ~25-instruction sequence: reciprocals of lineage counts come from a table
                const double inv_nc = __ldg(sP.rcp + ncomp);
                const double v1 = marginal_finish_rcp(num1, inv_nc, ncomp, (unsigned int)(D.n - 1), powl, sP.wsum);
                const double v2 = marginal_finish_rcp(num2, inv_nc, ncomp, D.nc_alt, powl, sP.wsum);
                LsDesc& Dw = m.desc[hh];
                if (qi == 0) { Dw.mA1 = v1; Dw.mA2 = v2; }
                else if (isC) { Dw.mB1 = v1; Dw.mB2 = v2; }
                else if (!partner) m.pv[C::at(qi - 1, hh)] = D.dtheta_own * v1;          // x 1/pi0 in the next S phase
                else { m.pv[C::at(qi - 1, hh)] = v1; m.aux[C::at(qi - Lx - 1, hh)] = v2; }   // pv[Lx + k], joined in Q2
            }
        }
        __syncthreads();
        LS_TICK(tQ1);

        // =================================== Q2: two-locus queries, one per lane ===================================
        //  C chosen: x and its Ltot one-site mutants against H' (:169, :179-187)
        //  A/B chosen: the joined haplotype (x, b_k) against H' - b_k (:325 / :402)
        {
            const int M2 = pre2[NH];
            const typename WO::T maskA = WO::narrow(sP.maskA), maskB = WO::narrow(sP.maskB);
            unsigned int* mine = m.scratch + tid;          // histogram / class list of this lane: element i at mine[i * NT]
            const bool dense = (LA + 1) * (LB + 1) <= kLsDenseMax;   // uniform over the launch (one locus pair)
#pragma unroll 1
            for (int it = tid; it < M2; it += NT) {
                const int hh = ls_find<NH>(pre2, it);
                const int qi = it - pre2[hh];
                const LsDesc& D = m.desc[hh];
                const bool isC = (D.xg == GC);
                const int ns = D.ns;
                // ---- class list (TypeContainer::getDifferences for a C query, TypeContainer.cpp:802-886) ----
                unsigned long long qw64 = D.xw, delw64 = 0ull; int delg = -1; int pslot = 0;
                if (isC) { if (qi > 0) qw64 = D.xw ^ (1ull << (qi - 1)); }
                else { LS_ASSERT(qi < D.np); pslot = (int)m.plist[C::at(qi, hh)]; LS_ASSERT(pslot < ns); delw64 = (unsigned long long)m.words[C::at(pslot, hh)]; delg = D.pg; qw64 = D.xw | delw64; }
                const typename WO::T qw = WO::narrow(qw64), delw = WO::narrow(delw64);
#pragma unroll 1
                for (int sidx = 0; sidx < nsum; sidx++) mine[sidx * NT] = 0xffff0000u;
                // branch-free (see Q1): a slot without lineages (the removed x, the left-out partner) adds count 0 and leaves
                // the class representative alone
#pragma unroll kLsUnrollQ
                for (int j = 0; j < ns; j++) {
                    const typename WO::T wj = m.words[C::at(j, hh)];
                    const unsigned int cgj = m.cg[C::at(j, hh)];
                    const int gj = (int)(cgj >> 30);
                    const unsigned int c = (cgj & kCntMask) - ((gj == delg && wj == delw) ? 1u : 0u);
                    const typename WO::T y = wj ^ qw;
                    const int da = WO::popc(y & maskA), db = WO::popc(y & maskB);
                    const bool jA = (gj == GA), jB = (gj == GB);
                    const int sidx = jA ? da : (jB ? db : da + db + 1);
                    const unsigned int rep = c ? (jA ? 0u : (jB ? 1u : 2u + (unsigned int)da)) : 0xffffu;
                    LS_ASSERT(sidx < nsum);
                    const unsigned int v = mine[sidx * NT];
                    mine[sidx * NT] = ((v & 0xffffu) + c) | (min(v >> 16, rep) << 16);
                }
                const double* __restrict__ row = D.crow;
                auto ldmy = [row](int k) {
                    const double2 v = kStaged ? *reinterpret_cast<const double2*>(row + 2 * k) : __ldg(reinterpret_cast<const double2*>(row + 2 * k));
                    MYPair e; e.m = v.x; e.y = v.y; return e;
                };
                int nH = 0;
                ClassTotals tot; tot.n = 0; tot.ncompA = 0; tot.ncompB = 0;
                double pi;
                if (dense) {
                    // short loci: scatter the classes (ascending dA+dB) into the two distance vectors, collect the totals and
                    // the term of the classes ancestral at both loci; then the (LA+1)(LB+1) bilinear form (coal_math.h)
                    unsigned int* vecA = mine + nsum * NT;
                    unsigned int* vecB = vecA + (LA + 1) * NT;
#pragma unroll 1
                    for (int i = 0; i < nsum; i++) vecA[i * NT] = 0u;
                    double t2 = 0.0;
#pragma unroll 1
                    for (int sidx = 0; sidx < nsum; sidx++) {
                        const unsigned int v = mine[sidx * NT];
                        const unsigned int cnt = v & 0xffffu;
                        if (cnt) {
                            const unsigned int rep = v >> 16;
                            int dA, dB;
                            if (rep == 0) { dA = sidx; dB = -1; } else if (rep == 1) { dA = -1; dB = sidx; } else { dA = (int)rep - 2; dB = sidx - 1 - dA; }
                            LS_ASSERT(dA <= LA && dB <= LB && (dA >= 0 || dB >= 0));
                            nH++;
                            tot.n += cnt;
                            if (dA >= 0) { tot.ncompA += cnt; vecA[dA * NT] += cnt | (dB < 0 ? cnt << 16 : 0u); }
                            if (dB >= 0) { tot.ncompB += cnt; vecB[dB * NT] += cnt | (dA < 0 ? cnt << 16 : 0u); }
                            if (dA >= 0 && dB >= 0) t2 += (double)cnt * ldmy(dA * sP.SB + dB).y;
                        }
                    }
                    if (D.cn == 0 || nH == 0) pi = sP.pow2negL;      // empty configuration, ConditionalSamplingHMM.cpp:277-298
                    else pi = pismc_bilinear_dense(ldmy, sP.SB, LA, LB, vecA, vecB, NT, t2, tot, __ldg(sP.rcp + tot.ncompA), __ldg(sP.rcp + tot.ncompB),
                                                   __ldg(sP.rcp + tot.n), sP.pow2negLA, sP.pow2negLB);
                } else {
                    // compact the classes in place (ascending dA+dB) and collect the totals
#pragma unroll 1
                    for (int sidx = 0; sidx < nsum; sidx++) {
                        const unsigned int v = mine[sidx * NT];
                        const unsigned int cnt = v & 0xffffu;
                        if (cnt) {
                            const unsigned int rep = v >> 16;
                            int dA, dB;
                            if (rep == 0) { dA = sidx; dB = -1; } else if (rep == 1) { dA = -1; dB = sidx; } else { dA = (int)rep - 2; dB = sidx - 1 - dA; }
                            mine[nH * NT] = pack_class(cnt, dA, dB);
                            nH++;
                            tot.n += cnt;
                            if (dA >= 0) tot.ncompA += cnt;
                            if (dB >= 0) tot.ncompB += cnt;
                        }
                    }
                    // ---- pi_SMC as a bilinear form over the class list (coal_math.h) ----
                    if (D.cn == 0 || nH == 0) pi = sP.pow2negL;
                    else pi = pismc_bilinear(ldmy, sP.SB, LA, LB, mine, NT, nH, tot, __ldg(sP.rcp + tot.ncompA), __ldg(sP.rcp + tot.ncompB),
                                             __ldg(sP.rcp + tot.n), sP.pow2negLA, sP.pow2negLB);
                }
                nh_lane += (unsigned int)nH;
                if (isC) { if (qi == 0) m.desc[hh].pi0c = pi; else m.pv[C::at(qi - 1, hh)] = pi; }
                else {
                    // joint: 0.5 [pi(x|H'-b) pi(b|H'-b+x) + pi(b|H'-b) pi(x|H')]   (:326 / :403 with :120-132)
                    const double v1 = m.pv[C::at(D.Lx + qi, hh)], v2 = m.aux[C::at(qi, hh)];
                    const double joint = 0.5 * D.mA2 * v1 + 0.5 * v2 * D.mA1;
                    m.pv[C::at(D.Lx + qi, hh)] = (double)D.own * (double)(m.cg[C::at(pslot, hh)] & kCntMask) * pi / joint;
                }
            }
        }
        __syncthreads();
        LS_TICK(tQ2);
#ifdef COAL_LS_TIMING
        tsteps++;
#endif
    }
#ifdef COAL_LS_TIMING
    if (tid == 0) {     // phase clocks of this CTA (cycles between the barriers, as seen by thread 0) and its step count
        atomicAdd(sA.counters + 10, (unsigned long long)tS); atomicAdd(sA.counters + 11, (unsigned long long)tG);
        atomicAdd(sA.counters + 12, (unsigned long long)tQ1); atomicAdd(sA.counters + 13, (unsigned long long)tQ2);
        atomicAdd(sA.counters + 14, (unsigned long long)tsteps);
    }
#endif

    // work counters of the launch
    unsigned int nh_warp = warp_sum_u(nh_lane);
    if ((tid & 31) == 0 && nh_warp) atomicAdd(sA.counters + 5, (unsigned long long)nh_warp);
    if (serial) {
        if (ev_total) atomicAdd(sA.counters + 1, ev_total);
        if (pi_total) atomicAdd(sA.counters + 2, pi_total);
        if (cq_total) atomicAdd(sA.counters + 4, cq_total);
    }
}

}  // namespace

// The launch: persistent CTAs of NT threads and NH histories pull history indices from a global counter until the batch is drawn.
// Shapes (NH, NT): (8, 32), (16, 64), (16, 128), (32, 64), (32, 128); the capacity tier (type store in global memory): (16, 64).
#ifndef COAL_LS_MIN_CTAS_32
#define COAL_LS_MIN_CTAS_32 12
#endif
#ifndef COAL_LS_MIN_CTAS_64
#define COAL_LS_MIN_CTAS_64 6
#endif
#ifndef COAL_LS_MIN_CTAS_128
#define COAL_LS_MIN_CTAS_128 3
#endif
template <int NH, int NT, int MD>
__global__ void __launch_bounds__(NT, NT == 32 ? COAL_LS_MIN_CTAS_32 : NT == 64 ? COAL_LS_MIN_CTAS_64 : COAL_LS_MIN_CTAS_128)
history_lockstep_kernel(const DevProblem P, const SampleArgs A) {
    if (threadIdx.x == 0) { sP = P; sA = A; }
    __syncthreads();
    run_lockstep<NH, NT, MD>();
}

// gstore: the capacity tier (type store in global memory), shape (16, 64) only
history_kernel_fn history_lockstep_select(int NH, int NT, bool narrow, bool gstore) {
    if (gstore) return narrow ? history_lockstep_kernel<16, 64, 3> : history_lockstep_kernel<16, 64, 2>;
    if (NH == 8) return narrow ? history_lockstep_kernel<8, 32, 1> : history_lockstep_kernel<8, 32, 0>;
    if (NH == 16 && NT == 128) return narrow ? history_lockstep_kernel<16, 128, 1> : history_lockstep_kernel<16, 128, 0>;
    if (NH == 16) return narrow ? history_lockstep_kernel<16, 64, 1> : history_lockstep_kernel<16, 64, 0>;
    if (NT == 64) return narrow ? history_lockstep_kernel<32, 64, 1> : history_lockstep_kernel<32, 64, 0>;
    return narrow ? history_lockstep_kernel<32, 128, 1> : history_lockstep_kernel<32, 128, 0>;
}
size_t history_lockstep_bytes(int NH, int NT, int kmax, int pvmax, int Ltot, int stage_doubles, bool narrow, bool gstore) {
    return ls_layout(NH, NT, kmax, pvmax, Ltot, stage_doubles, narrow, gstore).total;
}
size_t history_lockstep_store_bytes(int NH, int NT, int kmax, int pvmax, int Ltot, int stage_doubles, bool narrow) {
    return ls_layout(NH, NT, kmax, pvmax, Ltot, stage_doubles, narrow, true).gtotal;
}
