// Kernels of the importance-sampling likelihood path. See coal_device.cuh for the design summary and
// DESIGN.md §4 for the data layout and roofline of each kernel.
#include "coal_device.cuh"
#include "coal_math.h"

namespace coalgpu {

namespace {

constexpr int GA = COALGPU_GAMMA_A, GB = COALGPU_GAMMA_B, GC = COALGPU_GAMMA_C;
constexpr unsigned int kCntMask = 0x3fffffffu;

// per-warp shared-memory state
struct WarpMem {
    unsigned long long* words;   // [kmax]  packed haplotype of each stored type
    unsigned int* cg;            // [kmax]  count | gamma << 30
    unsigned int* xc;            // [kmax]  cross counts: C slot: cntA(A part) | cntB(B part) << 16;
                                 //         A slot: C lineages with this A part; B slot: likewise
    double* pv;                  // [pvmax] unnormalised candidate probabilities of the event
    double* aux;                 // [kmax]  joint conditional probability per coalescence partner
    unsigned int* plist;         // [kmax]  slot index of each partner
    unsigned int* hist;          // [(Ltot+2)*32] one class-histogram column per lane
};

struct HState {
    int nslots, nA, nB, nC;
    bool overflow;
};

__host__ __device__ inline size_t warp_mem_bytes(int kmax, int pvmax, int Ltot) {
    size_t b = 0;
    b += (size_t)kmax * 8;                 // words
    b += (size_t)pvmax * 8;                // pv
    b += (size_t)kmax * 8;                 // aux
    b += (size_t)kmax * 4 * 3;             // cg, xc, plist
    b += (size_t)(Ltot + 2) * 32 * 4;      // hist
    return (b + 15) & ~(size_t)15;
}

__device__ __forceinline__ WarpMem carve(unsigned char* base, int kmax, int pvmax) {
    WarpMem m;
    m.words = reinterpret_cast<unsigned long long*>(base); base += (size_t)kmax * 8;
    m.pv = reinterpret_cast<double*>(base); base += (size_t)pvmax * 8;
    m.aux = reinterpret_cast<double*>(base); base += (size_t)kmax * 8;
    m.cg = reinterpret_cast<unsigned int*>(base); base += (size_t)kmax * 4;
    m.xc = reinterpret_cast<unsigned int*>(base); base += (size_t)kmax * 4;
    m.plist = reinterpret_cast<unsigned int*>(base); base += (size_t)kmax * 4;
    m.hist = reinterpret_cast<unsigned int*>(base);
    return m;
}

// ---- type store ------------------------------------------------------------------------------------

// Change the multiplicity of type (g, w) by delta (TypeContainer::addType / removeType,
// TypeContainer.cpp:409-522) keeping the cross counts of every related slot current. A type that is
// not stored yet is appended; a slot whose count reaches 0 stays until commit(). Returns the slot.
__device__ int adjust(const DevProblem& P, const WarpMem& m, HState& s, int g, unsigned long long w, int delta, int lane) {
    int found = -1;
    unsigned int newxc = 0;
    const unsigned long long wa = w & P.maskA, wb = w & P.maskB;
    for (int base = 0; base < s.nslots; base += 32) {
        const int j = base + lane;
        const bool in = j < s.nslots;
        const unsigned long long wj = in ? m.words[j] : 0ull;
        const unsigned int cgj = in ? m.cg[j] : 0u;
        const int gj = (int)(cgj >> 30);
        const unsigned int cj = cgj & kCntMask;
        const unsigned int b = __ballot_sync(kFull, in && gj == g && wj == w);
        if (b) found = base + __ffs(b) - 1;
        if (in) {
            if (g == GA) {
                if (gj == GC && (wj & P.maskA) == w) { m.xc[j] += (unsigned int)delta; newxc += cj; }
            } else if (g == GB) {
                if (gj == GC && (wj & P.maskB) == w) { m.xc[j] += (unsigned int)(delta * 65536); newxc += cj; }
            } else {
                if (gj == GA && wj == wa) { m.xc[j] += (unsigned int)delta; newxc += cj; }
                if (gj == GB && wj == wb) { m.xc[j] += (unsigned int)delta; newxc += cj << 16; }
            }
        }
    }
    if (found < 0) {
        newxc = warp_sum_u(newxc);
        if (s.nslots >= P.kmax) { s.overflow = true; return -1; }
        found = s.nslots++;
        if (lane == 0) { m.words[found] = w; m.cg[found] = (unsigned int)g << 30; m.xc[found] = newxc; }
    }
    __syncwarp();
    if (lane == 0) m.cg[found] += (unsigned int)delta;
    __syncwarp();
    if (g == GA) s.nA += delta; else if (g == GB) s.nB += delta; else s.nC += delta;
    return found;
}

// Drop empty slots at the end of an event: highest empty slot first, overwritten by the last slot
// (the canonical order rule shared with the oracle's CANONICAL mode).
__device__ void commit(const WarpMem& m, HState& s, int lane) {
    for (;;) {
        int z = -1;
        for (int base = ((s.nslots - 1) / 32) * 32; base >= 0 && z < 0; base -= 32) {
            const int j = base + lane;
            const bool empty = j < s.nslots && (m.cg[j] & kCntMask) == 0;
            const unsigned int b = __ballot_sync(kFull, empty);
            if (b) z = base + 31 - __clz(b);
        }
        if (z < 0) break;
        const int last = s.nslots - 1;
        if (lane == 0 && z != last) { m.words[z] = m.words[last]; m.cg[z] = m.cg[last]; m.xc[z] = m.xc[last]; }
        s.nslots = last;
        __syncwarp();
        if (s.nslots == 0) break;
    }
}

// Lineage chooser (TypeContainer::sampleType, TypeContainer.cpp:641-745): slot order, weights of
// SURVEY.md Q10. Returns the slot; prob = weight / normalising constant.
__device__ int sample_type(const WarpMem& m, const HState& s, double thA, double thB, double rho, double u, int lane, double& prob) {
    double norm = 0.0;
    const int rounds = (s.nslots + 31) / 32;
    // pass 1: normalising constant (inclusive scans keep slot order)
    double w0 = 0.0, cum0 = 0.0;
    for (int r = 0; r < rounds; r++) {
        const int j = r * 32 + lane;
        double wgt = 0.0;
        if (j < s.nslots) {
            const unsigned int cgj = m.cg[j], xcj = m.xc[j];
            const unsigned int c = cgj & kCntMask;
            const int gj = (int)(cgj >> 30);
            if (c > 0) {
                if (gj == GC) wgt = ((((double)(c - 1) + (thA + thB)) + rho) + (double)(xcj & 0xffffu) + (double)(xcj >> 16)) * (double)c;
                else if (gj == GA) wgt = ((double)(c - 1 + (unsigned int)s.nB + xcj) + thA) * (double)c;
                else wgt = ((double)(c - 1 + (unsigned int)s.nA + xcj) + thB) * (double)c;
            }
        }
        const double cum = warp_incl_scan(wgt, lane) + norm;
        if (r == 0) { w0 = wgt; cum0 = cum; }
        norm = __shfl_sync(kFull, cum, 31);
    }
    const double thr = u * norm;
    int pick = -1; double wpick = 0.0;
    if (rounds == 1) {
        const unsigned int b = __ballot_sync(kFull, (lane < s.nslots) && (cum0 > thr) && (w0 > 0.0));
        if (b) { pick = __ffs(b) - 1; wpick = __shfl_sync(kFull, w0, pick); }
    } else {
        double carry = 0.0;
        for (int r = 0; r < rounds && pick < 0; r++) {
            const int j = r * 32 + lane;
            double wgt = 0.0;
            if (j < s.nslots) {
                const unsigned int cgj = m.cg[j], xcj = m.xc[j];
                const unsigned int c = cgj & kCntMask;
                const int gj = (int)(cgj >> 30);
                if (c > 0) {
                    if (gj == GC) wgt = ((((double)(c - 1) + (thA + thB)) + rho) + (double)(xcj & 0xffffu) + (double)(xcj >> 16)) * (double)c;
                    else if (gj == GA) wgt = ((double)(c - 1 + (unsigned int)s.nB + xcj) + thA) * (double)c;
                    else wgt = ((double)(c - 1 + (unsigned int)s.nA + xcj) + thB) * (double)c;
                }
            }
            const double cum = warp_incl_scan(wgt, lane) + carry;
            const unsigned int b = __ballot_sync(kFull, (j < s.nslots) && (cum > thr) && (wgt > 0.0));
            if (b) { const int l = __ffs(b) - 1; pick = r * 32 + l; wpick = __shfl_sync(kFull, wgt, l); }
            carry = __shfl_sync(kFull, cum, 31);
        }
    }
    if (pick < 0) {   // u*norm rounded past the last partial sum: take the last stored type
        pick = s.nslots - 1;
        const unsigned int cgj = m.cg[pick], xcj = m.xc[pick];
        const unsigned int c = cgj & kCntMask; const int gj = (int)(cgj >> 30);
        if (gj == GC) wpick = ((((double)(c - 1) + (thA + thB)) + rho) + (double)(xcj & 0xffffu) + (double)(xcj >> 16)) * (double)c;
        else if (gj == GA) wpick = ((double)(c - 1 + (unsigned int)s.nB + xcj) + thA) * (double)c;
        else wpick = ((double)(c - 1 + (unsigned int)s.nA + xcj) + thB) * (double)c;
    }
    prob = wpick / norm;
    return pick;
}

// ---- conditional sampling probabilities ----------------------------------------------------------------

// One-locus query per lane: query (qg in {A,B}, qw), optionally leaving out one copy of the query's own
// type (the H - b_k configurations of ProposalEngine.cpp:323-329). Evaluated against two table rows at once
// (the same distance histogram serves pi(.|H) and pi(.|H + other-locus lineage), ProposalEngine.cpp:120-132).
__device__ __forceinline__ void marginal_query(const DevProblem& P, const WarpMem& m, int nslots, int qg, unsigned long long qw,
                                               bool excl, const double* __restrict__ wE1, const double* __restrict__ wE2,
                                               double& num1, double& num2, unsigned int& ncomp) {
    num1 = 0.0; num2 = 0.0; ncomp = 0;
    const unsigned long long mask = (qg == GA) ? P.maskA : P.maskB;
    for (int j = 0; j < nslots; j++) {
        const unsigned long long wj = m.words[j];
        const unsigned int cgj = m.cg[j];
        const int gj = (int)(cgj >> 30);
        unsigned int c = cgj & kCntMask;
        if (gj == qg || gj == GC) {
            if (excl && gj == qg && wj == qw) c -= 1;
            const int d = __popcll((wj ^ qw) & mask);
            const double cf = (double)c;
            num1 += cf * __ldg(wE1 + d);
            num2 += cf * __ldg(wE2 + d);
            ncomp += c;
        }
    }
}

// Two-locus query. Lanes are organised in groups of p (power of two): every lane of a group builds the
// class histogram of the group's query in its own shared-memory column (TypeContainer::getDifferences,
// TypeContainer.cpp:802-886 incl. the dA+dB collapse of SURVEY.md Q1), then takes Q/p of the intervals of
// the forward pass. (delg, delw): one copy of that stored type is left out (delg < 0: none).
__device__ double cquery(const DevProblem& P, const WarpMem& m, int nslots, int lane, int p, unsigned long long qw,
                         int delg, unsigned long long delw, const double* __restrict__ row, unsigned int n_cond) {
    unsigned int* col = m.hist + lane;
    const int nsum = P.Ltot + 2;
    for (int sidx = 0; sidx < nsum; sidx++) col[sidx * 32] = 0xffff0000u;
    for (int j = 0; j < nslots; j++) {
        const unsigned long long wj = m.words[j];
        const unsigned int cgj = m.cg[j];
        const int gj = (int)(cgj >> 30);
        unsigned int c = cgj & kCntMask;
        if (gj == delg && wj == delw) c -= 1;
        if (c > 0) {
            const unsigned long long x = wj ^ qw;
            const int da = __popcll(x & P.maskA), db = __popcll(x & P.maskB);
            int sidx; unsigned int rep;
            if (gj == GA) { sidx = da; rep = 0; } else if (gj == GB) { sidx = db; rep = 1; } else { sidx = da + db + 1; rep = 2 + da; }
            const unsigned int v = col[sidx * 32];
            const unsigned int cnt = (v & 0xffffu) + c;
            const unsigned int r = min(v >> 16, rep);
            col[sidx * 32] = cnt | (r << 16);
        }
    }
    // compact the classes in place (ascending dA+dB) and collect the totals
    int nH = 0;
    ClassTotals tot; tot.n = 0; tot.ncompA = 0; tot.ncompB = 0;
    for (int sidx = 0; sidx < nsum; sidx++) {
        const unsigned int v = col[sidx * 32];
        const unsigned int cnt = v & 0xffffu;
        if (cnt) {
            const unsigned int rep = v >> 16;
            int dA, dB;
            if (rep == 0) { dA = sidx; dB = -1; } else if (rep == 1) { dA = -1; dB = sidx; } else { dA = (int)rep - 2; dB = sidx - 1 - dA; }
            col[nH * 32] = pack_class(cnt, dA, dB);
            nH++;
            tot.n += cnt;
            if (dA >= 0) tot.ncompA += cnt;
            if (dB >= 0) tot.ncompB += cnt;
        }
    }
    const bool empty = (n_cond == 0 || nH == 0);     // empty configuration, ConditionalSamplingHMM.cpp:277-298
    const int r = lane & (p - 1);
    const int per = kQd / p;
    auto ld = [row](int off) { return __ldg(row + off); };
    ChunkPartial cp = {0, 0, 0, 0, 0, 0};
    if (!empty) cp = chunk_forward(ld, c_w, P.off_yw, P.off_zlow, P.off_g, P.off_zdiag, P.off_EA, P.off_EB,
                                   P.LA, P.LB, col, 32, nH, tot, P.pow2negLA, P.pow2negLB, r * per, r * per + per);
    // exclusive prefixes of (pre, preB) over the lanes of the group
    double ipre = cp.pre, ipreB = cp.preB;
    for (int o = 1; o < p; o <<= 1) {
        const double t1 = __shfl_up_sync(kFull, ipre, o, p);
        const double t2 = __shfl_up_sync(kFull, ipreB, o, p);
        if (r >= o) { ipre += t1; ipreB += t2; }
    }
    double epre = __shfl_up_sync(kFull, ipre, 1, p), epreB = __shfl_up_sync(kFull, ipreB, 1, p);
    if (r == 0) { epre = 0.0; epreB = 0.0; }
    double acc = chunk_finish(cp, epre, epreB), acc2 = cp.acc2;
    for (int o = 1; o < p; o <<= 1) {
        acc += __shfl_xor_sync(kFull, acc, o);
        acc2 += __shfl_xor_sync(kFull, acc2, o);
    }
    return empty ? P.pow2negL : pismc_finish(acc, acc2, tot.n);
}

__device__ __forceinline__ int lanes_per_query(int nq) {
    int p = 1;
    while (p < kQd && nq * (p * 2) <= 32) p *= 2;
    return p;
}

__device__ __forceinline__ const double* row_ptr(const DevProblem& P, int set, int n_all) {
    const int n = n_all < 1 ? 1 : n_all;   // n_all == 0 is never dereferenced for a result
    return P.rows + ((size_t)set * P.n_max + (n - 1)) * P.row_doubles;
}

// ---- one history ------------------------------------------------------------------------------------------

__device__ void run_history(const DevProblem& P, const SampleArgs& A, const WarpMem& m, unsigned long long h, int lane,
                            unsigned long long& ev_total, unsigned long long& pi_total) {
    const unsigned long long hist = A.first + h;
    unsigned long long draw = 0;
    HState s; s.nslots = 0; s.nA = s.nB = s.nC = 0; s.overflow = false;
    const bool tracing = A.events != nullptr && h < A.n_trace;
    unsigned int nev = 0, npi = 0;

    // initial configuration (TypeContainer ctor, TypeContainer.cpp:42-99)
    for (int k = P.type_off[0]; k < P.type_off[1]; k++) {
        const unsigned int cg = P.type_cg[k];
        adjust(P, m, s, (int)(cg >> 30), P.type_w[k], (int)(cg & kCntMask), lane);
    }

    int e = 0;                                  // epoch = index into viral / HMM time_idx
    EpochConst ec = P.ep[0];
    int tset = P.set_of_epoch[0];
    double q = 0.0, consts = 0.0;
    double tsum = 0.0;
    bool post = true;
    EpochStats st; epoch_stats_clear(st);
    // weights are accumulated per epoch into the output buffer at the end of each epoch (lanes over g)
    double* wout = A.weights ? A.weights + h * (unsigned long long)P.G : nullptr;
    if (wout) for (int g = lane; g < P.G; g += 32) wout[g] = 0.0;

    auto flush_epoch = [&](int epoch) {
        if (wout) {
            const GridConst* gc = P.grid + (size_t)epoch * P.G;
            for (int g = lane; g < P.G; g += 32) {
                const GridConst c = gc[g];
                wout[g] += epoch_log_density(st, c.theta, c.rho, c.inv2Ntau, c.log_theta, c.log_rho, c.log_2Ntau,
                                             [](double x) { return log(x); });
            }
        }
        epoch_stats_clear(st);
    };

    for (;;) {
        // ---- choose a lineage and a waiting time (ProposalEngine.cpp:672-687 / 703-716) ----
        double p_hap;
        const double u_type = philox_uniform(A.seed, hist, draw++);
        const int sx = sample_type(m, s, ec.dthA, ec.dthB, ec.drho, u_type, lane, p_hap);
        const int n = s.nA + s.nB + s.nC, Ac = s.nA, Bc = s.nB, Cc = s.nC;
        const double Dd = (double)((unsigned int)n * (unsigned int)(n - 1)) + (double)(Ac + Cc) * ec.dthA + (double)(Bc + Cc) * ec.dthB + ec.drho * (double)Cc;
        const double rate = Dd * ec.inv2Ntau;
        const double u_wait = philox_uniform(A.seed, hist, draw++);
        const double wait = -log(u_wait) / rate;

        // ---- collection boundaries (ProposalEngine.cpp:694-783); a boundary redraws lineage and time ----
        if (e < P.T - 1) {
            if (!(tsum + wait < P.times[e + 1])) {
                const double excess = P.times[e + 1] - tsum;
                tsum = P.times[e + 1];
                q += -rate * excess;
                const double nn = (double)((unsigned int)n * (unsigned int)(n - 1));
                const double aa = (double)((Ac + Cc) * P.LA + (Bc + Cc) * P.LB);
                st.s_nn += nn * excess; st.s_a += aa * excess; st.s_c += (double)Cc * excess;
                st.bnd_nn = nn; st.bnd_a = aa; st.bnd_c = (double)Cc; st.has_bnd = 1; st.n_rate += 1;
                flush_epoch(e);
                if (A.lineages) { if (lane == 0) A.lineages[h * (unsigned long long)P.T + e] = (double)n; }
                // inject the next sample (TypeContainer::addTypeMap :525-591) + combinatorial terms (:66-86)
                const int a0 = s.nA, b0 = s.nB, c0 = s.nC;
                double comb = 0.0;
                for (int k = P.type_off[e + 1]; k < P.type_off[e + 2]; k++) {
                    const unsigned int cg = P.type_cg[k];
                    const int add = (int)(cg & kCntMask);
                    const int slot = adjust(P, m, s, (int)(cg >> 30), P.type_w[k], add, lane);
                    if (slot < 0) break;
                    const int now = (int)(m.cg[slot] & kCntMask);
                    comb += P.lfact[now] - P.lfact[add] - P.lfact[now - add];
                }
                comb -= P.lfact[s.nA] - P.lfact[s.nA - a0] - P.lfact[a0];
                comb -= P.lfact[s.nB] - P.lfact[s.nB - b0] - P.lfact[b0];
                comb -= P.lfact[s.nC] - P.lfact[s.nC - c0] - P.lfact[c0];
                consts += comb;
                post = true;
                e++;
                ec = P.ep[e];
                tset = P.set_of_epoch[e];
                if (s.overflow) break;
                continue;
            }
        } else if (n <= 5) {
            break;                               // ProposalEngine.cpp:792: stop at five lineages
        }
        tsum += wait;

        // ---- one event (ProposalEngine.cpp:134-566) ----
        q += (log(rate) - rate * wait) + log(p_hap);
        const unsigned long long xw = m.words[sx];
        const unsigned int xcg = m.cg[sx], xxc = m.xc[sx];
        const int xg = (int)(xcg >> 30);
        const unsigned int xcount = xcg & kCntMask;
        const unsigned long long xa = xw & P.maskA, xb = xw & P.maskB;
        unsigned int Cij = 0, Ci = 0, Cj = 0, Ai = 0, Bj = 0;
        if (xg == GC) {
            Cij = xcount; Ai = xxc & 0xffffu; Bj = xxc >> 16;
            unsigned int ci = 0, cj = 0;
            for (int base = 0; base < s.nslots; base += 32) {
                const int j = base + lane;
                if (j < s.nslots) {
                    const unsigned int cgj = m.cg[j];
                    if ((int)(cgj >> 30) == GC) {
                        const unsigned long long wj = m.words[j];
                        if ((wj & P.maskA) == xa) ci += cgj & kCntMask;
                        if ((wj & P.maskB) == xb) cj += cgj & kCntMask;
                    }
                }
            }
            Ci = warp_sum_u(ci); Cj = warp_sum_u(cj);
        } else if (xg == GA) { Ai = xcount; Ci = xxc; } else { Bj = xcount; Cj = xxc; }

        adjust(P, m, s, xg, xw, -1, lane);      // H' = H - x  (:168)
        const int ns = s.nslots;
        const double* row_m1 = row_ptr(P, tset, n - 1);
        const double* row_0 = row_ptr(P, tset, n);
        const double* row_m2 = row_ptr(P, tset, n - 2);
        int ncand = 0;
        double pi0 = 0.0;

        if (xg == GC) {
            // two-locus queries: x itself and its Ltot one-site mutants, all against H' (:169, :179-187)
            const int nq = P.Ltot + 1;
            for (int q0 = 0; q0 < nq; q0 += 32) {
                const int left = nq - q0;
                const int p = left >= 32 ? 1 : lanes_per_query(left);
                const int qi = q0 + lane / p;
                const bool valid = qi < nq;
                const unsigned long long qw = (valid && qi > 0) ? (xw ^ (1ull << (qi - 1))) : xw;
                const double pi = cquery(P, m, ns, lane, p, qw, -1, 0ull, row_m1, (unsigned int)(n - 1));
                if (valid && (lane & (p - 1)) == 0) { if (qi == 0) m.aux[0] = pi; else m.pv[qi - 1] = pi; }
            }
            npi += (unsigned int)nq;
            // one-locus queries: lane 0 the A part, lane 1 the B part, each against rows n-1 and n (:198-238)
            double num1 = 0.0, num2 = 0.0; unsigned int ncomp = 0;
            {
                const int qg = (lane == 1) ? GB : GA;
                const unsigned long long qw = (lane == 1) ? xb : xa;
                const int offw = (lane == 1) ? P.off_wEB : P.off_wEA;
                if (lane < 2) marginal_query(P, m, ns, qg, qw, false, row_m1 + offw, row_0 + offw, num1, num2, ncomp);
            }
            npi += 4u + (Ai > 0 ? 1u : 0u) + (Bj > 0 ? 1u : 0u);
            const double powl = (lane == 1) ? P.pow2negLB : P.pow2negLA;
            const double r_m1 = marginal_finish(num1, ncomp, (unsigned int)(n - 1), powl, P.wsum);
            const double r_0 = marginal_finish(num2, ncomp, (unsigned int)n, powl, P.wsum);
            const double piA_m1 = __shfl_sync(kFull, r_m1, 0), piA_0 = __shfl_sync(kFull, r_0, 0);
            const double piB_m1 = __shfl_sync(kFull, r_m1, 1), piB_0 = __shfl_sync(kFull, r_0, 1);
            __syncwarp();
            pi0 = m.aux[0];
            // 0.5 [pi(a|H') pi(b|H'+a) + pi(b|H') pi(a|H'+b)]   (piSMCJointly :120-132)
            const double pj = 0.5 * piA_m1 * piB_0 + 0.5 * piB_m1 * piA_0;
            ncand = P.Ltot + 4;
            for (int k = lane; k < P.Ltot; k += 32) m.pv[k] = ec.dtheta * (double)Cij * m.pv[k] / pi0;
            if (lane == 0) {
                m.pv[P.Ltot] = (Cij > 1) ? (double)(Cij * (Cij - 1)) / pi0 : 0.0;
                m.pv[P.Ltot + 1] = (Ai > 0) ? (double)(Ai * Cij) / piA_m1 : 0.0;   // pi(xa | H'+x-xa) == pi(xa | H')
                m.pv[P.Ltot + 2] = (Bj > 0) ? (double)(Bj * Cij) / piB_m1 : 0.0;
                m.pv[P.Ltot + 3] = ec.drho * (double)Cij * pj / pi0;
            }
        } else {
            // A-only (or, mirrored, B-only) lineage chosen (:298-449)
            const bool isA = (xg == GA);
            const int Lx = isA ? P.LA : P.LB, offb = isA ? 0 : P.LA;
            const int pg = isA ? GB : GA;
            const int off_own = isA ? P.off_wEA : P.off_wEB, off_oth = isA ? P.off_wEB : P.off_wEA;
            const double pow_own = isA ? P.pow2negLA : P.pow2negLB, pow_oth = isA ? P.pow2negLB : P.pow2negLA;
            const unsigned int own = isA ? Ai : Bj, ownC = isA ? Ci : Cj;
            // partner list = stored types of the other one-locus class, slot order (:300 / :377)
            int np = 0;
            for (int base = 0; base < ns; base += 32) {
                const int j = base + lane;
                const bool is = j < ns && (int)(m.cg[j] >> 30) == pg && (m.cg[j] & kCntMask) > 0;
                const unsigned int b = __ballot_sync(kFull, is);
                if (is) m.plist[np + __popc(b & ((1u << lane) - 1))] = (unsigned int)j;
                np += __popc(b);
            }
            __syncwarp();
            // one-locus queries, one per lane: 0 = x (rows n-1 and n-2), 1..Lx = mutants (row n-1),
            // then one per partner b_k: pi(b_k | H'-b_k+x) on row n-1 and pi(b_k | H'-b_k) on row n-2
            const int nm = 1 + Lx + np;
            double pi_x_m2 = 0.0;
            for (int q0 = 0; q0 < nm; q0 += 32) {
                const int qi = q0 + lane;
                const bool valid = qi < nm;
                const bool partner = valid && qi > Lx;
                int qg = xg; unsigned long long qw = xw; bool excl = false; int offw = off_own; double powl = pow_own;
                if (valid && qi >= 1 && qi <= Lx) qw = xw ^ (1ull << (offb + qi - 1));
                if (partner) { qg = pg; qw = m.words[m.plist[qi - Lx - 1]]; excl = true; offw = off_oth; powl = pow_oth; }
                double num1 = 0.0, num2 = 0.0; unsigned int ncomp = 0;
                if (valid) marginal_query(P, m, ns, qg, qw, excl, row_m1 + offw, row_m2 + offw, num1, num2, ncomp);
                const double v1 = marginal_finish(num1, ncomp, (unsigned int)(n - 1), powl, P.wsum);
                const double v2 = marginal_finish(num2, ncomp, (unsigned int)(n - 2 < 0 ? 0 : n - 2), powl, P.wsum);
                if (q0 == 0) { pi0 = __shfl_sync(kFull, v1, 0); pi_x_m2 = __shfl_sync(kFull, v2, 0); }
                if (valid && qi >= 1 && qi <= Lx) m.pv[qi - 1] = ec.dtheta * (double)own * v1 / pi0;
                // joint: 0.5 [pi(x|H'-b) pi(b|H'-b+x) + pi(b|H'-b) pi(x|H')]   (:326 / :403 with :120-132)
                if (partner) m.aux[qi - Lx - 1] = 0.5 * pi_x_m2 * v1 + 0.5 * v2 * pi0;
            }
            npi += (unsigned int)(1 + Lx + 5 * np);
            __syncwarp();
            // two-locus queries: the joined haplotype (x, b_k) against H' - b_k  (:325 / :402)
            for (int q0 = 0; q0 < np; q0 += 32) {
                const int left = np - q0;
                const int p = left >= 32 ? 1 : lanes_per_query(left);
                const int qi = q0 + lane / p;
                const bool valid = qi < np;
                const int pslot = (int)m.plist[valid ? qi : q0];
                const unsigned long long pw = m.words[pslot];
                const double pi = cquery(P, m, ns, lane, p, xw | pw, pg, pw, row_m2, (unsigned int)(n - 2));
                if (valid && (lane & (p - 1)) == 0) {
                    const double pc = (double)(m.cg[pslot] & kCntMask);
                    m.pv[Lx + qi] = (double)own * pc * pi / m.aux[qi];
                }
            }
            if (lane == 0) m.pv[Lx + np] = (own > 1 || ownC > 0) ? (double)(own * (own - 1 + ownC)) / pi0 : 0.0;
            ncand = Lx + np + 1;
        }
        __syncwarp();

        // ---- sample the move (sampleFromVector :568-583) ----
        double total = 0.0;
        for (int base = 0; base < ncand; base += 32) {
            const int k = base + lane;
            const double v = k < ncand ? m.pv[k] : 0.0;
            total = __shfl_sync(kFull, warp_incl_scan(v, lane) + total, 31);
        }
        const double u_move = philox_uniform(A.seed, hist, draw++);
        const double target = u_move * total;
        int idx = -1; double pidx = 0.0;
        {
            double carry = 0.0;
            for (int base = 0; base < ncand && idx < 0; base += 32) {
                const int k = base + lane;
                const double v = k < ncand ? m.pv[k] : 0.0;
                const double cum = warp_incl_scan(v, lane) + carry;
                const unsigned int b = __ballot_sync(kFull, k < ncand && !(cum < target));
                if (b) { const int l = __ffs(b) - 1; idx = base + l; pidx = __shfl_sync(kFull, v, l); }
                carry = __shfl_sync(kFull, cum, 31);
            }
            if (idx < 0) { idx = ncand - 1; pidx = m.pv[idx]; }
        }
        q += log(pidx) - log(total);

        // ---- apply the move, event constant (:242-294, :342-371, :418-448) ----
        double evc = 1.0; int kind = 0;          // kind: 0 coalescence, 1 mutation, 2 recombination
        int ev_kind = 0; unsigned long long ev_aux = 0ull; int ev_site = 0xff;
        if (xg == GC) {
            if (idx < P.Ltot) {
                const unsigned long long mw = xw ^ (1ull << idx);
                const int slot = adjust(P, m, s, GC, mw, +1, lane);
                evc = slot >= 0 ? (double)(m.cg[slot] & kCntMask) : 1.0; kind = 1;
                ev_kind = 0; ev_aux = mw; ev_site = idx;
            } else if (idx == P.Ltot) {
                evc = (double)((unsigned int)Cc * (Cij - 1)); ev_kind = 1;
            } else if (idx == P.Ltot + 1) {
                adjust(P, m, s, GC, xw, +1, lane); adjust(P, m, s, GA, xa, -1, lane);
                evc = (double)((unsigned int)Ac * Cij); ev_kind = 2; ev_aux = xa;
            } else if (idx == P.Ltot + 2) {
                adjust(P, m, s, GC, xw, +1, lane); adjust(P, m, s, GB, xb, -1, lane);
                evc = (double)((unsigned int)Bc * Cij); ev_kind = 3; ev_aux = xb;
            } else {
                const int sa = adjust(P, m, s, GA, xa, +1, lane);
                const int sb = adjust(P, m, s, GB, xb, +1, lane);
                const double Ai2 = sa >= 0 ? (double)(m.cg[sa] & kCntMask) : 1.0, Bj2 = sb >= 0 ? (double)(m.cg[sb] & kCntMask) : 1.0;
                evc = (double)Cc * Ai2 * Bj2 / (((double)Ac + 1.0) * ((double)Bc + 1.0)); kind = 2; ev_kind = 4;
            }
        } else {
            const bool isA = (xg == GA);
            const int Lx = isA ? P.LA : P.LB, offb = isA ? 0 : P.LA;
            const int np = ncand - Lx - 1;
            if (idx < Lx) {
                const unsigned long long mw = xw ^ (1ull << (offb + idx));
                const int slot = adjust(P, m, s, xg, mw, +1, lane);
                evc = slot >= 0 ? (double)(m.cg[slot] & kCntMask) : 1.0; kind = 1;
                ev_kind = 0; ev_aux = mw; ev_site = idx;
            } else if (idx < Lx + np) {
                const unsigned long long pw = m.words[m.plist[idx - Lx]];
                adjust(P, m, s, isA ? GB : GA, pw, -1, lane);
                const int slot = adjust(P, m, s, GC, xw | pw, +1, lane);
                const double Cnew = slot >= 0 ? (double)(m.cg[slot] & kCntMask) : 1.0;
                evc = (double)((unsigned int)Ac * (unsigned int)Bc) * Cnew / ((double)Cc + 1.0);
                ev_kind = 5; ev_aux = pw;
            } else {
                evc = isA ? (double)((unsigned int)Ac * (Ai + Ci - 1)) : (double)((unsigned int)Bc * (Bj + Cj - 1));
                ev_kind = 1;
            }
        }
        commit(m, s, lane);
        if (tracing && lane == 0) {
            if (nev < A.max_events) {
                coalgpu_event& r = A.events[(unsigned long long)h * A.max_events + nev];
                r.chosen_word = xw; r.aux_word = ev_aux; r.chosen_gamma = (uint8_t)xg; r.kind = (uint8_t)ev_kind;
                r.epoch = (uint8_t)e; r.site = (uint8_t)ev_site; r.n_before = (uint32_t)n;
            }
        }
        nev++;

        // ---- target density bookkeeping (:457-546) ----
        {
            const double nn = (double)((unsigned int)n * (unsigned int)(n - 1));
            const double aa = (double)((Ac + Cc) * P.LA + (Bc + Cc) * P.LB);
            st.slog += log(evc);
            st.s_nn += nn * wait; st.s_a += aa * wait; st.s_c += (double)Cc * wait;
            if (kind == 1) st.n_mut++; else if (kind == 2) st.n_rec++;
            if (post) { st.has_post = 1; st.post_nn = nn; st.post_a = aa; st.post_c = (double)Cc; post = false; }
            else st.n_rate++;
        }
        if (s.overflow) break;
        if (nev >= (1u << 18)) { s.overflow = true; break; }   // runaway guard: histories have O(10^2-10^3) events
    }

    // ---- final weight (ProposalEngine.cpp:817-858) ----
    flush_epoch(e);
    const int nt = s.nA + s.nB + s.nC;
    const double ln2 = 0.69314718055994530942;
    const double fin = consts + P.perm_const - q
                     - ((double)(P.LA + P.LB) * ln2 * (double)s.nC + (double)P.LA * ln2 * (double)s.nA + (double)P.LB * ln2 * (double)s.nB);
    if (wout) {
        __syncwarp();
        const double bad = __longlong_as_double(0x7ff8000000000000ll);
        for (int g = lane; g < P.G; g += 32) wout[g] = s.overflow ? bad : wout[g] + fin;
    }
    if (lane == 0) {
        if (A.tmrca) A.tmrca[h] = tsum;
        if (A.lineages) A.lineages[h * (unsigned long long)P.T + (P.T - 1)] = (double)nt;
        if (A.nevents) A.nevents[h] = nev;
        if (tracing && A.event_counts) A.event_counts[h] = nev;
        if (s.overflow) atomicAdd(A.counters + 3, 1ull);
    }
    ev_total += nev; pi_total += npi;
}

}  // namespace

// K1: persistent warps pull history indices from a global counter until the batch is drawn.
__global__ void __launch_bounds__(128) history_kernel(DevProblem P, SampleArgs A, int warp_bytes) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const WarpMem m = carve(smem + (size_t)warp * warp_bytes, P.kmax, P.pvmax);
    unsigned long long ev_total = 0, pi_total = 0;
    for (;;) {
        unsigned long long h = 0;
        if (lane == 0) h = atomicAdd(A.counters, 1ull);
        h = __shfl_sync(kFull, h, 0);
        if (h >= A.n) break;
        run_history(P, A, m, h, lane, ev_total, pi_total);
        __syncwarp();
    }
    if (lane == 0) { atomicAdd(A.counters + 1, ev_total); atomicAdd(A.counters + 2, pi_total); }
}

// K2: batched pi_SMC from class vectors (parity seam for ConditionalSamplingHMM::piSMC).
// One group of 16 lanes per query; classes are staged into the group's shared-memory column.
__global__ void __launch_bounds__(128) pismc_classes_kernel(DevProblem P, long long nq, const unsigned char* __restrict__ gamma,
                                                            const int* __restrict__ n_all, const int* __restrict__ tidx,
                                                            const long long* __restrict__ offs, const int* __restrict__ cnt,
                                                            const int* __restrict__ dA, const int* __restrict__ dB,
                                                            double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int* hist = reinterpret_cast<unsigned int*>(smem) + (size_t)warp * (P.Ltot + 2) * 32;
    const int p = 16, r = lane & 15;
    const long long qbase = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * 2;
    for (long long q0 = qbase; q0 < nq; q0 += (long long)gridDim.x * (blockDim.x >> 5) * 2) {
        const long long qi = q0 + (lane >> 4);
        const bool valid = qi < nq;
        const long long lo = valid ? offs[qi] : 0, hi = valid ? offs[qi + 1] : 0;
        const int nH = (int)(hi - lo);
        const int g = valid ? gamma[qi] : GC;
        const unsigned int n = valid ? (unsigned int)n_all[qi] : 1u;
        const int set = valid ? P.set_of_epoch[tidx[qi]] : 0;
        const double* row = row_ptr(P, set, (int)n);
        unsigned int* col = hist + lane;
        ClassTotals tot; tot.n = 0; tot.ncompA = 0; tot.ncompB = 0;
        double num = 0.0; unsigned int ncomp = 0;
        for (int k = 0; k < nH; k++) {
            const int c = cnt[lo + k], a = dA[lo + k], b = dB[lo + k];
            col[k * 32] = pack_class((unsigned int)c, a, b);
            tot.n += (unsigned int)c;
            if (a >= 0) tot.ncompA += (unsigned int)c;
            if (b >= 0) tot.ncompB += (unsigned int)c;
            if (g == GA && a >= 0) { num += (double)c * __ldg(row + P.off_wEA + a); ncomp += (unsigned int)c; }
            if (g == GB && b >= 0) { num += (double)c * __ldg(row + P.off_wEB + b); ncomp += (unsigned int)c; }
        }
        double pi;
        {
            ChunkPartial cp = {0, 0, 0, 0, 0, 0};
            auto ld = [row](int off) { return __ldg(row + off); };
            if (g == GC && nH > 0) cp = chunk_forward(ld, c_w, P.off_yw, P.off_zlow, P.off_g, P.off_zdiag, P.off_EA, P.off_EB, P.LA, P.LB,
                                           col, 32, nH, tot, P.pow2negLA, P.pow2negLB, r, r + 1);
            double ipre = cp.pre, ipreB = cp.preB;
            for (int o = 1; o < p; o <<= 1) {
                const double t1 = __shfl_up_sync(kFull, ipre, o, p), t2 = __shfl_up_sync(kFull, ipreB, o, p);
                if (r >= o) { ipre += t1; ipreB += t2; }
            }
            double epre = __shfl_up_sync(kFull, ipre, 1, p), epreB = __shfl_up_sync(kFull, ipreB, 1, p);
            if (r == 0) { epre = 0.0; epreB = 0.0; }
            double acc = chunk_finish(cp, epre, epreB), acc2 = cp.acc2;
            for (int o = 1; o < p; o <<= 1) { acc += __shfl_xor_sync(kFull, acc, o); acc2 += __shfl_xor_sync(kFull, acc2, o); }
            if (g == GC) pi = nH > 0 ? pismc_finish(acc, acc2, tot.n) : P.pow2negL;
            else pi = marginal_finish(num, ncomp, nH > 0 ? n : 0u, g == GA ? P.pow2negLA : P.pow2negLB, P.wsum);
        }
        if (valid && r == 0) out[qi] = pi;
        __syncwarp();
    }
}

// K3: per-block partial log-sum-exp of the (3+T) weighted statistics of SamplerOutput::addSample
// (SamplerOutput.cpp:76-93). grid = (blocks over histories, G, 3+T); block partials [stat][g][block][2].
__global__ void __launch_bounds__(256) lse_partial_kernel(unsigned long long n, int G, int T, const double* __restrict__ weights,
                                                          const double* __restrict__ tmrca, const double* __restrict__ lineages,
                                                          double* __restrict__ block_partials) {
    const int g = blockIdx.y, stat = blockIdx.z;
    double mx = -INFINITY, sm = 0.0;
    for (unsigned long long h = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; h < n; h += (unsigned long long)gridDim.x * blockDim.x) {
        const double w = weights[h * (unsigned long long)G + g];
        double x;
        if (stat == 0) x = w;
        else if (stat == 1) x = 2.0 * w;
        else if (stat == 2) x = w + log(tmrca[h]);
        else x = w + log(lineages[h * (unsigned long long)T + (stat - 3)]);
        if (x == x) {   // histories that overflowed carry NaN and are counted separately
            if (x > mx) { sm = sm * exp(mx - x) + 1.0; mx = x; } else { sm += exp(x - mx); }
        }
    }
    __shared__ double smx[256], ssm[256];
    smx[threadIdx.x] = mx; ssm[threadIdx.x] = sm;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const double m1 = smx[threadIdx.x], m2 = smx[threadIdx.x + o];
            const double s1 = ssm[threadIdx.x], s2 = ssm[threadIdx.x + o];
            const double mm = fmax(m1, m2);
            double ss = 0.0;
            if (mm > -INFINITY) ss = s1 * exp(m1 - mm) + s2 * exp(m2 - mm);
            smx[threadIdx.x] = mm; ssm[threadIdx.x] = ss;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double* o = block_partials + (((size_t)stat * G + g) * gridDim.x + blockIdx.x) * 2;
        o[0] = smx[0]; o[1] = ssm[0];
    }
}

// K3b: combine block partials -> partials[stat][g][2]
__global__ void lse_combine_kernel(int nblocks, int total, const double* __restrict__ block_partials, double* __restrict__ partials) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double* p = block_partials + (size_t)i * nblocks * 2;
    double mx = -INFINITY;
    for (int b = 0; b < nblocks; b++) mx = fmax(mx, p[2 * b]);
    double sm = 0.0;
    if (mx > -INFINITY) for (int b = 0; b < nblocks; b++) sm += p[2 * b + 1] * exp(p[2 * b] - mx);
    partials[2 * i] = mx; partials[2 * i + 1] = sm;
}

size_t history_warp_bytes(int kmax, int pvmax, int Ltot) { return warp_mem_bytes(kmax, pvmax, Ltot); }

}  // namespace coalgpu
