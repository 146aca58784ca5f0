// C ABI of libcoalgpu.so (declared in include/coalgpu.h). Host orchestration only: validate the
// problem, build the HMM tables (coal_tables.cpp), upload, launch the kernels of coal_kernels.cu.
// CUDA-only by design: no CPU fallback exists behind these entry points.
#include "coal_kernels.cu"
#include "coal_tables.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <memory>
#include <string>
#include <thread>
#include <vector>

using namespace coalgpu;

namespace {

constexpr int kCounters = 16;     // device counters per sampler (SampleArgs::counters)
thread_local std::string g_err;
std::atomic<int> g_emission_mode{COALGPU_EMISSION_REFERENCE};   // API value (coalgpu_set_emission_mode)
std::atomic<unsigned long long> g_launches{0};
std::atomic<unsigned long long> g_h2d_bytes{0}, g_d2h_bytes{0};      // bytes the drop-in entry points copied (coalgpu_transfer_bytes)

int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU_OK(expr)                                                                              \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            return fail(_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver ? COALGPU_ENODEV : COALGPU_ECUDA, \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
        }                                                                                        \
    } while (0)

// Device memory comes from the stream-ordered pool with an unlimited release threshold: cudaMalloc/cudaFree
// cost up to hundreds of milliseconds per run_sampler call on this driver (measured), the pool costs microseconds.
cudaError_t dev_alloc(void** p, size_t bytes, cudaStream_t stream = 0) {
    static std::atomic<unsigned long long> configured{0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(configured.load() & bit)) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        configured.fetch_or(bit);
    }
    return cudaMallocAsync(p, std::max<size_t>(bytes, 16), stream);
}
// Stream-ordered free: the release must be ordered after the last kernel that touches the buffer, i.e. issued on the
// stream that ran it (the legacy default stream does not order against cudaStreamNonBlocking streams).
void dev_free(void* p, cudaStream_t stream = 0) { if (p) cudaFreeAsync(p, stream); }

// The dynamic shared memory limit is a property of the kernel function, not of a launch: raise it to the device's
// opt-in maximum once per device for every instantiation. (Changing it per launch serialises against the launches of
// the same function that are still running -- measured: it made the pairs of a batch run one after another.)
int ensure_kernel_attributes(int device) {
    static std::atomic<unsigned long long> done{0};
    const unsigned long long bit = 1ull << (device & 63);
    if (done.load() & bit) return 0;
    int optin = 0;
    CU_OK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    // tuning knob: shared-memory carve-out of the lockstep kernels in percent (the rest of the 256 KB is L1); default: the driver's choice
    int carve = -1;
    if (const char* c = std::getenv("COALGPU_CARVEOUT")) carve = std::max(0, std::min(100, std::atoi(c)));
    for (int W : {8, 16, 32}) for (int narrow = 0; narrow < 2; narrow++) {
        const void* fn = (const void*)history_kernel_select(W, narrow != 0);
        cudaFuncAttributes fa;
        CU_OK(cudaFuncGetAttributes(&fa, fn));            // static shared memory (problem constants) counts against the limit
        CU_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
    }
    for (int shape : {8 | (32 << 8), 16 | (64 << 8), 16 | (128 << 8), 32 | (64 << 8), 32 | (128 << 8), 16 | (64 << 8) | (1 << 16)}) for (int narrow = 0; narrow < 2; narrow++) {
        const void* fn = (const void*)history_lockstep_select(shape & 0xff, (shape >> 8) & 0xff, narrow != 0, (shape >> 16) != 0);
        cudaFuncAttributes fa;
        CU_OK(cudaFuncGetAttributes(&fa, fn));
        CU_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
        if (carve >= 0) CU_OK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    }
    done.fetch_or(bit);
    return 0;
}

}  // namespace

struct coalgpu_sampler {
    int device = 0;
    int T = 0, LA = 0, LB = 0, G = 0;
    DevProblem P{};
    HostTables tab;
    std::vector<void*> allocs;
    unsigned long long* d_counters = nullptr;
    double* d_block_partials = nullptr;
    size_t block_partials_cap = 0;
    double* d_full_rows = nullptr;    // reference-layout rows on the device (device table builder only; owned through allocs)
    cudaStream_t last_stream = 0;     // stream of the most recent launch of this sampler: its buffers are freed in that stream's order
    // lockstep form of K1 (coal_lockstep.cuh): `group` then holds NH, the histories per CTA (16 or 32), CTAs have 4 * NH threads
    bool ls = false, fast_ls = false, wide_ls = false;
    int warp_bytes = 0, grid = 0, group = 32;   // warp_bytes = shared memory per history in flight; group = lanes per history of the full-store geometry (16 or 32; the reduced-store first pass may use 8)
    size_t smem_bytes = 0;
    // reduced-capacity first pass (see finish_device): same problem with a smaller type store
    int inj_max = 0, sm_count = 0;    // most input types of one sampling time; SMs of the device
    int type_estimate = 0;            // HostPrep::type_estimate
    bool rows_on_device_only = false; // emission parts were built on the device: coalgpu_tables fetches the rows first
    bool pilot_useful = false;        // choose_geometry: a measured peak could still raise the histories per SM
    bool fast = false, fast_forced = false;   // forced: COALGPU_KMAX_FAST > 0 (tests), two passes whatever the launch size
    DevProblem P_fast{};
    int fast_grid = 0, fast_group = 32;
    size_t fast_smem = 0;
    // the same without W = 8 (for launches that run side by side with other pairs' launches and cannot fill the
    // larger capacity of W = 8 on their own: coalgpu_run_sampler_batch with few samples per pair)
    bool wide = false;
    DevProblem P_wide{};
    int wide_grid = 0, wide_group = 32;
    size_t wide_smem = 0;
    unsigned long long tot_events = 0, tot_pismc = 0, tot_overflow = 0;
    unsigned long long n_nan_seen = 0;   // NaN log-weights the reduction kernel skipped (counters[9]) as of the last coalgpu_counters
    // K2 scratch smem
    size_t k2_smem = 0;
};

namespace {

// Everything coalgpu_create computes on the CPU (no CUDA call): the batch entry point runs this part for the
// next locus pairs on host threads while the GPU works on the current ones.
struct HostPrep {
    int T = 0, LA = 0, LB = 0, G = 0;
    int type_estimate = 0;    // distinct two-locus haplotypes + distinct A parts + distinct B parts of the input (see choose_geometry)
    bool device_emission = false;     // emission parts of the rows are built by emission_stable_kernel after the upload
    std::vector<double> thetas; int Q = 16;
    DevProblem P{};
    HostTables tab;
    std::vector<int> type_off;
    std::vector<unsigned long long> type_w;
    std::vector<unsigned int> type_cg;
    std::vector<EpochConst> ep;
    std::vector<GridConst> grid;
    std::vector<double> lfact, rcp, times;
};

int validate_problem(const coalgpu_problem* p, uint32_t Q, std::string& err) {
    auto bad = [&](const char* m) { err = m; return COALGPU_EINVAL; };
    if (Q != 4 && Q != 8 && Q != 16) return bad("quadrature_points must be 4, 8 or 16 (ConditionalSamplingHMM.cpp:39-58)");
    if (p->T < 1 || p->LA < 1 || p->LB < 1 || p->LA + p->LB > 64) return bad("need T >= 1, LA, LB >= 1 and LA + LB <= 64");
    if (p->n_mu < 1 || p->n_rho < 1 || p->n_lambda < 1) return bad("empty target grid");
    if (!(p->tau > 0) || !(p->driving_mu > 0) || !(p->driving_lambda > 0) || !(p->driving_rho > 0)) return bad("driving values and tau must be positive (main.cpp:147-150)");
    for (int t = 1; t < p->T; t++) if (!(p->times[t] > p->times[t - 1])) return bad("sampling times must be strictly ascending");
    if (!p->times || !p->viral || !p->mu || !p->rho || !p->lambda || !p->type_offsets) return bad("null array in the problem");
    for (int t = 0; t < p->T; t++) if (!(p->viral[t] > 0)) return bad("viral loads must be positive");
    // the target grid as main.cpp:165-182 asserts it: mu >= 0, rho >= 0, lambda > 0 (NaN fails every comparison)
    for (int i = 0; i < p->n_mu; i++) if (!(p->mu[i] >= 0) || std::isinf(p->mu[i])) return bad("target mutation rates must be finite and >= 0 (main.cpp:169)");
    for (int i = 0; i < p->n_rho; i++) if (!(p->rho[i] >= 0) || std::isinf(p->rho[i])) return bad("target recombination rates must be finite and >= 0 (main.cpp:181)");
    for (int i = 0; i < p->n_lambda; i++) if (!(p->lambda[i] > 0) || std::isinf(p->lambda[i])) return bad("target scale values must be finite and > 0 (main.cpp:175)");
    if (p->type_offsets[0] < 0) return bad("type_offsets must start at a non-negative index");
    for (int t = 0; t < p->T; t++) if (p->type_offsets[t + 1] < p->type_offsets[t]) return bad("type_offsets must be non-decreasing");
    if (p->type_offsets[p->T] > p->type_offsets[0] && (!p->type_gamma || !p->type_words || !p->type_counts)) return bad("null type arrays");
    return 0;
}

int device_check(int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(COALGPU_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(COALGPU_ENODEV, "device index out of range");
    return 0;
}

int prepare_host(const coalgpu_problem* p, uint32_t Q, int table_threads, HostPrep& h, std::string& err) {
    auto bad = [&](const char* m) { err = m; return COALGPU_EINVAL; };
    if (int rc = validate_problem(p, Q, err)) return rc;
    h.T = p->T; h.LA = p->LA; h.LB = p->LB; h.G = p->n_mu * p->n_rho * p->n_lambda;
    const int T = p->T, LA = p->LA, LB = p->LB, G = h.G;
    const uint64_t maskA = (LA == 64) ? ~0ull : ((1ull << LA) - 1);
    const uint64_t maskB = ((LB == 64) ? ~0ull : ((1ull << LB) - 1)) << LA;

    // merge duplicate input types per sampling time, keep insertion order; bound the lineage count:
    // every sampled two-locus lineage can split once into an A and a B lineage
    std::vector<int>& type_off = h.type_off; type_off.assign(T + 1, 0);
    std::vector<unsigned long long>& type_w = h.type_w;
    std::vector<unsigned int>& type_cg = h.type_cg;
    long long n_bound = 0;
    double perm = 0.0;
    for (int t = 0; t < T; t++) {
        const size_t start = type_w.size();
        for (int k = p->type_offsets[t]; k < p->type_offsets[t + 1]; k++) {
            const int g = p->type_gamma[k];
            const uint64_t w = p->type_words[k];
            const unsigned int c = p->type_counts[k];
            if (g > 2 || c == 0) return bad("bad type record");
            const uint64_t allowed = g == COALGPU_GAMMA_A ? maskA : (g == COALGPU_GAMMA_B ? maskB : (maskA | maskB));
            if (w & ~allowed) return bad("haplotype word has bits outside its loci");
            bool hit = false;
            for (size_t i = start; i < type_w.size(); i++)
                if (type_w[i] == w && (int)(type_cg[i] >> 30) == g) { type_cg[i] += c; hit = true; break; }
            if (!hit) { type_w.push_back(w); type_cg.push_back(((unsigned int)g << 30) | c); }
            n_bound += (g == COALGPU_GAMMA_C) ? 2ll * c : c;
        }
        // Canonical input order: the merged types of a sampling time sorted by (class, packed haplotype). The slot order of a
        // history's type store -- hence which history a given (seed, index) draws -- then no longer depends on the order in
        // which the caller happens to list the types (the reference binding lists them in unordered_map iteration order, the
        // host driver in first-appearance order: both now get the same numbers, byte-identical output files).
        {
            std::vector<std::pair<std::pair<int, unsigned long long>, unsigned int>> blk;
            for (size_t i = start; i < type_w.size(); i++) blk.push_back({{(int)(type_cg[i] >> 30), type_w[i]}, type_cg[i]});
            std::sort(blk.begin(), blk.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
            for (size_t i = start; i < type_w.size(); i++) { type_w[i] = blk[i - start].first.second; type_cg[i] = blk[i - start].second; }
        }
        type_off[t + 1] = (int)type_w.size();
        // probabilitySamplePermutation, ProposalEngine.cpp:88-118
        for (int g = 0; g < 3; g++) {
            unsigned int tot = 0;
            for (size_t i = start; i < type_w.size(); i++) if ((int)(type_cg[i] >> 30) == g) { const unsigned int c = type_cg[i] & 0x3fffffffu; tot += c; perm += std::lgamma((double)c + 1.0); }
            perm -= std::lgamma((double)tot + 1.0);
        }
    }
    {   // haplotype diversity of the input: what a history's type store has to hold is close to it (choose_geometry)
        std::vector<unsigned long long> c, a, b;
        for (size_t i = 0; i < type_w.size(); i++) {
            const int g = (int)(type_cg[i] >> 30);
            if (g == COALGPU_GAMMA_C) c.push_back(type_w[i]);
            if (g != COALGPU_GAMMA_B) a.push_back(type_w[i] & maskA);
            if (g != COALGPU_GAMMA_A) b.push_back(type_w[i] & maskB);
        }
        auto distinct = [](std::vector<unsigned long long>& v) { std::sort(v.begin(), v.end()); return (int)(std::unique(v.begin(), v.end()) - v.begin()); };
        h.type_estimate = distinct(c) + distinct(a) + distinct(b);
    }
    if (type_off[1] == 0) return bad("first sampling time has no sequences");
    if (n_bound < 1 || n_bound > 60000) return bad("lineage bound out of range (1..60000)");
    const int n_max = (int)n_bound;

    // HMM constants per (epoch, lineage count): ImportanceSampler.cpp:69-76, ConditionalSamplingHMM.cpp:36-226
    std::vector<double> thetas(T), rrates(T);
    for (int c = 0; c < T; c++) {
        thetas[c] = 4 * p->driving_lambda * p->viral[c] * p->driving_mu;
        rrates[c] = 4 * p->driving_lambda * p->viral[c] * p->driving_rho;
    }
    // emission tables: reference order (host, cached), stable on the host, or left empty for the device kernel (finish_device)
    const int emode = g_emission_mode.load();
    h.device_emission = emode == COALGPU_EMISSION_STABLE;
    h.thetas = thetas; h.Q = (int)Q;
    if (!build_tables((int)Q, LA, LB, n_max, thetas, rrates, h.tab, table_threads,
                      emode == COALGPU_EMISSION_STABLE ? kEmissionNone : emode == COALGPU_EMISSION_STABLE_HOST ? kEmissionStable : kEmissionReference))
        return bad("table construction failed");

    h.ep.resize(T);
    for (int c = 0; c < T; c++) {
        const double V = p->viral[c];
        EpochConst e{};
        e.dN = p->driving_lambda * V;                 // ProposalEngine.cpp:657-663
        e.dtheta = 4 * p->driving_mu * e.dN;
        e.drho = 4 * p->driving_rho * e.dN;
        e.dthA = e.dtheta * LA; e.dthB = e.dtheta * LB;
        e.inv2Ntau = 1.0 / (2 * e.dN * p->tau);
        h.ep[c] = e;
    }
    h.grid.resize((size_t)T * G);
    for (int c = 0; c < T; c++) {
        const double V = p->viral[c];
        size_t g = 0;
        for (int i = 0; i < p->n_mu; i++) for (int j = 0; j < p->n_rho; j++) for (int z = 0; z < p->n_lambda; z++, g++) {
            GridConst gc;
            const double N = p->lambda[z] * V;        // ProposalEngine.cpp:462-464
            gc.theta = 4 * p->mu[i] * p->lambda[z] * V;
            gc.rho = 4 * p->rho[j] * p->lambda[z] * V;
            gc.inv2Ntau = 1.0 / (2 * N * p->tau);
            gc.log_theta = std::log(gc.theta); gc.log_rho = std::log(gc.rho); gc.log_2Ntau = std::log(2 * N * p->tau);
            h.grid[(size_t)c * G + g] = gc;
        }
    }
    h.lfact.resize(n_max + 2);
    for (int i = 0; i < n_max + 2; i++) h.lfact[i] = std::lgamma((double)i + 1.0);
    h.rcp.assign(n_max + 2, 0.0);
    for (int i = 1; i < n_max + 2; i++) h.rcp[i] = 1.0 / (double)i;
    h.times.assign(p->times, p->times + T);

    DevProblem& P = h.P;
    const TableLayout& lay = h.tab.lay;
    P.T = T; P.LA = LA; P.LB = LB; P.Ltot = LA + LB; P.G = G;
    P.n_max = n_max;
    P.kmax = ((n_max + 2 + 31) / 32) * 32;
    P.pvmax = std::max(LA + LB + 4, std::max(LA, LB) + P.kmax + 1);
    P.pvmax = (P.pvmax + 1) & ~1;
    P.row_doubles = lay.dev_row_doubles;
    P.SB = lay.dev_sb; P.off_wEA = lay.dev_off_wEA; P.off_wEB = lay.dev_off_wEB;
#ifndef COAL_TMA_ROWS
    P.stage_doubles = 0;                                            // default build: every table row is read through L1 (ld.global.nc)
#else
    P.stage_doubles = (LA + LB <= 32) ? lay.dev_row_doubles : 0;    // matches kStaged in history_kernel: longer loci read their rows through L1
#endif
    P.maskA = maskA; P.maskB = maskB;
    P.perm_const = perm;
    P.pow2negLA = std::ldexp(1.0, -LA); P.pow2negLB = std::ldexp(1.0, -LB); P.pow2negL = std::ldexp(1.0, -(LA + LB));
    double wsum = 0; for (int i = 0; i < kQ; i++) wsum += h.tab.w[i];
    P.wsum = wsum;
    return 0;
}

// Launch geometry: persistent CTAs of cta_threads(W) threads = cta_threads(W)/W groups of W lanes, one history per group.
// W = 32, 16 or 8 lanes per history. More histories in flight per SM is better, but narrower groups take longer over
// the loops that the lanes of a group share (look-ups, draws, lineage weights: ceil(slots / W) steps). Measured per
// pair on cfg1/cfg2/cfg4 with tuned stores (profiles/r1_kernel_iterations.md): W = 8 (40-48 histories per SM) over
// W = 16 (24 per SM) is 1.24-1.42x when histories hold at most ~26 distinct types, 1.03-1.11x at 32-40, 1.00x at 51
// and 0.93x at 61. So W = 8 is used only once the sampler knows what its histories need (coalgpu_tune) and that is at
// most 44 types, and only if it puts at least 1.3x the histories on an SM. W = 16 needs a batch of two-locus queries
// (Ltot + 1) to fit 16 lanes.
//
// Two-pass capacity scheme. kmax is the worst case (every sampled two-locus lineage split by recombination and all
// types distinct); real histories hold far fewer distinct types. When a smaller type store lets more histories run
// per SM, the first pass uses it and defers the rare history that outgrows it to a second pass with the full store;
// a history's draws depend on (seed, index) only, so the results do not depend on the scheme. `peak` > 0 (coalgpu_tune)
// is the largest store any history of this sampler has needed so far: the reduced store is then sized from it.
struct Geometry { int group = 0, occ = 0, per_sm = 0; size_t smem = 0; bool ls = false; };
// threads per CTA / histories per CTA of a geometry: the group form packs cta_threads(W) / W groups of W lanes, the
// lockstep form advances NH = `group` histories with 4 * NH threads
// (lockstep: `group` = NH | NT << 8 | gstore << 16; gstore = the capacity tier with the type store in global memory)
inline int geo_threads(int group, bool ls) { return ls ? ((group >> 8) & 0xff) : cta_threads(group); }
inline bool geo_gstore(int group, bool ls) { return ls && (group >> 16) != 0; }
inline history_kernel_fn geo_kernel(int group, bool ls, bool narrow) {
    return ls ? history_lockstep_select(group & 0xff, (group >> 8) & 0xff, narrow, (group >> 16) != 0) : history_kernel_select(group, narrow);
}
// bytes of the global-memory type store one CTA of a capacity-tier launch needs (0 for every other geometry)
inline size_t geo_store_bytes(const DevProblem& P, int group, bool ls) {
    return geo_gstore(group, ls) ? history_lockstep_store_bytes(group & 0xff, (group >> 8) & 0xff, P.kmax, P.pvmax, P.Ltot, P.stage_doubles, P.Ltot <= 32) : 0;
}
inline int geo_hist_per_cta(int group, bool ls) { return ls ? (group & 0xff) : cta_threads(group) / group; }
// Which form of K1: the lockstep form (default) or the group form (COALGPU_KERNEL=group, or a forced group width
// COALGPU_GROUP=8|16|32, which the parity tests of the group form use).
bool want_lockstep() {
    if (const char* g = std::getenv("COALGPU_GROUP")) { const int W = std::atoi(g); if (W == 8 || W == 16 || W == 32) return false; }
    if (const char* k = std::getenv("COALGPU_KERNEL")) return std::string(k) != "group";
    return true;
}
Geometry geometry_for(const DevProblem& P, int kmax, int pvmax, bool allow8, bool gstore = false) {
    if (want_lockstep()) {
        int forced_nh = 0, forced_nt = 0;
        if (const char* g = std::getenv("COALGPU_LOCKSTEP_NH")) { const int v = std::atoi(g); if (v == 8 || v == 16 || v == 32) forced_nh = v; }
        if (const char* g = std::getenv("COALGPU_LOCKSTEP_NT")) { const int v = std::atoi(g); if (v == 64 || v == 128) forced_nt = v; }
        auto eval = [&](int NH, int NT, bool gstore) {
            Geometry g; g.group = NH | (NT << 8) | (gstore ? 1 << 16 : 0); g.ls = true;
            g.smem = history_lockstep_bytes(NH, NT, kmax, pvmax, P.Ltot, P.stage_doubles, P.Ltot <= 32, gstore);
            if (g.smem > 227 * 1024) return Geometry();
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.occ, (const void*)history_lockstep_select(NH, NT, P.Ltot <= 32, gstore), NT, g.smem) != cudaSuccess) { cudaGetLastError(); return Geometry(); }
            g.per_sm = g.occ * NH;
            return g;
        };
        if (gstore) return eval(16, 64, true);
        if (forced_nh) return eval(forced_nh, forced_nh == 8 ? 32 : (forced_nt ? forced_nt : (forced_nh == 16 ? 64 : 128)), false);
        const Geometry g16 = eval(16, 64, false), g32 = eval(32, forced_nt ? forced_nt : 128, false);
        if (g16.per_sm > 0 || g32.per_sm > 0) return g32.per_sm > g16.per_sm ? g32 : g16;
        // neither fits (type store beyond ~200 KB of shared memory per CTA): fall through to the group form
    }
    if (gstore) return Geometry();
    int forced_w = 0;
    if (const char* g = std::getenv("COALGPU_GROUP")) { const int W = std::atoi(g); if (W == 8 || W == 16 || W == 32) forced_w = W; }
    auto eval = [&](int W) {
        Geometry g; g.group = W;
        g.smem = history_group_bytes(kmax, pvmax, P.Ltot, W, P.stage_doubles) * (cta_threads(W) / W);
        if (g.smem > 227 * 1024) return Geometry();
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.occ, (const void*)history_kernel_select(W, P.Ltot <= 32), cta_threads(W), g.smem) != cudaSuccess) { cudaGetLastError(); return Geometry(); }
        g.per_sm = g.occ * (cta_threads(W) / W);
        return g;
    };
    if (forced_w) return eval(forced_w);
    Geometry best = eval(32);
    { const Geometry g16 = eval(16); if (g16.per_sm > best.per_sm) best = g16; }
    if (allow8) { const Geometry g8 = eval(8); if (10 * g8.per_sm >= 13 * best.per_sm) best = g8; }
    return best;
}

int choose_geometry(coalgpu_sampler* s, int peak) {
    DevProblem& P = s->P;
    Geometry full = geometry_for(P, P.kmax, P.pvmax, false);
    // Capacity tier: a worst-case type store beyond one CTA's shared memory (about 1,500 slots) goes to global memory. It is the
    // store of last resort -- the reduced-store first pass below still runs out of shared memory whenever one fits, and only the
    // histories that outgrow that are drawn here (COALGPU_GSTORE=1 forces the tier: parity tests).
    const char* gs_env = std::getenv("COALGPU_GSTORE");
    const bool gs = full.per_sm < 1 || (gs_env && std::atoi(gs_env) > 0);
    if (gs) full = geometry_for(P, P.kmax, P.pvmax, false, true);
    if (full.per_sm < 1) return fail(COALGPU_EINVAL, "problem too large: the per-event state of one CTA does not fit shared memory (lockstep form required: unset COALGPU_KERNEL / COALGPU_GROUP)");
    s->group = full.group; s->ls = full.ls; s->smem_bytes = full.smem; s->grid = s->sm_count * full.occ;
    s->warp_bytes = full.ls ? 0 : (int)history_group_bytes(P.kmax, P.pvmax, P.Ltot, s->group, P.stage_doubles);
    P.warp_bytes = s->warp_bytes;
    s->fast = false; s->fast_forced = false; s->wide = false;

    auto pv_for = [&](int k) { int pv = std::max(P.Ltot + 4, std::max(P.LA, P.LB) + k + 1); return (pv + 1) & ~1; };
    int forced = 0;
    if (const char* e = std::getenv("COALGPU_KMAX_FAST")) forced = std::atoi(e);
    if (forced < 0) return 0;                                       // single pass
    std::vector<int> cand; bool allow8 = false;
    if (forced > 0) cand.push_back(((forced + 7) / 8) * 8);          // tests force small stores so that histories ARE deferred
    else {
        // What the histories need: measured (coalgpu_tune: the most distinct types any history held so far) or, before
        // anything was drawn, estimated from the input: distinct two-locus haplotypes + distinct A parts + distinct
        // B parts, + 4. On cfg1/cfg2/cfg4 pairs the measured peak is between 0.64x and 1.1x of that sum and never
        // more than 2 above it (profiles/r1_kernel_iterations.md). The store is 1/8 (at least 4) above, in steps of 8:
        // a history that needs more is deferred to the full store, not lost.
        const int need = peak > 0 ? peak : s->type_estimate + 4;
        cand.push_back(std::max(32, ((need + std::max(4, need / 8) + 7) / 8) * 8));
        allow8 = need <= 44;
        if (peak == 0 && want_lockstep()) {
            // An ESTIMATED need is an upper bound more often than not (measured peaks are 0.64-1.1x of it). When the store sized
            // for it costs a CTA per SM against the smallest store, take the largest store that does not -- as long as that
            // still covers three quarters of the estimate; a history that needs more is deferred to the full store, not lost.
            const int kmin = std::min(32, P.kmax);
            const int best = geometry_for(P, kmin, pv_for(kmin), false).per_sm;
            int k = std::min(cand[0], P.kmax);
            while (k > kmin && geometry_for(P, k, pv_for(k), false).per_sm < best) k -= 8;
            if (k < cand[0] && 4 * k >= 3 * need) cand[0] = k;
        }
        for (int kf : {256, 128, 64}) if (kf > cand[0] && kf >= 2 * s->inj_max + 8) cand.push_back(kf);   // larger stores that still raise the occupancy
    }
    auto pick = [&](bool with8, DevProblem& Pout, int& grid, int& group, size_t& smem, bool& ls) -> bool {
        Geometry bestf = full; int best_k = 0;
        if (gs) bestf.per_sm = 0;        // any store that fits shared memory beats the capacity tier
        for (int kf : cand) {
            kf = std::min(kf, P.kmax);                               // kf == kmax: same store, possibly narrower groups
            const Geometry g = geometry_for(P, kf, pv_for(kf), with8);
            if (g.per_sm > bestf.per_sm || (forced > 0 && kf < P.kmax && g.per_sm > 0)) { bestf = g; best_k = kf; }
        }
        if (best_k == 0) return false;
        Pout = P;
        Pout.kmax = best_k;
        Pout.pvmax = pv_for(best_k);
        group = bestf.group; smem = bestf.smem; grid = s->sm_count * bestf.occ; ls = bestf.ls;
        Pout.warp_bytes = bestf.ls ? 0 : (int)history_group_bytes(best_k, Pout.pvmax, P.Ltot, bestf.group, P.stage_doubles);
        return true;
    };
    s->fast = pick(allow8, s->P_fast, s->fast_grid, s->fast_group, s->fast_smem, s->fast_ls);
    s->fast_forced = s->fast && forced > 0;
    s->wide = pick(false, s->P_wide, s->wide_grid, s->wide_group, s->wide_smem, s->wide_ls);
    // Would knowing the histories' real need (coalgpu_tune after a pilot chunk) put more histories on an SM than the
    // input-based estimate does? Only then is a pilot worth one history's latency (coalgpu_run_sampler). With the lockstep
    // form the register file caps the CTAs per SM long before a 32-48 slot store does, so usually it is not.
    {
        const int kmin = std::min(32, P.kmax);
        const Geometry gmin = geometry_for(P, kmin, pv_for(kmin), true);
        const int now = s->fast ? (s->fast_grid / std::max(1, s->sm_count)) * geo_hist_per_cta(s->fast_group, s->fast_ls) : full.per_sm;
        s->pilot_useful = peak == 0 && forced == 0 && gmin.per_sm > now;
    }
    return 0;
}

// Device half of coalgpu_create: checks the device, uploads what prepare_host built, picks the launch geometry.
// Everything goes into ONE stream-ordered allocation filled by asynchronous copies on `up`; nothing here waits for
// kernels of other samplers. The caller orders its launches after `up` (stream wait or synchronize).
int finish_device(HostPrep& h, int device, cudaStream_t up, coalgpu_sampler** out) {
    if (int rc = device_check(device)) return rc;
    CU_OK(cudaSetDevice(device));
    // cudaGetDeviceProperties costs tens of milliseconds per call; two attributes are all that is needed
    int cc_major = 0, sm_count = 0;
    CU_OK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
    CU_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    if (cc_major != 10) return fail(COALGPU_ENODEV, "this library carries sm_100a code only (Blackwell B200 required)");
    if (int rc = ensure_kernel_attributes(device)) return rc;

    coalgpu_sampler* s = new coalgpu_sampler;
    s->device = device; s->T = h.T; s->LA = h.LA; s->LB = h.LB; s->G = h.G;
    s->P = h.P;
    s->tab = std::move(h.tab);
    DevProblem& P = s->P;

    // arena layout (every part 256-byte aligned: the table rows are the source of bulk copies)
    size_t off = 0;
    auto place = [&off](size_t bytes) { const size_t o = off; off += (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255; return o; };
    const size_t dev_rows_doubles = (size_t)s->tab.n_sets * s->tab.n_max * s->tab.lay.dev_row_doubles;
    const size_t o_rows = place(dev_rows_doubles * sizeof(double)), o_set = place(s->tab.set_of_epoch.size() * sizeof(int));
    const size_t o_times = place(h.times.size() * sizeof(double)), o_ep = place(h.ep.size() * sizeof(EpochConst));
    const size_t o_grid = place(h.grid.size() * sizeof(GridConst)), o_lf = place(h.lfact.size() * sizeof(double));
    const size_t o_rcp = place(h.rcp.size() * sizeof(double)), o_toff = place(h.type_off.size() * sizeof(int));
    const size_t o_tw = place(h.type_w.size() * sizeof(unsigned long long)), o_tcg = place(h.type_cg.size() * sizeof(unsigned int));
    const size_t o_cnt = place(kCounters * sizeof(unsigned long long));
    char* base = nullptr;
    {
        cudaError_t e = dev_alloc((void**)&base, off, up);
        if (e != cudaSuccess) { delete s; return fail(e == cudaErrorMemoryAllocation ? COALGPU_ENOMEM : COALGPU_ECUDA, std::string("device arena: ") + cudaGetErrorString(e)); }
    }
    s->allocs.push_back(base);
    cudaError_t ue = cudaSuccess;
    auto put = [&](size_t o, const void* src, size_t bytes) { if (bytes && ue == cudaSuccess) { ue = cudaMemcpyAsync(base + o, src, bytes, cudaMemcpyHostToDevice, up); g_h2d_bytes += bytes; } };
    if (!h.device_emission) put(o_rows, s->tab.dev_rows.data(), dev_rows_doubles * sizeof(double));
    put(o_set, s->tab.set_of_epoch.data(), s->tab.set_of_epoch.size() * sizeof(int));
    put(o_times, h.times.data(), h.times.size() * sizeof(double));
    put(o_ep, h.ep.data(), h.ep.size() * sizeof(EpochConst));
    put(o_grid, h.grid.data(), h.grid.size() * sizeof(GridConst));
    put(o_lf, h.lfact.data(), h.lfact.size() * sizeof(double));
    put(o_rcp, h.rcp.data(), h.rcp.size() * sizeof(double));
    put(o_toff, h.type_off.data(), h.type_off.size() * sizeof(int));
    put(o_tw, h.type_w.data(), h.type_w.size() * sizeof(unsigned long long));
    put(o_tcg, h.type_cg.data(), h.type_cg.size() * sizeof(unsigned int));
    if (ue == cudaSuccess) ue = cudaMemsetAsync(base + o_cnt, 0, kCounters * sizeof(unsigned long long), up);
    if (ue == cudaSuccess && h.device_emission) {
        // Stable emission mode, table builder on the device: the reference-layout rows arrive with the transition constants
        // only (host), K4 fills their EA | EB | wEA | wEB from the table sets' thetas, K5 turns them into the rows the
        // history kernel reads. The reference-layout copy stays on the device for coalgpu_tables.
        std::vector<double> set_theta(s->tab.n_sets, 0.0);
        for (int c = 0; c < h.T; c++) set_theta[s->tab.set_of_epoch[c]] = h.thetas[c];
        double* d_theta = nullptr;
        ue = dev_alloc((void**)&d_theta, set_theta.size() * sizeof(double), up);
        if (ue == cudaSuccess) { s->allocs.push_back(d_theta); ue = cudaMemcpyAsync(d_theta, set_theta.data(), set_theta.size() * sizeof(double), cudaMemcpyHostToDevice, up); }
        if (ue == cudaSuccess) ue = dev_alloc((void**)&s->d_full_rows, s->tab.rows.size() * sizeof(double), up);
        if (ue == cudaSuccess) { s->allocs.push_back(s->d_full_rows); ue = cudaMemcpyAsync(s->d_full_rows, s->tab.rows.data(), s->tab.rows.size() * sizeof(double), cudaMemcpyHostToDevice, up); }
        if (ue == cudaSuccess) {
            const TableLayout& lay = s->tab.lay;
            EmissionArgs ea{};
            ea.rows = s->d_full_rows; ea.thetas = d_theta;
            ea.n_sets = s->tab.n_sets; ea.n_max = s->tab.n_max; ea.row_doubles = lay.row_doubles; ea.LA = h.LA; ea.LB = h.LB; ea.Q = h.Q;
            ea.off_EA = lay.off_EA; ea.off_EB = lay.off_EB; ea.off_wEA = lay.off_wEA; ea.off_wEB = lay.off_wEB;
            for (int i = 0; i <= kQ; i++) ea.t[i] = s->tab.quad[i];
            for (int i = 0; i < kQ; i++) ea.w[i] = s->tab.w[i];
            const long long total = (long long)ea.n_sets * ea.n_max * (h.LA + h.LB + 2);
            const int blocks = (int)std::min<long long>((total + 127) / 128, (long long)sm_count * 16);
            emission_stable_kernel<<<blocks, 128, 0, up>>>(ea);
            g_launches++;
            ue = cudaGetLastError();
            s->rows_on_device_only = true;
        }
        if (ue == cudaSuccess) {
            const TableLayout& lay = s->tab.lay;
            BilinearArgs ba{};
            ba.full_rows = s->d_full_rows; ba.dev_rows = (double*)(base + o_rows);
            ba.n_rows = (long long)s->tab.n_sets * s->tab.n_max;
            ba.row_doubles = lay.row_doubles; ba.dev_row_doubles = lay.dev_row_doubles; ba.LA = h.LA; ba.LB = h.LB;
            ba.off_yw = lay.off_yw; ba.off_zlow = lay.off_zlow; ba.off_g = lay.off_g; ba.off_zdiag = lay.off_zdiag;
            ba.off_EA = lay.off_EA; ba.off_EB = lay.off_EB; ba.off_wEA = lay.off_wEA; ba.off_wEB = lay.off_wEB;
            ba.dev_sb = lay.dev_sb; ba.dev_off_wEA = lay.dev_off_wEA; ba.dev_off_wEB = lay.dev_off_wEB;
            for (int i = 0; i < kQ; i++) ba.w[i] = s->tab.w[i];
            const long long total = ba.n_rows * (h.LA + 2);
            const int blocks = (int)std::min<long long>((total + 127) / 128, (long long)sm_count * 16);
            bilinear_rows_kernel<<<blocks, 128, 0, up>>>(ba);
            g_launches++;
            ue = cudaGetLastError();
        }
    }
    if (ue != cudaSuccess) { coalgpu_destroy(s); return fail(COALGPU_ECUDA, std::string("upload: ") + cudaGetErrorString(ue)); }
    P.rows = (const double*)(base + o_rows); P.set_of_epoch = (const int*)(base + o_set); P.times = (const double*)(base + o_times);
    P.ep = (const EpochConst*)(base + o_ep); P.grid = (const GridConst*)(base + o_grid); P.lfact = (const double*)(base + o_lf);
    P.rcp = (const double*)(base + o_rcp); P.type_off = (const int*)(base + o_toff);
    P.type_w = (const unsigned long long*)(base + o_tw); P.type_cg = (const unsigned int*)(base + o_tcg);
    s->d_counters = (unsigned long long*)(base + o_cnt);

    for (int t = 0; t < h.T; t++) s->inj_max = std::max(s->inj_max, h.type_off[t + 1] - h.type_off[t]);
    s->sm_count = sm_count;
    s->type_estimate = h.type_estimate;
    if (int grc = choose_geometry(s, 0)) { coalgpu_destroy(s); return grc; }
    s->k2_smem = (size_t)4 * (P.Ltot + 2) * 32 * sizeof(unsigned int);
    *out = s;
    return 0;
}

}  // namespace

extern "C" {

const char* coalgpu_last_error(void) { return g_err.c_str(); }
#ifdef COAL_TMA_ROWS
const char* coalgpu_version(void) { return "coalgpu 0.1 (sm_100a, TMA-staged table rows)"; }
#else
const char* coalgpu_version(void) { return "coalgpu 0.1 (sm_100a)"; }
#endif
uint64_t coalgpu_launch_count(void) { return g_launches.load(); }
void coalgpu_transfer_bytes(uint64_t* h2d, uint64_t* d2h) { if (h2d) *h2d = g_h2d_bytes.load(); if (d2h) *d2h = g_d2h_bytes.load(); }

int coalgpu_set_emission_mode(int mode) {
    if (mode != COALGPU_EMISSION_REFERENCE && mode != COALGPU_EMISSION_STABLE && mode != COALGPU_EMISSION_STABLE_HOST) return fail(COALGPU_EINVAL, "unknown emission mode");
    g_emission_mode.store(mode);
    return 0;
}

void coalgpu_destroy(coalgpu_sampler* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    // ordered after the sampler's last launch (the legacy default stream does not order against non-blocking streams)
    for (void* p : s->allocs) dev_free(p, s->last_stream);
    dev_free(s->d_block_partials, s->last_stream);
    delete s;
}

static int create_impl(const coalgpu_problem* p, uint32_t Q, int device, coalgpu_sampler** out) {
    if (!p || !out) return fail(COALGPU_EINVAL, "null argument");
    *out = nullptr;
    HostPrep h; std::string err;
    int rc = validate_problem(p, Q, err);
    if (rc) return fail(rc, err);
    if ((rc = device_check(device))) return rc;
    rc = prepare_host(p, Q, 0, h, err);
    if (rc) return fail(rc, err);
    rc = finish_device(h, device, 0, out);
    if (rc) return rc;
    cudaError_t e = cudaStreamSynchronize(0);   // uploads complete: the sampler may be used on any stream
    if (e != cudaSuccess) { coalgpu_destroy(*out); *out = nullptr; return fail(COALGPU_ECUDA, std::string("upload: ") + cudaGetErrorString(e)); }
    return 0;
}

// One launch of K1. pass = 0: single pass with the full type store; 1: first pass with the reduced store, histories that
// outgrow it go to d_defer (count in counters[6]); 3: the same with the W >= 16 variant of the reduced-store geometry;
// 2: the full store over the histories listed in d_defer.
static int sample_pass(coalgpu_sampler* s, int pass, uint64_t seed, uint64_t first, uint64_t n, cudaStream_t st,
                       double* d_weights, double* d_tmrca, double* d_lineages, uint32_t* d_nevents,
                       coalgpu_event* d_events, uint32_t* d_event_counts, uint32_t n_trace, uint32_t max_events,
                       unsigned long long* d_defer, uint64_t n_listed) {
    SampleArgs a{};
    a.seed = seed; a.first = first; a.n = n;
    a.weights = d_weights; a.tmrca = d_tmrca; a.lineages = d_lineages; a.nevents = d_nevents;
    a.events = d_events; a.event_counts = d_event_counts; a.n_trace = d_events ? n_trace : 0; a.max_events = max_events;
    a.counters = s->d_counters;
    const bool fast = pass == 1, wide = pass == 3;
    if (pass != 2) {
        CU_OK(cudaMemsetAsync(s->d_counters, 0, sizeof(unsigned long long), st));
        CU_OK(cudaMemsetAsync(s->d_counters + 6, 0, 2 * sizeof(unsigned long long), st));
    }
    if (pass == 1 || pass == 3) a.defer_list = d_defer;
    if (pass == 2) { a.list = d_defer; a.list_count = s->d_counters + 6; }
    const DevProblem& P = fast ? s->P_fast : wide ? s->P_wide : s->P;
    const int group = fast ? s->fast_group : wide ? s->wide_group : s->group;
    const bool ls = fast ? s->fast_ls : wide ? s->wide_ls : s->ls;
    const size_t smem = fast ? s->fast_smem : wide ? s->wide_smem : s->smem_bytes;
    int grid = fast ? s->fast_grid : wide ? s->wide_grid : s->grid;
    const int groups_per_cta = geo_hist_per_cta(group, ls);
    const unsigned long long todo = pass == 2 ? n_listed : n;
    const unsigned long long ctas_needed = (todo + groups_per_cta - 1) / groups_per_cta;
    if ((unsigned long long)grid > ctas_needed) grid = (int)std::max<unsigned long long>(ctas_needed, 1);
    history_kernel_fn kfn = geo_kernel(group, ls, P.Ltot <= 32);   // shared-memory limit: ensure_kernel_attributes
    void* d_store = nullptr;
    if (const size_t sb = geo_store_bytes(P, group, ls)) {           // capacity tier: one slice of global memory per CTA
        CU_OK(dev_alloc(&d_store, sb * (size_t)grid, st));
        a.gstore = (unsigned char*)d_store;
    }
    kfn<<<grid, geo_threads(group, ls), smem, st>>>(P, a);
    g_launches++;
    const cudaError_t le = cudaGetLastError();
    dev_free(d_store, st);
    CU_OK(le);
    s->last_stream = st;
    return 0;
}

// The reduced-store first pass pays off when the launch can fill the device; a launch smaller than what the full-store
// geometry keeps in flight is latency bound and runs single-pass with the wider groups.
static bool use_two_pass(const coalgpu_sampler* s, uint64_t n) {
    return s->fast && (s->fast_forced || n >= (uint64_t)s->grid * geo_hist_per_cta(s->group, s->ls));
}

int coalgpu_sample(coalgpu_sampler* s, uint64_t seed, uint64_t first, uint64_t n, void* stream,
                   double* d_weights, double* d_tmrca, double* d_lineages, uint32_t* d_nevents,
                   coalgpu_event* d_events, uint32_t* d_event_counts, uint32_t n_trace, uint32_t max_events) {
    if (!s) return fail(COALGPU_EINVAL, "null sampler");
    if (n == 0) return 0;
    CU_OK(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (!use_two_pass(s, n))
        return sample_pass(s, 0, seed, first, n, st, d_weights, d_tmrca, d_lineages, d_nevents, d_events, d_event_counts, n_trace, max_events, nullptr, 0);
    // Asynchronous contract: the host cannot look at the number of deferred histories, so the second pass is always
    // launched, with the full grid; when the list is empty (the usual case) its CTAs leave after one load.
    unsigned long long* d_defer = nullptr;
    CU_OK(dev_alloc((void**)&d_defer, n * sizeof(unsigned long long), st));
    int rc = sample_pass(s, 1, seed, first, n, st, d_weights, d_tmrca, d_lineages, d_nevents, d_events, d_event_counts, n_trace, max_events, d_defer, 0);
    if (!rc) rc = sample_pass(s, 2, seed, first, n, st, d_weights, d_tmrca, d_lineages, d_nevents, d_events, d_event_counts, n_trace, max_events, d_defer, n);
    cudaFreeAsync(d_defer, st);
    return rc;
}

int coalgpu_counters(coalgpu_sampler* s, uint64_t* n_events, uint64_t* n_pismc, uint64_t* n_overflow) {
    if (!s) return fail(COALGPU_EINVAL, "null sampler");
    CU_OK(cudaSetDevice(s->device));
    unsigned long long h[kCounters];
    CU_OK(cudaMemcpy(h, s->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    g_d2h_bytes += sizeof h;
    s->n_nan_seen = h[9];
    if (n_events) *n_events = h[1];
    if (n_pismc) *n_pismc = h[2];
    if (n_overflow) *n_overflow = h[3];
    return 0;
}

int coalgpu_work_counters(coalgpu_sampler* s, uint64_t* n_two_locus, uint64_t* n_classes) {
    if (!s) return fail(COALGPU_EINVAL, "null sampler");
    CU_OK(cudaSetDevice(s->device));
    unsigned long long h[8];
    CU_OK(cudaMemcpy(h, s->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    if (n_two_locus) *n_two_locus = h[4];
    if (n_classes) *n_classes = h[5];
    return 0;
}

int coalgpu_capacity_stats(coalgpu_sampler* s, uint64_t* n_deferred, uint32_t* store_first_pass, uint32_t* store_full,
                           uint32_t* peak_types, uint32_t* lanes_per_history) {
    if (!s) return fail(COALGPU_EINVAL, "null sampler");
    CU_OK(cudaSetDevice(s->device));
    unsigned long long h[kCounters];
    CU_OK(cudaMemcpy(h, s->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    if (n_deferred) *n_deferred = s->fast ? h[6] : 0;
    if (store_first_pass) *store_first_pass = (uint32_t)(s->fast ? s->P_fast.kmax : s->P.kmax);
    if (store_full) *store_full = (uint32_t)s->P.kmax;
    if (peak_types) *peak_types = (uint32_t)h[8];
    if (lanes_per_history) {      // lockstep form: histories per CTA (its serial phase runs one thread per history)
        const int g = s->fast ? s->fast_group : s->group; const bool ls = s->fast ? s->fast_ls : s->ls;
        *lanes_per_history = (uint32_t)(ls ? (g & 0xff) : g);
    }
    return 0;
}

// Parity seam of the device EVENT code (VERDICT r1 task 5): the given configurations go through the history kernel's own
// phases -- removal of the chosen lineage, partner list, one-locus and two-locus queries, candidate assembly -- and the
// unnormalised proposal probabilities of every backward move come back instead of a move being drawn.
int coalgpu_score_events(coalgpu_sampler* s, int64_t n_cfg, const int32_t* cfg_offsets, const uint8_t* gamma, const uint64_t* words,
                         const uint32_t* counts, const uint8_t* chosen_gamma, const uint64_t* chosen_words, const int32_t* epochs,
                         int32_t max_cand, double* out_pv, uint64_t* out_partner_words, int32_t* out_ncand) {
    if (!s || !cfg_offsets || !gamma || !words || !counts || !chosen_gamma || !chosen_words || !epochs || !out_pv || !out_partner_words || !out_ncand)
        return fail(COALGPU_EINVAL, "null argument");
    if (n_cfg <= 0) return 0;
    if (!s->ls) return fail(COALGPU_EINVAL, "coalgpu_score_events runs the lockstep form of the history kernel (unset COALGPU_KERNEL / COALGPU_GROUP)");
    if (max_cand < 1) return fail(COALGPU_EINVAL, "max_cand must be positive");
    const uint64_t maskA = s->P.maskA, maskB = s->P.maskB;
    for (int64_t q = 0; q < n_cfg; q++) {
        if (cfg_offsets[q + 1] < cfg_offsets[q] || cfg_offsets[0] != 0) return fail(COALGPU_EINVAL, "cfg_offsets must start at 0 and be non-decreasing");
        if (epochs[q] < 0 || epochs[q] >= s->T) return fail(COALGPU_EINVAL, "epoch out of range");
        long long n = 0; bool has = false;
        if (cfg_offsets[q + 1] - cfg_offsets[q] > s->P.kmax) return fail(COALGPU_EINVAL, "configuration has more types than the sampler's type store");
        for (int k = cfg_offsets[q]; k < cfg_offsets[q + 1]; k++) {
            if (gamma[k] > 2 || counts[k] == 0) return fail(COALGPU_EINVAL, "bad type record");
            const uint64_t allowed = gamma[k] == COALGPU_GAMMA_A ? maskA : (gamma[k] == COALGPU_GAMMA_B ? maskB : (maskA | maskB));
            if (words[k] & ~allowed) return fail(COALGPU_EINVAL, "haplotype word has bits outside its loci");
            n += counts[k];
            has = has || (gamma[k] == chosen_gamma[q] && words[k] == chosen_words[q]);
        }
        if (!has) return fail(COALGPU_EINVAL, "the chosen lineage is not in its configuration");
        if (n < 1 || n > s->tab.n_max) return fail(COALGPU_EINVAL, "configuration has more lineages than the sampler's tables cover");
    }
    CU_OK(cudaSetDevice(s->device));
    const size_t nt = (size_t)cfg_offsets[n_cfg];
    std::vector<void*> tmp;
    auto freeall = [&]() { for (void* p : tmp) dev_free(p); };
    int *d_off = nullptr, *d_ep = nullptr, *d_nc = nullptr; unsigned char *d_g = nullptr, *d_xg = nullptr;
    unsigned long long *d_w = nullptr, *d_xw = nullptr, *d_pw = nullptr; unsigned int* d_c = nullptr; double* d_pv = nullptr;
#define SC_TMP(ptr, bytes, src) do { cudaError_t _e = dev_alloc((void**)&ptr, std::max<size_t>(bytes, 8)); if (_e == cudaSuccess) { tmp.push_back(ptr); if (src) _e = cudaMemcpy(ptr, src, bytes, cudaMemcpyHostToDevice); } if (_e != cudaSuccess) { freeall(); return fail(COALGPU_ECUDA, cudaGetErrorString(_e)); } } while (0)
    SC_TMP(d_off, (size_t)(n_cfg + 1) * 4, cfg_offsets); SC_TMP(d_g, nt, gamma); SC_TMP(d_w, nt * 8, words); SC_TMP(d_c, nt * 4, counts);
    SC_TMP(d_xg, (size_t)n_cfg, chosen_gamma); SC_TMP(d_xw, (size_t)n_cfg * 8, chosen_words); SC_TMP(d_ep, (size_t)n_cfg * 4, epochs);
    SC_TMP(d_pv, (size_t)n_cfg * max_cand * 8, (const void*)nullptr); SC_TMP(d_pw, (size_t)n_cfg * max_cand * 8, (const void*)nullptr);
    SC_TMP(d_nc, (size_t)n_cfg * 4, (const void*)nullptr);
#undef SC_TMP
    cudaError_t e = cudaMemset(d_nc, 0xff, (size_t)n_cfg * 4);          // -1: not scored (type store overflow)
    if (e == cudaSuccess) e = cudaMemset(d_pv, 0, (size_t)n_cfg * max_cand * 8);
    if (e == cudaSuccess) e = cudaMemset(s->d_counters, 0, sizeof(unsigned long long));
    if (e == cudaSuccess) {
        SampleArgs a{};
        a.n = (unsigned long long)n_cfg; a.counters = s->d_counters;
        a.score_off = d_off; a.score_g = d_g; a.score_w = d_w; a.score_c = d_c; a.score_xg = d_xg; a.score_xw = d_xw; a.score_epoch = d_ep;
        a.score_pv = d_pv; a.score_pw = d_pw; a.score_ncand = d_nc; a.score_stride = max_cand;
        const int NH = s->group & 0xff, NT = geo_threads(s->group, true);
        const int grid = std::max((int)std::min<long long>(s->grid, (n_cfg + NH - 1) / NH), 1);
        history_kernel_fn kfn = geo_kernel(s->group, true, s->P.Ltot <= 32);
        void* d_store = nullptr;
        if (const size_t sb = geo_store_bytes(s->P, s->group, true)) { e = dev_alloc(&d_store, sb * (size_t)grid, 0); a.gstore = (unsigned char*)d_store; }
        if (e == cudaSuccess) {
            kfn<<<grid, NT, s->smem_bytes, 0>>>(s->P, a);
            g_launches++;
            e = cudaGetLastError();
        }
        dev_free(d_store, 0);
        s->last_stream = 0;
    }
    if (e == cudaSuccess) e = cudaMemcpy(out_pv, d_pv, (size_t)n_cfg * max_cand * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out_partner_words, d_pw, (size_t)n_cfg * max_cand * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out_ncand, d_nc, (size_t)n_cfg * 4, cudaMemcpyDeviceToHost);
    freeall();
    if (e != cudaSuccess) return fail(COALGPU_ECUDA, cudaGetErrorString(e));
    return 0;
}

// Debug builds with -DCOAL_LS_TIMING: phase clocks of the lockstep kernel (cycles in S, G, Q1, Q2 summed over CTAs, steps).
int coalgpu_phase_clocks(coalgpu_sampler* s, uint64_t* out5) {
    if (!s || !out5) return fail(COALGPU_EINVAL, "null argument");
    CU_OK(cudaSetDevice(s->device));
    unsigned long long h[kCounters];
    CU_OK(cudaMemcpy(h, s->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 5; i++) out5[i] = h[10 + i];
    return 0;
}

int coalgpu_tune(coalgpu_sampler* s) {
    if (!s) return fail(COALGPU_EINVAL, "null sampler");
    CU_OK(cudaSetDevice(s->device));
    CU_OK(cudaDeviceSynchronize());                 // no launch of this sampler may be in flight while its geometry changes
    unsigned long long h[kCounters];
    CU_OK(cudaMemcpy(h, s->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    if (h[8] == 0) return 0;                        // nothing drawn yet
    return choose_geometry(s, (int)h[8]);
}

int coalgpu_fp64_peak(int device, double* tflops) {
    if (!tflops) return fail(COALGPU_EINVAL, "null argument");
    CU_OK(cudaSetDevice(device));
    int sm_count = 0;
    CU_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    const int blocks = sm_count * 8, threads = 256, iters = 4096;
    double* d = nullptr;
    CU_OK(dev_alloc((void**)&d, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    CU_OK(cudaEventCreate(&e0)); CU_OK(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        CU_OK(cudaEventRecord(e0, 0));
        fp64_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        g_launches++;
        CU_OK(cudaEventRecord(e1, 0));
        CU_OK(cudaEventSynchronize(e1));
        float ms = 0; CU_OK(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); dev_free(d);
    *tflops = best;
    return 0;
}

int coalgpu_partials(coalgpu_sampler* s, uint64_t n, void* stream, const double* d_weights, const double* d_tmrca,
                     const double* d_lineages, double* d_partials) {
    if (!s || !d_weights || !d_tmrca || !d_lineages || !d_partials) return fail(COALGPU_EINVAL, "null argument");
    CU_OK(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int nstat = 3 + s->T;
    // block = 256 threads = (256 / cols) history rows x cols grid columns, cols = the power of two >= min(G, 32)
    int cols = 1;
    while (cols < 32 && cols < s->G) cols *= 2;
    const int rows = 256 / cols, col_tiles = (s->G + cols - 1) / cols;
    int nb = (int)std::min<uint64_t>((n + (uint64_t)rows * 4 - 1) / ((uint64_t)rows * 4), (uint64_t)std::max(1, 4 * s->sm_count / col_tiles));
    nb = std::max(1, std::min(nb, 1024));
    const size_t need = (size_t)nstat * s->G * nb * 2 * sizeof(double);
    if (need > s->block_partials_cap) {
        dev_free(s->d_block_partials, s->last_stream);      // its last reader ran on last_stream
        s->d_block_partials = nullptr; s->block_partials_cap = 0;
        CU_OK(dev_alloc((void**)&s->d_block_partials, need, st));
        s->block_partials_cap = need;
    }
    dim3 grid(nb, col_tiles, (s->T + kLseT - 1) / kLseT);
    lse_partial_kernel<<<grid, 256, 0, st>>>(n, s->G, s->T, cols, d_weights, d_tmrca, d_lineages, s->d_block_partials, s->d_counters);
    g_launches++;
    CU_OK(cudaGetLastError());
    const int total = nstat * s->G;
    lse_combine_kernel<<<(total + 127) / 128, 128, 0, st>>>(nb, total, s->d_block_partials, d_partials);
    g_launches++;
    CU_OK(cudaGetLastError());
    s->last_stream = st;
    return 0;
}

int coalgpu_finalize(int32_t G, int32_t T, const double* hp, coalgpu_result* r) {
    if (!hp || !r || !r->likelihood || !r->likelihood_sq || !r->tmrca || !r->lineages) return fail(COALGPU_EINVAL, "null argument");
    auto lse = [&](int stat, int g) { const double* q = hp + ((size_t)stat * G + g) * 2; return q[0] + std::log(q[1]); };
    for (int g = 0; g < G; g++) {
        const double like = lse(0, g);
        r->likelihood[g] = like;                      // log SUM (SamplerOutput.cpp:131 subtracts num_samples == 0)
        r->likelihood_sq[g] = lse(1, g);
        r->tmrca[g] = lse(2, g) - like;               // SamplerOutput.cpp:120-121
        for (int l = 0; l < T; l++) r->lineages[(size_t)l * G + g] = lse(3 + l, g) - like;   // :126-129
        if (r->ess) r->ess[g] = std::exp(2 * like - r->likelihood_sq[g]);
    }
    return 0;
}

// Histories [first0, first0 + count) of sampler s on the CURRENT device's default stream, chunked, reduced chunk by chunk
// into the (max, sum) pairs acc[(3+T)*G*2] on the host. Shared by coalgpu_run_sampler (the whole batch on one device) and
// coalgpu_run_sampler_comm (one index range per device). Chunking never changes which histories are drawn.
static int sample_range(coalgpu_sampler* s, uint64_t seed, uint64_t first0, uint64_t count, std::vector<double>& acc,
                        double* seconds_device, unsigned long long* n_deferred_out) {
    const int G = s->G, T = s->T, nstat = 3 + T;
    acc.assign((size_t)nstat * G * 2, 0.0);
    for (size_t i = 0; i < acc.size(); i += 2) acc[i] = -INFINITY;
    if (seconds_device) *seconds_device = 0.0;
    if (n_deferred_out) *n_deferred_out = 0;
    if (count == 0) return 0;
    // chunk the batch so that the per-history weight matrix stays below ~2 GiB
    uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(count, (2ull << 30) / ((size_t)(G + T + 1) * 8)));
    double *d_w = nullptr, *d_t = nullptr, *d_l = nullptr, *d_p = nullptr;
    unsigned long long* d_defer = nullptr; unsigned long long n_deferred_total = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<double> hp((size_t)nstat * G * 2);
    int rc = 0;
    auto cleanup = [&]() {
        dev_free(d_w); dev_free(d_t); dev_free(d_l); dev_free(d_p); dev_free(d_defer);
        if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1);
    };
#define RS_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { std::string m = std::string(#expr) + ": " + cudaGetErrorString(_e); cleanup(); return fail(_e == cudaErrorMemoryAllocation ? COALGPU_ENOMEM : COALGPU_ECUDA, m); } } while (0)
    RS_OK(dev_alloc((void**)&d_w, chunk * G * sizeof(double)));
    RS_OK(dev_alloc((void**)&d_t, chunk * sizeof(double)));
    RS_OK(dev_alloc((void**)&d_l, chunk * T * sizeof(double)));
    RS_OK(dev_alloc((void**)&d_p, hp.size() * sizeof(double)));
    RS_OK(cudaEventCreate(&e0)); RS_OK(cudaEventCreate(&e1));
    RS_OK(cudaEventRecord(e0, 0));
    // A long run starts with a pilot chunk; the reduced type store and the geometry of the rest are sized from what the
    // pilot histories needed (coalgpu_tune) instead of the input-based estimate. The pilot drains the device once, about
    // one history's latency: worth it from ~16x what the device holds in flight.
    const uint64_t pilot = (s->pilot_useful && count >= 16 * (uint64_t)s->grid * geo_hist_per_cta(s->group, s->ls)) ? 1024 : 0;
    uint64_t n = 0;
    for (uint64_t done = 0; done < count; done += n) {
        const uint64_t first = first0 + done;
        n = (done == 0 && pilot) ? pilot : std::min<uint64_t>(chunk, count - done);
        if (!use_two_pass(s, n)) {
            rc = sample_pass(s, 0, seed, first, n, 0, d_w, d_t, d_l, nullptr, nullptr, nullptr, 0, 0, nullptr, 0);
        } else {       // blocking call: look at the deferred count and launch the second pass only when it has work
            if (!d_defer) RS_OK(dev_alloc((void**)&d_defer, chunk * sizeof(unsigned long long)));
            rc = sample_pass(s, 1, seed, first, n, 0, d_w, d_t, d_l, nullptr, nullptr, nullptr, 0, 0, d_defer, 0);
            unsigned long long n_def = 0;
            if (!rc) { RS_OK(cudaMemcpy(&n_def, s->d_counters + 6, sizeof n_def, cudaMemcpyDeviceToHost)); g_d2h_bytes += sizeof n_def; }
            if (!rc && n_def) rc = sample_pass(s, 2, seed, first, n, 0, d_w, d_t, d_l, nullptr, nullptr, nullptr, 0, 0, d_defer, n_def);
            n_deferred_total += n_def;
        }
        if (!rc) rc = coalgpu_partials(s, n, nullptr, d_w, d_t, d_l, d_p);
        if (rc) { std::string m = g_err; cleanup(); return fail(rc, m); }
        RS_OK(cudaMemcpy(hp.data(), d_p, hp.size() * sizeof(double), cudaMemcpyDeviceToHost));
        g_d2h_bytes += hp.size() * sizeof(double);
        for (size_t i = 0; i < acc.size(); i += 2) {   // (max, sum) merge
            const double m1 = acc[i], m2 = hp[i], mm = std::max(m1, m2);
            if (mm > -INFINITY) acc[i + 1] = acc[i + 1] * std::exp(m1 - mm) + hp[i + 1] * std::exp(m2 - mm);
            acc[i] = mm;
        }
        if (done == 0 && pilot) {
            rc = coalgpu_tune(s);
            if (rc) { std::string m = g_err; cleanup(); return fail(rc, m); }
        }
    }
    RS_OK(cudaEventRecord(e1, 0));
    RS_OK(cudaEventSynchronize(e1));
    float ms = 0; RS_OK(cudaEventElapsedTime(&ms, e0, e1));
#undef RS_OK
    if (seconds_device) *seconds_device = ms * 1e-3;
    if (n_deferred_out) *n_deferred_out = n_deferred_total;
    cleanup();
    return 0;
}

static int run_sampler_impl(const coalgpu_problem* problem, uint64_t n_samples, uint32_t Q, uint64_t seed, int device, coalgpu_result* result) {
    if (!problem || !result) return fail(COALGPU_EINVAL, "null argument");
    if (n_samples == 0) return fail(COALGPU_EINVAL, "n_samples must be positive");
    const bool timing = std::getenv("COALGPU_TIMING") != nullptr;
    auto tick = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = tick();
    coalgpu_sampler* s = nullptr;
    int rc = create_impl(problem, Q, device, &s);
    if (rc) return rc;
    const double t_created = tick();
    std::vector<double> acc;
    double sec = 0; unsigned long long n_deferred_total = 0;
    rc = sample_range(s, seed, 0, n_samples, acc, &sec, &n_deferred_total);
    if (rc) { std::string m = g_err; coalgpu_destroy(s); return fail(rc, m); }
    const double t_sampled = tick();
    uint64_t nev = 0, npi = 0, nov = 0;
    rc = coalgpu_counters(s, &nev, &npi, &nov);
    const unsigned long long n_nan = s->n_nan_seen;
    const double t_cnt = tick();
    if (!rc) rc = coalgpu_finalize(s->G, s->T, acc.data(), result);
    result->n_histories = n_samples; result->n_events = nev; result->n_pismc = npi; result->seconds_device = sec;
    std::string m = g_err;
    if (timing && s->fast) {
        unsigned long long h[kCounters] = {0};
        cudaMemcpy(h, s->d_counters, sizeof h, cudaMemcpyDeviceToHost);
        fprintf(stderr, "coalgpu_run_sampler: first pass with a type store of %d (full %d, most needed %llu), W = %d (full %d), %d CTAs; %llu histories deferred to the full store\n",
                s->P_fast.kmax, s->P.kmax, h[8], s->fast_group, s->group, s->fast_grid, n_deferred_total);
    }
    const double t_fin = tick();
    coalgpu_destroy(s);
    if (timing) fprintf(stderr, "coalgpu_run_sampler: create %.2f ms, alloc+sample+reduce %.2f ms (device %.2f ms), counters %.2f ms, finalize %.2f ms, free %.2f ms\n",
                        1e3 * (t_created - t_begin), 1e3 * (t_sampled - t_created), 1e3 * sec, 1e3 * (t_cnt - t_sampled), 1e3 * (t_fin - t_cnt), 1e3 * (tick() - t_fin));
    if (rc) return fail(rc, m);
    if (nov) return fail(COALGPU_EOVERFLOW, "a history outgrew the type-store capacity");
    if (n_nan) return fail(COALGPU_ENAN, "a history weight is not a number (target grid / driving values out of range?)");
    return 0;
}

// Locus-pair batching: the reference runs its pairs one after another (main.cpp:242-311);
// a pair with a few thousand importance samples cannot fill 148 SMs and its HMM tables take the host longer to
// build than the GPU needs to sample it. So: host threads prepare the next pairs (tables, merged types) while the
// GPU samples the current ones, every pair runs on one of 16 streams, at most 32 pairs are in flight, and the only
// host<-device transfer per pair is one pinned copy of its (max, sum) partials and counters.
// Results are identical to n_problems calls of coalgpu_run_sampler.
static int run_sampler_batch_impl(int32_t n_problems, const coalgpu_problem* problems, uint64_t n_samples, uint32_t Q,
                                  uint64_t seed, int device, coalgpu_result* results) {
    if (n_problems < 0 || (n_problems > 0 && (!problems || !results))) return fail(COALGPU_EINVAL, "null argument");
    if (n_samples == 0) return fail(COALGPU_EINVAL, "n_samples must be positive");
    if (n_problems == 0) return 0;
    {   // argument and device errors before any thread is started
        std::string verr;
        for (int k = 0; k < n_problems; k++) if (int vrc = validate_problem(problems + k, Q, verr)) return fail(vrc, verr);
        if (int drc = device_check(device)) return drc;
    }
    constexpr int kWindow = 32, kStreams = 16;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int lookahead = (int)std::min<unsigned>(std::max(2u, hw), 32u);   // pairs being prepared ahead of the GPU
    struct Prep { coalgpu_sampler* s = nullptr; int rc = 0; std::string err; };
    struct Job {
        coalgpu_sampler* s = nullptr;
        double *d_w = nullptr, *d_t = nullptr, *d_l = nullptr, *d_p = nullptr;
        double* h_p = nullptr;            // slot of the pinned slab: partials [nstat*G*2] then 8 counters
        unsigned long long* d_defer = nullptr;
        cudaStream_t st = nullptr;
        size_t np = 0;
        cudaEvent_t e0 = nullptr, e1 = nullptr, eu = nullptr;
        int index = -1;
    };
    auto is_big = [&](const coalgpu_problem* p) {   // a pair whose weight matrix is large fills the device on its own
        const size_t G = (size_t)p->n_mu * p->n_rho * p->n_lambda;
        return n_samples * (G + p->T + 1) * 8 > (256ull << 20);
    };
    // Streams first: the workers upload on `up`. A pageable host-to-device copy blocks its caller until the stream gets
    // to it, and streams share hardware queues with each other (8 by default), so a copy can sit behind a running
    // kernel of another pair: that wait lands on a worker thread here, never on the thread that feeds the GPU.
    cudaStream_t streams[kStreams] = {}, up = nullptr;
    {
        cudaError_t e = cudaSetDevice(device);
        for (int i = 0; i < kStreams && e == cudaSuccess; i++) e = cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking);
        int arc = e == cudaSuccess ? ensure_kernel_attributes(device) : 0;
        if (e != cudaSuccess || arc) {
            for (int i = 0; i < kStreams; i++) if (streams[i]) cudaStreamDestroy(streams[i]);
            if (up) cudaStreamDestroy(up);
            return e != cudaSuccess ? fail(COALGPU_ECUDA, std::string("stream creation: ") + cudaGetErrorString(e)) : arc;
        }
    }
    std::vector<std::future<std::unique_ptr<Prep>>> preps(n_problems);
    int next_prep = 0;
    auto schedule = [&](int upto) {
        for (; next_prep < n_problems && next_prep < upto; next_prep++) {
            const coalgpu_problem* p = problems + next_prep;
            if (is_big(p)) continue;
            // one table thread per pair: the pairs are the parallel axis here
            auto work = [p, Q, device, up]() {
                std::unique_ptr<Prep> r(new Prep);
                try {
                    HostPrep h;
                    r->rc = prepare_host(p, Q, 1, h, r->err);
                    if (!r->rc) { r->rc = finish_device(h, device, up, &r->s); if (r->rc) r->err = g_err; }
                } catch (const std::bad_alloc&) { r->rc = COALGPU_ENOMEM; r->err = "out of host memory while preparing a locus pair"; }
                catch (const std::exception& e) { r->rc = COALGPU_EINVAL; r->err = std::string("preparing a locus pair: ") + e.what(); }
                return r;
            };
            try { preps[next_prep] = std::async(std::launch::async, work); }
            catch (...) { preps[next_prep] = std::async(std::launch::deferred, work); }   // no thread to be had: prepare on demand
        }
    };
    int rc = 0; std::string err;
    std::vector<Job> window;
    auto release = [&](Job& j) {
        dev_free(j.d_w); dev_free(j.d_t); dev_free(j.d_l); dev_free(j.d_p); dev_free(j.d_defer);
        if (j.e0) cudaEventDestroy(j.e0); if (j.e1) cudaEventDestroy(j.e1); if (j.eu) cudaEventDestroy(j.eu);
        if (j.s) coalgpu_destroy(j.s);
        j = Job();
    };
#define WB_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess && !rc) { rc = COALGPU_ECUDA; err = std::string(#expr) + ": " + cudaGetErrorString(_e); } } while (0)
    auto retire = [&](Job& j) {            // wait for the pair, turn its partials into the result grids
        if (j.s && !rc) {
            WB_OK(cudaEventSynchronize(j.e1));
            float ms = 0;
            if (!rc) WB_OK(cudaEventElapsedTime(&ms, j.e0, j.e1));
            if (!rc) {
                unsigned long long cnt[kCounters];
                std::memcpy(cnt, j.h_p + j.np, sizeof cnt);
                if (j.d_defer && cnt[6]) {       // histories outgrew the reduced store: second pass, reduce again
                    rc = sample_pass(j.s, 2, seed, 0, n_samples, j.st, j.d_w, j.d_t, j.d_l, nullptr, nullptr, nullptr, 0, 0, j.d_defer, cnt[6]);
                    WB_OK(cudaMemsetAsync(j.s->d_counters + 9, 0, sizeof(unsigned long long), j.st));   // the first reduction saw the deferred histories' NaN placeholders
                    if (!rc) rc = coalgpu_partials(j.s, n_samples, j.st, j.d_w, j.d_t, j.d_l, j.d_p);
                    if (rc) err = g_err;
                    WB_OK(cudaMemcpyAsync(j.d_p + j.np, j.s->d_counters, kCounters * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, j.st));
                    WB_OK(cudaMemcpyAsync(j.h_p, j.d_p, (j.np + kCounters) * sizeof(double), cudaMemcpyDeviceToHost, j.st));
                    g_d2h_bytes += (j.np + kCounters) * sizeof(double);
                    WB_OK(cudaStreamSynchronize(j.st));
                }
            }
            if (!rc) {
                coalgpu_result* r = results + j.index;
                unsigned long long cnt[kCounters];
                std::memcpy(cnt, j.h_p + j.np, sizeof cnt);
                rc = coalgpu_finalize(j.s->G, j.s->T, j.h_p, r);
                if (rc) err = g_err;
                r->n_histories = n_samples; r->n_events = cnt[1]; r->n_pismc = cnt[2]; r->seconds_device = ms * 1e-3;
                if (!rc && cnt[3]) { rc = COALGPU_EOVERFLOW; err = "a history outgrew the type-store capacity"; }
                if (!rc && cnt[9]) { rc = COALGPU_ENAN; err = "a history weight is not a number (target grid / driving values out of range?)"; }
            }
        }
        release(j);
    };
    // one pinned slab for the partials of the pairs in flight (pinned allocations synchronise the device: once per call)
    size_t slot_doubles = 0;
    for (int k = 0; k < n_problems; k++)
        if (!is_big(problems + k)) slot_doubles = std::max(slot_doubles, (size_t)(3 + problems[k].T) * problems[k].n_mu * problems[k].n_rho * problems[k].n_lambda * 2 + kCounters);
    double* slab = nullptr;
    if (slot_doubles) {
        cudaError_t e = cudaSetDevice(device);
        if (e == cudaSuccess) e = cudaMallocHost((void**)&slab, slot_doubles * kWindow * sizeof(double));
        if (e != cudaSuccess) {
            for (int i = 0; i < kStreams; i++) cudaStreamDestroy(streams[i]);
            cudaStreamDestroy(up);
            return fail(COALGPU_ECUDA, std::string("pinned host buffer: ") + cudaGetErrorString(e));
        }
    }
    long launched = 0;
    const bool timing = std::getenv("COALGPU_TIMING") != nullptr;
    auto tick = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_wait_prep = 0, t_finish = 0, t_alloc = 0, t_launch = 0, t_retire = 0; const double t_start = tick();
    schedule(lookahead);
    for (int k = 0; k < n_problems && !rc; k++) {
        const coalgpu_problem* p = problems + k;
        if (is_big(p)) {                   // sequential, chunked
            rc = run_sampler_impl(p, n_samples, Q, seed, device, results + k);
            if (rc) err = g_err;
            schedule(k + 1 + lookahead);
            continue;
        }
        double t0 = tick();
        std::unique_ptr<Prep> pr = preps[k].get();
        schedule(k + 1 + lookahead);
        t_wait_prep += tick() - t0; t0 = tick();
        if (pr->rc) { rc = pr->rc; err = pr->err; break; }
        if ((int)window.size() == kWindow) { retire(window.front()); window.erase(window.begin()); if (rc) break; }
        t_retire += tick() - t0; t0 = tick();
        window.emplace_back();
        Job& j = window.back();
        j.index = k;
        cudaStream_t st = streams[k % kStreams];
        j.s = pr->s; pr->s = nullptr;
        WB_OK(cudaEventCreateWithFlags(&j.eu, cudaEventDisableTiming));
        if (rc) break;
        WB_OK(cudaEventRecord(j.eu, up));
        WB_OK(cudaStreamWaitEvent(st, j.eu, 0));     // the pair's stream starts after its upload
        t_finish += tick() - t0; t0 = tick();
        const int T = j.s->T; const size_t G = j.s->G;
        j.np = (size_t)(3 + T) * G * 2;
        WB_OK(cudaMallocAsync((void**)&j.d_w, n_samples * G * sizeof(double), st));
        WB_OK(cudaMallocAsync((void**)&j.d_t, n_samples * sizeof(double), st));
        WB_OK(cudaMallocAsync((void**)&j.d_l, n_samples * T * sizeof(double), st));
        WB_OK(cudaMallocAsync((void**)&j.d_p, (j.np + kCounters) * sizeof(double), st));
        j.h_p = slab + (size_t)(launched++ % kWindow) * slot_doubles;   // free again: its previous user was retired above
        WB_OK(cudaEventCreate(&j.e0)); WB_OK(cudaEventCreate(&j.e1));
        if (rc) break;
        t_alloc += tick() - t0; t0 = tick();
        WB_OK(cudaEventRecord(j.e0, st));
        // Full grid per pair (one history per group when n_samples is small). Capping the grid so that every group pulls
        // several histories was measured 25 % slower: only ~8 kernels run side by side (hardware queues), so smaller grids
        // leave SMs empty (profiles/r1_batch_pipeline.md).
        // The pairs of a batch fill the device together, so every pair with a reduced store runs its first pass with it
        // (W = 8 only when the pair could fill the device on its own: about 8 launches run side by side, and a
        // thousand histories each cannot fill the larger capacity of W = 8 -- measured 227 k vs 288 k at cfg2);
        // the rare pair that deferred histories gets its second pass when it is retired.
        j.st = st;
        const int first_pass = use_two_pass(j.s, n_samples) ? 1 : (j.s->wide ? 3 : 0);
        if (first_pass) {
            WB_OK(cudaMallocAsync((void**)&j.d_defer, n_samples * sizeof(unsigned long long), st));
            if (rc) break;
            rc = sample_pass(j.s, first_pass, seed, 0, n_samples, st, j.d_w, j.d_t, j.d_l, nullptr, nullptr, nullptr, 0, 0, j.d_defer, 0);
        } else {
            rc = sample_pass(j.s, 0, seed, 0, n_samples, st, j.d_w, j.d_t, j.d_l, nullptr, nullptr, nullptr, 0, 0, nullptr, 0);
        }
        if (!rc) rc = coalgpu_partials(j.s, n_samples, st, j.d_w, j.d_t, j.d_l, j.d_p);
        if (rc) { err = g_err; break; }
        WB_OK(cudaMemcpyAsync(j.d_p + j.np, j.s->d_counters, kCounters * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
        WB_OK(cudaMemcpyAsync(j.h_p, j.d_p, (j.np + kCounters) * sizeof(double), cudaMemcpyDeviceToHost, st));
        g_d2h_bytes += (j.np + kCounters) * sizeof(double);
        WB_OK(cudaEventRecord(j.e1, st));
        t_launch += tick() - t0;
    }
    const double t_loop = tick();
    if (rc) cudaDeviceSynchronize();       // error path: nothing may still be running on buffers that are about to be freed
    for (Job& j : window) retire(j);
#undef WB_OK
    for (int k = 0; k < n_problems; k++)       // no worker may outlive the call; samplers that were never launched go away
        if (preps[k].valid()) { std::unique_ptr<Prep> pr = preps[k].get(); if (pr && pr->s) { cudaStreamSynchronize(up); coalgpu_destroy(pr->s); } }
    cudaDeviceSynchronize();
    for (int i = 0; i < kStreams; i++) cudaStreamDestroy(streams[i]);
    cudaStreamDestroy(up);
    if (slab) cudaFreeHost(slab);
    if (timing) fprintf(stderr, "coalgpu_run_sampler_batch: %d pairs, %.1f ms total: wait for host preparation + upload %.1f, retire (in loop) %.1f, stream order %.1f, alloc %.1f, launch %.1f, drain %.1f ms\n",
                        n_problems, 1e3 * (tick() - t_start), 1e3 * t_wait_prep, 1e3 * t_retire, 1e3 * t_finish, 1e3 * t_alloc, 1e3 * t_launch, 1e3 * (tick() - t_loop));
    if (rc) return fail(rc, err);
    return 0;
}

// No C++ exception crosses the C ABI (std::bad_alloc from the table vectors, std::system_error from std::thread /
// std::async): they become return codes like every other failure.
#define COALGPU_GUARD(call)                                                                                         \
    try { return call; }                                                                                            \
    catch (const std::bad_alloc&) { return fail(COALGPU_ENOMEM, "out of host memory"); }                           \
    catch (const std::exception& e) { return fail(COALGPU_EINVAL, std::string("internal error: ") + e.what()); }   \
    catch (...) { return fail(COALGPU_EINVAL, "internal error"); }
int coalgpu_create(const coalgpu_problem* p, uint32_t Q, int device, coalgpu_sampler** out) { COALGPU_GUARD(create_impl(p, Q, device, out)) }
int coalgpu_run_sampler(const coalgpu_problem* problem, uint64_t n_samples, uint32_t Q, uint64_t seed, int device, coalgpu_result* result) {
    COALGPU_GUARD(run_sampler_impl(problem, n_samples, Q, seed, device, result))
}
int coalgpu_run_sampler_batch(int32_t n_problems, const coalgpu_problem* problems, uint64_t n_samples, uint32_t Q,
                              uint64_t seed, int device, coalgpu_result* results) {
    COALGPU_GUARD(run_sampler_batch_impl(n_problems, problems, n_samples, Q, seed, device, results))
}

int coalgpu_tables(coalgpu_sampler* s, int32_t n_all, int32_t time_idx, double* w, double* y, double* z, double* EA, double* EB) {
    if (!s) return fail(COALGPU_EINVAL, "null sampler");
    if (time_idx < 0 || time_idx >= s->T || n_all < 1 || n_all > s->tab.n_max) return fail(COALGPU_EINVAL, "table index out of range");
    const TableLayout& lay = s->tab.lay;
    const int set = s->tab.set_of_epoch[time_idx];
    if (s->rows_on_device_only) {       // emission parts live on the device only: fetch the rows once
        if (s->tab.rows.empty()) return fail(COALGPU_EINVAL, "table rows were released (batch sampler)");
        CU_OK(cudaSetDevice(s->device));
        CU_OK(cudaDeviceSynchronize());
        CU_OK(cudaMemcpy(s->tab.rows.data(), s->d_full_rows, s->tab.rows.size() * sizeof(double), cudaMemcpyDeviceToHost));
        s->rows_on_device_only = false;
    }
    const double* row = s->tab.row(set, n_all);
    const double* yy = s->tab.y.data() + ((size_t)set * s->tab.n_max + (n_all - 1)) * kQ;
    if (w) std::memcpy(w, s->tab.w.data(), sizeof(double) * kQ);
    if (y) std::memcpy(y, yy, sizeof(double) * kQ);
    if (z) for (int i = 0; i < kQ; i++) for (int j = 0; j < kQ; j++)
        z[i * kQ + j] = (i > j) ? row[lay.off_zlow + j] : (i < j ? s->tab.w[j] * row[lay.off_g + i] : row[lay.off_zdiag + i]);
    if (EA) for (int i = 0; i < kQ; i++) for (int d = 0; d <= s->LA; d++) EA[i * (s->LA + 1) + d] = row[lay.off_EA + d * kEStride + i];
    if (EB) for (int i = 0; i < kQ; i++) for (int d = 0; d <= s->LB; d++) EB[i * (s->LB + 1) + d] = row[lay.off_EB + d * kEStride + i];
    return 0;
}

int coalgpu_pismc_classes(coalgpu_sampler* s, int64_t nq, const uint8_t* gamma, const int32_t* n_all, const int32_t* time_idx,
                          const int64_t* offs, const int32_t* counts, const int32_t* dA, const int32_t* dB, double* out_pi) {
    if (!s || !gamma || !n_all || !time_idx || !offs || !out_pi) return fail(COALGPU_EINVAL, "null argument");
    if (nq <= 0) return 0;
    for (int64_t q = 0; q < nq; q++) {
        if (offs[q + 1] - offs[q] > s->LA + s->LB + 2) return fail(COALGPU_EINVAL, "more classes than dA+dB sums");
        if (time_idx[q] < 0 || time_idx[q] >= s->T || n_all[q] < 1 || n_all[q] > s->tab.n_max) return fail(COALGPU_EINVAL, "query outside the table range");
        if (gamma[q] > 2) return fail(COALGPU_EINVAL, "bad gamma");
    }
    CU_OK(cudaSetDevice(s->device));
    const int64_t ncls = offs[nq];
    unsigned char* d_g = nullptr; int *d_n = nullptr, *d_t = nullptr, *d_c = nullptr, *d_a = nullptr, *d_b = nullptr; long long* d_o = nullptr; double* d_out = nullptr;
    std::vector<void*> tmp;
    auto freeall = [&]() { for (void* p : tmp) dev_free(p); };
#define TMP(ptr, bytes, src) do { cudaError_t _e = dev_alloc((void**)&ptr, std::max<size_t>(bytes, 8)); if (_e == cudaSuccess) { tmp.push_back(ptr); if (src) _e = cudaMemcpy(ptr, src, bytes, cudaMemcpyHostToDevice); } if (_e != cudaSuccess) { freeall(); return fail(COALGPU_ECUDA, cudaGetErrorString(_e)); } } while (0)
    TMP(d_g, (size_t)nq, gamma); TMP(d_n, (size_t)nq * 4, n_all); TMP(d_t, (size_t)nq * 4, time_idx);
    TMP(d_o, (size_t)(nq + 1) * 8, offs); TMP(d_c, (size_t)ncls * 4, counts); TMP(d_a, (size_t)ncls * 4, dA); TMP(d_b, (size_t)ncls * 4, dB);
    TMP(d_out, (size_t)nq * 8, (const void*)nullptr);
#undef TMP
    const int grid = (int)std::min<int64_t>((nq + 7) / 8, 148 * 8);
    pismc_classes_kernel<<<grid, 128, s->k2_smem>>>(s->P, (long long)nq, d_g, d_n, d_t, d_o, d_c, d_a, d_b, d_out);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(out_pi, d_out, (size_t)nq * 8, cudaMemcpyDeviceToHost);
    freeall();
    if (e != cudaSuccess) return fail(COALGPU_ECUDA, cudaGetErrorString(e));
    return 0;
}

}  // extern "C"

#include "coal_multi.cu"
