// Sample-sharded multi-GPU sampling behind the C ABI (included at the end of coal_api.cu: same translation unit).
//
// Reference analogue: ImportanceSampler::runSampler splits num_iterations over std::threads, every thread draws its
// share into a local SamplerOutput, merges it into the global one under a mutex and the caller log-sums the union
// (ImportanceSampler.cpp:110-145, SamplerOutput.cpp:97-133). Here a "thread" is a B200: one host thread per device
// draws the contiguous index range shard_range(n, r, R) (a history is a pure function of (seed, index), so the set of
// histories does not depend on R), reduces it on its device to (max, sum exp(x - max)) pairs per statistic and grid
// point, and the pairs of the R devices are combined over NVLink by NCCL: allreduce(MAX) of the maxima, a local
// rescale of every device's sums to the common maximum, allreduce(SUM) of the sums. There is no data-path collective.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch has already loaded when the caller is a Python
// process, else the system library), so libcoalgpu.so itself carries no link-time dependency on it and single-GPU
// callers never touch it. Only the types of <nccl.h> are used at compile time.
#include <dlfcn.h>
#include <nccl.h>

#include <map>
#include <mutex>

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        const char* names[] = {std::getenv("COALGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);          // the copy the process already holds (torch's), if any
            if (!api.handle) api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : ""); return; }
        auto sym = [&](const char* n) { void* p = dlsym(api.handle, n); if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + n; return p; };
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.CommCount = (decltype(api.CommCount))sym("ncclCommCount");
    });
    return &api;
}

std::atomic<unsigned long long> g_nccl_calls{0};

}  // namespace

struct coalgpu_comm;
namespace {
std::mutex g_comm_cache_mutex;
std::map<std::vector<int>, coalgpu_comm*> g_comm_cache;
}

struct coalgpu_comm {
    std::vector<int> devices;
    std::vector<ncclComm_t> comms;
    std::vector<cudaStream_t> streams;
    std::mutex busy;              // one collective sequence at a time per communicator
    int nccl_version = 0;
};

extern "C" {

uint64_t coalgpu_nccl_calls(void) { return g_nccl_calls.load(); }

void coalgpu_comm_finalize(coalgpu_comm* c) {
    if (!c) return;
    NcclApi* N = nccl_api();
    for (size_t r = 0; r < c->devices.size(); r++) {
        cudaSetDevice(c->devices[r]);
        if (r < c->streams.size() && c->streams[r]) { cudaStreamSynchronize(c->streams[r]); cudaStreamDestroy(c->streams[r]); }
        if (r < c->comms.size() && c->comms[r] && N->CommDestroy) N->CommDestroy(c->comms[r]);
    }
    delete c;
}

static int comm_init_impl(int32_t n_devices, const int32_t* devices, coalgpu_comm** out) {
    if (!out) return fail(COALGPU_EINVAL, "null argument");
    *out = nullptr;
    if (n_devices < 1) return fail(COALGPU_EINVAL, "n_devices must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(COALGPU_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(e));
    std::unique_ptr<coalgpu_comm> c(new coalgpu_comm);
    for (int r = 0; r < n_devices; r++) {
        const int d = devices ? devices[r] : r;
        if (d < 0 || d >= ndev) return fail(COALGPU_ENODEV, "device index out of range");
        for (int q : c->devices) if (q == d) return fail(COALGPU_EINVAL, "a device is listed twice");
        c->devices.push_back(d);
    }
    c->comms.assign(n_devices, nullptr);
    c->streams.assign(n_devices, nullptr);
    if (n_devices > 1) {
        NcclApi* N = nccl_api();
        if (!N->error.empty()) return fail(COALGPU_ENCCL, N->error);
        if (N->GetVersion) N->GetVersion(&c->nccl_version);
        const ncclResult_t nr = N->CommInitAll(c->comms.data(), n_devices, c->devices.data());
        if (nr != ncclSuccess) return fail(COALGPU_ENCCL, std::string("ncclCommInitAll: ") + N->GetErrorString(nr));
    }
    for (int r = 0; r < n_devices; r++) {
        e = cudaSetDevice(c->devices[r]);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->streams[r], cudaStreamNonBlocking);
        if (e != cudaSuccess) { const std::string m = std::string("stream creation: ") + cudaGetErrorString(e); coalgpu_comm_finalize(c.release()); return fail(COALGPU_ECUDA, m); }
    }
    *out = c.release();
    return 0;
}

int coalgpu_comm_size(const coalgpu_comm* c) { return c ? (int)c->devices.size() : 0; }

// [first, first + count) of rank r: floor(n / R) each, the remainder to the lowest ranks (the reference gives it to its
// first thread, ImportanceSampler.cpp:113-114,128). Same rule as coalescenator_b200/distributed.py: shard_range.
static void shard_range(uint64_t n, int r, int R, uint64_t* first, uint64_t* count) {
    const uint64_t base = n / (uint64_t)R, rem = n % (uint64_t)R;
    *count = base + ((uint64_t)r < rem ? 1 : 0);
    *first = (uint64_t)r * base + std::min<uint64_t>((uint64_t)r, rem);
}

static int run_sampler_comm_impl(coalgpu_comm* c, const coalgpu_problem* problem, uint64_t n_samples, uint32_t Q, uint64_t seed, coalgpu_result* result) {
    if (!c || !problem || !result) return fail(COALGPU_EINVAL, "null argument");
    if (n_samples == 0) return fail(COALGPU_EINVAL, "n_samples must be positive");
    const int R = (int)c->devices.size();
    std::lock_guard<std::mutex> lk(c->busy);
    // tables and merged types once on the host, then one copy per device
    HostPrep h0; std::string err;
    int rc = prepare_host(problem, Q, 0, h0, err);
    if (rc) return fail(rc, err);
    const int G = h0.G, T = h0.T;
    const size_t npairs = (size_t)(3 + T) * G;

    struct Rank {
        int rc = 0; std::string err;
        std::vector<double> acc;              // (max, sum) pairs of this device's histories
        uint64_t nev = 0, npi = 0, nov = 0, nnan = 0;
        double seconds = 0;
    };
    std::vector<Rank> ranks(R);
    auto work = [&](int r) {
        Rank& k = ranks[r];
        try {
            HostPrep h = h0;
            coalgpu_sampler* s = nullptr;
            k.rc = finish_device(h, c->devices[r], 0, &s);        // sets the device for this thread
            if (!k.rc) { cudaError_t e = cudaStreamSynchronize(0); if (e != cudaSuccess) { k.rc = fail(COALGPU_ECUDA, std::string("upload: ") + cudaGetErrorString(e)); } }
            if (!k.rc) {
                uint64_t first = 0, count = 0;
                shard_range(n_samples, r, R, &first, &count);
                k.rc = sample_range(s, seed, first, count, k.acc, &k.seconds, nullptr);
                if (!k.rc) { k.rc = coalgpu_counters(s, &k.nev, &k.npi, &k.nov); k.nnan = s->n_nan_seen; }
            }
            if (k.rc) k.err = g_err;
            if (s) coalgpu_destroy(s);
        } catch (const std::bad_alloc&) { k.rc = COALGPU_ENOMEM; k.err = "out of host memory"; }
        catch (const std::exception& e) { k.rc = COALGPU_EINVAL; k.err = std::string("internal error: ") + e.what(); }
    };
    {
        std::vector<std::thread> th;
        for (int r = 1; r < R; r++) th.emplace_back(work, r);
        work(0);
        for (auto& t : th) t.join();
    }
    for (int r = 0; r < R; r++) if (ranks[r].rc) return fail(ranks[r].rc, "device " + std::to_string(c->devices[r]) + ": " + ranks[r].err);

    // combine: (max, sum) pairs of the R devices -> pairs of all histories, over NCCL when R > 1
    std::vector<double> total(npairs * 2);
    if (R == 1) {
        total = ranks[0].acc;
    } else {
        NcclApi* N = nccl_api();
        std::vector<double*> d_buf(R, nullptr);       // per device: mx[npairs] | sm[npairs] | gmx[npairs]
        auto release = [&]() { for (int r = 0; r < R; r++) if (d_buf[r]) { cudaSetDevice(c->devices[r]); cudaFreeAsync(d_buf[r], c->streams[r]); } };
#define MG_OK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { const std::string m = std::string(#expr) + ": " + cudaGetErrorString(_e); release(); return fail(COALGPU_ECUDA, m); } } while (0)
#define NC_OK(expr) do { ncclResult_t _n = (expr); if (_n != ncclSuccess) { const std::string m = std::string(#expr) + ": " + N->GetErrorString(_n); release(); return fail(COALGPU_ENCCL, m); } } while (0)
        std::vector<std::vector<double>> split(R, std::vector<double>(npairs * 2));
        for (int r = 0; r < R; r++) {
            for (size_t i = 0; i < npairs; i++) { split[r][i] = ranks[r].acc[2 * i]; split[r][npairs + i] = ranks[r].acc[2 * i + 1]; }
            MG_OK(cudaSetDevice(c->devices[r]));
            MG_OK(cudaMallocAsync((void**)&d_buf[r], npairs * 3 * sizeof(double), c->streams[r]));
            MG_OK(cudaMemcpyAsync(d_buf[r], split[r].data(), npairs * 2 * sizeof(double), cudaMemcpyHostToDevice, c->streams[r]));
        }
        NC_OK(N->GroupStart());
        for (int r = 0; r < R; r++) NC_OK(N->AllReduce(d_buf[r], d_buf[r] + 2 * npairs, npairs, ncclDouble, ncclMax, c->comms[r], c->streams[r]));
        NC_OK(N->GroupEnd());
        g_nccl_calls += R;
        for (int r = 0; r < R; r++) {
            MG_OK(cudaSetDevice(c->devices[r]));
            lse_rescale_kernel<<<(int)((npairs + 127) / 128), 128, 0, c->streams[r]>>>((int)npairs, d_buf[r], d_buf[r] + 2 * npairs, d_buf[r] + npairs);
            g_launches++;
            MG_OK(cudaGetLastError());
        }
        NC_OK(N->GroupStart());
        for (int r = 0; r < R; r++) NC_OK(N->AllReduce(d_buf[r] + npairs, d_buf[r] + npairs, npairs, ncclDouble, ncclSum, c->comms[r], c->streams[r]));
        NC_OK(N->GroupEnd());
        g_nccl_calls += R;
        std::vector<double> back(npairs * 2);       // sm | gmx of device 0 (identical on every device)
        MG_OK(cudaSetDevice(c->devices[0]));
        MG_OK(cudaMemcpyAsync(back.data(), d_buf[0] + npairs, npairs * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->streams[0]));
        for (int r = 0; r < R; r++) { MG_OK(cudaSetDevice(c->devices[r])); MG_OK(cudaStreamSynchronize(c->streams[r])); }
        for (size_t i = 0; i < npairs; i++) { total[2 * i] = back[npairs + i]; total[2 * i + 1] = back[i]; }
        release();
#undef MG_OK
#undef NC_OK
    }
    rc = coalgpu_finalize(G, T, total.data(), result);
    if (rc) return rc;
    uint64_t nev = 0, npi = 0, nov = 0, nnan = 0; double sec = 0;
    for (const Rank& k : ranks) { nev += k.nev; npi += k.npi; nov += k.nov; nnan += k.nnan; sec = std::max(sec, k.seconds); }
    result->n_histories = n_samples; result->n_events = nev; result->n_pismc = npi; result->seconds_device = sec;
    if (nov) return fail(COALGPU_EOVERFLOW, "a history outgrew the type-store capacity");
    if (nnan) return fail(COALGPU_ENAN, "a history weight is not a number (target grid / driving values out of range?)");
    return 0;
}

int coalgpu_comm_init(int32_t n_devices, const int32_t* devices, coalgpu_comm** out) { COALGPU_GUARD(comm_init_impl(n_devices, devices, out)) }
int coalgpu_run_sampler_comm(coalgpu_comm* c, const coalgpu_problem* problem, uint64_t n_samples, uint32_t Q, uint64_t seed, coalgpu_result* result) {
    COALGPU_GUARD(run_sampler_comm_impl(c, problem, n_samples, Q, seed, result))
}

// The convenience form: the communicator for a device list is created on first use and kept for the life of the
// process (ncclCommInitAll costs about a second); coalgpu_multi_shutdown releases them.
static int run_sampler_multi_impl(const coalgpu_problem* problem, uint64_t n_samples, uint32_t Q, uint64_t seed, int32_t n_devices,
                                  const int32_t* devices, coalgpu_result* result) {
    if (n_devices < 1) return fail(COALGPU_EINVAL, "n_devices must be positive");
    std::vector<int> key;
    for (int r = 0; r < n_devices; r++) key.push_back(devices ? devices[r] : r);
    coalgpu_comm* c = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_comm_cache_mutex);
        auto it = g_comm_cache.find(key);
        if (it != g_comm_cache.end()) c = it->second;
        else {
            const int rc = comm_init_impl(n_devices, key.data(), &c);
            if (rc) return rc;
            g_comm_cache[key] = c;
        }
    }
    return run_sampler_comm_impl(c, problem, n_samples, Q, seed, result);
}
int coalgpu_run_sampler_multi(const coalgpu_problem* problem, uint64_t n_samples, uint32_t Q, uint64_t seed, int32_t n_devices,
                              const int32_t* devices, coalgpu_result* result) {
    COALGPU_GUARD(run_sampler_multi_impl(problem, n_samples, Q, seed, n_devices, devices, result))
}
void coalgpu_multi_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_comm_cache_mutex);
    for (auto& kv : g_comm_cache) coalgpu_comm_finalize(kv.second);
    g_comm_cache.clear();
}

}  // extern "C"
