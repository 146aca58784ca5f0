// TEST INFRASTRUCTURE (oracle/): driver around the UNMODIFIED reference objects compiled from
// /root/reference/src into oracle/_ref/. It performs what main.cpp:229-306 does for ONE locus pair
// without boost::program_options: build a Sample (Sample.hpp:56), fill ModelParameters
// (Utils.hpp:157-166), call ImportanceSampler::runSampler (ImportanceSampler.cpp:110) and print the
// four result grids (main.cpp:288-306) with full precision. Nothing in the product path links this.
//
// usage: coalref <problem-file> <n_samples> <n_threads> <seed> [trace-file [max_trace_histories]]
//   (the trace arguments only do something in the coalref_trace binary, see ref_trace.cpp)

#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "Utils.hpp"
#include "Sample.hpp"
#include "SamplerOutput.hpp"
#include "ImportanceSampler.hpp"
#include "Matrix3d.hpp"
#ifdef COALREF_GPU_BINDING
#include "GpuImportanceSampler.hpp"   // include/: the binding a maintainer adds
#endif

// implemented in ref_trace.cpp (tracing build) or ref_notrace.cpp (plain build)
void coalref_trace_open(const char* path, long max_histories, const std::vector<AncestralTypeMaps>* type_data);
void coalref_trace_close();

static void die(const std::string& m) { fprintf(stderr, "coalref: %s\n", m.c_str()); exit(2); }

struct Problem {
    uint T = 0, LA = 0, LB = 0;
    std::vector<double> times, viral;
    ModelParameters mp;
    std::vector<AncestralTypeMaps> maps;
};

static std::vector<double> read_list(std::istream& in, const char* key) {
    std::string k; size_t n;
    if (!(in >> k >> n) || k != key) die(std::string("expected list ") + key);
    std::vector<double> v(n);
    for (auto& x : v) if (!(in >> x)) die(std::string("short list ") + key);
    return v;
}

static Problem read_problem(const char* path) {
    std::ifstream in(path);
    if (!in) die(std::string("cannot open ") + path);
    Problem p; std::string k; int ver;
    if (!(in >> k >> ver) || k != "coalprob" || ver != 1) die("bad header");
    std::string kT, kA, kB;
    if (!(in >> kT >> p.T >> kA >> p.LA >> kB >> p.LB) || kT != "T" || kA != "LA" || kB != "LB") die("bad dims");
    if (!(in >> k) || k != "times") die("expected times");
    p.times.resize(p.T); for (auto& x : p.times) in >> x;
    if (!(in >> k) || k != "viral") die("expected viral");
    p.viral.resize(p.T); for (auto& x : p.viral) in >> x;
    if (!(in >> k >> p.mp.tau) || k != "tau") die("expected tau");
    if (!(in >> k >> p.mp.driving_mu >> p.mp.driving_lambda >> p.mp.driving_rho) || k != "driving") die("expected driving");
    p.mp.target_mu = read_list(in, "mu");
    p.mp.target_rho = read_list(in, "rho");
    p.mp.target_lambda = read_list(in, "lambda");
    p.maps.resize(p.T);
    for (uint t = 0; t < p.T; t++) {
        uint tt; size_t n;
        if (!(in >> k >> tt >> n) || k != "types" || tt != t) die("expected types block");
        for (size_t i = 0; i < n; i++) {
            std::string g, bits; uint cnt;
            if (!(in >> g >> bits >> cnt)) die("short types block");
            std::vector<bool> v; v.reserve(bits.size());
            for (char ch : bits) v.push_back(ch == '1');
            TypeMap* m = nullptr; size_t want = 0;
            if (g == "A") { m = &p.maps[t].A; want = p.LA; }
            else if (g == "B") { m = &p.maps[t].B; want = p.LB; }
            else if (g == "C") { m = &p.maps[t].C; want = p.LA + p.LB; }
            else die("bad gamma");
            if (v.size() != want) die("haplotype length mismatch");
            // same accumulate-on-duplicate rule as Sample.cpp:632-662
            if (!(m->insert({v, cnt}).second)) m->at(v) += cnt;
        }
    }
    return p;
}

static void print_grid(const char* name, Matrix3d<double>& m, bool last) {
    printf("\"%s\": [", name);
    bool first = true;
    for (uint i = 0; i < m.dim1(); i++) for (uint j = 0; j < m.dim2(); j++) for (uint k = 0; k < m.dim3(); k++) {
        printf("%s%.17g", first ? "" : ", ", m.getElement(i, j, k));
        first = false;
    }
    printf("]%s\n", last ? "" : ",");
}

int main(int argc, char** argv) {
    if (argc < 5) die("usage: coalref <problem-file> <n_samples> <n_threads> <seed> [trace-file [max_trace_histories]]");
    Problem p = read_problem(argv[1]);
    uint n_samples = (uint)strtoul(argv[2], nullptr, 10);
    uint n_threads = (uint)strtoul(argv[3], nullptr, 10);
    uint seed = (uint)strtoul(argv[4], nullptr, 10);
    if (argc > 5) coalref_trace_open(argv[5], argc > 6 ? atol(argv[6]) : 1000000000L, &p.maps);

    Sample sample(p.maps, p.times, p.viral, std::pair<uint, uint>(p.LA, p.LB));
#ifdef COALREF_GPU_BINDING
    GpuImportanceSampler sampler;       // the one-line change of main.cpp:281
#else
    ImportanceSampler sampler;
#endif
    auto t0 = std::chrono::steady_clock::now();
    SamplerOutput<double> res = sampler.runSampler(&sample, p.mp, n_samples, n_threads, 16, seed);
    auto t1 = std::chrono::steady_clock::now();
    double secs = std::chrono::duration<double>(t1 - t0).count();
    coalref_trace_close();

#ifdef COALREF_GPU_BINDING
    // the batch method of the binding on three copies of the pair: every output must equal the single call's
    bool batch_equal = true;
    {
        std::vector<Sample*> ss(3, &sample); std::vector<ModelParameters> ms(3, p.mp);
        std::vector<SamplerOutput<double> > bo = sampler.runSamplerBatch(ss, ms, n_samples, 16, seed);
        for (size_t b = 0; b < bo.size(); b++)
            for (uint i = 0; i < res.likelihood().dim1(); i++) for (uint j = 0; j < res.likelihood().dim2(); j++) for (uint k = 0; k < res.likelihood().dim3(); k++)
                batch_equal = batch_equal && bo[b].likelihood().getElement(i, j, k) == res.likelihood().getElement(i, j, k)
                                          && bo[b].tmrca().getElement(i, j, k) == res.tmrca().getElement(i, j, k);
        batch_equal = batch_equal && bo.size() == 3;
    }
    printf("{\n\"batch_equal\": %s,\n", batch_equal ? "true" : "false");
    const char* impl = "reference-binding-gpu";
#else
    const char* impl = "reference";
#endif
#ifndef COALREF_GPU_BINDING
    printf("{\n");
#endif
    printf("\"impl\": \"%s\", \"n_samples\": %u, \"n_threads\": %u, \"seed\": %u, \"seconds\": %.6f,\n", impl, n_samples, n_threads, seed, secs);
    printf("\"dims\": [%u, %u, %u], \"T\": %u,\n", res.likelihood().dim1(), res.likelihood().dim2(), res.likelihood().dim3(), p.T);
#ifdef COALREF_GPU_BINDING
    {   // the order in which the binding walked the reference's unordered_maps (the sampler's draws depend on it)
        std::vector<AncestralTypeMaps> maps = sample.getTypeMaps();
        printf("\"type_order\": [");
        for (size_t t = 0; t < maps.size(); t++) {
            printf("%s[", t ? ", " : "");
            bool first = true;
            TypeMap* ms[3] = {&maps[t].A, &maps[t].B, &maps[t].C};
            for (int g = 0; g < 3; g++) for (TypeMap::iterator it = ms[g]->begin(); it != ms[g]->end(); ++it) {
                std::string bits; for (size_t k = 0; k < it->first.size(); k++) bits += it->first[k] ? '1' : '0';
                printf("%s[\"%c\", \"%s\", %u]", first ? "" : ", ", "ABC"[g], bits.c_str(), it->second);
                first = false;
            }
            printf("]");
        }
        printf("],\n");
    }
#endif
    print_grid("likelihood", res.likelihood(), false);
    print_grid("likelihoodSq", res.likelihoodSq(), false);
    print_grid("tmrca", res.tmrca(), false);
    printf("\"lineages_remaining\": {\n");
    for (uint l = 0; l < p.T; l++) {
        char nm[32]; snprintf(nm, sizeof nm, "%u", l);
        print_grid(nm, res.lineages_remaining()[l], l + 1 == p.T);
    }
    printf("}\n}\n");
    return 0;
}
