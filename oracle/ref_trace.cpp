// TEST INFRASTRUCTURE (oracle/): call tracer for the UNMODIFIED reference objects, linked with
// `ld --wrap` so that no reference source is touched. It intercepts the cross-translation-unit calls
//   ProposalEngine::calculateLikelihood   (ProposalEngine.cpp:643)   -> history boundaries
//   TypeContainer::sampleType             (TypeContainer.cpp:641)    -> chosen lineage per event
//   TypeContainer::addType / removeType   (TypeContainer.cpp:409,459)-> net state change per event
//   TypeContainer::addTypeMap             (TypeContainer.cpp:525)    -> epoch injection
//   TypeContainer::getDifferences         (TypeContainer.cpp:748)    -> class vectors fed to the HMM
//   ConditionalSamplingHMM::piSMC         (ConditionalSamplingHMM.cpp:265) -> HMM value
//   SamplerOutput<...>::addSample         (SamplerOutput.cpp:76)     -> per-history grid weights
// and writes one text line per record (format in oracle/README.md). Single-threaded runs only.

#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "Utils.hpp"
#include "TypeContainer.hpp"
#include "ConditionalSamplingHMM.hpp"
#include "ProposalEngine.hpp"
#include "SamplerOutput.hpp"
#include "Matrix3d.hpp"

typedef SamplerOutput<LikelihoodSamplesType> SOut;

#define SYM_PISMC    "_ZN22ConditionalSamplingHMM5piSMCER9HaplotypeR13TypeContainerj"
#define SYM_GETDIFF  "_ZN13TypeContainer14getDifferencesER9Haplotype"
#define SYM_SAMPLE   "_ZN13TypeContainer10sampleTypeEddd"
#define SYM_ADD      "_ZN13TypeContainer7addTypeER9Haplotype"
#define SYM_REMOVE   "_ZN13TypeContainer10removeTypeER9Haplotype"
#define SYM_ADDMAP   "_ZN13TypeContainer10addTypeMapER17AncestralTypeMaps"
#define SYM_ADDSAMP  "_ZN13SamplerOutputI20LikelihoodSamplesSumE9addSampleER8Matrix3dIdEdRSt6vectorIdSaIdEE"
#define SYM_CALC     "_ZN14ProposalEngine19calculateLikelihoodEP13SamplerOutputI20LikelihoodSamplesSumE15ModelParametersSt6vectorI17AncestralTypeMapsSaIS6_EES5_IdSaIdEESA_St4pairIjjE"

double __real_piSMC(ConditionalSamplingHMM*, Haplotype&, TypeContainer&, uint) asm("__real_" SYM_PISMC);
double __wrap_piSMC(ConditionalSamplingHMM*, Haplotype&, TypeContainer&, uint) asm("__wrap_" SYM_PISMC);
vector<vector<int> > __real_getDiff(TypeContainer*, Haplotype&) asm("__real_" SYM_GETDIFF);
vector<vector<int> > __wrap_getDiff(TypeContainer*, Haplotype&) asm("__wrap_" SYM_GETDIFF);
pair<Haplotype, vector<double> > __real_sample(TypeContainer*, double, double, double) asm("__real_" SYM_SAMPLE);
pair<Haplotype, vector<double> > __wrap_sample(TypeContainer*, double, double, double) asm("__wrap_" SYM_SAMPLE);
void __real_add(TypeContainer*, Haplotype&) asm("__real_" SYM_ADD);
void __wrap_add(TypeContainer*, Haplotype&) asm("__wrap_" SYM_ADD);
void __real_remove(TypeContainer*, Haplotype&) asm("__real_" SYM_REMOVE);
void __wrap_remove(TypeContainer*, Haplotype&) asm("__wrap_" SYM_REMOVE);
void __real_addmap(TypeContainer*, AncestralTypeMaps&) asm("__real_" SYM_ADDMAP);
void __wrap_addmap(TypeContainer*, AncestralTypeMaps&) asm("__wrap_" SYM_ADDMAP);
void __real_addsamp(SOut*, Matrix3d<double>&, double, vector<double>&) asm("__real_" SYM_ADDSAMP);
void __wrap_addsamp(SOut*, Matrix3d<double>&, double, vector<double>&) asm("__wrap_" SYM_ADDSAMP);
uint __real_calc(ProposalEngine*, SOut*, ModelParameters, vector<AncestralTypeMaps>, vector<double>, vector<double>, pair<uint, uint>) asm("__real_" SYM_CALC);
uint __wrap_calc(ProposalEngine*, SOut*, ModelParameters, vector<AncestralTypeMaps>, vector<double>, vector<double>, pair<uint, uint>) asm("__wrap_" SYM_CALC);

namespace {
FILE* g_out = nullptr;
long g_max_hist = 0;
long g_hist = -1;           // index of the history being drawn
bool g_on = false;          // tracing active for this history
int g_epoch = 0;
std::map<std::string, int> g_delta;   // net state change since the last sampleType call
vector<vector<int> > g_last_diff;
bool g_have_diff = false;

std::string key_of(const Haplotype& h) {
    std::string s = h.gamma + ":";
    for (size_t i = 0; i < h.type.size(); i++) s.push_back(h.type[i] ? '1' : '0');
    return s;
}
void flush_delta() {
    bool any = false;
    for (auto& kv : g_delta) if (kv.second != 0) any = true;
    if (any) {
        fprintf(g_out, "D %ld", g_hist);
        for (auto& kv : g_delta) if (kv.second != 0) fprintf(g_out, " %+d%s", kv.second, kv.first.c_str());
        fprintf(g_out, "\n");
    }
    g_delta.clear();
}
}

void coalref_trace_open(const char* path, long max_histories, const std::vector<AncestralTypeMaps>*) {
    g_out = fopen(path, "w");
    if (!g_out) { fprintf(stderr, "coalref_trace: cannot open %s\n", path); exit(2); }
    g_max_hist = max_histories;
}
void coalref_trace_close() { if (g_out) { fclose(g_out); g_out = nullptr; } }

uint __wrap_calc(ProposalEngine* self, SOut* out, ModelParameters mp, vector<AncestralTypeMaps> td, vector<double> times, vector<double> viral, pair<uint, uint> lens) {
    g_hist++;
    g_on = (g_out != nullptr) && (g_hist < g_max_hist);
    g_epoch = 0;
    g_delta.clear();
    if (g_on) fprintf(g_out, "B %ld\n", g_hist);
    return __real_calc(self, out, mp, td, times, viral, lens);
}

vector<vector<int> > __wrap_getDiff(TypeContainer* self, Haplotype& h) {
    vector<vector<int> > d = __real_getDiff(self, h);
    if (g_on) { g_last_diff = d; g_have_diff = true; }
    return d;
}

double __wrap_piSMC(ConditionalSamplingHMM* self, Haplotype& h, TypeContainer& H, uint c) {
    g_have_diff = false;
    double pi = __real_piSMC(self, h, H, c);
    if (g_on && g_have_diff) {
        const vector<int>& cnt = g_last_diff[0];
        fprintf(g_out, "P %s %u %u %u %u %zu", h.gamma.c_str(), H.getLineageCount(), c, H.getLocusLengthA(), H.getLocusLengthB(), cnt.size());
        for (size_t i = 0; i < cnt.size(); i++) fprintf(g_out, " %d %d %d", cnt[i], g_last_diff[1][i], g_last_diff[2][i]);
        fprintf(g_out, " %.17g\n", pi);
    }
    return pi;
}

pair<Haplotype, vector<double> > __wrap_sample(TypeContainer* self, double ta, double tb, double rho) {
    if (g_on) flush_delta();
    pair<Haplotype, vector<double> > r = __real_sample(self, ta, tb, rho);
    if (g_on) fprintf(g_out, "S %ld %d %s %.17g %.17g\n", g_hist, g_epoch, key_of(r.first).c_str(), r.second[0], r.second[1]);
    return r;
}

void __wrap_add(TypeContainer* self, Haplotype& h) {
    if (g_on) g_delta[key_of(h)] += 1;
    __real_add(self, h);
}
void __wrap_remove(TypeContainer* self, Haplotype& h) {
    if (g_on) g_delta[key_of(h)] -= 1;
    __real_remove(self, h);
}
void __wrap_addmap(TypeContainer* self, AncestralTypeMaps& m) {
    if (g_on) { flush_delta(); g_epoch++; fprintf(g_out, "M %ld %d\n", g_hist, g_epoch); }
    __real_addmap(self, m);
}
void __wrap_addsamp(SOut* self, Matrix3d<double>& L, double tmrca, vector<double>& left) {
    if (g_on) {
        flush_delta();
        fprintf(g_out, "W %ld %.17g %zu", g_hist, tmrca, left.size());
        for (double v : left) fprintf(g_out, " %.17g", v);
        fprintf(g_out, " %u", L.dim1() * L.dim2() * L.dim3());
        for (uint i = 0; i < L.dim1(); i++) for (uint j = 0; j < L.dim2(); j++) for (uint k = 0; k < L.dim3(); k++) fprintf(g_out, " %.17g", L.getElement(i, j, k));
        fprintf(g_out, "\n");
    }
    __real_addsamp(self, L, tmrca, left);
}
