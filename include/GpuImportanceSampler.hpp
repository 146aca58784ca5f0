// Reference-side binding: a drop-in for the reference's ImportanceSampler
// (/root/reference/src/ImportanceSampler.hpp:42-53) that forwards runSampler to libcoalgpu.so.
// This header is compiled against the REFERENCE's headers (Sample.hpp, SamplerOutput.hpp, Utils.hpp): a
// maintainer adds it to the reference tree, links -lcoalgpu and changes the type at main.cpp:281.
// It is exercised by oracle/_ref/coalref_gpu (oracle/Makefile) in tests/test_gpu_parity.py.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "coalgpu.h"
#include "Sample.hpp"
#include "SamplerOutput.hpp"
#include "Utils.hpp"

class GpuImportanceSampler {
  public:
    explicit GpuImportanceSampler(int device = 0) : device_(device) {}

    // same signature as ImportanceSampler::runSampler (ImportanceSampler.cpp:110); num_threads is ignored
    SamplerOutput<double> runSampler(Sample* sample, ModelParameters mp, uint num_iterations, uint /*num_threads*/,
                                     uint quadrature_points, uint seed) {
        Flat f(sample, mp);
        coalgpu_result r = f.result();
        const int rc = coalgpu_run_sampler(&f.p, num_iterations, quadrature_points, seed, device_, &r);
        if (rc) throw std::runtime_error(std::string("coalgpu: ") + coalgpu_last_error());               // caught at main.cpp:343-346
        return f.output();
    }

    // All locus pairs of main.cpp's double loop (:242-311) in one call: collect the sub-samples and their
    // ModelParameters (rho already scaled per pair, :260-271) first, then write the rows in the original order.
    std::vector<SamplerOutput<double> > runSamplerBatch(const std::vector<Sample*>& samples, const std::vector<ModelParameters>& mps,
                                                        uint num_iterations, uint quadrature_points, uint seed) {
        if (samples.size() != mps.size()) throw std::runtime_error("coalgpu: one ModelParameters per Sample");
        std::vector<Flat*> flats;
        std::vector<coalgpu_problem> probs; std::vector<coalgpu_result> results;
        std::vector<SamplerOutput<double> > out;
        try {
            for (size_t k = 0; k < samples.size(); k++) { flats.push_back(new Flat(samples[k], mps[k])); }
            for (size_t k = 0; k < flats.size(); k++) { probs.push_back(flats[k]->p); results.push_back(flats[k]->result()); }
            const int rc = coalgpu_run_sampler_batch((int32_t)probs.size(), probs.data(), num_iterations, quadrature_points, seed, device_, results.data());
            if (rc) throw std::runtime_error(std::string("coalgpu: ") + coalgpu_last_error());
            for (size_t k = 0; k < flats.size(); k++) out.push_back(flats[k]->output());
        } catch (...) { for (size_t k = 0; k < flats.size(); k++) delete flats[k]; throw; }
        for (size_t k = 0; k < flats.size(); k++) delete flats[k];
        return out;
    }

  private:
    int device_;

    // Sample + ModelParameters flattened into a coalgpu_problem, and the buffers of its result
    struct Flat {
        std::vector<double> times, viral, mu, rho, lambda;
        std::vector<int32_t> offs;
        std::vector<uint8_t> gamma; std::vector<uint64_t> words; std::vector<uint32_t> counts;
        std::vector<double> like, likesq, tmrca, lin;
        coalgpu_problem p;
        Flat(Sample* sample, const ModelParameters& mp) : offs(1, 0), p() {
            std::vector<AncestralTypeMaps> maps = sample->getTypeMaps();          // Sample.hpp:56
            times = sample->getTimes(); viral = sample->getViralLoads();
            mu = mp.target_mu; rho = mp.target_rho; lambda = mp.target_lambda;
            const std::pair<uint, uint> lens = sample->getSeqLens();
            // the three TypeMaps of every sampling time; locus A site s -> bit s, locus B site s -> bit LA+s
            for (size_t t = 0; t < maps.size(); t++) {
                for (TypeMap::iterator it = maps[t].A.begin(); it != maps[t].A.end(); ++it) push(COALGPU_GAMMA_A, it->first, 0, it->second);
                for (TypeMap::iterator it = maps[t].B.begin(); it != maps[t].B.end(); ++it) push(COALGPU_GAMMA_B, it->first, lens.first, it->second);
                for (TypeMap::iterator it = maps[t].C.begin(); it != maps[t].C.end(); ++it) push(COALGPU_GAMMA_C, it->first, 0, it->second);
                offs.push_back((int32_t)gamma.size());
            }
            p.T = (int32_t)times.size(); p.LA = (int32_t)lens.first; p.LB = (int32_t)lens.second;
            p.times = times.data(); p.viral = viral.data(); p.tau = mp.tau;
            p.driving_mu = mp.driving_mu; p.driving_lambda = mp.driving_lambda; p.driving_rho = mp.driving_rho;   // rho already x (distance+1), main.cpp:260-271
            p.n_mu = (int32_t)mu.size(); p.n_rho = (int32_t)rho.size(); p.n_lambda = (int32_t)lambda.size();
            p.mu = mu.data(); p.rho = rho.data(); p.lambda = lambda.data();
            p.type_offsets = offs.data(); p.type_gamma = gamma.data(); p.type_words = words.data(); p.type_counts = counts.data();
            const size_t G = (size_t)p.n_mu * p.n_rho * p.n_lambda;
            like.resize(G); likesq.resize(G); tmrca.resize(G); lin.resize(G * p.T);
        }
        coalgpu_result result() {
            coalgpu_result r = coalgpu_result();
            r.likelihood = like.data(); r.likelihood_sq = likesq.data(); r.tmrca = tmrca.data(); r.lineages = lin.data();
            return r;
        }
        SamplerOutput<double> output() {
            SamplerOutput<double> out(p.n_mu, p.n_rho, p.n_lambda, 0, p.T);                              // ImportanceSampler.cpp:141
            const size_t G = like.size();
            size_t g = 0;
            for (int i = 0; i < p.n_mu; i++) for (int j = 0; j < p.n_rho; j++) for (int k = 0; k < p.n_lambda; k++, g++) {
                out.likelihood().addValueToElement(like[g], i, j, k);                                    // Matrix3d.cpp:67
                out.likelihoodSq().addValueToElement(likesq[g], i, j, k);
                out.tmrca().addValueToElement(tmrca[g], i, j, k);
                for (int l = 0; l < p.T; l++) out.lineages_remaining()[l].addValueToElement(lin[(size_t)l * G + g], i, j, k);
            }
            return out;
        }
        void push(uint8_t g, const std::vector<bool>& bits, uint shift, uint count) {
            uint64_t w = 0;
            for (size_t s = 0; s < bits.size(); s++) if (bits[s]) w |= (uint64_t)1 << (shift + s);
            gamma.push_back(g); words.push_back(w); counts.push_back(count);
        }
      private:
        Flat(const Flat&); Flat& operator=(const Flat&);     // p points into the vectors
    };
};
