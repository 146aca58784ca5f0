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
        std::vector<AncestralTypeMaps> maps = sample->getTypeMaps();          // Sample.hpp:56
        std::vector<double> times = sample->getTimes(), viral = sample->getViralLoads();
        const std::pair<uint, uint> lens = sample->getSeqLens();
        // flatten the three TypeMaps of every sampling time; locus A site s -> bit s, locus B site s -> bit LA+s
        std::vector<int32_t> offs(1, 0);
        std::vector<uint8_t> gamma; std::vector<uint64_t> words; std::vector<uint32_t> counts;
        for (size_t t = 0; t < maps.size(); t++) {
            for (TypeMap::iterator it = maps[t].A.begin(); it != maps[t].A.end(); ++it) push(gamma, words, counts, COALGPU_GAMMA_A, it->first, 0, it->second);
            for (TypeMap::iterator it = maps[t].B.begin(); it != maps[t].B.end(); ++it) push(gamma, words, counts, COALGPU_GAMMA_B, it->first, lens.first, it->second);
            for (TypeMap::iterator it = maps[t].C.begin(); it != maps[t].C.end(); ++it) push(gamma, words, counts, COALGPU_GAMMA_C, it->first, 0, it->second);
            offs.push_back((int32_t)gamma.size());
        }
        coalgpu_problem p = coalgpu_problem();
        p.T = (int32_t)times.size(); p.LA = (int32_t)lens.first; p.LB = (int32_t)lens.second;
        p.times = times.data(); p.viral = viral.data(); p.tau = mp.tau;
        p.driving_mu = mp.driving_mu; p.driving_lambda = mp.driving_lambda; p.driving_rho = mp.driving_rho;   // rho already x (distance+1), main.cpp:260-271
        p.n_mu = (int32_t)mp.target_mu.size(); p.n_rho = (int32_t)mp.target_rho.size(); p.n_lambda = (int32_t)mp.target_lambda.size();
        p.mu = mp.target_mu.data(); p.rho = mp.target_rho.data(); p.lambda = mp.target_lambda.data();
        p.type_offsets = offs.data(); p.type_gamma = gamma.data(); p.type_words = words.data(); p.type_counts = counts.data();

        const size_t G = (size_t)p.n_mu * p.n_rho * p.n_lambda;
        std::vector<double> like(G), likesq(G), tmrca(G), lin(G * p.T);
        coalgpu_result r = coalgpu_result();
        r.likelihood = like.data(); r.likelihood_sq = likesq.data(); r.tmrca = tmrca.data(); r.lineages = lin.data();
        const int rc = coalgpu_run_sampler(&p, num_iterations, quadrature_points, seed, device_, &r);
        if (rc) throw std::runtime_error(std::string("coalgpu: ") + coalgpu_last_error());               // caught at main.cpp:343-346

        SamplerOutput<double> out(p.n_mu, p.n_rho, p.n_lambda, 0, p.T);                                  // ImportanceSampler.cpp:141
        size_t g = 0;
        for (int i = 0; i < p.n_mu; i++) for (int j = 0; j < p.n_rho; j++) for (int k = 0; k < p.n_lambda; k++, g++) {
            out.likelihood().addValueToElement(like[g], i, j, k);                                        // Matrix3d.cpp:67
            out.likelihoodSq().addValueToElement(likesq[g], i, j, k);
            out.tmrca().addValueToElement(tmrca[g], i, j, k);
            for (int l = 0; l < p.T; l++) out.lineages_remaining()[l].addValueToElement(lin[(size_t)l * G + g], i, j, k);
        }
        return out;
    }

  private:
    int device_;
    static void push(std::vector<uint8_t>& gamma, std::vector<uint64_t>& words, std::vector<uint32_t>& counts, uint8_t g,
                     const std::vector<bool>& bits, uint shift, uint count) {
        uint64_t w = 0;
        for (size_t s = 0; s < bits.size(); s++) if (bits[s]) w |= (uint64_t)1 << (shift + s);
        gamma.push_back(g); words.push_back(w); counts.push_back(count);
    }
};
