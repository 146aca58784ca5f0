// coalescenator-gpu — C++ host driver with the reference's command line and output files
// (/root/reference/src/main.cpp:48-354), calling the B200 path through the C ABI of include/coalgpu.h
// where the reference calls ImportanceSampler::runSampler (main.cpp:281-283).
//
// Same options as the reference (main.cpp:54-88); additions: --device N, --gpus N (locus pairs dealt round robin to
// devices device..device+N-1, one host thread per device, no exchange needed: pairs are independent), --dump-problems DIR (write one
// .coalprob file per eligible locus pair and stop: needs no GPU), --ess (append an ESS column), --summary (side-car
// <prefix>_summary.txt: the grid point that maximises the composite likelihood, per-pair work and ESS, throughput).
#include <chrono>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/coalgpu.h"
#include "fasta_input.h"

using namespace coalhost;

namespace {

struct Options {
    unsigned num_threads = 1, seed = 0, loci_size = 1, mc_samples = 500000, min_dist = 1, max_dist = 400;
    std::string prefix = "coalescenator", fasta_list, time_list, viral_list, coordinate_list;
    std::string mutation_list = "1e-6,5e-6,1e-5,2.5e-5,5e-5,7.5e-5,1e-4";
    std::string scale_list = "250,500,750,1000,2000,4000";
    std::string rho_list = "1e-10,1e-9,1e-8,1e-7,1e-6,5e-6,1e-5,5e-5,1e-4,5e-4,1e-3";
    double generation_time = 0, mutation_driving = 2.5e-5, scale_driving = 1000, rho_driving = 5e-5;
    bool have_coords = false, have_generation_time = false, ess = false, summary = false;
    int device = 0, gpus = 1;
    bool shard_samples = false;   // --shard-samples: with --gpus N, split the importance samples of every pair over the GPUs (NCCL combine)
    std::string dump_dir;
};

std::string now() {   // Utils.hpp:211-265: dd/mm/yyyy hh:mm:ss
    time_t t = time(nullptr); struct tm lt; localtime_r(&t, &lt);
    char buf[32]; strftime(buf, sizeof buf, "%d/%m/%Y %H:%M:%S", &lt);
    return buf;
}

void usage() {
    std::cout << "\n#### Coalescenator (B200) ####\n"
              << "  -h [ --help ]  -p [ --num-threads ] N (ignored)  -s [ --seed ] N  --output-file-prefix S  --device N  --gpus N\n"
              << "  -f [ --fasta-list ] a.fa,b.fa  -t [ --time-list ] t1,t2  -v [ --viral-list ] v1,v2  -g [ --generation-time ] X\n"
              << "  --coordinate-list [l-r],... | bp1,bp2   --loci-size N (1)\n"
              << "  -i [ --mc-samples ] N (500000)  --mutation-list L  --mutation-driving X (2.5e-5)  --scale-list L  --scale-driving X (1000)\n"
              << "  --rho-list L  --rho-driving X (5e-5)  --min-comp-dist N (1)  --max-comp-dist N (400)\n"
              << "  --dump-problems DIR  --ess  --summary  --shard-samples (with --gpus N: every pair's samples over all GPUs, NCCL combine)\n";
}

Options parse(int argc, char** argv) {
    Options o; o.seed = (unsigned)time(nullptr);
    std::map<std::string, std::string> alias = {{"-p", "--num-threads"}, {"-s", "--seed"}, {"-f", "--fasta-list"}, {"-t", "--time-list"},
                                                {"-v", "--viral-list"}, {"-g", "--generation-time"}, {"-i", "--mc-samples"}, {"-h", "--help"}};
    for (int i = 1; i < argc; i++) {
        std::string k = argv[i], v;
        const size_t eq = k.find('=');
        if (k.rfind("--", 0) == 0 && eq != std::string::npos) { v = k.substr(eq + 1); k = k.substr(0, eq); }
        if (alias.count(k)) k = alias[k];
        if (k == "--help") { usage(); exit(1); }
        if (k == "--ess") { o.ess = true; continue; }
        if (k == "--summary") { o.summary = true; continue; }
        if (k == "--shard-samples") { o.shard_samples = true; continue; }
        if (v.empty() && eq == std::string::npos) { if (i + 1 >= argc) throw std::runtime_error("the required argument for option '" + k + "' is missing"); v = argv[++i]; }
        if (k == "--num-threads") o.num_threads = (unsigned)std::stoul(v);
        else if (k == "--seed") o.seed = (unsigned)std::stoul(v);
        else if (k == "--output-file-prefix") o.prefix = v;
        else if (k == "--fasta-list") o.fasta_list = v;
        else if (k == "--time-list") o.time_list = v;
        else if (k == "--viral-list") o.viral_list = v;
        else if (k == "--generation-time") { o.generation_time = std::stod(v); o.have_generation_time = true; }
        else if (k == "--coordinate-list") { o.coordinate_list = v; o.have_coords = true; }
        else if (k == "--loci-size") o.loci_size = (unsigned)std::stoul(v);
        else if (k == "--mc-samples") o.mc_samples = (unsigned)std::stoul(v);
        else if (k == "--mutation-list") o.mutation_list = v;
        else if (k == "--mutation-driving") o.mutation_driving = std::stod(v);
        else if (k == "--scale-list") o.scale_list = v;
        else if (k == "--scale-driving") o.scale_driving = std::stod(v);
        else if (k == "--rho-list") o.rho_list = v;
        else if (k == "--rho-driving") o.rho_driving = std::stod(v);
        else if (k == "--min-comp-dist") o.min_dist = (unsigned)std::stoul(v);
        else if (k == "--max-comp-dist") o.max_dist = (unsigned)std::stoul(v);
        else if (k == "--device") o.device = std::stoi(v);
        else if (k == "--gpus") o.gpus = std::max(1, std::stoi(v));
        else if (k == "--dump-problems") o.dump_dir = v;
        else throw std::runtime_error("unrecognised option '" + k + "'");
    }
    if (argc == 1) { usage(); exit(1); }
    for (auto& req : {std::make_pair("--fasta-list", &o.fasta_list), std::make_pair("--time-list", &o.time_list), std::make_pair("--viral-list", &o.viral_list)})
        if (req.second->empty()) throw std::runtime_error(std::string("the option '") + req.first + "' is required but missing");
    if (!o.have_generation_time) throw std::runtime_error("the option '--generation-time' is required but missing");
    return o;
}

std::vector<double> number_list(const std::string& s) {
    std::vector<double> v;
    for (const std::string& t : split(s, ",")) v.push_back(std::atof(t.c_str()));
    return v;
}

double log_addition(double a, double b) {   // Utils.hpp:88-107
    return a < b ? b + std::log(1 + std::exp(a - b)) : a + std::log(1 + std::exp(b - a));
}

std::string bits_of(const PairProblem& p, int g, uint64_t w) {
    const unsigned off = g == 1 ? p.LA : 0, len = g == 0 ? p.LA : (g == 1 ? p.LB : p.LA + p.LB);
    std::string s;
    for (unsigned i = 0; i < len; i++) s.push_back(((w >> (off + i)) & 1) ? '1' : '0');
    return s;
}

void dump_problem(const std::string& path, const PairProblem& p, double tau, double dmu, double dlambda, double drho,
                  const std::vector<double>& mu, const std::vector<double>& rho, const std::vector<double>& lambda) {
    std::ofstream f(path.c_str());
    f << std::setprecision(17);
    f << "coalprob 1\nT " << p.times.size() << " LA " << p.LA << " LB " << p.LB << "\ntimes";
    for (double x : p.times) f << " " << x;
    f << "\nviral";
    for (double x : p.viral) f << " " << x;
    f << "\ntau " << tau << "\ndriving " << dmu << " " << dlambda << " " << drho << "\n";
    f << "mu " << mu.size(); for (double x : mu) f << " " << x;
    f << "\nrho " << rho.size(); for (double x : rho) f << " " << x;
    f << "\nlambda " << lambda.size(); for (double x : lambda) f << " " << x;
    f << "\n";
    for (size_t t = 0; t < p.times.size(); t++) {
        f << "types " << t << " " << (p.type_offsets[t + 1] - p.type_offsets[t]) << "\n";
        for (int k = p.type_offsets[t]; k < p.type_offsets[t + 1]; k++)
            f << "ABC"[p.type_gamma[k]] << " " << bits_of(p, p.type_gamma[k], p.type_words[k]) << " " << p.type_counts[k] << "\n";
    }
}

}  // namespace

int main(int argc, char** argv) {
    try {
        Options o = parse(argc, argv);
        std::cout << "\n[" << now() << "] Seed used for pseudo-random generator: " << o.seed << "\n" << std::endl;

        Alignment aln(o.fasta_list, o.time_list, o.viral_list);                     // main.cpp:107
        if (o.have_coords) aln.parse_coordinates(o.coordinate_list); else aln.tile_loci(o.loci_size);   // :109-117
        const size_t s1 = aln.filter_ambiguous(), s2 = aln.filter_indels(), s3 = aln.filter_non_biallelic();   // :119-121
        for (const std::string& m : aln.messages) std::cout << "[" << now() << "] " << m << std::endl;
        std::cout << "[" << now() << "] " << s1 << " sites remaining after filtering ambiguous sites, " << s2 << " after filtering indels, "
                  << s3 << " after filtering non-bi-allelic positions." << std::endl;

        const std::vector<Locus>& loci = aln.loci();
        const std::vector<Locus>& orig = aln.original_loci();
        if (loci.empty()) { std::cout << "\n[" << now() << "] All loci was removed by the filtering process - exiting!" << std::endl; return 0; }   // :126-130
        std::cout << "\n[" << now() << "] Number of loci: " << loci.size() << std::endl;
        std::cout << "[" << now() << "] List of loci lengths: \n\n\t" << loci[0].second - loci[0].first + 1;
        for (size_t i = 1; i < loci.size(); i++) std::cout << ", " << loci[i].second - loci[i].first + 1;
        std::cout << std::endl;
        if (!(o.mutation_driving > 0) || !(o.scale_driving > 0) || !(o.rho_driving > 0) || !(o.generation_time > 0))
            throw std::runtime_error("driving values and generation time must be positive");   // :147-150

        const std::vector<double> mu_list = number_list(o.mutation_list), lambda_list = number_list(o.scale_list), rho_list = number_list(o.rho_list);
        // the reference asserts mu >= 0, lambda > 0, rho >= 0 (main.cpp:169,175,181)
        for (double v : mu_list) if (!(v >= 0)) throw std::runtime_error("--mutation-list: values must be >= 0");
        for (double v : lambda_list) if (!(v > 0)) throw std::runtime_error("--scale-list: values must be > 0");
        for (double v : rho_list) if (!(v >= 0)) throw std::runtime_error("--rho-list: values must be >= 0");
        if (loci.size() == 1) { std::cout << "\n[" << now() << "] Only a single loci left no remcobination estimation possible - exiting!" << std::endl; return 0; }   // :202-207

        struct Pair { size_t i, j; unsigned dist; };
        size_t n_skipped_long = 0;
        std::vector<Pair> pairs;                                                       // :211-225, :242-252
        for (size_t i = 0; i < loci.size(); i++) for (size_t j = i + 1; j < loci.size(); j++) {
            const unsigned dist = orig[j].first - orig[i].second - 1;
            if (dist < o.min_dist || dist > o.max_dist) continue;
            // libcoalgpu.so packs the two loci of a pair into one 64-bit haplotype word; a pair with more segregating sites is
            // reported and left out instead of failing the whole analysis (the reference itself loses accuracy beyond ~30
            // segregating sites per locus and cannot run beyond 1029, ConditionalSamplingHMM.cpp:73-82,196-209)
            const size_t lsum = (loci[i].second - loci[i].first + 1) + (loci[j].second - loci[j].first + 1);
            if (lsum > 64) {
                std::cout << "[" << now() << "] WARNING: skipping locus pair (" << orig[i].first << ", " << orig[i].second << ") and (" << orig[j].first << ", "
                          << orig[j].second << "): " << lsum << " segregating sites exceed the 64 a packed locus pair holds (use a smaller --loci-size)." << std::endl;
                n_skipped_long++;
                continue;
            }
            pairs.push_back({i, j, dist});
        }
        std::cout << "\n[" << now() << "] For each of the " << pairs.size() << " loci combination(s) generate " << o.mc_samples << " importance samples. "
                  << (unsigned long long)o.mc_samples * pairs.size() << " samples in total.\n" << std::endl;

        const size_t nmu = mu_list.size(), nrho = rho_list.size(), nlam = lambda_list.size(), G = nmu * nrho * nlam;
        std::vector<double> composite(G, 0.0), mean_tmrca(G, 0.0);                    // :229-230 (mean_tmrca starts at 0 in the log domain)

        std::ofstream pairwise;
        if (o.dump_dir.empty()) {
            pairwise.open((o.prefix + "_pairwise_loglikelihood.txt").c_str());
            if (!pairwise.is_open()) throw std::runtime_error("cannot open the pairwise output file");
            pairwise << "locus1\tlocus2\tmu\trho\tlambda\tloglikelihood\ttmrca\tlineages_remaining\tloglikelihoodSq" << (o.ess ? "\tess" : "") << "\n";   // :240
        }

        // Sub-samples of all eligible pairs first (main.cpp:254-279), then ONE batched call: the pairs run
        // concurrently on the GPU instead of one after another (coalgpu_run_sampler_batch), rows are written in
        // the reference's order afterwards.
        struct PairStat { unsigned long long events, pismc; double ess_min, ess_max; };
        std::vector<PairStat> pair_stats;
        double sampling_seconds = 0.0;
        struct Job { PairProblem sp; std::vector<double> rho_scaled; double drho; std::string l1, l2;
                     std::vector<double> like, likesq, tm, lin, ess; };
        std::vector<Job> jobs(pairs.size());
        size_t remaining = pairs.size();
        for (size_t q = 0; q < pairs.size(); q++) {
            const Pair& pr = pairs[q];
            Job& j = jobs[q];
            j.sp = aln.sub_sample(loci[pr.i], loci[pr.j]);                            // :254
            j.rho_scaled = rho_list;
            for (double& r : j.rho_scaled) r *= pr.dist + 1;                          // :260-271
            j.drho = o.rho_driving * (pr.dist + 1);
            remaining--;
            std::ostringstream l1, l2;
            l1 << "(" << orig[pr.i].first << ", " << orig[pr.i].second << ")";
            l2 << "(" << orig[pr.j].first << ", " << orig[pr.j].second << ")";
            j.l1 = l1.str(); j.l2 = l2.str();
            std::cout << "[" << now() << "] Working on locus " << j.l1 << " and " << j.l2 << " having " << j.sp.nA << ", " << j.sp.nB << " and " << j.sp.nC
                      << " sequences with ancentry A, B and C respectively taken at " << j.sp.times.size() << " different time points. " << remaining
                      << " remaining pairwise combinations." << std::endl;
            if (!o.dump_dir.empty()) {
                std::ostringstream path;
                path << o.dump_dir << "/pair_" << orig[pr.i].first << "-" << orig[pr.i].second << "_" << orig[pr.j].first << "-" << orig[pr.j].second << ".coalprob";
                dump_problem(path.str(), j.sp, o.generation_time, o.mutation_driving, o.scale_driving, j.drho, mu_list, j.rho_scaled, lambda_list);
            }
        }
        if (o.dump_dir.empty()) {
            std::vector<coalgpu_problem> probs(jobs.size());
            std::vector<coalgpu_result> results(jobs.size());
            for (size_t q = 0; q < jobs.size(); q++) {
                Job& j = jobs[q];
                const int T = (int)j.sp.times.size();
                coalgpu_problem p = {};
                p.T = T; p.LA = (int)j.sp.LA; p.LB = (int)j.sp.LB;
                p.times = j.sp.times.data(); p.viral = j.sp.viral.data(); p.tau = o.generation_time;
                p.driving_mu = o.mutation_driving; p.driving_lambda = o.scale_driving; p.driving_rho = j.drho;
                p.n_mu = (int)nmu; p.n_rho = (int)nrho; p.n_lambda = (int)nlam;
                p.mu = mu_list.data(); p.rho = j.rho_scaled.data(); p.lambda = lambda_list.data();
                p.type_offsets = j.sp.type_offsets.data(); p.type_gamma = j.sp.type_gamma.data();
                p.type_words = j.sp.type_words.data(); p.type_counts = j.sp.type_counts.data();
                probs[q] = p;
                j.like.resize(G); j.likesq.resize(G); j.tm.resize(G); j.lin.resize(G * T); j.ess.resize(G);
                results[q] = coalgpu_result{j.like.data(), j.likesq.data(), j.tm.data(), j.lin.data(), j.ess.data(), 0, 0, 0, 0.0};
            }
            const auto tb0 = std::chrono::steady_clock::now();
            if (o.gpus <= 1) {
                const int rc = coalgpu_run_sampler_batch((int)probs.size(), probs.data(), o.mc_samples, 16, o.seed, o.device, results.data());   // replaces :281-283
                if (rc) throw std::runtime_error(std::string("coalgpu: ") + coalgpu_last_error());
            } else if (probs.size() < (size_t)o.gpus || o.shard_samples || (double)o.mc_samples * (double)(G + probs[0].T + 1) * 8.0 > 256.0 * 1048576.0) {
                // Fewer pairs than GPUs, or so many importance samples per pair that one pair fills a GPU for a long time
                // (the reference's default is 500,000, main.cpp:75): the SAMPLES of each pair are sharded over the GPUs --
                // the thread fan-out of ImportanceSampler.cpp:125-139 with GPUs in the place of threads, one NCCL
                // allreduce of the (max, sum) pairs per locus pair (coalgpu_run_sampler_multi). Same histories as one
                // GPU would draw, so the files agree with the single-GPU run up to the rounding of the merge.
                std::vector<int> devs;
                for (int d = 0; d < o.gpus; d++) devs.push_back(o.device + d);
                for (size_t q = 0; q < probs.size(); q++)
                    if (coalgpu_run_sampler_multi(&probs[q], o.mc_samples, 16, o.seed, o.gpus, devs.data(), &results[q]))
                        throw std::runtime_error(std::string("coalgpu: ") + coalgpu_last_error());
                coalgpu_multi_shutdown();
            } else {
                // pairs are independent: device d takes pairs d, d+gpus, ... and runs its own batch; a pair's histories
                // depend on (seed, index) only, so the output does not depend on the number of GPUs
                std::vector<std::vector<coalgpu_problem>> dp(o.gpus);
                std::vector<std::vector<coalgpu_result>> dr(o.gpus);
                std::vector<std::vector<size_t>> di(o.gpus);
                for (size_t q = 0; q < probs.size(); q++) { const int d = (int)(q % o.gpus); dp[d].push_back(probs[q]); dr[d].push_back(results[q]); di[d].push_back(q); }
                std::vector<std::string> errs(o.gpus);
                std::vector<std::thread> th;
                for (int d = 0; d < o.gpus; d++)
                    th.emplace_back([&, d]() {
                        if (dp[d].empty()) return;
                        if (coalgpu_run_sampler_batch((int)dp[d].size(), dp[d].data(), o.mc_samples, 16, o.seed, o.device + d, dr[d].data()))
                            errs[d] = coalgpu_last_error();      // thread-local in the library: read on the thread that failed
                    });
                for (std::thread& t : th) t.join();
                for (int d = 0; d < o.gpus; d++) {
                    if (!errs[d].empty()) throw std::runtime_error("coalgpu (device " + std::to_string(o.device + d) + "): " + errs[d]);
                    for (size_t k = 0; k < di[d].size(); k++) results[di[d][k]] = dr[d][k];
                }
            }
            sampling_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tb0).count();
            for (size_t q = 0; q < jobs.size(); q++) {
                PairStat st;
                st.events = results[q].n_events; st.pismc = results[q].n_pismc;
                st.ess_min = *std::min_element(jobs[q].ess.begin(), jobs[q].ess.end());
                st.ess_max = *std::max_element(jobs[q].ess.begin(), jobs[q].ess.end());
                pair_stats.push_back(st);
            }
            for (size_t q = 0; q < jobs.size(); q++) {
                const Job& j = jobs[q];
                const int T = (int)j.sp.times.size();
                size_t g = 0;
                for (size_t i = 0; i < nmu; i++) for (size_t jj = 0; jj < nrho; jj++) for (size_t k = 0; k < nlam; k++, g++) {   // :285-310
                    composite[g] += j.like[g];
                    mean_tmrca[g] = log_addition(mean_tmrca[g], j.tm[g]);
                    pairwise << j.l1 << "\t" << j.l2 << "\t" << mu_list[i] << "\t" << rho_list[jj] << "\t" << lambda_list[k] << "\t" << j.like[g] << "\t" << std::exp(j.tm[g]) << "\t";
                    for (int l = 0; l < T; l++) pairwise << (l ? "," : "") << std::exp(j.lin[(size_t)l * G + g]);
                    pairwise << "\t" << j.likesq[g];
                    if (o.ess) pairwise << "\t" << j.ess[g];
                    pairwise << "\n";
                }
            }
        }
        if (!o.dump_dir.empty()) { std::cout << "[" << now() << "] Wrote " << pairs.size() << " problem file(s) to " << o.dump_dir << std::endl; return 0; }
        pairwise.close();

        std::ofstream comp((o.prefix + "_composite_loglikelihood.txt").c_str());     // :317-340
        if (!comp.is_open()) throw std::runtime_error("cannot open the composite output file");
        const double norm = std::log((double)(loci.size() * (loci.size() - 1) / 2));  // :325 (all locus pairs, filtered or not)
        comp << "mu\trho\tlambda\tlikelihood\ttmrca\n";
        size_t g = 0;
        for (size_t i = 0; i < nmu; i++) for (size_t j = 0; j < nrho; j++) for (size_t k = 0; k < nlam; k++, g++)
            comp << mu_list[i] << "\t" << rho_list[j] << "\t" << lambda_list[k] << "\t" << composite[g] << "\t" << std::exp(mean_tmrca[g] - norm) << "\n";
        comp.close();

        if (o.summary) {        // not in the reference: diagnostics next to the two result files
            std::ofstream sm((o.prefix + "_summary.txt").c_str());
            if (!sm.is_open()) throw std::runtime_error("cannot open the summary output file");
            size_t best = 0;
            for (size_t q = 1; q < G; q++) if (composite[q] > composite[best]) best = q;
            const size_t bi = best / (nrho * nlam), bj = (best / nlam) % nrho, bk = best % nlam;
            sm << "composite_maximum\tmu\t" << mu_list[bi] << "\trho\t" << rho_list[bj] << "\tlambda\t" << lambda_list[bk] << "\tloglikelihood\t" << composite[best]
               << "\ton_grid_edge\t" << (((nmu > 1 && (bi == 0 || bi + 1 == nmu)) || (nrho > 1 && (bj == 0 || bj + 1 == nrho)) || (nlam > 1 && (bk == 0 || bk + 1 == nlam))) ? 1 : 0) << "\n";
            sm << "pairs\t" << jobs.size() << "\tsamples_per_pair\t" << o.mc_samples << "\tsampling_seconds\t" << sampling_seconds << "\thistories_per_second\t"
               << (sampling_seconds > 0 ? (double)jobs.size() * (double)o.mc_samples / sampling_seconds : 0.0) << "\n";
            sm << "locus1\tlocus2\tevents_per_history\tpismc_per_history\tess_min\tess_max\n";
            for (size_t q = 0; q < jobs.size(); q++)
                sm << jobs[q].l1 << "\t" << jobs[q].l2 << "\t" << (double)pair_stats[q].events / (double)o.mc_samples << "\t" << (double)pair_stats[q].pismc / (double)o.mc_samples
                   << "\t" << pair_stats[q].ess_min << "\t" << pair_stats[q].ess_max << "\n";
        }
    } catch (std::exception& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;                              // :343-346
        return 1;
    }
    std::cout << "[" << now() << "] Coalescenated succesfully." << std::endl;
    return 0;
}
