/* coalgpu.h — C ABI of the B200 importance-sampling likelihood path.
 *
 * Drop-in boundary: the reference has no FFI layer; its seam is the C++ call
 *     SamplerOutput<double> ImportanceSampler::runSampler(Sample*, ModelParameters,
 *         uint num_iterations, uint num_threads, uint quadrature_points, uint seed);
 *     (/root/reference/src/ImportanceSampler.hpp:53, ImportanceSampler.cpp:110; called once per
 *      locus pair at main.cpp:281-283, results read at main.cpp:285-306)
 * `coalgpu_run_sampler` replaces exactly that call. Everything else here is the same work split
 * into stages so a caller can keep histories on the device, shard them over GPUs and combine the
 * per-GPU partial sums itself (coalescenator_b200/distributed.py does it with one NCCL exchange between
 * processes; coalgpu_run_sampler_multi does it inside this library for the GPUs of one process).
 *
 * Plain C types only. All functions return 0 on success or a negative COALGPU_E* code and never
 * abort; coalgpu_last_error() describes the last failure on the calling thread. The library is
 * CUDA-only: without a usable device every compute entry point fails with COALGPU_ENODEV — there
 * is no CPU fallback.
 */
#ifndef COALGPU_H
#define COALGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COALGPU_EINVAL   (-1)   /* bad argument / unsupported problem shape          */
#define COALGPU_ENODEV   (-2)   /* no CUDA device / wrong architecture               */
#define COALGPU_ECUDA    (-3)   /* CUDA runtime error (see coalgpu_last_error)       */
#define COALGPU_ENOMEM   (-4)
#define COALGPU_EOVERFLOW (-5)  /* a history outgrew the per-history state capacity  */
#define COALGPU_ENAN     (-6)   /* a history weight came out as NaN (never dropped silently) */
#define COALGPU_ENCCL    (-7)   /* NCCL missing or an NCCL call failed (multi-GPU entry points only) */

#define COALGPU_GAMMA_A 0       /* lineage ancestral at locus A only  (i,*) */
#define COALGPU_GAMMA_B 1       /* lineage ancestral at locus B only  (*,j) */
#define COALGPU_GAMMA_C 2       /* lineage ancestral at both loci     (i,j) */

/* One locus pair: the reference's Sample (Sample.hpp:44-63) + ModelParameters (Utils.hpp:157-166),
 * flattened. Haplotypes are bit-packed: locus A site s = bit s, locus B site s = bit LA+s
 * (LA+LB <= 64); the reference keeps vector<bool> with locus A first (ProposalEngine.cpp:596-640). */
typedef struct coalgpu_problem {
    int32_t T, LA, LB;             /* sampling times, segregating sites per locus                */
    const double* times;           /* [T] ascending; times[0] is unused (ProposalEngine.cpp:694) */
    const double* viral;           /* [T] viral loads; N = lambda * viral (main.cpp:229)          */
    double tau;                    /* generation time                                            */
    double driving_mu, driving_lambda, driving_rho;
    int32_t n_mu, n_rho, n_lambda; /* target grid; G = n_mu*n_rho*n_lambda                        */
    const double* mu;              /* [n_mu]                                                     */
    const double* rho;             /* [n_rho] already multiplied by distance+1 (main.cpp:260-271) */
    const double* lambda;          /* [n_lambda]                                                 */
    const int32_t* type_offsets;   /* [T+1] ranges of the arrays below per sampling time         */
    const uint8_t* type_gamma;     /* COALGPU_GAMMA_*                                            */
    const uint64_t* type_words;
    const uint32_t* type_counts;
} coalgpu_problem;

/* What SamplerOutput<double> holds after sumSamples (SamplerOutput.cpp:115-133), log domain, grid
 * index g = (i_mu*n_rho + i_rho)*n_lambda + i_lambda (the loop order of main.cpp:288-306):
 *   likelihood[g]   = log sum_h w_h            (a log SUM, not a mean: ImportanceSampler.cpp:141)
 *   likelihood_sq[g]= log sum_h w_h^2
 *   tmrca[g]        = log( sum_h w_h t_h / sum_h w_h )
 *   lineages[l*G+g] = log( sum_h w_h n_{h,l} / sum_h w_h ),  l < T
 * plus diagnostics the reference leaves to the user: ess[g] = exp(2*likelihood - likelihood_sq). */
typedef struct coalgpu_result {
    double* likelihood;     /* [G]   caller-allocated */
    double* likelihood_sq;  /* [G]   */
    double* tmrca;          /* [G]   */
    double* lineages;       /* [T*G] */
    double* ess;            /* [G]   may be NULL */
    uint64_t n_histories;   /* out */
    uint64_t n_events;      /* out: backward events drawn over all histories */
    uint64_t n_pismc;       /* out: pi_SMC evaluations */
    double seconds_device;  /* out: CUDA-event time of the sampling + reduction kernels */
} coalgpu_result;

/* Integer history encoding: one record per backward event of a traced history.
 * Same layout as the oracle's coalo_event. */
typedef struct coalgpu_event {
    uint64_t chosen_word;
    uint64_t aux_word;
    uint8_t chosen_gamma;
    uint8_t kind;      /* 0 mutation, 1 coalescence within class, 2 C^A, 3 C^B, 4 recombination, 5 A^B */
    uint8_t epoch;
    uint8_t site;
    uint32_t n_before;
} coalgpu_event;

typedef struct coalgpu_sampler coalgpu_sampler;

/* ---- the drop-in: replaces ImportanceSampler::runSampler (ImportanceSampler.cpp:110-145) -------
 * Blocking; host buffers in and out; uses CUDA device `device`. num_threads has no counterpart.
 * Histories h = 0..n_samples-1 use Philox4x32-10 streams keyed by (seed, h). */
int coalgpu_run_sampler(const coalgpu_problem* problem, uint64_t n_samples, uint32_t quadrature_points,
                        uint64_t seed, int device, coalgpu_result* result);

/* The same for several locus pairs at once (the double loop of main.cpp:242-311): pairs are launched on
 * separate CUDA streams so that small batches still fill the device. results[k] belongs to problems[k]. */
int coalgpu_run_sampler_batch(int32_t n_problems, const coalgpu_problem* problems, uint64_t n_samples,
                              uint32_t quadrature_points, uint64_t seed, int device, coalgpu_result* results);

/* ---- one locus pair over several GPUs of one box (sample sharding) --------------------------------
 * The reference fans num_iterations out over std::threads and merges their SamplerOutputs (ImportanceSampler.cpp:
 * 110-145, SamplerOutput.cpp:97-133). Here every device draws the contiguous index range
 *     first_r = r*floor(n/R) + min(r, n mod R),  count_r = floor(n/R) + (r < n mod R)
 * of the SAME histories coalgpu_run_sampler(n_samples, seed) would draw on one device (a history is a pure function
 * of (seed, index)), reduces it on its device, and the per-device (max, sum) pairs are combined over NVLink with NCCL:
 * allreduce(MAX) of the maxima, local rescale, allreduce(SUM). Results equal the single-device call up to the
 * rounding of that merge (1e-13 relative; tests/test_gpu_distributed.py).
 *
 * coalgpu_comm_init / coalgpu_comm_finalize own the NCCL communicator (ncclCommInitAll over `devices`; NULL = devices
 * 0..n_devices-1) and one stream per device; a communicator serves one call at a time. NCCL is bound at run time
 * (libnccl.so.2); with n_devices == 1 it is never loaded. coalgpu_run_sampler_multi is the same call with a
 * process-wide communicator per device list, created on first use and released by coalgpu_multi_shutdown. */
typedef struct coalgpu_comm coalgpu_comm;
int coalgpu_comm_init(int32_t n_devices, const int32_t* devices, coalgpu_comm** out);
void coalgpu_comm_finalize(coalgpu_comm* comm);
int coalgpu_comm_size(const coalgpu_comm* comm);
int coalgpu_run_sampler_comm(coalgpu_comm* comm, const coalgpu_problem* problem, uint64_t n_samples,
                             uint32_t quadrature_points, uint64_t seed, coalgpu_result* result);
int coalgpu_run_sampler_multi(const coalgpu_problem* problem, uint64_t n_samples, uint32_t quadrature_points,
                              uint64_t seed, int32_t n_devices, const int32_t* devices, coalgpu_result* result);
void coalgpu_multi_shutdown(void);
/* NCCL collectives issued by this library on the calling process so far (one per device and allreduce). */
uint64_t coalgpu_nccl_calls(void);

/* Emission-table mode of the samplers created AFTER the call (process-wide, thread-safe):
 *   COALGPU_EMISSION_REFERENCE (default): ConditionalSamplingHMM.cpp:184-226 in the reference's operation order --
 *       tables bit-identical to the reference's; entries below ~1e-13 are rounding noise / the epsilon clamp.
 *   COALGPU_EMISSION_STABLE: the same quantity from the closed form of the double sum, by quadrature (every entry to
 *       1e-14 relative, any locus length), built ON THE DEVICE by its own kernel. Agrees with the
 *       reference mode on every entry that is large enough to matter; histories are then no longer bit-identical to
 *       the oracle's.
 *   COALGPU_EMISSION_STABLE_HOST: the same quadrature evaluated on the host (seconds per data set; the CPU-testable
 *       statement of what the device kernel computes). */
#define COALGPU_EMISSION_REFERENCE 0
#define COALGPU_EMISSION_STABLE 1
#define COALGPU_EMISSION_STABLE_HOST 2
int coalgpu_set_emission_mode(int mode);

/* ---- staged API --------------------------------------------------------------------------------*/
/* Build the HMM tables on the host (ConditionalSamplingHMM.cpp:36-226, same operation order),
 * upload problem + tables to `device`. quadrature_points must be 4, 8 or 16 (ConditionalSamplingHMM.cpp:39-58). */
int coalgpu_create(const coalgpu_problem* problem, uint32_t quadrature_points, int device, coalgpu_sampler** out);
void coalgpu_destroy(coalgpu_sampler* s);

/* One sampler serves ONE stream at a time: coalgpu_sample resets the sampler's device counters and work list per call,
 * and its buffers are released in the order of the stream it was last used on; use one sampler per stream (or order the
 * streams with events). Samplers of different locus pairs are independent.
 *
 * Draw histories [first, first+n) on the sampler's device, asynchronously on `stream`
 * (a cudaStream_t cast to void*, NULL = default stream). Device outputs (any may be NULL):
 *   d_weights[n*G], d_tmrca[n], d_lineages[n*T], d_nevents[n] (uint32)
 * Optional event trace for the first n_trace histories of this call: d_events[n_trace*max_events],
 * d_event_counts[n_trace] (uint32). Replaces ProposalEngine::calculateLikelihood
 * (ProposalEngine.cpp:643-861) + SamplerOutput::addSample (SamplerOutput.cpp:76-93). */
int coalgpu_sample(coalgpu_sampler* s, uint64_t seed, uint64_t first, uint64_t n, void* stream,
                   double* d_weights, double* d_tmrca, double* d_lineages, uint32_t* d_nevents,
                   coalgpu_event* d_events, uint32_t* d_event_counts, uint32_t n_trace, uint32_t max_events);

/* Per-device partial log-sum-exp of n histories: d_partials[(3+T)*G*2] = (max, sum exp(x-max)) for
 * the statistics w, 2w, w+log t, w+log n_l (SamplerOutput.cpp:76-93), layout [stat][g][2]. */
int coalgpu_partials(coalgpu_sampler* s, uint64_t n, void* stream, const double* d_weights,
                     const double* d_tmrca, const double* d_lineages, double* d_partials);

/* Host-side finish: (max, sum) partials of ALL histories -> result grids (SamplerOutput.cpp:115-133). */
int coalgpu_finalize(int32_t G, int32_t T, const double* h_partials, coalgpu_result* result);

/* Work counters accumulated by coalgpu_sample since creation: events, pi_SMC evaluations, and the
 * number of histories that overflowed the type-store capacity (must be 0). */
int coalgpu_counters(coalgpu_sampler* s, uint64_t* n_events, uint64_t* n_pismc, uint64_t* n_overflow);

/* Two-locus pi_SMC evaluations drawn so far and the number of haplotype classes they summed over
 * (the quantities the algorithmic FLOP model of DESIGN.md §5 is written in). */
int coalgpu_work_counters(coalgpu_sampler* s, uint64_t* n_two_locus, uint64_t* n_classes);

/* Capacity statistics of the two-pass scheme: histories of the LAST coalgpu_sample call that outgrew the
 * reduced type store of the first pass and were drawn again with the full one; the two store sizes (equal when the
 * sampler runs single-pass); the most distinct types any history of this sampler has held; lanes per history of the
 * first pass. Any output may be NULL. Blocking. */
int coalgpu_capacity_stats(coalgpu_sampler* s, uint64_t* n_deferred, uint32_t* store_first_pass, uint32_t* store_full,
                           uint32_t* peak_types, uint32_t* lanes_per_history);

/* Re-size the reduced type store and the launch geometry from what the histories drawn so far needed (peak_types).
 * Blocking (synchronises the device). Never changes results, only occupancy. coalgpu_run_sampler does this itself
 * after a pilot chunk; callers of the staged API may call it after a first coalgpu_sample. */
int coalgpu_tune(coalgpu_sampler* s);

/* Measured FP64 FMA throughput of `device` in TFLOP/s (dependent-FMA chains, best of 3): the roofline
 * denominator of this fp64-ALU-bound path (MEASURED_PEAKS.json only carries HBM and bf16 peaks). */
int coalgpu_fp64_peak(int device, double* tflops);

/* ---- inner seams used by the parity tests ------------------------------------------------------*/
/* Host copy of the table row for (n_all, time_idx): y[Q], z[Q*Q] (expanded from the structured
 * form the kernels use), EA[Q*(LA+1)], EB[Q*(LB+1)], w[Q]. Any output may be NULL. */
int coalgpu_tables(coalgpu_sampler* s, int32_t n_all, int32_t time_idx, double* w, double* y, double* z,
                   double* EA, double* EB);

/* Batched pi_SMC from class vectors (the output of TypeContainer::getDifferences,
 * TypeContainer.cpp:748-894): query q has classes [class_offsets[q], class_offsets[q+1]).
 * Host buffers; replaces ConditionalSamplingHMM::piSMC (ConditionalSamplingHMM.cpp:265-448). */
int coalgpu_pismc_classes(coalgpu_sampler* s, int64_t n_queries, const uint8_t* gamma, const int32_t* n_all,
                          const int32_t* time_idx, const int64_t* class_offsets, const int32_t* counts,
                          const int32_t* dA, const int32_t* dB, double* out_pi);

/* The device EVENT code on given configurations: configuration q = the types [cfg_offsets[q], cfg_offsets[q+1]) (gamma, word,
 * count; slot order = list order) with the chosen lineage (chosen_gamma[q], chosen_words[q]) still in it, at sampling epoch
 * epochs[q]. The history kernel removes the chosen lineage and scores every backward move exactly as it does for a drawn
 * event (ProposalEngine.cpp:168-238, 298-338, 375-414); instead of drawing, the unnormalised proposal probabilities come
 * back: out_pv[q*max_cand + k], k < out_ncand[q], in the kernel's candidate order --
 *   C chosen:   k < LA+LB mutation of site k, then C^C, C^A, C^B, recombination;
 *   A/B chosen: k < L mutation of site k, then one A^B_k per stored type of the other one-locus class in slot order
 *               (its haplotype in out_partner_words[q*max_cand + k]), last the coalescence within the class.
 * out_ncand[q] = -1 if the configuration outgrew the type store. Host buffers, blocking. */
int coalgpu_score_events(coalgpu_sampler* s, int64_t n_cfg, const int32_t* cfg_offsets, const uint8_t* gamma,
                         const uint64_t* words, const uint32_t* counts, const uint8_t* chosen_gamma,
                         const uint64_t* chosen_words, const int32_t* epochs, int32_t max_cand, double* out_pv,
                         uint64_t* out_partner_words, int32_t* out_ncand);

/* Kernel launches issued by this library on the calling process so far (bench.py's gpu_launches). */
uint64_t coalgpu_launch_count(void);
/* Bytes the drop-in entry points (coalgpu_create / coalgpu_run_sampler*) have copied host -> device (problem, tables, constants)
 * and device -> host (partials, counters) on the calling process so far: bench.py's e2e.h2d_bytes_per_step / d2h_bytes_per_step. */
void coalgpu_transfer_bytes(uint64_t* h2d, uint64_t* d2h);
const char* coalgpu_last_error(void);
const char* coalgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif
