/* TEST INFRASTRUCTURE — CPU oracle for the importance-sampling likelihood hot path.
 *
 * This is a from-scratch CPU restatement of the reference algorithm (file:line citations are in
 * coal_oracle.cpp). It is the checker for tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg ONLY. Nothing in the product path (coalescenator_b200/csrc, libcoalgpu.so)
 * may include, link or call it.
 *
 * Parity status: PINNED. `mode = COALO_MODE_STRICT` reproduces the compiled reference
 * (oracle/_ref/coalref, built from the unmodified sources) bit-for-bit on whole runs: same
 * std::mt19937 stream, same libstdc++ unordered_map iteration order, same floating-point operation
 * order (tests/test_oracle_vs_reference.py, fixtures under tests/golden/).
 * `mode = COALO_MODE_CANONICAL` is the same algorithm with the two implementation-defined choices
 * replaced by the ones the CUDA path makes: counter-based Philox4x32-10 draws and slot-order
 * iteration of the type store with an order-independent class representative.
 */
#ifndef COAL_ORACLE_H
#define COAL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COALO_MODE_STRICT 0
#define COALO_MODE_CANONICAL 1

#define COALO_GAMMA_A 0
#define COALO_GAMMA_B 1
#define COALO_GAMMA_C 2

typedef struct coalo_problem {
    int32_t T, LA, LB;
    const double* times;        /* [T] ascending; times[0] unused (ProposalEngine.cpp:694) */
    const double* viral;        /* [T] */
    double tau, driving_mu, driving_lambda, driving_rho;
    int32_t n_mu, n_rho, n_lambda;
    const double* mu;           /* [n_mu]     */
    const double* rho;          /* [n_rho]    already scaled by distance+1 (main.cpp:260-271) */
    const double* lambda;       /* [n_lambda] */
    const int32_t* type_offsets;/* [T+1] ranges into the three arrays below, insertion order */
    const uint8_t* type_gamma;  /* 0=A (locus A only), 1=B, 2=C (both) */
    const uint64_t* type_words; /* locus A site s = bit s, locus B site s = bit LA+s */
    const uint32_t* type_counts;
} coalo_problem;

/* One record per event of a traced history (the "integer history encoding"). */
typedef struct coalo_event {
    uint64_t chosen_word;   /* lineage picked by the lineage chooser */
    uint64_t aux_word;      /* mutated haplotype / coalescence partner / 0 */
    uint8_t chosen_gamma;
    uint8_t kind;           /* COALO_EV_* */
    uint8_t epoch;          /* index c of ProposalEngine.cpp:694 minus 1 (HMM time_idx) */
    uint8_t site;           /* flipped site for mutations, else 0xff */
    uint32_t n_before;      /* lineage count before the event */
} coalo_event;

#define COALO_EV_MUT 0
#define COALO_EV_COAL_SAME 1   /* C^C, or A^A / A^C, or B^B / B^C */
#define COALO_EV_COAL_CA 2
#define COALO_EV_COAL_CB 3
#define COALO_EV_REC 4
#define COALO_EV_COAL_AB 5     /* A^B_k or B^A_k, aux_word = partner */

/* Gauss-Laguerre interval tables (ConditionalSamplingHMM.cpp:36-226). Outputs may be NULL.
 * y[Q], z[Q*Q] row-major (i,j), EA[Q*(LA+1)], EB[Q*(LB+1)] row-major (i,d), w[Q]. */
int coalo_tables(const coalo_problem* p, int Q, int n_all, int time_idx,
                 double* w, double* y, double* z, double* EA, double* EB);

/* pi_SMC from getDifferences' output (ConditionalSamplingHMM.cpp:329-447). */
double coalo_pismc_classes(const coalo_problem* p, int Q, int gamma, int n_all, int time_idx,
                           int n_classes, const int32_t* counts, const int32_t* dA, const int32_t* dB);

/* Draw histories [first, first+n) of the run (seed, mode). Per history h (0-based inside the call):
 *   weights[h*G + g]   log importance weight at grid point g = (i_mu*n_rho + i_rho)*n_lambda + i_lambda
 *   tmrca[h]           time at which <= 5 lineages remain (ProposalEngine.cpp:817-818)
 *   lineages[h*T + l]  lineages left at each sampling time (ProposalEngine.cpp:751,817)
 *   n_events[h], n_pismc[h]  work counters (may be NULL)
 * In STRICT mode `first` must be 0 (the mt19937 stream is sequential).
 * trace_path (may be NULL): text trace of the first trace_max histories in oracle/README.md format.
 * events (may be NULL): up to max_events records per history for the first n_event_hist histories,
 * event_counts[h] = number of events of history h. */
int coalo_run(const coalo_problem* p, int Q, int mode, uint64_t seed, uint64_t first, uint64_t n,
              double* weights, double* tmrca, double* lineages, int64_t* n_events, int64_t* n_pismc,
              const char* trace_path, int64_t trace_max,
              coalo_event* events, int64_t n_event_hist, int64_t max_events, int64_t* event_counts);

/* Candidate vectors of the events (tests of the device event code): the next coalo_run calls write one "V" line per event
 * of their first n_histories histories to `path` (NULL / "" switches the dump off); format in coal_oracle.cpp (Engine). */
void coalo_set_score_dump(const char* path, int64_t n_histories);

/* The unnormalised proposal probabilities of ONE event (ProposalEngine.cpp:171-238, 298-338, 375-414) in CANONICAL mode:
 * configuration = the given type list in slot order with the chosen lineage still in it. Returns the number of
 * candidates (labels: 24 bytes each, as in the V lines) or a negative error. */
int coalo_score_event(const coalo_problem* p, int Q, int epoch, int n_types, const uint8_t* gamma, const uint64_t* words,
                      const uint32_t* counts, int chosen_gamma, uint64_t chosen_word, int max_cand, double* out_vals,
                      char* out_labels);

/* sumSamples (SamplerOutput.cpp:115-133, LikelihoodSamples.cpp:56-70): descending-order
 * logAddition per grid point. out_* are [G], out_lineages is [T*G]; normalisation as the reference
 * (log-sum, not log-mean; tmrca and lineages divided by the likelihood). */
int coalo_reduce(int64_t n, int32_t G, int32_t T, const double* weights, const double* tmrca,
                 const double* lineages, double* out_like, double* out_likesq, double* out_tmrca,
                 double* out_lineages);

/* Philox4x32-10 uniform used by CANONICAL mode: draw number `k` of history `hist`. */
double coalo_uniform(uint64_t seed, uint64_t hist, uint64_t k);

#ifdef __cplusplus
}
#endif
#endif
