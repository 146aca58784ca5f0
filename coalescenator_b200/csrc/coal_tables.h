// Host-side construction of the conditional-sampling HMM constants, in the layout the kernels read.
// Follows ConditionalSamplingHMM.cpp:36-226 of the reference in the same floating-point operation
// order, so the emission tables are bit-identical to the reference's; the transition
// matrix is kept in the structured form z(i,j) = {zlow[j] (i>j), w[j]*g[i] (i<j), zdiag[i]} that the
// reference's formulas have (ConditionalSamplingHMM.cpp:110-173), which turns the interval-to-interval
// step of the forward pass from O(Q^2) into O(Q).
#pragma once
#include <cstdint>
#include <vector>

#include "coal_math.h"

namespace coalgpu {

constexpr int kQ = 16;   // interval slots the kernels sweep (main.cpp:103 hard-codes 16; 4 and 8 are padded)

struct TableLayout {
    int LA = 0, LB = 0;
    int off_yw = 0, off_zlow = kQ, off_g = 2 * kQ, off_zdiag = 3 * kQ;
    int off_EA = 4 * kQ, off_EB = 0, off_wEA = 0, off_wEB = 0;   // EA: [(LA+1)][kEStride], EB: [(LB+1)][kEStride] (distance-major, padded)
    int row_doubles = 0;   // multiple of 2
    // What the history kernel reads (coal_math.h "bilinear form"): per (table set, lineage count) one DEVICE row
    //   MY[(LA+2)][(LB+2)][2] | wEA[LA+1] | wEB[LB+1]
    // M[dA][dB] = sum_ij EA[dA][i] w_i z(i,j) EB[dB][j] and Y[dA][dB] = sum_j y_j w_j EA[dA][j] EB[dB][j], interleaved; index
    // LA+1 / LB+1 stands for "no lineage ancestral at this locus" (emission 1 at every interval).
    int dev_sb = 0;        // LB + 2: entries per dA row of MY
    int dev_off_wEA = 0, dev_off_wEB = 0, dev_row_doubles = 0;   // dev_row_doubles: multiple of 2 (16-byte rows for bulk copies)
    void init(int la, int lb) {
        LA = la; LB = lb;
        off_EB = off_EA + kEStride * (LA + 1);
        off_wEA = off_EB + kEStride * (LB + 1);
        off_wEB = off_wEA + (LA + 1);
        row_doubles = off_wEB + (LB + 1);
        row_doubles = (row_doubles + 1) & ~1;
        dev_sb = LB + 2;
        dev_off_wEA = 2 * (LA + 2) * (LB + 2);
        dev_off_wEB = dev_off_wEA + (LA + 1);
        dev_row_doubles = (dev_off_wEB + (LB + 1) + 1) & ~1;
    }
};

struct HostTables {
    TableLayout lay;
    int Q = kQ;                    // quadrature intervals in use (4, 8 or 16); tables are padded to kQ
    int n_max = 0;                 // rows for n_all = 1..n_max
    int n_sets = 0;                // distinct (theta, r) epochs
    std::vector<int> set_of_epoch; // [T]
    std::vector<double> w;         // [kQ] interval weights
    std::vector<double> quad;      // [kQ+1]
    std::vector<double> rows;      // [n_sets][n_max][row_doubles]: the reference's constants (coalgpu_tables, device table builder input)
    std::vector<double> dev_rows;  // [n_sets][n_max][dev_row_doubles]: what the kernels read; empty when the emission tables are built on the device
    std::vector<double> y;         // [n_sets][n_max][kQ] plain y (for coalgpu_tables)
    const double* row(int set, int n_all) const { return rows.data() + ((size_t)set * n_max + (n_all - 1)) * lay.row_doubles; }
    const double* dev_row(int set, int n_all) const { return dev_rows.data() + ((size_t)set * n_max + (n_all - 1)) * lay.dev_row_doubles; }
};

// thetas[c] = 4*lambda_d*V[c]*mu_d, r_rates[c] = 4*lambda_d*V[c]*rho_d (ImportanceSampler.cpp:69-76)
// Emission blocks are shared between calls through a process-wide cache keyed by (Q, L, theta) (coal_tables.cpp).
// threads: host threads used for blocks that are not cached yet (0 = all hardware threads, at most 16).
// emission_mode: kEmissionReference = the reference's double sum in its operation order (bit-identical tables, loses
// accuracy beyond ~20-30 sites per locus); kEmissionStable = the same quantity by quadrature of its closed form.
enum { kEmissionReference = 0, kEmissionStable = 1, kEmissionNone = 2 };   // None: emission parts left zero (built on the device)
bool build_tables(int Q, int LA, int LB, int n_max, const std::vector<double>& thetas,
                  const std::vector<double>& r_rates, HostTables& out, int threads = 0, int emission_mode = kEmissionReference);
void clear_table_cache();

}  // namespace coalgpu
