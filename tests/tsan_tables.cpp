// ThreadSanitizer harness for the process-wide emission cache of coal_tables.cpp: worker threads (the pairs of
// coalgpu_run_sampler_batch) call build_tables with the same (Q, L, theta) key and DIFFERENT n_max, so later
// callers grow a cached block while earlier ones still copy rows out of it. Exit code 0 = tables correct; TSan
// itself fails the process (exit 66) on a data race.
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../coalescenator_b200/csrc/coal_tables.h"

using namespace coalgpu;

int main() {
    const int kThreads = 16;
    std::vector<HostTables> tabs(kThreads);
    std::vector<int> ok(kThreads, 0);
    std::vector<std::thread> th;
    const std::vector<double> thetas{0.01, 0.02}, rr{0.02, 0.04};
    for (int k = 0; k < kThreads; k++)
        th.emplace_back([&, k]() { ok[k] = build_tables(16, 6, 5, 100 + 40 * k, thetas, rr, tabs[k], 1) ? 1 : 0; });
    for (auto& t : th) t.join();
    // every thread's rows for the lineage counts it shares with the largest request must be the same doubles
    const HostTables& big = tabs[kThreads - 1];
    for (int k = 0; k < kThreads; k++) {
        if (!ok[k]) { std::printf("build_tables failed in thread %d\n", k); return 1; }
        for (int s = 0; s < tabs[k].n_sets; s++)
            for (int n = 1; n <= tabs[k].n_max; n++)
                if (std::memcmp(tabs[k].row(s, n), big.row(s, n), sizeof(double) * tabs[k].lay.row_doubles) != 0) {
                    std::printf("row (set %d, n %d) of thread %d differs\n", s, n, k);
                    return 2;
                }
    }
    std::printf("ok\n");
    return 0;
}
