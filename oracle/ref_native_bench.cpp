// Native timing harness over the UNMODIFIED reference engine (test / measurement infrastructure only, never part of the
// product path).  It includes the reference's own headers from where they lie (-I/root/reference/src/cpp; nothing is copied)
// and drives AlphaZero::BatchedMCTS<Connect4> through the per-move loop of src/MCTS_cpp.py:217-357 with no Python in between:
//   prune_roots(-1) -> search_batch / backprop_batch (warm-up simulation) -> ceil((n-1)/K) x search_batch_vl / backprop_batch_vl
// with the same constant evaluator bench.py uses (uniform prior, fixed WDL / moves-left, terminal leaves as found).
// "a native C++ harness over BatchedMCTS.h (no Python)".
//   usage: ref_native_bench <boards_file> <n> <n_playout> <K> <steps> <warmup>     (boards_file: int8[n*42] then int32[n])
// prints one JSON object: seconds inside the engine's entry points and of the whole loop, OpenMP threads used.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "Connect4.h"
#include "BatchedMCTS.h"

using Clock = std::chrono::steady_clock;
static double secs(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); }

int main(int argc, char **argv) {
    if (argc < 7) { std::fprintf(stderr, "usage: %s boards_file n n_playout K steps warmup\n", argv[0]); return 2; }
    const int n = std::atoi(argv[2]), n_playout = std::atoi(argv[3]), K = std::atoi(argv[4]), steps = std::atoi(argv[5]),
              warmup = std::atoi(argv[6]);
    constexpr int S = 42, A = 7;
    std::vector<int8_t> boards((size_t)n * S);
    std::vector<int> turns(n);
    FILE *f = std::fopen(argv[1], "rb");
    if (!f || std::fread(boards.data(), 1, boards.size(), f) != boards.size() ||
        std::fread(turns.data(), sizeof(int), (size_t)n, f) != (size_t)n) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    std::fclose(f);

    AlphaZero::BatchedMCTS<AlphaZero::Connect4> eng(n);
    AlphaZero::SearchConfig &c = eng.config();          // server defaults (server.py:44-72,133-167), as in bench.py
    c.c_init = 1.4f; c.c_base = 1000.0f; c.fpu_reduction = 0.2f; c.dirichlet_alpha = 0.3f; c.noise_epsilon = 0.25f;
    c.mlh_slope = 0.1f; c.mlh_cap = 0.2f; c.use_symmetry = true; c.value_decay = 1.0f; c.vl_count = 1;
    eng.set_seed(0);

    const size_t rows = (size_t)n * K;
    std::vector<int8_t> lb(rows * S);
    std::vector<float> td(rows), tp1(rows), tp2(rows), pol(rows * A), d(rows), p1w(rows), p2w(rows), ml(rows);
    std::vector<uint8_t> it(rows), vm(rows * A);
    std::vector<int> lt(rows), sym(rows), reset(n, -1);

    auto evaluate = [&](size_t m) {                     // bench.py host_step's evaluate(), on all threads
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)m; ++i) {
            const bool t = it[i] != 0, first = lt[i] == 1;
            for (int a = 0; a < A; ++a) pol[i * A + a] = t ? 0.0f : 1.0f;
            d[i] = t ? td[i] : 0.25f;
            p1w[i] = t ? tp1[i] : (first ? 0.5f : 0.25f);
            p2w[i] = t ? tp2[i] : (first ? 0.25f : 0.5f);
            ml[i] = t ? 0.0f : 10.0f;
        }
    };
    double t_engine = 0.0, t_total = 0.0;
    for (int s = 0; s < warmup + steps; ++s) {
        double eng_s = 0.0;
        const auto t0 = Clock::now();
        auto a = Clock::now();
        eng.prune_roots(std::span<const int>(reset.data(), reset.size()));
        eng.search_batch(boards.data(), turns.data(), lb.data(), td.data(), tp1.data(), tp2.data(), it.data(), lt.data(), vm.data());
        eng_s += secs(a, Clock::now());
        evaluate((size_t)n);
        a = Clock::now();
        eng.backprop_batch(pol.data(), d.data(), p1w.data(), p2w.data(), ml.data(), it.data());
        eng_s += secs(a, Clock::now());
        for (int remaining = n_playout - 1; remaining > 0;) {
            const int cur = remaining < K ? remaining : K;
            remaining -= cur;
            a = Clock::now();
            eng.search_batch_vl(cur, boards.data(), turns.data(), lb.data(), td.data(), tp1.data(), tp2.data(), it.data(), lt.data(),
                                sym.data(), vm.data());
            eng_s += secs(a, Clock::now());
            evaluate((size_t)n * cur);
            a = Clock::now();
            eng.backprop_batch_vl(cur, pol.data(), d.data(), p1w.data(), p2w.data(), ml.data(), it.data(), sym.data());
            eng_s += secs(a, Clock::now());
        }
        if (s >= warmup) { t_engine += eng_s; t_total += secs(t0, Clock::now()); }
    }
    std::vector<int> counts((size_t)n * A);
    long long visits = 0;
    {   // visit counts of the last step: a checksum so the work cannot be optimised away, and a sanity figure for the caller
        auto st = eng.get_all_counts();
        for (int v : st) visits += v;
    }
    std::printf("{\"n\": %d, \"n_playout\": %d, \"K\": %d, \"steps\": %d, \"threads\": %d, \"engine_s\": %.6f, \"total_s\": %.6f, "
                "\"root_visits\": %lld}\n", n, n_playout, K, steps, omp_get_max_threads(), t_engine, t_total, visits);
    return 0;
}
