// Counter-based RNG.  Everything random in the engine (leaf symmetry ids, rollouts, Dirichlet noise) is a pure
// function of (seed, epoch, stream, env, counter), so results do not depend on how games are sharded over
// GPUs/threads.  The reference uses a thread_local std::mt19937 per OpenMP thread (src/cpp/MCTS.h:13-17,
// BatchedMCTS.h:68-84); that stream cannot be reproduced, so the integer draws here are
// matched bit-for-bit by the C restatement (oracle/az_oracle.c: orc_rand) instead.
#pragma once
#include <stdint.h>
#include "az_games.cuh"

namespace az {

enum { STREAM_SYM = 0, STREAM_ROLLOUT = 1, STREAM_NOISE = 2 };

AZ_HD uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
AZ_HD uint64_t az_rand(uint64_t seed, uint64_t epoch, uint64_t stream, uint64_t env, uint64_t ctr) {
    uint64_t h = splitmix64(seed ^ (stream * 0xD6E8FEB86659FD93ULL));
    h = splitmix64(h ^ epoch);
    h = splitmix64(h ^ (env << 24) ^ ctr);
    return h;
}
// config-2 rollout policy hash; twin of orc_rollout_hash
AZ_HD uint64_t rollout_hash(uint64_t seed, uint64_t gidx, uint64_t ply) {
    return splitmix64(splitmix64(seed ^ 0xA5A5A5A55A5A5A5AULL) ^ (gidx << 8) ^ ply);
}

}  // namespace az
