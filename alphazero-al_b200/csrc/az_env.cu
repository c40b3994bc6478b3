// Bitboard game environments: host-side single-game API (mirror of the reference's Env objects) and lockstep
// device kernels over az_root[n] records.  Game logic lives in az_games.cuh and is shared with the MCTS kernels.
//   reference: src/cpp/Connect4.h, src/cpp/Othello.h, src/cpp/env_common.h, env_connect4.h, env_othello.h
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/azb200_env.h"
#include "az_games.cuh"
#include "az_rng.cuh"

namespace az {

AZ_HD State to_state(const az_root &r) { State s; s.bb[0] = r.bb0; s.bb[1] = r.bb1; s.turn = r.turn; s.passes = r.passes; s.last = r.last; return s; }
AZ_HD void from_state(az_root &r, const State &s) { r.bb0 = s.bb[0]; r.bb1 = s.bb[1]; r.turn = s.turn; r.passes = s.passes; r.last = s.last; r.reserved = 0; }

// legal actions in ascending order, including Othello's pass (get_valid_moves: Connect4.h:209-218, Othello.h:282-296)
template <class G> AZ_HD int legal_count(const State &s, uint64_t &mask, bool &pass_only) {
    mask = G::legal(s);
    pass_only = false;
    if (G::GAME == GAME_OTH) {
        if (mask == 0ULL && !Oth::over(s)) { pass_only = true; return 1; }
    }
    return popc64(mask);
}
template <class G> AZ_HD int nth_action(uint64_t mask, bool pass_only, int idx) {
    if (pass_only) return 64;
    for (int i = 0; i < idx; ++i) mask &= mask - 1;
    return ctz64(mask);
}
AZ_HD uint64_t final_digest(const State &s, int winner, int plies) {
    uint64_t d = splitmix64(s.bb[0]) ^ splitmix64(s.bb[1] + 0x1234567ULL);
    return splitmix64(d ^ (uint64_t)(uint32_t)(winner + 1) ^ ((uint64_t)plies << 8) ^ ((uint64_t)(uint32_t)(s.turn + 1) << 20));
}

template <class G> __global__ void k_envs_reset(int n, az_root *st) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    State s; G::reset(s);
    az_root r; from_state(r, s);
    *reinterpret_cast<uint4 *>(st + i) = *reinterpret_cast<uint4 *>(&r);
    *(reinterpret_cast<uint4 *>(st + i) + 1) = *(reinterpret_cast<uint4 *>(&r) + 1);
}
template <class G> __global__ void k_envs_step(int n, az_root *st, const int32_t *__restrict__ actions, int32_t *winners, uint8_t *dones) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    az_root r;
    *reinterpret_cast<uint4 *>(&r) = *reinterpret_cast<const uint4 *>(st + i);
    *(reinterpret_cast<uint4 *>(&r) + 1) = *(reinterpret_cast<const uint4 *>(st + i) + 1);
    State s = to_state(r);
    const int a = actions[i];
    if (a >= 0 && !G::done(s)) {
        G::step(s, a);
        from_state(r, s);
        *reinterpret_cast<uint4 *>(st + i) = *reinterpret_cast<uint4 *>(&r);
        *(reinterpret_cast<uint4 *>(st + i) + 1) = *(reinterpret_cast<uint4 *>(&r) + 1);
    }
    if (winners) winners[i] = G::winner(s);
    if (dones) dones[i] = G::done(s) ? 1 : 0;
}
// one thread per (game, cell): coalesced byte-board / mask stores
template <class G> __global__ void k_envs_observe(int n, const az_root *__restrict__ st, int8_t *boards, uint8_t *masks, int32_t *turns,
                                                  int32_t *winners, uint8_t *dones) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t)n * G::S) return;
    const size_t g = t / G::S; const int j = (int)(t - g * G::S);
    az_root r;
    *reinterpret_cast<uint4 *>(&r) = *reinterpret_cast<const uint4 *>(st + g);
    *(reinterpret_cast<uint4 *>(&r) + 1) = *(reinterpret_cast<const uint4 *>(st + g) + 1);
    const State s = to_state(r);
    if (boards) boards[t] = (int8_t)G::cell(s, j);
    if (masks) {
        uint64_t m; bool po; legal_count<G>(s, m, po);
        if (G::GAME == GAME_C4) { if (j < 7) masks[g * 7 + j] = (uint8_t)((m >> j) & 1ULL); }
        else { masks[g * 65 + j] = (uint8_t)((m >> j) & 1ULL); if (j == 0) masks[g * 65 + 64] = po ? 1 : 0; }
    }
    if (j == 0) {
        if (turns) turns[g] = s.turn;
        if (winners) winners[g] = G::winner(s);
        if (dones) dones[g] = G::done(s) ? 1 : 0;
    }
}
template <class G> __global__ void k_envs_rollout(int n, uint64_t seed, uint64_t first, uint64_t *digest, int32_t *plies, int nrec, int maxp,
                                                  int8_t *rb, uint8_t *rm, int32_t *rt, int32_t *ra, int32_t *rw, uint8_t *rd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t g = first + (uint64_t)i;
    State s; G::reset(s);
    int ply = 0;
    const bool rec = i < nrec;
    while (ply < maxp && !G::done(s)) {
        uint64_t m; bool po;
        const int cnt = legal_count<G>(s, m, po);
        const int a = nth_action<G>(m, po, (int)(rollout_hash(seed, g, (uint64_t)ply) % (uint64_t)cnt));
        if (rec) {
            const size_t o = (size_t)i * maxp + ply;
            for (int j = 0; j < G::S; ++j) rb[o * G::S + j] = (int8_t)G::cell(s, j);
            for (int x = 0; x < G::A; ++x) rm[o * G::A + x] = (G::GAME == GAME_OTH && x == 64) ? (po ? 1 : 0) : (uint8_t)((m >> (x & 63)) & 1ULL);
            rt[o] = s.turn; ra[o] = a;
        }
        G::step(s, a);
        if (rec) { const size_t o = (size_t)i * maxp + ply; rw[o] = G::winner(s); rd[o] = G::done(s) ? 1 : 0; }
        ++ply;
    }
    if (digest) digest[i] = final_digest(s, G::winner(s), ply);
    if (plies) plies[i] = ply;
}

template <class G> void h_import(az_root *r, const int8_t *b) {
    uint64_t p0 = 0, p1 = 0;
    if (G::GAME == GAME_C4) {
        for (int c = 0; c < 7; ++c)
            for (int row = 5; row >= 0; --row) {
                int v = b[row * 7 + c];
                if (v == 0) break;
                uint64_t bit = 1ULL << (c * 7 + (5 - row));
                if (v == 1) p0 |= bit; else p1 |= bit;
            }
    } else {
        for (int j = 0; j < 64; ++j) { if (b[j] == 1) p0 |= 1ULL << j; else if (b[j] == -1) p1 |= 1ULL << j; }
    }
    State s; s.bb[0] = p0; s.bb[1] = p1;
    G::finish_import(s, r->turn);
    from_state(*r, s);
}

}  // namespace az

using namespace az;
#define AZ_GAME_SWITCH(game, EXPR_C4, EXPR_OTH) ((game) == GAME_C4 ? (EXPR_C4) : (EXPR_OTH))

extern "C" {

void az_env_reset(int game, az_root *r) { State s; if (game == GAME_C4) C4::reset(s); else Oth::reset(s); from_state(*r, s); }
void az_env_import(int game, az_root *r, const int8_t *b) { if (game == GAME_C4) h_import<C4>(r, b); else h_import<Oth>(r, b); }
void az_env_export(int game, const az_root *r, int8_t *b) {
    const State s = to_state(*r);
    if (game == GAME_C4) for (int j = 0; j < 42; ++j) b[j] = (int8_t)C4::cell(s, j);
    else for (int j = 0; j < 64; ++j) b[j] = (int8_t)Oth::cell(s, j);
}
int az_env_n_pieces(int game, const az_root *r) { const State s = to_state(*r); return AZ_GAME_SWITCH(game, C4::n_pieces(s), Oth::n_pieces(s)); }
void az_env_step(int game, az_root *r, int a) { State s = to_state(*r); if (game == GAME_C4) C4::step(s, a); else Oth::step(s, a); from_state(*r, s); }
int az_env_winner(int game, const az_root *r) { const State s = to_state(*r); return AZ_GAME_SWITCH(game, C4::winner(s), Oth::winner(s)); }
int az_env_full(int game, const az_root *r) { const State s = to_state(*r); return AZ_GAME_SWITCH(game, (int)C4::full(s), (int)Oth::full(s)); }
int az_env_done(int game, const az_root *r) { const State s = to_state(*r); return AZ_GAME_SWITCH(game, (int)C4::done(s), (int)Oth::done(s)); }
int az_env_valid_moves(int game, const az_root *r, int32_t *moves) {
    const State s = to_state(*r);
    uint64_t m; bool po; int n;
    if (game == GAME_C4) n = legal_count<C4>(s, m, po); else n = legal_count<Oth>(s, m, po);
    if (po) { moves[0] = 64; return 1; }
    int k = 0;
    for (; m; m &= m - 1) moves[k++] = ctz64(m);
    return n;
}
void az_env_apply_symmetry(int game, az_root *r, int sym) {
    State s = to_state(*r);
    if (game == GAME_C4) C4::symmetry(s, sym); else Oth::symmetry(s, sym);
    from_state(*r, s);
}
int az_env_inverse_symmetry_action(int game, int sym, int a) {
    if (game == GAME_C4) return sym == 0 ? a : 6 - a;                     /* env_connect4.h:46-49 */
    if (a == 64 || sym == 0) return a;                                    /* env_othello.h:46-56 */
    return Oth::xform_idx(Oth::inverse_sym(sym), a);
}

#define AZ_ENV_LAUNCH(game, KERNEL, GRID, BS, STREAM, ...)                                             \
    do {                                                                                               \
        if ((game) == GAME_C4) KERNEL<C4><<<GRID, BS, 0, (cudaStream_t)(STREAM)>>>(__VA_ARGS__);       \
        else if ((game) == GAME_OTH) KERNEL<Oth><<<GRID, BS, 0, (cudaStream_t)(STREAM)>>>(__VA_ARGS__); \
        else return AZ_ERR_INVALID;                                                                    \
        return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;                                \
    } while (0)

int az_envs_reset_dev(int game, int n, az_root *st, void *stream) {
    if (n <= 0) return AZ_OK;
    AZ_ENV_LAUNCH(game, k_envs_reset, (n + 127) / 128, 128, stream, n, st);
}
int az_envs_step_dev(int game, int n, az_root *st, const int32_t *actions, int32_t *winners, uint8_t *dones, void *stream) {
    if (n <= 0) return AZ_OK;
    AZ_ENV_LAUNCH(game, k_envs_step, (n + 127) / 128, 128, stream, n, st, actions, winners, dones);
}
int az_envs_observe_dev(int game, int n, const az_root *st, int8_t *boards, uint8_t *masks, int32_t *turns, int32_t *winners, uint8_t *dones,
                        void *stream) {
    if (n <= 0) return AZ_OK;
    const size_t cells = (size_t)n * (game == GAME_C4 ? 42 : 64);
    AZ_ENV_LAUNCH(game, k_envs_observe, (unsigned)((cells + 255) / 256), 256, stream, n, st, boards, masks, turns, winners, dones);
}
int az_envs_rollout_dev(int game, int n, uint64_t seed, uint64_t first, uint64_t *digest, int32_t *plies, int nrec, int maxp, int8_t *rb,
                        uint8_t *rm, int32_t *rt, int32_t *ra, int32_t *rw, uint8_t *rd, void *stream) {
    if (n <= 0) return AZ_OK;
    if (maxp <= 0) return AZ_ERR_INVALID;
    AZ_ENV_LAUNCH(game, k_envs_rollout, (n + 127) / 128, 128, stream, n, seed, first, digest, plies, nrec, maxp, rb, rm, rt, ra, rw, rd);
}

}  // extern "C"
