// On-device self-play driver: per-ply policy target / sampling / recording / env step (k_sp_ply) and per-game
// training-tuple construction into packed records (k_sp_flush).  Reference behaviour: src/game.py:65-164,
// src/player.py:333-375.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/azb200_selfplay.h"
#include "az_games.cuh"
#include "az_rng.cuh"

namespace az {

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }
template <class G> __host__ __device__ inline az_selfplay_layout make_layout() {
    az_selfplay_layout L;
    const int T1 = G::MAX_PLIES + 1;
    int o = 0;
    L.T1 = T1;
    L.off_header = o; o = align16(o + 16);
    L.off_state = o; o = align16(o + T1 * 3 * G::S);
    L.off_prob = o; o = align16(o + T1 * G::A * 4);
    L.off_root_wdl = o; o = align16(o + T1 * 12);
    L.off_future = o; o = align16(o + T1 * 12);
    L.off_winner = o; o = align16(o + T1);
    L.off_steps = o; o = align16(o + T1 * 2);
    L.off_aux = o; o = align16(o + T1 * 2);
    L.off_mask = o; o = align16(o + T1 * G::A);
    L.record_bytes = o;
    return L;
}

__device__ __forceinline__ State sp_state(const az_root &r) { State s; s.bb[0] = r.bb0; s.bb[1] = r.bb1; s.turn = r.turn; s.passes = r.passes; s.last = r.last; return s; }

// One thread per game slot.
template <class G> __global__ void k_sp_ply(az_selfplay sp, const int32_t *__restrict__ counts, const float *__restrict__ stats) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= sp.n) return;
    constexpr int A = G::A, S = G::S, T = G::MAX_PLIES;
    az_root r = sp.states[g];
    State s = sp_state(r);
    const int step = sp.steps[g];
    const int32_t *c = counts + (size_t)g * A;
    long long total = 0; int best = 0, bestv = -1;
    for (int a = 0; a < A; ++a) { const int v = c[a]; total += v; if (v > bestv) { bestv = v; best = a; } }   // np.argmax: first maximum
    int action = 0;
    if (total > 0) {
        const float temp = (sp.temp_decay_moves <= 0 || step < sp.temp_decay_moves) ? sp.temp_init : sp.temp_endgame;   // src/game.py:54-63
        if (temp <= 1e-6f) action = best;
        else {   // softmax(log(visits)/temp) over visited actions (src/player.py:364-369); RNG stream is ours
            const double invt = 1.0 / (double)temp, lmax = log((double)bestv) * invt;
            double z = 0.0;
            for (int a = 0; a < A; ++a) if (c[a] > 0) z += exp(log((double)c[a]) * invt - lmax);
            const uint64_t h = az_rand(sp.seed, (uint64_t)step, 3, sp.uids[g], 0);
            double u = ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0) * z, acc = 0.0;
            action = best;
            for (int a = 0; a < A; ++a) if (c[a] > 0) { acc += exp(log((double)c[a]) * invt - lmax); if (u < acc) { action = a; break; } }
        }
    }
    // ---- record the position (src/game.py:101-110) ----
    if (step < T) {
        const size_t o = (size_t)g * T + step;
        int8_t *ps = sp.st_state + o * 3 * S;
        const uint64_t own_bb = s.turn == 1 ? s.bb[0] : s.bb[1], opp_bb = s.turn == 1 ? s.bb[1] : s.bb[0];
        for (int j = 0; j < S; ++j) {
            const int bit = G::cell_bit(j);
            ps[j] = (int8_t)((own_bb >> bit) & 1ULL);
            ps[S + j] = (int8_t)((opp_bb >> bit) & 1ULL);
            ps[2 * S + j] = (int8_t)s.turn;
        }
        float *pp = sp.st_prob + o * A;
        for (int a = 0; a < A; ++a) pp[a] = total > 0 ? (float)((double)c[a] / (double)total) : 0.0f;   // visits / visits.sum() -> float32
        const float *st = stats + (size_t)g * (6 + 8 * A);
        sp.st_wdl[o * 3 + 0] = st[3]; sp.st_wdl[o * 3 + 1] = st[4]; sp.st_wdl[o * 3 + 2] = st[5];   // root_D, root_P1W, root_P2W
        uint64_t legal = G::legal(s);
        bool pass_only = false;
        if (G::GAME == GAME_OTH) pass_only = legal == 0ULL && !Oth::over(s);
        uint8_t *pm = sp.st_mask + o * A;
        for (int a = 0; a < A; ++a) pm[a] = (G::GAME == GAME_OTH && a == 64) ? (pass_only ? 1 : 0) : (uint8_t)((legal >> (a & 63)) & 1ULL);
        sp.st_player[o] = (int8_t)s.turn;
    }
    // ---- env.step + done (src/game.py:112-113) ----
    G::step(s, action);
    r.bb0 = s.bb[0]; r.bb1 = s.bb[1]; r.turn = s.turn; r.passes = s.passes; r.last = s.last;
    sp.states[g] = r;
    sp.steps[g] = step + 1;
    const bool done = G::done(s) || step + 1 >= T;
    sp.finished[g] = done ? 1 : 0;
    sp.actions[g] = done ? -1 : action;      // finished game: the tree is reset (src/game.py:158, reset_env)
}

// One CTA per game slot; only finished slots do work.
template <class G> __global__ void k_sp_flush(az_selfplay sp) {
    const int g = blockIdx.x;
    if (!sp.finished[g]) return;
    constexpr int A = G::A, S = G::S, T = G::MAX_PLIES;
    const az_selfplay_layout L = make_layout<G>();
    __shared__ int slot_s;
    if (threadIdx.x == 0) slot_s = atomicAdd(sp.out_count, 1);
    __syncthreads();
    const int slot = slot_s;
    const int Tn = sp.steps[g];                        // positions played
    const az_root r = sp.states[g];
    const State s = sp_state(r);
    const int winner = G::winner(s);
    const int diff = popc64(s.bb[0]) - popc64(s.bb[1]);
    if (slot < sp.out_capacity) {
        uint8_t *rec = sp.out + (size_t)slot * sp.record_bytes;
        for (int i = threadIdx.x; i < sp.record_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(rec)[i] = 0u;
        __syncthreads();
        if (threadIdx.x == 0) {
            reinterpret_cast<int32_t *>(rec + L.off_header)[0] = Tn + 1;
            reinterpret_cast<int32_t *>(rec + L.off_header)[1] = winner;
            reinterpret_cast<uint64_t *>(rec + L.off_header)[1] = sp.uids[g];
        }
        const size_t base = (size_t)g * T;
        for (int i = threadIdx.x; i < Tn * 3 * S; i += blockDim.x) rec[L.off_state + i] = (uint8_t)sp.st_state[base * 3 * S + i];
        for (int i = threadIdx.x; i < Tn * A; i += blockDim.x) {
            reinterpret_cast<float *>(rec + L.off_prob)[i] = sp.st_prob[base * A + i];
            rec[L.off_mask + i] = sp.st_mask[base * A + i];
        }
        for (int i = threadIdx.x; i < Tn * 3; i += blockDim.x) {
            reinterpret_cast<float *>(rec + L.off_root_wdl)[i] = sp.st_wdl[base * 3 + i];
            const int t = i / 3, ft = t + sp.td_steps;             // future_root_wdl (src/game.py:117-127)
            reinterpret_cast<float *>(rec + L.off_future)[i] = (sp.td_steps > 0 && ft < Tn) ? sp.st_wdl[(base + ft) * 3 + (i - t * 3)] : 0.0f;
        }
        for (int t = threadIdx.x; t < Tn; t += blockDim.x) {
            reinterpret_cast<int8_t *>(rec + L.off_winner)[t] = (int8_t)winner;
            const int ste = Tn - t;                               // steps_to_end = T..1 (src/game.py:116)
            reinterpret_cast<int16_t *>(rec + L.off_steps)[t] = (int16_t)ste;
            reinterpret_cast<int16_t *>(rec + L.off_aux)[t] = (int16_t)(G::GAME == GAME_OTH ? diff * (int)sp.st_player[base + t] : ste);   // src/game.py:17-23
        }
        // terminal tuple (src/game.py:135-148): end state, zero prob, winner, 0, terminal aux, zero wdl, all-ones mask
        const uint64_t own_bb = s.turn == 1 ? s.bb[0] : s.bb[1], opp_bb = s.turn == 1 ? s.bb[1] : s.bb[0];
        for (int j = threadIdx.x; j < S; j += blockDim.x) {
            const int bit = G::cell_bit(j);
            int8_t *ps = reinterpret_cast<int8_t *>(rec + L.off_state) + (size_t)Tn * 3 * S;
            ps[j] = (int8_t)((own_bb >> bit) & 1ULL); ps[S + j] = (int8_t)((opp_bb >> bit) & 1ULL); ps[2 * S + j] = (int8_t)s.turn;
        }
        for (int a = threadIdx.x; a < A; a += blockDim.x) rec[L.off_mask + Tn * A + a] = 1;
        if (threadIdx.x == 0) {
            reinterpret_cast<int8_t *>(rec + L.off_winner)[Tn] = (int8_t)winner;
            reinterpret_cast<int16_t *>(rec + L.off_aux)[Tn] = (int16_t)(G::GAME == GAME_OTH ? diff * s.turn : 0);   // src/game.py:25-30
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {                              // restart the slot with a fresh game
        State n0; G::reset(n0);
        az_root nr; nr.bb0 = n0.bb[0]; nr.bb1 = n0.bb[1]; nr.turn = n0.turn; nr.passes = n0.passes; nr.last = n0.last; nr.reserved = 0;
        sp.states[g] = nr;
        sp.steps[g] = 0;
        sp.uids[g] += sp.uid_stride;
        sp.finished[g] = 0;
    }
}

}  // namespace az

using namespace az;

extern "C" {

int az_selfplay_layout_for(int game, az_selfplay_layout *out) {
    if (game == GAME_C4) *out = make_layout<C4>();
    else if (game == GAME_OTH) *out = make_layout<Oth>();
    else return AZ_ERR_INVALID;
    return AZ_OK;
}
int az_selfplay_ply_dev(const az_selfplay *sp, const int32_t *counts, const float *stats, void *stream) {
    const int g = (sp->n + 127) / 128;
    if (sp->game == GAME_C4) k_sp_ply<C4><<<g, 128, 0, (cudaStream_t)stream>>>(*sp, counts, stats);
    else if (sp->game == GAME_OTH) k_sp_ply<Oth><<<g, 128, 0, (cudaStream_t)stream>>>(*sp, counts, stats);
    else return AZ_ERR_INVALID;
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}
int az_selfplay_flush_dev(const az_selfplay *sp, void *stream) {
    if (sp->game == GAME_C4) k_sp_flush<C4><<<sp->n, 128, 0, (cudaStream_t)stream>>>(*sp);
    else if (sp->game == GAME_OTH) k_sp_flush<Oth><<<sp->n, 128, 0, (cudaStream_t)stream>>>(*sp);
    else return AZ_ERR_INVALID;
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}

}  // extern "C"
