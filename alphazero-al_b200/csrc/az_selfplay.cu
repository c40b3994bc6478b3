// On-device self-play driver: per-ply policy target / sampling / recording / env step (k_sp_ply), the move of finished
// trajectories into the output ring (k_sp_flush) and the expansion of compact trajectory records into the reference's training
// tuples (k_sp_expand).  Reference behaviour: src/game.py:65-164, src/player.py:333-375, src/ReplayBuffer.py:12-19.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/azb200_selfplay.h"
#include "az_games.cuh"
#include "az_rng.cuh"

namespace az {

template <class G> struct SpFmt {
    static constexpr int POS_BYTES = (32 + 4 * G::A + 31) & ~31;      // az_sp_pos head + float prob[A], padded to a sector multiple
    static constexpr int CHUNKS = POS_BYTES / 16;
};
static_assert(sizeof(az_sp_game) == 32 && sizeof(az_sp_pos) == 32, "compact trajectory records are sector sized");
static_assert(SpFmt<C4>::POS_BYTES == 64 && SpFmt<Oth>::POS_BYTES == 320, "position record sizes stated in azb200_selfplay.h");

__device__ __forceinline__ State sp_state(const az_root &r) { State s; s.bb[0] = r.bb0; s.bb[1] = r.bb1; s.turn = r.turn; s.passes = r.passes; s.last = r.last; return s; }
__device__ __forceinline__ State sp_state(const az_sp_pos &p) { State s; s.bb[0] = p.bb0; s.bb[1] = p.bb1; s.turn = p.turn; s.passes = p.passes; s.last = -1; return s; }

// One thread per game slot.
template <class G> __global__ void k_sp_ply(az_selfplay sp, const int32_t *__restrict__ counts, const float *__restrict__ stats) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= sp.n) return;
    constexpr int A = G::A, PB = SpFmt<G>::POS_BYTES;
    const int T = sp.max_plies;
    az_root r = sp.states[g];
    State s = sp_state(r);
    const int step = sp.steps[g];
    const uint64_t uid = sp.uids[g];
    const int32_t *c = counts + (size_t)g * A;
    long long total = 0; int best = 0, bestv = -1;
    for (int a = 0; a < A; ++a) { const int v = c[a]; total += v; if (v > bestv) { bestv = v; best = a; } }   // np.argmax: first maximum
    int action = 0;
    if (total > 0) {
        const float temp = (sp.temp_decay_moves <= 0 || step < sp.temp_decay_moves) ? sp.temp_init : sp.temp_endgame;   // src/game.py:54-63
        if (temp <= 1e-6f) action = best;
        else {   // softmax(log(visits)/temp) over visited actions (src/player.py:364-369); RNG stream is ours, keyed by the game uid
            const double invt = 1.0 / (double)temp, lmax = log((double)bestv) * invt;
            double z = 0.0;
            for (int a = 0; a < A; ++a) if (c[a] > 0) z += exp(log((double)c[a]) * invt - lmax);
            const uint64_t h = az_rand(sp.seed, (uint64_t)step, 3, uid, 0);
            double u = ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0) * z, acc = 0.0;
            action = best;
            for (int a = 0; a < A; ++a) if (c[a] > 0) { acc += exp(log((double)c[a]) * invt - lmax); if (u < acc) { action = a; break; } }
        }
    }
    if (sp.forced && uid >= sp.forced_uid0 && uid - sp.forced_uid0 < (uint64_t)sp.forced_games && step < T) {   // opening script
        const int f = sp.forced[(uid - sp.forced_uid0) * (uint64_t)T + step];
        if (f >= 0) action = f;
    }
    // ---- record the position (src/game.py:101-110): bitboards + side to move, policy target, root WDL ----
    if (step < T) {
        uint8_t *rec = sp.st_pos + ((size_t)g * T + step) * PB;
        const float *st = stats + (size_t)g * (6 + 8 * A);
        az_sp_pos h;
        h.bb0 = s.bb[0]; h.bb1 = s.bb[1];
        h.root_wdl[0] = st[3]; h.root_wdl[1] = st[4]; h.root_wdl[2] = st[5];         // root_D, root_P1W, root_P2W
        h.turn = (int8_t)s.turn; h.passes = (uint8_t)s.passes; h.reserved[0] = h.reserved[1] = 0;
        *reinterpret_cast<uint4 *>(rec) = *reinterpret_cast<const uint4 *>(&h);
        *reinterpret_cast<uint4 *>(rec + 16) = *(reinterpret_cast<const uint4 *>(&h) + 1);
        float *pp = reinterpret_cast<float *>(rec + 32);
        for (int a = 0; a < A; ++a) pp[a] = total > 0 ? (float)((double)c[a] / (double)total) : 0.0f;   // visits / visits.sum() -> float32
        for (int a = A; a < (PB - 32) / 4; ++a) pp[a] = 0.0f;
    }
    // ---- env.step + done (src/game.py:112-113) ----
    G::step(s, action);
    r.bb0 = s.bb[0]; r.bb1 = s.bb[1]; r.turn = s.turn; r.passes = s.passes; r.last = s.last;
    sp.states[g] = r;
    sp.steps[g] = step + 1;
    const bool done = G::done(s) || step + 1 >= T;
    if (done) sp.fin_list[atomicAdd(sp.fin_count, 1)] = g;
    sp.actions[g] = done ? -1 : action;      // finished game: the tree is reset (src/game.py:158, reset_env)
}

// One CTA per finished slot (grid-stride over the list k_sp_ply built).
template <class G> __global__ void k_sp_flush(az_selfplay sp) {
    constexpr int PB = SpFmt<G>::POS_BYTES, CH = SpFmt<G>::CHUNKS;
    const int T = sp.max_plies;
    const int nfin = *sp.fin_count;
    __shared__ long long slot_s[2];
    for (int f = blockIdx.x; f < nfin; f += gridDim.x) {
        const int g = sp.fin_list[f];
        const int Tn = sp.steps[g];                        // plies played = recorded positions before the terminal one
        const az_root r = sp.states[g];
        const State s = sp_state(r);
        if (threadIdx.x == 0) {
            const unsigned long long gs = atomicAdd(&sp.out_counters[0], 1ULL);
            long long ps = -1;
            if (gs < (unsigned long long)sp.game_capacity) {
                ps = (long long)atomicAdd(&sp.out_counters[1], (unsigned long long)(Tn + 1));
                atomicAdd(&sp.out_counters[3], (unsigned long long)Tn);
            } else atomicAdd(&sp.out_counters[2], 1ULL);   // ring full: the game is dropped and counted (drain more often)
            slot_s[0] = (long long)gs; slot_s[1] = ps;
        }
        __syncthreads();
        const long long gs = slot_s[0], ps = slot_s[1];
        if (ps >= 0) {
            const uint4 *src = reinterpret_cast<const uint4 *>(sp.st_pos + (size_t)g * T * PB);
            uint4 *dst = reinterpret_cast<uint4 *>(sp.out_pos + (size_t)ps * PB);
            for (int i = threadIdx.x; i < Tn * CH; i += blockDim.x) dst[i] = src[i];
            // terminal position (src/game.py:135-148): the end state; zero policy and root WDL
            for (int i = threadIdx.x; i < CH; i += blockDim.x) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (i < 2) {
                    az_sp_pos h;
                    h.bb0 = s.bb[0]; h.bb1 = s.bb[1]; h.root_wdl[0] = h.root_wdl[1] = h.root_wdl[2] = 0.0f;
                    h.turn = (int8_t)s.turn; h.passes = (uint8_t)s.passes; h.reserved[0] = h.reserved[1] = 0;
                    v = *(reinterpret_cast<const uint4 *>(&h) + i);
                }
                dst[Tn * CH + i] = v;
            }
            if (threadIdx.x == 0) {
                az_sp_game hdr;
                hdr.uid = sp.uids[g]; hdr.pos_start = ps; hdr.length = Tn + 1; hdr.winner = G::winner(s); hdr.reserved[0] = hdr.reserved[1] = 0;
                uint4 *hp = reinterpret_cast<uint4 *>(sp.out_games + gs);
                hp[0] = *reinterpret_cast<const uint4 *>(&hdr); hp[1] = *(reinterpret_cast<const uint4 *>(&hdr) + 1);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {                              // restart the slot with a fresh game
            State n0; G::reset(n0);
            az_root nr; nr.bb0 = n0.bb[0]; nr.bb1 = n0.bb[1]; nr.turn = n0.turn; nr.passes = n0.passes; nr.last = n0.last; nr.reserved = 0;
            sp.states[g] = nr;
            sp.steps[g] = 0;
            sp.uids[g] += sp.uid_stride;
        }
        __syncthreads();
    }
}

// One CTA per game: compact records -> training tuples (rows of the replay-buffer tensors).
template <class G>
__global__ void k_sp_expand(int n_games, const az_sp_game *__restrict__ games, const uint8_t *__restrict__ pos, int td_steps,
                            int8_t *__restrict__ o_state, float *__restrict__ o_prob, int8_t *__restrict__ o_winner,
                            int16_t *__restrict__ o_steps, int16_t *__restrict__ o_aux, float *__restrict__ o_wdl,
                            uint8_t *__restrict__ o_mask, float *__restrict__ o_future) {
    constexpr int A = G::A, S = G::S, PB = SpFmt<G>::POS_BYTES;
    const int gi = blockIdx.x;
    if (gi >= n_games) return;
    const az_sp_game gm = games[gi];
    const int L = gm.length, Tn = L - 1;                   // Tn moves, L positions
    const uint8_t *base = pos + (size_t)gm.pos_start * PB;
    const az_sp_pos last = *reinterpret_cast<const az_sp_pos *>(base + (size_t)Tn * PB);
    const int diff = popc64(last.bb0) - popc64(last.bb1);  // final disc difference (src/game.py:17-30)
    for (int t = 0; t < L; ++t) {
        const az_sp_pos h = *reinterpret_cast<const az_sp_pos *>(base + (size_t)t * PB);
        const size_t row = (size_t)gm.pos_start + t;
        const bool terminal = t == Tn;
        if (o_state) {
            const uint64_t own = h.turn == 1 ? h.bb0 : h.bb1, opp = h.turn == 1 ? h.bb1 : h.bb0;
            int8_t *ps = o_state + row * 3 * S;
            for (int j = threadIdx.x; j < S; j += blockDim.x) {
                const int bit = G::cell_bit(j);
                ps[j] = (int8_t)((own >> bit) & 1ULL); ps[S + j] = (int8_t)((opp >> bit) & 1ULL); ps[2 * S + j] = h.turn;
            }
        }
        if (o_prob) {
            const float *pp = reinterpret_cast<const float *>(base + (size_t)t * PB + 32);
            for (int a = threadIdx.x; a < A; a += blockDim.x) o_prob[row * A + a] = terminal ? 0.0f : pp[a];
        }
        if (o_mask) {
            const State s = sp_state(h);
            const uint64_t legal = terminal ? 0ULL : G::legal(s);
            const bool pass_only = G::GAME == GAME_OTH && !terminal && legal == 0ULL;
            for (int a = threadIdx.x; a < A; a += blockDim.x)
                o_mask[row * A + a] = terminal ? 1 : ((G::GAME == GAME_OTH && a == 64) ? (pass_only ? 1 : 0) : (uint8_t)((legal >> (a & 63)) & 1ULL));
        }
        if (threadIdx.x < 3) {
            if (o_wdl) o_wdl[row * 3 + threadIdx.x] = terminal ? 0.0f : h.root_wdl[threadIdx.x];
            if (o_future) {                                 // future_root_wdl (src/game.py:117-127): root_wdl[t + k] while t + k < T
                const int ft = t + td_steps;
                float v = 0.0f;
                if (!terminal && td_steps > 0 && ft < Tn) v = reinterpret_cast<const az_sp_pos *>(base + (size_t)ft * PB)->root_wdl[threadIdx.x];
                o_future[row * 3 + threadIdx.x] = v;
            }
        }
        if (threadIdx.x == 0) {
            const int ste = terminal ? 0 : Tn - t;          // steps_to_end = T..1, terminal tuple 0 (src/game.py:116,140)
            if (o_winner) o_winner[row] = (int8_t)gm.winner;
            if (o_steps) o_steps[row] = (int16_t)ste;
            if (o_aux) o_aux[row] = (int16_t)(G::GAME == GAME_OTH ? diff * (int)h.turn : ste);
        }
    }
}

}  // namespace az

using namespace az;

extern "C" {

int az_selfplay_pos_bytes(int game) { return game == GAME_C4 ? SpFmt<C4>::POS_BYTES : (game == GAME_OTH ? SpFmt<Oth>::POS_BYTES : -1); }
int az_selfplay_max_plies(int game) { return game == GAME_C4 ? C4::MAX_PLIES : (game == GAME_OTH ? Oth::MAX_PLIES : -1); }

int az_selfplay_ply_dev(const az_selfplay *sp, const int32_t *counts, const float *stats, void *stream) {
    if (sp->pos_bytes != az_selfplay_pos_bytes(sp->game) || sp->max_plies != az_selfplay_max_plies(sp->game)) return AZ_ERR_INVALID;
    const int g = (sp->n + 127) / 128;
    if (cudaMemsetAsync(sp->fin_count, 0, sizeof(int32_t), (cudaStream_t)stream) != cudaSuccess) return AZ_ERR_CUDA;
    if (sp->game == GAME_C4) k_sp_ply<C4><<<g, 128, 0, (cudaStream_t)stream>>>(*sp, counts, stats);
    else k_sp_ply<Oth><<<g, 128, 0, (cudaStream_t)stream>>>(*sp, counts, stats);
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}
int az_selfplay_flush_dev(const az_selfplay *sp, void *stream) {
    if (sp->pos_bytes != az_selfplay_pos_bytes(sp->game)) return AZ_ERR_INVALID;
    const int g = sp->n < 148 * 8 ? sp->n : 148 * 8;       // finished games per ply are a few percent of the slots
    if (sp->game == GAME_C4) k_sp_flush<C4><<<g, 64, 0, (cudaStream_t)stream>>>(*sp);
    else k_sp_flush<Oth><<<g, 64, 0, (cudaStream_t)stream>>>(*sp);
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}
int az_selfplay_expand_dev(int game, int n_games, const az_sp_game *games, const uint8_t *pos, int td_steps, int8_t *o_state, float *o_prob,
                           int8_t *o_winner, int16_t *o_steps, int16_t *o_aux, float *o_wdl, uint8_t *o_mask, float *o_future, void *stream) {
    if (n_games <= 0) return AZ_OK;
    if (game == GAME_C4) k_sp_expand<C4><<<n_games, 64, 0, (cudaStream_t)stream>>>(n_games, games, pos, td_steps, o_state, o_prob, o_winner, o_steps, o_aux, o_wdl, o_mask, o_future);
    else if (game == GAME_OTH) k_sp_expand<Oth><<<n_games, 64, 0, (cudaStream_t)stream>>>(n_games, games, pos, td_steps, o_state, o_prob, o_winner, o_steps, o_aux, o_wdl, o_mask, o_future);
    else return AZ_ERR_INVALID;
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}

}  // extern "C"
