// Synthetic deterministic leaf evaluators on device pointers: CUDA twins of alphazero-al_b200/evaluators.py
// (HashEvaluator).  They stand in for CNN.predict (src/environments/Connect4/Network.py:267-288) plus the
// wrapper's rel->abs WDL conversion (src/MCTS_cpp.py:23-30, 275-297) when a test or the benchmark needs
// bit-reproducible priors without a host round trip.  They consume the engine's 32-byte az_leaf records directly
// (the hash is defined on the bitboards, so no byte board is ever materialised).  One thread per leaf; all
// arithmetic is exact in fp32 (small integers times powers of two and one IEEE division), so numpy and CUDA agree
// bit for bit.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/azb200.h"
#include "az_games.cuh"
#include "az_rng.cuh"

namespace az {

__device__ __forceinline__ uint64_t board_key(uint64_t own1, uint64_t own2, int turn) {
    const uint64_t t = turn == 1 ? 0x5555555555555555ULL : 0xAAAAAAAAAAAAAAAAULL;
    return splitmix64(own1 ^ splitmix64(own2 + 0x9E3779B97F4A7C15ULL) ^ t);
}

template <class G>
__global__ void k_eval_synth(int mode, int n, const az_leaf *__restrict__ leaves, float *__restrict__ policy, float *__restrict__ dv,
                             float *__restrict__ p1v, float *__restrict__ p2v, float *__restrict__ mlv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();                                      // the leaves are the previous kernel's output
    if (i >= n) return;
    constexpr int A = G::A;
    const uint64_t keep = l2_keep_policy();          // leaves / policy / value rows live in L2 (az_rng.cuh)
    az_leaf L;
    *reinterpret_cast<uint4 *>(&L) = ld_u4_keep(leaves + i, keep);
    *(reinterpret_cast<uint4 *>(&L) + 1) = ld_u4_keep(reinterpret_cast<const uint4 *>(leaves + i) + 1, keep);
    float *prow = policy + (size_t)i * A;
    if (L.flags & AZ_LEAF_TERMINAL) {
        for (int a = 0; a < A; ++a) st_f32_keep(prow + a, 0.0f, keep);
        const bool w1 = (L.flags & AZ_LEAF_P1_WINS) != 0, w2 = (L.flags & AZ_LEAF_P2_WINS) != 0;
        st_f32_keep(dv + i, (!w1 && !w2) ? 1.0f : 0.0f, keep); st_f32_keep(p1v + i, w1 ? 1.0f : 0.0f, keep);
        st_f32_keep(p2v + i, w2 ? 1.0f : 0.0f, keep); st_f32_keep(mlv + i, 0.0f, keep);
        return;
    }
    const int turn = L.turn;
    float wdl0, wdl1, wdl2, aux;
    if (mode == 2) {   // constant
        for (int a = 0; a < A; ++a) st_f32_keep(prow + a, 1.0f, keep);
        wdl0 = 0.25f; wdl1 = 0.5f; wdl2 = 0.25f;
        aux = G::GAME == GAME_C4 ? 10.0f : 0.125f;
    } else {
        const uint64_t own1 = L.bb0, own2 = L.bb1;
        bool canon = true, selfsym = false;
        uint64_t k1 = own1, k2 = own2;
        if (mode == 1) {   // flip-equivariant (Connect4 only)
            const uint64_t f1 = C4::flip_bb(own1), f2 = C4::flip_bb(own2);
            canon = (own1 < f1) || (own1 == f1 && own2 <= f2);
            selfsym = own1 == f1 && own2 == f2;
            if (!canon) { k1 = f1; k2 = f2; }
        }
        const uint64_t h = board_key(k1, k2, turn);
        for (int a = 0; a < A; ++a) {
            int ac = a;
            if (mode == 1) {
                if (!canon) ac = A - 1 - a;
                if (selfsym) ac = min(ac, A - 1 - ac);
            }
            const uint64_t ph = splitmix64(h + (uint64_t)ac + 1ULL);
            st_f32_keep(prow + a, ((float)((ph >> 40) & 0xFFFFULL) + 1.0f) * (1.0f / 65536.0f), keep);
        }
        const float w0 = (float)((splitmix64(h ^ 0x1111ULL) >> 40) & 0xFFULL) + 1.0f;
        const float w1 = (float)((splitmix64(h ^ 0x2222ULL) >> 40) & 0xFFULL) + 1.0f;
        const float w2 = (float)((splitmix64(h ^ 0x3333ULL) >> 40) & 0xFFULL) + 1.0f;
        const float s = (w0 + w1) + w2;
        wdl0 = w0 / s; wdl1 = w1 / s; wdl2 = w2 / s;
        if (G::GAME == GAME_C4) aux = (float)((h >> 20) & 31ULL);
        else aux = (float)((h >> 20) & 63ULL) * (1.0f / 32.0f) - 1.0f;
    }
    st_f32_keep(dv + i, wdl0, keep);
    st_f32_keep(p1v + i, turn == 1 ? wdl1 : wdl2, keep);
    st_f32_keep(p2v + i, turn == 1 ? wdl2 : wdl1, keep);
    st_f32_keep(mlv + i, aux, keep);
}

// The same evaluator with LPL lanes per leaf (Othello: 65 policy entries of one splitmix64 each are the bulk of the work; one
// thread per leaf made this stand-in for the network 10 us of every 68 us iteration at 4096 trees).  Lane l fills actions l, l + LPL, ...;
// lane 0 also writes the value tuple.  Same bits as k_eval_synth.
template <class G, int LPL>
__global__ void k_eval_synth_wide(int mode, int n, const az_leaf *__restrict__ leaves, float *__restrict__ policy, float *__restrict__ dv,
                                  float *__restrict__ p1v, float *__restrict__ p2v, float *__restrict__ mlv) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t / LPL, lane = t % LPL;
    pdl_wait();
    if (i >= n) return;
    constexpr int A = G::A;
    const uint64_t keep = l2_keep_policy();
    az_leaf L;
    *reinterpret_cast<uint4 *>(&L) = ld_u4_keep(leaves + i, keep);
    *(reinterpret_cast<uint4 *>(&L) + 1) = ld_u4_keep(reinterpret_cast<const uint4 *>(leaves + i) + 1, keep);
    float *prow = policy + (size_t)i * A;
    if (L.flags & AZ_LEAF_TERMINAL) {
        for (int a = lane; a < A; a += LPL) st_f32_keep(prow + a, 0.0f, keep);
        if (lane == 0) {
            const bool w1 = (L.flags & AZ_LEAF_P1_WINS) != 0, w2 = (L.flags & AZ_LEAF_P2_WINS) != 0;
            st_f32_keep(dv + i, (!w1 && !w2) ? 1.0f : 0.0f, keep); st_f32_keep(p1v + i, w1 ? 1.0f : 0.0f, keep);
            st_f32_keep(p2v + i, w2 ? 1.0f : 0.0f, keep); st_f32_keep(mlv + i, 0.0f, keep);
        }
        return;
    }
    const int turn = L.turn;
    float wdl0, wdl1, wdl2, aux;
    if (mode == 2) {   // constant
        for (int a = lane; a < A; a += LPL) st_f32_keep(prow + a, 1.0f, keep);
        wdl0 = 0.25f; wdl1 = 0.5f; wdl2 = 0.25f;
        aux = G::GAME == GAME_C4 ? 10.0f : 0.125f;
    } else {           // hash (mode 1, the flip-equivariant variant, is Connect4 only and stays on the one-thread kernel)
        const uint64_t h = board_key(L.bb0, L.bb1, turn);
        for (int a = lane; a < A; a += LPL) {
            const uint64_t ph = splitmix64(h + (uint64_t)a + 1ULL);
            st_f32_keep(prow + a, ((float)((ph >> 40) & 0xFFFFULL) + 1.0f) * (1.0f / 65536.0f), keep);
        }
        if (lane != 0) return;
        const float w0 = (float)((splitmix64(h ^ 0x1111ULL) >> 40) & 0xFFULL) + 1.0f;
        const float w1 = (float)((splitmix64(h ^ 0x2222ULL) >> 40) & 0xFFULL) + 1.0f;
        const float w2 = (float)((splitmix64(h ^ 0x3333ULL) >> 40) & 0xFFULL) + 1.0f;
        const float s = (w0 + w1) + w2;
        wdl0 = w0 / s; wdl1 = w1 / s; wdl2 = w2 / s;
        if (G::GAME == GAME_C4) aux = (float)((h >> 20) & 31ULL);
        else aux = (float)((h >> 20) & 63ULL) * (1.0f / 32.0f) - 1.0f;
    }
    if (lane != 0) return;
    st_f32_keep(dv + i, wdl0, keep);
    st_f32_keep(p1v + i, turn == 1 ? wdl1 : wdl2, keep);
    st_f32_keep(p2v + i, turn == 1 ? wdl2 : wdl1, keep);
    st_f32_keep(mlv + i, aux, keep);
}

// CNN.predict outputs -> backprop tuple: relative WDL [draw, win(to move), loss(to move)] to absolute [draw, p1w, p2w]
// (src/MCTS_cpp.py:23-30) with terminal leaves overridden by their cached result (src/MCTS_cpp.py:276-297).
__global__ void k_eval_finalize(int n, const az_leaf *__restrict__ leaves, const float *__restrict__ wdl_rel, const float *__restrict__ aux,
                                float *__restrict__ dv, float *__restrict__ p1v, float *__restrict__ p2v, float *__restrict__ mlv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t flags = leaves[i].flags;
    const int turn = leaves[i].turn;
    if (flags & AZ_LEAF_TERMINAL) {
        const bool w1 = (flags & AZ_LEAF_P1_WINS) != 0, w2 = (flags & AZ_LEAF_P2_WINS) != 0;
        dv[i] = (!w1 && !w2) ? 1.0f : 0.0f; p1v[i] = w1 ? 1.0f : 0.0f; p2v[i] = w2 ? 1.0f : 0.0f; mlv[i] = 0.0f;
        return;
    }
    const float d = wdl_rel[3 * i], w = wdl_rel[3 * i + 1], l = wdl_rel[3 * i + 2];
    dv[i] = d; p1v[i] = turn == 1 ? w : l; p2v[i] = turn == 1 ? l : w; mlv[i] = aux[i];
}

}  // namespace az

extern "C" int az_eval_finalize_dev(int n, const az_leaf *leaves, const float *wdl_rel, const float *aux, float *d, float *p1, float *p2,
                                    float *ml, void *stream) {
    if (n <= 0) return AZ_OK;
    az::k_eval_finalize<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, leaves, wdl_rel, aux, d, p1, p2, ml);
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}

extern "C" int az_eval_synthetic_dev(int game, int mode, int n, const az_leaf *leaves, float *pol, float *d, float *p1, float *p2, float *ml,
                                     void *stream) {
    if (n <= 0) return AZ_OK;
    if (mode < 0 || mode > 2 || (mode == 1 && game != AZ_GAME_CONNECT4)) return AZ_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const int bs = 128, g = (n + bs - 1) / bs;
    static const bool pdl = !(getenv("AZB200_PDL") && atoi(getenv("AZB200_PDL")) == 0);
    if (game == AZ_GAME_CONNECT4) az::launch_pdl(az::k_eval_synth<az::C4>, g, bs, 0, s, pdl, mode, n, leaves, pol, d, p1, p2, ml);
    else if (game == AZ_GAME_OTHELLO) {
        const int gw = (int)(((size_t)n * 16 + bs - 1) / bs);
        az::launch_pdl(az::k_eval_synth_wide<az::Oth, 16>, gw, bs, 0, s, pdl, mode, n, leaves, pol, d, p1, p2, ml);
    }
    else return AZ_ERR_INVALID;
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}
