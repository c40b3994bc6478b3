// Synthetic deterministic leaf evaluators on device pointers: CUDA twins of alphazero-al_b200/evaluators.py
// (HashEvaluator).  They stand in for CNN.predict (src/environments/Connect4/Network.py:267-288) plus the
// wrapper's rel->abs WDL conversion (src/MCTS_cpp.py:23-30, 275-297) when a test or the benchmark needs
// bit-reproducible priors without a host round trip.  One thread per leaf; all arithmetic is exact in fp32
// (small integers times powers of two and one IEEE division), so numpy and CUDA agree bit for bit.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/azb200.h"
#include "az_games.cuh"
#include "az_rng.cuh"

namespace az {

__device__ __forceinline__ uint64_t board_key(uint64_t own1, uint64_t own2, int turn) {
    const uint64_t t = turn == 1 ? 0x5555555555555555ULL : 0xAAAAAAAAAAAAAAAAULL;
    return splitmix64(own1 ^ splitmix64(own2 + 0x9E3779B97F4A7C15ULL) ^ t);
}

template <int S, int A, int COLS, bool IS_C4>
__global__ void k_eval_synth(int mode, int n, const int8_t *__restrict__ boards, const int32_t *__restrict__ turns,
                             const uint8_t *__restrict__ is_term, const float *__restrict__ td, const float *__restrict__ tp1,
                             const float *__restrict__ tp2, float *__restrict__ policy, float *__restrict__ dv,
                             float *__restrict__ p1v, float *__restrict__ p2v, float *__restrict__ mlv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float *prow = policy + (size_t)i * A;
    if (is_term[i]) {
        for (int a = 0; a < A; ++a) prow[a] = 0.0f;
        dv[i] = td[i]; p1v[i] = tp1[i]; p2v[i] = tp2[i]; mlv[i] = 0.0f;
        return;
    }
    const int turn = turns[i];
    float wdl0, wdl1, wdl2, aux;
    if (mode == 2) {   // constant
        for (int a = 0; a < A; ++a) prow[a] = 1.0f;
        wdl0 = 0.25f; wdl1 = 0.5f; wdl2 = 0.25f;
        aux = IS_C4 ? 10.0f : 0.125f;
    } else {
        const int8_t *b = boards + (size_t)i * S;
        uint64_t own1 = 0, own2 = 0, f1 = 0, f2 = 0;
        for (int j = 0; j < S; ++j) {
            const int v = b[j];
            const int r = j / COLS, c = j - r * COLS, jf = r * COLS + (COLS - 1 - c);
            if (v == 1) { own1 |= 1ULL << j; f1 |= 1ULL << jf; }
            else if (v == -1) { own2 |= 1ULL << j; f2 |= 1ULL << jf; }
        }
        bool canon = true, selfsym = false;
        uint64_t k1 = own1, k2 = own2;
        if (mode == 1) {
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
            prow[a] = ((float)((ph >> 40) & 0xFFFFULL) + 1.0f) * (1.0f / 65536.0f);
        }
        const float w0 = (float)((splitmix64(h ^ 0x1111ULL) >> 40) & 0xFFULL) + 1.0f;
        const float w1 = (float)((splitmix64(h ^ 0x2222ULL) >> 40) & 0xFFULL) + 1.0f;
        const float w2 = (float)((splitmix64(h ^ 0x3333ULL) >> 40) & 0xFFULL) + 1.0f;
        const float s = (w0 + w1) + w2;
        wdl0 = w0 / s; wdl1 = w1 / s; wdl2 = w2 / s;
        if (IS_C4) aux = (float)((h >> 20) & 31ULL);
        else aux = (float)((h >> 20) & 63ULL) * (1.0f / 32.0f) - 1.0f;
    }
    dv[i] = wdl0;
    p1v[i] = turn == 1 ? wdl1 : wdl2;
    p2v[i] = turn == 1 ? wdl2 : wdl1;
    mlv[i] = aux;
}

}  // namespace az

extern "C" int az_eval_synthetic_dev(int game, int mode, int n, const int8_t *b, const int32_t *t, const uint8_t *it, const float *td,
                                     const float *tp1, const float *tp2, float *pol, float *d, float *p1, float *p2, float *ml, void *stream) {
    if (n <= 0) return AZ_OK;
    if (mode < 0 || mode > 2 || (mode == 1 && game != AZ_GAME_CONNECT4)) return AZ_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const int bs = 128, g = (n + bs - 1) / bs;
    if (game == AZ_GAME_CONNECT4) az::k_eval_synth<42, 7, 7, true><<<g, bs, 0, s>>>(mode, n, b, t, it, td, tp1, tp2, pol, d, p1, p2, ml);
    else if (game == AZ_GAME_OTHELLO) az::k_eval_synth<64, 65, 8, false><<<g, bs, 0, s>>>(mode, n, b, t, it, td, tp1, tp2, pol, d, p1, p2, ml);
    else return AZ_ERR_INVALID;
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}
