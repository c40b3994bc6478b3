// Device evaluation cache: open hash table in HBM keyed by (bitboards, side to move) of the symmetrised leaf.
// One thread per leaf.  Entries are claimed with an atomic lock word during insertion, so two leaves hashing to the
// same entry in one batch cannot interleave their fields; lookups run in a different launch and are read-only.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/azb200_cache.h"
#include "az_games.cuh"
#include "az_rng.cuh"

namespace az {

template <int A> struct __align__(16) CacheEntry {
    uint64_t bb0, bb1;
    int32_t turn;          // 0 = empty, else +1 / -1
    uint32_t lock;
    float aux;
    float wdl[3];
    float probs[A];
};

__device__ __forceinline__ uint64_t leaf_hash(uint64_t bb0, uint64_t bb1, int turn) {
    return splitmix64(bb0 ^ splitmix64(bb1 + 0x632BE59BD9B4E019ULL) ^ (turn == 1 ? 0x9E3779B97F4A7C15ULL : 0xC2B2AE3D27D4EB4FULL));
}

template <int A>
__global__ void k_cache_lookup(CacheEntry<A> *tab, uint64_t mask, int n, const az_leaf *__restrict__ leaves, float *probs, float *wdl, float *aux,
                               int32_t *miss_idx, int32_t *miss_count, unsigned long long *stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const az_leaf L = leaves[i];
    if (L.flags & AZ_LEAF_TERMINAL) return;
    const CacheEntry<A> *e = tab + (leaf_hash(L.bb0, L.bb1, L.turn) & mask);
    const bool hit = e->turn == (int32_t)L.turn && e->bb0 == L.bb0 && e->bb1 == L.bb1;
    if (hit) {
        for (int a = 0; a < A; ++a) probs[(size_t)i * A + a] = e->probs[a];
        wdl[3 * i] = e->wdl[0]; wdl[3 * i + 1] = e->wdl[1]; wdl[3 * i + 2] = e->wdl[2];
        aux[i] = e->aux;
    } else miss_idx[atomicAdd(miss_count, 1)] = i;
    atomicAdd(stats + 0, 1ULL);
    if (hit) atomicAdd(stats + 1, 1ULL);
}

template <int A>
__global__ void k_cache_insert(CacheEntry<A> *tab, uint64_t mask, int m, const az_leaf *__restrict__ leaves, const int32_t *__restrict__ miss_idx,
                               const float *__restrict__ pm, const float *__restrict__ wm, const float *__restrict__ am, float *probs, float *wdl,
                               float *aux, unsigned long long *stats) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int i = miss_idx[j];
    for (int a = 0; a < A; ++a) probs[(size_t)i * A + a] = pm[(size_t)j * A + a];
    wdl[3 * i] = wm[3 * j]; wdl[3 * i + 1] = wm[3 * j + 1]; wdl[3 * i + 2] = wm[3 * j + 2];
    aux[i] = am[j];
    const az_leaf L = leaves[i];
    CacheEntry<A> *e = tab + (leaf_hash(L.bb0, L.bb1, L.turn) & mask);
    if (atomicCAS(&e->lock, 0u, 1u) != 0u) return;        // somebody else is writing this entry in this batch: skip
    e->turn = 0;                                           // invalidate while the fields change
    __threadfence();
    e->bb0 = L.bb0; e->bb1 = L.bb1; e->aux = am[j];
    e->wdl[0] = wm[3 * j]; e->wdl[1] = wm[3 * j + 1]; e->wdl[2] = wm[3 * j + 2];
    for (int a = 0; a < A; ++a) e->probs[a] = pm[(size_t)j * A + a];
    __threadfence();
    e->turn = L.turn;
    __threadfence();
    atomicExch(&e->lock, 0u);
    atomicAdd(stats + 2, 1ULL);
}

}  // namespace az

using namespace az;

struct az_evalcache {
    int game, device;
    uint64_t capacity;
    void *table = nullptr;
    size_t entry_bytes;
    unsigned long long *d_stats = nullptr;
};

extern "C" {

az_evalcache *az_evalcache_create(int game, int capacity_log2, int device) {
    if ((game != GAME_C4 && game != GAME_OTH) || capacity_log2 < 4 || capacity_log2 > 30) return nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    az_evalcache *c = new az_evalcache();
    c->game = game; c->device = device; c->capacity = 1ULL << capacity_log2;
    c->entry_bytes = game == GAME_C4 ? sizeof(CacheEntry<7>) : sizeof(CacheEntry<65>);
    if (cudaMalloc(&c->table, c->entry_bytes * c->capacity) != cudaSuccess || cudaMalloc((void **)&c->d_stats, 4 * sizeof(unsigned long long)) != cudaSuccess) {
        az_evalcache_destroy(c);
        return nullptr;
    }
    cudaMemset(c->table, 0, c->entry_bytes * c->capacity);
    cudaMemset(c->d_stats, 0, 4 * sizeof(unsigned long long));
    return c;
}
void az_evalcache_destroy(az_evalcache *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->table) cudaFree(c->table);
    if (c->d_stats) cudaFree(c->d_stats);
    delete c;
}
int az_evalcache_clear_dev(az_evalcache *c, void *stream) {
    return cudaMemsetAsync(c->table, 0, c->entry_bytes * c->capacity, (cudaStream_t)stream) == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}
int az_evalcache_lookup_dev(az_evalcache *c, int n, const az_leaf *leaves, float *probs, float *wdl, float *aux, int32_t *miss_idx, int32_t *miss_count,
                            void *stream) {
    if (n <= 0) return AZ_OK;
    const int g = (n + 127) / 128;
    if (c->game == GAME_C4) k_cache_lookup<7><<<g, 128, 0, (cudaStream_t)stream>>>((CacheEntry<7> *)c->table, c->capacity - 1, n, leaves, probs, wdl, aux, miss_idx, miss_count, c->d_stats);
    else k_cache_lookup<65><<<g, 128, 0, (cudaStream_t)stream>>>((CacheEntry<65> *)c->table, c->capacity - 1, n, leaves, probs, wdl, aux, miss_idx, miss_count, c->d_stats);
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}
int az_evalcache_insert_dev(az_evalcache *c, int m, const az_leaf *leaves, const int32_t *miss_idx, const float *pm, const float *wm, const float *am,
                            float *probs, float *wdl, float *aux, void *stream) {
    if (m <= 0) return AZ_OK;
    const int g = (m + 127) / 128;
    if (c->game == GAME_C4) k_cache_insert<7><<<g, 128, 0, (cudaStream_t)stream>>>((CacheEntry<7> *)c->table, c->capacity - 1, m, leaves, miss_idx, pm, wm, am, probs, wdl, aux, c->d_stats);
    else k_cache_insert<65><<<g, 128, 0, (cudaStream_t)stream>>>((CacheEntry<65> *)c->table, c->capacity - 1, m, leaves, miss_idx, pm, wm, am, probs, wdl, aux, c->d_stats);
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}
int az_evalcache_stats(az_evalcache *c, uint64_t *out4) {
    unsigned long long v[4];
    if (cudaSetDevice(c->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return AZ_ERR_CUDA;
    if (cudaMemcpy(v, c->d_stats, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return AZ_ERR_CUDA;
    out4[0] = v[0]; out4[1] = v[1]; out4[2] = v[2]; out4[3] = c->capacity;
    return AZ_OK;
}

}  // extern "C"
