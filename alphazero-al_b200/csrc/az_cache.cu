// Device evaluation cache: open hash table in HBM keyed by (bitboards, side to move) of the symmetrised leaf.
// One thread per leaf.  An entry is claimed through its 64-bit lock word = (epoch << 32 | leaf row + 1), the epoch being the
// number of the lookup launch: a claim of an earlier epoch counts as free, so locks never need releasing and a batch that
// was abandoned between lookup and insert leaves nothing behind.  The de-duplicating lookup claims the entry of every miss:
// a later miss of the same batch that finds the entry claimed compares its key with the claimant's leaf record and, if they
// are equal, becomes a duplicate (evaluated once, copied by k_cache_resolve) - the reference evaluates in-batch duplicates
// separately (src/MCTS_cpp.py:299-339); the network is a pure function of the row, so the results are the same.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/azb200_cache.h"
#include "az_games.cuh"
#include "az_rng.cuh"

namespace az {

template <int A> struct __align__(16) CacheEntry {
    uint64_t bb0, bb1;
    unsigned long long lock;   // epoch << 32 | (row + 1) of the leaf that owns the entry in that epoch
    int32_t turn;              // 0 = empty, else +1 / -1
    float aux;
    float wdl[3];
    float probs[A];
};

__device__ __forceinline__ uint64_t leaf_hash(uint64_t bb0, uint64_t bb1, int turn) {
    return splitmix64(bb0 ^ splitmix64(bb1 + 0x632BE59BD9B4E019ULL) ^ (turn == 1 ? 0x9E3779B97F4A7C15ULL : 0xC2B2AE3D27D4EB4FULL));
}

// Claim entry `e` for leaf row i in `epoch`; returns the row that owns it in this epoch (i itself on success).
__device__ __forceinline__ int cache_claim(unsigned long long *lock, uint32_t epoch, int i) {
    const unsigned long long tag = ((unsigned long long)epoch << 32) | (uint32_t)(i + 1);
    unsigned long long old = *(volatile unsigned long long *)lock;
    for (;;) {
        if ((uint32_t)(old >> 32) == epoch) return (int)(uint32_t)old - 1;
        const unsigned long long prev = atomicCAS(lock, old, tag);
        if (prev == old) return i;
        old = prev;
    }
}

// one atomic per warp and counter (a batch of 262 144 leaves would otherwise queue that many atomics on one address)
__device__ __forceinline__ void warp_count(unsigned long long *ctr, bool flag) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, flag);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(ctr, (unsigned long long)__popc(m));
}

// DEDUP: dup_of[i] = row whose evaluation leaf i shares (or -1); only one leaf per distinct position is appended to miss_idx.
template <int A, bool DEDUP>
__global__ void k_cache_lookup(CacheEntry<A> *tab, uint64_t mask, uint32_t epoch, int n, const az_leaf *__restrict__ leaves, float *probs, float *wdl,
                               float *aux, int32_t *miss_idx, int32_t *miss_count, int32_t *dup_of, unsigned long long *stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool looked = false, hit = false, dup = false, miss = false;
    if (i < n) {
        const az_leaf L = leaves[i];
        if (DEDUP) dup_of[i] = -1;
        if (!(L.flags & AZ_LEAF_TERMINAL)) {
            looked = true;
            CacheEntry<A> *e = tab + (leaf_hash(L.bb0, L.bb1, L.turn) & mask);
            hit = e->turn == (int32_t)L.turn && e->bb0 == L.bb0 && e->bb1 == L.bb1;
            if (hit) {
                for (int a = 0; a < A; ++a) probs[(size_t)i * A + a] = e->probs[a];
                wdl[3 * i] = e->wdl[0]; wdl[3 * i + 1] = e->wdl[1]; wdl[3 * i + 2] = e->wdl[2];
                aux[i] = e->aux;
            } else {
                if (DEDUP) {
                    const int owner = cache_claim(&e->lock, epoch, i);
                    if (owner != i) {
                        const az_leaf O = leaves[owner];
                        dup = O.bb0 == L.bb0 && O.bb1 == L.bb1 && O.turn == L.turn && !(O.flags & AZ_LEAF_TERMINAL);
                        if (dup) dup_of[i] = owner;
                    }       // else a different position owns the entry in this batch: evaluate, do not store
                }
                miss = !dup;
            }
        }
    }
    // the miss list: one atomic per warp, lanes take consecutive places
    const unsigned mm = __ballot_sync(0xFFFFFFFFu, miss);
    int base = 0;
    if ((threadIdx.x & 31) == 0 && mm) base = atomicAdd(miss_count, __popc(mm));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (miss) miss_idx[base + __popc(mm & ((1u << (threadIdx.x & 31)) - 1u))] = i;
    warp_count(stats + 0, looked);
    warp_count(stats + 1, hit);
    warp_count(stats + 3, dup);
}

template <int A>
__global__ void k_cache_insert(CacheEntry<A> *tab, uint64_t mask, uint32_t epoch, int m, const az_leaf *__restrict__ leaves,
                               const int32_t *__restrict__ miss_idx, const float *__restrict__ pm, const float *__restrict__ wm,
                               const float *__restrict__ am, float *probs, float *wdl, float *aux, unsigned long long *stats) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    bool stored = false;
    if (j < m) {
        const int i = miss_idx[j];
        for (int a = 0; a < A; ++a) probs[(size_t)i * A + a] = pm[(size_t)j * A + a];
        wdl[3 * i] = wm[3 * j]; wdl[3 * i + 1] = wm[3 * j + 1]; wdl[3 * i + 2] = wm[3 * j + 2];
        aux[i] = am[j];
        const az_leaf L = leaves[i];
        CacheEntry<A> *e = tab + (leaf_hash(L.bb0, L.bb1, L.turn) & mask);
        if (cache_claim(&e->lock, epoch, i) == i) {        // else another leaf of this batch owns the entry: skip
            // lookups run in other launches and every entry has one owner per epoch: plain stores
            e->bb0 = L.bb0; e->bb1 = L.bb1; e->aux = am[j];
            e->wdl[0] = wm[3 * j]; e->wdl[1] = wm[3 * j + 1]; e->wdl[2] = wm[3 * j + 2];
            for (int a = 0; a < A; ++a) e->probs[a] = pm[(size_t)j * A + a];
            e->turn = L.turn;
            stored = true;
        }
    }
    warp_count(stats + 2, stored);
}

// Rows that share another row's evaluation (k_cache_lookup<DEDUP>) copy it; runs after k_cache_insert wrote the owners' rows.
template <int A>
__global__ void k_cache_resolve(int n, const int32_t *__restrict__ dup_of, float *probs, float *wdl, float *aux) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int o = dup_of[i];
    if (o < 0) return;
    for (int a = 0; a < A; ++a) probs[(size_t)i * A + a] = probs[(size_t)o * A + a];
    wdl[3 * i] = wdl[3 * o]; wdl[3 * i + 1] = wdl[3 * o + 1]; wdl[3 * i + 2] = wdl[3 * o + 2];
    aux[i] = aux[o];
}

}  // namespace az

using namespace az;

struct az_evalcache {
    int game, device;
    uint64_t capacity;
    void *table = nullptr;
    size_t entry_bytes;
    unsigned long long *d_stats = nullptr;
    uint32_t epoch = 0;          // number of the last lookup launch; claims of earlier epochs are free
};

template <bool DEDUP>
static int cache_lookup(az_evalcache *c, int n, const az_leaf *leaves, float *probs, float *wdl, float *aux, int32_t *miss_idx, int32_t *miss_count,
                        int32_t *dup_of, cudaStream_t st) {
    if (n <= 0) return AZ_OK;
    if (++c->epoch == 0) {       // 2^32 lookups: claims of the first epochs would look current again
        if (cudaMemsetAsync(c->table, 0, c->entry_bytes * c->capacity, st) != cudaSuccess) return AZ_ERR_CUDA;
        c->epoch = 1;
    }
    const int g = (n + 127) / 128;
    if (c->game == GAME_C4) k_cache_lookup<7, DEDUP><<<g, 128, 0, st>>>((CacheEntry<7> *)c->table, c->capacity - 1, c->epoch, n, leaves, probs, wdl, aux, miss_idx, miss_count, dup_of, c->d_stats);
    else k_cache_lookup<65, DEDUP><<<g, 128, 0, st>>>((CacheEntry<65> *)c->table, c->capacity - 1, c->epoch, n, leaves, probs, wdl, aux, miss_idx, miss_count, dup_of, c->d_stats);
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}

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
    return cache_lookup<false>(c, n, leaves, probs, wdl, aux, miss_idx, miss_count, nullptr, (cudaStream_t)stream);
}
int az_evalcache_lookup_dedup_dev(az_evalcache *c, int n, const az_leaf *leaves, float *probs, float *wdl, float *aux, int32_t *miss_idx,
                                  int32_t *miss_count, int32_t *dup_of, void *stream) {
    if (!dup_of) return AZ_ERR_INVALID;
    return cache_lookup<true>(c, n, leaves, probs, wdl, aux, miss_idx, miss_count, dup_of, (cudaStream_t)stream);
}
int az_evalcache_insert_dev(az_evalcache *c, int m, const az_leaf *leaves, const int32_t *miss_idx, const float *pm, const float *wm, const float *am,
                            float *probs, float *wdl, float *aux, void *stream) {
    if (m <= 0) return AZ_OK;
    const int g = (m + 127) / 128;
    if (c->game == GAME_C4) k_cache_insert<7><<<g, 128, 0, (cudaStream_t)stream>>>((CacheEntry<7> *)c->table, c->capacity - 1, c->epoch, m, leaves, miss_idx, pm, wm, am, probs, wdl, aux, c->d_stats);
    else k_cache_insert<65><<<g, 128, 0, (cudaStream_t)stream>>>((CacheEntry<65> *)c->table, c->capacity - 1, c->epoch, m, leaves, miss_idx, pm, wm, am, probs, wdl, aux, c->d_stats);
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}
int az_evalcache_resolve_dups_dev(az_evalcache *c, int n, const int32_t *dup_of, float *probs, float *wdl, float *aux, void *stream) {
    if (n <= 0) return AZ_OK;
    const int g = (n + 127) / 128;
    if (c->game == GAME_C4) k_cache_resolve<7><<<g, 128, 0, (cudaStream_t)stream>>>(n, dup_of, probs, wdl, aux);
    else k_cache_resolve<65><<<g, 128, 0, (cudaStream_t)stream>>>(n, dup_of, probs, wdl, aux);
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}
int az_evalcache_stats(az_evalcache *c, uint64_t *out4) {
    unsigned long long v[4];
    if (cudaSetDevice(c->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return AZ_ERR_CUDA;
    if (cudaMemcpy(v, c->d_stats, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return AZ_ERR_CUDA;
    out4[0] = v[0]; out4[1] = v[1]; out4[2] = v[2]; out4[3] = c->capacity;
    return AZ_OK;
}
int az_evalcache_dups(az_evalcache *c, uint64_t *out) {
    unsigned long long v = 0;
    if (cudaSetDevice(c->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return AZ_ERR_CUDA;
    if (cudaMemcpy(&v, c->d_stats + 3, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return AZ_ERR_CUDA;
    *out = v;
    return AZ_OK;
}

}  // extern "C"
