/*
 * azb200_cache.h - device evaluation cache: the reference keeps an LRU "transposition table" of
 * network evaluations in Python (src/Cache.py:5-58, used by src/MCTS_cpp.py:146-189, 299-339: key = symmetrised leaf
 * board bytes + turn byte, value = (probs[A], wdl_rel[3], moves_left)).  Here the table lives in HBM and is probed by
 * one kernel per iteration, so only cache misses reach the network.  A cached value is exactly what the network
 * returns for that input, hence results do not depend on capacity or replacement policy (always-replace here, LRU in
 * the reference); `az_evalcache_clear_dev` is the equivalent of refresh_cache after a weight reload
 * (src/MCTS_cpp.py:361-377) and of the score_scale invalidation (src/MCTS_cpp.py:402-409).
 */
#ifndef AZB200_CACHE_H
#define AZB200_CACHE_H
#include "azb200.h"

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

typedef struct az_evalcache az_evalcache;

az_evalcache *az_evalcache_create(int game, int capacity_log2, int device);
void az_evalcache_destroy(az_evalcache *c);
int az_evalcache_clear_dev(az_evalcache *c, void *stream);
/* For every non-terminal leaf: hit  -> its cached (probs[A], wdl_rel[3], aux) are written to row i of the outputs;
 *                              miss -> i is appended to d_miss_idx (order unspecified) and *d_miss_count incremented.
 * Terminal leaves are neither (az_eval_finalize_dev overrides them).  *d_miss_count must be zero on entry. */
int az_evalcache_lookup_dev(az_evalcache *c, int n_leaves, const az_leaf *d_leaves, float *d_probs, float *d_wdl_rel,
                            float *d_aux, int32_t *d_miss_idx, int32_t *d_miss_count, void *stream);
/* De-duplicating lookup ("dedup of identical leaves within a batch"; the reference evaluates in-batch
 * duplicates separately, src/MCTS_cpp.py:299-339).  As az_evalcache_lookup_dev, but of all missing leaves that hold the same
 * position only one is appended to d_miss_idx; the others get d_dup_of[i] = row of that leaf (-1 for every other row) and
 * receive a copy of its outputs from az_evalcache_resolve_dups_dev, to be called after az_evalcache_insert_dev. */
int az_evalcache_lookup_dedup_dev(az_evalcache *c, int n_leaves, const az_leaf *d_leaves, float *d_probs, float *d_wdl_rel,
                                  float *d_aux, int32_t *d_miss_idx, int32_t *d_miss_count, int32_t *d_dup_of, void *stream);
int az_evalcache_resolve_dups_dev(az_evalcache *c, int n_leaves, const int32_t *d_dup_of, float *d_probs, float *d_wdl_rel,
                                  float *d_aux, void *stream);
/* Network outputs of the misses (row j belongs to leaf d_miss_idx[j]) are stored in the table and scattered to rows
 * d_miss_idx[j] of the full outputs. */
int az_evalcache_insert_dev(az_evalcache *c, int n_miss, const az_leaf *d_leaves, const int32_t *d_miss_idx,
                            const float *d_probs_m, const float *d_wdl_m, const float *d_aux_m, float *d_probs,
                            float *d_wdl_rel, float *d_aux, void *stream);
/* out[0..3] = lookups (non-terminal leaves probed), hits, inserts, capacity */
int az_evalcache_stats(az_evalcache *c, uint64_t *out4);
/* *out = leaves that shared another leaf's evaluation (az_evalcache_lookup_dedup_dev) */
int az_evalcache_dups(az_evalcache *c, uint64_t *out);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* AZB200_CACHE_H */
