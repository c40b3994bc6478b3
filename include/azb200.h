/*
 * azb200.h - C ABI of the B200-native self-play hot path (batched PUCT MCTS + bitboard envs).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / pybind types.  Every entry point names
 * the reference interface it replaces (paths relative to /root/reference/).  The reference exposes this path
 * through two pybind11 modules (src/cpp/mcts_bindings.cpp, src/cpp/env_bindings.cpp); INTEGRATION.md shows the
 * ctypes / pybind stub a maintainer adds on the reference side to bind these symbols instead.
 *
 * Conventions
 *   - every function returns AZ_OK (0) or a negative error; az_*_last_error() gives the message.  No C++
 *     exception crosses the ABI.  A handle is not re-entrant (one caller thread at a time, like the reference).
 *   - "host" entry points take HOST pointers (numpy buffers) and copy H2D/D2H inside the call; their `_dev`
 *     twins take CUDA DEVICE pointers plus a cudaStream_t (passed as void*) and never synchronise, so PyTorch
 *     tensors can be handed over by data_ptr() with no copies.  All tree state always lives in HBM; there is no
 *     CPU fallback - creating a handle without a CUDA device fails.
 *   - flat leaf index is env*K + k, exactly as in src/cpp/BatchedMCTS.h:221,251.
 */
#ifndef AZB200_H
#define AZB200_H
#include <stdint.h>

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

#define AZ_OK 0
#define AZ_ERR_INVALID (-1)
#define AZ_ERR_CUDA (-2)
#define AZ_ERR_NOMEM (-3)

#define AZ_GAME_CONNECT4 0
#define AZ_GAME_OTHELLO 1

#define AZ_EVAL_UNIFORM 0 /* IEvaluator default: uniform policy + uniform WDL (src/cpp/IEvaluator.h:56-64) */
#define AZ_EVAL_ROLLOUT 1 /* RolloutEvaluator (src/cpp/RolloutEvaluator.h:16-49) */

/* SearchConfig - src/cpp/MCTSNode.h:47-61 (same fields, same defaults; bool widened to int32). */
typedef struct az_search_config {
    float c_init;               /* 1.25   */
    float c_base;               /* 19652  */
    float dirichlet_alpha;      /* 0.3 (<=0 disables root noise) */
    float noise_epsilon;        /* 0.25   */
    float fpu_reduction;        /* 0.4    */
    float mlh_slope;            /* 0 (Connect4 moves-left utility) */
    float mlh_cap;              /* 0.2    */
    float score_utility_factor; /* 0 (Othello score utility) */
    float score_scale;          /* 8      */
    float value_decay;          /* 1      */
    int32_t use_symmetry;       /* 1      */
    int32_t vl_count;           /* 1      */
} az_search_config;

typedef struct az_mcts az_mcts; /* BatchedMCTS<Game> - src/cpp/BatchedMCTS.h:26-442 */

const char *az_version(void);
const char *az_global_last_error(void); /* error of the last failed az_*_create on this thread */

/* static traits - mcts_bindings.cpp:359-369 (action_size / board_size / board_shape class properties) */
int az_game_action_size(int game);
int az_game_board_size(int game);
int az_game_board_rows(int game);
int az_game_board_cols(int game);
int az_game_num_symmetries(int game);
void az_search_config_defaults(az_search_config *cfg);

/* BatchedMCTS(n_envs) - mcts_bindings.cpp:52, BatchedMCTS.h:52-58.  device = CUDA ordinal. */
az_mcts *az_mcts_create(int game, int n_envs, int device);
void az_mcts_destroy(az_mcts *h);
const char *az_mcts_last_error(const az_mcts *h);
int az_mcts_num_envs(const az_mcts *h);                          /* get_num_envs, mcts_bindings.cpp:68 */
int az_mcts_set_config(az_mcts *h, const az_search_config *cfg); /* .config setter, mcts_bindings.cpp:55-58 */
int az_mcts_get_config(const az_mcts *h, az_search_config *cfg);
int az_mcts_set_seed(az_mcts *h, int64_t seed);                  /* set_seed, BatchedMCTS.h:68-84 (seed<0: re-randomise) */
int az_mcts_reset_env(az_mcts *h, int env_idx);                  /* reset_env, BatchedMCTS.h:93-99 (out of range ignored) */
int az_mcts_prune_roots(az_mcts *h, const int32_t *actions);     /* prune_roots, BatchedMCTS.h:105-112; actions[n_envs] */

/* search_batch - BatchedMCTS.h:119-171 / mcts_bindings.cpp:89-134.  One simulation per tree.
 * in : boards int8[n,S], turns int32[n]
 * out: leaf boards int8[n,S], term_d/p1w/p2w f32[n], is_term u8[n], leaf turns i32[n], valid_mask u8[n,A] */
int az_mcts_search_batch(az_mcts *h, const int8_t *boards, const int32_t *turns, int8_t *out_boards, float *out_term_d,
                         float *out_term_p1w, float *out_term_p2w, uint8_t *out_is_term, int32_t *out_turns,
                         uint8_t *out_valid_mask);
/* backprop_batch - BatchedMCTS.h:176-199 / mcts_bindings.cpp:139-179 */
int az_mcts_backprop_batch(az_mcts *h, const float *policy, const float *d_vals, const float *p1w_vals,
                           const float *p2w_vals, const float *moves_left, const uint8_t *is_term);
/* remove_all_vl - BatchedMCTS.h:209-216 (idempotent clean-up) */
int az_mcts_remove_all_vl(az_mcts *h, int K);
/* Zero-copy form of search_batch (K == 0) / search_batch_vl (K >= 1): the leaf arrays are delivered in a pinned host block that
 * the CALLER owns until az_pinned_release(out->block) - the single device-to-host copy of the call lands there, so a binding can
 * hand out numpy arrays over it without a second pass (mcts_bindings.cpp:89-134,197-252 return fresh arrays owned by Python: the
 * binding releases the block when the last of them is collected).  Blocks come from a process-wide pool and outlive the engine. */
typedef struct az_host_leaves {
    int32_t block, rows;             /* pool block id; rows = n_envs * max(K, 1) */
    int8_t *boards;                  /* [rows, R, C] */
    float *term_d, *term_p1w, *term_p2w;
    uint8_t *is_term;
    int32_t *turns, *sym_ids;        /* sym_ids: the ids applied (all 0 when K == 0: the non-VL search keeps them internally) */
    uint8_t *valid_mask;             /* [rows, A] */
} az_host_leaves;
int az_mcts_search_batch_pinned(az_mcts *h, int K, const int8_t *boards, const int32_t *turns, az_host_leaves *out);
int az_pinned_release(int block);
/* search_batch_vl - BatchedMCTS.h:227-286 / mcts_bindings.cpp:197-252.  K virtual-loss simulations per tree,
 * sequential inside a tree; outputs have n*K rows plus sym_ids i32[n*K]. */
int az_mcts_search_batch_vl(az_mcts *h, int K, const int8_t *boards, const int32_t *turns, int8_t *out_boards,
                            float *out_term_d, float *out_term_p1w, float *out_term_p2w, uint8_t *out_is_term,
                            int32_t *out_turns, int32_t *out_sym_ids, uint8_t *out_valid_mask);
/* backprop_batch_vl - BatchedMCTS.h:296-332 / mcts_bindings.cpp:257-306 */
int az_mcts_backprop_batch_vl(az_mcts *h, int K, const float *policy, const float *d_vals, const float *p1w_vals,
                              const float *p2w_vals, const float *moves_left, const uint8_t *is_term,
                              const int32_t *sym_ids);
/* search(evaluator, boards, turns, n_playout) - BatchedMCTS.h:339-407 / mcts_bindings.cpp:313-337: the whole
 * playout loop with a built-in evaluator (AZ_EVAL_*), entirely on the device. */
int az_mcts_search(az_mcts *h, int evaluator, const int8_t *boards, const int32_t *turns, int n_playout);
/* get_all_counts - BatchedMCTS.h:413-427: int32[n*A] */
int az_mcts_get_counts(az_mcts *h, int32_t *out);
/* the same counts widened to int64[n*A] - what src/MCTS_cpp.py:81,442-443 builds with np.array(get_all_counts()) */
int az_mcts_get_counts64(az_mcts *h, int64_t *out);
/* get_all_root_stats - BatchedMCTS.h:435-441, layout MCTS.h:634-636: f32[n, 6+8A] */
int az_mcts_get_root_stats(az_mcts *h, float *out);

/* ---- device-pointer twins (no copies, no synchronisation; stream = cudaStream_t) ----
 * On the device the engine speaks bitboards, not byte boards: roots go in as az_root, leaves come out as az_leaf
 * (both 32 bytes, one sector).  az_pack_roots_dev / az_unpack_leaves_dev convert from / to the arrays of the
 * reference API with fully coalesced kernels; the host entry points above are exactly pack -> search -> unpack. */
typedef struct az_root {     /* a position to search from (import_board + set_turn, BatchedMCTS.h:133-137) */
    uint64_t bb0, bb1;       /* stones of player +1 / -1.  Connect4: bit = col*7 + (5-row); Othello: bit = row*8+col */
    int32_t turn;            /* side to move, +1 / -1 */
    int32_t passes;          /* Othello consecutive passes   } used by the env kernels only: the search re-derives them */
    int32_t last;            /* index of the last mover, -1  } like import_board does (Connect4.h:124-128, Othello.h:108-110) */
    int32_t reserved;
} az_root;
#define AZ_LEAF_TERMINAL 1u
#define AZ_LEAF_P1_WINS 2u
#define AZ_LEAF_P2_WINS 4u
typedef struct az_leaf {     /* a leaf to evaluate (SimResult + sym id, MCTS.h:22-28, BatchedMCTS.h:141-158) */
    uint64_t bb0, bb1;       /* leaf position AFTER the random symmetry (what the network sees) */
    int8_t turn;             /* side to move at the leaf */
    uint8_t flags;           /* AZ_LEAF_* : terminal flag and cached result (neither win bit = draw) */
    uint8_t sym;             /* symmetry id that was applied (0 for terminal leaves) */
    uint8_t passes;          /* Othello consecutive passes (needed to rebuild the legal mask) */
    int32_t reserved[3];
} az_leaf;

int az_pack_roots_dev(int game, int n, const int8_t *d_boards, const int32_t *d_turns, az_root *d_roots, void *stream);
/* Any output pointer may be NULL.  d_planes = CNN input f32[rows,3,R,C] (plane0 = side to move, plane1 = opponent,
 * plane2 = turn; src/MCTS_cpp.py:15-20) so the network consumes leaves with no host round trip. */
int az_unpack_leaves_dev(int game, int rows, const az_leaf *d_leaves, int8_t *d_boards, float *d_term_d, float *d_term_p1w,
                         float *d_term_p2w, uint8_t *d_is_term, int32_t *d_turns, int32_t *d_sym_ids, uint8_t *d_valid_mask,
                         float *d_planes, void *stream);

int az_mcts_prune_roots_dev(az_mcts *h, const int32_t *d_actions, void *stream);
/* prune_roots with every action < 0 (each tree back to a fresh root, MCTS.h:101-109), stream-ordered.  Unlike the call above -
 * which cannot see the device array - the host-side arena bookkeeping is reset too, so the next search starts from empty arenas
 * with no device synchronisation in between. */
int az_mcts_reset_all_dev(az_mcts *h, void *stream);
/* K == 0 selects the non-VL search_batch (1 leaf per tree, d_leaves[n]); K >= 1 the virtual-loss search (d_leaves[n*K]). */
int az_mcts_search_dev(az_mcts *h, int K, const az_root *d_roots, az_leaf *d_leaves, void *stream);
/* d_is_term / d_sym_ids may be NULL: the engine then uses the flags / ids it remembered from the matching search. */
int az_mcts_backprop_dev(az_mcts *h, int K, const float *d_policy, const float *d_d, const float *d_p1w,
                         const float *d_p2w, const float *d_moves_left, const uint8_t *d_is_term,
                         const int32_t *d_sym_ids, void *stream);
/* Shard variants: the same launches restricted to trees [first, first + count) (first a multiple of 32).  d_roots is
 * indexed by tree as always; the rows of the leaf / policy / value arrays that belong to the shard start at row0: row of
 * tree i, simulation k = row0 + (i - first)*K + k (row0 = first*K reproduces the whole-batch layout; a loop whose K varies
 * between iterations should give every shard a fixed region, row0 = first*Kmax, because shards run ahead of each other).
 * Independent shards may be driven on different streams so that the select of one overlaps the evaluation / back-prop of
 * another (the tree kernels are latency bound and leave issue slots idle; a CNN evaluator is compute bound).  Pass
 * new_epoch = 1 for the first shard of a search iteration and 0 for the others: every shard then draws its leaf
 * symmetries from the same epoch as an unsharded az_mcts_search_dev call, so results do not depend on the sharding.
 * az_mcts_stream_handover_dev orders the stream after everything queued through this handle so far (call it on the
 * stream the shards fork from, and again on the stream that joined them). */
int az_mcts_search_range_dev(az_mcts *h, int K, const az_root *d_roots, az_leaf *d_leaves, int first, int count, int64_t row0,
                             int new_epoch, void *stream);
int az_mcts_backprop_range_dev(az_mcts *h, int K, const float *d_policy, const float *d_d, const float *d_p1w,
                               const float *d_p2w, const float *d_moves_left, const uint8_t *d_is_term,
                               const int32_t *d_sym_ids, int first, int count, int64_t row0, void *stream);
int az_mcts_stream_handover_dev(az_mcts *h, void *stream);
/* The per-move playout loop of the reference wrapper (src/MCTS_cpp.py:217-357: one non-VL warm-up simulation, then
 * ceil((n-1)/K) virtual-loss iterations) with one of the synthetic evaluators of az_eval_synthetic_dev, driven natively:
 * 3 launches per iteration and shard, no per-launch host-language cost.  shards > 1 pipelines that many tree ranges on
 * internal streams (forked from and joined back into `stream`); a loop issued before with the same buffers / configuration is
 * replayed from a CUDA graph (AZB200_GRAPHS=0 disables).  With shards, shard j's rows live at [first_j*Kmax, ...).  Buffers: az_leaf[n*max(K,1)], policy f32[n*max(K,1)*A],
 * d/p1w/p2w/moves_left f32[n*max(K,1)].  *launches_out (optional) = kernels launched. */
int az_mcts_playout_synthetic_dev(az_mcts *h, int mode, int n_playout, int K, int shards, const az_root *d_roots,
                                  az_leaf *d_leaves, float *d_policy, float *d_d, float *d_p1w, float *d_p2w,
                                  float *d_moves_left, void *stream, int *launches_out);
/* The same loop from HOST arrays (boards i8[n, R, C], turns i32[n], as search_batch takes them: mcts_bindings.cpp:89-134), pipelined
 * shard by shard: staging, host-to-device copy, pack kernel, the shard's loop (its own CUDA graph), visit-count kernel and
 * device-to-host copy run on the shard's own stream, so the GPU searches shard 0 while the host still stages shard 1 and finished
 * shards' counts travel while the others search.  Synchronous (returns when every shard is done, like the reference's calls).
 * want_counts != 0: the int64 visit counts are left in a pinned block that the next az_mcts_get_counts64_pinned hands out
 * (src/player.py:333-343 calls get_visits_count right after batch_playout).  Uses the engine's own device buffers. */
int az_mcts_playout_synthetic_host(az_mcts *h, int mode, int n_playout, int K, int shards, const int8_t *boards,
                                   const int32_t *turns, int want_counts, int *launches_out);
/* get_all_counts as int64[n*A] (np.array(get_all_counts()), src/MCTS_cpp.py:81,442-443) in a pinned pool block the CALLER owns until
 * az_pinned_release(*block_out): the counts are widened on the device and copied once, straight into the memory the caller's
 * array lives in. */
int az_mcts_get_counts64_pinned(az_mcts *h, int64_t **out, int *block_out);
int az_mcts_search_eval_dev(az_mcts *h, int evaluator, const az_root *d_roots, int n_playout, void *stream);
int az_mcts_get_counts_dev(az_mcts *h, int32_t *d_out, void *stream);
int az_mcts_get_root_stats_dev(az_mcts *h, float *d_out, void *stream);

/* Lanes cooperating on one tree: Connect4 1/2/4/8 (0 = automatic: 1), Othello 8/16 (0 = choose from n_envs).  More trees per warp means
 * fewer replicated instructions; fewer means more warps to hide latency when n_envs is small. */
int az_mcts_set_lanes(az_mcts *h, int lanes);
int az_mcts_get_lanes(const az_mcts *h);
/* Generation of the thread-per-tree Connect4 kernels (lanes == 1): 0 = first generation, 1 = lean kernels (branch-free
 * IEEE divisions, software-pipelined block gather, staged back-prop, 256-bit slot accesses; default).  Both compute
 * bit-identical results; the setting exists for A/B measurements and the parity tests.  Env: AZB200_VARIANT. */
int az_mcts_set_variant(az_mcts *h, int variant);
int az_mcts_get_variant(const az_mcts *h);
/* Staggered-descent select (Connect4, lanes == 1, lean kernels, 1 <= K <= 8): small batches give every virtual-loss descent
 * its own lane, descent k starting one tree level after descent k-1 (K + depth - 1 dependent level steps instead of K x depth;
 * simulate_vl, MCTS.h:443-545, same results bit for bit).  Used when n_envs x group width (4 lanes per tree for K <= 4, 8 for
 * K <= 8) <= max_lanes; default 131072 (measured crossover: self-play with tree reuse at 32 768 trees), 0 = off.  Env: AZB200_WAVE_MAX.
 * Othello (2 <= K <= 4) has the same scheme on its lane-group kernels - a warp per tree, the K descents in 8-lane groups one
 * level apart - when n_envs x 16 <= max_lanes (8192 trees by default). */
int az_mcts_set_wave_max(az_mcts *h, int max_lanes);
int az_mcts_get_wave_max(const az_mcts *h);
/* Self-test of the branch-free division sequences against the compiler's IEEE division: mode 0 = 1/n for n = 1..count,
 * mode 1 = random a/b over the covered range, mode 2 = small-integer ratios (exact results and ties).  Writes the number
 * of bit mismatches (expected 0). */
int az_selftest_div(int mode, uint64_t count, uint64_t seed, uint64_t *mismatches);
/* Global index of env 0 of this handle: all RNG streams are keyed by (seed, global env index), so a game's result does
 * not depend on how games are sharded over GPUs. */
int az_mcts_set_env_base(az_mcts *h, uint64_t base);
/* Arena compaction.  The reference's node pools only grow within a game (MCTSNode.h:149-199); after a re-root only the
 * subtree below the played move is reachable.  mode 1 (default): when the growth of one more move might no longer fit at a
 * re-root, the live tree of every env is copied breadth-first into a second pool of the same size and the pools are swapped
 * (bounded memory per game: many more concurrent self-play games fit in HBM).  0 = never, 2 = at every re-root.  Search
 * results do not depend on it; leaf records of a search whose back-prop has not run yet are invalidated by a compaction.
 * Env: AZB200_COMPACTION.  az_mcts_compactions = compactions performed so far. */
int az_mcts_set_compaction(az_mcts *h, int mode);
uint64_t az_mcts_compactions(const az_mcts *h);
/* Lazy edge blocks (Connect4, lean thread-per-tree kernels, read-only selects): an expansion below the root stores a 32-byte header
 * {priors, legal mask} instead of num_edges x 32 bytes; the block is materialised on the node's second visit.  Fewer DRAM bytes
 * (most expanded nodes are never visited again), but one more dependent load per back-propagated path: measured slower on B200
 * (DESIGN.md), hence OFF by default.  Results are identical either way.  Env: AZB200_LAZY=1. */
int az_mcts_set_lazy(az_mcts *h, int on);
int az_mcts_get_lazy(const az_mcts *h);
/* Pre-size every tree arena (slots of 32 bytes per tree) so no reallocation happens later (e.g. under graph capture). */
int az_mcts_reserve(az_mcts *h, int slots_per_tree);

/* engine counters for the roofline model: out[0..7] = simulations, edges traversed (sum of depths), edges
 * scanned, edges created, expansions, max arena use (slots), arena capacity (slots/tree), kernel launches */
int az_mcts_enable_stats(az_mcts *h, int on);
int az_mcts_get_stats(az_mcts *h, uint64_t *out8);
/* Measurement: with timing on, every select launch is bracketed by CUDA events on the stream it is launched on;
 * az_mcts_get_select_time synchronises the device and returns the summed elapsed time (ms), the number of launches and
 * the leaf rows (trees x simulations) they produced since the last call (bench.py: roofline of the dominant kernel). */
int az_mcts_time_select(az_mcts *h, int on);
int az_mcts_get_select_time(az_mcts *h, float *ms_out, int *launches_out, uint64_t *rows_out);
/* the same for the back-prop launches (timed whenever az_mcts_time_select is on) */
int az_mcts_get_backprop_time(az_mcts *h, float *ms_out, int *launches_out, uint64_t *rows_out);
/* Diagnostics (stats mode, thread-per-tree Connect4 select): out[2w], out[2w+1] = %globaltimer (ns) at the start / end of
 * warp w of the most recent select launch.  Returns the number of warps written (<= max_warps) or a negative error. */
int az_mcts_get_warp_times(az_mcts *h, uint64_t *out, int max_warps);

/* Synthetic deterministic evaluators on device pointers (twins of alphazero-al_b200/evaluators.py): turn the leaves
 * of az_mcts_search_dev into the backprop tuple.  mode: 0 hash, 1 flip-equivariant hash (Connect4), 2 constant. */
int az_eval_synthetic_dev(int game, int mode, int n_leaves, const az_leaf *d_leaves, float *d_policy, float *d_d,
                          float *d_p1w, float *d_p2w, float *d_moves_left, void *stream);

/* Network outputs -> backprop tuple on device: d_wdl_rel f32[n,3] = [draw, win(to move), loss(to move)] (CNN.predict,
 * src/environments/Connect4/Network.py:267-288) becomes absolute d/p1w/p2w by the leaf's side to move
 * (src/MCTS_cpp.py:23-30); terminal leaves get their cached result and moves_left 0 (src/MCTS_cpp.py:276-282). */
int az_eval_finalize_dev(int n_leaves, const az_leaf *d_leaves, const float *d_wdl_rel, const float *d_aux, float *d_d,
                         float *d_p1w, float *d_p2w, float *d_moves_left, void *stream);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* AZB200_H */
