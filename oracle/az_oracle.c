/*
 * az_oracle.c - CPU restatement of the AlphaZero-AL self-play hot path (batched PUCT MCTS + bitboard envs).
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is the parity checker for the CUDA engine in
 * alphazero-al_b200/csrc/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load it; the product path never does (and fails loudly when its CUDA library is missing).
 *
 * It is a from-scratch, single-threaded, plain-C restatement of the reference algorithm; every
 * function cites the reference file:line (relative to /root/reference/) whose behaviour it follows.
 * Pinning: the reference ships no golden vectors, so this restatement is pinned
 * against the UNMODIFIED reference engine compiled into oracle/_ref/ (oracle/Makefile `ref`) by
 * tests/test_oracle_vs_ref.py, and against fixtures generated from that build (tests/golden/,
 * generator tests/golden/make_golden.py).
 *
 * Parity status: PINNED for everything deterministic (env stepping, legal masks, winners, leaf
 * boards, visit counts, root stats - bit exact with oracle/_ref/parity built -ffp-contract=off).
 * PARITY UNPINNED for the RNG-dependent features (Dirichlet root noise, random leaf symmetry ids,
 * random rollouts): the reference draws them from a thread_local std::mt19937 owned by whichever
 * OpenMP thread runs env i (src/cpp/MCTS.h:13-17, BatchedMCTS.h:68-84); here they come from a
 * counter-based splitmix64 stream keyed by (seed, epoch, env, k) that the CUDA engine reproduces
 * bit for bit, so oracle-vs-CUDA is exact while oracle-vs-reference is distributional only.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off  (no FMA contraction - SURVEY.md App. C.5).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_C4 0
#define ORC_OTH 1
#define ORC_MAX_A 65
#define ORC_EVAL_UNIFORM 0
#define ORC_EVAL_ROLLOUT 1

/* ------------------------------------------------------------------------------------------------
 * SearchConfig  (src/cpp/MCTSNode.h:47-61)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float c_init, c_base, dirichlet_alpha, noise_epsilon, fpu_reduction;
    float mlh_slope, mlh_cap, score_utility_factor, score_scale, value_decay;
    int32_t use_symmetry;
    int32_t vl_count;
} orc_config;

static void config_defaults(orc_config *c) {
    c->c_init = 1.25f; c->c_base = 19652.0f; c->dirichlet_alpha = 0.3f; c->noise_epsilon = 0.25f;
    c->fpu_reduction = 0.4f; c->mlh_slope = 0.0f; c->mlh_cap = 0.2f; c->score_utility_factor = 0.0f;
    c->score_scale = 8.0f; c->value_decay = 1.0f; c->use_symmetry = 1; c->vl_count = 1;
}

/* ------------------------------------------------------------------------------------------------
 * Counter-based RNG shared bit-for-bit with the CUDA engine (csrc/az_rng.cuh).
 * ---------------------------------------------------------------------------------------------- */
static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
/* stream: 0 = leaf symmetry ids, 1 = rollouts, 2 = dirichlet */
static uint64_t orc_rand(uint64_t seed, uint64_t epoch, uint64_t stream, uint64_t env, uint64_t ctr) {
    uint64_t h = splitmix64(seed ^ (stream * 0xD6E8FEB86659FD93ULL));
    h = splitmix64(h ^ epoch);
    h = splitmix64(h ^ (env << 24) ^ ctr);
    return h;
}

/* ------------------------------------------------------------------------------------------------
 * Game state.  One struct for both games (C has no templates); fields unused by a game stay zero.
 *   Connect4: src/cpp/Connect4.h:31-295     Othello: src/cpp/Othello.h:28-388
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int8_t board[64];   /* C4 uses the first 42 (row-major 6x7), Othello all 64 (8x8) */
    int turn;
    uint64_t bb[2];
    int height[7];      /* C4 only */
    int n_pieces;
    int passes;         /* Othello consecutive_passes */
    int last;           /* last_player_idx (-1 none) */
} orc_env;

typedef struct { int moves[ORC_MAX_A]; int count; } orc_moves;

static int game_A(int g) { return g == ORC_C4 ? 7 : 65; }
static int game_S(int g) { return g == ORC_C4 ? 42 : 64; }

/* ---- Connect4 ---- */
static void c4_reset(orc_env *e) {                       /* Connect4.h:62-72 */
    memset(e, 0, sizeof(*e));
    e->turn = 1; e->last = -1;
    for (int c = 0; c < 7; ++c) e->height[c] = c * 7;
}
static void c4_sync_from_board(orc_env *e) {              /* Connect4.h:100-129 */
    e->bb[0] = e->bb[1] = 0; e->n_pieces = 0; e->last = -1;
    for (int c = 0; c < 7; ++c) {
        e->height[c] = c * 7;
        for (int r = 5; r >= 0; --r) {
            int8_t v = e->board[r * 7 + c];
            if (v == 0) break;
            e->bb[v == 1 ? 0 : 1] |= 1ULL << e->height[c];
            e->height[c]++; e->n_pieces++;
        }
    }
    if (e->n_pieces > 0) e->last = (e->n_pieces % 2 == 1) ? 0 : 1;
}
static void c4_sync_to_board(orc_env *e) {                /* Connect4.h:135-150 */
    memset(e->board, 0, 42);
    for (int c = 0; c < 7; ++c)
        for (int bit = c * 7; bit < e->height[c]; ++bit)
            e->board[(5 - (bit - c * 7)) * 7 + c] = (e->bb[0] >> bit & 1) ? 1 : -1;
}
static void c4_step(orc_env *e, int col) {                /* Connect4.h:159-172 (no legality check) */
    int p = e->turn == 1 ? 0 : 1;
    e->bb[p] |= 1ULL << e->height[col];
    int row = 5 - (e->height[col] - col * 7);
    if (row >= 0 && row < 6) e->board[row * 7 + col] = (int8_t)e->turn;
    e->height[col]++; e->n_pieces++; e->last = p; e->turn = -e->turn;
}
static int c4_winner(const orc_env *e) {                  /* Connect4.h:182-203 */
    if (e->last == -1) return 0;
    uint64_t b = e->bb[e->last], t;
    int res = e->last == 0 ? 1 : -1;
    t = b & (b >> 1); if (t & (t >> 2))  return res;
    t = b & (b >> 7); if (t & (t >> 14)) return res;
    t = b & (b >> 6); if (t & (t >> 12)) return res;
    t = b & (b >> 8); if (t & (t >> 16)) return res;
    return 0;
}
static void c4_valid(const orc_env *e, orc_moves *m) {     /* Connect4.h:209-218 */
    m->count = 0;
    for (int c = 0; c < 7; ++c) if (e->height[c] < c * 7 + 6) m->moves[m->count++] = c;
}
static void c4_symmetry(orc_env *e, int sym) {             /* Connect4.h:249-280 */
    if (sym == 0) return;
    for (int p = 0; p < 2; ++p) {
        uint64_t s = e->bb[p], d = 0;
        for (int c = 0; c < 7; ++c) {
            uint64_t colbits = (s >> (c * 7)) & 0x7FULL;
            d |= colbits << ((6 - c) * 7);
        }
        e->bb[p] = d;
    }
    for (int c = 0; c < 3; ++c) {
        int m = 6 - c, hc = e->height[c] - c * 7, hm = e->height[m] - m * 7;
        e->height[c] = c * 7 + hm; e->height[m] = m * 7 + hc;
    }
    c4_sync_to_board(e);
}

/* ---- Othello ---- */
#define NOT_A 0xFEFEFEFEFEFEFEFEULL
#define NOT_H 0x7F7F7F7F7F7F7F7FULL
static uint64_t oth_shift(uint64_t b, int d) {             /* Othello.h:133-147 */
    switch (d) {
    case 0: return b >> 8;
    case 1: return (b >> 7) & NOT_A;
    case 2: return (b << 1) & NOT_A;
    case 3: return (b << 9) & NOT_A;
    case 4: return b << 8;
    case 5: return (b << 7) & NOT_H;
    case 6: return (b >> 1) & NOT_H;
    default: return (b >> 9) & NOT_H;
    }
}
static void oth_reset(orc_env *e) {                        /* Othello.h:62-75 */
    memset(e, 0, sizeof(*e));
    e->turn = 1;
    e->board[27] = -1; e->board[28] = 1; e->board[35] = 1; e->board[36] = -1;
    e->bb[0] = (1ULL << 28) | (1ULL << 35);
    e->bb[1] = (1ULL << 27) | (1ULL << 36);
    e->n_pieces = 4; e->last = -1;
}
static void oth_sync_from_board(orc_env *e) {              /* Othello.h:92-111 */
    e->bb[0] = e->bb[1] = 0;
    for (int i = 0; i < 64; ++i) {
        if (e->board[i] == 1) e->bb[0] |= 1ULL << i;
        else if (e->board[i] == -1) e->bb[1] |= 1ULL << i;
    }
    e->n_pieces = __builtin_popcountll(e->bb[0]) + __builtin_popcountll(e->bb[1]);
    e->passes = 0; e->last = -1;
}
static void oth_sync_to_board(orc_env *e) {                /* Othello.h:114-124 */
    memset(e->board, 0, 64);
    for (int i = 0; i < 64; ++i) {
        if (e->bb[0] >> i & 1) e->board[i] = 1;
        else if (e->bb[1] >> i & 1) e->board[i] = -1;
    }
}
static uint64_t oth_valid_positions(const orc_env *e) {    /* Othello.h:155-171 */
    int p = e->turn == 1 ? 0 : 1;
    uint64_t own = e->bb[p], opp = e->bb[1 - p], empty = ~(own | opp), valid = 0;
    for (int d = 0; d < 8; ++d) {
        uint64_t c = oth_shift(own, d) & opp;
        for (int i = 0; i < 5; ++i) c |= oth_shift(c, d) & opp;
        valid |= oth_shift(c, d) & empty;
    }
    return valid;
}
static uint64_t oth_flips(const orc_env *e, int pos) {     /* Othello.h:177-198 */
    int p = e->turn == 1 ? 0 : 1;
    uint64_t own = e->bb[p], opp = e->bb[1 - p], flipped = 0;
    for (int d = 0; d < 8; ++d) {
        uint64_t cand = 0, sq = oth_shift(1ULL << pos, d);
        while (sq & opp) { cand |= sq; sq = oth_shift(sq, d); }
        if (sq & own) flipped |= cand;
    }
    return flipped;
}
static void oth_step(orc_env *e, int action) {             /* Othello.h:206-235 */
    if (action == 64) { e->passes++; e->turn = -e->turn; return; }
    int p = e->turn == 1 ? 0 : 1;
    uint64_t flips = oth_flips(e, action);
    e->bb[p] |= (1ULL << action) | flips;
    e->bb[1 - p] &= ~flips;
    e->board[action] = (int8_t)e->turn;
    for (uint64_t f = flips; f; f &= f - 1) e->board[__builtin_ctzll(f)] = (int8_t)e->turn;
    e->n_pieces++; e->passes = 0; e->last = p; e->turn = -e->turn;
}
static int oth_over(const orc_env *e) { return e->n_pieces == 64 || e->passes >= 2; }   /* Othello.h:241-244 */
static int oth_winner(const orc_env *e) {                  /* Othello.h:250-258 */
    if (!oth_over(e)) return 0;
    int a = __builtin_popcountll(e->bb[0]), b = __builtin_popcountll(e->bb[1]);
    return a > b ? 1 : (b > a ? -1 : 0);
}
static void oth_valid(const orc_env *e, orc_moves *m) {    /* Othello.h:282-296 */
    m->count = 0;
    if (oth_over(e)) return;
    uint64_t v = oth_valid_positions(e);
    if (!v) { m->moves[m->count++] = 64; return; }
    for (; v; v &= v - 1) m->moves[m->count++] = __builtin_ctzll(v);
}
static void oth_xform(int sym, int r, int c, int *nr, int *nc) {   /* Othello.h:312-326 */
    switch (sym) {
    case 1: *nr = c;     *nc = 7 - r; break;
    case 2: *nr = 7 - r; *nc = 7 - c; break;
    case 3: *nr = 7 - c; *nc = r;     break;
    case 4: *nr = r;     *nc = 7 - c; break;
    case 5: *nr = 7 - r; *nc = c;     break;
    case 6: *nr = c;     *nc = r;     break;
    case 7: *nr = 7 - c; *nc = 7 - r; break;
    default: *nr = r;    *nc = c;     break;
    }
}
static uint64_t oth_xform_bb(uint64_t b, int sym) {        /* Othello.h:329-341 */
    uint64_t r = 0;
    for (; b; b &= b - 1) {
        int i = __builtin_ctzll(b), nr, nc;
        oth_xform(sym, i / 8, i % 8, &nr, &nc);
        r |= 1ULL << (nr * 8 + nc);
    }
    return r;
}
static void oth_symmetry(orc_env *e, int sym) {            /* Othello.h:347-353 */
    if (sym == 0) return;
    e->bb[0] = oth_xform_bb(e->bb[0], sym);
    e->bb[1] = oth_xform_bb(e->bb[1], sym);
    oth_sync_to_board(e);
}
static const int OTH_INV_SYM[8] = {0, 3, 2, 1, 4, 5, 6, 7};        /* Othello.h:356-361 */
static const int OTH_MCTS_SYMS[4] = {0, 2, 6, 7};                 /* Othello.h:45 */

/* ---- game dispatch (the MCTSGame concept, src/cpp/GameContext.h:44-72) ---- */
static void env_reset(int g, orc_env *e) { if (g == ORC_C4) c4_reset(e); else oth_reset(e); }
static void env_import(int g, orc_env *e, const int8_t *src) {     /* import_board: Connect4.h:87-91, Othello.h:83-87 */
    memset(e->board, 0, 64);
    memcpy(e->board, src, (size_t)game_S(g));
    if (g == ORC_C4) c4_sync_from_board(e); else oth_sync_from_board(e);
}
static void env_step(int g, orc_env *e, int a) { if (g == ORC_C4) c4_step(e, a); else oth_step(e, a); }
static int env_winner(int g, const orc_env *e) { return g == ORC_C4 ? c4_winner(e) : oth_winner(e); }
static int env_full(int g, const orc_env *e) { return g == ORC_C4 ? e->n_pieces == 42 : oth_over(e); }
static void env_valid(int g, const orc_env *e, orc_moves *m) { if (g == ORC_C4) c4_valid(e, m); else oth_valid(e, m); }
static void env_symmetry(int g, orc_env *e, int s) { if (g == ORC_C4) c4_symmetry(e, s); else oth_symmetry(e, s); }
static void env_sync_to_board(int g, orc_env *e) { if (g == ORC_C4) c4_sync_to_board(e); else oth_sync_to_board(e); }

/* inverse_symmetry_policy: Connect4.h:288-294, Othello.h:373-387 */
static void inverse_symmetry_policy(int g, int sym, float *policy) {
    if (sym == 0) return;
    if (g == ORC_C4) {
        for (int i = 0; i < 3; ++i) { float t = policy[i]; policy[i] = policy[6 - i]; policy[6 - i] = t; }
    } else {
        float tmp[65];
        int inv = OTH_INV_SYM[sym & 7];
        for (int i = 0; i < 64; ++i) {
            int nr, nc;
            oth_xform(inv, i / 8, i % 8, &nr, &nc);
            tmp[nr * 8 + nc] = policy[i];
        }
        tmp[64] = policy[64];
        memcpy(policy, tmp, sizeof(tmp));
    }
}
/* terminal_aux: Connect4.h:226-229 (0), Othello.h:260-266 (atan of signed disc difference) */
static float terminal_aux(int g, const orc_env *e, const orc_config *cfg) {
    if (g == ORC_C4) return 0.0f;
    int diff = __builtin_popcountll(e->bb[0]) - __builtin_popcountll(e->bb[1]);
    float raw = (float)(diff * e->turn);
    return atanf(raw / cfg->score_scale) * (2.0f / 3.14159265f);
}
static float clampf(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); }   /* std::clamp */
/* compute_aux_utility: Connect4.h:231-239, Othello.h:268-274 */
static float aux_utility(int g, float child_M, float parent_M, float child_Q, const orc_config *cfg) {
    if (g == ORC_C4) {
        if (cfg->mlh_slope <= 0.0f) return 0.0f;
        float m_diff = child_M - parent_M;
        float u = clampf(cfg->mlh_slope * m_diff, -cfg->mlh_cap, cfg->mlh_cap);
        return u * child_Q;
    }
    if (cfg->score_utility_factor <= 0.0f) return 0.0f;
    return cfg->score_utility_factor * child_M;
}

/* ------------------------------------------------------------------------------------------------
 * Tree storage  (src/cpp/MCTSNode.h:69-199)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float d, p1w, p2w; } wdl_t;
typedef struct { int32_t action, child; float prior, noise; } edge_t;
typedef struct {
    float W_d, W_p1w, W_p2w;
    int32_t n_visits, n_inflight;
    float M_sum;
    int32_t num_edges, edge_offset, parent, parent_edge_idx;
    int8_t turn; uint8_t is_expanded, is_terminal;
    float term_d, term_p1w, term_p2w;
} node_t;

typedef struct { int32_t node, edge; } vlent_t;

typedef struct {
    node_t *nodes; int ncap, ncount;
    edge_t *edges; int ecap, ecount;
    int root;
    orc_env sim_env;
    int cur_leaf, cur_leaf_turn;
    /* virtual-loss state (src/cpp/MCTS.h:60-64) */
    int vlK;                 /* vl_paths_.size() */
    vlent_t **paths; int *plen, *pcap;
    orc_env *vl_envs; int *vl_leaf; int *vl_turn; int vl_alloc;
    /* statistics for the roofline model (not in the reference) */
    uint64_t stat_depth, stat_edges, stat_expanded, stat_sims, stat_alloc, stat_seen, stat_expansions;
    uint64_t stat_lvl_nodes[8], stat_lvl_edges[8], stat_lvl_alloc[8];   /* per level of the descent (7 = deeper) */
    int stat_level;
} tree_t;

typedef struct {
    int game, n_envs;
    orc_config cfg;
    tree_t *trees;
    int *pending_sym;        /* BatchedMCTS.h:45 */
    uint64_t seed, epoch, noise_ctr;
} engine_t;

static wdl_t winner_to_wdl(int w) {                        /* MCTSNode.h:35-39 */
    wdl_t r = {0, 0, 0};
    if (w == 1) r.p1w = 1.0f; else if (w == -1) r.p2w = 1.0f; else r.d = 1.0f;
    return r;
}
static wdl_t mean_wdl(const node_t *n) {                   /* MCTSNode.h:118-122 */
    wdl_t r;
    if (n->n_visits == 0) { r.d = r.p1w = r.p2w = 1.f / 3; return r; }
    float inv = 1.0f / (float)n->n_visits;
    r.d = n->W_d * inv; r.p1w = n->W_p1w * inv; r.p2w = n->W_p2w * inv;
    return r;
}
static float wdl_q(wdl_t w, int turn) { return turn == 1 ? (w.p1w - w.p2w) : (w.p2w - w.p1w); }   /* MCTSNode.h:23-25 */
static float mean_q(const node_t *n) { return wdl_q(mean_wdl(n), n->turn); }                       /* MCTSNode.h:125 */
static float mean_M(const node_t *n) { return n->n_visits == 0 ? 0.0f : n->M_sum / (float)n->n_visits; }  /* :131-133 */

static int alloc_node(tree_t *t) {                         /* MCTSNode.h:161-168 */
    if (t->ncount >= t->ncap) { t->ncap *= 2; t->nodes = (node_t *)realloc(t->nodes, sizeof(node_t) * (size_t)t->ncap); }
    int i = t->ncount++;
    node_t *n = &t->nodes[i];
    memset(n, 0, sizeof(*n));
    n->edge_offset = -1; n->parent = -1; n->parent_edge_idx = -1; n->turn = 1;
    return i;
}
static int alloc_edges(tree_t *t, int count) {             /* MCTSNode.h:171-182 */
    int off = t->ecount;
    if (off + count > t->ecap) {
        while (off + count > t->ecap) t->ecap *= 2;
        t->edges = (edge_t *)realloc(t->edges, sizeof(edge_t) * (size_t)t->ecap);
    }
    for (int i = 0; i < count; ++i) { edge_t e = {-1, -1, 0.0f, 0.0f}; t->edges[off + i] = e; }
    t->ecount = off + count;
    return off;
}
static void tree_reset(tree_t *t) {                        /* MCTS.h:77-82 (fresh root always has turn=+1) */
    t->ncount = 0; t->ecount = 0;
    t->root = alloc_node(t);
    t->nodes[t->root].turn = 1;
}
static void tree_init(tree_t *t) {
    memset(t, 0, sizeof(*t));
    t->ncap = 2048; t->ecap = 8192;
    t->nodes = (node_t *)malloc(sizeof(node_t) * (size_t)t->ncap);
    t->edges = (edge_t *)malloc(sizeof(edge_t) * (size_t)t->ecap);
    t->cur_leaf = -1; t->cur_leaf_turn = 1;
    tree_reset(t);
}
static void tree_free(tree_t *t) {
    free(t->nodes); free(t->edges);
    for (int k = 0; k < t->vl_alloc; ++k) free(t->paths[k]);
    free(t->paths); free(t->plen); free(t->pcap); free(t->vl_envs); free(t->vl_leaf); free(t->vl_turn);
}

/* ---- Dirichlet noise: gamma(alpha,1) via Marsaglia-Tsang.  RNG-dependent => parity unpinned. ---- */
static double u01(engine_t *E, int env) {
    uint64_t h = orc_rand(E->seed, 0, 2, (uint64_t)env, E->noise_ctr++);
    return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
static double gauss(engine_t *E, int env) {
    double a = u01(E, env), b = u01(E, env);
    return sqrt(-2.0 * log(a)) * cos(6.283185307179586 * b);
}
static float gamma_draw(engine_t *E, int env, float alpha) {
    double a = alpha, boost = 1.0;
    if (a < 1.0) { boost = pow(u01(E, env), 1.0 / a); a += 1.0; }
    double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        double x = gauss(E, env), v = 1.0 + c * x;
        if (v <= 0) continue;
        v = v * v * v;
        double u = u01(E, env);
        if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return (float)(d * v * boost);
    }
}
static void draw_noise(engine_t *E, int env, int n, float *out) {   /* MCTS.h:119-131 / 350-363 */
    float sum = 0.0f;
    for (int i = 0; i < n; ++i) { out[i] = gamma_draw(E, env, E->cfg.dirichlet_alpha); sum += out[i]; }
    float inv = 1.0f / (sum + 1e-8f);
    for (int i = 0; i < n; ++i) out[i] = out[i] * inv;
}
static void apply_root_noise(engine_t *E, int env, tree_t *t) {     /* MCTS.h:113-132 */
    if (E->cfg.dirichlet_alpha <= 0.0f) return;
    node_t *root = &t->nodes[t->root];
    if (!root->is_expanded || root->num_edges == 0) return;
    float noise[ORC_MAX_A];
    draw_noise(E, env, root->num_edges, noise);
    for (int i = 0; i < root->num_edges; ++i) t->edges[root->edge_offset + i].noise = noise[i];
}
static void prune_root(engine_t *E, int env, tree_t *t, int action) {   /* MCTS.h:90-108 */
    node_t *root = &t->nodes[t->root];
    if (root->is_expanded) {
        for (int i = 0; i < root->num_edges; ++i) {
            edge_t *e = &t->edges[root->edge_offset + i];
            if (e->action == action && e->child != -1) {
                t->root = e->child;
                t->nodes[t->root].parent = -1;
                apply_root_noise(E, env, t);
                return;
            }
        }
    }
    tree_reset(t);
}

/* ---- selection (MCTS.h:140-234) ---- */
static float compute_fpu(const engine_t *E, const tree_t *t, int node_idx) {   /* MCTS.h:140-156 */
    const node_t *node = &t->nodes[node_idx];
    float parent_q = mean_q(node);
    float seen_policy = 0.0f;
    for (int i = 0; i < node->num_edges; ++i) {
        const edge_t *e = &t->edges[node->edge_offset + i];
        if (e->child != -1 && t->nodes[e->child].n_visits > 0) seen_policy += e->prior;
    }
    float scale = (1.0f + parent_q) / 2.0f;
    float effective_fpu = E->cfg.fpu_reduction * scale;
    float fpu_value = parent_q - effective_fpu * sqrtf(seen_policy);
    return (-1.0f < fpu_value) ? fpu_value : -1.0f;     /* std::max(-1.0f, fpu_value) */
}
static int select_edge(const engine_t *E, tree_t *t, int node_idx, float fpu_value) {   /* MCTS.h:163-234 */
    const orc_config *cfg = &E->cfg;
    const node_t *node = &t->nodes[node_idx];
    float parent_n = (float)(node->n_visits + node->n_inflight);
    float parent_M = mean_M(node);
    int is_root = node_idx == t->root;
    float ne = cfg->noise_epsilon;
    float best_score = -INFINITY;
    int best_edge = -1;
    for (int i = 0; i < node->num_edges; ++i) {
        const edge_t *e = &t->edges[node->edge_offset + i];
        float effective_prior = e->prior;
        if (is_root && ne > 0.0f) effective_prior = (1.0f - ne) * e->prior + ne * e->noise;
        float q_value, child_Q = 0.0f, child_M = 0.0f;
        int child_visits_total = 0;
        int seen = e->child != -1 && t->nodes[e->child].n_visits > 0;
        if (seen) {
            const node_t *child = &t->nodes[e->child];
            child_visits_total = child->n_visits + child->n_inflight;
            child_Q = mean_q(child);
            child_M = mean_M(child);
            if (E->game == ORC_OTH) child_M = -child_M;            /* AUX_NEGATE_PER_PLY, MCTS.h:197-198 */
            q_value = -child_Q;
        } else if (e->child != -1 && t->nodes[e->child].n_inflight > 0) {
            q_value = fpu_value;
            child_visits_total = t->nodes[e->child].n_inflight;
        } else {
            q_value = fpu_value;
        }
        float c_puct = cfg->c_init + logf((parent_n + cfg->c_base + 1.0f) / cfg->c_base);
        float u_score = c_puct * effective_prior * sqrtf(parent_n) / (1.0f + child_visits_total);
        float m_utility = 0.0f;
        if (seen) m_utility = aux_utility(E->game, child_M, parent_M, child_Q, cfg);
        float score = q_value + u_score + m_utility;
        if (score > best_score) { best_score = score; best_edge = i; }
        t->stat_edges++;
        t->stat_alloc += (uint64_t)(e->child != -1);
        t->stat_seen += (uint64_t)seen;
        t->stat_lvl_edges[t->stat_level]++;
        t->stat_lvl_alloc[t->stat_level] += (uint64_t)(e->child != -1);
    }
    return best_edge;
}

/* ---- simulate / simulate_vl (MCTS.h:242-322, 443-545).  vl<0 selects the non-VL variant. ---- */
typedef struct { orc_env board; wdl_t twdl; int is_terminal; } sim_result;

static void path_push(tree_t *t, int k, int node, int edge) {
    if (t->plen[k] >= t->pcap[k]) {
        t->pcap[k] = t->pcap[k] ? t->pcap[k] * 2 : 64;
        t->paths[k] = (vlent_t *)realloc(t->paths[k], sizeof(vlent_t) * (size_t)t->pcap[k]);
    }
    t->paths[k][t->plen[k]].node = node; t->paths[k][t->plen[k]].edge = edge; t->plen[k]++;
}
static sim_result simulate_common(engine_t *E, tree_t *t, const orc_env *start, int k /* -1 = non-VL */) {
    const int g = E->game;
    const int use_vl = k >= 0;
    if (use_vl) t->plen[k] = 0;
    t->sim_env = *start;
    int curr = t->root, winner = 0, full = 0, root_vl_applied = 0;
    t->stat_level = 0;
    while (t->nodes[curr].is_expanded) {
        node_t *node = &t->nodes[curr];
        if (node->is_terminal) break;
        if (node->num_edges == 0) break;
        float fpu = compute_fpu(E, t, curr);
        t->stat_lvl_nodes[t->stat_level]++;
        int best = select_edge(E, t, curr, fpu);
        if (t->stat_level < 7) t->stat_level++;
        if (best < 0) break;
        if (use_vl && !root_vl_applied) { t->nodes[t->root].n_inflight += E->cfg.vl_count; root_vl_applied = 1; }
        edge_t *e = &t->edges[node->edge_offset + best];
        env_step(g, &t->sim_env, e->action);
        if (e->child == -1) {                                       /* lazy child allocation */
            int c = alloc_node(t);
            node = &t->nodes[curr];                                 /* realloc may have moved nodes */
            e = &t->edges[node->edge_offset + best];
            e->child = c;
            t->nodes[c].parent = curr; t->nodes[c].parent_edge_idx = best;
            t->nodes[c].turn = (int8_t)t->sim_env.turn;
        }
        if (use_vl) { t->nodes[e->child].n_inflight += E->cfg.vl_count; path_push(t, k, curr, best); }
        curr = e->child;
        t->stat_depth++;
        winner = env_winner(g, &t->sim_env);
        full = env_full(g, &t->sim_env);
        if (winner != 0 || full) {
            wdl_t tw = winner != 0 ? winner_to_wdl(winner) : winner_to_wdl(0);
            node_t *leaf = &t->nodes[curr];
            leaf->is_terminal = 1; leaf->term_d = tw.d; leaf->term_p1w = tw.p1w; leaf->term_p2w = tw.p2w;
            break;
        }
    }
    t->stat_sims++;
    if (use_vl) { t->vl_leaf[k] = curr; t->vl_turn[k] = t->sim_env.turn; t->vl_envs[k] = t->sim_env; }
    else { t->cur_leaf = curr; t->cur_leaf_turn = t->sim_env.turn; }
    sim_result r; r.board = t->sim_env; r.is_terminal = 0; r.twdl.d = r.twdl.p1w = r.twdl.p2w = 0.0f;
    node_t *leaf = &t->nodes[curr];
    if (leaf->is_terminal) { r.is_terminal = 1; r.twdl.d = leaf->term_d; r.twdl.p1w = leaf->term_p1w; r.twdl.p2w = leaf->term_p2w; return r; }
    if (winner == 0 && !full) { winner = env_winner(g, &t->sim_env); full = env_full(g, &t->sim_env); }
    if (winner != 0 || full) {
        wdl_t tw = winner_to_wdl(winner != 0 ? winner : 0);
        leaf->is_terminal = 1; leaf->term_d = tw.d; leaf->term_p1w = tw.p1w; leaf->term_p2w = tw.p2w;
        r.is_terminal = 1; r.twdl = tw;
    }
    return r;
}

/* ---- expansion + backprop (MCTS.h:329-413, 591-609) ---- */
static void expand_leaf(engine_t *E, int env, tree_t *t, const float *policy) {   /* MCTS.h:329-375 */
    orc_moves valids;
    env_valid(E->game, &t->sim_env, &valids);
    int nv = valids.count;
    int off = alloc_edges(t, nv);
    node_t *leaf = &t->nodes[t->cur_leaf];
    leaf->edge_offset = off; leaf->num_edges = nv; leaf->is_expanded = 1;
    float policy_sum = 0.0f;
    for (int i = 0; i < nv; ++i) policy_sum += policy[valids.moves[i]];
    float noise[ORC_MAX_A];
    int has_noise = leaf->parent == -1 && E->cfg.dirichlet_alpha > 0.0f;
    if (has_noise) draw_noise(E, env, nv, noise);
    for (int i = 0; i < nv; ++i) {
        edge_t *e = &t->edges[off + i];
        e->action = valids.moves[i];
        e->prior = policy[valids.moves[i]] / (policy_sum + 1e-8f);
        e->child = -1;
        if (has_noise) e->noise = noise[i];
    }
    t->stat_expanded += (uint64_t)nv;
    t->stat_expansions++;
}
static void propagate(engine_t *E, tree_t *t, wdl_t w, float moves_left) {   /* MCTS.h:381-402 */
    int idx = t->cur_leaf;
    float ml = moves_left;
    const float gamma = E->cfg.value_decay;
    while (idx != -1) {
        node_t *n = &t->nodes[idx];
        n->n_visits++;
        n->W_d += w.d; n->W_p1w += w.p1w; n->W_p2w += w.p2w;
        n->M_sum += ml;
        if (E->game == ORC_C4) ml += 1.0f;     /* AUX_PLUS_ONE_PER_PLY */
        else ml = -ml;                          /* AUX_NEGATE_PER_PLY */
        idx = n->parent;
        if (gamma < 1.0f) {                     /* WDLValue::decayed, MCTSNode.h:28-31 */
            const float u = 1.0f / 3.0f;
            wdl_t d;
            d.d = gamma * w.d + (1 - gamma) * u; d.p1w = gamma * w.p1w + (1 - gamma) * u; d.p2w = gamma * w.p2w + (1 - gamma) * u;
            w = d;
        }
    }
}
static void backprop(engine_t *E, int env, tree_t *t, const float *policy, wdl_t w, float ml, int is_term) {  /* MCTS.h:407-413 */
    if (t->cur_leaf == -1) return;
    if (!is_term) expand_leaf(E, env, t, policy);
    propagate(E, t, w, is_term ? terminal_aux(E->game, &t->sim_env, &E->cfg) : ml);
}
static void prepare_vl(tree_t *t, int K) {                 /* MCTS.h:421-429 */
    if (K > t->vl_alloc) {
        t->paths = (vlent_t **)realloc(t->paths, sizeof(vlent_t *) * (size_t)K);
        t->plen = (int *)realloc(t->plen, sizeof(int) * (size_t)K);
        t->pcap = (int *)realloc(t->pcap, sizeof(int) * (size_t)K);
        t->vl_envs = (orc_env *)realloc(t->vl_envs, sizeof(orc_env) * (size_t)K);
        t->vl_leaf = (int *)realloc(t->vl_leaf, sizeof(int) * (size_t)K);
        t->vl_turn = (int *)realloc(t->vl_turn, sizeof(int) * (size_t)K);
        for (int k = t->vl_alloc; k < K; ++k) { t->paths[k] = NULL; t->plen[k] = 0; t->pcap[k] = 0; t->vl_leaf[k] = -1; t->vl_turn[k] = 1; }
        t->vl_alloc = K;
    }
    /* vector::resize(K) shrinks too; entries beyond the old size start at -1 */
    for (int k = t->vlK; k < K; ++k) { t->vl_leaf[k] = -1; t->vl_turn[k] = 1; }
    t->vlK = K;
    for (int k = 0; k < K; ++k) t->plen[k] = 0;
}
static void remove_all_vl(engine_t *E, tree_t *t, int K) { /* MCTS.h:561-581 */
    int safeK = K < t->vlK ? K : t->vlK;
    int vl = E->cfg.vl_count;
    for (int k = 0; k < safeK; ++k) {
        if (t->plen[k] > 0) t->nodes[t->root].n_inflight -= vl;
        for (int j = 0; j < t->plen[k]; ++j) {
            edge_t *e = &t->edges[t->nodes[t->paths[k][j].node].edge_offset + t->paths[k][j].edge];
            if (e->child != -1) t->nodes[e->child].n_inflight -= vl;
        }
        t->plen[k] = 0;
    }
}
static void backprop_vl(engine_t *E, int env, tree_t *t, int k, const float *policy, wdl_t w, float ml, int is_term) {  /* MCTS.h:591-609 */
    t->cur_leaf = t->vl_leaf[k]; t->cur_leaf_turn = t->vl_turn[k]; t->sim_env = t->vl_envs[k];
    if (t->cur_leaf == -1) return;
    if (!is_term && !t->nodes[t->cur_leaf].is_expanded) expand_leaf(E, env, t, policy);
    propagate(E, t, w, is_term ? terminal_aux(E->game, &t->sim_env, &E->cfg) : ml);
}

/* ------------------------------------------------------------------------------------------------
 * Batched manager  (src/cpp/BatchedMCTS.h:26-442)
 * ---------------------------------------------------------------------------------------------- */
static int sample_sym(const engine_t *E, int env, int k) {          /* BatchedMCTS.h:33-40, Othello.h:363-367 */
    uint64_t h = orc_rand(E->seed, E->epoch, 0, (uint64_t)env, (uint64_t)k);
    return E->game == ORC_C4 ? (int)(h & 1) : OTH_MCTS_SYMS[h & 3];
}
static void write_leaf(const engine_t *E, int env, int k, int flat, sim_result *r, int *sym_out,
                       int8_t *ob, float *td, float *tp1, float *tp2, uint8_t *it, int32_t *ot, uint8_t *vm) {
    const int g = E->game, A = game_A(g), S = game_S(g);
    it[flat] = (uint8_t)(r->is_terminal ? 1 : 0);
    td[flat] = r->twdl.d; tp1[flat] = r->twdl.p1w; tp2[flat] = r->twdl.p2w;
    ot[flat] = r->board.turn;
    int sym = 0;
    if (!r->is_terminal && E->cfg.use_symmetry) {                   /* BatchedMCTS.h:148-158 / 261-271 */
        sym = sample_sym(E, env, k);
        if (sym != 0) env_symmetry(g, &r->board, sym);
    }
    *sym_out = sym;
    memcpy(ob + (size_t)flat * S, r->board.board, (size_t)S);
    uint8_t *mask = vm + (size_t)flat * A;
    memset(mask, 0, (size_t)A);
    if (!r->is_terminal) {
        orc_moves m; env_valid(g, &r->board, &m);
        for (int i = 0; i < m.count; ++i) mask[m.moves[i]] = 1;
    }
}

void *orc_create(int game, int n_envs) {                    /* BatchedMCTS.h:52-58 */
    engine_t *E = (engine_t *)calloc(1, sizeof(engine_t));
    E->game = game; E->n_envs = n_envs;
    config_defaults(&E->cfg);
    E->trees = (tree_t *)malloc(sizeof(tree_t) * (size_t)n_envs);
    for (int i = 0; i < n_envs; ++i) tree_init(&E->trees[i]);
    E->pending_sym = (int *)calloc((size_t)n_envs, sizeof(int));
    return E;
}
void orc_destroy(void *h) {
    engine_t *E = (engine_t *)h;
    for (int i = 0; i < E->n_envs; ++i) tree_free(&E->trees[i]);
    free(E->trees); free(E->pending_sym); free(E);
}
void orc_set_config(void *h, const orc_config *c) { ((engine_t *)h)->cfg = *c; }
void orc_get_config(void *h, orc_config *c) { *c = ((engine_t *)h)->cfg; }
void orc_set_seed(void *h, int64_t seed) { engine_t *E = (engine_t *)h; E->seed = (uint64_t)seed; E->epoch = 0; E->noise_ctr = 0; }  /* BatchedMCTS.h:68-84 */
void orc_reset_env(void *h, int i) { engine_t *E = (engine_t *)h; if (i >= 0 && i < E->n_envs) tree_reset(&E->trees[i]); }        /* :93-99 */
void orc_prune_roots(void *h, const int32_t *actions) {     /* BatchedMCTS.h:105-112 */
    engine_t *E = (engine_t *)h;
    for (int i = 0; i < E->n_envs; ++i) prune_root(E, i, &E->trees[i], actions[i]);
}
void orc_search_batch(void *h, const int8_t *boards, const int32_t *turns, int8_t *ob, float *td, float *tp1, float *tp2,
                      uint8_t *it, int32_t *ot, uint8_t *vm) {   /* BatchedMCTS.h:119-171 */
    engine_t *E = (engine_t *)h;
    const int S = game_S(E->game);
    E->epoch++;
    for (int i = 0; i < E->n_envs; ++i) {
        orc_env cur; env_reset(E->game, &cur);
        env_import(E->game, &cur, boards + (size_t)i * S);
        cur.turn = turns[i];
        sim_result r = simulate_common(E, &E->trees[i], &cur, -1);
        write_leaf(E, i, 0, i, &r, &E->pending_sym[i], ob, td, tp1, tp2, it, ot, vm);
    }
}
void orc_backprop_batch(void *h, const float *pol, const float *d, const float *p1, const float *p2, const float *ml,
                        const uint8_t *it) {                  /* BatchedMCTS.h:176-199 */
    engine_t *E = (engine_t *)h;
    const int A = game_A(E->game);
    for (int i = 0; i < E->n_envs; ++i) {
        float policy[ORC_MAX_A];
        memcpy(policy, pol + (size_t)i * A, sizeof(float) * (size_t)A);
        if (E->pending_sym[i] != 0) inverse_symmetry_policy(E->game, E->pending_sym[i], policy);
        wdl_t w = {d[i], p1[i], p2[i]};
        backprop(E, i, &E->trees[i], policy, w, ml[i], it[i] != 0);
    }
}
void orc_remove_all_vl(void *h, int K) {                    /* BatchedMCTS.h:209-216 */
    engine_t *E = (engine_t *)h;
    for (int i = 0; i < E->n_envs; ++i) remove_all_vl(E, &E->trees[i], K);
}
void orc_search_batch_vl(void *h, int K, const int8_t *boards, const int32_t *turns, int8_t *ob, float *td, float *tp1,
                         float *tp2, uint8_t *it, int32_t *ot, int32_t *sym, uint8_t *vm) {   /* BatchedMCTS.h:227-286 */
    engine_t *E = (engine_t *)h;
    const int S = game_S(E->game);
    E->epoch++;
    for (int i = 0; i < E->n_envs; ++i) {
        orc_env cur; env_reset(E->game, &cur);
        env_import(E->game, &cur, boards + (size_t)i * S);
        cur.turn = turns[i];
        prepare_vl(&E->trees[i], K);
        for (int k = 0; k < K; ++k) {
            int flat = i * K + k, s;
            sim_result r = simulate_common(E, &E->trees[i], &cur, k);
            write_leaf(E, i, k, flat, &r, &s, ob, td, tp1, tp2, it, ot, vm);
            sym[flat] = s;
        }
    }
}
void orc_backprop_batch_vl(void *h, int K, const float *pol, const float *d, const float *p1, const float *p2,
                           const float *ml, const uint8_t *it, const int32_t *sym) {   /* BatchedMCTS.h:296-332 */
    engine_t *E = (engine_t *)h;
    const int A = game_A(E->game);
    for (int i = 0; i < E->n_envs; ++i) {
        remove_all_vl(E, &E->trees[i], K);
        for (int k = 0; k < K; ++k) {
            int flat = i * K + k;
            float policy[ORC_MAX_A];
            memcpy(policy, pol + (size_t)flat * A, sizeof(float) * (size_t)A);
            if (sym[flat] != 0) inverse_symmetry_policy(E->game, sym[flat], policy);
            wdl_t w = {d[flat], p1[flat], p2[flat]};
            backprop_vl(E, i, &E->trees[i], k, policy, w, ml[flat], it[flat] != 0);
        }
    }
}

/* RolloutEvaluator::evaluate_single (src/cpp/RolloutEvaluator.h:23-48); IEvaluator default (IEvaluator.h:56-64) */
static void evaluate_leaf(engine_t *E, int kind, int env, uint64_t playout, const orc_env *state, float *policy, wdl_t *w, float *ml) {
    const int A = game_A(E->game);
    for (int a = 0; a < A; ++a) policy[a] = 1.0f;
    *ml = 0.0f;
    if (kind == ORC_EVAL_UNIFORM) { w->d = w->p1w = w->p2w = 1.f / 3; return; }
    orc_env sim = *state;
    for (uint64_t step = 0;; ++step) {
        int win = env_winner(E->game, &sim);
        if (win != 0) { *w = winner_to_wdl(win); return; }
        if (env_full(E->game, &sim)) { *w = winner_to_wdl(0); return; }
        orc_moves m; env_valid(E->game, &sim, &m);
        uint64_t r = orc_rand(E->seed, E->epoch, 1, (uint64_t)env, (playout << 8) | step);
        env_step(E->game, &sim, m.moves[r % (uint64_t)m.count]);
    }
}
void orc_search(void *h, int kind, const int8_t *boards, const int32_t *turns, int n_playout) {   /* BatchedMCTS.h:339-407 */
    engine_t *E = (engine_t *)h;
    const int S = game_S(E->game);
    E->epoch++;
    for (int p = 0; p < n_playout; ++p) {
        for (int i = 0; i < E->n_envs; ++i) {
            orc_env cur; env_reset(E->game, &cur);
            env_import(E->game, &cur, boards + (size_t)i * S);
            cur.turn = turns[i];
            sim_result r = simulate_common(E, &E->trees[i], &cur, -1);
            float policy[ORC_MAX_A]; wdl_t w; float ml = 0.0f;
            if (r.is_terminal) { memset(policy, 0, sizeof(policy)); w = r.twdl; }
            else evaluate_leaf(E, kind, i, (uint64_t)p, &r.board, policy, &w, &ml);
            backprop(E, i, &E->trees[i], policy, w, ml, r.is_terminal);
        }
    }
}
void orc_get_counts(void *h, int32_t *out) {                /* MCTS.h:617-630, BatchedMCTS.h:413-427 */
    engine_t *E = (engine_t *)h;
    const int A = game_A(E->game);
    memset(out, 0, sizeof(int32_t) * (size_t)A * (size_t)E->n_envs);
    for (int i = 0; i < E->n_envs; ++i) {
        tree_t *t = &E->trees[i];
        const node_t *root = &t->nodes[t->root];
        if (!root->is_expanded) continue;
        for (int j = 0; j < root->num_edges; ++j) {
            const edge_t *e = &t->edges[root->edge_offset + j];
            if (e->child != -1) out[(size_t)i * A + e->action] = t->nodes[e->child].n_visits;
        }
    }
}
void orc_get_root_stats(void *h, float *out) {              /* MCTS.h:637-673, BatchedMCTS.h:435-441 */
    engine_t *E = (engine_t *)h;
    const int A = game_A(E->game), S = 6 + 8 * A;
    for (int i = 0; i < E->n_envs; ++i) {
        tree_t *t = &E->trees[i];
        float *o = out + (size_t)i * S;
        const node_t *root = &t->nodes[t->root];
        wdl_t rw = mean_wdl(root);
        o[0] = (float)root->n_visits; o[1] = mean_q(root); o[2] = mean_M(root);
        o[3] = rw.d; o[4] = rw.p1w; o[5] = rw.p2w;
        memset(o + 6, 0, sizeof(float) * 8 * (size_t)A);
        if (!root->is_expanded) continue;
        for (int j = 0; j < root->num_edges; ++j) {
            const edge_t *e = &t->edges[root->edge_offset + j];
            float *slot = o + 6 + e->action * 8;
            slot[2] = e->prior; slot[3] = e->noise;
            if (e->child != -1) {
                const node_t *c = &t->nodes[e->child];
                wdl_t cw = mean_wdl(c);
                float cm = mean_M(c);
                if (E->game == ORC_OTH) cm = -cm;
                slot[0] = (float)c->n_visits; slot[1] = mean_q(c); slot[4] = cm;
                slot[5] = cw.d; slot[6] = cw.p1w; slot[7] = cw.p2w;
            }
        }
    }
}
/* tree statistics for the roofline model: [sims, depth(edges traversed), edges scanned, edges created, nodes, edges, scanned
 * edges whose child is allocated, scanned edges whose child has visits, expansions, then per descent level 0..7: nodes scanned[8],
 * edges scanned[8], allocated among them[8], then: expanded nodes, expanded nodes with >= 2 visits, edges of the latter] summed */
void orc_get_tree_stats(void *h, uint64_t *out) {
    engine_t *E = (engine_t *)h;
    memset(out, 0, sizeof(uint64_t) * 36);
    for (int i = 0; i < E->n_envs; ++i) {
        tree_t *t = &E->trees[i];
        out[0] += t->stat_sims; out[1] += t->stat_depth; out[2] += t->stat_edges; out[3] += t->stat_expanded;
        out[4] += (uint64_t)t->ncount; out[5] += (uint64_t)t->ecount;
        out[6] += t->stat_alloc; out[7] += t->stat_seen; out[8] += t->stat_expansions;
        for (int l = 0; l < 8; ++l) { out[9 + l] += t->stat_lvl_nodes[l]; out[17 + l] += t->stat_lvl_edges[l]; out[25 + l] += t->stat_lvl_alloc[l]; }
        for (int j = 0; j < t->ncount; ++j) {            /* expanded nodes now, and those among them visited more than once */
            if (!t->nodes[j].is_expanded) continue;
            out[33]++;
            if (t->nodes[j].n_visits >= 2) { out[34]++; out[35] += (uint64_t)t->nodes[j].num_edges; }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Stand-alone env entry points (checker for env_cpp.<game>.Env and the lockstep env kernels).
 * state layout exported to Python = orc_env.
 * ---------------------------------------------------------------------------------------------- */
int orc_env_sizeof(void) { return (int)sizeof(orc_env); }
void orc_env_reset(int g, orc_env *e) { env_reset(g, e); }
void orc_env_import(int g, orc_env *e, const int8_t *b, int turn) { env_reset(g, e); env_import(g, e, b); e->turn = turn; }
void orc_env_step(int g, orc_env *e, int a) { env_step(g, e, a); }
int orc_env_winner(int g, const orc_env *e) { return env_winner(g, e); }
int orc_env_full(int g, const orc_env *e) { return env_full(g, e); }
int orc_env_done(int g, const orc_env *e) {                 /* env_connect4.h:43-44, env_othello.h:43-44 */
    return g == ORC_C4 ? (c4_winner(e) != 0 || e->n_pieces == 42) : oth_over(e);
}
int orc_env_valid(int g, const orc_env *e, int32_t *moves) {
    orc_moves m; env_valid(g, e, &m);
    for (int i = 0; i < m.count; ++i) moves[i] = m.moves[i];
    return m.count;
}
void orc_env_symmetry(int g, orc_env *e, int s) { env_symmetry(g, e, s); }
void orc_env_board(int g, orc_env *e, int8_t *out) { env_sync_to_board(g, e); memcpy(out, e->board, (size_t)game_S(g)); }
int orc_env_turn(const orc_env *e) { return e->turn; }

/*
 * Config-2 workload: lockstep random rollouts.  Game gidx plays
 * action = legal[r mod len(legal)], r = splitmix-hash(seed, gidx, ply), until done.  Per ply (before the move)
 * the checker records board/mask/turn and after the move winner/done; `digest` accumulates a hash of the
 * final state so a 1M-game device run can be cross-checked by a checksum of checksums.
 */
uint64_t orc_rollout_hash(uint64_t seed, uint64_t gidx, uint64_t ply) {
    return splitmix64(splitmix64(seed ^ 0xA5A5A5A55A5A5A5AULL) ^ (gidx << 8) ^ ply);
}
int orc_env_rollout(int g, uint64_t seed, uint64_t gidx, int max_plies, int8_t *boards, uint8_t *masks, int32_t *turns,
                    int32_t *actions, int32_t *winners, uint8_t *dones, uint64_t *digest) {
    const int A = game_A(g), S = game_S(g);
    orc_env e; env_reset(g, &e);
    int ply = 0;
    while (ply < max_plies) {
        int done = g == ORC_C4 ? (c4_winner(&e) != 0 || e.n_pieces == 42) : oth_over(&e);
        if (done) break;
        orc_moves m; env_valid(g, &e, &m);
        int a = m.moves[orc_rollout_hash(seed, gidx, (uint64_t)ply) % (uint64_t)m.count];
        if (boards) {
            env_sync_to_board(g, &e);
            memcpy(boards + (size_t)ply * S, e.board, (size_t)S);
            memset(masks + (size_t)ply * A, 0, (size_t)A);
            for (int i = 0; i < m.count; ++i) masks[(size_t)ply * A + m.moves[i]] = 1;
            turns[ply] = e.turn; actions[ply] = a;
        }
        env_step(g, &e, a);
        if (boards) {
            winners[ply] = env_winner(g, &e);
            dones[ply] = (uint8_t)(g == ORC_C4 ? (c4_winner(&e) != 0 || e.n_pieces == 42) : oth_over(&e));
        }
        ply++;
    }
    if (digest) {
        uint64_t d = splitmix64(e.bb[0]) ^ splitmix64(e.bb[1] + 0x1234567ULL);
        d = splitmix64(d ^ (uint64_t)(uint32_t)(env_winner(g, &e) + 1) ^ ((uint64_t)ply << 8) ^ ((uint64_t)(uint32_t)(e.turn + 1) << 20));
        *digest = d;
    }
    return ply;
}
/* digests of games [first, first+n) - the checksum-of-checksums checker for the 1M-game device run */
void orc_env_rollout_digests(int g, uint64_t seed, uint64_t first, int n, uint64_t *digests, int32_t *plies) {
    for (int i = 0; i < n; ++i) {
        uint64_t d = 0;
        int p = orc_env_rollout(g, seed, first + (uint64_t)i, g == ORC_C4 ? 42 : 128, NULL, NULL, NULL, NULL, NULL, NULL, &d);
        digests[i] = d;
        if (plies) plies[i] = p;
    }
}

/* ============================================================================================================
 * Gomoku (Env-only in the reference: src/cpp/Gomoku.h:11-296).  BYTE-BOARD restatement, deliberately in the
 * reference's own representation - the product keeps row bit masks (alphazero-al_b200/csrc/az_gomoku.cu), so the
 * two share no code.  Pinned against the compiled reference by tests/test_oracle_vs_ref.py.
 * ============================================================================================================ */
#define ORC_GMK_MAX 32
typedef struct orc_gomoku {
    int size, k;                                  /* board_size_, n_in_row_ */
    int8_t board[ORC_GMK_MAX * ORC_GMK_MAX];
    int turn, n_pieces, last_action, last_player, winner, done;
} orc_gomoku;

int orc_gmk_sizeof(void) { return (int)sizeof(orc_gomoku); }
void orc_gmk_reset(orc_gomoku *e) {                           /* Gomoku.h:30-39 */
    memset(e->board, 0, sizeof(e->board));
    e->turn = 1; e->n_pieces = 0; e->last_action = -1; e->last_player = 0; e->winner = 0; e->done = 0;
}
int orc_gmk_set_params(orc_gomoku *e, int size, int k) {      /* Gomoku.h:21-28, 214-222 */
    if (size <= 0 || k <= 1 || k > size || size > ORC_GMK_MAX) return -1;
    e->size = size; e->k = k;
    orc_gmk_reset(e);
    return 0;
}
static int gmk_in_bounds(const orc_gomoku *e, int r, int c) { return r >= 0 && r < e->size && c >= 0 && c < e->size; }   /* :224-227 */
static int gmk_count_direction(const orc_gomoku *e, int row, int col, int dr, int dc, int player) {   /* Gomoku.h:232-245 */
    int count = 0, r = row + dr, c = col + dc;
    while (gmk_in_bounds(e, r, c) && e->board[r * e->size + c] == player) { count++; r += dr; c += dc; }
    return count;
}
static int gmk_has_line_from(const orc_gomoku *e, int action, int player) {   /* Gomoku.h:247-263 */
    static const int DR[4] = {1, 0, 1, 1}, DC[4] = {0, 1, 1, -1};
    const int row = action / e->size, col = action % e->size;
    for (int i = 0; i < 4; ++i) {
        int fwd = gmk_count_direction(e, row, col, DR[i], DC[i], player);
        int bwd = gmk_count_direction(e, row, col, -DR[i], -DC[i], player);
        if (1 + fwd + bwd >= e->k) return 1;
    }
    return 0;
}
int orc_gmk_step(orc_gomoku *e, int action) {                 /* Gomoku.h:63-92; 1/2/3 = the three exceptions */
    const int S = e->size * e->size;
    if (e->done) return 1;
    if (action < 0 || action >= S) return 2;
    if (e->board[action] != 0) return 3;
    e->board[action] = (int8_t)e->turn;
    e->n_pieces++;
    e->last_action = action;
    e->last_player = e->turn;
    if (gmk_has_line_from(e, action, e->last_player)) { e->winner = e->last_player; e->done = 1; }
    else if (e->n_pieces == S) { e->winner = 0; e->done = 1; }
    e->turn = -e->turn;
    return 0;
}
int orc_gmk_import(orc_gomoku *e, const int8_t *src) {        /* import_board + sync_from_board: Gomoku.h:57-61,160-204 */
    const int S = e->size * e->size;
    int p1 = 0, p2 = 0;
    memcpy(e->board, src, (size_t)S);
    e->n_pieces = 0; e->last_action = -1; e->last_player = 0; e->winner = 0; e->done = 0;
    for (int i = 0; i < S; ++i) {
        int8_t v = e->board[i];
        if (v == 1) { p1++; e->n_pieces++; e->last_action = i; e->last_player = 1; }
        else if (v == -1) { p2++; e->n_pieces++; e->last_action = i; e->last_player = -1; }
        else if (v != 0) return 7;
    }
    if (p1 == p2) e->turn = 1;
    else if (p1 == p2 + 1) e->turn = -1;
    else e->turn = (e->n_pieces % 2 == 0) ? 1 : -1;
    for (int i = 0; i < S; ++i) {                             /* find_winner_full_scan: Gomoku.h:265-274 */
        int8_t v = e->board[i];
        if (v != 0 && gmk_has_line_from(e, i, v)) { e->winner = v; break; }
    }
    e->done = (e->winner != 0) || (e->n_pieces == S);
    return 0;
}
int orc_gmk_valid(const orc_gomoku *e, int32_t *moves) {      /* Gomoku.h:99-107 */
    int n = 0;
    for (int i = 0; i < e->size * e->size; ++i) if (e->board[i] == 0) moves[n++] = i;
    return n;
}
static void gmk_transform(int sym, int n, int r, int c, int *nr, int *nc) {   /* Gomoku.h:276-294 */
    switch (sym) {
        case 0: *nr = r; *nc = c; break;
        case 1: *nr = c; *nc = n - 1 - r; break;
        case 2: *nr = n - 1 - r; *nc = n - 1 - c; break;
        case 3: *nr = n - 1 - c; *nc = r; break;
        case 4: *nr = r; *nc = n - 1 - c; break;
        case 5: *nr = n - 1 - r; *nc = c; break;
        case 6: *nr = c; *nc = r; break;
        default: *nr = n - 1 - c; *nc = n - 1 - r; break;
    }
}
void orc_gmk_symmetry(orc_gomoku *e, int sym) {               /* Gomoku.h:130-158 */
    if (sym <= 0 || sym >= 8) return;
    const int n = e->size;
    int8_t nb[ORC_GMK_MAX * ORC_GMK_MAX];
    memset(nb, 0, sizeof(nb));
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < n; ++c) {
            int nr, nc; gmk_transform(sym, n, r, c, &nr, &nc);
            nb[nr * n + nc] = e->board[r * n + c];
        }
    memcpy(e->board, nb, (size_t)(n * n));
    if (e->last_action >= 0) {
        int nr, nc; gmk_transform(sym, n, e->last_action / n, e->last_action % n, &nr, &nc);
        e->last_action = nr * n + nc;
    }
}
void orc_gmk_board(const orc_gomoku *e, int8_t *out) { memcpy(out, e->board, (size_t)(e->size * e->size)); }
void orc_gmk_fields(const orc_gomoku *e, int32_t *out) {      /* turn, n_pieces, last_action, last_player, winner, done */
    out[0] = e->turn; out[1] = e->n_pieces; out[2] = e->last_action; out[3] = e->last_player; out[4] = e->winner; out[5] = e->done;
}
static uint64_t gmk_digest(const orc_gomoku *e, int plies) {
    uint64_t d = 0x9E3779B97F4A7C15ULL;
    for (int r = 0; r < e->size; ++r) {
        uint64_t lo = 0, hi = 0;
        for (int c = 0; c < e->size; ++c) {
            if (e->board[r * e->size + c] == 1) lo |= 1ULL << c;
            else if (e->board[r * e->size + c] == -1) hi |= 1ULL << c;
        }
        d = splitmix64(d ^ ((hi << 32) | lo));
    }
    return splitmix64(d ^ (uint64_t)(uint32_t)(e->winner + 1) ^ ((uint64_t)plies << 8) ^ ((uint64_t)(uint32_t)(e->turn + 1) << 20));
}
int orc_gmk_pick(uint64_t h, int n) { return (int)(((h >> 32) * (uint64_t)(uint32_t)n) >> 32); }   /* high half scaled to [0, n) */
/* lockstep random rollout of game gidx: the orc_gmk_pick(hash, #legal)-th legal move in ascending order until the game is over;
 * boards [plies, S*S] / turns / actions recorded BEFORE each move, winners / dones AFTER it (any may be NULL together) */
int orc_gmk_rollout(int size, int k, uint64_t seed, uint64_t gidx, int8_t *boards, int32_t *turns, int32_t *actions,
                    int32_t *winners, uint8_t *dones, uint64_t *digest, int8_t *final_board) {
    orc_gomoku e;
    if (orc_gmk_set_params(&e, size, k)) return -1;
    const int S = size * size;
    int32_t moves[ORC_GMK_MAX * ORC_GMK_MAX];
    int ply = 0;
    while (!e.done) {
        int n = orc_gmk_valid(&e, moves);
        int a = moves[orc_gmk_pick(orc_rollout_hash(seed, gidx, (uint64_t)ply), n)];
        if (boards) { memcpy(boards + (size_t)ply * S, e.board, (size_t)S); turns[ply] = e.turn; actions[ply] = a; }
        orc_gmk_step(&e, a);
        if (boards) { winners[ply] = e.winner; dones[ply] = (uint8_t)e.done; }
        ply++;
    }
    if (digest) *digest = gmk_digest(&e, ply);
    if (final_board) memcpy(final_board, e.board, (size_t)S);
    return ply;
}
void orc_gmk_rollout_digests(int size, int k, uint64_t seed, uint64_t first, int n, uint64_t *digests, int32_t *plies) {
    for (int i = 0; i < n; ++i) {
        uint64_t d = 0;
        int p = orc_gmk_rollout(size, k, seed, first + (uint64_t)i, NULL, NULL, NULL, NULL, NULL, &d, NULL);
        digests[i] = d;
        if (plies) plies[i] = p;
    }
}
