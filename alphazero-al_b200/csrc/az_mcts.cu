// Batched PUCT MCTS for sm_100a: N independent search trees resident in HBM, advanced in lockstep.
//
// What it computes follows the reference engine (citations relative to /root/reference/):
//   src/cpp/MCTS.h:46-675 (single tree), src/cpp/BatchedMCTS.h:26-442 (batch manager), src/cpp/MCTSNode.h.
// How it computes it is re-designed for the GPU:
//
//   * One LANE GROUP per tree for small batches (8 lanes for Connect4, 16 for Othello; 4 / 2 trees per warp).  Lane e
//     owns edge e of the node being scanned, so FPU + PUCT for all children is one coalesced sweep and the arg-max is a
//     shuffle reduction with lowest-index tie-break (the reference's strict `>` scan, MCTS.h:227).  From 8192 Connect4
//     trees up, one THREAD per tree (az_mcts_fast.cuh): every warp instruction then serves 32 trees.
//   * A node has no record of its own.  Its statistics (N, in-flight, W_d/W_p1/W_p2, M_sum), its flags and the
//     pointer to its edge block live in the 32-byte SLOT of the edge that leads to it, inside its parent's
//     block; the root's live in the per-tree TreeRec.  Descending one level is therefore ONE dependent
//     32-byte-per-lane load (the reference chases node -> edge -> child: 3 dependent cache lines, MCTS.h:140-234),
//     and back-propagation updates each statistic exactly once (no mirrored copies).
//   * The K virtual-loss descents of one tree stay sequential inside its lane group, and the K back-props are
//     applied in order k = 0..K-1 (MCTS.h:443-545, BatchedMCTS.h:309-330): that is what makes visit counts
//     bit-exact.  Parallelism comes from the thousands of trees, not from inside a tree; there are no atomics.
//   * fp32 PUCT arithmetic is written in the reference's operation order and this file is compiled with
//     -fmad=false (no FMA contraction).  log() of the integer parent visit count and
//     atan() of the integer disc difference come from host-libm look-up tables so they match glibc bit for bit.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <thread>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/azb200.h"
#include "az_games.cuh"
#include "az_rng.cuh"

namespace az {

// ------------------------------------------------------------------------------------------------
// HBM layout
// ------------------------------------------------------------------------------------------------
struct __align__(32) Slot {   // one edge + the statistics of the child it leads to (Edge + MCTSNode, MCTSNode.h:69-140)
    float prior;              // Edge.prior
    int n;                    // child n_visits
    uint32_t meta;            // [0,16) child n_inflight | [16,24) Edge.action | [24,32) flags
    uint32_t child;           // NONE = child not expanded, else (block offset in slots << 6) | num_edges
    float wd, wp1, wp2, msum; // child W_d, W_p1w, W_p2w, M_sum
};
static_assert(sizeof(Slot) == 32, "Slot must be one 32-byte sector");
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t F_ALLOC = 1u << 24;    // child node exists (Edge.child != -1)
constexpr uint32_t F_TERM = 1u << 25;     // child is_terminal
constexpr uint32_t F_WIN_P1 = 1u << 26;   // cached terminal result: neither bit = draw
constexpr uint32_t F_WIN_P2 = 1u << 27;
constexpr uint32_t F_TURN_P1 = 1u << 28;  // child node.turn == +1
// LAZY BLOCKS (Connect4, thread-per-tree lean kernels).  Only a quarter of the expanded nodes are ever visited again (measured:
// 137.6 expanded nodes per tree after a 200-simulation move, 35.0 with a second visit), yet an expansion wrote the node's whole edge
// block, num_edges x 32 bytes - most of back-prop's DRAM write traffic.  With F_LAZY set in the slot that owns the `child` pointer, the
// block's slots are RESERVED (the pointer, num_edges and the bump allocation are what they would be) but only a 32-byte HEADER
// {prior[0..6] in edge order, legal-move mask} is stored in the block's first slot.  The read-only selects score such a node from
// the header (every child is unvisited: N = 0, no sums, child = NONE, action = the e-th legal move); the back-prop of the node's
// second visit MATERIALISES the block - header -> num_edges slots - and clears the flag.  The root is never lazy (root expansions
// write the block; k_prune materialises a promoted lazy child), arena compaction materialises what it copies, and before a kernel
// that does not know about headers runs on such trees the host materialises everything (materialise_all) and stops creating them.
constexpr uint32_t F_LAZY = 1u << 29;
constexpr uint32_t INFL_MASK = 0xFFFFu;

struct __align__(64) TreeRec {   // per tree: the root's own statistics + allocator state
    Slot root;                   // root.prior unused
    uint32_t bump;               // slots used in this tree's arena
    uint32_t noise_ctr;          // Dirichlet draw counter
    uint32_t pad[6];
};
static_assert(sizeof(TreeRec) == 64, "TreeRec is one 64-byte record");

struct __align__(16) LeafHead {  // what backprop needs to know about a pending leaf (MCTS.h:56-64), un-symmetrised
    uint64_t bb0, bb1;
    int32_t turn;
    int16_t passes;
    int8_t last;
    uint8_t flags;               // LF_*
    uint32_t path_len;
    uint32_t sym;                // symmetry id handed to the evaluator for this leaf
};
static_assert(sizeof(LeafHead) == 32, "LeafHead is 32 bytes");
constexpr int PATH8 = 8;
struct __align__(64) LeafRec {   // head + the first 8 path entries: one 64-byte record gives backprop everything for depth <= 8
    LeafHead h;
    uint32_t path8[PATH8];
};
static_assert(sizeof(LeafRec) == 64, "LeafRec is 64 bytes");
static_assert(sizeof(az_root) == 32 && sizeof(az_leaf) == 32, "public records are one sector");
constexpr uint8_t LF_VALID = 1, LF_VLPENDING = 2, LF_TERM = 4, LF_WIN_P1 = 8, LF_WIN_P2 = 16;

struct Dev {   // kernel-visible view of an engine
    Slot *pool; uint32_t cap;            // cap = arena capacity (slots per tree)
    TreeRec *trees;
    float *noise; int noise_stride;      // root Dirichlet noise, [n_envs][noise_stride]
    LeafRec *leaf_vl; uint32_t *path_vl; int kcap;   // [n_envs][kcap], [n_envs][kcap][MAX_DEPTH]
    LeafRec *leaf_nv; uint32_t *path_nv;             // non-VL search_batch state
    const float *log_lut; int log_lut_n;
    const float2 *ls_lut;                // {log((n + c_base + 1) / c_base), sqrt(n)} interleaved, same length
    const float *atan_lut;               // 129 entries: disc difference -64..64
    unsigned long long *stats;           // nullptr = counters off
    int *err;                            // sticky device error flag (arena overflow)
    int n_envs;
    int env_lo, env_cnt;                 // trees [env_lo, env_lo + env_cnt) are processed by the select / back-prop launch
    int hints;                           // bit 0: streaming (.cs) stores for new edge blocks; bit 1: expansions below the root write lazy blocks (F_LAZY; off by default, AZB200_LAZY=1 / az_mcts_set_lazy).  AZB200_HINTS overrides (A/B runs)
    uint64_t seed, epoch;
    const unsigned long long *epoch_add;   // graph replays: added to `epoch` (kernel parameters are frozen in a captured graph)
    uint64_t env_base;                   // global index of env 0 (RNG keys are sharding-invariant)
};

constexpr int CTA = 128;
constexpr int AZ_DBG_WARPS = 4096;   // stats mode: per-warp start/end timestamps of the last thread-per-tree select launch

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
template <int W> __device__ __forceinline__ unsigned group_mask() {
    return (W == 32) ? 0xFFFFFFFFu : (((1u << W) - 1u) << ((threadIdx.x & 31) & ~(W - 1)));
}
// shuffle inside a lane group; a group of one lane needs no instruction at all
template <int W> __device__ __forceinline__ float gshfl(unsigned gm, float v, int src) { return W == 1 ? v : __shfl_sync(gm, v, src, W); }
template <int W> __device__ __forceinline__ int gshfl(unsigned gm, int v, int src) { return W == 1 ? v : __shfl_sync(gm, v, src, W); }
template <int W> __device__ __forceinline__ uint32_t gshfl(unsigned gm, uint32_t v, int src) { return W == 1 ? v : __shfl_sync(gm, v, src, W); }
template <int W> __device__ __forceinline__ void gsync(unsigned gm) { if (W > 1) __syncwarp(gm); }

__device__ __forceinline__ Slot ld_slot(const Slot *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    Slot s;
    s.prior = __uint_as_float(a.x); s.n = (int)a.y; s.meta = a.z; s.child = a.w;
    s.wd = __uint_as_float(b.x); s.wp1 = __uint_as_float(b.y); s.wp2 = __uint_as_float(b.z); s.msum = __uint_as_float(b.w);
    return s;
}
__device__ __forceinline__ void st_slot(Slot *p, const Slot &s) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(__float_as_uint(s.prior), (uint32_t)s.n, s.meta, s.child);
    q[1] = make_uint4(__float_as_uint(s.wd), __float_as_uint(s.wp1), __float_as_uint(s.wp2), __float_as_uint(s.msum));
}
// edge e of a lazy block, rebuilt from its header (stored in the block's first slot): {prior[e], N = 0, action = e-th legal move}
__device__ __forceinline__ Slot lazy_edge(const Slot &hdr, int e) {
    const float pr[7] = {hdr.prior, __int_as_float(hdr.n), __uint_as_float(hdr.meta), __uint_as_float(hdr.child), hdr.wd, hdr.wp1, hdr.wp2};
    uint32_t m = __float_as_uint(hdr.msum);
    float p = pr[0];
#pragma unroll
    for (int q = 1; q < 7; ++q) { if (q <= e) m &= m - 1; if (q == e) p = pr[q]; }
    Slot s; s.prior = p; s.n = 0; s.meta = (uint32_t)(__ffs((int)m) - 1) << 16; s.child = NONE; s.wd = s.wp1 = s.wp2 = s.msum = 0.0f;
    return s;
}
template <class T> __device__ __forceinline__ T ld32(const T *p) {      // 32-byte record as two 16-byte loads
    T v;
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    *reinterpret_cast<uint4 *>(&v) = q[0];
    *(reinterpret_cast<uint4 *>(&v) + 1) = q[1];
    return v;
}
template <class T> __device__ __forceinline__ void st32(T *p, const T &v) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = *reinterpret_cast<const uint4 *>(&v);
    q[1] = *(reinterpret_cast<const uint4 *>(&v) + 1);
}
// WDLValue::q on the running means (MCTSNode.h:23-25,118-128): 0 when unvisited (uniform thirds cancel)
__device__ __forceinline__ float mean_q(int n, float wp1, float wp2, bool turn_p1) {
    if (n == 0) return 0.0f;
    float inv = 1.0f / (float)n;
    float p1 = wp1 * inv, p2 = wp2 * inv;
    return turn_p1 ? (p1 - p2) : (p2 - p1);
}
__device__ __forceinline__ float mean_m(int n, float msum) { return n == 0 ? 0.0f : msum / (float)n; }   // :131-133

template <class G> __device__ __forceinline__ bool aux_enabled(const az_search_config &cfg) {
    return G::GAME == GAME_C4 ? cfg.mlh_slope > 0.0f : cfg.score_utility_factor > 0.0f;
}
template <class G> __device__ __forceinline__ float aux_utility(float child_M, float parent_M, float child_Q, const az_search_config &cfg) {
    if (G::GAME == GAME_C4) {      // Connect4.h:231-239
        if (cfg.mlh_slope <= 0.0f) return 0.0f;
        float m_diff = child_M - parent_M;
        float v = cfg.mlh_slope * m_diff, lo = -cfg.mlh_cap, hi = cfg.mlh_cap;
        float u = v < lo ? lo : (hi < v ? hi : v);   // std::clamp
        return u * child_Q;
    }
    if (cfg.score_utility_factor <= 0.0f) return 0.0f;   // Othello.h:268-274
    return cfg.score_utility_factor * child_M;
}

// ------------------------------------------------------------------------------------------------
// Dirichlet noise: gamma(alpha,1) by Marsaglia-Tsang on the counter-based stream.  (MCTS.h:113-132, 347-363.)
// RNG-dependent => distributional parity only.
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ double noise_u01(uint64_t seed, uint64_t env, uint32_t &ctr) {
    uint64_t h = az_rand(seed, 0, STREAM_NOISE, env, ctr++);
    return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
__device__ __noinline__ float gamma_draw(uint64_t seed, uint64_t env, uint32_t &ctr, float alpha) {
    double a = alpha, boost = 1.0;
    if (a < 1.0) { boost = pow(noise_u01(seed, env, ctr), 1.0 / a); a += 1.0; }
    double dd = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * dd);
    for (int it = 0; it < 64; ++it) {
        double u1 = noise_u01(seed, env, ctr), u2 = noise_u01(seed, env, ctr);
        double x = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2), v = 1.0 + c * x;
        if (v <= 0) continue;
        v = v * v * v;
        double u = noise_u01(seed, env, ctr);
        if (log(u) < 0.5 * x * x + dd - dd * v + dd * log(v)) return (float)(dd * v * boost);
    }
    return (float)(dd * boost);
}
__device__ __noinline__ void draw_root_noise(uint64_t seed, uint64_t env, uint32_t &ctr, float alpha, int ne, float *row) {
    float sum = 0.0f;
    for (int i = 0; i < ne; ++i) { float g = gamma_draw(seed, env, ctr, alpha); row[i] = g; sum += g; }
    float inv = 1.0f / (sum + 1e-8f);
    for (int i = 0; i < ne; ++i) row[i] = row[i] * inv;
}

// ------------------------------------------------------------------------------------------------
// pack / unpack: the byte-board arrays of the reference API <-> the engine's 32-byte bitboard records
// ------------------------------------------------------------------------------------------------
template <class G>
__global__ void k_pack_roots(int n, const int8_t *__restrict__ boards, const int32_t *__restrict__ turns, az_root *__restrict__ roots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;     // import_board: Connect4.h:100-129, Othello.h:92-111
    if (i >= n) return;
    const int8_t *b = boards + (size_t)i * G::S;
    uint64_t p0 = 0, p1 = 0;
    if (G::GAME == GAME_C4) {
        for (int c = 0; c < 7; ++c)
            for (int r = 5; r >= 0; --r) {                    // gravity scan: stop at the first empty cell
                int v = b[r * 7 + c];
                if (v == 0) break;
                uint64_t bit = 1ULL << (c * 7 + (5 - r));
                if (v == 1) p0 |= bit; else p1 |= bit;
            }
    } else {
        for (int j = 0; j < 64; ++j) { int v = b[j]; if (v == 1) p0 |= 1ULL << j; else if (v == -1) p1 |= 1ULL << j; }
    }
    State st; st.bb[0] = p0; st.bb[1] = p1; G::finish_import(st, turns[i]);
    az_root r; r.bb0 = p0; r.bb1 = p1; r.turn = st.turn; r.passes = st.passes; r.last = st.last; r.reserved = 0;
    st32(roots + i, r);
}
template <class G> __device__ __forceinline__ State leaf_state(const az_leaf &L) {
    State s; s.bb[0] = L.bb0; s.bb[1] = L.bb1; s.turn = L.turn; s.passes = L.passes; s.last = -1;
    return s;
}
// one thread per (leaf, cell): every store is coalesced
template <class G>
__global__ void k_unpack_leaves(int rows, const az_leaf *__restrict__ leaves, int8_t *__restrict__ ob, float *__restrict__ td,
                                float *__restrict__ tp1, float *__restrict__ tp2, uint8_t *__restrict__ it, int32_t *__restrict__ ot,
                                int32_t *__restrict__ sym, uint8_t *__restrict__ vm, float *__restrict__ planes) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t)rows * G::S) return;
    const size_t leaf = t / G::S; const int j = (int)(t - leaf * G::S);
    const az_leaf L = ld32(leaves + leaf);
    const bool term = (L.flags & AZ_LEAF_TERMINAL) != 0;
    const int bit = G::cell_bit(j);
    const int c0 = (int)((L.bb0 >> bit) & 1ULL), c1 = (int)((L.bb1 >> bit) & 1ULL);
    if (ob) ob[t] = (int8_t)(c0 - c1);
    if (planes) {
        float *pl = planes + leaf * 3 * G::S;
        const int own = L.turn == 1 ? c0 : c1, opp = L.turn == 1 ? c1 : c0;
        pl[j] = (float)own; pl[G::S + j] = (float)opp; pl[2 * G::S + j] = (float)L.turn;
    }
    if (vm) {      // valid mask of the (symmetrised) leaf; all zero for terminal leaves (BatchedMCTS.h:162-169)
        const State s = leaf_state<G>(L);
        const uint64_t legal = term ? 0ULL : G::legal(s);
        if (G::GAME == GAME_C4) { if (j < 7) vm[leaf * 7 + j] = (uint8_t)((legal >> j) & 1ULL); }
        else {
            vm[leaf * 65 + j] = (uint8_t)((legal >> j) & 1ULL);
            if (j == 0) vm[leaf * 65 + 64] = (uint8_t)((!term && legal == 0ULL && !Oth::over(s)) ? 1 : 0);
        }
    }
    if (j == 0) {
        if (td) td[leaf] = (term && !(L.flags & (AZ_LEAF_P1_WINS | AZ_LEAF_P2_WINS))) ? 1.0f : 0.0f;
        if (tp1) tp1[leaf] = (L.flags & AZ_LEAF_P1_WINS) ? 1.0f : 0.0f;
        if (tp2) tp2[leaf] = (L.flags & AZ_LEAF_P2_WINS) ? 1.0f : 0.0f;
        if (it) it[leaf] = term ? 1 : 0;
        if (ot) ot[leaf] = L.turn;
        if (sym) sym[leaf] = L.sym;
    }
}

// ------------------------------------------------------------------------------------------------
// SELECT: simulate / simulate_vl (MCTS.h:242-322, 443-545) + leaf export (BatchedMCTS.h:119-171, 227-286)
// W lanes cooperate on one tree; lane l owns edges l, l+W, l+2W, ... of the node being scanned.
// ------------------------------------------------------------------------------------------------
template <class G, int W, bool VL>
__global__ void __launch_bounds__(CTA) k_select(Dev d, az_search_config cfg, int K, const az_root *__restrict__ roots,
                                                az_leaf *__restrict__ leaves) {
    constexpr int NCH = (G::MAX_EDGES + W - 1) / W;
    __shared__ float seen_s[(G::GAME == GAME_OTH) ? CTA / W : 1][(G::GAME == GAME_OTH) ? G::MAX_EDGES + 2 : 1];   // Othello: visited children's priors in edge order
    const int gid = (blockIdx.x * CTA + threadIdx.x) / W;
    if (gid >= d.env_cnt) return;
    const int lane = threadIdx.x & (W - 1);
    const unsigned gm = group_mask<W>();
    const int env = d.env_lo + gid;
    Slot *arena = d.pool + (size_t)env * d.cap;
    TreeRec *tr = d.trees + env;
    const float *noise = d.noise + (size_t)env * d.noise_stride;
    const int vl = VL ? cfg.vl_count : 0;
    const bool use_aux = aux_enabled<G>(cfg);

    State start;
    { const az_root r = ld32(roots + env); start.bb[0] = r.bb0; start.bb[1] = r.bb1; G::finish_import(start, r.turn); }
    Slot root = ld_slot(&tr->root);        // identical copy in every lane of the group, written back once
    const uint32_t root_meta_in = root.meta;
    unsigned long long st_depth = 0, st_edges = 0;

    for (int k = 0; k < K; ++k) {
        State st = start;
        Slot cur = root;
        bool is_root = true, root_vl = false;
        uint32_t plen = 0;
        uint32_t *path = VL ? d.path_vl + ((size_t)env * d.kcap + k) * G::MAX_DEPTH : d.path_nv + (size_t)env * G::MAX_DEPTH;
        int winner = 0; bool full = false;
        uint32_t last_slot = 0;          // arena offset of the slot that leads to `cur` (valid when plen > 0)

        while (cur.child != NONE) {      // while (node.is_expanded)
            if (cur.meta & F_TERM) break;
            const int ne = (int)(cur.child & 63u);
            if (ne == 0 || plen >= (uint32_t)G::MAX_DEPTH) break;
            const uint32_t off = cur.child >> 6;
            // the edges are walked in chunks of W (lane l owns edges l, l + W, ...), only the running best stays in registers
            // (an array of all ceil(MAX_EDGES / W) chunks cost Othello ~100 registers per thread); second reads of a slot hit L1
            st_edges += (unsigned long long)ne;
            // ---- compute_fpu (MCTS.h:140-156): seen_policy summed sequentially in edge order over the visited children ----
            const int cur_infl = (int)(cur.meta & INFL_MASK);
            const float parent_q = mean_q(cur.n, cur.wp1, cur.wp2, (cur.meta & F_TURN_P1) != 0);
            float seen_policy = 0.0f;
            if (G::GAME == GAME_OTH) {
                // the priors of the visited children (0 for the others: + 0.0f is exact) go to the group's row in shared memory, then
                // every lane adds them in edge order (as k_select_ws does)
                float *sp_ = seen_s[threadIdx.x / W];
#pragma unroll 1
                for (int c0 = 0; c0 < ne; c0 += W) {
                    const int e = c0 + lane;
                    if (e < ne) { const uint2 pn = *reinterpret_cast<const uint2 *>(arena + off + e); sp_[e] = (int)pn.y > 0 ? __uint_as_float(pn.x) : 0.0f; }
                }
                gsync<W>(gm);
#pragma unroll 4
                for (int e = 0; e < ne; ++e) seen_policy += sp_[e];
                gsync<W>(gm);                                                 // (the next level of this descent overwrites the row)
            } else {
#pragma unroll 1
            for (int c0 = 0; c0 < ne; c0 += W) {
                const int e = c0 + lane;
                float pv = 0.0f;                                              // + 0.0f is exact: only visited children matter
                if (e < ne) { const uint2 pn = *reinterpret_cast<const uint2 *>(arena + off + e); if ((int)pn.y > 0) pv = __uint_as_float(pn.x); }
                unsigned vmask = (__ballot_sync(gm, pv != 0.0f) >> ((threadIdx.x & 31) & ~(W - 1))) & ((W == 32) ? 0xFFFFFFFFu : ((1u << W) - 1u));
                while (vmask) { const int l = __ffs((int)vmask) - 1; vmask &= vmask - 1; seen_policy += gshfl<W>(gm, pv, l); }
            }
            }
            const float fscale = (1.0f + parent_q) / 2.0f;
            const float eff_fpu = cfg.fpu_reduction * fscale;
            float fpu = parent_q - eff_fpu * sqrtf(seen_policy);
            fpu = (-1.0f < fpu) ? fpu : -1.0f;
            // ---- select_edge (MCTS.h:163-234) ----
            const int pn_i = cur.n + cur_infl;
            const float parent_n = (float)pn_i;
            const float parent_M = use_aux ? mean_m(cur.n, cur.msum) : 0.0f;
            const float lg = (pn_i >= 0 && pn_i < d.log_lut_n) ? d.log_lut[pn_i] : logf((parent_n + cfg.c_base + 1.0f) / cfg.c_base);
            const float c_puct = cfg.c_init + lg;
            const float sqrt_pn = sqrtf(parent_n);
            const float ne_eps = cfg.noise_epsilon;
            const bool mix_noise = is_root && ne_eps > 0.0f;
            float best_s = -INFINITY; int best_e = -1;
            Slot bs; bs.prior = 0.f; bs.n = 0; bs.meta = 0; bs.child = NONE; bs.wd = bs.wp1 = bs.wp2 = bs.msum = 0.f;   // my best edge's slot
#pragma unroll 1
            for (int c0 = 0; c0 < ne; c0 += W) {
                const int e = c0 + lane;
                if (e >= ne) continue;
                const Slot sl = ld_slot(arena + off + e);
                float eff_prior = sl.prior;
                if (mix_noise) eff_prior = (1.0f - ne_eps) * sl.prior + ne_eps * noise[e];
                const int cn = sl.n, cinf = (int)(sl.meta & INFL_MASK);
                float q_value = fpu, m_utility = 0.0f; int visits = cinf;     // unvisited: FPU, in-flight only
                if (cn > 0) {
                    visits = cn + cinf;
                    const float child_Q = mean_q(cn, sl.wp1, sl.wp2, (sl.meta & F_TURN_P1) != 0);
                    q_value = -child_Q;
                    if (use_aux) {
                        float child_M = mean_m(cn, sl.msum);
                        if (G::AUX_NEGATE) child_M = -child_M;
                        m_utility = aux_utility<G>(child_M, parent_M, child_Q, cfg);
                    }
                }
                const float u_score = c_puct * eff_prior * sqrt_pn / (1.0f + (float)visits);
                const float score = q_value + u_score + m_utility;
                if (score > best_s) { best_s = score; best_e = e; bs = sl; }   // ascending e per lane: strict > keeps the lowest index
            }
            // arg-max over the group; ties -> lowest edge index (the reference scans with a strict `>`)
#pragma unroll
            for (int o = W / 2; o > 0; o >>= 1) {
                const float os = __shfl_xor_sync(gm, best_s, o, W);
                const int oe = __shfl_xor_sync(gm, best_e, o, W);
                const bool take = oe >= 0 && (best_e < 0 || os > best_s || (os == best_s && oe < best_e));
                if (take) { best_s = os; best_e = oe; }
            }
            if (best_e < 0) break;
            if (VL && !root_vl) { root_vl = true; root.meta += (uint32_t)vl; }   // root virtual loss (MCTS.h:471-475)
            // broadcast the chosen child (held by the lane that owns the edge) to the whole group
            const int bl = best_e & (W - 1);
            Slot ch;
            ch.n = gshfl<W>(gm, bs.n, bl);
            ch.meta = gshfl<W>(gm, bs.meta, bl);
            ch.child = gshfl<W>(gm, bs.child, bl);
            ch.wp1 = gshfl<W>(gm, bs.wp1, bl);
            ch.wp2 = gshfl<W>(gm, bs.wp2, bl);
            ch.msum = gshfl<W>(gm, bs.msum, bl);
            ch.prior = 0.f; ch.wd = 0.f;
            step_group<G, W>(st, (int)((ch.meta >> 16) & 0xFFu), lane, gm);
            uint32_t nmeta = ch.meta;
            if (!(nmeta & F_ALLOC)) {      // lazy child allocation (MCTS.h:481-488): remember the child's side to move
                nmeta |= F_ALLOC;
                nmeta = st.turn == 1 ? (nmeta | F_TURN_P1) : (nmeta & ~F_TURN_P1);
            }
            nmeta += (uint32_t)vl;         // child virtual loss (MCTS.h:492)
            winner = G::winner(st);
            full = G::full(st);
            const bool term_now = winner != 0 || full;
            if (term_now) nmeta = (nmeta & ~(F_WIN_P1 | F_WIN_P2)) | F_TERM | (winner == 1 ? F_WIN_P1 : (winner == -1 ? F_WIN_P2 : 0u));
            last_slot = off + (uint32_t)best_e;
            if (lane == 0) {
                if (nmeta != ch.meta) arena[last_slot].meta = nmeta;
                path[plen] = last_slot;
            }
            ++plen;
            cur = ch; cur.meta = nmeta; is_root = false;
            if (term_now) break;
        }
        st_depth += plen;
        // ---- leaf classification (MCTS.h:512-544) ----
        bool leaf_term = (cur.meta & F_TERM) != 0;
        if (plen == 0) leaf_term = (root.meta & F_TERM) != 0;
        if (!leaf_term) {
            if (winner == 0 && !full) { winner = G::winner(st); full = G::full(st); }
            if (winner != 0 || full) {
                leaf_term = true;
                const uint32_t tf = F_TERM | (winner == 1 ? F_WIN_P1 : (winner == -1 ? F_WIN_P2 : 0u));
                if (plen == 0) { root.meta = (root.meta & ~(F_WIN_P1 | F_WIN_P2)) | tf; cur.meta = root.meta; }
                else { cur.meta = (cur.meta & ~(F_WIN_P1 | F_WIN_P2)) | tf; if (lane == 0) arena[last_slot].meta = cur.meta; }
            }
        }
        // ---- random symmetry for non-terminal leaves (BatchedMCTS.h:148-158 / 261-271) ----
        int sym = 0;
        State ex = st;
        if (!leaf_term && cfg.use_symmetry) {
            const uint64_t h = az_rand(d.seed, d.epoch + (d.epoch_add ? *d.epoch_add : 0ULL), STREAM_SYM, d.env_base + (uint64_t)env, (uint64_t)k);
            sym = G::GAME == GAME_C4 ? (int)(h & 1) : ((0x7620 >> (4 * (int)(h & 3))) & 0xF);   // Othello {0,2,6,7}
            G::symmetry(ex, sym);
        }
        if (lane == 0) {
            const uint8_t tflags = (uint8_t)(leaf_term ? (AZ_LEAF_TERMINAL | ((cur.meta & F_WIN_P1) ? AZ_LEAF_P1_WINS : 0u) |
                                                          ((cur.meta & F_WIN_P2) ? AZ_LEAF_P2_WINS : 0u)) : 0u);
            LeafHead L;    // remembered for backprop
            L.bb0 = st.bb[0]; L.bb1 = st.bb[1]; L.turn = st.turn; L.passes = (int16_t)st.passes; L.last = (int8_t)st.last;
            L.flags = (uint8_t)(LF_VALID | ((VL && plen > 0) ? LF_VLPENDING : 0) | (leaf_term ? LF_TERM : 0) |
                                ((tflags & AZ_LEAF_P1_WINS) ? LF_WIN_P1 : 0) | ((tflags & AZ_LEAF_P2_WINS) ? LF_WIN_P2 : 0));
            L.path_len = plen; L.sym = (uint32_t)sym;
            st32(&(VL ? d.leaf_vl + (size_t)env * d.kcap + k : d.leaf_nv + env)->h, L);
            az_leaf P;     // handed to the evaluator
            P.bb0 = ex.bb[0]; P.bb1 = ex.bb[1]; P.turn = (int8_t)st.turn; P.flags = tflags; P.sym = (uint8_t)sym; P.passes = (uint8_t)st.passes;
            P.reserved[0] = P.reserved[1] = P.reserved[2] = 0;
            st32(leaves + (size_t)env * K + k, P);
        }
        gsync<W>(gm);   // order this descent's in-flight updates before the next descent of the same tree
    }
    if (lane == 0 && root.meta != root_meta_in) tr->root.meta = root.meta;
    if (d.stats && lane == 0) {
        atomicAdd(d.stats + 0, (unsigned long long)K);
        atomicAdd(d.stats + 1, st_depth);
        atomicAdd(d.stats + 2, st_edges);
    }
}

// ------------------------------------------------------------------------------------------------
// BACKPROP: remove_all_vl + expand_leaf + propagate (MCTS.h:329-402, 561-609; BatchedMCTS.h:176-199, 296-332)
// ------------------------------------------------------------------------------------------------
template <class G, int W>
__device__ __forceinline__ void remove_vl_group(const Dev &d, const az_search_config &cfg, int env, int K, int lane, unsigned gm,
                                                Slot *arena, Slot &root) {
    const int vl = cfg.vl_count;
    for (int k = 0; k < K; ++k) {
        LeafHead *L = &(d.leaf_vl + (size_t)env * d.kcap + k)->h;
        const uint8_t fl = L->flags;
        if (!(fl & LF_VLPENDING)) continue;
        const uint32_t plen = L->path_len;
        const uint32_t *path = d.path_vl + ((size_t)env * d.kcap + k) * G::MAX_DEPTH;
        { int infl = (int)(root.meta & INFL_MASK) - vl; root.meta = (root.meta & ~INFL_MASK) | (uint32_t)max(infl, 0); }
        for (uint32_t j = lane; j < plen; j += W) {
            uint32_t *m = &arena[path[j]].meta;
            uint32_t v = *m;
            int infl = (int)(v & INFL_MASK) - vl;
            *m = (v & ~INFL_MASK) | (uint32_t)max(infl, 0);
        }
        gsync<W>(gm);
        if (lane == 0) L->flags = (uint8_t)(fl & ~LF_VLPENDING);
    }
}

template <class G, int W, bool VL>
__global__ void __launch_bounds__(CTA) k_backprop(Dev d, az_search_config cfg, int K, int removeK, int use_sym, const float *__restrict__ policy,
                                                  const float *__restrict__ dv, const float *__restrict__ p1v, const float *__restrict__ p2v,
                                                  const float *__restrict__ mlv, const uint8_t *__restrict__ is_term,
                                                  const int32_t *__restrict__ sym_ids) {
    constexpr int NJ = (G::A + W - 1) / W;
    __shared__ float psum_s[(G::GAME == GAME_OTH) ? CTA / W : 1][(G::GAME == GAME_OTH) ? G::MAX_EDGES + 2 : 1];   // Othello: legal policy values in order
    const int gid = (blockIdx.x * CTA + threadIdx.x) / W;
    if (gid >= d.env_cnt) return;
    const int lane = threadIdx.x & (W - 1);
    const unsigned gm = group_mask<W>();
    const int env = d.env_lo + gid;
    Slot *arena = d.pool + (size_t)env * d.cap;
    TreeRec *tr = d.trees + env;
    Slot root = ld_slot(&tr->root);     // kept in registers (identical in every lane), written back once
    uint32_t bump = tr->bump, noise_ctr = tr->noise_ctr;
    unsigned long long st_created = 0, st_expanded = 0;

    const int vl = cfg.vl_count;

    // The K simulations of a tree are applied one after the other (sums are order dependent), each a chain of dependent loads
    // (record -> path -> slots).  Start all of them now: the records' path lengths are read together, then every path slot of every
    // simulation is prefetched, so the chains below run on L1 / L2 hits.
    {
        const LeafRec *recs0 = VL ? d.leaf_vl + (size_t)env * d.kcap : d.leaf_nv + env;
        for (int k0 = 0; k0 < K; k0 += 4) {
            uint32_t pl[4]; uint32_t fl[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { pl[q] = 0; fl[q] = 0; if (k0 + q < K) { pl[q] = recs0[k0 + q].h.path_len; fl[q] = recs0[k0 + q].h.flags; } }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = k0 + q;
                if (k < K && (fl[q] & LF_VALID)) {
                    const uint32_t *pth = VL ? d.path_vl + ((size_t)env * d.kcap + k) * G::MAX_DEPTH : d.path_nv + (size_t)env * G::MAX_DEPTH;
                    for (uint32_t t = lane; t < pl[q]; t += W) asm volatile("prefetch.global.L1 [%0];" ::"l"(arena + pth[t]));
                }
            }
        }
    }

    for (int k = 0; k < K; ++k) {
        LeafRec *rp = VL ? d.leaf_vl + (size_t)env * d.kcap + k : d.leaf_nv + env;
        const LeafHead L = ld32(&rp->h);
        if (!(L.flags & LF_VALID)) continue;             // current_leaf_idx == -1 (MCTS.h:409,599)
        // The virtual loss of path k is removed in the same read-modify-write that adds simulation k's result to each
        // node (unobservable reordering: nothing in back-prop reads in-flight counts; at the end of the kernel every
        // pending loss of the first removeK paths is gone, as after remove_all_vl + K backprop_vl calls).
        const bool pending = VL && (L.flags & LF_VLPENDING) && k < removeK;
        const int dec = pending ? vl : 0;
        if (pending) { const int infl = (int)(root.meta & INFL_MASK) - vl; root.meta = (root.meta & ~INFL_MASK) | (uint32_t)max(infl, 0); }
        const size_t flat = (size_t)env * K + k;
        const bool term = is_term ? (is_term[flat] != 0) : ((L.flags & LF_TERM) != 0);
        const uint32_t plen = L.path_len;
        const uint32_t *path = VL ? d.path_vl + ((size_t)env * d.kcap + k) * G::MAX_DEPTH : d.path_nv + (size_t)env * G::MAX_DEPTH;
        State st; st.bb[0] = L.bb0; st.bb[1] = L.bb1; st.turn = L.turn; st.passes = L.passes; st.last = L.last;
        Slot *leaf_slot = plen > 0 ? arena + path[plen - 1] : nullptr;
        const uint32_t leaf_child = plen > 0 ? leaf_slot->child : root.child;

        // ---- expand_leaf (MCTS.h:329-375); VL: skipped when an earlier k already expanded it (MCTS.h:601-607) ----
        if (!term && (!VL || leaf_child == NONE)) {
            // search() never symmetrises (BatchedMCTS.h:404): use_sym == 0
            const int sym = use_sym ? (sym_ids ? sym_ids[flat] : (int)L.sym) : 0;
            const uint64_t legal = legal_group<G, W>(st, lane, gm);
            int ne; bool pass_only = false;
            if (G::GAME == GAME_OTH) {
                pass_only = legal == 0ULL && !Oth::over(st);
                ne = pass_only ? 1 : popc64(legal);
            } else ne = popc64(legal);
            const float *prow = policy + flat * G::A;
            // policy of the ORIGINAL frame: restored[a] = given[sym_action(a)]  (inverse_symmetry_policy)
            float pmine[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int a = j * W + lane;
                pmine[j] = (a < G::A) ? prow[G::sym_action(sym, a)] : 0.0f;
            }
            float psum = 0.0f;                             // summed over legal actions in ascending order
            if (G::GAME == GAME_C4) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int a = j * W + lane;
                    const float pv = (a < 7 && ((legal >> a) & 1ULL)) ? pmine[j] : 0.0f;   // + 0.0f is exact
#pragma unroll
                    for (int l = 0; l < W; ++l) if (j * W + l < 7) psum += gshfl<W>(gm, pv, l);
                }
            } else if (pass_only) {
                psum += gshfl<W>(gm, pmine[NJ - 1], 0);   // action 64 lives in lane 0, last register
            } else {
                // ascending legal actions: their policy values are compacted into shared memory in that order (rank = number of legal
                // actions below), then every lane adds them up sequentially - a load and a dependent add per legal move (a shuffle per
                // move from a bit-scan loop was 35 % of this kernel's stall samples, profiles/r2_ncu_full_oth.csv before this change)
                float *ps = psum_s[threadIdx.x / W];
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int a = j * W + lane;
                    if (a < 64 && ((legal >> a) & 1ULL)) ps[popc64(legal & ((1ULL << a) - 1ULL))] = pmine[j];
                }
                gsync<W>(gm);
#pragma unroll 4
                for (int i = 0; i < ne; ++i) psum += ps[i];
                gsync<W>(gm);                                // (the next simulation of this tree overwrites the row)
            }
            const float denom = psum + 1e-8f;
            const uint32_t alloc = (uint32_t)ne;
            if (bump + alloc > d.cap) {
                if (lane == 0) atomicExch(d.err, 1);              // host sizes the arenas so this never fires
            } else {
                const uint32_t off = bump;
                Slot ns; ns.n = 0; ns.child = NONE; ns.wd = ns.wp1 = ns.wp2 = ns.msum = 0.0f;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int a = j * W + lane;
                    if (a >= G::A) continue;
                    bool ok; int eidx;
                    if (G::GAME == GAME_OTH && a == Oth::PASS) { ok = pass_only; eidx = 0; }
                    else { ok = (legal >> (a & 63)) & 1ULL; eidx = popc64(legal & ((1ULL << (a & 63)) - 1ULL)); }
                    if (!ok) continue;
                    ns.prior = pmine[j] / denom;
                    ns.meta = (uint32_t)a << 16;
                    st_slot(arena + off + eidx, ns);
                }
                bump += alloc;
                const uint32_t cw = (off << 6) | (uint32_t)ne;
                if (plen > 0) { if (lane == 0) leaf_slot->child = cw; }
                else {
                    root.child = cw;
                    // root expansion draws Dirichlet noise when alpha > 0 (MCTS.h:347-363, leaf.parent == -1)
                    float *nrow = d.noise + (size_t)env * d.noise_stride;
                    if (cfg.dirichlet_alpha > 0.0f) {
                        if (lane == 0) draw_root_noise(d.seed, d.env_base + (uint64_t)env, noise_ctr, cfg.dirichlet_alpha, ne, nrow);
                        noise_ctr = gshfl<W>(gm, noise_ctr, 0);
                    } else {
                        for (int e = lane; e < ne; e += W) nrow[e] = 0.0f;
                    }
                }
                st_created += (unsigned long long)ne; st_expanded += 1;
            }
        }
        // ---- propagate (MCTS.h:381-402): leaf -> root, absolute WDL, per-level aux transform and decay ----
        float wd = dv[flat], w1 = p1v[flat], w2 = p2v[flat];
        float ml;
        if (term) {                                              // Game::terminal_aux (MCTS.h:412,608)
            if (G::GAME == GAME_C4) ml = 0.0f;
            else ml = d.atan_lut[(popc64(st.bb[0]) - popc64(st.bb[1])) * st.turn + 64];
        } else ml = mlv[flat];
        const float gamma = cfg.value_decay;
        const bool decay = gamma < 1.0f;
        const float u3 = 1.0f / 3.0f;
        for (uint32_t base = 0; base < plen; base += W) {
            const uint32_t t = base + lane;                      // t-th node counted from the leaf
            const bool mine = t < plen;
            Slot *sp = nullptr; int n = 0; uint32_t mt = 0; float4 w = make_float4(0, 0, 0, 0);
            if (mine) {
                sp = arena + path[plen - 1 - t];
                n = sp->n; mt = sp->meta;
                w = *reinterpret_cast<const float4 *>(&sp->wd);
            }
            float mwd = 0, mw1 = 0, mw2 = 0, mml = 0;
            const uint32_t lim = min((uint32_t)W, plen - base);
            for (uint32_t i = 0; i < lim; ++i) {                 // the value sequence is inherently sequential
                if (i == (uint32_t)lane) { mwd = wd; mw1 = w1; mw2 = w2; mml = ml; }
                if (G::AUX_PLUS_ONE) ml += 1.0f;
                if (G::AUX_NEGATE) ml = -ml;
                if (decay) { wd = gamma * wd + (1 - gamma) * u3; w1 = gamma * w1 + (1 - gamma) * u3; w2 = gamma * w2 + (1 - gamma) * u3; }
            }
            if (mine) {
                const int infl = (int)(mt & INFL_MASK) - dec;
                sp->n = n + 1;
                if (dec) sp->meta = (mt & ~INFL_MASK) | (uint32_t)max(infl, 0);
                w.x += mwd; w.y += mw1; w.z += mw2; w.w += mml;
                *reinterpret_cast<float4 *>(&sp->wd) = w;
            }
        }
        root.n += 1; root.wd += wd; root.wp1 += w1; root.wp2 += w2; root.msum += ml;
        if (pending && lane == 0) rp->h.flags = (uint8_t)(L.flags & ~LF_VLPENDING);
        gsync<W>(gm);   // the next k of this tree must see these updates (duplicate leaves, shared ancestors)
    }
    if (lane == 0) { st_slot(&tr->root, root); tr->bump = bump; tr->noise_ctr = noise_ctr; }
    if (d.stats && lane == 0) { atomicAdd(d.stats + 3, st_created); atomicAdd(d.stats + 4, st_expanded); }
}

// ================================================================================================
// Thread-per-tree variants (Connect4, lanes = 1) for large batches.
//
// With one thread per tree every warp instruction serves 32 trees, but naive per-thread loads of a tree's
// 224-byte node block touch 32 different cache lines per instruction (L1TEX wavefront bound).  k_select_t
// therefore gathers the 32 node blocks of a warp COOPERATIVELY: instruction i moves the 16-byte chunks of trees
// 2i and 2i+1 with consecutive lanes on consecutive chunks (2-4 lines per instruction instead of 32), stages
// them in shared memory (272-byte rows: conflict-free for 16-byte accesses), and each lane then reads its own
// block from shared memory.  The arithmetic is the same as k_select's, so results are bit-identical.
// ================================================================================================
template <class G, bool VL>
__global__ void __launch_bounds__(CTA, 4) k_select_t(Dev d, az_search_config cfg, int K, const az_root *__restrict__ roots,
                                                  az_leaf *__restrict__ leaves) {
    constexpr int NE = G::MAX_EDGES;       // 7
    constexpr int ROW = 17;                // uint4 per staged tree (16 + 1 pad)
    __shared__ uint4 stage[CTA / 32][32][ROW];
    const unsigned FULL = 0xFFFFFFFFu;
    const int tid = blockIdx.x * CTA + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool valid = tid < d.env_cnt;
    const int env = d.env_lo + (valid ? tid : d.env_cnt - 1);        // clamped: inactive lanes only help with the gather
    Slot *arena = d.pool + (size_t)env * d.cap;
    TreeRec *tr = d.trees + env;
    const int vl = VL ? cfg.vl_count : 0;
    const bool use_aux = aux_enabled<G>(cfg);
    const float ne_eps = cfg.noise_epsilon;

    State start;
    { const az_root r = ld32(roots + env); start.bb[0] = r.bb0; start.bb[1] = r.bb1; G::finish_import(start, r.turn); }
    Slot root = ld_slot(&tr->root);
    const uint32_t root_meta_in = root.meta;
    float nz[NE];                           // the root's Dirichlet noise, read once
#pragma unroll
    for (int e = 0; e < NE; ++e) nz[e] = ne_eps > 0.0f ? d.noise[(size_t)env * d.noise_stride + e] : 0.0f;
    unsigned long long st_depth = 0, st_edges = 0;

    for (int k = 0; k < K; ++k) {
        State st = start;
        Slot cur = root;
        bool is_root = true, root_vl = false;
        uint32_t plen = 0;
        uint32_t *path = VL ? d.path_vl + ((size_t)env * d.kcap + k) * G::MAX_DEPTH : d.path_nv + (size_t)env * G::MAX_DEPTH;
        uint32_t p8[PATH8];
#pragma unroll
        for (int j = 0; j < PATH8; ++j) p8[j] = 0;
        int winner = 0; bool full = false;
        uint32_t last_slot = 0;
        bool descending = valid && cur.child != NONE && !(cur.meta & F_TERM) && (cur.child & 63u) != 0;

        while (__any_sync(FULL, descending)) {
            // ---- cooperative gather of the warp's node blocks into shared memory ----
            const int ne = descending ? (int)(cur.child & 63u) : 0;
            const uint32_t off = descending ? (cur.child >> 6) : 0u;
            // 32-byte aligned block address with the chunk count packed into its low bits
            const unsigned long long src = (unsigned long long)(uintptr_t)(arena + off) | (unsigned long long)(2 * ne);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int t = 2 * i + (lane >> 4), part = lane & 15;
                const uint32_t lo = __shfl_sync(FULL, (uint32_t)src, t), hi = __shfl_sync(FULL, (uint32_t)(src >> 32), t);
                const int n_t = (int)(lo & 31u);
                if (part < n_t) {     // asynchronous 16-byte global->shared copy: all 16 rounds are in flight together
                    const uint4 *gp = reinterpret_cast<const uint4 *>((uintptr_t)(((unsigned long long)hi << 32) | (lo & ~31u))) + part;
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(&stage[warp][t][part]);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gp) : "memory");
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            if (descending) {
                Slot s[NE];
#pragma unroll
                for (int c = 0; c < NE; ++c) {
                    if (c < ne) {
                        const uint4 a = stage[warp][lane][2 * c], b = stage[warp][lane][2 * c + 1];
                        s[c].prior = __uint_as_float(a.x); s[c].n = (int)a.y; s[c].meta = a.z; s[c].child = a.w;
                        s[c].wd = __uint_as_float(b.x); s[c].wp1 = __uint_as_float(b.y); s[c].wp2 = __uint_as_float(b.z); s[c].msum = __uint_as_float(b.w);
                    } else { s[c].prior = 0.f; s[c].n = 0; s[c].meta = 0; s[c].child = NONE; s[c].wd = s[c].wp1 = s[c].wp2 = s[c].msum = 0.f; }
                }
                st_edges += (unsigned long long)ne;
                // ---- compute_fpu (MCTS.h:140-156) ----
                const int cur_infl = (int)(cur.meta & INFL_MASK);
                const float parent_q = mean_q(cur.n, cur.wp1, cur.wp2, (cur.meta & F_TURN_P1) != 0);
                float seen_policy = 0.0f;
#pragma unroll
                for (int c = 0; c < NE; ++c) seen_policy += (c < ne && s[c].n > 0) ? s[c].prior : 0.0f;   // + 0.0f is exact
                const float fscale = (1.0f + parent_q) / 2.0f;
                const float eff_fpu = cfg.fpu_reduction * fscale;
                float fpu = parent_q - eff_fpu * sqrtf(seen_policy);
                fpu = (-1.0f < fpu) ? fpu : -1.0f;
                // ---- select_edge (MCTS.h:163-234) ----
                const int pn_i = cur.n + cur_infl;
                const float parent_n = (float)pn_i;
                const float parent_M = use_aux ? mean_m(cur.n, cur.msum) : 0.0f;
                const float lg = (pn_i >= 0 && pn_i < d.log_lut_n) ? d.log_lut[pn_i] : logf((parent_n + cfg.c_base + 1.0f) / cfg.c_base);
                const float c_puct = cfg.c_init + lg;
                const float sqrt_pn = sqrtf(parent_n);
                const bool mix_noise = is_root && ne_eps > 0.0f;
                float best_s = -INFINITY; int best_e = -1;
#pragma unroll
                for (int c = 0; c < NE; ++c) {
                    if (c >= ne) continue;
                    float eff_prior = s[c].prior;
                    if (mix_noise) eff_prior = (1.0f - ne_eps) * s[c].prior + ne_eps * nz[c];
                    const int cn = s[c].n, cinf = (int)(s[c].meta & INFL_MASK);
                    float q_value = fpu, m_utility = 0.0f; int visits = cinf;
                    if (cn > 0) {
                        visits = cn + cinf;
                        const float child_Q = mean_q(cn, s[c].wp1, s[c].wp2, (s[c].meta & F_TURN_P1) != 0);
                        q_value = -child_Q;
                        if (use_aux) {
                            float child_M = mean_m(cn, s[c].msum);
                            if (G::AUX_NEGATE) child_M = -child_M;
                            m_utility = aux_utility<G>(child_M, parent_M, child_Q, cfg);
                        }
                    }
                    const float u_score = c_puct * eff_prior * sqrt_pn / (1.0f + (float)visits);
                    const float score = q_value + u_score + m_utility;
                    if (score > best_s) { best_s = score; best_e = c; }
                }
                if (best_e < 0) descending = false;
                else {
                    if (VL && !root_vl) { root_vl = true; root.meta += (uint32_t)vl; }
                    Slot ch = s[0];
#pragma unroll
                    for (int c = 1; c < NE; ++c) if (best_e == c) ch = s[c];
                    G::step(st, (int)((ch.meta >> 16) & 0xFFu));
                    uint32_t nmeta = ch.meta;
                    if (!(nmeta & F_ALLOC)) {
                        nmeta |= F_ALLOC;
                        nmeta = st.turn == 1 ? (nmeta | F_TURN_P1) : (nmeta & ~F_TURN_P1);
                    }
                    nmeta += (uint32_t)vl;
                    winner = G::winner(st);
                    full = G::full(st);
                    const bool term_now = winner != 0 || full;
                    if (term_now) nmeta = (nmeta & ~(F_WIN_P1 | F_WIN_P2)) | F_TERM | (winner == 1 ? F_WIN_P1 : (winner == -1 ? F_WIN_P2 : 0u));
                    last_slot = off + (uint32_t)best_e;
                    if (nmeta != ch.meta) arena[last_slot].meta = nmeta;
                    path[plen] = last_slot;
#pragma unroll
                    for (int j = 0; j < PATH8; ++j) if (plen == (uint32_t)j) p8[j] = last_slot;
                    ++plen;
                    cur = ch; cur.meta = nmeta; is_root = false;
                    descending = !term_now && cur.child != NONE && !(cur.meta & F_TERM) && (cur.child & 63u) != 0 && plen < (uint32_t)G::MAX_DEPTH;
                }
            }
            __syncwarp();      // the staging rows are reused by the next level
        }
        if (valid) {
            st_depth += plen;
            bool leaf_term = (cur.meta & F_TERM) != 0;
            if (plen == 0) leaf_term = (root.meta & F_TERM) != 0;
            if (!leaf_term) {
                if (winner == 0 && !full) { winner = G::winner(st); full = G::full(st); }
                if (winner != 0 || full) {
                    leaf_term = true;
                    const uint32_t tf = F_TERM | (winner == 1 ? F_WIN_P1 : (winner == -1 ? F_WIN_P2 : 0u));
                    if (plen == 0) { root.meta = (root.meta & ~(F_WIN_P1 | F_WIN_P2)) | tf; cur.meta = root.meta; }
                    else { cur.meta = (cur.meta & ~(F_WIN_P1 | F_WIN_P2)) | tf; arena[last_slot].meta = cur.meta; }
                }
            }
            int sym = 0;
            State ex = st;
            if (!leaf_term && cfg.use_symmetry) {
                const uint64_t h = az_rand(d.seed, d.epoch + (d.epoch_add ? *d.epoch_add : 0ULL), STREAM_SYM, d.env_base + (uint64_t)env, (uint64_t)k);
                sym = (int)(h & 1);
                G::symmetry(ex, sym);
            }
            const uint8_t tflags = (uint8_t)(leaf_term ? (AZ_LEAF_TERMINAL | ((cur.meta & F_WIN_P1) ? AZ_LEAF_P1_WINS : 0u) |
                                                          ((cur.meta & F_WIN_P2) ? AZ_LEAF_P2_WINS : 0u)) : 0u);
            LeafRec R;
            R.h.bb0 = st.bb[0]; R.h.bb1 = st.bb[1]; R.h.turn = st.turn; R.h.passes = (int16_t)st.passes; R.h.last = (int8_t)st.last;
            R.h.flags = (uint8_t)(LF_VALID | ((VL && plen > 0) ? LF_VLPENDING : 0) | (leaf_term ? LF_TERM : 0) |
                                  ((tflags & AZ_LEAF_P1_WINS) ? LF_WIN_P1 : 0) | ((tflags & AZ_LEAF_P2_WINS) ? LF_WIN_P2 : 0));
            R.h.path_len = plen; R.h.sym = (uint32_t)sym;
#pragma unroll
            for (int j = 0; j < PATH8; ++j) R.path8[j] = p8[j];
            LeafRec *dst = VL ? d.leaf_vl + (size_t)env * d.kcap + k : d.leaf_nv + env;
            st32(&dst->h, R.h);
            uint4 *pq = reinterpret_cast<uint4 *>(dst->path8);
            pq[0] = make_uint4(p8[0], p8[1], p8[2], p8[3]); pq[1] = make_uint4(p8[4], p8[5], p8[6], p8[7]);
            az_leaf P;
            P.bb0 = ex.bb[0]; P.bb1 = ex.bb[1]; P.turn = (int8_t)st.turn; P.flags = tflags; P.sym = (uint8_t)sym; P.passes = (uint8_t)st.passes;
            P.reserved[0] = P.reserved[1] = P.reserved[2] = 0;
            st32(leaves + (size_t)env * K + k, P);
        }
    }
    if (valid && root.meta != root_meta_in) tr->root.meta = root.meta;
    if (d.stats && valid) {
        atomicAdd(d.stats + 0, (unsigned long long)K);
        atomicAdd(d.stats + 1, st_depth);
        atomicAdd(d.stats + 2, st_edges);
    }
}

// Thread-per-tree back-propagation.  The virtual loss of path k is removed in the same read-modify-write that adds
// simulation k's result to each node (the order is unobservable: nothing in back-prop reads in-flight counts, and
// by the end of the kernel every pending loss is gone, exactly as after remove_all_vl + K backprop_vl calls).  The
// first 8 path entries travel inside the 64-byte LeafRec, and path nodes are updated four at a time so their loads
// overlap instead of forming a dependent chain.
template <class G, bool VL>
__global__ void __launch_bounds__(CTA, 4) k_backprop_t(Dev d, az_search_config cfg, int K, int removeK, int use_sym, const float *__restrict__ policy,
                                                    const float *__restrict__ dv, const float *__restrict__ p1v, const float *__restrict__ p2v,
                                                    const float *__restrict__ mlv, const uint8_t *__restrict__ is_term,
                                                    const int32_t *__restrict__ sym_ids) {
    static_assert(G::GAME == GAME_C4, "thread-per-tree back-prop is specialised for Connect4 (terminal aux = 0, <= 7 edges)");
    if ((int)(blockIdx.x * CTA + threadIdx.x) >= d.env_cnt) return;
    const int env = d.env_lo + (int)(blockIdx.x * CTA + threadIdx.x);
    constexpr int A = G::A;
    Slot *arena = d.pool + (size_t)env * d.cap;
    TreeRec *tr = d.trees + env;
    Slot root = ld_slot(&tr->root);
    uint32_t bump = tr->bump, noise_ctr = tr->noise_ctr;
    const int vl = cfg.vl_count;
    unsigned long long st_created = 0, st_expanded = 0;
    LeafRec *recs = VL ? d.leaf_vl + (size_t)env * d.kcap : d.leaf_nv + env;
    for (int k = 0; k < K; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(recs + k));

    for (int k = 0; k < K; ++k) {
        LeafRec *rp = recs + k;
        const LeafHead L = ld32(&rp->h);
        if (!(L.flags & LF_VALID)) continue;
        uint32_t p8[PATH8];
        { const uint4 *pq = reinterpret_cast<const uint4 *>(rp->path8); const uint4 a = pq[0], b = pq[1];
          p8[0] = a.x; p8[1] = a.y; p8[2] = a.z; p8[3] = a.w; p8[4] = b.x; p8[5] = b.y; p8[6] = b.z; p8[7] = b.w; }
        const size_t flat = (size_t)env * K + k;
        const bool term = is_term ? (is_term[flat] != 0) : ((L.flags & LF_TERM) != 0);
        const uint32_t plen = L.path_len;
        const uint32_t *path = VL ? d.path_vl + ((size_t)env * d.kcap + k) * G::MAX_DEPTH : d.path_nv + (size_t)env * G::MAX_DEPTH;
        const bool pending = VL && (L.flags & LF_VLPENDING) && k < removeK;
        const uint32_t dec = pending ? (uint32_t)vl : 0u;
        State st; st.bb[0] = L.bb0; st.bb[1] = L.bb1; st.turn = L.turn; st.passes = L.passes; st.last = L.last;
        auto path_at = [&](uint32_t j) -> uint32_t {             // j-th path entry (0 = first edge below the root)
            uint32_t v = p8[0];
#pragma unroll
            for (int q = 1; q < PATH8; ++q) if (j == (uint32_t)q) v = p8[q];
            return j < (uint32_t)PATH8 ? v : path[j];
        };
        if (pending) { const int infl = (int)(root.meta & INFL_MASK) - vl; root.meta = (root.meta & ~INFL_MASK) | (uint32_t)max(infl, 0); }
        Slot leaf = root; Slot *leaf_ptr = nullptr;
        if (plen > 0) { leaf_ptr = arena + path_at(plen - 1); leaf = ld_slot(leaf_ptr); }

        // ---- expand_leaf (MCTS.h:329-375) ----
        if (!term && (!VL || leaf.child == NONE)) {
            const int sym = use_sym ? (sym_ids ? sym_ids[flat] : (int)L.sym) : 0;
            const uint64_t legal = G::legal(st);
            const int ne = popc64(legal);
            const float *prow = policy + flat * A;
            float pm[A];
#pragma unroll
            for (int a = 0; a < A; ++a) pm[a] = prow[G::sym_action(sym, a)];
            float psum = 0.0f;
#pragma unroll
            for (int a = 0; a < A; ++a) psum += ((legal >> a) & 1ULL) ? pm[a] : 0.0f;     // ascending legal order; + 0.0f exact
            const float denom = psum + 1e-8f;
            const uint32_t alloc = (uint32_t)ne;
            if (bump + alloc > d.cap) atomicExch(d.err, 1);
            else {
                const uint32_t off = bump;
                Slot ns; ns.n = 0; ns.child = NONE; ns.wd = ns.wp1 = ns.wp2 = ns.msum = 0.0f;
                int eidx = 0;
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    if (!((legal >> a) & 1ULL)) continue;
                    ns.prior = pm[a] / denom;
                    ns.meta = (uint32_t)a << 16;
                    st_slot(arena + off + eidx, ns);
                    ++eidx;
                }
                bump += alloc;
                leaf.child = (off << 6) | (uint32_t)ne;
                if (plen == 0) {
                    float *nrow = d.noise + (size_t)env * d.noise_stride;
                    if (cfg.dirichlet_alpha > 0.0f) draw_root_noise(d.seed, d.env_base + (uint64_t)env, noise_ctr, cfg.dirichlet_alpha, ne, nrow);
                    else for (int e = 0; e < ne; ++e) nrow[e] = 0.0f;
                }
                st_created += (unsigned long long)ne; st_expanded += 1;
            }
        }
        // ---- propagate (MCTS.h:381-402) fused with the removal of this path's virtual loss ----
        float wd = dv[flat], w1 = p1v[flat], w2 = p2v[flat];
        float ml = term ? 0.0f : mlv[flat];                       // Connect4 terminal_aux = 0
        const float gamma = cfg.value_decay;
        const bool decay = gamma < 1.0f;
        const float u3 = 1.0f / 3.0f;
        auto advance = [&]() {
            if (G::AUX_PLUS_ONE) ml += 1.0f;
            if (G::AUX_NEGATE) ml = -ml;
            if (decay) { wd = gamma * wd + (1 - gamma) * u3; w1 = gamma * w1 + (1 - gamma) * u3; w2 = gamma * w2 + (1 - gamma) * u3; }
        };
        auto apply = [&](Slot &s) {
            s.n += 1; s.wd += wd; s.wp1 += w1; s.wp2 += w2; s.msum += ml;
            const int infl = (int)(s.meta & INFL_MASK) - (int)dec;
            s.meta = (s.meta & ~INFL_MASK) | (uint32_t)max(infl, 0);
        };
        if (plen == 0) {                                          // the leaf is the root
            root.child = leaf.child;
            root.n += 1; root.wd += wd; root.wp1 += w1; root.wp2 += w2; root.msum += ml;
        } else {
            apply(leaf);
            st_slot(leaf_ptr, leaf);
            advance();
            uint32_t t = 1;                                       // t-th node counted from the leaf
            while (t < plen) {
                const uint32_t cnt = min(4u, plen - t);
                Slot *sp[4]; Slot sv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) if ((uint32_t)q < cnt) sp[q] = arena + path_at(plen - 1 - (t + q));
#pragma unroll
                for (int q = 0; q < 4; ++q) if ((uint32_t)q < cnt) sv[q] = ld_slot(sp[q]);       // independent loads in flight together
#pragma unroll
                for (int q = 0; q < 4; ++q) if ((uint32_t)q < cnt) { apply(sv[q]); st_slot(sp[q], sv[q]); advance(); }
                t += cnt;
            }
            root.n += 1; root.wd += wd; root.wp1 += w1; root.wp2 += w2; root.msum += ml;
        }
        if (pending) rp->h.flags = (uint8_t)(L.flags & ~LF_VLPENDING);
    }
    st_slot(&tr->root, root); tr->bump = bump; tr->noise_ctr = noise_ctr;
    if (d.stats) { atomicAdd(d.stats + 3, st_created); atomicAdd(d.stats + 4, st_expanded); }
}

}  // namespace az
#include "az_mcts_fast.cuh"
#include "az_mcts_wave.cuh"
namespace az {

// remove_all_vl without backprop (BatchedMCTS.h:209-216)
template <class G, int W>
__global__ void __launch_bounds__(CTA) k_remove_vl(Dev d, az_search_config cfg, int K) {
    const int gid = (blockIdx.x * CTA + threadIdx.x) / W;
    if (gid >= d.n_envs) return;
    const int lane = threadIdx.x & (W - 1);
    const unsigned gm = group_mask<W>();
    TreeRec *tr = d.trees + gid;
    Slot root = ld_slot(&tr->root);
    const uint32_t before = root.meta;
    remove_vl_group<G, W>(d, cfg, gid, K, lane, gm, d.pool + (size_t)gid * d.cap, root);
    if (lane == 0 && root.meta != before) tr->root.meta = root.meta;
}

// ------------------------------------------------------------------------------------------------
// prune_root / apply_root_noise / reset (MCTS.h:77-132), one thread per tree
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void reset_tree(TreeRec *tr, float *nrow, int noise_stride) {
    Slot r; r.prior = 0.f; r.n = 0; r.meta = F_TURN_P1; r.child = NONE; r.wd = r.wp1 = r.wp2 = r.msum = 0.f;   // fresh root: turn = +1
    st_slot(&tr->root, r);
    tr->bump = 0;
    for (int e = 0; e < noise_stride; ++e) nrow[e] = 0.0f;
}
template <class G>
__global__ void k_prune(Dev d, az_search_config cfg, const int32_t *__restrict__ actions) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= d.n_envs) return;
    TreeRec *tr = d.trees + env;
    Slot *arena = d.pool + (size_t)env * d.cap;
    float *nrow = d.noise + (size_t)env * d.noise_stride;
    const Slot root = ld_slot(&tr->root);
    const int action = actions[env];
    if (root.child != NONE) {
        const int ne = (int)(root.child & 63u); const uint32_t off = root.child >> 6;
        for (int e = 0; e < ne; ++e) {
            Slot s = ld_slot(arena + off + e);
            if ((int)((s.meta >> 16) & 0xFFu) == action && (s.meta & F_ALLOC)) {
                s.prior = 0.f;
                const int cne = s.child == NONE ? 0 : (int)(s.child & 63u);
                if (s.meta & F_LAZY) {                      // the root's block is always a real one
                    Slot *blk = arena + (s.child >> 6);
                    const Slot hdr = ld_slot(blk);
                    for (int q = 0; q < cne; ++q) st_slot(blk + q, lazy_edge(hdr, q));
                    s.meta &= ~F_LAZY;
                }
                st_slot(&tr->root, s);                      // promoted child becomes the root (parent = -1)
                if (cfg.dirichlet_alpha > 0.0f && cne > 0) {  // apply_root_noise (MCTS.h:113-132)
                    uint32_t ctr = tr->noise_ctr;
                    draw_root_noise(d.seed, d.env_base + (uint64_t)env, ctr, cfg.dirichlet_alpha, cne, nrow);
                    tr->noise_ctr = ctr;
                } else {
                    for (int i = 0; i < d.noise_stride; ++i) nrow[i] = 0.0f;   // the promoted node's edges never had noise
                }
                return;
            }
        }
    }
    reset_tree(tr, nrow, d.noise_stride);
}
// ------------------------------------------------------------------------------------------------
// Arena compaction (new; the reference's node pools only ever grow within a game, MCTSNode.h:149-199).  After a re-root
// only the subtree below the played move is reachable, but the bump allocator never reclaims the rest: over a 40-ply
// self-play game a Connect4 arena grows to ~43 000 slots (1.4 MB per tree) of which a few hundred are live.  k_compact
// copies the live tree of every env breadth-first into a second pool and rewrites the child pointers; the pools are then
// swapped.  The copied blocks themselves are the BFS queue, so no extra memory is needed.  W lanes per tree: W slots of
// the queue are examined per step and their child blocks copied cooperatively.  Nothing observable depends on where a
// block lives, so search results are unchanged (tests/test_gpu_mcts.py).  Pending leaf records (a search without its
// back-prop) refer to the old offsets: they are invalidated.
// ------------------------------------------------------------------------------------------------
template <class G, int W>
__global__ void __launch_bounds__(CTA) k_compact(Dev d, Slot *__restrict__ dst_pool, unsigned int *max_bump_out) {
    constexpr int J = 4;                                         // queue windows examined per step (independent: more loads in flight)
    const int gid = (blockIdx.x * CTA + threadIdx.x) / W;
    if (gid >= d.n_envs) return;
    const int lane = threadIdx.x & (W - 1);
    const unsigned gm = group_mask<W>();
    const int gshift = (threadIdx.x & 31) & ~(W - 1);            // first lane of this group inside the warp
    const int env = gid;
    TreeRec *tr = d.trees + env;
    const Slot *src = d.pool + (size_t)env * d.cap;
    Slot *dst = dst_pool + (size_t)env * d.cap;
    if (lane == 0) {                                             // stale path offsets must never be back-propagated
        d.leaf_nv[env].h.flags = 0;
        for (int k = 0; k < d.kcap; ++k) d.leaf_vl[(size_t)env * d.kcap + k].h.flags = 0;
    }
    const uint32_t rc = tr->root.child;
    uint32_t nb = 0;
    if (rc != NONE) {
        const uint32_t ne = rc & 63u, off = rc >> 6;
        for (uint32_t e = lane; e < ne; e += W) st_slot(dst + e, ld_slot(src + off + e));
        nb = ne;
        gsync<W>(gm);
        uint32_t i = 0;
        while (i < nb) {
            const uint32_t lim = min(i + (uint32_t)(J * W), nb);
            uint32_t c[J], my_off[J], lz[J];
            uint32_t run = nb;                                   // next free slot of the new arena
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const uint32_t idx = i + (uint32_t)(j * W + lane);
                c[j] = idx < lim ? dst[idx].child : NONE;
                lz[j] = (G::GAME == GAME_C4 && idx < lim && c[j] != NONE) ? (dst[idx].meta & F_LAZY) : 0u;   // header-only block
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {                        // queue order: window j, then lane
                const uint32_t n2 = c[j] != NONE ? (c[j] & 63u) : 0u;
                uint32_t pre = n2;                               // inclusive prefix sum over the lane group
#pragma unroll
                for (int o = 1; o < W; o <<= 1) { const uint32_t v = __shfl_up_sync(gm, pre, o, W); if (lane >= o) pre += v; }
                my_off[j] = run + pre - n2;
                run += gshfl<W>(gm, pre, W - 1);
            }
#pragma unroll
            for (int j = 0; j < J; ++j) {                        // all lanes copy the child blocks of window j, four at a time
                unsigned todo = (__ballot_sync(gm, c[j] != NONE) >> gshift) & ((1u << W) - 1u);
                while (todo) {
                    Slot tmp[4]; uint32_t to[4]; bool mine[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        mine[q] = false;
                        if (!todo) continue;
                        const int l = __ffs((int)todo) - 1; todo &= todo - 1;
                        const uint32_t cl = gshfl<W>(gm, c[j], l), ol = gshfl<W>(gm, my_off[j], l);
                        const uint32_t lzl = gshfl<W>(gm, lz[j], l);
                        const uint32_t nl = cl & 63u, sl = cl >> 6;
                        if (G::MAX_EDGES <= W) {                 // one slot per lane
                            mine[q] = (uint32_t)lane < nl; to[q] = ol + (uint32_t)lane;
                            if (mine[q]) tmp[q] = lzl ? lazy_edge(ld_slot(src + sl), lane) : ld_slot(src + sl + lane);   // lazy blocks are materialised
                        } else {
                            for (uint32_t e = lane; e < nl; e += W) st_slot(dst + ol + e, ld_slot(src + sl + e));
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (mine[q]) st_slot(dst + to[q], tmp[q]);
                }
                if (c[j] != NONE) dst[i + (uint32_t)(j * W + lane)].child = (my_off[j] << 6) | (c[j] & 63u);
                if (lz[j]) dst[i + (uint32_t)(j * W + lane)].meta &= ~F_LAZY;
            }
            nb = run;
            i = lim;
            gsync<W>(gm);
        }
    }
    if (lane == 0) {
        if (rc != NONE) tr->root.child = rc & 63u;               // the root's block now starts at offset 0
        tr->bump = nb;
        atomicMax(max_bump_out, nb);
    }
}
__global__ void k_reset(Dev d, int env /* -1 = all */) {
    const int i = env >= 0 ? env : (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= d.n_envs || (env >= 0 && (blockIdx.x | threadIdx.x) != 0)) return;
    reset_tree(d.trees + i, d.noise + (size_t)i * d.noise_stride, d.noise_stride);
}
__global__ void k_init_leaf(LeafRec *leaf, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) { leaf[i].h.flags = 0; leaf[i].h.path_len = 0; leaf[i].h.sym = 0; }
}
__global__ void k_max_bump(Dev d, unsigned int *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int v = i < d.n_envs ? d.trees[i].bump : 0u;
    v = __reduce_max_sync(0xFFFFFFFFu, v);
    if ((threadIdx.x & 31) == 0) atomicMax(out, v);
}
__global__ void k_grow(const Slot *__restrict__ src, Slot *__restrict__ dst, const TreeRec *__restrict__ trees, uint32_t old_cap,
                       uint32_t new_cap) {   // one CTA per tree copies the used prefix of its arena
    const int env = blockIdx.x;
    const uint32_t used = trees[env].bump;
    const uint4 *s = reinterpret_cast<const uint4 *>(src + (size_t)env * old_cap);
    uint4 *t = reinterpret_cast<uint4 *>(dst + (size_t)env * new_cap);
    for (uint32_t i = threadIdx.x; i < used * 2; i += blockDim.x) t[i] = s[i];
}

// get_counts / get_root_stats (MCTS.h:617-673)
template <class G>
__global__ void k_counts(Dev d, int32_t *__restrict__ out) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= d.n_envs) return;
    int32_t *o = out + (size_t)env * G::A;
    for (int a = 0; a < G::A; ++a) o[a] = 0;
    const Slot root = ld_slot(&d.trees[env].root);
    if (root.child == NONE) return;
    const Slot *blk = d.pool + (size_t)env * d.cap + (root.child >> 6);
    const int ne = (int)(root.child & 63u);
    for (int e = 0; e < ne; ++e) {
        const Slot s = ld_slot(blk + e);
        if (s.meta & F_ALLOC) o[(s.meta >> 16) & 0xFFu] = s.n;
    }
}
// the same for trees [lo, lo + cnt), widened to int64 on the device (the dtype callers of the reference hold: np.array(get_all_counts()))
template <class G>
__global__ void k_counts64_range(Dev d, int lo, int cnt, int64_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    const int env = lo + i;
    int64_t *o = out + (size_t)env * G::A;
    int32_t c[G::A];
    for (int a = 0; a < G::A; ++a) c[a] = 0;
    const Slot root = ld_slot(&d.trees[env].root);
    if (root.child != NONE) {
        const Slot *blk = d.pool + (size_t)env * d.cap + (root.child >> 6);
        const int ne = (int)(root.child & 63u);
        for (int e = 0; e < ne; ++e) {
            const Slot s = ld_slot(blk + e);
            if (s.meta & F_ALLOC) c[(s.meta >> 16) & 0xFFu] = s.n;
        }
    }
    for (int a = 0; a < G::A; ++a) o[a] = (int64_t)c[a];
}
template <class G>
__global__ void k_root_stats(Dev d, float *__restrict__ out) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= d.n_envs) return;
    constexpr int SZ = 6 + 8 * G::A;
    float *o = out + (size_t)env * SZ;
    const Slot root = ld_slot(&d.trees[env].root);
    const float third = 1.f / 3;
    if (root.n == 0) { o[3] = o[4] = o[5] = third; }
    else { float inv = 1.0f / (float)root.n; o[3] = root.wd * inv; o[4] = root.wp1 * inv; o[5] = root.wp2 * inv; }
    o[0] = (float)root.n;
    o[1] = mean_q(root.n, root.wp1, root.wp2, (root.meta & F_TURN_P1) != 0);
    o[2] = mean_m(root.n, root.msum);
    for (int j = 6; j < SZ; ++j) o[j] = 0.0f;
    if (root.child == NONE) return;
    const Slot *blk = d.pool + (size_t)env * d.cap + (root.child >> 6);
    const float *nrow = d.noise + (size_t)env * d.noise_stride;
    const int ne = (int)(root.child & 63u);
    for (int e = 0; e < ne; ++e) {
        const Slot s = ld_slot(blk + e);
        float *sl = o + 6 + ((s.meta >> 16) & 0xFFu) * 8;
        sl[2] = s.prior; sl[3] = nrow[e];
        if (s.meta & F_ALLOC) {
            float cm = mean_m(s.n, s.msum);
            if (G::AUX_NEGATE) cm = -cm;
            sl[0] = (float)s.n;
            sl[1] = mean_q(s.n, s.wp1, s.wp2, (s.meta & F_TURN_P1) != 0);
            sl[4] = cm;
            if (s.n == 0) { sl[5] = sl[6] = sl[7] = third; }
            else { float inv = 1.0f / (float)s.n; sl[5] = s.wd * inv; sl[6] = s.wp1 * inv; sl[7] = s.wp2 * inv; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Built-in evaluators for search(): IEvaluator default and RolloutEvaluator (IEvaluator.h:56-64,
// RolloutEvaluator.h:23-48).  One thread per tree; integer RNG draws match oracle/az_oracle.c.
// ------------------------------------------------------------------------------------------------
template <class G>
__global__ void k_eval_builtin(Dev d, int kind, int playout, float *__restrict__ policy, float *__restrict__ dv, float *__restrict__ p1v,
                               float *__restrict__ p2v, float *__restrict__ mlv) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= d.n_envs) return;
    float *prow = policy + (size_t)env * G::A;
    const LeafHead L = ld32(&d.leaf_nv[env].h);
    mlv[env] = 0.0f;
    if (L.flags & LF_TERM) {
        for (int a = 0; a < G::A; ++a) prow[a] = 0.0f;
        const bool w1 = (L.flags & LF_WIN_P1) != 0, w2 = (L.flags & LF_WIN_P2) != 0;
        dv[env] = (!w1 && !w2) ? 1.0f : 0.0f; p1v[env] = w1 ? 1.0f : 0.0f; p2v[env] = w2 ? 1.0f : 0.0f;
        return;
    }
    for (int a = 0; a < G::A; ++a) prow[a] = 1.0f;
    if (kind == AZ_EVAL_UNIFORM) { dv[env] = p1v[env] = p2v[env] = 1.f / 3; return; }
    State s; s.bb[0] = L.bb0; s.bb[1] = L.bb1; s.turn = L.turn; s.passes = L.passes; s.last = L.last;
    int w = 0;
    for (uint64_t step = 0; step < 256; ++step) {
        w = G::winner(s);
        if (w != 0 || G::full(s)) break;
        uint64_t legal = G::legal(s);
        int cnt = popc64(legal), a;
        uint64_t r = az_rand(d.seed, d.epoch, STREAM_ROLLOUT, d.env_base + (uint64_t)env, ((uint64_t)playout << 8) | step);
        if (G::GAME == GAME_OTH && cnt == 0) a = Oth::PASS;
        else {
            int idx = (int)(r % (uint64_t)cnt);
            uint64_t v = legal;
            for (int i = 0; i < idx; ++i) v &= v - 1;
            a = ctz64(v);
        }
        G::step(s, a);
    }
    dv[env] = w == 0 ? 1.0f : 0.0f; p1v[env] = w == 1 ? 1.0f : 0.0f; p2v[env] = w == -1 ? 1.0f : 0.0f;
}

}  // namespace az

// ================================================================================================
// Host side: engine object + C ABI
// ================================================================================================
using namespace az;

static thread_local std::string g_global_err;

// Pinned host blocks handed out to callers of the *_pinned host entry points (zero-copy leaf arrays: the one device-to-host copy
// of a search lands in the memory the caller's arrays live in).  Process-wide so that a block outlives the engine it came from.
namespace {
struct PinnedBlock { void *ptr; size_t bytes; bool in_use; };
std::mutex g_pin_mu;
std::vector<PinnedBlock> g_pin;
int pinned_acquire(size_t bytes, void **out) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    int best = -1;
    for (size_t i = 0; i < g_pin.size(); ++i)
        if (g_pin[i].ptr && !g_pin[i].in_use && g_pin[i].bytes >= bytes && (best < 0 || g_pin[i].bytes < g_pin[(size_t)best].bytes)) best = (int)i;
    if (best < 0) {
        size_t slot = g_pin.size();
        for (size_t i = 0; i < g_pin.size(); ++i) if (!g_pin[i].in_use) { if (g_pin[i].ptr) cudaFreeHost(g_pin[i].ptr); g_pin[i].ptr = nullptr; slot = i; break; }
        void *p = nullptr;
        const size_t cap = bytes + bytes / 4;                 // a little slack: remainder iterations and K changes reuse the block
        if (cudaMallocHost(&p, cap) != cudaSuccess) { cudaGetLastError(); return -1; }
        if (slot == g_pin.size()) g_pin.push_back({p, cap, false}); else g_pin[slot] = {p, cap, false};
        best = (int)slot;
    }
    g_pin[(size_t)best].in_use = true;
    *out = g_pin[(size_t)best].ptr;
    return best;
}
}  // namespace

// Helper threads for the staging copies of az_mcts_playout_synthetic_host: the caller's boards are pageable memory, one core moves
// ~20 GB/s into pinned staging, and at 65 536 games that copy (2.75 MB) is what delays the start of the last shard.  The
// threads sleep on a condition variable between calls; the calling thread takes chunks from the same counter, so the call
// makes progress on its own when they are slow to wake.  AZB200_STAGE_THREADS (default 2, 0 = none).
// Copy into pinned staging that the CPU never reads back: streaming stores (no read-for-ownership of the destination lines).
#if defined(__x86_64__)
__attribute__((target("avx2"))) static void stage_copy_avx2(uint8_t *dst, const uint8_t *src, size_t n) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) { dst[i] = src[i]; ++i; }
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i)), b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 64)), d = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), a); _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 64), c); _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 96), d);
    }
    for (; i + 32 <= n; i += 32) _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i)));
    _mm_sfence();
    for (; i < n; ++i) dst[i] = src[i];
}
#endif
static void stage_copy(void *dst, const void *src, size_t n) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("AZB200_PLAIN_STAGE_COPY");
    if (avx2) { stage_copy_avx2((uint8_t *)dst, (const uint8_t *)src, n); return; }
#endif
    memcpy(dst, src, n);
}
struct StagePool {
    struct Chunk { void *dst[2]; const void *src[2]; size_t n[2]; };
    Chunk chunks[16];
    int total = 0;
    std::atomic<int> next{0};
    std::atomic<int> done[16];
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv, cv_idle;
    uint64_t gen = 0; int active = 0; bool quit = false;
    void copy_chunk(int i) {
        for (int q = 0; q < 2; ++q) if (chunks[i].n[q]) stage_copy(chunks[i].dst[q], chunks[i].src[q], chunks[i].n[q]);
        done[i].store(1, std::memory_order_release);
    }
    bool take_one() {
        const int i = next.fetch_add(1, std::memory_order_relaxed);
        if (i >= total) return false;
        copy_chunk(i);
        return true;
    }
    void worker() {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return quit || gen != seen; });
            if (quit) return;
            seen = gen; ++active;
            lk.unlock();
            while (take_one()) {}
            lk.lock();
            if (--active == 0) cv_idle.notify_all();
        }
    }
    void start(int n) {
        for (int i = 0; i < n; ++i) {
            try { threads.emplace_back([this] { worker(); }); }
            catch (...) { break; }             // no thread to be had: the calling thread copies everything itself (wait_chunk)
        }
    }
    // the job in chunks[0..count) has been filled in by the caller (only while no worker is active: begin() waits for that)
    void begin() { std::unique_lock<std::mutex> lk(mu); cv_idle.wait(lk, [&] { return active == 0; }); }
    void submit(int count) {
        { std::lock_guard<std::mutex> lk(mu); total = count; for (int i = 0; i < count; ++i) done[i].store(0, std::memory_order_relaxed); next.store(0); ++gen; }
        if (!threads.empty()) cv.notify_all();
    }
    void wait_chunk(int j) {          // the calling thread copies what is left while it waits
        while (!done[j].load(std::memory_order_acquire)) if (!take_one()) std::this_thread::yield();
    }
    ~StagePool() {
        { std::lock_guard<std::mutex> lk(mu); quit = true; }
        cv.notify_all();
        for (auto &t : threads) t.join();
    }
};

struct az_mcts {
    int game = 0, n = 0, device = 0;
    int A = 0, S = 0, W = 0, max_depth = 0, max_edges = 0;
    bool lanes_fixed = false;
    size_t bp_smem_set = 0;           // dynamic shared memory already granted to k_backprop_f
    az_search_config cfg;
    Dev d{};
    cudaStream_t stream = nullptr;
    // pools
    uint32_t cap = 0;
    uint64_t bump_bound = 0;          // conservative upper bound of max(TreeRec.bump) at the last refresh ...
    struct Bound { int lo, hi; uint64_t b; };
    std::vector<Bound> bounds;        // ... and of the tree ranges [lo, hi) that back-propagated since then
    cudaStream_t side[16] = {};        // shard streams of az_mcts_playout_synthetic_dev
    // CUDA graphs of whole playout loops (az_mcts_playout_synthetic_dev): the native loop issues ~600 launches per move at
    // ~6.5 us of host time each; a captured graph replays them with one call.  Keyed by everything a launch bakes in.
    struct GraphKey {
        int mode, n_playout, K, ns, W, variant, wave_max, hints, kcap, split; uint32_t cap; const void *ptrs[10]; az_search_config cfg; uint64_t seed, env_base;
    };
    // (split != 0: one graph per shard - `shard_exec` - launched on the shard's own stream, az_mcts_playout_synthetic_host)
    struct GraphEntry { GraphKey key; cudaGraphExec_t exec; uint64_t epoch0; int launches; uint64_t last_use; std::vector<cudaGraphExec_t> shard_exec; };
    std::vector<GraphEntry> graphs;
    unsigned long long *d_epoch_add = nullptr;
    int use_graphs = 1;               // AZB200_GRAPHS=0 disables
    bool capturing = false;           // inside a capture: no host-side arena accounting, no synchronising paths
    uint64_t graph_clock = 0, graph_replays = 0;
    Slot *pool_alt = nullptr;         // second arena pool (compaction target), allocated on first use, same capacity
    int compaction = 1;               // 0 never, 1 when the arenas are more than half full at a re-root, 2 at every re-root
    bool bound_stale = false;         // arena use changed on the device (compaction): refresh the host bound before the next back-prop
    uint64_t compactions = 0;
    uint64_t base_after_prune = 0, growth_est = 0; bool base_pending = false;   // arena-use bookkeeping between re-roots
    bool time_select = false;         // az_mcts_time_select: CUDA events around every select launch
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> sel_ev; size_t sel_used = 0; uint64_t sel_rows = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> bp_ev; size_t bp_used = 0; uint64_t bp_rows = 0;   // the same around every back-prop launch
    cudaEvent_t side_ev[16] = {};
    unsigned int *d_scratch_u32 = nullptr;
    // LUT state
    float lut_c_base = -1.0f, lut_scale = -1.0f;
    int lut_n = 0;
    float *d_log_lut = nullptr, *d_atan_lut = nullptr;
    float2 *d_ls_lut = nullptr;
    bool last_select_ro = false;      // the last select launch was read-only: its back-prop applies the leaf flags, removes no virtual loss
    // ... remembered per tree range: shards of one batch run ahead of each other on their own streams and may interleave launches of
    // different K in host order; a back-prop looks up the select of ITS range and is refused when none matches
    struct SelRec { int lo, hi, K; bool vl, ro; };
    std::vector<SelRec> sel_recs;
    bool pdl = true;                  // programmatic dependent launch for the lean kernels of the playout loop (AZB200_PDL=0 turns it off)
    bool lazy_live = false;           // the trees may hold lazy blocks (F_LAZY): only the read-only selects and k_backprop_f know them
    int variant = 1;                  // thread-per-tree kernels: 0 = first generation (k_*_t), 1 = lean (k_*_f)
    int wave_max = 131072;             // batches of at most this many descent lanes run the staggered-descent select (az_mcts_wave.cuh); 0 = off
    // VL bookkeeping
    int kcap = 0;
    int prepared_K = 0;               // vl_paths_.size() (MCTS.h:421-429)
    // host-API staging (device side)
    int io_rows = 0;                  // rows the io buffers can hold
    int8_t *io_boards_in = nullptr; int32_t *io_turns_in = nullptr; az_root *io_roots = nullptr; az_leaf *io_leaves = nullptr;
    uint8_t *io_out = nullptr;        // packed leaf arrays: boards | td | tp1 | tp2 | turns | sym | is_term | mask
    uint8_t *io_in = nullptr;         // packed eval arrays: policy | d | p1 | p2 | ml | sym | is_term
    int32_t *io_actions = nullptr; int32_t *io_counts = nullptr; float *io_stats = nullptr;
    uint8_t *h_out = nullptr, *h_in = nullptr;   // pinned mirrors of io_out / io_in
    uint8_t *h_in2 = nullptr; cudaEvent_t h_in_ev[2] = {nullptr, nullptr}; int h_in_sel = 0;   // back-prop inputs are double-buffered: the host call returns once the copy is queued
    bool err_check_pending = false;              // a back-prop was queued without reading the device error flag back
    std::vector<int8_t> last_boards; std::vector<int32_t> last_turns; bool roots_valid = false;   // host searches: roots already packed on the device
    int32_t *h_counts = nullptr;                 // pinned staging of the visit counts
    // pipelined host playout (az_mcts_playout_synthetic_host): pinned staging of the caller's boards / turns, per-shard events
    // (the copy that last read a shard's staging has completed), int64 visit counts on the device, per-shard device error flags
    uint8_t *h_stage = nullptr, *io_stage = nullptr; StagePool *stage_pool = nullptr;
    int64_t *io_counts64 = nullptr; int *h_err16 = nullptr;
    // visit counts already on the host (pinned pool block) and the tree generation they belong to: handed to the next
    // az_mcts_get_counts64_pinned if no kernel changed the trees in between
    uint64_t tree_gen = 1, counts_gen = 0; int counts_block = -1; int64_t *counts_ptr = nullptr;
    unsigned long long *d_stats = nullptr; int *d_err = nullptr;
    uint64_t launches = 0;
    bool stats_on = false;
    // stream hand-over between the host entry points (internal stream) and the *_dev entry points (caller streams)
    cudaEvent_t ev = nullptr;
    cudaStream_t user_stream = nullptr; bool user_pending = false;   // device-API work possibly still in flight
    bool internal_pending = false;                                   // un-synchronised work on the internal stream
    std::string err;
};

#define AZ_FAIL(h, code, ...)                                   \
    do {                                                        \
        char _b[512]; snprintf(_b, sizeof(_b), __VA_ARGS__);    \
        (h)->err = _b; return (code);                           \
    } while (0)
#define CU(h, call)                                                                                   \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess) AZ_FAIL(h, AZ_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e)); \
    } while (0)

static inline int grid_groups(int n, int W) { return (int)(((size_t)n * W + CTA - 1) / CTA); }
static inline int grid_threads(size_t n, int bs = 128) { return (int)((n + bs - 1) / bs); }
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// packed staging layout shared by the device buffers and their pinned mirrors
struct OutLayout { size_t boards, td, tp1, tp2, turns, sym, term, mask, total; };
struct InLayout { size_t policy, d, p1, p2, ml, sym, term, total; };
static OutLayout out_layout(size_t rows, int S, int A) {
    OutLayout L; size_t o = 0;
    L.boards = o; o = align256(o + rows * S);
    L.td = o; o = align256(o + rows * 4); L.tp1 = o; o = align256(o + rows * 4); L.tp2 = o; o = align256(o + rows * 4);
    L.turns = o; o = align256(o + rows * 4); L.sym = o; o = align256(o + rows * 4);
    L.term = o; o = align256(o + rows); L.mask = o; o = align256(o + rows * A);
    L.total = o; return L;
}
static InLayout in_layout(size_t rows, int A) {
    InLayout L; size_t o = 0;
    L.policy = o; o = align256(o + rows * A * 4);
    L.d = o; o = align256(o + rows * 4); L.p1 = o; o = align256(o + rows * 4); L.p2 = o; o = align256(o + rows * 4);
    L.ml = o; o = align256(o + rows * 4); L.sym = o; o = align256(o + rows * 4); L.term = o; o = align256(o + rows);
    L.total = o; return L;
}

template <class T> static int dev_alloc(az_mcts *h, T **p, size_t count) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) count = 1;
    CU(h, cudaMalloc((void **)p, count * sizeof(T)));
    return AZ_OK;
}

static int ensure_luts(az_mcts *h) {
    const az_search_config &c = h->cfg;
    if (!h->d_log_lut || h->lut_c_base != c.c_base || !h->d_atan_lut || h->lut_scale != c.score_scale)
        CU(h, cudaDeviceSynchronize());      // kernels in flight on caller streams may still read the old tables
    if (!h->d_log_lut || h->lut_c_base != c.c_base) {
        if (h->lut_n == 0) {
            const char *e = getenv("AZB200_LOG_LUT");
            h->lut_n = e ? std::max(1024, atoi(e)) : (1 << 18);
        }
        std::vector<float> lut((size_t)h->lut_n);
        std::vector<float2> ls((size_t)h->lut_n);
        for (int i = 0; i < h->lut_n; ++i) {
            lut[(size_t)i] = logf(((float)i + c.c_base + 1.0f) / c.c_base);   // MCTS.h:213-214, host libm
            ls[(size_t)i] = make_float2(lut[(size_t)i], sqrtf((float)i));     // MCTS.h:216 (sqrt is correctly rounded everywhere)
        }
        if (!h->d_log_lut) { int r = dev_alloc(h, &h->d_log_lut, (size_t)h->lut_n); if (r) return r; }
        if (!h->d_ls_lut) { int r = dev_alloc(h, &h->d_ls_lut, (size_t)h->lut_n); if (r) return r; }
        CU(h, cudaMemcpyAsync(h->d_log_lut, lut.data(), sizeof(float) * (size_t)h->lut_n, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaMemcpyAsync(h->d_ls_lut, ls.data(), sizeof(float2) * (size_t)h->lut_n, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        h->lut_c_base = c.c_base;
    }
    if (!h->d_atan_lut || h->lut_scale != c.score_scale) {
        float lut[129];
        for (int k = -64; k <= 64; ++k) lut[k + 64] = atanf((float)k / c.score_scale) * (2.0f / 3.14159265f);   // Othello.h:260-266
        if (!h->d_atan_lut) { int r = dev_alloc(h, &h->d_atan_lut, 129); if (r) return r; }
        CU(h, cudaMemcpyAsync(h->d_atan_lut, lut, sizeof(lut), cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        h->lut_scale = c.score_scale;
    }
    h->d.log_lut = h->d_log_lut; h->d.log_lut_n = h->lut_n; h->d.atan_lut = h->d_atan_lut; h->d.ls_lut = h->d_ls_lut;
    return AZ_OK;
}

static int ensure_kcap(az_mcts *h, int K) {
    if (K <= h->kcap) return AZ_OK;
    int nk = std::max(4, h->kcap * 2);           // a power of two: the staged back-prop addresses records with shifts
    while (nk < K) nk *= 2;
    LeafRec *nl = nullptr; uint32_t *np = nullptr;
    CU(h, cudaDeviceSynchronize());
    CU(h, cudaMalloc((void **)&nl, sizeof(LeafRec) * (size_t)h->n * nk));
    CU(h, cudaMalloc((void **)&np, sizeof(uint32_t) * (size_t)h->n * nk * h->max_depth));
    size_t cnt = (size_t)h->n * nk;
    k_init_leaf<<<grid_threads(cnt, 256), 256, 0, h->stream>>>(nl, cnt);
    // pending VL paths do not survive a regrow of K (the reference's resize keeps them, but callers never
    // change K between a search and its backprop)
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->d.leaf_vl) cudaFree(h->d.leaf_vl);
    if (h->d.path_vl) cudaFree(h->d.path_vl);
    h->d.leaf_vl = nl; h->d.path_vl = np; h->kcap = nk; h->d.kcap = nk;
    return AZ_OK;
}

static int ensure_io(az_mcts *h, int rows) {
    if (rows <= h->io_rows) return AZ_OK;
    int r = std::max(rows, h->io_rows * 2);
    CU(h, cudaDeviceSynchronize());
    const OutLayout ol = out_layout((size_t)r, h->S, h->A);
    const InLayout il = in_layout((size_t)r, h->A);
    int rc = 0;
    rc |= dev_alloc(h, &h->io_leaves, (size_t)r);
    rc |= dev_alloc(h, &h->io_out, ol.total);
    rc |= dev_alloc(h, &h->io_in, il.total);
    if (rc) return AZ_ERR_CUDA;
    if (h->h_out) { cudaFreeHost(h->h_out); h->h_out = nullptr; }
    if (h->h_in) { cudaFreeHost(h->h_in); h->h_in = nullptr; }
    if (h->h_in2) { cudaFreeHost(h->h_in2); h->h_in2 = nullptr; }
    CU(h, cudaMallocHost((void **)&h->h_out, ol.total));
    CU(h, cudaMallocHost((void **)&h->h_in, il.total));
    CU(h, cudaMallocHost((void **)&h->h_in2, il.total));
    for (int j = 0; j < 2; ++j) if (!h->h_in_ev[j]) CU(h, cudaEventCreateWithFlags(&h->h_in_ev[j], cudaEventDisableTiming));
    h->io_rows = r;
    return AZ_OK;
}

static int grow_arena(az_mcts *h, uint64_t ncap, cudaStream_t st) {
    if (ncap >= (1ull << 26)) AZ_FAIL(h, AZ_ERR_NOMEM, "tree arena would exceed 2^26 slots per tree");
    Slot *np = nullptr;
    cudaError_t e = cudaMalloc((void **)&np, sizeof(Slot) * (size_t)h->n * ncap);
    if (e != cudaSuccess) AZ_FAIL(h, AZ_ERR_NOMEM, "cannot grow tree arenas to %llu slots/tree: %s", (unsigned long long)ncap, cudaGetErrorString(e));
    CU(h, cudaDeviceSynchronize());
    k_grow<<<h->n, 256, 0, st>>>(h->d.pool, np, h->d.trees, h->cap, (uint32_t)ncap);
    CU(h, cudaStreamSynchronize(st));
    cudaFree(h->d.pool);
    if (h->pool_alt) { cudaFree(h->pool_alt); h->pool_alt = nullptr; }     // re-created with the new capacity on demand
    h->d.pool = np; h->cap = (uint32_t)ncap; h->d.cap = h->cap;
    return AZ_OK;
}

// Make sure no tree of the current range can overflow its arena during a back-prop of `sims` simulations per tree.
// The host keeps conservative upper bounds of max(TreeRec.bump) per tree range (shards of one batch advance independently).
static int ensure_arena(az_mcts *h, int sims, cudaStream_t st) {
    if (h->capturing) return AZ_OK;        // the caller reserved the whole loop's growth up front
    const uint64_t need = (uint64_t)sims * (uint64_t)(h->game == GAME_C4 ? 7 : 34);
    const int lo = h->d.env_lo, hi = lo + h->d.env_cnt;
    auto current = [&]() {
        uint64_t b = h->bump_bound;
        for (const auto &r : h->bounds) if (r.lo < hi && lo < r.hi) b = std::max(b, r.b);
        return b;
    };
    uint64_t b = current();
    if (h->bound_stale || b + need > h->cap) {
        h->bound_stale = false;
        // refresh the bound from the device (all streams: other shards of the batch may still be growing their trees)
        CU(h, cudaDeviceSynchronize());
        CU(h, cudaMemsetAsync(h->d_scratch_u32, 0, sizeof(unsigned int), st));
        k_max_bump<<<grid_threads((size_t)h->n, 256), 256, 0, st>>>(h->d, h->d_scratch_u32);
        unsigned int mx = 0;
        CU(h, cudaMemcpyAsync(&mx, h->d_scratch_u32, sizeof(mx), cudaMemcpyDeviceToHost, st));
        CU(h, cudaStreamSynchronize(st));
        h->bump_bound = mx; h->bounds.clear();
        if (h->base_pending) { h->base_after_prune = mx; h->base_pending = false; }
        b = mx;
        if (b + need > h->cap) {
            uint64_t ncap = h->cap;
            while (b + need > ncap) ncap *= 2;
            int rc = grow_arena(h, ncap, st); if (rc) return rc;
        }
    }
    b += need;
    h->bounds.erase(std::remove_if(h->bounds.begin(), h->bounds.end(), [&](const az_mcts::Bound &r) { return r.lo < hi && lo < r.hi; }),
                    h->bounds.end());
    h->bounds.push_back({lo, hi, b});
    return AZ_OK;
}

static int check_cfg(az_mcts *h, int K) {
    if (h->cfg.vl_count < 0 || (int64_t)h->cfg.vl_count * std::max(K, 1) > 65535)
        AZ_FAIL(h, AZ_ERR_INVALID, "vl_count * K must be in [0, 65535] (got vl_count=%d, K=%d)", h->cfg.vl_count, K);
    return ensure_luts(h);
}

static int auto_lanes(int game, int n) {
    const char *e = getenv("AZB200_LANES");
    if (game == GAME_OTH) {
        if (e) { int w = atoi(e); if (w == 8 || w == 16) return w; }
        return n >= 8192 ? 8 : 16;      // 8192 trees, n=400 K=4: 13.7 ms per move with 8 lanes, 17.1 with 16 (tools/exp_wave_oth.py)
    }
    if (e) { int w = atoi(e); if (w == 1 || w == 2 || w == 4 || w == 8) return w; }
    // measured with the lean thread-per-tree kernels (tools/bench_configs.py, tools/exp_wave.py, B200): one lane per tree wins
    // from 8192 trees up (16384 trees: 0.97 vs 0.89 G sims/s with 4 lanes; 32768: 1.58 vs 1.24 with 2 lanes) and, since the
    // lean kernels, below as well (n=200, K=4, ms per move, 8 lanes / 1 lane / 1 lane + staggered descents: 100 trees 2.08 /
    // 1.96 / 1.78, 1024 trees 2.26 / 2.06 / 1.91, 4096 trees 2.45 / 2.20 / 2.00)
    (void)n;
    return 1;
}

// ---- kernel dispatch over (game, lanes, VL) ----
#define AZ_DISPATCH_W(h, KERNEL, VLFLAG, GRID, STREAM, ...)                                                          \
    do {                                                                                                             \
        if ((h)->game == GAME_OTH) {                                                                                 \
            if ((h)->W == 8) KERNEL<Oth, 8, VLFLAG><<<GRID, CTA, 0, STREAM>>>(__VA_ARGS__);                          \
            else KERNEL<Oth, 16, VLFLAG><<<GRID, CTA, 0, STREAM>>>(__VA_ARGS__);                                     \
        } else switch ((h)->W) {                                                                                       \
            case 1: KERNEL<C4, 1, VLFLAG><<<GRID, CTA, 0, STREAM>>>(__VA_ARGS__); break;                             \
            case 2: KERNEL<C4, 2, VLFLAG><<<GRID, CTA, 0, STREAM>>>(__VA_ARGS__); break;                             \
            case 4: KERNEL<C4, 4, VLFLAG><<<GRID, CTA, 0, STREAM>>>(__VA_ARGS__); break;                             \
            default: KERNEL<C4, 8, VLFLAG><<<GRID, CTA, 0, STREAM>>>(__VA_ARGS__); break;                            \
        }                                                                                                            \
    } while (0)

// Small batches: one lane per descent, the K descents of a tree staggered by one level (az_mcts_wave.cuh)
// wave_max counts descent lanes = trees x group width (4 lanes per tree for K <= 4, 8 for K <= 8).  Measured (tools/exp_wave.py,
// ms per move, thread-per-tree / staggered): n=800 K=8: 2048 trees 10.5 / 6.4, 8192 trees 11.0 / 7.9, 16384 trees 12.5 / 11.7,
// 32768 trees 15.0 / 17.6; n=200 K=4: 100 trees 1.96 / 1.48, 8192 trees 2.33 / 2.01, 16384 trees 2.71 / 2.50, 32768 trees 3.33 / 3.36.
static bool use_wave(const az_mcts *h, bool vl, int K) {
    // K <= 4: trees x 4 <= wave_max (default 131 072: up to 32 768 trees), K <= 8: trees x 8.  Measured with the root-once thread-per-tree
    // select (ms per move, staggered / thread per tree): fresh roots, K = 4: 8192 trees 1.96 / 2.31, 16 384 2.43 / 2.63, 32 768 3.35 / 3.12;
    // K = 8, n = 800: 8192 7.77 / 10.5, 16 384 11.3 / 11.8, 32 768 17.7 / 14.5 (tools/exp_wave.py).  Self-play with tree reuse (deeper
    // trees: 3.3 levels per descent instead of 2.3), K = 4, ms per ply: 32 768 slots 4.15 / 4.60, 65 536 slots 6.66 / 5.89
    // (tools/exp_selfplay_breakdown.py) - the 32 768-tree case goes to the staggered kernel because self-play is what runs at that size.
    return vl && K >= 1 && K <= 8 && K <= h->kcap && h->wave_max > 0 && (int64_t)h->n * (K <= 4 ? 4 : 8) <= (int64_t)h->wave_max;
}
// Does the select launch for K simulations leave the tree untouched (read-only: back-prop applies the leaf flags)?
static bool select_is_ro(const az_mcts *h, bool vl, int K) {
    return h->game == GAME_C4 && h->W == 1 && h->variant != 0 && ((vl ? K : 1) <= RS_MAX || use_wave(h, vl, K));
}
// Copies the live tree of every env into the second pool (k_compact: breadth-first, child pointers rewritten, lazy blocks
// materialised) and swaps the pools.  must = the caller cannot go on without it (materialise_all).
static int run_compaction(az_mcts *h, cudaStream_t s, bool must) {
    if (!h->pool_alt) {
        cudaError_t e = cudaMalloc((void **)&h->pool_alt, sizeof(Slot) * (size_t)h->n * h->cap);
        if (e != cudaSuccess) {
            h->pool_alt = nullptr; cudaGetLastError();
            if (must) AZ_FAIL(h, AZ_ERR_NOMEM, "no memory for the second arena pool (needed to materialise lazy blocks)");
            return AZ_OK;                                  // no room for a second pool: keep growing instead
        }
    }
    CU(h, cudaMemsetAsync(h->d_scratch_u32 + 1, 0, sizeof(unsigned int), s));
    const int g = grid_groups(h->n, 8);
    if (h->game == GAME_C4) k_compact<C4, 8><<<g, CTA, 0, s>>>(h->d, h->pool_alt, h->d_scratch_u32 + 1);
    else k_compact<Oth, 8><<<g, CTA, 0, s>>>(h->d, h->pool_alt, h->d_scratch_u32 + 1);
    std::swap(h->d.pool, h->pool_alt);
    h->bound_stale = true; h->base_pending = true; h->compactions++; h->launches++;
    h->lazy_live = false;                                  // whatever was copied is a real block now
    h->sel_recs.clear();                                   // pending leaf records were invalidated
    CU(h, cudaGetLastError());
    return AZ_OK;
}
// A kernel that does not know lazy blocks is about to run on trees that may hold some: materialise them all (one compaction
// pass) and stop creating them for this engine (the configuration - lanes, kernel generation, K beyond the read-only select - is
// unlikely to change back).
static int materialise_all(az_mcts *h, cudaStream_t s) {
    if (!h->lazy_live) return AZ_OK;
    if (h->capturing) AZ_FAIL(h, AZ_ERR_INVALID, "internal: lazy blocks must be materialised before a graph capture");
    CU(h, cudaDeviceSynchronize());                        // other shards' streams may still be working on the old pool
    int rc = run_compaction(h, s, true); if (rc) return rc;
    h->d.hints &= ~2;
    return AZ_OK;
}
static void note_select(az_mcts *h, bool vl, int K) {
    const int lo = h->d.env_lo, hi = lo + h->d.env_cnt;
    auto &v = h->sel_recs;
    v.erase(std::remove_if(v.begin(), v.end(), [&](const az_mcts::SelRec &r) { return r.lo < hi && lo < r.hi; }), v.end());
    v.push_back({lo, hi, K, vl, h->last_select_ro});
}
static bool use_wave(const az_mcts *h, bool vl, int K);
// Will the select launch for these arguments be one that understands lazy blocks (the read-only lean kernels)?
static bool select_lazy_aware(const az_mcts *h, bool vl, int K, const az_leaf *leaves) {
    const bool lean = h->game == GAME_C4 && h->W == 1 && h->variant != 0 && (uint64_t)(h->n + 32) * h->cap < (1ull << 31) && ((uintptr_t)leaves & 31) == 0;
    return lean && (use_wave(h, vl, K) || (vl ? K : 1) <= 4 /* RS_MAX */);
}
static void launch_select_impl(az_mcts *h, bool vl, int K, const az_root *roots, az_leaf *leaves, cudaStream_t s);
static int launch_select(az_mcts *h, bool vl, int K, const az_root *roots, az_leaf *leaves, cudaStream_t s) {
    if (h->lazy_live && !select_lazy_aware(h, vl, K, leaves)) { int rc = materialise_all(h, s); if (rc) return rc; }
    launch_select_impl(h, vl, K, roots, leaves, s);
    note_select(h, vl, K);
    return AZ_OK;
}
static void launch_select_impl(az_mcts *h, bool vl, int K, const az_root *roots, az_leaf *leaves, cudaStream_t s) {
    const int cnt = h->d.env_cnt;                             // trees of this launch (a whole batch or one shard)
    const int g = grid_groups(cnt, h->W);
    if (h->game == GAME_C4 && h->W == 1 && h->variant != 0 && (uint64_t)(h->n + 32) * h->cap < (1ull << 31) &&
        ((uintptr_t)leaves & 31) == 0) {         // lean thread-per-tree kernel (32-bit chunk indices, 256-bit record stores)
        const int gf = (cnt + CTA_F - 1) / CTA_F;
        const int kk = vl ? K : 1;
        const bool aux = h->cfg.mlh_slope > 0.0f;            // aux_enabled<C4>
        const bool lz = (h->d.hints & 2) != 0;               // lazy blocks on: the variants that understand headers
        if (use_wave(h, vl, K)) {
            h->last_select_ro = true;
            const int kl = K <= 4 ? 4 : 8;
            const int gw = (int)(((size_t)cnt * kl + CTA_W - 1) / CTA_W);
#define AZ_SELECT_W(AX, KLV)                                                                                              \
    do {                                                                                                                  \
        if (lz) launch_pdl(k_select_w<C4, AX, KLV, true>, gw, CTA_W, 0, s, h->pdl, h->d, h->cfg, K, roots, leaves);       \
        else launch_pdl(k_select_w<C4, AX, KLV, false>, gw, CTA_W, 0, s, h->pdl, h->d, h->cfg, K, roots, leaves);         \
    } while (0)
            if (kl == 4) { if (aux) AZ_SELECT_W(true, 4); else AZ_SELECT_W(false, 4); }
            else { if (aux) AZ_SELECT_W(true, 8); else AZ_SELECT_W(false, 8); }
#undef AZ_SELECT_W
            h->launches++;
            return;
        }
        // K <= RS_MAX: read-only select (launch-local virtual loss, leaf flags applied by back-prop); the matching back-prop
        // launch must know (last_select_ro)
        const bool ro = kk <= RS_MAX;
        h->last_select_ro = ro;
        // launches that fill the machine on their own (more CTAs than 6 per SM) run the 128-register build, shard-sized
        // launches the spill-free 144-register build (see az_mcts_fast.cuh)
        const bool wide = gf > 6 * 148;
#define AZ_SELECT_F2(VLF, AX, ROF, LZF)                                                                                   \
    do {                                                                                                                  \
        if (wide) launch_pdl(k_select_f<C4, VLF, AX, ROF, LZF>, gf, CTA_F, 0, s, h->pdl, h->d, h->cfg, kk, roots, leaves);  \
        else launch_pdl(k_select_f_r<C4, VLF, AX, ROF, LZF>, gf, CTA_F, 0, s, h->pdl, h->d, h->cfg, kk, roots, leaves);     \
    } while (0)
#define AZ_SELECT_F(VLF, AX)                                                                                              \
    do {                                                                                                                  \
        if (ro) { if (lz) AZ_SELECT_F2(VLF, AX, true, true); else AZ_SELECT_F2(VLF, AX, true, false); }                   \
        else AZ_SELECT_F2(VLF, AX, false, false);      /* the read-write select never sees lazy blocks (materialise_all) */  \
    } while (0)
        if (vl) { if (aux) AZ_SELECT_F(true, true); else AZ_SELECT_F(true, false); }
        else { if (aux) AZ_SELECT_F(false, true); else AZ_SELECT_F(false, false); }
#undef AZ_SELECT_F
#undef AZ_SELECT_F2
        h->launches++;
        return;
    }
    h->last_select_ro = false;
    if (h->game == GAME_C4 && h->W == 1) {       // thread-per-tree kernels with cooperative block gather
        if (vl) k_select_t<C4, true><<<g, CTA, 0, s>>>(h->d, h->cfg, K, roots, leaves);
        else k_select_t<C4, false><<<g, CTA, 0, s>>>(h->d, h->cfg, 1, roots, leaves);
        h->launches++;
        return;
    }
    if (h->game == GAME_OTH && vl && K >= 2 && K <= 4 && h->wave_max > 0 && (int64_t)h->n * 16 <= (int64_t)h->wave_max) {
        // small Othello batches: a warp per tree, the K descents in 8-lane groups staggered by one level (az_mcts_wave.cuh).
        // Measured (tools/exp_wave_oth.py, n=400 K=4, ms per move, sequential / staggered): 256 trees 7.9 / 5.3, 1024 trees
        // 8.2 / 5.7, 4096 trees 9.8 / 8.6, 8192 trees 13.7 / 13.3, 16384 trees 22.7 / 25.0 - on up to 8192 trees by default
        k_select_ws<Oth, 8><<<(int)(((size_t)cnt * 32 + CTA - 1) / CTA), CTA, 0, s>>>(h->d, h->cfg, K, roots, leaves);
        h->launches++;
        return;
    }
    if (vl) AZ_DISPATCH_W(h, k_select, true, g, s, h->d, h->cfg, K, roots, leaves);
    else AZ_DISPATCH_W(h, k_select, false, g, s, h->d, h->cfg, 1, roots, leaves);
    h->launches++;
}
// Grants k_backprop_f its dynamic shared memory for K simulations per tree ahead of time (not allowed to happen lazily
// inside a stream capture).
static void bp_prepare(az_mcts *h, bool vl, int K);
static int launch_backprop(az_mcts *h, bool vl, int K, int removeK, int use_sym, const float *pol, const float *d, const float *p1,
                            const float *p2, const float *ml, const uint8_t *it, const int32_t *sym, cudaStream_t s) {
    const int cnt = h->d.env_cnt;
    const int g = grid_groups(cnt, h->W);
    h->tree_gen++;                          // visit counts change: counts already fetched to the host are stale
    {   // which select produced the leaves of this range?  (read-only selects leave work to the back-prop, the others do not)
        const int lo = h->d.env_lo, hi = lo + cnt;
        bool found = false, ro = false, mixed = false;
        for (const auto &r : h->sel_recs) {
            if (!(r.lo < hi && lo < r.hi)) continue;
            if (r.lo > lo || r.hi < hi) { mixed = true; continue; }        // covers only a part of the range
            found = true; ro = r.ro;
            if (r.vl != vl || (vl && K > r.K)) AZ_FAIL(h, AZ_ERR_INVALID, "back-prop of trees [%d, %d) with K=%d does not match their last select (K=%d)", lo, hi, vl ? K : 0, r.vl ? r.K : 0);
        }
        if (mixed && !found) AZ_FAIL(h, AZ_ERR_INVALID, "back-prop of trees [%d, %d): the range was searched in pieces; back-propagate the same pieces", lo, hi);
        if (found) h->last_select_ro = ro;
    }
    if (h->game == GAME_C4 && h->W == 1 && h->variant != 0 && ((uintptr_t)pol & 15) == 0) {
        // staged back-prop: leaf records + policy rows of a warp in shared memory (dynamic, sized by K and the record stride)
        const int kk = vl ? K : 1;
        int rec_shift = 2;                                   // non-VL: one 64-byte record (4 chunks) per tree
        if (vl) { rec_shift = 0; while ((1 << rec_shift) < 4 * h->kcap) ++rec_shift; }
        const size_t smem = backprop_f_smem_per_warp(kk, rec_shift) * (CTA_F / 32);
        if ((!vl || (1 << rec_shift) == 4 * h->kcap) && smem <= 200 * 1024) {
            const int gf = (cnt + CTA_F - 1) / CTA_F;
            if (smem > h->bp_smem_set) {                     // (never inside a graph capture: the playout loop calls bp_prepare first)
#define AZ_BP_ATTR(KF) cudaFuncSetAttribute(KF, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                AZ_BP_ATTR((k_backprop_f<C4, true, true>)); AZ_BP_ATTR((k_backprop_f<C4, true, false>));
                AZ_BP_ATTR((k_backprop_f<C4, false, true>)); AZ_BP_ATTR((k_backprop_f<C4, false, false>));
                AZ_BP_ATTR((k_backprop_f_r<C4, true, true>)); AZ_BP_ATTR((k_backprop_f_r<C4, true, false>));
                AZ_BP_ATTR((k_backprop_f_r<C4, false, true>)); AZ_BP_ATTR((k_backprop_f_r<C4, false, false>));
                AZ_BP_ATTR((k_backprop_f<C4, true, true, true>)); AZ_BP_ATTR((k_backprop_f<C4, false, true, true>));
                AZ_BP_ATTR((k_backprop_f_r<C4, true, true, true>)); AZ_BP_ATTR((k_backprop_f_r<C4, false, true, true>));
#undef AZ_BP_ATTR
                h->bp_smem_set = smem;
            }
            const bool ro = h->last_select_ro, wide = gf > 6 * 148;
            const bool lz = ro && (h->d.hints & 2) != 0;
            if (lz) h->lazy_live = true;                     // expansions below the root write headers only (F_LAZY)
#define AZ_BP_F2(VLF, ROF, LZF, KARG, RARG)                                                                               \
    do {                                                                                                                  \
        if (wide) launch_pdl(k_backprop_f<C4, VLF, ROF, LZF>, gf, CTA_F, smem, s, h->pdl, h->d, h->cfg, KARG, RARG, use_sym, rec_shift, pol, d, p1, p2, ml, it, sym);    \
        else launch_pdl(k_backprop_f_r<C4, VLF, ROF, LZF>, gf, CTA_F, smem, s, h->pdl, h->d, h->cfg, KARG, RARG, use_sym, rec_shift, pol, d, p1, p2, ml, it, sym);      \
    } while (0)
#define AZ_BP_F(VLF, KARG, RARG)                                                                                          \
    do {                                                                                                                  \
        if (ro) { if (lz) AZ_BP_F2(VLF, true, true, KARG, RARG); else AZ_BP_F2(VLF, true, false, KARG, RARG); }           \
        else AZ_BP_F2(VLF, false, false, KARG, RARG);                                                                     \
    } while (0)
            if (vl) AZ_BP_F(true, kk, removeK); else AZ_BP_F(false, 1, 0);
#undef AZ_BP_F2
#undef AZ_BP_F
            h->launches++;
            return AZ_OK;
        }
    }
    if (h->last_select_ro) AZ_FAIL(h, AZ_ERR_INVALID, "back-prop after a read-only select needs a 16-byte aligned policy buffer");
    if (h->game == GAME_C4 && h->W == 1) {
        if (vl) k_backprop_t<C4, true><<<g, CTA, 0, s>>>(h->d, h->cfg, K, removeK, use_sym, pol, d, p1, p2, ml, it, sym);
        else k_backprop_t<C4, false><<<g, CTA, 0, s>>>(h->d, h->cfg, 1, 0, use_sym, pol, d, p1, p2, ml, it, sym);
        h->launches++;
        return AZ_OK;
    }
    if (vl) AZ_DISPATCH_W(h, k_backprop, true, g, s, h->d, h->cfg, K, removeK, use_sym, pol, d, p1, p2, ml, it, sym);
    else AZ_DISPATCH_W(h, k_backprop, false, g, s, h->d, h->cfg, 1, 0, use_sym, pol, d, p1, p2, ml, it, sym);
    h->launches++;
    return AZ_OK;
}

static void bp_prepare(az_mcts *h, bool vl, int K) {
    if (!(h->game == GAME_C4 && h->W == 1 && h->variant != 0)) return;
    const int kk = vl ? K : 1;
    int rec_shift = 2;
    if (vl) { rec_shift = 0; while ((1 << rec_shift) < 4 * h->kcap) ++rec_shift; }
    const size_t smem = backprop_f_smem_per_warp(kk, rec_shift) * (CTA_F / 32);
    if (smem > h->bp_smem_set && smem <= 200 * 1024) {
#define AZ_BP_ATTR(KF) cudaFuncSetAttribute(KF, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
        AZ_BP_ATTR((k_backprop_f<C4, true, true>)); AZ_BP_ATTR((k_backprop_f<C4, true, false>));
        AZ_BP_ATTR((k_backprop_f<C4, false, true>)); AZ_BP_ATTR((k_backprop_f<C4, false, false>));
        AZ_BP_ATTR((k_backprop_f_r<C4, true, true>)); AZ_BP_ATTR((k_backprop_f_r<C4, true, false>));
        AZ_BP_ATTR((k_backprop_f_r<C4, false, true>)); AZ_BP_ATTR((k_backprop_f_r<C4, false, false>));
        AZ_BP_ATTR((k_backprop_f<C4, true, true, true>)); AZ_BP_ATTR((k_backprop_f<C4, false, true, true>));
        AZ_BP_ATTR((k_backprop_f_r<C4, true, true, true>)); AZ_BP_ATTR((k_backprop_f_r<C4, false, true, true>));
#undef AZ_BP_ATTR
        h->bp_smem_set = smem;
    }
}
static int set_range(az_mcts *h, int first, int count) {
    if (first < 0 || count <= 0 || first + count > h->n || (first & 31))
        AZ_FAIL(h, AZ_ERR_INVALID, "tree range [%d, %d) must lie inside [0, %d) and start at a multiple of 32", first, first + count, h->n);
    h->d.env_lo = first; h->d.env_cnt = count;
    return AZ_OK;
}
// Shards: rows of the caller's leaf / policy / value arrays that belong to tree range [first, first + count) start at row0
// (row of tree i, simulation k = row0 + (i - first) * K + k).  The kernels index rows by the global tree number, so the
// base pointers are shifted accordingly.  (row0 = first * K reproduces the whole-batch layout.)
static int do_search(az_mcts *h, int K, const az_root *d_roots, az_leaf *d_leaves, cudaStream_t s, int first = 0, int count = -1,
                     bool new_epoch = true, int64_t row0 = -1) {
    int rc = check_cfg(h, K); if (rc) return rc;
    CU(h, cudaSetDevice(h->device));
    rc = set_range(h, first, count < 0 ? h->n : count); if (rc) return rc;
    if (row0 >= 0) d_leaves += row0 - (int64_t)first * std::max(K, 1);
    const bool vl = K > 0;
    if (vl) { rc = ensure_kcap(h, K); if (rc) return rc; h->prepared_K = K; }
    if (new_epoch) h->d.epoch++;
    h->d.stats = h->stats_on ? h->d_stats : nullptr;
    if (h->time_select) {
        if (h->sel_used == h->sel_ev.size()) {
            cudaEvent_t a, b; CU(h, cudaEventCreate(&a)); CU(h, cudaEventCreate(&b));
            h->sel_ev.push_back({a, b});
        }
        CU(h, cudaEventRecord(h->sel_ev[h->sel_used].first, s));
    }
    rc = launch_select(h, vl, K, d_roots, d_leaves, s); if (rc) return rc;
    if (h->time_select) {
        CU(h, cudaEventRecord(h->sel_ev[h->sel_used].second, s));
        h->sel_used++; h->sel_rows += (uint64_t)h->d.env_cnt * (uint64_t)std::max(K, 1);
    }
    CU(h, cudaGetLastError());
    return AZ_OK;
}
static int do_backprop(az_mcts *h, int K, const float *pol, const float *d, const float *p1, const float *p2, const float *ml,
                       const uint8_t *it, const int32_t *sym, cudaStream_t s, int first = 0, int count = -1, int64_t row0 = -1) {
    int rc = check_cfg(h, K); if (rc) return rc;
    CU(h, cudaSetDevice(h->device));
    rc = set_range(h, first, count < 0 ? h->n : count); if (rc) return rc;
    if (row0 >= 0) {
        const int64_t sh = row0 - (int64_t)first * std::max(K, 1);
        pol += sh * h->A; d += sh; p1 += sh; p2 += sh; ml += sh;
        if (it) it += sh;
        if (sym) sym += sh;
    }
    const bool vl = K > 0;
    if (vl && K > h->prepared_K) AZ_FAIL(h, AZ_ERR_INVALID, "backprop_batch_vl: K (%d) exceeds the K of the last search_batch_vl (%d)", K, h->prepared_K);
    rc = ensure_arena(h, vl ? K : 1, s); if (rc) return rc;
    h->d.stats = h->stats_on ? h->d_stats : nullptr;
    const int removeK = vl ? std::min(K, h->prepared_K) : 0;
    if (h->time_select) {
        if (h->bp_used == h->bp_ev.size()) {
            cudaEvent_t a, b; CU(h, cudaEventCreate(&a)); CU(h, cudaEventCreate(&b));
            h->bp_ev.push_back({a, b});
        }
        CU(h, cudaEventRecord(h->bp_ev[h->bp_used].first, s));
    }
    // non-VL: the symmetry id is the one search_batch remembered (pending_sym_ids_, BatchedMCTS.h:45,194)
    rc = launch_backprop(h, vl, K, removeK, 1, pol, d, p1, p2, ml, it, vl ? sym : nullptr, s); if (rc) return rc;
    if (h->time_select) {
        CU(h, cudaEventRecord(h->bp_ev[h->bp_used].second, s));
        h->bp_used++; h->bp_rows += (uint64_t)h->d.env_cnt * (uint64_t)std::max(K, 1);
    }
    CU(h, cudaGetLastError());
    return AZ_OK;
}
static StagePool &stage_pool_of(az_mcts *h) {
    if (!h->stage_pool) {
        h->stage_pool = new StagePool();       // (threads are created here, on the first host-pipelined playout of this engine)
        const char *e = getenv("AZB200_STAGE_THREADS");
        h->stage_pool->start(std::min(std::max(e ? atoi(e) : 2, 0), 8));
    }
    return *h->stage_pool;
}
static int check_device_error(az_mcts *h) {
    int e = 0;
    CU(h, cudaMemcpyAsync(&e, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (e) AZ_FAIL(h, AZ_ERR_NOMEM, "device tree arena overflow (internal sizing error)");
    return AZ_OK;
}

// Host entry points run on the internal stream: first wait for whatever the caller queued through the *_dev API.
static int enter_host(az_mcts *h) {
    CU(h, cudaSetDevice(h->device));
    if (h->user_pending) {
        CU(h, cudaEventRecord(h->ev, h->user_stream));
        CU(h, cudaStreamWaitEvent(h->stream, h->ev, 0));
        h->user_pending = false;
    }
    return AZ_OK;
}
// *_dev entry points run on the caller's stream: first wait for un-synchronised work of the internal stream.
static int enter_dev(az_mcts *h, cudaStream_t s) {
    CU(h, cudaSetDevice(h->device));
    if (h->internal_pending && s != h->stream) {
        CU(h, cudaEventRecord(h->ev, h->stream));
        CU(h, cudaStreamWaitEvent(s, h->ev, 0));
        h->internal_pending = false;
    }
    if (s != h->stream) { h->user_stream = s; h->user_pending = true; }
    return AZ_OK;
}

template <class G> static void launch_pack(int n, const int8_t *b, const int32_t *t, az_root *r, cudaStream_t s) {
    k_pack_roots<G><<<grid_threads((size_t)n), 128, 0, s>>>(n, b, t, r);
}
template <class G> static void launch_unpack(int rows, const az_leaf *l, int8_t *ob, float *td, float *tp1, float *tp2, uint8_t *it, int32_t *ot,
                                             int32_t *sym, uint8_t *vm, float *planes, cudaStream_t s) {
    k_unpack_leaves<G><<<grid_threads((size_t)rows * G::S, 256), 256, 0, s>>>(rows, l, ob, td, tp1, tp2, it, ot, sym, vm, planes);
}

extern "C" {

const char *az_version(void) { return "azb200 0.2 (sm_100a)"; }
const char *az_global_last_error(void) { return g_global_err.c_str(); }
int az_game_action_size(int g) { return g == GAME_C4 ? C4::A : (g == GAME_OTH ? Oth::A : -1); }
int az_game_board_size(int g) { return g == GAME_C4 ? C4::S : (g == GAME_OTH ? Oth::S : -1); }
int az_game_board_rows(int g) { return g == GAME_C4 ? 6 : (g == GAME_OTH ? 8 : -1); }
int az_game_board_cols(int g) { return g == GAME_C4 ? 7 : (g == GAME_OTH ? 8 : -1); }
int az_game_num_symmetries(int g) { return g == GAME_C4 ? 2 : (g == GAME_OTH ? 8 : -1); }
void az_search_config_defaults(az_search_config *c) {
    c->c_init = 1.25f; c->c_base = 19652.0f; c->dirichlet_alpha = 0.3f; c->noise_epsilon = 0.25f; c->fpu_reduction = 0.4f;
    c->mlh_slope = 0.0f; c->mlh_cap = 0.2f; c->score_utility_factor = 0.0f; c->score_scale = 8.0f; c->value_decay = 1.0f;
    c->use_symmetry = 1; c->vl_count = 1;
}

int az_pack_roots_dev(int game, int n, const int8_t *b, const int32_t *t, az_root *r, void *stream) {
    if (n <= 0) return AZ_OK;
    if (game == GAME_C4) launch_pack<C4>(n, b, t, r, (cudaStream_t)stream);
    else if (game == GAME_OTH) launch_pack<Oth>(n, b, t, r, (cudaStream_t)stream);
    else return AZ_ERR_INVALID;
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}
int az_unpack_leaves_dev(int game, int rows, const az_leaf *l, int8_t *ob, float *td, float *tp1, float *tp2, uint8_t *it, int32_t *ot,
                         int32_t *sym, uint8_t *vm, float *planes, void *stream) {
    if (rows <= 0) return AZ_OK;
    if (game == GAME_C4) launch_unpack<C4>(rows, l, ob, td, tp1, tp2, it, ot, sym, vm, planes, (cudaStream_t)stream);
    else if (game == GAME_OTH) launch_unpack<Oth>(rows, l, ob, td, tp1, tp2, it, ot, sym, vm, planes, (cudaStream_t)stream);
    else return AZ_ERR_INVALID;
    return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA;
}

az_mcts *az_mcts_create(int game, int n_envs, int device) {
    if (game != GAME_C4 && game != GAME_OTH) { g_global_err = "unknown game id"; return nullptr; }
    if (n_envs <= 0) { g_global_err = "n_envs must be positive"; return nullptr; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_global_err = std::string("no CUDA device available (the MCTS engine has no CPU fallback): ") + cudaGetErrorString(e);
        return nullptr;
    }
    if (device < 0 || device >= ndev) { g_global_err = "invalid CUDA device ordinal"; return nullptr; }
    az_mcts *h = new az_mcts();
    h->game = game; h->n = n_envs; h->device = device;
    h->A = az_game_action_size(game); h->S = az_game_board_size(game);
    h->W = auto_lanes(game, n_envs);
    { const char *ve = getenv("AZB200_VARIANT"); if (ve) { int v = atoi(ve); if (v >= 0 && v <= 1) h->variant = v; } }
    { const char *we = getenv("AZB200_WAVE_MAX"); if (we) h->wave_max = std::max(0, atoi(we)); }
    { const char *ge = getenv("AZB200_GRAPHS"); if (ge) h->use_graphs = atoi(ge) != 0; }
    { const char *pe = getenv("AZB200_PDL"); if (pe) h->pdl = atoi(pe) != 0; }
    { const char *ce2 = getenv("AZB200_COMPACTION"); if (ce2) { int v = atoi(ce2); if (v >= 0 && v <= 2) h->compaction = v; } }
    h->max_depth = game == GAME_C4 ? C4::MAX_DEPTH : Oth::MAX_DEPTH;
    h->max_edges = game == GAME_C4 ? 8 : 48;
    az_search_config_defaults(&h->cfg);
    auto fail = [&](const char *what) -> az_mcts * {
        g_global_err = std::string(what) + ": " + h->err;
        az_mcts_destroy(h);
        return nullptr;
    };
    if (cudaSetDevice(device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return fail("create"); }
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { h->err = "stream create failed"; return fail("create"); }
    if (cudaEventCreateWithFlags(&h->ev, cudaEventDisableTiming) != cudaSuccess) { h->err = "event create failed"; return fail("create"); }
    const char *ce = getenv("AZB200_ARENA_SLOTS");
    h->cap = ce ? (uint32_t)std::max(256, atoi(ce)) : (game == GAME_C4 ? 2048u : 4096u);
    h->d.n_envs = n_envs; h->d.env_lo = 0; h->d.env_cnt = n_envs; h->d.cap = h->cap; h->d.noise_stride = h->max_edges;
    { const char *he = getenv("AZB200_HINTS"); h->d.hints = he ? atoi(he) : 1; }
    { const char *le = getenv("AZB200_LAZY"); if (le) h->d.hints = atoi(le) ? (h->d.hints | 2) : (h->d.hints & ~2); }   // lazy blocks: off by default (measured slower)
    h->d.seed = 0x243F6A8885A308D3ULL; h->d.epoch = 0; h->d.epoch_add = nullptr;
    int rc = 0;
    rc |= dev_alloc(h, &h->d.pool, (size_t)n_envs * h->cap);
    rc |= dev_alloc(h, &h->d.trees, (size_t)n_envs);
    rc |= dev_alloc(h, &h->d.noise, (size_t)n_envs * h->d.noise_stride);
    rc |= dev_alloc(h, &h->d.leaf_nv, (size_t)n_envs);
    rc |= dev_alloc(h, &h->d.path_nv, (size_t)n_envs * h->max_depth);
    rc |= dev_alloc(h, &h->d_stats, 8 + 2 * AZ_DBG_WARPS); rc |= dev_alloc(h, &h->d_err, 1); rc |= dev_alloc(h, &h->d_scratch_u32, 4);
    rc |= dev_alloc(h, &h->io_boards_in, (size_t)n_envs * h->S); rc |= dev_alloc(h, &h->io_turns_in, (size_t)n_envs);
    rc |= dev_alloc(h, &h->io_roots, (size_t)n_envs);
    rc |= dev_alloc(h, &h->io_actions, (size_t)n_envs); rc |= dev_alloc(h, &h->io_counts, (size_t)n_envs * h->A);
    rc |= dev_alloc(h, &h->io_stats, (size_t)n_envs * (6 + 8 * h->A));
    if (rc) return fail("device allocation");
    h->d.err = h->d_err;
    cudaMemsetAsync(h->d_stats, 0, (8 + 2 * AZ_DBG_WARPS) * sizeof(unsigned long long), h->stream);
    cudaMemsetAsync(h->d_err, 0, sizeof(int), h->stream);
    // TreeRec.noise_ctr (the per-tree Dirichlet draw counter) is not touched by a reset: it must start from 0, not from whatever
    // cudaMalloc handed out, or two engines with the same seed draw different noise (found by the sharding-invariance test)
    cudaMemsetAsync(h->d.trees, 0, sizeof(TreeRec) * (size_t)n_envs, h->stream);
    k_reset<<<grid_threads((size_t)n_envs), 128, 0, h->stream>>>(h->d, -1);
    k_init_leaf<<<grid_threads((size_t)n_envs), 128, 0, h->stream>>>(h->d.leaf_nv, (size_t)n_envs);
    if (ensure_io(h, n_envs) != AZ_OK) return fail("io allocation");
    if (ensure_luts(h) != AZ_OK) return fail("LUT upload");
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) { h->err = cudaGetErrorString(cudaGetLastError()); return fail("init kernels"); }
    return h;
}

void az_mcts_destroy(az_mcts *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    void *ptrs[] = {h->d.pool, h->pool_alt, h->d.trees, h->d.noise, h->d.leaf_vl, h->d.path_vl, h->d.leaf_nv, h->d.path_nv, h->d_log_lut, h->d_ls_lut,
                    h->d_atan_lut, h->d_stats, h->d_err, h->d_scratch_u32, h->io_boards_in, h->io_turns_in, h->io_roots, h->io_leaves,
                    h->io_out, h->io_in, h->io_actions, h->io_counts, h->io_stats};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->h_in) cudaFreeHost(h->h_in);
    if (h->h_in2) cudaFreeHost(h->h_in2);
    for (int j = 0; j < 2; ++j) if (h->h_in_ev[j]) cudaEventDestroy(h->h_in_ev[j]);
    if (h->h_counts) cudaFreeHost(h->h_counts);
    delete h->stage_pool;
    if (h->h_stage) cudaFreeHost(h->h_stage);
    if (h->io_stage) cudaFree(h->io_stage);
    if (h->h_err16) cudaFreeHost(h->h_err16);
    if (h->io_counts64) cudaFree(h->io_counts64);
    if (h->counts_block >= 0) az_pinned_release(h->counts_block);
    if (h->ev) cudaEventDestroy(h->ev);
    for (auto &pr : h->sel_ev) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto &pr : h->bp_ev) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (auto &g : h->graphs) { if (g.exec) cudaGraphExecDestroy(g.exec); for (auto e : g.shard_exec) if (e) cudaGraphExecDestroy(e); }
    if (h->d_epoch_add) cudaFree(h->d_epoch_add);
    for (int j = 0; j < 16; ++j) { if (h->side_ev[j]) cudaEventDestroy(h->side_ev[j]); if (h->side[j]) cudaStreamDestroy(h->side[j]); }
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}
const char *az_mcts_last_error(const az_mcts *h) { return h ? h->err.c_str() : g_global_err.c_str(); }
int az_mcts_num_envs(const az_mcts *h) { return h->n; }
int az_mcts_set_config(az_mcts *h, const az_search_config *c) { h->cfg = *c; return AZ_OK; }
int az_mcts_get_config(const az_mcts *h, az_search_config *c) { *c = h->cfg; return AZ_OK; }
int az_mcts_set_lanes(az_mcts *h, int lanes) {
    if (lanes == 0) { h->W = auto_lanes(h->game, h->n); return AZ_OK; }
    if (h->game == GAME_OTH) {
        if (lanes != 8 && lanes != 16) AZ_FAIL(h, AZ_ERR_INVALID, "Othello lanes must be 8 or 16");
        h->W = lanes;
        return AZ_OK;
    }
    if (lanes != 1 && lanes != 2 && lanes != 4 && lanes != 8) AZ_FAIL(h, AZ_ERR_INVALID, "Connect4 lanes must be 1, 2, 4 or 8");
    h->W = lanes;
    return AZ_OK;
}
int az_mcts_get_lanes(const az_mcts *h) { return h->W; }
int az_mcts_set_variant(az_mcts *h, int variant) {
    if (variant < 0 || variant > 1) AZ_FAIL(h, AZ_ERR_INVALID, "kernel variant must be 0 or 1");
    h->variant = variant;
    return AZ_OK;
}
int az_mcts_get_variant(const az_mcts *h) { return h->variant; }
int az_mcts_set_wave_max(az_mcts *h, int max_lanes) {
    if (max_lanes < 0) AZ_FAIL(h, AZ_ERR_INVALID, "wave_max must be >= 0");
    h->wave_max = max_lanes;
    return AZ_OK;
}
int az_mcts_get_wave_max(const az_mcts *h) { return h->wave_max; }
int az_selftest_div(int mode, uint64_t count, uint64_t seed, uint64_t *mismatches) {
    unsigned long long *dm = nullptr, hm = 0;
    if (mode < 0 || mode > 2 || !mismatches) return AZ_ERR_INVALID;
    if (cudaMalloc((void **)&dm, sizeof(hm)) != cudaSuccess) return AZ_ERR_CUDA;
    cudaMemset(dm, 0, sizeof(hm));
    k_selftest_div<<<148 * 8, 256>>>(mode, (unsigned long long)count, (unsigned long long)seed, dm);
    cudaError_t e = cudaMemcpy(&hm, dm, sizeof(hm), cudaMemcpyDeviceToHost);
    cudaFree(dm);
    if (e != cudaSuccess) return AZ_ERR_CUDA;
    *mismatches = hm;
    return AZ_OK;
}
int az_mcts_set_env_base(az_mcts *h, uint64_t base) { h->d.env_base = base; return AZ_OK; }
int az_mcts_reserve(az_mcts *h, int slots_per_tree) {
    CU(h, cudaSetDevice(h->device));
    if (slots_per_tree <= 0) AZ_FAIL(h, AZ_ERR_INVALID, "slots_per_tree must be positive");
    uint64_t ncap = h->cap;
    while (ncap < (uint64_t)slots_per_tree) ncap *= 2;
    if (ncap > h->cap) return grow_arena(h, ncap, h->stream);
    return AZ_OK;
}
int az_mcts_set_seed(az_mcts *h, int64_t seed) {
    if (seed < 0) {   // re-randomise (BatchedMCTS.h:73-76)
        uint64_t t = (uint64_t)clock() ^ ((uint64_t)(uintptr_t)h << 16);
        h->d.seed = splitmix64(t ^ splitmix64(h->d.seed));
    } else h->d.seed = (uint64_t)seed;
    h->d.epoch = 0;
    return AZ_OK;
}
int az_mcts_reset_env(az_mcts *h, int i) {
    if (i < 0 || i >= h->n) return AZ_OK;   // silently ignored (BatchedMCTS.h:93-99)
    { int rc = enter_host(h); if (rc) return rc; }
    h->tree_gen++;
    k_reset<<<1, 32, 0, h->stream>>>(h->d, i);
    h->internal_pending = true;
    h->launches++;
    CU(h, cudaGetLastError());
    return AZ_OK;
}
int az_mcts_prune_roots_dev(az_mcts *h, const int32_t *d_actions, void *stream) {
    int rc = check_cfg(h, 1); if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    rc = enter_dev(h, s); if (rc) return rc;
    h->tree_gen++;
    if (h->game == GAME_C4) k_prune<C4><<<grid_threads((size_t)h->n), 128, 0, s>>>(h->d, h->cfg, d_actions);
    else k_prune<Oth><<<grid_threads((size_t)h->n), 128, 0, s>>>(h->d, h->cfg, d_actions);
    h->launches++;
    CU(h, cudaGetLastError());
    // compaction: copy the surviving subtrees into the other pool when the arenas are filling up
    // ... i.e. when one more move's growth (estimated from the growth since the previous re-root) might not fit any more
    uint64_t bound = h->bump_bound;
    for (const auto &r : h->bounds) bound = std::max(bound, r.b);
    const uint64_t growth = bound > h->base_after_prune ? bound - h->base_after_prune : 0;
    h->growth_est = std::max(h->growth_est - h->growth_est / 4, growth);
    const bool due = h->compaction == 2 || (h->compaction == 1 && bound + h->growth_est + h->growth_est / 4 > h->cap && bound > 0);
    h->base_after_prune = bound;
    if (due) { rc = run_compaction(h, s, false); if (rc) return rc; }
    return AZ_OK;
}
// Every tree back to a fresh root (prune_roots with all actions < 0, MCTS.h:107 -> reset()), stream-ordered, and - unlike
// az_mcts_prune_roots_dev, which cannot see the device array - with the host's arena bookkeeping reset too: the arenas are empty,
// so the next playout needs no bound refresh (no device synchronisation between a reset and a search).
int az_mcts_reset_all_dev(az_mcts *h, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    int rc = enter_dev(h, s); if (rc) return rc;
    h->tree_gen++;
    k_reset<<<grid_threads((size_t)h->n), 128, 0, s>>>(h->d, -1);
    h->launches++;
    CU(h, cudaGetLastError());
    h->bump_bound = 0; h->bounds.clear(); h->base_after_prune = 0; h->bound_stale = false; h->base_pending = false; h->growth_est = 0;
    h->sel_recs.clear(); h->lazy_live = false;
    return AZ_OK;
}
int az_mcts_set_lazy(az_mcts *h, int on) {
    if (on) { h->d.hints |= 2; return AZ_OK; }
    if (h->lazy_live) { CU(h, cudaSetDevice(h->device)); int rc = materialise_all(h, h->stream); if (rc) return rc; CU(h, cudaStreamSynchronize(h->stream)); }
    h->d.hints &= ~2;
    return AZ_OK;
}
int az_mcts_get_lazy(const az_mcts *h) { return (h->d.hints & 2) ? 1 : 0; }
int az_mcts_set_compaction(az_mcts *h, int mode) {
    if (mode < 0 || mode > 2) AZ_FAIL(h, AZ_ERR_INVALID, "compaction mode must be 0 (never), 1 (auto) or 2 (every re-root)");
    h->compaction = mode;
    return AZ_OK;
}
uint64_t az_mcts_compactions(const az_mcts *h) { return h->compactions; }
int az_mcts_prune_roots(az_mcts *h, const int32_t *actions) {
    { int rc = enter_host(h); if (rc) return rc; }
    // no action can match an edge (all negative): every tree is reset (MCTS.h:107), the arenas are empty afterwards - nothing to
    // compact, the host's arena bound is exactly 0, and the actions need not travel at all
    int32_t signs = -1;
    for (int i = 0; i < h->n; ++i) signs &= actions[i];           // (no early exit: vectorised)
    if (signs < 0) {
        int rc = check_cfg(h, 1); if (rc) return rc;
        rc = az_mcts_reset_all_dev(h, h->stream); if (rc) return rc;
        h->internal_pending = true;
        return AZ_OK;
    }
    CU(h, cudaMemcpyAsync(h->io_actions, actions, sizeof(int32_t) * (size_t)h->n, cudaMemcpyHostToDevice, h->stream));
    int rc = az_mcts_prune_roots_dev(h, h->io_actions, h->stream);
    if (rc) return rc;
    // no synchronisation: `actions` (pageable) has been staged when cudaMemcpyAsync returns, io_actions is only touched on this
    // stream, and every later entry point is ordered after this one (the re-root overlaps the caller's next host work)
    h->internal_pending = true;
    return AZ_OK;
}

int az_pinned_release(int block);
// host entry points: pack -> search -> unpack into one packed staging buffer -> ONE D2H copy into pinned memory.
// `pinned` != NULL: the copy lands in a pinned block of the process-wide pool that the caller owns until az_pinned_release - its
// arrays ARE the copy's destination (no second pass over 2.3 MB per K = 4 iteration at 8192 games).  Otherwise the arrays the caller
// passed are filled from the engine's own pinned mirror.
static int host_search(az_mcts *h, int K, const int8_t *boards, const int32_t *turns, int8_t *ob, float *td, float *tp1, float *tp2,
                       uint8_t *it, int32_t *ot, int32_t *sym, uint8_t *vm, az_host_leaves *pinned = nullptr) {
    { int rc0 = enter_host(h); if (rc0) return rc0; }
    const int rowsK = K > 0 ? K : 1;
    const size_t rows = (size_t)h->n * rowsK;
    int rc = ensure_io(h, (int)rows); if (rc) return rc;
    cudaStream_t s = h->stream;
    // the wrapper hands the same boards to every iteration of a move (src/MCTS_cpp.py:217-357: 51 calls at n = 200, K = 4): the
    // packed roots of the previous call are still on the device when the bytes are unchanged (a 344 KB compare instead of a
    // pageable host-to-device copy + pack kernel)
    const size_t bbytes = (size_t)h->n * h->S, tbytes = sizeof(int32_t) * (size_t)h->n;
    const bool same = h->roots_valid && h->last_boards.size() == bbytes && memcmp(h->last_boards.data(), boards, bbytes) == 0 &&
                      memcmp(h->last_turns.data(), turns, tbytes) == 0;
    if (!same) {
        CU(h, cudaMemcpyAsync(h->io_boards_in, boards, bbytes, cudaMemcpyHostToDevice, s));
        CU(h, cudaMemcpyAsync(h->io_turns_in, turns, tbytes, cudaMemcpyHostToDevice, s));
        rc = az_pack_roots_dev(h->game, h->n, h->io_boards_in, h->io_turns_in, h->io_roots, s); if (rc) AZ_FAIL(h, rc, "pack_roots launch failed");
        h->last_boards.assign(boards, boards + bbytes); h->last_turns.assign(turns, turns + h->n); h->roots_valid = true;
    }
    rc = do_search(h, K, h->io_roots, h->io_leaves, s); if (rc) return rc;
    const OutLayout L = out_layout(rows, h->S, h->A);
    uint8_t *o = h->io_out;
    rc = az_unpack_leaves_dev(h->game, (int)rows, h->io_leaves, (int8_t *)(o + L.boards), (float *)(o + L.td), (float *)(o + L.tp1),
                              (float *)(o + L.tp2), o + L.term, (int32_t *)(o + L.turns), (int32_t *)(o + L.sym), o + L.mask, nullptr, s);
    if (rc) AZ_FAIL(h, rc, "unpack_leaves launch failed");
    h->launches += 2;
    uint8_t *dst = h->h_out;
    int block = -1;
    if (pinned) {
        void *bp = nullptr;
        block = pinned_acquire(L.total, &bp);
        if (block < 0) AZ_FAIL(h, AZ_ERR_NOMEM, "cannot allocate %zu bytes of pinned host memory for the leaf arrays", L.total);
        dst = (uint8_t *)bp;
    }
    int dev_err = 0;
    CU(h, cudaMemcpyAsync(dst, o, L.total, cudaMemcpyDeviceToHost, s));
    if (h->err_check_pending) CU(h, cudaMemcpyAsync(&dev_err, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    cudaError_t se = cudaStreamSynchronize(s);
    if (se != cudaSuccess || dev_err) {
        if (block >= 0) az_pinned_release(block);
        if (se != cudaSuccess) AZ_FAIL(h, AZ_ERR_CUDA, "cudaStreamSynchronize failed: %s", cudaGetErrorString(se));
        AZ_FAIL(h, AZ_ERR_NOMEM, "device tree arena overflow (internal sizing error)");
    }
    h->err_check_pending = false;
    const uint8_t *p = dst;
    if (pinned) {
        pinned->block = block; pinned->rows = (int32_t)rows;
        pinned->boards = (int8_t *)(dst + L.boards); pinned->term_d = (float *)(dst + L.td); pinned->term_p1w = (float *)(dst + L.tp1);
        pinned->term_p2w = (float *)(dst + L.tp2); pinned->is_term = dst + L.term; pinned->turns = (int32_t *)(dst + L.turns);
        pinned->sym_ids = (int32_t *)(dst + L.sym); pinned->valid_mask = dst + L.mask;
        return AZ_OK;
    }
    memcpy(ob, p + L.boards, rows * h->S); memcpy(td, p + L.td, rows * 4); memcpy(tp1, p + L.tp1, rows * 4); memcpy(tp2, p + L.tp2, rows * 4);
    memcpy(it, p + L.term, rows); memcpy(ot, p + L.turns, rows * 4); memcpy(vm, p + L.mask, rows * h->A);
    if (sym) memcpy(sym, p + L.sym, rows * 4);
    return AZ_OK;
}
// The evaluation tuple is packed into one of two pinned buffers, copied with ONE host-to-device copy, and the back-prop kernel is
// queued behind it: the call returns without waiting for either (the next host entry point is stream-ordered after them; the
// device error flag is read back by the next search).  The reference's call is synchronous, but nothing a caller can observe
// through the API happens before the next entry point anyway.
static int host_backprop(az_mcts *h, int K, const float *pol, const float *d, const float *p1, const float *p2, const float *ml,
                         const uint8_t *it, const int32_t *sym) {
    { int rc0 = enter_host(h); if (rc0) return rc0; }
    const int rowsK = K > 0 ? K : 1;
    const size_t rows = (size_t)h->n * rowsK;
    int rc = ensure_io(h, (int)rows); if (rc) return rc;
    cudaStream_t s = h->stream;
    const InLayout L = in_layout(rows, h->A);
    uint8_t *q = h->io_in;
    {   // (copying straight from the caller's pageable arrays - one cudaMemcpyAsync per array, staged by the driver - measured the same: 15.7 ms per step)
        const int sel = h->h_in_sel; h->h_in_sel ^= 1;
        uint8_t *p = sel ? h->h_in2 : h->h_in;
        CU(h, cudaEventSynchronize(h->h_in_ev[sel]));          // the copy that last read this buffer has completed
        // streaming stores: the pinned buffer is only read by the copy engine (0.115-0.118 ms per 8192-game K = 4 call; handing
        // pieces to the staging threads as well made the call itself faster, 0.085-0.103, but the caller's numpy evaluator and
        // the next search slower by more than that in three A/B runs - not used here)
        stage_copy(p + L.policy, pol, rows * h->A * 4); stage_copy(p + L.d, d, rows * 4); stage_copy(p + L.p1, p1, rows * 4);
        stage_copy(p + L.p2, p2, rows * 4); stage_copy(p + L.ml, ml, rows * 4); stage_copy(p + L.term, it, rows);
        if (sym) stage_copy(p + L.sym, sym, rows * 4);
        CU(h, cudaMemcpyAsync(h->io_in, p, L.total, cudaMemcpyHostToDevice, s));
        CU(h, cudaEventRecord(h->h_in_ev[sel], s));
    }
    rc = do_backprop(h, K, (const float *)(q + L.policy), (const float *)(q + L.d), (const float *)(q + L.p1), (const float *)(q + L.p2),
                     (const float *)(q + L.ml), q + L.term, sym ? (const int32_t *)(q + L.sym) : nullptr, s);
    if (rc) return rc;
    h->err_check_pending = true; h->internal_pending = true;
    return AZ_OK;
}

int az_mcts_search_batch(az_mcts *h, const int8_t *b, const int32_t *t, int8_t *ob, float *td, float *tp1, float *tp2, uint8_t *it,
                         int32_t *ot, uint8_t *vm) {
    return host_search(h, 0, b, t, ob, td, tp1, tp2, it, ot, nullptr, vm);
}
int az_mcts_backprop_batch(az_mcts *h, const float *pol, const float *d, const float *p1, const float *p2, const float *ml, const uint8_t *it) {
    return host_backprop(h, 0, pol, d, p1, p2, ml, it, nullptr);
}
int az_mcts_search_batch_vl(az_mcts *h, int K, const int8_t *b, const int32_t *t, int8_t *ob, float *td, float *tp1, float *tp2,
                            uint8_t *it, int32_t *ot, int32_t *sym, uint8_t *vm) {
    if (K < 1) AZ_FAIL(h, AZ_ERR_INVALID, "search_batch_vl: K must be >= 1");
    return host_search(h, K, b, t, ob, td, tp1, tp2, it, ot, sym, vm);
}
int az_mcts_backprop_batch_vl(az_mcts *h, int K, const float *pol, const float *d, const float *p1, const float *p2, const float *ml,
                              const uint8_t *it, const int32_t *sym) {
    if (K < 1) AZ_FAIL(h, AZ_ERR_INVALID, "backprop_batch_vl: K must be >= 1");
    return host_backprop(h, K, pol, d, p1, p2, ml, it, sym);
}
int az_mcts_search_batch_pinned(az_mcts *h, int K, const int8_t *b, const int32_t *t, az_host_leaves *out) {
    if (K < 0 || !out) AZ_FAIL(h, AZ_ERR_INVALID, "search_batch_pinned: K must be >= 0 and out non-NULL");
    return host_search(h, K, b, t, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, out);
}
int az_pinned_release(int block) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    if (block < 0 || (size_t)block >= g_pin.size() || !g_pin[(size_t)block].in_use) return AZ_ERR_INVALID;
    g_pin[(size_t)block].in_use = false;
    size_t idle = 0;                                         // keep a few idle blocks for reuse, give the rest back
    for (auto &b : g_pin) if (b.ptr && !b.in_use) ++idle;
    if (idle > 6) { cudaFreeHost(g_pin[(size_t)block].ptr); g_pin[(size_t)block].ptr = nullptr; g_pin[(size_t)block].bytes = 0; }
    return AZ_OK;
}
int az_mcts_remove_all_vl(az_mcts *h, int K) {
    int rc = check_cfg(h, 1); if (rc) return rc;
    rc = enter_host(h); if (rc) return rc;
    const int safeK = std::min(K, h->prepared_K);
    if (safeK <= 0) return AZ_OK;
    const int g = grid_groups(h->n, h->W);
    if (h->game == GAME_OTH) {
        if (h->W == 8) k_remove_vl<Oth, 8><<<g, CTA, 0, h->stream>>>(h->d, h->cfg, safeK);
        else k_remove_vl<Oth, 16><<<g, CTA, 0, h->stream>>>(h->d, h->cfg, safeK);
    } else switch (h->W) {
        case 1: k_remove_vl<C4, 1><<<g, CTA, 0, h->stream>>>(h->d, h->cfg, safeK); break;
        case 2: k_remove_vl<C4, 2><<<g, CTA, 0, h->stream>>>(h->d, h->cfg, safeK); break;
        case 4: k_remove_vl<C4, 4><<<g, CTA, 0, h->stream>>>(h->d, h->cfg, safeK); break;
        default: k_remove_vl<C4, 8><<<g, CTA, 0, h->stream>>>(h->d, h->cfg, safeK); break;
    }
    h->launches++;
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(h->stream));
    return AZ_OK;
}

int az_mcts_search_dev(az_mcts *h, int K, const az_root *d_roots, az_leaf *d_leaves, void *stream) {
    if (K < 0) AZ_FAIL(h, AZ_ERR_INVALID, "search_dev: K must be >= 0");
    { int rc = enter_dev(h, (cudaStream_t)stream); if (rc) return rc; }
    return do_search(h, K, d_roots, d_leaves, (cudaStream_t)stream);
}
int az_mcts_backprop_dev(az_mcts *h, int K, const float *pol, const float *d, const float *p1, const float *p2, const float *ml,
                         const uint8_t *it, const int32_t *sym, void *stream) {
    if (K < 0) AZ_FAIL(h, AZ_ERR_INVALID, "backprop_dev: K must be >= 0");
    { int rc = enter_dev(h, (cudaStream_t)stream); if (rc) return rc; }
    return do_backprop(h, K, pol, d, p1, p2, ml, it, sym, (cudaStream_t)stream);
}

int az_mcts_search_range_dev(az_mcts *h, int K, const az_root *d_roots, az_leaf *d_leaves, int first, int count, int64_t row0, int new_epoch,
                             void *stream) {
    if (K < 0 || row0 < 0) AZ_FAIL(h, AZ_ERR_INVALID, "search_range_dev: K and row0 must be >= 0");
    { int rc = enter_dev(h, (cudaStream_t)stream); if (rc) return rc; }
    return do_search(h, K, d_roots, d_leaves, (cudaStream_t)stream, first, count, new_epoch != 0, row0);
}
int az_mcts_backprop_range_dev(az_mcts *h, int K, const float *pol, const float *d, const float *p1, const float *p2, const float *ml,
                               const uint8_t *it, const int32_t *sym, int first, int count, int64_t row0, void *stream) {
    if (K < 0 || row0 < 0) AZ_FAIL(h, AZ_ERR_INVALID, "backprop_range_dev: K and row0 must be >= 0");
    { int rc = enter_dev(h, (cudaStream_t)stream); if (rc) return rc; }
    return do_backprop(h, K, pol, d, p1, p2, ml, it, sym, (cudaStream_t)stream, first, count, row0);
}
int az_mcts_stream_handover_dev(az_mcts *h, void *stream) { return enter_dev(h, (cudaStream_t)stream); }

// The whole playout loop of src/MCTS_cpp.py:217-357 with a synthetic evaluator, driven natively (no per-launch Python
// cost) and optionally pipelined over `shards` independent tree ranges on internal streams.
static int playout_issue(az_mcts *h, int mode, const std::vector<int> &iters, int ns, int per, const az_root *d_roots, az_leaf *d_leaves,
                         float *pol, float *d, float *p1, float *p2, float *ml, cudaStream_t main, int *launches_out) {
    cudaStream_t lanes[16];
    if (ns > 1) {
        for (int j = 0; j < ns; ++j) lanes[j] = h->side[j];
        CU(h, cudaEventRecord(h->ev, main));
        for (int j = 0; j < ns; ++j) CU(h, cudaStreamWaitEvent(lanes[j], h->ev, 0));
    } else lanes[0] = main;
    int launches = 0, rc;
    // every shard keeps its own rows of the leaf / policy / value arrays for the whole loop ([lo * kmax, (lo + cnt) * kmax)):
    // shards run ahead of each other, and with the whole-batch layout (row = tree * K + k) the rows of one shard's K = 4
    // iteration would overlap another shard's K = 1 or remainder iteration
    int kmax = 1;
    for (int k : iters) kmax = std::max(kmax, k);
    for (size_t it = 0; it < iters.size(); ++it) {
        const int k = iters[it], kk = std::max(k, 1);
        for (int j = 0; j < ns; ++j) {
            const int lo = j * per, cnt = std::min(per, h->n - lo);
            const size_t r0 = (size_t)lo * kmax;
            rc = do_search(h, k, d_roots, d_leaves, lanes[j], lo, cnt, j == 0, (int64_t)r0); if (rc) return rc;
            rc = az_eval_synthetic_dev(h->game, mode, cnt * kk, d_leaves + r0, pol + r0 * h->A, d + r0, p1 + r0, p2 + r0, ml + r0, lanes[j]);
            if (rc) AZ_FAIL(h, rc, "synthetic evaluator launch failed");
            rc = do_backprop(h, k, pol, d, p1, p2, ml, nullptr, nullptr, lanes[j], lo, cnt, (int64_t)r0); if (rc) return rc;
            launches += 3;
        }
    }
    if (ns > 1) {
        for (int j = 0; j < ns; ++j) { CU(h, cudaEventRecord(h->side_ev[j], lanes[j])); CU(h, cudaStreamWaitEvent(main, h->side_ev[j], 0)); }
    }
    *launches_out = launches;
    return AZ_OK;
}

// The whole playout loop of src/MCTS_cpp.py:217-357 with a synthetic evaluator, driven natively (no per-launch Python
// cost), optionally pipelined over `shards` independent tree ranges on internal streams, and replayed from a CUDA graph
// when the same loop was issued before (same buffers, configuration and arena pool).
int az_mcts_playout_synthetic_dev(az_mcts *h, int mode, int n_playout, int K, int shards, const az_root *d_roots, az_leaf *d_leaves,
                                  float *pol, float *d, float *p1, float *p2, float *ml, void *stream, int *launches_out) {
    if (n_playout < 0 || K < 0) AZ_FAIL(h, AZ_ERR_INVALID, "playout: n_playout and K must be >= 0");
    cudaStream_t main = (cudaStream_t)stream;
    int rc = enter_dev(h, main); if (rc) return rc;
    shards = std::min(std::max(shards, 1), 16);             // side[] / lanes[] hold 16 streams: more shards than that are merged
    const int per = (((h->n + shards - 1) / shards) + 31) / 32 * 32;
    const int ns = std::max(1, (h->n + per - 1) / per);
    std::vector<int> iters;
    if (K <= 1) iters.assign((size_t)n_playout, 0);
    else if (n_playout > 0) {
        iters.push_back(0);
        for (int rem = n_playout - 1; rem > 0; rem -= std::min(K, rem)) iters.push_back(std::min(K, rem));
    }
    if (ns > 1)
        for (int j = 0; j < ns; ++j)
            if (!h->side[j]) { CU(h, cudaStreamCreateWithFlags(&h->side[j], cudaStreamNonBlocking)); CU(h, cudaEventCreateWithFlags(&h->side_ev[j], cudaEventDisableTiming)); }
    int launches = 0;
    const bool graphable = h->use_graphs && !h->time_select && !h->stats_on && iters.size() >= 4;
    if (!graphable) {
        rc = playout_issue(h, mode, iters, ns, per, d_roots, d_leaves, pol, d, p1, p2, ml, main, &launches); if (rc) return rc;
    } else {
        rc = check_cfg(h, K); if (rc) return rc;
        if (K > 0) { rc = ensure_kcap(h, K); if (rc) return rc; }
        for (int k : iters) bp_prepare(h, k > 0, k);
        // the legacy default stream cannot be captured: run on the internal stream, ordered after / before the caller's
        cudaStream_t run = main;
        if (!main) {
            run = h->stream;
            CU(h, cudaEventRecord(h->ev, main));
            CU(h, cudaStreamWaitEvent(run, h->ev, 0));
        }
        if (!h->d_epoch_add) { CU(h, cudaMalloc((void **)&h->d_epoch_add, sizeof(unsigned long long))); }
        // reserve the growth of the whole loop up front (may synchronise / grow the arenas: before any capture)
        h->d.env_lo = 0; h->d.env_cnt = h->n;
        rc = ensure_arena(h, n_playout, run); if (rc) return rc;
        if (h->lazy_live) {                                  // (cannot happen inside the capture)
            bool all_aware = true;
            for (int k : iters) all_aware = all_aware && select_lazy_aware(h, k > 0, std::max(k, 1), d_leaves);
            if (!all_aware) { rc = materialise_all(h, run); if (rc) return rc; }
        }
        az_mcts::GraphKey key;
        memset(&key, 0, sizeof(key));
        key.mode = mode; key.n_playout = n_playout; key.K = K; key.ns = ns; key.W = h->W; key.variant = h->variant; key.wave_max = h->wave_max; key.hints = h->d.hints;
        key.kcap = h->kcap; key.cap = h->cap; key.cfg = h->cfg; key.seed = h->d.seed; key.env_base = h->d.env_base;
        const void *pp[10] = {d_roots, d_leaves, pol, d, p1, p2, ml, h->d.pool, h->d.leaf_vl, run};
        memcpy(key.ptrs, pp, sizeof(pp));
        az_mcts::GraphEntry *ge = nullptr;
        for (auto &g : h->graphs) if (memcmp(&g.key, &key, sizeof(key)) == 0) { ge = &g; break; }
        if (!ge) {
            if (h->graphs.size() >= 6) {                      // drop the least recently used
                size_t v = 0;
                for (size_t i = 1; i < h->graphs.size(); ++i) if (h->graphs[i].last_use < h->graphs[v].last_use) v = i;
                if (h->graphs[v].exec) cudaGraphExecDestroy(h->graphs[v].exec);
                for (auto se : h->graphs[v].shard_exec) if (se) cudaGraphExecDestroy(se);
                h->graphs.erase(h->graphs.begin() + (long)v);
            }
            const uint64_t epoch0 = h->d.epoch;
            h->d.epoch_add = h->d_epoch_add;
            h->capturing = true;
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamBeginCapture(run, cudaStreamCaptureModeThreadLocal);
            if (e == cudaSuccess) {
                rc = playout_issue(h, mode, iters, ns, per, d_roots, d_leaves, pol, d, p1, p2, ml, run, &launches);
                e = cudaStreamEndCapture(run, &graph);
            }
            h->capturing = false;
            h->d.epoch_add = nullptr;
            h->d.epoch = epoch0;                              // the replay below accounts for it
            cudaGraphExec_t exec = nullptr;
            if (rc == AZ_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (rc != AZ_OK || e != cudaSuccess || !exec) {       // no graph: issue the loop launch by launch from now on
                cudaGetLastError();
                h->use_graphs = 0;
                h->bounds.clear();                               // (the growth reserved above is re-counted by the issued loop; harmless)
                rc = playout_issue(h, mode, iters, ns, per, d_roots, d_leaves, pol, d, p1, p2, ml, run, &launches); if (rc) return rc;
                if (!main) { CU(h, cudaEventRecord(h->ev, run)); CU(h, cudaStreamWaitEvent(main, h->ev, 0)); }
                h->user_stream = main; h->user_pending = true;
                if (launches_out) *launches_out = launches;
                return AZ_OK;
            }
            h->graphs.push_back({key, exec, epoch0, launches, 0, {}});
            ge = &h->graphs.back();
        }
        const unsigned long long delta = (unsigned long long)(h->d.epoch - ge->epoch0);
        CU(h, cudaMemcpyAsync(h->d_epoch_add, &delta, sizeof(delta), cudaMemcpyHostToDevice, run));
        CU(h, cudaGraphLaunch(ge->exec, run));
        if (!main) { CU(h, cudaEventRecord(h->ev, run)); CU(h, cudaStreamWaitEvent(main, h->ev, 0)); }
        ge->last_use = ++h->graph_clock; h->graph_replays++;
        h->tree_gen++;
        launches = ge->launches;
        // host-side state the issued loop would have left behind
        h->d.epoch += (uint64_t)iters.size();
        if (K > 0) h->prepared_K = iters.back() > 0 ? iters.back() : h->prepared_K;
        h->last_select_ro = select_is_ro(h, iters.back() > 0, std::max(iters.back(), 1));
        for (int j = 0; j < ns; ++j) {                       // what the issued loop's last select of every shard would have noted
            h->d.env_lo = j * per; h->d.env_cnt = std::min(per, h->n - j * per);
            note_select(h, iters.back() > 0, iters.back());
        }
        h->d.env_lo = 0; h->d.env_cnt = h->n;
        h->launches += (uint64_t)launches;
    }
    h->user_stream = main; h->user_pending = true;
    if (launches_out) *launches_out = launches;
    return AZ_OK;
}

// ---- the same loop from HOST arrays, pipelined shard by shard --------------------------------------------------------------
// What a host caller pays around the device loop - staging the boards, the host-to-device copy, the pack kernel, enqueueing the
// loop, the visit-count kernel, the device-to-host copy - is done per tree shard on the shard's own stream: the GPU starts on
// shard 0 while the host still stages shard 1, and the counts of a finished shard travel while the others search.  One CUDA
// graph per shard (the single graph of az_mcts_playout_synthetic_dev cannot start before everything is staged).  The counts
// land as int64 in a pinned pool block that the next az_mcts_get_counts64_pinned hands to the caller (no second pass).
static int issue_shard_loop(az_mcts *h, int mode, const std::vector<int> &iters, int kmax, int lo, int cnt, cudaStream_t s, int *launches) {
    float *pol = (float *)(h->io_in), *d, *p1, *p2, *ml;
    {   const InLayout L = in_layout((size_t)h->io_rows, h->A);
        uint8_t *q = h->io_in;
        pol = (float *)(q + L.policy); d = (float *)(q + L.d); p1 = (float *)(q + L.p1); p2 = (float *)(q + L.p2); ml = (float *)(q + L.ml); }
    const size_t r0 = (size_t)lo * kmax;
    for (size_t it = 0; it < iters.size(); ++it) {
        const int k = iters[it], kk = std::max(k, 1);
        int rc = do_search(h, k, h->io_roots, h->io_leaves, s, lo, cnt, true, (int64_t)r0); if (rc) return rc;
        rc = az_eval_synthetic_dev(h->game, mode, cnt * kk, h->io_leaves + r0, pol + r0 * h->A, d + r0, p1 + r0, p2 + r0, ml + r0, s);
        if (rc) AZ_FAIL(h, rc, "synthetic evaluator launch failed");
        rc = do_backprop(h, k, pol, d, p1, p2, ml, nullptr, nullptr, s, lo, cnt, (int64_t)r0); if (rc) return rc;
        *launches += 3;
    }
    return AZ_OK;
}

static void drop_cached_counts(az_mcts *h) {
    if (h->counts_block >= 0) { az_pinned_release(h->counts_block); h->counts_block = -1; h->counts_ptr = nullptr; }
}

int az_mcts_playout_synthetic_host(az_mcts *h, int mode, int n_playout, int K, int shards, const int8_t *boards, const int32_t *turns,
                                   int want_counts, int *launches_out) {
    if (n_playout < 0 || K < 0) AZ_FAIL(h, AZ_ERR_INVALID, "playout: n_playout and K must be >= 0");
    int rc = enter_host(h); if (rc) return rc;
    drop_cached_counts(h);
    shards = std::min(std::max(shards, 1), 16);
    const int per = (((h->n + shards - 1) / shards) + 31) / 32 * 32;
    const int ns = std::max(1, (h->n + per - 1) / per);
    std::vector<int> iters;
    if (K <= 1) iters.assign((size_t)n_playout, 0);
    else if (n_playout > 0) {
        iters.push_back(0);
        for (int rem = n_playout - 1; rem > 0; rem -= std::min(K, rem)) iters.push_back(std::min(K, rem));
    }
    int kmax = 1;
    for (int k : iters) kmax = std::max(kmax, k);
    rc = ensure_io(h, h->n * kmax); if (rc) return rc;
    rc = check_cfg(h, K); if (rc) return rc;
    const size_t cnt_all = (size_t)h->n * h->A;
    // staging (pinned) and its device twin: shard j = [boards of its trees | turns of its trees] in one piece, so one copy moves it
    const size_t stage_bytes = (size_t)h->n * (h->S + 4) + 32 * 17;
    auto off_b = [&](int j, int lo) { return (size_t)lo * (h->S + 4) + 32 * (size_t)j; };
    auto off_t = [&](int j, int lo, int cnt) { return off_b(j, lo) + (((size_t)cnt * h->S + 15) & ~(size_t)15); };
    if (!h->h_stage) {
        CU(h, cudaMallocHost((void **)&h->h_stage, stage_bytes));
        CU(h, cudaMalloc((void **)&h->io_stage, stage_bytes));
        CU(h, cudaMallocHost((void **)&h->h_err16, sizeof(int) * 16));
        CU(h, cudaMalloc((void **)&h->io_counts64, sizeof(int64_t) * cnt_all));
    }
    for (int j = 0; j < ns; ++j) {
        if (!h->side[j]) { CU(h, cudaStreamCreateWithFlags(&h->side[j], cudaStreamNonBlocking)); CU(h, cudaEventCreateWithFlags(&h->side_ev[j], cudaEventDisableTiming)); }
    }
    h->roots_valid = false;                 // io_roots no longer holds what host_search packed last
    int64_t *cdst = nullptr; int block = -1;
    if (want_counts) {
        void *bp = nullptr;
        block = pinned_acquire(sizeof(int64_t) * cnt_all, &bp);
        if (block < 0) AZ_FAIL(h, AZ_ERR_NOMEM, "cannot allocate %zu bytes of pinned host memory for the visit counts", sizeof(int64_t) * cnt_all);
        cdst = (int64_t *)bp;
    }
    auto fail = [&](int code) { if (block >= 0) az_pinned_release(block); return code; };
    int launches = 0;
    cudaStream_t s0 = h->stream;
    const InLayout IL = in_layout((size_t)h->io_rows, h->A);
    float *pol = (float *)(h->io_in + IL.policy), *dv = (float *)(h->io_in + IL.d), *p1 = (float *)(h->io_in + IL.p1),
          *p2 = (float *)(h->io_in + IL.p2), *ml = (float *)(h->io_in + IL.ml);
    const bool pipelined = ns > 1 && h->use_graphs && !h->time_select && !h->stats_on && iters.size() >= 4 && !h->lazy_live;
    if (!pipelined) {
        // small batches / measurement modes: everything on the internal stream, the loop through the whole-batch entry point
        const size_t ot = off_t(0, 0, h->n), total = ot + sizeof(int32_t) * (size_t)h->n;
        memcpy(h->h_stage, boards, (size_t)h->n * h->S); memcpy(h->h_stage + ot, turns, sizeof(int32_t) * (size_t)h->n);
        CU(h, cudaMemcpyAsync(h->io_stage, h->h_stage, total, cudaMemcpyHostToDevice, s0));
        rc = az_pack_roots_dev(h->game, h->n, (const int8_t *)h->io_stage, (const int32_t *)(h->io_stage + ot), h->io_roots, s0);
        if (rc) { h->err = "pack_roots launch failed"; return fail(rc); }
        rc = az_mcts_playout_synthetic_dev(h, mode, n_playout, K, ns, h->io_roots, h->io_leaves, pol, dv, p1, p2, ml, s0, &launches); if (rc) return fail(rc);
        launches += 1;
        if (want_counts) {
            if (h->game == GAME_C4) k_counts64_range<C4><<<grid_threads((size_t)h->n), 128, 0, s0>>>(h->d, 0, h->n, h->io_counts64);
            else k_counts64_range<Oth><<<grid_threads((size_t)h->n), 128, 0, s0>>>(h->d, 0, h->n, h->io_counts64);
            CU(h, cudaMemcpyAsync(cdst, h->io_counts64, sizeof(int64_t) * cnt_all, cudaMemcpyDeviceToHost, s0));
            launches += 1;
        }
        h->launches += (uint64_t)(want_counts ? 2 : 1);
        rc = check_device_error(h); if (rc) return fail(rc);     // (synchronises the internal stream)
    } else {
        if (K > 0) { rc = ensure_kcap(h, K); if (rc) return fail(rc); }
        for (int k : iters) bp_prepare(h, k > 0, k);
        if (!h->d_epoch_add) { CU(h, cudaMalloc((void **)&h->d_epoch_add, sizeof(unsigned long long))); }
        h->d.env_lo = 0; h->d.env_cnt = h->n;
        rc = ensure_arena(h, n_playout, s0); if (rc) return fail(rc);          // the whole loop's growth, before any capture
        az_mcts::GraphKey key;
        memset(&key, 0, sizeof(key));
        key.mode = mode; key.n_playout = n_playout; key.K = K; key.ns = ns; key.W = h->W; key.variant = h->variant; key.wave_max = h->wave_max; key.hints = h->d.hints;
        key.kcap = h->kcap; key.split = 1; key.cap = h->cap; key.cfg = h->cfg; key.seed = h->d.seed; key.env_base = h->d.env_base;
        const void *pp[10] = {h->io_roots, h->io_leaves, pol, dv, p1, p2, ml, h->d.pool, h->d.leaf_vl, h->h_stage};
        memcpy(key.ptrs, pp, sizeof(pp));
        az_mcts::GraphEntry *ge = nullptr;
        for (auto &g : h->graphs) if (memcmp(&g.key, &key, sizeof(key)) == 0) { ge = &g; break; }
        if (!ge) {
            if (h->graphs.size() >= 6) {
                size_t v = 0;
                for (size_t i = 1; i < h->graphs.size(); ++i) if (h->graphs[i].last_use < h->graphs[v].last_use) v = i;
                if (h->graphs[v].exec) cudaGraphExecDestroy(h->graphs[v].exec);
                for (auto se : h->graphs[v].shard_exec) if (se) cudaGraphExecDestroy(se);
                h->graphs.erase(h->graphs.begin() + (long)v);
            }
            const uint64_t epoch0 = h->d.epoch;
            std::vector<cudaGraphExec_t> execs;
            int per_shard_launches = 0;
            bool ok = true;
            h->d.epoch_add = h->d_epoch_add;
            for (int j = 0; j < ns && ok; ++j) {
                const int lo = j * per, cnt = std::min(per, h->n - lo);
                h->d.epoch = epoch0;                                  // every shard walks the same epochs
                h->capturing = true;
                cudaGraph_t graph = nullptr;
                int l = 0;
                cudaError_t e = cudaStreamBeginCapture(h->side[j], cudaStreamCaptureModeThreadLocal);
                if (e == cudaSuccess) {
                    // the shard's staging -> device copy and its pack kernel are nodes of the graph too (fixed addresses): one call starts a shard
                    const size_t ob = off_b(j, lo), ot = off_t(j, lo, cnt);
                    e = cudaMemcpyAsync(h->io_stage + ob, h->h_stage + ob, ot - ob + sizeof(int32_t) * (size_t)cnt, cudaMemcpyHostToDevice, h->side[j]);
                    if (e == cudaSuccess) {
                        rc = az_pack_roots_dev(h->game, cnt, (const int8_t *)(h->io_stage + ob), (const int32_t *)(h->io_stage + ot), h->io_roots + lo, h->side[j]);
                        if (rc) h->err = "pack_roots launch failed";
                    }
                    if (e == cudaSuccess && rc == AZ_OK) rc = issue_shard_loop(h, mode, iters, kmax, lo, cnt, h->side[j], &l);
                    l += 1;
                    const cudaError_t e2 = cudaStreamEndCapture(h->side[j], &graph);
                    if (e == cudaSuccess) e = e2;
                }
                h->capturing = false;
                cudaGraphExec_t exec = nullptr;
                if (rc == AZ_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
                if (graph) cudaGraphDestroy(graph);
                if (rc != AZ_OK || e != cudaSuccess || !exec) { ok = false; cudaGetLastError(); break; }
                execs.push_back(exec);
                per_shard_launches += l;
            }
            h->d.epoch_add = nullptr;
            h->d.epoch = epoch0;
            if (!ok) {
                for (auto e : execs) cudaGraphExecDestroy(e);
                h->bounds.clear(); h->bound_stale = true;
                AZ_FAIL(h, fail(AZ_ERR_CUDA), "playout_synthetic_host: the shard loops could not be captured into CUDA graphs");
            }
            h->graphs.push_back({key, nullptr, epoch0, per_shard_launches, 0, execs});
            ge = &h->graphs.back();
        }
        const unsigned long long delta = (unsigned long long)(h->d.epoch - ge->epoch0);
        CU(h, cudaMemcpyAsync(h->d_epoch_add, &delta, sizeof(delta), cudaMemcpyHostToDevice, s0));
        CU(h, cudaEventRecord(h->ev, s0));
        for (int j = 0; j < ns; ++j) h->h_err16[j] = 0;
        static const bool TRACE = getenv("AZB200_TRACE") != nullptr;
        double tr_t[64]; int tr_n = 0;
        auto now = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        if (TRACE) tr_t[tr_n++] = now();
        // (the staging is free: every call returns only after all its shards - and with them their copies - have completed)
        StagePool &sp = stage_pool_of(h);
        sp.begin();
        for (int j = 0; j < ns; ++j) {
            const int lo = j * per, cnt = std::min(per, h->n - lo);
            sp.chunks[j] = {{h->h_stage + off_b(j, lo), h->h_stage + off_t(j, lo, cnt)}, {boards + (size_t)lo * h->S, turns + lo},
                            {(size_t)cnt * h->S, sizeof(int32_t) * (size_t)cnt}};
        }
        sp.submit(ns);
        for (int j = 0; j < ns; ++j) {                                     // fronts: a shard starts as soon as ITS boards are staged
            cudaStream_t sj = h->side[j];
            cudaError_t e = cudaStreamWaitEvent(sj, h->ev, 0);             // after the re-root / whatever the internal stream holds
            sp.wait_chunk(j);
            if (e != cudaSuccess) { for (int q = j + 1; q < ns; ++q) sp.wait_chunk(q); fail(0); AZ_FAIL(h, AZ_ERR_CUDA, "cudaStreamWaitEvent failed: %s", cudaGetErrorString(e)); }
            e = cudaGraphLaunch(ge->shard_exec[(size_t)j], sj);
            if (e != cudaSuccess) { for (int q = j + 1; q < ns; ++q) sp.wait_chunk(q); fail(0); AZ_FAIL(h, AZ_ERR_CUDA, "cudaGraphLaunch failed: %s", cudaGetErrorString(e)); }
            if (TRACE) tr_t[tr_n++] = now();
        }
        for (int j = 0; j < ns; ++j) {                                     // tails: queued while the GPU is already searching
            const int lo = j * per, cnt = std::min(per, h->n - lo);
            cudaStream_t sj = h->side[j];
            if (want_counts) {
                if (h->game == GAME_C4) k_counts64_range<C4><<<grid_threads((size_t)cnt), 128, 0, sj>>>(h->d, lo, cnt, h->io_counts64);
                else k_counts64_range<Oth><<<grid_threads((size_t)cnt), 128, 0, sj>>>(h->d, lo, cnt, h->io_counts64);
                CU(h, cudaMemcpyAsync(cdst + (size_t)lo * h->A, h->io_counts64 + (size_t)lo * h->A, sizeof(int64_t) * (size_t)cnt * h->A, cudaMemcpyDeviceToHost, sj));
            }
            CU(h, cudaMemcpyAsync(h->h_err16 + j, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, sj));
            CU(h, cudaEventRecord(h->side_ev[j], sj));
        }
        if (TRACE) tr_t[tr_n++] = now();
        ge->last_use = ++h->graph_clock; h->graph_replays++;
        launches = ge->launches + (want_counts ? ns : 0);
        // host-side state the issued loop would have left behind
        h->d.epoch += (uint64_t)iters.size();
        h->tree_gen++;
        if (K > 0) h->prepared_K = iters.back() > 0 ? iters.back() : h->prepared_K;
        h->last_select_ro = select_is_ro(h, iters.back() > 0, std::max(iters.back(), 1));
        for (int j = 0; j < ns; ++j) {
            h->d.env_lo = j * per; h->d.env_cnt = std::min(per, h->n - j * per);
            note_select(h, iters.back() > 0, iters.back());
        }
        h->d.env_lo = 0; h->d.env_cnt = h->n;
        h->launches += (uint64_t)launches;
        bool dev_err = false;
        for (int j = 0; j < ns; ++j) {
            cudaError_t se = cudaEventSynchronize(h->side_ev[j]);
            if (se != cudaSuccess) { fail(0); AZ_FAIL(h, AZ_ERR_CUDA, "playout_synthetic_host: shard %d failed: %s", j, cudaGetErrorString(se)); }
            dev_err = dev_err || h->h_err16[j] != 0;
            if (TRACE) tr_t[tr_n++] = now();
        }
        if (TRACE) { fprintf(stderr, "[trace]"); for (int i = 1; i < tr_n; ++i) fprintf(stderr, " %.0f", tr_t[i] - tr_t[0]); fprintf(stderr, "\n"); }
        if (dev_err) { fail(0); AZ_FAIL(h, AZ_ERR_NOMEM, "device tree arena overflow (internal sizing error)"); }
    }
    if (want_counts) { h->counts_block = block; h->counts_ptr = cdst; h->counts_gen = h->tree_gen; }
    if (launches_out) *launches_out = launches;
    return AZ_OK;
}

int az_mcts_search_eval_dev(az_mcts *h, int evaluator, const az_root *d_roots, int n_playout, void *stream) {
    if (evaluator != AZ_EVAL_UNIFORM && evaluator != AZ_EVAL_ROLLOUT) AZ_FAIL(h, AZ_ERR_INVALID, "unknown evaluator kind %d", evaluator);
    int rc = check_cfg(h, 1); if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    rc = enter_dev(h, s); if (rc) return rc;
    rc = ensure_io(h, h->n); if (rc) return rc;
    h->d.epoch++;                          // one epoch per search() call, like orc_search
    h->d.env_lo = 0; h->d.env_cnt = h->n;
    const InLayout L = in_layout((size_t)h->n, h->A);
    uint8_t *q = h->io_in;
    float *pol = (float *)(q + L.policy), *dv = (float *)(q + L.d), *p1 = (float *)(q + L.p1), *p2 = (float *)(q + L.p2), *ml = (float *)(q + L.ml);
    az_search_config saved = h->cfg;
    h->cfg.use_symmetry = 0;               // search() never symmetrises leaves (BatchedMCTS.h:357)
    h->d.stats = h->stats_on ? h->d_stats : nullptr;
    for (int p = 0; p < n_playout; ++p) {
        rc = ensure_arena(h, 1, s); if (rc) { h->cfg = saved; return rc; }
        rc = launch_select(h, false, 1, d_roots, h->io_leaves, s); if (rc) { h->cfg = saved; return rc; }
        if (h->game == GAME_C4) k_eval_builtin<C4><<<grid_threads((size_t)h->n), 128, 0, s>>>(h->d, evaluator, p, pol, dv, p1, p2, ml);
        else k_eval_builtin<Oth><<<grid_threads((size_t)h->n), 128, 0, s>>>(h->d, evaluator, p, pol, dv, p1, p2, ml);
        rc = launch_backprop(h, false, 1, 0, 0, pol, dv, p1, p2, ml, nullptr, nullptr, s); if (rc) { h->cfg = saved; return rc; }
        h->launches += 1;
    }
    h->cfg = saved;
    CU(h, cudaGetLastError());
    return AZ_OK;
}
int az_mcts_search(az_mcts *h, int evaluator, const int8_t *boards, const int32_t *turns, int n_playout) {
    { int rc0 = enter_host(h); if (rc0) return rc0; }
    CU(h, cudaMemcpyAsync(h->io_boards_in, boards, (size_t)h->n * h->S, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->io_turns_in, turns, sizeof(int32_t) * (size_t)h->n, cudaMemcpyHostToDevice, h->stream));
    h->roots_valid = false;                               // io_roots no longer holds what host_search packed last
    int rc = az_pack_roots_dev(h->game, h->n, h->io_boards_in, h->io_turns_in, h->io_roots, h->stream); if (rc) AZ_FAIL(h, rc, "pack_roots launch failed");
    rc = az_mcts_search_eval_dev(h, evaluator, h->io_roots, n_playout, h->stream); if (rc) return rc;
    return check_device_error(h);
}

int az_mcts_get_counts_dev(az_mcts *h, int32_t *d_out, void *stream) {
    { int rc = enter_dev(h, (cudaStream_t)stream); if (rc) return rc; }
    if (h->game == GAME_C4) k_counts<C4><<<grid_threads((size_t)h->n), 128, 0, (cudaStream_t)stream>>>(h->d, d_out);
    else k_counts<Oth><<<grid_threads((size_t)h->n), 128, 0, (cudaStream_t)stream>>>(h->d, d_out);
    h->launches++;
    CU(h, cudaGetLastError());
    return AZ_OK;
}
int az_mcts_get_counts(az_mcts *h, int32_t *out) {
    int rc = enter_host(h); if (rc) return rc;
    rc = az_mcts_get_counts_dev(h, h->io_counts, h->stream); if (rc) return rc;
    CU(h, cudaMemcpyAsync(out, h->io_counts, sizeof(int32_t) * (size_t)h->n * h->A, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return AZ_OK;
}
// int32 -> int64 on the host.  The destination is a fresh 3.7 MB numpy array at 65 536 trees: written with streaming stores
// (no read-for-ownership of lines that are overwritten entirely) when AVX2 is there.
#if defined(__x86_64__)
__attribute__((target("avx2"))) static void widen_avx2(const int32_t *src, int64_t *dst, size_t cnt) {
    size_t i = 0;
    while (i < cnt && ((uintptr_t)(dst + i) & 31)) { dst[i] = (int64_t)src[i]; ++i; }
    for (; i + 8 <= cnt; i += 8) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), _mm256_cvtepi32_epi64(_mm256_castsi256_si128(v)));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 4), _mm256_cvtepi32_epi64(_mm256_extracti128_si256(v, 1)));
    }
    _mm_sfence();
    for (; i < cnt; ++i) dst[i] = (int64_t)src[i];
}
#endif
static void widen_counts(const int32_t *src, int64_t *dst, size_t cnt) {
#if defined(__x86_64__)
    if (__builtin_cpu_supports("avx2")) { widen_avx2(src, dst, cnt); return; }
#endif
    for (size_t i = 0; i < cnt; ++i) dst[i] = (int64_t)src[i];
}
// Visit counts widened to int64 (what callers of the reference build from get_all_counts(): np.array(list of int)): one D2H
// copy into pinned memory, widened on the host.
int az_mcts_get_counts64(az_mcts *h, int64_t *out) {
    int rc = enter_host(h); if (rc) return rc;
    const size_t cnt = (size_t)h->n * h->A;
    if (!h->h_counts) CU(h, cudaMallocHost((void **)&h->h_counts, sizeof(int32_t) * cnt));
    rc = az_mcts_get_counts_dev(h, h->io_counts, h->stream); if (rc) return rc;
    CU(h, cudaMemcpyAsync(h->h_counts, h->io_counts, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    widen_counts(h->h_counts, out, cnt);
    return AZ_OK;
}
// The int64 counts in a pinned pool block that the CALLER owns until az_pinned_release(*block_out): either the block the last
// az_mcts_playout_synthetic_host already filled (nothing touched the trees since), or a fresh one (counts widened on the device, one copy).
int az_mcts_get_counts64_pinned(az_mcts *h, int64_t **out, int *block_out) {
    int rc = enter_host(h); if (rc) return rc;
    if (h->counts_block >= 0 && h->counts_gen == h->tree_gen) {
        *out = h->counts_ptr; *block_out = h->counts_block;
        h->counts_block = -1; h->counts_ptr = nullptr;
        return AZ_OK;
    }
    drop_cached_counts(h);
    const size_t cnt = (size_t)h->n * h->A;
    if (!h->io_counts64) CU(h, cudaMalloc((void **)&h->io_counts64, sizeof(int64_t) * cnt));
    void *bp = nullptr;
    const int block = pinned_acquire(sizeof(int64_t) * cnt, &bp);
    if (block < 0) AZ_FAIL(h, AZ_ERR_NOMEM, "cannot allocate %zu bytes of pinned host memory for the visit counts", sizeof(int64_t) * cnt);
    if (h->game == GAME_C4) k_counts64_range<C4><<<grid_threads((size_t)h->n), 128, 0, h->stream>>>(h->d, 0, h->n, h->io_counts64);
    else k_counts64_range<Oth><<<grid_threads((size_t)h->n), 128, 0, h->stream>>>(h->d, 0, h->n, h->io_counts64);
    h->launches++;
    cudaError_t e = cudaMemcpyAsync(bp, h->io_counts64, sizeof(int64_t) * cnt, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { az_pinned_release(block); AZ_FAIL(h, AZ_ERR_CUDA, "get_counts64_pinned: %s", cudaGetErrorString(e)); }
    *out = (int64_t *)bp; *block_out = block;
    return AZ_OK;
}
int az_mcts_get_root_stats_dev(az_mcts *h, float *d_out, void *stream) {
    { int rc = enter_dev(h, (cudaStream_t)stream); if (rc) return rc; }
    if (h->game == GAME_C4) k_root_stats<C4><<<grid_threads((size_t)h->n), 128, 0, (cudaStream_t)stream>>>(h->d, d_out);
    else k_root_stats<Oth><<<grid_threads((size_t)h->n), 128, 0, (cudaStream_t)stream>>>(h->d, d_out);
    h->launches++;
    CU(h, cudaGetLastError());
    return AZ_OK;
}
int az_mcts_get_root_stats(az_mcts *h, float *out) {
    int rc = enter_host(h); if (rc) return rc;
    rc = az_mcts_get_root_stats_dev(h, h->io_stats, h->stream); if (rc) return rc;
    CU(h, cudaMemcpyAsync(out, h->io_stats, sizeof(float) * (size_t)h->n * (6 + 8 * h->A), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return AZ_OK;
}
int az_mcts_enable_stats(az_mcts *h, int on) {
    CU(h, cudaSetDevice(h->device));
    h->stats_on = on != 0;
    CU(h, cudaDeviceSynchronize());
    CU(h, cudaMemset(h->d_stats, 0, 8 * sizeof(unsigned long long)));
    return AZ_OK;
}
int az_mcts_time_select(az_mcts *h, int on) { h->time_select = on != 0; h->sel_used = 0; h->sel_rows = 0; h->bp_used = 0; h->bp_rows = 0; return AZ_OK; }
int az_mcts_get_backprop_time(az_mcts *h, float *ms_out, int *launches_out, uint64_t *rows_out) {
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaDeviceSynchronize());
    float tot = 0.0f;
    for (size_t i = 0; i < h->bp_used; ++i) { float ms = 0.0f; CU(h, cudaEventElapsedTime(&ms, h->bp_ev[i].first, h->bp_ev[i].second)); tot += ms; }
    if (ms_out) *ms_out = tot;
    if (launches_out) *launches_out = (int)h->bp_used;
    if (rows_out) *rows_out = h->bp_rows;
    h->bp_used = 0; h->bp_rows = 0;
    return AZ_OK;
}
int az_mcts_get_select_time(az_mcts *h, float *ms_out, int *launches_out, uint64_t *rows_out) {
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaDeviceSynchronize());
    float tot = 0.0f;
    for (size_t i = 0; i < h->sel_used; ++i) { float ms = 0.0f; CU(h, cudaEventElapsedTime(&ms, h->sel_ev[i].first, h->sel_ev[i].second)); tot += ms; }
    if (ms_out) *ms_out = tot;
    if (launches_out) *launches_out = (int)h->sel_used;
    if (rows_out) *rows_out = h->sel_rows;
    h->sel_used = 0; h->sel_rows = 0;
    return AZ_OK;
}
int az_mcts_get_warp_times(az_mcts *h, uint64_t *out, int max_warps) {
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaDeviceSynchronize());
    const int nw = std::min(std::min(max_warps, AZ_DBG_WARPS), (h->n + 31) / 32);
    CU(h, cudaMemcpy(out, h->d_stats + 8, sizeof(uint64_t) * 2 * (size_t)nw, cudaMemcpyDeviceToHost));
    return nw;
}
int az_mcts_get_stats(az_mcts *h, uint64_t *out8) {
    CU(h, cudaSetDevice(h->device));
    unsigned long long v[8];
    CU(h, cudaDeviceSynchronize());
    CU(h, cudaMemcpy(v, h->d_stats, sizeof(v), cudaMemcpyDeviceToHost));
    CU(h, cudaMemsetAsync(h->d_scratch_u32, 0, sizeof(unsigned int), h->stream));
    k_max_bump<<<grid_threads((size_t)h->n, 256), 256, 0, h->stream>>>(h->d, h->d_scratch_u32);
    unsigned int mx = 0;
    CU(h, cudaMemcpyAsync(&mx, h->d_scratch_u32, sizeof(mx), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 5; ++i) out8[i] = v[i];
    out8[5] = mx; out8[6] = h->cap; out8[7] = h->launches;
    return AZ_OK;
}

}  // extern "C"
