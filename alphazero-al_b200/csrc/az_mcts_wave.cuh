// Staggered-descent select for SMALL Connect4 batches (included by az_mcts.cu after az_mcts_fast.cuh).
//
// The K virtual-loss descents of a tree are sequentially dependent (descent k sees the virtual loss of descents < k), so the
// thread-per-tree kernels walk K x depth tree levels one after the other: at 8192 trees or fewer that chain IS the kernel
// time (a few hundred warps cannot hide a DRAM round trip per level).  But descent k only needs to know which child the
// earlier descents took AT THE LEVEL IT IS SCORING - not their leaves.  Here every descent gets its own lane and starts one
// level step after its predecessor: in step s lane k scores level s - k, and the choices of lanes < k at that level were made
// in earlier steps (they are read from the paths kept in shared memory, exactly like the read-only select of k_select_f).
// K + depth - 1 dependent level steps instead of K x depth, same arithmetic, same bits.
//
// Read-only on the tree (launch-local virtual loss, first-visit flags applied by back-prop): the matching back-prop is
// k_backprop_f<..., RO = true>.  More warp instructions in total than the thread-per-tree kernel (lanes idle while they wait
// for their start step): only used when the whole batch is small (launch_select).
#pragma once

namespace az {

constexpr int CTA_W = 128;

// KL lanes per tree (4 or 8), lane k of a group runs descent k (k < K <= KL)
template <class G, bool AUX, int KL, bool LAZY = false>
__global__ void __launch_bounds__(CTA_W) k_select_w(Dev d, az_search_config cfg, int K, const az_root *__restrict__ roots,
                                                 az_leaf *__restrict__ leaves) {
    static_assert(G::GAME == GAME_C4, "staggered select is specialised for Connect4 (<= 7 edges)");
    constexpr int NE = G::MAX_EDGES;       // 7
    __shared__ uint32_t path_s[CTA_W][PATH8 + 1];          // first 8 path entries of every descent; odd stride
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int k = lane % KL;                               // my descent
    const int gbase = lane - k;                            // first lane of my tree's group
    const int tree = (blockIdx.x * CTA_W + threadIdx.x) / KL;
    const bool valid = tree < d.env_cnt && k < K;
    const int env = d.env_lo + (tree < d.env_cnt ? tree : d.env_cnt - 1);
    const Slot *arena = d.pool + (size_t)env * d.cap;
    const TreeRec *tr = d.trees + env;
    const int vl = cfg.vl_count;
    const float ne_eps = cfg.noise_epsilon;
    uint32_t *mypath = &path_s[threadIdx.x][0];
    const uint32_t *gpaths = &path_s[threadIdx.x - k][0];  // paths of my group: descent q at gpaths + q * (PATH8 + 1)
#pragma unroll
    for (int j = 0; j < PATH8; ++j) mypath[j] = 0;

    pdl_wait();                                            // the tree is the previous kernel's output
    // import_board (Connect4.h:100-129): the last mover is inferred from piece-count parity
    uint64_t b0, b1; int turn, last;
    { const az_root r = ld32(roots + env); b0 = r.bb0; b1 = r.bb1; turn = r.turn;
      const int np = popc64(r.bb0 | r.bb1); last = np > 0 ? ((np & 1) ? 0 : 1) : -1; }
    const Slot root = ld_slot(&tr->root);
    uint32_t root_meta = root.meta;
    float nz[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) nz[e] = ne_eps > 0.0f ? d.noise[(size_t)env * d.noise_stride + e] : 0.0f;
    const uint64_t keep = l2_keep_policy();

    int cur_n = root.n; uint32_t cur_meta = root.meta;
    float cur_Q = mean_q(root.n, root.wp1, root.wp2, (root.meta & F_TURN_P1) != 0);
    float cur_M = AUX ? mean_m(root.n, root.msum) : 0.0f;
    bool is_root = true;
    uint32_t plen = 0, last_slot = 0;
    int winner = 0; bool full = false;
    unsigned st_edges = 0;
    uint32_t *gpath = d.path_vl + ((size_t)env * d.kcap + k) * G::MAX_DEPTH;       // entries beyond the first 8
    // w = (block offset << 6) | num_edges of the node I score next, 0 = my descent has ended
    // (bit 5 of w: the node is a lazy block - only its 32-byte header exists, see F_LAZY; the root never is)
    constexpr uint32_t W_LAZY = 32u;
    uint32_t w = (valid && root.child != NONE && !(root.meta & F_TERM) && (root.child & 63u) != 0) ? root.child : 0u;
    __syncwarp();

    for (int step = 0; __any_sync(FULL, w != 0u); ++step) {
        const bool act = w != 0u && step >= k;
        const uint32_t off = w >> 6;
        const int ne = (int)(w & 7u);
        const bool lazy = LAZY && (w & W_LAZY) != 0u;
        // ---- in-flight counts from the earlier descents of my tree: a path that holds a slot of this block at this depth
        //      passed through this node (parent + 1) and through that child (child + 1); 4 bits per child ----
        uint32_t packed = 0u, cntp = 0u;
#pragma unroll
        for (int q = 0; q < KL - 1; ++q) {
            const uint32_t pq = __shfl_sync(FULL, plen, gbase + q);      // descent q is at least one level ahead of me (or done)
            if (act && q < k && pq > plen) {
                const uint32_t t = plen < (uint32_t)PATH8 ? gpaths[q * (PATH8 + 1) + plen]
                                                          : __ldcg(d.path_vl + ((size_t)env * d.kcap + q) * G::MAX_DEPTH + plen);
                const uint32_t dd = t - off;
                if (dd < (uint32_t)ne) { packed += 1u << (4 * dd); ++cntp; }
            }
        }
        uint32_t nw = 0u;
        if (act) {
            st_edges += (unsigned)ne;
            Slot s[NE];
            Slot hdr;
            if (lazy) hdr = ld_slot256(arena + off);
#pragma unroll
            for (int c = 0; c < NE; ++c) {
                if (c < ne) s[c] = lazy ? lazy_edge(hdr, c) : ld_slot256(arena + off + c);
                else { s[c].prior = 0.0f; s[c].n = 0; s[c].meta = 0u; s[c].child = NONE; s[c].wd = s[c].wp1 = s[c].wp2 = s[c].msum = 0.0f; }
            }
            // ---- compute_fpu (MCTS.h:140-156): seen_policy summed in edge order ----
            const float parent_q = cur_Q;
            float seen_policy = 0.0f;
#pragma unroll
            for (int c = 0; c < NE; ++c) seen_policy += (c < ne && s[c].n > 0) ? s[c].prior : 0.0f;   // + 0.0f is exact
            const float fscale = (1.0f + parent_q) / 2.0f;
            const float eff_fpu = cfg.fpu_reduction * fscale;
            float fpu = parent_q - eff_fpu * sqrtf(seen_policy);
            fpu = (-1.0f < fpu) ? fpu : -1.0f;
            // ---- select_edge (MCTS.h:163-234); a node below the root already carries this descent's own virtual loss
            //      (MCTS.h:492), the root gets its own after the first selection (:471-475) ----
            const int pn_i = cur_n + (int)(cntp + (is_root ? 0u : 1u)) * vl;
            const float parent_n = (float)pn_i;
            const float parent_M = cur_M;
            float lg, sqrt_pn;
            if (pn_i >= 0 && pn_i < d.log_lut_n) { const float2 v = __ldg(d.ls_lut + pn_i); lg = v.x; sqrt_pn = v.y; }
            else { lg = logf((parent_n + cfg.c_base + 1.0f) / cfg.c_base); sqrt_pn = sqrtf(parent_n); }
            const float c_puct = cfg.c_init + lg;
            const bool mix_noise = is_root && ne_eps > 0.0f;
            float best_s = -INFINITY, best_Q = 0.0f, best_M = 0.0f;
            int best_e = -1;
            // the branch-free correctly rounded divisions of k_select_f (az_mcts_fast.cuh): this kernel is instruction bound
            // (profiles: 51 % issue slots busy at 10 of 32 lanes active), and an IEEE `/` costs ~15 instructions and a branch
            SafeAcc safe;
#pragma unroll
            for (int c = 0; c < NE; ++c) {
                float eff_prior = s[c].prior;
                if (mix_noise) eff_prior = (1.0f - ne_eps) * s[c].prior + ne_eps * nz[c];
                const bool has = s[c].n > 0;
                const float nf = (float)max(s[c].n, 1);
                const float rn = rcp_refined(nf);                         // == 1.0f / nf
                const float p1 = s[c].wp1 * rn, p2 = s[c].wp2 * rn;
                const float dq = p1 - p2;                                  // p2 - p1 == -(p1 - p2) exactly
                const float child_Q = (s[c].meta & F_TURN_P1) ? dq : -dq;
                float child_M = 0.0f, m_utility = 0.0f;
                if (AUX) {
                    child_M = div_by_rcp(s[c].msum, nf, rn);
                    const float m_diff = child_M - parent_M;               // Connect4.h:231-239
                    const float v = cfg.mlh_slope * m_diff, lo = -cfg.mlh_cap, hi = cfg.mlh_cap;
                    const float u = v < lo ? lo : (hi < v ? hi : v);
                    m_utility = has ? u * child_Q : 0.0f;
                    if (c < ne) safe.add(s[c].msum);
                }
                const float q_value = has ? -child_Q : fpu;
                const int visits = s[c].n + (int)((packed >> (4 * c)) & 15u) * vl;
                const float den = 1.0f + (float)visits;
                const float num = c_puct * eff_prior * sqrt_pn;
                if (c < ne) safe.add(num);
                const float u_score = div_by_rcp(num, den, rcp_refined(den));
                const float score = q_value + u_score + m_utility;
                if (c < ne && score > best_s) { best_s = score; best_e = c; best_Q = child_Q; best_M = child_M; }
            }
            if (!safe.ok()) {      // rare: a numerator outside the range the fast division covers - the plain IEEE operators
                best_s = -INFINITY; best_e = -1; best_Q = 0.0f; best_M = 0.0f;
#pragma unroll 1
                for (int c = 0; c < ne; ++c) {
                    const Slot sc = lazy ? lazy_edge(hdr, c) : ld_slot256(arena + off + c);
                    float eff_prior = sc.prior;
                    if (mix_noise) eff_prior = (1.0f - ne_eps) * sc.prior + ne_eps * d.noise[(size_t)env * d.noise_stride + c];
                    float q_value = fpu, m_utility = 0.0f, child_Q = 0.0f, child_M = 0.0f;
                    if (sc.n > 0) {
                        child_Q = mean_q(sc.n, sc.wp1, sc.wp2, (sc.meta & F_TURN_P1) != 0);
                        q_value = -child_Q;
                        if (AUX) { child_M = mean_m(sc.n, sc.msum); m_utility = aux_utility<G>(child_M, parent_M, child_Q, cfg); }
                    }
                    const int visits = sc.n + (int)((packed >> (4 * c)) & 15u) * vl;
                    const float u_score = c_puct * eff_prior * sqrt_pn / (1.0f + (float)visits);
                    const float score = q_value + u_score + m_utility;
                    if (score > best_s) { best_s = score; best_e = c; best_Q = child_Q; best_M = child_M; }
                }
            }
            if (best_e >= 0) {
                const Slot ch = lazy ? lazy_edge(hdr, best_e) : ld_slot256(arena + off + best_e);   // re-read (L1 hit) instead of a 7-way select of 8 registers
                {   // Connect4::step (Connect4.h:159-172, no legality check): drop a stone of the side to move
                    const int col7 = (int)((ch.meta >> 16) & 0xFFu) * 7;
                    const uint64_t occ = b0 | b1;
                    const uint64_t bit = 1ULL << (col7 + popc64((occ >> col7) & 0x3FULL));
                    const bool p1_moves = turn == 1;
                    b0 |= p1_moves ? bit : 0ULL; b1 |= p1_moves ? 0ULL : bit;
                    last = p1_moves ? 0 : 1; turn = -turn;
                }
                uint32_t nmeta = ch.meta;
                if (!(nmeta & F_ALLOC)) {      // lazy child allocation (MCTS.h:481-488): remember the child's side to move
                    nmeta |= F_ALLOC;
                    nmeta = turn == 1 ? (nmeta | F_TURN_P1) : (nmeta & ~F_TURN_P1);
                }
                winner = c4_winner_of(last == 0 ? b0 : b1) ? (last == 0 ? 1 : -1) : 0;       // last mover only (:182-203)
                full = popc64(b0 | b1) == 42;
                const bool term_now = winner != 0 || full;
                if (term_now) nmeta = (nmeta & ~(F_WIN_P1 | F_WIN_P2)) | F_TERM | (winner == 1 ? F_WIN_P1 : (winner == -1 ? F_WIN_P2 : 0u));
                last_slot = off + (uint32_t)best_e;
                if (plen < (uint32_t)PATH8) mypath[plen] = last_slot; else gpath[plen] = last_slot;
                ++plen;
                cur_n = ch.n; cur_meta = nmeta; cur_Q = best_Q; cur_M = best_M; is_root = false;
                if (ch.child != NONE && !(ch.meta & F_TERM) && (ch.child & 63u) != 0 && plen < (uint32_t)G::MAX_DEPTH && !term_now)
                    nw = LAZY ? ((ch.child & ~63u) | (ch.child & 7u) | ((ch.meta & F_LAZY) ? W_LAZY : 0u)) : ch.child;
            }
        }
        if (act) w = nw;
        __syncwarp();                                  // my path entry is visible to the later descents of my tree
    }

    if (valid) {
        bool leaf_term = (cur_meta & F_TERM) != 0;
        if (plen == 0) leaf_term = (root_meta & F_TERM) != 0;
        if (!leaf_term) {
            if (winner == 0 && !full) {
                winner = (last >= 0 && c4_winner_of(last == 0 ? b0 : b1)) ? (last == 0 ? 1 : -1) : 0;
                full = popc64(b0 | b1) == 42;
            }
            if (winner != 0 || full) {
                leaf_term = true;
                const uint32_t tf = F_TERM | (winner == 1 ? F_WIN_P1 : (winner == -1 ? F_WIN_P2 : 0u));
                if (plen == 0) { root_meta = (root_meta & ~(F_WIN_P1 | F_WIN_P2)) | tf; cur_meta = root_meta; }
                else cur_meta = (cur_meta & ~(F_WIN_P1 | F_WIN_P2)) | tf;
            }
        }
        int sym = 0;
        uint64_t e0 = b0, e1 = b1;
        if (!leaf_term && cfg.use_symmetry) {
            const uint64_t h = az_rand(d.seed, d.epoch + (d.epoch_add ? *d.epoch_add : 0ULL), STREAM_SYM, d.env_base + (uint64_t)env, (uint64_t)k);
            sym = (int)(h & 1);
            if (sym) { e0 = G::flip_bb(b0); e1 = G::flip_bb(b1); }
        }
        const uint8_t tflags = (uint8_t)(leaf_term ? (AZ_LEAF_TERMINAL | ((cur_meta & F_WIN_P1) ? AZ_LEAF_P1_WINS : 0u) |
                                                      ((cur_meta & F_WIN_P2) ? AZ_LEAF_P2_WINS : 0u)) : 0u);
        const uint32_t lflags = LF_VALID | (leaf_term ? LF_TERM : 0u) | ((tflags & AZ_LEAF_P1_WINS) ? LF_WIN_P1 : 0u) |
                                ((tflags & AZ_LEAF_P2_WINS) ? LF_WIN_P2 : 0u);
        LeafRec *dst = d.leaf_vl + (size_t)env * d.kcap + k;
        st_words256_keep(&dst->h, (uint32_t)b0, (uint32_t)(b0 >> 32), (uint32_t)b1, (uint32_t)(b1 >> 32), (uint32_t)turn,
                         (((uint32_t)last & 0xFFu) << 16) | (lflags << 24), plen, (uint32_t)sym, keep);     // passes = 0
        st_words256_keep(dst->path8, mypath[0], mypath[1], mypath[2], mypath[3], mypath[4], mypath[5], mypath[6], mypath[7], keep);
        st_words256_keep(leaves + (size_t)env * K + k, (uint32_t)e0, (uint32_t)(e0 >> 32), (uint32_t)e1, (uint32_t)(e1 >> 32),
                         ((uint32_t)turn & 0xFFu) | ((uint32_t)tflags << 8) | ((uint32_t)sym << 16), 0u, 0u, 0u, keep);
    }
    if (d.stats) {                                   // warp-uniform; one atomic per warp and counter
        const unsigned sims = __reduce_add_sync(FULL, valid ? 1u : 0u);
        const unsigned dep = __reduce_add_sync(FULL, valid ? (unsigned)plen : 0u);
        const unsigned edg = __reduce_add_sync(FULL, valid ? st_edges : 0u);
        if (lane == 0) { atomicAdd(d.stats + 0, (unsigned long long)sims); atomicAdd(d.stats + 1, (unsigned long long)dep); atomicAdd(d.stats + 2, (unsigned long long)edg); }
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// The same staggering for the lane-group kernels (Othello): a warp per tree, 32 / W descents of W cooperating lanes each,
// descent k one level behind descent k - 1.  Virtual loss stays in the tree as in k_select (written into the slot when a
// descent passes, removed by back-prop): at every step each tree level is touched by at most one descent, and descents
// reach a given level in k order, so every load sees exactly what the sequential loop would have seen.  The level body is
// k_select's.
template <class G, int W>
__global__ void __launch_bounds__(CTA, 7) k_select_ws(Dev d, az_search_config cfg, int K, const az_root *__restrict__ roots,
                                                   az_leaf *__restrict__ leaves) {
    constexpr int NCH = (G::MAX_EDGES + W - 1) / W;
    __shared__ float seen_s[CTA / W][G::MAX_EDGES + 2];              // per lane group: the visited children's priors in edge order
    const unsigned FULL = 0xFFFFFFFFu;
    const int tree = (blockIdx.x * CTA + threadIdx.x) >> 5;          // one warp per tree
    if (tree >= d.env_cnt) return;
    const int k = (threadIdx.x & 31) / W;                            // my group's descent
    const int lane = threadIdx.x & (W - 1);
    const unsigned gm = group_mask<W>();
    const int env = d.env_lo + tree;
    Slot *arena = d.pool + (size_t)env * d.cap;
    TreeRec *tr = d.trees + env;
    const float *noise = d.noise + (size_t)env * d.noise_stride;
    const int vl = cfg.vl_count;
    const bool use_aux = aux_enabled<G>(cfg);
    const bool mine = k < K;

    State st;
    { const az_root r = ld32(roots + env); st.bb[0] = r.bb0; st.bb[1] = r.bb1; G::finish_import(st, r.turn); }
    const Slot root = ld_slot(&tr->root);
    // Every descent selects at the root or none does (an unexpanded / terminal / edgeless root ends all of them there), so the
    // root's in-flight count seen by descent k is what the k earlier descents added (MCTS.h:471-475).
    Slot cur = root;
    cur.meta += (uint32_t)(k * vl);
    bool is_root = true, done = !mine;
    uint32_t plen = 0, last_slot = 0;
    uint32_t *path = d.path_vl + ((size_t)env * d.kcap + k) * G::MAX_DEPTH;
    int winner = 0; bool full = false;
    unsigned long long st_edges = 0;

    for (int step = 0; __any_sync(FULL, !done); ++step) {
        if (!done && step >= k) {
            const int ne = cur.child != NONE ? (int)(cur.child & 63u) : 0;
            if (cur.child == NONE || (cur.meta & F_TERM) || ne == 0 || plen >= (uint32_t)G::MAX_DEPTH) done = true;
            else {
                const uint32_t off = cur.child >> 6;
                // The edges are walked in chunks of W (lane l owns edges l, l + W, ...) with nothing but the running best kept in
                // registers: an array of all ceil(48 / W) chunks cost 96 registers per thread, 20 resident warps per SM, and 4096
                // trees (one warp each) then need 1.4 waves; at 64 registers they are all resident (config 4: 8.8 -> 8.0 ms per move
                // already with a forced cap and its spills).  Second reads of a slot hit L1.
                st_edges += (unsigned long long)ne;
                // ---- compute_fpu (MCTS.h:140-156): seen_policy summed sequentially in edge order over the visited children ----
                const int cur_infl = (int)(cur.meta & INFL_MASK);
                const float parent_q = mean_q(cur.n, cur.wp1, cur.wp2, (cur.meta & F_TURN_P1) != 0);
                float seen_policy = 0.0f;
                {   // the priors of the visited children (0 for the others: + 0.0f is exact) go to the group's row in shared memory, then every
                    // lane adds them in edge order: a load and a dependent add per edge instead of a ballot / bit-scan / shuffle loop
                    float *sp_ = seen_s[threadIdx.x / W];
#pragma unroll 1
                    for (int c0 = 0; c0 < ne; c0 += W) {
                        const int e = c0 + lane;
                        if (e < ne) { const uint2 pn = *reinterpret_cast<const uint2 *>(arena + off + e); sp_[e] = (int)pn.y > 0 ? __uint_as_float(pn.x) : 0.0f; }
                    }
                    __syncwarp(gm);
#pragma unroll 4
                    for (int e = 0; e < ne; ++e) seen_policy += sp_[e];
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
                if (best_e < 0) done = true;
                else {
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
                    if (term_now) done = true;
                }
            }
        }
        __syncwarp();          // this step's in-flight / flag updates are visible to the descents that reach the level next
    }

    const unsigned sel = __ballot_sync(FULL, mine && lane == 0 && plen > 0);       // descents that added virtual loss to the root
    uint32_t root_meta = root.meta + (uint32_t)(__popc(sel) * vl);
    if (mine) {
        // ---- leaf classification (MCTS.h:512-544) ----
        bool leaf_term = (cur.meta & F_TERM) != 0;
        if (plen == 0) leaf_term = (root.meta & F_TERM) != 0;
        if (!leaf_term) {
            if (winner == 0 && !full) { winner = G::winner(st); full = G::full(st); }
            if (winner != 0 || full) {
                leaf_term = true;
                const uint32_t tf = F_TERM | (winner == 1 ? F_WIN_P1 : (winner == -1 ? F_WIN_P2 : 0u));
                if (plen == 0) { root_meta = (root_meta & ~(F_WIN_P1 | F_WIN_P2)) | tf; cur.meta = root_meta; }   // (then every descent ends at the root)
                else { cur.meta = (cur.meta & ~(F_WIN_P1 | F_WIN_P2)) | tf; if (lane == 0) arena[last_slot].meta = cur.meta; }
            }
        }
        // ---- random symmetry for non-terminal leaves (BatchedMCTS.h:261-271) ----
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
            L.flags = (uint8_t)(LF_VALID | (plen > 0 ? LF_VLPENDING : 0) | (leaf_term ? LF_TERM : 0) |
                                ((tflags & AZ_LEAF_P1_WINS) ? LF_WIN_P1 : 0) | ((tflags & AZ_LEAF_P2_WINS) ? LF_WIN_P2 : 0));
            L.path_len = plen; L.sym = (uint32_t)sym;
            st32(&(d.leaf_vl + (size_t)env * d.kcap + k)->h, L);
            az_leaf P;     // handed to the evaluator
            P.bb0 = ex.bb[0]; P.bb1 = ex.bb[1]; P.turn = (int8_t)st.turn; P.flags = tflags; P.sym = (uint8_t)sym; P.passes = (uint8_t)st.passes;
            P.reserved[0] = P.reserved[1] = P.reserved[2] = 0;
            st32(leaves + (size_t)env * K + k, P);
        }
    }
    if ((threadIdx.x & 31) == 0 && root_meta != root.meta) tr->root.meta = root_meta;
    if (d.stats) {
        const unsigned dep = __reduce_add_sync(FULL, (mine && lane == 0) ? (unsigned)plen : 0u);
        const unsigned edg = __reduce_add_sync(FULL, (mine && lane == 0) ? (unsigned)st_edges : 0u);
        if ((threadIdx.x & 31) == 0) { atomicAdd(d.stats + 0, (unsigned long long)K); atomicAdd(d.stats + 1, (unsigned long long)dep); atomicAdd(d.stats + 2, (unsigned long long)edg); }
    }
}

}  // namespace az
