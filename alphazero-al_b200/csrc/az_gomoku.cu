// Gomoku environment: host-side single-game API (mirror of the reference's env_cpp.gomoku.Env) and lockstep device
// kernels over az_gomoku[n] records.  Env-only, like the reference (no MCTS engine is registered for Gomoku,
// src/cpp/mcts_bindings.cpp:393-394).
//   reference: src/cpp/Gomoku.h:11-296 (byte board in a std::vector), src/cpp/env_gomoku.h:60-171
// Here a position is two sets of row bit masks; the game logic below is written once against a row accessor, so the
// host functions (pinned on the CPU against the compiled reference) and the kernels run the same code.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/azb200_gomoku.h"
#include "az_games.cuh"
#include "az_rng.cuh"

namespace az {

AZ_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
AZ_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

struct GmkMeta { int size, k, turn, n_pieces, last_action, last_player, winner, done; };

// rows of one game inside an az_gomoku record (host memory, or HBM for the lockstep step kernel)
struct RowsRec {
    uint32_t (*rows)[AZ_GOMOKU_MAX_SIZE];
    AZ_HD uint32_t get(int p, int r) const { return rows[p][r]; }
    AZ_HD void set(int p, int r, uint32_t v) { rows[p][r] = v; }
};
// rows of one game in shared memory, word (p, r) of thread t at base[(p*size + r)*stride + t]: bank = thread
struct RowsShared {
    uint32_t *base; int size, stride;
    AZ_HD uint32_t get(int p, int r) const { return base[(p * size + r) * stride]; }
    AZ_HD void set(int p, int r, uint32_t v) { base[(p * size + r) * stride] = v; }
};

AZ_HD int gmk_pidx(int player) { return player == 1 ? 0 : 1; }
template <class R> AZ_HD int gmk_cell(const R &b, int r, int c) {
    return (int)((b.get(0, r) >> c) & 1u) - (int)((b.get(1, r) >> c) & 1u);
}
// count_direction: Gomoku.h:232-245 (whole run of `player` stones starting next to (row, col))
template <class R> AZ_HD int gmk_count_dir(const R &b, int size, int row, int col, int dr, int dc, int p) {
    int count = 0, r = row + dr, c = col + dc;
    while (r >= 0 && r < size && c >= 0 && c < size && ((b.get(p, r) >> c) & 1u)) { ++count; r += dr; c += dc; }
    return count;
}
// has_line_from: Gomoku.h:247-263 (directions in the reference's order: vertical, horizontal, diagonal, anti-diagonal).
// (Tried: the 2k-1 cells of each line gathered into bit vectors and a shift-AND run test, no data-dependent loops - bit-exact,
// but 5.42 vs 4.03 ms per 1 M games of 15 x 15: most rays stop at their first probe, the fixed 2k-1 row gathers cost more.)
template <class R> AZ_HD bool gmk_has_line_rc(const R &b, int size, int k, int row, int col, int player) {
    const int p = gmk_pidx(player);
    const int DR[4] = {1, 0, 1, 1}, DC[4] = {0, 1, 1, -1};
    for (int i = 0; i < 4; ++i) {
        const int f = gmk_count_dir(b, size, row, col, DR[i], DC[i], p);
        const int bw = gmk_count_dir(b, size, row, col, -DR[i], -DC[i], p);
        if (1 + f + bw >= k) return true;
    }
    return false;
}
template <class R> AZ_HD bool gmk_has_line_from(const R &b, int size, int k, int action, int player) {
    return gmk_has_line_rc(b, size, k, action / size, action % size, player);
}
// step: Gomoku.h:63-92 (validated, unlike Connect4 / Othello); returns 0 or the status of the exception the reference throws
// (row, col) = action / size, action % size, passed in by callers that know them (runtime divisions are ~30 instructions each)
template <class R> AZ_HD int gmk_step_rc(R &b, GmkMeta &m, int action, int row, int col) {
    if (m.done) return AZ_GOMOKU_FINISHED;
    if (action < 0 || action >= m.size * m.size) return AZ_GOMOKU_OUT_OF_RANGE;
    if (((b.get(0, row) | b.get(1, row)) >> col) & 1u) return AZ_GOMOKU_OCCUPIED;
    const int p = gmk_pidx(m.turn);
    b.set(p, row, b.get(p, row) | (1u << col));
    m.n_pieces++;
    m.last_action = action;
    m.last_player = m.turn;
    if (gmk_has_line_rc(b, m.size, m.k, row, col, m.last_player)) { m.winner = m.last_player; m.done = 1; }
    else if (m.n_pieces == m.size * m.size) { m.winner = 0; m.done = 1; }
    m.turn = -m.turn;
    return 0;
}
template <class R> AZ_HD int gmk_step(R &b, GmkMeta &m, int action) {
    if (action < 0 || action >= m.size * m.size) return m.done ? AZ_GOMOKU_FINISHED : AZ_GOMOKU_OUT_OF_RANGE;
    return gmk_step_rc(b, m, action, action / m.size, action % m.size);
}
// the idx-th empty cell in ascending action order (get_valid_moves()[idx], Gomoku.h:99-107) without building the list
template <class R> AZ_HD int gmk_nth_empty(const R &b, int size, int idx, int &row, int &col) {
    const uint32_t full = size == 32 ? 0xFFFFFFFFu : ((1u << size) - 1u);
    for (int r = 0; r < size; ++r) {
        uint32_t e = ~(b.get(0, r) | b.get(1, r)) & full;
        const int c = popc32(e);
        if (idx < c) {
            for (int i = 0; i < idx; ++i) e &= e - 1;
            row = r; col = ctz32(e);
            return r * size + col;
        }
        idx -= c;
    }
    row = col = -1;
    return -1;
}
// rollout policy: index of the move among the n legal ones = high half of the 64-bit hash scaled to [0, n) (one wide
// multiply; a 64-bit modulo by a runtime value costs ~150 instructions on the device).  Twin: orc_gmk_pick.
AZ_HD int gmk_pick(uint64_t h, int n) { return (int)(((h >> 32) * (uint64_t)(uint32_t)n) >> 32); }
// transform_coord: Gomoku.h:276-294 (full D4 group of the square board)
AZ_HD void gmk_xform(int sym, int n, int r, int c, int &nr, int &nc) {
    switch (sym) {
        case 1: nr = c; nc = n - 1 - r; break;
        case 2: nr = n - 1 - r; nc = n - 1 - c; break;
        case 3: nr = n - 1 - c; nc = r; break;
        case 4: nr = r; nc = n - 1 - c; break;
        case 5: nr = n - 1 - r; nc = c; break;
        case 6: nr = c; nc = r; break;
        case 7: nr = n - 1 - c; nc = n - 1 - r; break;
        default: nr = r; nc = c; break;
    }
}
template <class R> AZ_HD uint64_t gmk_digest(const R &b, const GmkMeta &m, int plies) {
    uint64_t d = 0x9E3779B97F4A7C15ULL;
    for (int r = 0; r < m.size; ++r) d = splitmix64(d ^ (((uint64_t)b.get(1, r) << 32) | (uint64_t)b.get(0, r)));
    return splitmix64(d ^ (uint64_t)(uint32_t)(m.winner + 1) ^ ((uint64_t)plies << 8) ^ ((uint64_t)(uint32_t)(m.turn + 1) << 20));
}

AZ_HD GmkMeta gmk_meta(const az_gomoku *s) {
    GmkMeta m;
    m.size = s->size; m.k = s->n_in_row; m.turn = s->turn; m.n_pieces = s->n_pieces; m.last_action = s->last_action;
    m.last_player = s->last_player; m.winner = s->winner; m.done = s->done;
    return m;
}
AZ_HD void gmk_store_meta(az_gomoku *s, const GmkMeta &m) {
    s->size = m.size; s->n_in_row = m.k; s->turn = m.turn; s->n_pieces = m.n_pieces; s->last_action = m.last_action;
    s->last_player = m.last_player; s->winner = m.winner; s->done = m.done;
}
AZ_HD void gmk_reset_meta(GmkMeta &m) {            // reset: Gomoku.h:30-39
    m.turn = 1; m.n_pieces = 0; m.last_action = -1; m.last_player = 0; m.winner = 0; m.done = 0;
}
// apply_symmetry on a record: Gomoku.h:130-158 (stones scattered to their images, last action follows)
AZ_HD void gmk_symmetry(az_gomoku *s, int sym) {
    if (sym == 0) return;
    const int n = s->size;
    uint32_t nw[2][AZ_GOMOKU_MAX_SIZE];
    for (int p = 0; p < 2; ++p) for (int r = 0; r < n; ++r) nw[p][r] = 0;
    for (int p = 0; p < 2; ++p)
        for (int r = 0; r < n; ++r)
            for (uint32_t w = s->rows[p][r]; w; w &= w - 1) {
                int nr, nc; gmk_xform(sym, n, r, ctz32(w), nr, nc);
                nw[p][nr] |= 1u << nc;
            }
    for (int p = 0; p < 2; ++p) for (int r = 0; r < n; ++r) s->rows[p][r] = nw[p][r];
    if (s->last_action >= 0) {
        int nr, nc; gmk_xform(sym, n, s->last_action / n, s->last_action % n, nr, nc);
        s->last_action = nr * n + nc;
    }
}

// ---- kernels -----------------------------------------------------------------------------------------------------
__global__ void k_gmk_reset(int n, int size, int k, az_gomoku *st) {
    // one thread per 16-byte chunk of the records: fully coalesced stores (a record is 18 chunks)
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const int CH = (int)(sizeof(az_gomoku) / 16);
    if (t >= (size_t)n * CH) return;
    const int j = (int)(t % CH);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (j == CH - 2) v = make_uint4((uint32_t)size, (uint32_t)k, 1u, 0u);               // size, n_in_row, turn = +1, n_pieces
    if (j == CH - 1) v = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);                            // last_action = -1, last_player, winner, done
    reinterpret_cast<uint4 *>(st)[t] = v;
}
__global__ void k_gmk_step(int n, az_gomoku *st, const int32_t *__restrict__ actions, uint8_t *status, int32_t *winners, uint8_t *dones) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    az_gomoku *s = st + i;
    GmkMeta m = gmk_meta(s);
    const int a = actions[i];
    int rc = 0;
    if (a >= 0 && !m.done) {
        RowsRec b{s->rows};
        rc = gmk_step(b, m, a);
        if (rc == 0) gmk_store_meta(s, m);
    }
    if (status) status[i] = (uint8_t)rc;
    if (winners) winners[i] = m.winner;
    if (dones) dones[i] = (uint8_t)m.done;
}
// one thread per 16 consecutive cells of the flat [n, S*S] outputs (a chunk may straddle two games): 128-bit stores for
// boards and masks, each row mask loaded once per row; thread t < n also writes the per-game scalars
__global__ void k_gmk_observe(int n, int size, const az_gomoku *__restrict__ st, int8_t *boards, uint8_t *masks, int32_t *turns,
                              int32_t *winners, uint8_t *dones) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const int S = size * size;
    const size_t total = (size_t)n * S;
    if (t < (size_t)n) {
        const az_gomoku *s = st + t;
        if (turns) turns[t] = s->turn;
        if (winners) winners[t] = s->winner;
        if (dones) dones[t] = (uint8_t)s->done;
    }
    const size_t idx = t * 16;
    if (idx >= total || (!boards && !masks)) return;
    size_t g = idx / S;
    const int j = (int)(idx - g * S);
    int r = j / size, c = j % size;
    uint32_t w0 = st[g].rows[0][r], w1 = st[g].rows[1][r];
    const int cnt = total - idx < 16 ? (int)(total - idx) : 16;
    uint32_t bw[4] = {0, 0, 0, 0}, mw[4] = {0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        if (q < cnt) {
            const int v = (int)((w0 >> c) & 1u) - (int)((w1 >> c) & 1u);
            bw[q >> 2] |= ((uint32_t)v & 0xFFu) << ((q & 3) * 8);
            mw[q >> 2] |= (v == 0 ? 1u : 0u) << ((q & 3) * 8);
            if (++c == size) {
                c = 0;
                if (++r == size) { r = 0; ++g; }
                if (q + 1 < cnt) { w0 = st[g].rows[0][r]; w1 = st[g].rows[1][r]; }
            }
        }
    }
    if (cnt == 16) {
        if (boards) *reinterpret_cast<uint4 *>(boards + idx) = make_uint4(bw[0], bw[1], bw[2], bw[3]);
        if (masks) *reinterpret_cast<uint4 *>(masks + idx) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
    } else {
        for (int q = 0; q < cnt; ++q) {
            if (boards) boards[idx + q] = (int8_t)((bw[q >> 2] >> ((q & 3) * 8)) & 0xFFu);
            if (masks) masks[idx + q] = (uint8_t)((mw[q >> 2] >> ((q & 3) * 8)) & 0xFFu);
        }
    }
}
__global__ void k_gmk_symmetry(int n, az_gomoku *st, const int32_t *__restrict__ syms) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int sym = syms[i];
    if (sym <= 0 || sym >= 8) return;
    gmk_symmetry(st + i, sym);
}
// one thread per game; the row masks live in shared memory for the whole game (2*size words per thread, bank = thread).
// (Tried: a block-local queue from which a lane whose game is over takes the next one, 1..8 games per lane - 4.69..5.2 ms
// against 4.03 ms for 1 M games: the warp instructions go to the divergent row / ray walks inside a ply, not to idle tails.)
__global__ void k_gmk_rollout(int n, int size, int k, uint64_t seed, uint64_t first, uint64_t *digest, int32_t *plies, int nrec,
                              int8_t *rb, int32_t *rt, int32_t *ra, int32_t *rw, uint8_t *rd, az_gomoku *fin) {
    extern __shared__ uint32_t sh_rows[];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RowsShared b{sh_rows + threadIdx.x, size, (int)blockDim.x};
    for (int w = 0; w < 2 * size; ++w) sh_rows[w * blockDim.x + threadIdx.x] = 0;
    const uint64_t g = first + (uint64_t)i;
    const int S = size * size;
    GmkMeta m; m.size = size; m.k = k; gmk_reset_meta(m);
    const bool rec = i < nrec;
    int ply = 0;
    while (!m.done) {
        int row, col;
        const int a = gmk_nth_empty(b, size, gmk_pick(rollout_hash(seed, g, (uint64_t)ply), S - m.n_pieces), row, col);
        const size_t o = (size_t)i * S + ply;
        if (rec) {
            for (int j = 0; j < S; ++j) rb[o * S + j] = (int8_t)gmk_cell(b, j / size, j % size);
            rt[o] = m.turn; ra[o] = a;
        }
        gmk_step_rc(b, m, a, row, col);
        if (rec) { rw[o] = m.winner; rd[o] = (uint8_t)m.done; }
        ++ply;
    }
    if (digest) digest[i] = gmk_digest(b, m, ply);
    if (plies) plies[i] = ply;
    if (fin) {
        az_gomoku *s = fin + i;
        for (int p = 0; p < 2; ++p) for (int r = 0; r < AZ_GOMOKU_MAX_SIZE; ++r) s->rows[p][r] = r < size ? b.get(p, r) : 0u;
        gmk_store_meta(s, m);
    }
}

static int gmk_check_params(int size, int k) {      // validate_config: Gomoku.h:214-222, then this implementation's limit
    if (size <= 0) return AZ_GOMOKU_BAD_SIZE;
    if (k <= 1) return AZ_GOMOKU_BAD_N_LOW;
    if (k > size) return AZ_GOMOKU_BAD_N_HIGH;
    if (size > AZ_GOMOKU_MAX_SIZE) return AZ_ERR_GOMOKU_SIZE;
    return 0;
}

}  // namespace az

using namespace az;

extern "C" {

int az_gomoku_set_params(az_gomoku *s, int size, int k) {
    const int rc = gmk_check_params(size, k);
    if (rc) return rc;
    memset(s, 0, sizeof(*s));
    s->size = size; s->n_in_row = k;
    az_gomoku_reset(s);
    return 0;
}
void az_gomoku_reset(az_gomoku *s) {
    memset(s->rows, 0, sizeof(s->rows));
    GmkMeta m = gmk_meta(s); gmk_reset_meta(m); gmk_store_meta(s, m);
}
int az_gomoku_import(az_gomoku *s, const int8_t *board) {          // import_board + sync_from_board: Gomoku.h:57-61,160-204
    const int n = s->size, S = n * n;
    memset(s->rows, 0, sizeof(s->rows));
    int p1 = 0, p2 = 0;
    GmkMeta m = gmk_meta(s); gmk_reset_meta(m);
    for (int i = 0; i < S; ++i) {
        const int v = board[i];
        if (v == 1) { ++p1; s->rows[0][i / n] |= 1u << (i % n); m.last_action = i; m.last_player = 1; }
        else if (v == -1) { ++p2; s->rows[1][i / n] |= 1u << (i % n); m.last_action = i; m.last_player = -1; }
        else if (v != 0) { gmk_store_meta(s, m); return AZ_GOMOKU_BAD_CELL; }
    }
    m.n_pieces = p1 + p2;
    if (p1 == p2) m.turn = 1;
    else if (p1 == p2 + 1) m.turn = -1;
    else m.turn = (m.n_pieces % 2 == 0) ? 1 : -1;
    RowsRec b{s->rows};
    m.winner = 0;                                                   // find_winner_full_scan: Gomoku.h:265-274
    for (int i = 0; i < S; ++i) {
        const int v = gmk_cell(b, i / n, i % n);
        if (v != 0 && gmk_has_line_from(b, n, m.k, i, v)) { m.winner = v; break; }
    }
    m.done = (m.winner != 0) || (m.n_pieces == S);
    gmk_store_meta(s, m);
    return 0;
}
void az_gomoku_export(const az_gomoku *s, int8_t *board) {
    const int n = s->size;
    for (int i = 0; i < n * n; ++i) board[i] = (int8_t)((int)((s->rows[0][i / n] >> (i % n)) & 1u) - (int)((s->rows[1][i / n] >> (i % n)) & 1u));
}
int az_gomoku_step(az_gomoku *s, int action) {
    GmkMeta m = gmk_meta(s);
    RowsRec b{s->rows};
    const int rc = gmk_step(b, m, action);
    if (rc == 0) gmk_store_meta(s, m);
    return rc;
}
int az_gomoku_valid_moves(const az_gomoku *s, int32_t *moves) {
    const int n = s->size;
    int cnt = 0;
    for (int i = 0; i < n * n; ++i)
        if ((((s->rows[0][i / n] | s->rows[1][i / n]) >> (i % n)) & 1u) == 0) moves[cnt++] = i;
    return cnt;
}
int az_gomoku_apply_symmetry(az_gomoku *s, int sym) {
    if (sym < 0 || sym >= 8) return AZ_GOMOKU_BAD_SYM;
    gmk_symmetry(s, sym);
    return 0;
}
int az_gomoku_inverse_symmetry_action(int size, int sym, int action) {   // Gomoku.h:115-128 (applies transform_coord(sym_id))
    if (action < 0 || action >= size * size) return -AZ_GOMOKU_OUT_OF_RANGE;
    if (sym < 0 || sym >= 8) return -AZ_GOMOKU_BAD_SYM;
    int nr, nc; gmk_xform(sym, size, action / size, action % size, nr, nc);
    return nr * size + nc;
}

#define AZ_GMK_DONE() return cudaGetLastError() == cudaSuccess ? AZ_OK : AZ_ERR_CUDA

int az_gomoku_reset_dev(int n, int size, int k, az_gomoku *st, void *stream) {
    if (gmk_check_params(size, k)) return AZ_ERR_INVALID;
    if (n <= 0) return AZ_OK;
    const size_t chunks = (size_t)n * (sizeof(az_gomoku) / 16);
    k_gmk_reset<<<(unsigned)((chunks + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, size, k, st);
    AZ_GMK_DONE();
}
int az_gomoku_step_dev(int n, az_gomoku *st, const int32_t *actions, uint8_t *status, int32_t *winners, uint8_t *dones, void *stream) {
    if (n <= 0) return AZ_OK;
    k_gmk_step<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, st, actions, status, winners, dones);
    AZ_GMK_DONE();
}
int az_gomoku_observe_dev(int n, int size, const az_gomoku *st, int8_t *boards, uint8_t *masks, int32_t *turns, int32_t *winners,
                          uint8_t *dones, void *stream) {
    if (size <= 0 || size > AZ_GOMOKU_MAX_SIZE) return AZ_ERR_INVALID;
    if (n <= 0) return AZ_OK;
    if ((((uintptr_t)boards) | ((uintptr_t)masks)) & 15u) return AZ_ERR_INVALID;      /* 128-bit stores */
    const size_t chunks = ((size_t)n * size * size + 15) / 16;
    const size_t threads = chunks > (size_t)n ? chunks : (size_t)n;
    k_gmk_observe<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, size, st, boards, masks, turns, winners, dones);
    AZ_GMK_DONE();
}
int az_gomoku_symmetry_dev(int n, az_gomoku *st, const int32_t *syms, void *stream) {
    if (n <= 0) return AZ_OK;
    k_gmk_symmetry<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(n, st, syms);
    AZ_GMK_DONE();
}
int az_gomoku_rollout_dev(int n, int size, int k, uint64_t seed, uint64_t first, uint64_t *digest, int32_t *plies, int nrec, int8_t *rb,
                          int32_t *rt, int32_t *ra, int32_t *rw, uint8_t *rd, az_gomoku *fin, void *stream) {
    if (gmk_check_params(size, k)) return AZ_ERR_INVALID;
    if (n <= 0) return AZ_OK;
    const int bs = 128;
    const size_t smem = (size_t)2 * size * bs * sizeof(uint32_t);
    k_gmk_rollout<<<(n + bs - 1) / bs, bs, smem, (cudaStream_t)stream>>>(n, size, k, seed, first, digest, plies, nrec, rb, rt, ra, rw, rd, fin);
    AZ_GMK_DONE();
}

}  // extern "C"
