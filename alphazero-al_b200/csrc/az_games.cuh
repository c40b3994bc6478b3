// Bitboard game logic shared by the CUDA kernels and the host-side Env mirror.
//
// Semantics follow the reference's game classes (citations are relative to /root/reference/):
//   Connect4  src/cpp/Connect4.h:31-295   - 2 x u64, 7 bits per column (6 rows + sentinel)
//   Othello   src/cpp/Othello.h:28-388    - 2 x u64, bit i = (row i/8, col i%8), action 64 = pass
// but the state is re-designed for registers: no byte board, no height array (column heights are popcounts
// of the occupancy), 16 bytes of bitboards + three small ints.  The byte board the Python API exposes is
// materialised only at the boundary (export_cell / import helpers).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AZ_HD __host__ __device__ __forceinline__
#else
#define AZ_HD inline
#endif

namespace az {

AZ_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
AZ_HD int ctz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}

enum { GAME_C4 = 0, GAME_OTH = 1 };

// Register-resident state common to both games.
struct State {
    uint64_t bb[2];   // bb[0] = player +1, bb[1] = player -1
    int turn;         // side to move (+1 / -1)
    int passes;       // Othello consecutive passes
    int last;         // index (0/1) of the player who moved last, -1 = nobody
};

// ------------------------------------------------------------------------------------------------
// Connect4
// ------------------------------------------------------------------------------------------------
struct C4 {
    static constexpr int GAME = GAME_C4;
    static constexpr int ROWS = 6, COLS = 7;
    static constexpr int A = 7;          // ACTION_SIZE
    static constexpr int S = 42;         // BOARD_SIZE
    static constexpr int NUM_SYM = 2;
    static constexpr int MAX_EDGES = 7;
    static constexpr int MAX_DEPTH = 44; // 42 plies + slack
    static constexpr int MAX_PLIES = 42; // longest possible game
    static constexpr bool AUX_PLUS_ONE = true, AUX_NEGATE = false;   // Connect4.h:34-35

    AZ_HD static void reset(State &s) { s.bb[0] = s.bb[1] = 0; s.turn = 1; s.passes = 0; s.last = -1; }   // :62-72
    AZ_HD static int col_height(uint64_t occ, int c) { return popc64((occ >> (c * 7)) & 0x3FULL); }
    AZ_HD static int n_pieces(const State &s) { return popc64(s.bb[0] | s.bb[1]); }
    // import_board / sync_from_board (:87-129): the last mover is inferred from piece-count parity.
    AZ_HD static void finish_import(State &s, int turn) {
        int n = n_pieces(s);
        s.last = n > 0 ? ((n & 1) ? 0 : 1) : -1;
        s.turn = turn; s.passes = 0;
    }
    // (bb[] is only ever indexed with constants: a run-time index would move the boards to local memory in a kernel)
    AZ_HD static void step(State &s, int col) {                       // :159-172 (no legality check)
        uint64_t occ = s.bb[0] | s.bb[1];
        const bool p1 = s.turn == 1;
        const uint64_t bit = 1ULL << (col * 7 + col_height(occ, col));
        s.bb[0] |= p1 ? bit : 0ULL; s.bb[1] |= p1 ? 0ULL : bit;
        s.last = p1 ? 0 : 1; s.turn = -s.turn;
    }
    AZ_HD static int winner(const State &s) {                         // :182-203 (last mover only)
        if (s.last < 0) return 0;
        uint64_t b = s.last == 0 ? s.bb[0] : s.bb[1], t;
        int res = s.last == 0 ? 1 : -1;
        t = b & (b >> 1); if (t & (t >> 2))  return res;
        t = b & (b >> 7); if (t & (t >> 14)) return res;
        t = b & (b >> 6); if (t & (t >> 12)) return res;
        t = b & (b >> 8); if (t & (t >> 16)) return res;
        return 0;
    }
    AZ_HD static bool full(const State &s) { return n_pieces(s) == 42; }   // :221-224
    AZ_HD static bool done(const State &s) { return winner(s) != 0 || full(s); }   // env_connect4.h:43-44
    // legal moves as a bit mask over actions (bit c = column c not full), ascending = edge order (:209-218)
    AZ_HD static uint64_t legal(const State &s) {
        uint64_t occ = s.bb[0] | s.bb[1], top = occ >> 5;               // bit 7c+5 = top row of column c
        uint64_t m = 0;
#pragma unroll
        for (int c = 0; c < 7; ++c) m |= ((~top >> (c * 7)) & 1ULL) << c;
        return m;
    }
    AZ_HD static uint64_t flip_bb(uint64_t b) {                        // :249-268
        uint64_t d = 0;
#pragma unroll
        for (int c = 0; c < 7; ++c) d |= ((b >> (c * 7)) & 0x7FULL) << ((6 - c) * 7);
        return d;
    }
    AZ_HD static void symmetry(State &s, int sym) { if (sym) { s.bb[0] = flip_bb(s.bb[0]); s.bb[1] = flip_bb(s.bb[1]); } }
    // map an action of the ORIGINAL frame to its index in the symmetrised frame (self-inverse, :288-294)
    AZ_HD static int sym_action(int sym, int a) { return sym ? 6 - a : a; }
    // value of row-major cell j of the byte board (sync_to_board, :135-150)
    AZ_HD static int cell(const State &s, int j) {
        int r = j / 7, c = j - r * 7, bit = c * 7 + (5 - r);
        return (int)((s.bb[0] >> bit) & 1ULL) - (int)((s.bb[1] >> bit) & 1ULL);
    }
    AZ_HD static int cell_bit(int j) { int r = j / 7, c = j - r * 7; return c * 7 + (5 - r); }
    AZ_HD static float terminal_aux_dummy() { return 0.0f; }           // :226-229
};

// ------------------------------------------------------------------------------------------------
// Othello
// ------------------------------------------------------------------------------------------------
struct Oth {
    static constexpr int GAME = GAME_OTH;
    static constexpr int ROWS = 8, COLS = 8;
    static constexpr int A = 65;
    static constexpr int S = 64;
    static constexpr int NUM_SYM = 8;
    static constexpr int PASS = 64;
    static constexpr int MAX_EDGES = 48;  // known maximum mobility is 33; three 16-lane passes cover 48
    static constexpr int MAX_DEPTH = 128; // <= 60 placements + interleaved single passes
    static constexpr int MAX_PLIES = 128;
    static constexpr bool AUX_PLUS_ONE = false, AUX_NEGATE = true;     // Othello.h:31-32
    static constexpr uint64_t NOT_A = 0xFEFEFEFEFEFEFEFEULL, NOT_H = 0x7F7F7F7F7F7F7F7FULL;

    AZ_HD static void reset(State &s) {                                // :62-75
        s.bb[0] = (1ULL << 28) | (1ULL << 35); s.bb[1] = (1ULL << 27) | (1ULL << 36);
        s.turn = 1; s.passes = 0; s.last = -1;
    }
    AZ_HD static int n_pieces(const State &s) { return popc64(s.bb[0]) + popc64(s.bb[1]); }
    AZ_HD static void finish_import(State &s, int turn) { s.turn = turn; s.passes = 0; s.last = -1; }   // :92-111
    template <int D> AZ_HD static uint64_t shift(uint64_t b) {         // :133-147
        if (D == 0) return b >> 8;
        if (D == 1) return (b >> 7) & NOT_A;
        if (D == 2) return (b << 1) & NOT_A;
        if (D == 3) return (b << 9) & NOT_A;
        if (D == 4) return b << 8;
        if (D == 5) return (b << 7) & NOT_H;
        if (D == 6) return (b >> 1) & NOT_H;
        return (b >> 9) & NOT_H;
    }
    template <int D> AZ_HD static uint64_t valid_dir(uint64_t own, uint64_t opp, uint64_t empty) {
        uint64_t c = shift<D>(own) & opp;
#pragma unroll
        for (int i = 0; i < 5; ++i) c |= shift<D>(c) & opp;
        return shift<D>(c) & empty;
    }
    AZ_HD static uint64_t valid_positions(const State &s) {             // :155-171
        const bool p1 = s.turn == 1;
        uint64_t own = p1 ? s.bb[0] : s.bb[1], opp = p1 ? s.bb[1] : s.bb[0], empty = ~(own | opp);
        return valid_dir<0>(own, opp, empty) | valid_dir<1>(own, opp, empty) | valid_dir<2>(own, opp, empty) |
               valid_dir<3>(own, opp, empty) | valid_dir<4>(own, opp, empty) | valid_dir<5>(own, opp, empty) |
               valid_dir<6>(own, opp, empty) | valid_dir<7>(own, opp, empty);
    }
    template <int D> AZ_HD static uint64_t flips_dir(uint64_t placed, uint64_t own, uint64_t opp) {
        uint64_t cand = 0, sq = shift<D>(placed);
#pragma unroll
        for (int i = 0; i < 6; ++i) {          // a run of opponent discs is at most 6 long
            uint64_t hit = sq & opp;
            cand |= hit;
            sq = hit ? shift<D>(sq) : sq;
        }
        return (sq & own) ? cand : 0;
    }
    AZ_HD static uint64_t flips(const State &s, int pos) {              // :177-198
        const bool p1 = s.turn == 1;
        uint64_t own = p1 ? s.bb[0] : s.bb[1], opp = p1 ? s.bb[1] : s.bb[0], pl = 1ULL << pos;
        return flips_dir<0>(pl, own, opp) | flips_dir<1>(pl, own, opp) | flips_dir<2>(pl, own, opp) |
               flips_dir<3>(pl, own, opp) | flips_dir<4>(pl, own, opp) | flips_dir<5>(pl, own, opp) |
               flips_dir<6>(pl, own, opp) | flips_dir<7>(pl, own, opp);
    }
    AZ_HD static void step(State &s, int a) {                           // :206-235
        if (a == PASS) { s.passes++; s.turn = -s.turn; return; }
        const bool p1 = s.turn == 1;
        uint64_t f = flips(s, a);
        const uint64_t add = (1ULL << a) | f;
        s.bb[0] = p1 ? (s.bb[0] | add) : (s.bb[0] & ~f);
        s.bb[1] = p1 ? (s.bb[1] & ~f) : (s.bb[1] | add);
        s.passes = 0; s.last = p1 ? 0 : 1; s.turn = -s.turn;
    }
    AZ_HD static bool over(const State &s) { return n_pieces(s) == 64 || s.passes >= 2; }   // :241-244
    AZ_HD static int winner(const State &s) {                           // :250-258
        if (!over(s)) return 0;
        int a = popc64(s.bb[0]), b = popc64(s.bb[1]);
        return a > b ? 1 : (b > a ? -1 : 0);
    }
    AZ_HD static bool full(const State &s) { return over(s); }          // :299-302
    AZ_HD static bool done(const State &s) { return over(s); }          // env_othello.h:43-44
    // legal placements as a mask; pass is reported separately (get_valid_moves, :282-296)
    AZ_HD static uint64_t legal(const State &s) { return over(s) ? 0ULL : valid_positions(s); }
    AZ_HD static void xform(int sym, int r, int c, int &nr, int &nc) {  // :312-326
        switch (sym) {
        case 1: nr = c;     nc = 7 - r; break;
        case 2: nr = 7 - r; nc = 7 - c; break;
        case 3: nr = 7 - c; nc = r;     break;
        case 4: nr = r;     nc = 7 - c; break;
        case 5: nr = 7 - r; nc = c;     break;
        case 6: nr = c;     nc = r;     break;
        case 7: nr = 7 - c; nc = 7 - r; break;
        default: nr = r;    nc = c;     break;
        }
    }
    AZ_HD static int xform_idx(int sym, int i) { int nr, nc; xform(sym, i >> 3, i & 7, nr, nc); return nr * 8 + nc; }
    // transform_bb (:329-341) moves every stone with transform_coord; here the eight D4 maps are composed from three word-parallel
    // primitives on the 8x8 bit matrix (bit 8r + c): rows reversed (byte swap), all bits reversed (= rotate by 180 degrees) and the
    // transpose (three delta swaps) - ~20 instructions instead of ~15 per stone (the leaf symmetry was 19 % of the Othello select).
    AZ_HD static uint64_t bb_flip_rows(uint64_t x) {                    // (r, c) -> (7 - r, c)
#if defined(__CUDA_ARCH__)
        return ((uint64_t)__byte_perm((uint32_t)x, 0, 0x0123) << 32) | (uint64_t)__byte_perm((uint32_t)(x >> 32), 0, 0x0123);
#else
        return __builtin_bswap64(x);
#endif
    }
    AZ_HD static uint64_t bb_rot180(uint64_t x) {                       // (r, c) -> (7 - r, 7 - c): bit i -> bit 63 - i
#if defined(__CUDA_ARCH__)
        return __brevll(x);
#else
        x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
        x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
        x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
        return __builtin_bswap64(x);
#endif
    }
    AZ_HD static uint64_t bb_transpose(uint64_t x) {                    // (r, c) -> (c, r)
        uint64_t t;
        t = 0x0F0F0F0F00000000ULL & (x ^ (x << 28)); x ^= t ^ (t >> 28);
        t = 0x3333000033330000ULL & (x ^ (x << 14)); x ^= t ^ (t >> 14);
        t = 0x5500550055005500ULL & (x ^ (x << 7));  x ^= t ^ (t >> 7);
        return x;
    }
    AZ_HD static uint64_t xform_bb(uint64_t b, int sym) {               // == the stone-by-stone transform_bb with xform() above
        switch (sym) {
        case 1: return bb_flip_rows(bb_rot180(bb_transpose(b)));        // (c, 7 - r): transpose, then columns reversed
        case 2: return bb_rot180(b);                                    // (7 - r, 7 - c)
        case 3: return bb_flip_rows(bb_transpose(b));                   // (7 - c, r)
        case 4: return bb_flip_rows(bb_rot180(b));                      // (r, 7 - c): columns reversed
        case 5: return bb_flip_rows(b);                                 // (7 - r, c)
        case 6: return bb_transpose(b);                                 // (c, r)
        case 7: return bb_rot180(bb_transpose(b));                      // (7 - c, 7 - r)
        default: return b;
        }
    }
    AZ_HD static uint64_t xform_bb_by_stone(uint64_t b, int sym) {      // the definition, kept for the self-check (tests)
        if (sym == 0) return b;
        uint64_t r = 0;
        for (; b; b &= b - 1) r |= 1ULL << xform_idx(sym, ctz64(b));
        return r;
    }
    AZ_HD static void symmetry(State &s, int sym) { if (sym) { s.bb[0] = xform_bb(s.bb[0], sym); s.bb[1] = xform_bb(s.bb[1], sym); } }
    AZ_HD static int inverse_sym(int sym) { return (sym == 1) ? 3 : (sym == 3 ? 1 : sym); }   // :356-361
    // inverse_symmetry_policy (:373-387): restored[T_inv(i)] = given[i]  <=>  restored[a] = given[T_sym(a)]
    // every D4 map is "transpose?, then rows reversed?, then columns reversed?" (bits 0 / 1 / 2 of the table entry below; xform()
    // above is the definition, tests/test_env_host.py and the leaf-symmetry parity tests cover both forms)
    AZ_HD static int sym_action(int sym, int a) {
        if (a == PASS || sym == 0) return a;
        const int f = (0x71243650 >> (4 * sym)) & 7;       // sym 1: T+C (5), 2: R+C (6), 3: T+R (3), 4: C (4), 5: R (2), 6: T (1), 7: T+R+C (7)
        int r = a >> 3, c = a & 7;
        if (f & 1) { const int t = r; r = c; c = t; }
        if (f & 2) r = 7 - r;
        if (f & 4) c = 7 - c;
        return r * 8 + c;
    }
    AZ_HD static int cell(const State &s, int j) { return (int)((s.bb[0] >> j) & 1ULL) - (int)((s.bb[1] >> j) & 1ULL); }
    AZ_HD static int cell_bit(int j) { return j; }
};

#if defined(__CUDACC__)
// ---- Othello on a lane group (W >= 8 lanes per tree) ------------------------------------------------------------------
// flips() and valid_positions() walk the 8 compass directions one after the other (8 x 6 masked shifts, ~480 instructions on
// 64-bit boards); every lane of a tree's group would repeat all of it.  Here lane (l & 7) walks ONE direction, given as data
// (shift amount, left/right, wrap mask - no divergence), and the eight partial masks are OR-reduced with three shuffles.
struct OthDir { int sh; bool left; uint64_t mask; };
__device__ __forceinline__ OthDir oth_dir(int dd) {          // same numbering as Oth::shift<D>
    OthDir D;
    const int q = dd & 3;
    D.sh = q == 0 ? 8 : (q == 1 ? 7 : (q == 2 ? 1 : 9));
    D.left = dd >= 2 && dd <= 5;
    D.mask = q == 0 ? ~0ULL : ((dd >= 1 && dd <= 3) ? Oth::NOT_A : Oth::NOT_H);
    return D;
}
__device__ __forceinline__ uint64_t oth_shift(uint64_t b, const OthDir &D) { return (D.left ? (b << D.sh) : (b >> D.sh)) & D.mask; }
template <int W> __device__ __forceinline__ uint64_t oth_or8(uint64_t v, unsigned gm) {
    v |= __shfl_xor_sync(gm, v, 1, W); v |= __shfl_xor_sync(gm, v, 2, W); v |= __shfl_xor_sync(gm, v, 4, W);
    return v;
}
template <int W> __device__ __forceinline__ uint64_t oth_flips_group(const State &s, int pos, int lane, unsigned gm) {   // Othello.h:177-198
    const bool p1 = s.turn == 1;
    const uint64_t own = p1 ? s.bb[0] : s.bb[1], opp = p1 ? s.bb[1] : s.bb[0];
    const OthDir D = oth_dir(lane & 7);
    uint64_t cand = 0, sq = oth_shift(1ULL << pos, D);
#pragma unroll
    for (int i = 0; i < 6; ++i) { const uint64_t hit = sq & opp; cand |= hit; sq = hit ? oth_shift(sq, D) : sq; }
    return oth_or8<W>((sq & own) ? cand : 0ULL, gm);
}
template <int W> __device__ __forceinline__ uint64_t oth_valid_group(const State &s, int lane, unsigned gm) {             // Othello.h:155-171
    const bool p1 = s.turn == 1;
    const uint64_t own = p1 ? s.bb[0] : s.bb[1], opp = p1 ? s.bb[1] : s.bb[0], empty = ~(own | opp);
    const OthDir D = oth_dir(lane & 7);
    uint64_t c = oth_shift(own, D) & opp;
#pragma unroll
    for (int i = 0; i < 5; ++i) c |= oth_shift(c, D) & opp;
    return oth_or8<W>(oth_shift(c, D) & empty, gm);
}
// Game::step / legal moves for a lane group: group-uniform control flow required (every lane of the group calls it)
template <class G, int W> __device__ __forceinline__ void step_group(State &s, int a, int lane, unsigned gm) {
    if (G::GAME == GAME_OTH && W >= 8) {
        if (a == Oth::PASS) { s.passes++; s.turn = -s.turn; return; }
        const bool p1 = s.turn == 1;
        const uint64_t f = oth_flips_group<W>(s, a, lane, gm), add = (1ULL << a) | f;
        s.bb[0] = p1 ? (s.bb[0] | add) : (s.bb[0] & ~f);
        s.bb[1] = p1 ? (s.bb[1] & ~f) : (s.bb[1] | add);
        s.passes = 0; s.last = p1 ? 0 : 1; s.turn = -s.turn;
    } else G::step(s, a);
}
template <class G, int W> __device__ __forceinline__ uint64_t legal_group(const State &s, int lane, unsigned gm) {
    if (G::GAME == GAME_OTH && W >= 8) return Oth::over(s) ? 0ULL : oth_valid_group<W>(s, lane, gm);
    return G::legal(s);
}
#endif

}  // namespace az
