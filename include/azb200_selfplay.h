/*
 * azb200_selfplay.h - on-device self-play driver and the trajectory format of the one exchange step
 * of the path.
 *
 * Replaces the per-ply Python loops of src/game.py:65-164 (Game.batch_self_play) and src/player.py:333-375
 * (AlphaZeroPlayer.get_batch_action) with two kernels per ply: visit counts -> policy target, temperature sampling,
 * trajectory recording, env step; and, for the games that ended, a move of their trajectory into an output ring (the slot
 * restarts immediately so the batch never idles).
 *
 * A finished game travels as COMPACT records: one 32-byte az_sp_game header plus one position record per recorded position
 * (bitboards, side to move, root WDL, policy target: 64 bytes for Connect4, 320 for Othello; the terminal position included).
 * Everything else of the reference's training tuple (src/game.py:114-157, src/ReplayBuffer.py:12-19) - the int8 planes, legal
 * masks, winner_z, steps_to_end, aux target, future_root_wdl shifted by td_steps, the terminal tuple - is a function of those
 * and is produced on the receiving side by az_selfplay_expand_dev, straight into replay-buffer tensors.  That is what the NCCL
 * all-gather moves: ~1.4 KB per 22-position Connect4 game instead of 8.2 KB of padded tuples.
 */
#ifndef AZB200_SELFPLAY_H
#define AZB200_SELFPLAY_H
#include "azb200.h"

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

typedef struct az_sp_game {              /* 32 bytes: one finished game */
    uint64_t uid;                        /* global game id (slot's first game: uid_base + slot; then += uid_stride per restart) */
    int64_t pos_start;                   /* row of the game's first position in the position array the header travels with */
    int32_t length;                      /* recorded positions, the terminal one included (= plies + 1) */
    int32_t winner;                      /* env.winPlayer() at the end: +1 / -1 / 0 */
    int32_t reserved[2];
} az_sp_game;

typedef struct az_sp_pos {               /* head of a position record; float prob[A] follows, record padded to 32 bytes */
    uint64_t bb0, bb1;                   /* stones of player +1 / -1 (engine bit layout, az_root) */
    float root_wdl[3];                   /* root_D, root_P1W, root_P2W after the search (src/player.py:342); 0 at the terminal position */
    int8_t turn;                         /* side to move */
    uint8_t passes;                      /* Othello consecutive passes */
    uint8_t reserved[2];
} az_sp_pos;
/* bytes of one position record: 32 + 4*A rounded up to 32 (Connect4 64, Othello 320); < 0 for an unknown game */
int az_selfplay_pos_bytes(int game);
/* longest game in plies (Connect4 42, Othello 128): a slot's staging holds that many records */
int az_selfplay_max_plies(int game);

typedef struct az_selfplay {
    /* configuration */
    int32_t game, n, max_plies, pos_bytes;
    int32_t temp_decay_moves;            /* src/game.py:54-63 */
    float temp_init, temp_endgame;
    int32_t forced_games;                /* rows of `forced` */
    uint64_t seed;
    uint64_t uid_stride;                 /* a slot's next game gets uid += uid_stride (= total slots over all ranks) */
    uint64_t forced_uid0;                /* game uid of row 0 of `forced` */
    /* per-slot state (device) */
    az_root *states;                     /* [n] env states = search roots */
    int32_t *steps;                      /* [n] plies played in the current game */
    uint64_t *uids;                      /* [n] global id of the game in this slot */
    uint8_t *st_pos;                     /* [n][max_plies][pos_bytes] staging of the running trajectories */
    /* per-ply scratch (device) */
    int32_t *actions;                    /* [n] action played (written by ply; -1 = tree must reset) */
    int32_t *fin_list;                   /* [n] slots whose game ended at this ply ... */
    int32_t *fin_count;                  /* ... and how many (zeroed by az_selfplay_ply_dev) */
    /* opening script (optional, NULL = none): int8[forced_games][max_plies]; game uid u in [forced_uid0, forced_uid0 + forced_games)
     * plays forced[u - forced_uid0][t] at ply t when that entry is >= 0 (tests replay recorded games; openings books) */
    const int8_t *forced;
    /* output ring (device) */
    az_sp_game *out_games;               /* [game_capacity] */
    uint8_t *out_pos;                    /* [game_capacity * (max_plies + 1)][pos_bytes] */
    unsigned long long *out_counters;    /* [0] games written, [1] positions written, [2] games dropped (ring full), [3] plies of written games */
    int32_t game_capacity, reserved;
} az_selfplay;

/* one ply for every slot: counts int32[n,A] and root_stats f32[n,6+8A] come from az_mcts_get_counts_dev /
 * az_mcts_get_root_stats_dev; writes sp->actions (hand them to az_mcts_prune_roots_dev) and the list of finished slots. */
int az_selfplay_ply_dev(const az_selfplay *sp, const int32_t *d_counts, const float *d_root_stats, void *stream);
/* for every finished slot: move its trajectory (+ the terminal position) into the output ring, then restart the slot
 * (env reset, steps = 0, uid += stride).  One CTA per finished slot. */
int az_selfplay_flush_dev(const az_selfplay *sp, void *stream);

/* Compact records -> the reference's training tuples as replay-buffer tensors (src/ReplayBuffer.py:12-19), row r = position r of
 * `d_pos` (games[i] owns rows [pos_start, pos_start + length)): state int8[P,3,R,C], prob f32[P,A], winner int8[P],
 * steps_to_end int16[P], aux_target int16[P], root_wdl f32[P,3], valid_mask u8[P,A] (0/1), future_root_wdl f32[P,3]
 * (src/game.py:114-157: steps_to_end T..1, aux = steps_to_end for Connect4 / final disc difference x side to move for Othello,
 * future_root_wdl[t] = root_wdl[t + td_steps] while t + td_steps < T else 0, terminal tuple = end state, zero policy, all-ones
 * mask).  Any output pointer may be NULL. */
int az_selfplay_expand_dev(int game, int n_games, const az_sp_game *d_games, const uint8_t *d_pos, int td_steps, int8_t *d_state,
                           float *d_prob, int8_t *d_winner, int16_t *d_steps_to_end, int16_t *d_aux, float *d_root_wdl,
                           uint8_t *d_valid_mask, float *d_future_root_wdl, void *stream);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* AZB200_SELFPLAY_H */
