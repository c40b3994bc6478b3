/*
 * azb200_selfplay.h - on-device self-play driver: replaces the per-ply Python loops of
 * src/game.py:65-164 (Game.batch_self_play) and src/player.py:333-375 (AlphaZeroPlayer.get_batch_action) with two
 * kernels per ply.  Visit counts -> policy target, temperature sampling, trajectory recording, env step, and - when a
 * game ends - construction of the training tuples (winner_z, steps_to_end, aux target, root_wdl, valid_mask,
 * future_root_wdl shifted by td_steps, terminal tuple) straight into a packed per-game record that is what gets
 * all-gathered over NCCL.  Finished games restart immediately so the batch never idles.
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

typedef struct az_selfplay {
    /* configuration */
    int32_t game, n, max_plies;          /* max_plies = T (42 Connect4, 128 Othello); records hold T+1 positions */
    int32_t td_steps;                    /* future_root_wdl shift (src/game.py:117-133); 0 = zeros */
    int32_t temp_decay_moves;            /* src/game.py:54-63 */
    float temp_init, temp_endgame;
    uint64_t seed;
    uint64_t uid_stride;                 /* a slot's next game gets uid += uid_stride (= total games in flight over all ranks) */
    /* per-slot state (device) */
    az_root *states;                     /* [n] env states = search roots */
    int32_t *steps;                      /* [n] plies played in the current game */
    uint64_t *uids;                      /* [n] global id of the game in this slot */
    /* staging of the running trajectories (device) */
    int8_t *st_state;                    /* [n][T][3*S] */
    float *st_prob;                      /* [n][T][A] */
    float *st_wdl;                       /* [n][T][3] */
    uint8_t *st_mask;                    /* [n][T][A] */
    int8_t *st_player;                   /* [n][T] */
    /* per-ply scratch (device) */
    int32_t *actions;                    /* [n] action played (written by ply; -1 = tree must reset) */
    uint8_t *finished;                   /* [n] game ended at this ply */
    /* output (device) */
    uint8_t *out;                        /* [out_capacity][record_bytes] packed finished games */
    int32_t *out_count;                  /* number of records written (may exceed capacity: extra games are dropped) */
    int32_t out_capacity, record_bytes;
} az_selfplay;

/* packed record layout (all offsets in bytes from the start of a record; T1 = max_plies + 1 positions) */
typedef struct az_selfplay_layout {
    int32_t record_bytes, T1;
    int32_t off_header;   /* int32 length, int32 winner, uint64 uid */
    int32_t off_state;    /* int8  [T1][3*S] */
    int32_t off_prob;     /* f32   [T1][A]   */
    int32_t off_root_wdl; /* f32   [T1][3]   */
    int32_t off_future;   /* f32   [T1][3]   */
    int32_t off_winner;   /* int8  [T1]      */
    int32_t off_steps;    /* int16 [T1]      */
    int32_t off_aux;      /* int16 [T1]      */
    int32_t off_mask;     /* u8    [T1][A]   */
} az_selfplay_layout;

int az_selfplay_layout_for(int game, az_selfplay_layout *out);
/* one ply for every slot: counts int32[n,A] and root_stats f32[n,6+8A] come from az_mcts_get_counts_dev /
 * az_mcts_get_root_stats_dev; writes sp->actions (hand them to az_mcts_prune_roots_dev) and sp->finished. */
int az_selfplay_ply_dev(const az_selfplay *sp, const int32_t *d_counts, const float *d_root_stats, void *stream);
/* for every finished slot: build the training tuples into sp->out, then restart the slot (env reset, steps = 0, uid += stride) */
int az_selfplay_flush_dev(const az_selfplay *sp, void *stream);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* AZB200_SELFPLAY_H */
