/*
 * azb200_env.h - C ABI of the bitboard game environments (drop-in for the reference's `env_cpp` pybind module,
 * src/cpp/env_bindings.cpp + env_common.h + env_connect4.h + env_othello.h; game logic src/cpp/Connect4.h,
 * src/cpp/Othello.h).
 *
 *  - az_env_*      : ONE game held in a 32-byte az_root record on the HOST.  This is the host-side mirror of the
 *                    per-object `Env` API (`env.step`, `env.valid_move()`, ...) that src/game.py drives once per ply;
 *                    it is API glue, not the hot path.
 *  - az_envs_*_dev : N games advanced in lockstep ON THE DEVICE (one thread per game, az_root[n] in HBM - the same
 *                    records az_mcts_search_dev takes as roots, so self-play never leaves the GPU).
 */
#ifndef AZB200_ENV_H
#define AZB200_ENV_H
#include "azb200.h"

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

/* ---- single game, host memory (Env object API: env_common.h:133-249, env_connect4.h:30-65, env_othello.h:30-73) ---- */
void az_env_reset(int game, az_root *s);                              /* Env() / reset(): Connect4.h:62-72, Othello.h:62-75 */
void az_env_import(int game, az_root *s, const int8_t *board);        /* import_board; turn is left untouched */
void az_env_export(int game, const az_root *s, int8_t *board);        /* sync_to_board + board_data */
int az_env_n_pieces(int game, const az_root *s);
void az_env_step(int game, az_root *s, int action);                   /* step (no legality check, like the reference) */
int az_env_winner(int game, const az_root *s);                        /* winPlayer / check_winner */
int az_env_full(int game, const az_root *s);                          /* check_full */
int az_env_done(int game, const az_root *s);                          /* done */
int az_env_valid_moves(int game, const az_root *s, int32_t *moves);   /* valid_move: ascending, returns the count */
void az_env_apply_symmetry(int game, az_root *s, int sym_id);         /* apply_symmetry */
int az_env_inverse_symmetry_action(int game, int sym_id, int action); /* inverse_symmetry_action (static) */

/* ---- N games in lockstep, device memory ---- */
int az_envs_reset_dev(int game, int n, az_root *d_states, void *stream);
/* step game i with d_actions[i]; games that are already done (or whose action is < 0) are left untouched.
 * d_winners / d_dones (optional) receive winner and done flag AFTER the step. */
int az_envs_step_dev(int game, int n, az_root *d_states, const int32_t *d_actions, int32_t *d_winners, uint8_t *d_dones,
                     void *stream);
/* byte boards int8[n,S], legal masks u8[n,A], turns i32[n], winners i32[n], dones u8[n]; any pointer may be NULL */
int az_envs_observe_dev(int game, int n, const az_root *d_states, int8_t *d_boards, uint8_t *d_masks, int32_t *d_turns,
                        int32_t *d_winners, uint8_t *d_dones, void *stream);
/* Config-2 workload: game g = first_game + i plays action legal[hash(seed, g, ply) mod len(legal)] from
 * the initial position until it is over, entirely on the device.  d_digest[i] = checksum of the final state,
 * d_plies[i] = game length.  When n_record > 0 the first n_record games also record, per ply, the byte board / legal
 * mask / turn BEFORE the move, the action, and winner / done AFTER it into arrays shaped [n_record, max_plies, ...]. */
int az_envs_rollout_dev(int game, int n, uint64_t seed, uint64_t first_game, uint64_t *d_digest, int32_t *d_plies,
                        int n_record, int max_plies, int8_t *d_rec_boards, uint8_t *d_rec_masks, int32_t *d_rec_turns,
                        int32_t *d_rec_actions, int32_t *d_rec_winners, uint8_t *d_rec_dones, void *stream);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* AZB200_ENV_H */
