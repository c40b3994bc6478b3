/*
 * azb200_gomoku.h - C ABI of the Gomoku environment (drop-in for the reference's `env_cpp.gomoku.Env`,
 * src/cpp/Gomoku.h:11-296 + src/cpp/env_gomoku.h:60-171).  The reference registers no MCTS engine for Gomoku
 * (src/cpp/mcts_bindings.cpp:393-394), so this is Env-only, like the reference.
 *
 * The reference keeps a byte board (std::vector<int8_t>); here a position is two sets of ROW BIT MASKS
 * (bit c of rows[p][r] = a stone of player p at (r, c)), so "k-th empty cell" is a popcount walk, a win test is
 * <= 8 short ray walks over bit tests, and the whole record is 288 bytes (nine 32-byte sectors).  The byte board the
 * Python API exposes is materialised only at the boundary (az_gomoku_export / az_gomoku_observe_dev).
 *
 *  - az_gomoku_*      : ONE game in HOST memory - the per-object Env API src/game.py drives once per ply (API glue).
 *  - az_gomoku_*_dev  : N games advanced in lockstep ON THE DEVICE (one thread per game; the rollout kernel keeps the
 *                       row masks in shared memory for the whole game).
 *
 * Board sizes 1..32 (the reference accepts any positive size; a 32-bit row mask bounds it here - AZ_ERR_GOMOKU_SIZE).
 */
#ifndef AZB200_GOMOKU_H
#define AZB200_GOMOKU_H
#include "azb200.h"

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

#define AZ_GOMOKU_MAX_SIZE 32
typedef struct az_gomoku {                       /* Gomoku private state, Gomoku.h:206-217 */
    uint32_t rows[2][AZ_GOMOKU_MAX_SIZE];        /* rows[0] = stones of player +1, rows[1] = player -1 */
    int32_t size, n_in_row;                      /* board_size_, n_in_row_ */
    int32_t turn, n_pieces, last_action, last_player, winner, done;
} az_gomoku;

/* step / set_params / import status codes: the std::runtime_error cases of the reference, in its order of checks */
#define AZ_GOMOKU_FINISHED 1          /* "game is already finished"          Gomoku.h:65-66 */
#define AZ_GOMOKU_OUT_OF_RANGE 2      /* "action out of range"               Gomoku.h:67-68 */
#define AZ_GOMOKU_OCCUPIED 3          /* "cell is already occupied"          Gomoku.h:69-70 */
#define AZ_GOMOKU_BAD_SIZE 4          /* "board_size must be positive"       Gomoku.h:216-217 */
#define AZ_GOMOKU_BAD_N_LOW 5         /* "n_in_row must be >= 2"             Gomoku.h:218-219 */
#define AZ_GOMOKU_BAD_N_HIGH 6        /* "n_in_row must be <= board size"    Gomoku.h:220-221 */
#define AZ_GOMOKU_BAD_CELL 7          /* "board values must be -1, 0, or 1"  Gomoku.h:187-190 */
#define AZ_GOMOKU_BAD_SYM 8           /* "invalid symmetry id"               Gomoku.h:119-120,132-133 */
#define AZ_ERR_GOMOKU_SIZE 9          /* board_size > AZ_GOMOKU_MAX_SIZE (limit of this implementation) */

/* ---- single game, host memory ---- */
int az_gomoku_set_params(az_gomoku *s, int board_size, int n_in_row);   /* set_params: Gomoku.h:21-28 (validates, resets) */
void az_gomoku_reset(az_gomoku *s);                                     /* reset: Gomoku.h:30-39 */
int az_gomoku_import(az_gomoku *s, const int8_t *board);                /* import_board + sync_from_board: :57-61,160-204 */
void az_gomoku_export(const az_gomoku *s, int8_t *board);               /* board_data: :48 */
int az_gomoku_step(az_gomoku *s, int action);                           /* step: :63-92 (validated; 0 or a status code) */
int az_gomoku_valid_moves(const az_gomoku *s, int32_t *moves);          /* get_valid_moves: :99-107 (ascending), returns count */
int az_gomoku_apply_symmetry(az_gomoku *s, int sym_id);                 /* apply_symmetry: :130-158 */
int az_gomoku_inverse_symmetry_action(int board_size, int sym_id, int action);   /* :115-128; < 0 = -status */

/* ---- N games in lockstep, device memory (d_states = az_gomoku[n]) ---- */
int az_gomoku_reset_dev(int n, int board_size, int n_in_row, az_gomoku *d_states, void *stream);
/* step game i with d_actions[i].  Games that are already done, or whose action is < 0, are left untouched with status 0;
 * an illegal action (out of range / occupied cell - the cases where the reference throws) leaves the game untouched and
 * is reported in d_status[i] (optional).  d_winners / d_dones (optional) hold winner and done flag AFTER the step. */
int az_gomoku_step_dev(int n, az_gomoku *d_states, const int32_t *d_actions, uint8_t *d_status, int32_t *d_winners,
                       uint8_t *d_dones, void *stream);
/* byte boards int8[n,S*S], legal masks u8[n,S*S] (= empty cells, also after the game is over, like valid_mask of
 * env_gomoku.h:118-125), turns i32[n], winners i32[n], dones u8[n]; any pointer may be NULL; d_boards / d_masks must be
 * 16-byte aligned (128-bit stores).  board_size = the size all n games were reset with. */
int az_gomoku_observe_dev(int n, int board_size, const az_gomoku *d_states, int8_t *d_boards, uint8_t *d_masks, int32_t *d_turns,
                          int32_t *d_winners, uint8_t *d_dones, void *stream);
/* D4 symmetry sym_ids[i] applied to game i (apply_symmetry, Gomoku.h:130-158); an invalid id leaves the game untouched */
int az_gomoku_symmetry_dev(int n, az_gomoku *d_states, const int32_t *d_sym_ids, void *stream);
/* Lockstep random rollouts (the Gomoku twin of az_envs_rollout_dev): game g = first_game + i plays the
 * ((hash(seed, g, ply) >> 32) * #empty >> 32)-th empty cell in ascending action order from the empty board until it is over.
 * d_digest[i] = checksum of the final state, d_plies[i] = game length.  When n_record > 0 the first n_record games also
 * record, per ply, the byte board and turn BEFORE the move, the action, and winner / done AFTER it, into arrays shaped
 * [n_record, S*S, ...] (S*S = the longest possible game).  d_final (optional) receives the final az_gomoku records. */
int az_gomoku_rollout_dev(int n, int board_size, int n_in_row, uint64_t seed, uint64_t first_game, uint64_t *d_digest,
                          int32_t *d_plies, int n_record, int8_t *d_rec_boards, int32_t *d_rec_turns, int32_t *d_rec_actions,
                          int32_t *d_rec_winners, uint8_t *d_rec_dones, az_gomoku *d_final, void *stream);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* AZB200_GOMOKU_H */
