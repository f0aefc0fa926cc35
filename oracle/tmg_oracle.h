/*
 * TEST INFRASTRUCTURE ONLY.  CPU oracle for the tile-match board-transition path.
 *
 * Plain-C, literal (list-semantics) restatement of the reference
 *   /root/reference/src/tile_match_gym/board.py, tile_match_env.py, wrappers.py
 * used only as the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  The product (tile_match_gym_b200 / libtmg_b200.so) never links,
 * imports or calls anything in this directory.
 *
 * Parity pinning: tests/test_oracle_golden.py replays (a) every top-level engine call made
 * by the reference's own 16 tests (recorded by oracle/gen_golden.py), (b) the RNG-pinned
 * trajectories of tests/test_env.py, tests/board/test_move.py, tests/test_wrappers.py via
 * record-and-replay of PCG64 draws, (c) differential traces of the unmodified reference
 * driven by the shared Philox stream.
 */
#ifndef TMG_ORACLE_H
#define TMG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* special-enable bits (same values as include/tmg_b200.h) */
#define TMGO_SP_COOKIE 1u
#define TMGO_SP_VLASER 2u
#define TMGO_SP_HLASER 4u
#define TMGO_SP_BOMB 8u

/* status bits (same values as include/tmg_b200.h) */
#define TMGO_ST_BAD_ACTION 1u
#define TMGO_ST_NEEDS_RESET 2u
#define TMGO_ST_DRAWS_EXHAUSTED 4u
#define TMGO_ST_RESET_CAP 8u
#define TMGO_ST_LINE_OVERFLOW 16u
#define TMGO_ST_DFS_OVERFLOW 32u
#define TMGO_ST_INVALID_BOARD 64u
#define TMGO_ST_INTERNAL 128u

#define TMGO_AUTORESET_DISABLED 0
#define TMGO_AUTORESET_NEXT_STEP 1
#define TMGO_AUTORESET_SAME_STEP 2

typedef struct tmgo_board tmgo_board;

/* ---- single board (mirrors reference class Board, board.py:41-726) ---- */
tmgo_board *tmgo_board_create(int num_rows, int num_cols, int num_colours, uint32_t specials);
/* NOT the reference: generate_board is replaced by the constructive line-free sampler (include/tmg_b200.h) */
void tmgo_board_set_constructive(tmgo_board *b, int on);
void tmgo_board_destroy(tmgo_board *b);
/* stream: Philox (seed, env_id) or injected pre-drawn colours (draws may be NULL) */
void tmgo_board_set_stream(tmgo_board *b, uint64_t seed, uint32_t env_id, uint64_t draw_cursor,
                           uint64_t shuffle_cursor);
void tmgo_board_set_injected(tmgo_board *b, const uint8_t *draws, int64_t len, int64_t cursor);
void tmgo_board_get_cursors(const tmgo_board *b, uint64_t *draw_cursor, uint64_t *shuffle_cursor);
/* number of the last generated board (-1 before the first generate_board); selects the reset streams */
void tmgo_board_set_episode(tmgo_board *b, int32_t episode);
int32_t tmgo_board_get_episode(const tmgo_board *b);
uint32_t tmgo_board_status(const tmgo_board *b);
/* board I/O as int32 [2][R][C], the reference's dtype (board.py:96) */
void tmgo_board_set(tmgo_board *b, const int32_t *planes);
void tmgo_board_get(const tmgo_board *b, int32_t *planes);
void tmgo_board_set_counters(tmgo_board *b, int num_new_specials, int num_specials_activated);
void tmgo_board_get_counters(const tmgo_board *b, int *num_new_specials, int *num_specials_activated);
/* limits for generate_board / playability loops (0 = unlimited) */
void tmgo_board_set_iter_cap(tmgo_board *b, int64_t cap);
/* diagnostics gathered while running (max lines per round, max DFS depth) */
void tmgo_board_diag(const tmgo_board *b, int *max_lines, int *max_dfs_depth, int64_t *reset_iters);

/* primitives (each cites the reference function it restates in tmg_oracle.c) */
int tmgo_num_actions(const tmgo_board *b);
void tmgo_action_to_coords(const tmgo_board *b, int action, int *r1, int *c1, int *r2, int *c2);
void tmgo_generate_board(tmgo_board *b);
void tmgo_shuffle(tmgo_board *b);
void tmgo_gravity(tmgo_board *b);
void tmgo_refill(tmgo_board *b);
int tmgo_is_move_legal(const tmgo_board *b, int r1, int c1, int r2, int c2);
int tmgo_is_move_effective(tmgo_board *b, int r1, int c1, int r2, int c2);
int tmgo_possible_move(tmgo_board *b);
/* lines out: cells as r*C+c, line i occupies cells[offsets[i]..offsets[i+1]) ; returns n lines */
int tmgo_get_colour_lines(tmgo_board *b, int32_t *cells, int32_t *offsets, int max_cells, int max_lines);
/* matches out: same layout + names (0 normal,1 vertical_laser,2 horizontal_laser,3 bomb,4 cookie), colours */
int tmgo_detect_colour_matches(tmgo_board *b, int32_t *cells, int32_t *offsets, int32_t *names, int32_t *colours,
                               int max_cells, int max_matches);
/* one cascade round without gravity/refill: detect + resolve (board.py:369-373). returns n matches */
int tmgo_resolve_round(tmgo_board *b);
/* get_special_creation_pos (board.py:429-458) on explicit cell lists; returns r*C+c */
int tmgo_special_creation_pos(tmgo_board *b, const int32_t *cells, int n, const int32_t *taken_cells, int ntaken, int straight);
void tmgo_activate_special(tmgo_board *b, int r, int c, int tile_type, int is_combination_match);
void tmgo_combination_match(tmgo_board *b, int r1, int c1, int r2, int c2);
/* out[5] = num_eliminations, is_combination_match, num_new_specials, num_specials_activated, shuffled.
 * returns 0, or -1 for an illegal move (reference raises ValueError, board.py:349-350) */
int tmgo_move(tmgo_board *b, int r1, int c1, int r2, int c2, int32_t out[5]);
/* legal-move mask (tile_match_env.py:118-124 without the terminal rule): out[A] of 0/1 */
void tmgo_effective_mask(tmgo_board *b, uint8_t *out);
/* OneHotWrapper._one_hot_encode_board (wrappers.py:54-69): out[(K+S)][R][C] of 0/1 */
int tmgo_onehot_planes(const tmgo_board *b);
void tmgo_onehot(const tmgo_board *b, uint8_t *out);

/* ---- vectorised env with the same semantics as the product's C ABI (include/tmg_b200.h) ---- */
typedef struct tmgo_vec tmgo_vec;
typedef struct tmgo_vec_config {
    int32_t num_envs, num_rows, num_cols, num_colours, num_moves;
    uint32_t specials;
    int32_t autoreset;   /* TMGO_AUTORESET_* */
    int32_t refill_mode; /* 0 philox, 1 injected */
    uint64_t seed;
    uint64_t env_id_offset;
    int64_t max_reset_iters; /* 0 = unlimited */
    int32_t num_threads;     /* host threads used by tmgo_vec_step/reset */
    uint32_t flags;          /* TMGO_FLAG_* */
} tmgo_vec_config;
#define TMGO_FLAG_CONSTRUCTIVE_RESET 8u /* the product's TMG_FLAG_CONSTRUCTIVE_RESET contract (include/tmg_b200.h) */

typedef struct tmgo_vec_buffers { /* host arrays owned by the oracle, SoA over envs */
    int8_t *board;            /* [N][2][R][C] */
    int32_t *timer;           /* [N] */
    uint64_t *draw_cursor;    /* [N] */
    uint64_t *shuffle_cursor; /* [N] */
    int32_t *reward;          /* [N] */
    uint8_t *terminated;      /* [N] */
    uint8_t *is_combination_match;
    int32_t *num_new_specials;
    int32_t *num_specials_activated;
    uint8_t *shuffled;
    uint8_t *mask;            /* [N][A] */
    int32_t *num_moves_left;  /* [N] */
    uint32_t *status;         /* [N], sticky */
    int32_t *episode;         /* [N] number of the current board (-1 before the first reset) */
} tmgo_vec_buffers;

tmgo_vec *tmgo_vec_create(const tmgo_vec_config *cfg);
void tmgo_vec_destroy(tmgo_vec *v);
void tmgo_vec_get_buffers(tmgo_vec *v, tmgo_vec_buffers *out);
void tmgo_vec_set_injected_draws(tmgo_vec *v, const uint8_t *draws, int64_t per_env_len);
/* reset_mask NULL = all; init_boards NULL = generate_board, else int8 [N][2][R][C] */
void tmgo_vec_reset(tmgo_vec *v, const uint8_t *reset_mask, const int8_t *init_boards);
void tmgo_vec_step(tmgo_vec *v, const int32_t *actions);
/* every env stepped `steps` times with uniform actions from stream 2 (action k of env e = mulhi(word(action_seed,e,2,step0+k), A)); returns the reward sum */
int64_t tmgo_vec_rollout(tmgo_vec *v, int steps, uint64_t action_seed, uint64_t step0);
void tmgo_vec_onehot(tmgo_vec *v, uint8_t *out);
/* diagnostics over all envs */
void tmgo_vec_diag(tmgo_vec *v, int *max_lines, int *max_dfs_depth, int64_t *max_reset_iters);

/* stream helpers exposed for tests */
void tmgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint32_t tmgo_stream_word(uint64_t seed, uint32_t env_id, uint32_t stream, uint64_t k);

#ifdef __cplusplus
}
#endif
#endif
