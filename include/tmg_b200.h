/*
 * tmg_b200 -- C ABI of the B200-native batched tile-match board-transition engine.
 *
 * This is the drop-in boundary for the hot path of akshilpatel/tile-match-gym: what the reference's
 * `TileMatchEnv` calls on its `Board` (tile_match_env.py:50-51,59,87,98,122):
 *
 *   reference (Python)                                   this ABI (one call per batch of envs)
 *   ---------------------------------------------------  -------------------------------------
 *   Board(num_rows, ..., np_random)   board.py:42-93      tmg_create
 *   Board.generate_board()            board.py:95-112     tmg_reset (init_boards == NULL)
 *   env.board.board = b; timer = 0    board.py:65-74      tmg_reset (init_boards != NULL)
 *   Board.move(c1, c2) + timer/done   board.py:330-395,   tmg_step
 *     + _get_effective_actions          tile_match_env.py:93-124
 *   is_move_effective x num_actions   board.py:735-787    tmg_legal_mask
 *   OneHotWrapper.observation         wrappers.py:48-69   tmg_encode_onehot
 *   np_random.integers / .shuffle     board.py:97,116,    counter-based Philox stream (see below) or
 *                                       129,239             tmg_set_injected_draws
 *
 * All `dev` pointers are CUDA device pointers on the device the env was created on; every call is
 * asynchronous on the caller's stream (cudaStream_t passed as void*; NULL = default stream) unless it
 * says "host".  Plain C types only.  A handle is not thread-safe; distinct handles are independent.
 * There is no CPU fallback: every entry point needs a CUDA device of compute capability 10.0.
 *
 * Draw stream ("identical refill draws" contract).  Env e (global id = env_id_offset + local index)
 * draws its k-th refill colour as  1 + mulhi32(W(seed, e, 0, k), K)  and shuffles with Fisher-Yates
 * (i = P-1..1, j = mulhi32(W(seed, e, 1, next), i+1)), where W(seed, e, s, k) is word k&3 of
 * Philox4x32-10(key = (seed lo, seed hi), ctr = (k>>2 lo, k>>2 hi, e, s)).  Draws are consumed exactly
 * where the reference calls np_random.integers(1, K+1, n): row-major over the empty cells of each
 * refill (board.py:239-240), the R*C initial fill (board.py:97) and the (row+1)*C row-block redraws of
 * remove_colour_lines (board.py:129).
 * While a board is being generated (generate_board, board.py:95-112) the draws come from episode-indexed
 * streams that start at word 0 for every board: W_reset(seed, e, j, k) = Philox(key, ctr = (k>>2, j, e, 3))[k&3]
 * for integers() and ctr = (k>>2, j, e, 4) for shuffle(), j = number of the board (0 for the first reset).  The
 * j-th board of an env is therefore a pure function of (seed, e, j): the library generates it ahead of time on a
 * side stream ("pool") and tmg_step only copies it in; generating it inside tmg_step (TMG_FLAG_NO_PREGEN) gives the
 * same bytes.  With refill_mode = TMG_REFILL_INJECTED every draw, resets included, is draws[e][k] in consumption
 * order instead (pre-drawn by the caller, e.g. from numpy's PCG64, whose integers() stream is contiguous across
 * call sizes), and boards are generated inside the call.
 */
#ifndef TMG_B200_H
#define TMG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMG_ABI_VERSION 2   /* 2: tmg_host_io.board_packed */

/* return codes */
#define TMG_OK 0
#define TMG_ERR_INVALID_ARG 1
#define TMG_ERR_UNSUPPORTED_SHAPE 2
#define TMG_ERR_CUDA 3
#define TMG_ERR_NO_DEVICE 4
#define TMG_ERR_OOM 5
#define TMG_ERR_STATE 6

/* enabled specials (reference TILE_TYPES, board.py:18-25) */
#define TMG_SP_COOKIE 1u           /* type -1, colourless */
#define TMG_SP_VERTICAL_LASER 2u   /* type 2 */
#define TMG_SP_HORIZONTAL_LASER 4u /* type 3 */
#define TMG_SP_BOMB 8u             /* type 4 */

/* per-env sticky status bits (the vector replacement for the reference's exceptions) */
#define TMG_ST_BAD_ACTION 1u       /* action outside [0, A): IndexError at tile_match_env.py:97 */
#define TMG_ST_NEEDS_RESET 2u      /* step before reset / after done: Exception at tile_match_env.py:94-95 */
#define TMG_ST_DRAWS_EXHAUSTED 4u  /* injected draw stream ran out (colour 1 substituted) */
#define TMG_ST_RESET_CAP 8u        /* generate_board / playability loop hit max_reset_iters */
#define TMG_ST_LINE_OVERFLOW 16u   /* more colour lines in one cascade round than the on-chip table holds */
#define TMG_ST_DFS_OVERFLOW 32u    /* special-activation chain deeper than the on-chip stack */
#define TMG_ST_INVALID_BOARD 64u   /* injected board is not a full board of (1..K,1..4) tiles / (0,-1) cookies */
#define TMG_ST_INTERNAL 128u       /* a state the reference would raise on (e.g. no valid creation cell) */

#define TMG_AUTORESET_DISABLED 0   /* like the reference: stepping a finished env is an error */
#define TMG_AUTORESET_NEXT_STEP 1  /* gymnasium NEXT_STEP: the step after `terminated` resets, action ignored */
#define TMG_AUTORESET_SAME_STEP 2  /* gymnasium SAME_STEP: the terminal step also generates the next board */

#define TMG_REFILL_PHILOX 0
#define TMG_REFILL_INJECTED 1

#define TMG_FLAG_NO_MASK 1u        /* do not maintain the legal-move mask in tmg_step / tmg_reset */
#define TMG_FLAG_NO_PREGEN 2u      /* generate every board inside tmg_step instead of ahead of time on a side stream */
#define TMG_FLAG_BYTE_PLANES 4u    /* run the moves on the byte planes in shared memory instead of the register-resident
                                      bit-plane engine (same results; diagnostics / A-B measurements) */

#define TMG_FLAG_CONSTRUCTIVE_RESET 8u /* NOT the reference's generate_board: every board (tmg_reset, autoreset, pool) comes
                                      from a constructive line-free sampler, for shapes where the reference's redraw loop
                                      (board.py:99-109) does not terminate (e.g. 32x32 / 7 colours).  Board j of env e:
                                      cells in row-major order, each takes the next colour 1 + mulhi32(W5(k), K), k = 0, 1,
                                      ..., W5(k) = Philox4x32-10(key, ctr = (k>>2, j, e, 5))[k&3], that does not complete a
                                      triple with the two cells to its left or the two above it (at most 64 draws per cell,
                                      then TMG_ST_RESET_CAP); all types 1; a board without a possible move is drawn again
                                      from where the stream stands.  The env's refill stream is not touched. */

#define TMG_MAX_ROWS 32
#define TMG_MAX_COLS 32
#define TMG_MAX_COLOURS 31

typedef struct tmg_env tmg_env;

typedef struct tmg_config {
    uint32_t struct_size;      /* sizeof(tmg_config), for ABI evolution */
    int32_t device;            /* CUDA device ordinal */
    int32_t num_envs;          /* envs held by THIS handle (one shard) */
    int32_t num_rows;          /* TileMatchEnv(num_rows, num_cols, num_colours, num_moves, ...) */
    int32_t num_cols;
    int32_t num_colours;
    int32_t num_moves;
    uint32_t specials;         /* TMG_SP_* of colourless_specials + colour_specials */
    int32_t autoreset;         /* TMG_AUTORESET_* */
    int32_t refill_mode;       /* TMG_REFILL_* */
    uint32_t flags;            /* TMG_FLAG_* */
    int32_t max_reset_iters;   /* cap on redraw+shuffle iterations of one generate/playability loop; 0 = default 16384 */
    uint64_t seed;             /* Philox key */
    uint64_t env_id_offset;    /* global id of local env 0 (sharding by env index across GPUs) */
} tmg_config;

/* Device buffers owned by the handle, SoA over envs.  Valid until tmg_destroy; contents are valid
 * after the stream work of the last call completed and until the next call (the reference's obs
 * aliases live state the same way, tile_match_env.py:115). */
typedef struct tmg_buffers {
    int8_t *board;                   /* [N][2][R][C]  plane 0 colour 0..K, plane 1 type -1..4 (board.py:96) */
    int32_t *timer;                  /* [N]  moves made; -1 before the first reset */
    uint64_t *draw_cursor;           /* [N]  next index into the refill stream */
    uint64_t *shuffle_cursor;        /* [N]  next index into the shuffle stream */
    int32_t *reward;                 /* [N]  num_eliminations of the last step */
    uint8_t *terminated;             /* [N] */
    uint8_t *is_combination_match;   /* [N] */
    int32_t *num_new_specials;       /* [N] */
    int32_t *num_specials_activated; /* [N] */
    uint8_t *shuffled;               /* [N] */
    uint8_t *mask;                   /* [N][A]  1 = is_move_effective; all 0 on a terminal step */
    int32_t *num_moves_left;         /* [N] */
    uint32_t *status;                /* [N]  TMG_ST_* (sticky; clear with tmg_clear_status) */
    int32_t *episode;                /* [N]  number of the current board (-1 before the first reset) */
} tmg_buffers;

int tmg_abi_version(void);
/* hash of the sources the library was built from (the Python loader rebuilds the library when it differs from the tree) */
const char *tmg_build_id(void);
const char *tmg_error_string(int code);
/* "a|b|c" names of the TMG_ST_* bits set in `status` (static buffer per thread) */
const char *tmg_status_string(uint32_t status);

int tmg_num_actions(int32_t num_rows, int32_t num_cols);              /* board.py:77 */
int tmg_onehot_planes(int32_t num_colours, uint32_t specials);        /* wrappers.py:25 */
/* action id -> ((r1,c1),(r2,c2)), board.py:80-91 (host) */
int tmg_action_to_coords(int32_t num_rows, int32_t num_cols, int32_t action, int32_t out_r1c1r2c2[4]);

int tmg_create(const tmg_config *cfg, tmg_env **out);
int tmg_destroy(tmg_env *env);
int tmg_get_buffers(tmg_env *env, tmg_buffers *out);

/* draws: device uint8 [N][per_env_len], values in 1..K; kept by reference (caller owns the memory) */
int tmg_set_injected_draws(tmg_env *env, const uint8_t *draws_dev, int64_t per_env_len);

/* reset_mask_dev: device uint8 [N] (NULL = all envs).  init_boards_dev: device int8 [N][2][R][C]
 * (NULL = generate_board from the draw stream).  Sets timer=0, zeroes the step outputs, computes the mask. */
int tmg_reset(tmg_env *env, const uint8_t *reset_mask_dev, const int8_t *init_boards_dev, void *stream);

/* actions_dev: device int32 [N].  One TileMatchEnv.step per env, cascade loop inside the kernel. */
int tmg_step(tmg_env *env, const int32_t *actions_dev, void *stream);

/* Fused rollout: num_steps successive TileMatchEnv.step calls per env in ONE launch, env e taking
 * actions_dev[t][e] at step t (device int32 [num_steps][N]) -- the caller loop of src/examples/random_agent.py:12-31
 * with the actions drawn up front.  State, buffers and status afterwards are exactly those of num_steps tmg_step
 * calls (the per-step outputs in tmg_buffers are those of the last step); rewards_dev (int32) / terminated_dev
 * (uint8), each [num_steps][N] or NULL, receive every step's reward and termination flag.  Boards stay on chip
 * between the steps and no env waits for the longest cascade of the batch at every step. */
int tmg_step_many(tmg_env *env, const int32_t *actions_dev, int32_t num_steps, int32_t *rewards_dev,
                  uint8_t *terminated_dev, void *stream);

/* Fused rollout with the agent inside the kernel (the random agent of src/examples/random_agent.py:12-31, or one that
 * samples from info["effective_actions"], tile_match_env.py:118-124).  At a step of env e (global id g) on board number
 * j with `timer` moves made, the action is decided by word k = j * num_moves + timer of the env's action stream,
 * W(seed, g, 2, k) (see the draw-stream contract above):
 *   TMG_POLICY_UNIFORM  action = mulhi32(W, A)
 *   TMG_POLICY_MASK     the mulhi32(W, n)-th of the n effective actions in index order (uniform if n == 0)
 * Steps that take no action (the reset step of TMG_AUTORESET_NEXT_STEP, a finished env) record action 0.
 * actions_out_dev (int32 [num_steps][N] or NULL) receives the actions taken; everything else as tmg_step_many. */
#define TMG_POLICY_UNIFORM 1
#define TMG_POLICY_MASK 2
int tmg_rollout_policy(tmg_env *env, int32_t policy, int32_t num_steps, int32_t *actions_out_dev, int32_t *rewards_dev,
                       uint8_t *terminated_dev, void *stream);

/* recompute buffers.mask from the current boards (e.g. after the caller edited boards in place) */
int tmg_legal_mask(tmg_env *env, void *stream);

/* out_dev: device uint8 [N][K+S][R][C] of 0/1, plane order colours 1..K then enabled specials in the
 * order cookie, vertical_laser, horizontal_laser, bomb (wrappers.py:40-46) */
int tmg_encode_onehot(tmg_env *env, uint8_t *out_dev, void *stream);
/* same, float32 planes (the reference returns float64 0./1. values) */
int tmg_encode_onehot_f32(tmg_env *env, float *out_dev, void *stream);
/* same, float64 planes: the dtype OneHotWrapper.observation itself returns (np.zeros default, wrappers.py:57,64) */
int tmg_encode_onehot_f64(tmg_env *env, double *out_dev, void *stream);

int tmg_clear_status(tmg_env *env, void *stream);

/* Makes `stream` wait for every board generation the library has queued on its side stream so far (a benchmark
 * calls it before reading the clock; ordinary callers never need it). */
int tmg_join(tmg_env *env, void *stream);

/* TileMatchEnv.reset(seed=...) / set_seed (tile_match_env.py:79-86): new Philox key, both cursors back to 0 */
int tmg_set_seed(tmg_env *env, uint64_t seed, void *stream);

/* Host-buffer convenience path (the call a CPU-side consumer makes): copies actions from host memory,
 * steps, and copies back whatever output pointers are non-NULL; synchronises the stream before
 * returning.  Pinned host memory makes the copies asynchronous DMA. */
typedef struct tmg_host_io {
    const int32_t *actions;          /* in  [N] */
    int8_t *board;                   /* out [N][2][R][C] or NULL */
    int32_t *reward;                 /* out [N] or NULL */
    uint8_t *terminated;             /* out [N] or NULL */
    uint8_t *mask;                   /* out [N][A] or NULL */
    uint8_t *mask_bits;              /* out [N][(A+7)/8] or NULL: the same mask, bit j of byte b = action 8b+j (2.3x fewer PCIe bytes) */
    int32_t *num_moves_left;         /* out [N] or NULL */
    uint8_t *is_combination_match;   /* out [N] or NULL */
    int32_t *num_new_specials;       /* out [N] or NULL */
    int32_t *num_specials_activated; /* out [N] or NULL */
    uint8_t *shuffled;               /* out [N] or NULL */
    uint32_t *status;                /* out [N] or NULL */
    uint8_t *board_packed;           /* out [N][R][C] or NULL: the board as one byte per cell, colour | (type & 7) << 4
                                        (cookie type -1 -> 7), half the PCIe bytes of `board`; needs num_colours <= 15 */
} tmg_host_io;
int tmg_step_host(tmg_env *env, const tmg_host_io *io, void *stream);

/* Host mirror: the fast form of the host-buffer path.  Registers the page-locked host arrays of `io` (cudaHostAlloc /
 * cudaHostRegister memory, 16-byte aligned; any of them may be NULL) as a mirror of buffers.board (as byte planes and / or
 * packed, one byte per cell) / buffers.mask / the bit-packed mask / reward / terminated / num_moves_left.  The
 * call copies the current contents in full; from then on tmg_step's kernels write, straight into the arrays over PCIe,
 * the board and mask entries of exactly those envs whose board or mask they changed (a step that changes nothing moves
 * no board bytes) and the three per-env scalars of every env (coalesced), and tmg_reset / tmg_legal_mask /
 * tmg_debug_op re-copy them in full.  The arrays are complete and current whenever the stream work of the last call
 * has finished -- the reference's obs["board"] aliases live state the same way (tile_match_env.py:115).
 * tmg_step_host skips the copies of pointers that are the bound arrays and, while a mirror is bound, reads page-locked
 * actions in place instead of staging them.  The other fields of `io` are ignored;
 * io == NULL unbinds.  Returns TMG_ERR_INVALID_ARG for pageable memory. */
int tmg_host_bind(tmg_env *env, const tmg_host_io *io, void *stream);

/* Debug / known-answer entry point: runs ONE engine primitive on every env's device board, so that the
 * reference's function-level tests (tests/board/*.py) can be replayed on the GPU.  args_dev: int32 [N][4].
 * Counters num_new_specials / num_specials_activated are taken from and written back to the buffers;
 * op-specific results go to buffers.reward (and, for TMG_OP_MOVE, the other step outputs). */
#define TMG_OP_GRAVITY 1        /* Board.gravity           board.py:217-229 */
#define TMG_OP_REFILL 2         /* Board.refill            board.py:231-241 */
#define TMG_OP_RESOLVE_ROUND 3  /* detect_colour_matches + resolve_colour_matches, board.py:369-373; reward = #lines */
#define TMG_OP_ACTIVATE 4       /* activate_special((a0,a1), type=a2, is_combination_match=a3)  board.py:473-556 */
#define TMG_OP_COMBINE 5        /* combination_match((a0,a1),(a2,a3))                         board.py:600-719 */
#define TMG_OP_MOVE 6           /* Board.move((a0,a1),(a2,a3)) without the env timer             board.py:330-395 */
#define TMG_OP_EFFECTIVE 7      /* is_move_effective(board,(a0,a1),(a2,a3)) -> reward            board.py:735-787 */
#define TMG_OP_GENERATE 8       /* Board.generate_board                                       board.py:95-112 */
#define TMG_OP_SHUFFLE 9        /* Board.shuffle                                              board.py:114-118 */
#define TMG_OP_COUNT_LINES 10   /* len(get_colour_lines()) -> reward                          board.py:149-215 */
#define TMG_OP_LINES 11         /* get_colour_lines() itself: see tmg_debug_lines                board.py:149-215 */
#define TMG_OP_BYTE_PLANES 0x100 /* OR into `op`: run the primitive on the byte-plane implementation (boards of up to 10 rows
                                    and 7 colours otherwise run it on the register-resident engine the step kernels use) */
int tmg_debug_op(tmg_env *env, int32_t op, const int32_t *args_dev, void *stream);

/* get_colour_lines() (board.py:149-215) of every env's board, as the engine's line table.  out_dev: device uint32
 * [N][TMG_LINES_WORDS]: word 0 = number of lines n (at most 32 are reported), then n entries {info, cells}:
 *   info  bits 0-11 position of the line in the reference's list (ascending = list order; values >= 1024 are the
 *         crossing segments of board.py:198-214), bits 12-15 the row of the line's first cell, bit 16 kind
 *         (0 horizontal, 1 vertical), bits 17-21 its row (horizontal) or column (vertical), bits 22-24 its colour;
 *   cells bit set of the line's columns (horizontal) or rows (vertical).
 * byte_planes != 0 reads the table of the byte-plane implementation instead of the register-resident engine's. */
#define TMG_LINES_WORDS 65
int tmg_debug_lines(tmg_env *env, uint32_t *out_dev, int32_t byte_planes, void *stream);

/* Diagnostics: when set, every tmg_step writes per env {SM cycles spent, cycles in the general (non-fast) round path,
 * cascade rounds, redraw iterations} and ADDS the general path's cycles split into {scan, line table, classification,
 * resolution} to prof_dev (device uint32 [N][8], zero it before the step); NULL switches it off. */
int tmg_set_profile_buffer(tmg_env *env, uint32_t *prof_dev);

#ifdef __cplusplus
}
#endif
#endif /* TMG_B200_H */
