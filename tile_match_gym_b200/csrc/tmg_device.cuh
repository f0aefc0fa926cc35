// tmg_device.cuh -- device side of the B200 tile-match engine (sm_100a).
//
// One sub-warp ("group") of L lanes owns one board; lane c owns column c.  The board's colour and type
// planes live in shared memory in exactly the HBM layout ([2][R][C] int8), loaded/stored with the widest
// vector the env stride allows.  Parallel parts (line scan by ballot, gravity by per-lane compaction,
// refill ranks by ballot/popc + counter-based Philox, legal-move mask by per-colour row bitboards and
// shuffles) run on all lanes; the order-dependent list semantics of the reference (line classification,
// bomb merging, creation cells, the activation DFS, combination matches) run on the group's leader lane
// over small shared-memory tables.  The whole cascade loop stays inside the kernel.
//
// Reference being restated (never copied): /root/reference/src/tile_match_gym/board.py -- cited per
// function as "ref :NNN".  Semantics are literal for every FULL board whose cells are arbitrary
// (colour, type) pairs, so states the reference can reach through its shuffle/redraw path (e.g. a cookie
// that was given a colour by remove_colour_lines, ref :129) behave identically.
//
// This header is also compiled by tests/emu (a CPU lane emulator used only by the test-suite to debug
// kernel logic without a GPU); TMG_EMU selects the shim.  The product never builds with TMG_EMU.
#pragma once
#include <stdint.h>

#ifdef TMG_EMU
#include "emu_shim.h"
// the emulator checks that all lanes of a group are at the same collective (source line of the caller)
#define TMG_SITE_P0 int site_ = __builtin_LINE()
#define TMG_SITE_P , int site_ = __builtin_LINE()
#define TMG_SITE_SET emu::site_id = site_;
#define TMG_SITE_HERE emu::site_id = __LINE__;
#else
#include <cuda_runtime.h>
#define TMG_SITE_P0
#define TMG_SITE_P
#define TMG_SITE_SET
#define TMG_SITE_HERE
#endif

// Code-size knobs.  The step kernels are bound by instruction supply (ncu: stall_no_instruction is the top stall).  In
// round 1 rolling these loops measured 10 % SLOWER (more executed instructions); with the register-resident engine of
// round 2 the balance flipped: Philox rolled by 2 and the byte-plane row loops rolled measured +7 % at 65 536 envs and +8 %
// at 1 M (see tmg_rb.cuh for the whole ladder).
#ifndef TMG_UNROLL_PHILOX
#define TMG_UNROLL_PHILOX 2
#endif
#ifndef TMG_UNROLL_ROWS
#define TMG_UNROLL_ROWS 1
#endif
#ifndef TMG_FUSED_FALL
#define TMG_FUSED_FALL 1
#endif


namespace tmg {
constexpr int UNROLL_PHILOX = TMG_UNROLL_PHILOX, UNROLL_ROWS = TMG_UNROLL_ROWS;

// ---- status / config constants (values mirror include/tmg_b200.h) ---------------------------------
enum : uint32_t {
    ST_BAD_ACTION = 1u, ST_NEEDS_RESET = 2u, ST_DRAWS_EXHAUSTED = 4u, ST_RESET_CAP = 8u,
    ST_LINE_OVERFLOW = 16u, ST_DFS_OVERFLOW = 32u, ST_INVALID_BOARD = 64u, ST_INTERNAL = 128u
};
enum : uint32_t { SP_COOKIE = 1u, SP_VLASER = 2u, SP_HLASER = 4u, SP_BOMB = 8u };
enum { AUTORESET_DISABLED = 0, AUTORESET_NEXT_STEP = 1, AUTORESET_SAME_STEP = 2 };
enum : uint32_t { FLAG_NO_MASK = 1u, FLAG_NO_PREGEN = 2u, FLAG_BYTE_PLANES = 4u, FLAG_CONSTRUCTIVE_RESET = 8u };
enum { OP_GRAVITY = 1, OP_REFILL, OP_RESOLVE_ROUND, OP_ACTIVATE, OP_COMBINE, OP_MOVE, OP_EFFECTIVE, OP_GENERATE,
       OP_SHUFFLE, OP_COUNT_LINES, OP_LINES, OP_LAST = OP_LINES, OP_BYTE_PLANES = 0x100 };
enum { LINES_WORDS = 1 + 2 * 32 };   // OP_LINES output per env: n, then {info, cell set} per line (see RBoard::line_info)
enum { POLICY_GIVEN = 0, POLICY_UNIFORM = 1, POLICY_MASK = 2 };
enum { NAME_NORMAL = 0, NAME_VLASER = 2, NAME_HLASER = 3, NAME_BOMB = 4, NAME_COOKIE = -1 };  // = created tile type

struct Params {
    int N, R, C, K, P, A, num_moves;
    uint32_t specials;
    int autoreset, use_inj;
    uint32_t flags;
    int max_iters;
    uint32_t key0, key1;
    uint64_t env_id_offset;
    int board_vecw, mask_vecw, init_vecw;  // widest power-of-two vector (bytes) dividing the per-env strides / pointers
    // state / outputs (device)
    int8_t* board;
    int32_t* timer;
    uint64_t* draw_cursor;
    uint64_t* shuffle_cursor;
    int32_t* reward;
    uint8_t* terminated;
    uint8_t* is_comb;
    int32_t* new_specials;
    int32_t* activated;
    uint8_t* shuffled;
    uint8_t* mask;
    int32_t* moves_left;
    uint32_t* status;
    const uint8_t* inj;
    long long inj_len;
    int32_t* episode;        // [N] number of the current board, -1 before the first generate_board
    int8_t* pool_board;      // [N][2][R][C] the NEXT board of each env, generated ahead of time (k_pregen)
    uint8_t* pool_mask;      // [N][A]       its legal-move mask
    int32_t* pool_episode;   // [N]          which episode the pool entry belongs to (-1 = none)
    uint32_t* pool_status;   // [N]          status bits raised while generating it (merged when it is consumed)
    int pool_tag;            // number of the k_pregen launch that serves the requests of this launch (k_gate / k_reset)
    // scheduling state (device): see Ctl below
    uint32_t* ctl;           // control words
    uint2* wl_items;         // [N] work list of the current step: {env, action | flags << 12}; high priority from the front, rest from the back
    uint8_t* n_special;      // [N] special tiles on the env's board (scheduling hint only: a stale value costs time, never correctness)
    uint2* req_ring;         // [req_mask + 1] pool-refill requests {env, board number}, in request order (NULL: pool not in use).
                             // A request names the board it wants, so a late, repeated or overwritten entry is harmless.
    uint32_t req_mask;       // ring capacity - 1 (capacity = power of two >= 8 N)
    int pregen_one_shot;     // k_pregen: a group generates ONE board and leaves (short-lived blocks: the block scheduler can then
                             // give the SM slot to a step kernel of a higher-priority stream), instead of looping over the requests
    int commit_pregen;       // this launch closes a batch of requests: the next k_pregen launch serves [CTL_REQ_PREV, tail)
    int seq;                 // number of this tmg_step call: its parity selects the work-list counters
    // host mirror (tmg_host_bind): page-locked host arrays, as device-visible pointers, that the step kernel updates in
    // place over PCIe for exactly the envs whose board / mask changed (NULL = not bound)
    int8_t* h_board;         // [N][2][R][C]
    uint8_t* h_board_packed; // [N][R][C]  colour | (type & 7) << 4: the same board in half the PCIe bytes (K <= 15)
    uint8_t* h_mask;         // [N][A]
    uint8_t* h_mask_bits;    // [N][(A+7)/8]  bit j of byte b = action 8b + j
    int32_t* h_reward;       // [N]  per-env scalars: k_gate writes them coalesced, the workers overwrite the reward of a move
    uint8_t* h_terminated;   // [N]
    int32_t* h_moves_left;   // [N]
    // per-call inputs
    const int32_t* actions;
    const uint8_t* reset_mask;
    const int8_t* init_boards;
    int T;                   // tmg_step_many: steps in this call; actions is [T][N]
    int32_t* ro_reward;      // [T][N] or NULL
    uint8_t* ro_terminated;  // [T][N] or NULL
    int32_t* ro_actions;     // [T][N] or NULL: the actions taken (on-device policies)
    int policy;              // 0: actions given; POLICY_UNIFORM / POLICY_MASK: drawn inside the kernel from stream 2
    const int32_t* dbg_args;
    int dbg_op;
    uint32_t* dbg_out;       // OP_LINES: [N][LINES_WORDS]
    uint32_t* prof;  // optional [N][8]: cycles total, cycles in the general path, cascade rounds, redraw iterations,
                     // then general-path cycles split into scan / line table / classification / resolution (zero before each step)
};

// Control words.  A step is two launches: k_gate (one thread per env) decides which envs need board work and appends
// them to the work list; k_work (persistent groups) pops items until the list is empty, so a group that finishes a
// short item immediately takes the next one and the launch does not wait on whole blocks of idle groups.  Requests
// for pool refills go to a ring in request order; the launch that issued them records its [start, end) range under
// its tag so that the k_pregen launch with that tag serves exactly those.
enum { PG_RING = 32 };       // k_pregen launches whose ranges are kept (host: event ring of the same size)
enum {
    CTL_WL_COUNT = 0,        // [2] items appended by k_gate, by step parity
    CTL_WL_HEAD = 2,         // [2] pop cursor of k_work, by step parity
    CTL_DONE = 4,            // warps / groups of the running k_gate / k_reset that have finished
    CTL_REQ_TAIL = 5,        // refill requests appended so far (monotonic, wraps with the ring)
    CTL_REQ_PREV = 6,        // CTL_REQ_TAIL at the end of the previous requesting launch
    CTL_PG_RANGE = 8,        // [PG_RING][2] request range served by k_pregen launch `tag`, at tag % PG_RING
    CTL_PG_HEAD = CTL_PG_RANGE + 2 * PG_RING,   // [PG_RING] pop cursor of that launch
    CTL_RO_HEAD = CTL_PG_HEAD + PG_RING,        // pop cursor of k_rollout over the envs
    CTL_WL_COUNT_LO = CTL_RO_HEAD + 1,          // [2] normal-priority items (CTL_WL_COUNT counts the high-priority ones)
    CTL_WORDS = CTL_WL_COUNT_LO + 2
};
#ifndef TMG_PRI_SPECIALS
#define TMG_PRI_SPECIALS 4
#endif
enum { PRI_SPECIALS = TMG_PRI_SPECIALS };   // a move on a board with this many special tiles is scheduled first: 23 % of the effective
                             // moves, 80 % of the longest 1 % of the cascades (10x10, 4 colours, measured on the CPU restatement)
enum : uint32_t { IT_ACTION = 0xfffu, IT_EFF = 1u << 12, IT_REGEN = 1u << 13, IT_ZERO_MASK = 1u << 14, IT_FROM_POOL = 1u << 15 };

template <int L> struct Cfg {
    static constexpr int MAXR = (L == 32) ? 32 : 16;
    static constexpr int MAXP = MAXR * L;
    static constexpr int ML = (2 * L > 32) ? 2 * L : 32;  // line-table capacity per cascade round
    static constexpr int MLEN = (L > MAXR) ? L : MAXR;    // longest straight line
    static constexpr int DFS = (4 * L > 64) ? 4 * L : 64; // activation stack depth
    static constexpr int NW = 4 * (L - 1);                // stream words produced per Philox pass
    static constexpr unsigned LMASK = (L == 32) ? 0xffffffffu : ((1u << (L & 31)) - 1u);
    static constexpr int THREADS = 128;
    static constexpr int GPW = 32 / L;                    // groups per warp: 4, 3 (L = 10, two lanes idle), 2 or 1
    static constexpr int GPB = (THREADS / 32) * GPW;      // groups (boards) per block
};

template <int L> struct __align__(16) GroupSmem {
    int8_t board[2 * Cfg<L>::MAXP];  // colour plane [0,P), type plane [P,2P) -- the HBM layout
    uint8_t mask[2 * Cfg<L>::MAXP];  // legal-move mask staging (A < 2P)
    uint32_t wbuf[4 * L];            // stream words of the current Philox pass
    uint32_t stack[Cfg<L>::DFS];     // activation DFS frames
    uint32_t line_key[Cfg<L>::ML];   // (top row << 12) | list position  -> processing order
    uint32_t line_mask[Cfg<L>::ML];  // a line is straight: its columns (horizontal) or rows (vertical) as a bit set
    uint16_t cq_pos[Cfg<L>::ML];     // special-creation queue (ref :411)
    uint16_t taken[Cfg<L>::ML];
    uint16_t cnt[32];                // colour histogram for the cookie (ref :536)
    uint8_t line_kind[Cfg<L>::ML];   // 0 horizontal (line_idx = row), 1 vertical (line_idx = column)
    uint8_t line_idx[Cfg<L>::ML];
    uint8_t line_colour[Cfg<L>::ML];
    uint8_t order[Cfg<L>::ML];
    int8_t cq_type[Cfg<L>::ML];
    uint8_t cq_colour[Cfg<L>::ML];
};

// ---- Philox4x32-10 (Random123) ---------------------------------------------------------------------
template <int U = UNROLL_PHILOX>
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
    // rolled: the instruction caches (L0 ~6 KB, L1.5 32 KB per SM) bound this engine, not the issue rate
#pragma unroll U
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ bool not01(int t) { return t != 0 && t != 1; }

// ---- vector copies between global and the group's shared memory -------------------------------------
// (not inlined: the width switch would otherwise be replicated at every load/store site)
template <int L, typename V> __device__ __forceinline__ void copy_vec(void* dst, const void* src, int nbytes, int lane) {
    const int n = nbytes / (int)sizeof(V);
    V* d = reinterpret_cast<V*>(dst);
    const V* s = reinterpret_cast<const V*>(src);
#pragma unroll 1
    for (int i = lane; i < n; i += L) d[i] = s[i];
}
template <int L> __device__ __noinline__ void copy_bytes(void* dst, const void* src, int nbytes, int vecw, int lane) {
    switch (vecw) {
        case 16: copy_vec<L, uint4>(dst, src, nbytes, lane); break;
        case 8: copy_vec<L, uint2>(dst, src, nbytes, lane); break;
        case 4: copy_vec<L, uint32_t>(dst, src, nbytes, lane); break;
        case 2: copy_vec<L, uint16_t>(dst, src, nbytes, lane); break;
        default: copy_vec<L, uint8_t>(dst, src, nbytes, lane); break;
    }
}
// global -> shared copy with L2-only loads (ld.global.cg): for data another kernel may have written while this one runs
// (an L1 line shared with a neighbouring env's entry could be stale)
template <int L> __device__ __noinline__ void copy_bytes_cg(void* dst, const void* src, int nbytes, int vecw, int lane) {
#ifdef TMG_EMU
    copy_bytes<L>(dst, src, nbytes, vecw, lane);
#else
    if (vecw >= 4) {
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
#pragma unroll 1
        for (int i = lane; i < nbytes / 4; i += L) d[i] = __ldcg(s + i);
    } else {
        uint8_t* d = reinterpret_cast<uint8_t*>(dst);
        const uint8_t* s = reinterpret_cast<const uint8_t*>(src);
#pragma unroll 1
        for (int i = lane; i < nbytes; i += L) d[i] = __ldcg(s + i);
    }
#endif
}
template <int L> __device__ __forceinline__ void zero_bytes(void* dst, int nbytes, int vecw, int lane) {
    if (vecw >= 4) {
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        for (int i = lane; i < nbytes / 4; i += L) d[i] = 0u;
    } else {
        uint8_t* d = reinterpret_cast<uint8_t*>(dst);
        for (int i = lane; i < nbytes; i += L) d[i] = 0;
    }
}

// compile-time sized copy for the fixed board shapes: no call, no loop, the widest vector the size allows
template <int W> struct VecOf { typedef uint8_t T; };
template <> struct VecOf<2> { typedef uint16_t T; };
template <> struct VecOf<4> { typedef uint32_t T; };
template <> struct VecOf<8> { typedef uint2 T; };
template <> struct VecOf<16> { typedef uint4 T; };
__host__ __device__ constexpr int vecw_of(int nbytes) { return nbytes % 16 == 0 ? 16 : nbytes % 8 == 0 ? 8 : nbytes % 4 == 0 ? 4 : nbytes % 2 == 0 ? 2 : 1; }
template <int L, int NBYTES> __device__ __forceinline__ void copy_fixed(void* dst, const void* src, int lane) {
    constexpr int W = vecw_of(NBYTES), n = NBYTES / W;
    typedef typename VecOf<W>::T V;
    V* d = reinterpret_cast<V*>(dst);
    const V* s = reinterpret_cast<const V*>(src);
#pragma unroll
    for (int i = 0; i < (n + L - 1) / L; ++i) {
        const int k = lane + i * L;
        if (k < n) d[k] = s[k];
    }
}

// k-th pre-drawn colour of an env in injected mode, -1 when the stream is exhausted
__device__ __noinline__ int injected_draw(const uint8_t* inj, long long inj_len, int env, long long q) {
    if (q < inj_len) return inj[(size_t)env * (size_t)inj_len + (size_t)q];
    return -1;
}

// is_move_effective (ref :735-787): the literal window rule, evaluated by ONE lane with a virtual swap
// (the board is not touched).  Used for hand-made / unstable boards and by the debug entry point.
__device__ __noinline__ bool effective_literal(const int8_t* col, const int8_t* typ, int R, int C, int i1, int i2) {
    const int ta = typ[i1], tb = typ[i2];
    if (not01(ta) && not01(tb)) return true;             // ref :750
    if (ta < 0 || tb < 0) return true;                   // ref :754
    const int r1 = i1 / C, c1 = i1 % C, r2 = i2 / C, c2 = i2 % C;
    const int rmin_ = max(0, min(r1, r2) - 2), rmax_ = min(R - 1, max(r1, r2) + 2);  // ref :758-761
    const int cmin_ = max(0, min(c1, c2) - 2), cmax_ = min(C - 1, max(c1, c2) + 2);
    auto cs = [&](int i) -> int { return i == i1 ? col[i2] : (i == i2 ? col[i1] : col[i]); };
    auto ts = [&](int i) -> int { return i == i1 ? typ[i2] : (i == i2 ? typ[i1] : typ[i]); };
    if (cmin_ + 2 <= cmax_)                               // ref :767-771
        for (int r = rmin_; r <= rmax_; ++r)
            for (int c = cmin_; c + 2 <= cmax_; ++c) {
                const int i = r * C + c;
                if (cs(i) == cs(i + 1) && cs(i + 1) == cs(i + 2) && ts(i + 2) >= 0) return true;
            }
    if (rmin_ + 2 <= rmax_)                               // ref :777-781
        for (int r = rmin_; r + 2 <= rmax_; ++r)
            for (int c = cmin_; c <= cmax_; ++c) {
                const int i = r * C + c;
                if (cs(i) == cs(i + C) && cs(i + C) == cs(i + 2 * C) && ts(i + 2 * C) >= 0) return true;
            }
    return false;
}

// literal per-action mask for one column (boards whose mask depends on the exact window rule; rare)
struct MaskPair { unsigned v, h; };
__device__ __noinline__ MaskPair mask_literal_column(const int8_t* col, const int8_t* typ, int R, int C, int lane) {
    MaskPair o;
    o.v = 0u; o.h = 0u;
    if (lane < C)
        for (int r = 0; r < R; ++r) {
            const int i = r * C + lane;
            if (r + 1 < R) o.v |= (unsigned)effective_literal(col, typ, R, C, i, i + C) << r;
            if (lane + 1 < C) o.h |= (unsigned)effective_literal(col, typ, R, C, i, i + 1) << r;
        }
    return o;
}

// shuffle (ref :114-118): new[i] = old[perm[i]], perm = Fisher-Yates of arange(P) on stream 1.  Applying the
// same swaps to the cells themselves yields exactly old[perm[i]].  One lane.
// `episode` >= 0 selects the reset shuffle stream of that board (ctr = (blk, episode, env, 4)).
__device__ __noinline__ void shuffle_serial(int8_t* col, int8_t* typ, int P, uint32_t gid, uint32_t key0, uint32_t key1,
                                            uint64_t scur, long long episode) {
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    uint64_t have_blk = ~0ull;
    for (int i = P - 1; i >= 1; --i) {
        const uint64_t k = scur++;
        const uint64_t blk = k >> 2;
        if (blk != have_blk) {
            if (episode >= 0) philox4x32_10((uint32_t)blk, (uint32_t)episode, gid, 4u, key0, key1, w);
            else philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), gid, 1u, key0, key1, w);
            have_blk = blk;
        }
        const uint32_t word = (k & 3) == 0 ? w[0] : (k & 3) == 1 ? w[1] : (k & 3) == 2 ? w[2] : w[3];
        const int j = (int)__umulhi(word, (uint32_t)(i + 1));
        const int8_t a = col[i], b = typ[i];
        col[i] = col[j]; typ[i] = typ[j];
        col[j] = a; typ[j] = b;
    }
}

// packed rows of a fresh (all-normal) board: one word per row, BITS bits per cell, value = colour - 1
template <int BITS> __device__ __forceinline__ uint32_t packed_eq(uint32_t t) {   // cells of t that are all-zero -> their low bit
    if (BITS == 2) return ~(t | (t >> 1)) & 0x55555555u;
    return ~(t | (t >> 1) | (t >> 2)) & 0x09249249u;
}

// ======================================================================================================
// One board owned by a group of L lanes
// ======================================================================================================
// The rare, large parts of a cascade round run out of line on a private Board view, so that the common round
// (scan, fast path, gravity, refill) stays small enough for the instruction caches.
struct SlowOut { int n, n_new, n_act; uint32_t status; };
struct SlowScan {   // Board::Scan by value
    int rstar; unsigned mv, hs, hcells, m; int vtop; bool has_v; unsigned E, D, T, S;
};
template <int L, int RT, int CT>
__device__ __noinline__ SlowOut slow_round(GroupSmem<L>* sm, const Params* pp, int lane, unsigned gmask, int gshift, int env,
                                           int n_new, int n_act, SlowScan sc);
template <int L, int RT, int CT>
__device__ __noinline__ SlowOut slow_combination(GroupSmem<L>* sm, const Params* pp, int lane, unsigned gmask, int gshift,
                                                 int env, int i1, int i2);
// the cold ends of a work-list item: playability of a moved board whose mask is not known (shuffle / literal rule) and
// generate_board inside the step -- out of line so that the step kernel's hot code stays small (it is bound by instruction
// supply: every KB of inlined cold code that was removed from it measured as throughput)
struct FinishOut { unsigned effv, effh; uint32_t status; int shuffled; uint64_t dcur, scur; unsigned last_S; };
// host-mirror write-through of one env (tmg_host_bind) and the all-zero mask of a terminal step: out of line as well
enum { MIRROR_BOARD_SMEM = 1, MIRROR_BOARD_POOL = 2, MIRROR_MASK_SMEM = 4, MIRROR_MASK_POOL = 8, MIRROR_MASK_ZERO = 16, STORE_ZERO_MASK = 32 };
template <int L, int RT, int CT>
__device__ __noinline__ void mirror_item(GroupSmem<L>* sm, const Params* pp, int lane, unsigned gmask, int gshift, int env, int what);
template <int L, int RT, int CT>
__device__ __noinline__ FinishOut slow_finish(GroupSmem<L>* sm, const Params* pp, int lane, unsigned gmask, int gshift, int env,
                                              int do_playability, int do_generate, int next_ep, uint64_t dcur, uint64_t scur);

// RT/CT > 0 fix the board shape at compile time (full unrolling, immediate shared-memory offsets); 0 = runtime shape.
template <int L, int RT = 0, int CT = 0> struct Board {
    typedef Cfg<L> CF;
    GroupSmem<L>& s;
    const Params& p;
    const int lane;
    const unsigned gmask;
    const int gshift;
    int env;
    const int R, C, P, K;
    int8_t* const col;
    int8_t* const typ;
    uint64_t dcur, scur;
    uint32_t gid;
    uint32_t status;
    int n_new, n_act;  // counters, uniform after broadcast (ref :343-344)
    // Deferred deletions of a fast-path round: the cells [fg_top, fg_top + fg_len) of this lane's column.  They are not
    // zeroed -- fall_and_refill shifts the rows above them down over them and refills the top (fixed small shapes only).
    int fg_top = 0, fg_len = 0;
    bool fg_valid = false;
    static constexpr bool DEFER_GAPS = TMG_FUSED_FALL && RT > 0 && RT <= 16;
    // Unroll factors.  Where the byte planes ARE the hot path (boards above 10 rows, the runtime shape) the row loops and
    // Philox stay unrolled (rolled: config 5 fell from 71 M to 46 M env-steps/s); in the instantiations whose moves run on
    // the register-resident engine this code is the cold rest of the kernel and is kept small (TMG_UNROLL_*).
    static constexpr bool BYTES_HOT = RT == 0 || RT > 10;
    static constexpr int UR = BYTES_HOT ? 32 : UNROLL_ROWS, PHU = BYTES_HOT ? 10 : UNROLL_PHILOX;
    unsigned last_S = 0u;  // special tiles of this lane's column as of the last mask_bits (scheduling hint, see n_special)
    uint32_t prof_serial = 0u, prof_rounds = 0u, prof_iters = 0u;  // diagnostics (written only when p.prof is set)
    const uint32_t specials;   // copies of the Params fields the round code needs (no pointer chasing out of line)
    const bool prof_on;
    // While a board is generated the draws come from the episode-indexed reset streams (see include/tmg_b200.h):
    bool in_reset = false;
    uint32_t episode = 0u;
    uint64_t rdc = 0ull, rsc = 0ull;

    __device__ Board(GroupSmem<L>& sm, const Params& pp, int lane_, unsigned gmask_, int gshift_, int env_)
        : s(sm), p(pp), lane(lane_), gmask(gmask_), gshift(gshift_), env(env_), R(RT ? RT : pp.R), C(CT ? CT : pp.C),
          P(RT ? RT * CT : pp.P), K(pp.K), col(sm.board), typ(sm.board + (RT ? RT * CT : pp.P)), dcur(0), scur(0),
          gid((uint32_t)(pp.env_id_offset + (uint64_t)env_)), status(0), n_new(0), n_act(0), specials(pp.specials),
          prof_on(pp.prof != nullptr) {}

    // the same group moves on to another env (persistent loops that keep one Board object across their iterations)
    __device__ __forceinline__ void rebind(int env_) {
        env = env_;
        gid = (uint32_t)(p.env_id_offset + (uint64_t)env_);
        status = 0u; n_new = 0; n_act = 0; dcur = 0ull; scur = 0ull;
    }

    // ---- group collectives ---------------------------------------------------------------------------
    // a 32-lane group is the whole warp: a literal full mask lets the compiler emit the collectives without the
    // runtime membership check (MATCH.ANY / REDUX / VOTEU + divergent-path trampoline) a group mask in a register needs
    __device__ __forceinline__ unsigned gm() const { return L == 32 ? 0xffffffffu : gmask; }
    __device__ __forceinline__ unsigned ballot(bool pr TMG_SITE_P) const { TMG_SITE_SET return L == 32 ? __ballot_sync(0xffffffffu, pr) : ((__ballot_sync(gmask, pr) >> gshift) & CF::LMASK); }
    __device__ __forceinline__ void sync(TMG_SITE_P0) const { TMG_SITE_SET __syncwarp(gm()); }
    // groups need not be a power of two wide (10-lane groups: three 10-column boards per warp), so shuffles address
    // absolute lanes of the warp
    __device__ __forceinline__ int shfl(int v, int src TMG_SITE_P) const { TMG_SITE_SET return __shfl_sync(gm(), v, (L == 32 ? 0 : gshift) + src); }
    __device__ __forceinline__ int radd(int v TMG_SITE_P) const { TMG_SITE_SET return __reduce_add_sync(gm(), v); }
    __device__ __forceinline__ int rmax(int v TMG_SITE_P) const { TMG_SITE_SET return __reduce_max_sync(gm(), v); }
    __device__ __forceinline__ int rmin(int v TMG_SITE_P) const { TMG_SITE_SET return __reduce_min_sync(gm(), v); }
    __device__ __forceinline__ unsigned ror(unsigned v TMG_SITE_P) const { TMG_SITE_SET return __reduce_or_sync(gm(), v); }
    __device__ __forceinline__ unsigned lt_mask() const { return (1u << lane) - 1u; }
    // neighbour-lane bitboards: value of lane+d / lane-d, 0 outside [0,L)
    __device__ __forceinline__ unsigned from_right(unsigned v, int d TMG_SITE_P) const {
        TMG_SITE_SET
        const unsigned r = __shfl_down_sync(gm(), v, d);
        return (lane + d < L) ? r : 0u;
    }
    __device__ __forceinline__ unsigned from_left(unsigned v, int d TMG_SITE_P) const {
        TMG_SITE_SET
        const unsigned r = __shfl_up_sync(gm(), v, d);
        return (lane - d >= 0) ? r : 0u;
    }
    __device__ __forceinline__ unsigned rows_mask() const { return R >= 32 ? 0xffffffffu : ((1u << R) - 1u); }

    // ---- board / mask sized copies (compile-time size for the fixed shapes) ------------------------------------
    static constexpr int NB_FIXED = RT ? 2 * RT * CT : 1, NA_FIXED = RT ? 2 * RT * CT - RT - CT : 1;
    __device__ __forceinline__ void copy_board(void* dst, const void* src, int vecw) const {
        if (RT && vecw == vecw_of(NB_FIXED)) copy_fixed<L, NB_FIXED>(dst, src, lane);
        else copy_bytes<L>(dst, src, 2 * P, vecw, lane);
    }
    __device__ __forceinline__ void copy_mask(void* dst, const void* src) const {
        if (RT && p.mask_vecw == vecw_of(NA_FIXED)) copy_fixed<L, NA_FIXED>(dst, src, lane);
        else copy_bytes<L>(dst, src, p.A, p.mask_vecw, lane);
    }

    // ---- state I/O -----------------------------------------------------------------------------------
    __device__ __forceinline__ void load_board(const int8_t* src, int vecw) {
        copy_board(s.board, src + (size_t)env * 2 * P, vecw);
        sync();
    }
    __device__ __forceinline__ void store_board() {
        sync();
        copy_board(p.board + (size_t)env * 2 * P, s.board, p.board_vecw);
    }
    __device__ __forceinline__ void store_mask() {
        sync();
        copy_mask(p.mask + (size_t)env * p.A, s.mask);
    }
    __device__ __forceinline__ void store_zero_mask() { zero_bytes<L>(p.mask + (size_t)env * p.A, p.A, p.mask_vecw, lane); }
    // host mirror: this env's mask (s.mask, or all zero) as bytes and / or bits, straight into page-locked host memory
    __device__ __forceinline__ void mirror_mask(bool zero) {
        if (p.h_mask) {
            if (zero) zero_bytes<L>(p.h_mask + (size_t)env * p.A, p.A, p.mask_vecw, lane);
            else copy_mask(p.h_mask + (size_t)env * p.A, s.mask);
        }
        if (p.h_mask_bits) {
            const int bpe = (p.A + 7) >> 3;
            uint8_t* dst = p.h_mask_bits + (size_t)env * bpe;
#pragma unroll 1
            for (int b = lane; b < bpe; b += L) {
                unsigned v = 0u;
                if (!zero) {
                    const uint32_t* m32 = reinterpret_cast<const uint32_t*>(s.mask + 8 * b);
                    // bytes 0/1 at bits 0,8,16,24 -> 4 adjacent bits: the products land on distinct bit positions
                    v = ((((m32[0] & 0x01010101u) * 0x01020408u) >> 24) & 0xfu) | (((((m32[1] & 0x01010101u) * 0x01020408u) >> 24) & 0xfu) << 4);
                    const int valid = p.A - 8 * b;
                    if (valid < 8) v &= (1u << valid) - 1u;
                }
                dst[b] = (uint8_t)v;
            }
        }
    }
    // host mirror, packed form: one byte per cell, colour | (type & 7) << 4, from the byte planes at `src` (shared or global)
    __device__ __forceinline__ void mirror_board_packed(const int8_t* src) {
        uint8_t* dst = p.h_board_packed + (size_t)env * P;
        if ((P & 3) == 0) {
            const uint32_t* c4 = reinterpret_cast<const uint32_t*>(src);
            const uint32_t* t4 = reinterpret_cast<const uint32_t*>(src + P);
            uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
#pragma unroll 1
            for (int j = lane; j < (P >> 2); j += L) d4[j] = (c4[j] & 0x0f0f0f0fu) | ((t4[j] & 0x07070707u) << 4);
        } else {
#pragma unroll 1
            for (int i = lane; i < P; i += L) dst[i] = (uint8_t)((src[i] & 15) | ((src[P + i] & 7) << 4));
        }
    }
    __device__ __forceinline__ void load_cursors() { dcur = p.draw_cursor[env]; scur = p.shuffle_cursor[env]; }
    __device__ __forceinline__ void store_cursors() {
        if (lane == 0) { p.draw_cursor[env] = dcur; p.shuffle_cursor[env] = scur; }
    }

    // ---- draw stream -----------------------------------------------------------------------------------
    // words [start, start+n) of stream `stream` -> s.wbuf[0..n), n <= NW.  All lanes call.
    __device__ __forceinline__ void fill_words(uint32_t stream, uint64_t start, int n) {
        const uint64_t b0 = start >> 2;
        const int nb = (int)(((start + (uint64_t)n - 1) >> 2) - b0) + 1;  // <= L because n <= 4(L-1)
        if (lane < nb) {
            const uint64_t b = b0 + (uint64_t)lane;
            uint32_t w[4];
            philox4x32_10<PHU>((uint32_t)b, (uint32_t)(b >> 32), gid, stream, p.key0, p.key1, w);
            const int base = (int)((long long)(b << 2) - (long long)start);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = base + i;
                if (idx >= 0 && idx < n) s.wbuf[idx] = w[i];
            }
        }
        sync();
    }
    __device__ __forceinline__ int injected_colour(int k) {
        const int v = injected_draw(p.inj, p.inj_len, env, (long long)dcur + k);
        if (v < 0) { status |= ST_DRAWS_EXHAUSTED; return 1; }
        return v;
    }
    // cells [0, n) in row-major order <- next n draws (initial fill ref :97 / row-block redraw ref :129).
    // Cell i takes word dcur+i, so each lane turns whole Philox blocks straight into board bytes.
    __device__ void draw_cells(int n, bool set_type) {
        sync();
        if (set_type) {
            uint32_t* t32 = reinterpret_cast<uint32_t*>(s.board);  // type plane may start unaligned: bytes at the edges
            (void)t32;
            for (int i = lane; i < n; i += L) typ[i] = 1;
        }
        if (p.use_inj) {
            for (int i = lane; i < n; i += L) col[i] = (int8_t)injected_colour(i);
        } else {
            const uint64_t cur = in_reset ? rdc : dcur;
            const uint64_t b0 = cur >> 2, b1 = (cur + (uint64_t)n - 1) >> 2;
            for (uint64_t b = b0 + (uint64_t)lane; b <= b1; b += L) {
                uint32_t w[4];
                philox4x32_10<PHU>((uint32_t)b, in_reset ? episode : (uint32_t)(b >> 32), gid, in_reset ? 3u : 0u, p.key0, p.key1, w);
                const int base = (int)((long long)(b << 2) - (long long)cur);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int idx = base + i;
                    if (idx >= 0 && idx < n) col[idx] = (int8_t)(1 + (int)__umulhi(w[i], (uint32_t)K));
                }
            }
        }
        if (in_reset) rdc += (uint64_t)n;
        else dcur += (uint64_t)n;
        sync();
    }

    // ---- gravity (ref :217-229) + count of type==0 cells (ref :362,374) ------------------------------------
    // returns the number of empty cells now on top of this lane's column; *elim gets P - count_nonzero(type)
    __device__ int gravity(int* elim) {
        sync();
        int e = 0, nz = 0;
        if (lane < C) {
            int w = R - 1;
#pragma unroll UR
            for (int r = R - 1; r >= 0; --r) {
                const int i = r * C + lane;
                const int x = col[i], t = typ[i];
                nz += (t == 0);
                if (x != 0 || t != 0) {
                    if (w != r) { col[w * C + lane] = (int8_t)x; typ[w * C + lane] = (int8_t)t; }
                    --w;
                }
            }
            e = w + 1;
            for (; w >= 0; --w) { col[w * C + lane] = 0; typ[w * C + lane] = 0; }
        }
        if (elim) *elim = radd(nz);
        return e;
    }

    // ---- refill (ref :231-241): the k-th draw goes to the k-th empty cell in row-major order ---------------
    __device__ void refill(int e) {
        const int maxe = rmax(e);
        if (maxe == 0) return;  // ref :238: no rng call when nothing is empty
        const int total = radd(e);
        const unsigned lt = lt_mask();
        for (int ps = 0; ps < total; ps += CF::NW) {
            const int nw = min(CF::NW, total - ps);
            if (!p.use_inj) fill_words(0u, dcur + (uint64_t)ps, nw);
            int base = 0;
#pragma unroll 1
            for (int r = 0; r < maxe; ++r) {
                const unsigned m = ballot(e > r);
                if (e > r) {
                    const int rank = base + __popc(m & lt);
                    if (rank >= ps && rank < ps + nw) {
                        col[r * C + lane] = (int8_t)(p.use_inj ? injected_colour(rank)
                                                               : 1 + (int)__umulhi(s.wbuf[rank - ps], (uint32_t)K));
                        typ[r * C + lane] = 1;
                    }
                }
                base += __popc(m);
            }
            sync();
        }
        dcur += (uint64_t)total;
    }

    // ---- gravity + refill of a cascade round, fused (ref :362-364, :374-376) --------------------------------------
    // Fixed small shapes: the column is read into registers once, every surviving cell moves down by the number of
    // empties below it (independent stores, no carried write pointer), the emptied top rows are not zeroed because the
    // refill overwrites them, and the Philox block each lane will need is computed beside the loads.
    __device__ __forceinline__ void fall_and_refill(int& elim_add) {
        if (!TMG_FUSED_FALL || RT == 0 || RT > 16) {
            int cnt;
            const int e = gravity(&cnt);
            elim_add = cnt;
            refill(e);
            return;
        }
        constexpr int RR = (RT > 0 && RT <= 16) ? RT : 1;
        sync();
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        const uint64_t b0 = dcur >> 2;
        const int off = (int)(dcur & 3ull);
        if (!p.use_inj) {
            const uint64_t b = b0 + (uint64_t)lane;
            philox4x32_10<PHU>((uint32_t)b, (uint32_t)(b >> 32), gid, 0u, p.key0, p.key1, w);
        }
        int e;
        if (fg_valid) {
            // after a fast-path round every column has at most one gap and the round knows it: shift the rows above it down
            fg_valid = false;
            const int len = fg_len, top = fg_top;
            const int mt = rmax(len ? top : 0);
#pragma unroll 1
            for (int r = mt - 1; r >= 0; --r) {
                if (len && r < top) {
                    col[(r + len) * C + lane] = col[r * C + lane];
                    typ[(r + len) * C + lane] = typ[r * C + lane];
                }
            }
            e = len;
            elim_add = radd(len);                                  // ref :374: the deleted cells are the type-0 cells
        } else {
            int x[RR], t[RR];
            unsigned empt = 0u, tz = 0u;
            if (lane < C) {
#pragma unroll
                for (int r = 0; r < RR; ++r) {
                    x[r] = col[r * C + lane];
                    t[r] = typ[r * C + lane];
                    empt |= (unsigned)(x[r] == 0 && t[r] == 0) << r;
                    tz |= (unsigned)(t[r] == 0) << r;
                }
                if (empt) {
#pragma unroll
                    for (int r = RR - 1; r >= 0; --r) {
                        const int below = __popc(empt >> (r + 1));     // empties underneath: the cell falls that far
                        if (!((empt >> r) & 1u) && below) {
                            col[(r + below) * C + lane] = (int8_t)x[r];
                            typ[(r + below) * C + lane] = (int8_t)t[r];
                        }
                    }
                }
            }
            e = __popc(empt);
            elim_add = radd(__popc(tz));                           // ref :362,374: P - count_nonzero(type)
        }
        unsigned m = ballot(e > 0);
        if (!m) return;                                            // ref :238: no rng call when nothing is empty
        if (!p.use_inj) {
#pragma unroll
            for (int i = 0; i < 4; ++i) s.wbuf[4 * lane + i] = w[i];
        }
        sync();
        const unsigned lt = lt_mask();
        const int avail = 4 * L - off;                             // words of this pass, starting at dcur
        int base = 0;
#pragma unroll 1
        for (int r = 0; m; ++r) {                                  // ref :239-241: k-th draw -> k-th empty cell, row-major
            if (e > r) {
                const int rank = base + __popc(m & lt);
                if (rank < avail) {
                    col[r * C + lane] = (int8_t)(p.use_inj ? injected_colour(rank)
                                                           : 1 + (int)__umulhi(s.wbuf[off + rank], (uint32_t)K));
                    typ[r * C + lane] = 1;
                }
            }
            base += __popc(m);
            m = ballot(e > r + 1);
        }
        const int total = base;
#pragma unroll 1
        for (int ps = avail; ps < total; ps += CF::NW) {           // rare: more empties than one pass of words
            const int nw = min(CF::NW, total - ps);
            sync();
            if (!p.use_inj) fill_words(0u, dcur + (uint64_t)ps, nw);
            int bs = 0;
#pragma unroll 1
            for (int r = 0; r < RR; ++r) {
                const unsigned mm = ballot(e > r);
                if (e > r) {
                    const int rank = bs + __popc(mm & lt);
                    if (rank >= ps && rank < ps + nw) {
                        col[r * C + lane] = (int8_t)(p.use_inj ? injected_colour(rank)
                                                               : 1 + (int)__umulhi(s.wbuf[rank - ps], (uint32_t)K));
                        typ[r * C + lane] = 1;
                    }
                }
                bs += __popc(mm);
            }
        }
        dcur += (uint64_t)total;
        sync();
    }

    // ---- line detection (ref :149-215) ---------------------------------------------------------------------
    // Per-lane row bitboards of column c: E bit r: colour(r,c)==colour(r,c+1); D bit r: colour(r,c)==colour(r-1,c);
    // T bit r: type(r,c) > 0.  One pass over the column finds every anchored line of the board at once.
    struct Bits { unsigned E, D, T, S; };  // S bit r: type(r,c) not in {0,1} (a special tile)
    __device__ __forceinline__ Bits column_bits(bool all_normal) const {
        Bits b;
        b.E = b.D = b.T = b.S = 0u;
        if (L == 32 && CT > 0 && CT <= 16 && RT > 1) {
            // one board per warp: the upper half of the rows on lanes 0..C-1, the lower half on lanes 16..16+C-1, combined
            // with one xor-shuffle per bitboard -- half the serial row loop
            constexpr int RH = (RT + 1) / 2;
            const int c = lane & 15, half = lane >> 4;
            if (c < C) {
                const int r0 = half ? RH : 0;
                int prev = half ? (int)col[(RH - 1) * C + c] : -3;
                const bool has_right = c + 1 < C;
#pragma unroll
                for (int k = 0; k < RH; ++k) {
                    const int r = r0 + k;
                    if (r < RT) {
                        const int i = r * C + c;
                        const int x = col[i];
                        const int xr = has_right ? (int)col[i + 1] : -2;
                        b.E |= (unsigned)(x == xr) << r;
                        b.D |= (unsigned)(x == prev) << r;
                        if (!all_normal) {
                            const int t = typ[i];
                            b.T |= (unsigned)(t > 0) << r;
                            b.S |= (unsigned)not01(t) << r;
                        }
                        prev = x;
                    }
                }
            }
            TMG_SITE_HERE
            b.E |= __shfl_xor_sync(0xffffffffu, b.E, 16);
            b.D |= __shfl_xor_sync(0xffffffffu, b.D, 16);
            if (!all_normal) {
                b.T |= __shfl_xor_sync(0xffffffffu, b.T, 16);
                b.S |= __shfl_xor_sync(0xffffffffu, b.S, 16);
            }
            if (lane >= C) { b.E = b.D = b.T = b.S = 0u; }
            else if (all_normal) b.T = rows_mask();
            return b;
        }
        if (lane < C) {
            int prev = -3;
            const bool has_right = lane + 1 < C;
#pragma unroll UR
            for (int r = 0; r < (RT ? RT : R); ++r) {
                const int i = r * C + lane;
                const int x = col[i];
                const int xr = has_right ? (int)col[i + 1] : -2;
                b.E |= (unsigned)(x == xr) << r;
                b.D |= (unsigned)(x == prev) << r;
                if (!all_normal) {
                    const int t = typ[i];
                    b.T |= (unsigned)(t > 0) << r;
                    b.S |= (unsigned)not01(t) << r;
                }
                prev = x;
            }
            if (all_normal) b.T = rows_mask();
        }
        return b;
    }
    struct Scan {
        int rstar;        // row of the bottom-most anchored lines, -1 if the board has none (ref :158-160)
        unsigned mv;      // lanes with a vertical line whose anchor (bottom) is at rstar
        unsigned hs;      // start columns of the horizontal lines in row rstar
        unsigned hcells;  // all cells of those horizontal lines
        unsigned m;       // bit c: colour(rstar,c) == colour(rstar,c+1)
        int vtop;         // per lane: top row of its vertical line
        bool has_v;
        Bits bits;
    };
    // rows above `from` are known to be line-free
    __device__ Scan scan_lines(int from, bool all_normal) {
        Scan o;
        o.bits = column_bits(all_normal);
        const Bits& b = o.bits;
        const unsigned V = b.D & (b.D << 1) & b.T;            // vertical triple anchored (bottom) at r, anchor type > 0 (ref :163-173)
        const unsigned H = b.E & from_right(b.E, 1) & b.T;    // horizontal triple anchored (left) at (r,c)   (ref :179-189)
        const unsigned lim = (from >= 31) ? 0xffffffffu : ((2u << from) - 1u);
        const unsigned F = (V | H) & lim;
        o.rstar = rmax(F ? 31 - __clz((int)F) : -1);
        o.mv = o.hs = o.hcells = o.m = 0u; o.vtop = 0; o.has_v = false;
        if (o.rstar < 0) return o;
        const int rs = o.rstar;
        const unsigned m = ballot((b.E >> rs) & 1u);
        const unsigned T = ballot((b.T >> rs) & 1u);
        o.has_v = (V >> rs) & 1u;
        o.mv = ballot(o.has_v);
        unsigned cand = m & (m >> 1) & T, hs = 0u, hcells = 0u;
        while (cand) {  // left to right; cells of a found line cannot anchor another (ref :179,192)
            const int sidx = __ffs((int)cand) - 1;
            const int run = __ffs((int)~(m >> sidx)) - 1;  // line = sidx .. sidx+run
            const unsigned cells = ((2u << run) - 1u) << sidx;
            hs |= 1u << sidx;
            hcells |= cells;
            cand &= ~cells;
        }
        o.hs = hs; o.hcells = hcells; o.m = m;
        if (o.has_v) o.vtop = rs - __clz((int)~(b.D << (31 - rs)));  // extend upwards while equal (ref :168-172)
        return o;
    }
    // row of l[0][0] for the first line of get_colour_lines() (ref :127-128)
    __device__ __forceinline__ int first_line_top(const Scan& sc) {
        const int cv = sc.mv ? __ffs((int)sc.mv) - 1 : 64;
        const int ch = sc.hs ? __ffs((int)sc.hs) - 1 : 64;
        const int vt = shfl(sc.vtop, cv < 64 ? cv : 0);
        return (cv <= ch) ? vt : sc.rstar;  // vertical is listed before horizontal at the same column (ref :163,179)
    }

    // ---- line table in the reference's list order (closed form in SURVEY.md A.4) ---------------------------
    // returns the number of lines; keys give the processing order of process_colour_lines (ref :282)
    __device__ int build_line_table(const Scan& sc) {
        const int rs = sc.rstar;
        const unsigned lt = lt_mask();
        // phase 1: by column, vertical before horizontal
        const int before = __popc(sc.mv & lt) + __popc(sc.hs & lt);
        int n = __popc(sc.mv) + __popc(sc.hs);
        if (sc.has_v) {
            const int slot = before;
            if (slot < CF::ML) {
                s.line_mask[slot] = ((2u << rs) - 1u) & ~((1u << sc.vtop) - 1u);   // rows vtop..rs
                s.line_kind[slot] = 1;
                s.line_idx[slot] = (uint8_t)lane;
                s.line_colour[slot] = (uint8_t)col[rs * C + lane];
                s.line_key[slot] = ((uint32_t)sc.vtop << 12) | (uint32_t)slot;
            }
        }
        if ((sc.hs >> lane) & 1u) {
            const int slot = before + (sc.has_v ? 1 : 0);
            if (slot < CF::ML) {
                const int run = __ffs((int)~(sc.m >> lane)) - 1;
                s.line_mask[slot] = ((2u << run) - 1u) << lane;                        // columns lane..lane+run
                s.line_kind[slot] = 0;
                s.line_idx[slot] = (uint8_t)rs;
                s.line_colour[slot] = (uint8_t)col[rs * C + lane];
                s.line_key[slot] = ((uint32_t)rs << 12) | (uint32_t)slot;
            }
        }
        // phase 2 (ref :198-214): horizontal segments through the cells of the vertical lines, cut at phase-1 cells
        if (sc.mv) {
            const int rmin_ = rmin(sc.has_v ? sc.vtop : 1 << 20);
#pragma unroll 1
            for (int r = rmin_; r <= rs; ++r) {
                const unsigned m = ballot((sc.bits.E >> r) & 1u);
                const unsigned T = ballot((sc.bits.T >> r) & 1u);
                const bool origin = sc.has_v && sc.vtop <= r;
                unsigned Q = ballot(origin);
                if (r == rs) Q |= sc.hcells;
                const unsigned pass = T & ~Q;
                int left = 0, right = 0;
                bool seg = false;
                if (origin && ((sc.bits.T >> r) & 1u)) {  // match_color needs type > 0 on both cells (ref :199)
                    const unsigned chain_r = (m << 1) & pass;  // bit j: cell j equals cell j-1 and may be entered
                    const unsigned chain_l = m & pass;         // bit j: cell j equals cell j+1 and may be entered
                    if (lane + 1 < L) right = __ffs((int)~(chain_r >> (lane + 1))) - 1;
                    if (lane > 0) left = __clz((int)~(chain_l << (32 - lane)));
                    seg = (1 + left + right) >= 3;
                }
                const unsigned segm = ballot(seg);
                if (seg) {
                    const int slot = n + __popc(segm & lt);
                    if (slot < CF::ML) {
                        s.line_mask[slot] = ((2u << (left + right)) - 1u) << (lane - left);   // columns lane-left..lane+right
                        s.line_kind[slot] = 0;
                        s.line_idx[slot] = (uint8_t)r;
                        s.line_colour[slot] = (uint8_t)col[r * C + lane];
                        // same top row: phase 1 first, then phase 2 by (column of the vertical line, row)
                        s.line_key[slot] = ((uint32_t)r << 12) | (uint32_t)(1024 + lane * 32 + r);
                    }
                }
                n += __popc(segm);
            }
        }
        if (n > CF::ML) { status |= ST_LINE_OVERFLOW; n = CF::ML; }
        sync();
        return n;
    }

    // =====================================================================================================
    // Order-dependent part of a cascade round.  Classification of the lines (pure coordinate work on small
    // tables) runs on the leader lane; everything that touches the board -- deleting match cells, the special
    // activation DFS, combination matches -- runs on ALL lanes with uniform control flow, so that every sweep
    // (laser column/row, bomb window, cookie scans, colour histogram) is one parallel step.  The reference's
    // visit order is kept exactly: a sweep deletes the normal tiles up to the first special in its order in
    // one go (no activation can happen in between), then enters that special, then continues after it.
    // =====================================================================================================
    __device__ __forceinline__ uint16_t* mlist() { return reinterpret_cast<uint16_t*>(s.mask); }  // mask staging is free now
    static constexpr int MCAP = CF::MAXP;

    // number of cells whose colour is not 0 (ref :488 np.all(board[0] == 0), :532-533)
    __device__ __forceinline__ int count_nonzero_colours() {
        int n = 0;
        if (lane < C) {
#pragma unroll 1
            for (int r = 0; r < (RT ? RT : R); ++r) n += (col[r * C + lane] != 0);
        }
        return radd(n);
    }
    // frame: kind(2) | cell(10) << 2 | cursor(11) << 12 | colour(5) << 23
    __device__ __forceinline__ static uint32_t frame(int kind, int cell, int cursor, int mc) {
        return (uint32_t)kind | ((uint32_t)cell << 2) | ((uint32_t)cursor << 12) | ((uint32_t)mc << 23);
    }
    // entry of activate_special (ref :473-499 + the set-up of the cookie branch :530-544); pushes a frame.
    // All lanes, uniform arguments.  Board writes of other lanes must be visible (callers sync before).
    __device__ void enter_activation(int cell, int t, bool counted, int& sp) {
        const int own_nz = col[cell] != 0;
        // ref :488-489 "all colours zero": impossible when the target itself is coloured, so only colourless targets count
        int nz = own_nz ? 1 : count_nonzero_colours();
        if (nz == 0) return;
        if (t == 0 || t == 1) { status |= ST_INTERNAL; return; }  // ref :491-492 raises
        sync();                                                // every lane has read the cell
        if (lane == 0) { col[cell] = 0; typ[cell] = 0; }       // ref :496
        if (counted) ++n_act;                                  // ref :498-499
        int kind, mc = 0;
        if (t == 2) kind = 0;
        else if (t == 3) kind = 1;
        else if (t == 4) kind = 2;
        else if (t == -1) {
            kind = 3;
            if (own_nz) nz = count_nonzero_colours();          // (a coloured cookie only exists after a shuffle + redraw)
            nz -= own_nz;
            if (nz == 0) { sync(); return; }                   // ref :532-534
            sync();
            int best = 0, seen = 0;                            // ref :536-537: most common colour, lowest on ties
#pragma unroll 1
            for (int k = 1; k < 32; ++k) {
                if (k > K && seen >= nz) break;                // colours above K only exist on hand-made boards
                int c = 0;
                if (lane < C) {
#pragma unroll 1
                    for (int r = 0; r < R; ++r) c += (col[r * C + lane] == k);
                }
                c = radd(c);
                seen += c;
                if (c > best) { best = c; mc = k; }
            }
            if (lane < C) {                                    // ref :540-544: delete its normal tiles
#pragma unroll 1
                for (int r = 0; r < (RT ? RT : R); ++r) {
                    const int i = r * C + lane;
                    if (col[i] == mc && typ[i] == 1) { col[i] = 0; typ[i] = 0; }
                }
            }
        } else { status |= ST_INTERNAL; sync(); return; }      // ref :555-556 raises
        if (sp >= CF::DFS) { status |= ST_DFS_OVERFLOW; sync(); return; }
        if (lane == 0) s.stack[sp] = frame(kind, cell, 0, mc);
        ++sp;
        sync();
    }
    // first index i in [cur, n) (lanes test i = base + lane) for which pred(i) holds, else n
    template <typename F> __device__ __forceinline__ int first_in_order(int cur, int n, F&& pred) {
#pragma unroll 1
        for (int base = cur; base < n; base += L) {
            const int i = base + lane;
            const unsigned m = ballot(i < n && pred(i));
            if (m) return base + __ffs((int)m) - 1;
        }
        return n;
    }
    template <typename F> __device__ __forceinline__ void delete_in_order(int cur, int end, F&& cell_of) {
#pragma unroll 1
        for (int i = cur + lane; i < end; i += L) {
            const int q = cell_of(i);
            col[q] = 0;
            typ[q] = 0;
        }
    }
    // activate_special (ref :473-556) as an explicit-stack DFS; all lanes, uniform control flow
    __device__ void activate(int cell0, int t0, bool counted) {
        int sp = 0;
        int e_cell = cell0, e_t = t0;
        bool e_counted = counted;
#pragma unroll 1
        for (;;) {
            enter_activation(e_cell, e_t, e_counted, sp);     // one call site: nested calls always count (ref :505,513,526,554)
            e_counted = true;
            e_cell = -1;
#pragma unroll 1
          while (sp > 0) {
            sync();
            const uint32_t f = s.stack[sp - 1];
            const int kind = (int)(f & 3u), cell = (int)((f >> 2) & 1023u), mc = (int)(f >> 23);
            const int cur = (int)((f >> 12) & 2047u);
            const int r0 = cell / C, c0 = cell - r0 * C;
            int target = -1, next = 0;
            if (kind == 0) {                                   // vertical laser: rows top to bottom (ref :502-507)
                auto cell_of = [&](int i) { return i * C + c0; };
                const int first = first_in_order(cur, R, [&](int i) { return not01(typ[cell_of(i)]); });
                delete_in_order(cur, first, cell_of);
                if (first < R) { target = cell_of(first); next = first + 1; }
            } else if (kind == 1) {                            // horizontal laser: columns left to right (ref :510-515)
                auto cell_of = [&](int i) { return r0 * C + i; };
                const int first = first_in_order(cur, C, [&](int i) { return not01(typ[cell_of(i)]); });
                delete_in_order(cur, first, cell_of);
                if (first < C) { target = cell_of(first); next = first + 1; }
            } else if (kind == 2) {                            // bomb: clipped 3x3, row-major (ref :517-528)
                const int min_r = max(r0 - 1, 0), max_r = min(r0 + 1, R - 1);
                const int min_c = max(c0 - 1, 0), max_c = min(c0 + 1, C - 1);
                const int w = max_c - min_c + 1, n = w * (max_r - min_r + 1);
                auto cell_of = [&](int i) { return (min_r + i / w) * C + min_c + i % w; };
                const int first = first_in_order(cur, n, [&](int i) { return not01(typ[cell_of(i)]); });
                delete_in_order(cur, first, cell_of);
                if (first < n) { target = cell_of(first); next = first + 1; }
            } else {                                           // cookie: specials of its colour, row-major (ref :547-554)
                int cand = 1 << 20;
                if (lane < C)
#pragma unroll 1
                    for (int r = cur / C; r < R; ++r) {
                        const int i = r * C + lane;
                        if (i >= cur && col[i] == mc && typ[i] > 1) { cand = i; break; }
                    }
                cand = rmin(cand);
                if (cand < (1 << 20)) { target = cand; next = cand + 1; }
            }
            if (target < 0) { --sp; continue; }
            e_t = typ[target];
            e_cell = target;
            sync();                                            // deletions of this sweep are visible; frame can be updated
            if (lane == 0) s.stack[sp - 1] = frame(kind, cell, next, mc);
            break;
          }
            if (e_cell < 0) break;
        }
        sync();
    }

    // get_special_creation_pos for a bomb (ref :441-450); leader lane.  The match is a straight line plus up to three
    // cells of the crossing line, listed as out[0..n): the modal row / column (first element with the top count,
    // ref :445) follow from the geometry -- all cells of a horizontal line share its row, and the crossing column is
    // the only column that can repeat (shared cell + extra cells) -- so no counting is needed.
    __device__ __forceinline__ int creation_pos_bomb(const uint16_t* cells, int n, int ntaken, int kind, int idx, int k2, int i2,
                                                     int first_bit, int nextra) {
        int mrow, mcol;
        if (kind == 0) { mrow = idx; mcol = (k2 != kind && nextra >= 1) ? i2 : first_bit; }
        else { mcol = idx; mrow = (k2 != kind && nextra >= 1) ? i2 : first_bit; }
        const int corner = mrow * C + mcol;
        auto is_taken = [&](int cell) {
#pragma unroll 1
            for (int q = 0; q < ntaken; ++q) if (s.taken[q] == cell) return true;
            return false;
        };
        if (!is_taken(corner)) return corner;               // ref :446-447 (the corner is always a cell of the match)
        int best = -1, bestd = 0;
#pragma unroll 1
        for (int k = 0; k < n; ++k) {
            const int cell = cells[k];
            if (is_taken(cell)) continue;
            const int dr = cell / C - mrow, dc = cell % C - mcol, d = dr * dr + dc * dc;
            if (best < 0 || d < bestd) { best = cell; bestd = d; }  // stable: first minimum (ref :449)
        }
        return best;                                        // -1: no valid cell (reference: IndexError)
    }

    // process_colour_lines (ref :269-327) + the creation cells of resolve_colour_matches (ref :414-418); leader lane.
    // Lines are straight (also after a bomb took cells out of them), so a line is (kind, row|column, bit set) and
    // "do these two lines share a cell" is two bit tests.  Writes the cells of all matches, in resolve order, to
    // mlist() and the creation queue to s.cq_*.  Returns nm | ncq << 16 | flags << 24 (flag 1: table overflow,
    // flag 2: no valid creation cell).
    __device__ __forceinline__ int line_cell(int kind, int idx, int bit) const { return kind ? bit * C + idx : idx * C + bit; }
    __device__ __forceinline__ static unsigned lowest_bits(unsigned m, int k) {   // the k lowest set bits of m
        unsigned out = 0u;
#pragma unroll 1
        for (int i = 0; i < k && m; ++i) { const unsigned b = m & (0u - m); out |= b; m ^= b; }
        return out;
    }
    __device__ __forceinline__ static int nth_bit(unsigned m, int k) {            // position of the k-th (0-based) set bit
#pragma unroll 1
        for (int i = 0; i < k; ++i) m &= m - 1u;
        return __ffs((int)m) - 1;
    }
    __device__ __forceinline__ uint32_t classify_lines(int n) {
        // ref :282: stable sort by the first cell's row == sort by key (keys are unique)
#pragma unroll 1
        for (int i = 0; i < n; ++i) s.order[i] = (uint8_t)i;
#pragma unroll 1
        for (int i = 1; i < n; ++i) {
            const uint8_t v = s.order[i];
            const uint32_t kv = s.line_key[v];
            int j = i - 1;
#pragma unroll 1
            while (j >= 0 && s.line_key[s.order[j]] > kv) { s.order[j + 1] = s.order[j]; --j; }
            s.order[j + 1] = v;
        }
        uint16_t* out = mlist();
        int qh = 0, qn = n, nslots = n, ncq = 0, ntaken = 0, nm = 0;
        uint32_t flags = 0u;
        const bool sp_cookie = specials & SP_COOKIE, sp_v = specials & SP_VLASER, sp_h = specials & SP_HLASER,
                   sp_bomb = specials & SP_BOMB;
#pragma unroll 1
        while (qh < qn) {
            const int li = s.order[qh++];                    // ref :285 pop(0)
            const int kind = s.line_kind[li], idx = s.line_idx[li];
            const unsigned mask = s.line_mask[li];
            const int len = __popc(mask);
            int name = NAME_NORMAL, colour = s.line_colour[li];
            unsigned mm = 0u;                                // cells of the match that lie on this line
            int extra[3] = {-1, -1, -1}, nextra = 0;         // bomb: cells taken from the crossing line
            int bomb_k2 = 0, bomb_i2 = 0;
            if (len >= 5 && sp_cookie) {                     // ref :287-292
                mm = lowest_bits(mask, 5);
                name = NAME_COOKIE; colour = 0;
                const unsigned rest = mask & ~mm;
                if (__popc(rest) > 2) {
                    if (nslots < CF::ML && qn < CF::ML) {
                        const int ns = nslots++;
                        s.line_mask[ns] = rest; s.line_kind[ns] = (uint8_t)kind; s.line_idx[ns] = (uint8_t)idx;
                        s.line_colour[ns] = s.line_colour[li];
                        s.order[qn++] = (uint8_t)ns;
                    } else flags |= 1u;
                }
            } else if (len == 4) {                           // ref :294-302
                mm = mask;
                name = (kind == 0 && sp_h) ? NAME_HLASER : (sp_v ? NAME_VLASER : NAME_NORMAL);
            } else {
                int hit = -1, spos = 0;                      // spos: position of the shared cell along the crossing line
                if (sp_bomb) {                               // ref :304-308: first queued line sharing a cell
#pragma unroll 1
                    for (int q = qh; q < qn; ++q) {
                        const int lj = s.order[q];
                        const int k2 = s.line_kind[lj], i2 = s.line_idx[lj];
                        const unsigned m2 = s.line_mask[lj];
                        if (k2 != kind) {                    // a row and a column meet in one cell
                            if (((mask >> i2) & 1u) && ((m2 >> idx) & 1u)) { hit = q; spos = idx; break; }
                        } else if (i2 == idx && (mask & m2)) { // same row (crossing segments can overlap)
                            hit = q; spos = __ffs((int)(mask & m2)) - 1; break;   // first cell of `line` that is in l
                        }
                    }
                }
                if (hit >= 0) {                              // ref :309-320
                    const int lj = s.order[hit];
                    const int k2 = s.line_kind[lj], i2 = s.line_idx[lj];
                    unsigned m2 = s.line_mask[lj];
                    const int llen = __popc(m2);
                    mm = mask;
                    bomb_k2 = k2; bomb_i2 = i2;
                    // ref :310-312: the three cells of l closest to the shared cell (stable: lower position first on ties)
                    unsigned picked = 0u;
#pragma unroll 1
                    for (int d = 0; d < 32 && __popc(picked) < 3 && picked != m2; ++d) {
#pragma unroll 1
                        for (int sgn = 0; sgn < (d ? 2 : 1); ++sgn) {
                            const int pos = sgn ? spos + d : spos - d;
                            if (pos < 0 || pos > 31 || !((m2 >> pos) & 1u) || __popc(picked) >= 3) continue;
                            picked |= 1u << pos;
                            const bool in_line = (k2 != kind) ? (pos == idx) : (((mask >> pos) & 1u) != 0u);
                            if (!in_line) extra[nextra++] = line_cell(k2, i2, pos);
                        }
                    }
                    name = NAME_BOMB;
                    if (llen < 6) {                          // ref :315-316 (lines are distinct by value)
#pragma unroll 1
                        for (int q = hit; q + 1 < qn; ++q) s.order[q] = s.order[q + 1];
                        --qn;
                    } else {                                 // ref :317-319: drop the three cells
                        s.line_mask[lj] = m2 & ~picked;
                    }
                } else if (len >= 3) {                       // ref :322-325
                    mm = mask;
                } else continue;
            }
            // cells of the match in list order: the line's cells ascending, then the bomb's extra cells (ref :312)
            const int first_out = nm;
            {
                unsigned t = mm;
#pragma unroll 1
                while (t) {
                    const int bit = __ffs((int)t) - 1;
                    t &= t - 1u;
                    if (nm < MCAP) out[nm++] = (uint16_t)line_cell(kind, idx, bit);
                    else flags |= 1u;
                }
#pragma unroll 1
                for (int k = 0; k < nextra; ++k) {
                    if (nm < MCAP) out[nm++] = (uint16_t)extra[k];
                    else flags |= 1u;
                }
            }
            if (name != NAME_NORMAL) {                       // ref :414-418
                int pos;
                if (name == NAME_BOMB) {
                    pos = creation_pos_bomb(out + first_out, nm - first_out, ntaken, kind, idx, bomb_k2, bomb_i2,
                                            __ffs((int)mm) - 1, nextra);
                } else {                                     // straight: middle of the valid cells (ref :453-458)
                    unsigned valid = mm;
#pragma unroll 1
                    for (int q = 0; q < ntaken; ++q) {
                        const int tc = s.taken[q], tr = tc / C, tcol = tc - tr * C;
                        if (kind == 0 ? tr == idx : tcol == idx) valid &= ~(1u << (kind == 0 ? tcol : tr));
                    }
                    const int nv = __popc(valid);
                    pos = nv ? line_cell(kind, idx, nth_bit(valid, (nv % 2 == 0) ? nv / 2 - 1 : nv / 2)) : -1;
                }
                if (pos < 0) flags |= 2u;
                if (pos >= 0 && ntaken < CF::ML) s.taken[ntaken++] = (uint16_t)pos;
                if (ncq < CF::ML) {
                    s.cq_pos[ncq] = (uint16_t)(pos < 0 ? 0xffff : pos);
                    s.cq_type[ncq] = (int8_t)name;
                    s.cq_colour[ncq] = (uint8_t)colour;
                    ++ncq;
                }
            }
        }
        return (uint32_t)nm | ((uint32_t)ncq << 16) | (flags << 24);
    }

    // resolve_colour_matches (ref :397-427) on the classified matches; all lanes
    __device__ __forceinline__ void resolve_matches(uint32_t packed) {
        const int nm = (int)(packed & 0xffffu), ncq = (int)((packed >> 16) & 0xffu);
        const uint32_t flags = packed >> 24;
        if (flags & 1u) status |= ST_LINE_OVERFLOW;
        if (flags & 2u) status |= ST_INTERNAL;
        const uint16_t* ml = mlist();
        int pos = 0;
#pragma unroll 1
        while (pos < nm) {                                   // ref :421-423 -> resolve_colour_match :460-471
            auto cell_of = [&](int i) { return (int)ml[i]; };
            const int first = first_in_order(pos, nm, [&](int i) { return not01(typ[cell_of(i)]); });
            delete_in_order(pos, first, cell_of);
            if (first >= nm) break;
            const int cell = cell_of(first);
            const int t = typ[cell];
            sync();
            activate(cell, t, true);
            pos = first + 1;
        }
        sync();
#pragma unroll 1
        for (int i = lane; i < ncq; i += L) {                // ref :426-427 -> create_special :572-597
            if (s.cq_pos[i] == 0xffff) continue;
            col[s.cq_pos[i]] = (int8_t)s.cq_colour[i];
            typ[s.cq_pos[i]] = s.cq_type[i];
        }
        n_new += ncq;
    }

    // Fast path of a cascade round (most rounds): the lines anchored on row r* -- horizontal, vertical or both -- are all
    // 3 or 4 long, pairwise disjoint, contain no special tile and have no crossing segments.  Then process_colour_lines yields one normal / laser
    // match per line (ref :294-302,322-325), nothing is activated, deletions commute, and the creation cell of a
    // 4-line is its second cell (ref :453-456).  Returns the number of lines, or 0 if the general path must run.
    __device__ __forceinline__ int fast_round(const Scan& sc) {
        const Bits& b = sc.bits;
        const int rs = sc.rstar;
        const bool sp_v = specials & SP_VLASER, sp_h = specials & SP_HLASER;
        // horizontal lines of row rs
        const bool mine = (sc.hcells >> lane) & 1u;
        const bool start = (sc.hs >> lane) & 1u;
        const int hlen = start ? __ffs((int)~(sc.m >> lane)) : 0;             // run + 1
        bool bad = (mine && ((b.S >> rs) & 1u)) || hlen > 4;
        // vertical lines anchored on row rs.  A phase-2 segment (ref :198-214) is a horizontal run of >= 3 equal cells
        // through a cell of such a line: the cell with the two to its right, with one on either side, or with the two
        // to its left (ignoring the type and phase-1 cuts of :199-209 only makes the test stricter).  It also covers a
        // horizontal line of row rs that shares its cell with the vertical line, so the lines that pass are disjoint.
        const int vlen = sc.has_v ? rs - sc.vtop + 1 : 0;
        if (sc.mv) {
            const unsigned vrows = sc.has_v ? ((2u << rs) - 1u) & ~((1u << sc.vtop) - 1u) : 0u;
            const unsigned El = from_left(b.E, 1), Er = from_right(b.E, 1), Ell = from_left(El, 1);
            const unsigned cross = (b.E & Er) | (El & b.E) | (Ell & El);
            bad = bad || (vrows & (b.S | cross)) != 0u || vlen > 4;
        }
        if (ballot(bad)) return 0;
        // every line is 3 or 4 normal tiles and no two share a cell: one match per line, in any order
        const int laser = sp_h ? 3 : (sp_v ? 2 : 0);                          // horizontal 4-line (ref :297-302)
        const unsigned create = (laser && sc.hs) ? (ballot(start && hlen == 4) << 1) : 0u;   // its second cell (ref :453-456)
        const bool make = sc.has_v && vlen == 4 && sp_v;                      // vertical 4-line -> vertical laser or normal
        if (DEFER_GAPS) { fg_len = 0; fg_valid = true; }
        if (mine) {
            const int i = rs * C + lane;
            if ((create >> lane) & 1u) typ[i] = (int8_t)laser;               // keeps the line's colour (ref :596-597)
            else if (DEFER_GAPS) { fg_top = rs; fg_len = 1; }
            else { col[i] = 0; typ[i] = 0; }
        }
        if (sc.has_v) {
            if (DEFER_GAPS) {
                // the laser is created on the second cell (ref :453-456) and falls to the anchor row: write it there
                fg_top = sc.vtop;
                fg_len = make ? vlen - 1 : vlen;
                if (make) typ[rs * C + lane] = 2;
            } else {
                for (int r = sc.vtop; r <= rs; ++r) {
                    const int i = r * C + lane;
                    if (make && r == sc.vtop + 1) typ[i] = 2;
                    else { col[i] = 0; typ[i] = 0; }
                }
            }
        }
        n_new += __popc(create) + (sc.mv ? __popc(ballot(make)) : 0);
        return __popc(sc.hs) + __popc(sc.mv);
    }

    // general path of a cascade round: line table, classification, resolution with activations
    __device__ __forceinline__ int general_round(const Scan& sc) {
        const long long t0 = prof_on ? clock64() : 0;
        if (sc.rstar < 0) return 0;
        const long long t1 = prof_on ? clock64() : 0;
        const int n = build_line_table(sc);
        const long long t2 = prof_on ? clock64() : 0;
        uint32_t packed = 0u;
        if (lane == 0) packed = classify_lines(n);
        packed = (uint32_t)shfl((int)packed, 0);
        sync();
        const long long t3 = prof_on ? clock64() : 0;
        resolve_matches(packed);
        sync();
        if (prof_on && lane == 0) {   // diagnostics: cycles per phase of the general path, accumulated per env
            const long long t4 = clock64();
            atomicAdd(&p.prof[env * 8 + 4], (uint32_t)(t1 - t0));
            atomicAdd(&p.prof[env * 8 + 5], (uint32_t)(t2 - t1));
            atomicAdd(&p.prof[env * 8 + 6], (uint32_t)(t3 - t2));
            atomicAdd(&p.prof[env * 8 + 7], (uint32_t)(t4 - t3));
        }
        return n;
    }

    // one cascade round without gravity/refill (ref :369-373); returns the number of lines found
    __device__ __forceinline__ int resolve_round() {
        sync();
        const Scan sc = scan_lines(R - 1, false);
        if (sc.rstar < 0) return 0;
        ++prof_rounds;
        int n = fast_round(sc);
        if (n == 0) {
            const long long t0 = prof_on ? clock64() : 0;
            SlowScan ss;
            ss.rstar = sc.rstar; ss.mv = sc.mv; ss.hs = sc.hs; ss.hcells = sc.hcells; ss.m = sc.m; ss.vtop = sc.vtop; ss.has_v = sc.has_v;
            ss.E = sc.bits.E; ss.D = sc.bits.D; ss.T = sc.bits.T; ss.S = sc.bits.S;
            const SlowOut o = slow_round<L, RT, CT>(&s, &p, lane, gmask, gshift, env, n_new, n_act, ss);
            n = o.n; n_new = o.n_new; n_act = o.n_act; status |= o.status;
            if (prof_on) prof_serial += (uint32_t)(clock64() - t0);
        }
        sync();
        return n;
    }

    // combination_match (ref :600-719); all lanes, uniform control flow.  Every branch reduces to a sequence of
    // top-level activate_special calls (is_combination_match=True, i.e. not counted), produced by one driver loop so
    // that the activation machinery is instantiated once.
    __device__ __forceinline__ void combination(int i1, int i2) {
        n_act += 2;                                          // ref :609
        sync();
        const int t1 = typ[i1], k1 = col[i1], t2 = typ[i2], k2 = col[i2];
        const int r1 = i1 / C, c1 = i1 % C, r2 = i2 / C, c2 = i2 % C;
        const int r = min(r1, r2), c = min(c1, c2);
        sync();
        enum { NONE, LIST2, CROSS, ROWMAJOR, WINDOW };
        int mode = NONE, kk = 0;
        bool strict = false;                                 // ROWMAJOR: type > 1 (cookie+normal) vs type not in {0,1}
        int min_r = 0, max_r = 0, min_c = 0, max_c = 0;
        if (t1 == -1 && t2 == -1) {                          // ref :615-616
#pragma unroll 1
            for (int i = lane; i < P; i += L) { col[i] = 0; typ[i] = 0; }
        } else if ((t1 == -1 && t2 == 1) || (t1 == 1 && t2 == -1)) {  // ref :619-641
            const int ck = (t1 == -1) ? i1 : i2;
            kk = (t1 == -1) ? k2 : k1;
            if (lane == 0) { col[ck] = 0; typ[ck] = 0; }     // ref :626,628
            sync();
#pragma unroll 1
            for (int i = lane; i < P; i += L) if (col[i] == kk && typ[i] == 1) { col[i] = 0; typ[i] = 0; }  // ref :631-635
            // snapshot mask colour==kk & type>1, visited row-major; cells can only disappear meanwhile (ref :638-640)
            mode = ROWMAJOR; strict = true;
            n_act -= 1;                                      // ref :641
        } else if ((t1 == -1 && t2 >= 2) || (t1 >= 2 && t2 == -1)) {  // ref :644-660
            const int ck = (t1 == -1) ? i1 : i2;
            kk = (t1 == -1) ? k2 : k1;
            const int tt = (t1 == -1) ? t2 : t1;
            if (lane == 0) { col[ck] = 0; typ[ck] = 0; }     // ref :651
            sync();
            // ref :654 snapshot of colour==kk; a cell leaves the snapshot only by deletion (colour -> 0), and a cell
            // that was not in it never gains the colour, so a live test of colour==kk is the same set.
#pragma unroll 1
            for (int i = lane; i < P; i += L) if (col[i] == kk && typ[i] == 1) typ[i] = (int8_t)tt;  // ref :655-657
            mode = ROWMAJOR; strict = false;                 // ref :660
        } else if ((t1 == 2 || t1 == 3) && (t2 == 2 || t2 == 3)) {    // ref :663-674: v-laser then h-laser at (r,c)
            if (lane == 0) { col[i1] = 0; typ[i1] = 0; col[i2] = 0; typ[i2] = 0; }
            mode = LIST2;
        } else if ((t1 == 4 && (t2 == 2 || t2 == 3)) || (t2 == 4 && (t1 == 2 || t1 == 3))) {  // ref :677-696
            if (lane == 0) { col[i1] = 0; typ[i1] = 0; col[i2] = 0; typ[i2] = 0; }
            min_r = max(r - 1, 0); max_r = min(r + 1, R - 1);
            min_c = max(c - 1, 0); max_c = min(c + 1, C - 1);
            mode = CROSS;                                    // h-lasers on rows r-1..r+1, then v-lasers on columns c-1..c+1
        } else if (t1 == 4 && t2 == 4) {                     // ref :699-719
            if (lane == 0) { col[i1] = 0; typ[i1] = 0; col[i2] = 0; typ[i2] = 0; }
            min_r = max(r - 2, 0); max_r = min(r + 2, R - 1);
            min_c = max(c - 2, 0); max_c = min(c + 2, C - 1);
            mode = WINDOW;
        }
        const int nr = max_r - min_r + 1, w = max_c - min_c + 1;
        int cur = 0;
#pragma unroll 1
        while (mode != NONE) {
            sync();
            int cell = -1, t = 0;
            if (mode == LIST2) {
                if (cur < 2) { cell = r * C + c; t = 2 + cur; ++cur; }
            } else if (mode == CROSS) {
                if (cur < nr) { cell = (min_r + cur) * C + c; t = 3; ++cur; }
                else if (cur < nr + w) { cell = r * C + min_c + (cur - nr); t = 2; ++cur; }
            } else if (mode == ROWMAJOR) {                   // activate_specials_in_mask (ref :721-726)
                int cand = 1 << 20;
                if (lane < C) {
#pragma unroll 1
                    for (int rr = cur / C; rr < R; ++rr) {
                        const int i = rr * C + lane;
                        if (i >= cur && col[i] == kk && (strict ? typ[i] > 1 : not01(typ[i]))) { cand = i; break; }
                    }
                }
                cand = rmin(cand);
                if (cand < (1 << 20)) { cell = cand; t = typ[cand]; cur = cand + 1; }
            } else {                                         // 5x5 window, row-major: normal -> delete, other non-empty -> activate
                const int n = w * nr;
                auto cell_of = [&](int i) { return (min_r + i / w) * C + min_c + i % w; };
                const int first = first_in_order(cur, n, [&](int i) { const int tt = typ[cell_of(i)]; return tt != 1 && tt != 0; });
#pragma unroll 1
                for (int i = cur + lane; i < first; i += L) {
                    const int q = cell_of(i);
                    if (typ[q] == 1) { col[q] = 0; typ[q] = 0; }
                }
                if (first < n) { cell = cell_of(first); t = typ[cell]; cur = first + 1; }
            }
            if (cell < 0) break;
            sync();
            activate(cell, t, false);
        }
        sync();
    }

    __device__ void shuffle() {
        sync();
        if (lane == 0) shuffle_serial(col, typ, P, gid, p.key0, p.key1, in_reset ? rsc : scur, in_reset ? (long long)episode : -1ll);
        if (in_reset) rsc += (uint64_t)(P > 1 ? P - 1 : 0);  // one word per Fisher-Yates step, on every lane
        else scur += (uint64_t)(P > 1 ? P - 1 : 0);
        sync();
    }

    __device__ __forceinline__ void action_cells(int a, int& i1, int& i2) const {  // ref :80-91
        const int nv = C * (R - 1);
        if (a < nv) { i1 = a; i2 = a + C; }
        else {
            const int j = a - nv;
            const int r = j / (C - 1), c = j - r * (C - 1);
            i1 = r * C + c; i2 = i1 + 1;
        }
    }
    // is_move_effective for one action, all lanes cooperate: lane j tests window row j (horizontal triples) and
    // window column j (vertical triples) with a virtual swap (ref :735-787)
    __device__ bool effective_group(int a) {
        int i1, i2;
        action_cells(a, i1, i2);
        const int ta = typ[i1], tb = typ[i2];
        if (not01(ta) && not01(tb)) return true;
        if (ta < 0 || tb < 0) return true;
        const int r1 = i1 / C, c1 = i1 % C, r2 = i2 / C, c2 = i2 % C;
        const int rmin_ = max(0, min(r1, r2) - 2), rmax_ = min(R - 1, max(r1, r2) + 2);
        const int cmin_ = max(0, min(c1, c2) - 2), cmax_ = min(C - 1, max(c1, c2) + 2);
        auto cs = [&](int i) -> int { return i == i1 ? col[i2] : (i == i2 ? col[i1] : col[i]); };
        auto ts = [&](int i) -> int { return i == i1 ? typ[i2] : (i == i2 ? typ[i1] : typ[i]); };
        bool f = false;
        const int r = rmin_ + lane;
        if (r <= rmax_)
            for (int c = cmin_; c + 2 <= cmax_; ++c) {
                const int i = r * C + c;
                f |= (cs(i) == cs(i + 1) && cs(i + 1) == cs(i + 2) && ts(i + 2) >= 0);
            }
        const int c = cmin_ + lane;
        if (c <= cmax_)
            for (int rr = rmin_; rr + 2 <= rmax_; ++rr) {
                const int i = rr * C + c;
                f |= (cs(i) == cs(i + C) && cs(i + C) == cs(i + 2 * C) && ts(i + 2 * C) >= 0);
            }
        return ballot(f) != 0u;
    }

    // ===================================================================================================
    // legal-move mask (ref tile_match_env.py:118-124) by row bitboards: lane c holds, per colour k, the set of
    // rows of column c that carry k.  effv bit r <-> action r*C+c (swap with the cell below);
    // effh bit r <-> action C(R-1) + r(C-1) + c (swap with the cell to the right).
    // Boards with a pre-existing triple, or with a cell whose (colour==0) disagrees with (type<0), take the
    // literal per-action path (their masks depend on the exact window rule).
    // ===================================================================================================
    __device__ __forceinline__ bool mask_bits(unsigned& effv_out, unsigned& effh_out) {
        sync();
        const bool in = lane < C;
        unsigned S = 0u, Ng = 0u;    // rows with type not in {0,1} / type < 0
        bool odd = false;
        if (in)
#pragma unroll UR
            for (int r = 0; r < R; ++r) {
                const int t = typ[r * C + lane], x = col[r * C + lane];
                S |= (unsigned)not01(t) << r;
                Ng |= (unsigned)(t < 0) << r;
                odd |= (t < 0) != (x == 0);
            }
        last_S = S;
        const unsigned rows = rows_mask();
        const unsigned rows_v = rows >> 1;  // rows r with r+1 < R
        unsigned effv = ((S & (S >> 1)) | Ng | (Ng >> 1));                       // ref :750,754
        unsigned effh = (S & from_right(S, 1)) | Ng | from_right(Ng, 1);
        unsigned unstable = 0u;
#pragma unroll 1
        for (int k = 1; k <= K; ++k) {
            unsigned b = 0u;
            if (in)
#pragma unroll UR
                for (int r = 0; r < R; ++r) b |= (unsigned)(col[r * C + lane] == k) << r;
            const unsigned l1 = from_left(b, 1), l2 = from_left(b, 2), r1 = from_right(b, 1), r2 = from_right(b, 2);
            const unsigned hl = l1 & l2, hm = l1 & r1, hr = r1 & r2;      // a k-tile placed here completes a row triple
            const unsigned vu = (b << 1) & (b << 2), vm = (b << 1) & (b >> 1), vd = (b >> 1) & (b >> 2);
            const unsigned hany = hl | hm | hr, vany = vu | vm | vd;
            unstable |= (b & (b >> 1) & (b >> 2)) | (b & r1 & r2);
            // horizontal swap (r,c)<->(r,c+1): the k-tile on the right moves here, or the k-tile here moves right
            effh |= (r1 & (hl | vany)) | (b & from_right(hr | vany, 1));
            // vertical swap (r,c)<->(r+1,c): the k-tile below moves up, or the k-tile here moves down
            effv |= ((b >> 1) & (vu | hany)) | (b & ((vd | hany) >> 1));
        }
        if (!in || lane + 1 >= C) effh = 0u;
        if (!in) effv = 0u;
        effv &= rows_v;
        effh &= rows;
        if (ballot(unstable != 0u || odd)) {
            const MaskPair lit = mask_literal_column(col, typ, R, C, lane);
            effv = lit.v; effh = lit.h;
        }
        effv_out = effv; effh_out = effh;
        return ballot((effv | effh) != 0u) != 0u;
    }
    __device__ void mask_to_smem(unsigned effv, unsigned effh) {
        if (lane < C) {
            const int nv = C * (R - 1);
#pragma unroll UR
            for (int r = 0; r < R; ++r) {
                if (r + 1 < R) s.mask[r * C + lane] = (uint8_t)((effv >> r) & 1u);
                if (lane + 1 < C) s.mask[nv + r * (C - 1) + lane] = (uint8_t)((effh >> r) & 1u);
            }
        }
    }

    // ===================================================================================================
    // playability loop shared by generate_board (ref :99-109) and move (ref :381-391)
    // ===================================================================================================
    // `clean`: the caller knows the board has no lines.  `all_normal`: every tile has type 1 (fresh board).
    // Leaves the mask bits of the final board in effv/effh.
    __device__ __forceinline__ bool playability(bool clean, bool all_normal, unsigned& effv, unsigned& effh, int iters0 = 0) {
        bool shuffled = false;
        int iters = iters0, from = R - 1;
#pragma unroll 1
        for (;;) {
            bool capped = false;
            if (!clean) {
                sync();
                const Scan sc = scan_lines(from, all_normal);
                if (sc.rstar >= 0) {                         // remove_colour_lines (ref :120-131)
                    if (iters < p.max_iters) {
                        ++iters;
                        ++prof_iters;
                        const int top = first_line_top(sc);
                        const int row = min(R - 1, top + 1);
                        draw_cells((row + 1) * C, false);
                        // rows below max(rstar, row+2) were line-free before and none of their 3-windows changed
                        from = min(R - 1, max(sc.rstar, row + 2));
                        continue;
                    }
                    status |= ST_RESET_CAP;
                    capped = true;
                }
            }
            const bool any = mask_bits(effv, effh);          // possible_move (ref :558-569)
            if (any || capped) return shuffled;
            if (iters >= p.max_iters) { status |= ST_RESET_CAP; return shuffled; }
            ++iters;
            shuffled = true;
            shuffle();
            clean = false;
            all_normal = false;
            from = R - 1;
        }
    }

    // One iteration of remove_colour_lines on a fresh (all-normal) board (ref :120-131), the unit of work of k_pregen's
    // warp-converged loop.  Returns true if a line was found and the rows above it were redrawn, false when the board
    // is line-free (or the iteration cap was hit: `capped`).  Same arithmetic as the loop in playability().
    __device__ __forceinline__ bool redraw_iteration(int& from, int& iters, bool& capped) {
        sync();
        const Scan sc = scan_lines(from, true);
        if (sc.rstar < 0) return false;
        if (iters >= p.max_iters) { status |= ST_RESET_CAP; capped = true; return false; }
        ++iters;
        ++prof_iters;
        const int top = first_line_top(sc);
        const int row = min(R - 1, top + 1);
        draw_cells((row + 1) * C, false);
        from = min(R - 1, max(sc.rstar, row + 2));
        return true;
    }

    // generate_board's line removal (ref :95-101, :120-131) on PACKED ROWS, one board per warp: lane r holds row r of the
    // fresh board as one word of BITS-bit cells (a fresh board is colours only).  The line scan is then a handful of
    // xor / shift / and on the lane's own word plus two shuffles for the rows above, instead of a pass over the byte
    // planes, and a redraw assembles each row from the Philox words in shared memory.  Same draws, same order, same
    // result as begin_generate + the redraw loop of playability(); leaves the line-free board as bytes in shared
    // memory with the reset-stream cursor and the iteration count where that loop would have left them.
    static constexpr bool PACKED_GEN = L == 32 && RT > 0 && RT <= 32 && RT * CT + 3 <= 4 * L;
    template <int BITS> __device__ __forceinline__ void generate_packed(uint32_t ep, int& iters, bool& capped) {
        constexpr int RR = RT > 0 ? RT : 1, CC = CT > 0 ? CT : 1;
        constexpr uint32_t cells = (BITS == 2) ? 0x55555555u : 0x09249249u;
        constexpr uint32_t cmask = (CC * BITS >= 32) ? cells : (cells & ((1u << ((CC * BITS) & 31)) - 1u));   // low bit of cells 0..C-1
        constexpr uint32_t hmask = cmask & ((1u << (((CC - 1) * BITS) & 31)) - 1u);                          // cells 0..C-2
        episode = ep;
        in_reset = true;
        rsc = 0ull;
        uint32_t cur = 0u, row = 0u;       // reset-stream cursor (fits 32 bits: max_iters * P draws), this lane's row
        // Window of the reset stream kept in shared memory as COLOURS: draws [wb, wend), wb a multiple of 4, 4 * L draws per
        // Philox pass (lane j computes block wb / 4 + j).  A redraw takes on average half the board, so one pass serves about
        // two iterations.  BITS == 2: the window is a stream of 2-bit cells (lane j stores the byte of its four cells) and a
        // row is one funnel shift out of two words; BITS == 3: one byte per draw.
        uint32_t wb = 0u, wend = 0u;
        int n_rows = RR, from = RR - 1;    // ref :96-97: the initial fill draws every row
        iters = 0; capped = false;
#pragma unroll 1
        for (;;) {
            // cells [0, n_rows * C) in row-major order <- the next draws (ref :97 / :129): cell i takes word cur + i
            const uint32_t n = (uint32_t)(n_rows * CC);
            if (cur + n > wend) {                                              // (cur & 3) + n <= 4 * L: one pass always suffices
                wb = cur & ~3u;
                wend = wb + 4u * L;
                uint32_t w[4];
                philox4x32_10<PHU>((wb >> 2) + (uint32_t)lane, ep, gid, 3u, p.key0, p.key1, w);
                const uint32_t c0 = __umulhi(w[0], (uint32_t)K), c1 = __umulhi(w[1], (uint32_t)K), c2 = __umulhi(w[2], (uint32_t)K),
                               c3 = __umulhi(w[3], (uint32_t)K);
                sync();
                if (BITS == 2) reinterpret_cast<uint8_t*>(s.wbuf)[lane] = (uint8_t)(c0 | (c1 << 2) | (c2 << 4) | (c3 << 6));
                else s.wbuf[lane] = c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);
                sync();
            }
            if (lane < n_rows) {
                const uint32_t first = cur - wb + (uint32_t)(lane * CC);       // this row's first draw, relative to the window
                if (BITS == 2) {
                    constexpr uint32_t rowmask = (CC * 2 >= 32) ? 0xffffffffu : ((1u << ((CC * 2) & 31)) - 1u);
                    const uint32_t bit = first * 2u;
                    row = __funnelshift_r(s.wbuf[bit >> 5], s.wbuf[(bit >> 5) + 1], bit & 31u) & rowmask;
                } else {
                    const uint8_t* bp = reinterpret_cast<const uint8_t*>(s.wbuf) + first;
                    uint32_t acc = 0u;
#pragma unroll
                    for (int c = 0; c < CC; ++c) acc |= (uint32_t)bp[c] << (c * BITS);
                    row = acc;
                }
            }
            cur += n;
            // line scan (ref :149-196 with every tile normal): H = left ends of horizontal triples in this lane's row,
            // V = bottoms of vertical triples ending in it
            const uint32_t up1 = (uint32_t)__shfl_up_sync(0xffffffffu, (int)row, 1), up2 = (uint32_t)__shfl_up_sync(0xffffffffu, (int)row, 2);
            const uint32_t e = packed_eq<BITS>(row ^ (row >> BITS)) & hmask;
            const uint32_t H = e & (e >> BITS);
            const uint32_t V = lane >= 2 ? (packed_eq<BITS>(row ^ up1) & packed_eq<BITS>(up1 ^ up2) & cmask) : 0u;
            const unsigned has = ballot(lane <= from && lane < RR && (H | V) != 0u);
            if (!has) break;                                                   // line-free
            if (iters >= p.max_iters) { status |= ST_RESET_CAP; capped = true; break; }
            ++iters;
            ++prof_iters;
            const int rs = 31 - __clz((int)has);                               // bottom-most row with an anchored line (ref :158-160)
            // first line of the list (ref :127-128): lowest column, vertical before horizontal at the same column
            const int cvl = V ? (__ffs((int)V) - 1) : 1024, chl = H ? (__ffs((int)H) - 1) : 1024;
            const int cv = shfl(cvl, rs), ch = shfl(chl, rs);
            int top = rs;
            if (cv <= ch) {
                // vertical: extend upwards while the cell above has the colour (ref :168-172); D bit r: rows r and r-1 agree at cv
                const unsigned D = ballot(lane >= 1 && lane < RR && (((row ^ up1) >> cv) & ((1u << BITS) - 1u)) == 0u);
                top = rs - __clz((int)~(D << (31 - rs)));
            }
            const int ri = min(RR - 1, top + 1);
            n_rows = ri + 1;
            from = min(RR - 1, max(rs, ri + 2));                               // rows below were line-free and their 3-windows are unchanged
        }
        rdc = (uint64_t)cur;
        // the board as bytes, for possible_move / shuffle, the mask and the pool entry
        sync();
        if (lane < RR) {
#pragma unroll
            for (int c = 0; c < CC; ++c) col[lane * CC + c] = (int8_t)(1 + (int)((row >> (c * BITS)) & ((1u << BITS) - 1u)));
        }
#pragma unroll 1
        for (int i = lane; i < RR * CC; i += L) typ[i] = 1;
        sync();
    }

    // move (ref :330-378) after the effectiveness gate, in the pieces the kernels schedule:
    // move_begin = counters, swap and the combination match (ref :343-361), returns is_combination_match -- its
    // gravity + refill (ref :362-364) is left to the first cascade_trip; cascade_trip = one trip of the cascade
    // loop (ref :367-376), returns false when the board is stable.  One call site of fall_and_refill / resolve_round.
    __device__ __forceinline__ bool move_begin(int i1, int i2) {
        n_new = 0; n_act = 0;                                // ref :343-347
        sync();
        if (lane == 0) {                                     // swap_coords (ref :355, :729-732)
            const int8_t a = col[i1], b = typ[i1];
            col[i1] = col[i2]; typ[i1] = typ[i2];
            col[i2] = a; typ[i2] = b;
        }
        sync();
        const int t1 = typ[i1], t2 = typ[i2];
        const bool comb = (not01(t1) && not01(t2)) || t1 < 0 || t2 < 0;  // ref :357-359
        sync();  // every lane has read the swapped types before the leader starts deleting
        if (comb) {
            const SlowOut o = slow_combination<L, RT, CT>(&s, &p, lane, gmask, gshift, env, i1, i2);  // ref :361
            n_new = o.n_new; n_act = o.n_act; status |= o.status;
        }
        return comb;
    }
    __device__ __forceinline__ bool cascade_trip(bool& pending_fall, int& elim) {
        int lines = 0;
        if (!pending_fall) lines = resolve_round();          // ref :369-373
        if (!pending_fall && lines == 0) return false;
        int e_cnt;
        fall_and_refill(e_cnt);                              // ref :362-364 / :374-376
        elim += e_cnt;
        pending_fall = false;
        return true;
    }
    __device__ void move_core(int i1, int i2, int& elim_out, int& is_comb) {
        bool pending_fall = move_begin(i1, i2);
        is_comb = pending_fall;
        int elim = 0;
#pragma unroll 1
        while (cascade_trip(pending_fall, elim)) {}
        elim_out = elim + n_new;                             // ref :378 (counters are uniform across lanes)
    }

    // TMG_FLAG_CONSTRUCTIVE_RESET: board number `ep` of this env from a constructive line-free sampler instead of the
    // reference's generate_board (whose redraw loop, ref :99-109, does not terminate for e.g. 32x32 / 7 colours).
    // NOT the reference's algorithm; the contract (include/tmg_b200.h): cells in row-major
    // order, each takes the next colour 1 + mulhi32(W, K) of stream 5 of the board (W = Philox(key, ctr = (k>>2, ep, env,
    // 5))[k&3], k = 0, 1, ...) that does not complete a triple with the two cells to its left or the two above it (at most
    // 64 draws per cell); a board without a possible move is drawn again from where the stream stands.  All lanes walk
    // the cells together (the draws of a cell depend on the cells before it), each computing a Philox block of the window.
    __device__ void generate_constructive(uint32_t ep, unsigned& effv, unsigned& effh) {
        uint32_t cur = 0u, wbase = 0u;
        bool have = false;
        int attempts = 0;
#pragma unroll 1
        for (;;) {
            sync();
#pragma unroll 1
            for (int i = lane; i < P; i += L) typ[i] = 1;
#pragma unroll 1
            for (int r = 0; r < R; ++r) {
                sync();                                      // the rows above are visible
                int l1 = -1, l2 = -2;
#pragma unroll 1
                for (int c = 0; c < C; ++c) {
                    const int u1 = r >= 2 ? (int)col[(r - 1) * C + c] : -1, u2 = r >= 2 ? (int)col[(r - 2) * C + c] : -2;
                    const int forbid_v = (u1 == u2) ? u1 : -1, forbid_h = (c >= 2 && l1 == l2) ? l1 : -1;
                    int k = 1;
#pragma unroll 1
                    for (int tries = 0;; ++tries) {
                        if (!have || cur - wbase >= (uint32_t)(4 * L)) {      // next window of the stream: one block per lane
                            sync();
                            wbase = cur & ~3u;
                            uint32_t w[4];
                            philox4x32_10<PHU>((wbase >> 2) + (uint32_t)lane, ep, gid, 5u, p.key0, p.key1, w);
#pragma unroll
                            for (int q = 0; q < 4; ++q) s.wbuf[4 * lane + q] = w[q];
                            have = true;
                            sync();
                        }
                        k = 1 + (int)__umulhi(s.wbuf[cur - wbase], (uint32_t)K);
                        ++cur;
                        if (k != forbid_v && k != forbid_h) break;
                        if (tries >= 63) { status |= ST_RESET_CAP; break; }   // (K <= 2: a line-free colour may not exist)
                    }
                    if (lane == 0) col[r * C + c] = (int8_t)k;
                    l2 = l1; l1 = k;
                }
            }
            sync();
            const bool any = mask_bits(effv, effh);          // possible_move (ref :558-569)
            if (any) break;
            if (++attempts >= p.max_iters) { status |= ST_RESET_CAP; break; }
        }
    }

    // generate_board (ref :95-112) for board number `ep` of this env: a pure function of (seed, env, ep)
    __device__ __forceinline__ void begin_generate(uint32_t ep) {
        episode = ep;
        in_reset = !p.use_inj;   // injected draws are one sequential stream
        rdc = 0ull; rsc = 0ull;
        draw_cells(P, true);     // ref :96-97
    }
    __device__ __forceinline__ void end_generate() { in_reset = false; }

    __device__ bool board_is_valid() {
        bool bad = false;
        for (int i = lane; i < P; i += L) {
            const int k = col[i], t = typ[i];
            bad |= !((t == -1 && k == 0) || (t >= 1 && t <= 4 && k >= 1 && k <= K));
        }
        return ballot(bad) == 0u;
    }
};

template <int L, int RT, int CT>
__device__ __noinline__ SlowOut slow_round(GroupSmem<L>* sm, const Params* pp, int lane, unsigned gmask, int gshift, int env,
                                           int n_new, int n_act, SlowScan ss) {
    Board<L, RT, CT> b(*sm, *pp, lane, gmask, gshift, env);
    b.n_new = n_new; b.n_act = n_act;
    typename Board<L, RT, CT>::Scan sc;
    sc.rstar = ss.rstar; sc.mv = ss.mv; sc.hs = ss.hs; sc.hcells = ss.hcells; sc.m = ss.m; sc.vtop = ss.vtop; sc.has_v = ss.has_v;
    sc.bits.E = ss.E; sc.bits.D = ss.D; sc.bits.T = ss.T; sc.bits.S = ss.S;
    SlowOut o;
    o.n = b.general_round(sc);
    o.n_new = b.n_new; o.n_act = b.n_act; o.status = b.status;
    return o;
}
template <int L, int RT, int CT>
__device__ __noinline__ void mirror_item(GroupSmem<L>* sm, const Params* pp, int lane, unsigned gmask, int gshift, int env, int what) {
    Board<L, RT, CT> b(*sm, *pp, lane, gmask, gshift, env);
    const Params& p = *pp;
    b.sync();
    if (what & STORE_ZERO_MASK) b.store_zero_mask();
    if (what & (MIRROR_BOARD_SMEM | MIRROR_BOARD_POOL)) {
        const int8_t* src = (what & MIRROR_BOARD_POOL) ? p.pool_board + (size_t)env * 2 * p.P : b.s.board;
        if (p.h_board) b.copy_board(p.h_board + (size_t)env * 2 * p.P, src, p.board_vecw);
        if (p.h_board_packed) b.mirror_board_packed(src);
    }
    if (p.h_mask || p.h_mask_bits) {
        if (what & MIRROR_MASK_POOL) {
            b.copy_mask(b.s.mask, p.pool_mask + (size_t)env * p.A);
            b.sync();
            b.mirror_mask(false);
        } else if (what & MIRROR_MASK_SMEM) b.mirror_mask(false);
        else if (what & MIRROR_MASK_ZERO) b.mirror_mask(true);
    }
}
template <int L, int RT, int CT>
__device__ __noinline__ FinishOut slow_finish(GroupSmem<L>* sm, const Params* pp, int lane, unsigned gmask, int gshift, int env,
                                              int do_playability, int do_generate, int next_ep, uint64_t dcur, uint64_t scur) {
    Board<L, RT, CT> b(*sm, *pp, lane, gmask, gshift, env);
    b.dcur = dcur; b.scur = scur;
    FinishOut o;
    o.effv = 0u; o.effh = 0u; o.shuffled = 0;
    if (do_playability) o.shuffled = b.playability(true, false, o.effv, o.effh);           // ref board.py:381-391
    if (do_generate) {                                                                     // ref board.py:95-112 for board `next_ep`
        if (pp->flags & FLAG_CONSTRUCTIVE_RESET) b.generate_constructive((uint32_t)next_ep, o.effv, o.effh);
        else {
            b.begin_generate((uint32_t)next_ep);
            b.playability(false, true, o.effv, o.effh);
            b.end_generate();
        }
    }
    o.status = b.status; o.dcur = b.dcur; o.scur = b.scur; o.last_S = b.last_S;
    return o;
}
template <int L, int RT, int CT>
__device__ __noinline__ SlowOut slow_combination(GroupSmem<L>* sm, const Params* pp, int lane, unsigned gmask, int gshift,
                                                 int env, int i1, int i2) {
    Board<L, RT, CT> b(*sm, *pp, lane, gmask, gshift, env);
    b.combination(i1, i2);
    SlowOut o;
    o.n = 0; o.n_new = b.n_new; o.n_act = b.n_act; o.status = b.status;
    return o;
}

#include "tmg_rb.cuh"

// Which engine runs the moves of a step kernel instantiation: the register-resident one (tmg_rb.cuh) for boards of up to
// 10 rows and 7 colours, the byte planes in shared memory otherwise (or when TMG_FLAG_BYTE_PLANES asks for them).
// The step kernels take the choice as a template argument (RBK), made by the host with rb_supported(): an instantiation
// carries the move code of one engine only.
template <int L, int RT> struct UsesRB { static constexpr bool maybe = L == 32 && RT <= 10; };
__host__ __device__ inline bool rb_supported(int lanes, int R, int K, uint32_t flags, int use_inj) {
    return lanes == 32 && R <= 10 && K <= 7 && !(flags & FLAG_BYTE_PLANES) && !use_inj;   // (injected refill: byte planes)
}
// A move on the register-resident engine for a Board whose byte planes are staged in shared memory: pack, move, mask,
// unpack.  Returns the eliminations (without num_new_specials); mask_ok <- effv / effh are the mask of the final board
// and a move is possible (otherwise the caller runs Board::playability on the byte planes: shuffle / literal rule).
template <int RT, int CT, bool INJ, typename B>
__device__ __forceinline__ int rb_move(RBoard<RT, CT, INJ>& rb, B& b, int i1, int i2, int& is_comb, unsigned& effv, unsigned& effh, bool& mask_ok) {
    rb.dcur = b.dcur;
    rb.status = 0u;
    rb.pack_from_smem();
    const int elim = rb.move(i1, i2, is_comb);
    bool literal = false;
    const bool any = rb.mask_bits(effv, effh, literal);
    mask_ok = any && !literal;
    b.sync();
    rb.unpack_to_smem();
    b.sync();
    b.dcur = rb.dcur; b.n_new = rb.n_new; b.n_act = rb.n_act; b.status |= rb.status;
    b.last_S = rb.bits_S();
    b.prof_rounds += rb.prof_rounds; b.prof_serial += rb.prof_general;   // (diagnostics: general-path rounds in the slot of the byte planes' slow-path cycles)
    if (rb.prof_on && rb.lane == 0) {      // cycles in scan + fast round / general path / fall + refill of this move
        atomicAdd(&rb.p.prof[rb.env * 8 + 4], rb.prof_cyc[0]);
        atomicAdd(&rb.p.prof[rb.env * 8 + 5], rb.prof_cyc[1]);
        atomicAdd(&rb.p.prof[rb.env * 8 + 6], rb.prof_cyc[2]);
    }
    return elim;
}

// ======================================================================================================
// kernels
// ======================================================================================================
template <int L> struct GroupCtx {
    int g, lane, env, gshift;
    unsigned gmask;
    bool idle;   // lanes of a warp beyond its last whole group (30, 31 when L = 10)
    __device__ GroupCtx() {
        const int warp = (int)threadIdx.x >> 5, wl = (int)threadIdx.x & 31;
        const int gw = wl / L;
        idle = gw >= Cfg<L>::GPW;
        g = warp * Cfg<L>::GPW + gw;
        lane = wl - gw * L;
        env = idle ? 0x7fffffff : (int)blockIdx.x * Cfg<L>::GPB + g;
        gshift = gw * L;
        gmask = Cfg<L>::LMASK << (gshift & 31);
    }
};

template <int L> __device__ __forceinline__ GroupSmem<L>& group_smem(int g) {
    extern __shared__ __align__(16) unsigned char tmg_smem_raw[];
    return reinterpret_cast<GroupSmem<L>*>(tmg_smem_raw)[g];
}

template <int L> __device__ __forceinline__ void write_step_outputs(const Params& p, int env, int lane, int timer,
                                                                    int reward, int terminated, int is_comb, int n_new,
                                                                    int n_act, int shuffled, bool write_timer = true) {
    if (lane == 0) {
        if (write_timer) {
            p.timer[env] = timer;
            p.moves_left[env] = p.num_moves - timer;
        }
        p.reward[env] = reward;
        p.terminated[env] = (uint8_t)terminated;
        p.is_comb[env] = (uint8_t)is_comb;
        p.new_specials[env] = n_new;
        p.activated[env] = n_act;
        p.shuffled[env] = (uint8_t)shuffled;
    }
}
template <typename B> __device__ __forceinline__ void merge_status(B& b, const Params& p) {
    const unsigned st = b.ror(b.status);
    if (st && b.lane == 0) p.status[b.env] |= st;
}

// ---- scheduling helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ctl_read(uint32_t* w) { return atomicAdd(w, 0u); }   // coherent read of a counter
// Called once by every warp (k_gate) / group (k_reset) of a launch after its appends.  The last caller records the
// range of refill requests this launch issued under its tag and, for a step, clears the other parity's work counters
// for the next step.  Stream order makes all of it visible to the launches that follow.
__device__ __forceinline__ void launch_bookkeeping(const Params& p, bool is_step) {
    if (p.commit_pregen) {   // the host launches k_pregen `pool_tag` after this launch: it serves the requests since the last batch
        const uint32_t tail = ctl_read(&p.ctl[CTL_REQ_TAIL]), prev = p.ctl[CTL_REQ_PREV];
        const int slot = p.pool_tag % PG_RING;
        p.ctl[CTL_PG_RANGE + 2 * slot] = prev;
        p.ctl[CTL_PG_RANGE + 2 * slot + 1] = tail;
        p.ctl[CTL_PG_HEAD + slot] = 0u;
        p.ctl[CTL_REQ_PREV] = tail;
    }
    p.ctl[CTL_RO_HEAD] = 0u;
    if (is_step) {
        const int nq = (p.seq & 1) ^ 1;
        p.ctl[CTL_WL_COUNT + nq] = 0u;
        p.ctl[CTL_WL_COUNT_LO + nq] = 0u;
        p.ctl[CTL_WL_HEAD + nq] = 0u;
    }
}
__device__ __forceinline__ void commit_launch(const Params& p, uint32_t callers, bool is_step) {
    __threadfence();
    if (atomicAdd(&p.ctl[CTL_DONE], 1u) != callers - 1u) return;
    __threadfence();
    launch_bookkeeping(p, is_step);
    p.ctl[CTL_DONE] = 0u;
}
__device__ __forceinline__ void action_to_cells(int a, int R, int C, int& i1, int& i2) {  // ref :80-91
    const int nv = C * (R - 1);
    if (a < nv) { i1 = a; i2 = a + C; }
    else {
        const int j = a - nv;
        const int r = j / (C - 1), c = j - r * (C - 1);
        i1 = r * C + c; i2 = i1 + 1;
    }
}

// TileMatchEnv.reset (ref tile_match_env.py:84-91) for the selected envs
template <int L, int RT, int CT> __global__ void __launch_bounds__(Cfg<L>::THREADS) k_reset(const __grid_constant__ Params p) {
    const GroupCtx<L> gc;
    if (gc.env >= p.N) return;
    if (!p.reset_mask || p.reset_mask[gc.env]) {
        Board<L, RT, CT> b(group_smem<L>(gc.g), p, gc.lane, gc.gmask, gc.gshift, gc.env);
        b.load_cursors();
        unsigned effv = 0u, effh = 0u;
        if (p.init_boards) {
            b.load_board(p.init_boards, p.init_vecw);
            if (!b.board_is_valid()) b.status |= ST_INVALID_BOARD;
            b.mask_bits(effv, effh);
        } else {
            const int ep = p.episode[gc.env] + 1;    // generate_board (ref board.py:95-112)
            b.sync();
            if (p.flags & FLAG_CONSTRUCTIVE_RESET) {
                b.generate_constructive((uint32_t)ep, effv, effh);
            } else if (Board<L, RT, CT>::PACKED_GEN && !p.use_inj && p.K <= 8 && CT * (p.K <= 4 ? 2 : 3) <= 32) {
                int iters = 0;                       // line removal on packed rows, as in k_pregen
                bool capped = false;
                if (p.K <= 4) b.template generate_packed<2>((uint32_t)ep, iters, capped);
                else b.template generate_packed<3>((uint32_t)ep, iters, capped);
                if (capped) b.mask_bits(effv, effh);
                else b.playability(true, true, effv, effh, iters);
            } else {
                b.begin_generate((uint32_t)ep);
                b.playability(false, true, effv, effh);
            }
            b.end_generate();
            if (gc.lane == 0) {
                p.episode[gc.env] = ep;
                if (p.req_ring) p.req_ring[atomicAdd(&p.ctl[CTL_REQ_TAIL], 1u) & p.req_mask] = make_uint2((uint32_t)gc.env, (uint32_t)(ep + 1));   // next board -> pool
            }
        }
        b.store_board();
        b.store_cursors();
        if (!(p.flags & FLAG_NO_MASK)) { b.mask_to_smem(effv, effh); b.store_mask(); }
        merge_status(b, p);
        {
            const int nsp = min(255, b.radd(__popc(b.last_S)));
            if (gc.lane == 0) p.n_special[gc.env] = (uint8_t)nsp;
        }
        write_step_outputs<L>(p, gc.env, gc.lane, 0, 0, 0, 0, 0, 0, 0);
    }
    if (gc.lane == 0) commit_launch(p, (uint32_t)p.N, false);
}

// TileMatchEnv.step (ref tile_match_env.py:93-112), part 1: one thread per env.  Timer / termination / error checks, the
// effectiveness gate (ref board.py:352: the maintained mask IS is_move_effective of the current board, so a no-op
// step reads one byte and never touches its board) and the outputs of a step that changes nothing.  Envs that need
// board work -- an effective move, a new board, a zeroed mask -- go to the work list.
#ifndef TMG_GATE_EPT
#define TMG_GATE_EPT 2   // measured on B200 (65 536 envs): 1 and 2 envs per thread 371-373 M steps/s, 4 envs per thread 362 M
#endif
enum { GATE_EPT = TMG_GATE_EPT };   // envs per thread of k_gate: four independent chains of dependent loads per thread, a quarter of the atomics
// the envs [chunk * 32 GATE_EPT, (chunk + 1) * 32 GATE_EPT) by one warp (lane wl takes env0 + 32 k).
// (Gating inside the persistent step kernel -- a phase of k_work with a counter barrier behind it -- was built and measured
// SLOWER than this separate launch: 0.190 vs 0.171 ms per 65 536-env step; see DESIGN.md.)
__device__ __forceinline__ void gate_chunk(const Params& p, int chunk, int wl) {
    const int env0 = chunk * (32 * GATE_EPT) + wl;
    const int q = p.seq & 1;
    bool heavy[GATE_EPT], hi[GATE_EPT], req[GATE_EPT];
    uint32_t packed[GATE_EPT], req_ep[GATE_EPT];
#pragma unroll
    for (int k = 0; k < GATE_EPT; ++k) {
        const int env = env0 + 32 * k;
        heavy[k] = false; hi[k] = false; req[k] = false; packed[k] = 0u; req_ep[k] = 0u;
        if (env >= p.N) continue;
        int timer = p.timer[env];
        const int action = p.actions[env];
        bool eff = false, regenerate = false, fault = false;
        if (timer < 0 || timer >= p.num_moves) {
            if (p.autoreset == AUTORESET_NEXT_STEP && timer >= p.num_moves) regenerate = true;   // this call is the reset
            else { p.status[env] |= ST_NEEDS_RESET; fault = true; }                             // ref tile_match_env.py:94-95
        } else if (action < 0 || action >= p.A) {                                               // ref tile_match_env.py:97
            p.status[env] |= ST_BAD_ACTION;
            fault = true;
        } else if (!(p.flags & FLAG_NO_MASK)) {
            eff = p.mask[(size_t)env * p.A + action] != 0;
        } else {
            const int8_t* bd = p.board + (size_t)env * 2 * p.P;
            int i1, i2;
            action_to_cells(action, p.R, p.C, i1, i2);
            eff = effective_literal(bd, bd + p.P, p.R, p.C, i1, i2);
        }
        int terminated = 0;
        bool zero_mask = false;
        if (!fault) {
            if (!regenerate) {
                ++timer;                                                           // ref tile_match_env.py:100-101
                terminated = timer == p.num_moves;
                if (terminated) {
                    if (p.autoreset == AUTORESET_SAME_STEP) { regenerate = true; timer = 0; }
                    else zero_mask = !(p.flags & FLAG_NO_MASK);                    // ref tile_match_env.py:119-120
                }
            } else {
                timer = 0;
            }
            p.timer[env] = timer;
            p.moves_left[env] = p.num_moves - timer;
        }
        p.reward[env] = 0;
        p.terminated[env] = (uint8_t)terminated;
        p.is_comb[env] = 0;
        p.new_specials[env] = 0;
        p.activated[env] = 0;
        p.shuffled[env] = 0;
        if (p.h_reward) p.h_reward[env] = 0;                 // host mirror: coalesced stores over PCIe
        if (p.h_terminated) p.h_terminated[env] = (uint8_t)terminated;
        if (p.h_moves_left && !fault) p.h_moves_left[env] = p.num_moves - timer;
        // The next board is a pure function of (seed, env, episode): take it from the pool k_pregen filled ahead of
        // time when it is there, generate it inside the step otherwise (same result either way).
        bool from_pool = false;
        if (regenerate) {
            const int ep_now = p.episode[env];
            from_pool = !p.use_inj && p.pool_episode[env] == ep_now + 1;
            req_ep[k] = (uint32_t)(ep_now + 2);              // this step moves the env to board ep_now + 1: the pool needs the one after
        }
        heavy[k] = eff || regenerate || zero_mask;
        hi[k] = eff && p.n_special[env] >= PRI_SPECIALS;
        req[k] = regenerate && p.req_ring != nullptr;
        packed[k] = ((uint32_t)action & IT_ACTION) | (eff ? IT_EFF : 0u) | (regenerate ? IT_REGEN : 0u) |
                    (zero_mask ? IT_ZERO_MASK : 0u) | (from_pool ? IT_FROM_POOL : 0u);
    }
    // one append per list and warp: positions by ballot
    unsigned hm[GATE_EPT], lm[GATE_EPT], rm[GATE_EPT];
    uint32_t nh = 0u, nl = 0u, nr = 0u;
    TMG_SITE_HERE
#pragma unroll
    for (int k = 0; k < GATE_EPT; ++k) {
        hm[k] = __ballot_sync(0xffffffffu, hi[k]);
        lm[k] = __ballot_sync(0xffffffffu, heavy[k] && !hi[k]);
        rm[k] = __ballot_sync(0xffffffffu, req[k]);
        nh += (uint32_t)__popc(hm[k]); nl += (uint32_t)__popc(lm[k]); nr += (uint32_t)__popc(rm[k]);
    }
    uint32_t hbase = 0u, lbase = 0u, rbase = 0u;
    if (wl == 0) {
        if (nh) hbase = atomicAdd(&p.ctl[CTL_WL_COUNT + q], nh);
        if (nl) lbase = atomicAdd(&p.ctl[CTL_WL_COUNT_LO + q], nl);
        if (nr) rbase = atomicAdd(&p.ctl[CTL_REQ_TAIL], nr);
    }
    hbase = (uint32_t)__shfl_sync(0xffffffffu, (int)hbase, 0);
    lbase = (uint32_t)__shfl_sync(0xffffffffu, (int)lbase, 0);
    rbase = (uint32_t)__shfl_sync(0xffffffffu, (int)rbase, 0);
    const unsigned lt = (1u << wl) - 1u;
#pragma unroll
    for (int k = 0; k < GATE_EPT; ++k) {
        if (heavy[k]) {
            uint2 it; it.x = (uint32_t)(env0 + 32 * k); it.y = packed[k];
            p.wl_items[hi[k] ? hbase + (uint32_t)__popc(hm[k] & lt) : (uint32_t)p.N - 1u - (lbase + (uint32_t)__popc(lm[k] & lt))] = it;
        }
        if (req[k]) p.req_ring[(rbase + (uint32_t)__popc(rm[k] & lt)) & p.req_mask] = make_uint2((uint32_t)(env0 + 32 * k), req_ep[k]);
        hbase += (uint32_t)__popc(hm[k]); lbase += (uint32_t)__popc(lm[k]); rbase += (uint32_t)__popc(rm[k]);
    }
    __syncwarp(0xffffffffu);
}
__global__ void __launch_bounds__(128) k_gate(const __grid_constant__ Params p) {
    const int wl = (int)threadIdx.x & 31;
    gate_chunk(p, (int)blockIdx.x * 4 + ((int)threadIdx.x >> 5), wl);
    if (wl == 0) commit_launch(p, gridDim.x * (blockDim.x >> 5), true);
}

#ifndef TMG_STEP_MIN_BLOCKS
#define TMG_STEP_MIN_BLOCKS 8   // <= 64 registers/thread: 32 warps/SM.  Measured (65 536 envs): 4 blocks 183 M, 6 198 M, 8 228 M, 10 203 M steps/s
#endif
// groups of a warp: all done? -- also the point where the groups of the warp reconverge (a __syncwarp over the warp)
template <int L> __device__ __forceinline__ bool warp_all_done(bool done) {
#ifdef TMG_EMU
    return done;               // the emulator runs one group at a time
#else
    constexpr unsigned wmask = (Cfg<L>::GPW * L >= 32) ? 0xffffffffu : ((1u << ((Cfg<L>::GPW * L) & 31)) - 1u);
    return __all_sync(wmask, done);
#endif
}
// next index of a shared cursor for this group, or `n` and above when the list is exhausted
template <int L> __device__ __forceinline__ uint32_t pop_item(const GroupCtx<L>& gc, uint32_t* head) {
    uint32_t idx = 0u;
    TMG_SITE_HERE
    const unsigned gm = L == 32 ? 0xffffffffu : gc.gmask;
    __syncwarp(gm);                        // the previous item's shared-memory traffic is complete on every lane
    if (gc.lane == 0) idx = atomicAdd(head, 1u);
    return (uint32_t)__shfl_sync(gm, (int)idx, L == 32 ? 0 : gc.gshift);
}

// end of a work-list item: playability (ref board.py:381-391), the next board if the episode ended, state and outputs
template <int L, int RT, int CT, bool MIR> __device__ __forceinline__ void finish_item(Board<L, RT, CT>& b, const Params& p, uint32_t packed,
                                                                            int elim, int is_comb, long long prof_t0,
                                                                            bool mask_ok = false, unsigned effv0 = 0u, unsigned effh0 = 0u) {
    const int lane = b.lane, env = b.env;
    const bool want_mask = !(p.flags & FLAG_NO_MASK);
    const bool eff = packed & IT_EFF, regenerate = packed & IT_REGEN, zero_mask = packed & IT_ZERO_MASK,
               from_pool = packed & IT_FROM_POOL;
    unsigned effv = effv0, effh = effh0;                                   // mask_ok: the mask of the moved board is known and has a move
    int shuffled = 0;
    const int reward = elim + b.n_new;                                     // ref :378 (counters are uniform across lanes)
    const int n_new = b.n_new, n_act = b.n_act;
    const int next_ep = regenerate ? p.episode[env] + 1 : 0;
    const bool inline_gen = regenerate && !from_pool;
    const bool dirty = eff || inline_gen;
    // playability of the moved board (only when its mask is not known yet) and generate_board of the next episode when
    // the pool does not hold it: both rare, both out of line
    if ((eff && !mask_ok) || inline_gen) {
        b.sync();
        const FinishOut fo = slow_finish<L, RT, CT>(&b.s, &p, lane, b.gmask, b.gshift, env, eff && !mask_ok, inline_gen, next_ep, b.dcur, b.scur);
        effv = fo.effv; effh = fo.effh; shuffled = fo.shuffled;
        b.status |= fo.status; b.dcur = fo.dcur; b.scur = fo.scur; b.last_S = fo.last_S;
    }
    const bool mirrored = p.h_board || p.h_board_packed || p.h_mask || p.h_mask_bits;
    int mirror = 0;
    if (from_pool) {                       // board and mask of the new episode come straight from the pool
        b.sync();
        b.copy_board(p.board + (size_t)env * 2 * p.P, p.pool_board + (size_t)env * 2 * p.P, p.board_vecw);
        if (want_mask) b.copy_mask(p.mask + (size_t)env * p.A, p.pool_mask + (size_t)env * p.A);
        b.status |= p.pool_status[env];
        if (eff) b.store_cursors();
        mirror = MIRROR_BOARD_POOL | (want_mask ? MIRROR_MASK_POOL : 0);
    } else {
        if (dirty) {
            b.store_board(); b.store_cursors();
            mirror = MIRROR_BOARD_SMEM;
        }
        if (want_mask) {
            if (zero_mask) mirror |= STORE_ZERO_MASK | MIRROR_MASK_ZERO;
            else if (dirty) { b.mask_to_smem(effv, effh); b.store_mask(); mirror |= MIRROR_MASK_SMEM; }
        }
    }
    // MIR: the instantiation that runs while a host mirror is bound writes it through inline (it is on the path of every
    // changed env then); the other instantiation carries no mirror code at all.  The all-zero mask of a terminal step is
    // rare either way.
    if constexpr (MIR) {
        if (mirror & STORE_ZERO_MASK) b.store_zero_mask();
        if (mirror & (MIRROR_BOARD_SMEM | MIRROR_BOARD_POOL)) {
            b.sync();
            const int8_t* src = (mirror & MIRROR_BOARD_POOL) ? p.pool_board + (size_t)env * 2 * p.P : b.s.board;
            if (p.h_board) b.copy_board(p.h_board + (size_t)env * 2 * p.P, src, p.board_vecw);
            if (p.h_board_packed) b.mirror_board_packed(src);
        }
        if (p.h_mask || p.h_mask_bits) {
            if (mirror & MIRROR_MASK_POOL) {
                b.copy_mask(b.s.mask, p.pool_mask + (size_t)env * p.A);
                b.sync();
                b.mirror_mask(false);
            } else if (mirror & MIRROR_MASK_SMEM) { b.sync(); b.mirror_mask(false); }
            else if (mirror & MIRROR_MASK_ZERO) b.mirror_mask(true);
        }
    } else {
        if ((mirrored && mirror) || (mirror & STORE_ZERO_MASK)) mirror_item<L, RT, CT>(&b.s, &p, lane, b.gmask, b.gshift, env, mirror);
    }
    merge_status(b, p);
    if (dirty || from_pool) {              // scheduling hint for the next step: special tiles on the board (a fresh board has none)
        const int nsp = regenerate ? 0 : min(255, b.radd(__popc(b.last_S)));
        if (lane == 0) p.n_special[env] = (uint8_t)nsp;
    }
    if (lane == 0) {
        if (regenerate) p.episode[env] = next_ep;
        if (eff) {                         // k_gate wrote the outputs of a step that changes nothing
            if (p.h_reward) p.h_reward[env] = reward;
            p.reward[env] = reward;
            p.is_comb[env] = (uint8_t)is_comb;
            p.new_specials[env] = n_new;
            p.activated[env] = n_act;
            p.shuffled[env] = (uint8_t)shuffled;
        }
        if (p.prof) {
            p.prof[env * 8 + 0] = (uint32_t)(clock64() - prof_t0);
            p.prof[env * 8 + 1] = b.prof_serial;
            p.prof[env * 8 + 2] = b.prof_rounds;
            p.prof[env * 8 + 3] = b.prof_iters;
        }
    }
}

// Part 2 of a step: the envs of the work list -- the move (ref board.py:330-395) and / or the next board.
// Persistent groups pop items until the list is empty: first the items k_gate expects to cascade long (boards rich in
// special tiles), so that the longest item of the launch runs beside the bulk instead of trailing it.
// (A warp-level state machine that aligns the cascade rounds of the warp's groups, like k_pregen's loop, was measured
// no faster here: collectives with a group mask make the groups of a warp separate instruction streams anyway.)
template <int L, int RT, int CT, bool RBK, bool MIR> __global__ void __launch_bounds__(Cfg<L>::THREADS, TMG_STEP_MIN_BLOCKS) k_work(const __grid_constant__ Params p) {
    const GroupCtx<L> gc;
    if (gc.idle) return;
    const int q = p.seq & 1;
    const uint32_t n_hi = p.ctl[CTL_WL_COUNT + q], n = n_hi + p.ctl[CTL_WL_COUNT_LO + q];
    Board<L, RT, CT> b(group_smem<L>(gc.g), p, gc.lane, gc.gmask, gc.gshift, 0);
#pragma unroll 1
    for (;;) {
        const uint32_t idx = pop_item<L>(gc, &p.ctl[CTL_WL_HEAD + q]);
        if (idx >= n) break;
        const uint2 it = p.wl_items[idx < n_hi ? idx : (uint32_t)p.N - 1u - (idx - n_hi)];   // high priority from the front, the rest from the back
        b.rebind((int)it.x);
        b.prof_serial = 0u; b.prof_rounds = 0u; b.prof_iters = 0u;
        const uint32_t packed = it.y;
        const long long prof_t0 = p.prof ? clock64() : 0;
        int elim = 0, is_comb = 0;
        b.sync();
        if (packed & (IT_EFF | IT_REGEN)) b.load_cursors();
        bool mask_ok = false;
        unsigned effv = 0u, effh = 0u;
        if (packed & IT_EFF) {
            b.load_board(p.board, p.board_vecw);
            int i1, i2;
            b.action_cells((int)(packed & IT_ACTION), i1, i2);
            if constexpr (RBK && UsesRB<L, RT>::maybe) {
                RBoard<RT, CT> rb(b.s, p, gc.lane, b.env);
                elim = rb_move(rb, b, i1, i2, is_comb, effv, effh, mask_ok);
            } else {
                bool pending_fall = b.move_begin(i1, i2);
                is_comb = pending_fall;
#pragma unroll 1
                while (b.cascade_trip(pending_fall, elim)) {}
            }
        }
        finish_item<L, RT, CT, MIR>(b, p, packed, elim, is_comb, prof_t0, mask_ok, effv, effh);
    }
}

// Fused multi-step rollout (the caller loop of src/examples/random_agent.py:12-31 with given actions):
// T successive TileMatchEnv.step calls per env in ONE launch.  A group keeps its env's board in shared memory and its
// legal-move mask as bitboards in registers across the steps, so a step that changes nothing is a bit test, and no
// env ever waits for the slowest cascade of the batch.  State and outputs after the call are those of T tmg_step
// calls; per-step rewards / terminations go to [T][N] arrays.
template <int L, int RT, int CT, bool RBK> __global__ void __launch_bounds__(Cfg<L>::THREADS, TMG_STEP_MIN_BLOCKS) k_rollout(const __grid_constant__ Params p) {
    const GroupCtx<L> gc;
    if (gc.idle) return;
    const int lane = gc.lane;
    const bool want_mask = !(p.flags & FLAG_NO_MASK);
#pragma unroll 1
    for (;;) {
        const uint32_t idx = pop_item<L>(gc, &p.ctl[CTL_RO_HEAD]);
        if (idx >= (uint32_t)p.N) break;
        const int env = (int)idx;
        Board<L, RT, CT> b(group_smem<L>(gc.g), p, lane, gc.gmask, gc.gshift, env);
        b.load_cursors();
        int timer = p.timer[env], episode = p.episode[env];
        b.load_board(p.board, p.board_vecw);
        unsigned effv = 0u, effh = 0u;
        bool have_mask = false, regenerated = false, pool_used = false, touched = false;
        int reward = 0, is_comb = 0, shuffled = 0, terminated = 0, n_new = 0, n_act = 0;
        uint32_t aw[4] = {0u, 0u, 0u, 0u};          // on-device policy: the Philox block of the action stream in use
        uint64_t aw_blk = ~0ull;
#pragma unroll 1
        for (int t = 0; t < p.T; ++t) {
            int action = p.policy ? 0 : p.actions[(size_t)t * p.N + env];
            reward = 0; is_comb = 0; shuffled = 0; terminated = 0; n_new = 0; n_act = 0;
            bool eff = false, regenerate = false, fault = false;
            const bool live = !(timer < 0 || timer >= p.num_moves);
            if (live && p.policy) {
                // The agent of src/examples/random_agent.py:12-31 inside the kernel.  Word k = board number * num_moves
                // + timer of the env's action stream (stream 2) decides: uniform over all actions, or the
                // mulhi32(word, n)-th of the n effective actions in index order (ref tile_match_env.py:118-124).
                if (!have_mask) { b.mask_bits(effv, effh); have_mask = true; }
                const uint64_t k = (uint64_t)(unsigned)episode * (uint64_t)p.num_moves + (uint64_t)timer;
                if ((k >> 2) != aw_blk) {
                    aw_blk = k >> 2;
                    philox4x32_10((uint32_t)aw_blk, (uint32_t)(aw_blk >> 32), b.gid, 2u, p.key0, p.key1, aw);
                }
                const uint32_t word = (k & 3) == 0 ? aw[0] : (k & 3) == 1 ? aw[1] : (k & 3) == 2 ? aw[2] : aw[3];
                const int n_eff = p.policy == POLICY_MASK ? b.radd(__popc(effv) + __popc(effh)) : 0;
                if (n_eff == 0) action = (int)__umulhi(word, (uint32_t)p.A);
                else {
                    int kth = (int)__umulhi(word, (uint32_t)n_eff);
                    action = -1;
#pragma unroll 1
                    for (int r = 0; r < 2 * b.R && action < 0; ++r) {             // vertical actions row by row, then horizontal
                        const bool vert = r < b.R;
                        const int rr = vert ? r : r - b.R;
                        const unsigned m = b.ballot(((vert ? effv : effh) >> rr) & 1u);
                        const int c = __popc(m);
                        if (kth < c) action = (vert ? rr * b.C : b.C * (b.R - 1) + rr * (b.C - 1)) + b.nth_bit(m, kth);
                        else kth -= c;
                    }
                }
            }
            if (p.ro_actions && lane == 0) p.ro_actions[(size_t)t * p.N + env] = action;
            if (!live) {
                if (p.autoreset == AUTORESET_NEXT_STEP && timer >= p.num_moves) regenerate = true;   // this step is the reset
                else { b.status |= ST_NEEDS_RESET; fault = true; }                                  // ref tile_match_env.py:94-95
            } else if (action < 0 || action >= p.A) {                                               // ref tile_match_env.py:97
                b.status |= ST_BAD_ACTION;
                fault = true;
            } else {
                if (!have_mask) { b.mask_bits(effv, effh); have_mask = true; }
                // effectiveness gate (ref board.py:352): bit (r, c) of the mask bitboards, held by lane c
                const int nv = b.C * (b.R - 1);
                const bool vert = action < nv;
                const int j = vert ? action : action - nv, w = vert ? b.C : b.C - 1;
                const int r = j / w, c = j - r * w;
                eff = ((unsigned)b.shfl((int)(vert ? effv : effh), c) >> r) & 1u;
            }
            if (!fault) {
                if (!regenerate) {
                    ++timer;                                                       // ref tile_match_env.py:100-101
                    terminated = timer == p.num_moves;
                    if (terminated && p.autoreset == AUTORESET_SAME_STEP) { regenerate = true; timer = 0; }
                } else {
                    timer = 0;
                }
                if (eff) {
                    int i1, i2;
                    b.action_cells(action, i1, i2);
                    bool mask_ok = false;
                    if constexpr (RBK && UsesRB<L, RT>::maybe) {
                        RBoard<RT, CT> rb(b.s, p, lane, env);
                        reward = rb_move(rb, b, i1, i2, is_comb, effv, effh, mask_ok);
                        reward += b.n_new;                                         // ref board.py:378
                    } else {
                        b.move_core(i1, i2, reward, is_comb);
                    }
                    if (!mask_ok) shuffled = b.playability(true, false, effv, effh);   // ref board.py:381-391
                    n_new = b.n_new; n_act = b.n_act;
                    touched = true;
                }
                if (regenerate) {
                    ++episode;
                    // The pool entry may be completed by a k_pregen launch that runs beside this kernel: its episode number is
                    // published last (after a fence), so read it first, fence, then read the entry past the L1.
                    bool in_pool = false;
                    if (!p.use_inj && !pool_used) {
                        int pe = 0;
                        if (lane == 0) pe = (int)ctl_read(reinterpret_cast<uint32_t*>(p.pool_episode + env));
                        in_pool = b.shfl(pe, 0) == episode;
                    }
                    if (in_pool) {                                                  // the pool holds this board
                        __threadfence();
                        b.sync();
                        copy_bytes_cg<L>(b.s.board, p.pool_board + (size_t)env * 2 * p.P, 2 * p.P, p.board_vecw, lane);
                        if (lane == 0) b.status |= ctl_read(p.pool_status + env);
                        b.sync();
                        b.mask_bits(effv, effh);
                        pool_used = true;
                    } else if (p.flags & FLAG_CONSTRUCTIVE_RESET) {
                        b.generate_constructive((uint32_t)episode, effv, effh);
                    } else {
                        b.begin_generate((uint32_t)episode);                       // ref board.py:95-112
                        b.playability(false, true, effv, effh);
                        b.end_generate();
                    }
                    have_mask = true; regenerated = true; touched = true;
                }
            }
            if (lane == 0) {
                if (p.ro_reward) p.ro_reward[(size_t)t * p.N + env] = reward;
                if (p.ro_terminated) p.ro_terminated[(size_t)t * p.N + env] = (uint8_t)terminated;
            }
        }
        // state and last-step outputs, as T tmg_step calls would leave them
        const bool terminal = timer == p.num_moves;                               // all-zero mask (ref tile_match_env.py:119-120)
        if (touched) { b.store_board(); b.store_cursors(); }
        if (want_mask && p.T > 0) {
            if (terminal) b.store_zero_mask();
            else if (touched) { b.mask_to_smem(effv, effh); b.store_mask(); }
        }
        if (p.h_board && touched) b.copy_board(p.h_board + (size_t)env * 2 * p.P, b.s.board, p.board_vecw);
        if (p.h_board_packed && touched) b.mirror_board_packed(b.s.board);
        if (want_mask && p.T > 0 && (terminal || touched)) b.mirror_mask(terminal);
        merge_status(b, p);
        if (touched) {
            const int nsp = min(255, b.radd(__popc(b.last_S)));
            if (lane == 0) p.n_special[env] = (uint8_t)nsp;
        }
        if (lane == 0) {
            p.episode[env] = episode;
            if (regenerated && p.req_ring)    // next board -> pool
                p.req_ring[atomicAdd(&p.ctl[CTL_REQ_TAIL], 1u) & p.req_mask] = make_uint2((uint32_t)env, (uint32_t)(episode + 1));
        }
        write_step_outputs<L>(p, env, lane, timer, reward, terminated, is_comb, n_new, n_act, shuffled, timer >= 0);
        if (lane == 0) {
            if (p.h_reward) p.h_reward[env] = reward;
            if (p.h_terminated) p.h_terminated[env] = (uint8_t)terminated;
            if (p.h_moves_left && timer >= 0) p.h_moves_left[env] = p.num_moves - timer;
        }
        b.sync();
    }
    if (lane == 0) commit_launch(p, gridDim.x * (uint32_t)Cfg<L>::GPB, false);
}

// (Two boards per warp -- half a warp holds the <= 16 packed rows of a board, one instruction stream serves both -- was
// built and measured SLOWER on B200: 340 M vs 368 M steps/s at 65 536 envs, 691 M vs 706 M at 1 M.  A redraw of n rows needs
// 10 n / 4 Philox blocks, which is 40 % of an iteration; two redraws together overflow the warp's 32 lanes and take two
// passes, so nothing is shared there, and the bookkeeping of two boards costs more than the rest saves.)
// Pool refill: generate_board (ref board.py:95-112) of the next board of every env whose request this launch serves.
// Runs on a side stream, off the step path; touches no env state.  The fixed small shapes remove the lines on packed
// rows (Board::generate_packed); the others run the byte-plane loop, one scan + redraw iteration per trip.
// The kernel is issue-bound (ncu: 70 % issue slots busy, instruction-cache hit rate 99.98 %), and a board needs ~86
// identical scan + redraw iterations, so the groups of a warp are kept CONVERGED: the loop below is one iteration per
// trip for every group of the warp, with a warp-wide reconvergence point at the top, and a group that finishes its
// board takes the next request inside the same loop.  One warp instruction then serves all the boards of the warp.
template <int L, int RT, int CT> __global__ void __launch_bounds__(Cfg<L>::THREADS, TMG_STEP_MIN_BLOCKS) k_pregen(const __grid_constant__ Params p) {
    const GroupCtx<L> gc;
    if (gc.idle) return;
    const int slot = p.pool_tag % PG_RING;
    const uint32_t start = p.ctl[CTL_PG_RANGE + 2 * slot], n = p.ctl[CTL_PG_RANGE + 2 * slot + 1] - start;
    Board<L, RT, CT> b(group_smem<L>(gc.g), p, gc.lane, gc.gmask, gc.gshift, 0);
    bool have = false, done = false, capped = false, staged = false;   // staged: the line-free board is already in shared memory
    bool masked = false;                                                // ... and so are its mask bits (cv, ch)
    unsigned cv = 0u, ch = 0u;
    int from = 0, iters = 0, ep = 0;
#pragma unroll 1
    for (;;) {
        if (warp_all_done<L>(done)) break;     // also the reconvergence point of the warp's groups
        if (!done && !have) {                  // take the next request
            const uint32_t idx = pop_item<L>(gc, &p.ctl[CTL_PG_HEAD + slot]);
            if (idx >= n) done = true;
            else {
                // The request names its board (issued as {env, board number}), so the result does not depend on when this
                // launch runs relative to the steps; a request the env has already moved past is skipped.
                const uint2 rq = p.req_ring[(start + idx) & p.req_mask];
                const int env = (int)rq.x;
                ep = (int)rq.y;
                if ((uint32_t)env < (uint32_t)p.N && p.pool_episode[env] != ep && p.episode[env] < ep) {
                    b.rebind(env);
                    b.sync();
                    capped = false; from = b.R - 1; iters = 0; have = true;
                    staged = false;
                    if (p.flags & FLAG_CONSTRUCTIVE_RESET) {
                        b.generate_constructive((uint32_t)ep, cv, ch);
                        staged = true; masked = true;
                    } else if (Board<L, RT, CT>::PACKED_GEN && !p.use_inj && p.K <= 8 && CT * (p.K <= 4 ? 2 : 3) <= 32) {
                        // fixed small shapes: the whole line removal on packed rows, then straight to the finish
                        if (p.K <= 4) b.template generate_packed<2>((uint32_t)ep, iters, capped);
                        else b.template generate_packed<3>((uint32_t)ep, iters, capped);
                        staged = true;
                    } else {
                        b.begin_generate((uint32_t)ep);        // ref :96-97
                    }
                }
            }
        }
        if (!have) continue;
        if (!staged && b.redraw_iteration(from, iters, capped)) continue;   // ref :99-101, one iteration
        // line-free: possible_move / shuffle (ref :102-109, rare) and the mask, then publish the pool entry
        unsigned effv = cv, effh = ch;
        if (masked) masked = false;                            // the constructive generator left the mask of its board
        else if (capped) b.mask_bits(effv, effh);
        else b.playability(true, true, effv, effh, iters);
        b.end_generate();
        b.sync();
        const int env = b.env;
        b.copy_board(p.pool_board + (size_t)env * 2 * p.P, b.s.board, p.board_vecw);
        if (!(p.flags & FLAG_NO_MASK)) {
            b.mask_to_smem(effv, effh);
            b.sync();
            b.copy_mask(p.pool_mask + (size_t)env * p.A, b.s.mask);
        }
        const unsigned st = b.ror(b.status);
        if (gc.lane == 0) p.pool_status[env] = st;
        __threadfence();   // the board must be visible before the entry is declared valid
        b.sync();
        if (gc.lane == 0) p.pool_episode[env] = ep;
        b.sync();
        have = false;
        if (p.pregen_one_shot) done = true;
    }
}

// _get_effective_actions for every env from its current board (ref tile_match_env.py:118-124)
template <int L> __global__ void __launch_bounds__(Cfg<L>::THREADS) k_mask(const __grid_constant__ Params p) {
    const GroupCtx<L> gc;
    if (gc.env >= p.N) return;
    Board<L> b(group_smem<L>(gc.g), p, gc.lane, gc.gmask, gc.gshift, gc.env);
    const bool terminal = p.timer[gc.env] == p.num_moves;
    b.sync();
    if (terminal) { b.store_zero_mask(); return; }
    b.load_board(p.board, p.board_vecw);
    unsigned effv, effh;
    b.mask_bits(effv, effh);
    b.mask_to_smem(effv, effh);
    b.store_mask();
}

// one engine primitive per env (known-answer replays of the reference's function-level tests)
template <int L> __global__ void __launch_bounds__(Cfg<L>::THREADS) k_debug(const __grid_constant__ Params p) {
    const GroupCtx<L> gc;
    if (gc.env >= p.N) return;
    const int env = gc.env, lane = gc.lane;
    Board<L> b(group_smem<L>(gc.g), p, lane, gc.gmask, gc.gshift, env);
    b.load_board(p.board, p.board_vecw);
    b.load_cursors();
    b.n_new = p.new_specials[env];
    b.n_act = p.activated[env];
    const int a0 = p.dbg_args ? p.dbg_args[env * 4 + 0] : 0, a1 = p.dbg_args ? p.dbg_args[env * 4 + 1] : 0;
    const int a2 = p.dbg_args ? p.dbg_args[env * 4 + 2] : 0, a3 = p.dbg_args ? p.dbg_args[env * 4 + 3] : 0;
    int result = 0;
    const int C = p.C;
    b.sync();
    const int op = p.dbg_op & 0xff;
    // The primitives of the register-resident engine (tmg_rb.cuh) are exercised through the same entry point: boards it
    // supports run on it unless OP_BYTE_PLANES asks for the byte-plane implementation.
    if constexpr (L == 32) {
        const bool rb_op = op == OP_GRAVITY || op == OP_RESOLVE_ROUND || op == OP_ACTIVATE || op == OP_COMBINE || op == OP_MOVE ||
                           op == OP_COUNT_LINES || op == OP_LINES;
        if (rb_op && !(p.dbg_op & OP_BYTE_PLANES) && rb_supported(32, p.R, p.K, p.flags, 0)) {
            RBoard<0, 0, true> rb(b.s, p, lane, env);
            rb.dcur = b.dcur;
            rb.n_new = b.n_new; rb.n_act = b.n_act;
            rb.pack_from_smem();
            int is_comb = 0, shuffled = 0;
            bool to_bytes = true, fallback = false;
            unsigned effv = 0u, effh = 0u;
            switch (op) {
                case OP_GRAVITY:
                    result = rb.radd(__popc(rb.bits_tz() & (lane < C ? 0xffffffffu : 0u)));
                    rb.gravity_column();
                    break;
                case OP_RESOLVE_ROUND: rb.literal_rounds = true; result = rb.resolve_round(); break;
                case OP_ACTIVATE: rb.activate(a0 * C + a1, a2, a3 == 0); break;
                case OP_COMBINE: rb.combination(a0 * C + a1, a2 * C + a3); break;
                case OP_MOVE: {
                    const int i1 = a0 * C + a1, i2 = a2 * C + a3;
                    bool eff = false;
                    if (lane == 0) eff = effective_literal(b.col, b.typ, p.R, C, i1, i2);
                    eff = b.shfl((int)eff, 0) != 0;
                    rb.n_new = 0; rb.n_act = 0;
                    if (eff) {
                        result = rb.move(min(i1, i2), max(i1, i2), is_comb);
                        result += rb.n_new;
                        bool literal = false;
                        fallback = !rb.mask_bits(effv, effh, literal) || literal;
                    }
                    break;
                }
                default: {                                   // OP_COUNT_LINES / OP_LINES
                    const typename RBoard<0, 0, true>::Scan sc = rb.scan_lines();
                    result = sc.rstar < 0 ? 0 : rb.build_line_table(sc);
                    to_bytes = false;
                    if (op == OP_LINES && p.dbg_out) {
                        uint32_t* out = p.dbg_out + (size_t)env * LINES_WORDS;
                        if (lane == 0) out[0] = (uint32_t)result;
                        if (lane < result) { out[1 + 2 * lane] = b.s.line_key[lane]; out[2 + 2 * lane] = b.s.line_mask[lane]; }
                    }
                    break;
                }
            }
            if (to_bytes) {
                b.sync();
                rb.unpack_to_smem();
                b.sync();
            }
            b.dcur = rb.dcur; b.n_new = rb.n_new; b.n_act = rb.n_act; b.status |= rb.status;
            if (fallback) shuffled = b.playability(true, false, effv, effh);
            if (op == OP_MOVE && lane == 0) { p.is_comb[env] = (uint8_t)is_comb; p.shuffled[env] = (uint8_t)shuffled; }
            b.store_board();
            b.store_cursors();
            merge_status(b, p);
            if (lane == 0) {
                p.reward[env] = result;
                p.new_specials[env] = b.n_new;
                p.activated[env] = b.n_act;
            }
            return;
        }
    }
    switch (op) {
        case OP_GRAVITY: b.gravity(&result); break;
        case OP_REFILL: {
            // refill() expects the post-gravity layout; a general board is handled row by row instead:
            // count empties row-major and assign draws in that order (ref :231-241)
            int total = 0;
            for (int r = 0; r < p.R; ++r) {
                const bool em = lane < C && b.col[r * C + lane] == 0 && b.typ[r * C + lane] == 0;
                total += __popc(b.ballot(em));
            }
            for (int ps = 0; ps < total; ps += Cfg<L>::NW) {
                const int nw = min(Cfg<L>::NW, total - ps);
                if (!p.use_inj) b.fill_words(0u, b.dcur + (uint64_t)ps, nw);
                int base = 0;
                for (int r = 0; r < p.R; ++r) {
                    const bool em = lane < C && b.col[r * C + lane] == 0 && b.typ[r * C + lane] == 0;
                    const unsigned m = b.ballot(em);
                    if (em) {
                        const int rank = base + __popc(m & b.lt_mask());
                        if (rank >= ps && rank < ps + nw) {
                            b.col[r * C + lane] = (int8_t)(p.use_inj ? b.injected_colour(rank)
                                                                     : 1 + (int)__umulhi(b.s.wbuf[rank - ps], (uint32_t)p.K));
                            b.typ[r * C + lane] = 1;
                        }
                    }
                    base += __popc(m);
                }
                b.sync();
            }
            b.dcur += (uint64_t)total;
            result = total;
            break;
        }
        case OP_RESOLVE_ROUND: result = b.resolve_round(); break;
        case OP_ACTIVATE: b.activate(a0 * C + a1, a2, a3 == 0); break;
        case OP_COMBINE: b.combination(a0 * C + a1, a2 * C + a3); break;
        case OP_MOVE: {
            int reward = 0, is_comb = 0, shuffled = 0;
            unsigned effv = 0u, effh = 0u;
            const int i1 = a0 * C + a1, i2 = a2 * C + a3;
            b.n_new = 0; b.n_act = 0;
            bool eff = false;
            if (lane == 0) eff = effective_literal(b.col, b.typ, p.R, C, i1, i2);
            eff = b.shfl((int)eff, 0) != 0;
            if (eff) {
                b.move_core(i1, i2, reward, is_comb);
                shuffled = b.playability(true, false, effv, effh);
            }
            if (lane == 0) { p.is_comb[env] = (uint8_t)is_comb; p.shuffled[env] = (uint8_t)shuffled; }
            result = reward;
            break;
        }
        case OP_EFFECTIVE: {
            bool eff = false;
            if (lane == 0) eff = effective_literal(b.col, b.typ, p.R, C, a0 * C + a1, a2 * C + a3);
            result = b.shfl((int)eff, 0);
            break;
        }
        case OP_GENERATE: {
            unsigned v, h;
            const int ep = p.episode[env] + 1;
            b.sync();
            if (p.flags & FLAG_CONSTRUCTIVE_RESET) b.generate_constructive((uint32_t)ep, v, h);
            else {
                b.begin_generate((uint32_t)ep);
                b.playability(false, true, v, h);
                b.end_generate();
            }
            if (lane == 0) p.episode[env] = ep;
            break;
        }
        case OP_SHUFFLE: b.shuffle(); break;
        case OP_COUNT_LINES:
        case OP_LINES: {
            const typename Board<L>::Scan sc = b.scan_lines(p.R - 1, false);
            result = sc.rstar < 0 ? 0 : b.build_line_table(sc);
            if (op == OP_LINES && p.dbg_out) {               // the line table in the entry format of RBoard::line_info
                uint32_t* out = p.dbg_out + (size_t)env * LINES_WORDS;
                if (lane == 0) out[0] = (uint32_t)min(result, 32);
                for (int i = lane; i < min(result, 32); i += L)  {
                    out[1 + 2 * i] = (b.s.line_key[i] & 0xffffu) | ((uint32_t)b.s.line_kind[i] << 16) | ((uint32_t)b.s.line_idx[i] << 17) |
                                     ((uint32_t)(b.s.line_colour[i] & 7u) << 22);
                    out[2 + 2 * i] = b.s.line_mask[i];
                }
            }
            break;
        }
        default: break;
    }
    b.store_board();
    b.store_cursors();
    merge_status(b, p);
    if (lane == 0) {
        p.reward[env] = result;
        p.new_specials[env] = b.n_new;
        p.activated[env] = b.n_act;
    }
}

// OneHotWrapper._one_hot_encode_board (ref wrappers.py:54-69).  One thread per cell: two coalesced byte loads, then one
// store per output plane -- a warp writes 32 consecutive elements of a plane row per store.  HBM-bound: 2P bytes in,
// (K + S) P elements out per env.
template <typename OUT> __global__ void __launch_bounds__(256) k_onehot(const Params p, OUT* __restrict__ out, int planes) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (long long)p.N * p.P) return;
    const int env = (int)(q / p.P), cell = (int)(q - (long long)env * p.P);
    const int8_t* bd = p.board + (size_t)env * 2 * p.P;
    const int colour = bd[cell], type = bd[p.P + cell];
    OUT* o = out + (size_t)env * planes * p.P + cell;
#pragma unroll 4
    for (int k = 0; k < p.K; ++k) o[(size_t)k * p.P] = (OUT)(colour == k + 1);
    o += (size_t)p.K * p.P;
    // enabled specials in the order cookie(-1), v(2), h(3), bomb(4) (ref wrappers.py:40-46)
    if (p.specials & SP_COOKIE) { *o = (OUT)(type == -1); o += p.P; }
    if (p.specials & SP_VLASER) { *o = (OUT)(type == 2); o += p.P; }
    if (p.specials & SP_HLASER) { *o = (OUT)(type == 3); o += p.P; }
    if (p.specials & SP_BOMB) { *o = (OUT)(type == 4); o += p.P; }
}

// legal-move mask as bits for the host-buffer path: out[env][b] bit j = mask[env][8b + j]  (one thread per output byte)
__global__ void __launch_bounds__(256) k_pack_mask(const uint8_t* __restrict__ mask, uint8_t* __restrict__ out, int n_envs, int A,
                                                   int bytes_per_env) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_envs * bytes_per_env) return;
    const int env = (int)(i / bytes_per_env), b = (int)(i - (long long)env * bytes_per_env);
    const uint8_t* m = mask + (size_t)env * A + 8 * b;
    unsigned v = 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (8 * b + j < A) v |= (unsigned)(m[j] != 0) << j;
    out[i] = (uint8_t)v;
}

// closes the batch of pool requests collected so far under p.pool_tag (tmg_join: no step follows that could do it)
__global__ void k_commit_batch(const __grid_constant__ Params p) {
    const uint32_t tail = ctl_read(&p.ctl[CTL_REQ_TAIL]), prev = p.ctl[CTL_REQ_PREV];
    const int slot = p.pool_tag % PG_RING;
    p.ctl[CTL_PG_RANGE + 2 * slot] = prev;
    p.ctl[CTL_PG_RANGE + 2 * slot + 1] = tail;
    p.ctl[CTL_PG_HEAD + slot] = 0u;
    p.ctl[CTL_REQ_PREV] = tail;
}

// the packed form of every board (full refresh of a bound packed host mirror): one thread per 4 cells
__global__ void __launch_bounds__(256) k_pack_boards(const int8_t* __restrict__ board, uint8_t* __restrict__ out, int n_envs, int P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_envs * P) return;
    const int env = (int)(i / P), cell = (int)(i - (long long)env * P);
    const int8_t* b = board + (size_t)env * 2 * P;
    out[i] = (uint8_t)((b[cell] & 15) | ((b[P + cell] & 7) << 4));
}

__global__ void k_clear_status(uint32_t* st, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st[i] = 0u;
}

}  // namespace tmg
