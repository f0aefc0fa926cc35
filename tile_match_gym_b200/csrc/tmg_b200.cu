// tmg_b200.cu -- host side of the C ABI declared in include/tmg_b200.h (sm_100a only, no CPU fallback).
#include "tmg_b200.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "tmg_device.cuh"

using namespace tmg;

struct tmg_env {
    tmg_config cfg;
    Params p;
    int L;          // lanes per board: 8, 10, 16 or 32
    int planes;     // one-hot planes
    int32_t* actions_dev;  // staging for tmg_step_host
    uint8_t* mask_bits_dev;
    uint8_t* board_packed_dev;   // staging for the full refresh of a packed board mirror
    void* base;     // one allocation backs every buffer
    size_t bytes;
    // board pool: k_pregen runs on ONE side stream (refills are served strictly in request order, so the pool entry of an
    // env is only ever written by one launch at a time), off the step path.  Requests name their board ({env, board
    // number}), and a consumer takes an entry only if its number matches, so correctness never depends on when a refill
    // runs; the waits below only keep refills ahead of the steps that will want them and bound the launches in flight.
    static constexpr int RING = tmg::PG_RING;
    cudaStream_t side;
    cudaEvent_t ev_step, ev_pregen[RING];
    long long pregen_count;          // number of k_pregen launches so far (= tag of the next one)
    long long pregen_step[RING];     // step_count when launch `tag` was issued, at tag % RING
    long long waited_upto;           // every launch with tag <= this has been waited for by waited_stream
    cudaStream_t waited_stream;
    int pregen_every;                // a k_pregen launch serves the requests of this many tmg_step calls
    int steps_since_pregen;
    long long step_count;     // number of tmg_step calls so far (parity selects the work-list counters)
    int persistent_blocks;    // resident-block slots of the device for k_work / k_pregen (persistent groups)
    int pregen_grid_cap;      // diagnostics: cap on the blocks of a k_pregen launch
    bool pregen_one_shot;     // diagnostics: short-lived k_pregen blocks (one board per group), see Params::pregen_one_shot
    bool pregen;              // pool in use (philox refill, not disabled by flag)
    // host mirror (tmg_host_bind): the caller's page-locked arrays; p.h_* are their device-visible aliases
    int8_t* hm_board;
    uint8_t *hm_mask, *hm_mask_bits, *hm_terminated, *hm_board_packed;
    int32_t *hm_reward, *hm_moves_left;
    bool hm_bound;                  // any array bound: tmg_step_host then also reads page-locked actions in place
};

namespace {

int vec_width(size_t stride_bytes) {
    int w = 16;
    while (w > 1 && (stride_bytes % (size_t)w) != 0) w >>= 1;
    return w;
}
int ptr_vec_width(const void* ptr, int w) {
    while (w > 1 && (reinterpret_cast<uintptr_t>(ptr) % (uintptr_t)w) != 0) w >>= 1;
    return w;
}
size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int check_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return TMG_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return TMG_ERR_INVALID_ARG;
    return TMG_OK;
}

// (lanes per board, compile-time rows, compile-time cols); RT = CT = 0 is the runtime-shape kernel
template <int L_, int R_, int C_> struct Shape { static constexpr int L = L_, R = R_, C = C_; };
template <typename F> int launch_by_shape(const tmg_env* e, F&& f) {
    const int R = e->p.R, C = e->p.C;
    static const bool generic_only = getenv("TMG_B200_GENERIC_SHAPES") != nullptr;   // diagnostics: skip the fixed-shape kernels
    if (e->L == 32 && !generic_only) {
        if (R == 10 && C == 10) return f(Shape<32, 10, 10>());
        if (R == 9 && C == 9) return f(Shape<32, 9, 9>());
        if (R == 32 && C == 32) return f(Shape<32, 32, 32>());
    }
#ifdef TMG_SUBWARP_GROUPS   // 2-4 boards per warp: measured slower on B200 at every batch size, kept for experiments only
    switch (e->L) {
        case 8: return f(Shape<8, 0, 0>());
        case 10: return f(Shape<10, 0, 0>());
        case 16: return f(Shape<16, 0, 0>());
        default: break;
    }
#endif
    return f(Shape<32, 0, 0>());
}
template <typename F> int launch_by_lanes(const tmg_env* e, F&& f) {
#ifdef TMG_SUBWARP_GROUPS
    switch (e->L) {
        case 8: return f(std::integral_constant<int, 8>());
        case 10: return f(std::integral_constant<int, 10>());
        case 16: return f(std::integral_constant<int, 16>());
        default: break;
    }
#endif
    (void)e;
    return f(std::integral_constant<int, 32>());
}

int last_error() { return cudaGetLastError() == cudaSuccess ? TMG_OK : TMG_ERR_CUDA; }

template <int L> int grid_for(int n) { return (n + Cfg<L>::GPB - 1) / Cfg<L>::GPB; }

}  // namespace

// `st` waits until every k_pregen launch with tag <= upto has finished (one in-order side stream: one event is enough)
static bool wait_pregen(tmg_env* e, cudaStream_t st, long long upto) {
    if (!e->pregen) return true;
    if (upto >= e->pregen_count) upto = e->pregen_count - 1;
    if (st != e->waited_stream) { e->waited_stream = st; e->waited_upto = -1; }   // another caller stream has waited for nothing yet
    if (upto <= e->waited_upto) return true;
    if (e->pregen_count - upto < tmg_env::RING) {    // (an older slot was reused by a later launch: waiting for that one is stronger)
        if (cudaStreamWaitEvent(st, e->ev_pregen[upto % tmg_env::RING], 0) != cudaSuccess) return false;
    } else if (cudaStreamWaitEvent(st, e->ev_pregen[(e->pregen_count - 1) % tmg_env::RING], 0) != cudaSuccess) return false;
    e->waited_upto = upto;
    return true;
}
static bool join_pregen(tmg_env* e, cudaStream_t st) { return wait_pregen(e, st, e->pregen_count - 1); }
// Before a kernel on `st` that may consume pool entries: wait for the refills whose boards can be due by now (their
// requests are at least num_moves - pregen_every steps old) and keep at most 3 launches in flight.
static bool wait_due_pregen(tmg_env* e, cudaStream_t st) {
    if (!e->pregen) return true;
    long long upto = e->pregen_count - 4;
    const long long slack = e->p.num_moves - e->pregen_every;
    for (long long j = e->pregen_count - 1; j > upto && j >= 0; --j)
        if (e->step_count - e->pregen_step[j % tmg_env::RING] >= slack) { upto = j; break; }
    return upto < 0 || wait_pregen(e, st, upto);
}
// Does the launch about to be issued on the step stream close a batch of pool requests (so that k_pregen follows it)?
static bool closes_batch(tmg_env* e, bool is_step) {
    if (!e->pregen) return false;
    if (!is_step) return true;
    return e->steps_since_pregen + 1 >= e->pregen_every;
}
// after a kernel on `st` that closed the batch tagged e->pregen_count: serve it on the side stream
static int launch_pregen(tmg_env* e, cudaStream_t st) {
    if (!e->pregen) return TMG_OK;
    const long long tag = e->pregen_count;
    if (cudaEventRecord(e->ev_step, st) != cudaSuccess) return TMG_ERR_CUDA;
    if (cudaStreamWaitEvent(e->side, e->ev_step, 0) != cudaSuccess) return TMG_ERR_CUDA;
    Params p = e->p;
    p.pool_tag = (int)(tag & 0x7fffffff);
    p.pregen_one_shot = e->pregen_one_shot;
    const int rc = launch_by_shape(e, [&](auto shape) {
        typedef decltype(shape) S;
        constexpr int L = S::L;
        int grid = grid_for<L>(p.N);
        if (!e->pregen_one_shot) {
            if (grid > e->persistent_blocks) grid = e->persistent_blocks;
            if (grid > e->pregen_grid_cap) grid = e->pregen_grid_cap;
        }
        k_pregen<L, S::R, S::C><<<grid, Cfg<L>::THREADS, Cfg<L>::GPB * sizeof(GroupSmem<L>), e->side>>>(p);
        return last_error();
    });
    if (rc != TMG_OK) return rc;
    if (cudaEventRecord(e->ev_pregen[tag % tmg_env::RING], e->side) != cudaSuccess) return TMG_ERR_CUDA;
    e->pregen_step[tag % tmg_env::RING] = e->step_count;
    ++e->pregen_count;
    e->steps_since_pregen = 0;
    return TMG_OK;
}

extern "C" {

int tmg_abi_version(void) { return TMG_ABI_VERSION; }

#ifndef TMG_BUILD_ID_STR
#define TMG_BUILD_ID_STR "unidentified-build"
#endif
// hash of the sources this library was compiled from (tile_match_gym_b200/_native.py compares it with the tree)
const char* tmg_build_id(void) {
    static const char id[] = "TMG_BUILD_ID=" TMG_BUILD_ID_STR;
    return id + 13;
}

const char* tmg_error_string(int code) {
    switch (code) {
        case TMG_OK: return "ok";
        case TMG_ERR_INVALID_ARG: return "invalid argument";
        case TMG_ERR_UNSUPPORTED_SHAPE: return "unsupported board shape or colour count";
        case TMG_ERR_CUDA: return "CUDA error";
        case TMG_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
        case TMG_ERR_OOM: return "out of device memory";
        case TMG_ERR_STATE: return "call not valid in this state";
        default: return "unknown error";
    }
}

const char* tmg_status_string(uint32_t status) {
    static thread_local char buf[256];
    static const char* names[] = {"bad_action", "needs_reset", "draws_exhausted", "reset_cap",
                                  "line_overflow", "dfs_overflow", "invalid_board", "internal"};
    buf[0] = 0;
    for (int i = 0; i < 8; ++i)
        if (status & (1u << i)) {
            if (buf[0]) strncat(buf, "|", sizeof(buf) - strlen(buf) - 1);
            strncat(buf, names[i], sizeof(buf) - strlen(buf) - 1);
        }
    if (!buf[0]) strncpy(buf, "ok", sizeof(buf));
    return buf;
}

int tmg_num_actions(int32_t R, int32_t C) { return 2 * R * C - R - C; }

int tmg_onehot_planes(int32_t K, uint32_t specials) {
    return K + !!(specials & TMG_SP_COOKIE) + !!(specials & TMG_SP_VERTICAL_LASER) +
           !!(specials & TMG_SP_HORIZONTAL_LASER) + !!(specials & TMG_SP_BOMB);
}

int tmg_action_to_coords(int32_t R, int32_t C, int32_t a, int32_t out[4]) {
    if (R < 1 || C < 1 || a < 0 || a >= tmg_num_actions(R, C)) return TMG_ERR_INVALID_ARG;
    if (a < C * (R - 1)) { out[0] = a / C; out[1] = a % C; out[2] = out[0] + 1; out[3] = out[1]; }
    else { const int j = a - C * (R - 1); out[0] = j / (C - 1); out[1] = j % (C - 1); out[2] = out[0]; out[3] = out[1] + 1; }
    return TMG_OK;
}

int tmg_create(const tmg_config* cfg, tmg_env** out) {
    if (!cfg || !out || cfg->struct_size != sizeof(tmg_config)) return TMG_ERR_INVALID_ARG;
    *out = nullptr;
    const int R = cfg->num_rows, C = cfg->num_cols, K = cfg->num_colours, N = cfg->num_envs;
    if (N < 1 || cfg->num_moves < 1) return TMG_ERR_INVALID_ARG;
    if (R < 1 || C < 2 || R > TMG_MAX_ROWS || C > TMG_MAX_COLS || K < 1 || K > TMG_MAX_COLOURS || R * C < 2)
        return TMG_ERR_UNSUPPORTED_SHAPE;
    if (cfg->autoreset < 0 || cfg->autoreset > 2 || cfg->refill_mode < 0 || cfg->refill_mode > 1) return TMG_ERR_INVALID_ARG;
    int rc = check_device(cfg->device);
    if (rc != TMG_OK) return rc;
    if (cudaSetDevice(cfg->device) != cudaSuccess) return TMG_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return TMG_ERR_CUDA;
    if (prop.major != 10) return TMG_ERR_NO_DEVICE;  // built for sm_100a only

    tmg_env* e = new (std::nothrow) tmg_env();
    if (!e) return TMG_ERR_OOM;
    e->cfg = *cfg;
    // Lanes per board: ONE BOARD PER WARP (lane c owns column c, the other lanes help with Philox blocks and copies).
    // Sub-warp groups (2-4 boards per warp) are supported by the same code and were the first design, but measured
    // slower on B200 at every batch size up to 262 144 envs (65 536 envs: 232 M vs 253 M steps/s; 1 024 envs: 8.6 M vs
    // 11.5 M): the groups of a warp diverge into separate instruction streams anyway, their collectives need a
    // runtime membership check, and the step is bound by instruction supply, not by lanes.  A build with
    // -DTMG_SUBWARP_GROUPS instantiates them (TMG_B200_LANES=8|10|16 then selects one); the product build does not.
    e->L = 32;
#ifdef TMG_SUBWARP_GROUPS
    if (const char* lanes = getenv("TMG_B200_LANES")) {
        const int l = atoi(lanes);
        if ((l == 8 && C <= 8 && R <= 16) || (l == 10 && C <= 10 && R <= 16) || (l == 16 && C <= 16 && R <= 16)) e->L = l;
    }
#endif
    e->planes = tmg_onehot_planes(K, cfg->specials);
    Params& p = e->p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.R = R; p.C = C; p.K = K; p.P = R * C; p.A = tmg_num_actions(R, C);
    p.num_moves = cfg->num_moves;
    p.specials = cfg->specials;
    p.autoreset = cfg->autoreset;
    p.use_inj = cfg->refill_mode == TMG_REFILL_INJECTED;
    p.flags = cfg->flags;
    p.max_iters = cfg->max_reset_iters > 0 ? cfg->max_reset_iters : 16384;
    p.key0 = (uint32_t)cfg->seed;
    p.key1 = (uint32_t)(cfg->seed >> 32);
    p.env_id_offset = cfg->env_id_offset;
    p.board_vecw = vec_width((size_t)2 * p.P);
    p.mask_vecw = vec_width((size_t)p.A);
    p.init_vecw = p.board_vecw;

    // one slab, every array 256-byte aligned
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_board = take((size_t)N * 2 * p.P), o_timer = take((size_t)N * 4), o_dc = take((size_t)N * 8),
                 o_sc = take((size_t)N * 8), o_rew = take((size_t)N * 4), o_term = take(N), o_comb = take(N),
                 o_new = take((size_t)N * 4), o_act = take((size_t)N * 4), o_shuf = take(N),
                 o_mask = take((size_t)N * p.A), o_left = take((size_t)N * 4), o_stat = take((size_t)N * 4),
                 o_actions = take((size_t)N * 4), o_ep = take((size_t)N * 4), o_pool_ep = take((size_t)N * 4),
                 o_pool_board = take((size_t)N * 2 * p.P), o_pool_mask = take((size_t)N * p.A),
                 o_pool_status = take((size_t)N * 4), o_mask_bits = take((size_t)N * ((p.A + 7) / 8)),
                 o_ctl = take((size_t)CTL_WORDS * 4), o_items = take((size_t)N * sizeof(uint2)),
                 o_nsp = take((size_t)N), o_bpk = take((size_t)N * p.P);
    size_t req_cap = 1;   // requests of the launches in flight plus the batch being collected; an overwritten entry is harmless
    while (req_cap < (size_t)8 * N) req_cap <<= 1;
    const size_t o_ring = take(req_cap * sizeof(uint2));
    e->bytes = off;
    if (cudaMalloc(&e->base, e->bytes) != cudaSuccess) { cudaGetLastError(); delete e; return TMG_ERR_OOM; }
    if (cudaMemset(e->base, 0, e->bytes) != cudaSuccess) { cudaFree(e->base); delete e; return TMG_ERR_CUDA; }
    char* b = static_cast<char*>(e->base);
    p.board = reinterpret_cast<int8_t*>(b + o_board);
    p.timer = reinterpret_cast<int32_t*>(b + o_timer);
    p.draw_cursor = reinterpret_cast<uint64_t*>(b + o_dc);
    p.shuffle_cursor = reinterpret_cast<uint64_t*>(b + o_sc);
    p.reward = reinterpret_cast<int32_t*>(b + o_rew);
    p.terminated = reinterpret_cast<uint8_t*>(b + o_term);
    p.is_comb = reinterpret_cast<uint8_t*>(b + o_comb);
    p.new_specials = reinterpret_cast<int32_t*>(b + o_new);
    p.activated = reinterpret_cast<int32_t*>(b + o_act);
    p.shuffled = reinterpret_cast<uint8_t*>(b + o_shuf);
    p.mask = reinterpret_cast<uint8_t*>(b + o_mask);
    p.moves_left = reinterpret_cast<int32_t*>(b + o_left);
    p.status = reinterpret_cast<uint32_t*>(b + o_stat);
    e->actions_dev = reinterpret_cast<int32_t*>(b + o_actions);
    p.episode = reinterpret_cast<int32_t*>(b + o_ep);
    p.pool_episode = reinterpret_cast<int32_t*>(b + o_pool_ep);
    p.pool_board = reinterpret_cast<int8_t*>(b + o_pool_board);
    p.pool_mask = reinterpret_cast<uint8_t*>(b + o_pool_mask);
    p.pool_status = reinterpret_cast<uint32_t*>(b + o_pool_status);
    p.ctl = reinterpret_cast<uint32_t*>(b + o_ctl);
    p.wl_items = reinterpret_cast<uint2*>(b + o_items);
    p.req_mask = (uint32_t)(req_cap - 1);
    p.n_special = reinterpret_cast<uint8_t*>(b + o_nsp);
    e->mask_bits_dev = reinterpret_cast<uint8_t*>(b + o_mask_bits);
    e->board_packed_dev = reinterpret_cast<uint8_t*>(b + o_bpk);
    e->pregen = !p.use_inj && !(cfg->flags & TMG_FLAG_NO_PREGEN) && cfg->autoreset != TMG_AUTORESET_DISABLED;
    e->pregen_count = 0;
    e->waited_stream = nullptr;
    e->waited_upto = -1;
    e->step_count = 0;
    e->steps_since_pregen = 0;
    // One k_pregen launch per `pregen_every` steps: a board requested at an episode end is wanted num_moves steps later,
    // so the requests of a few steps are served together (fewer, fuller launches when the episode phases differ).
    // Measured (B200, config 2, staggered phases; bench 120/30, 20/5 steps at 65 536 envs, 60/10 at 1 M): every 2 steps
    // 428 / 423 M, 3: 443 / 431 / 1021 M, 4: 444 / 447 / 1021 M, 5: 445 / 450 M, 7: 443 / 418 / 980 M, 12: 447 / 406 M -- the
    // refill left open at the end of a run is drained inside its timed total, which is what a long interval costs a short run.
    e->pregen_every = cfg->num_moves >= 8 ? (cfg->num_moves / 4 < 4 ? cfg->num_moves / 4 : 4) : 1;
    if (const char* pe = getenv("TMG_B200_PREGEN_EVERY")) { if (atoi(pe) > 0) e->pregen_every = atoi(pe); }
    for (int i = 0; i < tmg_env::RING; ++i) e->pregen_step[i] = 0;
    e->hm_board = nullptr; e->hm_mask = nullptr; e->hm_mask_bits = nullptr; e->hm_terminated = nullptr;
    e->hm_reward = nullptr; e->hm_moves_left = nullptr; e->hm_bound = false; e->hm_board_packed = nullptr;
    p.req_ring = e->pregen ? reinterpret_cast<uint2*>(b + o_ring) : nullptr;
    e->persistent_blocks = prop.multiProcessorCount * TMG_STEP_MIN_BLOCKS;
    {
        const char* ppsm = getenv("TMG_B200_BLOCKS_PER_SM");   // tuning knob: persistent blocks per SM
        if (ppsm && atoi(ppsm) > 0) e->persistent_blocks = prop.multiProcessorCount * atoi(ppsm);
    }
    {
        // tuning knob: cap the refill kernel to this many blocks per SM.  Measured on B200 (65 536 envs): uncapped is
        // best overall -- a capped refill delays fewer steps but runs longer beside more of them.
        const char* capenv = getenv("TMG_B200_PREGEN_BLOCKS_PER_SM");
        const int per_sm = capenv ? atoi(capenv) : 0;
        e->pregen_grid_cap = per_sm > 0 ? prop.multiProcessorCount * per_sm : 0x7fffffff;
        e->pregen_one_shot = getenv("TMG_B200_PREGEN_ONE_SHOT") != nullptr;
    }
    e->side = nullptr;
    bool ok = cudaMemset(p.episode, 0xff, (size_t)N * 4) == cudaSuccess &&       // -1: no board generated yet
              cudaMemset(p.pool_episode, 0x80, (size_t)N * 4) == cudaSuccess;    // never equal to a real episode
    if (e->pregen) {
        ok = ok && cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&e->ev_step, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < tmg_env::RING; ++i) ok = ok && cudaEventCreateWithFlags(&e->ev_pregen[i], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) { cudaFree(e->base); delete e; return TMG_ERR_CUDA; }
    // timer = -1: "reset has never been called" (tile_match_env.py:75)
    if (cudaMemset(p.timer, 0xff, (size_t)N * 4) != cudaSuccess) { cudaFree(e->base); delete e; return TMG_ERR_CUDA; }
    if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(e->base); delete e; return TMG_ERR_CUDA; }
    *out = e;
    return TMG_OK;
}

int tmg_destroy(tmg_env* e) {
    if (!e) return TMG_ERR_INVALID_ARG;
    cudaSetDevice(e->cfg.device);
    if (e->pregen) {
        cudaStreamSynchronize(e->side);
        cudaEventDestroy(e->ev_step);
        for (int i = 0; i < tmg_env::RING; ++i) cudaEventDestroy(e->ev_pregen[i]);
        cudaStreamDestroy(e->side);
    }
    cudaFree(e->base);
    delete e;
    return TMG_OK;
}

int tmg_get_buffers(tmg_env* e, tmg_buffers* out) {
    if (!e || !out) return TMG_ERR_INVALID_ARG;
    const Params& p = e->p;
    out->board = p.board; out->timer = p.timer; out->draw_cursor = p.draw_cursor; out->shuffle_cursor = p.shuffle_cursor;
    out->reward = p.reward; out->terminated = p.terminated; out->is_combination_match = p.is_comb;
    out->num_new_specials = p.new_specials; out->num_specials_activated = p.activated; out->shuffled = p.shuffled;
    out->mask = p.mask; out->num_moves_left = p.moves_left; out->status = p.status; out->episode = p.episode;
    return TMG_OK;
}

int tmg_set_injected_draws(tmg_env* e, const uint8_t* draws_dev, int64_t per_env_len) {
    if (!e || per_env_len < 0 || (!draws_dev && per_env_len > 0)) return TMG_ERR_INVALID_ARG;
    if (!e->p.use_inj) return TMG_ERR_STATE;
    e->p.inj = draws_dev;
    e->p.inj_len = per_env_len;
    return TMG_OK;
}

// full copy of the bound mirror arrays (after calls that rewrite boards / masks wholesale)
static int refresh_mirror(tmg_env* e, cudaStream_t st) {
    const Params& p = e->p;
    const size_t N = (size_t)p.N;
    bool ok = true;
    if (e->hm_board) ok &= cudaMemcpyAsync(e->hm_board, p.board, N * 2 * p.P, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (e->hm_board_packed) {
        const long long total = (long long)N * p.P;
        k_pack_boards<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p.board, e->board_packed_dev, (int)N, p.P);
        ok &= cudaGetLastError() == cudaSuccess;
        ok &= cudaMemcpyAsync(e->hm_board_packed, e->board_packed_dev, (size_t)total, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    }
    if (e->hm_mask) ok &= cudaMemcpyAsync(e->hm_mask, p.mask, N * p.A, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (e->hm_mask_bits) {
        const int bpe = (p.A + 7) / 8;
        const long long total = (long long)N * bpe;
        k_pack_mask<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p.mask, e->mask_bits_dev, (int)N, p.A, bpe);
        ok &= cudaGetLastError() == cudaSuccess;
        ok &= cudaMemcpyAsync(e->hm_mask_bits, e->mask_bits_dev, (size_t)total, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    }
    if (e->hm_reward) ok &= cudaMemcpyAsync(e->hm_reward, p.reward, N * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (e->hm_terminated) ok &= cudaMemcpyAsync(e->hm_terminated, p.terminated, N, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (e->hm_moves_left) ok &= cudaMemcpyAsync(e->hm_moves_left, p.moves_left, N * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    return ok ? TMG_OK : TMG_ERR_CUDA;
}

int tmg_host_bind(tmg_env* e, const tmg_host_io* io, void* stream) {
    if (!e) return TMG_ERR_INVALID_ARG;
    if (cudaSetDevice(e->cfg.device) != cudaSuccess) return TMG_ERR_CUDA;
    static const tmg_host_io none = {};
    if (!io) io = &none;
    void* host[7] = {io->board, io->mask, io->mask_bits, io->reward, io->terminated, io->num_moves_left, io->board_packed};
    void* dev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (io->board_packed && e->p.K > 15) return TMG_ERR_INVALID_ARG;   // a colour must fit the low nibble
    bool any = false;
    for (int i = 0; i < 7; ++i) {
        if (!host[i]) continue;
        if (reinterpret_cast<uintptr_t>(host[i]) % 16 != 0) return TMG_ERR_INVALID_ARG;
        // page-locked (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory) memory only: the kernels access it directly
        if (cudaHostGetDevicePointer(&dev[i], host[i], 0) != cudaSuccess) { cudaGetLastError(); return TMG_ERR_INVALID_ARG; }
        any = true;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    e->hm_bound = any;
    e->hm_board = io->board; e->hm_mask = io->mask; e->hm_mask_bits = io->mask_bits;
    e->hm_reward = io->reward; e->hm_terminated = io->terminated; e->hm_moves_left = io->num_moves_left;
    e->hm_board_packed = io->board_packed;
    e->p.h_board = static_cast<int8_t*>(dev[0]);
    e->p.h_mask = static_cast<uint8_t*>(dev[1]);
    e->p.h_mask_bits = static_cast<uint8_t*>(dev[2]);
    e->p.h_reward = static_cast<int32_t*>(dev[3]);
    e->p.h_terminated = static_cast<uint8_t*>(dev[4]);
    e->p.h_moves_left = static_cast<int32_t*>(dev[5]);
    e->p.h_board_packed = static_cast<uint8_t*>(dev[6]);
    return refresh_mirror(e, st);
}

int tmg_reset(tmg_env* e, const uint8_t* reset_mask_dev, const int8_t* init_boards_dev, void* stream) {
    if (!e) return TMG_ERR_INVALID_ARG;
    if (cudaSetDevice(e->cfg.device) != cudaSuccess) return TMG_ERR_CUDA;
    Params p = e->p;
    p.reset_mask = reset_mask_dev;
    p.init_boards = init_boards_dev;
    p.init_vecw = init_boards_dev ? ptr_vec_width(init_boards_dev, p.board_vecw) : p.board_vecw;
    p.pool_tag = (int)(e->pregen_count & 0x7fffffff);
    p.commit_pregen = closes_batch(e, false);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!join_pregen(e, st)) return TMG_ERR_CUDA;
    const int rc = launch_by_shape(e, [&](auto shape) {
        typedef decltype(shape) S;
        constexpr int L = S::L;
        k_reset<L, S::R, S::C><<<grid_for<L>(p.N), Cfg<L>::THREADS, Cfg<L>::GPB * sizeof(GroupSmem<L>), st>>>(p);
        return last_error();
    });
    if (rc != TMG_OK) return rc;
    const int rc2 = refresh_mirror(e, st);
    if (rc2 != TMG_OK) return rc2;
    return launch_pregen(e, st);
}

int tmg_step(tmg_env* e, const int32_t* actions_dev, void* stream) {
    if (!e || !actions_dev) return TMG_ERR_INVALID_ARG;
    if (cudaSetDevice(e->cfg.device) != cudaSuccess) return TMG_ERR_CUDA;
    Params p = e->p;
    p.actions = actions_dev;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    p.pool_tag = (int)(e->pregen_count & 0x7fffffff);
    p.seq = (int)(e->step_count++ & 1);
    const bool close_batch = closes_batch(e, true);
    p.commit_pregen = close_batch;
    if (!wait_due_pregen(e, st)) return TMG_ERR_CUDA;
    k_gate<<<(p.N + 128 * GATE_EPT - 1) / (128 * GATE_EPT), 128, 0, st>>>(p);
    if (last_error() != TMG_OK) return TMG_ERR_CUDA;
    const int rc = launch_by_shape(e, [&](auto shape) {
        typedef decltype(shape) S;
        constexpr int L = S::L;
        int grid = grid_for<L>(p.N);
        if (grid > e->persistent_blocks) grid = e->persistent_blocks;
        const size_t smem = Cfg<L>::GPB * sizeof(GroupSmem<L>);
        const bool mir = e->hm_bound;   // a host mirror is bound: the instantiation that writes it through inline
        if constexpr (UsesRB<L, S::R>::maybe) {   // the register-resident engine where it applies (tmg_rb.cuh)
            if (rb_supported(L, p.R, p.K, p.flags, p.use_inj)) {
                if (mir) k_work<L, S::R, S::C, true, true><<<grid, Cfg<L>::THREADS, smem, st>>>(p);
                else k_work<L, S::R, S::C, true, false><<<grid, Cfg<L>::THREADS, smem, st>>>(p);
                return last_error();
            }
        }
        if (mir) k_work<L, S::R, S::C, false, true><<<grid, Cfg<L>::THREADS, smem, st>>>(p);
        else k_work<L, S::R, S::C, false, false><<<grid, Cfg<L>::THREADS, smem, st>>>(p);
        return last_error();
    });
    if (rc != TMG_OK) return rc;
    ++e->steps_since_pregen;
    return close_batch ? launch_pregen(e, st) : TMG_OK;
}

static int rollout(tmg_env* e, int policy, const int32_t* actions_dev, int32_t num_steps, int32_t* actions_out_dev,
                   int32_t* rewards_dev, uint8_t* terminated_dev, void* stream) {
    if (!e || num_steps < 0) return TMG_ERR_INVALID_ARG;
    if (num_steps == 0) return TMG_OK;
    if (cudaSetDevice(e->cfg.device) != cudaSuccess) return TMG_ERR_CUDA;
    Params p = e->p;
    p.actions = actions_dev;
    p.policy = policy;
    p.T = num_steps;
    p.ro_actions = actions_out_dev;
    p.ro_reward = rewards_dev;
    p.ro_terminated = terminated_dev;
    p.pool_tag = (int)(e->pregen_count & 0x7fffffff);
    p.commit_pregen = closes_batch(e, false);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // No wait for the recent pool refills: k_rollout takes a pool entry only if its board number says it is the one it
    // wants and complete (published last by k_pregen) and generates the board itself otherwise -- the same bytes either
    // way.  Only the number of refill launches in flight is bounded (range / event slots are rings of PG_RING).
    if (e->pregen && !wait_pregen(e, st, e->pregen_count - 4)) return TMG_ERR_CUDA;
    const int rc = launch_by_shape(e, [&](auto shape) {
        typedef decltype(shape) S;
        constexpr int L = S::L;
        int grid = grid_for<L>(p.N);
        if (grid > e->persistent_blocks) grid = e->persistent_blocks;
        if constexpr (UsesRB<L, S::R>::maybe) {   // the register-resident engine where it applies (tmg_rb.cuh)
            if (rb_supported(L, p.R, p.K, p.flags, p.use_inj)) {
                k_rollout<L, S::R, S::C, true><<<grid, Cfg<L>::THREADS, Cfg<L>::GPB * sizeof(GroupSmem<L>), st>>>(p);
                return last_error();
            }
        }
        k_rollout<L, S::R, S::C, false><<<grid, Cfg<L>::THREADS, Cfg<L>::GPB * sizeof(GroupSmem<L>), st>>>(p);
        return last_error();
    });
    if (rc != TMG_OK) return rc;
    return launch_pregen(e, st);
}

int tmg_step_many(tmg_env* e, const int32_t* actions_dev, int32_t num_steps, int32_t* rewards_dev, uint8_t* terminated_dev,
                  void* stream) {
    if (!actions_dev) return TMG_ERR_INVALID_ARG;
    return rollout(e, POLICY_GIVEN, actions_dev, num_steps, nullptr, rewards_dev, terminated_dev, stream);
}

int tmg_rollout_policy(tmg_env* e, int32_t policy, int32_t num_steps, int32_t* actions_out_dev, int32_t* rewards_dev,
                       uint8_t* terminated_dev, void* stream) {
    if (policy != TMG_POLICY_UNIFORM && policy != TMG_POLICY_MASK) return TMG_ERR_INVALID_ARG;
    return rollout(e, policy, nullptr, num_steps, actions_out_dev, rewards_dev, terminated_dev, stream);
}

int tmg_legal_mask(tmg_env* e, void* stream) {
    if (!e) return TMG_ERR_INVALID_ARG;
    if (cudaSetDevice(e->cfg.device) != cudaSuccess) return TMG_ERR_CUDA;
    const Params p = e->p;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rc = launch_by_lanes(e, [&](auto lanes) {
        constexpr int L = decltype(lanes)::value;
        k_mask<L><<<grid_for<L>(p.N), Cfg<L>::THREADS, Cfg<L>::GPB * sizeof(GroupSmem<L>), st>>>(p);
        return last_error();
    });
    return rc != TMG_OK ? rc : refresh_mirror(e, st);
}

static int onehot_launch(tmg_env* e, void* out, int elem_bytes, void* stream) {
    if (!e || !out) return TMG_ERR_INVALID_ARG;
    if (cudaSetDevice(e->cfg.device) != cudaSuccess) return TMG_ERR_CUDA;
    const Params p = e->p;
    const long long threads = (long long)p.N * p.P;   // one thread per cell
    const int block = 256;
    const long long grid = (threads + block - 1) / block;
    if (grid > 0x7fffffffLL) return TMG_ERR_INVALID_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (elem_bytes == 8) k_onehot<double><<<(unsigned)grid, block, 0, st>>>(p, static_cast<double*>(out), e->planes);
    else if (elem_bytes == 4) k_onehot<float><<<(unsigned)grid, block, 0, st>>>(p, static_cast<float*>(out), e->planes);
    else k_onehot<uint8_t><<<(unsigned)grid, block, 0, st>>>(p, static_cast<uint8_t*>(out), e->planes);
    return last_error();
}
int tmg_encode_onehot(tmg_env* e, uint8_t* out_dev, void* stream) { return onehot_launch(e, out_dev, 1, stream); }
int tmg_encode_onehot_f32(tmg_env* e, float* out_dev, void* stream) { return onehot_launch(e, out_dev, 4, stream); }
int tmg_encode_onehot_f64(tmg_env* e, double* out_dev, void* stream) { return onehot_launch(e, out_dev, 8, stream); }

int tmg_clear_status(tmg_env* e, void* stream) {
    if (!e) return TMG_ERR_INVALID_ARG;
    if (cudaSetDevice(e->cfg.device) != cudaSuccess) return TMG_ERR_CUDA;
    k_clear_status<<<(e->p.N + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(e->p.status, e->p.N);
    return last_error();
}

int tmg_join(tmg_env* e, void* stream) {
    if (!e) return TMG_ERR_INVALID_ARG;
    if (cudaSetDevice(e->cfg.device) != cudaSuccess) return TMG_ERR_CUDA;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (e->pregen && e->steps_since_pregen > 0) {   // requests collected since the last launch: serve them now
        Params p = e->p;
        p.pool_tag = (int)(e->pregen_count & 0x7fffffff);
        k_commit_batch<<<1, 1, 0, st>>>(p);
        if (last_error() != TMG_OK) return TMG_ERR_CUDA;
        const int rc = launch_pregen(e, st);
        if (rc != TMG_OK) return rc;
    }
    return join_pregen(e, st) ? TMG_OK : TMG_ERR_CUDA;
}

int tmg_set_seed(tmg_env* e, uint64_t seed, void* stream) {
    if (!e) return TMG_ERR_INVALID_ARG;
    if (cudaSetDevice(e->cfg.device) != cudaSuccess) return TMG_ERR_CUDA;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!join_pregen(e, st)) return TMG_ERR_CUDA;
    if (cudaMemsetAsync(e->p.episode, 0xff, (size_t)e->p.N * 4, st) != cudaSuccess) return TMG_ERR_CUDA;
    if (cudaMemsetAsync(e->p.pool_episode, 0x80, (size_t)e->p.N * 4, st) != cudaSuccess) return TMG_ERR_CUDA;
    e->cfg.seed = seed;
    e->p.key0 = (uint32_t)seed;
    e->p.key1 = (uint32_t)(seed >> 32);
    if (cudaMemsetAsync(e->p.draw_cursor, 0, (size_t)e->p.N * 8, st) != cudaSuccess) return TMG_ERR_CUDA;
    if (cudaMemsetAsync(e->p.shuffle_cursor, 0, (size_t)e->p.N * 8, st) != cudaSuccess) return TMG_ERR_CUDA;
    return TMG_OK;
}

int tmg_step_host(tmg_env* e, const tmg_host_io* io, void* stream) {
    if (!e || !io || !io->actions) return TMG_ERR_INVALID_ARG;
    if (cudaSetDevice(e->cfg.device) != cudaSuccess) return TMG_ERR_CUDA;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Params& p = e->p;
    const size_t N = (size_t)p.N;
    // With a host mirror bound, actions in page-locked memory are read in place by k_gate (one coalesced PCIe read per
    // warp) instead of being staged by a copy; pageable actions are staged as usual.
    const int32_t* actions = e->actions_dev;
    void* in_place = nullptr;
    if (e->hm_bound && reinterpret_cast<uintptr_t>(io->actions) % 4 == 0 &&
        cudaHostGetDevicePointer(&in_place, const_cast<int32_t*>(io->actions), 0) == cudaSuccess && in_place) {
        actions = static_cast<const int32_t*>(in_place);
    } else {
        cudaGetLastError();
        if (cudaMemcpyAsync(e->actions_dev, io->actions, N * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) return TMG_ERR_CUDA;
    }
    const int rc = tmg_step(e, actions, stream);
    if (rc != TMG_OK) return rc;
    bool ok = true;
    auto back = [&](void* dst, const void* src, size_t bytes) {
        if (dst) ok &= cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    };
    // arrays bound as the host mirror were already updated in place by the step kernel
    if (io->board != e->hm_board) back(io->board, p.board, N * 2 * p.P);
    if (io->board_packed && io->board_packed != e->hm_board_packed) {
        if (p.K > 15) return TMG_ERR_INVALID_ARG;
        const long long total = (long long)N * p.P;
        k_pack_boards<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p.board, e->board_packed_dev, (int)N, p.P);
        ok &= cudaGetLastError() == cudaSuccess;
        back(io->board_packed, e->board_packed_dev, (size_t)total);
    }
    if (io->reward != e->hm_reward) back(io->reward, p.reward, N * 4);
    if (io->terminated != e->hm_terminated) back(io->terminated, p.terminated, N);
    if (io->mask != e->hm_mask) back(io->mask, p.mask, N * p.A);
    if (io->mask_bits && io->mask_bits != e->hm_mask_bits) {
        const int bpe = (p.A + 7) / 8;
        const long long total = (long long)N * bpe;
        k_pack_mask<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p.mask, e->mask_bits_dev, (int)N, p.A, bpe);
        ok &= cudaGetLastError() == cudaSuccess;
        back(io->mask_bits, e->mask_bits_dev, (size_t)total);
    }
    if (io->num_moves_left != e->hm_moves_left) back(io->num_moves_left, p.moves_left, N * 4);
    back(io->is_combination_match, p.is_comb, N);
    back(io->num_new_specials, p.new_specials, N * 4);
    back(io->num_specials_activated, p.activated, N * 4);
    back(io->shuffled, p.shuffled, N);
    back(io->status, p.status, N * 4);
    if (!ok) return TMG_ERR_CUDA;
    // The caller blocks on every step, so the wake-up latency of the wait is on the critical path: poll the stream for
    // the usual length of a step before falling back to a blocking wait.
    for (int spins = 0; spins < 200000; ++spins) {
        const cudaError_t q = cudaStreamQuery(st);
        if (q == cudaSuccess) return TMG_OK;
        if (q != cudaErrorNotReady) { cudaGetLastError(); return TMG_ERR_CUDA; }
    }
    return cudaStreamSynchronize(st) == cudaSuccess ? TMG_OK : TMG_ERR_CUDA;
}

int tmg_set_profile_buffer(tmg_env* e, uint32_t* prof_dev) {
    if (!e) return TMG_ERR_INVALID_ARG;
    e->p.prof = prof_dev;
    return TMG_OK;
}

static int debug_launch(tmg_env* e, int32_t op, const int32_t* args_dev, uint32_t* out_dev, void* stream) {
    const int base = op & 0xff;
    if (!e || base < TMG_OP_GRAVITY || base > TMG_OP_LINES || (op & ~(0xff | TMG_OP_BYTE_PLANES))) return TMG_ERR_INVALID_ARG;
    if (cudaSetDevice(e->cfg.device) != cudaSuccess) return TMG_ERR_CUDA;
    Params p = e->p;
    p.dbg_op = op;
    p.dbg_args = args_dev;
    p.dbg_out = out_dev;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!join_pregen(e, st)) return TMG_ERR_CUDA;
    const int rc = launch_by_lanes(e, [&](auto lanes) {
        constexpr int L = decltype(lanes)::value;
        k_debug<L><<<grid_for<L>(p.N), Cfg<L>::THREADS, Cfg<L>::GPB * sizeof(GroupSmem<L>), st>>>(p);
        return last_error();
    });
    return rc != TMG_OK ? rc : refresh_mirror(e, st);
}

int tmg_debug_op(tmg_env* e, int32_t op, const int32_t* args_dev, void* stream) {
    if ((op & 0xff) == TMG_OP_LINES) return TMG_ERR_INVALID_ARG;   // needs an output buffer: tmg_debug_lines
    return debug_launch(e, op, args_dev, nullptr, stream);
}

int tmg_debug_lines(tmg_env* e, uint32_t* out_dev, int32_t byte_planes, void* stream) {
    if (!out_dev) return TMG_ERR_INVALID_ARG;
    return debug_launch(e, TMG_OP_LINES | (byte_planes ? TMG_OP_BYTE_PLANES : 0), nullptr, out_dev, stream);
}

}  // extern "C"
