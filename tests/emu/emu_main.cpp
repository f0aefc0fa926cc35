// TEST INFRASTRUCTURE ONLY -- see emu_shim.h.  Runs the product's kernels (tmg_device.cuh) on the CPU
// over host memory, one group at a time, so tests can fuzz the device logic against the oracle here.
#define TMG_EMU 1
#include "emu_shim.h"

#include <dlfcn.h>
#include <stdio.h>

#include <vector>

namespace tmg { alignas(16) unsigned char tmg_smem_raw[1 << 17]; }
#include "../../tile_match_gym_b200/csrc/tmg_device.cuh"

namespace emu {
Lane* cur = nullptr;
Dim block_idx, block_dim, grid_dim;
int group_lanes = 0, group_shift = 0, site_id = 0;

static ucontext_t sched_ctx;
static Lane lanes[32];
static uint64_t slots[32], results[2][32];
static volatile int arrived = 0, alive = 0;
static volatile unsigned long long gen = 0;
static unsigned expect_mask = 0;
static void (*entry_fn)() = nullptr;
static int ncoll[32];
static void* last_site[32];

static void __attribute__((noinline)) yield_to_sched() {
    swapcontext(&cur->ctx, &sched_ctx);
    asm volatile("" ::: "memory");
}

const uint64_t* gather(uint64_t v, unsigned mask) {
    if (mask != expect_mask) { fprintf(stderr, "emu: collective with mask %08x, group mask %08x\n", mask, expect_mask); abort(); }
    slots[cur->lane_in_warp] = v;
    ncoll[cur->lane_in_warp & 31]++;
    if (getenv("EMU_TRACE")) fprintf(stderr, "lane %d coll #%d site %p arrived %d alive %d gen %llu\n", cur->lane_in_warp, ncoll[cur->lane_in_warp & 31], __builtin_return_address(0), arrived, alive, gen);
    last_site[cur->lane_in_warp & 31] = __builtin_return_address(0);
    const unsigned long long my = gen;
    {   // every lane of the group must be at the same collective
        static int first_site, first_lane;
        if (arrived == 0) { first_site = site_id; first_lane = cur->lane_in_warp; }
        else if (site_id != first_site) {
            fprintf(stderr, "emu: divergent collective: lane %d at line %d, lane %d at line %d\n", first_lane, first_site,
                    cur->lane_in_warp, site_id);
            abort();
        }
    }
    arrived = arrived + 1;
    if (arrived == alive) {
        memcpy(results[my & 1], slots, sizeof(slots));
        arrived = 0;
        gen = gen + 1;
    } else {
        while (gen == my) yield_to_sched();
    }
    return results[my & 1];
}

static void trampoline() {
    if (getenv("EMU_TRACE")) fprintf(stderr, "lane %d start tid %u\n", cur->lane_in_warp, cur->tid.x);
    entry_fn();
    if (getenv("EMU_TRACE")) fprintf(stderr, "lane %d end\n", cur->lane_in_warp);
    cur->done = true;
    alive = alive - 1;
    if (arrived == alive && alive > 0) {  // a lane left while the others wait: a kernel bug in this design
        fprintf(stderr, "emu: lane %d exited while its group waits in a collective\n", cur->lane_in_warp);
        for (int i = 0; i < 32; ++i) if (ncoll[i]) { Dl_info di; dladdr(last_site[i], &di); fprintf(stderr, "  lane %d: %d collectives, last at +0x%lx\n", i, ncoll[i], (unsigned long)((char*)last_site[i] - (char*)di.dli_fbase)); }
        abort();
    }
    swapcontext(&cur->ctx, &sched_ctx);
}

// run one group: L lanes with threadIdx.x = first_tid + lane
static void run_group(int L, int first_tid, void (*fn)()) {
    entry_fn = fn;
    group_lanes = L;
    group_shift = first_tid & 31;   // groups need not be a power of two wide (10-lane groups: 0, 10, 20)
    expect_mask = (L == 32 ? 0xffffffffu : ((1u << L) - 1u)) << group_shift;
    alive = L; arrived = 0;
    memset(ncoll, 0, sizeof(ncoll));
    for (int i = 0; i < L; ++i) {
        Lane& ln = lanes[i];
        if (!ln.stack) ln.stack = (char*)malloc(1 << 18);
        getcontext(&ln.ctx);
        ln.ctx.uc_stack.ss_sp = ln.stack;
        ln.ctx.uc_stack.ss_size = 1 << 18;
        ln.ctx.uc_link = &sched_ctx;
        ln.tid = Dim{(unsigned)(first_tid + i), 0, 0};
        ln.lane_in_warp = group_shift + i;
        ln.done = false;
        makecontext(&ln.ctx, trampoline, 0);
    }
    bool any = true;
    while (any) {
        any = false;
        for (int i = 0; i < L; ++i) {
            if (lanes[i].done) continue;
            any = true;
            cur = &lanes[i];
            swapcontext(&sched_ctx, &lanes[i].ctx);
        }
    }
}
}  // namespace emu

using namespace tmg;

static Params g_params;
static bool mir() { return g_params.h_board || g_params.h_board_packed || g_params.h_mask || g_params.h_mask_bits || g_params.h_reward; }
static bool rbk() { return rb_supported(32, g_params.R, g_params.K, g_params.flags, g_params.use_inj); }
template <int L> static void e_reset() {
    if (g_params.R == 10 && g_params.C == 10 && L == 32) k_reset<32, 10, 10>(g_params);
    else if (g_params.R == 9 && g_params.C == 9 && L == 32) k_reset<32, 9, 9>(g_params);
    else if (g_params.R == 32 && g_params.C == 32 && L == 32) k_reset<32, 32, 32>(g_params);
    else k_reset<L, 0, 0>(g_params);
}
template <int L> static void e_step() {
    if (g_params.R == 10 && g_params.C == 10 && L == 32) (rbk() ? (mir() ? k_work<32, 10, 10, true, true>(g_params) : k_work<32, 10, 10, true, false>(g_params)) : (mir() ? k_work<32, 10, 10, false, true>(g_params) : k_work<32, 10, 10, false, false>(g_params)));
    else if (g_params.R == 9 && g_params.C == 9 && L == 32) (rbk() ? (mir() ? k_work<32, 9, 9, true, true>(g_params) : k_work<32, 9, 9, true, false>(g_params)) : (mir() ? k_work<32, 9, 9, false, true>(g_params) : k_work<32, 9, 9, false, false>(g_params)));
    else if (g_params.R == 32 && g_params.C == 32 && L == 32) (mir() ? k_work<32, 32, 32, false, true>(g_params) : k_work<32, 32, 32, false, false>(g_params));
    else if (rbk() && L == 32) (mir() ? k_work<L, 0, 0, true, true>(g_params) : k_work<L, 0, 0, true, false>(g_params));
    else (mir() ? k_work<L, 0, 0, false, true>(g_params) : k_work<L, 0, 0, false, false>(g_params));
}
static void e_gate() { k_gate(g_params); }
// one thread per env, whole warps (k_gate)
static void launch_threads(void (*fn)(), int n, int per_thread = 1) {
    const int grid = (n + 128 * per_thread - 1) / (128 * per_thread);
    emu::block_dim = emu::Dim{128u, 1, 1};
    emu::grid_dim = emu::Dim{(unsigned)grid, 1, 1};
    for (int b = 0; b < grid; ++b) {
        emu::block_idx = emu::Dim{(unsigned)b, 0, 0};
        for (int w = 0; w < 4; ++w) emu::run_group(32, w * 32, fn);
    }
}
template <int L> static void e_pregen() {
    if (g_params.R == 10 && g_params.C == 10 && L == 32) k_pregen<32, 10, 10>(g_params);
    else if (g_params.R == 9 && g_params.C == 9 && L == 32) k_pregen<32, 9, 9>(g_params);
    else if (g_params.R == 32 && g_params.C == 32 && L == 32) k_pregen<32, 32, 32>(g_params);
    else k_pregen<L, 0, 0>(g_params);
}
template <int L> static void e_rollout() {
    if (g_params.R == 10 && g_params.C == 10 && L == 32) (rbk() ? k_rollout<32, 10, 10, true>(g_params) : k_rollout<32, 10, 10, false>(g_params));
    else if (g_params.R == 9 && g_params.C == 9 && L == 32) (rbk() ? k_rollout<32, 9, 9, true>(g_params) : k_rollout<32, 9, 9, false>(g_params));
    else if (g_params.R == 32 && g_params.C == 32 && L == 32) k_rollout<32, 32, 32, false>(g_params);
    else if (rbk() && L == 32) k_rollout<L, 0, 0, true>(g_params);
    else k_rollout<L, 0, 0, false>(g_params);
}
template <int L> static void e_mask() { k_mask<L>(g_params); }
template <int L> static void e_debug() { k_debug<L>(g_params); }

template <int L> static void launch(void (*fn)(), int n) {
    const int gpb = Cfg<L>::GPB;
    const int grid = (n + gpb - 1) / gpb;
    emu::block_dim = emu::Dim{(unsigned)Cfg<L>::THREADS, 1, 1};
    emu::grid_dim = emu::Dim{(unsigned)grid, 1, 1};
    for (int b = 0; b < grid; ++b) {
        emu::block_idx = emu::Dim{(unsigned)b, 0, 0};
        for (int g = 0; g < gpb; ++g) emu::run_group(L, (g / Cfg<L>::GPW) * 32 + (g % Cfg<L>::GPW) * L, fn);
    }
}

struct EmuEnv {
    Params p;
    int L;
    int tag = 0, seq = 0;
    int pregen_every = 1, since_pregen = 0;   // batches of pool requests: k_pregen after every `pregen_every` steps
    std::vector<char> mem;
};

static int vec_width(size_t s) { int w = 16; while (w > 1 && s % (size_t)w) w >>= 1; return w; }

extern "C" {

struct emu_config {
    int32_t num_envs, num_rows, num_cols, num_colours, num_moves;
    uint32_t specials;
    int32_t autoreset, refill_mode;
    uint32_t flags;
    int32_t max_reset_iters;
    uint64_t seed, env_id_offset;
};
struct emu_buffers {
    void *board, *timer, *draw_cursor, *shuffle_cursor, *reward, *terminated, *is_combination_match, *num_new_specials,
        *num_specials_activated, *shuffled, *mask, *num_moves_left, *status, *episode;
};

void* emu_create(const emu_config* c) {
    EmuEnv* e = new EmuEnv();
    Params& p = e->p;
    memset(&p, 0, sizeof(p));
    p.N = c->num_envs; p.R = c->num_rows; p.C = c->num_cols; p.K = c->num_colours; p.P = p.R * p.C;
    p.A = 2 * p.R * p.C - p.R - p.C; p.num_moves = c->num_moves; p.specials = c->specials;
    p.autoreset = c->autoreset; p.use_inj = c->refill_mode == 1; p.flags = c->flags;
    p.max_iters = c->max_reset_iters > 0 ? c->max_reset_iters : 16384;
    p.key0 = (uint32_t)c->seed; p.key1 = (uint32_t)(c->seed >> 32); p.env_id_offset = c->env_id_offset;
    p.board_vecw = vec_width((size_t)2 * p.P); p.mask_vecw = vec_width((size_t)p.A); p.init_vecw = 1;
    // 10x10 / 9x9 / 32x32 run the fixed-shape one-board-per-warp kernels (as the library does); the other shapes keep
    // the sub-warp groups (8, 10, 16 lanes; the library's TMG_B200_LANES knob) covered in the emulator
    const bool fixed = (p.R == 10 && p.C == 10) || (p.R == 9 && p.C == 9);
    e->L = (p.R > 16 || fixed) ? 32 : (p.C <= 8 ? 8 : (p.C <= 10 ? 10 : (p.C <= 16 ? 16 : 32)));
    if (getenv("TMG_EMU_LANES32")) e->L = 32;
    if (const char* pe = getenv("TMG_EMU_PREGEN_EVERY")) e->pregen_every = atoi(pe) > 0 ? atoi(pe) : 1;
    const size_t N = (size_t)p.N;
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = (off + b + 255) / 256 * 256; return o; };
    size_t cap = 1;
    while (cap < 8 * N) cap <<= 1;
    size_t o[22] = {take(N * 2 * p.P), take(N * 4), take(N * 8), take(N * 8), take(N * 4), take(N), take(N), take(N * 4),
                    take(N * 4), take(N), take(N * p.A), take(N * 4), take(N * 4), take(N * 4), take(N * 4),
                    take(N * 2 * p.P), take(N * p.A), take(N * 4), take(CTL_WORDS * 4), take(N * 8), take(cap * 8),
                    take(N)};
    e->mem.assign(off + 256, 0);
    char* b = e->mem.data();
    b += (256 - ((uintptr_t)b & 255)) & 255;
    p.board = (int8_t*)(b + o[0]); p.timer = (int32_t*)(b + o[1]); p.draw_cursor = (uint64_t*)(b + o[2]);
    p.shuffle_cursor = (uint64_t*)(b + o[3]); p.reward = (int32_t*)(b + o[4]); p.terminated = (uint8_t*)(b + o[5]);
    p.is_comb = (uint8_t*)(b + o[6]); p.new_specials = (int32_t*)(b + o[7]); p.activated = (int32_t*)(b + o[8]);
    p.shuffled = (uint8_t*)(b + o[9]); p.mask = (uint8_t*)(b + o[10]); p.moves_left = (int32_t*)(b + o[11]);
    p.status = (uint32_t*)(b + o[12]);
    p.episode = (int32_t*)(b + o[13]); p.pool_episode = (int32_t*)(b + o[14]);
    p.pool_board = (int8_t*)(b + o[15]); p.pool_mask = (uint8_t*)(b + o[16]); p.pool_status = (uint32_t*)(b + o[17]);
    p.ctl = (uint32_t*)(b + o[18]); p.wl_items = (uint2*)(b + o[19]); p.req_mask = (uint32_t)(cap - 1);
    const bool pregen = !p.use_inj && !(p.flags & 2u) && p.autoreset != 0;
    p.req_ring = pregen ? (uint2*)(b + o[20]) : nullptr;
    p.n_special = (uint8_t*)(b + o[21]);
    for (size_t i = 0; i < N; ++i) { p.timer[i] = -1; p.episode[i] = -1; p.pool_episode[i] = (int32_t)0x80808080; }
    return e;
}
void emu_destroy(void* h) { delete (EmuEnv*)h; }
void emu_get_buffers(void* h, emu_buffers* o) {
    const Params& p = ((EmuEnv*)h)->p;
    o->board = p.board; o->timer = p.timer; o->draw_cursor = p.draw_cursor; o->shuffle_cursor = p.shuffle_cursor;
    o->reward = p.reward; o->terminated = p.terminated; o->is_combination_match = p.is_comb;
    o->num_new_specials = p.new_specials; o->num_specials_activated = p.activated; o->shuffled = p.shuffled;
    o->mask = p.mask; o->num_moves_left = p.moves_left; o->status = p.status; o->episode = p.episode;
}
void emu_host_bind_packed(void* h, uint8_t* board_packed) { ((EmuEnv*)h)->p.h_board_packed = board_packed; }
void emu_host_bind(void* h, int8_t* board, uint8_t* mask, uint8_t* mask_bits, int32_t* reward, uint8_t* terminated,
                   int32_t* moves_left) {   // the mirror is ordinary memory here
    Params& p = ((EmuEnv*)h)->p;
    p.h_board = board; p.h_mask = mask; p.h_mask_bits = mask_bits;
    p.h_reward = reward; p.h_terminated = terminated; p.h_moves_left = moves_left;
}
void emu_set_injected_draws(void* h, const uint8_t* d, int64_t len) { ((EmuEnv*)h)->p.inj = d; ((EmuEnv*)h)->p.inj_len = len; }

#define DISPATCH(FN)                                   \
    switch (e->L) {                                    \
        case 8: launch<8>(FN<8>, e->p.N); break;       \
        case 10: launch<10>(FN<10>, e->p.N); break;    \
        case 16: launch<16>(FN<16>, e->p.N); break;    \
        default: launch<32>(FN<32>, e->p.N); break;    \
    }

void emu_reset(void* h, const uint8_t* reset_mask, const int8_t* init_boards) {
    EmuEnv* e = (EmuEnv*)h;
    g_params = e->p; g_params.reset_mask = reset_mask; g_params.init_boards = init_boards; g_params.pool_tag = e->tag; g_params.commit_pregen = 1;
    DISPATCH(e_reset)
    if (!e->p.use_inj && !(e->p.flags & 2u) && e->p.autoreset != 0) {
        g_params = e->p; g_params.pool_tag = e->tag++;
        DISPATCH(e_pregen)
    }
}
void emu_step(void* h, const int32_t* actions) {
    EmuEnv* e = (EmuEnv*)h;
    const bool close_batch = ++e->since_pregen >= e->pregen_every;
    g_params = e->p; g_params.actions = actions; g_params.pool_tag = e->tag; g_params.seq = e->seq++ & 1; g_params.commit_pregen = close_batch;
    launch_threads(e_gate, e->p.N, GATE_EPT);
    DISPATCH(e_step)
    if (close_batch) e->since_pregen = 0;
    if (close_batch && !e->p.use_inj && !(e->p.flags & 2u) && e->p.autoreset != 0) {
        g_params = e->p; g_params.pool_tag = e->tag++;
        DISPATCH(e_pregen)
    }
}
void emu_step_many(void* h, const int32_t* actions, int T, int32_t* rewards, uint8_t* terminated, int policy, int32_t* actions_out) {
    EmuEnv* e = (EmuEnv*)h;
    g_params = e->p; g_params.actions = actions; g_params.T = T; g_params.ro_reward = rewards; g_params.ro_terminated = terminated;
    g_params.policy = policy; g_params.ro_actions = actions_out;
    g_params.pool_tag = e->tag; g_params.commit_pregen = 1;
    DISPATCH(e_rollout)
    if (!e->p.use_inj && !(e->p.flags & 2u) && e->p.autoreset != 0) {
        g_params = e->p; g_params.pool_tag = e->tag++;
        DISPATCH(e_pregen)
    }
}
void emu_legal_mask(void* h) {
    EmuEnv* e = (EmuEnv*)h;
    g_params = e->p;
    DISPATCH(e_mask)
}
void emu_debug_op(void* h, int op, const int32_t* args) {
    EmuEnv* e = (EmuEnv*)h;
    g_params = e->p; g_params.dbg_op = op; g_params.dbg_args = args;
    DISPATCH(e_debug)
}
void emu_debug_lines(void* h, uint32_t* out, int byte_planes) {
    EmuEnv* e = (EmuEnv*)h;
    g_params = e->p; g_params.dbg_op = OP_LINES | (byte_planes ? OP_BYTE_PLANES : 0); g_params.dbg_args = nullptr; g_params.dbg_out = out;
    DISPATCH(e_debug)
}
}
