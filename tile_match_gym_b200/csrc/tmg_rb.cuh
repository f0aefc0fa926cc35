// tmg_rb.cuh -- the register-resident board engine of the step kernels (included by tmg_device.cuh, namespace tmg).
//
// One warp owns one board and lane c owns column c, as in Board<32> -- but the column does not live in shared memory
// as bytes: it is TWO REGISTERS of bit-planes,
//     cw  bit (r + 10 j) = bit j of colour(r, c)          j = 0..2   (colours 0..7)
//     tw  bit (r + 10 j) = bit j of (type(r, c) & 7)                 (cookie -1 -> 7)
// for boards of up to 10 rows, 32 columns and 7 colours (BASELINE configs 1-3).  Everything a cascade round needs is
// then a handful of ALU operations on those two words plus warp collectives:
//   * "equal to the cell above / to the right" (get_colour_lines, ref :149-215) is one xor and a 3-plane fold,
//   * deleting cells is an AND with a row mask, gravity (ref :217-229) is a shift of the bits above a gap,
//   * refill (ref :231-241) ORs the drawn colour into the emptied top rows; the Philox blocks of the env's draw stream
//     are computed 128 words at a time and kept in shared memory across the rounds of a move,
//   * a laser / bomb / cookie sweep of activate_special (ref :473-556) finds "the first special in the reference's
//     visit order" with one ballot or shuffle and deletes everything before it with one mask; the DFS stack is held
//     across the lanes (frame d in lane d & 31), not in memory,
//   * process_colour_lines (ref :269-327) runs on all lanes at once: lane i holds line i, "first queued line that
//     shares a cell" is two bit tests per lane and a min-reduction over the queue keys, and the classified matches are
//     again one per lane,
//   * the legal-move mask (ref :735-787) gets its per-colour row sets from one xor + fold each.
// No step of a round reads or writes the byte planes; they exist only in HBM (the observation) and as the staging
// buffer of the coalesced load / store.  Rare paths (a board without a possible move -> shuffle, hand-made boards that
// need the literal window rule of is_move_effective) unpack to the byte planes and reuse Board<32>.
//
// Reference being restated (never copied): /root/reference/src/tile_match_gym/board.py, cited as "ref :NNN".
#pragma once

// The rare, large parts of a round -- the general path (line table, classification, resolution), the activation DFS and
// the combination match -- run OUT OF LINE on the board passed and returned by value in registers, one copy each per
// kernel, so that the common round (scan, fast path, gravity, refill) stays a small contiguous piece of code: the step
// kernels are bound by instruction supply (the SM's instruction caches hold 6 KB / 32 KB).
#ifndef TMG_RB_OUTLINE
#define TMG_RB_OUTLINE 0   // measured on B200 (65 536 envs): inline 358 M steps/s, out of line 346 M
#endif
struct RBState { uint32_t cw, tw; int n_new, n_act; uint32_t status; int aux; };
struct RBScanPack { int rstar; unsigned mv, hs, hcells, m; int vtop, has_v; unsigned E, D, T, S; };
template <int RT, int CT, bool INJ> __device__ __noinline__ RBState rb_activate_fn(RBState st, const Params* pp, GroupSmem<32>* sm, int lane, int env, int cell, int t, int counted);
template <int RT, int CT, bool INJ> __device__ __noinline__ RBState rb_general_fn(RBState st, const Params* pp, GroupSmem<32>* sm, int lane, int env, RBScanPack sc);
template <int RT, int CT, bool INJ> __device__ __noinline__ RBState rb_combination_fn(RBState st, const Params* pp, GroupSmem<32>* sm, int lane, int env, int i1, int i2);
// words [4 b0, 4 b0 + 128) of stream 0 of env `gid` -> wbuf (one Philox block per lane); one copy per kernel
#ifndef TMG_RB_FILL_NOINLINE
#define TMG_RB_FILL_NOINLINE 1
#endif
#if TMG_RB_FILL_NOINLINE
#define TMG_RB_FILL_ATTR __noinline__
#else
#define TMG_RB_FILL_ATTR __forceinline__
#endif
#ifndef TMG_RB_PROF
#define TMG_RB_PROF 0      // per-phase cycle counters of a move (tmg_set_profile_buffer): cost 6 % even when off, so a diagnostics build only
#endif
// Code-size knobs.  The step kernel is bound by instruction supply and every one of these measured as throughput on B200
// (65 536 / 1 048 576 envs, M env-steps/s): all loops unrolled and everything inline 372 / 692; refill rows rolled 375 / 702;
// + pack / unpack rolled 385 / 748; + Philox rolled by 2, byte-plane row loops rolled 402 / 809; + the cold ends of an item
// (playability fallback, in-step generate_board) and the combination match out of line 416 / 833; + host mirror out of
// line, injected refill out of this engine 440 / 935.  Not every cut pays: activate() at one call site and the cookie's
// colour count out of line measured 432 / 910, shortcuts for single-line rounds and lone lasers / bombs (MORE code, fewer
// executed instructions) 361 / 647 against 372 / 690 -- the kernel wants less code more than it wants fewer instructions.
#ifndef TMG_RB_ROLL_REFILL
#define TMG_RB_ROLL_REFILL 1
#endif
#ifndef TMG_RB_ROLL_PACK
#define TMG_RB_ROLL_PACK 1
#endif
#ifndef TMG_RB_OUTLINE_COMB
#define TMG_RB_OUTLINE_COMB 1
#endif
#ifndef TMG_RB_SINGLE_FALL
#define TMG_RB_SINGLE_FALL 1
#endif
__device__ TMG_RB_FILL_ATTR void rb_fill_words(uint32_t* wbuf, uint64_t b0, uint32_t gid, uint32_t key0, uint32_t key1, int lane) {
    TMG_SITE_HERE
    __syncwarp(0xffffffffu);                               // earlier reads of the buffer are done
    const uint64_t b = b0 + (uint64_t)lane;
    uint32_t w[4];
    philox4x32_10((uint32_t)b, (uint32_t)(b >> 32), gid, 0u, key0, key1, w);
    *reinterpret_cast<uint4*>(&wbuf[4 * lane]) = *reinterpret_cast<uint4*>(w);
    TMG_SITE_HERE
    __syncwarp(0xffffffffu);
}

// INJ: the instantiation can also take its draws from the injected tensor (tmg_set_injected_draws).  The step kernels
// instantiate INJ = false only -- injected refill runs on the byte planes -- so their hot code carries no trace of it;
// the known-answer entry point (k_debug) instantiates INJ = true.
template <int RT, int CT, bool INJ = false> struct RBoard {
    typedef Cfg<32> CF;
    static constexpr int PS = 10;                          // plane stride: rows per bit-plane
    static constexpr uint32_t REP = 0x00100401u;           // bit 0 of every plane
    static constexpr uint32_t SPREAD = 0x00040201u;        // (v * SPREAD) & REP puts bit j of v (0..7) at plane j
    static constexpr int DFS_CAP = 64;                     // activation frames: two registers per lane
    static constexpr int BIG = 0x7fffffff;

    GroupSmem<32>& s;
    const Params& p;
    const int lane;
    int env;
    const int R, C, P, K;
    uint32_t cw = 0u, tw = 0u;                             // this lane's column
    uint64_t dcur = 0ull;
    uint32_t gid;
    uint32_t status = 0u;
    int n_new = 0, n_act = 0;                              // counters, uniform (ref :343-344)
    const uint32_t specials;
    uint64_t pc_b0 = 0ull;                                 // s.wbuf holds stream words [4 pc_b0, 4 pc_b0 + 128) of this env
    bool pc_valid = false;
    int fg_top = 0, fg_len = 0;                            // the single gap a fast-path round left in this column
    bool fg_valid = false;
    bool literal_rounds = false;                           // debug entry point: a round leaves the board as the reference's resolve does
    uint32_t stk0 = 0u, stk1 = 0u;                         // DFS frames d (lane d) and 32 + d
    uint32_t prof_rounds = 0u, prof_general = 0u;
    uint32_t prof_cyc[4] = {0u, 0u, 0u, 0u};                 // diagnostics (p.prof set): cycles in scan + fast round, general path, fall + refill, rest
    const bool prof_on;

    __device__ RBoard(GroupSmem<32>& sm, const Params& pp, int lane_, int env_)
        : s(sm), p(pp), lane(lane_), env(env_), R(RT ? RT : pp.R), C(CT ? CT : pp.C), P(RT ? RT * CT : pp.P), K(pp.K),
          gid((uint32_t)(pp.env_id_offset + (uint64_t)env_)), specials(pp.specials), prof_on(TMG_RB_PROF && pp.prof != nullptr) {}

    static __device__ __forceinline__ bool supported(const Params& pp) { return pp.R <= PS && pp.K <= 7; }
    __device__ __forceinline__ RBState get_state() const { RBState st; st.cw = cw; st.tw = tw; st.n_new = n_new; st.n_act = n_act; st.status = status; st.aux = 0; return st; }
    __device__ __forceinline__ void set_state(const RBState& st) { cw = st.cw; tw = st.tw; n_new = st.n_new; n_act = st.n_act; status = st.status; }

    // ---- warp collectives (one board per warp: literal full mask) -------------------------------------------------------
    __device__ __forceinline__ unsigned ballot(bool pr TMG_SITE_P) const { TMG_SITE_SET return __ballot_sync(0xffffffffu, pr); }
    __device__ __forceinline__ void sync(TMG_SITE_P0) const { TMG_SITE_SET __syncwarp(0xffffffffu); }
    __device__ __forceinline__ int shfl(int v, int src TMG_SITE_P) const { TMG_SITE_SET return __shfl_sync(0xffffffffu, v, src); }
    __device__ __forceinline__ int radd(int v TMG_SITE_P) const { TMG_SITE_SET return __reduce_add_sync(0xffffffffu, v); }
    __device__ __forceinline__ int rmax(int v TMG_SITE_P) const { TMG_SITE_SET return __reduce_max_sync(0xffffffffu, v); }
    __device__ __forceinline__ int rmin(int v TMG_SITE_P) const { TMG_SITE_SET return __reduce_min_sync(0xffffffffu, v); }
    __device__ __forceinline__ unsigned ror(unsigned v TMG_SITE_P) const { TMG_SITE_SET return __reduce_or_sync(0xffffffffu, v); }
    __device__ __forceinline__ unsigned lt_mask() const { return (1u << lane) - 1u; }
    __device__ __forceinline__ unsigned from_right(unsigned v, int d TMG_SITE_P) const {
        TMG_SITE_SET
        const unsigned r = __shfl_down_sync(0xffffffffu, v, d);
        return (lane + d < 32) ? r : 0u;
    }
    __device__ __forceinline__ unsigned from_left(unsigned v, int d TMG_SITE_P) const {
        TMG_SITE_SET
        const unsigned r = __shfl_up_sync(0xffffffffu, v, d);
        return (lane - d >= 0) ? r : 0u;
    }

    // ---- bit-plane arithmetic ----------------------------------------------------------------------------------------------
    __device__ __forceinline__ unsigned rows_mask() const { return (1u << R) - 1u; }
    static __device__ __forceinline__ unsigned lowmask(int n) { return n >= 32 ? 0xffffffffu : ((1u << n) - 1u); }   // bits [0, n)
    __device__ __forceinline__ unsigned cols_mask() const { return lowmask(C); }
    static __device__ __forceinline__ uint32_t rep(unsigned m) { return m * REP; }                      // row set -> all three planes
    static __device__ __forceinline__ uint32_t spread(int v) { return ((uint32_t)(v & 7) * SPREAD) & REP; }
    static __device__ __forceinline__ int unspread(uint32_t y) { return (int)((y | (y >> (PS - 1)) | (y >> (2 * PS - 2))) & 7u); }
    __device__ __forceinline__ unsigned fold(uint32_t x) const { return (x | (x >> PS) | (x >> (2 * PS))) & rows_mask(); }
    __device__ __forceinline__ unsigned eqz(uint32_t x) const { return ~(x | (x >> PS) | (x >> (2 * PS))) & rows_mask(); }
    __device__ __forceinline__ unsigned colour_is(int k) const { return eqz(cw ^ (spread(k) * rows_mask())); }   // rows carrying colour k
    __device__ __forceinline__ unsigned colour_nonzero() const { return fold(cw); }
    // type sets of this column (type code = type & 7: 0 empty, 1 normal, 2 v-laser, 3 h-laser, 4 bomb, 7 cookie)
    __device__ __forceinline__ unsigned bits_T() const {      // type > 0
        const unsigned a = tw, b = tw >> PS, c = tw >> (2 * PS);
        return (a | b | c) & ~(a & b & c) & rows_mask();
    }
    __device__ __forceinline__ unsigned bits_S() const { return ((tw >> PS) | (tw >> (2 * PS))) & rows_mask(); }          // not in {0, 1}
    __device__ __forceinline__ unsigned bits_Ng() const { return tw & (tw >> PS) & (tw >> (2 * PS)) & rows_mask(); }     // type < 0
    __device__ __forceinline__ unsigned bits_normal() const { return tw & ~(tw >> PS) & ~(tw >> (2 * PS)) & rows_mask(); }
    __device__ __forceinline__ unsigned bits_tz() const { return eqz(tw); }                                              // type == 0
    __device__ __forceinline__ unsigned bits_gt1() const { return bits_S() & ~bits_Ng(); }                               // type > 1
    __device__ __forceinline__ unsigned bits_empty() const { return eqz(cw | tw); }                                      // (0, 0)
    __device__ __forceinline__ void clear_rows(unsigned m) { const uint32_t k = ~rep(m); cw &= k; tw &= k; }
    __device__ __forceinline__ void set_cell(int r, int colour, int type) {
        const uint32_t k = ~(REP << r);
        cw = (cw & k) | (spread(colour) << r);
        tw = (tw & k) | (spread(type) << r);
    }
    __device__ __forceinline__ int colour_at(int r) const { return unspread((cw >> r) & REP); }
    __device__ __forceinline__ int type_at(int r) const { const int t = unspread((tw >> r) & REP); return t == 7 ? -1 : t; }
    // uniform queries of one cell (every lane gets the answer)
    __device__ __forceinline__ int cell_colour(int r, int c TMG_SITE_P) const { TMG_SITE_SET return __shfl_sync(0xffffffffu, colour_at(r), c); }
    __device__ __forceinline__ int cell_type(int r, int c TMG_SITE_P) const { TMG_SITE_SET return __shfl_sync(0xffffffffu, type_at(r), c); }
    __device__ __forceinline__ void cell_rc(int cell, int& r, int& c) const { r = cell / C; c = cell - r * C; }

    // ---- byte planes <-> registers (s.board holds [2][R][C] int8, the HBM layout) ----------------------------------------------------
    __device__ __forceinline__ void pack_from_smem() {
        uint32_t a = 0u, b = 0u;
        if (lane < C) {
#if TMG_RB_ROLL_PACK
#pragma unroll 1
#else
#pragma unroll
#endif
            for (int r = 0; r < (RT ? RT : PS); ++r) {
                if (RT || r < R) {
                    a |= spread((int)s.board[r * C + lane]) << r;
                    b |= spread((int)s.board[P + r * C + lane]) << r;
                }
            }
        }
        cw = a; tw = b;
        fg_valid = false;
    }
    __device__ __forceinline__ void unpack_to_smem() {
        if (lane < C) {
#if TMG_RB_ROLL_PACK
#pragma unroll 1
#else
#pragma unroll
#endif
            for (int r = 0; r < (RT ? RT : PS); ++r) {
                if (RT || r < R) {
                    s.board[r * C + lane] = (int8_t)colour_at(r);
                    s.board[P + r * C + lane] = (int8_t)type_at(r);
                }
            }
        }
    }
    // ---- draw stream: words [4 pc_b0, 4 pc_b0 + 128) of stream 0 in s.wbuf -----------------------------------------------------------
    __device__ __forceinline__ void fill_cache(uint64_t b0) {
        rb_fill_words(s.wbuf, b0, gid, p.key0, p.key1, lane);
        pc_b0 = b0;
        pc_valid = true;
    }
    __device__ __forceinline__ int injected_colour(int k) {
        if constexpr (INJ) {
            const int v = injected_draw(p.inj, p.inj_len, env, (long long)dcur + k);
            if (v < 0) { status |= ST_DRAWS_EXHAUSTED; return 1; }
            return v;
        } else {
            return 1;
        }
    }
    __device__ __forceinline__ bool use_inj() const { return INJ && p.use_inj; }

    // ---- gravity (ref :217-229) + refill (ref :231-241) of a cascade round -----------------------------------------------------------
    // general gravity of this lane's column: every run of empty cells below a tile is closed by shifting what is above it down
    __device__ __forceinline__ void gravity_column() {
#pragma unroll 1
        for (;;) {
            const unsigned occ = fold(cw | tw);                       // cells that are not (0, 0)
            if (!occ) break;
            const unsigned below_top = ~((2u << (__ffs((int)occ) - 1)) - 1u) & rows_mask();   // rows under the topmost tile
            const unsigned holes = ~occ & below_top;
            if (!holes) break;
            const int pb = 31 - __clz((int)holes);                    // bottom-most hole
            const int len = __clz((int)~(holes << (31 - pb)));         // length of the run of holes ending there
            const int top = pb - len + 1;
            const uint32_t keep = rep(~((2u << pb) - 1u) & rows_mask()), low = rep((1u << top) - 1u);
            cw = (cw & keep) | ((cw & low) << len);
            tw = (tw & keep) | ((tw & low) << len);
        }
    }
    // returns P - count_nonzero(type) of the round (ref :362, :374)
    __device__ __forceinline__ int fall_and_refill() {
        int e, gone;                                                  // this column: empties on top after the fall, type-0 cells
        if (fg_valid) {                                               // a fast-path round left at most one known gap per column
            fg_valid = false;
            const int len = fg_len, top = fg_top;
            if (len) {
                const uint32_t keep = rep(~((1u << (top + len)) - 1u) & rows_mask()), low = rep((1u << top) - 1u);
                cw = (cw & keep) | ((cw & low) << len);
                tw = (tw & keep) | ((tw & low) << len);
            }
            e = len;
            gone = len;                                               // the deleted cells are the type-0 cells
        } else {
            gone = lane < C ? __popc(bits_tz()) : 0;
            gravity_column();
            e = lane < C ? __popc(bits_empty()) : 0;
        }
        const unsigned m0 = ballot(e > 0);
        const int elim = radd(gone);                                  // (only needed at the return: off the critical path)
        if (!m0) return elim;                                         // ref :238: no rng call when nothing is empty
        const unsigned lt = lt_mask();
        constexpr int RR = RT ? RT : PS;
        // Common case: the cached window of the draw stream covers this refill whatever its size.  The rows are unrolled:
        // the ballot of the next row is in flight while this row's word is loaded (a lone warp -- the longest cascade of
        // a launch -- pays every dependent latency in full).
        const long long rel = (long long)(dcur - 4ull * pc_b0);
        if (!use_inj() && pc_valid && rel >= 0 && rel + P <= 128) {
            const int off = (int)rel;
            int base = 0;
            unsigned m = m0;
#if TMG_RB_ROLL_REFILL
#pragma unroll 1
#else
#pragma unroll
#endif
            for (int r = 0; r < RR; ++r) {                            // ref :239-241: k-th draw -> k-th empty cell, row-major
                if (e > r) {
                    const int k = 1 + (int)__umulhi(s.wbuf[off + base + __popc(m & lt)], (uint32_t)K);
                    cw |= spread(k) << r;
                    tw |= 1u << r;
                }
                base += __popc(m);
                if (r + 1 < RR) {
                    m = ballot(e > r + 1);
                    if (!m) break;
                }
            }
            dcur += (uint64_t)base;
            return elim;
        }
        const int maxe = rmax(e);
        const int total = radd(e);
        // ranks [ps, ps + nw) of a pass take words of the cached window; one pass unless the move draws > 125 tiles
#pragma unroll 1
        for (int ps = 0; ps < total;) {
            const uint64_t start = dcur + (uint64_t)ps;
            int off = 0, nw = total - ps;
            if (!use_inj()) {
                const long long rl = (long long)(start - 4ull * pc_b0);
                const bool inside = pc_valid && rl >= 0 && rl + nw <= 128;
                if (!inside && !(pc_valid && rl >= 0 && rl < 4)) fill_cache(start >> 2);
                off = (int)(start - 4ull * pc_b0);
                nw = min(nw, 128 - off);
            }
            int base = 0;
#pragma unroll 1
            for (int r = 0; r < maxe; ++r) {
                const unsigned mm = ballot(e > r);
                if (e > r) {
                    const int rank = base + __popc(mm & lt);
                    if (rank >= ps && rank < ps + nw) {
                        const int k = use_inj() ? injected_colour(rank) : 1 + (int)__umulhi(s.wbuf[off + rank - ps], (uint32_t)K);
                        cw |= spread(k) << r;
                        tw |= 1u << r;
                    }
                }
                base += __popc(mm);
            }
            ps += nw;
        }
        dcur += (uint64_t)total;
        return elim;
    }

    // ---- line detection (ref :149-215): the same row bitboards as Board::column_bits, from the planes -----------------------------
    struct Bits { unsigned E, D, T, S; };
    __device__ __forceinline__ Bits column_bits() const {
        Bits b;
        TMG_SITE_HERE
        const uint32_t right = __shfl_down_sync(0xffffffffu, cw, 1);
        b.E = (lane + 1 < C) ? eqz(cw ^ right) : 0u;                  // colour(r,c) == colour(r,c+1)
        b.D = (lane < C) ? (eqz(cw ^ (cw << 1)) & ~1u) : 0u;          // colour(r,c) == colour(r-1,c)
        b.T = bits_T();
        b.S = bits_S();
        return b;
    }
    struct Scan {
        int rstar; unsigned mv, hs, hcells, m; int vtop; bool has_v; Bits bits;
    };
    __device__ __forceinline__ Scan scan_lines() {
        Scan o;
        o.bits = column_bits();
        const Bits& b = o.bits;
        const unsigned V = b.D & (b.D << 1) & b.T;                    // vertical triple anchored (bottom) at r (ref :163-173)
        const unsigned H = b.E & from_right(b.E, 1) & b.T;            // horizontal triple anchored (left) at (r,c) (ref :179-189)
        const unsigned F = V | H;
        o.rstar = rmax(F ? 31 - __clz((int)F) : -1);
        o.mv = o.hs = o.hcells = o.m = 0u; o.vtop = 0; o.has_v = false;
        if (o.rstar < 0) return o;
        const int rs = o.rstar;
        const unsigned m = ballot((b.E >> rs) & 1u);
        const unsigned T = ballot((b.T >> rs) & 1u);
        o.has_v = (V >> rs) & 1u;
        o.mv = ballot(o.has_v);
        unsigned cand = m & (m >> 1) & T, hs = 0u, hcells = 0u;
        while (cand) {                                                // left to right; cells of a found line cannot anchor another (ref :179,192)
            const int sidx = __ffs((int)cand) - 1;
            const int run = __ffs((int)~(m >> sidx)) - 1;
            const unsigned cells = ((2u << run) - 1u) << sidx;
            hs |= 1u << sidx;
            hcells |= cells;
            cand &= ~cells;
        }
        o.hs = hs; o.hcells = hcells; o.m = m;
        if (o.has_v) o.vtop = rs - __clz((int)~(b.D << (31 - rs)));    // extend upwards while equal (ref :168-172)
        return o;
    }

    // Fast path of a cascade round: see Board::fast_round -- lines of 3 or 4 normal tiles, pairwise disjoint, no crossing
    // segments.  Returns the number of lines, or 0 if the general path must run.
    __device__ __forceinline__ int fast_round(const Scan& sc) {
        const Bits& b = sc.bits;
        const int rs = sc.rstar;
        const bool sp_v = specials & SP_VLASER, sp_h = specials & SP_HLASER;
        const bool mine = (sc.hcells >> lane) & 1u;
        const bool start = (sc.hs >> lane) & 1u;
        const int hlen = start ? __ffs((int)~(sc.m >> lane)) : 0;
        bool bad = (mine && ((b.S >> rs) & 1u)) || hlen > 4;
        const int vlen = sc.has_v ? rs - sc.vtop + 1 : 0;
        if (sc.mv) {
            const unsigned vrows = sc.has_v ? ((2u << rs) - 1u) & ~((1u << sc.vtop) - 1u) : 0u;
            const unsigned El = from_left(b.E, 1), Er = from_right(b.E, 1), Ell = from_left(El, 1);
            const unsigned cross = (b.E & Er) | (El & b.E) | (Ell & El);
            bad = bad || (vrows & (b.S | cross)) != 0u || vlen > 4;
        }
        if (ballot(bad)) return 0;
        const int laser = sp_h ? 3 : (sp_v ? 2 : 0);                  // horizontal 4-line (ref :297-302)
        const unsigned create = (laser && sc.hs) ? (ballot(start && hlen == 4) << 1) : 0u;   // its second cell (ref :453-456)
        const bool make = sc.has_v && vlen == 4 && sp_v;              // vertical 4-line -> vertical laser or normal
        fg_len = 0; fg_top = 0; fg_valid = true;
        if (mine) {
            if ((create >> lane) & 1u) tw = (tw & ~(REP << rs)) | (spread(laser) << rs);   // keeps the line's colour (ref :596-597)
            else { fg_top = rs; fg_len = 1; }
        }
        if (sc.has_v) {
            // the laser is created on the second cell (ref :453-456) and falls to the anchor row: write it there
            fg_top = sc.vtop;
            fg_len = make ? vlen - 1 : vlen;
            if (make) tw = (tw & ~(REP << rs)) | (spread(2) << rs);
        }
        if (literal_rounds) {                                         // known-answer entry point: the board after resolve, before gravity
            if (sc.has_v && make) {                                   // the laser sits on the second cell, the anchor is deleted
                const int k = colour_at(rs);
                clear_rows(1u << rs);
                set_cell(sc.vtop + 1, k, 2);
                clear_rows((lowmask(rs + 1) & ~lowmask(sc.vtop)) & ~(2u << sc.vtop));
            } else if (fg_len) clear_rows(lowmask(fg_top + fg_len) & ~lowmask(fg_top));
            fg_valid = false;
        }
        n_new += __popc(create) + (sc.mv ? __popc(ballot(make)) : 0);
        return __popc(sc.hs) + __popc(sc.mv);
    }

    // ---- line table in the reference's list order, as in Board::build_line_table ------------------------------------
    // entry: key (top row << 12 | list position, 16 bits) | kind << 16 | row-or-column << 17 | colour << 22, and the cell set
    static __device__ __forceinline__ uint32_t line_info(uint32_t key, int kind, int idx, int colour) {
        return key | ((uint32_t)kind << 16) | ((uint32_t)idx << 17) | ((uint32_t)colour << 22);
    }
    __device__ __forceinline__ int build_line_table(const Scan& sc) {
        const int rs = sc.rstar;
        const unsigned lt = lt_mask();
        const int before = __popc(sc.mv & lt) + __popc(sc.hs & lt);
        int n = __popc(sc.mv) + __popc(sc.hs);
        sync();
        if (sc.has_v) {
            const int slot = before;
            if (slot < 32) {
                s.line_mask[slot] = ((2u << rs) - 1u) & ~((1u << sc.vtop) - 1u);
                s.line_key[slot] = line_info(((uint32_t)sc.vtop << 12) | (uint32_t)slot, 1, lane, colour_at(rs));
            }
        }
        if ((sc.hs >> lane) & 1u) {
            const int slot = before + (sc.has_v ? 1 : 0);
            if (slot < 32) {
                const int run = __ffs((int)~(sc.m >> lane)) - 1;
                s.line_mask[slot] = ((2u << run) - 1u) << lane;
                s.line_key[slot] = line_info(((uint32_t)rs << 12) | (uint32_t)slot, 0, rs, colour_at(rs));
            }
        }
        if (sc.mv) {                                                  // phase 2 (ref :198-214)
            // a crossing segment needs a horizontal run of >= 3 equal cells through a cell of a vertical line (see fast_round):
            // only the rows where some line has one are visited, top to bottom
            const unsigned El = from_left(sc.bits.E, 1), Er = from_right(sc.bits.E, 1), Ell = from_left(El, 1);
            const unsigned cross = (sc.bits.E & Er) | (El & sc.bits.E) | (Ell & El);
            const unsigned vrows = sc.has_v ? ((2u << rs) - 1u) & ~((1u << sc.vtop) - 1u) : 0u;
            unsigned todo = ror(vrows & cross);
#pragma unroll 1
            for (; todo; todo &= todo - 1u) {
                const int r = __ffs((int)todo) - 1;
                const unsigned m = ballot((sc.bits.E >> r) & 1u);
                const unsigned T = ballot((sc.bits.T >> r) & 1u);
                const bool origin = sc.has_v && sc.vtop <= r;
                unsigned Q = ballot(origin);
                if (r == rs) Q |= sc.hcells;
                const unsigned pass = T & ~Q;
                int left = 0, right = 0;
                bool seg = false;
                if (origin && ((sc.bits.T >> r) & 1u)) {
                    const unsigned chain_r = (m << 1) & pass;
                    const unsigned chain_l = m & pass;
                    if (lane + 1 < 32) right = __ffs((int)~(chain_r >> (lane + 1))) - 1;
                    if (lane > 0) left = __clz((int)~(chain_l << (32 - lane)));
                    seg = (1 + left + right) >= 3;
                }
                const unsigned segm = ballot(seg);
                if (seg) {
                    const int slot = n + __popc(segm & lt);
                    if (slot < 32) {
                        s.line_mask[slot] = ((2u << (left + right)) - 1u) << (lane - left);
                        s.line_key[slot] = line_info(((uint32_t)r << 12) | (uint32_t)(1024 + lane * 32 + r), 0, r, colour_at(r));
                    }
                }
                n += __popc(segm);
            }
        }
        if (n > 32) { status |= ST_LINE_OVERFLOW; n = 32; }
        sync();
        return n;
    }

    // ---- activate_special (ref :473-556): explicit-stack DFS, all lanes, uniform control flow ---------------------------------------
    static __device__ __forceinline__ uint32_t frame(int kind, int cell, int cursor, int mc) {
        return (uint32_t)kind | ((uint32_t)cell << 2) | ((uint32_t)cursor << 12) | ((uint32_t)mc << 23);
    }
    __device__ __forceinline__ uint32_t stack_get(int d TMG_SITE_P) const { TMG_SITE_SET return (uint32_t)__shfl_sync(0xffffffffu, d < 32 ? stk0 : stk1, d & 31); }
    __device__ __forceinline__ void stack_set(int d, uint32_t f) {
        if (lane == (d & 31)) { if (d < 32) stk0 = f; else stk1 = f; }
    }
    __device__ __forceinline__ int count_nonzero_colours() { return radd(__popc(colour_nonzero())); }
    // A clipped window swept row-major from position `cur` (bomb 3x3 ref :517-528, bomb+bomb 5x5 ref :699-719): deletes what
    // the reference deletes before the first special at or after `cur` and returns that special's position (n if none).
    __device__ __forceinline__ int window_sweep(int min_r, int nr, int min_c, int w, int cur, bool only_normal) {
        const int n = w * nr, dc = lane - min_c;
        const bool inw = dc >= 0 && dc < w;
        const unsigned S = bits_S();
        int cand = BIG;
        if (inw) {
#pragma unroll 1
            for (int rr = 0; rr < nr; ++rr) {
                const int i = rr * w + dc;
                if (i >= cur && ((S >> (min_r + rr)) & 1u)) { cand = i; break; }
            }
        }
        int first = rmin(cand);
        if (first == BIG) first = n;
        if (inw) {
            unsigned del = 0u;
#pragma unroll 1
            for (int rr = 0; rr < nr; ++rr) {
                const int i = rr * w + dc;
                if (i >= cur && i < first) del |= 1u << (min_r + rr);
            }
            if (only_normal) del &= bits_normal();
            clear_rows(del);
        }
        return first;
    }
    // first cell at or after `cur` (row-major) of colour `kk` whose type is > 1 (strict) / not in {0,1}; -1 if none
    __device__ __forceinline__ int next_special_of_colour(int kk, int cur, bool strict) {
        int cr, cc;
        cell_rc(cur, cr, cc);
        unsigned m = colour_is(kk) & (strict ? bits_gt1() : bits_S());
        m &= (lane >= cc) ? ~((1u << cr) - 1u) : ~((2u << cr) - 1u);
        const int best = rmin((lane < C && m) ? ((__ffs((int)m) - 1) * 32 + lane) : BIG);
        if (best == BIG) return -1;
        return (best >> 5) * C + (best & 31);
    }
    __device__ __forceinline__ void enter_activation(int cell, int t, bool counted, int& sp) {
        int r0, c0;
        cell_rc(cell, r0, c0);
        const int own_nz = cell_colour(r0, c0) != 0;
        // ref :488-489 "all colours zero": impossible when the target itself is coloured
        if (!own_nz && count_nonzero_colours() == 0) return;
        if (t == 0 || t == 1) { status |= ST_INTERNAL; return; }      // ref :491-492 raises
        if (lane == c0) clear_rows(1u << r0);                         // ref :496
        if (counted) ++n_act;                                         // ref :498-499
        int kind, mc = 0;
        if (t == 2) kind = 0;
        else if (t == 3) kind = 1;
        else if (t == 4) kind = 2;
        else if (t == -1) {
            kind = 3;
            // ref :532-537: most common non-zero colour after the cookie deleted itself, lowest on ties
            uint32_t lo = 0u, hi = 0u;                                // per-colour counts, 8 bits each: colours 1-4, 5-7
#pragma unroll
            for (int k = 1; k <= 7; ++k) {
                const uint32_t c = (uint32_t)__popc(colour_is(k));
                if (k <= 4) lo |= c << (8 * (k - 1)); else hi |= c << (8 * (k - 5));
            }
            lo = (uint32_t)radd((int)lo);
            hi = (uint32_t)radd((int)hi);
            int best = 0;
#pragma unroll
            for (int k = 1; k <= 7; ++k) {
                const int c = (int)(((k <= 4 ? lo >> (8 * (k - 1)) : hi >> (8 * (k - 5)))) & 0xffu);
                if (c > best) { best = c; mc = k; }
            }
            if (best == 0) return;                                    // ref :533-534
            clear_rows(colour_is(mc) & bits_normal());                // ref :540-544
        } else { status |= ST_INTERNAL; return; }                     // ref :555-556 raises
        if (sp >= DFS_CAP) { status |= ST_DFS_OVERFLOW; return; }
        stack_set(sp, frame(kind, cell, 0, mc));
        ++sp;
    }
    __device__ __forceinline__ void activate(int cell, int t, bool counted) {   // one call, wherever a special is hit
#if TMG_RB_OUTLINE
        set_state(rb_activate_fn<RT, CT, INJ>(get_state(), &p, &s, lane, env, cell, t, (int)counted));
#else
        activate_impl(cell, t, counted);
#endif
    }
    __device__ __forceinline__ void activate_impl(int cell0, int t0, bool counted) {
        int sp = 0;
        int e_cell = cell0, e_t = t0;
        bool e_counted = counted;
#pragma unroll 1
        for (;;) {
            enter_activation(e_cell, e_t, e_counted, sp);             // nested calls always count (ref :505,513,526,554)
            e_counted = true;
            e_cell = -1;
#pragma unroll 1
            while (sp > 0) {
                const uint32_t f = stack_get(sp - 1);
                const int kind = (int)(f & 3u), cell = (int)((f >> 2) & 1023u), mc = (int)(f >> 23);
                const int cur = (int)((f >> 12) & 2047u);
                int r0, c0;
                cell_rc(cell, r0, c0);
                int target = -1, next = 0;
                if (kind == 0) {                                      // vertical laser: rows top to bottom (ref :502-507)
                    const unsigned from = ~((1u << cur) - 1u);
                    const unsigned spm = (unsigned)shfl((int)bits_S(), c0) & from;
                    const int first = spm ? __ffs((int)spm) - 1 : R;
                    if (lane == c0) clear_rows(((1u << first) - 1u) & from);
                    if (first < R) { target = first * C + c0; next = first + 1; }
                } else if (kind == 1) {                               // horizontal laser: columns left to right (ref :510-515)
                    const unsigned from = ~lowmask(cur) & cols_mask();
                    const unsigned spm = ballot((bits_S() >> r0) & 1u) & from;
                    const int first = spm ? __ffs((int)spm) - 1 : C;
                    if (lane >= cur && lane < first && lane < C) clear_rows(1u << r0);
                    if (first < C) { target = r0 * C + first; next = first + 1; }
                } else if (kind == 2) {                               // bomb: clipped 3x3, row-major (ref :517-528)
                    const int min_r = max(r0 - 1, 0), max_r = min(r0 + 1, R - 1);
                    const int min_c = max(c0 - 1, 0), max_c = min(c0 + 1, C - 1);
                    const int w = max_c - min_c + 1, nr = max_r - min_r + 1;
                    const int first = window_sweep(min_r, nr, min_c, w, cur, false);
                    if (first < w * nr) { target = (min_r + first / w) * C + min_c + first % w; next = first + 1; }
                } else {                                              // cookie: specials of its colour, row-major (ref :547-554)
                    const int cand = next_special_of_colour(mc, cur, true);
                    if (cand >= 0) { target = cand; next = cand + 1; }
                }
                if (target < 0) { --sp; continue; }
                int tr, tc;
                cell_rc(target, tr, tc);
                e_t = cell_type(tr, tc);
                e_cell = target;
                stack_set(sp - 1, frame(kind, cell, next, mc));
                break;
            }
            if (e_cell < 0) break;
        }
    }

    // ---- process_colour_lines (ref :269-327) + creation cells (ref :414-418, :429-458) + resolve (ref :397-427) -----------------------
    static __device__ __forceinline__ unsigned lowest_bits(unsigned m, int k) {
        unsigned out = 0u;
#pragma unroll 1
        for (int i = 0; i < k && m; ++i) { const unsigned b = m & (0u - m); out |= b; m ^= b; }
        return out;
    }
    static __device__ __forceinline__ int nth_bit(unsigned m, int k) {
#pragma unroll 1
        for (int i = 0; i < k; ++i) m &= m - 1u;
        return __ffs((int)m) - 1;
    }
    __device__ __forceinline__ int line_cell(int kind, int idx, int bit) const { return kind ? bit * C + idx : idx * C + bit; }
    // is the cell taken by an earlier creation?  takenB: lane c holds the rows of column c that are taken
    __device__ __forceinline__ bool is_taken(unsigned takenB, int cell) {
        int r, c;
        cell_rc(cell, r, c);
        return ((unsigned)shfl((int)takenB, c) >> r) & 1u;
    }
    // the whole general path of a round: classification with lane i = line i, resolution with activations, creation
    __device__ __forceinline__ void classify_and_resolve(int n) {
        const bool sp_cookie = specials & SP_COOKIE, sp_v = specials & SP_VLASER, sp_h = specials & SP_HLASER,
                   sp_bomb = specials & SP_BOMB;
        uint32_t li = lane < n ? s.line_key[lane] : 0u, lm = lane < n ? s.line_mask[lane] : 0u;
        bool alive = lane < n;
        int qkey = (int)(li & 0xffffu);                               // queue order: ref :282 stable sort by the first cell's row == by key
        int nslots = n, tail = 0x10000, nm = 0, ncq = 0;
        uint32_t flags = 0u;
        unsigned takenB = 0u;
        uint32_t m_info = 0u, m_mm = 0u, m_ext = 0u;                  // match j in lane j: kind | idx<<1 | (type&7)<<6 | colour<<9 | nextra<<12 | cq<<14
#pragma unroll 1
        for (;;) {
            const int kmin = rmin(alive ? qkey : BIG);                // ref :285 pop(0)
            if (kmin == BIG) break;
            const int src = __ffs((int)ballot(alive && qkey == kmin)) - 1;
            const uint32_t pinfo = (uint32_t)shfl((int)li, src);
            const unsigned pm = (unsigned)shfl((int)lm, src);
            if (lane == src) alive = false;
            const int kind = (int)((pinfo >> 16) & 1u), idx = (int)((pinfo >> 17) & 31u);
            int colour = (int)((pinfo >> 22) & 7u);
            const int len = __popc(pm);
            int name = NAME_NORMAL, nextra = 0, bomb_k2 = 0, bomb_i2 = 0;
            unsigned mm = 0u;
            uint32_t ext = 0u;
            if (len >= 5 && sp_cookie) {                              // ref :287-292
                mm = lowest_bits(pm, 5);
                name = NAME_COOKIE; colour = 0;
                const unsigned rest = pm & ~mm;
                if (__popc(rest) > 2) {
                    if (nslots < 32) {
                        if (lane == nslots) { li = pinfo; lm = rest; qkey = tail; alive = true; }
                        ++nslots; ++tail;
                    } else flags |= 1u;
                }
            } else if (len == 4) {                                    // ref :294-302
                mm = pm;
                name = (kind == 0 && sp_h) ? NAME_HLASER : (sp_v ? NAME_VLASER : NAME_NORMAL);
            } else {
                bool hit = false;
                if (sp_bomb) {                                        // ref :304-308: first queued line sharing a cell
                    const int k2 = (int)((li >> 16) & 1u), i2 = (int)((li >> 17) & 31u);
                    const bool shares = alive && ((k2 != kind) ? (((pm >> i2) & 1u) && ((lm >> idx) & 1u))
                                                               : (i2 == idx && (pm & lm) != 0u));
                    const int hk = rmin(shares ? qkey : BIG);
                    if (hk != BIG) {                                  // ref :309-320
                        hit = true;
                        const int hsrc = __ffs((int)ballot(shares && qkey == hk)) - 1;
                        const uint32_t hinfo = (uint32_t)shfl((int)li, hsrc);
                        const unsigned m2 = (unsigned)shfl((int)lm, hsrc);
                        const int hk2 = (int)((hinfo >> 16) & 1u), hi2 = (int)((hinfo >> 17) & 31u);
                        const int spos = (hk2 != kind) ? idx : __ffs((int)(pm & m2)) - 1;   // the shared cell along the crossing line
                        bomb_k2 = hk2; bomb_i2 = hi2;
                        mm = pm;
                        // ref :310-312: the three cells of l closest to the shared cell (stable: lower position first on ties)
                        unsigned picked = 0u;
#pragma unroll 1
                        for (int d = 0; d < 32 && __popc(picked) < 3 && picked != m2; ++d) {
#pragma unroll 1
                            for (int sgn = 0; sgn < (d ? 2 : 1); ++sgn) {
                                const int pos = sgn ? spos + d : spos - d;
                                if (pos < 0 || pos > 31 || !((m2 >> pos) & 1u) || __popc(picked) >= 3) continue;
                                picked |= 1u << pos;
                                const bool in_line = (hk2 != kind) ? (pos == idx) : (((pm >> pos) & 1u) != 0u);
                                if (!in_line) { ext |= (uint32_t)line_cell(hk2, hi2, pos) << (10 * nextra); ++nextra; }
                            }
                        }
                        name = NAME_BOMB;
                        if (__popc(m2) < 6) { if (lane == hsrc) alive = false; }   // ref :315-316
                        else if (lane == hsrc) lm = m2 & ~picked;                  // ref :317-319
                    }
                }
                if (!hit) {
                    if (len >= 3) mm = pm;                            // ref :322-325
                    else continue;
                }
            }
            int cq = 0x3ff;                                           // no special created
            if (name != NAME_NORMAL) {                                // ref :414-418
                int pos;
                if (name == NAME_BOMB) {                              // ref :441-450: (modal row, modal column), see Board::creation_pos_bomb
                    const int first_bit = __ffs((int)mm) - 1;
                    int mrow, mcol;
                    if (kind == 0) { mrow = idx; mcol = (bomb_k2 != kind && nextra >= 1) ? bomb_i2 : first_bit; }
                    else { mcol = idx; mrow = (bomb_k2 != kind && nextra >= 1) ? bomb_i2 : first_bit; }
                    pos = mrow * C + mcol;
                    if (is_taken(takenB, pos)) {                      // nearest valid cell of the match, first minimum (ref :448-450)
                        int best = -1, bestd = 0;
                        const int ncell = __popc(mm) + nextra;
#pragma unroll 1
                        for (int q = 0; q < ncell; ++q) {
                            const int nl = __popc(mm);
                            const int cell = q < nl ? line_cell(kind, idx, nth_bit(mm, q)) : (int)((ext >> (10 * (q - nl))) & 1023u);
                            if (is_taken(takenB, cell)) continue;
                            const int dr = cell / C - mrow, dcl = cell % C - mcol, d = dr * dr + dcl * dcl;
                            if (best < 0 || d < bestd) { best = cell; bestd = d; }
                        }
                        pos = best;
                    }
                } else {                                              // straight: middle of the valid cells (ref :453-458)
                    const unsigned tk = kind == 0 ? ballot((takenB >> idx) & 1u) : (unsigned)shfl((int)takenB, idx);
                    const unsigned valid = mm & ~tk;
                    const int nv = __popc(valid);
                    pos = nv ? line_cell(kind, idx, nth_bit(valid, (nv % 2 == 0) ? nv / 2 - 1 : nv / 2)) : -1;
                }
                if (pos < 0) { flags |= 2u; cq = 0x3fe; }
                else {
                    int pr, pc;
                    cell_rc(pos, pr, pc);
                    if (lane == pc) takenB |= 1u << pr;
                    cq = pos;
                }
                ++ncq;
            }
            if (nm < 32) {
                if (lane == nm) {
                    m_info = (uint32_t)kind | ((uint32_t)idx << 1) | ((uint32_t)(name & 7) << 6) | ((uint32_t)colour << 9) |
                             ((uint32_t)nextra << 12) | ((uint32_t)cq << 14);
                    m_mm = mm; m_ext = ext;
                }
                ++nm;
            } else flags |= 1u;
        }
        if (flags & 1u) status |= ST_LINE_OVERFLOW;
        if (flags & 2u) status |= ST_INTERNAL;
        // resolve_colour_matches (ref :421-423 -> :460-471): every cell of every match in list order
#pragma unroll 1
        for (int j = 0; j < nm; ++j) {
            const uint32_t info = (uint32_t)shfl((int)m_info, j);
            const unsigned mm = (unsigned)shfl((int)m_mm, j);
            const uint32_t ext = (uint32_t)shfl((int)m_ext, j);
            const int kind = (int)(info & 1u), idx = (int)((info >> 1) & 31u), nextra = (int)((info >> 12) & 3u);
            unsigned rem = mm;                                        // the line's cells, ascending
#pragma unroll 1
            for (;;) {
                const unsigned S = bits_S();
                const unsigned spm = (kind == 0 ? ballot((S >> idx) & 1u) : (unsigned)shfl((int)S, idx)) & rem;
                const unsigned upto = spm ? ((spm & (0u - spm)) - 1u) : 0xffffffffu;   // the cells before the first special
                const unsigned del = rem & upto;
                if (kind == 0) { if ((del >> lane) & 1u) clear_rows(1u << idx); }
                else if (lane == idx) clear_rows(del);
                if (!spm) break;
                const int first = __ffs((int)spm) - 1;
                const int cell = line_cell(kind, idx, first);
                int cr, cc;
                cell_rc(cell, cr, cc);
                const int t = cell_type(cr, cc);
                activate(cell, t, true);
                rem &= ~(upto | (1u << first));
            }
#pragma unroll 1
            for (int k = 0; k < nextra; ++k) {                        // then the bomb's extra cells (ref :312)
                const int cell = (int)((ext >> (10 * k)) & 1023u);
                int cr, cc;
                cell_rc(cell, cr, cc);
                const int t = cell_type(cr, cc);
                if (not01(t)) activate(cell, t, true);
                else if (lane == cc) clear_rows(1u << cr);
            }
        }
#pragma unroll 1
        for (int j = 0; j < nm; ++j) {                                // ref :426-427 -> create_special :572-597
            const uint32_t info = (uint32_t)shfl((int)m_info, j);
            const int cq = (int)((info >> 14) & 1023u);
            if (cq >= 0x3fe) continue;
            int cr, cc;
            cell_rc(cq, cr, cc);
            if (lane == cc) set_cell(cr, (int)((info >> 9) & 7u), (int)((info >> 6) & 7u));
        }
        n_new += ncq;
    }

    __device__ __forceinline__ int general_impl(const RBScanPack& sp) {
        Scan sc;
        sc.rstar = sp.rstar; sc.mv = sp.mv; sc.hs = sp.hs; sc.hcells = sp.hcells; sc.m = sp.m; sc.vtop = sp.vtop; sc.has_v = sp.has_v != 0;
        sc.bits.E = sp.E; sc.bits.D = sp.D; sc.bits.T = sp.T; sc.bits.S = sp.S;
        const int n = build_line_table(sc);
        classify_and_resolve(n);
        return n;
    }

    // one cascade round without gravity / refill (ref :369-373); returns the number of lines found
    __device__ __forceinline__ int resolve_round() {
        const long long t0 = (TMG_RB_PROF && prof_on) ? clock64() : 0;
        const Scan sc = scan_lines();
        if (sc.rstar < 0) return 0;
        ++prof_rounds;
        int n = fast_round(sc);
        const long long t1 = (TMG_RB_PROF && prof_on) ? clock64() : 0;
        if (n == 0) {                                                 // general path, out of line
            ++prof_general;
            RBScanPack sp;
            sp.rstar = sc.rstar; sp.mv = sc.mv; sp.hs = sc.hs; sp.hcells = sc.hcells; sp.m = sc.m; sp.vtop = sc.vtop; sp.has_v = sc.has_v;
            sp.E = sc.bits.E; sp.D = sc.bits.D; sp.T = sc.bits.T; sp.S = sc.bits.S;
#if TMG_RB_OUTLINE
            const RBState st = rb_general_fn<RT, CT, INJ>(get_state(), &p, &s, lane, env, sp);
            set_state(st);
            n = st.aux;
#else
            n = general_impl(sp);
#endif
        }
        if (TMG_RB_PROF && prof_on) { prof_cyc[0] += (uint32_t)(t1 - t0); prof_cyc[1] += (uint32_t)(clock64() - t1); }
        return n;
    }

    // ---- combination_match (ref :600-719), as Board::combination -----------------------------------------------------------------------
    __device__ __forceinline__ void combination(int i1, int i2) {
#if TMG_RB_OUTLINE || TMG_RB_OUTLINE_COMB
        set_state(rb_combination_fn<RT, CT, INJ>(get_state(), &p, &s, lane, env, i1, i2));
#else
        combination_impl(i1, i2);
#endif
    }
    __device__ __forceinline__ void combination_impl(int i1, int i2) {
        n_act += 2;                                                   // ref :609
        int r1, c1, r2, c2;
        cell_rc(i1, r1, c1);
        cell_rc(i2, r2, c2);
        const int t1 = cell_type(r1, c1), k1 = cell_colour(r1, c1), t2 = cell_type(r2, c2), k2 = cell_colour(r2, c2);
        const int r = min(r1, r2), c = min(c1, c2);
        enum { NONE, LIST2, CROSS, ROWMAJOR, WINDOW };
        int mode = NONE, kk = 0;
        bool strict = false;
        int min_r = 0, max_r = 0, min_c = 0, max_c = 0;
        auto del_cell = [&](int rr, int cc) { if (lane == cc) clear_rows(1u << rr); };
        if (t1 == -1 && t2 == -1) {                                   // ref :615-616
            cw = 0u; tw = 0u;
        } else if ((t1 == -1 && t2 == 1) || (t1 == 1 && t2 == -1)) {  // ref :619-641
            kk = (t1 == -1) ? k2 : k1;
            if (t1 == -1) del_cell(r1, c1); else del_cell(r2, c2);    // ref :626,628
            clear_rows(colour_is(kk) & bits_normal());                // ref :631-635
            mode = ROWMAJOR; strict = true;
            n_act -= 1;                                               // ref :641
        } else if ((t1 == -1 && t2 >= 2) || (t1 >= 2 && t2 == -1)) {  // ref :644-660
            kk = (t1 == -1) ? k2 : k1;
            const int tt = (t1 == -1) ? t2 : t1;
            if (t1 == -1) del_cell(r1, c1); else del_cell(r2, c2);    // ref :651
            const unsigned m = colour_is(kk) & bits_normal();         // ref :655-657: its normal tiles take the special's type
            tw = (tw & ~rep(m)) | (spread(tt) * m);
            mode = ROWMAJOR; strict = false;                          // ref :660
        } else if ((t1 == 2 || t1 == 3) && (t2 == 2 || t2 == 3)) {    // ref :663-674
            del_cell(r1, c1); del_cell(r2, c2);
            mode = LIST2;
        } else if ((t1 == 4 && (t2 == 2 || t2 == 3)) || (t2 == 4 && (t1 == 2 || t1 == 3))) {  // ref :677-696
            del_cell(r1, c1); del_cell(r2, c2);
            min_r = max(r - 1, 0); max_r = min(r + 1, R - 1);
            min_c = max(c - 1, 0); max_c = min(c + 1, C - 1);
            mode = CROSS;
        } else if (t1 == 4 && t2 == 4) {                              // ref :699-719
            del_cell(r1, c1); del_cell(r2, c2);
            min_r = max(r - 2, 0); max_r = min(r + 2, R - 1);
            min_c = max(c - 2, 0); max_c = min(c + 2, C - 1);
            mode = WINDOW;
        }
        const int nr = max_r - min_r + 1, w = max_c - min_c + 1;
        int cur = 0;
#pragma unroll 1
        while (mode != NONE) {
            int cell = -1, t = 0;
            if (mode == LIST2) {
                if (cur < 2) { cell = r * C + c; t = 2 + cur; ++cur; }
            } else if (mode == CROSS) {
                if (cur < nr) { cell = (min_r + cur) * C + c; t = 3; ++cur; }
                else if (cur < nr + w) { cell = r * C + min_c + (cur - nr); t = 2; ++cur; }
            } else if (mode == ROWMAJOR) {                            // activate_specials_in_mask (ref :721-726)
                const int cand = cur < P ? next_special_of_colour(kk, cur, strict) : -1;
                if (cand >= 0) {
                    int cr, cc;
                    cell_rc(cand, cr, cc);
                    cell = cand; t = cell_type(cr, cc); cur = cand + 1;
                }
            } else {                                                  // 5x5 window: normal -> delete, other non-empty -> activate
                const int first = window_sweep(min_r, nr, min_c, w, cur, true);
                if (first < w * nr) {
                    cell = (min_r + first / w) * C + min_c + first % w;
                    int cr, cc;
                    cell_rc(cell, cr, cc);
                    t = cell_type(cr, cc); cur = first + 1;
                }
            }
            if (cell < 0) break;
            activate(cell, t, false);
        }
    }

    // ---- move (ref :330-378) after the effectiveness gate ------------------------------------------------------------------------------------
    __device__ __forceinline__ void swap_cells(int i1, int i2) {     // swap_coords (ref :355, :729-732): i2 is below or right of i1
        int r1, c1;
        cell_rc(i1, r1, c1);
        if (i2 == i1 + C) {                                           // vertical: rows r1, r1 + 1 of one column
            if (lane == c1) {
                const uint32_t xc = ((cw >> r1) ^ (cw >> (r1 + 1))) & REP, xt = ((tw >> r1) ^ (tw >> (r1 + 1))) & REP;
                cw ^= (xc << r1) | (xc << (r1 + 1));
                tw ^= (xt << r1) | (xt << (r1 + 1));
            }
        } else {                                                      // horizontal: columns c1, c1 + 1 of row r1
            const uint32_t v = ((cw >> r1) & REP) | (((tw >> r1) & REP) << 1);
            TMG_SITE_HERE
            const uint32_t fr = __shfl_down_sync(0xffffffffu, v, 1), fl = __shfl_up_sync(0xffffffffu, v, 1);
            if (lane == c1 || lane == c1 + 1) {
                const uint32_t nv = lane == c1 ? fr : fl;
                cw = (cw & ~(REP << r1)) | ((nv & REP) << r1);
                tw = (tw & ~(REP << r1)) | (((nv >> 1) & REP) << r1);
            }
        }
    }
    // returns the eliminations counted after each gravity (ref :362, :374); is_comb <- is_combination_match
    __device__ __forceinline__ int move(int i1, int i2, int& is_comb) {
        n_new = 0; n_act = 0;                                         // ref :343-347
        swap_cells(i1, i2);
        int r1, c1, r2, c2;
        cell_rc(i1, r1, c1);
        cell_rc(i2, r2, c2);
        const int t1 = cell_type(r1, c1), t2 = cell_type(r2, c2);
        const bool comb = (not01(t1) && not01(t2)) || t1 < 0 || t2 < 0;   // ref :357-359
        int elim = 0;
        fg_valid = false;
        is_comb = comb;
#if TMG_RB_SINGLE_FALL
        if (comb) combination(i1, i2);                                // ref :361; its gravity + refill (ref :362-364) is the loop's first trip
        bool pending = comb;
#pragma unroll 1
        for (;;) {                                                    // ref :367-376 (one call site of the round and of the fall)
            if (!pending && resolve_round() == 0) break;
            pending = false;
            const long long t0 = (TMG_RB_PROF && prof_on) ? clock64() : 0;
            elim += fall_and_refill();
            if (TMG_RB_PROF && prof_on) prof_cyc[2] += (uint32_t)(clock64() - t0);
        }
#else
        if (comb) {
            combination(i1, i2);                                      // ref :361
            elim += fall_and_refill();                                // ref :362-364
        }
#pragma unroll 1
        while (resolve_round() != 0) {                                // ref :367-376
            const long long t0 = (TMG_RB_PROF && prof_on) ? clock64() : 0;
            elim += fall_and_refill();
            if (TMG_RB_PROF && prof_on) prof_cyc[2] += (uint32_t)(clock64() - t0);
        }
#endif
        return elim;                                                  // the caller adds num_new_specials (ref :378)
    }

    // ---- legal-move mask (ref tile_match_env.py:118-124, board.py:735-787), as Board::mask_bits -------------------------------------------------
    // returns possible_move(); *literal <- the board needs the literal window rule (the caller unpacks and uses Board)
    __device__ __forceinline__ bool mask_bits(unsigned& effv_out, unsigned& effh_out, bool& literal) {
        const bool in = lane < C;
        const unsigned S = bits_S(), Ng = bits_Ng();
        const bool odd = in && (Ng != (~colour_nonzero() & rows_mask()));     // a cell whose (colour == 0) disagrees with (type < 0)
        const unsigned rows = rows_mask(), rows_v = rows >> 1;
        unsigned effv = ((S & (S >> 1)) | Ng | (Ng >> 1));            // ref :750,754
        unsigned effh = (S & from_right(S, 1)) | Ng | from_right(Ng, 1);
        unsigned unstable = 0u;
#pragma unroll 1
        for (int k = 1; k <= K; ++k) {
            const unsigned b = in ? colour_is(k) : 0u;
            const unsigned l1 = from_left(b, 1), l2 = from_left(b, 2), r1 = from_right(b, 1), r2 = from_right(b, 2);
            const unsigned hl = l1 & l2, hm = l1 & r1, hr = r1 & r2;
            const unsigned vu = (b << 1) & (b << 2), vm = (b << 1) & (b >> 1), vd = (b >> 1) & (b >> 2);
            const unsigned hany = hl | hm | hr, vany = vu | vm | vd;
            unstable |= (b & (b >> 1) & (b >> 2)) | (b & r1 & r2);
            effh |= (r1 & (hl | vany)) | (b & from_right(hr | vany, 1));
            effv |= ((b >> 1) & (vu | hany)) | (b & ((vd | hany) >> 1));
        }
        if (!in || lane + 1 >= C) effh = 0u;
        if (!in) effv = 0u;
        effv &= rows_v;
        effh &= rows;
        literal = ballot(unstable != 0u || odd) != 0u;
        effv_out = effv; effh_out = effh;
        return ballot((effv | effh) != 0u) != 0u;
    }
    __device__ __forceinline__ int special_count() { return radd(__popc(bits_S())); }
};

template <int RT, int CT, bool INJ> __device__ __noinline__ RBState rb_activate_fn(RBState st, const Params* pp, GroupSmem<32>* sm, int lane, int env, int cell, int t, int counted) {
    RBoard<RT, CT, INJ> b(*sm, *pp, lane, env);
    b.set_state(st);
    b.activate_impl(cell, t, counted != 0);
    return b.get_state();
}
template <int RT, int CT, bool INJ> __device__ __noinline__ RBState rb_general_fn(RBState st, const Params* pp, GroupSmem<32>* sm, int lane, int env, RBScanPack sc) {
    RBoard<RT, CT, INJ> b(*sm, *pp, lane, env);
    b.set_state(st);
    const int n = b.general_impl(sc);
    RBState o = b.get_state();
    o.aux = n;
    return o;
}
template <int RT, int CT, bool INJ> __device__ __noinline__ RBState rb_combination_fn(RBState st, const Params* pp, GroupSmem<32>* sm, int lane, int env, int i1, int i2) {
    RBoard<RT, CT, INJ> b(*sm, *pp, lane, env);
    b.set_state(st);
    b.combination_impl(i1, i2);
    return b.get_state();
}
