/*
 * TEST INFRASTRUCTURE ONLY -- see tmg_oracle.h.
 *
 * Literal plain-C restatement of the reference's list-processing algorithm.  Every function
 * cites the reference lines it follows ("ref board.py:NNN" = /root/reference/src/tile_match_gym/board.py).
 * Python lists become explicit ordered arrays; sets become membership flags; iteration
 * orders, stable sorts and "first maximal element" rules are kept exactly, because the
 * results depend on them.
 */
#include "tmg_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define MAXDIM 64
#define MAXLEN (MAXDIM + 8)

typedef struct {
    int n;
    int cells[MAXLEN]; /* r*C+c */
} line_t;

typedef struct {
    int n;
    line_t **l; /* ordered pointers into pool */
    line_t *pool;
    int pool_n, pool_cap;
} linelist_t;

enum { NAME_NORMAL = 0, NAME_VLASER = 1, NAME_HLASER = 2, NAME_BOMB = 3, NAME_COOKIE = 4 };

struct tmgo_board {
    int R, C, K, P;
    uint32_t specials;
    int32_t *colour, *type;
    /* counters (ref board.py:343-344) */
    int num_specials_activated, num_new_specials;
    /* stream */
    uint64_t seed;
    uint32_t env_id;
    uint64_t draw_cursor, shuffle_cursor;
    /* episode-indexed reset streams (oracle/stream.py): board j of an env is a pure function of (seed, env, j) */
    int32_t episode;
    int in_reset;
    int constructive;                               /* generate_board -> constructive sampler (not the reference) */
    uint64_t rdc, rsc;
    const uint8_t *inj;
    int64_t inj_len;
    int use_inj;
    uint32_t status;
    int64_t iter_cap;
    /* philox block cache */
    uint64_t cache_blk[2];
    uint32_t cache_w[2][4];
    int cache_ok[2];
    /* workspace */
    linelist_t lines, matches;
    int *match_name, *match_colour;
    uint8_t *flagA, *flagB, *flagC;
    int32_t *tmp;
    /* diagnostics */
    int max_lines, max_depth, depth;
    int64_t reset_iters;
};

/* -------------------------------------------------------------------------------------------- */
/* stream: Philox4x32-10 (Random123), see oracle/stream.py for the spec                           */
/* -------------------------------------------------------------------------------------------- */
void tmgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int i = 0; i < 10; i++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

uint32_t tmgo_stream_word(uint64_t seed, uint32_t env_id, uint32_t stream, uint64_t k) {
    uint64_t blk = k >> 2;
    uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), env_id, stream};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t out[4];
    tmgo_philox4x32_10(ctr, key, out);
    return out[k & 3];
}

static uint32_t reset_word(tmgo_board *b, uint32_t stream, uint64_t k) {
    uint32_t ctr[4] = {(uint32_t)(k >> 2), (uint32_t)b->episode, b->env_id, stream};
    uint32_t key[2] = {(uint32_t)b->seed, (uint32_t)(b->seed >> 32)};
    uint32_t out[4];
    tmgo_philox4x32_10(ctr, key, out);
    return out[k & 3];
}

static uint32_t board_word(tmgo_board *b, uint32_t stream, uint64_t k) {
    uint64_t blk = k >> 2;
    int s = stream & 1;
    if (!b->cache_ok[s] || b->cache_blk[s] != blk) {
        uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), b->env_id, stream};
        uint32_t key[2] = {(uint32_t)b->seed, (uint32_t)(b->seed >> 32)};
        tmgo_philox4x32_10(ctr, key, b->cache_w[s]);
        b->cache_blk[s] = blk;
        b->cache_ok[s] = 1;
    }
    return b->cache_w[s][k & 3];
}

/* np_random.integers(1, K+1, size=n) on the project stream (ref board.py:97,129,239) */
static void draw_colours(tmgo_board *b, int n, int32_t *out) {
    if (b->in_reset) {
        for (int i = 0; i < n; i++) {
            uint32_t w = reset_word(b, 3, b->rdc + (uint64_t)i);
            out[i] = 1 + (int32_t)(((uint64_t)w * (uint64_t)b->K) >> 32);
        }
        b->rdc += (uint64_t)n;
        return;
    }
    for (int i = 0; i < n; i++) {
        if (b->use_inj) {
            int64_t k = (int64_t)b->draw_cursor + i;
            if (k < b->inj_len) out[i] = b->inj[k];
            else { out[i] = 1; b->status |= TMGO_ST_DRAWS_EXHAUSTED; }
        } else {
            uint32_t w = board_word(b, 0, b->draw_cursor + (uint64_t)i);
            out[i] = 1 + (int32_t)(((uint64_t)w * (uint64_t)b->K) >> 32);
        }
    }
    b->draw_cursor += (uint64_t)n;
}

/* -------------------------------------------------------------------------------------------- */
/* list helpers                                                                                  */
/* -------------------------------------------------------------------------------------------- */
static void ll_init(linelist_t *ll, int cap) {
    ll->pool = (line_t *)malloc(sizeof(line_t) * (size_t)cap);
    ll->l = (line_t **)malloc(sizeof(line_t *) * (size_t)cap);
    ll->pool_cap = cap;
    ll->pool_n = 0;
    ll->n = 0;
}
static void ll_free(linelist_t *ll) { free(ll->pool); free(ll->l); }
static void ll_clear(linelist_t *ll) { ll->n = 0; ll->pool_n = 0; }
static line_t *ll_new(linelist_t *ll) { /* allocate + append at the back */
    if (ll->pool_n >= ll->pool_cap) return NULL;
    line_t *x = &ll->pool[ll->pool_n++];
    x->n = 0;
    ll->l[ll->n++] = x;
    return x;
}
static void ll_remove_at(linelist_t *ll, int i) {
    memmove(&ll->l[i], &ll->l[i + 1], sizeof(line_t *) * (size_t)(ll->n - i - 1));
    ll->n--;
}
static int line_eq(const line_t *a, const line_t *b) {
    if (a->n != b->n) return 0;
    for (int i = 0; i < a->n; i++) if (a->cells[i] != b->cells[i]) return 0;
    return 1;
}
static int line_has(const line_t *a, int cell) {
    for (int i = 0; i < a->n; i++) if (a->cells[i] == cell) return 1;
    return 0;
}
static void line_sort(line_t *a) { /* sorted(key=(row,col)) == ascending cell code; insertion sort */
    for (int i = 1; i < a->n; i++) {
        int v = a->cells[i], j = i - 1;
        while (j >= 0 && a->cells[j] > v) { a->cells[j + 1] = a->cells[j]; j--; }
        a->cells[j + 1] = v;
    }
}

/* -------------------------------------------------------------------------------------------- */
/* construction                                                                                  */
/* -------------------------------------------------------------------------------------------- */
tmgo_board *tmgo_board_create(int R, int C, int K, uint32_t specials) {
    if (R < 1 || C < 1 || R > MAXDIM || C > MAXDIM) return NULL;
    tmgo_board *b = (tmgo_board *)calloc(1, sizeof(*b));
    b->episode = -1;
    b->R = R; b->C = C; b->K = K; b->P = R * C; /* flat_size, ref board.py:56 */
    b->specials = specials;
    b->colour = (int32_t *)calloc((size_t)b->P, 4);
    b->type = (int32_t *)calloc((size_t)b->P, 4);
    int cap = 2 * b->P + 16;
    ll_init(&b->lines, cap);
    ll_init(&b->matches, cap);
    b->match_name = (int *)malloc(sizeof(int) * (size_t)cap);
    b->match_colour = (int *)malloc(sizeof(int) * (size_t)cap);
    b->flagA = (uint8_t *)malloc((size_t)b->P);
    b->flagB = (uint8_t *)malloc((size_t)b->P);
    b->flagC = (uint8_t *)malloc((size_t)b->P);
    b->tmp = (int32_t *)malloc(4 * (size_t)(b->P + 8));
    return b;
}
void tmgo_board_destroy(tmgo_board *b) {
    if (!b) return;
    free(b->colour); free(b->type);
    ll_free(&b->lines); ll_free(&b->matches);
    free(b->match_name); free(b->match_colour);
    free(b->flagA); free(b->flagB); free(b->flagC); free(b->tmp);
    free(b);
}
void tmgo_board_set_stream(tmgo_board *b, uint64_t seed, uint32_t env_id, uint64_t dc, uint64_t sc) {
    b->seed = seed; b->env_id = env_id; b->draw_cursor = dc; b->shuffle_cursor = sc;
    b->cache_ok[0] = b->cache_ok[1] = 0;
}
void tmgo_board_set_injected(tmgo_board *b, const uint8_t *draws, int64_t len, int64_t cursor) {
    b->inj = draws; b->inj_len = len; b->use_inj = draws != NULL;
    if (draws) b->draw_cursor = (uint64_t)cursor;
}
void tmgo_board_set_episode(tmgo_board *b, int32_t episode) { b->episode = episode; }
int32_t tmgo_board_get_episode(const tmgo_board *b) { return b->episode; }
void tmgo_board_get_cursors(const tmgo_board *b, uint64_t *dc, uint64_t *sc) { *dc = b->draw_cursor; *sc = b->shuffle_cursor; }
uint32_t tmgo_board_status(const tmgo_board *b) { return b->status; }
void tmgo_board_set(tmgo_board *b, const int32_t *planes) {
    memcpy(b->colour, planes, 4 * (size_t)b->P);
    memcpy(b->type, planes + b->P, 4 * (size_t)b->P);
}
void tmgo_board_get(const tmgo_board *b, int32_t *planes) {
    memcpy(planes, b->colour, 4 * (size_t)b->P);
    memcpy(planes + b->P, b->type, 4 * (size_t)b->P);
}
void tmgo_board_set_counters(tmgo_board *b, int nn, int na) { b->num_new_specials = nn; b->num_specials_activated = na; }
void tmgo_board_get_counters(const tmgo_board *b, int *nn, int *na) { *nn = b->num_new_specials; *na = b->num_specials_activated; }
void tmgo_board_set_iter_cap(tmgo_board *b, int64_t cap) { b->iter_cap = cap; }
void tmgo_board_diag(const tmgo_board *b, int *ml, int *md, int64_t *ri) { *ml = b->max_lines; *md = b->max_depth; *ri = b->reset_iters; }

/* ref board.py:77 */
int tmgo_num_actions(const tmgo_board *b) { return 2 * b->R * b->C - b->R - b->C; }

/* ref board.py:80-91 */
void tmgo_action_to_coords(const tmgo_board *b, int i, int *r1, int *c1, int *r2, int *c2) {
    int C = b->C, R = b->R;
    if (i < C * (R - 1)) { *r1 = i / C; *c1 = i % C; *r2 = *r1 + 1; *c2 = *c1; }
    else { int j = i - C * (R - 1); *r1 = j / (C - 1); *c1 = j % (C - 1); *r2 = *r1; *c2 = *c1 + 1; }
}

/* -------------------------------------------------------------------------------------------- */
/* swap / effectiveness (ref board.py:729-787)                                                   */
/* -------------------------------------------------------------------------------------------- */
static void swap_cells(tmgo_board *b, int i, int j) { /* ref board.py:729-732 */
    int32_t t = b->colour[i]; b->colour[i] = b->colour[j]; b->colour[j] = t;
    t = b->type[i]; b->type[i] = b->type[j]; b->type[j] = t;
}
static int not01(int t) { return t != 0 && t != 1; }

int tmgo_is_move_effective(tmgo_board *b, int r1, int c1, int r2, int c2) {
    int R = b->R, C = b->C;
    int i1 = r1 * C + c1, i2 = r2 * C + c2;
    if (not01(b->type[i1]) && not01(b->type[i2])) return 1;   /* ref :750 */
    if (b->type[i1] < 0 || b->type[i2] < 0) return 1;         /* ref :754 */
    int rmin = (r1 < r2 ? r1 : r2) - 2, rmax = (r1 > r2 ? r1 : r2) + 2; /* ref :758-761 */
    int cmin = (c1 < c2 ? c1 : c2) - 2, cmax = (c1 > c2 ? c1 : c2) + 2;
    if (rmin < 0) rmin = 0;
    if (rmax > R - 1) rmax = R - 1;
    if (cmin < 0) cmin = 0;
    if (cmax > C - 1) cmax = C - 1;
    swap_cells(b, i1, i2);                                     /* ref :764 */
    int found = 0;
    if (cmin + 2 <= cmax) {                                    /* ref :767-774 */
        for (int r = rmin; r <= rmax && !found; r++)
            for (int c = cmin; c + 2 <= cmax; c++) {
                const int32_t *x = &b->colour[r * C + c];
                if (x[0] == x[1] && x[1] == x[2] && b->type[r * C + c + 2] >= 0) { found = 1; break; }
            }
    }
    if (!found && rmin + 2 <= rmax) {                          /* ref :777-784 */
        for (int r = rmin; r + 2 <= rmax && !found; r++)
            for (int c = cmin; c <= cmax; c++) {
                int i = r * C + c;
                if (b->colour[i] == b->colour[i + C] && b->colour[i + C] == b->colour[i + 2 * C] && b->type[i + 2 * C] >= 0) { found = 1; break; }
            }
    }
    swap_cells(b, i1, i2);                                     /* ref :773,783,786 */
    return found;
}

/* ref board.py:243-266 */
int tmgo_is_move_legal(const tmgo_board *b, int r1, int c1, int r2, int c2) {
    if (!(0 <= r1 && r1 < b->R && 0 <= c1 && c1 < b->C)) return 0;
    if (!(0 <= r2 && r2 < b->R && 0 <= c2 && c2 < b->C)) return 0;
    if (r1 == r2 && c1 == c2) return 0;
    int dr = abs(r1 - r2), dc = abs(c1 - c2);
    if (!(r1 == r2 || c1 == c2) || dr > 1 || dc > 1) return 0;
    return 1;
}

/* ref board.py:558-569 */
int tmgo_possible_move(tmgo_board *b) {
    int A = tmgo_num_actions(b);
    for (int a = 0; a < A; a++) {
        int r1, c1, r2, c2;
        tmgo_action_to_coords(b, a, &r1, &c1, &r2, &c2);
        if (tmgo_is_move_effective(b, r1, c1, r2, c2)) return 1;
    }
    return 0;
}

void tmgo_effective_mask(tmgo_board *b, uint8_t *out) { /* ref tile_match_env.py:122-123 */
    int A = tmgo_num_actions(b);
    for (int a = 0; a < A; a++) {
        int r1, c1, r2, c2;
        tmgo_action_to_coords(b, a, &r1, &c1, &r2, &c2);
        out[a] = (uint8_t)tmgo_is_move_effective(b, r1, c1, r2, c2);
    }
}

/* -------------------------------------------------------------------------------------------- */
/* get_colour_lines (ref board.py:149-215)                                                       */
/* -------------------------------------------------------------------------------------------- */
static void get_colour_lines(tmgo_board *b) {
    int R = b->R, C = b->C;
    linelist_t *L = &b->lines;
    ll_clear(L);
    uint8_t *hset = b->flagB;                       /* ref :156 (vertical_line_coords can never hit: one row only) */
    uint8_t *vset = b->flagA;
    memset(vset, 0, (size_t)b->P);
    memset(hset, 0, (size_t)b->P);
    int found = 0;
    for (int row = R - 1; row >= 0; row--) {        /* ref :158 */
        if (found) break;                           /* ref :159-160 */
        for (int col = 0; col < C; col++) {         /* ref :161 */
            int i = row * C + col;
            if (1 < row && !vset[i]) {              /* ref :163 */
                if (b->type[i] > 0) {               /* ref :164 */
                    if (b->colour[i] == b->colour[i - C]) { /* ref :165 */
                        int start = row - 1, end = row;
                        while (start > 0) {         /* ref :168-172 */
                            if (b->colour[i] == b->colour[(start - 1) * C + col]) start--;
                            else break;
                        }
                        if (end - start >= 2) {     /* ref :173-177 */
                            found = 1;
                            line_t *ln = ll_new(L);
                            if (!ln) { b->status |= TMGO_ST_INTERNAL; return; }
                            for (int r = start; r <= end; r++) { ln->cells[ln->n++] = r * C + col; vset[r * C + col] = 1; }
                        }
                    }
                }
            }
            if (col < C - 2 && !hset[i]) {          /* ref :179 */
                if (b->type[i] > 0) {               /* ref :180 */
                    if (b->colour[i] == b->colour[i + 1]) { /* ref :181 */
                        int start = col, end = col + 1;
                        while (end < C - 1) {       /* ref :184-188 */
                            if (b->colour[i] == b->colour[row * C + end + 1]) end++;
                            else break;
                        }
                        if (end - start >= 2) {     /* ref :189-193 */
                            found = 1;
                            line_t *ln = ll_new(L);
                            if (!ln) { b->status |= TMGO_ST_INTERNAL; return; }
                            for (int c = start; c <= end; c++) { ln->cells[ln->n++] = row * C + c; hset[row * C + c] = 1; }
                        }
                    }
                }
            }
        }
    }
    /* phase 2 (ref :198-214): coords is a snapshot list of all phase-1 cells, in list order */
    uint8_t *incoords = b->flagC;
    memset(incoords, 0, (size_t)b->P);
    int32_t *coords = b->tmp;
    int ncoords = 0;
    for (int li = 0; li < L->n; li++)
        for (int k = 0; k < L->l[li]->n; k++) {
            coords[ncoords++] = L->l[li]->cells[k]; /* duplicates kept, as in the reference list */
            incoords[L->l[li]->cells[k]] = 1;
        }
    static const int DR[4] = {0, 1, 0, -1}, DC[4] = {1, 0, -1, 0}; /* ref :201 */
    for (int ci = 0; ci < ncoords; ci++) {          /* ref :203 */
        int c0 = coords[ci], r0 = c0 / C, q0 = c0 % C;
        for (int d = 0; d < 4; d++) {               /* ref :204 */
            line_t ln;
            ln.n = 0;
            ln.cells[ln.n++] = c0;                  /* ref :205 */
            for (int s = 0; s < 2; s++) {           /* ref :206: [d, -d] */
                int dr = s ? -DR[d] : DR[d], dc = s ? -DC[d] : DC[d];
                int nr = r0 + dr, nc = q0 + dc;
                for (;;) {                          /* ref :208 */
                    int valid = 0 <= nr && nr < R && 0 <= nc && nc < C;
                    if (valid && incoords[nr * C + nc]) break;      /* n not in coords */
                    if (!valid) break;
                    int ni = nr * C + nc;
                    if (!(b->colour[c0] == b->colour[ni] && b->type[c0] > 0 && b->type[ni] > 0)) break; /* ref :199 */
                    if (ln.n < MAXLEN) ln.cells[ln.n++] = ni;
                    nr += dr; nc += dc;
                }
            }
            if (ln.n >= 3) {                        /* ref :211-214 */
                line_sort(&ln);
                int present = 0;
                for (int li = 0; li < L->n; li++) if (line_eq(L->l[li], &ln)) { present = 1; break; }
                if (!present) {
                    line_t *nl = ll_new(L);
                    if (!nl) { b->status |= TMGO_ST_INTERNAL; return; }
                    *nl = ln;
                }
            }
        }
    }
    if (L->n > b->max_lines) b->max_lines = L->n;
}

static int export_lists(const linelist_t *L, int32_t *cells, int32_t *offsets, int max_cells, int max_lines) {
    int o = 0;
    if (L->n > max_lines) return -1;
    for (int i = 0; i < L->n; i++) {
        offsets[i] = o;
        for (int k = 0; k < L->l[i]->n; k++) { if (o >= max_cells) return -1; cells[o++] = L->l[i]->cells[k]; }
    }
    offsets[L->n] = o;
    return L->n;
}

int tmgo_get_colour_lines(tmgo_board *b, int32_t *cells, int32_t *offsets, int max_cells, int max_lines) {
    get_colour_lines(b);
    return export_lists(&b->lines, cells, offsets, max_cells, max_lines);
}

/* -------------------------------------------------------------------------------------------- */
/* process_colour_lines (ref board.py:269-327); consumes b->lines, fills b->matches               */
/* -------------------------------------------------------------------------------------------- */
static void process_colour_lines(tmgo_board *b) {
    int C = b->C;
    linelist_t *L = &b->lines, *M = &b->matches;
    ll_clear(M);
    /* ref :282: sort each line by (row,col); stable-sort the list by the first cell's row */
    for (int i = 0; i < L->n; i++) line_sort(L->l[i]);
    for (int i = 1; i < L->n; i++) {
        line_t *v = L->l[i];
        int j = i - 1;
        while (j >= 0 && L->l[j]->cells[0] / C > v->cells[0] / C) { L->l[j + 1] = L->l[j]; j--; }
        L->l[j + 1] = v;
    }
    while (L->n > 0) {                              /* ref :284 */
        line_t *line = L->l[0];                     /* ref :285 pop(0) */
        ll_remove_at(L, 0);
        int bomb_hit = 0;
        if (line->n >= 5 && (b->specials & TMGO_SP_COOKIE)) { /* ref :287-292 */
            line_t *m = ll_new(M);
            if (!m) { b->status |= TMGO_ST_INTERNAL; return; }
            for (int k = 0; k < 5; k++) m->cells[m->n++] = line->cells[k];
            b->match_name[M->n - 1] = NAME_COOKIE;
            b->match_colour[M->n - 1] = 0;
            if (line->n - 5 > 2) {                  /* ref :291-292: remainder re-queued at the back */
                line_t *rest = ll_new(L);
                if (!rest) { b->status |= TMGO_ST_INTERNAL; return; }
                for (int k = 5; k < line->n; k++) rest->cells[rest->n++] = line->cells[k];
            }
            continue;
        }
        if (line->n == 4) {                         /* ref :294-302 */
            line_t *m = ll_new(M);
            if (!m) { b->status |= TMGO_ST_INTERNAL; return; }
            *m = *line;
            b->match_colour[M->n - 1] = b->colour[line->cells[0]];
            if (line->cells[0] / C == line->cells[1] / C && (b->specials & TMGO_SP_HLASER)) b->match_name[M->n - 1] = NAME_HLASER;
            else if (b->specials & TMGO_SP_VLASER) b->match_name[M->n - 1] = NAME_VLASER;
            else b->match_name[M->n - 1] = NAME_NORMAL;
            continue;
        }
        if (b->specials & TMGO_SP_BOMB) {           /* ref :304: any(coord in l for coord in line for l in lines) */
            for (int li = 0; li < L->n && !bomb_hit; li++)
                for (int k = 0; k < line->n; k++) if (line_has(L->l[li], line->cells[k])) { bomb_hit = 1; break; }
        }
        if (bomb_hit) {
            for (int li = 0; li < L->n; li++) {     /* ref :305 */
                line_t *l = L->l[li];
                int shared = -1;                    /* ref :306-308: first cell of `line` that is in l */
                for (int k = 0; k < line->n; k++) if (line_has(l, line->cells[k])) { shared = line->cells[k]; break; }
                if (shared < 0) continue;
                int sr = shared / C, sc = shared % C;
                /* ref :310: stable sort of l by Manhattan distance to shared */
                line_t sorted = *l;
                for (int i = 1; i < sorted.n; i++) {
                    int v = sorted.cells[i];
                    int dv = abs(v / C - sr) + abs(v % C - sc);
                    int j = i - 1;
                    while (j >= 0) {
                        int u = sorted.cells[j];
                        int du = abs(u / C - sr) + abs(u % C - sc);
                        if (du > dv) { sorted.cells[j + 1] = u; j--; } else break;
                    }
                    sorted.cells[j + 1] = v;
                }
                int take = sorted.n < 3 ? sorted.n : 3;
                line_t *m = ll_new(M);              /* ref :312 */
                if (!m) { b->status |= TMGO_ST_INTERNAL; return; }
                for (int k = 0; k < line->n; k++) m->cells[m->n++] = line->cells[k];
                for (int k = 0; k < take; k++) if (!line_has(line, sorted.cells[k])) m->cells[m->n++] = sorted.cells[k];
                b->match_name[M->n - 1] = NAME_BOMB; /* ref :313 */
                b->match_colour[M->n - 1] = b->colour[line->cells[0]]; /* ref :314 */
                if (l->n < 6) {                     /* ref :315-316: lines.remove(l) = first element equal by value */
                    for (int q = 0; q < L->n; q++) if (line_eq(L->l[q], l)) { ll_remove_at(L, q); break; }
                } else {                            /* ref :317-319 */
                    for (int k = 0; k < take; k++) {
                        int cell = sorted.cells[k];
                        for (int q = 0; q < l->n; q++) if (l->cells[q] == cell) {
                            memmove(&l->cells[q], &l->cells[q + 1], sizeof(int) * (size_t)(l->n - q - 1));
                            l->n--;
                            break;
                        }
                    }
                }
                break;                              /* ref :320 */
            }
            continue;
        }
        if (line->n >= 3) {                         /* ref :322-325 */
            line_t *m = ll_new(M);
            if (!m) { b->status |= TMGO_ST_INTERNAL; return; }
            *m = *line;
            b->match_name[M->n - 1] = NAME_NORMAL;
            b->match_colour[M->n - 1] = b->colour[line->cells[0]];
        }
    }
}

/* ref board.py:133-147 */
static void detect_colour_matches(tmgo_board *b) {
    get_colour_lines(b);
    if (b->lines.n == 0) { ll_clear(&b->matches); return; }
    process_colour_lines(b);
}

int tmgo_detect_colour_matches(tmgo_board *b, int32_t *cells, int32_t *offsets, int32_t *names, int32_t *colours,
                               int max_cells, int max_matches) {
    detect_colour_matches(b);
    int n = export_lists(&b->matches, cells, offsets, max_cells, max_matches);
    for (int i = 0; i < n; i++) { names[i] = b->match_name[i]; colours[i] = b->match_colour[i]; }
    return n;
}

/* -------------------------------------------------------------------------------------------- */
/* activation (ref board.py:473-556)                                                             */
/* -------------------------------------------------------------------------------------------- */
static void activate_special(tmgo_board *b, int r0, int c0, int tile_type, int is_comb) {
    int R = b->R, C = b->C, P = b->P;
    int all_zero = 1;                               /* ref :488-489 */
    for (int i = 0; i < P; i++) if (b->colour[i] != 0) { all_zero = 0; break; }
    if (all_zero) return;
    if (tile_type == 0 || tile_type == 1) { b->status |= TMGO_ST_INTERNAL; return; } /* ref :491-492 raises */
    b->depth++;
    if (b->depth > b->max_depth) b->max_depth = b->depth;
    b->colour[r0 * C + c0] = 0;                     /* ref :496 */
    b->type[r0 * C + c0] = 0;
    if (!is_comb) b->num_specials_activated++;      /* ref :498-499 */
    if (tile_type == 2) {                           /* ref :502-507 */
        for (int row = 0; row < R; row++) {
            int i = row * C + c0;
            if (not01(b->type[i])) activate_special(b, row, c0, b->type[i], 0);
            else { b->colour[i] = 0; b->type[i] = 0; }
        }
    } else if (tile_type == 3) {                    /* ref :510-515 */
        for (int col = 0; col < C; col++) {
            int i = r0 * C + col;
            if (not01(b->type[i])) activate_special(b, r0, col, b->type[i], 0);
            else { b->colour[i] = 0; b->type[i] = 0; }
        }
    } else if (tile_type == 4) {                    /* ref :517-528 */
        int min_r = r0 - 1 > 0 ? r0 - 1 : 0, max_r = r0 + 1 < R - 1 ? r0 + 1 : R - 1;
        int min_c = c0 - 1 > 0 ? c0 - 1 : 0, max_c = c0 + 1 < C - 1 ? c0 + 1 : C - 1;
        for (int i = min_r; i <= max_r; i++)
            for (int j = min_c; j <= max_c; j++) {
                int q = i * C + j;
                if (not01(b->type[q])) activate_special(b, i, j, b->type[q], 0);
                else { b->colour[q] = 0; b->type[q] = 0; }
            }
    } else if (tile_type == -1) {                   /* ref :530-554 */
        int maxc = 0;
        for (int i = 0; i < P; i++) if (b->colour[i] > maxc) maxc = b->colour[i];
        if (maxc == 0) { b->depth--; return; }      /* ref :532-534 */
        int *counts = (int *)calloc((size_t)maxc + 1, sizeof(int)); /* ref :536 bincount over non-zero colours */
        for (int i = 0; i < P; i++) if (b->colour[i] != 0) counts[b->colour[i]]++;
        int mc = 0;                                 /* ref :537 argmax = first maximum */
        for (int k = 1; k <= maxc; k++) if (counts[k] > counts[mc]) mc = k;
        free(counts);
        uint8_t *cmask = (uint8_t *)malloc((size_t)P); /* local: recursion re-enters this branch */
        for (int i = 0; i < P; i++) cmask[i] = b->colour[i] == mc; /* ref :540 snapshot */
        for (int i = 0; i < P; i++) if (cmask[i] && b->type[i] == 1) { b->colour[i] = 0; b->type[i] = 0; } /* ref :541-544 */
        for (int i = 0; i < P; i++) cmask[i] = cmask[i] && b->type[i] > 1; /* ref :547-549 snapshot, row-major */
        for (int i = 0; i < P; i++)
            if (cmask[i] && not01(b->type[i])) activate_special(b, i / C, i % C, b->type[i], 0); /* ref :551-554 */
        free(cmask);
    } else {
        b->status |= TMGO_ST_INTERNAL;              /* ref :555-556 raises */
    }
    b->depth--;
}

void tmgo_activate_special(tmgo_board *b, int r, int c, int tile_type, int is_comb) { activate_special(b, r, c, tile_type, is_comb); }

/* ref board.py:721-726 */
static void activate_specials_in_mask(tmgo_board *b, const uint8_t *mask, int is_comb) {
    for (int i = 0; i < b->P; i++)
        if (mask[i] && not01(b->type[i])) activate_special(b, i / b->C, i % b->C, b->type[i], is_comb);
}

/* -------------------------------------------------------------------------------------------- */
/* combination_match (ref board.py:600-719)                                                      */
/* -------------------------------------------------------------------------------------------- */
static void combination_match(tmgo_board *b, int r1, int c1, int r2, int c2) {
    int R = b->R, C = b->C, P = b->P;
    b->num_specials_activated += 2;                 /* ref :609 */
    int i1 = r1 * C + c1, i2 = r2 * C + c2;
    int t1 = b->type[i1], k1 = b->colour[i1], t2 = b->type[i2], k2 = b->colour[i2]; /* ref :611-612 */
    if (t1 == -1 && t2 == -1) {                     /* ref :615-616 */
        memset(b->colour, 0, 4 * (size_t)P);
        memset(b->type, 0, 4 * (size_t)P);
    } else if ((t1 == -1 && t2 == 1) || (t1 == 1 && t2 == -1)) { /* ref :619-641 */
        if (t1 == 1) { int t; t = t1; t1 = t2; t2 = t; t = k1; k1 = k2; k2 = t; t = i1; i1 = i2; i2 = t; }
        b->colour[i1] = 0; b->type[i1] = 0;         /* ref :626 and (again) :628 */
        uint8_t *cmask = (uint8_t *)malloc((size_t)P), *smask = (uint8_t *)malloc((size_t)P);
        for (int i = 0; i < P; i++) cmask[i] = b->colour[i] == k2;     /* ref :631 */
        for (int i = 0; i < P; i++) if (cmask[i] && b->type[i] == 1) { b->colour[i] = 0; b->type[i] = 0; } /* ref :632-635 */
        for (int i = 0; i < P; i++) smask[i] = cmask[i] && b->type[i] > 1; /* ref :638-639 */
        activate_specials_in_mask(b, smask, 1);     /* ref :640 */
        b->num_specials_activated -= 1;             /* ref :641 */
        free(cmask); free(smask);
    } else if ((t1 == -1 && t2 >= 2) || (t1 >= 2 && t2 == -1)) { /* ref :644-660 */
        if (t2 == -1) { int t; t = t1; t1 = t2; t2 = t; t = k1; k1 = k2; k2 = t; t = i1; i1 = i2; i2 = t; }
        b->colour[i1] = 0; b->type[i1] = 0;         /* ref :651 */
        uint8_t *cmask = (uint8_t *)malloc((size_t)P);
        for (int i = 0; i < P; i++) cmask[i] = b->colour[i] == k2;     /* ref :654 */
        for (int i = 0; i < P; i++) if (cmask[i] && b->type[i] == 1) b->type[i] = t2; /* ref :655-657 */
        activate_specials_in_mask(b, cmask, 1);     /* ref :660 */
        free(cmask);
    } else if ((t1 == 2 || t1 == 3) && (t2 == 2 || t2 == 3)) {  /* ref :663-674 */
        b->colour[i1] = 0; b->type[i1] = 0;
        b->colour[i2] = 0; b->type[i2] = 0;
        int r = r1 < r2 ? r1 : r2, c = c1 < c2 ? c1 : c2;
        activate_special(b, r, c, 2, 1);
        activate_special(b, r, c, 3, 1);
    } else if ((t1 == 4 && t2 >= 2 && t2 <= 3) || (t2 == 4 && t1 >= 2 && t1 <= 3)) { /* ref :677-696 */
        b->colour[i1] = 0; b->type[i1] = 0;
        b->colour[i2] = 0; b->type[i2] = 0;
        int r = r1 < r2 ? r1 : r2, c = c1 < c2 ? c1 : c2;
        int min_r = r - 1 > 0 ? r - 1 : 0, max_r = r + 1 < R - 1 ? r + 1 : R - 1;
        int min_c = c - 1 > 0 ? c - 1 : 0, max_c = c + 1 < C - 1 ? c + 1 : C - 1;
        for (int i = min_r; i <= max_r; i++) activate_special(b, i, c, 3, 1);
        for (int j = min_c; j <= max_c; j++) activate_special(b, r, j, 2, 1);
    } else if (t1 == 4 && t2 == 4) {                /* ref :699-719 */
        b->colour[i1] = 0; b->type[i1] = 0;
        b->colour[i2] = 0; b->type[i2] = 0;
        int r = r1 < r2 ? r1 : r2, c = c1 < c2 ? c1 : c2;
        int min_r = r - 2 > 0 ? r - 2 : 0, max_r = r + 2 < R - 1 ? r + 2 : R - 1;
        int min_c = c - 2 > 0 ? c - 2 : 0, max_c = c + 2 < C - 1 ? c + 2 : C - 1;
        for (int i = min_r; i <= max_r; i++)
            for (int j = min_c; j <= max_c; j++) {
                int q = i * C + j;
                if (b->type[q] == 1) { b->colour[q] = 0; b->type[q] = 0; }
                else if (b->type[q] != 0) activate_special(b, i, j, b->type[q], 1);
            }
    }
}

void tmgo_combination_match(tmgo_board *b, int r1, int c1, int r2, int c2) { combination_match(b, r1, c1, r2, c2); }

/* -------------------------------------------------------------------------------------------- */
/* resolve (ref board.py:397-471, 572-597)                                                       */
/* -------------------------------------------------------------------------------------------- */
/* ref :429-458 */
static int get_special_creation_pos(tmgo_board *b, const line_t *coords, const uint8_t *taken, int straight) {
    int C = b->C;
    line_t valid;
    valid.n = 0;
    for (int k = 0; k < coords->n; k++) if (!taken[coords->cells[k]]) valid.cells[valid.n++] = coords->cells[k]; /* ref :439 */
    if (valid.n == 0) { b->status |= TMGO_ST_INTERNAL; return -1; } /* reference would raise IndexError */
    if (!straight) {                                /* ref :441-450 */
        int best_r = -1, best_rc = 0, best_c = -1, best_cc = 0;
        for (int k = 0; k < coords->n; k++) {       /* max(xs, key=xs.count): first element with the highest count */
            int rr = coords->cells[k] / C, cc = coords->cells[k] % C, nr = 0, nc = 0;
            for (int q = 0; q < coords->n; q++) { nr += coords->cells[q] / C == rr; nc += coords->cells[q] % C == cc; }
            if (nr > best_rc) { best_rc = nr; best_r = rr; }
            if (nc > best_cc) { best_cc = nc; best_c = cc; }
        }
        int corner = best_r * C + best_c;
        if (line_has(&valid, corner)) return corner; /* ref :446-447 */
        int best = -1, bestd = 0;                   /* ref :449: stable sort by squared distance, take first */
        for (int k = 0; k < valid.n; k++) {
            int dr = valid.cells[k] / C - best_r, dc = valid.cells[k] % C - best_c, d = dr * dr + dc * dc;
            if (best < 0 || d < bestd) { best = valid.cells[k]; bestd = d; }
        }
        return best;
    }
    line_sort(&valid);                              /* ref :453 */
    if (valid.n % 2 == 0) return valid.cells[valid.n / 2 - 1]; /* ref :454-456 */
    return valid.cells[valid.n / 2];                /* ref :457 */
}

static void resolve_colour_matches(tmgo_board *b) {
    int C = b->C;
    linelist_t *M = &b->matches;
    uint8_t *taken = b->flagA;
    memset(taken, 0, (size_t)b->P);
    int nq = 0;
    int *q_pos = (int *)malloc(sizeof(int) * (size_t)(M->n + 1)); /* special_creation_q, ref :411 */
    int *q_col = (int *)malloc(sizeof(int) * (size_t)(M->n + 1));
    int *q_nm = (int *)malloc(sizeof(int) * (size_t)(M->n + 1));
    for (int i = 0; i < M->n; i++) {                /* ref :414-418 */
        if (b->match_name[i] != NAME_NORMAL) {
            int pos = get_special_creation_pos(b, M->l[i], taken, b->match_name[i] != NAME_BOMB);
            if (pos >= 0) taken[pos] = 1;
            q_pos[nq] = pos; q_nm[nq] = b->match_name[i]; q_col[nq] = b->match_colour[i];
            nq++;
        }
    }
    for (int i = 0; i < M->n; i++) {                /* ref :421-423 -> resolve_colour_match :460-471 */
        const line_t *m = M->l[i];
        for (int k = 0; k < m->n; k++) {
            int cell = m->cells[k];
            if (not01(b->type[cell])) activate_special(b, cell / C, cell % C, b->type[cell], 0);
            else { b->colour[cell] = 0; b->type[cell] = 0; }
        }
    }
    for (int i = 0; i < nq; i++) {                  /* ref :426-427 -> create_special :572-597 */
        b->num_new_specials++;
        if (q_pos[i] < 0) continue;
        int t = q_nm[i] == NAME_COOKIE ? -1 : q_nm[i] == NAME_VLASER ? 2 : q_nm[i] == NAME_HLASER ? 3 : 4;
        b->colour[q_pos[i]] = q_col[i];
        b->type[q_pos[i]] = t;
    }
    free(q_pos); free(q_col); free(q_nm);
}

/* direct access for the reference's tests of get_special_creation_pos (tests/board/test_activation.py:437-543) */
int tmgo_special_creation_pos(tmgo_board *b, const int32_t *cells, int n, const int32_t *taken_cells, int ntaken, int straight) {
    line_t ln;
    ln.n = 0;
    for (int i = 0; i < n && i < MAXLEN; i++) ln.cells[ln.n++] = cells[i];
    memset(b->flagA, 0, (size_t)b->P);
    for (int i = 0; i < ntaken; i++) b->flagA[taken_cells[i]] = 1;
    return get_special_creation_pos(b, &ln, b->flagA, straight);
}

int tmgo_resolve_round(tmgo_board *b) {
    detect_colour_matches(b);
    int n = b->matches.n;
    if (n > 0) resolve_colour_matches(b);
    return n;
}

/* -------------------------------------------------------------------------------------------- */
/* gravity / refill / shuffle / remove_colour_lines / generate_board                             */
/* -------------------------------------------------------------------------------------------- */
void tmgo_gravity(tmgo_board *b) {                  /* ref board.py:217-229 */
    int R = b->R, C = b->C;
    for (int j = 0; j < C; j++) {
        int w = R - 1;
        for (int r = R - 1; r >= 0; r--) {          /* non-empty cells keep their order, packed to the bottom */
            int i = r * C + j;
            if (!(b->colour[i] == 0 && b->type[i] == 0)) {
                int32_t k = b->colour[i], t = b->type[i];
                b->colour[w * C + j] = k; b->type[w * C + j] = t;
                w--;
            }
        }
        for (; w >= 0; w--) { b->colour[w * C + j] = 0; b->type[w * C + j] = 0; }
    }
}

void tmgo_refill(tmgo_board *b) {                   /* ref board.py:231-241 */
    int n = 0;
    for (int i = 0; i < b->P; i++) n += (b->colour[i] == 0 && b->type[i] == 0);
    if (n > 0) {
        int32_t *vals = b->tmp;
        draw_colours(b, n, vals);
        int k = 0;
        for (int i = 0; i < b->P; i++)              /* boolean-mask assignment = row-major order */
            if (b->colour[i] == 0 && b->type[i] == 0) { b->colour[i] = vals[k++]; b->type[i] = 1; }
    }
}

void tmgo_shuffle(tmgo_board *b) {                  /* ref board.py:114-118 with the project's Fisher-Yates */
    int P = b->P;
    int32_t *idx = b->tmp;
    for (int i = 0; i < P; i++) idx[i] = i;
    for (int i = P - 1; i >= 1; i--) {
        uint32_t w = b->in_reset ? reset_word(b, 4, b->rsc++) : board_word(b, 1, b->shuffle_cursor++);
        int j = (int)(((uint64_t)w * (uint64_t)(i + 1)) >> 32);
        int32_t t = idx[i]; idx[i] = idx[j]; idx[j] = t;
    }
    int32_t *oc = (int32_t *)malloc(4 * (size_t)P), *ot = (int32_t *)malloc(4 * (size_t)P);
    memcpy(oc, b->colour, 4 * (size_t)P);
    memcpy(ot, b->type, 4 * (size_t)P);
    for (int i = 0; i < P; i++) { b->colour[i] = oc[idx[i]]; b->type[i] = ot[idx[i]]; } /* ref :118 */
    free(oc); free(ot);
}

/* ref board.py:120-131.  returns 0 if the iteration cap stopped it */
static int remove_colour_lines(tmgo_board *b, int64_t *iters) {
    int R = b->R, C = b->C;
    while (b->lines.n > 0) {                        /* ref :126 */
        if (b->iter_cap > 0 && *iters >= b->iter_cap) { b->status |= TMGO_ST_RESET_CAP; return 0; }
        (*iters)++;
        int top = b->lines.l[0]->cells[0] / C;      /* ref :127-128: l[0][0] */
        int row = top + 1 < R - 1 ? top + 1 : R - 1;
        int n = (row + 1) * C;
        draw_colours(b, n, b->tmp);                 /* ref :129 */
        for (int i = 0; i < n; i++) b->colour[i] = b->tmp[i];
        get_colour_lines(b);                        /* ref :130 */
    }
    return 1;
}

/* shared tail of generate_board (ref :102-109) and move (ref :381-391) */
static int playability_loop(tmgo_board *b, int have_lines) {
    int shuffled = 0;
    int64_t iters = 0;
    if (!have_lines) ll_clear(&b->lines);
    while (!tmgo_possible_move(b) || b->lines.n > 0) {
        if (b->lines.n > 0) {
            if (!remove_colour_lines(b, &iters)) break;
        } else {
            if (b->iter_cap > 0 && iters >= b->iter_cap) { b->status |= TMGO_ST_RESET_CAP; break; }
            iters++;
            shuffled = 1;
            tmgo_shuffle(b);
        }
        get_colour_lines(b);
    }
    if (iters > b->reset_iters) b->reset_iters = iters;
    return shuffled;
}

void tmgo_board_set_constructive(tmgo_board *b, int on) { b->constructive = on; }

/* NOT the reference's algorithm: the product's TMG_FLAG_CONSTRUCTIVE_RESET contract (include/tmg_b200.h), restated.
 * Cells in row-major order; each takes the next colour of stream 5 of the board that does not complete a triple with the
 * two cells to its left or the two above it (at most 64 draws); a board without a possible move is drawn again. */
static void generate_constructive(tmgo_board *b) {
    int R = b->R, C = b->C;
    uint64_t k = 0;
    int64_t attempts = 0;
    for (;;) {
        for (int i = 0; i < b->P; i++) b->type[i] = 1;
        for (int r = 0; r < R; r++)
            for (int c = 0; c < C; c++) {
                int fv = (r >= 2 && b->colour[(r - 1) * C + c] == b->colour[(r - 2) * C + c]) ? b->colour[(r - 1) * C + c] : -1;
                int fh = (c >= 2 && b->colour[r * C + c - 1] == b->colour[r * C + c - 2]) ? b->colour[r * C + c - 1] : -1;
                int col = 1;
                for (int tries = 0;; tries++) {
                    col = 1 + (int)(((uint64_t)reset_word(b, 5, k++) * (uint64_t)b->K) >> 32);
                    if (col != fv && col != fh) break;
                    if (tries >= 63) { b->status |= TMGO_ST_RESET_CAP; break; }
                }
                b->colour[r * C + c] = col;
            }
        if (tmgo_possible_move(b)) break;
        if (++attempts >= (b->iter_cap > 0 ? b->iter_cap : 16384)) { b->status |= TMGO_ST_RESET_CAP; break; }
    }
}

void tmgo_generate_board(tmgo_board *b) {           /* ref board.py:95-112 */
    b->episode++;
    if (b->constructive) { generate_constructive(b); return; }
    b->in_reset = !b->use_inj;                      /* injected draws are one sequential stream */
    b->rdc = b->rsc = 0;
    for (int i = 0; i < b->P; i++) b->type[i] = 1;
    draw_colours(b, b->P, b->colour);
    get_colour_lines(b);
    playability_loop(b, 1);
    b->in_reset = 0;
}

/* -------------------------------------------------------------------------------------------- */
/* move (ref board.py:330-395)                                                                   */
/* -------------------------------------------------------------------------------------------- */
static int count_empty_type(const tmgo_board *b) {  /* flat_size - count_nonzero(board[1]) */
    int n = 0;
    for (int i = 0; i < b->P; i++) n += b->type[i] == 0;
    return n;
}

int tmgo_move(tmgo_board *b, int r1, int c1, int r2, int c2, int32_t out[5]) {
    b->num_specials_activated = 0;                  /* ref :343-347 */
    b->num_new_specials = 0;
    int num_elim = 0, is_comb = 0, shuffled = 0;
    out[0] = out[1] = out[2] = out[3] = out[4] = 0;
    if (!tmgo_is_move_legal(b, r1, c1, r2, c2)) return -1; /* ref :349-350 */
    if (!tmgo_is_move_effective(b, r1, c1, r2, c2)) return 0; /* ref :352-353 */
    int i1 = r1 * b->C + c1, i2 = r2 * b->C + c2;
    swap_cells(b, i1, i2);                          /* ref :355 */
    int two = not01(b->type[i1]) && not01(b->type[i2]); /* ref :357 */
    int one = b->type[i1] < 0 || b->type[i2] < 0;       /* ref :358 */
    if (two || one) {                               /* ref :359-364 */
        is_comb = 1;
        combination_match(b, r1, c1, r2, c2);
        num_elim += count_empty_type(b);
        tmgo_gravity(b);
        tmgo_refill(b);
    }
    for (;;) {                                      /* ref :367-376 */
        detect_colour_matches(b);
        if (b->matches.n == 0) break;
        resolve_colour_matches(b);
        num_elim += count_empty_type(b);
        tmgo_gravity(b);
        tmgo_refill(b);
    }
    num_elim += b->num_new_specials;                /* ref :378 */
    shuffled = playability_loop(b, 0);              /* ref :381-391 */
    out[0] = num_elim; out[1] = is_comb; out[2] = b->num_new_specials; out[3] = b->num_specials_activated; out[4] = shuffled;
    return 0;
}

/* -------------------------------------------------------------------------------------------- */
/* one-hot (ref wrappers.py:40-46,54-69)                                                         */
/* -------------------------------------------------------------------------------------------- */
int tmgo_onehot_planes(const tmgo_board *b) {
    return b->K + !!(b->specials & TMGO_SP_COOKIE) + !!(b->specials & TMGO_SP_VLASER) + !!(b->specials & TMGO_SP_HLASER) + !!(b->specials & TMGO_SP_BOMB);
}
void tmgo_onehot(const tmgo_board *b, uint8_t *out) {
    int P = b->P, K = b->K;
    memset(out, 0, (size_t)tmgo_onehot_planes(b) * (size_t)P);
    for (int i = 0; i < P; i++) {                   /* colour_ohe[colour] = 1, plane 0 dropped (ref :57-59) */
        int k = b->colour[i];
        if (k >= 1 && k <= K) out[(k - 1) * P + i] = 1;
    }
    /* type_ohe index = type+1; kept slices sorted by type id: cookie(-1), v(2), h(3), bomb(4) (ref :40-46,63-67) */
    static const int tid[4] = {-1, 2, 3, 4};
    static const uint32_t bit[4] = {TMGO_SP_COOKIE, TMGO_SP_VLASER, TMGO_SP_HLASER, TMGO_SP_BOMB};
    int plane = K;
    for (int s = 0; s < 4; s++) {
        if (!(b->specials & bit[s])) continue;
        for (int i = 0; i < P; i++) if (b->type[i] == tid[s]) out[plane * P + i] = 1;
        plane++;
    }
}

/* -------------------------------------------------------------------------------------------- */
/* vectorised env (semantics of include/tmg_b200.h; ref tile_match_env.py:84-124)                */
/* -------------------------------------------------------------------------------------------- */
struct tmgo_vec {
    tmgo_vec_config cfg;
    int A, P;
    tmgo_vec_buffers buf;
    const uint8_t *inj;
    int64_t inj_len;
    tmgo_board **worker; /* one scratch board per thread */
    int nthreads;
    int max_lines, max_depth;
    int64_t max_reset_iters;
};

tmgo_vec *tmgo_vec_create(const tmgo_vec_config *cfg) {
    tmgo_vec *v = (tmgo_vec *)calloc(1, sizeof(*v));
    v->cfg = *cfg;
    int N = cfg->num_envs, R = cfg->num_rows, C = cfg->num_cols;
    v->P = R * C;
    v->A = 2 * R * C - R - C;
    v->nthreads = cfg->num_threads > 0 ? cfg->num_threads : 1;
    v->buf.board = (int8_t *)calloc((size_t)N * 2 * (size_t)v->P, 1);
    v->buf.timer = (int32_t *)malloc(4 * (size_t)N);
    for (int i = 0; i < N; i++) v->buf.timer[i] = -1; /* tile_match_env.py:75 timer = None */
    v->buf.draw_cursor = (uint64_t *)calloc((size_t)N, 8);
    v->buf.shuffle_cursor = (uint64_t *)calloc((size_t)N, 8);
    v->buf.episode = (int32_t *)malloc(4 * (size_t)N);
    for (int i = 0; i < N; i++) v->buf.episode[i] = -1;
    v->buf.reward = (int32_t *)calloc((size_t)N, 4);
    v->buf.terminated = (uint8_t *)calloc((size_t)N, 1);
    v->buf.is_combination_match = (uint8_t *)calloc((size_t)N, 1);
    v->buf.num_new_specials = (int32_t *)calloc((size_t)N, 4);
    v->buf.num_specials_activated = (int32_t *)calloc((size_t)N, 4);
    v->buf.shuffled = (uint8_t *)calloc((size_t)N, 1);
    v->buf.mask = (uint8_t *)calloc((size_t)N * (size_t)v->A, 1);
    v->buf.num_moves_left = (int32_t *)calloc((size_t)N, 4);
    v->buf.status = (uint32_t *)calloc((size_t)N, 4);
    v->worker = (tmgo_board **)calloc((size_t)v->nthreads, sizeof(tmgo_board *));
    for (int t = 0; t < v->nthreads; t++) {
        v->worker[t] = tmgo_board_create(R, C, cfg->num_colours, cfg->specials);
        if (v->worker[t]) v->worker[t]->constructive = (cfg->flags & TMGO_FLAG_CONSTRUCTIVE_RESET) != 0;
        tmgo_board_set_iter_cap(v->worker[t], cfg->max_reset_iters);
    }
    return v;
}
void tmgo_vec_destroy(tmgo_vec *v) {
    if (!v) return;
    free(v->buf.episode);
    free(v->buf.board); free(v->buf.timer); free(v->buf.draw_cursor); free(v->buf.shuffle_cursor);
    free(v->buf.reward); free(v->buf.terminated); free(v->buf.is_combination_match);
    free(v->buf.num_new_specials); free(v->buf.num_specials_activated); free(v->buf.shuffled);
    free(v->buf.mask); free(v->buf.num_moves_left); free(v->buf.status);
    for (int t = 0; t < v->nthreads; t++) tmgo_board_destroy(v->worker[t]);
    free(v->worker);
    free(v);
}
void tmgo_vec_get_buffers(tmgo_vec *v, tmgo_vec_buffers *out) { *out = v->buf; }
void tmgo_vec_set_injected_draws(tmgo_vec *v, const uint8_t *draws, int64_t per_env_len) { v->inj = draws; v->inj_len = per_env_len; }

static void load_env(tmgo_vec *v, tmgo_board *b, int e) {
    int P = v->P;
    const int8_t *src = v->buf.board + (size_t)e * 2 * (size_t)P;
    for (int i = 0; i < P; i++) { b->colour[i] = src[i]; b->type[i] = src[P + i]; }
    tmgo_board_set_stream(b, v->cfg.seed, (uint32_t)(v->cfg.env_id_offset + (uint64_t)e), v->buf.draw_cursor[e], v->buf.shuffle_cursor[e]);
    b->episode = v->buf.episode[e];
    b->in_reset = 0;
    if (v->cfg.refill_mode == 1) {
        b->inj = v->inj ? v->inj + (size_t)e * (size_t)v->inj_len : NULL;
        b->inj_len = v->inj ? v->inj_len : 0;
        b->use_inj = 1;
        static const uint8_t none = 0;
        if (!b->inj) b->inj = &none;
    } else b->use_inj = 0;
    b->status = 0;
}
static void store_env(tmgo_vec *v, tmgo_board *b, int e) {
    int P = v->P;
    int8_t *dst = v->buf.board + (size_t)e * 2 * (size_t)P;
    for (int i = 0; i < P; i++) { dst[i] = (int8_t)b->colour[i]; dst[P + i] = (int8_t)b->type[i]; }
    v->buf.draw_cursor[e] = b->draw_cursor;
    v->buf.shuffle_cursor[e] = b->shuffle_cursor;
    v->buf.episode[e] = b->episode;
    v->buf.status[e] |= b->status;
}
static int board_is_valid(const tmgo_board *b) {    /* full board of (1..K, 1..4) tiles or (0,-1) cookies */
    for (int i = 0; i < b->P; i++) {
        int k = b->colour[i], t = b->type[i];
        int ok = (t == -1 && k == 0) || (t >= 1 && t <= 4 && k >= 1 && k <= b->K);
        if (!ok) return 0;
    }
    return 1;
}
static void write_mask(tmgo_vec *v, tmgo_board *b, int e, int terminal) {
    uint8_t *m = v->buf.mask + (size_t)e * (size_t)v->A;
    if (terminal) memset(m, 0, (size_t)v->A);      /* tile_match_env.py:119-120 */
    else tmgo_effective_mask(b, m);
}
static void zero_outputs(tmgo_vec *v, int e) {
    v->buf.reward[e] = 0; v->buf.terminated[e] = 0; v->buf.is_combination_match[e] = 0;
    v->buf.num_new_specials[e] = 0; v->buf.num_specials_activated[e] = 0; v->buf.shuffled[e] = 0;
}
static void reset_one(tmgo_vec *v, tmgo_board *b, int e, const int8_t *init) {
    load_env(v, b, e);
    if (init) {
        int P = v->P;
        const int8_t *src = init + (size_t)e * 2 * (size_t)P;
        for (int i = 0; i < P; i++) { b->colour[i] = src[i]; b->type[i] = src[P + i]; }
        if (!board_is_valid(b)) b->status |= TMGO_ST_INVALID_BOARD;
    } else {
        tmgo_generate_board(b);                     /* tile_match_env.py:87 */
    }
    store_env(v, b, e);
    v->buf.timer[e] = 0;                            /* :88 */
    v->buf.num_moves_left[e] = v->cfg.num_moves;
    zero_outputs(v, e);
    write_mask(v, b, e, 0);                         /* :90 */
}
static void step_one(tmgo_vec *v, tmgo_board *b, int e, int action) {
    int nm = v->cfg.num_moves;
    int timer = v->buf.timer[e];
    if (timer < 0 || timer >= nm) {                 /* tile_match_env.py:94-95 raises */
        if (v->cfg.autoreset == TMGO_AUTORESET_NEXT_STEP && timer >= nm) { reset_one(v, b, e, NULL); return; }
        v->buf.status[e] |= TMGO_ST_NEEDS_RESET;
        zero_outputs(v, e);
        return;
    }
    if (action < 0 || action >= v->A) {             /* :97 IndexError */
        v->buf.status[e] |= TMGO_ST_BAD_ACTION;
        zero_outputs(v, e);
        return;
    }
    load_env(v, b, e);
    int r1, c1, r2, c2;
    int32_t out[5];
    tmgo_action_to_coords(b, action, &r1, &c1, &r2, &c2);
    tmgo_move(b, r1, c1, r2, c2, out);              /* :98 */
    timer += 1;                                     /* :100 */
    int done = timer == nm;                         /* :101 */
    v->buf.reward[e] = out[0];
    v->buf.is_combination_match[e] = (uint8_t)out[1];
    v->buf.num_new_specials[e] = out[2];
    v->buf.num_specials_activated[e] = out[3];
    v->buf.shuffled[e] = (uint8_t)out[4];
    v->buf.terminated[e] = (uint8_t)done;
    if (done && v->cfg.autoreset == TMGO_AUTORESET_SAME_STEP) {
        tmgo_generate_board(b);
        timer = 0;
        done = 0; /* mask/obs below describe the new episode; terminated stays 1 */
    }
    store_env(v, b, e);
    v->buf.timer[e] = timer;
    v->buf.num_moves_left[e] = nm - timer;
    write_mask(v, b, e, done);
}

typedef struct {
    tmgo_vec *v;
    int tid, lo, hi, mode; /* 0 reset, 1 step, 2 rollout */
    const uint8_t *reset_mask;
    const int8_t *init;
    const int32_t *actions;
    int steps;
    uint64_t action_seed, step0;
    int64_t reward_sum;
} job_t;

static void *worker_main(void *arg) {
    job_t *j = (job_t *)arg;
    tmgo_vec *v = j->v;
    tmgo_board *b = v->worker[j->tid];
    if (j->mode == 0) {
        for (int e = j->lo; e < j->hi; e++) if (!j->reset_mask || j->reset_mask[e]) reset_one(v, b, e, j->init);
    } else if (j->mode == 1) {
        for (int e = j->lo; e < j->hi; e++) step_one(v, b, e, j->actions[e]);
    } else {
        for (int e = j->lo; e < j->hi; e++) {
            uint32_t gid = (uint32_t)(v->cfg.env_id_offset + (uint64_t)e);
            for (int t = 0; t < j->steps; t++) {
                uint32_t w = tmgo_stream_word(j->action_seed, gid, 2, j->step0 + (uint64_t)t);
                int a = (int)(((uint64_t)w * (uint64_t)v->A) >> 32);
                step_one(v, b, e, a);
                j->reward_sum += v->buf.reward[e];
            }
        }
    }
    return NULL;
}

static int64_t run_jobs(tmgo_vec *v, job_t proto) {
    int T = v->nthreads, N = v->cfg.num_envs;
    if (T > N) T = N > 0 ? N : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)T);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)T);
    for (int t = 0; t < T; t++) {
        jobs[t] = proto;
        jobs[t].v = v; jobs[t].tid = t;
        jobs[t].lo = (int)((int64_t)N * t / T);
        jobs[t].hi = (int)((int64_t)N * (t + 1) / T);
        jobs[t].reward_sum = 0;
        if (T > 1) pthread_create(&th[t], NULL, worker_main, &jobs[t]);
        else worker_main(&jobs[t]);
    }
    int64_t sum = 0;
    for (int t = 0; t < T; t++) {
        if (T > 1) pthread_join(th[t], NULL);
        sum += jobs[t].reward_sum;
        int ml, md; int64_t ri;
        tmgo_board_diag(v->worker[t], &ml, &md, &ri);
        if (ml > v->max_lines) v->max_lines = ml;
        if (md > v->max_depth) v->max_depth = md;
        if (ri > v->max_reset_iters) v->max_reset_iters = ri;
    }
    free(th); free(jobs);
    return sum;
}

void tmgo_vec_reset(tmgo_vec *v, const uint8_t *reset_mask, const int8_t *init_boards) {
    job_t p; memset(&p, 0, sizeof(p));
    p.mode = 0; p.reset_mask = reset_mask; p.init = init_boards;
    run_jobs(v, p);
}
void tmgo_vec_step(tmgo_vec *v, const int32_t *actions) {
    job_t p; memset(&p, 0, sizeof(p));
    p.mode = 1; p.actions = actions;
    run_jobs(v, p);
}
/* steps every env `steps` times with uniform actions from stream 2 (cpu_baseline leg); returns sum of rewards */
int64_t tmgo_vec_rollout(tmgo_vec *v, int steps, uint64_t action_seed, uint64_t step0) {
    job_t p; memset(&p, 0, sizeof(p));
    p.mode = 2; p.steps = steps; p.action_seed = action_seed; p.step0 = step0;
    return run_jobs(v, p);
}
void tmgo_vec_onehot(tmgo_vec *v, uint8_t *out) {
    tmgo_board *b = v->worker[0];
    int planes = tmgo_onehot_planes(b);
    for (int e = 0; e < v->cfg.num_envs; e++) {
        load_env(v, b, e);
        tmgo_onehot(b, out + (size_t)e * (size_t)planes * (size_t)v->P);
    }
}
void tmgo_vec_diag(tmgo_vec *v, int *ml, int *md, int64_t *ri) { *ml = v->max_lines; *md = v->max_depth; *ri = v->max_reset_iters; }
