/*
 * syncerr_gpu.c -- row f2 of SURVEY.md section 8: read error correction on the all-syncmer graph.
 *
 * What the reference does (syncerr.c:679-757 find_error_syncmers, :342-612 the per-read pass, :144-288 the
 * graph search, :769-817 update_syncmer_db, :819-943 the driver; levdist.c:75-113, 156-225, 265-310 the
 * wavefront edit distance it searches with), restated from its behaviour:
 *
 *   suspects   a syncmer is suspect when its coverage is below err_mer_c, or -- below max_err_c -- when on one
 *              of its two sides it has arcs but none with coverage >= err_arc_c and >= max_arc_f * min(cov, cov');
 *              suspects are flagged in the database AND removed from the graph (vertex and all its arcs)
 *   blocks     on a read, trusted syncmers (not suspect, not already corrected) are anchors. A stretch of
 *              untrusted ones together with its flanking anchors is an error block; the anchor that closes a block
 *              must start at least 10 bases after the end of the syncmer before the block's first bad one. Blocks at
 *              the head of a read are searched backwards from their right anchor (reverse complement), blocks at
 *              the tail have no sink, and the bases after the last syncmer form one more tail block.
 *   search     depth-first over the graph from the left anchor: every arc appends the part of the next k-mer that
 *              does not overlap the current one to a candidate sequence, which is aligned to the read's bases of
 *              the block by a wavefront edit distance that is RESUMED from the parent's state (extension mode,
 *              band bw = max(6, ceil(0.02 * length))). A path counts when its score (edit distance plus unaligned
 *              target bases) is within the band and it ends in the sink (any vertex for tail blocks). The best one
 *              wins; a second one with the same score makes the block ambiguous (different bases: AMBISEQ,
 *              different syncmers: AMBISNQ). At most 10 000 search leaves.
 *   rewrite    only a SUCCESS replaces the block's syncmers by the path's inner vertices: k_mer = id << 1 | 1 and
 *              m_pos = 0xFFFFFFFE | strand (the "corrected" marker 0x7FFFFFFF as position). Everything else is kept.
 *   database   coverages and occurrence lists are rebuilt from the reads; a syncmer without any forward-strand
 *              occurrence left is flagged deleted (:811-812).
 *
 * The per-read pass is independent per read and runs on worker threads like the reference's kt_for; it only
 * reads the graph. Host code: it walks pointer-rich structures (per-read arrays, per-vertex strings).
 */
#include <stdlib.h>
#include <sys/time.h>
#include <sys/resource.h>
#include <string.h>
#include <math.h>
#include <assert.h>
#include <pthread.h>
#include <unistd.h>
#include "graph_gpu.h"
#include "syncgpu.h"

#define EC_FAILURE 0
#define EC_SUCCESS 1
#define EC_AMBISNQ 2
#define EC_AMBISEQ 3
#define EC_MAX_LEAVES 10000
#define EC_MIN_BLOCK 10          /* bases: shorter blocks are left alone (their flanking syncmers overlap) */
#define EC_MIN_BAND 6
#define NO_SINK UINT64_MAX
#define ID_MASK 0xFFFFFFFFFFFFFFFEULL

/* ---------------------------------------------------------------- wavefront edit distance, resumable
 * Landau-Vishkin / WFA in extension mode, as the graph search uses it (what the reference obtains from levdist.c's
 * wf_ed_core with is_ext = 1 and no traceback; its known answer, ED = 8 on the strings of levdist.c:445-446, and 3 600
 * random resumed comparisons against it are in tests/test_syncerr_cpu.py). Written from the recurrence:
 *
 *   far_e[d] = the last target index matched on diagonal d (query index - target index) with e edits
 *   far_{e+1}[d] = max(far_e[d-1], far_e[d] + 1, far_e[d+1] + 1), then slid along the matches of diagonal d
 *
 * The live diagonals of a wave are always contiguous, so the state is (lowest diagonal, count, far[]) plus the score;
 * it survives between calls, and a call with a longer query simply carries on -- which is how the search extends a
 * candidate by one vertex without re-aligning its prefix. A run ends when a diagonal reaches the last base of the
 * target or of the query (t_end / q_end = lengths aligned), or, with a band, when the score exceeds it.
 * The band rule is the reference's (levdist.c:99-113): once 2 bw + 1 diagonals are live, those outside
 * [max(-bw, -tl), max(bw, ql)] are dropped -- the upper bound WIDENS to the query length rather than clipping to bw;
 * before that only diagonals that left the matrix go. Observable through the scores, hence kept.
 */
typedef struct {
    const char *ts, *qs;
    int32_t tl, ql, bw;
    int32_t score, t_end, q_end;            /* t_end/q_end: aligned lengths when an end was reached, else 0 */
    int32_t d_lo;                           /* lowest live diagonal */
    int32_t *far, *nxt;                     /* far[j]: diagonal d_lo + j; nxt: scratch for the next wave */
    size_t n, m;                            /* live diagonals, capacity of far / nxt */
} wave_t;

#define WAVE_NONE (INT32_MIN / 2)

static void wave_reserve(wave_t *w, size_t need)
{
    if (w->m >= need) return;
    w->m = need + need / 2 + 16;
    w->far = (int32_t *) realloc(w->far, w->m * sizeof(int32_t));
    w->nxt = (int32_t *) realloc(w->nxt, w->m * sizeof(int32_t));
}

static void wave_begin(wave_t *w)           /* zero edits, nothing matched yet */
{
    wave_reserve(w, 8);
    w->score = 0; w->t_end = w->q_end = 0;
    w->d_lo = 0; w->n = 1; w->far[0] = -1;
}

static void wave_free(wave_t *w) { free(w->far); free(w->nxt); w->far = w->nxt = 0; w->m = 0; }

/* number of leading bytes two buffers share, at most `max` */
static inline int32_t shared_prefix(const char *x, const char *y, int32_t max)
{
    int32_t i = 0;
    for (; i + 8 <= max; i += 8) {
        uint64_t a, b;
        memcpy(&a, x + i, 8); memcpy(&b, y + i, 8);
        if (a != b) return i + (__builtin_ctzll(a ^ b) >> 3);
    }
    while (i < max && x[i] == y[i]) ++i;
    return i;
}

/* slides every live diagonal; 1 when one of them reached the end of the target or of the query */
static int wave_slide_all(wave_t *w, int32_t *t_end, int32_t *q_end)
{
    for (size_t j = 0; j < w->n; ++j) {
        const int32_t d = w->d_lo + (int32_t) j;
        int32_t k = w->far[j];
        if (k >= w->tl || k + d >= w->ql) continue;                     /* already outside */
        const int32_t room = (w->ql - d < w->tl ? w->ql - d : w->tl) - 1 - k;
        if (room > 0) k += shared_prefix(w->ts + k + 1, w->qs + k + d + 1, room);
        if (k == w->tl - 1 || k + d == w->ql - 1) { *t_end = k; *q_end = k + d; return 1; }
        w->far[j] = k;
    }
    return 0;
}

/* one more edit: the wave widens by a diagonal on either side, then the band rule trims it */
static void wave_widen(wave_t *w)
{
    const int32_t n = (int32_t) w->n, bw = w->bw;
    wave_reserve(w, (size_t) n + 4);
    const int32_t *f = w->far;
    int32_t *g = w->nxt, lo = 0, hi = n + 2, j;
    for (j = 0; j < n + 2; ++j) {                                       /* g[j]: diagonal d_lo - 1 + j */
        const int32_t ins = j >= 2 ? f[j - 2] : WAVE_NONE;              /* from the diagonal below */
        const int32_t sub = j >= 1 && j <= n ? f[j - 1] + 1 : WAVE_NONE;
        const int32_t del = j < n ? f[j] + 1 : WAVE_NONE;               /* from the diagonal above */
        int32_t best = ins > sub ? ins : sub;
        g[j] = best > del ? best : del;
    }
    const int32_t d0 = w->d_lo - 1;
    if (bw < 0 || n < 2 * bw + 1) {
        if (d0 < -w->tl) ++lo;
        if (d0 + n + 1 > w->ql) --hi;
    } else {
        const int32_t min_d = -bw > -w->tl ? -bw : -w->tl, max_d = bw > w->ql ? bw : w->ql;
        while (d0 + lo < min_d) ++lo;
        while (d0 + hi - 1 > max_d) --hi;
    }
    w->d_lo = d0 + lo;
    w->n = (size_t) (hi - lo);
    memcpy(w->far, g + lo, w->n * sizeof(int32_t));
}

static void wave_run(wave_t *w)
{
    int32_t t_end = w->t_end, q_end = w->q_end;
    for (;;) {
        t_end = q_end = -1;
        if (wave_slide_all(w, &t_end, &q_end)) break;
        wave_widen(w);
        ++w->score;
        if (w->bw >= 0 && w->score > w->bw) break;
    }
    w->t_end = t_end + 1; w->q_end = q_end + 1;
}

/* test hook: the extension-mode edit distance between two strings in one go, and resumed over a query that grows
 * piece by piece the way the graph search feeds it (the two must agree). bw < 0: no band. */
void oatk_wave_align(const char *ts, int32_t tl, const char *qs, int32_t ql, int32_t bw, int32_t grow, int32_t *out /* score, t_end, q_end */)
{
    wave_t w;
    memset(&w, 0, sizeof(w));
    w.ts = ts; w.tl = tl; w.bw = bw; w.qs = qs;
    wave_begin(&w);
    if (grow <= 0) grow = ql ? ql : 1;
    for (int32_t have = 0; ; ) {
        have = have + grow < ql ? have + grow : ql;
        w.ql = have;
        wave_run(&w);
        if (have >= ql || (w.t_end > 0 && w.t_end >= tl) || (bw >= 0 && w.score > bw)) break;
        /* an end of the QUERY was reached: the graph search appends more query and runs again from this state */
    }
    out[0] = w.score; out[1] = w.t_end; out[2] = w.q_end;
    wave_free(&w);
}

/* ---------------------------------------------------------------- small vectors */
typedef struct { size_t l, m; char *s; } str_t;
typedef struct { size_t n, m; uint64_t *a; } v64_t;
typedef struct { size_t n, m; uint32_t *a; } v32_t;

static inline void str_room(str_t *s, size_t extra)
{
    if (s->l + extra + 1 >= s->m) { s->m = (s->l + extra + 2) * 2; s->s = (char *) realloc(s->s, s->m); }
}
static inline void str_add(str_t *s, const char *p, size_t l) { str_room(s, l); memcpy(s->s + s->l, p, l); s->l += l; s->s[s->l] = 0; }
/* complement by table; the graph's strings hold ACGT (and N for syncmers without a copy), anything else stays as it is */
static char comp_tab[256];
static pthread_once_t comp_once = PTHREAD_ONCE_INIT;
static void comp_init(void)
{
    for (int c = 0; c < 256; ++c) comp_tab[c] = (char) c;
    comp_tab['A'] = 'T'; comp_tab['C'] = 'G'; comp_tab['G'] = 'C'; comp_tab['T'] = 'A';
}
static const char *comp_table(void)
{
    pthread_once(&comp_once, comp_init);               /* worker threads call this: built exactly once */
    return comp_tab;
}
static inline void str_add_rc(str_t *s, const char *p, size_t l)
{
    const char *t = comp_table();
    str_room(s, l);
    char *o = s->s + s->l;
    for (size_t i = 0; i < l; ++i) o[i] = t[(unsigned char) p[l - 1 - i]];
    s->l += l; s->s[s->l] = 0;
}
static inline void v64_push(v64_t *v, uint64_t x) { if (v->n == v->m) { v->m = v->m ? v->m * 2 : 16; v->a = (uint64_t *) realloc(v->a, v->m * 8); } v->a[v->n++] = x; }
static inline void v32_push(v32_t *v, uint32_t x) { if (v->n == v->m) { v->m = v->m ? v->m * 2 : 16; v->a = (uint32_t *) realloc(v->a, v->m * 4); } v->a[v->n++] = x; }
static inline void v64_copy(v64_t *d, const v64_t *s) { if (d->m < s->n) { d->m = s->n; d->a = (uint64_t *) realloc(d->a, d->m * 8); } d->n = s->n; if (s->n) memcpy(d->a, s->a, s->n * 8); }

/* ---------------------------------------------------------------- graph search */
typedef struct {
    int status, leaves, best, second;                   /* best / second best score so far */
    str_t cand, best_seq;
    v64_t path, best_path;
    int32_t *stash; size_t stash_n, stash_m;            /* wavefronts of the levels above, stacked */
} search_t;

static void search_from(const asmg_t *G, search_t *S, uint64_t sink, wave_t *w)
{
    if (S->leaves >= EC_MAX_LEAVES) return;
    const size_t l0 = S->cand.l, n0 = S->path.n, d0 = w->n;
    const uint64_t from = S->path.a[n0 - 1];
    const asmg_arc_t *arc = &G->arc[G->idx_p[from]];
    const uint64_t n_arc = G->idx_n[from];
    const int32_t t_end0 = w->t_end, q_end0 = w->q_end, s0 = w->score, lo0 = w->d_lo;
    const size_t stash0 = S->stash_n;
    if (S->stash_n + d0 > S->stash_m) { S->stash_m = (S->stash_n + d0) * 2 + 64; S->stash = (int32_t *) realloc(S->stash, S->stash_m * sizeof(int32_t)); }
    memcpy(S->stash + stash0, w->far, d0 * sizeof(int32_t));
    S->stash_n += d0;

    for (uint64_t i = 0; i < n_arc; ++i) {
        const asmg_arc_t *e = &arc[i];
        if (e->del) continue;
        const uint64_t v = e->w;
        const int ov = (int) e->ls, vl = (int) G->vtx[v >> 1].len;
        const char *vs = G->vtx[v >> 1].seq;
        v64_push(&S->path, v);
        if (v & 1) str_add_rc(&S->cand, vs, (size_t) (vl - ov));
        else str_add(&S->cand, vs + ov, (size_t) (vl - ov));
        w->qs = S->cand.s; w->ql = (int32_t) S->cand.l;
        wave_run(w);
        const int score = w->score + w->tl - w->t_end;   /* unaligned target bases count as edits */
        if (score <= w->bw && (sink == NO_SINK || sink == v)) {
            S->status = EC_SUCCESS;
            if (score <= S->best) {
                if (w->t_end > t_end0) S->second = S->best;   /* a real alternative, not an extension of its parent */
                S->best = score;
                /* tail blocks: a last vertex that is only partly covered by the read is not part of the answer */
                if (sink == NO_SINK && w->q_end < w->ql) --S->path.n;
                if (S->best == S->second) {
                    if ((size_t) w->q_end != S->best_seq.l || strncmp(S->cand.s, S->best_seq.s ? S->best_seq.s : "", (size_t) w->q_end)) S->status = EC_AMBISEQ;
                    if (S->status == EC_SUCCESS) {
                        int same = S->path.n == S->best_path.n;
                        for (size_t j = 0; same && j < S->path.n; ++j) same = S->path.a[j] == S->best_path.a[j];
                        if (!same) S->status = EC_AMBISNQ;
                    }
                }
                S->best_seq.l = 0;
                str_add(&S->best_seq, S->cand.s, (size_t) w->q_end);
                v64_copy(&S->best_path, &S->path);
            } else if (score < S->second) S->second = score;
        }
        if (w->score <= w->bw && w->ql - vl <= w->tl + w->bw && ((sink != NO_SINK && sink != v) || w->t_end < w->tl))
            search_from(G, S, sink, w);
        else ++S->leaves;
        /* back to the state in front of this arc */
        S->path.n = n0; S->cand.l = l0;
        w->t_end = t_end0; w->q_end = q_end0; w->score = s0; w->n = d0; w->d_lo = lo0;
        memcpy(w->far, S->stash + stash0, d0 * sizeof(int32_t));    /* (the stash may have moved: always through S) */
    }
    S->stash_n = stash0;
}

/* ---------------------------------------------------------------- one read */
typedef struct {
    wave_t w; search_t S; str_t seq; v64_t kk; v32_t pp;
    long stats[11];                                     /* tail blocks, their 4 outcomes, middle blocks, their 4 outcomes, too short */
} ec_worker_t;

static void correct_read(sr_db_t *db, const scg_t *g, double max_edist, uint64_t rid, ec_worker_t *W)
{
    const asmg_t *G = g->utg_asmg;
    const syncmer_t *scm = g->scm_db->a;
    sr_t *sr = &db->a[rid];
    const int ksz = db->k, n = (int) sr->n;
    const uint64_t *K = sr->k_mer;
    const uint32_t *P = sr->m_pos;
    int beg = -1, end, updated = 1;
    W->kk.n = 0; W->pp.n = 0;

    for (;;) {
        uint32_t from = beg < 1 ? 0 : (P[beg - 1] >> 1) + (uint32_t) ksz;
        from += EC_MIN_BLOCK;
        for (end = beg + 1; end < n; ++end)
            if (!scm[K[end] >> 1].del && !(K[end] & 1) && (P[end] >> 1) >= from) break;    /* the next anchor */

        if (beg >= 0 || end < n) {
            uint64_t src, sink;
            int l, rev, res;
            uint32_t at;
            if (beg < 0) {                              /* head of the read: search backwards from the first anchor */
                beg = end;
                src = (K[beg] & ID_MASK) | (uint64_t) !(P[beg] & 1);
                at = 0; sink = NO_SINK; l = (int) (P[beg] >> 1); rev = 1;
            } else {
                --beg;                                  /* the anchor in front of the block */
                src = (K[beg] & ID_MASK) | (P[beg] & 1);
                at = (P[beg] >> 1) + (uint32_t) ksz;
                if (end >= n) { sink = NO_SINK; l = (int) sr->hoco_l - (int) at; }
                else { sink = (K[end] & ID_MASK) | (P[end] & 1); l = (int) (P[end] >> 1) - (int) at; }
                rev = 0;
            }
            assert(l >= 0);
            if (W->seq.m < (size_t) l + 1) { W->seq.m = (size_t) l + 1; W->seq.s = (char *) realloc(W->seq.s, W->seq.m); }
            get_kmer_dna_seq(sr->hoco_s, at, l, (uint32_t) rev, W->seq.s);
            W->seq.l = (size_t) l;
            if (l >= EC_MIN_BLOCK) {
                wave_t *w = &W->w;
                w->ts = W->seq.s; w->tl = l; w->qs = 0; w->ql = 0;
                wave_begin(w);
                w->bw = (int32_t) ceil(l * max_edist);
                if (w->bw < EC_MIN_BAND) w->bw = EC_MIN_BAND;
                search_t *S = &W->S;
                S->status = EC_FAILURE; S->leaves = 0; S->best = S->second = INT32_MAX;
                S->cand.l = 0; S->best_seq.l = 0; S->path.n = 0; S->best_path.n = 0;
                v64_push(&S->path, src);
                search_from(G, S, sink, w);
                res = S->status;
                if (res) assert(src == S->best_path.a[0] && (sink == NO_SINK || sink == S->best_path.a[S->best_path.n - 1]));
                if (sink == NO_SINK) { ++W->stats[0]; ++W->stats[1 + res]; }
                else { ++W->stats[5]; ++W->stats[6 + res]; }
            } else { res = EC_FAILURE; ++W->stats[10]; }

            if (res == EC_SUCCESS) {                    /* the path's inner vertices replace the block */
                const v64_t *bp = &W->S.best_path;
                const int np = (int) bp->n;
                int j;
                if (rev) {
                    for (j = np - 1; j > 0; --j) { v64_push(&W->kk, (bp->a[j] & ID_MASK) | 1); v32_push(&W->pp, 0xFFFFFFFFu ^ (uint32_t) (bp->a[j] & 1)); }
                } else {
                    for (j = 1; j < np - 1; ++j) { v64_push(&W->kk, (bp->a[j] & ID_MASK) | 1); v32_push(&W->pp, 0xFFFFFFFEu | (uint32_t) (bp->a[j] & 1)); }
                    if (sink == NO_SINK && np > 1) { v64_push(&W->kk, (bp->a[j] & ID_MASK) | 1); v32_push(&W->pp, 0xFFFFFFFEu | (uint32_t) (bp->a[j] & 1)); }
                }
            } else if (rev) {                           /* keep what the read had */
                for (int j = 0; j < beg; ++j) { v64_push(&W->kk, K[j]); v32_push(&W->pp, P[j]); }
            } else if (beg + 1 < n) {
                for (int j = beg + 1; j < end; ++j) { v64_push(&W->kk, K[j]); v32_push(&W->pp, P[j]); }
            }
        } else updated = 0;                             /* not a single anchor on this read */

        /* the run of anchors that follows; note that the reference tests the flag of K[end] here, not of K[beg] (:579) */
        for (beg = end + 1; beg < n; ++beg)
            if (scm[K[beg] >> 1].del || (K[end] & 1)) break;
        if (beg > n) break;
        for (int j = end; j < beg; ++j) { v64_push(&W->kk, K[j]); v32_push(&W->pp, P[j]); }
    }

    if (updated) {
        const size_t nc = W->kk.n;
        sr->k_mer = (uint64_t *) realloc(sr->k_mer, nc * sizeof(uint64_t));
        sr->m_pos = (uint32_t *) realloc(sr->m_pos, nc * sizeof(uint32_t));
        sr->s_mer = (uint64_t *) realloc(sr->s_mer, nc * sizeof(uint64_t));
        if (nc) { memcpy(sr->k_mer, W->kk.a, nc * sizeof(uint64_t)); memcpy(sr->m_pos, W->pp.a, nc * sizeof(uint32_t)); }
        for (size_t j = 0; j < nc; ++j) sr->s_mer[j] = scm[sr->k_mer[j] >> 1].s;
        sr->n = (uint32_t) nc;
    }
}

/* ---------------------------------------------------------------- suspects */
/* Which syncmers look like sequencing errors (syncerr.c:679-757): fewer than err_mer_c occurrences, or -- below
 * max_err_c -- a side that has arcs but none of them reliable (>= err_arc_c reads and >= max_arc_f of the thinner
 * end). The test of one syncmer reads only coverages and the arcs' flags as they were on entry, so the candidates are
 * found in parallel into a byte map and the flags set afterwards; a deleted vertex takes every arc that touches it
 * with it (the graph is symmetric after asmg_finalize, so "its out-arcs and their complements" is "both ends"),
 * which one pass over the arc array does without chasing pointers. */
typedef struct { const scg_t *g; uint8_t *err; uint32_t err_mer_c, max_err_c, err_arc_c; double max_arc_f; } fes_t;

static void fes_mark(uint64_t lo, uint64_t hi, void *arg)
{
    const fes_t *F = (const fes_t *) arg;
    const asmg_t *G = F->g->utg_asmg;
    const syncmer_t *scm = F->g->scm_db->a;
    for (uint64_t i = lo; i < hi; ++i) {
        if (scm[i].del || scm[i].cov >= F->max_err_c) continue;
        if (scm[i].cov < F->err_mer_c) { F->err[i] = 1; continue; }
        int side_ok[2] = {-1, -1};                      /* -1: no live arc on that side */
        for (int side = 0; side < 2; ++side) {
            const uint64_t v = (uint64_t) i << 1 | (uint64_t) side;
            const asmg_arc_t *a = &G->arc[G->idx_p[v]];
            const uint64_t na = G->idx_n[v];
            uint64_t live = 0;
            for (uint64_t j = 0; j < na; ++j) live += !a[j].del;
            if (!live) continue;
            side_ok[side] = 0;
            for (uint64_t j = 0; j < na; ++j) {
                if (a[j].del) continue;
                const uint32_t cw = scm[a[j].w >> 1].cov, cv = scm[i].cov;
                if (a[j].cov >= F->err_arc_c && a[j].cov >= (cv < cw ? cv : cw) * F->max_arc_f) { side_ok[side] = 1; break; }
            }
        }
        if (!side_ok[0] || !side_ok[1]) F->err[i] = 1;
    }
}

static void fes_drop_arcs(uint64_t lo, uint64_t hi, void *arg)
{
    const fes_t *F = (const fes_t *) arg;
    asmg_arc_t *a = F->g->utg_asmg->arc;
    for (uint64_t i = lo; i < hi; ++i) if (F->err[a[i].v >> 1] || F->err[a[i].w >> 1]) a[i].del = 1;
}

int64_t find_error_syncmers(scg_t *g, uint32_t err_mer_c, uint32_t max_err_c, uint32_t err_arc_c, double max_arc_f, int del_err)
{
    asmg_t *G = g->utg_asmg;
    syncmer_t *scm = g->scm_db->a;
    const size_t n_scm = g->scm_db->n;
    fes_t F = {g, (uint8_t *) calloc(n_scm ? n_scm : 1, 1), err_mer_c, max_err_c, err_arc_c, max_arc_f};
    oatk_parallel_for(n_scm, fes_mark, &F);
    int64_t n_err = 0;
    uint32_t max_c = 0;
    for (size_t i = 0; i < n_scm; ++i) {
        if (F.err[i]) scm[i].del = 1;
        else F.err[i] = scm[i].del;                     /* deleted before this call: dropped from the graph as well */
        if (scm[i].del) { if (scm[i].cov > max_c) max_c = scm[i].cov; ++n_err; }
    }
    if (del_err) {
        for (size_t i = 0; i < n_scm; ++i) if (F.err[i]) G->vtx[i].del = 1;
        oatk_parallel_for(G->n_arc, fes_drop_arcs, &F);
    }
    free(F.err);
    fprintf(stderr, "[M::%s] error syncmer candidates: num = %ld, max_c = %u\n", __func__, (long) n_err, max_c);
    return n_err;
}

/* ---------------------------------------------------------------- the per-read pass on the device
 * sg_ec_correct (csrc/sg_ec.cu) runs the same blocks / search / rewrite, one warp per read, on the lists and hoco_s that are
 * resident in the batch behind sr_db. It gets what the error filter left of the graph: the live arcs in the graph's own
 * order (the search visits them in that order) with their overlaps, and for every arc target the occurrence its k hoco
 * bases are read from -- the first one no correction touched, which is what scg_consensus writes as the vertex text in
 * hoco mode (syncasm.c:911-931). Returns 0 when the lists were rewritten; anything else leaves the reads untouched and the
 * caller runs the host form. Conditions: every vertex is one syncmer with a text of k hoco bases (true for the graph
 * read_error_correction is given) and the arc array is sorted by its source vertex. */
static int g_last_on_device;
static uint64_t g_last_overflow;
/* where the per-read pass of the last read_error_correction ran (1: device) and how many reads needed the worst-case arena */
int oatk_ec_last_run(uint64_t *overflow_reads) { if (overflow_reads) *overflow_reads = g_last_overflow; return g_last_on_device; }

/* the live arcs of the filtered graph, in the graph's order, gathered by worker threads: every range counts its live arcs
 * first, the ranges are then placed one after another, and the same ranges fill their part */
typedef struct { uint64_t lo, n, at; } ec_part_t;
typedef struct {
    sr_db_t *db; const asmg_t *G; const syncmer_t *scm;
    uint64_t *av, *aw, *at; uint32_t *al;
    ec_part_t part[64]; int n_part; int fill; int bad;
    pthread_mutex_t lock;
} ec_pack_t;

static void ec_pack_range(uint64_t lo, uint64_t hi, void *arg)
{
    ec_pack_t *P = (ec_pack_t *) arg;
    const asmg_t *G = P->G;
    const int ksz = P->db->k;
    uint64_t i, j = 0;
    if (!P->fill) {
        uint64_t c = 0, prev = lo ? G->arc[lo - 1].v : 0;
        int bad = 0;
        for (i = lo; i < hi; ++i) {
            const asmg_arc_t *a = &G->arc[i];
            if (a->v < prev) bad = 1;                               /* the search relies on arcs sorted by their source */
            prev = a->v;
            if (a->del) continue;
            const asmg_vtx_t *vx = &G->vtx[a->w >> 1];
            if (vx->len != (uint64_t) ksz || vx->n != 1 || a->ls >= (uint64_t) ksz) bad = 1;
            ++c;
        }
        pthread_mutex_lock(&P->lock);
        if (P->n_part < 64) { P->part[P->n_part].lo = lo; P->part[P->n_part].n = c; ++P->n_part; } else bad = 1;
        P->bad |= bad;
        pthread_mutex_unlock(&P->lock);
        return;
    }
    for (int q = 0; q < P->n_part; ++q) if (P->part[q].lo == lo) j = P->part[q].at;
    for (i = lo; i < hi; ++i) {
        const asmg_arc_t *a = &G->arc[i];
        if (a->del) continue;
        P->av[j] = a->v; P->aw[j] = a->w; P->al[j] = (uint32_t) a->ls;
        /* the occurrence that supplies the bases of w's syncmer, in the syncmer's own orientation */
        const syncmer_t *m = &P->scm[a->w >> 1];
        uint64_t ref = UINT64_MAX;
        for (uint32_t c = 0; c < m->cov; ++c) {
            const uint64_t occ = m->m_pos[c];
            const sr_t *r = &P->db->a[occ >> 32];
            const uint64_t idx = occ >> 1 & MAX_RD_SCM;
            if (r->k_mer[idx] & 1) continue;
            ref = (occ >> 32) << 32 | (uint64_t) (r->m_pos[idx] >> 1) << 1 | (r->m_pos[idx] & 1);
            break;
        }
        P->at[j] = ref;
        ++j;
    }
}

typedef struct { sr_db_t *db; const syncmer_t *scm; const sg_ec_result_t *R; } ec_apply_t;
static void ec_apply_range(uint64_t lo, uint64_t hi, void *arg)
{
    const ec_apply_t *A = (const ec_apply_t *) arg;
    const sg_ec_result_t *R = A->R;
    for (uint64_t r = lo; r < hi; ++r) {
        if (R->out_n[r] == 0xffffffffu) continue;
        sr_t *sr = &A->db->a[r];
        const size_t nc = R->out_n[r];
        sr->k_mer = (uint64_t *) realloc(sr->k_mer, nc * sizeof(uint64_t));
        sr->m_pos = (uint32_t *) realloc(sr->m_pos, nc * sizeof(uint32_t));
        sr->s_mer = (uint64_t *) realloc(sr->s_mer, nc * sizeof(uint64_t));
        if (nc) { memcpy(sr->k_mer, R->out_k + R->out_off[r], nc * sizeof(uint64_t)); memcpy(sr->m_pos, R->out_p + R->out_off[r], nc * sizeof(uint32_t)); }
        for (size_t q = 0; q < nc; ++q) sr->s_mer[q] = A->scm[sr->k_mer[q] >> 1].s;
        sr->n = (uint32_t) nc;
    }
}

/* upload, search, and the rewritten lists back into the reads */
static int ec_device_run(sr_db_t *db, const syncmer_db_t *S, const uint8_t *del, uint64_t n_live, const uint64_t *av, const uint64_t *aw,
        const uint32_t *al, const uint64_t *at, double max_edist, long stats[11])
{
    sg_ec_graph_t E;
    memset(&E, 0, sizeof(E));
    E.n_syncmers = S->n; E.del = del; E.n_arcs = n_live; E.arc_v = av; E.arc_w = aw; E.arc_ls = al; E.arc_txt = at;
    sg_ec_result_t R;
    int rc = oatk_gpu_ec_correct(db, &E, max_edist, &R);
    oatk_tick("ec/device: upload, search kernel, download");
    if (rc == 0) {
        ec_apply_t A = {db, S->a, &R};
        oatk_parallel_for(db->n, ec_apply_range, &A);
        for (int q = 0; q < 11; ++q) stats[q] = (long) R.stats[q];
        g_last_overflow = R.n_overflow_reads;
        sg_ec_result_free(&R);
        oatk_tick("ec/device: lists back into the reads");
    }
    return rc;
}

static int cmp_part(const void *a, const void *b) { const ec_part_t *x = (const ec_part_t *) a, *y = (const ec_part_t *) b; return x->lo < y->lo ? -1 : x->lo > y->lo; }

static int correct_reads_on_device(sr_db_t *db, scg_t *g, double max_edist, long stats[11])
{
    const asmg_t *G = g->utg_asmg;
    const syncmer_db_t *S = g->scm_db;
    const syncmer_t *scm = S->a;
    if (!oatk_gpu_ec_available(db) || G->n_vtx != S->n) return -1;
    ec_pack_t P;
    memset(&P, 0, sizeof(P));
    P.db = db; P.G = G; P.scm = scm;
    pthread_mutex_init(&P.lock, 0);
    oatk_parallel_for(G->n_arc, ec_pack_range, &P);
    qsort(P.part, (size_t) P.n_part, sizeof(ec_part_t), cmp_part);
    uint64_t n_live = 0;
    for (int q = 0; q < P.n_part; ++q) { P.part[q].at = n_live; n_live += P.part[q].n; }
    int rc = -1;
    if (!P.bad) {
        uint8_t *del = (uint8_t *) malloc(S->n ? S->n : 1);
        P.av = (uint64_t *) malloc(8 * (n_live + 1)); P.aw = (uint64_t *) malloc(8 * (n_live + 1)); P.at = (uint64_t *) malloc(8 * (n_live + 1));
        P.al = (uint32_t *) malloc(4 * (n_live + 1));
        P.fill = 1;
        oatk_parallel_for(G->n_arc, ec_pack_range, &P);             /* same n, same ranges */
        for (uint64_t i = 0; i < S->n; ++i) del[i] = (uint8_t) scm[i].del;
        oatk_tick("ec/device: arrays of the filtered graph");
        rc = ec_device_run(db, S, del, n_live, P.av, P.aw, P.al, P.at, max_edist, stats);
        free(del); free(P.av); free(P.aw); free(P.at); free(P.al);
    }
    pthread_mutex_destroy(&P.lock);
    return rc;
}

/* ---------------------------------------------------------------- database after the rewrite */
/* The occurrence lists are rebuilt by worker threads that each own a range of syncmer ids and walk all reads: a list
 * is then filled by one thread, in read order, as the reference's single loop fills it (syncerr.c:769-817). */
typedef struct { sr_db_t *db; syncmer_db_t *S; uint32_t *cnt, *fwd; int pass; } rebuild_t;
static void rebuild_range(uint64_t lo, uint64_t hi, void *arg)
{
    rebuild_t *R = (rebuild_t *) arg;
    sr_db_t *db = R->db;
    syncmer_t *scm = R->S->a;
    if (R->pass == 0) {
        for (size_t r = 0; r < db->n; ++r) {
            const sr_t *sr = &db->a[r];
            for (uint32_t j = 0; j < sr->n; ++j) { const uint64_t id = sr->k_mer[j] >> 1; if (id >= lo && id < hi) ++R->cnt[id]; }
        }
        /* the lists are one malloc block per syncmer (syncmer.c:1359); a block that is large enough is kept -- nearly all
         * syncmers are sequencing errors whose list only shrinks */
        for (uint64_t i = lo; i < hi; ++i) {
            if (R->cnt[i] > scm[i].cov || !scm[i].m_pos) {
                free(scm[i].m_pos);
                scm[i].m_pos = (uint64_t *) malloc((R->cnt[i] ? R->cnt[i] : 1) * sizeof(uint64_t));
            }
            scm[i].cov = 0;
        }
        return;
    }
    for (size_t r = 0; r < db->n; ++r) {
        const sr_t *sr = &db->a[r];
        for (uint32_t j = 0; j < sr->n; ++j) {
            const uint64_t id = sr->k_mer[j] >> 1;
            if (id < lo || id >= hi) continue;
            scm[id].m_pos[scm[id].cov++] = sr->sid << 32 | (uint64_t) j << 1 | (sr->m_pos[j] & 1);
            if (!(sr->m_pos[j] & 1)) ++R->fwd[id];
        }
    }
    for (uint64_t i = lo; i < hi; ++i) scm[i].del = !R->fwd[i];       /* syncerr.c:805-812 */
}

static void rebuild_syncmer_db(sr_db_t *db, syncmer_db_t *S)
{
    rebuild_t R = {db, S, (uint32_t *) calloc(S->n ? S->n : 1, sizeof(uint32_t)), (uint32_t *) calloc(S->n ? S->n : 1, sizeof(uint32_t)), 0};
    free(S->c); S->c = 0;
    free(S->h); S->h = 0;
    oatk_parallel_for(S->n, rebuild_range, &R);
    R.pass = 1;
    oatk_parallel_for(S->n, rebuild_range, &R);
    free(R.fwd); free(R.cnt);
}

static void rebuild_syncmer_db(sr_db_t *db, syncmer_db_t *S);
/* the closing lines of read_error_correction (syncerr.c:899-923); they carry that function's name whichever form ran */
static void ec_print_summary(const long stats[11], int verbose, double cpu_at_entry, double wall_at_entry)
{
    const char *fn = "read_error_correction";
    fprintf(stderr, "[M::%s] Error Correction Summary Results\n", fn);
    fprintf(stderr, "[M::%s] total number of error blocks : %ld\n", fn, stats[0] + stats[5] + stats[10]);
    fprintf(stderr, "[M::%s]                - uncorrected : %ld\n", fn, stats[1] + stats[6]);
    fprintf(stderr, "[M::%s]                  - corrected : %ld\n", fn, stats[2] + stats[7]);
    fprintf(stderr, "[M::%s]             - ambiguous seqs : %ld\n", fn, stats[3] + stats[8]);
    fprintf(stderr, "[M::%s]             - ambiguous path : %ld\n", fn, stats[4] + stats[9]);
    if (verbose) {
        fprintf(stderr, "[M::%s] error blocks in the tail end : %ld\n", fn, stats[0]);
        fprintf(stderr, "[M::%s]                - uncorrected : %ld\n", fn, stats[1]);
        fprintf(stderr, "[M::%s]                  - corrected : %ld\n", fn, stats[2]);
        fprintf(stderr, "[M::%s]             - ambiguous seqs : %ld\n", fn, stats[3]);
        fprintf(stderr, "[M::%s]             - ambiguous path : %ld\n", fn, stats[4]);
        fprintf(stderr, "[M::%s]   error blocks in the middle : %ld\n", fn, stats[5]);
        fprintf(stderr, "[M::%s]                - uncorrected : %ld\n", fn, stats[6]);
        fprintf(stderr, "[M::%s]                  - corrected : %ld\n", fn, stats[7]);
        fprintf(stderr, "[M::%s]             - ambiguous seqs : %ld\n", fn, stats[8]);
        fprintf(stderr, "[M::%s]             - ambiguous path : %ld\n", fn, stats[9]);
        fprintf(stderr, "[M::%s]      error blocks overlapped : %ld\n", fn, stats[10]);
        {   /* syncerr.c:921-922: the step's own clocks */
            struct rusage ru;
            struct timeval tv;
            getrusage(RUSAGE_SELF, &ru);
            gettimeofday(&tv, 0);
            fprintf(stderr, "[M::%s]   error correction  CPU time : %.3f sec\n", fn,
                    ru.ru_utime.tv_sec + ru.ru_stime.tv_sec + 1e-6 * (ru.ru_utime.tv_usec + ru.ru_stime.tv_usec) - cpu_at_entry);
            fprintf(stderr, "[M::%s]   error correction real time : %.3f sec\n", fn, tv.tv_sec + 1e-6 * tv.tv_usec - wall_at_entry);
        }
    }
}

static void ec_clocks(double *cpu, double *wall)
{
    struct rusage ru;
    struct timeval tv;
    getrusage(RUSAGE_SELF, &ru);
    gettimeofday(&tv, 0);
    *cpu = ru.ru_utime.tv_sec + ru.ru_stime.tv_sec + 1e-6 * (ru.ru_utime.tv_usec + ru.ru_stime.tv_usec);
    *wall = tv.tv_sec + 1e-6 * tv.tv_usec;
}

/* ---------------------------------------------------------------- the whole step without a host graph
 * What syncasm() does for its error-correction step when the reads live on the device: the all-syncmer graph (one
 * vertex per distinct k-mer, ~10^7 arcs at 200 k reads) exists only to be filtered down to the ~2 % of it the search can
 * walk, so it is never built here. The device tallies the arcs (sg_arcs with no thresholds: exactly the arcs
 * make_syncmer_graph(sr_db, scm_db, 0, 0.) would receive, in the order asmg_finalize leaves them), flags the suspects
 * (sg_ec_filter) and sends back the flags and the surviving arcs; the host computes those arcs' overlaps the way
 * scg_consensus does in hoco space (oatk_syncmer_arc_overlaps), the search runs on the device (sg_ec_correct), and the
 * database is rebuilt as in read_error_correction. Same messages, same results (tests/test_gpu_ec.py compares the two
 * forms; the CLI tests compare stderr and both GFA files with the reference binary). Returns non-zero, having changed
 * nothing, when the conditions do not hold: no device batch behind sr_db, or syncmers already flagged deleted (the graph
 * the reference builds would then be renumbered by asmg_finalize). */
typedef struct { sr_db_t *db; const syncmer_t *scm; const uint64_t *arcs4; uint64_t *at; } ec_txt_t;
static void ec_txt_range(uint64_t lo, uint64_t hi, void *arg)
{
    ec_txt_t *T = (ec_txt_t *) arg;
    for (uint64_t i = lo; i < hi; ++i) {
        const syncmer_t *m = &T->scm[T->arcs4[4 * i + 1] >> 1];
        uint64_t ref = UINT64_MAX;
        for (uint32_t c = 0; c < m->cov; ++c) {
            const uint64_t occ = m->m_pos[c];
            const sr_t *r = &T->db->a[occ >> 32];
            const uint64_t idx = occ >> 1 & MAX_RD_SCM;
            if (r->k_mer[idx] & 1) continue;
            ref = (occ >> 32) << 32 | (uint64_t) (r->m_pos[idx] >> 1) << 1 | (r->m_pos[idx] & 1);
            break;
        }
        T->at[i] = ref;
    }
}

int read_error_correction_device(sr_db_t *sr_db, syncmer_db_t *scm_db, double max_edist, uint32_t err_mer_c, uint32_t max_err_c,
        uint32_t err_arc_c, double max_arc_f, int n_threads, int verbose)
{
    (void) n_threads;
    const char *force_host = getenv("OATK_EC_HOST"), *force_graph = getenv("OATK_EC_GRAPH");
    if ((force_host && atoi(force_host) > 0) || (force_graph && atoi(force_graph) > 0)) return -1;
    if (!oatk_gpu_ec_available(sr_db) || scm_db->n == 0) return -1;
    syncmer_t *scm = scm_db->a;
    for (size_t i = 0; i < scm_db->n; ++i) if (scm[i].del) return -1;
    double cpu_at_entry, wall_at_entry;
    ec_clocks(&cpu_at_entry, &wall_at_entry);
    oatk_tick(0);
    sg_ec_filter_out_t F;
    if (oatk_gpu_ec_filter(sr_db, err_mer_c, max_err_c, err_arc_c, max_arc_f, &F) != 0) return -1;
    oatk_tick("ec/graph-free: arc tally + error filter (device), flags and surviving arcs down");
    /* nothing has been changed so far; from here on the step is committed */
    int64_t n_err = 0;
    uint32_t max_c = 0;
    for (size_t i = 0; i < scm_db->n; ++i)
        if (F.err[i]) { scm[i].del = 1; if (scm[i].cov > max_c) max_c = scm[i].cov; ++n_err; }
    fprintf(stderr, "[M::%s] error syncmer candidates: num = %ld, max_c = %u\n", "find_error_syncmers", (long) n_err, max_c);
    const uint64_t n = F.n_live;
    uint64_t *av = (uint64_t *) malloc(8 * (n + 1)), *aw = (uint64_t *) malloc(8 * (n + 1)), *at = (uint64_t *) malloc(8 * (n + 1));
    uint32_t *al = (uint32_t *) malloc(4 * (n + 1));
    for (uint64_t i = 0; i < n; ++i) { av[i] = F.arcs4[4 * i]; aw[i] = F.arcs4[4 * i + 1]; }
    oatk_syncmer_arc_overlaps(sr_db, scm_db, n, F.arcs4, al);
    ec_txt_t T = {sr_db, scm, F.arcs4, at};
    oatk_parallel_for(n, ec_txt_range, &T);
    oatk_tick("ec/graph-free: overlaps and text sources of the surviving arcs");
    long stats[11] = {0};
    const int rc = ec_device_run(sr_db, scm_db, F.err, n, av, aw, al, at, max_edist, stats);
    free(av); free(aw); free(at); free(al);
    sg_ec_filter_free(&F);
    if (rc != 0) {
        fprintf(stderr, "[E::%s] the device search failed after the error filter had been applied\n", __func__);
        exit(EXIT_FAILURE);
    }
    g_last_on_device = 2;
    rebuild_syncmer_db(sr_db, scm_db);
    oatk_tick("ec: rebuild database");
    oatk_gpu_update_lists(sr_db, scm_db);
    oatk_tick("ec: refresh device lists");
    ec_print_summary(stats, verbose, cpu_at_entry, wall_at_entry);
    return 0;
}

/* ---------------------------------------------------------------- driver */
typedef struct { sr_db_t *db; scg_t *g; double max_edist; uint64_t *next; ec_worker_t W; pthread_t th; } ec_job_t;

static void *ec_job(void *arg)
{
    ec_job_t *J = (ec_job_t *) arg;
    for (;;) {
        const uint64_t r = __atomic_fetch_add(J->next, 1, __ATOMIC_RELAXED);
        if (r >= J->db->n) break;
        correct_read(J->db, J->g, J->max_edist, r, &J->W);
    }
    return 0;
}

void read_error_correction(sr_db_t *sr_db, scg_t *g, double max_edist, uint32_t err_mer_c, uint32_t max_err_c,
        uint32_t err_arc_c, double max_arc_f, int n_threads, FILE *fo, int verbose)
{
    (void) fo;                                          /* the corrected-read FASTA is a debugging aid of the reference */
    double cpu_at_entry, wall_at_entry;
    {
        struct rusage ru;
        struct timeval tv;
        getrusage(RUSAGE_SELF, &ru);
        gettimeofday(&tv, 0);
        cpu_at_entry = ru.ru_utime.tv_sec + ru.ru_stime.tv_sec + 1e-6 * (ru.ru_utime.tv_usec + ru.ru_stime.tv_usec);
        wall_at_entry = tv.tv_sec + 1e-6 * tv.tv_usec;
    }
    if (n_threads <= 0) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    /* The walk needs the hoco text of the vertices and the overlaps of the arcs it can reach. The reference computes
     * them for ALL syncmers before the error filter (run_syncasm.c:118, its largest single cost); a caller that has not
     * done so gets them here for what the filter leaves -- texts and overlaps are per vertex and per arc, so the
     * survivors' are the same (tests/test_syncerr_cpu.py::test_deferred_consensus). */
    int have_text = 0;
    for (uint64_t i = 0; i < g->utg_asmg->n_vtx && !have_text; ++i) if (!g->utg_asmg->vtx[i].del) { have_text = g->utg_asmg->vtx[i].seq != 0; break; }
    oatk_tick(0);
    find_error_syncmers(g, err_mer_c, max_err_c, err_arc_c, max_arc_f, 1);
    oatk_tick("ec: find error syncmers");
    if (!have_text) { scg_consensus(sr_db, g, 1, 1, 0); oatk_tick("ec: hoco consensus of the surviving graph"); }

    /* the per-read pass: on the device when the reads live there (the reference's kt_for over reads, syncerr.c:896),
     * on worker threads for read databases that were built on the host (OATK_EC_HOST=1 forces that form: parity tests) */
    long stats[11] = {0};
    ec_job_t *J = (ec_job_t *) calloc((size_t) n_threads, sizeof(ec_job_t));
    const char *force_host = getenv("OATK_EC_HOST");
    int on_device = 0;
    if (!(force_host && atoi(force_host) > 0)) on_device = correct_reads_on_device(sr_db, g, max_edist, stats) == 0;
    g_last_on_device = on_device;
    if (!on_device && oatk_gpu_bases_on_device(sr_db)) {
        fprintf(stderr, "[E::%s] the packed bases of the reads are on the device only and the device search did not run\n", __func__);
        exit(EXIT_FAILURE);
    }
    if (!on_device) {
        uint64_t next = 0;
        for (int t = 0; t < n_threads; ++t) { J[t].db = sr_db; J[t].g = g; J[t].max_edist = max_edist; J[t].next = &next; }
        if (n_threads == 1) ec_job(&J[0]);
        else {
            for (int t = 0; t < n_threads; ++t) pthread_create(&J[t].th, 0, ec_job, &J[t]);
            for (int t = 0; t < n_threads; ++t) pthread_join(J[t].th, 0);
        }
        for (int t = 0; t < n_threads; ++t) for (int j = 0; j < 11; ++j) stats[j] += J[t].W.stats[j];
    }

    oatk_tick(on_device ? "ec: correct reads (device)" : "ec: correct reads (host threads)");
    rebuild_syncmer_db(sr_db, g->scm_db);
    oatk_tick("ec: rebuild database");
    oatk_gpu_update_lists(sr_db, g->scm_db);            /* the device-resident lists follow the host's */
    oatk_tick("ec: refresh device lists");

    ec_print_summary(stats, verbose, cpu_at_entry, wall_at_entry);
    for (int t = 0; t < n_threads; ++t) {
        ec_worker_t *W = &J[t].W;
        wave_free(&W->w); free(W->S.cand.s); free(W->S.best_seq.s); free(W->S.path.a); free(W->S.best_path.a); free(W->S.stash);
        free(W->seq.s); free(W->kk.a); free(W->pp.a);
    }
    free(J);
}
