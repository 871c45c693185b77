/*
 * coverage_gpu.c -- row f3, second half: unitig and arc coverages from the read alignments
 * (reference syncasm.c:1882-2065 scg_ra_utg_coverage, 2067-2147 scg_ra_arc_coverage,
 * 2149-2261 scg_refine_arc_coverage; helpers 584-628, 1643-1880; graph.c:117-127, 237-248).
 *
 * Unitig coverage, three estimates in a row, each the outlier-trimmed mean (1.5 IQR fences) of per-syncmer values:
 *   1. reads with exactly one alignment record: how many of them cover each unitig position
 *   2. all aligned reads: every record of a read is cut into blocks in which all its records match the read
 *      (longest common subsequences of read and unitig syncmers, intersected over the records); a block of
 *      n syncmers is shared among the records' unitigs in proportion to their current estimates, iterated to a
 *      fixed point (at most 1000 rounds)
 *   3. every syncmer's k-mer coverage shared among its unitig occurrences in proportion to estimate 2
 * Arc coverage: the summed weights (1 / number of records) of the reads that walk an arc between two fragments
 * that each hold a syncmer unique in the graph; where the same syncmer pair also occurs on other arcs or inside
 * unitigs the value is scaled by this arc's share of the neighbouring unitig coverages; finally no arc may have
 * more than either of its unitigs.
 *
 * All sums run in the reference's order (doubles are not associative); two reference habits are kept because
 * they are visible in the numbers: fragment coordinates on a reverse-strand unitig are counted from the unitig's
 * other end, yet they index the forward syncmer list as they are (syncasm.c:1937-1939, 1789-1793, 2101-2103).
 * Position counts of estimate 1 are accumulated as integers (exact in doubles either way).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <limits.h>
#include "graph_gpu.h"

#define EM_ROUNDS 1000

typedef struct { uint32_t beg, len; } blk_t;
typedef struct { size_t n, m; blk_t *a; } blk_v;

#define PUSH(v, type, x) do { \
    if ((v).n == (v).m) { (v).m = (v).m ? (v).m << 1 : 16; (v).a = (type *) realloc((v).a, sizeof(type) * (v).m); } \
    (v).a[(v).n++] = (x); \
} while (0)

static int dbl_cmp(const void *a, const void *b)
{
    const double x = *(const double *) a, y = *(const double *) b;
    return (x > y) - (x < y);
}

/* linear-interpolated quantile of a sorted array */
static double quantile(const double *a, int n, double q)
{
    double whole, frac;
    int i;
    if (n == 1) return a[0];
    frac = modf(q * (n - 1), &whole);
    i = (int) lround(whole);
    if (i == n - 1) return a[i];
    return a[i] + (a[i + 1] - a[i]) * frac;
}

static double trimmed_mean(double *a, int n, int sorted)
{
    double lo, hi, iqr, sum = 0.;
    int i, kept = 0;
    if (n == 0) return 0.;
    if (!sorted) qsort(a, n, sizeof(double), dbl_cmp);
    lo = quantile(a, n, 0.25);
    hi = quantile(a, n, 0.75);
    iqr = hi - lo;
    lo -= 1.5 * iqr;
    hi += 1.5 * iqr;
    for (i = 0; i < n; ++i) if (a[i] >= lo && a[i] <= hi) { ++kept; sum += a[i]; }
    return kept ? sum / kept : 0.;
}

uint64_t asmg_max_link_id(asmg_t *g)
{
    uint64_t i, m = 0;
    for (i = 0; i < g->n_arc; ++i) if (g->arc[i].link_id > m) m = g->arc[i].link_id;
    return m;
}

void asmg_arc_fix_cov(asmg_t *g)
{
    uint64_t i;
    for (i = 0; i < g->n_arc; ++i) {
        asmg_arc_t *a = &g->arc[i];
        uint32_t cv, cw;
        if (a->del) continue;
        cv = g->vtx[a->v >> 1].cov; cw = g->vtx[a->w >> 1].cov;
        if (cw < cv) cv = cw;
        if (cv < a->cov) a->cov = cv;
    }
}

/* every unitig's coverage from its syncmers' k-mer coverages: the trimmed mean over the syncmers that occur once in
 * the graph, over all of them if there is none (syncasm.c:630-703) */
void scg_update_utg_cov(scg_t *scg)
{
    asmg_t *ug = scg->utg_asmg;
    const syncmer_t *scm = scg->scm_db->a;
    uint64_t i, j;
    for (i = 0; i < ug->n_vtx; ++i) {
        asmg_vtx_t *x = &ug->vtx[i];
        double *c, mean;
        uint64_t zero = 0;
        if (x->del) { x->cov = 0; continue; }
        c = (double *) calloc(x->n ? x->n : 1, sizeof(double));
        for (j = 0; j < x->n; ++j) {
            const uint64_t u = x->a[j] >> 1;
            if (scg->idx_u[u + 1] - scg->idx_u[u] == 1) c[j] = scm[u].cov;
        }
        qsort(c, x->n, sizeof(double), dbl_cmp);
        while (zero < x->n && c[zero] < DBL_EPSILON) ++zero;
        if (zero == x->n) {
            for (j = 0; j < x->n; ++j) c[j] = scm[x->a[j] >> 1].cov;
            qsort(c, x->n, sizeof(double), dbl_cmp);
            zero = 0;
        }
        mean = trimmed_mean(c + zero, (int) (x->n - zero), 1);
        x->cov = mean;
        free(c);
    }
}

/* ---------- matching blocks of one fragment: LCS of read and unitig syncmer ids (syncasm.c:1654-1751) ---------- */
static void match_blocks(const uint64_t *rd, int n_rd, const uint64_t *ut, int n_ut, int rev_ut, uint32_t offset, blk_v *out)
{
    /* ut is read backwards when rev_ut: U(j) is the j-th syncmer in read direction */
#define R(i) (rd[(i)] >> 1)
#define U(j) ((rev_ut ? ut[n_ut - 1 - (j)] : ut[(j)]) >> 1)
    const size_t first = out->n;
    int head = 0, r_end = n_rd - 1, u_end = n_ut - 1, nr, nu, i, j;
    size_t p, k;
    while (head < n_rd && head < n_ut && R(head) == U(head)) ++head;
    while (head <= r_end && head <= u_end && R(r_end) == U(u_end)) { --r_end; --u_end; }
    if (head > 0) { blk_t b = {offset, (uint32_t) head}; PUSH(*out, blk_t, b); }
    nr = r_end - head + 1; nu = u_end - head + 1;
    if (nr > 0 && nu > 0) {
        int *L = (int *) calloc((size_t) (nr + 1) * (nu + 1), sizeof(int));
        const size_t w = (size_t) nu + 1, mark = out->n;
        for (i = 1; i <= nr; ++i)
            for (j = 1; j <= nu; ++j)
                L[i * w + j] = R(head + i - 1) == U(head + j - 1) ? L[(i - 1) * w + j - 1] + 1
                             : (L[(i - 1) * w + j] > L[i * w + j - 1] ? L[(i - 1) * w + j] : L[i * w + j - 1]);
        /* walk back: a match is always taken; otherwise left only if strictly better than up */
        for (i = nr, j = nu; i > 0 && j > 0; ) {
            if (R(head + i - 1) == U(head + j - 1)) {
                blk_t b = {offset + (uint32_t) (head + i - 1), 1};
                PUSH(*out, blk_t, b);
                --i; --j;
            } else if (L[i * w + j - 1] > L[(i - 1) * w + j]) --j;
            else --i;
        }
        free(L);
        for (p = mark, k = out->n; p + 1 < k; ++p) { blk_t t = out->a[p]; out->a[p] = out->a[--k]; out->a[k] = t; }
    }
    if (nr < 0) nr = 0;
    if (head + nr < n_rd) { blk_t b = {offset + (uint32_t) (head + nr), (uint32_t) (n_rd - head - nr)}; PUSH(*out, blk_t, b); }
    /* join blocks that touch */
    if (out->n > first + 1) {
        for (p = first, k = first + 1; k < out->n; ++k) {
            if (out->a[p].beg + out->a[p].len == out->a[k].beg) out->a[p].len += out->a[k].len;
            else out->a[++p] = out->a[k];
        }
        out->n = p + 1;
    }
#undef R
#undef U
}

/* blocks of one read in which all its n records match: lengths into len_v, n unitig ids per block into utg_v
 * (syncasm.c:1756-1878) */
typedef struct { size_t n, m; uint32_t *a; } u32_v;
typedef struct { size_t n, m; uint64_t *a; } u64_v;

static void read_blocks(const scg_t *g, const sr_t *sr, const scg_ra_t *ra, uint32_t n, u32_v *len_v, u64_v *utg_v, uint32_t *n_blk)
{
    blk_v *bl = (blk_v *) calloc(n, sizeof(blk_v));
    uint64_t *uid = (uint64_t *) malloc(sizeof(uint64_t) * 5 * n), *frg = uid + n, *cur = frg + n, *beg = cur + n, *len = beg + n;
    uint32_t i, j;
    *n_blk = 0;
    for (i = 0; i < n; ++i)
        for (j = 0; j < ra[i].n; ++j) {
            const ra_frg_t *f = &ra[i].a[j];
            match_blocks(&sr->k_mer[f->s_beg], (int) (f->s_end - f->s_beg + 1), &g->utg_asmg->vtx[f->uid >> 1].a[f->u_beg],
                    (int) (f->u_end - f->u_beg + 1), (int) (f->uid & 1), f->s_beg, &bl[i]);
        }
#define LOAD(i) do { \
    beg[i] = bl[i].a[cur[i]].beg; len[i] = bl[i].a[cur[i]].len; \
    while (ra[i].a[frg[i]].s_end < beg[i]) ++frg[i]; \
    uid[i] = ra[i].a[frg[i]].uid >> 1; \
} while (0)
    for (i = 0; i < n; ++i) {
        if (bl[i].n == 0) goto done;
        frg[i] = cur[i] = 0;
        LOAD(i);
    }
    for (;;) {
        uint64_t left = 0;
        int ext, min_ext = INT_MAX;
        for (i = 0; i < n; ++i) if (beg[i] > left) left = beg[i];
        for (i = 0; i < n; ++i) {
            ext = (int) (len[i] - left + beg[i]);
            if (ext < min_ext) min_ext = ext;
        }
        if (min_ext > 0) {
            PUSH(*len_v, uint32_t, (uint32_t) min_ext);
            for (i = 0; i < n; ++i) PUSH(*utg_v, uint64_t, uid[i]);
            ++*n_blk;
            for (i = 0; i < n; ++i) {
                ext = (int) (len[i] - left + beg[i]);
                if (ext == min_ext) {
                    if (++cur[i] == bl[i].n) goto done;
                    LOAD(i);
                } else { beg[i] = left + min_ext; len[i] = ext - min_ext; }
            }
        } else {
            for (i = 0, j = 1; j < n; ++j) if (beg[j] < beg[i]) i = j;      /* leftmost block moves on */
            if (++cur[i] == bl[i].n) goto done;
            LOAD(i);
        }
    }
#undef LOAD
done:
    for (i = 0; i < n; ++i) free(bl[i].a);
    free(bl); free(uid);
}

void scg_ra_utg_coverage(scg_t *g, sr_db_t *sr_db, scg_ra_v *ra_v, int verbose)
{
    asmg_t *ug = g->utg_asmg;
    const uint64_t n_vtx = ug->n_vtx, n_scm = g->scm_db->n;
    const syncmer_t *scm = g->scm_db->a;
    uint64_t i, j, k, total = 0, *first;
    int64_t *delta;
    double *est, *val, *acc, whole;
    u32_v blk_len = {0, 0, 0}, rd_nblk = {0, 0, 0}, rd_naln = {0, 0, 0};
    u64_v blk_utg = {0, 0, 0};
    int round;

    if (ra_v->n == 0) {
        fprintf(stderr, "[W::%s] no read alignment, unitig coverage estimation skipped\n", __func__);
        return;
    }
    first = (uint64_t *) malloc(sizeof(uint64_t) * (n_vtx + 1));       /* where each unitig's positions start */
    for (i = 0; i < n_vtx; ++i) { first[i] = total; total += ug->vtx[i].n; }
    first[n_vtx] = total;
    est = (double *) calloc(n_vtx ? n_vtx : 1, sizeof(double));
    val = (double *) calloc(total ? total : 1, sizeof(double));

    /* 1: uniquely aligned reads per position (a difference array with one spare slot per unitig) */
    delta = (int64_t *) calloc(total + n_vtx + 1, sizeof(int64_t));
    for (i = 0; i < ra_v->n; ++i) {
        const scg_ra_t *r = &ra_v->a[i];
        if (modf(r->s, &whole) > DBL_EPSILON) continue;
        for (j = 0; j < r->n; ++j) {
            const uint64_t u = r->a[j].uid >> 1, b = first[u] + u;
            ++delta[b + r->a[j].u_beg];
            --delta[b + r->a[j].u_end + 1];
        }
    }
    for (i = 0; i < n_vtx; ++i) {
        int64_t run = 0;
        const uint64_t n = ug->vtx[i].n;
        double *c = val + first[i];
        uint64_t zero = 0;
        for (j = 0; j < n; ++j) { run += delta[first[i] + i + j]; c[j] = (double) run; }
        qsort(c, n, sizeof(double), dbl_cmp);
        while (zero < n && c[zero] < DBL_EPSILON) ++zero;
        est[i] = trimmed_mean(c + zero, (int) (n - zero), 1);
        if (est[i] < 1.) est[i] = 1.;
    }
    free(delta);

    /* 2: blocks of every aligned read (records of a read are adjacent), then the fixed point */
    for (i = 0; i < ra_v->n; i = j) {
        uint32_t nb;
        for (j = i + 1; j < ra_v->n && ra_v->a[j].sid == ra_v->a[i].sid; ++j) {}
        read_blocks(g, &sr_db->a[ra_v->a[i].sid], &ra_v->a[i], (uint32_t) (j - i), &blk_len, &blk_utg, &nb);
        PUSH(rd_nblk, uint32_t, nb);
        PUSH(rd_naln, uint32_t, (uint32_t) (j - i));
    }
    acc = (double *) malloc(sizeof(double) * (n_vtx ? n_vtx : 1));
    for (round = 0; round < EM_ROUNDS; ++round) {
        const uint32_t *bl = blk_len.a;
        const uint64_t *bu = blk_utg.a;
        double diff = 0.;
        memset(acc, 0, sizeof(double) * n_vtx);
        for (i = 0; i < rd_nblk.n; ++i) {
            const uint32_t na = rd_naln.a[i];
            for (k = 0; k < rd_nblk.a[i]; ++k, ++bl, bu += na) {
                double share = 0.;
                for (j = 0; j < na; ++j) share += est[bu[j]];
                if (share == 0.) continue;
                for (j = 0; j < na; ++j) acc[bu[j]] += est[bu[j]] / share * *bl;
            }
        }
        for (i = 0; i < n_vtx; ++i) {
            const double c = acc[i] / ug->vtx[i].n;
            diff += fabs(c - est[i]);
            est[i] = c;
        }
        if (verbose > 2) fprintf(stderr, "[M::%s] unitig coverage estimation iteration %d: diff = %.6f\n", __func__, round, diff);
        if (diff < DBL_EPSILON) break;
    }
    if (verbose > 2) fprintf(stderr, "[M::%s] unitig coverage estimation ended at iteration %d\n", __func__, round);
    free(acc);

    /* 3: k-mer coverages shared among the occurrences */
    memset(val, 0, sizeof(double) * total);
    for (i = 0; i < n_scm; ++i) {
        const uint128_t *o = g->idx_u[i], *end = g->idx_u[i + 1], *q;
        double share = 0.;
        if (o == end) continue;
        for (q = o; q < end; ++q) share += est[(uint64_t) (*q >> 36) & 0x3FFFFFFFFFFULL];
        if (share < DBL_EPSILON) continue;
        for (q = o; q < end; ++q) {
            const uint64_t u = (uint64_t) (*q >> 36) & 0x3FFFFFFFFFFULL;
            val[first[u] + ((uint64_t) *q & 0xFFFFFFFFFULL)] = est[u] / share * scm[i].cov;
        }
    }
    for (i = 0; i < n_vtx; ++i) {
        double c = trimmed_mean(val + first[i], (int) ug->vtx[i].n, 0);
        if (c < 1.) c = 1.;
        ug->vtx[i].cov = (uint32_t) c;
    }
    free(blk_len.a); free(blk_utg.a); free(rd_nblk.a); free(rd_naln.a);
    free(val); free(est); free(first);
}

static asmg_arc_t *any_arc(const asmg_t *g, uint64_t v, uint64_t w, int live_only)
{
    asmg_arc_t *a = &g->arc[g->idx_p[v]];
    uint64_t i, n = g->idx_n[v];
    for (i = 0; i < n; ++i) if (a[i].w == w && !(live_only && a[i].del)) return &a[i];
    return 0;
}

#define ARC_ID(a) ((a)->link_id << 1 | (a)->comp)

void scg_ra_arc_coverage(scg_t *g, sr_db_t *sr_db, scg_ra_v *ra_v, int refine, int verbose)
{
    asmg_t *ug = g->utg_asmg;
    const uint64_t n_id = 2 * (asmg_max_link_id(ug) + 1);
    double *w = (double *) calloc(n_id, sizeof(double)), whole;
    uint8_t *seen = (uint8_t *) calloc(n_id, 1), *anchored = 0;
    size_t m_anchored = 0;
    uint64_t i, j, s;
    (void) sr_db;

    for (i = 0; i < ra_v->n; ++i) {
        const scg_ra_t *r = &ra_v->a[i];
        double weight;
        if (r->n < 2) continue;
        weight = modf(r->s, &whole);
        if (weight < DBL_EPSILON) weight = 1.0;
        if (r->n > m_anchored) { m_anchored = r->n; anchored = (uint8_t *) realloc(anchored, m_anchored); }
        if (weight < .99) {
            /* several records: only fragments that hold a syncmer found nowhere else on the graph count */
            for (j = 0; j < r->n; ++j) {
                const uint64_t *a = ug->vtx[r->a[j].uid >> 1].a;
                anchored[j] = 0;
                for (s = r->a[j].u_beg; s <= r->a[j].u_end; ++s)
                    if (g->idx_u[(a[s] >> 1) + 1] - g->idx_u[a[s] >> 1] == 1) { anchored[j] = 1; break; }
            }
        } else memset(anchored, 1, r->n);
        for (j = 1; j < r->n; ++j) {
            const asmg_arc_t *arc = any_arc(ug, r->a[j - 1].uid, r->a[j].uid, 0);
            uint64_t id, cid;
            if (!anchored[j - 1] || !anchored[j]) continue;
            id = ARC_ID(arc);
            cid = ((arc->v ^ 1) != arc->w || (arc->w ^ 1) != arc->v) ? id ^ 1 : id;
            if (!seen[id]) { seen[id] = seen[cid] = 1; w[id] = weight; w[cid] = weight; }
            else { w[id] += weight; seen[cid] = 1; w[cid] += weight; }
        }
    }
    for (i = 0; i < ug->n_arc; ++i) {
        asmg_arc_t *arc = &ug->arc[i];
        if (arc->del) continue;
        arc->cov = (uint32_t) (seen[ARC_ID(arc)] ? w[ARC_ID(arc)] : 0);
    }
    free(w); free(seen); free(anchored);
    if (refine) scg_refine_arc_coverage(g, verbose);
    else asmg_arc_fix_cov(ug);
}

/* ---------- refinement: arcs and unitig-internal steps that join the same syncmer pair ---------- */
typedef struct { uint64_t v, w, order, link; } pair_key_t;

static int pair_cmp(const void *a, const void *b)
{
    const pair_key_t *x = (const pair_key_t *) a, *y = (const pair_key_t *) b;
    if (x->v != y->v) return x->v < y->v ? -1 : 1;
    if (x->w != y->w) return x->w < y->w ? -1 : 1;
    return (x->order > y->order) - (x->order < y->order);
}

static void canon_pair(uint64_t *v, uint64_t *w)
{
    if (*v > *w) { const uint64_t t = *v ^ 1; *v = *w ^ 1; *w = t; }
}

/* the syncmer pair an arc joins: last syncmer of v, first of w, both in walking direction */
static void arc_pair(const asmg_t *g, const asmg_arc_t *a, uint64_t *v, uint64_t *w)
{
    const asmg_vtx_t *x = &g->vtx[a->v >> 1], *y = &g->vtx[a->w >> 1];
    *v = (a->v & 1) ? x->a[0] ^ 1 : x->a[x->n - 1];
    *w = (a->w & 1) ? y->a[y->n - 1] ^ 1 : y->a[0];
    canon_pair(v, w);
}

/* group of a pair = link id of the first arc (in arc order) that joins it; UINT64_MAX if no arc does */
static uint64_t pair_group(const pair_key_t *key, size_t n, uint64_t v, uint64_t w)
{
    size_t lo = 0, hi = n;
    while (lo < hi) {
        const size_t mid = (lo + hi) >> 1;
        if (key[mid].v < v || (key[mid].v == v && key[mid].w < w)) lo = mid + 1;
        else hi = mid;
    }
    return lo < n && key[lo].v == v && key[lo].w == w ? key[lo].link : UINT64_MAX;
}

void scg_refine_arc_coverage(scg_t *g, int verbose)
{
    asmg_t *ug = g->utg_asmg;
    const uint64_t n_link = asmg_max_link_id(ug) + 1;
    pair_key_t *key = (pair_key_t *) malloc(sizeof(pair_key_t) * (ug->n_arc ? ug->n_arc : 1));
    uint64_t *members = (uint64_t *) calloc(n_link, sizeof(uint64_t)), *total = (uint64_t *) calloc(n_link, sizeof(uint64_t));
    uint64_t *own = (uint64_t *) calloc(n_link, sizeof(uint64_t));     /* per link id: the value its own arc put in */
    size_t n_key = 0;
    uint64_t i, j, v, w, grp;

    for (i = 0; i < ug->n_arc; ++i) {
        const asmg_arc_t *a = &ug->arc[i];
        if (a->del || a->comp) continue;
        arc_pair(ug, a, &key[n_key].v, &key[n_key].w);
        key[n_key].order = i; key[n_key].link = a->link_id;
        ++n_key;
    }
    qsort(key, n_key, sizeof(pair_key_t), pair_cmp);
    /* every arc of a pair reports to the pair's first link id, with the mean coverage of its two unitigs */
    for (i = 0; i < ug->n_arc; ++i) {
        const asmg_arc_t *a = &ug->arc[i];
        uint64_t c;
        if (a->del || a->comp) continue;
        arc_pair(ug, a, &v, &w);
        grp = pair_group(key, n_key, v, w);
        c = ((uint64_t) ug->vtx[a->v >> 1].cov + ug->vtx[a->w >> 1].cov) / 2;
        ++members[grp]; total[grp] += c;
        own[a->link_id] = c;
    }
    /* and so does every step inside a unitig that joins the same pair, with the unitig's coverage */
    for (i = 0; i < ug->n_vtx; ++i) {
        const uint64_t *a = ug->vtx[i].a;
        for (j = 1; j < ug->vtx[i].n; ++j) {
            v = a[j - 1]; w = a[j];
            canon_pair(&v, &w);
            if ((grp = pair_group(key, n_key, v, w)) == UINT64_MAX) continue;
            ++members[grp]; total[grp] += ug->vtx[i].cov;
        }
    }
    for (i = 0; i < ug->n_arc; ++i) {
        asmg_arc_t *a = &ug->arc[i], *c;
        uint64_t cov;
        if (a->del || a->comp) continue;
        arc_pair(ug, a, &v, &w);
        grp = pair_group(key, n_key, v, w);
        if (members[grp] == 1 || total[grp] == 0) continue;
        cov = (uint64_t) lround((double) a->cov / total[grp] * own[a->link_id]);
        if (verbose > 2)
            fprintf(stderr, "[M::%s] arc u%lu%c -> u%lu%c coverage updated: %u -> %lu\n", __func__, (unsigned long) (a->v >> 1),
                    "+-"[a->v & 1], (unsigned long) (a->w >> 1), "+-"[a->w & 1], a->cov, (unsigned long) cov);
        a->cov = cov;
        if ((c = any_arc(ug, a->w ^ 1, a->v ^ 1, 1))) c->cov = cov;
    }
    asmg_arc_fix_cov(ug);
    free(key); free(members); free(total); free(own);
}
