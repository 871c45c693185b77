/*
 * alignment_gpu.c -- row f3, first half: every read's syncmer list against the unitig graph
 * (reference alignment.c:177-594 for one read, 596-684 for the driver).
 *
 * What the reference computes, restated:
 *   hits       every (read syncmer, unitig occurrence) pair, as (oriented unitig, position on it, position on the
 *              read); the unitig is taken in the orientation that makes the syncmer's strand agree with the read
 *   links      inside one oriented unitig a hit points at the hit of the NEXT read position present on that unitig
 *              with the smallest larger unitig position
 *   fragments  following the links from every hit no earlier chain passed through gives a fragment (several
 *              fragments may share a tail); score = hits - max(unitig gaps, read gaps), kept when >= 0; hits
 *              without a link on either side are fragments of their own
 *   chains     fragments sorted by read interval; a fragment that reaches the end of its unitig may be followed
 *              by one that starts at position 0 of a unitig an arc leads to, if the read intervals overlap by
 *              exactly the arc's overlap; all best predecessors are kept
 *   records    every best-scoring chain that covers >= 90 % of the read's syncmers is one alignment record; the
 *              records of a read carry s = best score + 1 / (number of records)
 * The order of equal keys in the fragment sort is visible in the result (which predecessor lists form, in which
 * order chains are enumerated); the reference leaves it to libc's qsort and so does this file, with the same
 * comparator on the same input sequence.
 *
 * Differences in construction: reads are handed out to the threads in blocks from a shared counter instead of
 * one contiguous slice per thread (the records are put back in read order afterwards), predecessor lists live
 * in one arena per thread, and nothing is allocated per fragment.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <pthread.h>
#include "graph_gpu.h"

#define NO_LINK UINT32_MAX
#define READ_BLOCK 256
#define MIN_READ_FRAC .9          /* alignment.c:158 */

typedef struct {
    uint64_t uid, u_pos;
    uint32_t s_pos, next;
    int walked;
} hit_t;

typedef struct {
    uint64_t uid, u_beg, u_end;
    uint32_t s_beg, s_end, s_cnt;
    int64_t own, best;            /* score of the fragment alone / of the best chain that ends with it */
    int32_t p_head, p_tail;       /* its predecessors on best chains: a list in the arena, -1 = none */
} frag_t;

typedef struct { uint32_t frag; int32_t next; } pnode_t;

typedef struct { size_t n, m; scg_ra_t *a; } rec_v;

typedef struct {
    hit_t *hit; size_t n_hit, m_hit;
    frag_t *frg; size_t n_frg, m_frg;
    pnode_t *arena; size_t n_arena, m_arena;
    uint32_t *stack; size_t m_stack;
    uint32_t *path; size_t n_path, m_path;      /* chains found: length, then the fragments first to last */
    uint64_t n_mapped, n_unique;
} scratch_t;

typedef struct {
    sr_db_t *db;
    scg_t *g;
    const int64_t *old_ra;
    rec_v *block;                 /* records of each block of reads */
    uint64_t n_block;
    volatile uint64_t *next_block;
    scratch_t sc;
} job_t;

#define GROW(ptr, cap, need, type) do { \
    if ((need) > (cap)) { \
        (cap) = (cap) ? (cap) : 16; \
        while ((cap) < (need)) (cap) <<= 1; \
        (ptr) = (type *) realloc((ptr), sizeof(type) * (cap)); \
    } \
} while (0)

static int hit_cmp(const void *a, const void *b)
{
    const hit_t *x = (const hit_t *) a, *y = (const hit_t *) b;
    if (x->uid != y->uid) return x->uid < y->uid ? -1 : 1;
    if (x->s_pos != y->s_pos) return x->s_pos < y->s_pos ? -1 : 1;
    return (x->u_pos > y->u_pos) - (x->u_pos < y->u_pos);
}

static int frag_cmp(const void *a, const void *b)
{
    const frag_t *x = (const frag_t *) a, *y = (const frag_t *) b;
    if (x->s_beg != y->s_beg) return x->s_beg < y->s_beg ? -1 : 1;
    return (x->s_end > y->s_end) - (x->s_end < y->s_end);
}

static asmg_arc_t *live_arc(const asmg_t *g, uint64_t v, uint64_t w)
{
    asmg_arc_t *a = &g->arc[g->idx_p[v]];
    uint64_t i, n = g->idx_n[v];
    for (i = 0; i < n; ++i) if (a[i].w == w && !a[i].del) return &a[i];
    return 0;
}

static void add_frag(scratch_t *sc, uint64_t uid, uint32_t s_beg, uint32_t s_end, uint32_t s_cnt, uint64_t u_beg, uint64_t u_end, int64_t score)
{
    frag_t *f;
    GROW(sc->frg, sc->m_frg, sc->n_frg + 1, frag_t);
    f = &sc->frg[sc->n_frg++];
    f->uid = uid; f->u_beg = u_beg; f->u_end = u_end;
    f->s_beg = s_beg; f->s_end = s_end; f->s_cnt = s_cnt;
    f->own = f->best = score;
    f->p_head = f->p_tail = -1;
}

/* hits of one oriented unitig, [lo, hi) of the sorted list: links, then fragments (alignment.c:253-321) */
static void unitig_fragments(scratch_t *sc, size_t lo, size_t hi)
{
    hit_t *h = sc->hit;
    size_t a0, a1, b1, s, t, k;
    const uint64_t uid = h[lo].uid;
    /* consecutive read positions present on this unitig: [a0, a1) then [a1, b1) */
    for (a0 = lo, a1 = lo; a1 < hi && h[a1].s_pos == h[a0].s_pos; ++a1) {}
    while (a1 < hi) {
        for (b1 = a1; b1 < hi && h[b1].s_pos == h[a1].s_pos; ++b1) {}
        for (s = a0, t = a1; s < a1; ++s) {                       /* unitig positions ascend inside a group */
            while (t < b1 && h[t].u_pos <= h[s].u_pos) ++t;
            if (t < b1) h[s].next = (uint32_t) t;
        }
        a0 = a1; a1 = b1;
    }
    for (k = lo; k < hi; ++k) {
        int64_t u_gap = 0, s_gap = 0, score;
        uint32_t cnt = 1;
        if (h[k].walked || h[k].next == NO_LINK) continue;
        for (s = k; h[s].next != NO_LINK; s = t, ++cnt) {
            t = h[s].next;
            u_gap += (int64_t) (h[t].u_pos - h[s].u_pos) - 1;
            s_gap += (int64_t) h[t].s_pos - (int64_t) h[s].s_pos - 1;
            h[s].walked = 1;
        }
        h[s].walked = 1;
        if (u_gap < s_gap) u_gap = s_gap;
        score = (int64_t) cnt - u_gap;
        if (score >= 0) add_frag(sc, uid, h[k].s_pos, h[s].s_pos, cnt, h[k].u_pos, h[s].u_pos, score);
    }
    for (k = lo; k < hi; ++k)
        if (h[k].next == NO_LINK && !h[k].walked) add_frag(sc, uid, h[k].s_pos, h[k].s_pos, 1, h[k].u_pos, h[k].u_pos, 1);
}

static void add_pred(scratch_t *sc, frag_t *f, uint32_t j)
{
    int32_t id;
    GROW(sc->arena, sc->m_arena, sc->n_arena + 1, pnode_t);
    id = (int32_t) sc->n_arena++;
    sc->arena[id].frag = j; sc->arena[id].next = -1;
    if (f->p_tail < 0) f->p_head = id;
    else sc->arena[f->p_tail].next = id;
    f->p_tail = id;
}

/* all best chains that end with `node`, first fragment first (alignment.c:129-153) */
static void trace(scratch_t *sc, uint32_t node, uint32_t depth)
{
    int32_t p;
    GROW(sc->stack, sc->m_stack, (size_t) depth + 1, uint32_t);
    sc->stack[depth++] = node;
    if (sc->frg[node].p_head < 0) {
        uint32_t i;
        GROW(sc->path, sc->m_path, sc->n_path + depth + 1, uint32_t);
        sc->path[sc->n_path++] = depth;
        for (i = 0; i < depth; ++i) sc->path[sc->n_path++] = sc->stack[depth - 1 - i];
        return;
    }
    for (p = sc->frg[node].p_head; p >= 0; p = sc->arena[p].next) trace(sc, sc->arena[p].frag, depth);
}

static void align_read(scratch_t *sc, const scg_t *g, const sr_t *sr, int64_t old, rec_v *out)
{
    const asmg_t *ug = g->utg_asmg;
    const asmg_vtx_t *utg = ug->vtx;
    size_t j, k, m, lo, hi, p;
    int64_t best = 0;
    uint32_t n_rec = 0;

    sc->n_hit = 0;
    for (j = 0; j < sr->n; ++j) {
        const uint64_t scm = sr->k_mer[j] >> 1;
        const uint128_t *o = g->idx_u[scm], *end = g->idx_u[scm + 1];
        for (; o < end; ++o) {
            const uint64_t u = (uint64_t) (*o >> 36) & 0x3FFFFFFFFFFULL, pos = (uint64_t) *o & 0xFFFFFFFFFULL;
            const uint64_t t = ((uint64_t) (*o >> 78) & 1) ^ (sr->m_pos[j] & 1);
            hit_t *h;
            GROW(sc->hit, sc->m_hit, sc->n_hit + 1, hit_t);
            h = &sc->hit[sc->n_hit++];
            h->uid = u << 1 | t;
            h->u_pos = t ? utg[u].n - pos - 1 : pos;
            h->s_pos = (uint32_t) j;
            h->next = NO_LINK;
            h->walked = 0;
        }
    }
    if (sc->n_hit == 0) return;
    qsort(sc->hit, sc->n_hit, sizeof(hit_t), hit_cmp);

    sc->n_frg = 0;
    for (lo = 0; lo < sc->n_hit; lo = hi) {
        for (hi = lo + 1; hi < sc->n_hit && sc->hit[hi].uid == sc->hit[lo].uid; ++hi) {}
        unitig_fragments(sc, lo, hi);
    }
    if (sc->n_frg == 0) return;
    qsort(sc->frg, sc->n_frg, sizeof(frag_t), frag_cmp);

    /* chains across arcs (alignment.c:433-486): no clipping on either side of a junction, read overlap == arc overlap */
    m = sc->n_frg;
    sc->n_arena = 0;
    for (j = 0; j < m; ++j) {
        const frag_t *f = &sc->frg[j];
        const uint64_t last = f->s_end;
        const int64_t score = f->best;
        if (f->u_end + 1 != utg[f->uid >> 1].n) continue;
        for (k = j + 1; k < m; ++k) {
            frag_t *f1 = &sc->frg[k];
            const asmg_arc_t *arc;
            int64_t ovl, s1;
            if (f1->s_beg > last + 1) break;
            if (f1->u_beg != 0) continue;
            if (!(arc = live_arc(ug, f->uid, f1->uid))) continue;
            ovl = (int64_t) (arc->ln < last + 1 ? arc->ln : last + 1);
            if ((uint64_t) f1->s_beg + (uint64_t) ovl != last + 1) continue;
            s1 = score + f1->own - ovl;
            if (s1 <= score || s1 < f1->best || (s1 == f1->best && f1->p_head < 0)) continue;
            if (s1 > f1->best) { f1->best = s1; f1->p_head = f1->p_tail = -1; }
            add_pred(sc, f1, (uint32_t) j);
        }
    }
    for (j = 0; j < m; ++j) if (best < sc->frg[j].best) best = sc->frg[j].best;

    sc->n_path = 0;
    if (best >= (old >> 1))
        for (j = 0; j < m; ++j) if (sc->frg[j].best >= best) trace(sc, (uint32_t) j, 0);

    for (p = 0; p < sc->n_path; p += sc->path[p] + 1) {
        const uint32_t len = sc->path[p], *fr = &sc->path[p + 1];
        uint64_t covered = 0;
        scg_ra_t *r;
        for (k = 0; k < len; ++k) covered += sc->frg[fr[k]].s_cnt;
        if ((double) covered / sr->n < MIN_READ_FRAC) continue;
        GROW(out->a, out->m, out->n + 1, scg_ra_t);
        r = &out->a[out->n++];
        r->sid = sr->sid;
        r->n = len;
        r->a = (ra_frg_t *) malloc(sizeof(ra_frg_t) * len);
        for (k = 0; k < len; ++k) {
            const frag_t *f = &sc->frg[fr[k]];
            r->a[k].uid = f->uid; r->a[k].u_beg = f->u_beg; r->a[k].u_end = f->u_end;
            r->a[k].s_beg = f->s_beg; r->a[k].s_end = f->s_end;
        }
        ++n_rec;
    }
    for (k = 0; k < n_rec; ++k) out->a[out->n - 1 - k].s = 1.0 / n_rec + best;
    sc->n_mapped += n_rec > 0;
    sc->n_unique += n_rec == 1;
}

static void *align_worker(void *arg)
{
    job_t *job = (job_t *) arg;
    for (;;) {
        const uint64_t b = __sync_fetch_and_add(job->next_block, 1);
        uint64_t i, end;
        if (b >= job->n_block) break;
        end = (b + 1) * READ_BLOCK < job->db->n ? (b + 1) * READ_BLOCK : job->db->n;
        for (i = b * READ_BLOCK; i < end; ++i)
            if ((job->old_ra[i] & 1) && job->db->a[i].n) align_read(&job->sc, job->g, &job->db->a[i], job->old_ra[i], &job->block[b]);
    }
    return 0;
}

static void drop_records(scg_ra_v *ra_v)
{
    size_t i;
    for (i = 0; i < ra_v->n; ++i) free(ra_v->a[i].a);
    free(ra_v->a);
    ra_v->a = 0; ra_v->n = ra_v->m = 0;
}

void scg_ra_v_destroy(scg_ra_v *ra_v)
{
    if (!ra_v) return;
    drop_records(ra_v);
    free(ra_v);
}

void scg_read_alignment(sr_db_t *sr_db, scg_ra_v *ra_v, scg_t *g, int n_threads, int for_unzip)
{
    uint64_t i, n_live = 0, n_block, n_rec = 0, n_reads = 0, n_mapped = 0, n_unique = 0;
    volatile uint64_t next_block = 0;
    int64_t *old_ra;
    rec_v *block;
    job_t *job;
    pthread_t *th;
    int t;

    if (sr_db->n == 0 || !g->utg_asmg) return;
    for (i = 0; i < g->utg_asmg->n_vtx; ++i) n_live += !g->utg_asmg->vtx[i].del;
    if (!n_live) return;
    if (n_threads < 1) n_threads = 1;

    /* which reads to align and the score a new alignment has to reach (alignment.c:611-633) */
    old_ra = (int64_t *) calloc(sr_db->n, sizeof(int64_t));
    if (for_unzip && ra_v->n > 0) {
        for (i = 0; i < ra_v->n; ++i) {
            const scg_ra_t *r = &ra_v->a[i];
            double whole, frac;
            if (r->n <= 2 || (old_ra[r->sid] & 1)) continue;
            frac = modf(r->s, &whole);
            if (frac < DBL_EPSILON) whole -= 1;                     /* a single record: s = score + 1 */
            old_ra[r->sid] = (int64_t) ((uint64_t) whole << 1 | 1);
        }
    } else for (i = 0; i < sr_db->n; ++i) old_ra[i] = 1;

    n_block = (sr_db->n + READ_BLOCK - 1) / READ_BLOCK;
    block = (rec_v *) calloc(n_block, sizeof(rec_v));
    job = (job_t *) calloc(n_threads, sizeof(job_t));
    th = (pthread_t *) calloc(n_threads, sizeof(pthread_t));
    for (t = 0; t < n_threads; ++t) {
        job[t].db = sr_db; job[t].g = g; job[t].old_ra = old_ra;
        job[t].block = block; job[t].n_block = n_block; job[t].next_block = &next_block;
    }
    for (t = 1; t < n_threads; ++t) pthread_create(&th[t], 0, align_worker, &job[t]);
    align_worker(&job[0]);
    for (t = 1; t < n_threads; ++t) pthread_join(th[t], 0);

    drop_records(ra_v);
    for (i = 0; i < n_block; ++i) n_rec += block[i].n;
    ra_v->a = (scg_ra_t *) malloc(sizeof(scg_ra_t) * (n_rec ? n_rec : 1));
    ra_v->m = n_rec;
    for (i = 0; i < n_block; ++i) {
        if (block[i].n) memcpy(ra_v->a + ra_v->n, block[i].a, sizeof(scg_ra_t) * block[i].n);
        ra_v->n += block[i].n;
        free(block[i].a);
    }
    for (i = 0; i < sr_db->n; ++i) n_reads += sr_db->a[i].n > 0;
    for (t = 0; t < n_threads; ++t) {
        scratch_t *sc = &job[t].sc;
        n_mapped += sc->n_mapped; n_unique += sc->n_unique;
        free(sc->hit); free(sc->frg); free(sc->arena); free(sc->stack); free(sc->path);
    }
    fprintf(stderr, "[M::%s] %lu mappable reads, %lu mapped (%lu unique mapping)\n", __func__,
            (unsigned long) n_reads, (unsigned long) n_mapped, (unsigned long) n_unique);
    free(th); free(job); free(block); free(old_ra);
}
