/*
 * graph_gpu.c -- rows a7-a9 on the host, over the device arc tally. See graph_gpu.h.
 *
 * Written against the behaviour of the reference, not its text: adjacency is a
 * CSR view (idx_p / idx_n per oriented vertex) over the arc array sorted by
 * (v, w); "live" means not flagged del. Where the reference's result depends on
 * libc (the unstable qsort of the arc array, graph.c:70-83) the same libc call is
 * made on an identically laid out array.
 */
#include <stdlib.h>
#include <string.h>
#include "syncgpu.h"
#include "graph_gpu.h"

#define NO_ID UINT64_MAX

/* ---------- adjacency helpers ---------- */
static inline asmg_arc_t *out_arcs(const asmg_t *g, uint64_t v) { return &g->arc[g->idx_p[v]]; }
static inline uint64_t out_n(const asmg_t *g, uint64_t v) { return g->idx_n[v]; }

static uint64_t live_out(const asmg_t *g, uint64_t v)
{
    uint64_t i, n = out_n(g, v), c = 0;
    const asmg_arc_t *a = out_arcs(g, v);
    for (i = 0; i < n; ++i) c += !a[i].del;
    return c;
}

static asmg_arc_t *first_live(const asmg_t *g, uint64_t v)
{
    uint64_t i, n = out_n(g, v);
    asmg_arc_t *a = out_arcs(g, v);
    for (i = 0; i < n; ++i) if (!a[i].del) return &a[i];
    return 0;
}

static asmg_arc_t *find_arc(const asmg_t *g, uint64_t v, uint64_t w, int live_only)
{
    uint64_t i, n = out_n(g, v);
    asmg_arc_t *a = out_arcs(g, v);
    for (i = 0; i < n; ++i) if (a[i].w == w && !(live_only && a[i].del)) return &a[i];
    return 0;
}

static void flag_arcs(asmg_t *g, uint64_t v, uint64_t w)            /* every v->w arc becomes deleted */
{
    uint64_t i, n = out_n(g, v);
    asmg_arc_t *a = out_arcs(g, v);
    for (i = 0; i < n; ++i) if (a[i].w == w) a[i].del = 1;
}

void asmg_destroy(asmg_t *g)
{
    uint64_t i;
    if (!g) return;
    for (i = 0; i < g->n_vtx; ++i) { free(g->vtx[i].seq); free(g->vtx[i].a); }
    free(g->vtx); free(g->arc); free(g->idx_p); free(g->idx_n);
    free(g);
}

/* ---------- a8: asmg_finalize ---------- */
static int arc_vw_cmp(const void *x, const void *y)
{
    const asmg_arc_t *a = (const asmg_arc_t *) x, *b = (const asmg_arc_t *) y;
    if (a->v != b->v) return a->v < b->v ? -1 : 1;
    return (a->w > b->w) - (a->w < b->w);
}

/* Arcs that are already in strictly ascending (v, w) order stay as they are: with no two equal keys every sort gives
 * this one order (the device hands the arc tally over sorted, which saves sorting 10^7 48-byte arcs on the host). Equal
 * keys (graph.c:252 "multi-arc") leave the order to libc's qsort, as in the reference. */
void asmg_arc_sort(asmg_t *g)
{
    uint64_t i;
    for (i = 1; i < g->n_arc; ++i) if (arc_vw_cmp(&g->arc[i - 1], &g->arc[i]) >= 0) break;
    if (i < g->n_arc) qsort(g->arc, g->n_arc, sizeof(asmg_arc_t), arc_vw_cmp);
}

void asmg_arc_index(asmg_t *g)
{
    uint64_t i, j;
    free(g->idx_p); free(g->idx_n);
    g->idx_p = (uint64_t *) calloc(g->n_vtx * 2 + 1, sizeof(uint64_t));
    g->idx_n = (uint64_t *) calloc(g->n_vtx * 2 + 1, sizeof(uint64_t));
    for (i = 0; i < g->n_arc; i = j) {
        for (j = i + 1; j < g->n_arc && g->arc[j].v == g->arc[i].v; ++j) {}
        g->idx_p[g->arc[i].v] = i;
        g->idx_n[g->arc[i].v] = j - i;
    }
}

/* drop deleted vertices (and arcs that are deleted or touch one), renumber the survivors in order */
static void compact(asmg_t *g)
{
    uint64_t i, j, n = g->n_vtx, *map = (uint64_t *) malloc(sizeof(uint64_t) * (n + 1));
    for (i = j = 0; i < n; ++i) {
        if (g->vtx[i].del) { free(g->vtx[i].seq); free(g->vtx[i].a); map[i] = NO_ID; continue; }
        if (j < i) g->vtx[j] = g->vtx[i];
        map[i] = j++;
    }
    g->n_vtx = g->m_vtx = j;
    g->vtx = (asmg_vtx_t *) realloc(g->vtx, sizeof(asmg_vtx_t) * (j ? j : 1));
    for (i = j = 0; i < g->n_arc; ++i) {
        asmg_arc_t *a = &g->arc[i];
        if (a->del || map[a->v >> 1] == NO_ID || map[a->w >> 1] == NO_ID) continue;
        if (j < i) g->arc[j] = *a;
        g->arc[j].v = map[a->v >> 1] << 1 | (a->v & 1);
        g->arc[j].w = map[a->w >> 1] << 1 | (a->w & 1);
        ++j;
    }
    g->n_arc = g->m_arc = j;
    g->arc = (asmg_arc_t *) realloc(g->arc, sizeof(asmg_arc_t) * (j ? j : 1));
    free(map);
}

/* every arc must have its complement w^1 -> v^1, flagged the other way round (graph.c:205-235) */
static uint32_t repair_symmetry(asmg_t *g)
{
    uint32_t added = 0;
    uint64_t i, n = g->n_arc;
    for (i = 0; i < n; ++i) {
        asmg_arc_t *a = &g->arc[i], *c;
        if (a->del) continue;
        c = find_arc(g, a->w ^ 1, a->v ^ 1, 1);
        if (!c) {
            asmg_arc_t x = *a;
            if (g->n_arc == g->m_arc) {                      /* kroundup64(++m), as MYEXPAND does */
                uint64_t m = g->m_arc + 1;
                --m; m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16; m |= m >> 32; ++m;
                g->m_arc = m;
                g->arc = (asmg_arc_t *) realloc(g->arc, sizeof(asmg_arc_t) * m);
                a = &g->arc[i];
            }
            x.v = a->w ^ 1; x.w = a->v ^ 1; x.del = 0; x.comp = a->comp ^ 1;
            g->arc[g->n_arc++] = x;
            ++added;
        } else {
            c->comp = a->comp ^ 1;
            if (a->ln != c->ln) a->ln = c->ln = a->ln < c->ln ? a->ln : c->ln;
            if (a->ls != c->ls) a->ls = c->ls = a->ls < c->ls ? a->ls : c->ls;
        }
    }
    return added;
}

void asmg_shrink_link_id(asmg_t *g)
{
    uint64_t i, next = 0;
    const uint64_t PENDING = 0x8000000000000000ULL;
    for (i = 0; i < g->n_arc; ++i) g->arc[i].link_id |= PENDING;
    for (i = 0; i < g->n_arc; ++i) {
        asmg_arc_t *a = &g->arc[i];
        if (!(a->link_id & PENDING)) continue;
        a->link_id = next;
        find_arc(g, a->w ^ 1, a->v ^ 1, 0)->link_id = next;      /* an arc and its complement share the id */
        ++next;
    }
}

/* Large graphs (the all-syncmer graph of the error-correction step: ~10^7 arcs) spend their time in the two passes
 * that look every arc's complement up. The look-up is independent per arc, so it is done once, in parallel, into an
 * index; an arc array that is already symmetric and consistent (what the device tally delivers) then needs no repair
 * pass at all, and the link ids are handed out from the index in one sequential sweep. Anything else -- a missing
 * complement, flags or overlaps that disagree, deleted arcs still in the array -- takes the plain path above. */
typedef struct { const asmg_t *g; uint64_t *comp; int bad; } sym_t;

static void sym_scan(uint64_t lo, uint64_t hi, void *arg)
{
    sym_t *S = (sym_t *) arg;
    const asmg_t *g = S->g;
    uint64_t i;
    int bad = 0;
    for (i = lo; i < hi; ++i) {
        const asmg_arc_t *a = &g->arc[i], *c = find_arc(g, a->w ^ 1, a->v ^ 1, 1);
        if (a->del || !c || c->comp != (a->comp ^ 1u) || c->ln != a->ln || c->ls != a->ls) { bad = 1; break; }
        S->comp[i] = (uint64_t) (c - g->arc);
    }
    if (bad) __atomic_store_n(&S->bad, 1, __ATOMIC_RELAXED);
}

static int finalize_symmetric(asmg_t *g)
{
    sym_t S = {g, 0, 0};
    const uint64_t PENDING = 0x8000000000000000ULL;
    uint64_t i, next = 0;
    static long min_arcs_cached = -1;                    /* small graphs have nothing to gain; OATK_PF_MIN (tests) overrides */
    long min_arcs = __atomic_load_n(&min_arcs_cached, __ATOMIC_RELAXED);
    if (min_arcs < 0) { const char *e = getenv("OATK_PF_MIN"); min_arcs = e ? atol(e) : 1000000; __atomic_store_n(&min_arcs_cached, min_arcs, __ATOMIC_RELAXED); }
    if (g->n_arc < (uint64_t) min_arcs || g->n_arc == 0) return 0;
    S.comp = (uint64_t *) malloc(sizeof(uint64_t) * g->n_arc);
    oatk_parallel_for(g->n_arc, sym_scan, &S);
    if (S.bad) { free(S.comp); return 0; }
    /* asmg_shrink_link_id, with the complement taken from the index (no arc is deleted, so "first arc" and "first live
     * arc" are the same) */
    for (i = 0; i < g->n_arc; ++i) g->arc[i].link_id |= PENDING;
    for (i = 0; i < g->n_arc; ++i) {
        asmg_arc_t *a = &g->arc[i];
        if (!(a->link_id & PENDING)) continue;
        a->link_id = next;
        g->arc[S.comp[i]].link_id = next;
        ++next;
    }
    free(S.comp);
    return 1;
}

static int nothing_deleted(const asmg_t *g)
{
    uint64_t i;
    for (i = 0; i < g->n_vtx; ++i) if (g->vtx[i].del) return 0;
    for (i = 0; i < g->n_arc; ++i) if (g->arc[i].del) return 0;
    return 1;
}

void asmg_finalize(asmg_t *g, int do_cleanup)
{
    if (do_cleanup && !nothing_deleted(g)) compact(g);
    asmg_arc_sort(g);
    asmg_arc_index(g);
    if (finalize_symmetric(g)) return;
    if (repair_symmetry(g) > 0) { asmg_arc_sort(g); asmg_arc_index(g); }
    asmg_shrink_link_id(g);
}

/* ---------- a9: asmg_unitigging ---------- */
typedef struct { uint64_t n, m, *a; } u64v_t;
static void u64v_push(u64v_t *v, uint64_t x)
{
    if (v->n == v->m) { v->m = v->m ? v->m << 1 : 4; v->a = (uint64_t *) realloc(v->a, 8 * v->m); }
    v->a[v->n++] = x;
}
typedef struct { uint64_t n, m; asmg_vtx_t *a; } vtxv_t;
static void keep_path(vtxv_t *u, u64v_t *p, int circ)
{
    asmg_vtx_t *x;
    if (p->n < 2) { free(p->a); return; }                     /* a path of one vertex is not a merge */
    if (u->n == u->m) { u->m = u->m ? u->m << 1 : 2; u->a = (asmg_vtx_t *) realloc(u->a, sizeof(asmg_vtx_t) * u->m); }
    x = &u->a[u->n++];
    memset(x, 0, sizeof(*x));
    x->n = p->n;
    x->a = (uint64_t *) realloc(p->a, 8 * p->n);
    x->circ = circ;
}

asmg_t *asmg_unitigging(asmg_t *g)
{
    const uint64_t MID = UINT64_MAX - 1;
    uint64_t i, j, k, nv = g->n_vtx, v;
    uint8_t *seen = (uint8_t *) calloc(nv + 1, 1);
    vtxv_t utg = {0, 0, 0};
    asmg_vtx_t *vtx = g->vtx;
    uint64_t *where;
    asmg_t *out;
    struct { uint64_t n, m; asmg_arc_t *a; } arcs = {0, 0, 0};

    /* pass 1: paths that leave a junction (a vertex with more than one live arc on either side), in
     * vertex order, both orientations, arcs in sorted order */
    for (i = 0; i < nv; ++i) {
        if (vtx[i].del || !(live_out(g, i << 1) > 1 || live_out(g, i << 1 | 1) > 1)) continue;
        for (k = 0; k < 2; ++k) {
            uint64_t s = i << 1 | k, na = out_n(g, s), nl = live_out(g, s);
            asmg_arc_t *a = out_arcs(g, s);
            v = s;      /* not reset per arc: with more than one live arc nl != 1, so the start is never pushed anyway */
            for (j = 0; j < na; ++j) {
                u64v_t p = {0, 0, 0};
                if (a[j].del) continue;
                if (!seen[v >> 1] && nl == 1) u64v_push(&p, v);
                v = a[j].w;
                while (!seen[v >> 1] && live_out(g, v ^ 1) == 1) {
                    u64v_push(&p, v);
                    seen[v >> 1] = 1;
                    if (live_out(g, v) != 1) break;
                    v = first_live(g, v)->w;
                }
                keep_path(&utg, &p, 0);
            }
        }
        seen[i] = 1;
    }
    /* pass 2: linear paths from vertices with a dead end */
    for (i = 0; i < nv; ++i) {
        u64v_t p = {0, 0, 0};
        if (vtx[i].del || seen[i] || (live_out(g, i << 1) > 0 && live_out(g, i << 1 | 1) > 0)) continue;
        v = live_out(g, i << 1) > 0 ? i << 1 : i << 1 | 1;
        u64v_push(&p, v);
        seen[v >> 1] = 1;
        while (live_out(g, v) == 1) {
            v = first_live(g, v)->w;
            if (seen[v >> 1]) break;
            u64v_push(&p, v);
            seen[v >> 1] = 1;
        }
        keep_path(&utg, &p, 0);
    }
    /* pass 3: what is left lies on circles */
    for (i = 0; i < nv; ++i) {
        u64v_t p = {0, 0, 0};
        if (vtx[i].del || seen[i]) continue;
        v = i << 1;
        u64v_push(&p, v);
        seen[i] = 1;
        while (live_out(g, v) > 0) {
            v = first_live(g, v)->w;
            if (seen[v >> 1]) break;
            u64v_push(&p, v);
            seen[v >> 1] = 1;
        }
        keep_path(&utg, &p, 1);
    }
    free(seen);

    /* where each old vertex went: start of unitig u (u<<1), end (u<<1|1), inside (MID) or nowhere yet */
    where = (uint64_t *) malloc(8 * (nv + 1));
    memset(where, 0xff, 8 * (nv + 1));
    for (i = 0; i < utg.n; ++i) {
        asmg_vtx_t *u = &utg.a[i];
        where[u->a[0] >> 1] = i << 1;
        where[u->a[u->n - 1] >> 1] = i << 1 | 1;
        for (j = 1; j + 1 < u->n; ++j) where[u->a[j] >> 1] = MID;
        for (j = 1; j < u->n; ++j) {                              /* arcs inside a unitig disappear */
            flag_arcs(g, u->a[j - 1], u->a[j]);
            flag_arcs(g, u->a[j] ^ 1, u->a[j - 1] ^ 1);
        }
    }
    for (i = 0; i < nv; ++i) {                                    /* singletons, in vertex order */
        asmg_vtx_t *x;
        if (where[i] != NO_ID || vtx[i].del) continue;
        where[i] = utg.n << 1;
        if (utg.n == utg.m) { utg.m = utg.m ? utg.m << 1 : 2; utg.a = (asmg_vtx_t *) realloc(utg.a, sizeof(asmg_vtx_t) * utg.m); }
        x = &utg.a[utg.n++];
        memset(x, 0, sizeof(*x));
        x->n = 1;
        x->a = (uint64_t *) malloc(8);
        x->a[0] = i << 1;
        x->circ = find_arc(g, i << 1, i << 1, 1) != 0;
    }
    /* surviving arcs, re-addressed to unitig ends */
    for (i = 0; i < g->n_arc; ++i) {
        asmg_arc_t *a = &g->arc[i], *b;
        uint64_t pv, pw;
        if (a->del) continue;
        pv = where[a->v >> 1]; pw = where[a->w >> 1];
        if (pv == MID || pw == MID) continue;
        if (arcs.n == arcs.m) { arcs.m = arcs.m ? arcs.m << 1 : 2; arcs.a = (asmg_arc_t *) realloc(arcs.a, sizeof(asmg_arc_t) * arcs.m); }
        b = &arcs.a[arcs.n++];
        *b = *a;
        b->v = utg.a[pv >> 1].n > 1 ? pv ^ 1 : pv | (a->v & 1);
        b->w = utg.a[pw >> 1].n > 1 ? pw : pw | (a->w & 1);
        b->del = 0;
    }
    free(where);
    /* expand each unitig into the syncmer lists of its members, minus the overlaps */
    for (i = 0; i < utg.n; ++i) {
        u64v_t l = {0, 0, 0};
        uint64_t *path = utg.a[i].a, m = utg.a[i].n, t;
        for (j = 0; j < m; ++j) {
            const asmg_vtx_t *x = &vtx[path[j] >> 1];
            if (j > 0) l.n -= find_arc(g, path[j - 1], path[j], 0)->ln;
            if (path[j] & 1) for (t = x->n; t-- > 0; ) u64v_push(&l, x->a[t] ^ 1);
            else for (t = 0; t < x->n; ++t) u64v_push(&l, x->a[t]);
        }
        free(path);
        utg.a[i].n = l.n;
        utg.a[i].a = (uint64_t *) realloc(l.a, 8 * (l.n ? l.n : 1));
        utg.a[i].cov = 0;
    }
    out = (asmg_t *) calloc(1, sizeof(asmg_t));
    out->n_vtx = utg.n; out->m_vtx = utg.m; out->vtx = utg.a;
    out->n_arc = arcs.n; out->m_arc = arcs.m; out->arc = arcs.a;
    asmg_finalize(out, 1);
    return out;
}

/* ---------- syncmer -> unitig index (syncasm.c:116-190) ---------- */
static int u128_cmp(const void *a, const void *b)
{
    uint128_t x = *(const uint128_t *) a, y = *(const uint128_t *) b;
    return (x > y) - (x < y);
}

static void index_syncmers(scg_t *g)
{
    uint64_t i, j, tot = 0, n_scm = g->scm_db->n, s;
    uint128_t *u, *p, *end;
    asmg_t *a = g->utg_asmg;
    free(g->idx_u); free(g->scm_u);
    g->idx_u = 0; g->scm_u = 0;
    if (!a) return;
    for (i = 0; i < a->n_vtx; ++i) if (!a->vtx[i].del) tot += a->vtx[i].n;
    if (!tot) return;
    u = (uint128_t *) malloc(sizeof(uint128_t) * tot);
    for (i = 0, tot = 0; i < a->n_vtx; ++i) {
        if (a->vtx[i].del) continue;
        for (j = 0; j < a->vtx[i].n; ++j) u[tot++] = (uint128_t) a->vtx[i].a[j] << 78 | (uint128_t) i << 36 | j;
    }
    for (i = 1; i < tot; ++i) if (u[i - 1] >= u[i]) break;        /* singleton graphs come out in order (keys are distinct) */
    if (i < tot) qsort(u, tot, sizeof(uint128_t), u128_cmp);
    g->idx_u = (uint128_t **) malloc(sizeof(uint128_t *) * (n_scm + 1));
    p = u; end = u + tot;
    for (s = 0; s <= n_scm; ++s) {                /* idx_u[s] = first entry whose syncmer id is >= s */
        while (p < end && (uint64_t) (*p >> 79) < s) ++p;
        g->idx_u[s] = p;
    }
    g->scm_u = u;
}

typedef struct { asmg_t *g; const uint64_t *arcs4; } arc_fill_t;

/* arc records from the device's (v, w, cov, comp) quadruples */
static void arc_fill(uint64_t lo, uint64_t hi, void *arg)
{
    const arc_fill_t *F = (const arc_fill_t *) arg;
    asmg_t *g = F->g;
    uint64_t i;
    for (i = lo; i < hi; ++i) {
        asmg_arc_t *a = &g->arc[i];
        memset(a, 0, sizeof(*a));
        a->v = F->arcs4[4 * i]; a->w = F->arcs4[4 * i + 1];
        a->cov = (uint32_t) F->arcs4[4 * i + 2]; a->comp = (uint32_t) F->arcs4[4 * i + 3];
        a->link_id = UINT64_MAX;
        /* the device filter looked at coverages only; a syncmer deleted by other means (error
         * correction sets del, syncerr.c:811-812) also removes its arcs (syncasm.c:271) */
        if (g->vtx[a->v >> 1].del || g->vtx[a->w >> 1].del) a->del = 1;
    }
}

scg_t *make_syncmer_graph(sr_db_t *sr_db, syncmer_db_t *scm_db, uint32_t min_k_cov, double min_a_cov_f)
{
    scg_t *scg;
    asmg_t *g;
    uint64_t i, n, *arcs4 = 0, n_arcs = 0;
    if (scm_db->n == 0) return 0;
    n = scm_db->n;
    oatk_tick(0);
    /* one singleton unitig per syncmer; the coverage filter marks both the syncmer and its vertex (syncasm.c:223-232) */
    g = (asmg_t *) calloc(1, sizeof(asmg_t));
    g->vtx = (asmg_vtx_t *) calloc(n, sizeof(asmg_vtx_t));
    g->n_vtx = g->m_vtx = n;
    for (i = 0; i < n; ++i) {
        syncmer_t *m = &scm_db->a[i];
        m->del |= m->cov < min_k_cov;
        g->vtx[i].n = 1;
        g->vtx[i].cov = m->cov;
        g->vtx[i].del = m->del;
        /* a deleted vertex is dropped by asmg_finalize below before anybody can look at its syncmer list: with a coverage
         * threshold that is nearly all of them, and 10^6 eight-byte blocks they would be */
        if (m->del) continue;
        g->vtx[i].a = (uint64_t *) malloc(8);
        g->vtx[i].a[0] = i << 1;
    }
    oatk_tick("graph: vertices");
    /* arcs: counted, filtered and paired with their complements on the device */
    if (syncmer_graph_arcs(sr_db, scm_db, min_k_cov, min_a_cov_f, &arcs4, &n_arcs) != 0) {
        fprintf(stderr, "[E::%s] arc tally failed\n", __func__);
        asmg_destroy(g);
        free(arcs4);
        return 0;
    }
    oatk_tick("graph: arc tally (device) + download");
    g->arc = (asmg_arc_t *) malloc(sizeof(asmg_arc_t) * (n_arcs ? n_arcs : 1));
    g->n_arc = g->m_arc = n_arcs;
    { arc_fill_t F = {g, arcs4}; oatk_parallel_for(n_arcs, arc_fill, &F); }
    free(arcs4);
    oatk_tick("graph: arc records");
    asmg_finalize(g, 1);
    oatk_tick("graph: finalize");
    scg = (scg_t *) calloc(1, sizeof(scg_t));
    scg->scm_db = scm_db;
    scg->utg_asmg = g;
    index_syncmers(scg);
    oatk_tick("graph: syncmer index");
    return scg;
}

void process_mergeable_unitigs(scg_t *g)
{
    asmg_t *u = asmg_unitigging(g->utg_asmg);
    asmg_destroy(g->utg_asmg);
    g->utg_asmg = u;
    index_syncmers(g);
}

void scg_destroy(scg_t *g)
{
    if (!g) return;
    asmg_destroy(g->utg_asmg);
    free(g->scm_u); free(g->idx_u);
    free(g);
}

void scg_stat(scg_t *scg, FILE *fo, uint64_t *stats)
{
    uint64_t i, n_scm = 0, u_scm = scg->scm_db->n, n_utg = 0, n_arc = 0;
    asmg_t *a = scg->utg_asmg;
    for (i = 0; i < scg->scm_db->n; ++i) u_scm -= scg->scm_db->a[i].del;
    for (i = 0; i < a->n_arc; ++i) n_arc += !a->arc[i].del;
    for (i = 0; i < a->n_vtx; ++i) if (!a->vtx[i].del) { ++n_utg; n_scm += a->vtx[i].n; }
    if (fo) {
        fprintf(fo, "[M::%s] number unitigs  : %lu\n", __func__, (unsigned long) n_utg);
        fprintf(fo, "[M::%s] number syncmers : %lu\n", __func__, (unsigned long) n_scm);
        fprintf(fo, "[M::%s] number arcs     : %lu\n", __func__, (unsigned long) n_arc);
    }
    if (stats) { stats[0] = n_scm; stats[1] = u_scm; stats[2] = n_utg; stats[3] = n_arc; }
}
