/*
 * cleaning_gpu.c -- the three clean-up passes syncasm runs on the unitig graph between `.utg.gfa` and
 * `.utg.final.gfa` (reference run_syncasm.c:178-192, 273-282):
 *   asmg_drop_tip               graph.c:607-679   dead-end chains no longer than tip_len, unless they carry more than
 *                                                 half the coverage of what they compete with
 *   asmg_remove_weak_crosslink  graph.c:698-776   arcs dominated on both ends by a sibling with >= m_cov reads
 *   asmg_pop_bubble             graph.c:782-886   (with asmg_topo_ext :494-579, asmg_uext :347-380, asmg_cwt_len :594-604)
 *                                                 bubbles within `radius`: keep the heaviest path, drop the rest
 * All three work on vertex lengths / arc overlaps in bases, i.e. after scg_consensus filled them in.
 *
 * These passes are order-sensitive by design (vertices in id order, deletions of the tip pass deferred to its end,
 * bubbles resolved immediately) and so small (a few 10^3 unitigs) that there is nothing to parallelise; they are
 * restated here so that the whole `syncasm` command can run on this layer. Two reference habits are visible in the
 * result and kept: the comparison at graph.c:654 binds as `(del || (w ^ 1)) == tip_end`, which skips next to
 * nothing, and coverage-weighted lengths use the FIRST arc between two vertices whether or not it is deleted.
 */
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include "graph_gpu.h"

typedef struct { size_t n, m; uint64_t *a; } vec_t;

static void vpush(vec_t *v, uint64_t x)
{
    if (v->n == v->m) { v->m = v->m ? v->m << 1 : 16; v->a = (uint64_t *) realloc(v->a, 8 * v->m); }
    v->a[v->n++] = x;
}

static inline asmg_arc_t *arcs_of(const asmg_t *g, uint64_t v) { return &g->arc[g->idx_p[v]]; }

static uint64_t live_out(const asmg_t *g, uint64_t v)
{
    const asmg_arc_t *a = arcs_of(g, v);
    uint64_t i, n = g->idx_n[v], c = 0;
    for (i = 0; i < n; ++i) c += !a[i].del;
    return c;
}

static asmg_arc_t *first_arc(const asmg_t *g, uint64_t v, uint64_t w)
{
    asmg_arc_t *a = arcs_of(g, v);
    uint64_t i, n = g->idx_n[v];
    for (i = 0; i < n; ++i) if (a[i].w == w) return &a[i];
    return 0;
}

static void flag_arcs(asmg_t *g, uint64_t v, uint64_t w, uint32_t del)
{
    asmg_arc_t *a = arcs_of(g, v);
    uint64_t i, n = g->idx_n[v];
    for (i = 0; i < n; ++i) if (a[i].w == w) a[i].del = del;
}

static void flag_vertex(asmg_t *g, uint64_t s, uint32_t del)
{
    int o;
    g->vtx[s].del = del;
    for (o = 0; o < 2; ++o) {
        const uint64_t v = s << 1 | o;
        asmg_arc_t *a = arcs_of(g, v);
        uint64_t i, n = g->idx_n[v];
        for (i = 0; i < n; ++i) { a[i].del = del; flag_arcs(g, a[i].w ^ 1, v ^ 1, del); }
    }
}

/* where a walk from v stands after its vertex */
enum { END_TIP, END_CHAIN, END_SHARED, END_FORK };   /* no way out / unique way into an unshared vertex / into a shared one / several ways out */

/* follow the unbranched chain that starts with v for at most max_steps vertices; `len` = bases the chain adds
 * (each vertex minus its largest live overlap); with tip_only a chain that runs into a fork gives the fork back */
static int chain_from(const asmg_t *g, uint64_t v, int32_t max_steps, uint64_t *len, vec_t *path, int tip_only)
{
    uint64_t total = 0, step = 0;
    int kind;
    path->n = 0;
    vpush(path, v);
    do {
        const asmg_arc_t *a = arcs_of(g, v);
        uint64_t i, n = g->idx_n[v], live = 0, last = 0, ovl = 0, next = UINT64_MAX;
        step = 0;
        if (!g->vtx[v >> 1].del) {
            for (i = 0; i < n; ++i) if (!a[i].del) { ++live; last = i; if (a[i].ls > ovl) ovl = a[i].ls; }
            step = g->vtx[v >> 1].len - ovl;
            if (live == 1) next = a[last].w;
        }
        kind = live == 0 ? END_TIP : live > 1 ? END_FORK : live_out(g, next ^ 1) == 1 ? END_CHAIN : END_SHARED;
        total += step;
        if (kind != END_CHAIN) break;
        vpush(path, next);
        v = next;
    } while (--max_steps > 0);
    if (tip_only && kind == END_FORK) { total -= step; --path->n; }
    *len = total;
    return kind;
}

/* bases times coverage along a path */
static uint64_t weighted_len(const asmg_t *g, const uint64_t *v, size_t n)
{
    uint64_t wt;
    size_t i;
    if (n == 0) return 0;
    wt = g->vtx[v[0] >> 1].len * g->vtx[v[0] >> 1].cov;
    for (i = 1; i < n; ++i) wt += (g->vtx[v[i] >> 1].len - first_arc(g, v[i - 1], v[i])->ls) * g->vtx[v[i] >> 1].cov;
    return wt;
}

/* arcs grouped by the unbranched chain they lie on (graph.c:382-437): per link id a group number, chains first in
 * vertex order, then every remaining live arc a group of its own; caller frees */
uint32_t *asmg_uext_arc_group(asmg_t *g, uint32_t *n_group)
{
    const uint64_t n_link = asmg_max_link_id(g) + 1;
    uint32_t *group_of = (uint32_t *) malloc(sizeof(uint32_t) * n_link), group = 0;
    uint8_t *visited = (uint8_t *) calloc(g->n_vtx ? g->n_vtx : 1, 1);
    vec_t path = {0, 0, 0};
    uint64_t i, j, len;
    int o;
    memset(group_of, 0xff, sizeof(uint32_t) * n_link);
    for (i = 0; i < g->n_vtx; ++i) {
        uint32_t na = 0;
        if (visited[i] || g->vtx[i].del) continue;
        for (o = 0; o < 2; ++o) {
            const int kind = chain_from(g, i << 1 | (uint64_t) o, (int32_t) (g->n_vtx * 2 + 1), &len, &path, 0);
            for (j = 1; j < path.n; ++j) {
                const asmg_arc_t *a = arcs_of(g, path.a[j - 1]);
                uint64_t k = 0;
                while (a[k].w != path.a[j] || a[k].del) ++k;
                group_of[a[k].link_id] = group;
                visited[path.a[j] >> 1] = 1;
                ++na;
            }
            if (kind == END_SHARED) {
                const asmg_arc_t *a = arcs_of(g, path.a[path.n - 1]);
                uint64_t k = 0;
                while (a[k].del) ++k;
                group_of[a[k].link_id] = group;
                ++na;
            }
        }
        if (na > 0) ++group;
        visited[i] = 1;
    }
    for (i = 0; i < g->n_arc; ++i) if (!g->arc[i].del && group_of[g->arc[i].link_id] == UINT32_MAX) group_of[g->arc[i].link_id] = group++;
    if (n_group) *n_group = group;
    free(path.a); free(visited);
    return group_of;
}

uint64_t asmg_drop_tip(asmg_t *g, int32_t tip_cnt, uint64_t tip_len, int protect_super_tip, int do_cleanup, int VERBOSE)
{
    const uint64_t n_or = g->n_vtx << 1;
    vec_t tip = {0, 0, 0}, rival = {0, 0, 0}, doomed = {0, 0, 0};
    uint64_t v, i, len, cnt = 0;
    if ((uint64_t) tip_cnt > n_or) tip_cnt = (int32_t) n_or;
    for (v = 0; v < n_or; ++v) {
        int kind;
        if (g->vtx[v >> 1].del || live_out(g, v ^ 1) != 0) continue;           /* nothing may lead into a tip */
        kind = chain_from(g, v, tip_cnt, &len, &tip, 1);
        if (tip.n == 0 || kind == END_CHAIN || len > tip_len) continue;
        if (kind != END_TIP && protect_super_tip) {
            /* compare with every other way into the vertex the tip joins */
            const uint64_t end = tip.a[tip.n - 1], tip_bases = len, tip_wt = weighted_len(g, tip.a, tip.n);
            const asmg_arc_t *out = arcs_of(g, end), *in;
            uint64_t join, n_in, k = 0;
            int weaker = 0;
            while (out[k].del) ++k;
            join = out[k].w ^ 1;
            in = arcs_of(g, join);
            n_in = g->idx_n[join];
            for (i = 0; i < n_in; ++i) {
                if ((uint64_t) (in[i].del || (in[i].w ^ 1)) == end) continue;    /* graph.c:654, as it binds */
                chain_from(g, in[i].w, (int32_t) (n_or + 1), &len, &rival, 0);
                if (tip_bases <= len || tip_wt * 2 <= weighted_len(g, rival.a, rival.n)) { weaker = 1; break; }
            }
            if (!weaker) continue;
        }
        for (i = 0; i < tip.n; ++i) vpush(&doomed, tip.a[i]);
        ++cnt;
    }
    for (i = 0; i < doomed.n; ++i) flag_vertex(g, doomed.a[i] >> 1, 1);
    free(tip.a); free(rival.a); free(doomed.a);
    if (do_cleanup && cnt > 0) asmg_finalize(g, 1);
    if (VERBOSE) fprintf(stderr, "[M::%s] dropped %lu tips\n", __func__, (unsigned long) cnt);
    return cnt;
}

/* is arc `a` dominated among the live arcs leaving v: some sibling with >= m_cov reads has more than 1/c_thresh times its coverage */
static int dominated(const asmg_t *g, const asmg_arc_t *a, uint64_t v, double c_thresh, double m_cov)
{
    const asmg_arc_t *s = arcs_of(g, v);
    uint64_t k, n = g->idx_n[v];
    for (k = 0; k < n; ++k) {
        if (s[k].del || s[k].cov < m_cov) continue;
        if ((double) a->cov / s[k].cov < c_thresh) return 1;
    }
    return 0;
}

uint64_t asmg_remove_weak_crosslink(asmg_t *g, double c_thresh, double m_cov, int do_cleanup, int VERBOSE)
{
    vec_t weak = {0, 0, 0};
    uint64_t i, cnt;
    for (i = 0; i < g->n_arc; ++i) {
        const asmg_arc_t *a = &g->arc[i];
        if (a->del || a->comp) continue;
        if (dominated(g, a, a->v, c_thresh, m_cov) && dominated(g, a, a->w ^ 1, c_thresh, m_cov)) vpush(&weak, i);
    }
    for (i = 0; i < weak.n; ++i) {
        asmg_arc_t *a = &g->arc[weak.a[i]];
        a->del = 1;
        flag_arcs(g, a->w ^ 1, a->v ^ 1, 1);
    }
    cnt = weak.n;
    free(weak.a);
    if (do_cleanup && cnt > 0) asmg_finalize(g, 1);
    if (VERBOSE) fprintf(stderr, "[M::%s] dropped %lu weak cross links\n", __func__, (unsigned long) cnt);
    return cnt;
}

/* ---------- bubbles ---------- */
typedef struct {
    uint64_t from;       /* predecessor on the heaviest path */
    uint64_t dist;       /* fewest bases from the source's end */
    uint64_t weight;     /* most bases x coverage from the source */
    uint64_t waiting;    /* incoming arcs not yet walked */
    int seen;
} visit_t;

typedef struct {
    visit_t *at;
    vec_t ready, seen, arcs;
    uint64_t n_short_tip, n_sink, dist, sink;
    int self_cycle;
} walk_t;

#define THRU_SHORT_TIP 1
#define THRU_BUBBLE 2

/* topological walk from v0 while everything seen stays within max_dist; a vertex is expanded once all its incoming arcs
 * have been walked; when exactly one vertex is ready and none is waiting, everything has funnelled into it: a sink */
static uint64_t funnel_walk(const asmg_t *g, uint64_t v0, uint64_t max_dist, int thru, walk_t *b)
{
    uint64_t pending = 0, far = 0;
    visit_t *t;
    if (g->vtx[v0 >> 1].del) return 0;
    b->ready.n = b->seen.n = b->arcs.n = 0;
    b->n_short_tip = b->n_sink = b->dist = 0;
    b->self_cycle = 0;
    b->sink = UINT64_MAX;
    t = &b->at[v0];
    t->dist = t->weight = t->waiting = 0; t->seen = 0; t->from = UINT64_MAX;
    vpush(&b->ready, v0);
    while (b->ready.n > 0 && far <= max_dist) {
        const uint64_t v = b->ready.a[--b->ready.n], n = g->idx_n[v], d = b->at[v].dist, c = b->at[v].weight;
        const asmg_arc_t *a = arcs_of(g, v);
        uint64_t i;
        if (b->ready.n == 0 && pending == 0) {
            b->dist = d; b->sink = v;
            if (v != v0) { ++b->n_sink; if (!(thru & THRU_BUBBLE)) break; }
        }
        if (live_out(g, v) == 0) {
            if (d + g->vtx[v >> 1].len < max_dist) {
                if (b->ready.n || pending) ++b->n_short_tip;              /* not counted when it ends a bubble chain */
                if (thru & THRU_SHORT_TIP) continue;
            }
            break;
        }
        for (i = 0; i < n; ++i) {
            uint64_t w, step, gain;
            if (a[i].del) continue;
            w = a[i].w;
            step = g->vtx[v >> 1].len - a[i].ls;
            gain = g->vtx[v >> 1].cov * step;
            t = &b->at[w];
            if (w >> 1 == v0 >> 1) { b->self_cycle |= w == v0 ? 1 : 2; break; }
            vpush(&b->arcs, g->idx_p[v] + i);
            if (!t->seen) {
                vpush(&b->seen, w);
                t->from = v; t->seen = 1; t->dist = d + step; t->weight = c + gain;
                t->waiting = live_out(g, w ^ 1);
                ++pending;
            } else {
                if (c + gain > t->weight || (c + gain == t->weight && d + step > t->dist)) t->from = v;
                if (c + gain > t->weight) t->weight = c + gain;
                if (d + step < t->dist) t->dist = d + step;
            }
            if (t->dist > far) far = t->dist;
            assert(t->waiting > 0 && pending > 0);
            if (--t->waiting == 0) { vpush(&b->ready, w); --pending; }
        }
        if (i < n) break;
    }
    return b->n_sink;
}

/* keep the heaviest source->sink path of the walked region, delete the rest (unless it looks like real sequence) */
static int resolve(asmg_t *g, uint64_t v0, uint64_t max_del, int protect_super_bubble, walk_t *b)
{
    uint64_t i, v, w;
    assert(b->ready.n == 0);
    if (max_del > 0) {
        uint64_t kept = 0;
        v = b->sink;
        do { ++kept; v = b->at[v].from; } while (v != v0);
        if (b->seen.n > kept + max_del) return 0;
    }
    if (protect_super_bubble) {
        uint64_t b_kept = 0, c_kept = 0, b_tot = 0, c_tot = 0, left, right, left_wt, right_wt;
        vec_t p = {0, 0, 0};
        v = b->sink;
        do { b_kept += g->vtx[v >> 1].len; c_kept += g->vtx[v >> 1].len * g->vtx[v >> 1].cov; v = b->at[v].from; } while (v != v0);
        for (i = 0; i < b->seen.n; ++i) {
            const asmg_vtx_t *x = &g->vtx[b->seen.a[i] >> 1];
            b_tot += x->len; c_tot += x->len * x->cov;
        }
        chain_from(g, v0 ^ 1, (int32_t) (g->n_vtx * 2 + 1), &left, &p, 0);
        left_wt = weighted_len(g, p.a, p.n);
        chain_from(g, b->sink, (int32_t) (g->n_vtx * 2 + 1), &right, &p, 0);
        right_wt = weighted_len(g, p.a, p.n);
        free(p.a);
        /* what would go is covered more than half as deeply as the flanks, or as the path that stays: leave it */
        if ((c_tot - c_kept) * (left + right) * 2 > (left_wt + right_wt) * (b_tot - b_kept)) return 0;
        if ((c_tot - c_kept) * b_kept * 2 > c_kept * (b_tot - b_kept)) return 0;
    }
    for (i = 0; i < b->seen.n; ++i) g->vtx[b->seen.a[i] >> 1].del = 1;
    for (i = 0; i < b->arcs.n; ++i) {
        asmg_arc_t *a = &g->arc[b->arcs.a[i]];
        a->del = 1;
        flag_arcs(g, a->w ^ 1, a->v ^ 1, 1);
    }
    v = b->sink;
    do {
        w = b->at[v].from;
        g->vtx[v >> 1].del = 0;
        flag_arcs(g, w, v, 0);
        flag_arcs(g, v ^ 1, w ^ 1, 0);
        v = w;
    } while (v != v0);
    return 1;
}

uint64_t asmg_pop_bubble(asmg_t *g, uint64_t radius, uint64_t max_del, int protect_tip, int protect_super_bubble, int do_cleanup, int VERBOSE)
{
    const uint64_t n_or = g->n_vtx << 1;
    walk_t b;
    uint64_t v, i, n_pop = 0;
    memset(&b, 0, sizeof(b));
    b.at = (visit_t *) calloc(n_or ? n_or : 1, sizeof(visit_t));
    for (v = 0; v < n_or; ++v) b.at[v].from = UINT64_MAX;
    for (v = 0; v < n_or; ++v) {
        uint64_t ret = 0;
        if (g->vtx[v >> 1].del || live_out(g, v) < 2) continue;
        funnel_walk(g, v, g->vtx[v >> 1].len + radius, protect_tip ? 0 : THRU_SHORT_TIP, &b);
        if (b.n_sink && (ret = (uint64_t) resolve(g, v, max_del, protect_super_bubble, &b))) ret |= b.n_short_tip << 32;
        for (i = 0; i < b.seen.n; ++i) {
            visit_t *t = &b.at[b.seen.a[i]];
            t->dist = t->weight = t->waiting = 0; t->seen = 0; t->from = UINT64_MAX;
        }
        n_pop += ret;
    }
    free(b.at); free(b.ready.a); free(b.seen.a); free(b.arcs.a);
    if (do_cleanup && n_pop > 0) asmg_finalize(g, 1);
    if (VERBOSE)
        fprintf(stderr, "[M::%s] popped %u bubbles and trimmed %u short tips\n", __func__, (uint32_t) n_pop, (uint32_t) (n_pop >> 32));
    return n_pop;
}
