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

/* ---------- bubbles ----------
 * What asmg_pop_bubble does (reference graph.c:494-575, 782-882), stated as rules and written from them:
 *
 * EXPLORE. From an oriented unitig `src` with two or more ways out, unitigs are taken in dependency order: one becomes
 * "free" when every live arc into it has been followed, and the most recently freed one is taken next. Following an arc
 * x -> y offers y a route: its length (bases from the end of src, overlaps removed) and its mass (bases x coverage).
 * y remembers the SHORTEST length, the LARGEST mass, and as predecessor the x of the most massive offer, the longer
 * route winning a tie on mass. The exploration ends
 *   - at a JOIN: the unitig just taken is the only free one and nothing is still held back -- every route from src has
 *     funnelled into it. src itself qualifies trivially and does not count;
 *   - at an arc back into src (either strand): a cycle through the source, nothing is resolved;
 *   - at a dead end. A dead end closer than the radius is a short tip; it is counted unless it is where everything
 *     ended anyway, and the exploration carries on past it only when tips are not protected;
 *   - when some remembered length exceeds the radius.
 * RESOLVE. If a join was found, the predecessor chain join -> src is the path that stays; every other unitig and arc
 * the exploration touched goes -- unless more than max_del unitigs would go, or (super-bubble protection) what would go
 * is covered at least half as deeply per base as the flanking unbranched stretches, or as the path that stays.
 */
typedef struct {
    uint64_t pred;       /* where the most massive route comes from */
    uint64_t shortest;   /* bases, shortest route */
    uint64_t mass;       /* bases x coverage, most massive route */
    uint64_t held;       /* live arcs into it not followed yet */
    int known;
} route_t;

typedef struct {
    route_t *of;                   /* per oriented unitig */
    vec_t free_list, touched, arcs_used;
    uint64_t n_short_tip, n_join, join, join_dist;
    int cycle_through_source;
} region_t;

static void route_clear(route_t *r) { r->pred = UINT64_MAX; r->shortest = r->mass = r->held = 0; r->known = 0; }

/* y is reached from x over a route of `len` bases and `mass`; returns 1 when that was the last arc y was waiting for */
static int offer_route(const asmg_t *g, region_t *R, uint64_t x, uint64_t y, uint64_t len, uint64_t mass, uint64_t *n_held)
{
    route_t *r = &R->of[y];
    if (!r->known) {
        vpush(&R->touched, y);
        r->known = 1; r->pred = x; r->shortest = len; r->mass = mass;
        r->held = live_out(g, y ^ 1);
        ++*n_held;
    } else {
        const int heavier = mass > r->mass, as_heavy_but_longer = mass == r->mass && len > r->shortest;
        if (heavier || as_heavy_but_longer) r->pred = x;
        if (heavier) r->mass = mass;
        if (len < r->shortest) r->shortest = len;
    }
    assert(r->held > 0 && *n_held > 0);
    if (--r->held) return 0;
    --*n_held;
    return 1;
}

static uint64_t explore_region(const asmg_t *g, uint64_t src, uint64_t radius, int pass_short_tips, region_t *R)
{
    uint64_t n_held = 0, reach = 0;
    int stop = 0;
    R->free_list.n = R->touched.n = R->arcs_used.n = 0;
    R->n_short_tip = R->n_join = R->join_dist = 0;
    R->cycle_through_source = 0;
    R->join = UINT64_MAX;
    if (g->vtx[src >> 1].del) return 0;
    route_clear(&R->of[src]);
    vpush(&R->free_list, src);
    while (!stop && R->free_list.n > 0 && reach <= radius) {
        const uint64_t x = R->free_list.a[--R->free_list.n];
        const route_t here = R->of[x];
        const int funnelled = R->free_list.n == 0 && n_held == 0;
        if (funnelled) {
            R->join = x; R->join_dist = here.shortest;
            if (x != src) { ++R->n_join; break; }         /* the first join ends the exploration (bubble chains are not followed) */
        }
        if (live_out(g, x) == 0) {                         /* dead end */
            const int is_short = here.shortest + g->vtx[x >> 1].len < radius;
            if (is_short && !funnelled) ++R->n_short_tip;
            if (is_short && pass_short_tips) continue;
            break;
        }
        const asmg_arc_t *out = arcs_of(g, x);
        const uint64_t n_out = g->idx_n[x];
        for (uint64_t i = 0; i < n_out; ++i) {
            if (out[i].del) continue;
            const uint64_t y = out[i].w;
            if (y >> 1 == src >> 1) { R->cycle_through_source |= y == src ? 1 : 2; stop = 1; break; }
            const uint64_t step = g->vtx[x >> 1].len - out[i].ls;
            vpush(&R->arcs_used, g->idx_p[x] + i);
            if (offer_route(g, R, x, y, here.shortest + step, here.mass + g->vtx[x >> 1].cov * step, &n_held)) vpush(&R->free_list, y);
            if (R->of[y].shortest > reach) reach = R->of[y].shortest;
        }
    }
    return R->n_join;
}

/* bases and bases x coverage of a set of oriented unitigs */
static void bases_and_mass(const asmg_t *g, const uint64_t *v, uint64_t n, uint64_t *bases, uint64_t *mass)
{
    *bases = *mass = 0;
    for (uint64_t i = 0; i < n; ++i) { const asmg_vtx_t *x = &g->vtx[v[i] >> 1]; *bases += x->len; *mass += x->len * x->cov; }
}

/* does the part of the region that would be deleted look like real sequence? (graph.c:800-815) */
static int region_is_well_covered(const asmg_t *g, uint64_t src, const region_t *R, const vec_t *kept)
{
    uint64_t b_kept, m_kept, b_all, m_all, left_b, right_b;
    vec_t flank = {0, 0, 0};
    bases_and_mass(g, kept->a, kept->n, &b_kept, &m_kept);
    bases_and_mass(g, R->touched.a, R->touched.n, &b_all, &m_all);
    chain_from(g, src ^ 1, (int32_t) (g->n_vtx * 2 + 1), &left_b, &flank, 0);
    const uint64_t left_m = weighted_len(g, flank.a, flank.n);
    chain_from(g, R->join, (int32_t) (g->n_vtx * 2 + 1), &right_b, &flank, 0);
    const uint64_t right_m = weighted_len(g, flank.a, flank.n);
    free(flank.a);
    const uint64_t b_gone = b_all - b_kept, m_gone = m_all - m_kept;
    /* coverage per base of what goes, against half that of the flanks and half that of what stays */
    if (m_gone * (left_b + right_b) * 2 > (left_m + right_m) * b_gone) return 1;
    if (m_gone * b_kept * 2 > m_kept * b_gone) return 1;
    return 0;
}

static int keep_heaviest_path(asmg_t *g, uint64_t src, uint64_t max_del, int protect_super_bubble, region_t *R)
{
    vec_t kept = {0, 0, 0};                                /* join, ..., first unitig after src */
    uint64_t i, x;
    int done = 0;
    assert(R->free_list.n == 0);
    for (x = R->join; x != src; x = R->of[x].pred) vpush(&kept, x);
    if (!(max_del > 0 && R->touched.n > kept.n + max_del) && !(protect_super_bubble && region_is_well_covered(g, src, R, &kept))) {
        for (i = 0; i < R->touched.n; ++i) g->vtx[R->touched.a[i] >> 1].del = 1;
        for (i = 0; i < R->arcs_used.n; ++i) {
            asmg_arc_t *e = &g->arc[R->arcs_used.a[i]];
            e->del = 1;
            flag_arcs(g, e->w ^ 1, e->v ^ 1, 1);
        }
        for (i = 0; i < kept.n; ++i) {                     /* bring the path back, arc by arc and on both strands */
            const uint64_t y = kept.a[i], px = R->of[y].pred;
            g->vtx[y >> 1].del = 0;
            flag_arcs(g, px, y, 0);
            flag_arcs(g, y ^ 1, px ^ 1, 0);
        }
        done = 1;
    }
    free(kept.a);
    return done;
}

uint64_t asmg_pop_bubble(asmg_t *g, uint64_t radius, uint64_t max_del, int protect_tip, int protect_super_bubble, int do_cleanup, int VERBOSE)
{
    const uint64_t n_or = g->n_vtx << 1;
    region_t R;
    uint64_t v, i, n_pop = 0;
    memset(&R, 0, sizeof(R));
    R.of = (route_t *) calloc(n_or ? n_or : 1, sizeof(route_t));
    for (v = 0; v < n_or; ++v) R.of[v].pred = UINT64_MAX;
    for (v = 0; v < n_or; ++v) {
        uint64_t ret = 0;
        if (g->vtx[v >> 1].del || live_out(g, v) < 2) continue;
        explore_region(g, v, g->vtx[v >> 1].len + radius, !protect_tip, &R);
        if (R.n_join && keep_heaviest_path(g, v, max_del, protect_super_bubble, &R)) ret = 1 | R.n_short_tip << 32;
        for (i = 0; i < R.touched.n; ++i) route_clear(&R.of[R.touched.a[i]]);
        n_pop += ret;
    }
    free(R.of); free(R.free_list.a); free(R.touched.a); free(R.arcs_used.a);
    if (do_cleanup && n_pop > 0) asmg_finalize(g, 1);
    if (VERBOSE)
        fprintf(stderr, "[M::%s] popped %u bubbles and trimmed %u short tips\n", __func__, (uint32_t) n_pop, (uint32_t) (n_pop >> 32));
    return n_pop;
}
