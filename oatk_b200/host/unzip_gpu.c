/*
 * unzip_gpu.c -- repeat unzipping on the unitig graph (reference syncasm.c:1090-1484 scg_multiplex,
 * 1486-1641 scg_demultiplex), the step between clean-up and the final coverage estimates
 * (run_syncasm.c:207-262).
 *
 * scg_multiplex: reads that run through a unitig v1 with anchored alignments on both sides vote for (arc in, arc out)
 * pairs. Where one pairing clearly wins (the loser has < min_d_f of the best score of BOTH its arcs) the graph is
 * rewritten as its line graph around v1: every arc that touches such a unitig becomes a vertex (the two unitigs
 * glued), every surviving (arc in, arc out) pair an arc whose overlap is the whole middle unitig. Losing pairs simply
 * get no arc, which is what separates the repeat copies. Then the usual finalize + unitig merge.
 * scg_demultiplex: back to a graph without duplicated syncmers -- one vertex per distinct syncmer of every connected
 * component, arcs for consecutive syncmers inside unitigs and for non-overlapping arcs between them -- and merge.
 *
 * Vertex and arc numbering follow the reference's discovery order (it decides unitig ids in the GFA): arcs in index
 * order, components by breadth-first search from the lowest oriented vertex with a FIFO queue, syncmers in unitig
 * order. The reference's hash tables are only ever probed and filled (never iterated), so plain open addressing
 * does here.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include "graph_gpu.h"

#define NONE UINT64_MAX

/* ---------- a small map: 128-bit key -> 64-bit value ---------- */
typedef union { uint64_t u; double d; } val_t;
typedef struct { uint128_t *key; val_t *val; uint8_t *used; size_t cap, n; } map_t;

static size_t map_slot(const map_t *m, uint128_t k)
{
    uint64_t h = (uint64_t) k ^ (uint64_t) (k >> 64) * 0x9e3779b97f4a7c15ULL;
    size_t i;
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ULL; h ^= h >> 32;
    for (i = h & (m->cap - 1); m->used[i] && m->key[i] != k; i = (i + 1) & (m->cap - 1)) {}
    return i;
}

static void map_grow(map_t *m)
{
    map_t b;
    size_t i;
    b.cap = m->cap ? m->cap << 1 : 64; b.n = m->n;
    b.key = (uint128_t *) malloc(sizeof(uint128_t) * b.cap);
    b.val = (val_t *) malloc(sizeof(val_t) * b.cap);
    b.used = (uint8_t *) calloc(b.cap, 1);
    for (i = 0; i < m->cap; ++i) if (m->used[i]) {
        const size_t j = map_slot(&b, m->key[i]);
        b.used[j] = 1; b.key[j] = m->key[i]; b.val[j] = m->val[i];
    }
    free(m->key); free(m->val); free(m->used);
    *m = b;
}

/* slot of k, inserted if new (*is_new says which) */
static val_t *map_put(map_t *m, uint128_t k, int *is_new)
{
    size_t i;
    if ((m->n + 1) * 2 > m->cap) map_grow(m);
    i = map_slot(m, k);
    *is_new = !m->used[i];
    if (!m->used[i]) { m->used[i] = 1; m->key[i] = k; m->val[i].u = 0; ++m->n; }
    return &m->val[i];
}

static val_t *map_get(const map_t *m, uint128_t k)
{
    size_t i;
    if (!m->cap) return 0;
    i = map_slot(m, k);
    return m->used[i] ? &m->val[i] : 0;
}

static void map_clear(map_t *m) { if (m->cap) memset(m->used, 0, m->cap); m->n = 0; }
static void map_free(map_t *m) { free(m->key); free(m->val); free(m->used); memset(m, 0, sizeof(*m)); }

#define PAIR(a, b) ((uint128_t) (a) << 64 | (b))

/* ---------- graph helpers ---------- */
typedef struct { size_t n, m; uint64_t *a; } vec_t;

static void vpush(vec_t *v, uint64_t x)
{
    if (v->n == v->m) { v->m = v->m ? v->m << 1 : 8; v->a = (uint64_t *) realloc(v->a, 8 * v->m); }
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

static asmg_arc_t *find(const asmg_t *g, uint64_t v, uint64_t w, int live_only)
{
    asmg_arc_t *a = arcs_of(g, v);
    uint64_t i, n = g->idx_n[v];
    for (i = 0; i < n; ++i) if (a[i].w == w && !(live_only && a[i].del)) return &a[i];
    return 0;
}

#define ARC_ID(a) ((a)->link_id << 1 | (a)->comp)
static uint64_t comp_id(const asmg_arc_t *a) { return ((a->v ^ 1) != a->w || (a->w ^ 1) != a->v) ? ARC_ID(a) ^ 1 : ARC_ID(a); }

static uint64_t new_vertex(asmg_t *g)
{
    if (g->n_vtx == g->m_vtx) {
        g->m_vtx = g->m_vtx ? g->m_vtx << 1 : 16;
        g->vtx = (asmg_vtx_t *) realloc(g->vtx, sizeof(asmg_vtx_t) * g->m_vtx);
    }
    memset(&g->vtx[g->n_vtx], 0, sizeof(asmg_vtx_t));
    return g->n_vtx++;
}

static void new_arc(asmg_t *g, uint64_t v, uint64_t w, uint64_t ln, uint64_t ls, uint64_t link_id, uint32_t cov, uint32_t comp)
{
    asmg_arc_t *a;
    if (g->n_arc == g->m_arc) {
        g->m_arc = g->m_arc ? g->m_arc << 1 : 16;
        g->arc = (asmg_arc_t *) realloc(g->arc, sizeof(asmg_arc_t) * g->m_arc);
    }
    a = &g->arc[g->n_arc++];
    memset(a, 0, sizeof(*a));
    a->v = v; a->w = w; a->ln = ln; a->ls = ls; a->link_id = link_id; a->cov = cov; a->comp = comp;
}

/* the syncmers of oriented unitig v, in walking direction, appended to out */
static void append_oriented(vec_t *out, const asmg_vtx_t *x, int rev)
{
    uint64_t i;
    if (rev) for (i = x->n; i-- > 0; ) vpush(out, x->a[i] ^ 1);
    else for (i = 0; i < x->n; ++i) vpush(out, x->a[i]);
}

/* which fragments of a record hold a syncmer that occurs once in the graph (all of them for a read with one record) */
static void anchored_fragments(const scg_t *g, const scg_ra_t *r, double weight, uint8_t *flag)
{
    uint32_t j;
    uint64_t s;
    if (weight >= .99) { memset(flag, 0xff, r->n); return; }
    for (j = 0; j < r->n; ++j) {
        const uint64_t *a = g->utg_asmg->vtx[r->a[j].uid >> 1].a;
        flag[j] = 0;
        for (s = r->a[j].u_beg; s <= r->a[j].u_end; ++s)
            if (g->idx_u[(a[s] >> 1) + 1] - g->idx_u[a[s] >> 1] == 1) { flag[j] = 1; break; }
    }
}

int scg_multiplex(scg_t *g, scg_ra_v *ra_v, uint32_t max_n_scm, double min_n_r, double min_d_f)
{
    asmg_t *ug = g->utg_asmg;
    const uint64_t n_vtx = ug->n_vtx, n_arc = ug->n_arc, n_id = asmg_max_link_id(ug) * 2 + 2;
    map_t votes = {0, 0, 0, 0, 0}, made = {0, 0, 0, 0, 0};
    uint8_t *anchored = 0, *role = (uint8_t *) calloc(n_vtx ? n_vtx : 1, 1);   /* 1: rewritten around, 2: isolated */
    uint64_t *glued = (uint64_t *) malloc(8 * n_id);                         /* per arc id: the vertex that replaces it */
    vec_t *follow = (vec_t *) calloc(n_id, sizeof(vec_t));                   /* per arc id: where the walk may go next */
    size_t m_anchored = 0;
    uint64_t i, j, s, t;
    int updated = 0, is_new;
    double whole;

    /* votes of the reads: (arc in, arc out) around every middle fragment, both strands */
    for (i = 0; i < ra_v->n; ++i) {
        const scg_ra_t *r = &ra_v->a[i];
        const asmg_arc_t *arc;
        uint64_t l0, c0, l1, c1;
        double weight;
        val_t *v;
        if (r->n < 3) continue;
        weight = modf(r->s, &whole);
        if (weight < DBL_EPSILON) weight = 1.0;
        if (r->n > m_anchored) { m_anchored = r->n; anchored = (uint8_t *) realloc(anchored, m_anchored); }
        anchored_fragments(g, r, weight, anchored);
        arc = find(ug, r->a[0].uid, r->a[1].uid, 0);
        l0 = ARC_ID(arc); c0 = comp_id(arc);
        for (j = 2; j < r->n; ++j, l0 = l1, c0 = c1) {
            arc = find(ug, r->a[j - 1].uid, r->a[j].uid, 0);
            l1 = ARC_ID(arc); c1 = comp_id(arc);
            if (!anchored[j - 2] || !anchored[j - 1] || !anchored[j]) continue;
            v = map_put(&votes, PAIR(l0, l1), &is_new);
            if (is_new) {
                v->d = weight;
                map_put(&votes, PAIR(c1, c0), &is_new)->d = weight;
            } else {
                v->d += weight;
                map_put(&votes, PAIR(c1, c0), &is_new)->d += weight;
            }
        }
    }
    free(anchored);
    for (i = 0; i < n_id; ++i) glued[i] = NONE;

    for (i = 0; i < n_vtx; ++i) {
        const uint64_t v1 = i << 1;
        uint64_t n_in, n_out;
        asmg_arc_t **in, **out, *a;
        double *score, *best_in, *best_out, top = .0;
        if (ug->vtx[i].del) continue;
        n_in = live_out(ug, v1 ^ 1); n_out = live_out(ug, v1);
        if (n_in == 0 && n_out == 0) { role[i] = 2; continue; }
        if (n_in == 0 || n_out == 0) continue;
        in = (asmg_arc_t **) malloc(sizeof(asmg_arc_t *) * (n_in + n_out)); out = in + n_in;
        score = (double *) calloc(n_in * n_out + n_in + n_out, sizeof(double)); best_in = score + n_in * n_out; best_out = best_in + n_in;
        for (s = 0, j = 0, a = arcs_of(ug, v1 ^ 1); s < ug->idx_n[v1 ^ 1]; ++s) if (!a[s].del) in[j++] = &a[s];
        for (t = 0, j = 0, a = arcs_of(ug, v1); t < ug->idx_n[v1]; ++t) if (!a[t].del) out[j++] = &a[t];
        for (s = 0; s < n_in; ++s)
            for (t = 0; t < n_out; ++t) {
                const val_t *v = map_get(&votes, PAIR(comp_id(in[s]), ARC_ID(out[t])));
                const double sc = v ? v->d : .001;
                score[s * n_out + t] = sc;
                if (sc > best_in[s]) best_in[s] = sc;
                if (sc > best_out[t]) best_out[t] = sc;
                if (sc > top) top = sc;
            }
        /* long unitigs (reads spanning them are rare), self loops and thinly supported ones keep every pairing */
        role[i] = !(ug->vtx[i].n > max_n_scm || find(ug, v1, v1, 1) || top < min_n_r);
        for (s = 0; s < n_in; ++s)
            for (t = 0; t < n_out; ++t) {
                const double sc = score[s * n_out + t];
                if (role[i] && sc / best_in[s] < min_d_f && sc / best_out[t] < min_d_f) { ++updated; continue; }
                vpush(&follow[comp_id(in[s])], out[t]->w);
                vpush(&follow[ARC_ID(out[t]) ^ 1], in[s]->w);
            }
        free(in); free(score);
    }
    map_free(&votes);

    if (updated) {
        /* every arc next to a rewritten unitig becomes a vertex: its two unitigs glued over their overlap */
        for (i = 0; i < n_arc; ++i) {
            const asmg_arc_t *arc = &ug->arc[i];
            vec_t list = {0, 0, 0};
            uint64_t id, nv;
            if (arc->del || arc->comp || (role[arc->v >> 1] != 1 && role[arc->w >> 1] != 1)) continue;
            id = ARC_ID(arc);
            nv = new_vertex(ug);
            glued[id] = nv << 1; glued[id ^ 1] = nv << 1 | 1;
            append_oriented(&list, &ug->vtx[arc->v >> 1], (int) (arc->v & 1));
            list.n -= arc->ln;
            append_oriented(&list, &ug->vtx[arc->w >> 1], (int) (arc->w & 1));
            ug->vtx[nv].n = list.n;
            ug->vtx[nv].a = (uint64_t *) realloc(list.a, 8 * (list.n ? list.n : 1));
        }
        /* arcs between the new vertices (and from / to the old ones at the rim), one per allowed pairing */
        for (i = 0; i < n_arc; ++i) {
            uint64_t mid, id, from, c0;
            if (ug->arc[i].del) continue;
            mid = ug->arc[i].w; id = ARC_ID(&ug->arc[i]); c0 = ug->arc[i].cov;
            from = glued[id] == NONE ? mid : glued[id];
            for (j = 0; j < follow[id].n; ++j) {
                const asmg_arc_t *next = find(ug, mid, follow[id].a[j], 0);
                const uint64_t c1 = next->cov, to = glued[ARC_ID(next)] == NONE ? mid : glued[ARC_ID(next)];
                if (glued[id] == NONE && glued[ARC_ID(next)] == NONE) continue;
                map_put(&made, PAIR(from, to), &is_new);
                if (!is_new) continue;
                new_arc(ug, from, to, ug->vtx[mid >> 1].n, ug->vtx[mid >> 1].len, NONE, (uint32_t) ((c0 + c1) >> 1), 0);
            }
        }
        map_free(&made);
        for (i = 0; i < n_arc; ++i) if (!ug->arc[i].del && glued[ARC_ID(&ug->arc[i])] != NONE) ug->arc[i].del = 1;
        for (i = 0; i < n_vtx; ++i)
            if (!ug->vtx[i].del && role[i] != 2 && live_out(ug, i << 1 | 1) == 0 && live_out(ug, i << 1) == 0) ug->vtx[i].del = 1;
        asmg_finalize(ug, 1);
        process_mergeable_unitigs(g);
    }
    for (i = 0; i < n_id; ++i) free(follow[i].a);
    free(follow); free(glued); free(role);
    return updated;
}

void scg_demultiplex(scg_t *g)
{
    asmg_t *ug = g->utg_asmg, *dg = (asmg_t *) calloc(1, sizeof(asmg_t));
    const uint64_t n_or = ug->n_vtx * 2;
    uint8_t *done = (uint8_t *) calloc(n_or ? n_or : 1, 1);
    vec_t queue = {0, 0, 0}, members = {0, 0, 0};
    map_t vertex_of = {0, 0, 0, 0, 0}, made = {0, 0, 0, 0, 0};
    uint64_t i, j, k, head;
    int is_new;

    for (i = 0; i < n_or; ++i) {
        if (done[i] || ug->vtx[i >> 1].del) continue;
        /* the component of i, both strands, first in first out; a unitig joins when its reverse strand comes up */
        members.n = queue.n = 0; head = 0;
        vpush(&queue, i); vpush(&queue, i ^ 1);
        while (head < queue.n) {
            const uint64_t v = queue.a[head++];
            const asmg_arc_t *a = arcs_of(ug, v);
            if (done[v]) continue;
            if (v & 1) vpush(&members, v >> 1);
            for (j = 0; j < ug->idx_n[v]; ++j) {
                if (a[j].del) continue;
                if (!done[a[j].w]) vpush(&queue, a[j].w);
                if (!done[a[j].w ^ 1]) vpush(&queue, a[j].w ^ 1);
            }
            done[v] = 1;
        }
        /* one vertex per distinct syncmer, arcs along the unitigs */
        for (j = 0; j < members.n; ++j) {
            const asmg_vtx_t *x = &ug->vtx[members.a[j]];
            uint64_t prev = 0, cur = 0;
            for (k = 0; k < x->n; ++k, prev = cur) {
                val_t *slot = map_put(&vertex_of, x->a[k] >> 1, &is_new);
                if (is_new) {
                    cur = new_vertex(dg);
                    dg->vtx[cur].a = (uint64_t *) malloc(8);
                    dg->vtx[cur].n = 1;
                    dg->vtx[cur].a[0] = x->a[k] >> 1 << 1;
                    slot->u = cur;
                } else cur = slot->u;
                if (k > 0) {
                    const uint64_t v = prev << 1 | (x->a[k - 1] & 1), w = cur << 1 | (x->a[k] & 1);
                    if (map_get(&made, PAIR(v, w))) continue;
                    new_arc(dg, v, w, 0, 0, 0, 0, 0);
                    if (v != (w ^ 1)) new_arc(dg, w ^ 1, v ^ 1, 0, 0, 0, 0, 1);
                    map_put(&made, PAIR(v, w), &is_new);
                    map_put(&made, PAIR(w ^ 1, v ^ 1), &is_new);
                }
            }
        }
        /* arcs between unitigs that do not overlap: last syncmer of one -> first of the other. A component is closed
         * under live arcs, so walking the arcs of its members finds the same pairs as trying every member against
         * every other (syncasm.c:1596-1621); of several arcs v -> w only the first live one counts there, and here */
        for (j = 0; j < members.n * 2; ++j) {
            const uint64_t v = members.a[j >> 1] << 1 | (j & 1);
            const asmg_vtx_t *x = &ug->vtx[v >> 1];
            const asmg_arc_t *a = arcs_of(ug, v);
            uint64_t from = (v & 1) ? x->a[0] ^ 1 : x->a[x->n - 1], last_w = NONE;
            from = map_get(&vertex_of, from >> 1)->u << 1 | (from & 1);
            for (k = 0; k < ug->idx_n[v]; ++k) {
                const asmg_vtx_t *y;
                uint64_t to;
                if (a[k].del || a[k].w == last_w) continue;
                last_w = a[k].w;
                if (a[k].ln > 0) continue;
                y = &ug->vtx[a[k].w >> 1];
                to = (a[k].w & 1) ? y->a[y->n - 1] ^ 1 : y->a[0];
                to = map_get(&vertex_of, to >> 1)->u << 1 | (to & 1);
                if (map_get(&made, PAIR(from, to))) continue;
                new_arc(dg, from, to, 0, 0, 0, 0, 0);
                map_put(&made, PAIR(from, to), &is_new);
            }
        }
        map_clear(&vertex_of);
        map_clear(&made);
    }
    free(done); free(queue.a); free(members.a);
    map_free(&vertex_of); map_free(&made);
    asmg_finalize(dg, 1);
    asmg_destroy(ug);
    g->utg_asmg = dg;
    process_mergeable_unitigs(g);
}
