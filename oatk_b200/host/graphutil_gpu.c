/*
 * graphutil_gpu.c -- the remaining queries of the reference's graph.h, for callers beyond syncasm() (the path finder
 * works on the graph syncasm() hands over in scg_meta_t): asmg_arc_is_sorted graph.c:60, asmg_vtx_list :267,
 * asmg_print :281, asmg_subgraph :1111, asmg_path_exists :1230, asmg_tarjans_scc :1325.
 * (asmg_uext_arc_group :382 sits in cleaning_gpu.c next to the chain walk it uses.)
 *
 * Searches are breadth first with a first-in-first-out frontier and "first pop wins", as in the reference: with a
 * step or distance limit the set that is reached depends on that order. Strongly connected components are numbered
 * in the order Tarjan's algorithm completes them on a depth-first search that takes vertices and arcs in index order;
 * the search here keeps its own stack instead of recursing (a chain of 10^6 unitigs would overflow the C stack), which
 * changes nothing in the numbering.
 */
#include <stdlib.h>
#include <string.h>
#include "graph_gpu.h"

int asmg_arc_is_sorted(asmg_t *g)
{
    uint64_t e;
    for (e = 1; e < g->n_arc; ++e) {
        const asmg_arc_t *a = &g->arc[e - 1], *b = &g->arc[e];
        if (a->v > b->v || (a->v == b->v && a->w > b->w)) return 0;
    }
    return 1;
}

uint64_t *asmg_vtx_list(asmg_t *g, uint64_t *_n)
{
    uint64_t i, n = 0, *list;
    for (i = 0; i < g->n_vtx; ++i) n += !g->vtx[i].del;
    list = (uint64_t *) malloc(sizeof(uint64_t) * (n ? n : 1));
    for (i = 0, n = 0; i < g->n_vtx; ++i) if (!g->vtx[i].del) list[n++] = i;
    if (_n) *_n = n;
    return list;
}

void asmg_print(asmg_t *g, FILE *fo, int no_seq)
{
    scg_t view;
    memset(&view, 0, sizeof(view));
    view.utg_asmg = g;
    scg_print(&view, fo, no_seq);                      /* same text (syncasm.c:825 is a copy of graph.c:281) */
}

typedef struct { uint64_t v, dist; uint32_t step; } hop_t;
typedef struct { hop_t *a; size_t head, n, m; } fifo_t;

static void fifo_push(fifo_t *q, uint64_t v, uint32_t step, uint64_t dist)
{
    if (q->n == q->m) { q->m = q->m ? q->m << 1 : 64; q->a = (hop_t *) realloc(q->a, sizeof(hop_t) * q->m); }
    q->a[q->n].v = v; q->a[q->n].step = step; q->a[q->n].dist = dist;
    ++q->n;
}

uint32_t *asmg_subgraph(asmg_t *g, uint32_t *seeds, uint32_t n, uint32_t step, uint64_t dist, uint32_t *_nv, int modify_graph)
{
    const uint64_t n_or = g->n_vtx << 1;
    int8_t *state;                                     /* per oriented vertex: 0 unseen, 1 reached, -1 deleted */
    fifo_t q = {0, 0, 0, 0};
    uint32_t *list = 0;
    uint64_t i, nv = 0;
    if (n == 0) { if (_nv) *_nv = 0; return 0; }
    if (step == 0) step = UINT32_MAX;
    if (dist == 0) dist = UINT64_MAX;
    state = (int8_t *) calloc(n_or ? n_or : 1, 1);
    for (i = 0; i < g->n_vtx; ++i) if (g->vtx[i].del) state[i << 1] = state[i << 1 | 1] = -1;
    for (i = 0; i < n; ++i) if (seeds[i] < g->n_vtx) { fifo_push(&q, (uint64_t) seeds[i] << 1, 0, 0); fifo_push(&q, (uint64_t) seeds[i] << 1 | 1, 0, 0); }
    if (modify_graph) for (i = 0; i < g->n_vtx; ++i) g->vtx[i].del = 1;
    while (q.head < q.n) {
        const hop_t h = q.a[q.head++];
        const asmg_arc_t *a = &g->arc[g->idx_p[h.v]];
        if (state[h.v] != 0) continue;
        state[h.v] = 1;
        if (modify_graph) g->vtx[h.v >> 1].del = 0;
        if (!(h.step < step && h.dist < dist)) continue;
        for (i = 0; i < g->idx_n[h.v]; ++i) {
            uint64_t d;
            if (a[i].del) continue;
            d = h.dist + g->vtx[a[i].w >> 1].len - a[i].ls;
            if (state[a[i].w] == 0) fifo_push(&q, a[i].w, h.step + 1, d);
            if (state[a[i].w ^ 1] == 0) fifo_push(&q, a[i].w ^ 1, h.step + 1, d);
        }
    }
    for (i = 0; i < g->n_vtx; ++i) state[i] = state[i << 1] > 0 || state[i << 1 | 1] > 0;      /* per unitig now (i <= 2i: in place) */
    if (!modify_graph) {
        for (i = 0; i < g->n_vtx; ++i) nv += state[i];
        list = (uint32_t *) malloc(sizeof(uint32_t) * (nv ? nv : 1));
        for (i = 0, nv = 0; i < g->n_vtx; ++i) if (state[i]) list[nv++] = (uint32_t) i;
    } else {
        for (i = 0; i < g->n_arc; ++i) if (!state[g->arc[i].v >> 1] || !state[g->arc[i].w >> 1]) g->arc[i].del = 1;
        if (_nv) for (i = 0; i < g->n_vtx; ++i) nv += state[i];
    }
    if (_nv) *_nv = (uint32_t) nv;
    free(q.a); free(state);
    return list;
}

/* is there a walk from oriented vertex `source` to `sink` within the limits; deleted flags are not looked at (graph.c:1232) */
int asmg_path_exists(asmg_t *g, uint32_t source, uint32_t sink, uint32_t step, uint64_t dist, uint32_t *_step, uint64_t *_dist)
{
    const uint64_t n_or = g->n_vtx << 1;
    uint8_t *seen;
    fifo_t q = {0, 0, 0, 0};
    uint64_t i;
    int found = 0;
    if (source >= n_or || sink >= n_or) return 0;
    if (_step) *_step = 0;
    if (_dist) *_dist = 0;
    if (step == 0) step = UINT32_MAX;
    if (dist == 0) dist = UINT64_MAX;
    seen = (uint8_t *) calloc(n_or, 1);
    fifo_push(&q, source, 0, 0);
    while (q.head < q.n && !found) {
        const hop_t h = q.a[q.head++];
        const asmg_arc_t *a = &g->arc[g->idx_p[h.v]];
        if (seen[h.v]) continue;
        seen[h.v] = 1;
        if (!(h.step < step && h.dist < dist)) continue;
        for (i = 0; i < g->idx_n[h.v]; ++i) {
            if (a[i].w == sink) {
                found = 1;
                if (_step) *_step = h.step;
                if (_dist) *_dist = h.dist;
                break;
            }
            if (!seen[a[i].w]) fifo_push(&q, a[i].w, h.step + 1, h.dist + g->vtx[a[i].w >> 1].len - a[i].ls);
        }
    }
    free(q.a); free(seen);
    return found;
}

int asmg_tarjans_scc(asmg_t *g, int *scc)
{
    const uint64_t n_or = g->n_vtx << 1;
    int *low = (int *) malloc(sizeof(int) * (n_or ? n_or : 1)), *disc = (int *) malloc(sizeof(int) * (n_or ? n_or : 1));
    uint8_t *on_stack = (uint8_t *) calloc(n_or ? n_or : 1, 1);
    uint64_t *stack = (uint64_t *) malloc(sizeof(uint64_t) * (n_or ? n_or : 1)), *call = (uint64_t *) malloc(sizeof(uint64_t) * 2 * (n_or ? n_or : 1));
    uint64_t top = 0, depth_call, r;
    int n_scc = 0, clock = 0;
    for (r = 0; r < n_or; ++r) { scc[r] = low[r] = disc[r] = -1; }
    for (r = 0; r < n_or; ++r) {
        if (disc[r] != -1 || g->vtx[r >> 1].del) continue;
        depth_call = 0;
        call[0] = r; call[1] = 0;
        disc[r] = low[r] = ++clock; stack[top++] = r; on_stack[r] = 1;
        while (depth_call != (uint64_t) -1) {
            const uint64_t v = call[2 * depth_call];
            const asmg_arc_t *a = &g->arc[g->idx_p[v]];
            uint64_t i = call[2 * depth_call + 1], w = 0;
            int descend = 0;
            for (; i < g->idx_n[v]; ++i) {
                if (a[i].del) continue;
                w = a[i].w;
                if (g->vtx[w >> 1].del) continue;
                if (disc[w] == -1) { descend = 1; break; }
                if (on_stack[w] && disc[w] < low[v]) low[v] = disc[w];
            }
            if (descend) {
                call[2 * depth_call + 1] = i;             /* come back to this arc: its child's low-link is merged then */
                ++depth_call;
                call[2 * depth_call] = w; call[2 * depth_call + 1] = 0;
                disc[w] = low[w] = ++clock; stack[top++] = w; on_stack[w] = 1;
                continue;
            }
            if (low[v] == disc[v]) {
                uint64_t x;
                do { x = stack[--top]; on_stack[x] = 0; scc[x] = n_scc; } while (x != v);
                ++n_scc;
            }
            if (depth_call-- == 0) break;
            {   /* back in the parent: take the finished child's low-link, move past its arc */
                const uint64_t p = call[2 * depth_call];
                if (low[v] < low[p]) low[p] = low[v];
                ++call[2 * depth_call + 1];
            }
        }
    }
    free(low); free(disc); free(on_stack); free(stack); free(call);
    return n_scc;
}
