/*
 * report_gpu.c -- the reporting helpers of the reference's headers, so that callers that use them (verbose runs,
 * debugging builds) link against this layer unchanged:
 *   syncasm.h   scg_print :825, scg_print_unitig_syncmer_list :857, scg_subgraph_stat :423, scg_arc_coverage :309,
 *               scg_ra_print / scg_rv_print (alignment.c:686-708)
 *   syncmer.h   print_syncmer_on_seq :1164, print_all_syncmers_on_seq :1178, print_aligned_syncmers_on_seq :1184,
 *               get_hoco_seq :1211
 * Text formats are the reference's, byte for byte (tests/test_report_cpu.py compares them on the same structs).
 */
#include <stdlib.h>
#include <string.h>
#include "graph_gpu.h"

void scg_print(scg_t *g, FILE *fo, int no_seq)
{
    const asmg_t *a = g->utg_asmg;
    uint64_t i;
    fprintf(fo, "H\tVN:Z:1.0\n");
    for (i = 0; i < a->n_vtx; ++i) {
        const asmg_vtx_t *u = &a->vtx[i];
        const int l = (int) u->len, c = (int) u->cov;
        if (u->del) continue;
        fprintf(fo, "S\tu%lu\t", (unsigned long) i);
        if (u->seq && !no_seq) fprintf(fo, "%.*s", l, u->seq);
        else fputc('*', fo);
        fprintf(fo, "\tLN:i:%d\tKC:i:%ld\tSC:f:%.3f\n", l, (long) ((int64_t) l * c), (double) c);
    }
    for (i = 0; i < a->n_arc; ++i) {
        const asmg_arc_t *e = &a->arc[i];
        if (e->del || e->comp) continue;
        fprintf(fo, "L\tu%lu\t%c\tu%lu\t%c\t%ldM\tEC:i:%u\n", (unsigned long) (e->v >> 1), "+-"[e->v & 1], (unsigned long) (e->w >> 1), "+-"[e->w & 1], (long) e->ls, (unsigned) e->cov);
        fprintf(fo, "L\tu%lu\t%c\tu%lu\t%c\t%ldM\tEC:i:%u\n", (unsigned long) (e->w >> 1), "-+"[e->w & 1], (unsigned long) (e->v >> 1), "-+"[e->v & 1], (long) e->ls, (unsigned) e->cov);
    }
}

void scg_print_unitig_syncmer_list(scg_t *g, FILE *fo)
{
    const asmg_t *a = g->utg_asmg;
    const syncmer_t *scm = g->scm_db->a;
    const uint64_t *hap = g->scm_db->h;
    uint64_t i, j;
    for (i = 0; i < a->n_vtx; ++i) {
        const asmg_vtx_t *u = &a->vtx[i];
        if (u->del) continue;
        fprintf(fo, "u%lu syncmer list:", (unsigned long) i);
        for (j = 0; j < u->n; ++j) fprintf(fo, " %lu%c[%u]", (unsigned long) (u->a[j] >> 1), "+-"[u->a[j] & 1], (unsigned) scm[u->a[j] >> 1].cov);
        fputc('\n', fo);
        if (!hap) continue;
        fprintf(fo, "u%lu syncmer hap list:", (unsigned long) i);
        for (j = 0; j < u->n; ++j) {
            const uint64_t h = hap[u->a[j] >> 1];
            fprintf(fo, " %u:%u%c[%u]", (uint32_t) (h >> 33), (uint32_t) (h >> 1), "+-"[h & 1], (unsigned) scm[u->a[j] >> 1].cov);
        }
        fputc('\n', fo);
    }
}

void scg_ra_print(scg_ra_t *ra, FILE *fo)
{
    uint32_t i;
    fprintf(fo, "RID %lu [N %u S %.3f]:", (unsigned long) ra->sid, ra->n, ra->s);
    for (i = 0; i < ra->n; ++i) {
        const ra_frg_t *f = &ra->a[i];
        fprintf(fo, " [u%lu%c %lu %lu %u %u]", (unsigned long) (f->uid >> 1), "+-"[f->uid & 1], (unsigned long) f->u_beg, (unsigned long) f->u_end, f->s_beg, f->s_end);
    }
    fputc('\n', fo);
}

void scg_rv_print(scg_ra_v *rv, FILE *fo)
{
    size_t i;
    for (i = 0; i < rv->n; ++i) scg_ra_print(&rv->a[i], fo);
}

/* one block of four lines per connected component (both strands, live arcs), components in order of their lowest unitig */
void scg_subgraph_stat(scg_t *scg, FILE *fo)
{
    const asmg_t *a = scg->utg_asmg;
    const uint64_t n = a->n_vtx;
    uint32_t *comp = (uint32_t *) malloc(sizeof(uint32_t) * (n ? n : 1)), n_comp = 0, c;
    uint64_t *stack = (uint64_t *) malloc(sizeof(uint64_t) * (n ? n : 1)), top, i, j;
    uint64_t *n_utg, *n_scm, *n_arc, *seed;
    memset(comp, 0xff, sizeof(uint32_t) * (n ? n : 1));
    /* deleted unitigs are neither entered nor seeded (the reference seeds them too and then reads the first element of
     * an empty list, syncasm.c:440-452; unitig graphs have none after asmg_finalize) */
    for (i = 0; i < n; ++i) {
        if (comp[i] != UINT32_MAX || a->vtx[i].del) continue;
        comp[i] = n_comp; stack[0] = i; top = 1;
        while (top) {
            const uint64_t u = stack[--top];
            int o;
            for (o = 0; o < 2; ++o) {
                const asmg_arc_t *e = &a->arc[a->idx_p[u << 1 | o]];
                for (j = 0; j < a->idx_n[u << 1 | o]; ++j) {
                    const uint64_t w = e[j].w >> 1;
                    if (e[j].del || a->vtx[w].del || comp[w] != UINT32_MAX) continue;
                    comp[w] = n_comp; stack[top++] = w;
                }
            }
        }
        ++n_comp;
    }
    n_utg = (uint64_t *) calloc(4 * (size_t) (n_comp ? n_comp : 1), sizeof(uint64_t)); n_scm = n_utg + n_comp; n_arc = n_scm + n_comp; seed = n_arc + n_comp;
    for (i = n; i-- > 0; ) if (comp[i] != UINT32_MAX) { ++n_utg[comp[i]]; n_scm[comp[i]] += a->vtx[i].n; seed[comp[i]] = i; }
    for (i = 0; i < a->n_arc; ++i) {
        const asmg_arc_t *e = &a->arc[i];
        if (!e->del && comp[e->v >> 1] != UINT32_MAX && comp[e->v >> 1] == comp[e->w >> 1]) ++n_arc[comp[e->v >> 1]];
    }
    for (c = 0; c < n_comp; ++c) {
        fprintf(fo, "[M::%s] syncmer graph stats for subgraph %u - seeding u%u\n", __func__, c, (uint32_t) seed[c]);
        fprintf(fo, "[M::%s] number unitigs  : %u\n", __func__, (uint32_t) n_utg[c]);
        fprintf(fo, "[M::%s] number syncmers : %lu\n", __func__, (unsigned long) n_scm[c]);
        fprintf(fo, "[M::%s] number arcs     : %lu\n", __func__, (unsigned long) n_arc[c]);
    }
    free(comp); free(stack); free(n_utg);
}

/* arc coverage = how often the reads step from the last syncmer of v to the first of w, either strand (syncasm.c:309-368) */
typedef struct { uint64_t v, w, cnt; } step_t;

static int step_cmp(const void *x, const void *y)
{
    const step_t *a = (const step_t *) x, *b = (const step_t *) y;
    if (a->v != b->v) return a->v < b->v ? -1 : 1;
    return (a->w > b->w) - (a->w < b->w);
}

static step_t *step_find(step_t *t, size_t n, uint64_t v, uint64_t w)
{
    size_t lo = 0, hi = n;
    while (lo < hi) {
        const size_t mid = (lo + hi) >> 1;
        if (t[mid].v < v || (t[mid].v == v && t[mid].w < w)) lo = mid + 1; else hi = mid;
    }
    return lo < n && t[lo].v == v && t[lo].w == w ? &t[lo] : 0;
}

static void arc_ends(const asmg_t *g, const asmg_arc_t *a, uint64_t *v, uint64_t *w)
{
    const asmg_vtx_t *x = &g->vtx[a->v >> 1], *y = &g->vtx[a->w >> 1];
    *v = (a->v & 1) ? x->a[0] ^ 1 : x->a[x->n - 1];
    *w = (a->w & 1) ? y->a[y->n - 1] ^ 1 : y->a[0];
}

void scg_arc_coverage(scg_t *scg, sr_db_t *sr_db)
{
    asmg_t *g = scg->utg_asmg;
    step_t *t = (step_t *) malloc(sizeof(step_t) * (g->n_arc ? g->n_arc : 1)), *p;
    size_t n = 0, m = 0;
    uint64_t i, j, v, w;
    for (i = 0; i < g->n_arc; ++i) {
        if (g->arc[i].del) continue;
        arc_ends(g, &g->arc[i], &t[n].v, &t[n].w);
        t[n++].cnt = 0;
    }
    qsort(t, n, sizeof(step_t), step_cmp);
    for (i = 0; i < n; ++i) if (m == 0 || step_cmp(&t[m - 1], &t[i]) != 0) t[m++] = t[i];   /* one entry per syncmer pair */
    for (i = 0; i < sr_db->n; ++i) {
        const sr_t *s = &sr_db->a[i];
        if (s->n == 0) continue;
        v = s->k_mer[0] >> 1 << 1 | (s->m_pos[0] & 1);
        for (j = 1; j < s->n; ++j, v = w) {
            w = s->k_mer[j] >> 1 << 1 | (s->m_pos[j] & 1);
            if (!(p = step_find(t, m, v, w))) continue;
            ++p->cnt;
            if ((w ^ 1) != v && (p = step_find(t, m, w ^ 1, v ^ 1))) ++p->cnt;
        }
    }
    for (i = 0; i < g->n_arc; ++i) {
        if (g->arc[i].del) continue;
        arc_ends(g, &g->arc[i], &v, &w);
        g->arc[i].cov = (uint32_t) step_find(t, m, v, w)->cnt;
    }
    free(t);
}

/* ---------- syncmer.h ---------- */
void get_hoco_seq(sr_t *sr, kstring_t *s)
{
    uint32_t i;
    if (s->m < (size_t) sr->hoco_l + 1) { s->m = (size_t) sr->hoco_l + 1; s->s = (char *) realloc(s->s, s->m); }
    for (i = 0; i < sr->hoco_l; ++i) s->s[i] = char_nt4_table[(sr->hoco_s[i >> 2] >> (((i & 3) ^ 3) << 1)) & 3];
    s->l = sr->hoco_l;
}

void print_syncmer_on_seq(sr_t *sr, uint32_t n, int k, int w, FILE *fo)
{
    char *kmer;
    int i;
    if (n >= sr->n || (sr->m_pos[n] >> 1) == MAX_RD_LEN) return;                /* corrected entries have no position */
    fprintf(fo, ">%lu_%d_%u_%lu_%u\t", (unsigned long) sr->sid, (int) n, sr->m_pos[n] >> 1, (unsigned long) (sr->s_mer[n] & 1), sr->m_pos[n] & 1);
    fprintf(fo, "RD:Z:%lu\tMM:Z:", (unsigned long) sr->sid);
    for (i = 0; i < k; ++i) fputc(char_nt4_table[(sr->s_mer[n] >> 1 >> ((k - i - 1) * 2)) & 3], fo);
    fprintf(fo, "\tKH:Z:%lu\n", (unsigned long) sr->k_mer[n]);
    kmer = (char *) malloc((size_t) w + 1);
    get_kmer_dna_seq(sr->hoco_s, sr->m_pos[n] >> 1, w, sr->m_pos[n] & 1, kmer);
    fwrite(kmer, 1, (size_t) w, fo);
    fputc('\n', fo);
    free(kmer);
}

void print_all_syncmers_on_seq(sr_t *sr, int k, int w, FILE *fo)
{
    uint32_t i;
    for (i = 0; i < sr->n; ++i) print_syncmer_on_seq(sr, i, k, w, fo);
}

void print_aligned_syncmers_on_seq(sr_t *sr, int w, uint32_t beg, uint32_t end, FILE *fo)
{
    kstring_t s = {0, 0, 0};
    uint32_t i, j;
    if (end > sr->n) end = sr->n;
    get_hoco_seq(sr, &s);
    fprintf(fo, "%.*s\n", (int) s.l, s.s ? s.s : "");
    for (i = beg; i < end; ++i) {
        const uint32_t p = sr->m_pos[i] >> 1;
        if (p == MAX_RD_LEN) continue;
        for (j = 0; j < p; ++j) fputc('*', fo);
        fprintf(fo, "%.*s", w, &s.s[p]);
        for (j = p + (uint32_t) w; j < sr->hoco_l; ++j) fputc('*', fo);
        fputc('\n', fo);
    }
    free(s.s);
}
