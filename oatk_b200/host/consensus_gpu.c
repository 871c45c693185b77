/*
 * consensus_gpu.c -- row f1 of SURVEY.md section 8: unitig sequences and the GFA text.
 *
 * What the reference computes (syncasm.c:477-582 calc_syncmer_overlap, :888-1001
 * scg_syncmer_consensus, :1004-1046 scg_unitig_consensus, :630-670 utg_avg_cov,
 * :716-823 scg_consensus), restated from its behaviour:
 *
 *   offset(m1, m2)  For two syncmers that follow each other on a unitig: over all reads that
 *                   carry both as NEIGHBOURS in the right relative orientation, the difference
 *                   of their hoco start positions; the most frequent value wins. Ties go to the
 *                   value met first when the reference walks its khashl table in slot order, and
 *                   that table keeps its capacity while one unitig is processed -- so the table
 *                   below reproduces khashl 0.1's bucket function, probing, growth rule and its
 *                   in-place kick-out rehash.
 *   bases(m, from)  hoco bases [from, k) of a syncmer from its first occurrence that was not
 *                   error-corrected; every hoco base is written 1 + lround(mean run length - 1)
 *                   times, the mean taken over all such occurrences (0..255 from ho_rl, longer
 *                   runs from the ho_l_rl side list). A negative `from` pads with N.
 *   unitig          offsets accumulate into positions; a syncmer that starts before the end of
 *                   what is already written is skipped unless it is the last one that does.
 *   coverage        mean of the single-copy syncmers' coverages inside [Q1-1.5 IQR, Q3+1.5 IQR]
 *                   (all syncmers when none is single copy).
 *   GFA             H line, one S line per unitig (LN, KC = (int64)(len*cov), SC %.3f), two L
 *                   lines per arc with the overlap in bases and EC = arc coverage.
 *
 * Everything here runs on the host over the caller's sr_db_t / scg_t (the structs of
 * syncmer_gpu.h / graph_gpu.h): graphs after the coverage filter hold a few 10^4 syncmers.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <assert.h>
#include <pthread.h>
#include <unistd.h>
#include "graph_gpu.h"

/* test probes: how often the offset vote was decided by slot order, and how often a vote table grew past
 * its first four buckets (tests/test_consensus_cpu.py makes sure both paths are exercised) */
static uint64_t dbg_vote_ties, dbg_table_growths;   /* bumped with relaxed atomics: unitigs are processed in parallel */
void oatk_consensus_debug_counts(uint64_t *out) { out[0] = dbg_vote_ties; out[1] = dbg_table_growths; }

/* ---------------------------------------------------------------- int -> count, khashl slot order */
typedef struct { int key, val; } ovl_cell_t;
typedef struct {
    uint32_t bits, count;
    int live;                  /* 0: logically no table yet (khashl's keys == NULL), whatever storage is kept below */
    uint32_t alloc;            /* buckets the storage below can hold */
    uint32_t *used, *spare;    /* one bit per bucket; spare is the zeroed bitmap the next growth fills */
    ovl_cell_t *cell;
} ovl_tab_t;

static inline uint32_t ot_words(uint32_t n_buckets) { return n_buckets < 32 ? 1 : n_buckets >> 5; }
static inline int ot_used(const uint32_t *u, uint32_t i) { return (u[i >> 5] >> (i & 31)) & 1u; }
static inline void ot_set(uint32_t *u, uint32_t i) { u[i >> 5] |= 1u << (i & 31); }
static inline void ot_unset(uint32_t *u, uint32_t i) { u[i >> 5] &= ~(1u << (i & 31)); }
/* khashl.h:82 with the identity hash the reference gives this map (syncasm.c:63) */
static inline uint32_t ot_bucket(int key, uint32_t bits) { return (uint32_t) key * 2654435769u >> (32 - bits); }

/* kh_clear: same capacity, no entries */
static void ot_clear(ovl_tab_t *t)
{
    if (t->live) { memset(t->used, 0, ot_words(1u << t->bits) * sizeof(uint32_t)); t->count = 0; }
}
/* a table as kh_init leaves it (no buckets at all), keeping the storage for the next use */
static void ot_fresh(ovl_tab_t *t) { t->live = 0; t->bits = 0; t->count = 0; }

static void ot_free(ovl_tab_t *t) { free(t->used); free(t->spare); free(t->cell); memset(t, 0, sizeof(*t)); }

static void ot_reserve(ovl_tab_t *t, uint32_t buckets)
{
    if (buckets <= t->alloc) return;
    uint32_t *u = (uint32_t *) calloc(ot_words(buckets), sizeof(uint32_t)), *sp = (uint32_t *) calloc(ot_words(buckets), sizeof(uint32_t));
    if (t->live) memcpy(u, t->used, ot_words(1u << t->bits) * sizeof(uint32_t));
    free(t->used); free(t->spare);
    t->used = u; t->spare = sp;
    t->cell = (ovl_cell_t *) realloc(t->cell, buckets * sizeof(ovl_cell_t));
    t->alloc = buckets;
}

/* growth to the next power of two >= want (at least 4), moving the cells inside the same array the
 * way khashl does it (khashl.h:144-187): walk the old buckets in order, drop each cell at its new
 * home and carry on with whatever old cell was sitting there */
static void ot_grow(ovl_tab_t *t, uint32_t want)
{
    uint32_t j = 0, x = want, old_n = t->live ? 1u << t->bits : 0u, new_bits, new_n, mask;
    while ((x >>= 1) != 0) ++j;
    if (want & (want - 1)) ++j;
    new_bits = j > 2 ? j : 2;
    new_n = 1u << new_bits;
    if (t->count > (new_n >> 1) + (new_n >> 2)) return;
    if (old_n) __atomic_fetch_add(&dbg_table_growths, 1, __ATOMIC_RELAXED);
    ot_reserve(t, new_n);
    uint32_t *nu = t->spare;
    memset(nu, 0, ot_words(new_n) * sizeof(uint32_t));
    mask = new_n - 1;
    for (j = 0; j != old_n; ++j) {
        if (!ot_used(t->used, j)) continue;
        ovl_cell_t c = t->cell[j];
        ot_unset(t->used, j);
        for (;;) {
            uint32_t i = ot_bucket(c.key, new_bits);
            while (ot_used(nu, i)) i = (i + 1) & mask;
            ot_set(nu, i);
            if (i < old_n && ot_used(t->used, i)) {
                ovl_cell_t tmp = t->cell[i]; t->cell[i] = c; c = tmp;
                ot_unset(t->used, i);
            } else { t->cell[i] = c; break; }
        }
    }
    t->spare = t->used; t->used = nu;
    t->bits = new_bits; t->live = 1;
}

static void ot_count(ovl_tab_t *t, int key)
{
    uint32_t n = t->live ? 1u << t->bits : 0u;
    if (t->count >= (n >> 1) + (n >> 2)) { ot_grow(t, n + 1); n = 1u << t->bits; }
    const uint32_t mask = n - 1;
    uint32_t i = ot_bucket(key, t->bits), first = i;
    while (ot_used(t->used, i) && t->cell[i].key != key) { i = (i + 1) & mask; if (i == first) break; }
    if (!ot_used(t->used, i)) { t->cell[i].key = key; t->cell[i].val = 1; ot_set(t->used, i); ++t->count; }
    else ++t->cell[i].val;
}

/* the value with the highest count; among equals the first in slot order */
static int ot_mode(const ovl_tab_t *t)
{
    int best = 0, best_n = 0, tied = 0;
    if (!t->live) return 0;
    for (uint32_t i = 0, n = 1u << t->bits; i < n; ++i) {
        if (!ot_used(t->used, i)) continue;
        if (t->cell[i].val > best_n) { best_n = t->cell[i].val; best = t->cell[i].key; tied = 0; }
        else if (t->cell[i].val == best_n) tied = 1;
    }
    if (tied) __atomic_fetch_add(&dbg_vote_ties, 1, __ATOMIC_RELAXED);
    return best;
}

/* ---------------------------------------------------------------- distance between two neighbours */
static inline int occ_is_corrected(const sr_db_t *db, uint64_t occ)
{
    return (int) (db->a[occ >> 32].k_mer[occ >> 1 & MAX_RD_SCM] & 1);
}
static inline int64_t occ_start(const sr_db_t *db, uint64_t occ)
{
    return (int64_t) (db->a[occ >> 32].m_pos[occ >> 1 & MAX_RD_SCM] >> 1);
}

/* m1 (strand rc1) is followed by m2 (strand rc2): hoco distance between their starts */
/* `t` keeps its capacity from call to call when `carry` is set (one unitig), else it starts as an empty table */
static int neighbour_offset(const sr_db_t *db, const syncmer_t *m1, uint64_t rc1, const syncmer_t *m2, uint64_t rc2, ovl_tab_t *t, int carry)
{
    if (!carry) ot_fresh(t);
    const uint64_t *o1 = m1->m_pos, *o2 = m2->m_pos;
    const uint64_t n1 = m1->cov, n2 = m2->cov;
    assert(n1 > 0 && n2 > 0);
    ot_clear(t);
    /* every occurrence is looked up in its read (was it corrected, where does it start): bring the reads' headers in
     * first, then the list entries they point at, before the merge below asks for them one by one */
    for (uint64_t a = 0; a < n1; ++a) __builtin_prefetch(&db->a[o1[a] >> 32], 0, 1);
    for (uint64_t b = 0; b < n2; ++b) __builtin_prefetch(&db->a[o2[b] >> 32], 0, 1);
    for (uint64_t a = 0; a < n1; ++a) __builtin_prefetch(&db->a[o1[a] >> 32].m_pos[o1[a] >> 1 & MAX_RD_SCM], 0, 1);
    for (uint64_t b = 0; b < n2; ++b) __builtin_prefetch(&db->a[o2[b] >> 32].m_pos[o2[b] >> 1 & MAX_RD_SCM], 0, 1);
    uint64_t j0 = 0;                                  /* first occurrence of m2 on a read >= the current one */
    for (uint64_t a = 0; a < n1; ++a) {
        const uint64_t read = o1[a] >> 32, i1 = o1[a] >> 1 & MAX_RD_SCM, s1 = o1[a] & 1;
        if (occ_is_corrected(db, o1[a])) continue;
        while (j0 < n2 && (o2[j0] >> 32) < read) ++j0;
        for (uint64_t b = j0; b < n2 && (o2[b] >> 32) == read; ++b) {
            const uint64_t i2 = o2[b] >> 1 & MAX_RD_SCM, s2 = o2[b] & 1;
            if (occ_is_corrected(db, o2[b])) continue;
            /* read walks the pair forwards (m2 right after m1, both on the asked strands) or backwards */
            if (i1 == i2 + 1 && s1 != rc1 && s2 != rc2) ot_count(t, (int) (occ_start(db, o1[a]) - occ_start(db, o2[b])));
            else if (i1 + 1 == i2 && s1 == rc1 && s2 == rc2) ot_count(t, (int) (occ_start(db, o2[b]) - occ_start(db, o1[a])));
        }
    }
    return ot_mode(t);
}

/* ---------------------------------------------------------------- growing text buffer */
typedef struct { size_t l, m; char *s; } txt_t;
static inline void txt_put(txt_t *t, int c)
{
    if (t->l + 1 > t->m) { t->m = t->m ? t->m << 1 : 256; t->s = (char *) realloc(t->s, t->m); }
    t->s[t->l++] = (char) c;
}

/* ---------------------------------------------------------------- bases of one syncmer */
static inline void txt_room(txt_t *t, size_t extra)
{
    if (t->l + extra + 1 > t->m) { t->m = (t->l + extra + 1) * 2 + 256; t->s = (char *) realloc(t->s, t->m); }
}

__attribute__((optimize("O3"))) static void lane_add(uint16_t *restrict lane, const uint8_t *restrict rl, uint64_t l)
{
    for (uint64_t j = 0; j < l; ++j) lane[j] = (uint16_t) (lane[j] + rl[j]);
}

__attribute__((optimize("O3"))) static void lane_add_rev(uint16_t *restrict lane, const uint8_t *restrict rl, uint64_t l)
{
    for (uint64_t j = 0; j < l; ++j) lane[j] = (uint16_t) (lane[j] + rl[l - 1 - j]);
}

/* Run-length sums fetched from the device for a set of syncmers (read databases whose ho_rl stayed there): row i holds, for
 * syncmer ids[i] in its own orientation, sum over the uncorrected copies of (run length - 1) per hoco position. */
typedef struct {
    const int64_t *row_of;        /* per syncmer id: row, or -1 */
    const uint64_t *sums;         /* rows x w, or NULL when the reads carry ho_rl */
    const uint32_t *copies;       /* per row: uncorrected copies */
    const uint8_t *codes;         /* rows x w hoco codes of the row's first uncorrected occurrence as it lies on its read, or NULL
                                     when the reads carry hoco_s */
} rl_table_t;
/* l codes from position p of a k-mer whose w codes (as on the read) are `fwd`: forwards, or reverse-complemented like get_kmer_seq */
static void codes_from_row(const uint8_t *fwd, uint64_t p, uint64_t l, uint64_t r, uint8_t *out)
{
    for (uint64_t i = 0; i < l; ++i) out[i] = (uint8_t) (r ? 3 - fwd[p + l - 1 - i] : fwd[p + i]);
}

/* out == NULL: only the length is wanted (arc overlaps) */
static int64_t syncmer_text(const sr_db_t *db, const syncmer_t *m, int rev, int64_t from, txt_t *out, int hoco_only, const rl_table_t *rlt, uint64_t id)
{
    const int w = db->k;
    assert(from < w);
    int64_t written = from < 0 ? -from : 0;
    if (out) for (int64_t i = from; i < 0; ++i) txt_put(out, 'N');
    if (from < 0) from = 0;
    const uint64_t l = (uint64_t) (w - from);
    written += (int64_t) l;
    if (hoco_only && !out) return written;              /* one character per hoco base whatever the copies say */

    /* the first occurrence that read error correction left alone supplies the bases */
    uint32_t i;
    const sr_t *s = 0;
    uint64_t p = 0, r = 0;
    for (i = 0; i < m->cov; ++i) {
        if (occ_is_corrected(db, m->m_pos[i])) continue;
        s = &db->a[m->m_pos[i] >> 32];
        p = s->m_pos[m->m_pos[i] >> 1 & MAX_RD_SCM];
        r = (p & 1) ^ (uint64_t) rev;
        p >>= 1;
        break;
    }
    if (i == m->cov) {                                 /* every copy was corrected away: N (reference :925-931) */
        if (out) { txt_room(out, l); memset(out->s + out->l, 'N', l); out->l += l; }
        return written;
    }
    if (!r) p += (uint64_t) from;
    /* the occurrence's bases: from the read, or -- for reads whose packed bases stayed on the device -- from the row the
     * device sent for this syncmer (its w codes as they lie on the read, from the occurrence's start) */
    const uint8_t *row_codes = 0;
    if (!s->hoco_s) {
        assert(rlt && rlt->codes && rlt->row_of[id] >= 0);
        row_codes = rlt->codes + (size_t) rlt->row_of[id] * (size_t) w;
    }
    if (hoco_only) {
        txt_room(out, l);
        if (row_codes) {
            uint8_t *dst = (uint8_t *) out->s + out->l;
            codes_from_row(row_codes, r ? 0 : (uint64_t) from, l, r, dst);
            for (uint64_t j = 0; j < l; ++j) dst[j] = (uint8_t) char_nt4_table[dst[j]];
        } else get_kmer_dna_seq(s->hoco_s, (uint32_t) p, (int) l, (uint32_t) r, out->s + out->l);
        out->l += l;
        return written;
    }
    /* summed run lengths - 1 per hoco base over the uncorrected occurrences, in the syncmer's orientation. This is
     * the bulk of the consensus (bases x coverage additions): the one-byte run lengths are added into 16-bit lanes
     * with plain loops the compiler vectorises, folded into the 64-bit sums every 256 copies (255 * 256 < 2^16); the
     * rare runs of 255 or more, kept in a side list per read, are patched in on top */
    uint64_t *tot = (uint64_t *) calloc(l, sizeof(uint64_t));
    uint32_t copies = 0, pending = 0;
    if (rlt && rlt->sums) {
        /* the device summed the copies in the syncmer's forward frame F[0..w): position j of this view is F[from + j]
         * on the forward strand and F[w - 1 - from - j] on the reverse one */
        const int64_t row = rlt->row_of[id];
        assert(row >= 0);
        const uint64_t *F = rlt->sums + (uint64_t) row * (uint64_t) w;
        copies = rlt->copies[row];
        for (uint64_t j = 0; j < l; ++j) tot[j] = rev ? F[(uint64_t) w - 1 - (uint64_t) from - j] : F[(uint64_t) from + j];
    } else {
    uint16_t *lane = (uint16_t *) calloc(l, sizeof(uint16_t));
    /* the copies sit in as many different reads: their run lengths are fetched a batch ahead of the additions */
    enum { AHEAD = 16 };
    struct { const sr_t *t; const uint8_t *rl8; uint64_t q; int rr; } nxt[AHEAD];
    for (uint32_t i0 = 0; i0 < m->cov; i0 += AHEAD) {
        uint32_t nb = 0;
        for (i = i0; i < m->cov && i < i0 + AHEAD; ++i) {
            if (occ_is_corrected(db, m->m_pos[i])) continue;
            const sr_t *t = &db->a[m->m_pos[i] >> 32];
            uint64_t q = t->m_pos[m->m_pos[i] >> 1 & MAX_RD_SCM];
            const int rr = (int) ((q & 1) ^ (uint64_t) rev);
            q >>= 1;
            if (!rr) q += (uint64_t) from;
            nxt[nb].t = t; nxt[nb].q = q; nxt[nb].rr = rr; nxt[nb].rl8 = t->ho_rl + q;
            for (uint64_t o = 0; o < l; o += 64) __builtin_prefetch(nxt[nb].rl8 + o, 0, 1);
            ++nb;
        }
        for (uint32_t b = 0; b < nb; ++b) {
            const sr_t *t = nxt[b].t;
            const uint8_t *rl8 = nxt[b].rl8;
            const uint64_t q = nxt[b].q, rr = (uint64_t) nxt[b].rr;
            if (rr) lane_add_rev(lane, rl8, l);
            else lane_add(lane, rl8, l);
            if (t->ho_l_rl) {                          /* this read has long runs: swap the 255 marks for the real lengths */
                uint32_t side = 0;                     /* entries of the long-run list in front of q */
                for (uint64_t j = 0; j < q; ++j) side += t->ho_rl[j] == 255;
                for (const uint8_t *z = (const uint8_t *) memchr(rl8, 255, l); z; z = (const uint8_t *) memchr(z + 1, 255, l - (size_t) (z + 1 - rl8))) {
                    const uint64_t j = (uint64_t) (z - rl8);
                    tot[rr ? l - 1 - j : j] += (uint64_t) t->ho_l_rl[side++] - 255;
                }
            }
            ++copies;
            if (++pending == 256) { for (uint64_t j = 0; j < l; ++j) { tot[j] += lane[j]; lane[j] = 0; } pending = 0; }
        }
    }
    for (uint64_t j = 0; j < l; ++j) tot[j] += lane[j];
    free(lane);
    }
    if (!out) {
        for (uint64_t j = 0; j < l; ++j) written += lround((double) tot[j] / copies);
        free(tot);
        return written;
    }
    uint8_t *code = (uint8_t *) malloc(l);
    if (row_codes) codes_from_row(row_codes, r ? 0 : (uint64_t) from, l, r, code);
    else get_kmer_seq(s->hoco_s, (uint32_t) p, (int) l, (uint32_t) r, code);
    for (uint64_t j = 0; j < l; ++j) {
        const long extra = lround((double) tot[j] / copies);
        const char c = char_nt4_table[code[j]];
        txt_room(out, (size_t) extra + 1);
        memset(out->s + out->l, c, (size_t) extra + 1);
        out->l += (size_t) extra + 1;
        written += extra;
    }
    free(tot);
    free(code);
    return written;
}

/* ---------------------------------------------------------------- bases of a chain of syncmers */
static int64_t chain_text(const sr_db_t *db, const uint64_t *v, uint64_t n, const syncmer_t *scm, txt_t *out, int hoco_only, ovl_tab_t *tab, const rl_table_t *rlt)
{
    if (n == 0) return 0;
    const int w = db->k;
    ot_fresh(tab);                                      /* one table per chain: its capacity carries over inside it (reference :1011-1041) */
    int64_t *pos = (int64_t *) malloc(n * sizeof(int64_t));
    pos[0] = 0;
    for (uint64_t i = 1; i < n; ++i)
        pos[i] = pos[i - 1] + neighbour_offset(db, &scm[v[i - 1] >> 1], v[i - 1] & 1, &scm[v[i] >> 1], v[i] & 1, tab, 1);
    int64_t end = 0, len = 0;
    for (uint64_t i = 0; i < n; ++i) {
        while (i + 1 < n && pos[i + 1] <= end) ++i;     /* the next one still starts inside what is written: skip ahead */
        len += syncmer_text(db, &scm[v[i] >> 1], (int) (v[i] & 1), end - pos[i], out, hoco_only, rlt, v[i] >> 1);
        end = pos[i] + w;
    }
    free(pos);
    assert(len >= 0 && (uint64_t) len == out->l);
    return len;
}

/* ---------------------------------------------------------------- unitig coverage */
static int dbl_cmp(const void *a, const void *b)
{
    const double x = *(const double *) a, y = *(const double *) b;
    return (x > y) - (x < y);
}

static double quantile_sorted(const double *a, int n, double q)
{
    if (n == 1) return a[0];
    double whole;
    const double frac = modf(q * (n - 1), &whole);
    const int i = (int) lround(whole);
    return i == n - 1 ? a[i] : a[i] + (a[i + 1] - a[i]) * frac;
}

static double iqr_mean_sorted(const double *a, int n)
{
    if (n == 0) return 0.;
    double q1 = quantile_sorted(a, n, 0.25), q3 = quantile_sorted(a, n, 0.75);
    const double iqr = q3 - q1;
    q1 -= 1.5 * iqr; q3 += 1.5 * iqr;
    double sum = 0.;
    int kept = 0;
    for (int i = 0; i < n; ++i) if (a[i] >= q1 && a[i] <= q3) { sum += a[i]; ++kept; }
    return kept ? sum / kept : 0.;
}

static double unitig_coverage(const scg_t *g, const asmg_vtx_t *u)
{
    if (u->del) return 0.;
    const syncmer_t *scm = g->scm_db->a;
    double *c = (double *) calloc(u->n ? u->n : 1, sizeof(double));
    uint64_t i, first;
    for (i = 0; i < u->n; ++i) {                        /* syncmers that sit on exactly one unitig */
        const uint64_t id = u->a[i] >> 1;
        if (g->idx_u[id + 1] - g->idx_u[id] == 1) c[i] = scm[id].cov;
    }
    qsort(c, u->n, sizeof(double), dbl_cmp);
    for (first = 0; first < u->n && c[first] < DBL_EPSILON; ++first) {}
    if (first == u->n) {
        for (i = 0; i < u->n; ++i) c[i] = scm[u->a[i] >> 1].cov;
        qsort(c, u->n, sizeof(double), dbl_cmp);
        first = 0;
    }
    const double avg = iqr_mean_sorted(c + first, (int) (u->n - first));
    free(c);
    return avg;
}

/* ---------------------------------------------------------------- the graph as GFA */
static asmg_arc_t *arc_vw(asmg_t *g, uint64_t v, uint64_t w)
{
    asmg_arc_t *a = &g->arc[g->idx_p[v]];
    for (uint64_t i = 0, n = g->idx_n[v]; i < n; ++i) if (a[i].w == w) return &a[i];
    return 0;
}

/* Unitigs (and arcs) are independent of each other, so their texts and overlaps are computed by a few worker
 * threads that pull indices from a shared counter; the results are then written out in index order, which is
 * all the reference's single loop guarantees. */
typedef struct {
    sr_db_t *db; scg_t *scg; int hoco;
    char **text; int64_t *len; double *cov;            /* per unitig */
    int64_t *ovl;                                       /* per arc */
    uint64_t next;                                      /* shared work counter */
    int phase;
    const rl_table_t *rlt;                              /* run-length sums from the device, or NULL: the reads carry ho_rl */
} cons_job_t;

static int64_t arc_overlap(cons_job_t *J, const asmg_arc_t *a, txt_t *t, ovl_tab_t *tab)
{
    const sr_db_t *db = J->db;
    const syncmer_t *scm = J->scg->scm_db->a;
    asmg_t *G = J->scg->utg_asmg;
    const int w = db->k;
    int64_t l;
    if (a->ln > 0) {                                    /* the two unitigs share ln syncmers: their length in bases */
        const asmg_vtx_t *u = &G->vtx[a->v >> 1];
        t->l = 0;
        l = chain_text(db, (a->v & 1) ? u->a : &u->a[u->n - a->ln], a->ln, scm, t, J->hoco, tab, J->rlt);
    } else {                                            /* they abut: overlap of the two end syncmers */
        const asmg_vtx_t *u = &G->vtx[a->v >> 1];
        uint64_t z = a->v & 1;
        const uint64_t x = u->a[(u->n - 1) * (!z)] ^ z;
        u = &G->vtx[a->w >> 1];
        z = a->w & 1;
        const uint64_t y = u->a[(u->n - 1) * z] ^ z;
        l = neighbour_offset(db, &scm[x >> 1], x & 1, &scm[y >> 1], y & 1, tab, 0);
        if (l < w) l = syncmer_text(db, &scm[x >> 1], (int) (x & 1), l, 0, J->hoco, J->rlt, x >> 1);    /* length only */
        else l = 0;
    }
    return l;
}

#define CONS_BLOCK 64

static void *cons_worker(void *arg)
{
    cons_job_t *J = (cons_job_t *) arg;
    asmg_t *G = J->scg->utg_asmg;
    txt_t t = {0, 0, 0};
    ovl_tab_t tab;
    memset(&tab, 0, sizeof(tab));
    const uint64_t n_items = J->phase == 0 ? G->n_vtx : G->n_arc;
    for (;;) {                                           /* blocks of indices: most of an error-filtered graph is skipped */
        const uint64_t lo = __atomic_fetch_add(&J->next, CONS_BLOCK, __ATOMIC_RELAXED);
        const uint64_t hi = lo + CONS_BLOCK < n_items ? lo + CONS_BLOCK : n_items;
        if (lo >= n_items) break;
        for (uint64_t i = lo; i < hi; ++i) {
            if (J->phase == 0) {
                asmg_vtx_t *u = &G->vtx[i];
                if (u->del) continue;
                t.l = 0;
                J->len[i] = chain_text(J->db, u->a, u->n, J->scg->scm_db->a, &t, J->hoco, &tab, J->rlt);
                J->cov[i] = u->cov ? u->cov : unitig_coverage(J->scg, u);
                J->text[i] = (char *) malloc((size_t) J->len[i] + 1);
                memcpy(J->text[i], t.s, (size_t) J->len[i]);
                J->text[i][J->len[i]] = 0;
            } else {
                const asmg_arc_t *a = &G->arc[i];
                if (a->del || a->comp) continue;
                J->ovl[i] = arc_overlap(J, a, &t, &tab);
            }
        }
    }
    free(t.s);
    ot_free(&tab);
    return 0;
}

static void cons_run(cons_job_t *J, int phase, uint64_t n_items)
{
    long nt = oatk_host_threads();
    pthread_t th[16];
    if (n_items < 4) nt = 1;
    J->phase = phase; J->next = 0;
    if (nt == 1) { cons_worker(J); return; }
    for (long i = 0; i < nt; ++i) pthread_create(&th[i], 0, cons_worker, J);
    for (long i = 0; i < nt; ++i) pthread_join(th[i], 0);
}

/* run-length sums already fetched from the device for the syncmers of one read database (see scg_consensus) */
static struct { const sr_db_t *db; size_t n_scm; int64_t *row_of; uint64_t *sums; uint32_t *copies; uint8_t *codes; uint64_t n_rows; int with_sums, with_codes; } g_rl;
void oatk_cons_cache_drop(const sr_db_t *db)
{
    if (db && g_rl.db != db) return;
    free(g_rl.row_of); free(g_rl.sums); free(g_rl.copies); free(g_rl.codes);
    memset(&g_rl, 0, sizeof(g_rl));
}

/* Overlaps (asmg_arc_t.ls) of a list of arcs between single-syncmer vertices in homopolymer-compressed space, exactly as
 * scg_consensus(sr_db, scg, 1, ...) leaves them on the all-syncmer graph: for every arc that is not flagged as a
 * complement the most frequent start distance l of the two syncmers over the reads that carry both (neighbour_offset),
 * overlap = k - l when l < k, else 0, clipped to k; the value also goes to the FIRST arc (w^1, v^1) of the list -- in
 * list order, so that arcs nobody points at (a palindromic pair's second entry) keep 0 like they do there. The arcs must
 * be sorted by (v, w). Used by the graph-free form of read error correction (syncerr_gpu.c). */
typedef struct { const sr_db_t *db; const syncmer_t *scm; const uint64_t *arcs4; int64_t *l; const int32_t *dist; const uint8_t *flag; } ovl_job_t;
static void ovl_range(uint64_t lo, uint64_t hi, void *arg)
{
    ovl_job_t *J = (ovl_job_t *) arg;
    const int w = J->db->k;
    ovl_tab_t tab;
    memset(&tab, 0, sizeof(tab));
    for (uint64_t i = lo; i < hi; ++i) {
        const uint64_t x = J->arcs4[4 * i], y = J->arcs4[4 * i + 1];
        if (J->arcs4[4 * i + 3]) { J->l[i] = -1; continue; }        /* complement */
        /* the votes were counted on the device (sg_arc_votes) unless two distances tied there: the reference's answer
         * then depends on the slot order of its hash table, which only the table below reproduces */
        int64_t l = (J->flag && J->flag[i] == 0) ? J->dist[i]
                  : neighbour_offset(J->db, &J->scm[x >> 1], x & 1, &J->scm[y >> 1], y & 1, &tab, 0);
        if (l < w) l = syncmer_text(J->db, &J->scm[x >> 1], (int) (x & 1), l, 0, 1, 0, x >> 1);
        else l = 0;
        if (l > w) l = w;
        J->l[i] = l;
    }
    ot_free(&tab);
}

void oatk_syncmer_arc_overlaps(sr_db_t *sr_db, syncmer_db_t *scm_db, uint64_t n, const uint64_t *arcs4, uint32_t *ls)
{
    ovl_job_t J = {sr_db, scm_db->a, arcs4, (int64_t *) malloc(sizeof(int64_t) * (n ? n : 1)), 0, 0};
    int32_t *dist = (int32_t *) malloc(sizeof(int32_t) * (n ? n : 1));
    uint8_t *flag = (uint8_t *) malloc(n ? n : 1);
    if (n && !getenv("OATK_VOTES_HOST") && oatk_gpu_arc_votes(sr_db, n, arcs4, dist, flag) == 0) { J.dist = dist; J.flag = flag; }
    oatk_parallel_for(n, ovl_range, &J);
    free(dist); free(flag);
    for (uint64_t i = 0; i < n; ++i) ls[i] = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (J.l[i] < 0) continue;
        ls[i] = (uint32_t) J.l[i];
        const uint64_t mv = arcs4[4 * i + 1] ^ 1, mw = arcs4[4 * i] ^ 1;
        uint64_t lo = 0, hi = n;
        while (lo < hi) {                                            /* first arc with (v, w) >= (mv, mw) */
            const uint64_t mid = (lo + hi) >> 1;
            if (arcs4[4 * mid] < mv || (arcs4[4 * mid] == mv && arcs4[4 * mid + 1] < mw)) lo = mid + 1; else hi = mid;
        }
        if (lo < n && arcs4[4 * lo] == mv && arcs4[4 * lo + 1] == mw) ls[lo] = (uint32_t) J.l[i];
    }
    free(J.l);
}

void scg_consensus(sr_db_t *sr_db, scg_t *scg, int hoco_seq, int save_seq, FILE *fo)
{
    asmg_t *G = scg->utg_asmg;
    uint64_t i;
    cons_job_t J;

    for (i = 0; i < G->n_arc; ++i) G->arc[i].ls = 0;    /* graph.h:283-295 */
    for (i = 0; i < G->n_vtx; ++i) { free(G->vtx[i].seq); G->vtx[i].seq = 0; G->vtx[i].len = 0; }

    memset(&J, 0, sizeof(J));
    J.db = sr_db; J.scg = scg; J.hoco = hoco_seq;
    J.text = (char **) calloc(G->n_vtx ? G->n_vtx : 1, sizeof(char *));
    J.len = (int64_t *) calloc(G->n_vtx ? G->n_vtx : 1, sizeof(int64_t));
    J.cov = (double *) calloc(G->n_vtx ? G->n_vtx : 1, sizeof(double));
    J.ovl = (int64_t *) calloc(G->n_arc ? G->n_arc : 1, sizeof(int64_t));

    /* run lengths that stayed on the device: one request for every syncmer on a live unitig that was not asked for before,
     * answered by sg_runlen_sums. The sums of a syncmer depend on the reads' lists only, and syncasm() calls this function
     * three times on graphs made of the same syncmers (.utg.gfa, after unzipping, .utg.final.gfa), so the answers are kept
     * until the lists change (g_rl: dropped by oatk_cons_cache_drop from read_error_correction and sr_db_clean). */
    rl_table_t rlt;
    const int need_sums = !hoco_seq && oatk_gpu_run_lengths_on_device(sr_db);
    const int need_codes = oatk_gpu_bases_on_device(sr_db);                 /* sr_t.hoco_s == NULL: the bases come from the device too */
    if (need_sums || need_codes) {
        const syncmer_t *scm = scg->scm_db->a;
        const size_t n_scm = scg->scm_db->n;
        const size_t w = (size_t) sr_db->k;
        uint64_t n_new = 0, n_occ = 0, j, *ids, *occ_off, *occ, *refs;
        if (g_rl.db != sr_db || g_rl.n_scm != n_scm || (need_sums && !g_rl.with_sums) || (need_codes && !g_rl.with_codes)) {
            oatk_cons_cache_drop(0);
            g_rl.db = sr_db; g_rl.n_scm = n_scm; g_rl.with_sums = need_sums; g_rl.with_codes = need_codes;
            g_rl.row_of = (int64_t *) malloc(sizeof(int64_t) * (n_scm ? n_scm : 1));
            for (j = 0; j < n_scm; ++j) g_rl.row_of[j] = -1;
        }
        const uint64_t row0 = g_rl.n_rows;
        ids = 0;
        uint64_t m_ids = 0;
        for (i = 0; i < G->n_vtx; ++i) {
            const asmg_vtx_t *u = &G->vtx[i];
            if (u->del) continue;
            for (j = 0; j < u->n; ++j) {
                const uint64_t id = u->a[j] >> 1;
                if (g_rl.row_of[id] >= 0) continue;
                g_rl.row_of[id] = (int64_t) (row0 + n_new);
                if (n_new == m_ids) { m_ids = m_ids ? m_ids * 2 : 1024; ids = (uint64_t *) realloc(ids, sizeof(uint64_t) * m_ids); }
                ids[n_new++] = id;
                n_occ += scm[id].cov;
            }
        }
        if (n_new) {
            occ_off = (uint64_t *) malloc(sizeof(uint64_t) * (n_new + 1));
            occ = (uint64_t *) malloc(sizeof(uint64_t) * (n_occ ? n_occ : 1));
            refs = (uint64_t *) malloc(sizeof(uint64_t) * n_new);
            g_rl.copies = (uint32_t *) realloc(g_rl.copies, sizeof(uint32_t) * (row0 + n_new));
            if (g_rl.with_sums) g_rl.sums = (uint64_t *) realloc(g_rl.sums, sizeof(uint64_t) * (row0 + n_new) * w);
            if (g_rl.with_codes) g_rl.codes = (uint8_t *) realloc(g_rl.codes, (row0 + n_new) * w);
            n_occ = 0;
            for (j = 0; j < n_new; ++j) {
                const syncmer_t *m = &scm[ids[j]];
                occ_off[j] = n_occ;
                refs[j] = UINT64_MAX;
                for (uint32_t c = 0; c < m->cov; ++c) {
                    if (occ_is_corrected(sr_db, m->m_pos[c])) continue;
                    const uint64_t mp = sr_db->a[m->m_pos[c] >> 32].m_pos[m->m_pos[c] >> 1 & MAX_RD_SCM];
                    if (refs[j] == UINT64_MAX) refs[j] = (m->m_pos[c] >> 32) << 32 | (mp >> 1);    /* the first one supplies the bases */
                    occ[n_occ++] = (m->m_pos[c] >> 32) << 32 | mp;
                }
                g_rl.copies[row0 + j] = (uint32_t) (n_occ - occ_off[j]);
            }
            occ_off[n_new] = n_occ;
            oatk_tick("cons: run-length / base requests");
            if (g_rl.with_sums && oatk_gpu_runlen_sums(sr_db, n_new, occ_off, occ, g_rl.sums + row0 * w) != 0) {
                fprintf(stderr, "[E::%s] the run lengths could not be read from the device\n", __func__);
                exit(EXIT_FAILURE);
            }
            if (g_rl.with_codes) {
                /* syncmers every copy of which was corrected away have no text (N): they ask for nothing */
                uint64_t n_ref = 0, *pack = (uint64_t *) malloc(sizeof(uint64_t) * n_new);
                for (j = 0; j < n_new; ++j) if (refs[j] != UINT64_MAX) pack[n_ref++] = refs[j];
                uint8_t *tmp = (uint8_t *) malloc((n_ref ? n_ref : 1) * w);
                if (oatk_gpu_kmer_codes(sr_db, n_ref, pack, (int) w, tmp) != 0) {
                    fprintf(stderr, "[E::%s] the bases could not be read from the device\n", __func__);
                    exit(EXIT_FAILURE);
                }
                for (j = 0, n_ref = 0; j < n_new; ++j)
                    if (refs[j] != UINT64_MAX) memcpy(g_rl.codes + (row0 + j) * w, tmp + (n_ref++) * w, w);
                    else memset(g_rl.codes + (row0 + j) * w, 0, w);
                free(tmp); free(pack);
            }
            g_rl.n_rows = row0 + n_new;
            free(occ_off); free(occ); free(refs);
            oatk_tick("cons: run-length sums / bases from the device");
        }
        free(ids);
        rlt.row_of = g_rl.row_of; rlt.sums = g_rl.with_sums ? g_rl.sums : 0; rlt.copies = g_rl.copies; rlt.codes = g_rl.with_codes ? g_rl.codes : 0;
        J.rlt = &rlt;
    }

    cons_run(&J, 0, G->n_vtx);
    oatk_tick("cons: unitig texts");
    if (fo) fprintf(fo, "H\tVN:Z:1.0\n");
    for (i = 0; i < G->n_vtx; ++i) {
        asmg_vtx_t *u = &G->vtx[i];
        if (u->del) continue;
        const int64_t l = J.len[i];
        const double cov = J.cov[i];
        u->cov = cov;                                   /* the bit-field keeps the integer part, the text the double */
        u->len = (uint64_t) l;
        if (fo) fprintf(fo, "S\tu%lu\t%.*s\tLN:i:%ld\tKC:i:%ld\tSC:f:%.3f\n", (unsigned long) i, (int) l, J.text[i], (long) l, (long) (int64_t) (l * cov), cov);
        if (save_seq) u->seq = J.text[i]; else free(J.text[i]);
    }
    oatk_tick("cons: S lines");
    /* overlaps are clipped to the unitig lengths, which are all known now */
    cons_run(&J, 1, G->n_arc);
    oatk_tick("cons: arc overlaps");
    for (i = 0; i < G->n_arc; ++i) {
        asmg_arc_t *a = &G->arc[i];
        if (a->del || a->comp) continue;
        int64_t l = J.ovl[i];
        if ((uint64_t) l > G->vtx[a->v >> 1].len) l = (int64_t) G->vtx[a->v >> 1].len;
        if ((uint64_t) l > G->vtx[a->w >> 1].len) l = (int64_t) G->vtx[a->w >> 1].len;
        a->ls = (uint64_t) l;
        arc_vw(G, a->w ^ 1, a->v ^ 1)->ls = (uint64_t) l;
        if (fo) {
            fprintf(fo, "L\tu%lu\t%c\tu%lu\t%c\t%ldM\tEC:i:%u\n", (unsigned long) (a->v >> 1), "+-"[a->v & 1], (unsigned long) (a->w >> 1), "+-"[a->w & 1], (long) l, (unsigned) a->cov);
            fprintf(fo, "L\tu%lu\t%c\tu%lu\t%c\t%ldM\tEC:i:%u\n", (unsigned long) (a->w >> 1), "-+"[a->w & 1], (unsigned long) (a->v >> 1), "-+"[a->v & 1], (long) l, (unsigned) a->cov);
        }
    }
    free(J.text); free(J.len); free(J.cov); free(J.ovl);
}
