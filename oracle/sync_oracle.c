/*
 * sync_oracle.c -- CPU restatement of the syncasm hot path. TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * the library built from this file. The product never links or calls it.
 *
 * Parity status: PINNED against the unmodified reference compiled into
 * oracle/_ref/libref.so (tests/test_oracle.py) and against tests/golden/.
 *
 * This is a restatement, not a transcription: the reference's extractor
 * (reference syncmer.c:243-421) is a streaming loop over a ring buffer with a
 * running (minimum, slot-of-oldest-minimum) state. Here every decision is
 * written as a stateless predicate over the array m[] of per-position s-mer
 * hashes, which is also the form the CUDA kernels evaluate:
 *
 *   q = k - s + 1                       s-mers per k-mer
 *   m[p]                                hash64 of the canonical s-mer ENDING at
 *                                       hoco position p, or NONE (all ones) when
 *                                       fewer than s valid bases end at p or the
 *                                       s-mer is its own reverse complement
 *   l[p]                                valid bases in a row ending at p (0 on N)
 *   mo(p) = min m[p-q+1 .. p-1]         the window without its newest element
 *   e(p)  = m[p-q]                      the element that leaves the window
 *
 *   CLOSE(p): m[p] != NONE, l[p] >= k, m[p] <= mo(p) and
 *             ( m[p] <= e(p)  or  m[p] < mo(p)  or  m[p-q+1] == m[p] )
 *             -> k-mer starting at p-k+1, strand z of the LAST s-mer
 *   OPEN(p) : evaluated at step p in [k, H] for the k-mer starting at p-k:
 *             e(p) != NONE, e(p) <= mo(p), and (p < H: base p is not N and
 *             l[p] > k ; p == H: l[H-1] >= k) -> strand z of the FIRST s-mer
 *   a start emits iff exactly one of CLOSE / OPEN holds for it (the reference
 *   pushes both and then pops both, syncmer.c:337,393).
 *
 * The third clause of CLOSE is the reference's tie rule (syncmer.c:356-377):
 * after the old minimum expired, a new element that merely ties with the
 * window minimum only counts when the tied older copy sits in the oldest slot.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <assert.h>
#include "sync_oracle.h"

#define NONE UINT64_MAX

/* how often the tie clause of CLOSE rejected a position (tests check that it is exercised) */
static uint64_t g_tie_suppressed;
uint64_t or_debug_tie_suppressed(void) { return g_tie_suppressed; }

/* ---------- scalar pieces ---------- */

/* seq_nt4_table semantics (syncmer.c:47-64): bytes 0..3 are themselves,
 * ACGT/U in either case are 0..3, everything else is ambiguous (4) */
static int base_code(unsigned char ch)
{
    if (ch < 4) return ch;
    switch (ch) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
    }
    return 4;
}

/* invertible integer mix restricted to 2s bits (syncmer.c:116-126) */
uint64_t or_hash64(uint64_t x, uint64_t mask)
{
    x = ((x << 21) - x - 1) & mask;
    x ^= x >> 24;
    x = (x * 265) & mask;
    x ^= x >> 14;
    x = (x * 21) & mask;
    x ^= x >> 28;
    x = (x * 2147483649ULL) & mask;
    return x;
}

/* MurmurHash64A (Appleby), little-endian 8-byte blocks then a <=7-byte tail
 * (syncmer.c:131-170) */
uint64_t or_murmur64a(const void *key, uint32_t len, uint64_t seed)
{
    const uint64_t M = 0xc6a4a7935bd1e995ULL;
    const unsigned char *p = (const unsigned char *) key;
    uint64_t h = seed ^ ((uint64_t) len * M);
    uint32_t nblk = len / 8, i, t;
    for (i = 0; i < nblk; ++i, p += 8) {
        uint64_t w = 0;
        for (t = 0; t < 8; ++t) w |= (uint64_t) p[t] << (8 * t);
        w *= M; w ^= w >> 47; w *= M;
        h = (h ^ w) * M;
    }
    if (len & 7) {
        uint64_t w = 0;
        for (t = 0; t < (len & 7); ++t) w |= (uint64_t) p[t] << (8 * t);
        h = (h ^ w) * M;
    }
    h ^= h >> 47; h *= M; h ^= h >> 47;
    return h;
}

static inline int hoco_base(const uint8_t *hs, uint32_t p)
{
    return (hs[p >> 2] >> ((3 - (p & 3)) * 2)) & 3;
}

/* oriented, left-aligned, zero-padded packed k-mer: what the reference builds
 * by byte reversal + complement table + shifting (syncmer.c:186-211), written
 * here base by base */
static void oriented_kmer(const uint8_t *hs, uint32_t start, int k, int rev, uint8_t *out, int nbytes)
{
    int i;
    memset(out, 0, nbytes);
    for (i = 0; i < k; ++i) {
        int b = rev ? 3 - hoco_base(hs, start + k - 1 - i) : hoco_base(hs, start + i);
        out[i >> 2] |= (uint8_t) (b << ((3 - (i & 3)) * 2));
    }
}

uint64_t or_kmer_hash(const uint8_t *hs, uint32_t start, int k, int rev)
{
    int nb = (k + 3) / 4;
    uint8_t *buf = (uint8_t *) malloc(nb);
    uint64_t h;
    oriented_kmer(hs, start, k, rev, buf, nb);
    h = or_murmur64a(buf, nb, 1234);
    free(buf);
    return h;
}

/* ---------- a2-a4: one read ---------- */

static void extract_read(const char *seq, uint64_t len, int k, int s, or_read_t *r)
{
    uint64_t i, j;
    uint32_t H = 0, nl = 0, nn = 0, p;
    uint8_t *base = (uint8_t *) malloc(len + 1), *isn = (uint8_t *) malloc(len + 1);
    int q = k - s + 1;
    memset(r, 0, sizeof(*r));
    r->ho_rl = (uint8_t *) malloc(len + 1);
    r->ho_l_rl = (uint32_t *) malloc(sizeof(uint32_t) * (len / 256 + 1));
    r->n_nucl = (uint32_t *) malloc(sizeof(uint32_t) * (len + 1));

    /* a2: homopolymer compression of maximal runs of one unambiguous code;
     * an ambiguous character is a run of its own, stored as A with rl 1 */
    for (i = 0; i < len; ) {
        int c = base_code((unsigned char) seq[i]);
        if (c < 4) {
            uint64_t rl;
            for (j = i + 1; j < len && base_code((unsigned char) seq[j]) == c; ++j) {}
            rl = j - i;
            base[H] = (uint8_t) c; isn[H] = 0;
            r->ho_rl[H] = (uint8_t) ((rl > 256 ? 256 : rl) - 1);
            if (rl > 255) r->ho_l_rl[nl++] = (uint32_t) (rl - 1);
            i = j;
        } else {
            base[H] = 0; isn[H] = 1;
            r->ho_rl[H] = 0;
            r->n_nucl[nn++] = (uint32_t) i;
            ++i;
        }
        ++H;
    }
    r->hoco_l = H; r->n_lrl = nl; r->n_n = nn;
    r->hoco_s = (uint8_t *) calloc((H + 3) / 4 + 1, 1);
    for (p = 0; p < H; ++p) r->hoco_s[p >> 2] |= (uint8_t) (base[p] << ((3 - (p & 3)) * 2));

    /* a3: per-position hashes */
    {
        uint64_t *m = (uint64_t *) malloc(sizeof(uint64_t) * (H + 1));
        uint64_t *sv = (uint64_t *) malloc(sizeof(uint64_t) * (H + 1));
        uint32_t *l = (uint32_t *) malloc(sizeof(uint32_t) * (H + 1));
        uint64_t *pre = (uint64_t *) malloc(sizeof(uint64_t) * (H + 1));
        uint64_t *suf = (uint64_t *) malloc(sizeof(uint64_t) * (H + 1));
        uint8_t *fc = (uint8_t *) calloc(H + 2, 1), *fo = (uint8_t *) calloc(H + 2, 1);
        uint64_t mask = (1ULL << (2 * s)) - 1, fw = 0, rv = 0;
        uint32_t run = 0, B = (uint32_t) (q - 1), cap = 0;
        for (p = 0; p < H; ++p) {
            /* rolling both strands; an N contributes an A that is flushed out
             * again before l reaches s */
            fw = ((fw << 2) | base[p]) & mask;
            rv = (rv >> 2) | ((uint64_t) (3 - base[p]) << (2 * (s - 1)));
            run = isn[p] ? 0 : run + 1;
            l[p] = run;
            m[p] = NONE; sv[p] = NONE;
            if (run >= (uint32_t) s && fw != rv) {
                uint64_t z = fw < rv ? 0 : 1, c = z ? rv : fw;
                m[p] = or_hash64(c, mask);
                sv[p] = c << 1 | z;
            }
        }
        /* block prefix / suffix minima with block length q-1: the minimum of
         * any q-1 consecutive elements is min(suf[first], pre[last]) */
        for (p = 0; p < H; ++p)
            pre[p] = (p % B == 0) ? m[p] : (m[p] < pre[p - 1] ? m[p] : pre[p - 1]);
        for (p = H; p-- > 0; )
            suf[p] = (p % B == B - 1 || p == H - 1) ? m[p] : (m[p] < suf[p + 1] ? m[p] : suf[p + 1]);

        for (p = 0; p <= H; ++p) {
            uint64_t mo = NONE, e;
            if (p > 0) {
                /* window m[a..b] = the (at most q-1) elements before p */
                uint32_t b = p - 1, a = p >= (uint32_t) q ? p - q + 1 : 0;
                if (a / B != b / B) mo = suf[a] < pre[b] ? suf[a] : pre[b];
                else if (a % B == 0) mo = pre[b];
                else { uint32_t x; for (x = a; x <= b; ++x) if (m[x] < mo) mo = m[x]; }
            }
            e = p >= (uint32_t) q ? m[p - q] : NONE;
            if (p < H && m[p] != NONE && l[p] >= (uint32_t) k && m[p] <= mo) {
                if (m[p] <= e || m[p] < mo || m[p - q + 1] == m[p]) fc[p] = 1;
                else ++g_tie_suppressed;
            }
            if (p >= (uint32_t) k && e != NONE && e <= mo &&
                    (p < H ? (!isn[p] && l[p] > (uint32_t) k) : l[H - 1] >= (uint32_t) k))
                fo[p] = 1;
        }
        /* emit in start order; CLOSE and OPEN on one start cancel */
        for (p = 0; p + k <= H; ++p) cap += (fc[p + k - 1] ^ fo[p + k]);
        r->m_pos = (uint32_t *) malloc(sizeof(uint32_t) * (cap + 1));
        r->s_mer = (uint64_t *) malloc(sizeof(uint64_t) * (cap + 1));
        r->k_mer = (uint64_t *) malloc(sizeof(uint64_t) * (cap + 1));
        for (p = 0; p + k <= H; ++p) {
            int c = fc[p + k - 1], o = fo[p + k];
            uint64_t code;
            if (c == o) continue;
            code = c ? (sv[p + k - 1] ^ 1) : sv[p + s - 1];
            r->m_pos[r->n] = p << 1 | (uint32_t) ((c ? sv[p + k - 1] : sv[p + s - 1]) & 1);
            r->s_mer[r->n] = code;
            r->k_mer[r->n] = or_kmer_hash(r->hoco_s, p, k, r->m_pos[r->n] & 1);
            ++r->n;
        }
        free(m); free(sv); free(l); free(pre); free(suf); free(fc); free(fo);
    }
    free(base); free(isn);
}

or_db_t *or_extract(const char *bases, const uint64_t *off, uint64_t n_reads, int k, int s)
{
    uint64_t i;
    or_db_t *db;
    if (!(s > 0 && s < 32 && k > s)) return 0;   /* syncmer.c:251 */
    db = (or_db_t *) calloc(1, sizeof(or_db_t));
    db->n_reads = n_reads; db->k = k; db->s = s;
    db->a = (or_read_t *) calloc(n_reads ? n_reads : 1, sizeof(or_read_t));
    for (i = 0; i < n_reads; ++i)
        extract_read(bases + off[i], off[i + 1] - off[i], k, s, &db->a[i]);
    return db;
}

void or_db_free(or_db_t *db)
{
    uint64_t i;
    if (!db) return;
    for (i = 0; i < db->n_reads; ++i) {
        or_read_t *r = &db->a[i];
        free(r->hoco_s); free(r->ho_rl); free(r->ho_l_rl); free(r->n_nucl);
        free(r->m_pos); free(r->s_mer); free(r->k_mer);
    }
    free(db->a); free(db);
}

void or_totals(const or_db_t *db, uint64_t *t)
{
    uint64_t i;
    t[0] = t[1] = t[2] = t[3] = t[4] = 0;
    for (i = 0; i < db->n_reads; ++i) {
        const or_read_t *r = &db->a[i];
        t[0] += r->hoco_l; t[1] += r->n; t[2] += (r->hoco_l + 3) / 4; t[3] += r->n_lrl; t[4] += r->n_n;
    }
}

void or_flatten(const or_db_t *db, uint32_t *hoco_l, uint32_t *n_scm, uint32_t *n_lrl, uint32_t *n_n,
        uint8_t *hoco_s, uint8_t *ho_rl, uint32_t *ho_l_rl, uint32_t *n_nucl,
        uint32_t *m_pos, uint64_t *s_mer, uint64_t *k_mer)
{
    uint64_t i, ps = 0, pr = 0, pl = 0, pn = 0, pm = 0;
    for (i = 0; i < db->n_reads; ++i) {
        const or_read_t *r = &db->a[i];
        hoco_l[i] = r->hoco_l; n_scm[i] = r->n; n_lrl[i] = r->n_lrl; n_n[i] = r->n_n;
        memcpy(hoco_s + ps, r->hoco_s, (r->hoco_l + 3) / 4); ps += (r->hoco_l + 3) / 4;
        memcpy(ho_rl + pr, r->ho_rl, r->hoco_l); pr += r->hoco_l;
        memcpy(ho_l_rl + pl, r->ho_l_rl, 4 * (size_t) r->n_lrl); pl += r->n_lrl;
        memcpy(n_nucl + pn, r->n_nucl, 4 * (size_t) r->n_n); pn += r->n_n;
        memcpy(m_pos + pm, r->m_pos, 4 * (size_t) r->n);
        memcpy(s_mer + pm, r->s_mer, 8 * (size_t) r->n);
        memcpy(k_mer + pm, r->k_mer, 8 * (size_t) r->n);
        pm += r->n;
    }
}

/* ---------- a6: syncmer database ---------- */

typedef struct { uint64_t h, occ; } tup_t;

static int tup_cmp(const void *a, const void *b)
{
    const tup_t *x = (const tup_t *) a, *y = (const tup_t *) b;
    if (x->h != y->h) return x->h < y->h ? -1 : 1;
    if (x->occ != y->occ) return x->occ < y->occ ? -1 : 1;
    return 0;
}

or_scm_t *or_collect(or_db_t *db, int hash_bits)
{
    uint64_t N = 0, i, j, g0, g1, U = 0, hm = hash_bits >= 64 ? ~0ULL : ((1ULL << hash_bits) - 1);
    tup_t *t;
    or_scm_t *S;
    uint32_t *cls;        /* class of each tuple inside its hash group */
    int nb = (db->k + 3) / 4;
    for (i = 0; i < db->n_reads; ++i) N += db->a[i].n;
    if (N == 0) return 0;                                 /* syncmer.c:1414-1417 */
    t = (tup_t *) malloc(sizeof(tup_t) * N);
    cls = (uint32_t *) calloc(N, sizeof(uint32_t));
    for (i = 0, N = 0; i < db->n_reads; ++i)
        for (j = 0; j < db->a[i].n; ++j, ++N) {
            t[N].h = db->a[i].k_mer[j] & hm;
            t[N].occ = i << 32 | j << 1 | (db->a[i].m_pos[j] & 1);
        }
    qsort(t, N, sizeof(tup_t), tup_cmp);                  /* total order: result is unique */

    S = (or_scm_t *) calloc(1, sizeof(or_scm_t));
    S->h = (uint64_t *) malloc(8 * N); S->s = (uint64_t *) malloc(8 * N);
    S->cov = (uint32_t *) calloc(N, 4); S->off = (uint64_t *) malloc(8 * (N + 1));
    S->occ = (uint64_t *) malloc(8 * N);
    S->n_occ = N;

    for (g0 = 0; g0 < N; g0 = g1) {
        uint32_t ncls = 1;
        for (g1 = g0 + 1; g1 < N && t[g1].h == t[g0].h; ++g1) {}
        if (g1 - g0 > 1) {
            /* split the hash group into classes of identical oriented k-mers,
             * numbered in order of first appearance (syncmer.c:1283-1333) */
            uint8_t *seen = (uint8_t *) malloc((size_t) nb * (g1 - g0)), *cur = (uint8_t *) malloc(nb);
            ncls = 0;
            for (i = g0; i < g1; ++i) {
                uint64_t sid = t[i].occ >> 32; uint32_t idx = (uint32_t) t[i].occ >> 1, c;
                const or_read_t *r = &db->a[sid];
                oriented_kmer(r->hoco_s, r->m_pos[idx] >> 1, db->k, (int) (t[i].occ & 1), cur, nb);
                for (c = 0; c < ncls; ++c) if (!memcmp(seen + (size_t) c * nb, cur, nb)) break;
                if (c == ncls) memcpy(seen + (size_t) ncls++ * nb, cur, nb);
                cls[i] = c;
            }
            free(seen); free(cur);
        }
        for (i = 0; i < ncls; ++i) { S->h[U + i] = t[g0].h; S->s[U + i] = NONE; }
        for (i = g0; i < g1; ++i) ++S->cov[U + cls[i]];
        /* occurrence lists: classes laid out one after another, each in tuple order */
        {
            uint64_t o = g0;
            uint32_t c;
            for (c = 0; c < ncls; ++c) { S->off[U + c] = o; o += S->cov[U + c]; S->cov[U + c] = 0; }
        }
        for (i = g0; i < g1; ++i) {
            uint64_t id = U + cls[i], sid = t[i].occ >> 32; uint32_t idx = (uint32_t) t[i].occ >> 1;
            uint64_t sm = db->a[sid].s_mer[idx];
            S->occ[S->off[id] + S->cov[id]++] = t[i].occ;
            if (S->s[id] == NONE) S->s[id] = sm;
            else if (S->s[id] != sm) S->smer_conflict = 1;  /* reference exit(1)s here, :1370-1376 */
            db->a[sid].k_mer[idx] = id << 1;                /* :1378 */
        }
        U += ncls;
    }
    S->off[U] = N;
    S->n = U;
    free(t); free(cls);
    return S;
}

void or_scm_free(or_scm_t *S)
{
    if (!S) return;
    free(S->h); free(S->s); free(S->cov); free(S->off); free(S->occ); free(S);
}

/* ---------- a5: statistics ---------- */

static int u64_cmp(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *) a, y = *(const uint64_t *) b;
    return (x > y) - (x < y);
}

/* hifiasm-style peak finder as the reference carries it (syncmer.c:775-865),
 * without the histogram printing, which does not feed back into the result */
int or_analyze_count(int n, int start_cnt, const int64_t *cnt, int *peak_het)
{
    int i, low, top, left = -1, right = -1;
    int64_t vtop, vleft = -1, vright = -1, mn;
    *peak_het = -1;
    low = cnt[1] > 0 ? 1 : 2;
    if (low < start_cnt) low = start_cnt;
    for (i = low + 1; i < n && cnt[i] <= cnt[i - 1]; ++i) {}
    low = i - 1;
    if (low == n - 1) return -1;
    top = low + 1; vtop = cnt[top];
    for (i = low + 1; i < n; ++i) if (cnt[i] > vtop) vtop = cnt[i], top = i;
    for (i = top - 1; i > low; --i)
        if (cnt[i] >= cnt[i - 1] && cnt[i] >= cnt[i + 1] && cnt[i] > vleft) vleft = cnt[i], left = i;
    if (left > low && left < top) {
        for (i = left + 1, mn = vtop; i < top; ++i) if (cnt[i] < mn) mn = cnt[i];
        if (vleft < vtop * 0.05 || mn > vleft * 0.95) vleft = -1, left = -1;
    }
    for (i = top + 1; i < n - 1; ++i)
        if (cnt[i] >= cnt[i - 1] && cnt[i] >= cnt[i + 1] && cnt[i] > vright) vright = cnt[i], right = i;
    if (right > top) {
        for (i = top + 1, mn = vtop; i < right; ++i) if (cnt[i] < mn) mn = cnt[i];
        if (vright < vtop * 0.05 || mn > vright * 0.95 || right > top * 2.5) vright = -1, right = -1;
    }
    if (right > 0) { *peak_het = top; return right; }
    if (left > 0) *peak_het = left;
    return top;
}

/* The reference tabulates multiplicities in a khashl map (int -> int, syncmer.c:558) and reads
 * the number of singletons as the value stored under key 1 -- or, when no group has multiplicity
 * 1, as the value of whatever slot its scan over the table visited last (kh_ctab_stat,
 * syncmer.c:618-646). That makes the slot order observable, so the table is restated here:
 * power-of-two capacity, Fibonacci bucket of kh_hash_uint32, linear probing, growth at 75 %
 * load with khashl's in-place re-insertion (khashl.h:130-176, 186-211). */
typedef struct { uint32_t key; int val; } cslot_t;
typedef struct { uint32_t bits, count, cap; uint32_t *used; cslot_t *b; } ctab_t;

static uint32_t ct_hash(uint32_t k)
{
    k += ~(k << 15); k ^= k >> 10; k += k << 3; k ^= k >> 6; k += ~(k << 11); k ^= k >> 16;
    return k;
}
static uint32_t ct_bucket(uint32_t key, uint32_t bits) { return (ct_hash(key) * 2654435769U) >> (32 - bits); }
#define CT_USED(u, i) ((u)[(i) >> 5] >> ((i) & 31) & 1U)
#define CT_SET(u, i) ((u)[(i) >> 5] |= 1U << ((i) & 31))
#define CT_CLR(u, i) ((u)[(i) >> 5] &= ~(1U << ((i) & 31)))

static void ct_grow(ctab_t *t)
{
    uint32_t old_n = t->cap, nb = 2, n, j, *nu;
    while ((1U << nb) < old_n + 1) ++nb;
    n = 1U << nb;
    nu = (uint32_t *) calloc(n < 32 ? 1 : n >> 5, 4);
    t->b = (cslot_t *) realloc(t->b, sizeof(cslot_t) * n);
    for (j = 0; j < old_n; ++j) {
        cslot_t cur;
        if (!CT_USED(t->used, j)) continue;
        cur = t->b[j];
        CT_CLR(t->used, j);
        for (;;) {                       /* drop it at its new home, evicting a not-yet-moved element if one sits there */
            uint32_t i = ct_bucket(cur.key, nb);
            while (CT_USED(nu, i)) i = (i + 1) & (n - 1);
            CT_SET(nu, i);
            if (i < old_n && CT_USED(t->used, i)) {
                cslot_t tmp = t->b[i]; t->b[i] = cur; cur = tmp;
                CT_CLR(t->used, i);
            } else { t->b[i] = cur; break; }
        }
    }
    free(t->used);
    t->used = nu; t->bits = nb; t->cap = n;
}

static void ct_add1(ctab_t *t, uint32_t key)
{
    uint32_t i, last, mask;
    if (t->count >= (t->cap >> 1) + (t->cap >> 2)) ct_grow(t);
    mask = t->cap - 1;
    i = last = ct_bucket(key, t->bits);
    while (CT_USED(t->used, i) && t->b[i].key != key) { i = (i + 1) & mask; if (i == last) break; }
    if (!CT_USED(t->used, i)) { t->b[i].key = key; t->b[i].val = 1; CT_SET(t->used, i); ++t->count; }
    else ++t->b[i].val;
}

/* multiplicity-of-multiplicity summary of a sorted key array */
static void mult_table(const uint64_t *a, uint64_t n, int64_t *cnts, int *uniq, int *single, double *avg)
{
    uint64_t i, run = 1, u = 0;
    uint32_t j;
    int last_val = 0, have1 = 0;
    ctab_t t;
    memset(&t, 0, sizeof(t));
    memset(cnts, 0, sizeof(int64_t) * 1001);
    for (i = 1; i <= n; ++i) {
        if (i < n && a[i] == a[i - 1]) { ++run; continue; }
        cnts[run < 1000 ? run : 1000] += 1;
        ct_add1(&t, (uint32_t) run);
        ++u; run = 1;
    }
    for (j = 0; j < t.cap; ++j)
        if (CT_USED(t.used, j)) { last_val = t.b[j].val; if (t.b[j].key == 1) have1 = 1; }
    *single = have1 ? (int) cnts[1] : last_val;
    *uniq = (int) u; *avg = (double) n / (double) u;
    free(t.used); free(t.b);
}

int or_stat(const or_db_t *db, double *dout, int *iout, int64_t *s_cnts, int64_t *k_cnts)
{
    uint64_t N = 0, i, j, ngap = 0;
    double gap_sum = 0;
    uint64_t *ks, *ss;
    for (i = 0; i < db->n_reads; ++i) N += db->a[i].n;
    if (N == 0) return 1;                                  /* syncmer.c:909-912 */
    ks = (uint64_t *) malloc(8 * N); ss = (uint64_t *) malloc(8 * N);
    for (i = 0, N = 0; i < db->n_reads; ++i) {
        const or_read_t *r = &db->a[i];
        int64_t prev = 0x7FFFFFFF;
        for (j = 0; j < r->n; ++j, ++N) {
            int64_t cur = r->m_pos[j] >> 1;
            ks[N] = r->k_mer[j] >> 1;                      /* :896 drops the low bit */
            ss[N] = r->s_mer[j];
            if (prev != 0x7FFFFFFF && cur != 0x7FFFFFFF) { gap_sum += (double) (cur - prev - db->k); ++ngap; }
            prev = cur;
        }
    }
    qsort(ks, N, 8, u64_cmp); qsort(ss, N, 8, u64_cmp);
    dout[0] = (double) N / (double) db->n_reads;
    dout[1] = gap_sum / (double) ngap;                     /* 0/0 = NaN like the reference */
    mult_table(ss, N, s_cnts, &iout[0], &iout[1], &dout[2]);
    mult_table(ks, N, k_cnts, &iout[4], &iout[5], &dout[3]);
    iout[2] = or_analyze_count(1001, 5, s_cnts, &iout[3]);
    iout[6] = or_analyze_count(1001, 5, k_cnts, &iout[7]);
    free(ks); free(ss);
    return 0;
}

/* ---------- a7: arcs ---------- */

typedef struct { uint64_t v, w; } pair_t;
static int pair_cmp(const void *a, const void *b)
{
    const pair_t *x = (const pair_t *) a, *y = (const pair_t *) b;
    if (x->v != y->v) return x->v < y->v ? -1 : 1;
    if (x->w != y->w) return x->w < y->w ? -1 : 1;
    return 0;
}
typedef struct { uint64_t v, w, cov, comp; } arc4_t;
static int arc4_cmp(const void *a, const void *b)
{
    const arc4_t *x = (const arc4_t *) a, *y = (const arc4_t *) b;
    if (x->v != y->v) return x->v < y->v ? -1 : 1;
    if (x->w != y->w) return x->w < y->w ? -1 : 1;
    if (x->comp != y->comp) return x->comp < y->comp ? -1 : 1;
    if (x->cov != y->cov) return x->cov < y->cov ? -1 : 1;
    return 0;
}

uint64_t or_arcs(const or_db_t *db, const or_scm_t *scm, uint32_t min_k_cov, double min_a_cov_f, uint64_t *out4)
{
    uint64_t np = 0, i, j, g0, g1, na = 0;
    pair_t *pr;
    arc4_t *arcs;
    for (i = 0; i < db->n_reads; ++i) if (db->a[i].n) np += db->a[i].n - 1;
    pr = (pair_t *) malloc(sizeof(pair_t) * (np + 1));
    for (i = 0, np = 0; i < db->n_reads; ++i) {
        const or_read_t *r = &db->a[i];
        for (j = 1; j < r->n; ++j, ++np) {
            uint64_t v0 = (r->k_mer[j - 1] >> 1) << 1 | (r->m_pos[j - 1] & 1);
            uint64_t v1 = (r->k_mer[j] >> 1) << 1 | (r->m_pos[j] & 1);
            if (v0 <= v1) { pr[np].v = v0; pr[np].w = v1; }
            else { pr[np].v = v1 ^ 1; pr[np].w = v0 ^ 1; }          /* syncasm.c:256-257 */
        }
    }
    qsort(pr, np, sizeof(pair_t), pair_cmp);
    arcs = (arc4_t *) malloc(sizeof(arc4_t) * (2 * np + 1));
    for (g0 = 0; g0 < np; g0 = g1) {
        uint64_t v0 = pr[g0].v, v1 = pr[g0].w, c, cv0, cv1, mn;
        for (g1 = g0 + 1; g1 < np && pr[g1].v == v0 && pr[g1].w == v1; ++g1) {}
        c = (uint32_t) (g1 - g0);
        cv0 = scm->cov[v0 >> 1]; cv1 = scm->cov[v1 >> 1];
        mn = cv0 < cv1 ? cv0 : cv1;
        if ((double) (uint32_t) c < min_a_cov_f * (double) mn || cv0 < min_k_cov || cv1 < min_k_cov) continue;
        arcs[na].v = v0; arcs[na].w = v1; arcs[na].cov = c; arcs[na].comp = 0; ++na;
        if ((v1 ^ 1) != v0 || (v0 ^ 1) != v1) {
            arcs[na].v = v1 ^ 1; arcs[na].w = v0 ^ 1; arcs[na].cov = c; arcs[na].comp = 1; ++na;
        }
    }
    qsort(arcs, na, sizeof(arc4_t), arc4_cmp);
    if (out4) memcpy(out4, arcs, sizeof(arc4_t) * na);
    free(pr); free(arcs);
    return na;
}
