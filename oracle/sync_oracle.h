/*
 * sync_oracle.h -- CPU restatement of the syncasm hot path. TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * liboracle.so. The product (libsyncgpu.so) never links or calls it.
 *
 * Parity status: PINNED. Every function here is checked (tests/test_oracle.py)
 * against the unmodified reference built in oracle/_ref/libref.so and against
 * the golden vectors under tests/golden/ that were produced by that build
 * (tests/golden/make_golden.py), because the reference ships no test vectors
 * of its own for this path.
 */
#ifndef SYNC_ORACLE_H
#define SYNC_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* one read after a2-a4 (reference syncmer.h:48-70 fields, flat) */
typedef struct {
    uint32_t hoco_l;     /* homopolymer-compressed length */
    uint32_t n;          /* syncmers on the read */
    uint32_t n_lrl;      /* entries in ho_l_rl */
    uint32_t n_n;        /* entries in n_nucl */
    uint8_t *hoco_s;     /* 2-bit packed, 4 bases per byte, first base in bits 7:6 */
    uint8_t *ho_rl;      /* run length - 1, saturating at 255 */
    uint32_t *ho_l_rl;   /* run length - 1 of every run longer than 255 */
    uint32_t *n_nucl;    /* raw index of every non-ACGTU character */
    uint32_t *m_pos;     /* hoco start << 1 | rev */
    uint64_t *s_mer;     /* canonical s-mer << 1 | open/close orientation bit */
    uint64_t *k_mer;     /* MurmurHash64A of the oriented packed k-mer */
} or_read_t;

typedef struct {
    uint64_t n_reads;
    int k, s;
    or_read_t *a;
} or_db_t;

/* syncmer database after a6 (reference syncmer.h:79-89) */
typedef struct {
    uint64_t n;          /* distinct k-mers (= dense ids 0..n-1, hash-sorted) */
    uint64_t n_occ;      /* total occurrences */
    uint64_t *h, *s;     /* per id */
    uint32_t *cov;       /* per id */
    uint64_t *off;       /* per id, n+1: occurrence list start in occ[] */
    uint64_t *occ;       /* sid<<32 | idx<<1 | rev, sorted by (sid, idx) within an id */
    int smer_conflict;   /* 1 if identical k-mers carried different s-mer codes (reference exits) */
} or_scm_t;

uint64_t or_debug_tie_suppressed(void);
uint64_t or_hash64(uint64_t key, uint64_t mask);                       /* syncmer.c:116-126 */
uint64_t or_murmur64a(const void *key, uint32_t len, uint64_t seed);   /* syncmer.c:131-170 */
uint64_t or_kmer_hash(const uint8_t *hoco_s, uint32_t start, int k, int rev); /* syncmer.c:175-226 */

or_db_t *or_extract(const char *bases, const uint64_t *off, uint64_t n_reads, int k, int s);
void or_db_free(or_db_t *db);
void or_totals(const or_db_t *db, uint64_t *totals /* hoco_l, n, hoco bytes, n_lrl, n_n */);
void or_flatten(const or_db_t *db, uint32_t *hoco_l, uint32_t *n_scm, uint32_t *n_lrl, uint32_t *n_n,
        uint8_t *hoco_s, uint8_t *ho_rl, uint32_t *ho_l_rl, uint32_t *n_nucl,
        uint32_t *m_pos, uint64_t *s_mer, uint64_t *k_mer);

/* a6. hash_bits < 64 keeps only the low hash_bits of every k-mer hash before
 * grouping: a test hook that forces hash collisions so that the exact-sequence
 * split (syncmer.c:1270-1393) is exercised. 64 = reference behaviour. On
 * return db->a[].k_mer[] holds id<<1 (syncmer.c:1378). */
or_scm_t *or_collect(or_db_t *db, int hash_bits);
void or_scm_free(or_scm_t *s);

/* a5 (syncmer.c:867-1028). mult-of-mult tables are returned dense for
 * multiplicities < 1001 with the tail summed in [1000] (kh_ctab_cnt with
 * MAX_DEPTH 1000); dout: syncmer_per_read, avg_dist, smer_avg, kmer_avg;
 * iout: smer_unique, smer_singleton, smer_peak_hom, smer_peak_het, then the
 * same four for k-mers. Returns 1 for an empty collection. */
int or_stat(const or_db_t *db, double *dout, int *iout, int64_t *s_cnts, int64_t *k_cnts);
int or_analyze_count(int n_cnt, int start_cnt, const int64_t *cnt, int *peak_het); /* syncmer.c:775-865 */

/* a7 (syncasm.c:203-282): arcs before asmg_finalize, as a sorted list of
 * (v, w, cov, comp). Returns the arc count; out may be NULL to size. */
uint64_t or_arcs(const or_db_t *db, const or_scm_t *scm, uint32_t min_k_cov, double min_a_cov_f, uint64_t *out4);

#ifdef __cplusplus
}
#endif
#endif
