/*
 * syncmer_gpu.h -- host side of the drop-in: the reference's syncmer.h interface for
 * the hot path (reference syncmer.h:39-144), served by libsyncgpu.so.
 *
 * The four structs below are byte-compatible with the reference's (same member
 * order, types and bit-fields; tests/test_host_layer.py hands them to the
 * reference's own code to prove it), and the functions keep the reference's
 * names, argument meaning and ownership rules:
 *   - every per-read array is its own malloc block, NULL when empty
 *     (sr_destroy frees them one by one, syncmer.c:1047-1058; read error
 *     correction reallocs k_mer/m_pos/s_mer in place, syncerr.c:604-608)
 *   - syncmer_t.m_pos is one malloc block per syncmer (syncmer.c:1359)
 *   - sr_db_stat prints the reference's nine [M::sr_db_stat] lines
 * The reader comes in three forms: sr_read() over an sstream_t as in the reference
 * (sstream_gpu.c), sr_read_files() for file names (fastx_gpu.h) and sr_read_mem()
 * for records already in memory.
 */
#ifndef SYNCMER_GPU_H
#define SYNCMER_GPU_H
#include <stdio.h>
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __SYNCMER_H__          /* inside the reference tree its own header provides these */
#define uint128_t __uint128_t
#define MAX_RD_NUM 0xFFFFFFFFULL
#define MAX_RD_LEN 0x7FFFFFFFULL
#define MAX_RD_SCM 0x7FFFFFFFULL

typedef struct {
    uint64_t sid;
    char *sname;
    uint32_t hoco_l;
    uint8_t *hoco_s;
    uint8_t *ho_rl;
    uint32_t *ho_l_rl;
    uint32_t *n_nucl;
    uint32_t n;
    uint32_t *m_pos;
    uint64_t *s_mer;
    uint64_t *k_mer;
} sr_t;

typedef struct {
    uint64_t syncmer_n;
    double syncmer_per_read, syncmer_avg_dist, smer_avg_cnt, kmer_avg_cnt;
    int smer_unique, smer_singleton, smer_peak_hom, smer_peak_het;
    int kmer_unique, kmer_singleton, kmer_peak_hom, kmer_peak_het;
} sr_stat_t;

typedef struct {
    size_t n, m;
    sr_t *a;
    int k, s;
    sr_stat_t *stats;
} sr_db_t;

typedef struct {
    uint64_t h, s;
    uint32_t cov:31, del:1;
    uint64_t *m_pos;
} syncmer_t;

typedef struct {
    size_t n, m;
    syncmer_t *a;
    uint16_t *c;
    uint64_t *h;
} syncmer_db_t;
#endif

#ifndef KSTRING_H              /* klib's kstring.h provides this inside the reference tree */
typedef struct { size_t l, m; char *s; } kstring_t;
#endif

#ifndef __SSTREAM_H__           /* sstream.h:46-51: same members, same order; `s` is this layer's own state (sstream_gpu.c) */
typedef struct {
    uint64_t n_seq;             /* records delivered so far */
    char **files;
    int n_files, n;             /* n: index of the file being read */
    void *s;
} sstream_t;
sstream_t *sstream_open(char **files, int n_files);
void sstream_close(sstream_t *stream);
#endif
/* syncmer.h:120, syncmer.c:487-556: every record of the stream through the device pipeline into sr_db (read order
 * kept, sid == index); m_data = 0: no limit, else stop after the record that reaches it, with the reference's message */
void sr_read(sstream_t *s_stream, sr_db_t *sr_db, size_t m_data, int n_threads);

extern const unsigned char seq_nt4_table[256];
extern const char char_nt4_table[4];

void sr_db_init(sr_db_t *sr_db, int k, int s);
/* reads = n_reads records back to back in `bases`, off[n_reads+1]; names may be NULL.
 * Returns 0, or a negative SG_E_* code after printing an [E::sr_read_mem] line. */
int sr_read_mem(sr_db_t *sr_db, const char *bases, const uint64_t *off, char **names, uint64_t n_reads);
int sr_db_validate(sr_db_t *sr_db);
void sr_db_stat(sr_db_t *sr_db, FILE *fo, int more);
syncmer_db_t *collect_syncmer_from_reads(sr_db_t *sr_db);
void sr_destroy(sr_t *sr);
void sr_db_clean(sr_db_t *sr_db);
void sr_db_destroy(sr_db_t *sr_db);
void syncmer_db_init(syncmer_db_t *scm_db);
void syncmer_db_clean(syncmer_db_t *scm_db);
void syncmer_db_destroy(syncmer_db_t *scm_db);
void get_kmer_seq(uint8_t *hoco_s, uint32_t pos, int l, uint32_t rev, uint8_t *kmer_s);
void get_kmer_dna_seq(uint8_t *hoco_s, uint32_t pos, int l, uint32_t rev, char *dna_seq);
void print_hoco_seq(sr_t *sr, FILE *fo);
/* report_gpu.c: the debugging printers (syncmer.c:1164-1216). k = s-mer length, w = k-mer length, as the reference names them */
void get_hoco_seq(sr_t *sr, kstring_t *s);
void print_syncmer_on_seq(sr_t *sr, uint32_t n, int k, int w, FILE *fo);
void print_all_syncmers_on_seq(sr_t *sr, int k, int w, FILE *fo);
void print_aligned_syncmers_on_seq(sr_t *sr, int w, uint32_t beg, uint32_t end, FILE *fo);

/* arcs of make_syncmer_graph (syncasm.c:236-282) for the database just collected:
 * 4 uint64 per arc (v, w, cov, comp), sorted by (v, w, comp); caller frees. */
int syncmer_graph_arcs(sr_db_t *sr_db, syncmer_db_t *scm_db, uint32_t min_k_cov, double min_a_cov_f, uint64_t **arcs4, uint64_t *n_arcs);

/* which GPU the layer uses (default 0), and its release */
/* after read_error_correction (called by it): refresh the device-resident batch from the corrected host lists */
int oatk_gpu_update_lists(sr_db_t *sr_db, syncmer_db_t *scm_db);
void oatk_parallel_for(uint64_t n, void (*fn)(uint64_t lo, uint64_t hi, void *arg), void *arg);
void oatk_tick(const char *what);      /* OATK_TIMING=1: stage times on stderr */
int oatk_collect_conflict(void);                /* 1: the last collect_syncmer_from_reads returned NULL because identical k-mers had different s-mers (the reference exits there) */
void oatk_set_host_threads(int n);              /* the helper thread pools of this layer use at most n threads (0: the cores, up to 16); syncasm() passes -t */
long oatk_host_threads(void);
int oatk_gpu_set_device(int device);
/* run lengths stay on the device (sr_t.ho_rl == NULL) for the read databases made from now on; returns the previous setting */
int oatk_gpu_keep_run_lengths(int on);
int oatk_gpu_run_lengths_on_device(sr_db_t *sr_db);
int oatk_gpu_keep_packed_bases(int on);          /* sr_t.hoco_s stays on the device as well (NULL on the host) */
int oatk_gpu_bases_on_device(sr_db_t *sr_db);
int oatk_gpu_kmer_codes(sr_db_t *sr_db, uint64_t n, const uint64_t *refs, int len, uint8_t *codes);
int oatk_gpu_runlen_sums(sr_db_t *sr_db, uint64_t n_req, const uint64_t *occ_off, const uint64_t *occ, uint64_t *sums);
/* the per-read pass of read error correction on the device (syncerr_gpu.c builds the arguments) */
int oatk_gpu_ec_available(sr_db_t *sr_db);
int oatk_gpu_ec_correct(sr_db_t *sr_db, const void *graph, double max_edist, void *result);
int oatk_gpu_arc_votes(sr_db_t *sr_db, uint64_t n, const uint64_t *arcs4, int32_t *dist, uint8_t *flag);
int oatk_gpu_ec_filter(sr_db_t *sr_db, uint32_t err_mer_c, uint32_t max_err_c, uint32_t err_arc_c, double max_arc_f, void *result);
void oatk_cons_cache_drop(const sr_db_t *db);   /* consensus_gpu.c: forget the run-length sums kept for db (NULL: whatever is kept) */
void oatk_gpu_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif
