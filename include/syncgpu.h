/*
 * syncgpu.h -- C ABI of libsyncgpu.so, the B200 (sm_100a) implementation of the
 * oatk/syncasm hot path: closed-syncmer extraction, syncmer counting and the arc
 * tally of the sparse de Bruijn graph.
 *
 * The boundary is what the reference's own functions for this path would bind
 * if its host code stayed in C and called a device library (INTEGRATION.md shows
 * the stub a maintainer would add to reference syncmer.c / syncasm.c):
 *
 *   reference function (file:line)                    entry point here
 *   ------------------------------------------------  -------------------------
 *   sr_read + sr_read_analysis_thread                 sg_batch_set_reads_*,
 *     (syncmer.c:487-556, 243-421), kmer_hash64         sg_extract,
 *     (syncmer.c:175-226)                               sg_extract_sizes / _download
 *   sr_db_stat counting part (syncmer.c:867-987)      sg_stat
 *   collect_syncmer_from_reads + process_kmer_cluster sg_count,
 *     (syncmer.c:1397-1451, 1270-1393)                  sg_count_sizes / _download
 *   make_syncmer_graph arc tally (syncasm.c:236-282)  sg_arcs
 *
 * Plain pointers and sizes only; no C++ or torch types. All functions return 0
 * or a negative SG_E_* code and never call exit(). There is no CPU fallback:
 * without a CUDA device every compute call fails with SG_E_CUDA.
 *
 * Threading: a context and the batches made from it belong to one host thread
 * at a time. One context per GPU; one process per GPU for multi-GPU runs.
 */
#ifndef SYNCGPU_H
#define SYNCGPU_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_OK              0
#define SG_E_CUDA         -1   /* a CUDA runtime call failed (sg_last_error has the text) */
#define SG_E_ARG          -2   /* bad argument (k/s out of range, NULL pointer, state) */
#define SG_E_NOMEM        -3   /* host or device allocation failed */
#define SG_E_LIMIT        -4   /* reference limits exceeded: > 2^32-1 reads, read > 2^31-1 bases (syncmer.h:43-45) */
#define SG_E_KSIZE        -5   /* k too large for the shared-memory window of the scan kernel */
#define SG_E_SMER_CONFLICT -6  /* identical k-mers with different s-mer codes (reference exits, syncmer.c:1370-1376) */
#define SG_E_EMPTY        -7   /* no syncmers in the batch (reference returns NULL, syncmer.c:1414-1417) */
#define SG_E_STATE        -8   /* call order violated (e.g. sg_count before sg_extract) */
#define SG_E_COLLISION    -9   /* multi-GPU only: two different k-mers of different GPUs share a 64-bit hash */
#define SG_E_COMM        -10   /* NCCL missing or a collective failed (sg_comm_*) */

typedef struct sg_ctx sg_ctx;
typedef struct sg_batch sg_batch;

/* ---- context ---- */
int sg_ctx_create(int device, sg_ctx **out);
/* One process per GPU: pin the calling thread to the CPUs next to `device` and prefer its NUMA node for pages touched from
 * now on (call before allocating pinned host buffers). numa_node: the node, -1 unknown, < -1 policy refused; n_cpus: CPUs
 * bound, 0 = affinity left alone. Never fails for lack of sysfs. (The reference has no counterpart: it is host-only.) */
int sg_host_bind_near_device(int device, int flags, int *numa_node, int *n_cpus);   /* flags: 1 = CPUs, 2 = memory */
void sg_ctx_destroy(sg_ctx *ctx);
/* launch everything on this cudaStream_t (NULL = the legacy default stream) */
int sg_ctx_set_stream(sg_ctx *ctx, void *cuda_stream);
int sg_ctx_sync(sg_ctx *ctx);
const char *sg_strerror(int code);
const char *sg_last_error(sg_ctx *ctx);
/* number of kernels launched by this context since creation (bench.py's gpu_launches) */
uint64_t sg_ctx_launches(sg_ctx *ctx);
/* per-stage device time of the most recent call, CUDA events on the context's stream */
enum { SG_T_ENCODE = 0, SG_T_SCAN, SG_T_KMERHASH, SG_T_PLACE, SG_T_SORT, SG_T_GROUP, SG_T_STAT, SG_T_ARCS, SG_T_PACK, SG_T_EC, SG_T_EXCH, SG_T_IDS, SG_T_N };
int sg_ctx_enable_timing(sg_ctx *ctx, int on);
int sg_ctx_timings(sg_ctx *ctx, float *ms /* SG_T_N */, uint32_t *launches /* SG_T_N */);

/* ---- a batch of reads and everything derived from it (device resident) ---- */
int sg_batch_create(sg_ctx *ctx, sg_batch **out);
void sg_batch_destroy(sg_batch *b);
/* bases: all reads back to back, 1 byte per base as in the FASTA/FASTQ record;
 * off: n_reads+1 byte offsets into bases. Host version copies (bases need not be
 * pinned; pinned makes the copy asynchronous). Device version borrows the
 * pointers, which must stay valid until the batch is destroyed or reset; the
 * bases pointer must be 16-byte aligned and the buffer readable up to the next
 * multiple of 16 bytes after total_bases (the encode kernel moves whole 16-byte
 * units, by 128-bit loads before and by the TMA unit's bulk copies now). */
int sg_batch_set_reads_host(sg_batch *b, const char *bases, const uint64_t *off, uint64_t n_reads);
int sg_batch_set_reads_device(sg_batch *b, const void *d_bases, const uint64_t *d_off, uint64_t n_reads, uint64_t total_bases);
/* read ids of this batch start here (sid = sid_base + index); default 0 */
int sg_batch_set_sid_base(sg_batch *b, uint64_t sid_base);

/* a2-a4: homopolymer compression, 2-bit packing, closed-syncmer selection and
 * MurmurHash64A of every selected k-mer. Synchronises once (the syncmer count sizes what follows). */
int sg_extract(sg_batch *b, int k, int s);

typedef struct {
    uint64_t n_reads;
    uint64_t n_syncmers;      /* sum of sr_t.n */
    uint64_t hoco_bases;      /* sum of sr_t.hoco_l */
    uint64_t hoco_s_bytes;    /* size of the packed download buffer (reads 16-byte aligned) */
    uint64_t ho_rl_bytes;     /* size of the run-length download buffer (reads 16-byte aligned) */
    uint64_t n_ambiguous;     /* total entries of all sr_t.n_nucl */
    uint64_t n_long_runs;     /* total entries of all sr_t.ho_l_rl */
} sg_extract_sizes_t;
int sg_extract_sizes(sg_batch *b, sg_extract_sizes_t *out);   /* synchronises */

/* Host destinations; any pointer may be NULL to skip that array.
 * Per read r (reference sr_t, syncmer.h:48-70):
 *   hoco_s  = hoco_s_buf + hoco_s_off[r], ceil(hoco_l[r]/4) bytes
 *   ho_rl   = ho_rl_buf  + ho_rl_off[r],  hoco_l[r] bytes
 *   m_pos/s_mer/k_mer = arrays + scm_off[r], n_scm[r] entries
 *   n_nucl  = the amb_pos entries whose amb_sid == r (sorted by sid, then position)
 *   ho_l_rl = the lrl_val entries whose lrl_sid == r (sorted by sid, then hoco index) */
typedef struct {
    uint32_t *hoco_l;        /* n_reads */
    uint32_t *n_scm;         /* n_reads */
    uint64_t *hoco_s_off;    /* n_reads + 1 */
    uint64_t *ho_rl_off;     /* n_reads + 1 */
    uint64_t *scm_off;       /* n_reads + 1 */
    uint8_t *hoco_s_buf;     /* hoco_s_bytes */
    uint8_t *ho_rl_buf;      /* ho_rl_bytes */
    uint32_t *m_pos;         /* n_syncmers: hoco start << 1 | rev */
    uint64_t *s_mer;         /* n_syncmers: canonical s-mer << 1 | orientation */
    uint64_t *k_mer;         /* n_syncmers: k-mer hash (after sg_count: id << 1) */
    uint32_t *amb_sid, *amb_pos;             /* n_ambiguous */
    uint32_t *lrl_sid, *lrl_idx, *lrl_val;   /* n_long_runs */
} sg_extract_out_t;
int sg_extract_download(sg_batch *b, const sg_extract_out_t *out);

/* f1, the bulk of scg_syncmer_consensus (reference syncasm.c:946-998) on the device. For each of n_req syncmers the caller lists the
 * occurrences that count (occ[occ_off[i] .. occ_off[i+1]), each read << 32 | hoco start << 1 | strand -- i.e. sr_t.sid and
 * sr_t.m_pos[] of the copies read error correction left alone); sums[i * k + j] = sum over strand-0 copies of (run length - 1) at
 * start + j plus sum over strand-1 copies at start + k - 1 - j, long runs resolved through the side list. Needs the run lengths
 * on the device: a batch after sg_extract, or a pipe's master batch with sg_pipe_keep_run_lengths. Synchronises. */
int sg_runlen_sums(sg_batch *b, uint64_t n_req, const uint64_t *occ_off, const uint64_t *occ, uint64_t *sums /* host, n_req * k */);
int sg_runlen_resident(sg_batch *b);                          /* 1 when sg_runlen_sums can be served */
/* hoco bases of n_req stretches of `len` positions, one code 0..3 per byte (n_req x len), read from the device-resident
 * packed bases: refs[i] = read << 32 | first hoco position. What get_kmer_seq (syncmer.c:1105-1125) reads from sr_t.hoco_s,
 * for read databases whose packed bases stayed on the device (sg_pipe_keep_packed_bases). */
int sg_kmer_codes(sg_batch *b, uint64_t n_req, const uint64_t *refs, int len, uint8_t *codes);

/* ---- f2: the per-read pass of read error correction on the device ----
 * Replaces the kt_for over reads inside read_error_correction (reference syncerr.c:342-612 per read, :144-288 dfs_search,
 * levdist.c:156-225, 265-310 wf_ed_core in extension mode). The caller has run the error filter (syncerr.c:679-757) and
 * passes what is left of the all-syncmer graph; the reads, their syncmer lists (k_mer = id << 1 | corrected, m_pos) and
 * hoco_s are those resident in the batch (after sg_count, or after sg_batch_set_lists_host). Results come back as the
 * rewritten list of every read that has at least one anchor (out_n[r] = 0xffffffff: read r keeps its list). */
typedef struct {
    uint64_t n_syncmers;            /* entries of del */
    const uint8_t *del;             /* per syncmer id: flagged deleted (syncmer_t.del after the filter) */
    uint64_t n_arcs;                /* live arcs of the filtered graph, in the graph's own order (sorted by v, then as asmg_finalize left them) */
    const uint64_t *arc_v, *arc_w;  /* oriented vertices, id << 1 | strand */
    const uint32_t *arc_ls;         /* overlap of v and w in hoco bases (asmg_arc_t.ls) */
    const uint64_t *arc_txt;        /* where the k hoco bases of w's syncmer are read from: batch-local read << 32 | hoco start << 1 |
                                       strand of that occurrence (the first one no correction touched, syncasm.c:911-921); ~0: none left, the text is all N */
} sg_ec_graph_t;
typedef struct {
    int64_t stats[11];              /* tail blocks, their 4 outcomes, middle blocks, their 4 outcomes, blocks shorter than 10 bases */
    uint64_t n_out;                 /* entries of out_k / out_p */
    uint64_t *out_off;              /* per read: first entry of its rewritten list */
    uint32_t *out_n;                /* per read: entries, or 0xffffffff when the read keeps its list */
    uint64_t *out_k;                /* k_mer of the rewritten lists */
    uint32_t *out_p;                /* m_pos of the rewritten lists */
    uint64_t n_overflow_reads;      /* reads whose search outgrew the ordinary arena and ran again in the worst-case one */
} sg_ec_result_t;
int sg_ec_correct(sg_batch *b, const sg_ec_graph_t *g, double max_edist, sg_ec_result_t *res);
void sg_ec_result_free(sg_ec_result_t *res);
/* The error filter of read error correction on the device (reference find_error_syncmers, syncerr.c:679-757): runs
 * sg_arcs(b, 0, 0) -- the arcs of the all-syncmer graph, which never have to become a host graph -- flags the suspect
 * syncmers (coverage below err_mer_c, or below max_err_c with a side whose arcs are all unreliable) and returns the arcs
 * both of whose ends survive, in the list's (v, w, comp) order. del_prev: syncmer_t.del on entry, or NULL for none. */
typedef struct {
    uint64_t n_syncmers;            /* entries of err */
    uint8_t *err;                   /* 1: suspect (malloc'ed) */
    uint64_t n_arcs_all;            /* arcs before the filter */
    uint64_t n_live;                /* arcs left */
    uint64_t *arcs4;                /* v, w, cov, comp per live arc (malloc'ed) */
} sg_ec_filter_out_t;
int sg_ec_filter(sg_batch *b, const uint8_t *del_prev, uint32_t err_mer_c, uint32_t max_err_c, uint32_t err_arc_c, double max_arc_f,
        sg_ec_filter_out_t *out);
void sg_ec_filter_free(sg_ec_filter_out_t *out);
/* f1: the votes of calc_syncmer_overlap (reference syncasm.c:477-582) for n arcs between single syncmers (v, w, cov, comp
 * records as sg_ec_filter returns them), over the occurrence lists of the last sg_count: dist[i] = the most frequent
 * difference of the two syncmers' start positions over the reads that carry them as neighbours (0 when there is none);
 * flag[i] = 1 when two values share the highest count (the reference breaks the tie in the slot order of its hash table: the
 * caller decides those on the host), 2 for a complement arc (no vote: it takes its mirror's value), else 0. */
int sg_arc_votes(sg_batch *b, uint64_t n, const uint64_t *arcs4, int32_t *dist, uint8_t *flag);
/* test hook: vertices per path and stacked wavefront entries of the first-pass search arena (default 256 / 16384) */
int sg_debug_set_ec_arena(uint32_t path, uint32_t stash);

/* a5: the counting part of sr_db_stat. Dense multiplicity-of-multiplicity tables
 * for distinct s-mer codes and distinct (k_mer >> 1) keys, multiplicities >= 1000
 * summed in [1000] (reference kh_ctab_cnt with MAX_DEPTH 1000, syncmer.c:639-659),
 * and the gap sum of consecutive syncmers on a read. Peak finding stays on the host. */
typedef struct {
    uint64_t n_syncmers;
    uint64_t n_gaps;          /* pairs of consecutive syncmers on one read */
    int64_t gap_sum;          /* sum of (p1 - p0 - k) over those pairs */
    uint64_t smer_unique, smer_singleton;
    uint64_t kmer_unique, kmer_singleton;
    int64_t smer_cnts[1001];
    int64_t kmer_cnts[1001];
} sg_stat_t;
int sg_stat(sg_batch *b, sg_stat_t *out);                     /* synchronises */
/* after sg_stat: the multiplicity of every distinct key in ascending key order (which = 0: k_mer >> 1, 1: s-mer codes),
 * i.e. in the order the reference's sr_db_stat feeds its multiplicity tables (syncmer.c:916-926, 967-977). Only that
 * order is not in sg_stat_t; the host layer needs it to print the reference's singleton figure when no key occurs
 * once (see oatk_b200/host/syncmer_gpu.c). *host_out is malloc'ed (NULL when empty), the caller frees it. */
int sg_stat_multiplicities(sg_batch *b, int which, uint32_t **host_out, uint64_t *n);

/* a6: syncmer database. Sorts the (hash, sid, idx, rev) tuples, groups them by
 * hash, verifies with an exact sequence comparison that a group holds one
 * k-mer (splitting it like process_kmer_cluster when it does not), assigns
 * dense ids in hash order and rewrites k_mer[] to id << 1. */
int sg_count(sg_batch *b);
/* After sg_count returned SG_E_SMER_CONFLICT: what the reference prints before it exits (syncmer.c:1371-1374), for the first
 * class in hash order that holds two s-mer codes: out[0] the k-mer hash, out[1] / out[2] the s-mer code and the read of the
 * class's first tuple, out[3] / out[4] the code and the read of the first tuple that disagrees. SG_E_STATE otherwise. */
int sg_count_conflict(sg_batch *b, uint64_t out[5]);
/* How sg_count decides that two tuples of equal hash are the same k-mer. 0 (default): equal hash and
 * equal 64-bit fingerprint (a second, independent hash of the same packed k-mer, computed by
 * sg_extract); a group with differing fingerprints is then split by exact sequence comparison like
 * process_kmer_cluster. 1: the packed k-mers themselves are compared for every tuple, as the
 * reference does (syncmer.c:1283-1333). Same results unless two different k-mers agree in all 128 bits. */
int sg_batch_set_exact_verify(sg_batch *b, int on);
typedef struct {
    uint64_t n_syncmers;      /* occurrences */
    uint64_t n_unique;        /* syncmer_db_t.n */
    uint64_t n_hash_collisions; /* hash groups that had to be split */
} sg_count_sizes_t;
int sg_count_sizes(sg_batch *b, sg_count_sizes_t *out);       /* synchronises */
typedef struct {
    uint64_t *h;              /* n_unique: syncmer_t.h */
    uint64_t *s;              /* n_unique: syncmer_t.s */
    uint32_t *cov;            /* n_unique: syncmer_t.cov */
    uint64_t *occ_off;        /* n_unique + 1 */
    uint64_t *occ;            /* n_syncmers: concatenated syncmer_t.m_pos lists (sid<<32 | idx<<1 | rev) */
    uint64_t *k_mer_id;       /* n_syncmers, read order: id << 1 (sr_t.k_mer after collect) */
} sg_count_out_t;
int sg_count_download(sg_batch *b, const sg_count_out_t *out);

/* a7: arc tally of make_syncmer_graph. Counts canonical (v0,v1) neighbour pairs
 * in a warp-cooperative open-addressing table, applies the coverage filters and
 * returns arcs and their complements sorted by (v, w, comp): 4 uint64 per arc
 * = v, w, cov, comp. Vertex ids are syncmer ids (before asmg_cleanup renumbers). */
int sg_arcs(sg_batch *b, uint32_t min_k_cov, double min_a_cov_f, uint64_t *n_arcs);  /* synchronises */
int sg_arcs_download(sg_batch *b, uint64_t *arcs4);

/* ---- host buffers at PCIe rate: the chunked copy / compute / copy pipeline ----
 * sr_read's role (syncmer.c:487-556) for inputs that live in host memory: the reads are cut into
 * chunks of `chunk_reads`, which flow through n_slots (stream, batch) pairs driven by one host
 * thread each, so uploads, kernels and downloads overlap. Per-read results are written to `out`
 * in read order exactly as sg_extract_download would (offset arrays are global); the syncmer
 * tuples and the packed reads are appended to a device-resident master batch on which sg_stat,
 * sg_count, sg_count_download, sg_arcs run afterwards. `bases` and the big arrays of `out` should
 * be pinned (cudaHostAlloc / cudaHostRegister) or the copies serialise. */
typedef struct sg_pipe sg_pipe;
typedef struct {
    uint64_t max_syncmers;    /* capacity of out->m_pos / s_mer / k_mer */
    uint64_t hoco_s_bytes;    /* capacity of out->hoco_s_buf (<= total_bases/4 + 16 per read always suffices) */
    uint64_t ho_rl_bytes;     /* capacity of out->ho_rl_buf (<= total_bases + 16 per read always suffices) */
    uint64_t max_ambiguous, max_long_runs;
} sg_pipe_caps_t;
int sg_pipe_create(int device, int n_slots, sg_pipe **out);
void sg_pipe_destroy(sg_pipe *p);
int sg_pipe_run_host(sg_pipe *p, const char *bases, const uint64_t *off, uint64_t n_reads, int k, int s,
        uint64_t chunk_reads, const sg_extract_out_t *out, const sg_pipe_caps_t *caps, sg_extract_sizes_t *sizes);
/* The same pipeline, handing every chunk to the caller instead of filling flat arrays: `fn` is called once per
 * chunk of reads [r0, r0 + n_reads) with chunk-local arrays (offsets start at 0, amb_sid / lrl_sid count reads
 * from r0) that live in pinned staging memory and are valid only during the call. Calls for different chunks
 * may run concurrently on different threads; a non-zero return value stops the run and is returned. This is
 * what the reference-API layer uses to build its per-read malloc blocks while later chunks are still in flight.
 * `bases` may be ordinary pageable memory in both forms (it is staged through pinned buffers by the slot
 * threads); pinned input goes straight to the copy engine. */
typedef int (*sg_pipe_chunk_fn)(void *user, uint64_t r0, uint64_t n_reads, const sg_extract_out_t *chunk, const sg_extract_sizes_t *sizes);
int sg_pipe_run_host_cb(sg_pipe *p, const char *bases, const uint64_t *off, uint64_t n_reads, int k, int s,
        uint64_t chunk_reads, sg_pipe_chunk_fn fn, void *user, sg_extract_sizes_t *sizes);
sg_batch *sg_pipe_master(sg_pipe *p);
sg_ctx *sg_pipe_ctx(sg_pipe *p);
const char *sg_pipe_last_error(sg_pipe *p);
uint64_t sg_pipe_launches(sg_pipe *p);
/* 1: the run lengths (ho_rl, 1 byte per hoco base -- as large as the input) stay on the device in the master batch instead of
 * being downloaded; sg_runlen_sums then serves the one consumer they have, the run-length consensus. sg_pipe_run_host does the
 * same on its own when the caller passes no ho_rl buffer. */
int sg_pipe_keep_run_lengths(sg_pipe *p, int on);
/* 1: the packed bases (hoco_s, a quarter byte per hoco base) are not downloaded either; sg_kmer_codes serves the consensus.
 * sg_pipe_run_host does the same on its own when the caller passes no hoco_s buffer. */
int sg_pipe_keep_packed_bases(sg_pipe *p, int on);
/* Callback form: the master batch has room for `factor` times the expected number of syncmers (2 per window of k - s + 1
 * bases; default 16, never more than one per base). Low-complexity input (short tandem repeats tie at every position) can
 * exceed that: the run then ends with SG_E_NOMEM and sg_pipe_syncmer_overflow() = 1, and the caller repeats it with a
 * larger factor (sr_read_mem does). */
int sg_pipe_set_capacity_factor(sg_pipe *p, unsigned factor);
int sg_pipe_syncmer_overflow(sg_pipe *p);
/* multi-GPU: global index of this pipe's first read (sid of read i = sid_base + i); default 0 */
int sg_pipe_set_sid_base(sg_pipe *p, uint64_t sid_base);

/* device pointers of the batch's result arrays (for device-side consumers such as the multi-GPU
 * exchange); valid until the next call that recomputes them. n = number of elements. */
enum { SG_BUF_KEY = 0,      /* uint64 k-mer hash, read order */
       SG_BUF_OCC,          /* uint64 sid<<32 | idx<<1 | rev, read order */
       SG_BUF_SMER,         /* uint64 s-mer code, read order */
       SG_BUF_MPOS,         /* uint32 m_pos, read order */
       SG_BUF_KID,          /* uint64 id<<1 per tuple of the counted set (read order, or adopted order) */
       SG_BUF_SORTED_OCC,   /* uint64 occurrences in hash-sorted order = concatenated syncmer_t.m_pos */
       SG_BUF_SCM_H,        /* uint64 per distinct k-mer */
       SG_BUF_SCM_COV,      /* uint32 per distinct k-mer */
       SG_BUF_ADOPTED_OCC   /* uint64 occ of the adopted tuples, adopted order */ };
int sg_batch_buffer(sg_batch *b, int which, void **d_ptr, uint64_t *n);

/* test hook: keep only the low `bits` bits of every k-mer hash when grouping, which forces hash
 * collisions so that the exact-sequence split of process_kmer_cluster is exercised. 64 = off. */
int sg_debug_set_hash_bits(sg_batch *b, int bits);
/* tests only: the sort of sg_count/sg_stat runs radix passes over hash bits [low_bits, 64) and repairs the
 * rare runs that differ below (default 24; 0 = sort on all 64 bits). sg_debug_sort_info reports how many
 * out-of-order pairs the last sort repaired and whether it fell back to the full sort. */
int sg_debug_set_sort_low_bits(sg_batch *b, int low_bits);
/* test hook: hash bits ordered by the packed sort (8, 16, 24 or 32; default 32) -- fewer bits leave longer runs to the repair pass */
int sg_debug_set_pack_bits(sg_batch *b, int bits);
int sg_debug_sort_info(sg_batch *b, uint64_t *repairs, int *fell_back);
/* tests only: how many reads of the last sg_extract were finished by scan_exact_kernel (full-hash window
 * minimum; taken by reads on which identical s-mers keep tying for the minimum: tandem repeats) */
int sg_debug_scan_info(sg_batch *b, uint64_t *deferred_reads);

/* After read error correction has rewritten the reads' syncmer lists on the host (reference syncerr.c:592-606, 769-817):
 * replace the batch's per-syncmer arrays (read order: k_mer = id << 1 | corrected, m_pos, s_mer; scm_off = n_reads + 1
 * offsets) and the per-k-mer coverages, so that the second sr_db_stat (keys = ids, corrected entries skipped in the gap
 * sum, syncmer.c:896-902) and the arc tally of the final graph run on the corrected lists. sg_count is not valid on
 * such a batch (the ids exist). */
int sg_batch_set_lists_host(sg_batch *b, uint64_t n_reads, const uint64_t *scm_off, const uint64_t *k_mer, const uint32_t *m_pos,
        const uint64_t *s_mer, const uint32_t *cov, uint64_t n_unique);

/* multi-GPU statistics: after the tuple exchange every rank holds a share of the occurrences of an s-mer code,
 * so the s-mer part of sg_stat is per rank. sg_smer_counts_pack exposes the distinct codes of this rank's tuples
 * with their local counts (2 x uint64 per code, device memory owned by the batch, valid until the next sg_stat /
 * merge); after an all-gather, sg_smer_counts_merge adds the pairs of all ranks up and overwrites the smer_*
 * fields of `out` with the global tables (reference sr_db_stat, syncmer.c:916-926). k-mer tables and gap sums of
 * the ranks simply add up. */
int sg_smer_counts_pack(sg_batch *b, void **d_pairs, uint64_t *n);
int sg_smer_counts_merge(sg_batch *b, const void *d_pairs, uint64_t n, sg_stat_t *out);

/* ---- multi-GPU exchange (one process per GPU; the transport is the caller's
 * collective, e.g. NCCL all-to-all; see oatk_b200/dist.py) ---- */
/* partition this batch's tuples by hash range into n_parts buckets: counts[p] tuples for part p,
 * laid out contiguously in an internal device buffer of 4 x uint64 per tuple (hash, occ, s_mer, fingerprint) */
int sg_tuples_partition(sg_batch *b, int n_parts, uint64_t *counts /* host, n_parts */, void **d_tuples /* device ptr out */);
/* replace this batch's tuple set by tuples received from the peers (device pointer, n tuples of 4 x uint64) */
int sg_tuples_adopt(sg_batch *b, const void *d_tuples, uint64_t n);
/* after sg_count on adopted tuples: (occ, (id + id_base) << 1) pairs in adopted order, 2 x uint64 each,
 * to be sent back to the ranks the tuples came from (same split sizes, reversed) */
int sg_ids_pack(sg_batch *b, uint64_t id_base, void **d_pairs, uint64_t *n);
/* the pairs received back for this rank's own reads: fills k_mer[] (read order) with id << 1 */
int sg_ids_scatter(sg_batch *b, const void *d_pairs, uint64_t n);

/* ---- the same exchange in C over NCCL (csrc/sg_comm.cu; NCCL is dlopen'ed at the first call) ----
 * One sg_comm per GPU. Either one process per GPU (sg_comm_unique_id on rank 0, the 128 bytes handed to the
 * other ranks by the launcher, sg_comm_init_rank everywhere) or one process with one thread per GPU
 * (sg_comm_init_all, then each thread drives its own context). All data calls are collective: every rank makes
 * the same calls in the same order; they run on the context's stream. */
typedef struct sg_comm sg_comm;
int sg_comm_unique_id(void *id128 /* 128 bytes out */);
int sg_comm_init_rank(sg_ctx *ctx, int world, int rank, const void *id128, sg_comm **out);
int sg_comm_init_all(sg_ctx **ctxs, int n, sg_comm **out /* n */);
void sg_comm_destroy(sg_comm *c);
int sg_comm_rank(sg_comm *c);
int sg_comm_world(sg_comm *c);
uint64_t sg_comm_bytes_sent(sg_comm *c);                    /* payload this rank put on the wire so far */
/* between sg_extract and sg_stat / sg_count: sg_tuples_partition -> all-to-all-v -> sg_tuples_adopt. Counts are
 * exchanged on the device; the host reads the world x world count matrix once. */
int sg_comm_exchange_tuples(sg_comm *c, sg_batch *b);
/* after sg_stat on every rank: *st (this rank's result) becomes the table over ALL reads (reference sr_db_stat on
 * the whole input): k-mer tables, gap sums and totals are summed, s-mer codes merged as (code, count) pairs */
int sg_comm_global_stat(sg_comm *c, sg_batch *b, sg_stat_t *st);
/* after sg_count on every rank: global ids (local rank in the hash range + distinct k-mers of the lower ranges,
 * summed on the device) go back to the ranks that hold the reads; k_mer[] of sg_extract_download then holds them */
int sg_comm_return_ids(sg_comm *c, sg_batch *b, uint64_t *id_base, uint64_t *n_unique_total);
/* a7 over all ranks (reference syncasm.c:236-282), after sg_comm_return_ids: local pair tally -> all-to-all-v keyed by
 * the canonical pair -> filter with the coverages of all ranges -> the whole arc list, (v, w, comp) order, on `root`
 * (sg_arcs_download there; *n_arcs is 0 on the other ranks) */
int sg_comm_arcs(sg_comm *c, sg_batch *b, uint32_t min_k_cov, double min_a_cov_f, int root, uint64_t *n_arcs);

#ifdef __cplusplus
}
#endif
#endif
