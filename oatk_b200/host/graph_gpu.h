/*
 * graph_gpu.h -- host side of rows a7-a9: the reference's graph.h / syncasm.h interface
 * for building the syncmer graph and merging unitigs (reference graph.h:39-96,
 * syncasm.h:51-64, 94-97). Structs are byte-compatible with the reference's.
 *
 * The arc tally runs on the GPU (sg_arcs); what is left is small (a few 10^4 vertices
 * after the coverage filter) and order-sensitive, so it stays on the host:
 *   make_syncmer_graph        syncasm.c:203-299   vertices + arcs -> asmg_finalize -> index
 *   asmg_finalize             graph.c:250-263     cleanup, sort, index, symmetry repair, link ids
 *   asmg_unitigging           graph.c:905-1105    three ordered walks, singletons, arc remap, list expansion
 *   process_mergeable_unitigs syncasm.c:1048-1061
 *   scg_consensus             syncasm.c:716-823   unitig sequences, overlaps, GFA (consensus_gpu.c)
 */
#ifndef GRAPH_GPU_H
#define GRAPH_GPU_H
#include "syncmer_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __GRAPH_H__
typedef struct {
    uint64_t v, w;
    uint64_t ln, ls;
    uint32_t cov:30, del:1, comp:1;
    uint64_t link_id;
} asmg_arc_t;

typedef struct {
    uint64_t n;
    uint64_t *a;
    char *seq;
    uint64_t len;
    uint32_t cov:30, del:1, circ:1;
} asmg_vtx_t;

typedef struct {
    uint64_t n_vtx, m_vtx;
    asmg_vtx_t *vtx;
    uint64_t n_arc, m_arc;
    asmg_arc_t *arc;
    uint64_t *idx_p;
    uint64_t *idx_n;
} asmg_t;
#endif

#ifndef __SYNCASM_H__
typedef struct {
    syncmer_db_t *scm_db;      /* borrowed */
    asmg_t *utg_asmg;
    uint128_t *scm_u;          /* scm_id[49] | scm_rev[1] | utg_id[42] | utg_pos[36], sorted */
    uint128_t **idx_u;         /* n_scm + 1 pointers into scm_u */
} scg_t;

/* one read against the unitig graph (syncasm.h:62-77): a chain of fragments, each a stretch of one oriented unitig */
typedef struct {
    uint64_t uid;              /* unitig << 1 | strand */
    uint64_t u_beg, u_end;     /* first / last syncmer on the oriented unitig (inclusive) */
    uint32_t s_beg, s_end;     /* first / last syncmer on the read (inclusive) */
} ra_frg_t;

typedef struct {
    uint64_t sid;
    uint32_t n;
    ra_frg_t *a;               /* its own malloc block */
    double s;                  /* best chain score + 1 / (number of equally good alignments of this read) */
} scg_ra_t;

typedef struct { size_t n, m; scg_ra_t *a; } scg_ra_v;

/* what syncasm() hands to a caller that wants to go on with the graph (syncasm.h:79-85) */
typedef struct {
    int k, s;
    scg_t *scg;
    syncmer_db_t *scm_db;
    sr_db_t *sr_db;
    scg_ra_v *ra_db;
} scg_meta_t;
#endif

int oatk_ec_last_run(uint64_t *overflow_reads);
/* read error correction without the all-syncmer graph on the host (filter, search on the device; syncerr_gpu.c): 0 when it ran,
 * anything else when the caller has to take make_syncmer_graph + read_error_correction */
int read_error_correction_device(sr_db_t *sr_db, syncmer_db_t *scm_db, double max_edist, uint32_t err_mer_c, uint32_t max_err_c,
        uint32_t err_arc_c, double max_arc_f, int n_threads, int verbose);
void oatk_syncmer_arc_overlaps(sr_db_t *sr_db, syncmer_db_t *scm_db, uint64_t n, const uint64_t *arcs4, uint32_t *ls);   /* 1: the last read_error_correction searched on the device */
void asmg_destroy(asmg_t *g);
/* graphutil_gpu.c / cleaning_gpu.c: the queries of graph.h:72-96 that callers beyond syncasm() use */
int asmg_arc_is_sorted(asmg_t *g);
uint64_t *asmg_vtx_list(asmg_t *g, uint64_t *_n);
void asmg_print(asmg_t *g, FILE *fo, int no_seq);
uint32_t *asmg_uext_arc_group(asmg_t *g, uint32_t *n);
uint32_t *asmg_subgraph(asmg_t *g, uint32_t *seeds, uint32_t n, uint32_t step, uint64_t dist, uint32_t *_nv, int modify_graph);
int asmg_tarjans_scc(asmg_t *g, int *scc);
int asmg_path_exists(asmg_t *g, uint32_t source, uint32_t sink, uint32_t step, uint64_t dist, uint32_t *_step, uint64_t *_dist);
void asmg_arc_sort(asmg_t *g);
void asmg_arc_index(asmg_t *g);
void asmg_shrink_link_id(asmg_t *g);
void asmg_finalize(asmg_t *g, int do_cleanup);
asmg_t *asmg_unitigging(asmg_t *g);

scg_t *make_syncmer_graph(sr_db_t *sr_db, syncmer_db_t *scm_db, uint32_t min_k_cov, double min_a_cov_f);
void process_mergeable_unitigs(scg_t *g);
void scg_destroy(scg_t *g);
void scg_stat(scg_t *scg, FILE *fo, uint64_t *stats);
/* row f1 (consensus_gpu.c): unitig sequences, arc overlaps and the GFA text (reference syncasm.c:716-823).
 * hoco_seq: write homopolymer-compressed bases; save_seq: keep each unitig's sequence in vtx[i].seq;
 * fo may be NULL (lengths, coverages and overlaps are still filled in) */
void scg_consensus(sr_db_t *sr_db, scg_t *scg, int hoco_seq, int save_seq, FILE *fo);
/* row f2 (syncerr_gpu.c): read error correction on the all-syncmer graph (reference syncerr.c:679-757, 819-943).
 * g must be the graph of ALL syncmers (make_syncmer_graph(sr_db, scm_db, 0, 0.)), either after scg_consensus(sr_db, g, 1, 1, 0)
 * as in the reference, or without it: then the consensus is computed inside, only for what the error filter leaves;
 * rewrites the reads' syncmer lists and rebuilds the coverages / occurrence lists of g->scm_db */
int64_t find_error_syncmers(scg_t *g, uint32_t err_mer_c, uint32_t max_err_c, uint32_t err_arc_c, double max_arc_f, int del_err);
void read_error_correction(sr_db_t *sr_db, scg_t *g, double max_edist, uint32_t err_mer_c, uint32_t max_err_c,
        uint32_t err_arc_c, double max_arc_f, int n_threads, FILE *fo, int verbose);

/* row f3 (alignment_gpu.c, coverage_gpu.c): reads against the unitig graph and the coverage estimates drawn from
 * the alignments (reference alignment.c:596-684, syncasm.c:1882-2261, graph.c:117-127, 237-248).
 * ra_v is replaced by the new records, ordered by read and, within a read, in the reference's enumeration order;
 * for_unzip re-aligns only reads that spanned more than two unitigs before and keeps a result only if it scores
 * at least what the old one did */
void scg_read_alignment(sr_db_t *sr_db, scg_ra_v *ra_v, scg_t *g, int n_threads, int for_unzip);
void scg_ra_v_destroy(scg_ra_v *ra_v);
void scg_ra_utg_coverage(scg_t *g, sr_db_t *sr_db, scg_ra_v *ra_v, int verbose);
void scg_ra_arc_coverage(scg_t *g, sr_db_t *sr_db, scg_ra_v *ra_v, int refine, int verbose);
void scg_refine_arc_coverage(scg_t *g, int verbose);
uint64_t asmg_max_link_id(asmg_t *g);
void asmg_arc_fix_cov(asmg_t *g);
/* the whole command (run_syncasm_gpu.c; reference run_syncasm.c:52-326): files -> <out>.utg.gfa, <out>.utg.final.gfa.
 * min_k_cov 0 = ten times the k-mer coverage peak; do_unzip = rounds of repeat unzipping (0: none, bubbles are popped
 * straight away); meta, if given, receives the structures instead of their being freed. Returns 0, or 1 after an
 * [E::syncasm] line */
int syncasm(char **file_in, int n_file, size_t m_data, int k, int s, int bubble_size, int tip_size, int min_k_cov, double min_a_cov_f,
        double weak_cross, int do_ec, int do_unzip, int n_threads, char *out, scg_meta_t *meta, int VERBOSE);
int oatk_syncasm_graph_stage(sr_db_t *sr_db, scg_t *scg, scg_ra_v *ra_db, int bubble_size, int tip_size, double weak_cross,
        int do_unzip, int n_threads, char *out, int VERBOSE);
int scg_is_empty(scg_t *scg);
/* report_gpu.c: GFA of the graph as it stands (no consensus run), unitig syncmer lists, per-component statistics, alignment
 * records, and the arc coverage recount from the reads (syncasm.c:825, 857, 423, 309; alignment.c:686-708) */
void scg_print(scg_t *g, FILE *fo, int no_seq);
void scg_print_unitig_syncmer_list(scg_t *g, FILE *fo);
void scg_subgraph_stat(scg_t *scg, FILE *fo);
void scg_arc_coverage(scg_t *scg, sr_db_t *sr_db);
void scg_ra_print(scg_ra_t *ra, FILE *fo);
void scg_rv_print(scg_ra_v *rv, FILE *fo);
void scg_meta_clean(scg_meta_t *meta);
void scg_meta_destroy(scg_meta_t *meta);
/* repeat unzipping (unzip_gpu.c; reference syncasm.c:682, 1090, 1486): scg_multiplex returns the number of
 * (arc in, arc out) pairings it dropped, 0 = graph untouched */
void scg_update_utg_cov(scg_t *scg);
int scg_multiplex(scg_t *g, scg_ra_v *ra_v, uint32_t max_n_scm, double min_n_r, double min_d_f);
void scg_demultiplex(scg_t *g);
/* clean-up of the unitig graph between .utg.gfa and .utg.final.gfa (cleaning_gpu.c; reference graph.c:607, 698, 855):
 * return the number of tips / links / bubbles (| short tips << 32) removed; do_cleanup re-finalizes the graph */
uint64_t asmg_drop_tip(asmg_t *g, int32_t tip_cnt, uint64_t tip_len, int protect_super_tip, int do_cleanup, int VERBOSE);
uint64_t asmg_remove_weak_crosslink(asmg_t *g, double c_thresh, double m_cov, int do_cleanup, int VERBOSE);
uint64_t asmg_pop_bubble(asmg_t *g, uint64_t radius, uint64_t max_del, int protect_tip, int protect_super_bubble, int do_cleanup, int VERBOSE);

#ifdef __cplusplus
}
#endif
#endif
