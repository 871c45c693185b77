/*
 * run_syncasm_gpu.c -- the caller of the whole path: reference run_syncasm.c:52-326 `syncasm()`, same signature,
 * same order of steps, same two output files (<out>.utg.gfa, <out>.utg.final.gfa) and the same [M::syncasm] progress
 * lines, over this layer:
 *   reads             sr_read_files              fastx_gpu.c + the device pipeline (rows a1-a4, f4)
 *   statistics        sr_db_stat                 device (a5)
 *   syncmer database  collect_syncmer_from_reads device (a6)
 *   error correction  make_syncmer_graph(0, 0) -> scg_consensus(hoco) -> read_error_correction (a7-a9, f1, f2)
 *   graph, unitigs    make_syncmer_graph -> process_mergeable_unitigs -> scg_consensus -> .utg.gfa
 *   clean-up          asmg_pop_bubble / asmg_remove_weak_crosslink / asmg_drop_tip until nothing changes (cleaning_gpu.c)
 *   unzipping         scg_read_alignment / scg_update_utg_cov / scg_multiplex rounds, scg_demultiplex (f3, unzip_gpu.c)
 *   final coverages   scg_read_alignment -> scg_ra_utg_coverage -> scg_ra_arc_coverage -> scg_consensus -> .utg.final.gfa
 * Errors come back as 1 after an [E::syncasm] line, as in the reference (2: the reference would have left the process from
 * inside -- a file that cannot be opened, identical k-mers with different s-mers -- and its lines are printed); nothing here calls exit().
 */
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <math.h>
#include "graph_gpu.h"
#include "fastx_gpu.h"

static FILE *open_out(const char *prefix, const char *suffix)
{
    char *path = (char *) malloc(strlen(prefix) + strlen(suffix) + 1);
    FILE *fo;
    sprintf(path, "%s%s", prefix, suffix);
    fo = fopen(path, "w");
    if (!fo) fprintf(stderr, "[E::%s] failed to open file '%s' to write: %s\n", __func__, path, strerror(errno));
    free(path);
    return fo;
}

int scg_is_empty(scg_t *scg)
{
    uint64_t i, live = scg->scm_db->n;
    for (i = 0; i < scg->scm_db->n; ++i) live -= scg->scm_db->a[i].del;
    return live == 0;
}

void scg_meta_clean(scg_meta_t *meta)
{
    if (!meta) return;
    scg_destroy(meta->scg);
    if (meta->scm_db) syncmer_db_destroy(meta->scm_db);
    if (meta->sr_db) sr_db_destroy(meta->sr_db);
    scg_ra_v_destroy(meta->ra_db);
    meta->scg = 0; meta->scm_db = 0; meta->sr_db = 0; meta->ra_db = 0;
}

void scg_meta_destroy(scg_meta_t *meta)
{
    if (!meta) return;
    scg_meta_clean(meta);
    free(meta);
}

/* the three passes until a round removes nothing (run_syncasm.c:178-191, 273-281) */
static void clean_graph(scg_t *scg, int with_bubbles, int bubble_size, int tip_size, double weak_cross, int VERBOSE)
{
    uint64_t cleaned = 1;
    while (cleaned) {
        cleaned = 0;
        if (with_bubbles) {
            cleaned += asmg_pop_bubble(scg->utg_asmg, bubble_size, 0, 0, 1, 0, VERBOSE);
            cleaned += asmg_remove_weak_crosslink(scg->utg_asmg, weak_cross, 10, 0, VERBOSE);
        }
        cleaned += asmg_drop_tip(scg->utg_asmg, INT32_MAX, tip_size, 1, 0, VERBOSE);
    }
    process_mergeable_unitigs(scg);
}

/* everything after unitigging (run_syncasm.c:164-303): .utg.gfa, clean-up, unzipping, final coverages, .utg.final.gfa.
 * Host code only; ra_db must point at an empty record vector and receives the final alignments */
int oatk_syncasm_graph_stage(sr_db_t *sr_db, scg_t *scg, scg_ra_v *ra_db, int bubble_size, int tip_size, double weak_cross,
        int do_unzip, int n_threads, char *out, int VERBOSE)
{
    const int k = sr_db->k;
    FILE *fo;
    oatk_tick(0);
    if (!(fo = open_out(out, ".utg.gfa"))) return 1;
    scg_consensus(sr_db, scg, 0, 0, fo);
    fclose(fo);
    oatk_tick("stage: consensus + .utg.gfa");
    if (VERBOSE > 1) scg_subgraph_stat(scg, stderr);

    /* bubbles are haplotypes until the repeats are unzipped: only tips go before that */
    fprintf(stderr, "[M::syncasm] syncmer graph cleanup\n");
    clean_graph(scg, do_unzip <= 0, bubble_size, tip_size, weak_cross, VERBOSE);
    oatk_tick("stage: clean-up");

    if (do_unzip > 0) {
        const uint32_t max_n_scm = (uint32_t) ceil(30000.0 / k);      /* repeats up to ~15 kb: what a HiFi read can span */
        int round = 0, updated = 1;
        fprintf(stderr, "[M::syncasm] assembly graph unzipping\n");
        while (updated != 0 && round < do_unzip) {
            ++round;
            scg_read_alignment(sr_db, ra_db, scg, n_threads, 1);
            scg_update_utg_cov(scg);
            updated = scg_multiplex(scg, ra_db, max_n_scm, 10, .3);
            if (VERBOSE > 0) {
                fprintf(stderr, "[M::syncasm] syncmer graph stats after multiplexing round %d\n", round);
                scg_stat(scg, stderr, 0);
            }
        }
        /* arcs that only reads of another copy support */
        scg_read_alignment(sr_db, ra_db, scg, n_threads, 1);
        scg_ra_arc_coverage(scg, sr_db, ra_db, 0, VERBOSE);
        asmg_remove_weak_crosslink(scg->utg_asmg, weak_cross, 10, 0, VERBOSE);
        oatk_tick("stage: unzip rounds (alignment, multiplex)");

        scg_demultiplex(scg);
        scg_read_alignment(sr_db, ra_db, scg, n_threads, 0);
        scg_ra_utg_coverage(scg, sr_db, ra_db, VERBOSE);
        scg_ra_arc_coverage(scg, sr_db, ra_db, 1, VERBOSE);
        oatk_tick("stage: demultiplex, alignment, coverages");
        scg_consensus(sr_db, scg, 0, 0, 0);                            /* lengths and overlaps for the clean-up */
        clean_graph(scg, 1, bubble_size, tip_size, weak_cross, VERBOSE);
        oatk_tick("stage: consensus + clean-up");
    }

    scg_read_alignment(sr_db, ra_db, scg, n_threads, 0);
    scg_ra_utg_coverage(scg, sr_db, ra_db, VERBOSE);
    scg_ra_arc_coverage(scg, sr_db, ra_db, 1, VERBOSE);
    oatk_tick("stage: final alignment + coverages");

    fprintf(stderr, "[M::syncasm] syncmer graph stats after final processing\n");
    scg_stat(scg, stderr, 0);
    if (!(fo = open_out(out, ".utg.final.gfa"))) return 1;
    scg_consensus(sr_db, scg, 0, 0, fo);
    fclose(fo);
    oatk_tick("stage: consensus + .utg.final.gfa");

    return 0;
}

int syncasm(char **file_in, int n_file, size_t m_data, int k, int s, int bubble_size, int tip_size, int min_k_cov, double min_a_cov_f,
        double weak_cross, int do_ec, int do_unzip, int n_threads, char *out, scg_meta_t *meta, int VERBOSE)
{
    scg_t *scg = 0;
    sr_db_t *sr_db = 0;
    syncmer_db_t *scm_db = 0;
    scg_ra_v *ra_db = 0;
    int ret = 0, rc;

    oatk_set_host_threads(n_threads);                 /* -t bounds the helper pools of this layer too */
    sr_db = (sr_db_t *) malloc(sizeof(sr_db_t));
    sr_db_init(sr_db, k, s);
    {
        /* A caller that takes no structures back (meta == NULL: the syncasm command) never looks at sr_t.ho_rl, whose one
         * consumer, the run-length consensus, is served on the device: the array -- as large as the input -- is not
         * downloaded. OATK_FULL_READS=1 keeps the full records; a caller with meta always gets them. */
        const char *full = getenv("OATK_FULL_READS");
        const int lean = !meta && !(full && atoi(full) > 0);
        const int was = oatk_gpu_keep_run_lengths(lean);
        /* ... and the packed bases (sr_t.hoco_s) with them when the reads are going to be corrected on the device: their
         * other consumer, the consensus, asks the device for the few k-mers it writes out */
        const int was_hs = oatk_gpu_keep_packed_bases(lean && !getenv("OATK_EC_HOST"));
        rc = sr_read_files(sr_db, (const char *const *) file_in, n_file, m_data);
        oatk_gpu_keep_run_lengths(was);
        oatk_gpu_keep_packed_bases(was_hs);
    }
    if (rc == FASTX_E_OPEN) { ret = 2; goto done; }   /* the reference has printed its line and left from inside sstream (sstream.c:46-49) */
    if (rc != 0) {
        fprintf(stderr, "[E::%s] failed to read the input files (%d)\n", __func__, rc);
        ret = 1;
        goto done;
    }
    fprintf(stderr, "[M::%s] collected syncmers from %lu target sequence(s)\n", __func__, (unsigned long) sr_db->n);
    if (sr_db_validate(sr_db)) { ret = 1; goto done; }
    sr_db_stat(sr_db, stderr, VERBOSE);
    if (min_k_cov == 0) {
        if (!sr_db->stats) {                          /* no syncmers, no spectrum to read the threshold from */
            fprintf(stderr, "[E::%s] empty syncmer graph\n", __func__);
            ret = 1;
            goto done;
        }
        min_k_cov = sr_db->stats->kmer_peak_het > 0 ? sr_db->stats->kmer_peak_het * 10 : sr_db->stats->kmer_peak_hom * 10;
        fprintf(stderr, "[M::%s] set minimum kmer coverage as %d\n", __func__, min_k_cov);
    }

    scm_db = collect_syncmer_from_reads(sr_db);
    if (!scm_db) {
        /* an empty database: the reference dereferences NULL here (syncasm.c:205). Identical k-mers with different s-mers: the
         * reference has printed its four lines and left the process with EXIT_FAILURE; the lines are out, 2 tells main */
        if (oatk_collect_conflict()) { ret = 2; goto done; }
        fprintf(stderr, "[E::%s] empty syncmer graph\n", __func__);
        ret = 1;
        goto done;
    }

    if (do_ec) {
        /* the graph of ALL syncmers; the reads are corrected against its consensus in homopolymer-compressed space, which
         * read_error_correction computes itself for the part of the graph that survives its error filter */
        /* ... unless the reads live on the device: then that graph is never built on the host -- the device tallies its
         * arcs, filters them and searches, and only what survives the filter comes down (read_error_correction_device) */
        if (read_error_correction_device(sr_db, scm_db, 0.02, min_k_cov, min_k_cov * 10, min_k_cov, min_a_cov_f, n_threads, VERBOSE) == 0) {
            sr_db_stat(sr_db, stderr, VERBOSE);
        } else {
            scg = make_syncmer_graph(sr_db, scm_db, 0, 0.);
            if (scg) {
                read_error_correction(sr_db, scg, 0.02, min_k_cov, min_k_cov * 10, min_k_cov, min_a_cov_f, n_threads, 0, VERBOSE);
                sr_db_stat(sr_db, stderr, VERBOSE);
                scg_destroy(scg); scg = 0;
            }
        }
    }

    fprintf(stderr, "[M::%s] make syncmer graph\n", __func__);
    scg = make_syncmer_graph(sr_db, scm_db, min_k_cov, min_a_cov_f);
    if (!scg || scg_is_empty(scg)) {
        fprintf(stderr, "[E::%s] empty syncmer graph\n", __func__);
        ret = 1;
        goto done;
    }
    fprintf(stderr, "[M::%s] syncmer graph stats\n", __func__);
    scg_stat(scg, stderr, 0);
    if (VERBOSE > 1) scg_subgraph_stat(scg, stderr);

    fprintf(stderr, "[M::%s] syncmer graph unitigging\n", __func__);
    process_mergeable_unitigs(scg);
    fprintf(stderr, "[M::%s] syncmer graph stats after unitigging\n", __func__);
    scg_stat(scg, stderr, 0);
    ra_db = (scg_ra_v *) calloc(1, sizeof(scg_ra_v));
    ret = oatk_syncasm_graph_stage(sr_db, scg, ra_db, bubble_size, tip_size, weak_cross, do_unzip, n_threads, out, VERBOSE);

done:
    if (meta) {
        scg_meta_clean(meta);
        meta->k = k; meta->s = s;
        meta->scg = scg; meta->scm_db = scm_db; meta->sr_db = sr_db; meta->ra_db = ra_db;
    } else {
        scg_destroy(scg);
        if (scm_db) syncmer_db_destroy(scm_db);
        if (sr_db) sr_db_destroy(sr_db);
        scg_ra_v_destroy(ra_db);
    }
    return ret;
}
