/*
 * ref_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C harness around the UNMODIFIED reference sources, compiled where they
 * lie under /root/reference (see oracle/Makefile; output goes to oracle/_ref/).
 * It pulls the reference's syncmer.c into this translation unit with #include
 * so that the file-local functions on the hot path (sr_read_analysis_thread,
 * kmer_hash64, hash64) can be driven from memory buffers without going through
 * a FASTA file. No reference source is copied into this repository.
 *
 * Everything exported here is prefixed ref_ and has a flat C ABI for ctypes.
 */
#include <sys/time.h>
#include "syncmer.c"      /* reference/syncmer.c, unmodified, via -I$(REF) */
#include "syncasm.h"
#include "graph.h"

static double now_s(void)
{
    struct timeval tv;
    gettimeofday(&tv, 0);
    return tv.tv_sec + tv.tv_usec * 1e-6;
}

/* ---- unit functions (SURVEY appendix A.4 vectors come from these) ---- */
uint64_t ref_hash64(uint64_t key, uint64_t mask) { return hash64(key, mask); }
/* the peak finder behind sr_db_stat's peak_hom / peak_het lines (syncmer.c:775-865), quiet */
int ref_analyze_count(int n_cnt, int start_cnt, const int64_t *cnt, int *peak_het) { return ha_analyze_count(n_cnt, start_cnt, cnt, peak_het, 0); }
uint64_t ref_murmur64a(const void *p, uint32_t len, uint64_t seed) { return MurmurHash64A(p, len, seed); }
uint64_t ref_kmer_hash64(uint8_t *hoco_s, uint32_t pos_rev, int k)
{
    uint8_t *key = (uint8_t *) malloc((k - 1) / 4 + 2);
    uint64_t h = kmer_hash64(hoco_s, pos_rev, k, key);
    free(key);
    return h;
}

/* ---- a1-a4: extraction from an in-memory batch (bases + n+1 offsets) ---- */
sr_db_t *ref_extract_mem(const char *bases, const uint64_t *off, uint64_t n_reads, int k, int s)
{
    sr_db_t *db;
    p_data_t dat;
    uint64_t i;
    char nm[32];
    MYMALLOC(db, 1);
    sr_db_init(db, k, s);
    dat.n_reads = 1;
    MYMALLOC(dat.sid, 1);
    MYMALLOC(dat.name, 1);
    MYMALLOC(dat.seq, 1);
    MYMALLOC(dat.len, 1);
    dat.sr_db = db;
    for (i = 0; i < n_reads; ++i) {
        uint64_t l = off[i + 1] - off[i];
        char *sq = (char *) malloc(l + 1);
        memcpy(sq, bases + off[i], l);
        sq[l] = 0;
        snprintf(nm, sizeof(nm), "r%lu", (unsigned long) i);
        dat.sid[0] = i;
        dat.name[0] = strdup(nm);
        dat.seq[0] = sq;           /* freed by the reference thread function */
        dat.len[0] = (int) l;
        sr_read_analysis_thread(&dat);
    }
    free(dat.sid); free(dat.name); free(dat.seq); free(dat.len);
    return db;
}

/* ---- a1: the reference's own file path (kseq + pthread batches) ---- */
sr_db_t *ref_extract_file(const char *path, int k, int s, int n_threads, size_t m_data)
{
    char *files[1];
    sstream_t *rdr;
    sr_db_t *db;
    files[0] = (char *) path;
    rdr = sstream_open(files, 1);
    if (!rdr) return 0;
    MYMALLOC(db, 1);
    sr_db_init(db, k, s);
    sr_read(rdr, db, m_data, n_threads);
    sstream_close(rdr);
    return db;
}

void ref_sr_db_free(sr_db_t *db) { sr_db_destroy(db); }
void ref_scm_db_free(syncmer_db_t *db) { syncmer_db_destroy(db); }
void ref_scg_free(scg_t *g) { scg_destroy(g); }

/* ---- flat views for ctypes ---- */
uint64_t ref_n_reads(sr_db_t *db) { return db->n; }

/* totals[0]=sum hoco_l, [1]=sum n, [2]=sum ceil(hoco_l/4), [3]=#ho_rl==255 */
void ref_totals(sr_db_t *db, uint64_t *totals)
{
    size_t i; uint32_t j;
    totals[0] = totals[1] = totals[2] = totals[3] = 0;
    for (i = 0; i < db->n; ++i) {
        sr_t *r = &db->a[i];
        totals[0] += r->hoco_l;
        totals[1] += r->n;
        totals[2] += (r->hoco_l + 3) / 4;
        for (j = 0; j < r->hoco_l; ++j) totals[3] += r->ho_rl[j] == 255;
    }
}

/* concatenate the per-read arrays, in read order, no padding */
void ref_flatten(sr_db_t *db, uint32_t *hoco_l, uint32_t *n_scm, uint32_t *n_lrl,
        uint8_t *hoco_s, uint8_t *ho_rl, uint32_t *ho_l_rl,
        uint32_t *m_pos, uint64_t *s_mer, uint64_t *k_mer)
{
    size_t i, ps = 0, pr = 0, pl = 0, pm = 0; uint32_t j, c;
    for (i = 0; i < db->n; ++i) {
        sr_t *r = &db->a[i];
        hoco_l[i] = r->hoco_l;
        n_scm[i] = r->n;
        if (r->hoco_l) {
            memcpy(hoco_s + ps, r->hoco_s, (r->hoco_l + 3) / 4); ps += (r->hoco_l + 3) / 4;
            memcpy(ho_rl + pr, r->ho_rl, r->hoco_l); pr += r->hoco_l;
        }
        for (j = 0, c = 0; j < r->hoco_l; ++j) c += r->ho_rl[j] == 255;
        n_lrl[i] = c;
        for (j = 0; j < c; ++j) ho_l_rl[pl++] = r->ho_l_rl[j];
        for (j = 0; j < r->n; ++j, ++pm) {
            m_pos[pm] = r->m_pos[j];
            s_mer[pm] = r->s_mer[j];
            k_mer[pm] = r->k_mer[j];
        }
    }
}

/* n_nucl has no stored length in sr_t; the caller passes the count per read
 * (number of non-ACGTU characters in the raw read) */
void ref_flatten_nnucl(sr_db_t *db, const uint32_t *cnt, uint32_t *out)
{
    size_t i, p = 0; uint32_t j;
    for (i = 0; i < db->n; ++i)
        for (j = 0; j < cnt[i]; ++j) out[p++] = db->a[i].n_nucl[j];
}

/* ---- a5 ---- */
/* out[0..4]: syncmer_n, per_read, avg_dist, smer_avg, kmer_avg ; iout[0..7]: the eight ints */
int ref_stat(sr_db_t *db, int verbose, double *out, int *iout)
{
    FILE *fo = fopen("/dev/null", "w");
    sr_stat_t *st;
    sr_db_stat(db, verbose > 0 ? stderr : fo, verbose);
    fclose(fo);
    st = db->stats;
    if (!st) return 1;
    out[0] = (double) st->syncmer_n; out[1] = st->syncmer_per_read; out[2] = st->syncmer_avg_dist;
    out[3] = st->smer_avg_cnt; out[4] = st->kmer_avg_cnt;
    iout[0] = st->smer_unique; iout[1] = st->smer_singleton; iout[2] = st->smer_peak_hom; iout[3] = st->smer_peak_het;
    iout[4] = st->kmer_unique; iout[5] = st->kmer_singleton; iout[6] = st->kmer_peak_hom; iout[7] = st->kmer_peak_het;
    return 0;
}

/* ---- a6 ---- */
syncmer_db_t *ref_collect(sr_db_t *db) { return collect_syncmer_from_reads(db); }
uint64_t ref_scm_n(syncmer_db_t *s) { return s ? s->n : 0; }
void ref_scm_flatten(syncmer_db_t *s, uint64_t *h, uint64_t *sm, uint32_t *cov, uint64_t *m_pos)
{
    size_t i, p = 0; uint32_t j;
    for (i = 0; i < s->n; ++i) {
        h[i] = s->a[i].h; sm[i] = s->a[i].s; cov[i] = s->a[i].cov;
        for (j = 0; j < s->a[i].cov; ++j) m_pos[p++] = s->a[i].m_pos[j];
    }
}
/* k_mer[] after collect holds id<<1 */
void ref_kmer_ids(sr_db_t *db, uint64_t *k_mer)
{
    size_t i, p = 0; uint32_t j;
    for (i = 0; i < db->n; ++i)
        for (j = 0; j < db->a[i].n; ++j) k_mer[p++] = db->a[i].k_mer[j];
}

/* ---- a7-a9 ---- */
scg_t *ref_make_graph(sr_db_t *db, syncmer_db_t *s, uint32_t min_k_cov, double a)
{
    return make_syncmer_graph(db, s, min_k_cov, a);
}
void ref_unitig(scg_t *g) { process_mergeable_unitigs(g); }
uint64_t ref_graph_n_vtx(scg_t *g) { return g->utg_asmg->n_vtx; }
uint64_t ref_graph_n_arc(scg_t *g) { return g->utg_asmg->n_arc; }
/* arcs: 6 uint64 per arc: v, w, ln, ls, cov|del<<30|comp<<31, link_id */
void ref_graph_arcs(scg_t *g, uint64_t *out)
{
    uint64_t i;
    asmg_t *a = g->utg_asmg;
    for (i = 0; i < a->n_arc; ++i) {
        asmg_arc_t *e = &a->arc[i];
        out[i * 6 + 0] = e->v; out[i * 6 + 1] = e->w; out[i * 6 + 2] = e->ln; out[i * 6 + 3] = e->ls;
        out[i * 6 + 4] = (uint64_t) e->cov | (uint64_t) e->del << 30 | (uint64_t) e->comp << 31;
        out[i * 6 + 5] = e->link_id;
    }
}
/* vertices: n (syncmers), cov|del<<30|circ<<31 ; then the concatenated syncmer lists */
uint64_t ref_graph_vtx_total(scg_t *g)
{
    uint64_t i, t = 0;
    for (i = 0; i < g->utg_asmg->n_vtx; ++i) t += g->utg_asmg->vtx[i].n;
    return t;
}
void ref_graph_vtx(scg_t *g, uint64_t *n, uint64_t *flags, uint64_t *lists)
{
    uint64_t i, j, p = 0;
    asmg_t *a = g->utg_asmg;
    for (i = 0; i < a->n_vtx; ++i) {
        n[i] = a->vtx[i].n;
        flags[i] = (uint64_t) a->vtx[i].cov | (uint64_t) a->vtx[i].del << 30 | (uint64_t) a->vtx[i].circ << 31;
        for (j = 0; j < a->vtx[i].n; ++j) lists[p++] = a->vtx[i].a[j];
    }
}
/* GFA text of the current graph (scg_consensus prints S/L lines) */
int ref_write_gfa(sr_db_t *db, scg_t *g, const char *path)
{
    FILE *fo = fopen(path, "w");
    if (!fo) return 1;
    scg_consensus(db, g, 0, 0, fo);
    fclose(fo);
    return 0;
}

/* the same with the reference's switches: hoco_seq (homopolymer-compressed text), save_seq */
int ref_write_gfa2(sr_db_t *db, scg_t *g, int hoco_seq, int save_seq, const char *path)
{
    FILE *fo = fopen(path, "w");
    if (!fo) return 1;
    scg_consensus(db, g, hoco_seq, save_seq, fo);
    fclose(fo);
    return 0;
}

/* read error correction of the reference on its own structures (run_syncasm.c:126), and what it leaves behind */
void ref_read_ec(sr_db_t *db, scg_t *g, double max_edist, uint32_t err_mer_c, uint32_t max_err_c, uint32_t err_arc_c, double max_arc_f, int n_threads)
{
    read_error_correction(db, g, max_edist, err_mer_c, max_err_c, err_arc_c, max_arc_f, n_threads, 0, 0);
}
void ref_scm_flags(syncmer_db_t *s, uint8_t *del) { size_t i; for (i = 0; i < s->n; ++i) del[i] = s->a[i].del; }
uint64_t ref_scm_total_cov(syncmer_db_t *s) { size_t i; uint64_t t = 0; for (i = 0; i < s->n; ++i) t += s->a[i].cov; return t; }

/* row f3: read -> unitig-graph alignment (alignment.c:596) and the coverage estimators (syncasm.c:1882-2261),
 * run by the reference on its own structures; alignment records flattened for comparison */
scg_ra_v *ref_ra_new(void) { return (scg_ra_v *) calloc(1, sizeof(scg_ra_v)); }
void ref_ra_free(scg_ra_v *ra) { scg_ra_v_destroy(ra); }
void ref_read_alignment(sr_db_t *db, scg_ra_v *ra, scg_t *g, int n_threads, int for_unzip) { scg_read_alignment(db, ra, g, n_threads, for_unzip); }
uint64_t ref_ra_n(scg_ra_v *ra) { return ra->n; }
uint64_t ref_ra_total(scg_ra_v *ra) { size_t i; uint64_t t = 0; for (i = 0; i < ra->n; ++i) t += ra->a[i].n; return t; }
/* per record: sid, n, s; per fragment 5 uint64: uid, u_beg, u_end, s_beg, s_end */
void ref_ra_flatten(scg_ra_v *ra, uint64_t *sid, uint32_t *n, double *s, uint64_t *frg)
{
    size_t i, p = 0;
    uint32_t j;
    for (i = 0; i < ra->n; ++i) {
        sid[i] = ra->a[i].sid; n[i] = ra->a[i].n; s[i] = ra->a[i].s;
        for (j = 0; j < ra->a[i].n; ++j, ++p) {
            ra_frg_t *f = &ra->a[i].a[j];
            frg[5 * p] = f->uid; frg[5 * p + 1] = f->u_beg; frg[5 * p + 2] = f->u_end; frg[5 * p + 3] = f->s_beg; frg[5 * p + 4] = f->s_end;
        }
    }
}
/* the reference's repeat unzipping (run_syncasm.c:207-233) on its own graph: leaves unitigs duplicated along the
 * read paths, i.e. syncmers with several unitig occurrences -- the input that makes alignments ambiguous */
int ref_multiplex_rounds(sr_db_t *db, scg_t *g, int rounds, int n_threads)
{
    scg_ra_v *ra = (scg_ra_v *) calloc(1, sizeof(scg_ra_v));
    int r = 0, updated = 1, total = 0;
    uint32_t max_n_scm = ceil(30000.0 / db->k);
    while (updated != 0 && r < rounds) {
        ++r;
        scg_read_alignment(db, ra, g, n_threads, 1);
        scg_update_utg_cov(g);
        updated = scg_multiplex(g, ra, max_n_scm, 10, .3);
        total += updated;
    }
    scg_ra_v_destroy(ra);
    return total;
}
/* the clean-up passes between .utg.gfa and .utg.final.gfa (graph.c:607, 698, 855; syncasm.c:682, 1090, 1486) */
uint64_t ref_drop_tip(scg_t *g, int32_t tip_cnt, uint64_t tip_len, int protect, int cleanup) { return asmg_drop_tip(g->utg_asmg, tip_cnt, tip_len, protect, cleanup, 0); }
uint64_t ref_weak_crosslink(scg_t *g, double c_thresh, double m_cov, int cleanup) { return asmg_remove_weak_crosslink(g->utg_asmg, c_thresh, m_cov, cleanup, 0); }
uint64_t ref_pop_bubble(scg_t *g, uint64_t radius, uint64_t max_del, int protect_tip, int protect_super, int cleanup) { return asmg_pop_bubble(g->utg_asmg, radius, max_del, protect_tip, protect_super, cleanup, 0); }
void ref_update_utg_cov(scg_t *g) { scg_update_utg_cov(g); }
int ref_multiplex(scg_t *g, scg_ra_v *ra, uint32_t max_n_scm, double min_n_r, double min_d_f) { return scg_multiplex(g, ra, max_n_scm, min_n_r, min_d_f); }
void ref_demultiplex(scg_t *g) { scg_demultiplex(g); }
void ref_ra_utg_coverage(scg_t *g, sr_db_t *db, scg_ra_v *ra) { scg_ra_utg_coverage(g, db, ra, 0); }
void ref_ra_arc_coverage(scg_t *g, sr_db_t *db, scg_ra_v *ra, int refine) { scg_ra_arc_coverage(g, db, ra, refine, 0); }

/* the reference's resumable wavefront edit distance (levdist.c:265 wf_ed_core) in extension mode without traceback,
 * driven the way syncerr.c:444-485 and dfs_search drive it: one diagonal to start with, the query growing by `grow`
 * characters per call */
#include "levdist.h"
void ref_wave_align(char *ts, int tl, char *qs, int ql, int bw, int grow, int *out)
{
    wf_config_t conf;
    wf_diag_t diag;
    int have = 0;
    memset(&conf, 0, sizeof(conf));
    memset(&diag, 0, sizeof(diag));
    conf.ts = ts; conf.tl = tl; conf.qs = qs; conf.is_ext = 1; conf.bw = bw;
    conf.wf_diag = &diag;
    diag.m = 4 * (size_t) (tl + ql + 4);
    diag.a = malloc(diag.m * sizeof(wf_diag1_t));
    diag.n = 1; diag.a[0].d = 0; diag.a[0].k = -1;
    if (grow <= 0) grow = ql ? ql : 1;
    for (;;) {
        have = have + grow < ql ? have + grow : ql;
        conf.ql = have;
        wf_ed_core(&conf);
        if (have >= ql || (conf.t_end > 0 && conf.t_end >= tl) || (bw >= 0 && conf.score > bw)) break;
    }
    out[0] = conf.score; out[1] = conf.t_end; out[2] = conf.q_end;
    free(diag.a);
}

/* the reference's reader alone (sstream_open / sstream_read, the loop of sr_read with its -D rule): every record's
 * name and sequence, flat. Arrays are malloc()ed; the caller frees them. */
int ref_parse_files(char **files, int n_files, size_t mD, char **bases_out, uint64_t **off_out, char **names_out, uint64_t **name_off_out, uint64_t *n_out)
{
    sstream_t *ss = sstream_open(files, n_files);
    size_t nb = 0, mb = 1 << 16, nn = 0, mn = 1 << 12, n = 0, m = 1024, D = 0;
    char *bases = malloc(mb), *names = malloc(mn);
    uint64_t *off = malloc((m + 1) * 8), *noff = malloc((m + 1) * 8);
    int l;
    off[0] = noff[0] = 0;
    if (mD == 0) mD = SIZE_MAX;
    while ((l = sstream_read(ss)) >= 0) {
        const char *nm = ss->s->ks->name.s, *sq = ss->s->ks->seq.s;
        size_t ln = strlen(nm);
        while (nb + l + 1 > mb) { mb *= 2; bases = realloc(bases, mb); }
        while (nn + ln + 1 > mn) { mn *= 2; names = realloc(names, mn); }
        if (n + 2 > m) { m *= 2; off = realloc(off, (m + 1) * 8); noff = realloc(noff, (m + 1) * 8); }
        memcpy(bases + nb, sq, l); nb += l;
        memcpy(names + nn, nm, ln); nn += ln;
        ++n; off[n] = nb; noff[n] = nn;
        D += l;
        if (D >= mD) break;
    }
    sstream_close(ss);
    *bases_out = bases; *off_out = off; *names_out = names; *name_off_out = noff; *n_out = n;
    return 0;
}

/* ---- CPU baseline timing: the reference's own sr_read -> sr_db_stat -> collect,
 *      called in the order run_syncasm.c does. t[0..2] = seconds per stage,
 *      t[3] = raw bases read, t[4] = syncmers, t[5] = distinct k-mers ---- */
int ref_time_extract_count(const char *path, int k, int s, int n_threads, double *t)
{
    double t0, t1, t2, t3;
    uint64_t tot[4];
    sr_db_t *db;
    syncmer_db_t *scm;
    FILE *fo;
    t0 = now_s();
    db = ref_extract_file(path, k, s, n_threads, 0);
    if (!db) return 1;
    t1 = now_s();
    fo = fopen("/dev/null", "w");
    sr_db_stat(db, fo, 0);
    fclose(fo);
    t2 = now_s();
    scm = collect_syncmer_from_reads(db);
    t3 = now_s();
    t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2;
    ref_totals(db, tot);
    t[3] = 0; t[4] = (double) tot[1]; t[5] = scm ? (double) scm->n : 0;
    if (scm) syncmer_db_destroy(scm);
    sr_db_destroy(db);
    return 0;
}
