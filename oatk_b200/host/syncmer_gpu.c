/*
 * syncmer_gpu.c -- the reference's syncmer.h functions for the hot path, implemented
 * over the C ABI of libsyncgpu.so (include/syncgpu.h). Host code only: it moves
 * flat device results into the reference's data structures and prints what the
 * reference prints. No sequence arithmetic happens here.
 */
#include <stdlib.h>
#include <unistd.h>
#include <time.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include "syncgpu.h"
#include "syncmer_gpu.h"

const unsigned char seq_nt4_table[256] = {
#define R16(x) x, x, x, x, x, x, x, x, x, x, x, x, x, x, x, x
    0, 1, 2, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, R16(4), R16(4), R16(4),
    4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    R16(4), R16(4), R16(4), R16(4), R16(4), R16(4), R16(4), R16(4)
#undef R16
};
const char char_nt4_table[4] = {'A', 'C', 'G', 'T'};

/* one device context for the process, one live batch per sr_db_t */
static sg_ctx *g_ctx;
static int g_device;
typedef struct reg_s { sr_db_t *db; sg_batch *b; sg_pipe *pipe; struct reg_s *next; } reg_t;
static reg_t *g_reg;
static sg_pipe *g_spare;          /* a released pipeline (device buffers, pinned staging, streams) waits here for the next sr_read_mem */
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

int oatk_gpu_set_device(int device) { g_device = device; return 0; }

static sg_batch *batch_of(sr_db_t *db, int create)
{
    reg_t *r;
    sg_batch *b = 0;
    pthread_mutex_lock(&g_lock);
    for (r = g_reg; r; r = r->next) if (r->db == db) { b = r->b; break; }
    if (!b && create) {
        if (!g_ctx && sg_ctx_create(g_device, &g_ctx) != SG_OK) {
            fprintf(stderr, "[E::%s] no usable CUDA device %d (libsyncgpu has no CPU path)\n", __func__, g_device);
        } else if (sg_batch_create(g_ctx, &b) == SG_OK) {
            r = (reg_t *) malloc(sizeof(reg_t));
            r->db = db; r->b = b; r->pipe = 0; r->next = g_reg; g_reg = r;
        }
    }
    pthread_mutex_unlock(&g_lock);
    return b;
}

/* whose error text to print */
static sg_ctx *ctx_of(sr_db_t *db)
{
    reg_t *r;
    sg_ctx *c = g_ctx;
    pthread_mutex_lock(&g_lock);
    for (r = g_reg; r; r = r->next) if (r->db == db && r->pipe) { c = sg_pipe_ctx(r->pipe); break; }
    pthread_mutex_unlock(&g_lock);
    return c;
}

static void batch_drop(sr_db_t *db)
{
    reg_t **pp, *r;
    pthread_mutex_lock(&g_lock);
    for (pp = &g_reg; (r = *pp); pp = &r->next)
        if (r->db == db) {
            *pp = r->next;
            if (r->pipe) { if (g_spare) sg_pipe_destroy(g_spare); g_spare = r->pipe; }
            else sg_batch_destroy(r->b);
            free(r);
            break;
        }
    pthread_mutex_unlock(&g_lock);
}

void oatk_gpu_shutdown(void)
{
    pthread_mutex_lock(&g_lock);
    while (g_reg) { reg_t *r = g_reg; g_reg = r->next; if (r->pipe) sg_pipe_destroy(r->pipe); else sg_batch_destroy(r->b); free(r); }
    if (g_spare) { sg_pipe_destroy(g_spare); g_spare = 0; }
    if (g_ctx) { sg_ctx_destroy(g_ctx); g_ctx = 0; }
    pthread_mutex_unlock(&g_lock);
}

void sr_db_init(sr_db_t *sr_db, int k, int s)
{
    if (!sr_db) return;
    sr_db->n = sr_db->m = 0;
    sr_db->a = 0;
    sr_db->k = k;
    sr_db->s = s;
    sr_db->stats = 0;
}

static void *dup_block(const void *src, size_t bytes)
{
    void *p;
    if (!bytes) return 0;                 /* kvec semantics: an empty array is NULL */
    p = malloc(bytes);
    memcpy(p, src, bytes);
    return p;
}

/* the pipeline of libsyncgpu hands over one chunk of reads at a time (chunk-local arrays in pinned staging);
 * this turns it into the reference's per-read malloc blocks. Runs on the pipeline's slot threads, one chunk per
 * call, disjoint read ranges. */
typedef struct { sr_db_t *db; char **names; } fill_t;

static int fill_chunk(void *user, uint64_t r0, uint64_t nr, const sg_extract_out_t *c, const sg_extract_sizes_t *z)
{
    fill_t *f = (fill_t *) user;
    uint64_t i, ia = 0, il = 0;
    for (i = 0; i < nr; ++i) {
        sr_t *r = &f->db->a[r0 + i];
        const uint64_t a0 = ia, l0 = il;
        char nm[32];
        r->sid = r0 + i;                                        /* asserted by the reference, syncmer.c:1407 */
        if (f->names && f->names[r0 + i]) r->sname = strdup(f->names[r0 + i]);
        else { snprintf(nm, sizeof(nm), "r%lu", (unsigned long) (r0 + i)); r->sname = strdup(nm); }
        r->hoco_l = c->hoco_l[i];
        r->hoco_s = (uint8_t *) dup_block(c->hoco_s_buf + c->hoco_s_off[i], (c->hoco_l[i] + 3) / 4);
        r->ho_rl = (uint8_t *) dup_block(c->ho_rl_buf + c->ho_rl_off[i], c->hoco_l[i]);
        while (ia < z->n_ambiguous && c->amb_sid[ia] == i) ++ia;
        r->n_nucl = (uint32_t *) dup_block(c->amb_pos + a0, 4 * (ia - a0));
        while (il < z->n_long_runs && c->lrl_sid[il] == i) ++il;
        r->ho_l_rl = (uint32_t *) dup_block(c->lrl_val + l0, 4 * (il - l0));
        r->n = c->n_scm[i];
        r->m_pos = (uint32_t *) dup_block(c->m_pos + c->scm_off[i], 4 * (size_t) r->n);
        r->s_mer = (uint64_t *) dup_block(c->s_mer + c->scm_off[i], 8 * (size_t) r->n);
        r->k_mer = (uint64_t *) dup_block(c->k_mer + c->scm_off[i], 8 * (size_t) r->n);
    }
    return 0;
}

#define SR_READ_SLOTS 6
#define SR_READ_CHUNK 4096

int sr_read_mem(sr_db_t *sr_db, const char *bases, const uint64_t *off, char **names, uint64_t n_reads)
{
    sg_pipe *pipe = 0;
    sg_extract_sizes_t z;
    fill_t f;
    reg_t *r;
    int rc, k = sr_db->k, s = sr_db->s;

    sr_db_clean(sr_db);
    sr_db_init(sr_db, k, s);
    pthread_mutex_lock(&g_lock);
    pipe = g_spare; g_spare = 0;
    pthread_mutex_unlock(&g_lock);
    int slots = getenv("OATK_SR_SLOTS") ? atoi(getenv("OATK_SR_SLOTS")) : SR_READ_SLOTS;      /* tuning knob; the pipeline takes 1..8 */
    if (slots < 1) slots = 1;
    if (slots > 8) slots = 8;
    if (!pipe && sg_pipe_create(g_device, slots, &pipe) != SG_OK) {
        fprintf(stderr, "[E::%s] no usable CUDA device %d (libsyncgpu has no CPU path)\n", __func__, g_device);
        return SG_E_CUDA;
    }
    /* sr_db_stat / collect / the arc tally work on the pipeline's device-resident master batch */
    pthread_mutex_lock(&g_lock);
    r = (reg_t *) malloc(sizeof(reg_t));
    r->db = sr_db; r->b = sg_pipe_master(pipe); r->pipe = pipe; r->next = g_reg; g_reg = r;
    pthread_mutex_unlock(&g_lock);
    sr_db->a = (sr_t *) calloc(n_reads ? n_reads : 1, sizeof(sr_t));
    sr_db->n = sr_db->m = n_reads;
    f.db = sr_db; f.names = names;
    rc = sg_pipe_run_host_cb(pipe, bases, off, n_reads, k, s, SR_READ_CHUNK, fill_chunk, &f, &z);
    if (rc != SG_OK) fprintf(stderr, "[E::%s] %s: %s\n", __func__, sg_strerror(rc), sg_pipe_last_error(pipe));
    return rc;
}

int sr_db_validate(sr_db_t *sr_db)
{
    size_t i;
    if (sr_db->n > MAX_RD_NUM) {
        fprintf(stderr, "[E::%s] read number exceeds the limit %llu\n", __func__, MAX_RD_NUM);
        return 1;
    }
    for (i = 0; i < sr_db->n; ++i)
        if (sr_db->a[i].n > MAX_RD_SCM) {
            fprintf(stderr, "[E::%s] syncmer number (%u) on read exceeds the limit %llu: %s\n", __func__,
                    sr_db->a[i].n, MAX_RD_SCM, sr_db->a[i].sname);
            return 2;
        }
    return 0;
}

/* peak finder the reference borrows from hifiasm (syncmer.c:775-865): lowest point from the left,
 * highest peak after it, then a smaller peak on either side that is separated by a real valley */
static int find_peaks(int n, int start_cnt, const int64_t *cnt, int *peak_het)
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

void sr_db_stat(sr_db_t *sr_db, FILE *fo, int more)
{
    sg_batch *b = batch_of(sr_db, 0);
    sg_stat_t st;
    sr_stat_t *s;
    int rc;
    (void) more;
    if (!sr_db->stats) sr_db->stats = (sr_stat_t *) calloc(1, sizeof(sr_stat_t));
    s = sr_db->stats;
    if (!b) { fprintf(stderr, "[E::%s] the read database was not produced by sr_read_mem\n", __func__); return; }
    rc = sg_stat(b, &st);
    if (rc == SG_E_EMPTY) { fprintf(fo, "[M::%s] empty syncmer collection\n", __func__); return; }   /* syncmer.c:909-912 */
    if (rc != SG_OK) { fprintf(stderr, "[E::%s] %s: %s\n", __func__, sg_strerror(rc), sg_last_error(ctx_of(sr_db))); return; }
    s->syncmer_n = st.n_syncmers;
    s->syncmer_per_read = (double) st.n_syncmers / (double) sr_db->n;
    s->syncmer_avg_dist = (double) st.gap_sum / (double) st.n_gaps;
    s->smer_unique = (int) st.smer_unique; s->smer_singleton = (int) st.smer_singleton;
    s->smer_avg_cnt = (double) st.n_syncmers / (double) st.smer_unique;
    s->kmer_unique = (int) st.kmer_unique; s->kmer_singleton = (int) st.kmer_singleton;
    s->kmer_avg_cnt = (double) st.n_syncmers / (double) st.kmer_unique;
    s->smer_peak_hom = find_peaks(1001, 5, st.smer_cnts, &s->smer_peak_het);
    s->kmer_peak_hom = find_peaks(1001, 5, st.kmer_cnts, &s->kmer_peak_het);
    fprintf(fo, "[M::%s] number syncmers collected: %lu\n", __func__, (unsigned long) s->syncmer_n);
    fprintf(fo, "[M::%s] number syncmers per read: %.3f\n", __func__, s->syncmer_per_read);
    fprintf(fo, "[M::%s] average kmer space: %.3f\n", __func__, s->syncmer_avg_dist);
    fprintf(fo, "[M::%s] number uniqe smer: %d; singletons: %d (%.3f%%)\n", __func__, s->smer_unique, s->smer_singleton,
            (double) s->smer_singleton * 100 / s->smer_unique);
    fprintf(fo, "[M::%s] average smer count: %.3f\n", __func__, s->smer_avg_cnt);
    fprintf(fo, "[M::%s] smer peak_hom: %d; peak_het: %d\n", __func__, s->smer_peak_hom, s->smer_peak_het);
    fprintf(fo, "[M::%s] number uniqe kmer: %d; singletons: %d (%.3f%%)\n", __func__, s->kmer_unique, s->kmer_singleton,
            (double) s->kmer_singleton * 100 / s->kmer_unique);
    fprintf(fo, "[M::%s] average kmer count: %.3f\n", __func__, s->kmer_avg_cnt);
    fprintf(fo, "[M::%s] kmer peak_hom: %d; peak_het: %d\n", __func__, s->kmer_peak_hom, s->kmer_peak_het);
}

syncmer_db_t *collect_syncmer_from_reads(sr_db_t *sr_db)
{
    sg_batch *b = batch_of(sr_db, 0);
    sg_count_sizes_t z;
    sg_count_out_t o;
    syncmer_db_t *db;
    uint64_t *h, *s, *off, *occ, *kid, i, p;
    uint32_t *cov;
    int rc;
    if (!b) { fprintf(stderr, "[E::%s] the read database was not produced by sr_read_mem\n", __func__); return 0; }
    rc = sg_count(b);
    if (rc == SG_E_EMPTY) return 0;                            /* syncmer.c:1414-1417 */
    if (rc == SG_E_SMER_CONFLICT) {
        fprintf(stderr, "[E::%s] identical kmers have different smers\n", __func__);   /* the reference exits here */
        return 0;
    }
    if (rc != SG_OK || (rc = sg_count_sizes(b, &z)) != SG_OK) {
        fprintf(stderr, "[E::%s] %s: %s\n", __func__, sg_strerror(rc), sg_last_error(ctx_of(sr_db)));
        return 0;
    }
    h = malloc(8 * (z.n_unique + 1)); s = malloc(8 * (z.n_unique + 1)); cov = malloc(4 * (z.n_unique + 1));
    off = malloc(8 * (z.n_unique + 2)); occ = malloc(8 * (z.n_syncmers + 1)); kid = malloc(8 * (z.n_syncmers + 1));
    o.h = h; o.s = s; o.cov = cov; o.occ_off = off; o.occ = occ; o.k_mer_id = kid;
    if ((rc = sg_count_download(b, &o)) != SG_OK) {
        fprintf(stderr, "[E::%s] %s: %s\n", __func__, sg_strerror(rc), sg_last_error(ctx_of(sr_db)));
        free(h); free(s); free(cov); free(off); free(occ); free(kid);
        return 0;
    }
    db = (syncmer_db_t *) malloc(sizeof(syncmer_db_t));
    db->n = db->m = z.n_unique;
    db->a = (syncmer_t *) malloc(sizeof(syncmer_t) * (z.n_unique ? z.n_unique : 1));
    db->c = (uint16_t *) malloc(sizeof(uint16_t) * (z.n_unique ? z.n_unique : 1));
    db->h = 0;
    for (i = 0; i < z.n_unique; ++i) {
        syncmer_t *m = &db->a[i];
        m->h = h[i]; m->s = s[i]; m->cov = cov[i]; m->del = 0;
        m->m_pos = (uint64_t *) malloc(8 * (size_t) cov[i]);
        memcpy(m->m_pos, occ + off[i], 8 * (size_t) cov[i]);
        db->c[i] = 1;                                           /* syncmer.c:1443 */
    }
    /* k_mer[] now carries id << 1 (syncmer.c:1378) */
    for (i = 0, p = 0; i < sr_db->n; ++i) {
        sr_t *r = &sr_db->a[i];
        if (r->n) memcpy(r->k_mer, kid + p, 8 * (size_t) r->n);
        p += r->n;
    }
    free(h); free(s); free(cov); free(off); free(occ); free(kid);
    return db;
}

int syncmer_graph_arcs(sr_db_t *sr_db, syncmer_db_t *scm_db, uint32_t min_k_cov, double min_a_cov_f, uint64_t **arcs4, uint64_t *n_arcs)
{
    sg_batch *b = batch_of(sr_db, 0);
    int rc;
    (void) scm_db;
    *arcs4 = 0; *n_arcs = 0;
    if (!b) return SG_E_STATE;
    if ((rc = sg_arcs(b, min_k_cov, min_a_cov_f, n_arcs)) != SG_OK) return rc;
    *arcs4 = (uint64_t *) malloc(32 * (*n_arcs + 1));
    return sg_arcs_download(b, *arcs4);
}

/* read error correction rewrote the syncmer lists of the reads and the coverages of the database on the host: bring
 * the device-resident batch behind this sr_db_t up to date, so that the second sr_db_stat and the arc tally of the
 * final graph (both on the device) see the corrected lists. A database that was not produced by sr_read_mem has no
 * batch and nothing to refresh. */
/* fn(lo, hi, arg) over [0, n) cut into contiguous ranges, on up to 16 threads (the caller's included) */
typedef struct { void (*fn)(uint64_t, uint64_t, void *); void *arg; uint64_t lo, hi; } pf_job_t;
static void *pf_run(void *p) { pf_job_t *j = (pf_job_t *) p; j->fn(j->lo, j->hi, j->arg); return 0; }
void oatk_parallel_for(uint64_t n, void (*fn)(uint64_t lo, uint64_t hi, void *arg), void *arg)
{
    long nt = sysconf(_SC_NPROCESSORS_ONLN), t;
    pf_job_t job[16];
    pthread_t th[16];
    static long min_n = -1;                            /* OATK_PF_MIN: smallest n worth threads (tests lower it) */
    if (min_n < 0) { const char *e = getenv("OATK_PF_MIN"); min_n = e ? atol(e) : 65536; }
    if (nt > 16) nt = 16;
    if (nt < 1 || n < (uint64_t) min_n) nt = 1;
    if (nt == 1 && min_n <= 1 && n > 1) nt = 4;        /* ... and force threads even on a one-core box */
    for (t = 0; t < nt; ++t) { job[t].fn = fn; job[t].arg = arg; job[t].lo = n * (uint64_t) t / (uint64_t) nt; job[t].hi = n * (uint64_t) (t + 1) / (uint64_t) nt; }
    for (t = 1; t < nt; ++t) pthread_create(&th[t], 0, pf_run, &job[t]);
    pf_run(&job[0]);
    for (t = 1; t < nt; ++t) pthread_join(th[t], 0);
}

/* stage timer for tuning: with OATK_TIMING set, prints the time since the previous call to stderr */
void oatk_tick(const char *what)
{
    static int on = -1;
    static struct timespec last;
    struct timespec now;
    if (on < 0) { on = getenv("OATK_TIMING") != 0; clock_gettime(CLOCK_MONOTONIC, &last); }
    if (!on) return;
    clock_gettime(CLOCK_MONOTONIC, &now);
    if (what) fprintf(stderr, "[T::%s] %.3f s\n", what, (double) (now.tv_sec - last.tv_sec) + 1e-9 * (double) (now.tv_nsec - last.tv_nsec));
    last = now;
}

int oatk_gpu_update_lists(sr_db_t *sr_db, syncmer_db_t *scm_db)
{
    sg_batch *b = batch_of(sr_db, 0);
    uint64_t i, N = 0, *off, *km, *sm;
    uint32_t *mp, *cov;
    int rc;
    if (!b) return 0;
    off = (uint64_t *) malloc(8 * (sr_db->n + 1));
    off[0] = 0;
    for (i = 0; i < sr_db->n; ++i) { N += sr_db->a[i].n; off[i + 1] = N; }
    km = (uint64_t *) malloc(8 * (N + 1)); sm = (uint64_t *) malloc(8 * (N + 1)); mp = (uint32_t *) malloc(4 * (N + 1));
    cov = (uint32_t *) malloc(4 * (scm_db->n + 1));
    for (i = 0; i < sr_db->n; ++i) {
        const sr_t *r = &sr_db->a[i];
        if (!r->n) continue;
        memcpy(km + off[i], r->k_mer, 8 * (size_t) r->n);
        memcpy(sm + off[i], r->s_mer, 8 * (size_t) r->n);
        memcpy(mp + off[i], r->m_pos, 4 * (size_t) r->n);
    }
    for (i = 0; i < scm_db->n; ++i) cov[i] = scm_db->a[i].cov;
    rc = sg_batch_set_lists_host(b, sr_db->n, off, km, mp, sm, cov, scm_db->n);
    if (rc != SG_OK) fprintf(stderr, "[E::%s] %s: %s\n", __func__, sg_strerror(rc), sg_last_error(ctx_of(sr_db)));
    free(off); free(km); free(sm); free(mp); free(cov);
    return rc;
}

void sr_destroy(sr_t *sr)
{
    if (!sr) return;
    free(sr->sname); free(sr->hoco_s); free(sr->ho_rl); free(sr->ho_l_rl); free(sr->n_nucl);
    free(sr->s_mer); free(sr->k_mer); free(sr->m_pos);
}

void sr_db_clean(sr_db_t *sr_db)
{
    size_t i;
    if (!sr_db) return;
    for (i = 0; i < sr_db->n; ++i) sr_destroy(&sr_db->a[i]);
    free(sr_db->a);
    free(sr_db->stats);
    sr_db->a = 0; sr_db->stats = 0; sr_db->n = sr_db->m = 0;
    batch_drop(sr_db);
}

void sr_db_destroy(sr_db_t *sr_db)
{
    if (!sr_db) return;
    sr_db_clean(sr_db);
    free(sr_db);
}

void syncmer_db_init(syncmer_db_t *scm_db)
{
    if (!scm_db) return;
    scm_db->n = scm_db->m = 0; scm_db->a = 0; scm_db->c = 0; scm_db->h = 0;
}

void syncmer_db_clean(syncmer_db_t *scm_db)
{
    size_t i;
    if (!scm_db) return;
    for (i = 0; i < scm_db->n; ++i) free(scm_db->a[i].m_pos);
    free(scm_db->a); free(scm_db->c); free(scm_db->h);
}

void syncmer_db_destroy(syncmer_db_t *scm_db)
{
    if (!scm_db) return;
    syncmer_db_clean(scm_db);
    free(scm_db);
}

static inline int hoco_base(const uint8_t *hs, uint32_t p) { return (hs[p >> 2] >> ((3 - (p & 3)) << 1)) & 3; }

void get_kmer_seq(uint8_t *hoco_s, uint32_t pos, int l, uint32_t rev, uint8_t *kmer_s)
{
    int i;
    for (i = 0; i < l; ++i) kmer_s[i] = (uint8_t) (rev ? 3 - hoco_base(hoco_s, pos + l - 1 - i) : hoco_base(hoco_s, pos + i));
}

/* l bases from hoco position pos as text, reverse-complemented when rev. Whole packed bytes go through a table of
 * four characters per byte (the error correction unpacks every block of every read with this). */
void get_kmer_dna_seq(uint8_t *hoco_s, uint32_t pos, int l, uint32_t rev, char *dna_seq)
{
    static uint32_t fwd4[256], rc4[256];
    static volatile int ready;
    int i = 0;
    if (!ready) {
        int b, j;
        for (b = 0; b < 256; ++b) {
            char f[4], r[4];
            for (j = 0; j < 4; ++j) { const int c = (b >> ((3 - j) * 2)) & 3; f[j] = char_nt4_table[c]; r[3 - j] = char_nt4_table[3 - c]; }
            memcpy(&fwd4[b], f, 4); memcpy(&rc4[b], r, 4);
        }
        __sync_synchronize();
        ready = 1;
    }
    if (!rev) {
        uint32_t p = pos;
        for (; i < l && (p & 3); ++i, ++p) dna_seq[i] = char_nt4_table[hoco_base(hoco_s, p)];
        for (; i + 4 <= l; i += 4, p += 4) memcpy(dna_seq + i, &fwd4[hoco_s[p >> 2]], 4);
        for (; i < l; ++i, ++p) dna_seq[i] = char_nt4_table[hoco_base(hoco_s, p)];
    } else {
        uint32_t p = pos + (uint32_t) l;                   /* one past the base that comes out first */
        for (; i < l && (p & 3); ++i) dna_seq[i] = char_nt4_table[3 - hoco_base(hoco_s, --p)];
        for (; i + 4 <= l; i += 4) { p -= 4; memcpy(dna_seq + i, &rc4[hoco_s[p >> 2]], 4); }
        for (; i < l; ++i) dna_seq[i] = char_nt4_table[3 - hoco_base(hoco_s, --p)];
    }
}

void print_hoco_seq(sr_t *sr, FILE *fo)
{
    uint32_t i;
    for (i = 0; i < sr->hoco_l; ++i) fputc(char_nt4_table[hoco_base(sr->hoco_s, i)], fo);
    fputc('\n', fo);
}
