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

/* 1: sr_read_mem / sr_read_files / sr_read leave the run lengths (sr_t.ho_rl, one byte per hoco base: as large as the input)
 * on the device; sr_t.ho_rl is NULL and scg_consensus gets its run-length sums from there (oatk_gpu_runlen_sums). syncasm()
 * switches it on; callers of the API get the full read records unless they ask. */
static int g_keep_rl, g_keep_hs;
int oatk_gpu_keep_run_lengths(int on) { const int was = g_keep_rl; g_keep_rl = on != 0; return was; }
/* the same for the packed bases (sr_t.hoco_s == NULL): their consumers under syncasm() -- the consensus texts and the
 * read error correction -- are served on the device (oatk_gpu_kmer_codes, sg_ec_correct) */
int oatk_gpu_keep_packed_bases(int on) { const int was = g_keep_hs; g_keep_hs = on != 0; return was; }

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
        r->hoco_s = c->hoco_s_buf ? (uint8_t *) dup_block(c->hoco_s_buf + c->hoco_s_off[i], (c->hoco_l[i] + 3) / 4) : 0;   /* NULL: resident on the device */
        r->ho_rl = c->ho_rl_buf ? (uint8_t *) dup_block(c->ho_rl_buf + c->ho_rl_off[i], c->hoco_l[i]) : 0;   /* NULL: resident on the device */
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
    sg_pipe_keep_run_lengths(pipe, g_keep_rl);
    sg_pipe_keep_packed_bases(pipe, g_keep_hs);
    /* the pipeline's master batch is cut for 16 times the expected number of syncmers; low-complexity reads (a
     * dinucleotide repeat ties at every position: one syncmer per base) can need more, where the reference simply goes on:
     * the run is repeated with more room until there is one slot per base */
    for (unsigned factor = 16; ; factor *= 8) {
        sg_pipe_set_capacity_factor(pipe, factor);
        rc = sg_pipe_run_host_cb(pipe, bases, off, n_reads, k, s, SR_READ_CHUNK, fill_chunk, &f, &z);
        if (!(rc == SG_E_NOMEM && sg_pipe_syncmer_overflow(pipe)) || factor > (unsigned) (k - s + 1)) break;
        for (uint64_t i = 0; i < n_reads; ++i) sr_destroy(&sr_db->a[i]);        /* what the chunks before the overflow built */
        memset(sr_db->a, 0, (n_reads ? n_reads : 1) * sizeof(sr_t));
    }
    sg_pipe_set_capacity_factor(pipe, 16);
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

/*
 * Peaks of a multiplicity spectrum cnt[0..n) -- what the reference's ha_analyze_count computes (syncmer.c:775-865;
 * a heuristic it took from hifiasm) and, at verbose levels above 1, prints. The numbers and the printed lines are
 * part of the contract (-c 0 derives the coverage threshold from the k-mer peak, run_syncasm.c:90-93), so the rule
 * is restated here piece by piece:
 *   floor   the spectrum falls from its first used bin (1, or 2 when bin 1 is empty, and not below `first_bin`)
 *           until it first rises; the bin before the rise is the floor. A spectrum that never rises has no peak (-1).
 *   summit  the first highest bin after the floor.
 *   a side peak is the highest local maximum (>= both neighbours, first one wins ties) strictly between floor and
 *   summit (left) or after the summit, last bin excluded (right). It only counts when it reaches 5 % of the summit and
 *   the valley between the two dips below 95 % of it; a right peak must also lie within 2.5 x the summit's bin.
 * With a right peak the summit is the heterozygous peak and the right peak the homozygous one; otherwise the summit
 * is homozygous and a left peak, if any, heterozygous.
 */
typedef struct { int bin; int64_t height; } peak_t;

static int64_t lowest_between(const int64_t *cnt, int from, int to, int64_t start)
{
    int i;
    for (i = from; i < to; ++i) if (cnt[i] < start) start = cnt[i];
    return start;
}

/* highest local maximum among bins [from, to), scanned in the given direction; bin -1 if there is none */
static peak_t best_local_maximum(const int64_t *cnt, int from, int to, int step)
{
    peak_t best = { -1, -1 };
    int i;
    for (i = step > 0 ? from : to - 1; i >= from && i < to; i += step)
        if (cnt[i] >= cnt[i - 1] && cnt[i] >= cnt[i + 1] && cnt[i] > best.height) best.bin = i, best.height = cnt[i];
    return best;
}

static void spectrum_line(int bin, int64_t value, int64_t summit)
{
    const int width = 100;
    int stars = (int) ((double) width * value / summit + .499), over = stars > width, j;
    if (over) stars = width;
    if (bin >= 0) fprintf(stderr, "[M::ha_hist_line] %5d: ", bin);
    else fprintf(stderr, "[M::ha_hist_line] %5s: ", "rest");
    for (j = 0; j < stars; ++j) fputc('*', stderr);
    if (over) fputc('>', stderr);
    fprintf(stderr, " %lld\n", (long long) value);
}

static int find_peaks(int n, int first_bin, const int64_t *cnt, int *peak_het, int verbose)
{
    const int first_used = cnt[1] > 0 ? 1 : 2;
    int floor_bin = first_used > first_bin ? first_used : first_bin, i;
    peak_t summit, left, right;
    *peak_het = -1;
    while (floor_bin + 1 < n && cnt[floor_bin + 1] <= cnt[floor_bin]) ++floor_bin;
    if (verbose > 0) fprintf(stderr, "[M::ha_analyze_count] lowest: count[%d] = %ld\n", floor_bin, (long) cnt[floor_bin]);
    if (floor_bin == n - 1) return -1;

    summit.bin = floor_bin + 1; summit.height = cnt[summit.bin];
    for (i = floor_bin + 2; i < n; ++i) if (cnt[i] > summit.height) summit.bin = i, summit.height = cnt[i];
    if (verbose > 0) {
        int64_t rest = 0;
        fprintf(stderr, "[M::ha_analyze_count] highest: count[%d] = %ld\n", summit.bin, (long) cnt[summit.bin]);
        /* bins up to the first empty-looking one after the summit, the remainder as one line */
        for (i = first_used; i < n; ++i) {
            if (i > summit.bin && (int) ((double) 100 * cnt[i] / summit.height + .499) == 0) break;
            spectrum_line(i, cnt[i], summit.height);
        }
        for (; i < n; ++i) rest += cnt[i];
        spectrum_line(-1, rest, summit.height);
    }

    left = best_local_maximum(cnt, floor_bin + 1, summit.bin, -1);
    if (left.bin > floor_bin && left.bin < summit.bin) {
        const int64_t valley = lowest_between(cnt, left.bin + 1, summit.bin, summit.height);
        if (left.height < summit.height * 0.05 || valley > left.height * 0.95) left.bin = -1, left.height = -1;
    }
    if (verbose > 0) {
        if (left.height > 0) fprintf(stderr, "[M::ha_analyze_count] left: count[%d] = %ld\n", left.bin, (long) cnt[left.bin]);
        else fprintf(stderr, "[M::ha_analyze_count] left: none\n");
    }
    right = best_local_maximum(cnt, summit.bin + 1, n - 1, 1);
    if (right.bin > summit.bin) {
        const int64_t valley = lowest_between(cnt, summit.bin + 1, right.bin, summit.height);
        if (right.height < summit.height * 0.05 || valley > right.height * 0.95 || right.bin > summit.bin * 2.5)
            right.bin = -1, right.height = -1;
    }
    if (verbose > 0) {
        if (right.height > 0) fprintf(stderr, "[M::ha_analyze_count] right: count[%d] = %ld\n", right.bin, (long) cnt[right.bin]);
        else fprintf(stderr, "[M::ha_analyze_count] right: none\n");
    }
    if (right.bin > 0) { *peak_het = summit.bin; return right.bin; }
    if (left.bin > 0) *peak_het = left.bin;
    return summit.bin;
}

/*
 * The three tables sr_db_stat prints at verbose levels above 1 (syncmer.c:1018-1022 -> kh_ctab_print :736-757 ->
 * hist_plot :669-734): gap between neighbouring syncmers of a read, s-mer multiplicities, k-mer multiplicities, each
 * as (value, how often) pairs in ascending order of value. The device keeps only what the statistics need (bins up
 * to 1000, the gap sum), so for this diagnostic output the full tables are rebuilt from the per-read lists on the host.
 */
typedef struct { int64_t value; int64_t times; } tally_t;

static int cmp_i64(const void *a, const void *b) { int64_t x = *(const int64_t *) a, y = *(const int64_t *) b; return (x > y) - (x < y); }
static int cmp_u64(const void *a, const void *b) { uint64_t x = *(const uint64_t *) a, y = *(const uint64_t *) b; return (x > y) - (x < y); }

/* sorted values -> (value, times) pairs; returns how many */
static size_t tally_sorted_i64(const int64_t *v, size_t n, tally_t *out)
{
    size_t i, m = 0;
    for (i = 0; i < n; ++i) {
        if (m && out[m - 1].value == v[i]) ++out[m - 1].times;
        else { out[m].value = v[i]; out[m].times = 1; ++m; }
    }
    return m;
}

/* keys -> multiplicity of every distinct key -> (multiplicity, times) pairs */
static size_t multiplicity_table(uint64_t *keys, size_t n, tally_t **out)
{
    int64_t *mult = (int64_t *) malloc(sizeof(int64_t) * (n ? n : 1));
    size_t i, m = 0, run = 0;
    qsort(keys, n, sizeof(uint64_t), cmp_u64);
    for (i = 0; i < n; ++i) {
        ++run;
        if (i + 1 == n || keys[i + 1] != keys[i]) { mult[m++] = (int64_t) run; run = 0; }
    }
    qsort(mult, m, sizeof(int64_t), cmp_i64);
    *out = (tally_t *) malloc(sizeof(tally_t) * (m ? m : 1));
    m = tally_sorted_i64(mult, m, *out);
    free(mult);
    return m;
}

static int printed_width(int32_t v)
{
    int w = v > 0 ? 0 : 1;              /* the reference counts a sign position for zero as well */
    do { v /= 10; ++w; } while (v != 0);
    return w;
}

static void bar(FILE *fo, double count, double per_dot)
{
    const double dots = count / per_dot;
    uint32_t j, n = (uint32_t) dots;
    for (j = 0; j < (n < 100 ? n : 100); ++j) fputc('*', fo);
    n = dots > 100 ? (uint32_t) log10(dots / 100) : 0;
    for (j = 0; j < n; ++j) fputc('+', fo);
    fprintf(fo, " %d\n", (int) count);
}

static void print_table(const tally_t *t, size_t n, const char *title, FILE *fo, int list_all)
{
    size_t i, shown = 0;
    if (n >= 5) {
        /* rows until 99 % of the mass outside the three smallest values is covered; the rest as one line */
        double mass = 0, acc = 0, rest = 0;
        uint32_t tallest = 0;
        int width = 0;
        for (i = 3; i < n; ++i) mass += (uint32_t) t[i].times;
        mass *= .99;
        for (i = 0; i < n; ++i) {
            if (i >= 3) acc += (uint32_t) t[i].times;
            if (acc >= mass) { shown = i + 1; break; }
        }
        for (i = 0; i < shown; ++i) {
            const int w = printed_width((int32_t) t[i].value);
            if (i >= 3 && (uint32_t) t[i].times > tallest) tallest = (uint32_t) t[i].times;
            if (w > width) width = w;
        }
        if (shown < n) ++width;
        {
            const double per_dot = tallest / 100 > 1 ? tallest / 100 : 1;
            for (i = 0; i < shown; ++i) {
                fprintf(fo, "[M::hist_plot] [%s] %*d: ", title, width, (int32_t) t[i].value);
                bar(fo, (uint32_t) t[i].times, per_dot);
            }
            if (shown < n) {
                for (i = shown; i < n; ++i) rest += (uint32_t) t[i].times;
                fprintf(fo, "[M::hist_plot] [%s] >%*d: ", title, width - 1, (int32_t) t[shown - 1].value);
                bar(fo, rest, per_dot);
            }
        }
    }
    if (list_all)
        for (i = 0; i < n; ++i) fprintf(fo, "[M::kh_ctab_print] [%s CNTS] %ld %d\n", title, (long) t[i].value, (int) t[i].times);
}

static void print_verbose_tables(sr_db_t *db, FILE *fo, int list_all)
{
    uint64_t total = 0, i, j, n_gap = 0, p;
    int64_t *gap;
    uint64_t *key;
    tally_t *t;
    size_t m;
    for (i = 0; i < db->n; ++i) total += db->a[i].n;
    gap = (int64_t *) malloc(sizeof(int64_t) * (total ? total : 1));
    key = (uint64_t *) malloc(sizeof(uint64_t) * (total ? total : 1));
    for (i = 0; i < db->n; ++i) {
        const sr_t *r = &db->a[i];
        int prev = (int) MAX_RD_LEN;
        for (j = 0; j < r->n; ++j) {                           /* corrected entries carry MAX_RD_LEN and break the chain (syncmer.c:896-902) */
            const int here = (int) (r->m_pos[j] >> 1);
            if (prev != (int) MAX_RD_LEN && here != (int) MAX_RD_LEN) gap[n_gap++] = (int64_t) here - prev - db->k;
            prev = here;
        }
    }
    qsort(gap, n_gap, sizeof(int64_t), cmp_i64);
    t = (tally_t *) malloc(sizeof(tally_t) * (n_gap ? n_gap : 1));
    m = tally_sorted_i64(gap, n_gap, t);
    print_table(t, m, "DIST", fo, list_all);
    free(t); free(gap);
    for (i = 0, p = 0; i < db->n; ++i) for (j = 0; j < db->a[i].n; ++j) key[p++] = db->a[i].s_mer[j];
    m = multiplicity_table(key, total, &t);
    print_table(t, m, "SMER", fo, list_all);
    free(t);
    for (i = 0, p = 0; i < db->n; ++i) for (j = 0; j < db->a[i].n; ++j) key[p++] = db->a[i].k_mer[j] >> 1;
    m = multiplicity_table(key, total, &t);
    print_table(t, m, "KMER", fo, list_all);
    free(t); free(key);
}

/*
 * The singleton count of sr_db_stat when NO key occurs exactly once. The reference tabulates "how many keys have
 * multiplicity m" in a khashl map and reads the singletons as the value under key 1; when that key is absent it
 * reports whatever value its scan over the table met last (kh_ctab_stat, syncmer.c:618-646: `c` keeps the value of the
 * last occupied slot). After read error correction there are normally no singletons left, so that number is on
 * stderr in every default run and is reproduced here: the multiplicities are replayed in the order the reference
 * inserts them (ascending key order, from the device) into a table with khashl 0.1's geometry -- capacity a power
 * of two from 4 up, bucket = Wang hash x 2654435769 >> (32 - bits), linear probing, doubling at 75 % load where the old
 * slots are re-seated in slot order and an element that lands on a not yet re-seated one takes its place and
 * sends it on (khashl.h:130-211).
 */
typedef struct { uint32_t key; int val; uint8_t full; } mslot_t;

static uint32_t mslot_home(uint32_t key, int bits)
{
    key += ~(key << 15); key ^= key >> 10; key += key << 3; key ^= key >> 6; key += ~(key << 11); key ^= key >> 16;
    return (key * 2654435769u) >> (32 - bits);
}

static int value_of_last_slot(const uint32_t *mult, uint64_t n)
{
    mslot_t *tab = 0;
    uint32_t cap = 0, live = 0, j;
    int bits = 0, last = 0;
    uint64_t x;
    for (x = 0; x < n; ++x) {
        uint32_t i, start;
        if (live >= (cap >> 1) + (cap >> 2)) {              /* also true for the empty table: 0 >= 0 */
            int nbits = 2;
            uint32_t ncap;
            mslot_t *next;
            while ((1u << nbits) < cap + 1) ++nbits;
            ncap = 1u << nbits;
            next = (mslot_t *) calloc(ncap, sizeof(mslot_t));
            for (j = 0; j < cap; ++j) {
                mslot_t moving;
                if (!tab[j].full) continue;
                moving = tab[j]; tab[j].full = 0;
                for (;;) {
                    i = mslot_home(moving.key, nbits);
                    while (next[i].full) i = (i + 1) & (ncap - 1);
                    next[i] = moving;
                    if (i < cap && tab[i].full) { moving = tab[i]; tab[i].full = 0; }   /* the old tenant of that slot goes next */
                    else break;
                }
            }
            free(tab);
            tab = next; cap = ncap; bits = nbits;
        }
        i = start = mslot_home(mult[x], bits);
        while (tab[i].full && tab[i].key != mult[x]) { i = (i + 1) & (cap - 1); if (i == start) break; }
        if (!tab[i].full) { tab[i].key = mult[x]; tab[i].val = 1; tab[i].full = 1; ++live; }
        else ++tab[i].val;
    }
    for (j = 0; j < cap; ++j) if (tab[j].full) last = tab[j].val;
    free(tab);
    return last;
}

static int singletons_as_reported(sg_batch *b, int which, int64_t true_count)
{
    uint32_t *mult = 0;
    uint64_t n = 0;
    int v;
    if (true_count > 0) return (int) true_count;
    if (sg_stat_multiplicities(b, which, &mult, &n) != SG_OK || n == 0) { free(mult); return 0; }
    v = value_of_last_slot(mult, n);
    free(mult);
    return v;
}

void sr_db_stat(sr_db_t *sr_db, FILE *fo, int verbose)
{
    sg_batch *b = batch_of(sr_db, 0);
    sg_stat_t st;
    sr_stat_t *s;
    int rc;
    if (!sr_db->stats) sr_db->stats = (sr_stat_t *) calloc(1, sizeof(sr_stat_t));
    s = sr_db->stats;
    if (!b) { fprintf(stderr, "[E::%s] the read database was not produced by sr_read_mem\n", __func__); return; }
    rc = sg_stat(b, &st);
    if (rc == SG_E_EMPTY) { fprintf(fo, "[M::%s] empty syncmer collection\n", __func__); return; }   /* syncmer.c:909-912 */
    if (rc != SG_OK) { fprintf(stderr, "[E::%s] %s: %s\n", __func__, sg_strerror(rc), sg_last_error(ctx_of(sr_db))); return; }
    s->syncmer_n = st.n_syncmers;
    s->syncmer_per_read = (double) st.n_syncmers / (double) sr_db->n;
    s->syncmer_avg_dist = (double) st.gap_sum / (double) st.n_gaps;
    s->smer_unique = (int) st.smer_unique; s->smer_singleton = singletons_as_reported(b, 1, st.smer_cnts[1]);
    s->smer_avg_cnt = (double) st.n_syncmers / (double) st.smer_unique;
    s->kmer_unique = (int) st.kmer_unique; s->kmer_singleton = singletons_as_reported(b, 0, st.kmer_cnts[1]);
    s->kmer_avg_cnt = (double) st.n_syncmers / (double) st.kmer_unique;
    s->smer_peak_hom = find_peaks(1001, 5, st.smer_cnts, &s->smer_peak_het, verbose - 1);
    s->kmer_peak_hom = find_peaks(1001, 5, st.kmer_cnts, &s->kmer_peak_het, verbose - 1);
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
    if (verbose > 1) print_verbose_tables(sr_db, fo, verbose - 1 > 0);
}

typedef struct { syncmer_db_t *db; const uint64_t *h, *s; const uint32_t *cov; const uint64_t *off, *occ; } collect_fill_t;
static void collect_fill(uint64_t lo, uint64_t hi, void *arg)
{
    const collect_fill_t *F = (const collect_fill_t *) arg;
    for (uint64_t i = lo; i < hi; ++i) {
        syncmer_t *m = &F->db->a[i];
        m->h = F->h[i]; m->s = F->s[i]; m->cov = F->cov[i]; m->del = 0;
        m->m_pos = (uint64_t *) malloc(8 * (size_t) F->cov[i]);
        memcpy(m->m_pos, F->occ + F->off[i], 8 * (size_t) F->cov[i]);
        F->db->c[i] = 1;                                        /* syncmer.c:1443 */
    }
}
typedef struct { sr_db_t *db; const uint64_t *kid, *first; } collect_ids_t;
static void collect_ids(uint64_t lo, uint64_t hi, void *arg)
{
    const collect_ids_t *K = (const collect_ids_t *) arg;
    for (uint64_t i = lo; i < hi; ++i) {
        sr_t *r = &K->db->a[i];
        if (r->n) memcpy(r->k_mer, K->kid + K->first[i], 8 * (size_t) r->n);
    }
}

static int oatk_collect_conflict_flag = 0;
int oatk_collect_conflict(void) { return oatk_collect_conflict_flag; }

/* the peak finder alone, for the tests that compare it with the reference's ha_analyze_count on histograms of their own */
int oatk_find_peaks(int n, int first_bin, const int64_t *cnt, int *peak_het) { return find_peaks(n, first_bin, cnt, peak_het, 0); }

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
    oatk_collect_conflict_flag = 0;
    if (rc == SG_E_SMER_CONFLICT) {
        /* the reference prints these four lines from process_kmer_cluster and exits (syncmer.c:1370-1376); this layer returns
         * NULL and leaves the decision to the caller (oatk_collect_conflict() tells it apart from an empty database) */
        uint64_t c[5] = {0, 0, 0, 0, 0};
        sg_count_conflict(b, c);
        fprintf(stderr, "[E::process_kmer_cluster] identical kmers have different smers\n");
        fprintf(stderr, "[E::process_kmer_cluster] kmer hash  : %lu\n", (unsigned long) c[0]);
        fprintf(stderr, "[E::process_kmer_cluster] smer code 0: %lu; read id: %lu\n", (unsigned long) c[1], (unsigned long) c[2]);
        fprintf(stderr, "[E::process_kmer_cluster] smer code 1: %lu; read id: %lu\n", (unsigned long) c[3], (unsigned long) c[4]);
        oatk_collect_conflict_flag = 1;
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
    /* one malloc block per occurrence list (syncmer.c:1359; freed one by one, syncmer.c:1099): 10^6 small blocks, built
     * by worker threads over ranges of ids */
    { collect_fill_t F = {db, h, s, cov, off, occ}; oatk_parallel_for(z.n_unique, collect_fill, &F); }
    /* k_mer[] now carries id << 1 (syncmer.c:1378) */
    {
        uint64_t *first = (uint64_t *) malloc(8 * (sr_db->n + 1));
        for (i = 0, p = 0; i < sr_db->n; ++i) { first[i] = p; p += sr_db->a[i].n; }
        collect_ids_t K = {sr_db, kid, first};
        oatk_parallel_for(sr_db->n, collect_ids, &K);
        free(first);
    }
    free(h); free(s); free(cov); free(off); free(occ); free(kid);
    return db;
}

/* 1 when the run lengths of this read database live on the device only (see oatk_gpu_keep_run_lengths) */
int oatk_gpu_run_lengths_on_device(sr_db_t *sr_db)
{
    sg_batch *b = batch_of(sr_db, 0);
    uint64_t i;
    if (!b || !sg_runlen_resident(b)) return 0;
    for (i = 0; i < sr_db->n; ++i) if (sr_db->a[i].hoco_l) return sr_db->a[i].ho_rl == 0;
    return 0;
}

/* 1 when the packed bases of this read database live on the device only (see oatk_gpu_keep_packed_bases) */
int oatk_gpu_bases_on_device(sr_db_t *sr_db)
{
    uint64_t i;
    if (!batch_of(sr_db, 0)) return 0;
    for (i = 0; i < sr_db->n; ++i) if (sr_db->a[i].hoco_l) return sr_db->a[i].hoco_s == 0;
    return 0;
}

/* hoco bases (codes 0..3, one per byte) of n stretches of len positions: refs[i] = read << 32 | first position (sg_kmer_codes) */
int oatk_gpu_kmer_codes(sr_db_t *sr_db, uint64_t n, const uint64_t *refs, int len, uint8_t *codes)
{
    sg_batch *b = batch_of(sr_db, 0);
    int rc;
    if (!b) return SG_E_STATE;
    rc = sg_kmer_codes(b, n, refs, len, codes);
    if (rc != SG_OK) fprintf(stderr, "[E::%s] %s: %s\n", __func__, sg_strerror(rc), sg_last_error(ctx_of(sr_db)));
    return rc;
}

/* sums of (run length - 1) per hoco position for n_req syncmers over the listed occurrences (sg_runlen_sums) */
int oatk_gpu_runlen_sums(sr_db_t *sr_db, uint64_t n_req, const uint64_t *occ_off, const uint64_t *occ, uint64_t *sums)
{
    sg_batch *b = batch_of(sr_db, 0);
    int rc;
    if (!b) return SG_E_STATE;
    rc = sg_runlen_sums(b, n_req, occ_off, occ, sums);
    if (rc != SG_OK) fprintf(stderr, "[E::%s] %s: %s\n", __func__, sg_strerror(rc), sg_last_error(ctx_of(sr_db)));
    return rc;
}

/* f2 on the device (sg_ec_correct): 1 when this read database has a device-resident batch whose lists can be searched */
int oatk_gpu_ec_available(sr_db_t *sr_db)
{
    return batch_of(sr_db, 0) != 0;
}

/* graph / result are sg_ec_graph_t / sg_ec_result_t of include/syncgpu.h (kept opaque here so that the headers of the host
 * layer do not pull the device ABI in) */
int oatk_gpu_ec_correct(sr_db_t *sr_db, const void *graph, double max_edist, void *result)
{
    sg_batch *b = batch_of(sr_db, 0);
    int rc;
    if (!b) return SG_E_STATE;
    rc = sg_ec_correct(b, (const sg_ec_graph_t *) graph, max_edist, (sg_ec_result_t *) result);
    if (rc != SG_OK) fprintf(stderr, "[E::%s] %s: %s\n", __func__, sg_strerror(rc), sg_last_error(ctx_of(sr_db)));
    return rc;
}

/* the distance votes of a list of arcs on the device (sg_arc_votes); non-zero: count them on the host */
int oatk_gpu_arc_votes(sr_db_t *sr_db, uint64_t n, const uint64_t *arcs4, int32_t *dist, uint8_t *flag)
{
    sg_batch *b = batch_of(sr_db, 0);
    if (!b) return SG_E_STATE;
    return sg_arc_votes(b, n, arcs4, dist, flag);          /* a state error (lists replaced since sg_count) is not an error here */
}

/* the error filter on the device (sg_ec_filter); result is an sg_ec_filter_out_t */
int oatk_gpu_ec_filter(sr_db_t *sr_db, uint32_t err_mer_c, uint32_t max_err_c, uint32_t err_arc_c, double max_arc_f, void *result)
{
    sg_batch *b = batch_of(sr_db, 0);
    int rc;
    if (!b) return SG_E_STATE;
    rc = sg_ec_filter(b, 0, err_mer_c, max_err_c, err_arc_c, max_arc_f, (sg_ec_filter_out_t *) result);
    if (rc != SG_OK) fprintf(stderr, "[E::%s] %s: %s\n", __func__, sg_strerror(rc), sg_last_error(ctx_of(sr_db)));
    return rc;
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
/* how many host threads the helper pools of this layer may use: the cores of the machine, at most 16, at most what the
 * caller of syncasm() asked for with -t (0: no wish) */
static int g_host_threads = 0;
void oatk_set_host_threads(int n) { __atomic_store_n(&g_host_threads, n > 0 ? n : 0, __ATOMIC_RELAXED); }
long oatk_host_threads(void)
{
    long nt = sysconf(_SC_NPROCESSORS_ONLN);
    const int cap = __atomic_load_n(&g_host_threads, __ATOMIC_RELAXED);
    if (nt > 16) nt = 16;
    if (cap > 0 && nt > cap) nt = cap;
    return nt < 1 ? 1 : nt;
}

void oatk_parallel_for(uint64_t n, void (*fn)(uint64_t lo, uint64_t hi, void *arg), void *arg)
{
    long nt = oatk_host_threads(), t;
    pf_job_t job[16];
    pthread_t th[16];
    static long min_n_cached = -1;                     /* OATK_PF_MIN: smallest n worth threads (tests lower it) */
    long min_n = __atomic_load_n(&min_n_cached, __ATOMIC_RELAXED);
    if (min_n < 0) { const char *e = getenv("OATK_PF_MIN"); min_n = e ? atol(e) : 65536; __atomic_store_n(&min_n_cached, min_n, __ATOMIC_RELAXED); }
    if (n < (uint64_t) min_n) nt = 1;
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
    oatk_cons_cache_drop(sr_db);                       /* the lists changed: run-length sums kept for them are stale */
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
    oatk_cons_cache_drop(sr_db);
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
static uint32_t fwd4[256], rc4[256];
static pthread_once_t dna4_once = PTHREAD_ONCE_INIT;
static void dna4_init(void)
{
    int b, j;
    for (b = 0; b < 256; ++b) {
        char f[4], r[4];
        for (j = 0; j < 4; ++j) { const int c = (b >> ((3 - j) * 2)) & 3; f[j] = char_nt4_table[c]; r[3 - j] = char_nt4_table[3 - c]; }
        memcpy(&fwd4[b], f, 4); memcpy(&rc4[b], r, 4);
    }
}

void get_kmer_dna_seq(uint8_t *hoco_s, uint32_t pos, int l, uint32_t rev, char *dna_seq)
{
    int i = 0;
    pthread_once(&dna4_once, dna4_init);                   /* called from worker threads: the tables are built exactly once */
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
