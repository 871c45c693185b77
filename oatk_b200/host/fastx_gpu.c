/*
 * fastx_gpu.c -- row f4 of SURVEY.md section 8: the input path in front of sr_read_mem.
 *
 * The reference pulls records one at a time through sstream_read -> kseq_read (sstream.c:84-104,
 * kseq.h:193-240) over zlib's gzread and strdup()s name and sequence per record (syncmer.c:522-526): about
 * 0.9 GB/s on one thread, which is most of its sr_read wall time once the analysis threads are fast. Here a
 * file is mapped (or inflated) once and cut into records with memchr/memcpy straight into the flat
 * (bases, offsets, names) arrays that sr_read_mem takes. The record grammar is kseq's, restated:
 *
 *   header    after a FASTQ record or at the start, skip to the next '>' or '@' ANYWHERE; otherwise the header
 *             character was the first character of the line that ended the previous sequence
 *   name      up to the first white space; the rest of the line is a comment (dropped, like the reference keeps
 *             only name.s)
 *   sequence  every following line, until a line that STARTS with '>', '@' or '+'; empty lines are skipped; one
 *             trailing '\r' per line is dropped once the sequence is longer than one character
 *   quality   after '+': the rest of that line is skipped, then lines are appended until the quality is at least as
 *             long as the sequence (at least one line is consumed); a different length, or no line at all, is a
 *             truncated record: kseq returns -2, sstream_read passes it on, and the reference's read loop stops
 *             with that FILE (the next file, if any, is opened)
 *   -D cap    reading stops after the record that makes the running total reach the limit (syncmer.c:537-541)
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <ctype.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <zlib.h>
#include <pthread.h>
#include "syncmer_gpu.h"
#include "fastx_gpu.h"

#define KS_BUFSIZE 16384u         /* kseq's stream buffer: end of file is only known once a short read has happened */

typedef struct { const char *p; size_t n; int mapped; char *owned; } blob_t;

/* whole gzip stream (possibly several members; plain data passes through unchanged) into one buffer */
static int blob_inflate(int fd, size_t size_hint, const char *path, blob_t *b)
{
    gzFile g = gzdopen(fd, "r");
    size_t cap = size_hint * 4 + (1 << 20), n = 0;
    char *buf;
    int r = 0, zerr = 0;
    if (!g) { close(fd); return -1; }
    gzbuffer(g, 1 << 20);
    buf = (char *) malloc(cap);
    while (buf) {
        if (cap - n < (1 << 20)) {
            char *grown = (char *) realloc(buf, cap + cap / 2);
            if (!grown) { free(buf); buf = 0; break; }
            buf = grown; cap += cap / 2;
        }
        r = gzread(g, buf + n, (unsigned) ((cap - n) > (1u << 30) ? (1u << 30) : (cap - n)));
        if (r <= 0) break;
        n += (size_t) r;
    }
    if (!buf) {
        fprintf(stderr, "[E::%s] out of memory while reading \"%s\"\n", __func__, path);
        gzclose(g);
        return -1;
    }
    if (r < 0) {
        /* kseq treats a read error as the end of the stream (kseq.h:96-101) and so does this reader: the records
         * before the damage are kept. Said aloud here, silently there. */
        const char *msg = gzerror(g, &zerr);
        fprintf(stderr, "[W::%s] \"%s\": %s; keeping what was read before the error\n", __func__, path, msg ? msg : "read error");
    }
    gzclose(g);
    b->p = b->owned = buf; b->n = n;
    return 0;
}

static int blob_open(const char *path, blob_t *b)
{
    memset(b, 0, sizeof(*b));
    if (strcmp(path, "-") == 0) return blob_inflate(dup(STDIN_FILENO), 1 << 24, path, b);   /* kopen.c reads "-" from stdin */
    int fd = open(path, O_RDONLY);
    if (fd < 0) return -1;
    unsigned char magic[2] = {0, 0};
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return -1; }
    if (!S_ISREG(st.st_mode)) return blob_inflate(fd, 1 << 24, path, b);                      /* pipes, process substitution */
    ssize_t got = pread(fd, magic, 2, 0);
    if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b)
        return blob_inflate(fd, (size_t) st.st_size, path, b);          /* zlib is the reference's reader too */
    if (st.st_size == 0) { close(fd); b->p = ""; b->n = 0; return 0; }
    /* not MAP_POPULATE: the parser's threads fault their own pieces in, in parallel */
    void *m = mmap(0, (size_t) st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return -1;
    madvise(m, (size_t) st.st_size, MADV_SEQUENTIAL);
    b->p = (const char *) m; b->n = (size_t) st.st_size; b->mapped = 1;
    return 0;
}

int fastx_can_open(const char *path)
{
    if (strcmp(path, "-") == 0) return 0;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return -1;
    close(fd);
    return 0;
}

static void blob_close(blob_t *b)
{
    if (b->mapped) munmap((void *) b->p, b->n);
    free(b->owned);
    memset(b, 0, sizeof(*b));
}

/* where parsed records go: either appended to growing arrays (bases/off/names of a fastx_t) or, in the parallel
 * path, counted first and then written at known places */
typedef struct {
    char *bases; uint64_t n_bases, m_bases;            /* m_bases = 0: fixed buffer, never grown */
    uint64_t *off; char **names; uint64_t n, m;        /* m = 0: fixed arrays */
    int count_only;
} sink_t;

static inline void sink_bases(sink_t *k, const char *src, size_t l)
{
    if (!k->count_only) {
        if (k->m_bases && k->n_bases + l + 1 > k->m_bases) {
            k->m_bases = (k->n_bases + l + 1) + (k->n_bases + l + 1) / 2 + 4096;
            k->bases = (char *) realloc(k->bases, k->m_bases);
        }
        memcpy(k->bases + k->n_bases, src, l);
    }
    k->n_bases += l;
}

static inline void sink_record(sink_t *k, const char *name, size_t name_l)
{
    if (!k->count_only) {
        if (k->m && k->n + 2 > k->m) {
            k->m = k->m * 2;
            k->off = (uint64_t *) realloc(k->off, (k->m + 1) * sizeof(uint64_t));
            k->names = (char **) realloc(k->names, k->m * sizeof(char *));
        }
        char *s = (char *) malloc(name_l + 1);
        memcpy(s, name, name_l);
        s[name_l] = 0;
        k->names[k->n] = s;
        k->off[k->n + 1] = k->n_bases;
    }
    ++k->n;
}

/* one file (or a run of whole FASTA records of it) held in memory; returns 0 at its end, 1 when the data limit was
 * reached, -2 on a truncated record. *odd is set when a line starts with '@' or '+' (not plain FASTA). */
static int parse_blob(const char *buf, size_t n, int eof_known_at_end, sink_t *x, uint64_t max_bases, uint64_t *total, int *odd)
{
    size_t pos = 0;
    int last_char = 0;
    for (;;) {
        if (!last_char) {                               /* jump to the next header character, wherever it is */
            const char *a = (const char *) memchr(buf + pos, '>', n - pos), *b = (const char *) memchr(buf + pos, '@', n - pos);
            const char *h = !a ? b : (!b ? a : (a < b ? a : b));
            if (!h) return 0;
            if (*h == '@') *odd = 1;
            pos = (size_t) (h - buf) + 1;
        }
        if (pos >= n && eof_known_at_end) return 0;     /* a header character at the very end: no record */
        size_t q = pos;
        while (q < n && !isspace((unsigned char) buf[q])) ++q;
        const char *name = buf + pos;
        const size_t name_l = q - pos;
        int c = q < n ? buf[q] : 0;
        pos = q < n ? q + 1 : n;
        if (c != '\n') {                                /* comment: the rest of the line */
            const char *e = (const char *) memchr(buf + pos, '\n', n - pos);
            pos = e ? (size_t) (e - buf) + 1 : n;
        }
        const uint64_t seq0 = x->n_bases;
        c = -1;
        while (pos < n) {
            c = buf[pos];
            if (c == '>' || c == '+' || c == '@') break;
            ++pos;
            if (c == '\n') { c = -1; continue; }        /* empty line */
            const int had_more = !(pos >= n && eof_known_at_end);
            const char *e = (const char *) memchr(buf + pos, '\n', n - pos);
            const size_t end = e ? (size_t) (e - buf) : n;
            size_t l = end - (pos - 1);
            /* one trailing '\r' goes once the sequence is longer than a character; decided before the copy so that a
             * piece of the parallel path never writes past its own window */
            if (had_more && (x->n_bases - seq0) + l > 1 && buf[end - 1] == '\r') --l;
            sink_bases(x, buf + pos - 1, l);
            pos = e ? end + 1 : n;
            c = -1;
        }
        if (pos < n && (c == '>' || c == '+' || c == '@')) ++pos; else c = -1;
        if (c == '+' || c == '@') *odd = 1;
        last_char = (c == '>' || c == '@') ? c : last_char;
        const uint64_t seq_l = x->n_bases - seq0;
        if (c == '+') {
            const char *e = (const char *) memchr(buf + pos, '\n', n - pos);
            if (!e) { x->n_bases = seq0; return -2; }   /* no quality string */
            pos = (size_t) (e - buf) + 1;
            uint64_t qual_l = 0;
            do {
                if (pos >= n && eof_known_at_end) break;
                e = (const char *) memchr(buf + pos, '\n', n - pos);
                const size_t end = e ? (size_t) (e - buf) : n;
                qual_l += end - pos;
                if (qual_l > 1 && end > pos && buf[end - 1] == '\r') --qual_l;
                pos = e ? end + 1 : n;
                if (!e && pos >= n) { if (qual_l < seq_l) break; }
            } while (qual_l < seq_l && pos < n);
            last_char = 0;
            if (qual_l != seq_l) { x->n_bases = seq0; return -2; }
        } else if (c != '>' && c != '@') last_char = 0; /* end of file inside a FASTA record */
        sink_record(x, name, name_l);
        *total += seq_l;
        if (max_bases && *total >= max_bases) return 1;
        if (c == -1 && pos >= n) return 0;
    }
}

/* ---- plain FASTA in parallel: records start exactly at the lines that begin with '>', so the file can be cut
 * there; every piece is parsed twice (count, then write at the places the counts give). Any line that begins with
 * '@' or '+' sends the whole file back to the sequential parser. ---- */
typedef struct {
    const char *buf; size_t n; int eof_known; sink_t sink; int odd; int rc; pthread_t th;
} piece_t;

static void *piece_run(void *arg)
{
    piece_t *p = (piece_t *) arg;
    uint64_t total = 0;
    p->rc = parse_blob(p->buf, p->n, p->eof_known, &p->sink, 0, &total, &p->odd);
    return 0;
}

static int parse_fasta_parallel(const char *buf, size_t n, int eof_known_at_end, fastx_t *x, int n_threads)
{
    enum { MAXP = 64 };
    piece_t pc[MAXP];
    int np = n_threads * 2 > MAXP ? MAXP : n_threads * 2, i, k = 0;
    size_t cut[MAXP + 1];
    if (n < (8u << 20) || np < 2 || buf[0] != '>') return -1;
    cut[0] = 0;
    for (i = 1; i < np; ++i) {                          /* the next line that starts with '>' after the i-th share */
        size_t from = n / np * i;
        const char *h = 0;
        if (from < cut[k]) from = cut[k];
        while (from < n && (h = (const char *) memchr(buf + from, '>', n - from)) != 0 && h[-1] != '\n') from = (size_t) (h - buf) + 1;
        if (!h || from >= n) break;
        if ((size_t) (h - buf) > cut[k]) cut[++k] = (size_t) (h - buf);
    }
    np = k + 1;
    cut[np] = n;
    for (int pass = 0; pass < 2; ++pass) {
        for (i = 0; i < np; ++i) {
            piece_t *p = &pc[i];
            if (pass == 0) {
                memset(p, 0, sizeof(*p));
                p->buf = buf + cut[i]; p->n = cut[i + 1] - cut[i];
                p->eof_known = i == np - 1 ? eof_known_at_end : 1;
                p->sink.count_only = 1;
            } else {
                p->sink.count_only = 0; p->sink.n = 0; p->sink.n_bases = 0;
            }
            pthread_create(&p->th, 0, piece_run, p);
        }
        for (i = 0; i < np; ++i) pthread_join(pc[i].th, 0);
        if (pass == 0) {
            uint64_t nr = 0, nb = 0;
            for (i = 0; i < np; ++i) { if (pc[i].odd || pc[i].rc) return -1; nr += pc[i].sink.n; nb += pc[i].sink.n_bases; }
            /* room for everything, then every piece gets its window into the final arrays */
            if (x->n_bases + nb + 1 > x->m_bases) {
                x->m_bases = x->n_bases + nb + 1; x->bases = (char *) realloc(x->bases, x->m_bases);
#ifdef MADV_HUGEPAGE
                if (x->m_bases >= (64u << 20)) {           /* gigabytes of fresh memory: 2 MB pages cut the page faults 512-fold */
                    const uintptr_t a0 = ((uintptr_t) x->bases + 0x1fffff) & ~(uintptr_t) 0x1fffff, a1 = ((uintptr_t) x->bases + x->m_bases) & ~(uintptr_t) 0x1fffff;
                    if (a1 > a0) madvise((void *) a0, a1 - a0, MADV_HUGEPAGE);
                }
#endif
            }
            if (x->n + nr + 2 > x->m) {
                x->m = x->n + nr + 2;
                x->off = (uint64_t *) realloc(x->off, (x->m + 1) * sizeof(uint64_t));
                x->names = (char **) realloc(x->names, x->m * sizeof(char *));
            }
            uint64_t r0 = x->n, b0 = x->n_bases;
            for (i = 0; i < np; ++i) {
                const uint64_t cr = pc[i].sink.n, cb = pc[i].sink.n_bases;
                pc[i].sink.bases = x->bases + b0; pc[i].sink.m_bases = 0;      /* fixed windows: never grown */
                pc[i].sink.off = x->off + r0; pc[i].sink.names = x->names + r0; pc[i].sink.m = 0;
                r0 += cr; b0 += cb;
            }
            /* the second pass writes into the windows; offsets are window-relative until the fix-up below */
        }
    }
    {
        uint64_t r0 = x->n, b0 = x->n_bases;
        for (i = 0; i < np; ++i) {
            for (uint64_t j = 1; j <= pc[i].sink.n; ++j) x->off[r0 + j] = b0 + pc[i].sink.off[j];
            r0 += pc[i].sink.n; b0 += pc[i].sink.n_bases;
        }
        x->n = r0; x->n_bases = b0;
    }
    return 0;
}

int fastx_load(const char *const *files, int n_files, uint64_t max_bases, fastx_t *x)
{
    memset(x, 0, sizeof(*x));
    x->m = 1024;
    x->off = (uint64_t *) malloc((x->m + 1) * sizeof(uint64_t));
    x->off[0] = 0;
    x->names = (char **) malloc(x->m * sizeof(char *));
    long nt = oatk_host_threads();
    if (getenv("OATK_FASTX_THREADS")) nt = atol(getenv("OATK_FASTX_THREADS"));      /* tests: force a thread count */
    if (nt > 16) nt = 16;
    uint64_t total = 0;
    for (int f = 0; f < n_files && !x->limit_reached; ++f) {
        blob_t b;
        if (blob_open(files[f], &b) != 0) {
            fprintf(stderr, "[E::%s] fail to open file \"%s\"\n", "make_kseq_stream", files[f]);   /* the reference exits here, sstream.c:46-49 */
            fastx_free(x);
            return FASTX_E_OPEN;
        }
        /* kseq learns about the end of file from a short read of its 16 KB buffer: when the size is a multiple of
         * that, one more (empty) read happens before "end of file" is known. Only visible in corner cases. */
        const int eof_known = (b.n % KS_BUFSIZE) != 0 || b.n == 0;
        const uint64_t n_before = x->n;
        if (b.n && parse_fasta_parallel(b.p, b.n, eof_known, x, (int) nt) == 0) {
            /* the cap is applied afterwards: keep the records up to the one that reaches it */
            for (uint64_t i = n_before; i < x->n; ++i) {
                total += x->off[i + 1] - x->off[i];
                if (max_bases && total >= max_bases) {
                    for (uint64_t j = i + 1; j < x->n; ++j) free(x->names[j]);
                    x->n = i + 1; x->n_bases = x->off[i + 1];
                    x->limit_reached = 1;
                    break;
                }
            }
        } else {
            sink_t k;
            int odd = 0;
            memset(&k, 0, sizeof(k));
            if (x->n_bases + b.n + 1 > x->m_bases) { x->m_bases = x->n_bases + b.n + 1; x->bases = (char *) realloc(x->bases, x->m_bases); }
            k.bases = x->bases; k.n_bases = x->n_bases; k.m_bases = x->m_bases;
            k.off = x->off; k.names = x->names; k.n = x->n; k.m = x->m;
            const int rc = parse_blob(b.p, b.n, eof_known, &k, max_bases, &total, &odd);
            x->bases = k.bases; x->n_bases = k.n_bases; x->m_bases = k.m_bases;
            x->off = k.off; x->names = k.names; x->n = k.n; x->m = k.m;
            if (rc == 1) x->limit_reached = 1;
            /* rc == -2: a truncated record ends this file; like sstream_read the next file is still opened */
        }
        blob_close(&b);
    }
    if (x->bases) x->bases[x->n_bases] = 0; else { x->bases = (char *) calloc(1, 1); x->m_bases = 1; }
    return 0;
}

void fastx_free(fastx_t *x)
{
    if (!x) return;
    for (uint64_t i = 0; i < x->n; ++i) free(x->names[i]);
    free(x->names); free(x->off); free(x->bases);
    memset(x, 0, sizeof(*x));
}

/* the reference's sr_read (syncmer.c:487-556) for files on disk: parse, then one pass of the device pipeline */
int sr_read_files(sr_db_t *sr_db, const char *const *files, int n_files, size_t max_bases)
{
    fastx_t x;
    oatk_tick(0);
    { const int lrc = fastx_load(files, n_files, max_bases, &x); if (lrc != 0) return lrc; }
    oatk_tick("reads: parse files");
    if (x.limit_reached)
        fprintf(stderr, "[M::%s] data limit (%lu) reached. Discard the remaining sequences...\n", "sr_read", (unsigned long) max_bases);
    const int rc = sr_read_mem(sr_db, x.bases, x.off, x.names, x.n);
    oatk_tick("reads: device pipeline + per-read blocks");
    fastx_free(&x);
    oatk_tick("reads: free the parse buffers");
    return rc;
}
