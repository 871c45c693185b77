/*
 * syncasm_main.c -- the `syncasm` command over liboatk_gpu.so / libsyncgpu.so.
 *
 * Command-line contract of the reference's run_syncasm.c:329-454: same option letters and long names
 * (-k -s -c -a -D -t -o -v, --max-bubble --max-tip --weak-cross --unzip-round --no-read-ec --threads
 * --verbose --version --help), same defaults (:356-367), same usage text, same messages and exit codes
 * ([E::main] unknown / missing option -> 1, usage on stderr -> 1, on stdout with -h -> 0, a failed run ->
 * "[E::main] failed to constrcut assembly" and EXIT_FAILURE, the Version / CMD / Real time lines at the end).
 *
 * The reference parses with klib's ketopt in permuting mode. The scanner below is written from ketopt's
 * documented behaviour rather than from its code, but keeps every observable habit of it, because scripts
 * and the tests of this repository compare stderr with the reference binary's:
 *   - options may follow file names; a lone "-" is a file name; "--" ends the options
 *   - long options may be shortened to any prefix that fits exactly one table entry ("--max-b 5"); a prefix
 *     that fits two ("--ver") is an unknown option; "--name=value" and "--name value" are both accepted
 *   - short options cluster ("-Vh"), an argument may be attached ("-k501") or follow
 *   - the token quoted in an error message is whatever stands in front of the scanner's cursor once the
 *     offending token has been stepped over and moved in front of the file names seen so far; after
 *     "reads.fa --bogus" that is "reads.fa", after "-xk 5" (an unknown letter that is not the last of its
 *     cluster) it is the token before the cluster. Quirks, but the reference's output.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <sys/resource.h>
#include "../graph_gpu.h"
#include "../fastx_gpu.h"

#define SYNCASM_GPU_VERSION "1.0"        /* reference version.h:34: the version string scripts check */

enum { OPT_MAX_BUBBLE = 301, OPT_MAX_TIP, OPT_WEAK_CROSS, OPT_UNZIP_ROUND, OPT_NO_READ_EC };
enum { ARG_NONE = 0, ARG_NEEDED = 1 };
#define RET_DONE (-1)

typedef struct { const char *name; int arg; int code; } long_opt_t;

static const long_opt_t LONG_OPTS[] = {          /* run_syncasm.c:329-340, same order (prefix matching sees all) */
    { "max-bubble", ARG_NEEDED, OPT_MAX_BUBBLE }, { "max-tip", ARG_NEEDED, OPT_MAX_TIP },
    { "weak-cross", ARG_NEEDED, OPT_WEAK_CROSS }, { "unzip-round", ARG_NEEDED, OPT_UNZIP_ROUND },
    { "no-read-ec", ARG_NONE, OPT_NO_READ_EC }, { "threads", ARG_NEEDED, 't' }, { "verbose", ARG_NEEDED, 'v' },
    { "version", ARG_NONE, 'V' }, { "help", ARG_NONE, 'h' }, { NULL, 0, 0 } };
static const char SHORT_OPTS[] = "k:s:c:a:D:t:v:o:Vh";

typedef struct {
    int argc;
    char **argv;
    int cur;            /* token under the cursor */
    int in_cluster;     /* next letter of a short-option cluster, 0 = at the start of a token */
    int waiting;        /* file names the cursor has passed; they sit right in front of it */
    char *value;        /* argument of the option just returned */
    int first_file;     /* valid after RET_DONE */
} scanner_t;

static int is_file_name(const char *t) { return t[0] != '-' || t[1] == '\0'; }

/* token at `from` jumps in front of the `n` tokens before it */
static void hop_left(char **argv, int from, int n)
{
    char *t = argv[from];
    memmove(argv + from - n + 1, argv + from - n, (size_t) n * sizeof(char *));
    argv[from - n] = t;
}

/* one option per call: its code, '?' (unknown / ambiguous), ':' (argument missing) or RET_DONE */
static int next_option(scanner_t *z)
{
    char **argv = z->argv;
    int code, start, done_with_token = 0, j;
    while (z->cur < z->argc && is_file_name(argv[z->cur])) { ++z->cur; ++z->waiting; }
    z->value = NULL;
    start = z->cur;
    if (z->cur >= z->argc) { z->first_file = z->cur - z->waiting; return RET_DONE; }
    if (argv[z->cur][1] == '-') {
        const char *name = argv[z->cur] + 2;
        size_t len = strcspn(name, "=");
        const long_opt_t *hit = NULL, *o;
        int hits = 0;
        if (*name == '\0') {                               /* "--": the rest are file names */
            hop_left(argv, z->cur, z->waiting);
            ++z->cur;
            z->first_file = z->cur - z->waiting;
            return RET_DONE;
        }
        for (o = LONG_OPTS; o->name; ++o)
            if (strncmp(name, o->name, len) == 0) { ++hits; hit = o; }
        code = '?';
        if (hits == 1) {
            code = hit->code;
            if (name[len] == '=') z->value = (char *) name + len + 1;
            else if (hit->arg == ARG_NEEDED) {
                if (z->cur + 1 < z->argc) z->value = argv[++z->cur];
                else code = ':';
            }
        }
        done_with_token = 1;
    } else {
        const char *spec;
        if (z->in_cluster == 0) z->in_cluster = 1;
        code = (unsigned char) argv[z->cur][z->in_cluster++];
        spec = strchr(SHORT_OPTS, code);
        if (!spec) code = '?';
        else if (spec[1] == ':') {
            if (argv[z->cur][z->in_cluster] != '\0') z->value = argv[z->cur] + z->in_cluster;
            else if (z->cur + 1 < z->argc) z->value = argv[++z->cur];
            else code = ':';
            done_with_token = 1;
        }
        if (!done_with_token && argv[z->cur][z->in_cluster] == '\0') done_with_token = 1;
    }
    if (done_with_token) {
        ++z->cur;
        z->in_cluster = 0;
        for (j = start; j < z->cur && z->waiting > 0; ++j) hop_left(argv, j, z->waiting);
    }
    return code;
}

static double wall_now(void)
{
    struct timeval t;
    gettimeofday(&t, NULL);
    return t.tv_sec + 1e-6 * t.tv_usec;
}

static void usage_text(FILE *fo, int k, int s, int c, double a, size_t D, int t, const char *o, int bubble, int tip,
        double weak, int unzip, int verbose)
{
    fprintf(fo, "\n");
    fprintf(fo, "Usage: syncasm [options] <target.fa[stq][.gz]> [...]\n");
    fprintf(fo, "Options:\n");
    fprintf(fo, "    -k INT               kmer size [%d]\n", k);
    fprintf(fo, "    -s INT               smer size (no larger than 31) [%d]\n", s);
    fprintf(fo, "    -c INT               minimum kmer coverage [%d]\n", c);
    fprintf(fo, "    -a FLOAT             minimum arc coverage [%.2f]\n", a);
    fprintf(fo, "    -D INT               maximum amount of data to use; suffix K/M/G recognized [%lu]\n", (unsigned long) D);
    fprintf(fo, "    -t INT               number of threads [%d]\n", t);
    fprintf(fo, "    -o FILE              prefix of output files [%s]\n", o);
    fprintf(fo, "    --max-bubble  INT    maximum bubble size for assembly graph clean [%d]\n", bubble);
    fprintf(fo, "    --max-tip     INT    maximum tip size for assembly graph clean [%d]\n", tip);
    fprintf(fo, "    --weak-cross  FLOAT  maximum relative edge coverage for weak crosslink clean [%.2f]\n", weak);
    fprintf(fo, "    --unzip-round INT    maximum round of assembly graph unzipping [%d]\n", unzip);
    fprintf(fo, "    --no-read-ec         do not do read error correction\n");
    fprintf(fo, "    -v INT               verbose level [%d]\n", verbose);
    fprintf(fo, "    --version            show version number\n");
    fprintf(fo, "\n");
    fprintf(fo, "Example: ./syncasm -k 1001 -c 50 -t 8 -o syncasm.asm hifi.fa.gz\n\n");
}

int main(int argc, char *argv[])
{
    /* defaults of run_syncasm.c:356-367 */
    int k = 1001, s = 31, min_k_cov = 3, n_threads = 1, bubble = 100000, tip = 10000, do_ec = 1, unzip = 3, verbose = 0;
    double min_a_cov_f = .35, weak_cross = 0.3;
    size_t m_data = 0;
    char *out = "syncasm.asm";
    FILE *help_to = stderr;
    scanner_t z = { argc, argv, 1, 0, 0, NULL, 1 };
    const double t_start = wall_now();
    struct rusage ru;
    int c, i, ret;

    while ((c = next_option(&z)) != RET_DONE) {
        switch (c) {
            case 'k': k = atoi(z.value); break;
            case 's': s = atoi(z.value); break;
            case 'c': min_k_cov = atoi(z.value); break;
            case 'a': min_a_cov_f = atof(z.value); break;
            case 'D': {
                char *unit;
                m_data = strtol(z.value, &unit, 0);
                if (*unit == 'k' || *unit == 'K') m_data <<= 10;
                else if (*unit == 'm' || *unit == 'M') m_data <<= 20;
                else if (*unit == 'g' || *unit == 'G') m_data <<= 30;
                break;
            }
            case 't': n_threads = atoi(z.value); break;
            case OPT_MAX_BUBBLE: bubble = atoi(z.value); break;
            case OPT_MAX_TIP: tip = atoi(z.value); break;
            case OPT_WEAK_CROSS: weak_cross = atof(z.value); break;
            case OPT_UNZIP_ROUND: unzip = atoi(z.value); break;
            case OPT_NO_READ_EC: do_ec = 0; break;
            case 'o': if (strcmp(z.value, "-") != 0) out = z.value; break;
            case 'v': verbose = atoi(z.value); break;
            case 'V': puts(SYNCASM_GPU_VERSION); return 0;
            case 'h': help_to = stdout; break;
            case '?': fprintf(stderr, "[E::%s] unknown option: \"%s\"\n", __func__, argv[z.cur - 1]); return 1;
            case ':': fprintf(stderr, "[E::%s] missing option: \"%s\"\n", __func__, argv[z.cur - 1]); return 1;
        }
    }

    if (argc == z.first_file || help_to == stdout) {
        usage_text(help_to, k, s, min_k_cov, min_a_cov_f, m_data, n_threads, out, bubble, tip, weak_cross, unzip, verbose);
        return help_to == stdout ? 0 : 1;
    }

    ret = syncasm(argv + z.first_file, argc - z.first_file, m_data, k, s, bubble, tip, min_k_cov, min_a_cov_f, weak_cross,
                  do_ec, unzip, n_threads, out, 0, verbose);
    if (ret == 2) exit(EXIT_FAILURE);              /* the reference left from inside process_kmer_cluster: nothing more is printed */
    if (ret) {
        fprintf(stderr, "[E::%s] failed to constrcut assembly\n", __func__);
        exit(EXIT_FAILURE);
    }
    if (fflush(stdout) == EOF) {
        fprintf(stderr, "[E::%s] failed to write the results\n", __func__);
        exit(EXIT_FAILURE);
    }
    if (verbose >= 0) {
        getrusage(RUSAGE_SELF, &ru);
        fprintf(stderr, "[M::%s] Version: %s\n", __func__, SYNCASM_GPU_VERSION);
        fprintf(stderr, "[M::%s] CMD:", __func__);
        for (i = 0; i < argc; ++i) fprintf(stderr, " %s", argv[i]);
        fprintf(stderr, "\n[M::%s] Real time: %.3f sec; CPU: %.3f sec; Peak RSS: %.3f GB\n", __func__, wall_now() - t_start,
                ru.ru_utime.tv_sec + ru.ru_stime.tv_sec + 1e-6 * (ru.ru_utime.tv_usec + ru.ru_stime.tv_usec),
                ru.ru_maxrss * 1024.0 / 1024.0 / 1024.0 / 1024.0);
    }
    return 0;
}
