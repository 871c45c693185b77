/*
 * fastx_gpu.h -- FASTA / FASTQ (plain or gzip) files into the flat arrays sr_read_mem takes, with the record
 * grammar of the reference's kseq / sstream reader (row f4 of SURVEY.md section 8; see fastx_gpu.c).
 */
#ifndef FASTX_GPU_H
#define FASTX_GPU_H
#include <stdint.h>
#include <stddef.h>
#include "syncmer_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    char *bases;            /* all sequences back to back, as they stand in the file (case, IUPAC codes kept) */
    uint64_t *off;          /* n + 1 offsets into bases */
    char **names;           /* n names (up to the first white space of the header) */
    uint64_t n, m;          /* records, capacity */
    uint64_t n_bases, m_bases;
    int limit_reached;      /* the -D cap stopped the reading */
} fastx_t;

/* reads the files in order; max_bases = 0: no limit. Returns 0, or FASTX_E_OPEN when a file cannot be opened (after the
 * reference's own line for that, "[E::make_kseq_stream] fail to open file ...", sstream.c:46-49, where the reference exits) */
#define FASTX_E_OPEN (-100)
int fastx_load(const char *const *files, int n_files, uint64_t max_bases, fastx_t *out);
void fastx_free(fastx_t *x);
/* 0 when the file can be opened for reading the way fastx_load would */
int fastx_can_open(const char *path);
/* sr_read for files: fastx_load + sr_read_mem, printing the reference's data-limit message; FASTX_E_OPEN or sr_read_mem's code */
int sr_read_files(sr_db_t *sr_db, const char *const *files, int n_files, size_t max_bases);

#ifdef __cplusplus
}
#endif
#endif
