/*
 * sstream_gpu.c -- the reference's sequence-stream entry points for this path: sstream_open / sstream_close
 * (sstream.h:56-57, sstream.c:62-81) and sr_read (syncmer.h:120, syncmer.c:487-556), so that the three lines of
 * run_syncasm.c:79-86
 *     s_stream = sstream_open(file_in, n_file);  sr_read(s_stream, sr_db, m_data, n_threads);  sstream_close(s_stream);
 * compile and behave unchanged against this layer.
 *
 * sstream_t keeps the reference's public members (n_seq, files, n_files, n, s; sstream.h:46-51) in the same
 * order and types. What hangs off `s` differs: the reference keeps an open gz handle and a kseq buffer there and
 * hands out one record per sstream_read; here the records are cut out of whole mapped / inflated files by
 * fastx_gpu.c and go to the device in one pipeline run, so `s` only remembers that the first file could be opened
 * (the reference opens it in sstream_open and exits when it cannot, sstream.c:45-49 -- kept, message included).
 * sr_read consumes the stream from its current file to the end, like the reference's loop; n_seq and n are left
 * as the reference leaves them (records delivered, index of the last file).
 */
#include <stdio.h>
#include <stdlib.h>
#include "syncmer_gpu.h"
#include "fastx_gpu.h"

typedef struct { int opened; } stream_state_t;

sstream_t *sstream_open(char **files, int n_files)
{
    sstream_t *ss = (sstream_t *) calloc(1, sizeof(sstream_t));
    stream_state_t *st = (stream_state_t *) calloc(1, sizeof(stream_state_t));
    if (fastx_can_open(files[0]) != 0) {
        fprintf(stderr, "[E::%s] fail to open file \"%s\"\n", "make_kseq_stream", files[0]);
        exit(EXIT_FAILURE);
    }
    st->opened = 1;
    ss->n_seq = 0;
    ss->files = files;
    ss->n_files = n_files;
    ss->n = 0;
    ss->s = st;
    return ss;
}

void sstream_close(sstream_t *ss)
{
    if (!ss) return;
    free(ss->s);
    free(ss);
}

void sr_read(sstream_t *s_stream, sr_db_t *sr_db, size_t mD, int n_threads)
{
    (void) n_threads;            /* the device pipeline stands in for the reference's batches of 10 000 reads per thread */
    /* a file that cannot be opened ends the reference's process inside sstream_read (sstream.c:45-49); the same
     * happens to a device failure here: sr_read has no way to report either */
    if (sr_read_files(sr_db, (const char *const *) s_stream->files + s_stream->n, s_stream->n_files - s_stream->n, mD) != 0)
        exit(EXIT_FAILURE);
    s_stream->n_seq += sr_db->n;
    s_stream->n = s_stream->n_files - 1;
}
