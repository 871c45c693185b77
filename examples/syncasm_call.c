/*
 * The smallest C caller of the host layer: what the reference's main() does after option parsing
 * (run_syncasm.c:429-431), with the reference's default values.
 *
 *   gcc -O2 -o syncasm_call examples/syncasm_call.c -Ioatk_b200/host -Iinclude \
 *       -Loatk_b200/host -loatk_gpu -Loatk_b200 -lsyncgpu -Wl,-rpath,$PWD/oatk_b200/host -Wl,-rpath,$PWD/oatk_b200
 *   ./syncasm_call out_prefix reads.fa[.gz] [more files]       ->  out_prefix.utg.gfa, out_prefix.utg.final.gfa
 */
#include <stdio.h>
#include "graph_gpu.h"

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <out prefix> <reads.fa|fq[.gz]> [...]\n", argv[0]); return 2; }
    /* -k 1001 -s 31 --max-bubble 100000 --max-tip 10000 -c 30 -a 0.35 --weak-cross 0.3, read EC on, 3 unzip rounds, -t 8 */
    int rc = syncasm(argv + 2, argc - 2, 0, 1001, 31, 100000, 10000, 30, 0.35, 0.3, 1, 3, 8, argv[1], 0, 0);
    oatk_gpu_shutdown();
    return rc;
}
