#!/usr/bin/env python
"""BASELINE.json configs[2] without read error correction: reads -> syncmers -> syncmer database -> graph
(-c, -a) -> unitigs -> GFA, once through the host layer over libsyncgpu (GPU) and once through the unmodified
reference on the host cores; the two GFA files must be byte-identical. Prints one JSON line with the stage times.

  python tests/tools/config3_run.py [--reads 20000] [--genome 1000000] [--c 30] [--k 1001] [--threads N]

Needs a CUDA device and oracle/_ref/libref.so (test infrastructure: it is the checker and the CPU baseline)."""
import argparse
import ctypes as C
import hashlib
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np                                   # noqa: E402
from oatk_b200 import synth                          # noqa: E402
from oatk_b200.host import build_host                # noqa: E402
from pyoracle import Ref, pack_reads                 # noqa: E402


class SrDb(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.c_void_p), ("k", C.c_int), ("s", C.c_int), ("stats", C.c_void_p)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=20000)
    ap.add_argument("--genome", type=int, default=1000000)
    ap.add_argument("--len", type=int, default=15000)
    ap.add_argument("--err", type=float, default=1e-3)
    ap.add_argument("--k", type=int, default=1001)
    ap.add_argument("--s", type=int, default=31)
    ap.add_argument("--c", type=int, default=30)
    ap.add_argument("--a", type=float, default=0.35)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--ec", action="store_true", help="with read error correction (the reference's default); without: --no-read-ec")
    args = ap.parse_args()

    reads = synth.hifi_reads(2, args.genome, args.reads, args.len, args.err)
    bases, off = pack_reads(reads)
    total = int(off[-1])
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)

    # ---- ours -------------------------------------------------------------------------------------
    H = C.CDLL(build_host.build())
    H.sr_read_mem.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    H.sr_read_files.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.c_size_t]
    H.sr_db_init.argtypes = [C.c_void_p, C.c_int, C.c_int]
    H.sr_db_stat.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    H.collect_syncmer_from_reads.restype = C.c_void_p
    H.collect_syncmer_from_reads.argtypes = [C.c_void_p]
    H.make_syncmer_graph.restype = C.c_void_p
    H.make_syncmer_graph.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double]
    H.process_mergeable_unitigs.argtypes = [C.c_void_p]
    H.scg_consensus.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    H.scg_consensus.restype = None
    H.scg_destroy.argtypes = [C.c_void_p]
    H.read_error_correction.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_void_p, C.c_int]
    H.read_error_correction.restype = None
    H.syncmer_db_destroy.argtypes = [C.c_void_p]
    H.sr_db_clean.argtypes = [C.c_void_p]

    fa = os.path.join(tmp, "reads.fa")
    with open(fa, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">r%d\n" % i)
            f.write(r)
            f.write(b"\n")
    fa_arr = (C.c_char_p * 1)(fa.encode())

    def ours():
        t = {}
        db = SrDb()
        H.sr_db_init(C.byref(db), args.k, args.s)
        t0 = time.perf_counter()
        assert H.sr_read_files(C.byref(db), fa_arr, 1, 0) == 0           # parse the FASTA + device pipeline + per-read blocks
        t["sr_read_s"] = time.perf_counter() - t0
        nul = libc.fopen(b"/dev/null", b"w")
        t0 = time.perf_counter()
        H.sr_db_stat(C.byref(db), nul, 0)
        t["sr_db_stat_s"] = time.perf_counter() - t0
        libc.fclose(nul)
        t0 = time.perf_counter()
        scm = H.collect_syncmer_from_reads(C.byref(db))
        t["collect_s"] = time.perf_counter() - t0
        assert scm
        if args.ec:
            t0 = time.perf_counter()
            g = H.make_syncmer_graph(C.byref(db), scm, 0, 0.0)
            t["ec_graph_s"] = time.perf_counter() - t0
            t["ec_consensus_s"] = 0.0                 # computed inside read_error_correction, for the surviving graph only
            t0 = time.perf_counter()
            H.read_error_correction(C.byref(db), g, 0.02, args.c, args.c * 10, args.c, args.a, args.threads, None, 0)
            t["read_ec_s"] = time.perf_counter() - t0
            nul = libc.fopen(b"/dev/null", b"w")
            t0 = time.perf_counter()
            H.sr_db_stat(C.byref(db), nul, 0)
            t["sr_db_stat2_s"] = time.perf_counter() - t0
            libc.fclose(nul)
            H.scg_destroy(g)
        t0 = time.perf_counter()
        g = H.make_syncmer_graph(C.byref(db), scm, args.c, args.a)
        t["graph_s"] = time.perf_counter() - t0
        assert g
        t0 = time.perf_counter()
        H.process_mergeable_unitigs(g)
        t["unitig_s"] = time.perf_counter() - t0
        path = os.path.join(tmp, "ours.utg.gfa")
        fo = libc.fopen(path.encode(), b"w")
        t0 = time.perf_counter()
        H.scg_consensus(C.byref(db), g, 0, 0, fo)
        t["consensus_gfa_s"] = time.perf_counter() - t0
        libc.fclose(fo)
        H.scg_destroy(g)
        H.syncmer_db_destroy(scm)
        H.sr_db_clean(C.byref(db))
        t["total_s"] = sum(t.values())
        return t, path

    ours()                                            # warm-up: context creation, first allocations
    t_ours, p_ours = ours()

    # ---- the unmodified reference -------------------------------------------------------------------
    R = Ref()
    R.L.ref_write_gfa.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
    t_ref = {}
    t0 = time.perf_counter()
    rdb = R.L.ref_extract_file(fa.encode(), args.k, args.s, args.threads, 0)       # sr_read: kseq parse + extract, -t threads
    t_ref["sr_read_s"] = time.perf_counter() - t0
    R.L.sr_db_stat.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    nul = libc.fopen(b"/dev/null", b"w")
    t0 = time.perf_counter()
    R.L.sr_db_stat(rdb, nul, 0)
    t_ref["sr_db_stat_s"] = time.perf_counter() - t0
    libc.fclose(nul)
    t0 = time.perf_counter()
    rscm = R.L.ref_collect(rdb)
    t_ref["collect_s"] = time.perf_counter() - t0
    if args.ec:
        R.L.ref_write_gfa2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        R.L.ref_read_ec.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int]
        t0 = time.perf_counter()
        g = R.L.ref_make_graph(rdb, rscm, 0, 0.0)
        t_ref["ec_graph_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        assert R.L.ref_write_gfa2(rdb, g, 1, 1, b"/dev/null") == 0
        t_ref["ec_consensus_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        R.L.ref_read_ec(rdb, g, 0.02, args.c, args.c * 10, args.c, args.a, args.threads)
        t_ref["read_ec_s"] = time.perf_counter() - t0
        nul = libc.fopen(b"/dev/null", b"w")
        t0 = time.perf_counter()
        R.L.sr_db_stat(rdb, nul, 0)
        t_ref["sr_db_stat2_s"] = time.perf_counter() - t0
        libc.fclose(nul)
        R.L.ref_scg_free(g)
    t0 = time.perf_counter()
    g = R.L.ref_make_graph(rdb, rscm, args.c, args.a)
    t_ref["graph_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    R.L.ref_unitig(g)
    t_ref["unitig_s"] = time.perf_counter() - t0
    p_ref = os.path.join(tmp, "ref.utg.gfa")
    t0 = time.perf_counter()
    assert R.L.ref_write_gfa(rdb, g, p_ref.encode()) == 0
    t_ref["consensus_gfa_s"] = time.perf_counter() - t0
    t_ref["total_s"] = sum(t_ref.values())

    a, b = open(p_ours, "rb").read(), open(p_ref, "rb").read()
    line = {"config": "BASELINE.json configs[2] (%s): %d x %d b reads of a %d b genome, k=%d s=%d -c %d -a %.2f" % (
                "default: with read error correction" if args.ec else "--no-read-ec", args.reads, args.len, args.genome, args.k, args.s, args.c, args.a),
            "raw_bases": total, "gfa_identical": a == b, "gfa_md5": hashlib.md5(a).hexdigest(), "gfa_md5_reference": hashlib.md5(b).hexdigest(),
            "gfa_S_lines": a.count(b"\nS\t"), "gfa_L_lines": a.count(b"\nL\t"), "gfa_bytes": len(a),
            "ours_s": t_ours, "reference_s": t_ref, "reference_threads": args.threads,
            "speedup_total": t_ref["total_s"] / t_ours["total_s"],
            "note": "both arms read the same FASTA in /dev/shm; ours: sr_read_files (native parser, f4) + host layer over libsyncgpu; reference: sr_read -t threads"}
    print(json.dumps(line))
    for p in (p_ours, p_ref, fa):
        os.unlink(p)
    os.rmdir(tmp)
    return 0 if a == b else 1


if __name__ == "__main__":
    sys.exit(main())
