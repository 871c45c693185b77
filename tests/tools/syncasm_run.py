#!/usr/bin/env python
"""BASELINE.json configs[2] through the WHOLE command: FASTA -> <out>.utg.gfa and <out>.utg.final.gfa, once by the host
layer's syncasm() (device rows on the GPU) and once by the unmodified reference's syncasm() on the host cores, same
arguments (the reference's defaults: -k 1001 -s 31 -a 0.35, read error correction, 3 unzip rounds, --max-bubble 100000
--max-tip 10000 --weak-cross 0.3). Both files must be byte-identical. Prints one JSON line.

  python tests/tools/syncasm_run.py [--reads 20000] [--genome 1000000] [--c 30] [--threads N]

Needs a CUDA device and oracle/_ref/libref.so (test infrastructure: the checker and the CPU baseline)."""
import argparse
import ctypes as C
import hashlib
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from oatk_b200 import synth                          # noqa: E402
from oatk_b200.host import build_host                # noqa: E402
from pyoracle import Ref                             # noqa: E402
import make_golden_syncasm as G                      # noqa: E402


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
    ap.add_argument("--no-read-ec", action="store_true")
    ap.add_argument("--unzip-round", type=int, default=3)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()

    reads = synth.hifi_reads(2, args.genome, args.reads, args.len, args.err)
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fa = os.path.join(tmp, "reads.fa")
    with open(fa, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">r%d\n%s\n" % (i, r))
    a = dict(k=args.k, s=args.s, mkc=args.c, af=args.a, ec=0 if args.no_read_ec else 1, unzip=args.unzip_round, bubble=100000, tip=10000, weak=0.3)

    H = C.CDLL(build_host.build())
    G.bind(H)
    p_ours, p_ref = os.path.join(tmp, "ours"), os.path.join(tmp, "ref")
    assert G.run(H, fa, a, p_ours, args.threads) == 0          # warm-up: context creation, first allocations
    t0 = time.perf_counter()
    assert G.run(H, fa, a, p_ours, args.threads) == 0
    t_ours = time.perf_counter() - t0
    R = Ref().L
    G.bind(R)
    t0 = time.perf_counter()
    assert G.run(R, fa, a, p_ref, args.threads) == 0
    t_ref = time.perf_counter() - t0

    out = {"config": "BASELINE.json configs[2], whole command: %d x %d b reads of a %d b genome, -k %d -s %d -c %d -a %.2f%s --unzip-round %d" % (
               args.reads, args.len, args.genome, args.k, args.s, args.c, args.a, " --no-read-ec" if args.no_read_ec else "", args.unzip_round),
           "raw_bases": sum(len(r) for r in reads), "ours_s": t_ours, "reference_s": t_ref, "reference_threads": args.threads, "speedup": t_ref / t_ours}
    ok = True
    for suffix in (".utg.gfa", ".utg.final.gfa"):
        x, y = open(p_ours + suffix, "rb").read(), open(p_ref + suffix, "rb").read()
        out[suffix] = {"identical": x == y, "md5": hashlib.md5(x).hexdigest(), "md5_reference": hashlib.md5(y).hexdigest(),
                       "S": x.count(b"\nS\t"), "L": x.count(b"\nL\t"), "bytes": len(x)}
        ok &= x == y
        os.unlink(p_ours + suffix)
        os.unlink(p_ref + suffix)
    print(json.dumps(out))
    os.unlink(fa)
    os.rmdir(tmp)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
