#!/usr/bin/env python
"""Randomised hunt for divergences in the WHOLE command on the GPU: the scenarios of tests/tools/fuzz_graph_stage.py (random
mixtures of repeats, haplotypes, rare molecules, recombinants and tandem arrays; random k / s, coverage thresholds, clean-up
limits, with and without read error correction and unzipping) through this repository's syncasm() -- device extraction,
counting, arc tally, error filter, graph search, votes, run-length sums -- against the unmodified reference's syncasm() on the
host; both GFA files must be byte-identical. Prints one line per seed and a JSON summary.

  python tests/tools/fuzz_syncasm_gpu.py [--seeds 0:40] [--hifi]

Needs a CUDA device and oracle/_ref/libref.so (test infrastructure: the checker)."""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "tools")):
    sys.path.insert(0, p)
import numpy as np                                   # noqa: E402
from oatk_b200.host import build_host                # noqa: E402
from pyoracle import Ref                             # noqa: E402
from test_alignment_cpu import _sample               # noqa: E402
from fuzz_graph_stage import random_genomes          # noqa: E402

PROTO = [C.POINTER(C.c_char_p), C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
         C.c_double, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_int]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", default="0:40")
    ap.add_argument("--hifi", action="store_true", help="k = 1001, s = 31 and 12-20 kb reads (the reference's defaults)")
    args = ap.parse_args()
    lo, hi = (int(x) for x in args.seeds.split(":"))
    R = Ref().L
    H = C.CDLL(build_host.build())
    for L in (R, H):
        L.syncasm.restype = C.c_int
        L.syncasm.argtypes = PROTO
    H.oatk_ec_last_run.argtypes = [C.POINTER(C.c_uint64)]
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    bad, n_ok, n_skip, n_dev = [], 0, 0, 0
    for seed in range(lo, hi):
        rng = np.random.default_rng(1000 + seed)
        k, s = [(101, 11), (201, 15), (301, 21), (501, 31), (151, 13)][int(rng.integers(0, 5))]
        if args.hifi:
            k, s = 1001, 31
        genomes = random_genomes(rng, k)
        n = int(rng.integers(400, 1600))
        L1, L2 = int(rng.integers(6000, 14000)), int(rng.integers(1500, 6000))
        if args.hifi:
            n, L1, L2 = int(rng.integers(500, 1200)), int(rng.integers(12000, 20000)), int(rng.integers(6000, 12000))
        err = float(rng.choice([0.0001, 0.0003, 0.001]))
        reads = _sample(rng, genomes, n // 2, L1, err) + _sample(rng, genomes, n - n // 2, L2, err)
        mkc, af = int(rng.integers(2, 6)), float(rng.choice([0.0, 0.05, 0.2, 0.35]))
        ec, unzip = int(rng.integers(0, 2)), int(rng.integers(0, 4))
        bubble, tip, weak = int(rng.choice([1000, 20000, 100000])), int(rng.choice([500, 3000, 10000])), float(rng.choice([0.2, 0.3, 0.5]))
        fa = os.path.join(tmp, "r.fa")
        with open(fa, "wb") as f:
            for i, r in enumerate(reads):
                f.write(b">r%d\n%s\n" % (i, r))
        files = (C.c_char_p * 1)(fa.encode())
        p_ref, p_ours = os.path.join(tmp, "ref"), os.path.join(tmp, "ours")
        for suffix in (".utg.gfa", ".utg.final.gfa"):
            for p in (p_ref, p_ours):
                if os.path.exists(p + suffix):
                    os.unlink(p + suffix)
        desc = dict(seed=seed, k=k, s=s, n=n, L=(L1, L2), err=err, mkc=mkc, af=af, ec=ec, unzip=unzip, bubble=bubble, tip=tip, weak=weak)
        sys.stderr.flush()
        rc_ref = R.syncasm(files, 1, 0, k, s, bubble, tip, mkc, af, weak, ec, unzip, 4, p_ref.encode(), None, 0)
        rc = H.syncasm(files, 1, 0, k, s, bubble, tip, mkc, af, weak, ec, unzip, 4, p_ours.encode(), None, 0)
        if ec:
            n_dev += H.oatk_ec_last_run(None) == 2
        if rc_ref != 0 or rc != 0:
            ok = rc_ref == rc or (rc_ref != 0 and rc != 0)
            print("skip (empty graph on both sides)" if ok else "BAD (return codes %d / %d)" % (rc_ref, rc), desc, flush=True)
            n_skip += ok
            if not ok:
                bad.append(desc)
            continue
        res = {}
        for suffix in (".utg.gfa", ".utg.final.gfa"):
            a, b = open(p_ours + suffix, "rb").read(), open(p_ref + suffix, "rb").read()
            res[suffix] = (a == b, b.count(b"\nS\t"), b.count(b"\nL\t"))
        ok = all(v[0] for v in res.values())
        print("ok " if ok else "BAD", desc, res, flush=True)
        n_ok += ok
        if not ok:
            bad.append(desc)
    print(json.dumps({"seeds": "%d:%d" % (lo, hi), "hifi": bool(args.hifi), "identical": n_ok, "both_empty": n_skip, "divergent": len(bad),
                      "error_correction_runs_on_the_device_without_a_host_graph": n_dev, "divergences": bad}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
