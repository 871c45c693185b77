#!/usr/bin/env python
"""Random-parameter hunt on the GPU: the device path (sg_extract, sg_stat, sg_count, sg_arcs through the C ABI) against the
CPU oracle on the read sets and (k, s) pairs of tests/tools/fuzz_oracle_vs_reference.py (k - s in [1, 2500], s in [1, 31] odd and
even, HiFi-like reads, the adversarial set, short-period tandem arrays with ambiguous bases); every sr_t field, the
multiplicity tables, the syncmer database and the arc list must be equal, an s-mer conflict must be reported as one.
Needs a CUDA device; the oracle is the checker (test infrastructure).

  python tests/tools/fuzz_extract_gpu.py <first seed> <last seed + 1> [seconds]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np                                   # noqa: E402
import parity                                        # noqa: E402
from oatk_b200 import lib, synth                     # noqa: E402
from pyoracle import Oracle, pack_reads              # noqa: E402


def make(seed):                                      # the generator of tests/tools/fuzz_oracle_vs_reference.py
    rng = np.random.default_rng(seed)
    s = int(rng.integers(1, 32))
    k = int(s + rng.integers(1, 400)) if rng.random() < 0.7 else int(s + rng.integers(1, 2500))
    reads = []
    if rng.random() < 0.5:
        reads += synth.adversarial_reads(int(rng.integers(0, 1000)), k, s)
    reads += synth.hifi_reads(int(rng.integers(0, 1000)), int(rng.integers(2000, 40000)), int(rng.integers(1, 20)), int(rng.integers(50, 9000)), float(rng.choice([0, 0.0005, 0.003, 0.02])))
    for _ in range(int(rng.integers(0, 4))):
        unit = bytes(rng.choice(list(b"ACGT"), int(rng.integers(1, 40))).tolist())
        r = bytearray(unit * int(rng.integers(5, 400)))
        for p in rng.integers(0, len(r), int(rng.integers(0, 5))): r[p] = ord("N")
        reads.append(bytes(r))
    return k, s, reads


def one(ctx, oracle, seed):
    k, s, reads = make(seed)
    bases, off = pack_reads(reads)
    db, exp = oracle.extract(bases, off, k, s)
    why = []
    b = lib.Batch(ctx)
    b.set_reads_host(bases, off)
    b.extract(k, s)
    d = parity.diff(b.extract_download(), exp, parity.EXTRACT_FIELDS)
    if d: why.append(("extract", d[:3]))
    rc, dd, ii, sc, kc = oracle.stat(db)
    oc = oracle.collect(db, len(reads))
    if rc == 0:
        st = b.stat()
        if not (np.array_equal(np.array(st.smer_cnts[:], np.int64), sc) and np.array_equal(np.array(st.kmer_cnts[:], np.int64), kc)):
            why.append(("stat tables",))
    if oc is None:
        try:
            b.count(); why.append(("count went on where the oracle has nothing",))
        except lib.SgError:
            pass
    elif oc["smer_conflict"]:
        try:
            b.count(); why.append(("no conflict reported",))
        except lib.SgError as e:
            if e.code != -6: why.append(("conflict code", e.code))
    else:
        b.count()
        d = parity.diff(b.count_download(), oc, parity.SCM_FIELDS)
        if d: why.append(("count", d[:3]))
        rng = np.random.default_rng(seed + 7)
        mkc, af = int(rng.choice([0, 1, 2, 3, 5])), float(rng.choice([0.0, 0.05, 0.35, 1.0]))
        ga, oa = b.arcs(mkc, af), oracle.arcs(db, oc, mkc, af)
        if ga.shape != oa.shape or not np.array_equal(ga, oa): why.append(("arcs", mkc, af, ga.shape, oa.shape))
    b.close()
    oracle.free(db, oc)
    return k, s, len(reads), why


if __name__ == "__main__":
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    budget = float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
    ctx, oracle = lib.Context(0), Oracle()
    t0, bad, n = time.time(), 0, 0
    for seed in range(lo, hi):
        if time.time() - t0 > budget:
            break
        res = one(ctx, oracle, seed)
        n += 1
        if res[3]:
            bad += 1
            print("DIVERGENCE seed", seed, res, flush=True)
        if n % 10 == 0:                                # a run cut short by a time limit still says how far it got
            print("... %d cases, %d divergent, %.1f s" % (n, bad, time.time() - t0), flush=True)
    print("done: seeds %d:%d, %d cases, %d divergent, %.1f s" % (lo, lo + n, n, bad, time.time() - t0))
