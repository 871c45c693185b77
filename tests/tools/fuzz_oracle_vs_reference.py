#!/usr/bin/env python
"""Randomised check of the oracle (oracle/sync_oracle.c, the CPU restatement the GPU tests compare with) against the
unmodified reference compiled in place (oracle/_ref/libref.so): random k in (s, s + 2500], random s in [1, 31] (odd and
even), mixtures of HiFi-like reads (error rates up to 2 %), the adversarial set, short-period tandem arrays with
ambiguous bases; every sr_t field, the sr_db_stat figures, the syncmer database and the arc list of make_syncmer_graph (random
coverage thresholds) must be equal. Each case runs in a
forked child because the reference calls exit() on inputs it rejects ("identical kmers have different smers").
Test infrastructure, CPU only.

  python tests/tools/fuzz_oracle_vs_reference.py <first seed> <last seed + 1> [mean]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
from oatk_b200 import synth
import parity
from pyoracle import Oracle, Ref, pack_reads
oracle, ref = Oracle(), Ref()

MEAN = False          # third argument `mean`: also long homopolymers (past 255 and past 65535 bases), runs of N, lower case and
                      # IUPAC codes, empty and one-base reads


def make(seed):
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
    if MEAN:
        for _ in range(int(rng.integers(0, 4))):
            parts = []
            for _ in range(int(rng.integers(1, 12))):
                c = rng.random()
                if c < 0.25: parts.append(bytes([rng.choice(list(b"ACGTacgtN"))]) * int(rng.choice([1, 2, 254, 255, 256, 257, 300, 1000, 65535, 65536, 70000])))
                elif c < 0.5: parts.append(bytes(rng.choice(list(b"ACGTacgtNRYKMSWBDHVn-*"), int(rng.integers(1, 600))).tolist()))
                else: parts.append(bytes(rng.choice(list(b"ACGT"), int(rng.integers(1, 3000))).tolist()))
            reads.append(b"".join(parts))
        if rng.random() < 0.3: reads.insert(int(rng.integers(0, len(reads) + 1)), b"")
        if rng.random() < 0.3: reads.append(b"A")
    return k, s, reads

def one(seed, note=None):
    k, s, reads = make(seed)
    bases, off = pack_reads(reads)
    odb, of = oracle.extract(bases, off, k, s)
    rdb, rf = ref.extract(bases, off, k, s)
    d = parity.diff(of, rf, parity.EXTRACT_FIELDS)
    why = []
    if d: why.append(("extract", d[:3]))
    orc, od, oi, _, _ = oracle.stat(odb)
    rrc, rd, ri = ref.stat(rdb)
    if not np.array_equal(od, rd, equal_nan=True): why.append(("stat d", od.tolist(), rd.tolist()))
    elif not np.array_equal(oi, ri): why.append(("stat i", oi.tolist(), ri.tolist()))
    oc = oracle.collect(odb, len(reads))
    if oc is not None and oc["smer_conflict"] and note:
        note("conflict")                         # the reference is about to exit(1): tell the parent that this is expected
    rc = ref.collect(rdb)
    if oc is not None and oc["smer_conflict"]: why.append(("the oracle saw an s-mer conflict, the reference went on",))
    if (oc is None) != (rc is None): why.append(("collect none", oc is None, rc is None))
    elif oc is not None:
        dd = parity.diff(oc, rc, parity.SCM_FIELDS)
        if dd: why.append(("collect", dd[:3]))
        # a7: the arc list of make_syncmer_graph for a random (minimum k-mer coverage, arc coverage fraction)
        rng = np.random.default_rng(seed + 7)
        mkc, af = int(rng.choice([0, 1, 2, 3, 5])), float(rng.choice([0.0, 0.05, 0.35, 1.0]))
        g = ref.graph(rdb, rc, mkc, af)
        gd = ref.graph_dump(g)
        first = gd["vtx_lists"][np.concatenate([[0], np.cumsum(gd["vtx_n"])[:-1]]).astype(np.int64)] if len(gd["vtx_n"]) else np.zeros(0, np.uint64)
        ra = gd["arcs"]
        if len(ra):
            v = first[(ra[:, 0] >> 1).astype(np.int64)] | (ra[:, 0] & 1)
            w = first[(ra[:, 1] >> 1).astype(np.int64)] | (ra[:, 1] & 1)
            a = np.stack([v, w, ra[:, 4] & 0x3FFFFFFF, (ra[:, 4] >> 31) & 1], axis=1).astype(np.uint64)
            a[(a[:, 1] ^ 1) == a[:, 0], 3] = 0          # asmg_arc_fix_symm flips comp of v+ -> v- (tests/golden_util.py)
            a = a[np.lexsort((a[:, 2], a[:, 3], a[:, 1], a[:, 0]))]
        else:
            a = np.zeros((0, 4), np.uint64)
        oa = oracle.arcs(odb, oc, mkc, af)
        oa = oa[np.lexsort((oa[:, 2], oa[:, 3], oa[:, 1], oa[:, 0]))] if len(oa) else oa
        if oa.shape != a.shape or not np.array_equal(oa, a): why.append(("arcs", mkc, af, oa.shape, a.shape))
        ref.free(g=g)
    return k, s, len(reads), why


if __name__ == "__main__":
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    MEAN = len(sys.argv) > 3 and sys.argv[3] == "mean"
    bad = conflicts = 0
    for seed in range(lo, hi):
        r, w = os.pipe()
        pid = os.fork()
        if pid == 0:
            os.close(r)
            try:
                res = one(seed, lambda msg: os.write(w, (msg + "\n").encode()))
                os.write(w, ("result " + repr(res) + "\n").encode())
            except BaseException as e:
                os.write(w, ("raised " + repr(e)[:300] + "\n").encode())
            finally:
                os._exit(0)
        os.close(w)
        out = b""
        while True:
            c = os.read(r, 65536)
            if not c: break
            out += c
        os.close(r)
        _, st = os.waitpid(pid, 0)
        lines = out.decode().splitlines()
        last = lines[-1] if lines else ""
        if last.startswith("result "):
            res = eval(last[7:], {"nan": float("nan"), "inf": float("inf")})
            if res[3]:
                bad += 1
                print("DIVERGENCE seed", seed, res, flush=True)
        elif last == "conflict" and os.WIFEXITED(st) and os.WEXITSTATUS(st) == 1:
            conflicts += 1                         # both see identical k-mers with different s-mers; the reference exits there
            print("seed", seed, "s-mer conflict: the oracle flags it, the reference exits with 1", flush=True)
        else:
            bad += 1
            print("DIVERGENCE seed", seed, "child ended with status", st, "after", lines[-2:], flush=True)
    print("done", lo, hi, "bad", bad, "conflicts", conflicts)
