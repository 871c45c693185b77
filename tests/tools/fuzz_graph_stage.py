#!/usr/bin/env python
"""Randomised hunt for divergences in the graph stage of syncasm() (host code, no GPU): random mixtures of repeats,
haplotypes, rare molecules and recombinants, random k / coverage thresholds / clean-up limits; the reference's whole
command against this layer's read error correction and graph stage, fed with the reference's extraction, database and
graph construction (the device rows, which have their own GPU tests); both GFA files compared.

  python tests/tools/fuzz_graph_stage.py [--seeds 0:40] [--threads 3]

Needs oracle/_ref/libref.so (test infrastructure)."""
import argparse
import ctypes as C
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
os.environ.setdefault("OATK_PF_MIN", "1")
import numpy as np                                   # noqa: E402
from oatk_b200 import synth                          # noqa: E402
from oatk_b200.host import build_host                # noqa: E402
from pyoracle import Ref                             # noqa: E402
from test_alignment_cpu import _mutate, _sample      # noqa: E402


def random_genomes(rng, k):
    rnd = lambda n: bytes(b"ACGT"[i] for i in rng.integers(0, 4, n))
    base = rnd(int(rng.integers(20000, 60000)))
    parts = [base]
    for _ in range(int(rng.integers(0, 4))):             # dispersed repeats, some inverted, some shorter than k
        rep = rnd(int(rng.choice([k - 8, k + 50, 3 * k, 3000, 6000])))
        for _ in range(int(rng.integers(2, 4))):
            parts.append(rep if rng.random() < .6 else synth.revcomp(rep))
            parts.append(rnd(int(rng.integers(3000, 12000))))
    if rng.random() < .4:                                 # a tandem array
        parts.append(rnd(int(rng.integers(150, 900))) * int(rng.integers(3, 12)))
        parts.append(rnd(5000))
    g = b"".join(parts)
    out = [g] * int(rng.integers(3, 12))
    if rng.random() < .6:
        out += [_mutate(rng, g, float(rng.choice([0.0003, 0.001, 0.003])))] * int(rng.integers(1, 6))
    for _ in range(int(rng.integers(0, 4))):              # rare side molecules: tips, chimeras
        a, b = sorted(int(x) for x in rng.integers(0, len(g) - 4000, 2))
        out.append(g[a:a + 4000] + (rnd(int(rng.integers(500, 6000))) if rng.random() < .5 else g[b:b + 4000]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", default="0:20")
    ap.add_argument("--threads", type=int, default=3)
    ap.add_argument("--ref-ec", action="store_true", help="read error correction of our arm by the reference (default: by this layer, without the up-front consensus)")
    ap.add_argument("--hifi", action="store_true", help="k = 1001, s = 31 and 15 kb reads (the reference's defaults) on larger genomes")
    args = ap.parse_args()
    lo, hi = (int(x) for x in args.seeds.split(":"))
    R = Ref().L
    H = C.CDLL(build_host.build())
    R.syncasm.restype = C.c_int
    R.syncasm.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                          C.c_double, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_int]
    R.ref_write_gfa2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    R.ref_read_ec.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int]
    R.ref_ra_new.restype = C.c_void_p
    R.ref_make_graph.restype = C.c_void_p
    R.ref_make_graph.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double]
    R.ref_collect.restype = C.c_void_p
    R.ref_extract_file.restype = C.c_void_p
    H.oatk_syncasm_graph_stage.restype = C.c_int
    H.oatk_syncasm_graph_stage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_char_p, C.c_int]
    H.scg_ra_v_destroy.argtypes = [C.c_void_p]
    H.read_error_correction.restype = None
    H.read_error_correction.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_void_p, C.c_int]
    bad = []
    tmp = tempfile.mkdtemp()
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
        rc = R.syncasm(files, 1, 0, k, s, bubble, tip, mkc, af, weak, ec, unzip, 2, p_ref.encode(), None, 0)
        desc = dict(seed=seed, k=k, s=s, n=n, L=(L1, L2), err=err, mkc=mkc, af=af, ec=ec, unzip=unzip, bubble=bubble, tip=tip, weak=weak)
        if rc != 0:
            print("skip (reference: empty graph)", desc, flush=True)
            continue
        rdb = R.ref_extract_file(fa.encode(), k, s, 2, 0)
        scm = R.ref_collect(rdb)
        if ec:
            g = R.ref_make_graph(rdb, scm, 0, 0.0)
            if args.ref_ec:
                R.ref_write_gfa2(rdb, g, 1, 1, b"/dev/null")
                R.ref_read_ec(rdb, g, 0.02, mkc, mkc * 10, mkc, af, 2)
            else:
                H.read_error_correction(rdb, g, 0.02, mkc, mkc * 10, mkc, af, args.threads, None, 0)
            R.ref_scg_free(g)
        g = R.ref_make_graph(rdb, scm, mkc, af)
        R.ref_unitig(g)
        ra = R.ref_ra_new()
        assert H.oatk_syncasm_graph_stage(rdb, g, ra, bubble, tip, weak, unzip, args.threads, p_ours.encode(), 0) == 0
        res = {}
        for suffix in (".utg.gfa", ".utg.final.gfa"):
            a, b = open(p_ours + suffix, "rb").read(), open(p_ref + suffix, "rb").read()
            res[suffix] = (a == b, b.count(b"\nS\t"), b.count(b"\nL\t"))
        ok = all(v[0] for v in res.values())
        print("ok " if ok else "BAD", desc, res, flush=True)
        if not ok:
            bad.append(desc)
        H.scg_ra_v_destroy(ra)
        R.ref_scg_free(g)
        R.ref_scm_db_free(scm)
        R.ref_sr_db_free(rdb)
    print("divergences:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
