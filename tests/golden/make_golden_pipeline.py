"""Generates tests/golden/pipeline.json from the UNMODIFIED reference (oracle/_ref/libref.so): for a few seeded
read sets, what run_syncasm.c:79-166 produces up to <out>.utg.gfa -- with read error correction (the default) and
without (--no-read-ec): md5 of the GFA text, its S/L line counts, and the [M::sr_db_stat] lines of both passes.
Run in the build container:  python tests/golden/make_golden_pipeline.py
tests/test_gpu_golden_pipeline.py replays the same inputs through the host layer on the GPU and compares."""
import ctypes as C
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
from oatk_b200 import synth            # noqa: E402
from pyoracle import Ref, pack_reads    # noqa: E402

CASES = {
    # name: (seed, genome, reads, read length, error rate, k, s, min_k_cov)
    "hifi_k1001": (13, 60000, 300, 15000, 0.002, 1001, 31, 10),
    "hifi_k301": (13, 40000, 240, 9000, 0.004, 301, 15, 8),
    "hifi_k501_noisy": (21, 50000, 260, 12000, 0.006, 501, 31, 8),
}


def reads_of(case):
    seed, G, n, L, err, k, s, mkc = CASES[case]
    return synth.hifi_reads(seed, G, n, L, err) + synth.adversarial_reads(3, k, s)


def _lines(fn, db):
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    path = tempfile.mktemp()
    fo = libc.fopen(path.encode(), b"w")
    fn(db, fo, 0)
    libc.fclose(fo)
    txt = open(path).read()
    os.unlink(path)
    return txt.splitlines()


def main():
    R = Ref()
    L = R.L
    L.sr_db_stat.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.ref_write_gfa.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
    L.ref_write_gfa2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    L.ref_read_ec.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int]
    out = {}
    for name, (seed, G, n, Lr, err, k, s, mkc) in CASES.items():
        bases, off = pack_reads(reads_of(name))
        rec = {}
        for ec in (False, True):
            db, _ = R.extract(bases, off, k, s)
            stat1 = _lines(L.sr_db_stat, db)
            scm = R.collect(db)
            stat2 = None
            if ec:
                g = R.graph(db, scm, 0, 0.0)
                assert L.ref_write_gfa2(db, g, 1, 1, b"/dev/null") == 0
                L.ref_read_ec(db, g, 0.02, mkc, mkc * 10, mkc, 0.35, 2)
                stat2 = _lines(L.sr_db_stat, db)
                R.free(g=g)
            g = R.graph(db, scm, mkc, 0.35)
            R.unitig(g)
            path = tempfile.mktemp()
            assert L.ref_write_gfa(db, g, path.encode()) == 0
            txt = open(path, "rb").read()
            os.unlink(path)
            rec["ec" if ec else "no_ec"] = {"gfa_md5": hashlib.md5(txt).hexdigest(), "S": txt.count(b"\nS\t"), "L": txt.count(b"\nL\t"),
                                            "bytes": len(txt), "stat1": stat1, "stat2": stat2}
            R.free(g=g)
            R.free(db, scm)
        out[name] = rec
        print(name, {k_: (v["gfa_md5"], v["S"], v["L"]) for k_, v in rec.items()})
    json.dump(out, open(os.path.join(HERE, "pipeline.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
