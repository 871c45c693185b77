"""Generates tests/golden/syncasm.json from the UNMODIFIED reference (oracle/_ref/libref.so, which holds
run_syncasm.c's syncasm()): for a few seeded read sets with repeats, a second haplotype and rare molecules, the md5 /
size / S,L line counts of <out>.utg.gfa and <out>.utg.final.gfa the reference command writes.
Run in the build container:  python tests/golden/make_golden_syncasm.py
tests/test_gpu_syncasm.py runs the host layer's syncasm() on the same FASTA on the GPU and compares."""
import ctypes as C
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np                                  # noqa: E402

CASES = {
    # name: (kind, k, s, min_k_cov (0 = from the k-mer spectrum), arc fraction, reads, (long, short) read length, error rate, seed,
    #        read EC, unzip rounds, max bubble, max tip, weak cross)
    "repeats_default": ("repeats", 201, 15, 3, 0.35, 900, (11000, 2500), 0.0002, 7, 1, 3, 100000, 10000, 0.3),
    "repeats_k101_no_ec": ("repeats", 101, 11, 2, 0.05, 1200, (10000, 1500), 0.0003, 8, 0, 3, 100000, 10000, 0.3),
    "diploid_unzip": ("diploid", 201, 15, 3, 0.2, 500, (9000, 3000), 0.0003, 10, 1, 3, 100000, 10000, 0.3),
    "diploid_no_unzip": ("diploid", 201, 15, 3, 0.2, 500, (9000, 3000), 0.0003, 10, 1, 0, 100000, 10000, 0.3),
    "organelle_auto_cov": ("organelle", 201, 15, 0, 0.35, 3600, (8000, 5000), 0.0003, 11, 1, 3, 100000, 10000, 0.3),
    "minor": ("minor", 201, 15, 2, 0.05, 900, (6000, 6000), 0.0003, 12, 1, 1, 100000, 10000, 0.3),
    "chimera": ("chimera", 201, 15, 2, 0.05, 2500, (5000, 5000), 0.0002, 17, 0, 1, 1000, 3000, 0.3),
    "hifi_k1001": ("plain", 1001, 31, 5, 0.35, 240, (15000, 15000), 0.0005, 19, 1, 3, 100000, 10000, 0.3),
}


def write_fasta(name, path):
    from test_alignment_cpu import _sample, _genome
    from test_cleaning_cpu import _genome as _genome2
    kind, k, s, mkc, af, n, L, err, seed = CASES[name][:9]
    rng = np.random.default_rng(seed)
    if kind == "organelle":                         # a thin nuclear background and a deep circular organelle with an inverted repeat
        rnd = lambda m: bytes(b"ACGT"[i] for i in rng.integers(0, 4, m))
        ir = rnd(4000)
        from oatk_b200 import synth
        genomes = [rnd(150000)] + [rnd(12000) + ir + rnd(5000) + synth.revcomp(ir)] * 2
    else:
        genomes = _genome2(kind, rng) if kind in ("minor", "branches", "chimera") else _genome(kind, rng)
    reads = _sample(rng, genomes, n // 2, L[0], err) + _sample(rng, genomes, n - n // 2, L[1], err)
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            if i % 3 == 2:                           # every third record as FASTQ-less multi-line FASTA
                f.write(b">r%d some comment\n" % i)
                for o in range(0, len(r), 70):
                    f.write(r[o:o + 70] + b"\n")
            else:
                f.write(b">r%d\n%s\n" % (i, r))


def args_of(name):
    kind, k, s, mkc, af, n, L, err, seed, ec, unzip, bubble, tip, weak = CASES[name]
    return dict(k=k, s=s, mkc=mkc, af=af, ec=ec, unzip=unzip, bubble=bubble, tip=tip, weak=weak)


def bind(L):
    L.syncasm.restype = C.c_int
    L.syncasm.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                          C.c_double, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_int]


def run(L, fa, a, prefix, threads):
    files = (C.c_char_p * 1)(fa.encode())
    return L.syncasm(files, 1, 0, a["k"], a["s"], a["bubble"], a["tip"], a["mkc"], a["af"], a["weak"], a["ec"], a["unzip"], threads, prefix.encode(), None, 0)


def summary(path):
    txt = open(path, "rb").read()
    return {"md5": hashlib.md5(txt).hexdigest(), "bytes": len(txt), "S": txt.count(b"\nS\t"), "L": txt.count(b"\nL\t")}


def main():
    from pyoracle import Ref
    L = Ref().L
    bind(L)
    out = {}
    tmp = tempfile.mkdtemp()
    for name in CASES:
        fa, prefix = os.path.join(tmp, name + ".fa"), os.path.join(tmp, name)
        write_fasta(name, fa)
        assert run(L, fa, args_of(name), prefix, 2) == 0
        out[name] = {suffix: summary(prefix + suffix) for suffix in (".utg.gfa", ".utg.final.gfa")}
        out[name]["fasta_md5"] = hashlib.md5(open(fa, "rb").read()).hexdigest()
        print(name, out[name])
        for suffix in (".utg.gfa", ".utg.final.gfa"):
            os.unlink(prefix + suffix)
        os.unlink(fa)
    os.rmdir(tmp)
    json.dump(out, open(os.path.join(HERE, "syncasm.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
