"""f1 on the CPU: scg_consensus of the host layer (oatk_b200/host/consensus_gpu.c) run on graphs that the
UNMODIFIED reference built (its structs are byte-compatible with ours), against the reference's own
scg_consensus on an identical second graph: GFA text, unitig lengths / coverages and arc overlaps.
Needs oracle/_ref/libref.so; no GPU (the consensus is host code over host structs)."""
import ctypes as C
import os
import tempfile
import numpy as np
import pytest
from oatk_b200 import synth
from pyoracle import pack_reads


@pytest.fixture(scope="module")
def host():
    from oatk_b200.host import build_host
    try:
        L = C.CDLL(build_host.build())
    except OSError as e:                      # libsyncgpu.so / libcudart not loadable here
        pytest.skip("host layer not loadable: %s" % e)
    L.scg_consensus.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.scg_consensus.restype = None
    return L


def _gfa(fn):
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    path = tempfile.mktemp()
    fo = libc.fopen(path.encode(), b"w")
    fn(fo)
    libc.fclose(fo)
    txt = open(path, "rb").read()
    os.unlink(path)
    return txt


def _first_diff(a, b):
    la, lb = a.split(b"\n"), b.split(b"\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            j = next((t for t in range(min(len(x), len(y))) if x[t] != y[t]), min(len(x), len(y)))
            return "line %d col %d: %r vs %r" % (i, j, x[max(0, j - 30):j + 30], y[max(0, j - 30):j + 30])
    return "line counts %d vs %d" % (len(la), len(lb))


CASES = [
    # k, s, min_k_cov, a, genome, reads, read length, error rate, seed
    (1001, 31, 10, 0.35, 60000, 300, 15000, 0.002, 77),
    (501, 31, 30, 0.35, 40000, 400, 12000, 0.001, 5),
    (301, 15, 2, 0.0, 30000, 60, 9000, 0.003, 9),       # low coverage: ties in the offset vote
    (101, 11, 3, 0.35, 20000, 120, 5000, 0.01, 21),     # noisy: many short unitigs and overlapping neighbours
]


def _rc(b):
    return bytes({65: 84, 67: 71, 71: 67, 84: 65}[x] for x in reversed(b))


def test_offset_vote_ties_and_table_growth(host, ref):
    """Neighbouring syncmers normally overlap, so every read votes for the same offset. Microsatellites of varying
    length are the exception: the same pair of k-mers sits at several distances, the vote splits (ties are decided
    by khashl slot order) and the per-unitig vote table grows past four buckets while it is reused."""
    bad = []
    host.oatk_consensus_debug_counts.argtypes = [C.c_void_p]
    c0 = np.zeros(2, np.uint64)
    host.oatk_consensus_debug_counts(c0.ctypes.data)
    rng = np.random.default_rng(5)
    rnd = lambda n: bytes(b"ACGT"[i] for i in rng.integers(0, 4, n))
    for trial, unit in enumerate([b"ACG", b"ACAG", b"AC", b"ACGTG", b"AGC", b"ATCG", b"ACTG", b"AG"]):
        left, right = rnd(400), rnd(400)
        reads = []
        for n in range(8, 30):
            seq = left + unit * n + right
            for _ in range(int(rng.integers(1, 4))):
                reads.append(seq if rng.integers(0, 2) else _rc(seq))
        bases, off = pack_reads(reads)
        rdb, _ = ref.extract(bases, off, 31, 7)
        rscm = ref.collect(rdb)
        g1, g2 = ref.graph(rdb, rscm, 2, 0.0), ref.graph(rdb, rscm, 2, 0.0)
        assert g1 and g2
        ref.unitig(g1)
        ref.unitig(g2)
        ref.L.ref_write_gfa.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
        path = tempfile.mktemp()
        assert ref.L.ref_write_gfa(rdb, g2, path.encode()) == 0
        want = open(path, "rb").read()
        os.unlink(path)
        got = _gfa(lambda fo: host.scg_consensus(rdb, g1, 0, 0, fo))
        if got != want:
            bad.append((trial, _first_diff(got, want)))
        ref.free(g=g1)
        ref.free(g=g2)
        ref.free(rdb, rscm)
    assert not bad, bad
    c1 = np.zeros(2, np.uint64)
    host.oatk_consensus_debug_counts(c1.ctypes.data)
    assert c1[0] > c0[0], "no tied votes in the stress set"
    assert c1[1] > c0[1], "no vote table ever grew in the stress set"


@pytest.mark.parametrize("k,s,mkc,a,G,n,L,err,seed", CASES)
@pytest.mark.parametrize("unitig", [False, True])
def test_gfa_matches_reference(host, ref, k, s, mkc, a, G, n, L, err, seed, unitig):
    reads = synth.hifi_reads(seed, G, n, L, err) + synth.adversarial_reads(3, k, s)
    bases, off = pack_reads(reads)
    rdb, _ = ref.extract(bases, off, k, s)
    rscm = ref.collect(rdb)
    g1, g2 = ref.graph(rdb, rscm, mkc, a), ref.graph(rdb, rscm, mkc, a)
    assert g1 and g2
    if unitig:
        ref.unitig(g1)
        ref.unitig(g2)
    ref.L.ref_write_gfa.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
    path = tempfile.mktemp()
    assert ref.L.ref_write_gfa(rdb, g2, path.encode()) == 0
    want = open(path, "rb").read()
    os.unlink(path)
    got = _gfa(lambda fo: host.scg_consensus(rdb, g1, 0, 0, fo))
    assert got.count(b"\nS\t") > 0 and got.count(b"\nL\t") > 0
    assert got == want, _first_diff(got, want)
    d1, d2 = ref.graph_dump(g1), ref.graph_dump(g2)           # ls, cov written back into the graph
    for f in d1:
        assert np.array_equal(d1[f], d2[f]), f
    # homopolymer-compressed text with saved sequences (what the error-correction graph uses, run_syncasm.c:118)
    ref.L.ref_write_gfa2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    assert ref.L.ref_write_gfa2(rdb, g2, 1, 1, path.encode()) == 0
    want = open(path, "rb").read()
    os.unlink(path)
    hoco = _gfa(lambda fo: host.scg_consensus(rdb, g1, 1, 1, fo))
    assert hoco == want, _first_diff(hoco, want)
    ref.free(g=g1)
    ref.free(g=g2)
    ref.free(rdb, rscm)


def test_long_homopolymer_runs(host, ref):
    """runs of 255 bases and more live in a side list per read (ho_l_rl); their lengths vary from read to read, so the
    consensus has to average the real lengths, not the 255 marks"""
    rng = np.random.default_rng(11)
    rnd = lambda n: bytes(b"ACGT"[i] for i in rng.integers(0, 4, n))
    parts = [rnd(3000), rnd(2500), rnd(3500), rnd(2000), rnd(3000)]
    runs = [(b"A", 300), (b"C", 255), (b"T", 700), (b"G", 254)]
    reads = []
    for _ in range(60):
        seq = parts[0]
        for (base, n), nxt in zip(runs, parts[1:]):
            seq += base * (n + int(rng.integers(-3, 4))) + nxt
        a = int(rng.integers(0, 2000))
        b = len(seq) - int(rng.integers(0, 2000))
        r = seq[a:b]
        reads.append(r if rng.integers(0, 2) else _rc(r))
    bases, off = pack_reads(reads)
    for k, s in ((301, 15), (101, 11)):
        rdb, _ = ref.extract(bases, off, k, s)
        rscm = ref.collect(rdb)
        g1, g2 = ref.graph(rdb, rscm, 3, 0.2), ref.graph(rdb, rscm, 3, 0.2)
        assert g1 and g2
        ref.unitig(g1)
        ref.unitig(g2)
        ref.L.ref_write_gfa.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
        path = tempfile.mktemp()
        assert ref.L.ref_write_gfa(rdb, g2, path.encode()) == 0
        want = open(path, "rb").read()
        os.unlink(path)
        got = _gfa(lambda fo: host.scg_consensus(rdb, g1, 0, 0, fo))
        assert got == want, _first_diff(got, want)
        assert b"A" * 290 in want or b"T" * 290 in want, "no long run made it into a unitig"
        ref.free(g=g1)
        ref.free(g=g2)
        ref.free(rdb, rscm)
