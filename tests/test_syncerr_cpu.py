"""f2 on the CPU: read error correction of the host layer (oatk_b200/host/syncerr_gpu.c) against the reference's
read_error_correction on identical structures: two pipelines are built by the UNMODIFIED reference (reads ->
syncmer database -> all-syncmer graph -> hoco consensus), one is corrected by the reference, the other by our
code, and every per-read array and the rebuilt syncmer database must agree. Needs oracle/_ref/libref.so; no GPU."""
import ctypes as C
import numpy as np
import pytest
from oatk_b200 import synth
from pyoracle import pack_reads, count_ambiguous


@pytest.fixture(scope="module")
def host():
    from oatk_b200.host import build_host
    try:
        L = C.CDLL(build_host.build())
    except OSError as e:
        pytest.skip("host layer not loadable: %s" % e)
    L.read_error_correction.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_void_p, C.c_int]
    L.read_error_correction.restype = None
    return L


def _pipeline(ref, bases, off, k, s, consensus=True):
    rdb, _ = ref.extract(bases, off, k, s)
    rscm = ref.collect(rdb)
    g = ref.graph(rdb, rscm, 0, 0.0)                      # all syncmers: vertex id == syncmer id
    ref.L.ref_write_gfa2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    if consensus:
        assert ref.L.ref_write_gfa2(rdb, g, 1, 1, b"/dev/null") == 0      # hoco consensus saved in the graph (run_syncasm.c:118)
    return rdb, rscm, g


def _state(ref, rdb, rscm, n, n_n):
    f = ref._flat(rdb, n, n_n)
    S = rscm["_handle"]
    U = int(ref.L.ref_scm_n(S))
    ref.L.ref_scm_total_cov.restype = C.c_uint64
    ref.L.ref_scm_total_cov.argtypes = [C.c_void_p]
    N = int(ref.L.ref_scm_total_cov(S))
    d = dict(h=np.zeros(U, np.uint64), s=np.zeros(U, np.uint64), cov=np.zeros(U, np.uint32), occ=np.zeros(N + 1, np.uint64), dele=np.zeros(U, np.uint8))
    ref.L.ref_scm_flatten(S, d["h"].ctypes.data, d["s"].ctypes.data, d["cov"].ctypes.data, d["occ"].ctypes.data)
    ref.L.ref_scm_flags.argtypes = [C.c_void_p, C.c_void_p]
    ref.L.ref_scm_flags(S, d["dele"].ctypes.data)
    d["occ"] = d["occ"][:N]
    for key in ("n_scm", "m_pos", "s_mer", "k_mer"):
        d["read_" + key] = f[key]
    return d


CASES = [
    # k, s, genome, reads, read length, error rate, seed, min_k_cov, threads
    (301, 15, 40000, 200, 9000, 0.004, 3, 8, 1),
    (301, 15, 40000, 200, 9000, 0.004, 3, 8, 4),
    (1001, 31, 60000, 300, 15000, 0.002, 7, 10, 2),
    (101, 11, 20000, 250, 4000, 0.01, 11, 6, 3),
    (63, 9, 15000, 300, 3000, 0.02, 5, 5, 2),
]


@pytest.mark.parametrize("k,s,G,n,L,err,seed,mkc,threads", CASES)
def test_read_error_correction_matches_reference(host, ref, k, s, G, n, L, err, seed, mkc, threads):
    reads = synth.hifi_reads(seed, G, n, L, err) + synth.adversarial_reads(3, k, s)
    bases, off = pack_reads(reads)
    n_n = count_ambiguous(bases, off)
    a = _pipeline(ref, bases, off, k, s)
    b = _pipeline(ref, bases, off, k, s)
    before = _state(ref, a[0], a[1], len(reads), n_n)
    ref.L.ref_read_ec.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int]
    ref.L.ref_read_ec(a[0], a[2], 0.02, mkc, mkc * 10, mkc, 0.35, threads)
    host.read_error_correction(b[0], b[2], 0.02, mkc, mkc * 10, mkc, 0.35, threads, None, 0)
    want = _state(ref, a[0], a[1], len(reads), n_n)
    got = _state(ref, b[0], b[1], len(reads), n_n)
    assert not np.array_equal(want["read_k_mer"], before["read_k_mer"]), "the case corrects nothing"
    for key in want:
        assert np.array_equal(got[key], want[key]), key
    # the graphs were pruned the same way (suspect vertices and their arcs)
    d1, d2 = ref.graph_dump(a[2]), ref.graph_dump(b[2])
    for f in d1:
        assert np.array_equal(d1[f], d2[f]), f
    for x in (a, b):
        ref.free(g=x[2])
        ref.free(x[0], x[1])


@pytest.mark.parametrize("k,s,G,n,L,err,seed,mkc,threads", CASES)
def test_deferred_consensus(host, ref, k, s, G, n, L, err, seed, mkc, threads):
    """read_error_correction on a graph WITHOUT the up-front all-syncmer consensus: it computes the texts and overlaps of
    what its error filter leaves, and the corrected reads, the database and the pruning are the reference's all the same"""
    reads = synth.hifi_reads(seed, G, n, L, err) + synth.adversarial_reads(3, k, s)
    bases, off = pack_reads(reads)
    n_n = count_ambiguous(bases, off)
    a = _pipeline(ref, bases, off, k, s)
    b = _pipeline(ref, bases, off, k, s, consensus=False)
    ref.L.ref_read_ec.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int]
    ref.L.ref_read_ec(a[0], a[2], 0.02, mkc, mkc * 10, mkc, 0.35, threads)
    host.read_error_correction(b[0], b[2], 0.02, mkc, mkc * 10, mkc, 0.35, threads, None, 0)
    want = _state(ref, a[0], a[1], len(reads), n_n)
    got = _state(ref, b[0], b[1], len(reads), n_n)
    for key in want:
        assert np.array_equal(got[key], want[key]), key
    d1, d2 = ref.graph_dump(a[2]), ref.graph_dump(b[2])
    assert np.array_equal(d1["vtx_flags"] >> 30, d2["vtx_flags"] >> 30)                  # same vertices pruned
    assert np.array_equal(d1["arcs"][:, 4] >> 30, d2["arcs"][:, 4] >> 30)                # same arcs pruned
    live = (d1["arcs"][:, 4] >> 30 & 1) == 0
    assert np.array_equal(d1["arcs"][live], d2["arcs"][live])                            # and the same overlaps on the live ones
    for x in (a, b):
        ref.free(g=x[2])
        ref.free(x[0], x[1])


def test_randomised_inputs_with_repeats(host, ref):
    """random genome sizes, read lengths, error rates and coverage thresholds, half of the cases with a planted repeat
    (two contexts around one shared segment: alternative paths, ambiguous blocks)"""
    rng = np.random.default_rng(0)
    ref.L.ref_read_ec.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int]
    bad = []
    for it in range(10):
        k, s = [(301, 15), (101, 11), (63, 9), (501, 31), (41, 7)][it % 5]
        G, L, n = int(rng.integers(8000, 50000)), int(rng.integers(2000, 12000)), int(rng.integers(60, 300))
        err, mkc = float(rng.choice([0.002, 0.005, 0.01, 0.02])), int(rng.integers(3, 12))
        reads = synth.hifi_reads(100 + it, G, n, L, err)
        if it % 2:
            rep = reads[0][:min(len(reads[0]), 3000)]
            reads += [rep + synth.hifi_reads(500 + it, 5000, 1, 3000, 0)[0] for _ in range(mkc + 2)]
            reads += [synth.hifi_reads(600 + it, 5000, 1, 3000, 0)[0] + rep for _ in range(mkc + 2)]
        bases, off = pack_reads(reads)
        n_n = count_ambiguous(bases, off)
        a, b = _pipeline(ref, bases, off, k, s), _pipeline(ref, bases, off, k, s)
        ref.L.ref_read_ec(a[0], a[2], 0.02, mkc, mkc * 10, mkc, 0.35, 2)
        host.read_error_correction(b[0], b[2], 0.02, mkc, mkc * 10, mkc, 0.35, 3, None, 0)
        want, got = _state(ref, a[0], a[1], len(reads), n_n), _state(ref, b[0], b[1], len(reads), n_n)
        diff = [x for x in want if not np.array_equal(got[x], want[x])]
        if diff:
            bad.append((it, k, s, G, L, n, err, mkc, diff))
        for x in (a, b):
            ref.free(g=x[2])
            ref.free(x[0], x[1])
    assert not bad, bad


def test_wavefront_edit_distance(host, ref):
    """the resumable wavefront edit distance under the graph search, alone: the reference's built-in strings
    (levdist.c:445-446) and random pairs, in one go and with the query growing piece by piece, with and without band"""
    host.oatk_wave_align.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    ref.L.ref_wave_align.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p]

    def both(t, q, bw, grow):
        a, b = np.zeros(3, np.int32), np.zeros(3, np.int32)
        host.oatk_wave_align(t, len(t), q, len(q), bw, grow, a.ctypes.data)
        ref.L.ref_wave_align(t, len(t), q, len(q), bw, grow, b.ctypes.data)
        return tuple(a), tuple(b)

    t = b"AATGCTCTCATGACATATGAGATAGATACATAGAGACAGATATAGATACACACAGAGATATATGACGTCTGTATGCTCTCTCTCATAGATATACTCTGTAGACTGTCATATACATGCAGAAAAA"
    q = b"CGCTCTCATGACANATGAGATAGATACATAGAGNCAGATATAGATACACACAGTTT"
    got, want = both(t, q, -1, 0)
    assert got == want and got[0] == 8                     # the known answer of the reference's own test main: ED = 8
    rng = np.random.default_rng(4)
    for it in range(400):
        L = int(rng.integers(10, 400))
        base = rng.integers(0, 4, L)
        tt = bytes(b"ACGT"[i] for i in base)
        qq = bytearray(tt)
        for _ in range(int(rng.integers(0, 12))):             # a few edits
            p = int(rng.integers(0, max(len(qq), 1)))
            kind = int(rng.integers(0, 3))
            if kind == 0 and qq:
                qq[p] = b"ACGT"[int(rng.integers(0, 4))]
            elif kind == 1 and qq:
                del qq[p]
            else:
                qq.insert(p, b"ACGT"[int(rng.integers(0, 4))])
        qq = bytes(qq) + bytes(b"ACGT"[i] for i in rng.integers(0, 4, int(rng.integers(0, 60))))
        if not qq:
            continue
        for bw in (-1, 6, int(np.ceil(0.02 * L)) + 1):
            for grow in (0, 37, 1001):
                got, want = both(tt, qq, bw, grow)
                assert got == want, (it, L, len(qq), bw, grow, got, want)
