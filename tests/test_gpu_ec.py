"""f2 on the device (csrc/sg_ec.cu, sg_ec_correct) against its host form (host/syncerr_gpu.c, worker threads): the same
reads, database and filtered graph go through read_error_correction twice -- once with OATK_EC_HOST=1, once on the
device -- and every read's rewritten list (k_mer, m_pos, s_mer), the rebuilt database and the summary lines must agree.
The host form itself is pinned to the unmodified reference in tests/test_syncerr_cpu.py and tests/test_host_layer.py."""
import ctypes as C
import os
import numpy as np
import pytest
from oatk_b200 import synth
from pyoracle import pack_reads

pytestmark = pytest.mark.gpu


class Sr(C.Structure):
    _fields_ = [("sid", C.c_uint64), ("sname", C.c_void_p), ("hoco_l", C.c_uint32), ("hoco_s", C.c_void_p), ("ho_rl", C.c_void_p),
                ("ho_l_rl", C.c_void_p), ("n_nucl", C.c_void_p), ("n", C.c_uint32), ("m_pos", C.POINTER(C.c_uint32)),
                ("s_mer", C.POINTER(C.c_uint64)), ("k_mer", C.POINTER(C.c_uint64))]


class SrDb(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.POINTER(Sr)), ("k", C.c_int), ("s", C.c_int), ("stats", C.c_void_p)]


class Scm(C.Structure):
    _fields_ = [("h", C.c_uint64), ("s", C.c_uint64), ("covdel", C.c_uint32), ("m_pos", C.POINTER(C.c_uint64))]


class ScmDb(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.POINTER(Scm)), ("c", C.c_void_p), ("h", C.c_void_p)]


@pytest.fixture(scope="module")
def host():
    from oatk_b200.host import build_host
    L = C.CDLL(build_host.build())
    L.sr_read_mem.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.sr_db_init.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.collect_syncmer_from_reads.restype = C.c_void_p
    L.collect_syncmer_from_reads.argtypes = [C.c_void_p]
    L.sr_db_clean.argtypes = [C.c_void_p]
    L.syncmer_db_destroy.argtypes = [C.c_void_p]
    L.make_syncmer_graph.restype = C.c_void_p
    L.make_syncmer_graph.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double]
    L.scg_destroy.argtypes = [C.c_void_p]
    L.scg_consensus.restype = None
    L.scg_consensus.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.read_error_correction.restype = None
    L.read_error_correction.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_void_p, C.c_int]
    L.oatk_ec_last_run.argtypes = [C.POINTER(C.c_uint64)]
    L.read_error_correction_device.restype = C.c_int
    L.read_error_correction_device.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_int]
    return L


def run_ec(host, reads, k, s, mkc, on_host, up_front_consensus, graph_free=False):
    bases, off = pack_reads(reads)
    db = SrDb()
    host.sr_db_init(C.byref(db), k, s)
    assert host.sr_read_mem(C.byref(db), bases.ctypes.data, off.ctypes.data, None, len(reads)) == 0
    scm = host.collect_syncmer_from_reads(C.byref(db))
    if graph_free:
        # what syncasm() does: no all-syncmer graph on the host at all
        os.environ.pop("OATK_EC_HOST", None)
        assert host.read_error_correction_device(C.byref(db), scm, 0.02, mkc, mkc * 10, mkc, 0.35, 4, 1) == 0
        over = C.c_uint64(0)
        where = host.oatk_ec_last_run(C.byref(over))
        lists = []
        for r in range(db.n):
            a = db.a[r]
            n = a.n
            lists.append((np.ctypeslib.as_array(a.k_mer, (n,)).copy() if n else np.zeros(0, np.uint64),
                          np.ctypeslib.as_array(a.m_pos, (n,)).copy() if n else np.zeros(0, np.uint32),
                          np.ctypeslib.as_array(a.s_mer, (n,)).copy() if n else np.zeros(0, np.uint64)))
        S = C.cast(scm, C.POINTER(ScmDb)).contents
        covdel = np.array([S.a[i].covdel for i in range(S.n)], dtype=np.uint32)
        host.syncmer_db_destroy(scm)
        host.sr_db_clean(C.byref(db))
        return where, int(over.value), lists, covdel
    g = host.make_syncmer_graph(C.byref(db), scm, 0, 0.0)
    if up_front_consensus:
        host.scg_consensus(C.byref(db), g, 1, 1, None)      # what run_syncasm.c:118 does; without it the texts follow the filter
    if on_host:
        os.environ["OATK_EC_HOST"] = "1"
    else:
        os.environ.pop("OATK_EC_HOST", None)
    try:
        host.read_error_correction(C.byref(db), g, 0.02, mkc, mkc * 10, mkc, 0.35, 4, None, 1)
    finally:
        os.environ.pop("OATK_EC_HOST", None)
    over = C.c_uint64(0)
    where = host.oatk_ec_last_run(C.byref(over))
    lists = []
    for r in range(db.n):
        a = db.a[r]
        n = a.n
        lists.append((np.ctypeslib.as_array(a.k_mer, (n,)).copy() if n else np.zeros(0, np.uint64),
                      np.ctypeslib.as_array(a.m_pos, (n,)).copy() if n else np.zeros(0, np.uint32),
                      np.ctypeslib.as_array(a.s_mer, (n,)).copy() if n else np.zeros(0, np.uint64)))
    S = C.cast(scm, C.POINTER(ScmDb)).contents
    covdel = np.array([S.a[i].covdel for i in range(S.n)], dtype=np.uint32)
    host.scg_destroy(g)
    host.syncmer_db_destroy(scm)
    host.sr_db_clean(C.byref(db))
    return where, int(over.value), lists, covdel


def same(a, b):
    assert len(a) == len(b)
    for r, (x, y) in enumerate(zip(a, b)):
        for f, u, v in zip(("k_mer", "m_pos", "s_mer"), x, y):
            assert np.array_equal(u, v), "read %d: %s differs (%d vs %d entries)" % (r, f, len(u), len(v))


@pytest.mark.parametrize("k,s,G,n,L,err,mkc,front", [(1001, 31, 60000, 300, 15000, 0.002, 10, 1), (301, 15, 40000, 240, 9000, 0.004, 8, 0),
                                                      (501, 31, 30000, 400, 12000, 0.003, 12, 1), (127, 31, 20000, 500, 4000, 0.006, 10, 0)])
def test_device_search_equals_host_search(host, k, s, G, n, L, err, mkc, front):
    reads = synth.hifi_reads(13, G, n, L, err) + synth.adversarial_reads(3, k, s)
    wh, _, lh, ch = run_ec(host, reads, k, s, mkc, True, front)
    wd, over, ld, cd = run_ec(host, reads, k, s, mkc, False, front)
    assert wh == 0 and wd == 1, "the second run was expected on the device"
    same(lh, ld)
    assert np.array_equal(ch, cd)
    corrected = sum(int((x[0] & 1).sum()) for x in ld)
    assert corrected > 0, "no block was corrected: the test does not exercise the search"
    # the graph-free form of syncasm(): filter on the device, overlaps of the surviving arcs only
    wf, _, lf, cf = run_ec(host, reads, k, s, mkc, False, front, graph_free=True)
    assert wf == 2
    same(lh, lf)
    assert np.array_equal(ch, cf)


def test_repeats_and_haplotypes(host):
    """tandem repeats shorter than the window and two haplotypes: blocks with several equally good paths (AMBISEQ /
    AMBISNQ), cycles in the graph (the leaf cap and deep paths) and tail blocks"""
    rng = np.random.default_rng(5)
    unit = synth._nohp(rng, 700)
    hapA = synth._rand(rng, 20000) + unit * 6 + synth._rand(rng, 20000)
    hapB = bytearray(hapA)
    for p in rng.integers(0, len(hapB), 25):
        hapB[p] = ord("ACGT"[(b"ACGT".index(hapB[p]) + 1) % 4])
    reads = []
    for hap in (hapA, bytes(hapB)):
        for _ in range(160):
            st = int(rng.integers(0, len(hap) - 9000))
            r = bytearray(hap[st:st + 9000])
            for p in rng.integers(0, len(r), rng.binomial(len(r), 0.004)):
                r[p] = ord("ACGT"[(b"ACGT".index(r[p]) + 1 + int(rng.integers(0, 3))) % 4])
            r = bytes(r)
            reads.append(synth.revcomp(r) if rng.integers(0, 2) else r)
    for k, s, mkc in ((301, 15, 6), (501, 31, 6)):
        wh, _, lh, ch = run_ec(host, reads, k, s, mkc, True, 0)
        wd, over, ld, cd = run_ec(host, reads, k, s, mkc, False, 0)
        assert wh == 0 and wd == 1
        same(lh, ld)
        assert np.array_equal(ch, cd)
        wf, _, lf, cf = run_ec(host, reads, k, s, mkc, False, 0, graph_free=True)
        assert wf == 2
        same(lh, lf)
        assert np.array_equal(ch, cf)


def test_worst_case_arena_pass(host):
    """a first-pass arena too small for almost any search: the reads go through the overflow list and the worst-case
    arena, with the same result"""
    from oatk_b200 import lib
    reads = synth.hifi_reads(13, 40000, 240, 9000, 0.004) + synth.adversarial_reads(3, 301, 15)
    wh, _, lh, ch = run_ec(host, reads, 301, 15, 8, True, 0)
    L = lib.library()
    L.sg_debug_set_ec_arena.argtypes = [C.c_uint32, C.c_uint32]
    assert L.sg_debug_set_ec_arena(2, 24) == 0
    try:
        wd, over, ld, cd = run_ec(host, reads, 301, 15, 8, False, 0)
    finally:
        L.sg_debug_set_ec_arena(256, 16384)
    assert wd == 1 and over > 0, "no read took the worst-case pass (%d)" % over
    same(lh, ld)
    assert np.array_equal(ch, cd)


def test_votes_on_the_device_equal_the_host_votes(host):
    """sg_arc_votes (the distance votes of calc_syncmer_overlap on the device) against the host's vote table: the graph-free
    error-correction step run twice, once with the votes counted on the host (OATK_VOTES_HOST=1), on reads with tandem
    repeats of varying copy number -- where the same pair of syncmers sits at several distances and votes tie"""
    rng = np.random.default_rng(9)
    unit = synth._nohp(rng, 90)
    reads = []
    for copies in (3, 4, 5):
        g = synth._rand(rng, 9000) + unit * copies + synth._rand(rng, 9000)
        for _ in range(50):
            st = int(rng.integers(0, len(g) - 8000))
            r = bytearray(g[st:st + 8000])
            for p in rng.integers(0, len(r), rng.binomial(len(r), 0.003)):
                r[p] = ord("ACGT"[(b"ACGT".index(r[p]) + 1 + int(rng.integers(0, 3))) % 4])
            r = bytes(r)
            reads.append(synth.revcomp(r) if rng.integers(0, 2) else r)
    reads += synth.hifi_reads(21, 30000, 200, 8000, 0.003)
    for k, s, mkc in ((201, 15, 6), (301, 31, 8)):
        os.environ["OATK_VOTES_HOST"] = "1"
        try:
            w0, _, l0, c0 = run_ec(host, reads, k, s, mkc, False, 0, graph_free=True)
        finally:
            os.environ.pop("OATK_VOTES_HOST", None)
        w1, _, l1, c1 = run_ec(host, reads, k, s, mkc, False, 0, graph_free=True)
        wh, _, lh, ch = run_ec(host, reads, k, s, mkc, True, 0)
        assert w0 == 2 and w1 == 2 and wh == 0
        same(l0, l1)
        same(lh, l1)
        assert np.array_equal(c0, c1) and np.array_equal(ch, c1)
