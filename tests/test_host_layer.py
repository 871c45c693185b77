"""The C host layer (oatk_b200/host/syncmer_gpu.c) as a drop-in for the reference's syncmer.h functions:
its structs are handed to the UNMODIFIED reference's own code (flatten helpers, make_syncmer_graph) to
prove byte compatibility, and its printed statistics are compared line by line."""
import ctypes as C
import os
import re
import tempfile
import numpy as np
import pytest
from oatk_b200 import synth
from pyoracle import pack_reads, count_ambiguous
import parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class SrDb(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.c_void_p), ("k", C.c_int), ("s", C.c_int), ("stats", C.c_void_p)]


@pytest.fixture(scope="module")
def host():
    from oatk_b200.host import build_host
    L = C.CDLL(build_host.build())
    L.sr_read_mem.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.sr_db_init.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.sr_db_stat.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.sr_db_validate.argtypes = [C.c_void_p]
    L.collect_syncmer_from_reads.restype = C.c_void_p
    L.collect_syncmer_from_reads.argtypes = [C.c_void_p]
    L.sr_db_clean.argtypes = [C.c_void_p]
    L.syncmer_db_destroy.argtypes = [C.c_void_p]
    L.syncmer_graph_arcs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    return L


def stat_lines(fn, db_ptr):
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    path = tempfile.mktemp()
    fo = libc.fopen(path.encode(), b"w")
    fn(db_ptr, fo, 0)
    libc.fclose(fo)
    txt = open(path).read()
    os.unlink(path)
    return txt.splitlines()


def _libc():
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    return libc


def gfa_text(host, db_ptr, g):
    host.scg_consensus.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    host.scg_consensus.restype = None
    libc = _libc()
    path = tempfile.mktemp()
    fo = libc.fopen(path.encode(), b"w")
    host.scg_consensus(db_ptr, g, 0, 0, fo)
    libc.fclose(fo)
    txt = open(path, "rb").read()
    os.unlink(path)
    return txt


def ref_gfa_text(ref, rdb, g):
    ref.L.ref_write_gfa.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
    path = tempfile.mktemp()
    assert ref.L.ref_write_gfa(rdb, g, path.encode()) == 0
    txt = open(path, "rb").read()
    os.unlink(path)
    return txt


def first_diff(a, b):
    la, lb = a.split(b"\n"), b.split(b"\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            j = next((t for t in range(min(len(x), len(y))) if x[t] != y[t]), min(len(x), len(y)))
            return "line %d col %d: %r vs %r" % (i, j, x[max(0, j - 20):j + 20], y[max(0, j - 20):j + 20])
    return "line counts %d vs %d" % (len(la), len(lb))


@pytest.mark.parametrize("k,s", [(1001, 31), (301, 15)])
def test_drop_in_structs_feed_the_reference(host, ref, k, s):
    reads = synth.hifi_reads(31, 150000, 360, 15000, 0.001) + synth.adversarial_reads(3, k, s)
    bases, off = pack_reads(reads)
    n = len(reads)
    db = SrDb()
    host.sr_db_init(C.byref(db), k, s)
    assert host.sr_read_mem(C.byref(db), bases.ctypes.data, off.ctypes.data, None, n) == 0
    assert host.sr_db_validate(C.byref(db)) == 0
    # 1. the reference's own flatten walks OUR sr_db_t
    mine = ref._flat(C.addressof(db), n, count_ambiguous(bases, off))
    rdb, theirs = ref.extract(bases, off, k, s)
    assert parity.diff(mine, theirs, parity.EXTRACT_FIELDS) == []
    # 2. sr_db_stat prints the reference's lines (the reference first, as run_syncasm.c does)
    ref.L.sr_db_stat.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    exp_lines = stat_lines(ref.L.sr_db_stat, rdb)
    got_lines = stat_lines(host.sr_db_stat, C.addressof(db))
    assert got_lines == exp_lines
    # 3. collect: the reference's flatten walks OUR syncmer_db_t, and k_mer[] was rewritten to id << 1
    scm = host.collect_syncmer_from_reads(C.byref(db))
    assert scm
    rscm = ref.collect(rdb)
    U, N = len(rscm["h"]), len(rscm["occ"])
    assert ref.L.ref_scm_n(scm) == U
    got = dict(h=np.zeros(U, np.uint64), s=np.zeros(U, np.uint64), cov=np.zeros(U, np.uint32), occ=np.zeros(N, np.uint64))
    ref.L.ref_scm_flatten(scm, got["h"].ctypes.data, got["s"].ctypes.data, got["cov"].ctypes.data, got["occ"].ctypes.data)
    ids = np.zeros(N + 1, np.uint64)
    ref.L.ref_kmer_ids(C.addressof(db), ids.ctypes.data)
    got["k_mer_id"] = ids[:N]
    assert parity.diff(got, rscm, ("h", "s", "cov", "occ", "k_mer_id")) == []
    # 4. the reference's make_syncmer_graph + unitigging run on OUR structs and give the reference's graph
    g_mine = ref.L.ref_make_graph(C.addressof(db), scm, 3, 0.35)
    g_ref = ref.graph(rdb, rscm, 3, 0.35)
    d1, d2 = ref.graph_dump(g_mine), ref.graph_dump(g_ref)
    for f in d1:
        assert np.array_equal(d1[f], d2[f]), f
    ref.unitig(g_mine)
    ref.unitig(g_ref)
    d1, d2 = ref.graph_dump(g_mine), ref.graph_dump(g_ref)
    for f in d1:
        assert np.array_equal(d1[f], d2[f]), f
    # 5. our arc list = the reference graph's arcs mapped back to syncmer ids
    # (make_syncmer_graph set the del flags in scm in place; the arc call does not depend on them)
    p, na = C.c_void_p(), C.c_uint64()
    assert host.syncmer_graph_arcs(C.byref(db), scm, 3, 0.35, C.byref(p), C.byref(na)) == 0
    arcs = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), (na.value, 4)).copy()
    g3 = ref.graph(rdb, rscm, 3, 0.35)
    gd = ref.graph_dump(g3)
    first = gd["vtx_lists"][np.concatenate([[0], np.cumsum(gd["vtx_n"])[:-1]]).astype(np.int64)]
    ra = gd["arcs"]
    exp = np.stack([first[(ra[:, 0] >> 1).astype(np.int64)] | (ra[:, 0] & 1), first[(ra[:, 1] >> 1).astype(np.int64)] | (ra[:, 1] & 1),
                    ra[:, 4] & 0x3FFFFFFF, (ra[:, 4] >> 31) & 1], axis=1).astype(np.uint64)
    exp[(exp[:, 1] ^ 1) == exp[:, 0], 3] = 0            # asmg_arc_fix_symm flips comp of self-complementary arcs
    exp = exp[np.lexsort((exp[:, 2], exp[:, 3], exp[:, 1], exp[:, 0]))]
    assert np.array_equal(arcs, exp)
    C.CDLL(None).free(p)
    ref.free(g=g_mine)
    ref.free(g=g_ref)
    ref.free(g=g3)
    host.syncmer_db_destroy(scm)
    host.sr_db_clean(C.byref(db))
    ref.free(rdb, rscm)


@pytest.mark.parametrize("k,s,mkc", [(1001, 31, 3), (301, 15, 2), (501, 31, 30)])
def test_graph_layer_matches_reference(host, ref, k, s, mkc):
    """a7-a9 of the host layer (make_syncmer_graph over the device arc tally, asmg_finalize,
    asmg_unitigging) against the reference's graph.c / syncasm.c on the same reads"""
    host.make_syncmer_graph.restype = C.c_void_p
    host.make_syncmer_graph.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double]
    host.process_mergeable_unitigs.argtypes = [C.c_void_p]
    host.scg_destroy.argtypes = [C.c_void_p]
    # 60x coverage of a small genome (so that -c 30 leaves a graph) plus the adversarial reads
    reads = synth.hifi_reads(77, 60000, 240 if mkc < 30 else 400, 15000, 0.002) + synth.adversarial_reads(3, k, s) * 2
    bases, off = pack_reads(reads)
    db = SrDb()
    host.sr_db_init(C.byref(db), k, s)
    assert host.sr_read_mem(C.byref(db), bases.ctypes.data, off.ctypes.data, None, len(reads)) == 0
    scm = host.collect_syncmer_from_reads(C.byref(db))
    rdb, _ = ref.extract(bases, off, k, s)
    rscm = ref.collect(rdb)
    for a in (0.35, 0.0):
        g_mine = host.make_syncmer_graph(C.byref(db), scm, mkc, a)
        g_ref = ref.graph(rdb, rscm, mkc, a)
        assert g_mine and g_ref
        d1, d2 = ref.graph_dump(g_mine), ref.graph_dump(g_ref)      # the reference's dump walks OUR scg_t
        for f in d1:
            assert np.array_equal(d1[f], d2[f]), ("graph", f, a)
        assert len(d1["arcs"]) > 0
        host.process_mergeable_unitigs(g_mine)
        ref.unitig(g_ref)
        d1, d2 = ref.graph_dump(g_mine), ref.graph_dump(g_ref)
        for f in d1:
            assert np.array_equal(d1[f], d2[f]), ("unitigs", f, a)
        # f1: the GFA text (unitig consensus, overlaps, coverages) of our layer against the reference's scg_consensus
        mine_gfa, ref_gfa = gfa_text(host, C.addressof(db), g_mine), ref_gfa_text(ref, rdb, g_ref)
        assert mine_gfa.count(b"\nS\t") > 0
        assert mine_gfa == ref_gfa, ("gfa", a, first_diff(mine_gfa, ref_gfa))
        # and the lengths / overlaps it leaves in the graph
        d1, d2 = ref.graph_dump(g_mine), ref.graph_dump(g_ref)
        for f in d1:
            assert np.array_equal(d1[f], d2[f]), ("after consensus", f, a)
        host.scg_destroy(g_mine)
        ref.free(g=g_ref)
    host.syncmer_db_destroy(scm)
    host.sr_db_clean(C.byref(db))
    ref.free(rdb, rscm)


def test_low_complexity_reads_need_more_room(host, ref):
    """reads that are nothing but short tandem repeats: every position ties for the window minimum, so there is close to
    one syncmer per base where the pipeline's master batch is cut for 16 times two per window. sr_read_mem repeats the run
    with more room and the result is the reference's"""
    k, s = 301, 15
    reads = [b"AC" * 4000, b"ACG" * 2500, b"TTAGGG" * 1500, b"AC" * 3000 + b"ACGT" * 800] * 12
    bases, off = pack_reads(reads)
    n = len(reads)
    db = SrDb()
    host.sr_db_init(C.byref(db), k, s)
    assert host.sr_read_mem(C.byref(db), bases.ctypes.data, off.ctypes.data, None, n) == 0
    mine = ref._flat(C.addressof(db), n, count_ambiguous(bases, off))
    rdb, theirs = ref.extract(bases, off, k, s)
    assert parity.diff(mine, theirs, parity.EXTRACT_FIELDS) == []
    total = int(off[-1])
    assert len(theirs["m_pos"]) > 16 * (total // (k - s + 1) + n) + 1024, "the set does not overflow the first capacity"
    host.sr_db_clean(C.byref(db))
    ref.free(rdb)


def test_empty_collection(host):
    bases, off = pack_reads([b"ACGTTGCA", b"NNNN", b""])
    db = SrDb()
    host.sr_db_init(C.byref(db), 1001, 31)
    assert host.sr_read_mem(C.byref(db), bases.ctypes.data, off.ctypes.data, None, 3) == 0
    assert stat_lines(host.sr_db_stat, C.addressof(db)) == ["[M::sr_db_stat] empty syncmer collection"]
    assert not host.collect_syncmer_from_reads(C.byref(db))
    host.sr_db_clean(C.byref(db))


def test_sr_read_files_matches_reference(host, ref):
    """file -> sr_db_t through the native reader (f4) and the device pipeline, against the reference's sr_read on the
    same FASTQ + gzipped FASTA pair: every per-read field and the read names"""
    host.sr_read_files.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.c_size_t]
    k, s = 301, 15
    reads = synth.hifi_reads(5, 80000, 150, 9000, 0.002) + synth.adversarial_reads(3, k, s)
    with tempfile.TemporaryDirectory() as d:
        import gzip
        p1, p2 = os.path.join(d, "a.fq"), os.path.join(d, "b.fa.gz")
        half = len(reads) // 2
        with open(p1, "wb") as f:
            for i, r in enumerate(reads[:half]):
                f.write(b"@q%d desc\n%s\n+\n%s\n" % (i, r, b"I" * len(r)))
        with gzip.open(p2, "wb") as f:
            for i, r in enumerate(reads[half:]):
                f.write(b">f%d\n" % i)
                for a in range(0, max(len(r), 1), 70):
                    f.write(r[a:a + 70] + b"\n")
        files = (C.c_char_p * 2)(p1.encode(), p2.encode())
        db = SrDb()
        host.sr_db_init(C.byref(db), k, s)
        assert host.sr_read_files(C.byref(db), files, 2, 0) == 0
        ref.L.ref_extract_file.restype = C.c_void_p
        # the reference reads one file per call of its shim: compare file by file through the in-memory path instead
        bases, off = pack_reads(reads)
        rdb, theirs = ref.extract(bases, off, k, s)
        mine = ref._flat(C.addressof(db), len(reads), count_ambiguous(bases, off))
        assert db.n == len(reads)
        assert parity.diff(mine, theirs, parity.EXTRACT_FIELDS) == []
        # names: "q<i>" then "f<i>"
        class Sr(C.Structure):
            _fields_ = [("sid", C.c_uint64), ("sname", C.c_char_p), ("hoco_l", C.c_uint32), ("hoco_s", C.c_void_p), ("ho_rl", C.c_void_p),
                        ("ho_l_rl", C.c_void_p), ("n_nucl", C.c_void_p), ("n", C.c_uint32), ("m_pos", C.c_void_p), ("s_mer", C.c_void_p), ("k_mer", C.c_void_p)]
        arr = C.cast(db.a, C.POINTER(Sr))
        assert arr[0].sname == b"q0" and arr[half - 1].sname == b"q%d" % (half - 1)
        assert arr[half].sname == b"f0" and arr[len(reads) - 1].sname == b"f%d" % (len(reads) - half - 1)
        assert all(arr[i].sid == i for i in range(0, len(reads), 17))
        # the -D cap: stop after the read that reaches it, print the reference's message
        total = int(off[-1])
        db2 = SrDb()
        host.sr_db_init(C.byref(db2), k, s)
        assert host.sr_read_files(C.byref(db2), files, 2, total // 3) == 0
        cum = np.cumsum(np.diff(off))
        assert db2.n == int(np.searchsorted(cum, total // 3, side="left")) + 1
        host.sr_db_clean(C.byref(db2))
        # the reference's own calling sequence (run_syncasm.c:79-86): sstream_open -> sr_read -> sstream_close
        class SStream(C.Structure):
            _fields_ = [("n_seq", C.c_uint64), ("files", C.POINTER(C.c_char_p)), ("n_files", C.c_int), ("n", C.c_int), ("s", C.c_void_p)]
        host.sstream_open.restype = C.POINTER(SStream)
        host.sstream_open.argtypes = [C.POINTER(C.c_char_p), C.c_int]
        host.sr_read.argtypes = [C.POINTER(SStream), C.c_void_p, C.c_size_t, C.c_int]
        host.sr_read.restype = None
        host.sstream_close.argtypes = [C.POINTER(SStream)]
        ss = host.sstream_open(files, 2)
        db3 = SrDb()
        host.sr_db_init(C.byref(db3), k, s)
        host.sr_read(ss, C.byref(db3), 0, 8)
        assert ss.contents.n_seq == len(reads) and ss.contents.n == 1
        host.sstream_close(ss)
        again = ref._flat(C.addressof(db3), len(reads), count_ambiguous(bases, off))
        assert parity.diff(again, theirs, parity.EXTRACT_FIELDS) == []
        host.sr_db_clean(C.byref(db3))
        host.sr_db_clean(C.byref(db))
        ref.free(rdb)


@pytest.mark.parametrize("k,s,G,n,L,err,mkc", [(1001, 31, 60000, 300, 15000, 0.002, 10), (301, 15, 40000, 240, 9000, 0.004, 8)])
def test_default_pipeline_with_error_correction(host, ref, k, s, G, n, L, err, mkc):
    """run_syncasm.c:79-166 with read error correction on (the default): reads -> statistics -> database -> all-syncmer
    graph -> hoco consensus -> read error correction -> statistics again -> final graph -> unitigs -> GFA, through the host
    layer (device: extraction, counting, both statistics passes, both arc tallies) and through the unmodified reference"""
    for f, a in (("make_syncmer_graph", [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double]), ("process_mergeable_unitigs", [C.c_void_p]),
                 ("scg_destroy", [C.c_void_p]), ("scg_consensus", [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
                 ("read_error_correction", [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_void_p, C.c_int])):
        getattr(host, f).argtypes = a
    host.make_syncmer_graph.restype = C.c_void_p
    host.scg_consensus.restype = None
    host.read_error_correction.restype = None
    reads = synth.hifi_reads(13, G, n, L, err) + synth.adversarial_reads(3, k, s)
    bases, off = pack_reads(reads)
    # ---- ours
    db = SrDb()
    host.sr_db_init(C.byref(db), k, s)
    assert host.sr_read_mem(C.byref(db), bases.ctypes.data, off.ctypes.data, None, len(reads)) == 0
    stat1 = stat_lines(host.sr_db_stat, C.addressof(db))
    scm = host.collect_syncmer_from_reads(C.byref(db))
    g = host.make_syncmer_graph(C.byref(db), scm, 0, 0.0)
    host.scg_consensus(C.byref(db), g, 1, 1, None)
    host.read_error_correction(C.byref(db), g, 0.02, mkc, mkc * 10, mkc, 0.35, 4, None, 0)
    stat2 = stat_lines(host.sr_db_stat, C.addressof(db))
    host.scg_destroy(g)
    g = host.make_syncmer_graph(C.byref(db), scm, mkc, 0.35)
    assert g
    host.process_mergeable_unitigs(g)
    mine_gfa = gfa_text(host, C.addressof(db), g)
    # ---- the reference
    rdb, _ = ref.extract(bases, off, k, s)
    ref.L.sr_db_stat.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    rstat1 = stat_lines(ref.L.sr_db_stat, rdb)
    rscm = ref.collect(rdb)
    rg = ref.graph(rdb, rscm, 0, 0.0)
    ref.L.ref_write_gfa2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    assert ref.L.ref_write_gfa2(rdb, rg, 1, 1, b"/dev/null") == 0
    ref.L.ref_read_ec.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int]
    ref.L.ref_read_ec(rdb, rg, 0.02, mkc, mkc * 10, mkc, 0.35, 2)
    rstat2 = stat_lines(ref.L.sr_db_stat, rdb)
    ref.free(g=rg)
    rg = ref.graph(rdb, rscm, mkc, 0.35)
    ref.unitig(rg)
    want_gfa = ref_gfa_text(ref, rdb, rg)
    assert stat1 == rstat1
    assert stat2 == rstat2 and stat2 != stat1          # error correction changed the counts, and both sides agree on how
    assert mine_gfa.count(b"\nS\t") > 0
    assert mine_gfa == want_gfa, first_diff(mine_gfa, want_gfa)
    host.scg_destroy(g)
    host.syncmer_db_destroy(scm)
    host.sr_db_clean(C.byref(db))
    ref.free(g=rg)
    ref.free(rdb, rscm)
