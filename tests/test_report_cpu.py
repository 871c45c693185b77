"""The reporting helpers of the host layer (oatk_b200/host/report_gpu.c) against the UNMODIFIED reference's on the same
structs: text compared byte for byte, recounted arc coverages compared in the graph. No GPU."""
import ctypes as C
import os
import tempfile
import numpy as np
import pytest
from oatk_b200 import synth
from pyoracle import pack_reads
from test_alignment_cpu import _sample, _genome, _bind


@pytest.fixture(scope="module")
def host():
    from oatk_b200.host import build_host
    try:
        return C.CDLL(build_host.build())
    except OSError as e:
        pytest.skip("host layer not loadable: %s" % e)


def _text(lib, name, *args):
    """calls lib.name(*args with FILE* in place of the placeholder 'FO') and returns what it wrote"""
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    path = tempfile.mktemp()
    fo = libc.fopen(path.encode(), b"w")
    f = getattr(lib, name)
    f.restype = None
    f(*[C.c_void_p(fo) if a == "FO" else a for a in args])
    libc.fclose(fo)
    out = open(path, "rb").read()
    os.unlink(path)
    return out


class SrT(C.Structure):
    _fields_ = [("sid", C.c_uint64), ("sname", C.c_char_p), ("hoco_l", C.c_uint32), ("hoco_s", C.c_void_p), ("ho_rl", C.c_void_p),
                ("ho_l_rl", C.c_void_p), ("n_nucl", C.c_void_p), ("n", C.c_uint32), ("m_pos", C.c_void_p), ("s_mer", C.c_void_p), ("k_mer", C.c_void_p)]


class SrDb(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.POINTER(SrT)), ("k", C.c_int), ("s", C.c_int), ("stats", C.c_void_p)]


@pytest.mark.parametrize("kind,k,s,mkc,seed", [("repeats", 101, 11, 2, 4), ("diploid", 201, 15, 3, 2), ("mixture", 301, 21, 2, 5)])
def test_reports_match_reference(host, ref, kind, k, s, mkc, seed):
    R = _bind(ref)
    R.ref_write_gfa2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    rng = np.random.default_rng(seed)
    reads = _sample(rng, _genome(kind, rng), 300, 5000, 0.0005) + synth.adversarial_reads(3, k, s)
    bases, off = pack_reads(reads)
    rdb, _ = ref.extract(bases, off, k, s)
    rscm = ref.collect(rdb)
    g1, g2 = ref.graph(rdb, rscm, mkc, 0.05), ref.graph(rdb, rscm, mkc, 0.05)
    P = C.c_void_p
    for unitig in (False, True):
        if unitig:
            ref.unitig(g1)
            ref.unitig(g2)
        for g in (g1, g2):
            assert R.ref_write_gfa2(rdb, g, 0, 1, b"/dev/null") == 0            # sequences saved in the vertices
        for no_seq in (0, 1):
            a, b = _text(host, "scg_print", P(g1), "FO", no_seq), _text(R, "scg_print", P(g2), "FO", no_seq)
            assert a == b and a.count(b"\nS\t") > 0
        a, b = _text(host, "scg_print_unitig_syncmer_list", P(g1), "FO"), _text(R, "scg_print_unitig_syncmer_list", P(g2), "FO")
        assert a == b and len(a) > 100
        a, b = _text(host, "scg_subgraph_stat", P(g1), "FO"), _text(R, "scg_subgraph_stat", P(g2), "FO")
        assert a.replace(b"scg_subgraph_stat", b"X") == b.replace(b"scg_subgraph_stat", b"X") and a.count(b"seeding") >= 1
        # arc coverage recounted from the reads
        host.scg_arc_coverage.argtypes = [C.c_void_p, C.c_void_p]
        R.scg_arc_coverage.argtypes = [C.c_void_p, C.c_void_p]
        host.scg_arc_coverage(g1, rdb)
        R.scg_arc_coverage(g2, rdb)
        d1, d2 = ref.graph_dump(g1), ref.graph_dump(g2)
        assert np.array_equal(d1["arcs"], d2["arcs"])
    # alignment records
    ra = R.ref_ra_new()
    R.ref_read_alignment(rdb, ra, g2, 2, 0)
    a, b = _text(host, "scg_rv_print", P(ra), "FO"), _text(R, "scg_rv_print", P(ra), "FO")
    assert a == b and a.count(b"RID") > 50
    R.ref_ra_free(ra)
    # per-read printers
    db = C.cast(rdb, C.POINTER(SrDb)).contents
    shown = 0
    for i in range(0, db.n, 37):
        sr = C.byref(db.a[i])
        a, b = _text(host, "print_all_syncmers_on_seq", sr, s, k, "FO"), _text(R, "print_all_syncmers_on_seq", sr, s, k, "FO")
        assert a == b
        shown += a.count(b">")
        a, b = _text(host, "print_aligned_syncmers_on_seq", sr, k, 0, 5, "FO"), _text(R, "print_aligned_syncmers_on_seq", sr, k, 0, 5, "FO")
        assert a == b
        a, b = _text(host, "print_hoco_seq", sr, "FO"), _text(R, "print_hoco_seq", sr, "FO")
        assert a == b
    assert shown > 20
    ref.free(g=g1)
    ref.free(g=g2)
    ref.free(rdb, rscm)


def test_get_kmer_dna_seq_matches_reference(host, ref):
    """every alignment of start and length against the packed bytes, both strands"""
    rng = np.random.default_rng(9)
    packed = rng.integers(0, 256, 400, dtype=np.uint8)
    for L in (host, ref.L):
        L.get_kmer_dna_seq.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_char_p]
        L.get_kmer_dna_seq.restype = None
    for pos in list(range(0, 9)) + [int(x) for x in rng.integers(0, 1000, 40)]:
        for ln in list(range(0, 14)) + [int(x) for x in rng.integers(14, 500, 25)]:
            for rev in (0, 1):
                a, b = C.create_string_buffer(ln + 8), C.create_string_buffer(ln + 8)
                host.get_kmer_dna_seq(packed.ctypes.data, pos, ln, rev, a)
                ref.L.get_kmer_dna_seq(packed.ctypes.data, pos, ln, rev, b)
                assert a.raw[:ln] == b.raw[:ln] and a.raw[ln:] == b"\0" * 8, (pos, ln, rev)
