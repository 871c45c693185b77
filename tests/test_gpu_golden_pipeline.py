"""The whole path up to <out>.utg.gfa through the host layer on the GPU (extraction, counting, statistics and arc
tallies on the device; graph, consensus and read error correction on the host) against the golden outputs of the
unmodified reference in tests/golden/pipeline.json (made by tests/golden/make_golden_pipeline.py). Needs no
reference build at run time."""
import ctypes as C
import hashlib
import json
import os
import sys
import tempfile
import pytest
from pyoracle import pack_reads

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_pipeline import CASES, reads_of   # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "pipeline.json")))


class SrDb(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.c_void_p), ("k", C.c_int), ("s", C.c_int), ("stats", C.c_void_p)]


@pytest.fixture(scope="module")
def host():
    from oatk_b200.host import build_host
    L = C.CDLL(build_host.build())
    sig = {"sr_read_mem": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64], "sr_db_init": [C.c_void_p, C.c_int, C.c_int],
           "sr_db_stat": [C.c_void_p, C.c_void_p, C.c_int], "collect_syncmer_from_reads": [C.c_void_p], "sr_db_clean": [C.c_void_p],
           "syncmer_db_destroy": [C.c_void_p], "make_syncmer_graph": [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double],
           "process_mergeable_unitigs": [C.c_void_p], "scg_destroy": [C.c_void_p], "scg_consensus": [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p],
           "read_error_correction": [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int, C.c_void_p, C.c_int]}
    for f, a in sig.items():
        getattr(L, f).argtypes = a
    L.collect_syncmer_from_reads.restype = C.c_void_p
    L.make_syncmer_graph.restype = C.c_void_p
    L.scg_consensus.restype = None
    L.read_error_correction.restype = None
    return L


def _to_file(fn):
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    path = tempfile.mktemp()
    fo = libc.fopen(path.encode(), b"w")
    fn(fo)
    libc.fclose(fo)
    data = open(path, "rb").read()
    os.unlink(path)
    return data


@pytest.mark.parametrize("ec", [False, True])
@pytest.mark.parametrize("case", sorted(CASES))
def test_utg_gfa_matches_golden(host, case, ec):
    seed, G, n, L, err, k, s, mkc = CASES[case]
    want = GOLD[case]["ec" if ec else "no_ec"]
    bases, off = pack_reads(reads_of(case))
    db = SrDb()
    host.sr_db_init(C.byref(db), k, s)
    assert host.sr_read_mem(C.byref(db), bases.ctypes.data, off.ctypes.data, None, len(off) - 1) == 0
    stat1 = _to_file(lambda fo: host.sr_db_stat(C.byref(db), fo, 0)).decode().splitlines()
    assert stat1 == want["stat1"]
    scm = host.collect_syncmer_from_reads(C.byref(db))
    assert scm
    if ec:
        g = host.make_syncmer_graph(C.byref(db), scm, 0, 0.0)
        host.scg_consensus(C.byref(db), g, 1, 1, None)
        host.read_error_correction(C.byref(db), g, 0.02, mkc, mkc * 10, mkc, 0.35, 3, None, 0)
        stat2 = _to_file(lambda fo: host.sr_db_stat(C.byref(db), fo, 0)).decode().splitlines()
        assert stat2 == want["stat2"]
        host.scg_destroy(g)
    g = host.make_syncmer_graph(C.byref(db), scm, mkc, 0.35)
    assert g
    host.process_mergeable_unitigs(g)
    gfa = _to_file(lambda fo: host.scg_consensus(C.byref(db), g, 0, 0, fo))
    assert (gfa.count(b"\nS\t"), gfa.count(b"\nL\t"), len(gfa)) == (want["S"], want["L"], want["bytes"])
    assert hashlib.md5(gfa).hexdigest() == want["gfa_md5"]
    host.scg_destroy(g)
    host.syncmer_db_destroy(scm)
    host.sr_db_clean(C.byref(db))
