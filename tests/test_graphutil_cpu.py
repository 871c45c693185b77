"""The remaining graph.h queries of the host layer (oatk_b200/host/graphutil_gpu.c, asmg_uext_arc_group in
cleaning_gpu.c) against the UNMODIFIED reference's on the same graphs. No GPU."""
import ctypes as C
import numpy as np
import pytest
from oatk_b200 import synth
from pyoracle import pack_reads
from test_alignment_cpu import _sample, _genome, _bind
from test_cleaning_cpu import _utg, _genome as _genome2
from test_report_cpu import _text


@pytest.fixture(scope="module")
def host():
    from oatk_b200.host import build_host
    try:
        return C.CDLL(build_host.build())
    except OSError as e:
        pytest.skip("host layer not loadable: %s" % e)


def _sig(L):
    L.asmg_arc_is_sorted.argtypes = [C.c_void_p]
    L.asmg_vtx_list.restype = C.POINTER(C.c_uint64)
    L.asmg_vtx_list.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.asmg_uext_arc_group.restype = C.POINTER(C.c_uint32)
    L.asmg_uext_arc_group.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    L.asmg_subgraph.restype = C.POINTER(C.c_uint32)
    L.asmg_subgraph.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint32), C.c_int]
    L.asmg_tarjans_scc.argtypes = [C.c_void_p, C.c_void_p]
    L.asmg_path_exists.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    L.asmg_max_link_id.restype = C.c_uint64
    L.asmg_max_link_id.argtypes = [C.c_void_p]
    return L


@pytest.mark.parametrize("kind,k,s,mkc,seed", [("repeats", 101, 11, 2, 4), ("diploid", 201, 15, 3, 2), ("branches", 201, 15, 2, 13), ("chimera", 201, 15, 2, 17)])
def test_graph_queries_match_reference(host, ref, kind, k, s, mkc, seed):
    R = _bind(ref)
    H = _sig(host)
    _sig(R)
    R.ref_write_gfa2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    rng = np.random.default_rng(seed)
    genomes = _genome2(kind, rng) if kind in ("branches", "chimera") else _genome(kind, rng)
    reads = _sample(rng, genomes, 600, 5000, 0.0005) + synth.adversarial_reads(3, k, s)
    bases, off = pack_reads(reads)
    rdb, _ = ref.extract(bases, off, k, s)
    rscm = ref.collect(rdb)
    g1, g2 = ref.graph(rdb, rscm, mkc, 0.05), ref.graph(rdb, rscm, mkc, 0.05)
    for unitig in (False, True):
        if unitig:
            ref.unitig(g1)
            ref.unitig(g2)
        for g in (g1, g2):
            assert R.ref_write_gfa2(rdb, g, 0, 1, b"/dev/null") == 0            # lengths, overlaps, sequences
        u1, u2 = _utg(g1), _utg(g2)
        nv = int(ref.L.ref_graph_n_vtx(g1))
        assert H.asmg_arc_is_sorted(u1) == R.asmg_arc_is_sorted(u2) == 1
        n1, n2 = C.c_uint64(), C.c_uint64()
        l1, l2 = H.asmg_vtx_list(u1, C.byref(n1)), R.asmg_vtx_list(u2, C.byref(n2))
        assert n1.value == n2.value and l1[:n1.value] == l2[:n2.value]
        for no_seq in (0, 1):
            assert _text(H, "asmg_print", C.c_void_p(u1), "FO", no_seq) == _text(R, "asmg_print", C.c_void_p(u2), "FO", no_seq)
        m1, m2 = C.c_uint32(), C.c_uint32()
        a1, a2 = H.asmg_uext_arc_group(u1, C.byref(m1)), R.asmg_uext_arc_group(u2, C.byref(m2))
        nl = int(H.asmg_max_link_id(u1)) + 1
        assert m1.value == m2.value and a1[:nl] == a2[:nl]
        s1, s2 = np.zeros(2 * nv, np.int32), np.zeros(2 * nv, np.int32)
        assert H.asmg_tarjans_scc(u1, s1.ctypes.data) == R.asmg_tarjans_scc(u2, s2.ctypes.data)
        assert np.array_equal(s1, s2)
        # neighbourhoods and reachability with and without limits
        for trial in range(40):
            seeds = (C.c_uint32 * 2)(int(rng.integers(0, nv)), int(rng.integers(0, nv + 3)))
            step = int(rng.choice([0, 1, 2, 5]))
            dist = int(rng.choice([0, 500, 3000, 20000]))
            c1, c2 = C.c_uint32(), C.c_uint32()
            v1, v2 = H.asmg_subgraph(u1, seeds, 2, step, dist, C.byref(c1), 0), R.asmg_subgraph(u2, seeds, 2, step, dist, C.byref(c2), 0)
            assert c1.value == c2.value and v1[:c1.value] == v2[:c2.value], (trial, step, dist)
            src, snk = int(rng.integers(0, 2 * nv)), int(rng.integers(0, 2 * nv))
            st1, st2, d1, d2 = C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_uint64()
            e1 = H.asmg_path_exists(u1, src, snk, step * 3, dist * 4, C.byref(st1), C.byref(d1))
            e2 = R.asmg_path_exists(u2, src, snk, step * 3, dist * 4, C.byref(st2), C.byref(d2))
            assert (e1, st1.value, d1.value) == (e2, st2.value, d2.value)
    # the variant that prunes the graph to the neighbourhood
    seeds = (C.c_uint32 * 1)(0)
    c1, c2 = C.c_uint32(), C.c_uint32()
    H.asmg_subgraph(_utg(g1), seeds, 1, 3, 0, C.byref(c1), 1)
    R.asmg_subgraph(_utg(g2), seeds, 1, 3, 0, C.byref(c2), 1)
    assert c1.value == c2.value
    d1, d2 = ref.graph_dump(g1), ref.graph_dump(g2)
    for f in d1:
        assert np.array_equal(d1[f], d2[f]), f
    ref.free(g=g1)
    ref.free(g=g2)
    ref.free(rdb, rscm)
