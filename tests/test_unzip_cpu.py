"""Repeat unzipping of the host layer (oatk_b200/host/unzip_gpu.c: scg_multiplex, scg_demultiplex; coverage_gpu.c:
scg_update_utg_cov) against the UNMODIFIED reference's (syncasm.c:1090, 1486, 682), in the order run_syncasm.c:207-262
runs them -- every step on our side done by our code, on the reference's side by its own, graphs and alignment records
compared after each. Needs oracle/_ref/libref.so; no GPU."""
import ctypes as C
import math
import numpy as np
import pytest
from oatk_b200 import synth
from pyoracle import pack_reads
from test_alignment_cpu import _sample, _bind, _flat, _same, _genome
from test_cleaning_cpu import _utg


@pytest.fixture(scope="module")
def host():
    from oatk_b200.host import build_host
    try:
        L = C.CDLL(build_host.build())
    except OSError as e:
        pytest.skip("host layer not loadable: %s" % e)
    L.scg_read_alignment.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.scg_ra_v_destroy.argtypes = [C.c_void_p]
    L.scg_ra_utg_coverage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.scg_ra_arc_coverage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.scg_update_utg_cov.argtypes = [C.c_void_p]
    L.scg_multiplex.restype = C.c_int
    L.scg_multiplex.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_double]
    L.scg_demultiplex.argtypes = [C.c_void_p]
    L.asmg_remove_weak_crosslink.restype = C.c_uint64
    L.asmg_remove_weak_crosslink.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int]
    L.process_mergeable_unitigs.argtypes = [C.c_void_p]
    L.scg_consensus.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return L


CASES = [
    # kind, k, s, min_k_cov, arc fraction, reads, read lengths, error, seed
    ("repeats", 201, 15, 3, 0.1, 900, (11000, 2500), 0.0002, 7),
    ("repeats", 101, 11, 2, 0.0, 1200, (10000, 1500), 0.0003, 8),
    ("repeats", 301, 21, 3, 0.2, 700, (12000, 4000), 0.0001, 9),
    ("diploid", 201, 15, 3, 0.2, 500, (9000, 3000), 0.0002, 10),
    ("mixture", 201, 15, 2, 0.05, 600, (9000, 3000), 0.0002, 11),
]

DONE = []


@pytest.mark.parametrize("kind,k,s,mkc,af,n,L,err,seed", CASES)
def test_unzip_matches_reference(host, ref, kind, k, s, mkc, af, n, L, err, seed):
    R = _bind(ref)
    R.ref_update_utg_cov.argtypes = [C.c_void_p]
    R.ref_multiplex.restype = C.c_int
    R.ref_multiplex.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_double]
    R.ref_demultiplex.argtypes = [C.c_void_p]
    R.ref_weak_crosslink.restype = C.c_uint64
    R.ref_weak_crosslink.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
    R.ref_write_gfa2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    rng = np.random.default_rng(seed)
    genomes = _genome(kind, rng)
    reads = _sample(rng, genomes, n // 2, L[0], err) + _sample(rng, genomes, n - n // 2, L[1], err) + synth.adversarial_reads(3, k, s)
    bases, off = pack_reads(reads)
    rdb, _ = ref.extract(bases, off, k, s)
    rscm = ref.collect(rdb)
    g1, g2 = ref.graph(rdb, rscm, mkc, af), ref.graph(rdb, rscm, mkc, af)
    assert g1 and g2
    for g in (g1, g2):
        ref.unitig(g)
        assert R.ref_write_gfa2(rdb, g, 0, 0, b"/dev/null") == 0

    def same(what):
        d1, d2 = ref.graph_dump(g1), ref.graph_dump(g2)
        for f in d1:
            assert d1[f].shape == d2[f].shape and np.array_equal(d1[f], d2[f]), (what, f)

    ours, theirs = R.ref_ra_new(), R.ref_ra_new()
    max_n_scm = math.ceil(30000.0 / k)
    rounds, updated, total = 0, 1, 0
    while updated and rounds < 3:
        rounds += 1
        host.scg_read_alignment(rdb, ours, g1, 3, 1)
        R.ref_read_alignment(rdb, theirs, g2, 2, 1)
        _same(_flat(R, ours), _flat(R, theirs), "alignment, round %d" % rounds)
        host.scg_update_utg_cov(g1)
        R.ref_update_utg_cov(g2)
        same("unitig coverage, round %d" % rounds)
        updated, u2 = host.scg_multiplex(g1, ours, max_n_scm, 10, .3), R.ref_multiplex(g2, theirs, max_n_scm, 10, .3)
        assert updated == u2
        same("multiplex, round %d" % rounds)
        total += updated
    host.scg_read_alignment(rdb, ours, g1, 2, 1)
    R.ref_read_alignment(rdb, theirs, g2, 2, 1)
    _same(_flat(R, ours), _flat(R, theirs), "alignment after unzipping")
    host.scg_ra_arc_coverage(g1, rdb, ours, 0, 0)
    R.ref_ra_arc_coverage(g2, rdb, theirs, 0)
    same("arc coverage")
    assert host.asmg_remove_weak_crosslink(_utg(g1), .3, 10, 0, 0) == R.ref_weak_crosslink(g2, .3, 10, 0)
    same("weak links")
    host.scg_demultiplex(g1)
    R.ref_demultiplex(g2)
    same("demultiplex")
    host.scg_read_alignment(rdb, ours, g1, 2, 0)
    R.ref_read_alignment(rdb, theirs, g2, 2, 0)
    _same(_flat(R, ours), _flat(R, theirs), "final alignment")
    host.scg_ra_utg_coverage(g1, rdb, ours, 0)
    R.ref_ra_utg_coverage(g2, rdb, theirs)
    host.scg_ra_arc_coverage(g1, rdb, ours, 1, 0)
    R.ref_ra_arc_coverage(g2, rdb, theirs, 1)
    same("final coverage")
    print(kind, dict(rounds=rounds, dropped_pairings=total))
    DONE.append(total)
    host.scg_ra_v_destroy(ours)
    R.ref_ra_free(theirs)
    ref.free(g=g1)
    ref.free(g=g2)
    ref.free(rdb, rscm)


def test_something_was_unzipped():
    if not DONE:
        pytest.skip("parity cases did not run")
    assert sum(1 for t in DONE if t > 0) >= 2
