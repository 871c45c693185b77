"""The graph stage of syncasm() (oatk_b200/host/run_syncasm_gpu.c: .utg.gfa -> clean-up -> unzipping -> final coverages ->
.utg.final.gfa) against the UNMODIFIED reference's whole command (run_syncasm.c:52 syncasm(), compiled into
oracle/_ref/libref.so) on the same FASTA: both GFA files must be byte-identical. Our side starts from structures the
reference's front half built (reads, syncmer database, corrected reads, graph, unitigs -- the device rows have their own
GPU tests), so this runs without a GPU; tests/test_gpu_syncasm.py runs our syncasm() end to end."""
import ctypes as C
import os
import tempfile
import numpy as np
import pytest
from test_alignment_cpu import _sample, _genome
from test_cleaning_cpu import _genome as _genome2


def genomes_for(kind, rng):
    if kind == "linear":                                 # one linear molecule: a single unitig and no arc at all
        return [bytes(b"ACGT"[i] for i in rng.integers(0, 4, 30000))]
    return _genome2(kind, rng) if kind in ("minor", "branches", "chimera") else _genome(kind, rng)


@pytest.fixture(scope="module")
def host():
    from oatk_b200.host import build_host
    try:
        L = C.CDLL(build_host.build())
    except OSError as e:
        pytest.skip("host layer not loadable: %s" % e)
    L.oatk_syncasm_graph_stage.restype = C.c_int
    L.oatk_syncasm_graph_stage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.scg_ra_v_destroy.argtypes = [C.c_void_p]
    return L


CASES = [
    # kind, k, s, min_k_cov, arc fraction, reads, read lengths, error, seed, read EC, unzip rounds, bubble, tip, weak
    ("repeats", 201, 15, 3, 0.1, 900, (11000, 2500), 0.0002, 7, 1, 3, 100000, 10000, 0.3),
    ("repeats", 101, 11, 2, 0.05, 1200, (10000, 1500), 0.0003, 8, 0, 3, 100000, 10000, 0.3),
    ("diploid", 201, 15, 3, 0.2, 500, (9000, 3000), 0.0003, 10, 1, 3, 100000, 10000, 0.3),
    ("diploid", 201, 15, 3, 0.2, 500, (9000, 3000), 0.0003, 10, 1, 0, 100000, 10000, 0.3),
    ("minor", 201, 15, 2, 0.05, 500, (6000, 6000), 0.0003, 11, 0, 0, 100000, 10000, 0.3),
    ("chimera", 201, 15, 2, 0.05, 2500, (5000, 5000), 0.0002, 17, 0, 1, 1000, 3000, 0.3),
    ("branches", 201, 15, 2, 0.05, 700, (6000, 6000), 0.0003, 13, 1, 2, 100000, 10000, 0.3),
    ("linear", 201, 15, 3, 0.35, 200, (6000, 6000), 0.0, 21, 1, 3, 100000, 10000, 0.3),
    ("linear", 201, 15, 3, 0.35, 200, (6000, 6000), 0.0, 21, 0, 0, 100000, 10000, 0.3),
]


def _first_diff(a, b):
    la, lb = a.split(b"\n"), b.split(b"\n")
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            j = next((t for t in range(min(len(x), len(y))) if x[t] != y[t]), min(len(x), len(y)))
            return "line %d col %d: %r vs %r" % (i, j, x[max(0, j - 40):j + 40], y[max(0, j - 40):j + 40])
    return "line counts %d vs %d" % (len(la), len(lb))


@pytest.mark.parametrize("kind,k,s,mkc,af,n,L,err,seed,ec,unzip,bubble,tip,weak", CASES)
def test_graph_stage_matches_reference_command(host, ref, kind, k, s, mkc, af, n, L, err, seed, ec, unzip, bubble, tip, weak):
    R = ref.L
    R.syncasm.restype = C.c_int
    R.syncasm.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                          C.c_double, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_int]
    R.ref_write_gfa2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    R.ref_read_ec.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_int]
    R.ref_ra_new.restype = C.c_void_p
    R.ref_make_graph.restype = C.c_void_p
    rng = np.random.default_rng(seed)
    genomes = genomes_for(kind, rng)
    if kind == "linear":
        g0 = genomes[0]
        reads = [g0[p:p + L[0]] for p in (int(x) for x in rng.integers(0, len(g0) - L[0], n))]
    else:
        reads = _sample(rng, genomes, n // 2, L[0], err) + _sample(rng, genomes, n - n // 2, L[1], err)
    tmp = tempfile.mkdtemp()
    fa = os.path.join(tmp, "reads.fa")
    with open(fa, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">r%d\n%s\n" % (i, r))
    files = (C.c_char_p * 1)(fa.encode())
    p_ref, p_ours = os.path.join(tmp, "ref"), os.path.join(tmp, "ours")
    assert R.syncasm(files, 1, 0, k, s, bubble, tip, mkc, af, weak, ec, unzip, 2, p_ref.encode(), None, 0) == 0

    rdb = R.ref_extract_file(fa.encode(), k, s, 2, 0)
    scm = R.ref_collect(rdb)
    if ec:
        g = R.ref_make_graph(rdb, scm, 0, 0.0)
        assert R.ref_write_gfa2(rdb, g, 1, 1, b"/dev/null") == 0
        R.ref_read_ec(rdb, g, 0.02, mkc, mkc * 10, mkc, af, 2)
        R.ref_scg_free(g)
    g = R.ref_make_graph(rdb, scm, mkc, af)
    R.ref_unitig(g)
    ra = R.ref_ra_new()
    assert host.oatk_syncasm_graph_stage(rdb, g, ra, bubble, tip, weak, unzip, 3, p_ours.encode(), 0) == 0
    for suffix in (".utg.gfa", ".utg.final.gfa"):
        a, b = open(p_ours + suffix, "rb").read(), open(p_ref + suffix, "rb").read()
        assert b.count(b"\nS\t") > 0 or kind == "linear"      # (everything of a short linear molecule ends up trimmed as tips)
        assert a == b, (suffix, _first_diff(a, b))
        os.unlink(p_ours + suffix)
        os.unlink(p_ref + suffix)
    host.scg_ra_v_destroy(ra)
    R.ref_scg_free(g)
    R.ref_scm_db_free(scm)
    R.ref_sr_db_free(rdb)
    os.unlink(fa)
    os.rmdir(tmp)
