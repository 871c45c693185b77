"""Graph clean-up of the host layer (oatk_b200/host/cleaning_gpu.c: asmg_pop_bubble, asmg_remove_weak_crosslink,
asmg_drop_tip) against the UNMODIFIED reference's (graph.c:855, 698, 607), driven the way run_syncasm.c:178-192 and
273-282 drive them: on the unitig graph after consensus, repeated until nothing changes, the graph compared after
every pass. Needs oracle/_ref/libref.so; no GPU."""
import ctypes as C
import numpy as np
import pytest
from oatk_b200 import synth
from pyoracle import pack_reads
from test_alignment_cpu import _mutate, _sample, _bind


@pytest.fixture(scope="module")
def host():
    from oatk_b200.host import build_host
    try:
        L = C.CDLL(build_host.build())
    except OSError as e:
        pytest.skip("host layer not loadable: %s" % e)
    L.asmg_drop_tip.restype = C.c_uint64
    L.asmg_drop_tip.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_int, C.c_int, C.c_int]
    L.asmg_remove_weak_crosslink.restype = C.c_uint64
    L.asmg_remove_weak_crosslink.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int]
    L.asmg_pop_bubble.restype = C.c_uint64
    L.asmg_pop_bubble.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int]
    L.process_mergeable_unitigs.argtypes = [C.c_void_p]
    return L


def _utg(g):
    """scg_t.utg_asmg: second pointer of the struct"""
    return C.cast(g, C.POINTER(C.c_void_p))[1]


CHIMERA_K = 201


def _genome(kind, rng):
    rnd = lambda n: bytes(b"ACGT"[i] for i in rng.integers(0, 4, n))
    if kind == "minor":          # a 1-in-9 variant molecule: low-coverage bubbles
        a = rnd(40000)
        return [a] * 8 + [_mutate(rng, a, 0.0007)]
    if kind == "branches":       # rare molecules that leave the main one and end: tips of several lengths; a chimera: a weak cross link
        a = rnd(50000)
        return [a] * 12 + [a[5000:9000] + rnd(2500), rnd(1500) + a[20000:24000], a[30000:33000] + rnd(6000), a[10000:13000] + a[40000:43000]]
    if kind == "chimera":        # rare recombinants across repeats a little shorter than k: no k-mer of their own at the
        parts, rare = [], []     # junction (or hardly any), so the graph gets a direct, thinly covered arc between the flanks
        for i in range(6):
            rep = rnd(CHIMERA_K - 6 - 2 * i)
            a, b, c, d = rnd(5000), rnd(5000), rnd(5000), rnd(5000)
            parts += [a, rep, b, c, rep, d]
            rare += [a[-3500:] + rep + d[:3500], c[-3500:] + rep + b[:3500]]
        return [b"".join(parts)] * (8 * len(rare)) + rare
    if kind == "diploid":
        a = rnd(40000)
        return [a, _mutate(rng, a, 0.001)]
    raise ValueError(kind)


CASES = [
    # kind, k, s, min_k_cov, arc fraction, reads, read length, error, seed, bubble, tip, weak
    ("minor", 201, 15, 2, 0.05, 500, 6000, 0.0003, 11, 100000, 10000, 0.3),
    ("minor", 101, 11, 2, 0.0, 500, 4000, 0.0005, 12, 100000, 10000, 0.3),
    ("branches", 201, 15, 2, 0.05, 700, 6000, 0.0003, 13, 100000, 10000, 0.3),
    ("branches", 101, 11, 2, 0.0, 800, 4000, 0.0005, 14, 100000, 3000, 0.5),
    ("chimera", 201, 15, 2, 0.05, 2500, 5000, 0.0002, 17, 1000, 3000, 0.3),
    ("chimera", 201, 21, 2, 0.0, 3000, 4000, 0.0003, 18, 500, 2000, 0.4),
    ("diploid", 201, 15, 2, 0.1, 400, 6000, 0.0005, 15, 100000, 10000, 0.3),
    ("diploid", 101, 11, 2, 0.0, 300, 3000, 0.002, 16, 2000, 500, 0.3),
]

DONE = []


@pytest.mark.parametrize("kind,k,s,mkc,af,n,L,err,seed,bubble,tip,weak", CASES)
def test_cleanup_matches_reference(host, ref, kind, k, s, mkc, af, n, L, err, seed, bubble, tip, weak):
    R = _bind(ref)
    R.ref_drop_tip.restype = C.c_uint64
    R.ref_drop_tip.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_int, C.c_int]
    R.ref_weak_crosslink.restype = C.c_uint64
    R.ref_weak_crosslink.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
    R.ref_pop_bubble.restype = C.c_uint64
    R.ref_pop_bubble.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int]
    R.ref_write_gfa2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    rng = np.random.default_rng(seed)
    reads = _sample(rng, _genome(kind, rng), n, L, err) + synth.adversarial_reads(3, k, s)
    bases, off = pack_reads(reads)
    rdb, _ = ref.extract(bases, off, k, s)
    rscm = ref.collect(rdb)
    g1, g2 = ref.graph(rdb, rscm, mkc, af), ref.graph(rdb, rscm, mkc, af)
    assert g1 and g2
    for g in (g1, g2):
        ref.unitig(g)
        assert R.ref_write_gfa2(rdb, g, 0, 0, b"/dev/null") == 0          # lengths and overlaps in bases

    def same(what):
        d1, d2 = ref.graph_dump(g1), ref.graph_dump(g2)
        for f in d1:
            assert np.array_equal(d1[f], d2[f]), (what, f)

    tot = dict(bubbles=0, weak=0, tips=0, rounds=0)
    for with_bubbles in (False, True):                                    # before unzipping only tips go (run_syncasm.c:183-190)
        cleaned = 1
        while cleaned:
            cleaned = 0
            if with_bubbles:
                a, b = host.asmg_pop_bubble(_utg(g1), bubble, 0, 0, 1, 0, 0), R.ref_pop_bubble(g2, bubble, 0, 0, 1, 0)
                assert a == b
                same("bubbles")
                tot["bubbles"] += a & 0xffffffff
                cleaned += a
                a, b = host.asmg_remove_weak_crosslink(_utg(g1), weak, 10, 0, 0), R.ref_weak_crosslink(g2, weak, 10, 0)
                assert a == b
                same("weak links")
                tot["weak"] += a
                cleaned += a
            a, b = host.asmg_drop_tip(_utg(g1), 2 ** 31 - 1, tip, 1, 0, 0), R.ref_drop_tip(g2, 2 ** 31 - 1, tip, 1, 0)
            assert a == b
            same("tips")
            tot["tips"] += a
            cleaned += a
            tot["rounds"] += 1
        for g in (g1, g2):
            ref.unitig(g)
            assert R.ref_write_gfa2(rdb, g, 0, 0, b"/dev/null") == 0
        same("merged")
    # the variants that finalize on their own, and tip protection off / bubble size limits
    a, b = host.asmg_pop_bubble(_utg(g1), bubble, 3, 1, 0, 1, 0), R.ref_pop_bubble(g2, bubble, 3, 1, 0, 1)
    assert a == b
    same("bubbles, max_del, unprotected")
    a, b = host.asmg_drop_tip(_utg(g1), 3, 10 * tip, 0, 1, 0), R.ref_drop_tip(g2, 3, 10 * tip, 0, 1)
    assert a == b
    same("tips, unprotected")
    print(kind, tot)
    DONE.append(tot)
    ref.free(g=g1)
    ref.free(g=g2)
    ref.free(rdb, rscm)


def test_cases_had_something_to_clean():
    if not DONE:
        pytest.skip("parity cases did not run")
    for what in ("bubbles", "weak", "tips"):
        assert sum(t[what] for t in DONE) > 0, what
