"""f3 on the CPU: read -> unitig-graph alignment and the coverage estimators of the host layer
(oatk_b200/host/alignment_gpu.c, coverage_gpu.c) against the UNMODIFIED reference's scg_read_alignment,
scg_ra_utg_coverage and scg_ra_arc_coverage (alignment.c:596, syncasm.c:1882, 2067) on graphs the reference built
(structs are byte-compatible, so both sides read the same sr_db / scg and either side's records feed the other).
Needs oracle/_ref/libref.so; no GPU."""
import ctypes as C
import numpy as np
import pytest
from oatk_b200 import synth
from pyoracle import pack_reads


@pytest.fixture(scope="module")
def host():
    from oatk_b200.host import build_host
    try:
        L = C.CDLL(build_host.build())
    except OSError as e:
        pytest.skip("host layer not loadable: %s" % e)
    L.scg_read_alignment.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.scg_read_alignment.restype = None
    L.scg_ra_v_destroy.argtypes = [C.c_void_p]
    L.scg_ra_utg_coverage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.scg_ra_arc_coverage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    return L


def _bind(ref):
    L = ref.L
    L.ref_ra_new.restype = C.c_void_p
    L.ref_ra_free.argtypes = [C.c_void_p]
    L.ref_read_alignment.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.ref_ra_n.restype = C.c_uint64
    L.ref_ra_n.argtypes = [C.c_void_p]
    L.ref_ra_total.restype = C.c_uint64
    L.ref_ra_total.argtypes = [C.c_void_p]
    L.ref_ra_flatten.argtypes = [C.c_void_p] * 5
    L.ref_ra_utg_coverage.argtypes = [C.c_void_p] * 3
    L.ref_ra_arc_coverage.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return L


def _flat(L, ra):
    n, t = int(L.ref_ra_n(ra)), int(L.ref_ra_total(ra))
    sid, cnt, s, frg = np.zeros(n + 1, np.uint64), np.zeros(n + 1, np.uint32), np.zeros(n + 1, np.float64), np.zeros((t + 1, 5), np.uint64)
    L.ref_ra_flatten(ra, sid.ctypes.data, cnt.ctypes.data, s.ctypes.data, frg.ctypes.data)
    return sid[:n], cnt[:n], s[:n], frg[:t]


def _same(a, b, what):
    for x, y, name in zip(a, b, ("sid", "n", "s", "fragments")):
        assert x.shape == y.shape, (what, name, x.shape, y.shape)
        assert np.array_equal(x, y), (what, name, np.flatnonzero((x != y).reshape(len(x), -1).any(1))[:5])


def _mutate(rng, seq, rate):
    """point differences and small indels at `rate` per base: the second haplotype"""
    out = bytearray()
    for ch in seq:
        r = rng.random()
        if r < rate / 3:
            out.append(b"ACGT"[(b"ACGT".index(ch) + int(rng.integers(1, 4))) % 4])
        elif r < 2 * rate / 3:
            continue
        elif r < rate:
            out.append(ch)
            out.append(b"ACGT"[int(rng.integers(0, 4))])
        else:
            out.append(ch)
    return bytes(out)


def _sample(rng, genomes, n, L, err):
    """reads of length L from circular genomes, random strand, sub/ins/del errors"""
    reads = []
    for _ in range(n):
        g = genomes[int(rng.integers(0, len(genomes)))]
        p = int(rng.integers(0, len(g)))
        r = (g + g)[p:p + L] if 2 * len(g) >= p + L else g
        if err:
            r = _mutate(rng, r, err)
        reads.append(r if rng.integers(0, 2) else synth.revcomp(r))
    return reads


def _genome(kind, rng):
    rnd = lambda n: bytes(b"ACGT"[i] for i in rng.integers(0, 4, n))
    if kind == "plain":
        return [rnd(60000)]
    if kind == "diploid":                       # bubbles: two haplotypes 0.1 % apart
        a = rnd(50000)
        return [a, _mutate(rng, a, 0.001)]
    if kind == "repeats":                       # a 6 kb repeat in three copies (one inverted) and a tandem array
        rep, unit = rnd(6000), rnd(700)
        return [rnd(15000) + rep + rnd(12000) + synth.revcomp(rep) + rnd(9000) + unit * 9 + rnd(8000) + rep + rnd(10000)]
    if kind == "mixture":                       # an abundant and a rare molecule that share a segment
        shared = rnd(8000)
        return [rnd(20000) + shared + rnd(15000)] * 4 + [rnd(12000) + shared + rnd(9000)]
    raise ValueError(kind)


CASES = [
    # kind, k, s, min_k_cov, arc fraction, reads, read length, error rate, seed
    ("plain", 501, 31, 3, 0.35, 150, 9000, 0.0001, 1),
    ("diploid", 301, 21, 3, 0.2, 300, 8000, 0.0002, 2),
    ("repeats", 201, 15, 3, 0.1, 400, 7000, 0.0003, 3),
    ("repeats", 101, 11, 2, 0.0, 300, 4000, 0.0005, 4),
    ("mixture", 301, 21, 2, 0.05, 350, 8000, 0.0002, 5),
    ("diploid", 101, 11, 2, 0.0, 250, 3000, 0.001, 6),
    ("repeats", 201, 15, 3, 0.1, 900, (11000, 2500), 0.0002, 7),
    ("repeats", 101, 11, 2, 0.0, 1200, (10000, 1500), 0.0003, 8),
]


SEEN = []


@pytest.mark.parametrize("kind,k,s,mkc,af,n,L,err,seed", CASES)
@pytest.mark.parametrize("unitig", ["unitigs", "syncmers", "unzipped"])
def test_alignment_and_coverage_match_reference(host, ref, kind, k, s, mkc, af, n, L, err, seed, unitig):
    R = _bind(ref)
    rng = np.random.default_rng(seed)
    genomes = _genome(kind, rng)
    if isinstance(L, tuple):                    # long reads span the repeats (they get unzipped), short ones fall inside
        reads = _sample(rng, genomes, n // 2, L[0], err) + _sample(rng, genomes, n - n // 2, L[1], err)
    else:
        reads = _sample(rng, genomes, n, L, err)
    reads += synth.adversarial_reads(3, k, s)
    bases, off = pack_reads(reads)
    rdb, _ = ref.extract(bases, off, k, s)
    rscm = ref.collect(rdb)
    g1, g2 = ref.graph(rdb, rscm, mkc, af), ref.graph(rdb, rscm, mkc, af)
    assert g1 and g2
    if unitig != "syncmers":
        ref.unitig(g1)
        ref.unitig(g2)
    if unitig == "unzipped":
        # the reference's own repeat unzipping on both copies: unitigs duplicated along read paths, so that
        # syncmers occur on several unitigs and reads have several equally good alignments
        R.ref_multiplex_rounds.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        u1, u2 = R.ref_multiplex_rounds(rdb, g1, 3, 2), R.ref_multiplex_rounds(rdb, g2, 3, 2)
        assert u1 == u2
        d1, d2 = ref.graph_dump(g1), ref.graph_dump(g2)
        assert all(np.array_equal(d1[f], d2[f]) for f in d1)
    ours, theirs = R.ref_ra_new(), R.ref_ra_new()
    host.scg_read_alignment(rdb, ours, g1, 3, 0)
    R.ref_read_alignment(rdb, theirs, g2, 2, 0)
    a, b = _flat(R, ours), _flat(R, theirs)
    assert len(b[0]) > n // 6, "hardly any read aligned: the case tests nothing"
    _same(a, b, "alignment")
    stats = dict(records=len(b[0]), multi=int(len(b[0]) - len(np.unique(b[0]))), frac=int((np.modf(b[2])[0] > 1e-9).sum()), chains=int((b[1] > 1).sum()))

    # re-alignment for unzipping: only reads that spanned > 2 unitigs, and only if the score does not drop
    host.scg_read_alignment(rdb, ours, g1, 2, 1)
    R.ref_read_alignment(rdb, theirs, g2, 3, 1)
    _same(_flat(R, ours), _flat(R, theirs), "for_unzip")
    host.scg_read_alignment(rdb, ours, g1, 4, 0)
    R.ref_read_alignment(rdb, theirs, g2, 1, 0)

    # coverage estimators: ours on g1 with our records, the reference's on g2 with its own
    host.scg_ra_utg_coverage(g1, rdb, ours, 0)
    R.ref_ra_utg_coverage(g2, rdb, theirs)
    d1, d2 = ref.graph_dump(g1), ref.graph_dump(g2)
    assert np.array_equal(d1["vtx_flags"], d2["vtx_flags"]), "unitig coverage"
    for refine in (0, 1):
        host.scg_ra_arc_coverage(g1, rdb, ours, refine, 0)
        R.ref_ra_arc_coverage(g2, rdb, theirs, refine)
        d1, d2 = ref.graph_dump(g1), ref.graph_dump(g2)
        assert np.array_equal(d1["arcs"], d2["arcs"]), "arc coverage, refine=%d" % refine
    print(kind, unitig, stats)
    SEEN.append((unitig, stats))
    host.scg_ra_v_destroy(ours)
    R.ref_ra_free(theirs)
    ref.free(g=g1)
    ref.free(g=g2)
    ref.free(rdb, rscm)


def test_cases_were_ambiguous():
    """the cases above must have held reads with several records and chains over several unitigs"""
    if not SEEN:
        pytest.skip("parity cases did not run")
    assert sum(st["multi"] for _, st in SEEN) > 0
    assert sum(st["chains"] for _, st in SEEN) > 0
