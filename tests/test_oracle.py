"""CPU: the oracle (oracle/sync_oracle.c) against the golden vectors produced by the unmodified
reference, and - where oracle/_ref/libref.so exists - against the live reference on fresh inputs."""
import numpy as np
import pytest
from oatk_b200 import synth
from pyoracle import pack_reads
import golden_util as gu
import parity


def test_unit_vectors(oracle):
    g = np.load(gu.GOLD + "/unit_vectors.npz")
    b = bytes((37 * i + 11) % 256 for i in range(256))
    for n, h in zip(g["murmur_len"], g["murmur"]):
        assert oracle.murmur(b[:int(n)]) == int(h)
    for x, h in zip(g["hash62_key"], g["hash62"]):
        assert oracle.hash64(int(x), (1 << 62) - 1) == int(h)
    for x, h in zip(g["hash22_key"], g["hash22"]):
        assert oracle.hash64(int(x), (1 << 22) - 1) == int(h)
    # SURVEY.md appendix A.4, printed by the reference during the survey
    assert oracle.murmur(b[:251]) == 15130853122908940662
    assert oracle.hash64(0x0123456789ABCDEF, (1 << 62) - 1) == 2609805317204882043


def test_hash64_is_a_bijection(oracle):
    """ties between window hashes mean identical s-mers only because hash64 permutes 2s-bit values"""
    for bits in (6, 10, 16):
        m = (1 << bits) - 1
        out = {oracle.hash64(x, m) for x in range(1 << bits)}
        assert len(out) == 1 << bits and max(out) <= m


@pytest.mark.parametrize("name", sorted(gu.CASES))
def test_golden(oracle, name):
    gen, args, k, s, mkc = gu.CASES[name]
    reads = gu.make_reads(gen, args)
    bases, off = pack_reads(reads)
    g = gu.load(name)
    db, f = oracle.extract(bases, off, k, s)
    assert gu.check_extract(f, g) == []
    rc, d, i, sc, kc = oracle.stat(db)
    assert rc == int(g["stat_rc"][0])
    assert np.array_equal(d, g["stat_d"], equal_nan=True) and np.array_equal(i, g["stat_i"])
    scm = oracle.collect(db, len(reads))
    assert gu.check_scm(scm, g) == []
    arcs = oracle.arcs(db, scm, mkc, 0.35)
    assert np.array_equal(arcs, gu.golden_arcs(g))
    oracle.free(db, scm)


@pytest.mark.parametrize("k,s", [(1001, 31), (301, 15), (64, 31), (40, 1), (12, 11), (5, 3), (2001, 31)])
def test_live_reference(oracle, ref, k, s):
    reads = synth.adversarial_reads(11, k, s) + synth.hifi_reads(3, 30000, 12, 6000, 0.003)
    bases, off = pack_reads(reads)
    odb, of = oracle.extract(bases, off, k, s)
    rdb, rf = ref.extract(bases, off, k, s)
    assert parity.diff(of, rf, parity.EXTRACT_FIELDS) == []
    orc, od, oi, _, _ = oracle.stat(odb)
    rrc, rd, ri = ref.stat(rdb)
    assert orc == rrc and np.array_equal(od, rd, equal_nan=True) and np.array_equal(oi, ri)
    oc, rc = oracle.collect(odb, len(reads)), ref.collect(rdb)
    assert (oc is None) == (rc is None)
    if oc is not None:
        assert parity.diff(oc, rc, parity.SCM_FIELDS) == []
    oracle.free(odb, oc)
    ref.free(rdb, rc)


def test_tie_rule_is_exercised(oracle):
    """the restatement's third CLOSE clause (reference syncmer.c:356-377) must actually reject
    positions on the adversarial set, so that the comparisons above pin it"""
    before = oracle.tie_suppressed()
    for k, s in ((101, 11), (1001, 31)):
        bases, off = pack_reads(synth.adversarial_reads(3, k, s))
        db, _ = oracle.extract(bases, off, k, s)
        oracle.free(db)
    assert oracle.tie_suppressed() > before


def test_empty_and_short(oracle):
    bases, off = pack_reads([b"", b"ACGT", b"N" * 10])
    db, f = oracle.extract(bases, off, 1001, 31)
    assert f["n_scm"].tolist() == [0, 0, 0] and f["hoco_l"].tolist() == [0, 4, 10]
    assert oracle.collect(db, 3) is None          # reference returns NULL (syncmer.c:1414-1417)
    assert oracle.stat(db)[0] == 1                # "empty syncmer collection" (syncmer.c:909-912)
    oracle.free(db)


def test_bad_parameters(oracle):
    bases, off = pack_reads([b"ACGT"])
    for k, s in ((31, 31), (10, 32), (10, 0)):
        with pytest.raises(ValueError):
            oracle.extract(bases, off, k, s)


def test_forced_collisions_are_split_in_first_seen_order(oracle):
    reads = synth.hifi_reads(9, 60000, 40, 8000, 0.002)
    bases, off = pack_reads(reads)
    db, f = oracle.extract(bases, off, 301, 15)
    full = oracle.collect(db, len(reads), 64)
    db2, _ = oracle.extract(bases, off, 301, 15)
    cut = oracle.collect(db2, len(reads), 4)
    # truncating the hash must not merge distinct k-mers: same classes, same coverages as a multiset
    assert len(cut["h"]) == len(full["h"])
    assert sorted(cut["cov"].tolist()) == sorted(full["cov"].tolist())
    assert len(np.unique(cut["h"])) <= 16
    oracle.free(db, full)
    oracle.free(db2, cut)


def test_peak_finder(oracle):
    cnt = np.zeros(1001, np.int64)
    cnt[1] = 5000
    cnt[2] = 800
    for i in range(3, 200):
        cnt[i] = int(1000 * np.exp(-((i - 60) / 12.0) ** 2)) + int(400 * np.exp(-((i - 30) / 6.0) ** 2))
    hom, het = oracle.analyze_count(cnt)
    assert (hom, het) == (60, 30)
    flat = np.zeros(1001, np.int64)
    flat[1] = 10
    assert oracle.analyze_count(flat)[0] == -1     # low coverage: no rise after the first minimum


def test_smer_conflict_is_where_the_reference_gives_up(oracle, ref):
    """identical k-mers with different s-mer codes: the reference prints four lines from process_kmer_cluster and exits
    (syncmer.c:1370-1376); the oracle flags the same input (and nothing else in the suites does)"""
    bases, off = pack_reads([parity.CONFLICT_READ])
    db, f = oracle.extract(bases, off, parity.CONFLICT_K, parity.CONFLICT_S)
    c = oracle.collect(db, 1)
    assert f["n_scm"].tolist() == [2] and c["smer_conflict"] == 1 and len(c["h"]) == 1
    rc, lines, out = parity.reference_on_conflict()
    assert rc == 1 and out == ""
    assert lines[0] == "[E::process_kmer_cluster] identical kmers have different smers"
    assert lines[1] == "[E::process_kmer_cluster] kmer hash  : %d" % int(c["h"][0])
    assert lines[2] == "[E::process_kmer_cluster] smer code 0: %d; read id: 0" % int(c["s"][0])
    assert lines[3] == "[E::process_kmer_cluster] smer code 1: %d; read id: 0" % (int(c["s"][0]) ^ 1)
    oracle.free(db, c)


def test_peak_finder_three_ways(oracle, ref):
    """the peak finder behind sr_db_stat's peak_hom / peak_het (and behind the automatic -c): the reference's
    ha_analyze_count (syncmer.c:775-865), the oracle's restatement and the host layer's find_peaks on 3000 synthetic
    multiplicity histograms (error spike, up to three coverage peaks, Poisson noise, truncated / flat / tiny tables)"""
    import ctypes as C
    from oatk_b200.host import build_host
    host = C.CDLL(build_host.build())
    host.oatk_find_peaks.restype = C.c_int
    host.oatk_find_peaks.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    x = np.arange(1001, dtype=np.float64)
    for it in range(3000):
        cnt = rng.choice([0, 1e3, 1e5, 1e7]) * np.exp(-x / rng.uniform(0.5, 3))
        for _ in range(int(rng.integers(0, 4))):
            mu, sd, amp = rng.uniform(6, 400), rng.uniform(1, 40), rng.choice([10, 1e3, 1e5])
            cnt += amp * np.exp(-0.5 * ((x - mu) / sd) ** 2)
        if rng.random() < 0.5: cnt = rng.poisson(np.maximum(cnt, 0)).astype(np.float64)
        if rng.random() < 0.2: cnt[int(rng.integers(0, 1001)):] = 0
        if rng.random() < 0.1: cnt[:] = rng.integers(0, 3, 1001)
        if rng.random() < 0.3: cnt[1000] = rng.choice([0, 5, 1e6])
        cnt[0] = 0
        c = np.ascontiguousarray(cnt.astype(np.int64))
        het = C.c_int(0)
        ours = (host.oatk_find_peaks(len(c), 5, c.ctypes.data, C.byref(het)), het.value)
        assert ours == ref.analyze_count(c) == oracle.analyze_count(c), (it, c[:40].tolist())
