"""a5/a6 on the GPU (sg_stat, sg_count through the C ABI) against the CPU oracle, bit for bit."""
import numpy as np
import pytest
from oatk_b200 import synth
from pyoracle import pack_reads
import parity

pytestmark = pytest.mark.gpu


def gpu_batch(gpu_ctx, bases, off, k, s):
    from oatk_b200 import lib
    b = lib.Batch(gpu_ctx)
    b.set_reads_host(bases, off)
    b.extract(k, s)
    return b


def check_count(gpu_ctx, oracle, reads, k, s, hash_bits=64):
    bases, off = pack_reads(reads)
    db, _ = oracle.extract(bases, off, k, s)
    exp = oracle.collect(db, len(reads), hash_bits)
    b = gpu_batch(gpu_ctx, bases, off, k, s)
    if hash_bits != 64:
        b.debug_set_hash_bits(hash_bits)
    # both group tests: packed k-mers compared for every tuple (the reference's way), then the default
    # hash + fingerprint test with exact comparison only inside groups whose fingerprints differ
    b.set_exact_verify(True)
    b.count()
    exact = b.count_download()
    assert not parity.diff(exact, exp, parity.SCM_FIELDS)
    b.set_exact_verify(False)
    b.count()
    got = b.count_download()
    assert got["n_hash_collisions"] == exact["n_hash_collisions"]
    d = parity.diff(got, exp, parity.SCM_FIELDS)
    # k_mer[] as the host sees it after collect
    f = b.extract_download(want_seq=False)
    d += parity.diff({"k_mer_id": f["k_mer"]}, exp, ("k_mer_id",), tag="download:")
    assert not d, "\n".join(d)
    oracle.free(db, exp)
    b.close()
    return got


@pytest.mark.parametrize("k,s", [(1001, 31), (501, 31), (101, 11), (64, 31), (33, 31)])
def test_count_adversarial(gpu_ctx, oracle, k, s):
    check_count(gpu_ctx, oracle, synth.adversarial_reads(3, k, s), k, s)


def test_count_hifi(gpu_ctx, oracle):
    reads = synth.hifi_reads(7, 200000, 400, 15000, 0.001)
    got = check_count(gpu_ctx, oracle, reads, 1001, 31)
    assert got["n_hash_collisions"] == 0
    assert int(got["cov"].sum()) == len(got["occ"])


@pytest.mark.parametrize("bits", [3, 8, 16])
def test_forced_hash_collisions(gpu_ctx, oracle, bits):
    """truncating the hash forces distinct k-mers into one hash group: the exact-sequence
    split must number the classes like process_kmer_cluster (syncmer.c:1270-1393)"""
    reads = synth.hifi_reads(9, 60000, 60, 8000, 0.002) + synth.adversarial_reads(4, 301, 15)[:20]
    got = check_count(gpu_ctx, oracle, reads, 301, 15, hash_bits=bits)
    assert got["n_hash_collisions"] > 0


@pytest.mark.parametrize("k,s", [(1001, 31), (301, 15)])
def test_stat(gpu_ctx, oracle, k, s):
    reads = synth.hifi_reads(5, 100000, 300, 12000, 0.001) + synth.adversarial_reads(3, k, s)
    bases, off = pack_reads(reads)
    db, _ = oracle.extract(bases, off, k, s)
    rc, d, i, sc, kc = oracle.stat(db)
    assert rc == 0
    b = gpu_batch(gpu_ctx, bases, off, k, s)
    st = b.stat()
    assert np.array_equal(np.array(st.smer_cnts[:], np.int64), sc)
    assert np.array_equal(np.array(st.kmer_cnts[:], np.int64), kc)
    assert (st.smer_unique, st.smer_singleton, st.kmer_unique, st.kmer_singleton) == (i[0], i[1], i[4], i[5])
    assert st.gap_sum / st.n_gaps == d[1]
    assert st.n_syncmers / len(reads) == d[0]
    oracle.free(db)
    b.close()


def test_stat_then_count_order(gpu_ctx, oracle):
    """the reference calls sr_db_stat before collect (run_syncasm.c:88-103); both orders must agree"""
    reads = synth.hifi_reads(3, 50000, 100, 9000, 0.001)
    bases, off = pack_reads(reads)
    b = gpu_batch(gpu_ctx, bases, off, 501, 31)
    s1 = b.stat()
    b.count()
    c1 = b.count_download()
    b2 = gpu_batch(gpu_ctx, bases, off, 501, 31)
    b2.count()
    c2 = b2.count_download()
    s2 = b2.stat()
    assert not parity.diff(c1, c2, parity.SCM_FIELDS)
    assert s1.kmer_cnts[:] == s2.kmer_cnts[:] and s1.smer_cnts[:] == s2.smer_cnts[:]


def test_count_empty(gpu_ctx):
    from oatk_b200 import lib
    b = lib.Batch(gpu_ctx)
    b.set_reads_host(np.frombuffer(b"ACGTACGT", np.uint8), np.array([0, 8], np.uint64))
    b.extract(1001, 31)
    with pytest.raises(lib.SgError) as e:
        b.count()
    assert e.value.code == -7


@pytest.mark.parametrize("low_bits,expect", [(24, "repair-or-clean"), (48, "repair"), (56, "fallback"), (0, "full")])
def test_partial_sort_and_repair(gpu_ctx, oracle, low_bits, expect):
    """sg_count sorts on the top 64 - low_bits hash bits and repairs the runs that differ below; moving the split
    makes the repair pass (and, past its buffers, the fall-back to the full sort) do real work"""
    from oatk_b200 import lib
    # one owner per run now repairs it whole, however many inversions it holds: only runs past 4096 tuples (here: 256
    # buckets of ~4600) send the sort back to all 64 bits
    n_reads, read_len = (6000, 12000) if expect == "fallback" else (300, 15000)
    reads = synth.hifi_reads(11, 100000, n_reads, read_len, 0.001) + synth.adversarial_reads(5, 101, 11)
    bases, off = pack_reads(reads)
    db, _ = oracle.extract(bases, off, 101, 11)
    exp = oracle.collect(db, len(reads), 64)
    b = gpu_batch(gpu_ctx, bases, off, 101, 11)
    b.debug_set_sort_low_bits(low_bits)
    b.count()
    got = b.count_download()
    repairs, fell_back = b.debug_sort_info()
    d = parity.diff(got, exp, parity.SCM_FIELDS)
    assert not d, "\n".join(d)
    if expect == "repair":
        assert repairs > 0 and not fell_back
    elif expect == "fallback":
        assert fell_back
    oracle.free(db, exp)
    b.close()


@pytest.mark.parametrize("bits,expect,deep", [(32, "none", False), (16, "repair", False), (8, "repair", False), (16, "repair", True), (8, "repair", True)])
def test_packed_sort_and_repair(gpu_ctx, oracle, bits, expect, deep):
    """the default sort orders (top hash bits | tuple index) words and repairs the runs that differ below; with 16 or 8
    bits in the word nearly every tuple sits in such a run, so the five-array repair does real work. The order, ids and
    occurrence lists must not depend on where the split is."""
    if deep:   # few distinct k-mers, each several hundred copies deep (what one GPU of eight adopts): long runs, few hashes per run
        reads = synth.hifi_reads(13, 12000, 1500, 4000, 0.0002)
    else:
        reads = synth.hifi_reads(12, 100000, 300, 15000, 0.001) + synth.adversarial_reads(5, 101, 11)
    bases, off = pack_reads(reads)
    db, _ = oracle.extract(bases, off, 101, 11)
    exp = oracle.collect(db, len(reads), 64)
    b = gpu_batch(gpu_ctx, bases, off, 101, 11)
    b.debug_set_pack_bits(bits)
    st = b.stat()
    b.count()
    got = b.count_download()
    repairs, fell_back = b.debug_sort_info()
    d = parity.diff(got, exp, parity.SCM_FIELDS)
    assert not d, "\n".join(d)
    rc, dd, ii, sc, kc = oracle.stat(db)
    assert np.array_equal(np.array(st.kmer_cnts[:]), kc) and np.array_equal(np.array(st.smer_cnts[:]), sc)
    assert not fell_back
    if expect == "repair":
        assert repairs > 0
    oracle.free(db, exp)
    b.close()


def test_host_binding_reports_and_never_fails(gpu_ctx):
    """sg_host_bind_near_device with no flags only reports (NUMA node or -1, 0 CPUs bound); with the memory flag it may set the
    page policy of this thread, which no test depends on. A box that hides its topology is not an error."""
    from oatk_b200 import lib
    node, ncpu = lib.bind_host_near_device(0, cpus=False, memory=False)
    assert isinstance(node, int) and ncpu == 0
    node2, ncpu2 = lib.bind_host_near_device(0, cpus=False, memory=True)
    assert ncpu2 == 0 and (node2 == node or node2 < -1)


def test_smer_conflict_is_reported_like_the_reference(gpu_ctx, oracle, ref):
    """the one input class on which the reference exits (identical k-mers, different s-mer codes; tests/parity.py): sg_count
    returns SG_E_SMER_CONFLICT and sg_count_conflict the figures of the reference's four lines; with other reads around it,
    the FIRST conflicting class in hash order is the one reported"""
    from oatk_b200 import lib
    k, s = parity.CONFLICT_K, parity.CONFLICT_S
    other = synth.hifi_reads(21, 30000, 40, 3000, 0.001)
    for reads, rid in (([parity.CONFLICT_READ], 0), (other[:20] + [parity.CONFLICT_READ] + other[20:], 20)):
        bases, off = pack_reads(reads)
        b = gpu_batch(gpu_ctx, bases, off, k, s)
        with pytest.raises(lib.SgError) as e:
            b.count()
        assert e.value.code == -6
        got = b.count_conflict()
        rc, lines, _ = parity.reference_on_conflict()
        assert rc == 1
        assert lines[1].endswith(": %d" % got[0])
        assert lines[2] == "[E::process_kmer_cluster] smer code 0: %d; read id: 0" % got[1] and got[2] == rid
        assert lines[3] == "[E::process_kmer_cluster] smer code 1: %d; read id: 0" % got[3] and got[4] == rid
        b.close()
    # a batch without a conflict has nothing to report
    bases, off = pack_reads(other)
    b = gpu_batch(gpu_ctx, bases, off, k, s)
    b.count()
    with pytest.raises(lib.SgError):
        b.count_conflict()
    b.close()
