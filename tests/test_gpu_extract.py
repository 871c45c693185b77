"""a2-a4 on the GPU (sg_extract through the C ABI) against the CPU oracle, bit for bit."""
import numpy as np
import pytest
from oatk_b200 import synth
from pyoracle import pack_reads
import parity

pytestmark = pytest.mark.gpu

KS = [(1001, 31), (501, 31), (2001, 31), (101, 11), (301, 15), (64, 31), (33, 31), (40, 1), (12, 11), (5, 3)]


def run_gpu(gpu_ctx, bases, off, k, s):
    from oatk_b200 import lib
    b = lib.Batch(gpu_ctx)
    b.set_reads_host(bases, off)
    b.extract(k, s)
    f = b.extract_download()
    b.close()
    return f


def check(gpu_ctx, oracle, reads, k, s):
    bases, off = pack_reads(reads)
    db, exp = oracle.extract(bases, off, k, s)
    oracle.free(db)
    got = run_gpu(gpu_ctx, bases, off, k, s)
    d = parity.diff(got, exp, parity.EXTRACT_FIELDS)
    if d:
        d += parity.per_read_report(got, exp)
    assert not d, "\n".join(d)
    return got


@pytest.mark.parametrize("k,s", KS)
def test_adversarial(gpu_ctx, oracle, k, s):
    check(gpu_ctx, oracle, synth.adversarial_reads(3, k, s), k, s)


@pytest.mark.parametrize("k,s", [(1001, 31), (501, 31), (2001, 31), (301, 15), (40, 1), (12, 11), (20001, 31),
                                 (127, 31), (1007, 31), (47, 31), (142, 31), (1006, 31)])   # q - 1 a multiple of 16, and one off on either side
def test_tandem_repeats(gpu_ctx, oracle, k, s):
    """reads carrying tandem arrays (period 2, 3, TTAGGG, 37, 171; 2 kb up to the whole read): every window position
    ties for the minimum. scan_kernel hands such reads to scan_exact_kernel; both must agree with the reference rules"""
    from oatk_b200 import lib
    L = 15000 if k < 20000 else 60000
    reads = synth.repeat_reads(5, 60, L) + synth.hifi_reads(6, 200000, 20, L, 0.001)
    r = bytearray(synth.repeat_reads(7, 5, L)[4])
    for p in (100, L // 2, L // 2 + 1, L - 50):
        r[p] = ord("N")
    reads.append(bytes(r))                                    # ambiguous bases inside and around an array
    bases, off = pack_reads(reads)
    db, exp = oracle.extract(bases, off, k, s)
    oracle.free(db)
    b = lib.Batch(gpu_ctx)
    b.set_reads_host(bases, off)
    b.extract(k, s)
    got = b.extract_download()
    n_def = b.debug_scan_info()
    b.close()
    d = parity.diff(got, exp, parity.EXTRACT_FIELDS)
    if d:
        d += parity.per_read_report(got, exp)
    assert not d, "\n".join(d)
    if k - s + 1 > 64:
        assert n_def >= 30, "the repeat reads were expected on the exact path, only %d went" % n_def
    assert n_def <= len(reads)


@pytest.mark.parametrize("k,s", [(63, 31), (95, 21), (127, 31), (96, 31), (1023, 31)])
def test_kmer_block_shapes(gpu_ctx, oracle, k, s):
    """k-mer lengths around multiples of 32 bases: a last 8-byte Murmur block that is not full of bases, no tail,
    a tail of every size class (the interior fast path of kmerhash_kernel and the checked one)"""
    reads = synth.hifi_reads(9, 60000, 40, 6000, 0.002) + synth.adversarial_reads(3, k, s)
    check(gpu_ctx, oracle, reads, k, s)


@pytest.mark.parametrize("k,s", [(1001, 31), (501, 31), (2001, 31)])
def test_hifi_small(gpu_ctx, oracle, k, s):
    reads = synth.hifi_reads(42, 500000, 300, 15000, 0.001)
    got = check(gpu_ctx, oracle, reads, k, s)
    if (k, s) == (1001, 31):
        # first read of reads10k.fa as extracted by the reference
        assert got["hoco_l"][0] == 11296 and got["n_scm"][0] == 21
        assert (int(got["m_pos"][0]), int(got["s_mer"][0]), int(got["k_mer"][0])) == \
            (447, 1089221421968769627, 12466468101431059233)


def test_empty_batch(gpu_ctx):
    from oatk_b200 import lib
    b = lib.Batch(gpu_ctx)
    b.set_reads_host(np.zeros(1, np.uint8), np.zeros(1, np.uint64))
    b.extract(1001, 31)
    z = b.extract_sizes()
    assert z.n_reads == 0 and z.n_syncmers == 0


def test_bad_params(gpu_ctx):
    from oatk_b200 import lib
    b = lib.Batch(gpu_ctx)
    b.set_reads_host(np.frombuffer(b"ACGT", np.uint8), np.array([0, 4], np.uint64))
    for k, s in ((31, 31), (10, 32), (10, 0), (5, 9)):
        with pytest.raises(lib.SgError):
            b.extract(k, s)


def test_long_read(gpu_ctx, oracle):
    """one read much longer than a tile: exercises the ring buffer wrap and the tile carries"""
    rng = np.random.default_rng(11)
    reads = [synth._rand(rng, 300000), synth._rand(rng, 70000).replace(b"AC", b"AAAAC")]
    check(gpu_ctx, oracle, reads, 1001, 31)
    check(gpu_ctx, oracle, reads, 2001, 31)


def test_unaligned_offsets(gpu_ctx, oracle):
    """reads whose raw offsets are not 16-byte aligned, of every length mod 16"""
    rng = np.random.default_rng(5)
    reads = [synth._rand(rng, 1200 + i) for i in range(40)]
    check(gpu_ctx, oracle, reads, 301, 15)


def test_large_k_and_the_window_limit(gpu_ctx, oracle):
    """k far above the default still fits the shared-memory window; beyond the limit the call fails loudly"""
    from oatk_b200 import lib
    rng = np.random.default_rng(2)
    reads = [synth._rand(rng, 60000), synth._rand(rng, 9000), synth._rand(rng, 45000)]
    check(gpu_ctx, oracle, reads, 8001, 31)
    check(gpu_ctx, oracle, reads, 20001, 25)
    check(gpu_ctx, oracle, reads, 40001, 31)      # one warp per CTA, ring of 4096 chunks
    b = lib.Batch(gpu_ctx)
    bases, off = pack_reads(reads)
    b.set_reads_host(bases, off)
    with pytest.raises(lib.SgError) as e:
        b.extract(200001, 31)
    assert e.value.code == -5


def test_many_tiny_reads(gpu_ctx, oracle):
    rng = np.random.default_rng(4)
    reads = [synth._rand(rng, int(n)) for n in rng.integers(0, 200, 20000)]
    check(gpu_ctx, oracle, reads, 64, 31)
    check(gpu_ctx, oracle, reads, 1001, 31)
