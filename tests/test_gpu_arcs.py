"""a7 on the GPU (sg_arcs: warp-cooperative arc table) against the oracle and the reference goldens."""
import numpy as np
import pytest
from oatk_b200 import synth
from pyoracle import pack_reads
import golden_util as gu

pytestmark = pytest.mark.gpu


def gpu_arcs(gpu_ctx, bases, off, k, s, mkc, a):
    from oatk_b200 import lib
    b = lib.Batch(gpu_ctx)
    b.set_reads_host(bases, off)
    b.extract(k, s)
    b.count()
    out = b.arcs(mkc, a)
    b.close()
    return out


@pytest.mark.parametrize("name", sorted(gu.CASES))
def test_arcs_golden(gpu_ctx, name):
    gen, args, k, s, mkc = gu.CASES[name]
    bases, off = pack_reads(gu.make_reads(gen, args))
    got = gpu_arcs(gpu_ctx, bases, off, k, s, mkc, 0.35)
    assert np.array_equal(got, gu.golden_arcs(gu.load(name)))


@pytest.mark.parametrize("mkc,a", [(0, 0.0), (3, 0.35), (30, 0.35), (5, 1.0), (2, 0.5)])
def test_arcs_oracle(gpu_ctx, oracle, mkc, a):
    """a 40x genome so that the -c / -a filters really cut; includes the EC graph setting (0, 0.)"""
    reads = synth.hifi_reads(21, 120000, 320, 15000, 0.001)
    bases, off = pack_reads(reads)
    db, _ = oracle.extract(bases, off, 1001, 31)
    scm = oracle.collect(db, len(reads))
    exp = oracle.arcs(db, scm, mkc, a)
    got = gpu_arcs(gpu_ctx, bases, off, 1001, 31, mkc, a)
    assert got.shape == exp.shape and np.array_equal(got, exp)
    oracle.free(db, scm)


def test_arcs_tandem_and_palindromes(gpu_ctx, oracle):
    """self-complementary arcs (v+ -> v-) and repeated neighbours from tandem repeats / hairpins"""
    for k, s in ((101, 11), (64, 31)):
        reads = synth.adversarial_reads(3, k, s, scale=2) * 3
        bases, off = pack_reads(reads)
        db, _ = oracle.extract(bases, off, k, s)
        scm = oracle.collect(db, len(reads))
        exp = oracle.arcs(db, scm, 2, 0.35)
        got = gpu_arcs(gpu_ctx, bases, off, k, s, 2, 0.35)
        assert np.array_equal(got, exp)
        oracle.free(db, scm)
