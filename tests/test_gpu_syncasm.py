"""The whole command on the GPU: syncasm() of the host layer (oatk_b200/host/run_syncasm_gpu.c -- FASTA in,
<out>.utg.gfa and <out>.utg.final.gfa out; extraction, counting, statistics and arc tallies on the device, the rest on
the host) against the golden outputs of the unmodified reference's syncasm() in tests/golden/syncasm.json (made by
tests/golden/make_golden_syncasm.py). Where oracle/_ref/libref.so is present the reference command is also run on the
spot and the files are compared byte by byte."""
import ctypes as C
import hashlib
import json
import os
import sys
import tempfile
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_syncasm as G   # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "syncasm.json")))


@pytest.fixture(scope="module")
def host():
    from oatk_b200.host import build_host
    L = C.CDLL(build_host.build())
    G.bind(L)
    return L


@pytest.mark.parametrize("case", sorted(G.CASES))
def test_both_gfa_files_match_the_reference(host, case):
    tmp = tempfile.mkdtemp()
    fa, prefix = os.path.join(tmp, "reads.fa"), os.path.join(tmp, "ours")
    G.write_fasta(case, fa)
    assert hashlib.md5(open(fa, "rb").read()).hexdigest() == GOLD[case]["fasta_md5"], "generator drifted: regenerate the golden file"
    assert G.run(host, fa, G.args_of(case), prefix, 4) == 0
    for suffix in (".utg.gfa", ".utg.final.gfa"):
        assert G.summary(prefix + suffix) == GOLD[case][suffix], (case, suffix)
    from pyoracle import have_ref
    if have_ref():
        from pyoracle import Ref
        R = Ref().L
        G.bind(R)
        ref_prefix = os.path.join(tmp, "ref")
        assert G.run(R, fa, G.args_of(case), ref_prefix, 4) == 0
        for suffix in (".utg.gfa", ".utg.final.gfa"):
            assert open(prefix + suffix, "rb").read() == open(ref_prefix + suffix, "rb").read(), (case, suffix)
            os.unlink(ref_prefix + suffix)
    for suffix in (".utg.gfa", ".utg.final.gfa"):
        os.unlink(prefix + suffix)
    os.unlink(fa)
    os.rmdir(tmp)


def test_errors_come_back(host):
    """a missing file and a read set without syncmers: 1, no exit(), nothing written"""
    tmp = tempfile.mkdtemp()
    prefix = os.path.join(tmp, "x")
    a = dict(k=201, s=15, mkc=3, af=0.35, ec=1, unzip=3, bubble=100000, tip=10000, weak=0.3)
    assert G.run(host, os.path.join(tmp, "missing.fa"), a, prefix, 2) == 1
    fa = os.path.join(tmp, "short.fa")
    with open(fa, "wb") as f:
        f.write(b">a\nACGTACGTACGT\n>b\nTTTTGGGGCCCCAAAA\n")
    assert G.run(host, fa, a, prefix, 2) == 1
    assert not os.path.exists(prefix + ".utg.final.gfa")
    os.unlink(fa)
    for f in os.listdir(tmp):
        os.unlink(os.path.join(tmp, f))
    os.rmdir(tmp)
