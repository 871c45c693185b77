"""BASELINE.json configs[0] against the answers the unmodified reference gave when SURVEY.md was written (appendix A.3):
reads10k.fa through the CUDA extraction (the per-read dump's md5) and through the `syncasm` command (-k 1001 -s 31 -c 30 -t 8:
both GFA files' md5, the sr_db_stat lines before and after read error correction, the final graph's size)."""
import hashlib
import os
import subprocess
import tempfile
import numpy as np
import pytest
from pyoracle import pack_reads
import survey_reads as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def reads10k():
    K = S.READS10K
    reads, fa = S.generate(K["seed"], K["G"], K["N"], K["L"], K["err"])
    if hashlib.md5(fa).hexdigest() != K["fasta_md5"]:
        pytest.skip("numpy no longer reproduces the survey's FASTA (generator stream changed)")
    return reads, fa


def test_extraction_dump(gpu_ctx, reads10k):
    from oatk_b200 import lib
    K = S.READS10K
    reads, _ = reads10k
    bases, off = pack_reads(reads)
    b = lib.Batch(gpu_ctx)
    b.set_reads_host(bases, off)
    b.extract(1001, 31)
    f = b.extract_download(want_seq=False)
    assert int(f["hoco_l"].astype(np.int64).sum()) == K["hoco_total"] and len(f["m_pos"]) == K["syncmers"]
    assert S.dump_md5(f["hoco_l"], f["n_scm"], f["m_pos"], f["s_mer"], f["k_mer"]) == K["dump_md5"]
    b.count()
    assert b.count_sizes().n_unique == K["distinct_kmers"]
    b.close()


def test_syncasm_command_configs0(reads10k):
    from oatk_b200.host import build_host
    build_host.build()
    K = S.READS10K
    exe = os.path.join(ROOT, "oatk_b200", "host", "syncasm")
    d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        fa = os.path.join(d, "reads10k.fa")
        open(fa, "wb").write(reads10k[1])
        p = subprocess.run([exe, "-k", "1001", "-s", "31", "-c", "30", "-t", "8", "-o", os.path.join(d, "out"), fa],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
        err = p.stderr.decode()
        assert p.returncode == 0, err[-2000:]
        assert hashlib.md5(open(os.path.join(d, "out.utg.gfa"), "rb").read()).hexdigest() == K["utg_gfa_md5"]
        assert hashlib.md5(open(os.path.join(d, "out.utg.final.gfa"), "rb").read()).hexdigest() == K["final_gfa_md5"]
        blocks = err.split("[M::read_error_correction] Error Correction Summary Results")
        assert len(blocks) == 2
        for want in K["stat1"]:
            assert want in blocks[0], want
        for want in K["after_ec"]:
            assert want in blocks[1], want
        tail = err.split("syncmer graph stats after final processing")[1]
        for want in K["final"]:
            assert want in tail, want
    finally:
        import shutil
        shutil.rmtree(d, ignore_errors=True)


def test_reads80k(gpu_ctx):
    """the survey's second set (80 k reads, 1.2 Gbases): the extraction dump and both GFA files of the whole command"""
    from oatk_b200 import lib
    from oatk_b200.host import build_host
    build_host.build()
    K = S.READS80K
    reads, fa = S.generate(K["seed"], K["G"], K["N"], K["L"], K["err"])
    if len(fa) != K["fasta_bytes"] or hashlib.md5(fa).hexdigest() != K["fasta_md5"]:
        pytest.skip("numpy no longer reproduces the survey's FASTA (generator stream changed)")
    bases, off = pack_reads(reads)
    del reads
    b = lib.Batch(gpu_ctx)
    b.set_reads_host(bases, off)
    b.extract(1001, 31)
    f = b.extract_download(want_seq=False)
    assert int(f["hoco_l"].astype(np.int64).sum()) == K["hoco_total"] and len(f["m_pos"]) == K["syncmers"]
    assert S.dump_md5(f["hoco_l"], f["n_scm"], f["m_pos"], f["s_mer"], f["k_mer"]) == K["dump_md5"]
    b.count()
    assert b.count_sizes().n_unique == K["distinct_kmers"]
    b.close()
    exe = os.path.join(ROOT, "oatk_b200", "host", "syncasm")
    d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        path = os.path.join(d, "reads80k.fa")
        open(path, "wb").write(fa)
        del fa
        p = subprocess.run([exe, "-k", "1001", "-s", "31", "-c", "30", "-t", "8", "-o", os.path.join(d, "out"), path],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
        assert p.returncode == 0, p.stderr.decode()[-2000:]
        assert hashlib.md5(open(os.path.join(d, "out.utg.gfa"), "rb").read()).hexdigest() == K["utg_gfa_md5"]
        assert hashlib.md5(open(os.path.join(d, "out.utg.final.gfa"), "rb").read()).hexdigest() == K["final_gfa_md5"]
    finally:
        import shutil
        shutil.rmtree(d, ignore_errors=True)
