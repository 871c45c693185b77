"""f1 on the device: sg_runlen_sums (the run-length sums behind scg_syncmer_consensus, reference syncasm.c:946-998) against a
numpy restatement over the downloaded ho_rl / ho_l_rl of the same reads -- plain batches, reads with homopolymers past 255,
both strands, k below and above the 1024 positions a CTA keeps in registers, and the pipeline's master batch with the run
lengths left on the device (sg_pipe_keep_run_lengths / no ho_rl buffer)."""
import ctypes as C
import numpy as np
import pytest
from oatk_b200 import synth
from pyoracle import pack_reads

pytestmark = pytest.mark.gpu


def expected_sums(f, scm, k, want):
    """f: extract_download dict (compact per-read layout), scm: count_download dict; want: syncmer ids"""
    hoco_l = f["hoco_l"].astype(np.int64)
    rl_off = np.concatenate([[0], np.cumsum(hoco_l)])            # extract_download hands the arrays over unpadded
    scm_off = np.concatenate([[0], np.cumsum(f["n_scm"].astype(np.int64))])
    # long runs: per read, in order of the 255 marks
    lrl_off = np.concatenate([[0], np.cumsum(f["n_lrl"].astype(np.int64))])
    out = np.zeros((len(want), k), np.uint64)
    occ_off, occ = [0], []
    for row, u in enumerate(want):
        for o in scm["occ"][scm["off"][u]:scm["off"][u + 1]]:
            sid, idx = int(o >> np.uint64(32)), int(o >> np.uint64(1)) & 0x7FFFFFFF
            mp = int(f["m_pos"][scm_off[sid] + idx])
            start, strand = mp >> 1, mp & 1
            rl = f["ho_rl"][rl_off[sid]:rl_off[sid] + hoco_l[sid]].astype(np.uint64)
            marks = np.nonzero(rl == 255)[0]
            if len(marks):
                rl[marks] = f["ho_l_rl"][lrl_off[sid]:lrl_off[sid] + len(marks)]
            seg = rl[start:start + k]
            out[row] += seg[::-1] if strand else seg
            occ.append((sid << 32) | mp)
        occ_off.append(len(occ))
    return out, np.array(occ_off, np.uint64), np.array(occ, np.uint64)


@pytest.mark.parametrize("k,s", [(301, 15), (1001, 31), (1501, 31)])
def test_runlen_sums_match_host_arithmetic(gpu_ctx, k, s):
    from oatk_b200 import lib
    rng = np.random.default_rng(k)
    L = 6 * k
    genome = synth._rand(rng, 3 * L)
    genome = genome[:L] + b"A" * 300 + genome[L:2 * L] + b"C" * 1000 + genome[2 * L:]          # runs past 255 inside k-mers
    reads = []
    for i in range(60):
        a = int(rng.integers(0, len(genome) - L))
        r = genome[a:a + L]
        reads.append(synth.revcomp(r) if i % 2 else r)
    bases, off = pack_reads(reads)
    b = lib.Batch(gpu_ctx)
    b.set_reads_host(bases, off)
    b.extract(k, s)
    f = b.extract_download()
    b.count()
    scm = b.count_download()
    assert f["n_lrl"].sum() > 0
    want = np.argsort(-scm["cov"].astype(np.int64), kind="stable")[:80]                            # the deepest syncmers
    exp, occ_off, occ = expected_sums(f, scm, k, want)
    got = b.runlen_sums(occ_off, occ, k)
    assert np.array_equal(got, exp)
    # an occurrence outside its read is an error, not a silent zero
    bad = occ.copy()
    bad[0] = (int(bad[0]) & ~0xFFFFFFFF) | ((int(f["hoco_l"][int(bad[0]) >> 32]) - 3) << 1)
    with pytest.raises(lib.SgError):
        b.runlen_sums(occ_off, bad, k)
    b.close()


def test_pipe_master_keeps_run_lengths(gpu_ctx):
    """the host-buffer pipeline without a ho_rl buffer: nothing of it is downloaded, the master batch serves the sums"""
    from oatk_b200 import lib
    k, s = 501, 31
    rng = np.random.default_rng(5)
    reads = synth.hifi_reads(77, 40000, 700, 9000, 0.002)
    reads[3] = reads[3][:2000] + b"G" * 400 + reads[3][2000:]
    reads[500] = reads[500][:100] + b"T" * 260 + reads[500][100:]
    bases, off = pack_reads(reads)
    ref = lib.Batch(gpu_ctx)
    ref.set_reads_host(bases, off)
    ref.extract(k, s)
    f = ref.extract_download()
    ref.count()
    scm = ref.count_download()
    want = np.argsort(-scm["cov"].astype(np.int64), kind="stable")[:200]
    exp, occ_off, occ = expected_sums(f, scm, k, want)

    n, N = len(reads), int(f["n_scm"].sum())
    o = lib.ExtractOut()
    keep = {}
    def arr(name, size, dt):
        keep[name] = np.zeros(max(int(size), 1), dt)
        setattr(o, name, keep[name].ctypes.data)
    arr("hoco_l", n, np.uint32); arr("n_scm", n, np.uint32)
    arr("hoco_s_off", n + 1, np.uint64); arr("ho_rl_off", n + 1, np.uint64); arr("scm_off", n + 1, np.uint64)
    arr("hoco_s_buf", len(f["hoco_s"]) + 64 * n + 4096, np.uint8)
    o.ho_rl_buf = None                                         # <- run lengths stay on the device
    arr("m_pos", N + 16, np.uint32); arr("s_mer", N + 16, np.uint64); arr("k_mer", N + 16, np.uint64)
    for nm in ("amb_sid", "amb_pos", "lrl_sid", "lrl_idx", "lrl_val"):
        arr(nm, 1024, np.uint32)
    caps = lib.PipeCaps(N + 16, len(keep["hoco_s_buf"]), 0, 1024, 1024)
    pipe = lib.Pipe(gpu_ctx.device, 3)
    z = pipe.run_host(bases.ctypes.data, off.ctypes.data, n, k, s, 128, o, caps)
    assert z.n_syncmers == N
    assert np.array_equal(keep["m_pos"][:N], f["m_pos"])
    got = pipe.master.runlen_sums(occ_off, occ, k)
    assert np.array_equal(got, exp)
    pipe.close()
    ref.close()
