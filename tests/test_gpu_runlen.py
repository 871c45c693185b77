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


@pytest.mark.parametrize("k,s", [(301, 15), (1001, 31)])
def test_kmer_codes_match_the_packed_bases(gpu_ctx, k, s):
    """sg_kmer_codes (the bases behind get_kmer_seq for read databases whose hoco_s stayed on the device) against the
    downloaded packed bases of the same reads: stretches at the start, in the middle and at the very end of reads, and a
    request outside its read must be refused"""
    from oatk_b200 import lib
    rng = np.random.default_rng(k + 1)
    reads = [synth._rand(rng, int(n)) for n in rng.integers(2 * k, 6 * k, 40)]
    bases, off = pack_reads(reads)
    b = lib.Batch(gpu_ctx)
    b.set_reads_host(bases, off)
    b.extract(k, s)
    f = b.extract_download()
    hoco_l = f["hoco_l"].astype(np.int64)
    hs_off = np.concatenate([[0], np.cumsum((hoco_l + 3) // 4)])      # extract_download hands the packed bases over unpadded, read after read
    L = lib.library()
    L.sg_kmer_codes.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]
    refs, exp = [], []
    for sid in range(len(reads)):
        # unpack the read's bases from the compact download (4 per byte, first in the top bits)
        o0 = int(hs_off[sid])
        nb = (hoco_l[sid] + 3) // 4
        packed = f["hoco_s"][o0:o0 + nb]
        codes = np.stack([(packed >> 6) & 3, (packed >> 4) & 3, (packed >> 2) & 3, packed & 3], axis=1).reshape(-1)[:hoco_l[sid]]
        for start in (0, int(hoco_l[sid] - k) // 2, int(hoco_l[sid] - k)):
            refs.append((sid << 32) | start)
            exp.append(codes[start:start + k])
    refs = np.array(refs, np.uint64)
    out = np.zeros((len(refs), k), np.uint8)
    assert L.sg_kmer_codes(b.h, len(refs), refs.ctypes.data, k, out.ctypes.data) == 0
    assert np.array_equal(out, np.stack(exp).astype(np.uint8))
    bad = np.array([(0 << 32) | int(hoco_l[0] - k + 1)], np.uint64)
    assert L.sg_kmer_codes(b.h, 1, bad.ctypes.data, k, out.ctypes.data) != 0
    b.close()
