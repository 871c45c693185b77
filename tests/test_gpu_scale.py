"""BASELINE-size runs (configs[1]: 1 M x 15 kb, generated on the device) checked through properties that
do not need the oracle to process the whole input, plus exact oracle parity on a random sample of reads."""
import numpy as np
import pytest
import torch
from pyoracle import pack_reads

pytestmark = pytest.mark.gpu
K, S, L = 1001, 31, 15000


@pytest.fixture(scope="module")
def big(gpu_ctx):
    from oatk_b200 import lib, synth_gpu
    free, _ = torch.cuda.mem_get_info()
    n = 1_000_000 if free > 100e9 else 200_000
    dev = torch.device("cuda", 0)
    bases, off = synth_gpu.hifi_reads_gpu(1000, 50_000_000, n, L, 1e-3, dev, genome_seed=1)
    b = lib.Batch(gpu_ctx)
    b.set_reads_device(bases.data_ptr(), off.data_ptr(), n, n * L)
    b.extract(K, S)
    kp, kn = b.buffer("key")
    from oatk_b200 import dist as sgdist
    keys = sgdist.tensor_from_ptr(kp, kn, dev).clone().cpu().numpy().view(np.uint64)
    st = b.stat()
    b.count()
    scm = b.count_download()
    f = b.extract_download(want_seq=False)          # k_mer[] now holds id << 1
    return dict(n=n, bases=bases, f=f, keys=keys, st=st, scm=scm, batch=b)


def test_counts_add_up(big):
    f, scm, st = big["f"], big["scm"], big["st"]
    N = len(f["m_pos"])
    assert N == int(f["n_scm"].sum()) == st.n_syncmers
    assert int(scm["cov"].astype(np.int64).sum()) == N                 # the reference's own assert, syncmer.c:1448
    assert int(scm["off"][-1]) == N and np.array_equal(np.diff(scm["off"].astype(np.int64)), scm["cov"])
    assert sum(st.kmer_cnts[:]) == st.kmer_unique and sum(st.smer_cnts[:]) == st.smer_unique
    assert sum(i * c for i, c in enumerate(st.kmer_cnts[:1000])) + 0 <= N
    # about 2 syncmers per window of q hoco positions (21.4 per 15 kb read at k=1001)
    assert 19.0 < N / big["n"] < 24.0


def test_database_is_sorted_and_consistent(big):
    f, scm, keys = big["f"], big["scm"], big["keys"]
    h = scm["h"]
    assert np.all(h[1:] > h[:-1])                                      # ids are ranks in hash order, no hash group was split
    assert scm["n_hash_collisions"] == 0
    ids = (f["k_mer"] >> np.uint64(1)).astype(np.int64)                 # k_mer[] after collect = id << 1
    assert np.all((f["k_mer"] & np.uint64(1)) == 0)
    assert np.array_equal(h[ids], keys)                                 # every occurrence points at its own hash
    assert np.array_equal(np.bincount(ids, minlength=len(h)), scm["cov"])
    # occurrence lists: (sid, idx) ascending inside every id, and they invert k_mer[]
    occ = scm["occ"]
    same = np.repeat(np.arange(len(h)), scm["cov"])
    inc = occ[1:] > occ[:-1]
    assert np.all(inc | (same[1:] != same[:-1]))
    scm_off = np.concatenate([[0], np.cumsum(f["n_scm"].astype(np.int64))])
    o = scm_off[(occ >> np.uint64(32)).astype(np.int64)] + ((occ >> np.uint64(1)) & np.uint64(0x7FFFFFFF)).astype(np.int64)
    assert np.array_equal(ids[o], same)
    assert np.array_equal((occ & np.uint64(1)).astype(np.uint32), f["m_pos"][o] & 1)


def test_sampled_reads_match_the_oracle(big, oracle):
    rng = np.random.default_rng(3)
    n, f = big["n"], big["f"]
    pick = np.sort(rng.choice(n, 300, replace=False))
    rows = big["bases"].view(n, L)[torch.from_numpy(pick).to(big["bases"].device)].cpu().numpy()
    bases, off = pack_reads([r.tobytes() for r in rows])
    db, exp = oracle.extract(bases, off, K, S)
    oracle.free(db)
    scm_off = np.concatenate([[0], np.cumsum(f["n_scm"].astype(np.int64))])
    sel = np.concatenate([np.arange(scm_off[r], scm_off[r + 1]) for r in pick])
    assert np.array_equal(f["hoco_l"][pick], exp["hoco_l"])
    assert np.array_equal(f["n_scm"][pick], exp["n_scm"])
    assert np.array_equal(f["m_pos"][sel], exp["m_pos"])
    assert np.array_equal(f["s_mer"][sel], exp["s_mer"])
    assert np.array_equal(big["keys"][sel], exp["k_mer"])


def test_rerun_is_bit_identical(big, gpu_ctx):
    """atomics only allocate slots; nothing observable may depend on scheduling"""
    from oatk_b200 import lib
    n = min(big["n"], 200_000)
    b = lib.Batch(gpu_ctx)
    off = (torch.arange(0, n + 1, dtype=torch.int64, device=big["bases"].device) * L)
    outs = []
    for _ in range(2):
        b.set_reads_device(big["bases"].data_ptr(), off.data_ptr(), n, n * L)
        b.extract(K, S)
        b.count()
        f = b.extract_download(want_seq=False)
        s = b.count_download()
        outs.append((f["m_pos"].copy(), f["s_mer"].copy(), f["k_mer"].copy(), s["h"].copy(), s["occ"].copy()))
    for x, y in zip(*outs):
        assert np.array_equal(x, y)
    b.close()
