"""CPU, world_size 2, gloo: the host logic of the multi-GPU exchange (split sizes, range partition,
global id bases, id return) with the oracle standing in for the per-rank extraction and counting."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORLD = 2
K, S = 301, 15


def _worker(rank, port, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oatk_b200 import synth, dist as sgdist
    from pyoracle import Oracle, pack_reads
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        reads = synth.hifi_reads(17, 60000, 80, 9000, 0.001)
        per = len(reads) // WORLD
        mine = reads[rank * per:(rank + 1) * per]
        O = Oracle()
        bases, off = pack_reads(mine)
        db, f = O.extract(bases, off, K, S)
        # tuples of the local reads, sid = global read index
        sid = np.repeat(np.arange(per, dtype=np.uint64) + np.uint64(rank * per), f["n_scm"])
        idx = np.concatenate([np.arange(n, dtype=np.uint64) for n in f["n_scm"]]) if len(sid) else np.zeros(0, np.uint64)
        occ = sid << np.uint64(32) | idx << np.uint64(1) | (f["m_pos"].astype(np.uint64) & np.uint64(1))
        keys = torch.from_numpy(f["k_mer"].view(np.int64).copy())
        part = sgdist.range_part(keys, WORLD).numpy()
        order = np.argsort(part, kind="stable")
        rows = np.stack([f["k_mer"][order], occ[order], f["s_mer"][order]], axis=1).astype(np.uint64)
        counts = np.bincount(part, minlength=WORLD).tolist()
        dev = torch.device("cpu")
        recv = sgdist.exchange_counts(dist, counts, dev)
        got = sgdist.exchange_rows(dist, torch.from_numpy(rows.view(np.int64).reshape(-1).copy()), counts, recv, 3)
        got = got.numpy().view(np.uint64).reshape(-1, 3)
        # local count of the owned hash range (stable sort by hash keeps (sid, idx) order)
        o2 = np.argsort(got[:, 0], kind="stable")
        sk, so = got[o2, 0], got[o2, 1]
        head = np.ones(len(sk), bool)
        head[1:] = sk[1:] != sk[:-1]
        local_id = np.cumsum(head) - 1
        base, allc = sgdist.id_base(dist, int(head.sum()), rank, WORLD, dev)
        # ids back to the owners of the reads: adopted order, same splits reversed
        kid = np.empty(len(sk), np.uint64)
        kid[o2] = (local_id + base).astype(np.uint64) << np.uint64(1)
        pairs = np.stack([got[:, 1], kid], axis=1).astype(np.uint64)
        back = sgdist.exchange_rows(dist, torch.from_numpy(pairs.view(np.int64).reshape(-1).copy()), recv, counts, 2)
        back = back.numpy().view(np.uint64).reshape(-1, 2)
        mykid = np.empty(len(occ), np.uint64)
        pos = {int(o): i for i, o in enumerate(occ)}
        for o, v in back:
            mykid[pos[int(o)]] = v
        ret[rank] = dict(h=sk[head], cov=np.diff(np.concatenate([np.nonzero(head)[0], [len(sk)]])), occ=so, kid=mykid,
                         base=base, counts=allc)
        O.free(db)
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_matches_single_process():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oatk_b200 import synth
    from pyoracle import Oracle, pack_reads
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(port, ret), nprocs=WORLD, join=True)
    reads = synth.hifi_reads(17, 60000, 80, 9000, 0.001)
    O = Oracle()
    bases, off = pack_reads(reads)
    db, f = O.extract(bases, off, K, S)
    exp = O.collect(db, len(reads))
    r0, r1 = ret[0], ret[1]
    assert np.array_equal(np.concatenate([r0["h"], r1["h"]]), exp["h"])
    assert np.array_equal(np.concatenate([r0["cov"], r1["cov"]]).astype(np.uint32), exp["cov"])
    assert np.array_equal(np.concatenate([r0["occ"], r1["occ"]]), exp["occ"])
    assert np.array_equal(np.concatenate([r0["kid"], r1["kid"]]), exp["k_mer_id"])
    assert r0["base"] == 0 and r1["base"] == len(r0["h"]) and sum(r0["counts"]) == len(exp["h"])
    O.free(db, exp)


def test_range_part_is_monotone_and_balanced():
    sys.path.insert(0, ROOT)
    from oatk_b200 import dist as sgdist
    rng = np.random.default_rng(1)
    k = np.sort(rng.integers(0, 2**63, 20000, dtype=np.int64).astype(np.uint64) * np.uint64(2) + rng.integers(0, 2, 20000).astype(np.uint64))
    for world in (1, 2, 3, 8):
        p = sgdist.range_part(torch.from_numpy(k.view(np.int64).copy()), world).numpy()
        assert p.min() >= 0 and p.max() == world - 1 or world == 1
        assert np.all(np.diff(p) >= 0)
        assert np.bincount(p, minlength=world).min() > 20000 / world * 0.8
        # the two hashes sr_db_stat merges (h and h^1) never straddle a boundary
        p2 = sgdist.range_part(torch.from_numpy((k ^ np.uint64(1)).view(np.int64).copy()), world).numpy()
        assert np.array_equal(p, p2)
