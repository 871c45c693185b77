"""The multi-GPU parity check, shared by tests/run_multigpu_parity.py (pytest, torchrun) and bench.py (which runs it before
timing whenever WORLD_SIZE > 1, so that the driver's scaling record carries it).

Every rank extracts its contiguous block of a seeded read set (HiFi-like reads + the adversarial set + reads with tandem
arrays) on its own GPU; the tuples are exchanged in C over NCCL (sg_comm_exchange_tuples), every rank counts its hash range,
ids come back (sg_comm_return_ids), the arc tally is exchanged and filtered (sg_comm_arcs). Rank 0 runs the CPU oracle on
the WHOLE read set and compares: syncmer_t.{h, s, cov}, the concatenated occurrence lists, k_mer ids of every read, the
sr_db_stat tables over all reads, and the arc list for two (min_k_cov, a) settings. TEST INFRASTRUCTURE: uses oracle/."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def share_unique_id(dist, rank, dev):
    """rank 0's NCCL unique id to everybody, over the process group the launcher already has"""
    import torch
    from oatk_b200 import lib
    raw = lib.comm_unique_id() if rank == 0 else bytes(128)
    t = torch.tensor(list(raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


def gather_var(dist, world, dev, a):
    """variable-length all-gather of a numpy array (as int64 words)"""
    import torch
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint64:
        a = a.view(np.int64)
    else:
        a = a.astype(np.int64)
    t = torch.from_numpy(a.reshape(-1).copy()).to(dev)
    n = torch.tensor([t.numel()], device=dev)
    ns = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    mx = max(int(x.item()) for x in ns)
    pad = torch.zeros(max(mx, 1), dtype=torch.int64, device=dev)
    pad[:t.numel()] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return np.concatenate([o[:int(m.item())].cpu().numpy() for o, m in zip(outs, ns)]).view(np.uint64)


def run(dist, rank, world, local, comm=None, ctx=None, k=501, s=31, reads_per_rank=96):
    """returns {"ok": bool, "world": n, "checks": {...}} on rank 0, {"ok": bool} elsewhere"""
    import torch
    from oatk_b200 import lib, synth
    from pyoracle import Oracle, pack_reads
    dev = torch.device("cuda", local)
    reads = (synth.hifi_reads(23, 150000, reads_per_rank * world, 12000, 0.001) + synth.adversarial_reads(3, k, s)
             + synth.repeat_reads(29, 10, 9000))
    reads = reads[:len(reads) // world * world]
    per = len(reads) // world
    bases, off = pack_reads(reads[rank * per:(rank + 1) * per])
    own_ctx = ctx is None
    if own_ctx:
        ctx = lib.Context(local)
    own_comm = comm is None
    if own_comm:
        comm = lib.Comm(ctx, world, rank, share_unique_id(dist, rank, dev))
    b = lib.Batch(ctx)
    b.set_sid_base(rank * per)
    b.set_reads_host(bases, off)
    b.extract(k, s)
    comm.exchange_tuples(b)
    st = b.stat()
    loc_kc = np.array(st.kmer_cnts[:], np.int64)
    gst = comm.global_stat(b, st)
    b.count()
    got = b.count_download()
    base, total = comm.return_ids(b)
    f = b.extract_download(want_seq=False)
    arcs = {}
    for mkc, a in ((0, 0.0), (3, 0.35)):
        arcs[(mkc, a)] = comm.arcs(b, mkc, a, root=0)
    H, S_, COV, OCC, KID = (gather_var(dist, world, dev, x) for x in (got["h"], got["s"], got["cov"], got["occ"], f["k_mer"]))
    bases_all = gather_var(dist, world, dev, np.array([base, len(got["h"])], np.uint64))
    ok = True
    res = {"ok": True}
    if rank == 0:
        O = Oracle()
        ab, ao = pack_reads(reads)
        db, _ = O.extract(ab, ao, k, s)
        rc, d, i, sc, kcx = O.stat(db)
        exp = O.collect(db, len(reads))
        checks = {
            "h": bool(np.array_equal(H, exp["h"])), "s": bool(np.array_equal(S_, exp["s"])),
            "cov": bool(np.array_equal(COV.astype(np.uint32), exp["cov"])), "occ": bool(np.array_equal(OCC, exp["occ"])),
            "k_mer_id": bool(np.array_equal(KID, exp["k_mer_id"])),
            "id_bases": bool(int(bases_all[1::2].sum()) == len(exp["h"]) == total and
                             np.array_equal(bases_all[0::2], np.concatenate([[0], np.cumsum(bases_all[1::2])[:-1]]).astype(np.uint64))),
            "stat_kmer_cnts": bool(np.array_equal(np.array(gst.kmer_cnts[:], np.int64), kcx)),
            "stat_smer_cnts": bool(np.array_equal(np.array(gst.smer_cnts[:], np.int64), sc)),
            "stat_smer_unique": int(gst.smer_unique) == int(np.sum(sc)),
            "stat_gaps": bool(gst.gap_sum / gst.n_gaps == d[1]),
            "stat_syncmers": int(gst.n_syncmers) == len(exp["occ"]),
        }
        for (mkc, a), mine in arcs.items():
            theirs = O.arcs(db, exp, mkc, a)
            checks["arcs_c%d_a%g" % (mkc, a)] = bool(mine.shape == theirs.shape and np.array_equal(mine, theirs))
            checks["arcs_c%d_a%g_n" % (mkc, a)] = int(len(theirs))
        ok = all(v for k_, v in checks.items() if not k_.endswith("_n"))
        res = {"ok": ok, "world": world, "reads": len(reads), "distinct_kmers": int(len(exp["h"])), "k": k, "s": s,
               "transport": "sg_comm_* (C, NCCL grouped send/recv)", "checks": checks}
        O.free(db, exp)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    res["ok"] = bool(flag.item())
    b.close()
    if own_comm:
        comm.close()
    return res
