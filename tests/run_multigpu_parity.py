"""torchrun entry: N ranks extract their shard of reads on their own GPU, exchange tuples with one
NCCL all-to-all, count their hash range, return ids; rank 0 checks everything against the oracle
run on the whole read set. Launched by tests/test_gpu_multi.py (needs >= 2 GPUs)."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
from oatk_b200 import lib, synth, dist as sgdist   # noqa: E402
from pyoracle import Oracle, pack_reads            # noqa: E402

K, S = 501, 31


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    reads = synth.hifi_reads(23, 150000, 96 * world, 12000, 0.001) + synth.adversarial_reads(3, K, S)
    reads = reads[:len(reads) // world * world]
    per = len(reads) // world
    mine = reads[rank * per:(rank + 1) * per]
    bases, off = pack_reads(mine)
    ctx = lib.Context(local)
    b = lib.Batch(ctx)
    b.set_sid_base(rank * per)
    b.set_reads_host(bases, off)
    b.extract(K, S)
    ex = sgdist.TupleExchange(ctx, dist, rank, world)
    ex.run(b)
    st = b.stat()
    gst = ex.global_stat(b, b.stat())          # every rank ends up with the whole-input tables
    b.count()
    got = b.count_download()
    base, allc = ex.return_ids(b, len(got["h"]))
    f = b.extract_download(want_seq=False)
    # gather on rank 0
    def gather(a):
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.int64).copy() if a.dtype == np.uint64 else a.astype(np.int64)).to(dev)
        n = torch.tensor([t.numel()], device=dev)
        ns = [torch.empty_like(n) for _ in range(world)]
        dist.all_gather(ns, n)
        mx = max(int(x.item()) for x in ns)
        pad = torch.zeros(mx, dtype=torch.int64, device=dev)
        pad[:t.numel()] = t
        outs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(outs, pad)
        return np.concatenate([o[:int(m.item())].cpu().numpy() for o, m in zip(outs, ns)]).view(np.uint64)
    H, COV, OCC, KID = gather(got["h"]), gather(got["cov"]), gather(got["occ"]), gather(f["k_mer"])
    kc = torch.tensor(np.array(st.kmer_cnts[:], np.int64), device=dev)
    dist.all_reduce(kc)
    gaps = torch.tensor([st.gap_sum, st.n_gaps], device=dev, dtype=torch.int64)
    dist.all_reduce(gaps)
    ok = True
    if rank == 0:
        O = Oracle()
        ab, ao = pack_reads(reads)
        db, _ = O.extract(ab, ao, K, S)
        rc, d, i, sc, kcx = O.stat(db)
        exp = O.collect(db, len(reads))
        checks = {"h": np.array_equal(H, exp["h"]), "cov": np.array_equal(COV.astype(np.uint32), exp["cov"]),
                  "occ": np.array_equal(OCC, exp["occ"]), "k_mer_id": np.array_equal(KID, exp["k_mer_id"]),
                  "kmer_cnts": np.array_equal(kc.cpu().numpy(), kcx),
                  "avg_dist": gaps[0].item() / gaps[1].item() == d[1],
                  "global_kmer_cnts": np.array_equal(np.array(gst.kmer_cnts[:], np.int64), kcx),
                  "global_smer_cnts": np.array_equal(np.array(gst.smer_cnts[:], np.int64), sc),
                  "global_smer_unique": int(gst.smer_unique) == int(np.sum(sc)),
                  "global_gaps": gst.gap_sum / gst.n_gaps == d[1]}
        ok = all(checks.values())
        print("MULTIGPU_PARITY", "OK" if ok else "FAIL", checks, "world", world, "distinct", len(exp["h"]), "per-rank", allc)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
