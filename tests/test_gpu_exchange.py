"""the multi-GPU exchange kernels on one GPU: partition by hash range, adopt, count per range,
pack / scatter ids -- stitched together here the way oatk_b200/dist.py does across ranks"""
import numpy as np
import pytest
import torch
from oatk_b200 import synth, dist as sgdist
from pyoracle import pack_reads
import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_adopt_count(gpu_ctx, oracle, world):
    from oatk_b200 import lib
    reads = synth.hifi_reads(13, 150000, 300, 12000, 0.001) + synth.adversarial_reads(3, 501, 31)
    bases, off = pack_reads(reads)
    db, _ = oracle.extract(bases, off, 501, 31)
    exp = oracle.collect(db, len(reads))

    b = lib.Batch(gpu_ctx)
    b.set_reads_host(bases, off)
    b.extract(501, 31)
    counts, ptr = b.tuples_partition(world)
    N = sum(counts)
    assert N == len(exp["occ"])
    dev = torch.device("cuda", 0)
    rows = sgdist.tensor_from_ptr(ptr, N * 4, dev).clone().view(-1, 4)
    # every part holds exactly the keys of its range, in (sid, idx) order
    keys = rows[:, 0].contiguous()
    part = sgdist.range_part(keys, world).numpy()
    assert np.array_equal(part, np.repeat(np.arange(world), counts))
    occ = rows[:, 1].cpu().numpy().view(np.uint64)
    startp = np.concatenate([[0], np.cumsum(counts)])
    for p in range(world):
        o = occ[startp[p]:startp[p + 1]]
        assert np.all(o[1:] > o[:-1])

    # count every range as its owner would, then stitch
    hs, covs, occs, pairs = [], [], [], []
    base = 0
    for p in range(world):
        seg = rows[startp[p]:startp[p + 1]].contiguous()
        if counts[p] == 0:
            continue
        b.tuples_adopt(seg.data_ptr(), counts[p])
        b.count()
        got = b.count_download()
        hs.append(got["h"]); covs.append(got["cov"]); occs.append(got["occ"])
        pp, n = b.ids_pack(base)
        pairs.append(sgdist.tensor_from_ptr(pp, n * 2, dev).clone())
        base += len(got["h"])
    assert np.array_equal(np.concatenate(hs), exp["h"])
    assert np.array_equal(np.concatenate(covs), exp["cov"])
    assert np.array_equal(np.concatenate(occs), exp["occ"])
    allpairs = torch.cat(pairs)
    b.ids_scatter(allpairs.data_ptr(), N)
    f = b.extract_download(want_seq=False)
    assert np.array_equal(f["k_mer"], exp["k_mer_id"])
    oracle.free(db, exp)
    b.close()
