"""Multi-GPU plumbing for the syncmer counter: one process per GPU, torch.distributed for transport.

The path shards by read (contiguous blocks, so sid = global read index). Extraction needs no
communication. Counting needs every occurrence of a k-mer on one GPU, so the (hash, occ, s_mer, fingerprint)
tuples are range-partitioned on the hash and exchanged with ONE all-to-all over NVLink; each GPU
then owns a contiguous hash range and its local hash order is the reference's global order
restricted to that range. Dense ids are local ranks plus the number of distinct k-mers on the
lower ranks (one all-gather of a scalar). A second all-to-all returns (occ, id) pairs to the
GPU that holds the read.

The functions that take tensors are device agnostic so that the split/offset logic is covered by
world_size-2 gloo tests on CPU (tests/test_dist_cpu.py).
"""
import torch


class _DevPtr:
    """zero-copy view of a device allocation owned by libsyncgpu"""

    def __init__(self, ptr, n, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def tensor_from_ptr(ptr, n, device):
    if n == 0 or not ptr:
        return torch.empty(0, dtype=torch.int64, device=device)
    return torch.as_tensor(_DevPtr(ptr, n), device=device)


def exchange_counts(dist, send_counts, device):
    """send_counts[p] = items this rank sends to rank p; returns items received from every rank"""
    send = torch.tensor(send_counts, dtype=torch.int64, device=device)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    return [int(x) for x in recv.tolist()]


def exchange_rows(dist, rows, send_counts, recv_counts, width):
    """rows: flat int64 tensor of len(sum(send_counts)) * width, grouped by destination rank"""
    out = torch.empty(sum(recv_counts) * width, dtype=torch.int64, device=rows.device)
    dist.all_to_all_single(out, rows, output_split_sizes=[c * width for c in recv_counts],
                           input_split_sizes=[c * width for c in send_counts])
    return out


def id_base(dist, n_unique, rank, world, device):
    """distinct k-mers owned by the lower ranks = what to add to a local id to make it global"""
    mine = torch.tensor([n_unique], dtype=torch.int64, device=device)
    allc = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine)
    counts = [int(t.item()) for t in allc]
    return sum(counts[:rank]), counts


def range_part(keys, world):
    """the partition rule of sg_tuples_partition (csrc/sg_arcs.cu part_key_kernel) on a tensor of
    uint64 hashes held as int64: part = floor(((key >> 1) << 1) * world / 2^64)"""
    import numpy as np
    k = keys.cpu().numpy().view(np.uint64)
    k = (k >> np.uint64(1)) << np.uint64(1)
    hi = (k >> np.uint64(32)).astype(np.uint64)
    lo = (k & np.uint64(0xFFFFFFFF)).astype(np.uint64)
    w = np.uint64(world)
    # 64x64 -> high 64 bits with 32-bit limbs (world < 2^32)
    t = lo * w
    part = (hi * w + (t >> np.uint64(32))) >> np.uint64(32)
    return torch.from_numpy(part.astype(np.int64))


class TupleExchange:
    """the exchange step between sg_extract and sg_stat / sg_count on every rank"""

    def __init__(self, ctx, dist, rank, world):
        self.ctx, self.dist, self.rank, self.world = ctx, dist, rank, world
        self.device = torch.device("cuda", ctx.device)
        self.send_counts = self.recv_counts = None
        self.bytes_sent = 0

    def run(self, batch):
        counts, ptr = batch.tuples_partition(self.world)
        self.send_counts = counts
        self.recv_counts = exchange_counts(self.dist, counts, self.device)
        rows = tensor_from_ptr(ptr, sum(counts) * 4, self.device)
        got = exchange_rows(self.dist, rows, counts, self.recv_counts, 4)
        self.bytes_sent = (sum(counts) - counts[self.rank]) * 32
        batch.tuples_adopt(got.data_ptr(), sum(self.recv_counts))
        torch.cuda.current_stream().synchronize()      # `got` may be freed once the adopt kernel has read it
        return sum(self.recv_counts)

    def global_stat(self, batch, st):
        """sr_db_stat's tables over ALL reads from the per-rank sg_stat results: k-mer tables and gap sums add up
        (every k-mer lives on one rank), s-mer codes are all-gathered as (code, local count) pairs and merged"""
        dev = self.device
        ptr, n = batch.smer_counts_pack()
        mine = tensor_from_ptr(ptr, n * 2, dev)
        ns = [torch.empty(1, dtype=torch.int64, device=dev) for _ in range(self.world)]
        self.dist.all_gather(ns, torch.tensor([n], dtype=torch.int64, device=dev))
        ns = [int(x.item()) for x in ns]
        pad = torch.zeros(2 * max(max(ns), 1), dtype=torch.int64, device=dev)
        pad[:2 * n] = mine
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(parts, pad)
        allp = torch.cat([p[:2 * m] for p, m in zip(parts, ns)]) if sum(ns) else torch.empty(0, dtype=torch.int64, device=dev)
        batch.smer_counts_merge(allp.data_ptr() if allp.numel() else None, sum(ns), st)
        torch.cuda.current_stream().synchronize()
        kc = torch.tensor(list(st.kmer_cnts), dtype=torch.int64, device=dev)
        misc = torch.tensor([st.gap_sum, st.n_gaps, st.kmer_unique, st.kmer_singleton, st.n_syncmers], dtype=torch.int64, device=dev)
        self.dist.all_reduce(kc)
        self.dist.all_reduce(misc)
        for i, v in enumerate(kc.tolist()):
            st.kmer_cnts[i] = v
        st.gap_sum, st.n_gaps, st.kmer_unique, st.kmer_singleton, st.n_syncmers = (int(x) for x in misc.tolist())
        return st

    def return_ids(self, batch, n_unique):
        """after sg_count: global ids back to the ranks that hold the reads"""
        base, all_counts = id_base(self.dist, n_unique, self.rank, self.world, self.device)
        ptr, n = batch.ids_pack(base)
        rows = tensor_from_ptr(ptr, n * 2, self.device)
        back = exchange_rows(self.dist, rows, self.recv_counts, self.send_counts, 2)
        batch.ids_scatter(back.data_ptr(), sum(self.send_counts))
        return base, all_counts
