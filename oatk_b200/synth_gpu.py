"""Synthetic HiFi-like reads generated on the GPU with torch (bench.py workload).

Same model as oatk_b200/synth.py: uniform random circular genome,
reads of exactly `read_len` bases from uniform random starts, random strand, per-base
error rate `err` split evenly over substitution / insertion / deletion. torch is used
here only to fill device memory; nothing on the measured path goes through it.
"""
import torch


def hifi_reads_gpu(seed, genome_len, n_reads, read_len, err, device, chunk=16384, genome_seed=None):
    """returns (bases uint8 [n_reads*read_len] on device, offsets uint64-as-int64 [n_reads+1] on device)"""
    g = torch.Generator(device=device)
    g.manual_seed(seed if genome_seed is None else genome_seed)      # all ranks sequence the same genome
    genome = torch.randint(0, 4, (genome_len,), dtype=torch.uint8, device=device, generator=g)
    g.manual_seed(seed)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=device)
    out = torch.empty(n_reads * read_len, dtype=torch.uint8, device=device)
    for r0 in range(0, n_reads, chunk):
        b = min(chunk, n_reads - r0)
        start = torch.randint(0, genome_len, (b, 1), device=device, generator=g)
        u = torch.rand((b, read_len), device=device, generator=g)
        ins = u < err / 3
        dele = (u >= err / 3) & (u < 2 * err / 3)
        sub = (u >= 2 * err / 3) & (u < err)
        adv = torch.ones((b, read_len), dtype=torch.int32, device=device)
        adv -= ins.to(torch.int32)
        adv += dele.to(torch.int32)
        src = torch.cumsum(adv, dim=1, dtype=torch.int64)
        src += start - 1
        src %= genome_len
        base = genome[src]
        del src, adv, u
        rnd = torch.randint(0, 4, (b, read_len), dtype=torch.uint8, device=device, generator=g)
        base = torch.where(ins, rnd, base)
        base = torch.where(sub, (base + 1 + rnd % 3) % 4, base)
        rev = torch.rand((b, 1), device=device, generator=g) < 0.5
        base = torch.where(rev, (3 - base).flip(1), base)
        out[r0 * read_len:(r0 + b) * read_len] = lut[base.long()].reshape(-1)
        del base, rnd, ins, dele, sub
    off = torch.arange(0, n_reads + 1, dtype=torch.int64, device=device) * read_len
    return out, off
