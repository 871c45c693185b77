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


REPEAT_PERIODS = (2, 3, 6, 37, 171)


def plant_repeats_gpu(bases, n_reads, read_len, every, seed, device):
    """Overwrites part of every `every`-th read with one tandem array (period 2, 3, 37, 171 or the
    telomere unit TTAGGG; 2 kb up to the whole read), in place: bench.py --workload repeats."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    sel = torch.arange(0, n_reads, every, device=device)
    m = sel.numel()
    per = torch.tensor(REPEAT_PERIODS, device=device)[torch.arange(m, device=device) % len(REPEAT_PERIODS)]
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=device)
    unit = torch.randint(0, 4, (m, 171), dtype=torch.uint8, device=device, generator=g)
    for j in range(1, 171):                                   # no homopolymer inside a unit
        same = unit[:, j] == unit[:, j - 1]
        unit[:, j] = torch.where(same, (unit[:, j] + 1) % 4, unit[:, j])
    tel = torch.tensor([3, 3, 0, 2, 2, 2], dtype=torch.uint8, device=device)
    unit[per == 6, :6] = tel
    lo = min(2000, read_len)
    la = torch.randint(lo, read_len + 1, (m,), device=device, generator=g)
    st = (torch.rand(m, device=device, generator=g) * (read_len - la + 1).float()).long().clamp_(min=0)
    st = torch.minimum(st, read_len - la)
    view = bases.view(n_reads, read_len)
    idx = torch.arange(read_len, device=device).view(1, -1)
    for c0 in range(0, m, 4096):
        c1 = min(m, c0 + 4096)
        rel = idx - st[c0:c1].view(-1, 1)
        inside = (rel >= 0) & (rel < la[c0:c1].view(-1, 1))
        rep = lut[unit[c0:c1].gather(1, (rel.clamp(min=0) % per[c0:c1].view(-1, 1))).long()]
        rows = sel[c0:c1]
        view[rows] = torch.where(inside, rep, view[rows])
    return int(m)
