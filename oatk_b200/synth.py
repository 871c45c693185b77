"""Deterministic synthetic HiFi-like reads and the
adversarial parity set (tandem repeats, N runs, long homopolymers, short and empty
reads, lower case, IUPAC codes, palindromes)."""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def hifi_reads(seed, genome_len, n_reads, read_len, err):
    """The survey's generator: uniform random circular genome, fixed-length windows,
    random strand, per-base error rate `err` split evenly over sub/ins/del."""
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, genome_len, dtype=np.uint8)
    g2 = np.concatenate([genome, genome[:read_len]]) if read_len <= genome_len else np.tile(genome, read_len // genome_len + 2)[:genome_len + read_len]
    out = []
    for _ in range(n_reads):
        p = int(rng.integers(0, genome_len))
        r = g2[p:p + read_len].copy()
        if rng.integers(0, 2):
            r = (3 - r)[::-1]
        ne = rng.binomial(read_len, err)
        if ne:
            pos = rng.integers(0, read_len, ne)
            kind = rng.integers(0, 3, ne)
            sub = pos[kind == 0]
            r[sub] = (r[sub] + rng.integers(1, 4, len(sub))) % 4
            dele = pos[kind == 1]
            keep = np.ones(read_len, bool)
            keep[dele] = False
            ins = np.sort(pos[kind == 2])
            r = r[keep]
            if len(ins):
                r = np.insert(r, np.minimum(ins, len(r)), rng.integers(0, 4, len(ins)).astype(np.uint8))
        out.append(ACGT[r].tobytes())
    return out


def _rand(rng, n):
    return ACGT[rng.integers(0, 4, n)].tobytes()


def _nohp(rng, n):
    """random sequence without homopolymer runs (hoco length == raw length)"""
    x = rng.integers(0, 4, n)
    step = rng.integers(1, 4, n)
    for i in range(1, n):
        if x[i] == x[i - 1]:
            x[i] = (x[i - 1] + step[i]) % 4
    return ACGT[x].tobytes()


def hash64(x, mask):
    """the reference's invertible s-mer hash (syncmer.c:116-126) on python ints; used only to
    plant minimisers in the adversarial reads below"""
    x = ((x << 21) - x - 1) & mask
    x ^= x >> 24
    x = (x * 265) & mask
    x ^= x >> 14
    x = (x * 21) & mask
    x ^= x >> 28
    x = (x * 2147483649) & mask
    return x


def _smer_hashes(seq, s):
    """hash of the canonical s-mer ending at every position of an ACGT byte string (None = none)"""
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    mask = (1 << (2 * s)) - 1
    fw = rv = 0
    out = []
    for i, ch in enumerate(seq):
        c = code[ch]
        fw = ((fw << 2) | c) & mask
        rv = (rv >> 2) | ((3 - c) << (2 * (s - 1)))
        out.append(None if i + 1 < s or fw == rv else hash64(min(fw, rv), mask))
    return out


def tie_rule_read(rng, k, s):
    """A read that makes the reference's tie clause (syncmer.c:356-377) reject a CLOSE: an s-mer Y
    occurs twice inside one window, is the window minimum, its older copy is not in the oldest
    slot, and the element that just left the window (X) is smaller still. Returns None when
    k - s + 1 is too small to hold the construction."""
    q = k - s + 1
    if q < 2 * s + 4 or s < 4:
        return None
    for _ in range(200):
        cands = [_nohp(rng, s) for _ in range(400 if s > 8 else 60)]
        hs = sorted((h[-1], c) for c in cands for h in [_smer_hashes(c, s)] if h[-1] is not None)
        (hx, X), (hy, Y) = hs[0], hs[1]
        if hx == hy:
            continue
        a = (q - 2 * s) // 3
        b = q - 2 * s - a
        for _ in range(300):
            read = _nohp(rng, k) + X + _nohp(rng, a) + Y + _nohp(rng, b) + Y + _nohp(rng, k)
            if any(read[i] == read[i + 1] for i in range(len(read) - 1)):
                continue
            h = _smer_hashes(read, s)
            p = k + s + a + s + b + s - 1                      # end of the second Y
            win = [v for v in h[p - q + 1:p] if v is not None]
            if h[p] == hy and h[p - q] == hx and min(win) == hy and h[p - q + 1] != hy:
                return read
    return None


def revcomp(b):
    t = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")
    return b.translate(t)[::-1]


def adversarial_reads(seed, k, s, scale=1):
    """Reads built to hit every branch of the extractor (syncmer.c:284-394).
    Lengths are expressed in units of k so the same set works for any (k, s)."""
    rng = np.random.default_rng(seed)
    L = max(4 * k, 600) * scale
    out = []
    out.append(b"")                                            # empty read
    out.append(b"A")                                           # single base
    out.append(_rand(rng, s))                                  # shorter than k
    out.append(_nohp(rng, k - 1))                              # one short of a k-mer
    out.append(_nohp(rng, k))                                  # exactly one k-mer
    out.append(_nohp(rng, k + 1))
    out.append(_rand(rng, L))                                  # plain random
    out.append(_rand(rng, L).lower())                          # lower case
    r = bytearray(_rand(rng, L))
    for p in rng.integers(0, L, 5):
        r[p] = ord("N")
    out.append(bytes(r))                                       # isolated Ns
    r = bytearray(_rand(rng, L))
    r[L // 2:L // 2 + 40] = b"N" * 40
    out.append(bytes(r))                                       # an N run (not compressed)
    r = bytearray(_rand(rng, L))
    for p, ch in zip(rng.integers(0, L, 8), b"RYKMSWBD"):
        r[p] = ch
    out.append(bytes(r))                                       # IUPAC codes, all ambiguous
    out.append(_rand(rng, L // 2) + b"A" * 300 + _rand(rng, L // 2))     # homopolymer > 255
    out.append(_rand(rng, L // 2) + b"C" * 256 + _rand(rng, L // 2))     # exactly 256
    out.append(_rand(rng, L // 2) + b"G" * 255 + _rand(rng, L // 2))     # exactly 255
    out.append(b"T" * 1000 + _rand(rng, L))                    # starts with a long run
    out.append(_rand(rng, L) + b"a" * 700)                     # ends with a long run
    unit = _nohp(rng, 37)
    out.append(unit * (L // 37 + 2))                           # tandem repeat, period 37 (< q)
    unit = _nohp(rng, max(k - s + 1, 8))
    out.append(unit * 5)                                       # period exactly q: ties at both window ends
    unit = _nohp(rng, max(k - s, 7))
    out.append(unit * 5)                                       # period q-1
    unit = _nohp(rng, max(k - s + 2, 9))
    out.append(unit * 5)                                       # period q+1
    unit = _nohp(rng, max(k // 3, 5))
    out.append(_rand(rng, k) + unit * 9 + _rand(rng, k))       # repeat embedded in unique sequence
    out.append(b"AC" * (L // 2))                               # dinucleotide repeat (period 2)
    out.append(b"ACG" * (L // 3))
    half = _nohp(rng, L // 2)
    out.append(half + revcomp(half))                           # perfect palindrome (hairpin)
    out.append(half + b"N" + revcomp(half))
    x = _nohp(rng, L)
    out.append(x)
    out.append(revcomp(x))                                     # the same read from the other strand
    r = bytearray(_nohp(rng, 2 * k + 10))
    r[k] = ord("N")
    out.append(bytes(r))                                       # k-mer immediately followed by N: no OPEN
    r = bytearray(_nohp(rng, 2 * k + 10))
    r[k + 5] = ord("n")
    out.append(bytes(r))
    out.append(b"N" * 50)                                      # only ambiguous
    out.append(b"ACGT" * 3 + b"U" * 5 + b"u" * 3 + _rand(rng, L))        # U is T
    out.append(_rand(rng, L).replace(b"A", b"AA"))             # many short runs
    out.append(b"\x00\x01\x02\x03" * 50 + _rand(rng, L))       # raw codes 0..3 are bases too (nt4 table)
    for _ in range(6 * scale):                                 # a few long reads with sparse errors
        out.append(_rand(rng, int(rng.integers(L, 3 * L))))
    t = tie_rule_read(rng, k, s)                               # planted tie: the CLOSE tie clause must reject
    if t is not None:
        out.append(t)
        out.append(revcomp(t))
    return out


REPEAT_PERIODS = (2, 3, 6, 37, 171)


def repeat_reads(seed, n_reads, read_len, min_arr=2000):
    """Reads that carry one tandem array each (period 2, 3, 37, 171 or the telomere unit TTAGGG,
    2 kb up to the whole read) inside random sequence: the input class on which every window
    position ties for the minimum (the scan kernel's cliff)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_reads):
        p = REPEAT_PERIODS[i % len(REPEAT_PERIODS)]
        unit = b"TTAGGG" if p == 6 else _nohp(rng, p)
        la = int(rng.integers(min(min_arr, read_len), read_len + 1))
        st = int(rng.integers(0, read_len - la + 1))
        arr = (unit * (la // p + 2))[:la]
        r = _rand(rng, st) + arr + _rand(rng, read_len - st - la)
        if rng.integers(0, 2):
            r = revcomp(r)
        out.append(r)
    return out
