"""The s = 31 kernel path (oatk_b200/csrc/sg_hash31.cuh) hashes x in a left-aligned frame X = x << 2 and takes
every canonical 31-mer straight out of three packed words. Both steps are restated here in numpy/Python and
checked against the reference formulation (hash64, reference syncmer.c:116-126; rolling canonical s-mer,
syncmer.c:307-315) so that the algebra is pinned on the CPU before the GPU parity tests run."""
import random

import numpy as np

M64 = (1 << 64) - 1
M32 = (1 << 32) - 1


def hash64_ref(x, mask):
    x = ((x << 21) - x - 1) & mask
    x ^= x >> 24
    x = (x * 265) & mask
    x ^= x >> 14
    x = (x * 21) & mask
    x ^= x >> 28
    x = (x + (x << 31)) & mask
    return x


def hash_aligned(X):
    X = (X * ((1 << 21) - 1) - 4) & M64
    X ^= (X >> 24) & ~3
    X = (X * 265) & M64
    X ^= (X >> 14) & ~3
    X = (X * 21) & M64
    X ^= (X >> 28) & ~3
    X = (X * ((1 << 31) + 1)) & M64
    return X


def test_aligned_hash_equals_hash64():
    rnd = random.Random(7)
    mask = (1 << 62) - 1
    xs = [0, 1, mask, mask - 1, 0x0123456789ABCDEF & mask] + [rnd.getrandbits(62) for _ in range(20000)]
    for x in xs:
        assert hash_aligned(x << 2) == hash64_ref(x, mask) << 2
    # SURVEY appendix A.4 known answers
    assert hash_aligned(0) >> 2 == 2158324264573792932
    assert hash_aligned(1 << 2) >> 2 == 2002549777813010638
    assert hash_aligned(mask << 2) >> 2 == 4015643844226056017


def test_aligned_hash_vectorised_matches_oracle_words():
    # the kernel works on (hi, lo) 32-bit halves with IMAD.WIDE / IMAD.HI; restate that split exactly
    rnd = np.random.default_rng(3)
    x = rnd.integers(0, 1 << 62, 5000, dtype=np.uint64)
    X = x << np.uint64(2)
    hi = (X >> np.uint64(32)).astype(np.uint64)
    lo = (X & np.uint64(M32)).astype(np.uint64)

    def mul(hi, lo, c, add):
        w = lo * np.uint64(c) + np.uint64(add)          # < 2^64 for c < 2^31 and add < 2^64 - c*2^32: wraps like the GPU
        return (hi * np.uint64(c) + (w >> np.uint64(32))) & np.uint64(M32), w & np.uint64(M32)

    def xs(hi, lo, n):
        t = hi >> np.uint64(n)
        u = ((hi << np.uint64(32 - n)) | (lo >> np.uint64(n))) & np.uint64(M32)
        return hi ^ t, lo ^ (u & np.uint64(0xfffffffc))

    with np.errstate(over="ignore"):
        hi, lo = mul(hi, lo, 0x1fffff, 0xfffffffffffffffc)
        hi, lo = xs(hi, lo, 24)
        hi, lo = mul(hi, lo, 265, 0)
        hi, lo = xs(hi, lo, 14)
        hi, lo = mul(hi, lo, 21, 0)
        hi, lo = xs(hi, lo, 28)
        top = (hi * np.uint64(0x80000001) + ((lo * np.uint64(0x80000001)) >> np.uint64(32))) & np.uint64(M32)
    mask = (1 << 62) - 1
    exp = np.array([hash64_ref(int(v), mask) >> 30 for v in x], dtype=np.uint64)
    assert np.array_equal(top, exp)


def _rev2(x):
    r = 0
    for i in range(16):
        r |= ((x >> (2 * i)) & 3) << (2 * (15 - i))
    return r


def _shf_l(lo, hi, n):
    return (((hi << 32 | lo) << n) >> 32) & M32


def test_canonical_31mer_extraction():
    rnd = random.Random(11)
    for _ in range(300):
        a, b, w0 = (rnd.getrandbits(32) for _ in range(3))
        bases = [(w >> (30 - 2 * i)) & 3 for w in (a, b, w0) for i in range(16)]
        ra, rb, rc = _rev2(~w0 & M32), _rev2(~b & M32), _rev2(~a & M32)
        for j in range(16):
            sm = bases[32 + j - 30:32 + j + 1]
            fw = 0
            for c in sm:
                fw = fw << 2 | c
            rv = 0
            for c in reversed(sm):
                rv = rv << 2 | (3 - c)
            n = 2 * j + 4
            if n < 32:
                fh, fl = _shf_l(b, a, n), _shf_l(w0, b, n)
            elif n == 32:
                fh, fl = b, w0
            else:
                fh, fl = _shf_l(w0, b, 2), (w0 << 2) & M32
            m = 30 - 2 * j
            rh, rl = (ra, rb) if j == 15 else (_shf_l(rb, ra, m), _shf_l(rc, rb, m))
            assert fh != rh                                   # the middle base pairs with itself: one 32-bit compare decides
            lt = fh < rh
            hi, lo = (fh, fl) if lt else (rh, rl)
            assert (hi << 32 | (lo & 0xfffffffc)) == min(fw, rv) << 2
