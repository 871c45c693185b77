// sg_hash31.cuh -- s = 31: canonical 31-mer of a position and its hash64, in a left-aligned frame.
//
// The reference hashes x (62 bits) with hash64(x, 2^62 - 1) (reference syncmer.c:116-126). Here the
// value is kept as X = x << 2 in a 64-bit pair (hi:lo): multiplications mod 2^64 are then
// multiplications mod 2^62 of x with no masking, and only the right shifts have to keep the two
// alignment bits clear. The result is bit-identical: hashA(x << 2) == hash64(x) << 2
// (tests/test_hash31.py restates both on the CPU).
//
// Cost model measured on B200 (tools/microbench/int_pipes.cu): LOP3/SHF/ISETP/SEL/VIMNMX issue at
// 0.5 warp-instructions per clock per SM sub-partition on the ALU pipe, IMAD at 0.5 on the FMA pipe
// (concurrently with the ALU pipe); IMAD.WIDE and IMAD.HI run at 0.25 AND take an ALU slot each
// ("3 LOP3 + 1 IMAD.HI" sustains 0.46, not 0.67). The ALU pipe is what bounds this code, so the
// selection below minimises ALU slots: three IMAD.WIDE + IMAD for the multiplications, plain shifts
// for the xor-shifts, one 32-bit compare for the strand (see h31_canon).
#pragma once
#include <stdint.h>

namespace sg {

// 2^(32-24), 2^(32-14), 2^(32-28): multipliers that turn "x >> n" into the high word of a product
// (kept for the microbenchmarks that compare both instruction selections)
struct H31Consts { uint32_t p8, p18, p4; };
__host__ __device__ inline H31Consts h31_consts() { return H31Consts{1u << 8, 1u << 18, 1u << 4}; }

// (hi:lo) >> 32 after a left shift by n < 32: the top word of a two-word window
__device__ __forceinline__ uint32_t h31_shf_l(uint32_t lo, uint32_t hi, uint32_t n) { return __funnelshift_l(lo, hi, n); }

// (hi:lo) *= c, + add  (mod 2^64): IMAD.WIDE + IMAD
__device__ __forceinline__ void h31_mul(uint32_t &hi, uint32_t &lo, uint32_t c, uint64_t add)
{
    const uint64_t w = (uint64_t) lo * c + add;
    hi = hi * c + (uint32_t) (w >> 32);
    lo = (uint32_t) w;
}

// X ^= X >> N with the alignment bits kept clear; pw = 2^(32-N)
template <int N>
__device__ __forceinline__ void h31_xorshift(uint32_t &hi, uint32_t &lo, uint32_t pw)
{
#ifdef SG_H31_FMA_SHIFTS
    const uint32_t t = __umulhi(hi, pw);                       // hi >> N on the FMA pipe (costs an ALU slot as well)
#else
    (void) pw;
    const uint32_t t = hi >> N;
#endif
    const uint32_t u = __funnelshift_r(lo, hi, N);
    lo ^= u & 0xfffffffcu;
    hi ^= t;
}

// hash64 of x = X >> 2, returned as the top 32 bits of hash << 2 (i.e. hash >> 30)
__device__ __forceinline__ uint32_t h31_hash_top(uint32_t hi, uint32_t lo, const H31Consts &K)
{
    h31_mul(hi, lo, 0x1fffffu, 0xfffffffffffffffcull);         // x = (x << 21) - x - 1
    h31_xorshift<24>(hi, lo, K.p8);
    h31_mul(hi, lo, 265u, 0);                                  // x = x + (x << 3) + (x << 8)
    h31_xorshift<14>(hi, lo, K.p18);
    h31_mul(hi, lo, 21u, 0);                                   // x = x + (x << 2) + (x << 4)
    h31_xorshift<28>(hi, lo, K.p4);
    return hi * 0x80000001u + __umulhi(lo, 0x80000001u);       // x += x << 31: only the top word is needed
}

// the same for a full 64-bit result (hash << 2)
__device__ __forceinline__ uint64_t h31_hash_full(uint32_t hi, uint32_t lo)
{
    uint64_t x = (uint64_t) hi << 32 | lo;
    x = x * 0x1fffffull - 4ull;
    x ^= (x >> 24) & ~3ull;
    x *= 265ull;
    x ^= (x >> 14) & ~3ull;
    x *= 21ull;
    x ^= (x >> 28) & ~3ull;
    x *= 0x80000001ull;
    return x;
}

// Canonical 31-mer ending at position J (0..15) of the sixteen bases of word w0; a, b are the two
// words before it (32 bases), ra/rb/rc the reverse-complement words of w0/b/a. All words hold their
// first base in bits 31:30. Returns the smaller strand left-aligned in (hi:lo), low two bits clear.
// The middle base of a 31-mer pairs with itself under reverse complement and comp(x) != x, and that
// base lies among the top 16 bases of both strands: the HIGH WORDS of the two strands always differ,
// so one 32-bit compare orders them exactly.
template <int J>
__device__ __forceinline__ void h31_canon(uint32_t a, uint32_t b, uint32_t w0, uint32_t ra, uint32_t rb, uint32_t rc,
        uint32_t &hi, uint32_t &lo)
{
    uint32_t fh, fl;
    if (J < 14) { fh = h31_shf_l(b, a, 2 * J + 4); fl = h31_shf_l(w0, b, 2 * J + 4); }
    else if (J == 14) { fh = b; fl = w0; }
    else { fh = h31_shf_l(w0, b, 2); fl = w0 << 2; }
    const uint32_t rh = J == 15 ? ra : h31_shf_l(rb, ra, 30 - 2 * J);
    const uint32_t rl = J == 15 ? rb : h31_shf_l(rc, rb, 30 - 2 * J);
    const bool lt = fh < rh;
    hi = min(fh, rh);
    lo = (lt ? fl : rl) & 0xfffffffcu;
}

} // namespace sg
