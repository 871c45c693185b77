// hash31.cuh -- s = 31: hash of the canonical 31-mer in a left-aligned 64-bit frame
#pragma once
#include <stdint.h>
namespace sg {

#ifndef H31_FMA_SHIFTS
#define H31_FMA_SHIFTS 0
#endif

#ifndef H31_FMA_FUNNELS
#define H31_FMA_FUNNELS 0      // how many of the three xorshifts build their funnel shift from two IMADs
#endif
// powers of two kept in registers (loaded from kernel parameters) so that ptxas keeps the multiplies
// on the FMA pipe instead of turning them back into shifts; m4 = -4 as a 64-bit addend
struct H31Consts { uint32_t p8, p18, p4, pad; uint64_t m4; };

__device__ __forceinline__ uint32_t shf_l(uint32_t lo, uint32_t hi, uint32_t n)
{
    uint32_t d; asm("shf.l.clamp.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(n)); return d;
}
__device__ __forceinline__ uint32_t shf_r(uint32_t lo, uint32_t hi, uint32_t n)
{
    uint32_t d; asm("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(lo), "r"(hi), "r"(n)); return d;
}
// (hi:lo) * c + add  (mod 2^64)
__device__ __forceinline__ void mul64c(uint32_t &hi, uint32_t &lo, uint32_t c, uint64_t add)
{
    uint64_t w;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w) : "r"(lo), "r"(c), "l"(add));
    uint32_t wl, wh;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(wl), "=r"(wh) : "l"(w));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hi) : "r"(hi), "r"(c), "r"(wh));
    lo = wl;
}
__device__ __forceinline__ void mul64c0(uint32_t &hi, uint32_t &lo, uint32_t c)
{
    uint64_t w;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(lo), "r"(c));
    uint32_t wl, wh;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(wl), "=r"(wh) : "l"(w));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hi) : "r"(hi), "r"(c), "r"(wh));
    lo = wl;
}
// X ^= (X >> n) with the two alignment bits kept clear
template <int N, bool FUNNEL_FMA>
__device__ __forceinline__ void xorshift(uint32_t &hi, uint32_t &lo, uint32_t pw)
{
    uint32_t t, u;
#if H31_FMA_SHIFTS
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(t) : "r"(hi), "r"(pw));      // hi >> N on the FMA pipe
#else
    t = hi >> N;
#endif
    if (FUNNEL_FMA) {                                                 // (hi:lo) >> N = hi * 2^(32-N) + (lo >> N)
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(u) : "r"(lo), "r"(pw));
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(u) : "r"(hi), "r"(pw), "r"(u));
    } else u = shf_r(lo, hi, N);
    lo = lo ^ (u & 0xfffffffcu);
    hi = hi ^ t;
}
// hash64 (reference syncmer.c:116-126) of x = X >> 2 with mask 2^62 - 1, returned as (hash << 2) >> 32
__device__ __forceinline__ uint32_t hash31_hi(uint32_t hi, uint32_t lo, const H31Consts &K)
{
    mul64c(hi, lo, 0x1fffffu, K.m4);     // x = (x << 21) - x - 1
    xorshift<24, (H31_FMA_FUNNELS > 0)>(hi, lo, K.p8);
    mul64c0(hi, lo, 265u);
    xorshift<14, (H31_FMA_FUNNELS > 1)>(hi, lo, K.p18);
    mul64c0(hi, lo, 21u);
    xorshift<28, (H31_FMA_FUNNELS > 2)>(hi, lo, K.p4);
    uint32_t r;                                            // x += x << 31: only the high word is needed
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(lo), "r"(0x80000001u));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(hi), "r"(0x80000001u), "r"(r));
    return r;
}
// canonical aligned 31-mer ending at position J of the word w0 (a, b = the two words before it);
// ra, rb, rc = reverse-complement words of w0, b, a
template <int J>
__device__ __forceinline__ void canon31(uint32_t a, uint32_t b, uint32_t w0, uint32_t ra, uint32_t rb, uint32_t rc,
        uint32_t &hi, uint32_t &lo, bool &palin)
{
    uint32_t fh, fl;
    if (J < 14) { fh = shf_l(b, a, 2 * J + 4); fl = shf_l(w0, b, 2 * J + 4) & 0xfffffffcu; }
    else if (J == 14) { fh = b; fl = w0 & 0xfffffffcu; }
    else { fh = shf_l(w0, b, 2); fl = w0 << 2; }
    const uint32_t rh = J == 15 ? ra : shf_l(rb, ra, 30 - 2 * J);
    const uint32_t rl = (J == 15 ? rb : shf_l(rc, rb, 30 - 2 * J)) & 0xfffffffcu;
    const uint64_t f64 = (uint64_t) fh << 32 | fl, r64 = (uint64_t) rh << 32 | rl;
    const bool lt = f64 < r64;
    palin = f64 == r64;
    hi = lt ? fh : rh;
    lo = lt ? fl : rl;
}
} // namespace sg
