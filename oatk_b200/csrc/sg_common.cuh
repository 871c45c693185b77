// sg_common.cuh -- shared device helpers and internal declarations of libsyncgpu.
//
// Device data layout ("sparse" layout, indexed by capacity offsets):
//   hoff[r]   = sum_{j<r} roundup64(len_j)      capacity offset of read r in hoco positions
//   hoco_s    : 2 bits / hoco base, 4 per byte, first base in bits 7:6 (reference
//               syncmer.h:53-55); read r starts at byte hoff[r]/4 (16-byte aligned)
//   ho_rl     : 1 byte / hoco base (run length - 1, saturating at 255); read r at hoff[r]
//   nbits     : 1 bit / hoco base, set for an ambiguous base, LSB-first; read r at byte hoff[r]/8
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SG_FULL 0xffffffffu
#define SG_NONE64 0xffffffffffffffffull

namespace sg {

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// reverse the order of the sixteen 2-bit groups of a word
__device__ __forceinline__ uint32_t rev2(uint32_t x)
{
    x = __brev(x);
    return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
}

// reverse complement of 32 bases held big-endian (first base in bits 63:62)
__device__ __forceinline__ uint64_t rc64(uint64_t v)
{
    v = ~v;
    return (uint64_t) rev2((uint32_t) v) << 32 | rev2((uint32_t) (v >> 32));
}

// invertible mix on 2s-bit values (reference syncmer.c:116-126), written as the
// three multiplications it is
__device__ __forceinline__ uint64_t hash64(uint64_t x, uint64_t mask)
{
    x = ((x << 21) - x - 1) & mask;
    x ^= x >> 24;
    x = (x * 265) & mask;
    x ^= x >> 14;
    x = (x * 21) & mask;
    x ^= x >> 28;
    x = (x + (x << 31)) & mask;
    return x;
}

// sixteen hoco bases starting at a multiple of 16, first base in bits 31:30; words
// outside [0, nwords) read as zero
__device__ __forceinline__ uint32_t hoco_word(const uint32_t *hs, int64_t w, int64_t nwords)
{
    return (w >= 0 && w < nwords) ? bswap32(__ldg(hs + w)) : 0u;
}

// 32 consecutive bases starting at hoco position p (any p >= -32), first base on top
__device__ __forceinline__ uint64_t hoco_window(const uint32_t *hs, int64_t p, int64_t nwords)
{
    int64_t w = p >> 4;               // arithmetic shift: floor for negatives
    int sh = (int) (p & 15) * 2;
    uint32_t a = hoco_word(hs, w, nwords), b = hoco_word(hs, w + 1, nwords), c = hoco_word(hs, w + 2, nwords);
    uint32_t hi = __funnelshift_l(b, a, sh), lo = __funnelshift_l(c, b, sh);
    return (uint64_t) hi << 32 | lo;
}

// canonical s-mer ending at hoco position p: returns canon << 1 | z (z = 1 when
// the reverse strand is the smaller one), the reference's s_mer code (syncmer.c:307-315)
__device__ __forceinline__ uint64_t smer_code_at(const uint32_t *hs, int64_t p, int s, int64_t nwords)
{
    uint64_t v = hoco_window(hs, p - 31, nwords);          // bases p-31 .. p
    uint64_t mask = (1ull << (2 * s)) - 1;
    uint64_t fw = v & mask;
    uint64_t rv = rc64(v) >> (64 - 2 * s);
    return fw < rv ? fw << 1 : (rv << 1 | 1);
}

__device__ __forceinline__ uint64_t bswap64(uint64_t v)
{
    return (uint64_t) bswap32((uint32_t) v) << 32 | bswap32((uint32_t) (v >> 32));
}

// 32 oriented bases number 32j .. 32j+31 of the k-mer, first base on top, bases
// past the end of the k-mer zeroed
__device__ __forceinline__ uint64_t oriented_block(const uint32_t *hs, int64_t nwords, int64_t start, int k, int rev, int j)
{
    uint64_t v;
    if (!rev) v = hoco_window(hs, start + 32 * (int64_t) j, nwords);
    else v = rc64(hoco_window(hs, start + k - 32 * (int64_t) (j + 1), nwords));
    const int left = k - 32 * j;                       // bases of the k-mer in this block
    if (left < 32) v &= ~0ull << (64 - 2 * left);
    return v;
}

struct BlockScanU32 {
    // exclusive scan of one value per thread over a block of NW warps; returns the
    // exclusive prefix, total through *total. smem must hold NW+1 words.
    template <int NW>
    static __device__ __forceinline__ uint32_t run(uint32_t v, uint32_t *smem, uint32_t *total)
    {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(SG_FULL, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) smem[wid] = inc;
        __syncthreads();
        uint32_t base = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            uint32_t t = smem[w];
            if (w < wid) base += t;
            tot += t;
        }
        __syncthreads();
        *total = tot;
        return base + inc - v;
    }
};

} // namespace sg
