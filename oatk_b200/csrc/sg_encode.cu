// sg_encode.cu -- kernel 1a: nt4 encoding, homopolymer compression, 2-bit packing.
//
// Replaces the encode part of the reference's per-read loop (reference
// syncmer.c:284-323; table semantics syncmer.c:47-64): a run of one unambiguous
// base becomes one hoco base with run length - 1 in ho_rl (saturating at 255,
// longer runs also go to the ho_l_rl side list, :301-304); an ambiguous
// character is never compressed, is stored as A with ho_rl 0 and its RAW index
// goes to n_nucl (:316-322).
//
// One CTA per read, walking the read in tiles of NT*16 raw bytes. Per tile:
//   classify   16 bytes per thread from one aligned 128-bit load, SIMD-in-register
//              for plain ACGT/acgt; anything else (N, IUPAC, U, bytes 0..3, tile
//              edges) takes an exact per-byte path
//   run starts 2-bit packed codes XOR their 1-base shift -> start mask
//   scan       block prefix sum of start counts = hoco index
//   stage      (code, raw start) per hoco base into shared memory
//   finalise   run length = next start - this start; 16 codes -> one 32-bit word,
//              16 N flags -> one 16-bit word, coalesced stores; the open last run
//              and an incomplete group of 16 are carried into the next tile
#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

__device__ __forceinline__ int base_code_slow(uint32_t ch)
{
    if (ch < 4) return (int) ch;          // the nt4 table maps raw 0..3 to themselves
    switch (ch) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
    }
    return 4;
}

// four ASCII bytes -> 8 bits of packed codes (first byte in bits 7:6); `bad`
// collects any byte that is not one of ACGTacgt
__device__ __forceinline__ uint32_t classify4(uint32_t w, uint32_t &bad)
{
    const uint32_t K = 0x01010101u;
    uint32_t a = w >> 1, b = w >> 2;
    uint32_t c1 = b & K, c0 = (a ^ b) & K;
    uint32_t t = c0 & c1, o = c0 | c1, n = c1 & ~c0;
    uint32_t expect = 0x41414141u + t * 17u + o * 2u + n * 4u;   // 'A','C','G','T' per byte
    bad |= (w & 0xDFDFDFDFu) ^ expect;
    return ((c0 + 2u * c1) * 0x40100401u) >> 24;
}

template <int NT>
__global__ void __launch_bounds__(NT) encode_kernel(EncodeArgs A)
{
    constexpr int TILE = NT * 16;
    constexpr int NW = NT / 32;
    __shared__ __align__(16) uint8_t s_code[TILE + 48];
    __shared__ __align__(16) uint32_t s_pos[TILE + 48];
    __shared__ uint32_t s_scan[NW + 1];
    __shared__ uint32_t s_namb;

    const uint64_t r = blockIdx.x;
    const int tid = threadIdx.x;
    const uint64_t raw0 = A.off[r], raw1 = A.off[r + 1];
    const uint32_t len = (uint32_t) (raw1 - raw0);
    const uint64_t hb = A.hoff[r];                      // capacity offset, multiple of 64
    const uint64_t a0 = raw0 & ~15ull;
    const uint32_t ntiles = (uint32_t) ((raw1 - a0 + TILE - 1) / TILE);
    uint32_t *hs32 = reinterpret_cast<uint32_t *>(A.hoco_s + hb / 4);
    uint16_t *nb16 = reinterpret_cast<uint16_t *>(A.nbits + hb / 8);
    uint8_t *rl8 = A.ho_rl + hb;
    const uint32_t sid = (uint32_t) r;

    uint32_t n_stage = 0;        // staged entries carried from the previous tile (uniform)
    uint32_t g_done = 0;         // hoco entries already written (uniform, multiple of 16)
    if (tid == 0) s_namb = 0;

    for (uint32_t t = 0; t < ntiles; ++t) {
        const uint64_t g = a0 + (uint64_t) t * TILE + (uint64_t) tid * 16;   // global byte index of my chunk
        uint32_t P = 0, NM = 0, VM = 0;    // packed codes; ambiguous / void masks (bit 2*(15-i) for byte i)
        bool fast = false;
        if (g < raw1 && g + 16 > raw0) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(A.bases + g));
            uint32_t bad = 0;
            uint32_t p0 = classify4(q.x, bad), p1 = classify4(q.y, bad), p2 = classify4(q.z, bad), p3 = classify4(q.w, bad);
            if (bad == 0 && g >= raw0 && g + 16 <= raw1) {
                P = p0 << 24 | p1 << 16 | p2 << 8 | p3;
                fast = true;
            } else {
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t bit = 1u << (2 * (15 - i));
                    if (g + i < raw0 || g + i >= raw1) { VM |= bit; continue; }
                    int c = base_code_slow((w[i >> 2] >> (8 * (i & 3))) & 0xFFu);
                    if (c == 4) NM |= bit; else P |= (uint32_t) c << (2 * (15 - i));
                }
            }
        } else {
            VM = 0x55555555u;
        }
        // code of the byte in front of my chunk: 0..3, 4 = ambiguous, 5 = outside the read
        int pc = 5;
        if (g > raw0 && g <= raw1) pc = base_code_slow(__ldg(A.bases + g - 1));
        uint32_t M;
        {
            uint32_t D = P ^ ((P >> 2) | ((uint32_t) (pc & 3) << 30));
            M = (D | (D >> 1)) & 0x55555555u;
            if (pc >= 4) M |= 0x40000000u;
            if (!fast) M = (M | NM | (NM >> 2) | (VM >> 2)) & ~VM;
        }
        const uint32_t cnt = __popc(M);
        uint32_t tot;
        const uint32_t ex = BlockScanU32::run<NW>(cnt, s_scan, &tot);

        // stage (code, raw start) of every run that starts in my chunk: 16 predicated steps, no loop
        {
            uint32_t hl = n_stage + ex;
            const uint32_t rel = (uint32_t) (g - raw0);        // wraps for bytes before the read; those are void
            if (fast) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (M & (1u << (2 * (15 - i)))) {
                        s_code[hl] = (uint8_t) ((P >> (2 * (15 - i))) & 3u);
                        s_pos[hl] = rel + (uint32_t) i;
                        ++hl;
                    }
                }
            } else {
                uint32_t m = M;
                while (m) {
                    const int b = 31 - __clz(m);
                    m &= ~(1u << b);
                    s_code[hl] = (uint8_t) (((P >> b) & 3u) | ((NM >> b) & 1u) << 2);
                    s_pos[hl] = rel + (uint32_t) (15 - (b >> 1));
                    ++hl;
                }
            }
        }
        const bool last = (t + 1 == ntiles);
        const uint32_t n_avail = n_stage + tot;
        if (last && tid == 0) s_pos[n_avail] = len;
        __syncthreads();
        const uint32_t closed = last ? n_avail : (n_avail ? n_avail - 1 : 0);
        const uint32_t fin = last ? closed : (closed & ~15u);

        // finalise groups of 16 hoco bases: run lengths -> 16 bytes, codes -> one big-endian word,
        // ambiguity flags -> 16 bits; everything from vector loads of the staging arrays
        const uint32_t ngrp = (fin + 15) >> 4;
        for (uint32_t gi = tid; gi < ngrp; gi += NT) {
            const uint32_t e0 = gi * 16;
            const uint32_t nval = min(16u, fin - e0);             // entries of this group that exist
            uint32_t ps[17];
            {
                const uint4 *q4 = reinterpret_cast<const uint4 *>(s_pos + e0);
                const uint4 a = q4[0], b4 = q4[1], c4 = q4[2], d4 = q4[3];
                ps[0] = a.x; ps[1] = a.y; ps[2] = a.z; ps[3] = a.w; ps[4] = b4.x; ps[5] = b4.y; ps[6] = b4.z; ps[7] = b4.w;
                ps[8] = c4.x; ps[9] = c4.y; ps[10] = c4.z; ps[11] = c4.w; ps[12] = d4.x; ps[13] = d4.y; ps[14] = d4.z; ps[15] = d4.w;
                ps[16] = s_pos[e0 + 16];
            }
            uint32_t rl[16], big = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                rl[j] = (uint32_t) j < nval ? ps[j + 1] - ps[j] - 1u : 0u;     // run length - 1
                big |= rl[j];
            }
            const uint4 cw = *reinterpret_cast<const uint4 *>(s_code + e0);
            const uint32_t w[4] = {cw.x, cw.y, cw.z, cw.w};
            uint32_t word = 0, nbw = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t x = w[j];
                if ((uint32_t) (4 * j) >= nval) x = 0;
                else if (nval - 4 * j < 4) x &= (1u << (8 * (nval - 4 * j))) - 1u;
                word |= (((x & 0x03030303u) * 0x40100401u) >> 24) << (24 - 8 * j);
                nbw |= ((((x >> 2) & 0x01010101u) * 0x00204081u >> 21) & 0xFu) << (4 * j);   // gather bit 2 of 4 bytes
            }
            if (big >= 255u) {
                // a run longer than 256 saturates ho_rl and goes to the side list (syncmer.c:301-304)
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (rl[j] >= 255u) {
                        unsigned long long o = atomicAdd(A.lrl_count, 1ull);
                        if (o < A.lrl_cap) { A.lrl_sid[o] = sid; A.lrl_idx[o] = g_done + e0 + j; A.lrl_val[o] = rl[j]; }
                        rl[j] = 255u;
                    }
                }
            }
            if (nbw) {
                for (uint32_t mm = nbw; mm; mm &= mm - 1) {
                    const int j = __ffs(mm) - 1;
                    unsigned long long o = atomicAdd(A.amb_count, 1ull);
                    if (o < A.amb_cap) { A.amb_sid[o] = sid; A.amb_pos[o] = s_pos[e0 + j]; }
                }
                atomicAdd(&s_namb, (uint32_t) __popc(nbw));
            }
            uint4 out;
            out.x = rl[0] | rl[1] << 8 | rl[2] << 16 | rl[3] << 24;
            out.y = rl[4] | rl[5] << 8 | rl[6] << 16 | rl[7] << 24;
            out.z = rl[8] | rl[9] << 8 | rl[10] << 16 | rl[11] << 24;
            out.w = rl[12] | rl[13] << 8 | rl[14] << 16 | rl[15] << 24;
            *reinterpret_cast<uint4 *>(rl8 + g_done + e0) = out;      // capacity is padded to 64: a full store always fits
            hs32[(g_done >> 4) + gi] = bswap32(word);
            nb16[(g_done >> 4) + gi] = (uint16_t) nbw;
        }
        // carry the tail (open run and incomplete group) to the front
        const uint32_t n_carry = n_avail - fin;
        uint8_t cc = 0; uint32_t cp = 0;
        if ((uint32_t) tid < n_carry) { cc = s_code[fin + tid]; cp = s_pos[fin + tid]; }
        __syncthreads();
        if ((uint32_t) tid < n_carry) { s_code[tid] = cc; s_pos[tid] = cp; }
        __syncthreads();
        n_stage = n_carry;
        g_done += fin;
    }
    if (tid == 0) {
        A.hoco_l[r] = g_done;
        A.n_amb[r] = s_namb;
    }
}

int launch_encode(const EncodeArgs &A, uint64_t n_reads, cudaStream_t st)
{
    if (n_reads == 0) return 0;
    encode_kernel<256><<<(unsigned) n_reads, 256, 0, st>>>(A);
    return 1;
}

} // namespace sg
