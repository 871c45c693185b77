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
//   scan       block prefix sum of start counts = hoco index, block prefix maximum
//              of "latest start" = where the run before my chunk began
//   close      every start closes the run before it (length = distance between
//              starts) with one byte store into shared memory and shifts its 2-bit
//              code into a register; the thread's codes are then OR-ed into the
//              shared big-endian word stream with at most two shared atomics
//   flush      finished groups of 16 entries leave as one 32-bit code word, one
//              128-bit store of run lengths and 16 N flags; the open last run and an
//              incomplete group stay staged for the next tile
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
    constexpr int NWORD = TILE / 16 + 4;               // staged 16-entry groups (tile + carry)
    __shared__ __align__(16) uint8_t s_rl_arr[TILE + 64];   // run length - 1 per staged hoco entry; [-16, 0) is scratch
    __shared__ uint32_t s_w[NWORD];                    // 2-bit codes, 16 per word, first entry in bits 31:30
    __shared__ uint32_t s_nbw[NWORD / 2 + 2];          // ambiguity flags, entry e -> bit e & 31 of word e >> 5
    __shared__ uint32_t s_cnt[NW];
    __shared__ int s_last[NW];
    __shared__ uint32_t s_namb;
    uint8_t *s_rl = s_rl_arr + 16;

    const uint64_t r = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t raw0 = A.off[r], raw1 = A.off[r + 1];
    const uint32_t len = (uint32_t) (raw1 - raw0);
    const uint64_t hb = A.hoff[r];                      // capacity offset, multiple of 64
    const uint64_t a0 = raw0 & ~15ull;
    const uint32_t ntiles = (uint32_t) ((raw1 - a0 + TILE - 1) / TILE);
    uint32_t *hs32 = reinterpret_cast<uint32_t *>(A.hoco_s + hb / 4);
    uint16_t *nb16 = reinterpret_cast<uint16_t *>(A.nbits + hb / 8);
    uint8_t *rl8 = A.ho_rl + hb;
    const uint32_t sid = (uint32_t) r;

    uint32_t n_stage = 0;        // staged entries carried from the previous tile, the open run last (uniform)
    uint32_t g_done = 0;         // hoco entries already written (uniform, multiple of 16)
    int carry_last = -1;         // raw position of the latest run start seen so far (uniform)
    for (int i = tid; i < NWORD; i += NT) s_w[i] = 0;
    for (int i = tid; i < NWORD / 2 + 2; i += NT) s_nbw[i] = 0;
    if (tid == 0) s_namb = 0;
    __syncthreads();

    // a run longer than 256 saturates ho_rl and goes to the side list (syncmer.c:301-304)
    auto close_run = [&](uint32_t e, uint32_t rl1) {
        if (rl1 >= 255u) {
            unsigned long long o = atomicAdd(A.lrl_count, 1ull);
            if (o < A.lrl_cap) { A.lrl_sid[o] = sid; A.lrl_idx[o] = g_done + e; A.lrl_val[o] = rl1; }
            rl1 = 255u;
        }
        s_rl[(int) e] = (uint8_t) rl1;
    };

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
        const int rel = (int) (uint32_t) (g - raw0);   // wraps for bytes before the read; those are void
        const uint32_t cnt = __popc(M);
        const int mylast = M ? rel + 15 - ((__ffs(M) - 1) >> 1) : -1;

        // block scan: hoco index of my first start, and the raw position of the latest start before my chunk
        uint32_t inc = cnt;
        int lmax = mylast;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t a = __shfl_up_sync(SG_FULL, inc, d);
            const int bq = __shfl_up_sync(SG_FULL, lmax, d);
            if (lane >= d) { inc += a; lmax = max(lmax, bq); }
        }
        if (lane == 31) { s_cnt[wid] = inc; s_last[wid] = lmax; }
        __syncthreads();
        uint32_t ex = inc - cnt, tot = 0;
        int prev = __shfl_up_sync(SG_FULL, lmax, 1), blast = carry_last;
        if (lane == 0) prev = -1;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const uint32_t c = s_cnt[w];
            const int l = s_last[w];
            if (w < wid) { ex += c; prev = max(prev, l); }
            tot += c; blast = max(blast, l);
        }
        prev = max(prev, carry_last);

        // every run start closes the run before it (run length = distance between starts) and
        // appends its code; 16 predicated steps, everything in registers but the byte store
        {
            uint32_t e = n_stage + ex, CP = 0;
            if (fast) {
                // branch-free steps: selects and one predicated byte store each; runs of 256+ are rare and
                // are only noticed here (big), then pushed to the side list by a second pass
                uint32_t big = 0;
                const int prev0 = prev;
                const uint32_t e0 = e;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const bool st = (M >> (2 * (15 - i))) & 1u;
                    const uint32_t rl1 = (uint32_t) (rel + i - prev - 1);
                    if (st) s_rl[(int) e - 1] = (uint8_t) min(rl1, 255u);
                    big |= st ? rl1 : 0u;
                    prev = st ? rel + i : prev;
                    CP = st ? CP * 4u + ((P >> (2 * (15 - i))) & 3u) : CP;
                    e += st;
                }
                if (big >= 255u) {
                    int pv = prev0;
                    uint32_t ee = e0, m = M;
                    while (m) {
                        const int b = 31 - __clz(m);
                        m &= ~(1u << b);
                        const int i = 15 - (b >> 1);
                        const uint32_t rl1 = (uint32_t) (rel + i - pv - 1);
                        if (rl1 >= 255u) close_run(ee - 1, rl1);
                        pv = rel + i;
                        ++ee;
                    }
                }
            } else {
                uint32_t m = M;
                while (m) {
                    const int b = 31 - __clz(m);
                    m &= ~(1u << b);
                    const int i = 15 - (b >> 1);
                    close_run(e - 1, (uint32_t) (rel + i - prev - 1));   // entry -1 of the read lands in scratch
                    prev = rel + i;
                    CP = CP * 4u + ((P >> b) & 3u);
                    if ((NM >> b) & 1u) {
                        atomicOr(&s_nbw[e >> 5], 1u << (e & 31));
                        unsigned long long o = atomicAdd(A.amb_count, 1ull);
                        if (o < A.amb_cap) { A.amb_sid[o] = sid; A.amb_pos[o] = (uint32_t) (rel + i); }
                        atomicAdd(&s_namb, 1u);
                    }
                    ++e;
                }
            }
            if (cnt) {
                // my cnt codes sit in the low 2*cnt bits of CP: drop them at bit 2*(n_stage+ex) of the big-endian stream
                const uint32_t bitpos = 2u * (n_stage + ex);
                const uint64_t v = ((uint64_t) CP << (64 - 2 * cnt)) >> (bitpos & 31u);
                atomicOr(&s_w[bitpos >> 5], (uint32_t) (v >> 32));
                if ((uint32_t) v) atomicOr(&s_w[(bitpos >> 5) + 1], (uint32_t) v);
            }
        }
        const bool last = (t + 1 == ntiles);
        const uint32_t n_avail = n_stage + tot;
        if (last && tid == 0 && n_avail) close_run(n_avail - 1, (uint32_t) ((int) len - blast - 1));
        __syncthreads();
        const uint32_t fin = last ? n_avail : (n_avail ? ((n_avail - 1) & ~15u) : 0);

        // write the finished groups of 16: one code word, 16 run-length bytes, 16 flags
        const uint32_t ngrp = (fin + 15) >> 4;
        for (uint32_t gi = tid; gi < ngrp; gi += NT) {
            hs32[(g_done >> 4) + gi] = bswap32(s_w[gi]);
            nb16[(g_done >> 4) + gi] = (uint16_t) (s_nbw[gi >> 1] >> (16 * (gi & 1)));
            *reinterpret_cast<uint4 *>(rl8 + g_done + 16 * gi) = *reinterpret_cast<const uint4 *>(s_rl + 16 * gi);
        }
        // what stays staged: the open run and an incomplete group (at most 16 entries, one word)
        const uint32_t n_carry = n_avail - fin;
        const uint32_t cw = s_w[fin >> 4], cn = (s_nbw[fin >> 5] >> (fin & 16u)) & 0xffffu;
        uint8_t cb = 0;
        if ((uint32_t) tid < n_carry) cb = s_rl[fin + tid];
        __syncthreads();
        if ((uint32_t) tid < n_carry) s_rl[tid] = cb;
        for (int i = tid; i < NWORD; i += NT) s_w[i] = (i == 0 && n_carry) ? cw : 0u;
        for (int i = tid; i < NWORD / 2 + 2; i += NT) s_nbw[i] = (i == 0 && n_carry) ? cn : 0u;
        n_stage = n_carry;
        g_done += fin;
        carry_last = blast;
        // the two barriers of the next tile's scan order these writes before anybody stages again
    }
    if (tid == 0) {
        A.hoco_l[r] = g_done;
        A.n_amb[r] = s_namb;
    }
}

int launch_encode(const EncodeArgs &A, uint64_t n_reads, cudaStream_t st)
{
    if (n_reads == 0) return 0;
    encode_kernel<128><<<(unsigned) n_reads, 128, 0, st>>>(A);
    return 1;
}

} // namespace sg
