// sg_encode.cu -- kernel 1a: nt4 encoding, homopolymer compression, 2-bit packing.
//
// Replaces the encode part of the reference's per-read loop (reference
// syncmer.c:284-323; table semantics syncmer.c:47-64): a run of one unambiguous
// base becomes one hoco base with run length - 1 in ho_rl (saturating at 255,
// longer runs also go to the ho_l_rl side list, :301-304); an ambiguous
// character is never compressed, is stored as A with ho_rl 0 and its RAW index
// goes to n_nucl (:316-322).
//
// One WARP per read (reads are handed out through an atomic counter, the grid is
// persistent), no block-wide barriers. The warp walks the read in tiles of 1024
// raw bytes. A tile travels global -> shared memory as ONE bulk copy of the TMA
// unit (cp.async.bulk, issued by lane 0 two tiles ahead, completion counted in
// bytes on an mbarrier of the stage); lane l then owns 32 consecutive bytes of the
// tile (two 128-bit shared-memory loads). Per tile:
//   classify   four bytes at a time in registers: codes from bits 1-2 of the ASCII
//              byte, validity by re-deriving the three bits that separate ACGT/acgt
//              from everything else; anything else (N, IUPAC, U, bytes 0..3) sends
//              the lane through an exact per-byte path
//   run starts 2-bit packed codes XOR their one-base shift -> start mask; warp
//              prefix sum of the start counts = hoco index of the lane's first run
//   stage      every start stores ONE 16-bit entry {code of the new run, length - 1
//              of the run it closes} into a warp-private shared array: 32 unrolled
//              steps of predicate / length / byte-merge / pointer bump
//   emit       finished groups of 16 entries leave as one 128-bit store of run
//              lengths, one 32-bit word of codes and 16 N flags; the open last run
//              and an incomplete group stay staged for the next tile
#include <algorithm>
#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

constexpr int ENC_WARPS = 8;
constexpr int ENC_TILE = 1024;                 // raw bytes per warp and tile
constexpr int ENC_SLOTS = ENC_TILE + 64;       // staged 16-bit entries per warp (tile + carry + slack)
constexpr int ENC_STAGES = 2;                  // raw tiles in flight per warp

// ---- TMA bulk copy + mbarrier (PTX; SASS: UBLKCP / SYNCS) ----
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(arrivals) : "memory");
}
// one arrival that also announces `bytes` of asynchronous traffic: the phase completes when they have landed
__device__ __forceinline__ void mbar_arrive_expect(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nSG_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra SG_DONE;\nbra SG_WAIT;\nSG_DONE:\n}"
            :: "r"(bar), "r"(parity) : "memory");
}

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

// four ASCII bytes -> their codes 0..3 in the same byte lanes; `bad` collects any byte that is not
// one of ACGTacgt. A=0x41 C=0x43 G=0x47 T=0x54: bits 2:1 give the code (Gray-decoded), bit 5 is the
// case, and the other bits must read 0 1 . t 0 . . !t with t = "is T".
__device__ __forceinline__ uint32_t classify4(uint32_t w, uint32_t &bad)
{
    const uint32_t K = 0x01010101u;
    const uint32_t c1 = (w >> 2) & K, c0 = ((w >> 1) ^ (w >> 2)) & K;
    const uint32_t t17 = (c0 & c1) * 17u;                 // T: bit 4 set, bit 0 clear
    bad |= ((w ^ t17) & 0xD9D9D9D9u) ^ 0x41414141u;
    return c0 + 2u * c1;
}
// validity word of four bytes (zero byte = one of ACGTacgt), not accumulated
__device__ __forceinline__ uint32_t invalid4(uint32_t w)
{
    const uint32_t K = 0x01010101u;
    const uint32_t t17 = ((w >> 2) & ((w >> 1) ^ (w >> 2)) & K) * 17u;
    return ((w ^ t17) & 0xD9D9D9D9u) ^ 0x41414141u;
}
// codes in byte lanes -> 8 bits, first byte in bits 7:6
__device__ __forceinline__ uint32_t pack4(uint32_t cc) { return ((cc & 0x03030303u) * 0x40100401u) >> 24; }

__global__ void __launch_bounds__(32 * ENC_WARPS) encode_kernel(EncodeArgs A)
{
    __shared__ __align__(16) uint16_t s_ent_all[ENC_WARPS][ENC_SLOTS];
    __shared__ __align__(128) uint8_t s_raw_all[ENC_WARPS][ENC_STAGES][ENC_TILE];   // raw tiles, filled by the TMA unit
    __shared__ __align__(8) uint64_t s_bar_all[ENC_WARPS][ENC_STAGES];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint16_t *ent = s_ent_all[wid];
    uint32_t *ent32 = reinterpret_cast<uint32_t *>(ent);
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t raw_s = smem_addr(s_raw_all[wid][0]), bar_s = smem_addr(&s_bar_all[wid][0]);
    if (lane == 0) {
        for (int i = 0; i < ENC_STAGES; ++i) mbar_init(bar_s + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    uint32_t n_used = 0;             // tiles this warp has consumed so far: stage = n_used % STAGES, parity = (n_used / STAGES) & 1

    for (;;) {
        unsigned int r32 = 0;
        if (lane == 0) r32 = atomicAdd(A.work, 1u);
        r32 = __shfl_sync(SG_FULL, r32, 0);
        if ((uint64_t) r32 >= A.n_reads) break;
        const uint64_t r = r32;
        const uint32_t sid = r32;
        const uint64_t raw0 = A.off[r], raw1 = A.off[r + 1];
        const int len = (int) (raw1 - raw0);
        const uint64_t hb = A.hoff[r];                      // capacity offset, multiple of 64
        const uint64_t a0 = raw0 & ~15ull;
        const int ntiles = (int) ((raw1 - a0 + ENC_TILE - 1) / ENC_TILE);
        uint32_t *hs32 = reinterpret_cast<uint32_t *>(A.hoco_s + hb / 4);
        uint16_t *nb16 = reinterpret_cast<uint16_t *>(A.nbits + hb / 8);
        uint4 *rl128 = reinterpret_cast<uint4 *>(A.ho_rl + hb);

        uint32_t n_stage = 0;        // staged entries carried from the previous tile, the open run last (uniform)
        uint32_t g_done = 0;         // hoco entries already written (uniform, multiple of 16)
        int lastpos = -1;            // read position of the latest run start seen so far (uniform)
        uint32_t last_cc = 5;        // code of the last byte of the previous tile: 0..3, 4 ambiguous, 5 outside (uniform)
        uint32_t n_amb = 0;          // ambiguous characters seen by this lane
        bool any_n = false;          // the staged entries may carry N flags (uniform)

        // tile t of this read into the stage it will be consumed from (lane 0 only). The copy ends at the 16-byte
        // boundary after the read's last base (the input buffer is readable up to there, include/syncgpu.h).
        // (One lane issues it, so every instruction here costs the warp a whole issue slot: the sizes are 32-bit and come
        // from one subtraction per read.)
        const uint32_t span = (uint32_t) (((raw1 + 15ull) & ~15ull) - a0);     // bytes from the first tile's start to the end of the last copy
        const uint8_t *src0 = A.bases + a0;
        auto issue_tile = [&](int t, uint32_t use, bool dirty) {
            const uint32_t at = (uint32_t) t * ENC_TILE, bytes = min((uint32_t) ENC_TILE, span - at);
            const uint32_t st = use % ENC_STAGES;
            // the slow path writes codes over the stage with ordinary stores: they are ordered before the unit's writes
            if (dirty) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect(bar_s + 8 * st, bytes);
            bulk_load(raw_s + st * ENC_TILE, src0 + at, bytes, bar_s + 8 * st);
        };
        // a run of 256 or more saturates ho_rl and goes to the side list (syncmer.c:301-304)
        auto side_list = [&](uint32_t idx, uint32_t rl1) {
            unsigned long long o = atomicAdd(A.lrl_count, 1ull);
            if (o < A.lrl_cap) { A.lrl_sid[o] = sid; A.lrl_idx[o] = idx; A.lrl_val[o] = rl1; }
        };

        if (lane == 0)
            for (int t = 0; t < ENC_STAGES && t < ntiles; ++t) issue_tile(t, n_used + t, true);
        for (int t = 0; t < ntiles; ++t, ++n_used) {
            const uint32_t st = n_used % ENC_STAGES;
            mbar_wait(bar_s + 8 * st, (n_used / ENC_STAGES) & 1u);
            // bytes the copy did not bring (beyond the read's end) read as zero, as do chunks outside the read
            const uint64_t g = a0 + (uint64_t) t * ENC_TILE + (uint64_t) lane * 32;
            const bool in0 = g < raw1 && g + 32 > raw0, in1 = in0 && g + 16 < raw1;
            const uint4 *tp = reinterpret_cast<const uint4 *>(s_raw_all[wid][st]) + 2 * lane;
            const uint4 zero4 = make_uint4(0, 0, 0, 0);
            const uint4 q0 = in0 ? tp[0] : zero4, q1 = in1 ? tp[1] : zero4;
            const int rel = (int) (int64_t) (a0 + (uint64_t) t * ENC_TILE + (uint64_t) lane * 32 - raw0);   // read position of my first byte
            const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            uint32_t cc[8];
            uint32_t bad = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) cc[j] = classify4(w[j], bad);
            const bool inside = rel >= 0 && rel + 32 <= len;
            // valid positions of my chunk: [vlo, vhi)
            const int vlo = min(32, max(0, -rel)), vhi = max(0, min(32, len - rel));
            uint32_t NM0 = 0, NM1 = 0;                     // ambiguous positions, bit 2*(15-i) like the start masks
            if (!inside) {
                // chunk at the head or tail of the read: only the bytes inside the read decide whether it is plain ACGT
                bad = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int lo = max(vlo - 4 * j, 0), hi = min(vhi - 4 * j, 4);       // valid bytes [lo, hi) of word j
                    const uint32_t bm = hi > lo ? (0xffffffffu >> (8 * (4 - hi + lo))) << (8 * lo) : 0u;
                    bad |= invalid4(w[j]) & bm;
                }
            }
            const bool slow = bad != 0;
            if (bad) {
                // exact per-byte codes (also fixes U, raw 0..3) for the valid bytes; everything else becomes 0. The words
                // are taken from the staged tile and the codes written back over them, so that no register array is
                // indexed at run time (it would move to local memory); my 32 bytes of the stage are mine alone.
                bad = 0;
                uint32_t *mine = reinterpret_cast<uint32_t *>(s_raw_all[wid][st]) + 8 * lane;
#pragma unroll 1
                for (int j = 0; j < 8; ++j) {
                    uint32_t cw = 0;
                    const uint32_t wj = mine[j];
                    for (int b = 0; b < 4; ++b) {
                        const int i = 4 * j + b;
                        if (i < vlo || i >= vhi) continue;
                        const int c = base_code_slow((wj >> (8 * b)) & 0xffu);
                        if (c == 4) {
                            cw |= 4u << (8 * b);               // N flag rides in bit 2 of the code byte; the code itself is A
                            if (i < 16) NM0 |= 1u << (2 * (15 - i)); else NM1 |= 1u << (2 * (31 - i));
                            unsigned long long o = atomicAdd(A.amb_count, 1ull);
                            if (o < A.amb_cap) { A.amb_sid[o] = sid; A.amb_pos[o] = (uint32_t) (rel + i); }
                            ++n_amb;
                        } else cw |= (uint32_t) c << (8 * b);
                    }
                    mine[j] = cw;
                }
                const uint4 c0 = reinterpret_cast<const uint4 *>(mine)[0], c1 = reinterpret_cast<const uint4 *>(mine)[1];
                cc[0] = c0.x; cc[1] = c0.y; cc[2] = c0.z; cc[3] = c0.w; cc[4] = c1.x; cc[5] = c1.y; cc[6] = c1.z; cc[7] = c1.w;
            }
            // every lane is done with the staged bytes (the vote is the warp's rendezvous): refill the stage
            const bool dirty = __any_sync(SG_FULL, slow);
            if (dirty && __any_sync(SG_FULL, (NM0 | NM1) != 0)) any_n = true;
            if (lane == 0 && t + ENC_STAGES < ntiles) issue_tile(t + ENC_STAGES, n_used + ENC_STAGES, dirty);

            // packed codes: positions 0..15 in P0, 16..31 in P1, first position in bits 31:30
            const uint32_t P0 = pack4(cc[0]) << 24 | pack4(cc[1]) << 16 | pack4(cc[2]) << 8 | pack4(cc[3]);
            const uint32_t P1 = pack4(cc[4]) << 24 | pack4(cc[5]) << 16 | pack4(cc[6]) << 8 | pack4(cc[7]);
            // code of the byte in front of my chunk: 0..3, 4 = ambiguous, 5 = outside the read
            uint32_t mylastc = (NM1 & 1u) ? 4u : (P1 & 3u);
            if (rel + 31 < 0 || rel + 31 >= len) mylastc = 5u;
            uint32_t pc = __shfl_up_sync(SG_FULL, mylastc, 1);
            if (lane == 0) pc = last_cc;
            last_cc = __shfl_sync(SG_FULL, mylastc, 31);
            uint32_t M0, M1;
            {
                const uint32_t D0 = P0 ^ ((P0 >> 2) | (pc & 3u) << 30), D1 = P1 ^ ((P1 >> 2) | P0 << 30);
                M0 = (D0 | (D0 >> 1)) & 0x55555555u;
                M1 = (D1 | (D1 >> 1)) & 0x55555555u;
                if (pc >= 4u) M0 |= 0x40000000u;
                if (!inside || (NM0 | NM1)) {
                    // void positions never start a run, the first valid position always does; an ambiguous
                    // character is its own run and so is whatever follows it
                    const uint64_t all = 0x5555555555555555ull;
                    const uint64_t keep = (vhi > vlo) ? ((all >> (2 * vlo)) & ~(vhi < 32 ? all >> (2 * vhi) : 0ull)) : 0ull;
                    const uint64_t NM = (uint64_t) NM0 << 32 | NM1;
                    uint64_t M = (uint64_t) M0 << 32 | M1;
                    M |= NM | (NM >> 2);
                    if (vlo > 0 && vlo < 32) M |= 1ull << (2 * (31 - vlo));
                    M &= keep;
                    M0 = (uint32_t) (M >> 32); M1 = (uint32_t) M;
                }
            }
            const uint32_t cnt = __popc(M0) + __popc(M1);
            // read position of my last start, and of the latest start in front of my chunk
            int mylast = -1;
            if (M1) mylast = rel + 31 - ((__ffs(M1) - 1) >> 1);
            else if (M0) mylast = rel + 15 - ((__ffs(M0) - 1) >> 1);
            const uint32_t has = __ballot_sync(SG_FULL, cnt != 0);
            const uint32_t below = has & lt_mask;
            int prev = __shfl_sync(SG_FULL, mylast, below ? 31 - __clz(below) : 0);
            if (!below) prev = lastpos;
            if (has) lastpos = __shfl_sync(SG_FULL, mylast, 31 - __clz(has));
            // hoco index of my first start
            uint32_t inc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t a = __shfl_up_sync(SG_FULL, inc, d); if (lane >= d) inc += a; }
            const uint32_t tot = __shfl_sync(SG_FULL, inc, 31);
            uint32_t e = n_stage + inc - cnt;              // slot of my first start

            if (cnt) {
                if (rel + 31 - prev < 255) {
                    // fast: no run that closes in my chunk can reach 256. nb = -(position of the latest start
                    // relative to my chunk), so a start at i closes a run of length - 1 = i - 1 + nb.
                    uint16_t *pe = ent + e;
                    int nb = rel - prev;
#define SG_ENC_STEP(I, MW, CW) { \
                        const bool st = (MW >> (2 * (15 - ((I) & 15)))) & 1u; \
                        const uint32_t v = __byte_perm((uint32_t) (nb + (I) - 1), CW, 0x0040 | ((4 + ((I) & 3)) << 4)); \
                        if (st) { *pe = (uint16_t) v; ++pe; nb = -(I); } }
                    SG_ENC_STEP(0, M0, cc[0]) SG_ENC_STEP(1, M0, cc[0]) SG_ENC_STEP(2, M0, cc[0]) SG_ENC_STEP(3, M0, cc[0])
                    SG_ENC_STEP(4, M0, cc[1]) SG_ENC_STEP(5, M0, cc[1]) SG_ENC_STEP(6, M0, cc[1]) SG_ENC_STEP(7, M0, cc[1])
                    SG_ENC_STEP(8, M0, cc[2]) SG_ENC_STEP(9, M0, cc[2]) SG_ENC_STEP(10, M0, cc[2]) SG_ENC_STEP(11, M0, cc[2])
                    SG_ENC_STEP(12, M0, cc[3]) SG_ENC_STEP(13, M0, cc[3]) SG_ENC_STEP(14, M0, cc[3]) SG_ENC_STEP(15, M0, cc[3])
                    SG_ENC_STEP(16, M1, cc[4]) SG_ENC_STEP(17, M1, cc[4]) SG_ENC_STEP(18, M1, cc[4]) SG_ENC_STEP(19, M1, cc[4])
                    SG_ENC_STEP(20, M1, cc[5]) SG_ENC_STEP(21, M1, cc[5]) SG_ENC_STEP(22, M1, cc[5]) SG_ENC_STEP(23, M1, cc[5])
                    SG_ENC_STEP(24, M1, cc[6]) SG_ENC_STEP(25, M1, cc[6]) SG_ENC_STEP(26, M1, cc[6]) SG_ENC_STEP(27, M1, cc[6])
                    SG_ENC_STEP(28, M1, cc[7]) SG_ENC_STEP(29, M1, cc[7]) SG_ENC_STEP(30, M1, cc[7]) SG_ENC_STEP(31, M1, cc[7])
#undef SG_ENC_STEP
                } else {
                    // a long run may close here: exact lengths, saturation and the side list
                    int pv = prev;
                    for (int i = 0; i < 32; ++i) {
                        const uint32_t MW = i < 16 ? M0 : M1;
                        if (!((MW >> (2 * (15 - (i & 15)))) & 1u)) continue;
                        uint32_t rl1 = (uint32_t) (rel + i - pv - 1);
                        if (rl1 >= 255u) { if (g_done + e > 0) side_list(g_done + e - 1, rl1); rl1 = 255u; }
                        // code byte of position i from the packed words (no run-time index into cc[]): code | N flag << 2
                        const int sh = 2 * (15 - (i & 15));
                        const uint32_t cb = (((i < 16 ? P0 : P1) >> sh) & 3u) | ((((i < 16 ? NM0 : NM1) >> sh) & 1u) << 2);
                        ent[e] = (uint16_t) (rl1 | cb << 8);
                        pv = rel + i;
                        ++e;
                    }
                }
            }
            const bool last = (t + 1 == ntiles);
            const uint32_t n_avail = n_stage + tot;
            __syncwarp();
            if (last && lane == 0 && n_avail) {
                // the end of the read closes the last run
                uint32_t rl1 = (uint32_t) (len - lastpos - 1);
                if (rl1 >= 255u) { side_list(g_done + n_avail - 1, rl1); rl1 = 255u; }
                ent[n_avail] = (uint16_t) rl1;
            }
            __syncwarp();
            const uint32_t fin = last ? n_avail : (n_avail ? ((n_avail - 1) & ~15u) : 0);

            // emit the finished groups of 16: slot e = {low byte: length - 1 of entry e-1, high byte: code of entry e}
            const uint32_t ngrp = (fin + 15) >> 4;
            for (uint32_t gi = lane; gi < ngrp; gi += 32) {
                const uint4 A0 = *reinterpret_cast<const uint4 *>(ent32 + 8 * gi), A1 = *reinterpret_cast<const uint4 *>(ent32 + 8 * gi + 4);
                const uint32_t W8 = ent32[8 * gi + 8];
                const uint32_t W[9] = {A0.x, A0.y, A0.z, A0.w, A1.x, A1.y, A1.z, A1.w, W8};
                uint32_t R[4], C[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    C[m] = __byte_perm(W[2 * m], W[2 * m + 1], 0x7531);                    // codes of entries 4m .. 4m+3
                    const uint32_t x = __byte_perm(W[2 * m], W[2 * m + 1], 0x0642);        // lengths of 4m, 4m+1, 4m+2
                    R[m] = __byte_perm(x, W[2 * m + 2], 0x4210);                           // and 4m+3 from the next word
                }
                uint32_t code = pack4(C[0]) | pack4(C[1]) << 8 | pack4(C[2]) << 16 | pack4(C[3]) << 24;
                uint32_t nflag = 0;
                if (any_n) {
#pragma unroll
                    for (int m = 0; m < 4; ++m) nflag |= ((((C[m] >> 2) & 0x01010101u) * 0x10204080u) >> 28) << (4 * m);   // bytes 0..3 -> bits 0..3
                }
                const uint32_t left = fin - 16 * gi;       // entries of this group that exist (>= 16 except in the last group of a read)
                if (left < 16) {
                    // keep the unused tail of the last word clean: bases beyond hoco_l read as zero
                    // (entry i sits in bits 7:6 >> 2 (i & 3) of byte i >> 2: whole bytes for the first left / 4, the top
                    // 2 (left & 3) bits of the next)
                    const uint32_t fullb = left >> 2, part = left & 3u;
                    const uint32_t keepc = ((1u << (8 * fullb)) - 1u) | (((0xff00u >> (2 * part)) & 0xffu) << (8 * fullb));
                    code &= keepc;
                    nflag &= (1u << left) - 1u;
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int nb_ = min(4, max(0, (int) left - 4 * m));        // run lengths of this word that exist
                        R[m] &= nb_ == 4 ? 0xffffffffu : (1u << (8 * nb_)) - 1u;
                    }
                }
                const uint32_t go = (g_done >> 4) + gi;
                hs32[go] = code;
                nb16[go] = (uint16_t) nflag;
                rl128[go] = make_uint4(R[0], R[1], R[2], R[3]);
            }
            // what stays staged: the open run and an incomplete group (at most 16 entries)
            const uint32_t n_carry = n_avail - fin;
            uint16_t cv = 0;
            if ((uint32_t) lane < n_carry) cv = ent[fin + lane];
            __syncwarp();
            if ((uint32_t) lane < n_carry) ent[lane] = cv;
            n_stage = n_carry;
            g_done += fin;
            __syncwarp();
        }
        n_amb = __reduce_add_sync(SG_FULL, n_amb);
        if (lane == 0) {
            A.hoco_l[r] = g_done;
            A.n_amb[r] = n_amb;
        }
    }
}

int launch_encode(const EncodeArgs &A, cudaStream_t st)
{
    if (A.n_reads == 0) return 0;
    static int ctas_per_sm = 0, n_sm = 0;
    if (!ctas_per_sm) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return -1;
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, encode_kernel, 32 * ENC_WARPS, 0) != cudaSuccess) return -1;
        if (ctas_per_sm < 1) return -1;
    }
    const uint64_t want = (A.n_reads + ENC_WARPS - 1) / ENC_WARPS;
    const unsigned grid = (unsigned) std::min<uint64_t>(want, (uint64_t) n_sm * ctas_per_sm);
    encode_kernel<<<grid, 32 * ENC_WARPS, 0, st>>>(A);
    return 1;
}

} // namespace sg
