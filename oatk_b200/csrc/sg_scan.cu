// sg_scan.cu -- kernel 1b: rolling s-mer hash, k-window minimum, closed-syncmer selection.
//
// Replaces the minimiser/emission part of the reference's per-read loop (reference
// syncmer.c:276-283, 307-394) with the stateless rules derived in DESIGN.md section 3:
//
//   m[p]     hash64 of the canonical s-mer ending at hoco position p, or NONE
//   mo(p)    min m[p-q+1 .. p-1],  e(p) = m[p-q],  q = k-s+1
//   CLOSE(p) m[p] valid, l[p] >= k, m[p] <= mo(p) and (m[p] <= e(p) or m[p] < mo(p) or m[p-q+1] == m[p])
//   OPEN(p)  e(p) valid, e(p) <= mo(p), l[p-1] >= k and (p == H or base p unambiguous)
//   start t emits iff CLOSE(t+k-1) xor OPEN(t+k)
//
// One WARP per read (reads are handed out through an atomic counter, the grid is persistent), no
// block-wide barriers. The warp walks the read in tiles of 32 chunks x 16 positions; lane l owns one
// chunk (one 32-bit word of packed bases).
//   1. hash the 16 positions of the chunk; only a 32-bit key that is monotone in the 62-bit hash is
//      kept (s = 31: hash >> 30 clamped below NONE, see sg_hash31.cuh; otherwise hash >> 32), in a
//      warp-private shared-memory ring of the last q + 512 positions laid out [16][chunks] so that
//      every access pattern is bank-conflict free
//   2. window minimum over whole chunks: chunk minima are prefix- and suffix-min scanned inside blocks
//      of B lanes (B = 32 for k = 1001) with shuffles; the minimum r0 over the n_full chunks in front
//      of a chunk is then suffix(first block) , totals of the blocks in between , prefix(own block)
//      -- three or four shared-memory words, no tree
//   3. a chunk can hold a candidate only if its own minimum (CLOSE) or the minimum of the two chunks
//      its leaving elements e(p) come from (OPEN) is <= r0: about one chunk in 20
//   4. the warp takes flagged chunks one at a time: lanes 0-15 test CLOSE at position i, lanes 16-31
//      OPEN at step i against r0 and the chunk's own earlier positions; survivors are settled by the
//      whole warp over the < 32 window positions r0 does not cover. A tie on the key (identical
//      s-mers, i.e. tandem repeats; otherwise 2^-30) is resolved exactly by re-hashing only the tied
//      positions from the packed read
//   5. CLOSE/OPEN bits of a tile are combined in registers (E = ((C << 1) | carry) ^ O), chunks with
//      emissions are queued in a 64-entry list and written out as (sid, idx, start << 1 | open)
//      records when the list fills or the read ends; s-mer codes and k-mer hashes follow in sg_kmer.cu
#include <algorithm>
#include <cuda_pipeline.h>
#include "sg_common.cuh"
#include "sg_hash31.cuh"
#include "sg_internal.h"
#include "../../include/syncgpu.h"

namespace sg {

constexpr uint32_t HNONE = 0xffffffffu;  // "no hash" among chunk minima (32-bit)
constexpr uint32_t KNONE = 0xffffu;      // "no hash" among the per-position keys (15 significant bits in 16)
constexpr int LISTCAP = 64;            // queued chunks with emissions (a tile adds at most 32)
constexpr uint32_t TIE_TILE = 3;       // exact tie settlements a read may ask for inside one tile ...
constexpr uint32_t TIE_READ = 12;      // ... and in all (+ one per four tiles) before it is deferred to scan_exact_kernel

// Exact decision for a candidate whose key ties with the window minimum: the full 62-bit hashes of
// the tied positions are recomputed from the packed read and compared under the reference's rules.
// Whole warp; rare, so it is kept out of line to keep the tile loop inside the instruction cache.
__device__ __noinline__ bool settle_tie(const uint16_t *ring, int RCH, const uint32_t *hs32, int nwords, int s,
        int p, int q, bool is_open, uint32_t tgt, int lane)
{
    const int RM = RCH - 1, RS = RCH + 4;
    const uint64_t mask = (1ull << (2 * s)) - 1;
    auto ring_at = [&](int x) -> uint32_t { return ring[(x & 15) * RS + ((x >> 4) & RM)]; };
    auto m64_at = [&](int x) -> uint64_t {
        if (x < 0 || ring_at(x) == KNONE) return SG_NONE64;
        return hash64(smer_code_at(hs32, x, s, nwords) >> 1, mask);
    };
    // positions whose key equals tgt are the only ones that can hold the window minimum; the key is a
    // monotone function of the hash (its top bits, clamped below NONE), so the full hashes decide
    uint64_t best = SG_NONE64;
    for (int x = p - q + 1 + lane; x < p; x += 32)
        if (ring_at(x) == tgt) best = min(best, m64_at(x));
    const uint32_t bh = (uint32_t) (best >> 32), mh = __reduce_min_sync(SG_FULL, bh);
    const uint32_t ml = __reduce_min_sync(SG_FULL, bh == mh ? (uint32_t) best : 0xffffffffu);
    const uint64_t mo = (uint64_t) mh << 32 | ml;
    const uint64_t e64 = m64_at(p - q);
    if (is_open) return e64 <= mo;
    const uint64_t mp = m64_at(p);
    return mp <= mo && (mp <= e64 || mp < mo || m64_at(p - q + 1) == mp);
}

template <int S_FIXED, int RCH_FIXED, int LOGB_FIXED, int OSP>
__global__ void __launch_bounds__(32 * SYNC_SCAN_WARPS, 6) scan_kernel(ScanArgs A, ScanGeom G)
{
    extern __shared__ __align__(16) uint32_t smem[];
    // ring rows are RCH + 4 keys apart: a chunk's 16 positions (one column) and 32 consecutive positions
    // (two columns) then fall into distinct banks, like the 32 chunks of a tile (one row)
    const int RCH = RCH_FIXED ? RCH_FIXED : G.rch, RM = RCH - 1, RS = RCH + 4;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int NLV = OSP < 0 ? 1 : 2;               // arrays of chunk minima: one without the split
    uint32_t *wbase = smem + (size_t) wid * (8 * RS + (NLV + 1) * RCH + 2 * LISTCAP + 64);
    uint16_t *ring = reinterpret_cast<uint16_t *>(wbase);   // [16][RS] 15-bit keys of m[]
    // The elements e(p) = m[p - q] that leave the windows of a chunk's 16 positions are the positions of offset >= osp in
    // one chunk and those of offset < osp in the next (osp = -q mod 16, the same for every chunk): the minimum of a chunk
    // is kept in these two halves, so that the OPEN flag of step 3 looks at exactly the 16 leaving elements. With the
    // offset known at compile time (OSP >= 0: the shapes of the benchmark sweep) the split costs nothing in the hash loop;
    // otherwise both halves hold the whole chunk's minimum (the looser test of the two chunks).
    uint32_t *LvHi = wbase + 8 * RS;                   // [RCH] minimum over the chunk's positions of offset >= osp (32-bit)
    uint32_t *LvLo = OSP < 0 ? LvHi : LvHi + RCH;      // [RCH] ... of offset < osp
    uint32_t *sfxA = LvHi + NLV * RCH;                       // [RCH] suffix minimum of the chunk's block from the chunk on
    uint32_t *list_c = sfxA + RCH;                     // [LISTCAP] chunk index
    uint32_t *list_e = list_c + LISTCAP;               // [LISTCAP] E | Om << 16
    uint32_t *wbuf = list_e + LISTCAP;                 // [2][32] packed words of the next tile, filled by cp.async

    const int k = A.k, s = S_FIXED ? S_FIXED : A.s, q = k - s + 1;
    const uint64_t mask = (1ull << (2 * s)) - 1;
    const int rsh = 2 * s - 2;
    const bool small_q = q < 16;       // a thread's earlier positions fall out of the window: no running bound
    // per-position keys: the top 15 bits of the hash (a monotone map, so "<" on keys implies "<" on hashes and
    // equal keys go to the exact path); chunk minima keep 32 bits (s = 31) or the key itself (other s)
    const int kshift = max(0, 2 * s - 15);
    auto to_key = [&](uint32_t cm) -> uint32_t { return S_FIXED == 31 ? cm >> 17 : min(cm, 0x7fffu); };
    const int n_full = G.n_full, logB = LOGB_FIXED >= 0 ? LOGB_FIXED : G.logB, B = 1 << logB;
    // tile-invariant per lane: my place in my block of B chunks, and how many whole blocks lie between the
    // block that holds the first chunk of my window (c - n_full) and my own block
    const int lb = lane & (B - 1);
    const int n_between = n_full > 0 ? ((lane >> logB) - ((lane - n_full) >> logB) - 1) : 0;
    const int nb_common = __reduce_min_sync(SG_FULL, n_between);
    auto ring_at = [&](int x) -> uint32_t { return ring[(x & 15) * RS + ((x >> 4) & RM)]; };

    for (;;) {
        unsigned int r32 = 0;
        if (lane == 0) r32 = atomicAdd(A.work, 1u);
        r32 = __shfl_sync(SG_FULL, r32, 0);
        if ((uint64_t) r32 >= A.n_reads) break;
        const uint64_t r = r32;
        const int H = (int) A.hoco_l[r];
        if (H < k) { if (lane == 0) A.n_scm[r] = 0; continue; }
        const uint64_t hb = A.hoff[r];
        const uint32_t *hs32 = reinterpret_cast<const uint32_t *>(A.hoco_s + hb / 4);
        const uint16_t *nb16 = reinterpret_cast<const uint16_t *>(A.nbits + hb / 8);
        const int nwords = (H + 15) >> 4;
        const bool has_n = A.n_amb[r] != 0;

        // chunk minima in front of the read must say "no hash"; the position ring needs no clearing: every
        // position a window can reach (>= 0) is written, hash or NONE, by the tile that holds it
        for (int i = lane; i < (NLV + 1) * RCH; i += 32) LvHi[i] = HNONE;        // LvHi, (LvLo,) sfxA
        __syncwarp();

        uint32_t n_emitted = 0, carryC = 0;
        int n_list = 0;
        int last_n = -1;                                   // last ambiguous position seen so far
        // ties on the key are settled exactly one at a time (settle_tie, O(q) each): fine for the odd key
        // collision, ruinous inside a tandem repeat where every position ties. A read that ties more than
        // TIE_TILE times in one tile, or more than TIE_READ + tiles / 4 times in all, is handed from that
        // tile on to scan_exact_kernel, which keeps full hashes and costs the same whatever the input.
        uint32_t ties_read = 0;
        bool deferred = false;
        int tl = 0;
        uint32_t prev30 = 0, prev31 = 0;                   // the two words in front of the tile
        const int n_tiles = (H + 1 + 511) >> 9;
        // the packed word of the next tile travels global -> shared with cp.async one tile ahead, so no
        // register waits on it while the current tile is hashed
        auto fetch_word = [&](int cw, int slot) {
            if (cw < nwords) __pipeline_memcpy_async(wbuf + 32 * slot + lane, hs32 + cw, 4);
            else wbuf[32 * slot + lane] = 0u;
            __pipeline_commit();
        };
        fetch_word(lane, 0);

        // writes the records of the queued chunks
        auto flush = [&]() {
            uint32_t done = 0;
            for (int base = 0; base < n_list; base += 32) {
                const int j = base + lane;
                const uint32_t ev = j < n_list ? list_e[j] : 0u, cj = j < n_list ? list_c[j] : 0u;
                const uint32_t cnt = __popc(ev & 0xffffu);
                uint32_t inc = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(SG_FULL, inc, d); if (lane >= d) inc += t; }
                const uint32_t tot = __shfl_sync(SG_FULL, inc, 31);
                unsigned long long b = 0;
                if (lane == 0) b = atomicAdd(A.rec_count, (unsigned long long) tot);
                b = __shfl_sync(SG_FULL, b, 0);
                uint32_t idx = inc - cnt;                  // rank inside this batch of 32 queued chunks
                uint32_t E = ev & 0xffffu;
                const uint32_t Om = ev >> 16;
                while (E) {
                    const int i = __ffs(E) - 1;
                    E &= E - 1;
                    const uint32_t t = (uint32_t) ((int) (cj << 4) + i - k);   // k-mer start
                    const uint64_t o = b + idx;
                    // the s-mer code and the strand bit follow in kmerhash_kernel (one thread per syncmer, no divergence)
                    if (o < A.rec_cap) {
                        A.rec_sid[o] = (uint32_t) r;
                        A.rec_idx[o] = n_emitted + done + idx;
                        A.rec_mpos[o] = t << 1 | ((Om >> i) & 1u);             // bit 0: emitted by OPEN (first s-mer) or CLOSE (last s-mer)
                    }
                    ++idx;
                }
                done += tot;
            }
            n_emitted += done;
            n_list = 0;
            __syncwarp();
        };

        for (; tl < n_tiles; ++tl) {
            const int c = (tl << 5) + lane;                // my chunk
            uint32_t ties_tile = 0;
            const int P = c << 4;                          // its first position
            const int cs = c & RM;
            __pipeline_wait_prior(0);
            const uint32_t w0 = bswap32(wbuf[32 * (tl & 1) + lane]);
            fetch_word(c + 32, (tl + 1) & 1);
            uint32_t wb = __shfl_up_sync(SG_FULL, w0, 1), wa = __shfl_up_sync(SG_FULL, w0, 2);
            if (lane == 0) { wb = prev31; wa = prev30; }
            if (lane == 1) wa = prev31;
            prev30 = __shfl_sync(SG_FULL, w0, 30);
            prev31 = __shfl_sync(SG_FULL, w0, 31);

            // which of my 16 positions can carry a hash: >= s valid bases in a row, inside the read
            uint32_t vm, nb = 0;
            int l0 = P;                                    // valid bases in a row ending just before my chunk
            {
                const int to = min(16, max(0, H - P));
                if (!has_n) {
                    const int from = max(0, s - 1 - P);
                    vm = (from < to) ? ((0xffffu << from) & (0xffffu >> (16 - to))) : 0u;
                } else {
                    nb = c < nwords ? nb16[c] : 0u;
                    const int mine = nb ? P + 31 - __clz(nb) : -1;
                    int inc = mine;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(SG_FULL, inc, d); if (lane >= d) inc = max(inc, t); }
                    int exc = __shfl_up_sync(SG_FULL, inc, 1);
                    if (lane == 0) exc = -1;
                    const int before = max(last_n, exc);
                    last_n = max(last_n, __shfl_sync(SG_FULL, inc, 31));
                    l0 = P - 1 - before;
                    vm = 0;
                    int l = l0;
                    for (int i = 0; i < to; ++i) { l = ((nb >> i) & 1u) ? 0 : l + 1; vm |= (uint32_t) (l >= s) << i; }
                }
            }

            // 1. keys of my 16 positions
            uint32_t cmin = HNONE, cmLo = HNONE;           // minimum of the offsets >= OSP (all of them without a split) / < OSP
            uint32_t scan_up = 0;                          // see the partial chunks below: what the minimum may be short of
            uint16_t *own = ring + cs;
            if (vm && S_FIXED == 31) {
                // s = 31: every position is extracted straight from the three words around it (no rolling
                // dependency between positions) and hashed in the left-aligned frame of sg_hash31.cuh.
                // An odd s-mer cannot be its own reverse complement.
                const uint32_t ra = rev2(~w0), rb = rev2(~wb), rc = rev2(~wa);
                {
#define SG_H31_PAIR(J) { uint32_t hi, lo, hi2, lo2; h31_canon<J>(wa, wb, w0, ra, rb, rc, hi, lo); h31_canon<(J) + 1>(wa, wb, w0, ra, rb, rc, hi2, lo2); \
                        const uint32_t hv = h31_hash_top(hi, lo, G.h31), hv2 = h31_hash_top(hi2, lo2, G.h31); \
                        own[(J) * RS] = (uint16_t) (hv >> 17); own[((J) + 1) * RS] = (uint16_t) (hv2 >> 17); \
                        if (OSP < 0 || (J) >= OSP) cmin = min(cmin, min(hv, hv2)); \
                        else if ((J) + 1 < OSP) cmLo = min(cmLo, min(hv, hv2)); \
                        else { cmLo = min(cmLo, hv); cmin = min(cmin, hv2); } }
                    SG_H31_PAIR(0) SG_H31_PAIR(2) SG_H31_PAIR(4) SG_H31_PAIR(6) SG_H31_PAIR(8) SG_H31_PAIR(10) SG_H31_PAIR(12) SG_H31_PAIR(14)
#undef SG_H31_PAIR
                }
                if (vm != 0xffffu) {
                    // A chunk at either end of the read, or next to an ambiguous base: some of its positions carry no hash.
                    // It went through the same sixteen hashes as everybody else (a separate loop for it would make the
                    // other 31 lanes wait: two tiles of every read paid for both paths); the keys of the positions that
                    // do not count are now overwritten, and the chunk's minima are rebuilt from the KEYS of the others.
                    // A minimum known to 15 bits stands for a range of hashes: as the chunk's own value in the flag tests of
                    // step 3 the bottom of the range is used, as part of the window minimum of later chunks the top
                    // (scan_up), so that both tests can only flag more, never less; the decisions themselves are taken
                    // on keys, with ties settled exactly, whatever the flags were computed from.
                    uint32_t kh = KNONE, kl = KNONE;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {                 // unrolled: one lane runs this while 31 wait
                        const bool ok = (vm >> j) & 1u;
                        const uint32_t kj = ok ? (uint32_t) own[j * RS] : KNONE;
                        if (!ok) own[j * RS] = (uint16_t) KNONE;
                        if (OSP < 0 || j >= OSP) kh = min(kh, kj); else kl = min(kl, kj);
                    }
                    cmin = kh == KNONE ? HNONE : kh << 17;
                    cmLo = kl == KNONE ? HNONE : kl << 17;
                    scan_up = 0x1ffffu;
                }
            } else if (vm) {
                uint32_t ww = w0;
                const uint64_t V = (uint64_t) wa << 32 | wb;
                uint64_t fw = V & mask, rv = rc64(V) >> (64 - 2 * s);
                uint32_t vmr = vm;
#pragma unroll 4
                for (int j = 0; j < 16; ++j) {
                    const uint32_t b = ww >> 30;
                    ww <<= 2;
                    fw = ((fw << 2) | b) & mask;
                    rv = (rv >> 2) | ((uint64_t) (3u - b) << rsh);
                    const bool ok = (vmr & 1u) && fw != rv;
                    vmr >>= 1;
                    const uint32_t h = (uint32_t) (hash64(fw < rv ? fw : rv, mask) >> kshift);
                    own[j * RS] = (uint16_t) (ok ? h : KNONE);
                    if (ok) cmin = min(cmin, h);
                }
            } else {
#pragma unroll 4
                for (int i = 0; i < 16; ++i) own[i * RS] = (uint16_t) KNONE;
            }

            // 2. minimum over the n_full whole chunks in front of mine: prefix / suffix minima inside blocks of B lanes
            const uint32_t cmHi = cmin;
            cmin = min(cmin, cmLo);
            const uint32_t cscan = cmin == HNONE ? HNONE : cmin | scan_up;
            uint32_t pfx = cscan, sfx = cscan;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                if (d < B) {
                    const uint32_t up = __shfl_up_sync(SG_FULL, pfx, d), dn = __shfl_down_sync(SG_FULL, sfx, d);
                    if (lb >= d) pfx = min(pfx, up);
                    if (lb + d < B) sfx = min(sfx, dn);
                }
            }
            LvHi[cs] = cmHi;
            if (OSP >= 0) LvLo[cs] = cmLo;
            sfxA[cs] = sfx;
            uint32_t pex = __shfl_up_sync(SG_FULL, pfx, 1);                // my block's chunks in front of me
            if (lb == 0) pex = HNONE;
            __syncwarp();
            uint32_t r0 = HNONE;
            if (n_full > 0) {
                int sl = (c - n_full) & RM;                                  // first chunk of my window
                r0 = min(pex, sfxA[sl]);
                sl &= ~(B - 1);                                              // its block; block totals sit at block starts
                // the lanes of a warp differ by at most one in the number of whole blocks between (it steps once where the
                // window start crosses a block boundary): the common part runs unpredicated, the extra block is a select
                for (int j = 0; j < nb_common; ++j) { sl = (sl + B) & RM; r0 = min(r0, sfxA[sl]); }
                { const uint32_t t = sfxA[(sl + B) & RM]; r0 = n_between > nb_common ? min(r0, t) : r0; }
            }

            // 3. which chunks can hold a candidate at all
            uint32_t mC = 0xffffu, mO = 0xffffu;           // positions that may close / open at all
            if (P < k || P + 16 > H) {                     // only the first and last chunks of a read are partial
                mC = (P + 15 < k - 1 || P >= H) ? 0u :
                    ((0xffffu << max(0, k - 1 - P)) & (0xffffu >> (16 - min(16, H - P)))) & 0xffffu;
                mO = (P + 15 < k || P > H) ? 0u :
                    ((0xffffu << max(0, k - P)) & (0xffffu >> (16 - min(16, H + 1 - P)))) & 0xffffu;
            }
            bool flagged;
            if (small_q) flagged = (mC | mO) != 0;
            else {
                const int cA = (P - q) >> 4;
                const uint32_t emin = min(LvHi[cA & RM], LvLo[(cA + 1) & RM]);
                flagged = (mC && cmin <= r0) || (mO && emin <= r0);
            }

            // 4. the warp takes the flagged chunks one at a time
            uint32_t Cm = 0, Om = 0;
            {
                uint32_t any = __ballot_sync(SG_FULL, flagged);
                while (any) {
                    const int src = __ffs(any) - 1;
                    any &= any - 1;
                    const int lch = c - lane + src, lP = lch << 4;
                    const uint32_t lr0 = small_q ? KNONE : to_key(__shfl_sync(SG_FULL, r0, src));
                    const uint32_t lmask = __shfl_sync(SG_FULL, mC | mO << 16, src);
                    const int fc = small_q ? 0x7fffffff : ((n_full > 0 ? lch - n_full : lch) << 4);
                    const int li = lane & 15;
                    const uint32_t h = ring[li * RS + (lch & RM)];
                    uint32_t Rin = h;                      // inclusive prefix minimum inside each half warp
#pragma unroll
                    for (int d = 1; d < 16; d <<= 1) { const uint32_t t = __shfl_up_sync(SG_FULL, Rin, d, 16); if (li >= d) Rin = min(Rin, t); }
                    uint32_t Rex = __shfl_up_sync(SG_FULL, Rin, 1, 16);
                    Rex = min(li == 0 ? KNONE : Rex, lr0); // r0 and the chunk's earlier positions
                    const uint32_t mine = lane < 16 ? h : ring_at(lP + li - q);
                    const bool cand = mine != KNONE && ((lmask >> lane) & 1u) && (small_q || mine <= Rex);
                    uint32_t lc = __ballot_sync(SG_FULL, cand);
                    while (lc) {
                        const int bit = __ffs(lc) - 1;
                        lc &= lc - 1;
                        const int i = bit & 15, p = lP + i;
                        const bool is_open = bit >> 4;
                        // minimum of the keys over m[p-q+1 .. p-1]: Rex covers [fc, p), the rest is scanned
                        uint32_t m = __shfl_sync(SG_FULL, Rex, bit);
                        if (small_q) m = KNONE;
                        // (at most q mod 16 + 15 positions: one per lane)
                        { const int x = p - q + 1 + lane; if (x < min(fc, p)) m = min(m, ring_at(x)); }
                        const uint32_t Mhi = __reduce_min_sync(SG_FULL, m);
                        const uint32_t tgt = __shfl_sync(SG_FULL, mine, bit);
                        bool yes = tgt < Mhi;
                        if (tgt == Mhi) {
                            ++ties_tile;
                            if (ties_tile > TIE_TILE || ties_read + ties_tile > TIE_READ + ((uint32_t) tl >> 2)) {
                                deferred = true; lc = 0; any = 0; yes = false;
                            } else yes = settle_tie(ring, RCH, hs32, nwords, s, p, q, is_open, tgt, lane);
                        }
                        if (lane == src && yes) {
                            if (has_n) {
                                // run-length conditions that the position masks only imply for reads without N
                                auto run_len = [&](int ii) -> int {
                                    if (ii < 0) return l0;
                                    const uint32_t ml = nb & ((2u << ii) - 1u);
                                    return ml ? ii - (31 - __clz(ml)) : l0 + ii + 1;
                                };
                                if (is_open) yes = run_len(i - 1) >= k && (p == H || !((nb >> i) & 1u));
                                else yes = run_len(i) >= k;
                            }
                            if (yes) { if (is_open) Om |= 1u << i; else Cm |= 1u << i; }
                        }
                    }
                }
            }

            ties_read += ties_tile;
            if (deferred) break;                           // warp-uniform

            // 5. emissions of this tile: step i of chunk c emits iff CLOSE at the position before xor OPEN at step i
            {
                uint32_t pc = (__shfl_up_sync(SG_FULL, Cm, 1) >> 15) & 1u;
                if (lane == 0) pc = carryC;
                carryC = (__shfl_sync(SG_FULL, Cm, 31) >> 15) & 1u;
                const uint32_t E = (((Cm << 1) | pc) ^ Om) & 0xffffu;
                const uint32_t hit = __ballot_sync(SG_FULL, E != 0);
                if (hit) {
                    if (E) {
                        const int at = n_list + __popc(hit & ((1u << lane) - 1u));
                        list_c[at] = (uint32_t) c;
                        list_e[at] = E | Om << 16;
                    }
                    n_list += __popc(hit);
                    __syncwarp();
                    if (n_list > LISTCAP - 32) flush();
                }
            }
            __syncwarp();                                  // ring and minima slots are reused by later tiles
        }
        if (n_list) flush();
        if (deferred) {
            __pipeline_wait_prior(0);                      // the word of the next tile is still in flight
            if (lane == 0) {
                const unsigned int j = atomicAdd(A.defer_count, 1u);
                A.defer[4 * (size_t) j + 0] = (uint32_t) r;
                A.defer[4 * (size_t) j + 1] = (uint32_t) tl;        // first tile the exact kernel decides
                A.defer[4 * (size_t) j + 2] = n_emitted;            // records written so far (tiles < tl)
                A.defer[4 * (size_t) j + 3] = carryC;               // CLOSE at the last position of tile tl - 1
            }
        } else if (lane == 0) A.n_scm[r] = n_emitted;
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------------------
// scan_exact_kernel -- the same selection rules with FULL 62-bit hashes, for the reads scan_kernel
// deferred (tandem repeats, microsatellites, low-complexity sequence: anything where identical s-mers
// tie for the window minimum again and again). Nothing here depends on how often values tie, so the
// cost per position is a constant (about twice scan_kernel's on random sequence).
//
// One warp per deferred read, same tiles of 32 chunks x 16 positions. Per chunk the lane keeps
//   ringM[x]  m[x], the hash of the s-mer ending at x (or NONE)
//   ringS[x]  min m[x .. end of x's chunk]            (suffix minimum inside the chunk)
//   cm[c]     minimum of chunk c                       (shared memory)
// in rings of RLEN >= q + 544 positions (ringM / ringS in global scratch, one slice per resident warp:
// they stay in L1/L2). The window of position p = P + i, [a, p - 1] with a = p - q + 1, is then
//   ringS[a]  (rest of a's chunk)  ,  cm[chunk(a) + 1 .. c - 1]  ,  own positions P .. p - 1
// and e(p) = ringM[a - 1], m[p - q + 1] = ringM[a]. The whole-chunk part differs between the 16
// positions of a chunk only in whether chunk(a) is cA or cA + 1 (cA = chunk of the window start of the
// chunk's first position), so it is one loop over q / 16 shared-memory words per lane and tile.
// Windows shorter than two chunks read ringM position by position.
struct ExactEmit {
    uint32_t *list_c, *list_e;
    int n_list;
    uint32_t n_emitted;
};

__device__ __forceinline__ void exact_flush(const ScanArgs &A, ExactEmit &S, uint32_t r, int k, int lane)
{
    uint32_t done = 0;
    for (int base = 0; base < S.n_list; base += 32) {
        const int j = base + lane;
        const uint32_t ev = j < S.n_list ? S.list_e[j] : 0u, cj = j < S.n_list ? S.list_c[j] : 0u;
        const uint32_t cnt = __popc(ev & 0xffffu);
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(SG_FULL, inc, d); if (lane >= d) inc += t; }
        const uint32_t tot = __shfl_sync(SG_FULL, inc, 31);
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(A.rec_count, (unsigned long long) tot);
        b = __shfl_sync(SG_FULL, b, 0);
        uint32_t idx = inc - cnt;
        uint32_t E = ev & 0xffffu;
        const uint32_t Om = ev >> 16;
        while (E) {
            const int i = __ffs(E) - 1;
            E &= E - 1;
            const uint32_t t = (uint32_t) ((int) (cj << 4) + i - k);
            const uint64_t o = b + idx;
            if (o < A.rec_cap) {
                A.rec_sid[o] = r;
                A.rec_idx[o] = S.n_emitted + done + idx;
                A.rec_mpos[o] = t << 1 | ((Om >> i) & 1u);
            }
            ++idx;
        }
        done += tot;
    }
    S.n_emitted += done;
    S.n_list = 0;
    __syncwarp();
}

__global__ void __launch_bounds__(32 * SYNC_SCAN_WARPS) scan_exact_kernel(ScanArgs A, int rlen_log)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int RLEN = 1 << rlen_log, RMASK = RLEN - 1, CCH = RLEN >> 4, CM = CCH - 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t *cm = reinterpret_cast<uint64_t *>(smem) + (size_t) wid * (CCH + LISTCAP);
    ExactEmit S;
    S.list_c = reinterpret_cast<uint32_t *>(cm + CCH);
    S.list_e = S.list_c + LISTCAP;
    uint64_t *ringM = A.xring + ((size_t) blockIdx.x * SYNC_SCAN_WARPS + wid) * 2 * RLEN;
    uint64_t *ringS = ringM + RLEN;

    const int k = A.k, s = A.s, q = k - s + 1;
    const uint64_t mask = (1ull << (2 * s)) - 1;
    const int rsh = 2 * s - 2;
    const unsigned int n_def = *A.defer_count;

    for (;;) {
        unsigned int j32 = 0;
        if (lane == 0) j32 = atomicAdd(A.work2, 1u);
        j32 = __shfl_sync(SG_FULL, j32, 0);
        if (j32 >= n_def) break;
        const uint32_t r = A.defer[4 * (size_t) j32], t0 = A.defer[4 * (size_t) j32 + 1];
        S.n_emitted = A.defer[4 * (size_t) j32 + 2];
        S.n_list = 0;
        uint32_t carryC = A.defer[4 * (size_t) j32 + 3];
        const int H = (int) A.hoco_l[r];
        const uint64_t hb = A.hoff[r];
        const uint32_t *hs32 = reinterpret_cast<const uint32_t *>(A.hoco_s + hb / 4);
        const uint16_t *nb16 = reinterpret_cast<const uint16_t *>(A.nbits + hb / 8);
        const int nwords = (H + 15) >> 4;
        const bool has_n = A.n_amb[r] != 0;
        const int n_tiles = (H + 1 + 511) >> 9;
        // hashes of the q + 1 positions in front of tile t0 are needed before anything is decided; with
        // ambiguous bases the run lengths need the whole prefix, so such a read is re-hashed from its start
        const int tw = has_n ? 0 : max(0, ((int) t0 * 512 - q - 16) >> 9);
        int last_n = -1;

        for (int i = lane; i < CCH; i += 32) cm[i] = SG_NONE64;
        __syncwarp();

        for (int tl = tw; tl < n_tiles; ++tl) {
            const int c = (tl << 5) + lane, P = c << 4;
            const uint32_t w0 = hoco_word(hs32, c, nwords), wb = hoco_word(hs32, c - 1, nwords), wa = hoco_word(hs32, c - 2, nwords);

            uint32_t vm, nb = 0;
            int l0 = P;
            {
                const int to = min(16, max(0, H - P));
                if (!has_n) {
                    const int from = max(0, s - 1 - P);
                    vm = (from < to) ? ((0xffffu << from) & (0xffffu >> (16 - to))) : 0u;
                } else {
                    nb = c < nwords ? nb16[c] : 0u;
                    const int mine = nb ? P + 31 - __clz(nb) : -1;
                    int inc = mine;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(SG_FULL, inc, d); if (lane >= d) inc = max(inc, t); }
                    int exc = __shfl_up_sync(SG_FULL, inc, 1);
                    if (lane == 0) exc = -1;
                    const int before = max(last_n, exc);
                    last_n = max(last_n, __shfl_sync(SG_FULL, inc, 31));
                    l0 = P - 1 - before;
                    vm = 0;
                    int l = l0;
                    for (int i = 0; i < to; ++i) { l = ((nb >> i) & 1u) ? 0 : l + 1; vm |= (uint32_t) (l >= s) << i; }
                }
            }

            // full hashes of my 16 positions and their suffix minima
            uint64_t m[16];
            {
                uint32_t ww = w0;
                const uint64_t V = (uint64_t) wa << 32 | wb;
                uint64_t fw = V & mask, rv = rc64(V) >> (64 - 2 * s);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t b = ww >> 30;
                    ww <<= 2;
                    fw = ((fw << 2) | b) & mask;
                    rv = (rv >> 2) | ((uint64_t) (3u - b) << rsh);
                    const bool ok = ((vm >> j) & 1u) && fw != rv;
                    m[j] = ok ? hash64(fw < rv ? fw : rv, mask) : SG_NONE64;
                }
            }
            {
                uint64_t sm = SG_NONE64;
                uint64_t *dm = ringM + (P & RMASK), *ds = ringS + (P & RMASK);
#pragma unroll
                for (int j = 15; j >= 0; --j) { sm = min(sm, m[j]); dm[j] = m[j]; ds[j] = sm; }
                cm[c & CM] = sm;
            }
            __syncwarp();
            if (tl < (int) t0) continue;                   // warm-up tile: hashes only

            uint32_t mC = 0xffffu, mO = 0xffffu;
            if (P < k || P + 16 > H) {
                mC = (P + 15 < k - 1 || P >= H) ? 0u :
                    ((0xffffu << max(0, k - 1 - P)) & (0xffffu >> (16 - min(16, H - P)))) & 0xffffu;
                mO = (P + 15 < k || P > H) ? 0u :
                    ((0xffffu << max(0, k - P)) & (0xffffu >> (16 - min(16, H + 1 - P)))) & 0xffffu;
            }
            uint32_t Cm = 0, Om = 0;
            if (mC | mO) {
                const int a0 = P - q + 1, cA = a0 >> 4;    // arithmetic shift: floor
                uint64_t R2 = SG_NONE64, c1 = SG_NONE64;
                const bool wide = cA + 1 < c;              // every window of this chunk starts in an earlier chunk
                if (wide) {
                    c1 = cm[(cA + 1) & CM];
                    for (int x = cA + 2; x < c; ++x) R2 = min(R2, cm[x & CM]);
                    // nothing can close unless it is at most every whole chunk in its window, nothing can open
                    // unless the element leaving is: whole-chunk minima decide for most chunks
                    const uint64_t own = ringS[P & RMASK];
                    // the leaving elements e(p) = m[a0 - 1 + i] sit in the two chunks from (a0 - 1) >> 4 on (that is
                    // cA - 1 and cA when q - 1 is a multiple of 16)
                    const int cE = (a0 - 1) >> 4;
                    const uint64_t lv = min(cE >= 0 ? cm[cE & CM] : SG_NONE64, cE + 1 >= 0 ? cm[(cE + 1) & CM] : SG_NONE64);
                    if (own > R2) mC = 0;
                    if (lv > R2) mO = 0;
                }
                uint64_t pre = SG_NONE64;                  // min m[P .. p-1]
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (((mC | mO) >> i) & 1u) {
                        const int p = P + i, a = p - q + 1;   // a >= 0 whenever bit i of mC or mO is set
                        uint64_t mo;
                        if (wide) {
                            mo = min(min(ringS[a & RMASK], pre), (a >> 4) == cA ? min(c1, R2) : R2);
                        } else {
                            mo = SG_NONE64;
                            for (int x = a; x < p; ++x) mo = min(mo, ringM[x & RMASK]);
                        }
                        const uint64_t e = a >= 1 ? ringM[(a - 1) & RMASK] : SG_NONE64;
                        const uint64_t me = m[i];
                        bool cl = ((mC >> i) & 1u) && me != SG_NONE64 && me <= mo &&
                                  (me <= e || me < mo || ringM[a & RMASK] == me);
                        bool op = ((mO >> i) & 1u) && e != SG_NONE64 && e <= mo;
                        if (has_n && (cl || op)) {
                            auto run_len = [&](int ii) -> int {
                                if (ii < 0) return l0;
                                const uint32_t ml = nb & ((2u << ii) - 1u);
                                return ml ? ii - (31 - __clz(ml)) : l0 + ii + 1;
                            };
                            if (op) op = run_len(i - 1) >= k && (p == H || !((nb >> i) & 1u));
                            if (cl) cl = run_len(i) >= k;
                        }
                        Cm |= (uint32_t) cl << i;
                        Om |= (uint32_t) op << i;
                    }
                    pre = min(pre, m[i]);
                }
            }

            {
                uint32_t pc = (__shfl_up_sync(SG_FULL, Cm, 1) >> 15) & 1u;
                if (lane == 0) pc = carryC;
                carryC = (__shfl_sync(SG_FULL, Cm, 31) >> 15) & 1u;
                const uint32_t E = (((Cm << 1) | pc) ^ Om) & 0xffffu;
                const uint32_t hit = __ballot_sync(SG_FULL, E != 0);
                if (hit) {
                    if (E) {
                        const int at = S.n_list + __popc(hit & ((1u << lane) - 1u));
                        S.list_c[at] = (uint32_t) c;
                        S.list_e[at] = E | Om << 16;
                    }
                    S.n_list += __popc(hit);
                    __syncwarp();
                    if (S.n_list > LISTCAP - 32) exact_flush(A, S, r, k, lane);
                }
            }
            __syncwarp();
        }
        if (S.n_list) exact_flush(A, S, r, k, lane);
        if (lane == 0) A.n_scm[r] = S.n_emitted;
        __syncwarp();
    }
}

// ring length (log2, in positions) of scan_exact_kernel and the global scratch it wants
static int exact_rlen_log(int k, int s)
{
    const int need = (k - s + 1) + 512 + 32;
    int lg = 10;
    while ((1 << lg) < need) ++lg;
    return lg;
}

size_t scan_exact_scratch_bytes(int k, int s, int *grid_out)
{
    const size_t per_warp = (size_t) 2 * sizeof(uint64_t) << exact_rlen_log(k, s);
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    size_t grid = (size_t) n_sm * 4;
    const size_t budget = (size_t) 512 << 20;
    while (grid > 1 && grid * SYNC_SCAN_WARPS * per_warp > budget) grid >>= 1;
    if (grid_out) *grid_out = (int) grid;
    return grid * SYNC_SCAN_WARPS * per_warp;
}

static int launch_scan_exact(const ScanArgs &A, cudaStream_t st)
{
    const int lg = exact_rlen_log(A.k, A.s);
    const size_t smem = (size_t) SYNC_SCAN_WARPS * ((size_t) (1 << lg) / 16 + LISTCAP) * sizeof(uint64_t);
    if (smem > 227 * 1024) return SG_E_KSIZE;
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem_set < smem) {
        if (cudaFuncSetAttribute(scan_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return SG_E_CUDA;
        smem_set = smem;
    }
    int grid = 1;
    scan_exact_scratch_bytes(A.k, A.s, &grid);
    const uint64_t want = (A.n_reads + SYNC_SCAN_WARPS - 1) / SYNC_SCAN_WARPS;
    scan_exact_kernel<<<(unsigned) std::min<uint64_t>(want, (uint64_t) grid), 32 * SYNC_SCAN_WARPS, smem, st>>>(A, lg);
    return 1;
}

int scan_geometry(int k, int s, ScanGeom *g, size_t *smem_per_warp)
{
    const int q = k - s + 1;
    int n_full = q / 16 - 1;
    if (n_full < 0) n_full = 0;
    int logB = 0;
    while (logB < 5 && (2 << logB) <= n_full) ++logB;      // largest B = 2^logB <= min(32, n_full)
    const int need = (q + 15) / 16 + 32 + 2;
    int rch = 64;
    while (rch < need) rch <<= 1;
    g->rch = rch; g->n_full = n_full; g->logB = logB; g->h31 = h31_consts();
    *smem_per_warp = sizeof(uint32_t) * ((size_t) 8 * (rch + 4) + 2 * rch + 2 * LISTCAP + 64);
    return *smem_per_warp <= 227 * 1024 ? 0 : SG_E_KSIZE;
}

template <int S_FIXED, int RCH_FIXED, int LOGB_FIXED, int OSP>
static int launch_scan_t(const ScanArgs &A, const ScanGeom &g, size_t smem_per_warp, cudaStream_t st)
{
    auto kern = scan_kernel<S_FIXED, RCH_FIXED, LOGB_FIXED, OSP>;
    static int ctas_per_sm = 0, n_sm = 0;
    static size_t smem_set = 0;
    int warps = SYNC_SCAN_WARPS;
    while (warps > 1 && smem_per_warp * warps > 227 * 1024) warps >>= 1;
    const size_t smem = smem_per_warp * warps;
    if (smem_set != smem) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return SG_E_CUDA;
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return SG_E_CUDA;
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return SG_E_CUDA;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, 32 * warps, smem) != cudaSuccess) return SG_E_CUDA;
        if (ctas_per_sm < 1) return SG_E_KSIZE;
        smem_set = smem;
    }
    const uint64_t want = (A.n_reads + warps - 1) / warps;
    const unsigned grid = (unsigned) std::min<uint64_t>(want, (uint64_t) n_sm * ctas_per_sm);
    kern<<<grid, 32 * warps, smem, st>>>(A, g);
    return 1;
}

int launch_scan(const ScanArgs &A, cudaStream_t st)
{
    ScanGeom g;
    size_t spw;
    if (scan_geometry(A.k, A.s, &g, &spw)) return SG_E_KSIZE;
    if (A.n_reads == 0) return 0;
    int rc;
    if (A.s == 31) {
        // the shapes of the benchmark sweep (k = 501, 1001, 2001) get compile-time ring and block sizes
        // ... and a compile-time split of the chunk minima at -q mod 16 (k = 501: 9, k = 1001: 5, k = 2001: 13)
        const int osp = (int) ((16u - ((unsigned) (A.k - A.s + 1) & 15u)) & 15u);
        // (one more array of chunk minima: taken where the shared memory it costs does not cost a resident CTA)
        const size_t spw_split = spw + sizeof(uint32_t) * (size_t) g.rch;
        if (g.rch == 64 && g.logB == 4 && osp == 9) rc = launch_scan_t<31, 64, 4, 9>(A, g, spw_split, st);
        else if (g.rch == 128 && g.logB == 5 && osp == 5) rc = launch_scan_t<31, 128, 5, 5>(A, g, spw_split, st);
        else if (g.rch == 64 && g.logB == 4) rc = launch_scan_t<31, 64, 4, -1>(A, g, spw, st);
        else if (g.rch == 128 && g.logB == 5) rc = launch_scan_t<31, 128, 5, -1>(A, g, spw, st);
        else if (g.rch == 256 && g.logB == 5) rc = launch_scan_t<31, 256, 5, -1>(A, g, spw, st);
        else rc = launch_scan_t<31, 0, -1, -1>(A, g, spw, st);
    } else rc = launch_scan_t<0, 0, -1, -1>(A, g, spw, st);
    if (rc < 0) return rc;
    // the reads scan_kernel deferred (none on ordinary sequence: the kernel then finds an empty list and returns)
    const int rc2 = launch_scan_exact(A, st);
    return rc2 < 0 ? rc2 : rc + rc2;
}

} // namespace sg
