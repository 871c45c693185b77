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
// One CTA per read walks it in tiles of NT*16 positions; thread t owns 16
// consecutive positions (one 32-bit word of packed bases).
//   1. roll both strands through the 16 bases and hash; only the HIGH 32 bits of
//      each 62-bit hash are kept, in a shared-memory ring that holds the last
//      q + tile positions, transposed [16][chunks] so every access is conflict free
//   2. a radix-4 sparse table over the per-chunk minima gives every thread
//      r0 = min over the chunks that lie fully inside the window of all its 16 positions
//   3. a position is a candidate when its high word (or that of e(p)) is <= the
//      running minimum of r0 and the thread's own earlier positions: about 2/q of
//      all positions pass
//   4. candidates are settled by the whole warp: the < 48 window positions that r0
//      and the running minimum do not cover are min-reduced on the high word. A tie
//      on the high word (identical s-mers, i.e. tandem repeats; otherwise 2^-30) is
//      resolved exactly by re-hashing just the tied positions from the packed read
//   5. CLOSE/OPEN bits are parked per chunk; once per read (or every 32 k
//      positions) they are combined, ranked with one block scan and written as
//      (sid, idx, m_pos, s_mer) records; k-mer hashes follow in sg_kmer.cu
#include "sg_common.cuh"
#include "sg_hash31.cuh"
#include "sg_internal.h"
#include "../../include/syncgpu.h"

namespace sg {

constexpr int COCAP = 1024;            // chunks of parked CLOSE/OPEN bits (16 k positions)
constexpr uint32_t HNONE = 0xffffffffu;

// Exact decision for a candidate whose high word ties with the window minimum: the
// full 62-bit hashes of the tied positions are recomputed from the packed read and
// compared under the reference's rules. Whole warp; rare (identical s-mers inside
// one window), so it is kept out of line to keep the tile loop inside the I-cache.
__device__ __noinline__ bool settle_tie(const uint32_t *ring, int RCH, const uint32_t *hs32, int nwords, int s,
        int p, int q, bool is_open, uint32_t tgt, int lane)
{
    const int RM = RCH - 1;
    const uint64_t mask = (1ull << (2 * s)) - 1;
    auto ring_at = [&](int x) -> uint32_t { return ring[(x & 15) * RCH + ((x >> 4) & RM)]; };
    auto m64_at = [&](int x) -> uint64_t {
        if (x < 0 || ring_at(x) == HNONE) return SG_NONE64;
        return hash64(smer_code_at(hs32, x, s, nwords) >> 1, mask);
    };
    // positions whose ring word equals tgt are the only ones that can hold the window minimum; the ring
    // word is a monotone function of the hash (its top bits, clamped below NONE), so the full hashes decide
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

template <int NT, int S_FIXED>
__global__ void __launch_bounds__(NT) scan_kernel(ScanArgs A, ScanGeom G)
{
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) uint32_t smem[];
    const int RCH = G.rch, RM = RCH - 1;
    uint32_t *ring = smem;                             // [16][RCH] high words of m[]
    uint32_t *Lv = ring + 16 * RCH;                    // [T+1][RCH] radix-4 sparse table, level 0 = chunk minima
    uint32_t *co = Lv + (G.T + 1) * RCH;               // [COCAP] Cm | Om << 16 per chunk
    uint32_t *s_scan = co + COCAP;                     // [NW + 1]
    int *s_misc = reinterpret_cast<int *>(s_scan + NW + 1);   // [NW + 4]

    const uint64_t r = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int H = (int) A.hoco_l[r];
    const int k = A.k, s = S_FIXED ? S_FIXED : A.s, q = k - s + 1;
    if (H < k) { if (tid == 0) A.n_scm[r] = 0; return; }
    const uint64_t hb = A.hoff[r];
    const uint32_t *hs32 = reinterpret_cast<const uint32_t *>(A.hoco_s + hb / 4);
    const uint16_t *nb16 = reinterpret_cast<const uint16_t *>(A.nbits + hb / 8);
    const int nwords = (H + 15) >> 4;
    const bool has_n = A.n_amb[r] != 0;
    const uint64_t mask = (1ull << (2 * s)) - 1;
    const int rsh = 2 * s - 2;
    const bool small_q = q < 16;       // a thread's earlier positions fall out of the window: no running bound
    const int n_full = G.n_full, T = G.T, W = 1 << (2 * T);

    for (int i = tid; i < 16 * RCH; i += NT) ring[i] = HNONE;
    for (int i = tid; i < (T + 1) * RCH; i += NT) Lv[i] = HNONE;
    if (tid == 0) s_misc[NW] = -1;                     // last ambiguous position seen so far
    __syncthreads();

    auto ring_at = [&](int x) -> uint32_t { return ring[(x & 15) * RCH + ((x >> 4) & RM)]; };
    uint32_t n_emitted = 0, carryC = 0;
    int co_base = 0;                                   // chunk index of co[0]
    const int n_sub = (H + 1 + NT * 16 - 1) / (NT * 16);

    // ranks the parked bits of chunks [co_base, c_end) and writes their records
    auto flush = [&](int c_end) {
        const int nb = c_end - co_base;
        const int cpt = (nb + NT - 1) / NT;            // chunks per thread, contiguous
        const int j0 = min(tid * cpt, nb), j1 = min(j0 + cpt, nb);
        uint32_t cnt = 0;
        for (int j = j0; j < j1; ++j) {
            const uint32_t v = co[j], pv = j ? (co[j - 1] >> 15) & 1u : carryC;
            cnt += __popc((((v << 1) | pv) ^ (v >> 16)) & 0xffffu);
        }
        uint32_t tot;
        uint32_t idx = BlockScanU32::run<NW>(cnt, s_scan, &tot);
        if (tot) {
            if (tid == 0) {
                unsigned long long b = atomicAdd(A.rec_count, (unsigned long long) tot);
                s_misc[NW + 1] = (int) (uint32_t) b;
                s_misc[NW + 2] = (int) (uint32_t) (b >> 32);
            }
            __syncthreads();
            const uint64_t base = (uint64_t) (uint32_t) s_misc[NW + 1] | (uint64_t) (uint32_t) s_misc[NW + 2] << 32;
            for (int j = j0; j < j1; ++j) {
                const uint32_t v = co[j], pv = j ? (co[j - 1] >> 15) & 1u : carryC, Om = v >> 16;
                uint32_t E = (((v << 1) | pv) ^ Om) & 0xffffu;
                while (E) {
                    const int i = __ffs(E) - 1;
                    E &= E - 1;
                    const int t = ((co_base + j) << 4) + i - k;      // k-mer start
                    const bool open = (Om >> i) & 1u;
                    // OPEN: code of the first s-mer; CLOSE: code of the last s-mer with its low bit flipped
                    const uint64_t raw = smer_code_at(hs32, open ? t + s - 1 : t + k - 1, s, nwords);
                    const uint64_t o = base + idx;
                    if (o < A.rec_cap) {
                        A.rec_sid[o] = (uint32_t) r;
                        A.rec_idx[o] = n_emitted + idx;
                        A.rec_mpos[o] = (uint32_t) t << 1 | (uint32_t) (raw & 1ull);
                        A.rec_smer[o] = open ? raw : raw ^ 1ull;
                    }
                    ++idx;
                }
            }
        }
        __syncthreads();
        if (nb) carryC = (co[nb - 1] >> 15) & 1u;
        n_emitted += tot;
        co_base = c_end;
        __syncthreads();
    };

    for (int sub = 0; ; ++sub) {
        // single call site: when the parking buffer would overflow, and once at the end of the read
        if (sub == n_sub || sub * NT + NT - co_base > COCAP) flush(sub * NT);      // uniform: depends on sub only
        if (sub == n_sub) break;
        const int c = sub * NT + tid;                  // my chunk
        const int P = c << 4;                          // its first position
        const int cs = c & RM;

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
                int mine = nb ? P + 31 - __clz(nb) : -1, inc = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(SG_FULL, inc, d); if (lane >= d) inc = max(inc, t); }
                if (lane == 31) s_misc[wid] = inc;
                __syncthreads();
                int before = s_misc[NW];
                for (int w = 0; w < wid; ++w) before = max(before, s_misc[w]);
                int exc = __shfl_up_sync(SG_FULL, inc, 1);
                if (lane == 0) exc = -1;
                before = max(before, exc);
                l0 = P - 1 - before;
                __syncthreads();
                if (tid == NT - 1) s_misc[NW] = max(before, mine);
                vm = 0;
                int l = l0;
                for (int i = 0; i < to; ++i) { l = ((nb >> i) & 1u) ? 0 : l + 1; vm |= (uint32_t) (l >= s) << i; }
            }
        }

        // 1. hashes of my 16 positions (4 x 4: the loop keeps the tile body inside the instruction cache).
        //    An s-mer of odd length cannot be its own reverse complement, so a chunk whose 16 positions
        //    are all valid needs no per-position checks.
        uint32_t cmin = HNONE;
        uint32_t *own = ring + cs;
        if (vm && S_FIXED == 31) {
            // s = 31: every position is extracted straight from the three words around it (no rolling
            // dependency between positions) and hashed in the left-aligned frame of sg_hash31.cuh. The ring
            // word is hash >> 30 clamped below NONE. An odd s-mer cannot be its own reverse complement.
            const uint32_t a = hoco_word(hs32, c - 2, nwords), b = hoco_word(hs32, c - 1, nwords), w0 = hoco_word(hs32, c, nwords);
            const uint32_t ra = rev2(~w0), rb = rev2(~b), rc = rev2(~a);
            if (vm == 0xffffu) {
#define SG_H31_POS(J) { uint32_t hi, lo; h31_canon<J>(a, b, w0, ra, rb, rc, hi, lo); \
                        const uint32_t hv = min(h31_hash_top(hi, lo, G.h31), 0xfffffffeu); own[(J) * RCH] = hv; cmin = min(cmin, hv); }
                SG_H31_POS(0) SG_H31_POS(1) SG_H31_POS(2) SG_H31_POS(3) SG_H31_POS(4) SG_H31_POS(5) SG_H31_POS(6) SG_H31_POS(7)
                SG_H31_POS(8) SG_H31_POS(9) SG_H31_POS(10) SG_H31_POS(11) SG_H31_POS(12) SG_H31_POS(13) SG_H31_POS(14) SG_H31_POS(15)
#undef SG_H31_POS
            } else {
#pragma unroll 1
                for (int j = 0; j < 16; ++j) {
                    uint32_t hi, lo;
                    h31_canon_rt(j, a, b, w0, ra, rb, rc, hi, lo);
                    const uint32_t hv = ((vm >> j) & 1u) ? min(h31_hash_top(hi, lo, G.h31), 0xfffffffeu) : HNONE;
                    own[j * RCH] = hv;
                    cmin = min(cmin, hv);
                }
            }
        } else if (vm) {
            uint32_t w0 = hoco_word(hs32, c, nwords);
            const uint64_t V = (uint64_t) hoco_word(hs32, c - 2, nwords) << 32 | hoco_word(hs32, c - 1, nwords);
            uint64_t fw = V & mask, rv = rc64(V) >> (64 - 2 * s);
            uint32_t *dst = own;
            if (vm == 0xffffu && (s & 1)) {
                for (int i4 = 0; i4 < 4; ++i4) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t b = w0 >> 30;
                        w0 <<= 2;
                        fw = ((fw << 2) | b) & mask;
                        rv = (rv >> 2) | ((uint64_t) (3u - b) << rsh);
                        const uint32_t hv = (uint32_t) (hash64(fw < rv ? fw : rv, mask) >> 32);
                        dst[j * RCH] = hv;
                        cmin = min(cmin, hv);
                    }
                    dst += 4 * RCH;
                }
            } else {
                uint32_t vmr = vm;
                for (int i4 = 0; i4 < 4; ++i4) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t b = w0 >> 30;
                        w0 <<= 2;
                        fw = ((fw << 2) | b) & mask;
                        rv = (rv >> 2) | ((uint64_t) (3u - b) << rsh);
                        const bool ok = (vmr & 1u) && fw != rv;
                        vmr >>= 1;
                        const uint32_t h = (uint32_t) (hash64(fw < rv ? fw : rv, mask) >> 32);
                        const uint32_t hv = ok ? h : HNONE;
                        dst[j * RCH] = hv;
                        cmin = min(cmin, hv);
                    }
                    dst += 4 * RCH;
                }
            }
        } else {
#pragma unroll 4
            for (int i = 0; i < 16; ++i) own[i * RCH] = HNONE;
        }
        Lv[cs] = cmin;
        __syncthreads();

        // 2. radix-4 sparse table over chunk minima: level t holds the minimum of 4^t chunks ending at c
        {
            uint32_t v = cmin;
            for (int t = 1; t <= T; ++t) {
                const int st = 1 << (2 * (t - 1));
                const uint32_t *L = Lv + (t - 1) * RCH;
                v = min(min(v, L[(c - st) & RM]), min(L[(c - 2 * st) & RM], L[(c - 3 * st) & RM]));
                Lv[t * RCH + cs] = v;
                __syncthreads();
            }
        }
        uint32_t r0 = HNONE;
        if (n_full > 0) {
            const uint32_t *L = Lv + T * RCH;
            for (int e = c - 1; e - W + 1 > c - n_full; e -= W) r0 = min(r0, L[e & RM]);
            r0 = min(r0, L[(c - n_full + W - 1) & RM]);
        }

        // 3. which chunks can hold a candidate at all: its own minimum (CLOSE) or the minimum of the two
        //    chunks its leaving elements e(p) come from (OPEN) must not exceed r0. About 1 chunk in 20.
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
            const uint32_t emin = min(Lv[cA & RM], Lv[(cA + 1) & RM]);
            flagged = (mC && cmin <= r0) || (mO && emin <= r0);
        }

        // 4. the warp takes the flagged chunks one at a time: lanes 0-15 test CLOSE at position i,
        //    lanes 16-31 test OPEN at step i against the running minimum, then the survivors are settled
        uint32_t Cm = 0, Om = 0;
        {
            uint32_t any = __ballot_sync(SG_FULL, flagged);
            while (any) {
                const int src = __ffs(any) - 1;
                any &= any - 1;
                const int lch = c - lane + src, lP = lch << 4;
                const uint32_t lr0 = small_q ? HNONE : __shfl_sync(SG_FULL, r0, src);
                const uint32_t lmask = __shfl_sync(SG_FULL, mC | mO << 16, src);
                const int fc = small_q ? 0x7fffffff : ((n_full > 0 ? lch - n_full : lch) << 4);
                const int li = lane & 15;
                const uint32_t h = ring[li * RCH + (lch & RM)];
                uint32_t Rin = h;                      // inclusive prefix minimum inside each half warp
#pragma unroll
                for (int d = 1; d < 16; d <<= 1) { const uint32_t t = __shfl_up_sync(SG_FULL, Rin, d, 16); if (li >= d) Rin = min(Rin, t); }
                uint32_t Rex = __shfl_up_sync(SG_FULL, Rin, 1, 16);
                Rex = min(li == 0 ? HNONE : Rex, lr0); // r0 and the chunk's earlier positions
                const uint32_t mine = lane < 16 ? h : ring_at(lP + li - q);
                const bool cand = mine != HNONE && ((lmask >> lane) & 1u) && (small_q || mine <= Rex);
                uint32_t lc = __ballot_sync(SG_FULL, cand);
                while (lc) {
                    const int bit = __ffs(lc) - 1;
                    lc &= lc - 1;
                    const int i = bit & 15, p = lP + i;
                    const bool is_open = bit >> 4;
                    // minimum of the high words over m[p-q+1 .. p-1]: Rex covers [fc, p), the rest is scanned
                    uint32_t m = __shfl_sync(SG_FULL, Rex, bit);
                    if (small_q) m = HNONE;
                    for (int x = p - q + 1 + lane; x < min(fc, p); x += 32) m = min(m, ring_at(x));
                    const uint32_t Mhi = __reduce_min_sync(SG_FULL, m);
                    const uint32_t tgt = __shfl_sync(SG_FULL, mine, bit);
                    bool yes = tgt < Mhi;
                    if (tgt == Mhi) yes = settle_tie(ring, RCH, hs32, nwords, s, p, q, is_open, tgt, lane);
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
        co[c - co_base] = Cm | Om << 16;
        __syncthreads();                               // ring and table slots are reused by the next tile
    }
    if (tid == 0) A.n_scm[r] = n_emitted;
}

int scan_geometry(int k, int s, int nt, ScanGeom *g, size_t *smem)
{
    const int q = k - s + 1;
    int n_full = q / 16 - 1;
    if (n_full < 0) n_full = 0;
    int T = 0;
    while (n_full > 0 && (4 << (2 * T)) <= n_full) ++T;  // largest T with 4^T <= n_full
    int need = (q + 15) / 16 + nt + 2, rch = 64;
    while (rch < need) rch <<= 1;
    g->rch = rch; g->n_full = n_full; g->T = T; g->h31 = h31_consts();
    *smem = sizeof(uint32_t) * ((size_t) 16 * rch + (size_t) (T + 1) * rch + COCAP + (nt / 32 + 1) + (nt / 32 + 4));
    return *smem <= 227 * 1024 ? 0 : SG_E_KSIZE;
}

template <int NT>
static int launch_scan_nt(const ScanArgs &A, uint64_t n_reads, cudaStream_t st)
{
    ScanGeom g;
    size_t smem;
    if (scan_geometry(A.k, A.s, NT, &g, &smem)) return SG_E_KSIZE;
    if (n_reads == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(scan_kernel<NT, 31>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return SG_E_CUDA;
        if (cudaFuncSetAttribute(scan_kernel<NT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return SG_E_CUDA;
        attr_set = true;
    }
    if (A.s == 31) scan_kernel<NT, 31><<<(unsigned) n_reads, NT, smem, st>>>(A, g);
    else scan_kernel<NT, 0><<<(unsigned) n_reads, NT, smem, st>>>(A, g);
    return 1;
}

int launch_scan(const ScanArgs &A, uint64_t n_reads, cudaStream_t st)
{
    return launch_scan_nt<SYNC_SCAN_NT>(A, n_reads, st);
}

} // namespace sg
