// sg_scan.cu -- kernel 1b: rolling s-mer hash, k-window minimum, closed-syncmer selection.
//
// Replaces the minimiser/emission part of the reference's per-read loop (reference
// syncmer.c:276-283, 307-394) with the stateless rules of oracle/sync_oracle.c:
//
//   m[p]     hash64 of the canonical s-mer ending at hoco position p, or NONE
//   mo(p)    min m[p-q+1 .. p-1],  e(p) = m[p-q],  q = k-s+1
//   CLOSE(p) m[p] valid, l[p] >= k, m[p] <= mo(p) and (m[p] <= e(p) or m[p] < mo(p) or m[p-q+1] == m[p])
//   OPEN(p)  e(p) valid, e(p) <= mo(p), l[p-1] >= k and (p == H or base p unambiguous)
//   start t emits iff CLOSE(t+k-1) xor OPEN(t+k)
//
// One CTA per read walks it in tiles of NT*16 positions; thread t owns 16
// consecutive positions (one 32-bit word of packed bases).
//   1. roll both strands through the 16 bases, hash, store (hi, lo) words of m[]
//      in a shared-memory ring that always holds the last q + tile positions
//      (transposed [16][chunks] so that every access is bank-conflict free)
//   2. sparse-table doubling over the per-chunk minima of the HIGH words gives
//      every thread r0 = min over the chunks that lie fully inside the window of
//      all its 16 positions
//   3. a position is a candidate when its high word (or that of e(p)) is <= the
//      running minimum of r0 and the thread's own earlier positions; about 2/q
//      of all positions pass
//   4. candidates are settled exactly: the < 32 window positions not covered by
//      r0 are scanned on the high word; a tie on the high word (only identical
//      s-mers in practice, i.e. tandem repeats) falls back to a full 64-bit scan
//   5. CLOSE/OPEN bits are combined, ranked with a block scan and written as
//      (sid, idx, m_pos, s_mer) records; k-mer hashes follow in sg_kmer.cu
#include "sg_common.cuh"
#include "sg_internal.h"
#include "../../include/syncgpu.h"

namespace sg {

template <int NT>
__global__ void __launch_bounds__(NT) scan_kernel(ScanArgs A, ScanGeom G)
{
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) uint32_t smem[];
    const int RCH = G.rch, RM = RCH - 1;
    uint32_t *ring_hi = smem;                          // [16][RCH]
    uint32_t *ring_lo = ring_hi + 16 * RCH;            // [16][RCH]
    uint32_t *D = ring_lo + 16 * RCH;                  // [J+1][RCH] (at least one level)
    const int nlev = G.J >= 0 ? G.J + 1 : 1;
    uint32_t *cflag = D + nlev * RCH;                  // [RCH]
    uint32_t *s_scan = cflag + RCH;                    // [NW + 1]
    int *s_misc = reinterpret_cast<int *>(s_scan + NW + 1);   // [NW + 4]

    const uint64_t r = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t H = A.hoco_l[r];
    const int k = A.k, s = A.s, q = k - s + 1;
    if (H < k) { if (tid == 0) A.n_scm[r] = 0; return; }
    const uint64_t hb = A.hoff[r];
    const uint32_t *hs32 = reinterpret_cast<const uint32_t *>(A.hoco_s + hb / 4);
    const uint16_t *nb16 = reinterpret_cast<const uint16_t *>(A.nbits + hb / 8);
    const int64_t nwords = (H + 15) >> 4;
    const bool has_n = A.n_amb[r] != 0;
    const uint64_t mask = (1ull << (2 * s)) - 1;
    const int rsh = 2 * s - 2;

    for (int i = tid; i < 16 * RCH; i += NT) { ring_hi[i] = 0xffffffffu; ring_lo[i] = 0xffffffffu; }
    for (int i = tid; i < nlev * RCH; i += NT) D[i] = 0xffffffffu;
    for (int i = tid; i < RCH; i += NT) cflag[i] = 0;
    if (tid == 0) s_misc[NW] = -1;                     // last ambiguous position seen so far
    __syncthreads();

    auto ring_at = [&](int64_t x) -> int { return (int) (x & 15) * RCH + (int) ((x >> 4) & RM); };
    auto m_at = [&](int64_t x) -> uint64_t { int a = ring_at(x); return (uint64_t) ring_hi[a] << 32 | ring_lo[a]; };

    uint32_t n_emitted = 0;
    const int64_t n_sub = (H + 1 + NT * 16 - 1) / (NT * 16);
    for (int64_t sub = 0; sub < n_sub; ++sub) {
        const int64_t c = sub * NT + tid;              // my chunk
        const int64_t P = c << 4;                      // its first position
        const int cs = (int) (c & RM);
        const uint32_t w0 = hoco_word(hs32, c, nwords);
        const uint32_t nb = (has_n && c < nwords) ? nb16[c] : 0u;

        // valid bases in a row ending just before my chunk
        int64_t l0 = P;
        if (has_n) {
            int mine = nb ? (int) P + 31 - __clz(nb) : -1, inc = mine;
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
        }

        // 1. hashes of my 16 positions
        uint32_t hi[16];
        uint32_t cmin = 0xffffffffu;
        {
            const uint64_t V = (uint64_t) hoco_word(hs32, c - 2, nwords) << 32 | hoco_word(hs32, c - 1, nwords);
            uint64_t fw = V & mask, rv = rc64(V) >> (64 - 2 * s);
            int64_t l = l0;
            const int nvalid = (int) min((int64_t) 16, max((int64_t) 0, H - P));
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint32_t b = (w0 >> (30 - 2 * i)) & 3u;
                fw = ((fw << 2) | b) & mask;
                rv = (rv >> 2) | ((uint64_t) (3u - b) << rsh);
                l = ((nb >> i) & 1u) ? 0 : l + 1;
                const bool ok = i < nvalid && l >= s && fw != rv;
                const uint64_t m = ok ? hash64(fw < rv ? fw : rv, mask) : SG_NONE64;
                hi[i] = (uint32_t) (m >> 32);
                ring_hi[i * RCH + cs] = hi[i];
                ring_lo[i * RCH + cs] = (uint32_t) m;
                cmin = min(cmin, hi[i]);
            }
        }
        D[cs] = cmin;
        __syncthreads();

        // 2. sparse table over chunk minima
        {
            uint32_t v = cmin;
            for (int j = 1; j <= G.J; ++j) {
                v = min(v, D[(j - 1) * RCH + (int) ((c - (1 << (j - 1))) & RM)]);
                D[j * RCH + cs] = v;
                __syncthreads();
            }
        }
        uint32_t R = 0xffffffffu;
        if (G.n_full > 0)
            R = min(D[G.J * RCH + (int) ((c - 1) & RM)], D[G.J * RCH + (int) ((c - G.n_full + (1 << G.J) - 1) & RM)]);

        // valid-run length ending at position P+i
        auto run_len = [&](int i) -> int64_t {
            if (i < 0) return l0;
            const uint32_t ml = nb & ((2u << i) - 1u);
            return ml ? (int64_t) (i - (31 - __clz(ml))) : l0 + i + 1;
        };
        const int64_t first_cov = (G.n_full > 0 ? (c - G.n_full) : c) << 4;   // first position covered by R at i = 0
        // for q < 16 a thread's earlier positions fall out of the window, so the running
        // minimum is not a bound: every position is settled by scanning its (short) window
        const bool small_q = q < 16;

        // 3./4. candidates
        uint32_t Cm = 0, Om = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int64_t p = P + i;
            const uint32_t e_hi = ring_hi[ring_at(p - q)];
            const bool cand_o = p <= H && p >= k && e_hi != 0xffffffffu && e_hi <= R;
            const bool cand_c = p < H && hi[i] != 0xffffffffu && hi[i] <= R;
            if (cand_o || cand_c) {
                // exact minimum of the high words over m[p-q+1 .. p-1]
                uint32_t Mhi = R;
                for (int64_t x = p - q + 1, xe = small_q ? p : first_cov; x < xe; ++x) Mhi = min(Mhi, ring_hi[ring_at(x)]);
                uint64_t mo = 0; bool have_mo = false;
                auto full_min = [&]() {
                    if (!have_mo) {
                        mo = SG_NONE64;
                        for (int64_t x = p - q + 1; x < p; ++x) mo = min(mo, m_at(x));
                        have_mo = true;
                    }
                    return mo;
                };
                if (cand_o && run_len(i - 1) >= k && (p == H || !((nb >> i) & 1u))) {
                    bool yes = e_hi < Mhi;
                    if (!yes && e_hi == Mhi) yes = m_at(p - q) <= full_min();
                    if (yes) Om |= 1u << i;
                }
                if (cand_c && run_len(i) >= k) {
                    bool yes = hi[i] < Mhi;
                    if (!yes && hi[i] == Mhi) {
                        const uint64_t mp = m_at(p), mm = full_min();
                        yes = mp <= mm && (mp <= m_at(p - q) || mp < mm || m_at(p - q + 1) == mp);
                    }
                    if (yes) Cm |= 1u << i;
                }
            }
            if (!small_q) R = min(R, hi[i]);
        }

        // 5. combine, rank, write
        cflag[cs] = (Cm >> 15) & 1u;
        __syncthreads();
        uint32_t E = (((Cm << 1) | cflag[(int) ((c - 1) & RM)]) ^ Om) & 0xffffu;
        uint32_t tot;
        const uint32_t ex = BlockScanU32::run<NW>(__popc(E), s_scan, &tot);
        if (tot) {
            if (tid == 0) {
                unsigned long long b = atomicAdd(A.rec_count, (unsigned long long) tot);
                s_misc[NW + 1] = (int) (uint32_t) b;
                s_misc[NW + 2] = (int) (uint32_t) (b >> 32);
            }
            __syncthreads();
            const uint64_t base = (uint64_t) (uint32_t) s_misc[NW + 1] | (uint64_t) (uint32_t) s_misc[NW + 2] << 32;
            uint32_t j = ex;
            while (E) {
                const int i = __ffs(E) - 1;
                E &= E - 1;
                const int64_t t = P + i - k;           // k-mer start
                uint64_t code;
                if ((Om >> i) & 1u) code = smer_code_at(hs32, t + s - 1, s, nwords);          // first s-mer
                else code = smer_code_at(hs32, t + k - 1, s, nwords) ^ 1ull;                 // last s-mer, flipped
                const uint32_t z = (uint32_t) (((Om >> i) & 1u) ? code & 1ull : (code ^ 1ull) & 1ull);
                const uint64_t o = base + j;
                if (o < A.rec_cap) {
                    A.rec_sid[o] = (uint32_t) r;
                    A.rec_idx[o] = n_emitted + j;
                    A.rec_mpos[o] = (uint32_t) t << 1 | z;
                    A.rec_smer[o] = code;
                }
                ++j;
            }
            n_emitted += tot;
            __syncthreads();
        }
    }
    if (tid == 0) A.n_scm[r] = n_emitted;
}

int scan_geometry(int k, int s, int nt, ScanGeom *g, size_t *smem)
{
    const int q = k - s + 1;
    int n_full = q / 16 - 1;
    if (n_full < 0) n_full = 0;
    int J = -1;
    while (n_full > 0 && (2 << J) <= n_full) ++J;        // largest J with 2^J <= n_full
    int need = (q + 15) / 16 + nt + 2, rch = 64;
    while (rch < need) rch <<= 1;
    const int nlev = J >= 0 ? J + 1 : 1;
    g->rch = rch; g->n_full = n_full; g->J = J;
    *smem = sizeof(uint32_t) * ((size_t) 32 * rch + (size_t) nlev * rch + rch + (nt / 32 + 1) + (nt / 32 + 4));
    return *smem <= 227 * 1024 ? 0 : SG_E_KSIZE;
}

int launch_scan(const ScanArgs &A, uint64_t n_reads, cudaStream_t st)
{
    constexpr int NT = 128;
    ScanGeom g;
    size_t smem;
    if (scan_geometry(A.k, A.s, NT, &g, &smem)) return SG_E_KSIZE;
    if (n_reads == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(scan_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return SG_E_CUDA;
        attr_set = true;
    }
    scan_kernel<NT><<<(unsigned) n_reads, NT, smem, st>>>(A, g);
    return 1;
}

} // namespace sg
