// sg_count.cu -- kernel group 2: the syncmer database (a6) and the counting part of
// sr_db_stat (a5).
//
// Reference: collect_syncmer_from_reads + process_kmer_cluster (syncmer.c:1397-1451,
// 1270-1393) sorts 128-bit tuples hash<<64 | sid<<32 | idx<<1 | rev, forms one
// cluster per hash, splits a cluster by exact sequence comparison when it holds
// more than one tuple, and numbers the resulting classes 0..U-1 in that order.
// sr_db_stat (syncmer.c:867-987) sorts all syncmers twice to tabulate how many
// distinct s-mer codes / k-mer keys occur once, twice, ...
//
// Here: the tuples exist in (sid, idx) order, so one stable radix sort by hash
// gives the reference's order. Every tuple is then compared, 2-bit-packed word by
// word, with its predecessor of equal hash (a warp walks 31 tuples plus one of
// overlap and hands each lane its neighbour's words by shuffle, so every k-mer is
// unpacked once). Equal neighbours everywhere means one class per hash; a hash
// group with an unequal pair is re-classified serially like the reference does.
// Ids are a prefix sum over class heads.
#include "sg_common.cuh"
#include "sg_internal.h"
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include "sg_host.h"
#include "sg_table.cuh"

namespace sg {

__global__ void __launch_bounds__(256) run_heads_kernel(const uint64_t *k, uint64_t n, int shift, uint32_t *head);

__global__ void __launch_bounds__(256) tuple_init_kernel(const uint64_t *key, uint64_t *skey, uint64_t *sval, uint64_t n, uint64_t hmask)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { skey[i] = key[i] & hmask; sval[i] = i; }
}

// ---- partial sort + repair ------------------------------------------------------------------
// The k-mer hashes are uniform 64-bit values, so five radix passes over bits 24..63 already put almost
// every tuple in its final place: what is left are the few runs that agree on those 40 bits and differ
// below (expected N^2 / 2^41 pairs, a few hundred at 21 M tuples). sort_detect_kernel lists the adjacent
// pairs that are out of order inside such a run, sort_repair_kernel insertion-sorts each listed run once
// (stable: strict comparisons, and the radix passes kept the input order inside a run). Anything the
// small fixed buffers cannot hold is reported, and the caller then sorts on all 64 bits instead.
constexpr uint32_t SORT_FIX_CAP = 8192;       // listed pairs
constexpr uint64_t SORT_FIX_MAXRUN = 4096;    // longest run a single thread is allowed to repair

// Lists, for every run (maximal stretch of equal top bits) that is out of order, the position of its FIRST inversion.
// Ownership is decided here, on data nobody is writing: a thread that sees an inversion walks back through its run and
// stays silent if an earlier inversion exists. Runs are disjoint, so the repair threads never touch the same element.
__global__ void __launch_bounds__(256) sort_detect_kernel(const uint64_t *k, uint64_t n, int SORT_LOW_BITS, uint32_t *fix /* [0] count, [1] overflow, [2..] positions */)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    const uint64_t a = k[i], b = k[i + 1];
    if ((a >> SORT_LOW_BITS) != (b >> SORT_LOW_BITS) || a <= b) return;
    const uint64_t top = a >> SORT_LOW_BITS;
    for (uint64_t j = i; j > 0 && (k[j - 1] >> SORT_LOW_BITS) == top; --j) {
        if (k[j - 1] > k[j]) return;                       // an earlier inversion owns the run
        if (i - j > SORT_FIX_MAXRUN) { fix[1] = 1u; return; }
    }
    const uint32_t o = atomicAdd(&fix[0], 1u);
    if (o < SORT_FIX_CAP) fix[2 + o] = (uint32_t) i; else fix[1] = 1u;
}

// any inversion left inside a run? (run after the repair: its answer decides whether the partial sort stands)
__global__ void __launch_bounds__(256) sort_check_kernel(const uint64_t *k, uint64_t n, int SORT_LOW_BITS, uint32_t *fix)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    const uint64_t a = k[i], b = k[i + 1];
    if ((a >> SORT_LOW_BITS) == (b >> SORT_LOW_BITS) && a > b) fix[1] = 1u;
}

__global__ void __launch_bounds__(128) sort_repair_kernel(uint64_t *k, uint64_t *v, uint64_t n, int SORT_LOW_BITS, uint32_t *fix)
{
    const uint32_t cnt = min(fix[0], SORT_FIX_CAP);
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < cnt; e += gridDim.x * blockDim.x) {
        const uint64_t i = fix[2 + e];
        const uint64_t top = k[i] >> SORT_LOW_BITS;
        uint64_t s0 = i, s1 = i + 1;                       // the run: everything before i is in order already
        while (s0 > 0 && (k[s0 - 1] >> SORT_LOW_BITS) == top) --s0;
        while (s1 + 1 < n && (k[s1 + 1] >> SORT_LOW_BITS) == top) ++s1;
        if (s1 - s0 > SORT_FIX_MAXRUN) { fix[1] = 1u; continue; }
        for (uint64_t a = i + 1; a <= s1; ++a) {           // stable insertion sort of [s0, s1]
            const uint64_t kk = k[a], vv = v[a];
            uint64_t b = a;
            while (b > s0 && k[b - 1] > kk) { k[b] = k[b - 1]; v[b] = v[b - 1]; --b; }
            k[b] = kk; v[b] = vv;
        }
    }
}

// ---- packed sort: the top 32 hash bits and the tuple index share one 64-bit word -----------------------------
// Tuples exist in (sid, idx) order and a stable sort keeps that order among equal keys, so sorting the words
// (hash >> 32) << 32 | index on their top half alone -- four passes over 8-byte words instead of five over
// 16-byte pairs -- leaves every tuple in its final place except inside the runs that agree on 32 hash bits
// (N^2 / 2^33 pairs: a few 10^4 at 21 M tuples). The gather then brings the whole 32-byte record (hash, occurrence,
// s-mer code, fingerprint) of every word in, and the detect / repair pair below puts the few runs right.
// Which hash bits the packed sort orders: the top ones -- of the hash itself, or, for tuples adopted from other GPUs, of the
// hash's place inside this GPU's hash range ((h - sub) << lsh: every hash of the range shares its leading bits, which
// would otherwise use up the word's 32 bits and leave eight times as many runs to the repair at eight GPUs).
struct RunKey { uint64_t sub; int lsh, low_bits; };
__device__ __forceinline__ uint64_t run_of(const RunKey &K, uint64_t h) { return ((h - K.sub) << K.lsh) >> K.low_bits; }

__global__ void __launch_bounds__(256) pack_init_kernel(const ulonglong4 *tup, uint64_t *pk, uint64_t n, RunKey K)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pk[i] = run_of(K, tup[i].w) << 32 | i;            // low_bits = 32: the top half stays where it is
}

__global__ void __launch_bounds__(256) sort_check_rk_kernel(const uint64_t *k, uint64_t n, RunKey K, uint32_t *fix)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    const uint64_t a = k[i], b = k[i + 1];
    if (run_of(K, a) == run_of(K, b) && a > b) fix[1] = 1u;
}

__global__ void __launch_bounds__(256) pack_gather_kernel(const uint64_t *pk, const ulonglong4 *tup, uint64_t *skey, uint64_t *sval, uint64_t *socc,
        uint64_t *ssmer, uint64_t *sfp, uint64_t n)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint64_t idx = pk[i] & 0xffffffffull;
        const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(tup + idx);
        const ulonglong2 a = __ldg(p), b = __ldg(p + 1);
        skey[i] = b.y; sval[i] = idx; socc[i] = a.x; ssmer[i] = a.y; sfp[i] = b.x;
    }
}

// The runs that need it, put right over all five gathered arrays: one WARP per run. A run holds the occurrences of the few
// distinct k-mers that share their top hash bits, each k-mer's occurrences already in tuple order but interleaved with
// the others' -- and a genomic k-mer brings its whole coverage along: hundreds of copies on one GPU, eight times that when
// eight GPUs send theirs, so nothing here may walk a run one element at a time. Three launches:
//   sort_detect_all_kernel   lists EVERY adjacent out-of-order pair inside a run;
//   sort_owner_kernel        keeps, of the pairs of one run, the first (a warp looks back 32 tuples at a time for an earlier
//                            pair or the start of the run); it only reads, so ownership never depends on a repair in flight;
//   sort_repair5_kernel      finds the run's ends 32 tuples at a time, then takes the distinct hashes in increasing order:
//                            one pass for the next smallest hash, one that moves its tuples (ballot + prefix count: stable)
//                            behind those already placed, in scratch arrays; the run is copied back at the end.
struct Repair5 { uint64_t *k, *v, *o, *sm, *fp, *tk, *tv, *to, *tsm, *tfp; };
constexpr uint32_t NOT_OWNER = 0xffffffffu;                 // a listed index is below 2^32 - 1

__global__ void __launch_bounds__(256) sort_detect_all_kernel(const uint64_t *k, uint64_t n, RunKey K, uint32_t cap, uint32_t *fix)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    const uint64_t a = k[i], b = k[i + 1];
    if (run_of(K, a) != run_of(K, b) || a <= b) return;
    const uint32_t o = atomicAdd(&fix[0], 1u);
    if (o < cap) fix[2 + o] = (uint32_t) i; else fix[1] = 1u;
}

// how many of the tuples before `hi` (exclusive, going down) belong to the run `top`, 32 at a time; *stop: an out-of-order pair
// was seen among them (only looked for when find_pair)
__device__ __forceinline__ uint64_t run_back(const uint64_t *k, RunKey K, uint64_t top, uint64_t hi, int lane, bool find_pair, bool *pair)
{
    uint64_t cnt = 0;
    *pair = false;
    for (;;) {
        const bool ok = hi - cnt > (uint64_t) lane;             // j = hi - cnt - 1 - lane exists
        const uint64_t j = hi - cnt - 1 - lane;
        const uint64_t kj = ok ? k[j] : 0;
        const bool in = ok && run_of(K, kj) == top;
        const unsigned m = __ballot_sync(SG_FULL, in);
        const int t = m == SG_FULL ? 32 : __ffs(~m) - 1;         // the run reaches this far back
        if (find_pair) {
            const bool inv = lane < t && kj > k[j + 1];
            if (__any_sync(SG_FULL, inv)) { *pair = true; return cnt; }
        }
        cnt += t;
        if (t < 32 || cnt > SORT_FIX_MAXRUN) return cnt;
    }
}

__global__ void __launch_bounds__(128) sort_owner_kernel(const uint64_t *k, RunKey K, uint32_t cap, uint32_t *fix)
{
    const uint32_t cnt = min(fix[0], cap);
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t e = warp; e < cnt; e += nwarp) {
        const uint64_t i = fix[2 + e];
        bool pair;
        const uint64_t back = run_back(k, K, run_of(K, k[i]), i, lane, true, &pair);
        if (lane == 0) {
            if (pair) fix[2 + e] = NOT_OWNER;
            else if (back > SORT_FIX_MAXRUN) fix[1] = 1u;
        }
    }
}

__global__ void __launch_bounds__(128) sort_repair5_kernel(Repair5 R, uint64_t n, RunKey K, uint32_t cap, uint32_t *fix)
{
    const uint32_t cnt = min(fix[0], cap);
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t e = warp; e < cnt; e += nwarp) {
        const uint32_t ent = fix[2 + e];
        if (ent == NOT_OWNER) continue;
        const uint64_t i = ent;
        const uint64_t top = run_of(K, R.k[i]);
        bool unused;
        const uint64_t s0 = i - run_back(R.k, K, top, i, lane, false, &unused);
        uint64_t s1 = i + 1;
        for (;;) {                                              // forward, 32 at a time
            const uint64_t j = s1 + 1 + lane;
            const bool in = j < n && run_of(K, R.k[j]) == top;
            const unsigned m = __ballot_sync(SG_FULL, in);
            const int t = m == SG_FULL ? 32 : __ffs(~m) - 1;
            s1 += t;
            if (t < 32 || s1 - s0 > SORT_FIX_MAXRUN) break;
        }
        if (s1 - s0 > SORT_FIX_MAXRUN) { if (lane == 0) fix[1] = 1u; continue; }
        uint64_t placed = s0, prev = 0;
        bool first = true;
        while (placed <= s1) {
            uint64_t cur = ~0ull;                               // the smallest hash not placed yet
            for (uint64_t a0 = s0; a0 <= s1; a0 += 32) {
                const uint64_t a = a0 + lane;
                if (a <= s1) { const uint64_t ka = R.k[a]; if (first || ka > prev) cur = min(cur, ka); }
            }
            for (int d = 16; d; d >>= 1) cur = min(cur, __shfl_xor_sync(SG_FULL, cur, d));
            for (uint64_t a0 = s0; a0 <= s1; a0 += 32) {       // its tuples, in the order they stand
                const uint64_t a = a0 + lane;
                const bool eq = a <= s1 && R.k[a] == cur;
                const unsigned m = __ballot_sync(SG_FULL, eq);
                if (eq) {
                    const uint64_t d = placed + __popc(m & ((1u << lane) - 1u));
                    R.tk[d] = cur; R.tv[d] = R.v[a]; R.to[d] = R.o[a]; R.tsm[d] = R.sm[a]; R.tfp[d] = R.fp[a];
                }
                placed += __popc(m);
            }
            prev = cur; first = false;
        }
        __syncwarp();
        for (uint64_t a = s0 + lane; a <= s1; a += 32) { R.k[a] = R.tk[a]; R.v[a] = R.tv[a]; R.o[a] = R.to[a]; R.sm[a] = R.tsm[a]; R.fp[a] = R.tfp[a]; }
        __syncwarp();
    }
}

// the same from the 32-byte records written by kmerhash_kernel: one sector per tuple
__global__ void __launch_bounds__(256) tuple_gather_aos_kernel(const uint64_t *sval, const ulonglong4 *tup, uint64_t *socc, uint64_t *ssmer, uint64_t *sfp, uint64_t n)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(tup + sval[i]);
        const ulonglong2 a = __ldg(p), b = __ldg(p + 1);
        socc[i] = a.x; ssmer[i] = a.y; sfp[i] = b.x;
    }
}

__global__ void __launch_bounds__(256) tuple_gather_kernel(const uint64_t *sval, const uint64_t *occ, const uint64_t *smer, const uint64_t *fp,
        uint64_t *socc, uint64_t *ssmer, uint64_t *sfp, uint64_t n)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const uint64_t o = sval[i]; socc[i] = occ[o]; ssmer[i] = smer[o]; sfp[i] = fp[o]; }
}

// default group test: a tuple continues its predecessor's class when hash AND fingerprint agree
__global__ void __launch_bounds__(256) verify_fp_kernel(const uint64_t *skey, const uint64_t *sfp, uint64_t n, uint32_t *newid,
        uint8_t *differs, unsigned long long *status)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool same = i > 0 && skey[i] == skey[i - 1];
    const bool neq = same && sfp[i] != sfp[i - 1];
    newid[i] = same ? 0u : 1u;
    differs[i] = neq ? 1 : 0;
    if (neq) atomicAdd(status + 1, 1ull);
}

struct VerifyArgs {
    const uint64_t *skey, *sval, *socc;
    const uint32_t *m_pos;
    const uint64_t *hoff;
    const uint8_t *hoco_s;
    const uint32_t *hoco_l;
    uint64_t sid_base, n;
    int k;
    uint32_t *newid;               // 1 for the first tuple of a hash group
    uint8_t *differs;              // 1 when the tuple's k-mer differs from its predecessor of equal hash
    unsigned long long *status;    // [1] += number of differing neighbours
};

__global__ void __launch_bounds__(256) verify_kernel(VerifyArgs A)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t e = (int64_t) (warp * 31) + lane - 1;        // lane 0 repeats the previous warp's last tuple
    const bool ok = e >= 0 && (uint64_t) e < A.n;
    uint64_t key = 0;
    const uint32_t *hs = nullptr;
    int64_t nwords = 0, start = 0;
    int rev = 0;
    if (ok) {
        key = A.skey[e];
        const uint64_t oc = A.socc[e];
        const uint64_t sid = (oc >> 32) - A.sid_base;
        hs = reinterpret_cast<const uint32_t *>(A.hoco_s + A.hoff[sid] / 4);
        nwords = ((int64_t) A.hoco_l[sid] + 15) >> 4;
        start = A.m_pos[A.sval[e]] >> 1;
        rev = (int) (oc & 1ull);
    }
    const uint64_t pkey = __shfl_up_sync(SG_FULL, key, 1);
    const bool pok = __shfl_up_sync(SG_FULL, (int) ok, 1) != 0;
    const bool same = ok && lane > 0 && pok && pkey == key;
    const int nblk = (A.k + 31) >> 5;
    bool neq = false;
    for (int j = 0; j < nblk; ++j) {
        const uint64_t mine = ok ? oriented_block(hs, nwords, start, A.k, rev, j) : 0ull;
        const uint64_t prev = __shfl_up_sync(SG_FULL, mine, 1);
        neq |= same && mine != prev;
    }
    if (ok && lane > 0) {
        A.newid[e] = same ? 0u : 1u;
        A.differs[e] = neq ? 1 : 0;
        if (neq) atomicAdd(A.status + 1, 1ull);
    }
}

// serial re-classification of the (practically non-existent) hash groups that hold
// more than one distinct k-mer; classes are numbered by first appearance and laid
// out one after another, each in tuple order (syncmer.c:1283-1333, 1342-1379)
struct SplitArgs {
    uint64_t *skey, *sval, *socc, *ssmer;
    uint64_t *t_val, *t_occ, *t_smer;     // scratch of the same size
    uint32_t *cls;                        // scratch, one per tuple
    const uint32_t *m_pos;
    const uint64_t *hoff;
    const uint8_t *hoco_s;
    const uint32_t *hoco_l;
    uint64_t sid_base, n;
    int k;
    uint32_t *newid;
    const uint8_t *differs;
    unsigned long long *status;           // [2] += groups split
};

__device__ bool same_kmer(const SplitArgs &A, uint64_t a, uint64_t b)
{
    const uint64_t oa = A.socc[a], ob = A.socc[b];
    const uint64_t sa = (oa >> 32) - A.sid_base, sb = (ob >> 32) - A.sid_base;
    const uint32_t *ha = reinterpret_cast<const uint32_t *>(A.hoco_s + A.hoff[sa] / 4);
    const uint32_t *hb = reinterpret_cast<const uint32_t *>(A.hoco_s + A.hoff[sb] / 4);
    const int64_t na = ((int64_t) A.hoco_l[sa] + 15) >> 4, nb = ((int64_t) A.hoco_l[sb] + 15) >> 4;
    const int64_t pa = A.m_pos[A.sval[a]] >> 1, pb = A.m_pos[A.sval[b]] >> 1;
    const int nblk = (A.k + 31) >> 5;
    for (int j = 0; j < nblk; ++j)
        if (oriented_block(ha, na, pa, A.k, (int) (oa & 1), j) != oriented_block(hb, nb, pb, A.k, (int) (ob & 1), j)) return false;
    return true;
}

__global__ void __launch_bounds__(128) split_kernel(SplitArgs A)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n || (i > 0 && A.skey[i] == A.skey[i - 1])) return;      // hash-group heads only (skey is never rewritten)
    uint64_t j = i + 1;
    bool any = false;
    while (j < A.n && A.skey[j] == A.skey[i]) { any |= A.differs[j] != 0; ++j; }
    if (!any) return;
    atomicAdd(A.status + 2, 1ull);
    // 1. classes by first appearance; t_val[i + c] remembers the tuple that represents class c
    uint32_t ncls = 0;
    for (uint64_t x = i; x < j; ++x) {
        uint32_t c = 0;
        for (; c < ncls; ++c) if (same_kmer(A, x, A.t_val[i + c])) break;
        if (c == ncls) A.t_val[i + ncls++] = x;
        A.cls[x] = c;
    }
    // 2. stable partition by class into the scratch arrays (the representatives are no longer needed);
    //    bit 63 of the permuted sval marks the first tuple of a class
    uint64_t w = i;
    for (uint32_t c = 0; c < ncls; ++c) {
        bool first = true;
        for (uint64_t x = i; x < j; ++x) {
            if (A.cls[x] != c) continue;
            A.t_occ[w] = A.socc[x];
            A.t_smer[w] = A.ssmer[x];
            A.t_val[w] = A.sval[x] | (first ? 1ull << 63 : 0ull);
            first = false;
            ++w;
        }
    }
    // 3. copy back
    for (uint64_t x = i; x < j; ++x) {
        const uint64_t v = A.t_val[x];
        A.sval[x] = v & ~(1ull << 63);
        A.newid[x] = (uint32_t) (v >> 63);
        A.socc[x] = A.t_occ[x];
        A.ssmer[x] = A.t_smer[x];
    }
}

struct FillArgs {
    const uint64_t *skey, *sval, *socc, *ssmer;
    const uint32_t *newid;
    const uint64_t *ex;                  // exclusive scan of newid (n + 1)
    uint64_t n;
    uint64_t *kid;                       // read order: id << 1
    uint64_t *scm_h, *scm_s, *occ_off;
    unsigned long long *status;          // [0] |= s-mer conflict, [3] = n - (first tuple in hash order that disagrees with its class)
};

__global__ void __launch_bounds__(256) fill_kernel(FillArgs A)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    const uint32_t head = A.newid[i];
    const uint64_t id = A.ex[i] + head - 1;
    A.kid[A.sval[i]] = id << 1;                                  // syncmer.c:1378
    if (head) { A.scm_h[id] = A.skey[i]; A.scm_s[id] = A.ssmer[i]; A.occ_off[id] = i; }
    else if (A.ssmer[i] != A.ssmer[i - 1]) { atomicOr(A.status, 1ull); atomicMax(A.status + 3, (unsigned long long) (A.n - i)); }   // syncmer.c:1370-1376
    if (i == A.n - 1) A.occ_off[id + 1] = A.n;
}

// what the reference prints before it gives up (syncmer.c:1371-1374): the class's hash, the s-mer code and read of its first
// tuple, the code and read of the first tuple that disagrees
__global__ void conflict_info_kernel(const uint64_t *skey, const uint64_t *ssmer, const uint64_t *socc, const uint32_t *newid, uint64_t i, unsigned long long *out)
{
    uint64_t h = i;
    while (h > 0 && !newid[h]) --h;
    out[0] = skey[i]; out[1] = ssmer[h]; out[2] = socc[h] >> 32; out[3] = ssmer[i]; out[4] = socc[i] >> 32;
}

__global__ void __launch_bounds__(256) cov_kernel(const uint64_t *occ_off, uint32_t *cov, uint64_t u)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < u) cov[i] = (uint32_t) (occ_off[i + 1] - occ_off[i]);
}

// ---- a5 ----
__global__ void __launch_bounds__(256) run_heads_kernel(const uint64_t *k, uint64_t n, int shift, uint32_t *head)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || (k[i] >> shift) != (k[i - 1] >> shift)) ? 1u : 0u;
}

// one thread per run head: length = distance to the next head, found through the scan
__global__ void __launch_bounds__(256) run_starts_kernel(const uint32_t *head, const uint64_t *ex, uint64_t n, uint64_t *starts)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (head[i]) starts[ex[i]] = i;
    if (i == n - 1) starts[ex[n]] = n;
}

__global__ void __launch_bounds__(256) run_hist_kernel(const uint64_t *starts, uint64_t g, unsigned long long *hist /* 1001 */)
{
    __shared__ uint32_t h[1001];
    for (int i = threadIdx.x; i < 1001; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < g; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t len = starts[i + 1] - starts[i];
        atomicAdd(&h[len < 1000 ? len : 1000], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1001; i += blockDim.x) if (h[i]) atomicAdd(hist + i, (unsigned long long) h[i]);
}

// multiplicity of every distinct key through the warp-cooperative table
__global__ void __launch_bounds__(256) key_tally_kernel(const uint64_t *keys, uint64_t n, uint64_t *tk, uint32_t *tv, uint64_t nslot_mask)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarp = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    for (uint64_t base = warp0 * 32; base < n; base += nwarp * 32)
        table_add_warp(tk, tv, nslot_mask, base + lane < n ? keys[base + lane] : EMPTY_KEY, lane);
}

// the same in a table sized for FEW distinct keys (it then stays in L2): the lanes of a warp merge equal keys (match), then every
// leader probes for itself (slot_add_bounded); gives up when a key finds no slot nearby
__global__ void __launch_bounds__(256) key_tally_small_kernel(const uint64_t *keys, uint64_t n, uint64_t *tk, uint32_t *tv, uint64_t nslot_mask,
        unsigned int *full)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarp = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    for (uint64_t base = warp0 * 32; base < n; base += nwarp * 32) {
        const uint64_t key = base + lane < n ? keys[base + lane] : EMPTY_KEY;
        const unsigned int gave_up = lane == 0 ? *((volatile unsigned int *) full) : 0u;
        const uint32_t peers = __match_any_sync(SG_FULL, key);
        const bool leader = key != EMPTY_KEY && lane == __ffs(peers) - 1;
        if (__any_sync(SG_FULL, gave_up != 0)) return;          // somebody gave up: the result will be thrown away
        const bool lost = leader && !slot_add_bounded(tk, tv, nslot_mask, key, __popc(peers), 128);
        if (__any_sync(SG_FULL, lost)) { if (lane == 0) *full = 1u; return; }
    }
}

// multiplicity-of-multiplicity histogram over the occupied slots; hist[1001] counts the distinct keys
__global__ void __launch_bounds__(256) slot_hist_kernel(const uint64_t *tk, const uint32_t *tv, uint64_t nslots, unsigned long long *hist)
{
    __shared__ uint32_t h[1002];
    for (int i = threadIdx.x; i < 1002; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < nslots; i += (uint64_t) gridDim.x * blockDim.x) {
        if (tk[i] == EMPTY_KEY) continue;
        const uint32_t c = tv[i];
        atomicAdd(&h[c < 1000 ? c : 1000], 1u);
        atomicAdd(&h[1001], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1002; i += blockDim.x) if (h[i]) atomicAdd(hist + i, (unsigned long long) h[i]);
}

// ---- multi-GPU: s-mer multiplicities need every occurrence of a code in one table -----------------
// the occupied slots of a tally table as (code, count) pairs
__global__ void __launch_bounds__(256) slot_pack_kernel(const uint64_t *tk, const uint32_t *tv, uint64_t nslots, uint64_t *pairs, unsigned long long *n_out)
{
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < nslots; i += (uint64_t) gridDim.x * blockDim.x) {
        if (tk[i] == EMPTY_KEY) continue;
        const unsigned long long o = atomicAdd(n_out, 1ull);
        pairs[2 * o] = tk[i];
        pairs[2 * o + 1] = tv[i];
    }
}
// add the pairs of all ranks into one table
__global__ void __launch_bounds__(256) pair_tally_kernel(const uint64_t *pairs, uint64_t n, uint64_t *tk, uint32_t *tv, uint64_t nslot_mask)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarp = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    for (uint64_t base = warp0 * 32; base < n; base += nwarp * 32) {
        const bool ok = base + lane < n;
        const uint64_t key = ok ? pairs[2 * (base + lane)] : EMPTY_KEY;
        const uint32_t cnt = ok ? (uint32_t) pairs[2 * (base + lane) + 1] : 0u;
        uint32_t work = __ballot_sync(SG_FULL, ok);
        while (work) {
            const int src = __ffs(work) - 1;
            work &= work - 1;
            table_add(tk, tv, nslot_mask, __shfl_sync(SG_FULL, key, src), __shfl_sync(SG_FULL, cnt, src), lane);
        }
    }
}

__global__ void __launch_bounds__(256) gap_kernel(const uint64_t *occ, const uint32_t *m_pos, uint64_t n, int k, unsigned long long *out /* [0]=sum (two's complement), [1]=count */)
{
    __shared__ long long ssum[8];
    __shared__ unsigned long long scnt[8];
    long long s = 0; unsigned long long c = 0;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t idx = (uint32_t) (occ[i] >> 1) & 0x7FFFFFFFu;
        if (idx == 0) continue;
        const uint32_t p1 = m_pos[i] >> 1, p0 = m_pos[i - 1] >> 1;
        if (p0 == 0x7FFFFFFFu || p1 == 0x7FFFFFFFu) continue;          // error-corrected entries, syncmer.c:900
        s += (long long) p1 - (long long) p0 - k; ++c;
    }
    for (int d = 16; d; d >>= 1) { s += __shfl_down_sync(SG_FULL, s, d); c += __shfl_down_sync(SG_FULL, c, d); }
    if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = s; scnt[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { s += ssum[w]; c += scnt[w]; }
        atomicAdd(out, (unsigned long long) s);
        atomicAdd(out + 1, c);
    }
}

} // namespace sg

using namespace sg;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return SG_E_CUDA; } } while (0)
#define RS(buf, bytes) do { if ((buf).reserve(bytes)) { ctx->err = "device allocation of " + std::to_string((size_t)(bytes)) + " bytes failed"; return SG_E_NOMEM; } } while (0)
#define LAUNCHED(stage, expr) do { int n_ = (expr); if (n_ < 0) return n_; ctx->count_launch(stage, n_); } while (0)
static inline unsigned nblk(uint64_t n, unsigned t) { return (unsigned) ((n + t - 1) / t); }

// sort the tuples of the batch by hash (cached until the next extract/adopt)
static int ensure_sorted(sg_batch *b)
{
    sg_ctx *ctx = b->ctx;
    if (b->sorted) return SG_OK;
    cudaStream_t st = ctx->stream;
    const uint64_t N = b->t_n();
    RS(b->skey, (N + 1) * 8); RS(b->sval, (N + 1) * 8); RS(b->skey_alt, (N + 1) * 8); RS(b->sval_alt, (N + 1) * 8);
    RS(b->sort_tmp, sort_tmp_words(N) * 4);
    RS(b->socc, (N + 1) * 8); RS(b->ssmer, (N + 1) * 8); RS(b->sfp, (N + 1) * 8);
    ctx->t_begin(SG_T_SORT);
    const uint64_t hmask = b->hash_bits >= 64 ? ~0ull : ((1ull << b->hash_bits) - 1);
    RS(b->sort_fix, (SORT_FIX_CAP + 2) * 4);
    const int SORT_LOW_BITS = b->keys_are_ids ? 0 : b->sort_low_bits;   // 24 unless a test moves it (multiple of 8); dense ids: all bits
    bool full = b->hash_bits < 64 || SORT_LOW_BITS == 0;     // truncated hashes (tests) collide by design
    // the ordinary case: whole hashes, records of this batch's own extract, fewer than 2^32 tuples (the index shares a word
    // with the top half of the hash)
    if (!full && SORT_LOW_BITS == 24 && (b->adopted ? b->atup_valid : b->tup_valid) && !b->keys_are_ids && N < (1ull << 32) && !getenv("SG_SORT_PAIRS")) {
        const uint32_t cap = (uint32_t) std::min<uint64_t>(N + 2, 0xfffffff0u);   // every out-of-order pair inside a run is listed: fewer than N
        RS(b->sort_fix, ((size_t) cap + 2) * 4);
        uint64_t *pk = (uint64_t *) b->skey_alt.p, *pk_alt = (uint64_t *) b->sval_alt.p;
        const int plow = 64 - b->pack_bits;                   // hash bits the packed sort leaves to the repair (32 unless a test moves it)
        const RunKey RK = {b->adopted ? b->range_lo : 0ull, b->adopted ? b->range_lsh : 0, plow};
        ctx->lap("begin");
        pack_init_kernel<<<nblk(N, 256), 256, 0, st>>>((const ulonglong4 *) b->tup.p, pk, N, RK);
        ctx->count_launch(SG_T_SORT, 1);
        ctx->lap("pack_init");
        LAUNCHED(SG_T_SORT, launch_sort_keys(pk, pk_alt, N, 32, 32 + b->pack_bits, (uint32_t *) b->sort_tmp.p, st));
        ctx->lap("sort_keys");
        pack_gather_kernel<<<nblk(N, 256), 256, 0, st>>>(pk, (const ulonglong4 *) b->tup.p, (uint64_t *) b->skey.p, (uint64_t *) b->sval.p,
                (uint64_t *) b->socc.p, (uint64_t *) b->ssmer.p, (uint64_t *) b->sfp.p, N);
        ctx->lap("gather");
        CK(cudaMemsetAsync(b->sort_fix.p, 0, 8, st));
        sort_detect_all_kernel<<<nblk(N, 256), 256, 0, st>>>((const uint64_t *) b->skey.p, N, RK, cap, (uint32_t *) b->sort_fix.p);
        sort_owner_kernel<<<std::min<unsigned>(nblk((uint64_t) cap * 32, 128), 148u * 16u), 128, 0, st>>>((const uint64_t *) b->skey.p, RK, cap, (uint32_t *) b->sort_fix.p);
        // one warp per listed run (their number is known on the device only: a grid for a full list, warps without a run leave at once);
        // scratch: the two word buffers of the sort and three arrays that are filled later in the step
        RS(b->ids, (N + 2) * 8); RS(b->starts, (N + 2) * 8); RS(b->kid, (N + 1) * 8);
        ctx->lap("detect");
        Repair5 R5 = {(uint64_t *) b->skey.p, (uint64_t *) b->sval.p, (uint64_t *) b->socc.p, (uint64_t *) b->ssmer.p, (uint64_t *) b->sfp.p,
                      pk, pk_alt, (uint64_t *) b->ids.p, (uint64_t *) b->starts.p, (uint64_t *) b->kid.p};
        sort_repair5_kernel<<<std::min<unsigned>(nblk((uint64_t) cap * 32, 128), 148u * 16u), 128, 0, st>>>(R5, N, RK, cap, (uint32_t *) b->sort_fix.p);
        ctx->lap("repair");
        sort_check_rk_kernel<<<nblk(N, 256), 256, 0, st>>>((const uint64_t *) b->skey.p, N, RK, (uint32_t *) b->sort_fix.p);
        ctx->count_launch(SG_T_SORT, 5);
        ctx->lap("check");
        uint32_t hf[2];
        CK(cudaMemcpyAsync(hf, b->sort_fix.p, sizeof(hf), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        b->n_sort_repairs = hf[0];
        b->sort_fell_back = false;
        if (ctx->laps_on > 0) fprintf(stderr, "[sg laps dev %d] sort: N=%llu runs=%u bad=%u cap=%u\n", ctx->device, (unsigned long long) N, hf[0], hf[1], cap);
        ctx->laps_print("sort");
        if (hf[1] == 0 && hf[0] <= cap) {
            ctx->t_end(SG_T_SORT);
            CK(cudaGetLastError());
            b->sorted = true;
            return SG_OK;
        }
        b->sort_fell_back = true;                             // too many or too long runs: the pair sort on all 64 bits below
        full = true;
    }
    { const int rc_ = sg::ensure_adopted_soa(b); if (rc_) return rc_; }       // the pair sort reads the four arrays
    for (int attempt = 0; attempt < 2; ++attempt) {
        tuple_init_kernel<<<nblk(N, 256), 256, 0, st>>>(b->t_key(), (uint64_t *) b->skey.p, (uint64_t *) b->sval.p, N, hmask);
        ctx->count_launch(SG_T_SORT, 1);
        LAUNCHED(SG_T_SORT, launch_sort_pairs((uint64_t *) b->skey.p, (uint64_t *) b->sval.p, (uint64_t *) b->skey_alt.p,
                (uint64_t *) b->sval_alt.p, N, full ? 0 : SORT_LOW_BITS, b->hash_bits >= 64 ? 64 : ((b->hash_bits + 7) & ~7), (uint32_t *) b->sort_tmp.p, st));
        if (full) break;
        CK(cudaMemsetAsync(b->sort_fix.p, 0, 8, st));
        sort_detect_kernel<<<nblk(N, 256), 256, 0, st>>>((const uint64_t *) b->skey.p, N, SORT_LOW_BITS, (uint32_t *) b->sort_fix.p);
        sort_repair_kernel<<<8, 128, 0, st>>>((uint64_t *) b->skey.p, (uint64_t *) b->sval.p, N, SORT_LOW_BITS, (uint32_t *) b->sort_fix.p);
        sort_check_kernel<<<nblk(N, 256), 256, 0, st>>>((const uint64_t *) b->skey.p, N, SORT_LOW_BITS, (uint32_t *) b->sort_fix.p);
        ctx->count_launch(SG_T_SORT, 3);
        uint32_t hf[2];
        CK(cudaMemcpyAsync(hf, b->sort_fix.p, sizeof(hf), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        b->n_sort_repairs = hf[0];
        b->sort_fell_back = false;
        if (hf[1] == 0 && hf[0] <= SORT_FIX_CAP) break;
        b->sort_fell_back = true;
        full = true;                                          // too many or too long runs: sort on all 64 bits
    }
    if (b->tup_valid && !b->adopted && !b->keys_are_ids)
        tuple_gather_aos_kernel<<<nblk(N, 256), 256, 0, st>>>((const uint64_t *) b->sval.p, (const ulonglong4 *) b->tup.p,
                (uint64_t *) b->socc.p, (uint64_t *) b->ssmer.p, (uint64_t *) b->sfp.p, N);
    else
        tuple_gather_kernel<<<nblk(N, 256), 256, 0, st>>>((const uint64_t *) b->sval.p, b->t_occ(),
            b->t_smer(), b->t_fp(), (uint64_t *) b->socc.p, (uint64_t *) b->ssmer.p, (uint64_t *) b->sfp.p, N);
    ctx->count_launch(SG_T_SORT, 1);
    ctx->t_end(SG_T_SORT);
    CK(cudaGetLastError());
    b->sorted = true;
    return SG_OK;
}

// multiplicity-of-multiplicity table of a sorted key array (keys compared after >> shift)
static int mult_table(sg_batch *b, const uint64_t *sorted, uint64_t N, int shift, unsigned long long *d_hist, uint64_t *groups)
{
    sg_ctx *ctx = b->ctx;
    *groups = 0;
    if (N == 0) return SG_OK;
    cudaStream_t st = ctx->stream;
    RS(b->flags, (N + 1) * 4); RS(b->ids, (N + 2) * 8); RS(b->ids_tmp, scan_tmp_words(N) * 8);
    RS(b->starts, (N + 2) * 8);
    run_heads_kernel<<<nblk(N, 256), 256, 0, st>>>(sorted, N, shift, (uint32_t *) b->flags.p);
    ctx->count_launch(SG_T_STAT, 1);
    LAUNCHED(SG_T_STAT, launch_scan_u32_u64((const uint32_t *) b->flags.p, (uint64_t *) b->ids.p, N, (uint64_t *) b->ids_tmp.p, st));
    uint64_t G = 0;
    CK(cudaMemcpyAsync(&G, (uint64_t *) b->ids.p + N, 8, cudaMemcpyDeviceToHost, st));
    run_starts_kernel<<<nblk(N, 256), 256, 0, st>>>((const uint32_t *) b->flags.p, (const uint64_t *) b->ids.p, N, (uint64_t *) b->starts.p);
    CK(cudaStreamSynchronize(st));
    run_hist_kernel<<<std::min<unsigned>(nblk(G, 256), 1184u), 256, 0, st>>>((const uint64_t *) b->starts.p, G, d_hist);
    ctx->count_launch(SG_T_STAT, 2);
    *groups = G;
    return SG_OK;
}

extern "C" {

int sg_batch_set_exact_verify(sg_batch *b, int on)
{
    if (!b) return SG_E_ARG;
    b->exact_verify = on != 0;
    b->counted = false;
    return SG_OK;
}

int sg_debug_set_sort_low_bits(sg_batch *b, int bits)
{
    if (!b || bits < 0 || bits > 56 || (bits & 7)) return SG_E_ARG;
    b->sort_low_bits = bits;
    b->sorted = b->counted = false;
    return SG_OK;
}

int sg_debug_set_pack_bits(sg_batch *b, int bits)
{
    if (!b || bits < 8 || bits > 32 || (bits & 7)) return SG_E_ARG;
    b->pack_bits = bits;
    b->sorted = b->counted = false;
    return SG_OK;
}

int sg_debug_sort_info(sg_batch *b, uint64_t *repairs, int *fell_back)
{
    if (!b) return SG_E_ARG;
    if (repairs) *repairs = b->n_sort_repairs;
    if (fell_back) *fell_back = b->sort_fell_back ? 1 : 0;
    return SG_OK;
}

int sg_debug_set_hash_bits(sg_batch *b, int bits)
{
    if (!b || bits < 1 || bits > 64) return SG_E_ARG;
    b->hash_bits = bits;
    b->sorted = b->counted = false;
    return SG_OK;
}

int sg_stat(sg_batch *b, sg_stat_t *out)
{
    if (!b || !out) return SG_E_ARG;
    if (!b->extracted) return SG_E_STATE;
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof(*out));
    const uint64_t N = b->t_n();
    out->n_syncmers = b->n_syncmers;
    if (N == 0 && b->n_syncmers == 0) return SG_E_EMPTY;            // "empty syncmer collection", syncmer.c:909-912
    int rc = ensure_sorted(b);
    if (rc) return rc;
    ctx->t_begin(SG_T_STAT);
    RS(b->stat_dev, (2 * 1001 + 4) * 8);
    unsigned long long *d = (unsigned long long *) b->stat_dev.p;
    CK(cudaMemsetAsync(d, 0, (2 * 1001 + 4) * 8, st));
    // k-mer keys: hash with its low bit dropped (syncmer.c:896); the hash-sorted order keeps equal keys adjacent
    uint64_t gk = 0, gs = 0;
    rc = mult_table(b, (const uint64_t *) b->skey.p, N, 1, d + 1001, &gk);
    if (rc) return rc;
    // s-mer codes: no order is needed, only how often each distinct code occurs -> hash table
    // (the reference sorts all syncmers a second time for this, syncmer.c:916-926)
    if (N) {
        // Syncmers of one genome share few s-mers (the minimum of a window survives most sequencing errors in it), so
        // the table is first cut for N / 16 distinct codes: 48 MB at 21 M syncmers, resident in L2, where a table for
        // the worst case (every code distinct) is 768 MB of DRAM probes. A tally that finds the small table crowded
        // (a key without a free slot within 128 probes, or more than half of the slots taken) is repeated in the large one.
        uint64_t nslots = 1024, nfull = 1024;
        while (nfull < 2 * N) nfull <<= 1;
        while (nslots < N / 8) nslots <<= 1;
        if (nslots > nfull) nslots = nfull;
        RS(b->stat_dev2, 1004 * 8);
        bool tallied = false;
        for (int attempt = 0; attempt < 2; ++attempt) {
            RS(b->arc_keys, nslots * 8); RS(b->arc_vals, nslots * 4);
            b->smer_slots = nslots;
            CK(cudaMemsetAsync(b->arc_keys.p, 0xff, nslots * 8, st));
            CK(cudaMemsetAsync(b->arc_vals.p, 0, nslots * 4, st));
            CK(cudaMemsetAsync(b->stat_dev2.p, 0, 1004 * 8, st));
            // in k-mer hash order the occurrences of one k-mer (hence of its s-mer code) are neighbours: the warps merge
            // them with one match_any before touching the table
            if (nslots == nfull) {
                key_tally_kernel<<<std::min<unsigned>(nblk(N, 256), 148u * 16u), 256, 0, st>>>((const uint64_t *) b->ssmer.p, N, (uint64_t *) b->arc_keys.p,
                        (uint32_t *) b->arc_vals.p, nslots - 1);
                ctx->count_launch(SG_T_STAT, 1);
                break;
            }
            unsigned long long *flag = (unsigned long long *) b->stat_dev2.p + 1002;      // [1002] a key found no slot (low word)
            key_tally_small_kernel<<<std::min<unsigned>(nblk(N, 256), 148u * 16u), 256, 0, st>>>((const uint64_t *) b->ssmer.p, N, (uint64_t *) b->arc_keys.p,
                    (uint32_t *) b->arc_vals.p, nslots - 1, (unsigned int *) flag);
            // the histogram pass counts the keys in the table as well ([1001]): no counter is hammered during the tally
            slot_hist_kernel<<<std::min<unsigned>(nblk(nslots, 256), 148u * 16u), 256, 0, st>>>((const uint64_t *) b->arc_keys.p,
                    (const uint32_t *) b->arc_vals.p, nslots, (unsigned long long *) b->stat_dev2.p);
            ctx->count_launch(SG_T_STAT, 2);
            unsigned long long hf[2];
            CK(cudaMemcpyAsync(hf, (unsigned long long *) b->stat_dev2.p + 1001, sizeof(hf), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if ((hf[1] & 0xffffffffull) == 0 && hf[0] <= nslots / 2) { tallied = true; break; }
            nslots = nfull;
        }
        if (!tallied) {
            slot_hist_kernel<<<std::min<unsigned>(nblk(nslots, 256), 148u * 16u), 256, 0, st>>>((const uint64_t *) b->arc_keys.p,
                    (const uint32_t *) b->arc_vals.p, nslots, (unsigned long long *) b->stat_dev2.p);
            ctx->count_launch(SG_T_STAT, 1);
        }
        CK(cudaMemcpyAsync(d, b->stat_dev2.p, 1001 * 8, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(&gs, (unsigned long long *) b->stat_dev2.p + 1001, 8, cudaMemcpyDeviceToHost, st));
    }
    // gaps are a property of the local reads whichever tuple set is being counted
    if (b->n_syncmers) gap_kernel<<<std::min<unsigned>(nblk(b->n_syncmers, 256), 1184u), 256, 0, st>>>((const uint64_t *) b->occ.p, (const uint32_t *) b->m_pos.p, b->n_syncmers, b->k, d + 2002);
    ctx->count_launch(SG_T_STAT, 1);
    ctx->t_end(SG_T_STAT);
    std::vector<unsigned long long> h(2 * 1001 + 4);
    CK(cudaMemcpyAsync(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    for (int i = 0; i < 1001; ++i) { out->smer_cnts[i] = (int64_t) h[i]; out->kmer_cnts[i] = (int64_t) h[1001 + i]; }
    out->gap_sum = (int64_t) h[2002];
    out->n_gaps = h[2003];
    out->smer_unique = gs; out->smer_singleton = h[1];
    out->kmer_unique = gk; out->kmer_singleton = h[1001 + 1];
    return SG_OK;
}

// run lengths of a starts[] array, or the second word of sorted (key, count) pairs, as uint32 in key order
__global__ void __launch_bounds__(256) run_len_kernel(const uint64_t *starts, uint64_t g, uint32_t *out)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < g) out[i] = (uint32_t) (starts[i + 1] - starts[i]);
}
__global__ void __launch_bounds__(256) pair_split_kernel(const uint64_t *pairs, uint64_t n, uint64_t *key, uint64_t *val)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { key[i] = pairs[2 * i]; val[i] = pairs[2 * i + 1]; }
}
__global__ void __launch_bounds__(256) narrow_kernel(const uint64_t *val, uint64_t n, uint32_t *out)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t) val[i];
}

int sg_stat_multiplicities(sg_batch *b, int which, uint32_t **host_out, uint64_t *n_out)
{
    if (!b || !host_out || !n_out || (which != 0 && which != 1)) return SG_E_ARG;
    if (!b->extracted) return SG_E_STATE;
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    *host_out = nullptr; *n_out = 0;
    uint64_t G = 0;
    if (which == 0) {
        const uint64_t N = b->t_n();
        int rc = ensure_sorted(b);
        if (rc) return rc;
        RS(b->stat_dev2, 1002 * 8);
        CK(cudaMemsetAsync(b->stat_dev2.p, 0, 1002 * 8, st));
        rc = mult_table(b, (const uint64_t *) b->skey.p, N, 1, (unsigned long long *) b->stat_dev2.p, &G);   // leaves starts[0..G]
        if (rc) return rc;
        b->smer_slots = 0;
        if (G) {
            RS(b->flags, (G + 1) * 4);
            run_len_kernel<<<nblk(G, 256), 256, 0, st>>>((const uint64_t *) b->starts.p, G, (uint32_t *) b->flags.p);
            ctx->count_launch(SG_T_STAT, 1);
        }
    } else {
        void *pairs = nullptr;
        int rc = sg_smer_counts_pack(b, &pairs, &G);
        if (rc) return rc;
        if (G) {
            RS(b->skey2, (G + 1) * 8); RS(b->sval2, (G + 1) * 8);
            RS(b->ids, (G + 2) * 8); RS(b->starts, (G + 2) * 8);
            RS(b->sort_tmp, sort_tmp_words(G) * 4);
            pair_split_kernel<<<nblk(G, 256), 256, 0, st>>>((const uint64_t *) pairs, G, (uint64_t *) b->skey2.p, (uint64_t *) b->sval2.p);
            LAUNCHED(SG_T_STAT, launch_sort_pairs((uint64_t *) b->skey2.p, (uint64_t *) b->sval2.p, (uint64_t *) b->ids.p, (uint64_t *) b->starts.p,
                    G, 0, 64, (uint32_t *) b->sort_tmp.p, st));
            RS(b->flags, (G + 1) * 4);
            narrow_kernel<<<nblk(G, 256), 256, 0, st>>>((const uint64_t *) b->sval2.p, G, (uint32_t *) b->flags.p);
            ctx->count_launch(SG_T_STAT, 2);
        }
    }
    if (G) {
        uint32_t *h = (uint32_t *) malloc(G * sizeof(uint32_t));
        if (!h) return SG_E_NOMEM;
        CK(cudaMemcpyAsync(h, b->flags.p, G * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        b->d2h_bytes += G * 4;
        *host_out = h;
    }
    *n_out = G;
    return SG_OK;
}

int sg_smer_counts_pack(sg_batch *b, void **d_pairs, uint64_t *n)
{
    if (!b || !d_pairs || !n) return SG_E_ARG;
    if (!b->smer_slots) return SG_E_STATE;                          // sg_stat builds the table
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    RS(b->smer_pairs, (b->smer_slots / 2 + 1) * 16);                // load factor <= 0.5
    RS(b->status, 4 * 8);
    CK(cudaMemsetAsync(b->status.p, 0, 8, st));
    slot_pack_kernel<<<std::min<unsigned>(nblk(b->smer_slots, 256), 148u * 16u), 256, 0, st>>>((const uint64_t *) b->arc_keys.p,
            (const uint32_t *) b->arc_vals.p, b->smer_slots, (uint64_t *) b->smer_pairs.p, (unsigned long long *) b->status.p);
    ctx->count_launch(SG_T_STAT, 1);
    unsigned long long c = 0;
    CK(cudaMemcpyAsync(&c, b->status.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    *d_pairs = b->smer_pairs.p;
    *n = c;
    return SG_OK;
}

int sg_smer_counts_merge(sg_batch *b, const void *d_pairs, uint64_t n, sg_stat_t *out)
{
    if (!b || !out || (!d_pairs && n)) return SG_E_ARG;
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    uint64_t nslots = 1024;
    while (nslots < 2 * n) nslots <<= 1;
    RS(b->arc_keys, nslots * 8); RS(b->arc_vals, nslots * 4); RS(b->stat_dev2, 1002 * 8);
    b->smer_slots = 0;                                              // the local table is gone
    CK(cudaMemsetAsync(b->arc_keys.p, 0xff, nslots * 8, st));
    CK(cudaMemsetAsync(b->arc_vals.p, 0, nslots * 4, st));
    CK(cudaMemsetAsync(b->stat_dev2.p, 0, 1002 * 8, st));
    if (n) pair_tally_kernel<<<std::min<unsigned>(nblk(n, 256), 148u * 16u), 256, 0, st>>>((const uint64_t *) d_pairs, n, (uint64_t *) b->arc_keys.p,
            (uint32_t *) b->arc_vals.p, nslots - 1);
    slot_hist_kernel<<<std::min<unsigned>(nblk(nslots, 256), 148u * 16u), 256, 0, st>>>((const uint64_t *) b->arc_keys.p,
            (const uint32_t *) b->arc_vals.p, nslots, (unsigned long long *) b->stat_dev2.p);
    ctx->count_launch(SG_T_STAT, 2);
    unsigned long long h[1002];
    CK(cudaMemcpyAsync(h, b->stat_dev2.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    for (int i = 0; i < 1001; ++i) out->smer_cnts[i] = (int64_t) h[i];
    out->smer_unique = h[1001];
    out->smer_singleton = h[1];
    return SG_OK;
}

int sg_batch_set_lists_host(sg_batch *b, uint64_t n_reads, const uint64_t *scm_off, const uint64_t *k_mer, const uint32_t *m_pos,
        const uint64_t *s_mer, const uint32_t *cov, uint64_t n_unique)
{
    if (!b || !scm_off || (n_reads && scm_off[n_reads] && (!k_mer || !m_pos || !s_mer)) || (n_unique && !cov)) return SG_E_ARG;
    if (n_reads > 0xFFFFFFFFull) return SG_E_LIMIT;
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    const uint64_t N = n_reads ? scm_off[n_reads] : 0;
    RS(b->key, (N + 1) * 8); RS(b->kid, (N + 1) * 8); RS(b->occ, (N + 1) * 8); RS(b->m_pos, (N + 1) * 4); RS(b->s_mer, (N + 1) * 8); RS(b->fp, (N + 1) * 8);
    RS(b->scm_cov, (n_unique + 1) * 4);
    std::vector<uint64_t> occ(N);
    for (uint64_t r = 0; r < n_reads; ++r)
        for (uint64_t i = scm_off[r], j = 0; i < scm_off[r + 1]; ++i, ++j)
            occ[i] = (b->sid_base + r) << 32 | j << 1 | (m_pos[i] & 1u);
    if (N) {
        CK(cudaMemcpyAsync(b->key.p, k_mer, N * 8, cudaMemcpyHostToDevice, st));     // sr_db_stat keys on k_mer >> 1 = the id (syncmer.c:896)
        CK(cudaMemcpyAsync(b->kid.p, k_mer, N * 8, cudaMemcpyHostToDevice, st));     // the arc tally drops the low bit as well
        CK(cudaMemcpyAsync(b->occ.p, occ.data(), N * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(b->m_pos.p, m_pos, N * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(b->s_mer.p, s_mer, N * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(b->fp.p, 0, N * 8, st));
    }
    if (n_unique) CK(cudaMemcpyAsync(b->scm_cov.p, cov, n_unique * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    b->h2d_bytes += N * 36 + n_unique * 4;
    b->n_reads = n_reads; b->n_syncmers = N; b->n_unique = n_unique;
    b->extracted = true; b->counted = true; b->sorted = false; b->adopted = false; b->sizes_known = false; b->have_kid_local = false;
    b->keys_are_ids = true;
    b->smer_slots = 0;
    return SG_OK;
}

int sg_count(sg_batch *b)
{
    if (!b) return SG_E_ARG;
    if (!b->extracted) return SG_E_STATE;
    if (b->keys_are_ids) return SG_E_STATE;                         // the ids exist already (sg_batch_set_lists_host)
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    const uint64_t N = b->t_n();
    if (N == 0) return SG_E_EMPTY;                                  // reference returns NULL, syncmer.c:1414-1417
    int rc = ensure_sorted(b);
    if (rc) return rc;
    ctx->t_begin(SG_T_GROUP);
    RS(b->flags, (N + 1) * 4); RS(b->differs, N + 1); RS(b->ids, (N + 2) * 8); RS(b->ids_tmp, scan_tmp_words(N) * 8);
    RS(b->status, 16 * 8); RS(b->kid, (N + 1) * 8);
    b->has_conflict = false;
    unsigned long long *status = (unsigned long long *) b->status.p;
    CK(cudaMemsetAsync(status, 0, 4 * 8, st));
    CK(cudaMemsetAsync(b->differs.p, 0, N + 1, st));
    {
        // element 0 is always a head; the kernel writes flags for lanes >= 1 only
        uint32_t one = 1;
        CK(cudaMemcpyAsync(b->flags.p, &one, 4, cudaMemcpyHostToDevice, st));
    }
    VerifyArgs V;
    memset(&V, 0, sizeof(V));
    V.skey = (const uint64_t *) b->skey.p; V.sval = (const uint64_t *) b->sval.p; V.socc = (const uint64_t *) b->socc.p;
    V.m_pos = (const uint32_t *) b->m_pos.p; V.hoff = (const uint64_t *) b->hoff.p; V.hoco_s = (const uint8_t *) b->hoco_s.p;
    V.hoco_l = (const uint32_t *) b->hoco_l.p; V.sid_base = b->sid_base; V.n = N; V.k = b->k;
    V.newid = (uint32_t *) b->flags.p; V.differs = (uint8_t *) b->differs.p; V.status = status;
    if (b->exact_verify && !b->adopted) {
        const uint64_t warps = (N + 30) / 31 + 1;
        verify_kernel<<<nblk(warps * 32, 256), 256, 0, st>>>(V);
        ctx->count_launch(SG_T_GROUP, 1);
    } else {
        // hash + independent 64-bit fingerprint (also what travels with tuples adopted from other GPUs,
        // whose reads are not here)
        verify_fp_kernel<<<nblk(N, 256), 256, 0, st>>>((const uint64_t *) b->skey.p, (const uint64_t *) b->sfp.p, N,
                (uint32_t *) b->flags.p, (uint8_t *) b->differs.p, status);
        ctx->count_launch(SG_T_GROUP, 1);
    }
    unsigned long long hs[4];
    CK(cudaMemcpyAsync(hs, status, sizeof(hs), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    b->n_collisions = 0;
    if (hs[1] && b->adopted) { ctx->err = "64-bit hash collision between k-mers of different GPUs: count this input on one GPU"; return SG_E_COLLISION; }
    if (hs[1]) {
        // some hash group holds more than one k-mer: rebuild those groups the way process_kmer_cluster does
        RS(b->cls, (N + 1) * 4);
        SplitArgs S;
        S.skey = (uint64_t *) b->skey.p; S.sval = (uint64_t *) b->sval.p; S.socc = (uint64_t *) b->socc.p; S.ssmer = (uint64_t *) b->ssmer.p;
        S.t_val = (uint64_t *) b->sval_alt.p; S.t_occ = (uint64_t *) b->skey_alt.p; S.t_smer = (uint64_t *) b->ids.p;
        S.cls = (uint32_t *) b->cls.p; S.m_pos = V.m_pos; S.hoff = V.hoff; S.hoco_s = V.hoco_s; S.hoco_l = V.hoco_l;
        S.sid_base = b->sid_base; S.n = N; S.k = b->k; S.newid = (uint32_t *) b->flags.p; S.differs = (const uint8_t *) b->differs.p;
        S.status = status;
        split_kernel<<<nblk(N, 128), 128, 0, st>>>(S);
        ctx->count_launch(SG_T_GROUP, 1);
        CK(cudaMemcpyAsync(hs, status, sizeof(hs), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        b->n_collisions = hs[2];
    }
    LAUNCHED(SG_T_GROUP, launch_scan_u32_u64((const uint32_t *) b->flags.p, (uint64_t *) b->ids.p, N, (uint64_t *) b->ids_tmp.p, st));
    uint64_t U = 0;
    CK(cudaMemcpyAsync(&U, (uint64_t *) b->ids.p + N, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    b->n_unique = U;
    RS(b->scm_h, (U + 1) * 8); RS(b->scm_s, (U + 1) * 8); RS(b->scm_cov, (U + 1) * 4); RS(b->scm_occ_off, (U + 2) * 8);
    FillArgs F;
    F.skey = (const uint64_t *) b->skey.p; F.sval = (const uint64_t *) b->sval.p; F.socc = (const uint64_t *) b->socc.p;
    F.ssmer = (const uint64_t *) b->ssmer.p; F.newid = (const uint32_t *) b->flags.p; F.ex = (const uint64_t *) b->ids.p;
    F.n = N; F.kid = (uint64_t *) b->kid.p; F.scm_h = (uint64_t *) b->scm_h.p; F.scm_s = (uint64_t *) b->scm_s.p;
    F.occ_off = (uint64_t *) b->scm_occ_off.p; F.status = status;
    fill_kernel<<<nblk(N, 256), 256, 0, st>>>(F);
    cov_kernel<<<nblk(U, 256), 256, 0, st>>>((const uint64_t *) b->scm_occ_off.p, (uint32_t *) b->scm_cov.p, U);
    ctx->count_launch(SG_T_GROUP, 2);
    ctx->t_end(SG_T_GROUP);
    CK(cudaMemcpyAsync(hs, status, sizeof(hs), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    b->counted = true;
    if (hs[0]) {
        conflict_info_kernel<<<1, 1, 0, st>>>((const uint64_t *) b->skey.p, (const uint64_t *) b->ssmer.p, (const uint64_t *) b->socc.p,
                (const uint32_t *) b->flags.p, N - hs[3], status + 8);
        unsigned long long ci[5];
        CK(cudaMemcpyAsync(ci, status + 8, sizeof(ci), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int j = 0; j < 5; ++j) b->conflict[j] = ci[j];
        b->has_conflict = true;
        ctx->err = "identical kmers have different smers";
        return SG_E_SMER_CONFLICT;
    }
    return SG_OK;
}

int sg_count_conflict(sg_batch *b, uint64_t out[5])
{
    if (!b || !out) return SG_E_ARG;
    if (!b->has_conflict) return SG_E_STATE;
    for (int j = 0; j < 5; ++j) out[j] = b->conflict[j];
    return SG_OK;
}

int sg_batch_buffer(sg_batch *b, int which, void **ptr, uint64_t *n)
{
    if (!b || !ptr || !n) return SG_E_ARG;
    switch (which) {
        case SG_BUF_KEY: *ptr = b->key.p; *n = b->n_syncmers; break;
        case SG_BUF_OCC: *ptr = b->occ.p; *n = b->n_syncmers; break;
        case SG_BUF_SMER: *ptr = b->s_mer.p; *n = b->n_syncmers; break;
        case SG_BUF_MPOS: *ptr = b->m_pos.p; *n = b->n_syncmers; break;
        case SG_BUF_KID: *ptr = b->kid.p; *n = b->counted ? b->t_n() : 0; break;
        case SG_BUF_SORTED_OCC: *ptr = b->socc.p; *n = b->sorted ? b->t_n() : 0; break;
        case SG_BUF_SCM_H: *ptr = b->scm_h.p; *n = b->counted ? b->n_unique : 0; break;
        case SG_BUF_SCM_COV: *ptr = b->scm_cov.p; *n = b->counted ? b->n_unique : 0; break;
        case SG_BUF_ADOPTED_OCC: { const int rc_ = sg::ensure_adopted_soa(b); if (rc_) return rc_; *ptr = b->aocc.p; *n = b->n_adopted; break; }
        default: return SG_E_ARG;
    }
    return SG_OK;
}

int sg_count_sizes(sg_batch *b, sg_count_sizes_t *out)
{
    if (!b || !out) return SG_E_ARG;
    if (!b->counted) return SG_E_STATE;
    out->n_syncmers = b->t_n();
    out->n_unique = b->n_unique;
    out->n_hash_collisions = b->n_collisions;
    return SG_OK;
}

int sg_count_download(sg_batch *b, const sg_count_out_t *o)
{
    if (!b || !o) return SG_E_ARG;
    if (!b->counted) return SG_E_STATE;
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    const uint64_t N = b->t_n(), U = b->n_unique;
    if (o->h) CK(cudaMemcpyAsync(o->h, b->scm_h.p, U * 8, cudaMemcpyDeviceToHost, st));
    if (o->s) CK(cudaMemcpyAsync(o->s, b->scm_s.p, U * 8, cudaMemcpyDeviceToHost, st));
    if (o->cov) CK(cudaMemcpyAsync(o->cov, b->scm_cov.p, U * 4, cudaMemcpyDeviceToHost, st));
    if (o->occ_off) CK(cudaMemcpyAsync(o->occ_off, b->scm_occ_off.p, (U + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (o->occ) CK(cudaMemcpyAsync(o->occ, b->socc.p, N * 8, cudaMemcpyDeviceToHost, st));
    if (o->k_mer_id) CK(cudaMemcpyAsync(o->k_mer_id, b->kid.p, N * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    b->d2h_bytes += U * 28 + N * 16;
    return SG_OK;
}

} // extern "C"
