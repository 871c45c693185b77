// sg_ec.cu -- f2 on the device: the per-read pass of read error correction.
//
// Reference: syncerr.c:342-612 (error blocks of a read and what replaces them), :144-288 (dfs_search over
// the all-syncmer graph), levdist.c:75-113, 156-225, 265-310 (the wavefront edit distance in extension mode
// that the search resumes arc by arc). The rules are those stated at the head of
// oatk_b200/host/syncerr_gpu.c, which is the host form of the same pass; this file restates them for a warp.
//
// One WARP per read (reads are handed out through an atomic counter). The read's syncmer list, its trusted
// flags, the block's bases, the candidate sequence, the search stack and the stacked wavefronts live in a
// per-warp arena in global memory (they stay in L1/L2); the live wavefront -- far[d] = last target index
// matched on diagonal d -- lives in shared memory, DIAGONALS ACROSS LANES:
//   slide   every lane extends its diagonal over at most 8 matching bases; a diagonal that is still
//           matching after that (the main diagonal of a block that is mostly right) is finished by the
//           whole warp, 32 bases per step with one ballot
//   widen   far'[d] = max(far[d-1], far[d] + 1, far[d+1] + 1), one diagonal per lane, then the reference's
//           band rule (levdist.c:99-113: beyond 2 bw + 1 live diagonals the upper bound WIDENS to the query
//           length, it does not clip to bw)
// The reference slides the diagonals in ascending order and stops at the first one that reaches the end of
// the target or of the query, leaving the later ones for the resumed call. A slide depends on nothing but
// its own diagonal and is idempotent, and a resumed call only ever sees a LONGER query with the same
// prefix, so sliding all of them at once and reporting the lowest one that reached an end leaves the same
// scores, the same ends and the same choices.
//
// The depth-first search is iterative (explicit frames); its bookkeeping -- best / second-best score, the
// ambiguity tests on bases and on paths, the cap of 10 000 leaves, the tail-block rule that drops a last
// vertex the read covers only partly -- follows the host form statement by statement, and so does the
// rewrite of the read's list, including syncerr.c:579's test of k_mer[end] where k_mer[beg] is meant.
//
// A vertex' text is the k hoco bases of the first occurrence of its syncmer that no earlier correction
// touched (what scg_consensus writes in hoco mode, syncasm.c:911-931); the host sends that occurrence per
// arc, the bases are read straight from the device-resident hoco_s.
//
// A read whose search outgrows the arena (path deeper, wavefronts more numerous, or a longer list than the
// arena was cut for) is put on an overflow list and run again, alone, in an arena cut for the worst case.
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <cmath>
#include <ctime>
#include <cstdio>
#include <cstdlib>
#include "sg_common.cuh"
#include "sg_internal.h"
#include "sg_host.h"

namespace sg {

constexpr int EC_WARPS = 4;
constexpr int EC_MAX_LEAVES = 10000;
constexpr int EC_MIN_BLOCK = 10;
constexpr int EC_MIN_BAND = 6;
constexpr int WAVE_NONE = INT32_MIN / 2;
constexpr uint64_t NO_SINK = ~0ull;
constexpr uint64_t ID_MASK = ~1ull;
constexpr uint64_t NO_TEXT = ~0ull;
enum { EC_FAILURE = 0, EC_SUCCESS = 1, EC_AMBISNQ = 2, EC_AMBISEQ = 3 };

struct EcCaps {
    uint32_t seq;        // bytes of each of the three sequence buffers (block, candidate, best candidate)
    uint32_t path;       // vertices on a path == search frames
    uint32_t stash;      // stacked wavefront entries
    uint32_t list;       // syncmers on a read (input) / of the rewritten list (output)
    uint32_t wave;       // live diagonals (shared memory, two buffers of this size per warp)
};

struct EcArgs {
    const uint64_t *hoff; const uint8_t *hoco_s; const uint32_t *hoco_l;
    const uint64_t *scm_off; const uint64_t *kid; const uint32_t *m_pos;
    uint64_t n_reads;
    int k;
    const uint8_t *del;
    const uint64_t *arc_v, *arc_w, *arc_txt; const uint32_t *arc_ls;
    uint64_t n_arc;
    double max_edist;
    // work: either all reads through a counter, or the reads of a list
    unsigned int *work;
    const uint32_t *todo; uint32_t n_todo;
    uint8_t *arena; uint64_t arena_per_warp;
    EcCaps cap;
    // results
    uint64_t *out_k; uint32_t *out_p; unsigned long long *out_count; uint64_t out_cap;
    uint64_t *out_off; uint32_t *out_n;            // per read; out_n = 0xffffffff: list unchanged
    unsigned long long *stats;                     // 11 counters, as the host's
    unsigned int *over_count; uint32_t *over_reads; uint32_t over_cap;
};

// the arena of one warp, cut from one allocation
struct EcArena {
    uint8_t *tseq, *cand, *best_seq;
    uint64_t *path, *best_path;
    int *frames;                                   // 10 ints per frame
    int *stash;
    uint64_t *lk; uint32_t *lp; uint8_t *lt;       // the read's list: k_mer, m_pos, trusted
    uint64_t *ok; uint32_t *op;                    // the rewritten list
};
__host__ __device__ inline uint64_t ec_arena_bytes(const EcCaps &c)
{
    auto up = [](uint64_t x) { return (x + 15) & ~15ull; };
    return 3 * up(c.seq) + 2 * up(8ull * c.path) + up(40ull * c.path) + up(4ull * c.stash) + up(8ull * c.list) + up(4ull * c.list) + up(c.list)
         + up(8ull * c.list) + up(4ull * c.list);
}
__device__ inline EcArena ec_cut(uint8_t *p, const EcCaps &c)
{
    auto up = [](uint64_t x) { return (x + 15) & ~15ull; };
    EcArena a;
    a.tseq = p; p += up(c.seq); a.cand = p; p += up(c.seq); a.best_seq = p; p += up(c.seq);
    a.path = (uint64_t *) p; p += up(8ull * c.path); a.best_path = (uint64_t *) p; p += up(8ull * c.path);
    a.frames = (int *) p; p += up(40ull * c.path);
    a.stash = (int *) p; p += up(4ull * c.stash);
    a.lk = (uint64_t *) p; p += up(8ull * c.list); a.lp = (uint32_t *) p; p += up(4ull * c.list); a.lt = p; p += up(c.list);
    a.ok = (uint64_t *) p; p += up(8ull * c.list); a.op = (uint32_t *) p;
    return a;
}

__device__ __forceinline__ int hoco_code(const uint8_t *hs, uint32_t p) { return (hs[p >> 2] >> ((3 - (p & 3)) << 1)) & 3; }

struct Wave {
    int tl, ql, bw, score, t_end, q_end, d_lo, n;
    int *far, *buf0, *buf1;                        // far points into buf0 or buf1
    int which;
};

// matching bases from a[0], b[0] on, at most `room`, by the whole warp
__device__ __forceinline__ int coop_prefix(const uint8_t *a, const uint8_t *b, int room, int lane)
{
    for (int base = 0; base < room; base += 32) {
        const int i = base + lane;
        const bool ne = i < room ? a[i] != b[i] : true;
        const uint32_t m = __ballot_sync(SG_FULL, ne);
        if (m) return min(room, base + __ffs(m) - 1);
    }
    return room;
}

// the wavefront carries on until a diagonal reaches the end of the target or of the query, or the band is left
__device__ void wave_run(Wave &w, const uint8_t *ts, const uint8_t *qs, int lane)
{
    for (;;) {
        int te = -1, qe = -1;
        bool done = false;
        for (int j0 = 0; j0 < w.n && !done; j0 += 32) {
            const int j = j0 + lane, d = w.d_lo + j;
            int k = 0, room = 0;
            bool live = false;
            if (j < w.n) {
                k = w.far[j];
                live = !(k >= w.tl || k + d >= w.ql);
                if (live) room = min(w.ql - d, w.tl) - 1 - k;
            }
            bool more = false;
            if (live && room > 0) {
                const uint8_t *a = ts + k + 1, *b = qs + k + d + 1;
                const int lim = min(room, 8);
                int c = 0;
                while (c < lim && a[c] == b[c]) ++c;
                k += c; room -= c;
                more = c == 8 && room > 0;
            }
            uint32_t mm = __ballot_sync(SG_FULL, more);
            while (mm) {
                const int src = __ffs(mm) - 1;
                mm &= mm - 1;
                const int sk = __shfl_sync(SG_FULL, k, src), sd = __shfl_sync(SG_FULL, d, src), sroom = __shfl_sync(SG_FULL, room, src);
                const int c = coop_prefix(ts + sk + 1, qs + sk + sd + 1, sroom, lane);
                if (lane == src) k += c;
            }
            const bool reached = live && (k == w.tl - 1 || k + d == w.ql - 1);
            if (live) w.far[j] = k;
            const uint32_t rm = __ballot_sync(SG_FULL, reached);
            if (rm) {
                const int src = __ffs(rm) - 1;
                te = __shfl_sync(SG_FULL, k, src);
                qe = te + w.d_lo + j0 + src;
                done = true;
            }
        }
        __syncwarp();
        if (done) { w.t_end = te + 1; w.q_end = qe + 1; return; }
        // one more edit: the wave widens by a diagonal on either side, then the band rule trims it
        const int n = w.n;
        int *g = w.which ? w.buf0 : w.buf1;
        const int *f = w.far;
        for (int j = lane; j < n + 2; j += 32) {
            const int ins = j >= 2 ? f[j - 2] : WAVE_NONE;
            const int sub = (j >= 1 && j <= n) ? f[j - 1] + 1 : WAVE_NONE;
            const int del = j < n ? f[j] + 1 : WAVE_NONE;
            g[j] = max(max(ins, sub), del);
        }
        const int d0 = w.d_lo - 1;
        int lo = 0, hi = n + 2;
        if (w.bw < 0 || n < 2 * w.bw + 1) {
            if (d0 < -w.tl) ++lo;
            if (d0 + n + 1 > w.ql) --hi;
        } else {
            const int min_d = max(-w.bw, -w.tl), max_d = max(w.bw, w.ql);
            while (d0 + lo < min_d) ++lo;
            while (d0 + hi - 1 > max_d) --hi;
        }
        w.d_lo = d0 + lo;
        w.n = hi - lo;
        w.far = g + lo;
        w.which ^= 1;
        __syncwarp();
        ++w.score;
        if (w.bw >= 0 && w.score > w.bw) { w.t_end = 0; w.q_end = 0; return; }
    }
}

// first arc out of oriented vertex v (arcs are sorted by v)
__device__ __forceinline__ uint64_t arc_lower(const EcArgs &A, uint64_t v)
{
    uint64_t lo = 0, hi = A.n_arc;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (A.arc_v[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

struct Search {
    int status, leaves, best, second;
    int cand_l, best_seq_l, path_n, best_path_n, stash_n;
    bool overflow;
};

// the search below `src` for one block; S.path[0] = src and the wave are set up by the caller
__device__ void search_block(const EcArgs &A, const EcArena &M, Search &S, Wave &w, uint64_t sink, int lane)
{
    int depth = 0;
    // frame layout: arc_i_lo, arc_i_hi (64-bit index split), arc_end_lo, arc_end_hi, l0, n0, d0, t_end0, q_end0|s0 packed? -- kept plain:
    // [0] arc_i  [1] arc_end  [2] l0  [3] n0  [4] d0  [5] t_end0  [6] q_end0  [7] s0  [8] lo0  [9] stash0      (arc indices fit 32 bits: n_arc < 2^31 is checked on the host)
    auto enter = [&]() -> bool {                       // search_from's prologue; false: nothing to do at this node
        if (S.leaves >= EC_MAX_LEAVES) return false;
        if ((uint32_t) depth >= A.cap.path || (uint64_t) S.stash_n + (uint64_t) w.n > A.cap.stash) { S.overflow = true; return false; }
        const uint64_t from = M.path[S.path_n - 1];
        const uint64_t a0 = arc_lower(A, from);
        uint64_t a1 = a0;
        while (a1 < A.n_arc && A.arc_v[a1] == from) ++a1;
        int *F = M.frames + 10 * depth;
        if (lane == 0) {
            F[0] = (int) a0; F[1] = (int) a1; F[2] = S.cand_l; F[3] = S.path_n; F[4] = w.n; F[5] = w.t_end; F[6] = w.q_end; F[7] = w.score; F[8] = w.d_lo; F[9] = S.stash_n;
        }
        for (int j = lane; j < w.n; j += 32) M.stash[S.stash_n + j] = w.far[j];
        S.stash_n += w.n;
        __syncwarp();
        ++depth;
        return true;
    };
    if (!enter()) return;
    while (depth > 0) {
        int *F = M.frames + 10 * (depth - 1);
        const int ai = F[0], aend = F[1], l0 = F[2], n0 = F[3], d0 = F[4], t_end0 = F[5], q_end0 = F[6], s0 = F[7], lo0 = F[8], stash0 = F[9];
        if (ai >= aend) {                               // search_from's epilogue
            S.stash_n = stash0;
            --depth;
            if (depth > 0) {
                // back in the parent, after the recursive call: restore the state in front of the parent's arc
                int *P = M.frames + 10 * (depth - 1);
                S.path_n = P[3]; S.cand_l = P[2];
                w.t_end = P[5]; w.q_end = P[6]; w.score = P[7]; w.n = P[4]; w.d_lo = P[8];
                w.far = w.which ? w.buf1 : w.buf0;      // any buffer will do: the whole wave is rewritten
                for (int j = lane; j < w.n; j += 32) w.far[j] = M.stash[P[9] + j];
                __syncwarp();
            }
            continue;
        }
        __syncwarp();
        if (lane == 0) F[0] = ai + 1;
        __syncwarp();
        const uint64_t v = A.arc_w[ai];
        const int ov = (int) A.arc_ls[ai], vl = A.k;
        const uint64_t txt = A.arc_txt[ai];
        const int add = vl - ov;
        if ((uint32_t) S.path_n >= A.cap.path || (uint64_t) S.cand_l + (uint64_t) add + 1 > A.cap.seq) { S.overflow = true; return; }
        if (lane == 0) M.path[S.path_n] = v;
        ++S.path_n;
        // the part of v's text that does not overlap the vertex before it
        if (txt == NO_TEXT) {
            for (int i = lane; i < add; i += 32) M.cand[S.cand_l + i] = 4;
        } else {
            const uint64_t rd = txt >> 32;
            const uint32_t p = (uint32_t) (txt & 0xffffffffu) >> 1;
            const int rr = (int) ((txt & 1) ^ (v & 1));
            const uint8_t *hs = A.hoco_s + A.hoff[rd] / 4;
            for (int i = lane; i < add; i += 32) {
                const int x = ov + i;
                M.cand[S.cand_l + i] = (uint8_t) (rr ? 3 - hoco_code(hs, p + (uint32_t) (vl - 1 - x)) : hoco_code(hs, p + (uint32_t) x));
            }
        }
        S.cand_l += add;
        __syncwarp();
        w.ql = S.cand_l;
        wave_run(w, M.tseq, M.cand, lane);
        const int score = w.score + w.tl - w.t_end;     // unaligned target bases count as edits
        if (score <= w.bw && (sink == NO_SINK || sink == v)) {
            S.status = EC_SUCCESS;
            if (score <= S.best) {
                if (w.t_end > t_end0) S.second = S.best;
                S.best = score;
                if (sink == NO_SINK && w.q_end < w.ql) --S.path_n;
                if (S.best == S.second) {
                    bool differ = w.q_end != S.best_seq_l;
                    if (!differ) {
                        bool ne = false;
                        for (int i = lane; i < w.q_end; i += 32) ne |= M.cand[i] != M.best_seq[i];
                        differ = __any_sync(SG_FULL, ne);
                    }
                    if (differ) S.status = EC_AMBISEQ;
                    if (S.status == EC_SUCCESS) {
                        bool same = S.path_n == S.best_path_n;
                        if (same) {
                            bool ne = false;
                            for (int i = lane; i < S.path_n; i += 32) ne |= M.path[i] != M.best_path[i];
                            same = !__any_sync(SG_FULL, ne);
                        }
                        if (!same) S.status = EC_AMBISNQ;
                    }
                }
                __syncwarp();
                for (int i = lane; i < w.q_end; i += 32) M.best_seq[i] = M.cand[i];
                S.best_seq_l = w.q_end;
                for (int i = lane; i < S.path_n; i += 32) M.best_path[i] = M.path[i];
                S.best_path_n = S.path_n;
                __syncwarp();
            } else if (score < S.second) S.second = score;
        }
        bool descended = false;
        if (w.score <= w.bw && w.ql - vl <= w.tl + w.bw && ((sink != NO_SINK && sink != v) || w.t_end < w.tl)) {
            descended = enter();
            if (S.overflow) return;
        } else ++S.leaves;
        if (!descended) {
            // back to the state in front of this arc
            S.path_n = n0; S.cand_l = l0;
            w.t_end = t_end0; w.q_end = q_end0; w.score = s0; w.n = d0; w.d_lo = lo0;
            w.far = w.which ? w.buf1 : w.buf0;
            for (int j = lane; j < d0; j += 32) w.far[j] = M.stash[stash0 + j];
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(32 * EC_WARPS) ec_read_kernel(EcArgs A)
{
    extern __shared__ __align__(16) int ec_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t gw = (uint64_t) blockIdx.x * EC_WARPS + wid;
    const EcArena M = ec_cut(A.arena + gw * A.arena_per_warp, A.cap);
    int *wb0 = ec_smem + (size_t) wid * 2 * A.cap.wave, *wb1 = wb0 + A.cap.wave;
    int st[11];
    for (int i = 0; i < 11; ++i) st[i] = 0;

    for (;;) {
        unsigned int wi = 0;
        if (lane == 0) wi = atomicAdd(A.work, 1u);
        wi = __shfl_sync(SG_FULL, wi, 0);
        uint32_t rid;
        if (A.todo) { if (wi >= A.n_todo) break; rid = A.todo[wi]; }
        else { if ((uint64_t) wi >= A.n_reads) break; rid = wi; }
        const uint64_t s0 = A.scm_off[rid];
        const int n = (int) (A.scm_off[rid + 1] - s0);
        const int H = (int) A.hoco_l[rid];
        const uint8_t *hs = A.hoco_s + A.hoff[rid] / 4;
        bool overflow = (uint32_t) n > A.cap.list;
        int st_read[11];
        for (int i = 0; i < 11; ++i) st_read[i] = 0;
        int nc = 0;                                     // rewritten list so far
        int updated = 1;
        if (!overflow) {
            for (int j = lane; j < n; j += 32) {
                const uint64_t kk = A.kid[s0 + j];
                M.lk[j] = kk; M.lp[j] = A.m_pos[s0 + j];
                M.lt[j] = (uint8_t) (A.del[kk >> 1] ? 1 : 0);      // 1: the syncmer is flagged deleted
            }
            __syncwarp();
            const uint64_t *K = M.lk;
            const uint32_t *P = M.lp;
            auto push = [&](uint64_t kk, uint32_t pp) {
                if ((uint32_t) nc >= A.cap.list) { overflow = true; return; }
                if (lane == 0) { M.ok[nc] = kk; M.op[nc] = pp; }
                ++nc;
            };
            int beg = -1, end;
            for (;;) {
                uint32_t from = beg < 1 ? 0 : (P[beg - 1] >> 1) + (uint32_t) A.k;
                from += EC_MIN_BLOCK;
                for (end = beg + 1; end < n; ++end)
                    if (!M.lt[end] && !(K[end] & 1) && (P[end] >> 1) >= from) break;        // the next anchor
                if (beg >= 0 || end < n) {
                    uint64_t src, sink;
                    int l, rev, res;
                    uint32_t at;
                    if (beg < 0) {
                        beg = end;
                        src = (K[beg] & ID_MASK) | (uint64_t) !(P[beg] & 1);
                        at = 0; sink = NO_SINK; l = (int) (P[beg] >> 1); rev = 1;
                    } else {
                        --beg;
                        src = (K[beg] & ID_MASK) | (P[beg] & 1);
                        at = (P[beg] >> 1) + (uint32_t) A.k;
                        if (end >= n) { sink = NO_SINK; l = H - (int) at; }
                        else { sink = (K[end] & ID_MASK) | (P[end] & 1); l = (int) (P[end] >> 1) - (int) at; }
                        rev = 0;
                    }
                    Search S;
                    S.overflow = false;
                    if (l < 0 || (uint32_t) l + 1 > A.cap.seq) { overflow = true; break; }
                    if (l >= EC_MIN_BLOCK) {
                        for (int i = lane; i < l; i += 32)
                            M.tseq[i] = (uint8_t) (rev ? 3 - hoco_code(hs, at + (uint32_t) (l - 1 - i)) : hoco_code(hs, at + (uint32_t) i));
                        Wave w;
                        w.tl = l; w.ql = 0; w.score = 0; w.t_end = 0; w.q_end = 0; w.d_lo = 0; w.n = 1;
                        w.buf0 = wb0; w.buf1 = wb1; w.far = wb0; w.which = 0;
                        if (lane == 0) wb0[0] = -1;
                        w.bw = (int) ceil(l * A.max_edist);
                        if (w.bw < EC_MIN_BAND) w.bw = EC_MIN_BAND;
                        if ((uint32_t) (2 * w.bw + 8) > A.cap.wave) { overflow = true; break; }
                        S.status = EC_FAILURE; S.leaves = 0; S.best = S.second = INT32_MAX;
                        S.cand_l = 0; S.best_seq_l = 0; S.path_n = 1; S.best_path_n = 0; S.stash_n = 0;
                        if (lane == 0) M.path[0] = src;
                        __syncwarp();
                        search_block(A, M, S, w, sink, lane);
                        if (S.overflow) { overflow = true; break; }
                        res = S.status;
                        if (sink == NO_SINK) { ++st_read[0]; ++st_read[1 + res]; }
                        else { ++st_read[5]; ++st_read[6 + res]; }
                    } else { res = EC_FAILURE; ++st_read[10]; S.best_path_n = 0; }

                    if (res == EC_SUCCESS) {
                        const uint64_t *bp = M.best_path;
                        const int np = S.best_path_n;
                        int j;
                        if (rev) {
                            for (j = np - 1; j > 0; --j) push((bp[j] & ID_MASK) | 1, 0xFFFFFFFFu ^ (uint32_t) (bp[j] & 1));
                        } else {
                            for (j = 1; j < np - 1; ++j) push((bp[j] & ID_MASK) | 1, 0xFFFFFFFEu | (uint32_t) (bp[j] & 1));
                            if (sink == NO_SINK && np > 1) push((bp[j] & ID_MASK) | 1, 0xFFFFFFFEu | (uint32_t) (bp[j] & 1));
                        }
                    } else if (rev) {
                        for (int j = 0; j < beg; ++j) push(K[j], P[j]);
                    } else if (beg + 1 < n) {
                        for (int j = beg + 1; j < end; ++j) push(K[j], P[j]);
                    }
                    if (overflow) break;
                } else updated = 0;

                for (beg = end + 1; beg < n; ++beg)
                    if (M.lt[beg] || (K[end] & 1)) break;                                 // syncerr.c:579
                if (beg > n) break;
                for (int j = end; j < beg; ++j) push(K[j], P[j]);
                if (overflow) break;
            }
        }
        __syncwarp();
        if (overflow) {
            if (lane == 0) {
                const unsigned int o = atomicAdd(A.over_count, 1u);
                if (o < A.over_cap) A.over_reads[o] = rid;
                A.out_n[rid] = 0xffffffffu;
            }
            continue;
        }
        for (int i = 0; i < 11; ++i) st[i] += st_read[i];
        if (!updated) { if (lane == 0) A.out_n[rid] = 0xffffffffu; continue; }
        unsigned long long o = 0;
        if (lane == 0) o = atomicAdd(A.out_count, (unsigned long long) nc);
        o = __shfl_sync(SG_FULL, o, 0);
        if (o + (unsigned long long) nc <= A.out_cap)
            for (int j = lane; j < nc; j += 32) { A.out_k[o + j] = M.ok[j]; A.out_p[o + j] = M.op[j]; }
        if (lane == 0) { A.out_off[rid] = o; A.out_n[rid] = (uint32_t) nc; }
        __syncwarp();
    }
    if (lane == 0) for (int i = 0; i < 11; ++i) if (st[i]) atomicAdd(A.stats + i, (unsigned long long) st[i]);
}

// ---- the error filter (syncerr.c:679-757) over the device-resident arc list -------------------------------------------
// A syncmer is suspect when its coverage is below err_mer_c, or -- below max_err_c -- when one of its two sides has arcs
// but none with coverage >= err_arc_c and >= max_arc_f * min(cov, cov') (the product in double precision, as the
// reference writes it). Arcs are the (v, w, cov, comp) records sg_arcs leaves sorted by (v, w, comp): the arcs of an
// oriented vertex are found by bisection. One thread per syncmer; most leave at the first test.
__device__ __forceinline__ uint64_t arc4_lower(const uint64_t *arcs4, uint64_t n, uint64_t v)
{
    uint64_t lo = 0, hi = n;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (arcs4[4 * mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}
__global__ void __launch_bounds__(256) ec_filter_kernel(const uint64_t *arcs4, uint64_t n_arc, const uint32_t *cov, const uint8_t *del_prev, uint64_t n_scm,
        uint32_t err_mer_c, uint32_t max_err_c, uint32_t err_arc_c, double max_arc_f, uint8_t *err)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_scm) return;
    uint8_t e = 0;
    const uint32_t cv = cov[i];
    if (!(del_prev && del_prev[i]) && cv < max_err_c) {
        if (cv < err_mer_c) e = 1;
        else {
            int side_ok[2] = {-1, -1};
            for (int side = 0; side < 2; ++side) {
                const uint64_t v = i << 1 | (uint64_t) side;
                uint64_t a = arc4_lower(arcs4, n_arc, v);
                if (a >= n_arc || arcs4[4 * a] != v) continue;                 // no arc on that side
                side_ok[side] = 0;
                for (; a < n_arc && arcs4[4 * a] == v; ++a) {
                    const uint32_t ac = (uint32_t) arcs4[4 * a + 2], cw = cov[arcs4[4 * a + 1] >> 1];
                    if (ac >= err_arc_c && (double) ac >= (double) (cv < cw ? cv : cw) * max_arc_f) { side_ok[side] = 1; break; }
                }
            }
            if (!side_ok[0] || !side_ok[1]) e = 1;
        }
    }
    err[i] = e;
}
// an arc survives when neither end is suspect or was deleted before
__global__ void __launch_bounds__(256) ec_live_flag_kernel(const uint64_t *arcs4, uint64_t n_arc, const uint8_t *err, const uint8_t *del_prev, uint32_t *flag)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_arc) return;
    const uint64_t v = arcs4[4 * i] >> 1, w = arcs4[4 * i + 1] >> 1;
    flag[i] = (err[v] || err[w] || (del_prev && (del_prev[v] || del_prev[w]))) ? 0u : 1u;
}
__global__ void __launch_bounds__(256) ec_live_pack_kernel(const uint64_t *arcs4, uint64_t n_arc, const uint32_t *flag, const uint64_t *ex, uint64_t *out4)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_arc || !flag[i]) return;
    const uint64_t o = ex[i];
    out4[4 * o] = arcs4[4 * i]; out4[4 * o + 1] = arcs4[4 * i + 1]; out4[4 * o + 2] = arcs4[4 * i + 2]; out4[4 * o + 3] = arcs4[4 * i + 3];
}

// ---- f1: the votes of calc_syncmer_overlap (syncasm.c:477-582) for a list of arcs between single syncmers ----------------
// The distance between two neighbouring syncmers is the most frequent difference of their hoco start positions over the
// reads that carry both next to each other on the asked strands. One WARP per arc: a lane takes one occurrence of the first
// syncmer, looks at the entry in front of it and the one behind it on its read, and votes when that entry is the second
// syncmer (entries an earlier correction touched do not vote). Equal votes are merged with ballots into a table of four
// values; more than four, or two values with the same highest count, are left to the host, which breaks the tie in the
// slot order of the reference's hash table. No votes: 0, like the empty table there.
struct VoteArgs {
    const uint64_t *arcs4; uint64_t n;
    const uint64_t *socc, *occ_off;           // occurrences of every syncmer, (sid, idx) order
    const uint64_t *scm_off, *kid; const uint32_t *m_pos;
    uint64_t sid_base, n_reads;
    int32_t *dist; uint8_t *flag;             // flag 1: the host decides
};
__global__ void __launch_bounds__(128) arc_vote_kernel(VoteArgs A)
{
    const int lane = threadIdx.x & 31;
    const uint64_t w0 = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    for (uint64_t a = w0; a < A.n; a += nw) {
        if (A.arcs4[4 * a + 3]) { if (lane == 0) { A.dist[a] = 0; A.flag[a] = 2; } continue; }     // a complement: takes its mirror's value
        const uint64_t x = A.arcs4[4 * a], y = A.arcs4[4 * a + 1];
        const uint64_t id1 = x >> 1, id2 = y >> 1, rc1 = x & 1, rc2 = y & 1;
        const uint64_t o0 = A.occ_off[id1], o1 = A.occ_off[id1 + 1];
        int tv[4] = {0, 0, 0, 0}, tc[4] = {0, 0, 0, 0}, nt = 0;
        bool over = false;
        for (uint64_t ob = o0; ob < o1; ob += 32) {
            bool valid = false;
            int vote = 0;
            if (ob + lane < o1) {
                const uint64_t oc = A.socc[ob + lane];
                const uint64_t rd = (oc >> 32) - A.sid_base, i1 = (oc >> 1) & 0x7FFFFFFFull, s1 = oc & 1;
                const uint64_t e0 = A.scm_off[rd], e1 = A.scm_off[rd + 1];
                const uint64_t e = e0 + i1;
                if (rd < A.n_reads && e < e1 && !(A.kid[e] & 1)) {
                    const int p1 = (int) (A.m_pos[e] >> 1);
                    if (s1 != rc1) {
                        if (i1 >= 1) { const uint64_t kk = A.kid[e - 1]; const uint32_t mp = A.m_pos[e - 1];
                            if ((kk >> 1) == id2 && !(kk & 1) && (uint64_t) (mp & 1) != rc2) { valid = true; vote = p1 - (int) (mp >> 1); } }
                    } else {
                        if (e + 1 < e1) { const uint64_t kk = A.kid[e + 1]; const uint32_t mp = A.m_pos[e + 1];
                            if ((kk >> 1) == id2 && !(kk & 1) && (uint64_t) (mp & 1) == rc2) { valid = true; vote = (int) (mp >> 1) - p1; } }
                    }
                }
            }
            uint32_t left = __ballot_sync(SG_FULL, valid);
            while (left) {
                const int src = __ffs(left) - 1;
                const int v = __shfl_sync(SG_FULL, vote, src);
                const uint32_t same = __ballot_sync(SG_FULL, valid && vote == v);
                left &= ~same;
                int t = 0;
                for (; t < nt; ++t) if (tv[t] == v) break;
                if (t == nt) { if (nt < 4) { tv[nt] = v; tc[nt] = 0; ++nt; } else { over = true; t = 3; } }
                tc[t] += __popc(same);
            }
        }
        int best = 0, best_n = 0;
        bool tied = false;
        for (int t = 0; t < nt; ++t) {
            if (tc[t] > best_n) { best_n = tc[t]; best = tv[t]; tied = false; }
            else if (tc[t] == best_n) tied = true;
        }
        if (lane == 0) { A.dist[a] = best; A.flag[a] = (over || tied) ? 1 : 0; }
    }
}

} // namespace sg

using namespace sg;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return SG_E_CUDA; } } while (0)

// first-pass arena (tests shrink it to drive reads through the worst-case pass)
static uint32_t g_ec_path_cap = 256, g_ec_stash_cap = 16384;
extern "C" int sg_debug_set_ec_arena(uint32_t path, uint32_t stash)
{
    if (path < 1 || stash < 1) return SG_E_ARG;
    g_ec_path_cap = path; g_ec_stash_cap = stash;
    return SG_OK;
}

extern "C" int sg_ec_correct(sg_batch *b, const sg_ec_graph_t *g, double max_edist, sg_ec_result_t *res)
{
    if (!b || !g || !res || (g->n_arcs && (!g->arc_v || !g->arc_w || !g->arc_ls || !g->arc_txt)) || (g->n_syncmers && !g->del)) return SG_E_ARG;
    if (!b->extracted || !b->counted) return SG_E_STATE;
    if (g->n_arcs >= (1ull << 31)) return SG_E_LIMIT;
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    memset(res, 0, sizeof(*res));
    const uint64_t n_reads = b->n_reads, N = b->n_syncmers;
    if (n_reads == 0) return SG_OK;
    if (n_reads > 0xfffffff0ull) return SG_E_LIMIT;

    // per-read figures the arena is cut for (the host mirrors of the last extract / list update may be stale: ask the device)
    std::vector<uint32_t> hl(n_reads);
    std::vector<uint64_t> so(n_reads + 1);
    CK(cudaMemcpyAsync(hl.data(), b->hoco_l.p, n_reads * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(so.data(), b->scm_off.p, (n_reads + 1) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (so[n_reads] != N) { ctx->err = "sg_ec_correct: the device lists are not those of the last extract / list update"; return SG_E_STATE; }
    uint32_t max_l = 0, max_n = 0;
    for (uint64_t r = 0; r < n_reads; ++r) { max_l = std::max(max_l, hl[r]); max_n = std::max<uint32_t>(max_n, (uint32_t) (so[r + 1] - so[r])); }
    const int k = b->k;
    const int bw_max = std::max<int>(EC_MIN_BAND, (int) std::ceil(max_l * max_edist));

    sg::DevBuf &d_del = b->ec_del, &d_av = b->ec_av, &d_aw = b->ec_aw, &d_at = b->ec_at, &d_al = b->ec_al, &d_arena = b->ec_arena, &d_outk = b->ec_outk,
        &d_outp = b->ec_outp, &d_off = b->ec_off, &d_n = b->ec_n, &d_misc = b->ec_misc, &d_over = b->ec_over;
#define RSV(buf, bytes) do { if ((buf).reserve(bytes)) { ctx->err = "device allocation failed in sg_ec_correct"; return SG_E_NOMEM; } } while (0)
    RSV(d_del, g->n_syncmers + 16); RSV(d_av, g->n_arcs * 8 + 16); RSV(d_aw, g->n_arcs * 8 + 16); RSV(d_at, g->n_arcs * 8 + 16); RSV(d_al, g->n_arcs * 4 + 16);
    CK(cudaMemcpyAsync(d_del.p, g->del, g->n_syncmers, cudaMemcpyHostToDevice, st));
    if (g->n_arcs) {
        CK(cudaMemcpyAsync(d_av.p, g->arc_v, g->n_arcs * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_aw.p, g->arc_w, g->n_arcs * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_at.p, g->arc_txt, g->n_arcs * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_al.p, g->arc_ls, g->n_arcs * 4, cudaMemcpyHostToDevice, st));
    }
    b->h2d_bytes += g->n_syncmers + g->n_arcs * 28;
    // results: the rewritten lists can be longer than the originals (a corrected block holds what the read should have
    // had); twice the input plus slack, checked after the run
    const uint64_t out_cap = 2 * N + 64 * n_reads + 1024;
    RSV(d_outk, out_cap * 8); RSV(d_outp, out_cap * 4); RSV(d_off, n_reads * 8 + 16); RSV(d_n, n_reads * 4 + 16);
    RSV(d_misc, 64 * 8); RSV(d_over, n_reads * 4 + 16);
    CK(cudaMemsetAsync(d_misc.p, 0, 64 * 8, st));
    unsigned long long *d_stats = (unsigned long long *) d_misc.p;            // [0..10] stats, [11] out_count, [12] work (u32), [13] over_count (u32)

    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);

    EcArgs A;
    memset(&A, 0, sizeof(A));
    A.hoff = (const uint64_t *) b->hoff.p; A.hoco_s = (const uint8_t *) b->hoco_s.p; A.hoco_l = (const uint32_t *) b->hoco_l.p;
    A.scm_off = (const uint64_t *) b->scm_off.p; A.kid = (const uint64_t *) b->kid.p; A.m_pos = (const uint32_t *) b->m_pos.p;
    A.n_reads = n_reads; A.k = k; A.del = (const uint8_t *) d_del.p;
    A.arc_v = (const uint64_t *) d_av.p; A.arc_w = (const uint64_t *) d_aw.p; A.arc_txt = (const uint64_t *) d_at.p; A.arc_ls = (const uint32_t *) d_al.p;
    A.n_arc = g->n_arcs; A.max_edist = max_edist;
    A.out_k = (uint64_t *) d_outk.p; A.out_p = (uint32_t *) d_outp.p; A.out_count = d_stats + 11; A.out_cap = out_cap;
    A.out_off = (uint64_t *) d_off.p; A.out_n = (uint32_t *) d_n.p; A.stats = d_stats;
    A.work = (unsigned int *) (d_stats + 12); A.over_count = (unsigned int *) (d_stats + 13);
    A.over_reads = (uint32_t *) d_over.p; A.over_cap = (uint32_t) n_reads;

    // pass 1: every read, in an arena cut for ordinary searches; pass 2: the reads that outgrew it, in one cut for the worst case
    uint32_t n_over = 0;
    std::vector<uint32_t> over;
    for (int pass = 0; pass < 2; ++pass) {
        EcCaps c;
        c.seq = (uint32_t) (((uint64_t) max_l + (uint64_t) bw_max + 2ull * (uint64_t) k + 64 + 15) & ~15ull);   // a candidate grows to tl + bw + k before the search stops descending, plus one more vertex
        c.list = pass == 0 ? std::min<uint32_t>(2 * max_n + 64, 4096) : 0;
        c.wave = (uint32_t) (2 * bw_max + 8);
        if (pass == 0) { c.path = g_ec_path_cap; c.stash = g_ec_stash_cap; }
        else {
            // worst case: every arc adds one base, every wave is full
            c.path = (uint32_t) std::min<uint64_t>((uint64_t) max_l + (uint64_t) bw_max + 8, 1u << 20);
            c.stash = (uint32_t) std::min<uint64_t>((uint64_t) c.path * (uint64_t) c.wave, 1ull << 28);
            c.list = (uint32_t) std::min<uint64_t>((uint64_t) c.path + 2ull * max_n + 64, 1u << 24);
        }
        int warps = EC_WARPS;
        const size_t smem = (size_t) warps * 2 * c.wave * sizeof(int);
        if (smem > 200 * 1024) { ctx->err = "sg_ec_correct: band too wide for the shared-memory wavefront"; return SG_E_LIMIT; }
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(ec_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        const uint64_t per_warp = (ec_arena_bytes(c) + 255) & ~255ull;
        uint64_t n_work = pass == 0 ? n_reads : n_over;
        uint64_t grid = std::min<uint64_t>((n_work + warps - 1) / warps, (uint64_t) n_sm * (pass == 0 ? 8 : 1));
        // the arena must fit: fewer warps rather than a failed allocation
        const uint64_t budget = pass == 0 ? (8ull << 30) : (16ull << 30);
        while (grid > 1 && grid * warps * per_warp > budget) grid = (grid + 1) / 2;
        if (grid * warps * per_warp > (64ull << 30)) { ctx->err = "sg_ec_correct: a read needs more search memory than the device arena allows"; return SG_E_NOMEM; }
        RSV(d_arena, grid * warps * per_warp);
        A.arena = (uint8_t *) d_arena.p; A.arena_per_warp = per_warp; A.cap = c;
        CK(cudaMemsetAsync(A.work, 0, 4, st));
        CK(cudaMemsetAsync(A.over_count, 0, 4, st));
        if (pass == 1) {
            CK(cudaMemcpyAsync(d_over.p, over.data(), (size_t) n_over * 4, cudaMemcpyHostToDevice, st));
            A.todo = (const uint32_t *) d_over.p; A.n_todo = n_over;
            // the overflow list of this pass goes behind the work list
            A.over_reads = (uint32_t *) d_over.p + n_over; A.over_cap = (uint32_t) (n_reads - n_over);
        }
        ctx->t_begin(SG_T_EC);
        ec_read_kernel<<<(unsigned) grid, 32 * warps, smem, st>>>(A);
        ctx->count_launch(SG_T_EC, 1);
        ctx->t_end(SG_T_EC);
        CK(cudaGetLastError());
        unsigned int oc = 0;
        CK(cudaMemcpyAsync(&oc, A.over_count, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (pass == 0) {
            n_over = oc;
            res->n_overflow_reads = oc;
            if (!oc) break;
            over.resize(oc);
            CK(cudaMemcpyAsync(over.data(), d_over.p, (size_t) oc * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            std::sort(over.begin(), over.end());
        } else if (oc) { ctx->err = "sg_ec_correct: a read outgrew the worst-case search arena"; return SG_E_LIMIT; }
    }

    unsigned long long hs[12];
    CK(cudaMemcpyAsync(hs, d_stats, sizeof(hs), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (hs[11] > out_cap) { ctx->err = "sg_ec_correct: rewritten lists exceed the output buffer"; return SG_E_LIMIT; }
    for (int i = 0; i < 11; ++i) res->stats[i] = (int64_t) hs[i];
    const uint64_t n_out = hs[11];
    res->n_out = n_out;
    res->out_off = (uint64_t *) malloc((n_reads + 1) * 8);
    res->out_n = (uint32_t *) malloc((n_reads + 1) * 4);
    res->out_k = (uint64_t *) malloc((n_out + 1) * 8);
    res->out_p = (uint32_t *) malloc((n_out + 1) * 4);
    if (!res->out_off || !res->out_n || !res->out_k || !res->out_p) { sg_ec_result_free(res); return SG_E_NOMEM; }
    CK(cudaMemcpyAsync(res->out_off, d_off.p, n_reads * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(res->out_n, d_n.p, n_reads * 4, cudaMemcpyDeviceToHost, st));
    if (n_out) {
        CK(cudaMemcpyAsync(res->out_k, d_outk.p, n_out * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(res->out_p, d_outp.p, n_out * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    b->d2h_bytes += n_reads * 12 + n_out * 12;
    return SG_OK;
}


// The error filter on the device: sg_arcs(b, 0, 0) (the all-syncmer graph's arcs) must be what the batch holds, which
// this call makes sure of itself. Returns the suspect flags and what is left of the arc list, in the list's order.
extern "C" int sg_ec_filter(sg_batch *b, const uint8_t *del_prev, uint32_t err_mer_c, uint32_t max_err_c, uint32_t err_arc_c, double max_arc_f,
        sg_ec_filter_out_t *out)
{
    if (!b || !out) return SG_E_ARG;
    if (!b->counted || b->adopted) return SG_E_STATE;
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof(*out));
    uint64_t na = 0;
    const bool dbg = getenv("SG_EC_TIMING") != nullptr;
    auto now = []() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; };
    double t0 = now();
    int rc = sg_arcs(b, 0, 0.0, &na);
    if (rc) return rc;
    if (dbg) { fprintf(stderr, "[sg_ec_filter] sg_arcs %.3f s (%llu arcs)\n", now() - t0, (unsigned long long) na); t0 = now(); }
    const uint64_t U = b->n_unique;
    out->n_syncmers = U;
    sg::DevBuf &d_err = b->ec_err, &d_prev = b->ec_prev, &d_flag = b->ec_flag, &d_ex = b->ec_ex, &d_tmp = b->ec_tmp, &d_live = b->ec_live;
    if (d_err.reserve(U + 16) || d_flag.reserve((na + 1) * 4) || d_ex.reserve((na + 2) * 8) || d_tmp.reserve(scan_tmp_words(na) * 8) || (del_prev && d_prev.reserve(U + 16))) {
        ctx->err = "device allocation failed in sg_ec_filter"; return SG_E_NOMEM;
    }
    if (del_prev) CK(cudaMemcpyAsync(d_prev.p, del_prev, U, cudaMemcpyHostToDevice, st));
    if (dbg) { fprintf(stderr, "[sg_ec_filter] allocations %.3f s\n", now() - t0); t0 = now(); }
    ctx->t_begin(SG_T_EC);
    if (U) ec_filter_kernel<<<(unsigned) ((U + 255) / 256), 256, 0, st>>>((const uint64_t *) b->arc_out.p, na, (const uint32_t *) b->scm_cov.p,
            del_prev ? (const uint8_t *) d_prev.p : nullptr, U, err_mer_c, max_err_c, err_arc_c, max_arc_f, (uint8_t *) d_err.p);
    uint64_t n_live = 0;
    if (na) {
        ec_live_flag_kernel<<<(unsigned) ((na + 255) / 256), 256, 0, st>>>((const uint64_t *) b->arc_out.p, na, (const uint8_t *) d_err.p,
                del_prev ? (const uint8_t *) d_prev.p : nullptr, (uint32_t *) d_flag.p);
        int nl = launch_scan_u32_u64((const uint32_t *) d_flag.p, (uint64_t *) d_ex.p, na, (uint64_t *) d_tmp.p, st);
        if (nl < 0) return nl;
        ctx->count_launch(SG_T_EC, nl + 1);
        CK(cudaMemcpyAsync(&n_live, (uint64_t *) d_ex.p + na, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (d_live.reserve((n_live + 1) * 32)) { ctx->err = "device allocation failed in sg_ec_filter"; return SG_E_NOMEM; }
        ec_live_pack_kernel<<<(unsigned) ((na + 255) / 256), 256, 0, st>>>((const uint64_t *) b->arc_out.p, na, (const uint32_t *) d_flag.p,
                (const uint64_t *) d_ex.p, (uint64_t *) d_live.p);
        ctx->count_launch(SG_T_EC, 1);
    }
    ctx->count_launch(SG_T_EC, 1);
    ctx->t_end(SG_T_EC);
    out->err = (uint8_t *) malloc(U ? U : 1);
    out->arcs4 = (uint64_t *) malloc((n_live + 1) * 32);
    if (!out->err || !out->arcs4) { sg_ec_filter_free(out); return SG_E_NOMEM; }
    if (U) CK(cudaMemcpyAsync(out->err, d_err.p, U, cudaMemcpyDeviceToHost, st));
    if (n_live) CK(cudaMemcpyAsync(out->arcs4, d_live.p, n_live * 32, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (dbg) { fprintf(stderr, "[sg_ec_filter] kernels + download %.3f s\n", now() - t0); t0 = now(); }
    out->n_arcs_all = na;
    out->n_live = n_live;
    b->h2d_bytes += del_prev ? U : 0;
    b->d2h_bytes += U + n_live * 32;
    return SG_OK;
}


extern "C" int sg_arc_votes(sg_batch *b, uint64_t n, const uint64_t *arcs4, int32_t *dist, uint8_t *flag)
{
    if (!b || (n && (!arcs4 || !dist || !flag))) return SG_E_ARG;
    if (!b->counted || b->adopted || b->keys_are_ids || !b->sorted) return SG_E_STATE;     // the occurrence lists of the last sg_count
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return SG_OK;
    if (b->ec_live.reserve(n * 32 + 32) || b->ec_flag.reserve(n * 4 + 16) || b->ec_err.reserve(n + 16)) { ctx->err = "device allocation failed in sg_arc_votes"; return SG_E_NOMEM; }
    CK(cudaMemcpyAsync(b->ec_live.p, arcs4, n * 32, cudaMemcpyHostToDevice, st));
    VoteArgs A;
    A.arcs4 = (const uint64_t *) b->ec_live.p; A.n = n;
    A.socc = (const uint64_t *) b->socc.p; A.occ_off = (const uint64_t *) b->scm_occ_off.p;
    A.scm_off = (const uint64_t *) b->scm_off.p; A.kid = (const uint64_t *) b->kid.p; A.m_pos = (const uint32_t *) b->m_pos.p;
    A.sid_base = b->sid_base; A.n_reads = b->n_reads;
    A.dist = (int32_t *) b->ec_flag.p; A.flag = (uint8_t *) b->ec_err.p;
    ctx->t_begin(SG_T_EC);
    arc_vote_kernel<<<(unsigned) std::min<uint64_t>((n + 3) / 4, 148ull * 16ull), 128, 0, st>>>(A);
    ctx->count_launch(SG_T_EC, 1);
    ctx->t_end(SG_T_EC);
    CK(cudaMemcpyAsync(dist, b->ec_flag.p, n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(flag, b->ec_err.p, n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    b->h2d_bytes += n * 32; b->d2h_bytes += n * 5;
    return SG_OK;
}

extern "C" void sg_ec_filter_free(sg_ec_filter_out_t *o)
{
    if (!o) return;
    free(o->err); free(o->arcs4);
    o->err = 0; o->arcs4 = 0;
}

extern "C" void sg_ec_result_free(sg_ec_result_t *r)
{
    if (!r) return;
    free(r->out_off); free(r->out_n); free(r->out_k); free(r->out_p);
    r->out_off = 0; r->out_n = 0; r->out_k = 0; r->out_p = 0;
}
