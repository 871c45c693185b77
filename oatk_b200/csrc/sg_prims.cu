// sg_prims.cu -- device-wide primitives written for this library: exclusive scans
// and a stable LSD radix sort of (uint64 key, uint64 value) pairs.
//
// The reference orders its syncmer tuples with one 128-bit qsort (reference
// syncmer.c:1410-1419). Tuples are generated here already in (sid, idx) order,
// so a STABLE sort on the 64-bit hash alone reproduces that order; the sort is
// eight 8-bit counting passes (histogram per tile, scan, stable scatter with
// warp match-any ranking).
#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

// ---------------------------------------------------------------- scans
constexpr int SCAN_NT = 256, SCAN_IPT = 8, SCAN_TILE = SCAN_NT * SCAN_IPT;

struct LoadU32 {
    const uint32_t *p;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return p[i]; }
};
struct LoadCap64 {      // roundup64(off[i+1] - off[i])
    const uint64_t *off;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return (off[i + 1] - off[i] + 63ull) & ~63ull; }
};

template <typename T>
__device__ __forceinline__ T block_exscan(T v, T *smem /* NT/32 */, T *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    T inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T t = __shfl_up_sync(SG_FULL, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    T base = 0, tot = 0;
    for (int w = 0; w < nw; ++w) { T t = smem[w]; if (w < wid) base += t; tot += t; }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

template <typename Load>
__global__ void __launch_bounds__(SCAN_NT) scan_reduce_kernel(Load ld, uint64_t n, uint64_t *sums)
{
    __shared__ uint64_t sm[SCAN_NT / 32];
    const uint64_t base = (uint64_t) blockIdx.x * SCAN_TILE;
    uint64_t v = 0;
    for (int j = 0; j < SCAN_IPT; ++j) {
        uint64_t i = base + (uint64_t) j * SCAN_NT + threadIdx.x;
        if (i < n) v += ld(i);
    }
    uint64_t tot;
    block_exscan<uint64_t>(v, sm, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(uint64_t *sums, uint64_t nb)
{
    __shared__ uint64_t sm[32];
    uint64_t carry = 0;
    for (uint64_t b0 = 0; b0 < nb; b0 += 1024) {
        uint64_t i = b0 + threadIdx.x;
        uint64_t v = i < nb ? sums[i] : 0, tot;
        uint64_t ex = block_exscan<uint64_t>(v, sm, &tot);
        if (i < nb) sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) sums[nb] = carry;
}

template <typename Load>
__global__ void __launch_bounds__(SCAN_NT) scan_apply_kernel(Load ld, uint64_t n, const uint64_t *sums, uint64_t *out)
{
    __shared__ uint64_t sm[SCAN_NT / 32];
    // thread-contiguous items so that the scan order equals the index order
    const uint64_t base = (uint64_t) blockIdx.x * SCAN_TILE + (uint64_t) threadIdx.x * SCAN_IPT;
    uint64_t v[SCAN_IPT], s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_IPT; ++j) { v[j] = base + j < n ? ld(base + j) : 0; s += v[j]; }
    uint64_t tot;
    uint64_t ex = block_exscan<uint64_t>(s, sm, &tot) + sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_IPT; ++j) { if (base + j < n) out[base + j] = ex; ex += v[j]; }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = sums[gridDim.x];
}

size_t scan_tmp_words(uint64_t n) { return (size_t) ((n + SCAN_TILE - 1) / SCAN_TILE + 2); }

template <typename Load>
static int scan_generic(Load ld, uint64_t *out, uint64_t n, uint64_t *tmp, cudaStream_t st)
{
    if (n == 0) { cudaMemsetAsync(out, 0, sizeof(uint64_t), st); return 0; }
    const uint64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_reduce_kernel<Load><<<(unsigned) nb, SCAN_NT, 0, st>>>(ld, n, tmp);
    scan_sums_kernel<<<1, 1024, 0, st>>>(tmp, nb);
    scan_apply_kernel<Load><<<(unsigned) nb, SCAN_NT, 0, st>>>(ld, n, tmp, out);
    return 3;
}

int launch_scan_u32_u64(const uint32_t *in, uint64_t *out, uint64_t n, uint64_t *tmp, cudaStream_t st)
{
    return scan_generic(LoadU32{in}, out, n, tmp, st);
}

int launch_capacity_offsets(const uint64_t *off, uint64_t *hoff, uint64_t n, uint64_t *tmp, cudaStream_t st)
{
    return scan_generic(LoadCap64{off}, hoff, n, tmp, st);
}

// ---------------------------------------------------------------- radix sort
constexpr int RS_NT = 256, RS_NW = RS_NT / 32, RS_IPT = 8, RS_TILE = RS_NT * RS_IPT, RS_BINS = 256;

// counts[d * ntiles + tile]
__global__ void __launch_bounds__(RS_NT) rs_hist_kernel(const uint64_t *key, uint64_t n, int shift, uint32_t *counts, uint32_t ntiles)
{
    __shared__ uint32_t h[RS_BINS];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t) blockIdx.x * RS_TILE;
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        uint64_t i = base + (uint64_t) j * RS_NT + threadIdx.x;
        if (i < n) atomicAdd(&h[(key[i] >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    counts[(uint64_t) threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// in-place exclusive scan of a uint32 array (total < 2^32), three-phase like the scans above
__global__ void __launch_bounds__(SCAN_NT) rs_scan_reduce(const uint32_t *a, uint64_t n, uint32_t *sums)
{
    __shared__ uint32_t sm[SCAN_NT / 32];
    const uint64_t base = (uint64_t) blockIdx.x * SCAN_TILE;
    uint32_t v = 0;
    for (int j = 0; j < SCAN_IPT; ++j) { uint64_t i = base + (uint64_t) j * SCAN_NT + threadIdx.x; if (i < n) v += a[i]; }
    uint32_t tot;
    block_exscan<uint32_t>(v, sm, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(1024) rs_scan_sums(uint32_t *sums, uint64_t nb)
{
    __shared__ uint32_t sm[32];
    uint32_t carry = 0;
    for (uint64_t b0 = 0; b0 < nb; b0 += 1024) {
        uint64_t i = b0 + threadIdx.x;
        uint32_t v = i < nb ? sums[i] : 0, tot;
        uint32_t ex = block_exscan<uint32_t>(v, sm, &tot);
        if (i < nb) sums[i] = carry + ex;
        carry += tot;
    }
}
__global__ void __launch_bounds__(SCAN_NT) rs_scan_apply(uint32_t *a, uint64_t n, const uint32_t *sums)
{
    __shared__ uint32_t sm[SCAN_NT / 32];
    const uint64_t base = (uint64_t) blockIdx.x * SCAN_TILE + (uint64_t) threadIdx.x * SCAN_IPT;
    uint32_t v[SCAN_IPT], s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_IPT; ++j) { v[j] = base + j < n ? a[base + j] : 0; s += v[j]; }
    uint32_t tot;
    uint32_t ex = block_exscan<uint32_t>(s, sm, &tot) + sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_IPT; ++j) { if (base + j < n) a[base + j] = ex; ex += v[j]; }
}

// stable scatter: warp w of a tile owns elements [w*32*IPT, (w+1)*32*IPT) of the tile and
// walks them 32 at a time in index order; match_any groups equal digits inside
// a round, a warp-private counter carries the rank across rounds
__global__ void __launch_bounds__(RS_NT) rs_scatter_kernel(const uint64_t *key, const uint64_t *val, uint64_t *okey, uint64_t *oval,
        uint64_t n, int shift, const uint32_t *offsets, uint32_t ntiles)
{
    __shared__ uint32_t wc[RS_NW][RS_BINS];      // per-warp running digit counts
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_NW * RS_BINS; i += RS_NT) (&wc[0][0])[i] = 0;
    __syncthreads();
    const uint64_t wbase = (uint64_t) blockIdx.x * RS_TILE + (uint64_t) wid * 32 * RS_IPT;
    uint64_t kreg[RS_IPT], vreg[RS_IPT];
    uint32_t rank[RS_IPT];
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        const uint64_t i = wbase + (uint64_t) j * 32 + lane;
        const bool ok = i < n;
        kreg[j] = ok ? key[i] : 0;
        vreg[j] = ok ? val[i] : 0;
        const uint32_t d = ok ? (uint32_t) ((kreg[j] >> shift) & 0xFFu) : 0x100u;   // 0x100: out of range, own group
        const uint32_t peers = __match_any_sync(SG_FULL, d);
        const uint32_t before = __popc(peers & ((1u << lane) - 1u));
        uint32_t base = 0;
        if (ok) base = wc[wid][d];
        __syncwarp();
        if (ok && before == 0) wc[wid][d] = base + __popc(peers);
        __syncwarp();
        rank[j] = base + before;
    }
    __syncthreads();
    // exclusive prefix over warps, per digit (thread d handles digit d)
    {
        uint32_t run = offsets[(uint64_t) threadIdx.x * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_NW; ++w) { uint32_t t = wc[w][threadIdx.x]; wc[w][threadIdx.x] = run; run += t; }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        const uint64_t i = wbase + (uint64_t) j * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t) ((kreg[j] >> shift) & 0xFFu);
            const uint64_t o = (uint64_t) wc[wid][d] + rank[j];
            okey[o] = kreg[j];
            oval[o] = vreg[j];
        }
    }
}

// in-place exclusive scan of n uint32 (total < 2^32); sums must hold n / 2048 + 8 words
int launch_exscan_u32(uint32_t *a, uint64_t n, uint32_t *sums, cudaStream_t st)
{
    if (n == 0) return 0;
    const uint64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    rs_scan_reduce<<<(unsigned) nb, SCAN_NT, 0, st>>>(a, n, sums);
    rs_scan_sums<<<1, 1024, 0, st>>>(sums, nb);
    rs_scan_apply<<<(unsigned) nb, SCAN_NT, 0, st>>>(a, n, sums);
    return 3;
}

// keys only (the value rides in the low bits of the key word): same stable scatter, half the traffic
__global__ void __launch_bounds__(RS_NT) rs_scatter_keys_kernel(const uint64_t *key, uint64_t *okey, uint64_t n, int shift, const uint32_t *offsets, uint32_t ntiles)
{
    __shared__ uint32_t wc[RS_NW][RS_BINS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_NW * RS_BINS; i += RS_NT) (&wc[0][0])[i] = 0;
    __syncthreads();
    const uint64_t wbase = (uint64_t) blockIdx.x * RS_TILE + (uint64_t) wid * 32 * RS_IPT;
    uint64_t kreg[RS_IPT];
    uint32_t rank[RS_IPT];
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        const uint64_t i = wbase + (uint64_t) j * 32 + lane;
        const bool ok = i < n;
        kreg[j] = ok ? key[i] : 0;
        const uint32_t d = ok ? (uint32_t) ((kreg[j] >> shift) & 0xFFu) : 0x100u;
        const uint32_t peers = __match_any_sync(SG_FULL, d);
        const uint32_t before = __popc(peers & ((1u << lane) - 1u));
        uint32_t base = 0;
        if (ok) base = wc[wid][d];
        __syncwarp();
        if (ok && before == 0) wc[wid][d] = base + __popc(peers);
        __syncwarp();
        rank[j] = base + before;
    }
    __syncthreads();
    {
        uint32_t run = offsets[(uint64_t) threadIdx.x * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_NW; ++w) { uint32_t t = wc[w][threadIdx.x]; wc[w][threadIdx.x] = run; run += t; }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RS_IPT; ++j) {
        const uint64_t i = wbase + (uint64_t) j * 32 + lane;
        if (i < n) okey[(uint64_t) wc[wid][(uint32_t) ((kreg[j] >> shift) & 0xFFu)] + rank[j]] = kreg[j];
    }
}

// stable LSD radix sort of 64-bit words on bits [begin_bit, end_bit); the result ends in `key`
int launch_sort_keys(uint64_t *key, uint64_t *key_alt, uint64_t n, int begin_bit, int end_bit, uint32_t *tmp, cudaStream_t st)
{
    if (n == 0) return 0;
    const uint32_t ntiles = (uint32_t) ((n + RS_TILE - 1) / RS_TILE);
    const uint64_t m = (uint64_t) ntiles * RS_BINS;
    const uint64_t nb = (m + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t *counts = tmp, *sums = tmp + m;
    int launches = 0, passes = 0;
    uint64_t *k0 = key, *k1 = key_alt;
    for (int shift = begin_bit; shift < end_bit; shift += 8, ++passes) {
        rs_hist_kernel<<<ntiles, RS_NT, 0, st>>>(k0, n, shift, counts, ntiles);
        rs_scan_reduce<<<(unsigned) nb, SCAN_NT, 0, st>>>(counts, m, sums);
        rs_scan_sums<<<1, 1024, 0, st>>>(sums, nb);
        rs_scan_apply<<<(unsigned) nb, SCAN_NT, 0, st>>>(counts, m, sums);
        rs_scatter_keys_kernel<<<ntiles, RS_NT, 0, st>>>(k0, k1, n, shift, counts, ntiles);
        launches += 5;
        uint64_t *t = k0; k0 = k1; k1 = t;
    }
    if (passes & 1) cudaMemcpyAsync(key, k0, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st);
    return launches;
}

size_t sort_tmp_words(uint64_t n)
{
    const uint64_t ntiles = (n + RS_TILE - 1) / RS_TILE;
    const uint64_t m = ntiles * RS_BINS;
    return (size_t) (m + (m + SCAN_TILE - 1) / SCAN_TILE + 8);
}

int launch_sort_pairs(uint64_t *key, uint64_t *val, uint64_t *key_alt, uint64_t *val_alt, uint64_t n,
        int begin_bit, int end_bit, uint32_t *tmp, cudaStream_t st)
{
    if (n == 0) return 0;
    const uint32_t ntiles = (uint32_t) ((n + RS_TILE - 1) / RS_TILE);
    const uint64_t m = (uint64_t) ntiles * RS_BINS;
    const uint64_t nb = (m + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t *counts = tmp, *sums = tmp + m;
    int launches = 0, passes = 0;
    uint64_t *k0 = key, *v0 = val, *k1 = key_alt, *v1 = val_alt;
    for (int shift = begin_bit; shift < end_bit; shift += 8, ++passes) {
        rs_hist_kernel<<<ntiles, RS_NT, 0, st>>>(k0, n, shift, counts, ntiles);
        rs_scan_reduce<<<(unsigned) nb, SCAN_NT, 0, st>>>(counts, m, sums);
        rs_scan_sums<<<1, 1024, 0, st>>>(sums, nb);
        rs_scan_apply<<<(unsigned) nb, SCAN_NT, 0, st>>>(counts, m, sums);
        rs_scatter_kernel<<<ntiles, RS_NT, 0, st>>>(k0, v0, k1, v1, n, shift, counts, ntiles);
        launches += 5;
        uint64_t *t;
        t = k0; k0 = k1; k1 = t;
        t = v0; v0 = v1; v1 = t;
    }
    if (passes & 1) {       // result sits in the alternate buffers: bring it home
        cudaMemcpyAsync(key, k0, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(val, v0, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st);
    }
    return launches;
}

} // namespace sg
