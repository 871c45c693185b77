// sg_comm.cu -- the multi-GPU step of the syncmer counter and of the arc tally, in C over NCCL.
//
// The reference is one process (no communication layer exists); what shards here is
// the read set: GPU r holds a contiguous block of reads (sid = global read index) and extracts it alone.
// Two things need all occurrences of a key on one GPU:
//   * collect_syncmer_from_reads (syncmer.c:1397-1451): ids are ranks in hash order, so the tuples
//     (hash, occ, s_mer, fingerprint) are range-partitioned on the hash and exchanged once; GPU r then owns
//     hash range r, its local hash order is the reference's global order restricted to that range, and
//     id = local rank + distinct k-mers of the lower ranges. (occ, id) pairs travel back the same way.
//   * make_syncmer_graph's arc tally (syncasm.c:236-282): every GPU tallies the neighbouring pairs of its
//     own reads (ids are global by then), the (canonical pair, count) entries are partitioned on a hash of
//     the pair and exchanged once, the owner adds them up and applies count >= a * min(cov(v0), cov(v1))
//     with the coverages of all ranges (one all-gather); the survivors are collected and sorted on one rank.
// Transport: grouped ncclSend / ncclRecv on the context's stream (an all-to-all-v), ncclAllGather for the
// count matrix, ncclBroadcast for the coverage ranges. Counts live on the device; the host reads the
// world x world count matrix ONCE per exchange (NCCL takes element counts as host arguments) and nothing
// else: the id base is summed on the device from the all-gathered distinct counts.
//
// NCCL is loaded with dlopen at the first sg_comm_* call, so libsyncgpu.so has no link-time dependency on
// it and single-GPU users never touch it. Two ways to form the communicator:
//   sg_comm_init_rank  one process per GPU (bench.py under torchrun): rank 0 makes the unique id with
//                      sg_comm_unique_id and hands it to the others by whatever means the launcher has
//   sg_comm_init_all   one process, one host thread per GPU (every data call below blocks until its peers have made it too, so
//                      the calls for different GPUs must come from different threads); nothing in this repository uses it yet:
//                      the host layer's syncasm() runs on one GPU
// Every sg_comm_* data call is collective: all ranks (or all threads) make the same calls in the same order.
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <mutex>
#include <nccl.h>
#include "sg_common.cuh"
#include "sg_internal.h"
#include "sg_host.h"
#include "sg_table.cuh"

namespace sg {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};

static NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // a process that already carries NCCL (torch bundles one) gets that copy: same soname
        const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; names[i] && !api.lib; ++i) api.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) { api.why = std::string("libnccl.so.2 not found: ") + dlerror(); return; }
#define SG_SYM(field, name) do { *(void **) &api.field = dlsym(api.lib, name); \
        if (!api.field) { api.why = std::string("NCCL lacks ") + name; api.lib = nullptr; return; } } while (0)
        SG_SYM(GetUniqueId, "ncclGetUniqueId"); SG_SYM(CommInitRank, "ncclCommInitRank"); SG_SYM(CommInitAll, "ncclCommInitAll");
        SG_SYM(CommDestroy, "ncclCommDestroy"); SG_SYM(GroupStart, "ncclGroupStart"); SG_SYM(GroupEnd, "ncclGroupEnd");
        SG_SYM(Send, "ncclSend"); SG_SYM(Recv, "ncclRecv"); SG_SYM(AllGather, "ncclAllGather"); SG_SYM(AllReduce, "ncclAllReduce");
        SG_SYM(Broadcast, "ncclBroadcast"); SG_SYM(GetErrorString, "ncclGetErrorString");
#undef SG_SYM
    });
    return api.lib ? &api : (api.why.empty() ? nullptr : &api);
}

} // namespace sg

using namespace sg;

struct sg_comm {
    sg_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    DevBuf cnt_dev, mat_dev, uniq_dev, recv, back, gather, covs, scratch, pcount, pcursor, ppairs, split_counts, split_sums, perm, sent_dev;
    bool fast = false;                       // the last exchange split the extract's records directly (perm holds where each sent record came from)
    uint64_t *mat_host = nullptr;            // pinned, world x world (+ world for the distinct counts)
    std::vector<uint64_t> send_counts, recv_counts;   // of the last tuple exchange
    std::vector<uint64_t> uniq;              // distinct k-mers per rank after sg_comm_return_ids
    uint64_t bytes_sent = 0;
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return SG_E_CUDA; } } while (0)
#define NK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { \
    ctx->err = std::string(#call) + ": " + N->GetErrorString(r_); return SG_E_COMM; } } while (0)
#define RS(buf, bytes) do { if ((buf).reserve(bytes)) { ctx->err = "device allocation of " + std::to_string((size_t)(bytes)) + " bytes failed"; return SG_E_NOMEM; } } while (0)
static inline unsigned nblk(uint64_t n, unsigned t) { return (unsigned) ((n + t - 1) / t); }

namespace sg {

// ---- the exchange without detours: the 32-byte records kmerhash_kernel wrote are split by destination in one stable
// pass (per-tile counts, one scan, ranked scatter as in the radix sort) straight into the send buffer, together with the
// index each record came from; what comes back after counting is then just the id of every sent record, in the order it
// was sent, and the saved indices put the ids into read order. No (key, value) sort, no gather, no (occ, id) pairs.
constexpr int SP_NT = 256, SP_NW = SP_NT / 32, SP_IPT = 8, SP_TILE = SP_NT * SP_IPT;
__device__ __forceinline__ uint32_t part_of(uint64_t h, uint32_t world) { return (uint32_t) __umul64hi((h >> 1) << 1, (uint64_t) world); }

__global__ void __launch_bounds__(SP_NT) split_hist_kernel(const uint64_t *key, uint64_t n, uint32_t world, uint32_t *counts, uint32_t ntiles)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t) blockIdx.x * SP_TILE;
#pragma unroll
    for (int j = 0; j < SP_IPT; ++j) {
        const uint64_t i = base + (uint64_t) j * SP_NT + threadIdx.x;
        if (i < n) atomicAdd(&h[part_of(key[i], world)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < world) counts[(uint64_t) threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// counts[] holds exclusive offsets now; part p starts at counts[p * ntiles]
__global__ void split_counts_kernel(const uint32_t *offsets, uint32_t ntiles, int world, uint64_t n, uint64_t *cnt)
{
    const int p = threadIdx.x;
    if (p >= world) return;
    const uint64_t a = offsets[(uint64_t) p * ntiles], b = p + 1 < world ? offsets[(uint64_t) (p + 1) * ntiles] : n;
    cnt[p] = b - a;
}

__global__ void __launch_bounds__(SP_NT) split_scatter_kernel(const uint64_t *key, const ulonglong4 *rec, uint64_t n, uint32_t world,
        const uint32_t *offsets, uint32_t ntiles, ulonglong4 *out, uint32_t *perm)
{
    __shared__ uint32_t wc[SP_NW][256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < SP_NW * 256; i += SP_NT) (&wc[0][0])[i] = 0;
    __syncthreads();
    const uint64_t wbase = (uint64_t) blockIdx.x * SP_TILE + (uint64_t) wid * 32 * SP_IPT;
    uint32_t rank[SP_IPT], dig[SP_IPT];
#pragma unroll
    for (int j = 0; j < SP_IPT; ++j) {
        const uint64_t i = wbase + (uint64_t) j * 32 + lane;
        const bool ok = i < n;
        const uint32_t d = ok ? part_of(key[i], world) : 0x100u;
        const uint32_t peers = __match_any_sync(SG_FULL, d);
        const uint32_t before = __popc(peers & ((1u << lane) - 1u));
        uint32_t base = 0;
        if (ok) base = wc[wid][d];
        __syncwarp();
        if (ok && before == 0) wc[wid][d] = base + __popc(peers);
        __syncwarp();
        rank[j] = base + before;
        dig[j] = d;
    }
    __syncthreads();
    if (threadIdx.x < world) {
        uint32_t run = offsets[(uint64_t) threadIdx.x * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < SP_NW; ++w) { const uint32_t t = wc[w][threadIdx.x]; wc[w][threadIdx.x] = run; run += t; }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SP_IPT; ++j) {
        const uint64_t i = wbase + (uint64_t) j * 32 + lane;
        if (i < n) {
            const uint64_t o = (uint64_t) wc[wid][dig[j]] + rank[j];
            out[o] = rec[i];
            perm[o] = (uint32_t) i;
        }
    }
}

// ids of the records this rank sent, back in the order they were sent (part after part): kid_local[perm[j]] = id + base of the owner
__global__ void __launch_bounds__(256) ids_unpermute_kernel(const uint64_t *back, const uint32_t *perm, uint64_t n, const uint64_t *sent /* world counts */,
        const uint64_t *uniq /* world */, int world, uint64_t *kid_local)
{
    __shared__ uint64_t s_end[256], s_base[256];
    if (threadIdx.x == 0) {
        uint64_t e = 0, b = 0;
        for (int p = 0; p < world; ++p) { e += sent[p]; s_end[p] = e; s_base[p] = b; b += uniq[p]; }
    }
    __syncthreads();
    const uint64_t j = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int lo = 0, hi = world - 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_end[mid] <= j) lo = mid + 1; else hi = mid; }
    kid_local[perm[j]] = back[j] + (s_base[lo] << 1);
}

// end offsets of the parts (0 = the part is empty) -> counts
__global__ void ends_to_counts_kernel(const unsigned long long *ends, int n_parts, uint64_t *counts)
{
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long prev = 0;
    for (int p = 0; p < n_parts; ++p) {
        const unsigned long long e = ends[p] ? ends[p] : prev;
        counts[p] = e - prev;
        prev = e;
    }
}

// (occ, (local id + distinct k-mers of the lower ranks) << 1) for the adopted tuples, the base summed on the device
__global__ void __launch_bounds__(256) pair_pack_base_kernel(const uint64_t *occ, const uint64_t *kid, uint64_t n, const uint64_t *uniq, int rank, uint64_t *pairs)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t base = 0;
    for (int r = 0; r < rank; ++r) base += uniq[r];
    pairs[2 * i] = occ[i]; pairs[2 * i + 1] = kid[i] + (base << 1);
}

// owner of a canonical arc key: a multiplicative mix spreads the dense ids over the ranks
__device__ __forceinline__ uint32_t arc_owner(uint64_t key, uint32_t world)
{
    key ^= key >> 33; key *= 0xff51afd7ed558ccdull; key ^= key >> 33;
    return (uint32_t) __umul64hi(key, (uint64_t) world);
}
__global__ void __launch_bounds__(256) arc_part_count_kernel(const uint64_t *tk, uint64_t nslots, uint32_t world, unsigned long long *counts)
{
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < nslots; i += (uint64_t) gridDim.x * blockDim.x)
        if (tk[i] != EMPTY_KEY) atomicAdd(&h[arc_owner(tk[i], world)], 1u);
    __syncthreads();
    if (threadIdx.x < world && h[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long) h[threadIdx.x]);
}
__global__ void offsets_kernel(const unsigned long long *counts, int n, unsigned long long *cursor)
{
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long s = 0;
    for (int p = 0; p < n; ++p) { cursor[p] = s; s += counts[p]; }
}
__global__ void __launch_bounds__(256) arc_part_scatter_kernel(const uint64_t *tk, const uint32_t *tv, uint64_t nslots, uint32_t world,
        unsigned long long *cursor, uint64_t *pairs)
{
    const int lane = threadIdx.x & 31;
    for (uint64_t i0 = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) - lane; i0 < nslots; i0 += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t i = i0 + lane;
        const uint64_t key = i < nslots ? tk[i] : EMPTY_KEY;
        const bool live = key != EMPTY_KEY;
        const uint32_t own = live ? arc_owner(key, world) : 0xffffffffu;
        // lanes bound for the same rank take consecutive places with one atomic
        const unsigned same = __match_any_sync(SG_FULL, own);
        const int leader = __ffs(same) - 1;
        unsigned long long base = 0;
        if (live && lane == leader) base = atomicAdd(cursor + own, (unsigned long long) __popc(same));
        base = __shfl_sync(SG_FULL, base, leader);
        if (live) {
            const unsigned long long o = base + __popc(same & ((1u << lane) - 1u));
            pairs[2 * o] = key; pairs[2 * o + 1] = tv[i];
        }
    }
}

// host-side steps of sg_arcs.cu shared with the multi-GPU path
int arcs_tally_local(sg_batch *b, const uint64_t *kid, uint64_t *nslots_out);
int arcs_merge_pairs(sg_batch *b, const uint64_t *pairs, uint64_t n, uint64_t *nslots_out);
int arcs_emit(sg_batch *b, uint64_t nslots, const uint32_t *cov, uint32_t min_k_cov, double min_a_cov_f, uint64_t *n_arcs);
int arcs_sort_unpack(sg_batch *b, uint64_t na);
int tuples_partition_device(sg_batch *b, int n_parts);

} // namespace sg

extern "C" {

int sg_comm_unique_id(void *id128)
{
    NcclApi *N = nccl_api();
    if (!N || !N->lib || !id128) return SG_E_COMM;
    ncclUniqueId id;
    if (N->GetUniqueId(&id) != ncclSuccess) return SG_E_COMM;
    static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
    memcpy(id128, &id, 128);
    return SG_OK;
}

static sg_comm *comm_new(sg_ctx *ctx, int world, int rank)
{
    sg_comm *c = new sg_comm();
    c->ctx = ctx; c->world = world; c->rank = rank;
    if (cudaMallocHost(&c->mat_host, sizeof(uint64_t) * ((size_t) world * world + 2 * world + 1024 + 8)) != cudaSuccess) { delete c; return nullptr; }
    return c;
}

int sg_comm_init_rank(sg_ctx *ctx, int world, int rank, const void *id128, sg_comm **out)
{
    if (!ctx || !out || !id128 || world < 1 || world > 256 || rank < 0 || rank >= world) return SG_E_ARG;
    *out = nullptr;
    NcclApi *N = nccl_api();
    if (!N || !N->lib) { ctx->err = N ? N->why : "NCCL not available"; return SG_E_COMM; }
    CK(cudaSetDevice(ctx->device));
    sg_comm *c = comm_new(ctx, world, rank);
    if (!c) { ctx->err = "pinned allocation failed"; return SG_E_NOMEM; }
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = N->CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) { ctx->err = std::string("ncclCommInitRank: ") + N->GetErrorString(r); cudaFreeHost(c->mat_host); delete c; return SG_E_COMM; }
    *out = c;
    return SG_OK;
}

int sg_comm_init_all(sg_ctx **ctxs, int n, sg_comm **out)
{
    if (!ctxs || !out || n < 1 || n > 256) return SG_E_ARG;
    NcclApi *N = nccl_api();
    sg_ctx *ctx = ctxs[0];
    if (!N || !N->lib) { ctx->err = N ? N->why : "NCCL not available"; return SG_E_COMM; }
    std::vector<int> devs(n);
    std::vector<ncclComm_t> comms(n);
    for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
    NK(N->CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) {
        cudaSetDevice(devs[i]);
        out[i] = comm_new(ctxs[i], n, i);
        if (!out[i]) { ctx->err = "pinned allocation failed"; return SG_E_NOMEM; }
        out[i]->comm = comms[i];
    }
    return SG_OK;
}

void sg_comm_destroy(sg_comm *c)
{
    if (!c) return;
    NcclApi *N = nccl_api();
    cudaSetDevice(c->ctx->device);
    if (N && N->lib && c->comm) N->CommDestroy(c->comm);
    if (c->mat_host) cudaFreeHost(c->mat_host);
    delete c;
}

int sg_comm_rank(sg_comm *c) { return c ? c->rank : -1; }
int sg_comm_world(sg_comm *c) { return c ? c->world : 0; }
uint64_t sg_comm_bytes_sent(sg_comm *c) { return c ? c->bytes_sent : 0; }

// all-to-all-v of `width` uint64 per item: items for rank p start at soff[p] of `send`, those from rank p land at roff[p] of `recv`
static int all_to_all_v(sg_comm *c, const uint64_t *send, const std::vector<uint64_t> &scount, uint64_t *recv, const std::vector<uint64_t> &rcount, int width)
{
    sg_ctx *ctx = c->ctx;
    NcclApi *N = nccl_api();
    cudaStream_t st = ctx->stream;
    uint64_t so = 0, ro = 0;
    NK(N->GroupStart());
    for (int p = 0; p < c->world; ++p) {
        if (scount[p]) NK(N->Send(send + so * width, scount[p] * width, ncclUint64, p, c->comm, st));
        if (rcount[p]) NK(N->Recv(recv + ro * width, rcount[p] * width, ncclUint64, p, c->comm, st));
        if (p != c->rank) c->bytes_sent += scount[p] * width * 8;
        so += scount[p]; ro += rcount[p];
    }
    NK(N->GroupEnd());
    return SG_OK;
}

// every rank contributes `world` counts (device); returns the world x world matrix on the host: m[r * world + p] = items r sends to p
static int gather_count_matrix(sg_comm *c, const uint64_t *counts_dev)
{
    sg_ctx *ctx = c->ctx;
    NcclApi *N = nccl_api();
    cudaStream_t st = ctx->stream;
    const int W = c->world;
    RS(c->mat_dev, (size_t) W * W * 8);
    NK(N->AllGather(counts_dev, c->mat_dev.p, W, ncclUint64, c->comm, st));
    CK(cudaMemcpyAsync(c->mat_host, c->mat_dev.p, (size_t) W * W * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SG_OK;
}

int sg_comm_exchange_tuples(sg_comm *c, sg_batch *b)
{
    if (!c || !b) return SG_E_ARG;
    if (!b->extracted) return SG_E_STATE;
    sg_ctx *ctx = c->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    const int W = c->world;
    ctx->t_begin(SG_T_EXCH);
    ctx->lap("begin");
    int rc;
    const uint64_t N = b->n_syncmers;
    c->fast = b->tup_valid && N < (1ull << 32) && !getenv("SG_COMM_PAIRS");
    RS(c->cnt_dev, (size_t) W * 8);
    if (c->fast) {
        const uint32_t ntiles = (uint32_t) ((N + SP_TILE - 1) / SP_TILE);
        const uint64_t m = (uint64_t) std::max<uint32_t>(ntiles, 1) * W;
        RS(c->split_counts, (m + 8) * 4); RS(c->split_sums, (m / 2048 + 16) * 4); RS(c->perm, (N + 1) * 4); RS(b->tuples, (N + 1) * 32);
        if (N) {
            split_hist_kernel<<<ntiles, SP_NT, 0, st>>>((const uint64_t *) b->key.p, N, (uint32_t) W, (uint32_t *) c->split_counts.p, ntiles);
            const int nl = launch_exscan_u32((uint32_t *) c->split_counts.p, m, (uint32_t *) c->split_sums.p, st);
            split_counts_kernel<<<1, 256, 0, st>>>((const uint32_t *) c->split_counts.p, ntiles, W, N, (uint64_t *) c->cnt_dev.p);
            split_scatter_kernel<<<ntiles, SP_NT, 0, st>>>((const uint64_t *) b->key.p, (const ulonglong4 *) b->tup.p, N, (uint32_t) W,
                    (const uint32_t *) c->split_counts.p, ntiles, (ulonglong4 *) b->tuples.p, (uint32_t *) c->perm.p);
            ctx->count_launch(SG_T_EXCH, 3 + nl);
        } else CK(cudaMemsetAsync(c->cnt_dev.p, 0, (size_t) W * 8, st));
        b->sorted = false;
    } else {
        rc = tuples_partition_device(b, W);                 // b->tuples grouped by destination, b->part_counts = end offsets
        if (rc) return rc;
        ends_to_counts_kernel<<<1, 32, 0, st>>>((const unsigned long long *) b->part_counts.p, W, (uint64_t *) c->cnt_dev.p);
        ctx->count_launch(SG_T_SORT, 1);
    }
    ctx->lap("split");
    rc = gather_count_matrix(c, (const uint64_t *) c->cnt_dev.p);
    if (rc) return rc;
    ctx->lap("counts");
    c->send_counts.assign(W, 0); c->recv_counts.assign(W, 0);
    uint64_t total = 0;
    for (int p = 0; p < W; ++p) {
        c->send_counts[p] = c->mat_host[(size_t) c->rank * W + p];
        c->recv_counts[p] = c->mat_host[(size_t) p * W + c->rank];
        total += c->recv_counts[p];
    }
    if (c->fast) {
        RS(c->sent_dev, (size_t) W * 8);                     // how many records went to each rank: the id return needs it after mat_dev has been reused
        CK(cudaMemcpyAsync(c->sent_dev.p, c->cnt_dev.p, (size_t) W * 8, cudaMemcpyDeviceToDevice, st));
        // the records arrive in the layout the packed sort of sg_count gathers from: they land where it looks for them.
        // (The extract's own records were consumed by the split; room for a quarter more than a fair share, grow-only.)
        RS(b->tup, (std::max<uint64_t>(total, N) + std::max<uint64_t>(total, N) / 4 + 1024) * 32);
        b->tup_valid = false;
        rc = all_to_all_v(c, (const uint64_t *) b->tuples.p, c->send_counts, (uint64_t *) b->tup.p, c->recv_counts, 4);
        if (rc) return rc;
        b->atup_valid = true; b->asoa_valid = false;
        b->adopted = true; b->n_adopted = total;
        {
            // this rank's hash range starts at ceil(rank * 2^64 / W) and is at most 2^64 / W + 1 long: inside it the hashes can be
            // moved up by floor(log2 W) - 1 bits, which is what the packed sort of sg_count orders them by
            const unsigned __int128 lo = (((unsigned __int128) c->rank << 64) + (unsigned) W - 1) / (unsigned) W;
            int lg = 0;
            while ((2 << lg) <= W) ++lg;
            b->range_lo = (uint64_t) lo;
            b->range_lsh = lg > 0 ? lg - 1 : 0;
        }
        b->sorted = b->counted = false;
        ctx->lap("all_to_all");
        ctx->t_end(SG_T_EXCH);
        ctx->laps_print("exchange");
        return SG_OK;
    }
    RS(c->recv, (total + 1) * 32);
    rc = all_to_all_v(c, (const uint64_t *) b->tuples.p, c->send_counts, (uint64_t *) c->recv.p, c->recv_counts, 4);
    if (rc) return rc;
    rc = sg_tuples_adopt(b, c->recv.p, total);
    ctx->t_end(SG_T_EXCH);
    return rc;
}

int sg_comm_return_ids(sg_comm *c, sg_batch *b, uint64_t *id_base, uint64_t *n_unique_total)
{
    if (!c || !b) return SG_E_ARG;
    if (!b->counted || !b->adopted) return SG_E_STATE;
    if ((int) c->send_counts.size() != c->world) return SG_E_STATE;
    sg_ctx *ctx = c->ctx;
    NcclApi *N = nccl_api();
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    const int W = c->world;
    RS(c->uniq_dev, (size_t) (W + 1) * 8);
    ctx->t_begin(SG_T_IDS);
    ctx->lap("begin");
    uint64_t *mine = c->mat_host + (size_t) W * W + W;       // pinned scratch word
    *mine = b->n_unique;
    CK(cudaMemcpyAsync((uint64_t *) c->uniq_dev.p + W, mine, 8, cudaMemcpyHostToDevice, st));
    NK(N->AllGather((const uint64_t *) c->uniq_dev.p + W, c->uniq_dev.p, 1, ncclUint64, c->comm, st));
    ctx->lap("allgather");
    const uint64_t n = b->n_adopted;
    int rc;
    if (c->fast) {
        // ids of the adopted records, in the order they arrived, go back as they are (8 bytes each)
        RS(c->back, (b->n_syncmers + 1) * 8);
        rc = all_to_all_v(c, (const uint64_t *) b->kid.p, c->recv_counts, (uint64_t *) c->back.p, c->send_counts, 1);
        if (rc) return rc;
        ctx->lap("all_to_all");
        CK(cudaMemcpyAsync(c->mat_host + (size_t) W * W, c->uniq_dev.p, (size_t) W * 8, cudaMemcpyDeviceToHost, st));
        RS(b->kid_local, (b->n_syncmers + 1) * 8);
        if (b->n_syncmers) {
            ids_unpermute_kernel<<<nblk(b->n_syncmers, 256), 256, 0, st>>>((const uint64_t *) c->back.p, (const uint32_t *) c->perm.p, b->n_syncmers,
                    (const uint64_t *) c->sent_dev.p, (const uint64_t *) c->uniq_dev.p, W, (uint64_t *) b->kid_local.p);
            ctx->count_launch(SG_T_IDS, 1);
        }
        ctx->lap("unpermute");
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        b->have_kid_local = true;
        ctx->t_end(SG_T_IDS);
        ctx->laps_print("ids");
    } else {
    RS(b->tuples, (n + 1) * 32);
    if (n) {
        { const int rc_ = sg::ensure_adopted_soa(b); if (rc_) return rc_; }
        pair_pack_base_kernel<<<nblk(n, 256), 256, 0, st>>>((const uint64_t *) b->aocc.p, (const uint64_t *) b->kid.p, n,
                (const uint64_t *) c->uniq_dev.p, c->rank, (uint64_t *) b->tuples.p);
        ctx->count_launch(SG_T_GROUP, 1);
    }
    RS(c->back, (b->n_syncmers + 1) * 16);
    rc = all_to_all_v(c, (const uint64_t *) b->tuples.p, c->recv_counts, (uint64_t *) c->back.p, c->send_counts, 2);
    if (rc) return rc;
    CK(cudaMemcpyAsync(c->mat_host + (size_t) W * W, c->uniq_dev.p, (size_t) W * 8, cudaMemcpyDeviceToHost, st));
    rc = sg_ids_scatter(b, c->back.p, b->n_syncmers);       // synchronises
    ctx->t_end(SG_T_IDS);
    if (rc) return rc;
    }
    c->uniq.assign(c->mat_host + (size_t) W * W, c->mat_host + (size_t) W * W + W);
    uint64_t base = 0, tot = 0;
    for (int r = 0; r < W; ++r) { if (r < c->rank) base += c->uniq[r]; tot += c->uniq[r]; }
    if (tot >= (1ull << 31)) { ctx->err = "more than 2^31 - 1 distinct k-mers: ids no longer fit the 64-bit arc keys"; return SG_E_LIMIT; }
    if (id_base) *id_base = base;
    if (n_unique_total) *n_unique_total = tot;
    return SG_OK;
}

int sg_comm_global_stat(sg_comm *c, sg_batch *b, sg_stat_t *st_io)
{
    if (!c || !b || !st_io) return SG_E_ARG;
    sg_ctx *ctx = c->ctx;
    NcclApi *N = nccl_api();
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    const int W = c->world;
    // s-mer codes: (code, local count) pairs of every rank into one table
    void *pairs = nullptr;
    uint64_t n = 0;
    int rc = sg_smer_counts_pack(b, &pairs, &n);
    if (rc) return rc;
    RS(c->cnt_dev, (size_t) W * 8);
    std::vector<uint64_t> same(W, n);
    CK(cudaMemcpyAsync(c->cnt_dev.p, same.data(), (size_t) W * 8, cudaMemcpyHostToDevice, st));
    rc = gather_count_matrix(c, (const uint64_t *) c->cnt_dev.p);   // synchronises: `same` may go
    if (rc) return rc;
    std::vector<uint64_t> sc(W, n), rcnt(W, 0);
    uint64_t total = 0;
    for (int p = 0; p < W; ++p) { rcnt[p] = c->mat_host[(size_t) p * W + c->rank]; total += rcnt[p]; }
    RS(c->gather, (total + 1) * 16);
    {   // all-gather-v: the same block goes to every rank
        uint64_t ro = 0;
        NK(N->GroupStart());
        for (int p = 0; p < W; ++p) {
            if (n) NK(N->Send(pairs, n * 2, ncclUint64, p, c->comm, st));
            if (rcnt[p]) NK(N->Recv((uint64_t *) c->gather.p + ro * 2, rcnt[p] * 2, ncclUint64, p, c->comm, st));
            ro += rcnt[p];
        }
        NK(N->GroupEnd());
    }
    rc = sg_smer_counts_merge(b, c->gather.p, total, st_io);
    if (rc) return rc;
    // k-mer tables, gap sums and totals add up: every k-mer lives on one rank, every read on one rank
    RS(c->scratch, 1008 * 8);
    int64_t *h = (int64_t *) (c->mat_host + (size_t) W * W + 2 * W);
    for (int i = 0; i < 1001; ++i) h[i] = st_io->kmer_cnts[i];
    h[1001] = st_io->gap_sum; h[1002] = (int64_t) st_io->n_gaps; h[1003] = (int64_t) st_io->kmer_unique;
    h[1004] = (int64_t) st_io->kmer_singleton; h[1005] = (int64_t) st_io->n_syncmers;
    CK(cudaMemcpyAsync(c->scratch.p, h, 1006 * 8, cudaMemcpyHostToDevice, st));
    NK(N->AllReduce(c->scratch.p, c->scratch.p, 1006, ncclInt64, ncclSum, c->comm, st));
    CK(cudaMemcpyAsync(h, c->scratch.p, 1006 * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int i = 0; i < 1001; ++i) st_io->kmer_cnts[i] = h[i];
    st_io->gap_sum = h[1001]; st_io->n_gaps = (uint64_t) h[1002]; st_io->kmer_unique = (uint64_t) h[1003];
    st_io->kmer_singleton = (uint64_t) h[1004]; st_io->n_syncmers = (uint64_t) h[1005];
    return SG_OK;
}

int sg_comm_arcs(sg_comm *c, sg_batch *b, uint32_t min_k_cov, double min_a_cov_f, int root, uint64_t *n_arcs)
{
    if (!c || !b || !n_arcs || root < 0 || root >= c->world) return SG_E_ARG;
    if (!b->counted || !b->have_kid_local || (int) c->uniq.size() != c->world) return SG_E_STATE;   // after sg_comm_return_ids
    sg_ctx *ctx = c->ctx;
    NcclApi *N = nccl_api();
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    const int W = c->world;
    *n_arcs = 0;
    ctx->t_begin(SG_T_ARCS);
    // 1. neighbouring pairs of the local reads, global ids
    uint64_t nslots = 0;
    int rc = arcs_tally_local(b, (const uint64_t *) b->kid_local.p, &nslots);
    if (rc) return rc;
    // 2. entries grouped by owner
    RS(c->pcount, 257 * 8); RS(c->pcursor, 257 * 8);
    CK(cudaMemsetAsync(c->pcount.p, 0, 257 * 8, st));
    arc_part_count_kernel<<<std::min<unsigned>(nblk(nslots, 256), 148u * 8u), 256, 0, st>>>((const uint64_t *) b->arc_keys.p, nslots, (uint32_t) W,
            (unsigned long long *) c->pcount.p);
    offsets_kernel<<<1, 32, 0, st>>>((const unsigned long long *) c->pcount.p, W, (unsigned long long *) c->pcursor.p);
    RS(c->ppairs, (nslots / 2 + 2) * 16);                     // load factor <= 0.5
    arc_part_scatter_kernel<<<std::min<unsigned>(nblk(nslots, 256), 148u * 8u), 256, 0, st>>>((const uint64_t *) b->arc_keys.p,
            (const uint32_t *) b->arc_vals.p, nslots, (uint32_t) W, (unsigned long long *) c->pcursor.p, (uint64_t *) c->ppairs.p);
    ctx->count_launch(SG_T_ARCS, 3);
    // 3. one exchange
    rc = gather_count_matrix(c, (const uint64_t *) c->pcount.p);
    if (rc) return rc;
    std::vector<uint64_t> sc(W), rcnt(W);
    uint64_t total = 0;
    for (int p = 0; p < W; ++p) { sc[p] = c->mat_host[(size_t) c->rank * W + p]; rcnt[p] = c->mat_host[(size_t) p * W + c->rank]; total += rcnt[p]; }
    RS(c->gather, (total + 1) * 16);
    rc = all_to_all_v(c, (const uint64_t *) c->ppairs.p, sc, (uint64_t *) c->gather.p, rcnt, 2);
    if (rc) return rc;
    // 4. the owner adds the counts up
    rc = arcs_merge_pairs(b, (const uint64_t *) c->gather.p, total, &nslots);
    if (rc) return rc;
    // 5. coverages of all hash ranges, in id order
    uint64_t U = 0;
    for (int r = 0; r < W; ++r) U += c->uniq[r];
    RS(c->covs, (U + 1) * 4);
    {
        uint64_t o = 0;
        NK(N->GroupStart());
        for (int r = 0; r < W; ++r) {
            if (c->uniq[r]) NK(N->Broadcast(b->scm_cov.p, (uint32_t *) c->covs.p + o, c->uniq[r], ncclUint32, r, c->comm, st));
            o += c->uniq[r];
        }
        NK(N->GroupEnd());
        c->bytes_sent += c->uniq[c->rank] * 4 * (uint64_t) (W - 1);
    }
    // 6. filter, complement arcs, local order
    uint64_t na = 0;
    rc = arcs_emit(b, nslots, (const uint32_t *) c->covs.p, min_k_cov, min_a_cov_f, &na);   // synchronises
    if (rc) return rc;
    // 7. every rank's arcs to the root, which puts them in (v, w, comp) order
    RS(c->cnt_dev, (size_t) W * 8);
    std::vector<uint64_t> to_root(W, 0);
    to_root[root] = na;
    CK(cudaMemcpyAsync(c->cnt_dev.p, to_root.data(), (size_t) W * 8, cudaMemcpyHostToDevice, st));
    rc = gather_count_matrix(c, (const uint64_t *) c->cnt_dev.p);
    if (rc) return rc;
    std::vector<uint64_t> rfrom(W, 0);
    uint64_t all = 0;
    for (int p = 0; p < W; ++p) { rfrom[p] = c->mat_host[(size_t) p * W + c->rank]; all += rfrom[p]; }
    // keys and values travel as two all-to-all-v's of one word per arc
    RS(c->gather, (all + 1) * 16);
    uint64_t *gk = (uint64_t *) c->gather.p, *gv = gk + all;
    rc = all_to_all_v(c, (const uint64_t *) b->arc_okey.p, to_root, gk, rfrom, 1);
    if (rc) return rc;
    rc = all_to_all_v(c, (const uint64_t *) b->arc_oval.p, to_root, gv, rfrom, 1);
    if (rc) return rc;
    if (c->rank == root) {
        RS(b->arc_okey, (all + 2) * 8); RS(b->arc_oval, (all + 2) * 8); RS(b->arc_okey_alt, (all + 2) * 8); RS(b->arc_oval_alt, (all + 2) * 8);
        if (all) {
            CK(cudaMemcpyAsync(b->arc_okey.p, gk, all * 8, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(b->arc_oval.p, gv, all * 8, cudaMemcpyDeviceToDevice, st));
        }
        rc = arcs_sort_unpack(b, all);
        if (rc) return rc;
        b->n_arcs = all;
        *n_arcs = all;
    } else b->n_arcs = 0;
    ctx->t_end(SG_T_ARCS);
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return SG_OK;
}

} // extern "C"
