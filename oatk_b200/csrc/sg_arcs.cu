// sg_arcs.cu -- kernel group 3: the arc tally of make_syncmer_graph (a7), and the
// hash-range partition of syncmer tuples used by the multi-GPU exchange.
//
// Reference (syncasm.c:236-282): every pair of neighbouring syncmers on a read
// gives an oriented pair (v0, v1), v = id << 1 | rev; it is counted under its
// canonical form ((v0,v1) if v0 <= v1, else (v1^1, v0^1)) in a khashl map keyed by
// the 128-bit pair; an entry becomes an arc when its count reaches
// min_a_cov_f * min(cov(v0), cov(v1)) and both syncmers pass the coverage filter,
// and every arc that is not its own complement also yields the complement arc.
//
// Here the map is a warp-cooperative open-addressing table: ids are < 2^31, so the
// pair fits one 64-bit key; a warp first merges equal keys among its 32 lanes with
// __match_any_sync, then takes the distinct keys one at a time and probes 32
// consecutive slots at once (one coalesced 256-byte read, __ballot_sync to find the
// key or the first empty slot, one atomicCAS by the elected lane). Arcs come out in
// table order and are sorted by (v, w, comp), the order asmg_finalize needs anyway.
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include "sg_common.cuh"
#include "sg_internal.h"
#include "sg_host.h"
#include "sg_table.cuh"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return SG_E_CUDA; } } while (0)
#define RS(buf, bytes) do { if ((buf).reserve(bytes)) { ctx->err = "device allocation of " + std::to_string((size_t)(bytes)) + " bytes failed"; return SG_E_NOMEM; } } while (0)
#define LAUNCHED(stage, expr) do { int n_ = (expr); if (n_ < 0) return n_; ctx->count_launch(stage, n_); } while (0)
static inline unsigned nblk(uint64_t n, unsigned t) { return (unsigned) ((n + t - 1) / t); }

namespace sg {

struct ArcTallyArgs {
    const uint64_t *occ, *kid;
    const uint32_t *m_pos;
    uint64_t n;
    uint64_t *tk; uint32_t *tv; uint64_t nslot_mask;
};

__global__ void __launch_bounds__(256) arc_tally_kernel(ArcTallyArgs A)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarp = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    for (uint64_t base = warp0 * 32; base < A.n; base += nwarp * 32) {
        const uint64_t o = base + lane;
        uint64_t key = EMPTY_KEY;
        if (o < A.n && o > 0) {
            const uint32_t idx = (uint32_t) (A.occ[o] >> 1) & 0x7FFFFFFFu;
            if (idx > 0) {                                 // same read as o - 1
                const uint64_t v0 = (A.kid[o - 1] >> 1) << 1 | (A.m_pos[o - 1] & 1u);
                const uint64_t v1 = (A.kid[o] >> 1) << 1 | (A.m_pos[o] & 1u);
                key = v0 <= v1 ? (v0 << 32 | v1) : ((v1 ^ 1) << 32 | (v0 ^ 1));      // syncasm.c:256-257
            }
        }
        table_add_warp(A.tk, A.tv, A.nslot_mask, key, lane);
    }
}

struct ArcEmitArgs {
    const uint64_t *tk; const uint32_t *tv; uint64_t nslots;
    const uint32_t *cov;
    uint32_t min_k_cov; double min_a_cov_f;
    unsigned long long *n_out;
    uint64_t cap;
    uint64_t *okey, *oval;          // key = v << 32 | w ; val = cov << 1 | comp
};

__global__ void __launch_bounds__(256) arc_emit_kernel(ArcEmitArgs A)
{
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < A.nslots; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t key = A.tk[i];
        if (key == EMPTY_KEY) continue;
        const uint64_t v0 = key >> 32, v1 = key & 0xFFFFFFFFull;
        const uint32_t c = A.tv[i], c0 = A.cov[v0 >> 1], c1 = A.cov[v1 >> 1];
        if ((double) c < A.min_a_cov_f * (double) min(c0, c1) || c0 < A.min_k_cov || c1 < A.min_k_cov) continue;   // syncasm.c:270-272
        const bool selfc = (v1 ^ 1) == v0;                 // the arc is its own complement, syncasm.c:275
        const unsigned long long o = atomicAdd(A.n_out, selfc ? 1ull : 2ull);
        if (o + (selfc ? 1 : 2) <= A.cap) {
            A.okey[o] = key; A.oval[o] = (uint64_t) c << 1;
            if (!selfc) { A.okey[o + 1] = (v1 ^ 1) << 32 | (v0 ^ 1); A.oval[o + 1] = (uint64_t) c << 1 | 1ull; }
        }
    }
}

__global__ void __launch_bounds__(256) arc_unpack_kernel(const uint64_t *key, const uint64_t *val, uint64_t n, uint64_t *out4)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out4[4 * i + 0] = key[i] >> 32; out4[4 * i + 1] = key[i] & 0xFFFFFFFFull;
    out4[4 * i + 2] = val[i] >> 1; out4[4 * i + 3] = val[i] & 1ull;
}

// ---- hash-range partition for the multi-GPU exchange ----
__global__ void __launch_bounds__(256) part_key_kernel(const uint64_t *key, uint64_t n, uint32_t n_parts, uint64_t *pkey, uint64_t *pval)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // range partition on key >> 1 so that the two hashes sr_db_stat merges (syncmer.c:896) stay together
    pkey[i] = __umul64hi((key[i] >> 1) << 1, (uint64_t) n_parts);
    pval[i] = i;
}

__global__ void __launch_bounds__(256) part_gather_kernel(const uint64_t *pkey, const uint64_t *pval, const uint64_t *key, const uint64_t *occ,
        const uint64_t *smer, const uint64_t *fp, uint64_t n, uint64_t *tuples, unsigned long long *counts)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t o = pval[i];
    tuples[4 * i + 0] = key[o]; tuples[4 * i + 1] = occ[o]; tuples[4 * i + 2] = smer[o]; tuples[4 * i + 3] = fp[o];
    if (i == n - 1 || pkey[i + 1] != pkey[i]) atomicMax(counts + pkey[i], (unsigned long long) (i + 1));   // end offset of this part
}

__global__ void __launch_bounds__(256) adopt_kernel(const uint64_t *tuples, uint64_t n, uint64_t *key, uint64_t *occ, uint64_t *smer, uint64_t *fp, ulonglong4 *rec)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t h = tuples[4 * i], o = tuples[4 * i + 1], sm = tuples[4 * i + 2], f = tuples[4 * i + 3];   // (hash, occurrence, s-mer code, fingerprint)
    key[i] = h; occ[i] = o; smer[i] = sm; fp[i] = f;
    rec[i] = make_ulonglong4(o, sm, f, h);                             // the record layout of kmerhash_kernel: the packed sort of sg_count gathers from it
}

// ids coming back from the GPU that owns the hash range: pairs (occ, id << 1) for occurrences on local reads
__global__ void __launch_bounds__(256) kid_scatter_kernel(const uint64_t *pairs, uint64_t n, const uint64_t *scm_off, uint64_t sid_base,
        uint64_t n_reads, uint64_t id_add, uint64_t *kid, unsigned long long *bad)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t oc = pairs[2 * i], sid = (oc >> 32) - sid_base, idx = (oc >> 1) & 0x7FFFFFFFull;
    if (sid >= n_reads || scm_off[sid] + idx >= scm_off[sid + 1]) { atomicAdd(bad, 1ull); return; }
    kid[scm_off[sid] + idx] = pairs[2 * i + 1] + (id_add << 1);
}

__global__ void __launch_bounds__(256) pair_pack_kernel(const uint64_t *occ, const uint64_t *kid, uint64_t n, uint64_t id_add, uint64_t *pairs)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pairs[2 * i] = occ[i]; pairs[2 * i + 1] = kid[i] + (id_add << 1);
}

} // namespace sg

using namespace sg;


namespace sg {

// (key, count) pairs added into a tally table (the owner's side of the multi-GPU arc exchange)
__global__ void __launch_bounds__(256) arc_pair_tally_kernel(const uint64_t *pairs, uint64_t n, uint64_t *tk, uint32_t *tv, uint64_t nslot_mask)
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

static int arcs_table(sg_batch *b, uint64_t items, uint64_t *nslots_out)
{
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    uint64_t nslots = 1024;
    while (nslots < 2 * items) nslots <<= 1;      // load factor <= 0.5 even if every item is a distinct key
    RS(b->arc_keys, nslots * 8); RS(b->arc_vals, nslots * 4);
    RS(b->status, 4 * 8);
    CK(cudaMemsetAsync(b->arc_keys.p, 0xff, nslots * 8, st));
    CK(cudaMemsetAsync(b->arc_vals.p, 0, nslots * 4, st));
    b->smer_slots = 0;                            // sg_stat's s-mer table lived in the same buffers
    *nslots_out = nslots;
    return SG_OK;
}

// neighbouring pairs of this batch's reads -> table; kid = id << 1 per syncmer in read order (ids < 2^31)
int arcs_tally_local(sg_batch *b, const uint64_t *kid, uint64_t *nslots_out)
{
    sg_ctx *ctx = b->ctx;
    const uint64_t N = b->n_syncmers;
    int rc = arcs_table(b, N, nslots_out);
    if (rc) return rc;
    ArcTallyArgs T;
    T.occ = (const uint64_t *) b->occ.p; T.kid = kid; T.m_pos = (const uint32_t *) b->m_pos.p; T.n = N;
    T.tk = (uint64_t *) b->arc_keys.p; T.tv = (uint32_t *) b->arc_vals.p; T.nslot_mask = *nslots_out - 1;
    if (N) arc_tally_kernel<<<std::min<unsigned>(nblk(N, 256), 148u * 16u), 256, 0, ctx->stream>>>(T);
    ctx->count_launch(SG_T_ARCS, 1);
    return SG_OK;
}

int arcs_merge_pairs(sg_batch *b, const uint64_t *pairs, uint64_t n, uint64_t *nslots_out)
{
    sg_ctx *ctx = b->ctx;
    int rc = arcs_table(b, n, nslots_out);
    if (rc) return rc;
    if (n) arc_pair_tally_kernel<<<std::min<unsigned>(nblk(n, 256), 148u * 16u), 256, 0, ctx->stream>>>(pairs, n, (uint64_t *) b->arc_keys.p,
            (uint32_t *) b->arc_vals.p, *nslots_out - 1);
    ctx->count_launch(SG_T_ARCS, 1);
    return SG_OK;
}

// table -> arcs that pass the filter, with their complements, in table order (arc_okey / arc_oval). Synchronises.
int arcs_emit(sg_batch *b, uint64_t nslots, const uint32_t *cov, uint32_t min_k_cov, double min_a_cov_f, uint64_t *n_arcs)
{
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    unsigned long long *d_n = (unsigned long long *) b->status.p + 3;
    uint64_t cap = std::max<uint64_t>(b->arc_cap, 1 << 16);
    uint64_t na = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        RS(b->arc_okey, (cap + 2) * 8); RS(b->arc_oval, (cap + 2) * 8);
        RS(b->arc_okey_alt, (cap + 2) * 8); RS(b->arc_oval_alt, (cap + 2) * 8);
        CK(cudaMemsetAsync(d_n, 0, 8, st));
        ArcEmitArgs E;
        E.tk = (const uint64_t *) b->arc_keys.p; E.tv = (const uint32_t *) b->arc_vals.p; E.nslots = nslots; E.cov = cov;
        E.min_k_cov = min_k_cov; E.min_a_cov_f = min_a_cov_f; E.n_out = d_n; E.cap = cap;
        E.okey = (uint64_t *) b->arc_okey.p; E.oval = (uint64_t *) b->arc_oval.p;
        arc_emit_kernel<<<std::min<unsigned>(nblk(nslots, 256), 148u * 16u), 256, 0, st>>>(E);
        ctx->count_launch(SG_T_ARCS, 1);
        CK(cudaMemcpyAsync(&na, d_n, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (na <= cap) break;
        cap = na + 16;
    }
    b->arc_cap = cap;
    *n_arcs = na;
    return SG_OK;
}

// arc_okey / arc_oval (na entries) -> (v, w, comp) order -> arc_out as 4 words per arc
int arcs_sort_unpack(sg_batch *b, uint64_t na)
{
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    if (!na) return SG_OK;
    // order by (v, w, comp): stable LSD, minor key first
    RS(b->sort_tmp, sort_tmp_words(na) * 4);
    uint64_t *k0 = (uint64_t *) b->arc_okey.p, *v0 = (uint64_t *) b->arc_oval.p;
    uint64_t *k1 = (uint64_t *) b->arc_okey_alt.p, *v1 = (uint64_t *) b->arc_oval_alt.p;
    LAUNCHED(SG_T_ARCS, launch_sort_pairs(v0, k0, v1, k1, na, 0, 40, (uint32_t *) b->sort_tmp.p, st));   // by cov<<1|comp
    LAUNCHED(SG_T_ARCS, launch_sort_pairs(k0, v0, k1, v1, na, 0, 64, (uint32_t *) b->sort_tmp.p, st));   // by v<<32|w
    RS(b->arc_out, na * 32);
    arc_unpack_kernel<<<nblk(na, 256), 256, 0, st>>>(k0, v0, na, (uint64_t *) b->arc_out.p);
    ctx->count_launch(SG_T_ARCS, 1);
    return SG_OK;
}

// this batch's tuples grouped by hash range (b->tuples, 4 words each, (sid, idx) order kept inside a part);
// b->part_counts[p] = end offset of part p, 0 for an empty part. No host synchronisation.
// records (occ, s_mer, fp, hash) of the adopted set -> the four arrays the pair sort and sg_ids_pack read
__global__ void __launch_bounds__(256) unrecord_kernel(const ulonglong4 *rec, uint64_t n, uint64_t *key, uint64_t *occ, uint64_t *smer, uint64_t *fp)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ulonglong4 t = rec[i];
    occ[i] = t.x; smer[i] = t.y; fp[i] = t.z; key[i] = t.w;
}

int ensure_adopted_soa(sg_batch *b)
{
    if (!b->adopted || b->asoa_valid) return SG_OK;
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    const uint64_t n = b->n_adopted;
    RS(b->akey, (n + 1) * 8); RS(b->aocc, (n + 1) * 8); RS(b->asmer, (n + 1) * 8); RS(b->afp, (n + 1) * 8);
    if (n) {
        unrecord_kernel<<<nblk(n, 256), 256, 0, st>>>((const ulonglong4 *) b->tup.p, n, (uint64_t *) b->akey.p, (uint64_t *) b->aocc.p,
                (uint64_t *) b->asmer.p, (uint64_t *) b->afp.p);
        ctx->count_launch(SG_T_SORT, 1);
    }
    b->asoa_valid = true;
    return SG_OK;
}

int tuples_partition_device(sg_batch *b, int n_parts)
{
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    const uint64_t N = b->n_syncmers;
    RS(b->skey, (N + 1) * 8); RS(b->sval, (N + 1) * 8); RS(b->skey_alt, (N + 1) * 8); RS(b->sval_alt, (N + 1) * 8);
    RS(b->sort_tmp, sort_tmp_words(std::max<uint64_t>(N, 1)) * 4);
    RS(b->tuples, (N + 1) * 32);
    RS(b->status, 4 * 8);
    RS(b->part_counts, 257 * 8);
    CK(cudaMemsetAsync(b->part_counts.p, 0, 257 * 8, st));
    if (N) {
        part_key_kernel<<<nblk(N, 256), 256, 0, st>>>((const uint64_t *) b->key.p, N, (uint32_t) n_parts, (uint64_t *) b->skey.p, (uint64_t *) b->sval.p);
        ctx->count_launch(SG_T_SORT, 1);
        // one stable counting pass on the part index keeps the (sid, idx) order inside every part
        LAUNCHED(SG_T_SORT, launch_sort_pairs((uint64_t *) b->skey.p, (uint64_t *) b->sval.p, (uint64_t *) b->skey_alt.p,
                (uint64_t *) b->sval_alt.p, N, 0, 8, (uint32_t *) b->sort_tmp.p, st));
        part_gather_kernel<<<nblk(N, 256), 256, 0, st>>>((const uint64_t *) b->skey.p, (const uint64_t *) b->sval.p,
                (const uint64_t *) b->key.p, (const uint64_t *) b->occ.p, (const uint64_t *) b->s_mer.p, (const uint64_t *) b->fp.p, N,
                (uint64_t *) b->tuples.p, (unsigned long long *) b->part_counts.p);
        ctx->count_launch(SG_T_SORT, 1);
    }
    b->sorted = false;                                     // skey/sval were used as scratch
    return SG_OK;
}

} // namespace sg

extern "C" {

int sg_arcs(sg_batch *b, uint32_t min_k_cov, double min_a_cov_f, uint64_t *n_arcs)
{
    if (!b || !n_arcs) return SG_E_ARG;
    if (!b->counted) return SG_E_STATE;
    if (b->adopted) return SG_E_STATE;            // needs the per-read id sequences of the local reads: sg_comm_arcs
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    if (b->n_unique >= (1ull << 31)) { ctx->err = "2^31 or more distinct k-mers: vertex ids no longer fit the 64-bit arc keys"; return SG_E_LIMIT; }
    ctx->t_begin(SG_T_ARCS);
    uint64_t nslots = 0, na = 0;
    int rc = arcs_tally_local(b, (const uint64_t *) b->kid.p, &nslots);
    if (rc) return rc;
    rc = arcs_emit(b, nslots, (const uint32_t *) b->scm_cov.p, min_k_cov, min_a_cov_f, &na);
    if (rc) return rc;
    b->n_arcs = na;
    rc = arcs_sort_unpack(b, na);
    if (rc) return rc;
    ctx->t_end(SG_T_ARCS);
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    *n_arcs = na;
    return SG_OK;
}

int sg_arcs_download(sg_batch *b, uint64_t *arcs4)
{
    if (!b || !arcs4) return SG_E_ARG;
    sg_ctx *ctx = b->ctx;
    if (b->n_arcs == 0) return SG_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(arcs4, b->arc_out.p, b->n_arcs * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    b->d2h_bytes += b->n_arcs * 32;
    return SG_OK;
}

int sg_tuples_partition(sg_batch *b, int n_parts, uint64_t *counts, void **d_tuples)
{
    if (!b || !counts || !d_tuples || n_parts < 1 || n_parts > 256) return SG_E_ARG;
    if (!b->extracted) return SG_E_STATE;
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    int rc = tuples_partition_device(b, n_parts);
    if (rc) return rc;
    std::vector<unsigned long long> ends(256, 0);
    CK(cudaMemcpyAsync(ends.data(), b->part_counts.p, 256 * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    uint64_t prev = 0;
    for (int p = 0; p < n_parts; ++p) {
        const uint64_t e = ends[p] ? ends[p] : prev;       // empty part: no end was recorded
        counts[p] = e - prev;
        prev = e;
    }
    *d_tuples = b->tuples.p;
    return SG_OK;
}

int sg_tuples_adopt(sg_batch *b, const void *d_tuples, uint64_t n)
{
    if (!b || (!d_tuples && n)) return SG_E_ARG;
    if (!b->extracted) return SG_E_STATE;
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    // keep the local read-order arrays for the id write-back; the tuple set being counted is replaced
    RS(b->akey, (n + 1) * 8); RS(b->aocc, (n + 1) * 8); RS(b->asmer, (n + 1) * 8); RS(b->afp, (n + 1) * 8);
    RS(b->tup, (n + 1) * 32);                                      // the records of the local extract are not needed again before the next one
    b->tup_valid = false;
    if (n) {
        adopt_kernel<<<nblk(n, 256), 256, 0, st>>>((const uint64_t *) d_tuples, n, (uint64_t *) b->akey.p, (uint64_t *) b->aocc.p, (uint64_t *) b->asmer.p,
                (uint64_t *) b->afp.p, (ulonglong4 *) b->tup.p);
        ctx->count_launch(SG_T_SORT, 1);
    }
    CK(cudaGetLastError());
    b->atup_valid = true;
    b->asoa_valid = true;
    b->range_lo = 0; b->range_lsh = 0;                             // the caller's parts are its own business: the whole hash orders the sort
    b->adopted = true;
    b->n_adopted = n;
    b->sorted = b->counted = false;
    return SG_OK;
}

int sg_ids_pack(sg_batch *b, uint64_t id_base, void **d_pairs, uint64_t *n)
{
    if (!b || !d_pairs || !n) return SG_E_ARG;
    if (!b->counted || !b->adopted) return SG_E_STATE;
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    const uint64_t N = b->n_adopted;
    { const int rc_ = ensure_adopted_soa(b); if (rc_) return rc_; }
    RS(b->tuples, (N + 1) * 32);
    if (N) {
        pair_pack_kernel<<<nblk(N, 256), 256, 0, st>>>((const uint64_t *) b->aocc.p, (const uint64_t *) b->kid.p, N, id_base, (uint64_t *) b->tuples.p);
        ctx->count_launch(SG_T_GROUP, 1);
    }
    CK(cudaGetLastError());
    *d_pairs = b->tuples.p;
    *n = N;
    return SG_OK;
}

int sg_ids_scatter(sg_batch *b, const void *d_pairs, uint64_t n)
{
    if (!b || (!d_pairs && n)) return SG_E_ARG;
    if (!b->extracted) return SG_E_STATE;
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    if (n != b->n_syncmers) { ctx->err = "id pairs do not cover the local syncmers"; return SG_E_ARG; }
    RS(b->kid_local, (b->n_syncmers + 1) * 8);
    RS(b->status, 4 * 8);
    CK(cudaMemsetAsync(b->status.p, 0, 4 * 8, st));
    if (n) {
        kid_scatter_kernel<<<nblk(n, 256), 256, 0, st>>>((const uint64_t *) d_pairs, n, (const uint64_t *) b->scm_off.p, b->sid_base,
                b->n_reads, 0, (uint64_t *) b->kid_local.p, (unsigned long long *) b->status.p);
        ctx->count_launch(SG_T_GROUP, 1);
    }
    unsigned long long bad = 0;
    CK(cudaMemcpyAsync(&bad, b->status.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (bad) { ctx->err = "id pairs refer to reads of another rank"; return SG_E_ARG; }
    b->have_kid_local = true;
    return SG_OK;
}

} // extern "C"
