// sg_runlen.cu -- f1 on the device: the run-length sums behind scg_syncmer_consensus.
//
// The reference turns a syncmer back into bases by writing every hoco base 1 + lround(mean(run length - 1))
// times, the mean taken over all occurrences of the k-mer that read error correction left alone
// (reference syncasm.c:946-998; ho_rl holds run length - 1, 255 = "look in ho_l_rl", syncmer.h:56-61).
// That is bases x coverage byte additions over ho_rl, the one array of the read database that is as large as
// the input (1 byte per hoco base) -- and the only consumer of it. With this kernel ho_rl never leaves the
// device: the host sends, per requested syncmer, the list of its uncorrected occurrences
// (read << 32 | hoco start << 1 | strand) and gets back k sums in the syncmer's own orientation
//     F[j] = sum over strand-0 copies of rl[start + j]  +  sum over strand-1 copies of rl[start + k-1-j]
// from which every view the consensus asks for follows on the host (reverse: F[k-1-j]; a start offset: a
// suffix). Integer sums, so the result is exact whatever the order of the additions.
//
// One CTA per request; a thread owns positions j, j + 256, ...; the occurrence list is walked by all
// threads together, so the bytes of one copy are read by consecutive threads (coalesced, forwards or
// backwards). A byte of 255 is looked up in the side list of long runs, kept sorted by (read, hoco index).
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include "sg_common.cuh"
#include "sg_internal.h"
#include "sg_host.h"

namespace sg {

struct RunlenArgs {
    const uint64_t *hoff;          // capacity offset per local read
    const uint8_t *ho_rl;
    const uint32_t *hoco_l;
    uint64_t sid_base, n_reads;
    int k;
    const uint64_t *rq_off, *rq_occ;
    uint64_t n_req;
    const uint64_t *lrl_key;       // read << 32 | hoco index, ascending
    const uint64_t *lrl_val;       // run length - 1
    uint64_t n_lrl;
    uint64_t *out;                 // n_req x k
    unsigned long long *bad;       // occurrences that point outside their read
};

__device__ __forceinline__ uint64_t long_run(const RunlenArgs &A, uint64_t key)
{
    uint64_t lo = 0, hi = A.n_lrl;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (A.lrl_key[mid] < key) lo = mid + 1; else hi = mid; }
    return (lo < A.n_lrl && A.lrl_key[lo] == key) ? A.lrl_val[lo] : 255ull;
}

__global__ void __launch_bounds__(256) runlen_sum_kernel(RunlenArgs A)
{
    __shared__ uint64_t s_base[64];
    __shared__ uint32_t s_start[64], s_flag[64];
    for (uint64_t rq = blockIdx.x; rq < A.n_req; rq += gridDim.x) {
        const uint64_t o0 = A.rq_off[rq], o1 = A.rq_off[rq + 1];
        uint64_t acc[4] = {0, 0, 0, 0};                   // positions tid, tid + 256, ... (k <= 1024 in registers, beyond that in passes)
        for (int j0 = 0; j0 < A.k; j0 += 1024) {
            for (int q = 0; q < 4; ++q) acc[q] = 0;
            for (uint64_t ob = o0; ob < o1; ob += 64) {
                const int nb = (int) min((uint64_t) 64, o1 - ob);
                __syncthreads();
                if (threadIdx.x < nb) {
                    const uint64_t e = A.rq_occ[ob + threadIdx.x];
                    const uint64_t sid = (e >> 32) - A.sid_base;
                    const uint32_t start = (uint32_t) (e >> 1) & 0x7FFFFFFFu;
                    bool ok = sid < A.n_reads;
                    if (ok) ok = (uint64_t) start + (uint64_t) A.k <= (uint64_t) A.hoco_l[sid];
                    if (!ok) atomicAdd(A.bad, 1ull);
                    s_base[threadIdx.x] = ok ? A.hoff[sid] : 0;
                    s_start[threadIdx.x] = start;
                    s_flag[threadIdx.x] = (uint32_t) (e & 1u) | (ok ? 2u : 0u);
                }
                __syncthreads();
                for (int c = 0; c < nb; ++c) {
                    if (!(s_flag[c] & 2u)) continue;
                    const uint64_t base = s_base[c];
                    const uint32_t start = s_start[c];
                    const bool rev = s_flag[c] & 1u;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int j = j0 + q * 256 + threadIdx.x;
                        if (j < A.k) {
                            const uint32_t pos = rev ? start + (uint32_t) (A.k - 1 - j) : start + (uint32_t) j;
                            uint64_t rl = A.ho_rl[base + pos];
                            if (rl == 255 && A.n_lrl) {
                                const uint64_t e = A.rq_occ[ob + c];
                                rl = long_run(A, (e >> 32) << 32 | pos);
                            }
                            acc[q] += rl;
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j0 + q * 256 + threadIdx.x;
                if (j < A.k) A.out[rq * (uint64_t) A.k + j] = acc[q];
            }
        }
    }
}

// hoco bases of requested stretches, one code (0..3) per byte: what get_kmer_seq reads from sr_t.hoco_s when the packed
// bases stayed on the device. One warp per request.
__global__ void __launch_bounds__(256) kmer_codes_kernel(const uint64_t *hoff, const uint8_t *hoco_s, const uint32_t *hoco_l, uint64_t sid_base, uint64_t n_reads,
        const uint64_t *refs, uint64_t n_req, int len, uint8_t *out, unsigned long long *bad)
{
    const int lane = threadIdx.x & 31;
    const uint64_t w0 = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    for (uint64_t rq = w0; rq < n_req; rq += nw) {
        const uint64_t e = refs[rq];
        const uint64_t sid = (e >> 32) - sid_base;
        const uint32_t start = (uint32_t) e;
        const bool ok = sid < n_reads && (uint64_t) start + (uint64_t) len <= (uint64_t) hoco_l[sid < n_reads ? sid : 0];
        if (!ok) { if (lane == 0) atomicAdd(bad, 1ull); continue; }
        const uint8_t *hs = hoco_s + hoff[sid] / 4;
        for (int i = lane; i < len; i += 32) {
            const uint32_t p = start + (uint32_t) i;
            out[rq * (uint64_t) len + i] = (uint8_t) ((hs[p >> 2] >> ((3 - (p & 3)) << 1)) & 3);
        }
    }
}

__global__ void __launch_bounds__(256) lrl_key_kernel(const uint32_t *sid, const uint32_t *idx, const uint32_t *val, uint64_t n, uint64_t sid_base,
        uint64_t *key, uint64_t *v)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { key[i] = ((uint64_t) sid[i] + sid_base) << 32 | idx[i]; v[i] = val[i]; }
}

} // namespace sg

using namespace sg;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return SG_E_CUDA; } } while (0)
#define RS(buf, bytes) do { if ((buf).reserve(bytes)) { ctx->err = "device allocation of " + std::to_string((size_t)(bytes)) + " bytes failed"; return SG_E_NOMEM; } } while (0)
#define LAUNCHED(stage, expr) do { int n_ = (expr); if (n_ < 0) return n_; ctx->count_launch(stage, n_); } while (0)
static inline unsigned nblk(uint64_t n, unsigned t) { return (unsigned) ((n + t - 1) / t); }

extern "C" {

int sg_runlen_resident(sg_batch *b) { return b && b->extracted && b->rl_resident ? 1 : 0; }

int sg_kmer_codes(sg_batch *b, uint64_t n_req, const uint64_t *refs, int len, uint8_t *codes)
{
    if (!b || len < 1 || (n_req && (!refs || !codes))) return SG_E_ARG;
    if (!b->extracted) return SG_E_STATE;
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    if (n_req == 0) return SG_OK;
    RS(b->rq_occ, (n_req + 1) * 8); RS(b->rq_out, n_req * (uint64_t) len + 16); RS(b->status, 4 * 8);
    CK(cudaMemsetAsync(b->status.p, 0, 8, st));
    CK(cudaMemcpyAsync(b->rq_occ.p, refs, n_req * 8, cudaMemcpyHostToDevice, st));
    ctx->t_begin(SG_T_PACK);
    kmer_codes_kernel<<<(unsigned) std::min<uint64_t>((n_req + 7) / 8, 148ull * 16ull), 256, 0, st>>>((const uint64_t *) b->hoff.p, (const uint8_t *) b->hoco_s.p,
            (const uint32_t *) b->hoco_l.p, b->sid_base, b->n_reads, (const uint64_t *) b->rq_occ.p, n_req, len, (uint8_t *) b->rq_out.p, (unsigned long long *) b->status.p);
    ctx->count_launch(SG_T_PACK, 1);
    ctx->t_end(SG_T_PACK);
    unsigned long long bad = 0;
    CK(cudaMemcpyAsync(codes, b->rq_out.p, n_req * (uint64_t) len, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&bad, b->status.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    b->h2d_bytes += n_req * 8;
    b->d2h_bytes += n_req * (uint64_t) len;
    if (bad) { ctx->err = std::to_string(bad) + " requested stretches lie outside their read"; return SG_E_ARG; }
    return SG_OK;
}

int sg_runlen_sums(sg_batch *b, uint64_t n_req, const uint64_t *occ_off, const uint64_t *occ, uint64_t *sums)
{
    if (!b || (n_req && (!occ_off || !sums))) return SG_E_ARG;
    if (!b->extracted || !b->rl_resident) return SG_E_STATE;
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    if (n_req == 0) return SG_OK;
    const uint64_t n_occ = occ_off[n_req];
    if (n_occ && !occ) return SG_E_ARG;
    // the side list of long runs, ordered by (read, hoco index), once per extraction
    const uint64_t nl = b->n_lrl_total;
    if (nl && !b->lrl_sorted) {
        RS(b->lrl_key, (nl + 1) * 8); RS(b->lrl_sval, (nl + 1) * 8); RS(b->lrl_key_alt, (nl + 1) * 8); RS(b->lrl_val_alt, (nl + 1) * 8);
        RS(b->sort_tmp, sort_tmp_words(nl) * 4);
        lrl_key_kernel<<<nblk(nl, 256), 256, 0, st>>>((const uint32_t *) b->lrl_sid.p, (const uint32_t *) b->lrl_idx.p, (const uint32_t *) b->lrl_val.p,
                nl, b->sid_base, (uint64_t *) b->lrl_key.p, (uint64_t *) b->lrl_sval.p);
        ctx->count_launch(SG_T_PACK, 1);
        LAUNCHED(SG_T_PACK, launch_sort_pairs((uint64_t *) b->lrl_key.p, (uint64_t *) b->lrl_sval.p, (uint64_t *) b->lrl_key_alt.p,
                (uint64_t *) b->lrl_val_alt.p, nl, 0, 64, (uint32_t *) b->sort_tmp.p, st));
        b->lrl_sorted = true;
    }
    RS(b->rq_off, (n_req + 1) * 8); RS(b->rq_occ, (n_occ + 1) * 8); RS(b->rq_out, n_req * (uint64_t) b->k * 8);
    RS(b->status, 4 * 8);
    CK(cudaMemsetAsync(b->status.p, 0, 8, st));
    CK(cudaMemcpyAsync(b->rq_off.p, occ_off, (n_req + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_occ) CK(cudaMemcpyAsync(b->rq_occ.p, occ, n_occ * 8, cudaMemcpyHostToDevice, st));
    b->h2d_bytes += (n_req + 1 + n_occ) * 8;
    RunlenArgs A;
    A.hoff = (const uint64_t *) b->hoff.p; A.ho_rl = (const uint8_t *) b->ho_rl.p; A.hoco_l = (const uint32_t *) b->hoco_l.p;
    A.sid_base = b->sid_base; A.n_reads = b->n_reads; A.k = b->k;
    A.rq_off = (const uint64_t *) b->rq_off.p; A.rq_occ = (const uint64_t *) b->rq_occ.p; A.n_req = n_req;
    A.lrl_key = (const uint64_t *) b->lrl_key.p; A.lrl_val = (const uint64_t *) b->lrl_sval.p; A.n_lrl = nl;
    A.out = (uint64_t *) b->rq_out.p; A.bad = (unsigned long long *) b->status.p;
    ctx->t_begin(SG_T_PACK);
    runlen_sum_kernel<<<(unsigned) std::min<uint64_t>(n_req, 148ull * 8ull), 256, 0, st>>>(A);
    ctx->count_launch(SG_T_PACK, 1);
    ctx->t_end(SG_T_PACK);
    unsigned long long bad = 0;
    CK(cudaMemcpyAsync(sums, b->rq_out.p, n_req * (uint64_t) b->k * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&bad, b->status.p, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    b->d2h_bytes += n_req * (uint64_t) b->k * 8;
    if (bad) { ctx->err = std::to_string(bad) + " occurrences point outside their read"; return SG_E_ARG; }
    return SG_OK;
}

} // extern "C"
