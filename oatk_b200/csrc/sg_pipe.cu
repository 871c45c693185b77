// sg_pipe.cu -- the host-buffer path at full PCIe rate: reads are cut into chunks that flow
// through n_slots independent (context, stream, batch) triples, one host thread each, so the
// host-to-device copy of chunk i+1, the kernels of chunk i and the device-to-host copy of
// chunk i-1 overlap. What the reference does in sr_read (syncmer.c:487-556: a batch of
// 10 000 reads per pthread, joined per super-batch) becomes a 3-deep copy/compute/copy
// pipeline; per-read results land in the caller's host arrays in read order, and the small
// per-syncmer tuples (plus the 2-bit packed reads, which the exact k-mer verification of
// sg_count needs) are appended to a device-resident master batch on which sg_stat,
// sg_count and sg_arcs then run as usual.
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include "sg_common.cuh"
#include "sg_internal.h"
#include "sg_host.h"

namespace sg {

__global__ void __launch_bounds__(256) add_offset_kernel(const uint64_t *src, uint64_t *dst, uint64_t n, uint64_t add)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] + add;
}
__global__ void __launch_bounds__(256) add_offset32_kernel(const uint32_t *src, uint32_t *dst, uint64_t n, uint32_t add)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] + add;
}

} // namespace sg

using namespace sg;

// per slot: two device input buffers, so the upload of the slot's next chunk (on its own stream) runs
// while the current chunk is computed and downloaded
struct SlotIn {
    cudaStream_t up = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr};
    DevBuf bases[2], off[2];
    uint64_t *h_off[2] = {nullptr, nullptr};       // pinned chunk-local offsets
    size_t h_off_cap = 0;
    char *h_bases[2] = {nullptr, nullptr};         // pinned copy of a chunk when the caller's reads are pageable
    size_t h_bases_cap[2] = {0, 0};
    // pinned result staging of the callback form (one chunk)
    void *stage[16] = {};
    size_t stage_cap[16] = {};
    void *staged(int i, size_t bytes)
    {
        if (stage_cap[i] < bytes) {
            if (stage[i]) cudaFreeHost(stage[i]);
            stage[i] = nullptr; stage_cap[i] = 0;
            const size_t want = bytes + bytes / 4 + 4096;
            if (cudaMallocHost(&stage[i], want) != cudaSuccess) return nullptr;
            stage_cap[i] = want;
        }
        return stage[i];
    }
};

struct sg_pipe {
    int device = 0, n_slots = 0;
    std::vector<sg_ctx *> ctx;
    std::vector<sg_batch *> slot;
    std::vector<SlotIn *> in;
    sg_ctx *mctx = nullptr;
    sg_batch *master = nullptr;
    std::string err;
    uint64_t sid_base = 0;       // global index of the first read (multi-GPU: this GPU's block of the read set)
    unsigned cap_factor = 16;    // callback form: room for this many times the expected number of syncmers in the master batch
    bool overflowed = false;     // the last run ended because that room was not enough (sg_pipe_syncmer_overflow)
    bool keep_hs = false;        // hoco_s is not downloaded either (it is in the master batch anyway): sg_kmer_codes serves the consensus
    bool keep_rl = false;        // ho_rl stays on the device (master batch) instead of travelling to the host: sg_runlen_sums serves the consensus
};

static inline unsigned nblk(uint64_t n, unsigned t) { return (unsigned) ((n + t - 1) / t); }

extern "C" {

int sg_pipe_create(int device, int n_slots, sg_pipe **out)
{
    if (!out || n_slots < 1 || n_slots > 8) return SG_E_ARG;
    *out = nullptr;
    sg_pipe *p = new sg_pipe();
    p->device = device; p->n_slots = n_slots;
    int rc = sg_ctx_create(device, &p->mctx);
    if (rc) { delete p; return rc; }
    if ((rc = sg_batch_create(p->mctx, &p->master))) { sg_ctx_destroy(p->mctx); delete p; return rc; }
    for (int i = 0; i < n_slots; ++i) {
        sg_ctx *c = nullptr; sg_batch *b = nullptr;
        cudaStream_t st;
        if (sg_ctx_create(device, &c) || cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess || sg_batch_create(c, &b)) {
            rc = SG_E_CUDA;
            break;
        }
        sg_ctx_set_stream(c, st);
        p->ctx.push_back(c); p->slot.push_back(b);
        SlotIn *si = new SlotIn();
        p->in.push_back(si);
        if (cudaStreamCreateWithFlags(&si->up, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&si->ready[0], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&si->ready[1], cudaEventDisableTiming) != cudaSuccess) { rc = SG_E_CUDA; break; }
    }
    if (rc) { sg_pipe_destroy(p); return rc; }
    *out = p;
    return SG_OK;
}

void sg_pipe_destroy(sg_pipe *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    for (size_t i = 0; i < p->slot.size(); ++i) {
        sg_batch_destroy(p->slot[i]);
        cudaStream_t st = p->ctx[i]->stream;
        sg_ctx_destroy(p->ctx[i]);
        if (st) cudaStreamDestroy(st);
    }
    for (SlotIn *si : p->in) {
        if (si->up) cudaStreamDestroy(si->up);
        for (int j = 0; j < 2; ++j) {
            if (si->ready[j]) cudaEventDestroy(si->ready[j]);
            if (si->h_off[j]) cudaFreeHost(si->h_off[j]);
            if (si->h_bases[j]) cudaFreeHost(si->h_bases[j]);
        }
        for (int j = 0; j < 16; ++j) if (si->stage[j]) cudaFreeHost(si->stage[j]);
        delete si;
    }
    if (p->master) sg_batch_destroy(p->master);
    if (p->mctx) sg_ctx_destroy(p->mctx);
    delete p;
}

sg_batch *sg_pipe_master(sg_pipe *p) { return p ? p->master : nullptr; }
sg_ctx *sg_pipe_ctx(sg_pipe *p) { return p ? p->mctx : nullptr; }
const char *sg_pipe_last_error(sg_pipe *p) { return p ? p->err.c_str() : "no pipe"; }

int sg_pipe_set_capacity_factor(sg_pipe *p, unsigned factor)
{
    if (!p || factor < 1) return SG_E_ARG;
    p->cap_factor = factor;
    return SG_OK;
}
int sg_pipe_syncmer_overflow(sg_pipe *p) { return p && p->overflowed ? 1 : 0; }

int sg_pipe_keep_run_lengths(sg_pipe *p, int on) { if (!p) return SG_E_ARG; p->keep_rl = on != 0; return SG_OK; }
int sg_pipe_keep_packed_bases(sg_pipe *p, int on) { if (!p) return SG_E_ARG; p->keep_hs = on != 0; return SG_OK; }

int sg_pipe_set_sid_base(sg_pipe *p, uint64_t sid_base) { if (!p) return SG_E_ARG; p->sid_base = sid_base; return SG_OK; }

uint64_t sg_pipe_launches(sg_pipe *p)
{
    if (!p) return 0;
    uint64_t n = p->mctx->launches;
    for (auto c : p->ctx) n += c->launches;
    return n;
}

static int pipe_run(sg_pipe *p, const char *bases, const uint64_t *off, uint64_t n_reads, int k, int s,
        uint64_t chunk_reads, const sg_extract_out_t *out, const sg_pipe_caps_t *caps, sg_pipe_chunk_fn cb, void *cb_user, sg_extract_sizes_t *sizes)
{
    if (!p || !off || (!cb && (!out || !caps)) || !sizes || chunk_reads == 0) return SG_E_ARG;
    if (!(s > 0 && s < 32 && k > s)) return SG_E_ARG;
    if (n_reads > 0xFFFFFFFFull) return SG_E_LIMIT;
    cudaSetDevice(p->device);
    const uint64_t n_chunks = (n_reads + chunk_reads - 1) / chunk_reads;
    const uint64_t total = n_reads ? off[n_reads] - off[0] : 0;
    sg_batch *M = p->master;
    sg_ctx *mctx = p->mctx;
    auto fail = [&](int rc, const std::string &what) { p->err = what; return rc; };

    // pageable reads are staged through pinned per-slot buffers by the slot threads (the driver would do
    // the same, serially and synchronously)
    bool in_pinned = true;
    if (total) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, bases) != cudaSuccess) { cudaGetLastError(); in_pinned = false; }
        else in_pinned = pa.type == cudaMemoryTypeHost || pa.type == cudaMemoryTypeManaged;
    }
    // master arrays: per read and per position capacities are known, per syncmer ones come from the caller
    // (callback form: 2 closed syncmers per window is the expectation; room for 16, never more than one per base)
    const uint64_t qwin = (uint64_t) (k - s + 1);
    const uint64_t cap_pos = total + 64 * n_reads + 64;
    const uint64_t capN = cb ? std::min<uint64_t>(total + n_reads, (uint64_t) p->cap_factor * (total / qwin + n_reads)) + 1024 : caps->max_syncmers;
    p->overflowed = false;
    // run lengths stay on the device when asked to (sg_pipe_keep_run_lengths) or when the caller gave no buffer for them
    const bool keep_rl = p->keep_rl || (!cb && out && !out->ho_rl_buf);
    const uint64_t cap_lrl = total / 256 + 1024;             // a listed run is at least 256 bases long
    if (keep_rl && (M->ho_rl.reserve(cap_pos + 64) || M->lrl_sid.reserve(cap_lrl * 4) || M->lrl_idx.reserve(cap_lrl * 4) || M->lrl_val.reserve(cap_lrl * 4)))
        return fail(SG_E_NOMEM, "master allocation failed");
    if (M->hoff.reserve((n_reads + 1) * 8) || M->hoco_s.reserve(cap_pos / 4 + 64) || M->hoco_l.reserve((n_reads + 1) * 4) ||
            M->n_scm.reserve((n_reads + 1) * 4) || M->scm_off.reserve((n_reads + 1) * 8) || M->n_amb.reserve((n_reads + 1) * 4) ||
            M->key.reserve((capN + 1) * 8) || M->occ.reserve((capN + 1) * 8) || M->m_pos.reserve((capN + 1) * 4) || M->s_mer.reserve((capN + 1) * 8) || M->fp.reserve((capN + 1) * 8))
        return fail(SG_E_NOMEM, "master allocation failed");

    // running totals, advanced in chunk order
    struct Totals { uint64_t hoff = 0, hs = 0, rl = 0, scm = 0, amb = 0, lrl = 0, hoco = 0; } tot;
    std::mutex mu;
    std::condition_variable cv;
    uint64_t next_chunk = 0;
    int first_err = 0;
    std::string first_msg;

    const bool trace = getenv("SG_PIPE_TRACE") != nullptr;
    std::vector<double> tphase(p->n_slots * 6, 0.0);
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const bool last_writes_end = true;
    (void) last_writes_end;

    auto worker = [&](int si) {
        cudaSetDevice(p->device);
        double *tp = &tphase[si * 6];
        sg_batch *b = p->slot[si];
        sg_ctx *ctx = p->ctx[si];
        cudaStream_t st = ctx->stream;
        SlotIn *in = p->in[si];
        // chunk-local offsets live in pinned memory so that their copy is asynchronous too
        if (in->h_off_cap < chunk_reads + 1) {
            for (int j = 0; j < 2; ++j) {
                if (in->h_off[j]) cudaFreeHost(in->h_off[j]);
                in->h_off[j] = nullptr;
                cudaMallocHost((void **) &in->h_off[j], (chunk_reads + 1) * sizeof(uint64_t));
            }
            in->h_off_cap = chunk_reads + 1;
        }
        // uploads chunk c into input buffer j on the slot's upload stream
        auto upload = [&](uint64_t c, int j) -> int {
            const uint64_t r0 = c * chunk_reads, r1 = std::min(n_reads, r0 + chunk_reads), nr = r1 - r0;
            const uint64_t nb = off[r1] - off[r0];
            if (!in->h_off[j] || in->bases[j].reserve(nb + 64) || in->off[j].reserve((nr + 1) * sizeof(uint64_t))) return SG_E_NOMEM;
            for (uint64_t i = 0; i <= nr; ++i) in->h_off[j][i] = off[r0 + i] - off[r0];
            const char *src = bases + off[r0];
            if (!in_pinned && nb) {
                if (in->h_bases_cap[j] < nb) {
                    if (in->h_bases[j]) cudaFreeHost(in->h_bases[j]);
                    in->h_bases[j] = nullptr; in->h_bases_cap[j] = 0;
                    if (cudaMallocHost((void **) &in->h_bases[j], nb + nb / 8 + 4096) != cudaSuccess) return SG_E_NOMEM;
                    in->h_bases_cap[j] = nb + nb / 8 + 4096;
                }
                // the buffer was last read by the upload of two rounds ago, which the extract of that chunk waited for
                memcpy(in->h_bases[j], src, nb);
                src = in->h_bases[j];
            }
            if (nb && cudaMemcpyAsync(in->bases[j].p, src, nb, cudaMemcpyHostToDevice, in->up) != cudaSuccess) return SG_E_CUDA;
            if (cudaMemcpyAsync(in->off[j].p, in->h_off[j], (nr + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, in->up) != cudaSuccess) return SG_E_CUDA;
            if (cudaEventRecord(in->ready[j], in->up) != cudaSuccess) return SG_E_CUDA;
            b->h2d_bytes += nb + (nr + 1) * sizeof(uint64_t);
            return SG_OK;
        };
        int cur = 0;
        int up_rc = (uint64_t) si < n_chunks ? upload(si, 0) : SG_OK;
        for (uint64_t c = si; c < n_chunks; c += p->n_slots, cur ^= 1) {
            const uint64_t r0 = c * chunk_reads, r1 = std::min(n_reads, r0 + chunk_reads), nr = r1 - r0;
            int rc = up_rc;
            std::string msg;
            if (rc) msg = "upload failed";
            double t0 = now();
            std::vector<uint64_t> l_hs(nr + 1), l_rl(nr + 1), l_scm(nr + 1);
            const uint64_t *loff = in->h_off[cur];
            sg_batch_set_sid_base(b, p->sid_base + r0);
            sg_extract_sizes_t z;
            memset(&z, 0, sizeof(z));
            {
                std::lock_guard<std::mutex> g(mu);
                if (first_err && !rc) rc = first_err;
            }
            if (!rc && cudaStreamWaitEvent(st, in->ready[cur], 0) != cudaSuccess) rc = SG_E_CUDA;
            if (!rc) rc = sg_batch_set_reads_device(b, in->bases[cur].p, (const uint64_t *) in->off[cur].p, nr, off[r1] - off[r0]);
            // the slot's next chunk starts to travel now: its buffer was last read by the extract of two rounds ago
            if (!rc && c + p->n_slots < n_chunks) up_rc = upload(c + p->n_slots, cur ^ 1);
            if (!rc) rc = sg_extract(b, k, s);
            if (!rc) rc = sg_extract_sizes(b, &z);
            if (rc && msg.empty()) msg = ctx->err;
            double t1 = now(); tp[0] += t1 - t0;
            // ---- ordered section: claim this chunk's place in every output ----
            Totals base;
            uint64_t hcap = 0;                             // capacity positions of this chunk (multiple of 64)
            {
                std::unique_lock<std::mutex> g(mu);
                cv.wait(g, [&] { return next_chunk == c; });
                base = tot;
                if (!rc && !first_err) {
                    if (tot.scm + z.n_syncmers > capN || (!cb && ((out->hoco_s_buf && tot.hs + z.hoco_s_bytes > caps->hoco_s_bytes) ||
                            (out->ho_rl_buf && tot.rl + z.ho_rl_bytes > caps->ho_rl_bytes) || tot.amb + z.n_ambiguous > caps->max_ambiguous ||
                            tot.lrl + z.n_long_runs > caps->max_long_runs))) {
                        rc = SG_E_NOMEM; msg = "caller capacities too small";
                        if (cb) p->overflowed = true;
                    }
                }
                if (!rc && !first_err) {
                    for (uint64_t i = 0; i < nr; ++i) hcap += (loff[i + 1] - loff[i] + 63) & ~63ull;
                    tot.hoff += hcap; tot.hs += z.hoco_s_bytes; tot.rl += z.ho_rl_bytes; tot.scm += z.n_syncmers;
                    tot.amb += z.n_ambiguous; tot.lrl += z.n_long_runs; tot.hoco += z.hoco_bases;
                } else if (!first_err) { first_err = rc; first_msg = msg; }
                ++next_chunk;
                cv.notify_all();
            }
            double t2 = now(); tp[1] += t2 - t1;
            if (rc || first_err) continue;
            // ---- append to the master batch (device to device, this slot's stream) ----
            const uint64_t N = z.n_syncmers;
            if (hcap) cudaMemcpyAsync((uint8_t *) M->hoco_s.p + base.hoff / 4, b->hoco_s.p, hcap / 4, cudaMemcpyDeviceToDevice, st);
            if (keep_rl) {
                if (hcap) cudaMemcpyAsync((uint8_t *) M->ho_rl.p + base.hoff, b->ho_rl.p, hcap, cudaMemcpyDeviceToDevice, st);
                if (z.n_long_runs) {
                    add_offset32_kernel<<<nblk(z.n_long_runs, 256), 256, 0, st>>>((const uint32_t *) b->lrl_sid.p, (uint32_t *) M->lrl_sid.p + base.lrl, z.n_long_runs, (uint32_t) r0);
                    cudaMemcpyAsync((uint32_t *) M->lrl_idx.p + base.lrl, b->lrl_idx.p, z.n_long_runs * 4, cudaMemcpyDeviceToDevice, st);
                    cudaMemcpyAsync((uint32_t *) M->lrl_val.p + base.lrl, b->lrl_val.p, z.n_long_runs * 4, cudaMemcpyDeviceToDevice, st);
                    ctx->count_launch(SG_T_PLACE, 1);
                }
            }
            cudaMemcpyAsync((uint32_t *) M->hoco_l.p + r0, b->hoco_l.p, nr * 4, cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync((uint32_t *) M->n_scm.p + r0, b->n_scm.p, nr * 4, cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync((uint32_t *) M->n_amb.p + r0, b->n_amb.p, nr * 4, cudaMemcpyDeviceToDevice, st);
            add_offset_kernel<<<nblk(nr + 1, 256), 256, 0, st>>>((const uint64_t *) b->hoff.p, (uint64_t *) M->hoff.p + r0, nr + 1, base.hoff);
            add_offset_kernel<<<nblk(nr + 1, 256), 256, 0, st>>>((const uint64_t *) b->scm_off.p, (uint64_t *) M->scm_off.p + r0, nr + 1, base.scm);
            ctx->count_launch(SG_T_PLACE, 2);
            if (N) {
                cudaMemcpyAsync((uint64_t *) M->key.p + base.scm, b->key.p, N * 8, cudaMemcpyDeviceToDevice, st);
                cudaMemcpyAsync((uint64_t *) M->occ.p + base.scm, b->occ.p, N * 8, cudaMemcpyDeviceToDevice, st);
                cudaMemcpyAsync((uint32_t *) M->m_pos.p + base.scm, b->m_pos.p, N * 4, cudaMemcpyDeviceToDevice, st);
                cudaMemcpyAsync((uint64_t *) M->s_mer.p + base.scm, b->s_mer.p, N * 8, cudaMemcpyDeviceToDevice, st);
                cudaMemcpyAsync((uint64_t *) M->fp.p + base.scm, b->fp.p, N * 8, cudaMemcpyDeviceToDevice, st);
            }
            if (cb) {
                // ---- callback form: the chunk lands in this slot's pinned staging and is handed over ----
                sg_extract_out_t o;
                memset(&o, 0, sizeof(o));
                o.hoco_l = (uint32_t *) in->staged(0, (nr + 1) * 4); o.n_scm = (uint32_t *) in->staged(1, (nr + 1) * 4);
                o.hoco_s_off = (uint64_t *) in->staged(2, (nr + 2) * 8); o.ho_rl_off = (uint64_t *) in->staged(3, (nr + 2) * 8);
                o.scm_off = (uint64_t *) in->staged(4, (nr + 2) * 8);
                o.hoco_s_buf = (uint8_t *) in->staged(5, z.hoco_s_bytes + 64); o.ho_rl_buf = (uint8_t *) in->staged(6, keep_rl ? 64 : z.ho_rl_bytes + 64);
                o.m_pos = (uint32_t *) in->staged(7, (N + 1) * 4); o.s_mer = (uint64_t *) in->staged(8, (N + 1) * 8); o.k_mer = (uint64_t *) in->staged(9, (N + 1) * 8);
                o.amb_sid = (uint32_t *) in->staged(10, (z.n_ambiguous + 1) * 4); o.amb_pos = (uint32_t *) in->staged(11, (z.n_ambiguous + 1) * 4);
                o.lrl_sid = (uint32_t *) in->staged(12, (z.n_long_runs + 1) * 4); o.lrl_idx = (uint32_t *) in->staged(13, (z.n_long_runs + 1) * 4);
                o.lrl_val = (uint32_t *) in->staged(14, (z.n_long_runs + 1) * 4);
                bool ok = true;
                for (int j = 0; j < 15; ++j) ok = ok && in->stage[j];
                double t3 = now(); tp[2] += t3 - t2;
                if (keep_rl) o.ho_rl_buf = nullptr;        // stays on the device
                if (p->keep_hs) o.hoco_s_buf = nullptr;
                rc = ok ? sg_extract_download(b, &o) : SG_E_NOMEM;
                double t4 = now(); tp[3] += t4 - t3;
                if (!rc) rc = cb(cb_user, r0, nr, &o, &z);
                if (rc) {
                    std::lock_guard<std::mutex> g(mu);
                    if (!first_err) { first_err = rc; first_msg = rc == SG_E_NOMEM ? "pinned staging allocation failed" : ctx->err; }
                }
                tp[4] += now() - t4;
                continue;
            }
            // ---- this chunk's slice of the caller's arrays ----
            sg_extract_out_t o = *out;
            if (o.hoco_l) o.hoco_l += r0;
            if (o.n_scm) o.n_scm += r0;
            // offset arrays go through chunk-local copies: entry nr of one chunk is entry 0 of the next
            o.hoco_s_off = l_hs.data(); o.ho_rl_off = l_rl.data(); o.scm_off = l_scm.data();
            if (o.hoco_s_buf) o.hoco_s_buf += base.hs;
            if (o.ho_rl_buf) o.ho_rl_buf += base.rl;
            if (o.m_pos) o.m_pos += base.scm;
            if (o.s_mer) o.s_mer += base.scm;
            if (o.k_mer) o.k_mer += base.scm;
            if (o.amb_sid) o.amb_sid += base.amb;
            if (o.amb_pos) o.amb_pos += base.amb;
            if (o.lrl_sid) o.lrl_sid += base.lrl;
            if (o.lrl_idx) o.lrl_idx += base.lrl;
            if (o.lrl_val) o.lrl_val += base.lrl;
            double t3 = now(); tp[2] += t3 - t2;
            rc = sg_extract_download(b, &o);              // synchronises this slot's stream
            double t4 = now(); tp[3] += t4 - t3;
            if (!rc) {
                // chunk-local offsets and read ids -> global; only the last chunk writes the closing entry
                const uint64_t ne = nr + (r1 == n_reads ? 1 : 0);
                for (uint64_t i = 0; i < ne; ++i) {
                    if (out->hoco_s_off) out->hoco_s_off[r0 + i] = l_hs[i] + base.hs;
                    if (out->ho_rl_off) out->ho_rl_off[r0 + i] = l_rl[i] + base.rl;
                    if (out->scm_off) out->scm_off[r0 + i] = l_scm[i] + base.scm;
                }
                for (uint64_t i = 0; i < z.n_ambiguous && o.amb_sid; ++i) o.amb_sid[i] += (uint32_t) r0;
                for (uint64_t i = 0; i < z.n_long_runs && o.lrl_sid; ++i) o.lrl_sid[i] += (uint32_t) r0;
            } else {
                std::lock_guard<std::mutex> g(mu);
                if (!first_err) { first_err = rc; first_msg = ctx->err; }
            }
            tp[4] += now() - t4;
        }
        cudaStreamSynchronize(in->up);
        cudaStreamSynchronize(st);
    };

    std::vector<std::thread> th;
    for (int i = 0; i < p->n_slots; ++i) th.emplace_back(worker, i);
    for (auto &t : th) t.join();
    if (trace)
        for (int i = 0; i < p->n_slots; ++i)
            fprintf(stderr, "[sg_pipe] slot %d: upload+extract %.1f ms, ordered wait %.1f, append enqueue %.1f, download %.1f, fix-up %.1f\n", i,
                    tphase[i * 6] * 1e3, tphase[i * 6 + 1] * 1e3, tphase[i * 6 + 2] * 1e3, tphase[i * 6 + 3] * 1e3, tphase[i * 6 + 4] * 1e3);
    if (first_err) return fail(first_err, first_msg);
    if (cudaDeviceSynchronize() != cudaSuccess) return fail(SG_E_CUDA, cudaGetErrorString(cudaGetLastError()));

    // the master batch now looks like the result of one big sg_extract
    M->d_bases = nullptr; M->d_off = nullptr;
    M->n_reads = n_reads; M->total_bases = total; M->sid_base = p->sid_base;
    M->k = k; M->s = s;
    M->n_syncmers = tot.scm; M->n_amb_total = tot.amb; M->n_lrl_total = tot.lrl; M->hoco_bases = tot.hoco;
    M->extracted = true; M->counted = M->sorted = M->adopted = M->sizes_known = M->have_kid_local = false;
    M->pipe_fed = true;
    M->rl_resident = keep_rl;
    M->lrl_sorted = false;
    M->keys_are_ids = false;
    sizes->n_reads = n_reads; sizes->n_syncmers = tot.scm; sizes->hoco_bases = tot.hoco;
    sizes->hoco_s_bytes = tot.hs; sizes->ho_rl_bytes = tot.rl; sizes->n_ambiguous = tot.amb; sizes->n_long_runs = tot.lrl;
    (void) mctx;
    return SG_OK;
}

int sg_pipe_run_host(sg_pipe *p, const char *bases, const uint64_t *off, uint64_t n_reads, int k, int s,
        uint64_t chunk_reads, const sg_extract_out_t *out, const sg_pipe_caps_t *caps, sg_extract_sizes_t *sizes)
{
    return pipe_run(p, bases, off, n_reads, k, s, chunk_reads, out, caps, nullptr, nullptr, sizes);
}

int sg_pipe_run_host_cb(sg_pipe *p, const char *bases, const uint64_t *off, uint64_t n_reads, int k, int s,
        uint64_t chunk_reads, sg_pipe_chunk_fn fn, void *user, sg_extract_sizes_t *sizes)
{
    if (!fn) return SG_E_ARG;
    return pipe_run(p, bases, off, n_reads, k, s, chunk_reads, nullptr, nullptr, fn, user, sizes);
}

} // extern "C"
