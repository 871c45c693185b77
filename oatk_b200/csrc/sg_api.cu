// sg_api.cu -- the C ABI declared in include/syncgpu.h: contexts, batches, buffer
// management, kernel sequencing and host transfers. No compute happens here.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cctype>
#include <sched.h>
#include <unistd.h>
#include <sys/syscall.h>
#include <string>
#include <vector>
#include <algorithm>
#include "sg_common.cuh"
#include "sg_internal.h"
#include "sg_host.h"
#include "../../include/syncgpu.h"

using namespace sg;

namespace sg {

int DevBuf::reserve(size_t n)
{
    if (n <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) {
        if (cudaMalloc(&p, n) != cudaSuccess) { cudaGetLastError(); return SG_E_NOMEM; }
        want = n;
    }
    cap = want;
    return 0;
}
void DevBuf::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }

} // namespace sg

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return SG_E_CUDA; } } while (0)
#define RS(buf, bytes) do { if ((buf).reserve(bytes)) { ctx->err = "device allocation of " + std::to_string((size_t)(bytes)) + " bytes failed"; return SG_E_NOMEM; } } while (0)
#define LAUNCHED(stage, expr) do { int n_ = (expr); if (n_ < 0) return n_; ctx->count_launch(stage, n_); } while (0)

void sg_ctx::count_launch(int stage, int n)
{
    launches += n;
    stage_launch[stage] += n;
}
void sg_ctx::t_begin(int stage)
{
    if (!timing) return;
    cudaEventRecord(ev[stage][0], stream);
}
void sg_ctx::t_end(int stage)
{
    if (!timing) return;
    cudaEventRecord(ev[stage][1], stream);
    ev_used[stage] = true;
}

void sg_ctx::lap(const char *name)
{
    if (laps_on < 0) { const char *e = getenv("SG_LAPS"); laps_on = e && *e && *e != '0'; }
    if (!laps_on || !timing) return;
    if (lap_n == lap_ev.size()) { cudaEvent_t e; cudaEventCreate(&e); lap_ev.push_back(e); lap_name.push_back(name); }
    lap_name[lap_n] = name;
    cudaEventRecord(lap_ev[lap_n++], stream);
}
void sg_ctx::laps_print(const char *tag)
{
    if (laps_on <= 0 || lap_n < 2) { lap_n = 0; return; }
    cudaEventSynchronize(lap_ev[lap_n - 1]);
    std::string line = std::string("[sg laps dev ") + std::to_string(device) + "] " + tag + ":";
    for (size_t i = 1; i < lap_n; ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, lap_ev[i - 1], lap_ev[i]);
        char buf[96];
        snprintf(buf, sizeof(buf), " %s=%.3f", lap_name[i], ms);
        line += buf;
    }
    fprintf(stderr, "%s\n", line.c_str());
    lap_n = 0;
}

extern "C" {

const char *sg_strerror(int code)
{
    switch (code) {
        case SG_OK: return "ok";
        case SG_E_CUDA: return "CUDA error";
        case SG_E_ARG: return "bad argument";
        case SG_E_NOMEM: return "out of memory";
        case SG_E_LIMIT: return "reference size limit exceeded";
        case SG_E_KSIZE: return "k too large for the scan kernel's shared-memory window";
        case SG_E_SMER_CONFLICT: return "identical kmers have different smers";
        case SG_E_EMPTY: return "empty syncmer collection";
        case SG_E_STATE: return "call order violated";
        case SG_E_COLLISION: return "64-bit k-mer hash collision across GPUs";
        case SG_E_COMM: return "NCCL unavailable or a collective failed";
    }
    return "unknown error";
}

int sg_ctx_create(int device, sg_ctx **out)
{
    if (!out) return SG_E_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); return SG_E_CUDA; }
    if (cudaSetDevice(device) != cudaSuccess) return SG_E_CUDA;
    sg_ctx *ctx = new sg_ctx();
    ctx->device = device;
    for (int i = 0; i < SG_T_N; ++i) { cudaEventCreate(&ctx->ev[i][0]); cudaEventCreate(&ctx->ev[i][1]); }
    *out = ctx;
    return SG_OK;
}

void sg_ctx_destroy(sg_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < SG_T_N; ++i) { cudaEventDestroy(ctx->ev[i][0]); cudaEventDestroy(ctx->ev[i][1]); }
    delete ctx;
}

// Host placement for one process per GPU: run the calling thread (and the threads it starts later) on the CPUs next to the device and
// prefer that NUMA node for the pages it touches from now on, so that pinned buffers allocated afterwards are read by the copy
// engines without crossing the socket link. Everything comes from sysfs; a container that hides it leaves the process as it was.
int sg_host_bind_near_device(int device, int flags, int *numa_node, int *n_cpus)
{
    if (numa_node) *numa_node = -1;
    if (n_cpus) *n_cpus = 0;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int) sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return SG_E_CUDA; }
    for (char *c = bus; *c; ++c) *c = (char) tolower((unsigned char) *c);
    char path[160];
    int node = -1;
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    if (FILE *f = fopen(path, "r")) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    cpu_set_t now, want;
    CPU_ZERO(&want);
    int bound = 0;
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE *fl = (flags & 1) ? fopen(path, "r") : nullptr;
    if (FILE *f = fl) {
        if (sched_getaffinity(0, sizeof(now), &now) == 0) {
            int a, b2;
            // "0-15,32-47": ranges and single numbers separated by commas
            while (fscanf(f, "%d", &a) == 1) {
                b2 = a;
                int ch = fgetc(f);
                if (ch == '-') { if (fscanf(f, "%d", &b2) != 1) break; ch = fgetc(f); }
                for (int c = a; c <= b2 && c < CPU_SETSIZE; ++c) if (c >= 0 && CPU_ISSET(c, &now)) { CPU_SET(c, &want); ++bound; }
                if (ch != ',') break;
            }
            if (bound > 0 && sched_setaffinity(0, sizeof(want), &want) != 0) bound = 0;
        }
        fclose(f);
    }
    if ((flags & 2) && node >= 0 && node < 1024) {
        unsigned long mask[16] = {0};
        mask[node / (8 * sizeof(unsigned long))] = 1ul << (node % (8 * sizeof(unsigned long)));
        // MPOL_PREFERRED = 1: fall back to other nodes rather than fail when the node is full or not allowed
        if (syscall(SYS_set_mempolicy, 1, mask, (unsigned long) (sizeof(mask) * 8)) != 0) node = -1 - node;
    }
    if (numa_node) *numa_node = node;
    if (n_cpus) *n_cpus = bound;
    return SG_OK;
}

int sg_ctx_set_stream(sg_ctx *ctx, void *s) { if (!ctx) return SG_E_ARG; ctx->stream = (cudaStream_t) s; return SG_OK; }
int sg_ctx_sync(sg_ctx *ctx) { if (!ctx) return SG_E_ARG; CK(cudaStreamSynchronize(ctx->stream)); return SG_OK; }
const char *sg_last_error(sg_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }
uint64_t sg_ctx_launches(sg_ctx *ctx) { return ctx ? ctx->launches : 0; }
int sg_ctx_enable_timing(sg_ctx *ctx, int on) { if (!ctx) return SG_E_ARG; ctx->timing = on != 0; return SG_OK; }

int sg_ctx_timings(sg_ctx *ctx, float *ms, uint32_t *launches)
{
    if (!ctx) return SG_E_ARG;
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < SG_T_N; ++i) {
        float t = 0;
        if (ctx->ev_used[i]) cudaEventElapsedTime(&t, ctx->ev[i][0], ctx->ev[i][1]);
        if (ms) ms[i] = t;
        if (launches) launches[i] = ctx->stage_launch[i];
        ctx->ev_used[i] = false;
        ctx->stage_launch[i] = 0;
    }
    return SG_OK;
}

int sg_batch_create(sg_ctx *ctx, sg_batch **out)
{
    if (!ctx || !out) return SG_E_ARG;
    sg_batch *b = new sg_batch();
    b->ctx = ctx;
    *out = b;
    return SG_OK;
}

void sg_batch_destroy(sg_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    delete b;
}

static void reset_state(sg_batch *b)
{
    b->extracted = b->counted = b->sizes_known = b->sorted = b->adopted = false;
    b->n_adopted = 0;
    b->have_kid_local = false;
    b->pipe_fed = false;
    b->rl_resident = false;
    b->lrl_sorted = false;
    b->keys_are_ids = false;
    b->tup_valid = false;
    b->atup_valid = false;
    b->asoa_valid = true;
    b->k = b->s = 0;
    b->n_syncmers = 0;
}

int sg_batch_set_reads_host(sg_batch *b, const char *bases, const uint64_t *off, uint64_t n_reads)
{
    if (!b || !off || (!bases && n_reads && off[n_reads])) return SG_E_ARG;
    sg_ctx *ctx = b->ctx;
    if (n_reads > 0xFFFFFFFFull) return SG_E_LIMIT;
    for (uint64_t i = 0; i < n_reads; ++i) {
        if (off[i + 1] < off[i]) return SG_E_ARG;
        if (off[i + 1] - off[i] > 0x7FFFFFFFull) return SG_E_LIMIT;
    }
    CK(cudaSetDevice(ctx->device));
    const uint64_t total = n_reads ? off[n_reads] : 0;
    RS(b->own_bases, total + 64);
    RS(b->own_off, (n_reads + 1) * sizeof(uint64_t));
    if (total) CK(cudaMemcpyAsync(b->own_bases.p, bases, total, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(b->own_off.p, off, (n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    b->d_bases = (const uint8_t *) b->own_bases.p;
    b->d_off = (const uint64_t *) b->own_off.p;
    b->n_reads = n_reads;
    b->total_bases = total;
    b->h2d_bytes += total + (n_reads + 1) * sizeof(uint64_t);
    reset_state(b);
    return SG_OK;
}

int sg_batch_set_reads_device(sg_batch *b, const void *d_bases, const uint64_t *d_off, uint64_t n_reads, uint64_t total_bases)
{
    if (!b || !d_off || (!d_bases && total_bases)) return SG_E_ARG;
    if (((uintptr_t) d_bases & 15u) != 0) return SG_E_ARG;
    if (n_reads > 0xFFFFFFFFull) return SG_E_LIMIT;
    b->d_bases = (const uint8_t *) d_bases;
    b->d_off = d_off;
    b->n_reads = n_reads;
    b->total_bases = total_bases;
    reset_state(b);
    return SG_OK;
}

int sg_batch_set_sid_base(sg_batch *b, uint64_t sid_base) { if (!b) return SG_E_ARG; b->sid_base = sid_base; return SG_OK; }

// ------------------------------------------------------------------ a2-a4
static int run_extract(sg_batch *b, uint64_t rec_cap_hint)
{
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    const uint64_t n = b->n_reads;
    const uint64_t cap = b->total_bases + 64 * n + 64;            // >= hoff[n]
    RS(b->hoff, (n + 1) * sizeof(uint64_t));
    RS(b->scan_tmp, scan_tmp_words(std::max<uint64_t>(n, 1)) * sizeof(uint64_t));
    RS(b->hoco_s, cap / 4 + 64);
    RS(b->ho_rl, cap + 64);
    RS(b->nbits, cap / 8 + 64);
    RS(b->hoco_l, (n + 1) * sizeof(uint32_t));
    RS(b->n_amb, (n + 1) * sizeof(uint32_t));
    RS(b->n_scm, (n + 1) * sizeof(uint32_t));
    RS(b->scm_off, (n + 1) * sizeof(uint64_t));
    RS(b->counters, 8 * sizeof(unsigned long long));
    if (b->amb_cap == 0) b->amb_cap = 1 << 16;
    if (b->lrl_cap == 0) b->lrl_cap = 1 << 14;
    RS(b->amb_sid, b->amb_cap * 4); RS(b->amb_pos, b->amb_cap * 4);
    RS(b->lrl_sid, b->lrl_cap * 4); RS(b->lrl_idx, b->lrl_cap * 4); RS(b->lrl_val, b->lrl_cap * 4);
    // expected syncmers: 2 per window of q hoco positions; leave generous room, grow on overflow
    const uint64_t q = (uint64_t) (b->k - b->s + 1);
    uint64_t rec_cap = std::max<uint64_t>(rec_cap_hint, 4 * (b->total_bases / q + n) + 1024);
    b->rec_cap = rec_cap;
    RS(b->rec_sid, rec_cap * 4); RS(b->rec_idx, rec_cap * 4); RS(b->rec_mpos, rec_cap * 4);

    CK(cudaMemsetAsync(b->counters.p, 0, 8 * sizeof(unsigned long long), st));
    unsigned long long *cnt = (unsigned long long *) b->counters.p;

    ctx->t_begin(SG_T_ENCODE);
    LAUNCHED(SG_T_ENCODE, launch_capacity_offsets(b->d_off, (uint64_t *) b->hoff.p, n, (uint64_t *) b->scan_tmp.p, st));
    EncodeArgs E;
    E.bases = b->d_bases; E.off = b->d_off; E.hoff = (const uint64_t *) b->hoff.p;
    E.n_reads = n; E.work = reinterpret_cast<unsigned int *>(cnt + 4);
    E.hoco_s = (uint8_t *) b->hoco_s.p; E.ho_rl = (uint8_t *) b->ho_rl.p; E.nbits = (uint8_t *) b->nbits.p;
    E.hoco_l = (uint32_t *) b->hoco_l.p; E.n_amb = (uint32_t *) b->n_amb.p;
    E.amb_count = cnt + 0; E.lrl_count = cnt + 1;
    E.amb_cap = b->amb_cap; E.lrl_cap = b->lrl_cap;
    E.amb_sid = (uint32_t *) b->amb_sid.p; E.amb_pos = (uint32_t *) b->amb_pos.p;
    E.lrl_sid = (uint32_t *) b->lrl_sid.p; E.lrl_idx = (uint32_t *) b->lrl_idx.p; E.lrl_val = (uint32_t *) b->lrl_val.p;
    LAUNCHED(SG_T_ENCODE, launch_encode(E, st));
    ctx->t_end(SG_T_ENCODE);

    ctx->t_begin(SG_T_SCAN);
    ScanArgs S;
    S.hoff = (const uint64_t *) b->hoff.p; S.hoco_s = (const uint8_t *) b->hoco_s.p; S.nbits = (const uint8_t *) b->nbits.p;
    S.hoco_l = (const uint32_t *) b->hoco_l.p; S.n_amb = (const uint32_t *) b->n_amb.p;
    S.k = b->k; S.s = b->s;
    S.n_reads = n; S.work = reinterpret_cast<unsigned int *>(cnt + 3);
    S.n_scm = (uint32_t *) b->n_scm.p;
    S.rec_count = cnt + 2; S.rec_cap = rec_cap;
    S.rec_sid = (uint32_t *) b->rec_sid.p; S.rec_idx = (uint32_t *) b->rec_idx.p; S.rec_mpos = (uint32_t *) b->rec_mpos.p;
    RS(b->scan_defer, (n + 1) * 4 * sizeof(uint32_t));
    RS(b->scan_xring, scan_exact_scratch_bytes(b->k, b->s, nullptr));
    S.defer_count = reinterpret_cast<unsigned int *>(cnt + 5); S.work2 = reinterpret_cast<unsigned int *>(cnt + 6);
    S.defer = (uint32_t *) b->scan_defer.p; S.xring = (uint64_t *) b->scan_xring.p;
    LAUNCHED(SG_T_SCAN, launch_scan(S, st));
    ctx->t_end(SG_T_SCAN);

    ctx->t_begin(SG_T_PLACE);
    LAUNCHED(SG_T_PLACE, launch_scan_u32_u64((const uint32_t *) b->n_scm.p, (uint64_t *) b->scm_off.p, n, (uint64_t *) b->scan_tmp.p, st));
    ctx->t_end(SG_T_PLACE);
    return SG_OK;
}

int sg_extract(sg_batch *b, int k, int s)
{
    if (!b || !b->d_off) return SG_E_ARG;
    if (!(s > 0 && s < 32 && k > s)) return SG_E_ARG;               // reference assert, syncmer.c:251
    sg_ctx *ctx = b->ctx;
    CK(cudaSetDevice(ctx->device));
    {
        ScanGeom g; size_t smem;
        if (scan_geometry(k, s, &g, &smem)) { ctx->err = "k - s + 1 exceeds the scan window"; return SG_E_KSIZE; }
    }
    reset_state(b);
    b->k = k; b->s = s;
    cudaStream_t st = ctx->stream;
    uint64_t hint = 0;
    for (int attempt = 0; attempt < 4; ++attempt) {
        int rc = run_extract(b, hint);
        if (rc) return rc;
        // the record count decides the size of everything downstream: read it back
        unsigned long long hc[6];
        CK(cudaMemcpyAsync(hc, b->counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        bool again = false;
        if (hc[0] > b->amb_cap) { b->amb_cap = hc[0] + 16; again = true; }
        if (hc[1] > b->lrl_cap) { b->lrl_cap = hc[1] + 16; again = true; }
        if (hc[2] > b->rec_cap) { hint = hc[2] + 16; again = true; }
        b->n_amb_total = hc[0]; b->n_lrl_total = hc[1]; b->n_syncmers = hc[2];
        b->n_scan_deferred = (uint32_t) hc[5];
        if (!again) break;
        if (attempt == 3) { ctx->err = "side-list capacity did not converge"; return SG_E_NOMEM; }
    }
    const uint64_t N = b->n_syncmers;
    RS(b->key, (N + 1) * 8); RS(b->occ, (N + 1) * 8); RS(b->m_pos, (N + 1) * 4); RS(b->s_mer, (N + 1) * 8); RS(b->fp, (N + 1) * 8);
    ctx->t_begin(SG_T_KMERHASH);
    KmerArgs K;
    K.hoff = (const uint64_t *) b->hoff.p; K.hoco_s = (const uint8_t *) b->hoco_s.p; K.hoco_l = (const uint32_t *) b->hoco_l.p;
    K.k = k; K.s = s; K.n_rec = N;
    K.rec_sid = (const uint32_t *) b->rec_sid.p; K.rec_idx = (const uint32_t *) b->rec_idx.p; K.rec_mpos = (const uint32_t *) b->rec_mpos.p;
    K.scm_off = (const uint64_t *) b->scm_off.p;
    K.sid_base = b->sid_base;
    K.fp = (uint64_t *) b->fp.p;
    RS(b->tup, (N + 1) * 32);
    K.tup = (uint64_t *) b->tup.p;
    b->tup_valid = true;
    K.key = (uint64_t *) b->key.p; K.occ = (uint64_t *) b->occ.p; K.m_pos = (uint32_t *) b->m_pos.p; K.s_mer = (uint64_t *) b->s_mer.p;
    LAUNCHED(SG_T_KMERHASH, launch_kmerhash(K, st));
    ctx->t_end(SG_T_KMERHASH);
    CK(cudaGetLastError());
    b->extracted = true;
    b->rl_resident = true;                     // ho_rl and the side list of this batch stay on the device until the next extract
    b->lrl_sorted = false;
    return SG_OK;
}

static int ensure_sizes(sg_batch *b)
{
    sg_ctx *ctx = b->ctx;
    if (b->sizes_known) return SG_OK;
    const uint64_t n = b->n_reads;
    cudaStream_t st = ctx->stream;
    b->h_hoco_l.resize(n);
    b->h_n_scm.resize(n);
    if (n) {
        CK(cudaMemcpyAsync(b->h_hoco_l.data(), b->hoco_l.p, n * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(b->h_n_scm.data(), b->n_scm.p, n * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    b->h_hs_off.assign(n + 1, 0); b->h_rl_off.assign(n + 1, 0); b->h_scm_off.assign(n + 1, 0);
    uint64_t hb = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t L = b->h_hoco_l[i];
        b->h_hs_off[i + 1] = b->h_hs_off[i] + ((((L + 3) / 4) + 15) & ~15ull);
        b->h_rl_off[i + 1] = b->h_rl_off[i] + ((L + 15) & ~15ull);
        b->h_scm_off[i + 1] = b->h_scm_off[i] + b->h_n_scm[i];
        hb += L;
    }
    b->hoco_bases = hb;
    b->sizes_known = true;
    return SG_OK;
}

int sg_debug_scan_info(sg_batch *b, uint64_t *deferred_reads)
{
    if (!b) return SG_E_ARG;
    if (!b->extracted) return SG_E_STATE;
    if (deferred_reads) *deferred_reads = b->n_scan_deferred;
    return SG_OK;
}

int sg_extract_sizes(sg_batch *b, sg_extract_sizes_t *out)
{
    if (!b || !out) return SG_E_ARG;
    if (!b->extracted) return SG_E_STATE;
    int rc = ensure_sizes(b);
    if (rc) return rc;
    out->n_reads = b->n_reads;
    out->n_syncmers = b->n_syncmers;
    out->hoco_bases = b->hoco_bases;
    out->hoco_s_bytes = b->h_hs_off[b->n_reads];
    out->ho_rl_bytes = b->h_rl_off[b->n_reads];
    out->n_ambiguous = b->n_amb_total;
    out->n_long_runs = b->n_lrl_total;
    return SG_OK;
}

int sg_extract_download(sg_batch *b, const sg_extract_out_t *o)
{
    if (!b || !o) return SG_E_ARG;
    if (!b->extracted) return SG_E_STATE;
    if (b->pipe_fed && (o->hoco_s_buf || o->ho_rl_buf || o->amb_sid || o->amb_pos || o->lrl_sid || o->lrl_idx || o->lrl_val))
        return SG_E_STATE;                        // the pipeline already delivered these per chunk
    sg_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_sizes(b);
    if (rc) return rc;
    const uint64_t n = b->n_reads, N = b->n_syncmers;
    if (o->hoco_l && n) memcpy(o->hoco_l, b->h_hoco_l.data(), n * 4);
    if (o->n_scm && n) memcpy(o->n_scm, b->h_n_scm.data(), n * 4);
    if (o->hoco_s_off) memcpy(o->hoco_s_off, b->h_hs_off.data(), (n + 1) * 8);
    if (o->ho_rl_off) memcpy(o->ho_rl_off, b->h_rl_off.data(), (n + 1) * 8);
    if (o->scm_off) memcpy(o->scm_off, b->h_scm_off.data(), (n + 1) * 8);
    if ((o->hoco_s_buf || o->ho_rl_buf) && n) {
        // pack the capacity-indexed device layout into the 16-byte aligned compact one, then one copy each
        const uint64_t hsb = b->h_hs_off[n], rlb = b->h_rl_off[n];
        if (o->hoco_s_buf) RS(b->pk_hs, hsb + 16);
        if (o->ho_rl_buf) RS(b->pk_rl, rlb + 16);
        RS(b->pk_hs_off, (n + 1) * 8); RS(b->pk_rl_off, (n + 1) * 8);
        CK(cudaMemcpyAsync(b->pk_hs_off.p, b->h_hs_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(b->pk_rl_off.p, b->h_rl_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
        ctx->t_begin(SG_T_PACK);
        LAUNCHED(SG_T_PACK, launch_pack(b, st, o->hoco_s_buf != nullptr, o->ho_rl_buf != nullptr));
        ctx->t_end(SG_T_PACK);
        if (o->hoco_s_buf && hsb) CK(cudaMemcpyAsync(o->hoco_s_buf, b->pk_hs.p, hsb, cudaMemcpyDeviceToHost, st));
        if (o->ho_rl_buf && rlb) CK(cudaMemcpyAsync(o->ho_rl_buf, b->pk_rl.p, rlb, cudaMemcpyDeviceToHost, st));
        b->d2h_bytes += (o->hoco_s_buf ? hsb : 0) + (o->ho_rl_buf ? rlb : 0);
    }
    if (N) {
        if (o->m_pos) CK(cudaMemcpyAsync(o->m_pos, b->m_pos.p, N * 4, cudaMemcpyDeviceToHost, st));
        if (o->s_mer) CK(cudaMemcpyAsync(o->s_mer, b->s_mer.p, N * 8, cudaMemcpyDeviceToHost, st));
        const void *ksrc = b->have_kid_local ? b->kid_local.p : ((b->counted && !b->adopted) ? b->kid.p : b->key.p);
        if (o->k_mer) CK(cudaMemcpyAsync(o->k_mer, ksrc, N * 8, cudaMemcpyDeviceToHost, st));
        b->d2h_bytes += N * 20;
    }
    CK(cudaStreamSynchronize(st));
    // side lists are tiny and arrive in atomic order: sort them on the host
    if (b->n_amb_total && (o->amb_sid || o->amb_pos)) {
        const uint64_t m = b->n_amb_total;
        std::vector<uint32_t> a(m), p(m);
        CK(cudaMemcpy(a.data(), b->amb_sid.p, m * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(p.data(), b->amb_pos.p, m * 4, cudaMemcpyDeviceToHost));
        std::vector<uint64_t> key(m);
        for (uint64_t i = 0; i < m; ++i) key[i] = (uint64_t) a[i] << 32 | p[i];
        std::sort(key.begin(), key.end());
        for (uint64_t i = 0; i < m; ++i) {
            if (o->amb_sid) o->amb_sid[i] = (uint32_t) (key[i] >> 32);
            if (o->amb_pos) o->amb_pos[i] = (uint32_t) key[i];
        }
    }
    if (b->n_lrl_total && (o->lrl_sid || o->lrl_idx || o->lrl_val)) {
        const uint64_t m = b->n_lrl_total;
        std::vector<uint32_t> a(m), p(m), v(m);
        CK(cudaMemcpy(a.data(), b->lrl_sid.p, m * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(p.data(), b->lrl_idx.p, m * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(v.data(), b->lrl_val.p, m * 4, cudaMemcpyDeviceToHost));
        std::vector<uint64_t> ord(m);
        for (uint64_t i = 0; i < m; ++i) ord[i] = i;
        std::sort(ord.begin(), ord.end(), [&](uint64_t x, uint64_t y) {
            return a[x] != a[y] ? a[x] < a[y] : p[x] < p[y]; });
        for (uint64_t i = 0; i < m; ++i) {
            if (o->lrl_sid) o->lrl_sid[i] = a[ord[i]];
            if (o->lrl_idx) o->lrl_idx[i] = p[ord[i]];
            if (o->lrl_val) o->lrl_val[i] = v[ord[i]];
        }
    }
    return SG_OK;
}

} // extern "C"
