// sg_host.h -- host-side state behind the opaque handles of include/syncgpu.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/syncgpu.h"

namespace sg {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);     // grow-only; contents are not preserved
    void release();
    ~DevBuf() { release(); }
};

} // namespace sg

struct sg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    bool timing = false;
    cudaEvent_t ev[SG_T_N][2];
    bool ev_used[SG_T_N] = {};
    uint32_t stage_launch[SG_T_N] = {};
    void count_launch(int stage, int n);
    void t_begin(int stage);
    void t_end(int stage);
    // SG_LAPS=1 in the environment: event-timed laps inside a stage, printed to stderr when the stage closes (diagnostics only)
    int laps_on = -1;
    std::vector<cudaEvent_t> lap_ev;
    std::vector<const char *> lap_name;
    size_t lap_n = 0;
    void lap(const char *name);
    void laps_print(const char *tag);
};

struct sg_batch {
    sg_ctx *ctx = nullptr;
    // input
    const uint8_t *d_bases = nullptr;
    const uint64_t *d_off = nullptr;
    uint64_t n_reads = 0, total_bases = 0, sid_base = 0;
    sg::DevBuf own_bases, own_off;
    // state
    bool extracted = false, counted = false, sizes_known = false;
    int k = 0, s = 0;
    uint64_t n_syncmers = 0, n_amb_total = 0, n_lrl_total = 0, hoco_bases = 0;
    uint64_t n_scan_deferred = 0;                // reads of the last sg_extract that went through scan_exact_kernel
    uint64_t h2d_bytes = 0, d2h_bytes = 0;
    // a2-a4 device results (capacity-indexed layout, see sg_common.cuh)
    sg::DevBuf hoff, scan_tmp, hoco_s, ho_rl, nbits, hoco_l, n_amb, n_scm, scm_off, counters;
    uint64_t amb_cap = 0, lrl_cap = 0, rec_cap = 0;
    sg::DevBuf amb_sid, amb_pos, lrl_sid, lrl_idx, lrl_val;
    sg::DevBuf rec_sid, rec_idx, rec_mpos, rec_smer;
    sg::DevBuf scan_defer, scan_xring;           // reads deferred to scan_exact_kernel and its hash rings
    sg::DevBuf key, occ, m_pos, s_mer, fp;      // read order, one entry per syncmer (fp: second hash)
    sg::DevBuf tup;                              // read order: (occ, s_mer, fp, key) as 32-byte records (sg_extract only)
    bool tup_valid = false;
    bool atup_valid = false;
    uint64_t range_lo = 0;                       // adopted tuples: first hash of this GPU's range and how far (h - range_lo) can be shifted up
    int range_lsh = 0;                           // without overflow (set by the exchange; 0 / 0 otherwise)
    bool asoa_valid = true;                      // akey / aocc / asmer / afp hold the adopted tuples (false: only the records in tup do; sg::ensure_adopted_soa)                     // tup holds the records of the ADOPTED tuple set (sg_tuples_adopt)
    // download staging
    sg::DevBuf pk_hs, pk_rl, pk_hs_off, pk_rl_off;
    std::vector<uint32_t> h_hoco_l, h_n_scm;
    std::vector<uint64_t> h_hs_off, h_rl_off, h_scm_off;
    // a5/a6 device results
    sg::DevBuf kid;                              // read order: id << 1
    sg::DevBuf skey, sval, skey_alt, sval_alt, sort_tmp, sort_fix;
    int sort_low_bits = 24;                      // radix passes skip these low hash bits, a repair pass handles them (0: full sort)
    int pack_bits = 32;                          // hash bits the packed (hash-top | index) sort orders; a repair pass handles the rest
    bool sort_fell_back = false;
    bool keys_are_ids = false;                   // sg_batch_set_lists_host: key[] holds id << 1 | corrected, not hashes
    uint64_t n_sort_repairs = 0;                 // out-of-order pairs the repair pass saw
    sg::DevBuf socc, ssmer, flags, ids, ids_tmp, differs, cls, starts, stat_dev, stat_dev2, skey2, sval2;
    bool sorted = false;
    bool exact_verify = false;                   // sg_batch_set_exact_verify: compare packed k-mers instead of fingerprints
    int hash_bits = 64;                          // < 64 only through sg_debug_set_hash_bits (tests)
    sg::DevBuf scm_h, scm_s, scm_cov, scm_occ_off, status;
    uint64_t n_unique = 0, n_collisions = 0;
    bool has_conflict = false;                   // the last sg_count ended with SG_E_SMER_CONFLICT: what the reference would have printed
    uint64_t conflict[5] = {};                   // k-mer hash, first s-mer code, its read, disagreeing code, its read
    sg::DevBuf smer_pairs;                       // multi-GPU: (s-mer code, local count) pairs of the last sg_stat
    uint64_t smer_slots = 0;                     // slots of the s-mer tally table left by sg_stat (0: none)
    // a7
    sg::DevBuf arc_keys, arc_vals, arc_out, arc_okey, arc_oval, arc_okey_alt, arc_oval_alt;
    uint64_t n_arcs = 0, arc_cap = 0;
    // multi-GPU exchange: tuples adopted from the peers replace the local tuple set for a5/a6
    sg::DevBuf tuples, part_counts, akey, aocc, asmer, afp, kid_local, sfp;
    bool have_kid_local = false;
    bool adopted = false;
    bool pipe_fed = false;                       // filled by sg_pipe_run_host: no raw reads on the device, ho_rl only if rl_resident
    bool rl_resident = false;                    // ho_rl (capacity layout) and the long-run side list are on the device: sg_runlen_sums works
    bool lrl_sorted = false;                     // lrl_key / lrl_sval hold the side list ordered by (read, hoco index)
    sg::DevBuf lrl_key, lrl_sval, lrl_key_alt, lrl_val_alt, rq_off, rq_occ, rq_out;
    // f2 (sg_ec_filter / sg_ec_correct): kept with the batch, because cudaMalloc / cudaFree per call cost more than the kernels
    sg::DevBuf ec_del, ec_av, ec_aw, ec_at, ec_al, ec_arena, ec_outk, ec_outp, ec_off, ec_n, ec_misc, ec_over, ec_err, ec_prev, ec_flag, ec_ex, ec_tmp, ec_live;
    uint64_t n_adopted = 0;
    // the tuple set a5/a6 work on
    const uint64_t *t_key() const { return (const uint64_t *) (adopted ? akey.p : key.p); }
    const uint64_t *t_occ() const { return (const uint64_t *) (adopted ? aocc.p : occ.p); }
    const uint64_t *t_smer() const { return (const uint64_t *) (adopted ? asmer.p : s_mer.p); }
    const uint64_t *t_fp() const { return (const uint64_t *) (adopted ? afp.p : fp.p); }
    uint64_t t_n() const { return adopted ? n_adopted : n_syncmers; }
};

namespace sg {
int ensure_adopted_soa(sg_batch *b);
int launch_pack(sg_batch *b, cudaStream_t st, bool want_hs, bool want_rl);
}
