// sg_internal.h -- kernel argument blocks and launchers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "sg_hash31.cuh"

namespace sg {

struct EncodeArgs {
    const uint8_t *bases;        // all reads back to back
    const uint64_t *off;         // n_reads + 1 raw offsets
    const uint64_t *hoff;        // n_reads + 1 capacity offsets (multiples of 64)
    uint64_t n_reads;
    unsigned int *work;          // zeroed before the launch: next read to hand to a warp
    uint8_t *hoco_s, *ho_rl, *nbits;
    uint32_t *hoco_l;            // per read
    uint32_t *n_amb;             // per read: ambiguous characters
    unsigned long long *amb_count, *lrl_count;
    uint64_t amb_cap, lrl_cap;
    uint32_t *amb_sid, *amb_pos;
    uint32_t *lrl_sid, *lrl_idx, *lrl_val;
};
int launch_encode(const EncodeArgs &A, cudaStream_t st);

struct ScanArgs {
    const uint64_t *hoff;
    const uint8_t *hoco_s, *nbits;
    const uint32_t *hoco_l, *n_amb;
    int k, s;
    uint64_t n_reads;
    unsigned int *work;          // zeroed before the launch: next read to hand to a warp
    uint32_t *n_scm;             // per read: syncmers emitted
    unsigned long long *rec_count;
    uint64_t rec_cap;
    uint32_t *rec_sid, *rec_idx, *rec_mpos;   // unordered records (one per syncmer): read, rank on the read, start << 1 | open
    // reads handed over to scan_exact_kernel (see sg_scan.cu): (read, first tile, records so far, CLOSE carry) each
    unsigned int *defer_count, *work2;        // zeroed before the launch
    uint32_t *defer;                          // 4 * n_reads
    uint64_t *xring;                          // scan_exact_scratch_bytes(k, s)
};
size_t scan_exact_scratch_bytes(int k, int s, int *grid_out);
constexpr int SYNC_SCAN_WARPS = 4;  // warps per CTA of the syncmer scan kernel (one read per warp, 16 positions per lane and tile)
struct ScanGeom {
    int rch;        // ring size in chunks of 16 positions (power of two, >= window + one tile)
    int n_full;     // chunks fully inside every window of a lane's 16 positions
    int logB;       // chunk minima are prefix/suffix-min scanned in blocks of 2^logB lanes (2^logB <= n_full)
    H31Consts h31;  // shift multipliers of the s = 31 hash (kept as parameters: see sg_hash31.cuh)
};
int scan_geometry(int k, int s, ScanGeom *g, size_t *smem_per_warp);
// returns launches (>= 0) or a negative SG_E_* code
int launch_scan(const ScanArgs &A, cudaStream_t st);

struct KmerArgs {
    const uint64_t *hoff;
    const uint8_t *hoco_s;
    const uint32_t *hoco_l;
    int k, s;
    uint64_t n_rec;
    const uint32_t *rec_sid, *rec_idx, *rec_mpos;
    const uint64_t *scm_off;     // n_reads + 1: exclusive scan of n_scm
    uint64_t sid_base;
    // ordered outputs (read order)
    uint64_t *key;               // k-mer hash
    uint64_t *fp;                // independent second hash of the same k-mer
    uint64_t *tup;               // optional: (occ, s_mer, fp, key) per syncmer as one 32-byte record
    uint64_t *occ;               // sid << 32 | idx << 1 | rev
    uint32_t *m_pos;
    uint64_t *s_mer;
};
int launch_kmerhash(const KmerArgs &A, cudaStream_t st);

// ---- primitives (sg_prims.cu) ----
// exclusive scan of n uint32 values into n+1 uint64 (out[n] = total); tmp must hold scan_tmp_words(n) uint64
size_t scan_tmp_words(uint64_t n);
int launch_scan_u32_u64(const uint32_t *in, uint64_t *out, uint64_t n, uint64_t *tmp, cudaStream_t st);
// hoff[r] = sum_{j<r} roundup64(off[j+1]-off[j]); same tmp requirement
int launch_capacity_offsets(const uint64_t *off, uint64_t *hoff, uint64_t n, uint64_t *tmp, cudaStream_t st);
// stable LSD radix sort of (key, val) pairs by the full 64-bit key; *_alt are ping-pong buffers.
// On return the sorted data is in key/val. tmp must hold sort_tmp_words(n) uint32.
size_t sort_tmp_words(uint64_t n);
int launch_sort_pairs(uint64_t *key, uint64_t *val, uint64_t *key_alt, uint64_t *val_alt, uint64_t n,
        int begin_bit, int end_bit, uint32_t *tmp, cudaStream_t st);
int launch_exscan_u32(uint32_t *a, uint64_t n, uint32_t *sums, cudaStream_t st);
// the same for bare 64-bit words (a payload can ride in the bits below begin_bit); result in key
int launch_sort_keys(uint64_t *key, uint64_t *key_alt, uint64_t n, int begin_bit, int end_bit, uint32_t *tmp, cudaStream_t st);

} // namespace sg
