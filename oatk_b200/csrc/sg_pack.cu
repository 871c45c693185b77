// sg_pack.cu -- gathers the capacity-indexed hoco_s / ho_rl arrays into the compact,
// 16-byte aligned per-read layout that sg_extract_download hands to the host
// (one device-to-host copy per array instead of one per read).
#include "sg_common.cuh"
#include "sg_internal.h"
#include "sg_host.h"

namespace sg {

__device__ __forceinline__ uint4 mask_tail(uint4 v, int64_t keep)     // keep the first `keep` bytes of 16
{
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t left = keep - 4 * j;
        if (left <= 0) w[j] = 0;
        else if (left < 4) w[j] &= (1u << (8 * left)) - 1u;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(128) pack_kernel(const uint64_t *hoff, const uint32_t *hoco_l,
        const uint8_t *hoco_s, const uint8_t *ho_rl, const uint64_t *hs_off, const uint64_t *rl_off,
        uint8_t *pk_hs, uint8_t *pk_rl)
{
    const uint64_t r = blockIdx.x;
    const int64_t L = hoco_l[r], hsb = (L + 3) >> 2;
    const uint64_t hb = hoff[r];
    const uint4 *src_s = reinterpret_cast<const uint4 *>(hoco_s + hb / 4);
    const uint4 *src_r = reinterpret_cast<const uint4 *>(ho_rl + hb);
    uint4 *dst_s = reinterpret_cast<uint4 *>(pk_hs + hs_off[r]);
    uint4 *dst_r = reinterpret_cast<uint4 *>(pk_rl + rl_off[r]);
    const int64_t ns = (hsb + 15) >> 4, nr = (L + 15) >> 4;
    if (pk_hs) for (int64_t i = threadIdx.x; i < ns; i += blockDim.x) dst_s[i] = mask_tail(src_s[i], hsb - 16 * i);
    if (pk_rl) for (int64_t i = threadIdx.x; i < nr; i += blockDim.x) dst_r[i] = mask_tail(src_r[i], L - 16 * i);
}

int launch_pack(sg_batch *b, cudaStream_t st, bool want_hs, bool want_rl)
{
    if (b->n_reads == 0) return 0;
    pack_kernel<<<(unsigned) b->n_reads, 128, 0, st>>>((const uint64_t *) b->hoff.p, (const uint32_t *) b->hoco_l.p,
            (const uint8_t *) b->hoco_s.p, (const uint8_t *) b->ho_rl.p,
            (const uint64_t *) b->pk_hs_off.p, (const uint64_t *) b->pk_rl_off.p,
            want_hs ? (uint8_t *) b->pk_hs.p : nullptr, want_rl ? (uint8_t *) b->pk_rl.p : nullptr);
    return 1;
}

} // namespace sg
