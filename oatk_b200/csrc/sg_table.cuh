// sg_table.cuh -- warp-cooperative open-addressing hash table (64-bit keys, 32-bit counters).
//
// Replaces the khashl maps the reference counts with (arc tally: kh_u128 in syncasm.c:240-261;
// multiplicity tables: kh_ctab in syncmer.c:558-667 fed from two qsorts). A warp first merges
// equal keys among its 32 lanes with __match_any_sync, then takes the distinct keys one at a
// time and probes 32 consecutive slots at once: one coalesced 256-byte read, __ballot_sync to
// find the key or the first empty slot, one atomicCAS by the elected lane, one atomicAdd.
#pragma once
#include "sg_common.cuh"

namespace sg {

constexpr uint64_t EMPTY_KEY = ~0ull;

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// add `cnt` to the counter of `key`; whole warp, key and cnt uniform across lanes
__device__ __forceinline__ void table_add(uint64_t *tk, uint32_t *tv, uint64_t nslot_mask, uint64_t key, uint32_t cnt, int lane)
{
    uint64_t g = (mix64(key) << 5) & nslot_mask;           // group of 32 slots
    for (;;) {
        const uint64_t cur = *((volatile uint64_t *) (tk + g + lane));
        const uint32_t hit = __ballot_sync(SG_FULL, cur == key);
        if (hit) {
            if (lane == __ffs(hit) - 1) atomicAdd(tv + g + lane, cnt);
            return;
        }
        const uint32_t emp = __ballot_sync(SG_FULL, cur == EMPTY_KEY);
        if (emp) {
            const int e = __ffs(emp) - 1;
            uint64_t old = 0;
            if (lane == e) old = atomicCAS((unsigned long long *) (tk + g + lane), (unsigned long long) EMPTY_KEY, (unsigned long long) key);
            old = __shfl_sync(SG_FULL, old, e);
            if (old == EMPTY_KEY || old == key) {
                if (lane == e) atomicAdd(tv + g + lane, cnt);
                return;
            }
            continue;                                      // somebody else took the slot: look at this group again
        }
        g = (g + 32) & nslot_mask;
    }
}


// One lane, one key, in a table that is mostly empty and resident in L2: plain linear probing, a slot at a time. Every lane of a
// warp is on its own probe sequence, so 32 reads are in flight where the group probe above has one (the s-mer tally of sg_stat
// is bound by that latency, not by bytes). false: no slot within max_probes.
__device__ __forceinline__ bool slot_add_bounded(uint64_t *tk, uint32_t *tv, uint64_t nslot_mask, uint64_t key, uint32_t cnt, int max_probes)
{
    uint64_t g = mix64(key) & nslot_mask;
    for (int t = 0; t < max_probes; ++t) {
        uint64_t cur = *((volatile uint64_t *) (tk + g));
        if (cur == EMPTY_KEY) {
            cur = atomicCAS((unsigned long long *) (tk + g), (unsigned long long) EMPTY_KEY, (unsigned long long) key);
            if (cur == EMPTY_KEY) cur = key;
        }
        if (cur == key) { atomicAdd(tv + g, cnt); return true; }
        g = (g + 1) & nslot_mask;
    }
    return false;
}

// one key per lane (EMPTY_KEY = nothing): add 1 for every lane's key; the whole warp must call
__device__ __forceinline__ void table_add_warp(uint64_t *tk, uint32_t *tv, uint64_t nslot_mask, uint64_t key, int lane)
{
    const uint32_t peers = __match_any_sync(SG_FULL, key);
    const bool leader = key != EMPTY_KEY && lane == __ffs(peers) - 1;
    const uint32_t cnt = __popc(peers);
    uint32_t work = __ballot_sync(SG_FULL, leader);
    while (work) {
        const int src = __ffs(work) - 1;
        work &= work - 1;
        table_add(tk, tv, nslot_mask, __shfl_sync(SG_FULL, key, src), __shfl_sync(SG_FULL, cnt, src), lane);
    }
}

} // namespace sg
