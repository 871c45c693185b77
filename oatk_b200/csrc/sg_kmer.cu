// sg_kmer.cu -- kernel 1c: MurmurHash64A of every selected k-mer, and placement of
// the syncmer records into read order.
//
// Replaces kmer_hash64 + MurmurHash64A (reference syncmer.c:175-226, 131-170):
// the k bases starting at hoco position `start` are taken in read orientation
// (rev = 0) or reverse-complemented (rev = 1), packed 4 per byte first base on
// top, the unused low bits of the last byte are zero, and ceil(k/4) bytes are
// hashed with seed 1234 as little-endian 8-byte blocks plus a <= 7-byte tail.
//
// Here one thread hashes one k-mer straight from the packed read: 32 bases at a
// time are pulled out of the big-endian word stream with a funnel shift (the
// reverse strand through a 2-bit-group reversal of the mirrored window), byte
// swapped into the little-endian block Murmur expects, and mixed. The hash
// chain is serial by construction (xor and multiply do not commute), so the
// parallelism is across k-mers.
#include "sg_common.cuh"
#include "sg_internal.h"

namespace sg {

// Besides the reference's MurmurHash64A, a second, independent 64-bit hash of the same oriented
// words (multiply-xorshift chain with other constants, over the big-endian blocks). Hash and
// fingerprint together decide "same k-mer" in sg_count's default mode; see DESIGN.md section 8.
// MurmurHash64A over the oriented k-mer with a three-word window that slides by two words per block (each block costs two
// loads and no re-conversion of the word it shares with its neighbour). Every load is bounds-checked like hoco_word: the first and the last syncmer of
// a read reach past its words, and with 21 syncmers per read nearly every warp holds one of them -- a separate unchecked
// path for the others made those warps run both.
__device__ __forceinline__ uint64_t kmer_murmur_sliding(const uint32_t *hs, int64_t nwords, int64_t start, int k, int rev, uint64_t *fp_out)
{
    const uint64_t M = 0xc6a4a7935bd1e995ull, F = 0x9e3779b97f4a7c15ull;
    const uint32_t nbytes = (uint32_t) (k + 3) >> 2, nblk = nbytes >> 3;
    uint64_t h = 1234ull ^ ((uint64_t) nbytes * M), f = 0x243f6a8885a308d3ull ^ (uint64_t) k;
    const int64_t p0 = rev ? start + k - 32 : start;                // first window; later ones are 32 bases further on / back
    const int sh = (int) (p0 & 15) * 2;
    int64_t wi = p0 >> 4;                                           // arithmetic shift: floor, the last reverse window may start before the read
    uint32_t a = hoco_word(hs, wi, nwords), b = hoco_word(hs, wi + 1, nwords), c = hoco_word(hs, wi + 2, nwords);
    const uint32_t nall = nblk + ((nbytes & 7u) ? 1u : 0u);
    // the strands of the k-mers of a warp are a coin toss each: both directions go through the same instructions
    // (selects, no branches), or every iteration would run twice
    const int step = rev ? -2 : 2, o1 = rev ? 0 : 1;
    for (uint32_t j = 0; j < nall; ++j) {
        uint64_t be = (uint64_t) __funnelshift_l(b, a, sh) << 32 | __funnelshift_l(c, b, sh);
        const uint64_t rcbe = rc64(be);
        be = rev ? rcbe : be;
        // next window: two words on (forward) or two words back (reverse); one word is shared
        if (j + 1 < nall) {
            wi += step;
            const uint32_t x = hoco_word(hs, wi + o1, nwords), y = hoco_word(hs, wi + o1 + 1, nwords);
            const uint32_t na = rev ? x : c, nb = rev ? y : x, nc = rev ? a : y;
            a = na; b = nb; c = nc;
        }
        const int left = k - 32 * (int) j;                            // bases of the k-mer in this block
        if (left < 32) be &= ~0ull << (64 - 2 * left);
        if (j < nblk) {
            uint64_t x = bswap64(be);
            x *= M; x ^= x >> 47; x *= M;
            h = (h ^ x) * M;
        } else h = (h ^ bswap64(be)) * M;                             // the <= 7-byte tail is mixed in unhashed
        f = (f ^ be) * F; f ^= f >> 32;
    }
    h ^= h >> 47; h *= M; h ^= h >> 47;
    f *= 0xd6e8feb86659fd93ull; f ^= f >> 29;
    *fp_out = f;
    return h;
}

__global__ void __launch_bounds__(256) kmerhash_kernel(KmerArgs A)
{
    const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_rec) return;
    const uint32_t sid = A.rec_sid[i], idx = A.rec_idx[i], rec = A.rec_mpos[i];
    const uint64_t hb = A.hoff[sid];
    const uint32_t *hs32 = reinterpret_cast<const uint32_t *>(A.hoco_s + hb / 4);
    const int64_t nwords = ((int64_t) A.hoco_l[sid] + 15) >> 4;
    // the scan kernel says where the k-mer starts and which rule selected it. OPEN: code of the first s-mer;
    // CLOSE: code of the last s-mer with its low bit flipped (reference syncmer.c:325-338, 356-377); the strand
    // bit of that s-mer orients the k-mer
    const int64_t t = rec >> 1;
    const bool open = rec & 1u;
    const uint64_t raw = smer_code_at(hs32, open ? t + A.s - 1 : t + A.k - 1, A.s, nwords);
    const uint32_t mp = (uint32_t) t << 1 | (uint32_t) (raw & 1ull);
    uint64_t fp;
    const uint64_t h = kmer_murmur_sliding(hs32, nwords, t, A.k, (int) (mp & 1u), &fp);
    const uint64_t o = A.scm_off[sid] + idx;
    A.key[o] = h;
    A.fp[o] = fp;
    A.occ[o] = (A.sid_base + sid) << 32 | (uint64_t) idx << 1 | (mp & 1u);
    A.m_pos[o] = mp;
    A.s_mer[o] = open ? raw : raw ^ 1ull;
    // the same three words once more as one 32-byte record: the hash-order gather of sg_count then touches one
    // sector per tuple instead of three
    if (A.tup) {
        ulonglong4 t;
        t.x = (A.sid_base + sid) << 32 | (uint64_t) idx << 1 | (mp & 1u); t.y = open ? raw : raw ^ 1ull; t.z = fp; t.w = h;
        reinterpret_cast<ulonglong4 *>(A.tup)[o] = t;
    }
}

int launch_kmerhash(const KmerArgs &A, cudaStream_t st)
{
    if (A.n_rec == 0) return 0;
    kmerhash_kernel<<<(unsigned) ((A.n_rec + 255) / 256), 256, 0, st>>>(A);
    return 1;
}

} // namespace sg
