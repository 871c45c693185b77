#if VARIANT >= 10
#include "../../oatk_b200/csrc/sg_hash31.cuh"
#else
#include "hash31.cuh"
#endif
#include <cstdio>
#include <vector>
using namespace sg;
__device__ __forceinline__ uint32_t rev2(uint32_t x){ x = __brev(x); return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1); }
#if VARIANT == 9 || VARIANT == 12
__device__ __forceinline__ uint64_t hash64(uint64_t x, uint64_t mask)
{
    x = ((x << 21) - x - 1) & mask; x ^= x >> 24; x = (x * 265) & mask; x ^= x >> 14; x = (x * 21) & mask; x ^= x >> 28; x = (x + (x << 31)) & mask; return x;
}
#endif
#if VARIANT >= 10
template <int J> __device__ __forceinline__ void one(uint32_t a, uint32_t b, uint32_t w0, uint32_t ra, uint32_t rb, uint32_t rc, uint32_t *dst, int RCH, uint32_t &cmin, const H31Consts &K)
{
    uint32_t hi, lo;
    h31_canon<J>(a, b, w0, ra, rb, rc, hi, lo);
#if VARIANT == 12
    uint32_t hv = (uint32_t) (hash64(((uint64_t) hi << 32 | lo) >> 2, (1ull << 62) - 1) >> 32);
#else
    uint32_t hv = min(h31_hash_top(hi, lo, K), 0xfffffffeu);
#endif
    dst[J * RCH] = hv; cmin = min(cmin, hv);
}
#else
template <int J> __device__ __forceinline__ void one(uint32_t a, uint32_t b, uint32_t w0, uint32_t ra, uint32_t rb, uint32_t rc, uint32_t *dst, int RCH, uint32_t &cmin, const H31Consts &K)
{
    uint32_t hi, lo; bool pal;
    canon31<J>(a, b, w0, ra, rb, rc, hi, lo, pal);
    uint32_t hv = min(hash31_hi(hi, lo, K), 0xfffffffeu);
    dst[J * RCH] = hv; cmin = min(cmin, hv);
}
#endif
__global__ void __launch_bounds__(64) tk(const uint32_t *in, uint32_t *out, H31Consts K, int n)
{
    __shared__ uint32_t ring[16 * 128]; __shared__ uint64_t s_m4; 
#if VARIANT < 10
 if (threadIdx.x == 0) s_m4 = K.m4;
#endif
 __syncthreads(); 
#if VARIANT < 10
 K.m4 = *(volatile uint64_t *) &s_m4;
#endif

    uint32_t acc = 0xffffffffu;
    for (int it = 0; it < n; ++it) {
        int c = ((it * 64 + threadIdx.x) & 1023) + 2;
        uint32_t a = __ldg(in + c - 2), b = __ldg(in + c - 1), w0 = __ldg(in + c);
        uint32_t *dst = ring + threadIdx.x + (it & 1) * 64; uint32_t cmin = 0xffffffffu;
#if VARIANT == 11
        {
            uint64_t V = (uint64_t) a << 32 | b;
            uint64_t fw = V << 2, rv = ((uint64_t) rev2(~(uint32_t) V) << 32 | rev2(~(uint32_t) (V >> 32))) & ~3ull;
            for (int i4 = 0; i4 < 4; ++i4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t bb = w0 >> 30; w0 <<= 2;
                    fw = (fw << 2) | (bb << 2); rv = ((rv >> 2) & ~3ull) | ((uint64_t) (3u - bb) << 62);
                    const uint64_t x = fw < rv ? fw : rv;
                    const uint32_t hv = min(h31_hash_top((uint32_t) (x >> 32), (uint32_t) x, K), 0xfffffffeu);
                    dst[j * 128] = hv; cmin = min(cmin, hv);
                }
                dst += 4 * 128;
            }
        }
#elif VARIANT == 9
        const uint64_t mask = (1ull << 62) - 1;
        uint64_t V = (uint64_t) a << 32 | b;
        uint64_t fw = V & mask, rv = ((uint64_t) rev2(~(uint32_t) V) << 32 | rev2(~(uint32_t) (V >> 32))) >> 2;
        for (int i4 = 0; i4 < 4; ++i4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t bb = w0 >> 30; w0 <<= 2;
                fw = ((fw << 2) | bb) & mask; rv = (rv >> 2) | ((uint64_t) (3u - bb) << 60);
                const uint32_t hv = (uint32_t) (hash64(fw < rv ? fw : rv, mask) >> 32);
                dst[j * 128] = hv; cmin = min(cmin, hv);
            }
            dst += 4 * 128;
        }
#else
        uint32_t ra = rev2(~w0), rb = rev2(~b), rc = rev2(~a);
        one<0>(a,b,w0,ra,rb,rc,dst,128,cmin,K); one<1>(a,b,w0,ra,rb,rc,dst,128,cmin,K); one<2>(a,b,w0,ra,rb,rc,dst,128,cmin,K); one<3>(a,b,w0,ra,rb,rc,dst,128,cmin,K);
        one<4>(a,b,w0,ra,rb,rc,dst,128,cmin,K); one<5>(a,b,w0,ra,rb,rc,dst,128,cmin,K); one<6>(a,b,w0,ra,rb,rc,dst,128,cmin,K); one<7>(a,b,w0,ra,rb,rc,dst,128,cmin,K);
        one<8>(a,b,w0,ra,rb,rc,dst,128,cmin,K); one<9>(a,b,w0,ra,rb,rc,dst,128,cmin,K); one<10>(a,b,w0,ra,rb,rc,dst,128,cmin,K); one<11>(a,b,w0,ra,rb,rc,dst,128,cmin,K);
        one<12>(a,b,w0,ra,rb,rc,dst,128,cmin,K); one<13>(a,b,w0,ra,rb,rc,dst,128,cmin,K); one<14>(a,b,w0,ra,rb,rc,dst,128,cmin,K); one<15>(a,b,w0,ra,rb,rc,dst,128,cmin,K);
#endif
        acc = min(acc, cmin);
        __syncthreads();
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + ring[(threadIdx.x * 7) & 1023];
}
int main()
{
    const int nblk = 148 * 16 * 4, n = 256;
    uint32_t *in, *out;
    cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, (size_t) nblk * 64 * 4);
    std::vector<uint32_t> h(4096); uint32_t x = 12345; for (auto &v : h) { x = x * 1664525u + 1013904223u; v = x; }
    cudaMemcpy(in, h.data(), 4096 * 4, cudaMemcpyHostToDevice);
#if VARIANT >= 10
    H31Consts K = h31_consts();
#else
    H31Consts K = {1u << 8, 1u << 18, 1u << 4, 0, 0xfffffffffffffffcull};
#endif
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) tk<<<nblk, 64>>>(in, out, K, n);
    cudaEventRecord(e0);
    for (int w = 0; w < 5; ++w) tk<<<nblk, 64>>>(in, out, K, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    std::vector<uint32_t> o((size_t) nblk * 64); cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    uint64_t cs = 0; for (auto v : o) cs = cs * 31 + v;
    double pos = (double) nblk * 64 * 16 * n;
    printf("variant %d shifts %d funnels %d: %.3f ms, %.1f Gpos/s, %.1f ms per 11.25 Gpos, checksum %llx, err %d\n", VARIANT, H31_FMA_SHIFTS, H31_FMA_FUNNELS, ms, pos / ms * 1e-6, 11.25e9 / (pos / ms), (unsigned long long) cs, (int) cudaGetLastError());
    return 0;
}
