// int_pipes.cu -- per-SM issue rates of the integer instructions the syncmer kernels are made of.
// Each kernel runs 8 independent dependency chains per thread; results are warp-instructions per
// clock per SM sub-partition (SMSP), measured with 1024 threads per SM-resident CTA set.
#include <cstdio>
#include <cstdint>
#define CHAINS 8
#define REPS 64
#define ITERS 256
template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t p, uint32_t q)
{
    uint32_t x[CHAINS], y[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { x[i] = threadIdx.x * 7 + i + p; y[i] = blockIdx.x + i * 3 + q; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < REPS / 2; ++r) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                if (OP == 0) { asm volatile("shf.r.clamp.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(y[i])); asm volatile("shf.r.clamp.b32 %0, %0, %1, 9;" : "+r"(y[i]) : "r"(x[i])); }
                if (OP == 1) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(p)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(y[i]) : "r"(x[i]), "r"(q)); }
                if (OP == 2) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(p), "r"(y[i])); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(q), "r"(x[i])); }
                if (OP == 3) { asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(p)); asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(q)); }
                if (OP == 4) { uint64_t w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x[i]), "r"(p)); x[i] = (uint32_t) w; y[i] ^= (uint32_t) (w >> 32);
                               asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(y[i]), "r"(q)); y[i] = (uint32_t) w; x[i] ^= (uint32_t) (w >> 32); }
                if (OP == 5) { asm volatile("shf.r.clamp.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(p)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(q), "r"(p)); }
                if (OP == 6) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(q), "r"(p)); asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(q)); }
                if (OP == 7) { asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(x[i])); }
                if (OP == 8) { asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(p)); asm volatile("max.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(q)); }
                if (OP == 9) { asm volatile("{.reg .pred pp; setp.lt.u32 pp, %0, %1; selp.u32 %0, %1, %2, pp;}" : "+r"(x[i]) : "r"(y[i]), "r"(p)); asm volatile("{.reg .pred pp; setp.lt.u32 pp, %0, %1; selp.u32 %0, %1, %2, pp;}" : "+r"(y[i]) : "r"(x[i]), "r"(q)); }
                if (OP == 10) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(q), "r"(p)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(q), "r"(p)); }
                if (OP == 11) { float fx = __uint_as_float(x[i]), fy = __uint_as_float(y[i]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(fx) : "f"(1.0001f), "f"(0.5f)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(fy) : "f"(0.9999f), "f"(0.25f)); x[i] = __float_as_uint(fx); y[i] = __float_as_uint(fy); }
                if (OP == 13) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(q), "r"(p)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(x[i]) : "r"(p), "r"(q)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(q), "r"(p)); asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(q)); }
                if (OP == 14) { uint64_t w; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(q), "r"(p)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(x[i]) : "r"(p), "r"(q)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(q), "r"(p)); asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(y[i]), "r"(q)); y[i] = (uint32_t) (w >> 32); }
                if (OP == 15) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(q), "r"(p)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(x[i]) : "r"(p), "r"(q)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(q), "r"(p)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(p), "r"(q)); }
                if (OP == 16) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(q), "r"(p)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(x[i]) : "r"(p), "r"(q)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(q), "r"(p)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(q), "r"(p)); }
                if (OP == 17) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(p)); asm volatile("shf.r.clamp.b32 %0, %0, %1, 7;" : "+r"(y[i]) : "r"(x[i])); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(p)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(q), "r"(x[i])); }
                if (OP == 12) { float fy = __uint_as_float(y[i]); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(q), "r"(p)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(fy) : "f"(0.9999f), "f"(0.25f)); y[i] = __float_as_uint(fy); }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char *name, uint32_t *out, int instr_per_pair)
{
    const int nblk = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<nblk, 256>>>(out, 3, 5);
    cudaEventRecord(e0);
    k<OP><<<nblk, 256>>>(out, 3, 5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winstr = (double) nblk * 8 /*warps*/ * ITERS * (REPS / 2) * CHAINS * instr_per_pair;
    double cyc = ms * 1e-3 * 1.965e9 * 148 * 4;
    printf("%-28s %.3f ms  %.3f warp-instr/clk/SMSP (at 1965 MHz)\n", name, ms, winstr / cyc);
}
int main()
{
    uint32_t *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("SHF", out, 2); run<1>("LOP3", out, 2); run<7>("IADD", out, 2); run<8>("VIMNMX", out, 2); run<9>("ISETP+SEL", out, 4);
    run<2>("IMAD", out, 2); run<3>("IMAD.HI", out, 2); run<4>("IMAD.WIDE (+LOP3)", out, 4);
    run<5>("SHF + IMAD", out, 2); run<10>("LOP3 + IMAD", out, 2); run<6>("LOP3 + IMAD.HI", out, 2);
    run<13>("3 LOP3 + 1 IMAD.HI", out, 4); run<14>("3 LOP3 + 1 IMAD.WIDE", out, 4); run<15>("2 LOP3 + 2 IMAD", out, 4); run<16>("3 LOP3 + 1 IMAD", out, 4); run<17>("LOP3 SHF LOP3 IMAD dependent regs", out, 4);
    run<11>("FFMA", out, 2); run<12>("LOP3 + FFMA", out, 2);
    printf("err %d\n", (int) cudaGetLastError());
    return 0;
}
