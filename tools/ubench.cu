// Integer-pipe micro-benchmarks for sm_100a (development aid; results quoted in DESIGN.md).
// Each kernel issues a long stream of one instruction mix; we report warp-instructions per
// cycle per SM sub-partition derived from clock64().
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048
#define REP8(x) x x x x x x x x

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t two, uint32_t three, long long* cyc)
{
    uint32_t a0 = threadIdx.x, a1 = a0 * 3 + 1, a2 = a0 ^ 0x5555, a3 = a0 + 77, a4 = a0 * 5, a5 = a0 | 9, a6 = a0 + 1234, a7 = ~a0;
    uint32_t b = blockIdx.x + 12345, c = threadIdx.x * 7 + 3;
    long long t0 = clock64();
    for (int i = 0; i < ITER; ++i) {
        if (MODE == 0) {   // LOP3 only (8 independent chains)
            REP8(asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b), "r"(c));
                 asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a1) : "r"(b), "r"(c));
                 asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(b), "r"(c));
                 asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a3) : "r"(b), "r"(c));)
        } else if (MODE == 1) {   // IMAD reg form
            REP8(asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a0) : "r"(three), "r"(c));
                 asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a1) : "r"(three), "r"(c));
                 asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a2) : "r"(three), "r"(c));
                 asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a3) : "r"(three), "r"(c));)
        } else if (MODE == 2) {   // IMAD.HI reg form
            REP8(asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a0) : "r"(three), "r"(c));
                 asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a1) : "r"(three), "r"(c));
                 asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a2) : "r"(three), "r"(c));
                 asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a3) : "r"(three), "r"(c));)
        } else if (MODE == 3) {   // 1 LOP3 : 1 IMAD
            REP8(asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b), "r"(c));
                 asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a1) : "r"(three), "r"(c));
                 asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(b), "r"(c));
                 asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a3) : "r"(three), "r"(c));)
        } else if (MODE == 4) {   // 2 LOP3 : 1 IMAD : 1 IMAD.HI  (variant B mix)
            REP8(asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b), "r"(c));
                 asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a1) : "r"(three), "r"(c));
                 asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(b), "r"(c));
                 asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a3) : "r"(two), "r"(c));)
        } else if (MODE == 5) {   // IADD3 (independent)
            REP8(asm volatile("add.u32 %0, %0, %1;" : "+r"(a0) : "r"(b));
                 asm volatile("add.u32 %0, %0, %1;" : "+r"(a1) : "r"(c));
                 asm volatile("add.u32 %0, %0, %1;" : "+r"(a2) : "r"(b));
                 asm volatile("add.u32 %0, %0, %1;" : "+r"(a3) : "r"(c));)
        } else if (MODE == 6) {   // mad.wide
            unsigned long long w0 = a0, w1 = a1, w2 = a2, w3 = a3;
            REP8(asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w0) : "r"(a4), "r"(three));
                 asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w1) : "r"(a5), "r"(three));
                 asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w2) : "r"(a6), "r"(three));
                 asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w3) : "r"(a7), "r"(three));)
            a0 = (uint32_t)w0 ^ (uint32_t)(w0 >> 32); a1 = (uint32_t)w1; a2 = (uint32_t)w2; a3 = (uint32_t)w3;
        } else if (MODE == 7) {   // carry chain add.cc/addc.cc (what the tile kernel uses)
            REP8(asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(a0) : "r"(b));
                 asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(a1) : "r"(c));
                 asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(a2) : "r"(b));
                 asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(a3) : "r"(c));)
        } else if (MODE == 8) {   // 3 LOP3 : 1 IMAD
            REP8(asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b), "r"(c));
                 asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a1) : "r"(b), "r"(c));
                 asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(b), "r"(c));
                 asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a3) : "r"(three), "r"(c));)
        } else if (MODE == 9) {   // IMAD with immediate multiplier 3 (imm form)
            REP8(asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(a0) : "r"(c));
                 asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(a1) : "r"(c));
                 asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(a2) : "r"(c));
                 asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(a3) : "r"(c));)
        } else if (MODE == 10) {   // popc
            REP8(asm volatile("popc.b32 %0, %0;" : "+r"(a0));
                 asm volatile("popc.b32 %0, %0;" : "+r"(a1));
                 asm volatile("popc.b32 %0, %0;" : "+r"(a2));
                 asm volatile("popc.b32 %0, %0;" : "+r"(a3));)
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, uint32_t* out, long long* cyc, int warps_per_smsp)
{
    const int blocks_per_sm = warps_per_smsp * 4 * 32 / 256;
    const int blocks = 148 * (blocks_per_sm < 1 ? 1 : blocks_per_sm);
    k<MODE><<<blocks, 256>>>(out, 2, 3, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 2, 3, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148 * 8]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    const double instr_per_warp = (double)ITER * 32;
    const double warps_per_smsp_d = (double)blocks / 148 * 8 / 4;
    printf("%-28s warps/SMSP=%2d  cycles=%9.0f  warp-instr/clk/SMSP=%.3f  (%.3f ms)\n", name, warps_per_smsp, avg,
           instr_per_warp * warps_per_smsp_d / avg, ms);
}

int main()
{
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&cyc, 148 * 8 * 8);
    for (int w : {2, 8}) {
        run<0>("LOP3", out, cyc, w);
        run<5>("IADD3", out, cyc, w);
        run<7>("IADD3.X carry chain", out, cyc, w);
        run<1>("IMAD (reg)", out, cyc, w);
        run<9>("IMAD (imm)", out, cyc, w);
        run<2>("IMAD.HI (reg)", out, cyc, w);
        run<6>("IMAD.WIDE", out, cyc, w);
        run<10>("POPC", out, cyc, w);
        run<3>("1 LOP3 : 1 IMAD", out, cyc, w);
        run<8>("3 LOP3 : 1 IMAD", out, cyc, w);
        run<4>("2 LOP3 : 1 IMAD : 1 IMAD.HI", out, cyc, w);
    }
    return 0;
}
