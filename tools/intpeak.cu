// Integer-pipe issue-rate microbenchmark for sm_100a.
// Measures warp-instructions per clock per SM for the instruction classes the Goldilocks/Poseidon
// kernels are made of, so the Poseidon roofline denominator is measured, not assumed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/intpeak tools/intpeak.cu
// Output: one JSON object on stdout.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8   // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(1024) bench(uint32_t* out, uint32_t seed, unsigned long long* clk) {
    uint32_t a[UNROLL], b = seed | 1u, c = seed * 3u + 7u;
    unsigned long long w[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; i++) { a[i] = threadIdx.x + i * seed; w[i] = a[i] * 0x9e3779b97f4a7c15ULL; }
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < UNROLL; i++) {
            if (MODE == 0) {          // IMAD (32-bit lo)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            } else if (MODE == 1) {   // IMAD.WIDE.U32 with 64-bit accumulate
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b));
            } else if (MODE == 2) {   // IADD3
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            } else if (MODE == 3) {   // 64-bit add = IADD3 + IADD3.X
                asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((unsigned long long)b << 13 | c));
            } else if (MODE == 4) {   // 1:1 IMAD.WIDE + IADD3
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));
            } else if (MODE == 5) {   // IMAD.HI.U32
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            } else if (MODE == 6) {   // LOP3
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            } else if (MODE == 7) {   // SHF (funnel shift)
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b));
            } else if (MODE == 8) {   // 1:1 IMAD + IADD3
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));
            } else if (MODE == 9) {   // 1:2 IMAD.WIDE + IADD3 pair (64-bit add)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            } else if (MODE == 10) {  // ISETP + SEL (compare-select)
                asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(a[i]) : "r"(b), "r"(c));
            } else if (MODE == 11) {  // IMAD.WIDE with small immediate multiplier
                asm volatile("mad.wide.u32 %0, %1, 41, %0;" : "+l"(w[i]) : "r"(a[i]));
            } else if (MODE == 12) {  // mad.lo.cc + madc.hi (expected: IMAD.WIDE.U32 with 64-bit addend)
                asm volatile("{ .reg .u32 l, h; mov.b64 {l,h}, %0; mad.lo.cc.u32 l, %1, %2, l; madc.hi.u32 h, %1, %2, h; mov.b64 %0, {l,h}; }" : "+l"(w[i]) : "r"(a[i]), "r"(b));
            } else if (MODE == 13) {  // mul.wide only (IMAD.WIDE.U32 with RZ addend), result xor-folded by LOP3
                unsigned long long t;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[i]), "r"(b));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"((uint32_t)t), "r"((uint32_t)(t >> 32)));
            } else if (MODE == 14) {  // 2x mul.wide + 1 LOP3
                unsigned long long t, u;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[i]), "r"(b));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(u) : "r"(a[i]), "r"(c));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[i]) : "r"((uint32_t)t), "r"((uint32_t)(t >> 32)), "r"((uint32_t)u));
            } else if (MODE == 16) {  // IMAD.WIDE.U32 (RZ addend), result never consumed in the loop
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(b));
            } else if (MODE == 17) {  // one accumulating IMAD.WIDE per chain per iteration (mad.lo.cc/madc.hi)
                uint32_t l = (uint32_t)w[i], h = (uint32_t)(w[i] >> 32);
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(l), "+r"(h) : "r"(a[i]), "r"(b));
                w[i] = ((unsigned long long)h << 32) | l;
            } else if (MODE == 18) {  // accumulating IMAD.WIDE with carry-out + carry counter (lazy dot product step)
                uint32_t l = (uint32_t)w[i], h = (uint32_t)(w[i] >> 32);
                asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;" : "+r"(l), "+r"(h), "+r"(a[(i + 1) % UNROLL]) : "r"(c), "r"(b));
                w[i] = ((unsigned long long)h << 32) | l;
            } else if (MODE == 15) {  // mad.lo.cc chain of 2 accumulations + 2 LOP3
                asm volatile("{ .reg .u32 l, h; mov.b64 {l,h}, %0; mad.lo.cc.u32 l, %1, %2, l; madc.hi.u32 h, %1, %2, h; mad.lo.cc.u32 l, %1, %3, l; madc.hi.u32 h, %1, %3, h; mov.b64 %0, {l,h}; }" : "+l"(w[i]) : "r"(a[i]), "r"(b), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(a[i]) : "r"(b), "r"(c));
            }
        }
    }
    unsigned long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; i++) acc ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

static const char* NAMES[] = {"imad_lo", "imad_wide_acc", "iadd3", "add64_pair", "imad_wide+iadd3", "imad_hi",
                              "lop3", "shf", "imad_lo+iadd3", "imad_wide+2iadd3", "isetp+sel", "imad_wide_imm", "madcc_pair", "mulwide+lop3", "2mulwide+lop3", "2madcc+2lop3", "mulwide_noconsumer", "madcc_acc", "madcc_acc_carry"};
static const int INSTR_PER_SLOT[] = {1, 1, 1, 2, 2, 1, 1, 1, 2, 3, 2, 1, 1, 2, 3, 4, 1, 1, 2};

template <int MODE>
void run(int sms, uint32_t* out, unsigned long long* clk, bool last) {
    int threads = 1024, blocks = sms;   // 32 warps/SM = 8 per SMSP
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench<MODE><<<blocks, threads>>>(out, 12345u, clk);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<MODE><<<blocks, threads>>>(out, 12345u, clk);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[1024]; cudaMemcpy(h, clk, sizeof(unsigned long long) * blocks, cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < blocks; i++) cyc += h[i]; cyc /= blocks;
    double warp_instr = (double)ITERS * UNROLL * INSTR_PER_SLOT[MODE] * (threads / 32);   // per SM
    double ipc = warp_instr / cyc;
    double thread_ops_per_s = warp_instr * 32.0 * blocks / (ms * 1e-3);
    printf("  \"%s\": {\"warp_instr_per_clk_per_sm\": %.3f, \"thread_instr_per_s\": %.4e, \"ms\": %.4f, \"cycles\": %.0f}%s\n",
           NAMES[MODE], ipc, thread_ops_per_s, ms, cyc, last ? "" : ",");
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    uint32_t* out; unsigned long long* clk;
    cudaMalloc(&out, sizeof(uint32_t) * 1024 * sms); cudaMalloc(&clk, sizeof(unsigned long long) * sms);
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, sms, p.clockRate);
    run<0>(sms, out, clk, false); run<1>(sms, out, clk, false); run<2>(sms, out, clk, false);
    run<3>(sms, out, clk, false); run<4>(sms, out, clk, false); run<5>(sms, out, clk, false);
    run<6>(sms, out, clk, false); run<7>(sms, out, clk, false); run<8>(sms, out, clk, false);
    run<9>(sms, out, clk, false); run<10>(sms, out, clk, false); run<11>(sms, out, clk, false);
    run<12>(sms, out, clk, false); run<13>(sms, out, clk, false); run<14>(sms, out, clk, false); run<15>(sms, out, clk, false);
    run<16>(sms, out, clk, false); run<17>(sms, out, clk, false); run<18>(sms, out, clk, true);
    printf("}\n");
    return 0;
}
