// Second integer-pipe microbenchmark (round 1b): how IMAD.WIDE.U32 behaves with and without a register addend and
// how much IADD3 / IADD3.X work overlaps with it.  Every mode's inner loop is checked in SASS (cuobjdump) before use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/intpeak2 tools/intpeak2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define U 8
typedef unsigned long long u64;
typedef unsigned int u32;

template <int MODE>
__global__ void __launch_bounds__(1024) bench(u32* out, u32 seed, u64* clk) {
    u32 a[U], lo[U], hi[U], cnt[U], b = seed | 1u, c = seed * 3u + 7u;
    u64 w[U];
#pragma unroll
    for (int i = 0; i < U; i++) { a[i] = threadIdx.x + i * seed; w[i] = a[i] * 0x9e3779b97f4a7c15ULL; lo[i] = a[i] * 3; hi[i] = a[i] * 5; cnt[i] = 0; }
    u64 t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < U; i++) {
            if (MODE == 0) {            // mul.wide (RZ addend), products xor-folded into the multiplicand (1 LOP3 per 2 WIDE)
                u64 t, u;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(a[i]), "r"(b));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(u) : "r"(a[i]), "r"(c));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a[i]) : "r"((u32)t), "r"((u32)(t >> 32)), "r"((u32)(u >> 32)));
            } else if (MODE == 1) {     // accumulating IMAD.WIDE: w = a*b + w
                asm volatile("mad.lo.cc.u32 %0, %1, %2, %0; madc.hi.u32 %1, %1, %2, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(b));
            } else if (MODE == 2) {     // accumulating IMAD.WIDE with carry-out + carry counter
                asm volatile("mad.lo.cc.u32 %0, %1, %3, %0; madc.hi.cc.u32 %1, %1, %3, %1; addc.u32 %2, %2, 0;" : "+r"(lo[i]), "+r"(hi[i]), "+r"(cnt[i]) : "r"(b));
            } else if (MODE == 3) {     // mul.wide + explicit 64-bit accumulate with carry counter on the ALU (add.cc/addc.cc/addc)
                u64 t;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(hi[i]), "r"(b));
                asm volatile("add.cc.u32 %0, %0, %3; addc.cc.u32 %1, %1, %4; addc.u32 %2, %2, 0;" : "+r"(lo[i]), "+r"(hi[i]), "+r"(cnt[i]) : "r"((u32)t), "r"((u32)(t >> 32)));
            } else if (MODE == 4) {     // carry chains only: 3 ALU instructions per slot
                asm volatile("add.cc.u32 %0, %0, %3; addc.cc.u32 %1, %1, %4; addc.u32 %2, %2, 0;" : "+r"(lo[i]), "+r"(hi[i]), "+r"(cnt[i]) : "r"(b), "r"(c));
            } else if (MODE == 5) {     // mul.wide + 6 ALU (two explicit accumulations of the same product)
                u64 t;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(hi[i]), "r"(b));
                asm volatile("add.cc.u32 %0, %0, %3; addc.cc.u32 %1, %1, %4; addc.u32 %2, %2, 0;" : "+r"(lo[i]), "+r"(hi[i]), "+r"(cnt[i]) : "r"((u32)t), "r"((u32)(t >> 32)));
                asm volatile("add.cc.u32 %0, %0, %4; addc.cc.u32 %1, %1, %3; addc.u32 %2, %2, 0;" : "+r"(lo[i]), "+r"(hi[i]), "+r"(cnt[i]) : "r"((u32)t), "r"((u32)(t >> 32)));
            } else if (MODE == 6) {     // 32-bit IMAD with register addend
                asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(lo[i]) : "r"(b));
            } else if (MODE == 7) {     // mul.wide whose addend is a zero-extended 32-bit value: w = a*b + (u64)hi32(w)
                asm volatile("{ .reg .u32 l, h; .reg .u64 z; mov.b64 {l,h}, %0; cvt.u64.u32 z, h; mad.wide.u32 %0, l, %1, z; }" : "+l"(w[i]) : "r"(b));
            } else if (MODE == 8) {     // lop3 only (ALU reference)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            } else if (MODE == 9) {     // mul.wide (RZ) + 3 LOP3
                u64 t;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(cnt[i]), "r"(b));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lo[i]) : "r"((u32)t), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(hi[i]) : "r"((u32)(t >> 32)), "r"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(cnt[i]) : "r"(lo[i]), "r"(c));
            }
        }
    }
    u64 t1 = clock64();
    u32 acc = 0;
#pragma unroll
    for (int i = 0; i < U; i++) acc ^= a[i] ^ (u32)w[i] ^ (u32)(w[i] >> 32) ^ lo[i] ^ hi[i] ^ cnt[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
static const char* NAMES[] = {"2 mul.wide(RZ) + lop3", "imad.wide acc", "imad.wide acc carry-out + addc", "mul.wide + add.cc/addc.cc/addc",
                              "add.cc/addc.cc/addc", "mul.wide + 6 carry-chain adds", "imad.lo acc", "mad.wide zext addend", "lop3", "mul.wide + 3 lop3"};
static const int SLOT[] = {3, 1, 2, 4, 3, 7, 1, 1, 1, 4};
template <int MODE> void run(int sms, u32* out, u64* clk) {
    bench<MODE><<<sms, 1024>>>(out, 12345u, clk);
    cudaDeviceSynchronize();
    bench<MODE><<<sms, 1024>>>(out, 12345u, clk);
    cudaDeviceSynchronize();
    u64 h[1024]; cudaMemcpy(h, clk, sizeof(u64) * sms, cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < sms; i++) cyc += h[i]; cyc /= sms;
    double slots_per_smsp = (double)ITERS * U * 8;          // 32 warps per SM = 8 per SMSP
    printf("  \"%s\": {\"cycles_per_slot_per_smsp\": %.3f, \"ptx_instr_per_slot\": %d},\n", NAMES[MODE], cyc / slots_per_smsp, SLOT[MODE]);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    u32* out; u64* clk;
    cudaMalloc(&out, sizeof(u32) * 1024 * sms); cudaMalloc(&clk, sizeof(u64) * sms);
    printf("{\n");
    run<0>(sms, out, clk); run<1>(sms, out, clk); run<2>(sms, out, clk); run<3>(sms, out, clk); run<4>(sms, out, clk);
    run<5>(sms, out, clk); run<6>(sms, out, clk); run<7>(sms, out, clk); run<8>(sms, out, clk); run<9>(sms, out, clk);
    printf("  \"device\": \"%s\"\n}\n", p.name);
    return 0;
}
