// S-box throughput microbenchmark: 12 independent x^7 per thread per iteration, as in a Poseidon full round.
// Reports SMSP cycles per x^7 for several formulations and occupancies.  Build with -DGL_REDUCE_ALU to compare.
#include <cstdio>
#include <cuda_runtime.h>
#include "../vectorx_b200/csrc/gl.cuh"
#define ITERS 256
template <int MODE, int LANES>
__global__ void bench(u64* out, u64 seed, u64* clk) {
    u64 s[LANES];
#pragma unroll
    for (int i = 0; i < LANES; i++) s[i] = seed * (threadIdx.x + 1 + 977 * i) + blockIdx.x;
    u64 t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < LANES; i++) {
            if (MODE == 0) s[i] = gl_pow7_cc(s[i]);
            else if (MODE == 1) s[i] = gl_pow7(s[i]);
            else if (MODE == 2) s[i] = gl_mul_cc(s[i], s[(i + 1) % LANES]);
            else if (MODE == 3) s[i] = gl_sqr_cc(s[i]);
        }
    }
    u64 t1 = clock64();
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < LANES; i++) acc ^= s[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int MODE, int LANES> void run(const char* name, int sms, int threads, u64* out, u64* clk, int ops_per_lane) {
    bench<MODE, LANES><<<sms, threads>>>(out, 0x9e3779b97f4a7c15ULL, clk);
    cudaDeviceSynchronize();
    bench<MODE, LANES><<<sms, threads>>>(out, 0x9e3779b97f4a7c15ULL, clk);
    cudaDeviceSynchronize();
    u64 h[1024]; cudaMemcpy(h, clk, sizeof(u64) * sms, cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < sms; i++) cyc += h[i]; cyc /= sms;
    double per_smsp = (double)ITERS * LANES * (threads / 32) / 4.0;
    printf("  {\"mode\": \"%s\", \"lanes\": %d, \"warps_per_smsp\": %d, \"cycles_per_op_per_smsp\": %.2f},\n", name, LANES, threads / 128, cyc / per_smsp);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    u64 *out, *clk;
    cudaMalloc(&out, 8 * 1024 * sms); cudaMalloc(&clk, 8 * sms);
    printf("[\n");
    for (int th : {128, 256, 512, 1024}) {
        run<0, 12>("pow7_cc", sms, th, out, clk, 1);
        run<1, 12>("pow7_c", sms, th, out, clk, 1);
        run<2, 12>("mul_cc", sms, th, out, clk, 1);
        run<3, 12>("sqr_cc", sms, th, out, clk, 1);
        run<0, 4>("pow7_cc", sms, th, out, clk, 1);
    }
    printf("  {}\n]\n");
    return 0;
}
