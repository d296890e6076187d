// Microbenchmark: how does the B200 FP64 pipe share issue slots with integer / shared-memory work?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix_bench fp64_mix_bench.cu && ./fp64_mix_bench
// Each thread runs CH independent DFMA chains; per DFMA it also issues NI integer ops (LOP3/IADD chains) and
// optionally one LDS.64.  Reports fp64 thread-instructions per clock per SM and total instructions per clock.
#include <cstdio>
#include <cuda_runtime.h>

template <int CH, int NI, int LDS>
__global__ void mix(double *out, int iters, double a, double b) {
    __shared__ double sh[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = i * 1e-3;
    __syncthreads();
    double x[CH];
    unsigned k[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 11u};
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-6 + c;
    double acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            x[c] = __fma_rn(x[c], a, b);
#pragma unroll
            for (int q = 0; q < NI; ++q) k[q & 3] = (k[q & 3] ^ (k[(q + 1) & 3] + 0x9e3779b9u)) + it;
            if (LDS) acc += sh[(k[0] + c * 32 + threadIdx.x) & 2047];
        }
    }
    double s = acc;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double) (k[0] ^ k[1] ^ k[2] ^ k[3]);
}

template <int CH, int NI, int LDS>
void run(int threads, int ctas_per_sm, const char *name) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * ctas_per_sm * threads);
    const int iters = 20000;
    mix<CH, NI, LDS><<<sms * ctas_per_sm, threads>>>(out, 100, 1.0000001, 1e-9);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mix<CH, NI, LDS><<<sms * ctas_per_sm, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double fp64 = (double) iters * CH * threads * ctas_per_sm;       // thread instr per SM
    // integer ops per inner step: NI x (LOP3 + IADD + IADD) roughly 2-3 instr
    printf("%-34s threads/SM=%4d  fp64 %.1f lanes/clk/SM  (%.2f ms)\n", name, threads * ctas_per_sm, fp64 / cycles, ms);
    cudaFree(out);
}

int main() {
    run<8, 0, 0>(1024, 2, "DFMA x8 chains");
    run<8, 0, 0>(640, 1, "DFMA x8 chains");
    run<8, 0, 0>(256, 1, "DFMA x8 chains");
    run<2, 0, 0>(640, 1, "DFMA x2 chains");
    run<4, 0, 0>(640, 1, "DFMA x4 chains");
    run<8, 1, 0>(640, 1, "DFMA x8 + 1 int group each");
    run<8, 2, 0>(640, 1, "DFMA x8 + 2 int groups each");
    run<8, 4, 0>(640, 1, "DFMA x8 + 4 int groups each");
    run<8, 0, 1>(640, 1, "DFMA x8 + LDS.64 each");
    run<8, 2, 1>(640, 1, "DFMA x8 + 2 int + LDS.64 each");
    run<4, 2, 0>(640, 1, "DFMA x4 + 2 int groups each");
    run<4, 2, 0>(1024, 1, "DFMA x4 + 2 int groups each");
    return 0;
}
