// Stand-alone channel kernels (y and codeword dumps for parity runs) and the fp64
// issue-rate microbenchmark that provides the FP64 roofline denominator.
#include <algorithm>

#include "bpmath.cuh"
#include "frame.cuh"

namespace ldpc {

// y[f][i] for local frames f in [0, frames): transmit(), utils/channel.h:19-26
__global__ void channel_kernel(uint64_t seed, uint64_t frame_begin, long long frames, int n, double sigma,
                               const uint8_t *__restrict__ codewords, double *__restrict__ y) {
    const int half = (n + 1) / 2;
    const long long total = frames * half;
    for (long long idx = blockIdx.x * (long long) blockDim.x + threadIdx.x; idx < total;
         idx += (long long) gridDim.x * blockDim.x) {
        const long long f = idx / half;
        const int blk = (int) (idx - f * half);
        double z[2];
        noise_pair(seed, frame_begin + (uint64_t) f, (uint32_t) blk, z[0], z[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = 2 * blk + h;
            if (i < n) {
                const int c = codewords ? codewords[(size_t) f * n + i] : 0;
                y[(size_t) f * n + i] = __fma_rn(sigma, z[h], c ? -1.0 : 1.0);
            }
        }
    }
}

// c = u * G with u from Philox stream 0: gen_random_codeword(), utils/channel.h:29-36
__global__ void generator_kernel(uint64_t seed, uint64_t frame_begin, long long frames, int n, int k, int k_words,
                                 const uint32_t *__restrict__ gen_cols, uint8_t *__restrict__ out) {
    const long long total = frames * n;
    for (long long idx = blockIdx.x * (long long) blockDim.x + threadIdx.x; idx < total;
         idx += (long long) gridDim.x * blockDim.x) {
        const long long f = idx / n;
        const int j = (int) (idx - f * n);
        unsigned int acc = 0;
        for (int w = 0; w < k_words; w += 4) {
            const uint4 u = info_block(seed, frame_begin + (uint64_t) f, (uint32_t) (w / 4));
            const unsigned int uw[4] = {u.x, u.y, u.z, u.w};
            for (int q = 0; q < 4 && w + q < k_words; ++q) acc ^= uw[q] & gen_cols[(size_t) j * k_words + w + q];
        }
        out[idx] = (uint8_t) (__popc(acc) & 1);
    }
}

int launch_channel(const ldpc_code *c, uint64_t seed, uint64_t frame_begin, int64_t frames, double sigma,
                   const uint8_t *d_codewords, double *d_y, cudaStream_t stream) {
    if (frames <= 0) return LDPC_OK;
    long long total = frames * ((c->n + 1) / 2);
    int grid = (int) std::min<long long>((total + 255) / 256, 148 * 16);
    channel_kernel<<<grid, 256, 0, stream>>>(seed, frame_begin, frames, c->n, sigma, d_codewords, d_y);
    LDPC_CUDA(cudaGetLastError());
    return LDPC_OK;
}

int launch_generator_codewords(const ldpc_code *c, uint64_t seed, uint64_t frame_begin, int64_t frames,
                               uint8_t *d_codewords, cudaStream_t stream) {
    if (frames <= 0) return LDPC_OK;
    if (!c->d.gen_cols || c->k <= 0) return fail(LDPC_E_INVALID, "no generator attached (ldpc_code_set_generator)");
    long long total = frames * c->n;
    int grid = (int) std::min<long long>((total + 255) / 256, 148 * 16);
    generator_kernel<<<grid, 256, 0, stream>>>(seed, frame_begin, frames, c->n, c->k, c->k_words, c->d.gen_cols,
                                               d_codewords);
    LDPC_CUDA(cudaGetLastError());
    return LDPC_OK;
}

__global__ void bpmath_kernel(int count, const double *a, const double *ev, const double *od, double *out_exp,
                              double *out_log) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        out_exp[i] = exp_neg(a[i]);
        out_log[i] = log_ratio(ev[i], od[i]);
    }
}

int debug_bpmath(int device, int count, const double *a, const double *ev, const double *od, double *out_exp,
                 double *out_log) {
    LDPC_CUDA(cudaSetDevice(device));
    double *d = nullptr;
    const size_t bytes = sizeof(double) * (size_t) count;
    LDPC_CUDA(cudaMalloc((void **) &d, 5 * bytes));
    LDPC_CUDA(cudaMemcpy(d, a, bytes, cudaMemcpyHostToDevice));
    LDPC_CUDA(cudaMemcpy(d + count, ev, bytes, cudaMemcpyHostToDevice));
    LDPC_CUDA(cudaMemcpy(d + 2 * (size_t) count, od, bytes, cudaMemcpyHostToDevice));
    bpmath_kernel<<<148, 256>>>(count, d, d + count, d + 2 * (size_t) count, d + 3 * (size_t) count,
                                d + 4 * (size_t) count);
    LDPC_CUDA(cudaGetLastError());
    LDPC_CUDA(cudaMemcpy(out_exp, d + 3 * (size_t) count, bytes, cudaMemcpyDeviceToHost));
    LDPC_CUDA(cudaMemcpy(out_log, d + 4 * (size_t) count, bytes, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return LDPC_OK;
}

// 8 independent DFMA chains per thread: measures lanes x clock of the FP64 pipe
__global__ void fp64_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
        x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

int measure_fp64_peak(int device, double *gfma_per_s) {
    LDPC_CUDA(cudaSetDevice(device));
    int sms = 0;
    LDPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int threads = 256, blocks = sms * 8, iters = 1 << 15;
    double *out = nullptr;
    LDPC_CUDA(cudaMalloc((void **) &out, sizeof(double) * threads * blocks));
    cudaEvent_t e0, e1;
    LDPC_CUDA(cudaEventCreate(&e0));
    LDPC_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        LDPC_CUDA(cudaEventRecord(e0));
        fp64_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        LDPC_CUDA(cudaEventRecord(e1));
        LDPC_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        LDPC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double rate = (double) threads * blocks * 8.0 * iters / (ms * 1e-3) / 1e9;
        if (rep >= 2 && rate > best) best = rate;   // two warm-ups
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *gfma_per_s = best;
    return LDPC_OK;
}

}  // namespace ldpc

namespace ldpc {

// conflict-free 16-byte shared-memory loads, 4 independent per iteration: measures bytes/clk x clock of the LSU path
__global__ void smem_peak_kernel(double *out, int iters) {
    __shared__ __align__(16) double2 buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_double2(i, -i);
    __syncthreads();
    double2 a0 = make_double2(0, 0), a1 = a0, a2 = a0, a3 = a0;
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
        const double2 x0 = buf[idx], x1 = buf[(idx + 256) & 2047], x2 = buf[(idx + 512) & 2047], x3 = buf[(idx + 768) & 2047];
        a0.x += x0.x; a1.y += x1.y; a2.x += x2.x; a3.y += x3.y;
        idx = (idx + 1024 + (int) (x0.y == 12345.0)) & 2047;      // data dependence keeps the loads in the loop
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (a0.x + a1.y) + (a2.x + a3.y);
}

int measure_smem_peak(int device, double *gbytes_per_s) {
    LDPC_CUDA(cudaSetDevice(device));
    int sms = 0;
    LDPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int threads = 256, blocks = sms * 4, iters = 1 << 14;
    double *out = nullptr;
    LDPC_CUDA(cudaMalloc((void **) &out, sizeof(double) * threads * blocks));
    cudaEvent_t e0, e1;
    LDPC_CUDA(cudaEventCreate(&e0));
    LDPC_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 6; ++rep) {
        LDPC_CUDA(cudaEventRecord(e0));
        smem_peak_kernel<<<blocks, threads>>>(out, iters);
        LDPC_CUDA(cudaEventRecord(e1));
        LDPC_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        LDPC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double rate = (double) threads * blocks * 4.0 * 16.0 * iters / (ms * 1e-3) / 1e9;
        if (rep >= 2 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *gbytes_per_s = best;
    return LDPC_OK;
}

}  // namespace ldpc
