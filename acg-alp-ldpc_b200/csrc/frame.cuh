// Per-frame prologue / epilogue shared by the BP and QP-ADMM kernels.
//
// Both decoders run as PERSISTENT CTAs: one CTA decodes one frame at a time and
// pulls the next frame index from a global atomic queue when it finishes, so the
// heavy-tailed iteration counts never idle a slot.  Results are
// written per frame index (or folded into integer counters), so the outcome does
// not depend on the schedule.
//
//   decode mode      y comes from HBM, bits/ok/iters/soft go back to HBM
//                    (Decoder::decode, algo/algo.h:6-11, batched)
//   experiment mode  the codeword, AWGN and LLRs are produced on device and only
//                    the counter block leaves the GPU (exp(), experiment.h:80-123)
#ifndef LDPC_B200_FRAME_CUH
#define LDPC_B200_FRAME_CUH

#include "channel.cuh"
#include "ldpc_internal.h"

namespace ldpc {

struct KernelIO {
    // decode mode
    const double *y;
    uint8_t *bits;
    uint8_t *ok;
    int32_t *iters;
    double *soft;
    // experiment mode
    int experiment;
    int cw_source;
    uint64_t seed, frame_begin;
    const uint8_t *words;
    uint64_t n_words;
    unsigned long long *counters;
    const uint32_t *gen_cols;
    int k, k_words;
    // common
    long long frames;
    unsigned long long *queue;
    double var, sigma;
    int n, m;
    const uint16_t *row_ptr, *col_idx;
};

// small per-CTA scratch in shared memory
struct FrameScratch {
    long long frame;                 // frame index fetched for this round
    int hamming;                     // channel hard-decision errors of the frame (experiment.h:33-46)
    int mismatches;                  // decoded bits != transmitted bits
    unsigned int info[16];           // up to 512 information bits (generator mode)
    unsigned long long cnt[LDPC_CNT_COUNT];  // CTA-local counters, flushed once at kernel exit
};

__device__ __forceinline__ void scratch_init(FrameScratch *s) {
    if (threadIdx.x < LDPC_CNT_COUNT) s->cnt[threadIdx.x] = 0ull;
}

__device__ __forceinline__ void scratch_flush(const KernelIO &io, FrameScratch *s) {
    __syncthreads();
    if (io.experiment && threadIdx.x < LDPC_CNT_COUNT && s->cnt[threadIdx.x])
        atomicAdd(&io.counters[threadIdx.x], s->cnt[threadIdx.x]);
}

// Fetch the next frame for this CTA (-1 when the queue is drained).  Contains a barrier.
__device__ __forceinline__ long long next_frame(const KernelIO &io, FrameScratch *s) {
    __syncthreads();  // everyone is done with the previous frame's scratch
    if (threadIdx.x == 0) {
        s->frame = (long long) atomicAdd(io.queue, 1ull);
        s->hamming = 0;
        s->mismatches = 0;
    }
    __syncthreads();
    const long long f = s->frame;
    return f < io.frames ? f : -1;
}

// Fill llr[0..n) (= 2*y/sigma^2, utils/channel.h:14-16) for local frame f.
// Experiment mode also fills cw[0..n) with the transmitted codeword and counts
// the channel hard-decision errors.  Ends with a barrier.
__device__ __forceinline__ void load_frame(const KernelIO &io, long long f, double *llr, uint8_t *cw,
                                           FrameScratch *s) {
    const int n = io.n;
    if (!io.experiment) {
        const double *y = io.y + (size_t) f * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) llr[i] = __ddiv_rn(__dmul_rn(2.0, y[i]), io.var);
        __syncthreads();
        return;
    }
    const uint64_t gf = io.frame_begin + (uint64_t) f;
    if (io.cw_source == LDPC_CW_GENERATOR) {
        const int nblk = (io.k + 127) / 128;
        if ((int) threadIdx.x < nblk) {
            const uint4 w = info_block(io.seed, gf, threadIdx.x);
            s->info[4 * threadIdx.x + 0] = w.x;
            s->info[4 * threadIdx.x + 1] = w.y;
            s->info[4 * threadIdx.x + 2] = w.z;
            s->info[4 * threadIdx.x + 3] = w.w;
        }
        __syncthreads();
        const int tail = io.k & 31;
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            unsigned int acc = 0;
            for (int w = 0; w < io.k_words; ++w) {
                unsigned int u = s->info[w];
                if (w == io.k_words - 1 && tail) u &= (1u << tail) - 1u;
                acc ^= u & io.gen_cols[(size_t) j * io.k_words + w];
            }
            cw[j] = (uint8_t) (__popc(acc) & 1);
        }
    } else if (io.cw_source == LDPC_CW_TABLE) {
        const uint8_t *src = io.words + (size_t) (gf % io.n_words) * n;
        for (int j = threadIdx.x; j < n; j += blockDim.x) cw[j] = src[j] ? 1 : 0;
    } else {
        for (int j = threadIdx.x; j < n; j += blockDim.x) cw[j] = 0;
    }
    __syncthreads();
    int ham = 0;
    for (int blk = threadIdx.x; 2 * blk < n; blk += blockDim.x) {
        double z[2];
        noise_pair(io.seed, gf, (uint32_t) blk, z[0], z[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = 2 * blk + h;
            if (i < n) {
                const int c = cw[i];
                const double y = __fma_rn(io.sigma, z[h], c ? -1.0 : 1.0);
                ham += c ? (y > 0) : (y <= 0);
                llr[i] = __ddiv_rn(__dmul_rn(2.0, y), io.var);
            }
        }
    }
    ham = __reduce_add_sync(0xffffffffu, ham);
    if ((threadIdx.x & 31) == 0 && ham) atomicAdd(&s->hamming, ham);
    __syncthreads();
}

// Publish one decoded frame.  hard[0..n) = decisions in shared memory, soft[0..n)
// = posterior LLR / relaxed solution in shared memory (may be null).
//   ok        the decoder's bool
//   has_bits  0 when the reference would return an empty word (BP failure, bp.h:198)
//   valid     hard[] satisfies every check (IsCodeword, utils/codeword.h:90-95)
// Contains barriers; all threads must call it with identical ok/has_bits/valid/iters.
__device__ __forceinline__ void finish_frame(const KernelIO &io, long long f, const uint8_t *hard,
                                             const uint8_t *cw, const double *soft, int ok, int has_bits,
                                             int valid, int iters, FrameScratch *s) {
    const int n = io.n;
    if (!io.experiment) {
        uint8_t *out = io.bits + (size_t) f * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = has_bits ? hard[i] : (uint8_t) 0;
        if (io.soft && soft) {
            double *so = io.soft + (size_t) f * n;
            for (int i = threadIdx.x; i < n; i += blockDim.x) so[i] = soft[i];
        }
        if (threadIdx.x == 0) {
            io.ok[f] = (uint8_t) ok;
            io.iters[f] = iters;
        }
        return;
    }
    int mism = 0;
    if (has_bits)
        for (int i = threadIdx.x; i < n; i += blockDim.x) mism += hard[i] != cw[i];
    mism = __reduce_add_sync(0xffffffffu, mism);
    if ((threadIdx.x & 31) == 0 && mism) atomicAdd(&s->mismatches, mism);
    __syncthreads();
    if (threadIdx.x == 0) {
        // verdict, experiment.h:109-118
        const int is_codeword = ok && has_bits && valid;
        const int correct = is_codeword && s->mismatches == 0;
        s->cnt[LDPC_CNT_TOTAL] += 1;
        s->cnt[LDPC_CNT_CORRECT] += correct;
        s->cnt[LDPC_CNT_PSEUDO] += is_codeword && !correct;
        s->cnt[LDPC_CNT_DECODER_FAIL] += !ok;
        s->cnt[LDPC_CNT_BIT_ERRORS] += has_bits ? s->mismatches : 0;
        s->cnt[LDPC_CNT_SUM_HAMMING] += s->hamming;
        s->cnt[correct ? LDPC_CNT_SUM_HAMMING_OK : LDPC_CNT_SUM_HAMMING_WRONG] += s->hamming;
        s->cnt[LDPC_CNT_SUM_ITERS] += iters;
        s->cnt[LDPC_CNT_FRAMES_WITH_BITS] += has_bits;
    }
}

// every check has even parity over hard[]?  (block-wide; contains a barrier)
__device__ __forceinline__ int syndrome_ok(const KernelIO &io, const uint8_t *hard) {
    int bad = 0;
    for (int c = threadIdx.x; c < io.m; c += blockDim.x) {
        int parity = 0;
        for (int e = io.row_ptr[c]; e < io.row_ptr[c + 1]; ++e) parity ^= hard[io.col_idx[e]];
        bad |= parity;
    }
    return !__syncthreads_or(bad);
}

}  // namespace ldpc

#endif
