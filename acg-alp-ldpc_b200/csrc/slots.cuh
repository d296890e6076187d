// Frame slots: the scheduling skeleton shared by the decoding kernels.
//
// A persistent CTA keeps F frames ("slots") in flight and advances them in lock
// step, one decoder iteration per trip of its main loop; a lane works on one
// (graph element, frame) pair, so F frames fill the warps of a CTA even when the
// graph has few nodes of one kind.  Frames finish at
// very different iteration counts, so a slot whose frame has
// finished is refilled from the global atomic frame queue at the top of the next
// trip while its neighbours keep iterating: no slot waits for the slowest frame.
// Results are written per frame index (decode mode) or folded into integer
// counters (experiment mode), so the outcome is independent of the schedule.
#ifndef LDPC_B200_SLOTS_CUH
#define LDPC_B200_SLOTS_CUH

#include "frame.cuh"

namespace ldpc {

enum { SLOT_EMPTY = 0, SLOT_NEW = 1, SLOT_ACTIVE = 2, SLOT_DEAD = 3 };

template <int F>
struct SlotBlock {
    long long frame[F];
    int iter[F];        // decoder iterations completed by the frame in this slot
    int state[F];
    int hamming[F];     // channel hard-decision errors of the frame (experiment.h:33-46)
    int red[32];        // block-reduction scratch
    unsigned bad;       // BP: bit f set <=> some check of slot f is unsatisfied
    int alive;          // slots that still hold or may get a frame
    unsigned info[F][16];
    unsigned long long cnt[LDPC_CNT_COUNT];
};

template <int F>
__device__ __forceinline__ void slots_init(SlotBlock<F> *S) {
    if (threadIdx.x < F) S->state[threadIdx.x] = SLOT_EMPTY;
    if (threadIdx.x < LDPC_CNT_COUNT) S->cnt[threadIdx.x] = 0ull;
}

// Thread 0: retire finished frames' bookkeeping is done by the caller; here empty
// slots pull new frame indices.  Ends with a barrier; returns nothing (read S after).
template <int F>
__device__ __forceinline__ void slots_refill(const KernelIO &io, SlotBlock<F> *S) {
    if (threadIdx.x == 0) {
        int alive = 0;
#pragma unroll
        for (int f = 0; f < F; ++f) {
            if (S->state[f] == SLOT_EMPTY) {
                const long long fr = (long long) atomicAdd(io.queue, 1ull);
                if (fr < io.frames) {
                    S->frame[f] = fr;
                    S->iter[f] = 0;
                    S->hamming[f] = 0;
                    S->state[f] = SLOT_NEW;
                } else {
                    S->state[f] = SLOT_DEAD;
                }
            }
            alive += S->state[f] != SLOT_DEAD;
        }
        S->alive = alive;
        S->bad = 0u;
    }
    __syncthreads();
}

// Sum of x over the CTA (two barriers).
template <int F>
__device__ __forceinline__ int block_sum(int x, SlotBlock<F> *S) {
    x = __reduce_add_sync(0xffffffffu, x);
    if ((threadIdx.x & 31) == 0) S->red[threadIdx.x >> 5] = x;
    __syncthreads();
    int total = 0;
    for (int w = 0; w < (int) (blockDim.x >> 5); ++w) total += S->red[w];
    __syncthreads();
    return total;
}

// LLRs (2 y / sigma^2, utils/channel.h:14-16) of the frames entering the slots of
// `newmask`, frame-major as llr[f * stride + i]; experiment mode also produces the
// transmitted codeword cw[f * n + i] and the channel Hamming count.  `per_var(i, f,
// llr)` lets the kernel initialise per-variable state.  Ends with a barrier.
template <int F, typename PerVar>
__device__ __forceinline__ void slots_load(const KernelIO &io, SlotBlock<F> *S, unsigned newmask, double *llr,
                                           int stride, uint8_t *cw, PerVar per_var) {
    const int n = io.n;
    if (!io.experiment) {
#pragma unroll
        for (int f = 0; f < F; ++f) {
            if (!((newmask >> f) & 1u)) continue;
            const double *y = io.y + (size_t) S->frame[f] * n;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const double l = __ddiv_rn(__dmul_rn(2.0, y[i]), io.var);
                if (llr) llr[(size_t) f * stride + i] = l;
                per_var(i, f, l);
            }
        }
        __syncthreads();
        return;
    }
    if (io.cw_source == LDPC_CW_GENERATOR) {
        const int nblk = (io.k + 127) / 128;
        if ((int) threadIdx.x < nblk * F) {
            const int f = threadIdx.x / nblk, b = threadIdx.x - f * nblk;
            if ((newmask >> f) & 1u) {
                const uint4 w = info_block(io.seed, io.frame_begin + (uint64_t) S->frame[f], b);
                S->info[f][4 * b + 0] = w.x;
                S->info[f][4 * b + 1] = w.y;
                S->info[f][4 * b + 2] = w.z;
                S->info[f][4 * b + 3] = w.w;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int f = 0; f < F; ++f) {
        if (!((newmask >> f) & 1u)) continue;
        const uint64_t gf = io.frame_begin + (uint64_t) S->frame[f];
        uint8_t *c = cw + (size_t) f * n;
        if (io.cw_source == LDPC_CW_GENERATOR) {
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                unsigned int acc = 0;
                for (int w = 0; w < io.k_words; ++w) acc ^= S->info[f][w] & io.gen_cols[(size_t) j * io.k_words + w];
                c[j] = (uint8_t) (__popc(acc) & 1);
            }
        } else if (io.cw_source == LDPC_CW_TABLE) {
            const uint8_t *src = io.words + (size_t) (gf % io.n_words) * n;
            for (int j = threadIdx.x; j < n; j += blockDim.x) c[j] = src[j] ? 1 : 0;
        } else {
            for (int j = threadIdx.x; j < n; j += blockDim.x) c[j] = 0;
        }
    }
    __syncthreads();
#pragma unroll
    for (int f = 0; f < F; ++f) {
        if (!((newmask >> f) & 1u)) continue;
        const uint64_t gf = io.frame_begin + (uint64_t) S->frame[f];
        const uint8_t *c = cw + (size_t) f * n;
        int ham = 0;
        for (int blk = threadIdx.x; 2 * blk < n; blk += blockDim.x) {
            double z[2];
            noise_pair(io.seed, gf, (uint32_t) blk, z[0], z[1]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = 2 * blk + h;
                if (i < n) {
                    const int bit = c[i];
                    const double y = __fma_rn(io.sigma, z[h], bit ? -1.0 : 1.0);
                    ham += bit ? (y > 0) : (y <= 0);
                    const double l = __ddiv_rn(__dmul_rn(2.0, y), io.var);
                    if (llr) llr[(size_t) f * stride + i] = l;
                    per_var(i, f, l);
                }
            }
        }
        ham = __reduce_add_sync(0xffffffffu, ham);
        if ((threadIdx.x & 31) == 0 && ham) atomicAdd(&S->hamming[f], ham);
    }
    __syncthreads();
}

// Publish the frame of slot f.  hard(i) / soft(i) read the decisions and the soft
// output of variable i.  ok = decoder bool; has_bits = 0 when the reference
// returns an empty word (BP failure, bp.h:198); valid = decisions satisfy every
// check.  All threads call with identical arguments; experiment mode contains
// barriers.
template <int F, typename Hard, typename Soft>
__device__ __forceinline__ void slot_finish(const KernelIO &io, SlotBlock<F> *S, int f, int ok, int has_bits,
                                            int valid, int iters, const uint8_t *cw, Hard hard, Soft soft) {
    const int n = io.n;
    const long long frame = S->frame[f];
    if (!io.experiment) {
        uint8_t *out = io.bits + (size_t) frame * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = has_bits ? (uint8_t) hard(i) : (uint8_t) 0;
        if (io.soft) {
            double *so = io.soft + (size_t) frame * n;
            for (int i = threadIdx.x; i < n; i += blockDim.x) so[i] = soft(i);
        }
        if (threadIdx.x == 0) {
            io.ok[frame] = (uint8_t) ok;
            io.iters[frame] = iters;
        }
        return;
    }
    int mism = 0;
    if (has_bits) {
        const uint8_t *c = cw + (size_t) f * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) mism += hard(i) != c[i];
    }
    mism = block_sum(mism, S);
    if (threadIdx.x == 0) {
        // verdict, experiment.h:109-118
        const int is_codeword = ok && has_bits && valid;
        const int correct = is_codeword && mism == 0;
        const int ham = S->hamming[f];
        S->cnt[LDPC_CNT_TOTAL] += 1;
        S->cnt[LDPC_CNT_CORRECT] += correct;
        S->cnt[LDPC_CNT_PSEUDO] += is_codeword && !correct;
        S->cnt[LDPC_CNT_DECODER_FAIL] += !ok;
        S->cnt[LDPC_CNT_BIT_ERRORS] += has_bits ? mism : 0;
        S->cnt[LDPC_CNT_SUM_HAMMING] += ham;
        S->cnt[correct ? LDPC_CNT_SUM_HAMMING_OK : LDPC_CNT_SUM_HAMMING_WRONG] += ham;
        S->cnt[LDPC_CNT_SUM_ITERS] += iters;
        S->cnt[LDPC_CNT_FRAMES_WITH_BITS] += has_bits;
    }
}

template <int F>
__device__ __forceinline__ void slots_flush(const KernelIO &io, SlotBlock<F> *S) {
    __syncthreads();
    if (io.experiment && threadIdx.x < LDPC_CNT_COUNT && S->cnt[threadIdx.x])
        atomicAdd(&io.counters[threadIdx.x], S->cnt[threadIdx.x]);
}

// F interleaved doubles <-> registers (16-byte accesses when F >= 2)
template <int F>
__device__ __forceinline__ void ldv(const double *p, double (&x)[F]) {
    if constexpr (F == 1) {
        x[0] = p[0];
    } else {
#pragma unroll
        for (int i = 0; i < F; i += 2) {
            const double2 t = *reinterpret_cast<const double2 *>(p + i);
            x[i] = t.x;
            x[i + 1] = t.y;
        }
    }
}

template <int F>
__device__ __forceinline__ void stv(double *p, const double (&x)[F]) {
    if constexpr (F == 1) {
        p[0] = x[0];
    } else {
#pragma unroll
        for (int i = 0; i < F; i += 2) *reinterpret_cast<double2 *>(p + i) = make_double2(x[i], x[i + 1]);
    }
}

}  // namespace ldpc

#endif
