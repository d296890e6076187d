// Frame slots for a TEAM of warps inside a CTA (the likelihood-ratio BP kernel runs several independent teams per
// CTA, each with its own frames, behind its own named barrier, so that the read-only tables are held once per SM and
// no team ever waits for another).  Same arithmetic and the same per-frame results as the per-slot versions of
// slots.cuh, with (threadIdx.x, blockDim.x, __syncthreads) replaced by (team.tid, team.nt, team.sync), and ALL the frames
// of a slot mask loaded / published in one pass (the per-slot versions cost two to three barriers per frame, which
// dominates when frames finish after three or four iterations -- SNR >= -1 dB).
#ifndef LDPC_B200_SLOTS_TEAM_CUH
#define LDPC_B200_SLOTS_TEAM_CUH

#include "slots.cuh"

namespace ldpc {

// slot index of the k-th set bit of mask (k < popc(mask))
__device__ __forceinline__ int nth_slot(unsigned mask, int k) { return (int) __fns(mask, 0, k + 1); }

struct Team {
    int tid, nt;        // thread index inside the team, threads of the team (a multiple of 32)
    int bar;            // named barrier of the team (1..15; 0 is __syncthreads)
    __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nt) : "memory"); }
};

template <int F>
__device__ __forceinline__ void team_slots_init(const Team &t, SlotBlock<F> *S) {
    if (t.tid < F) S->state[t.tid] = SLOT_EMPTY;
    if (t.tid < LDPC_CNT_COUNT) S->cnt[t.tid] = 0ull;
}

// LLRs (2 y / sigma^2, utils/channel.h:14-16) of the frames entering the slots of `mask`; experiment mode also produces
// the transmitted codewords cw[f * n + i] and the channel Hamming counts.  per_var(i, f, llr) initialises the kernel's
// per-variable state.  Ends with a team barrier.
template <int F, typename PerVar>
__device__ __forceinline__ void team_slots_load_all(const Team &t, const KernelIO &io, SlotBlock<F> *S, unsigned mask,
                                                    uint8_t *cw, PerVar per_var) {
    const int n = io.n, nf = __popc(mask), tid = t.tid, nt = t.nt;
    if (!io.experiment) {
        for (int idx = tid; idx < nf * n; idx += nt) {
            const int k = idx / n, i = idx - k * n, f = nth_slot(mask, k);
            const double y = io.y[(size_t) S->frame[f] * n + i];
            per_var(i, f, __ddiv_rn(__dmul_rn(2.0, y), io.var));
        }
        t.sync();
        return;
    }
    if (io.cw_source == LDPC_CW_GENERATOR) {
        const int nblk = (io.k + 127) / 128;
        for (int idx = tid; idx < nf * nblk; idx += nt) {
            const int k = idx / nblk, b = idx - k * nblk, f = nth_slot(mask, k);
            const uint4 w = info_block(io.seed, io.frame_begin + (uint64_t) S->frame[f], b);
            S->info[f][4 * b + 0] = w.x;
            S->info[f][4 * b + 1] = w.y;
            S->info[f][4 * b + 2] = w.z;
            S->info[f][4 * b + 3] = w.w;
        }
        t.sync();
    }
    for (int idx = tid; idx < nf * n; idx += nt) {
        const int k = idx / n, j = idx - k * n, f = nth_slot(mask, k);
        uint8_t bit = 0;
        if (io.cw_source == LDPC_CW_GENERATOR) {
            unsigned int acc = 0;
            for (int w = 0; w < io.k_words; ++w) acc ^= S->info[f][w] & io.gen_cols[(size_t) j * io.k_words + w];
            bit = (uint8_t) (__popc(acc) & 1);
        } else if (io.cw_source == LDPC_CW_TABLE) {
            const uint64_t gf = io.frame_begin + (uint64_t) S->frame[f];
            bit = io.words[(size_t) (gf % io.n_words) * n + j] ? 1 : 0;
        }
        cw[(size_t) f * n + j] = bit;
    }
    t.sync();
    const int half = (n + 1) / 2;
    for (int idx = tid; idx < nf * half; idx += nt) {
        const int k = idx / half, blk = idx - k * half, f = nth_slot(mask, k);
        const uint64_t gf = io.frame_begin + (uint64_t) S->frame[f];
        const uint8_t *c = cw + (size_t) f * n;
        double z[2];
        noise_pair(io.seed, gf, (uint32_t) blk, z[0], z[1]);
        int ham = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = 2 * blk + h;
            if (i < n) {
                const int bit = c[i];
                const double y = __fma_rn(io.sigma, z[h], bit ? -1.0 : 1.0);
                ham += bit ? (y > 0) : (y <= 0);
                per_var(i, f, __ddiv_rn(__dmul_rn(2.0, y), io.var));
            }
        }
        if (ham) atomicAdd(&S->hamming[f], ham);
    }
    t.sync();
}

// Publish the frames of `mask`.  okmask: decoder flags (BP: also "has bits" and "is a codeword").  hard(i, f) /
// soft(i, f): decision and soft output of variable i of slot f.  All threads of the team call with identical arguments.
template <int F, typename Hard, typename Soft>
__device__ __forceinline__ void team_slots_finish_all(const Team &t, const KernelIO &io, SlotBlock<F> *S, unsigned mask,
                                                      unsigned okmask, const uint8_t *cw, Hard hard, Soft soft) {
    const int n = io.n, nf = __popc(mask), tid = t.tid, nt = t.nt;
    if (!io.experiment) {
        for (int idx = tid; idx < nf * n; idx += nt) {
            const int k = idx / n, i = idx - k * n, f = nth_slot(mask, k);
            const size_t o = (size_t) S->frame[f] * n + i;
            io.bits[o] = ((okmask >> f) & 1u) ? (uint8_t) hard(i, f) : (uint8_t) 0;
            if (io.soft) io.soft[o] = soft(i, f);
        }
        if (tid < F && ((mask >> tid) & 1u)) {
            io.ok[S->frame[tid]] = (uint8_t) ((okmask >> tid) & 1u);
            io.iters[S->frame[tid]] = S->iter[tid];
        }
        return;
    }
    if (tid < F) S->red[tid] = 0;
    t.sync();
    for (int idx = tid; idx < nf * n; idx += nt) {
        const int k = idx / n, i = idx - k * n, f = nth_slot(mask, k);
        if (((okmask >> f) & 1u) && hard(i, f) != cw[(size_t) f * n + i]) atomicAdd(&S->red[f], 1);
    }
    t.sync();
    if (tid == 0) {
        for (int f = 0; f < F; ++f) {
            if (!((mask >> f) & 1u)) continue;
            // verdict, experiment.h:109-118 (BP: a frame with ok has bits and satisfies every check)
            const int ok = (okmask >> f) & 1u, mism = S->red[f];
            const int correct = ok && mism == 0;
            const int ham = S->hamming[f];
            S->cnt[LDPC_CNT_TOTAL] += 1;
            S->cnt[LDPC_CNT_CORRECT] += correct;
            S->cnt[LDPC_CNT_PSEUDO] += ok && !correct;
            S->cnt[LDPC_CNT_DECODER_FAIL] += !ok;
            S->cnt[LDPC_CNT_BIT_ERRORS] += ok ? mism : 0;
            S->cnt[LDPC_CNT_SUM_HAMMING] += ham;
            S->cnt[correct ? LDPC_CNT_SUM_HAMMING_OK : LDPC_CNT_SUM_HAMMING_WRONG] += ham;
            S->cnt[LDPC_CNT_SUM_ITERS] += S->iter[f];
            S->cnt[LDPC_CNT_FRAMES_WITH_BITS] += ok;
        }
    }
}

template <int F>
__device__ __forceinline__ void team_slots_flush(const Team &t, const KernelIO &io, SlotBlock<F> *S) {
    t.sync();
    if (io.experiment && t.tid < LDPC_CNT_COUNT && S->cnt[t.tid]) atomicAdd(&io.counters[t.tid], S->cnt[t.tid]);
}

}  // namespace ldpc

#endif
