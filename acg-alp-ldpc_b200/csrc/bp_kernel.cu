// Belief-propagation (flooding sum-product) kernel (sm_100a) -- algo/bp.h:155-222.
//
// A persistent CTA keeps F frames in flight (slots.cuh) with both message arrays
// in shared memory, interleaved by frame (msg[edge * F + f]); work is distributed
// per EDGE so every lane produces one message for each of the F frames:
//   variable phase  V->C message of CSC edge p from the channel LLR and the other
//                   C->V messages of its variable (VNode::message, bp.h:77-83),
//                   written to the edge's CSR position; the first edge of each
//                   variable also forms the posterior (estimate(), bp.h:85-90) and
//                   the hard decision
//   syndrome        per check parity of the decisions (IsCodeword, bp.h:195); the
//                   decisions of the F frames of a variable share one 32-bit word
//   check phase     C->V message of CSR edge e from the V->C messages of the other
//                   edges of its check (CNode::message, bp.h:49-57), written to the
//                   edge's CSC position
// A trip of the main loop is V, S, C for all slots; a frame that entered a slot
// this trip has all C->V messages zero, so its V phase is the reference's initial
// send (bp.h:184) and its syndrome is ignored, exactly the reference's schedule.
//
// Message algebra.  The reference sends phi(|t|) = -log tanh(|t|/2) in long double.
// fp64 cannot evaluate that form accurately once tanh rounds to 1, so the V->C
// message is carried as  s * E,  E = exp(-|t|)  (sign bit = sign of t, with t <= 0
// counted negative as bp.h:82 does), for which
//     tanh(|t|/2) = (1 - E) / (1 + E)
//     prod_i tanh(|t_i|/2) = (ev - od) / (ev + od),   (ev + od) = prod_i (1 + E_i),
//     phi(sum_i phi(|t_i|)) = 2 atanh(prod_i tanh(|t_i|/2)) = log(ev / od)
// where ev / od are the even / odd elementary symmetric sums of the E_i, built
// with positive terms only (no cancellation), so the C->V magnitude log(ev/od) has
// full relative accuracy at every magnitude.  This is the same function the
// reference computes; per frame agreement with it is checked in tests/.
// Magnitudes are capped near 700 (bpmath.cuh) instead of saturating to infinity.
#include <algorithm>
#include <cstdlib>

#include "bpmath.cuh"
#include "slots.cuh"

namespace ldpc {

struct BpParams {
    KernelIO io;
    const BpEdgeC *edge_c;
    const BpEdgeV *edge_v;
    const uint16_t *col_ptr;
    int E;
    int max_iter;
    int early_exit;
};

template <int F, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) bp_kernel(const BpParams p) {
    extern __shared__ __align__(16) double smem[];
    const KernelIO &io = p.io;
    const int n = io.n, E = p.E;
    const int tid = threadIdx.x, nt = blockDim.x;

    double *v2c = smem;                     // E * F, CSR order
    double *c2v = v2c + (size_t) E * F;     // E * F, CSC order
    double *llr = c2v + (size_t) E * F;     // n * F
    SlotBlock<F> *S = reinterpret_cast<SlotBlock<F> *>(llr + (size_t) n * F);
    uint32_t *hard = reinterpret_cast<uint32_t *>(S + 1);   // n words: byte f = decision of slot f
    uint8_t *hard8 = reinterpret_cast<uint8_t *>(hard);
    uint8_t *cw = hard8 + 4 * (size_t) n;                    // F * n (experiment mode)

    slots_init(S);
    for (int e = tid; e < E * F; e += nt) c2v[e] = 0.0;     // CNode/VNode::init, bp.h:42-45
    for (int v = tid; v < n; v += nt) hard[v] = 0u;
    __syncthreads();

    for (;;) {
        slots_refill(io, S);
        if (S->alive == 0) break;
        unsigned newmask = 0, livemask = 0;
        int iter[F];
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const int st = S->state[f];
            newmask |= (st == SLOT_NEW ? 1u : 0u) << f;
            livemask |= (st != SLOT_DEAD ? 1u : 0u) << f;
            iter[f] = S->iter[f];
        }
        if (newmask)   // decisions of variables without edges never change: channel hard decision
            slots_load(io, S, newmask, llr, cw, [&](int i, int f, double l) { hard8[4 * i + f] = l <= 0.0 ? 1 : 0; });

        // ---- variable phase: bp.h:77-83 (+ estimate and decision, bp.h:85-90, :191-193)
        for (int pos = tid; pos < E; pos += nt) {
            const BpEdgeV ed = p.edge_v[pos];
            double sum[F], x[F];
#pragma unroll
            for (int f = 0; f < F; ++f) sum[f] = 0.0;
            for (int o = ed.begin; o < pos; ++o) {
                ldv<F>(c2v + o * F, x);
#pragma unroll
                for (int f = 0; f < F; ++f) sum[f] += x[f];
            }
            for (int o = pos + 1; o < ed.end; ++o) {
                ldv<F>(c2v + o * F, x);
#pragma unroll
                for (int f = 0; f < F; ++f) sum[f] += x[f];
            }
            double l[F], out[F];
            ldv<F>(llr + ed.var * F, l);
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const double t = l[f] + sum[f];
                const double mag = exp_neg(fabs(t));
                out[f] = (t <= 0.0) ? -mag : mag;
            }
            stv<F>(v2c + ed.dst * F, out);
            if (pos == ed.begin) {
                ldv<F>(c2v + pos * F, x);
                uint32_t word = 0;
#pragma unroll
                for (int f = 0; f < F; ++f) word |= (l[f] + (sum[f] + x[f]) <= 0.0 ? 1u : 0u) << (8 * f);
                hard[ed.var] = word;
            }
        }
        __syncthreads();

        // ---- syndrome: bp.h:195 -> utils/codeword.h:90-95, all slots at once
        {
            uint32_t acc_any = 0;
            for (int c = tid; c < io.m; c += nt) {
                uint32_t acc = 0;
                for (int e = io.row_ptr[c]; e < io.row_ptr[c + 1]; ++e) acc ^= hard[io.col_idx[e]];
                acc_any |= acc;
            }
            acc_any = __reduce_or_sync(0xffffffffu, acc_any);
            if ((tid & 31) == 0 && acc_any) atomicOr(&S->bad, acc_any);
        }
        __syncthreads();
        const unsigned bad = S->bad;

        // ---- per slot: converged / out of iterations?  (uniform over the CTA)
        unsigned cmask = 0;
#pragma unroll
        for (int f = 0; f < F; ++f) {
            if (!((livemask >> f) & 1u)) continue;
            const int it = iter[f];
            const int ok = it >= 1 && ((bad >> (8 * f)) & 0xffu) == 0;
            const bool finished = (ok && p.early_exit) || it >= p.max_iter;
            if (!finished) {
                cmask |= 1u << f;
                continue;
            }
            slot_finish<F>(io, S, f, ok, ok, ok, it, cw,
                           [&](int i) { return (int) hard8[4 * i + f]; },
                           [&](int i) {
                               double sum = 0.0;
                               for (int o = p.col_ptr[i]; o < p.col_ptr[i + 1]; ++o) sum += c2v[o * F + f];
                               return llr[i * F + f] + sum;
                           });
        }

        // ---- check phase: bp.h:49-57; slots that are not iterating get zero messages,
        // which is the initial state of the next frame entering them
        for (int e = tid; e < E; e += nt) {
            const BpEdgeC ed = p.edge_c[e];
            double ev[F], od[F], x[F];
            int sg[F];
#pragma unroll
            for (int f = 0; f < F; ++f) { ev[f] = 1.0; od[f] = 0.0; sg[f] = 0; }
            for (int o = ed.begin; o < ed.end; ++o) {
                if (o == e) continue;
                ldv<F>(v2c + o * F, x);
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    sg[f] ^= __double2hiint(x[f]);
                    const double a = fabs(x[f]);
                    const double ne = __fma_rn(od[f], a, ev[f]);
                    od[f] = __fma_rn(ev[f], a, od[f]);
                    ev[f] = ne;
                }
            }
            double out[F];
#pragma unroll
            for (int f = 0; f < F; ++f) {
                out[f] = 0.0;
                if ((cmask >> f) & 1u) {
                    const double mag = log_ratio(ev[f], od[f]);
                    out[f] = sg[f] < 0 ? -mag : mag;
                }
            }
            stv<F>(c2v + ed.dst * F, out);
        }
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int f = 0; f < F; ++f) {
                if (!((livemask >> f) & 1u)) continue;
                if ((cmask >> f) & 1u) { S->iter[f] = iter[f] + 1; S->state[f] = SLOT_ACTIVE; }
                else S->state[f] = SLOT_EMPTY;
            }
        }
    }
    slots_flush(io, S);
}

// ---------------------------------------------------------------- host side

static size_t bp_smem_bytes(const ldpc_code *c, int F) {
    return sizeof(double) * F * (2 * (size_t) c->E + (size_t) c->n) + sizeof(SlotBlock<4>) + 4 * (size_t) c->n +
           (size_t) F * c->n + 32;
}

// CTA size: the multiple of 32 in [128, 512] that wastes the fewest lanes on E edges
static int bp_threads(const ldpc_code *c) {
    const char *force = getenv("LDPC_BP_THREADS");
    if (force && atoi(force) >= 32 && atoi(force) <= 512) return atoi(force) / 32 * 32;
    int best_nt = 128;
    double best = -1;
    for (int nt = 128; nt <= 512; nt += 32) {
        int rounds = (c->E + nt - 1) / nt;
        double eff = (double) c->E / ((double) rounds * nt);
        if (eff > best + 1e-9) { best = eff; best_nt = nt; }
    }
    return best_nt;
}

template <int F, int MAXT, int MINB>
static int launch_bp_f(const BpParams &p, const ldpc_code *c, int threads, int64_t frames, cudaStream_t stream) {
    const size_t smem = bp_smem_bytes(c, F);
    auto kernel = bp_kernel<F, MAXT, MINB>;
    LDPC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int per_sm = 0, sms = 0;
    LDPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    LDPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    if (per_sm < 1) return fail(LDPC_E_UNSUPPORTED, "BP state of this code does not fit on one SM");
    const long long want = (frames + F - 1) / F;
    const long long grid = std::min<long long>((long long) per_sm * sms, want);
    kernel<<<(unsigned) grid, threads, smem, stream>>>(p);
    LDPC_CUDA(cudaGetLastError());
    return LDPC_OK;
}

int launch_bp(const ldpc_code *c, const FrameIO &fio, int64_t frames, double var, int max_iter, int early_exit,
              unsigned long long *queue, cudaStream_t stream) {
    if (frames <= 0) return LDPC_OK;
    BpParams p;
    KernelIO &io = p.io;
    io.y = fio.y; io.bits = fio.bits; io.ok = fio.ok; io.iters = fio.iters; io.soft = fio.soft;
    io.experiment = fio.experiment; io.cw_source = fio.cw_source; io.seed = fio.seed;
    io.frame_begin = fio.frame_begin; io.words = fio.words; io.n_words = fio.n_words;
    io.counters = fio.counters; io.gen_cols = c->d.gen_cols; io.k = c->k; io.k_words = c->k_words;
    io.frames = frames; io.queue = queue; io.var = var; io.sigma = std::sqrt(var);
    io.n = c->n; io.m = c->m; io.row_ptr = c->d.row_ptr; io.col_idx = c->d.col_idx;
    p.edge_c = c->d.bp_c; p.edge_v = c->d.bp_v; p.col_ptr = c->d.col_ptr;
    p.E = c->E; p.max_iter = max_iter; p.early_exit = early_exit;

    // frames in flight per CTA: as many as keep at least two CTAs of messages on one SM
    int F = 4;
    while (F > 1 && 2 * bp_smem_bytes(c, F) > 227 * 1024) F >>= 1;
    if (frames < 4 * 148 * F) F = 1;                 // tiny batches: spread the frames over the SMs instead
    if (const char *force = getenv("LDPC_BP_F")) {
        const int v = atoi(force);
        if (v == 1 || v == 2 || v == 4) F = v;
    }
    if (bp_smem_bytes(c, F) > 227 * 1024)
        return fail(LDPC_E_UNSUPPORTED, "BP messages of this code exceed 227 KB of shared memory");
    const int threads = bp_threads(c);
    switch (F) {
        case 4: return launch_bp_f<4, 512, 1>(p, c, threads, frames, stream);
        case 2: return launch_bp_f<2, 512, 2>(p, c, threads, frames, stream);
        default: return launch_bp_f<1, 512, 2>(p, c, threads, frames, stream);
    }
}

}  // namespace ldpc
