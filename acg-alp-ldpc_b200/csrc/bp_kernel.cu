// Belief-propagation (flooding sum-product) kernel (sm_100a) -- algo/bp.h:155-222.
//
// One persistent CTA decodes one frame at a time with both message arrays in
// shared memory; work is distributed per EDGE so every lane does one message:
//   check phase     C->V message of CSR edge e from the V->C messages of the other
//                   edges of its check (CNode::message, bp.h:49-57), written to the
//                   edge's CSC position
//   variable phase  V->C message of CSC edge p from the channel LLR and the other
//                   C->V messages of its variable (VNode::message, bp.h:77-83),
//                   written to the edge's CSR position; then per variable the
//                   posterior (estimate(), bp.h:85-90) and the hard decision
//   syndrome        per check parity of the decisions (IsCodeword, bp.h:195)
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
#include <algorithm>
#include <cstdlib>

#include "frame.cuh"

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

// V->C for every edge (CSC order in, CSR order out); with_estimate also produces
// posterior + decisions per variable.
__device__ __forceinline__ void variable_phase(const BpParams &p, const double *llr, const double *c2v,
                                               double *v2c, double *post, uint8_t *hard, bool with_estimate) {
    for (int e = threadIdx.x; e < p.E; e += blockDim.x) {
        const BpEdgeV ed = p.edge_v[e];
        double sum = 0.0;
        for (int o = ed.begin; o < ed.end; ++o)
            if (o != e) sum += c2v[o];
        const double t = llr[ed.var] + sum;
        const double mag = exp(-fabs(t));
        v2c[ed.dst] = (t <= 0.0) ? -mag : mag;
    }
    if (with_estimate) {
        for (int v = threadIdx.x; v < p.io.n; v += blockDim.x) {
            double sum = 0.0;
            for (int o = p.col_ptr[v]; o < p.col_ptr[v + 1]; ++o) sum += c2v[o];
            const double est = llr[v] + sum;
            post[v] = est;
            hard[v] = (est <= 0.0) ? 1 : 0;
        }
    }
}

// C->V for every edge (CSR order in, CSC order out)
__device__ __forceinline__ void check_phase(const BpParams &p, const double *v2c, double *c2v) {
    for (int e = threadIdx.x; e < p.E; e += blockDim.x) {
        const BpEdgeC ed = p.edge_c[e];
        double ev = 1.0, od = 0.0;
        int sign = 0;
        for (int o = ed.begin; o < ed.end; ++o) {
            if (o == e) continue;
            const double x = v2c[o];
            sign ^= __double2hiint(x);
            const double a = fabs(x);
            const double ne = fma(od, a, ev);
            od = fma(ev, a, od);
            ev = ne;
        }
        const double mag = log(ev / od);
        c2v[ed.dst] = (sign < 0) ? -mag : mag;
    }
}

__global__ void __launch_bounds__(512, 2) bp_kernel(const BpParams p) {
    extern __shared__ __align__(16) double smem[];
    const KernelIO &io = p.io;
    const int n = io.n, E = p.E;

    double *v2c = smem;            // E, CSR order
    double *c2v = v2c + E;         // E, CSC order
    double *llr = c2v + E;         // n
    double *post = llr + n;        // n
    FrameScratch *scratch = (FrameScratch *) (post + n);
    uint8_t *hard = (uint8_t *) (scratch + 1);
    uint8_t *cw = hard + n;

    scratch_init(scratch);
    for (;;) {
        const long long f = next_frame(io, scratch);
        if (f < 0) break;
        load_frame(io, f, llr, cw, scratch);
        for (int e = threadIdx.x; e < E; e += blockDim.x) c2v[e] = 0.0;     // CNode/VNode::init, bp.h:42-45
        for (int v = threadIdx.x; v < n; v += blockDim.x) { post[v] = llr[v]; hard[v] = 0; }
        __syncthreads();
        variable_phase(p, llr, c2v, v2c, post, hard, false);                // initial send, bp.h:184
        __syncthreads();
        int ok = 0, iters = 0;
        for (int it = 1; it <= p.max_iter; ++it) {
            iters = it;
            check_phase(p, v2c, c2v);                                       // bp.h:187
            __syncthreads();
            variable_phase(p, llr, c2v, v2c, post, hard, true);             // bp.h:188-193
            __syncthreads();
            ok = syndrome_ok(io, hard);                                     // bp.h:195
            if (ok && p.early_exit) break;
        }
        finish_frame(io, f, hard, cw, post, ok, ok, ok, iters, scratch);
    }
    scratch_flush(io, scratch);
}

// ---------------------------------------------------------------- host side

static size_t bp_smem_bytes(const ldpc_code *c) {
    return sizeof(double) * (2 * (size_t) c->E + 2 * (size_t) c->n) + sizeof(FrameScratch) + 2 * (size_t) c->n + 16;
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

    size_t smem = bp_smem_bytes(c);
    if (smem > 227 * 1024) return fail(LDPC_E_UNSUPPORTED, "BP messages of this code exceed 227 KB of shared memory");
    int threads = bp_threads(c);
    LDPC_CUDA(cudaFuncSetAttribute(bp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int per_sm = 0, sms = 0;
    LDPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bp_kernel, threads, smem));
    LDPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    if (per_sm < 1) return fail(LDPC_E_UNSUPPORTED, "BP state of this code does not fit on one SM");
    long long grid = std::min<long long>((long long) per_sm * sms, frames);
    bp_kernel<<<(unsigned) grid, threads, smem, stream>>>(p);
    LDPC_CUDA(cudaGetLastError());
    return LDPC_OK;
}

}  // namespace ldpc
