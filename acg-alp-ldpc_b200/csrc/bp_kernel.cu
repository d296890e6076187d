// Belief-propagation (flooding sum-product) kernel (sm_100a) -- algo/bp.h:155-222.
//
// Layout.  A persistent CTA keeps F frames in flight (slots.cuh).  Per frame the
// state in shared memory is ONE message array indexed by the CSR position of the
// edge, the channel LLRs and the posteriors:  (E + 2 n) * 8 bytes (11 KB for the
// 160 x 280 codes).  The variable phase reads the C->V message of an edge and
// overwrites it with the V->C message, the check phase does the opposite: every
// node updates its own edges in place.
//
// Mapping.  Work is per NODE and a lane is a (node, frame) pair; nodes are ranked
// by degree and every warp-job covers nodes of one degree, so the node update is a
// degree-templated, fully unrolled straight-line routine: each message is loaded
// once, the leave-one-out sums / products come from prefix-suffix passes
// (3(d-2) combines instead of d(d-1)), and the d transcendental evaluations of a
// node are d independent dependency chains.
//   variable pass  VNode::message (bp.h:77-83) for every edge of the variable, the
//                  posterior (estimate(), bp.h:85-90) and the hard decision; the
//                  decision rides in bit 62 of each outgoing message
//   syndrome pass  parity of the decision bits per check (IsCodeword, bp.h:195)
//   check pass     CNode::message (bp.h:49-57) for every edge of the check
// A trip of the main loop is V, S, C for all slots; a frame that entered a slot
// this trip has all messages zero, so its V pass is the reference's initial send
// (bp.h:184) and its syndrome is ignored, exactly the reference's schedule.
//
// Message algebra.  The reference sends phi(|t|) = -log tanh(|t|/2) in long double.
// fp64 cannot evaluate that form accurately once tanh rounds to 1, so the V->C
// message is carried as  s * E,  E = exp(-|t|) in [0,1]  (sign bit = sign of t, with
// t <= 0 counted negative as bp.h:82 does; bit 62, always clear for values below 2,
// carries the variable's hard decision), for which
//     tanh(|t|/2) = (1 - E) / (1 + E)
//     prod_i tanh(|t_i|/2) = (ev - od) / (ev + od),   (ev + od) = prod_i (1 + E_i),
//     phi(sum_i phi(|t_i|)) = 2 atanh(prod_i tanh(|t_i|/2)) = log(ev / od)
// where ev / od are the even / odd elementary symmetric sums of the E_i, built
// with positive terms only (no cancellation), so the C->V magnitude log(ev/od) has
// full relative accuracy at every magnitude.  This is the same function the
// reference computes; per frame agreement with it is checked in tests/.
// Magnitudes are capped near 700 (bpmath.cuh) instead of saturating to infinity.
#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "bpmath.cuh"
#include "slots.cuh"

namespace ldpc {

constexpr int BP_MAX_UNROLLED_DEGREE = 8;   // larger degrees take the generic (local-memory) routines
constexpr int BP_MAX_DEGREE = 64;

struct BpParams {
    KernelIO io;
    const uint16_t *chk_rs;
    const BpVarRec *var_rec;
    const uint16_t *var_edges;
    const BpJob *jobs_v, *jobs_c;
    int rounds_v, rounds_c;
    int E, stride_e, stride_n;   // per-frame strides (doubles) of the message and LLR/posterior arrays
    int max_iter;
    int early_exit;
};

__device__ __forceinline__ double with_hi(double x, int hi) { return __hiloint2double(hi, __double2loint(x)); }

// ---- variable node of degree D (D is a compile-time constant in the unrolled instances)
template <int D>
__device__ __forceinline__ void var_update(double *msg, const double *llr, double *post, const BpVarRec rec,
                                           const uint16_t *var_edges, int d_runtime) {
    const int d = D > 0 ? D : d_runtime;
    constexpr int CAP = D > 0 ? D : BP_MAX_DEGREE;
    int pos[CAP];
    double m[CAP], suf[CAP];
#pragma unroll
    for (int j = 0; j < d; ++j) {
        pos[j] = var_edges[rec.off + j];
        m[j] = msg[pos[j]];
    }
    const double l = llr[rec.var];
    double total = m[0];                       // sum in edge order, then llr + sum (bp.h:85-90)
#pragma unroll
    for (int j = 1; j < d; ++j) total += m[j];
    const double est = l + total;
    post[rec.var] = est;
    const int hard_bit = est <= 0.0 ? 0x40000000 : 0;      // decision (bp.h:193) rides in bit 62
    suf[d - 1] = 0.0;
#pragma unroll
    for (int j = d - 2; j >= 0; --j) suf[j] = (j == d - 2) ? m[j + 1] : suf[j + 1] + m[j + 1];
    double pre = 0.0;
#pragma unroll
    for (int j = 0; j < d; ++j) {
        const double others = (j == 0) ? suf[0] : (j == d - 1 ? pre : pre + suf[j]);
        const double t = l + others;                       // bp.h:81
        const double e = exp_neg_abs(t);
        const int sign_bit = t <= 0.0 ? (int) 0x80000000 : 0;   // bp.h:82: zero counts as negative
        msg[pos[j]] = with_hi(e, __double2hiint(e) | sign_bit | hard_bit);
        pre = (j == 0) ? m[0] : pre + m[j];
    }
}

// ---- check node of degree D
template <int D>
__device__ __forceinline__ void chk_update(double *msg, int rs, bool active, int d_runtime) {
    const int d = D > 0 ? D : d_runtime;
    constexpr int CAP = D > 0 ? D : BP_MAX_DEGREE;
    double *edge = msg + rs;
    if (!active) {                 // slot is not iterating: zero messages = initial state of its next frame
#pragma unroll
        for (int j = 0; j < d; ++j) edge[j] = 0.0;
        return;
    }
    double a[CAP], se[CAP], so[CAP];
    int h[CAP], tot = 0;
#pragma unroll
    for (int j = 0; j < d; ++j) {
        const double x = edge[j];
        h[j] = __double2hiint(x);
        tot ^= h[j];
        a[j] = with_hi(x, h[j] & 0x3fffffff);              // |x| without the decision bit
    }
    // suffix products S_j = prod_{i>j} (1, a_i) in the (even, odd) representation
    se[d - 1] = 1.0;
    so[d - 1] = 0.0;
#pragma unroll
    for (int j = d - 2; j >= 0; --j) {
        if (j == d - 2) { se[j] = 1.0; so[j] = a[j + 1]; }
        else { se[j] = __fma_rn(so[j + 1], a[j + 1], se[j + 1]); so[j] = __fma_rn(se[j + 1], a[j + 1], so[j + 1]); }
    }
    double pe = 1.0, po = 0.0;                              // prefix product P_j = prod_{i<j}
#pragma unroll
    for (int j = 0; j < d; ++j) {
        double ev, od;
        if (j == 0) { ev = se[0]; od = so[0]; }
        else if (j == d - 1) { ev = pe; od = po; }
        else { ev = __fma_rn(po, so[j], pe * se[j]); od = __fma_rn(po, se[j], pe * so[j]); }
        const double mag = log_ratio(ev, od);               // phi(sum of the others' phi), bp.h:56
        edge[j] = with_hi(mag, __double2hiint(mag) | ((tot ^ h[j]) & (int) 0x80000000));
        if (j == 0) { pe = 1.0; po = a[0]; }
        else { const double ne = __fma_rn(po, a[j], pe); po = __fma_rn(pe, a[j], po); pe = ne; }
    }
}

template <int F, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) bp_kernel(const BpParams p) {
    extern __shared__ __align__(16) double smem[];
    const KernelIO &io = p.io;
    const int n = io.n;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int warp = tid >> 5, nwarps = nt >> 5, lane = tid & 31;
    constexpr int NODES_PER_WARP = 32 / F;
    const int f_lane = lane % F, node_lane = lane / F;

    double *msg = smem;                                    // F x stride_e
    double *llr = msg + (size_t) F * p.stride_e;           // F x stride_n
    double *post = llr + (size_t) F * p.stride_n;          // F x stride_n
    SlotBlock<F> *S = reinterpret_cast<SlotBlock<F> *>(post + (size_t) F * p.stride_n);
    uint8_t *cw = reinterpret_cast<uint8_t *>(S + 1);      // F x n (experiment mode)

    slots_init(S);
    for (int e = tid; e < F * p.stride_e; e += nt) msg[e] = 0.0;    // CNode/VNode::init, bp.h:42-45
    __syncthreads();

    double *msg_f = msg + (size_t) f_lane * p.stride_e;
    const double *llr_f = llr + (size_t) f_lane * p.stride_n;
    double *post_f = post + (size_t) f_lane * p.stride_n;

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
        if (newmask)   // variables without edges keep posterior = channel LLR
            slots_load<F>(io, S, newmask, llr, p.stride_n, cw,
                          [&](int i, int f, double l) { post[(size_t) f * p.stride_n + i] = l; });

        // ---- variable pass
        for (int r = 0; r < p.rounds_v; ++r) {
            const BpJob job = p.jobs_v[r * nwarps + warp];
            if (job.degree == 0 || node_lane >= job.count) continue;
            const BpVarRec rec = p.var_rec[job.first + node_lane];
            switch (job.degree) {
                case 1: var_update<1>(msg_f, llr_f, post_f, rec, p.var_edges, 1); break;
                case 2: var_update<2>(msg_f, llr_f, post_f, rec, p.var_edges, 2); break;
                case 3: var_update<3>(msg_f, llr_f, post_f, rec, p.var_edges, 3); break;
                case 4: var_update<4>(msg_f, llr_f, post_f, rec, p.var_edges, 4); break;
                case 5: var_update<5>(msg_f, llr_f, post_f, rec, p.var_edges, 5); break;
                case 6: var_update<6>(msg_f, llr_f, post_f, rec, p.var_edges, 6); break;
                case 7: var_update<7>(msg_f, llr_f, post_f, rec, p.var_edges, 7); break;
                case 8: var_update<8>(msg_f, llr_f, post_f, rec, p.var_edges, 8); break;
                default: var_update<0>(msg_f, llr_f, post_f, rec, p.var_edges, job.degree); break;
            }
        }
        __syncthreads();

        // ---- syndrome pass: parity of the decision bits (bit 62 of the V->C messages) per check
        {
            unsigned bad = 0;
            for (int r = 0; r < p.rounds_c; ++r) {
                const BpJob job = p.jobs_c[r * nwarps + warp];
                if (job.degree == 0 || node_lane >= job.count) continue;
                const int *hi = reinterpret_cast<const int *>(msg_f + p.chk_rs[job.first + node_lane]) + 1;
                int acc = 0;
                for (int j = 0; j < job.degree; ++j) acc ^= hi[2 * j];
                bad |= ((unsigned) acc >> 30 & 1u) << f_lane;
            }
            bad = __reduce_or_sync(0xffffffffu, bad);
            if (lane == 0 && bad) atomicOr(&S->bad, bad);
        }
        __syncthreads();
        const unsigned bad = S->bad;

        // ---- per slot: converged / out of iterations?  (uniform over the CTA)
        unsigned cmask = 0;
#pragma unroll
        for (int f = 0; f < F; ++f) {
            if (!((livemask >> f) & 1u)) continue;
            const int it = iter[f];
            const int ok = it >= 1 && ((bad >> f) & 1u) == 0;           // bp.h:195 (not before iteration 1)
            const bool finished = (ok && p.early_exit) || it >= p.max_iter;
            if (!finished) {
                cmask |= 1u << f;
                continue;
            }
            const double *pf = post + (size_t) f * p.stride_n;
            slot_finish<F>(io, S, f, ok, ok, ok, it, cw, [&](int i) { return pf[i] <= 0.0 ? 1 : 0; },
                           [&](int i) { return pf[i]; });
        }

        // ---- check pass
        const bool active = (cmask >> f_lane) & 1u;
        for (int r = 0; r < p.rounds_c; ++r) {
            const BpJob job = p.jobs_c[r * nwarps + warp];
            if (job.degree == 0 || node_lane >= job.count) continue;
            const int rs = p.chk_rs[job.first + node_lane];
            switch (job.degree) {
                case 1: chk_update<1>(msg_f, rs, active, 1); break;
                case 2: chk_update<2>(msg_f, rs, active, 2); break;
                case 3: chk_update<3>(msg_f, rs, active, 3); break;
                case 4: chk_update<4>(msg_f, rs, active, 4); break;
                case 5: chk_update<5>(msg_f, rs, active, 5); break;
                case 6: chk_update<6>(msg_f, rs, active, 6); break;
                case 7: chk_update<7>(msg_f, rs, active, 7); break;
                case 8: chk_update<8>(msg_f, rs, active, 8); break;
                default: chk_update<0>(msg_f, rs, active, job.degree); break;
            }
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

static int odd_stride(int count) { return count | 1; }   // frames land in different banks

static size_t bp_smem_bytes(const ldpc_code *c, int F) {
    return sizeof(double) * F * ((size_t) odd_stride(c->E) + 2 * (size_t) odd_stride(c->n)) + sizeof(SlotBlock<8>) +
           (size_t) F * c->n + 32;
}

// One warp-job = 32/F node ranks of one degree; jobs sorted by degree (descending) so that the
// warps of a round do equal work, padded to rounds x warps.
static std::vector<BpJob> make_jobs(const std::vector<BpClass> &classes, int F, int nwarps, int *rounds) {
    std::vector<BpJob> jobs;
    const int per_warp = 32 / F;
    for (const BpClass &cl : classes)
        for (int start = 0; start < cl.count; start += per_warp)
            jobs.push_back(BpJob{(uint16_t) cl.degree, (uint16_t) (cl.first + start),
                                 (uint16_t) std::min(per_warp, cl.count - start), 0});
    std::stable_sort(jobs.begin(), jobs.end(), [](const BpJob &a, const BpJob &b) { return a.degree > b.degree; });
    *rounds = ((int) jobs.size() + nwarps - 1) / nwarps;
    jobs.resize((size_t) *rounds * nwarps, BpJob{0, 0, 0, 0});
    return jobs;
}

static int get_schedule(const ldpc_code *c, int F, int nwarps, BpSchedule *out) {
    std::lock_guard<std::mutex> lock(c->sched_mu);
    auto it = c->bp_sched.find({F, nwarps});
    if (it == c->bp_sched.end()) {
        BpSchedule s;
        std::vector<BpJob> jv = make_jobs(c->var_classes, F, nwarps, &s.rounds_v);
        std::vector<BpJob> jc = make_jobs(c->chk_classes, F, nwarps, &s.rounds_c);
        LDPC_CUDA(dev_malloc((void **) &s.jobs_v, sizeof(BpJob) * std::max<size_t>(jv.size(), 1)));
        LDPC_CUDA(dev_malloc((void **) &s.jobs_c, sizeof(BpJob) * std::max<size_t>(jc.size(), 1)));
        LDPC_CUDA(upload_sync(s.jobs_v, jv.data(), sizeof(BpJob) * jv.size()));
        LDPC_CUDA(upload_sync(s.jobs_c, jc.data(), sizeof(BpJob) * jc.size()));
        it = c->bp_sched.emplace(std::make_pair(F, nwarps), s).first;
    }
    *out = it->second;
    return LDPC_OK;
}

template <int F, int MAXT, int MINB>
static int launch_bp_f(BpParams &p, const ldpc_code *c, int threads, int64_t frames, cudaStream_t stream) {
    BpSchedule s;
    int st = get_schedule(c, F, threads / 32, &s);
    if (st) return st;
    p.jobs_v = s.jobs_v; p.jobs_c = s.jobs_c; p.rounds_v = s.rounds_v; p.rounds_c = s.rounds_c;
    const size_t smem = bp_smem_bytes(c, F);
    auto kernel = bp_kernel<F, MAXT, MINB>;
    LDPC_CUDA(allow_max_dynamic_smem(kernel));
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

// Decoder::decode / exp() for BP: the likelihood-ratio kernel unless the node degrees would force its message
// cap below 50 (the reference's own saturation point is ~45.7); LDPC_BP_KERNEL=log|lr overrides (A/B runs).
// which kernel served the last BP launch of this process (ldpc_debug_last_bp_kernel): 1 likelihood-ratio, 2 log-domain
std::atomic<int> g_last_bp_kernel{0};

int launch_bp(const ldpc_code *c, const FrameIO &fio, int64_t frames, double var, int max_iter, int early_exit,
              unsigned long long *queue, cudaStream_t stream) {
    bool lr = bp_lr_cap(c, nullptr) >= 50.0;
    if (const char *k = getenv("LDPC_BP_KERNEL")) {
        if (k[0] == 'l' && k[1] == 'o') lr = false;
        else if (k[0] == 'l' && k[1] == 'r') lr = true;
    }
    if (lr) {
        const int st = launch_bp_lr(c, fio, frames, var, max_iter, early_exit, queue, stream);
        if (st != LDPC_E_UNSUPPORTED) {               // codes beyond its shared-memory layout take the log-domain kernel
            g_last_bp_kernel = 1;
            return st;
        }
    }
    g_last_bp_kernel = 2;
    return launch_bp_log(c, fio, frames, var, max_iter, early_exit, queue, stream);
}

int launch_bp_log(const ldpc_code *c, const FrameIO &fio, int64_t frames, double var, int max_iter, int early_exit,
                  unsigned long long *queue, cudaStream_t stream) {
    if (frames <= 0) return LDPC_OK;
    if (c->max_row_deg > BP_MAX_DEGREE || c->max_col_deg > BP_MAX_DEGREE)
        return fail(LDPC_E_UNSUPPORTED, "node degree above 64 is not supported by the BP kernel");
    BpParams p;
    KernelIO &io = p.io;
    io.y = fio.y; io.bits = fio.bits; io.ok = fio.ok; io.iters = fio.iters; io.soft = fio.soft;
    io.experiment = fio.experiment; io.cw_source = fio.cw_source; io.seed = fio.seed;
    io.frame_begin = fio.frame_begin; io.words = fio.words; io.n_words = fio.n_words;
    io.counters = fio.counters; io.gen_cols = c->d.gen_cols; io.k = c->k; io.k_words = c->k_words;
    io.frames = frames; io.queue = queue; io.var = var; io.sigma = std::sqrt(var);
    io.n = c->n; io.m = c->m; io.row_ptr = c->d.row_ptr; io.col_idx = c->d.col_idx;
    p.chk_rs = c->d.chk_rs; p.var_rec = c->d.var_rec; p.var_edges = c->d.var_edges;
    p.E = c->E; p.stride_e = odd_stride(c->E); p.stride_n = odd_stride(c->n);
    p.max_iter = max_iter; p.early_exit = early_exit;

    // frames in flight per CTA: enough (check, frame) lanes to fill the CTA's warps, within shared memory
    int threads = 256;
    if (const char *force = getenv("LDPC_BP_THREADS")) {
        const int v = atoi(force) / 32 * 32;
        if (v >= 32 && v <= 512) threads = v;
    }
    int F = 8;
    while (F > 1 && 2 * bp_smem_bytes(c, F) > 227 * 1024) F >>= 1;
    if (frames < 2ll * 148 * F) F = 1;               // tiny batches: spread the frames over the SMs instead
    if (const char *force = getenv("LDPC_BP_F")) {
        const int v = atoi(force);
        if (v == 1 || v == 2 || v == 4 || v == 8) F = v;
    }
    while (F > 1 && bp_smem_bytes(c, F) > 227 * 1024) F >>= 1;
    if (bp_smem_bytes(c, F) > 227 * 1024)
        return fail(LDPC_E_UNSUPPORTED, "BP messages of this code exceed 227 KB of shared memory");
    // register budget: 256-thread CTAs at 2 / 3 / 4 CTAs per SM (128 / 85 / 64 registers per thread)
    int minb = 2;
    if (const char *force = getenv("LDPC_BP_MINB")) minb = atoi(force);
    if (threads > 256) minb = 0;
    switch (F * 8 + minb) {
        case 8 * 8 + 2: return launch_bp_f<8, 256, 2>(p, c, threads, frames, stream);
        case 8 * 8 + 3: return launch_bp_f<8, 256, 3>(p, c, threads, frames, stream);
        case 8 * 8 + 4: return launch_bp_f<8, 256, 4>(p, c, threads, frames, stream);
        case 4 * 8 + 2: return launch_bp_f<4, 256, 2>(p, c, threads, frames, stream);
        case 4 * 8 + 3: return launch_bp_f<4, 256, 3>(p, c, threads, frames, stream);
        case 4 * 8 + 4: return launch_bp_f<4, 256, 4>(p, c, threads, frames, stream);
        case 2 * 8 + 2: return launch_bp_f<2, 256, 2>(p, c, threads, frames, stream);
        case 2 * 8 + 3: return launch_bp_f<2, 256, 3>(p, c, threads, frames, stream);
        case 1 * 8 + 2: return launch_bp_f<1, 256, 2>(p, c, threads, frames, stream);
        case 1 * 8 + 3: return launch_bp_f<1, 256, 3>(p, c, threads, frames, stream);
        default: break;
    }
    switch (F) {
        case 8: return launch_bp_f<8, 512, 1>(p, c, threads, frames, stream);
        case 4: return launch_bp_f<4, 512, 1>(p, c, threads, frames, stream);
        case 2: return launch_bp_f<2, 512, 1>(p, c, threads, frames, stream);
        default: return launch_bp_f<1, 512, 1>(p, c, threads, frames, stream);
    }
}

}  // namespace ldpc
