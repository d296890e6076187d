// Belief propagation (flooding sum-product) in the likelihood-ratio domain (sm_100a)
// -- algo/bp.h:155-222, the same schedule and the same function of the inputs as the
// reference, evaluated without a single transcendental inside the iteration.
//
// Algebra.  The reference sends phi(|t|) = -log tanh(|t|/2) and sign(t) from variables to
// checks and sign * phi(sum phi) back (bp.h:49-57, 77-83).  With L = exp(t):
//     tanh(t/2) = (L - 1) / (L + 1)
//     prod_i tanh(t_i/2) = (Pe - Po) / (Pe + Po)
// where Pe / Po are the sums over all ways of picking the "1" from an even / odd number
// of the factors (L_i + 1) -- positive terms only, so there is no cancellation at any
// magnitude -- and the check-to-variable message, as a likelihood ratio, is
//     x = exp(sign * phi(sum phi)) = (1 + prod tanh) / (1 - prod tanh) = Pe / Po.
// The variable-to-check message is L_j = L_ch * prod_{k != j} x_k with L_ch = exp(llr),
// the posterior is L_j * x_j and the hard decision (estimate <= 0, bp.h:193) is
// posterior <= 1.  Leave-one-out products come from prefix/suffix passes.  Per edge and
// iteration this is ~13 fp64 instructions (11 of them the Pe/Po recurrences and one
// division) instead of the ~48 of the exp / log-ratio form (bp_kernel.cu), and every
// value carries full fp64 relative accuracy, i.e. ~1e-16 absolute in the LLR domain.
// exp() is evaluated once per variable and frame (L_ch), log() once per variable when a
// frame's soft output is written.
//
// Range.  Messages are clamped to exp(+-C1) on the high word (integer min/max); C1 = 100
// when the degrees allow it (products of dv-1 / dc-1 messages must stay inside the
// double range), else smaller; launch_bp falls back to the log-domain kernel when C1
// would drop below 50 (the reference itself saturates at |t| ~ 45.7).
//
// Layout.  A persistent CTA keeps F frames in flight; a lane is a (node, frame) pair with
// the frame index fastest, and all per-frame arrays are stored [element][frame], so the
// F lanes of a node read F consecutive doubles: with F = 16 every 64-bit shared-memory
// access of a half-warp is one conflict-free 128-byte wavefront, whatever the graph.
//   msg   E x F doubles   C->V likelihood ratios x, overwritten in place by the V->C
//                         messages L_j (sign bit = hard decision of the variable) and back
//   lch   n x F doubles   L_ch
//   dec   n x F bytes     hard decisions of the last variable pass
//   post  n x F doubles   posterior likelihood ratios (only when soft output is requested)
// A trip of the main loop is: check pass (which also yields the syndrome of the previous
// variable pass from the sign bits it loads), barrier, frames that converged or ran out
// of iterations are published and their slots refilled, variable pass, barrier.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "bpmath.cuh"
#include "slots.cuh"

namespace ldpc {

constexpr int LR_MAX_DEGREE = 64;

struct BpLrParams {
    KernelIO io;
    const uint32_t *rec_v;      // variable records: [var * F * 8, pos_0 * F * 8, ..., pos_{d-1} * F * 8], padded to 4 words
    const BpLrRun *runs_v, *runs_c;   // per warp: max_runs entries, terminated by degree 0
    int max_runs_v, max_runs_c;
    int E;
    int max_iter, early_exit;
    int clamp_lo, clamp_hi;     // high words of exp(-C1), exp(+C1)
    double llr_cap;
    int chunk;                  // frames claimed from the global queue at a time
    int soft;                   // posterior array present
};

// control words of a trip, double-buffered by trip parity (written by warp 0 during the
// variable phase of trip k for trip k+1)
struct LrCtl {
    unsigned active;    // slots whose messages are valid V->C messages (they take part in the check pass)
    unsigned elig;      // slots with iter >= 1 (the reference tests the syndrome from iteration 1 on, bp.h:195)
    unsigned atmax;     // slots with iter >= max_iter
    unsigned bad;       // slots with an unsatisfied check (accumulated by the check pass)
};

template <int F>
struct LrShared {
    SlotBlock<F> S;
    LrCtl ctl[2];
    unsigned fresh;               // slots whose frame has just been loaded (initial send, bp.h:184)
    unsigned live;                // slots holding a frame
    long long q_next, q_end;      // locally claimed range of the global frame queue
};

__device__ __forceinline__ double ld_f64(const char *p) { return *reinterpret_cast<const double *>(p); }
__device__ __forceinline__ void st_f64(char *p, double v) { *reinterpret_cast<double *>(p) = v; }

__device__ __forceinline__ double clamp_hi_word(double x, int lo, int hi) {
    return __hiloint2double(min(max(__double2hiint(x), lo), hi), __double2loint(x));
}

// node record of a variable of degree D: word 0 = var * F * 8, words 1..D = message slot * F * 8
template <int D>
struct LrRec {
    static constexpr int CAP = D > 0 ? D : LR_MAX_DEGREE;
    static constexpr int WORDS = D > 0 ? ((D + 1 + 3) / 4) * 4 : 4;
    uint32_t w[CAP + 4];
    __device__ __forceinline__ void load(const uint32_t *rec, int d) {
        if (D > 0) {
#pragma unroll
            for (int q = 0; q < WORDS / 4; ++q) {
                const uint4 t = __ldg(reinterpret_cast<const uint4 *>(rec) + q);
                w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
            }
        } else {
            for (int q = 0; q <= d; ++q) w[q] = __ldg(rec + q);
        }
    }
};

// ---- variable node of degree D: VNode::message (bp.h:77-83), estimate (bp.h:85-90), decision (bp.h:193)
template <int D, int FB, bool SOFT>   // FB = F * 8: byte stride between consecutive elements
__device__ __forceinline__ void lr_var_update(char *msg_f, const char *lch_f, char *post_f, uint8_t *dec_f,
                                              const LrRec<D> &r, int d_runtime, int clamp_lo, int clamp_hi) {
    const int d = D > 0 ? D : d_runtime;
    constexpr int CAP = D > 0 ? D : LR_MAX_DEGREE;
    const uint32_t *w = r.w;
    const double lc = ld_f64(lch_f + w[0]);
    double x[CAP], suf[CAP];
#pragma unroll
    for (int j = 0; j < d; ++j) x[j] = ld_f64(msg_f + w[1 + j]);
    suf[d - 1] = 1.0;
#pragma unroll
    for (int j = d - 2; j >= 0; --j) suf[j] = (j == d - 2) ? x[j + 1] : suf[j + 1] * x[j + 1];
    double pre = lc;                           // L_ch * prod_{k < j} x_k
    double lam[CAP];
#pragma unroll
    for (int j = 0; j < d; ++j) {
        lam[j] = (j == d - 1) ? pre : pre * suf[j];
        if (j < d - 1) pre *= x[j];
    }
    const double tot = lam[d - 1] * x[d - 1];  // posterior likelihood ratio
    const bool one = tot <= 1.0;               // estimate <= 0 -> bit 1 (bp.h:193)
    const int hard = one ? (int) 0x80000000 : 0;
#pragma unroll
    for (int j = 0; j < d; ++j) {
        const int hi = min(max(__double2hiint(lam[j]), clamp_lo), clamp_hi) | hard;
        st_f64(msg_f + w[1 + j], __hiloint2double(hi, __double2loint(lam[j])));
    }
    dec_f[w[0] / 8] = (uint8_t) one;
    if (SOFT) st_f64(post_f + w[0], tot);
}

// the initial send (bp.h:184): all C->V messages are zero (CNode::init, bp.h:42-45), i.e. x = 1
template <int D, int FB, bool SOFT>
__device__ __forceinline__ void lr_var_first(char *msg_f, const char *lch_f, char *post_f, uint8_t *dec_f,
                                             const LrRec<D> &r, int d_runtime, int clamp_lo, int clamp_hi) {
    const int d = D > 0 ? D : d_runtime;
    const uint32_t *w = r.w;
    const double lc = ld_f64(lch_f + w[0]);
    const double l = clamp_hi_word(lc, clamp_lo, clamp_hi);
    const int hard = lc <= 1.0 ? (int) 0x80000000 : 0;
    const double m = __hiloint2double(__double2hiint(l) | hard, __double2loint(l));
#pragma unroll
    for (int j = 0; j < d; ++j) st_f64(msg_f + w[1 + j], m);
    dec_f[w[0] / 8] = (uint8_t) (lc <= 1.0);
    if (SOFT) st_f64(post_f + w[0], lc);
}

// one step of the variable pass: 32/F consecutive ranks of degree D, one per lane group
template <int D, int F, bool SOFT>
__device__ __forceinline__ void lr_var_step(char *msg_f, const char *lch_f, char *post_f, uint8_t *dec_f,
                                            const uint32_t *rec, int degree, int node_lane, bool fresh,
                                            int clamp_lo, int clamp_hi) {
    const int stride = D > 0 ? ((D + 1 + 3) / 4) * 4 : ((degree + 1 + 3) / 4) * 4;
    LrRec<D> r;
    r.load(rec + node_lane * stride, degree);
    if (fresh) lr_var_first<D, F * 8, SOFT>(msg_f, lch_f, post_f, dec_f, r, degree, clamp_lo, clamp_hi);
    else lr_var_update<D, F * 8, SOFT>(msg_f, lch_f, post_f, dec_f, r, degree, clamp_lo, clamp_hi);
}

// ---- check node of degree D: CNode::message (bp.h:49-57); returns the parity of the decisions
template <int D, int FB>
__device__ __forceinline__ int lr_chk_update(char *edge, int d_runtime, int clamp_hi) {
    const int d = D > 0 ? D : d_runtime;
    constexpr int CAP = D > 0 ? D : LR_MAX_DEGREE;
    double a[CAP], se[CAP], so[CAP];
    int par = 0;
#pragma unroll
    for (int j = 0; j < d; ++j) {
        const double m = ld_f64(edge + j * FB);
        const int hi = __double2hiint(m);
        par ^= hi;
        a[j] = __hiloint2double(hi & 0x7fffffff, __double2loint(m));
    }
    if (d == 1) {                               // no other variable: phi(0) = +inf in the reference; here the cap
        st_f64(edge, __hiloint2double(clamp_hi, 0));
        return (unsigned) par >> 31;
    }
    // suffix pairs S_j = (Pe, Po) over the inputs i > j
    se[d - 1] = 1.0;
    so[d - 1] = 0.0;
#pragma unroll
    for (int j = d - 2; j >= 0; --j) {
        if (j == d - 2) { se[j] = a[j + 1]; so[j] = 1.0; }
        else { se[j] = __fma_rn(se[j + 1], a[j + 1], so[j + 1]); so[j] = __fma_rn(so[j + 1], a[j + 1], se[j + 1]); }
    }
    double pe = 1.0, po = 0.0;                  // prefix pair over the inputs i < j
#pragma unroll
    for (int j = 0; j < d; ++j) {
        double ev, od;
        if (j == 0) { ev = se[0]; od = so[0]; }
        else if (j == d - 1) { ev = pe; od = po; }
        else if (j == 1) { ev = __fma_rn(pe, se[j], so[j]); od = __fma_rn(pe, so[j], se[j]); }   // (pe, po) = (a0, 1)
        else { ev = __fma_rn(pe, se[j], po * so[j]); od = __fma_rn(pe, so[j], po * se[j]); }
        st_f64(edge + j * FB, div_pos(ev, od));
        if (j == 0) { pe = a[0]; po = 1.0; }
        else if (j < d - 1) { const double ne = __fma_rn(pe, a[j], po); po = __fma_rn(po, a[j], pe); pe = ne; }
    }
    return (unsigned) par >> 31;
}

template <int F, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) bp_lr_kernel(const BpLrParams p) {
    extern __shared__ __align__(16) double smem[];
    const KernelIO &io = p.io;
    const int n = io.n;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int warp = tid >> 5, nwarps = nt >> 5, lane = tid & 31;
    constexpr int FB = F * 8;
    const int f_lane = lane % F, node_lane = lane / F;

    char *msg = reinterpret_cast<char *>(smem);                          // E x F doubles
    char *lch = msg + (size_t) p.E * FB;                                 // n x F doubles
    char *post = lch + (size_t) n * FB;                                  // n x F doubles (soft output only)
    uint8_t *dec = reinterpret_cast<uint8_t *>(post + (p.soft ? (size_t) n * FB : 0));   // n x F bytes
    uint8_t *cw = dec + (size_t) n * F;                                  // F x n bytes (experiment mode)
    LrShared<F> *L = reinterpret_cast<LrShared<F> *>(
        (reinterpret_cast<uintptr_t>(cw + (io.experiment ? (size_t) n * F : 0)) + 15) & ~(uintptr_t) 15);
    SlotBlock<F> *S = &L->S;

    slots_init(S);
    if (tid == 0) {
        L->ctl[0] = LrCtl{0u, 0u, 0u, 0u};
        L->ctl[1] = LrCtl{0u, 0u, 0u, 0u};
        L->fresh = 0u;
        L->live = 0u;
        L->q_next = L->q_end = 0;
        S->alive = F;
    }
    __syncthreads();

    char *msg_f = msg + f_lane * 8;
    const char *lch_f = lch + f_lane * 8;
    char *post_f = post + f_lane * 8;
    const bool soft = p.soft != 0;
    // this warp's steps of the two passes, heaviest degree first, terminated by degree 0
    const BpLrRun *my_steps_v = p.runs_v + (size_t) warp * p.max_runs_v;
    const BpLrRun *my_steps_c = p.runs_c + (size_t) warp * p.max_runs_c;
    uint8_t *dec_f = dec + f_lane;

    for (unsigned trip = 0;; ++trip) {
        LrCtl *ctl = &L->ctl[trip & 1];
        // ---- check pass (also the syndrome of the previous variable pass)
        const unsigned active = ctl->active;
        if (active) {
            int bad = 0;
            if ((active >> f_lane) & 1u) {
                const BpLrRun *q = my_steps_c;
                BpLrRun e = *q;
                while (e.degree > 8) {
                    const BpLrRun nx = *++q;
                    if (node_lane < e.nodes)
                        bad |= lr_chk_update<0, FB>(msg_f + e.first + node_lane * e.degree * FB, e.degree, p.clamp_hi);
                    e = nx;
                }
#define LDPC_CHK_STEPS(D)                                                                                   \
    while (e.degree == D) {                                                                                 \
        const BpLrRun nx = *++q;                                                                            \
        if (node_lane < e.nodes) bad |= lr_chk_update<D, FB>(msg_f + e.first + node_lane * D * FB, D, p.clamp_hi); \
        e = nx;                                                                                             \
    }
                LDPC_CHK_STEPS(8) LDPC_CHK_STEPS(7) LDPC_CHK_STEPS(6) LDPC_CHK_STEPS(5)
                LDPC_CHK_STEPS(4) LDPC_CHK_STEPS(3) LDPC_CHK_STEPS(2) LDPC_CHK_STEPS(1)
#undef LDPC_CHK_STEPS
            }
            unsigned b = __ballot_sync(0xffffffffu, bad);
#pragma unroll
            for (int s = F; s < 32; s <<= 1) b |= b >> s;      // fold the node lanes: bit f = slot f
            b &= (F == 32) ? 0xffffffffu : ((1u << F) - 1u);
            if (lane == 0 && b) atomicOr(&ctl->bad, b);
        }
        __syncthreads();

        // ---- publish finished frames, refill their slots
        const unsigned okmask = ctl->elig & ~ctl->bad & active;          // syndrome vanished (iteration >= 1)
        const unsigned finmask = (ctl->atmax | (p.early_exit ? okmask : 0u)) & active;
        if (finmask || trip == 0) {
            for (int f = 0; f < F; ++f) {
                if (!((finmask >> f) & 1u)) continue;
                const int ok = (okmask >> f) & 1u;
                const uint8_t *df = dec + f;
                const char *pf = post + f * 8;
                slot_finish<F>(io, S, f, ok, ok, ok, S->iter[f], cw, [&](int i) { return (int) df[(size_t) i * F]; },
                               [&](int i) {
                                   const double t = ld_f64(pf + (size_t) i * FB);
                                   return log_pos(fmin(fmax(t, 1e-300), 1e300));
                               });
            }
            __syncthreads();
            if (warp == 0) {
                // lanes f < F own slot f: empty slots take the next frame of the locally claimed range
                const unsigned live_before = L->live & ~finmask;
                const bool want = lane < F && !((live_before >> lane) & 1u) && S->state[lane] != SLOT_DEAD;
                const unsigned wmask = __ballot_sync(0xffffffffu, want);
                const int need = __popc(wmask), rank = __popc(wmask & ((1u << lane) - 1u));
                const long long next = L->q_next, end = L->q_end;
                __syncwarp();
                const long long left = end - next;
                long long got = 0, amt = 0;
                if (need > left) {                 // claim a new chunk (at least what is missing) from the global queue
                    amt = max((long long) p.chunk, need - left);
                    if (lane == 0) got = (long long) atomicAdd(io.queue, (unsigned long long) amt);
                    got = __shfl_sync(0xffffffffu, got, 0);
                }
                if (want) {
                    const long long fr = rank < left ? next + rank : got + (rank - left);
                    if (fr < io.frames) { S->frame[lane] = fr; S->iter[lane] = 0; S->hamming[lane] = 0; S->state[lane] = SLOT_NEW; }
                    else S->state[lane] = SLOT_DEAD;
                }
                if (lane == 0) {
                    if (need > left) { L->q_next = got + (need - left); L->q_end = got + amt; }
                    else L->q_next = next + need;
                }
                __syncwarp();
                const int st = lane < F ? S->state[lane] : SLOT_DEAD;
                const unsigned fresh = __ballot_sync(0xffffffffu, st == SLOT_NEW);
                const unsigned dead = __ballot_sync(0xffffffffu, lane < F && st == SLOT_DEAD);
                if (lane == 0) {
                    L->fresh = fresh;
                    L->live = live_before | fresh;
                    S->alive = F - __popc(dead);
                }
            }
            __syncthreads();
            if (S->alive == 0) break;
            const unsigned fresh = L->fresh;
            if (fresh) {
                // L_ch = exp(llr), llr clamped to +-llr_cap; variables without edges keep decision / posterior of the channel
                slots_load<F>(io, S, fresh, nullptr, 0, cw, [&](int i, int f, double l) {
                    const double lc = exp_signed(fmin(fmax(l, -p.llr_cap), p.llr_cap));
                    st_f64(lch + (size_t) i * FB + f * 8, lc);
                    dec[(size_t) i * F + f] = (uint8_t) (lc <= 1.0);
                    if (p.soft) st_f64(post + (size_t) i * FB + f * 8, lc);
                });
            }
        }

        // ---- variable pass
        const unsigned live = L->live, fresh = L->fresh;
        if (warp == 0) {
            // control words of the next trip
            int it = 0;
            if (lane < F && ((live >> lane) & 1u)) {
                it = ((fresh >> lane) & 1u) ? 0 : S->iter[lane] + 1;
                S->iter[lane] = it;
                S->state[lane] = SLOT_ACTIVE;
            }
            const unsigned elig = __ballot_sync(0xffffffffu, it >= 1) & live;
            const unsigned atmax = __ballot_sync(0xffffffffu, it >= p.max_iter) & live;
            if (lane == 0) L->ctl[(trip + 1) & 1] = LrCtl{live, elig, atmax, 0u};
        }
        if ((live >> f_lane) & 1u) {
            const bool is_fresh = (fresh >> f_lane) & 1u;
            const BpLrRun *q = my_steps_v;
            BpLrRun e = *q;
#define LDPC_VAR_STEP(D)                                                                                          \
    do {                                                                                                          \
        const BpLrRun nx = *++q;                                                                                  \
        if (node_lane < e.nodes) {                                                                                \
            if (soft) lr_var_step<D, F, true>(msg_f, lch_f, post_f, dec_f, p.rec_v + e.first, e.degree, node_lane, \
                                              is_fresh, p.clamp_lo, p.clamp_hi);                                  \
            else lr_var_step<D, F, false>(msg_f, lch_f, post_f, dec_f, p.rec_v + e.first, e.degree, node_lane,    \
                                          is_fresh, p.clamp_lo, p.clamp_hi);                                      \
        }                                                                                                         \
        e = nx;                                                                                                   \
    } while (0)
            while (e.degree > 8) LDPC_VAR_STEP(0);
            while (e.degree == 8) LDPC_VAR_STEP(8);
            while (e.degree == 7) LDPC_VAR_STEP(7);
            while (e.degree == 6) LDPC_VAR_STEP(6);
            while (e.degree == 5) LDPC_VAR_STEP(5);
            while (e.degree == 4) LDPC_VAR_STEP(4);
            while (e.degree == 3) LDPC_VAR_STEP(3);
            while (e.degree == 2) LDPC_VAR_STEP(2);
            while (e.degree == 1) LDPC_VAR_STEP(1);
#undef LDPC_VAR_STEP
        }
        __syncthreads();
        if (tid == 0 && fresh) L->fresh = 0u;      // read again only after the next barrier
    }
    slots_flush(io, S);
}

// ---------------------------------------------------------------- host side

static size_t lr_smem_bytes(const ldpc_code *c, int F, bool soft, bool experiment) {
    return (size_t) F * 8 * ((size_t) c->E + (size_t) c->n * (soft ? 2 : 1)) + (size_t) F * c->n * (experiment ? 2 : 1) +
           16 + sizeof(SlotBlock<32>) + 256;
}

// message cap C1: products of (dv - 1) messages times L_ch, and of (dc - 1) messages, must stay inside the double range
double bp_lr_cap(const ldpc_code *c, double *llr_cap_out) {
    const double llr_cap = 100.0;
    double c1 = 100.0;
    if (c->max_col_deg > 1) c1 = std::min(c1, (700.0 - llr_cap) / (c->max_col_deg - 1));
    if (c->max_row_deg > 1) c1 = std::min(c1, (700.0 - 0.7 * c->max_row_deg) / (c->max_row_deg - 1));
    if (llr_cap_out) *llr_cap_out = llr_cap;
    return c1;
}

// Deals the steps (32/F consecutive node ranks of one degree class = one node per lane group) to the warps
// round-robin, heaviest classes first, so the warps of a pass finish together.  Per warp: its steps in that
// order, terminated by a degree-0 entry.
template <typename FirstOf>
static std::vector<BpLrRun> make_runs(const std::vector<BpClass> &classes, int F, int nwarps, int *max_runs,
                                      FirstOf first_of) {
    const int G = 32 / F;
    std::vector<int> order(classes.size());
    for (size_t k = 0; k < classes.size(); ++k) order[k] = (int) k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return classes[a].degree > classes[b].degree; });
    std::vector<std::vector<BpLrRun>> per_warp(nwarps);
    int s_global = 0;
    for (int k : order)
        for (int n0 = 0; n0 < classes[k].count; n0 += G, ++s_global)
            per_warp[s_global % nwarps].push_back(BpLrRun{(uint16_t) classes[k].degree,
                                                          (uint16_t) std::min(G, classes[k].count - n0), first_of(k, n0)});
    size_t mr = 1;
    for (auto &rw : per_warp) mr = std::max(mr, rw.size() + 1);
    std::vector<BpLrRun> flat((size_t) nwarps * mr, BpLrRun{0, 0, 0});
    for (int w = 0; w < nwarps; ++w) std::copy(per_warp[w].begin(), per_warp[w].end(), flat.begin() + (size_t) w * mr);
    *max_runs = (int) mr;
    return flat;
}

template <typename T>
static int upload_vec(T **dst, const std::vector<T> &src) {
    LDPC_CUDA(cudaMalloc((void **) dst, sizeof(T) * std::max<size_t>(src.size(), 1)));
    if (!src.empty()) LDPC_CUDA(cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
    return LDPC_OK;
}

static int get_lr_schedule(const ldpc_code *c, int F, int nwarps, BpLrSchedule *out) {
    std::lock_guard<std::mutex> lock(c->sched_mu);
    auto it = c->bp_lr_sched.find({F, nwarps});
    if (it == c->bp_lr_sched.end()) {
        BpLrSchedule s;
        // message slots in check-rank order: the checks of one degree class are stored back to back
        std::vector<int> slot_of_edge(c->E, 0), class_slot0;
        {
            int slot = 0;
            for (const BpClass &cl : c->chk_classes) {
                class_slot0.push_back(slot);
                for (int k = 0; k < cl.count; ++k) {
                    const int chk = c->chk_order[cl.first + k];
                    for (int e = c->row_ptr[chk]; e < c->row_ptr[chk + 1]; ++e) slot_of_edge[e] = slot++;
                }
            }
        }
        // variable records in rank order (classes of equal degree are adjacent, code.cu)
        std::vector<uint32_t> rec, first_v;
        for (const BpClass &cl : c->var_classes) {
            first_v.push_back((uint32_t) rec.size());
            const int stride = ((cl.degree + 1 + 3) / 4) * 4;
            for (int k = 0; k < cl.count; ++k) {
                const int v = c->var_order[cl.first + k];
                const size_t base = rec.size();
                rec.resize(base + stride, 0u);
                rec[base] = (uint32_t) v * F * 8;
                for (int j = 0; j < cl.degree; ++j)
                    rec[base + 1 + j] = (uint32_t) slot_of_edge[c->csc_edge[c->col_ptr[v] + j]] * F * 8;
            }
        }
        std::vector<BpLrRun> rv = make_runs(c->var_classes, F, nwarps, &s.max_runs_v, [&](int cls, int node0) {
            return first_v[cls] + (uint32_t) node0 * (uint32_t) (((c->var_classes[cls].degree + 1 + 3) / 4) * 4);
        });
        std::vector<BpLrRun> rc = make_runs(c->chk_classes, F, nwarps, &s.max_runs_c, [&](int cls, int node0) {
            return (uint32_t) (class_slot0[cls] + node0 * c->chk_classes[cls].degree) * F * 8;
        });
        int st;
        if ((st = upload_vec(&s.rec_v, rec))) return st;
        if ((st = upload_vec(&s.runs_v, rv))) return st;
        if ((st = upload_vec(&s.runs_c, rc))) return st;
        it = c->bp_lr_sched.emplace(std::make_pair(F, nwarps), s).first;
    }
    *out = it->second;
    return LDPC_OK;
}

template <int F, int MAXT>
static int launch_lr_ft(BpLrParams &p, const ldpc_code *c, int threads, size_t smem, int64_t frames, cudaStream_t stream) {
    BpLrSchedule s;
    int st = get_lr_schedule(c, F, threads / 32, &s);
    if (st) return st;
    p.rec_v = s.rec_v; p.runs_v = s.runs_v; p.runs_c = s.runs_c;
    p.max_runs_v = s.max_runs_v; p.max_runs_c = s.max_runs_c;
    auto kernel = bp_lr_kernel<F, MAXT>;
    LDPC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int per_sm = 0, sms = 0;
    LDPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    LDPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    if (per_sm < 1) return fail(LDPC_E_UNSUPPORTED, "BP state of this code does not fit on one SM");
    const long long want = (frames + F - 1) / F;
    const long long grid = std::min<long long>((long long) per_sm * sms, want);
    // frames are claimed from the global queue in chunks; small enough that the tail stays balanced
    p.chunk = (int) std::max<long long>(1, std::min<long long>(F, frames / (grid * 4 * F) * F));
    kernel<<<(unsigned) grid, threads, smem, stream>>>(p);
    LDPC_CUDA(cudaGetLastError());
    return LDPC_OK;
}

// register budget by CTA size: 128 / 85 / 64 registers per thread
template <int F>
static int launch_lr_f(BpLrParams &p, const ldpc_code *c, int threads, size_t smem, int64_t frames, cudaStream_t stream) {
    if (threads <= 512) return launch_lr_ft<F, 512>(p, c, threads, smem, frames, stream);
    if (threads <= 768) return launch_lr_ft<F, 768>(p, c, threads, smem, frames, stream);
    return launch_lr_ft<F, 1024>(p, c, threads, smem, frames, stream);
}

int launch_bp_lr(const ldpc_code *c, const FrameIO &fio, int64_t frames, double var, int max_iter, int early_exit,
                 unsigned long long *queue, cudaStream_t stream) {
    if (frames <= 0) return LDPC_OK;
    BpLrParams p;
    KernelIO &io = p.io;
    io.y = fio.y; io.bits = fio.bits; io.ok = fio.ok; io.iters = fio.iters; io.soft = fio.soft;
    io.experiment = fio.experiment; io.cw_source = fio.cw_source; io.seed = fio.seed;
    io.frame_begin = fio.frame_begin; io.words = fio.words; io.n_words = fio.n_words;
    io.counters = fio.counters; io.gen_cols = c->d.gen_cols; io.k = c->k; io.k_words = c->k_words;
    io.frames = frames; io.queue = queue; io.var = var; io.sigma = std::sqrt(var);
    io.n = c->n; io.m = c->m; io.row_ptr = c->d.row_ptr; io.col_idx = c->d.col_idx;
    p.E = c->E; p.max_iter = max_iter; p.early_exit = early_exit;
    p.soft = fio.soft != nullptr;
    const double c1 = bp_lr_cap(c, &p.llr_cap);
    {
        const double lo = std::exp(-c1), hi = std::exp(c1);
        uint64_t blo, bhi;
        memcpy(&blo, &lo, 8);
        memcpy(&bhi, &hi, 8);
        p.clamp_lo = (int) (blo >> 32);
        p.clamp_hi = (int) (bhi >> 32);
    }
    const bool exp_mode = fio.experiment != 0;
    int F = 16;
    if (const char *force = getenv("LDPC_BP_F")) {
        const int v = atoi(force);
        if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) F = v;
    } else {
        while (F > 1 && frames < 2ll * 148 * F) F >>= 1;      // small batches: spread the frames over the SMs
    }
    while (F > 1 && lr_smem_bytes(c, F, p.soft, exp_mode) > 227 * 1024) F >>= 1;
    const size_t smem = lr_smem_bytes(c, F, p.soft, exp_mode);
    if (smem > 227 * 1024) return fail(LDPC_E_UNSUPPORTED, "BP messages of this code exceed 227 KB of shared memory");
    int threads = (long long) c->n * F >= 2048 ? 512 : 256;
    if (const char *force = getenv("LDPC_BP_THREADS")) {
        const int v = atoi(force) / 32 * 32;
        if (v >= 32 && v <= 1024) threads = v;
    }
    switch (F) {
        case 16: return launch_lr_f<16>(p, c, threads, smem, frames, stream);
        case 8: return launch_lr_f<8>(p, c, threads, smem, frames, stream);
        case 4: return launch_lr_f<4>(p, c, threads, smem, frames, stream);
        case 2: return launch_lr_f<2>(p, c, threads, smem, frames, stream);
        default: return launch_lr_f<1>(p, c, threads, smem, frames, stream);
    }
}

}  // namespace ldpc
