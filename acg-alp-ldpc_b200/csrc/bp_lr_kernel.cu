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
// Layout.  A persistent CTA keeps F frames in flight; a lane is a (node, frame PAIR) and all
// per-frame arrays are stored [element][frame] with the frame index fastest, so a lane moves
// its two frames with one 128-bit access and the F/2 lanes of a node cover F consecutive
// doubles: with F = 16 every quarter-warp access is one conflict-free 128-byte wavefront,
// whatever the graph.  Two frames per lane halve the address arithmetic, the shared-memory
// instructions and the per-node bookkeeping per frame and give every thread two independent
// dependency chains.
//   msg   E x F doubles   C->V likelihood ratios x, overwritten in place by the V->C
//                         messages L_j (sign bit = hard decision of the variable) and back
//   lch   n x F doubles   L_ch
//   dec   n x F bytes     hard decisions of the last variable pass
//   post  n x F doubles   posterior likelihood ratios (only when soft output is requested)
// A trip of the main loop is: check pass (which also yields the syndrome of the previous
// variable pass from the sign bits it loads), barrier, frames that converged or ran out
// of iterations are published and their slots refilled, variable pass, barrier.
#include <algorithm>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "bpmath.cuh"
#include "slots_team.cuh"
#include "smem_ptx.cuh"

namespace ldpc {

constexpr int LR_MAX_UNROLLED = 8;           // larger degrees take the generic routines (kernel variant WIDE)
constexpr int LR_MAX_DEGREE = 16;            // bp_lr_cap() >= 50 implies degrees <= 14

struct BpLrParams {
    KernelIO io;
    const uint32_t *rec_v;      // variable records (words): [var * F * 8, slot_0 * F * 8, ..., slot_{d-1} * F * 8], padded to 4 words
    const uint32_t *steps;      // step words of one team: check pass (steps_c rows of `gt` words), then variable pass (steps_v rows)
    const uint16_t *var_store;  // variable index -> storage index of its L_ch / decision / posterior (rank order)
    int rec_words;              // total words of rec_v
    int steps_c, steps_v;       // rows of the two step tables (the last row of each is all zero)
    int gt;                     // threads per team
    int E;                      // message slots per frame (including the idle ones of padded classes)
    // byte offsets inside a team's region of shared memory, and the size of a region
    uint32_t off_lch, off_post, off_dec, off_cw, off_ctl, team_bytes;
    uint32_t off_teams;         // start of the teams' regions (behind the shared tables)
    int max_iter, early_exit;
    int clamp_lo, clamp_hi;     // high words of exp(-C1), exp(+C1)
    double llr_cap;
    int chunk;                  // frames claimed from the global queue at a time
};

// A step = one warp instruction's worth of nodes of one degree: 64/F consecutive ranks, one per lane group.  Every
// thread of a team has its own word per step (row k of the table at [k * gt + thread]):
//   bits 0-17  byte offset of the lane's work: check pass -- the first message of its node; variable pass -- the node's
//              record in the record table
//   bit 23     the lane is idle in this step (the class does not fill the step)
//   bits 24-31 degree (0 = end of the list); the lists are sorted by descending degree
constexpr uint32_t LR_STEP_OFF = 0x3ffffu, LR_STEP_IDLE = 1u << 23;
__host__ __device__ __forceinline__ uint32_t lr_step_word(uint32_t off, bool idle, int degree) {
    return off | (idle ? LR_STEP_IDLE : 0u) | ((uint32_t) degree << 24);
}

// control words of a trip, double-buffered by trip parity (written by warp 0 of the team during the variable phase of
// trip k for trip k+1)
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

// two frames of one element; the hot loops address shared memory through 32-bit shared-window addresses (smem_ptx.cuh):
// one register per base address, register + immediate per access
struct P2 {
    double a, b;
};
__device__ __forceinline__ P2 ld_p2(uint32_t a) {
    const double2 t = lds_f64x2(a);
    return P2{t.x, t.y};
}
__device__ __forceinline__ void st_p2(uint32_t a, P2 v) { sts_f64x2(a, v.a, v.b); }
__device__ __forceinline__ P2 operator*(P2 x, P2 y) { return P2{x.a * y.a, x.b * y.b}; }
__device__ __forceinline__ P2 fma2(P2 x, P2 y, P2 z) { return P2{__fma_rn(x.a, y.a, z.a), __fma_rn(x.b, y.b, z.b)}; }
__device__ __forceinline__ P2 div2(P2 x, P2 y) { return P2{div_pos(x.a, y.a), div_pos(x.b, y.b)}; }

// clamp to [exp(-C1), exp(C1)] on the high word and put the variable's decision into the sign bit, in two integer
// instructions: t = clamp(hi - lo, 0, range) is one VIADDMNMX.RELU, and (t + lo) | sign == t + (lo + sign) because
// t + lo < 2^31.  neg_lo = -lo, range = hi - lo, lo_sign = lo + (decision << 31).
__device__ __forceinline__ double clamp_sign(double x, int neg_lo, int range, int lo_sign) {
    return __hiloint2double(__viaddmin_s32_relu(__double2hiint(x), neg_lo, range) + lo_sign, __double2loint(x));
}

// shared-window base addresses of a lane (team region + the lane's frame pair)
struct LrAddr {
    uint32_t msg, dec;          // messages (L_ch: + off_lch, posteriors: + off_post, both uniform), decisions
    uint32_t off_lch, off_post;
    uint32_t rec;               // the record table
    int neg_lo, range, lo;      // clamp constants: -hi word of exp(-C1), hi(exp(C1)) - hi(exp(-C1)), hi word of exp(-C1)
    int lo_neg;                 // lo with the sign bit set (decision 1)
};

// ---- variable node of degree D, two frames: VNode::message (bp.h:77-83), estimate (bp.h:85-90), decision (bp.h:193)
// rec = shared-window address of the node's record
template <int D, bool SOFT>
__device__ __forceinline__ void lr_var_update(const LrAddr &A, uint32_t rec, int d_runtime) {
    const int d = D > 0 ? D : d_runtime;
    constexpr int CAP = D > 0 ? D : LR_MAX_DEGREE;
    uint32_t w[CAP + 4];
    if (D > 0) {
#pragma unroll
        for (int q = 0; q < (D + 1 + 3) / 4; ++q) {
            const uint4 t = lds_u32x4(rec + 16 * q);
            w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
        }
    } else {
        for (int q = 0; q <= d; ++q) w[q] = lds_u32(rec + 4 * q);
    }
    const P2 lc = ld_p2(A.msg + A.off_lch + w[0]);
    P2 x[CAP], suf[CAP], lam[CAP];
#pragma unroll
    for (int j = 0; j < d; ++j) x[j] = ld_p2(A.msg + w[1 + j]);
    suf[d - 1] = P2{1.0, 1.0};
#pragma unroll
    for (int j = d - 2; j >= 0; --j) suf[j] = (j == d - 2) ? x[j + 1] : suf[j + 1] * x[j + 1];
    P2 pre = lc;                                // L_ch * prod_{k < j} x_k
#pragma unroll
    for (int j = 0; j < d; ++j) {
        lam[j] = (j == d - 1) ? pre : pre * suf[j];
        if (j < d - 1) pre = pre * x[j];
    }
    const P2 tot = lam[d - 1] * x[d - 1];       // posterior likelihood ratios
    const bool one_a = tot.a <= 1.0, one_b = tot.b <= 1.0;   // estimate <= 0 -> bit 1 (bp.h:193)
    const int sa = one_a ? A.lo_neg : A.lo, sb = one_b ? A.lo_neg : A.lo;     // one select each: both constants live in registers
    const int neg_lo = A.neg_lo, range = A.range;
#pragma unroll
    for (int j = 0; j < d; ++j)
        st_p2(A.msg + w[1 + j], P2{clamp_sign(lam[j].a, neg_lo, range, sa), clamp_sign(lam[j].b, neg_lo, range, sb)});
    uint32_t bits2 = one_b ? 0x100u : 0u;
    if (one_a) bits2 += 1u;
    sts_u16(A.dec + (w[0] >> 3), bits2);
    if (SOFT) st_p2(A.msg + A.off_post + w[0], tot);
}

// ---- check node of degree D, two frames: CNode::message (bp.h:49-57); returns the parities of the decisions
// (bit 0 / bit 1 = first / second frame of the pair).  edge = shared-window address of the first message of the node,
// lane's frame pair.
template <int D, int FB>
__device__ __forceinline__ int lr_chk_update(uint32_t edge, int d_runtime, int clamp_hi) {
    const int d = D > 0 ? D : d_runtime;
    constexpr int CAP = D > 0 ? D : LR_MAX_DEGREE;
    P2 a[CAP], se[CAP], so[CAP];
    int par_a = 0, par_b = 0;
#pragma unroll
    for (int j = 0; j < d; ++j) {
        const P2 m = ld_p2(edge + j * FB);
        par_a ^= __double2hiint(m.a);
        par_b ^= __double2hiint(m.b);
        a[j] = P2{fabs(m.a), fabs(m.b)};
    }
    const int par = (int) ((unsigned) par_a >> 31) | (int) (((unsigned) par_b >> 31) << 1);
    if (d == 1) {                               // no other variable: phi(0) = +inf in the reference; here the cap
        const double cap = __hiloint2double(clamp_hi, 0);
        st_p2(edge, P2{cap, cap});
        return par;
    }
    const P2 one{1.0, 1.0};
    // suffix pairs S_j = (Pe, Po) over the inputs i > j
    se[d - 1] = one;
    so[d - 1] = P2{0.0, 0.0};
#pragma unroll
    for (int j = d - 2; j >= 0; --j) {
        if (j == d - 2) { se[j] = a[j + 1]; so[j] = one; }
        else { se[j] = fma2(se[j + 1], a[j + 1], so[j + 1]); so[j] = fma2(so[j + 1], a[j + 1], se[j + 1]); }
    }
    P2 pe = one, po{0.0, 0.0};                  // prefix pair over the inputs i < j
#pragma unroll
    for (int j = 0; j < d; ++j) {
        P2 ev, od;
        if (j == 0) { ev = se[0]; od = so[0]; }
        else if (j == d - 1) { ev = pe; od = po; }
        else if (j == 1) { ev = fma2(pe, se[j], so[j]); od = fma2(pe, so[j], se[j]); }   // (pe, po) = (a0, 1)
        else { ev = fma2(pe, se[j], po * so[j]); od = fma2(pe, so[j], po * se[j]); }
        st_p2(edge + j * FB, div2(ev, od));
        if (j == 0) { pe = a[0]; po = one; }
        else if (j < d - 1) { const P2 ne = fma2(pe, a[j], po); po = fma2(po, a[j], pe); pe = ne; }
    }
    return par;
}

// Several TEAMS per CTA: a team is `gt` threads with F frames in flight, its own region of shared memory and its own
// named barrier; the read-only tables (variable records, step lists) are held once per CTA.  One CTA per SM, as many
// teams as shared memory and registers allow: the teams never wait for one another, so their FP64-bound check passes and
// shared-memory-bound variable passes overlap.
// SOFT: posterior likelihood ratios are kept (soft output requested).  WIDE: degrees above 8 occur (generic routines).
template <int F, int MAXT, bool SOFT, bool WIDE>
__global__ void __launch_bounds__(MAXT, 1) bp_lr_kernel(const BpLrParams p) {
    extern __shared__ __align__(16) double smem[];
    const KernelIO &io = p.io;
    constexpr int FB = F * 8;                   // bytes between consecutive elements
    constexpr int LPN = F / 2;                  // lanes per node
    const int gt = p.gt;
    const int team = threadIdx.x / gt;
    Team T;
    T.tid = threadIdx.x - team * gt;
    T.nt = gt;
    T.bar = 1 + team;
    const int tid = T.tid, nt = gt;
    const int warp = tid >> 5, lane = tid & 31;
    const int pair = lane % LPN;

    // shared tables
    uint32_t *rec = reinterpret_cast<uint32_t *>(smem);                  // variable records
    uint32_t *steps = rec + p.rec_words;                                 // step words: check pass, variable pass
    for (int i = threadIdx.x; i < p.rec_words; i += blockDim.x) rec[i] = p.rec_v[i];
    for (int i = threadIdx.x; i < (p.steps_c + p.steps_v) * gt; i += blockDim.x) steps[i] = p.steps[i];

    // this team's region
    char *base = reinterpret_cast<char *>(smem) + p.off_teams + (size_t) team * p.team_bytes;
    char *msg = base;                                                    // E x F doubles
    char *lch = base + p.off_lch;                                        // n x F doubles
    char *post = base + p.off_post;                                      // n x F doubles (soft output only)
    uint8_t *dec = reinterpret_cast<uint8_t *>(base + p.off_dec);        // n x F bytes
    uint8_t *cw = reinterpret_cast<uint8_t *>(base + p.off_cw);          // F x n bytes (experiment mode)
    LrShared<F> *L = reinterpret_cast<LrShared<F> *>(base + p.off_ctl);
    SlotBlock<F> *S = &L->S;

    // shared-window addresses of this lane's share
    // (They take a round trip through shared memory: a value ptxas can derive from the thread index is recomputed at
    // every step instead of being kept -- fourteen instructions per step; a loaded value stays in its register.)
    LrAddr A;
    {
        const uint32_t scratch = smem_addr(msg) + tid * 32;
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(scratch), "r"(smem_addr(msg) + pair * 16),
                     "r"(smem_addr(dec) + pair * 2), "r"(smem_addr(steps) + tid * 4), "r"(2u * pair) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(scratch + 16), "r"(smem_addr(rec)), "r"(-p.clamp_lo),
                     "r"(p.clamp_hi - p.clamp_lo), "r"(p.clamp_lo) : "memory");
    }
    T.sync();
    const uint4 kept = lds_u32x4(smem_addr(msg) + tid * 32), kept2 = lds_u32x4(smem_addr(msg) + tid * 32 + 16);
    T.sync();
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(smem_addr(msg) + tid * 8), "r"((uint32_t) gt * 4),
                 "r"(p.clamp_lo + (int) 0x80000000) : "memory");
    T.sync();
    const uint32_t row = lds_u32(smem_addr(msg) + tid * 8);      // bytes between the rows of a step table
    A.lo_neg = (int) lds_u32(smem_addr(msg) + tid * 8 + 4);
    T.sync();
    A.msg = kept.x;
    A.dec = kept.y;
    A.off_lch = p.off_lch;
    A.off_post = p.off_post;
    A.rec = kept2.x;
    A.neg_lo = (int) kept2.y;
    A.range = (int) kept2.z;
    A.lo = (int) kept2.w;
    const uint32_t a_steps_c = kept.z, a_steps_v = a_steps_c + p.steps_c * gt * 4;
    const unsigned pair_shift = kept.w, pair_bits = 3u << pair_shift;

    team_slots_init(T, S);                      // (after the round trip: its scratch may reach into the control block)
    if (tid == 0) {
        L->ctl[0] = LrCtl{0u, 0u, 0u, 0u};
        L->ctl[1] = LrCtl{0u, 0u, 0u, 0u};
        L->fresh = 0u;
        L->live = 0u;
        L->q_next = L->q_end = 0;
        S->alive = F;
    }
    __syncthreads();                            // the only CTA-wide barrier: tables are in place

    for (unsigned trip = 0;; ++trip) {
        LrCtl *ctl = &L->ctl[trip & 1];
        // ---- check pass (also the syndrome of the previous variable pass)
        const unsigned active = ctl->active;
        if (active) {
            unsigned bad = 0;
            if (active & pair_bits) {
                // the step word of a busy lane is offset + (degree << 24): inside the loop of degree D the address is one add
                uint32_t q = a_steps_c;
                uint32_t e = lds_u32(q);
                if (WIDE) {
                    while (e >= ((uint32_t) (LR_MAX_UNROLLED + 1) << 24)) {
                        const uint32_t nx = lds_u32(q += row);
                        if (!(e & LR_STEP_IDLE)) bad |= (unsigned) lr_chk_update<0, FB>(A.msg + (e & LR_STEP_OFF), (int) (e >> 24), p.clamp_hi);
                        e = nx;
                    }
                }
#define LDPC_CHK_STEPS(D)                                                                                   \
    while (e >= ((uint32_t) (D) << 24)) {                                                                   \
        const uint32_t nx = lds_u32(q += row);                                                              \
        if (!(e & LR_STEP_IDLE)) bad |= (unsigned) lr_chk_update<D, FB>(A.msg + e - ((uint32_t) (D) << 24), D, p.clamp_hi); \
        e = nx;                                                                                             \
    }
                LDPC_CHK_STEPS(8) LDPC_CHK_STEPS(7) LDPC_CHK_STEPS(6) LDPC_CHK_STEPS(5)
                LDPC_CHK_STEPS(4) LDPC_CHK_STEPS(3) LDPC_CHK_STEPS(2) LDPC_CHK_STEPS(1)
#undef LDPC_CHK_STEPS
            }
            const unsigned b = __reduce_or_sync(0xffffffffu, bad << pair_shift);
            if (lane == 0 && b) atomicOr(&ctl->bad, b);
        }
        T.sync();

        // ---- publish finished frames, refill their slots
        const unsigned okmask = ctl->elig & ~ctl->bad & active;          // syndrome vanished (iteration >= 1)
        const unsigned finmask = (ctl->atmax | (p.early_exit ? okmask : 0u)) & active;
        if (finmask || trip == 0) {
            if (finmask)
                team_slots_finish_all<F>(T, io, S, finmask, okmask, cw,
                                         [&](int i, int f) { return (int) dec[(size_t) p.var_store[i] * F + f]; },
                                         [&](int i, int f) {
                                             const double t = ld_f64(post + (size_t) p.var_store[i] * FB + f * 8);
                                             return log_pos(fmin(fmax(t, 1e-300), 1e300));
                                         });
            T.sync();
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
            T.sync();
            if (S->alive == 0) break;
            const unsigned fresh = L->fresh;
            if (fresh) {
                // all C->V messages of a new frame are zero (CNode::init, bp.h:42-45): x = 1, which makes its first
                // variable pass the reference's initial send (bp.h:184)
                for (int i = tid; i < p.E * F; i += nt)
                    if ((fresh >> (i % F)) & 1u) st_f64(msg + (size_t) i * 8, 1.0);
                // L_ch = exp(llr), llr clamped to +-llr_cap; variables without edges keep decision / posterior of the channel
                team_slots_load_all<F>(T, io, S, fresh, cw, [&](int i, int f, double l) {
                    const double lc = exp_signed(fmin(fmax(l, -p.llr_cap), p.llr_cap));
                    const size_t st = p.var_store[i];
                    st_f64(lch + st * FB + f * 8, lc);
                    dec[st * F + f] = (uint8_t) (lc <= 1.0);
                    if (SOFT) st_f64(post + st * FB + f * 8, lc);
                });
            }
        }

        // ---- variable pass
        const unsigned live = L->live;
        if (warp == 0) {
            // control words of the next trip
            const unsigned fresh = L->fresh;
            int it = 0;
            if (lane < F && ((live >> lane) & 1u)) {
                it = ((fresh >> lane) & 1u) ? 0 : S->iter[lane] + 1;
                S->iter[lane] = it;
                S->state[lane] = SLOT_ACTIVE;
            }
            const unsigned elig = __ballot_sync(0xffffffffu, it >= 1) & live;
            const unsigned atmax = __ballot_sync(0xffffffffu, it >= p.max_iter) & live;
            if (lane == 0) {
                L->ctl[(trip + 1) & 1] = LrCtl{live, elig, atmax, 0u};
                L->fresh = 0u;
            }
        }
        if (live & pair_bits) {
            uint32_t q = a_steps_v;
            uint32_t e = lds_u32(q);
            if (WIDE) {
                while (e >= ((uint32_t) (LR_MAX_UNROLLED + 1) << 24)) {
                    const uint32_t nx = lds_u32(q += row);
                    if (!(e & LR_STEP_IDLE)) lr_var_update<0, SOFT>(A, A.rec + (e & LR_STEP_OFF), (int) (e >> 24));
                    e = nx;
                }
            }
#define LDPC_VAR_STEPS(D)                                                                                         \
    while (e >= ((uint32_t) (D) << 24)) {                                                                         \
        const uint32_t nx = lds_u32(q += row);                                                                    \
        if (!(e & LR_STEP_IDLE)) lr_var_update<D, SOFT>(A, A.rec + e - ((uint32_t) (D) << 24), D);           \
        e = nx;                                                                                                   \
    }
            LDPC_VAR_STEPS(8) LDPC_VAR_STEPS(7) LDPC_VAR_STEPS(6) LDPC_VAR_STEPS(5)
            LDPC_VAR_STEPS(4) LDPC_VAR_STEPS(3) LDPC_VAR_STEPS(2) LDPC_VAR_STEPS(1)
#undef LDPC_VAR_STEPS
        }
        T.sync();
    }
    team_slots_flush(T, io, S);
}

// ---------------------------------------------------------------- host side

static size_t up16(size_t x) { return (x + 15) & ~(size_t) 15; }

// a team's region of shared memory: messages, L_ch, posteriors (soft output), decisions, codewords (experiment mode),
// control block; fills the offsets of `p` when given
static size_t lr_team_bytes(const ldpc_code *c, int n_slots, int F, bool soft, bool experiment, BpLrParams *p) {
    size_t off = (size_t) n_slots * F * 8;
    const size_t off_lch = off; off += (size_t) c->n * F * 8;
    const size_t off_post = off; off += soft ? (size_t) c->n * F * 8 : 0;
    const size_t off_dec = off; off += (size_t) c->n * F;
    const size_t off_cw = off; off += experiment ? (size_t) c->n * F : 0;
    const size_t off_ctl = up16(off); off = std::max<size_t>(up16(off_ctl + sizeof(LrShared<16>)), 768 * 32);   // >= the kernel's start-up scratch
    if (p) {
        p->off_lch = (uint32_t) off_lch; p->off_post = (uint32_t) off_post; p->off_dec = (uint32_t) off_dec;
        p->off_cw = (uint32_t) off_cw; p->off_ctl = (uint32_t) off_ctl; p->team_bytes = (uint32_t) off;
    }
    return off;
}

constexpr size_t LR_SMEM_MAX = 227 * 1024;

// message cap C1: products of (dv - 1) messages times L_ch, and of (dc - 1) messages, must stay inside the double range
double bp_lr_cap(const ldpc_code *c, double *llr_cap_out) {
    const double llr_cap = 100.0;
    double c1 = 100.0;
    if (c->max_col_deg > 1) c1 = std::min(c1, (700.0 - llr_cap) / (c->max_col_deg - 1));
    if (c->max_row_deg > 1) c1 = std::min(c1, (700.0 - 0.7 * c->max_row_deg) / (c->max_row_deg - 1));
    if (llr_cap_out) *llr_cap_out = llr_cap;
    return c1;
}

static int rec_words_of(const ldpc_code *c) {
    int words = 0;
    for (const BpClass &cl : c->var_classes) words += cl.count * (((cl.degree + 1 + 3) / 4) * 4);
    return words;
}

// Deals the steps (64/F consecutive node ranks of one degree class = one node per lane group) to the warps of a team so
// that the warps of a pass finish together: longest-processing-time first onto the least loaded warp, with the executed
// instructions of a step as its cost.  Per warp: its steps by descending degree (the kernel runs one loop per degree).
// The result is one word per (step row, thread of the team): row k at [k * gt + thread], the row after a warp's last
// step is zero.  offset_of(class, node rank in class) = byte offset of the node's work (check pass: its first message;
// variable pass: its record); the lane's frame pair is part of the lane's base address.
template <typename OffsetOf>
static std::vector<uint32_t> deal_steps(const std::vector<BpClass> &classes, int F, int nwarps, bool check_pass, OffsetOf offset_of,
                                        int *rows_out) {
    const int G = 64 / F, LPN = F / 2;
    struct Step { int cls, node0, degree, cost; };
    std::vector<Step> all;
    for (size_t k = 0; k < classes.size(); ++k)
        for (int n0 = 0; n0 < classes[k].count; n0 += G) {
            const int d = classes[k].degree;
            all.push_back(Step{(int) k, n0, d, check_pass ? 12 + 33 * d : 32 + 18 * d});
        }
    std::stable_sort(all.begin(), all.end(), [](const Step &a, const Step &b) { return a.cost > b.cost; });
    std::vector<std::vector<Step>> mine(nwarps);
    std::vector<long> load(nwarps, 0);
    // warp 0 of a team also writes the control words of the next trip in front of its variable pass; a head start for
    // the other warps changes nothing measurable (0 / 100 / 200 / 300: 37.8 / 37.1 / 37.3 / 37.9 ms on H05), so none by default
    if (!check_pass)
        if (const char *e = getenv("LDPC_BP_W0_HANDICAP")) load[0] = atoi(e);
    for (size_t i = 0; i < all.size(); ++i) {
        const int w = (int) (std::min_element(load.begin(), load.end()) - load.begin());
        mine[w].push_back(all[i]);
        load[w] += all[i].cost;
    }
    size_t rows = 1;
    for (int w = 0; w < nwarps; ++w) {
        std::stable_sort(mine[w].begin(), mine[w].end(), [](const Step &a, const Step &b) { return a.degree > b.degree; });
        rows = std::max(rows, mine[w].size() + 1);
    }
    const int gt = nwarps * 32;
    std::vector<uint32_t> words(rows * gt, 0u);
    for (int w = 0; w < nwarps; ++w)
        for (size_t k = 0; k < mine[w].size(); ++k) {
            const Step &st = mine[w][k];
            for (int lane = 0; lane < 32; ++lane) {
                const int node_lane = lane / LPN;
                const bool idle = st.node0 + node_lane >= classes[st.cls].count;
                const uint32_t off = idle ? 0u : offset_of(st.cls, st.node0 + node_lane);
                words[k * gt + w * 32 + lane] = lr_step_word(off, idle, st.degree);
            }
        }
    *rows_out = (int) rows;
    return words;
}

template <typename T>
static int upload_vec(T **dst, const std::vector<T> &src) {
    LDPC_CUDA(dev_malloc((void **) dst, sizeof(T) * std::max<size_t>(src.size(), 1)));
    if (!src.empty()) LDPC_CUDA(upload_sync(*dst, src.data(), sizeof(T) * src.size()));
    return LDPC_OK;
}

// Balanced 2-colouring of the edges of the Tanner graph: every variable and every check gets as many edges of colour
// 0 as of colour 1 (+-1 for odd degrees).  Euler partition: nodes of odd degree are joined to a dummy node of the
// other side (the two dummies to each other if their degrees are odd), every node then has even degree, the edge set
// splits into closed trails, and a trail of a bipartite graph has even length, so colouring it alternately gives every
// node visit one edge of each colour.
static void lr_balanced_colouring(const ldpc_code *c, std::vector<int> &colour) {
    const int n = c->n, m = c->m, E = c->E;
    const int dummy_var = n + m, dummy_chk = n + m + 1, nodes = n + m + 2;
    struct Ed { int a, b; };
    std::vector<Ed> edges;
    for (int r = 0; r < m; ++r)
        for (int e = c->row_ptr[r]; e < c->row_ptr[r + 1]; ++e) edges.push_back(Ed{c->col_idx[e], n + r});
    int odd = 0;
    for (int v = 0; v < n; ++v)
        if ((c->col_ptr[v + 1] - c->col_ptr[v]) & 1) { edges.push_back(Ed{v, dummy_chk}); ++odd; }
    for (int r = 0; r < m; ++r)
        if ((c->row_ptr[r + 1] - c->row_ptr[r]) & 1) edges.push_back(Ed{dummy_var, n + r});
    if (odd & 1) edges.push_back(Ed{dummy_var, dummy_chk});
    std::vector<std::vector<int>> adj(nodes);
    for (int e = 0; e < (int) edges.size(); ++e) { adj[edges[e].a].push_back(e); adj[edges[e].b].push_back(e); }
    std::vector<char> used(edges.size(), 0);
    std::vector<size_t> next(nodes, 0);
    std::vector<int> col(edges.size(), 0);
    for (int start = 0; start < nodes; ++start) {
        for (;;) {
            while (next[start] < adj[start].size() && used[adj[start][next[start]]]) ++next[start];
            if (next[start] >= adj[start].size()) break;
            int at = start, k = 0;          // walk a closed trail from `start`
            for (;;) {
                while (next[at] < adj[at].size() && used[adj[at][next[at]]]) ++next[at];
                if (next[at] >= adj[at].size()) break;      // only possible back at `start`
                const int e = adj[at][next[at]];
                used[e] = 1;
                col[e] = k++ & 1;
                at = edges[e].a == at ? edges[e].b : edges[e].a;
            }
        }
    }
    colour.assign(E, 0);
    for (int e = 0; e < E; ++e) colour[e] = col[e];
}

static int get_lr_schedule(const ldpc_code *c, int F, int nwarps, BpLrSchedule *out) {
    std::lock_guard<std::mutex> lock(c->sched_mu);
    auto it = c->bp_lr_sched.find({F, nwarps});
    if (it == c->bp_lr_sched.end()) {
        BpLrSchedule s;
        // Message slots in check-rank order: the checks of one degree class are stored back to back.  With fewer than
        // 16 frames per CTA a node's frames fill only part of a 128-byte line and the two (F = 8) nodes of a quarter-warp
        // must use opposite halves of the line, i.e. slots of opposite parity, at every access:
        //  * a class of even degree gets the stride degree + 1 (one idle slot per check), so consecutive checks -- the
        //    lane groups of a step -- start on slots of alternating parity;
        //  * the edges get a balanced 2-colouring (every node has as many edges of one colour as of the other, +-1:
        //    Euler partition of the Tanner graph); a check lists its edges with alternating colours, so colour = slot
        //    parity; variables are paired so that one has as many even slots as the other has odd ones and list their
        //    edges in opposite parity order.
        // The order of the nodes inside a degree class and of the edges inside a node only permutes independent work
        // and the order of exact-in-any-order products' roundings; results stay deterministic.
        const int pad_even = F < 16 ? 1 : 0;
        std::vector<int> chk_ord(c->chk_order), var_ord(c->var_order);
        std::vector<int> colour(c->E, 0);
        if (pad_even) lr_balanced_colouring(c, colour);
        std::vector<int> slot_of_edge(c->E, 0), class_slot0;
        int n_slots = 0;
        {
            int slot = 0;
            for (const BpClass &cl : c->chk_classes) {
                const int stride = cl.degree | pad_even;
                class_slot0.push_back(slot);
                if (pad_even) {
                    // order the checks of the class so that the majority colour of check k is the parity of its first slot
                    std::vector<int> maj[2], out;
                    for (int k = 0; k < cl.count; ++k) {
                        const int chk = c->chk_order[cl.first + k];
                        int ones = 0;
                        for (int e = c->row_ptr[chk]; e < c->row_ptr[chk + 1]; ++e) ones += colour[e];
                        maj[2 * ones > cl.degree ? 1 : 0].push_back(chk);
                    }
                    size_t i0 = 0, i1 = 0;
                    for (int k = 0; k < cl.count; ++k) {
                        const int want = (slot + k * stride) & 1;
                        std::vector<int> &pref = maj[want], &other = maj[want ^ 1];
                        size_t &ip = want ? i1 : i0, &io_ = want ? i0 : i1;
                        if (ip < pref.size()) out.push_back(pref[ip++]);
                        else out.push_back(other[io_++]);
                    }
                    for (int k = 0; k < cl.count; ++k) chk_ord[cl.first + k] = out[k];
                }
                for (int k = 0; k < cl.count; ++k) {
                    const int chk = chk_ord[cl.first + k];
                    const int e0 = c->row_ptr[chk], d = cl.degree;
                    if (pad_even) {
                        // positions of parity (slot & 1) take the edges of that colour first
                        std::vector<int> by_col[2], pos_of(d, -1);
                        for (int j = 0; j < d; ++j) by_col[colour[e0 + j]].push_back(j);
                        std::vector<int> rest;
                        size_t used[2] = {0, 0};
                        for (int q = 0; q < d; ++q) {
                            const int par = (slot + q) & 1;
                            if (used[par] < by_col[par].size()) pos_of[by_col[par][used[par]++]] = q;
                            else rest.push_back(q);
                        }
                        size_t r = 0;
                        for (int j = 0; j < d; ++j)
                            if (pos_of[j] < 0) pos_of[j] = rest[r++];
                        for (int j = 0; j < d; ++j) slot_of_edge[e0 + j] = slot + pos_of[j];
                    } else {
                        for (int j = 0; j < d; ++j) slot_of_edge[e0 + j] = slot + j;
                    }
                    slot += stride;
                }
            }
            n_slots = std::max(slot, 1);
        }
        // variable order inside a class: pairs (2i, 2i+1) with complementary numbers of even slots
        std::vector<std::vector<int>> var_edges(c->n);      // the edges of a variable in record order
        for (const BpClass &cl : c->var_classes) {
            std::vector<int> members(c->var_order.begin() + cl.first, c->var_order.begin() + cl.first + cl.count);
            auto evens = [&](int v) {
                int k = 0;
                for (int q = c->col_ptr[v]; q < c->col_ptr[v + 1]; ++q) k += !(slot_of_edge[c->csc_edge[q]] & 1);
                return k;
            };
            std::vector<int> order;
            if (pad_even) {
                std::vector<std::vector<int>> bucket(cl.degree + 1);
                for (int v : members) bucket[evens(v)].push_back(v);
                std::vector<int> left;
                for (int t = 0; t <= cl.degree; ++t) {
                    std::vector<int> &a = bucket[t], &b = bucket[cl.degree - t];
                    if (t > cl.degree - t) break;
                    if (t == cl.degree - t) {
                        while (a.size() >= 2) { order.push_back(a.back()); a.pop_back(); order.push_back(a.back()); a.pop_back(); }
                    } else {
                        while (!a.empty() && !b.empty()) { order.push_back(a.back()); a.pop_back(); order.push_back(b.back()); b.pop_back(); }
                    }
                }
                for (auto &bk : bucket) for (int v : bk) left.push_back(v);
                order.insert(order.end(), left.begin(), left.end());
            } else {
                order = members;
            }
            for (int k = 0; k < cl.count; ++k) {
                const int v = order[k];
                var_ord[cl.first + k] = v;
                std::vector<int> ev, od;
                for (int q = c->col_ptr[v]; q < c->col_ptr[v + 1]; ++q) {
                    const int e = c->csc_edge[q];
                    ((slot_of_edge[e] & 1) ? od : ev).push_back(e);
                }
                std::vector<int> &first = (pad_even && (k & 1)) ? od : ev, &second = (pad_even && (k & 1)) ? ev : od;
                if (pad_even) {
                    var_edges[v] = first;
                    var_edges[v].insert(var_edges[v].end(), second.begin(), second.end());
                } else {
                    for (int q = c->col_ptr[v]; q < c->col_ptr[v + 1]; ++q) var_edges[v].push_back(c->csc_edge[q]);
                }
            }
        }
        // Fewer than 16 frames per CTA: a 16-byte access is served quarter-warp by quarter-warp (8 lanes = 16/F nodes of one
        // step), and a node's frames fill only F/16 of a 128-byte line, so the 16/F nodes of a quarter-warp must sit in
        // different positions of the line (slot mod 16/F) at every access or the wavefront is replayed (34 % of all
        // shared-memory wavefronts of the (3,6)-1008 code at F = 2, profiles/r01_bp_lr_1008_ncu.txt).  The check pass is
        // conflict-free by the odd class strides; for the variable pass the parity construction above is the starting point
        // (at F = 8 it leaves 38 of 430 accesses of H05 replayed, 4 after the polish) and a deterministic annealing pass permutes the
        // variables inside their degree classes, the edges inside a variable's record and the edge positions inside a
        // check -- all of it reorders independent work only.  Cost = replays (largest multiplicity - 1 per access).
        if (F <= 8 && !getenv("LDPC_BP_NO_ANNEAL")) {
            const int Q = 16 / F;                           // nodes per quarter-warp = positions per line
            std::vector<int> cls_of(c->n, -1), pos_of(c->n, 0), chk_of_edge(c->E, 0);
            for (size_t k = 0; k < c->var_classes.size(); ++k)
                for (int i = 0; i < c->var_classes[k].count; ++i) {
                    cls_of[var_ord[c->var_classes[k].first + i]] = (int) k;
                    pos_of[var_ord[c->var_classes[k].first + i]] = i;
                }
            for (int r = 0; r < c->m; ++r)
                for (int e = c->row_ptr[r]; e < c->row_ptr[r + 1]; ++e) chk_of_edge[e] = r;
            auto group_cost = [&](int k, int g) {
                const BpClass &cl = c->var_classes[k];
                int cost = 0;
                for (int j = 0; j < cl.degree; ++j) {
                    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, worst = 1;
                    for (int i = g * Q; i < std::min(cl.count, g * Q + Q); ++i)
                        worst = std::max(worst, ++cnt[slot_of_edge[var_edges[var_ord[cl.first + i]][j]] % Q]);
                    cost += worst - 1;
                }
                return cost;
            };
            long now = 0;
            for (size_t k = 0; k < c->var_classes.size(); ++k)
                for (int g = 0; g * Q < c->var_classes[k].count; ++g) now += group_cost((int) k, g);
            const long before = now;
            long best = now;
            std::vector<int> best_ord(var_ord), best_slot(slot_of_edge);
            std::vector<std::vector<int>> best_edges(var_edges);
            uint64_t rs = 0x9E3779B97F4A7C15ull;            // xorshift: the layout must not depend on the C++ library
            auto rnd = [&]() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t) (rs >> 11); };
            int moves_per_node = 300;
            if (const char *e = getenv("LDPC_BP_ANNEAL_MOVES")) moves_per_node = std::max(1, atoi(e));
            const int moves = now > 0 ? moves_per_node * c->n : 0;
            // checks of one class (equal degree): first slot of each, to let two checks trade their slot ranges
            std::vector<int> chk_cls(c->m, -1), chk_base(c->m, 0);
            for (size_t k = 0; k < c->chk_classes.size(); ++k)
                for (int i = 0; i < c->chk_classes[k].count; ++i) {
                    const int r = chk_ord[c->chk_classes[k].first + i];
                    chk_cls[r] = (int) k;
                    int lo = 1 << 30;
                    for (int e = c->row_ptr[r]; e < c->row_ptr[r + 1]; ++e) lo = std::min(lo, slot_of_edge[e]);
                    chk_base[r] = lo;
                }
            std::vector<int> touched, gk, gg;
            for (int it = 0; it < moves && best > 0; ++it) {
                const double temp = 0.5 * std::pow(0.02, (double) it / moves);
                const uint32_t kind = rnd() % 4u;
                int v1 = -1, v2 = -1, a = 0, b = 0, e1 = 0, e2 = 0, r1 = -1, r2 = -1;
                touched.clear();
                if (kind == 0) {                            // two variables of one class trade places
                    v1 = (int) (rnd() % (uint32_t) c->n);
                    if (cls_of[v1] < 0) continue;
                    const BpClass &cl = c->var_classes[cls_of[v1]];
                    v2 = var_ord[cl.first + (int) (rnd() % (uint32_t) cl.count)];
                    if (v1 == v2) continue;
                    touched = {v1, v2};
                } else if (kind == 1) {                     // two edges of one variable trade places in its record
                    v1 = (int) (rnd() % (uint32_t) c->n);
                    const int d = (int) var_edges[v1].size();
                    if (cls_of[v1] < 0 || d < 2) continue;
                    a = (int) (rnd() % (uint32_t) d); b = (int) (rnd() % (uint32_t) d);
                    if (a == b) continue;
                    touched = {v1};
                } else if (kind == 2) {                     // two edges of one check trade slots
                    const int r = (int) (rnd() % (uint32_t) c->m), d = c->row_ptr[r + 1] - c->row_ptr[r];
                    if (d < 2) continue;
                    e1 = c->row_ptr[r] + (int) (rnd() % (uint32_t) d); e2 = c->row_ptr[r] + (int) (rnd() % (uint32_t) d);
                    if (e1 == e2) continue;
                    touched = {c->col_idx[e1], c->col_idx[e2]};
                } else {                                    // two checks of one class trade their slot ranges
                    r1 = (int) (rnd() % (uint32_t) c->m);
                    if (chk_cls[r1] < 0) continue;
                    const BpClass &cl = c->chk_classes[chk_cls[r1]];
                    r2 = chk_ord[cl.first + (int) (rnd() % (uint32_t) cl.count)];
                    if (r1 == r2 || ((chk_base[r1] ^ chk_base[r2]) & 1)) continue;      // keep the parity of the first slot
                    for (int r : {r1, r2})
                        for (int e = c->row_ptr[r]; e < c->row_ptr[r + 1]; ++e) touched.push_back(c->col_idx[e]);
                }
                auto apply = [&]() {
                    if (kind == 0) {
                        const int k = cls_of[v1], p1 = pos_of[v1], p2 = pos_of[v2], f0 = c->var_classes[k].first;
                        std::swap(var_ord[f0 + p1], var_ord[f0 + p2]);
                        pos_of[v1] = p2; pos_of[v2] = p1;
                    } else if (kind == 1) {
                        std::swap(var_edges[v1][a], var_edges[v1][b]);
                    } else if (kind == 2) {
                        std::swap(slot_of_edge[e1], slot_of_edge[e2]);
                    } else {
                        const int shift = chk_base[r2] - chk_base[r1];
                        for (int e = c->row_ptr[r1]; e < c->row_ptr[r1 + 1]; ++e) slot_of_edge[e] += shift;
                        for (int e = c->row_ptr[r2]; e < c->row_ptr[r2 + 1]; ++e) slot_of_edge[e] -= shift;
                        std::swap(chk_base[r1], chk_base[r2]);
                    }
                };
                // groups touched (a trade of places keeps the set of groups)
                gk.clear(); gg.clear();
                for (int v : touched) {
                    const int k = cls_of[v], g = pos_of[v] / Q;
                    bool seen = false;
                    for (size_t i = 0; i < gk.size(); ++i) seen |= gk[i] == k && gg[i] == g;
                    if (!seen) { gk.push_back(k); gg.push_back(g); }
                }
                int c0 = 0, c1 = 0;
                for (size_t i = 0; i < gk.size(); ++i) c0 += group_cost(gk[i], gg[i]);
                apply();
                for (size_t i = 0; i < gk.size(); ++i) c1 += group_cost(gk[i], gg[i]);
                const int delta = c1 - c0;
                if (delta > 0 && (rnd() & 0xffffff) / 16777216.0 >= std::exp(-delta / temp)) {
                    apply();                                // rejected: undo
                } else {
                    now += delta;
                    if (now < best) { best = now; best_ord = var_ord; best_slot = slot_of_edge; best_edges = var_edges; }
                }
            }
            var_ord = best_ord; slot_of_edge = best_slot; var_edges = best_edges;
            if (getenv("LDPC_BP_LAYOUT_STATS"))
                fprintf(stderr, "bp layout F=%d: variable-pass replays per iteration %ld -> %ld (%d moves)\n", F, before, best, moves);
        }
        // L_ch / decision / posterior storage in variable-rank order (neighbouring lane groups -> neighbouring rows);
        // variables without edges follow
        std::vector<uint16_t> var_store(c->n, 0);
        {
            std::vector<char> ranked(c->n, 0);
            int next = 0;
            for (int v : var_ord) { var_store[v] = (uint16_t) next++; ranked[v] = 1; }
            for (int v = 0; v < c->n; ++v)
                if (!ranked[v]) var_store[v] = (uint16_t) next++;
        }
        // variable records in rank order (classes of equal degree are adjacent, code.cu)
        std::vector<uint32_t> rec, first_v;
        for (const BpClass &cl : c->var_classes) {
            first_v.push_back((uint32_t) rec.size() * 4);     // byte offset
            const int stride = ((cl.degree + 1 + 3) / 4) * 4;
            for (int k = 0; k < cl.count; ++k) {
                const int v = var_ord[cl.first + k];
                const size_t base = rec.size();
                rec.resize(base + stride, 0u);
                rec[base] = (uint32_t) var_store[v] * F * 8;
                for (int j = 0; j < cl.degree; ++j) rec[base + 1 + j] = (uint32_t) slot_of_edge[var_edges[v][j]] * F * 8;
            }
        }
        // layout statistics: paired accesses (lane groups 2i, 2i+1 of a class, same edge position) whose slots have
        // equal parity, i.e. replayed wavefronts at F = 8
        int clash_v = 0, pairs_v = 0, clash_c = 0, pairs_c = 0;
        for (const BpClass &cl : c->var_classes)
            for (int k = 0; k + 1 < cl.count; k += 2)
                for (int j = 0; j < cl.degree; ++j, ++pairs_v)
                    clash_v += !((slot_of_edge[var_edges[var_ord[cl.first + k]][j]] ^ slot_of_edge[var_edges[var_ord[cl.first + k + 1]][j]]) & 1);
        for (size_t ci = 0; ci < c->chk_classes.size(); ++ci) {
            const BpClass &cl = c->chk_classes[ci];
            for (int k = 0; k + 1 < cl.count; k += 2, pairs_c += cl.degree)
                clash_c += ((cl.degree | pad_even) & 1) ? 0 : cl.degree;
        }
        if (getenv("LDPC_BP_LAYOUT_STATS"))
            fprintf(stderr, "bp layout F=%d: %d of %d paired variable-pass and %d of %d paired check-pass accesses share a half line\n",
                    F, clash_v, pairs_v, clash_c, pairs_c);
        if (rec.size() * 4 >= (1u << 18) || (size_t) n_slots * F * 8 >= (1u << 18))
            return fail(LDPC_E_UNSUPPORTED, "code too large for the 18-bit step offsets of the BP kernel");
        int rows_c = 0, rows_v = 0;
        std::vector<uint32_t> sc = deal_steps(c->chk_classes, F, nwarps, true, [&](int cls, int node) {
            return (uint32_t) (class_slot0[cls] + node * (c->chk_classes[cls].degree | pad_even)) * F * 8;
        }, &rows_c);
        std::vector<uint32_t> sv = deal_steps(c->var_classes, F, nwarps, false, [&](int cls, int node) {
            return first_v[cls] + (uint32_t) node * (uint32_t) (((c->var_classes[cls].degree + 1 + 3) / 4) * 16);
        }, &rows_v);
        std::vector<uint32_t> steps(sc);
        steps.insert(steps.end(), sv.begin(), sv.end());
        s.rec_words = (int) rec.size();
        s.n_slots = n_slots;
        s.pad_even = pad_even;
        s.clash_v = clash_v; s.pairs_v = pairs_v; s.clash_c = clash_c; s.pairs_c = pairs_c;
        s.steps_c = rows_c;
        s.steps_v = rows_v;
        int st;
        if ((st = upload_vec(&s.var_store, var_store))) return st;
        if ((st = upload_vec(&s.rec_v, rec))) return st;
        if ((st = upload_vec(&s.steps, steps))) return st;
        it = c->bp_lr_sched.emplace(std::make_pair(F, nwarps), s).first;
    }
    *out = it->second;
    return LDPC_OK;
}

using LrKernel = void (*)(const BpLrParams);

template <int F, int MAXT>
static LrKernel lr_kernel_ft(bool soft) { return soft ? bp_lr_kernel<F, MAXT, true, false> : bp_lr_kernel<F, MAXT, false, false>; }

// register budget by CTA size: 128 registers per thread up to 512 threads, 102 up to 640 (the launcher never asks for
// more); the generic-degree routines only at 128
template <int F>
static LrKernel lr_kernel_f(int threads, bool soft, bool wide) {
    if (wide) return soft ? bp_lr_kernel<F, 512, true, true> : bp_lr_kernel<F, 512, false, true>;
    if (threads <= 512) return lr_kernel_ft<F, 512>(soft);
    return lr_kernel_ft<F, 640>(soft);
}

static LrKernel lr_kernel(int F, int threads, bool soft, bool wide) {
    switch (F) {
        case 16: return lr_kernel_f<16>(threads, soft, wide);
        case 8: return lr_kernel_f<8>(threads, soft, wide);
        case 4: return lr_kernel_f<4>(threads, soft, wide);
        default: return lr_kernel_f<2>(threads, soft, wide);
    }
}

// testing hook: statistics of the shared-memory layout for F frames per team (ldpc_debug_bp_layout)
int bp_lr_layout_stats(const ldpc_code *c, int F, int32_t out[6]) {
    if (F != 2 && F != 4 && F != 8 && F != 16) return fail(LDPC_E_INVALID, "F must be 2, 4, 8 or 16");
    BpLrSchedule s;
    LDPC_CUDA(cudaSetDevice(c->device));
    int st = get_lr_schedule(c, F, 8, &s);
    if (st) return st;
    out[0] = s.n_slots; out[1] = s.pad_even; out[2] = s.clash_v; out[3] = s.pairs_v; out[4] = s.clash_c; out[5] = s.pairs_c;
    return LDPC_OK;
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
    p.max_iter = max_iter; p.early_exit = early_exit;
    const bool soft = fio.soft != nullptr, exp_mode = fio.experiment != 0;
    const bool wide = std::max(c->max_row_deg, c->max_col_deg) > LR_MAX_UNROLLED;
    const double c1 = bp_lr_cap(c, &p.llr_cap);
    {
        const double lo = std::exp(-c1), hi = std::exp(c1);
        uint64_t blo, bhi;
        memcpy(&blo, &lo, 8);
        memcpy(&bhi, &hi, 8);
        p.clamp_lo = (int) (blo >> 32);
        p.clamp_hi = (int) (bhi >> 32);
    }
    // Launch shape: ONE CTA per SM made of independent teams (F frames, `gt` threads, a named barrier and a region of
    // shared memory each) that share the read-only tables.  Several small teams beat one large one: the FP64-bound check
    // pass of one overlaps the shared-memory-bound variable pass of another, and a small team refills its frame slots
    // sooner (profiles/r01_bp_lr_sweep.txt, when teams were CTAs: H05, fixed iterations: one CTA of 16 frames 50.5 ms, two
    // of 8 45.6 ms, four of 4 45.6 ms).  Holding the tables once per SM makes room for a fifth team of 4 frames on the
    // 160 x 280 codes (20 frames and 20 warps per SM instead of 16).
    const int rec_words = rec_words_of(c);
    const int max_threads = wide ? 512 : 640;            // registers: 128 (generic-degree routines) / 102 per thread
    auto gt_of = [&](int f) {
        const int lanes = std::max(c->n, c->m) * (f / 2);
        int gt = std::min(f <= 4 ? (f == 4 ? 128 : 256) : (f == 8 ? 256 : 512), std::max(64, (int) (lanes / 2.9 + 16) / 32 * 32));
        if (const char *force = getenv("LDPC_BP_THREADS")) {
            const int v = atoi(force) / 32 * 32;
            if (v >= 32 && v <= 768) gt = v;
        }
        return std::min(gt, max_threads);
    };
    auto teams_of = [&](int f) {          // teams of f frames that fit an SM (tables estimated from above: 24 step rows)
        const size_t tables = (size_t) 4 * (rec_words + 24 * gt_of(f));
        const size_t team = lr_team_bytes(c, c->E + c->m, f, soft, exp_mode, nullptr);
        if (tables + team > LR_SMEM_MAX) return 0;
        return (int) std::min<size_t>(std::min<size_t>(15, max_threads / gt_of(f)), (LR_SMEM_MAX - tables) / team);
    };
    int F = 8;
    if (const char *force = getenv("LDPC_BP_F")) {
        const int v = atoi(force);
        if (v == 2 || v == 4 || v == 8 || v == 16) F = v;
    } else {
        if (teams_of(4) >= 4) F = 4;
        while (F > 2 && frames < 2ll * 148 * F) F >>= 1;      // small batches: spread the frames over the SMs
        while (F > 2 && teams_of(F) < 2) F >>= 1;
    }
    while (F > 2 && teams_of(F) < 1) F >>= 1;
    const int gt = gt_of(F);
    BpLrSchedule s;
    int st = get_lr_schedule(c, F, gt / 32, &s);
    if (st) return st;
    p.rec_v = s.rec_v; p.steps = s.steps; p.rec_words = s.rec_words; p.steps_c = s.steps_c; p.steps_v = s.steps_v;
    p.var_store = s.var_store; p.E = s.n_slots;             // message slots including the idle ones
    p.gt = gt;
    const size_t tables = up16((size_t) 4 * (s.rec_words + (s.steps_c + s.steps_v) * gt));
    const size_t team_bytes = lr_team_bytes(c, s.n_slots, F, soft, exp_mode, &p);
    p.off_teams = (uint32_t) tables;
    if (tables + team_bytes > LR_SMEM_MAX) return fail(LDPC_E_UNSUPPORTED, "BP messages of this code exceed 227 KB of shared memory");
    int teams = (int) std::min<size_t>(std::min<size_t>(15, max_threads / gt), (LR_SMEM_MAX - tables) / team_bytes);
    if (const char *force = getenv("LDPC_BP_TEAMS")) teams = std::max(1, std::min(teams, atoi(force)));
    int sms = 0;
    LDPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    const long long want = (frames + F - 1) / F;            // teams the batch can keep busy
    const long long grid = std::min<long long>(sms, want);
    teams = (int) std::min<long long>(teams, (want + grid - 1) / grid);
    const int threads = teams * gt;
    const size_t smem = tables + (size_t) teams * team_bytes;
    LrKernel kernel = lr_kernel(F, threads, soft, wide);
    LDPC_CUDA(allow_max_dynamic_smem(kernel));
    int per_sm = 0;
    LDPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) return fail(LDPC_E_UNSUPPORTED, "BP state of this code does not fit on one SM");
    // frames are claimed from the global queue in chunks; small enough that the tail stays balanced
    p.chunk = (int) std::max<long long>(1, std::min<long long>(F, frames / (grid * teams * 4 * F) * F));
    kernel<<<(unsigned) grid, threads, smem, stream>>>(p);
    LDPC_CUDA(cudaGetLastError());
    return LDPC_OK;
}

}  // namespace ldpc
