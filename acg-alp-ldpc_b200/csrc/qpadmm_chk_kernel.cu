// QP-ADMM decoding, check-centric kernel (sm_100a) -- DecodeQPADMM, algo/qp_admm.h:104-178, for codes
// whose checks have degree <= 12 and whose variables have at most 15 edges (every code of BASELINE.json and the
// proposals of optimize_H.cpp, including their degree-0/1/2 checks).  qpadmm_kernel.cu serves the rest.
//
// ConstructADMMProblem (qp_admm.h:59-92) splits a check of degree d into a chain of d - 2 three-variable
// blocks that are linked by d - 3 auxiliary variables, and every auxiliary variable belongs to exactly two
// consecutive blocks of ONE check.  So one lane owns one (check, frame) pair for the whole life of the
// frame and keeps in registers
//     yl[k][4]   the duals of its blocks (the reference's z is a function of the same number, admm_rows.cuh)
//     aux[k]     the values of its auxiliary variables
// and only what crosses between checks and ORIGINAL variables goes through shared memory:
//     w01 / w23  the four row terms w = yl + mu (z - b) of every block (two 16-byte chunks per frame), written by
//                the check lanes, gathered by the variable lanes in ascending row order (qp_admm.h:133-138)
//     v          the variable values (two buffers, see below), written by the variable lanes, gathered by the
//                check lanes in ascending variable order (qp_admm.h:144-151)
//     qa, inv    q_i + alpha/2 and -1/(mu e_i - alpha)
// Compared with the block-per-lane kernel this drops the auxiliary variables' round trip (half of all gathers),
// all per-iteration table loads of the check phase (the lane <-> check mapping is static) and the sign-flip
// arithmetic of the residuals and of the auxiliary updates (signs are static there).
// Arithmetic is fp64, every operation an _rn intrinsic in the reference's order: v, the hard decisions and the
// iteration count are bit-identical to the reference (the summation order of the stop test is a tree).
//
// A trip of the main loop:
//   variable phase   v[trip & 1] of the slots that hold a frame; warp 0 meanwhile adds up the stop sums of the
//                    previous check phase (qp_admm.h:161-163)
//   barrier          frames whose stop test fired (or that ran out of iterations) are published from
//                    v[(trip - 1) & 1] -- the variable phase just executed for them is discarded -- and their
//                    slots refilled
//   check phase      residuals, duals, row terms, the auxiliary variables of the NEXT iteration, stop-sum partials
//   barrier
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "admm_rows.cuh"
#include "slots.cuh"
#include "smem_ptx.cuh"

namespace ldpc {

constexpr int CHK_MAX_NB = 10;    // blocks per check (degree <= 12)
constexpr uint32_t CHK_EMPTY_SLOT = 0xffffffffu;

struct AdmmChkParams {
    KernelIO io;
    const uint32_t *chk_tab;      // per check rank: tab_stride words: degree, then the STORAGE SLOTS of its variables (ascending variable index)
    const uint32_t *var_words;    // per variable slot: (offset of its incidence records) << 10 | e << 4 | incidences, or CHK_EMPTY_SLOT
    const uint4 *var_inc;         // incidence records {chunk index, flip mask w0, flip mask w1, flip mask w2}, row order
    const uint16_t *var_slot;     // variable index -> storage slot
    const uint16_t *slot_e;       // per slot: sum of squared coefficients of the variable's column (qp_admm.h:94-99)
    uint32_t plane_base[CHK_MAX_NB];   // chunk index of check rank 0's k-th block
    int n_chk, n_slots, n_chunks, n_inc, tab_stride;
    int special_lo, special_hi;   // chunks [lo, hi) belong to one- and two-variable checks (no row with b = 2)
    // byte offsets of the arrays in dynamic shared memory
    uint32_t off_w23, off_v, off_qa, off_inv, off_red, off_inc, off_vw, off_cw, off_ctl;
    int max_iter;
    double alpha, mu, eps_stop;
    int chunk;                    // frames claimed from the global queue at a time
    // grid mode (qpadmm_params.cpp:51-67): work item q of the queue = frame q % grid_frames under the parameters of
    // point q / grid_frames; counters per point
    const double *grid_alpha, *grid_mu;
    long long grid_frames;
};

struct ChkCtl {
    unsigned live;                // slots holding a frame
    unsigned ran;                 // slots that took part in the last check phase
    unsigned done[2];             // by trip parity: slots whose frame is finished (set by warps 0..F-1 during the variable phase)
    unsigned fresh;
    long long q_next, q_end;
};

template <int F>
struct ChkShared {
    SlotBlock<F> S;
    ChkCtl c;
    double alpha[F], mu[F];       // grid mode: parameters of the slot's work item
    double inv_tab[64][F];        // grid mode: inv_coef by e (sum of squared coefficients of the column, <= 60)
    int point[F];
};

// One check with NBK blocks (degree NBK + 2), one frame.  va = the values of its variables in ascending index
// order.  Writes the row terms of its blocks, updates yl / aux, returns the partial stop sum.
template <int NBK, int NB>
__device__ __forceinline__ double chk_update(const double (&va)[NB + 2], double (&yl)[NB][4], double (&aux)[NB],
                                             uint32_t a_w01, uint32_t off_w23, const uint32_t (&plane_off)[NB],
                                             double mu, double half_mu, double half_alpha, double inv_aux) {
    constexpr int D = NBK + 2;
    double part = 0.0, P = 0.0;
#pragma unroll
    for (int k = 0; k < NBK; ++k) {
        double r0, r1, r2, r3;
        // the block's variables in ascending index order (originals before auxiliaries) and their slots
        if (NBK == 1) residual_rows<0, 1, 2>(va[0], va[1], va[2], 2.0, r0, r1, r2, r3);
        else if (k == 0) residual_rows<0, 1, 2>(va[0], va[1], aux[0], 2.0, r0, r1, r2, r3);
        else if (k == NBK - 1) residual_rows<1, 2, 0>(va[D - 2], va[D - 1], aux[k - 1], 2.0, r0, r1, r2, r3);
        else residual_rows<1, 0, 2>(va[k + 1], aux[k - 1], aux[k], 2.0, r0, r1, r2, r3);
        const double w0 = row_update_fp<false>(r0, yl[k][0], part, mu, half_mu);
        const double w1 = row_update_fp<false>(r1, yl[k][1], part, mu, half_mu);
        const double w2 = row_update_fp<false>(r2, yl[k][2], part, mu, half_mu);
        const double w3 = row_update_fp<true>(r3, yl[k][3], part, mu, half_mu);
        sts_f64x2(a_w01 + plane_off[k], w0, w1);
        sts_f64x2(a_w01 + off_w23 + plane_off[k], w2, w3);
        // the auxiliary variable between blocks k-1 and k, for the next iteration (qp_admm.h:132-142 with q = 0):
        // rows of block k-1 (slot 2: -,-,+,+) then rows of block k (slot 0: +,-,-,+)
        if (k > 0) {
            double B = __dadd_rn(P, w0);
            B = __dadd_rn(B, -w1);
            B = __dadd_rn(B, -w2);
            B = __dadd_rn(B, w3);
            aux[k - 1] = clip01_int(__dmul_rn(B, inv_aux));
        }
        if (k < NBK - 1) {
            P = __dadd_rn(half_alpha, -w0);
            P = __dadd_rn(P, -w1);
            P = __dadd_rn(P, w2);
            P = __dadd_rn(P, w3);
        }
    }
    return part;
}

// inv_coef, qp_admm.h:123-127 (A = (mu e - alpha)/2; inv = -1/(2A))
__device__ __forceinline__ double inv_coef(double mu, double alpha, double e) {
    const double A = __dmul_rn(__dadd_rn(__dmul_rn(mu, e), -alpha), 0.5);
    return __ddiv_rn(-1.0, __dmul_rn(2.0, A));
}

// value of an auxiliary variable in iteration 0: z = yl = 0, so w = (0, 0, 0, mu (0 - 2)) in both of its blocks
__device__ __forceinline__ double aux_start(double mu, double half_alpha, double inv_aux) {
    const double w3 = __fma_rn(mu, __dadd_rn(0.0, -2.0), 0.0);
    double B = __dadd_rn(half_alpha, -0.0);
    B = __dadd_rn(B, -0.0);
    B = __dadd_rn(B, 0.0);
    B = __dadd_rn(B, w3);
    B = __dadd_rn(B, 0.0);
    B = __dadd_rn(B, -0.0);
    B = __dadd_rn(B, -0.0);
    B = __dadd_rn(B, w3);
    return clip01_int(__dmul_rn(B, inv_aux));
}

// A check of degree 1 or 2 (qp_admm.h:70-83): ROWS = degree inequality rows with b = 0 and no auxiliary variable;
// the missing rows of its chunk stay zero, so the variable phase needs no special case.
template <int ROWS>
__device__ __forceinline__ double chk_special(double v0, double v1, double &yl0, double &yl1, uint32_t a_w01,
                                              uint32_t off_w23, uint32_t plane_off0, double mu, double half_mu) {
    double part = 0.0;
    // r = b - A v in ascending variable order: row 0 = (+1, -1), row 1 = (-1, +1)
    const double r0 = ROWS == 2 ? __dadd_rn(-v0, v1) : -v0;
    const double w0 = row_update_fp<false>(r0, yl0, part, mu, half_mu);
    double w1 = 0.0;
    if (ROWS == 2) w1 = row_update_fp<false>(__dadd_rn(v0, -v1), yl1, part, mu, half_mu);
    sts_f64x2(a_w01 + plane_off0, w0, w1);
    sts_f64x2(a_w01 + off_w23 + plane_off0, 0.0, 0.0);
    return part;
}

template <int F, int NB, bool GRID>
__global__ void __launch_bounds__(NB <= 6 ? 640 : 320, 1) qpadmm_chk_kernel(const AdmmChkParams p) {
    extern __shared__ __align__(16) double smem[];
    const KernelIO &io = p.io;
    const int n = io.n;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int warp = tid >> 5, nwarps = nt >> 5, lane = tid & 31;
    const int f = tid % F, cr = tid / F;              // this lane's frame slot and check rank / variable column (static)
    const int cpt = nt / F;                           // variable slots between the steps of a lane
    constexpr int LPF = 32 / F;                       // lanes of one frame in a warp

    char *sm = reinterpret_cast<char *>(smem);
    const uint32_t sbase = smem_addr(smem);
    uint8_t *cw = reinterpret_cast<uint8_t *>(sm + p.off_cw);          // F x n (experiment mode)
    ChkShared<F> *L = reinterpret_cast<ChkShared<F> *>(sm + p.off_ctl);
    SlotBlock<F> *S = &L->S;
    double *v_gen = reinterpret_cast<double *>(sm + p.off_v);          // generic views for the cold paths
    double *qa_gen = reinterpret_cast<double *>(sm + p.off_qa);
    double *red_gen = reinterpret_cast<double *>(sm + p.off_red);

    slots_init(S);
    for (int r = tid; r < p.n_slots; r += nt) {
        reinterpret_cast<double *>(sm + p.off_inv)[r] = inv_coef(p.mu, p.alpha, (double) p.slot_e[r]);
        reinterpret_cast<uint32_t *>(sm + p.off_vw)[r] = p.var_words[r];
    }
    for (int a = tid; a < p.n_inc; a += nt) {
        uint4 rec = p.var_inc[a];
        rec.x = rec.x * (F * 16);                     // chunk index -> byte offset
        reinterpret_cast<uint4 *>(sm + p.off_inc)[a] = rec;
    }
    if (tid == 0) {
        L->c.live = L->c.ran = L->c.done[0] = L->c.done[1] = L->c.fresh = 0u;
        L->c.q_next = L->c.q_end = 0;
        S->alive = F;
    }
    // penalty parameters of this lane's frame slot (grid mode: reloaded whenever a new work item enters the slot)
    double mu = p.mu, half_mu = __dmul_rn(p.mu, 0.5), half_alpha = __dmul_rn(p.alpha, 0.5);
    double inv_aux = inv_coef(p.mu, p.alpha, 8.0);     // auxiliary variables: e = 8 (two blocks x four rows)
    double aux_init = aux_start(mu, half_alpha, inv_aux);

    // ---- this lane's check (static): degree, variable offsets, chunk offsets
    const bool has_chk = cr < p.n_chk;
    int nb = -2;                                   // degree - 2: -1 / 0 = the one- and two-variable checks
    uint32_t voff[NB + 2], plane_off[NB];
#pragma unroll
    for (int j = 0; j < NB + 2; ++j) voff[j] = 0;
    if (has_chk) {
        const uint32_t *tab = p.chk_tab + (size_t) cr * p.tab_stride;
        nb = (int) tab[0] - 2;
#pragma unroll
        for (int j = 0; j < NB + 2; ++j)
            if (j < nb + 2) voff[j] = tab[1 + j] * (F * 8);
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) plane_off[k] = (p.plane_base[k] + cr) * (F * 16);
    double yl[NB][4], aux[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        yl[k][0] = yl[k][1] = yl[k][2] = yl[k][3] = 0.0;
        aux[k] = aux_init;
    }
    // shared-window addresses of this lane's frame column
    const uint32_t a_w01 = sbase + f * 16;                          // + chunk * F * 16 (w23: + off_w23)
    const uint32_t a_v0 = sbase + p.off_v + f * 8;                  // + slot * F * 8   (second buffer: + vbuf)
    const uint32_t vbuf = (uint32_t) p.n_slots * F * 8;
    const uint32_t a_qa = sbase + p.off_qa + f * 8;
    const uint32_t a_inv = sbase + p.off_inv, a_inc = sbase + p.off_inc, a_vw = sbase + p.off_vw;
    const uint32_t a_invtab = smem_addr(&L->inv_tab[0][0]) + f * 8;
    __syncthreads();

    for (unsigned trip = 0;; ++trip) {
        const uint32_t a_vcur = a_v0 + ((trip & 1) ? vbuf : 0u);
        double *vcur_gen = v_gen + (size_t) (trip & 1) * p.n_slots * F;
        const double *vprev_gen = v_gen + (size_t) ((trip & 1) ^ 1) * p.n_slots * F;
        const unsigned live = L->c.live, ran = L->c.ran;

        // ---- warp q < F: stop test of slot q after the previous check phase (qp_admm.h:161-163) / out of iterations;
        // lane j adds the partial of warp j, then a shuffle tree
        if (warp < F) {
            const int q = warp;
            double sum2 = (((ran >> q) & 1u) && lane < nwarps) ? red_gen[q * 32 + lane] : 0.0;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) sum2 += __shfl_xor_sync(0xffffffffu, sum2, off);
            if (lane == 0 && ((live >> q) & 1u)) {
                const bool fin = (((ran >> q) & 1u) && sum2 < p.eps_stop) || S->iter[q] >= p.max_iter;
                if (fin) atomicOr(&L->c.done[trip & 1], 1u << q);
            }
        }
        if (tid == 0) L->c.fresh = 0u;    // everybody read it before the last barrier; the refill below sets it again

        // ---- variable phase, qp_admm.h:132-142: this lane's column of variable slots
        // (prefetching the next incidence's records and row terms costs more in registers than it hides: 153 vs 141 ms)
        if ((live >> f) & 1u) {
            for (int slot = cr; slot < p.n_slots; slot += cpt) {
                const uint32_t word = lds_u32(a_vw + slot * 4);
                if (word == CHK_EMPTY_SLOT) continue;
                uint32_t rec = a_inc + (word >> 10) * 16;
                const uint32_t rec_end = rec + (word & 15u) * 16;
                double B = lds_f64(a_qa + slot * (F * 8));
                for (; rec != rec_end; rec += 16) {
                    const uint4 r = lds_u32x4(rec);
                    const double2 a01 = lds_f64x2(a_w01 + r.x), a23 = lds_f64x2(a_w01 + p.off_w23 + r.x);
                    B = __dadd_rn(B, __hiloint2double(__double2hiint(a01.x) ^ (int) r.y, __double2loint(a01.x)));
                    B = __dadd_rn(B, __hiloint2double(__double2hiint(a01.y) ^ (int) r.z, __double2loint(a01.y)));
                    B = __dadd_rn(B, __hiloint2double(__double2hiint(a23.x) ^ (int) r.w, __double2loint(a23.x)));
                    B = __dadd_rn(B, a23.y);
                }
                const double ic = GRID ? lds_f64(a_invtab + ((word >> 4) & 63u) * (F * 8)) : lds_f64(a_inv + slot * 8);
                sts_f64(a_vcur + slot * (F * 8), clip01_int(__dmul_rn(B, ic)));
            }
        }
        __syncthreads();

        // ---- publish finished frames (from the previous buffer), refill their slots
        const unsigned done = L->c.done[trip & 1];
        if (tid == 0) L->c.done[(trip & 1) ^ 1] = 0u;      // the buffer of the next trip; last read a trip ago
        if (done || trip == 0) {
            for (int q = 0; q < F; ++q) {
                if (!((done >> q) & 1u)) continue;
                const double *vf = vprev_gen + q;
                int valid = 1;
                if (io.experiment) {
                    int bad = 0;
                    for (int c = tid; c < io.m; c += nt) {
                        int parity = 0;
                        for (int e = io.row_ptr[c]; e < io.row_ptr[c + 1]; ++e)
                            parity ^= vf[(size_t) p.var_slot[io.col_idx[e]] * F] > 0.5 ? 1 : 0;
                        bad |= parity;
                    }
                    valid = !__syncthreads_or(bad);
                }
                slot_finish<F>(io, S, q, 1, 1, valid, S->iter[q], cw,
                               [&](int i) { return vf[(size_t) p.var_slot[i] * F] > 0.5 ? 1 : 0; },
                               [&](int i) { return vf[(size_t) p.var_slot[i] * F]; });
                if (GRID) {                          // the frame's counts go to its point's block
                    __syncthreads();
                    if (tid < LDPC_CNT_COUNT) {
                        const unsigned long long cnt = S->cnt[tid];
                        if (cnt) atomicAdd(&io.counters[(size_t) L->point[q] * LDPC_CNT_COUNT + tid], cnt);
                        S->cnt[tid] = 0ull;
                    }
                    __syncthreads();
                }
            }
            __syncthreads();
            if (warp == 0) {
                const unsigned live_before = live & ~done;
                const bool want = lane < F && !((live_before >> lane) & 1u) && S->state[lane] != SLOT_DEAD;
                const unsigned wmask = __ballot_sync(0xffffffffu, want);
                const int need = __popc(wmask), rank = __popc(wmask & ((1u << lane) - 1u));
                const long long next = L->c.q_next, end = L->c.q_end;
                __syncwarp();
                const long long left = end - next;
                long long got = 0, amt = 0;
                if (need > left) {
                    amt = max((long long) p.chunk, need - left);
                    if (lane == 0) got = (long long) atomicAdd(io.queue, (unsigned long long) amt);
                    got = __shfl_sync(0xffffffffu, got, 0);
                }
                if (want) {
                    const long long fr = rank < left ? next + rank : got + (rank - left);
                    if (fr < io.frames) {
                        S->iter[lane] = 0; S->hamming[lane] = 0; S->state[lane] = SLOT_NEW;
                        if (GRID) {
                            const long long pt = fr / p.grid_frames;
                            S->frame[lane] = fr - pt * p.grid_frames;
                            L->point[lane] = (int) pt;
                            L->alpha[lane] = p.grid_alpha[pt];
                            L->mu[lane] = p.grid_mu[pt];
                        } else {
                            S->frame[lane] = fr;
                        }
                    } else {
                        S->state[lane] = SLOT_DEAD;
                    }
                }
                if (lane == 0) {
                    if (need > left) { L->c.q_next = got + (need - left); L->c.q_end = got + amt; }
                    else L->c.q_next = next + need;
                }
                __syncwarp();
                const int st = lane < F ? S->state[lane] : SLOT_DEAD;
                const unsigned fresh = __ballot_sync(0xffffffffu, st == SLOT_NEW);
                const unsigned dead = __ballot_sync(0xffffffffu, lane < F && st == SLOT_DEAD);
                if (lane == 0) {
                    L->c.fresh = fresh;
                    L->c.live = live_before | fresh;
                    S->alive = F - __popc(dead);
                }
                if (lane < F && st == SLOT_NEW) S->state[lane] = SLOT_ACTIVE;
            }
            __syncthreads();
            if (S->alive == 0) break;
            const unsigned fresh = L->c.fresh;
            if (fresh) {
                if (GRID && tid < 64 * F) {          // inv_coef of the slot's parameters by e
                    const int q = tid % F, e = tid / F;
                    if ((fresh >> q) & 1u) L->inv_tab[e][q] = inv_coef(L->mu[q], L->alpha[q], (double) e);
                }
                // z = yl = 0 (qp_admm.h:120-121): w = mu (0 - b) -- the first variable phase of the frame reads it
                for (int i = tid; i < p.n_chunks * F; i += nt)
                    if ((fresh >> (i % F)) & 1u) {
                        const int chunk = i / F;
                        const double b3 = (chunk >= p.special_lo && chunk < p.special_hi) ? 0.0 : 2.0;
                        const double w3 = __fma_rn(GRID ? L->mu[i % F] : p.mu, __dadd_rn(0.0, -b3), 0.0);
                        sts_f64x2(sbase + i * 16, 0.0, 0.0);
                        sts_f64x2(sbase + p.off_w23 + i * 16, 0.0, w3);
                    }
                // q + alpha/2, and the v before the first update (qp_admm.h:116-119, visible only if max_iter == 0):
                // it goes to the buffer a frame that finishes at once is published from
                slots_load<F>(io, S, fresh, nullptr, 0, cw, [&](int i, int q, double l) {
                    const int r = p.var_slot[i];
                    qa_gen[r * F + q] = __dadd_rn(l, __dmul_rn(GRID ? L->alpha[q] : p.alpha, 0.5));
                    vcur_gen[r * F + q] = l > 0.0 ? 1.0 : 0.0;
                });
            }
        }

        // ---- check phase, qp_admm.h:144-159, for the slots that are iterating
        const unsigned fresh = L->c.fresh;
        const unsigned run = live & ~done;          // had a variable phase this trip and are not finished
        double part = 0.0;
        if ((fresh >> f) & 1u) {                    // a new frame moved into this lane's slot: z = yl = 0
            if (GRID) {
                mu = L->mu[f];
                half_mu = __dmul_rn(mu, 0.5);
                half_alpha = __dmul_rn(L->alpha[f], 0.5);
                inv_aux = inv_coef(mu, L->alpha[f], 8.0);
                aux_init = aux_start(mu, half_alpha, inv_aux);
            }
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                yl[k][0] = yl[k][1] = yl[k][2] = yl[k][3] = 0.0;
                aux[k] = aux_init;
            }
        }
        if (has_chk && ((run >> f) & 1u)) {
            double va[NB + 2];
#pragma unroll
            for (int j = 0; j < NB + 2; ++j) va[j] = (j < nb + 2) ? lds_f64(a_vcur + voff[j]) : 0.0;
#define LDPC_CHK_CASE(K)                                                                                          \
    case K:                                                                                                       \
        if (NB >= K)                                                                                              \
            part = chk_update<(NB >= K ? K : 1), NB>(va, yl, aux, a_w01, p.off_w23, plane_off, mu, half_mu,       \
                                                     half_alpha, inv_aux);                                        \
        break;
            switch (nb) {
                case -1: part = chk_special<1>(va[0], va[1], yl[0][0], yl[0][1], a_w01, p.off_w23, plane_off[0], mu, half_mu); break;
                case 0: part = chk_special<2>(va[0], va[1], yl[0][0], yl[0][1], a_w01, p.off_w23, plane_off[0], mu, half_mu); break;
                LDPC_CHK_CASE(1) LDPC_CHK_CASE(2) LDPC_CHK_CASE(3) LDPC_CHK_CASE(4) LDPC_CHK_CASE(5) LDPC_CHK_CASE(6)
                LDPC_CHK_CASE(7) LDPC_CHK_CASE(8) LDPC_CHK_CASE(9) LDPC_CHK_CASE(10)
                default: break;
            }
#undef LDPC_CHK_CASE
        }
        // the lanes of one frame are the lanes with equal lane % F
#pragma unroll
        for (int off = 16; off >= F; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        if (lane < F) red_gen[lane * 32 + warp] = part;
        if (warp == 0) {
            if (lane < F && ((run >> lane) & 1u)) S->iter[lane] += 1;
            if (lane == 0) L->c.ran = run;
        }
        __syncthreads();
    }
    slots_flush(io, S);
}

// ---------------------------------------------------------------- host side

template <typename T>
static int upload_chk(T **dst, const std::vector<T> &src) {
    LDPC_CUDA(cudaMalloc((void **) dst, sizeof(T) * std::max<size_t>(src.size(), 1)));
    if (!src.empty()) LDPC_CUDA(cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
    return LDPC_OK;
}

static int live_checks(const ldpc_code *c) {
    int k = 0;
    for (int r = 0; r < c->m; ++r) k += c->row_ptr[r + 1] > c->row_ptr[r];
    return k;
}

static int chk_threads(const ldpc_code *c, int F) { return (std::max(1, live_checks(c)) * F + 31) / 32 * 32; }

// Tables of the check-centric kernel for F frames per CTA (the variable slots depend on the CTA's column count);
// `supported` = every check has degree <= 12 and every variable at most 15 edges.
static int get_chk_tables(const ldpc_code *c, int F, const AdmmChkTables **out) {
    std::lock_guard<std::mutex> lock(c->sched_mu);
    AdmmChkTables &t = c->admm_chk[F == 4 ? 2 : (F == 2 ? 1 : 0)];
    if (!t.built) {
        t.built = true;
        t.supported = c->m > 0 && c->n < 65535;
        int live_checks = 0;
        for (int r = 0; r < c->m && t.supported; ++r) {
            const int d = c->row_ptr[r + 1] - c->row_ptr[r];
            if (d > CHK_MAX_NB + 2) t.supported = false;
            live_checks += d > 0;
        }
        for (int v = 0; v < c->n && t.supported; ++v)
            if (c->col_ptr[v + 1] - c->col_ptr[v] > 15) t.supported = false;
        if (live_checks == 0) t.supported = false;
        if (t.supported) {
            const int n = c->n;
            // checks with at least one edge by degree (descending, stable; checks without edges have no rows,
            // qp_admm.h:67-69), variables by degree (descending, stable)
            auto cdeg = [&](int r) { return c->row_ptr[r + 1] - c->row_ptr[r]; };
            auto vdeg = [&](int v) { return c->col_ptr[v + 1] - c->col_ptr[v]; };
            std::vector<int> chk, var(n), rank_of_chk(c->m, -1);
            for (int i = 0; i < c->m; ++i)
                if (cdeg(i) > 0) chk.push_back(i);
            const int m = (int) chk.size();
            for (int i = 0; i < n; ++i) var[i] = i;
            std::stable_sort(chk.begin(), chk.end(), [&](int a, int b) { return cdeg(a) > cdeg(b); });
            std::stable_sort(var.begin(), var.end(), [&](int a, int b) { return vdeg(a) > vdeg(b); });
            for (int i = 0; i < m; ++i) rank_of_chk[chk[i]] = i;
            t.n_chk = m;
            t.max_nb = std::max(1, cdeg(chk[0]) - 2);
            // Variable slots: lane column c of the CTA updates the slots c, c + cols, c + 2 cols, ...  The variables
            // are dealt to the columns in boustrophedon order, heaviest first, so that every column gets about the
            // same number of incidences (the variable phase ends at a barrier) and neighbouring columns -- the lanes
            // of one warp -- get variables of equal degree.
            const int cols = chk_threads(c, F) / F;
            const int steps = (n + cols - 1) / cols;
            t.n_slots = steps * cols;
            std::vector<int> slot_of_var(n, 0), var_of_slot(t.n_slots, -1);
            for (int i = 0; i < n; ++i) {
                const int j = i / cols, k = i % cols;
                const int col = (j & 1) ? cols - 1 - k : k;
                slot_of_var[var[i]] = j * cols + col;
                var_of_slot[j * cols + col] = var[i];
            }
            // block k of every check that has one: a plane of consecutive chunks indexed by check rank
            int base = 0;
            for (int k = 0; k < CHK_MAX_NB; ++k) {
                t.plane_base[k] = (uint32_t) base;
                int cnt = 0;
                for (int i = 0; i < m; ++i) cnt += std::max(1, cdeg(chk[i]) - 2) > k;    // one- and two-variable checks: one chunk
                base += (cnt + 1) & ~1;            // even bases: neighbouring checks write neighbouring rows
            }
            t.n_chunks = base;
            {   // the one- and two-variable checks are the last ranks (degree descending): their chunks in plane 0
                int m3 = 0;
                for (int i = 0; i < m; ++i) m3 += cdeg(chk[i]) >= 3;
                t.special_lo = (int) t.plane_base[0] + m3;
                t.special_hi = (int) t.plane_base[0] + m;
            }
            t.tab_stride = CHK_MAX_NB + 3;
            std::vector<uint32_t> tab((size_t) m * t.tab_stride, 0u);
            for (int i = 0; i < m; ++i) {
                const int r = chk[i];
                tab[(size_t) i * t.tab_stride] = (uint32_t) cdeg(r);
                for (int e = c->row_ptr[r], j = 0; e < c->row_ptr[r + 1]; ++e, ++j)
                    tab[(size_t) i * t.tab_stride + 1 + j] = (uint32_t) slot_of_var[c->col_idx[e]];
            }
            // variable incidences in ascending row order = ascending check index (a variable is in one block per check)
            std::vector<uint32_t> words(t.n_slots, CHK_EMPTY_SLOT);
            std::vector<uint4> inc;
            std::vector<uint16_t> vslot(n), se(t.n_slots, 4);
            int e_min = 1000000000;                       // over ALL variables, as qp_admm.h:108-111
            for (int r : chk)
                if (cdeg(r) >= 4) e_min = std::min(e_min, 8);     // auxiliary variables: two blocks x four rows
            for (int sl = 0; sl < t.n_slots; ++sl) {
                const int v = var_of_slot[sl];
                if (v < 0) continue;
                vslot[v] = (uint16_t) sl;
                const uint32_t first = (uint32_t) inc.size();
                int e = 0;
                for (int q = c->col_ptr[v]; q < c->col_ptr[v + 1]; ++q) {
                    const int edge = c->csc_edge[q];
                    const int r = (int) (std::upper_bound(c->row_ptr.begin(), c->row_ptr.end(), edge) - c->row_ptr.begin()) - 1;
                    const int d = cdeg(r), j = edge - c->row_ptr[r];
                    uint4 rec;
                    if (d >= 3) {               // block and slot of the variable in the chain (qp_admm.h:84-91)
                        const int k = j == 0 ? 0 : (j == d - 1 ? d - 3 : j - 1);
                        const int slot = j == 0 ? 0 : (j == d - 1 ? 2 : 1);
                        rec.x = t.plane_base[k] + (uint32_t) rank_of_chk[r];
                        rec.y = slot == 0 ? 0u : 0x80000000u;
                        rec.z = slot == 1 ? 0u : 0x80000000u;
                        rec.w = slot == 2 ? 0u : 0x80000000u;
                        e += 4;
                    } else {                    // qp_admm.h:70-83: rows (+1) or (+1, -1) / (-1, +1); the chunk's other rows are zero
                        rec.x = t.plane_base[0] + (uint32_t) rank_of_chk[r];
                        rec.y = j == 0 ? 0u : 0x80000000u;
                        rec.z = j == 1 ? 0u : 0x80000000u;
                        rec.w = 0u;
                        e += d;
                    }
                    inc.push_back(rec);
                }
                words[sl] = (first << 10) | ((uint32_t) e << 4) | (uint32_t) vdeg(v);
                se[sl] = (uint16_t) e;
                e_min = std::min(e_min, e);
            }
            if (inc.size() >= (1u << 22)) t.supported = false;
            t.n_inc = (int) inc.size();
            t.e_min = e_min;
            int st;
            if ((st = upload_chk(&t.chk_tab, tab))) return st;
            if ((st = upload_chk(&t.var_words, words))) return st;
            if ((st = upload_chk(&t.var_inc, inc))) return st;
            if ((st = upload_chk(&t.var_rank, vslot))) return st;
            if ((st = upload_chk(&t.var_e, se))) return st;
        }
    }
    *out = &t;
    return LDPC_OK;
}

static size_t up16(size_t x) { return (x + 15) & ~(size_t) 15; }

// carve-up of the dynamic shared memory; returns the total
static size_t chk_smem_layout(const ldpc_code *c, const AdmmChkTables &t, int F, bool experiment, AdmmChkParams *p) {
    size_t off = 0;
    off += (size_t) t.n_chunks * F * 16;                 // w01
    const size_t off_w23 = off; off += (size_t) t.n_chunks * F * 16;
    const size_t off_v = off; off += (size_t) 2 * t.n_slots * F * 8;
    const size_t off_qa = off; off += (size_t) t.n_slots * F * 8;
    const size_t off_inv = off; off += (size_t) t.n_slots * 8;
    const size_t off_red = off; off += (size_t) F * 32 * 8;
    const size_t off_inc = up16(off); off = off_inc + (size_t) t.n_inc * 16;
    const size_t off_vw = off; off += (size_t) t.n_slots * 4;
    const size_t off_cw = off; off += experiment ? (size_t) F * c->n : 0;
    const size_t off_ctl = up16(off); off = off_ctl + sizeof(ChkShared<4>);
    if (p) {
        p->off_w23 = (uint32_t) off_w23; p->off_v = (uint32_t) off_v; p->off_qa = (uint32_t) off_qa;
        p->off_inv = (uint32_t) off_inv; p->off_red = (uint32_t) off_red; p->off_inc = (uint32_t) off_inc;
        p->off_vw = (uint32_t) off_vw; p->off_cw = (uint32_t) off_cw; p->off_ctl = (uint32_t) off_ctl;
    }
    return off + 16;
}

using ChkKernel = void (*)(const AdmmChkParams);

template <int F, bool GRID>
static ChkKernel chk_kernel_for(int nb) {
    if (nb <= 2) return qpadmm_chk_kernel<F, 2, GRID>;
    if (nb <= 4) return qpadmm_chk_kernel<F, 4, GRID>;
    if (nb == 5) return qpadmm_chk_kernel<F, 5, GRID>;
    if (nb == 6) return qpadmm_chk_kernel<F, 6, GRID>;
    if (nb <= 8) return qpadmm_chk_kernel<F, 8, GRID>;
    return qpadmm_chk_kernel<F, 10, GRID>;
}

static ChkKernel chk_kernel_for(int F, int nb, bool grid) {
    if (grid) return F == 4 ? chk_kernel_for<4, true>(nb) : (F == 2 ? chk_kernel_for<2, true>(nb) : chk_kernel_for<1, true>(nb));
    return F == 4 ? chk_kernel_for<4, false>(nb) : (F == 2 ? chk_kernel_for<2, false>(nb) : chk_kernel_for<1, false>(nb));
}

// smallest 4 x column degree: DecodeQPADMM answers {zeros, false} when e_min * mu <= alpha (qp_admm.h:108-114)
int qpadmm_chk_e_min(const ldpc_code *c) {
    const AdmmChkTables *t = nullptr;
    if (get_chk_tables(c, 1, &t) || !t->supported) return -1;
    return t->e_min;
}

int launch_qpadmm_chk(const ldpc_code *c, const FrameIO &fio, int64_t frames, double var, double alpha, double mu,
                      int max_iter, double eps_stop, unsigned long long *queue, cudaStream_t stream,
                      const double *grid_alpha, const double *grid_mu, int64_t grid_points) {
    if (frames <= 0) return LDPC_OK;
    const bool grid = grid_points > 0;
    const int64_t frames_per_point = frames;
    if (grid) frames *= grid_points;             // work items of the queue
    // frames per CTA: one lane per (check, frame), at most 640 lanes.  Two CTAs of two frames per SM beat one CTA
    // of four (profiles/r01_admm_chk_sweep.txt): their barriers and their FP64-bound / latency-bound phases overlap.
    int F = 2;
    if (const char *force = getenv("LDPC_ADMM_F")) {
        const int v = atoi(force);
        if (v == 1 || v == 2 || v == 4) F = v;
    } else {
        while (F > 1 && frames < 2ll * 148 * F) F >>= 1;
    }
    const bool exp_mode = fio.experiment != 0;
    const AdmmChkTables *t = nullptr;
    for (;; F >>= 1) {
        int st = get_chk_tables(c, F, &t);
        if (st) return st;
        if (!t->supported) return LDPC_E_UNSUPPORTED;
        if (chk_threads(c, F) <= (t->max_nb <= 6 ? 640 : 320) && chk_smem_layout(c, *t, F, exp_mode, nullptr) <= 227 * 1024) break;
        if (F == 1) return LDPC_E_UNSUPPORTED;
    }
    if (!grid && (double) t->e_min * mu <= alpha) return LDPC_E_UNSUPPORTED;   // infeasible: the general kernel answers {zeros, false}
    AdmmChkParams p;
    KernelIO &io = p.io;
    io.y = fio.y; io.bits = fio.bits; io.ok = fio.ok; io.iters = fio.iters; io.soft = fio.soft;
    io.experiment = fio.experiment; io.cw_source = fio.cw_source; io.seed = fio.seed;
    io.frame_begin = fio.frame_begin; io.words = fio.words; io.n_words = fio.n_words;
    io.counters = fio.counters; io.gen_cols = c->d.gen_cols; io.k = c->k; io.k_words = c->k_words;
    io.frames = frames; io.queue = queue; io.var = var; io.sigma = std::sqrt(var);     // grid mode: frames = work items
    io.n = c->n; io.m = c->m; io.row_ptr = c->d.row_ptr; io.col_idx = c->d.col_idx;
    p.chk_tab = t->chk_tab; p.var_words = t->var_words; p.var_inc = t->var_inc; p.var_slot = t->var_rank;
    p.slot_e = t->var_e;
    for (int k = 0; k < CHK_MAX_NB; ++k) p.plane_base[k] = t->plane_base[k];
    p.special_lo = t->special_lo; p.special_hi = t->special_hi;
    p.n_chk = t->n_chk; p.n_slots = t->n_slots; p.n_chunks = t->n_chunks; p.n_inc = t->n_inc; p.tab_stride = t->tab_stride;
    p.max_iter = max_iter; p.alpha = alpha; p.mu = mu; p.eps_stop = eps_stop;
    p.grid_alpha = grid_alpha; p.grid_mu = grid_mu; p.grid_frames = frames_per_point;
    const int threads = chk_threads(c, F);
    const size_t smem = chk_smem_layout(c, *t, F, exp_mode, &p);
    ChkKernel fn = chk_kernel_for(F, t->max_nb, grid);
    LDPC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int per_sm = 0, sms = 0;
    LDPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
    LDPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    if (per_sm < 1) return LDPC_E_UNSUPPORTED;
    const long long want = (frames + F - 1) / F;
    const long long ctas = std::min<long long>((long long) per_sm * sms, want);
    p.chunk = (int) std::max<long long>(1, std::min<long long>(F, frames / (ctas * 4 * F) * F));
    fn<<<(unsigned) ctas, threads, smem, stream>>>(p);
    LDPC_CUDA(cudaGetLastError());
    return LDPC_OK;
}

void free_chk_tables(ldpc_code *c) {
    for (AdmmChkTables &t : c->admm_chk) {
        cudaFree(t.chk_tab); cudaFree(t.var_words); cudaFree(t.var_inc); cudaFree(t.var_rank); cudaFree(t.var_e);
    }
}

}  // namespace ldpc
