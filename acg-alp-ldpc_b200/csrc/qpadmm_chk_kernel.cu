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
// The variable phase of a lane is a static stream of one 32-bit word per incidence (chunk offset, sign flips, end
// of variable / end of column), software-pipelined one incidence ahead: its dependent chain is the reference's
// own chain of additions and nothing else (profiles/r01_admm_chk_sweep.txt: 140.7 -> 132.4 ms against per-slot
// records that were looked up on the way).
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
#include <cstdio>
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
    const uint16_t *var_slot;     // variable index -> storage slot
    const uint16_t *slot_e;       // per slot: sum of squared coefficients of the variable's column (qp_admm.h:94-99)
    const uint32_t *var_stream;   // variable phase as one word per incidence and lane column (see chk_stream_word)
    int stream_rows;              // words per lane column (row k of column c at index k * columns + c)
    uint32_t plane_base[CHK_MAX_NB];   // chunk index of check rank 0's k-th block
    int n_chk, n_slots, n_chunks, tab_stride;
    int special_lo, special_hi;   // chunks [lo, hi) belong to one- and two-variable checks (no row with b = 2)
    // byte offsets of the arrays in dynamic shared memory
    uint32_t off_w23, off_v, off_qa, off_inv, off_red, off_str, off_cw, off_ctl;
    int max_iter;
    double alpha, mu, eps_stop;
    // derived penalty constants, computed on the host with the device's operations (IEEE double, no contraction): as
    // kernel parameters they are constant-bank operands of the FP64 instructions instead of ten live registers
    double half_mu, half_alpha, inv_aux, aux_init;
    int chunk;                    // frames claimed from the global queue at a time
    // grid mode (qpadmm_params.cpp:51-67): work item q of the queue = frame q % grid_frames under the parameters of
    // point q / grid_frames; counters per point
    const double *grid_alpha, *grid_mu;
    long long grid_frames;
};

// Variable phase as a stream: the incidences of the variables of one lane column (slots c, c + columns, ...) back to
// back, one word each, in the reference's gather order (ascending row, qp_admm.h:133-138):
//   bits 4-19   byte offset of the block's chunk in w01 (w23: + off_w23)
//   bits 20-25  e of the variable (sum of squared coefficients of its column; grid mode looks inv_coef up by it)
//   bits 31/30/29  flip the sign of the row term 0 / 1 / 2 (coefficient -1)
//   bit 1 FIRST  first word of a variable (informative: the kernel loads q + alpha/2 of the next variable at LAST)
//   bit 0 LAST   last word of a variable: scale, clip, store v, move on to the column's next slot
//   bit 2 NOADD  no row terms (a variable without edges, or a column without variables): the offset is that of the
//                all-zero chunk behind the last real one
//   bit 3 END    last word of the column
constexpr uint32_t STR_LAST = 1u, STR_FIRST = 2u, STR_NOADD = 4u, STR_END = 8u, STR_OFF_MASK = 0xffff0u;

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
    double inv_tab[64][F];        // inv_coef by e (sum of squared coefficients of the column, <= 60) under the slot's parameters
    int point[F];
};

// One check with NBK blocks (degree NBK + 2), one frame.  va = the values of its variables in ascending index
// order.  Writes the row terms of its blocks, updates yl / aux, returns the partial stop sum.
// MIXED: the warp holds checks of NBK and of NBK - 1 blocks (the boundary between two degree classes).  Instead of
// running both code paths one after the other, the shorter checks (shrt) ride along: their last block takes the last
// step (its variables, chunk and duals were placed there when the lane was set up), the step before it -- the longer
// checks' last middle block -- is computed and discarded, and the auxiliary variable that joins their last two blocks
// is moved into the place the last step reads and back.  Every lane performs exactly the operations of its own
// check in the same order.
template <int NBK, int NB, bool MIXED>
__device__ __forceinline__ double chk_update(const double (&va)[NB + 2], double (&yl)[NB][4], double (&aux)[NB],
                                             uint32_t a_w01, uint32_t off_w23, const uint32_t (&plane_off)[NB],
                                             double mu, double half_mu, double half_alpha, double inv_aux, bool shrt) {
    constexpr int D = NBK + 2;
    double part = 0.0, P = 0.0;
#pragma unroll
    for (int k = 0; k < NBK; ++k) {
        const bool skip = MIXED && shrt && k == NBK - 2;
        if (MIXED && NBK >= 3 && k == NBK - 1) aux[k - 1] = shrt ? aux[k - 2] : aux[k - 1];
        double r0, r1, r2, r3;
        // the block's variables in ascending index order (originals before auxiliaries) and their slots
        if (NBK == 1) residual_rows<0, 1, 2>(va[0], va[1], va[2], 2.0, r0, r1, r2, r3);
        else if (k == 0) residual_rows<0, 1, 2>(va[0], va[1], aux[0], 2.0, r0, r1, r2, r3);
        else if (k == NBK - 1) residual_rows<1, 2, 0>(va[D - 2], va[D - 1], aux[k - 1], 2.0, r0, r1, r2, r3);
        else residual_rows<1, 0, 2>(va[k + 1], aux[k - 1], aux[k], 2.0, r0, r1, r2, r3);
        const double part_in = part;
        const double w0 = row_update_fp<false>(r0, yl[k][0], part, mu, half_mu);
        const double w1 = row_update_fp<false>(r1, yl[k][1], part, mu, half_mu);
        const double w2 = row_update_fp<false>(r2, yl[k][2], part, mu, half_mu);
        const double w3 = row_update_fp<true>(r3, yl[k][3], part, mu, half_mu);
        if (MIXED) part = skip ? part_in : part;
        if (!skip) {
            sts_f64x2(a_w01 + plane_off[k], w0, w1);
            sts_f64x2(a_w01 + off_w23 + plane_off[k], w2, w3);
        }
        // the auxiliary variable between blocks k-1 and k, for the next iteration (qp_admm.h:132-142 with q = 0):
        // rows of block k-1 (slot 2: -,-,+,+) then rows of block k (slot 0: +,-,-,+)
        if (k > 0) {
            double B = __dadd_rn(P, w0);
            B = __dadd_rn(B, -w1);
            B = __dadd_rn(B, -w2);
            B = __dadd_rn(B, w3);
            const double anew = clip01_int(__dmul_rn(B, inv_aux));
            aux[k - 1] = skip ? aux[k - 1] : anew;
            if (MIXED && NBK >= 3 && k == NBK - 1) aux[k - 2] = shrt ? aux[k - 1] : aux[k - 2];
        }
        if (k < NBK - 1) {
            double Pn = __dadd_rn(half_alpha, -w0);
            Pn = __dadd_rn(Pn, -w1);
            Pn = __dadd_rn(Pn, w2);
            Pn = __dadd_rn(Pn, w3);
            P = skip ? P : Pn;
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

// TWO: 64 registers per thread, so that two CTAs of up to 512 lanes share an SM (large codes: one frame per CTA fills
// 342..512 lanes, and at 96 registers a single CTA per SM has nobody to overlap its phases with)
// SHAPE 2: 80 registers per thread, so that FIVE CTAs of up to 160 lanes (the 160 x 280 codes, one frame each) share an SM
// instead of four at 96 registers.
template <int F, int NB, bool GRID, int SHAPE>
__global__ void __launch_bounds__(SHAPE == 1 ? 512 : (SHAPE == 2 ? 192 : (NB <= 6 ? 640 : 320)), SHAPE == 1 ? 2 : (SHAPE == 2 ? 4 : 1))
qpadmm_chk_kernel(const AdmmChkParams p) {
    constexpr bool TWO = SHAPE == 1;
    extern __shared__ __align__(16) double smem[];
    const KernelIO &io = p.io;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int warp = tid >> 5, nwarps = nt >> 5, lane = tid & 31;
    const int f = tid % F, cr = tid / F;              // this lane's frame slot and check rank / variable column (static)
    const int cpt = nt / F;                           // variable slots between the steps of a lane

    char *sm = reinterpret_cast<char *>(smem);
    const uint32_t sbase = smem_addr(smem);
    uint8_t *cw = reinterpret_cast<uint8_t *>(sm + p.off_cw);          // F x n (experiment mode)
    ChkShared<F> *L = reinterpret_cast<ChkShared<F> *>(sm + p.off_ctl);
    SlotBlock<F> *S = &L->S;
    double *v_gen = reinterpret_cast<double *>(sm + p.off_v);          // generic views for the cold paths
    double *qa_gen = reinterpret_cast<double *>(sm + p.off_qa);
    double *red_gen = reinterpret_cast<double *>(sm + p.off_red);

    slots_init(S);
    // inv_coef by e (sum of squared coefficients of the variable's column, <= 60): one table per frame slot (grid mode
    // refills a slot's column when a work item with other parameters enters it)
    for (int i = tid; i < 64 * F; i += nt) L->inv_tab[i / F][i % F] = inv_coef(p.mu, p.alpha, (double) (i / F));
    // ... and, where shared memory is not the limit, per slot: one load from a running address instead of a look-up
    if (!TWO && !GRID)
        for (int r = tid; r < p.n_slots; r += nt)
            reinterpret_cast<double *>(sm + p.off_inv)[r] = inv_coef(p.mu, p.alpha, (double) p.slot_e[r]);
    for (int a = tid; a < p.stream_rows * (nt / F); a += nt)
        reinterpret_cast<uint32_t *>(sm + p.off_str)[a] = p.var_stream[a];
    if (tid < F) {                                        // the all-zero chunk, the all-zero v slots
        sts_f64x2(sbase + (p.n_chunks * F + tid) * 16, 0.0, 0.0);
        sts_f64x2(sbase + p.off_w23 + (p.n_chunks * F + tid) * 16, 0.0, 0.0);
        v_gen[(size_t) p.n_slots * F + tid] = 0.0;
        v_gen[(size_t) (2 * p.n_slots + 1) * F + tid] = 0.0;
    }
    if (tid == 0) {
        L->c.live = L->c.ran = L->c.done[0] = L->c.done[1] = L->c.fresh = 0u;
        L->c.q_next = L->c.q_end = 0;
        S->alive = F;
    }
    // penalty parameters of this lane's frame slot (grid mode: reloaded whenever a new work item enters the slot)
    double mu = p.mu, half_mu = p.half_mu, half_alpha = p.half_alpha;
    double inv_aux = p.inv_aux;                        // auxiliary variables: e = 8 (two blocks x four rows)
    double aux_init = p.aux_init;

    // ---- this lane's check (static): degree, variable offsets, chunk offsets
    const bool has_chk = cr < p.n_chk;
    int nb = -2;                                   // degree - 2: -1 / 0 = the one- and two-variable checks
    uint32_t voff[NB + 2], plane_off[NB];
#pragma unroll
    for (int j = 0; j < NB + 2; ++j) voff[j] = (uint32_t) p.n_slots * (F * 8);      // the all-zero slot
    if (has_chk) nb = (int) p.chk_tab[(size_t) cr * p.tab_stride] - 2;
    // a warp at the boundary of two degree classes whose block counts differ by one runs ONE code path (chk_update,
    // MIXED): the shorter checks keep their last two variables and the chunk of their last block where the longer
    // checks' last step looks for them
    int mx = 0;                                    // 0: uniform warp, 1 / 2: a longer / shorter check of a mixed warp
    {
        const int nb_hi = __reduce_max_sync(0xffffffffu, has_chk ? nb : -100), nb_lo = __reduce_min_sync(0xffffffffu, has_chk ? nb : 100);
        if (nb_hi == nb_lo + 1 && nb_lo >= 2) mx = nb == nb_lo ? 2 : 1;
    }
    const bool shrt = mx == 2;
#pragma unroll
    for (int k = 0; k < NB; ++k) plane_off[k] = (p.plane_base[k] + cr) * (F * 16);
    if (has_chk) {
        const uint32_t *tab = p.chk_tab + (size_t) cr * p.tab_stride;
        // entry j of va: variable j of the check; for the shorter checks of a mixed warp the last two variables move
        // up by one (entries nb + 1, nb + 2) and entry nb is a hole
#pragma unroll
        for (int j = 0; j < NB + 2; ++j) {
            const int src = !shrt || j < nb ? j : (j == nb ? -1 : j - 1);
            if (src >= 0 && src < nb + 2) voff[j] = tab[1 + src] * (F * 8);
        }
        if (shrt) {                                 // the chunk of their last block, where the last step stores
#pragma unroll
            for (int k = 2; k < NB; ++k)
                if (k == nb) plane_off[k] = (p.plane_base[k - 1] + cr) * (F * 16);
        }
    }
    double yl[NB][4], aux[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        yl[k][0] = yl[k][1] = yl[k][2] = yl[k][3] = 0.0;
        aux[k] = aux_init;
    }
    // shared-window addresses of this lane's frame column
    const uint32_t a_w01 = sbase + f * 16;                          // + chunk * F * 16 (w23: + off_w23)
    const uint32_t a_v0 = sbase + p.off_v + f * 8;                  // + slot * F * 8   (second buffer: + vbuf)
    const uint32_t vbuf = (uint32_t) (p.n_slots + 1) * F * 8;       // (one all-zero slot behind each buffer: absent variables read it)
    const uint32_t a_qa = sbase + p.off_qa + f * 8;
    const uint32_t a_invtab = smem_addr(&L->inv_tab[0][0]) + f * 8;
    const uint32_t a_str = sbase + p.off_str + cr * 4;
    __syncthreads();

    for (unsigned trip = 0;; ++trip) {
        const uint32_t a_vcur = a_v0 + ((trip & 1) ? vbuf : 0u);
        double *vcur_gen = v_gen + (size_t) (trip & 1) * (p.n_slots + 1) * F;
        const double *vprev_gen = v_gen + (size_t) ((trip & 1) ^ 1) * (p.n_slots + 1) * F;
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
        if ((live >> f) & 1u) {
            // One word per incidence.  The row terms of the next incidence and the word after it are in flight while the
            // four additions of this one retire; two steps per trip of the loop so that the row-term registers alternate
            // instead of being copied.  B always holds q + alpha/2 of the variable about to start (loaded when its
            // predecessor is stored), so a variable needs no "first" test.
            // (Two interleaved streams per lane -- two addition chains -- change nothing, 330.6 vs 331.0 ms on the
            // (3,6)-1008 code: with the loads in flight early the phase is bound by shared-memory wavefronts.)
            const uint32_t a_w23 = a_w01 + p.off_w23;
            const uint32_t s_str = (uint32_t) cpt * 4, s_slot = (uint32_t) cpt * (F * 8), s_inv = (uint32_t) cpt * 8;
#define LDPC_VAR_STEP(W, A01, A23, B, AQ, AV, AI)                                                                  \
    B = __dadd_rn(B, flip_by(A01.x, W));                                                                           \
    B = __dadd_rn(B, flip_by(A01.y, W << 1));                                                                      \
    B = __dadd_rn(B, flip_by(A23.x, W << 2));                                                                      \
    B = __dadd_rn(B, A23.y);                                                                                       \
    if (W & STR_LAST) {                                                                                            \
        const double ic = (GRID || TWO) ? lds_f64(a_invtab + ((W >> 20) & 63u) * (F * 8)) : lds_f64(AI);           \
        AQ += s_slot;                                                                                              \
        const double vnew = clip01_int(__dmul_rn(B, ic));                                                          \
        B = lds_f64(AQ);                                                                                           \
        sts_f64(AV, vnew);                                                                                         \
        AV += s_slot;                                                                                              \
        AI += s_inv;                                                                                               \
    }
            uint32_t a_s = a_str, a_q = a_qa + cr * (F * 8), a_vv = a_vcur + cr * (F * 8), a_i = sbase + p.off_inv + cr * 8;
            uint32_t w = lds_u32(a_s);
            double2 a01 = lds_f64x2(a_w01 + (w & STR_OFF_MASK)), a23 = lds_f64x2(a_w23 + (w & STR_OFF_MASK));
            double B = lds_f64(a_q);
            a_s += s_str;
            uint32_t wn = lds_u32(a_s);
            for (;;) {
                const double2 n01 = lds_f64x2(a_w01 + (wn & STR_OFF_MASK)), n23 = lds_f64x2(a_w23 + (wn & STR_OFF_MASK));
                a_s += s_str;
                const uint32_t wnn = lds_u32(a_s);
                LDPC_VAR_STEP(w, a01, a23, B, a_q, a_vv, a_i)
                if (w & STR_END) break;
                a01 = lds_f64x2(a_w01 + (wnn & STR_OFF_MASK));
                a23 = lds_f64x2(a_w23 + (wnn & STR_OFF_MASK));
                a_s += s_str;
                w = lds_u32(a_s);
                LDPC_VAR_STEP(wn, n01, n23, B, a_q, a_vv, a_i)
                if (wn & STR_END) break;
                wn = w;
                w = wnn;
            }
#undef LDPC_VAR_STEP
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
                if (GRID)                            // inv_coef of the slot's parameters by e (a CTA may have as few as 32 threads)
                    for (int i = tid; i < 64 * F; i += nt) {
                        const int q = i % F, e = i / F;
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
            for (int j = 0; j < NB + 2; ++j) va[j] = lds_f64(a_vcur + voff[j]);       // absent entries: the all-zero slot
#define LDPC_CHK_CASE(K, MIX)                                                                                     \
    case K:                                                                                                       \
        if (NB >= K)                                                                                              \
            part = chk_update<(NB >= K ? K : 1), NB, MIX>(va, yl, aux, a_w01, p.off_w23, plane_off, mu, half_mu,  \
                                                          half_alpha, inv_aux, mx == 2);                          \
        break;
            if (mx) {
                switch (nb + (mx >> 1)) {
                    LDPC_CHK_CASE(3, true) LDPC_CHK_CASE(4, true) LDPC_CHK_CASE(5, true) LDPC_CHK_CASE(6, true)
                    LDPC_CHK_CASE(7, true) LDPC_CHK_CASE(8, true) LDPC_CHK_CASE(9, true) LDPC_CHK_CASE(10, true)
                    default: break;
                }
            } else {
                switch (nb) {
                    case -1: part = chk_special<1>(va[0], va[1], yl[0][0], yl[0][1], a_w01, p.off_w23, plane_off[0], mu, half_mu); break;
                    case 0: part = chk_special<2>(va[0], va[1], yl[0][0], yl[0][1], a_w01, p.off_w23, plane_off[0], mu, half_mu); break;
                    LDPC_CHK_CASE(1, false) LDPC_CHK_CASE(2, false) LDPC_CHK_CASE(3, false) LDPC_CHK_CASE(4, false)
                    LDPC_CHK_CASE(5, false) LDPC_CHK_CASE(6, false) LDPC_CHK_CASE(7, false) LDPC_CHK_CASE(8, false)
                    LDPC_CHK_CASE(9, false) LDPC_CHK_CASE(10, false)
                    default: break;
                }
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

// Bank-aware ranks for codes that run one frame per CTA (more than 320 live checks): with F = 1 a warp's gathers go to
// 32 unrelated addresses -- 43 % of all shared-memory wavefronts of the (3,6)-1008 code were replays
// (profiles/r01b_admm_chk_1008_ncu.txt).  The order of the checks inside a degree class and of the variables inside
// a degree class is free (warps stay uniform, the columns stay balanced, the arithmetic does not change), so a
// deterministic annealing pass picks both to spread every warp-wide gather over the banks:
//   v gather of the check phase   G = 32/F consecutive check ranks read, for position j, the slots of their j-th
//                                 variables: F x 8 bytes each, 16/F positions per 128-byte wavefront
//   w gather of the variable phase  G consecutive lane columns read, at row k of their streams, one chunk each:
//                                 F x 16 bytes, 8/F positions per wavefront
// Cost of an access = wavefronts beyond the minimum for its number of active lanes.
static void chk_anneal(const ldpc_code *c, int F, int cols, const uint32_t *plane_base, std::vector<int> &chk,
                       std::vector<int> &var) {
    const int n = c->n, m = (int) chk.size(), G = 32 / F, bins_a = 16 / F, bins_c = 8 / F;
    const int steps = (n + cols - 1) / cols;
    auto cdeg = [&](int r) { return c->row_ptr[r + 1] - c->row_ptr[r]; };
    auto vdeg = [&](int v) { return c->col_ptr[v + 1] - c->col_ptr[v]; };
    std::vector<int> edge_chk(c->E), edge_plane(c->E);
    int maxdeg = 0;
    for (int r = 0; r < c->m; ++r) {
        const int d = cdeg(r);
        maxdeg = std::max(maxdeg, d);
        for (int j = 0; j < d; ++j) {
            edge_chk[c->row_ptr[r] + j] = r;
            edge_plane[c->row_ptr[r] + j] = d >= 3 ? (j == 0 ? 0 : (j == d - 1 ? d - 3 : j - 1)) : 0;
        }
    }
    std::vector<int> rank_of_chk(c->m, -1), pos_of_var(n);
    for (int i = 0; i < m; ++i) rank_of_chk[chk[i]] = i;
    for (int i = 0; i < n; ++i) pos_of_var[var[i]] = i;
    auto col_of_pos = [&](int i) { const int j = i / cols, k = i % cols; return (j & 1) ? cols - 1 - k : k; };
    auto slot_of_pos = [&](int i) { return (i / cols) * cols + col_of_pos(i); };
    // An 8-byte access is served half-warp by half-warp (16 lanes x 8 B = one 128-byte wavefront), a 16-byte access
    // quarter-warp by quarter-warp (8 lanes x 16 B): a group is conflict-free iff its addresses fall into different
    // positions of the 128-byte line, and costs (largest multiplicity - 1) replays otherwise.
    const int ha = 16 / F, qc = 8 / F;               // checks per half-warp, columns per quarter-warp
    auto cost_a = [&](int g) {
        int cost = 0;
        const int r1 = std::min(m, g * G + G);
        for (int j = 0; j < maxdeg; ++j)
            for (int h0 = g * G; h0 < r1; h0 += ha) {
                int cnt[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, worst = 1;
                for (int i = h0; i < std::min(r1, h0 + ha); ++i) {
                    const int r = chk[i];
                    if (j >= cdeg(r)) continue;
                    worst = std::max(worst, ++cnt[slot_of_pos(pos_of_var[c->col_idx[c->row_ptr[r] + j]]) % bins_a]);
                }
                cost += worst - 1;
            }
        return cost;
    };
    std::vector<int> chunks;                         // scratch: [column in warp][row] -> chunk, -1 beyond the end
    auto cost_c = [&](int h) {
        const int c1 = std::min(cols, h * G + G), width = c1 - h * G;
        int rows = 0;
        chunks.assign((size_t) G * 16 * steps, -1);
        const int cap = 16 * steps;
        for (int col = h * G; col < c1; ++col) {
            int k = 0;
            for (int j = 0; j < steps; ++j) {
                const int i = j * cols + ((j & 1) ? cols - 1 - col : col);
                if (i >= n) continue;
                const int v = var[i];
                for (int q = c->col_ptr[v]; q < c->col_ptr[v + 1] && k < cap; ++q, ++k) {
                    const int e = c->csc_edge[q];
                    chunks[(size_t) (col - h * G) * cap + k] = (int) plane_base[edge_plane[e]] + rank_of_chk[edge_chk[e]];
                }
            }
            rows = std::max(rows, k);
        }
        int cost = 0;
        for (int k = 0; k < rows; ++k)
            for (int x0 = 0; x0 < width; x0 += qc) {
                int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, worst = 1;
                for (int x = x0; x < std::min(width, x0 + qc); ++x) {
                    const int q = chunks[(size_t) x * cap + k];
                    if (q >= 0) worst = std::max(worst, ++cnt[q % bins_c]);
                }
                cost += worst - 1;
            }
        return 2 * cost;                             // two 16-byte loads (w01, w23) per stream word
    };
    // class ranges (equal degree) by position
    std::vector<int> c_lo(m), c_hi(m), v_lo(n), v_hi(n);
    for (int i = 0, s0 = 0; i <= m; ++i)
        if (i == m || cdeg(chk[i]) != cdeg(chk[s0])) { for (int k = s0; k < i; ++k) { c_lo[k] = s0; c_hi[k] = i; } s0 = i; }
    for (int i = 0, s0 = 0; i <= n; ++i)
        if (i == n || vdeg(var[i]) != vdeg(var[s0])) { for (int k = s0; k < i; ++k) { v_lo[k] = s0; v_hi[k] = i; } s0 = i; }
    uint64_t rs = 0x9E3779B97F4A7C15ull;             // xorshift: the layout must not depend on the C++ library
    auto rnd = [&]() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t) (rs >> 11); };
    int moves_per_node = 40;
    if (const char *e = getenv("LDPC_ADMM_ANNEAL_MOVES")) moves_per_node = std::max(1, atoi(e));
    const int moves = moves_per_node * (n + m);
    std::vector<int> ga, gc;
    auto uniq = [](std::vector<int> &x) { std::sort(x.begin(), x.end()); x.erase(std::unique(x.begin(), x.end()), x.end()); };
    auto replays = [&]() {
        long t = 0;
        for (int g = 0; g * G < m; ++g) t += cost_a(g);
        for (int h = 0; h * G < cols; ++h) t += cost_c(h);
        return t;
    };
    const bool stats = getenv("LDPC_ADMM_LAYOUT_STATS") != nullptr;
    const long replays_before = replays();
    long now = replays_before, best = replays_before;
    std::vector<int> best_chk(chk), best_var(var);
    for (int it = 0; it < moves; ++it) {
        const double temp = 0.6 * std::pow(0.02, (double) it / moves);
        ga.clear(); gc.clear();
        const bool swap_chk = (rnd() & 1u) != 0;
        int a, b;
        if (swap_chk) {
            a = (int) (rnd() % (uint32_t) m);
            b = c_lo[a] + (int) (rnd() % (uint32_t) (c_hi[a] - c_lo[a]));
            if (a == b) continue;
            ga = {a / G, b / G};
            for (int x : {chk[a], chk[b]})
                for (int e = c->row_ptr[x]; e < c->row_ptr[x + 1]; ++e) gc.push_back(col_of_pos(pos_of_var[c->col_idx[e]]) / G);
        } else {
            a = (int) (rnd() % (uint32_t) n);
            b = v_lo[a] + (int) (rnd() % (uint32_t) (v_hi[a] - v_lo[a]));
            if (a == b) continue;
            gc = {col_of_pos(a) / G, col_of_pos(b) / G};
            for (int x : {var[a], var[b]})
                for (int q = c->col_ptr[x]; q < c->col_ptr[x + 1]; ++q) ga.push_back(rank_of_chk[edge_chk[c->csc_edge[q]]] / G);
        }
        uniq(ga); uniq(gc);
        auto total = [&]() { int t = 0; for (int g : ga) t += cost_a(g); for (int h : gc) t += cost_c(h); return t; };
        auto apply = [&]() {
            if (swap_chk) { std::swap(chk[a], chk[b]); rank_of_chk[chk[a]] = a; rank_of_chk[chk[b]] = b; }
            else { std::swap(var[a], var[b]); pos_of_var[var[a]] = a; pos_of_var[var[b]] = b; }
        };
        const int before = total();
        apply();
        const int delta = total() - before;
        if (delta > 0 && (rnd() & 0xffffff) / 16777216.0 >= std::exp(-delta / temp)) {
            apply();                                     // rejected: swap back
        } else {
            now += delta;
            if (now < best) { best = now; best_chk = chk; best_var = var; }
        }
    }
    chk = best_chk;                                      // the best state seen (the start if nothing beat it)
    var = best_var;
    for (int i = 0; i < m; ++i) rank_of_chk[chk[i]] = i;
    for (int i = 0; i < n; ++i) pos_of_var[var[i]] = i;
    if (stats)
        fprintf(stderr, "admm check-kernel layout F=%d: %ld -> %ld replayed wavefronts per iteration (%d moves)\n", F,
                replays_before, replays(), moves);
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
            t.n_chk = m;
            t.max_nb = std::max(1, cdeg(chk[0]) - 2);
            const int cols = chk_threads(c, F) / F;
            // block k of every check that has one: a plane of consecutive chunks indexed by check rank
            int base = 0;
            for (int k = 0; k < CHK_MAX_NB; ++k) {
                t.plane_base[k] = (uint32_t) base;
                int cnt = 0;
                for (int i = 0; i < m; ++i) cnt += std::max(1, cdeg(chk[i]) - 2) > k;    // one- and two-variable checks: one chunk
                base += (cnt + 1) & ~1;            // even bases: neighbouring checks write neighbouring rows
            }
            t.n_chunks = base;
            // codes too large for two frames per CTA run one CTA per SM, bound by shared-memory wavefronts: bank-aware ranks
            bool anneal = F == 1 && chk_threads(c, 2) > 640;
            if (const char *force = getenv("LDPC_ADMM_ANNEAL")) anneal = atoi(force) != 0;
            if (anneal) chk_anneal(c, F, cols, t.plane_base, chk, var);
            for (int i = 0; i < m; ++i) rank_of_chk[chk[i]] = i;
            // Variable slots: lane column c of the CTA updates the slots c, c + cols, c + 2 cols, ...  The variables
            // are dealt to the columns in boustrophedon order, heaviest first, so that every column gets about the
            // same number of incidences (the variable phase ends at a barrier) and neighbouring columns -- the lanes
            // of one warp -- get variables of equal degree.
            const int steps = (n + cols - 1) / cols;
            t.n_slots = steps * cols;
            std::vector<int> slot_of_var(n, 0), var_of_slot(t.n_slots, -1);
            for (int i = 0; i < n; ++i) {
                const int j = i / cols, k = i % cols;
                const int col = (j & 1) ? cols - 1 - k : k;
                slot_of_var[var[i]] = j * cols + col;
                var_of_slot[j * cols + col] = var[i];
            }
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
            // the same incidences as a stream per lane column (chk_stream_word): column c owns the slots c, c + cols, ...
            if ((size_t) (t.n_chunks + 1) * F * 16 > STR_OFF_MASK) t.supported = false;
            const uint32_t zero_off = (uint32_t) t.n_chunks * (uint32_t) (F * 16);   // a chunk of zeros behind the last one: NOADD words add it
            std::vector<std::vector<uint32_t>> col_words(cols);
            for (int col = 0; col < cols; ++col) {
                std::vector<uint32_t> &cwds = col_words[col];
                for (int j = 0; j < steps; ++j) {
                    const int sl = j * cols + col;
                    if (var_of_slot[sl] < 0) continue;           // only in the last row: nothing follows in this column
                    const uint32_t first = words[sl] >> 10, cnt = words[sl] & 15u, e = (words[sl] >> 4) & 63u;
                    if (cnt == 0) cwds.push_back(STR_FIRST | STR_LAST | STR_NOADD | (e << 20) | zero_off);
                    for (uint32_t q = 0; q < cnt; ++q) {
                        const uint4 &rec = inc[first + q];
                        uint32_t wd = (rec.x * (uint32_t) (F * 16)) | (e << 20);
                        wd |= (rec.y & 0x80000000u) | ((rec.z & 0x80000000u) >> 1) | ((rec.w & 0x80000000u) >> 2);
                        if (q == 0) wd |= STR_FIRST;
                        if (q + 1 == cnt) wd |= STR_LAST;
                        cwds.push_back(wd);
                    }
                }
                if (cwds.empty()) cwds.push_back(STR_NOADD | zero_off);
                cwds.back() |= STR_END;
            }
            size_t kmax = 0;
            for (const auto &cwds : col_words) kmax = std::max(kmax, cwds.size());
            t.stream_rows = (int) kmax + 2;                  // the kernel reads up to two words past a column's end
            std::vector<uint32_t> stream((size_t) t.stream_rows * cols, STR_NOADD | STR_END | zero_off);
            for (int col = 0; col < cols; ++col)
                for (size_t k = 0; k < col_words[col].size(); ++k) stream[k * cols + col] = col_words[col][k];
            t.e_min = e_min;
            int st;
            TableStager stage;
            stage.add(&t.chk_tab, tab);
            stage.add(&t.var_stream, stream);
            stage.add(&t.var_rank, vslot);
            stage.add(&t.var_e, se);
            if ((st = stage.commit(&t.blob))) return st;
        }
    }
    *out = &t;
    return LDPC_OK;
}

static size_t up16(size_t x) { return (x + 15) & ~(size_t) 15; }

// carve-up of the dynamic shared memory; returns the total
static size_t chk_smem_layout(const ldpc_code *c, const AdmmChkTables &t, int F, bool experiment, bool two, AdmmChkParams *p) {
    size_t off = 0;
    off += (size_t) (t.n_chunks + 1) * F * 16;           // w01 (+ the all-zero chunk)
    const size_t off_w23 = off; off += (size_t) (t.n_chunks + 1) * F * 16;
    const size_t off_v = off; off += (size_t) 2 * (t.n_slots + 1) * F * 8;
    const size_t off_qa = off; off += (size_t) t.n_slots * F * 8;
    const size_t off_inv = off; off += two ? 0 : (size_t) t.n_slots * 8;     // two CTAs per SM: inv_coef is looked up by e instead
    const size_t off_red = off; off += (size_t) F * 32 * 8;
    const size_t off_str = off; off += (size_t) t.stream_rows * (chk_threads(c, F) / F) * 4;
    const size_t off_cw = off; off += experiment ? (size_t) F * c->n : 0;
    const size_t off_ctl = up16(off); off = off_ctl + sizeof(ChkShared<4>);
    if (p) {
        p->off_w23 = (uint32_t) off_w23; p->off_v = (uint32_t) off_v; p->off_qa = (uint32_t) off_qa;
        p->off_inv = (uint32_t) off_inv; p->off_red = (uint32_t) off_red;
        p->off_str = (uint32_t) off_str; p->off_cw = (uint32_t) off_cw; p->off_ctl = (uint32_t) off_ctl;
    }
    return off + 16;
}

using ChkKernel = void (*)(const AdmmChkParams);

template <int F, bool GRID>
static ChkKernel chk_kernel_for(int nb, int shape) {
    if (shape == 1 && F == 1) {
        if (nb <= 2) return qpadmm_chk_kernel<1, 2, GRID, 1>;
        if (nb <= 4) return qpadmm_chk_kernel<1, 4, GRID, 1>;
    }
    if (shape == 2 && F == 1) {
        if (nb <= 4) return qpadmm_chk_kernel<1, 4, GRID, 2>;
        if (nb == 5) return qpadmm_chk_kernel<1, 5, GRID, 2>;
    }
    if (nb <= 2) return qpadmm_chk_kernel<F, 2, GRID, 0>;
    if (nb <= 4) return qpadmm_chk_kernel<F, 4, GRID, 0>;
    if (nb == 5) return qpadmm_chk_kernel<F, 5, GRID, 0>;
    if (nb == 6) return qpadmm_chk_kernel<F, 6, GRID, 0>;
    if (nb <= 8) return qpadmm_chk_kernel<F, 8, GRID, 0>;
    return qpadmm_chk_kernel<F, 10, GRID, 0>;
}

static ChkKernel chk_kernel_for(int F, int nb, bool grid, int shape) {
    if (grid) return F == 4 ? chk_kernel_for<4, true>(nb, shape) : (F == 2 ? chk_kernel_for<2, true>(nb, shape) : chk_kernel_for<1, true>(nb, shape));
    return F == 4 ? chk_kernel_for<4, false>(nb, shape) : (F == 2 ? chk_kernel_for<2, false>(nb, shape) : chk_kernel_for<1, false>(nb, shape));
}

// smallest 4 x column degree: DecodeQPADMM answers {zeros, false} when e_min * mu <= alpha (qp_admm.h:108-114)
int qpadmm_chk_e_min(const ldpc_code *c) {
    const AdmmChkTables *t = nullptr;
    if (get_chk_tables(c, 1, &t) || !t->supported) return -1;
    return t->e_min;
}

// host twins of inv_coef / aux_start / clip01_int (the same IEEE operations in the same order; volatile keeps the
// host compiler from contracting or reassociating them)
static double host_inv_coef(double mu, double alpha, double e) {
    volatile double t = mu * e;
    volatile double A = (t - alpha) * 0.5;
    volatile double d = 2.0 * A;
    return -1.0 / d;
}
static double host_aux_start(double mu, double half_alpha, double inv_aux) {
    volatile double w3 = mu * -2.0;                  // fma(mu, 0 - 2, 0): exact
    volatile double B = half_alpha + w3;             // the additions of +-0 in between leave the sum unchanged
    B = B + w3;
    volatile double v = B * inv_aux;
    if (!(v > 0.0)) return 0.0;
    return v >= 1.0 ? 1.0 : (double) v;
}

int launch_qpadmm_chk(const ldpc_code *c, const FrameIO &fio, int64_t frames, double var, double alpha, double mu,
                      int max_iter, double eps_stop, unsigned long long *queue, cudaStream_t stream,
                      const double *grid_alpha, const double *grid_mu, int64_t grid_points) {
    if (frames <= 0) return LDPC_OK;
    const bool grid = grid_points > 0;
    const int64_t frames_per_point = frames;
    if (grid) frames *= grid_points;             // work items of the queue
    // frames per CTA: one lane per (check, frame), at most 640 lanes.  Four CTAs of one frame per SM beat two of two and
    // one of four (profiles/r01_admm_chk_sweep.txt: 127.8 / 135.4 / 151 ms on optimalH, 124.6 / 129.2 / 150 ms on H05):
    // their barriers and their FP64-bound / latency-bound phases overlap.
    int F = 1;
    if (const char *force = getenv("LDPC_ADMM_F")) {
        const int v = atoi(force);
        if (v == 1 || v == 2 || v == 4) F = v;
    }
    const bool exp_mode = fio.experiment != 0;
    const AdmmChkTables *t = nullptr;
    for (;; F >>= 1) {
        int st = get_chk_tables(c, F, &t);
        if (st) return st;
        if (!t->supported) return LDPC_E_UNSUPPORTED;
        if (chk_threads(c, F) <= (t->max_nb <= 6 ? 640 : 320) && chk_smem_layout(c, *t, F, exp_mode, false, nullptr) <= 227 * 1024) break;
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
    p.chk_tab = t->chk_tab; p.var_slot = t->var_rank;
    for (int k = 0; k < CHK_MAX_NB; ++k) p.plane_base[k] = t->plane_base[k];
    p.special_lo = t->special_lo; p.special_hi = t->special_hi;
    p.n_chk = t->n_chk; p.n_slots = t->n_slots; p.n_chunks = t->n_chunks; p.tab_stride = t->tab_stride;
    p.max_iter = max_iter; p.alpha = alpha; p.mu = mu; p.eps_stop = eps_stop;
    p.half_mu = mu * 0.5; p.half_alpha = alpha * 0.5;
    p.inv_aux = host_inv_coef(mu, alpha, 8.0);
    p.aux_init = host_aux_start(mu, p.half_alpha, p.inv_aux);
    p.grid_alpha = grid_alpha; p.grid_mu = grid_mu; p.grid_frames = frames_per_point;
    const int threads = chk_threads(c, F);
    // one frame per CTA on 342..512 lanes: two CTAs per SM at 64 registers (a few spills) instead of one at 96 -- they
    // overlap their phases: 277.7 -> 254.7 ms per 16384 x 1000 frame-iterations on the (3,6)-1008 code
    bool two = F == 1 && threads > 341 && threads <= 512 && t->max_nb <= 4 &&
               2 * chk_smem_layout(c, *t, F, exp_mode, true, nullptr) <= 227 * 1024;
    if (const char *force = getenv("LDPC_ADMM_TWO")) two = two && atoi(force) != 0;
    // the 160 x 280 codes (one frame per CTA on at most 160 lanes): five CTAs per SM at 80 registers instead of four at 96
    // is SLOWER (optimalH 120.2 -> 124.6 ms, H05 120.2 -> 126.3 ms per 32768 x 1000 frame-iterations: the shared-memory
    // pipe is already 77 % busy and the spills add to it), so it is only an experiment knob (LDPC_ADMM_FIVE=1)
    bool five = false;
    if (const char *force = getenv("LDPC_ADMM_FIVE"))
        five = atoi(force) != 0 && !two && F == 1 && threads <= 160 && t->max_nb >= 3 && t->max_nb <= 5 &&
               5 * chk_smem_layout(c, *t, F, exp_mode, false, nullptr) <= 227 * 1024;
    const size_t smem = chk_smem_layout(c, *t, F, exp_mode, two, &p);
    p.var_stream = t->var_stream;
    p.stream_rows = t->stream_rows;
    p.slot_e = t->var_e;
    ChkKernel fn = chk_kernel_for(F, t->max_nb, grid, two ? 1 : (five ? 2 : 0));
    LDPC_CUDA(allow_max_dynamic_smem(fn));
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
        dev_free(t.blob);
    }
}

}  // namespace ldpc
