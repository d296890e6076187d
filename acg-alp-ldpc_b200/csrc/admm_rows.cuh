// Row arithmetic of the QP-ADMM iteration shared by the two kernels (qpadmm_kernel.cu: one lane per
// three-variable block, any check degree; qpadmm_chk_kernel.cu: one lane per check, state in registers).
// Every operation is an _rn intrinsic in the reference's order (algo/qp_admm.h:130-164): the iteration
// is chaotic, so the results are only reproducible bit for bit.
#ifndef LDPC_B200_ADMM_ROWS_CUH
#define LDPC_B200_ADMM_ROWS_CUH

#include <cstdint>

namespace ldpc {

// x with its sign flipped when bit 31 of `word` is set
__device__ __forceinline__ double flip_by(double x, uint32_t word) {
    return __hiloint2double(__double2hiint(x) ^ (int) (word & 0x80000000u), __double2loint(x));
}

// r[q] = b[q] - sum_k cf(q, S_k) * vs[k] with the block's variables visited in
// ascending index order; S_k = slot of the k-th visited variable; cf(q, s) = +1 if
// q == 3 or q == s, else -1.  The leading "0 -/+ v" of rows 0..2 is folded into a
// sign (differs from the reference only in the sign of an exact zero).
template <int S0, int S1, int S2>
__device__ __forceinline__ void residual_rows(double v0, double v1, double v2, double b3, double &r0, double &r1,
                                              double &r2, double &r3) {
    // row q: x = (q == 3 ? b3 - v0 : -+v0), then -+ v1, then -+ v2, minus where cf = +1
    r0 = __dadd_rn(__dadd_rn(S0 == 0 ? -v0 : v0, S1 == 0 ? -v1 : v1), S2 == 0 ? -v2 : v2);
    r1 = __dadd_rn(__dadd_rn(S0 == 1 ? -v0 : v0, S1 == 1 ? -v1 : v1), S2 == 1 ? -v2 : v2);
    r2 = __dadd_rn(__dadd_rn(S0 == 2 ? -v0 : v0, S1 == 2 ? -v1 : v1), S2 == 2 ? -v2 : v2);
    r3 = __dadd_rn(__dadd_rn(__dadd_rn(b3, -v0), -v1), -v2);
}

// One inequality row (qp_admm.h:154-159): yl is the dual of the previous iteration (state), r the new
// residual.  t = r - yl; z = max(0, t); yl' = max(0, -t) (== max(0, yl - r) exactly); stop-sum term (z - r)^2;
// w = yl' + mu (z - b).  Returns w, updates yl and part.
template <bool ROW3>
__device__ __forceinline__ double row_update(double r, double &yl, double &part, double mu, double b3) {
    const double t = __dadd_rn(r, -yl);
    // both maxima from the sign bit alone, on the integer pipe (bit-identical to std::max, signed zeros included)
    const int hi = __double2hiint(t), lo = __double2loint(t);
    const int neg = hi >> 31;                                          // all ones iff t < 0 (or t == -0)
    const double z = __hiloint2double(hi & ~neg, lo & ~neg);           // max(0, t)
    yl = __hiloint2double(hi & neg & 0x7fffffff, lo & neg);            // max(0, -t)
    const double d = __dadd_rn(z, -r);
    part = __fma_rn(d, d, part);
    return __fma_rn(mu, ROW3 ? __dadd_rn(z, -b3) : z, yl);
}

// The same row on the FP64 pipe alone (six instructions, no integer work): with s = t + |t| (= 2 max(0, t),
// exact; the |.| is an operand modifier), z = s/2 is never formed -- every use folds the exact factor 1/2 into a
// fused multiply-add, which rounds the same real number the reference rounds:
//   d = z - r = fma(1/2, s, -r),   yl' = max(0, -t) = z - t = fma(1/2, s, -t)  (exact),
//   w = yl' + mu z = fma(mu/2, s, yl'),   row 3: w = yl' + mu (z - 2) = fma(mu, fma(1/2, s, -2), yl').
// half_mu = mu / 2 (exact).  Bit-identical to row_update (tests/test_gpu_parity.py compares v bit for bit).
template <bool ROW3>
__device__ __forceinline__ double row_update_fp(double r, double &yl, double &part, double mu, double half_mu) {
    const double t = __dadd_rn(r, -yl);
    const double s = __dadd_rn(t, fabs(t));
    const double d = __fma_rn(0.5, s, -r);
    part = __fma_rn(d, d, part);
    yl = __fma_rn(0.5, s, -t);
    return ROW3 ? __fma_rn(mu, __fma_rn(0.5, s, -2.0), yl) : __fma_rn(half_mu, s, yl);
}

// std::max(v, 0.0) then std::min(v, 1.0) (qp_admm.h:140-141) on the high word: negative (sign bit set) -> +0,
// >= 1.0 -> 1.0, else unchanged
__device__ __forceinline__ double clip01_int(double x) {
    const int hi = __double2hiint(x);
    const bool keep = (unsigned) hi < 0x3ff00000u;                     // 0 <= x < 1
    return __hiloint2double(min(max(hi, 0), 0x3ff00000), keep ? __double2loint(x) : 0);
}

}  // namespace ldpc

#endif
