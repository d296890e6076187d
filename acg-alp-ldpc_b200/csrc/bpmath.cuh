// fp64 transcendental kernels of the BP message update, written for instruction
// count: the decoder is bound by the FP64 pipe (64 lanes/SM) and by issue slots,
// so both functions are straight-line DFMA chains with a handful of integer ops,
// no special-case branches and no libm calls.
#ifndef LDPC_B200_BPMATH_CUH
#define LDPC_B200_BPMATH_CUH

namespace ldpc {

// Polynomial coefficients live in constant memory: an FP64 instruction reads a constant-bank
// operand directly, whereas a 64-bit immediate costs two extra move instructions.
static __constant__ double K_EXP[11] = {
    0x1.1eed8eff8d898p-29,   // 1/12!
    0x1.ae64567f544e4p-26,   // 1/11!
    0x1.27e4fb7789f5cp-22,   // 1/10!
    0x1.71de3a556c734p-19,   // 1/9!
    0x1.a01a01a01a01ap-16,   // 1/8!
    0x1.a01a01a01a01ap-13,   // 1/7!
    0x1.6c16c16c16c17p-10,   // 1/6!
    0x1.1111111111111p-7,    // 1/5!
    0x1.5555555555555p-5,    // 1/4!
    0x1.5555555555555p-3,    // 1/3!
    0.5};
static __constant__ double K_EXP_RED[3] = {-1.4426950408889634, -0x1.62e42fefa0000p-1, -0x1.cf79abc9e3b3ap-40};
static __constant__ double K_ATANH[9] = {
    0x1.af286bca1af28p-5,    // 1/19
    0x1.e1e1e1e1e1e1ep-5,    // 1/17
    0x1.1111111111111p-4,    // 1/15
    0x1.3b13b13b13b14p-4,    // 1/13
    0x1.745d1745d1746p-4,    // 1/11
    0x1.c71c71c71c71cp-4,    // 1/9
    0x1.2492492492492p-3,    // 1/7
    0x1.999999999999ap-3,    // 1/5
    0x1.5555555555555p-2};   // 1/3
static __constant__ double K_LOG[3] = {1.4142135623730951, 0x1.62e42fefa39efp-1, 0x1.62e42fefa39f0p-2};

// exp(-|t|) (t = +-inf allowed), relative error < 2 ulp.
// |t| is clamped to 700 on its high word, so the result never underflows to a
// denormal and no message magnitude downstream reaches infinity:
//   k = rint(-a log2 e) through the 1.5*2^52 trick, r = -a - k ln2 (two-part ln2),
//   degree-12 Taylor polynomial on |r| <= ln2/2, exponent patched in.
__device__ __forceinline__ double exp_neg_abs(double t) {
    // |t| clamped to 700, both on the high word (integer pipe, no FP64 instruction)
    const int hi_a = min(__double2hiint(t) & 0x7fffffff, 0x4085e000);
    const double a = __hiloint2double(hi_a, __double2loint(t));
    const double magic = 6755399441055744.0;                      // 1.5 * 2^52
    double kf = __fma_rn(a, K_EXP_RED[0], magic);
    const int k = __double2loint(kf);
    kf -= magic;
    double r = __fma_rn(kf, K_EXP_RED[1], -a);                    // ln2 high part (trailing bits zero)
    r = __fma_rn(kf, K_EXP_RED[2], r);                            // ln2 low part
    double p = K_EXP[0];
#pragma unroll
    for (int i = 1; i < 11; ++i) p = __fma_rn(p, r, K_EXP[i]);
    p = __fma_rn(p, r, 1.0);
    p = __fma_rn(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// exp(-a) for a >= 0
__device__ __forceinline__ double exp_neg(double a) { return exp_neg_abs(a); }

// exp(t) for |t| <= 700 (no clamp, no abs): same reduction and polynomial as exp_neg_abs
__device__ __forceinline__ double exp_signed(double t) {
    const double magic = 6755399441055744.0;                      // 1.5 * 2^52
    const double a = -t;
    double kf = __fma_rn(a, K_EXP_RED[0], magic);
    const int k = __double2loint(kf);
    kf -= magic;
    double r = __fma_rn(kf, K_EXP_RED[1], -a);
    r = __fma_rn(kf, K_EXP_RED[2], r);
    double p = K_EXP[0];
#pragma unroll
    for (int i = 1; i < 11; ++i) p = __fma_rn(p, r, K_EXP[i]);
    p = __fma_rn(p, r, 1.0);
    p = __fma_rn(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

__device__ __forceinline__ double rcp_fast(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));        // MUFU.RCP64H, ~2^-23
    double e = __fma_rn(-d, r, 1.0);
    r = __fma_rn(r, e, r);
    e = __fma_rn(-d, r, 1.0);
    return __fma_rn(r, e, r);
}

// log(ev / od) for ev >= od >= 0, ev in [1, 2^60): absolute error ~1e-14.
// od is clamped to the smallest normal, so od == 0 (a degree-1 check, or every
// other input saturated) gives ~log(ev) + 708 instead of +inf.
//   ev = 2^ke me, od = 2^ko mo, me, mo in [1,2);  if me < mo: me *= 2, ke -= 1, so me/mo in [1,2)
//   log(me/mo) = log(c) + 2 atanh(s),  c = sqrt(2) rounded,  s = (me - c mo)/(me + c mo),  |s| <= 0.1716
__device__ __forceinline__ double log_ratio(double ev, double od) {
    int hi_e = __double2hiint(ev);
    int hi_o = max(__double2hiint(od), 0x00100000);
    int d = (hi_e >> 20) - (hi_o >> 20);
    hi_e = (hi_e & 0x000fffff) | 0x3ff00000;
    hi_o = (hi_o & 0x000fffff) | 0x3ff00000;
    double me = __hiloint2double(hi_e, __double2loint(ev));
    const double mo = __hiloint2double(hi_o, __double2loint(od));
    if (me < mo) {
        me = __hiloint2double(hi_e + 0x00100000, __double2loint(ev));
        d -= 1;
    }
    const double c = K_LOG[0];
    const double num = __fma_rn(-c, mo, me);
    const double den = __fma_rn(c, mo, me);
    const double s = num * rcp_fast(den);
    const double s2 = s * s;
    double p = K_ATANH[0];
#pragma unroll
    for (int i = 1; i < 9; ++i) p = __fma_rn(p, s2, K_ATANH[i]);
    const double two_s = s + s;
    const double u = __fma_rn(two_s, s2 * p, two_s);
    // log(c) rounds to 0x1.62e42fefa39f0p-2 (c is sqrt(2) rounded to double; residual 2.4e-17)
    return __fma_rn((double) d, K_LOG[1], K_LOG[2] + u);
}

// log(x) for a positive normal x (used once per variable when a frame's soft output is written)
__device__ __forceinline__ double log_pos(double x) {
    return x >= 1.0 ? log_ratio(x, 1.0) : -log_ratio(1.0, x);
}

// a / b for positive normal a, b with b in [2^-1000, 2^1000], four FP64 instructions behind the MUFU seed:
//   r0 = 1/b (1 - e) with |e| <= 2^-22 (MUFU.RCP64H), so a/b = a r0 / (1 - e) = q0 (1 + e + e^2 + e^3 + ...);
//   q = q0 + q0 (e + e^2) drops e^3 <= 2^-66 and carries the roundings of q0 and of the final fma: ~1 ulp, no special cases.
// (The Newton form -- refine r, multiply, correct the quotient's residual -- takes five.)
__device__ __forceinline__ double div_pos(double a, double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    const double q0 = a * r;
    const double e = __fma_rn(-b, r, 1.0);
    const double e2 = __fma_rn(e, e, e);
    return __fma_rn(q0, e2, q0);
}

}  // namespace ldpc

#endif
