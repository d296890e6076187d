// Device-side channel model: counter-based Philox4x32-10 AWGN and information bits.
//
// Replaces transmit() / gen_random_codeword() (utils/channel.h:19-36), which draw
// from mt19937 + std::normal_distribution on the host, by a stream addressed with
// the GLOBAL frame index, so any sharding over GPUs produces the same frames:
//     key     = (seed lo, seed hi)
//     counter = (frame lo, frame hi, block, stream)     stream 0: info bits, 1: noise
// One noise block = two standard normals (Box-Muller).  Every floating-point step
// below is a single correctly rounded IEEE operation written with the _rn
// intrinsics (nvcc never contracts those), in a fixed order, so a CPU can replay
// y bit for bit (the test oracle does).
#ifndef LDPC_B200_CHANNEL_CUH
#define LDPC_B200_CHANNEL_CUH

#include <cstdint>

namespace ldpc {

enum { STREAM_INFO = 0, STREAM_NOISE = 1 };

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// 52 random bits -> (0,1) exclusive: (a + 1/2) * 2^-52, all exact
__device__ __forceinline__ double unit_open(uint32_t hi, uint32_t lo) {
    const unsigned long long a = ((unsigned long long) hi << 20) | (lo >> 12);
    return __dmul_rn(__dadd_rn((double) a, 0.5), 0x1p-52);
}

// ln(u), u in (0,1): u = 2^ex * mant, mant in [sqrt(1/2), sqrt(2)), ln(mant) = 2 atanh(s),
// s = (mant-1)/(mant+1), odd series to s^23
__device__ __forceinline__ double det_log(double u) {
    int hi = __double2hiint(u);
    int ex = ((hi >> 20) & 0x7ff) - 1023;
    double mant = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(u));
    if (mant > 1.4142135623730951) {
        mant = __dmul_rn(mant, 0.5);
        ex += 1;
    }
    const double f = __dadd_rn(mant, -1.0);
    const double s = __ddiv_rn(f, __dadd_rn(2.0, f));
    const double s2 = __dmul_rn(s, s);
    double p = 0x1.642c8590b2164p-5;               // 1/23
    p = __fma_rn(p, s2, 0x1.8618618618618p-5);     // 1/21
    p = __fma_rn(p, s2, 0x1.af286bca1af28p-5);     // 1/19
    p = __fma_rn(p, s2, 0x1.e1e1e1e1e1e1ep-5);     // 1/17
    p = __fma_rn(p, s2, 0x1.1111111111111p-4);     // 1/15
    p = __fma_rn(p, s2, 0x1.3b13b13b13b14p-4);     // 1/13
    p = __fma_rn(p, s2, 0x1.745d1745d1746p-4);     // 1/11
    p = __fma_rn(p, s2, 0x1.c71c71c71c71cp-4);     // 1/9
    p = __fma_rn(p, s2, 0x1.2492492492492p-3);     // 1/7
    p = __fma_rn(p, s2, 0x1.999999999999ap-3);     // 1/5
    p = __fma_rn(p, s2, 0x1.5555555555555p-2);     // 1/3
    const double two_s = __dadd_rn(s, s);
    const double lm = __fma_rn(two_s, __dmul_rn(p, s2), two_s);
    return __fma_rn((double) ex, 0x1.62e42fefa39efp-1, lm);
}

// sin and cos of 2*pi*u, u in (0,1): quadrant reduction is exact, then Taylor to x^19 / x^18
__device__ __forceinline__ void det_sincos2pi(double u, double &sn, double &cs) {
    const double t = __dmul_rn(u, 4.0);
    const int quad = (int) __dadd_rn(t, 0.5);
    const double f = __dadd_rn(t, -(double) quad);
    const double x = __dmul_rn(f, 0x1.921fb54442d18p+0);
    const double x2 = __dmul_rn(x, x);
    double ps = -0x1.2f49b46814157p-57;
    ps = __fma_rn(ps, x2, 0x1.952c77030ad4ap-49);
    ps = __fma_rn(ps, x2, -0x1.ae7f3e733b81fp-41);
    ps = __fma_rn(ps, x2, 0x1.6124613a86d09p-33);
    ps = __fma_rn(ps, x2, -0x1.ae64567f544e4p-26);
    ps = __fma_rn(ps, x2, 0x1.71de3a556c734p-19);
    ps = __fma_rn(ps, x2, -0x1.a01a01a01a01ap-13);
    ps = __fma_rn(ps, x2, 0x1.1111111111111p-7);
    ps = __fma_rn(ps, x2, -0x1.5555555555555p-3);
    const double s = __fma_rn(__dmul_rn(x, x2), ps, x);
    double pc = 0x1.6827863b97d97p-53;
    pc = __fma_rn(pc, x2, -0x1.ae7f3e733b81fp-45);
    pc = __fma_rn(pc, x2, 0x1.93974a8c07c9dp-37);
    pc = __fma_rn(pc, x2, -0x1.1eed8eff8d898p-29);
    pc = __fma_rn(pc, x2, 0x1.27e4fb7789f5cp-22);
    pc = __fma_rn(pc, x2, -0x1.a01a01a01a01ap-16);
    pc = __fma_rn(pc, x2, 0x1.6c16c16c16c17p-10);
    pc = __fma_rn(pc, x2, -0x1.5555555555555p-5);
    pc = __fma_rn(pc, x2, 0x1p-1);
    const double c = __fma_rn(-x2, pc, 1.0);
    switch (quad & 3) {
        case 0: sn = s; cs = c; break;
        case 1: sn = c; cs = -s; break;
        case 2: sn = -s; cs = -c; break;
        default: sn = -c; cs = s; break;
    }
}

// the two standard normals of noise block `blk` of global frame `frame`
__device__ __forceinline__ void noise_pair(uint64_t seed, uint64_t frame, uint32_t blk, double &z0, double &z1) {
    const uint4 w = philox4x32_10(make_uint4((uint32_t) frame, (uint32_t) (frame >> 32), blk, STREAM_NOISE),
                                  make_uint2((uint32_t) seed, (uint32_t) (seed >> 32)));
    const double u1 = unit_open(w.x, w.y);
    const double u2 = unit_open(w.z, w.w);
    const double r = __dsqrt_rn(__dmul_rn(-2.0, det_log(u1)));
    double sn, cs;
    det_sincos2pi(u2, sn, cs);
    z0 = __dmul_rn(r, cs);
    z1 = __dmul_rn(r, sn);
}

// 128 information bits: block `blk` of stream 0
__device__ __forceinline__ uint4 info_block(uint64_t seed, uint64_t frame, uint32_t blk) {
    return philox4x32_10(make_uint4((uint32_t) frame, (uint32_t) (frame >> 32), blk, STREAM_INFO),
                         make_uint2((uint32_t) seed, (uint32_t) (seed >> 32)));
}

}  // namespace ldpc

#endif
