// BPSK + AWGN channel helpers with the reference's names (utils/channel.h:8-44).
// These are the HOST versions (mt19937 + std::normal_distribution) used by the
// per-frame Decoder path and by drivers that bring their own generator; the
// batched experiment path generates the same model on the GPU from a Philox
// stream (csrc/channel.cuh).
#ifndef LDPC_B200_UTILS_CHANNEL_H
#define LDPC_B200_UTILS_CHANNEL_H

#include <cmath>

#include "codeword.h"

using namespace std;

typedef vector<double> TFVector;

const double EPS = 1e-8;

// sigma^2 for Es/N0 = snr dB, Es = 1 (utils/channel.h:12)
inline double llr_variance(double snr) { return pow(10, -(snr / 10)) / 2; }

// log-likelihood ratio of one received sample (utils/channel.h:14-16)
inline double llr(double v, double snr) { return 2 * v / llr_variance(snr); }

// y_i = (c_i ? -1 : +1) + N(0, sigma^2)   (utils/channel.h:19-26)
template <typename Gen>
TFVector transmit(double snr, const TCodeword &c, Gen &rnd) {
    normal_distribution<double> noise(0, sqrt(llr_variance(snr)));
    TFVector y(c.size());
    for (size_t i = 0; i < c.size(); ++i) {
        const double symbol = c[i] ? -1.0 : 1.0;
        y[i] = symbol + noise(rnd);
    }
    return y;
}

// random combination of the rows of G: row i is taken when rnd() is even
// (utils/channel.h:29-36)
template <typename Gen>
TCodeword gen_random_codeword(const vector<TCodeword> &G, Gen &rnd) {
    assert(!G.empty());
    gf2::Packed acc((G[0].size() + 63) / 64, 0);
    for (const TCodeword &row : G)
        if (rnd() % 2 == 0) gf2::xor_into(acc, gf2::pack(row));
    return gf2::unpack(acc, G[0].size());
}

// n codewords (utils/channel.h:38-44); the rows of G are packed once, the generator is consumed exactly as by n
// calls of gen_random_codeword (one draw per row, row order)
template <typename Gen>
vector<TCodeword> gen_random_codewords(const TMatrix &G, int n, Gen &rnd) {
    assert(!G.empty());
    vector<gf2::Packed> rows;
    rows.reserve(G.size());
    for (const TCodeword &row : G) rows.push_back(gf2::pack(row));
    vector<TCodeword> words;
    words.reserve(n);
    while ((int) words.size() < n) {
        gf2::Packed acc((G[0].size() + 63) / 64, 0);
        for (const gf2::Packed &row : rows)
            if (rnd() % 2 == 0) gf2::xor_into(acc, row);
        words.push_back(gf2::unpack(acc, G[0].size()));
    }
    return words;
}

#endif
