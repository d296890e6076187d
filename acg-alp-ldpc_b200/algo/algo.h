// The decoder interface of the reference (algo/algo.h:6-20), unchanged in shape:
// a per-frame virtual decode(H, channel_word, snr) -> {codeword, ok} plus name().
// GpuDecoder is the extension the batched experiment path looks for.
#ifndef LDPC_B200_ALGO_ALGO_H
#define LDPC_B200_ALGO_ALGO_H

#include "../utils/channel.h"
#include "ldpc_b200.h"

class Decoder {
public:
    virtual pair<TCodeword, bool> decode(const TMatrix &H, const TFVector &channel_word, double snr) = 0;

    virtual string name() const = 0;

    virtual ~Decoder() = default;
};

// Decoders that run on the GPU expose their configuration so that
// multithread_experiment() can hand whole Monte-Carlo points to the device
// instead of calling decode() frame by frame.
class GpuDecoder : public Decoder {
public:
    virtual ldpc_algo_cfg_t config() const = 0;
};

// y -> LLR, algo/algo.h:13-20
inline vector<double> CalculateCoef(const vector<double> &y, double snr) {
    vector<double> coef(y.size());
    for (size_t i = 0; i < y.size(); ++i) coef[i] = llr(y[i], snr);
    return coef;
}

#endif
