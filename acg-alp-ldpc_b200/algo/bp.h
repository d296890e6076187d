// BeliefPropagationDecoder with the reference's constructor and semantics
// (algo/bp.h:208-222): flooding sum-product, at most max_iter iterations, exits on
// a zero syndrome, returns {EMPTY codeword, false} when it never converges
// (bp.h:198).  The work happens in the CUDA kernel behind ldpc_bp_decode.
#ifndef LDPC_B200_ALGO_BP_H
#define LDPC_B200_ALGO_BP_H

#include "gpu_code.h"

class BeliefPropagationDecoder : public GpuDecoder {
public:
    explicit BeliefPropagationDecoder(int max_iter) : _max_iter(max_iter) {}

    pair<TCodeword, bool> decode(const TMatrix &H, const TFVector &channel_word, double snr) override {
        ldpc_host::CodeRef code = ldpc_host::CodeCache::instance().get(H);
        const size_t n = channel_word.size();
        vector<uint8_t> bits(n);
        uint8_t ok = 0;
        int32_t iters = 0;
        if (ldpc_bp_decode(code.get(), channel_word.data(), 1, snr, _max_iter, 1, bits.data(), &ok, &iters, nullptr))
            ldpc_host::die("ldpc_bp_decode");
        if (!ok) return {TCodeword(), false};
        return {TCodeword(bits.begin(), bits.end()), true};
    }

    string name() const override { return "BP"; }

    ldpc_algo_cfg_t config() const override { return ldpc_algo_cfg_t{LDPC_ALGO_BP, _max_iter, 1, 0, 0.0, 0.0, 0.0}; }

private:
    int _max_iter;
};

#endif
