// QPADMMDecoder with the reference's constructor and semantics
// (algo/qp_admm.h:180-194 -> DecodeQPADMM :104-178): penalised-QP ADMM on the
// three-variable-check decomposition, fp64, stops when the squared residual drops
// below eps_stop or after max_iter iterations; {zeros, false} when
// min(e) * mu <= alpha.  The work happens in the CUDA kernel behind ldpc_qpadmm_decode.
#ifndef LDPC_B200_ALGO_QP_ADMM_H
#define LDPC_B200_ALGO_QP_ADMM_H

#include "gpu_code.h"

class QPADMMDecoder : public GpuDecoder {
public:
    explicit QPADMMDecoder(double alpha, double mu, int max_iter = 2000, double eps_stop = 1e-5)
        : _alpha(alpha), _mu(mu), _eps_stop(eps_stop), _max_iter(max_iter) {}

    pair<TCodeword, bool> decode(const TMatrix &H, const TFVector &channel_word, double snr) override {
        ldpc_host::CodeRef code = ldpc_host::CodeCache::instance().get(H);
        const size_t n = channel_word.size();
        vector<uint8_t> bits(n);
        uint8_t ok = 0;
        int32_t iters = 0;
        if (ldpc_qpadmm_decode(code.get(), channel_word.data(), 1, snr, _alpha, _mu, _max_iter, _eps_stop, bits.data(), &ok,
                               &iters, nullptr))
            ldpc_host::die("ldpc_qpadmm_decode");
        return {TCodeword(bits.begin(), bits.end()), ok != 0};
    }

    string name() const override { return "QP-ADMM"; }

    ldpc_algo_cfg_t config() const override {
        return ldpc_algo_cfg_t{LDPC_ALGO_QPADMM, _max_iter, 1, 0, _alpha, _mu, _eps_stop};
    }

private:
    double _alpha, _mu, _eps_stop;
    int _max_iter;
};

#endif
