// Grid search over the QP-ADMM penalty parameters (alpha, mu) on data/optimalH -- the
// reference's qpadmm_params.cpp (:12-84) with the FER evaluations on the GPU.
// Same grid (61 x 61 over [0,3]^2, alpha outer / mu inner), same 1000 codewords
// from mt19937(239), SNR -3 dB, QPADMMDecoder(alpha, mu, 1000, 1e-5), same "first
// strictly smaller FER wins" rule and the same stdout / stderr lines.
//
// Environment: LDPC_GRID (points per axis, default 61), LDPC_TESTS_NUM (default 1000).
#include <memory>
#include <utility>

#include "experiment.h"
#include "utils/parse_data.h"
#include "algo/algo.h"
#include "algo/qp_admm.h"

using namespace std;

const int THREADS_NUM = 8;
const int TESTS_NUM = 1000;

// FER of QPADMMDecoder(alpha, mu, 1000, 1e-5) at `snr` over the given codewords
double estimate_qpadmm(const TMatrix &H, const vector<TCodeword> &codewords, double snr, double alpha, double mu) {
    auto decoder = make_shared<QPADMMDecoder>(alpha, mu, 1000, 1e-5);
    return multithread_experiment(decoder, codewords, H, snr, THREADS_NUM).FER();
}

// i-th of cnt equally spaced points of [L, R], evaluated as the reference does (qpadmm_params.cpp:32-34)
double linear_function(double L, double R, int cnt, int i) { return L + ((R - L) / (cnt - 1)) * i; }

int main() {
    std::ios::sync_with_stdio(0);
    cout.precision(5);
    cout << fixed;

    TMatrix H = load_matrix("data/optimalH");
    TMatrix G = GetOrtogonal(H).first;
    if (H.empty() || G.empty()) {
        cerr << "cannot load data/optimalH" << endl;
        return 1;
    }
    const int tests_num = getenv("LDPC_TESTS_NUM") ? atoi(getenv("LDPC_TESTS_NUM")) : TESTS_NUM;
    const int grid = getenv("LDPC_GRID") ? max(2, atoi(getenv("LDPC_GRID"))) : 61;

    mt19937 rnd(239);
    vector<TCodeword> codewords = gen_random_codewords(G, tests_num, rnd);
    cerr << "n=" << H[0].size() << " k=" << H.size() << endl;

    const double alpha_l = 0, alpha_r = 3.0, mu_l = 0, mu_r = 3.0, snr = -3.0;
    // all pairs of the grid in the reference's scan order (alpha outer, mu inner), evaluated in one launch per GPU
    // (LDPC_GRID_BATCHED=0: one launch per pair, the reference's structure)
    vector<double> alphas, mus;
    for (int ai = 0; ai < grid; ++ai)
        for (int mi = 0; mi < grid; ++mi) {
            alphas.push_back(linear_function(alpha_l, alpha_r, grid, ai));
            mus.push_back(linear_function(mu_l, mu_r, grid, mi));
        }
    const bool batched = !(getenv("LDPC_GRID_BATCHED") && atoi(getenv("LDPC_GRID_BATCHED")) == 0);
    vector<ExperimentResult> results;
    if (batched) results = ldpc_host::gpu_qpadmm_grid(alphas, mus, 1000, 1e-5, codewords, H, snr);
    double best_fer = 2.0, best_alpha = -1, best_mu = -1;
    for (size_t i = 0; i < alphas.size(); ++i) {
        const double alpha = alphas[i], mu = mus[i];
        const double fer = batched ? results[i].FER() : estimate_qpadmm(H, codewords, snr, alpha, mu);
        cerr << "alpha=" << alpha << ", mu=" << mu << ": fer=" << fer << endl;
        if (fer < best_fer) {
            best_fer = fer;
            best_alpha = alpha;
            best_mu = mu;
            cout << "new best fer found: " << fer << "| alpha=" << alpha << ", mu=" << mu << endl;
        }
    }

    cout << "Best parameters:" << endl;
    cout << "alpha=" << best_alpha << endl;
    cout << "mu=" << best_mu << endl;
    cout << "fer=" << best_fer << endl;
    return 0;
}
