// FER / time sweep of the decoders over an SNR grid -> stdout, log.txt (stderr),
// report.csv.  Same flow, constants and output formats as the reference's main.cpp
// (:23-91), with the GPU-backed BP and QP-ADMM decoders; the glpk-based decoders
// (ALP, AGC-ALP) are out of scope and not listed.
//
// Environment overrides (defaults = the reference's constants):
//   LDPC_TESTS_NUM  frames per SNR point (10000)      LDPC_GPUS  number of GPUs to use (all)
//   LDPC_MATRIX     optimalH | H05 (optimalH)          LDPC_SEED  noise / codeword stream seed
#define OPTIMAL

#include <iomanip>
#include <memory>
#include <utility>

#include "experiment.h"
#include "utils/parse_data.h"
#include "utils/codeword.h"
#include "algo/algo.h"
#include "algo/bp.h"
#include "algo/qp_admm.h"

using namespace std;

const int THREADS_NUM = 8;
const int LOG_FREQ = 1000000;
const int TESTS_NUM = 10000;

const vector<double> SNRS = {-5, -4.5, -4, -3.5, -3, -2.5, -2, -1.5, -1, -0.5, 0.0};

int main() {
    std::ios::sync_with_stdio(0);
    cout.precision(5);
    cout << fixed;

    const char *matrix_env = getenv("LDPC_MATRIX");
#ifdef OPTIMAL
    const string matrix = matrix_env ? matrix_env : "optimalH";
#else
    const string matrix = matrix_env ? matrix_env : "H05";
#endif
    const int tests_num = getenv("LDPC_TESTS_NUM") ? atoi(getenv("LDPC_TESTS_NUM")) : TESTS_NUM;

    // alpha/mu as main.cpp:30-34 pairs them with the two matrices
    vector<shared_ptr<Decoder>> decoders{make_shared<BeliefPropagationDecoder>(100)};
    if (matrix == "optimalH") decoders.push_back(make_shared<QPADMMDecoder>(1.2, 0.55, 10000, 1e-5));
    else decoders.push_back(make_shared<QPADMMDecoder>(1.95, 0.5, 10000, 1e-5));

    ofstream fdata("report.csv");
    fdata << "Method,SNR,Sigma,FER,Time,AvgHamming,AvgHammingCorrect,AvgHammingWrong" << endl;
    fdata << fixed << setprecision(12);

    for (auto snr : SNRS) cerr << "snr=" << snr << ": var=" << llr_variance(snr) << endl;

    TMatrix H = load_matrix("data/" + matrix);
    TMatrix G = (matrix == "H05") ? load_matrix("data/G05") : GetOrtogonal(H).first;
    if (H.empty() || G.empty()) {
        cerr << "cannot load data/" << matrix << endl;
        return 1;
    }

    mt19937 rnd(239'239'239);
    vector<TCodeword> codewords = gen_random_codewords(G, tests_num, rnd);

    cerr << "n=" << H[0].size() << " k=" << H.size() << "\n";

    for (auto decoder : decoders) {
        cout << "Algo: " << decoder->name() << endl;
        for (double snr : SNRS) {
            ExperimentResult res = multithread_experiment(decoder, codewords, H, snr, THREADS_NUM, LOG_FREQ);

            cout << "\tSNR: " << snr << ", FER: " << res.FER() << ", (time=" << res.avg_time() << "s)" << endl;
            cerr << "\tSNR: " << snr << ", FER: " << res.FER() << ", (time=" << res.avg_time() << "s)" << endl;
            cerr << "\t\tAverage hamming distance: " << res.mean_hamming() << std::endl;
            cerr << "\t\tAverage hamming distance, correctly decoded: " << res.mean_hamming_ok() << endl;
            cerr << "\t\tAverage hamming distance, incorrectly decoded: " << res.mean_hamming_wrong() << endl;
            cerr << "\t\tBER (frames that returned bits): " << scientific << res.BER(H[0].size()) << fixed
                 << ", mean iterations: " << (double) res.sum_iters / res.total << endl;

            fdata << decoder->name() << "," << snr << "," << sqrt(llr_variance(snr)) << "," << res.FER() << ","
                  << res.avg_time() << "," << res.mean_hamming() << "," << res.mean_hamming_ok() << ","
                  << res.mean_hamming_wrong() << endl;
        }
        cerr << string(30, '_') << endl;
    }
    return 0;
}
