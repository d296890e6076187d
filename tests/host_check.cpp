// Test helper (compiled by tests/test_host_cpp.py): exercises the host-side mirror
// of the reference's utils/ and experiment.h WITHOUT touching the GPU, and prints
// results for the Python side to compare with the golden fixtures.
#include <cinttypes>
#include <cstring>

#include "experiment.h"
#include "utils/parse_data.h"

// a CPU stand-in decoder: hard decision of the channel word, "ok" iff it is a codeword
class HardDecision : public Decoder {
public:
    pair<TCodeword, bool> decode(const TMatrix &H, const TFVector &y, double) override {
        TCodeword c(y.size());
        for (size_t i = 0; i < y.size(); ++i) c[i] = y[i] <= 0;
        if (!IsCodeword(H, c)) return {TCodeword(), false};
        return {c, true};
    }
    string name() const override { return "HARD"; }
};

int main(int argc, char **argv) {
    const string dir = argv[1];
    TMatrix H = read_pcm_rows(dir + "/optimalH.rows");
    cout << "shape " << H.size() << " " << H[0].size() << "\n";
    auto orth = GetOrtogonal(H);
    cout << "orth_ok " << orth.second << " " << orth.first.size() << "\n";
    for (const TCodeword &row : orth.first) cout << "G " << row << "\n";
    mt19937 rnd(239'239'239);
    vector<TCodeword> words = gen_random_codewords(orth.first, 8, rnd);
    for (const TCodeword &w : words) cout << "W " << w << " " << IsCodeword(H, w) << "\n";
    // transmit with the reference's per-frame seeding
    for (int f = 0; f < 3; ++f) {
        mt19937 noise(f + 1);
        TFVector y = transmit(-3.0, words[f], noise);
        cout << "Y";
        for (double v : y) {
            uint64_t bits;
            memcpy(&bits, &v, 8);
            cout << " " << hex << bits << dec;
        }
        cout << "\n";
    }
    // dense text round trip + parser quirks
    save_matrix(H, dir + "/_roundtrip.txt");
    cout << "roundtrip " << (read_pcm(dir + "/_roundtrip.txt") == H) << "\n";
    remove((dir + "/_roundtrip.txt").c_str());
    {
        ofstream q(dir + "/_quirks.txt");
        q << "1,0,2,1,\n0,1,1,0\n";
    }
    TMatrix Q = read_pcm(dir + "/_quirks.txt");
    remove((dir + "/_quirks.txt").c_str());
    cout << "quirks " << Q.size() << " " << Q[0] << " " << Q[1] << "\n";
    // rank-deficient H
    TMatrix bad = {H[0], H[1], H[0] ^ H[1]};
    cout << "deficient " << GetOrtogonal(bad).second << "\n";
    // products
    TCodeword s = H * words[0];
    cout << "syndrome_zero " << (s == TCodeword(H.size(), false)) << "\n";
    TCodeword u(orth.first.size(), false);
    u[0] = u[5] = true;
    cout << "vM " << ((u * orth.first) == (orth.first[0] ^ orth.first[5])) << "\n";
    // generic (non-GPU) experiment path: thread-count invariant counters
    auto dec = make_shared<HardDecision>();
    vector<TCodeword> many = gen_random_codewords(orth.first, 200, rnd);
    ExperimentResult a = multithread_experiment(dec, many, H, 4.0, 1);
    ExperimentResult b = multithread_experiment(dec, many, H, 4.0, 5);
    cout << "exp " << a.total << " " << a.correct << " " << a.pseudo << " " << a.tr.sum_hamming << " "
         << a.tr.sum_hamming_ok << " " << a.tr.sum_hamming_wrong << "\n";
    cout << "exp " << b.total << " " << b.correct << " " << b.pseudo << " " << b.tr.sum_hamming << " "
         << b.tr.sum_hamming_ok << " " << b.tr.sum_hamming_wrong << "\n";
    return 0;
}
