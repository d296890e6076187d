// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" shim around the UNMODIFIED reference headers, which are
// compiled from where they lie (-I$(REF), default /root/reference).  Nothing
// from the reference is copied into this repository: this file only *calls*
//   BeliefPropagationDecoder::decode   (algo/bp.h:208-222)
//   QPADMMDecoder::decode              (algo/qp_admm.h:180-194)
//   multithread_experiment             (experiment.h:125-139), threads_num = 1
//   read_pcm / GetOrtogonal / gen_random_codewords / transmit / llr / IsCodeword
// and marshals plain arrays in and out.  The output lives in oracle/_ref/
// (git-ignored, but shipped to the GPU box by gpurun).
//
// Rules that come from SURVEY.md section 0/8c:
//   * never compile with -DNDEBUG (experiment.h keeps pthread calls in assert());
//   * one thread per process only: Node::counter (bp.h:13) is a racy global;
//   * < 2.4 M BP frames per process (Node::counter is an int) -- callers reset
//     it through ref_reset_node_counter().
#include <cstdint>
#include <cstring>
#include <memory>
#include <iomanip>

#include "experiment.h"
#include "utils/parse_data.h"
#include "utils/codeword.h"
#include "algo/algo.h"
#include "algo/bp.h"
#include "algo/qp_admm.h"

namespace {

TMatrix to_tmatrix(const uint8_t *dense, int rows, int cols) {
    TMatrix M(rows, TCodeword(cols, false));
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) M[r][c] = dense[(size_t) r * cols + c] != 0;
    return M;
}

void from_tmatrix(const TMatrix &M, uint8_t *dense) {
    size_t k = 0;
    for (const auto &row : M)
        for (bool bit : row) dense[k++] = bit ? 1 : 0;
}

}  // namespace

extern "C" {

void ref_reset_node_counter() { Node::counter = 0; }

double ref_llr_variance(double snr) { return llr_variance(snr); }
double ref_llr(double v, double snr) { return llr(v, snr); }

// read_pcm (utils/parse_data.h:6-25).  Call once with dense == NULL to get the
// shape, then again with a buffer of rows*cols bytes.  Returns 0 on success.
int ref_read_pcm(const char *path, int *rows, int *cols, uint8_t *dense) {
    TMatrix M = read_pcm(path);
    if (M.empty()) return 1;
    *rows = (int) M.size();
    *cols = (int) M[0].size();
    if (dense) from_tmatrix(M, dense);
    return 0;
}

// save_matrix (utils/parse_data.h:44-54)
void ref_save_matrix(const uint8_t *dense, int rows, int cols, const char *path) {
    save_matrix(to_tmatrix(dense, rows, cols), path);
}

// GetOrtogonal (utils/codeword.h:97-128).  G is (n-m) x n.  Returns the bool.
int ref_get_orthogonal(const uint8_t *H, int m, int n, uint8_t *G) {
    auto res = GetOrtogonal(to_tmatrix(H, m, n));
    if (!res.second) return 0;
    from_tmatrix(res.first, G);
    return 1;
}

int ref_is_codeword(const uint8_t *H, int m, int n, const uint8_t *c) {
    TCodeword w(n);
    for (int i = 0; i < n; ++i) w[i] = c[i] != 0;
    return IsCodeword(to_tmatrix(H, m, n), w) ? 1 : 0;
}

// gen_random_codewords (utils/channel.h:29-44) with mt19937(seed), as
// main.cpp:63-64 / qpadmm_params.cpp:46-47 do.
void ref_gen_random_codewords(const uint8_t *G, int k, int n, int count, uint32_t seed, uint8_t *out) {
    mt19937 rnd(seed);
    auto words = gen_random_codewords(to_tmatrix(G, k, n), count, rnd);
    from_tmatrix(words, out);
}

// transmit (utils/channel.h:19-26) with a fresh mt19937(seed), the way
// experiment.h:97-99 calls it (seed = 1-based frame index at one thread).
void ref_transmit(double snr, const uint8_t *codeword, int n, uint32_t seed, double *y) {
    TCodeword c(n);
    for (int i = 0; i < n; ++i) c[i] = codeword[i] != 0;
    mt19937 rnd(seed);
    TFVector out = transmit(snr, c, rnd);
    memcpy(y, out.data(), sizeof(double) * n);
}

// One frame through BeliefPropagationDecoder.  bits gets n entries when the
// decoder succeeded; on failure the reference returns an EMPTY codeword
// (bp.h:198) and bits is left untouched.  Return value = the bool of the pair.
int ref_bp_decode(const uint8_t *H, int m, int n, const double *y, double snr, int max_iter, uint8_t *bits) {
    TMatrix Hm = to_tmatrix(H, m, n);
    TFVector yy(y, y + n);
    BeliefPropagationDecoder dec(max_iter);
    auto res = dec.decode(Hm, yy, snr);
    for (size_t i = 0; i < res.first.size(); ++i) bits[i] = res.first[i] ? 1 : 0;
    return res.second ? 1 : 0;
}

int ref_qpadmm_decode(const uint8_t *H, int m, int n, const double *y, double snr, double alpha, double mu,
                      int max_iter, double eps_stop, uint8_t *bits) {
    TMatrix Hm = to_tmatrix(H, m, n);
    TFVector yy(y, y + n);
    QPADMMDecoder dec(alpha, mu, max_iter, eps_stop);
    auto res = dec.decode(Hm, yy, snr);
    for (size_t i = 0; i < res.first.size(); ++i) bits[i] = res.first[i] ? 1 : 0;
    return res.second ? 1 : 0;
}

// Batched forms (B frames, y is B x n row-major).  Used for the golden
// fixtures and as the `--impl reference` CPU arm of bench.py.  ok[f] is the
// decoder's bool; bits rows of failed BP frames are zero-filled.
// Returns the seconds spent inside decode() only, timed as experiment.h:100-103.
double ref_bp_decode_batch(const uint8_t *H, int m, int n, const double *y, int frames, double snr, int max_iter,
                           uint8_t *bits, uint8_t *ok) {
    TMatrix Hm = to_tmatrix(H, m, n);
    BeliefPropagationDecoder dec(max_iter);
    double secs = 0;
    for (int f = 0; f < frames; ++f) {
        TFVector yy(y + (size_t) f * n, y + (size_t) (f + 1) * n);
        auto t0 = chrono::steady_clock::now();
        auto res = dec.decode(Hm, yy, snr);
        secs += (double) (chrono::steady_clock::now() - t0).count() / 1e9;
        ok[f] = res.second ? 1 : 0;
        uint8_t *row = bits + (size_t) f * n;
        memset(row, 0, n);
        for (size_t i = 0; i < res.first.size(); ++i) row[i] = res.first[i] ? 1 : 0;
        if (Node::counter > (1 << 30)) Node::counter = 0;
    }
    return secs;
}

double ref_qpadmm_decode_batch(const uint8_t *H, int m, int n, const double *y, int frames, double snr,
                               double alpha, double mu, int max_iter, double eps_stop, uint8_t *bits,
                               uint8_t *ok) {
    TMatrix Hm = to_tmatrix(H, m, n);
    QPADMMDecoder dec(alpha, mu, max_iter, eps_stop);
    double secs = 0;
    for (int f = 0; f < frames; ++f) {
        TFVector yy(y + (size_t) f * n, y + (size_t) (f + 1) * n);
        auto t0 = chrono::steady_clock::now();
        auto res = dec.decode(Hm, yy, snr);
        secs += (double) (chrono::steady_clock::now() - t0).count() / 1e9;
        ok[f] = res.second ? 1 : 0;
        uint8_t *row = bits + (size_t) f * n;
        memset(row, 0, n);
        for (size_t i = 0; i < res.first.size(); ++i) row[i] = res.first[i] ? 1 : 0;
    }
    return secs;
}

// The reference Monte-Carlo harness, run race-free (threads_num = 1).
// algo: 0 = BP(max_iter), 1 = QP-ADMM(alpha, mu, max_iter, eps_stop).
// out[0..6] = correct, pseudo, total, sum_hamming, sum_hamming_ok,
// sum_hamming_wrong, (int) round(time_sec * 1e6).
void ref_experiment(int algo, int max_iter, double alpha, double mu, double eps_stop, const uint8_t *H, int m,
                    int n, const uint8_t *codewords, int count, double snr, int64_t *out) {
    TMatrix Hm = to_tmatrix(H, m, n);
    vector<TCodeword> words = to_tmatrix(codewords, count, n);
    shared_ptr<Decoder> dec;
    if (algo == 0)
        dec = make_shared<BeliefPropagationDecoder>(max_iter);
    else
        dec = make_shared<QPADMMDecoder>(alpha, mu, max_iter, eps_stop);
    Node::counter = 0;
    ExperimentResult res = multithread_experiment(dec, words, Hm, snr, 1);
    out[0] = res.correct;
    out[1] = res.pseudo;
    out[2] = res.total;
    out[3] = res.tr.sum_hamming;
    out[4] = res.tr.sum_hamming_ok;
    out[5] = res.tr.sum_hamming_wrong;
    out[6] = (int64_t) (res.time_sec * 1e6 + 0.5);
}

}  // extern "C"
