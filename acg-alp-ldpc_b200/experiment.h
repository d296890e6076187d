// Monte-Carlo harness with the reference's public surface (experiment.h:25-139):
// HammingDistanceTracker, ExperimentResult, merge_exp_results and
// multithread_experiment(decoder, codewords, H, snr, threads_num, log_freq).
//
// For a GpuDecoder the whole point -- codeword selection, AWGN, decoding, verdict
// and counting -- runs on the GPUs: frames are sharded over the visible devices
// by GLOBAL frame index (any device count yields the same frames), each device
// produces one counter block, and the blocks are all-reduced over NVLink by NCCL
// (ldpc_experiment_run_multi; the reference's merge_exp_results after pthread_join).  Noise comes from the device Philox
// stream instead of mt19937(frame index + 1), so FER agrees with the reference
// statistically, not frame by frame; decoder parity is checked on identical y
// through Decoder::decode / the C ABI (tests/).
// Any other Decoder takes the generic path: a race-free thread pool calling
// decode() per frame exactly as exp() does (experiment.h:80-123).
#ifndef LDPC_B200_EXPERIMENT_H
#define LDPC_B200_EXPERIMENT_H

#include <atomic>
#include <chrono>
#include <memory>
#include <thread>

#include "algo/algo.h"
#include "algo/gpu_code.h"
#include "utils/codeword.h"

using namespace std;

// experiment.h:25-47.  Counters are 64-bit here: the reference's ints overflow
// beyond ~5e7 frames.
struct HammingDistanceTracker {
    HammingDistanceTracker(long long sum_hamming = 0, long long sum_hamming_ok = 0, long long sum_hamming_wrong = 0)
        : sum_hamming(sum_hamming), sum_hamming_ok(sum_hamming_ok), sum_hamming_wrong(sum_hamming_wrong) {}

    long long sum_hamming;
    long long sum_hamming_ok;
    long long sum_hamming_wrong;

    // distance between the transmitted word and the channel's hard decisions
    void new_experiment(const TMatrix &H, const TCodeword &c, const TFVector &y, bool correct) {
        long long flipped = 0;
        for (size_t i = 0; i < H[0].size(); ++i) flipped += c[i] ? (y[i] > 0) : (y[i] <= 0);
        sum_hamming += flipped;
        (correct ? sum_hamming_ok : sum_hamming_wrong) += flipped;
    }
};

// experiment.h:49-68
struct ExperimentResult {
    ExperimentResult(HammingDistanceTracker tr, long long correct = 0, long long pseudo = 0, long long total = 0,
                     double time_sec = 0)
        : tr(tr), correct(correct), pseudo(pseudo), total(total), time_sec(time_sec) {}

    HammingDistanceTracker tr;
    long long correct;
    long long pseudo;
    long long total;
    double time_sec;
    // extensions (not in the reference): post-decoding bit errors over frames that returned n bits, iterations
    long long bit_errors = 0, frames_with_bits = 0, sum_iters = 0;

    double FER() { return (double) (total - correct) / total; }

    double avg_time() { return time_sec / total; }

    double mean_hamming() { return (double) tr.sum_hamming / total; }

    double mean_hamming_ok() { return (double) tr.sum_hamming_ok / max(1LL, correct); }

    double mean_hamming_wrong() { return (double) tr.sum_hamming_wrong / max(1LL, total - correct); }

    double BER(size_t n) { return frames_with_bits ? (double) bit_errors / ((double) frames_with_bits * n) : 0.0; }
};

// experiment.h:70-78
inline void merge_exp_results(ExperimentResult &a, const ExperimentResult &b) {
    a.tr.sum_hamming += b.tr.sum_hamming;
    a.tr.sum_hamming_ok += b.tr.sum_hamming_ok;
    a.tr.sum_hamming_wrong += b.tr.sum_hamming_wrong;
    a.correct += b.correct;
    a.pseudo += b.pseudo;
    a.total += b.total;
    a.time_sec += b.time_sec;
    a.bit_errors += b.bit_errors;
    a.frames_with_bits += b.frames_with_bits;
    a.sum_iters += b.sum_iters;
}

namespace ldpc_host {

inline uint64_t experiment_seed() {
    if (const char *s = getenv("LDPC_SEED")) return strtoull(s, nullptr, 10);
    return 239239239ull;   // echoes main.cpp:63
}

inline ExperimentResult from_counters(const uint64_t *c, double seconds) {
    ExperimentResult r(HammingDistanceTracker((long long) c[LDPC_CNT_SUM_HAMMING], (long long) c[LDPC_CNT_SUM_HAMMING_OK],
                                              (long long) c[LDPC_CNT_SUM_HAMMING_WRONG]),
                       (long long) c[LDPC_CNT_CORRECT], (long long) c[LDPC_CNT_PSEUDO], (long long) c[LDPC_CNT_TOTAL],
                       seconds);
    r.bit_errors = (long long) c[LDPC_CNT_BIT_ERRORS];
    r.frames_with_bits = (long long) c[LDPC_CNT_FRAMES_WITH_BITS];
    r.sum_iters = (long long) c[LDPC_CNT_SUM_ITERS];
    return r;
}

// Pins the experiments of the calling host thread to one GPU (-1: frames are split over all visible GPUs).
// optimize_H.cpp evaluates several proposals concurrently, one host thread and one GPU per proposal.
inline int &pinned_gpu() {
    static thread_local int device = -1;
    return device;
}

// One Monte-Carlo point on all visible GPUs (or on the pinned one); frame f transmits codewords[f].
inline ExperimentResult gpu_experiment(const GpuDecoder &decoder, const vector<TCodeword> &codewords, const TMatrix &H,
                                       double snr) {
    const size_t frames = codewords.size(), n = H[0].size();
    const auto tp = chrono::steady_clock::now();
    vector<uint8_t> words(frames * n);
    for (size_t f = 0; f < frames; ++f) copy(codewords[f].begin(), codewords[f].end(), words.begin() + f * n);
    if (getenv("LDPC_EXP_TRACE"))
        cerr << "gpu_experiment: codeword table " << chrono::duration<double, milli>(chrono::steady_clock::now() - tp).count() << " ms" << endl;
    if (pinned_gpu() >= 0) {
        const auto t0 = chrono::steady_clock::now();
        CodeRef code = CodeCache::instance().get(H, pinned_gpu());
        const auto t1 = chrono::steady_clock::now();
        const ldpc_algo_cfg_t cfg1 = decoder.config();
        uint64_t cnt[LDPC_CNT_COUNT];
        double secs = 0;
        if (ldpc_experiment_run(code.get(), &cfg1, snr, experiment_seed(), 0, frames, LDPC_CW_TABLE, words.data(), frames, cnt,
                                &secs) != LDPC_OK)
            die("ldpc_experiment_run");
        if (getenv("LDPC_EXP_TRACE"))
            cerr << "gpu_experiment: code handle " << chrono::duration<double, milli>(t1 - t0).count() << " ms, run "
                 << chrono::duration<double, milli>(chrono::steady_clock::now() - t1).count() << " ms (kernel " << secs * 1e3
                 << " ms)" << endl;
        return from_counters(cnt, secs);
    }
    // several GPUs: shards by global frame index, counter blocks all-reduced over NVLink by NCCL inside the library
    // (the reference's merge_exp_results after pthread_join, experiment.h:70-78, 133-137)
    // (a GPU is worth a shard of a few thousand frames: the 1000-frame evaluations of optimize_H.cpp stay on one device)
    const int gpus = max(1, min<int>(visible_gpus(), (int) (frames / 4096)));
    const ldpc_algo_cfg_t cfg = decoder.config();
    vector<CodeRef> codes;
    vector<const ldpc_code_t *> handles;
    for (int g = 0; g < gpus; ++g) {
        codes.push_back(CodeCache::instance().get(H, g));
        handles.push_back(codes.back().get());
    }
    uint64_t cnt[LDPC_CNT_COUNT];
    double secs = 0;
    if (ldpc_experiment_run_multi(handles.data(), gpus, &cfg, snr, experiment_seed(), 0, frames, LDPC_CW_TABLE, words.data(),
                                  frames, cnt, &secs) != LDPC_OK)
        die("ldpc_experiment_run_multi");
    return from_counters(cnt, secs);
}

// The (alpha, mu) double loop of qpadmm_params.cpp:51-67 in one launch per GPU: the parameter pairs are
// dealt to the visible GPUs, each GPU evaluates its pairs on all frames (ldpc_qpadmm_grid_run).  Returns the
// result of every pair, in input order.
inline vector<ExperimentResult> gpu_qpadmm_grid(const vector<double> &alphas, const vector<double> &mus, int max_iter,
                                                double eps_stop, const vector<TCodeword> &codewords, const TMatrix &H,
                                                double snr) {
    const size_t points = alphas.size(), frames = codewords.size(), n = H[0].size();
    vector<uint8_t> words(frames * n);
    for (size_t f = 0; f < frames; ++f)
        for (size_t i = 0; i < n; ++i) words[f * n + i] = codewords[f][i];
    const int gpus = max(1, min<int>(visible_gpus(), (int) max<size_t>(points, 1)));
    const uint64_t seed = experiment_seed();
    vector<uint64_t> cnt(points * LDPC_CNT_COUNT, 0);
    vector<double> secs(gpus, 0.0);
    vector<thread> workers;
    for (int g = 0; g < gpus; ++g)
        workers.emplace_back([&, g] {
            const size_t begin = points * g / gpus, end = points * (g + 1) / gpus;
            if (begin == end) return;
            CodeRef code = CodeCache::instance().get(H, g);
            if (ldpc_qpadmm_grid_run(code.get(), (int32_t) (end - begin), alphas.data() + begin, mus.data() + begin, max_iter,
                                     eps_stop, snr, seed, 0, frames, LDPC_CW_TABLE, words.data(), frames,
                                     cnt.data() + begin * LDPC_CNT_COUNT, &secs[g]) != LDPC_OK)
                die("ldpc_qpadmm_grid_run");
        });
    for (thread &w : workers) w.join();
    vector<ExperimentResult> out;
    for (size_t i = 0; i < points; ++i) out.push_back(from_counters(cnt.data() + i * LDPC_CNT_COUNT, 0.0));
    return out;
}

// The reference's exp() loop for arbitrary decoders, without its races: the frame
// index comes from an atomic, and the noise generator is seeded with the 1-based
// index of the frame it belongs to (what experiment.h:90-97 yields at one thread).
inline ExperimentResult cpu_experiment(const shared_ptr<Decoder> &decoder, const vector<TCodeword> &codewords,
                                       const TMatrix &H, double snr, int threads_num, int log_freq) {
    atomic<size_t> next(0);
    vector<ExperimentResult> parts(max(threads_num, 1), ExperimentResult(HammingDistanceTracker()));
    vector<thread> workers;
    for (int t = 0; t < max(threads_num, 1); ++t)
        workers.emplace_back([&, t] {
            ExperimentResult &r = parts[t];
            for (;;) {
                const size_t f = next.fetch_add(1);
                if (f >= codewords.size()) break;
                const TCodeword &codeword = codewords[f];
                mt19937 rnd((uint32_t) (f + 1));
                TFVector y = transmit(snr, codeword, rnd);
                auto t0 = chrono::steady_clock::now();
                pair<TCodeword, bool> out = decoder->decode(H, y, snr);
                r.time_sec += chrono::duration<double>(chrono::steady_clock::now() - t0).count();
                bool correct = false;
                if (out.second && IsCodeword(H, out.first)) {
                    if (out.first == codeword) { r.correct++; correct = true; }
                    else r.pseudo++;
                }
                if (out.first.size() == codeword.size()) {
                    r.frames_with_bits++;
                    for (size_t i = 0; i < codeword.size(); ++i) r.bit_errors += out.first[i] != codeword[i];
                }
                r.total++;
                r.tr.new_experiment(H, codeword, y, correct);
                if (log_freq > 0 && (f + 1) % (size_t) log_freq == 0) cout << (f + 1) << endl;
            }
        });
    for (thread &w : workers) w.join();
    ExperimentResult total{HammingDistanceTracker()};
    for (const ExperimentResult &p : parts) merge_exp_results(total, p);
    return total;
}

}  // namespace ldpc_host

// experiment.h:125-139
inline ExperimentResult multithread_experiment(shared_ptr<Decoder> decoder, const vector<TCodeword> &codewords,
                                               const TMatrix &H, double snr, int threads_num, int log_freq = 1e9) {
    if (const GpuDecoder *gpu = dynamic_cast<const GpuDecoder *>(decoder.get())) {
        ExperimentResult res = ldpc_host::gpu_experiment(*gpu, codewords, H, snr);
        for (long long uuid = log_freq; log_freq > 0 && uuid <= (long long) codewords.size(); uuid += log_freq)
            cout << uuid << endl;   // the progress lines of experiment.h:104-108
        return res;
    }
    return ldpc_host::cpu_experiment(decoder, codewords, H, snr, threads_num, log_freq);
}

#endif
