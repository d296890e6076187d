// Test helper for the GPU box (compiled by tests/test_gpu_host_drivers.py): drives the C++ drop-in layer --
// BeliefPropagationDecoder / QPADMMDecoder behind the reference's Decoder interface (algo/algo.h:6-11) and
// multithread_experiment (experiment.h:125-139) -- the way the reference's harness does: ONE decoder instance
// shared by many pthreads, one frame per decode() call (experiment.h:101, 128-130; 200 threads in optimize_H.cpp:12).
//
//   gpu_host_check words  <data dir> <matrix> <seed> <count>          codewords of gen_random_codewords(G, count, mt19937(seed))
//   gpu_host_check decode <data dir> <matrix> <in.bin> <out.bin> <threads> <alpha> <mu> <admm iters>
//        in.bin : int32 tasks, then per task: double snr, n doubles y
//        out.bin: per task: BP ok byte, n BP bits (all 0 when the decoder returned an EMPTY codeword, and the
//                 byte after them = 1 iff it was empty), QP-ADMM ok byte, n QP-ADMM bits
#include <atomic>
#include <cstdio>
#include <cstring>
#include <pthread.h>

#include "experiment.h"
#include "utils/parse_data.h"
#include "algo/bp.h"
#include "algo/qp_admm.h"

struct Task {
    double snr;
    TFVector y;
    pair<TCodeword, bool> bp, admm;
};

struct Shared {
    TMatrix H;
    vector<Task> tasks;
    atomic<size_t> next{0};
    shared_ptr<Decoder> bp, admm;
};

static void *worker(void *arg) {
    Shared *s = static_cast<Shared *>(arg);
    for (;;) {
        const size_t i = s->next.fetch_add(1);
        if (i >= s->tasks.size()) break;
        Task &t = s->tasks[i];
        t.bp = s->bp->decode(s->H, t.y, t.snr);         // the same instance from every thread
        t.admm = s->admm->decode(s->H, t.y, t.snr);
    }
    return nullptr;
}

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const string mode = argv[1], dir = argv[2], matrix = argv[3];
    TMatrix H = load_matrix(dir + "/" + matrix);
    if (H.empty()) return 3;
    const size_t n = H[0].size();
    if (mode == "words" && argc >= 6) {
        TMatrix G = (matrix == "H05") ? load_matrix(dir + "/G05") : GetOrtogonal(H).first;
        mt19937 rnd((uint32_t) strtoul(argv[4], nullptr, 10));
        for (const TCodeword &w : gen_random_codewords(G, atoi(argv[5]), rnd)) cout << w << "\n";
        return 0;
    }
    if (mode == "decode" && argc >= 10) {
        Shared s;
        s.H = H;
        FILE *in = fopen(argv[4], "rb");
        if (!in) return 4;
        int32_t tasks = 0;
        if (fread(&tasks, 4, 1, in) != 1) return 4;
        s.tasks.resize(tasks);
        for (Task &t : s.tasks) {
            t.y.resize(n);
            if (fread(&t.snr, 8, 1, in) != 1 || fread(t.y.data(), 8, n, in) != n) return 4;
        }
        fclose(in);
        const int threads = atoi(argv[6]);
        s.bp = make_shared<BeliefPropagationDecoder>(100);
        s.admm = make_shared<QPADMMDecoder>(atof(argv[7]), atof(argv[8]), atoi(argv[9]), 1e-5);
        vector<pthread_t> ids(threads);
        for (pthread_t &id : ids) pthread_create(&id, nullptr, worker, &s);
        for (pthread_t &id : ids) pthread_join(id, nullptr);
        FILE *out = fopen(argv[5], "wb");
        if (!out) return 5;
        for (const Task &t : s.tasks) {
            vector<uint8_t> rec(2 * n + 3, 0);
            rec[0] = t.bp.second;
            for (size_t i = 0; i < t.bp.first.size() && i < n; ++i) rec[1 + i] = t.bp.first[i];
            rec[1 + n] = t.bp.first.empty();
            rec[2 + n] = t.admm.second;
            if (t.admm.first.size() != n) return 6;
            for (size_t i = 0; i < n; ++i) rec[3 + n + i] = t.admm.first[i];
            fwrite(rec.data(), 1, rec.size(), out);
        }
        fclose(out);
        cout << "decoded " << s.tasks.size() << " tasks on " << threads << " threads; names " << s.bp->name() << " "
             << s.admm->name() << "\n";
        return 0;
    }
    return 2;
}
