// Local-descent search over quasi-cyclic parity-check matrices (circulant blocks of
// size 20 on an 8 x 14 grid) that minimises the QP-ADMM frame error rate -- the
// reference's optimize_H.cpp (:12-137) with the FER evaluation on the GPU.
// The proposal chain is the reference's: mt19937(239) drives random_permute, a
// proposal is accepted iff its FER is strictly lower, every accepted matrix is
// written to the save path.  Every proposal is a new H, hence a new code handle.
//
// Environment: LDPC_OPT_ITERS (default 10000), LDPC_OPT_SAVE (default data/optimalH.txt), LDPC_OPT_WINDOW (proposals
//              evaluated speculatively at once, default = twice the number of GPUs; the trajectory does not depend on it),
//              LDPC_OPT_START (a matrix stem to start from instead of a random one), LDPC_OPT_PROCS (evaluation processes
//              instead of threads), LDPC_OPT_TRACE / LDPC_OPT_TIMELINE (where the time goes, stderr).
#include <memory>
#include <mutex>
#include <functional>
#include <condition_variable>
#include <thread>
#include <utility>

#include <cstring>
#include <sys/wait.h>
#include <unistd.h>

#include "experiment.h"
#include "utils/parse_data.h"
#include "utils/codeword.h"
#include "algo/algo.h"
#include "algo/qp_admm.h"

using namespace std;

const int THREADS_NUM = 200;
double SNR = -3.0;
shared_ptr<QPADMMDecoder> decoder = make_shared<QPADMMDecoder>(1.95, 0.5, 1000, 1e-5);

// LDPC_OPT_TRACE=1: where a proposal's time goes (microseconds summed over all evaluations, printed to stderr at exit)
static atomic<long long> g_us_orth(0), g_us_words(0), g_us_exp(0), g_us_kernel(0), g_evals(0);
static inline long long now_us() {
    return chrono::duration_cast<chrono::microseconds>(chrono::steady_clock::now().time_since_epoch()).count();
}

// FER of the decoder on `tests_num` codewords of H drawn from mt19937(239); 1.0 when H is rank deficient
double FER(const TMatrix &H, int tests_num = 1000) {
    const long long t0 = now_us();
    pair<TMatrix, bool> gen = GetOrtogonal(H);
    const long long t1 = now_us();
    g_us_orth += t1 - t0;
    ++g_evals;
    if (!gen.second) return 1.0;
    mt19937 rnd(239);
    vector<TCodeword> codewords = gen_random_codewords(gen.first, tests_num, rnd);
    const long long t2 = now_us();
    g_us_words += t2 - t1;
    ExperimentResult res = multithread_experiment(decoder, codewords, H, SNR, THREADS_NUM);
    const double fer = res.FER();
    g_us_exp += now_us() - t2;
    g_us_kernel += (long long) (res.time_sec * 1e6);
    if (getenv("LDPC_OPT_TIMELINE") && (g_evals < 24 || now_us() - t0 > 20000)) {
        static const long long t_origin = t0;
        cerr << "timeline: evaluation " << g_evals << " thread " << this_thread::get_id() << " from " << (t0 - t_origin) / 1000.0 << " ms to "
             << (now_us() - t_origin) / 1000.0 << " ms (host work until " << (t2 - t_origin) / 1000.0 << ")" << endl;
    }
    if (getenv("LDPC_EXP_TRACE")) cerr << "FER(): multithread_experiment " << (now_us() - t2) / 1000.0 << " ms" << endl;
    return fer;
}

// A block matrix whose (i, j) block is either zero or the identity cyclically shifted by diagonals[i][j].
struct PermutationsMatrix {
    PermutationsMatrix(int block_size, const TMatrix &blocks, const vector<vector<int>> &diagonals)
        : _block_size(block_size), _blocks(blocks), _diagonals(diagonals) {}

    // recover the block structure of a quasi-cyclic H (asserts that H is one)
    PermutationsMatrix(int block_size, const TMatrix &H) : _block_size(block_size) {
        assert(H.size() % block_size == 0 && H[0].size() % block_size == 0);
        const int rows = (int) H.size() / block_size, cols = (int) H[0].size() / block_size;
        _blocks.assign(rows, TCodeword(cols, false));
        _diagonals.assign(rows, vector<int>(cols, -block_size));
        for (int r = 0; r < (int) H.size(); ++r)
            for (int c = 0; c < (int) H[0].size(); ++c) {
                if (!H[r][c]) continue;
                const int bi = r / block_size, bj = c / block_size;
                const int shift = ((c % block_size) - (r % block_size) + block_size) % block_size;
                assert(!_blocks[bi][bj] || _diagonals[bi][bj] == shift);
                _blocks[bi][bj] = true;
                _diagonals[bi][bj] = shift;
            }
        assert(to_tmatrix() == H);
    }

    TMatrix to_tmatrix() const {
        const int rows = (int) _blocks.size(), cols = (int) _blocks[0].size();
        TMatrix H(_block_size * rows, TCodeword(_block_size * cols, false));
        for (int bi = 0; bi < rows; ++bi)
            for (int bj = 0; bj < cols; ++bj) {
                if (!_blocks[bi][bj]) continue;
                const int shift = _diagonals[bi][bj];
                assert(0 <= shift && shift < _block_size);
                for (int k = 0; k < _block_size; ++k) H[bi * _block_size + k][bj * _block_size + (shift + k) % _block_size] = true;
            }
        return H;
    }

    // one move of the search: pick a block; switch it on if it is off, otherwise toss a coin on switching it
    // off; then redraw its shift.  The generator is consumed in exactly this order (optimize_H.cpp:71-80).
    template <typename Gen>
    PermutationsMatrix random_permute(Gen &rnd) const {
        const int i = rnd() % (int) _blocks.size();
        const int j = rnd() % (int) _blocks[0].size();
        PermutationsMatrix next(*this);
        if (!next._blocks[i][j] or rnd() % 2 == 0) next._blocks[i][j] = !next._blocks[i][j];
        next._diagonals[i][j] = rnd() % _block_size;
        return next;
    }

private:
    int _block_size;
    TMatrix _blocks;
    vector<vector<int>> _diagonals;
};

// Persistent host threads for the concurrent evaluations: worker k is pinned to GPU k % gpus and keeps its CUDA
// stream and buffers (the library's workspaces are per host thread) for the whole search.
class ProposalPool {
public:
    ProposalPool(int workers, int gpus) {
        for (int k = 0; k < workers; ++k)
            threads_.emplace_back([this, k, gpus] {
                ldpc_host::pinned_gpu() = k % gpus;
                int seen = 0;
                for (;;) {
                    unique_lock<mutex> lock(mu_);
                    cv_.wait(lock, [&] { return stop_ || generation_ != seen; });
                    if (stop_) return;
                    seen = generation_;
                    const bool mine = k < count_;
                    lock.unlock();
                    if (mine) job_(k);
                    lock.lock();
                    if (mine && --pending_ == 0) done_.notify_all();
                }
            });
    }
    ~ProposalPool() {
        { lock_guard<mutex> lock(mu_); stop_ = true; }
        cv_.notify_all();
        for (thread &t : threads_) t.join();
    }
    // job(k) for k in [0, count) on worker k; returns when all are done
    void run(int count, function<void(int)> job) {
        unique_lock<mutex> lock(mu_);
        job_ = job;
        count_ = pending_ = count;
        ++generation_;
        cv_.notify_all();
        done_.wait(lock, [&] { return pending_ == 0; });
    }

private:
    vector<thread> threads_;
    mutex mu_;
    condition_variable cv_, done_;
    function<void(int)> job_;
    int generation_ = 0, count_ = 0, pending_ = 0;
    bool stop_ = false;
};

// CUDA start-up grows with the number of visible GPUs (14 s for the first call on an 8 x B200 box against 1.5-4.5 s with one
// GPU visible), so a process that will use `count` GPUs starting at position `first` of the visible ones narrows
// CUDA_VISIBLE_DEVICES to them BEFORE it touches CUDA.  (This is the application, not the library: the library never
// changes the environment.)
static void narrow_visible_gpus(int first, int count) {
    vector<string> ids;
    if (const char *cur = getenv("CUDA_VISIBLE_DEVICES")) {
        string item;
        for (const char *c = cur;; ++c) {
            if (*c == ',' || *c == 0) { if (!item.empty()) ids.push_back(item); item.clear(); if (*c == 0) break; }
            else item.push_back(*c);
        }
    }
    string value;
    for (int k = 0; k < count; ++k) {
        const int pos = first + k;
        const string id = pos < (int) ids.size() ? ids[pos] : (ids.empty() ? to_string(pos) : string());
        if (id.empty()) break;
        value += (value.empty() ? "" : ",") + id;
    }
    if (!value.empty()) setenv("CUDA_VISIBLE_DEVICES", value.c_str(), 1);
}

// Evaluation PROCESSES (several GPUs): worker k is a child process pinned to GPU k % gpus that reads a matrix from a pipe,
// evaluates its FER and writes the number back.  Threads of one process share the CUDA runtime's process-wide locks, and
// with a new code handle per proposal the evaluations queue for them (8 GPUs, 16 threads: 2.1 x one GPU); processes do not.
// The children are forked BEFORE anything touches CUDA in this process (a CUDA context does not survive fork).
class ProposalProcs {
public:
    ProposalProcs(int workers, int gpus) {
        for (int k = 0; k < workers; ++k) {
            int to_child[2], from_child[2];
            if (pipe(to_child) != 0 || pipe(from_child) != 0) { perror("pipe"); exit(1); }
            const pid_t pid = fork();
            if (pid < 0) { perror("fork"); exit(1); }
            if (pid == 0) {
                for (const Worker &w : workers_) { close(w.wr); close(w.rd); }      // the siblings' ends
                close(to_child[1]);
                close(from_child[0]);
                narrow_visible_gpus(k % gpus, 1);              // this process sees ONE GPU, as device 0
                ldpc_host::pinned_gpu() = 0;
                serve(to_child[0], from_child[1]);
                _exit(0);
            }
            close(to_child[0]);
            close(from_child[1]);
            workers_.push_back(Worker{pid, to_child[1], from_child[0]});
        }
    }
    ~ProposalProcs() {
        for (const Worker &w : workers_) { close(w.wr); close(w.rd); }
        for (const Worker &w : workers_) waitpid(w.pid, nullptr, 0);
    }
    int size() const { return (int) workers_.size(); }
    // send matrix H to worker k (FER on tests_num codewords); the answer is read with collect(k)
    void submit(int k, const TMatrix &H, int tests_num) {
        vector<unsigned char> msg(12 + H.size() * H[0].size());
        const int32_t hdr[3] = {(int32_t) H.size(), (int32_t) H[0].size(), tests_num};
        memcpy(msg.data(), hdr, 12);
        size_t at = 12;
        for (const TCodeword &row : H)
            for (bool bit : row) msg[at++] = bit;
        write_all(workers_[k].wr, msg.data(), msg.size());
    }
    double collect(int k) {
        double fer = 1.0;
        read_all(workers_[k].rd, &fer, sizeof(fer));
        return fer;
    }

private:
    struct Worker { pid_t pid; int wr, rd; };
    vector<Worker> workers_;
    static void write_all(int fd, const void *buf, size_t n) {
        const char *p = static_cast<const char *>(buf);
        while (n) { const ssize_t k = write(fd, p, n); if (k <= 0) { perror("write"); exit(1); } p += k; n -= (size_t) k; }
    }
    static bool read_all(int fd, void *buf, size_t n) {
        char *p = static_cast<char *>(buf);
        while (n) { const ssize_t k = read(fd, p, n); if (k <= 0) return false; p += k; n -= (size_t) k; }
        return true;
    }
    static void serve(int rd, int wr) {
        for (;;) {
            int32_t hdr[3];
            if (!read_all(rd, hdr, 12)) return;                   // the parent closed the pipe: done
            vector<unsigned char> bits((size_t) hdr[0] * hdr[1]);
            if (!read_all(rd, bits.data(), bits.size())) return;
            TMatrix H(hdr[0], TCodeword(hdr[1]));
            for (int r = 0; r < hdr[0]; ++r)
                for (int c = 0; c < hdr[1]; ++c) H[r][c] = bits[(size_t) r * hdr[1] + c] != 0;
            const double fer = FER(H, hdr[2]);
            write_all(wr, &fer, sizeof(fer));
        }
    }
};

static ProposalProcs *g_procs = nullptr;      // non-null: evaluations go to the worker processes
static long long g_t_after_initial = 0;

// FER through worker 0 when the evaluation processes exist (this process then never touches CUDA), else here
static double FER_anywhere(const TMatrix &H, int tests_num = 1000) {
    if (!g_procs) return FER(H, tests_num);
    g_procs->submit(0, H, tests_num);
    return g_procs->collect(0);
}

// The reference's chain (optimize_H.cpp:89-104), evaluated speculatively: the next `window` proposals are all drawn
// from the CURRENT matrix with the generator states the sequential loop would have, their FERs are evaluated
// concurrently (one host thread each, dealt round-robin to the visible GPUs), and the first improving one is
// accepted -- the later ones are discarded and the generator is rewound to the state right after the accepted draw.
// The accepted sequence, the printed lines and the saved matrices are those of the sequential loop (window = 1).
template <typename Gen>
PermutationsMatrix optimize(PermutationsMatrix H, Gen &rnd, int iters, const string &save_filepath, int window) {
    // (with evaluation processes every worker evaluates the start matrix once: CUDA start-up -- context, module load,
    // seconds per process -- is paid by all of them at the same time instead of inside the first windows)
    if (g_procs) {
        for (int k = 1; k < g_procs->size(); ++k) g_procs->submit(k, H.to_tmatrix(), 1000);
    }
    double error = FER_anywhere(H.to_tmatrix());
    if (g_procs)
        for (int k = 1; k < g_procs->size(); ++k) g_procs->collect(k);
    cout << "initial FER=" << error << endl;
    g_t_after_initial = now_us();
    ProposalPool pool(!g_procs && window > 1 ? window : 0, g_procs ? 1 : ldpc_host::visible_gpus());
    for (int i = 0; i < iters;) {
        const int w = max(1, min(window, iters - i));
        vector<PermutationsMatrix> candidates;
        vector<Gen> state_after;
        Gen draw = rnd;
        for (int k = 0; k < w; ++k) {
            candidates.push_back(H.random_permute(draw));
            state_after.push_back(draw);
        }
        vector<double> errors(w, 1.0);
        if (g_procs) {
            for (int k = 0; k < w; ++k) g_procs->submit(k, candidates[k].to_tmatrix(), 1000);
            for (int k = 0; k < w; ++k) errors[k] = g_procs->collect(k);
        } else if (w == 1) {
            errors[0] = FER(candidates[0].to_tmatrix());
        } else {
            pool.run(w, [&](int k) { errors[k] = FER(candidates[k].to_tmatrix()); });
        }
        int used = w;
        for (int k = 0; k < w; ++k) {
            cout << "\tproposal: FER=" << errors[k] << endl;
            if (errors[k] < error) {
                H = candidates[k];
                error = errors[k];
                cout << "accept, FER=" << error << endl;
                save_matrix(H.to_tmatrix(), save_filepath);
                used = k + 1;
                break;
            }
        }
        rnd = state_after[used - 1];
        i += used;
    }
    return H;
}

// random start with unseeded rand(), redrawn until H has full row rank (optimize_H.cpp:106-122)
PermutationsMatrix random_permutation_matrix(int block_size, int n, int m) {
    for (;;) {
        TMatrix blocks(n, TCodeword(m));
        vector<vector<int>> shifts(n, vector<int>(m));
        for (int i = 0; i < n; i++)
            for (int j = 0; j < m; j++) {
                blocks[i][j] = rand() % 2;
                shifts[i][j] = rand() % block_size;
            }
        PermutationsMatrix H(block_size, blocks, shifts);
        if (GetOrtogonal(H.to_tmatrix()).second) return H;
    }
}

int main() {
    std::ios::sync_with_stdio(0);
    cout.precision(5);
    cout << fixed;

    const int iters = getenv("LDPC_OPT_ITERS") ? atoi(getenv("LDPC_OPT_ITERS")) : 10000;
    const string save = getenv("LDPC_OPT_SAVE") ? getenv("LDPC_OPT_SAVE") : "data/optimalH.txt";
    PermutationsMatrix H0 = getenv("LDPC_OPT_START") ? PermutationsMatrix(20, load_matrix(getenv("LDPC_OPT_START")))
                                                     : random_permutation_matrix(20, 8, 14);
    const long long t_start = now_us();
    mt19937 rnd(239);
    // LDPC_OPT_PROCS=<n>: n evaluation processes instead of evaluation threads in this process.
    // The GPU count is asked for in a short-lived child, so that this process has not touched CUDA when it forks.
    int gpus = 1;
    {
        int fd[2];
        if (pipe(fd) == 0) {
            const pid_t pid = fork();
            if (pid == 0) {
                const int n = ldpc_host::visible_gpus();
                if (write(fd[1], &n, sizeof(n)) != (ssize_t) sizeof(n)) _exit(1);
                _exit(0);
            }
            close(fd[1]);
            if (pid > 0) {
                if (read(fd[0], &gpus, sizeof(gpus)) != (ssize_t) sizeof(gpus)) gpus = 1;
                waitpid(pid, nullptr, 0);
            }
            close(fd[0]);
        }
    }
    // Measured on 8 x B200, 3000 proposals (profiles/r02_optimize_H_8gpu.txt): evaluation threads 0.97 ms per proposal after
    // 13 s of CUDA start-up, 16 evaluation processes 0.69 ms after 21 s (one GPU: 4.33 ms after 8.6 s) -- the processes
    // pay off beyond ~28000 proposals, so threads are the default and LDPC_OPT_PROCS=<n> asks for processes.
    const int procs = getenv("LDPC_OPT_PROCS") ? max(0, atoi(getenv("LDPC_OPT_PROCS"))) : 0;
    // proposals evaluated concurrently (1 = the reference's sequential loop; the trajectory is the same for any value);
    // default: two in flight per GPU, so that the host work of one (GetOrtogonal, 1000 codewords, compiling and
    // uploading the new H) overlaps the evaluation of the other
    int window = getenv("LDPC_OPT_WINDOW") ? max(1, atoi(getenv("LDPC_OPT_WINDOW"))) : 2 * gpus;
    ProposalProcs *procs_owner = nullptr;
    if (procs > 0) {
        g_procs = procs_owner = new ProposalProcs(procs, gpus);
        window = min(window, procs);
    } else {
        narrow_visible_gpus(0, gpus);                  // evaluation threads in this process: the GPUs it will use
    }
    TMatrix H = optimize(H0, rnd, iters, save, window).to_tmatrix();
    const long long t_search = now_us();

    cout << FER_anywhere(H, 10000) << endl;
    if (getenv("LDPC_OPT_TRACE"))
        cerr << "trace: initial FER (CUDA start-up included) " << (g_t_after_initial - t_start) / 1000 << " ms; " << iters
             << " proposals " << (t_search - g_t_after_initial) / 1000 << " ms = "
             << (double) (t_search - g_t_after_initial) / 1000.0 / iters << " ms per proposal; final FER on 10000 frames "
             << (now_us() - t_search) / 1000 << " ms; evaluation processes " << procs << ", window " << window << endl;
    delete procs_owner;
    if (getenv("LDPC_OPT_TRACE"))
        cerr << "trace: " << g_evals << " evaluations; per evaluation: GetOrtogonal " << g_us_orth / max(1LL, (long long) g_evals)
             << " us, codewords " << g_us_words / max(1LL, (long long) g_evals) << " us, experiment (code handle + upload + "
             << "kernel) " << g_us_exp / max(1LL, (long long) g_evals) << " us, of which the kernel (CUDA events) "
             << g_us_kernel / max(1LL, (long long) g_evals) << " us; window " << window << endl;
    return 0;
}
