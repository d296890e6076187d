// extern "C" entry points of libldpc_b200.so (see include/ldpc_b200.h).
//
// Host-pointer entry points stream the batch through two slots (copy-in of chunk
// i+1 overlaps the decode of chunk i and the copy-out of chunk i-1); device-pointer
// entry points just enqueue the kernel.  All workspaces are per host thread, so a
// code handle can be shared by any number of threads.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <cstdio>
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>

#include "ldpc_internal.h"

using namespace ldpc;

namespace {

struct Slot {
    cudaStream_t stream = nullptr;
    unsigned long long *queue = nullptr;   // frame queue head
    unsigned long long *counters = nullptr;
    double *y = nullptr, *soft = nullptr;
    uint8_t *bits = nullptr, *ok = nullptr;
    int32_t *iters = nullptr;
    size_t cap_frames = 0, cap_n = 0;
    bool has_soft = false;
    cudaEvent_t e0 = nullptr, e1 = nullptr;   // timing events of the experiment entry points (created once)
};

struct ThreadCtx {
    std::map<int, Slot[2]> per_device;
    // host threads come and go (optimize_H.cpp evaluates proposals on short-lived threads): give the streams and
    // buffers back when the thread ends (errors are ignored: at process exit the runtime may already be gone)
    ~ThreadCtx() {
        for (auto &kv : per_device) {
            if (cudaSetDevice(kv.first) != cudaSuccess) continue;
            for (Slot &s : kv.second) {
                if (s.stream) cudaStreamSynchronize(s.stream);
                cudaFree(s.queue); cudaFree(s.counters); cudaFree(s.y); cudaFree(s.soft);
                cudaFree(s.bits); cudaFree(s.ok); cudaFree(s.iters);
                if (s.e0) cudaEventDestroy(s.e0);
                if (s.e1) cudaEventDestroy(s.e1);
                if (s.stream) cudaStreamDestroy(s.stream);
            }
        }
        cudaGetLastError();
    }
};

thread_local ThreadCtx g_ctx;

int slot_init(Slot &s) {
    if (s.stream) return LDPC_OK;
    LDPC_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    LDPC_CUDA(cudaMalloc((void **) &s.queue, sizeof(unsigned long long)));
    LDPC_CUDA(cudaMalloc((void **) &s.counters, sizeof(unsigned long long) * LDPC_CNT_COUNT));
    return LDPC_OK;
}

int slot_reserve(Slot &s, size_t frames, size_t n, bool soft) {
    if (frames <= s.cap_frames && n <= s.cap_n && (!soft || s.has_soft)) return LDPC_OK;
    LDPC_CUDA(cudaStreamSynchronize(s.stream));
    cudaFree(s.y); cudaFree(s.soft); cudaFree(s.bits); cudaFree(s.ok); cudaFree(s.iters);
    s.y = s.soft = nullptr; s.bits = s.ok = nullptr; s.iters = nullptr;
    s.cap_frames = std::max(frames, s.cap_frames);
    s.cap_n = std::max(n, s.cap_n);
    s.has_soft = soft || s.has_soft;
    LDPC_CUDA(cudaMalloc((void **) &s.y, sizeof(double) * s.cap_frames * s.cap_n));
    if (s.has_soft) LDPC_CUDA(cudaMalloc((void **) &s.soft, sizeof(double) * s.cap_frames * s.cap_n));
    LDPC_CUDA(cudaMalloc((void **) &s.bits, s.cap_frames * s.cap_n));
    LDPC_CUDA(cudaMalloc((void **) &s.ok, s.cap_frames));
    LDPC_CUDA(cudaMalloc((void **) &s.iters, sizeof(int32_t) * s.cap_frames));
    return LDPC_OK;
}

struct DecodeCfg {
    int algo;
    double snr, alpha, mu, eps_stop;
    int max_iter, early_exit;
};

int enqueue_decode(const ldpc_code *c, const DecodeCfg &cfg, const FrameIO &io, int64_t frames,
                   unsigned long long *queue, cudaStream_t stream) {
    const double var = llr_variance(cfg.snr);
    LDPC_CUDA(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), stream));
    if (cfg.algo == LDPC_ALGO_BP) return launch_bp(c, io, frames, var, cfg.max_iter, cfg.early_exit, queue, stream);
    return launch_qpadmm(c, io, frames, var, cfg.alpha, cfg.mu, cfg.max_iter, cfg.eps_stop, queue, stream);
}

int check_common(const ldpc_code *c, const void *y, int64_t frames, const void *bits, const void *ok,
                 const void *iters, int max_iter) {
    if (!c) return fail(LDPC_E_INVALID, "code is NULL");
    if (frames < 0) return fail(LDPC_E_INVALID, "frames < 0");
    if (frames > 0 && (!y || !bits || !ok || !iters)) return fail(LDPC_E_INVALID, "NULL buffer");
    if (max_iter < 0) return fail(LDPC_E_INVALID, "max_iter < 0");
    return LDPC_OK;
}

int decode_device(const ldpc_code *c, const DecodeCfg &cfg, const double *d_y, int64_t frames, uint8_t *d_bits,
                  uint8_t *d_ok, int32_t *d_iters, double *d_soft, cudaStream_t stream) {
    LDPC_CUDA(cudaSetDevice(c->device));
    if (frames == 0) return LDPC_OK;
    int st;
    FrameIO io;
    io.y = d_y; io.bits = d_bits; io.ok = d_ok; io.iters = d_iters; io.soft = d_soft;
    // the frame-queue head is stream-ordered scratch of this launch
    unsigned long long *queue = nullptr;
    LDPC_CUDA(cudaMallocAsync((void **) &queue, sizeof(unsigned long long), stream));
    st = enqueue_decode(c, cfg, io, frames, queue, stream);
    cudaFreeAsync(queue, stream);
    return st;
}

int decode_host(const ldpc_code *c, const DecodeCfg &cfg, const double *y, int64_t frames, uint8_t *bits,
                uint8_t *ok, int32_t *iters, double *soft) {
    LDPC_CUDA(cudaSetDevice(c->device));
    if (frames == 0) return LDPC_OK;
    const size_t n = c->n;
    Slot *slots = g_ctx.per_device[c->device];
    // chunk: <= 256 MiB of y per slot, and at least a few waves of CTAs
    int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(frames, (256ll << 20) / (int64_t) (n * sizeof(double))));
    if (frames > chunk && frames < 2 * chunk) chunk = (frames + 1) / 2;
    if ((int64_t) (frames * n * sizeof(double)) >= (64ll << 20))
        chunk = std::min<int64_t>(chunk, (frames + 3) / 4);   // large batches: at least four chunks, so copies overlap kernels
    int st;
    for (int s = 0; s < 2; ++s) {
        if ((st = slot_init(slots[s]))) return st;
        if ((st = slot_reserve(slots[s], (size_t) chunk, n, soft != nullptr))) return st;
    }
    // the first chunk is small: its copy-in is the only one no kernel hides
    // On a failure in the middle of the loop the work already queued still reads the caller's y and writes the caller's
    // output buffers: both streams are drained before the status is returned (secondary errors are ignored).
    auto pipeline = [&]() -> int {
        int which = 0;
        int64_t cnt = 0;
        for (int64_t begin = 0; begin < frames; begin += cnt, which ^= 1) {
            Slot &s = slots[which];
            cnt = std::min(begin == 0 && frames > chunk ? std::max<int64_t>(chunk / 8, 1) : chunk, frames - begin);
            LDPC_CUDA(cudaMemcpyAsync(s.y, y + begin * n, sizeof(double) * cnt * n, cudaMemcpyHostToDevice, s.stream));
            FrameIO io;
            io.y = s.y; io.bits = s.bits; io.ok = s.ok; io.iters = s.iters; io.soft = soft ? s.soft : nullptr;
            int st2;
            if ((st2 = enqueue_decode(c, cfg, io, cnt, s.queue, s.stream))) return st2;
            LDPC_CUDA(cudaMemcpyAsync(bits + begin * n, s.bits, cnt * n, cudaMemcpyDeviceToHost, s.stream));
            LDPC_CUDA(cudaMemcpyAsync(ok + begin, s.ok, cnt, cudaMemcpyDeviceToHost, s.stream));
            LDPC_CUDA(cudaMemcpyAsync(iters + begin, s.iters, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, s.stream));
            if (soft)
                LDPC_CUDA(cudaMemcpyAsync(soft + begin * n, s.soft, sizeof(double) * cnt * n, cudaMemcpyDeviceToHost,
                                          s.stream));
        }
        return LDPC_OK;
    };
    st = pipeline();
    const cudaError_t e0 = cudaStreamSynchronize(slots[0].stream), e1 = cudaStreamSynchronize(slots[1].stream);
    if (st) return st;
    if (e0 != cudaSuccess) return cuda_fail(e0, "decode stream 0", __FILE__, __LINE__);
    if (e1 != cudaSuccess) return cuda_fail(e1, "decode stream 1", __FILE__, __LINE__);
    return LDPC_OK;
}

// ---- one Monte-Carlo shard: checks, launch (asynchronous, counters stay on the device), collection

int experiment_check(const ldpc_code *c, const ldpc_algo_cfg_t *cfg, int codeword_source, const uint8_t *words, uint64_t n_words) {
    if (!c || !cfg) return fail(LDPC_E_INVALID, "NULL argument");
    if (cfg->algo != LDPC_ALGO_BP && cfg->algo != LDPC_ALGO_QPADMM) return fail(LDPC_E_INVALID, "unknown algo");
    if (cfg->max_iter < 0) return fail(LDPC_E_INVALID, "max_iter < 0");
    if (codeword_source < LDPC_CW_ZERO || codeword_source > LDPC_CW_GENERATOR)
        return fail(LDPC_E_INVALID, "unknown codeword source");
    if (codeword_source == LDPC_CW_TABLE && (!words || n_words == 0))
        return fail(LDPC_E_INVALID, "LDPC_CW_TABLE needs a codeword table");
    if (codeword_source == LDPC_CW_GENERATOR && (c->k <= 0 || !c->d.gen_cols))
        return fail(LDPC_E_INVALID, "LDPC_CW_GENERATOR needs ldpc_code_set_generator");
    return LDPC_OK;
}

struct ExperimentJob {
    int device = 0;
    cudaStream_t stream = nullptr;
    unsigned long long *counters = nullptr;      // the calling thread's per-device counter block (device memory)
    uint8_t *d_words = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

int experiment_enqueue(const ldpc_code *c, const ldpc_algo_cfg_t *cfg, double snr, uint64_t seed, uint64_t frame_begin,
                       uint64_t frame_count, int codeword_source, const uint8_t *words, uint64_t n_words, ExperimentJob *job) {
    LDPC_CUDA(cudaSetDevice(c->device));
    DecodeCfg dc{cfg->algo, snr, cfg->alpha, cfg->mu, cfg->eps_stop, cfg->max_iter, cfg->early_exit};
    Slot &s = g_ctx.per_device[c->device][0];
    int st = slot_init(s);
    if (st) return st;
    job->device = c->device;
    job->stream = s.stream;
    job->counters = s.counters;
    if (codeword_source == LDPC_CW_TABLE) {
        LDPC_CUDA(dev_malloc((void **) &job->d_words, n_words * (size_t) c->n));
        LDPC_CUDA(cudaMemcpyAsync(job->d_words, words, n_words * (size_t) c->n, cudaMemcpyHostToDevice, s.stream));
    }
    if (!s.e0) LDPC_CUDA(cudaEventCreate(&s.e0));
    if (!s.e1) LDPC_CUDA(cudaEventCreate(&s.e1));
    job->e0 = s.e0;
    job->e1 = s.e1;
    LDPC_CUDA(cudaMemsetAsync(s.counters, 0, sizeof(unsigned long long) * LDPC_CNT_COUNT, s.stream));
    FrameIO io;
    io.experiment = 1; io.seed = seed; io.frame_begin = frame_begin; io.cw_source = codeword_source;
    io.words = job->d_words; io.n_words = n_words; io.counters = s.counters;
    LDPC_CUDA(cudaEventRecord(job->e0, s.stream));
    if (frame_count > 0) st = enqueue_decode(c, dc, io, (int64_t) frame_count, s.queue, s.stream);
    LDPC_CUDA(cudaEventRecord(job->e1, s.stream));
    return st;
}

// waits for the shard's stream and reads its counter block (also after a failed enqueue: the stream is drained)
int experiment_collect(ExperimentJob *job, uint64_t counters[LDPC_CNT_COUNT], double *gpu_seconds) {
    LDPC_CUDA(cudaSetDevice(job->device));
    unsigned long long host_cnt[LDPC_CNT_COUNT];
    LDPC_CUDA(cudaMemcpyAsync(host_cnt, job->counters, sizeof(host_cnt), cudaMemcpyDeviceToHost, job->stream));
    LDPC_CUDA(cudaStreamSynchronize(job->stream));
    for (int i = 0; i < LDPC_CNT_COUNT; ++i) counters[i] = host_cnt[i];
    float ms = 0;
    if (job->e0 && job->e1 && cudaEventElapsedTime(&ms, job->e0, job->e1) == cudaSuccess && gpu_seconds) *gpu_seconds = ms * 1e-3;
    cudaGetLastError();
    return LDPC_OK;
}

void experiment_release(ExperimentJob *job) {
    if (job->stream) { cudaSetDevice(job->device); cudaStreamSynchronize(job->stream); }
    dev_free(job->d_words);                   // (the events belong to the thread's slot)
    cudaGetLastError();
    *job = ExperimentJob();
}

}  // namespace

extern "C" {

int ldpc_bp_decode(const ldpc_code_t *c, const double *y, int64_t frames, double snr, int32_t max_iter,
                   int32_t early_exit, uint8_t *bits, uint8_t *ok, int32_t *iters, double *post_llr) {
    int st = check_common(c, y, frames, bits, ok, iters, max_iter);
    if (st) return st;
    DecodeCfg cfg{LDPC_ALGO_BP, snr, 0, 0, 0, max_iter, early_exit};
    return decode_host(c, cfg, y, frames, bits, ok, iters, post_llr);
}

int ldpc_bp_decode_device(const ldpc_code_t *c, const double *d_y, int64_t frames, double snr, int32_t max_iter,
                          int32_t early_exit, uint8_t *d_bits, uint8_t *d_ok, int32_t *d_iters,
                          double *d_post_llr, void *stream) {
    int st = check_common(c, d_y, frames, d_bits, d_ok, d_iters, max_iter);
    if (st) return st;
    DecodeCfg cfg{LDPC_ALGO_BP, snr, 0, 0, 0, max_iter, early_exit};
    return decode_device(c, cfg, d_y, frames, d_bits, d_ok, d_iters, d_post_llr, (cudaStream_t) stream);
}

int ldpc_qpadmm_decode(const ldpc_code_t *c, const double *y, int64_t frames, double snr, double alpha, double mu,
                       int32_t max_iter, double eps_stop, uint8_t *bits, uint8_t *ok, int32_t *iters,
                       double *v_out) {
    int st = check_common(c, y, frames, bits, ok, iters, max_iter);
    if (st) return st;
    DecodeCfg cfg{LDPC_ALGO_QPADMM, snr, alpha, mu, eps_stop, max_iter, 1};
    return decode_host(c, cfg, y, frames, bits, ok, iters, v_out);
}

int ldpc_qpadmm_decode_device(const ldpc_code_t *c, const double *d_y, int64_t frames, double snr, double alpha,
                              double mu, int32_t max_iter, double eps_stop, uint8_t *d_bits, uint8_t *d_ok,
                              int32_t *d_iters, double *d_v_out, void *stream) {
    int st = check_common(c, d_y, frames, d_bits, d_ok, d_iters, max_iter);
    if (st) return st;
    DecodeCfg cfg{LDPC_ALGO_QPADMM, snr, alpha, mu, eps_stop, max_iter, 1};
    return decode_device(c, cfg, d_y, frames, d_bits, d_ok, d_iters, d_v_out, (cudaStream_t) stream);
}

int ldpc_channel_generate(const ldpc_code_t *c, uint64_t seed, uint64_t frame_begin, int64_t frames, double snr,
                          const uint8_t *codewords, double *y) {
    if (!c || frames < 0 || (frames > 0 && !y)) return fail(LDPC_E_INVALID, "bad argument");
    if (frames == 0) return LDPC_OK;
    LDPC_CUDA(cudaSetDevice(c->device));
    const size_t n = c->n;
    double *d_y = nullptr;
    uint8_t *d_cw = nullptr;
    LDPC_CUDA(cudaMalloc((void **) &d_y, sizeof(double) * frames * n));
    int st = LDPC_OK;
    if (codewords) {
        cudaError_t e = cudaMalloc((void **) &d_cw, frames * n);
        if (e == cudaSuccess) e = cudaMemcpy(d_cw, codewords, frames * n, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) st = cuda_fail(e, "codeword upload", __FILE__, __LINE__);
    }
    if (!st) st = launch_channel(c, seed, frame_begin, frames, std::sqrt(llr_variance(snr)), d_cw, d_y, 0);
    if (!st) {
        cudaError_t e = cudaMemcpy(y, d_y, sizeof(double) * frames * n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) st = cuda_fail(e, "y download", __FILE__, __LINE__);
    }
    cudaFree(d_y);
    cudaFree(d_cw);
    return st;
}

int ldpc_channel_generate_device(const ldpc_code_t *c, uint64_t seed, uint64_t frame_begin, int64_t frames,
                                 double snr, const uint8_t *d_codewords, double *d_y, void *stream) {
    if (!c || frames < 0 || (frames > 0 && !d_y)) return fail(LDPC_E_INVALID, "bad argument");
    LDPC_CUDA(cudaSetDevice(c->device));
    return launch_channel(c, seed, frame_begin, frames, std::sqrt(llr_variance(snr)), d_codewords, d_y,
                          (cudaStream_t) stream);
}

int ldpc_generator_codewords(const ldpc_code_t *c, uint64_t seed, uint64_t frame_begin, int64_t frames,
                             uint8_t *codewords) {
    if (!c || frames < 0 || (frames > 0 && !codewords)) return fail(LDPC_E_INVALID, "bad argument");
    if (frames == 0) return LDPC_OK;
    LDPC_CUDA(cudaSetDevice(c->device));
    uint8_t *d = nullptr;
    LDPC_CUDA(cudaMalloc((void **) &d, frames * (size_t) c->n));
    int st = launch_generator_codewords(c, seed, frame_begin, frames, d, 0);
    if (!st) {
        cudaError_t e = cudaMemcpy(codewords, d, frames * (size_t) c->n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) st = cuda_fail(e, "codeword download", __FILE__, __LINE__);
    }
    cudaFree(d);
    return st;
}

int ldpc_experiment_run(const ldpc_code_t *c, const ldpc_algo_cfg_t *cfg, double snr, uint64_t seed,
                        uint64_t frame_begin, uint64_t frame_count, int32_t codeword_source, const uint8_t *words,
                        uint64_t n_words, uint64_t counters[LDPC_CNT_COUNT], double *gpu_seconds) {
    if (!counters) return fail(LDPC_E_INVALID, "NULL argument");
    int st = experiment_check(c, cfg, codeword_source, words, n_words);
    if (st) return st;
    for (int i = 0; i < LDPC_CNT_COUNT; ++i) counters[i] = 0;
    if (gpu_seconds) *gpu_seconds = 0.0;
    if (frame_count == 0) return LDPC_OK;
    ExperimentJob job;
    st = experiment_enqueue(c, cfg, snr, seed, frame_begin, frame_count, codeword_source, words, n_words, &job);
    if (!st) st = experiment_collect(&job, counters, gpu_seconds);
    experiment_release(&job);
    return st;
}

// ---- multi-GPU: NCCL through dlopen (the library has no link-time dependency on it)

namespace {

struct Nccl {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

Nccl &nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        // inside a torch process the bundled libnccl.so.2 is already loaded and this returns it
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (n.lib) break;
        }
        if (!n.lib) return;
        auto sym = [&](const char *name) { return dlsym(n.lib, name); };
        n.GetUniqueId = (decltype(n.GetUniqueId)) sym("ncclGetUniqueId");
        n.CommInitRank = (decltype(n.CommInitRank)) sym("ncclCommInitRank");
        n.CommInitAll = (decltype(n.CommInitAll)) sym("ncclCommInitAll");
        n.AllReduce = (decltype(n.AllReduce)) sym("ncclAllReduce");
        n.CommDestroy = (decltype(n.CommDestroy)) sym("ncclCommDestroy");
        n.GroupStart = (decltype(n.GroupStart)) sym("ncclGroupStart");
        n.GroupEnd = (decltype(n.GroupEnd)) sym("ncclGroupEnd");
        n.GetErrorString = (decltype(n.GetErrorString)) sym("ncclGetErrorString");
        n.ok = n.GetUniqueId && n.CommInitRank && n.CommInitAll && n.AllReduce && n.CommDestroy && n.GroupStart && n.GroupEnd &&
               n.GetErrorString;
    });
    return n;
}

// NCCL prints its version banner to STDOUT at the first communicator when NCCL_DEBUG asks for it (the GPU boxes set it);
// the drivers' stdout is part of the drop-in surface (main.cpp / optimize_H.cpp / qpadmm_params.cpp print results
// there), so file descriptor 1 points at stderr while a communicator is being created.
struct StdoutToStderr {
    int saved = -1;
    StdoutToStderr() {
        fflush(stdout);
        saved = dup(1);
        if (saved >= 0) dup2(2, 1);
    }
    ~StdoutToStderr() {
        if (saved >= 0) {
            fflush(stdout);
            dup2(saved, 1);
            close(saved);
        }
    }
};

int nccl_fail(ncclResult_t r, const char *what) {
    return fail(LDPC_E_CUDA, std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error"));
}

#define LDPC_NCCL(call)                                         \
    do {                                                        \
        ncclResult_t r__ = (call);                              \
        if (r__ != ncclSuccess) return nccl_fail(r__, #call);   \
    } while (0)

// single-process communicators, one per device list (ncclCommInitAll is expensive: 0.1-1 s)
struct CommSet {
    std::vector<ncclComm_t> comms;
};
std::mutex g_comm_mu;
std::map<std::vector<int>, CommSet> g_comm_sets;

}  // namespace

struct ldpc_comm {
    ncclComm_t comm = nullptr;
    int device = 0, world = 1;
    cudaStream_t stream = nullptr;
    unsigned long long *buf = nullptr;
    size_t cap = 0;
};

int ldpc_experiment_run_multi(const ldpc_code_t *const *codes, int32_t n_devices, const ldpc_algo_cfg_t *cfg, double snr,
                              uint64_t seed, uint64_t frame_begin, uint64_t frame_count, int32_t codeword_source,
                              const uint8_t *words, uint64_t n_words, uint64_t counters[LDPC_CNT_COUNT],
                              double *gpu_seconds) {
    if (!codes || n_devices < 1 || !counters) return fail(LDPC_E_INVALID, "bad argument");
    std::vector<int> devices;
    for (int g = 0; g < n_devices; ++g) {
        int st = experiment_check(codes[g], cfg, codeword_source, words, n_words);
        if (st) return st;
        if (codes[g]->n != codes[0]->n || codes[g]->m != codes[0]->m || codes[g]->E != codes[0]->E)
            return fail(LDPC_E_INVALID, "the code handles of ldpc_experiment_run_multi must describe the same H");
        for (int d : devices)
            if (d == codes[g]->device) return fail(LDPC_E_INVALID, "two code handles on one device");
        devices.push_back(codes[g]->device);
    }
    for (int i = 0; i < LDPC_CNT_COUNT; ++i) counters[i] = 0;
    if (gpu_seconds) *gpu_seconds = 0.0;
    if (frame_count == 0) return LDPC_OK;
    if (n_devices == 1)
        return ldpc_experiment_run(codes[0], cfg, snr, seed, frame_begin, frame_count, codeword_source, words, n_words, counters,
                                   gpu_seconds);
    Nccl &nc = nccl();
    if (!nc.ok) return fail(LDPC_E_UNSUPPORTED, "libnccl.so.2 could not be loaded");
    CommSet *set = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_comm_mu);
        auto it = g_comm_sets.find(devices);
        if (it == g_comm_sets.end()) {
            CommSet cs;
            cs.comms.resize(n_devices);
            StdoutToStderr quiet;
            LDPC_NCCL(nc.CommInitAll(cs.comms.data(), n_devices, devices.data()));
            it = g_comm_sets.emplace(devices, cs).first;
        }
        set = &it->second;
    }
    // all shards are enqueued from this thread (launches are asynchronous), then one grouped all-reduce on the shards' streams
    std::vector<ExperimentJob> jobs(n_devices);
    int st = LDPC_OK;
    for (int g = 0; g < n_devices && !st; ++g) {
        const uint64_t b = frame_count * (uint64_t) g / (uint64_t) n_devices, e = frame_count * (uint64_t) (g + 1) / (uint64_t) n_devices;
        st = experiment_enqueue(codes[g], cfg, snr, seed, frame_begin + b, e - b, codeword_source, words, n_words, &jobs[g]);
    }
    if (!st) {
        std::lock_guard<std::mutex> lock(g_comm_mu);        // one collective at a time per communicator set
        ncclResult_t r = nc.GroupStart();
        for (int g = 0; g < n_devices && r == ncclSuccess; ++g)
            r = nc.AllReduce(jobs[g].counters, jobs[g].counters, LDPC_CNT_COUNT, ncclUint64, ncclSum, set->comms[g], jobs[g].stream);
        const ncclResult_t r2 = nc.GroupEnd();
        if (r != ncclSuccess || r2 != ncclSuccess) st = nccl_fail(r != ncclSuccess ? r : r2, "ncclAllReduce of the counter blocks");
    }
    uint64_t tmp[LDPC_CNT_COUNT];
    double longest = 0.0;
    for (int g = 0; g < n_devices; ++g) {                   // every stream is drained, also after a failure
        double secs = 0.0;
        const int sg = jobs[g].stream ? experiment_collect(&jobs[g], g == 0 ? counters : tmp, &secs) : LDPC_OK;
        if (!st) st = sg;
        longest = std::max(longest, secs);
        experiment_release(&jobs[g]);
    }
    if (gpu_seconds) *gpu_seconds = longest;
    return st;
}

int ldpc_comm_unique_id(uint8_t id[LDPC_COMM_ID_BYTES]) {
    if (!id) return fail(LDPC_E_INVALID, "NULL argument");
    Nccl &nc = nccl();
    if (!nc.ok) return fail(LDPC_E_UNSUPPORTED, "libnccl.so.2 could not be loaded");
    static_assert(sizeof(ncclUniqueId) == LDPC_COMM_ID_BYTES, "NCCL id size");
    ncclUniqueId u;
    LDPC_NCCL(nc.GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return LDPC_OK;
}

int ldpc_comm_init(int32_t rank, int32_t world, const uint8_t id[LDPC_COMM_ID_BYTES], int device, ldpc_comm_t **out) {
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return fail(LDPC_E_INVALID, "bad argument");
    Nccl &nc = nccl();
    if (!nc.ok) return fail(LDPC_E_UNSUPPORTED, "libnccl.so.2 could not be loaded");
    LDPC_CUDA(cudaSetDevice(device));
    ldpc_comm *c = new ldpc_comm;
    c->device = device;
    c->world = world;
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclResult_t r;
    {
        StdoutToStderr quiet;
        r = nc.CommInitRank(&c->comm, world, u, rank);
    }
    if (r != ncclSuccess) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { nc.CommDestroy(c->comm); delete c; return cuda_fail(e, "stream", __FILE__, __LINE__); }
    *out = c;
    return LDPC_OK;
}

int ldpc_allreduce_counters(ldpc_comm_t *c, uint64_t *counters, int32_t count) {
    if (!c || !counters || count < 0) return fail(LDPC_E_INVALID, "bad argument");
    if (count == 0) return LDPC_OK;
    Nccl &nc = nccl();
    LDPC_CUDA(cudaSetDevice(c->device));
    if ((size_t) count > c->cap) {
        cudaFree(c->buf);
        c->buf = nullptr;
        c->cap = 0;
        LDPC_CUDA(cudaMalloc((void **) &c->buf, sizeof(unsigned long long) * (size_t) count));
        c->cap = (size_t) count;
    }
    LDPC_CUDA(cudaMemcpyAsync(c->buf, counters, sizeof(uint64_t) * (size_t) count, cudaMemcpyHostToDevice, c->stream));
    LDPC_NCCL(nc.AllReduce(c->buf, c->buf, (size_t) count, ncclUint64, ncclSum, c->comm, c->stream));
    LDPC_CUDA(cudaMemcpyAsync(counters, c->buf, sizeof(uint64_t) * (size_t) count, cudaMemcpyDeviceToHost, c->stream));
    LDPC_CUDA(cudaStreamSynchronize(c->stream));
    return LDPC_OK;
}

void ldpc_comm_destroy(ldpc_comm_t *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && nccl().CommDestroy) nccl().CommDestroy(c->comm);
    cudaFree(c->buf);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
}

int ldpc_qpadmm_grid_run(const ldpc_code_t *c, int32_t points, const double *alpha, const double *mu,
                         int32_t max_iter, double eps_stop, double snr, uint64_t seed, uint64_t frame_begin,
                         uint64_t frame_count, int32_t codeword_source, const uint8_t *words, uint64_t n_words,
                         uint64_t *counters, double *gpu_seconds) {
    if (!c || !alpha || !mu || !counters || points < 0) return fail(LDPC_E_INVALID, "bad argument");
    if (max_iter < 0) return fail(LDPC_E_INVALID, "max_iter < 0");
    if (codeword_source < LDPC_CW_ZERO || codeword_source > LDPC_CW_GENERATOR)
        return fail(LDPC_E_INVALID, "unknown codeword source");
    if (codeword_source == LDPC_CW_TABLE && (!words || n_words == 0))
        return fail(LDPC_E_INVALID, "LDPC_CW_TABLE needs a codeword table");
    if (codeword_source == LDPC_CW_GENERATOR && (c->k <= 0 || !c->d.gen_cols))
        return fail(LDPC_E_INVALID, "LDPC_CW_GENERATOR needs ldpc_code_set_generator");
    if (gpu_seconds) *gpu_seconds = 0.0;
    for (int64_t i = 0; i < (int64_t) points * LDPC_CNT_COUNT; ++i) counters[i] = 0;
    if (points == 0 || frame_count == 0) return LDPC_OK;
    // pairs the check-centric kernel cannot take (infeasible, or a code outside its range) run one by one
    const int e_min = qpadmm_chk_e_min(c);
    std::vector<int> batch;
    std::vector<double> ba, bm;
    ldpc_algo_cfg_t cfg{LDPC_ALGO_QPADMM, max_iter, 1, 0, alpha[0], mu[0], eps_stop};
    for (int i = 0; i < points; ++i) {
        if (e_min > 0 && (double) e_min * mu[i] > alpha[i]) {
            batch.push_back(i); ba.push_back(alpha[i]); bm.push_back(mu[i]);
            continue;
        }
        cfg.alpha = alpha[i]; cfg.mu = mu[i];
        double secs = 0;
        int st = ldpc_experiment_run(c, &cfg, snr, seed, frame_begin, frame_count, codeword_source, words, n_words,
                                     counters + (size_t) i * LDPC_CNT_COUNT, &secs);
        if (st) return st;
        if (gpu_seconds) *gpu_seconds += secs;
    }
    if (batch.empty()) return LDPC_OK;
    LDPC_CUDA(cudaSetDevice(c->device));
    Slot &s = g_ctx.per_device[c->device][0];
    int st = slot_init(s);
    if (st) return st;
    const size_t nb = batch.size();
    uint8_t *d_words = nullptr;
    double *d_par = nullptr;
    unsigned long long *d_cnt = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<unsigned long long> host_cnt(nb * LDPC_CNT_COUNT);
    auto cleanup = [&]() {
        dev_free(d_words); dev_free(d_par); dev_free(d_cnt);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    };
    cudaError_t e = cudaSuccess;
    if (codeword_source == LDPC_CW_TABLE) {
        e = dev_malloc((void **) &d_words, n_words * (size_t) c->n);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_words, words, n_words * (size_t) c->n, cudaMemcpyHostToDevice, s.stream);
    }
    if (e == cudaSuccess) e = dev_malloc((void **) &d_par, sizeof(double) * 2 * nb);
    if (e == cudaSuccess) e = dev_malloc((void **) &d_cnt, sizeof(unsigned long long) * nb * LDPC_CNT_COUNT);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_par, ba.data(), sizeof(double) * nb, cudaMemcpyHostToDevice, s.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_par + nb, bm.data(), sizeof(double) * nb, cudaMemcpyHostToDevice, s.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * nb * LDPC_CNT_COUNT, s.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(s.queue, 0, sizeof(unsigned long long), s.stream);
    if (e != cudaSuccess) { cleanup(); return cuda_fail(e, "grid setup", __FILE__, __LINE__); }
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    FrameIO io;
    io.experiment = 1; io.seed = seed; io.frame_begin = frame_begin; io.cw_source = codeword_source;
    io.words = d_words; io.n_words = n_words; io.counters = d_cnt;
    cudaEventRecord(e0, s.stream);
    st = launch_qpadmm_chk(c, io, (int64_t) frame_count, llr_variance(snr), ba[0], bm[0], max_iter, eps_stop, s.queue,
                           s.stream, d_par, d_par + nb, (int64_t) nb);
    cudaEventRecord(e1, s.stream);
    if (st == LDPC_E_UNSUPPORTED) st = fail(LDPC_E_UNSUPPORTED, "grid launch: code outside the check-centric kernel's range");
    if (!st) {
        e = cudaMemcpyAsync(host_cnt.data(), d_cnt, sizeof(unsigned long long) * host_cnt.size(), cudaMemcpyDeviceToHost, s.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
        if (e != cudaSuccess) st = cuda_fail(e, "grid run", __FILE__, __LINE__);
    }
    if (!st) {
        for (size_t b = 0; b < nb; ++b)
            for (int k = 0; k < LDPC_CNT_COUNT; ++k)
                counters[(size_t) batch[b] * LDPC_CNT_COUNT + k] = host_cnt[b * LDPC_CNT_COUNT + k];
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (gpu_seconds) *gpu_seconds += ms * 1e-3;
    }
    cleanup();
    return st;
}

int ldpc_host_alloc(void **ptr, uint64_t bytes) {
    if (!ptr) return fail(LDPC_E_INVALID, "ptr is NULL");
    LDPC_CUDA(cudaMallocHost(ptr, bytes));
    return LDPC_OK;
}

int ldpc_host_free(void *ptr) {
    LDPC_CUDA(cudaFreeHost(ptr));
    return LDPC_OK;
}

int ldpc_debug_bpmath(int device, int32_t count, const double *a, const double *ev, const double *od, double *out_exp,
                      double *out_log) {
    if (count < 0 || (count > 0 && (!a || !ev || !od || !out_exp || !out_log))) return fail(LDPC_E_INVALID, "bad argument");
    if (count == 0) return LDPC_OK;
    return debug_bpmath(device, count, a, ev, od, out_exp, out_log);
}

int ldpc_debug_last_bp_kernel(void) { return ldpc::g_last_bp_kernel.load(); }

int ldpc_debug_last_qpadmm_kernel(void) { return ldpc::g_last_qpadmm_kernel.load(); }

int ldpc_debug_bp_layout(const ldpc_code_t *c, int32_t frames_per_cta, int32_t out[6]) {
    if (!c || !out) return fail(LDPC_E_INVALID, "NULL argument");
    return bp_lr_layout_stats(c, frames_per_cta, out);
}

int ldpc_measure_smem_peak(int device, double *gbytes_per_s) {
    if (!gbytes_per_s) return fail(LDPC_E_INVALID, "NULL argument");
    return measure_smem_peak(device, gbytes_per_s);
}

int ldpc_measure_fp64_peak(int device, double *gfma_per_s) {
    if (!gfma_per_s) return fail(LDPC_E_INVALID, "NULL argument");
    return measure_fp64_peak(device, gfma_per_s);
}

}  // extern "C"
