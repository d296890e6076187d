// Internal definitions shared by the translation units of libldpc_b200.so.
// Nothing here is part of the ABI (include/ldpc_b200.h is).
#ifndef LDPC_B200_INTERNAL_H
#define LDPC_B200_INTERNAL_H

#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "ldpc_b200.h"

namespace ldpc {

// ---- error plumbing -------------------------------------------------------
void set_error(const std::string &msg);
int fail(int status, const std::string &msg);
int cuda_fail(cudaError_t err, const char *what, const char *file, int line);

#define LDPC_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t err__ = (call);                                            \
        if (err__ != cudaSuccess) return ::ldpc::cuda_fail(err__, #call, __FILE__, __LINE__); \
    } while (0)

// ---- device tables ----------------------------------------------------------
// BP works per NODE.  All messages of a frame live in ONE array indexed by the CSR
// position of the edge: the variable phase reads the C->V message of an edge and
// overwrites it with the V->C message, the check phase does the opposite, so each
// phase updates its node's edges in place.  Nodes are ranked by degree so that a
// warp runs one degree-specialised, fully unrolled code path.
struct BpVarRec {
    uint16_t var;     // variable index
    uint16_t off;     // start of its edge list (CSR positions, rows ascending) in var_edges
};
struct BpJob {        // the work of one warp in one round: node ranks [first, first + count) x F frames
    uint16_t degree;  // 0 = nothing to do
    uint16_t first;
    uint16_t count;
    uint16_t pad;
};
struct BpClass { int degree, first, count; };
struct BpSchedule {   // device arrays, rounds x warps jobs each (log-domain kernel)
    BpJob *jobs_v = nullptr, *jobs_c = nullptr;
    int rounds_v = 0, rounds_c = 0;
};
// likelihood-ratio BP kernel (bp_lr_kernel.cu): tables depend on the frames per CTA (byte offsets are pre-scaled)
struct BpLrSchedule {
    uint32_t *rec_v = nullptr;    // variable node records
    uint32_t *steps = nullptr;    // one word per (step row, thread of a team): check-pass rows, then variable-pass rows (lr_step_word)
    uint16_t *var_store = nullptr; // variable index -> storage index of its per-variable arrays
    int rec_words = 0, steps_c = 0, steps_v = 0, n_slots = 0, pad_even = 0;
    int clash_v = 0, pairs_v = 0, clash_c = 0, pairs_c = 0;   // layout statistics (ldpc_debug_bp_layout)
};

// QP-ADMM works per BLOCK: one three-variable check of the chain decomposition
// (qp_admm.h:34-57, 84-91) with its 4 inequality rows, or a degree-2 / degree-1
// check (2 rows / 1 row, qp_admm.h:70-83).  Row q of a block has coefficient +1
// for slot q and for every slot in row 3, else -1.
// Variables and blocks are addressed by RANK (their position in the kernel's processing
// and storage order, chosen by admm_layout.cu to avoid shared-memory bank conflicts).
struct AdmmBlock {    // indexed by block rank
    uint16_t var[3];  // RANKS of the block's variables in ASCENDING variable-index order (the residual
                      // is accumulated in that order, qp_admm.h:144-151); absent -> n_var (a zero)
    uint16_t meta;    // bits 0-1 / 2-3 / 4-5: slot of var[0] / var[1] / var[2]; bits 8-10: rows
};
struct AdmmVarRec {   // indexed by variable rank
    uint16_t inc_start;   // first incidence word of the variable in admm_inc
    uint8_t inc_count;    // number of blocks the variable belongs to
    uint8_t pad0;
    uint16_t e;           // sum of squared coefficients of its column (qp_admm.h:94-99)
    uint16_t pad1;
};

// tables of the check-centric QP-ADMM kernel (qpadmm_chk_kernel.cu), built on first use
struct AdmmChkTables {
    bool built = false, supported = false;
    uint32_t *chk_tab = nullptr, *var_stream = nullptr;
    uint16_t *var_rank = nullptr, *var_e = nullptr;
    uint32_t plane_base[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    int special_lo = 0, special_hi = 0;
    int n_chk = 0, n_chunks = 0, n_inc = 0, n_slots = 0, tab_stride = 0, max_nb = 0, e_min = 0;
    int stream_rows = 0;          // rows of var_stream (words per lane column, two rows of padding included)
    void *blob = nullptr;         // the one device allocation behind the four tables
};

struct DeviceTables {
    // BP
    uint16_t *chk_rs = nullptr;    // check rank -> CSR position of its first edge
    BpVarRec *var_rec = nullptr;   // variable rank -> record
    uint16_t *var_edges = nullptr; // E: CSR positions of the edges of each variable
    uint16_t *col_ptr = nullptr;   // n+1, CSC offsets
    uint16_t *row_ptr = nullptr;   // m+1
    uint16_t *col_idx = nullptr;   // E, variable of CSR position e
    // QP-ADMM
    AdmmBlock *blocks = nullptr;       // n_blocks, rank order
    AdmmVarRec *admm_var = nullptr;    // n_var, rank order
    uint32_t *admm_inc = nullptr;      // incidence words: block rank | sign-flip bits 31/30/29 for rows 0/1/2,
                                       // blocks in ascending row order per variable (the reference's gather order)
    uint16_t *admm_var_id = nullptr;   // variable rank -> variable index
    uint16_t *admm_var_rank = nullptr; // variable index (< n) -> rank
    // generator (optional): column j of G packed over k bits, k_words words per column
    uint32_t *gen_cols = nullptr;
    // the device allocations behind the BP and the QP-ADMM tables (one blob each, TableStager)
    void *blob_bp = nullptr, *blob_admm = nullptr;
};

}  // namespace ldpc

struct ldpc_code {
    int device = 0;
    int m = 0, n = 0, E = 0;
    int max_row_deg = 0, max_col_deg = 0;
    int n_blocks = 0, n_var = 0, n_rows = 0, nnz = 0, n_inc = 0, e_min = 0;
    int k = 0, k_words = 0;
    long admm_conflicts_before = 0, admm_conflicts_after = 0;   // replayed wavefronts per iteration (layout cost)
    // host copies (also used by tests through ldpc_code_info)
    std::vector<int> row_ptr, col_idx, col_ptr, csc_edge;
    std::vector<ldpc::BpClass> chk_classes, var_classes;   // nodes of equal degree are adjacent in rank order
    std::vector<int> chk_order, var_order;                 // rank -> node index (degree-0 nodes dropped)
    ldpc::DeviceTables d;
    // BP launch schedules, built on first use per (frames per CTA or team, warps per CTA or team)
    mutable std::mutex sched_mu;
    mutable std::map<std::pair<int, int>, ldpc::BpSchedule> bp_sched;
    mutable std::map<std::pair<int, int>, ldpc::BpLrSchedule> bp_lr_sched;
    mutable ldpc::AdmmChkTables admm_chk[3];     // frames per CTA 1, 2, 4
};

namespace ldpc {

// kernels / launchers (bp_kernel.cu, qpadmm_kernel.cu, channel_kernel.cu)
struct FrameIO {
    // decode mode: inputs/outputs per frame (device pointers; outputs may be null)
    const double *y = nullptr;
    uint8_t *bits = nullptr;
    uint8_t *ok = nullptr;
    int32_t *iters = nullptr;
    double *soft = nullptr;  // BP: posterior LLR, QP-ADMM: v[0..n)
    // experiment mode: y is generated on device
    int experiment = 0;
    uint64_t seed = 0, frame_begin = 0;
    int cw_source = LDPC_CW_ZERO;
    const uint8_t *words = nullptr;
    uint64_t n_words = 0;
    unsigned long long *counters = nullptr;  // LDPC_CNT_COUNT device words
};

int compile_admm(ldpc_code *code);

int launch_bp(const ldpc_code *code, const FrameIO &io, int64_t frames, double var, int max_iter,
              int early_exit, unsigned long long *queue, cudaStream_t stream);
// the two BP kernels behind launch_bp: likelihood-ratio domain (default) and log domain (large node degrees)
int launch_bp_lr(const ldpc_code *code, const FrameIO &io, int64_t frames, double var, int max_iter,
                 int early_exit, unsigned long long *queue, cudaStream_t stream);
int launch_bp_log(const ldpc_code *code, const FrameIO &io, int64_t frames, double var, int max_iter,
                  int early_exit, unsigned long long *queue, cudaStream_t stream);
double bp_lr_cap(const ldpc_code *code, double *llr_cap_out);
int bp_lr_layout_stats(const ldpc_code *code, int F, int32_t out[6]);
extern std::atomic<int> g_last_bp_kernel;        // 1 likelihood-ratio, 2 log-domain (testing hook)
extern std::atomic<int> g_last_qpadmm_kernel;    // 1 check-centric, 2 block-per-lane (testing hook)
int launch_qpadmm(const ldpc_code *code, const FrameIO &io, int64_t frames, double var, double alpha,
                  double mu, int max_iter, double eps_stop, unsigned long long *queue, cudaStream_t stream);
// the two QP-ADMM kernels behind launch_qpadmm: check-centric (checks of degree 0..12, variables of at most 15 edges) and block-per-lane (any code)
// grid_points > 0: `frames` frames under each of grid_points (alpha, mu) pairs (device arrays), counters per point
int launch_qpadmm_chk(const ldpc_code *code, const FrameIO &io, int64_t frames, double var, double alpha,
                      double mu, int max_iter, double eps_stop, unsigned long long *queue, cudaStream_t stream,
                      const double *grid_alpha = nullptr, const double *grid_mu = nullptr, int64_t grid_points = 0);
int qpadmm_chk_e_min(const ldpc_code *code);
int launch_qpadmm_blk(const ldpc_code *code, const FrameIO &io, int64_t frames, double var, double alpha,
                      double mu, int max_iter, double eps_stop, unsigned long long *queue, cudaStream_t stream);
void free_chk_tables(ldpc_code *code);
int launch_channel(const ldpc_code *code, uint64_t seed, uint64_t frame_begin, int64_t frames, double sigma,
                   const uint8_t *d_codewords, double *d_y, cudaStream_t stream);
int launch_generator_codewords(const ldpc_code *code, uint64_t seed, uint64_t frame_begin, int64_t frames,
                               uint8_t *d_codewords, cudaStream_t stream);
int measure_fp64_peak(int device, double *gfma_per_s);
int measure_smem_peak(int device, double *gbytes_per_s);
int debug_bpmath(int device, int count, const double *a, const double *ev, const double *od, double *out_exp,
                 double *out_log);

// Device memory of tables and per-call scratch comes from a small caching pool (code.cu): cudaFree synchronises the whole
// device -- it waits for the kernels of every other host thread -- and cudaMalloc / cudaFree take a process-wide lock, so
// a host thread that builds and drops a code handle per evaluation (optimize_H.cpp: a new H per proposal, several
// proposals in flight) would serialise all the others.  Freed blocks are kept per device and size class and handed out
// again; at most LDPC_POOL_CAP_MB (default 256) of idle blocks are kept per device.
cudaError_t dev_malloc(void **ptr, size_t bytes);
void dev_free(void *ptr);

// Host-to-device upload of a table that a kernel on ANOTHER (non-blocking) stream will read.  cudaMemcpy from pageable
// memory returns once the source has been staged -- the DMA to the device may still be in flight on the legacy stream,
// and a kernel launched on a non-blocking stream does not wait for it.  (Seen once in 9000 evaluations of optimize_H
// with two processes per GPU: a whole evaluation ran on a partly uploaded table blob.)  So: copy on the calling thread's
// own stream and wait for THAT stream -- complete on return, and no other thread's kernels are waited for.
inline cudaError_t upload_sync(void *dst, const void *src, size_t bytes) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, cudaStreamPerThread);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(cudaStreamPerThread);
}

// Several tables, ONE allocation and ONE host-to-device copy: every CUDA runtime call takes a process-wide lock, and
// with a code handle per proposal and a host thread per GPU (optimize_H.cpp) the ~20 small uploads of a handle were what
// the evaluation threads queued for.  add() stages a table and remembers where its device pointer goes; commit()
// allocates the blob (256-byte aligned parts), copies once and sets the pointers.  Only the blob is freed.
class TableStager {
public:
    template <typename T>
    void add(T **dst, const std::vector<T> &src) {
        const size_t off = (host_.size() + 255) & ~(size_t) 255;
        host_.resize(off + std::max<size_t>(sizeof(T) * src.size(), 1));
        if (!src.empty()) memcpy(host_.data() + off, src.data(), sizeof(T) * src.size());
        items_.push_back(Item{reinterpret_cast<void **>(dst), off});
    }
    int commit(void **blob) {
        LDPC_CUDA(dev_malloc(blob, std::max<size_t>(host_.size(), 1)));
        if (!host_.empty()) LDPC_CUDA(upload_sync(*blob, host_.data(), host_.size()));
        for (const Item &it : items_) *it.dst = static_cast<char *>(*blob) + it.off;
        return LDPC_OK;
    }

private:
    struct Item { void **dst; size_t off; };
    std::vector<char> host_;
    std::vector<Item> items_;
};

// Opts a kernel into the full dynamic shared memory of an SM (227 KB minus its static share).  The attribute is
// per-function state shared by every host thread: setting it to the size of ONE launch races with a concurrent launch
// of another code that needs more (optimize_H.cpp evaluates several H at once) -- "invalid argument" at launch; the
// constant maximum cannot race.
template <typename Kernel>
inline cudaError_t allow_max_dynamic_smem(Kernel kernel) {
    // once per (kernel, device): the two runtime calls take the process-wide lock the launches of other threads wait for
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, bool> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const std::pair<const void *, int> key(reinterpret_cast<const void *>(kernel), dev);
    {
        std::lock_guard<std::mutex> lock(mu);
        if (done.count(key)) return cudaSuccess;
    }
    cudaFuncAttributes attr;
    e = cudaFuncGetAttributes(&attr, kernel);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int) attr.sharedSizeBytes);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    done[key] = true;
    return cudaSuccess;
}

// sigma^2 exactly as utils/channel.h:12 computes it on the host
double llr_variance(double snr);

}  // namespace ldpc

#endif
