// Host-side graph compiler: H (CSR) -> device tables, once per H.
//
// Replaces the reference's per-FRAME structure builds:
//   from_biadjacency_matrix + TannerGraph/VNode/CNode  (algo/bp.h:97-153)  -> BP node tables (here)
//   ConstructADMMProblem                               (algo/qp_admm.h:13-102) -> ADMM block tables (admm_layout.cu)
// Only q (the LLRs) depends on the frame; everything else depends on H alone.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "ldpc_internal.h"

namespace ldpc {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int fail(int status, const std::string &msg) {
    g_last_error = msg;
    return status;
}

int cuda_fail(cudaError_t err, const char *what, const char *file, int line) {
    g_last_error = std::string("CUDA error: ") + cudaGetErrorString(err) + " in " + what + " (" + file + ":" +
                   std::to_string(line) + ")";
    return LDPC_E_CUDA;
}

double llr_variance(double snr) { return std::pow(10, -(snr / 10)) / 2; }

// ---- caching pool for device tables (see ldpc_internal.h)
namespace {
struct DevPool {
    std::mutex mu;
    std::map<std::pair<int, size_t>, std::vector<void *>> idle;      // (device, size class) -> blocks
    std::map<void *, std::pair<int, size_t>> owner;                   // every block handed out or idle
    std::map<int, size_t> idle_bytes;
};
DevPool &dev_pool() {
    static DevPool *p = new DevPool();        // leaked on purpose: host threads may free blocks while the process exits
    return *p;
}
size_t pool_class(size_t bytes) {             // powers of two from 512 B; above 1 MiB multiples of 1 MiB
    if (bytes > (1u << 20)) return (bytes + (1u << 20) - 1) & ~(size_t) ((1u << 20) - 1);
    size_t c = 512;
    while (c < bytes) c <<= 1;
    return c;
}
}  // namespace

cudaError_t dev_malloc(void **ptr, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const size_t cls = pool_class(std::max<size_t>(bytes, 1));
    DevPool &P = dev_pool();
    {
        std::lock_guard<std::mutex> lock(P.mu);
        auto it = P.idle.find({dev, cls});
        if (it != P.idle.end() && !it->second.empty()) {
            *ptr = it->second.back();
            it->second.pop_back();
            P.idle_bytes[dev] -= cls;
            return cudaSuccess;
        }
    }
    e = cudaMalloc(ptr, cls);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(P.mu);
    P.owner[*ptr] = {dev, cls};
    return cudaSuccess;
}

void dev_free(void *ptr) {
    if (!ptr) return;
    static const size_t cap = [] {
        const char *e = getenv("LDPC_POOL_CAP_MB");
        return (size_t) (e ? std::max(0, atoi(e)) : 256) << 20;
    }();
    DevPool &P = dev_pool();
    int dev = -1;
    {
        std::lock_guard<std::mutex> lock(P.mu);
        auto it = P.owner.find(ptr);
        if (it != P.owner.end()) {
            dev = it->second.first;
            const size_t cls = it->second.second;
            if (P.idle_bytes[dev] + cls <= cap) {
                P.idle[{dev, cls}].push_back(ptr);
                P.idle_bytes[dev] += cls;
                return;
            }
            P.owner.erase(it);
        }
    }
    int cur = 0;                              // over the cap (or not ours): really free, on the block's device
    cudaGetDevice(&cur);
    if (dev >= 0 && dev != cur) cudaSetDevice(dev);
    cudaFree(ptr);
    if (dev >= 0 && dev != cur) cudaSetDevice(cur);
}

template <typename T>
static int upload(T **dst, const std::vector<T> &src) {
    size_t bytes = sizeof(T) * std::max<size_t>(src.size(), 1);
    LDPC_CUDA(dev_malloc((void **) dst, bytes));
    if (!src.empty()) LDPC_CUDA(upload_sync(*dst, src.data(), sizeof(T) * src.size()));
    return LDPC_OK;
}

static int compile_bp(ldpc_code *c) {
    const int m = c->m, n = c->n, E = c->E;
    // CSC: positions p enumerate edges variable-major, rows ascending inside a variable
    c->col_ptr.assign(n + 1, 0);
    for (int e = 0; e < E; ++e) c->col_ptr[c->col_idx[e] + 1]++;
    for (int v = 0; v < n; ++v) {
        c->max_col_deg = std::max(c->max_col_deg, c->col_ptr[v + 1]);
        c->col_ptr[v + 1] += c->col_ptr[v];
    }
    c->csc_edge.assign(E, 0);
    std::vector<int> fill(c->col_ptr.begin(), c->col_ptr.end() - 1);
    for (int e = 0; e < E; ++e) c->csc_edge[fill[c->col_idx[e]]++] = e;
    for (int r = 0; r < m; ++r) c->max_row_deg = std::max(c->max_row_deg, c->row_ptr[r + 1] - c->row_ptr[r]);
    // rank nodes by degree (stable), dropping degree-0 nodes, and record the degree classes
    auto rank_by_degree = [](int count, const std::vector<int> &ptr, std::vector<int> &order,
                             std::vector<BpClass> &classes) {
        order.clear();
        for (int i = 0; i < count; ++i)
            if (ptr[i + 1] > ptr[i]) order.push_back(i);
        std::stable_sort(order.begin(), order.end(),
                         [&](int a, int b) { return ptr[a + 1] - ptr[a] < ptr[b + 1] - ptr[b]; });
        classes.clear();
        for (int k = 0; k < (int) order.size(); ++k) {
            const int d = ptr[order[k] + 1] - ptr[order[k]];
            if (classes.empty() || classes.back().degree != d) classes.push_back(BpClass{d, k, 0});
            classes.back().count++;
        }
    };
    std::vector<int> &chk_order = c->chk_order, &var_order = c->var_order;
    rank_by_degree(m, c->row_ptr, chk_order, c->chk_classes);
    rank_by_degree(n, c->col_ptr, var_order, c->var_classes);
    std::vector<uint16_t> chk_rs(chk_order.size());
    for (size_t k = 0; k < chk_order.size(); ++k) chk_rs[k] = (uint16_t) c->row_ptr[chk_order[k]];
    std::vector<BpVarRec> var_rec(var_order.size());
    std::vector<uint16_t> var_edges;
    for (size_t k = 0; k < var_order.size(); ++k) {
        const int v = var_order[k];
        var_rec[k] = BpVarRec{(uint16_t) v, (uint16_t) var_edges.size()};
        for (int p = c->col_ptr[v]; p < c->col_ptr[v + 1]; ++p) var_edges.push_back((uint16_t) c->csc_edge[p]);
    }
    std::vector<uint16_t> colp(c->col_ptr.begin(), c->col_ptr.end());
    std::vector<uint16_t> rowp(c->row_ptr.begin(), c->row_ptr.end());
    std::vector<uint16_t> coli(c->col_idx.begin(), c->col_idx.end());
    int st;
    TableStager stage;
    stage.add(&c->d.chk_rs, chk_rs);
    stage.add(&c->d.var_rec, var_rec);
    stage.add(&c->d.var_edges, var_edges);
    stage.add(&c->d.col_ptr, colp);
    stage.add(&c->d.row_ptr, rowp);
    stage.add(&c->d.col_idx, coli);
    if ((st = stage.commit(&c->d.blob_bp))) return st;
    return LDPC_OK;
}

}  // namespace ldpc

using namespace ldpc;

extern "C" {

int ldpc_abi_version(void) { return LDPC_B200_ABI_VERSION; }

const char *ldpc_last_error(void) { return g_last_error.c_str(); }

int ldpc_device_count(int *count) {
    if (!count) return fail(LDPC_E_INVALID, "count is NULL");
    LDPC_CUDA(cudaGetDeviceCount(count));
    return LDPC_OK;
}

int ldpc_code_create(int32_t m, int32_t n, const int32_t *row_ptr, const int32_t *col_idx, int device,
                     ldpc_code_t **out) {
    if (!out) return fail(LDPC_E_INVALID, "out is NULL");
    *out = nullptr;
    if (m <= 0 || n <= 0 || !row_ptr || !col_idx) return fail(LDPC_E_INVALID, "empty parity-check matrix");
    if (row_ptr[0] != 0) return fail(LDPC_E_INVALID, "row_ptr[0] must be 0");
    int E = row_ptr[m];
    if (n >= 65535 || m >= 65535 || E >= 65535)
        return fail(LDPC_E_UNSUPPORTED, "code too large for the 16-bit edge tables");
    for (int r = 0; r < m; ++r) {
        if (row_ptr[r + 1] < row_ptr[r]) return fail(LDPC_E_INVALID, "row_ptr must be non-decreasing");
        for (int e = row_ptr[r]; e < row_ptr[r + 1]; ++e) {
            if (col_idx[e] < 0 || col_idx[e] >= n) return fail(LDPC_E_INVALID, "column index out of range");
            if (e > row_ptr[r] && col_idx[e] <= col_idx[e - 1])
                return fail(LDPC_E_INVALID, "columns must be strictly ascending within a row");
        }
    }
    int ndev = 0;
    LDPC_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(LDPC_E_CUDA, "no such CUDA device");
    LDPC_CUDA(cudaSetDevice(device));

    ldpc_code *c = new ldpc_code();
    c->device = device;
    c->m = m;
    c->n = n;
    c->E = E;
    c->row_ptr.assign(row_ptr, row_ptr + m + 1);
    c->col_idx.assign(col_idx, col_idx + E);
    int st = compile_bp(c);
    if (st == LDPC_OK) st = compile_admm(c);
    if (st != LDPC_OK) {
        ldpc_code_destroy(c);
        return st;
    }
    *out = c;
    return LDPC_OK;
}

int ldpc_code_create_dense(int32_t m, int32_t n, const uint8_t *H, int device, ldpc_code_t **out) {
    if (!H || m <= 0 || n <= 0) return fail(LDPC_E_INVALID, "empty parity-check matrix");
    std::vector<int32_t> row_ptr(m + 1, 0), col_idx;
    for (int r = 0; r < m; ++r) {
        for (int c = 0; c < n; ++c)
            if (H[(size_t) r * n + c]) col_idx.push_back(c);
        row_ptr[r + 1] = (int32_t) col_idx.size();
    }
    if (col_idx.empty()) col_idx.push_back(0);
    return ldpc_code_create(m, n, row_ptr.data(), col_idx.data(), device, out);
}

void ldpc_code_destroy(ldpc_code_t *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    dev_free(c->d.blob_bp);
    dev_free(c->d.blob_admm);
    dev_free(c->d.gen_cols);
    free_chk_tables(c);
    for (auto &kv : c->bp_sched) { dev_free(kv.second.jobs_v); dev_free(kv.second.jobs_c); }
    for (auto &kv : c->bp_lr_sched) {
        dev_free(kv.second.rec_v); dev_free(kv.second.steps); dev_free(kv.second.var_store);
    }
    delete c;
}

int ldpc_code_info(const ldpc_code_t *c, ldpc_code_info_t *info) {
    if (!c || !info) return fail(LDPC_E_INVALID, "NULL argument");
    info->m = c->m; info->n = c->n; info->edges = c->E;
    info->max_row_deg = c->max_row_deg; info->max_col_deg = c->max_col_deg;
    info->admm_blocks = c->n_blocks; info->admm_n_var = c->n_var; info->admm_rows = c->n_rows;
    info->admm_nnz = c->nnz; info->admm_e_min = c->e_min; info->k = c->k; info->device = c->device;
    info->admm_conflicts_natural = (int32_t) c->admm_conflicts_before;
    info->admm_conflicts_laid_out = (int32_t) c->admm_conflicts_after;
    return LDPC_OK;
}

int ldpc_code_set_generator(ldpc_code_t *c, int32_t k, const uint8_t *G) {
    if (!c || !G || k <= 0) return fail(LDPC_E_INVALID, "bad generator");
    if (k > 512) return fail(LDPC_E_UNSUPPORTED, "generators with more than 512 rows are not supported");
    LDPC_CUDA(cudaSetDevice(c->device));
    int kw = (k + 31) / 32;
    std::vector<uint32_t> cols((size_t) c->n * kw, 0u);
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < c->n; ++j)
            if (G[(size_t) i * c->n + j]) cols[(size_t) j * kw + i / 32] |= 1u << (i % 32);
    dev_free(c->d.gen_cols);
    c->d.gen_cols = nullptr;
    int st = upload(&c->d.gen_cols, cols);
    if (st) return st;
    c->k = k;
    c->k_words = kw;
    return LDPC_OK;
}

}  // extern "C"
