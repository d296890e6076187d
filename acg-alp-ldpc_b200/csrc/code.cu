// Host-side graph compiler: H (CSR) -> device tables, once per H.
//
// Replaces the reference's per-FRAME structure builds:
//   from_biadjacency_matrix + TannerGraph/VNode/CNode  (algo/bp.h:97-153)  -> BP edge tables
//   ConstructADMMProblem                               (algo/qp_admm.h:13-102) -> ADMM block tables
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

template <typename T>
static int upload(T **dst, const std::vector<T> &src) {
    size_t bytes = sizeof(T) * std::max<size_t>(src.size(), 1);
    LDPC_CUDA(cudaMalloc((void **) dst, bytes));
    if (!src.empty()) LDPC_CUDA(cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
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
    std::vector<int> csc_pos(E);
    for (int e = 0; e < E; ++e) {
        int p = fill[c->col_idx[e]]++;
        c->csc_edge[p] = e;
        csc_pos[e] = p;
    }
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
    std::vector<int> chk_order, var_order;
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
    if ((st = upload(&c->d.chk_rs, chk_rs))) return st;
    if ((st = upload(&c->d.var_rec, var_rec))) return st;
    if ((st = upload(&c->d.var_edges, var_edges))) return st;
    if ((st = upload(&c->d.col_ptr, colp))) return st;
    if ((st = upload(&c->d.row_ptr, rowp))) return st;
    if ((st = upload(&c->d.col_idx, coli))) return st;
    return LDPC_OK;
}

// The chain decomposition of qp_admm.h:59-92 expressed as blocks.
static int compile_admm(ldpc_code *c) {
    const int m = c->m, n = c->n;
    struct RawBlock { int var[3]; int nvars; int rows; };
    std::vector<RawBlock> raw;
    int next_aux = n;
    for (int r = 0; r < m; ++r) {
        const int *idx = c->col_idx.data() + c->row_ptr[r];
        int d = c->row_ptr[r + 1] - c->row_ptr[r];
        if (d == 0) continue;                                    // qp_admm.h:67-69
        if (d == 1) { raw.push_back({{idx[0], -1, -1}, 1, 1}); continue; }       // :70-74
        if (d == 2) { raw.push_back({{idx[0], idx[1], -1}, 2, 2}); continue; }   // :75-83
        int last = idx[0];                                       // :84-91
        for (int j = 1; j <= d - 2; ++j) {
            int third = (j == d - 2) ? idx[d - 1] : next_aux++;
            raw.push_back({{last, idx[j], third}, 3, 4});
            last = third;
        }
    }
    c->n_blocks = (int) raw.size();
    c->n_var = next_aux;
    if (c->n_var >= 65535 || c->n_blocks >= 16384)
        return fail(LDPC_E_UNSUPPORTED, "code too large for the 16-bit QP-ADMM tables");

    std::vector<AdmmBlock> blocks(raw.size());
    std::vector<std::vector<uint16_t>> inc(c->n_var);
    std::vector<int> e(c->n_var, 0);
    c->n_rows = 0;
    c->nnz = 0;
    for (size_t b = 0; b < raw.size(); ++b) {
        const RawBlock &rb = raw[b];
        int order[3] = {0, 1, 2};
        // ascending variable index; absent slots (-1) go last and point at the zero sentinel
        std::sort(order, order + 3, [&](int x, int y) {
            int vx = rb.var[x] < 0 ? 1 << 30 : rb.var[x], vy = rb.var[y] < 0 ? 1 << 30 : rb.var[y];
            return vx < vy || (vx == vy && x < y);
        });
        AdmmBlock ab;
        ab.meta = (uint16_t) (rb.rows << 8);
        for (int k = 0; k < 3; ++k) {
            int slot = order[k];
            ab.var[k] = (uint16_t) (rb.var[slot] < 0 ? c->n_var : rb.var[slot]);
            ab.meta |= (uint16_t) (slot << (2 * k));
        }
        blocks[b] = ab;
        for (int slot = 0; slot < rb.nvars; ++slot) {
            inc[rb.var[slot]].push_back((uint16_t) ((b << 2) | slot));
            e[rb.var[slot]] += rb.rows;          // one +-1 coefficient per row of the block
        }
        c->n_rows += rb.rows;
        c->nnz += rb.rows * rb.nvars;
    }
    std::vector<uint32_t> var_ptr(c->n_var + 1, 0);
    std::vector<uint16_t> inc_flat;
    for (int v = 0; v < c->n_var; ++v) {
        var_ptr[v] = (uint32_t) inc_flat.size();
        inc_flat.insert(inc_flat.end(), inc[v].begin(), inc[v].end());
    }
    var_ptr[c->n_var] = (uint32_t) inc_flat.size();
    c->n_inc = (int) inc_flat.size();
    // e_min as DecodeQPADMM computes it: over ALL variables, starting from 1e9 (qp_admm.h:108-111)
    c->e_min = 1000000000;
    for (int v = 0; v < c->n_var; ++v) c->e_min = std::min(c->e_min, e[v]);
    // balance: threads take variables round-robin from this order, so a warp sees equal degrees
    std::vector<uint16_t> order(c->n_var);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](uint16_t a, uint16_t b) { return inc[a].size() > inc[b].size(); });
    std::vector<uint8_t> e8(c->n_var);
    for (int v = 0; v < c->n_var; ++v) {
        if (e[v] > 255) return fail(LDPC_E_UNSUPPORTED, "column weight too large for the QP-ADMM tables");
        e8[v] = (uint8_t) e[v];
    }
    // blocks grouped by the order in which their slots are visited (at most three patterns occur)
    std::vector<uint16_t> blk_order(raw.size());
    std::iota(blk_order.begin(), blk_order.end(), 0);
    std::stable_sort(blk_order.begin(), blk_order.end(), [&](uint16_t a, uint16_t b) {
        return (blocks[a].meta & 0x3f) < (blocks[b].meta & 0x3f);
    });
    int st;
    if ((st = upload(&c->d.blocks, blocks))) return st;
    if ((st = upload(&c->d.blk_order, blk_order))) return st;
    if ((st = upload(&c->d.var_ptr, var_ptr))) return st;
    if ((st = upload(&c->d.inc, inc_flat))) return st;
    if ((st = upload(&c->d.var_order, order))) return st;
    if ((st = upload(&c->d.var_e, e8))) return st;
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
    cudaFree(c->d.chk_rs); cudaFree(c->d.var_rec); cudaFree(c->d.var_edges); cudaFree(c->d.col_ptr);
    for (auto &kv : c->bp_sched) { cudaFree(kv.second.jobs_v); cudaFree(kv.second.jobs_c); } cudaFree(c->d.row_ptr);
    cudaFree(c->d.col_idx); cudaFree(c->d.blocks); cudaFree(c->d.blk_order); cudaFree(c->d.var_ptr); cudaFree(c->d.inc);
    cudaFree(c->d.var_order); cudaFree(c->d.var_e); cudaFree(c->d.gen_cols);
    delete c;
}

int ldpc_code_info(const ldpc_code_t *c, ldpc_code_info_t *info) {
    if (!c || !info) return fail(LDPC_E_INVALID, "NULL argument");
    info->m = c->m; info->n = c->n; info->edges = c->E;
    info->max_row_deg = c->max_row_deg; info->max_col_deg = c->max_col_deg;
    info->admm_blocks = c->n_blocks; info->admm_n_var = c->n_var; info->admm_rows = c->n_rows;
    info->admm_nnz = c->nnz; info->admm_e_min = c->e_min; info->k = c->k; info->device = c->device;
    return LDPC_OK;
}

int ldpc_code_set_generator(ldpc_code_t *c, int32_t k, const uint8_t *G) {
    if (!c || !G || k <= 0) return fail(LDPC_E_INVALID, "bad generator");
    LDPC_CUDA(cudaSetDevice(c->device));
    int kw = (k + 31) / 32;
    std::vector<uint32_t> cols((size_t) c->n * kw, 0u);
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < c->n; ++j)
            if (G[(size_t) i * c->n + j]) cols[(size_t) j * kw + i / 32] |= 1u << (i % 32);
    cudaFree(c->d.gen_cols);
    c->d.gen_cols = nullptr;
    int st = upload(&c->d.gen_cols, cols);
    if (st) return st;
    c->k = k;
    c->k_words = kw;
    return LDPC_OK;
}

}  // extern "C"
