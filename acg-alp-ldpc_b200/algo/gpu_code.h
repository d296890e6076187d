// Host-side cache of compiled code handles.  The reference API passes the dense H
// on EVERY decode call (algo/algo.h:8) and rebuilds its graph each time
// (bp.h:136-153, qp_admm.h:13-102); here H is hashed, compiled once per
// (content, device) with ldpc_code_create and reused.  Thread-safe: decode() is
// called concurrently from many pthreads (experiment.h:128-130).
#ifndef LDPC_B200_ALGO_GPU_CODE_H
#define LDPC_B200_ALGO_GPU_CODE_H

#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>

#include "algo.h"

namespace ldpc_host {

[[noreturn]] inline void die(const char *what) {
    cerr << "ldpc_b200: " << what << ": " << ldpc_last_error() << endl;
    abort();   // the reference signals nothing either (asserts only); fail loudly, never fall back
}

struct CsrMatrix {
    int m = 0, n = 0;
    vector<int32_t> row_ptr, col_idx;
    bool operator<(const CsrMatrix &o) const {
        if (m != o.m) return m < o.m;
        if (n != o.n) return n < o.n;
        if (row_ptr != o.row_ptr) return row_ptr < o.row_ptr;
        return col_idx < o.col_idx;
    }
};

inline CsrMatrix to_csr(const TMatrix &H) {
    CsrMatrix c;
    c.m = (int) H.size();
    c.n = H.empty() ? 0 : (int) H[0].size();
    c.row_ptr.push_back(0);
    for (const TCodeword &row : H) {
        for (int j = 0; j < (int) row.size(); ++j)
            if (row[j]) c.col_idx.push_back(j);
        c.row_ptr.push_back((int32_t) c.col_idx.size());
    }
    return c;
}

// A handle stays alive while somebody holds the shared_ptr: the cache may evict it while another host thread is still
// decoding with it (optimize_H.cpp evaluates several proposals, i.e. several H, concurrently).
typedef shared_ptr<ldpc_code_t> CodeRef;

class CodeCache {
public:
    static CodeCache &instance() {
        static CodeCache cache;
        return cache;
    }

    CodeRef get(const TMatrix &H, int device = 0) {
        CsrMatrix key = to_csr(H);
        {
            lock_guard<mutex> lock(mu_);
            auto it = codes_.find(make_pair(device, key));
            if (it != codes_.end()) return it->second;
        }
        // compile and upload OUTSIDE the lock: optimize_H.cpp evaluates several proposals (several new H) at once, one
        // host thread and one GPU each, and a compile takes milliseconds
        vector<int32_t> cols = key.col_idx;
        if (cols.empty()) cols.push_back(0);
        ldpc_code_t *raw = nullptr;
        if (ldpc_code_create(key.m, key.n, key.row_ptr.data(), cols.data(), device, &raw) != LDPC_OK)
            die("ldpc_code_create");
        CodeRef fresh(raw, [](ldpc_code_t *c) { ldpc_code_destroy(c); });
        lock_guard<mutex> lock(mu_);
        CodeRef &slot = codes_[make_pair(device, key)];
        if (!slot) {                                      // (else another thread was faster: ours is dropped)
            slot = fresh;
            if (codes_.size() > 64) evict(device, key);   // optimize_H.cpp proposes a new H every step
        }
        return slot;
    }

private:
    void evict(int device, const CsrMatrix &keep) {
        for (auto it = codes_.begin(); it != codes_.end();) {
            if (it->first.first == device && !(it->first.second < keep) && !(keep < it->first.second)) ++it;
            else it = codes_.erase(it);                   // destroyed when its last user lets go
        }
    }
    mutex mu_;
    map<pair<int, CsrMatrix>, CodeRef> codes_;
};

inline int visible_gpus() {
    int count = 0;
    if (ldpc_device_count(&count) != LDPC_OK || count < 1) die("no CUDA device (there is no CPU fallback)");
    if (const char *cap = getenv("LDPC_GPUS")) count = max(1, min(count, atoi(cap)));
    return count;
}

}  // namespace ldpc_host

#endif
