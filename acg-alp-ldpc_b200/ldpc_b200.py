"""ctypes binding of libldpc_b200.so (the C ABI in include/ldpc_b200.h).

This is the Python face of the product, used by tests/, bench.py and
__graft_entry__.py.  It mirrors the reference's decoder objects:

    BeliefPropagationDecoder(max_iter)                      algo/bp.h:208-222
    QPADMMDecoder(alpha, mu, max_iter=2000, eps_stop=1e-5)  algo/qp_admm.h:180-194

with ``decode(code, y, snr)`` taking a batch of channel words (frames x n raw
samples, not LLRs) and a compiled ``Code`` instead of the dense H of every call.
There is no CPU fallback: a missing library or missing GPU raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LDPC_B200_LIB") or os.path.join(HERE, "libldpc_b200.so")     # (the override is for A/B timing of two builds)

CNT_NAMES = ["total", "correct", "pseudo", "decoder_fail", "bit_errors", "sum_hamming", "sum_hamming_ok",
             "sum_hamming_wrong", "sum_iters", "frames_with_bits"]
ALGO_BP, ALGO_QPADMM = 0, 1
CW_ZERO, CW_TABLE, CW_GENERATOR = 0, 1, 2


class LdpcError(RuntimeError):
    pass


class CodeInfo(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("m", "n", "edges", "max_row_deg", "max_col_deg", "admm_blocks",
                                          "admm_n_var", "admm_rows", "admm_nnz", "admm_e_min", "k", "device",
                                          "admm_conflicts_natural", "admm_conflicts_laid_out")]


class AlgoCfg(C.Structure):
    _fields_ = [("algo", C.c_int32), ("max_iter", C.c_int32), ("early_exit", C.c_int32), ("reserved", C.c_int32),
                ("alpha", C.c_double), ("mu", C.c_double), ("eps_stop", C.c_double)]


_lib = None


def lib():
    """Load the library once; raise (never fall back) if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LdpcError("%s is missing: run `python acg-alp-ldpc_b200/build.py` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64, u64, i32, dbl = C.c_void_p, C.c_int64, C.c_uint64, C.c_int32, C.c_double
    L.ldpc_last_error.restype = C.c_char_p
    L.ldpc_abi_version.restype = C.c_int
    L.ldpc_device_count.argtypes = [C.POINTER(C.c_int)]
    L.ldpc_code_create.argtypes = [i32, i32, vp, vp, C.c_int, C.POINTER(vp)]
    L.ldpc_code_create_dense.argtypes = [i32, i32, vp, C.c_int, C.POINTER(vp)]
    L.ldpc_code_destroy.argtypes = [vp]
    L.ldpc_code_destroy.restype = None
    L.ldpc_code_info.argtypes = [vp, C.POINTER(CodeInfo)]
    L.ldpc_code_set_generator.argtypes = [vp, i32, vp]
    L.ldpc_bp_decode.argtypes = [vp, vp, i64, dbl, i32, i32, vp, vp, vp, vp]
    L.ldpc_bp_decode_device.argtypes = [vp, vp, i64, dbl, i32, i32, vp, vp, vp, vp, vp]
    L.ldpc_qpadmm_decode.argtypes = [vp, vp, i64, dbl, dbl, dbl, i32, dbl, vp, vp, vp, vp]
    L.ldpc_qpadmm_decode_device.argtypes = [vp, vp, i64, dbl, dbl, dbl, i32, dbl, vp, vp, vp, vp, vp]
    L.ldpc_channel_generate.argtypes = [vp, u64, u64, i64, dbl, vp, vp]
    L.ldpc_channel_generate_device.argtypes = [vp, u64, u64, i64, dbl, vp, vp, vp]
    L.ldpc_generator_codewords.argtypes = [vp, u64, u64, i64, vp]
    L.ldpc_experiment_run.argtypes = [vp, C.POINTER(AlgoCfg), dbl, u64, u64, u64, i32, vp, u64, vp,
                                      C.POINTER(dbl)]
    L.ldpc_qpadmm_grid_run.argtypes = [vp, i32, vp, vp, i32, dbl, dbl, u64, u64, u64, i32, vp, u64, vp, C.POINTER(dbl)]
    L.ldpc_experiment_run_multi.argtypes = [vp, i32, C.POINTER(AlgoCfg), dbl, u64, u64, u64, i32, vp, u64, vp,
                                            C.POINTER(dbl)]
    L.ldpc_comm_unique_id.argtypes = [vp]
    L.ldpc_comm_init.argtypes = [i32, i32, vp, C.c_int, C.POINTER(vp)]
    L.ldpc_allreduce_counters.argtypes = [vp, vp, i32]
    L.ldpc_comm_destroy.argtypes = [vp]
    L.ldpc_comm_destroy.restype = None
    L.ldpc_host_alloc.argtypes = [C.POINTER(vp), u64]
    L.ldpc_host_free.argtypes = [vp]
    L.ldpc_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(dbl)]
    L.ldpc_measure_smem_peak.argtypes = [C.c_int, C.POINTER(dbl)]
    L.ldpc_debug_bpmath.argtypes = [C.c_int, i32, vp, vp, vp, vp, vp]
    L.ldpc_debug_bp_layout.argtypes = [vp, i32, vp]
    L.ldpc_debug_last_qpadmm_kernel.argtypes = []
    L.ldpc_debug_last_bp_kernel.argtypes = []
    _lib = L
    return L


def _check(status):
    if status != 0:
        raise LdpcError("ldpc_b200 error %d: %s" % (status, lib().ldpc_last_error().decode()))


def device_count():
    n = C.c_int()
    _check(lib().ldpc_device_count(C.byref(n)))
    return n.value


def measure_fp64_peak(device=0):
    out = C.c_double()
    _check(lib().ldpc_measure_fp64_peak(device, C.byref(out)))
    return out.value


def measure_smem_peak(device=0):
    out = C.c_double()
    _check(lib().ldpc_measure_smem_peak(device, C.byref(out)))
    return out.value


def debug_bpmath(a, ev, od, device=0):
    a, ev, od = (np.ascontiguousarray(x, np.float64) for x in (a, ev, od))
    out_exp, out_log = np.empty_like(a), np.empty_like(a)
    _check(lib().ldpc_debug_bpmath(device, a.size, a.ctypes.data, ev.ctypes.data, od.ctypes.data, out_exp.ctypes.data,
                                   out_log.ctypes.data))
    return out_exp, out_log


def last_bp_kernel():
    """1: the likelihood-ratio kernel served the last BP launch, 2: the log-domain kernel, 0: none yet"""
    return int(lib().ldpc_debug_last_bp_kernel())


def last_qpadmm_kernel():
    """1: the check-centric kernel served the last QP-ADMM launch, 2: the block-per-lane kernel, 0: none yet"""
    return int(lib().ldpc_debug_last_qpadmm_kernel())


def _ptr(a):
    return None if a is None else a.ctypes.data


DATA = os.path.join(HERE, "data")


def load_rows(name):
    """Dense uint8 parity-check matrix from acg-alp-ldpc_b200/data/<name>.rows (the sparse text format the repo ships
    its matrices in: '#' comments, "m n", then "deg c0 c1 ..." per row; utils/parse_data.h: read_pcm_rows)."""
    lines = [l for l in open(os.path.join(DATA, name + ".rows")) if not l.startswith("#")]
    m, n = (int(x) for x in lines[0].split())
    H = np.zeros((m, n), np.uint8)
    for r, line in enumerate(lines[1:1 + m]):
        vals = [int(x) for x in line.split()]
        assert vals[0] == len(vals) - 1
        H[r, vals[1:]] = 1
    return H


class Comm:
    """One rank of the path's only collective: the all-reduce of the counter blocks, done by the library's
    own NCCL communicator.  `make_id()` on rank 0, hand the 128 bytes to every rank, then Comm(rank, world, id, device)."""

    @staticmethod
    def make_id():
        buf = np.zeros(128, np.uint8)
        _check(lib().ldpc_comm_unique_id(buf.ctypes.data))
        return buf

    def __init__(self, rank, world, comm_id, device=0):
        self._h = C.c_void_p()
        comm_id = np.ascontiguousarray(comm_id, np.uint8)
        assert comm_id.size == 128
        _check(lib().ldpc_comm_init(rank, world, comm_id.ctypes.data, device, C.byref(self._h)))

    def allreduce(self, counters):
        """sum of a dict (or array) of unsigned 64-bit counters over all ranks"""
        keys = sorted(counters) if isinstance(counters, dict) else None
        vec = np.array([counters[k] for k in keys] if keys else counters, np.uint64)
        _check(lib().ldpc_allreduce_counters(self._h, vec.ctypes.data, vec.size))
        return dict(zip(keys, (int(x) for x in vec))) if keys else vec

    def close(self):
        if self._h:
            lib().ldpc_comm_destroy(self._h)
            self._h = C.c_void_p()


def experiment_multi(codes, decoder, snr, seed, frame_begin, frame_count, source=CW_ZERO, words=None):
    """ldpc_experiment_run_multi: one Monte-Carlo point sharded over the devices of `codes` (handles of the same H)"""
    cfg = decoder.cfg()
    cnt = np.zeros(len(CNT_NAMES), np.uint64)
    secs = C.c_double()
    w = None if words is None else np.ascontiguousarray(words, np.uint8)
    handles = (C.c_void_p * len(codes))(*[c._h for c in codes])
    _check(lib().ldpc_experiment_run_multi(handles, len(codes), C.byref(cfg), snr, seed, frame_begin, frame_count, source,
                                           _ptr(w), 0 if w is None else w.shape[0], cnt.ctypes.data, C.byref(secs)))
    res = dict(zip(CNT_NAMES, (int(x) for x in cnt)))
    res["gpu_seconds"] = secs.value
    return res


class Code:
    """A parity-check matrix compiled and uploaded once (replaces the per-frame
    graph builds of algo/bp.h:136-153 and algo/qp_admm.h:13-102)."""

    def __init__(self, H=None, csr=None, shape=None, device=0):
        L = lib()
        self._h = C.c_void_p()
        if H is not None:
            H = np.ascontiguousarray(H, np.uint8)
            m, n = H.shape
            _check(L.ldpc_code_create_dense(m, n, H.ctypes.data, device, C.byref(self._h)))
        else:
            row_ptr, col_idx = (np.ascontiguousarray(a, np.int32) for a in csr)
            m, n = shape
            _check(L.ldpc_code_create(m, n, row_ptr.ctypes.data, col_idx.ctypes.data, device, C.byref(self._h)))
        info = CodeInfo()
        _check(L.ldpc_code_info(self._h, C.byref(info)))
        self.info = {k: getattr(info, k) for k, _ in CodeInfo._fields_}
        self.m, self.n = info.m, info.n

    def close(self):
        if self._h:
            lib().ldpc_code_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_generator(self, G):
        G = np.ascontiguousarray(G, np.uint8)
        _check(lib().ldpc_code_set_generator(self._h, G.shape[0], G.ctypes.data))
        self.info["k"] = G.shape[0]

    # ---- channel ----------------------------------------------------------
    def channel(self, seed, frame_begin, frames, snr, codewords=None):
        y = np.empty((frames, self.n), np.float64)
        cw = None if codewords is None else np.ascontiguousarray(codewords, np.uint8)
        _check(lib().ldpc_channel_generate(self._h, seed, frame_begin, frames, snr, _ptr(cw), y.ctypes.data))
        return y

    def channel_device(self, seed, frame_begin, frames, snr, d_y, d_codewords=0, stream=0):
        _check(lib().ldpc_channel_generate_device(self._h, seed, frame_begin, frames, snr, d_codewords or None, d_y,
                                                  stream or None))

    def generator_codewords(self, seed, frame_begin, frames):
        out = np.empty((frames, self.n), np.uint8)
        _check(lib().ldpc_generator_codewords(self._h, seed, frame_begin, frames, out.ctypes.data))
        return out

    # ---- decoders (host buffers) --------------------------------------------
    def _out(self, frames, soft):
        bits = np.empty((frames, self.n), np.uint8)
        ok = np.empty(frames, np.uint8)
        iters = np.empty(frames, np.int32)
        s = np.empty((frames, self.n), np.float64) if soft else None
        return bits, ok, iters, s

    def bp_decode(self, y, snr, max_iter, early_exit=True, soft=True):
        y = np.ascontiguousarray(y, np.float64).reshape(-1, self.n)
        bits, ok, iters, post = self._out(y.shape[0], soft)
        _check(lib().ldpc_bp_decode(self._h, y.ctypes.data, y.shape[0], snr, max_iter, int(early_exit),
                                    bits.ctypes.data, ok.ctypes.data, iters.ctypes.data, _ptr(post)))
        return bits, ok, iters, post

    def qpadmm_decode(self, y, snr, alpha, mu, max_iter=2000, eps_stop=1e-5, soft=True):
        y = np.ascontiguousarray(y, np.float64).reshape(-1, self.n)
        bits, ok, iters, v = self._out(y.shape[0], soft)
        _check(lib().ldpc_qpadmm_decode(self._h, y.ctypes.data, y.shape[0], snr, alpha, mu, max_iter, eps_stop,
                                        bits.ctypes.data, ok.ctypes.data, iters.ctypes.data, _ptr(v)))
        return bits, ok, iters, v

    # ---- decoders (device pointers, asynchronous on `stream`) -----------------
    def bp_decode_device(self, d_y, frames, snr, max_iter, early_exit, d_bits, d_ok, d_iters, d_post=0, stream=0):
        _check(lib().ldpc_bp_decode_device(self._h, d_y, frames, snr, max_iter, int(early_exit), d_bits, d_ok,
                                           d_iters, d_post or None, stream or None))

    def qpadmm_decode_device(self, d_y, frames, snr, alpha, mu, max_iter, eps_stop, d_bits, d_ok, d_iters, d_v=0,
                             stream=0):
        _check(lib().ldpc_qpadmm_decode_device(self._h, d_y, frames, snr, alpha, mu, max_iter, eps_stop, d_bits,
                                               d_ok, d_iters, d_v or None, stream or None))

    # ---- Monte-Carlo point ----------------------------------------------------
    def experiment(self, decoder, snr, seed, frame_begin, frame_count, source=CW_ZERO, words=None):
        cfg = decoder.cfg()
        cnt = np.zeros(len(CNT_NAMES), np.uint64)
        secs = C.c_double()
        w = None if words is None else np.ascontiguousarray(words, np.uint8)
        _check(lib().ldpc_experiment_run(self._h, C.byref(cfg), snr, seed, frame_begin, frame_count, source,
                                         _ptr(w), 0 if w is None else w.shape[0], cnt.ctypes.data,
                                         C.byref(secs)))
        res = dict(zip(CNT_NAMES, (int(x) for x in cnt)))
        res["gpu_seconds"] = secs.value
        return res


    def bp_layout(self, frames_per_cta):
        """layout statistics of the likelihood-ratio BP kernel (ldpc_debug_bp_layout)"""
        out = np.zeros(6, np.int32)
        _check(lib().ldpc_debug_bp_layout(self._h, frames_per_cta, out.ctypes.data))
        return dict(zip(("slots", "pad_even", "clash_v", "pairs_v", "clash_c", "pairs_c"), (int(x) for x in out)))

    def qpadmm_grid(self, alphas, mus, snr, max_iter, eps_stop, seed, frame_begin, frame_count, source=CW_ZERO,
                    words=None):
        """qpadmm_params.cpp:51-67 in one launch -> (list of counter dicts in the order of alphas/mus, gpu seconds)"""
        a = np.ascontiguousarray(alphas, np.float64)
        m = np.ascontiguousarray(mus, np.float64)
        assert a.shape == m.shape and a.ndim == 1
        cnt = np.zeros((a.size, len(CNT_NAMES)), np.uint64)
        secs = C.c_double()
        w = None if words is None else np.ascontiguousarray(words, np.uint8)
        _check(lib().ldpc_qpadmm_grid_run(self._h, a.size, a.ctypes.data, m.ctypes.data, max_iter, eps_stop, snr, seed,
                                          frame_begin, frame_count, source, _ptr(w), 0 if w is None else w.shape[0],
                                          cnt.ctypes.data, C.byref(secs)))
        return [dict(zip(CNT_NAMES, (int(x) for x in row))) for row in cnt], secs.value


class BeliefPropagationDecoder:
    """algo/bp.h:208-222; early_exit=False is the fixed-iteration measurement mode."""

    def __init__(self, max_iter, early_exit=True):
        self.max_iter, self.early_exit = int(max_iter), bool(early_exit)

    def name(self):
        return "BP"

    def cfg(self):
        return AlgoCfg(ALGO_BP, self.max_iter, int(self.early_exit), 0, 0.0, 0.0, 0.0)

    def decode(self, code, y, snr):
        """-> (bits, ok): ok[f] False means the reference returns an EMPTY word (bp.h:198)."""
        bits, ok, _, _ = code.bp_decode(y, snr, self.max_iter, self.early_exit, soft=False)
        return bits, ok.astype(bool)


class QPADMMDecoder:
    """algo/qp_admm.h:180-194"""

    def __init__(self, alpha, mu, max_iter=2000, eps_stop=1e-5):
        self.alpha, self.mu, self.max_iter, self.eps_stop = float(alpha), float(mu), int(max_iter), float(eps_stop)

    def name(self):
        return "QP-ADMM"

    def cfg(self):
        return AlgoCfg(ALGO_QPADMM, self.max_iter, 1, 0, self.alpha, self.mu, self.eps_stop)

    def decode(self, code, y, snr):
        bits, ok, _, _ = code.qpadmm_decode(y, snr, self.alpha, self.mu, self.max_iter, self.eps_stop, soft=False)
        return bits, ok.astype(bool)
