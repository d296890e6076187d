"""ctypes loaders for the two CHECKERS under oracle/ -- test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product
(acg-alp-ldpc_b200/) never does.

  Oracle  -> oracle/libldpc_oracle.so   our plain-C restatement (oracle/ldpc_oracle.c)
  Ref     -> oracle/_ref/libref_oracle.so   the unmodified reference headers behind
             oracle/ref_harness.cpp (present when built in the dev container)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libldpc_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_oracle.so")
REFERENCE_ROOT = os.environ.get("LDPC_REFERENCE_ROOT", "/root/reference")

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")

CNT_NAMES = ["total", "correct", "pseudo", "decoder_fail", "bit_errors", "sum_hamming", "sum_hamming_ok",
             "sum_hamming_wrong", "sum_iters", "frames_with_bits"]


def build(ref=True):
    """Compile the checkers (idempotent).  The reference harness is only built
    where the reference tree exists (the dev container); elsewhere the prebuilt
    oracle/_ref/libref_oracle.so that travelled with the snapshot is used."""
    subprocess.run(["make", "-s", "-C", HERE, "all"], check=True)
    if ref and os.path.isdir(REFERENCE_ROOT):
        subprocess.run(["make", "-s", "-C", HERE, "ref", "REF=" + REFERENCE_ROOT], check=True)


def have_ref():
    return os.path.exists(REF_SO)


def dense_to_csr(H):
    H = np.ascontiguousarray(H, dtype=np.uint8)
    m, n = H.shape
    row_ptr = np.zeros(m + 1, np.int32)
    cols = []
    for r in range(m):
        idx = np.flatnonzero(H[r])
        cols.append(idx)
        row_ptr[r + 1] = row_ptr[r] + len(idx)
    col_idx = np.concatenate(cols).astype(np.int32) if cols else np.zeros(0, np.int32)
    return row_ptr, np.ascontiguousarray(col_idx)


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = self.lib = C.CDLL(ORACLE_SO)
        L.orc_llr_variance.restype = C.c_double
        L.orc_llr_variance.argtypes = [C.c_double]
        L.orc_llr.restype = C.c_double
        L.orc_llr.argtypes = [C.c_double, C.c_double]
        L.orc_syndrome_ok.argtypes = [C.c_int, i32p, i32p, u8p]
        bp_args = [C.c_int, C.c_int, i32p, i32p, f64p, C.c_double, C.c_int, C.c_int, u8p,
                   C.POINTER(C.c_int), C.c_void_p]
        L.orc_bp_decode_fp80.argtypes = bp_args
        L.orc_bp_decode_fp64.argtypes = bp_args
        L.orc_admm_build.argtypes = [C.c_int, C.c_int, i32p, i32p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]
        L.orc_qpadmm_decode.argtypes = [C.c_int, C.c_int, i32p, i32p, f64p, C.c_double, C.c_double, C.c_double,
                                        C.c_int, C.c_double, u8p, C.POINTER(C.c_int), C.c_void_p]
        L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.orc_gauss_pair.argtypes = [u32p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_channel_frame.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_double, f64p]
        L.orc_info_bits.argtypes = [C.c_uint64, C.c_uint64, C.c_int, u8p]
        L.orc_encode.argtypes = [u8p, C.c_int, C.c_int, u8p, u8p]
        L.orc_experiment.argtypes = [C.c_int, C.c_int, C.c_int, i32p, i32p, C.c_double, C.c_int, C.c_int,
                                     C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint64, C.c_uint64,
                                     C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, u64p]

    # -- channel ------------------------------------------------------------
    def llr_variance(self, snr):
        return self.lib.orc_llr_variance(snr)

    def philox(self, ctr, key):
        out = np.zeros(4, np.uint32)
        self.lib.orc_philox4x32_10(np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), out)
        return out

    def gauss_pair(self, words):
        a, b = C.c_double(), C.c_double()
        self.lib.orc_gauss_pair(np.asarray(words, np.uint32), C.byref(a), C.byref(b))
        return a.value, b.value

    def channel(self, seed, frame_begin, frames, n, snr, codewords=None):
        """y for global frames [frame_begin, frame_begin + frames); codewords is
        None (all-zero) or a (frames, n) uint8 array."""
        sigma = float(np.sqrt(self.llr_variance(snr)))
        y = np.zeros((frames, n), np.float64)
        for f in range(frames):
            cw = None
            if codewords is not None:
                cw = np.ascontiguousarray(codewords[f], np.uint8).ctypes.data
            self.lib.orc_channel_frame(seed, frame_begin + f, n, cw, sigma, y[f])
        return y

    def info_bits(self, seed, frame, k):
        u = np.zeros(k, np.uint8)
        self.lib.orc_info_bits(seed, frame, k, u)
        return u

    def encode(self, G, u):
        G = np.ascontiguousarray(G, np.uint8)
        c = np.zeros(G.shape[1], np.uint8)
        self.lib.orc_encode(G, G.shape[0], G.shape[1], np.ascontiguousarray(u, np.uint8), c)
        return c

    # -- decoders -----------------------------------------------------------
    def bp_decode(self, H_csr, m, n, y, snr, max_iter, early_exit=True, precision="fp80"):
        row_ptr, col_idx = H_csr
        y = np.ascontiguousarray(y, np.float64).reshape(-1, n)
        B = y.shape[0]
        bits = np.zeros((B, n), np.uint8)
        ok = np.zeros(B, np.uint8)
        iters = np.zeros(B, np.int32)
        post = np.zeros((B, n), np.float64)
        fn = self.lib.orc_bp_decode_fp80 if precision == "fp80" else self.lib.orc_bp_decode_fp64
        it = C.c_int()
        for f in range(B):
            ok[f] = fn(m, n, row_ptr, col_idx, y[f], snr, max_iter, int(early_exit), bits[f], C.byref(it),
                       post[f].ctypes.data)
            iters[f] = it.value
        return bits, ok, iters, post

    def qpadmm_decode(self, H_csr, m, n, y, snr, alpha, mu, max_iter, eps_stop):
        row_ptr, col_idx = H_csr
        y = np.ascontiguousarray(y, np.float64).reshape(-1, n)
        B = y.shape[0]
        bits = np.zeros((B, n), np.uint8)
        ok = np.zeros(B, np.uint8)
        iters = np.zeros(B, np.int32)
        v = np.zeros((B, n), np.float64)
        it = C.c_int()
        for f in range(B):
            ok[f] = self.lib.orc_qpadmm_decode(m, n, row_ptr, col_idx, y[f], snr, alpha, mu, max_iter, eps_stop,
                                               bits[f], C.byref(it), v[f].ctypes.data)
            iters[f] = it.value
        return bits, ok, iters, v

    def admm_build(self, H_csr, m, n):
        row_ptr, col_idx = H_csr
        nv, R, nnz = C.c_int(), C.c_int(), C.c_int()
        self.lib.orc_admm_build(m, n, row_ptr, col_idx, C.byref(nv), C.byref(R), C.byref(nnz), None, None, None,
                                None, None)
        col_ptr = np.zeros(nv.value + 1, np.int32)
        col_row = np.zeros(max(nnz.value, 1), np.int32)
        col_cf = np.zeros(max(nnz.value, 1), np.float64)
        b = np.zeros(max(R.value, 1), np.float64)
        e = np.zeros(nv.value, np.float64)
        self.lib.orc_admm_build(m, n, row_ptr, col_idx, C.byref(nv), C.byref(R), C.byref(nnz),
                                col_ptr.ctypes.data, col_row.ctypes.data, col_cf.ctypes.data, b.ctypes.data,
                                e.ctypes.data)
        return dict(n_var=nv.value, R=R.value, nnz=nnz.value, col_ptr=col_ptr, col_row=col_row[:nnz.value],
                    col_cf=col_cf[:nnz.value], b=b[:R.value], e=e)

    def experiment(self, algo, H_csr, m, n, snr, max_iter, seed, frame_begin, count, early_exit=True, alpha=0.0,
                   mu=0.0, eps_stop=0.0, G=None, words=None):
        row_ptr, col_idx = H_csr
        cnt = np.zeros(len(CNT_NAMES), np.uint64)
        Gp, k = None, 0
        if G is not None:
            G = np.ascontiguousarray(G, np.uint8)
            Gp, k = G.ctypes.data, G.shape[0]
        wp, nw = None, 0
        if words is not None:
            words = np.ascontiguousarray(words, np.uint8)
            wp, nw = words.ctypes.data, words.shape[0]
        self.lib.orc_experiment(0 if algo == "bp" else 1, m, n, row_ptr, col_idx, snr, max_iter, int(early_exit),
                                alpha, mu, eps_stop, seed, frame_begin, count, Gp, k, wp, nw, cnt)
        return dict(zip(CNT_NAMES, (int(x) for x in cnt)))


class Ref:
    """The unmodified reference, via oracle/ref_harness.cpp."""

    def __init__(self):
        if not have_ref():
            build(ref=True)
        if not have_ref():
            raise FileNotFoundError(REF_SO)
        L = self.lib = C.CDLL(REF_SO)
        L.ref_llr_variance.restype = C.c_double
        L.ref_llr_variance.argtypes = [C.c_double]
        L.ref_llr.restype = C.c_double
        L.ref_llr.argtypes = [C.c_double, C.c_double]
        L.ref_read_pcm.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
        L.ref_save_matrix.argtypes = [u8p, C.c_int, C.c_int, C.c_char_p]
        L.ref_get_orthogonal.argtypes = [u8p, C.c_int, C.c_int, u8p]
        L.ref_is_codeword.argtypes = [u8p, C.c_int, C.c_int, u8p]
        L.ref_gen_random_codewords.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_uint32, u8p]
        L.ref_transmit.argtypes = [C.c_double, u8p, C.c_int, C.c_uint32, f64p]
        L.ref_bp_decode.argtypes = [u8p, C.c_int, C.c_int, f64p, C.c_double, C.c_int, u8p]
        L.ref_qpadmm_decode.argtypes = [u8p, C.c_int, C.c_int, f64p, C.c_double, C.c_double, C.c_double, C.c_int,
                                        C.c_double, u8p]
        L.ref_bp_decode_batch.restype = C.c_double
        L.ref_bp_decode_batch.argtypes = [u8p, C.c_int, C.c_int, f64p, C.c_int, C.c_double, C.c_int, u8p, u8p]
        L.ref_qpadmm_decode_batch.restype = C.c_double
        L.ref_qpadmm_decode_batch.argtypes = [u8p, C.c_int, C.c_int, f64p, C.c_int, C.c_double, C.c_double,
                                              C.c_double, C.c_int, C.c_double, u8p, u8p]
        L.ref_experiment.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, u8p, C.c_int, C.c_int,
                                     u8p, C.c_int, C.c_double, i64p]

    def read_pcm(self, path):
        r, c = C.c_int(), C.c_int()
        if self.lib.ref_read_pcm(path.encode(), C.byref(r), C.byref(c), None):
            raise IOError(path)
        H = np.zeros((r.value, c.value), np.uint8)
        self.lib.ref_read_pcm(path.encode(), C.byref(r), C.byref(c), H.ctypes.data)
        return H

    def save_matrix(self, H, path):
        H = np.ascontiguousarray(H, np.uint8)
        self.lib.ref_save_matrix(H, H.shape[0], H.shape[1], path.encode())

    def get_orthogonal(self, H):
        H = np.ascontiguousarray(H, np.uint8)
        m, n = H.shape
        G = np.zeros((n - m, n), np.uint8)
        ok = self.lib.ref_get_orthogonal(H, m, n, G)
        return (G, True) if ok else (None, False)

    def is_codeword(self, H, c):
        H = np.ascontiguousarray(H, np.uint8)
        return bool(self.lib.ref_is_codeword(H, H.shape[0], H.shape[1], np.ascontiguousarray(c, np.uint8)))

    def gen_random_codewords(self, G, count, seed):
        G = np.ascontiguousarray(G, np.uint8)
        out = np.zeros((count, G.shape[1]), np.uint8)
        self.lib.ref_gen_random_codewords(G, G.shape[0], G.shape[1], count, seed, out)
        return out

    def transmit(self, snr, codeword, seed):
        c = np.ascontiguousarray(codeword, np.uint8)
        y = np.zeros(len(c), np.float64)
        self.lib.ref_transmit(snr, c, len(c), seed, y)
        return y

    def bp_decode(self, H, y, snr, max_iter):
        H = np.ascontiguousarray(H, np.uint8)
        m, n = H.shape
        y = np.ascontiguousarray(y, np.float64).reshape(-1, n)
        bits = np.zeros((y.shape[0], n), np.uint8)
        ok = np.zeros(y.shape[0], np.uint8)
        secs = self.lib.ref_bp_decode_batch(H, m, n, y, y.shape[0], snr, max_iter, bits, ok)
        return bits, ok, secs

    def qpadmm_decode(self, H, y, snr, alpha, mu, max_iter, eps_stop):
        H = np.ascontiguousarray(H, np.uint8)
        m, n = H.shape
        y = np.ascontiguousarray(y, np.float64).reshape(-1, n)
        bits = np.zeros((y.shape[0], n), np.uint8)
        ok = np.zeros(y.shape[0], np.uint8)
        secs = self.lib.ref_qpadmm_decode_batch(H, m, n, y, y.shape[0], snr, alpha, mu, max_iter, eps_stop, bits,
                                                ok)
        return bits, ok, secs

    def experiment(self, algo, H, codewords, snr, max_iter, alpha=0.0, mu=0.0, eps_stop=0.0):
        H = np.ascontiguousarray(H, np.uint8)
        cw = np.ascontiguousarray(codewords, np.uint8)
        out = np.zeros(7, np.int64)
        self.lib.ref_experiment(0 if algo == "bp" else 1, max_iter, alpha, mu, eps_stop, H, H.shape[0], H.shape[1],
                                cw, cw.shape[0], snr, out)
        names = ["correct", "pseudo", "total", "sum_hamming", "sum_hamming_ok", "sum_hamming_wrong", "time_us"]
        return dict(zip(names, (int(x) for x in out)))
