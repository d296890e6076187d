"""CPU: the oracle (oracle/ldpc_oracle.c) against the committed golden vectors
produced by the unmodified reference (tests/golden/make_golden.py), and against
published known answers for Philox4x32-10."""
import os

import numpy as np
import pytest

from oracle.oracle import dense_to_csr
from tests.helpers import GOLDEN, load_rows


@pytest.mark.parametrize("name", ["optimalH", "H05"])
def test_oracle_reproduces_reference_decoders(oracle, name):
    g = np.load(os.path.join(GOLDEN, "ref_%s.npz" % name))
    H = load_rows(name)
    m, n = H.shape
    csr = dense_to_csr(H)
    alpha, mu = g["alpha_mu"]
    for si, snr in enumerate(g["snrs"]):
        y = g["y_%d" % si]
        bits, ok, iters, post = oracle.bp_decode(csr, m, n, y, snr, 100)
        assert (ok == g["bp_ok_%d" % si]).all()
        assert (bits == g["bp_bits_%d" % si]).all()       # failed frames: zero rows on both sides
        assert ((post <= 0) == bits.astype(bool))[ok.astype(bool)].all()
        bits, ok, iters, v = oracle.qpadmm_decode(csr, m, n, y, snr, alpha, mu, 1000, 1e-5)
        assert (ok == g["admm_ok_%d" % si]).all()
        assert (bits == g["admm_bits_%d" % si]).all()
        assert ((v > 0.5) == bits.astype(bool)).all()


@pytest.mark.parametrize("name", ["optimalH", "H05"])
def test_oracle_verdict_counters_match_reference_harness(oracle, name):
    """exp() verdict + HammingDistanceTracker (experiment.h:33-46, 109-120)."""
    import ctypes as C
    g = np.load(os.path.join(GOLDEN, "ref_%s.npz" % name))
    H = load_rows(name)
    m, n = H.shape
    row_ptr, col_idx = dense_to_csr(H)
    alpha, mu = g["alpha_mu"]
    y, cw = g["exp_y"], g["exp_codewords"]
    keys = list(g["exp_keys"])
    for algo in ("bp", "qpadmm"):
        want = dict(zip(keys, g["exp_" + algo]))
        if algo == "bp":
            bits, ok, iters, _ = oracle.bp_decode((row_ptr, col_idx), m, n, y, -3.0, 100)
            has_bits = ok
        else:
            bits, ok, iters, _ = oracle.qpadmm_decode((row_ptr, col_idx), m, n, y, -3.0, alpha, mu, 1000, 1e-5)
            has_bits = np.ones_like(ok)
        cnt = np.zeros(10, np.uint64)
        fn = oracle.lib.orc_account_frame
        fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                       C.c_void_p, C.c_int, C.c_void_p]
        for f in range(len(y)):
            fn(m, n, row_ptr.ctypes.data, col_idx.ctypes.data, cw[f].ctypes.data, y[f].ctypes.data, int(ok[f]),
               int(has_bits[f]), bits[f].ctypes.data, int(iters[f]), cnt.ctypes.data)
        got = dict(total=cnt[0], correct=cnt[1], pseudo=cnt[2], sum_hamming=cnt[5], sum_hamming_ok=cnt[6],
                   sum_hamming_wrong=cnt[7])
        for k in keys:
            assert int(got[k]) == int(want[k]), (algo, k)


def test_philox_known_answers(oracle):
    """Random123 kat_vectors for philox4x32_10."""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert list(oracle.philox(ctr, key)) == want


def test_gaussian_transform_accuracy_and_moments(oracle):
    rng = np.random.default_rng(5)
    for _ in range(500):
        w = rng.integers(0, 2 ** 32, 4, dtype=np.uint64).astype(np.uint32)
        z0, z1 = oracle.gauss_pair(w)
        u1 = (((int(w[0]) << 20) | (int(w[1]) >> 12)) + 0.5) * 2.0 ** -52
        u2 = (((int(w[2]) << 20) | (int(w[3]) >> 12)) + 0.5) * 2.0 ** -52
        r = np.sqrt(-2 * np.log(u1))
        assert abs(z0 - r * np.cos(2 * np.pi * u2)) < 1e-13
        assert abs(z1 - r * np.sin(2 * np.pi * u2)) < 1e-13
    y = oracle.channel(239239239, 0, 2000, 280, -3.0)
    var = oracle.llr_variance(-3.0)
    assert abs(y.mean() - 1.0) < 4 * np.sqrt(var / y.size)
    assert abs(y.var() / var - 1.0) < 0.01
    # frames are addressed by global index: a shifted window replays the same rows
    y2 = oracle.channel(239239239, 100, 50, 280, -3.0)
    assert (y2 == y[100:150]).all()


def test_generator_fixture_and_info_bits(oracle):
    g = np.load(os.path.join(GOLDEN, "ref_generator_optimalH.npz"))
    n = int(g["n"])
    G = np.unpackbits(g["G"], axis=1)[:, :n]
    H = load_rows("optimalH")
    assert ((G.astype(int) @ H.T.astype(int)) % 2 == 0).all()
    row_ptr, col_idx = dense_to_csr(H)
    for w in g["first_words"]:
        assert oracle.lib.orc_syndrome_ok(H.shape[0], row_ptr, col_idx, np.ascontiguousarray(w))
    u = oracle.info_bits(1234, 77, G.shape[0])
    c = oracle.encode(G, u)
    assert (c == (u.astype(int) @ G.astype(int)) % 2).all()
    assert 20 < u.sum() < 100


def test_oracle_posterior_against_60_digit_arithmetic(oracle):
    """the fp80 oracle's posterior LLRs (the reference's VNode::estimate, bp.h:85-90, instrumented) against 60-digit
    arithmetic running the reference's flooding schedule for the same number of iterations: within 1e-4 relative --
    the yardstick the GPU posterior tests use is itself pinned"""
    import mpmath as mp
    from oracle.oracle import dense_to_csr
    from tests.helpers import load_rows
    from tests.posterior_accuracy import mp_bp
    mp.mp.dps = 60
    H = load_rows("optimalH")
    m, n = H.shape
    for snr, frames in ((-1.0, 2), (3.0, 1)):
        y = oracle.channel(239239239, 41000, frames, n, snr)
        bits, ok, iters, post = oracle.bp_decode(dense_to_csr(H), m, n, y, snr, 100)
        assert (ok == 1).all()
        for f in range(frames):
            truth = np.array([float(t) for t in mp_bp(H, y[f], snr, int(iters[f]), mp)])
            fin = np.isfinite(post[f])
            assert fin.mean() > 0.9
            assert np.max(np.abs(post[f][fin] - truth[fin]) / np.abs(truth[fin])) < 1e-4
            assert ((truth <= 0) == (bits[f] == 1)).all()
