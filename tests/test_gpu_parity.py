"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle
on identical inputs, against the reference's golden vectors, and -- at sizes the
oracle cannot follow -- through size-independent properties.

Bars (BASELINE.json north_star):
  * channel samples, QP-ADMM hard decisions / iteration counts / v: bit-exact
  * BP: hard decisions, success flag and converging iteration identical per frame
    (>= 99.99 % required; every mismatch is printed), posterior LLR within 1e-4 relative
"""
import os

import numpy as np
import pytest

from oracle.oracle import dense_to_csr
from tests.helpers import GOLDEN, load_rows, small_irregular_code, wilson_interval

pytestmark = pytest.mark.gpu

SEED = 239239239
ADMM = {"optimalH": (1.2, 0.55), "H05": (1.95, 0.5), "reg_3_6_1008": (1.2, 0.55)}


@pytest.fixture(scope="module")
def codes(gpu_lib):
    out = {}
    for name in ("optimalH", "H05", "reg_3_6_1008"):
        H = load_rows(name)
        out[name] = (H, gpu_lib.Code(H=H), dense_to_csr(H))
    return out


def test_code_info_matches_survey_table(codes):
    want = {"optimalH": (900, 580, 700, 2320, 6960, 4), "H05": (860, 540, 660, 2160, 6480, 4),
            "reg_3_6_1008": (3024, 2016, 2520, 8064, 24192, 8)}
    for name, (E, T, nv, R, nnz, emin) in want.items():
        info = codes[name][1].info
        assert (info["edges"], info["admm_blocks"], info["admm_n_var"], info["admm_rows"], info["admm_nnz"],
                info["admm_e_min"]) == (E, T, nv, R, nnz, emin)


def test_channel_bit_exact_with_oracle(codes, oracle):
    for name in ("optimalH", "reg_3_6_1008"):
        H, code, _ = codes[name]
        n = H.shape[1]
        rng = np.random.default_rng(1)
        cw = rng.integers(0, 2, (64, n)).astype(np.uint8)      # any bits: the channel does not care
        for snr, begin in ((-3.0, 0), (0.5, (1 << 33) + 12345)):
            y_gpu = code.channel(SEED, begin, 64, snr, cw)
            y_cpu = oracle.channel(SEED, begin, 64, n, snr, cw)
            assert y_gpu.tobytes() == y_cpu.tobytes()
        y0 = code.channel(SEED + 1, 7, 16, -1.0)
        assert y0.tobytes() == oracle.channel(SEED + 1, 7, 16, n, -1.0).tobytes()


@pytest.mark.parametrize("name,snrs,frames,max_iter", [
    ("optimalH", (-4.0, -3.0, -1.0), 96, 1000),
    ("H05", (-3.0, 0.0), 64, 1000),
    ("reg_3_6_1008", (-2.0, 0.0), 24, 300),
])
def test_qpadmm_bit_exact(codes, oracle, name, snrs, frames, max_iter):
    H, code, csr = codes[name]
    m, n = H.shape
    alpha, mu = ADMM[name]
    for snr in snrs:
        y = code.channel(SEED, 1000, frames, snr)
        gb, gok, git, gv = code.qpadmm_decode(y, snr, alpha, mu, max_iter, 1e-5)
        # every code of BASELINE.json is inside the check-centric kernel's range: a silent fall-back is a failure
        assert _lib(codes).last_qpadmm_kernel() == 1, "the block-per-lane kernel served %s" % name
        ob, ook, oit, ov = oracle.qpadmm_decode(csr, m, n, y, snr, alpha, mu, max_iter, 1e-5)
        bad = np.flatnonzero((gb != ob).any(1) | (git != oit) | (gok != ook))
        for f in bad:
            print("QP-ADMM mismatch %s snr=%g frame=%d iters gpu/cpu=%d/%d" % (name, snr, f, git[f], oit[f]))
        assert len(bad) == 0
        assert (gv == ov).all(), "v differs by up to %g" % np.abs(gv - ov).max()
    # fixed-iteration mode (eps_stop = 0) runs every iteration
    y = code.channel(SEED, 5, 8, 0.0)
    gb, gok, git, gv = code.qpadmm_decode(y, 0.0, alpha, mu, 50, 0.0)
    ob, ook, oit, ov = oracle.qpadmm_decode(csr, m, n, y, 0.0, alpha, mu, 50, 0.0)
    assert (git == 50).all() and (oit == 50).all() and (gv == ov).all() and (gb == ob).all()


@pytest.mark.parametrize("name,snrs,frames", [
    ("optimalH", (-4.0, -3.0, -2.0, 0.0), 128),
    ("H05", (-3.0, -1.0), 96),
    ("reg_3_6_1008", (-1.0, 1.0), 24),
])
def test_bp_matches_fp80_oracle(codes, oracle, name, snrs, frames):
    H, code, csr = codes[name]
    m, n = H.shape
    total = mism = 0
    for snr in snrs:
        y = code.channel(SEED, 2000, frames, snr)
        gb, gok, git, gpost = code.bp_decode(y, snr, 100)
        # every code of BASELINE.json is served by the likelihood-ratio kernel: a silent fall-back is a failure
        assert _lib(codes).last_bp_kernel() == 1, "the log-domain kernel served %s" % name
        ob, ook, oit, opost = oracle.bp_decode(csr, m, n, y, snr, 100)
        bad = np.flatnonzero((gb != ob).any(1) | (gok != ook) | (git != oit))
        for f in bad:
            print("BP mismatch %s snr=%g frame=%d ok gpu/cpu=%d/%d iters=%d/%d" %
                  (name, snr, f, gok[f], ook[f], git[f], oit[f]))
        total += frames
        mism += len(bad)
        conv = (gok == 1) & (ook == 1)
        rel = np.abs(gpost[conv] - opost[conv]) / np.maximum(np.abs(opost[conv]), 1e-300)
        assert rel.size == 0 or rel.max() < 1e-4, "posterior LLR off by %g relative" % rel.max()
        assert (gb[gok == 0] == 0).all()      # failed frames return no bits (bp.h:198)
    assert mism == 0, "%d of %d frames differ" % (mism, total)


@pytest.mark.parametrize("slots", ["2", "4", "8", "16", "log"])
def test_bp_multi_slot_kernels(codes, oracle, slots, monkeypatch):
    """F frames per CTA with slot refill: same per-frame results as the oracle, whatever the schedule
    ("log" = the log-domain kernel that serves codes with very large node degrees)"""
    if slots == "log":
        monkeypatch.setenv("LDPC_BP_KERNEL", "log")
        monkeypatch.setenv("LDPC_BP_F", "4")
    else:
        monkeypatch.setenv("LDPC_BP_F", slots)
    for name, frames, snr in (("optimalH", 301, -3.0), ("H05", 150, -2.0), ("reg_3_6_1008", 21, -1.0)):
        H, code, csr = codes[name]
        m, n = H.shape
        y = code.channel(SEED, 4000, frames, snr)
        gb, gok, git, gpost = code.bp_decode(y, snr, 100)
        assert _lib(codes).last_bp_kernel() == (2 if slots == "log" else 1)
        ob, ook, oit, opost = oracle.bp_decode(csr, m, n, y, snr, 100)
        bad = np.flatnonzero((gb != ob).any(1) | (gok != ook) | (git != oit))
        for f in bad:
            print("BP(F=%s) mismatch %s frame=%d ok %d/%d iters %d/%d" % (slots, name, f, gok[f], ook[f], git[f], oit[f]))
        assert len(bad) == 0
        conv = gok == 1
        assert np.allclose(gpost[conv], opost[conv], rtol=1e-4, atol=0)
    # experiment counters are schedule independent too
    H, code, csr = codes["optimalH"]
    bp = _lib(codes).BeliefPropagationDecoder(100)
    got = code.experiment(bp, -3.0, SEED, 0, 90)
    want = oracle.experiment("bp", csr, H.shape[0], H.shape[1], -3.0, 100, SEED, 0, 90)
    assert all(got[k] == want[k] for k in want)


def _lib(codes):
    import ldpc_b200
    return ldpc_b200


def test_bp_math_kernels(gpu_lib):
    """exp(-a) and log(ev/od) of csrc/bpmath.cuh against long double"""
    rng = np.random.default_rng(11)
    a = np.concatenate([rng.uniform(0, 60, 200000), rng.uniform(0, 1e-3, 1000), rng.uniform(0, 699, 20000),
                        [0.0, 700.0, 5000.0, np.inf]])
    od = np.exp(-rng.uniform(0, 80, a.size)) * rng.integers(0, 2, a.size)
    ev = 1 + rng.uniform(0, 30, a.size) * rng.integers(0, 2, a.size) + od * rng.uniform(0, 1, a.size)
    ev = np.maximum(ev, od)
    got_exp, got_log = gpu_lib.debug_bpmath(a, ev, od)
    fin = a <= 699
    ref = np.exp(-a[fin].astype(np.longdouble))
    assert float(np.max(np.abs(got_exp[fin] - ref) / ref)) < 1e-15
    assert (got_exp[~fin] > 0).all() and (got_exp[~fin] < 1e-300).all()       # capped at exp(-700), never 0 / denormal
    pos = od > 0
    ref = np.log(ev[pos].astype(np.longdouble) / od[pos].astype(np.longdouble))
    assert float(np.max(np.abs(got_log[pos] - ref))) < 5e-14
    assert (got_log[~pos] > 700).all() and np.isfinite(got_log).all()          # od = 0: capped, not inf


def test_bp_fixed_iteration_mode(codes, oracle):
    H, code, csr = codes["optimalH"]
    m, n = H.shape
    y = code.channel(SEED, 0, 32, -1.0)
    gb, gok, git, gpost = code.bp_decode(y, -1.0, 20, early_exit=False)
    ob, ook, oit, opost = oracle.bp_decode(csr, m, n, y, -1.0, 20, early_exit=False)
    assert (git == 20).all() and (gok == ook).all() and (gb == ob).all()
    # Without the syndrome exit converged frames keep iterating and their LLRs grow without bound.  The
    # reference's long-double phi(x) = -log(tanh(x/2)) loses relative accuracy from |x| ~ 30 and saturates
    # to +inf near 45 (tanhl rounds to 1); the kernel's exp-domain form does not.  Compare where the oracle
    # is still accurate, and require the saturated entries to be large and of the right sign on the GPU.
    ok = (gok == 1)[:, None] & np.ones_like(gpost, bool)
    small = ok & (np.abs(opost) < 20)
    assert np.allclose(gpost[small], opost[small], rtol=1e-4, atol=0)
    big = ok & ~small
    assert (np.sign(gpost[big]) == np.sign(opost[big])).all() and (np.abs(gpost[big]) > 19.9).all()


@pytest.mark.parametrize("name", ["optimalH", "H05"])
def test_reference_golden_vectors(codes, name):
    """outputs of the unmodified reference on its own mt19937 channel words"""
    g = np.load(os.path.join(GOLDEN, "ref_%s.npz" % name))
    H, code, _ = codes[name]
    alpha, mu = g["alpha_mu"]
    for si, snr in enumerate(g["snrs"]):
        y = g["y_%d" % si]
        bits, ok, _, _ = code.bp_decode(y, float(snr), 100)
        assert (ok == g["bp_ok_%d" % si]).all() and (bits == g["bp_bits_%d" % si]).all()
        bits, ok, _, _ = code.qpadmm_decode(y, float(snr), alpha, mu, 1000, 1e-5)
        assert (ok == g["admm_ok_%d" % si]).all() and (bits == g["admm_bits_%d" % si]).all()


def test_edge_cases(gpu_lib, codes, oracle):
    H, code, csr = codes["optimalH"]
    m, n = H.shape
    # empty batch, single frame, ragged batch sizes
    b, ok, it, s = code.bp_decode(np.zeros((0, n)), 0.0, 10)
    assert b.shape == (0, n) and ok.shape == (0,)
    b, ok, it, s = code.qpadmm_decode(np.zeros((0, n)), 0.0, 1.2, 0.55, 10, 1e-5)
    assert b.shape == (0, n)
    for frames in (1, 3, 149, 1000):
        y = code.channel(SEED, 99, frames, 0.0)
        b, ok, it, _ = code.bp_decode(y, 0.0, 100)
        assert ok.all() and (b == 0).all()           # all-zero codeword at 0 dB
        b2, ok2, it2, _ = code.qpadmm_decode(y, 0.0, 1.2, 0.55, 2000, 1e-5)
        assert ok2.all() and (b2 == 0).all()
        if frames == 149:
            ob, ook, oit, _ = oracle.bp_decode(csr, m, n, y[:20], 0.0, 100)
            assert (oit == it[:20]).all()
    # infeasible parameters: min(e) mu <= alpha -> {zeros, false}, qp_admm.h:108-114
    y = code.channel(SEED, 0, 5, 0.0)
    b, ok, it, v = code.qpadmm_decode(y, 0.0, 3.0, 0.5, 100, 1e-5)
    assert (ok == 0).all() and (b == 0).all() and (it == 0).all() and (v == 0).all()
    # max_iter = 0: BP fails at once, QP-ADMM returns the initial v = (q > 0)
    b, ok, it, _ = code.bp_decode(y, 0.0, 0)
    assert (ok == 0).all() and (b == 0).all() and (it == 0).all()
    ob, ook, oit, ov = oracle.qpadmm_decode(csr, m, n, y, 0.0, 1.2, 0.55, 0, 1e-5)
    b, ok, it, v = code.qpadmm_decode(y, 0.0, 1.2, 0.55, 0, 1e-5)
    assert (b == ob).all() and (ok == ook).all() and (it == 0).all()


def test_irregular_code_special_blocks(gpu_lib, oracle):
    """checks of degree 0/1/2 and isolated variables (qp_admm.h:67-83; bp.h handles them implicitly)"""
    H = small_irregular_code()
    m, n = H.shape
    code = gpu_lib.Code(H=H)
    csr = dense_to_csr(H)
    rng = np.random.default_rng(3)
    y = 1.0 + 0.8 * rng.standard_normal((40, n))
    gb, gok, git, gv = code.qpadmm_decode(y, 1.0, -0.3, 0.6, 300, 1e-5)
    ob, ook, oit, ov = oracle.qpadmm_decode(csr, m, n, y, 1.0, -0.3, 0.6, 300, 1e-5)
    assert (gb == ob).all() and (gok == ook).all() and (git == oit).all() and (gv == ov).all()
    gb, gok, git, gv = code.qpadmm_decode(y, 1.0, 0.5, 0.6, 300, 1e-5)     # e_min = 0 -> infeasible
    assert (gok == 0).all() and (gb == 0).all()
    gb, gok, git, gpost = code.bp_decode(y, 1.0, 30)
    ob, ook, oit, opost = oracle.bp_decode(csr, m, n, y, 1.0, 30)
    assert (gb == ob).all() and (gok == ook).all() and (git == oit).all()
    fin = np.isfinite(opost) & (ook == 1)[:, None]
    assert np.allclose(gpost[fin], opost[fin], rtol=1e-4, atol=0)
    # degree-1 checks pin a bit: the reference's LLR is +inf, the kernels cap message magnitudes (100 in the
    # likelihood-ratio kernel, ~700 in the log-domain one)
    pinned = np.isinf(opost) & (ook == 1)[:, None]
    assert pinned.any() and (gpost[pinned] > 90).all()


def test_experiment_counters_match_oracle(gpu_lib, codes, oracle):
    """device-side codeword + AWGN + decode + verdict + counting == the oracle's replay"""
    H, code, csr = codes["optimalH"]
    m, n = H.shape
    g = np.load(os.path.join(GOLDEN, "ref_generator_optimalH.npz"))
    G = np.unpackbits(g["G"], axis=1)[:, :n]
    code.set_generator(G)
    # the generated codewords themselves
    cw_gpu = code.generator_codewords(SEED, 10, 16)
    for f in range(16):
        assert (cw_gpu[f] == oracle.encode(G, oracle.info_bits(SEED, 10 + f, G.shape[0]))).all()
    bp = gpu_lib.BeliefPropagationDecoder(100)
    admm = gpu_lib.QPADMMDecoder(1.2, 0.55, 1000, 1e-5)
    keys = ["total", "correct", "pseudo", "decoder_fail", "bit_errors", "sum_hamming", "sum_hamming_ok",
            "sum_hamming_wrong", "sum_iters", "frames_with_bits"]
    for snr in (-3.0, -1.5):
        for source, kw in ((gpu_lib.CW_GENERATOR, dict(G=G)), (gpu_lib.CW_ZERO, {}),
                           (gpu_lib.CW_TABLE, dict(words=cw_gpu))):
            for dec, algo in ((bp, "bp"), (admm, "qpadmm")):
                got = code.experiment(dec, snr, SEED, 500, 60, source, kw.get("words"))
                want = oracle.experiment(algo, csr, m, n, snr, dec.max_iter, SEED, 500, 60, alpha=1.2, mu=0.55,
                                         eps_stop=1e-5, **kw)
                assert {k: got[k] for k in keys} == want, (snr, source, algo)
    # sharding invariance: counters over [0, 300) == sum over three disjoint shards (any GPU count)
    whole = code.experiment(admm, -2.0, SEED, 0, 300, gpu_lib.CW_GENERATOR)
    parts = [code.experiment(admm, -2.0, SEED, b, c, gpu_lib.CW_GENERATOR) for b, c in ((0, 100), (100, 37), (137, 163))]
    for k in keys:
        assert whole[k] == sum(p[k] for p in parts)


def test_full_size_properties(gpu_lib, codes):
    """sizes the CPU oracle cannot follow: round trips and statistics"""
    H, code, _ = codes["optimalH"]
    n = H.shape[1]
    g = np.load(os.path.join(GOLDEN, "ref_generator_optimalH.npz"))
    G = np.unpackbits(g["G"], axis=1)[:, :n]
    code.set_generator(G)
    bp = gpu_lib.BeliefPropagationDecoder(100)
    admm = gpu_lib.QPADMMDecoder(1.2, 0.55, 10000, 1e-5)
    # encode -> (almost noiseless) channel -> decode returns the transmitted word, every frame
    for dec in (bp, admm):
        r = code.experiment(dec, 6.0, SEED, 0, 200000, gpu_lib.CW_GENERATOR)
        assert r["total"] == 200000 and r["correct"] == 200000 and r["bit_errors"] == 0 and r["pseudo"] == 0
    # FER at -3 dB agrees with the race-free reference numbers (BP 529/1000, QP-ADMM 744/1000
    # correct on mt19937 noise) within binomial confidence intervals
    for dec, ref_correct in ((bp, 529), (admm, 744)):
        r = code.experiment(dec, -3.0, SEED, 0, 20000, gpu_lib.CW_GENERATOR)
        lo, hi = wilson_interval(ref_correct, 1000, z=3.3)
        p = r["correct"] / r["total"]
        assert lo <= p <= hi, (dec.name(), p, lo, hi)
        # channel symmetry: the all-zero codeword sees the same FER
        r0 = code.experiment(dec, -3.0, SEED + 1, 0, 20000, gpu_lib.CW_ZERO)
        lo, hi = wilson_interval(r["correct"], r["total"], z=4.0)
        assert lo - 0.01 <= r0["correct"] / r0["total"] <= hi + 0.01
    # mean channel Hamming distance ~ n * Q(1/sigma)
    from math import erfc, sqrt
    sigma = sqrt(10 ** 0.3 / 2)
    expect = n * 0.5 * erfc(1 / sigma / sqrt(2))
    assert abs(r["sum_hamming"] / r["total"] - expect) < 0.2


@pytest.mark.parametrize("shape", [("1", "2"), ("1", "8"), ("2", "4"), ("2", "5"), ("4", "8")])
def test_qpadmm_launch_shapes(codes, oracle, shape, monkeypatch):
    """frames per CTA x blocks per lane: every shape is bit-identical to the oracle, whatever the schedule"""
    monkeypatch.setenv("LDPC_ADMM_F", shape[0])
    monkeypatch.setenv("LDPC_ADMM_KB", shape[1])
    for name, frames, snr, iters in (("optimalH", 75, -3.0, 600), ("H05", 33, -2.0, 400), ("reg_3_6_1008", 9, -1.0, 200)):
        H, code, csr = codes[name]
        m, n = H.shape
        alpha, mu = ADMM[name]
        y = code.channel(SEED, 7000, frames, snr)
        gb, gok, git, gv = code.qpadmm_decode(y, snr, alpha, mu, iters, 1e-5)
        ob, ook, oit, ov = oracle.qpadmm_decode(csr, m, n, y, snr, alpha, mu, iters, 1e-5)
        assert (git == oit).all() and (gb == ob).all() and (gok == ook).all() and (gv == ov).all(), (name, shape)
    H, code, csr = codes["optimalH"]
    import ldpc_b200
    dec = ldpc_b200.QPADMMDecoder(1.2, 0.55, 500, 1e-5)
    got = code.experiment(dec, -2.5, SEED, 0, 70)
    want = oracle.experiment("qpadmm", csr, H.shape[0], H.shape[1], -2.5, 500, SEED, 0, 70, alpha=1.2, mu=0.55, eps_stop=1e-5)
    assert all(got[k] == want[k] for k in want)


def test_admm_layout_removes_bank_conflicts(codes):
    """the graph compiler's rank assignment (csrc/admm_layout.cu) must not be worse than the natural order"""
    for name in ("optimalH", "H05", "reg_3_6_1008"):
        info = codes[name][1].info
        print(name, "replayed wavefronts per iteration: natural", info["admm_conflicts_natural"], "laid out",
              info["admm_conflicts_laid_out"])
        assert info["admm_conflicts_laid_out"] <= info["admm_conflicts_natural"]


@pytest.mark.gpu
def test_qpadmm_grid_equals_point_by_point(codes, oracle):
    """qpadmm_params.cpp:51-67 batched into one launch: every (alpha, mu) pair gets exactly the counters of its own
    ldpc_experiment_run (same frames for every pair), including infeasible pairs (min(e) * mu <= alpha)"""
    L = _lib(codes)
    H, code, csr = codes["optimalH"]
    m, n = H.shape
    alphas = np.array([0.0, 1.2, 1.95, 3.0, 0.3, 2.25, 1.5], np.float64)
    mus = np.array([0.5, 0.55, 0.5, 0.7, 0.0, 3.0, 0.4], np.float64)    # pairs 3 and 4 are infeasible (e_min = 4)
    frames, snr, iters, eps = 257, -3.0, 300, 1e-5
    grid, secs = code.qpadmm_grid(alphas, mus, snr, iters, eps, SEED, 11, frames)
    assert secs > 0
    for a, mu_, got in zip(alphas, mus, grid):
        want = code.experiment(L.QPADMMDecoder(a, mu_, iters, eps), snr, SEED, 11, frames)
        for k in got:
            assert got[k] == want[k], (a, mu_, k, got[k], want[k])
    # and one pair against the CPU oracle's replay
    want = oracle.experiment("qpadmm", csr, m, n, snr, iters, SEED, 11, frames, alpha=1.2, mu=0.55, eps_stop=eps)
    assert all(grid[1][k] == want[k] for k in want)


@pytest.mark.gpu
def test_bp_layout_for_8_frames_per_cta_is_nearly_conflict_free(codes):
    """odd class strides + balanced edge 2-colouring: paired lane groups use opposite halves of a 128-byte line"""
    for name in ("optimalH", "H05", "reg_3_6_1008"):
        H, code, csr = codes[name]
        st = code.bp_layout(8)
        print(name, st)
        assert st["pad_even"] == 1 and st["slots"] >= int(H.sum())
        assert st["clash_c"] == 0
        assert st["clash_v"] <= 0.12 * st["pairs_v"], st
        if name != "reg_3_6_1008":     # (16 frames of the n = 1008 code exceed the kernel's shared-memory offsets)
            assert code.bp_layout(16)["slots"] == int(H.sum())        # 16 frames fill a line: no padding


@pytest.mark.gpu
@pytest.mark.parametrize("max_row_deg", [10, 12, 16])
def test_wide_check_degrees(gpu_lib, oracle, max_row_deg):
    """checks of degree 9..12 take the check-centric QP-ADMM kernel's wide variants and the generic-degree paths of the
    likelihood-ratio BP kernel; degree 16 takes the fallbacks (block-per-lane QP-ADMM, log-domain BP)"""
    rng = np.random.default_rng(100 + max_row_deg)
    m, n = 24, 60
    H = np.zeros((m, n), np.uint8)
    degs = [3, 4, 5, 6, 7, 8, 9, max_row_deg] * 3
    for r, d in enumerate(degs):
        H[r, rng.choice(n, size=d, replace=False)] = 1
    for v in np.flatnonzero(H.sum(0) == 0):          # every variable in at least one check
        H[rng.integers(m), v] = 1
    code = gpu_lib.Code(H=H)
    csr = dense_to_csr(H)
    y = 1.0 + 0.75 * rng.standard_normal((64, n))
    snr = 0.0
    gb, gok, git, gv = code.qpadmm_decode(y, snr, 1.2, 0.55, 400, 1e-5)
    ob, ook, oit, ov = oracle.qpadmm_decode(csr, m, n, y, snr, 1.2, 0.55, 400, 1e-5)
    assert (gb == ob).all() and (gok == ook).all() and (git == oit).all() and (gv == ov).all()
    gb, gok, git, gpost = code.bp_decode(y, snr, 50)
    ob, ook, oit, opost = oracle.bp_decode(csr, m, n, y, snr, 50)
    assert (gb == ob).all() and (gok == ook).all() and (git == oit).all()
    fin = np.isfinite(opost) & (np.abs(opost) < 30) & (ook == 1)[:, None]
    assert np.allclose(gpost[fin], opost[fin], rtol=1e-4, atol=0)
    code.close()


@pytest.mark.gpu
def test_qpadmm_block_kernel_fallback(codes, oracle, monkeypatch):
    """the block-per-lane kernel (codes outside the check-centric kernel's range) stays bit-identical to the oracle"""
    monkeypatch.setenv("LDPC_ADMM_KERNEL", "block")
    for name, frames, snr, iters in (("optimalH", 75, -3.0, 600), ("reg_3_6_1008", 9, -1.0, 200)):
        H, code, csr = codes[name]
        m, n = H.shape
        alpha, mu = ADMM[name]
        y = code.channel(SEED, 9000, frames, snr)
        gb, gok, git, gv = code.qpadmm_decode(y, snr, alpha, mu, iters, 1e-5)
        assert _lib(codes).last_qpadmm_kernel() == 2
        ob, ook, oit, ov = oracle.qpadmm_decode(csr, m, n, y, snr, alpha, mu, iters, 1e-5)
        assert (git == oit).all() and (gb == ob).all() and (gok == ook).all() and (gv == ov).all(), name


@pytest.mark.gpu
def test_bp_posterior_against_60_digit_arithmetic(codes, oracle):
    """north_star: output LLRs within 1e-4 relative.  The reference's long-double phi form is itself only accurate to
    ~1e-6 for |LLR| > 30 and saturates to inf near 45 (tanhl(x/2) rounds to 1), so the judge of both is 60-digit
    arithmetic (mpmath) running the reference's flooding schedule (bp.h:155-199) for the same number of iterations on
    the same channel samples: the CUDA posterior must be within 1e-4 (it is within 1e-10) of the TRUE value at every
    magnitude, and the fp80 oracle must be within 1e-4 of it wherever it is finite."""
    import mpmath as mp
    from tests.posterior_accuracy import mp_bp
    mp.mp.dps = 60
    H, code, csr = codes["optimalH"]
    m, n = H.shape
    worst_gpu = worst_orc = 0.0
    big = 0
    for snr, frames in ((-1.0, 3), (2.0, 2), (5.0, 2)):
        y = code.channel(SEED, 31000, frames, snr)
        gb, gok, git, gpost = code.bp_decode(y, snr, 100)
        ob, ook, oit, opost = oracle.bp_decode(csr, m, n, y, snr, 100)
        assert (gok == 1).all() and (ook == 1).all() and (git == oit).all() and (gb == ob).all()
        for f in range(frames):
            truth = np.array([float(t) for t in mp_bp(H, y[f], snr, int(git[f]), mp)])
            assert ((truth <= 0) == (gb[f] == 1)).all()
            worst_gpu = max(worst_gpu, float(np.max(np.abs(gpost[f] - truth) / np.abs(truth))))
            fin = np.isfinite(opost[f])
            worst_orc = max(worst_orc, float(np.max(np.abs(opost[f][fin] - truth[fin]) / np.abs(truth[fin]))))
            big += int((np.abs(truth) > 30).sum())
    print("posterior LLR vs 60-digit arithmetic: CUDA max rel err %.3g, fp80 oracle %.3g (%d values beyond 30)" %
          (worst_gpu, worst_orc, big))
    assert big > 0
    assert worst_gpu < 1e-9
    assert worst_orc < 1e-4


@pytest.mark.gpu
def test_bp_fixed_iteration_mode_against_60_digit_arithmetic(codes):
    """Fixed-iteration mode at high SNR, where the reference's messages run into inf and the kernel clamps likelihood
    ratios to exp(+-100) (the one deviation from SURVEY.md 7.3-2): against 60-digit arithmetic the hard
    decisions are identical, every posterior below the clamp is within 1e-4 relative, and the others are large."""
    import mpmath as mp
    from tests.posterior_accuracy import mp_bp
    mp.mp.dps = 60
    H, code, _ = codes["optimalH"]
    clamped = 0
    for snr, iters in ((1.0, 6), (4.0, 4)):
        y = code.channel(SEED, 32000, 2, snr)
        gb, gok, git, gpost = code.bp_decode(y, snr, iters, early_exit=False)
        assert (git == iters).all() and (gok == 1).all()
        for f in range(2):
            truth = np.array([float(t) for t in mp_bp(H, y[f], snr, iters, mp)])
            assert ((truth <= 0) == (gb[f] == 1)).all()
            low = np.abs(truth) < 90
            assert np.allclose(gpost[f][low], truth[low], rtol=1e-4, atol=0)
            assert (np.abs(gpost[f][~low]) > 89.9).all() and (np.sign(gpost[f][~low]) == np.sign(truth[~low])).all()
            clamped += int((~low).sum())
    print("fixed-iteration mode: %d posteriors beyond the clamp" % clamped)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["optimalH", "H05"])
def test_high_snr_parity(codes, oracle, name):
    """+2 / +4 / +6 dB (the round-1 parity range stopped at +1 dB): BP flag, bits, converging iteration against the fp80
    oracle and, where oracle/_ref travelled with the snapshot, flag and bits against the UNMODIFIED reference; QP-ADMM
    bit-identical v."""
    from oracle.oracle import Ref, have_ref
    H, code, csr = codes[name]
    m, n = H.shape
    alpha, mu = ADMM[name]
    ref = Ref() if have_ref() else None
    rng = np.random.default_rng(5)
    for snr in (2.0, 4.0, 6.0):
        cw = None
        if name == "optimalH":          # random codewords, not only the all-zero word
            g = np.load(os.path.join(GOLDEN, "ref_optimalH.npz"))
            cw = np.ascontiguousarray(g["exp_codewords"][rng.integers(0, 60, 48)])
        y = code.channel(SEED, 33000, 48, snr, cw)
        gb, gok, git, gpost = code.bp_decode(y, snr, 100)
        ob, ook, oit, opost = oracle.bp_decode(csr, m, n, y, snr, 100)
        assert (gb == ob).all() and (gok == ook).all() and (git == oit).all(), (name, snr)
        assert (gok == 1).all()
        if cw is not None:
            assert (gb == cw).all()
        ab, aok, ait, av = code.qpadmm_decode(y, snr, alpha, mu, 1000, 1e-5)
        qb, qok, qit, qv = oracle.qpadmm_decode(csr, m, n, y, snr, alpha, mu, 1000, 1e-5)
        assert (ab == qb).all() and (aok == qok).all() and (ait == qit).all() and (av == qv).all(), (name, snr)
        if ref is not None:
            rb, rok, _ = ref.bp_decode(H, y[:16], snr, 100)
            assert (rok == gok[:16]).all() and (rb == gb[:16]).all()
            rb, rok, _ = ref.qpadmm_decode(H, y[:16], snr, alpha, mu, 1000, 1e-5)
            assert (rok == aok[:16]).all() and (rb == ab[:16]).all()


@pytest.mark.gpu
def test_qpadmm_grid_on_a_code_with_few_checks(gpu_lib, oracle):
    """Grid mode refills a slot's inv_coef table when a work item with other parameters enters it.  A code with at most
    32 live checks runs CTAs of 32 threads, fewer than the 64 table rows: every row must still be rewritten (a variable
    of degree >= 8 reads row e = 4 * degree >= 32)."""
    rng = np.random.default_rng(21)
    m, n = 30, 40
    H = np.zeros((m, n), np.uint8)
    for v in range(8):                                  # eight variables of degree 9
        H[rng.choice(m, size=9, replace=False), v] = 1
    for v in range(8, n):
        H[rng.choice(m, size=3, replace=False), v] = 1
    assert H.sum(1).max() <= 12 and H.sum(1).min() >= 1
    code = gpu_lib.Code(H=H)
    csr = dense_to_csr(H)
    alphas = np.array([0.4, 1.2, 0.0, 2.0, 0.8], np.float64)
    mus = np.array([0.9, 0.55, 0.3, 0.7, 1.5], np.float64)
    frames, snr, iters, eps = 120, 1.0, 200, 1e-5
    grid, _ = code.qpadmm_grid(alphas, mus, snr, iters, eps, SEED, 0, frames)
    for a, mu_, got in zip(alphas, mus, grid):
        want = oracle.experiment("qpadmm", csr, m, n, snr, iters, SEED, 0, frames, alpha=a, mu=mu_, eps_stop=eps)
        assert all(got[k] == want[k] for k in want), (a, mu_, got, want)
    code.close()


@pytest.mark.gpu
def test_multi_gpu_entry_points(gpu_lib, codes, oracle):
    """ldpc_experiment_run_multi / ldpc_comm_* / ldpc_allreduce_counters.  The driver's GPU tests see one
    device, so this covers the single-device case of both (a world of one rank) and -- when more devices are visible --
    the sharded run: identical counters for any device count."""
    L = gpu_lib
    H, code, csr = codes["optimalH"]
    m, n = H.shape
    dec = L.QPADMMDecoder(1.2, 0.55, 300, 1e-5)
    want = oracle.experiment("qpadmm", csr, m, n, -2.0, 300, SEED, 40, 90, alpha=1.2, mu=0.55, eps_stop=1e-5)
    got = L.experiment_multi([code], dec, -2.0, SEED, 40, 90)
    assert all(got[k] == want[k] for k in want)
    ndev = L.device_count()
    if ndev > 1:
        handles = [code] + [L.Code(H=H, device=g) for g in range(1, ndev)]
        for cnt in (2, ndev):
            got = L.experiment_multi(handles[:cnt], dec, -2.0, SEED, 40, 90)
            assert all(got[k] == want[k] for k in want), cnt
        bp = L.BeliefPropagationDecoder(100)
        one = code.experiment(bp, -3.0, SEED, 0, 5000)
        many = L.experiment_multi(handles, bp, -3.0, SEED, 0, 5000)
        assert all(one[k] == many[k] for k in want)
    # a communicator of one rank: the all-reduce is the identity
    comm = L.Comm(0, 1, L.Comm.make_id(), 0)
    assert comm.allreduce(want) == want
    vec = comm.allreduce(np.arange(37, dtype=np.uint64) << np.uint64(40))
    assert (vec == np.arange(37, dtype=np.uint64) << np.uint64(40)).all()
    comm.close()


@pytest.mark.gpu
def test_results_do_not_depend_on_the_schedule(codes, monkeypatch):
    """Stand-in for compute-sanitizer's racecheck (closed on this pool, profiles/r02_compute_sanitizer_closed.txt):
    every launch shape -- frames per team / CTA, teams per CTA, soft output on and off -- decodes the same ragged batches
    several times and must reproduce bits, flags, iteration counts and soft outputs BIT FOR BIT.  The kernels update
    messages in place, double-buffer their control words by trip parity and refill slots between two barriers: a race
    there shows up as a run-to-run or shape-to-shape difference."""
    for name, snr, sizes in (("optimalH", -3.0, (7, 301, 2500)), ("H05", -2.5, (301,)), ("reg_3_6_1008", -1.5, (5, 130))):
        H, code, _ = codes[name]
        alpha, mu = ADMM[name]
        for frames in sizes:
            y = code.channel(SEED, 77000, frames, snr)
            base = None
            for F in ("2", "4", "8"):
                for teams in ("1", "2", "5"):
                    monkeypatch.setenv("LDPC_BP_F", F)
                    monkeypatch.setenv("LDPC_BP_TEAMS", teams)
                    for rep in range(2):
                        b, ok, it, post = code.bp_decode(y, snr, 100, soft=(rep == 0))
                        sig = (b.tobytes(), ok.tobytes(), it.tobytes())
                        if base is None:
                            base, base_post = sig, post.copy()
                        assert sig == base, (name, frames, F, teams, rep)
                        # the products of a node are taken in layout order, which depends on F: soft outputs agree to rounding
                        if rep == 0:
                            conv = ok == 1
                            assert np.allclose(post[conv], base_post[conv], rtol=1e-9, atol=0), (name, frames, F, teams)
            monkeypatch.delenv("LDPC_BP_F")
            monkeypatch.delenv("LDPC_BP_TEAMS")
            base = None
            for F in ("1", "2", "4"):
                if name == "reg_3_6_1008" and F != "1":
                    continue
                monkeypatch.setenv("LDPC_ADMM_F", F)
                for rep in range(2):
                    b, ok, it, v = code.qpadmm_decode(y[:min(frames, 400)], snr, alpha, mu, 300, 1e-5)
                    sig = (b.tobytes(), ok.tobytes(), it.tobytes(), v.tobytes())
                    base = base or sig
                    assert sig == base, (name, frames, F, rep)
            monkeypatch.delenv("LDPC_ADMM_F")
