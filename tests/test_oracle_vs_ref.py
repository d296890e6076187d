"""CPU: pins the oracle against the UNMODIFIED reference compiled into
oracle/_ref (skipped where that library was not built)."""
import numpy as np
import pytest

from oracle.oracle import dense_to_csr
from tests.helpers import load_rows, small_irregular_code


def test_llr_and_variance(oracle, ref):
    for snr in (-5.0, -3.0, -0.5, 0.0, 2.25):
        assert oracle.llr_variance(snr) == ref.lib.ref_llr_variance(snr)
        for v in (-1.3, 0.0, 0.77):
            assert oracle.lib.orc_llr(v, snr) == ref.lib.ref_llr(v, snr)


@pytest.mark.parametrize("name,alpha,mu", [("optimalH", 1.2, 0.55), ("H05", 1.95, 0.5)])
def test_decoders_frame_by_frame(oracle, ref, name, alpha, mu):
    H = load_rows(name)
    m, n = H.shape
    csr = dense_to_csr(H)
    G = ref.get_orthogonal(H)[0] if name == "optimalH" else load_rows("G05")
    cw = ref.gen_random_codewords(G, 40, 239)
    for snr in (-4.0, -2.0, 0.0):
        y = np.stack([ref.transmit(snr, cw[i], 1000 + i) for i in range(len(cw))])
        rb, rok, _ = ref.bp_decode(H, y, snr, 60)
        ob, ook, _, _ = oracle.bp_decode(csr, m, n, y, snr, 60)
        assert (rok == ook).all() and (rb == ob).all()
        rb, rok, _ = ref.qpadmm_decode(H, y, snr, alpha, mu, 600, 1e-5)
        ob, ook, _, _ = oracle.qpadmm_decode(csr, m, n, y, snr, alpha, mu, 600, 1e-5)
        assert (rok == ook).all() and (rb == ob).all()
        for iters in (0, 1, 7):             # max_iter = 0 returns the initial v = (q > 0), qp_admm.h:116-119
            rb, rok, _ = ref.qpadmm_decode(H, y[:6], snr, alpha, mu, iters, 1e-5)
            ob, ook, oit, _ = oracle.qpadmm_decode(csr, m, n, y[:6], snr, alpha, mu, iters, 1e-5)
            assert (rok == ook).all() and (rb == ob).all() and (oit == iters).all()


def test_irregular_code_special_cases(oracle, ref):
    """checks of degree 0/1/2 (qp_admm.h:67-83), isolated variables, infeasible alpha."""
    H = small_irregular_code()
    m, n = H.shape
    csr = dense_to_csr(H)
    rng = np.random.default_rng(3)
    y = 1.0 + 0.8 * rng.standard_normal((30, n))
    # e_min = 0 here (isolated variables): DecodeQPADMM returns {zeros, false} for alpha >= 0
    rb, rok, _ = ref.qpadmm_decode(H, y, 1.0, 0.5, 0.6, 300, 1e-5)
    ob, ook, oit, _ = oracle.qpadmm_decode(csr, m, n, y, 1.0, 0.5, 0.6, 300, 1e-5)
    assert (rok == 0).all() and (ook == 0).all() and (rb == 0).all() and (ob == 0).all() and (oit == 0).all()
    # negative alpha is feasible for every mu > 0
    rb, rok, _ = ref.qpadmm_decode(H, y, 1.0, -0.3, 0.6, 300, 1e-5)
    ob, ook, _, _ = oracle.qpadmm_decode(csr, m, n, y, 1.0, -0.3, 0.6, 300, 1e-5)
    assert (rok == ook).all() and (rb == ob).all()
    Hc = H[:, :-2].copy()          # no isolated variables, still degree-0/1/2 checks
    Hc[0, 5] = 1                   # make every column used at least... (not required) keep structure
    csr2 = dense_to_csr(Hc)
    y2 = y[:, :-2]
    used = Hc.sum(0) > 0
    if used.all():
        rb, rok, _ = ref.qpadmm_decode(Hc, y2, 1.0, 0.4, 0.9, 300, 1e-5)
        ob, ook, _, _ = oracle.qpadmm_decode(csr2, Hc.shape[0], Hc.shape[1], y2, 1.0, 0.4, 0.9, 300, 1e-5)
        assert (rok == ook).all() and (rb == ob).all()
    rb, rok, _ = ref.bp_decode(H, y, 1.0, 30)
    ob, ook, _, _ = oracle.bp_decode(csr, m, n, y, 1.0, 30)
    assert (rok == ook).all() and (rb == ob).all()


def test_admm_problem_shape(oracle):
    """sizes of SURVEY.md section 8 for the shipped codes."""
    want = {"optimalH": (580, 700, 2320, 6960), "H05": (540, 660, 2160, 6480),
            "reg_3_6_1008": (2016, 2520, 8064, 24192)}
    for name, (T, n_var, R, nnz) in want.items():
        H = load_rows(name)
        p = oracle.admm_build(dense_to_csr(H), *H.shape)
        assert (p["n_var"], p["R"], p["nnz"]) == (n_var, R, nnz)
        assert R == 4 * T
        assert p["e"].min() == (8 if name == "reg_3_6_1008" else 4)


@pytest.mark.parametrize("name,alpha,mu", [("optimalH", 1.2, 0.55), ("H05", 1.95, 0.5), ("reg_3_6_1008", 1.2, 0.55)])
def test_decoders_at_high_snr_and_on_the_large_code(oracle, ref, name, alpha, mu):
    """The oracle is also pinned where the GPU parity tests of round 2 reach: +2 / +4 / +6 dB (where the reference's
    phi form saturates, bp.h:34) and the (3,6)-1008 code, on the reference's own mt19937 channel words."""
    H = load_rows(name)
    m, n = H.shape
    csr = dense_to_csr(H)
    frames = 24 if name != "reg_3_6_1008" else 6
    rng = np.random.default_rng(11)
    for snr in ((2.0, 4.0, 6.0) if name != "reg_3_6_1008" else (-1.0, 2.0)):
        cw = np.zeros((frames, n), np.uint8)
        y = np.stack([ref.transmit(snr, cw[i], 5000 + i) for i in range(frames)])
        rb, rok, _ = ref.bp_decode(H, y, snr, 100)
        ob, ook, oit, opost = oracle.bp_decode(csr, m, n, y, snr, 100)
        assert (rok == ook).all() and (rb == ob).all()
        rb, rok, _ = ref.qpadmm_decode(H, y, snr, alpha, mu, 1000, 1e-5)
        ob, ook, _, _ = oracle.qpadmm_decode(csr, m, n, y, snr, alpha, mu, 1000, 1e-5)
        assert (rok == ook).all() and (rb == ob).all()
