"""GPU tests of the C++ drop-in layer -- what a user of the reference touches: `make run` (main.cpp ->
report.csv), Decoder::decode called concurrently on one instance from 200 pthreads (experiment.h:128-130,
optimize_H.cpp:12), `make run_qpadmm_params` and `make optimize`.  The executables are built from the package's
sources against libldpc_b200.so and run on the GPU; their outputs are compared with the CPU oracle's replay of the
same frames (the Philox noise is replayable, so the comparison is exact, not statistical), with the golden vectors
of the unmodified reference, and with the reference's output formats (main.cpp:47-49, 79-86).
"""
import os
import struct
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from oracle.oracle import dense_to_csr
from tests.helpers import DATA, GOLDEN, PKG, ROOT, load_rows

pytestmark = pytest.mark.gpu

SEED = 239239239
SNRS = [-5, -4.5, -4, -3.5, -3, -2.5, -2, -1.5, -1, -0.5, 0.0]          # main.cpp:27
HEADER = "Method,SNR,Sigma,FER,Time,AvgHamming,AvgHammingCorrect,AvgHammingWrong"   # main.cpp:48


def _build(tmp, source, name):
    exe = str(tmp / name)
    subprocess.run(["g++", "-std=c++17", "-pthread", "-O2", "-I" + PKG, "-I" + os.path.join(ROOT, "include"), source,
                    "-o", exe, "-L" + PKG, "-lldpc_b200", "-Wl,-rpath," + PKG], check=True)
    return exe


@pytest.fixture(scope="module")
def work(tmp_path_factory, gpu_lib):
    """a scratch directory laid out like the package directory the drivers run in (data/ next to the executable)"""
    tmp = tmp_path_factory.mktemp("drivers")
    os.symlink(DATA, tmp / "data")
    return tmp


@pytest.fixture(scope="module")
def helper(work):
    return _build(work, os.path.join(ROOT, "tests", "gpu_host_check.cpp"), "gpu_host_check")


def _words(helper, matrix, seed, count):
    out = subprocess.run([helper, "words", DATA, matrix, str(seed), str(count)], check=True, capture_output=True, text=True)
    return np.array([[int(c) for c in line.strip()] for line in out.stdout.splitlines()], np.uint8)


def _env(**kw):
    env = dict(os.environ)
    env.update({k: str(v) for k, v in kw.items()})
    return env


def test_make_run_report_csv_equals_oracle_replay(work, helper, oracle):
    """main.cpp on the GPU: report.csv has the reference's header and row format, and every FER / Hamming column equals
    the CPU oracle's replay of the same frames (same mt19937(239239239) codewords, same Philox noise)."""
    frames = 60
    exe = _build(work, os.path.join(PKG, "main.cpp"), "main")
    run = subprocess.run([exe], cwd=work, env=_env(LDPC_TESTS_NUM=frames, LDPC_GPUS=1), capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-2000:]
    lines = open(work / "report.csv").read().splitlines()
    assert lines[0] == HEADER
    assert len(lines) == 1 + 2 * len(SNRS)
    # the codewords main.cpp draws are the reference's (golden: first 60 words of the stream)
    words = _words(helper, "optimalH", SEED, frames)
    g = np.load(os.path.join(GOLDEN, "ref_optimalH.npz"))
    assert (words == g["exp_codewords"][:frames]).all()
    H = load_rows("optimalH")
    m, n = H.shape
    csr = dense_to_csr(H)
    jobs = [("BP", snr) for snr in SNRS] + [("QP-ADMM", snr) for snr in SNRS]

    def replay(job):
        name, snr = job
        if name == "BP":
            return oracle.experiment("bp", csr, m, n, float(snr), 100, SEED, 0, frames, words=words)
        return oracle.experiment("qpadmm", csr, m, n, float(snr), 10000, SEED, 0, frames, alpha=1.2, mu=0.55,
                                 eps_stop=1e-5, words=words)                 # main.cpp:29-31

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        want = list(pool.map(replay, jobs))
    some_wrong = some_right = False
    for line, (name, snr), w in zip(lines[1:], jobs, want):
        cells = line.split(",")
        assert len(cells) == 8 and cells[0] == name
        sigma = float(np.sqrt(oracle.llr_variance(float(snr))))
        total, correct = w["total"], w["correct"]
        expect = ["%.12f" % float(snr), "%.12f" % sigma, "%.12f" % ((total - correct) / total), None,
                  "%.12f" % (w["sum_hamming"] / total), "%.12f" % (w["sum_hamming_ok"] / max(1, correct)),
                  "%.12f" % (w["sum_hamming_wrong"] / max(1, total - correct))]
        for col, (got, exp) in enumerate(zip(cells[1:], expect)):
            if exp is not None:
                assert got == exp, (name, snr, col, got, exp)
        assert float(cells[4]) > 0 and len(cells[4].split(".")[1]) == 12      # Time: seconds per frame, same format
        some_wrong |= correct < total
        some_right |= correct > 0
    assert some_wrong and some_right                                         # the sweep crosses the waterfall
    # stdout lines of main.cpp:69-72
    out = run.stdout.splitlines()
    assert out[0] == "Algo: BP" and out[1 + len(SNRS)] == "Algo: QP-ADMM"
    fer0 = (want[0]["total"] - want[0]["correct"]) / want[0]["total"]
    assert out[1].startswith("\tSNR: -5.00000, FER: %.5f, (time=" % fer0)


@pytest.mark.parametrize("matrix,alpha,mu", [("optimalH", 1.2, 0.55), ("H05", 1.95, 0.5)])
def test_decoder_adapters_from_200_pthreads(work, helper, oracle, matrix, alpha, mu):
    """BeliefPropagationDecoder::decode / QPADMMDecoder::decode, one shared instance, 200 pthreads, one frame per call:
    outputs equal the unmodified reference's on its own mt19937 channel words (golden vectors) and the oracle's."""
    g = np.load(os.path.join(GOLDEN, "ref_%s.npz" % matrix))
    H = load_rows(matrix)
    m, n = H.shape
    csr = dense_to_csr(H)
    tasks = []
    for si, snr in enumerate(g["snrs"]):
        tasks += [(float(snr), g["y_%d" % si][f], g["bp_ok_%d" % si][f], g["bp_bits_%d" % si][f], g["admm_ok_%d" % si][f],
                   g["admm_bits_%d" % si][f]) for f in range(g["y_%d" % si].shape[0])]
    y60 = np.ascontiguousarray(g["exp_y"])
    ob, ook, _, _ = oracle.bp_decode(csr, m, n, y60, -3.0, 100)
    ab, aok, _, _ = oracle.qpadmm_decode(csr, m, n, y60, -3.0, alpha, mu, 1000, 1e-5)
    tasks += [(-3.0, y60[f], ook[f], ob[f], aok[f], ab[f]) for f in range(60)]
    tasks = tasks * 3                                   # 324 calls: all 200 threads find work
    fin, fout = str(work / ("in_%s.bin" % matrix)), str(work / ("out_%s.bin" % matrix))
    with open(fin, "wb") as f:
        f.write(struct.pack("<i", len(tasks)))
        for t in tasks:
            f.write(struct.pack("<d", t[0]))
            f.write(np.ascontiguousarray(t[1], np.float64).tobytes())
    run = subprocess.run([helper, "decode", DATA, matrix, fin, fout, "200", repr(alpha), repr(mu), "1000"],
                         capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-2000:]
    assert "names BP QP-ADMM" in run.stdout
    rec = np.fromfile(fout, np.uint8).reshape(len(tasks), 2 * n + 3)
    converged = 0
    for i, t in enumerate(tasks):
        bp_ok, bp_bits, bp_empty = rec[i, 0], rec[i, 1:1 + n], rec[i, 1 + n]
        ad_ok, ad_bits = rec[i, 2 + n], rec[i, 3 + n:]
        assert bp_ok == t[2] and (bp_bits == t[3]).all(), (matrix, i)
        assert bp_empty == (0 if bp_ok else 1)          # failure returns an EMPTY codeword (bp.h:198)
        assert ad_ok == t[4] and (ad_bits == t[5]).all(), (matrix, i)
        converged += int(bp_ok)
    assert 0 < converged < len(tasks)


def test_qpadmm_params_grid_equals_oracle(work, helper, oracle):
    """qpadmm_params.cpp on a 5 x 5 grid: every "alpha=, mu=: fer=" line equals the oracle's FER for that pair on the same
    frames, in the reference's scan order, and the winner follows the strict `<` first-best rule (qpadmm_params.cpp:51-81)."""
    frames, grid = 50, 5
    exe = _build(work, os.path.join(PKG, "qpadmm_params.cpp"), "qpadmm_params")
    run = subprocess.run([exe], cwd=work, env=_env(LDPC_TESTS_NUM=frames, LDPC_GRID=grid, LDPC_GPUS=1),
                         capture_output=True, text=True)
    assert run.returncode == 0, run.stderr[-2000:]
    words = _words(helper, "optimalH", 239, frames)                          # qpadmm_params.cpp:45-46
    H = load_rows("optimalH")
    m, n = H.shape
    csr = dense_to_csr(H)
    pts = [(0.0 + ((3.0 - 0.0) / (grid - 1)) * ai, 0.0 + ((3.0 - 0.0) / (grid - 1)) * mi) for ai in range(grid) for mi in range(grid)]

    def replay(p):
        r = oracle.experiment("qpadmm", csr, m, n, -3.0, 1000, SEED, 0, frames, alpha=p[0], mu=p[1], eps_stop=1e-5, words=words)
        return (r["total"] - r["correct"]) / r["total"]

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        fers = list(pool.map(replay, pts))
    err = [l for l in run.stderr.splitlines() if l.startswith("alpha=")]
    assert len(err) == grid * grid
    for line, (a, mu_), fer in zip(err, pts, fers):
        assert line == "alpha=%g, mu=%g: fer=%g" % (a, mu_, fer), line       # cerr keeps the default format, as the reference's
    best = min(range(len(pts)), key=lambda i: (fers[i], i))                   # first strictly smaller wins
    out = run.stdout.splitlines()
    assert out[-4:] == ["Best parameters:", "alpha=%.5f" % pts[best][0], "mu=%.5f" % pts[best][1], "fer=%.5f" % fers[best]]
    assert any(f < 1.0 for f in fers) and any(f == 1.0 for f in fers)         # feasible and infeasible pairs both occur


def test_optimize_H_trajectory_is_window_invariant(work):
    """optimize_H.cpp: the speculative window (proposals evaluated concurrently, generator rewound on accept) must not
    change the chain: stdout and every saved matrix are byte-identical for windows 1, 3 and 8 -- with evaluation threads
    and with evaluation processes (the multi-GPU mode, forced here on one GPU)."""
    exe = _build(work, os.path.join(PKG, "optimize_H.cpp"), "optimize_H")
    outs, mats = [], []
    for window, procs in ((1, 0), (3, 0), (8, 0), (4, 4), (2, 5)):
        save = str(work / ("opt_w%d_p%d.txt" % (window, procs)))
        run = subprocess.run([exe], cwd=work, capture_output=True, text=True,
                             env=_env(LDPC_OPT_ITERS=40, LDPC_OPT_WINDOW=window, LDPC_OPT_PROCS=procs, LDPC_OPT_SAVE=save,
                                      LDPC_OPT_START="data/H05"))
        assert run.returncode == 0, run.stderr[-2000:]
        outs.append(run.stdout)
        mats.append(open(save).read() if os.path.exists(save) else "")
    assert all(o == outs[0] for o in outs)
    assert all(m == mats[0] for m in mats)
    lines = outs[0].splitlines()
    assert lines[0].startswith("initial FER=") and sum(l.startswith("\tproposal: FER=") for l in lines) == 40
    assert any(l.startswith("accept, FER=") for l in lines), "no proposal was accepted: the rewind path was not exercised"
    assert mats[0] != ""
