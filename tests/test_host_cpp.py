"""CPU: the C++ host mirror of the reference's utils/ and experiment.h
(acg-alp-ldpc_b200/{utils,algo,experiment.h,main.cpp}) against golden vectors
produced by the unmodified reference."""
import os
import subprocess

import numpy as np
import pytest

from tests.helpers import GOLDEN, PKG, ROOT, DATA, load_rows


@pytest.fixture(scope="module")
def host_out(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("host") / "host_check")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ldpc_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = mod.build()
    subprocess.run(["g++", "-std=c++17", "-pthread", "-O1", "-I" + PKG, "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "host_check.cpp"), "-o", exe, "-L" + PKG, "-lldpc_b200",
                    "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    out = subprocess.run([exe, DATA], check=True, capture_output=True, text=True).stdout
    return [l.split() for l in out.splitlines()]


def rows(host_out, tag):
    return [l[1:] for l in host_out if l[0] == tag]


def test_get_orthogonal_and_codewords_match_reference(host_out):
    g = np.load(os.path.join(GOLDEN, "ref_generator_optimalH.npz"))
    n = int(g["n"])
    G_ref = np.unpackbits(g["G"], axis=1)[:, :n]
    assert rows(host_out, "shape")[0] == ["160", "280"]
    assert rows(host_out, "orth_ok")[0] == ["1", "120"]
    G = np.array([[int(c) for c in r[0]] for r in rows(host_out, "G")], np.uint8)
    assert (G == G_ref).all()
    W = rows(host_out, "W")
    words = np.array([[int(c) for c in r[0]] for r in W], np.uint8)
    assert (words == g["first_words"]).all()          # same mt19937 draws, same row selection rule
    assert all(r[1] == "1" for r in W)


def test_transmit_matches_reference_bit_for_bit(host_out):
    g = np.load(os.path.join(GOLDEN, "ref_optimalH.npz"))
    y_ref = g["y_0"]                                  # snr -3, seeds 1.., same codewords
    for f, r in enumerate(rows(host_out, "Y")):
        y = np.array([int(x, 16) for x in r], np.uint64).view(np.float64)
        assert y.tobytes() == y_ref[f].tobytes()


def test_text_formats_and_algebra(host_out):
    assert rows(host_out, "roundtrip")[0] == ["1"]
    # "2" parses as 0; trailing comma optional (utils/parse_data.h:15-21)
    assert rows(host_out, "quirks")[0] == ["2", "1001", "0110"]
    assert rows(host_out, "deficient")[0] == ["0"]
    assert rows(host_out, "syndrome_zero")[0] == ["1"]
    assert rows(host_out, "vM")[0] == ["1"]


def test_generic_experiment_path_is_thread_count_invariant(host_out):
    a, b = rows(host_out, "exp")
    assert a == b and a[0] == "200"
    total, correct, pseudo, ham, ham_ok, ham_wrong = (int(x) for x in a)
    assert ham == ham_ok + ham_wrong and correct + pseudo <= total and correct > 0


def test_rows_files_agree_with_dense_expansion():
    H = load_rows("H05")
    assert H.shape == (160, 280) and H.sum() == 860
    assert sorted(set(H.sum(1))) == [4, 5, 6, 7]
    H = load_rows("reg_3_6_1008")
    assert (H.sum(0) == 3).all() and (H.sum(1) == 6).all()
    ov = H.astype(np.int32) @ H.T.astype(np.int32)
    np.fill_diagonal(ov, 0)
    assert ov.max() <= 1                               # no 4-cycles


@pytest.mark.parametrize("driver", ["main.cpp", "qpadmm_params.cpp", "optimize_H.cpp"])
def test_drivers_compile_and_link(driver, tmp_path):
    """the reference's three executables (Makefile:8-24) build against the C ABI library"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ldpc_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = mod.build()
    exe = str(tmp_path / driver.replace(".cpp", ""))
    subprocess.run(["g++", "-std=c++17", "-pthread", "-O1", "-I" + PKG, "-I" + os.path.join(ROOT, "include"),
                    os.path.join(PKG, driver), "-o", exe, "-L" + PKG, "-lldpc_b200",
                    "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    assert os.path.exists(exe)
