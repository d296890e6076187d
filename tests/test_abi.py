"""CPU: the C-ABI library builds, loads and exports every symbol that
include/ldpc_b200.h declares; without a GPU it fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.helpers import ROOT, PKG, load_rows


@pytest.fixture(scope="module")
def cdll():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ldpc_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return C.CDLL(mod.build())


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ldpc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ldpc_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(cdll):
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(cdll, name), name
    assert cdll.ldpc_abi_version() == 1


def test_no_cpu_fallback_without_gpu(cdll):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    H = load_rows("optimalH")
    handle = C.c_void_p()
    st = cdll.ldpc_code_create_dense(H.shape[0], H.shape[1], H.ctypes.data_as(C.c_void_p), 0, C.byref(handle))
    assert st == -2 and not handle.value          # LDPC_E_CUDA
    cdll.ldpc_last_error.restype = C.c_char_p
    assert b"CUDA" in cdll.ldpc_last_error()


def test_argument_validation(cdll):
    cdll.ldpc_last_error.restype = C.c_char_p
    handle = C.c_void_p()
    assert cdll.ldpc_code_create(0, 0, None, None, 0, C.byref(handle)) == -1
    rp = np.array([0, 2, 1], np.int32)
    ci = np.array([0, 1], np.int32)
    assert cdll.ldpc_code_create(2, 4, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p), 0,
                                 C.byref(handle)) == -1
    rp = np.array([0, 2], np.int32)
    ci = np.array([1, 1], np.int32)     # not strictly ascending
    assert cdll.ldpc_code_create(1, 4, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p), 0,
                                 C.byref(handle)) == -1
    assert cdll.ldpc_bp_decode(None, None, 0, C.c_double(0), 1, 1, None, None, None, None) == -1


def test_multi_gpu_entry_points_validate_and_fail_loudly(cdll):
    """ldpc_experiment_run_multi / ldpc_comm_*: argument checks, and no silent success without a GPU"""
    import torch
    assert cdll.ldpc_experiment_run_multi(None, 0, None, C.c_double(0), 0, 0, 0, 0, None, 0, None, None) == -1
    assert cdll.ldpc_comm_init(0, 0, None, 0, None) == -1
    assert cdll.ldpc_allreduce_counters(None, None, 0) == -1
    cdll.ldpc_comm_destroy(None)                                   # a no-op
    if not torch.cuda.is_available():
        comm = C.c_void_p()
        ident = (C.c_uint8 * 128)()
        assert cdll.ldpc_comm_init(0, 1, ident, 0, C.byref(comm)) in (-2, -4) and not comm.value   # LDPC_E_CUDA / no NCCL
