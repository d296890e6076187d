"""Builds acg-alp-ldpc_b200/libldpc_b200.so in-tree with nvcc for sm_100a.

    python acg-alp-ldpc_b200/build.py [--force] [--verbose]

The library is the product: hand-written CUDA kernels behind the C ABI of
include/ldpc_b200.h.  It is git-ignored (source-only history) but travels to
the GPU box with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libldpc_b200.so")
SOURCES = ["code.cu", "admm_layout.cu", "api.cu", "bp_kernel.cu", "bp_lr_kernel.cu", "qpadmm_kernel.cu", "qpadmm_chk_kernel.cu", "channel_kernel.cu"]
HEADERS = ["ldpc_internal.h", "frame.cuh", "channel.cuh", "slots.cuh", "slots_team.cuh", "bpmath.cuh", "admm_rows.cuh", "smem_ptx.cuh", os.path.join(ROOT, "include", "ldpc_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps + [os.path.abspath(__file__)])


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        obj = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write("== %s\n%s\n" % (s, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
