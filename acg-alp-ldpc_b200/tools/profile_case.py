"""One short launch of a decode kernel, for ncu (--set full replays a kernel ~40
times, so the case must be small).

    python acg-alp-ldpc_b200/tools/profile_case.py --algo bp|qpadmm [--code H05] [--frames N] [--iters K]
"""
import argparse
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ldpc_b200 as L  # noqa: E402
from ldpc_b200 import load_rows  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--algo", default="bp")
ap.add_argument("--code", default=None)
ap.add_argument("--frames", type=int, default=148 * 24)
ap.add_argument("--iters", type=int, default=None)
ap.add_argument("--snr", type=float, default=None)
ap.add_argument("--early-exit", action="store_true")
a = ap.parse_args()
name = a.code or ("H05" if a.algo == "bp" else "optimalH")
code = L.Code(H=load_rows(name))
snr = a.snr if a.snr is not None else (-5.0 if a.algo == "bp" else -3.0)
y = code.channel(239239239, 0, a.frames, snr)
for rep in range(2):
    t0 = time.perf_counter()
    if a.algo == "bp":
        bits, ok, it, _ = code.bp_decode(y, snr, a.iters or 100, early_exit=a.early_exit, soft=False)
    else:
        bits, ok, it, _ = code.qpadmm_decode(y, snr, 1.2, 0.55, a.iters or 200, 1e-5 if a.early_exit else 0.0, soft=False)
    dt = time.perf_counter() - t0
print("%s %s frames=%d mean_iters=%.1f ok=%d  %.3f ms (host call, incl. copies)" % (
    a.algo, name, a.frames, it.mean(), int(ok.sum()), dt * 1e3))
