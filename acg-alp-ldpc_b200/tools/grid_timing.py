"""GPU time of the (alpha, mu) grid search of qpadmm_params.cpp: one launch for all pairs (ldpc_qpadmm_grid_run)
against one launch per pair (ldpc_experiment_run).

    python acg-alp-ldpc_b200/tools/grid_timing.py [--grid 13] [--frames 1000]
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ldpc_b200 as L  # noqa: E402
from ldpc_b200 import load_rows  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=13)
ap.add_argument("--frames", type=int, default=1000)
ap.add_argument("--per-point", action="store_true", help="also time one launch per pair")
a = ap.parse_args()
code = L.Code(H=load_rows("optimalH"))
pts = [(3.0 / (a.grid - 1) * i, 3.0 / (a.grid - 1) * j) for i in range(a.grid) for j in range(a.grid)]
alphas = np.array([p[0] for p in pts])
mus = np.array([p[1] for p in pts])
for rep in range(2):
    t0 = time.perf_counter()
    res, secs = code.qpadmm_grid(alphas, mus, -3.0, 1000, 1e-5, 239239239, 0, a.frames)
    wall = time.perf_counter() - t0
    iters = sum(r["sum_iters"] for r in res)
    feas = sum(1 for r in res if r["decoder_fail"] == 0)
    print("batched: %d pairs (%d feasible) x %d frames: wall %.3f s, gpu %.3f s, %.3e frame-iterations, %.3e frame-iter/s" % (
        len(pts), feas, a.frames, wall, secs, iters, iters / max(secs, 1e-9)))
best = min(range(len(res)), key=lambda i: (res[i]["total"] - res[i]["correct"], i))
print("best pair: alpha=%.5f mu=%.5f fer=%.5f" % (alphas[best], mus[best], 1 - res[best]["correct"] / res[best]["total"]))
if a.per_point:
    t0 = time.perf_counter()
    gsecs = 0.0
    for (al, mu), want in zip(pts, res):
        got = code.experiment(L.QPADMMDecoder(al, mu, 1000, 1e-5), -3.0, 239239239, 0, a.frames)
        gsecs += got["gpu_seconds"]
        assert all(got[k] == want[k] for k in want), (al, mu)
    print("per pair: wall %.3f s, gpu %.3f s (identical counters)" % (time.perf_counter() - t0, gsecs))
