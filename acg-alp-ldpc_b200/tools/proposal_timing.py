"""Where does one optimize_H.cpp proposal spend its time?  (code handle per H, 1000-frame QP-ADMM FER evaluation)"""
import os, sys, time
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ldpc_b200 as L
from ldpc_b200 import load_rows
H = load_rows("optimalH")
rng = np.random.default_rng(1)
dec = L.QPADMMDecoder(1.95, 0.5, 1000, 1e-5)
words = np.zeros((1000, H.shape[1]), np.uint8)
L.Code(H=H).close()
for rep in range(5):
    Hp = H.copy()
    r0 = 20 * rng.integers(8); c0 = 20 * rng.integers(14); sh = rng.integers(20)     # toggle / shift one circulant
    Hp[r0:r0 + 20, c0:c0 + 20] = np.roll(np.eye(20, dtype=np.uint8), sh, axis=1)
    t0 = time.perf_counter(); code = L.Code(H=Hp); t1 = time.perf_counter()
    r = code.experiment(dec, -3.0, 239239239, 0, 1000, source=L.CW_TABLE, words=words); t2 = time.perf_counter()
    r2 = code.experiment(dec, -3.0, 239239239, 0, 1000, source=L.CW_TABLE, words=words); t3 = time.perf_counter()
    code.close(); t4 = time.perf_counter()
    print("create %.1f ms, first run %.1f ms (gpu %.1f), second run %.1f ms (gpu %.1f), destroy %.1f ms, FER %.3f" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * r["gpu_seconds"], 1e3 * (t3 - t2), 1e3 * r2["gpu_seconds"], 1e3 * (t4 - t3),
        1 - r["correct"] / r["total"]))
