"""Shared helpers for the test-suite (not collected by pytest)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "acg-alp-ldpc_b200")
DATA = os.path.join(PKG, "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


from ldpc_b200 import load_rows  # noqa: E402,F401  (the .rows parser lives in the package)


def small_irregular_code(seed=7, m=12, n=24):
    """A small H with checks of degree 0, 1, 2 and >= 3 (the special cases of
    qp_admm.h:67-83) and variables of degree 0."""
    rng = np.random.default_rng(seed)
    H = np.zeros((m, n), np.uint8)
    degs = [0, 1, 2, 2, 3, 3, 4, 5, 6, 3, 4, 7][:m]
    for r, d in enumerate(degs):
        H[r, rng.choice(n - 2, size=d, replace=False)] = 1   # last two variables stay isolated
    return H


def wilson_interval(k, n, z=1.96):
    if n == 0:
        return 0.0, 1.0
    p = k / n
    den = 1 + z * z / n
    centre = (p + z * z / (2 * n)) / den
    half = z * np.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / den
    return max(0.0, centre - half), min(1.0, centre + half)
