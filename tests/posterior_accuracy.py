"""Which posterior LLR is right at large magnitudes?  (run on the GPU box; not collected by pytest)

    python tests/posterior_accuracy.py [--frames 12] [--snr -1.0] [--out gpurun_out/posterior_accuracy.txt]

The per-frame parity campaign finds the CUDA posterior LLRs within 1e-5 of the fp80 oracle's for |LLR| <= 30 and up
to 1e-2 away above.  This tool decodes the same frames a third time with 60-digit arithmetic (mpmath; the reference's
flooding schedule and phi form, bp.h:155-205) and measures both against it: the reference's phi form keeps
exp(-x) only to 2^-64 exp(x) / 2 relative (tanh(x/2) = 1 - 2 exp(-x) in fp80), the likelihood-ratio kernel keeps full
fp64 relative accuracy at every magnitude.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "acg-alp-ldpc_b200"))
from tests.helpers import load_rows  # noqa: E402


def mp_bp(H, y, snr, iters, mp):
    """flooding sum-product for exactly `iters` iterations in the phi domain, arbitrary precision"""
    m, n = H.shape
    var = mp.mpf(10) ** (-mp.mpf(snr) / 10) / 2
    llr = [2 * mp.mpf(float(v)) / var for v in y]
    rows = [np.flatnonzero(H[r]) for r in range(m)]
    # phi(x) = -log(tanh(x/2)) = 2 atanh(exp(-x)): this form keeps full relative accuracy at every magnitude (the
    # log/tanh form returns 0 once exp(-x) drops below the working precision)
    phi = lambda x: 2 * mp.atanh(mp.exp(-x))
    v2c = {(r, v): llr[v] for r in range(m) for v in rows[r]}
    post = None
    for _ in range(iters):
        c2v = {}
        for r in range(m):
            mags = {v: phi(abs(v2c[(r, v)])) for v in rows[r]}
            sgn = {v: (-1 if v2c[(r, v)] <= 0 else 1) for v in rows[r]}
            for v in rows[r]:
                s = mp.mpf(0)
                sg = 1
                for u in rows[r]:
                    if u != v:
                        s += mags[u]
                        sg *= sgn[u]
                c2v[(r, v)] = sg * phi(s)
        tot = [llr[v] for v in range(n)]
        for (r, v), msg in c2v.items():
            tot[v] += msg
        post = tot
        for (r, v) in v2c:
            v2c[(r, v)] = tot[v] - c2v[(r, v)]
    return post


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--snr", type=float, default=-1.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "posterior_accuracy.txt"))
    a = ap.parse_args()
    import mpmath as mp
    import ldpc_b200 as L
    from oracle.oracle import Oracle, dense_to_csr
    mp.mp.dps = 60
    H = load_rows("optimalH")
    m, n = H.shape
    code = L.Code(H=H)
    orc = Oracle()
    y = code.channel(239239239, 0, a.frames, a.snr)
    gb, gok, git, gpost = code.bp_decode(y, a.snr, 100)
    ob, ook, oit, opost = orc.bp_decode(dense_to_csr(H), m, n, y, a.snr, 100)
    rows = []
    for f in range(a.frames):
        if not (gok[f] and ook[f] and git[f] == oit[f]):
            continue
        truth = np.array([float(t) for t in mp_bp(H, y[f], a.snr, int(git[f]), mp)])
        for lo, hi in ((0, 30), (30, 40), (40, 1e9)):
            sel = (np.abs(truth) >= lo) & (np.abs(truth) < hi) & np.isfinite(opost[f])
            if sel.any():
                rows.append((lo, hi, int(sel.sum()), float(np.max(np.abs(gpost[f][sel] - truth[sel]) / np.abs(truth[sel]))),
                             float(np.max(np.abs(opost[f][sel] - truth[sel]) / np.abs(truth[sel])))))
    lines = ["# posterior LLR against 60-digit arithmetic, optimalH @ %g dB, %d frames (same channel samples, same iteration count)" % (a.snr, a.frames),
             "%-18s %8s %22s %22s" % ("|LLR| range", "values", "max rel err CUDA", "max rel err fp80 oracle")]
    for lo, hi in ((0, 30), (30, 40), (40, 1e9)):
        sel = [r for r in rows if r[0] == lo]
        if sel:
            lines.append("%-18s %8d %22.3g %22.3g" % ("[%g, %s)" % (lo, "inf" if hi > 1e8 else "%g" % hi), sum(r[2] for r in sel),
                                                       max(r[3] for r in sel), max(r[4] for r in sel)))
    print("\n".join(lines))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
