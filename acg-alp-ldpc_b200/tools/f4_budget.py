"""SURVEY.md 8 f-4 (fp32 fast BP with fp64 re-decode of the fragile frames): what could it buy?

    python acg-alp-ldpc_b200/tools/f4_budget.py > profiles/r02_f4_budget.txt      (on the GPU box)

Parity needs every frame whose outcome may depend on the precision -- frames that do not converge, frames that
converge late -- decoded again by the fp64 kernel.  This tool measures, with the fp64 kernel as it runs (experiment
mode, early exit), the share of all BP iterations spent in such frames, and from it the best case of the two-pass
scheme:  time = T_fp32(all frames) + T_fp64(fragile frames), with the fp32 pass r times faster per iteration than the
fp64 kernel (r = 1.5: half the shared-memory bytes and one-cycle instead of two-cycle arithmetic against the same
address / control instructions; r = infinity: a free first pass)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ldpc_b200 as L  # noqa: E402

SEED = 239239239
print(__doc__.split("\n\n")[0])
print("%-14s %6s %10s %10s %12s %14s %12s %12s" % ("code", "SNR", "mean its", "FER", "late (>50)", "fragile share", "best r=1.5", "best r=inf"))
for name in ("H05", "optimalH", "reg_3_6_1008"):
    code = L.Code(H=L.load_rows(name))
    frames = 200000 if name != "reg_3_6_1008" else 50000
    for snr in (-5.0, -3.0, -2.0, -1.0, 0.0):
        y = code.channel(SEED, 0, frames, snr)
        _, ok, it, _ = code.bp_decode(y, snr, 100, soft=False)
        total = float(it.sum())
        failed = ok == 0
        late = (ok == 1) & (it > 50)
        fragile = float(it[failed | late].sum())
        share = fragile / total
        print("%-14s %6.1f %10.1f %10.4f %12.4f %14.3f %12.2f %12.2f" % (
            name, snr, it.mean(), failed.mean(), late.mean(), share, 1.0 / (1.0 / 1.5 + share), 1.0 / max(share, 1e-9)))
    code.close()
print("""
Reading: "fragile share" = iterations of frames that fail or converge after more than 50 iterations / all iterations.
Where decoding is expensive (-2 dB and below, including the headline's -5 dB) the fragile frames ARE the work: even with a
first pass 1.5 x faster per iteration the two-pass scheme is SLOWER than the fp64 kernel alone (best case 0.6-0.87), and
a free first pass would buy at most 1.0-2.1 x.  Where nearly all frames converge (>= -1 dB) the decoder alone could gain
1.2-1.5 x, but there a frame takes 3-5 iterations and the Monte-Carlo point is dominated by the channel (Philox + fp64
Box-Muller, bench.py "as_run" / "channel"), which an fp32 decoder does not touch.""")
