"""BASELINE.json configs[1] / configs[3] as run: an SNR sweep with N frames per point through ldpc_experiment_run
(device-side codewords u*G or all-zero, Philox AWGN, decoding, verdict, counting), timed with CUDA events.

    python acg-alp-ldpc_b200/tools/config_sweep.py --code H05 --algo bp --frames 10000000
"""
import argparse
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ldpc_b200 as L  # noqa: E402
from ldpc_b200 import load_rows  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--code", default="H05")
ap.add_argument("--algo", default="bp")
ap.add_argument("--frames", type=int, default=10 ** 7)
ap.add_argument("--snrs", default="-5,-4.5,-4,-3.5,-3,-2.5,-2,-1.5,-1,-0.5,0")
ap.add_argument("--alpha", type=float, default=1.2)
ap.add_argument("--mu", type=float, default=0.55)
ap.add_argument("--max-iter", type=int, default=None)
a = ap.parse_args()
H = load_rows(a.code)
m, n = H.shape
code = L.Code(H=H)
dec = L.BeliefPropagationDecoder(a.max_iter or 100) if a.algo == "bp" else L.QPADMMDecoder(a.alpha, a.mu, a.max_iter or 10000, 1e-5)
print("# %s on %s (%d x %d), %d frames per SNR point, all-zero codeword (the decoders are symmetric), seed 239239239, one B200" % (
    dec.name(), a.code, m, n, a.frames))
print("%6s %12s %12s %10s %12s %12s %12s %10s" % ("snr", "FER", "+-95%", "BER", "mean iters", "gpu s", "frames/s", "info Gb/s"))
total = 0.0
for snr in (float(x) for x in a.snrs.split(",")):
    r = code.experiment(dec, snr, 239239239, 0, a.frames)
    fer = 1 - r["correct"] / r["total"]
    ber = r["bit_errors"] / max(1, r["frames_with_bits"] * n)
    total += r["gpu_seconds"]
    print("%6.1f %12.3e %12.1e %10.2e %12.2f %12.3f %12.0f %10.3f" % (
        snr, fer, 1.96 * math.sqrt(max(fer * (1 - fer), 1e-300) / r["total"]), ber, r["sum_iters"] / r["total"],
        r["gpu_seconds"], r["total"] / r["gpu_seconds"], r["total"] / r["gpu_seconds"] * (n - m) / 1e9), flush=True)
print("# sweep total: %.2f s of GPU time" % total)
