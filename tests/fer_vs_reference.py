"""(run on the GPU box; not collected by pytest)  FER of the CUDA decoders against the UNMODIFIED reference (oracle/_ref, one single-threaded process per host
core -- the reference's own multi-threaded harness is racy) at every SNR of main.cpp:27, with 95 %
binomial confidence intervals.  BASELINE.json's bar: the two FERs agree within the intervals at every SNR.

    python tests/fer_vs_reference.py [--gpu-frames 20000] [--ref-frames-per-core 100] [--out FILE]

GPU side: ldpc_experiment_run with codewords u*G (G = the reference's GetOrtogonal(H)) and the Philox channel.
Reference side: the single-threaded exp() loop (experiment.h:85-121) replayed frame by frame with the reference's own
functions -- codeword f of gen_random_codewords(G, N, mt19937(239239239)), noise transmit(.., mt19937(f + 1)), the
unmodified decoder, IsCodeword -- split over the host cores by frame range.  The two sides use different noise
streams: the comparison is statistical.
"""
import argparse
import math
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "acg-alp-ldpc_b200"))
sys.path.insert(0, ROOT)
from ldpc_b200 import load_rows  # noqa: E402

SNRS = [-5.0 + 0.5 * i for i in range(11)]          # main.cpp:27
DECODERS = [("BP", "bp", dict(max_iter=100)), ("QP-ADMM", "qpadmm", dict(max_iter=10000, alpha=1.2, mu=0.55, eps_stop=1e-5))]


def _ref_chunk(job):
    """frames [begin, begin + count) of the reference's single-threaded exp() loop (experiment.h:85-121): codeword
    f of gen_random_codewords(G, total, mt19937(239239239)) (main.cpp:63-64), noise from mt19937(f + 1), the
    unmodified decoder, the verdict of experiment.h:109-118"""
    name, algo, kw, snr, total, begin, count = job
    from oracle.oracle import Ref
    ref = Ref()
    H = load_rows(name)
    G, ok = ref.get_orthogonal(H)
    cw = ref.gen_random_codewords(G, total, 239239239)[begin:begin + count]
    y = np.stack([ref.transmit(snr, cw[k], begin + k + 1) for k in range(count)])
    t0 = time.time()
    if algo == "bp":
        bits, okf, _ = ref.bp_decode(H, y, snr, kw["max_iter"])
    else:
        bits, okf, _ = ref.qpadmm_decode(H, y, snr, kw["alpha"], kw["mu"], kw["max_iter"], kw["eps_stop"])
    secs = time.time() - t0
    correct = 0
    for k in range(count):
        if okf[k] and ref.is_codeword(H, bits[k]) and (bits[k] == cw[k]).all():
            correct += 1
    return correct, count, secs


def ci(p, n):
    return 1.96 * math.sqrt(max(p * (1 - p), 1e-12) / n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", default="optimalH")
    ap.add_argument("--gpu-frames", type=int, default=20000)
    ap.add_argument("--ref-frames-per-core", type=int, default=100)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fer_vs_reference.txt"))
    a = ap.parse_args()
    import ldpc_b200 as L
    from oracle.oracle import Ref
    cores = len(os.sched_getaffinity(0))
    H = load_rows(a.code)
    G, ok = Ref().get_orthogonal(H)
    assert ok
    code = L.Code(H=H)
    code.set_generator(G)
    lines = ["# FER: CUDA kernels (%d frames per point, Philox channel, codewords u*G) vs the unmodified reference (%d cores x %d "
             "frames, 1 thread per process), code %s" % (a.gpu_frames, cores, a.ref_frames_per_core, a.code),
             "# agree = |FER_gpu - FER_ref| <= 1.96 sqrt(var_gpu + var_ref)",
             "%-8s %5s  %10s %8s  %10s %8s  %8s %6s  %12s %12s" % ("decoder", "snr", "FER_gpu", "+-", "FER_ref", "+-", "diff", "agree",
                                                                   "gpu frames/s", "ref frames/s")]
    ctx = mp.get_context("fork")
    all_ok = True
    for label, algo, kw in DECODERS:
        dec = L.BeliefPropagationDecoder(kw["max_iter"]) if algo == "bp" else L.QPADMMDecoder(kw["alpha"], kw["mu"], kw["max_iter"], kw["eps_stop"])
        for snr in SNRS:
            g = code.experiment(dec, snr, 239239239, 0, a.gpu_frames, source=L.CW_GENERATOR)
            pg = 1.0 - g["correct"] / g["total"]
            t0 = time.time()
            with ctx.Pool(cores) as pool:
                parts = pool.map(_ref_chunk, [(a.code, algo, kw, snr, cores * a.ref_frames_per_core, i * a.ref_frames_per_core,
                                               a.ref_frames_per_core) for i in range(cores)])
            wall = time.time() - t0
            rc, rt = sum(p[0] for p in parts), sum(p[1] for p in parts)
            pr = 1.0 - rc / rt
            diff = abs(pg - pr)
            agree = diff <= 1.96 * math.sqrt(max(pg * (1 - pg), 1e-12) / g["total"] + max(pr * (1 - pr), 1e-12) / rt) + 1e-12
            all_ok &= agree
            lines.append("%-8s %5.1f  %10.5f %8.5f  %10.5f %8.5f  %8.5f %6s  %12.0f %12.1f" % (
                label, snr, pg, ci(pg, g["total"]), pr, ci(pr, rt), diff, "yes" if agree else "NO",
                g["total"] / max(g["gpu_seconds"], 1e-9), rt / wall))
            print(lines[-1], flush=True)
    lines.append("# all points agree: %s" % all_ok)
    print(lines[-1])
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
