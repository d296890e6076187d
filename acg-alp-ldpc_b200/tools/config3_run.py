"""BASELINE.json configs[3] as stated: the synthetic (3,6)-regular n = 1008 code, BP and QP-ADMM, 1e9 frames sharded over
all visible GPUs through the Monte-Carlo path (ldpc_experiment_run_multi: shards by global frame index, NCCL all-reduce
of the counter blocks on the devices).

    python acg-alp-ldpc_b200/tools/config3_run.py [--frames 1000000000] [--gpus N] > profiles/rNN_config3_1e9.txt
"""
import argparse
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ldpc_b200 as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=10 ** 9)
ap.add_argument("--gpus", type=int, default=0)
ap.add_argument("--points", default="bp:0.0,bp:-1.5,qpadmm:1.0")
a = ap.parse_args()
gpus = a.gpus or L.device_count()
H = L.load_rows("reg_3_6_1008")
codes = [L.Code(H=H, device=g) for g in range(gpus)]
print("# configs[3]: (3,6)-regular 504 x 1008, %d frames per point, %d GPU(s), all-zero codeword, seed 239239239" % (a.frames, gpus))
for pt in a.points.split(","):
    algo, snr = pt.split(":")
    snr = float(snr)
    dec = L.BeliefPropagationDecoder(100) if algo == "bp" else L.QPADMMDecoder(1.2, 0.55, 1000, 1e-5)
    L.experiment_multi(codes, dec, snr, 239239239, 0, 4096 * gpus)            # warm-up: tables, communicator
    t0 = time.perf_counter()
    r = L.experiment_multi(codes, dec, snr, 239239239, 0, a.frames)
    wall = time.perf_counter() - t0
    print("%-8s snr %5.1f  frames %d  wall %8.2f s  %.4g frames/s  FER %.4g  BER %.4g  mean iterations %.2f  counters %s" % (
        dec.name(), snr, r["total"], wall, r["total"] / wall, 1 - r["correct"] / r["total"],
        r["bit_errors"] / max(1, r["frames_with_bits"] * 1008), r["sum_iters"] / r["total"],
        {k: r[k] for k in L.CNT_NAMES}), flush=True)
