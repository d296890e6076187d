"""Device-resident throughput of the decode kernels under different launch shapes
(the LDPC_* environment knobs), for tuning on a B200.

    python acg-alp-ldpc_b200/tools/sweep.py --algo bp --code H05 --frames 65536 \
        --set LDPC_BP_F=1,2,4 --set LDPC_BP_THREADS=192,224,288,448
"""
import argparse
import itertools
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import torch  # noqa: E402
import ldpc_b200 as L  # noqa: E402
from ldpc_b200 import load_rows  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--algo", default="bp")
ap.add_argument("--code", default=None)
ap.add_argument("--frames", type=int, default=65536)
ap.add_argument("--iters", type=int, default=None)
ap.add_argument("--snr", type=float, default=None)
ap.add_argument("--early-exit", action="store_true")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--set", action="append", default=[], help="ENV=v1,v2,... (cartesian product)")
a = ap.parse_args()

name = a.code or ("H05" if a.algo == "bp" else "optimalH")
code = L.Code(H=load_rows(name))
n = code.n
snr = a.snr if a.snr is not None else (-5.0 if a.algo == "bp" else -3.0)
iters = a.iters or (100 if a.algo == "bp" else 1000)
dev = torch.device("cuda", 0)
y = torch.empty((a.frames, n), dtype=torch.float64, device=dev)
bits = torch.empty((a.frames, n), dtype=torch.uint8, device=dev)
ok = torch.empty(a.frames, dtype=torch.uint8, device=dev)
its = torch.empty(a.frames, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream()
code.channel_device(239239239, 0, a.frames, snr, y.data_ptr(), 0, stream.cuda_stream)


def run():
    if a.algo == "bp":
        code.bp_decode_device(y.data_ptr(), a.frames, snr, iters, a.early_exit, bits.data_ptr(), ok.data_ptr(),
                              its.data_ptr(), 0, stream.cuda_stream)
    else:
        code.qpadmm_decode_device(y.data_ptr(), a.frames, snr, 1.2, 0.55, iters, 1e-5 if a.early_exit else 0.0,
                                  bits.data_ptr(), ok.data_ptr(), its.data_ptr(), 0, stream.cuda_stream)


keys = [s.split("=")[0] for s in a.set]
vals = [s.split("=")[1].split(",") for s in a.set]
print("# %s %s frames=%d iters=%d snr=%g early_exit=%s" % (a.algo, name, a.frames, iters, snr, a.early_exit))
ref_sig = None
for combo in itertools.product(*vals) if vals else [()]:
    for k, v in zip(keys, combo):
        os.environ[k] = v
    try:
        run()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run()
            e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        sig = (int(ok.sum().item()), int(its.sum().item()), int(bits.sum().item()))
        if ref_sig is None:
            ref_sig = sig
        mit = its.double().mean().item()
        print("%-40s %9.3f ms  %10.0f frames/s  %.3e frame-iter/s  mean_iters=%.1f  %s" % (
            " ".join("%s=%s" % kv for kv in zip(keys, combo)), best, a.frames / best * 1e3,
            a.frames * mit / best * 1e3, mit, "same-results" if sig == ref_sig else "RESULTS-DIFFER %s" % (sig,)))
    except Exception as ex:  # noqa: BLE001
        print("%-40s FAILED %s" % (" ".join("%s=%s" % kv for kv in zip(keys, combo)), ex))
