#!/usr/bin/env python
"""Throughput benchmark of the hot path: decoded frames/s at fixed iterations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): BP sum-product, 100 fixed iterations, on the
H05 code (160 x 280), AWGN frames at -5 dB.  One "step" = one batch of
--frames frames through the decode kernel; 1e7 frames per SNR point = 1e7/frames
steps.  -5 dB is used because there the reference's BP (which cannot switch its
syndrome exit off) also runs all 100 iterations on >= 99.8 % of frames, so the CPU
arm does the same work on the same kind of input.  QP-ADMM (optimalH, 1000 fixed
iterations, eps_stop = 0) is measured in the same run and reported under "qpadmm"; both decoders on the
synthetic (3,6)-regular n = 1008 code (configs[3]) under "reg_3_6_1008".

Also in the line: BP on optimalH, QP-ADMM at main.cpp's 10000 iterations, the decoders as
they really run (experiment mode, early exit, -3 dB and 0 dB, mean iterations), the channel
kernel against the HBM roofline, and configs[3] through the Monte-Carlo path: a fixed total of
frames of the (3,6)-1008 code sharded over the ranks by global frame index, the counter blocks
all-reduced by the library's NCCL communicator inside the timed region ("experiment_scaling").

One JSON line on stdout (rank 0).  Keys follow the driver's contract.  `value` is the throughput
of the whole job over K launches (CUDA events) plus the counter all-reduce (wall clock), max over
ranks; kernel_* are the launches alone, wall_ms_per_step the raw wall clock of the region.  The roofline of BP is
shared-memory bandwidth, that of QP-ADMM the FP64 pipe (messages never leave the SM, so HBM
traffic is ~0.1 % of peak -- reported under roofline.hbm for completeness).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "acg-alp-ldpc_b200"))
from ldpc_b200 import load_rows  # noqa: E402

SEED = 239239239
BP_SNR, BP_ITERS = -5.0, 100
ADMM_SNR, ADMM_ITERS, ADMM_ALPHA, ADMM_MU = -3.0, 1000, 1.2, 0.55
# algorithmic fp64 instructions per unit of work (DESIGN.md "rooflines"):
# BP in the likelihood-ratio domain: per edge and iteration
BP_FP64_PER_EDGE_ITER = 11.0          # check: 5.3 Pe/Po recurrences + 4 division; variable: 2.7 products + decision
BP_SMEM_BYTES_PER_EDGE_ITER = 32.0    # each message is read and written once per pass (8 B), two passes
BP_SMEM_BYTES_PER_VAR_ITER = 8.0      # channel likelihood ratio
# DRAM traffic per frame from the ncu --set full captures of the two kernels (dram__bytes_read.sum + dram__bytes_write.sum
# over the frames of the captured launch, profiles/r01_bp_lr_final_ncu.txt / r01_admm_chk_final_ncu.txt): the y samples
# (n x 8 B = 2240 B for n = 280) and nothing else -- the outputs stay in L2 until after the kernel; scaled by n / 280 for the
# (3,6)-1008 code (its captures, profiles/r01_bp_lr_1008_ncu.txt / r01_admm_chk_1008_ncu.txt: 8.1 KB per frame)
NCU_DRAM_BYTES_PER_FRAME = {"bp": 10.724e6 / 4736, "qpadmm": 5.396e6 / 2368}
# QP-ADMM, check-centric kernel: per three-variable block and iteration
ADMM_FP64_PER_BLOCK_ITER = 47.0       # 9 residual + 25 row updates (6 per row, 7 for row 3) + 6.5 auxiliary update + 6.5 variable update


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured", d.get("sm_max_mhz", 1965.0)
    return 6650.0, "fallback", 1965.0


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  In-process NVML from a thread (three light queries
    every 50 ms); an `nvidia-smi -lms` child is the fall-back -- its polls take the driver's lock long enough to show up
    as an occasional +10 % step on 0.4 s timed regions (profiles/r01_bench_notes.txt)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.nvml, self.stop_flag = [], None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
            except Exception:
                pass
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if uuid else b"")
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        bits = [("hw_slowdown", getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4))]
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.rows.append((time.perf_counter(), [str(sm), str(self.max_sm), ""] +
                                  ["Active" if mask & b else "Not Active" for _, b in bits]))
            except Exception:
                pass
            time.sleep(0.05)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def wait_ready(self, timeout=3.0):
        """blocks until the first sample has arrived (the sampler's start-up is over), at most `timeout` seconds"""
        t_end = time.perf_counter() + timeout
        while (self.proc or self.nvml) and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.01)

    def stop(self, t0, t1):
        if not self.proc and not self.nvml:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi"]}
        time.sleep(0.15)
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi"}


# ----------------------------------------------------------------- reference arm

def _ref_worker(args):
    """one single-threaded process of the UNMODIFIED reference (race-free)"""
    algo, code_name, frames, begin = args
    from oracle.oracle import Oracle, Ref
    ref, orc = Ref(), Oracle()
    H = load_rows(code_name)
    n = H.shape[1]
    if algo == "bp":
        y = orc.channel(SEED, begin, frames, n, BP_SNR)
        t0 = time.perf_counter()
        _, ok, secs = ref.bp_decode(H, y, BP_SNR, BP_ITERS)
    else:
        y = orc.channel(SEED, begin, frames, n, ADMM_SNR)
        t0 = time.perf_counter()
        _, ok, secs = ref.qpadmm_decode(H, y, ADMM_SNR, ADMM_ALPHA, ADMM_MU, ADMM_ITERS, 0.0)
    return time.perf_counter() - t0, secs, frames


def reference_sample(algo, code_name, frames_per_proc, procs, begin=0):
    """frames/s of the reference's CPU decoder on `procs` host cores (wall clock over all processes)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    jobs = [(algo, code_name, frames_per_proc, begin + i * frames_per_proc) for i in range(procs)]
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    inner = max(r[0] for r in res)
    return frames_per_proc * procs / inner, wall


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle import have_ref, build
    build(ref=True)
    cores = host_cores()
    kind = "reference" if have_ref() else "port"
    if not have_ref():
        emit({"impl": "reference", "unavailable": "oracle/_ref/libref_oracle.so was not built"})
        return
    per_proc = max(2, int(round(args.ref_seconds / 0.07)))          # ~0.07 s per BP(100) frame per core
    for _ in range(args.warmup):
        reference_sample("bp", "H05", 2, cores)
    t0 = time.perf_counter()
    vals = []
    for s in range(args.steps):
        v, _ = reference_sample("bp", "H05", per_proc, cores, begin=s * per_proc * cores)
        vals.append(v)
    elapsed = time.perf_counter() - t0
    value = float(np.mean(vals))
    sample = "%d frames/step = %d procs x %d frames, BP(100) H05 @ %g dB, 1 thread per process" % (
        per_proc * cores, cores, per_proc, BP_SNR)
    emit({
        "impl": "reference", "metric": "decoded frames/sec (BP, fixed 100 iters)", "value": value,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * elapsed / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f80", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": per_proc * cores,
                   "note": "the reference cannot switch its syndrome exit off; at -5 dB >= 99.8 % of its frames run all 100 iterations"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ----------------------------------------------------------------------- GPU arm

WORKLOAD = "BP(100 fixed iters, syndrome exit off) on H05 160x280, AWGN @ %g dB" % BP_SNR


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import ldpc_b200 as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        # the path's only collective is done by the LIBRARY's communicator (ldpc_allreduce_counters, NCCL over NVLink);
        # torch.distributed is plumbing: it hands the communicator id to the ranks and provides the barrier
        cid = torch.from_numpy(L.Comm.make_id() if rank == 0 else np.zeros(128, np.uint8)).to(dev)
        dist.broadcast(cid, 0)
        comm = L.Comm(rank, world, cid.cpu().numpy(), local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allreduce(counters):
        return comm.allreduce(counters) if comm else counters

    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    hbm_peak, peak_kind, _ = peaks()
    fp64_peak = L.measure_fp64_peak(local)            # G fp64 FMA instr/s, measured on this GPU
    smem_peak = L.measure_smem_peak(local)            # GB/s of conflict-free shared-memory loads, measured

    def bench_algo(algo, code_name, frames, steps, warmup, snr=None, n_iter=None, e2e=True):
        H = load_rows(code_name)
        m, n = H.shape
        code = L.Code(H=H, device=local)
        snr = (BP_SNR if algo == "bp" else ADMM_SNR) if snr is None else snr
        n_iter = (BP_ITERS if algo == "bp" else ADMM_ITERS) if n_iter is None else n_iter
        # inputs resident in HBM: two alternating batches, each larger than L2 (126 MB) for BP
        ys = []
        for b in range(2):
            y = torch.empty((frames, n), dtype=torch.float64, device=dev)
            begin = (rank * 2 + b) * frames
            code.channel_device(SEED, begin, frames, snr, y.data_ptr(), 0, sptr)
            ys.append(y)
        bits = torch.empty((frames, n), dtype=torch.uint8, device=dev)
        ok = torch.empty(frames, dtype=torch.uint8, device=dev)
        iters = torch.empty(frames, dtype=torch.int32, device=dev)

        def step_device(i):
            y = ys[i & 1]
            if algo == "bp":
                code.bp_decode_device(y.data_ptr(), frames, snr, n_iter, False, bits.data_ptr(), ok.data_ptr(),
                                      iters.data_ptr(), 0, sptr)
            else:
                code.qpadmm_decode_device(y.data_ptr(), frames, snr, ADMM_ALPHA, ADMM_MU, n_iter, 0.0,
                                          bits.data_ptr(), ok.data_ptr(), iters.data_ptr(), 0, sptr)

        # the clock sampler starts BEFORE the warm-up: its start-up (NVML initialisation over all GPUs of the box) must
        # not fall into the timed region
        sampler = ClockSampler(local) if rank == 0 else None
        for i in range(warmup):
            step_device(i)
        int(ok.sum().item()) + int(iters.sum().item())       # the reductions of the timed region are loaded and warm too
        allreduce({"ok": 0, "iters": 0})
        torch.cuda.synchronize()
        if sampler:
            sampler.wait_ready()
        # which kernel served the launches (testing hooks of the C ABI): the headline must not come from a fall-back
        if algo == "bp":
            kernel = {1: "bp_lr_kernel", 2: "bp_kernel (log domain)"}[L.last_bp_kernel()]
            assert L.last_bp_kernel() == 1, "the log-domain BP kernel served the benchmark code"
        else:
            kernel = {1: "qpadmm_chk_kernel", 2: "qpadmm_kernel (block per lane)"}[L.last_qpadmm_kernel()]
            assert L.last_qpadmm_kernel() == 1, "the block-per-lane QP-ADMM kernel served the benchmark code"
        gc.collect()                  # nothing of the previous measurement (pinned buffers, code handles) is released ...
        gc.disable()                  # ... inside the timed region
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_host0 = time.perf_counter()
        e0.record(stream)
        for i in range(steps):
            step_device(i)
        e1.record(stream)
        t_a = time.perf_counter()
        # the path's only collective, inside the timed region: decoder-flag and iteration counters of the last step,
        # summed over the ranks by the library's NCCL communicator
        mine = {"ok": int(ok.sum().item()), "iters": int(iters.sum().item())}
        t_b = time.perf_counter()
        counts = allreduce(mine)
        t_c = time.perf_counter()
        barrier()
        t_host1 = time.perf_counter()
        if os.environ.get("BENCH_TRACE"):
            sys.stderr.write("trace %s %s: launches %.1f ms, kernels+sums %.1f ms, all-reduce %.1f ms, barrier %.1f ms\n" % (
                algo, code_name, 1e3 * (t_a - t_host0), 1e3 * (t_b - t_a), 1e3 * (t_c - t_b), 1e3 * (t_host1 - t_c)))
        gc.enable()
        clocks = sampler.stop(t_host0, t_host1) if sampler else None
        dev_ms = max_over_ranks(e0.elapsed_time(e1))
        coll_ms = max_over_ranks(1e3 * (t_c - t_b))               # the all-reduce, host-synchronous
        raw_wall_ms = max_over_ranks(1e3 * (t_host1 - t_host0))
        # The timed step = the launches (CUDA events on the launching stream, as the contract asks) + the collective
        # (wall clock: it runs on the library's communicator stream and returns when the sums are on the host).  The raw
        # wall clock of the whole region is reported next to it: on these boxes it occasionally carries a host-side stall of
        # 50-600 ms between the last kernel and the return of the first reduction (profiles/r02_bench_notes.txt).
        wall_ms = dev_ms + coll_ms
        assert int(iters.min().item()) == n_iter, "fixed-iteration mode broken"
        assert counts["iters"] == world * frames * n_iter

        info = code.info
        e2e_res = None
        if e2e:
            # end to end through the public host-buffer API: pinned y in, bits/ok/iters out, every step
            y_pin = torch.empty((frames, n), dtype=torch.float64).pin_memory()
            y_pin.copy_(ys[0].cpu())
            b_pin = torch.empty((frames, n), dtype=torch.uint8).pin_memory()
            ok_pin = torch.empty(frames, dtype=torch.uint8).pin_memory()
            it_pin = torch.empty(frames, dtype=torch.int32).pin_memory()
            lib = L.lib()

            def step_e2e():
                if algo == "bp":
                    st = lib.ldpc_bp_decode(code._h, y_pin.data_ptr(), frames, snr, n_iter, 0, b_pin.data_ptr(),
                                            ok_pin.data_ptr(), it_pin.data_ptr(), None)
                else:
                    st = lib.ldpc_qpadmm_decode(code._h, y_pin.data_ptr(), frames, snr, ADMM_ALPHA, ADMM_MU, n_iter, 0.0,
                                                b_pin.data_ptr(), ok_pin.data_ptr(), it_pin.data_ptr(), None)
                assert st == 0, lib.ldpc_last_error()

            e2e_steps = max(1, min(steps, 3))
            step_e2e()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                step_e2e()
            barrier()
            e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
            e2e_res = {"value": world * frames * e2e_steps / (e2e_ms * 1e-3), "unit": "frames/s",
                       "h2d_bytes_per_step": frames * n * 8, "d2h_bytes_per_step": frames * (n + 5)}
            del y_pin, b_pin, ok_pin, it_pin

        units = info["edges"] if algo == "bp" else info["admm_blocks"]
        per_unit = BP_FP64_PER_EDGE_ITER if algo == "bp" else ADMM_FP64_PER_BLOCK_ITER
        fps_gpu = frames * steps / (dev_ms * 1e-3)                      # this rank's kernel throughput (CUDA events)
        value = world * frames * steps / (wall_ms * 1e-3)               # whole job: launches (events) + all-reduce
        achieved = fps_gpu * n_iter * units * per_unit / 1e9            # G fp64 instr/s on one GPU
        bytes_per_frame = n * 8 + n + 1 + 4                             # y in, bits + flag + iteration count out
        k = info["n"] - info["m"]
        fp64 = {"achieved": achieved, "peak": fp64_peak, "unit": "Gop/s (fp64 instr)", "frac": achieved / fp64_peak,
                "peak_source": "ldpc_measure_fp64_peak on this GPU (8 independent DFMA chains/thread)",
                "work_per_launch": "%d frames x %d iters x %d %s x %.0f fp64 instr" % (
                    frames, n_iter, units, "edges" if algo == "bp" else "blocks", per_unit)}
        hbm = {"achieved": fps_gpu * bytes_per_frame / 1e9, "peak": hbm_peak, "unit": "GB/s",
               "frac": fps_gpu * bytes_per_frame / 1e9 / hbm_peak, "peak_kind": peak_kind}
        # DRAM bytes the launch must move (the ncu captures under profiles/ show dram__bytes within 2 % of it: y is read
        # once, the outputs are written once, every message stays on chip)
        traffic = frames * bytes_per_frame
        if algo == "bp":
            # messages never leave the SM: the bounding resource of the likelihood-ratio kernel is shared-memory
            # bandwidth (algorithmic bytes below), then the FP64 pipe; HBM carries 2.5 KB per frame
            smem_bytes = info["edges"] * BP_SMEM_BYTES_PER_EDGE_ITER + info["n"] * BP_SMEM_BYTES_PER_VAR_ITER
            got = fps_gpu * n_iter * smem_bytes / 1e9
            roof = {"bound": "smem", "achieved": got, "peak": smem_peak, "unit": "GB/s", "frac": got / smem_peak,
                    "traffic": traffic, "bytes_per_frame_iter": smem_bytes,
                    "peak_source": "ldpc_measure_smem_peak on this GPU (conflict-free LDS.128)",
                    "work_per_launch": "%d frames x %d iters x %.0f B of shared-memory traffic" % (frames, n_iter, smem_bytes),
                    "fp64": fp64, "hbm": hbm}
        else:
            roof = dict(fp64, bound="fp64", traffic=traffic, hbm=hbm)
            # shared-memory traffic of the check-centric kernel per frame-iteration: the row terms w are written once
            # per block (32 B) and gathered once per (original variable, check) incidence (32 B); v is written once
            # per original variable and gathered once per incidence (8 B); q + alpha/2 and inv_coef are read per variable
            inc = info["edges"]
            smem_bytes = info["admm_blocks"] * 32 + inc * 32 + inc * 8 + info["n"] * 24
            got = fps_gpu * n_iter * smem_bytes / 1e9
            roof["smem"] = {"achieved": got, "peak": smem_peak, "unit": "GB/s", "frac": got / smem_peak,
                            "bytes_per_frame_iter": smem_bytes,
                            "peak_source": "ldpc_measure_smem_peak on this GPU (conflict-free LDS.128)"}
        res = {
            "value": value, "ms_per_step": wall_ms / steps, "kernel_ms_per_step": dev_ms / steps,
            "allreduce_ms": coll_ms, "wall_ms_per_step": raw_wall_ms / steps, "kernel_value": world * fps_gpu,
            "info_gbit_per_s": value * k / 1e9, "roofline": roof, "clocks": clocks, "frames_per_step": frames,
            "iters": n_iter, "snr_db": snr, "mean_ok": counts["ok"] / (world * frames), "kernel": kernel,
        }
        if e2e_res:
            res["e2e"] = e2e_res
        code.close()
        del ys, bits, ok, iters
        torch.cuda.empty_cache()
        return res

    def as_run(algo, code_name, snr, frames, max_iter):
        """experiment mode (device-side codewords, AWGN, decoding with the reference's stopping rules, verdict,
        counters): frames/s as the decoders really run, with the mean iterations per frame"""
        H = load_rows(code_name)
        code = L.Code(H=H, device=local)
        dec = L.BeliefPropagationDecoder(max_iter) if algo == "bp" else L.QPADMMDecoder(ADMM_ALPHA, ADMM_MU, max_iter, 1e-5)
        begin = rank * frames
        code.experiment(dec, snr, SEED, begin, min(frames, 4096))            # warm-up (tables, schedules)
        barrier()
        t0 = time.perf_counter()
        r = code.experiment(dec, snr, SEED, begin, frames)
        tot = allreduce({k: r[k] for k in L.CNT_NAMES})
        barrier()
        wall = max_over_ranks(time.perf_counter() - t0)
        code.close()
        return {"workload": "%s on %s @ %g dB, early exit as the reference" % (dec.name(), code_name, snr),
                "frames": world * frames, "value": world * frames / wall, "unit": "frames/s",
                "kernel_seconds": max_over_ranks(r["gpu_seconds"]), "mean_iters": tot["sum_iters"] / tot["total"],
                "fer": 1.0 - tot["correct"] / tot["total"]}

    def channel_roofline(code_name, frames):
        """device Philox4x32-10 + Box-Muller channel (utils/channel.h:19-26): y written to HBM, 8 n bytes per frame"""
        H = load_rows(code_name)
        n = H.shape[1]
        code = L.Code(H=H, device=local)
        y = torch.empty((frames, n), dtype=torch.float64, device=dev)
        code.channel_device(SEED, 0, frames, -3.0, y.data_ptr(), 0, sptr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(3):
            code.channel_device(SEED, (i + 1) * frames, frames, -3.0, y.data_ptr(), 0, sptr)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        code.close()
        gbs = frames * n * 8 / (ms * 1e-3) / 1e9
        return {"workload": "channel_kernel: %d frames x %d samples (Philox4x32-10 + fp64 Box-Muller), y to HBM" % (frames, n),
                "frames_per_s": frames / (ms * 1e-3), "samples_per_s": frames * n / (ms * 1e-3),
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                             "peak_kind": peak_kind, "traffic": frames * n * 8}}

    def experiment_scaling(total_bp, total_admm):
        """BASELINE.json configs[3]: the (3,6)-1008 code, a FIXED total of frames sharded over the ranks by global frame
        index through ldpc_experiment_run, counter blocks all-reduced by NCCL inside the timed wall.  The counters are
        functions of the global frame range only: identical for 1/2/4/8 GPUs (compare `counters` across the SCALE lines);
        each run also checks the invariance itself on a small range (sharded + all-reduced == one rank alone)."""
        import sharding
        H = load_rows("reg_3_6_1008")
        code = L.Code(H=H, device=local)
        out = {}
        for name, dec, snr, total in (("bp", L.BeliefPropagationDecoder(100), -2.0, total_bp),
                                      ("qpadmm", L.QPADMMDecoder(ADMM_ALPHA, ADMM_MU, 1000, 1e-5), -1.0, total_admm)):
            small = 1536
            b, e = sharding.shard_range(small, rank, world)
            part = code.experiment(dec, snr, SEED, 10 ** 9 + b, e - b)
            summed = allreduce({k: part[k] for k in L.CNT_NAMES})
            alone = code.experiment(dec, snr, SEED, 10 ** 9, small)
            assert all(summed[k] == alone[k] for k in L.CNT_NAMES), "sharded counters differ from a single rank's"
            b, e = sharding.shard_range(total, rank, world)
            barrier()
            t0 = time.perf_counter()
            r = code.experiment(dec, snr, SEED, b, e - b)
            tot = allreduce({k: r[k] for k in L.CNT_NAMES})
            barrier()
            wall = max_over_ranks(time.perf_counter() - t0)
            out[name] = {"workload": "%s on reg_3_6_1008 @ %g dB, %d frames in total (strong scaling), sharded by global frame "
                                     "index, NCCL all-reduce of the counter block inside the timed region" % (dec.name(), snr, total),
                         "value": total / wall, "unit": "frames/s", "wall_s": wall, "scaling": "strong",
                         "kernel_seconds_max": max_over_ranks(r["gpu_seconds"]), "counters": tot,
                         "mean_iters": tot["sum_iters"] / tot["total"], "invariance_checked_on_frames": small}
        code.close()
        return out

    bp = bench_algo("bp", "H05", args.frames, args.steps, args.warmup)
    extra = {}
    if not args.headline_only:
        extra["bp_optimalH"] = bench_algo("bp", "optimalH", args.frames, args.steps, args.warmup, e2e=False)
        admm = bench_algo("qpadmm", "optimalH", max(1024, args.frames // 8), args.steps, args.warmup)
        # main.cpp:31 runs QP-ADMM with max_iter = 10000: the same kernel, ten times the iterations per frame
        extra["qpadmm_10000"] = bench_algo("qpadmm", "optimalH", max(1024, args.frames // 64), max(1, args.steps // 2), 1,
                                           n_iter=10000, e2e=False)
        # the other code north_star names: synthetic (3,6)-regular n = 1008 (BASELINE.json configs[3]), same settings
        big_bp = bench_algo("bp", "reg_3_6_1008", max(1024, args.frames // 4), args.steps, args.warmup)
        big_admm = bench_algo("qpadmm", "reg_3_6_1008", max(512, args.frames // 32), args.steps, args.warmup)
        as_run_pts = [as_run("bp", "H05", -3.0, args.frames, 100), as_run("bp", "H05", 0.0, 4 * args.frames, 100),
                      as_run("qpadmm", "optimalH", -3.0, max(1024, args.frames // 8), 10000),
                      as_run("qpadmm", "optimalH", 0.0, args.frames, 10000)]
        chan = channel_roofline("H05", args.frames) if rank == 0 else None
        barrier()
        exp_scale = experiment_scaling(args.exp_frames, max(2048, args.exp_frames // 16))

    exp_only = None
    if args.headline_only and args.with_experiment:
        exp_only = experiment_scaling(args.exp_frames, max(2048, args.exp_frames // 16))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.oracle import have_ref
        if have_ref():
            cores = host_cores()
            per_proc = max(2, int(round(args.ref_seconds / 0.07)))
            v, wall = reference_sample("bp", "H05", per_proc, cores)
            cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": "reference",
                   "sample": "%d procs x %d frames, unmodified reference BP(100) on H05 @ %g dB, 1 thread/process, "
                             "%.1f s wall" % (cores, per_proc, BP_SNR, wall)}
    if rank == 0:
        pick = ("value", "ms_per_step", "kernel_ms_per_step", "wall_ms_per_step", "allreduce_ms", "kernel_value", "info_gbit_per_s", "roofline", "clocks",
                "kernel", "frames_per_step", "iters", "snr_db")
        line = {
            "metric": "decoded frames/sec (BP, fixed 100 iters)", "value": bp["value"], "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": bp["ms_per_step"],
            "kernel_ms_per_step": bp["kernel_ms_per_step"], "wall_ms_per_step": bp["wall_ms_per_step"],
            "allreduce_ms": bp["allreduce_ms"], "kernel_value": bp["kernel_value"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": args.frames,
                       "steps_for_1e7_frames_per_point": -(-10 ** 7 // args.frames),
                       "parallelism": "frames sharded over %d GPU(s)" % world,
                       "timed_region": "barrier, K decode launches (CUDA events on the launching stream), NCCL all-reduce of the "
                                       "counters (ldpc_allreduce_counters, wall clock), barrier; max over ranks; value = work / (events + "
                                       "all-reduce); wall_ms_per_step = raw wall clock of the region; kernel_* = the launches alone",
                       "l2_policy": "inputs larger than L2: two alternating %d MB batches" % (args.frames * 280 * 8 >> 20)},
            "info_gbit_per_s": bp["info_gbit_per_s"], "roofline": bp["roofline"], "e2e": bp["e2e"],
            "clocks": bp["clocks"], "gpu_launches": args.steps, "kernel": bp["kernel"], "cpu_baseline": cpu,
        }
        if not args.headline_only:
            line["bp_optimalH"] = dict({k: extra["bp_optimalH"][k] for k in pick},
                                       workload="BP(100 fixed iters) on optimalH 160x280 @ %g dB" % BP_SNR)
            line["qpadmm"] = dict({k: admm[k] for k in pick + ("e2e",)},
                                  metric="decoded frames/sec (QP-ADMM, fixed 1000 iters, eps_stop=0)",
                                  config={"workload": "QP-ADMM(alpha=%g, mu=%g, 1000 iters) on optimalH 160x280 @ %g dB" % (
                                      ADMM_ALPHA, ADMM_MU, ADMM_SNR), "frames_per_step_per_gpu": admm["frames_per_step"]})
            line["qpadmm_10000"] = dict({k: extra["qpadmm_10000"][k] for k in pick},
                                        workload="QP-ADMM(alpha=%g, mu=%g, 10000 fixed iters, main.cpp:31) on optimalH @ %g dB" % (
                                            ADMM_ALPHA, ADMM_MU, ADMM_SNR))
            line["reg_3_6_1008"] = {
                "config": {"workload": "synthetic (3,6)-regular 504x1008 (configs[3]): BP(100 fixed iters) @ %g dB, %d frames/step/GPU; "
                                       "QP-ADMM(alpha=%g, mu=%g, 1000 fixed iters) @ %g dB, %d frames/step/GPU" % (
                                           BP_SNR, big_bp["frames_per_step"], ADMM_ALPHA, ADMM_MU, ADMM_SNR,
                                           big_admm["frames_per_step"])},
                "bp": {k: big_bp[k] for k in pick + ("e2e",)},
                "qpadmm": {k: big_admm[k] for k in pick + ("e2e",)},
            }
            line["as_run"] = as_run_pts
            line["channel"] = chan
            line["experiment_scaling"] = exp_scale
        elif exp_only:
            line["experiment_scaling"] = exp_only
        emit(line)
    if comm:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line, on the process's real stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def guard_stdout():
    """Everything any library prints to file descriptor 1 during the run (NCCL's version banner, a stray warning of
    a child process) goes to stderr; only emit() reaches the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=1 << 20,
                    help="frames per step per GPU (BP); QP-ADMM uses 1/8.  A million frames (0.18 s per step) so that the "
                         "30-50 ms stalls these boxes show now and then (one run in five, whatever samples the clocks) "
                         "stay below a few per cent of the timed region")
    ap.add_argument("--ref-seconds", type=float, default=12.0, help="CPU work per core of one reference sample")
    ap.add_argument("--exp-frames", type=int, default=1 << 23,
                    help="total BP frames of the experiment-mode scaling arm on the (3,6)-1008 code (QP-ADMM: 1/16)")
    ap.add_argument("--headline-only", action="store_true", help="only the headline workload (BP on H05)")
    ap.add_argument("--with-experiment", action="store_true",
                    help="with --headline-only: also the experiment-mode strong-scaling arm (counters comparable across N)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
