"""prints the interesting numbers of a bench.py JSON line:  python tools/show_bench.py gpurun_out/x.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
print("headline %.4g frames/s (kernel %.4g, e2e %.4g), roofline %.3f, n_gpus %d, clocks %s" % (
    d["value"], d["kernel_value"], d["e2e"]["value"], d["roofline"]["frac"], d["n_gpus"], d["clocks"]))
print("cpu_baseline", d.get("cpu_baseline"))
for k in ("bp_optimalH", "qpadmm", "qpadmm_10000"):
    if k in d:
        print(k, "%.4g frames/s, frac %.3f, smem frac %s" % (d[k]["value"], d[k]["roofline"]["frac"],
                                                             d[k]["roofline"].get("smem", {}).get("frac")))
if "reg_3_6_1008" in d:
    for k in ("bp", "qpadmm"):
        r = d["reg_3_6_1008"][k]
        print("1008", k, "%.4g frames/s, frac %.3f, smem frac %s, e2e %.4g" % (
            r["value"], r["roofline"]["frac"], r["roofline"].get("smem", {}).get("frac"), r["e2e"]["value"]))
for p in d.get("as_run", []):
    print("as run:", p["workload"], "%.4g frames/s, mean iters %.1f, FER %.4g" % (p["value"], p["mean_iters"], p["fer"]))
if d.get("channel"):
    print("channel: %.4g frames/s, HBM frac %.3f" % (d["channel"]["frames_per_s"], d["channel"]["roofline"]["frac"]))
for k, v in (d.get("experiment_scaling") or {}).items():
    print("experiment", k, "%.4g frames/s, wall %.3f s, mean iters %.1f, counters %s" % (
        v["value"], v["wall_s"], v["mean_iters"], v["counters"]))
