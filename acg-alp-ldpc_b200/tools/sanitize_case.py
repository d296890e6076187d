"""Small batches through every kernel and launch shape, for compute-sanitizer (one tool per run):

    compute-sanitizer --tool racecheck python acg-alp-ldpc_b200/tools/sanitize_case.py [bp|admm|all]

Both BP kernels (likelihood-ratio: 2 / 4 / 8 / 16 frames per team, several teams per CTA, soft output on and off, the
generic-degree variant; log domain), both QP-ADMM kernels (check-centric: 1 / 2 / 4 frames per CTA, two CTAs per SM,
grid mode; block per lane), decode and experiment mode, slot refill with early exit.  Sizes are tiny: the sanitizer
slows the kernels down by two orders of magnitude."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ldpc_b200 as L  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
SEED = 239239239


def env(**kw):
    for k in [k for k in os.environ if k.startswith("LDPC_")]:
        del os.environ[k]
    for k, v in kw.items():
        os.environ[k] = str(v)


codes = {name: L.Code(H=L.load_rows(name)) for name in ("optimalH", "H05", "reg_3_6_1008")}
rng = np.random.default_rng(5)
wide = np.zeros((24, 60), np.uint8)
for r, d in enumerate([3, 4, 5, 6, 7, 8, 9, 12] * 3):
    wide[r, rng.choice(60, size=d, replace=False)] = 1
for v in np.flatnonzero(wide.sum(0) == 0):
    wide[rng.integers(24), v] = 1
codes["wide"] = L.Code(H=wide)

if which in ("bp", "all"):
    for F in (2, 4, 8, 16):
        for name, frames in (("optimalH", 45), ("reg_3_6_1008", 9)):
            if F == 16 and name == "reg_3_6_1008":
                continue
            for teams in (1, 3):
                env(LDPC_BP_F=F, LDPC_BP_TEAMS=teams)
                c = codes[name]
                y = c.channel(SEED, 0, frames, -2.0)
                for soft in (True, False):
                    c.bp_decode(y, -2.0, 12, soft=soft)
                c.bp_decode(y, -2.0, 6, early_exit=False, soft=False)
                c.experiment(L.BeliefPropagationDecoder(10), -1.0, SEED, 0, frames)
                print("bp lr F=%d teams=%d %s ok" % (F, teams, name), flush=True)
    env()
    y = codes["wide"].channel(SEED, 0, 40, 0.0)
    codes["wide"].bp_decode(y, 0.0, 10)
    print("bp lr generic degrees ok", flush=True)
    env(LDPC_BP_KERNEL="log", LDPC_BP_F=4)
    y = codes["optimalH"].channel(SEED, 0, 30, -2.0)
    codes["optimalH"].bp_decode(y, -2.0, 8)
    codes["optimalH"].experiment(L.BeliefPropagationDecoder(8), -1.0, SEED, 0, 30)
    print("bp log-domain ok", flush=True)

if which in ("admm", "all"):
    for F in (1, 2, 4):
        env(LDPC_ADMM_F=F)
        c = codes["optimalH"]
        y = c.channel(SEED, 0, 22, -2.0)
        c.qpadmm_decode(y, -2.0, 1.2, 0.55, 40, 1e-5)
        c.qpadmm_decode(y, -2.0, 1.2, 0.55, 15, 0.0, soft=False)
        c.experiment(L.QPADMMDecoder(1.2, 0.55, 30, 1e-5), 0.0, SEED, 0, 22)
        print("admm check-centric F=%d ok" % F, flush=True)
    env()
    c = codes["reg_3_6_1008"]
    y = c.channel(SEED, 0, 5, 0.0)
    c.qpadmm_decode(y, 0.0, 1.2, 0.55, 12, 1e-5)                       # two CTAs per SM, 64 registers
    env(LDPC_ADMM_TWO=0)
    c.qpadmm_decode(y, 0.0, 1.2, 0.55, 12, 1e-5)
    print("admm check-centric (3,6)-1008 ok", flush=True)
    env()
    codes["optimalH"].qpadmm_grid(np.array([0.0, 1.2, 3.0]), np.array([0.5, 0.55, 0.7]), -3.0, 20, 1e-5, SEED, 0, 17)
    codes["wide"].qpadmm_decode(codes["wide"].channel(SEED, 0, 12, 0.0), 0.0, 1.2, 0.55, 20, 1e-5)
    print("admm grid mode and wide checks ok", flush=True)
    for F, kb in ((1, 2), (2, 4), (4, 8)):
        env(LDPC_ADMM_KERNEL="block", LDPC_ADMM_F=F, LDPC_ADMM_KB=kb)
        c = codes["optimalH"]
        y = c.channel(SEED, 0, 14, -2.0)
        c.qpadmm_decode(y, -2.0, 1.2, 0.55, 20, 1e-5)
        c.experiment(L.QPADMMDecoder(1.2, 0.55, 20, 1e-5), 0.0, SEED, 0, 14)
        print("admm block-per-lane F=%d KB=%d ok" % (F, kb), flush=True)
    env()
    c = codes["optimalH"]
    c.channel(SEED, 5, 8, 0.0)
print("done", flush=True)
