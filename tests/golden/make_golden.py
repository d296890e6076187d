"""Generates the golden fixtures in tests/golden/ from the UNMODIFIED reference
(oracle/_ref/libref_oracle.so, built from /root/reference by oracle/Makefile).

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these are
outputs of the reference itself, run single-threaded in this container:
codewords from gen_random_codewords(G, ., mt19937(239239239)) (main.cpp:63-64),
channel words from transmit(snr, c, mt19937(frame index + 1)) (experiment.h:90-99),
decoded by BeliefPropagationDecoder(100) and QPADMMDecoder(alpha, mu, 1000, 1e-5)
with the alpha/mu main.cpp:28-34 pairs with each matrix.  The .npz files travel to
the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.helpers import GOLDEN, load_rows  # noqa: E402
from oracle.oracle import Ref  # noqa: E402

FRAMES = 24
SNRS = [-3.0, -1.0]
ADMM = {"optimalH": (1.2, 0.55), "H05": (1.95, 0.5)}


def main():
    ref = Ref()
    for name in ("optimalH", "H05"):
        H = load_rows(name)
        G = ref.get_orthogonal(H)[0] if name == "optimalH" else load_rows("G05")
        cw = ref.gen_random_codewords(G, FRAMES, 239239239)
        out = {"H_name": name, "snrs": np.array(SNRS), "codewords": cw, "alpha_mu": np.array(ADMM[name])}
        for si, snr in enumerate(SNRS):
            y = np.stack([ref.transmit(snr, cw[i], i + 1) for i in range(FRAMES)])
            bp_bits, bp_ok, _ = ref.bp_decode(H, y, snr, 100)
            a, mu = ADMM[name]
            ad_bits, ad_ok, _ = ref.qpadmm_decode(H, y, snr, a, mu, 1000, 1e-5)
            out["y_%d" % si] = y
            out["bp_bits_%d" % si] = bp_bits
            out["bp_ok_%d" % si] = bp_ok
            out["admm_bits_%d" % si] = ad_bits
            out["admm_ok_%d" % si] = ad_ok
            print(name, snr, "BP ok", int(bp_ok.sum()), "ADMM frames == codeword",
                  int((ad_bits == cw).all(1).sum()))
        # the reference's Monte-Carlo harness, one thread: counters of 60 frames at -3 dB
        cw60 = ref.gen_random_codewords(G, 60, 239239239)
        for algo in ("bp", "qpadmm"):
            a, mu = ADMM[name]
            res = ref.experiment(algo, H, cw60, -3.0, 100 if algo == "bp" else 1000, a, mu, 1e-5)
            res.pop("time_us")
            out["exp_" + algo] = np.array([res[k] for k in sorted(res)], np.int64)
            out["exp_keys"] = np.array(sorted(res))
            print(name, algo, res)
        out["exp_codewords"] = cw60
        out["exp_y"] = np.stack([ref.transmit(-3.0, cw60[i], i + 1) for i in range(60)])
        np.savez_compressed(os.path.join(GOLDEN, "ref_%s.npz" % name), **out)
    # GetOrtogonal / read_pcm fixtures: G of optimalH as index lists hash + a few codewords
    H = load_rows("optimalH")
    G = ref.get_orthogonal(H)[0]
    np.savez_compressed(os.path.join(GOLDEN, "ref_generator_optimalH.npz"), G=np.packbits(G, axis=1),
                        n=np.array(G.shape[1]), first_words=ref.gen_random_codewords(G, 8, 239239239))


if __name__ == "__main__":
    main()
