"""Every launch shape of the likelihood-ratio BP kernel (frames per team x teams per CTA x soft output) against the
log-domain kernel on the same channel samples: flags, bits and iteration counts must be identical.

    python acg-alp-ldpc_b200/tools/bp_shape_matrix.py      (on the GPU box; prints one line per shape)
"""
import os, sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/acg-alp-ldpc_b200')
import ldpc_b200 as L
from ldpc_b200 import load_rows
for name in ("optimalH", "H05", "reg_3_6_1008"):
    code = L.Code(H=load_rows(name))
    for frames in (301, 6000):
        snr = -3.0 if name != "reg_3_6_1008" else -1.5
        y = code.channel(239239239, 4000, frames, snr)
        os.environ["LDPC_BP_KERNEL"] = "log"; os.environ["LDPC_BP_F"] = "4"
        rb, rok, rit, rp = code.bp_decode(y, snr, 100)
        os.environ["LDPC_BP_KERNEL"] = "lr"
        for F in ("2", "4", "8"):
            for teams in ("1", "2", "5"):
                for soft in (True, False):
                    os.environ["LDPC_BP_F"] = F; os.environ["LDPC_BP_TEAMS"] = teams
                    b, ok, it, p = code.bp_decode(y, snr, 100, soft=soft)
                    bad = np.flatnonzero((b != rb).any(1) | (ok != rok) | (it != rit))
                    print(name, frames, "F", F, "teams", teams, "soft", soft, "mismatches", len(bad), bad[:8], flush=True)
