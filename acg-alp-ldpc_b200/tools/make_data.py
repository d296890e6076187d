"""Writes the parity-check matrices of BASELINE.json's configs in the sparse
``.rows`` text format used by this repo (one line per check: degree, then the
ascending column indices).

    python acg-alp-ldpc_b200/tools/make_data.py [--reference /root/reference]

* optimalH / H05 / G05 are read from the reference's dense comma-separated files
  with its own parsing rule (utils/parse_data.h:15-21: a cell is 1 iff its last
  character is '1') and re-emitted as index lists -- no file is copied.
  ``make data`` (acg-alp-ldpc_b200/Makefile) expands them back to the dense
  ``data/*.txt`` layout that read_pcm() and `make run` expect.
* reg_3_6_1008 is the synthetic (3,6)-regular n = 1008 code of SURVEY.md 8:
  configuration model from a fixed seed, double edges and 4-cycles removed by
  edge swaps, regenerated until H has full row rank (GetOrtogonal needs that).
"""
import argparse
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(os.path.dirname(HERE), "data")


def read_dense(path):
    rows = []
    for tok in open(path).read().split():
        cells = tok[:-1].split(",") if tok.endswith(",") else tok.split(",")
        rows.append([1 if (c and c[-1] == "1") else 0 for c in cells])
    return np.array(rows, np.uint8)


def write_rows(path, H, comment):
    with open(path, "w") as f:
        f.write("# %s\n" % comment)
        f.write("%d %d\n" % H.shape)
        for r in H:
            idx = np.flatnonzero(r)
            f.write(" ".join(str(x) for x in [len(idx)] + list(idx)) + "\n")


def gf2_rank(H):
    A = H.copy().astype(np.uint8)
    m, n = A.shape
    rank = 0
    for col in range(n):
        piv = np.flatnonzero(A[rank:, col])
        if len(piv) == 0:
            continue
        p = piv[0] + rank
        if p != rank:
            A[[rank, p]] = A[[p, rank]]
        others = np.flatnonzero(A[:, col])
        others = others[others != rank]
        A[others] ^= A[rank]
        rank += 1
        if rank == m:
            break
    return rank


def four_cycles(H):
    """pairs of checks sharing >= 2 variables"""
    ov = H.astype(np.int32) @ H.T.astype(np.int32)
    np.fill_diagonal(ov, 0)
    return np.argwhere(np.triu(ov) >= 2)


def regular_code(n, dv, dc, seed):
    m = n * dv // dc
    rng = np.random.default_rng(seed)
    attempt = 0
    while True:
        attempt += 1
        sockets = np.repeat(np.arange(n), dv)
        rng.shuffle(sockets)
        checks = np.repeat(np.arange(m), dc)
        edges = list(zip(checks.tolist(), sockets.tolist()))
        for _ in range(200):
            H = np.zeros((m, n), np.int32)
            for c, v in edges:
                H[c, v] += 1
            bad = [i for i, (c, v) in enumerate(edges) if H[c, v] > 1]
            cyc = four_cycles((H > 0).astype(np.uint8))
            if not bad and len(cyc) == 0:
                break
            # swap the variable ends of offending edges with random other edges
            victims = set(bad)
            for a, b in cyc:
                shared = np.flatnonzero((H[a] > 0) & (H[b] > 0))
                for i, (c, v) in enumerate(edges):
                    if c == a and v == shared[0]:
                        victims.add(i)
                        break
            for i in victims:
                j = int(rng.integers(len(edges)))
                (ci, vi), (cj, vj) = edges[i], edges[j]
                edges[i], edges[j] = (ci, vj), (cj, vi)
        else:
            continue
        Hb = (H > 0).astype(np.uint8)
        if (Hb.sum(0) == dv).all() and (Hb.sum(1) == dc).all() and gf2_rank(Hb) == m:
            return Hb, attempt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    os.makedirs(DATA, exist_ok=True)
    for name in ("optimalH", "H05", "G05"):
        src = os.path.join(args.reference, "data", name + ".txt")
        H = read_dense(src)
        write_rows(os.path.join(DATA, name + ".rows"), H,
                   "%s: %dx%d, %d ones (index lists of the reference's data/%s.txt)" % (name, H.shape[0], H.shape[1],
                                                                                          int(H.sum()), name))
        print(name, H.shape, int(H.sum()))
    H, attempt = regular_code(1008, 3, 6, seed=239239239)
    write_rows(os.path.join(DATA, "reg_3_6_1008.rows"), H,
               "synthetic (3,6)-regular n=1008 code, girth >= 6, full rank; make_data.py seed 239239239 attempt %d"
               % attempt)
    print("reg_3_6_1008", H.shape, int(H.sum()), "attempt", attempt)


if __name__ == "__main__":
    main()
