"""CPU estimate of what exact tile culling could save in the Chamfer search on benchmark-like data.

Both clouds are reordered into k-d leaves of `leaf` points (median split along the longest axis); a (row leaf, column
leaf) pair has to be evaluated only if the distance between the two bounding boxes does not exceed the largest
nearest-neighbour distance of the rows in the row leaf or of the columns in the column leaf (bounds taken from the
exact answer here, i.e. the best any bound-propagation scheme can do).  Prints the fraction of leaf pairs that survive.

    python scripts/culling_potential.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reart_b200.synth import make_sequence                                               # noqa: E402


def kd_order(p, leaf):
    out = []

    def rec(ix):
        if len(ix) <= leaf:
            out.append(ix)
            return
        ax = np.ptp(p[ix], 0).argmax()
        o = ix[np.argsort(p[ix, ax], kind="stable")]
        rec(o[:len(o) // 2])
        rec(o[len(o) // 2:])

    rec(np.arange(len(p)))
    return np.concatenate(out)


def nn_d2(a, b):
    out = np.empty(len(a), np.float32)
    for s in range(0, len(a), 2048):
        out[s:s + 2048] = ((a[s:s + 2048, None, :] - b[None]) ** 2).sum(-1).min(1)
    return out


def main():
    seq = make_sequence(T=4, N=16384, P=15, seed=2)
    cano, part, frames, pose = seq["cano"], seq["part"], seq["frames"], seq["pose"]
    for state in ("identity pose (first iteration)", "ground-truth pose (converged)"):
        for leaf in (256, 128, 64, 32):
            frac = []
            for t in range(2):
                src = cano
                if state.startswith("ground"):
                    src = np.einsum("nij,nj->ni", pose[t, part, :3, :3], cano) + pose[t, part, :3, 3]
                s, g = src[kd_order(src, leaf)], frames[t][kd_order(frames[t], leaf)]
                sb, gb = s.reshape(-1, leaf, 3), g.reshape(-1, leaf, 3)
                gap = np.maximum(0, np.maximum(sb.min(1)[:, None] - gb.max(1)[None], gb.min(1)[None] - sb.max(1)[:, None]))
                gap = (gap * gap).sum(-1)
                rows, cols = nn_d2(s, g).reshape(-1, leaf).max(1), nn_d2(g, s).reshape(-1, leaf).max(1)
                frac.append(((gap <= rows[:, None]) | (gap <= cols[None, :])).mean())
            print(f"{state:34s} leaf {leaf:4d}: {100 * np.mean(frac):5.1f} % of leaf pairs must be evaluated")


if __name__ == "__main__":
    main()
