"""CPU prototype of an EXACT tile-culled bidirectional nearest-neighbour search (design study for DESIGN.md section 8).

Rows (a skinned cloud) are grouped into k-d leaves of `Lr` points, columns (an observed frame, constant over the
optimisation) into k-d leaves of `Lc` points.  A (row leaf, column leaf) pair is skipped when the squared gap between the
two boxes is STRICTLY larger than both the largest running row minimum of the row leaf and the largest running column
minimum of the column leaf.  The gap is formed with the same expression as a point distance, so it is a lower bound of
every computed distance of the pair bit for bit (subtraction, multiplication and addition are monotone under rounding),
and the strict ">" keeps exact ties alive.

Schedule: a seeding pass evaluates, for every row leaf, its `seed` nearest column leaves and, for every column leaf, its
`seed` nearest row leaves (by box gap); then
  static  -- every other pair is tested against the bounds the seeding left (no further tightening: what a kernel gets
             with no communication between CTAs);
  dynamic -- row leaves are visited in turn, nearest column leaves first, bounds tightened after every evaluated pair.
Both reproduce the brute-force minima exactly (asserted).  Row order: k-d leaves of the POSED cloud ("posed") or one
fixed order from the canonical cloud ("cano").

    python scripts/culling_prototype.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reart_b200.synth import make_sequence                                               # noqa: E402


def kd_leaves(p, leaf):
    """Permutation that groups p into k-d leaves of exactly `leaf` points (median split on the longest axis)."""
    out = []

    def rec(ix):
        if len(ix) <= leaf:
            out.append(ix)
            return
        ax = np.ptp(p[ix], 0).argmax()
        o = ix[np.argsort(p[ix, ax], kind="stable")]
        h = (-(-len(o) // leaf) // 2) * leaf
        rec(o[:h])
        rec(o[h:])

    rec(np.arange(len(p)))
    return np.concatenate(out)


def d2(a, b):
    dx = a[:, None, 0] - b[None, :, 0]
    dy = a[:, None, 1] - b[None, :, 1]
    dz = a[:, None, 2] - b[None, :, 2]
    return (dx * dx + dy * dy) + dz * dz


def box_gap2(alo, ahi, blo, bhi):
    g = np.maximum(np.float32(0), np.maximum(alo[:, None] - bhi[None], blo[None] - ahi[:, None])).astype(np.float32)
    return (g[..., 0] * g[..., 0] + g[..., 1] * g[..., 1]) + g[..., 2] * g[..., 2]


def brute(a, b):
    rows = np.full(len(a), np.inf, np.float32)
    cols = np.full(len(b), np.inf, np.float32)
    for s in range(0, len(a), 1024):
        d = d2(a[s:s + 1024], b)
        rows[s:s + 1024] = d.min(1)
        cols = np.minimum(cols, d.min(0))
    return rows, cols


def culled(a, b, Lr, Lc, seed, dynamic):
    A, B = a.reshape(-1, Lr, 3), b.reshape(-1, Lc, 3)
    gap = box_gap2(A.min(1), A.max(1), B.min(1), B.max(1))
    nr, nc = gap.shape
    rows = np.full((nr, Lr), np.inf, np.float32)
    cols = np.full((nc, Lc), np.inf, np.float32)
    done = np.zeros((nr, nc), bool)

    def evaluate(i, j):
        if not done[i, j]:
            d = d2(A[i], B[j])
            rows[i] = np.minimum(rows[i], d.min(1))
            cols[j] = np.minimum(cols[j], d.min(0))
            done[i, j] = True

    by_row, by_col = np.argsort(gap, axis=1, kind="stable"), np.argsort(gap, axis=0, kind="stable")
    for i in range(nr):
        for j in by_row[i, :seed]:
            evaluate(i, j)
    for j in range(nc):
        for i in by_col[:seed, j]:
            evaluate(i, j)
    seeded = done.mean()
    if dynamic:
        for i in range(nr):
            for j in by_row[i]:
                if not (gap[i, j] > rows[i].max() and gap[i, j] > cols[j].max()):
                    evaluate(i, j)
    else:
        rb, cb = rows.max(1).copy(), cols.max(1).copy()
        for i, j in zip(*np.nonzero(~((gap > rb[:, None]) & (gap > cb[None, :])))):
            evaluate(i, j)
    return rows.reshape(-1), cols.reshape(-1), done.mean(), seeded


def main():
    seq = make_sequence(T=4, N=16384, P=15, seed=2)
    cano, part, frames, pose = seq["cano"], seq["part"], seq["frames"], seq["pose"]
    t = 0
    posed = (np.einsum("nij,nj->ni", pose[t, part, :3, :3], cano) + pose[t, part, :3, 3]).astype(np.float32)
    print("state      row order  (Lr, Lc, seed)   seeding   static   dynamic   [% of leaf pairs evaluated]")
    for state, src in (("identity", cano), ("converged", posed)):
        for row_mode in ("posed", "cano"):
            if state == "identity" and row_mode == "cano":
                continue                                         # same thing: the posed cloud IS the canonical one
            for Lr, Lc, seed in ((256, 256, 2), (256, 64, 2), (256, 32, 2), (128, 32, 2)):
                a = src[kd_leaves(src if row_mode == "posed" else cano, Lr)]
                b = frames[t][kd_leaves(frames[t], Lc)]
                r0, c0 = brute(a, b)
                out = []
                for dynamic in (False, True):
                    r, c, frac, seeded = culled(a, b, Lr, Lc, seed, dynamic)
                    assert np.array_equal(r, r0) and np.array_equal(c, c0)
                    out.append(100 * frac)
                print(f"{state:10s} {row_mode:9s}  {str((Lr, Lc, seed)):15s} {100 * seeded:6.1f}   {out[0]:6.1f}   {out[1]:6.1f}",
                      flush=True)


if __name__ == "__main__":
    main()
