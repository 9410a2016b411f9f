"""Offline model of the neighbour-gather access pattern for different particle orderings (numpy/scipy, no GPU).

For a disordered fluid block (jittered 2r lattice, support radius 4r) it sorts the particles by several cell orderings,
builds the index-sorted neighbour rows the ELL lists hold, and reports per ordering
  lines/gather : distinct 128-byte lines (4 x 32-byte records) touched by one warp-wide gather of the k-th neighbour
  reuse/block  : record accesses per distinct line within one thread block (upper bound of what L1 can give)
Usage: python tools/gather_locality.py [n_side] [jitter]
"""
import sys

import numpy as np
from scipy.spatial import cKDTree


def order_rows(c, dims):
    return (c[:, 2] * dims[1] + c[:, 1]) * dims[0] + c[:, 0]


def order_bricks(c, dims, b):
    bx, by, bz = b
    nb = [(dims[k] + b[k] - 1) // b[k] for k in range(3)]
    q = c // np.array(b)
    r = c % np.array(b)
    brick = (q[:, 2] * nb[1] + q[:, 1]) * nb[0] + q[:, 0]
    inner = (r[:, 2] * by + r[:, 1]) * bx + r[:, 0]
    return brick * (bx * by * bz) + inner


def morton(c):
    def spread(v):
        o = np.zeros_like(v)
        for k in range(10):
            o |= ((v >> k) & 1) << (3 * k)
        return o
    return spread(c[:, 0]) | (spread(c[:, 1]) << 1) | (spread(c[:, 2]) << 2)


def analyse(x, key, h, block=128):
    perm = np.argsort(key, kind="stable")
    xs = x[perm]
    tree = cKDTree(xs)
    nb = tree.query_ball_point(xs, h * 0.999999)
    n = len(xs)
    cap = max(len(a) for a in nb)
    rows = np.full((n, cap), -1, dtype=np.int64)
    for i, a in enumerate(nb):
        a = np.sort([j for j in a if j != i])
        rows[i, : len(a)] = a
    lines = rows // 4
    lines[rows < 0] = -1
    nw = n // 32
    tot_lines = 0
    tot_gathers = 0
    for w in range(nw):
        blk = lines[w * 32:(w + 1) * 32]
        for k in range(cap):
            col = blk[:, k]
            col = col[col >= 0]
            if col.size:
                tot_lines += np.unique(col).size
                tot_gathers += 1
    reuse = []
    for b0 in range(0, n - block + 1, block):
        blk = lines[b0:b0 + block]
        v = blk[blk >= 0]
        reuse.append(v.size / np.unique(v).size)
    return tot_lines / tot_gathers, float(np.mean(reuse)), float(np.mean([len(a) - 1 for a in nb]))


def main():
    ns = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    jit = float(sys.argv[2]) if len(sys.argv) > 2 else 0.35
    r = 0.025
    d = 2 * r
    h = 4 * r
    rng = np.random.default_rng(0)
    g = np.stack(np.meshgrid(*[np.arange(ns)] * 3, indexing="ij"), -1).reshape(-1, 3) * d
    x = g + rng.uniform(-jit * d, jit * d, g.shape)
    cell = h / 2 * (1 + 1e-7)
    c = np.floor((x - x.min(0) + 1e-9) / cell).astype(np.int64)
    dims = c.max(0) + 1
    cfull = np.floor((x - x.min(0) + 1e-9) / (h * (1 + 1e-7))).astype(np.int64)
    dfull = cfull.max(0) + 1
    cases = {
        "rows, cell h/2 (current)": order_rows(c, dims),
        "rows, cell h": order_rows(cfull, dfull),
        "morton, cell h/2": morton(c),
        "bricks 4x4x4, cell h/2": order_bricks(c, dims, (4, 4, 4)),
        "bricks 4x4x2, cell h/2": order_bricks(c, dims, (4, 4, 2)),
        "bricks 8x4x4, cell h/2": order_bricks(c, dims, (8, 4, 4)),
        "bricks 8x2x2, cell h/2": order_bricks(c, dims, (8, 2, 2)),
        "bricks 4x2x2, cell h/2": order_bricks(c, dims, (4, 2, 2)),
        "bricks 8x8x8, cell h/2": order_bricks(c, dims, (8, 8, 8)),
    }
    print(f"{len(x)} particles, jitter {jit}")
    for name, key in cases.items():
        for blk in (128, 256):
            l, ru, nn = analyse(x, key, h, blk)
            print(f"{name:28s} block {blk}: lines/gather {l:5.2f}  reuse/block {ru:5.2f}  mean nbrs {nn:.1f}")


if __name__ == "__main__":
    main()
