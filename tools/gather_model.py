"""Offline model of the neighbour gathers of the list kernels (k_rho / k_push) on a real particle state (numpy/scipy,
no GPU).  Input: positions dumped by tests/dump_bench_state.py (the bench scene family after a few steps).

What it reports, per particle ordering / list-slot order / CTA shape:
  wavefronts/gather  L1 data-stage wavefronts of one warp-wide gather of 32-byte records IF every sector hits: per group of
                     four adjacent lanes, the largest number of distinct records that share a 32-byte position of their
                     128-byte lines (the rule tools/probe/gather_probe.cu measured on B200: P2/P10 8.1, P3 14.1, P4 19.2)
  lines/gather       distinct 128-byte lines touched by one warp-wide gather
  max hit %          1 - distinct sectors / sector accesses of one CTA: what the L1 can give if nothing is evicted
  broadcast          lane-candidate evaluations per true pair of the "warp = 32 targets, every lane reads the same staged
                     candidate" formulation (candidates = union of the targets' 5x5x5 stencils)
Usage: python tools/gather_model.py /tmp/bench_state.npz
"""
import sys

import numpy as np
from scipy.spatial import cKDTree


def wavefront_model(rows, n):
    cap = max(len(a) for a in rows)
    R = np.full((n, cap), -1, dtype=np.int64)
    for i, a in enumerate(rows):
        R[i, : len(a)] = a
    nw = n // 32
    tot_wf = tot_g = tot_lines = act = 0
    for w in range(nw):
        blk = R[w * 32:(w + 1) * 32]
        for k in range(cap):
            col = blk[:, k]
            m = col >= 0
            if not m.any():
                continue
            tot_g += 1
            act += int(m.sum())
            for q in range(8):
                cc = col[q * 4:(q + 1) * 4]
                cc = cc[cc >= 0]
                if cc.size:
                    tot_wf += int(np.bincount(np.unique(cc) & 3, minlength=4).max())
            tot_lines += np.unique(col[m] >> 2).size
    return tot_wf / tot_g, tot_lines / tot_g, act / tot_g, tot_g / nw


def main():
    z = np.load(sys.argv[1] if len(sys.argv) > 1 else "/tmp/bench_state.npz")
    x, r = z["x"], float(z["radius"])
    h, cell = 4 * r, 2 * r
    n = len(x)
    c = np.floor((x - (x.min(0) - 1e-9)) / cell).astype(np.int64)
    dims = c.max(0) + 1

    def rows_key(cc):
        return (cc[:, 2] * dims[1] + cc[:, 1]) * dims[0] + cc[:, 0]

    def brick_key(b):
        def f(cc):
            nbk = [(dims[k] + b[k] - 1) // b[k] for k in range(3)]
            q, rr = cc // np.array(b), cc % np.array(b)
            return ((q[:, 2] * nbk[1] + q[:, 1]) * nbk[0] + q[:, 0]) * (b[0] * b[1] * b[2]) + (rr[:, 2] * b[1] + rr[:, 1]) * b[0] + rr[:, 0]
        return f

    def ordered(keyfn):
        p = np.argsort(keyfn(c), kind="stable")
        xs, cs = x[p], c[p]
        nb = cKDTree(xs).query_ball_point(xs, h * (1 - 1e-12))
        return xs, cs, [np.sort([j for j in a if j != i]).astype(np.int64) for i, a in enumerate(nb)]

    xs, cs, rows = ordered(rows_key)
    print(f"{n} particles, mean fluid neighbours {np.mean([len(a) for a in rows]):.1f}, rows of {dims[0]} cells")
    print("slot order (row-major particle order, all sectors assumed to hit):")
    print("  index-sorted (what the kernels use): wavefronts/gather %.2f  lines/gather %.2f  active lanes %.1f  slots/warp %.1f" % wavefront_model(rows, n))

    def core_first(i, a):
        d = cs[a] - cs[i]
        core = np.abs(d).max(1) <= 1
        off = ((d[:, 2] + 2) * 5 + (d[:, 1] + 2)) * 5 + (d[:, 0] + 2)
        k = np.lexsort((a, off))
        return np.concatenate([a[k][core[k]], a[k][~core[k]]])

    def fixed_slots(i, a):
        d = cs[a] - cs[i]
        core = np.abs(d).max(1) <= 1
        off = ((d[:, 2] + 1) * 3 + (d[:, 1] + 1)) * 3 + (d[:, 0] + 1)
        out, extra = np.full(27, -1, dtype=np.int64), []
        for j, o, cf in zip(a, off, core):
            if cf and out[o] < 0:
                out[o] = j
            else:
                extra.append(j)
        return np.concatenate([np.delete(out, 13), np.array(extra, dtype=np.int64)])

    print("  3x3x3 core cells first, by cell offset:   wavefronts/gather %.2f  lines/gather %.2f  active lanes %.1f  slots/warp %.1f"
          % wavefront_model([core_first(i, a) if len(a) else a for i, a in enumerate(rows)], n))
    print("  one fixed slot per core cell offset:      wavefronts/gather %.2f  lines/gather %.2f  active lanes %.1f  slots/warp %.1f"
          % wavefront_model([fixed_slots(i, a) if len(a) else a for i, a in enumerate(rows)], n))

    print("particle order / CTA size (index-sorted slots):")
    cases = [("row-major", rows_key, 128), ("row-major", rows_key, 256), ("row-major", rows_key, 512), ("row-major", rows_key, 1024),
             ("bricks 8x4x4", brick_key((8, 4, 4)), 128), ("bricks 4x4x4", brick_key((4, 4, 4)), 128), ("bricks 16x4x2", brick_key((16, 4, 2)), 128),
             ("bricks 8x4x4", brick_key((8, 4, 4)), 256)]
    cache = {}
    for name, fn, block in cases:
        if name not in cache:
            _, _, rw = ordered(fn)
            cache[name] = (rw, wavefront_model(rw, n))
        rw, (wf, lines, _, _) = cache[name]
        tot = dsec = 0
        for b0 in range(0, n - block + 1, block):
            allj = np.concatenate(rw[b0:b0 + block])
            tot += allj.size
            dsec += np.unique(allj).size
        print(f"  {name:14s} {block:5d} particles per CTA: max hit {100 * (1 - dsec / tot):4.1f} %   wavefronts/gather {wf:5.2f}  lines/gather {lines:5.2f}")

    print("broadcast formulation (exact distance test of every staged candidate by every target lane):")
    cell_count = np.bincount(rows_key(c), minlength=int(dims.prod()))
    rng = np.random.default_rng(0)

    def inflation(keyfn):
        _, cs2, rw = ordered(keyfn)
        tot_c = tot_p = 0
        for w in rng.choice(n // 32, size=min(200, n // 32), replace=False):
            cells = set()
            for cx, cy, cz in set(map(tuple, cs2[w * 32:(w + 1) * 32])):
                for dz in range(-2, 3):
                    for dy in range(-2, 3):
                        for dx in range(-2, 3):
                            a, b, d = cx + dx, cy + dy, cz + dz
                            if 0 <= a < dims[0] and 0 <= b < dims[1] and 0 <= d < dims[2]:
                                cells.add((d * dims[1] + b) * dims[0] + a)
            tot_c += 32 * sum(int(cell_count[k]) for k in cells)
            tot_p += sum(len(rw[i]) for i in range(w * 32, (w + 1) * 32))
        return tot_c / tot_p

    print(f"  warp = 32 consecutive particles of a row:  {inflation(rows_key):5.1f} lane-candidate evaluations per true pair")
    print(f"  warp = the particles of a 4x4x2-cell brick: {inflation(brick_key((4, 4, 2))):5.1f}")


if __name__ == "__main__":
    main()
