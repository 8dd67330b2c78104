"""numpy model of the DISTRIBUTED selection of the top g kd levels planned for the multi-GPU rebuild (DESIGN.md section
6b, item 1): every rank only ever looks at its own particles; what crosses ranks is small (bounding boxes, 2048-bin
histograms, the particles tied with a pivot key, one minimum per segment).  Test helper: tests/test_dist_top_model.py
checks that it reproduces the oracle's partition and boxes, ties included, so that the CUDA version of round 2 has a
pinned algorithm to follow.  Follows the single-GPU build of csrc/kdtree.cu (3-pass radix select 11 + 11 + 10 bits,
ties ranked by the previous split axes and then by the input index; fmm_cart3_kdtree.cuh:89-202 for the geometry)."""
import numpy as np

NO_AXIS = 3


def ordered_bits(f):
    u = np.asarray(f, np.float32).view(np.uint32)
    return np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.uint32)


def unordered_bits(k):
    k = np.uint32(k)
    u = (k & np.uint32(0x7FFFFFFF)) if (k & np.uint32(0x80000000)) else ~k
    return np.array([u], np.uint32).view(np.float32)[0]


def seg_start(n, i, l):
    return 0 if i <= 0 else ((n * i - 1) >> l) + 1


def widest_axis(d):
    dx, dy, dz = d
    return (0 if dx > dz else 2) if dx > dy else (1 if dy > dz else 2)


def chain_push(axis, parent_chain):
    c, k = [axis], 1
    for a in parent_chain:
        if a != NO_AXIS and a != axis and k < 3:
            c.append(a)
            k += 1
    return c + [NO_AXIS] * (3 - len(c))


def distributed_top(pos, g, world):
    """pos (n,3) float32; rank r holds indices [seg_start(n,r,log2 world), ...).  Returns (dest segment of every particle at
    level g, {node: (lbound, rbound, axis)} for levels 0..g, bytes exchanged per rank)."""
    n = len(pos)
    gw = world.bit_length() - 1
    own = [np.arange(seg_start(n, r, gw), seg_start(n, r + 1, gw)) for r in range(world)]
    keys = [ordered_bits(pos[:, a]) for a in range(3)]
    exchanged = 0
    # bounding box: one min/max all-reduce
    lo = np.min([pos[o].min(0) for o in own], 0)
    hi = np.max([pos[o].max(0) for o in own], 0)
    exchanged += 24
    boxes = {0: (lo.copy(), hi.copy(), widest_axis(hi - lo), chain_push(widest_axis(hi - lo), [NO_AXIS] * 3))}
    seg = np.zeros(n, np.int64)   # segment of every particle at the current level (known to its owner only)
    for l in range(g):
        new_seg = seg.copy()
        for i in range(1 << l):
            node = (1 << l) - 1 + i
            lb, rb, axis, chain = boxes[node]
            k = seg_start(n, 2 * i + 1, l + 1) - seg_start(n, 2 * i, l + 1) - 1   # rank of the pivot inside the segment
            mine = [o[seg[o] == i] for o in own]
            kk = [keys[axis][m] for m in mine]
            # 3-pass radix select: the ranks exchange one histogram per pass
            prefix, krem, less = np.uint32(0), k, 0
            for shift, bits in ((21, 11), (10, 11), (0, 10)):
                hist = np.zeros(1 << bits, np.int64)
                for q in kk:
                    sel = q if shift == 21 else q[(q >> np.uint32(shift + bits)) == (prefix >> np.uint32(shift + bits))]
                    hist += np.bincount((sel >> np.uint32(shift)) & np.uint32((1 << bits) - 1), minlength=1 << bits)
                exchanged += 4 * (1 << bits)
                cum = np.cumsum(hist)
                b = int(np.searchsorted(cum, krem, side="right"))
                before = int(cum[b - 1]) if b else 0
                prefix |= np.uint32(b << shift)
                less += before
                krem -= before
                eq = int(hist[b])
            need = krem + 1
            go_left = [q < prefix for q in kk]
            if eq != need:
                # tied with the pivot key on both sides: the ranks pool the tied particles (few) and rank them by the
                # rest of the total order
                tied = np.concatenate([m[q == prefix] for m, q in zip(mine, kk)])
                exchanged += 16 * len(tied)
                cols = [tied]
                for a in reversed([c for c in chain[1:] if c != NO_AXIS]):
                    cols.append(keys[a][tied])
                order = np.lexsort(cols)          # last key = most recent previous axis, then older, then index
                left_tied = set(tied[order[:need]].tolist())
                go_left = [gl | np.isin(m, list(left_tied)) & (q == prefix) for gl, m, q in zip(go_left, mine, kk)]
            else:
                go_left = [gl | (q == prefix) for gl, q in zip(go_left, kk)]
            # first key of the right child: one min all-reduce
            rmin = min([int(q[~gl].min()) if (~gl).any() else 0xFFFFFFFF for q, gl in zip(kk, go_left)])
            exchanged += 4
            for m, gl in zip(mine, go_left):
                new_seg[m[gl]] = 2 * i
                new_seg[m[~gl]] = 2 * i + 1
            lrb = rb.copy(); lrb[axis] = unordered_bits(prefix)
            rlb = lb.copy(); rlb[axis] = unordered_bits(rmin)
            for child, (cl, cr) in ((2 * node + 1, (lb.copy(), lrb)), (2 * node + 2, (rlb, rb.copy()))):
                ax = widest_axis(cr - cl)
                boxes[child] = (cl, cr, ax, chain_push(ax, chain))
        seg = new_seg
    return seg, boxes, exchanged
