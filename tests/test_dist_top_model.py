"""The distributed top-level selection planned for the multi-GPU rebuild (tests/dist_top_model.py) reproduces the oracle's
partition and node boxes -- including inputs full of tied coordinates -- while exchanging only kilobytes per rank.  CPU only."""
import numpy as np
import pytest

import coulomb_oscillators_b200 as nb
from dist_top_model import distributed_top, seg_start
from refs import Oracle


def quantised(n, seed):
    """coordinates on a coarse lattice: most split planes cut through groups of equal keys"""
    rng = np.random.default_rng(seed)
    return (rng.integers(-6, 7, size=(n, 3)) * np.float32(0.125)).astype(np.float32)


@pytest.mark.parametrize("n,world,kind", [(20000, 8, "ga"), (12345, 4, "ga"), (9000, 8, "lattice"), (4099, 2, "lattice"), (8192, 8, "cube")])
def test_distributed_top_matches_oracle(n, world, kind):
    pos = {"ga": lambda: nb.init_ga(n)[0], "cube": lambda: nb.init_test_cube(n)[0], "lattice": lambda: quantised(n, n)}[kind]()
    g = world.bit_length() - 1
    dest, boxes, exchanged = distributed_top(pos, g, world)
    orc = Oracle(order=3, unsort=0)
    orc.fmm3_kd(pos.copy(), None, None)
    T = orc.tree()
    assert T["levels"] >= g
    for i in range(1 << g):
        want = np.sort(T["perm"][seg_start(n, i, g):seg_start(n, i + 1, g)])
        assert np.array_equal(np.flatnonzero(dest == i), want), i
    for node, (lb, rb, axis, _) in boxes.items():
        assert np.array_equal(lb, T["lbound"][node]) and np.array_equal(rb, T["rbound"][node]) and axis == T["splitdim"][node], node
    # histograms (7 segments x 3 passes) + boxes + the tied particles; never the positions themselves.  The lattice input is
    # adversarial (13 distinct coordinates per axis: every pivot is tied with hundreds of particles), there only correctness counts
    if kind != "lattice":
        assert exchanged < 150_000
