"""Randomised fills through z2d_submit against the CPU oracle, byte for byte.

A batch of a few hundred painter.fill calls mixes what the spec scenes show one at a time: several sub-paths per call, lines and
curves, repeated points, two-point "polygons" (move_to, line_to, close_path: the unpaired-crossing quirk of
fill_plotter.zig:78-81 / multisample.zig:156), empty sub-paths, the trailing move_to Path.close leaves behind, shapes hanging
over the surface edge, both fill rules, several tolerances and operators.  Some calls qualify for the node-parallel flattening
kernels and some do not, and one batch holds both (the parallel recorder takes runs of plain fills; here runs are short, so
both recorders are exercised by the two batch sizes).
"""
import numpy as np
import pytest

from tests.fuzz_util import assert_scene_matches
from z2d_b200 import abi, workloads
from z2d_b200.abi import AntiAliasMode, Format, NodeTag

pytestmark = pytest.mark.gpu

SIZE = 320


def fuzz_scene(seed, n_paths, aa, ops=None):
    rng = np.random.default_rng(seed)
    tags, pts, node_off = [], [], [0]

    def q(v):
        return float(np.round(v * 16) / 16)

    for _ in range(n_paths):
        for _sub in range(int(rng.integers(1, 4))):
            x, y = rng.uniform(-20, SIZE + 20, 2)
            x0, y0 = x, y
            tags.append(int(NodeTag.move_to)); pts.append((q(x), q(y), 0, 0, 0, 0))
            for _k in range(int(rng.choice([0, 1, 2, 3, 3, 4, 6, 9]))):
                kind = rng.integers(0, 10)
                if kind == 0:
                    tags.append(int(NodeTag.line_to)); pts.append((q(x), q(y), 0, 0, 0, 0))
                    continue
                step = rng.uniform(-70, 70, (3, 2))
                if kind < 6:
                    x, y = x + step[0, 0], y + step[0, 1]
                    if kind == 1:
                        y = pts[-1][1] if tags[-1] != int(NodeTag.curve_to) else pts[-1][5]  # horizontal edge
                    tags.append(int(NodeTag.line_to)); pts.append((q(x), q(y), 0, 0, 0, 0))
                else:
                    c1 = (x + step[0, 0], y + step[0, 1])
                    c2 = (c1[0] + step[1, 0], c1[1] + step[1, 1])
                    x, y = c2[0] + step[2, 0], c2[1] + step[2, 1]
                    tags.append(int(NodeTag.curve_to)); pts.append((q(c1[0]), q(c1[1]), q(c2[0]), q(c2[1]), q(x), q(y)))
            tags.append(int(NodeTag.close_path)); pts.append((0, 0, 0, 0, 0, 0))  # painter.fill requires closed sub-paths
            if rng.integers(0, 2) == 0:
                tags.append(int(NodeTag.move_to)); pts.append((q(x0), q(y0), 0, 0, 0, 0))
        node_off.append(len(tags))
    nodes = np.zeros(len(tags), dtype=workloads.NODE_DT)
    nodes["tag"] = np.array(tags, dtype=np.uint32)
    nodes["p"] = np.array(pts, dtype=np.float64)
    patterns = workloads._premultiplied_colours(rng, n_paths)
    fo = np.zeros(n_paths, dtype=workloads.FILLOPTS_DT)
    fo["anti_aliasing_mode"] = int(aa)
    fo["fill_rule"] = rng.integers(0, 2, n_paths)
    fo["op"] = rng.choice(ops or ([int(abi.Operator.src_over)] * 6 + [int(abi.Operator.src), int(abi.Operator.xor), int(abi.Operator.multiply),
                                                                      int(abi.Operator.dst_out)]), n_paths)
    fo["precision"] = int(abi.Precision.integer)
    fo["tolerance"] = rng.choice([0.1, 0.1, 0.01, 0.5, 2.0], n_paths)
    kind = np.zeros(n_paths, dtype=np.uint32)
    return workloads.Scene(SIZE, SIZE, nodes, np.array(node_off, dtype=np.int64), patterns, kind, fo,
                           np.zeros(0, dtype=workloads.STROKEOPTS_DT), np.arange(n_paths, dtype=np.int64))


# Every call is compared: a seed is never excused as a whole.  Calls with a dangling edge (two-point "polygons") and calls with an
# unbounded operator and anti-aliasing none depend on the ORDER the reference's per-scanline sort leaves equal crossings in
# (Polygon.zig:275-353, multisample.zig:156, direct.zig:112-124); the device replays that loop for exactly those calls
# (k_edge_sim), and the oracle restates it with a stable sort -- identical to the reference's pdq sort while a scanline holds at
# most 12 active edges (insertion sort); beyond that the reference's tie order is not pinned by anything it ships.
SEEDS = list(range(31, 55))
CASES = [pytest.param(seed, 3000 if seed == 34 else 300, aa, id=f"{seed}-{aa.name}")
         for seed in SEEDS for aa in [AntiAliasMode.default, AntiAliasMode.none, AntiAliasMode.supersample_4x]]


@pytest.mark.parametrize("seed,n_paths,aa", CASES)
def test_random_fills_match_oracle(cuda, seed, n_paths, aa):
    scene = fuzz_scene(seed, n_paths, aa)
    assert_scene_matches(cuda, scene, min_covered=SIZE * SIZE // 10)


@pytest.mark.parametrize("seed", [61, 62, 63, 64, 65, 66])
def test_random_fills_unbounded_operators_without_aa(cuda, seed):
    """direct.zig with src_in / dst_in / src_out / dst_atop: every span pair clears the rest of its row (row records)."""
    scene = fuzz_scene(seed, 120, AntiAliasMode.none, ops=[int(abi.Operator.src_over)] * 3 + [int(abi.Operator.src_in), int(abi.Operator.dst_in),
                                                                                             int(abi.Operator.src_out), int(abi.Operator.dst_atop)])
    assert_scene_matches(cuda, scene)
