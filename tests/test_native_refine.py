"""The native builder state and refineCell (include/loco_grid.hpp through its C ABI include/loco_grid.h and
locointerpolate_b200/grid.py) against the compiled reference
(oracle/_ref: LCAdaptiveGrid::bootstrapSample / refineCell, LCAdaptiveGrid.cpp:256-456, 629-914) on the CPU.

Grids are compared as SETS -- samples by centre, basis functions by (centre, support) key, per-cell sample sets, boxes, centre
values -- with weights to 1e-12 relative: the reference iterates unordered containers, so sample numbering and summation
order are not defined by it (the weights are sums of dyadic fractions and come out bit-identical in practice)."""
import numpy as np
import pytest

from locointerpolate_b200 import capi, refine
from locointerpolate_b200.grid import HostGrid
from oracle import oracleapi, refapi

pytestmark = pytest.mark.skipif(not refapi.available(), reason="oracle/_ref (the compiled reference) is not built")


def _key(x):
    return tuple(np.round(np.asarray(x, dtype=np.float64) * 2.0 ** 30).astype(np.int64))


def canonical(g):
    """plain-array graph -> order-free description"""
    samples = {}
    for i in range(len(g.sample_center)):
        fns = {}
        for b in range(g.bf_offset[i], g.bf_offset[i + 1]):
            k = _key(g.bf_center[b]) + _key(g.bf_support[b])
            assert k not in fns, "a sample holds two functions of one key"
            fns[k] = g.bf_weight[b]
        k = _key(g.sample_center[i])
        assert k not in samples, "two samples at one centre"
        samples[k] = (g.sample_value[i].copy(), fns)
    cells = []
    for c in range(len(g.cell_split_dir)):
        leaf = g.cell_split_dir[c] < 0
        members = sorted(_key(g.sample_center[s]) for s in g.cell_sample[g.cell_sample_offset[c]:g.cell_sample_offset[c + 1]]) if leaf else []
        cells.append((tuple(g.cell_ranges[c]), int(g.cell_split_dir[c]), members, tuple(g.cell_center_value[c]) if leaf else None))
    border = sorted((tuple(g.border_ranges[c]), sorted(_key(g.sample_center[s]) for s in g.border_sample[g.border_sample_offset[c]:g.border_sample_offset[c + 1]]))
                    for c in range(len(g.border_ranges)))
    return samples, cells, border


def assert_same_grid(ga, gb, tag=""):
    sa, ca, ba = canonical(ga)
    sb, cb, bb = canonical(gb)
    assert set(sa) == set(sb), (tag, "sample sets differ", len(sa), len(sb))
    worst = 0.0
    for k in sa:
        (va, fa), (vb, fb) = sa[k], sb[k]
        assert np.array_equal(va, vb), (tag, "sample value", k)
        assert set(fa) == set(fb), (tag, "basis functions of the sample at", np.array(k) / 2.0 ** 30, len(fa), len(fb))
        for x in fa:
            worst = max(worst, abs(fa[x] - fb[x]) / abs(fa[x]))
    assert worst < 1e-12, (tag, worst)
    assert len(ca) == len(cb), (tag, "number of tree cells")
    for i, (x, y) in enumerate(zip(ca, cb)):
        assert x == y, (tag, "tree cell", i, x[0], y[0], len(x[2]), len(y[2]))
    assert ba == bb, (tag, "border cells")


@pytest.mark.parametrize("N,basis,boot,eval_corners", [(2, capi.LINEAR, 2, True), (2, capi.CUBIC, 2, True), (2, capi.CUBIC, 1, False),
                                                       (3, capi.CUBIC, 1, False), (3, capi.LINEAR, 2, True), (1, capi.CUBIC, 3, True),
                                                       (4, capi.LINEAR, 1, True), (3, capi.CUBIC, 2, True)])
def test_native_bootstrap_equals_the_reference_builder(N, basis, boot, eval_corners):
    ref = refapi.RefGrid(N, basis, False, 0, 0.5, boot, eval_corners=eval_corners)
    g0 = oracleapi.graph_from_ref(ref.sampled_function().graph())
    nat = HostGrid.bootstrap(N, basis, boot, g0.domain, ref.true_values, eval_corners=eval_corners)
    g1 = nat.graph()
    assert_same_grid(g0, g1)
    # beyond the set comparison: the same sample ORDER (samples_ order = proto ids), border-cell order and value sharing
    assert np.array_equal(g0.sample_center, g1.sample_center) and np.array_equal(g0.sample_value, g1.sample_value)
    first = lambda grp: np.array([np.flatnonzero(grp == v)[0] for v in grp])
    assert np.array_equal(first(g0.sample_value_group), first(g1.sample_value_group))
    assert np.array_equal(g0.border_ranges, g1.border_ranges)
    fn = ref.sampled_function()
    for leaf in range(nat.num_leaves()):   # neighbour lists of the live builder: leaves + border cells
        assert sum(nat.leaf_neighbors(leaf)) == fn.num_neighbors(leaf), leaf


@pytest.mark.parametrize("N,basis,boot,steps,seed,start", [
    (2, capi.LINEAR, 2, 40, 3, "graph"), (2, capi.CUBIC, 2, 30, 5, "graph"), (2, capi.CUBIC, 2, 30, 6, "bootstrap"),
    (2, capi.LINEAR, 1, 30, 8, "bootstrap"), (2, capi.CUBIC, 1, 20, 9, "graph"), (1, capi.CUBIC, 2, 10, 2, "bootstrap"),
    (3, capi.LINEAR, 1, 25, 7, "graph"), (3, capi.CUBIC, 1, 12, 9, "bootstrap"), (3, capi.LINEAR, 2, 15, 11, "bootstrap"), (4, capi.LINEAR, 1, 8, 13, "graph")])
def test_native_refine_cell_equals_the_reference(N, basis, boot, steps, seed, start):
    """the same random refineCell sequence (TestExamples.cpp:225-240 style) on the reference's grid and on the native one"""
    ref = refapi.RefGrid(N, basis, False, 0, 0.5, boot)
    g0 = oracleapi.graph_from_ref(ref.sampled_function().graph())
    nat = HostGrid.from_graph(g0, ref.true_values) if start == "graph" else HostGrid.bootstrap(N, basis, boot, g0.domain, ref.true_values)
    rng = np.random.default_rng(seed)
    for step in range(steps):
        L = ref.num_leaves()
        assert nat.num_leaves() == L
        leaf, direction = int(rng.integers(0, L)), int(rng.integers(0, N))
        ref.refine(leaf, direction)
        nat.refine(leaf, direction)
        if step % 3 == 2 or step == steps - 1:
            assert_same_grid(oracleapi.graph_from_ref(ref.sampled_function().graph()), nat.graph(), f"step {step}: leaf {leaf} direction {direction}")
    fn = ref.sampled_function()
    for leaf in range(nat.num_leaves()):
        assert sum(nat.leaf_neighbors(leaf)) == fn.num_neighbors(leaf), leaf


@pytest.mark.parametrize("N,basis,boot,steps", [(3, capi.LINEAR, 1, 25), (3, capi.CUBIC, 1, 12), (4, capi.LINEAR, 1, 10), (2, capi.CUBIC, 2, 40)])
def test_native_refinement_keeps_linear_precision(N, basis, boot, steps):
    """the property refineCell exists to preserve (TestExamples.cpp linearPrecisionTest): after any sequence of splits the
    interpolant of a LINEAR function is that function; evaluated with the CPU restatement on the native grid's graph"""
    ref = refapi.RefGrid(N, basis, True, 0, 0.5, boot)   # linear_fn=True: the function being sampled is linear
    g0 = oracleapi.graph_from_ref(ref.sampled_function().graph())
    nat = HostGrid.bootstrap(N, basis, boot, g0.domain, ref.true_values)
    rng = np.random.default_rng(11)
    dom = g0.domain.reshape(N, 2)
    for step in range(steps):
        nat.refine(int(rng.integers(0, nat.num_leaves())), int(rng.integers(0, N)))
    g = nat.graph()
    q = dom[:, 0] + rng.random((4000, N)) * (dom[:, 1] - dom[:, 0])
    out, _, status, _ = oracleapi.Oracle(g).eval(q)
    truth = ref.true_values(q)
    assert (status == 0).all()
    assert np.abs(out - truth).max() <= 1e-12 * max(1.0, np.abs(truth).max())
    # the flattener accepts the native graph (support sets, term lists) and numbers its leaves as the grid does
    m = capi.Model.from_graph(g, device=capi.HOST_ONLY)
    assert m.L == nat.num_leaves()


@pytest.mark.parametrize("basis,max_level,threshold", [(capi.LINEAR, 8, 0.05), (capi.CUBIC, 7, 0.1)])
def test_level_synchronous_loop_with_native_surgery(basis, max_level, threshold):
    """The refinement loop of SURVEY.md 8f.3 without the reference library in the loop: pass selection by
    locointerpolate_b200.refine, surgery by the native grid.  (The device error check is replaced by the CPU restatement's
    centre errors here; tests/cpp/test_facade.cpp runs the same loop through LCAdaptiveGrid::adaptiveSample on the GPU.)
    The same passes driven with the REFERENCE's refineCell give the same grid."""
    ref = refapi.RefGrid(2, basis, False, 0, threshold, 2)
    g0 = oracleapi.graph_from_ref(ref.sampled_function().graph())
    nat = HostGrid.bootstrap(2, basis, 2, g0.domain, ref.true_values)

    def model_of(graph_fn):
        def get_model():
            g = graph_fn()
            m = capi.Model.from_graph(g, device=capi.HOST_ONLY)
            orc = oracleapi.Oracle(g)
            m.center_error = lambda t: orc.center_error(t)   # stands in for loco_center_error
            return m
        return get_model

    log_nat = refine.adaptive_sample_passes(model_of(nat.graph), nat.refine, threshold, max_level)
    log_ref = refine.adaptive_sample_passes(model_of(lambda: oracleapi.graph_from_ref(ref.sampled_function().graph())), ref.refine, threshold, max_level)
    assert [p["flagged"] for p in log_nat] == [p["flagged"] for p in log_ref]
    assert log_nat[-1]["flagged"] == 0 and len(log_nat) > 3 and nat.num_leaves() == ref.num_leaves() > 16
    assert_same_grid(oracleapi.graph_from_ref(ref.sampled_function().graph()), nat.graph(), "after the loop")


@pytest.mark.parametrize("N,basis,boot,steps", [(2, capi.CUBIC, 2, 25), (2, capi.LINEAR, 2, 25), (3, capi.LINEAR, 1, 15)])
def test_optimal_split_direction_equals_the_reference(N, basis, boot, steps):
    """ERROR_BASED split policy on randomly refined grids: every leaf, against LCAdaptiveGridCell::getOptimalSplitDirection"""
    ref = refapi.RefGrid(N, basis, False, 0, 0.5, boot)
    g0 = oracleapi.graph_from_ref(ref.sampled_function().graph())
    nat = HostGrid.from_graph(g0, ref.true_values)
    rng = np.random.default_rng(21)
    for step in range(steps):
        L = ref.num_leaves()
        want = [ref.optimal_direction(l) for l in range(L)]
        assert min(want) >= 0
        assert [nat.optimal_direction(l) for l in range(L)] == want, step
        leaf = int(rng.integers(0, L))
        ref.refine(leaf, want[leaf])
        nat.refine(leaf, want[leaf])
    assert_same_grid(oracleapi.graph_from_ref(ref.sampled_function().graph()), nat.graph())


def test_optimal_split_direction():
    """ERROR_BASED split policy (getOptimalSplitDirection, LCAdaptiveGridCell.cpp:977-1041) on a function that varies in one
    direction only: the direction of the variation is chosen in every leaf"""
    for vary in (0, 1, 2):
        nat = HostGrid.bootstrap(3, capi.LINEAR, 1, [0, 1, 0, 2, -1, 1], lambda p, v=vary: np.sin(3 * p[:, v]))
        assert [nat.optimal_direction(l) for l in range(nat.num_leaves())] == [vary] * 8
