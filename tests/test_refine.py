"""Host logic of the GPU-assisted refinement pass (`locointerpolate_b200.refine`) without a GPU: the
device error check is replaced by the reference's own per-leaf centre errors, the surgery is the reference's refineCell
(oracle/_ref, test infrastructure).  The GPU version of this test is tests/test_gpu_refine.py."""
import numpy as np
import pytest

from locointerpolate_b200 import capi, refine
from oracle import oracleapi, refapi


@pytest.mark.skipif(not refapi.available(), reason="oracle/_ref (the compiled reference) is not built")
@pytest.mark.parametrize("basis,max_level,threshold", [(capi.LINEAR, 8, 0.05), (capi.CUBIC, 7, 0.1)])
def test_pass_selection_and_leaf_bookkeeping(basis, max_level, threshold):
    grid = refapi.RefGrid(2, basis, False, 0, threshold, 2)
    seen = []

    def get_model():
        fn = grid.sampled_function()
        m = capi.Model.from_graph(oracleapi.graph_from_ref(fn.graph()), device=capi.HOST_ONLY)
        ref_err = np.array([fn.center_error(l) for l in range(fn.L)])
        m.center_error = lambda t: (ref_err, (~((ref_err - t) <= 0)).astype(np.uint8))  # stands in for loco_center_error
        seen.append(m)
        return m

    def surgery(leaf, direction):
        m = seen[-1]
        volume, level = refine.leaf_geometry(m)
        box = m.leaf_box(leaf)
        assert level[leaf] < max_level and volume[leaf] == volume[[l for l, _ in picked]].max()
        assert direction == int(np.argmax(box[:, 1] - box[:, 0]))
        grid.refine(leaf, direction)

    log = []
    for _ in range(64):
        m = get_model()
        leaves, dirs, err = refine.select_splits(m, threshold, max_level)
        log.append(len(leaves))
        if not leaves:
            break
        assert leaves == sorted(leaves, reverse=True)  # descending: a split never renumbers a leaf still to come
        picked = list(zip(leaves, dirs))
        for leaf, direction in picked:
            surgery(leaf, direction)
        assert grid.num_leaves() == m.L + len(leaves)
    assert log[-1] == 0 and len(log) > 3
    # same loop through the packaged driver from a fresh grid: same number of leaves
    grid2 = refapi.RefGrid(2, basis, False, 0, threshold, 2)

    def get_model2():
        fn = grid2.sampled_function()
        m = capi.Model.from_graph(oracleapi.graph_from_ref(fn.graph()), device=capi.HOST_ONLY)
        ref_err = np.array([fn.center_error(l) for l in range(fn.L)])
        m.center_error = lambda t: (ref_err, (~((ref_err - t) <= 0)).astype(np.uint8))
        return m
    plog = refine.adaptive_sample_passes(get_model2, grid2.refine, threshold, max_level)
    assert [p["flagged"] for p in plog] == log and plog[-1]["leaves"] == grid.num_leaves()
    # levels: every leaf within the limit, the root cell's children start at level 4 (bootstrap 2 in 2-D)
    _, level = refine.leaf_geometry(seen[-1])
    assert level.min() >= 4 and level.max() <= max_level


@pytest.mark.skipif(not refapi.available(), reason="oracle/_ref (the compiled reference) is not built")
@pytest.mark.parametrize("basis", [capi.CUBIC, capi.LINEAR])
def test_incremental_reflatten_equals_a_fresh_flatten(basis):
    """loco_model_update_from_graph (re-flatten after refineCell, LCAdaptiveGrid.cpp:629-914): the handle of an
    earlier grid state updated to the refined state must hold exactly what a handle created from the refined state holds --
    leaf numbering, influencing-sample sets, merged basis functions, evaluation cells with their term lists, polynomial
    cells -- while leaves outside the doubly-adjacent sets of the split cells (getDoubleAdjacentCells, :174) are reused."""
    grid = refapi.RefGrid(2, basis, False, 6, 0.1, 2)     # an adaptively refined tree built by the reference
    m = capi.Model.from_graph(oracleapi.graph_from_ref(grid.sampled_function().graph()), device=capi.HOST_ONLY)
    assert m.info["n_eval_cells"] > m.L and m.info["n_reused_leaves"] == 0
    rng = np.random.default_rng(7)
    for step in range(3):
        for _ in range(2):
            grid.refine(int(rng.integers(0, grid.num_leaves())), int(rng.integers(0, 2)))
        ga = oracleapi.graph_from_ref(grid.sampled_function().graph())
        m.update_from_graph(ga)
        fresh = capi.Model.from_graph(ga, device=capi.HOST_ONLY)
        for k in ("n_leaves", "n_samples", "n_cells", "n_support_entries", "n_terms", "n_eval_cells", "n_eval_terms", "n_poly_cells", "depth", "max_support"):
            assert m.info[k] == fresh.info[k], (step, k)
        assert 0 < m.info["n_reused_leaves"] < m.L, (step, m.info["n_reused_leaves"], m.L)   # untouched leaves were reused, touched ones not
        for l in range(fresh.L):
            assert np.array_equal(m.support_set(l), fresh.support_set(l))
            assert np.array_equal(m.leaf_box(l), fresh.leaf_box(l))
            c1, s1, f1 = m.merged_terms(l)
            c2, s2, f2 = fresh.merged_terms(l)
            assert f1 == f2 and np.array_equal(c1, c2) and np.array_equal(s1, s2)
            b1, first1, n1 = m.eval_cells(l)
            b2, first2, n2 = fresh.eval_cells(l)
            assert np.array_equal(b1, b2) and first1 == first2 and n1 == n2
            for c in range(first1, first1 + n1):
                assert np.array_equal(m.eval_cell_terms(c), fresh.eval_cell_terms(c)), (step, l, c)
                p1, p2 = m.eval_cell_poly(c), fresh.eval_cell_poly(c)
                assert (p1 is None) == (p2 is None)
                if p1 is not None:
                    assert np.array_equal(p1[0], p2[0]) and np.array_equal(p1[1], p2[1])
    # an update with an unchanged graph reuses every leaf
    m.update_from_graph(ga)
    assert m.info["n_reused_leaves"] == m.L
