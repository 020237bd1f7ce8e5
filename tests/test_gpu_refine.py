"""GPU-assisted refinement pass: the batched decideSplit of `locointerpolate_b200.refine` drives the
REFERENCE's own refineCell (compiled under oracle/_ref — test infrastructure playing the host that owns the grid).
Every pass checks the device's per-leaf centre errors and split flags against the reference's decideSplit data on the
same grid state; at the end the refined tree evaluates like the reference's own LCSampledFunction over that grid."""
import numpy as np
import pytest

from locointerpolate_b200 import capi, refine
from oracle import oracleapi, refapi

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not refapi.available(), reason="oracle/_ref (the compiled reference) is not built")
@pytest.mark.parametrize("basis,max_level,threshold", [(capi.LINEAR, 8, 0.05), (capi.CUBIC, 7, 0.1)])
def test_level_synchronous_refinement_with_reference_surgery(basis, max_level, threshold):
    # bootstrap only: maxLevel 0 keeps the reference's own adaptiveSample from splitting anything (LCAdaptiveGrid.cpp:583)
    grid = refapi.RefGrid(2, basis, False, 0, threshold, 2)
    state = {}

    def get_model():
        fn = grid.sampled_function()
        m = capi.Model.from_graph(oracleapi.graph_from_ref(fn.graph()), device=0)
        err, split = m.center_error(threshold)
        ref_err = np.array([fn.center_error(l) for l in range(fn.L)])
        assert m.L == fn.L == grid.num_leaves()
        assert np.allclose(err, ref_err, rtol=0, atol=1e-13)
        assert np.array_equal(split.astype(bool), ~((ref_err - threshold) <= 0))  # LCAdaptiveGrid.cpp:601
        state["fn"], state["m"] = fn, m
        return m

    log = refine.adaptive_sample_passes(get_model, grid.refine, threshold, max_level)
    assert len(log) > 2 and log[-1]["flagged"] == 0 and log[-1]["leaves"] > log[0]["leaves"] == 16
    assert sum(p["flagged"] for p in log) == log[-1]["leaves"] - 16  # every split adds one leaf
    # no leaf below the level limit is left with an error above the threshold
    m, fn = state["m"], state["fn"]
    err, split = m.center_error(threshold)
    _, level = refine.leaf_geometry(m)
    assert not (split.astype(bool) & (level < max_level)).any()
    # the refined tree evaluates like the reference's LCSampledFunction over the same grid
    q = np.random.default_rng(3).random((4000, 2))
    out, leaf, status = m.eval(q)
    r_out, r_leaf, r_status = fn.eval(q)
    assert np.array_equal(leaf, r_leaf) and np.array_equal(status, r_status)
    ok = status == 0
    assert np.abs(out[ok] - r_out[ok]).max() <= 1e-12 * max(1.0, np.abs(r_out[ok]).max())
    # and it approximates the sampled function about as well as the reference's sequential schedule does
    seq = refapi.RefGrid(2, basis, False, max_level, threshold, 2)
    truth = grid.true_values(q)
    s_out, _, s_status = seq.sampled_function().eval(q)
    e_pass, e_seq = np.abs(out[ok] - truth[ok]).max(), np.abs(s_out[s_status == 0] - truth[s_status == 0]).max()
    assert e_pass <= 2.0 * max(e_seq, threshold), (e_pass, e_seq, log[-1]["leaves"], seq.num_leaves())


@pytest.mark.skipif(not refapi.available(), reason="oracle/_ref (the compiled reference) is not built")
def test_refinement_loop_on_one_handle_updated_incrementally():
    """The same loop with ONE device handle that is updated in place after every pass (loco_model_update_from_graph): the
    pass decisions and the final evaluation must equal those of the loop that builds a fresh handle per pass."""
    threshold, max_level = 0.1, 7

    def run(incremental):
        grid = refapi.RefGrid(2, capi.CUBIC, False, 0, threshold, 2)
        graph = lambda: oracleapi.graph_from_ref(grid.sampled_function().graph())
        if incremental:
            log = refine.adaptive_sample_passes(None, grid.refine, threshold, max_level, get_graph=graph, device=0)
        else:
            log = refine.adaptive_sample_passes(lambda: capi.Model.from_graph(graph(), device=0), grid.refine, threshold, max_level)
        return grid, log
    g_inc, log_inc = run(True)
    g_new, log_new = run(False)
    assert [(p["leaves"], p["flagged"]) for p in log_inc] == [(p["leaves"], p["flagged"]) for p in log_new]
    assert np.allclose([p["max_center_error"] for p in log_inc], [p["max_center_error"] for p in log_new], rtol=0, atol=1e-13)
    assert max(p["reused_leaves"] for p in log_inc) > 0 and all(p["reused_leaves"] == 0 for p in log_new)
    # one more update of a handle, then evaluation against a fresh handle of the same state: identical
    ga = oracleapi.graph_from_ref(g_inc.sampled_function().graph())
    m = capi.Model.from_graph(oracleapi.graph_from_ref(refapi.RefGrid(2, capi.CUBIC, False, 0, threshold, 2).sampled_function().graph()), device=0)
    m.update_from_graph(ga)
    fresh = capi.Model.from_graph(ga, device=0)
    q = np.random.default_rng(5).random((300_000, 2))
    for path in (0, 1, 4):
        m.set_option("path", path)
        fresh.set_option("path", path)
        a, la, sa = m.eval(q)
        b, lb, sb = fresh.eval(q)
        assert np.array_equal(la, lb) and np.array_equal(sa, sb) and np.array_equal(a, b, equal_nan=True), path


@pytest.mark.skipif(not refapi.available(), reason="oracle/_ref (the compiled reference) is not built")
@pytest.mark.parametrize("basis,max_level,threshold", [(capi.LINEAR, 8, 0.05), (capi.CUBIC, 7, 0.1)])
def test_refinement_loop_with_native_surgery(basis, max_level, threshold):
    """The loop without the reference library in it: error check on the device, surgery by the native host grid
    (include/loco_grid.hpp, through locointerpolate_b200/grid.py), one handle updated incrementally.  The same loop with the
    reference's refineCell as the host makes the same decisions and ends in the same grid."""
    from locointerpolate_b200.grid import HostGrid, adaptive_sample
    from test_native_refine import assert_same_grid
    ref = refapi.RefGrid(2, basis, False, 0, threshold, 2)
    g0 = oracleapi.graph_from_ref(ref.sampled_function().graph())
    nat = HostGrid.bootstrap(2, basis, 2, g0.domain, ref.true_values)
    log_nat = adaptive_sample(nat, threshold, max_level, device=0)
    log_ref = refine.adaptive_sample_passes(None, ref.refine, threshold, max_level, device=0,
                                            get_graph=lambda: oracleapi.graph_from_ref(ref.sampled_function().graph()))
    assert [(p["leaves"], p["flagged"]) for p in log_nat] == [(p["leaves"], p["flagged"]) for p in log_ref]
    assert np.allclose([p["max_center_error"] for p in log_nat], [p["max_center_error"] for p in log_ref], rtol=0, atol=1e-13)
    assert log_nat[-1]["flagged"] == 0 and log_nat[-1]["leaves"] > 16
    if basis == capi.CUBIC:   # evaluation cells (the derived data the incremental update carries over) exist on cubic scalar trees
        assert max(p["reused_leaves"] for p in log_nat) > 0
    assert_same_grid(oracleapi.graph_from_ref(ref.sampled_function().graph()), nat.graph(), "after the loop")
    # the native grid's model evaluates like the reference's LCSampledFunction over the reference-refined grid
    m = capi.Model.from_graph(nat.graph(), device=0)
    fn = ref.sampled_function()
    q = np.random.default_rng(3).random((4000, 2))
    out, leaf, status = m.eval(q)
    r_out, r_leaf, r_status = fn.eval(q)
    assert np.array_equal(leaf, r_leaf) and np.array_equal(status, r_status)
    ok = status == 0
    assert np.abs(out[ok] - r_out[ok]).max() <= 1e-12 * max(1.0, np.abs(r_out[ok]).max())
