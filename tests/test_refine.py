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
