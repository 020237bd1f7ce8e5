"""GPU-assisted refinement pass: the error check of the adaptive sampling loop runs batched on the
device, the topology surgery stays with the host that owns the grid.

The reference's LCAdaptiveGrid::adaptiveSample (LCAdaptiveGrid.cpp:481-545) pops leaves largest-first from a heap and
runs decideSplit (:578-626) -> decideSplitDirection (:547-562) -> refineCell (:629-914) one cell at a time.  Here one
PASS decides a whole size class at once: `loco_center_error` evaluates the interpolant at every leaf centre against the
stored centre value in one launch (the CENTER_BASED metric, :594-602, NaN => split), this module picks the largest
flagged leaves that are still below the level limit (`level < maxLevel_`, :583) with the reference's CYCLIC direction
rule (:553), and the caller's `surgery(leaf, direction)` callback performs refineCell on its own grid object — in the
reference that is LCAdaptiveGrid::refineCell itself.  After a pass the caller hands back the re-flattened model.

The schedule differs from the reference's on purpose: within one size class the reference re-evaluates every cell after
the refinements that precede it in heap order, a pass uses the state at the start of the pass for all of them.  The
per-leaf decisions of a pass are exactly the reference's decideSplit on that state (tests/test_gpu_refine.py).
"""
from typing import Callable, List, Tuple

import numpy as np

from .capi import Model


def leaf_geometry(model: Model) -> Tuple[np.ndarray, np.ndarray]:
    """(volume [L], level [L]) of every leaf from its box in cube coordinates; the root [-1,1]^N has level 0 and every
    split adds one (LCAdaptiveGridCell::size / getLevel, LCAdaptiveGridCell.cpp:202-210, :121-131)."""
    boxes = np.stack([model.leaf_box(l) for l in range(model.L)])  # [L, N, 2]
    width = boxes[:, :, 1] - boxes[:, :, 0]
    level = np.rint(np.log2(2.0 / width).sum(axis=1)).astype(np.int64)
    return width.prod(axis=1), level


def cyclic_direction(model: Model, leaf: int, literal: bool = False) -> int:
    """CYCLIC split policy.  Default: the direction after the parent's in cyclic order, LCAdaptiveGrid.cpp:553 generalised
    from 2 to N parameters -- on the trees the reference builds (bootstrapUniformSplit cycles 0..N-1) that is the first
    dimension of largest extent.  literal=True: the reference's expression as written, `(parentDir + 1) % 2`, which for
    N > 2 only ever alternates between directions 0 and 1 (SURVEY.md 8a hazard H7); identical to the default for N = 2."""
    if literal:
        return (model.leaf_parent_split(leaf) + 1) % 2
    box = model.leaf_box(leaf)
    return int(np.argmax(box[:, 1] - box[:, 0]))


def select_splits(model: Model, threshold: float, max_level: int, literal_cyclic: bool = False):
    """One batched decideSplit over all leaves.  Returns (leaves, directions, err): the flagged leaves of the LARGEST
    flagged size class (what the reference's heap would pop next), in descending leaf order so that splitting one of them
    leaves the indices of those still to come untouched (leaves are numbered depth-first, child1 first)."""
    err, split = model.center_error(threshold)
    volume, level = leaf_geometry(model)
    flagged = np.flatnonzero((split != 0) & (level < max_level))
    if len(flagged) == 0:
        return [], [], err
    top = volume[flagged].max()
    chosen = [int(l) for l in flagged[volume[flagged] >= top * (1 - 1e-12)]][::-1]
    return chosen, [cyclic_direction(model, l, literal_cyclic) for l in chosen], err


def adaptive_sample_passes(get_model: Callable[[], Model], surgery: Callable[[int, int], None], threshold: float, max_level: int,
                           max_passes: int = 256, literal_cyclic: bool = False, get_graph: Callable[[], object] = None,
                           device: int = 0) -> List[dict]:
    """Level-synchronous adaptiveSample: repeat {flatten -> batched error check on the device -> host surgery on the
    flagged leaves of the largest size class} until no leaf is flagged.  `get_model()` returns the model of the grid's
    current state (from_graph of whatever the host's grid serialises to); `surgery(leaf, direction)` is refineCell.

    Incremental form: pass `get_graph()` (-> capi.GraphArrays of the grid's current state) instead of `get_model`
    (get_model=None).  The first pass creates the handle; later passes call Model.update_from_graph on it
    (loco_model_update_from_graph): leaves outside the split cells' doubly-adjacent sets keep their evaluation cells."""
    log = []
    model = None
    for _ in range(max_passes):
        if get_graph is not None:
            if model is None:
                model = Model.from_graph(get_graph(), device=device)
            else:
                model.update_from_graph(get_graph())
        else:
            model = get_model()
        leaves, dirs, err = select_splits(model, threshold, max_level, literal_cyclic)
        log.append({"leaves": model.L, "flagged": len(leaves), "max_center_error": float(np.nanmax(err)) if len(err) else 0.0,
                    "reused_leaves": int(model.info.get("n_reused_leaves", 0))})
        if not leaves:
            return log
        for leaf, direction in zip(leaves, dirs):
            surgery(leaf, direction)
    raise RuntimeError("adaptive_sample_passes: max_passes reached")
