/* loco_grid.h -- C ABI of the HOST side of the refinement loop: the builder's grid state with
 * bootstrapSample and refineCell's topology surgery (include/loco_grid.hpp), for hosts that are not C++
 * (locointerpolate_b200/grid.py binds it with ctypes).  Implemented by locointerpolate_b200/libloco_grid.so, a
 * host-only library (no CUDA, no kernels; built by g++ from hostsrc/loco_grid_capi.cpp).  It never evaluates the
 * interpolant: the error check of the loop is loco_center_error / loco_error_check* of loco_b200.h on the device,
 * and the grid's state reaches the device as a loco_graph_t (loco_model_from_graph / loco_model_update_from_graph).
 *
 * Return values: LOCO_OK or LOCO_ERR_*; loco_grid_last_error gives the text (the reference's LCError descriptions
 * where it has one: "cannot split border cells", "split direction out of range", "error, the refinement is not half",
 * "Not found on border", "adjacent cells is empty", "weights do not sum to one: ...", "could not find corner shape"). */
#ifndef LOCO_GRID_H
#define LOCO_GRID_H

#include "loco_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct loco_grid_handle loco_grid_t;

/* LCAdaptiveGrid(f, basis, params, evalCorners) up to and including bootstrapSample(level)
 * (LCAdaptiveSampling/LCAdaptiveGrid.cpp:16-31, 256-456).  fn: the function at ORIGINAL parameters
 * (LCFunction::evalShapeInfo, LCFunction/LCFunction.h:32); domain [2N] = min,max per parameter. */
int loco_grid_bootstrap(int32_t n_params, int32_t basis_type, int32_t level, int32_t value_dim, int32_t eval_corners,
                        const double *domain, loco_value_fn fn, void *user, loco_grid_t **out);

/* The builder state of a grid given as plain arrays (e.g. dumped from a reference-side LCAdaptiveGrid: getRoot,
 * getSamples, getBorderCells, LCAdaptiveGrid.h:30-36), with the LIVE builder's neighbour relations (border cells linked
 * to the leaves and border cells they touch, LCAdaptiveGrid.cpp:374-451).  fn is called for new samples / leaf centres. */
int loco_grid_from_graph(const loco_graph_t *graph, loco_value_fn fn, void *user, loco_grid_t **out);

void loco_grid_free(loco_grid_t *g);
const char *loco_grid_last_error(const loco_grid_t *g); /* g == NULL: the error of the last failed constructor */

/* LCAdaptiveGrid::refineCell(cell, splitDirection, &affected) (LCAdaptiveGrid.cpp:629-914) for the leaf-th leaf of
 * LCAdaptiveGridCell::getAllLeafCells order (:247-258; the leaf numbering of the flat model). */
int loco_grid_refine_leaf(loco_grid_t *g, int32_t leaf, int32_t direction);

/* The requests of one level-synchronous pass: leaf indices of ONE numbering in DESCENDING order (so that a split leaves
 * the indices still to come valid) -- what LCAdaptiveGrid::selectSplits (loco_lc.hpp) / refine.select_splits return. */
int loco_grid_refine_leaves(loco_grid_t *g, const int32_t *leaf, const int32_t *direction, int64_t n);

/* LCAdaptiveGridCell::getOptimalSplitDirection (LCAdaptiveGridCell.cpp:977-1041): the ERROR_BASED split policy */
int loco_grid_optimal_split_direction(const loco_grid_t *g, int32_t leaf, int32_t *direction);

/* Plain-array image of the current state (pointers owned by the grid, valid until its next call): tree cells in
 * pre-order, samples in samples_ order, for loco_model_from_graph / loco_model_update_from_graph. */
int loco_grid_graph(loco_grid_t *g, loco_graph_t *out);

/* counts: leaves, samples, calls of fn so far; neighbours of a leaf (tree leaves / border cells, LCAdaptiveGridCell::getNeighbors) */
int64_t loco_grid_num_leaves(const loco_grid_t *g);
int64_t loco_grid_num_samples(const loco_grid_t *g);
int64_t loco_grid_num_function_evaluations(const loco_grid_t *g);
int loco_grid_leaf_neighbors(const loco_grid_t *g, int32_t leaf, int32_t *n_leaf, int32_t *n_border);

#ifdef __cplusplus
}
#endif
#endif /* LOCO_GRID_H */
