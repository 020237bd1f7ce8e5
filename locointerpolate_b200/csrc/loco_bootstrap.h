// loco_bootstrap.h -- see loco_bootstrap.cpp
#pragma once
#include "loco_graph.h"

namespace loco {

// f : original parameter space [N] -> value [D]   (LCFunction::evalShapeInfo, LCFunction.h:32)
typedef void (*ValueFn)(const double *params, double *out, void *user);

// Graph of LCAdaptiveGrid(f, basis, LCAdaptiveSamplingParams(maxDepth<=N*level, ., level, .), evalCorners)
// before any adaptive refinement.  fn == nullptr leaves the value table virtual (filled on the device).
Status bootstrap_graph(int N, int basis, int level, int D, bool eval_corners, const double *domain,
                       ValueFn fn, void *user, Graph *out);

} // namespace loco
