"""The adaptive BUILD as a caller of the hot path: the reference's sequential builder
(LCAdaptiveGrid's constructor: bootstrapSample + adaptiveSample, per-cell decideSplit + refineCell on the CPU) against the
level-synchronous loop of this repo (error check of all leaves on the GPU, refineCell on the native host grid of
include/loco_grid.hpp, incremental re-flatten), on the reference's own test function and the TestDebug parameters
(TestDebug.cpp:153-157: 2-D cubic, maxDepth 10, threshold 0.1, bootstrap 2).

    python tools/refine_bench.py [--max-depth 10] [--threshold 0.1] [--bootstrap 2] [--basis 1]

Needs a GPU and oracle/_ref (the compiled reference: its function is the one sampled, its builder the one timed).
Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locointerpolate_b200 import capi  # noqa: E402
from locointerpolate_b200.grid import HostGrid, adaptive_sample  # noqa: E402
from oracle import refapi  # noqa: E402  (measurement harness: the reference arm and the sampled function)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-depth", type=int, default=10)
    ap.add_argument("--threshold", type=float, default=0.1)
    ap.add_argument("--bootstrap", type=int, default=2)
    ap.add_argument("--basis", type=int, default=capi.CUBIC)
    ap.add_argument("--n-params", type=int, default=2)
    ap.add_argument("--skip-reference", action="store_true")
    a = ap.parse_args()
    N = a.n_params
    boot = refapi.RefGrid(N, a.basis, False, 0, a.threshold, a.bootstrap)   # bootstrap only: the function and the domain
    dom = boot.sampled_function().graph().ranges
    n_fn = [0]

    def f(p):
        n_fn[0] += len(p)
        return boot.true_values(p)

    # warm the device path once (context, module load) outside the timed region
    warm = HostGrid.bootstrap(N, a.basis, a.bootstrap, dom, f)
    capi.Model.from_graph(warm.graph(), device=0).center_error(a.threshold)
    n_fn[0] = 0
    t0 = time.perf_counter()
    grid = HostGrid.bootstrap(N, a.basis, a.bootstrap, dom, f)
    t_boot = time.perf_counter() - t0
    log = adaptive_sample(grid, a.threshold, a.max_depth, device=0, error_based_direction=True)
    t_ours = time.perf_counter() - t0
    model = capi.Model.from_graph(grid.graph(), device=0)
    rng = np.random.default_rng(1)
    lo, hi = np.asarray(dom).reshape(N, 2).T
    q = lo + rng.random((200_000, N)) * (hi - lo)
    truth = boot.true_values(q)
    out, _, status = model.eval(q)
    ok = status == 0
    res = {"what": "adaptive build: level-synchronous loop (GPU error check + native refineCell + incremental re-flatten)",
           "config": {"n_params": N, "basis": "cubic" if a.basis == capi.CUBIC else "linear", "max_depth": a.max_depth, "threshold": a.threshold,
                      "bootstrap": a.bootstrap, "split_direction": "ERROR_BASED", "error_metric": "CENTER_BASED"},
           "ours": {"seconds": t_ours, "bootstrap_seconds": t_boot, "passes": len(log), "leaves": grid.num_leaves(), "samples": grid.num_samples(),
                    "function_evaluations": n_fn[0], "max_abs_error_200k_queries": float(np.abs(out[ok] - truth[ok]).max()),
                    "mean_abs_error": float(np.abs(out[ok] - truth[ok]).mean()), "pass_log": log}}
    if not a.skip_reference:
        t0 = time.perf_counter()
        ref = refapi.RefGrid(N, a.basis, False, a.max_depth, a.threshold, a.bootstrap)
        t_ref = time.perf_counter() - t0
        fn = ref.sampled_function()
        r_out, _, r_status = fn.eval(q[:20_000])
        rk = r_status == 0
        res["reference"] = {"seconds": t_ref, "leaves": ref.num_leaves(), "what": "LCAdaptiveGrid constructor of the compiled reference (sequential, one core)",
                            "max_abs_error_20k_queries": float(np.abs(r_out[rk] - truth[:20_000][rk]).max()),
                            "mean_abs_error": float(np.abs(r_out[rk] - truth[:20_000][rk]).mean())}
        res["speedup"] = t_ref / t_ours
    print(json.dumps(res))


if __name__ == "__main__":
    main()
