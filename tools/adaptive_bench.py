"""Adaptive-tree check (not a BASELINE config): the reference's own builder produces the TestDebug tree
(TestDebug.cpp:153-157: 2-D cubic, maxDepth 10, threshold 0.1, bootstrap 2 -> ~724 leaves, ~54k basis functions),
the product evaluates it through the GENERIC term-list kernels, parity is checked against the reference on a sample
and throughput is reported next to the reference's single-thread rate.  Needs oracle/_ref/liblocoref.so and a GPU.

    python tools/adaptive_bench.py [--depth 10] [--queries 10000000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=10)
    ap.add_argument("--basis", type=int, default=1)
    ap.add_argument("--queries", type=int, default=10_000_000)
    args = ap.parse_args()
    import torch
    from locointerpolate_b200 import capi
    from oracle import oracleapi, refapi
    t0 = time.perf_counter()
    grid = refapi.RefGrid(2, args.basis, False, args.depth, 0.1, 2)
    fn = grid.sampled_function()
    build_s = time.perf_counter() - t0
    g = oracleapi.graph_from_ref(fn.graph())
    m = capi.Model.from_graph(g, device=0)
    info = m.info
    rng = np.random.default_rng(1)
    qs = rng.random((20_000, 2))
    t0 = time.perf_counter()
    r_out, r_leaf, r_status = fn.eval(qs)
    ref_qps = len(qs) / (time.perf_counter() - t0)
    res = {}
    Q = args.queries
    q = torch.rand((Q, 2), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    out = torch.empty(Q, dtype=torch.float64, device="cuda")
    leaf = torch.empty(Q, dtype=torch.int32, device="cuda")
    status = torch.empty(Q, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for path, name in ((1, "direct"), (2, "tma_tiles"), (4, "sorted_chunks")):
        m.set_option("path", path)
        o, l, s = m.eval(qs)
        ok = s == 0
        worst = float(np.max(np.abs(o[ok] - r_out[ok]) / np.maximum(np.abs(r_out[ok]), 1.0)))
        assert np.array_equal(l, r_leaf) and np.array_equal(s, r_status) and worst < 1e-12, (name, worst)
        for _ in range(2):
            m.eval_device(q.data_ptr(), Q, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), st)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            m.eval_device(q.data_ptr(), Q, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), st)
        b.record()
        torch.cuda.synchronize()
        res[name] = {"queries_per_s": Q / (a.elapsed_time(b) / 3 * 1e-3), "worst_rel_diff_vs_reference": worst}
    print(json.dumps({"tree": {"leaves": info["n_leaves"], "samples": info["n_samples"], "basis_functions": info["n_basis_functions"],
                               "terms_per_leaf": info["n_terms"] / info["n_leaves"], "max_terms": info["max_terms"], "depth": info["depth"],
                               "reference_build_s": build_s},
                      "reference_single_thread_qps": ref_qps, "gpu": res}))


if __name__ == "__main__":
    main()
