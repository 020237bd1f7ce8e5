"""Throughput of the batched evalDeriv on the c2 model (3-D cubic, bootstrap 5, 32,768 leaves):
thread-per-query term-list kernel (path 1) against the leaf-bucket path with the derivative factor (automatic dispatch).

    python tools/deriv_bench.py [--queries 100000000]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=100_000_000)
    ap.add_argument("--level", type=int, default=5)
    args = ap.parse_args()
    import torch
    from locointerpolate_b200 import capi
    N, Q = 3, args.queries

    def f(p):
        return 1 + np.sin(3 * p + np.arange(len(p))).sum()
    m = capi.Model.bootstrap(N, capi.CUBIC, args.level, 1, True, fn=f, device=0)
    q = torch.rand((Q, N), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    out = torch.empty(Q, dtype=torch.float64, device="cuda")
    leaf = torch.empty(Q, dtype=torch.int32, device="cuda")
    status = torch.empty(Q, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    ref = None
    for path, name in ((1, "direct"), (0, "buckets")):
        m.set_option("path", path)
        for _ in range(2):
            m.eval_deriv_device(q.data_ptr(), Q, 1, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), st)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            m.eval_deriv_device(q.data_ptr(), Q, 1, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), st)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        res[name] = {"ms": ms, "queries_per_s": Q / (ms * 1e-3)}
        if ref is None:
            ref = out.clone()
        else:
            res[name]["max_abs_diff_vs_direct"] = float((out - ref).abs().max())
            res[name]["max_abs_value"] = float(ref.abs().max())
    print(json.dumps({"model": {"leaves": m.L, "N": N}, "queries": Q, "derivative": res}))


if __name__ == "__main__":
    main()
