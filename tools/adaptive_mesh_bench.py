"""mesh-valued function on an ADAPTIVE tree (generic, non-tensor leaves): parity vs the oracle and throughput.
Tree: the reference's own builder, 2-D cubic, maxDepth 7, threshold 0.1, bootstrap 2; values: random [S x D]."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from locointerpolate_b200 import capi
from oracle import oracleapi, refapi

D = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 7
fn = refapi.RefGrid(2, 1, False, depth, 0.1, 2).sampled_function()
g = oracleapi.graph_from_ref(fn.graph())
rng = np.random.default_rng(2)
g.D = D
g.sample_value = rng.standard_normal((fn.S, D))
g.sample_value_group = None
g.cell_sample_value = g.border_sample_value = g.cell_center_value = g.border_center_value = None
g.normalise()
m = capi.Model.from_graph(g, device=0)
info = m.info
orc = oracleapi.Oracle(g)
qs = rng.random((300, 2))
o_out, o_leaf, o_status, scale = orc.eval(qs)
out, leaf, status = m.eval(qs)
ok = status == 0
worst = float(np.max(np.abs(out[ok] - o_out[ok]) / np.maximum(np.abs(o_out[ok]), scale[ok][:, None])))
assert np.array_equal(leaf, o_leaf) and worst < 1e-12, worst
q = torch.rand((Q, 2), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
outd = torch.empty((Q, D + (D & 1)), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    m.eval_device(q.data_ptr(), Q, outd.data_ptr(), 0, 0, st)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    m.eval_device(q.data_ptr(), Q, outd.data_ptr(), 0, 0, st)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(json.dumps({"leaves": info["n_leaves"], "samples": info["n_samples"], "max_support": info["max_support"], "avg_support": info["n_support_entries"] / info["n_leaves"],
                  "D": D, "Q": Q, "ms": ms, "queries_per_s": Q / (ms * 1e-3), "results_GBps": Q * D * 8 / (ms * 1e-3) / 1e9, "worst_rel": worst}))
