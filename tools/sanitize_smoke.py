"""small end-to-end pass over every kernel family, meant to be run under `compute-sanitizer --tool memcheck`"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from locointerpolate_b200 import capi  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
rng = np.random.default_rng(0)
for name in ("cub2d_adapt6", "lin3d_adapt5", "cub3d_boot2"):
    m = capi.Model.from_proto(os.path.join(G, name + ".pb"), device=0)
    q = rng.random((3001, m.N)) * 1.02 - 0.01
    ref = None
    for path in (1, 2, 3, 4, 5):
        m.set_option("path", path)
        m.set_option("chunk", 1000)
        out, leaf, status = m.eval(q)
        if ref is None:
            ref = out
        assert np.allclose(out, ref, rtol=1e-12, atol=1e-12, equal_nan=True)
    for d in range(m.N):
        m.eval_deriv(q, d)
    m.center_error(0.05)
    m.error_check(q, rng.random(len(q)), 0.1)
    m.free()
for N, basis, level, corners, D in ((3, capi.CUBIC, 1, False, 10), (4, capi.LINEAR, 1, True, 131), (2, capi.CUBIC, 2, False, 260), (1, capi.LINEAR, 2, True, 6)):
    m = capi.Model.bootstrap(N, basis, level, D, corners, device=0)
    m.fill_values_synthetic(0.3)
    q = rng.random((777, N)) * 1.02 - 0.01
    outs = []
    for simt, path in ((0, 0), (1, 0), (0, 1)):
        m.set_option("mesh_simt", simt)
        m.set_option("path", path)
        outs.append(m.eval(q)[0])
    assert np.allclose(outs[0], outs[1], rtol=1e-12, atol=1e-12, equal_nan=True) and np.allclose(outs[0], outs[2], rtol=1e-12, atol=1e-12, equal_nan=True)
    import torch
    tq = torch.from_numpy(q).cuda()
    ring = torch.empty((128, D + (D & 1)), dtype=torch.float64, device="cuda")
    chk = torch.empty(len(q), dtype=torch.float64, device="cuda")
    m.set_option("mesh_simt", 0)
    m.set_option("path", 0)
    m.eval_ring_device(tq.data_ptr(), len(q), ring.data_ptr(), 128, chk.data_ptr(), 0, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    wts = 1 + np.arange(D) % 7
    ok = ~np.isnan(outs[0][:, 0])
    assert np.allclose(chk.cpu().numpy()[ok], (outs[0][ok] * wts).sum(axis=1), rtol=1e-11)
    m.free()
print("sanitize_smoke ok")
