"""Parity at the kernel instantiations the BASELINE configs time (VERDICT r01 "next round" item 2).

The mesh-valued generalisation has no reference implementation (LCFunctionValue.cpp:20-31 accepts LCRealFunctionValue only),
so its oracle is built from what the reference DOES have: the influencing samples and their weights for a query in the
reference's own support-list order (LCAdaptiveGridCell::getAllInfluencingSamples, LCAdaptiveGridCell.cpp:558-607 +
LCSample::getInterpolationWeight, LCSample.cpp:76-89, exposed by oracle/ref_adapter.cpp as ref_fn_weights) combined with the
value rows as LCRealFunctionValue::add does per component (`val += w * other`, LCRealFunctionValue.cpp:51-61) -- and,
independently, the CPU restatement in oracle/loco_oracle.cpp with D components.

Bars: leaf indices / status bytes bit-exact; values |a-b| <= 1e-12 * max(|ref|, sum_k |w_k| max_j |v_k[j]|) (fp64),
1e-5 relative for the optional fp32 mode."""
import os
import tempfile

import numpy as np
import pytest

import protoschema
from conftest import assert_values_close
from locointerpolate_b200 import capi
from oracle import oracleapi, refapi

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _scalar_twin_graph(N, basis, level, corners):
    """graph of the same bootstrap tree with scalar values (host-only handle -> proto -> arrays): tree, samples and basis
    functions do not depend on the value type, and the sample order is the generator's"""
    twin = capi.Model.bootstrap(N, basis, level, 1, corners, fn=lambda p: float(np.sin(3 * p).sum()), device=capi.HOST_ONLY)
    tmp = tempfile.mktemp(suffix=".pb")
    try:
        twin.save_proto(tmp)
        return protoschema.read_binary(tmp), twin
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)


def _mesh_graph(g, per_sample_values):
    """the twin's graph with D-component values per sample (per-cell copies = the sample's own value)"""
    g.D = per_sample_values.shape[1]
    g.sample_value = per_sample_values
    g.cell_sample_value = g.border_sample_value = g.cell_center_value = g.border_center_value = None
    g.sample_value_group = None
    return g.normalise()


def _reference_mesh_rows(fn, q_cube, leaf, per_sample_values, cols=None):
    """sum_k w_k * v_k[j] with the reference's own (sample, weight) lists; returns (rows [n, len(cols)], scale [n])"""
    out, scale = [], []
    for x, l in zip(q_cube, leaf):
        ids, w = fn.weights(int(l), x)
        v = per_sample_values(ids) if callable(per_sample_values) else per_sample_values[ids]
        if cols is not None and not callable(per_sample_values):
            v = v[:, cols]
        acc = np.zeros(v.shape[1])
        for k in range(len(ids)):   # the reference's order: r += w * v, one sample after the other
            acc += w[k] * v[k]
        out.append(acc)
        scale.append(float((np.abs(w) * np.abs(v).max(axis=1)).sum()))
    return np.array(out), np.array(scale)


@pytest.mark.parametrize("D", [257, 1000])
def test_c4_shape_6d_cubic_mesh_against_reference_weights(D):
    """BASELINE configs[3] kernel instantiation: mesh_tensor_mma_kernel<6, cubic, exact, ONE=false> (4 high dimensions in the
    mixed-radix weight walk, folded rows because evalCorners=false shares border rows), D not a multiple of the 128-component
    chunk; every leaf of the bootstrap-1 grid touches the border.  MMA, DFMA-tile, gather and ring paths."""
    torch = _torch()
    N, level = 6, 1
    g, twin = _scalar_twin_graph(N, capi.CUBIC, level, False)
    mesh = capi.Model.bootstrap(N, capi.CUBIC, level, D, False, fn=None, device=0)
    assert mesh.info["n_fold_entries"] < mesh.info["n_support_entries"]  # rows are folded
    rows = mesh.sample_value_rows()
    assert np.array_equal(rows, twin.sample_value_rows())
    rng = np.random.default_rng(61)
    table = rng.standard_normal((mesh.info["n_value_rows"], D))
    mesh.set_values(table)
    per_sample = table[rows]
    # queries: ~70 per leaf (one full 64-query tile + a partial one per leaf), plus points on faces / in the +-1e-6 band / outside
    Q = 64 * 70
    q = rng.random((Q, N)) * 1.0000004 - 0.0000002   # up to 4e-7 outside the cube: inside the reference's +-1e-6 acceptance band
    q[:64] = np.round(q[:64] * 4) / 4
    q[64:72, 0] = 1.5
    st = torch.cuda.current_stream().cuda_stream
    tq = torch.from_numpy(q).cuda()
    res = {}
    for key, (simt, path) in {"mma": (0, 0), "simt": (1, 0), "gather": (0, 1)}.items():
        mesh.set_option("mesh_simt", simt)
        mesh.set_option("path", path)
        out = torch.empty((Q, D), dtype=torch.float64, device="cuda")
        leaf = torch.empty(Q, dtype=torch.int32, device="cuda")
        status = torch.empty(Q, dtype=torch.uint8, device="cuda")
        mesh.profile_begin()
        mesh.eval_device(tq.data_ptr(), Q, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), st)
        prof = mesh.profile_end()
        res[key] = (out.cpu().numpy(), leaf.cpu().numpy(), status.cpu().numpy())
        want = {"mma": "mesh_tensor_mma", "simt": "mesh_tensor_tiles", "gather": "mesh_gather"}[key]
        assert want in prof, (key, prof)
    out, leaf, status = res["mma"]
    ok = status == 0
    assert (~ok).sum() == 8 and np.isnan(out[~ok]).all()
    fn = refapi.RefFn.from_graph(g) if refapi.available() else None   # (scalar graph: before g receives the mesh values)
    # the CPU restatement with D components, on a subset (4,096 samples x D per query)
    n_orc = 400 if D > 500 else 1200
    orc = oracleapi.Oracle(_mesh_graph(g, per_sample))
    o_out, o_leaf, o_status, o_scale = orc.eval(q[:n_orc])
    for key, (r_out, r_leaf, r_status) in res.items():
        assert np.array_equal(r_leaf, leaf) and np.array_equal(r_status, status), key
        assert np.array_equal(r_leaf[:n_orc], o_leaf) and np.array_equal(r_status[:n_orc], o_status), key
        oko = o_status == 0
        assert_values_close(r_out[:n_orc][oko], o_out[oko], o_scale[oko])
    # the reference's own weights, two queries per leaf
    assert fn is not None, "oracle/_ref/liblocoref.so must travel with the snapshot"
    if fn is not None:
        pick = np.concatenate([np.flatnonzero(ok & (leaf == l))[:2] for l in range(mesh.L)])
        assert len(pick) == 2 * mesh.L
        r_out_ref, r_leaf_ref, r_status_ref = fn.eval(q[pick])
        assert np.array_equal(r_leaf_ref, leaf[pick]) and (r_status_ref == 0).all()
        ref_rows, scale = _reference_mesh_rows(fn, q[pick] * 2 - 1, leaf[pick], per_sample)
        for key, (r_out, _, _) in res.items():
            assert_values_close(r_out[pick], ref_rows, scale)
    # ring path: per-query checksums of results streamed through a 128-query ring, against the kept results
    mesh.set_option("mesh_simt", 0)
    mesh.set_option("path", 0)
    ring = torch.empty((128, D + (D & 1)), dtype=torch.float64, device="cuda")
    chk = torch.empty(Q, dtype=torch.float64, device="cuda")
    mesh.eval_ring_device(tq.data_ptr(), Q, ring.data_ptr(), 128, chk.data_ptr(), 0, 0, st)
    torch.cuda.synchronize()
    wts = 1.0 + np.arange(D) % 7
    ref_chk = (out * wts).sum(axis=1)
    got = chk.cpu().numpy()
    assert np.isnan(got[~ok]).all()
    bound = 1e-12 * np.maximum(np.abs(ref_chk[ok]), (np.abs(out[ok]) * wts).sum(axis=1))
    assert (np.abs(got[ok] - ref_chk[ok]) <= bound).all()


def test_c3_shape_4d_linear_mesh_d30000():
    """BASELINE configs[2] at its own size: N=4 linear, bootstrap 3 (4,096 leaves, K=16), D = 30,000 on the synthetic table
    the bench uses -- mesh_tensor_mma_kernel<4, linear, exact, ONE=true>; 2,000 queries, 64 sampled components against the
    reference's own weights."""
    torch = _torch()
    N, level, D, Q, seed = 4, 3, 30_000, 2000, 0.5
    g, twin = _scalar_twin_graph(N, capi.LINEAR, level, True)
    mesh = capi.Model.bootstrap(N, capi.LINEAR, level, D, True, fn=None, device=0)
    mesh.fill_values_synthetic(seed)
    rows = mesh.sample_value_rows()
    rng = np.random.default_rng(62)
    cols = np.sort(rng.choice(D, 64, replace=False))
    cols[0], cols[-1] = 0, D - 1

    # the sampled columns of the device's table (the device forms sin(0.37 j + 0.61 row + seed) with fused multiply-adds, so a
    # host recomputation of the ARGUMENT differs by an ulp of ~1e4, i.e. ~2e-12 in the value: read the table back instead)
    R = mesh.info["n_value_rows"]
    tab = np.concatenate([mesh.get_values(r0, min(512, R - r0))[:, cols] for r0 in range(0, R, 512)])
    j = cols[None, :]
    assert np.abs(tab - (np.sin(0.37 * j + 0.61 * np.arange(R)[:, None] + seed) + 0.001 * (j % 97))).max() < 1e-10

    def synthetic(ids):
        return tab[rows[ids]]
    q = rng.random((Q, N)) * 1.0000004 - 0.0000002
    q[:256] = np.round(q[:256] * 8) / 8   # on sample positions / split planes
    q[256:260, 1] = -0.5
    tq = torch.from_numpy(q).cuda()
    out = torch.empty((Q, D), dtype=torch.float64, device="cuda")
    leaf = torch.empty(Q, dtype=torch.int32, device="cuda")
    status = torch.empty(Q, dtype=torch.uint8, device="cuda")
    mesh.profile_begin()
    mesh.eval_device(tq.data_ptr(), Q, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert "mesh_tensor_mma" in mesh.profile_end()
    got = out[:, torch.from_numpy(cols).cuda()].cpu().numpy()
    leaf, status = leaf.cpu().numpy(), status.cpu().numpy()
    ok = status == 0
    assert (~ok).sum() == 4 and bool(torch.isnan(out[torch.from_numpy(~ok).cuda()]).all())
    orc = oracleapi.Oracle(g)
    _, o_leaf, o_status, _ = orc.eval(q)
    assert np.array_equal(leaf, o_leaf) and np.array_equal(status, o_status)
    assert refapi.available(), "oracle/_ref/liblocoref.so must travel with the snapshot"
    fn = refapi.RefFn.from_graph(g)
    _, r_leaf, r_status = fn.eval(q)
    assert np.array_equal(leaf[ok], r_leaf[ok]) and np.array_equal(status, r_status)
    ref_rows, scale = _reference_mesh_rows(fn, q[ok] * 2 - 1, leaf[ok], synthetic)
    assert_values_close(got[ok], ref_rows, scale)


def test_c4b_ring_path_6d_cubic_boot2():
    """The c4b bench path (loco_eval_batch_ring_device: whole batch sorted once, 64-query tiles streamed through a ring,
    per-query checksum) on the c4b TREE (6-D cubic, bootstrap 2, evalCorners=false: 4,096 leaves, 117,649 samples sharing
    15,625 value rows) at a reduced D; checksums against the kept results of the same handle and against the CPU
    restatement on a sample of queries.  (The reference's loader needs ~10 min for this tree; its weights pin the same
    kernel instantiation on the bootstrap-1 tree in test_c4_shape_6d_cubic_mesh_against_reference_weights.)"""
    torch = _torch()
    N, level, D = 6, 2, 130
    g, twin = _scalar_twin_graph(N, capi.CUBIC, level, False)
    mesh = capi.Model.bootstrap(N, capi.CUBIC, level, D, False, fn=None, device=0)
    rows = mesh.sample_value_rows()
    rng = np.random.default_rng(63)
    table = rng.standard_normal((mesh.info["n_value_rows"], D))
    mesh.set_values(table)
    Q = 20_000
    q = rng.random((Q, N))
    q[:Q // 2] = q[:Q // 2] * 0.26   # half of the batch in the corner leaves (folded rows, tiles of one leaf fill up)
    tq = torch.from_numpy(q).cuda()
    st = torch.cuda.current_stream().cuda_stream
    out = torch.empty((Q, D), dtype=torch.float64, device="cuda")
    leaf = torch.empty(Q, dtype=torch.int32, device="cuda")
    status = torch.empty(Q, dtype=torch.uint8, device="cuda")
    mesh.eval_device(tq.data_ptr(), Q, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), st)
    ring = torch.empty((1024, D), dtype=torch.float64, device="cuda")
    chk = torch.empty(Q, dtype=torch.float64, device="cuda")
    mesh.profile_begin()
    mesh.eval_ring_device(tq.data_ptr(), Q, ring.data_ptr(), 1024, chk.data_ptr(), 0, 0, st)
    prof = mesh.profile_end()
    assert "mesh_tensor_mma" in prof and prof["mesh_tensor_mma"]["launches"] > 1
    out, leaf, status, chk = out.cpu().numpy(), leaf.cpu().numpy(), status.cpu().numpy(), chk.cpu().numpy()
    assert (status == 0).all()
    wts = 1.0 + np.arange(D) % 7
    ref_chk = (out * wts).sum(axis=1)
    assert (np.abs(chk - ref_chk) <= 1e-12 * (np.abs(out) * wts).sum(axis=1)).all()
    n = 300
    pick = np.concatenate([np.arange(n // 2), Q // 2 + np.arange(n // 2)])
    orc = oracleapi.Oracle(_mesh_graph(g, table[rows]))
    o_out, o_leaf, o_status, o_scale = orc.eval(q[pick])
    assert np.array_equal(leaf[pick], o_leaf) and (o_status == 0).all()
    assert_values_close(out[pick], o_out, o_scale)
    o_chk = (o_out * wts).sum(axis=1)
    assert (np.abs(chk[pick] - o_chk) <= 1e-12 * o_scale * wts.sum()).all()


def _splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _generated_points(m, ppl, seed):
    """host regeneration of the device's jitter points (contract in include/loco_b200.h); returns (cube points [L*ppl, N], leaf [L*ppl])"""
    L, N = m.L, m.N
    boxes = np.stack([m.leaf_box(l) for l in range(L)])  # [L, N, 2]
    p = np.arange(L * ppl, dtype=np.uint64)
    leaf = (p // np.uint64(ppl)).astype(np.int64)
    pts = np.empty((L * ppl, N))
    with np.errstate(over="ignore"):
        for d in range(N):
            h = _splitmix64(np.uint64(seed) ^ _splitmix64(p * np.uint64(N) + np.uint64(d)))
            r = ((h >> np.uint64(40)).astype(np.float32) / np.float32(16777215.0)).astype(np.float64)
            lo, hi = boxes[leaf, d, 0], boxes[leaf, d, 1]
            pts[:, d] = lo + r * (hi - lo)
    return pts, leaf.astype(np.int32), boxes


@pytest.mark.parametrize("N,basis,level,ppl", [(8, capi.LINEAR, 1, 300), (3, capi.CUBIC, 3, 700), (2, capi.LINEAR, 4, 1000)])
def test_generated_error_check(N, basis, level, ppl):
    """loco_error_check_generated_device (RANDOM_SAMPLES metric at scale, BASELINE configs[4]): the points are regenerated on the
    host from the documented counter-based generator; each must lie in its leaf, the truth must be the reference's sine sum
    (LCRealFunction.cpp:50-63) to 1e-12, and err_max / split must equal a host recomputation with the CPU restatement's
    in-cell evaluation -- for the fused kernel and for the three-kernel form -- and loco_error_check on the same points."""
    torch = _torch()
    rng = np.random.default_rng(5)
    amp, phase = rng.uniform(-1, 1, (N, 10)), rng.uniform(-1, 1, (N, 10))
    j = np.arange(10)

    def f(p):
        return 1.0 + float((amp * np.sin(3 * j[None, :] * np.asarray(p)[:, None] + phase)).sum())
    dom = np.array([[0.0, 1.0], [-2.0, 3.0]] * N)[:N]   # a non-trivial domain exercises mapFromStandardHypercube
    m = capi.Model.bootstrap(N, basis, level, 1, True, domain=dom.reshape(-1), fn=f, device=0)
    seed, thr = 1234567, 0.05
    pts, leaf, boxes = _generated_points(m, ppl, seed)
    assert (pts >= boxes[leaf, :, 0]).all() and (pts <= boxes[leaf, :, 1]).all(), "a generated point left its leaf"
    alpha = (pts + 1.0) / 2.0
    xo = dom[:, 0] * (1 - alpha) + dom[:, 1] * alpha
    truth = 1.0 + (amp[None] * np.sin(3 * j[None, None, :] * xo[:, :, None] + phase[None])).sum(axis=(1, 2))
    tmp = tempfile.mktemp(suffix=".pb")
    m.save_proto(tmp)
    orc = oracleapi.Oracle(protoschema.read_binary(tmp))
    os.unlink(tmp)
    approx, st_in = orc.eval_in_cell(pts, leaf)
    assert (st_in == 0).all()
    err_pt = np.abs(approx - truth)
    want_err = np.zeros(m.L)
    np.maximum.at(want_err, leaf, err_pt)
    want_split = np.zeros(m.L, dtype=bool)
    np.logical_or.at(want_split, leaf, ~((err_pt - thr) <= 0))
    d_amp, d_phase = torch.from_numpy(amp).cuda(), torch.from_numpy(phase).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    for fused in (1, 0):
        m.set_option("fused_errgen", fused)
        err = torch.full((m.L,), -1.0, dtype=torch.float64, device="cuda")
        split = torch.full((m.L,), 7, dtype=torch.uint8, device="cuda")
        m.profile_begin()
        m.error_check_generated_device(ppl, seed, d_amp.data_ptr(), d_phase.data_ptr(), 10, thr, err.data_ptr(), split.data_ptr(), stream)
        prof = m.profile_end()
        assert ("error_check_generated" in prof) == bool(fused), prof
        err, split = err.cpu().numpy(), split.cpu().numpy().astype(bool)
        assert np.abs(err - want_err).max() <= 1e-12 * max(1.0, np.abs(truth).max()), (fused, np.abs(err - want_err).max())
        flips = split != want_split  # a flag may only differ where the error sits within rounding of the threshold
        assert (np.abs(want_err[flips] - thr) < 1e-12).all(), fused
    # the located check on the same points in original coordinates: only leaves none of whose points sits on a leaf face
    # can be compared (a point on a shared face is located in the neighbouring leaf)
    interior = (pts > boxes[leaf, :, 0]).all(axis=1) & (pts < boxes[leaf, :, 1]).all(axis=1)
    clean = np.ones(m.L, dtype=bool)
    clean[leaf[~interior]] = False
    e2, s2, bad = m.error_check(xo, truth, thr)
    assert bad == 0
    assert np.abs(e2[clean] - want_err[clean]).max() <= 1e-11 * max(1.0, np.abs(truth).max())


@pytest.mark.parametrize("N,basis,level", [(3, capi.CUBIC, 5), (2, capi.CUBIC, 5), (4, capi.LINEAR, 3)])
def test_optional_fp32_mode_bucket_path(N, basis, level):
    """north star: 1e-5 relative in the optional fp32 mode -- on the leaf-bucket path (path 5, the default dispatch of the c2
    workload: bucket_eval_kernel<..., FP32=true>), against the CPU restatement; leaf indices stay bit-exact."""
    torch = _torch()

    def f(p):
        return 1 + np.sin(3 * p + np.arange(len(p))).sum()
    m = capi.Model.bootstrap(N, basis, level, 1, True, fn=f, device=0)
    Q = 1_500_000
    tq = torch.rand((Q, N), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(19)) * 1.000002 - 0.000001
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for path in (5, 0):
        for fp32 in (0, 1):
            m.set_option("path", path)
            m.set_option("fp32", fp32)
            out = torch.empty(Q, dtype=torch.float64, device="cuda")
            leaf = torch.empty(Q, dtype=torch.int32, device="cuda")
            status = torch.empty(Q, dtype=torch.uint8, device="cuda")
            m.profile_begin()
            m.eval_device(tq.data_ptr(), Q, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), st)
            prof = m.profile_end()
            assert "bucket_eval" in prof, (path, prof)   # the automatic dispatch picks the buckets at this size too
            res[(path, fp32)] = (out.cpu().numpy(), leaf.cpu().numpy(), status.cpu().numpy())
    a, b = res[(5, 0)], res[(5, 1)]
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    ok = a[2] == 0
    diff = np.abs(a[0][ok] - b[0][ok])
    assert diff.max() > 0, "the fp32 mode did not run"
    assert (diff <= 1e-5 * np.maximum(np.abs(a[0][ok]), 1.0)).all(), diff.max()
    assert np.array_equal(res[(0, 1)][0], b[0], equal_nan=True)
    # against the CPU restatement on a sample
    tmp = tempfile.mktemp(suffix=".pb")
    m.save_proto(tmp)
    orc = oracleapi.Oracle(protoschema.read_binary(tmp))
    os.unlink(tmp)
    n = 20_000
    o_out, o_leaf, o_status, scale = orc.eval(tq[:n].cpu().numpy())
    assert np.array_equal(b[1][:n], o_leaf) and np.array_equal(b[2][:n], o_status)
    oko = o_status == 0
    assert_values_close(a[0][:n][oko], o_out[oko], scale[oko])
    assert (np.abs(b[0][:n][oko] - o_out[oko]) <= 1e-5 * np.maximum(np.abs(o_out[oko]), scale[oko])).all()


def test_nan_query_on_the_bucket_path():
    """A NaN coordinate: the reference descends to child2 everywhere, passes the containment test (every comparison is
    false) and every basis factor of that dimension is 0 (`d <= 1` false) -> value 0.0, status OK.  The uniform-knot fast
    path of the cubic tensor leaves must not turn that into NaN."""
    m = capi.Model.bootstrap(3, capi.CUBIC, 4, 1, True, fn=lambda p: 1 + float(np.sin(3 * p).sum()), device=0)
    tmp = tempfile.mktemp(suffix=".pb")
    m.save_proto(tmp)
    orc = oracleapi.Oracle(protoschema.read_binary(tmp))
    os.unlink(tmp)
    rng = np.random.default_rng(3)
    q = rng.random((300_000, 3))
    nan_rows = rng.choice(len(q), 500, replace=False)
    q[nan_rows, rng.integers(0, 3, 500)] = np.nan
    q[nan_rows[:20]] = np.nan
    o_out, o_leaf, o_status, scale = orc.eval(q[nan_rows])
    assert (o_status == 0).all() and (o_out == 0.0).all()
    for path in (1, 4, 5):
        m.set_option("path", path)
        out, leaf, status = m.eval(q)
        assert np.array_equal(leaf[nan_rows], o_leaf) and (status[nan_rows] == 0).all(), path
        assert (out[nan_rows] == 0.0).all(), path


def test_save_proto_follows_set_values(tmp_path):
    """loco_model_save_proto must serialise the values the kernels evaluate, also after loco_model_set_values /
    loco_model_fill_values_synthetic replaced the table on the device."""
    m = capi.Model.bootstrap(2, capi.CUBIC, 2, 1, True, fn=lambda p: float(p.sum()), device=0)
    rng = np.random.default_rng(1)
    new = rng.standard_normal((m.info["n_value_rows"], 1))
    m.set_values(new)
    q = rng.random((2000, 2))
    want, _, _ = m.eval(q)
    p1 = str(tmp_path / "a.pb")
    m.save_proto(p1)
    got, _, _ = capi.Model.from_proto(p1, device=0).eval(q)
    assert np.array_equal(want, got)
    orc = oracleapi.Oracle(protoschema.read_binary(p1))
    o_out, _, _, scale = orc.eval(q)
    assert_values_close(want, o_out, scale)
    # mesh-valued, table filled on the device only: the file carries the synthetic table
    mesh = capi.Model.bootstrap(2, capi.LINEAR, 2, 6, True, fn=lambda p: np.arange(6) + float(p.sum()), device=0)
    mesh.fill_values_synthetic(0.75)
    p2 = str(tmp_path / "b.pb")
    mesh.save_proto(p2)
    a, _, _ = mesh.eval(q)
    b, _, _ = capi.Model.from_proto(p2, device=0).eval(q)
    assert np.array_equal(a, b)


def test_error_check_nan_component_is_sticky():
    """with D > 1 a NaN error component must survive the later finite components of the same point
    (err_max NaN, split set) -- the all_reduce(MAX) over ranks relies on NaN having the largest bit pattern."""
    D = 3
    m = capi.Model.bootstrap(2, capi.LINEAR, 2, D, True, fn=lambda p: np.array([p.sum(), p[0], 1.0]), device=0)
    rng = np.random.default_rng(2)
    q = rng.random((4000, 2))
    out, leaf, status = m.eval(q)
    truth = out.copy()
    truth[17, 0] = np.nan      # component 0 is NaN, components 1 and 2 are exact
    err, split, bad = m.error_check(q, truth, 10.0)
    assert bad == 0
    assert np.isnan(err[leaf[17]]) and split[leaf[17]] == 1
    others = np.arange(m.L) != leaf[17]
    assert (err[others] <= 1e-13).all() and (split[others] == 0).all()


def test_adaptive_dispatch_between_buckets_and_counting_sort():
    """Automatic path selection for big scalar batches on tensor-product leaves: uniform batches take the single-pass leaf
    buckets; once most of the batches seen so far (> 65 %) overflows the buckets' capacity (lattice sweeps, trajectories) the next calls take the exact
    counting sort, and go back when the batches turn uniform again.  Decided from counters the kernels keep (no
    synchronisation), so a switch takes effect a call later; results are the same on every path."""
    torch = _torch()
    m = capi.Model.bootstrap(3, capi.CUBIC, 4, 1, True, fn=lambda p: 1 + float(np.sin(3 * p).sum()), device=0)
    Q = 2_000_000
    gen = torch.Generator("cuda").manual_seed(31)
    uniform = torch.rand((Q, 3), dtype=torch.float64, device="cuda", generator=gen)
    skewed = uniform.clone()
    sel = torch.arange(Q, device="cuda") % 10 != 0
    skewed[sel] = skewed[sel] * 0.1 + 0.45   # nine tenths of the batch in 0.1 % of the domain (the switch needs > 65 % overflow)
    st = torch.cuda.current_stream().cuda_stream

    def run(tq):
        out = torch.empty(Q, dtype=torch.float64, device="cuda")
        leaf = torch.empty(Q, dtype=torch.int32, device="cuda")
        status = torch.empty(Q, dtype=torch.uint8, device="cuda")
        m.profile_begin()
        m.eval_device(tq.data_ptr(), Q, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), st)
        prof = m.profile_end()   # synchronises: the counters of this call are visible to the next one
        return out, leaf, prof

    m.set_option("path", 1)
    ref_u, leaf_u, _ = run(uniform)
    ref_s, leaf_s, _ = run(skewed)
    m.set_option("path", 0)
    o, l, prof = run(uniform)
    assert "bucket_eval" in prof and torch.equal(o, ref_u) and torch.equal(l, leaf_u)
    o, l, prof = run(skewed)                      # decided before its own statistics exist: still the buckets (overflow list)
    assert "bucket_eval" in prof and torch.equal(o, ref_s) and torch.equal(l, leaf_s)
    o, l, prof = run(skewed)                      # the previous call overflowed: counting sort
    assert "sorted_eval" in prof and "bucket_eval" not in prof, prof
    assert torch.equal(o, ref_s) and torch.equal(l, leaf_s)
    o, l, prof = run(uniform)                     # the previous (sorted) call was still skewed
    assert "sorted_eval" in prof and torch.equal(o, ref_u)
    o, l, prof = run(uniform)                     # ... and the last one was uniform: back to the buckets
    assert "bucket_eval" in prof and torch.equal(o, ref_u)
    m.set_option("adaptive", 0)
    o, l, prof = run(skewed)
    o, l, prof = run(skewed)
    assert "bucket_eval" in prof and torch.equal(o, ref_s)
