"""Parity of the CUDA path (through the C ABI) with the reference: committed golden vectors produced by the
reference's own code, the CPU restatement in oracle/, and size-independent properties at larger sizes.

Bars (BASELINE.json north_star): leaf indices, status bytes and support sets bit-exact; fp64 values within
1e-12 relative (see conftest.VALUE_RTOL for what "relative" is taken against)."""
import os

import numpy as np
import pytest

import protoschema
from conftest import GOLDEN, GOLDEN_CASES, assert_values_close
from locointerpolate_b200 import capi
from oracle import oracleapi

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.mark.parametrize("path", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_vectors(golden, name, path):
    pb, z = golden(name)
    m = capi.Model.from_proto(pb, device=0)
    m.set_option("path", path)
    orc = oracleapi.Oracle(protoschema.read_binary(pb))
    out, leaf, status = m.eval(z["q"])
    assert m.kernel_launches > 0
    assert np.array_equal(leaf, z["leaf"]), "containing leaf differs from the reference"
    assert np.array_equal(status, z["status"])
    ok = status == 0
    o_out, _, _, scale = orc.eval(z["q"])
    assert_values_close(out[ok], z["out"][ok], scale[ok])   # vs the reference's own output
    assert_values_close(out[ok], o_out[ok], scale[ok])      # vs the CPU restatement
    assert np.isnan(out[~ok]).all()
    for l in range(m.L):
        assert np.array_equal(m.support_set(l), z["sup_ids"][z["sup_off"][l]:z["sup_off"][l + 1]])
    err, split = m.center_error(0.05)
    assert np.allclose(err, z["center_err"], rtol=0, atol=1e-13)
    assert np.array_equal(split.astype(bool), ~((z["center_err"] - 0.05) <= 0))


@pytest.mark.parametrize("name", ["c1_lin2d_boot2", "cub2d_adapt6", "cub3d_boot2"])
def test_device_buffers_and_ragged_sizes(golden, name):
    torch = _torch()
    pb, z = golden(name)
    m = capi.Model.from_proto(pb, device=0)
    orc = oracleapi.Oracle(protoschema.read_binary(pb))
    rng = np.random.default_rng(5)
    for Q in (0, 1, 31, 257, 4099):
        q = rng.random((Q, m.N))
        tq = torch.from_numpy(q).cuda()
        out = torch.empty(Q, dtype=torch.float64, device="cuda")
        leaf = torch.empty(Q, dtype=torch.int32, device="cuda")
        status = torch.empty(Q, dtype=torch.uint8, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        m.eval_device(tq.data_ptr(), Q, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), stream)
        torch.cuda.synchronize()
        o_out, o_leaf, o_status, scale = orc.eval(q)
        assert np.array_equal(leaf.cpu().numpy(), o_leaf) and np.array_equal(status.cpu().numpy(), o_status)
        assert_values_close(out.cpu().numpy(), o_out, scale)
        # leaf / status outputs are optional
        out2 = torch.empty_like(out)
        m.eval_device(tq.data_ptr(), Q, out2.data_ptr(), 0, 0, stream)
        torch.cuda.synchronize()
        assert torch.equal(out, out2)


@pytest.mark.parametrize("name,D", [("c1_lin2d_boot2", 6), ("cub2d_adapt6", 7), ("cub2d_boot2", 5), ("cub3d_boot2", 30), ("cub3d_boot1_nocorner", 131), ("lin4d_boot1", 300)])
def test_mesh_valued(golden, name, D):
    """D > 1: per-vertex mesh vectors.  Oracle = LCRealFunctionValue::add applied component-wise."""
    pb, z = golden(name)
    g = protoschema.read_binary(pb)
    rng = np.random.default_rng(11)
    g.D = D
    g.sample_value = rng.standard_normal((len(g.sample_center), D))
    g.cell_sample_value = g.border_sample_value = g.cell_center_value = g.border_center_value = None
    g.normalise()
    m = capi.Model.from_graph(g, device=0)
    orc = oracleapi.Oracle(g)
    q = z["q"][:400]
    o_out, o_leaf, o_status, scale = orc.eval(q)
    for simt in (0, 1):  # fp64 MMA tiles (default) and DFMA register tiles
        m.set_option("mesh_simt", simt)
        out, leaf, status = m.eval(q)
        assert np.array_equal(leaf, o_leaf) and np.array_equal(status, o_status)
        ok = status == 0
        assert_values_close(out[ok], o_out[ok], scale[ok])
        assert np.isnan(out[~ok]).all()


def test_error_check_matches_oracle(golden):
    pb, z = golden("cub2d_adapt6")
    m = capi.Model.from_proto(pb, device=0)
    orc = oracleapi.Oracle(protoschema.read_binary(pb))
    q, truth = z["q"], z["true_values"]
    for thr in (0.01, 0.1):
        err, split, bad = m.error_check(q, truth, thr)
        o_err, o_split, o_bad = orc.error_check(q, truth, thr)
        assert bad == o_bad == 16
        assert np.allclose(err, o_err, rtol=0, atol=1e-12)
        flips = split != o_split  # a flag may only differ where the error sits within rounding of the threshold
        assert (np.abs(o_err[flips] - thr) < 1e-12).all()
    # NaN truth => split (decideSplit is written as !(max(diff - allowed) <= 0))
    truth2 = truth.copy()
    truth2[200] = np.nan
    err, split, _ = m.error_check(q, truth2, 10.0)
    _, leaf, _, _ = orc.eval(q[200:201])
    assert split[leaf[0]] == 1 and split.sum() == 1 and np.isnan(err[leaf[0]])


@pytest.mark.parametrize("N,basis,level", [(3, capi.CUBIC, 3), (4, capi.LINEAR, 2), (6, capi.LINEAR, 1), (2, capi.CUBIC, 5)])
def test_bootstrap_models_against_oracle(N, basis, level, tmp_path):
    def f(p):
        return 1 + np.sin(3 * p + np.arange(len(p))).sum()
    m = capi.Model.bootstrap(N, basis, level, 1, True, fn=f, device=0)
    path = str(tmp_path / "m.pb")
    m.save_proto(path)
    orc = oracleapi.Oracle(protoschema.read_binary(path))
    q = np.random.default_rng(2).random((3000, N)) * 1.000002 - 0.000001
    out, leaf, status = m.eval(q)
    o_out, o_leaf, o_status, scale = orc.eval(q)
    assert np.array_equal(leaf, o_leaf) and np.array_equal(status, o_status)
    assert_values_close(out, o_out, scale)


@pytest.mark.parametrize("N,basis,level,Q", [(3, capi.CUBIC, 5, 4_000_000), (4, capi.LINEAR, 3, 2_000_000), (6, capi.CUBIC, 1, 200_000),
                                             (8, capi.LINEAR, 1, 500_000)])
def test_full_size_properties(N, basis, level, Q):
    """BASELINE-size trees, checked through properties that need no oracle run:
    linear precision (f = 1 + sum (i+1) x_i is reproduced; TestExamples.cpp:313-448) and partition of
    unity (LCAdaptiveGridCell.cpp:655) -- both to 1e-12 instead of the reference's 1e-5 / 1e-6."""
    torch = _torch()
    coef = np.arange(1, N + 1, dtype=np.float64)
    m = capi.Model.bootstrap(N, basis, level, 1, True, fn=lambda p: 1 + (coef * p).sum(), device=0)
    tq = torch.rand((Q, N), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    out = torch.empty(Q, dtype=torch.float64, device="cuda")
    leaf = torch.empty(Q, dtype=torch.int32, device="cuda")
    status = torch.empty(Q, dtype=torch.uint8, device="cuda")
    m.eval_device(tq.data_ptr(), Q, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int(status.sum()) == 0
    exact = 1 + (tq * torch.from_numpy(coef).cuda()).sum(dim=1)
    assert float((out - exact).abs().max()) < 1e-12 * (1 + coef.sum())
    assert int(leaf.min()) == 0 and int(leaf.max()) == m.L - 1   # every leaf is hit, indices in range
    m.set_values(np.ones((m.info["n_value_rows"], 1)))
    m.eval_device(tq.data_ptr(), Q, out.data_ptr(), 0, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert float((out - 1).abs().max()) < 1e-12


def test_mesh_ring_checksum():
    """Mesh batches too large to keep are streamed through a ring; the per-query checksum must equal the
    checksum of the directly evaluated result."""
    torch = _torch()
    N, D, Q = 4, 3000, 1000
    m = capi.Model.bootstrap(N, capi.LINEAR, 2, D, True, device=0)
    m.fill_values_synthetic(0.25)
    tq = torch.rand((Q, N), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    out = torch.empty((Q, D), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    m.eval_device(tq.data_ptr(), Q, out.data_ptr(), 0, 0, st)
    ring = torch.empty((128, D), dtype=torch.float64, device="cuda")
    chk = torch.empty(Q, dtype=torch.float64, device="cuda")
    m.eval_ring_device(tq.data_ptr(), Q, ring.data_ptr(), 128, chk.data_ptr(), 0, 0, st)
    torch.cuda.synchronize()
    wts = (1 + torch.arange(D, device="cuda") % 7).double()
    ref = (out * wts).sum(dim=1)
    assert float(((chk - ref).abs() / ref.abs().clamp_min(1)).max()) < 1e-12
    # the synthetic table is what the header says it is
    v = m.get_values(5, 2)
    j = np.arange(D)
    assert np.allclose(v[0], np.sin(0.37 * j + 0.61 * 5 + 0.25) + 0.001 * (j % 97), atol=1e-12)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_derivative_golden_vectors(golden, name):
    """LCSampledFunction::evalDeriv batch vs the reference's own output and vs the restatement."""
    pb, z = golden(name)
    ref = np.load(os.path.join(GOLDEN, name + ".deriv.npz"))["deriv"]
    m = capi.Model.from_proto(pb, device=0)
    orc = oracleapi.Oracle(protoschema.read_binary(pb))
    for d in range(m.N):
        o_out, o_leaf, o_status, scale = orc.eval_deriv(z["q"], d)
        for path in (1, 5, 4):  # thread per query; leaf buckets / sorted term lists with the derivative factor (cubic basis, else falls back)
            m.set_option("path", path)
            out, leaf, status = m.eval_deriv(z["q"], d)
            assert np.array_equal(leaf, z["leaf"]) and np.array_equal(status, z["status"])
            ok = status == 0
            assert_values_close(out[ok], ref[d][ok], scale[ok])
            assert_values_close(out[ok], o_out[ok], scale[ok])
            assert np.isnan(out[~ok]).all()
    with pytest.raises(capi.LocoError):
        m.eval_deriv(z["q"], m.N)


def _adversarial_queries(N, n, seed):
    """uniform points plus points on / one ulp either side of dyadic planes (where split planes and grid boundaries lie)"""
    rng = np.random.default_rng(seed)
    q = rng.random((n, N)) * 1.00001 - 0.000005
    k = n // 2
    planes = rng.integers(0, 65, (k, N)) / 64.0
    nudge = rng.integers(-1, 2, (k, N))
    q[:k] = np.where(nudge == 0, planes, np.nextafter(planes, np.where(nudge > 0, 2.0, -1.0)))
    q[k] = np.nan
    q[k + 1, 0] = np.nan
    q[k + 2, -1] = np.inf
    q[k + 3, 0] = -np.inf
    return q


@pytest.mark.parametrize("name", ["cub2d_adapt6", "lin3d_adapt5", "lin2d_adapt6", "cub3d_boot2", "lin4d_boot1", "cub1d_linprec"])
def test_locate_grid_equals_descent_golden(golden, name):
    """The locate grid (Flat::lut) must name exactly the leaf the k-d descent from the root reaches."""
    pb, z = golden(name)
    m = capi.Model.from_proto(pb, device=0)
    q = np.concatenate([z["q"], _adversarial_queries(m.N, 200_000, 3)])
    for path in (1, 2, 4, 5):
        m.set_option("path", path)
        m.set_option("chunk", 50_000)
        m.set_option("lut", 0)
        out0, leaf0, status0 = m.eval(q)
        m.set_option("lut", 1)
        out1, leaf1, status1 = m.eval(q)
        assert np.array_equal(leaf0, leaf1) and np.array_equal(status0, status1)
        assert np.array_equal(out0, out1, equal_nan=True)
    assert np.array_equal(leaf1[:len(z["q"])], z["leaf"])


@pytest.mark.parametrize("N,basis,level", [(3, capi.CUBIC, 5), (2, capi.LINEAR, 9), (6, capi.LINEAR, 2), (8, capi.LINEAR, 1)])
def test_locate_grid_equals_descent_bootstrap(N, basis, level):
    m = capi.Model.bootstrap(N, basis, level, 1, True, fn=lambda p: float(p.sum()), device=0)
    q = _adversarial_queries(N, 1_000_000, 4)
    m.set_option("path", 1)
    m.set_option("lut", 0)
    _, leaf0, status0 = m.eval(q)
    m.set_option("lut", 1)
    _, leaf1, status1 = m.eval(q)
    assert np.array_equal(leaf0, leaf1) and np.array_equal(status0, status1)
    assert leaf1.max() == m.L - 1


@pytest.mark.parametrize("basis,N,level,corners", [(capi.CUBIC, 3, 4, True), (capi.CUBIC, 2, 6, False), (capi.CUBIC, 4, 2, True),
                                                   (capi.LINEAR, 3, 4, True), (capi.LINEAR, 2, 6, True), (capi.LINEAR, 5, 2, True)])
def test_derivative_throughput_path_equals_direct(basis, N, level, corners):
    """evalDeriv on tensor-product leaves through the leaf-bucket path (one derivative factor in the contraction; the linear
    basis with the reference's signed comparison, LCBasisFunction.cpp:295-332) against the thread-per-query term-list
    kernel, at a batch size where the automatic dispatch picks the buckets; a skewed part of the batch overflows its buckets."""
    torch = _torch()

    def f(p):
        return 1 + np.sin(3 * p + np.arange(len(p))).sum()
    m = capi.Model.bootstrap(N, basis, level, 1, corners, fn=f, device=0)
    Q = 600_000
    tq = torch.rand((Q, N), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(17)) * 1.000002 - 0.000001
    tq[:100_000] = tq[:100_000] * 0.05 + 0.3  # concentrated in a few leaves
    tq[100_000:100_040, 0] = 1.5               # outside
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for path in (1, 0, 5):
        m.set_option("path", path)
        for d in (0, N - 1):
            out = torch.empty(Q, dtype=torch.float64, device="cuda")
            leaf = torch.empty(Q, dtype=torch.int32, device="cuda")
            status = torch.empty(Q, dtype=torch.uint8, device="cuda")
            m.eval_deriv_device(tq.data_ptr(), Q, d, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), st)
            torch.cuda.synchronize()
            res[(path, d)] = (out.cpu().numpy(), leaf.cpu().numpy(), status.cpu().numpy())
    for d in (0, N - 1):
        a = res[(1, d)]
        assert (a[2] != 0).sum() >= 40 and np.abs(a[0][a[2] == 0]).max() > 0.1
        for path in (0, 5):
            b = res[(path, d)]
            assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
            ok = a[2] == 0
            assert np.isnan(b[0][~ok]).all()
            assert np.abs(a[0][ok] - b[0][ok]).max() <= 1e-12 * max(1.0, np.abs(a[0][ok]).max())
    # the bucket kernels really ran for path 5
    m.set_option("path", 5)
    m.profile_begin()
    m.eval_deriv_device(tq.data_ptr(), Q, 0, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), st)
    prof = m.profile_end()
    assert "bucket_eval" in prof and "bucket_scatter" in prof, prof


@pytest.mark.parametrize("name", ["cub2d_testdebug", "lin3d_adapt5", "cub2d_prototest", "lin2d_adapt6"])
def test_polynomial_evaluation_cells(golden, name):
    """Option "ecell_poly": evaluation cells that no knot crosses are evaluated from their single tensor-product polynomial
    (csrc/loco_cellpoly.h; coefficients formed on the device from the current values).  Same leaves and status bytes,
    values within the 1e-12 bar of the oracle and of the term-list evaluation."""
    pb, z = golden(name)
    m = capi.Model.from_proto(pb, device=0)
    assert m.info["n_poly_cells"] > 0
    dom = m.domain.reshape(m.N, 2)
    rng = np.random.default_rng(43)
    q = np.concatenate([z["q"], dom[:, 0] + (dom[:, 1] - dom[:, 0]) * rng.random((300_000, m.N))])
    m.set_option("path", 4)
    m.set_option("ecell_poly", 0)
    a, la, sa = m.eval(q)
    m.set_option("ecell_poly", 1)
    b, lb, sb = m.eval(q)
    assert np.array_equal(la, lb) and np.array_equal(sa, sb)
    ok = sa == 0
    assert np.isnan(b[~ok]).all()
    assert (a[ok] != b[ok]).any(), "the polynomial cells did not run"
    orc = oracleapi.Oracle(protoschema.read_binary(pb))
    n = 40_000
    o_out, o_leaf, o_status, scale = orc.eval(q[:n])
    okn = o_status == 0
    assert np.array_equal(lb[:n], o_leaf) and np.array_equal(sb[:n], o_status)
    assert_values_close(b[:n][okn], o_out[okn], scale[okn])
    assert np.abs(a[ok] - b[ok]).max() <= 1e-12 * max(1.0, scale[okn].max())


def test_derivative_on_adaptive_tree_sorted_path(golden):
    """evalDeriv on the reference's TestDebug tree (generic leaves, cubic): sorted path keyed by evaluation cell with the
    derivative factor in the cells' term lists, against the thread-per-query kernel and the reference's own derivatives."""
    torch = _torch()
    pb, z = golden("cub2d_testdebug")
    m = capi.Model.from_proto(pb, device=0)
    ref = np.load(os.path.join(GOLDEN, "cub2d_testdebug.deriv.npz"))["deriv"]
    Q = 1_000_000
    tq = torch.rand((Q, 2), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(23))
    tq[:len(z["q"])] = torch.from_numpy(z["q"]).cuda()
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for path in (1, 0):
        m.set_option("path", path)
        for d in (0, 1):
            out = torch.empty(Q, dtype=torch.float64, device="cuda")
            leaf = torch.empty(Q, dtype=torch.int32, device="cuda")
            status = torch.empty(Q, dtype=torch.uint8, device="cuda")
            m.eval_deriv_device(tq.data_ptr(), Q, d, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), st)
            torch.cuda.synchronize()
            res[(path, d)] = (out.cpu().numpy(), leaf.cpu().numpy(), status.cpu().numpy())
    nz = len(z["q"])
    for d in (0, 1):
        a, b = res[(1, d)], res[(0, d)]
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        ok = a[2] == 0
        assert np.abs(a[0][ok] - b[0][ok]).max() <= 1e-11 * np.abs(a[0][ok]).max()
        okz = z["status"] == 0
        assert np.array_equal(b[1][:nz], z["leaf"]) and np.array_equal(b[2][:nz], z["status"])
        assert np.abs(b[0][:nz][okz] - ref[d][okz]).max() <= 1e-11 * np.abs(ref[d][okz]).max()
    m.set_option("path", 0)
    m.profile_begin()
    m.eval_deriv_device(tq.data_ptr(), Q, 0, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), st)
    assert "sorted_eval" in m.profile_end()


@pytest.mark.parametrize("N,level", [(9, 1), (10, 0), (12, 0)])
def test_more_than_eight_parameters_every_path(N, level, tmp_path):
    """N > 8 runs the runtime-N instantiation of every kernel (coordinates in local memory instead of unrolled
    registers): scalar values on every path, derivatives, the centre error check and a mesh-valued function."""
    def f(p):
        return 1 + np.sin(3 * p + np.arange(len(p))).sum()
    m = capi.Model.bootstrap(N, capi.LINEAR, level, 1, True, fn=f, device=0)
    path = str(tmp_path / "m.pb")
    m.save_proto(path)
    orc = oracleapi.Oracle(protoschema.read_binary(path))
    rng = np.random.default_rng(21)
    q = rng.random((8_000, N)) * 1.000002 - 0.000001
    q[:512] = np.round(q[:512] * 2) / 2  # on split planes and sample positions
    q[np.arange(600, 640), rng.integers(0, N, 40)] += 1.5 * rng.choice([-1.0, 1.0], 40)  # outside the domain in one coordinate
    o_out, o_leaf, o_status, scale = orc.eval(q)
    ok = o_status == 0
    assert ok.sum() > 7_000 and (~ok).sum() >= 30
    for p_ in (0, 1, 2, 3, 4, 5):
        m.set_option("path", p_)
        out, leaf, status = m.eval(q)
        assert np.array_equal(leaf, o_leaf) and np.array_equal(status, o_status), p_
        assert_values_close(out[ok], o_out[ok], scale[ok])
        assert np.isnan(out[~ok]).all()
    m.set_option("path", 0)
    for d in (0, N - 1):
        out, leaf, status = m.eval_deriv(q[:4000], d)
        d_out, d_leaf, d_status, d_scale = orc.eval_deriv(q[:4000], d)
        assert np.array_equal(leaf, d_leaf) and np.array_equal(status, d_status)
        assert_values_close(out[d_status == 0], d_out[d_status == 0], d_scale[d_status == 0])
    err, split = m.center_error(0.05)
    o_err, o_split = orc.center_error(0.05)
    assert np.allclose(err, o_err, rtol=0, atol=1e-12) and np.array_equal(split, o_split)
    # mesh-valued on the same tree: component j equals the scalar function scaled by (j + 1)
    D = 5
    mesh = capi.Model.bootstrap(N, capi.LINEAR, level, D, True, fn=lambda p: f(p) * (1 + np.arange(D)), device=0)
    for simt, p_ in ((0, 0), (1, 0), (0, 1)):
        mesh.set_option("mesh_simt", simt)
        mesh.set_option("path", p_)
        mo, ml, ms = mesh.eval(q[:3000])
        assert np.array_equal(ml, o_leaf[:3000]) and np.array_equal(ms, o_status[:3000])
        for j in range(D):
            assert_values_close(mo[ok[:3000], j], o_out[:3000][ok[:3000]] * (j + 1), scale[:3000][ok[:3000]] * (j + 1))


@pytest.mark.parametrize("N,level", [(2, 1), (3, 2), (4, 1)])
def test_mesh_valued_shared_border_rows(N, level):
    """Cubic grids built with evalCorners=false: border samples share the value object of their clamped in-domain
    sample (LCAdaptiveGrid.cpp:345-355), which the fp64-MMA mesh kernel exploits by folding equal rows.  A D-component
    function must equal D scalar functions on the same tree, on every mesh path."""
    D = 4
    fns = [lambda p, j=j: float(np.sin(3 * p * (j + 1) + j).sum() + j) for j in range(D)]
    mesh = capi.Model.bootstrap(N, capi.CUBIC, level, D, False, fn=lambda p: np.array([f(p) for f in fns]), device=0)
    assert mesh.info["n_value_rows"] - mesh.L < mesh.info["n_samples"]  # rows are shared (centre values add L rows)
    q = np.random.default_rng(8).random((5000, N)) * 1.000002 - 0.000001
    outs = {}
    for simt, path in ((0, 0), (1, 0), (0, 1)):  # folded MMA tiles, DFMA tiles, generic weights + gather
        mesh.set_option("mesh_simt", simt)
        mesh.set_option("path", path)
        outs[(simt, path)], leaf, status = mesh.eval(q)
    for j in range(D):
        scalar = capi.Model.bootstrap(N, capi.CUBIC, level, 1, False, fn=fns[j], device=0)
        ref, leaf_s, status_s = scalar.eval(q)
        assert np.array_equal(leaf, leaf_s) and np.array_equal(status, status_s)
        for key, out in outs.items():
            assert np.abs(out[:, j] - ref).max() <= 1e-12 * (np.abs(ref).max() + 1), key


def test_mesh_valued_proto_file(golden, tmp_path):
    """A mesh-valued function stored in the reference's own schema (ShapeInfo.tetMesh.vertices) loads and evaluates."""
    pb, z = golden("cub2d_adapt6")
    g = protoschema.read_binary(pb)
    V = 7
    rng = np.random.default_rng(12)
    g.D = 3 * V
    g.sample_value = rng.standard_normal((len(g.sample_center), 3 * V))
    g.cell_center_value = rng.standard_normal((len(g.cell_split_dir), 3 * V))
    g.border_center_value = rng.standard_normal((len(g.border_ranges), 3 * V))
    g.cell_sample_value = g.border_sample_value = None
    src = str(tmp_path / "mesh.pb")
    protoschema.write_binary(g, src)
    m = capi.Model.from_proto(src, device=0)
    orc = oracleapi.Oracle(protoschema.read_binary(src))
    q = z["q"][:600]
    out, leaf, status = m.eval(q)
    o_out, o_leaf, o_status, scale = orc.eval(q)
    assert np.array_equal(leaf, o_leaf) and np.array_equal(status, o_status)
    ok = status == 0
    assert_values_close(out[ok], o_out[ok], scale[ok])


@pytest.mark.parametrize("name", ["cub3d_boot2", "cub2d_boot2", "cub3d_boot1_nocorner"])
def test_uniform_knot_factors_equal_piecewise(golden, name):
    """The four cubic factors of a dimension formed together from one spline parameter (uniform knots) against four
    separate evaluations of the reference's piecewise form: rounding-level agreement, incl. the +-1e-6 acceptance band."""
    pb, z = golden(name)
    m = capi.Model.from_proto(pb, device=0)
    orc = oracleapi.Oracle(protoschema.read_binary(pb))
    o_out, _, o_status, scale = orc.eval(z["q"])
    ok = o_status == 0
    for path in (1, 4):
        m.set_option("path", path)
        m.set_option("uniform", 0)
        a, _, _ = m.eval(z["q"])
        m.set_option("uniform", 1)
        b, _, _ = m.eval(z["q"])
        assert_values_close(a[ok], o_out[ok], scale[ok])
        assert_values_close(b[ok], o_out[ok], scale[ok])
        assert np.abs(a[ok] - b[ok]).max() <= 1e-14 * max(1.0, np.abs(o_out[ok]).max())


@pytest.mark.parametrize("N,basis,level", [(3, capi.CUBIC, 4), (2, capi.CUBIC, 5), (4, capi.LINEAR, 3)])
def test_optional_fp32_mode(N, basis, level):
    """north star: interpolated values within 1e-5 relative in the optional fp32 mode; leaf indices stay bit-exact."""
    torch = _torch()

    def f(p):
        return 1 + np.sin(3 * p + np.arange(len(p))).sum()
    m = capi.Model.bootstrap(N, basis, level, 1, True, fn=f, device=0)
    Q = 400_000
    tq = torch.rand((Q, N), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(9))
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for fp32 in (0, 1):
        m.set_option("path", 4)
        m.set_option("fp32", fp32)
        out = torch.empty(Q, dtype=torch.float64, device="cuda")
        leaf = torch.empty(Q, dtype=torch.int32, device="cuda")
        status = torch.empty(Q, dtype=torch.uint8, device="cuda")
        m.eval_device(tq.data_ptr(), Q, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), st)
        torch.cuda.synchronize()
        res[fp32] = (out.cpu().numpy(), leaf.cpu().numpy(), status.cpu().numpy())
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    diff = np.abs(res[0][0] - res[1][0])
    assert diff.max() > 0, "the fp32 mode did not run"
    assert (diff <= 1e-5 * np.maximum(np.abs(res[0][0]), 1.0)).all(), diff.max()


@pytest.mark.parametrize("name", ["cub2d_prototest", "lin2d_adapt6", "lin3d_adapt5", "cub2d_testdebug"])
def test_evaluation_cells_drop_only_zero_terms(golden, name):
    """Sorted path keyed by evaluation cell (Flat::ecell_*): a cell's term list may only leave out terms whose weight is
    exactly zero everywhere in the cell, so the sums must equal the per-leaf sums bit for bit — on random points and on
    points on / next to every dyadic face of every leaf (where the cell index may round either way)."""
    pb, z = golden(name)
    m = capi.Model.from_proto(pb, device=0)
    assert m.info["n_eval_cells"] > m.L and m.info["n_eval_terms"] > 0
    rng = np.random.default_rng(11)
    dom = m.domain.reshape(m.N, 2)
    lo, w = dom[:, 0], dom[:, 1] - dom[:, 0]
    pts = [lo + w * rng.random((200_000, m.N)), z["q"]]
    fr = np.arange(33) / 32.0
    for l in range(m.L):
        box = m.leaf_box(l)  # cube coordinates
        grid = np.stack(np.meshgrid(*[box[d, 0] + (box[d, 1] - box[d, 0]) * fr for d in range(m.N)], indexing="ij"), -1).reshape(-1, m.N)
        if len(grid) > 6000:
            grid = grid[rng.choice(len(grid), 6000, replace=False)]
        for jitter in (0.0, 1e-15, -1e-15, 3e-7, -3e-7):
            cube = grid + jitter * rng.choice([-1.0, 1.0], size=grid.shape)
            pts.append(lo + w * (cube + 1.0) * 0.5)
    q = np.concatenate(pts)
    m.set_option("path", 4)
    m.set_option("ecell_poly", 0)  # term lists only: the polynomial form of a cell changes the summation order
    m.set_option("ecells", 0)
    a, la, sa = m.eval(q)
    m.set_option("ecells", 1)
    b, lb, sb = m.eval(q)
    m.set_option("path", 1)
    c, lc, sc = m.eval(q)
    assert np.array_equal(la, lb) and np.array_equal(sa, sb) and np.array_equal(la, lc) and np.array_equal(sa, sc)
    assert (sa == 0).sum() > 200_000
    assert np.array_equal(a, b, equal_nan=True)
    orc = oracleapi.Oracle(protoschema.read_binary(pb))
    o_out, o_leaf, o_status, scale = orc.eval(q[:50_000])
    ok = o_status == 0
    assert np.array_equal(lb[:50_000], o_leaf) and np.array_equal(sb[:50_000], o_status)
    assert_values_close(b[:50_000][ok], o_out[ok], scale[ok])


def test_bucket_path_overflow_with_skewed_queries():
    """path 5 (single-pass fixed-capacity leaf buckets): a batch concentrated in a few leaves overflows their buckets;
    the overflow list must give exactly the results of the direct path."""
    torch = _torch()
    m = capi.Model.bootstrap(3, capi.CUBIC, 4, 1, True, fn=lambda p: 1 + np.sin(3 * p).sum(), device=0)
    Q = 600_000
    gen = torch.Generator("cuda").manual_seed(21)
    tq = torch.rand((Q, 3), dtype=torch.float64, device="cuda", generator=gen)
    tq[: Q // 2] = tq[: Q // 2] * 0.05 + 0.4     # half of the batch inside a 5 % cube (a handful of leaves)
    tq[Q // 2: Q // 2 + 1000] = 1.5              # and some queries outside the domain
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    for path in (1, 5):
        m.set_option("path", path)
        m.set_option("chunk", 200_000)
        out = torch.empty(Q, dtype=torch.float64, device="cuda")
        leaf = torch.empty(Q, dtype=torch.int32, device="cuda")
        status = torch.empty(Q, dtype=torch.uint8, device="cuda")
        m.eval_device(tq.data_ptr(), Q, out.data_ptr(), leaf.data_ptr(), status.data_ptr(), st)
        torch.cuda.synchronize()
        res[path] = (out.cpu().numpy(), leaf.cpu().numpy(), status.cpu().numpy())
    assert np.array_equal(res[1][1], res[5][1]) and np.array_equal(res[1][2], res[5][2])
    assert int((res[5][2] != 0).sum()) == 1000
    assert np.array_equal(res[1][0], res[5][0], equal_nan=True)  # same arithmetic per query on both paths
