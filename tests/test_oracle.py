"""The oracle (oracle/loco_oracle.cpp) pinned against the reference: committed golden vectors produced by
the reference's own code, and -- where oracle/_ref/liblocoref.so is present -- the compiled reference live."""
import os

import numpy as np
import pytest

import protoschema
from conftest import GOLDEN, GOLDEN_CASES, assert_values_close
from oracle import oracleapi, refapi


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_golden(golden, name):
    pb, z = golden(name)
    orc = oracleapi.Oracle(protoschema.read_binary(pb))
    out, leaf, status, scale = orc.eval(z["q"])
    assert np.array_equal(leaf, z["leaf"]), "containing leaf differs from the reference"
    assert np.array_equal(status, z["status"])
    ok = status == 0
    assert ok.sum() > 1000 and (~ok).sum() == 16
    assert_values_close(out[ok], z["out"][ok], scale[ok])
    assert np.isnan(out[~ok]).all()
    # influencing-sample sets and neighbour counts, leaf by leaf (bit-exact objects)
    for l in range(orc.L):
        assert np.array_equal(orc.support(l), z["sup_ids"][z["sup_off"][l]:z["sup_off"][l + 1]]), f"support set of leaf {l}"
        assert sum(orc.neighbors(l)) == z["n_neighbors"][l]
    # CENTER_BASED error metric of decideSplit
    err, split = orc.center_error(0.05)
    assert np.allclose(err, z["center_err"], rtol=0, atol=1e-13)
    assert np.array_equal(split.astype(bool), ~((z["center_err"] - 0.05) <= 0))


@pytest.mark.parametrize("name", ["cub2d_prototest", "cub1d_linprec", "lin2d_prototest"])
def test_linear_precision(golden, name):
    """Known-answer property the reference tests check (TestExamples.cpp:313-448, bound 1e-5 there):
    the interpolant of f = 1 + sum (i+1) x_i reproduces it."""
    pb, z = golden(name)
    orc = oracleapi.Oracle(protoschema.read_binary(pb))
    out, _, status, _ = orc.eval(z["q"])
    # inside the domain proper: within the +-EPSILON acceptance band outside it the hats are clipped
    ok = (status == 0) & (z["q"] >= 0).all(axis=1) & (z["q"] <= 1).all(axis=1)
    N = z["q"].shape[1]
    exact = 1 + (z["q"] * np.arange(1, N + 1)).sum(axis=1)
    assert np.abs(out[ok] - exact[ok]).max() < 1e-12


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_partition_of_unity(golden, name):
    """LCAdaptiveGridCell.cpp:655: the weights of a query sum to one.  With all sample values set to 1 the
    interpolant must be 1 everywhere inside the domain."""
    pb, z = golden(name)
    g = protoschema.read_binary(pb)
    g.sample_value = np.ones_like(g.sample_value)
    g.cell_sample_value = None
    g.border_sample_value = None
    orc = oracleapi.Oracle(g)
    out, _, status, _ = orc.eval(z["q"])
    assert np.abs(out[status == 0] - 1).max() < 2e-6  # the reference's own tolerance (EPSILON), also in the acceptance band
    inside = (status == 0) & (z["q"] >= 0).all(axis=1) & (z["q"] <= 1).all(axis=1)
    assert np.abs(out[inside] - 1).max() < 1e-12


@pytest.mark.skipif(not refapi.available(), reason="oracle/_ref/liblocoref.so not built (needs /root/reference)")
@pytest.mark.parametrize("cfg", [
    dict(N=2, basis=1, linear_fn=False, max_depth=5, threshold=0.2, bootstrap=1),
    dict(N=2, basis=0, linear_fn=False, max_depth=7, threshold=0.05, bootstrap=2),
    dict(N=3, basis=1, linear_fn=False, max_depth=0, threshold=0.1, bootstrap=1, eval_corners=False),
    dict(N=4, basis=1, linear_fn=False, max_depth=0, threshold=0.1, bootstrap=1),
    dict(N=5, basis=0, linear_fn=False, max_depth=0, threshold=0.1, bootstrap=1),
])
def test_oracle_matches_live_reference(cfg):
    g = refapi.RefGrid(**cfg)
    fn = g.sampled_function()
    for f in (fn, fn.reload()):  # in-memory grid and the loader's reconstruction of it
        orc = oracleapi.Oracle(oracleapi.graph_from_ref(f.graph()))
        q = np.random.default_rng(7).random((1500, f.N)) * 1.02 - 0.01
        r_out, r_leaf, r_status = f.eval(q)
        o_out, o_leaf, o_status, scale = orc.eval(q)
        assert np.array_equal(r_leaf, o_leaf) and np.array_equal(r_status, o_status)
        ok = r_status == 0
        assert_values_close(o_out[ok], r_out[ok], scale[ok])
        for l in range(f.L):
            assert np.array_equal(f.support(l), orc.support(l))
            assert f.num_neighbors(l) == sum(orc.neighbors(l))


@pytest.mark.skipif(not refapi.available(), reason="oracle/_ref/liblocoref.so not built")
def test_mesh_valued_restatement_matches_scalar_reference():
    """D > 1 has no reference implementation (LCFunctionValue.cpp:20-31).  The restatement applies
    LCRealFunctionValue::add per component; it is pinned here against the reference's own weights
    (getAllInfluencingSamples + getInterpolationWeight) applied to a D-column value table."""
    g = refapi.RefGrid(2, 1, False, 5, 0.2, 1)
    fn = g.sampled_function()
    ga = oracleapi.graph_from_ref(fn.graph())
    D = 6
    rng = np.random.default_rng(3)
    table = rng.standard_normal((fn.S, D))
    ga.D = D
    ga.sample_value = table
    ga.sample_value_group = None
    ga.cell_sample_value = None
    ga.border_sample_value = None
    ga.cell_center_value = None
    ga.border_center_value = None
    orc = oracleapi.Oracle(ga.normalise())
    q = rng.random((300, 2))
    out, leaf, status, scale = orc.eval(q)
    for i in range(len(q)):
        ids, w = fn.weights(int(leaf[i]), q[i] * 2 - 1)
        ref = (w[:, None] * table[ids]).sum(axis=0)
        assert np.abs(out[i] - ref).max() <= 1e-12 * max(scale[i], 1e-300)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_derivative_matches_golden(golden, name):
    """LCSampledFunction::evalDeriv (LCSampledFunction.cpp:61-67): restatement vs the reference's own output
    (tests/golden/make_golden_deriv.py), every cube direction.  The linear basis is restated literally,
    including the reference's signed-distance test (LCBasisFunction.cpp:295-332)."""
    import os
    pb, z = golden(name)
    ref = np.load(os.path.join(GOLDEN, name + ".deriv.npz"))["deriv"]
    orc = oracleapi.Oracle(protoschema.read_binary(pb))
    for d in range(orc.N):
        out, leaf, status, scale = orc.eval_deriv(z["q"], d)
        assert np.array_equal(leaf, z["leaf"]) and np.array_equal(status, z["status"])
        ok = status == 0
        assert_values_close(out[ok], ref[d][ok], scale[ok])
        assert np.isnan(out[~ok]).all() and np.isnan(ref[d][~ok]).all()


@pytest.mark.parametrize("name", ["cub2d_prototest", "cub1d_linprec"])
def test_cubic_derivative_of_linear_function(golden, name):
    """known answer: the cubic interpolant reproduces f = 1 + sum (i+1) x_i, so d/d(cube_i) = (i+1)/2 on [0,1]^N"""
    pb, z = golden(name)
    orc = oracleapi.Oracle(protoschema.read_binary(pb))
    inside = (z["q"] > 0).all(axis=1) & (z["q"] < 1).all(axis=1)
    for d in range(orc.N):
        out, _, status, _ = orc.eval_deriv(z["q"], d)
        assert np.abs(out[inside & (status == 0)] - (d + 1) / 2).max() < 1e-11


@pytest.mark.skipif(not refapi.available(), reason="oracle/_ref/liblocoref.so not built")
@pytest.mark.parametrize("N,basis,level,corners", [(6, 1, 1, False), (4, 0, 3, True)])
def test_mesh_restatement_at_baseline_shapes(N, basis, level, corners):
    """The mesh oracle of the GPU parity tests (tests/test_gpu_baseline_shapes.py) at the BASELINE trees: 6-D cubic with
    evalCorners=false (4,096 influencing samples per leaf, border samples sharing value rows) and 4-D linear bootstrap 3.
    CPU restatement with D components == the reference's own (sample, weight) lists applied to the value rows."""
    import tempfile
    import os
    import protoschema
    from locointerpolate_b200 import capi
    twin = capi.Model.bootstrap(N, basis, level, 1, corners, fn=lambda p: float(np.sin(3 * p).sum()), device=capi.HOST_ONLY)
    tmp = tempfile.mktemp(suffix=".pb")
    twin.save_proto(tmp)
    g = protoschema.read_binary(tmp)
    os.unlink(tmp)
    fn = refapi.RefFn.from_graph(g)
    rows = twin.sample_value_rows()
    assert len(rows) == fn.S
    D = 5
    rng = np.random.default_rng(4)
    table = rng.standard_normal((int(rows.max()) + 1, D))
    per_sample = table[rows]   # samples sharing a value object carry equal rows
    g.D = D
    g.sample_value = per_sample
    g.sample_value_group = None
    g.cell_sample_value = g.border_sample_value = g.cell_center_value = g.border_center_value = None
    orc = oracleapi.Oracle(g.normalise())
    q = rng.random((40, N)) * 1.0000004 - 0.0000002
    q[:8] = np.round(q[:8] * 4) / 4
    out, leaf, status, scale = orc.eval(q)
    r_out, r_leaf, r_status = fn.eval(q)
    assert np.array_equal(leaf, r_leaf) and np.array_equal(status, r_status) and (status == 0).all()
    for i in range(len(q)):
        ids, w = fn.weights(int(leaf[i]), q[i] * 2 - 1)
        ref = np.zeros(D)
        for k in range(len(ids)):
            ref += w[k] * per_sample[ids[k]]
        assert np.abs(out[i] - ref).max() <= 1e-12 * max(scale[i], 1e-300)
        assert abs(w.sum() - 1) < 1e-12
