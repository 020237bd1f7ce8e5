"""Evaluation cells (Flat::ecell_*) on the host: a cell's term list may leave out a merged basis function of
its leaf only if that function is exactly zero everywhere the cell can be selected — including the +-1e-6 acceptance
band outside the leaf and either side of a cell face, where the rounding of the cell index may go both ways."""
import numpy as np
import pytest

from locointerpolate_b200 import capi

CASES = ["cub2d_prototest", "lin2d_adapt6", "lin3d_adapt5", "cub2d_testdebug"]


def _cell_index(x, lo, hi, bits):
    """the device's rule (loco_sorted.cu eval_cell): k_d = clamp(int((x_d - lo_d) * 2^b_d / (hi_d - lo_d)))"""
    idx, shift = np.zeros(len(x), dtype=np.int64), 0
    for d, b in enumerate(bits):
        if b:
            g = 1 << int(b)
            k = np.clip(((x[:, d] - lo[d]) * g / (hi[d] - lo[d])).astype(np.int64), 0, g - 1)
            idx |= k << shift
            shift += int(b)
    return idx


@pytest.mark.parametrize("name", CASES)
def test_cell_lists_cover_every_nonzero_term(golden, name):
    pb, _ = golden(name)
    m = capi.Model.from_proto(pb, device=capi.HOST_ONLY)
    assert m.info["n_eval_cells"] > m.L
    rng = np.random.default_rng(31)
    leaves = range(m.L) if m.L <= 80 else rng.choice(m.L, 80, replace=False)
    total_listed = total_merged = checked = 0
    for leaf in leaves:
        leaf = int(leaf)
        centre, support, first = m.merged_terms(leaf)
        bits, first_cell, n_cells = m.eval_cells(leaf)
        assert n_cells == 1 << int(bits.sum())
        box = m.leaf_box(leaf)
        lo, hi = box[:, 0], box[:, 1]
        lists = [m.eval_cell_terms(first_cell + c) for c in range(n_cells)]
        for ids in lists:
            assert ((ids >= first) & (ids < first + len(centre))).all() and len(np.unique(ids)) == len(ids)
            total_listed += len(ids)
        total_merged += len(centre) * n_cells
        if n_cells == 1:
            assert len(lists[0]) == len(centre)  # an undivided leaf keeps its whole list
            continue
        # points: uniform in the leaf, in the acceptance band around it, and on / next to every possible cell face
        pts = [lo + (hi - lo) * rng.random((3000, m.N)), lo - 1e-6 + (hi - lo + 2e-6) * rng.random((500, m.N))]
        face = lo + (hi - lo) * rng.integers(0, 33, (3000, m.N)) / 32.0
        pts += [face, np.nextafter(face, -np.inf), np.nextafter(face, np.inf), face + 3e-7 * rng.choice([-1.0, 1.0], face.shape)]
        x = np.concatenate(pts)
        cell = _cell_index(x, lo, hi, bits)
        member = np.zeros((n_cells, len(centre)), dtype=bool)
        for c, ids in enumerate(lists):
            member[c, ids - first] = True
        # a basis function is non-zero at x iff |x_d - c_d| < s_d in every dimension (LCBasisFunction.cpp:283-288, 384-409)
        nonzero = (np.abs(x[:, None, :] - centre[None, :, :]) < support[None, :, :]).all(axis=2)
        missing = nonzero & ~member[cell]
        assert not missing.any(), f"leaf {leaf}: {int(missing.sum())} non-zero terms missing from their cell's list"
        checked += len(x)
    assert checked > 0 and total_listed < total_merged  # the cells do shorten the lists


@pytest.mark.parametrize("name", CASES)
def test_cell_polynomials_equal_the_oracle(golden, name):
    """Polynomial form of the evaluation cells (csrc/loco_cellpoly.h, SURVEY 8f.4 per cell): on a cell that no knot crosses,
    the F^N coefficients computed by the code the device kernel shares must reproduce the reference restatement's values
    at random points of the cell within the 1e-12 bar."""
    import protoschema
    from conftest import assert_values_close
    from oracle import oracleapi
    pb, _ = golden(name)
    m = capi.Model.from_proto(pb, device=capi.HOST_ONLY)
    orc = oracleapi.Oracle(protoschema.read_binary(pb))
    cubic = m.info["basis_type"] == capi.CUBIC
    F = 4 if cubic else 2
    n_poly = m.info["n_poly_cells"]
    if name == "cub2d_testdebug":
        assert n_poly > 0.5 * m.info["n_eval_cells"], (n_poly, m.info["n_eval_cells"])
    assert n_poly > 0
    dom = m.domain.reshape(m.N, 2)
    rng = np.random.default_rng(41)
    leaves = range(m.L) if m.L <= 60 else rng.choice(m.L, 60, replace=False)
    pts, vals, seen = [], [], 0
    for leaf in leaves:
        bits, first_cell, n_cells = m.eval_cells(int(leaf))
        for c in range(n_cells):
            pc = m.eval_cell_poly(first_cell + c)
            if pc is None:
                continue
            seen += 1
            if seen % 3:  # a third of the polynomial cells is plenty
                continue
            geom, coef = pc
            assert len(coef) == F ** m.N
            t = rng.uniform(-1, 1, (8, m.N))
            t[0] = 1.0   # corners of the cell: the pieces must still hold on the closed cell
            t[1] = -1.0
            x = geom[:, 0] + t / geom[:, 1]
            powers = np.stack([t ** k for k in range(F)], axis=-1)  # [8, N, F]
            v = np.zeros(len(t))
            for j in range(len(coef)):
                term, r = np.full(len(t), coef[j]), j
                for d in range(m.N):
                    term = term * powers[:, d, r % F]
                    r //= F
                v += term
            pts.append(x)
            vals.append(v)
    x, v = np.concatenate(pts), np.concatenate(vals)
    q = dom[:, 0] + (dom[:, 1] - dom[:, 0]) * (x + 1.0) * 0.5
    o_out, _, o_status, scale = orc.eval(q)
    assert (o_status == 0).all() and len(q) > 100
    assert_values_close(v, o_out, scale)
