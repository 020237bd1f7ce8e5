"""Host-side logic of the product, testable without a GPU: the C-ABI library loads and exports what
include/loco_b200.h declares, the hand-written proto codec agrees with google.protobuf, the flattener
reproduces the reference's support sets, the bootstrap generator reproduces the reference's builder, and
evaluation without a CUDA device fails loudly (there is no CPU path)."""
import ctypes
import os
import re

import numpy as np
import pytest

import protoschema
from conftest import GOLDEN_CASES, ROOT
from locointerpolate_b200 import capi
from oracle import oracleapi, refapi


def test_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "loco_b200.h")).read()
    declared = set(re.findall(r"\b(loco_[a-z0-9_]+)\s*\(", header))
    declared -= {"loco_value_fn"}
    assert declared == set(capi.SIGNATURES), (declared ^ set(capi.SIGNATURES))
    L = ctypes.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_headers_compile_as_c(tmp_path):
    """the boundary is a C ABI: both headers are valid C99 (plain pointers and sizes, no C++ in the signatures)"""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "loco_b200.h"\n#include "loco_grid.h"\nint main(void) { loco_model_t *m = 0; loco_grid_t *g = 0; loco_graph_t gr; (void)m; (void)g; (void)gr; return 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "abi.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_grid_abi_exports_every_declared_symbol():
    """the host-only companion library (include/loco_grid.h: the builder's grid state and refineCell) likewise"""
    from locointerpolate_b200 import build as product_build, grid
    header = open(os.path.join(ROOT, "include", "loco_grid.h")).read()
    declared = set(re.findall(r"\b(loco_grid_[a-z0-9_]+)\s*\(", header))
    assert declared == set(grid.SYMBOLS), (declared ^ set(grid.SYMBOLS))
    L = ctypes.CDLL(product_build.build_grid())
    for name in declared:
        assert hasattr(L, name), name
    # constructor errors come back through loco_grid_last_error(NULL)
    with pytest.raises(capi.LocoError, match="number of levels must be positive"):
        grid.HostGrid.bootstrap(2, capi.LINEAR, -1, [0, 1, 0, 1], lambda p: p[:, 0])
    g = grid.HostGrid.bootstrap(2, capi.CUBIC, 1, [0, 1, 0, 1], lambda p: p[:, 0])
    with pytest.raises(capi.LocoError, match="split direction out of range"):
        g.refine(0, 2)
    with pytest.raises(capi.LocoError, match="leaf index out of range"):
        g.refine(4, 0)
    with pytest.raises(capi.LocoError, match="descending"):
        g.refine_pass([0, 1], [0, 0])
    g.refine_pass([3, 1], [0, 1])
    assert g.num_leaves() == 6


def test_no_cpu_path():
    """Without a usable device, construction fails; a host-only handle refuses to evaluate."""
    pb = os.path.join(ROOT, "tests", "golden", "c1_lin2d_boot2.pb")
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        with pytest.raises(capi.LocoError) as e:
            capi.Model.from_proto(pb, device=0)
        assert e.value.code == 3 and "no CPU path" in e.value.msg
    m = capi.Model.from_proto(pb, device=capi.HOST_ONLY)
    with pytest.raises(capi.LocoError) as e:
        m.eval(np.zeros((4, 2)))
    assert e.value.code == 3 and "no CPU path" in e.value.msg
    with pytest.raises(capi.LocoError):
        m.center_error(0.1)


def test_product_does_not_link_the_oracle():
    """The shipped library must not reference anything under oracle/ (a product path through the oracle
    would void every parity claim)."""
    import subprocess
    syms = subprocess.run(["nm", "-D", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in syms and "ref_fn_" not in syms
    for root, _, files in os.walk(os.path.join(ROOT, "locointerpolate_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(root, f)).read()
                assert "oracleapi" not in src and "refapi" not in src and "loco_oracle" not in src, f


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_proto_reader_and_flattener(golden, name, tmp_path):
    pb, z = golden(name)
    m = capi.Model.from_proto(pb, device=capi.HOST_ONLY)
    g = protoschema.read_binary(pb)
    assert m.N == g.N and m.D == 1 and m.info["basis_type"] == g.basis
    assert m.info["n_samples"] == len(g.sample_center) and m.info["n_cells"] == len(g.cell_split_dir)
    assert m.info["n_border_cells"] == len(g.border_ranges) and m.info["n_basis_functions"] == g.bf_offset[-1]
    assert m.L == len(z["sup_off"]) - 1
    for l in range(m.L):
        assert np.array_equal(m.support_set(l), z["sup_ids"][z["sup_off"][l]:z["sup_off"][l + 1]]), f"support set of leaf {l}"
        assert sum(m.leaf_neighbors(l)) == z["n_neighbors"][l]
    # writer: binary output parses with google.protobuf into the same graph, and is byte-identical to
    # what google.protobuf serialises for it (field order, unpacked proto2 doubles)
    out = str(tmp_path / "rt.pb")
    m.save_proto(out)
    g2 = protoschema.read_binary(out)
    for f in ("domain", "sample_center", "sample_value", "bf_offset", "bf_center", "bf_support", "bf_weight", "cell_ranges",
              "cell_split_dir", "cell_split_val", "cell_sample_offset", "cell_sample", "border_ranges", "border_sample_offset", "border_sample"):
        assert np.array_equal(getattr(g, f), getattr(g2, f)), f
    assert open(out, "rb").read() == protoschema.graph_to_message(g2).SerializePartialToString()
    # text format both ways
    txt = str(tmp_path / "rt.txt")
    m.save_proto(txt, ascii=True)
    g3 = protoschema.read_text(txt)
    assert np.array_equal(g3.bf_weight, g.bf_weight) and np.array_equal(g3.cell_ranges, g.cell_ranges)
    txt2 = str(tmp_path / "py.txt")
    protoschema.write_text(g, txt2)
    m2 = capi.Model.from_proto(txt2, ascii=True, device=capi.HOST_ONLY)
    assert m2.info["n_terms"] == m.info["n_terms"] and m2.info["n_support_entries"] == m.info["n_support_entries"]


def test_proto_errors_use_reference_messages(tmp_path):
    with pytest.raises(capi.LocoError) as e:
        capi.Model.from_proto(str(tmp_path / "missing.pb"), device=capi.HOST_ONLY)
    assert e.value.code == 2 and "could not open file" in e.value.msg
    bad = tmp_path / "bad.pb"
    bad.write_bytes(b"\x0a\xff\xff\xff\x01garbage")
    with pytest.raises(capi.LocoError) as e:
        capi.Model.from_proto(str(bad), device=capi.HOST_ONLY)
    assert "could not parse the file" in e.value.msg
    # a leaf that names an unknown sample id (LCProtoConverter.cpp:326-330)
    g = protoschema.read_binary(os.path.join(ROOT, "tests", "golden", "c1_lin2d_boot2.pb"))
    msg = protoschema.graph_to_message(g)
    cell = msg.root
    while cell.HasField("interNode"):
        cell = cell.interNode.child1
    cell.leaf.homeomorphicSamples[0].precomputedSampleID = 9999
    p = tmp_path / "badid.pb"
    p.write_bytes(msg.SerializePartialToString())
    with pytest.raises(capi.LocoError) as e:
        capi.Model.from_proto(str(p), device=capi.HOST_ONLY)
    assert "could not map the sample from id" in e.value.msg
    # sample without id (LCProtoConverter.cpp:282-285)
    msg = protoschema.graph_to_message(g)
    msg.samples[3].ClearField("id")
    p.write_bytes(msg.SerializePartialToString())
    with pytest.raises(capi.LocoError) as e:
        capi.Model.from_proto(str(p), device=capi.HOST_ONLY)
    assert "sample does not have id" in e.value.msg


def test_corrupted_proto_files_never_crash_the_loader(tmp_path):
    """Truncated / bit-flipped / spliced files must either load (the damage hit a value) or fail with an LCError-style
    message — the loader validates every length and index it reads (loco_proto.cpp, validate_graph)."""
    import random
    src = open(os.path.join(ROOT, "tests", "golden", "cub2d_boot1_adapt.pb"), "rb").read()
    rnd = random.Random(11)
    outcomes = {"ok": 0, "err": 0}
    for it in range(120):
        b = bytearray(src)
        mode = rnd.choice(["trunc", "flip", "flip", "insert", "zero"])
        if mode == "trunc":
            b = b[:rnd.randrange(0, len(b))]
        elif mode == "flip":
            for _ in range(rnd.randrange(1, 6)):
                b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
        elif mode == "insert":
            i = rnd.randrange(len(b))
            b[i:i] = bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 9)))
        else:
            i, n = rnd.randrange(len(b)), rnd.randrange(1, 64)
            b[i:i + n] = bytes(n)
        path = str(tmp_path / "damaged.pb")
        open(path, "wb").write(bytes(b))
        try:
            m = capi.Model.from_proto(path, device=capi.HOST_ONLY)
            assert m.L >= 1
            outcomes["ok"] += 1
        except capi.LocoError as e:
            assert str(e)
            outcomes["err"] += 1
    assert outcomes["err"] > 20 and outcomes["ok"] + outcomes["err"] == 120


def test_packed_doubles_are_accepted(tmp_path):
    """proto3-style writers pack repeated doubles; the reader must take both encodings."""
    g = protoschema.read_binary(os.path.join(ROOT, "tests", "golden", "cub2d_boot2.pb"))
    raw = protoschema.graph_to_message(g).SerializePartialToString()

    def varint(v):
        o = bytearray()
        while v >= 0x80:
            o.append((v & 0x7f) | 0x80); v >>= 7
        o.append(v)
        return bytes(o)
    dom = np.asarray(g.domain, dtype="<f8").tobytes()
    unpacked = b"".join(b"\x29" + dom[i:i + 8] for i in range(0, len(dom), 8))  # field 5, wire type 1
    assert raw.endswith(unpacked)
    packed = raw[:-len(unpacked)] + b"\x2a" + varint(len(dom)) + dom               # field 5, wire type 2
    p = tmp_path / "packed.pb"
    p.write_bytes(packed)
    m = capi.Model.from_proto(str(p), device=capi.HOST_ONLY)
    assert m.info["n_leaves"] == 16


@pytest.mark.skipif(not refapi.available(), reason="oracle/_ref/liblocoref.so not built")
@pytest.mark.parametrize("N,basis,level,corners", [(1, 0, 3, True), (2, 0, 2, True), (2, 1, 2, True), (2, 1, 0, True), (3, 1, 1, False),
                                                   (3, 1, 2, True), (4, 0, 1, True), (4, 1, 1, False)])
def test_bootstrap_generator_matches_reference_builder(N, basis, level, corners, tmp_path):
    """loco_model_bootstrap vs LCAdaptiveGrid::bootstrapSample: same samples (order, coordinates bit for
    bit), basis functions, tree, leaf / border-cell sample sets, value sharing of border samples."""
    g = refapi.RefGrid(N, basis, False, 0, 0.1, level, 0, corners)
    fn = g.sampled_function()
    rg = oracleapi.graph_from_ref(fn.graph())
    m = capi.Model.bootstrap(N, basis, level, 1, corners, fn=lambda p: g.true_values(p[None, :])[0], device=capi.HOST_ONLY)
    out = str(tmp_path / "boot.pb")
    m.save_proto(out)
    mine = protoschema.read_binary(out)
    for f in ("domain", "sample_center", "sample_value", "bf_offset", "bf_center", "bf_support", "bf_weight", "cell_ranges",
              "cell_split_dir", "cell_split_val", "cell_center_value", "border_ranges", "border_center_value"):
        a, b = getattr(rg, f), getattr(mine, f)
        if f in ("cell_center_value",):  # only leaves carry a centre value
            leafs = rg.cell_split_dir < 0
            a, b = a[leafs], b[leafs]
        assert np.array_equal(a, b), f

    def sets(off, smp):
        return [sorted(smp[off[i]:off[i + 1]].tolist()) for i in range(len(off) - 1)]
    assert sets(rg.cell_sample_offset, rg.cell_sample) == sets(mine.cell_sample_offset, mine.cell_sample)
    assert sets(rg.border_sample_offset, rg.border_sample) == sets(mine.border_sample_offset, mine.border_sample)
    assert m.info["n_value_rows"] - m.info["n_leaves"] - m.info["n_border_cells"] == len(np.unique(rg.sample_value_group))
    for l in range(fn.L):
        assert np.array_equal(m.support_set(l), fn.support(l))
        assert sum(m.leaf_neighbors(l)) == fn.num_neighbors(l)


@pytest.mark.skipif(not refapi.available(), reason="oracle/_ref/liblocoref.so not built")
@pytest.mark.parametrize("N,basis,boot,seed,n,max_level", [(2, 1, 1, 3, 14, 6), (2, 0, 1, 5, 20, 7), (3, 0, 0, 2, 10, 5), (3, 1, 0, 4, 6, 4),
                                                         (1, 1, 1, 6, 8, 6)])
def test_flattener_on_randomly_refined_live_grids(N, basis, boot, seed, n, max_level):
    """Host flattener against the LIVE reference on grids refined at random leaves in random directions (the recipe of
    TestExamples.cpp:225-240 with other seeds): leaf count, influencing-sample sets (bit-exact), neighbour counts, and the
    evaluation-cell lists still cover every non-zero term (introspection)."""
    grid = refapi.RefGrid(N, basis, False, 0, 0.1, boot)
    done = grid.refine_random(seed, n, max_level)
    assert done > 0
    fn = grid.sampled_function()
    m = capi.Model.from_graph(oracleapi.graph_from_ref(fn.graph()), device=capi.HOST_ONLY)
    assert m.L == fn.L == grid.num_leaves()
    for l in range(m.L):
        assert np.array_equal(m.support_set(l), fn.support(l)), f"support set of leaf {l}"
        assert sum(m.leaf_neighbors(l)) == fn.num_neighbors(l)
    if m.info["n_eval_cells"]:
        rng = np.random.default_rng(seed)
        for l in range(m.L):
            bits, first_cell, n_cells = m.eval_cells(l)
            if n_cells == 1:
                continue
            centre, support, first = m.merged_terms(l)
            box = m.leaf_box(l)
            x = box[:, 0] + (box[:, 1] - box[:, 0]) * rng.random((400, N))
            idx, shift = np.zeros(len(x), dtype=np.int64), 0
            for d, b in enumerate(bits):
                if b:
                    g = 1 << int(b)
                    idx |= np.clip(((x[:, d] - box[d, 0]) * g / (box[d, 1] - box[d, 0])).astype(np.int64), 0, g - 1) << shift
                    shift += int(b)
            nonzero = (np.abs(x[:, None, :] - centre[None]) < support[None]).all(axis=2)
            for c in np.unique(idx):
                listed = np.zeros(len(centre), dtype=bool)
                listed[m.eval_cell_terms(first_cell + int(c)) - first] = True
                assert not (nonzero[idx == c] & ~listed).any()
    # polynomial form of the cells no knot crosses, against the reference's own evaluation at random points of the cell
    if m.info["n_poly_cells"]:
        F = 4 if basis == 1 else 2
        dom = m.domain.reshape(N, 2)
        rng = np.random.default_rng(seed + 100)
        pts, vals = [], []
        for cell in rng.choice(m.info["n_eval_cells"], min(300, m.info["n_eval_cells"]), replace=False):
            pc = m.eval_cell_poly(int(cell))
            if pc is None:
                continue
            geom, coef = pc
            t = rng.uniform(-1, 1, (4, N))
            v = np.zeros(len(t))
            for j in range(len(coef)):
                term, r = np.full(len(t), coef[j]), j
                for d in range(N):
                    term = term * t[:, d] ** (r % F)
                    r //= F
                v += term
            pts.append(geom[:, 0] + t / geom[:, 1])
            vals.append(v)
        x, v = np.concatenate(pts), np.concatenate(vals)
        r_out, _, r_status = fn.eval(dom[:, 0] + (dom[:, 1] - dom[:, 0]) * (x + 1.0) * 0.5)
        assert (r_status == 0).all()
        assert np.abs(v - r_out).max() <= 1e-12 * max(1.0, np.abs(r_out).max())


def test_bootstrap_sizes_of_baseline_configs():
    """Tree sizes quoted for the BASELINE configs, built host-only with a virtual value table."""
    m = capi.Model.bootstrap(3, capi.CUBIC, 3, 1, True, device=capi.HOST_ONLY)  # C2 tree at boot 3
    assert (m.info["n_leaves"], m.info["n_samples"], m.info["n_border_cells"], m.info["max_support"]) == (512, 1331, 488, 64)
    m = capi.Model.bootstrap(4, capi.LINEAR, 3, 30, True, device=capi.HOST_ONLY)  # C3 tree
    assert (m.info["n_leaves"], m.info["n_samples"], m.info["max_support"], m.info["depth"]) == (4096, 6561, 16, 12)
    m = capi.Model.bootstrap(6, capi.CUBIC, 1, 3, False, device=capi.HOST_ONLY)  # C4 boot-1 tree
    assert (m.info["n_leaves"], m.info["n_samples"], m.info["n_border_cells"], m.info["max_support"]) == (64, 15625, 4032, 4096)
    assert m.info["n_value_rows"] == 3 ** 6  # border samples share the rows of their clamped in-domain samples
    m = capi.Model.bootstrap(8, capi.LINEAR, 1, 1, True, device=capi.HOST_ONLY)  # C5 tree
    assert (m.info["n_leaves"], m.info["n_samples"], m.info["max_support"]) == (256, 6561, 256)


def test_graph_argument_validation():
    g = protoschema.read_binary(os.path.join(ROOT, "tests", "golden", "c1_lin2d_boot2.pb"))
    g.cell_sample = g.cell_sample.copy()
    g.cell_sample[0] = 10 ** 6
    with pytest.raises(capi.LocoError) as e:
        capi.Model.from_graph(g, device=capi.HOST_ONLY)
    assert "could not map the sample from id" in e.value.msg


@pytest.mark.parametrize("name,cells", [("cub2d_prototest", True), ("lin2d_adapt6", True), ("lin3d_adapt5", True), ("cub3d_boot2", False),
                                        ("c1_lin2d_boot2", False)])
def test_evaluation_cells_are_built_for_adaptive_scalar_trees(golden, name, cells):
    """Flat::ecell_*: leaves with many merged terms are cut into evaluation cells; tensor-product trees have none."""
    pb, _ = golden(name)
    info = capi.Model.from_proto(pb, device=capi.HOST_ONLY).info
    if cells:
        assert info["n_eval_cells"] > info["n_leaves"]
        assert 0 < info["n_eval_terms"] <= 32 * info["n_terms"]
    else:
        assert info["n_eval_cells"] == 0 and info["n_eval_terms"] == 0


def test_mesh_valued_proto_round_trip(golden, tmp_path):
    """Mesh values travel in ShapeInfo.tetMesh.vertices (PrecomputedParametricShape.proto:87-99; the slot the public
    reference no longer reads, LCProtoConverter.cpp:211-263): D = 3 * #vertices doubles per value.  A file written by
    google.protobuf must load, and what the product writes must parse back into the same values."""
    pb, _ = golden("cub2d_boot1_adapt")
    g = protoschema.read_binary(pb)
    V = 5
    rng = np.random.default_rng(4)
    g.D = 3 * V
    g.sample_value = rng.standard_normal((len(g.sample_center), 3 * V))
    g.cell_center_value = rng.standard_normal((len(g.cell_split_dir), 3 * V))
    g.border_center_value = rng.standard_normal((len(g.border_ranges), 3 * V))
    g.cell_sample_value = g.border_sample_value = None
    src = str(tmp_path / "mesh.pb")
    protoschema.write_binary(g, src)
    m = capi.Model.from_proto(src, device=capi.HOST_ONLY)
    assert m.D == 3 * V and m.info["n_samples"] == len(g.sample_center)
    out = str(tmp_path / "mesh_rt.pb")
    m.save_proto(out)
    g2 = protoschema.read_binary(out)
    assert g2.D == 3 * V
    assert np.array_equal(g2.sample_value, g.sample_value)
    leaves = np.asarray(g.cell_split_dir) < 0
    assert np.array_equal(np.asarray(g2.cell_center_value)[leaves], g.cell_center_value[leaves])
    txt = str(tmp_path / "mesh.txt")
    m.save_proto(txt, ascii=True)
    g3 = protoschema.read_text(txt)
    assert np.array_equal(g3.sample_value, g.sample_value)
