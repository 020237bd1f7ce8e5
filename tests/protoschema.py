"""The reference's on-disk schema (LCFunction/PrecomputedParametricShape.proto, proto2, package
OptCAD.proto) rebuilt as a dynamic google.protobuf descriptor -- an encoder/decoder that is fully
independent of the hand-written C++ codec in locointerpolate_b200/csrc/loco_proto.cpp.

Only the messages the format actually uses are declared (field numbers and names as in the .proto;
unused TetMesh members are left out -- unknown fields are skipped by both codecs).
"""
from __future__ import annotations

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory, text_format

from locointerpolate_b200.capi import GraphArrays

_F = descriptor_pb2.FieldDescriptorProto
_OPT, _REP = _F.LABEL_OPTIONAL, _F.LABEL_REPEATED


def _build_pool():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "PrecomputedParametricShape.proto"
    fd.package = "OptCAD.proto"
    fd.syntax = "proto2"

    def msg(name, fields):
        m = fd.message_type.add()
        m.name = name
        for num, fname, label, ftype, tname in fields:
            f = m.field.add()
            f.name, f.number, f.label, f.type = fname, num, label, ftype
            if tname:
                f.type_name = ".OptCAD.proto." + tname

    D, I, M = _F.TYPE_DOUBLE, _F.TYPE_INT32, _F.TYPE_MESSAGE
    msg("PrecomputedParametricShape", [(1, "root", _OPT, M, "AdaptiveGridCell"), (2, "samples", _REP, M, "PrecomputedSample"),
                                       (3, "bootstrapNSamples", _OPT, I, None), (4, "borderCells", _REP, M, "AdaptiveGridCell"),
                                       (5, "range", _REP, D, None)])
    msg("PrecomputedSample", [(1, "id", _OPT, I, None), (2, "center", _REP, D, None), (3, "shapeInfo", _OPT, M, "ShapeInfo"),
                              (4, "basisFuntions", _REP, M, "BasisFunction")])
    msg("FunctionTestShapeInfo", [(1, "val", _OPT, D, None)])
    msg("Vector3d", [(1, "x", _OPT, D, None), (2, "y", _OPT, D, None), (3, "z", _OPT, D, None)])
    msg("TetMesh", [(1, "vertices", _REP, M, "Vector3d")])
    msg("ShapeInfo", [(1, "functionShapeInfo", _OPT, M, "FunctionTestShapeInfo"), (2, "tetMesh", _OPT, M, "TetMesh")])
    msg("BasisFunction", [(1, "linearBspline", _OPT, M, "LinearBSpline"), (2, "cubicBspline", _OPT, M, "CubicBSpline")])
    for n in ("LinearBSpline", "CubicBSpline"):
        msg(n, [(1, "center", _REP, D, None), (2, "support", _REP, D, None), (3, "weight", _OPT, D, None)])
    msg("AdaptiveGridCell", [(1, "ranges", _REP, D, None), (2, "leaf", _OPT, M, "AdaptiveGridLeaf"),
                             (3, "interNode", _OPT, M, "AdaptiveGridInterNode")])
    msg("AdaptiveGridLeaf", [(1, "centerShapeInfo", _OPT, M, "ShapeInfo"), (2, "homeomorphicSamples", _REP, M, "HomeomorphicSample")])
    msg("HomeomorphicSample", [(1, "precomputedSampleID", _OPT, I, None), (2, "shapeInfo", _OPT, M, "ShapeInfo")])
    msg("AdaptiveGridInterNode", [(1, "child1", _OPT, M, "AdaptiveGridCell"), (2, "child2", _OPT, M, "AdaptiveGridCell"),
                                  (3, "splidDirection", _OPT, I, None), (4, "splitVal", _OPT, D, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return pool


_pool = _build_pool()
Shape = message_factory.GetMessageClass(_pool.FindMessageTypeByName("OptCAD.proto.PrecomputedParametricShape"))


def _put_value(si, v):
    v = np.atleast_1d(v)
    if len(v) == 1:
        si.functionShapeInfo.val = float(v[0])
    else:
        for j in range(0, len(v), 3):
            p = si.tetMesh.vertices.add()
            p.x, p.y, p.z = float(v[j]), float(v[j + 1]), float(v[j + 2])


def _get_value(si):
    if si.HasField("functionShapeInfo"):
        return [si.functionShapeInfo.val]
    out = []
    for p in si.tetMesh.vertices:
        out += [p.x, p.y, p.z]
    return out


def graph_to_message(g: GraphArrays):
    """What convertToProto (LCProtoConverter.cpp:141-168) would write for this graph."""
    g.normalise()
    msg = Shape()
    for s in range(len(g.sample_center)):
        ps = msg.samples.add()
        ps.id = s
        ps.center.extend(g.sample_center[s].tolist())
        _put_value(ps.shapeInfo, g.sample_value[s])
        for b in range(g.bf_offset[s], g.bf_offset[s + 1]):
            bf = ps.basisFuntions.add()
            tgt = bf.cubicBspline if g.basis == 1 else bf.linearBspline
            tgt.center.extend(g.bf_center[b].tolist())
            tgt.support.extend(g.bf_support[b].tolist())
            tgt.weight = float(g.bf_weight[b])

    def leaf_part(cell, center_value, off, sample, copies, i):
        _put_value(cell.leaf.centerShapeInfo, center_value[i])
        for e in range(off[i], off[i + 1]):
            h = cell.leaf.homeomorphicSamples.add()
            h.precomputedSampleID = int(sample[e])
            _put_value(h.shapeInfo, copies[e] if copies is not None else g.sample_value[sample[e]])

    def put_cell(cell, i):
        cell.ranges.extend(g.cell_ranges[i].tolist())
        if g.cell_split_dir[i] < 0:
            leaf_part(cell, g.cell_center_value, g.cell_sample_offset, g.cell_sample, g.cell_sample_value, i)
        else:
            put_cell(cell.interNode.child1, i + 1)
            put_cell(cell.interNode.child2, int(g.cell_child2[i]))
            cell.interNode.splidDirection = int(g.cell_split_dir[i])
            cell.interNode.splitVal = float(g.cell_split_val[i])

    for b in range(len(g.border_ranges)):
        bc = msg.borderCells.add()
        bc.ranges.extend(g.border_ranges[b].tolist())
        leaf_part(bc, g.border_center_value, g.border_sample_offset, g.border_sample, g.border_sample_value, b)
    put_cell(msg.root, 0)
    msg.range.extend(g.domain.tolist())
    return msg


def message_to_graph(msg) -> GraphArrays:
    """convertToCplusplus (LCProtoConverter.cpp:356-403) into plain arrays."""
    N = len(msg.root.ranges) // 2
    id_to_index = {}
    centers, values, bf_off, bf_c, bf_s, bf_w = [], [], [0], [], [], []
    basis = None
    for i, ps in enumerate(msg.samples):
        id_to_index[ps.id] = i
        centers.append(list(ps.center))
        values.append(_get_value(ps.shapeInfo))
        for bf in ps.basisFuntions:
            cub = bf.HasField("cubicBspline")
            src = bf.cubicBspline if cub else bf.linearBspline
            if basis is None:
                basis = 1 if cub else 0
            bf_c.append(list(src.center)); bf_s.append(list(src.support)); bf_w.append(src.weight)
        bf_off.append(len(bf_w))
    D = len(values[0])
    ranges, sdir, sval, cval, off, smp, cpy = [], [], [], [], [0], [], []

    def walk(cell):
        ranges.append(list(cell.ranges))
        if cell.HasField("leaf"):
            sdir.append(-1); sval.append(0.0)
            cval.append(_get_value(cell.leaf.centerShapeInfo))
            for h in cell.leaf.homeomorphicSamples:
                smp.append(id_to_index[h.precomputedSampleID]); cpy.append(_get_value(h.shapeInfo))
            off.append(len(smp))
        else:
            sdir.append(cell.interNode.splidDirection); sval.append(cell.interNode.splitVal)
            cval.append([0.0] * D); off.append(len(smp))
            walk(cell.interNode.child1); walk(cell.interNode.child2)

    walk(msg.root)
    tree = (ranges, sdir, sval, cval, off, smp, cpy)
    ranges, sdir, sval, cval, off, smp, cpy = [], [], [], [], [0], [], []
    for bc in msg.borderCells:
        walk(bc)
    dom = list(msg.range) if len(msg.range) else tree[0][0]
    f = lambda a, w: np.asarray(a, dtype=np.float64).reshape(-1, w)
    return GraphArrays(
        N=N, D=D, basis=basis or 0, domain=np.asarray(dom), sample_center=f(centers, N), sample_value=f(values, D),
        bf_offset=np.asarray(bf_off, dtype=np.int64), bf_center=f(bf_c, N), bf_support=f(bf_s, N), bf_weight=np.asarray(bf_w, dtype=np.float64),
        cell_ranges=f(tree[0], 2 * N), cell_split_dir=np.asarray(tree[1], dtype=np.int32), cell_split_val=np.asarray(tree[2], dtype=np.float64),
        cell_center_value=f(tree[3], D), cell_sample_offset=np.asarray(tree[4], dtype=np.int64), cell_sample=np.asarray(tree[5], dtype=np.int32),
        cell_sample_value=f(tree[6], D) if tree[6] else None,
        border_ranges=f(ranges, 2 * N), border_center_value=f(cval, D) if cval else None,
        border_sample_offset=np.asarray(off, dtype=np.int64), border_sample=np.asarray(smp, dtype=np.int32),
        border_sample_value=f(cpy, D) if cpy else None).normalise()


def write_binary(g: GraphArrays, path: str):
    with open(path, "wb") as fh:
        fh.write(graph_to_message(g).SerializePartialToString())  # SerializePartialToOstream, LCShapeInfoProtoConverter.h:38


def write_text(g: GraphArrays, path: str):
    with open(path, "w") as fh:
        fh.write(text_format.MessageToString(graph_to_message(g)))  # TextFormat::PrintToString, :20


def read_binary(path: str) -> GraphArrays:
    m = Shape()
    with open(path, "rb") as fh:
        m.ParseFromString(fh.read())
    return message_to_graph(m)


def read_text(path: str) -> GraphArrays:
    m = Shape()
    with open(path) as fh:
        text_format.Parse(fh.read(), m)
    return message_to_graph(m)
