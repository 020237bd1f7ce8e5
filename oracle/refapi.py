"""ctypes front-end for oracle/_ref/liblocoref.so -- TEST INFRASTRUCTURE ONLY.

liblocoref.so is the reference's own C++ (LCUtils, LCFunction minus protobuf, LCRealFunction,
LCAdaptiveSampling minus LCProtoConverter) compiled unmodified by oracle/Makefile, plus the
extern "C" adapter oracle/ref_adapter.cpp.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "liblocoref.so")

_lib = None


def available() -> bool:
    return os.path.exists(REF_SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(REF_SO)
        vp, i, d, l = C.c_void_p, C.c_int, C.c_double, C.c_long
        P = C.POINTER
        L.ref_grid_new.restype = vp
        L.ref_grid_new.argtypes = [i, i, i, i, d, i, i, i]
        L.ref_grid_num_leaves.argtypes = [vp]
        L.ref_grid_refine_random.argtypes = [vp, C.c_uint, i, i]
        L.ref_grid_refine.argtypes = [vp, i, i]
        L.ref_grid_optimal_direction.argtypes = [vp, i]
        L.ref_grid_true_values.argtypes = [vp, vp, l, vp]
        L.ref_fn_from_grid.restype = vp
        L.ref_fn_from_grid.argtypes = [vp]
        L.ref_fn_from_graph.restype = vp
        L.ref_fn_from_graph.argtypes = [vp]
        L.ref_fn_reload.restype = vp
        L.ref_fn_reload.argtypes = [vp]
        L.ref_fn_sizes.argtypes = [vp, vp]
        L.ref_fn_ranges.argtypes = [vp, vp]
        L.ref_fn_dump_samples.argtypes = [vp] * 9
        L.ref_fn_dump_cells.argtypes = [vp] * 9
        L.ref_fn_dump_border.argtypes = [vp] * 6
        L.ref_fn_eval.argtypes = [vp, vp, l, vp, vp, vp]
        L.ref_fn_eval_deriv.argtypes = [vp, vp, l, i, vp]
        L.ref_fn_eval_mesh_sum.restype = d
        L.ref_fn_eval_mesh_sum.argtypes = [vp, vp, l, vp, l]
        L.ref_fn_eval_mesh_fork.argtypes = [vp, vp, l, vp, l, i]
        L.ref_fn_eval_sum.restype = d
        L.ref_fn_eval_sum.argtypes = [vp, vp, l]
        L.ref_fn_eval_fork.argtypes = [vp, vp, l, i]
        L.ref_fn_weights.argtypes = [vp, i, vp, vp, vp, i]
        L.ref_fn_support.argtypes = [vp, i, vp, i]
        L.ref_fn_num_neighbors.argtypes = [vp, i]
        L.ref_fn_center_error.restype = d
        L.ref_fn_center_error.argtypes = [vp, i]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class Graph:
    """Plain-array image of an LCSampledFunction object graph (what a proto file carries)."""
    N: int
    basis: int                    # 0 linear / 1 cubic
    ranges: np.ndarray            # [2N] domain (min,max per parameter)
    centers: np.ndarray           # [S,N]
    values: np.ndarray            # [S]
    value_group: np.ndarray       # [S] lowest sample index sharing the same value object
    bf_off: np.ndarray            # [S+1]
    bf_center: np.ndarray         # [B,N]
    bf_support: np.ndarray        # [B,N]
    bf_weight: np.ndarray         # [B]
    bf_type: np.ndarray           # [B]
    cell_ranges: np.ndarray       # [C,2N]   tree cells, pre-order
    cell_is_leaf: np.ndarray      # [C]
    cell_split_dir: np.ndarray    # [C]
    cell_split_val: np.ndarray    # [C]
    cell_center_val: np.ndarray   # [C]
    cell_ls_off: np.ndarray       # [C+1]
    cell_ls_sample: np.ndarray    # [E]
    cell_ls_value: np.ndarray     # [E]
    border_ranges: np.ndarray     # [Bc,2N]
    border_center_val: np.ndarray # [Bc]
    border_ls_off: np.ndarray     # [Bc+1]
    border_ls_sample: np.ndarray
    border_ls_value: np.ndarray


class RefGrid:
    """LCAdaptiveGrid built by the reference (LCAdaptiveGrid.cpp:16-39)."""

    def __init__(self, N, basis, linear_fn=False, max_depth=0, threshold=0.5, bootstrap=0,
                 n_error_samples=0, eval_corners=True):
        self.N, self.basis = N, basis
        self.h = lib().ref_grid_new(N, basis, int(linear_fn), max_depth, float(threshold), bootstrap,
                                    n_error_samples, int(eval_corners))

    def num_leaves(self):
        return lib().ref_grid_num_leaves(self.h)

    def refine_random(self, seed=1, n=12, max_level=5):
        r = lib().ref_grid_refine_random(self.h, seed, n, max_level)
        if r < 0:
            raise RuntimeError(f"reference refineCell failed after {-1 - r} refinements")
        return r

    def refine(self, leaf, direction):
        r = lib().ref_grid_refine(self.h, leaf, direction)
        if r != 0:
            raise RuntimeError(f"reference refineCell failed ({r})")

    def optimal_direction(self, leaf):
        """LCAdaptiveGridCell::getOptimalSplitDirection of the leaf-th leaf (-1: the reference reports an error)"""
        return lib().ref_grid_optimal_direction(self.h, int(leaf))

    def true_values(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.N)
        out = np.empty(len(q))
        lib().ref_grid_true_values(self.h, _p(q), len(q), _p(out))
        return out

    def sampled_function(self):
        return RefFn(lib().ref_fn_from_grid(self.h))


class RefFn:
    """LCSampledFunction of the reference (LCSampledFunction.cpp)."""

    def __init__(self, h):
        self.h = h
        s = np.zeros(9, dtype=np.int64)
        lib().ref_fn_sizes(h, _p(s))
        (self.N, self.S, self.C, self.L, self.Bc, self.B, self.E, self.Eb, self.basis) = (int(x) for x in s)

    @classmethod
    def from_graph(cls, g):
        """Reference objects built from a GraphArrays through the loader's constructor sequence."""
        s = g.struct()
        h = lib().ref_fn_from_graph(C.byref(s))
        if not h:
            raise RuntimeError("the reference only has scalar (LCRealFunctionValue) values")
        return cls(h)

    def reload(self):
        """Rebuild through the loader's code path (LCProtoConverter.cpp:356-403 without protobuf)."""
        return RefFn(lib().ref_fn_reload(self.h))

    def eval(self, q, want_leaf=True):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.N)
        Q = len(q)
        out = np.empty(Q)
        leaf = np.empty(Q, dtype=np.int32)
        status = np.empty(Q, dtype=np.uint8)
        lib().ref_fn_eval(self.h, _p(q), Q, _p(out), _p(leaf) if want_leaf else None,
                          _p(status) if want_leaf else None)
        return out, leaf, status

    def eval_deriv(self, q, direction):
        """LCSampledFunction::evalDeriv along cube direction `direction` (NaN where the reference returned an error)."""
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.N)
        out = np.empty(len(q))
        lib().ref_fn_eval_deriv(self.h, _p(q), len(q), direction, _p(out))
        return out

    def eval_mesh_fork(self, q, table, procs):
        """mesh-valued CPU baseline: reference weights + `val += w * other` over the D columns of table [S, D]"""
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.N)
        table = np.ascontiguousarray(table, dtype=np.float64)
        assert table.shape[0] == self.S
        rc = lib().ref_fn_eval_mesh_fork(self.h, _p(q), len(q), _p(table), table.shape[1], procs)
        if rc != 0:
            raise RuntimeError(f"fork-parallel mesh evaluation failed ({rc})")

    def eval_sum(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.N)
        return lib().ref_fn_eval_sum(self.h, _p(q), len(q))

    def eval_fork(self, q, procs):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.N)
        rc = lib().ref_fn_eval_fork(self.h, _p(q), len(q), procs)
        if rc != 0:
            raise RuntimeError(f"fork-parallel reference evaluation failed ({rc})")

    def support(self, leaf, cap=1 << 20):
        ids = np.empty(cap, dtype=np.int32)
        k = lib().ref_fn_support(self.h, leaf, _p(ids), cap)
        if k < 0:
            raise RuntimeError(f"ref_fn_support failed ({k})")
        return ids[:k].copy()

    def weights(self, leaf, cube_point, cap=1 << 20):
        x = np.ascontiguousarray(cube_point, dtype=np.float64)
        ids = np.empty(cap, dtype=np.int32)
        w = np.empty(cap)
        k = lib().ref_fn_weights(self.h, leaf, _p(x), _p(ids), _p(w), cap)
        if k < 0:
            raise RuntimeError(f"ref_fn_weights failed ({k})")
        return ids[:k].copy(), w[:k].copy()

    def num_neighbors(self, leaf):
        return lib().ref_fn_num_neighbors(self.h, leaf)

    def center_error(self, leaf):
        return lib().ref_fn_center_error(self.h, leaf)

    def graph(self) -> Graph:
        N, S, Cn, Bc, B, E, Eb = self.N, self.S, self.C, self.Bc, self.B, self.E, self.Eb
        ranges = np.empty(2 * N)
        lib().ref_fn_ranges(self.h, _p(ranges))
        centers = np.empty((S, N)); values = np.empty(S); vg = np.empty(S, dtype=np.int32)
        bf_off = np.empty(S + 1, dtype=np.int32)
        bf_c = np.empty((B, N)); bf_s = np.empty((B, N)); bf_w = np.empty(B); bf_t = np.empty(B, dtype=np.int32)
        lib().ref_fn_dump_samples(self.h, _p(centers), _p(values), _p(vg), _p(bf_off), _p(bf_c), _p(bf_s), _p(bf_w), _p(bf_t))
        cr = np.empty((Cn, 2 * N)); il = np.empty(Cn, dtype=np.int32); sd = np.empty(Cn, dtype=np.int32)
        sv = np.empty(Cn); cv = np.empty(Cn); lo = np.empty(Cn + 1, dtype=np.int32)
        ls = np.empty(max(E, 1), dtype=np.int32); lv = np.empty(max(E, 1))
        lib().ref_fn_dump_cells(self.h, _p(cr), _p(il), _p(sd), _p(sv), _p(cv), _p(lo), _p(ls), _p(lv))
        br = np.empty((Bc, 2 * N)); bcv = np.empty(Bc); bo = np.empty(Bc + 1, dtype=np.int32)
        bs = np.empty(max(Eb, 1), dtype=np.int32); bv = np.empty(max(Eb, 1))
        lib().ref_fn_dump_border(self.h, _p(br), _p(bcv), _p(bo), _p(bs), _p(bv))
        return Graph(N, self.basis, ranges, centers, values, vg, bf_off, bf_c, bf_s, bf_w, bf_t,
                     cr, il, sd, sv, cv, lo, ls[:E], lv[:E], br, bcv, bo, bs[:Eb], bv[:Eb])
