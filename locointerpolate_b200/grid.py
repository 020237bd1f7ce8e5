"""ctypes front-end of libloco_grid.so (C ABI include/loco_grid.h): the HOST side of the refinement loop -- the builder's
grid state with bootstrapSample and refineCell's topology surgery (LCAdaptiveGrid.cpp:256-456, 629-914; include/loco_grid.hpp).
Host-only plumbing, no evaluation: the error check of the loop runs on the device through capi.Model
(loco_center_error / loco_error_check*), the grid's state reaches it as a GraphArrays (Model.from_graph / update_from_graph).

    grid = HostGrid.bootstrap(N, capi.CUBIC, level, domain, f)          # f: points [n, N] in ORIGINAL parameters -> values [n] or [n, D]
    log = adaptive_sample(grid, threshold, max_level, device=0)         # the level-synchronous loop of refine.py on this grid
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List

import numpy as np

from . import build as _build
from .capi import VALUE_FN, GraphArrays, LocoError, Model, _GraphStruct

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build_grid())
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.loco_grid_bootstrap.argtypes = [i32, i32, i32, i32, i32, vp, VALUE_FN, vp, C.POINTER(vp)]
        L.loco_grid_from_graph.argtypes = [C.POINTER(_GraphStruct), VALUE_FN, vp, C.POINTER(vp)]
        L.loco_grid_free.argtypes = [vp]
        L.loco_grid_free.restype = None
        L.loco_grid_last_error.argtypes = [vp]
        L.loco_grid_last_error.restype = C.c_char_p
        L.loco_grid_refine_leaf.argtypes = [vp, i32, i32]
        L.loco_grid_refine_leaves.argtypes = [vp, vp, vp, i64]
        L.loco_grid_optimal_split_direction.argtypes = [vp, i32, C.POINTER(i32)]
        L.loco_grid_graph.argtypes = [vp, C.POINTER(_GraphStruct)]
        for n in ("loco_grid_num_leaves", "loco_grid_num_samples", "loco_grid_num_function_evaluations"):
            getattr(L, n).argtypes = [vp]
            getattr(L, n).restype = i64
        L.loco_grid_leaf_neighbors.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
        _lib = L
    return _lib


# the symbols include/loco_grid.h declares (checked by tests/test_host.py)
SYMBOLS = ("loco_grid_bootstrap", "loco_grid_from_graph", "loco_grid_free", "loco_grid_last_error", "loco_grid_refine_leaf",
           "loco_grid_refine_leaves", "loco_grid_optimal_split_direction", "loco_grid_graph", "loco_grid_num_leaves", "loco_grid_num_samples",
           "loco_grid_num_function_evaluations", "loco_grid_leaf_neighbors")


def _arr(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    ct = {np.float64: C.c_double, np.int32: C.c_int32, np.int64: C.c_int64}[dtype]
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()


class HostGrid:
    """loco_grid_t: the builder's grid (cells with neighbour lists and sample maps, samples with their basis functions)."""

    def __init__(self):
        raise TypeError("use HostGrid.bootstrap or HostGrid.from_graph")

    @classmethod
    def _new(cls, N: int, D: int, true_values: Callable[[np.ndarray], np.ndarray]):
        self = cls.__new__(cls)
        self.N, self.D, self.h = N, D, None

        def fn(p, out, _user):
            x = np.array([p[i] for i in range(N)])
            v = np.ravel(np.atleast_1d(true_values(x[None, :])))
            for j in range(D):
                out[j] = float(v[j])
        self._fn = VALUE_FN(fn)   # kept alive with the grid: the library calls it for new samples and leaf centres
        return self

    def _check(self, rc):
        if rc != 0:
            raise LocoError(rc, lib().loco_grid_last_error(self.h).decode())

    @classmethod
    def bootstrap(cls, N, basis, level, domain, true_values, eval_corners=True, D=1) -> "HostGrid":
        """LCAdaptiveGrid(f, basis, params, evalCorners) up to bootstrapSample(level)"""
        self = cls._new(N, D, true_values)
        dom = np.ascontiguousarray(domain, dtype=np.float64).reshape(2 * N)
        h = C.c_void_p()
        self._check(lib().loco_grid_bootstrap(N, basis, level, D, int(eval_corners), dom.ctypes.data_as(C.c_void_p), self._fn, None, C.byref(h)))
        self.h = h
        return self

    @classmethod
    def from_graph(cls, g: GraphArrays, true_values) -> "HostGrid":
        """the builder state of a grid given as plain arrays (live-builder neighbour relations)"""
        self = cls._new(g.N, g.D, true_values)
        s = g.struct()
        h = C.c_void_p()
        self._check(lib().loco_grid_from_graph(C.byref(s), self._fn, None, C.byref(h)))
        self.h = h
        return self

    def __del__(self):
        if getattr(self, "h", None):
            lib().loco_grid_free(self.h)
            self.h = None

    def refine(self, leaf: int, direction: int):
        """LCAdaptiveGrid::refineCell for the leaf-th leaf (getAllLeafCells order) along cube direction `direction`"""
        self._check(lib().loco_grid_refine_leaf(self.h, int(leaf), int(direction)))

    def refine_pass(self, leaves, directions):
        """the requests of one level-synchronous pass (descending leaf order)"""
        a = np.ascontiguousarray(leaves, dtype=np.int32)
        b = np.ascontiguousarray(directions, dtype=np.int32)
        self._check(lib().loco_grid_refine_leaves(self.h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), len(a)))

    def optimal_direction(self, leaf: int) -> int:
        d = C.c_int32()
        self._check(lib().loco_grid_optimal_split_direction(self.h, int(leaf), C.byref(d)))
        return d.value

    def num_leaves(self) -> int:
        return lib().loco_grid_num_leaves(self.h)

    def num_samples(self) -> int:
        return lib().loco_grid_num_samples(self.h)

    def fn_calls(self) -> int:
        return lib().loco_grid_num_function_evaluations(self.h)

    def leaf_neighbors(self, leaf: int):
        a, b = C.c_int32(), C.c_int32()
        self._check(lib().loco_grid_leaf_neighbors(self.h, int(leaf), C.byref(a), C.byref(b)))
        return a.value, b.value

    def graph(self) -> GraphArrays:
        """copy of the current state as GraphArrays (for Model.from_graph / Model.update_from_graph)"""
        s = _GraphStruct()
        self._check(lib().loco_grid_graph(self.h, C.byref(s)))
        N, D, S, Cn, Bc = s.n_params, s.value_dim, s.n_samples, s.n_cells, s.n_border_cells
        f, i32, i64 = np.float64, np.int32, np.int64
        bf_off = _arr(s.bf_offset, S + 1, i64)
        B = int(bf_off[-1])
        c_off = _arr(s.cell_sample_offset, Cn + 1, i64)
        E = int(c_off[-1])
        b_off = _arr(s.border_sample_offset, Bc + 1, i64) if Bc else np.zeros(1, dtype=i64)
        Eb = int(b_off[-1])
        return GraphArrays(
            N=N, D=D, basis=s.basis_type, domain=_arr(s.domain, 2 * N, f), sample_center=_arr(s.sample_center, S * N, f),
            sample_value=_arr(s.sample_value, S * D, f), sample_value_group=_arr(s.sample_value_group, S, i32), bf_offset=bf_off,
            bf_center=_arr(s.bf_center, B * N, f), bf_support=_arr(s.bf_support, B * N, f), bf_weight=_arr(s.bf_weight, B, f),
            cell_ranges=_arr(s.cell_ranges, Cn * 2 * N, f), cell_split_dir=_arr(s.cell_split_dir, Cn, i32), cell_split_val=_arr(s.cell_split_val, Cn, f),
            cell_child2=_arr(s.cell_child2, Cn, i32), cell_center_value=_arr(s.cell_center_value, Cn * D, f) if s.cell_center_value else None,
            cell_sample_offset=c_off,
            cell_sample=_arr(s.cell_sample, E, i32), cell_sample_value=_arr(s.cell_sample_value, E * D, f) if s.cell_sample_value else None,
            border_ranges=_arr(s.border_ranges, Bc * 2 * N, f) if Bc else None, border_center_value=_arr(s.border_center_value, Bc * D, f) if Bc and s.border_center_value else None,
            border_sample_offset=b_off if Bc else None, border_sample=_arr(s.border_sample, Eb, i32) if Bc else None,
            border_sample_value=_arr(s.border_sample_value, Eb * D, f) if Bc and s.border_sample_value else None).normalise()


def adaptive_sample(grid: HostGrid, threshold: float, max_level: int, device: int = 0, error_based_direction: bool = False,
                    max_passes: int = 256) -> List[dict]:
    """LCAdaptiveGrid::adaptiveSample (LCAdaptiveGrid.cpp:481-545) as level-synchronous passes on `grid`: one device handle,
    updated incrementally after every pass (refine.adaptive_sample_passes with the native surgery).  error_based_direction:
    the ERROR_BASED split policy (getOptimalSplitDirection) instead of the cyclic rule.  Returns the per-pass log."""
    from . import refine

    def surgery(leaf, direction):
        grid.refine(leaf, grid.optimal_direction(leaf) if error_based_direction else direction)
    return refine.adaptive_sample_passes(None, surgery, threshold, max_level, max_passes=max_passes, get_graph=grid.graph, device=device)
