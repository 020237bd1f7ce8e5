"""ctypes front-end of tests/cpp/grid_shim.cpp: the native builder state of include/loco_grid.hpp (test helper)."""
import ctypes as C
import os
import subprocess

import numpy as np

from locointerpolate_b200.capi import VALUE_FN, GraphArrays, _GraphStruct

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "grid_shim.cpp")
SO = os.path.join(ROOT, "tests", "cpp", "libgrid_shim.so")
_lib = None


def build(force: bool = False) -> str:
    deps = [SRC, os.path.join(ROOT, "include", "loco_grid.hpp"), os.path.join(ROOT, "include", "loco_b200.h")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(SO) < os.path.getmtime(d) for d in deps):
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), SRC, "-o", SO],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return SO


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.gs_from_graph.restype = vp
        L.gs_from_graph.argtypes = [C.POINTER(_GraphStruct), VALUE_FN, vp]
        L.gs_error.restype = C.c_char_p
        L.gs_error.argtypes = [vp]
        L.gs_free.argtypes = [vp]
        L.gs_refine.argtypes = [vp, C.c_int, C.c_int]
        L.gs_refine_pass.argtypes = [vp, vp, vp, C.c_long]
        for n in ("gs_num_leaves", "gs_num_samples", "gs_fn_calls"):
            getattr(L, n).restype = C.c_long
            getattr(L, n).argtypes = [vp]
        L.gs_graph.restype = C.POINTER(_GraphStruct)
        L.gs_graph.argtypes = [vp]
        L.gs_leaf_neighbors.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.gs_bootstrap.restype = vp
        L.gs_bootstrap.argtypes = [C.c_int] * 5 + [vp, VALUE_FN, vp]
        L.gs_optimal_direction.argtypes = [vp, C.c_int]
        _lib = L
    return _lib


def _arr(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    ct = {np.float64: C.c_double, np.int32: C.c_int32, np.int64: C.c_int64}[dtype]
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()


class NativeGrid:
    """loco_grid::Grid started from a plain-array graph; `true_values(points [n, N] in ORIGINAL parameters) -> [n]` is the function"""

    def __init__(self, g: GraphArrays, true_values):
        self.N, self.D = g.N, g.D
        self._g = g
        self._fn = self._callback(true_values)
        s = g.struct()
        self.h = lib().gs_from_graph(C.byref(s), self._fn, None)
        self._check()

    def _callback(self, true_values):
        def fn(p, out, _user):
            x = np.array([p[i] for i in range(self.N)])
            v = np.atleast_1d(true_values(x[None, :]))
            for j in range(self.D):
                out[j] = float(np.ravel(v)[j])
        return VALUE_FN(fn)

    def _check(self):
        err = lib().gs_error(self.h).decode()
        if err:
            raise RuntimeError(err)

    @classmethod
    def bootstrap(cls, N, basis, level, domain, true_values, eval_corners=True, D=1):
        """LCAdaptiveGrid::bootstrapSample natively (Grid::bootstrap)"""
        self = cls.__new__(cls)
        self.N, self.D = N, D
        self._fn = self._callback(true_values)
        dom = np.ascontiguousarray(domain, dtype=np.float64).reshape(2 * N)
        self.h = lib().gs_bootstrap(N, basis, level, D, int(eval_corners), dom.ctypes.data_as(C.c_void_p), self._fn, None)
        self._check()
        return self

    def optimal_direction(self, leaf):
        d = lib().gs_optimal_direction(self.h, int(leaf))
        if d < 0:
            raise RuntimeError(lib().gs_error(self.h).decode())
        return d

    def __del__(self):
        if getattr(self, "h", None):
            lib().gs_free(self.h)
            self.h = None

    def refine(self, leaf, direction):
        if lib().gs_refine(self.h, int(leaf), int(direction)) != 0:
            raise RuntimeError("native refineCell failed: " + lib().gs_error(self.h).decode())

    def refine_pass(self, leaves, directions):
        a = np.ascontiguousarray(leaves, dtype=np.int32)
        b = np.ascontiguousarray(directions, dtype=np.int32)
        if lib().gs_refine_pass(self.h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), len(a)) != 0:
            raise RuntimeError("native refineCell failed: " + lib().gs_error(self.h).decode())

    def num_leaves(self):
        return lib().gs_num_leaves(self.h)

    def num_samples(self):
        return lib().gs_num_samples(self.h)

    def fn_calls(self):
        return lib().gs_fn_calls(self.h)

    def leaf_neighbors(self, leaf):
        a, b = C.c_int(), C.c_int()
        lib().gs_leaf_neighbors(self.h, int(leaf), C.byref(a), C.byref(b))
        return a.value, b.value

    def graph(self) -> GraphArrays:
        s = lib().gs_graph(self.h).contents
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
            cell_child2=_arr(s.cell_child2, Cn, i32), cell_center_value=_arr(s.cell_center_value, Cn * D, f), cell_sample_offset=c_off,
            cell_sample=_arr(s.cell_sample, E, i32), cell_sample_value=_arr(s.cell_sample_value, E * D, f),
            border_ranges=_arr(s.border_ranges, Bc * 2 * N, f) if Bc else None, border_center_value=_arr(s.border_center_value, Bc * D, f) if Bc else None,
            border_sample_offset=b_off if Bc else None, border_sample=_arr(s.border_sample, Eb, i32) if Bc else None,
            border_sample_value=_arr(s.border_sample_value, Eb * D, f) if Bc else None).normalise()
