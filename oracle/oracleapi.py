"""ctypes front-end for oracle/libloco_oracle.so (the CPU restatement) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from locointerpolate_b200.capi import GraphArrays, _GraphStruct

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libloco_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "loco_oracle.cpp")
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "oracle"], check=True, capture_output=True)
    return ORACLE_SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(ORACLE_SO)
        vp, i64 = C.c_void_p, C.c_int64
        L.oracle_new.restype = vp
        L.oracle_new.argtypes = [C.POINTER(_GraphStruct)]
        L.oracle_free.argtypes = [vp]
        L.oracle_num_leaves.argtypes = [vp]
        L.oracle_eval.argtypes = [vp, vp, i64, vp, vp, vp, vp]
        L.oracle_eval_deriv.argtypes = [vp, vp, i64, C.c_int, vp, vp, vp, vp]
        L.oracle_support.argtypes = [vp, C.c_int, vp, C.c_int]
        L.oracle_neighbors.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_error_check.argtypes = [vp, vp, vp, i64, C.c_double, vp, vp, C.POINTER(i64)]
        L.oracle_center_error.argtypes = [vp, C.c_double, vp, vp]
        L.oracle_eval_in_cell.argtypes = [vp, vp, vp, i64, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """CPU restatement of LCSampledFunction built from a plain-array graph."""

    def __init__(self, g: GraphArrays):
        self.g = g
        s = g.struct()
        self.h = lib().oracle_new(C.byref(s))
        self.N, self.D = g.N, g.D
        self.L = lib().oracle_num_leaves(self.h)

    def __del__(self):
        try:
            if self.h:
                lib().oracle_free(self.h)
                self.h = None
        except Exception:
            pass

    def eval(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.N)
        Q = len(q)
        out = np.empty((Q, self.D)); leaf = np.empty(Q, dtype=np.int32)
        status = np.empty(Q, dtype=np.uint8); scale = np.empty(Q)
        lib().oracle_eval(self.h, _p(q), Q, _p(out), _p(leaf), _p(status), _p(scale))
        return (out[:, 0] if self.D == 1 else out), leaf, status, scale

    def eval_deriv(self, q, direction):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.N)
        Q = len(q)
        out = np.empty(Q); leaf = np.empty(Q, dtype=np.int32)
        status = np.empty(Q, dtype=np.uint8); scale = np.empty(Q)
        lib().oracle_eval_deriv(self.h, _p(q), Q, direction, _p(out), _p(leaf), _p(status), _p(scale))
        return out, leaf, status, scale

    def support(self, leaf, cap=1 << 20):
        ids = np.empty(cap, dtype=np.int32)
        k = lib().oracle_support(self.h, leaf, _p(ids), cap)
        if k < 0:
            raise RuntimeError(f"oracle_support failed ({k})")
        return ids[:k].copy()

    def neighbors(self, leaf):
        a, b = C.c_int(), C.c_int()
        lib().oracle_neighbors(self.h, leaf, C.byref(a), C.byref(b))
        return a.value, b.value

    def error_check(self, q, truth, threshold):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.N)
        truth = np.ascontiguousarray(truth, dtype=np.float64).reshape(len(q), self.D)
        err = np.empty(self.L); split = np.empty(self.L, dtype=np.uint8); bad = C.c_int64()
        lib().oracle_error_check(self.h, _p(q), _p(truth), len(q), threshold, _p(err), _p(split), C.byref(bad))
        return err, split, bad.value

    def center_error(self, threshold):
        err = np.empty(self.L); split = np.empty(self.L, dtype=np.uint8)
        rc = lib().oracle_center_error(self.h, threshold, _p(err), _p(split))
        if rc != 0:
            raise RuntimeError("oracle: model has no centre values")
        return err, split

    def eval_in_cell(self, x, leaf):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, self.N)
        leaf = np.ascontiguousarray(leaf, dtype=np.int32)
        out = np.empty((len(x), self.D)); status = np.empty(len(x), dtype=np.uint8)
        lib().oracle_eval_in_cell(self.h, _p(x), _p(leaf), len(x), _p(out), _p(status))
        return (out[:, 0] if self.D == 1 else out), status


def graph_from_ref(rg) -> GraphArrays:
    """refapi.Graph (dump of the compiled reference) -> GraphArrays"""
    return GraphArrays(
        N=rg.N, D=1, basis=rg.basis, domain=rg.ranges, sample_center=rg.centers, sample_value=rg.values,
        sample_value_group=rg.value_group, bf_offset=rg.bf_off.astype(np.int64), bf_center=rg.bf_center,
        bf_support=rg.bf_support, bf_weight=rg.bf_weight, cell_ranges=rg.cell_ranges,
        cell_split_dir=np.where(rg.cell_is_leaf == 1, -1, rg.cell_split_dir).astype(np.int32),
        cell_split_val=rg.cell_split_val, cell_center_value=rg.cell_center_val,
        cell_sample_offset=rg.cell_ls_off.astype(np.int64), cell_sample=rg.cell_ls_sample, cell_sample_value=rg.cell_ls_value,
        border_ranges=rg.border_ranges, border_center_value=rg.border_center_val,
        border_sample_offset=rg.border_ls_off.astype(np.int64), border_sample=rg.border_ls_sample,
        border_sample_value=rg.border_ls_value).normalise()
