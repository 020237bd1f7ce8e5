"""ctypes binding of include/loco_b200.h (libloco_b200.so).

This is plumbing for tests, bench.py and Python callers: every call goes straight through the C ABI.
There is no Python or CPU implementation of the evaluation path here -- if the CUDA library is missing
or no GPU is visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libloco_b200.so")

HOST_ONLY = -1
LINEAR, CUBIC = 0, 1
ST_OK, ST_OUTSIDE_CELL, ST_BAD_ARITY, ST_NO_COMMON_SAMPLE, ST_EMPTY_SUPPORT = range(5)

_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double


class LocoError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[loco error {code}] {msg}")
        self.code, self.msg = code, msg


class _GraphStruct(C.Structure):
    _fields_ = [
        ("n_params", _i32), ("value_dim", _i32), ("basis_type", _i32), ("domain", _vp),
        ("n_samples", _i64), ("sample_center", _vp), ("sample_value", _vp), ("sample_value_group", _vp),
        ("bf_offset", _vp), ("bf_center", _vp), ("bf_support", _vp), ("bf_weight", _vp),
        ("n_cells", _i64), ("cell_ranges", _vp), ("cell_split_dir", _vp), ("cell_split_val", _vp), ("cell_child2", _vp),
        ("cell_center_value", _vp), ("cell_sample_offset", _vp), ("cell_sample", _vp), ("cell_sample_value", _vp),
        ("n_border_cells", _i64), ("border_ranges", _vp), ("border_center_value", _vp), ("border_sample_offset", _vp),
        ("border_sample", _vp), ("border_sample_value", _vp),
    ]


class _Info(C.Structure):
    _fields_ = [(n, _i32) for n in ("n_params", "value_dim", "basis_type", "exact_rcp", "depth", "max_support", "max_terms", "device")] + \
               [(n, _i64) for n in ("n_samples", "n_value_rows", "n_leaves", "n_cells", "n_border_cells", "n_basis_functions",
                                    "n_support_entries", "n_terms", "device_bytes", "n_fold_entries", "n_eval_cells", "n_eval_terms", "n_poly_cells",
                                    "n_reused_leaves")]


VALUE_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp)

_lib = None

# name -> (restype, argtypes); also the list of symbols the header declares (checked by tests)
SIGNATURES = {
    "loco_model_from_proto": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(_vp)]),
    "loco_model_from_graph": (C.c_int, [C.POINTER(_GraphStruct), C.c_int, C.POINTER(_vp)]),
    "loco_model_update_from_graph": (C.c_int, [_vp, C.POINTER(_GraphStruct)]),
    "loco_model_bootstrap": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, C.c_int, C.POINTER(_vp)]),
    "loco_model_save_proto": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "loco_free": (None, [_vp]),
    "loco_model_info": (C.c_int, [_vp, C.POINTER(_Info)]),
    "loco_support_set": (C.c_int, [_vp, _i32, _vp, _i32, C.POINTER(_i32)]),
    "loco_leaf_neighbors": (C.c_int, [_vp, _i32, C.POINTER(_i32), C.POINTER(_i32)]),
    "loco_last_error": (C.c_char_p, [_vp]),
    "loco_status_string": (C.c_char_p, [C.c_int]),
    "loco_sample_value_rows": (C.c_int, [_vp, _vp]),
    "loco_leaf_parent_split": (C.c_int, [_vp, _i32, C.POINTER(_i32)]),
    "loco_model_set_values": (C.c_int, [_vp, _vp]),
    "loco_model_fill_values_synthetic": (C.c_int, [_vp, _f64]),
    "loco_model_get_values": (C.c_int, [_vp, _i64, _i64, _vp]),
    "loco_eval_batch": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "loco_eval_batch_device": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "loco_eval_batch_ring_device": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp]),
    "loco_eval_batch_ring": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "loco_model_get_value_columns": (C.c_int, [_vp, _i64, _i64, _vp]),
    "loco_model_domain": (C.c_int, [_vp, _vp]),
    "loco_leaf_box": (C.c_int, [_vp, _i32, _vp]),
    "loco_merged_terms": (C.c_int, [_vp, _i32, _vp, _vp, _i32, C.POINTER(_i32), C.POINTER(_i64)]),
    "loco_eval_cells": (C.c_int, [_vp, _i32, _vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "loco_eval_cell_terms": (C.c_int, [_vp, _i64, _vp, _i32, C.POINTER(_i32)]),
    "loco_eval_cell_poly": (C.c_int, [_vp, _i64, _vp, _vp, _i32, C.POINTER(_i32)]),
    "loco_eval_deriv_batch": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "loco_eval_deriv_batch_device": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "loco_error_check": (C.c_int, [_vp, _vp, _vp, _i64, _f64, _vp, _vp, C.POINTER(_i64)]),
    "loco_error_check_device": (C.c_int, [_vp, _vp, _vp, _i64, _f64, _vp, _vp, _vp, _vp]),
    "loco_center_error": (C.c_int, [_vp, _f64, _vp, _vp]),
    "loco_error_check_generated_device": (C.c_int, [_vp, _i64, C.c_uint64, _vp, _vp, _i32, _f64, _vp, _vp, _vp]),
    "loco_peer_buffer_create": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(_vp), C.c_char_p]),
    "loco_peer_buffer_open": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(_vp)]),
    "loco_peer_buffer_close": (C.c_int, [C.c_int, _vp]),
    "loco_peer_buffer_free": (C.c_int, [C.c_int, _vp]),
    "loco_peer_copy_async": (C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    "loco_set_option": (C.c_int, [_vp, C.c_char_p, _i64]),
    "loco_kernel_launches": (_i64, [_vp]),
    "loco_profile_begin": (C.c_int, [_vp]),
    "loco_profile_end": (C.c_int, [_vp, C.c_char_p, C.c_size_t]),
}


def lib():
    """Load libloco_b200.so (built by locointerpolate_b200.build); raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not built: run `python -m locointerpolate_b200.build` (there is no fallback path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_vp)


@dataclass
class GraphArrays:
    """numpy image of loco_graph_t (see include/loco_b200.h)."""
    N: int
    D: int
    basis: int
    domain: np.ndarray
    sample_center: np.ndarray
    sample_value: np.ndarray
    bf_offset: np.ndarray
    bf_center: np.ndarray
    bf_support: np.ndarray
    bf_weight: np.ndarray
    cell_ranges: np.ndarray
    cell_split_dir: np.ndarray
    cell_split_val: np.ndarray
    cell_sample_offset: np.ndarray
    cell_sample: np.ndarray
    sample_value_group: Optional[np.ndarray] = None
    cell_child2: Optional[np.ndarray] = None
    cell_center_value: Optional[np.ndarray] = None
    cell_sample_value: Optional[np.ndarray] = None
    border_ranges: Optional[np.ndarray] = None
    border_center_value: Optional[np.ndarray] = None
    border_sample_offset: Optional[np.ndarray] = None
    border_sample: Optional[np.ndarray] = None
    border_sample_value: Optional[np.ndarray] = None
    _keep: list = field(default_factory=list, repr=False)

    def normalise(self):
        N, D = self.N, self.D
        f = lambda a: None if a is None else _np(a, np.float64)
        self.domain = f(self.domain).reshape(2 * N)
        self.sample_center = f(self.sample_center).reshape(-1, N)
        self.sample_value = f(self.sample_value).reshape(-1, D)
        self.bf_offset = _np(self.bf_offset, np.int64)
        self.bf_center = f(self.bf_center).reshape(-1, N)
        self.bf_support = f(self.bf_support).reshape(-1, N)
        self.bf_weight = f(self.bf_weight).reshape(-1)
        self.cell_ranges = f(self.cell_ranges).reshape(-1, 2 * N)
        self.cell_split_dir = _np(self.cell_split_dir, np.int32)
        self.cell_split_val = f(self.cell_split_val)
        self.cell_sample_offset = _np(self.cell_sample_offset, np.int64)
        self.cell_sample = _np(self.cell_sample, np.int32)
        if self.sample_value_group is not None:
            self.sample_value_group = _np(self.sample_value_group, np.int32)
        if self.cell_child2 is None:
            self.cell_child2 = child2_from_preorder(self.cell_split_dir)
        self.cell_child2 = _np(self.cell_child2, np.int32)
        if self.cell_center_value is not None:
            self.cell_center_value = f(self.cell_center_value).reshape(-1, D)
        if self.cell_sample_value is not None:
            self.cell_sample_value = f(self.cell_sample_value).reshape(-1, D)
        if self.border_ranges is None or len(self.border_ranges) == 0:
            self.border_ranges = np.zeros((0, 2 * N))
            self.border_sample_offset = np.zeros(1, dtype=np.int64)
            self.border_sample = np.zeros(0, dtype=np.int32)
            self.border_center_value = None
            self.border_sample_value = None
        self.border_ranges = f(self.border_ranges).reshape(-1, 2 * N)
        self.border_sample_offset = _np(self.border_sample_offset, np.int64)
        self.border_sample = _np(self.border_sample, np.int32)
        if self.border_center_value is not None:
            self.border_center_value = f(self.border_center_value).reshape(-1, D)
        if self.border_sample_value is not None:
            self.border_sample_value = f(self.border_sample_value).reshape(-1, D)
        return self

    def struct(self) -> _GraphStruct:
        self.normalise()
        g = _GraphStruct()
        g.n_params, g.value_dim, g.basis_type = self.N, self.D, self.basis
        g.domain = _ptr(self.domain)
        g.n_samples = len(self.sample_center)
        g.sample_center, g.sample_value = _ptr(self.sample_center), _ptr(self.sample_value)
        g.sample_value_group = _ptr(self.sample_value_group)
        g.bf_offset, g.bf_center, g.bf_support, g.bf_weight = map(_ptr, (self.bf_offset, self.bf_center, self.bf_support, self.bf_weight))
        g.n_cells = len(self.cell_split_dir)
        g.cell_ranges, g.cell_split_dir, g.cell_split_val, g.cell_child2 = map(
            _ptr, (self.cell_ranges, self.cell_split_dir, self.cell_split_val, self.cell_child2))
        g.cell_center_value = _ptr(self.cell_center_value)
        g.cell_sample_offset, g.cell_sample = _ptr(self.cell_sample_offset), _ptr(self.cell_sample)
        g.cell_sample_value = _ptr(self.cell_sample_value)
        g.n_border_cells = len(self.border_ranges)
        g.border_ranges = _ptr(self.border_ranges)
        g.border_center_value = _ptr(self.border_center_value)
        g.border_sample_offset, g.border_sample = _ptr(self.border_sample_offset), _ptr(self.border_sample)
        g.border_sample_value = _ptr(self.border_sample_value)
        return g


def child2_from_preorder(split_dir) -> np.ndarray:
    """child2 index of every inner cell of a pre-order cell list (child1 = index+1)."""
    split_dir = np.asarray(split_dir)
    n = len(split_dir)
    size = np.ones(n, dtype=np.int64)
    child2 = np.full(n, -1, dtype=np.int32)
    for i in range(n - 1, -1, -1):
        if split_dir[i] >= 0:
            c1 = i + 1
            c2 = c1 + size[c1]
            child2[i] = c2
            size[i] = 1 + size[c1] + size[c2]
    return child2


def _check0(rc):
    if rc != 0:
        raise LocoError(rc, lib().loco_last_error(None).decode())


def peer_buffer_create(device: int, nbytes: int):
    """(device pointer, 64-byte handle) of a fresh device buffer other processes of the box can map"""
    p = _vp()
    h = C.create_string_buffer(64)
    _check0(lib().loco_peer_buffer_create(device, nbytes, C.byref(p), h))
    return p.value, h.raw


def peer_buffer_open(device: int, handle: bytes) -> int:
    p = _vp()
    _check0(lib().loco_peer_buffer_open(device, C.create_string_buffer(handle, 64), C.byref(p)))
    return p.value


def peer_buffer_close(device: int, ptr: int):
    _check0(lib().loco_peer_buffer_close(device, ptr))


def peer_buffer_free(device: int, ptr: int):
    _check0(lib().loco_peer_buffer_free(device, ptr))


def peer_copy_async(dst: int, src: int, nbytes: int, stream: int = 0):
    _check0(lib().loco_peer_copy_async(dst, src, nbytes, stream or None))


class Model:
    """Handle to a loco_model_t."""

    def __init__(self, handle):
        self._h = handle
        self._refresh_info()

    def _refresh_info(self):
        info = _Info()
        self._check(lib().loco_model_info(self._h, C.byref(info)))
        self.info = {n: getattr(info, n) for n, _ in _Info._fields_}
        self.N, self.D, self.L = info.n_params, info.value_dim, info.n_leaves

    def update_from_graph(self, g: "GraphArrays"):
        """the same handle after the grid changed (refinement loop): incremental re-flatten + re-upload"""
        s = g.struct()
        self._check(lib().loco_model_update_from_graph(self._h, C.byref(s)))
        self._refresh_info()

    # ---- constructors
    @staticmethod
    def _new(rc, h):
        if rc != 0:
            raise LocoError(rc, lib().loco_last_error(None).decode())
        return Model(h)

    @classmethod
    def from_proto(cls, path: str, ascii: bool = False, device: int = 0) -> "Model":
        h = _vp()
        return cls._new(lib().loco_model_from_proto(path.encode(), int(ascii), device, C.byref(h)), h)

    @classmethod
    def from_graph(cls, g: GraphArrays, device: int = 0) -> "Model":
        h = _vp()
        s = g.struct()
        return cls._new(lib().loco_model_from_graph(C.byref(s), device, C.byref(h)), h)

    @classmethod
    def bootstrap(cls, N: int, basis: int, level: int, D: int = 1, eval_corners: bool = True, domain=None, fn=None,
                  device: int = 0) -> "Model":
        dom = _np(domain if domain is not None else [0.0, 1.0] * N, np.float64)
        cb = None
        if fn is not None:
            def _cb(p, o, _u):
                v = np.atleast_1d(np.asarray(fn(np.ctypeslib.as_array(p, shape=(N,))), dtype=np.float64))
                for j in range(D):
                    o[j] = v[j]
            cb = VALUE_FN(_cb)
        h = _vp()
        rc = lib().loco_model_bootstrap(N, basis, level, D, int(eval_corners), _ptr(dom), C.cast(cb, _vp) if cb else None, None, device, C.byref(h))
        return cls._new(rc, h)

    # ---- plumbing
    def _check(self, rc):
        if rc != 0:
            raise LocoError(rc, lib().loco_last_error(self._h).decode())

    def free(self):
        if self._h:
            lib().loco_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def save_proto(self, path: str, ascii: bool = False):
        self._check(lib().loco_model_save_proto(self._h, path.encode(), int(ascii)))

    def support_set(self, leaf: int) -> np.ndarray:
        k = _i32()
        self._check(lib().loco_support_set(self._h, leaf, None, 0, C.byref(k)))
        ids = np.empty(k.value, dtype=np.int32)
        self._check(lib().loco_support_set(self._h, leaf, _ptr(ids), k.value, C.byref(k)))
        return ids

    def leaf_neighbors(self, leaf: int):
        a, b = _i32(), _i32()
        self._check(lib().loco_leaf_neighbors(self._h, leaf, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_option(self, name: str, value: int):
        self._check(lib().loco_set_option(self._h, name.encode(), value))

    @property
    def kernel_launches(self) -> int:
        return lib().loco_kernel_launches(self._h)

    def profile_begin(self):
        self._check(lib().loco_profile_begin(self._h))

    def profile_end(self) -> dict:
        """{kernel name: {"launches": n, "total_ms": t}} for every kernel launched since profile_begin (device synchronised)"""
        import json
        buf = C.create_string_buffer(1 << 16)
        self._check(lib().loco_profile_end(self._h, buf, len(buf)))
        return json.loads(buf.value.decode())

    def sample_value_rows(self) -> np.ndarray:
        """value-table row of every sample (samples sharing one value object share a row)"""
        rows = np.empty(self.info["n_samples"], dtype=np.int32)
        self._check(lib().loco_sample_value_rows(self._h, _ptr(rows)))
        return rows

    def leaf_parent_split(self, leaf: int) -> int:
        d = _i32()
        self._check(lib().loco_leaf_parent_split(self._h, leaf, C.byref(d)))
        return d.value

    def set_values(self, values):
        v = _np(values, np.float64).reshape(self.info["n_value_rows"], self.D)
        self._check(lib().loco_model_set_values(self._h, _ptr(v)))

    def fill_values_synthetic(self, seed: float = 0.0):
        self._check(lib().loco_model_fill_values_synthetic(self._h, float(seed)))

    def get_values(self, row0: int, n: int) -> np.ndarray:
        out = np.empty((n, self.D))
        self._check(lib().loco_model_get_values(self._h, row0, n, _ptr(out)))
        return out

    # ---- hot path, host buffers (numpy)
    def eval(self, q, want_leaf=True, want_status=True):
        q = _np(q, np.float64).reshape(-1, self.N)
        Q = len(q)
        out = np.empty((Q, self.D))
        leaf = np.empty(Q, dtype=np.int32) if want_leaf else None
        status = np.empty(Q, dtype=np.uint8) if want_status else None
        self._check(lib().loco_eval_batch(self._h, _ptr(q), Q, _ptr(out), _ptr(leaf), _ptr(status)))
        return (out[:, 0] if self.D == 1 else out), leaf, status

    def eval_raw(self, q_ptr: int, Q: int, out_ptr: int, leaf_ptr: int = 0, status_ptr: int = 0):
        """host pointers (e.g. pinned torch tensors' data_ptr())"""
        self._check(lib().loco_eval_batch(self._h, q_ptr, Q, out_ptr, leaf_ptr or None, status_ptr or None))

    # ---- hot path, device buffers (raw pointers; torch tensors' data_ptr())
    def eval_device(self, q_ptr: int, Q: int, out_ptr: int, leaf_ptr: int = 0, status_ptr: int = 0, stream: int = 0):
        self._check(lib().loco_eval_batch_device(self._h, q_ptr, Q, out_ptr, leaf_ptr or None, status_ptr or None, stream or None))

    def eval_ring_device(self, q_ptr: int, Q: int, ring_ptr: int, ring_queries: int, checksum_ptr: int, leaf_ptr: int = 0,
                         status_ptr: int = 0, stream: int = 0):
        self._check(lib().loco_eval_batch_ring_device(self._h, q_ptr, Q, ring_ptr, ring_queries, checksum_ptr, leaf_ptr or None,
                                                      status_ptr or None, stream or None))

    def eval_ring_raw(self, q_ptr: int, Q: int, ring_queries: int, checksum_ptr: int, leaf_ptr: int = 0, status_ptr: int = 0):
        """host pointers: queries in, per-query checksums (+ leaf, status) out; the ring lives inside the handle"""
        self._check(lib().loco_eval_batch_ring(self._h, q_ptr, Q, ring_queries, checksum_ptr, leaf_ptr or None, status_ptr or None))

    def get_value_columns(self, col0: int, n: int) -> np.ndarray:
        out = np.empty((self.info["n_value_rows"], n))
        self._check(lib().loco_model_get_value_columns(self._h, col0, n, _ptr(out)))
        return out

    # ---- derivative along one cube direction (scalar functions)
    def eval_deriv(self, q, direction: int):
        q = _np(q, np.float64).reshape(-1, self.N)
        Q = len(q)
        out = np.empty(Q)
        leaf = np.empty(Q, dtype=np.int32)
        status = np.empty(Q, dtype=np.uint8)
        self._check(lib().loco_eval_deriv_batch(self._h, _ptr(q), Q, direction, _ptr(out), _ptr(leaf), _ptr(status)))
        return out, leaf, status

    def eval_deriv_device(self, q_ptr: int, Q: int, direction: int, out_ptr: int, leaf_ptr: int = 0, status_ptr: int = 0, stream: int = 0):
        self._check(lib().loco_eval_deriv_batch_device(self._h, q_ptr, Q, direction, out_ptr, leaf_ptr or None, status_ptr or None, stream or None))

    @property
    def domain(self) -> np.ndarray:
        r = np.empty(2 * self.N)
        self._check(lib().loco_model_domain(self._h, _ptr(r)))
        return r

    def leaf_box(self, leaf: int) -> np.ndarray:
        r = np.empty(2 * self.N)
        self._check(lib().loco_leaf_box(self._h, leaf, _ptr(r)))
        return r.reshape(self.N, 2)

    # ---- compacted term lists (introspection; host-only handles)
    def merged_terms(self, leaf: int):
        """(centre [n, N], support [n, N], first_id) of the leaf's merged basis functions"""
        n, first = _i32(), _i64()
        self._check(lib().loco_merged_terms(self._h, leaf, None, None, 0, C.byref(n), C.byref(first)))
        c, s = np.empty((n.value, self.N)), np.empty((n.value, self.N))
        if n.value:
            self._check(lib().loco_merged_terms(self._h, leaf, _ptr(c), _ptr(s), n.value, C.byref(n), C.byref(first)))
        return c, s, first.value

    def eval_cells(self, leaf: int):
        """(bits [N], first_cell, n_cells) of the leaf's evaluation sub-grid"""
        bits = np.zeros(self.N, dtype=np.int32)
        first, cnt = _i64(), _i64()
        self._check(lib().loco_eval_cells(self._h, leaf, _ptr(bits), C.byref(first), C.byref(cnt)))
        return bits, first.value, cnt.value

    def eval_cell_terms(self, cell: int) -> np.ndarray:
        n = _i32()
        self._check(lib().loco_eval_cell_terms(self._h, cell, None, 0, C.byref(n)))
        ids = np.empty(n.value, dtype=np.int64)
        if n.value:
            self._check(lib().loco_eval_cell_terms(self._h, cell, _ptr(ids), n.value, C.byref(n)))
        return ids

    def eval_cell_poly(self, cell: int):
        """(geom [N, 2] = mid, 1/halfwidth; coef [F^N], dimension 0 fastest) or None if the cell keeps its term list"""
        n = _i32()
        self._check(lib().loco_eval_cell_poly(self._h, cell, None, None, 0, C.byref(n)))
        if n.value == 0:
            return None
        geom, coef = np.empty(2 * self.N), np.empty(n.value)
        self._check(lib().loco_eval_cell_poly(self._h, cell, _ptr(geom), _ptr(coef), n.value, C.byref(n)))
        return geom.reshape(self.N, 2), coef

    # ---- error check
    def error_check(self, q, truth, threshold: float):
        q = _np(q, np.float64).reshape(-1, self.N)
        truth = _np(truth, np.float64).reshape(len(q), self.D)
        err = np.empty(self.L)
        split = np.empty(self.L, dtype=np.uint8)
        bad = _i64()
        self._check(lib().loco_error_check(self._h, _ptr(q), _ptr(truth), len(q), threshold, _ptr(err), _ptr(split), C.byref(bad)))
        return err, split, bad.value

    def error_check_device(self, q_ptr, truth_ptr, P, threshold, err_ptr, split_ptr, nbad_ptr=0, stream=0):
        self._check(lib().loco_error_check_device(self._h, q_ptr, truth_ptr, P, threshold, err_ptr, split_ptr, nbad_ptr or None, stream or None))

    def center_error(self, threshold: float):
        err = np.empty(self.L)
        split = np.empty(self.L, dtype=np.uint8)
        self._check(lib().loco_center_error(self._h, threshold, _ptr(err), _ptr(split)))
        return err, split

    def error_check_generated_device(self, points_per_leaf, seed, amp_ptr, phase_ptr, n_terms, threshold, err_ptr, split_ptr, stream=0):
        self._check(lib().loco_error_check_generated_device(self._h, points_per_leaf, seed, amp_ptr, phase_ptr, n_terms, threshold,
                                                            err_ptr, split_ptr, stream or None))
