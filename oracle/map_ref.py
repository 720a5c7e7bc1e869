"""TEST INFRASTRUCTURE ONLY -- ctypes loader of oracle/_ref/libmapref.so: the reference's own map::Grid (built by
oracle/build_map_ref.py from the sources where they lie).  Used to pin oracle/map_grid.py and to make tests/golden/map_grid_*.npz."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libmapref.so")
_lib = None


def available():
    return os.path.exists(LIB)


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        dp, ip, bp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int8)
        _lib.mapref_build.argtypes = [dp, ip, C.c_int, C.c_double, C.c_double, C.c_double, bp, C.c_int, ip, ip, dp]
        _lib.mapref_build.restype = C.c_int
        _lib.mapref_reveal.argtypes = [dp, ip, C.c_int, C.c_double, C.c_double, C.c_double, ip, C.c_int, C.c_int, bp, C.c_int]
        _lib.mapref_reveal.restype = C.c_int
    return _lib


def _flat(obstacles):
    xy = np.ascontiguousarray(np.array([c for ob in obstacles for v in ob for c in v], dtype=np.float64))
    nv = np.ascontiguousarray(np.array([len(ob) for ob in obstacles], dtype=np.int32))
    return xy, nv


def build(obstacles, scale, resolution, inflate):
    """-> (cells int8 (H, W), resolution, origin (2,)) from the reference's Grid::build_map / occupancy_grid."""
    lib = _load()
    xy, nv = _flat(obstacles)
    dp, ip, bp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int8)
    W, H = C.c_int(), C.c_int()
    org = np.zeros(2)
    n = lib.mapref_build(xy.ctypes.data_as(dp), nv.ctypes.data_as(ip), len(nv), scale, resolution, inflate, None, 0,
                         C.byref(W), C.byref(H), org.ctypes.data_as(dp))
    cells = np.zeros(n, dtype=np.int8)
    lib.mapref_build(xy.ctypes.data_as(dp), nv.ctypes.data_as(ip), len(nv), scale, resolution, inflate, cells.ctypes.data_as(bp), n,
                     C.byref(W), C.byref(H), org.ctypes.data_as(dp))
    return cells.reshape(H.value, W.value), float(resolution), org


def reveal(obstacles, scale, resolution, inflate, path_ixiy, visibility):
    lib = _load()
    xy, nv = _flat(obstacles)
    dp, ip, bp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int8)
    g, _, _ = build(obstacles, scale, resolution, inflate)
    path = np.ascontiguousarray(np.asarray(path_ixiy, dtype=np.int32).reshape(-1, 2))
    fake = np.zeros(g.size, dtype=np.int8)
    lib.mapref_reveal(xy.ctypes.data_as(dp), nv.ctypes.data_as(ip), len(nv), scale, resolution, inflate,
                      path.ctypes.data_as(ip), len(path), visibility, fake.ctypes.data_as(bp), fake.size)
    return fake.reshape(g.shape)
