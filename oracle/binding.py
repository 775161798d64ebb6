"""ctypes binding of the CPU oracle (test infrastructure — see oracle/te_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libte_oracle.so")


class Geometry(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("resolution", C.c_double),
                ("length_x", C.c_double), ("length_y", C.c_double),
                ("position_x", C.c_double), ("position_y", C.c_double)]

    @classmethod
    def make(cls, rows, cols, resolution, position=(0.0, 0.0)):
        # GridMap::setGeometry: length = size * resolution (doubles)
        return cls(rows, cols, resolution, rows * resolution, cols * resolution, position[0], position[1])


class ChainParams(C.Structure):
    _fields_ = [("normals_radius", C.c_double), ("normals_algorithm", C.c_int32),
                ("normals_positive_axis", C.c_int32), ("slope_critical", C.c_double),
                ("step_critical", C.c_double), ("step_first_radius", C.c_double),
                ("step_second_radius", C.c_double), ("step_critical_cells", C.c_int32),
                ("reserved0", C.c_int32), ("roughness_critical", C.c_double),
                ("roughness_radius", C.c_double), ("fuse_weight", C.c_float), ("reserved1", C.c_int32)]

    @classmethod
    def yaml_defaults(cls, algorithm=0):
        """traversability_estimation/config/robot_filter_parameter.yaml:2-37"""
        return cls(0.05, algorithm, 2, 1.0, 0.12, 0.04, 0.04, 4, 0, 0.05, 0.05,
                   np.float32(1.0) / np.float32(3.0), 0)


class FootprintParams(C.Structure):
    _fields_ = [("radius", C.c_double), ("offset", C.c_double), ("traversability_default", C.c_double),
                ("max_gap_width", C.c_double), ("critical_step_height", C.c_double),
                ("radius_is_integer_norm", C.c_int32), ("verify_roughness", C.c_int32)]

    @classmethod
    def yaml_defaults(cls):
        """robot_footprint_parameter.yaml:5-8, robot.yaml:10, robot_filter_parameter.yaml:18"""
        return cls(0.30, 0.15, 0.3, 0.3, 0.12, 1, 0)


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int32)
        G, P, F = C.POINTER(Geometry), C.POINTER(ChainParams), C.POINTER(FootprintParams)
        L.teo_normals.argtypes = [G, P, fp, fp, fp, fp, C.c_int]
        L.teo_slope.argtypes = [G, C.c_double, fp, fp, C.c_int]
        L.teo_step.argtypes = [G, P, fp, fp, fp, C.c_int]
        L.teo_roughness.argtypes = [G, P, fp, fp, fp, fp, fp, C.c_int]
        L.teo_fuse.argtypes = [C.c_int64, C.c_float, fp, fp, fp, fp]
        L.teo_chain.argtypes = [G, P, fp, fp, fp, fp, fp, fp, fp, fp, C.c_int]
        L.teo_footprint.argtypes = [G, F, fp, fp, fp, fp, fp, fp, fp, C.c_int]
        L.teo_footprint2.argtypes = [G, F, fp, fp, fp, fp, fp, fp, fp, fp, fp, C.c_int]
        L.teo_spiral_offsets.argtypes = [C.c_double, C.c_double, ip, ip, C.c_int]
        L.teo_circle_cells.argtypes = [G, C.c_int, C.c_int, C.c_double, ip, ip, C.c_int]
        L.teo_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _f(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _layer(g, a):
    a = np.asfortranarray(a, dtype=np.float32)
    assert a.shape == (g.rows, g.cols), (a.shape, g.rows, g.cols)
    return a


def _new(g):
    return np.empty((g.rows, g.cols), dtype=np.float32, order="F")


def normals(g, p, elevation, nthreads=0):
    e = _layer(g, elevation)
    nx, ny, nz = _new(g), _new(g), _new(g)
    rc = lib().teo_normals(C.byref(g), C.byref(p), _f(e), _f(nx), _f(ny), _f(nz), nthreads)
    assert rc == 0, rc
    return nx, ny, nz


def slope(g, critical, nz, nthreads=0):
    z = _layer(g, nz)
    out = _new(g)
    assert lib().teo_slope(C.byref(g), critical, _f(z), _f(out), nthreads) == 0
    return out


def step(g, p, elevation, nthreads=0, return_step_height=False):
    e = _layer(g, elevation)
    out, sh = _new(g), _new(g)
    assert lib().teo_step(C.byref(g), C.byref(p), _f(e), _f(out), _f(sh), nthreads) == 0
    return (out, sh) if return_step_height else out


def roughness(g, p, elevation, nx, ny, nz, nthreads=0):
    e, a, b, c = (_layer(g, v) for v in (elevation, nx, ny, nz))
    out = _new(g)
    assert lib().teo_roughness(C.byref(g), C.byref(p), _f(e), _f(a), _f(b), _f(c), _f(out), nthreads) == 0
    return out


def fuse(weight, s, t, r):
    s, t, r = (np.asfortranarray(v, dtype=np.float32) for v in (s, t, r))
    out = np.empty_like(s)
    assert lib().teo_fuse(s.size, np.float32(weight), _f(s), _f(t), _f(r), _f(out)) == 0
    return out


def chain(g, p, elevation, nthreads=0, with_normals=False):
    """Returns dict(slope, step, roughness, traversability[, nx, ny, nz])."""
    e = _layer(g, elevation)
    o = {k: _new(g) for k in ("slope", "step", "roughness", "traversability")}
    n = {k: _new(g) for k in ("nx", "ny", "nz")} if with_normals else {"nx": None, "ny": None, "nz": None}
    rc = lib().teo_chain(C.byref(g), C.byref(p), _f(e), _f(o["slope"]), _f(o["step"]), _f(o["roughness"]),
                         _f(o["traversability"]), _f(n["nx"]), _f(n["ny"]), _f(n["nz"]), nthreads)
    assert rc == 0, rc
    if with_normals:
        o.update(n)
    return o


def footprint(g, fp, traversability, slope_l, step_l, elevation, nthreads=0, roughness=None):
    """Returns (traversability_footprint, slope_footprint, step_footprint[, roughness_footprint when `roughness` is given])."""
    t, s, st, e = (_layer(g, v) for v in (traversability, slope_l, step_l, elevation))
    out, sfp, stfp = _new(g), _new(g), _new(g)
    if roughness is None:
        rc = lib().teo_footprint(C.byref(g), C.byref(fp), _f(t), _f(s), _f(st), _f(e), _f(out), _f(sfp), _f(stfp), nthreads)
        assert rc == 0, rc
        return out, sfp, stfp
    r, rfp = _layer(g, roughness), _new(g)
    rc = lib().teo_footprint2(C.byref(g), C.byref(fp), _f(t), _f(s), _f(st), _f(r), _f(e), _f(out), _f(sfp), _f(stfp), _f(rfp), nthreads)
    assert rc == 0, rc
    return out, sfp, stfp, rfp


def footprint_polygon(g, fp, polygon_xy, yaw, traversability, slope_l, step_l, elevation, nthreads=0, roughness=None):
    """TraversabilityMap::traversabilityFootprint(yaw): (traversability_x, traversability_rot) for the footprint polygon
    `polygon_xy` (vertices (x, y) in the footprint frame)."""
    t, s, st, e = (_layer(g, v) for v in (traversability, slope_l, step_l, elevation))
    r = _layer(g, roughness) if roughness is not None else None
    pts = np.ascontiguousarray(polygon_xy, dtype=np.float64).reshape(-1, 2)
    ox, orot = _new(g), _new(g)
    L = lib()
    L.teo_footprint_polygon.argtypes = [C.POINTER(Geometry), C.POINTER(FootprintParams), C.c_int, C.c_void_p, C.c_double] + [C.c_void_p] * 7 + [C.c_int]
    rc = L.teo_footprint_polygon(C.byref(g), C.byref(fp), len(pts), pts.ctypes.data, float(yaw), t.ctypes.data, s.ctypes.data,
                                 st.ctypes.data, r.ctypes.data if r is not None else None, e.ctypes.data, ox.ctypes.data,
                                 orot.ctypes.data, nthreads)
    assert rc == 0, rc
    return ox, orot


def check_circular_paths(g, footprint_layer, traversability_default, path_begin, poses_xy, robot_slope=None):
    """(is_safe uint8[npaths], traversability float64[npaths]) of TraversabilityMap::checkCircularFootprintPath per path;
    robot_slope: the layer checkInclination reads when checkRobotInclination_ is set (None: off)."""
    f = _layer(g, footprint_layer)
    rs = _layer(g, robot_slope) if robot_slope is not None else None
    pb = np.ascontiguousarray(path_begin, dtype=np.int32)
    xy = np.ascontiguousarray(poses_xy, dtype=np.float64)
    n = len(pb) - 1
    safe = np.zeros(n, dtype=np.uint8)
    trav = np.zeros(n, dtype=np.float64)
    L = lib()
    L.teo_check_circular_paths2.argtypes = [C.POINTER(Geometry), C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = L.teo_check_circular_paths2(C.byref(g), f.ctypes.data, rs.ctypes.data if rs is not None else None, traversability_default, n,
                                     pb.ctypes.data, xy.ctypes.data, safe.ctypes.data, trav.ctypes.data)
    assert rc == 0, rc
    return safe, trav


def spiral_offsets(radius, resolution):
    cap = int((2 * np.ceil(radius / resolution) + 3) ** 2)
    di = np.empty(cap, dtype=np.int32)
    dj = np.empty(cap, dtype=np.int32)
    n = lib().teo_spiral_offsets(radius, resolution, di.ctypes.data_as(C.POINTER(C.c_int32)),
                                 dj.ctypes.data_as(C.POINTER(C.c_int32)), cap)
    assert 0 <= n <= cap, n
    return di[:n].copy(), dj[:n].copy()


def circle_cells(g, i, j, radius):
    cap = int((2 * np.floor(radius / g.resolution) + 5) ** 2)
    a = np.empty(cap, dtype=np.int32)
    b = np.empty(cap, dtype=np.int32)
    n = lib().teo_circle_cells(C.byref(g), i, j, radius, a.ctypes.data_as(C.POINTER(C.c_int32)),
                               b.ctypes.data_as(C.POINTER(C.c_int32)), cap)
    assert 0 <= n <= cap
    return a[:n].copy(), b[:n].copy()


def max_threads():
    return lib().teo_max_threads()
