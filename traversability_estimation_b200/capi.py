"""ctypes view of include/te_b200.h."""
from __future__ import annotations

import ctypes as C
import os
import sys
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# TE_B200_LIBRARY: development only — load a variant built with `make -C csrc variant NAME=... EXTRA=...` (several builds
# of the kernels can then be timed in one GPU session); the product library is the one next to this file.
_LIB = os.environ.get("TE_B200_LIBRARY") or os.path.join(_HERE, "libte_b200.so")

MEM_HOST, MEM_DEVICE = 0, 1
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_FUSED = 0, 1, 2

EXPORTS = ["te_create", "te_destroy", "te_last_error", "te_abi_version", "te_set_stream", "te_synchronize",
           "te_set_kernel", "te_get_stats", "te_enable_timing", "te_get_timing", "te_get_flag_counters", "te_get_escalation_stats", "te_fused_plan", "te_slope", "te_normals", "te_step", "te_roughness", "te_chain",
           "te_chain_batched", "te_footprint", "te_footprint2", "te_footprint_polygon", "te_check_footprint_paths", "te_check_footprint_paths2", "te_ipc_export", "te_ipc_open", "te_ipc_close", "te_event_create_ipc", "te_event_open_ipc",
           "te_event_record", "te_event_destroy", "te_halo_pull", "te_host_alloc", "te_host_free"]


IPC_HANDLE_BYTES = 80  # TE_IPC_HANDLE_BYTES


class TEError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"te_b200 error {code}: {msg}")
        self.code = code


class Geometry(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("resolution", C.c_double),
                ("length_x", C.c_double), ("length_y", C.c_double),
                ("position_x", C.c_double), ("position_y", C.c_double),
                ("start_row", C.c_int32), ("start_col", C.c_int32)]

    @classmethod
    def make(cls, rows, cols, resolution, position=(0.0, 0.0)):
        # GridMap::setGeometry: length = size * resolution
        return cls(rows, cols, resolution, rows * resolution, cols * resolution, position[0], position[1], 0, 0)


class Slab(C.Structure):
    _fields_ = [("col_begin", C.c_int32), ("col_count", C.c_int32), ("halo_left", C.c_int32), ("halo_right", C.c_int32)]


class HaloPeer(C.Structure):
    """te_halo_peer: a neighbour's slab buffer as mapped into this process."""
    _fields_ = [("layer", C.c_void_p), ("slab", Slab), ("ready_event", C.c_void_p)]


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


def library_path() -> str:
    return _LIB


def build_library(force: bool = False) -> str:
    """nvcc-compile csrc/ for sm_100a into libte_b200.so (in-tree, travels to the GPU box)."""
    args = ["make", "-C", os.path.join(_HERE, "csrc"), "-s"]
    if force:
        args.append("-B")
    subprocess.check_call(args)
    return _LIB


_lib = None


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise ImportError(f"{_LIB} is missing: run traversability_estimation_b200.build_library() "
                              "(there is no CPU fallback)")
        L = C.CDLL(_LIB)
        vp, fp = C.c_void_p, C.c_void_p  # layers are passed as raw addresses (host or device)
        G, S = C.POINTER(Geometry), C.POINTER(Slab)
        P, F = C.POINTER(ChainParams), C.POINTER(FootprintParams)
        L.te_create.argtypes = [C.POINTER(vp), C.c_int]
        L.te_destroy.argtypes = [vp]
        L.te_last_error.restype = C.c_char_p
        L.te_set_stream.argtypes = [vp, vp]
        L.te_synchronize.argtypes = [vp]
        L.te_set_kernel.argtypes = [vp, C.c_int]
        L.te_get_stats.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.te_enable_timing.argtypes = [vp, C.c_int]
        L.te_get_timing.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.te_slope.argtypes = [vp, G, C.c_double, fp, fp, C.c_int]
        L.te_normals.argtypes = [vp, G, P, fp, fp, fp, fp, C.c_int]
        L.te_step.argtypes = [vp, G, P, fp, fp, C.c_int]
        L.te_roughness.argtypes = [vp, G, P, fp, fp, fp, fp, fp, C.c_int]
        L.te_chain.argtypes = [vp, G, S, P, fp, fp, fp, fp, fp, fp, fp, fp, C.c_int]
        L.te_chain_batched.argtypes = [vp, G, P, C.c_int32, fp, fp, fp, fp, fp, C.c_int]
        L.te_footprint.argtypes = [vp, G, S, F, fp, fp, fp, fp, fp, fp, fp, C.c_int]
        L.te_footprint2.argtypes = [vp, G, S, F, fp, fp, fp, fp, fp, fp, fp, fp, fp, C.c_int]
        L.te_ipc_export.argtypes = [vp, vp]
        L.te_ipc_open.argtypes = [vp, C.POINTER(vp)]
        L.te_ipc_close.argtypes = [vp]
        L.te_event_create_ipc.argtypes = [vp, C.POINTER(vp), vp]
        L.te_event_open_ipc.argtypes = [vp, C.POINTER(vp)]
        L.te_event_record.argtypes = [vp, vp]
        L.te_event_destroy.argtypes = [vp]
        L.te_halo_pull.argtypes = [vp, G, S, fp, C.POINTER(HaloPeer), C.POINTER(HaloPeer)]
        _lib = L
    return _lib


def fused_plan(rows: int, out_ncols: int, nmaps: int = 1, sms: int = 148) -> dict:
    """te_fused_plan: the (level, map, segment, strip) work units of the fused launch; host arithmetic, no GPU needed."""
    L = load_library()
    out = (C.c_int32 * 19)()
    L.te_fused_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    rc = L.te_fused_plan(rows, out_ncols, nmaps, sms, out)
    if rc != 0:
        raise TEError(rc, "te_fused_plan: invalid argument")
    nl = out[1]
    levels = [dict(unit0=out[2 + 4 * i], col0=out[3 + 4 * i], seg_len=out[4 + 4 * i], nseg=out[5 + 4 * i]) for i in range(nl)]
    return {"strips": out[0], "levels": levels, "units": out[2 + 4 * nl]}


def _addr(a):
    """Address of a numpy array (host) / torch tensor (device) / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


class Context:
    """One te_ctx (one per rank / plugin instance)."""

    def __init__(self, device: int = 0):
        self._L = load_library()
        h = C.c_void_p()
        self._check(self._L.te_create(C.byref(h), device))
        self._h = h
        self._own_stream = True   # device-memory calls run on the context's own (non-blocking) stream until set_stream(ptr)

    def _order_after_torch(self, memory):
        """This ctypes view is used with torch tensors as device memory.  While the context runs on its own stream nothing orders
        its kernels after the torch kernels that produce their inputs: drain torch's current stream first (callers that pass
        their stream with set_stream need no such thing, nor does the C ABI itself — ordering is the caller's there)."""
        if memory == MEM_DEVICE and self._own_stream and "torch" in sys.modules:
            torch = sys.modules["torch"]
            if torch.cuda.is_available():
                torch.cuda.current_stream().synchronize()

    def _check(self, rc):
        if rc != 0:
            raise TEError(rc, self._L.te_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.te_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        self._check(self._L.te_set_stream(self._h, stream_ptr))
        self._own_stream = not stream_ptr

    def synchronize(self):
        self._check(self._L.te_synchronize(self._h))

    def set_kernel(self, choice):
        self._check(self._L.te_set_kernel(self._h, choice))

    def stats(self):
        a, b = C.c_int64(), C.c_int64()
        self._check(self._L.te_get_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def enable_timing(self, on=True):
        self._check(self._L.te_enable_timing(self._h, 1 if on else 0))

    def timing(self):
        """(main kernel ms, fix-up kernel ms, timed launches) since the last call."""
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        self._check(self._L.te_get_timing(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def flag_counters(self):
        a = (C.c_uint32 * 5)()
        self._L.te_get_flag_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        self._check(self._L.te_get_flag_counters(self._h, a))
        return list(a)

    def escalation_stats(self):
        """(cells by escalation reason [16], cells by number of valid window cells [26]) of the last fused launch."""
        a, b = (C.c_uint32 * 16)(), (C.c_uint32 * 26)()
        self._L.te_get_escalation_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        self._check(self._L.te_get_escalation_stats(self._h, a, b))
        return list(a), list(b)

    def slope(self, g, critical, nz, out, memory):
        self._order_after_torch(memory)
        self._check(self._L.te_slope(self._h, C.byref(g), critical, _addr(nz), _addr(out), memory))

    def normals(self, g, p, elevation, nx, ny, nz, memory):
        self._order_after_torch(memory)
        self._check(self._L.te_normals(self._h, C.byref(g), C.byref(p), _addr(elevation), _addr(nx), _addr(ny), _addr(nz), memory))

    def step(self, g, p, elevation, out, memory):
        self._order_after_torch(memory)
        self._check(self._L.te_step(self._h, C.byref(g), C.byref(p), _addr(elevation), _addr(out), memory))

    def roughness(self, g, p, elevation, nx, ny, nz, out, memory):
        self._order_after_torch(memory)
        self._check(self._L.te_roughness(self._h, C.byref(g), C.byref(p), _addr(elevation), _addr(nx), _addr(ny), _addr(nz),
                                         _addr(out), memory))

    def chain(self, g, p, elevation, slope, step, roughness, traversability, memory, slab=None, nx=None, ny=None, nz=None):
        self._order_after_torch(memory)
        self._check(self._L.te_chain(self._h, C.byref(g), C.byref(slab) if slab is not None else None, C.byref(p),
                                     _addr(elevation), _addr(slope), _addr(step), _addr(roughness), _addr(traversability),
                                     _addr(nx), _addr(ny), _addr(nz), memory))

    def chain_batched(self, g, p, nmaps, elevation, slope, step, roughness, traversability, memory):
        self._order_after_torch(memory)
        self._check(self._L.te_chain_batched(self._h, C.byref(g), C.byref(p), nmaps, _addr(elevation), _addr(slope),
                                             _addr(step), _addr(roughness), _addr(traversability), memory))

    def footprint(self, g, fp, traversability, slope, step, elevation, out, memory, slab=None, slope_fp=None, step_fp=None,
                  roughness=None, roughness_fp=None):
        self._order_after_torch(memory)
        if roughness is None and roughness_fp is None:
            self._check(self._L.te_footprint(self._h, C.byref(g), C.byref(slab) if slab is not None else None, C.byref(fp),
                                             _addr(traversability), _addr(slope), _addr(step), _addr(elevation), _addr(out),
                                             _addr(slope_fp), _addr(step_fp), memory))
        else:
            self._check(self._L.te_footprint2(self._h, C.byref(g), C.byref(slab) if slab is not None else None, C.byref(fp),
                                              _addr(traversability), _addr(slope), _addr(step), _addr(roughness), _addr(elevation),
                                              _addr(out), _addr(slope_fp), _addr(step_fp), _addr(roughness_fp), memory))

    def footprint_polygon(self, g, fp, polygon_xy, yaw, traversability, slope, step, elevation, out_x, out_rot, memory, slab=None, roughness=None):
        """TraversabilityMap::traversabilityFootprint(yaw): layers traversability_x / traversability_rot for the footprint polygon."""
        pts = np.ascontiguousarray(polygon_xy, dtype=np.float64).reshape(-1, 2)
        self._L.te_footprint_polygon.argtypes = [C.c_void_p, C.POINTER(Geometry), C.c_void_p, C.POINTER(FootprintParams), C.c_int32, C.c_void_p,
                                                 C.c_double] + [C.c_void_p] * 7 + [C.c_int]
        self._check(self._L.te_footprint_polygon(self._h, C.byref(g), C.byref(slab) if slab is not None else None, C.byref(fp), len(pts),
                                                 pts.ctypes.data, float(yaw), _addr(traversability), _addr(slope), _addr(step), _addr(roughness),
                                                 _addr(elevation), _addr(out_x), _addr(out_rot), memory))

    def check_footprint_paths(self, g, footprint_layer, traversability_default, path_begin, poses_xy, robot_slope=None):
        """Host convenience: (is_safe uint8[npaths], traversability float64[npaths]); footprint_layer is a column-major host layer;
        robot_slope (optional layer) switches checkRobotInclination_ on."""
        f = np.asfortranarray(footprint_layer, dtype=np.float32)
        rs = np.asfortranarray(robot_slope, dtype=np.float32) if robot_slope is not None else None
        pb = np.ascontiguousarray(path_begin, dtype=np.int32)
        xy = np.ascontiguousarray(poses_xy, dtype=np.float64)
        n = len(pb) - 1
        safe, trav = np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.float64)
        self._L.te_check_footprint_paths2.argtypes = [C.c_void_p, C.POINTER(Geometry), C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_void_p,
                                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        self._check(self._L.te_check_footprint_paths2(self._h, C.byref(g), f.ctypes.data, rs.ctypes.data if rs is not None else None,
                                                      traversability_default, n, pb.ctypes.data, xy.ctypes.data, safe.ctypes.data,
                                                      trav.ctypes.data, MEM_HOST))
        return safe, trav

    # ---- multi-GPU halo (te_halo_pull and the IPC helpers around it)
    def ipc_export(self, device_ptr) -> bytes:
        h = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._check(self._L.te_ipc_export(_addr(device_ptr), h))
        return h.raw

    def ipc_open(self, handle: bytes) -> int:
        out = C.c_void_p()
        self._check(self._L.te_ipc_open(C.create_string_buffer(handle, IPC_HANDLE_BYTES), C.byref(out)))
        return out.value

    def ipc_close(self, ptr: int):
        self._check(self._L.te_ipc_close(ptr))

    def event_create_ipc(self):
        """(event, 64-byte handle) of an interprocess event recorded with event_record()."""
        ev, h = C.c_void_p(), C.create_string_buffer(64)
        self._check(self._L.te_event_create_ipc(self._h, C.byref(ev), h))
        return ev.value, h.raw

    def event_open_ipc(self, handle: bytes) -> int:
        ev = C.c_void_p()
        self._check(self._L.te_event_open_ipc(C.create_string_buffer(handle, 64), C.byref(ev)))
        return ev.value

    def event_record(self, event: int):
        self._check(self._L.te_event_record(self._h, event))

    def event_destroy(self, event: int):
        self._check(self._L.te_event_destroy(event))

    def halo_pull(self, g, slab, layer, left=None, right=None):
        self._order_after_torch(MEM_DEVICE)
        self._check(self._L.te_halo_pull(self._h, C.byref(g), C.byref(slab), _addr(layer),
                                         C.byref(left) if left is not None else None, C.byref(right) if right is not None else None))

    # Convenience for host numpy layers (column-major float32), used by tests.
    def chain_host(self, g, p, elevation, with_normals=False):
        e = np.asfortranarray(elevation, dtype=np.float32)
        assert e.shape == (g.rows, g.cols)
        o = {k: np.empty((g.rows, g.cols), np.float32, order="F") for k in ("slope", "step", "roughness", "traversability")}
        n = {k: np.empty((g.rows, g.cols), np.float32, order="F") for k in ("nx", "ny", "nz")} if with_normals else {}
        self.chain(g, p, e, o["slope"], o["step"], o["roughness"], o["traversability"], MEM_HOST, **n)
        o.update(n)
        return o
