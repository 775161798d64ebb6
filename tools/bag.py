"""Minimal rosbag-v2 reader for the reference's only data fixture.

The fixture `traversability_estimation/maps/elevation_map.bag` of the reference
(installed by `traversability_estimation/CMakeLists.txt:132-134`, loaded by the
`load_elevation_map` service `TraversabilityEstimation.cpp:125-152`) holds one
`grid_map_msgs/GridMap` message with the chain input (`elevation`) AND the chain
outputs, so it is the golden vector of the filter chain.

Only what that file needs is implemented: uncompressed chunks, one connection,
`grid_map_msgs/GridMap` deserialisation.  No ROS is required.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np


@dataclass
class GridMapMessage:
    frame_id: str
    stamp: tuple
    resolution: float
    length_x: float
    length_y: float
    pose: tuple
    layers: list
    basic_layers: list
    rows: int
    cols: int
    outer_start_index: int
    inner_start_index: int
    # layer name -> float32 array of shape (rows, cols), Fortran (column-major) order,
    # i.e. `a[i, j]` is `data[j * rows + i]` exactly like Eigen::MatrixXf.
    data: dict = field(default_factory=dict)


def _records(buf: bytes, pos: int, end: int):
    while pos < end:
        (hlen,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        hend = pos + hlen
        header = {}
        while pos < hend:
            (flen,) = struct.unpack_from("<I", buf, pos)
            pos += 4
            fieldb = buf[pos:pos + flen]
            pos += flen
            k, _, v = fieldb.partition(b"=")
            header[k.decode()] = v
        (dlen,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        yield header, buf[pos:pos + dlen]
        pos += dlen


class _Cursor:
    def __init__(self, b: bytes):
        self.b = b
        self.p = 0

    def take(self, fmt: str):
        v = struct.unpack_from("<" + fmt, self.b, self.p)
        self.p += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def string(self) -> str:
        n = self.take("I")
        s = self.b[self.p:self.p + n].decode()
        self.p += n
        return s

    def strings(self) -> list:
        return [self.string() for _ in range(self.take("I"))]


def _parse_gridmap(msg: bytes) -> GridMapMessage:
    c = _Cursor(msg)
    c.take("I")  # seq
    sec, nsec = c.take("II")
    frame = c.string()
    res, lx, ly = c.take("ddd")
    pose = c.take("7d")
    layers = c.strings()
    basic = c.strings()
    narr = c.take("I")
    arrays = []
    for _ in range(narr):
        ndim = c.take("I")
        dims = []
        for _ in range(ndim):
            label = c.string()
            size, stride = c.take("II")
            dims.append((label, size, stride))
        c.take("I")  # data_offset
        n = c.take("I")
        a = np.frombuffer(c.b, dtype="<f4", count=n, offset=c.p).copy()
        c.p += 4 * n
        arrays.append((dims, a))
    outer, inner = c.take("HH")
    dims0 = arrays[0][0]
    assert dims0[0][0] == "column_index" and dims0[1][0] == "row_index", dims0
    cols, rows = dims0[0][1], dims0[1][1]
    out = GridMapMessage(frame, (sec, nsec), res, lx, ly, pose, layers, basic, rows, cols, outer, inner)
    for name, (dims, a) in zip(layers, arrays):
        assert dims[0][1] == cols and dims[1][1] == rows
        out.data[name] = np.asfortranarray(a.reshape((cols, rows)).T)
    return out


def read_gridmap_bag(path: str) -> GridMapMessage:
    """Return the first grid_map_msgs/GridMap message in an uncompressed v2 bag."""
    with open(path, "rb") as f:
        buf = f.read()
    magic = b"#ROSBAG V2.0\n"
    assert buf.startswith(magic), "not a rosbag v2 file"
    for header, data in _records(buf, len(magic), len(buf)):
        op = header.get("op", b"\xff")[0]
        if op == 5:  # chunk
            assert header["compression"] == b"none", "compressed chunks unsupported"
            for h2, d2 in _records(data, 0, len(data)):
                if h2["op"][0] == 2:  # message data
                    return _parse_gridmap(d2)
    raise ValueError("no message record found")


if __name__ == "__main__":
    import sys
    import zlib

    m = read_gridmap_bag(sys.argv[1])
    print(m.frame_id, m.stamp, m.resolution, m.length_x, m.length_y, m.rows, m.cols, m.pose)
    for k, v in m.data.items():
        raw = np.ascontiguousarray(v.T).tobytes()
        print(f"{k:28s} crc32={zlib.crc32(raw):08x} sum={np.nansum(v.astype(np.float64)):.9f} "
              f"nan={int(np.isnan(v).sum())}")
