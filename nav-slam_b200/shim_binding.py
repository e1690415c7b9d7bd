"""ctypes caller of the per-shape reference-ABI shim (libnavslam_shim_<R>x<C>.so).

The shim exports the reference's own symbols (headers/slam.h:22-28, utils/kdtree.h:21-30,
utils/pointcloud.h:55-57) with byte-identical signatures; this module calls the three `slam.h`
entry points exactly the way the reference's main.c does (src/main.c:253,309,317): `Pos` by value,
`PointCloud*` / `SLAM_attr*` as caller-owned memory of the reference's layout.  Used by bench.py
(`e2e_shim`) and the shim tests; it never touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build


class Pos(C.Structure):  # utils/pointcloud.h:32-35
    _fields_ = [(n, C.c_double) for n in ("x", "y", "z", "roll", "pitch", "yaw")]

    @classmethod
    def of(cls, v):
        return cls(*[float(t) for t in v])

    def arr(self):
        return np.array([self.x, self.y, self.z, self.roll, self.pitch, self.yaw])


class ShimSlam:
    """One SLAM_attr driven through init_slam / slam_localization / slam_mapping of the shim."""

    def __init__(self, rows: int, cols: int):
        path = _build.shim_path(rows, cols)
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run __graft_entry__.build()")
        self.rows, self.cols = rows, cols
        L = self.lib = C.CDLL(path)
        assert L.navslam_abi_rows() == rows and L.navslam_abi_cols() == cols
        for f in ("navslam_abi_sizeof_pointcloud", "navslam_abi_sizeof_slam_attr", "navslam_abi_offsetof_frame_count",
                  "navslam_abi_offsetof_error"):
            getattr(L, f).restype = C.c_size_t
        L.init_slam.argtypes = [C.c_void_p, Pos, C.c_void_p]
        L.slam_mapping.argtypes = [C.c_void_p, Pos, C.c_void_p]
        L.slam_localization.restype = Pos
        L.slam_localization.argtypes = [C.c_void_p, C.c_void_p, Pos, Pos]
        L.navslam_release.argtypes = [C.c_void_p]
        self.sizeof_pointcloud = L.navslam_abi_sizeof_pointcloud()
        self.sizeof_slam_attr = L.navslam_abi_sizeof_slam_attr()
        self.off_frame_count = L.navslam_abi_offsetof_frame_count()
        self.off_error = L.navslam_abi_offsetof_error()
        self.attr = np.zeros(self.sizeof_slam_attr, dtype=np.uint8)  # SLAM_attr (315 MB at 64x2048, lazily mapped)

    def pack_cloud(self, cloud, ts=0) -> np.ndarray:
        """PointCloud = {int ToF_timestamps; Point ToF_position[R][C]} with the points at byte offset 8."""
        cloud = np.ascontiguousarray(cloud, dtype=np.float64)
        buf = np.zeros(self.sizeof_pointcloud, dtype=np.uint8)
        buf[:4] = np.frombuffer(np.int32(ts).tobytes(), dtype=np.uint8)
        buf[8:] = np.frombuffer(cloud.tobytes(), dtype=np.uint8)
        return buf

    @property
    def frame_count(self) -> int:
        return int(self.attr[self.off_frame_count:self.off_frame_count + 4].view(np.int32)[0])

    @frame_count.setter
    def frame_count(self, v: int):
        self.attr[self.off_frame_count:self.off_frame_count + 4] = np.frombuffer(np.int32(v).tobytes(), dtype=np.uint8)

    @property
    def error(self) -> float:
        return float(self.attr[self.off_error:self.off_error + 8].view(np.float64)[0])

    def global_cloud(self, frame: int) -> np.ndarray:
        off = frame * self.sizeof_pointcloud + 8
        n = self.rows * self.cols * 24
        return self.attr[off:off + n].view(np.float64).reshape(self.rows, self.cols, 3)

    def init_slam(self, pos, packed_cloud):
        self.lib.init_slam(self.attr.ctypes.data, Pos.of(pos), packed_cloud.ctypes.data)

    def slam_localization(self, packed_cloud, pos_predict, pos_last) -> np.ndarray:
        return self.lib.slam_localization(self.attr.ctypes.data, packed_cloud.ctypes.data, Pos.of(pos_predict),
                                          Pos.of(pos_last)).arr()

    def slam_mapping(self, pos, packed_cloud):
        self.lib.slam_mapping(self.attr.ctypes.data, Pos.of(pos), packed_cloud.ctypes.data)

    def release(self):
        self.lib.navslam_release(self.attr.ctypes.data)
