"""ORACLE (test/bench infrastructure only): ctypes wrapper of oracle/locate_ref.cpp, the C++ port of
the reference's CPU locate path used as the timed CPU baseline and as a second checker."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "build", "liblocate_ref.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            subprocess.check_call(["make", "-C", HERE])
        lib = C.CDLL(SO)
        fp = C.POINTER(C.c_float)
        lib.locref_create.restype = C.c_void_p
        lib.locref_create.argtypes = [C.c_int, C.c_int, fp, fp, fp, C.c_float, C.c_int, C.c_float, C.c_float,
                                      C.c_float, C.c_int, C.c_int, C.c_float, C.c_int]
        lib.locref_destroy.argtypes = [C.c_void_p]
        lib.locref_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.locref_cluster.argtypes = [C.c_void_p]
        lib.locref_search.argtypes = [C.c_void_p, fp, fp, C.POINTER(C.c_int)]
        lib.locref_search_many.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.locref_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.locref_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.locref_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib = lib
    return _lib


class LocatorRef:
    def __init__(self, image_width, image_height, intrinsic, lidar_to_camera, world_to_camera, zoom_factor=0.5,
                 queue_size=3, min_depth_diff=500, max_depth_diff=4000, cluster_tolerance=400, min_cluster_size=8,
                 max_cluster_size=1000, max_distance=29300, threads=0):
        lib = load()
        fp = C.POINTER(C.c_float)
        K = np.ascontiguousarray(np.asarray(intrinsic, np.float32).reshape(9))
        L = np.ascontiguousarray(np.asarray(lidar_to_camera, np.float32).reshape(16))
        W = np.ascontiguousarray(np.asarray(world_to_camera, np.float32).reshape(16))
        self._h = lib.locref_create(image_width, image_height, K.ctypes.data_as(fp), L.ctypes.data_as(fp),
                                    W.ctypes.data_as(fp), zoom_factor, queue_size, min_depth_diff, max_depth_diff,
                                    cluster_tolerance, min_cluster_size, max_cluster_size, max_distance, threads)
        if not self._h:
            raise ValueError("singular calibration")
        w, h = C.c_int(), C.c_int()
        lib.locref_size(self._h, C.byref(w), C.byref(h))
        self.Wz, self.Hz = w.value, h.value

    def __del__(self):
        if getattr(self, "_h", None):
            load().locref_destroy(self._h)
            self._h = None

    def update(self, cloud):
        if cloud is None or len(cloud) == 0:
            load().locref_update(self._h, None, 0, 3)
            return
        pts = np.ascontiguousarray(cloud, np.float32)
        load().locref_update(self._h, pts.ctypes.data, pts.shape[0], pts.shape[1])

    def cluster(self):
        load().locref_cluster(self._h)

    def search(self, rects):
        r = np.ascontiguousarray(rects, np.float32).reshape(-1, 4)
        xyz = np.zeros((len(r), 3), np.float32)
        loc = np.zeros(len(r), np.int32)
        load().locref_search_many(self._h, r.ctypes.data, len(r), xyz.ctypes.data, loc.ctypes.data)
        return [tuple(xyz[i]) if loc[i] else None for i in range(len(r))]

    def image(self, which):
        idx = {"depth": 0, "background": 1, "diff": 2, "labels": 3}[which]
        out = np.empty((self.Hz, self.Wz), np.int32 if idx == 3 else np.float32)
        load().locref_read(self._h, idx, out.ctypes.data)
        return out

    def counts(self):
        a, b = C.c_int(), C.c_int()
        load().locref_counts(self._h, C.byref(a), C.byref(b))
        return a.value, b.value
