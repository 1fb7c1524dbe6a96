"""rm_radar_b200 — B200-native detect + locate hot path of zmsbruce/rm_radar.

Python host mirror of the reference's public classes (`/root/reference/src/radar.h:15-18`):
`Detector`, `RobotDetector`, `Locator`, `Robot`, `Detection`, `Label` — same names, argument order,
defaults and error behaviour as `detector.h:87-93,173-184`, `locator.h:59-71`, `robot.h:32-164` —
over the C ABI in `include/rm_radar_b200.h`.  All computation happens in the CUDA library; there is
no CPU fallback (importing works without a GPU, constructing an object does not).
"""
from __future__ import annotations

import ctypes as C
import enum
import os
import time
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import CapacityError, RadarError  # noqa: F401

__all__ = ["Detector", "RobotDetector", "Locator", "Robot", "Detection", "Label", "RadarError", "Comm",
           "engine_path_for", "build_engine", "run_once", "run_once_records"]


class Label(enum.IntEnum):
    # enum Label — /root/reference/src/robot/robot.h:32-45 (same order as armor.onnx class names)
    BlueHero = 0
    BlueEngineer = 1
    BlueInfantryThree = 2
    BlueInfantryFour = 3
    BlueInfantryFive = 4
    RedHero = 5
    RedEngineer = 6
    RedInfantryThree = 7
    RedInfantryFour = 8
    RedInfantryFive = 9
    BlueSentry = 10
    RedSentry = 11


@dataclass
class Detection:
    # radar::Detection — detection.h:25-68
    x: float
    y: float
    width: float
    height: float
    label: float
    confidence: float

    def as_array(self):
        return np.array([self.x, self.y, self.width, self.height, self.label, self.confidence], np.float32)


@dataclass
class Robot:
    # radar::Robot — robot.h:53-164 (optional-valued fields are None when unset)
    rect: tuple | None = None
    label: int | None = None
    confidence: float | None = None
    armors: list | None = None
    location: tuple | None = None
    cluster: int | None = None
    cluster_points: int = 0
    track_state: int | None = None   # 0 tentative, 1 confirmed (robot.h:139-141)
    track_id: int | None = None

    def isTracked(self) -> bool:    # robot.h:81
        return self.track_state is not None

    def isDetected(self) -> bool:   # robot.h:65
        return self.armors is not None

    def isLocated(self) -> bool:    # robot.h:73
        return self.location is not None


def build_engine(onnx_path: str, engine_path: str, input_width: int = 640, input_height: int = 640) -> None:
    """ONNX → `.rmeng` plan through the library's own builder (csrc/engine.cu, `rmr_engine_build`); stands where the
    reference builds and caches a TensorRT engine (detector.cpp:177-243, 281-311).  Host only."""
    _lib.check(_lib.load().rmr_engine_build(os.fspath(onnx_path).encode(), os.fspath(engine_path).encode(),
                                            input_width, input_height))


def engine_path_for(path: str, input_width: int = 640, input_height: int = 640) -> str:
    """Resolve the reference's `engine_path` argument.  The reference loads `<x>.engine` or builds it
    from the sibling `<x>.onnx` (detector.cpp:74-99).  We accept `.rmeng`, `.engine` or `.onnx` and
    build `<x>.rmeng` from `<x>.onnx` when it does not exist yet (`rmr_engine_resolve`; the C++ constructors
    do the same on their own).  A read-only model directory falls back to a cache next to the package."""
    lib = _lib.load()
    buf = C.create_string_buffer(4096)
    rc = lib.rmr_engine_resolve(os.fspath(path).encode(), input_width, input_height, buf, len(buf))
    if rc == 0:
        return buf.value.decode()
    if rc == -1:
        _lib.check(rc)       # ValueError = std::invalid_argument, detector.cpp:80
    base = os.path.splitext(path)[0]
    onnx = base + ".onnx"
    if not os.path.exists(onnx):
        _lib.check(rc)
    cache = os.path.join(os.path.dirname(os.path.abspath(__file__)), "engines")
    os.makedirs(cache, exist_ok=True)
    eng = os.path.join(cache, os.path.basename(base) + ".rmeng")
    if not os.path.exists(eng):
        build_engine(onnx, eng, input_width, input_height)
    return eng


def _as_bgr(image: np.ndarray) -> np.ndarray:
    if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
        raise ValueError("image must be HxWx3 uint8 (BGR)")
    if not image.flags.c_contiguous and image.strides[1:] != (3, 1):
        image = np.ascontiguousarray(image)
    return image


def _dets(buf, n):
    return [Detection(*buf[i].astuple()) for i in range(n)]


class Detector:
    """radar::Detector — detector.h:84-134."""

    def __init__(self, engine_path, classes, image_size, max_batch_size, opt_batch_size=None, nms_thresh=0.65,
                 conf_thresh=0.25, input_width=640, input_height=640, input_name="images", input_channels=3,
                 opt_level=3, *, compat=True, device=0):
        if input_channels != 3:
            raise ValueError("input_channels must be 3")
        self._lib = _lib.load()
        self._h = C.c_void_p()
        w, h = image_size
        eng = engine_path_for(os.fspath(engine_path))
        _lib.check(self._lib.rmr_detector_create(C.byref(self._h), eng.encode(), classes, w, h, max_batch_size,
                                                 nms_thresh, conf_thresh, input_width, input_height, int(compat),
                                                 device))
        self.classes = classes
        self.max_batch_size = max_batch_size
        self.input_size = (input_width, input_height)
        self._cap = 1024   # kMaxOut of the library: every survivor comes back

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.rmr_detector_destroy(self._h)
            self._h.value = None

    def detect(self, images):
        """One HxWx3 BGR uint8 array → list[Detection]; a sequence of arrays → list[list[Detection]]."""
        if isinstance(images, np.ndarray):
            img = _as_bgr(images)
            out = (_lib.Detection * self._cap)()
            n = C.c_int()
            _lib.check(self._lib.rmr_detector_detect(self._h, img.ctypes.data, img.shape[1], img.shape[0],
                                                     img.strides[0], out, self._cap, C.byref(n)))
            return _dets(out, min(n.value, self._cap))
        imgs = [_as_bgr(i) for i in images]
        k = len(imgs)
        if k == 0:
            return []
        ptrs = (C.c_void_p * k)(*[i.ctypes.data for i in imgs])
        ws = (C.c_int * k)(*[i.shape[1] for i in imgs])
        hs = (C.c_int * k)(*[i.shape[0] for i in imgs])
        ss = (C.c_int * k)(*[i.strides[0] for i in imgs])
        out = (_lib.Detection * (self._cap * k))()
        counts = (C.c_int * k)()
        _lib.check(self._lib.rmr_detector_detect_batch(self._h, ptrs, ws, hs, ss, k, out, self._cap, counts))
        return [[Detection(*out[i * self._cap + j].astuple()) for j in range(min(counts[i], self._cap))]
                for i in range(k)]

    # -- inspection (tests) --
    def last_input(self, n=1):
        w, h = self.input_size
        out = np.empty((n, 3, h, w), np.float32)
        _lib.check(self._lib.rmr_detector_last_input(self._h, out.ctypes.data, n))
        return out

    def last_output(self, n=1):
        a = C.c_int()
        _lib.check(self._lib.rmr_detector_info(self._h, C.byref(a), None, None, None))
        out = np.empty((n, 4 + self.classes, a.value), np.float32)
        _lib.check(self._lib.rmr_detector_last_output(self._h, out.ctypes.data, n))
        return out

    def set_stream(self, cuda_stream: int):
        _lib.check(self._lib.rmr_detector_set_stream(self._h, C.c_void_p(cuda_stream)))

    def time_forward(self, batch: int, iters: int = 20) -> float:
        """ms per replay of the conv-stack graph at `batch` (CUDA events on the detector's stream)."""
        ms = C.c_float()
        _lib.check(self._lib.rmr_detector_time_forward(self._h, batch, iters, C.byref(ms)))
        return ms.value

    def profile_ops(self, batch: int = 1, iters: int = 20):
        """Per-op table of the engine plan: list of dicts (type, umma, shapes, flops, ms)."""
        cap = 512
        rows = np.zeros((cap, 12), np.float64)
        n = C.c_int()
        _lib.check(self._lib.rmr_detector_profile_ops(self._h, batch, iters,
                                                      rows.ctypes.data_as(C.POINTER(C.c_double)), cap, C.byref(n)))
        keys = ("type", "umma", "h_in", "w_in", "cin", "h_out", "w_out", "cout", "k", "stride", "flops", "ms")
        return [dict(zip(keys, rows[i].tolist())) for i in range(min(n.value, cap))]

    def plan_stats(self, batch: int = 1):
        n, u, l = C.c_int(), C.c_int(), C.c_int()
        _lib.check(self._lib.rmr_detector_plan_stats(self._h, batch, C.byref(n), C.byref(u), C.byref(l)))
        return dict(launches=n.value, umma_convs=u.value, graph_lanes=l.value)

    def info(self):
        a, c, k, f = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        _lib.check(self._lib.rmr_detector_info(self._h, C.byref(a), C.byref(c), C.byref(k), C.byref(f)))
        return dict(anchors=a.value, classes=c.value, kernel_launches=k.value, flops_per_image=f.value)


class _DetectorView(Detector):
    def __init__(self, lib, handle, classes, input_size):   # borrowed handle: never destroyed
        self._lib, self._h, self.classes, self.input_size, self._cap = lib, C.c_void_p(handle), classes, input_size, 1024

    def __del__(self):
        pass


def _robot_from_rec(r) -> Robot:
    robot = Robot()
    if r.has_rect:
        robot.rect = tuple(r.rect)
    if r.is_detected:
        robot.label = int(r.label)
        robot.confidence = float(r.confidence)
        robot.armors = [Detection(*r.armors[i].astuple()) for i in range(r.n_armors)]
    if r.is_located:
        robot.location = tuple(r.location)
        robot.cluster = int(r.cluster)
        robot.cluster_points = int(r.cluster_points)
    return robot


class RobotDetector:
    """radar::RobotDetector — detector.h:171-190."""

    def __init__(self, car_path, armor_path, image_size, armor_classes, max_cars, opt_cars, iou_thresh=0.75,
                 car_nms_thresh=0.65, car_conf_thresh=0.25, armor_nms_thresh=0.65, armor_conf_thresh=0.50,
                 input_width=640, input_height=640, input_name="images", input_channels=3, opt_level=5, *,
                 compat=True, device=0, frames=1):
        """frames > 1: throughput mode (BASELINE config[2]) — detect_frames / run_batch take that many images per call."""
        self._lib = _lib.load()
        self._h = C.c_void_p()
        w, h = image_size
        car = engine_path_for(os.fspath(car_path))
        armor = engine_path_for(os.fspath(armor_path))
        _lib.check(self._lib.rmr_robot_detector_create_batched(
            C.byref(self._h), car.encode(), armor.encode(), w, h, armor_classes, max_cars, iou_thresh,
            car_nms_thresh, car_conf_thresh, armor_nms_thresh, armor_conf_thresh, input_width, input_height,
            int(compat), device, frames))
        self.frames = frames
        self._frame_recs = (_lib.RobotRec * (max_cars * frames))() if frames > 1 else None
        self.max_cars = max_cars
        self.armor_classes = armor_classes
        self.input_size = (input_width, input_height)
        self._recs = (_lib.RobotRec * max_cars)()

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.rmr_robot_detector_destroy(self._h)
            self._h.value = None

    def detect(self, image: np.ndarray) -> list:
        """RobotDetector::detect(const cv::Mat&) — detector.cpp:413-455."""
        img = _as_bgr(image)
        n = C.c_int()
        _lib.check(self._lib.rmr_robot_detector_detect(self._h, img.ctypes.data, img.shape[1], img.shape[0],
                                                       img.strides[0], self._recs, self.max_cars, C.byref(n)))
        return [_robot_from_rec(self._recs[i]) for i in range(min(n.value, self.max_cars))]

    def detect_frames(self, images) -> list:
        """RobotDetector::detect on each of n same-sized frames, the networks batched over them -> list of list[Robot]."""
        batch = np.ascontiguousarray(np.stack([_as_bgr(im) for im in images]))
        n, h, w = batch.shape[:3]
        counts = (C.c_int * n)()
        recs = self._frame_recs if self._frame_recs is not None else (_lib.RobotRec * (self.max_cars * n))()
        _lib.check(self._lib.rmr_robot_detector_detect_frames(self._h, batch.ctypes.data, 0, n, w, h, w * 3, recs,
                                                              self.max_cars, counts))
        return [[_robot_from_rec(recs[f * self.max_cars + i]) for i in range(min(counts[f], self.max_cars))] for f in range(n)]

    def detect_jpeg(self, decoder: "JpegDecoder", file_bytes: bytes) -> list:
        """cv::imread + RobotDetector::detect (samples/main.cpp:24-40, detector.cpp:413-455): the JPEG is decoded on the
        device and the frame never crosses PCIe."""
        n = C.c_int()
        _lib.check(self._lib.rmr_robot_detector_detect_jpeg(self._h, decoder._h, file_bytes, len(file_bytes), self._recs,
                                                            self.max_cars, C.byref(n)))
        return [_robot_from_rec(self._recs[i]) for i in range(min(n.value, self.max_cars))]

    def detect_records(self, ptr: int, width: int, height: int, stride: int, device_ptr: bool):
        """Raw-record variant used by the bench: returns (ctypes RobotRec array, count)."""
        n = C.c_int()
        fn = self._lib.rmr_robot_detector_detect_device if device_ptr else self._lib.rmr_robot_detector_detect
        _lib.check(fn(self._h, C.c_void_p(ptr), width, height, stride, self._recs, self.max_cars, C.byref(n)))
        return self._recs, min(n.value, self.max_cars)

    def last_cars(self):
        out = (_lib.Detection * 64)()
        n = C.c_int()
        _lib.check(self._lib.rmr_robot_detector_last_cars(self._h, out, 64, C.byref(n)))
        return _dets(out, min(n.value, 64))

    def last_armors(self, car_index):
        out = (_lib.Detection * 64)()
        n = C.c_int()
        _lib.check(self._lib.rmr_robot_detector_last_armors(self._h, car_index, out, 64, C.byref(n)))
        return _dets(out, min(n.value, 64))

    def last_stats(self):
        k, f, c = C.c_int(), C.c_double(), C.c_int()
        _lib.check(self._lib.rmr_robot_detector_last_stats(self._h, C.byref(k), C.byref(f), C.byref(c)))
        return dict(kernel_launches=k.value, conv_flops=f.value, n_cars=c.value)

    def last_timing(self):
        """Device ms of the car / armor network replays of the last detect call."""
        a, b = C.c_float(), C.c_float()
        _lib.check(self._lib.rmr_robot_detector_last_timing(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def car_detector(self):
        return _DetectorView(self._lib, self._lib.rmr_robot_detector_car(self._h), 1, self.input_size)

    def armor_detector(self):
        return _DetectorView(self._lib, self._lib.rmr_robot_detector_armor(self._h), self.armor_classes,
                             self.input_size)

    def set_stream(self, cuda_stream: int):
        _lib.check(self._lib.rmr_robot_detector_set_stream(self._h, C.c_void_p(cuda_stream)))


class Locator:
    """radar::Locator — locator.h:53-71."""

    def __init__(self, image_width, image_height, intrinsic, lidar_to_camera, world_to_camera, zoom_factor=0.5,
                 queue_size=3, min_depth_diff=500, max_depth_diff=4000, cluster_tolerance=400, min_cluster_size=8,
                 max_cluster_size=1000, max_distance=29300, *, device=0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        K = np.ascontiguousarray(np.asarray(intrinsic, np.float32).reshape(9))
        L = np.ascontiguousarray(np.asarray(lidar_to_camera, np.float32).reshape(16))
        W = np.ascontiguousarray(np.asarray(world_to_camera, np.float32).reshape(16))
        fp = C.POINTER(C.c_float)
        _lib.check(self._lib.rmr_locator_create(
            C.byref(self._h), image_width, image_height, K.ctypes.data_as(fp), L.ctypes.data_as(fp),
            W.ctypes.data_as(fp), zoom_factor, queue_size, min_depth_diff, max_depth_diff, cluster_tolerance,
            min_cluster_size, max_cluster_size, max_distance, device))
        w, h = C.c_int(), C.c_int()
        _lib.check(self._lib.rmr_locator_image_size(self._h, C.byref(w), C.byref(h)))
        self.image_size_zoomed = (w.value, h.value)

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.rmr_locator_destroy(self._h)
            self._h.value = None

    def update(self, cloud):
        """Locator::update — locate.cpp:158-220.  cloud: [n,3] or [n,4] float32 (PointXYZ) or None."""
        if cloud is None or len(cloud) == 0:
            _lib.check(self._lib.rmr_locator_update(self._h, None, 0, 16))
            return
        pts = np.asarray(cloud)
        if pts.dtype != np.float32 or pts.ndim != 2 or pts.shape[1] not in (3, 4) or not pts.flags.c_contiguous:
            pts = np.ascontiguousarray(pts[:, :3], np.float32)
        _lib.check(self._lib.rmr_locator_update(self._h, pts.ctypes.data, pts.shape[0], pts.strides[0]))

    def update_pcd(self, file_bytes: bytes) -> int:
        """Locator::update fed from a PCD v0.7 file image parsed on the device; returns the point count."""
        n = C.c_int()
        _lib.check(self._lib.rmr_locator_update_pcd(self._h, file_bytes, len(file_bytes), C.byref(n)))
        return n.value

    def update_device(self, dev_ptr: int, n_points: int, stride_bytes: int):
        _lib.check(self._lib.rmr_locator_update_device(self._h, C.c_void_p(dev_ptr), n_points, stride_bytes))

    def cluster(self):
        """Locator::cluster — locate.cpp:231-264."""
        _lib.check(self._lib.rmr_locator_cluster(self._h))

    def search(self, robots):
        """Locator::search(std::vector<Robot>&) — locate.cpp:276-326: sets `location` in place."""
        n = len(robots)
        if n == 0:
            return
        recs = (_lib.RobotRec * n)()
        for i, r in enumerate(robots):
            if r.rect is not None:
                recs[i].rect = (C.c_float * 4)(*r.rect)
                recs[i].has_rect = 1
        _lib.check(self._lib.rmr_locator_search(self._h, recs, n))
        for i, r in enumerate(robots):
            if recs[i].is_located:
                r.location = tuple(recs[i].location)
                r.cluster = int(recs[i].cluster)
                r.cluster_points = int(recs[i].cluster_points)

    def search_records(self, recs, n):
        _lib.check(self._lib.rmr_locator_search(self._h, recs, n))

    # -- inspection (tests) --
    def image(self, which: str) -> np.ndarray:
        idx = {"depth": 0, "background": 1, "diff": 2, "labels": 3}[which]
        w, h = self.image_size_zoomed
        out = np.empty((h, w), np.int32 if idx == 3 else np.float32)
        _lib.check(self._lib.rmr_locator_read_image(self._h, idx, out.ctypes.data))
        return out

    def save_background(self) -> np.ndarray:
        """The running-max background depth image (the only long-lived Locator state)."""
        return self.image("background")

    def load_background(self, image: np.ndarray):
        img = np.ascontiguousarray(image, np.float32)
        _lib.check(self._lib.rmr_locator_load_background(self._h, img.ctypes.data, img.shape[1], img.shape[0]))

    def stats(self):
        f, c = C.c_int(), C.c_int()
        _lib.check(self._lib.rmr_locator_stats(self._h, C.byref(f), C.byref(c)))
        return dict(foreground=f.value, clusters=c.value)

    def foreground(self):
        n = self.stats()["foreground"]
        out = np.empty((max(n, 1), 4), np.float32)
        _lib.check(self._lib.rmr_locator_read_foreground(self._h, out.ctypes.data, n))
        return out[:n, :3].copy(), out[:n, 3].view(np.int32).copy()

    def set_stream(self, cuda_stream: int):
        _lib.check(self._lib.rmr_locator_set_stream(self._h, C.c_void_p(cuda_stream)))


def run_once_records(detector: "RobotDetector", locator: "Locator", frame_ptr: int, frame_on_device: bool, width: int,
                     height: int, stride: int, cloud_ptr: int, cloud_on_device: bool, n_points: int, point_stride: int):
    """SampleRadar::runOnce (sample_radar.h:106-127) on raw pointers: returns (ctypes RobotRec array, count)."""
    n = C.c_int()
    _lib.check(detector._lib.rmr_run_once(detector._h, locator._h, C.c_void_p(frame_ptr), int(frame_on_device), width,
                                          height, stride, C.c_void_p(cloud_ptr), int(cloud_on_device), n_points,
                                          point_stride, detector._recs, detector.max_cars, C.byref(n)))
    return detector._recs, min(n.value, detector.max_cars)


def run_batch_records(detector: "RobotDetector", locators, frames_ptr: int, frames_on_device: bool, n_frames: int, width: int,
                      height: int, stride: int, clouds_ptr: int, clouds_on_device: bool, n_points: int, point_stride: int):
    """n camera + LiDAR streams at once (throughput mode, rmr_run_batch): returns (RobotRec array [n][max_cars], counts)."""
    counts = (C.c_int * n_frames)()
    handles = (C.c_void_p * n_frames)(*[loc._h for loc in locators])
    _lib.check(detector._lib.rmr_run_batch(detector._h, handles, n_frames, C.c_void_p(frames_ptr), int(frames_on_device), width,
                                           height, stride, C.c_void_p(clouds_ptr), int(clouds_on_device), n_points,
                                           point_stride, detector._frame_recs, detector.max_cars, counts))
    return detector._frame_recs, counts


def run_batch(detector: "RobotDetector", locators, images, clouds) -> list:
    """run_once for n streams at once on host arrays -> list of list[Robot] (frame i with locators[i])."""
    batch = np.ascontiguousarray(np.stack([_as_bgr(im) for im in images]))
    pts = np.ascontiguousarray(np.stack([np.asarray(c, np.float32)[:, :3] for c in clouds]))
    n, h, w = batch.shape[:3]
    recs, counts = run_batch_records(detector, locators, batch.ctypes.data, False, n, w, h, w * 3, pts.ctypes.data, False,
                                     pts.shape[1], 12)
    return [[_robot_from_rec(recs[f * detector.max_cars + i]) for i in range(min(counts[f], detector.max_cars))] for f in range(n)]


def run_once(detector: "RobotDetector", locator: "Locator", image: np.ndarray, cloud, tracker: "Tracker | None" = None,
             timestamp_ns: int | None = None) -> list:
    """One frame of the whole path on host arrays: detect + update + cluster + search (+ Tracker::update when a
    tracker is given, sample_radar.h:121-123) -> list[Robot]."""
    img = _as_bgr(image)
    pts = np.ascontiguousarray(np.asarray(cloud, np.float32)[:, :3]) if cloud is not None and len(cloud) else None
    recs, n = run_once_records(detector, locator, img.ctypes.data, False, img.shape[1], img.shape[0], img.strides[0],
                               pts.ctypes.data if pts is not None else 0, False, len(pts) if pts is not None else 0, 12)
    if tracker is None:
        return [_robot_from_rec(recs[i]) for i in range(n)]
    state, tid = (C.c_int32 * max(n, 1))(), (C.c_int32 * max(n, 1))()
    tracker.update_records(recs, n, timestamp_ns if timestamp_ns is not None else time.monotonic_ns(), state, tid)
    robots = [_robot_from_rec(recs[i]) for i in range(n)]
    for i, r in enumerate(robots):
        if state[i] >= 0:
            r.track_state, r.track_id = int(state[i]), int(tid[i])
            r.label = int(recs[i].label)        # a tracked, undetected robot carries the track's label
    return robots


def auction(values, max_iter: int = 100) -> list:
    """radar::track::auction — auction.h:33-126: value matrix [agents, tasks] -> task per agent (-1 = unmatched)."""
    lib = _lib.load()
    v = np.ascontiguousarray(values, np.float32)
    n_agents = v.shape[0]
    n_tasks = v.shape[1] if v.ndim == 2 else 0
    out = (C.c_int32 * max(n_agents, 1))()
    _lib.check(lib.rmr_auction(v.ctypes.data, n_agents, n_tasks, max_iter, out))
    return [int(out[i]) for i in range(n_agents)]


class Tracker:
    """radar::Tracker — tracker.h:23-53, tracker.cpp:47-220.  Host code inside the same library."""

    def __init__(self, observation_noise, class_num, init_thresh=4, miss_thresh=10, max_acceleration=2.0,
                 acceleration_correlation_time=1.0, distance_weight=0.40, feature_weight=0.60, max_iter=100,
                 distance_thresh=0.8):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        noise = (C.c_float * 3)(*[float(x) for x in observation_noise])
        _lib.check(self._lib.rmr_tracker_create(C.byref(self._h), noise, class_num, init_thresh, miss_thresh, max_acceleration,
                                                acceleration_correlation_time, distance_weight, feature_weight, max_iter,
                                                distance_thresh))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.rmr_tracker_destroy(self._h)
            self._h.value = None

    def update(self, robots: list, timestamp_ns: int):
        """Tracker::update(std::vector<Robot>&, time_point): robots are updated in place (Robot::setTrack)."""
        n = len(robots)
        recs = (_lib.RobotRec * max(n, 1))()
        for i, r in enumerate(robots):
            rec = recs[i]
            rec.label = -1 if r.label is None else int(r.label)
            if r.rect is not None:
                rec.rect = (C.c_float * 4)(*r.rect)
                rec.has_rect = 1
            if r.armors is not None:
                rec.is_detected = 1
                rec.confidence = float(r.confidence or 0.0)
                rec.n_armors = min(len(r.armors), _lib.MAX_ARMORS)
                for k in range(rec.n_armors):
                    a = r.armors[k]
                    rec.armors[k] = _lib.Detection(a.x, a.y, a.width, a.height, a.label, a.confidence)
            if r.location is not None:
                rec.is_located = 1
                rec.location = (C.c_float * 3)(*r.location)
        state = (C.c_int32 * max(n, 1))()
        tid = (C.c_int32 * max(n, 1))()
        self.update_records(recs, n, timestamp_ns, state, tid)
        for i, r in enumerate(robots):
            if state[i] >= 0:
                r.track_state, r.track_id = int(state[i]), int(tid[i])
                r.label = int(recs[i].label)
                r.location = tuple(recs[i].location)

    def update_records(self, recs, n: int, timestamp_ns: int, state=None, tid=None):
        """Raw-record variant: the array `rmr_run_once` filled goes straight in."""
        _lib.check(self._lib.rmr_tracker_update(self._h, recs, n, int(timestamp_ns), state, tid))

    def tracks(self) -> list:
        n = C.c_int()
        out = (_lib.TrackRec * 64)()
        _lib.check(self._lib.rmr_tracker_tracks(self._h, out, 64, C.byref(n)))
        return [dict(id=t.id, label=t.label, state=t.state, init_count=t.init_count, miss_count=t.miss_count,
                     location=tuple(t.location), filter_state=tuple(t.filter_state)) for t in out[:min(n.value, 64)]]


def jpeg_info(file_bytes: bytes) -> dict:
    """Header fields of a JPEG file image (host only)."""
    lib = _lib.load()
    v = [C.c_int() for _ in range(6)]
    _lib.check(lib.rmr_jpeg_info(file_bytes, len(file_bytes), *[C.byref(x) for x in v]))
    return dict(zip(("width", "height", "components", "h_samp", "v_samp", "restart_interval"), (x.value for x in v)))


class JpegDecoder:
    """Device-side cv::imread for baseline JPEG (samples/main.cpp:24-40); bit-exact with libjpeg-turbo's defaults."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._lib.rmr_jpeg_decoder_create(C.byref(self._h), device))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.rmr_jpeg_decoder_destroy(self._h)
            self._h.value = None

    def set_stream(self, cuda_stream: int):
        _lib.check(self._lib.rmr_jpeg_decoder_set_stream(self._h, C.c_void_p(cuda_stream)))

    def decode(self, file_bytes: bytes) -> np.ndarray:
        """-> BGR uint8 [H, W, 3] on the host (what cv2.imdecode / cv::imread returns)."""
        meta = jpeg_info(file_bytes)
        out = np.empty((meta["height"], meta["width"], 3), np.uint8)
        w, h = C.c_int(), C.c_int()
        _lib.check(self._lib.rmr_jpeg_decode(self._h, file_bytes, len(file_bytes), out.ctypes.data, out.nbytes,
                                             C.byref(w), C.byref(h)))
        return out

    def decode_device(self, file_bytes: bytes, dev_ptr: int = 0, stride: int = 0):
        """Asynchronous decode into device memory; returns (device pointer, width, height)."""
        frame = C.c_void_p()
        w, h = C.c_int(), C.c_int()
        _lib.check(self._lib.rmr_jpeg_decode_device(self._h, file_bytes, len(file_bytes), C.c_void_p(dev_ptr or None),
                                                    stride, C.byref(frame), C.byref(w), C.byref(h)))
        return frame.value, w.value, h.value

    def status(self) -> dict:
        st, rounds, dec, launches, up = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        _lib.check(self._lib.rmr_jpeg_decoder_status(self._h, C.byref(st), C.byref(rounds), C.byref(dec), C.byref(launches),
                                                     C.byref(up)))
        return dict(status=st.value, sync_rounds=rounds.value, loop_decodes=dec.value, kernel_launches=launches.value,
                    upload_bytes=up.value)

    def profile(self, file_bytes: bytes) -> dict:
        """Device ms per stage of one decode (CUDA events between the launches)."""
        ms = (C.c_float * 13)()
        _lib.check(self._lib.rmr_jpeg_decoder_profile(self._h, file_bytes, len(file_bytes), ms))
        names = ("upload", "clear", "unstuff", "entropy", "dc_scan", "idct", "colour", "entropy.pass0", "entropy.pass1",
                 "entropy.chase", "entropy.verify", "entropy.scan", "entropy.write")
        return dict(zip(names, (float(x) for x in ms)))

    def coefficients(self, n_blocks: int) -> np.ndarray:
        """Quantised coefficient blocks of the last decode, int16 [n_blocks, 64] (scan order x natural order)."""
        out = np.empty((n_blocks, 64), np.int16)
        n = C.c_long()
        _lib.check(self._lib.rmr_jpeg_decoder_read_coefficients(self._h, out.ctypes.data, n_blocks, C.byref(n)))
        return out[:n.value]


def pcd_parse(file_bytes: bytes, capacity: int = 1 << 21, device: int = 0) -> np.ndarray:
    """PCD v0.7 file image -> [n, 3] float32, parsed on the device (pcl::io::loadPCDFile stand-in)."""
    lib = _lib.load()
    out = np.empty((capacity, 3), np.float32)
    n = C.c_int()
    _lib.check(lib.rmr_pcd_parse(file_bytes, len(file_bytes), out.ctypes.data, capacity, C.byref(n), device))
    return out[:n.value].copy()


def conv_selftest(n, h, w, cin, cout, k, stride, act=1, residual=0, out_f32=0, seed=0, iters=0):
    """tcgen05 conv vs the CUDA-core checker on random data; returns (max_abs_diff, max_ref, ms)."""
    lib = _lib.load()
    d, r, ms = C.c_float(), C.c_float(), C.c_float()
    _lib.check(lib.rmr_conv_selftest(n, h, w, cin, cout, k, stride, act, residual, out_f32, seed, iters,
                                     C.byref(d), C.byref(r), C.byref(ms)))
    return d.value, r.value, ms.value


class Comm:
    """The multi-GPU exchange inside the library (rmr_comm_*): one NCCL all-gather per step of every rank's block of
    robot records [max_robots, 8] = [valid, label, confidence, is_located, x, y, z, rect area]."""

    ID_BYTES, RECORD_FLOATS = 128, 8

    @staticmethod
    def _nccl_first():
        """The library resolves NCCL at run time (`dlopen libnccl.so.2`) and takes the instance the process already has.
        torch bundles a newer NCCL under the same soname: if the system one were loaded first, a later `import torch`
        in the same process would fail to resolve its symbols.  So, where torch is installed, it loads its NCCL first."""
        try:
            import torch  # noqa: F401
        except ImportError:
            pass

    @staticmethod
    def unique_id() -> bytes:
        Comm._nccl_first()
        buf = (C.c_uint8 * Comm.ID_BYTES)()
        _lib.check(_lib.load().rmr_comm_unique_id(buf))
        return bytes(buf)

    def __init__(self, unique_id: bytes, rank: int, world: int, device: int = 0, max_robots: int = 20):
        Comm._nccl_first()
        self._lib = _lib.load()
        self._h = C.c_void_p()
        buf = (C.c_uint8 * Comm.ID_BYTES).from_buffer_copy(unique_id)
        _lib.check(self._lib.rmr_comm_create(C.byref(self._h), buf, rank, world, device, max_robots))
        self.rank, self.world, self.max_robots = rank, world, max_robots
        self._out = np.zeros((world, max_robots, Comm.RECORD_FLOATS), np.float32)

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.rmr_comm_destroy(self._h)
            self._h.value = None

    def close(self):
        """Orderly shutdown; every rank calls it at the same point of the program."""
        if getattr(self, "_h", None) and self._h.value:
            _lib.check(self._lib.rmr_comm_close(self._h))

    def publish(self, recs, n: int, after_stream: int = 0):
        """recs: the ctypes RobotRec array rmr_run_once filled; returns at once (the exchange runs on its own stream)."""
        _lib.check(self._lib.rmr_comm_publish(self._h, recs, n, C.c_void_p(after_stream)))

    def collect(self) -> np.ndarray:
        """[world, max_robots, 8] of the last publish (waits for it)."""
        _lib.check(self._lib.rmr_comm_collect(self._h, self._out.ctypes.data))
        return self._out

    @staticmethod
    def pack(recs, n: int, max_robots: int) -> np.ndarray:
        out = np.zeros((max_robots, Comm.RECORD_FLOATS), np.float32)
        _lib.check(_lib.load().rmr_comm_pack(recs, n, max_robots, out.ctypes.data))
        return out


def postprocess_selftest(candidates: np.ndarray, nms_thresh: float) -> np.ndarray:
    """NMS + restore kernel on caller-supplied candidates [n, 6] (tests only) -> surviving rows [m, 6] in anchor order."""
    c = np.ascontiguousarray(candidates, np.float32).reshape(-1, 6)
    out = np.zeros((max(len(c), 1), 6), np.float32)
    n = C.c_int()
    _lib.check(_lib.load().rmr_postprocess_selftest(c.ctypes.data, len(c), nms_thresh, out.ctypes.data, len(out), C.byref(n)))
    return out[:min(n.value, len(out))]


def conv_timeline(n, h, w, cin, cout, k, stride, max_ctas=4096):
    """Per-CTA clock64 timeline of one tcgen05 conv launch: int64 array [ctas, 64] (profiling aid)."""
    lib = _lib.load()
    out = np.zeros((max_ctas, 64), np.int64)
    nc = C.c_int()
    _lib.check(lib.rmr_conv_timeline(n, h, w, cin, cout, k, stride, out.ctypes.data, max_ctas, C.byref(nc)))
    return out[:min(abs(nc.value), max_ctas)]   # conv2.cu launches report a negative count (other slot map)
