"""CPU-only: the C-ABI library loads and exports every symbol include/rm_radar_b200.h declares, and
the Python host mirror keeps the reference's constructor signatures.  No compute calls."""
import ctypes
import inspect
import os
import re

import pytest

from rm_radar_b200 import _lib
import rm_radar_b200 as rr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rm_radar_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rmr_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.SYMBOLS)


@pytest.mark.skipif(not os.path.exists(_lib.LIB_PATH), reason="library not built")
def test_library_exports_every_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), s


@pytest.mark.skipif(not os.path.exists(_lib.LIB_PATH), reason="library not built")
def test_struct_layouts():
    # Detection is the reference's 6-float POD (detection.h:27-28 static_asserts)
    assert ctypes.sizeof(_lib.Detection) == 24
    assert ctypes.sizeof(_lib.RobotRec) == 16 + 4 * 5 + 24 * 16 + 4 + 12 + 8


def test_reference_signatures():
    # detector.h:87-93
    p = list(inspect.signature(rr.Detector.__init__).parameters)
    assert p[1:13] == ["engine_path", "classes", "image_size", "max_batch_size", "opt_batch_size", "nms_thresh",
                       "conf_thresh", "input_width", "input_height", "input_name", "input_channels", "opt_level"]
    d = inspect.signature(rr.Detector.__init__).parameters
    assert d["nms_thresh"].default == 0.65 and d["conf_thresh"].default == 0.25 and d["input_width"].default == 640
    # detector.h:173-180
    r = inspect.signature(rr.RobotDetector.__init__).parameters
    assert list(r)[1:7] == ["car_path", "armor_path", "image_size", "armor_classes", "max_cars", "opt_cars"]
    assert r["iou_thresh"].default == 0.75 and r["armor_conf_thresh"].default == 0.50
    # locator.h:59-65
    l = inspect.signature(rr.Locator.__init__).parameters
    assert [l[k].default for k in ("zoom_factor", "queue_size", "min_depth_diff", "max_depth_diff",
                                   "cluster_tolerance", "min_cluster_size", "max_cluster_size", "max_distance")] == \
        [0.5, 3, 500, 4000, 400, 8, 1000, 29300]
    assert [x.name for x in rr.Label][:3] == ["BlueHero", "BlueEngineer", "BlueInfantryThree"] and len(rr.Label) == 12


@pytest.mark.skipif(not os.path.exists(_lib.LIB_PATH), reason="library not built")
def test_no_cpu_fallback_errors_loudly():
    """Without a CUDA device (this container) construction must fail, not silently compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    with pytest.raises((rr.RadarError, ValueError)):
        rr.Locator(640, 480, np.eye(3), np.eye(4), np.eye(4))
    with pytest.raises(ValueError):
        rr.Detector("/nonexistent/model.engine", 1, (640, 480), 1)


def test_headers_compile_standalone(tmp_path):
    """rm_radar_b200.h is plain C (what cgo / JNI / ctypes bind); radar.hpp is C++20 and needs neither
    OpenCV nor PCL.  Syntax-only: no library or GPU required."""
    import shutil
    import subprocess
    inc = os.path.join(ROOT, "include")
    if shutil.which("gcc") is None or shutil.which("g++") is None:
        pytest.skip("no host compiler")
    c = tmp_path / "t.c"
    c.write_text('#include "rm_radar_b200.h"\nint main(void){ rmr_robot_t r; (void)r; return sizeof(rmr_detection_t) != 24; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I" + inc, str(c)])
    cpp = tmp_path / "t.cpp"
    cpp.write_text('#include "radar.hpp"\nint main(){ radar::Robot r; return r.isDetected() || r.isLocated(); }\n')
    subprocess.check_call(["g++", "-std=c++20", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I" + inc, str(cpp)])


def test_jpeg_header_parser_scope_matches_the_oracle():
    """Host-side half of the JPEG stage (no GPU): rmr_jpeg_info accepts and rejects exactly what the oracle does,
    including the metadata that changes what cv::imread returns (EXIF orientation, RGB-coded files)."""
    import glob

    import numpy as np

    import rm_radar_b200 as rr
    from oracle import jpeg_oracle as jo
    from tests.test_oracle_jpeg import JPEG_DIR, metadata_cases
    for path in sorted(glob.glob(os.path.join(JPEG_DIR, "*.jpg"))):
        data = open(path, "rb").read()
        assert rr.jpeg_info(data) == jo.info(data), path
    for name, data, ok in metadata_cases():
        if ok:
            assert rr.jpeg_info(data) == jo.info(data), name
        else:
            with pytest.raises(ValueError):
                rr.jpeg_info(data)
            with pytest.raises(ValueError):
                jo.info(data)
    for bad in (b"", b"\xff\xd8", b"not a jpeg", open(os.path.join(JPEG_DIR, "photo_420_q90.jpg"), "rb").read()[:100]):
        with pytest.raises(ValueError):
            rr.jpeg_info(bad)


def test_comm_pack_matches_the_python_record_layout():
    """rmr_comm_pack (the block csrc/comm.cu all-gathers) == rm_radar_b200.dist.pack_records, field by field."""
    import numpy as np
    import rm_radar_b200 as rr
    from rm_radar_b200 import dist
    rng = np.random.default_rng(3)
    recs = (_lib.RobotRec * 6)()
    for i in range(5):
        r = recs[i]
        r.has_rect = 1
        for k in range(4):
            r.rect[k] = float(rng.uniform(1, 500))
        r.is_detected = int(i % 2)
        r.label = int(rng.integers(0, 12))
        r.confidence = float(rng.uniform(0.5, 1))
        r.is_located = int(i % 3 != 0)
        for k in range(3):
            r.location[k] = float(rng.uniform(-10, 10))
    for n, cap in ((5, 6), (5, 3), (0, 4)):
        assert np.array_equal(rr.Comm.pack(recs, n, cap), dist.pack_records(recs, n, cap).numpy())
