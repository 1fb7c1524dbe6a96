"""The library's engine builder (csrc/engine.cu: `rmr_engine_build`, `rmr_engine_resolve`) — host only, no GPU.

The reference resolves `<x>.engine` to a cached TensorRT engine or builds it from the sibling `<x>.onnx`
(/root/reference/src/detect/detector.cpp:74-99, 177-243); a path with neither file is std::invalid_argument
(detector.cpp:80).  The C++ builder must write the very bytes the Python restatement (rm_radar_b200/engine.py)
writes, so plans are interchangeable between the two."""
import ctypes as C
import os
import shutil

import pytest

import rm_radar_b200 as rr
from rm_radar_b200 import _lib, engine
from tests import fixtures as fx

needs_onnx = pytest.mark.skipif(not fx.have_onnx(), reason="ONNX models not in this snapshot")


@needs_onnx
@pytest.mark.parametrize("name", ["car", "armor"])
def test_cc_builder_writes_the_python_builders_bytes(name, tmp_path):
    out = tmp_path / f"{name}.rmeng"
    rr.build_engine(fx.onnx(name), str(out))
    want = engine.serialize(engine.compile_onnx(fx.onnx(name)))
    got = out.read_bytes()
    assert len(got) == len(want)
    assert got == want
    if fx.have_models():       # and the plan the GPU tests run is this one
        assert got == open(fx.engine(name), "rb").read()
    assert not (tmp_path / f"{name}.rmeng.tmp").exists()


@needs_onnx
def test_resolve_builds_from_the_sibling_onnx(tmp_path):
    os.symlink(fx.onnx("car"), tmp_path / "car.onnx")
    for asked in ("car.engine", "car.onnx", "car.rmeng"):
        got = rr.engine_path_for(str(tmp_path / asked))
        assert got == str(tmp_path / "car.rmeng")
        assert os.path.getsize(got) > 1 << 20
    first = os.path.getmtime(tmp_path / "car.rmeng")
    rr.engine_path_for(str(tmp_path / "car.engine"))          # cached: not rebuilt
    assert os.path.getmtime(tmp_path / "car.rmeng") == first


def test_neither_engine_nor_onnx_is_invalid_argument(tmp_path):
    with pytest.raises(ValueError, match="neither"):
        rr.engine_path_for(str(tmp_path / "nothing.engine"))
    lib = _lib.load()
    buf = C.create_string_buffer(16)
    assert lib.rmr_engine_resolve(str(tmp_path / "nothing.engine").encode(), 640, 640, buf, len(buf)) == -1
    # the detector constructor reports it the same way, before touching the device (detector.cpp:80)
    h = C.c_void_p()
    rc = lib.rmr_detector_create(C.byref(h), str(tmp_path / "nothing.engine").encode(), 1, 1280, 1024, 1, 0.65, 0.25,
                                 640, 640, 1, 0)
    assert rc == -1 and not h.value


@needs_onnx
def test_resolved_path_must_fit_the_buffer(tmp_path):
    os.symlink(fx.onnx("car"), tmp_path / "car.onnx")
    buf = C.create_string_buffer(4)
    assert _lib.load().rmr_engine_resolve(str(tmp_path / "car.engine").encode(), 640, 640, buf, len(buf)) == -4


def test_malformed_onnx_is_an_error_not_a_crash(tmp_path):
    lib = _lib.load()
    bad = tmp_path / "bad.onnx"
    for blob in (b"", b"\x3a\xff\xff\xff\xff\x0f", b"\x08\x07\x3a\x02\x0a\x7f", os.urandom(4096)):
        bad.write_bytes(blob)
        rc = lib.rmr_engine_build(str(bad).encode(), str(tmp_path / "bad.rmeng").encode(), 640, 640)
        assert rc in (-1, -2), rc      # invalid_argument / runtime_error (detector.cpp:184,199,205)
        assert not (tmp_path / "bad.rmeng").exists()
    assert lib.rmr_engine_build(str(tmp_path / "absent.onnx").encode(), str(tmp_path / "x.rmeng").encode(), 640, 640) == -1
    assert lib.rmr_engine_build(str(bad).encode(), str(tmp_path / "x.rmeng").encode(), 0, 640) == -1


@needs_onnx
def test_truncated_onnx_is_an_error(tmp_path):
    data = open(fx.onnx("car"), "rb").read()
    cut = tmp_path / "cut.onnx"
    cut.write_bytes(data[: len(data) // 2])
    assert _lib.load().rmr_engine_build(str(cut).encode(), str(tmp_path / "cut.rmeng").encode(), 640, 640) in (-1, -2)


@pytest.mark.gpu
@needs_onnx
def test_detector_given_an_engine_path_builds_the_plan_and_detects(tmp_path):
    """radar::Detector("car.engine", ...) next to car.onnx — the reference's first-run flow (detector.cpp:74-99):
    the C++ constructor itself builds the plan; the detections are those of the prebuilt plan."""
    import numpy as np
    os.symlink(fx.onnx("car"), tmp_path / "car.onnx")
    lib = _lib.load()
    h = C.c_void_p()
    _lib.check(lib.rmr_detector_create(C.byref(h), str(tmp_path / "car.engine").encode(), 1, fx.IMAGE_SIZE[0],
                                       fx.IMAGE_SIZE[1], 1, 0.65, 0.25, 640, 640, 1, 0))
    lib.rmr_detector_destroy(h)
    assert (tmp_path / "car.rmeng").exists()
    img = fx.load_frame(0)
    fresh = rr.Detector(str(tmp_path / "car.engine"), 1, fx.IMAGE_SIZE, 1).detect(img)
    ready = rr.Detector(fx.engine("car"), 1, fx.IMAGE_SIZE, 1).detect(img)
    assert len(fresh) == len(ready) > 0
    assert np.array_equal(np.array([d.as_array() for d in fresh]), np.array([d.as_array() for d in ready]))
