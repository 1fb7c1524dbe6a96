"""GPU: the C++ host mirror (include/radar.hpp) driven like SampleRadar::runOnce, compared with the
oracle (golden cars / labels / positions) — the parity test a reference maintainer would read."""
import json
import os
import subprocess

import numpy as np
import pytest

from tests import fixtures as fx

pytestmark = pytest.mark.gpu
BIN = os.path.join(fx.ROOT, "tests", "cpp", "build", "radar_hpp_test")


@pytest.mark.skipif(not (fx.have_models() and os.path.exists(BIN)), reason="engines / C++ test binary not built")
def test_cpp_host_run_once(tmp_path):
    from oracle import locate_oracle as lo
    exp = np.load(os.path.join(fx.GOLDEN, "expected.npz"))
    img = fx.load_frame(0)
    clouds = fx.load_clouds()
    (tmp_path / "frame.bgr").write_bytes(np.ascontiguousarray(img).tobytes())
    (tmp_path / "bg.f32").write_bytes(np.ascontiguousarray(clouds["background"][:, :3], np.float32).tobytes())
    (tmp_path / "c0.f32").write_bytes(np.ascontiguousarray(clouds["c0"][:, :3], np.float32).tobytes())
    out = subprocess.run([BIN, fx.engine("car"), fx.engine("armor"), str(tmp_path / "frame.bgr"),
                          str(img.shape[1]), str(img.shape[0]), str(tmp_path / "bg.f32"), str(tmp_path / "c0.f32"),
                          os.path.join(fx.GOLDEN, "frames", "0.jpg")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    # the document is the last line (a library banner, e.g. NCCL's version line, may precede it on stdout)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, (out.stdout[-2000:], out.stderr[-2000:])
    doc = json.loads(lines[-1])
    assert doc["ctor_throws"] is True
    assert doc["run_once_same"] is True      # radar::runOnce == update + cluster + detect + search
    assert doc["run_batch_same"] is True     # radar::runBatch over two streams == the single-stream robots, twice
    assert doc["exchange_ok"] is True        # radar::Exchange (NCCL all-gather, world of one) returns the record block
    assert doc["jpeg_same"] is True          # radar::JpegDecoder::imdecode == cv2.imread, byte for byte
    assert doc["jpeg_detect_same"] is True and doc["jpeg_rejects"] is True
    robots = doc["robots"]
    rects = np.array([r["rect"] for r in robots], np.float32)
    assert len(rects) == len(exp["f0_robot_rects"])
    for a, b in zip(rects, exp["f0_robot_rects"]):
        assert fx.iou_xywh(a, b) >= 0.99
    assert [r["label"] for r in robots if r["n_armors"] >= 0] == exp["f0_robot_labels"].tolist()
    assert np.allclose([r["confidence"] for r in robots if r["n_armors"] >= 0], exp["f0_robot_conf"], atol=5e-3)
    for r in robots:     # Robot::rect(): Rect2f -> Rect by cvRound (robot.h:111)
        assert r["rect_int"] == [int(np.rint(np.float32(v))) for v in r["rect"]]
    ora = lo.LocatorOracle(fx.IMAGE_SIZE[0], fx.IMAGE_SIZE[1], fx.INTRINSIC, fx.LIDAR_TO_CAMERA, fx.WORLD_TO_CAMERA)
    ora.update(clouds["background"]); ora.update(clouds["c0"]); ora.cluster()
    want = ora.search([tuple(r["rect"]) for r in robots])
    assert any(w is not None for w in want)
    for r, w in zip(robots, want):
        assert r["located"] == (w is not None)
        if w is not None:
            assert np.abs(np.asarray(r["location"]) - w).max() < 1e-3     # metres (north_star tolerance)
