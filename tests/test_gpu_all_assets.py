"""GPU: detect + locate on ALL ten frames and ten clouds the reference ships (assets/images/0-9.jpg paired with
assets/clouds/0-9.pcd; north_star: "outputs match ... on the same assets/images + assets/clouds inputs").
Expected values: tests/golden/expected_all.npz, written by tests/golden/make_golden.py from the fp32 oracle; the input
files are copied by __graft_entry__.build() into the git-ignored tests/golden/_assets/ and travel with the snapshot
(frames 0 and 5 and clouds 0-3, 5 are also committed under tests/golden/)."""
import os

import numpy as np
import pytest

import rm_radar_b200 as rr
from tests import fixtures as fx

pytestmark = pytest.mark.gpu

ASSETS = os.path.join(fx.GOLDEN, "_assets")
EXPECTED = os.path.join(fx.GOLDEN, "expected_all.npz")


def _frame(i):
    import cv2
    for d in (ASSETS, os.path.join(fx.GOLDEN, "frames")):
        p = os.path.join(d, f"{i}.jpg")
        if os.path.exists(p):
            return cv2.imread(p, cv2.IMREAD_COLOR)
    pytest.skip(f"frame {i} not in this snapshot (tests/golden/_assets is filled by build() where /root/reference is mounted)")


def _clouds():
    p = os.path.join(ASSETS, "clouds_all.npz")
    if os.path.exists(p):
        z = np.load(p)
        return {k: z[k] for k in z.files}
    return fx.load_clouds()


@pytest.fixture(scope="module")
def detector():
    if not fx.have_models():
        pytest.skip("engines missing")
    return rr.RobotDetector(fx.engine("car"), fx.engine("armor"), fx.IMAGE_SIZE, fx.CLASS_NUM, fx.MAX_BATCH, fx.OPT_BATCH)


@pytest.mark.parametrize("i", range(10))
def test_detect_and_locate_match_the_oracle_on_asset_pair(i):
    """A fresh detector per pair, like the oracle run that wrote the expected values: in the reference-compatible
    letterbox mode the u8 staging buffer is persistent (SURVEY B#1: bytes of earlier ROIs survive where nothing is
    written), so a detector's history reaches the armour confidences in the fourth decimal — frame 8 has an armour at
    0.4987 against the 0.50 threshold."""
    if not fx.have_models():
        pytest.skip("engines missing")
    detector = rr.RobotDetector(fx.engine("car"), fx.engine("armor"), fx.IMAGE_SIZE, fx.CLASS_NUM, fx.MAX_BATCH, fx.OPT_BATCH)
    exp = np.load(EXPECTED)
    img = _frame(i)
    clouds = _clouds()
    if f"c{i}" not in clouds:
        pytest.skip(f"cloud {i} not in this snapshot")
    robots = detector.detect(img)
    # detection: class-exact, IoU >= 0.99, confidence within 5e-3 (north_star gate), cars and every armour
    cars = [d.as_array() for d in detector.last_cars()]
    fx.match_detections(cars, exp[f"f{i}_cars"])
    # armours per car.  A detection whose confidence sits within the 5e-3 gate of the 0.50 threshold may exist on one
    # side only (fp16 network vs fp32 oracle; the reference's own TensorRT FP16 engine has the same property): such
    # borderline detections are set aside on both sides, everything else must match one to one
    tol, thr = 5e-3, 0.50
    want_counts = exp[f"f{i}_armor_counts"].tolist()
    want_armors = exp[f"f{i}_armors"]
    borderline = 0
    off = 0
    for k in range(len(cars)):
        got_k = [d.as_array() for d in detector.last_armors(k)]
        want_k = [a for a in want_armors[off:off + want_counts[k]]]
        off += want_counts[k]
        solid_got = [g for g in got_k if abs(g[5] - thr) > tol]
        solid_want = [w for w in want_k if abs(w[5] - thr) > tol]
        borderline += (len(got_k) - len(solid_got)) + (len(want_k) - len(solid_want))
        fx.match_detections(solid_got, solid_want)
    assert borderline <= 1
    if borderline == 0:
        assert [r.label for r in robots if r.isDetected()] == exp[f"f{i}_robot_labels"].tolist()
        assert np.allclose([r.confidence for r in robots if r.isDetected()], exp[f"f{i}_robot_conf"], atol=5e-3)
        assert len(robots) == len(exp[f"f{i}_robot_rects"])
        for r, want in zip(robots, exp[f"f{i}_robot_rects"]):
            assert fx.iou_xywh(np.asarray(r.rect, np.float32), want) >= 0.99
    # locate: world xyz within 1e-3 m, the same robots located, the same foreground / cluster counts
    loc = rr.Locator(*fx.IMAGE_SIZE, fx.INTRINSIC, fx.LIDAR_TO_CAMERA, fx.WORLD_TO_CAMERA)
    loc.update(clouds["background"]); loc.update(clouds[f"c{i}"]); loc.cluster()
    loc.search(robots)
    st = loc.stats()
    assert [st["foreground"], st["clusters"]] == exp[f"f{i}_fg_clusters"].tolist()
    if borderline == 0:
        located = [r.location is not None for r in robots]
        assert located == exp[f"f{i}_located"].tolist()
        for r, want in zip(robots, exp[f"f{i}_locations"]):
            if r.location is not None:
                assert np.allclose(r.location, want, atol=1e-3), (r.location, want)
    else:
        # the robot list differs by the borderline armour's robot: compare the located robots by rectangle
        want_rects, want_loc = exp[f"f{i}_robot_rects"], exp[f"f{i}_locations"]
        for r in robots:
            j = int(np.argmax([fx.iou_xywh(np.asarray(r.rect, np.float32), w) for w in want_rects]))
            if r.location is not None and not np.isnan(want_loc[j]).any() and fx.iou_xywh(np.asarray(r.rect, np.float32), want_rects[j]) >= 0.99:
                assert np.allclose(r.location, want_loc[j], atol=1e-3)


def test_armor_head_matches_the_fp32_oracle_on_real_rois(detector):
    """armor.onnx (84 convs, 12 classes) at batch > 1 on real ROIs: the whole head tensor [n, 16, 8400] against the fp32
    ONNX oracle, not only the few post-NMS detections — a near-threshold class flip cannot hide."""
    if not fx.have_onnx():
        pytest.skip("fp32 ONNX copies not in this snapshot")
    from oracle import detect_oracle as do
    from oracle.onnx_torch import OnnxNet
    img = _frame(0)
    detector.detect(img)
    cars = [d.as_array() for d in detector.last_cars()]
    rois = [img[int(c[1]):int(c[1]) + int(c[3]), int(c[0]):int(c[0]) + int(c[2])] for c in cars[:5]]
    rois = [np.ascontiguousarray(r) for r in rois if r.size]
    assert len(rois) >= 3
    size = (max(r.shape[1] for r in rois), max(r.shape[0] for r in rois))
    det = rr.Detector(fx.engine("armor"), fx.CLASS_NUM, size, len(rois), conf_thresh=0.5)
    det.detect(rois)
    got = det.last_output(len(rois))
    armor = OnnxNet(fx.onnx("armor"))
    for k, roi in enumerate(rois):
        x, _ = do.preprocess(roi)
        ref = armor(x[None]).numpy()[0]
        assert got[k].shape == ref.shape and ref.shape[0] == 4 + fx.CLASS_NUM
        assert np.abs(got[k][4:] - ref[4:]).max() < 4e-3                    # every class score of every anchor
        hot = ref[4:].max(axis=0) > 0.05
        if hot.any():
            assert np.abs(got[k][:4, hot] - ref[:4, hot]).max() < 0.25          # boxes where there is anything to box
            assert (got[k][4:, hot].argmax(axis=0) == ref[4:, hot].argmax(axis=0)).all()   # class-exact
