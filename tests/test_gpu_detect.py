"""GPU: detect path vs the oracle through the C ABI — preprocess bit-exact (u8 stage exact, fp16
rounding of the /255 blob), network head output within fp16 tolerance of the fp32 ONNX oracle,
final detections class-exact with IoU >= 0.99 (BASELINE north_star gate)."""
import os

import numpy as np
import pytest

import rm_radar_b200 as rr
from oracle import detect_oracle as do
from tests import fixtures as fx

pytestmark = pytest.mark.gpu

needs_models = pytest.mark.skipif(not fx.have_models(), reason="engines not built")
needs_onnx = pytest.mark.skipif(not fx.have_onnx(), reason="fp32 onnx copies for the live oracle not present")


@pytest.fixture(scope="module")
def robot_detector():
    return rr.RobotDetector(fx.engine("car"), fx.engine("armor"), fx.IMAGE_SIZE, fx.CLASS_NUM, fx.MAX_BATCH, fx.OPT_BATCH)


@pytest.fixture(scope="module")
def oracle_nets():
    from oracle.onnx_torch import OnnxNet
    car, armor = OnnxNet(fx.onnx("car")), OnnxNet(fx.onnx("armor"))
    return (lambda x: car(x).numpy()), (lambda x: armor(x).numpy())


def fp16_blob(img, compat=True, border=None):
    b, pp = do.preprocess(img, compat=compat, border_buf=border)
    return b.astype(np.float16).astype(np.float32), pp


@needs_models
@pytest.mark.parametrize("size", [(1920, 1080), (1280, 1280), (2592, 2048), (810, 1080), (1280, 720), (57, 100),
                                  (133, 178), (29, 44)])
def test_preprocess_bit_exact(size):
    """Letterbox kernel vs oracle for clean (1920x1080, 1280^2, the reference's bus/zidane sizes) and
    bug-compatible 639-row / 639-column geometries, on a fresh detector (staging = zeros)."""
    w, h = size
    rng = np.random.default_rng(w * 7 + h)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    det = rr.Detector(fx.engine("car"), 1, (w, h), 1)
    det.detect(img)
    got = det.last_input(1)[0]
    want, _ = fp16_blob(img)
    assert np.array_equal(got, want), f"{np.sum(got != want)} of {got.size} input values differ"


@needs_models
def test_preprocess_corrected_mode_and_stale_staging():
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (100, 57, 3), dtype=np.uint8)     # 639-column shear case
    b = rng.integers(0, 256, (90, 61, 3), dtype=np.uint8)
    det = rr.Detector(fx.engine("car"), 1, (64, 100), 1)
    border = np.zeros(640 * 640 * 3, np.uint8)
    for img in (a, b, a):       # stale bytes of the previous call must survive exactly like the reference buffer
        det.detect(img)
        want, _ = fp16_blob(img, border=border)
        assert np.array_equal(det.last_input(1)[0], want)
    det2 = rr.Detector(fx.engine("car"), 1, (64, 100), 1, compat=False)
    det2.detect(a)
    want, _ = fp16_blob(a, compat=False)
    assert np.array_equal(det2.last_input(1)[0], want)


@needs_models
@needs_onnx
def test_network_output_vs_fp32_oracle(oracle_nets):
    """Head output [1,5,34000] of the tcgen05 conv stack vs the fp32 ONNX graph on the golden frame.
    fp16 operands / fp32 accumulate: SURVEY C.4 measured 0.009 px / 3e-4; gate at 0.25 px / 4e-3."""
    car_net, _ = oracle_nets
    img = fx.load_frame(0)
    det = rr.Detector(fx.engine("car"), 1, fx.IMAGE_SIZE, 1)
    dets = det.detect(img)
    x, pp = do.preprocess(img)
    ref = car_net(x[None])[0]
    got = det.last_output(1)[0]
    assert got.shape == ref.shape == (5, 34000)
    hot = ref[4] > 0.05
    assert np.abs(got[4] - ref[4]).max() < 4e-3
    assert np.abs(got[:4, hot] - ref[:4, hot]).max() < 0.25
    want = do.postprocess(ref, 1, pp, 0.65, 0.25)
    fx.match_detections([d.as_array() for d in dets], want)


@needs_models
def test_cascade_golden_frames(robot_detector):
    exp = np.load(os.path.join(fx.GOLDEN, "expected.npz"))
    for f in (0, 5, 0):
        img = fx.load_frame(f)
        robots = robot_detector.detect(img)
        cars = [d.as_array() for d in robot_detector.last_cars()]
        fx.match_detections(cars, exp[f"f{f}_cars"])
        counts = [len(robot_detector.last_armors(i)) for i in range(len(cars))]
        assert counts == exp[f"f{f}_armor_counts"].tolist()
        armors = [d.as_array() for i in range(len(cars)) for d in robot_detector.last_armors(i)]
        fx.match_detections(armors, exp[f"f{f}_armors"], min_iou=0.99)
        labels = [r.label for r in robots if r.isDetected()]
        assert labels == exp[f"f{f}_robot_labels"].tolist()
        conf = [r.confidence for r in robots if r.isDetected()]
        assert np.allclose(conf, exp[f"f{f}_robot_conf"], atol=5e-3)
        rects = np.array([r.rect for r in robots], np.float32)
        for a, b in zip(rects, exp[f"f{f}_robot_rects"]):
            assert fx.iou_xywh(a, b) >= 0.99
        # output order: undetected robots in car order, then labelled ascending (detector.cpp:431-453)
        det_flags = [r.isDetected() for r in robots]
        assert det_flags == sorted(det_flags) and labels == sorted(labels)
        # the event timing bench.py's roofline uses: both replays ran on the device for this call
        car_ms, armor_ms = robot_detector.last_timing()
        assert 0.05 < car_ms < 50 and (armor_ms > 0.05) == (len(cars) > 0)


@needs_models
@needs_onnx
def test_cascade_1080p_matches_live_oracle(robot_detector, oracle_nets):
    """BASELINE config C2 geometry (1920x1080): CUDA path vs the oracle run live on the same frame."""
    exp = np.load(os.path.join(fx.GOLDEN, "expected.npz"))
    img = fx.resize_frame(fx.load_frame(0), 1920, 1080)
    det = rr.RobotDetector(fx.engine("car"), fx.engine("armor"), (1920, 1080), fx.CLASS_NUM, fx.MAX_BATCH, fx.OPT_BATCH)
    robots = det.detect(img)
    cars = [d.as_array() for d in det.last_cars()]
    fx.match_detections(cars, exp["f0_1080_cars"])
    assert [len(det.last_armors(i)) for i in range(len(cars))] == exp["f0_1080_armor_counts"].tolist()
    assert [r.label for r in robots if r.isDetected()] == exp["f0_1080_robot_labels"].tolist()
    car_net, armor_net = oracle_nets
    tr = do.CascadeTrace()
    want = do.robot_detect(img, car_net, armor_net, trace=tr)
    assert [r.label for r in want if r.is_detected()] == [r.label for r in robots if r.isDetected()]


@needs_models
def test_detector_batch_and_edge_cases():
    det = rr.Detector(fx.engine("armor"), 12, (640, 640), 4, conf_thresh=0.5)
    assert det.detect([]) == []
    rng = np.random.default_rng(1)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (w, h) in [(133, 178), (64, 64), (300, 200)]]
    out = det.detect(imgs)
    assert len(out) == 3 and all(isinstance(o, list) for o in out)
    got = det.last_input(3)
    border = [np.zeros(640 * 640 * 3, np.uint8) for _ in range(3)]
    for i, im in enumerate(imgs):
        want, _ = fp16_blob(im, border=border[i])
        assert np.array_equal(got[i], want)
    with pytest.raises(ValueError):
        det.detect(imgs + imgs)            # batch > max_batch_size
    with pytest.raises(ValueError):
        rr.Detector(fx.engine("armor"), 3, (640, 640), 1)   # class count mismatch
    # a black frame: no cars -> empty armor batch (the reference aborts inside TensorRT here, B#8)
    rd = rr.RobotDetector(fx.engine("car"), fx.engine("armor"), (640, 480), 12, 20, 4)
    assert rd.detect(np.zeros((480, 640, 3), np.uint8)) == []


@needs_models
def test_batch_of_real_frames_equals_single_calls():
    """BASELINE config C3 geometry (1280x1280 frames, batch): Detector.detect(list) gives every image the
    detections the single-image call gives it (different batch plans / tile shapes, same results)."""
    det = rr.Detector(fx.engine("car"), 1, (1280, 1280), 4)
    f0 = fx.resize_frame(fx.load_frame(0), 1280, 1280)
    f5 = fx.resize_frame(fx.load_frame(5), 1280, 1280)
    imgs = [f0, f5, np.ascontiguousarray(f0[:, ::-1]), f5]
    batch = det.detect(imgs)
    assert len(batch) == 4 and len(batch[0]) >= 4
    for img, dets in zip(imgs, batch):
        single = det.detect(img)
        fx.match_detections([d.as_array() for d in dets], [d.as_array() for d in single])
    assert [d.as_array().tolist() for d in batch[1]] == [d.as_array().tolist() for d in batch[3]]   # same image, same slot-independent result


@needs_models
def test_run_once_equals_separate_calls():
    """rmr_run_once (SampleRadar::runOnce body) == update + cluster + detect + search on the same inputs."""
    img = fx.load_frame(0)
    clouds = fx.load_clouds()
    kw = dict(device=0)
    det_a = rr.RobotDetector(fx.engine("car"), fx.engine("armor"), fx.IMAGE_SIZE, fx.CLASS_NUM, fx.MAX_BATCH, fx.OPT_BATCH, **kw)
    det_b = rr.RobotDetector(fx.engine("car"), fx.engine("armor"), fx.IMAGE_SIZE, fx.CLASS_NUM, fx.MAX_BATCH, fx.OPT_BATCH, **kw)
    loc_a = rr.Locator(*fx.IMAGE_SIZE, fx.INTRINSIC, fx.LIDAR_TO_CAMERA, fx.WORLD_TO_CAMERA)
    loc_b = rr.Locator(*fx.IMAGE_SIZE, fx.INTRINSIC, fx.LIDAR_TO_CAMERA, fx.WORLD_TO_CAMERA)
    loc_a.update(clouds["background"]); loc_b.update(clouds["background"])
    for key in ("c0", "c1"):
        loc_a.update(clouds[key]); loc_a.cluster()
        want = det_a.detect(img)
        loc_a.search(want)
        got = rr.run_once(det_b, loc_b, img, clouds[key])
        assert len(got) == len(want) and any(r.location is not None for r in got)
        for g, w in zip(got, want):
            assert g.rect == w.rect and g.label == w.label and g.confidence == w.confidence
            assert g.location == w.location and g.cluster == w.cluster


@needs_models
def test_run_batch_equals_run_once_per_stream():
    """Throughput mode (BASELINE config[2], rmr_run_batch): n camera + LiDAR streams in one call — the car network batched over
    the frames, the armor network over all their ROIs, one Locator per stream — gives every stream what rmr_run_once
    gives it alone (Detector::detect(container) vs detect(Mat), /root/reference/src/detect/detector.cu:380-502)."""
    clouds = fx.load_clouds()
    frames = [fx.load_frame(0), fx.load_frame(5), np.ascontiguousarray(fx.load_frame(0)[:, ::-1])]
    keys = ["c0", "c1", "c2"]
    n = len(frames)
    det1 = rr.RobotDetector(fx.engine("car"), fx.engine("armor"), fx.IMAGE_SIZE, fx.CLASS_NUM, fx.MAX_BATCH, fx.OPT_BATCH)
    detn = rr.RobotDetector(fx.engine("car"), fx.engine("armor"), fx.IMAGE_SIZE, fx.CLASS_NUM, fx.MAX_BATCH, fx.OPT_BATCH, frames=n)
    mk = lambda: rr.Locator(*fx.IMAGE_SIZE, fx.INTRINSIC, fx.LIDAR_TO_CAMERA, fx.WORLD_TO_CAMERA)   # noqa: E731
    locs_a, locs_b = [mk() for _ in range(n)], [mk() for _ in range(n)]
    for la, lb in zip(locs_a, locs_b):
        la.update(clouds["background"]); lb.update(clouds["background"])
    npts = min(len(clouds[k]) for k in keys)
    cl = [np.ascontiguousarray(clouds[k][:npts]) for k in keys]
    want = [rr.run_once(det1, locs_a[i], frames[i], cl[i]) for i in range(n)]
    got = rr.run_batch(detn, locs_b, frames, cl)
    assert len(got) == n and sum(len(w) for w in want) >= 6
    for g_robots, w_robots in zip(got, want):
        assert len(g_robots) == len(w_robots)
        for g, w in zip(g_robots, w_robots):
            assert g.label == w.label and (g.confidence is None) == (w.confidence is None)
            if g.confidence is not None:
                assert abs(g.confidence - w.confidence) <= 5e-3
            assert fx.iou_xywh(g.rect, w.rect) >= 0.99
            assert (g.location is None) == (w.location is None)
            if g.location is not None:
                assert np.allclose(g.location, w.location, atol=1e-3)
    # detect_frames (detection only) agrees with the cascade inside run_batch
    only = detn.detect_frames(frames)
    assert [[(r.label, r.rect) for r in f] for f in only] == [[(r.label, r.rect) for r in f] for f in got]


@needs_models
def test_whole_chain_jpeg_detect_locate_track():
    """SampleRadar::runOnce end to end (sample_radar.h:106-127): JPEG in, tracked robots out, over repeated frames."""
    from oracle import track_oracle as to
    clouds = fx.load_clouds()
    det = rr.RobotDetector(fx.engine("car"), fx.engine("armor"), fx.IMAGE_SIZE, fx.CLASS_NUM, fx.MAX_BATCH, fx.OPT_BATCH)
    loc = rr.Locator(fx.IMAGE_SIZE[0], fx.IMAGE_SIZE[1], fx.INTRINSIC, fx.LIDAR_TO_CAMERA, fx.WORLD_TO_CAMERA)
    dec = rr.JpegDecoder(0)
    trk = rr.Tracker([0.2, 0.2, 0.2], fx.CLASS_NUM)
    ora = to.Tracker([0.2, 0.2, 0.2], fx.CLASS_NUM)
    loc.update(clouds["background"])
    jpg = open(os.path.join(fx.GOLDEN, "frames", "0.jpg"), "rb").read()
    img = dec.decode(jpg)
    t = 0
    import copy
    for frame in range(6):
        t += 40_000_000
        plain = rr.run_once(det, loc, img, clouds["c0"])          # the observation of this frame
        obs = [to.RobotObs(armors=[(a.label, a.confidence) for a in r.armors] if r.armors is not None else None,
                           location=r.location, label=r.label) for r in plain]
        robots = copy.deepcopy(plain)
        trk.update(robots, t)
        ora.update(obs, t)
        assert [r.track_state for r in robots] == [o.track_state for o in obs]
        assert [r.label for r in robots] == [o.label for o in obs]
        for r, o in zip(robots, obs):
            if o.location is not None:
                assert np.allclose(r.location, o.location, atol=1e-3)
    located = [r for r in robots if r.isDetected() and r.isLocated()]
    assert located and all(r.track_state == 1 for r in located)   # confirmed after init_thresh = 4 frames
    assert len(trk.tracks()) == len(ora.tracks) > 0
    # the fused entry point with the tracker attached: same tracks keep being matched
    t += 40_000_000
    fused = rr.run_once(det, loc, img, clouds["c0"], tracker=trk, timestamp_ns=t)
    assert len(fused) == len(robots)
    assert [r.track_id for r in fused if r.isDetected() and r.isLocated()] == [r.track_id for r in located]


# ---------------------------------------------------------------------------------------------------------------
# NMS + restore kernel on synthetic candidates (rmr_postprocess_selftest): the cases real frames never produce
# ---------------------------------------------------------------------------------------------------------------
def _nms_check(cand, thresh=0.65):
    got = rr.postprocess_selftest(cand, thresh)
    keep = do.nms(cand.astype(np.float32), thresh, 0.0)        # every row is already above the score threshold
    assert np.array_equal(got, cand[keep].astype(np.float32))   # survivors, in anchor order, untouched (identity restore)
    return keep


def test_nms_kernel_suppression_chain_is_all_pairs_not_greedy():
    """A suppresses B, B would suppress C: greedy NMS keeps C, the reference's all-pairs rule kills it
    (NMSKernel compares every row with every column whatever became of the column, SURVEY B#5)."""
    cand = np.array([[0, 0, 100, 100, 0, 0.9], [20, 0, 100, 100, 0, 0.8], [40, 0, 100, 100, 0, 0.7]], np.float32)
    keep = _nms_check(cand, 0.6)
    assert keep.tolist() == [0]


def test_nms_kernel_ties_labels_and_order():
    rng = np.random.default_rng(5)
    cand = []
    for c in range(40):                                   # clusters with equal confidences and mixed labels
        x, y = rng.uniform(5, 600, 2)     # away from 0: restoreDetection clamps negative corners
        for k in range(int(rng.integers(1, 6))):
            cand.append([x + rng.uniform(-2, 2), y + rng.uniform(-2, 2), 40, 40, float(rng.integers(0, 3)),
                         float(rng.choice([0.5, 0.6, 0.6, 0.9]))])
    cand = np.array(cand, np.float32)
    cand = cand[rng.permutation(len(cand))]
    keep = _nms_check(cand)
    assert 40 <= len(keep) < len(cand)
    same = np.array([[10, 10, 20, 20, 1, 0.7]] * 3, np.float32)   # identical boxes, identical confidence: all survive (B#6)
    assert _nms_check(same).tolist() == [0, 1, 2]


def test_nms_kernel_large_random_sets_match_the_oracle():
    rng = np.random.default_rng(11)
    for n in (1, 255, 256, 257, 3000):
        xy = rng.uniform(0, 300, (n, 2))
        wh = rng.uniform(10, 120, (n, 2))
        cand = np.concatenate([xy, wh, rng.integers(0, 12, (n, 1)).astype(np.float64), rng.uniform(0.25, 1, (n, 1))], axis=1)
        _nms_check(cand.astype(np.float32))
    assert len(rr.postprocess_selftest(np.zeros((0, 6), np.float32), 0.65)) == 0


def test_capacities_fail_loudly_instead_of_truncating():
    """The reference returns every survivor (detector.cu:561-579); running out of room is RMR_ERR_CAPACITY, never a
    silently shorter (and atomic-order dependent) list."""
    many = np.tile(np.array([[0, 0, 10, 10, 0, 0.9]], np.float32), (16385, 1))
    with pytest.raises(rr.CapacityError, match="capacity 16384"):
        rr.postprocess_selftest(many, 0.65)
    # product path: with the confidence threshold at 0 every one of the 34 000 anchors is a candidate
    det = rr.Detector(fx.engine("car"), 1, fx.IMAGE_SIZE, 1, conf_thresh=0.0)
    with pytest.raises(rr.CapacityError, match="confidence threshold"):
        det.detect(fx.load_frame(0))
    det_ok = rr.Detector(fx.engine("car"), 1, fx.IMAGE_SIZE, 1)
    assert len(det_ok.detect(fx.load_frame(0))) >= 4        # and the library is still usable afterwards
