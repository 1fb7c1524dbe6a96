"""Pins oracle/detect_oracle.py against the reference's own golden vectors
(/root/reference/test/detect/kernel_test.cu, detector_test.cpp).  CPU only."""
import numpy as np
import pytest

from oracle import detect_oracle as do

# fixture of kernel_test.cu:17-37: 4x4x3 u8 ramp 0..47
RAMP = np.arange(48, dtype=np.uint8).reshape(4, 4, 3)

# kernel_test.cu:71-85
RESIZE_DOUBLE = [
    0, 1, 2, 1, 2, 3, 3, 4, 5, 4, 5, 6, 6, 7, 8, 7, 8, 9,
    9, 10, 11, 9, 10, 11, 6, 7, 8, 7, 8, 9, 9, 10, 11, 10, 11, 12,
    12, 13, 14, 13, 14, 15, 15, 16, 17, 15, 16, 17, 12, 13, 14, 13, 14, 15,
    15, 16, 17, 16, 17, 18, 18, 19, 20, 19, 20, 21, 21, 22, 23, 21, 22, 23,
    18, 19, 20, 19, 20, 21, 21, 22, 23, 22, 23, 24, 24, 25, 26, 25, 26, 27,
    27, 28, 29, 27, 28, 29, 24, 25, 26, 25, 26, 27, 27, 28, 29, 28, 29, 30,
    30, 31, 32, 31, 32, 33, 33, 34, 35, 33, 34, 35, 30, 31, 32, 31, 32, 33,
    33, 34, 35, 34, 35, 36, 36, 37, 38, 37, 38, 39, 39, 40, 41, 39, 40, 41,
    36, 37, 38, 37, 38, 39, 39, 40, 41, 40, 41, 42, 42, 43, 44, 43, 44, 45,
    45, 46, 47, 45, 46, 47, 36, 37, 38, 37, 38, 39, 39, 40, 41, 40, 41, 42,
    42, 43, 44, 43, 44, 45, 45, 46, 47, 45, 46, 47]
# kernel_test.cu:87-90
RESIZE_HALF = [0, 1, 2, 6, 7, 8, 24, 25, 26, 30, 31, 32]
# kernel_test.cu:125-139  (top=2,bottom=2,left=1,right=1)
BORDER = [128] * 39 + list(range(0, 12)) + [128] * 6 + list(range(12, 24)) + [128] * 6 + \
    list(range(24, 36)) + [128] * 6 + list(range(36, 48)) + [128] * 39


def test_resize_double_golden():
    assert do.resize(RAMP, 8, 8).reshape(-1).tolist() == RESIZE_DOUBLE


def test_resize_half_golden():
    assert do.resize(RAMP, 2, 2).reshape(-1).tolist() == RESIZE_HALF


def test_copy_make_border_golden():
    out = do.copy_make_border(RAMP, 2, 2, 1, 1)
    assert out.shape == (8, 6, 3)
    assert out.reshape(-1).tolist() == BORDER


def test_blob_equals_blobFromImage():
    # kernel_test.cu:141-173: Blob == cv::dnn::blobFromImage(src, 0.01, Size(), Scalar(), swapRB=true)
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, (32, 48, 3), dtype=np.uint8)
    ref = cv2.dnn.blobFromImage(src, 0.01, (0, 0), (0, 0, 0), True)
    got = do.blob(src, np.float32(0.01))
    assert np.array_equal(ref[0], got)


@pytest.mark.parametrize("w,h,dw,dh", [(810, 1080, 80, 0), (1280, 720, 0, 140)])
def test_preparam_golden(w, h, dw, dh):
    # detector_test.cpp:38-41, 58-67
    pp = do.preparam(w, h)
    assert pp.width == w and pp.height == h
    assert pp.dw == np.float32(dw) and pp.dh == np.float32(dh)


def test_letterbox_geometry_survey():
    # 1920x1080 -> 640x360 + 140/140; 2592x2048 -> 505.68 -> 505 rows (compat) + 67/67
    u8, pp = do.letterbox_u8(np.zeros((1080, 1920, 3), np.uint8))
    assert (u8[:140] == 128).all() and (u8[140:500] == 0).all() and (u8[500:] == 128).all()
    u8, pp = do.letterbox_u8(np.full((2048, 2592, 3), 7, np.uint8), compat=True)
    assert (u8[:67] == 128).all() and (abs(u8[67:67 + 505].astype(int) - 6.5) < 1).all() and (u8[572:639] == 128).all()
    assert (u8[639] == 0).all()          # stale row: never written (fresh buffer = 0)
    u8, pp = do.letterbox_u8(np.full((2048, 2592, 3), 7, np.uint8), compat=False)
    assert (abs(u8[67:67 + 506].astype(int) - 6.5) < 1).all() and (u8[573:] == 128).all()


def test_letterbox_shear_compat():
    # 639-column case: stride 639*3 written, 640*3 read -> one pixel shear per row
    img = np.full((100, 57, 3), 9, np.uint8)  # h-limited: ratio = 100/640, w/ratio = 364.8 -> 364 (compat)
    u8, pp = do.letterbox_u8(img, compat=True)
    pw = int(np.float32(pp.width / pp.ratio))
    left = int(do.c_round(np.float32(float(pp.dw) - 0.1)))
    right = int(do.c_round(np.float32(float(pp.dw) + 0.1)))
    bw = pw + left + right
    assert bw == 639
    flat = u8.reshape(-1)
    row5 = flat[5 * bw * 3:(5 * bw + bw) * 3].reshape(bw, 3)
    assert (row5[:left] == 128).all() and (abs(row5[left:left + pw].astype(int) - 8.5) < 1).all()


def test_decode_and_nms_semantics():
    # decodeKernel: clamp x,y without shrinking w,h; first max wins
    out = np.zeros((4 + 3, 3), np.float32)
    out[:4, 0] = [5, 5, 20, 4]      # cx-w/2 < 0 -> x=0, w stays 20
    out[4:, 0] = [0.3, 0.9, 0.9]    # tie -> label 1
    out[:4, 1] = [100, 100, 10, 10]
    out[4:, 1] = [0.8, 0.1, 0.1]
    out[:4, 2] = [101, 100, 10, 10]
    out[4:, 2] = [0.7, 0.1, 0.1]
    d = do.decode(out, 3)
    assert d[0].tolist() == [0, 3, 20, 4, 1, np.float32(0.9)]
    keep = do.nms(d, 0.65, 0.25)
    assert keep.tolist() == [0, 1]   # row 2 suppressed by row 1 (same label, higher conf, IoU .818)
    # all-pairs, not greedy: chain A>B>C where A kills B, B kills C, A does not overlap C
    d = np.array([[0, 0, 10, 10, 0, .9], [3, 0, 10, 10, 0, .8], [6, 0, 10, 10, 0, .7]], np.float32)
    assert do.nms(d, 0.5, 0.25).tolist() == [0]          # greedy NMS would keep [0, 2]
    # equal confidences both survive
    d = np.array([[0, 0, 10, 10, 0, .9], [0, 0, 10, 10, 0, .9]], np.float32)
    assert do.nms(d, 0.5, 0.25).tolist() == [0, 1]


def test_restore_clamps():
    pp = do.preparam(1920, 1080)
    d = np.array([[0, 100, 700, 100, 0, .5]], np.float32)
    r = do.restore(d, pp)
    assert r[0, 0] == 0 and r[0, 1] == 0 and r[0, 2] == 1920 and r[0, 3] == np.float32(300)


def test_set_detection_vote():
    car = np.array([10, 20, 100, 100, 0, .9], np.float32)
    arm = np.array([[1, 1, 5, 5, 3, .6], [2, 2, 5, 5, 7, .5], [3, 3, 5, 5, 7, .4]], np.float32)
    r = do.set_detection(car, arm)
    assert r.label == 7 and abs(r.confidence - 0.45) < 1e-6
    assert r.armors[0, 0] == 11 and r.armors[0, 1] == 21
    assert not do.set_detection(car, np.zeros((0, 6), np.float32)).is_detected()


def test_compute_iou_bounding():
    # intersection / bounding rectangle area (detector.cpp:335-345)
    assert do.compute_iou_bounding((0, 0, 10, 10), (5, 5, 10, 10)) == np.float32(25 / 225)
    assert do.compute_iou_bounding((0, 0, 10, 10), (20, 20, 5, 5)) == 0
