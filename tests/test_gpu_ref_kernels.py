"""GPU: the oracle's pre/post-processing restatement against the REFERENCE'S OWN CUDA kernels
(resizeKernel, copyMakeBorderKernel, blobKernel, transposeKernel, decodeKernel, IoU, NMSKernel —
/root/reference/src/detect/detector.cu:40-360), compiled from the reference source into the git-ignored
oracle/_ref/libref_kernels.so by oracle/Makefile (only where /root/reference is mounted; the .so travels).
This pins oracle/detect_oracle.py at the decode / IoU / NMS boundary, where the reference holds no vectors.

NMSKernel races with itself (it writes label = NaN into the array other blocks are reading, SURVEY B#5/#6),
so the NMS comparison uses chain-free inputs: clusters in which the best box overlaps every other member
above the threshold — then every execution order gives the all-pairs result the oracle computes."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import detect_oracle as do
from tests import fixtures as fx

pytestmark = pytest.mark.gpu

REF_DIR = os.path.join(fx.ROOT, "oracle", "_ref")


def _load(name):
    so = os.path.join(REF_DIR, name)
    if not os.path.exists(so):
        pytest.skip(f"oracle/_ref/{name} not built (needs /root/reference at build time)")
    lib = C.CDLL(so)
    lib.ref_iou.restype = C.c_float
    lib.ref_iou.argtypes = [C.c_float] * 8
    return lib


@pytest.fixture(scope="module")
def ref():
    """the reference kernels compiled as their source reads: IEEE arithmetic, no FMA contraction"""
    return _load("libref_kernels_ieee.so")


@pytest.fixture(scope="module")
def ref_fast():
    """the same source under the reference's release flags (-O3 --use_fast_math, CMakeLists.txt:18)"""
    return _load("libref_kernels.so")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("sw,sh,dw,dh", [(1920, 1080, 640, 360), (2592, 2048, 640, 505), (37, 91, 260, 640), (640, 640, 640, 640),
                                          (5, 3, 13, 7)])
def test_resize_matches_reference_kernel(ref, sw, sh, dw, dh):
    rng = np.random.default_rng(sw * 7 + dh)
    src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    out = np.zeros((dh, dw, 3), np.uint8)
    assert ref.ref_resize(_p(src), _p(out), 3, sw, sh, dw, dh) == 0
    assert np.array_equal(out, do.resize(src, dw, dh))


@pytest.mark.parametrize("sw,sh,dw,dh", [(1920, 1080, 640, 360), (2592, 2048, 640, 505), (37, 91, 260, 640)])
def test_resize_under_the_release_flags_differs_by_at_most_one_lsb(ref_fast, sw, sh, dw, dh):
    """SURVEY B#2: --use_fast_math contracts the blend into FMAs and divides approximately; against the IEEE
    restatement that moves a pixel by at most one grey level (truncating cast next to an integer), rarely."""
    rng = np.random.default_rng(sw * 7 + dh)
    src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    out = np.zeros((dh, dw, 3), np.uint8)
    assert ref_fast.ref_resize(_p(src), _p(out), 3, sw, sh, dw, dh) == 0
    want = do.resize(src, dw, dh)
    d = np.abs(out.astype(np.int16) - want.astype(np.int16))
    # an approximate division can also flip floor() at an exactly-integer sample position: then the sampled
    # neighbourhood moves by one source pixel and the value by more than one level; those positions are
    # exactly the ones where dst * src / dst is an integer
    exact_x = (np.arange(dw) * sw) % dw == 0
    exact_y = (np.arange(dh) * sh) % dh == 0
    generic = ~(exact_y[:, None] | exact_x[None, :])
    if generic.any():
        assert d[generic].max() <= 1
        assert (d[generic] != 0).mean() < 0.02
    # everywhere, exact positions included, the two builds disagree on a small minority of pixels only
    assert (d != 0).mean() < 0.05


@pytest.mark.parametrize("w,h,top,bottom,left,right", [(640, 360, 140, 140, 0, 0), (505, 640, 0, 0, 67, 68), (639, 360, 140, 140, 0, 0)])
def test_copy_make_border_matches_reference_kernel(ref, w, h, top, bottom, left, right):
    rng = np.random.default_rng(w + h)
    src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    dw, dh = w + left + right, h + top + bottom
    buf = np.zeros(640 * 640 * 3, np.uint8)
    assert ref.ref_copy_make_border(_p(src), _p(buf), 3, w, h, top, bottom, left, right, 640, 640, buf.size) == 0
    want = do.copy_make_border(src, top, bottom, left, right)
    got = buf[: dh * dw * 3].reshape(dh, dw, 3)     # the kernel writes with dst_step = dst_w * channels
    assert np.array_equal(got, want)
    assert not buf[dh * dw * 3:].any()              # nothing beyond dst_h x dst_w is touched (stale bytes survive)


def test_blob_matches_reference_kernel(ref):
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, (640, 640, 3), dtype=np.uint8)
    out = np.zeros((3, 640, 640), np.float32)
    assert ref.ref_blob(_p(src), _p(out), 640, 640, 3, C.c_float(1 / 255.0)) == 0
    assert np.array_equal(out, do.blob(src))


def test_whole_preprocess_matches_reference_kernels(ref):
    """resize -> border -> blob chained exactly like Detector::preprocess (detector.cu:392-414), incl. the 639-px geometry."""
    img = fx.load_frame(0)[:1080, :1920]
    pp = do.preparam(img.shape[1], img.shape[0])
    pad_w, pad_h = int(np.float32(pp.width / pp.ratio)), int(np.float32(pp.height / pp.ratio))
    resized = np.zeros((pad_h, pad_w, 3), np.uint8)
    assert ref.ref_resize(_p(np.ascontiguousarray(img)), _p(resized), 3, img.shape[1], img.shape[0], pad_w, pad_h) == 0
    top, bottom = int(do.c_round(np.float32(np.float64(pp.dh) - 0.1))), int(do.c_round(np.float32(np.float64(pp.dh) + 0.1)))
    left, right = int(do.c_round(np.float32(np.float64(pp.dw) - 0.1))), int(do.c_round(np.float32(np.float64(pp.dw) + 0.1)))
    buf = np.zeros(640 * 640 * 3, np.uint8)
    assert ref.ref_copy_make_border(_p(resized), _p(buf), 3, pad_w, pad_h, top, bottom, left, right, 640, 640, buf.size) == 0
    out = np.zeros((3, 640, 640), np.float32)
    assert ref.ref_blob(_p(buf), _p(out), 640, 640, 3, C.c_float(1 / 255.0)) == 0
    want, _ = do.preprocess(np.ascontiguousarray(img), compat=True)
    assert np.array_equal(out, want)


@pytest.mark.parametrize("rows,cols", [(5, 34000), (16, 8400), (33, 65), (1, 1)])
def test_transpose_matches_reference_kernel(ref, rows, cols):
    rng = np.random.default_rng(rows)
    src = rng.standard_normal((rows, cols)).astype(np.float32)
    out = np.zeros((cols, rows), np.float32)
    assert ref.ref_transpose(_p(src), _p(out), rows, cols) == 0
    assert np.array_equal(out, src.T)


@pytest.mark.parametrize("build", ["ieee", "release"])
@pytest.mark.parametrize("classes,anchors", [(1, 34000), (12, 8500), (3, 37)])
def test_decode_matches_reference_kernel(build, classes, anchors):
    ref = _load("libref_kernels_ieee.so" if build == "ieee" else "libref_kernels.so")   # 0.5 * w is exact: same under both
    rng = np.random.default_rng(classes)
    ch = 4 + classes
    net = rng.uniform(0, 640, (ch, anchors)).astype(np.float32)
    net[4:] = rng.uniform(0, 1, (classes, anchors)).astype(np.float32)
    net[4:, ::7] = net[4, ::7]          # equal scores: the first maximum must win
    net[2:4, ::5] *= 4                  # boxes whose corner clamps at 0
    rows = np.ascontiguousarray(net.T)
    out = np.zeros((anchors, 6), np.float32)
    assert ref.ref_decode(_p(rows), _p(out), ch, anchors, classes) == 0
    assert np.array_equal(out, do.decode(net, classes))


def test_iou_matches_reference_function(ref):
    rng = np.random.default_rng(11)
    a = rng.uniform(0, 100, (200, 4)).astype(np.float32)
    b = rng.uniform(0, 100, (200, 4)).astype(np.float32)
    b[:20] = a[:20]                               # identical boxes
    b[20:40, 0] = a[20:40, 0] + a[20:40, 2]       # touching edges: x_right == x_left -> area 0, not the early-out
    want = np.array([ref.ref_iou(*a[i], *b[i]) for i in range(200)], np.float32)
    got = np.array([do.iou_xywh(a[i:i + 1], b[i:i + 1])[0, 0] for i in range(200)], np.float32)
    assert np.array_equal(got, want)


def test_iou_under_the_release_flags_within_rounding(ref_fast):
    rng = np.random.default_rng(12)
    a = rng.uniform(0, 100, (200, 4)).astype(np.float32)
    b = (a + rng.uniform(-5, 5, (200, 4))).astype(np.float32)
    b[:, 2:] = np.abs(b[:, 2:]) + 1
    want = np.array([do.iou_xywh(a[i:i + 1], b[i:i + 1])[0, 0] for i in range(200)], np.float32)
    got = np.array([ref_fast.ref_iou(*a[i], *b[i]) for i in range(200)], np.float32)
    assert np.allclose(got, want, rtol=0, atol=4e-7)


def _chain_free_scene(rng, n_clusters, per_cluster, n_noise, classes):
    dets = []
    for c in range(n_clusters):
        cx, cy = 60 + 45 * (c % 12), 60 + 45 * (c // 12)
        label = float(rng.integers(0, classes))
        best_conf = rng.uniform(0.8, 0.99)
        dets.append([cx, cy, 30, 30, label, best_conf])
        for _ in range(per_cluster - 1):        # jitter <= 1 px: IoU with the best box > 0.8
            dets.append([cx + rng.uniform(-1, 1), cy + rng.uniform(-1, 1), 30, 30, label, rng.uniform(0.3, 0.79)])
    for _ in range(n_noise):                    # below the score threshold: die whatever the order
        dets.append([rng.uniform(0, 600), rng.uniform(0, 600), 20, 20, float(rng.integers(0, classes)), rng.uniform(0, 0.2)])
    dets = np.array(dets, np.float32)
    return dets[rng.permutation(len(dets))]


@pytest.mark.parametrize("build", ["ieee", "release"])
@pytest.mark.parametrize("seed,n_clusters,per,noise,classes", [(1, 30, 6, 800, 1), (2, 100, 3, 3000, 12), (3, 1, 40, 10, 2)])
def test_nms_matches_reference_kernel_on_chain_free_scenes(build, seed, n_clusters, per, noise, classes):
    ref = _load("libref_kernels_ieee.so" if build == "ieee" else "libref_kernels.so")   # IoUs are far from the threshold
    rng = np.random.default_rng(seed)
    dets = _chain_free_scene(rng, n_clusters, per, noise, classes)
    work = dets.copy()
    assert ref.ref_nms(_p(work), C.c_float(0.65), C.c_float(0.25), len(work)) == 0   # reference defaults, sample_radar.h
    ref_keep = np.nonzero(~np.isnan(work[:, 4]))[0]
    keep = do.nms(dets, 0.65, 0.25)
    assert np.array_equal(keep, ref_keep)
    assert len(keep) == n_clusters
    assert np.array_equal(work[ref_keep], dets[ref_keep])   # survivors are untouched


def test_nms_equal_confidence_tie_matches_reference_kernel(ref):
    """two identical boxes with the same confidence: `comp_conf > row_conf` is false both ways, both survive (B#6)."""
    dets = np.array([[10, 10, 20, 20, 0, 0.9], [10, 10, 20, 20, 0, 0.9], [10, 10, 20, 20, 1, 0.5]], np.float32)
    work = dets.copy()
    assert ref.ref_nms(_p(work), C.c_float(0.65), C.c_float(0.25), 3) == 0
    assert np.array_equal(np.nonzero(~np.isnan(work[:, 4]))[0], do.nms(dets, 0.65, 0.25))
    assert len(do.nms(dets, 0.65, 0.25)) == 3
