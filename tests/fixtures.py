"""Shared fixtures for tests / smoke / bench: the sample's calibration, golden assets, synthetic
generators and comparison helpers.  Nothing here reads /root/reference at run time
(it does not exist on the GPU box); tests/golden/make_golden.py is the only script that does."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
ENGINES = os.path.join(ROOT, "rm_radar_b200", "engines")

# calibration of the sample — /root/reference/samples/main.cpp:12-21
IMAGE_SIZE = (2592, 2048)
INTRINSIC = np.array([1685.51538398561, 0, 1278.99324114319, 0, 1685.26471848220, 1037.21273138299, 0, 0, 1],
                     np.float32)
LIDAR_TO_CAMERA = np.array([0, -1, 0, 0.85443, 0, 0, -1, -37.6845, 1, 0, 0, 12.2631, 0.0, 0.0, 0.0, 1.0], np.float32)
WORLD_TO_CAMERA = np.array([0.05975021, 0.99807031, 0.01689906, -7179.65399136, 0.28962566, -0.00113262,
                            -0.95713933, -4671.34956587, -0.9552732, 0.06208368, -0.28913445, 28286.8920291,
                            0.0, 0.0, 0.0, 1.0], np.float32)
# sample_radar.h:32-34
CLASS_NUM, MAX_BATCH, OPT_BATCH = 12, 20, 4


def scaled_intrinsic(w, h):
    """fx, cx scaled by W/2592; fy, cy by H/2048."""
    K = INTRINSIC.copy()
    K[0] *= w / 2592.0; K[2] *= w / 2592.0
    K[4] *= h / 2048.0; K[5] *= h / 2048.0
    return K.astype(np.float32)


def engine(name):
    p = os.path.join(ENGINES, name + ".rmeng")
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} missing: run __graft_entry__.build() where /root/reference is mounted")
    return p


def onnx(name):
    return os.path.join(ENGINES, name + ".onnx")


def have_models():
    return all(os.path.exists(os.path.join(ENGINES, n)) for n in ("car.rmeng", "armor.rmeng"))


def have_onnx():
    """fp32 ONNX copies for the live torch oracle (114 MB; may be left out of a GPU-box snapshot)."""
    return all(os.path.exists(os.path.join(ENGINES, n)) for n in ("car.onnx", "armor.onnx"))


def load_frame(i):
    import cv2
    img = cv2.imread(os.path.join(GOLDEN, "frames", f"{i}.jpg"), cv2.IMREAD_COLOR)
    assert img is not None
    return img


def load_clouds():
    z = np.load(os.path.join(GOLDEN, "clouds.npz"))
    return {k: z[k] for k in z.files}


def resize_frame(img, w, h):
    """Resize an asset frame with the oracle's own resize so that real robots stay in view (§8d)."""
    from oracle import detect_oracle as do
    return do.resize(img, w, h)


# ---------------------------------------------------------------------------------------------
# synthetic LiDAR scene: static background (ground + far wall) and R boxes in front
# ---------------------------------------------------------------------------------------------
def synthetic_scene(n_points, seed, n_boxes=8, w=1920, h=1080, dense_factor=10):
    """Returns (background cloud, frame cloud, boxes as full-res pixel rects [x,y,w,h]) in the
    LiDAR frame of the sample calibration (x forward, mm)."""
    rng = np.random.default_rng(seed)
    K = scaled_intrinsic(w, h).reshape(3, 3).astype(np.float64)
    L = LIDAR_TO_CAMERA.reshape(4, 4).astype(np.float64)

    def rays(n):
        # uniform over the camera field of view, expressed in camera pixel space then back to lidar
        u = rng.uniform(0, w, n); v = rng.uniform(0, h, n)
        return u, v

    def backproject(u, v, depth):
        cam = np.stack([(u - K[0, 2]) / K[0, 0] * depth, (v - K[1, 2]) / K[1, 1] * depth, depth], 1)
        Linv = np.linalg.inv(L)
        return (cam @ Linv[:3, :3].T + Linv[:3, 3]).astype(np.float32)

    def background_depth(u, v):
        # far wall at 22-27 m with a slow gradient, floor-like ramp in the lower half
        d = 24000 + 2000 * np.sin(u / w * 3.0) + 1000 * (v / h)
        return d

    ub, vb = rays(n_points * dense_factor)
    bg = backproject(ub, vb, background_depth(ub, vb))
    uf, vf = rays(n_points)
    depth = background_depth(uf, vf)
    boxes = []
    for b in range(n_boxes):
        bw, bh = rng.uniform(60, 160), rng.uniform(60, 160)
        bx, by = rng.uniform(0, w - bw), rng.uniform(0, h - bh)
        inside = (uf >= bx) & (uf < bx + bw) & (vf >= by) & (vf < by + bh)
        depth[inside] = background_depth(uf[inside], vf[inside]) - rng.uniform(1000, 3000) + rng.normal(0, 30, inside.sum())
        boxes.append((bx, by, bw, bh))
    fr = backproject(uf, vf, depth)
    return bg, fr, np.asarray(boxes, np.float32)


def iou_xywh(a, b):
    ax2, ay2, bx2, by2 = a[0] + a[2], a[1] + a[3], b[0] + b[2], b[1] + b[3]
    iw = max(0.0, min(ax2, bx2) - max(a[0], b[0])); ih = max(0.0, min(ay2, by2) - max(a[1], b[1]))
    inter = iw * ih
    return inter / (a[2] * a[3] + b[2] * b[3] - inter + 1e-12)


def match_detections(got, ref, min_iou=0.99, conf_tol=5e-3):
    """BASELINE gate: same count, class-exact, IoU >= 0.99 pairwise in order."""
    assert len(got) == len(ref), f"count {len(got)} != {len(ref)}"
    for g, r in zip(got, ref):
        assert int(g[4]) == int(r[4]), f"class {g[4]} != {r[4]}"
        i = iou_xywh(g, r)
        assert i >= min_iou, f"IoU {i:.4f} < {min_iou} ({g} vs {r})"
        assert abs(g[5] - r[5]) <= conf_tol, f"conf {g[5]} vs {r[5]}"


def run_smoke(verbose=False):
    """One detect + locate pass of the hot path on cuda:0, checked against the oracle on the golden
    frame (fp32 ONNX oracle when the model copies travelled, else the committed expected values)."""
    import rm_radar_b200 as rr
    from oracle import locate_oracle as lo
    exp = np.load(os.path.join(GOLDEN, "expected.npz"), allow_pickle=True)
    img = load_frame(0)
    clouds = load_clouds()
    det = rr.RobotDetector(engine("car"), engine("armor"), IMAGE_SIZE, CLASS_NUM, MAX_BATCH, OPT_BATCH)
    loc = rr.Locator(IMAGE_SIZE[0], IMAGE_SIZE[1], INTRINSIC, LIDAR_TO_CAMERA, WORLD_TO_CAMERA)
    ora = lo.LocatorOracle(IMAGE_SIZE[0], IMAGE_SIZE[1], INTRINSIC, LIDAR_TO_CAMERA, WORLD_TO_CAMERA)
    for c in (clouds["background"], clouds["c0"]):
        loc.update(c); ora.update(c)
    loc.cluster(); ora.cluster()
    # the frame enters as the JPEG file the reference would cv::imread: decoded on the device, checked against the oracle
    from oracle import jpeg_oracle as jo
    jpg = open(os.path.join(GOLDEN, "frames", "0.jpg"), "rb").read()
    dec = rr.JpegDecoder(0)
    assert np.array_equal(dec.decode(jpg), jo.decode(jpg)), "device JPEG decode differs from the oracle"
    assert np.array_equal(img, jo.decode(jpg)), "oracle JPEG decode differs from cv2.imread"
    robots = det.detect_jpeg(dec, jpg)
    loc.search(robots)
    cars = [d.as_array() for d in det.last_cars()]
    match_detections(cars, exp["f0_cars"])
    assert np.array_equal(loc.image("diff"), ora.diff), "diff image differs from the oracle"
    assert np.array_equal(loc.image("labels"), ora.label_image), "cluster labels differ from the oracle"
    want = ora.search([r.rect for r in robots])
    n_loc = 0
    for r, w in zip(robots, want):
        assert (r.location is None) == (w is None)
        if w is not None:
            n_loc += 1
            assert np.abs(np.asarray(r.location) - w).max() < 1e-3, (r.location, w)
    labels = sorted(int(r.label) for r in robots if r.isDetected())
    assert labels == sorted(int(x) for x in exp["f0_robot_labels"]), (labels, exp["f0_robot_labels"])
    if verbose:
        print(f"smoke ok: {len(cars)} cars, robot labels {labels}, {n_loc} located, "
              f"{loc.stats()['clusters']} clusters; stats {det.last_stats()}")
