"""ORACLE (test infrastructure only — never imported by the product path).

numpy restatement of the reference's detect path, function by function, each citing the
reference lines it follows.  Pinned against the reference's own golden vectors in
tests/test_oracle_detect.py (ResizeDouble / ResizeHalf / CopyMakeBorder truth tables from
`/root/reference/test/detect/kernel_test.cu:71-139`, the Blob ≡ cv::dnn::blobFromImage
definition `:141-173`, the letterbox numbers of `test/detect/detector_test.cpp:38-67`).
The network boundary (TensorRT) and NMS/decode have no reference vectors: "parity unpinned"
there (see oracle/onnx_torch.py header and DESIGN.md).

All arithmetic is float32 unless the reference promotes (the `0.5 *` in decodeKernel is double).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------------------
# PreParam — /root/reference/src/detect/preparam.h:46-52
# --------------------------------------------------------------------------------------
@dataclass
class PreParam:
    width: np.float32
    height: np.float32
    ratio: np.float32
    dw: np.float32
    dh: np.float32


def c_round(x) -> np.float32:
    """std::round: half away from zero (numpy's round is half-to-even)."""
    x = f32(x)
    return f32(np.trunc(x + np.copysign(f32(0.5), x))) if np.isfinite(x) else x


def preparam(in_w: int, in_h: int, out_w: int = 640, out_h: int = 640) -> PreParam:
    height = f32(in_h)
    width = f32(in_w)
    ratio = f32(1) / min(f32(out_h) / height, f32(out_w) / width)
    ratio = f32(ratio)
    dw = f32((f32(out_w) - c_round(width / ratio)) * f32(0.5))
    dh = f32((f32(out_h) - c_round(height / ratio)) * f32(0.5))
    return PreParam(width, height, ratio, dw, dh)


# --------------------------------------------------------------------------------------
# resizeKernel — /root/reference/src/detect/detector.cu:40-81
# top-left aligned bilinear, float blend in the order tl+tr+bl+br, truncating cast.
# --------------------------------------------------------------------------------------
def resize(src: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    src_h, src_w, ch = src.shape
    dy = np.arange(dst_h, dtype=f32)
    dx = np.arange(dst_w, dtype=f32)
    # `dst_y * static_cast<float>(src_h) / dst_h` evaluates left to right in float
    sy = (dy * f32(src_h)) / f32(dst_h)
    sx = (dx * f32(src_w)) / f32(dst_w)
    y0 = sy.astype(np.int32)
    x0 = sx.astype(np.int32)
    y1 = np.minimum(y0 + 1, src_h - 1)
    x1 = np.minimum(x0 + 1, src_w - 1)
    ly = (sy - y0.astype(f32)).astype(f32)[:, None, None]
    lx = (sx - x0.astype(f32)).astype(f32)[None, :, None]
    hy = (f32(1) - ly).astype(f32)
    hx = (f32(1) - lx).astype(f32)
    s = src.astype(f32)
    tl = s[y0][:, x0] * hy * hx
    tr = s[y0][:, x1] * hy * lx
    bl = s[y1][:, x0] * ly * hx
    br = s[y1][:, x1] * ly * lx
    val = ((tl + tr) + bl) + br
    return val.astype(np.uint8)  # static_cast<unsigned char>: truncation (values are in [0,255])


# --------------------------------------------------------------------------------------
# copyMakeBorderKernel — /root/reference/src/detect/detector.cu:102-133
# --------------------------------------------------------------------------------------
def copy_make_border(src: np.ndarray, top: int, bottom: int, left: int, right: int) -> np.ndarray:
    h, w, ch = src.shape
    dst = np.full((h + top + bottom, w + left + right, ch), 128, np.uint8)
    dst[top:top + h, left:left + w] = src
    return dst


# --------------------------------------------------------------------------------------
# blobKernel — /root/reference/src/detect/detector.cu:151-171  (BGR u8 HWC → RGB f32 CHW × scale)
# --------------------------------------------------------------------------------------
def blob(src: np.ndarray, scale=f32(1 / 255.0)) -> np.ndarray:
    return (src[:, :, ::-1].transpose(2, 0, 1).astype(f32) * f32(scale)).astype(f32)


# --------------------------------------------------------------------------------------
# Detector::preprocess call sites — /root/reference/src/detect/detector.cu:380-421, 439-502
#
# compat=True reproduces the float→int truncation of padding_width/height at the kernel call
# sites: the three kernels run on a persistent 640*640*3 u8 border
# buffer with the *actual* (possibly 639-wide) stride, and blobKernel reads it with stride 640.
# `border_buf` is that persistent buffer (previous contents survive where nothing is written;
# a fresh Detector starts from zeros).  compat=False letterboxes to the rounded size.
# --------------------------------------------------------------------------------------
def letterbox_u8(image: np.ndarray, out_w=640, out_h=640, compat=True, border_buf=None):
    pp = preparam(image.shape[1], image.shape[0], out_w, out_h)
    pad_w_f = f32(pp.width / pp.ratio)
    pad_h_f = f32(pp.height / pp.ratio)
    top = int(c_round(f32(np.float64(pp.dh) - 0.1)))      # std::round(pparam.dh - 0.1): double math
    bottom = int(c_round(f32(np.float64(pp.dh) + 0.1)))
    left = int(c_round(f32(np.float64(pp.dw) - 0.1)))
    right = int(c_round(f32(np.float64(pp.dw) + 0.1)))
    if compat:
        pw, ph = int(pad_w_f), int(pad_h_f)  # float → int parameter: truncation
    else:
        pw, ph = int(c_round(pad_w_f)), int(c_round(pad_h_f))
    pw = max(pw, 1)
    ph = max(ph, 1)
    resized = resize(image, pw, ph)
    if border_buf is None:
        border_buf = np.zeros(out_w * out_h * 3, np.uint8)
    bordered = copy_make_border(resized, top, bottom, left, right)
    bh, bw = bordered.shape[:2]
    # the border kernel's grid covers 640x640 threads but only dst_h x dst_w are written,
    # with dst_step = bw*3 into the flat buffer
    bh_c, bw_c = min(bh, out_h), min(bw, out_w)
    flat = bordered[:bh_c, :bw_c].reshape(bh_c, bw_c * 3)
    if bw == out_w:
        n = bh_c * out_w * 3
        border_buf[:n] = flat.reshape(-1)
    else:
        for y in range(bh_c):
            o = y * bw * 3
            border_buf[o:o + bw_c * 3] = flat[y]
    return border_buf.reshape(out_h, out_w, 3), pp


def preprocess(image: np.ndarray, out_w=640, out_h=640, compat=True, border_buf=None):
    """→ (f32 [3,out_h,out_w] RGB/255, PreParam)."""
    u8, pp = letterbox_u8(image, out_w, out_h, compat, border_buf)
    return blob(u8), pp


# --------------------------------------------------------------------------------------
# decodeKernel — /root/reference/src/detect/detector.cu:219-251   (after transposeKernel :185-203)
# --------------------------------------------------------------------------------------
def decode(net_out: np.ndarray, classes: int) -> np.ndarray:
    """net_out [4+classes, A] f32 → [A, 6] f32 (x, y, w, h, label, conf)."""
    rows = np.ascontiguousarray(net_out.T.astype(f32))  # transposeKernel
    box = rows[:, :4]
    scores = rows[:, 4:4 + classes]
    label = np.argmax(scores, axis=1)  # first maximum wins (strict > in the reference loop)
    conf = scores[np.arange(rows.shape[0]), label]
    x = np.maximum(box[:, 0].astype(np.float64) - 0.5 * box[:, 2].astype(np.float64), 0.0).astype(f32)
    y = np.maximum(box[:, 1].astype(np.float64) - 0.5 * box[:, 3].astype(np.float64), 0.0).astype(f32)
    out = np.stack([x, y, box[:, 2], box[:, 3], label.astype(f32), conf], axis=1).astype(f32)
    return out


# --------------------------------------------------------------------------------------
# IoU — /root/reference/src/detect/detector.cu:271-293
# --------------------------------------------------------------------------------------
def iou_xywh(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """a [n,4], b [m,4] → [n,m] f32, float32 arithmetic in the reference's order."""
    ax, ay, aw, ah = (a[:, None, i].astype(f32) for i in range(4))
    bx, by, bw, bh = (b[None, :, i].astype(f32) for i in range(4))
    xl = np.maximum(ax, bx)
    yt = np.maximum(ay, by)
    xr = np.minimum(ax + aw, bx + bw)
    yb = np.minimum(ay + ah, by + bh)
    inter = ((xr - xl) * (yb - yt)).astype(f32)
    union = ((aw * ah + bw * bh) - inter).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = (inter / union).astype(f32)
    return np.where((xr < xl) | (yb < yt), f32(0), v)


# --------------------------------------------------------------------------------------
# NMSKernel — /root/reference/src/detect/detector.cu:315-360, race-free snapshot semantics
# (SURVEY.md Appendix B#5/#6): row dies if conf < score_thresh, or if ANY same-label column with
# strictly greater conf has IoU > nms_thresh (columns are never filtered by score: a column
# below the threshold has conf < every passing row, so it cannot suppress one).
# Returns indices (anchor order) of surviving rows.
# --------------------------------------------------------------------------------------
def nms(dets: np.ndarray, nms_thresh, score_thresh) -> np.ndarray:
    conf = dets[:, 5]
    cand = np.nonzero(~(conf < f32(score_thresh)))[0]
    if cand.size == 0:
        return cand
    d = dets[cand]
    ious = iou_xywh(d[:, :4], d[:, :4])
    same = d[:, None, 4] == d[None, :, 4]
    higher = d[None, :, 5] > d[:, None, 5]
    killed = (same & higher & (ious > f32(nms_thresh))).any(axis=1)
    return cand[~killed]


# --------------------------------------------------------------------------------------
# restoreDetection — /root/reference/src/detect/detector.cpp:258-268
# --------------------------------------------------------------------------------------
def restore(det: np.ndarray, pp: PreParam) -> np.ndarray:
    d = det.astype(f32).copy()
    d[:, 0] = np.clip((d[:, 0] - pp.dw) * pp.ratio, f32(0), pp.width)
    d[:, 1] = np.clip((d[:, 1] - pp.dh) * pp.ratio, f32(0), pp.height)
    # std::clamp(v, 0, hi) with hi < 0 is UB in C++; cannot happen since x <= width
    d[:, 2] = np.minimum(np.maximum(d[:, 2] * pp.ratio, f32(0)), pp.width - d[:, 0])
    d[:, 3] = np.minimum(np.maximum(d[:, 3] * pp.ratio, f32(0)), pp.height - d[:, 1])
    return d.astype(f32)


def postprocess(net_out: np.ndarray, classes: int, pp: PreParam, nms_thresh, conf_thresh) -> np.ndarray:
    """Detector::postprocess — /root/reference/src/detect/detector.cu:522-582. → [n,6] anchor order."""
    dets = decode(net_out, classes)
    keep = nms(dets, nms_thresh, conf_thresh)
    return restore(dets[keep], pp)


# --------------------------------------------------------------------------------------
# computeIoU (dedup) — /root/reference/src/detect/detector.cpp:324-349  (intersection / bounding rect)
# --------------------------------------------------------------------------------------
def compute_iou_bounding(r1, r2) -> np.float32:
    r1 = [f32(v) for v in r1]
    r2 = [f32(v) for v in r2]
    x1 = max(r1[0], r2[0]); y1 = max(r1[1], r2[1])
    x2 = min(f32(r1[0] + r1[2]), f32(r2[0] + r2[2])); y2 = min(f32(r1[1] + r1[3]), f32(r2[1] + r2[3]))
    if x1 < x2 and y1 < y2:
        iw, ih = f32(x2 - x1), f32(y2 - y1)
    else:
        iw = ih = f32(0)
    ux1 = min(r1[0], r2[0]); uy1 = min(r1[1], r2[1])
    ux2 = max(f32(r1[0] + r1[2]), f32(r2[0] + r2[2])); uy2 = max(f32(r1[1] + r1[3]), f32(r2[1] + r2[3]))
    ia = f32(iw * ih)
    ua = f32(f32(ux2 - ux1) * f32(uy2 - uy1))
    return f32(ia / ua) if ua > 0 else f32(0)


# --------------------------------------------------------------------------------------
# Robot::setDetection — /root/reference/src/robot/robot.cpp:41-74
# --------------------------------------------------------------------------------------
@dataclass
class Robot:
    rect: tuple | None = None            # Rect2f (x, y, w, h)
    label: int | None = None
    confidence: np.float32 | None = None
    armors: np.ndarray | None = None     # [n,6] in frame coordinates
    location: np.ndarray | None = None   # metres, world

    def is_detected(self) -> bool:
        # Robot::isDetected — /root/reference/src/robot/robot.h:65 (armors_.has_value())
        return self.armors is not None


def rect_int(rect):
    """Robot::rect() — /root/reference/src/robot/robot.h:111: optional<Rect2f> → optional<cv::Rect>,
    i.e. cv::saturate_cast<int>(float) = cvRound = round-half-to-even on each of x, y, w, h."""
    return tuple(int(np.rint(f32(v))) for v in rect)


def set_detection(car: np.ndarray, armors: np.ndarray) -> Robot:
    r = Robot(rect=(f32(car[0]), f32(car[1]), f32(car[2]), f32(car[3])))
    if armors.shape[0] == 0:
        return r
    score = {}
    for a in armors:
        k = int(a[4])
        score[k] = f32(score.get(k, f32(0)) + f32(a[5]))
    label = None
    best = None
    for k in sorted(score):  # std::map order; max_element keeps the first maximum
        if best is None or best < score[k]:
            best, label = score[k], k
    count = int(sum(1 for a in armors if a[4] == label))
    r.label = label
    r.confidence = f32(best / f32(count))
    arm = armors.astype(f32).copy()
    arm[:, 0] += f32(car[0])
    arm[:, 1] += f32(car[1])
    r.armors = arm
    return r


# --------------------------------------------------------------------------------------
# RobotDetector::detect — /root/reference/src/detect/detector.cpp:413-455
# --------------------------------------------------------------------------------------
@dataclass
class CascadeTrace:
    car_dets: np.ndarray = None
    rois: list = field(default_factory=list)
    armor_dets: list = field(default_factory=list)
    car_net_out: np.ndarray = None
    armor_net_out: list = field(default_factory=list)


def robot_detect(image: np.ndarray, car_net, armor_net, armor_classes=12, max_cars=20, iou_thresh=0.75,
                 car_nms=0.65, car_conf=0.25, armor_nms=0.65, armor_conf=0.50, compat=True,
                 car_border=None, armor_borders=None, trace: CascadeTrace | None = None):
    """car_net / armor_net: callables f32 [B,3,640,640] → [B,4+nc,A] (oracle.onnx_torch.OnnxNet).

    `car_border`, `armor_borders[i]`: persistent u8 staging buffers of the reference Detector
    objects (compat mode); None = fresh detector (zeros)."""
    x, pp = preprocess(image, compat=compat, border_buf=car_border)
    out = np.asarray(car_net(x[None]))[0]
    cars = postprocess(out, 1, pp, car_nms, car_conf)
    # B#8: more cars than max_batch is UB in the reference; the build keeps the first max_cars
    cars = cars[:max_cars]
    if trace is not None:
        trace.car_dets = cars
        trace.car_net_out = out
    armor_dets = []
    if cars.shape[0] > 0:
        blobs, pps = [], []
        for i, c in enumerate(cars):
            # cv::Rect(float, float, float, float): truncation  (detector.cpp:420-421)
            rx, ry, rw, rh = int(c[0]), int(c[1]), int(c[2]), int(c[3])
            roi = image[ry:ry + rh, rx:rx + rw]
            if trace is not None:
                trace.rois.append((rx, ry, rw, rh))
            if rw <= 0 or rh <= 0:
                blobs.append(None); pps.append(None)
                continue
            bb = armor_borders[i] if armor_borders is not None else None
            b, p = preprocess(roi, compat=compat, border_buf=bb)
            blobs.append(b); pps.append(p)
        valid = [i for i, b in enumerate(blobs) if b is not None]
        outs = {}
        if valid:
            o = np.asarray(armor_net(np.stack([blobs[i] for i in valid])))
            for j, i in enumerate(valid):
                outs[i] = o[j]
        for i in range(cars.shape[0]):
            if i in outs:
                armor_dets.append(postprocess(outs[i], armor_classes, pps[i], armor_nms, armor_conf))
                if trace is not None:
                    trace.armor_net_out.append(outs[i])
            else:
                armor_dets.append(np.zeros((0, 6), f32))
    if trace is not None:
        trace.armor_dets = armor_dets
    robots = []
    by_label = {}
    for i in range(cars.shape[0]):
        robot = set_detection(cars[i], armor_dets[i])
        if not robot.is_detected():
            robots.append(robot)
            continue
        lab = robot.label
        if lab not in by_label:
            by_label[lab] = robot
        else:
            ex = by_label[lab]
            if compute_iou_bounding(rect_int(ex.rect), rect_int(robot.rect)) > f32(iou_thresh):
                continue
            elif ex.confidence < robot.confidence:
                by_label[lab] = robot
    for lab in sorted(by_label):
        robots.append(by_label[lab])
    return robots
