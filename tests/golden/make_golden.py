"""Generates the committed fixtures under tests/golden/ from the reference's assets and the oracle.
Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
  frames/0.jpg, frames/5.jpg   verbatim copies of assets/images (input data, not source)
  clouds.npz                    assets/clouds 0..3 + every 4th point of background.pcd
  expected.npz                  oracle outputs (fp32 ONNX via oracle/onnx_torch.py, compat letterbox)
  jpeg/*.jpg, jpeg/expected.npz what cv2.imdecode (= the reference's cv::imread) returns for small synthetic files
                                (4:4:4 / 4:2:2 / 4:2:0 / grayscale, restart intervals, optimised tables, odd sizes)
                                and sha256 of the decoded reference frames 0 and 5
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import detect_oracle as do  # noqa: E402
from oracle import locate_oracle as lo  # noqa: E402
from oracle.onnx_torch import OnnxNet  # noqa: E402
from tests import fixtures as fx  # noqa: E402


def make_jpeg_fixtures():
    """Golden vectors for the JPEG stage: outputs of cv2.imdecode, the call the reference makes (samples/main.cpp:24-40)."""
    import hashlib

    import cv2
    d = os.path.join(HERE, "jpeg")
    os.makedirs(d, exist_ok=True)
    src = cv2.imread(os.path.join(HERE, "frames", "0.jpg"), cv2.IMREAD_COLOR)
    rng = np.random.default_rng(7)
    S = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
         "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420}
    cases = [("photo_420_q90", (97, 61), "420", 90, 0, 0, False), ("photo_444_q75_opt", (97, 61), "444", 75, 0, 1, False),
             ("photo_422_q50_rst3", (130, 47), "422", 50, 3, 0, False), ("photo_420_q95_rst1_opt", (64, 48), "420", 95, 1, 1, False),
             ("noise_420_q100", (33, 17), "420", 100, 0, 0, True), ("noise_444_q30_rst2", (40, 24), "444", 30, 2, 0, True),
             ("tiny_420", (3, 2), "420", 90, 0, 0, False), ("gray_q80", (75, 50), None, 80, 0, 0, False)]
    out = {}
    for name, (w, h), samp, q, rst, opt, noise in cases:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8) if noise else \
            cv2.resize(src[500:1500, 800:2200], (w, h), interpolation=cv2.INTER_AREA)
        flags = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_RST_INTERVAL, rst, cv2.IMWRITE_JPEG_OPTIMIZE, opt]
        if samp is None:
            img = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        else:
            flags += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, S[samp]]
        ok, enc = cv2.imencode(".jpg", img, flags)
        assert ok
        open(os.path.join(d, name + ".jpg"), "wb").write(enc.tobytes())
        out[name] = cv2.imdecode(enc, cv2.IMREAD_COLOR)
    for i in (0, 5):
        img = cv2.imread(os.path.join(HERE, "frames", f"{i}.jpg"), cv2.IMREAD_COLOR)
        out[f"frame{i}_sha256"] = np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8)
        out[f"frame{i}_shape"] = np.asarray(img.shape, np.int32)
    np.savez_compressed(os.path.join(d, "expected.npz"), **out)
    print("jpeg fixtures:", sorted(out))


def main():
    import cv2
    os.makedirs(os.path.join(HERE, "frames"), exist_ok=True)
    for i in (0, 5):
        shutil.copyfile(f"{REF}/assets/images/{i}.jpg", os.path.join(HERE, "frames", f"{i}.jpg"))
    clouds = {f"c{i}": lo.read_pcd(f"{REF}/assets/clouds/{i}.pcd") for i in range(4)}
    clouds["c5"] = lo.read_pcd(f"{REF}/assets/clouds/5.pcd")
    clouds["background"] = lo.read_pcd(f"{REF}/assets/clouds/background.pcd")[::4].copy()
    np.savez_compressed(os.path.join(HERE, "clouds.npz"), **clouds)

    car = OnnxNet(f"{REF}/models/car.onnx")
    armor = OnnxNet(f"{REF}/models/armor.onnx")
    out = {}
    for i in (0, 5):
        img = cv2.imread(os.path.join(HERE, "frames", f"{i}.jpg"), cv2.IMREAD_COLOR)
        tr = do.CascadeTrace()
        robots = do.robot_detect(img, lambda x: car(x).numpy(), lambda x: armor(x).numpy(), trace=tr)
        out[f"f{i}_cars"] = tr.car_dets
        out[f"f{i}_rois"] = np.asarray(tr.rois, np.int32)
        out[f"f{i}_armor_counts"] = np.asarray([len(a) for a in tr.armor_dets], np.int32)
        out[f"f{i}_armors"] = np.concatenate(tr.armor_dets) if tr.armor_dets else np.zeros((0, 6), np.float32)
        out[f"f{i}_robot_labels"] = np.asarray([r.label for r in robots if r.is_detected()], np.int32)
        out[f"f{i}_robot_conf"] = np.asarray([r.confidence for r in robots if r.is_detected()], np.float32)
        out[f"f{i}_robot_rects"] = np.asarray([r.rect for r in robots], np.float32)
        print(i, "cars", len(tr.car_dets), "armors", out[f"f{i}_armor_counts"], "labels", out[f"f{i}_robot_labels"])
        if i == 0:
            # 1920x1080 variant of frame 0 (BASELINE config C2 geometry), resized with the oracle's resize
            small = do.resize(img, 1920, 1080)
            tr2 = do.CascadeTrace()
            robots2 = do.robot_detect(small, lambda x: car(x).numpy(), lambda x: armor(x).numpy(), trace=tr2)
            out["f0_1080_cars"] = tr2.car_dets
            out["f0_1080_armor_counts"] = np.asarray([len(a) for a in tr2.armor_dets], np.int32)
            out["f0_1080_armors"] = np.concatenate(tr2.armor_dets) if tr2.armor_dets else np.zeros((0, 6), np.float32)
            out["f0_1080_robot_labels"] = np.asarray([r.label for r in robots2 if r.is_detected()], np.int32)
            print("1080p cars", len(tr2.car_dets), "armors", out["f0_1080_armor_counts"], out["f0_1080_robot_labels"])
    # locate: background (subsampled) then cloud 0, boxes = frame-0 cars
    loc = lo.LocatorOracle(fx.IMAGE_SIZE[0], fx.IMAGE_SIZE[1], fx.INTRINSIC, fx.LIDAR_TO_CAMERA, fx.WORLD_TO_CAMERA)
    loc.update(clouds["background"]); loc.update(clouds["c0"]); loc.cluster()
    res = loc.search([tuple(c[:4]) for c in out["f0_cars"]])
    out["f0_located"] = np.asarray([r is not None for r in res])
    out["f0_locations"] = np.asarray([r if r is not None else (np.nan,) * 3 for r in res], np.float64)
    out["f0_fg_clusters"] = np.asarray([len(loc.fg_points), loc.num_clusters], np.int32)
    print("locate", out["f0_fg_clusters"], out["f0_locations"])
    np.savez_compressed(os.path.join(HERE, "expected.npz"), **out)


def make_all_assets():
    """expected_all.npz: the oracle on ALL ten frames and ten clouds of the reference (north_star: "the same
    assets/images + assets/clouds inputs"); frame i is paired with cloud i, boxes = that frame's cars.  The eight extra
    frames and the extra clouds are input data, copied by __graft_entry__.build() into the git-ignored
    tests/golden/_assets/ (they travel to the GPU box with the snapshot); only the oracle's outputs are committed."""
    import cv2
    car = OnnxNet(f"{REF}/models/car.onnx")
    armor = OnnxNet(f"{REF}/models/armor.onnx")
    out = {}
    bg = lo.read_pcd(f"{REF}/assets/clouds/background.pcd")[::4].copy()
    for i in range(10):
        img = cv2.imread(f"{REF}/assets/images/{i}.jpg", cv2.IMREAD_COLOR)
        tr = do.CascadeTrace()
        robots = do.robot_detect(img, lambda x: car(x).numpy(), lambda x: armor(x).numpy(), trace=tr)
        out[f"f{i}_cars"] = tr.car_dets
        out[f"f{i}_armor_counts"] = np.asarray([len(a) for a in tr.armor_dets], np.int32)
        out[f"f{i}_armors"] = np.concatenate(tr.armor_dets) if tr.armor_dets else np.zeros((0, 6), np.float32)
        out[f"f{i}_robot_labels"] = np.asarray([r.label for r in robots if r.is_detected()], np.int32)
        out[f"f{i}_robot_conf"] = np.asarray([r.confidence for r in robots if r.is_detected()], np.float32)
        out[f"f{i}_robot_rects"] = np.asarray([r.rect for r in robots], np.float32)
        # locate: a fresh Locator per pair (background, then cloud i), boxes = the frame's robots in output order
        cloud = lo.read_pcd(f"{REF}/assets/clouds/{i}.pcd")
        loc = lo.LocatorOracle(fx.IMAGE_SIZE[0], fx.IMAGE_SIZE[1], fx.INTRINSIC, fx.LIDAR_TO_CAMERA, fx.WORLD_TO_CAMERA)
        loc.update(bg); loc.update(cloud); loc.cluster()
        res = loc.search([tuple(r.rect) for r in robots])
        out[f"f{i}_located"] = np.asarray([r is not None for r in res])
        out[f"f{i}_locations"] = np.asarray([r if r is not None else (np.nan,) * 3 for r in res], np.float64)
        out[f"f{i}_fg_clusters"] = np.asarray([len(loc.fg_points), loc.num_clusters], np.int32)
        print(i, "cars", len(tr.car_dets), "armors", out[f"f{i}_armor_counts"].tolist(), "labels", out[f"f{i}_robot_labels"].tolist(),
              "located", int(out[f"f{i}_located"].sum()), "fg/clusters", out[f"f{i}_fg_clusters"].tolist(), flush=True)
    np.savez_compressed(os.path.join(HERE, "expected_all.npz"), **out)


def copy_assets(dst=None):
    """Input data of the all-assets tests: frames 0-9 and clouds 0-9 + background (subsampled like clouds.npz)."""
    dst = dst or os.path.join(HERE, "_assets")
    os.makedirs(dst, exist_ok=True)
    for i in range(10):
        if not os.path.exists(os.path.join(dst, f"{i}.jpg")):
            shutil.copyfile(f"{REF}/assets/images/{i}.jpg", os.path.join(dst, f"{i}.jpg"))
    cl = os.path.join(dst, "clouds_all.npz")
    if not os.path.exists(cl):
        clouds = {f"c{i}": lo.read_pcd(f"{REF}/assets/clouds/{i}.pcd") for i in range(10)}
        clouds["background"] = lo.read_pcd(f"{REF}/assets/clouds/background.pcd")[::4].copy()
        np.savez_compressed(cl, **clouds)


def make_pcd_fixtures():
    """tests/golden/pcd/: a 1500-point prefix of assets/clouds/0.pcd (ASCII), an ASCII file exercising the number
    grammar (decimals, exponents, signs, CRLF, an extra column, nan / inf, missing final newline) and a binary file
    whose 16-byte records carry a leading intensity field."""
    out_dir = os.path.join(HERE, "pcd")
    os.makedirs(out_dir, exist_ok=True)
    lines = open(f"{REF}/assets/clouds/0.pcd", "rb").read().split(b"\n")
    hdr_end = [i for i, ln in enumerate(lines) if ln.startswith(b"DATA")][0]
    n = 1500
    hdr = [(b"WIDTH %d" % n if ln.startswith(b"WIDTH") else b"POINTS %d" % n if ln.startswith(b"POINTS") else ln)
           for ln in lines[:hdr_end + 1]]
    open(os.path.join(out_dir, "asset0_head1500_ascii.pcd"), "wb").write(
        b"\n".join(hdr + lines[hdr_end + 1:hdr_end + 1 + n]) + b"\n")
    rng = np.random.default_rng(3)
    pts = rng.uniform(-30000, 30000, (400, 3))
    rows = []
    for i, (x, y, z) in enumerate(pts):
        fmt = [("%.3f", "%.6g", "%d"), ("%e", "%.1f", "%+.4f"), ("%.9g", "%.2e", "%.0f")][i % 3]
        rows.append(" ".join(f % (v if "%d" not in f else int(v)) for f, v in zip(fmt, (x, y, z))) + " %d" % (i % 255))
    rows[5] = "nan 1 2 0"
    rows[6] = "-inf 0.5 +7 1"
    rows[7] = "  12.5\t-3e2   0004.250   9"
    txt = ("# .PCD v0.7 - Point Cloud Data file format\r\nVERSION 0.7\r\nFIELDS x y z intensity\r\nSIZE 4 4 4 4\r\n"
           "TYPE F F F U\r\nCOUNT 1 1 1 1\r\nWIDTH 400\r\nHEIGHT 1\r\nVIEWPOINT 0 0 0 1 0 0 0\r\nPOINTS 400\r\n"
           "DATA ascii\r\n" + "\r\n".join(rows))
    open(os.path.join(out_dir, "variants_crlf_ascii.pcd"), "wb").write(txt.encode())
    n = 1000
    xyz = rng.uniform(-30000, 30000, (n, 3)).astype("<f4")
    rec = np.zeros(n, dtype=[("intensity", "<u4"), ("x", "<f4"), ("y", "<f4"), ("z", "<f4")])
    rec["intensity"] = np.arange(n)
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS intensity x y z\nSIZE 4 4 4 4\nTYPE U F F F\n"
           "COUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (n, n))
    open(os.path.join(out_dir, "ixyz_binary.pcd"), "wb").write(hdr.encode() + rec.tobytes())
    np.save(os.path.join(out_dir, "ixyz_binary_expected.npy"), xyz)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "all":
    make_all_assets()
    copy_assets()
    sys.exit(0)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "jpeg":
        make_jpeg_fixtures()
        sys.exit(0)
    main()
    make_pcd_fixtures()
