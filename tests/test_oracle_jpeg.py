"""CPU: the JPEG oracle (oracle/jpeg_ref.c) against the reference's own decoder.

The reference reads camera frames with cv::imread (samples/main.cpp:24-40) = OpenCV's bundled libjpeg-turbo.
Pinned three ways: committed outputs of cv2.imdecode for small synthetic files (tests/golden/jpeg), sha256 of
the decoded reference frames, and -- cv2 is in this image -- a live sweep over encoder settings.  Bit-exact.
"""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import jpeg_oracle as jo
from tests import fixtures as fx

JPEG_DIR = os.path.join(fx.GOLDEN, "jpeg")


def test_oracle_matches_committed_cv2_outputs():
    exp = np.load(os.path.join(JPEG_DIR, "expected.npz"))
    files = sorted(glob.glob(os.path.join(JPEG_DIR, "*.jpg")))
    assert len(files) == 8
    for path in files:
        name = os.path.splitext(os.path.basename(path))[0]
        got = jo.decode(open(path, "rb").read())
        assert np.array_equal(got, exp[name]), name


def test_oracle_matches_reference_frames():
    exp = np.load(os.path.join(JPEG_DIR, "expected.npz"))
    for i in (0, 5):
        data = open(os.path.join(fx.GOLDEN, "frames", f"{i}.jpg"), "rb").read()
        meta = jo.info(data)
        assert (meta["width"], meta["height"], meta["h_samp"], meta["v_samp"]) == (2592, 2048, 2, 2)
        img = jo.decode(data)
        assert img.shape == tuple(exp[f"frame{i}_shape"])
        assert hashlib.sha256(img.tobytes()).digest() == exp[f"frame{i}_sha256"].tobytes()


def jpeg_sweep(sizes, qualities=(30, 90, 100), restarts=(0, 1, 7)):
    """(name, bytes) for photo and noise content over sampling x quality x restart interval x optimised tables."""
    import cv2
    src = fx.load_frame(0)
    rng = np.random.default_rng(0)
    S = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
         "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420}
    for (w, h) in sizes:
        photo = cv2.resize(src[500:1500, 800:2200], (w, h), interpolation=cv2.INTER_AREA)
        noise = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for cname, img in (("photo", photo), ("noise", noise)):
            for sname, code in S.items():
                for q in qualities:
                    for rst in restarts:
                        for opt in (0, 1):
                            ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, code,
                                                                 cv2.IMWRITE_JPEG_RST_INTERVAL, rst, cv2.IMWRITE_JPEG_OPTIMIZE, opt])
                            assert ok
                            yield f"{cname} {w}x{h} {sname} q{q} rst{rst} opt{opt}", enc.tobytes()
        ok, enc = cv2.imencode(".jpg", cv2.cvtColor(photo, cv2.COLOR_BGR2GRAY), [cv2.IMWRITE_JPEG_QUALITY, 80])
        yield f"gray {w}x{h}", enc.tobytes()


def test_oracle_matches_live_cv2_sweep():
    cv2 = pytest.importorskip("cv2")
    n = 0
    for name, data in jpeg_sweep([(640, 480), (333, 217), (17, 9), (8, 8), (5, 3), (3, 2), (2, 2), (1, 1), (1000, 31)]):
        want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        assert np.array_equal(jo.decode(data), want), name
        n += 1
    assert n > 900


def test_oracle_sampling_is_what_the_name_says():
    exp = {"photo_420_q90": (2, 2), "photo_444_q75_opt": (1, 1), "photo_422_q50_rst3": (2, 1), "gray_q80": (1, 1)}
    for name, hv in exp.items():
        meta = jo.info(open(os.path.join(JPEG_DIR, name + ".jpg"), "rb").read())
        assert (meta["h_samp"], meta["v_samp"]) == hv, name
    assert jo.info(open(os.path.join(JPEG_DIR, "photo_422_q50_rst3.jpg"), "rb").read())["restart_interval"] == 3


def test_oracle_rejects_what_it_does_not_decode():
    cv2 = pytest.importorskip("cv2")
    img = fx.load_frame(0)[:64, :64]
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(ValueError):
        jo.decode(enc.tobytes())
    good = open(os.path.join(JPEG_DIR, "photo_420_q90.jpg"), "rb").read()
    with pytest.raises(ValueError):
        jo.decode(good[:200])
    with pytest.raises(ValueError):
        jo.decode(b"not a jpeg at all")


def _insert_segment(jpeg: bytes, marker: int, payload: bytes) -> bytes:
    """A marker segment placed right after SOI."""
    seg = bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + payload
    return jpeg[:2] + seg + jpeg[2:]


def exif_orientation(value: int, little_endian: bool = True) -> bytes:
    bo = "little" if little_endian else "big"
    tiff = (b"II" if little_endian else b"MM") + (42).to_bytes(2, bo) + (8).to_bytes(4, bo)
    entry = (0x0112).to_bytes(2, bo) + (3).to_bytes(2, bo) + (1).to_bytes(4, bo) + value.to_bytes(2, bo) + b"\x00\x00"
    return b"Exif\x00\x00" + tiff + (1).to_bytes(2, bo) + entry + (0).to_bytes(4, bo)


def metadata_cases():
    """(name, file, decodes_like_the_plain_file) for the metadata that changes what cv::imread returns."""
    good = open(os.path.join(JPEG_DIR, "photo_420_q90.jpg"), "rb").read()
    yield "exif orientation 1", _insert_segment(good, 0xE1, exif_orientation(1)), True
    yield "exif orientation 1 big endian", _insert_segment(good, 0xE1, exif_orientation(1, False)), True
    yield "exif orientation 6", _insert_segment(good, 0xE1, exif_orientation(6)), False
    yield "exif orientation 3 big endian", _insert_segment(good, 0xE1, exif_orientation(3, False)), False
    yield "adobe transform 1", _insert_segment(good, 0xEE, b"Adobe\x00\x64\x00\x00\x00\x00\x01"), True
    yield "adobe transform 0 beside JFIF (JFIF wins)", _insert_segment(good, 0xEE, b"Adobe\x00\x64\x00\x00\x00\x00\x00"), True
    assert good[2:4] == b"\xff\xe0" and good[6:10] == b"JFIF"
    no_jfif = good[:2] + good[4 + int.from_bytes(good[4:6], "big"):]
    yield "no JFIF, no Adobe (component ids 1 2 3)", no_jfif, True
    yield "adobe transform 0 without JFIF (RGB)", _insert_segment(no_jfif, 0xEE, b"Adobe\x00\x64\x00\x00\x00\x00\x00"), False
    yield "comment", _insert_segment(good, 0xFE, b"hello"), True


def test_metadata_that_changes_the_reference_output_is_rejected():
    """cv::imread rotates by the EXIF orientation and skips the colour transform for RGB-coded files: both are outside
    the decoder's scope and must fail loudly instead of returning other pixels than the reference would."""
    cv2 = pytest.importorskip("cv2")
    good = open(os.path.join(JPEG_DIR, "photo_420_q90.jpg"), "rb").read()
    plain = jo.decode(good)
    for name, data, same in metadata_cases():
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        if same:
            assert np.array_equal(ref, plain), name                     # the reference is unaffected by this segment
            assert np.array_equal(jo.decode(data), plain), name
        else:
            assert ref is None or ref.shape != plain.shape or not np.array_equal(ref, plain), name   # the reference differs
            with pytest.raises(ValueError):
                jo.decode(data)
