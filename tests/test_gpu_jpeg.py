"""GPU parity: device JPEG decode (rm_radar_b200/csrc/jpeg.cu) through the C ABI vs the oracle and vs cv2.imdecode.

Replaces the reference's cv::imread (samples/main.cpp:24-40).  Integer pipeline -> every comparison is bit-exact.
"""
import glob
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

import rm_radar_b200 as rr
from oracle import jpeg_oracle as jo
from tests import fixtures as fx
from tests.test_oracle_jpeg import JPEG_DIR, jpeg_sweep

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def decoder():
    return rr.JpegDecoder(0)


def check(decoder, data, want, name=""):
    got = decoder.decode(data)
    st = decoder.status()
    assert st["status"] == 0, (name, st)
    assert got.shape == want.shape, name
    if not np.array_equal(got, want):
        bad = np.argwhere((got != want).any(axis=2))
        raise AssertionError(f"{name}: {len(bad)} pixels differ, first at {bad[0]}, sync rounds {st['sync_rounds']}")
    return st


def test_committed_fixtures_bit_exact(decoder):
    exp = np.load(os.path.join(JPEG_DIR, "expected.npz"))
    for path in sorted(glob.glob(os.path.join(JPEG_DIR, "*.jpg"))):
        name = os.path.splitext(os.path.basename(path))[0]
        data = open(path, "rb").read()
        check(decoder, data, exp[name], name)
        # the entropy stage on its own: quantised coefficients, DC predictors resolved
        _, coef = jo.decode(data, want_coefficients=True)
        assert np.array_equal(decoder.coefficients(len(coef)), coef), name


def test_reference_frames_bit_exact(decoder):
    exp = np.load(os.path.join(JPEG_DIR, "expected.npz"))
    for i in (0, 5):
        data = open(os.path.join(fx.GOLDEN, "frames", f"{i}.jpg"), "rb").read()
        assert rr.jpeg_info(data) == jo.info(data)
        want, coef = jo.decode(data, want_coefficients=True)
        st = check(decoder, data, want, f"frame {i}")
        assert np.array_equal(decoder.coefficients(len(coef)), coef)
        got = decoder.decode(data)
        assert hashlib.sha256(got.tobytes()).digest() == exp[f"frame{i}_sha256"].tobytes()
        # 1 MB of file instead of 16 MB of pixels over PCIe; a handful of synchronisation rounds
        assert st["upload_bytes"] < len(data) + 65536 and st["sync_rounds"] <= 40, st


def test_sweep_matches_oracle_and_cv2(decoder):
    import cv2
    n = 0
    for name, data in jpeg_sweep([(640, 480), (333, 217), (17, 9), (8, 8), (3, 2), (1, 1), (1000, 31)], qualities=(30, 100), restarts=(0, 1, 7)):
        want = jo.decode(data)
        check(decoder, data, want, name)
        if n % 16 == 0:
            assert np.array_equal(want, cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)), name
        n += 1
    assert n > 400


def test_large_frames_with_restart_markers_and_custom_tables(decoder):
    import cv2
    src = fx.load_frame(5)
    for (w, h), flags in (((1920, 1080), [cv2.IMWRITE_JPEG_RST_INTERVAL, 120, cv2.IMWRITE_JPEG_QUALITY, 92]),
                          ((2592, 2048), [cv2.IMWRITE_JPEG_OPTIMIZE, 1, cv2.IMWRITE_JPEG_QUALITY, 98]),
                          ((2591, 2047), [cv2.IMWRITE_JPEG_RST_INTERVAL, 1, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422]),
                          ((1280, 720), [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_QUALITY, 100])):
        img = cv2.resize(src, (w, h), interpolation=cv2.INTER_AREA)
        ok, enc = cv2.imencode(".jpg", img, flags)
        data = enc.tobytes()
        want = cv2.imdecode(enc, cv2.IMREAD_COLOR)
        assert np.array_equal(jo.decode(data), want)
        check(decoder, data, want, f"{w}x{h} {flags}")
    gray = cv2.cvtColor(src, cv2.COLOR_BGR2GRAY)
    ok, enc = cv2.imencode(".jpg", gray, [cv2.IMWRITE_JPEG_RST_INTERVAL, 3])
    check(decoder, enc.tobytes(), cv2.imdecode(enc, cv2.IMREAD_COLOR), "gray rst3")


def test_decode_device_into_caller_buffer_with_pitch(decoder):
    import torch
    data = open(os.path.join(JPEG_DIR, "photo_420_q90.jpg"), "rb").read()
    want = jo.decode(data)
    h, w = want.shape[:2]
    pitch = w * 3 + 13
    buf = torch.full((h, pitch), 7, dtype=torch.uint8, device="cuda:0")
    ptr, ww, hh = decoder.decode_device(data, buf.data_ptr(), pitch)
    assert (ptr, ww, hh) == (buf.data_ptr(), w, h)
    assert decoder.status()["status"] == 0
    got = buf.cpu().numpy()
    assert np.array_equal(got[:, :w * 3].reshape(h, w, 3), want)
    assert (got[:, w * 3:] == 7).all()          # the padding of each row is left alone


def test_rejects_unsupported_and_flags_corrupt_streams(decoder):
    import cv2
    img = fx.load_frame(0)[:64, :64]
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(ValueError, match="progressive"):      # std::invalid_argument, like the oracle's ValueError
        decoder.decode(enc.tobytes())
    with pytest.raises(ValueError):
        decoder.decode(b"\x89PNG not a jpeg")
    good = open(os.path.join(JPEG_DIR, "photo_420_q90.jpg"), "rb").read()
    with pytest.raises(ValueError):
        decoder.decode(good[:300])
    # entropy-coded data cut short (EOI kept): the block count cannot come out right
    cut = good[:len(good) // 2] + b"\xff\xd9"
    with pytest.raises(rr.RadarError, match="corrupt"):
        decoder.decode(cut)
    # over-subscribed Huffman table (a DHT whose code-length counts break the Kraft bound, e.g. 255 codes of
    # length 1): must be rejected before any table entry is written (it used to index far past the lookup table)
    i = good.index(b"\xff\xc4")
    bad_dht = bytearray(good)
    bits = bad_dht[i + 5:i + 21]                  # marker(2) length(2) class/id(1) then bits[1..16]
    donor = max(range(16), key=lambda l: bits[l])
    assert bits[donor] >= 3
    bad_dht[i + 5 + donor] -= 3                   # same symbol count, so the segment length still parses
    bad_dht[i + 5] += 3                           # three codes of length 1: only two exist
    with pytest.raises(ValueError, match="Huffman"):
        decoder.decode(bytes(bad_dht))
    # and the decoder is still usable afterwards
    check(decoder, good, jo.decode(good), "after errors")


def test_metadata_segments(decoder):
    """EXIF orientation 1 / JFIF-beside-Adobe / comments decode like the plain file; rotated or RGB-coded files raise."""
    from tests.test_oracle_jpeg import metadata_cases
    plain = jo.decode(open(os.path.join(JPEG_DIR, "photo_420_q90.jpg"), "rb").read())
    for name, data, ok in metadata_cases():
        if ok:
            check(decoder, data, plain, name)
        else:
            with pytest.raises(ValueError):
                decoder.decode(data)


def test_result_does_not_depend_on_the_subsequence_size():
    """The fixed point of the synchronisation is the sequential decode whatever the cut (jpeg.cu header)."""
    code = (
        "import sys, hashlib; sys.path.insert(0, %r)\n"
        "import rm_radar_b200 as rr\n"
        "d = rr.JpegDecoder(0); data = open(%r, 'rb').read()\n"
        "img = d.decode(data); st = d.status()\n"
        "print(hashlib.sha256(img.tobytes()).hexdigest(), st['status'], st['sync_rounds'])\n"
    ) % (fx.ROOT, os.path.join(fx.GOLDEN, "frames", "0.jpg"))
    exp = np.load(os.path.join(JPEG_DIR, "expected.npz"))["frame0_sha256"].tobytes().hex()
    for bits, simple in (("256", "0"), ("1024", "0"), ("8192", "0"), ("65536", "0"), ("512", "1"), ("1024", "1"), ("4096", "1")):
        env = dict(os.environ, RMR_JPEG_SUB_BITS=bits, RMR_JPEG_SIMPLE=simple)   # hypothesis kernel and single-guess kernel
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        digest, status, rounds = out.stdout.split()[-3:]
        assert (digest, status) == (exp, "0"), (bits, simple, rounds)


@pytest.mark.skipif(not fx.have_models(), reason="engines not built")
def test_detect_jpeg_equals_detect_on_the_decoded_frame(decoder):
    data = open(os.path.join(fx.GOLDEN, "frames", "0.jpg"), "rb").read()
    img = fx.load_frame(0)
    det = rr.RobotDetector(fx.engine("car"), fx.engine("armor"), (img.shape[1], img.shape[0]), fx.CLASS_NUM, fx.MAX_BATCH, fx.OPT_BATCH)
    a = det.detect(img)
    b = det.detect_jpeg(decoder, data)
    assert len(a) == len(b) > 0
    for x, y in zip(a, b):
        assert x.label == y.label and x.rect == y.rect and x.confidence == y.confidence
