"""GPU: PCD ingestion — the device parser against the oracle's reader
(oracle/locate_oracle.py: read_pcd, the pcl::io::loadPCDFile stand-in) on committed file images:
a prefix of the reference's assets/clouds/0.pcd (ASCII), an ASCII file with decimals / exponents / signs /
CRLF / an extra column / nan / inf, and a binary file whose records carry a leading intensity field."""
import os

import numpy as np
import pytest

import rm_radar_b200 as rr
from oracle import locate_oracle as lo
from tests import fixtures as fx

pytestmark = pytest.mark.gpu
PCD = os.path.join(fx.GOLDEN, "pcd")


def ascii_expected(path, columns=(0, 1, 2)):
    """Token-wise restatement for files the oracle's 3-column reader does not cover: text -> float -> float32."""
    data = open(path, "rb").read()
    body = data[data.index(b"DATA ascii") :].split(b"\n", 1)[1]
    rows = [ln.split() for ln in body.replace(b"\r", b"").split(b"\n") if ln.strip()]
    return np.array([[np.float32(float(r[c])) for c in columns] for r in rows], np.float32)


def same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32)[~np.isnan(b)], b.view(np.uint32)[~np.isnan(b)]) and \
        np.array_equal(np.isnan(a), np.isnan(b))


def test_asset_prefix_ascii_matches_oracle_reader():
    path = os.path.join(PCD, "asset0_head1500_ascii.pcd")
    got = rr.pcd_parse(open(path, "rb").read())
    want = lo.read_pcd(path)
    assert want.shape == (1500, 3) and same(got, want)


def test_ascii_variants_and_binary_records():
    path = os.path.join(PCD, "variants_crlf_ascii.pcd")
    got = rr.pcd_parse(open(path, "rb").read())
    want = ascii_expected(path)
    assert want.shape == (400, 3) and np.isnan(want[5, 0]) and np.isinf(want[6, 0])
    assert same(got, want), np.argwhere(got.view(np.uint32) != want.view(np.uint32))[:5]
    got = rr.pcd_parse(open(os.path.join(PCD, "ixyz_binary.pcd"), "rb").read())
    assert same(got, np.load(os.path.join(PCD, "ixyz_binary_expected.npy")))
    with pytest.raises(ValueError):
        rr.pcd_parse(b"VERSION 0.7\nFIELDS x y\nPOINTS 1\nDATA ascii\n1 2\n")          # no z field
    with pytest.raises(ValueError):
        rr.pcd_parse(b"VERSION 0.7\nFIELDS x y z\nPOINTS 3\nDATA ascii\n1 2 3\n")       # fewer lines than POINTS
    with pytest.raises(ValueError):
        rr.pcd_parse(b"VERSION 0.7\nFIELDS x y z\nPOINTS 1\nDATA binary_compressed\n")


def test_update_from_pcd_equals_update_from_array():
    """Locator.update_pcd(file) leaves exactly the state Locator.update(read_pcd(file)) leaves."""
    path = os.path.join(PCD, "asset0_head1500_ascii.pcd")
    blob = open(path, "rb").read()
    cloud = lo.read_pcd(path)
    a = rr.Locator(*fx.IMAGE_SIZE, fx.INTRINSIC, fx.LIDAR_TO_CAMERA, fx.WORLD_TO_CAMERA)
    b = rr.Locator(*fx.IMAGE_SIZE, fx.INTRINSIC, fx.LIDAR_TO_CAMERA, fx.WORLD_TO_CAMERA)
    for _ in range(2):
        assert a.update_pcd(blob) == 1500
        b.update(cloud)
    for which in ("depth", "background", "diff"):
        assert np.array_equal(a.image(which).view(np.uint32), b.image(which).view(np.uint32))
