"""Pins oracle/locate_oracle.py against the properties the reference's own locator tests check
(/root/reference/test/locate/locator_test.cpp) and against a brute-force BFS statement of PCL's
Euclidean clustering.  CPU only."""
import numpy as np
import pytest

from oracle import locate_oracle as lo

I3 = np.eye(3, dtype=np.float32)
I4 = np.eye(4, dtype=np.float32)


def fixture_locator():
    # locator_test.cpp:17-29
    return lo.LocatorOracle(640, 480, I3, I4, I4, zoom_factor=0.5, queue_size=5, min_depth_diff=0.05,
                            max_depth_diff=5.0, cluster_tolerance=100, min_cluster_size=10,
                            max_cluster_size=1000, max_distance=20)


def test_zoom():
    # locator_test.cpp:43-51
    loc = fixture_locator()
    r = loc.zoom_rect((100, 100, 50, 60))
    assert r[2] == int(50 * 0.5) and r[3] == int(60 * 0.5)


def test_coordinate_round_trip_identity():
    # locator_test.cpp:53-74 (identity calibration: lidar->camera->lidar is exact up to float)
    loc = fixture_locator()
    p = np.array([[100.0, 200.0, 50.0]], np.float32)
    u, v, d = loc.lidar_to_camera(p)
    assert np.allclose([u[0], v[0], d[0]], [100 * 0.5 / 50, 200 * 0.5 / 50, 50], rtol=1e-6)
    back = loc.camera_to_lidar(u, v, d)
    assert np.allclose(back[0], p[0], rtol=1e-5)
    w = loc.lidar_to_world(p[0])
    assert np.allclose(w, p[0], rtol=1e-6)


def two_blobs(loc, seed):
    # locator_test.cpp:76-119: two Gaussian pixel blobs (sigma 10 px, 500 draws each) written
    # straight into diff_depth_image_, depths U(5,6) and U(1,2)
    rng = np.random.default_rng(seed)
    for (cx, cy, lo_d, hi_d) in [(160, 120, 5.0, 6.0), (80, 60, 1.0, 2.0)]:
        xs = np.clip(rng.normal(cx, 10, 500).astype(int), 0, loc.Wz - 1)
        ys = np.clip(rng.normal(cy, 10, 500).astype(int), 0, loc.Hz - 1)
        loc.diff[ys, xs] = rng.uniform(lo_d, hi_d, 500).astype(np.float32)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_cloud_cluster_two_blobs(seed):
    loc = fixture_locator()
    two_blobs(loc, seed)
    loc.cluster()
    assert loc.num_clusters == 2          # locator_test.cpp:118


@pytest.mark.parametrize("seed", [0, 1])
def test_robot_search_has_location(seed):
    # locator_test.cpp:121-167: rect_ = (140,100,40,40) at full res -> zoomed covers blob 2 (80,60)
    loc = fixture_locator()
    two_blobs(loc, seed)
    loc.cluster()
    xyz, info = loc.search_rect((140, 100, 40, 40))
    assert xyz is not None and info["n"] > 0


def test_components_match_bruteforce_bfs():
    rng = np.random.default_rng(3)
    pts = rng.uniform(0, 3000, (600, 3)).astype(np.float32)
    a = lo.radius_components(pts, 400)
    b = lo.radius_components_bruteforce(pts, 400)
    m = {}
    assert all(m.setdefault(x, y) == y for x, y in zip(a, b)) and len(set(a)) == len(set(b))


def test_update_last_writer_and_background_max():
    loc = lo.LocatorOracle(64, 64, I3, I4, I4, zoom_factor=1.0, queue_size=3, min_depth_diff=1, max_depth_diff=10)
    # two points on the same pixel (u=10.2, v=5.5) with depths 8 then 6; (0,0,0) skipped; x>max skipped
    pts = np.array([[10.2 * 8, 5.5 * 8, 8], [10.6 * 6, 5.1 * 6, 6], [0, 0, 0], [40000, 1, 1]], np.float32)
    loc.update(pts)
    assert loc.depth[5, 10] == 6 and loc.background[5, 10] == 8
    assert loc.diff[5, 10] == 6          # bg - depth = 2 in [1, 10]
    assert loc.stats["valid"] == 2 and loc.stats["collisions"] == 1
    loc.update(None)                      # locate.cpp:163-166: images cleared, nothing else
    assert not loc.diff.any() and len(loc.ring) == 1


def test_cluster_order_and_size_filter():
    loc = lo.LocatorOracle(200, 200, I3, I4, I4, zoom_factor=1.0, cluster_tolerance=2.0,
                           min_cluster_size=3, max_cluster_size=6)
    # depth 1 everywhere: pixel distance == lidar distance; three runs of 3, 5, 8 px and a singleton
    loc.diff[10, 10:13] = 1
    loc.diff[50, 10:15] = 1
    loc.diff[90, 10:18] = 1      # 8 > max -> dropped whole
    loc.diff[130, 10] = 1        # 1 < min
    loc.diff[150, 10:13] = 1     # same size as the first: tie -> ascending first index
    loc.cluster()
    assert loc.num_clusters == 3 and loc.cluster_sizes == [5, 3, 3]
    assert loc.label_image[50, 10] == 0 and loc.label_image[10, 10] == 1 and loc.label_image[150, 10] == 2
    assert loc.label_image[90, 10] == -1 and loc.label_image[130, 10] == -1 and loc.label_image[0, 0] == -2
    # unclustered (-1) group competes and wins on ties by lowest id (locate.cpp:303-306)
    xyz, info = loc.search_rect((8, 88, 12, 70))   # covers the 8-run (-1), the singleton (-1), cluster 2
    assert info["cluster"] == -1 and info["n"] == 9


def test_committed_pcd_prefix_matches_reference_asset():
    """tests/golden/pcd/asset0_head1500_ascii.pcd is the first 1500 points of the reference's assets/clouds/0.pcd."""
    import os
    from tests.conftest import GOLDEN, REFERENCE, has_reference
    a = lo.read_pcd(os.path.join(GOLDEN, "pcd", "asset0_head1500_ascii.pcd"))
    assert a.shape == (1500, 3) and a.dtype == np.float32
    if has_reference():
        assert np.array_equal(a, lo.read_pcd(os.path.join(REFERENCE, "assets", "clouds", "0.pcd"))[:1500])
