"""CPU: the tracker oracle (oracle/track_oracle.py) against the reference's own component tests
(test/track/auction_test.cpp, singer_test.cpp, features_test.cpp), restated one to one."""
import numpy as np
import pytest

from oracle import track_oracle as to


# ---- test/track/auction_test.cpp:14-63 ----
def test_auction_equal_agents_and_tasks():
    assert to.auction(np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]], np.float32), 100) == [2, 1, 0]


def test_auction_more_agents_than_tasks():
    res = to.auction(np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9], [1, 4, 7]], np.float32), 100)
    assert len(res) == 4 and all(t in res for t in range(3))


def test_auction_more_tasks_than_agents():
    res = to.auction(np.arange(1, 13, dtype=np.float32).reshape(3, 4), 100)
    assert len(res) == 3 and all(r != to.NOT_MATCHED for r in res)


def test_auction_zero_iterations():
    assert to.auction(np.arange(1, 10, dtype=np.float32).reshape(3, 3), 0) == [to.NOT_MATCHED] * 3


# ---- test/track/singer_test.cpp:15-121 ----
def make_filter():
    return to.SingerEKF(np.zeros(9, np.float32), np.eye(9, dtype=np.float32) * 0.5, 2.0, 1.0, np.eye(3, dtype=np.float32) * 0.2)


def approx(a, b, prec):            # Eigen isApprox: ||a - b|| <= prec * min(||a||, ||b||)
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) <= prec * min(np.linalg.norm(a), np.linalg.norm(b))


def test_singer_stable():
    f = make_filter()
    z = np.array([10, 20, 30], np.float32)
    for _ in range(10):
        f.predict(1.0)
        f.update(z)
    assert approx(f.x[[0, 3, 6]], z, 1e-1)


def test_singer_uniform_motion():
    f = make_filter()
    p0, v = np.array([10, 20, 30], np.float32), np.array([2, 4, 6], np.float32)
    for i in range(10):
        f.predict(1.0)
        f.update(p0 + i * v)
    assert approx(f.x[[0, 3, 6]], p0 + 9 * v, 1e-1)
    assert approx(f.x[[1, 4, 7]], v, 1e-1)
    assert np.all(np.abs(f.x[[2, 5, 8]]) < 1e-1)


def test_singer_accelerated_motion():
    f = make_filter()
    p0, v, a = np.array([10, 20, 30.]), np.array([2, 4, 6.]), np.array([0, 0.5, 1.0])
    for i in range(10):
        f.predict(1.0)
        f.update((p0 + v * i + 0.5 * a * i * i).astype(np.float32))
    assert approx(f.x[[0, 3, 6]], p0 + v * 9 + 0.5 * a * 81, 1e-1)
    assert approx(f.x[[1, 4, 7]], v + a * 9, 1e-1)


# ---- test/track/features_test.cpp:14-111 ----
def test_features_constructors_and_growth():
    f = to.Features(feature_size=5, capacity=10)
    assert (f.size, f.capacity) == (0, 10)
    g = to.Features(feature=[1, 2, 3, 4, 5], capacity=5)
    assert (g.size, g.capacity) == (1, 5) and np.array_equal(g.get(0), [1, 2, 3, 4, 5])
    h = to.Features(feature_size=3)
    for want_size, want_cap in ((1, 1), (2, 2), (3, 4)):          # capacity doubles
        h.push_back([1, 2, 3])
        assert (h.size, h.capacity) == (want_size, want_cap)
    with pytest.raises(IndexError):
        g.get(1)
    g.clear()
    assert (g.size, g.capacity) == (0, 5) and not g.m.any()


def test_features_label_and_feature():
    f = to.Features(feature=[0.2, 0.8, 0.0])
    f.push_back([0.6, 0.4, 0.0])
    f.push_back([0.5, 0.5, 0.0])
    assert f.label() == 1                                          # row sums 1.3 / 1.7 / 0
    assert np.allclose(f.feature(), np.array([1.3, 1.7, 0]) / 3.0, atol=1e-6)
    assert not to.Features(feature_size=3).feature().any()         # iszero(sum): no division


# ---- Tracker life cycle (tracker.cpp:126-220; the reference has no test at this level) ----
def test_tracker_life_cycle():
    trk = to.Tracker([0.2, 0.2, 0.2], 12, init_thresh=3, miss_thresh=2)
    t = 0

    def frame(robots):
        nonlocal t
        t += 50_000_000
        trk.update(robots, t)
        return robots

    r = frame([to.RobotObs(armors=[(3, 0.9)], location=[1, 2, 0], label=3)])
    assert r[0].track_state == to.TENTATIVE and len(trk.tracks) == 1 and trk.tracks[0].track_id == 0
    for _ in range(3):
        r = frame([to.RobotObs(armors=[(3, 0.9)], location=[1.02, 2.0, 0], label=3)])
    assert r[0].track_state == to.CONFIRMED and trk.tracks[0].init_count == 3
    # an undetected, located robot next to the track inherits the track's label and filtered position
    r = frame([to.RobotObs(armors=None, location=[1.05, 2.0, 0])])
    assert r[0].label == 3 and r[0].track_state == to.CONFIRMED
    # two misses delete a confirmed track; an unlocated robot never matches
    frame([to.RobotObs(armors=[(3, 0.9)], location=None, label=3)])
    assert len(trk.tracks) == 1 and trk.tracks[0].miss_count == 1
    frame([])
    assert trk.tracks == []
    # a tentative track that misses once is dropped at once
    frame([to.RobotObs(armors=[(5, 0.7)], location=[4, 4, 0], label=5)])
    assert len(trk.tracks) == 1 and trk.tracks[0].track_id == 1
    frame([])
    assert trk.tracks == []
