"""oracle/track_oracle.py -- TEST INFRASTRUCTURE ONLY (checker; never imported by the product).

numpy float32 restatement of the reference's tracker, the step after the hot path:
  Features            src/track/features.h:27-208   (growing column store, label = arg-max of the row sums,
                                                     feature = row sums / total)
  SingerEKF           src/track/singer.h:27-131 over ExtendedKalmanFilter, src/track/kalman_filter.h:170-296
                      (9-state position / velocity / acceleration per axis, Singer transition + process noise)
  auction             src/track/auction.h:33-126    (forward auction without epsilon, virtual zero-value tasks)
  Robot.feature       src/robot/robot.cpp:102-122,  Robot.setTrack  src/robot/robot.cpp:81-94
  Tracker             src/track/tracker.cpp:47-220  (cost = distance score + cosine feature score, match gating,
                                                     tentative / confirmed / deleted life cycle)
Pinned by the reference's own component tests (test/track/auction_test.cpp, singer_test.cpp, features_test.cpp)
restated in tests/test_oracle_track.py; the reference has no tracker-level test or vector, so Tracker.update as a
whole is "parity unpinned" beyond those components.
`iszero` in the reference is glibc's classification macro: an exact comparison with zero.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
NOT_MATCHED = -1
TENTATIVE, CONFIRMED, DELETED = 0, 1, 2


class Features:
    """features.h:27-208."""

    def __init__(self, feature=None, feature_size=None, capacity=1):
        if feature is not None:
            feature = np.asarray(feature, f32)
            self.m = np.zeros((feature.size, capacity), f32)
            self.m[:, 0] = feature
            self.size = 1
        else:
            self.m = np.zeros((feature_size, capacity), f32)
            self.size = 0
        self.capacity = capacity

    def push_back(self, feature):
        feature = np.asarray(feature, f32)
        if feature.size != self.m.shape[0]:
            raise RuntimeError("row of feature is not the same")
        if self.size >= self.capacity:                      # features.h:103-110: double, zero-filled
            self.capacity *= 2
            grown = np.zeros((self.m.shape[0], self.capacity), f32)
            grown[:, :self.m.shape[1]] = self.m
            self.m = grown
        self.m[:, self.size] = feature
        self.size += 1

    def get(self, index):
        if index < 0 or index >= self.size:
            raise IndexError("index out of range")
        return self.m[:, index].copy()

    def clear(self):
        self.size = 0
        self.m[:] = 0

    def label(self):
        return int(np.argmax(self.m.sum(axis=1, dtype=f32)))      # maxCoeff: first maximum

    def feature(self):
        total = self.m.sum(dtype=f32)
        if total == 0:
            return np.zeros(self.m.shape[0], f32)
        return (self.m.sum(axis=1, dtype=f32) / total).astype(f32)


class SingerEKF:
    """singer.h:27-131 + kalman_filter.h:206-293."""

    def __init__(self, state, covariance, max_a, tau, observation_noise):
        self.x = np.asarray(state, f32).reshape(9).copy()
        self.P = np.asarray(covariance, f32).reshape(9, 9).copy()
        self.R = np.asarray(observation_noise, f32).reshape(3, 3).copy()
        self.max_a, self.tau = f32(max_a), f32(tau)

    def transition(self, dt):
        dt = f32(dt)
        F = np.eye(9, dtype=f32)
        for i in range(3):
            F[3 * i, 3 * i + 1] = dt
            F[3 * i, 3 * i + 2] = dt * dt / f32(2)
            F[3 * i + 1, 3 * i + 2] = dt
            F[3 * i + 2, 3 * i + 2] = np.exp(-dt / self.tau, dtype=f32)
        return F

    def process_noise(self, dt):
        dt = f32(dt)
        Q = np.zeros((9, 9), f32)
        e1 = f32(1) - np.exp(-dt / self.tau, dtype=f32)
        e2 = (f32(1) - np.exp(f32(-2) * dt / self.tau, dtype=f32)) / f32(2)
        for i in range(3):
            b = 3 * i
            Q[b, b] = f32(float(dt) ** 3 / 3)                    # std::pow(float, int) is evaluated in double
            Q[b + 1, b] = Q[b, b + 1] = f32(float(dt) ** 2 / 2)
            Q[b + 2, b] = Q[b, b + 2] = dt / f32(2)
            Q[b + 1, b + 1] = dt
            Q[b + 2, b + 1] = Q[b + 1, b + 2] = e1
            Q[b + 2, b + 2] = e2
        return (Q * f32(float(self.max_a) ** 2)).astype(f32)

    def predict(self, dt):
        F, Q = self.transition(dt), self.process_noise(dt)
        self.x = (F @ self.x).astype(f32)
        self.P = (F @ self.P @ F.T + Q).astype(f32)

    def update(self, z):
        z = np.asarray(z, f32).reshape(3)
        H = np.zeros((3, 9), f32)
        H[0, 0] = H[1, 3] = H[2, 6] = 1
        residual = z - self.x[[0, 3, 6]]
        S = (H @ self.P @ H.T + self.R).astype(f32)
        K = (self.P @ H.T @ np.linalg.inv(S).astype(f32)).astype(f32)
        self.x = (self.x + K @ residual).astype(f32)
        self.P = ((np.eye(9, dtype=f32) - K @ H) @ self.P).astype(f32)


def auction(value_matrix, max_iter):
    """auction.h:33-126.  value_matrix [agents, tasks] -> task index per agent, -1 = unmatched."""
    V = np.asarray(value_matrix, f32)
    n_agents = V.shape[0]
    n_tasks = V.shape[1] if V.ndim == 2 else 0
    n_real = n_tasks
    if n_agents > n_tasks:                                   # virtual zero-value tasks
        E = np.zeros((n_agents, n_agents), f32)
        E[:, :n_tasks] = V.reshape(n_agents, n_tasks)
        V, n_tasks = E, n_agents
    prices = np.zeros(n_tasks, f32)
    assignment = [NOT_MATCHED] * n_agents
    it = 0
    while it < max_iter:
        if sum(1 for a in assignment if 0 <= a <= n_real) >= n_agents:      # auction.h:58-61 (<=, as written)
            break
        changed = False
        for agent in range(n_agents):
            if assignment[agent] != NOT_MATCHED:
                continue
            best_task, best_value = NOT_MATCHED, -np.inf
            for task in range(n_tasks):
                value = f32(V[agent, task] - prices[task])
                if value > best_value:
                    best_value, best_task = value, task
            if best_task != NOT_MATCHED:
                prices[best_task] = f32(prices[best_task] + best_value)
                for other in range(n_agents):
                    if assignment[other] == best_task:
                        assignment[other] = NOT_MATCHED
                        break
                assignment[agent] = best_task
                changed = True
        if not changed:
            break
        it += 1
    return [NOT_MATCHED if a >= n_real else a for a in assignment]


class RobotObs:
    """The fields of radar::Robot the tracker reads and writes (robot.h:53-164)."""

    def __init__(self, armors=None, location=None, label=None):
        self.armors = armors              # list of (label, confidence) or None  (isDetected = armors is not None)
        self.location = None if location is None else np.asarray(location, f32)
        self.label = label
        self.track_state = None

    def is_detected(self):
        return self.armors is not None

    def is_located(self):
        return self.location is not None

    def feature(self, class_num):         # robot.cpp:102-122
        v = np.zeros(class_num, f32)
        if not self.is_detected():
            return v
        for label, conf in self.armors:
            v[int(label)] = f32(v[int(label)] + f32(conf))
        s = v.sum(dtype=f32)
        return v if s == 0 else (v / s).astype(f32)

    def set_track(self, track):           # robot.cpp:81-94
        self.track_state = track.state
        if track.state == CONFIRMED:
            self.label = track.label()
            self.location = track.location()
        else:
            if self.label is None:
                self.label = track.label()
            if self.location is None:
                self.location = track.location()


class Track:
    """track.h:27-196."""

    def __init__(self, location, feature, timestamp_ns, track_id, max_acc, tau, observe_noise):
        self.features = Features(feature=feature)
        self.timestamp = int(timestamp_ns)
        self.track_id = track_id
        self.init_count = 0
        self.miss_count = 0
        self.state = TENTATIVE
        x0 = np.zeros(9, f32)
        x0[0], x0[3], x0[6] = location
        self.filter = SingerEKF(x0, np.eye(9, dtype=f32) * f32(0.1), max_acc, tau, np.diag(np.asarray(observe_noise, f32)))

    def predict(self, timestamp_ns):
        dt = f32(float(f32(int(timestamp_ns) - self.timestamp)) * 1e-9)   # float(ns) * 1e-9 in double -> float, track.h:111-116
        self.filter.predict(dt)
        self.timestamp = int(timestamp_ns)

    def update(self, location, feature):
        self.features.push_back(feature)
        self.filter.update(location)

    def label(self):
        return self.features.label()

    def feature(self):
        return self.features.feature()

    def location(self):
        return self.filter.x[[0, 3, 6]].copy()


class Tracker:
    """tracker.cpp:47-220."""

    def __init__(self, observation_noise, class_num, init_thresh=4, miss_thresh=10, max_acceleration=2.0,
                 acceleration_correlation_time=1.0, distance_weight=0.40, feature_weight=0.60, max_iter=100,
                 distance_thresh=0.8):
        self.noise = np.asarray(observation_noise, f32)
        self.class_num = class_num
        self.init_thresh, self.miss_thresh = init_thresh, miss_thresh
        self.max_acc, self.tau = f32(max_acceleration), f32(acceleration_correlation_time)
        self.wd, self.wf = f32(distance_weight), f32(feature_weight)
        self.max_iter = max_iter
        self.dthr = f32(distance_thresh)
        self.tracks = []
        self.latest_id = 0

    @staticmethod
    def distance(a, b):
        d = np.asarray(a, f32) - np.asarray(b, f32)
        return f32(np.sqrt(f32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])))

    def cost(self, track, robot):
        if not robot.is_located() and not robot.is_detected():
            return f32(0)
        if not robot.is_located():
            ds = f32(0)
        else:
            d = self.distance(robot.location, track.location())
            ds = f32(1) if d < self.dthr else (f32(-d / self.dthr + f32(2)) if d < f32(2) * self.dthr else f32(0))
        fr, ft = robot.feature(self.class_num), track.feature()
        denom = f32(np.sqrt(f32(np.dot(fr, fr))) * np.sqrt(f32(np.dot(ft, ft))))
        fs = f32(0) if denom == 0 else f32((f32(np.dot(fr, ft)) / denom + f32(1)) / f32(2))
        return f32(ds * self.wd + fs * self.wf)

    def update(self, robots, timestamp_ns):
        for t in self.tracks:
            t.predict(timestamp_ns)
        C = np.zeros((len(robots), len(self.tracks)), f32)
        for r, robot in enumerate(robots):
            for t, track in enumerate(self.tracks):
                C[r, t] = self.cost(track, robot)
        unmatched, matched = [], []
        for r, t in enumerate(auction(C, self.max_iter)):
            robot = robots[r]
            if not robot.is_located() or t == NOT_MATCHED:
                unmatched.append(r)
                continue
            track = self.tracks[t]
            far = self.distance(robot.location, track.location()) > f32(2) * self.dthr
            if far and (robot.label if robot.label is not None else -1) != track.label():
                unmatched.append(r)
                continue
            track.update(robot.location, robot.feature(self.class_num))
            if track.state == TENTATIVE:
                track.init_count += 1
                if track.init_count >= self.init_thresh:
                    track.state = CONFIRMED
            track.miss_count = 0
            robot.set_track(track)
            matched.append(t)
        for i, track in enumerate(self.tracks):
            if i not in matched:
                if track.state == TENTATIVE:
                    track.state = DELETED
                elif track.state == CONFIRMED:
                    track.miss_count += 1
                    if track.miss_count >= self.miss_thresh:
                        track.state = DELETED
        self.tracks = [t for t in self.tracks if t.state != DELETED]
        for r in unmatched:
            robot = robots[r]
            if robot.is_detected() and robot.is_located():
                track = Track(robot.location, robot.feature(self.class_num), timestamp_ns, self.latest_id,
                              self.max_acc, self.tau, self.noise)
                self.latest_id += 1
                robot.set_track(track)
                self.tracks.append(track)
