// Tracker: the step after the hot path (reference src/track/tracker.cpp:47-220, track.h,
// singer.h, kalman_filter.h, auction.h, features.h and Robot::feature / Robot::setTrack in src/robot/robot.cpp:81-122).
//
// Host code on purpose: the work is a <= 20 x 20 cost matrix and a 9-state filter per track -- microseconds of scalar
// arithmetic that consumes the robot records `rmr_run_once` has just brought back, so a kernel would only add a launch and a
// round trip.  No Eigen: the Singer model is block diagonal (three independent position / velocity / acceleration
// triples), the observation picks x[0], x[3], x[6], so the filter is written out on fixed 9 x 9 float arrays.  The
// reference keeps every feature vector of a track in a growing matrix but only ever reads its row sums
// (features.h:173-197): the row sums are what is kept here.  float32 like the reference; summation orders differ from
// Eigen's, parity against oracle/track_oracle.py is to 1e-4 (tests/test_track.py).
#include "track.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <stdexcept>

namespace rmr {

namespace {
constexpr int kNotMatched = -1;

inline float distance3(const float a[3], const float b[3]) {
    const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return std::sqrt(dx * dx + dy * dy + dz * dz);
}

// 3 x 3 inverse by cofactors (what Eigen's fixed-size inverse() does)
bool invert3(const float m[9], float out[9]) {
    const float c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    if (det == 0.f) return false;
    const float inv = 1.f / det;
    out[0] = c00 * inv;
    out[1] = (m[2] * m[7] - m[1] * m[8]) * inv;
    out[2] = (m[1] * m[5] - m[2] * m[4]) * inv;
    out[3] = c01 * inv;
    out[4] = (m[0] * m[8] - m[2] * m[6]) * inv;
    out[5] = (m[2] * m[3] - m[0] * m[5]) * inv;
    out[6] = c02 * inv;
    out[7] = (m[1] * m[6] - m[0] * m[7]) * inv;
    out[8] = (m[0] * m[4] - m[1] * m[3]) * inv;
    return true;
}
}  // namespace

std::vector<int> auction(const std::vector<float>& values, int n_agents, int n_tasks, int max_iter) {
    const int n_real = n_tasks;
    const int cols = std::max(n_agents, n_tasks);          // more agents than tasks: virtual tasks of value 0
    std::vector<float> v(static_cast<size_t>(n_agents) * cols, 0.f);
    for (int a = 0; a < n_agents; ++a)
        for (int t = 0; t < n_tasks; ++t) v[static_cast<size_t>(a) * cols + t] = values[static_cast<size_t>(a) * n_tasks + t];
    n_tasks = cols;
    std::vector<float> prices(static_cast<size_t>(n_tasks), 0.f);
    std::vector<int> assignment(static_cast<size_t>(n_agents), kNotMatched);
    for (int it = 0; it < max_iter; ++it) {
        int settled = 0;
        for (int a : assignment) settled += (a >= 0 && a <= n_real) ? 1 : 0;      // `<=` as the reference writes it (auction.h:58-61)
        if (settled >= n_agents) break;
        bool changed = false;
        for (int agent = 0; agent < n_agents; ++agent) {
            if (assignment[agent] != kNotMatched) continue;
            int best = kNotMatched;
            float best_value = -std::numeric_limits<float>::infinity();
            for (int t = 0; t < n_tasks; ++t) {
                const float value = v[static_cast<size_t>(agent) * cols + t] - prices[t];
                if (value > best_value) { best_value = value; best = t; }
            }
            if (best == kNotMatched) continue;
            prices[best] += best_value;                       // no epsilon: the bidder pays its whole margin
            for (int other = 0; other < n_agents; ++other)
                if (assignment[other] == best) { assignment[other] = kNotMatched; break; }
            assignment[agent] = best;
            changed = true;
        }
        if (!changed) break;
    }
    for (int& a : assignment)
        if (a >= n_real) a = kNotMatched;
    return assignment;
}

Tracker::Tracker(const float observation_noise[3], int class_num, int init_thresh, int miss_thresh, float max_acceleration,
                 float acceleration_correlation_time, float distance_weight, float feature_weight, int max_iter,
                 float distance_thresh)
    : class_num_(class_num), init_thresh_(init_thresh), miss_thresh_(miss_thresh), max_acc_(max_acceleration),
      tau_(acceleration_correlation_time), wd_(distance_weight), wf_(feature_weight), max_iter_(max_iter), dthr_(distance_thresh) {
    if (observation_noise == nullptr) throw std::invalid_argument("Tracker: null observation noise");
    if (class_num <= 0) throw std::invalid_argument("Tracker: class_num must be positive");
    for (int i = 0; i < 3; ++i) noise_[i] = observation_noise[i];
}

int Tracker::Track::label() const {       // Features::label(): first maximum of the row sums (features.h:173-178)
    int best = 0;
    for (size_t i = 1; i < feature_sum.size(); ++i)
        if (feature_sum[i] > feature_sum[best]) best = static_cast<int>(i);
    return best;
}

void Tracker::Track::feature(std::vector<float>& out) const {   // features.h:186-197
    float total = 0.f;
    for (float f : feature_sum) total += f;
    out.assign(feature_sum.size(), 0.f);
    if (total == 0.f) return;
    for (size_t i = 0; i < feature_sum.size(); ++i) out[i] = feature_sum[i] / total;
}

// SingerEKF::predict (singer.h:57-101, kalman_filter.h:233-243): x = F x, P = F P F^T + Q
void Tracker::predict(Track& t, int64_t timestamp_ns) const {
    const float dt = static_cast<float>(static_cast<double>(static_cast<float>(timestamp_ns - t.timestamp_ns)) * 1e-9);   // track.h:111-116
    t.timestamp_ns = timestamp_ns;
    const float f02 = dt * dt / 2, f22 = std::exp(-dt / tau_);
    const float a2 = static_cast<float>(std::pow(static_cast<double>(max_acc_), 2));
    const float q00 = static_cast<float>(std::pow(static_cast<double>(dt), 3) / 3) * a2;
    const float q01 = static_cast<float>(std::pow(static_cast<double>(dt), 2) / 2) * a2;
    const float q02 = dt / 2 * a2, q11 = dt * a2, q12 = (1 - std::exp(-dt / tau_)) * a2;
    const float q22 = (1 - std::exp(-2 * dt / tau_)) / 2 * a2;
    float F[81] = {};
    for (int i = 0; i < 9; ++i) F[i * 9 + i] = 1.f;
    float Q[81] = {};
    for (int b = 0; b < 9; b += 3) {
        F[b * 9 + b + 1] = dt;
        F[b * 9 + b + 2] = f02;
        F[(b + 1) * 9 + b + 2] = dt;
        F[(b + 2) * 9 + b + 2] = f22;
        Q[b * 9 + b] = q00;
        Q[b * 9 + b + 1] = Q[(b + 1) * 9 + b] = q01;
        Q[b * 9 + b + 2] = Q[(b + 2) * 9 + b] = q02;
        Q[(b + 1) * 9 + b + 1] = q11;
        Q[(b + 1) * 9 + b + 2] = Q[(b + 2) * 9 + b + 1] = q12;
        Q[(b + 2) * 9 + b + 2] = q22;
    }
    float x[9], FP[81];
    for (int i = 0; i < 9; ++i) {
        float s = 0.f;
        for (int k = 0; k < 9; ++k) s += F[i * 9 + k] * t.x[k];
        x[i] = s;
        for (int j = 0; j < 9; ++j) {
            float p = 0.f;
            for (int k = 0; k < 9; ++k) p += F[i * 9 + k] * t.P[k * 9 + j];
            FP[i * 9 + j] = p;
        }
    }
    for (int i = 0; i < 9; ++i) {
        t.x[i] = x[i];
        for (int j = 0; j < 9; ++j) {
            float p = 0.f;
            for (int k = 0; k < 9; ++k) p += FP[i * 9 + k] * F[j * 9 + k];
            t.P[i * 9 + j] = p + Q[i * 9 + j];
        }
    }
}

// SingerEKF::update (singer.h:103-115, kalman_filter.h:272-293) with H = rows 0, 3, 6 of the identity
void Tracker::correct(Track& t, const float z[3]) const {
    static const int obs[3] = {0, 3, 6};
    float S[9], Sinv[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) S[a * 3 + b] = t.P[obs[a] * 9 + obs[b]] + (a == b ? noise_[a] : 0.f);
    if (!invert3(S, Sinv)) return;
    float K[27];     // K = P H^T S^-1: [9][3]
    for (int i = 0; i < 9; ++i)
        for (int b = 0; b < 3; ++b) {
            float s = 0.f;
            for (int a = 0; a < 3; ++a) s += t.P[i * 9 + obs[a]] * Sinv[a * 3 + b];
            K[i * 3 + b] = s;
        }
    float r[3];
    for (int a = 0; a < 3; ++a) r[a] = z[a] - t.x[obs[a]];
    for (int i = 0; i < 9; ++i) t.x[i] += K[i * 3] * r[0] + K[i * 3 + 1] * r[1] + K[i * 3 + 2] * r[2];
    float P[81];     // P = (I - K H) P
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j) {
            float s = 0.f;
            for (int a = 0; a < 3; ++a) s += K[i * 3 + a] * t.P[obs[a] * 9 + j];
            P[i * 9 + j] = t.P[i * 9 + j] - s;
        }
    std::copy(P, P + 81, t.P);
}

void Tracker::robot_feature(const rmr_robot_t& r, std::vector<float>& out) const {    // robot.cpp:102-122
    out.assign(static_cast<size_t>(class_num_), 0.f);
    if (!r.is_detected) return;
    for (int i = 0; i < r.n_armors && i < RMR_MAX_ARMORS; ++i) {
        const int label = static_cast<int>(r.armors[i].label);
        if (label >= 0 && label < class_num_) out[static_cast<size_t>(label)] += r.armors[i].confidence;
    }
    float sum = 0.f;
    for (float f : out) sum += f;
    if (sum == 0.f) return;
    for (float& f : out) f /= sum;
}

float Tracker::cost(const Track& t, const rmr_robot_t& r, const std::vector<float>& fr) const {   // tracker.cpp:85-118
    if (!r.is_located && !r.is_detected) return 0.f;
    float ds = 0.f;
    if (r.is_located) {
        const float loc[3] = {t.x[0], t.x[3], t.x[6]};
        const float d = distance3(r.location, loc);
        ds = d < dthr_ ? 1.f : (d < 2 * dthr_ ? -d / dthr_ + 2.f : 0.f);
    }
    std::vector<float> ft;
    t.feature(ft);
    float nr = 0.f, nt = 0.f, dot = 0.f;
    for (size_t i = 0; i < fr.size(); ++i) { nr += fr[i] * fr[i]; nt += ft[i] * ft[i]; dot += fr[i] * ft[i]; }
    const float denom = std::sqrt(nr) * std::sqrt(nt);
    const float fs = denom == 0.f ? 0.f : (dot / denom + 1.f) / 2.f;
    return ds * wd_ + fs * wf_;
}

void Tracker::set_track(rmr_robot_t& r, const Track& t, int32_t* state, int32_t* id) {    // robot.cpp:81-94
    if (state) *state = t.state;
    if (id) *id = t.id;
    const bool confirmed = t.state == kConfirmed;
    if (confirmed || r.label < 0) r.label = t.label();
    if (confirmed || !r.is_located) {
        r.location[0] = t.x[0];
        r.location[1] = t.x[3];
        r.location[2] = t.x[6];
        r.is_located = 1;
    }
}

void Tracker::update(rmr_robot_t* robots, int n, int64_t timestamp_ns, int32_t* track_state, int32_t* track_id) {
    if (n < 0 || (n > 0 && robots == nullptr)) throw std::invalid_argument("Tracker::update: bad robot array");
    for (Track& t : tracks_) predict(t, timestamp_ns);
    const int n_tracks = static_cast<int>(tracks_.size());
    std::vector<std::vector<float>> features(static_cast<size_t>(n));
    std::vector<float> costs(static_cast<size_t>(n) * n_tracks);
    for (int r = 0; r < n; ++r) {
        robot_feature(robots[r], features[r]);
        if (track_state) track_state[r] = -1;
        if (track_id) track_id[r] = -1;
        for (int t = 0; t < n_tracks; ++t) costs[static_cast<size_t>(r) * n_tracks + t] = cost(tracks_[t], robots[r], features[r]);
    }
    const std::vector<int> match = auction(costs, n, n_tracks, max_iter_);
    std::vector<int> unmatched;
    std::vector<char> matched(static_cast<size_t>(n_tracks), 0);
    for (int r = 0; r < n; ++r) {
        rmr_robot_t& robot = robots[r];
        const int ti = match[r];
        if (!robot.is_located || ti == kNotMatched) { unmatched.push_back(r); continue; }
        Track& track = tracks_[ti];
        // the auction assigns every agent something, however poor: far away AND another label = not this track (tracker.cpp:158-169)
        const float loc[3] = {track.x[0], track.x[3], track.x[6]};
        if (distance3(robot.location, loc) > 2 * dthr_ && robot.label != track.label()) { unmatched.push_back(r); continue; }
        for (int c = 0; c < class_num_; ++c) track.feature_sum[c] += features[r][c];
        correct(track, robot.location);
        if (track.state == kTentative && ++track.init_count >= init_thresh_) track.state = kConfirmed;
        track.miss_count = 0;
        set_track(robot, track, track_state ? track_state + r : nullptr, track_id ? track_id + r : nullptr);
        matched[ti] = 1;
    }
    for (int t = 0; t < n_tracks; ++t) {
        if (matched[t]) continue;
        Track& track = tracks_[t];
        if (track.state == kTentative) track.state = kDeleted;
        else if (track.state == kConfirmed && ++track.miss_count >= miss_thresh_) track.state = kDeleted;
    }
    tracks_.erase(std::remove_if(tracks_.begin(), tracks_.end(), [](const Track& t) { return t.state == kDeleted; }), tracks_.end());
    for (int r : unmatched) {
        rmr_robot_t& robot = robots[r];
        if (!(robot.is_detected && robot.is_located)) continue;
        Track t;
        t.feature_sum = features[r];
        t.timestamp_ns = timestamp_ns;
        t.id = latest_id_++;
        std::fill(t.x, t.x + 9, 0.f);
        t.x[0] = robot.location[0];
        t.x[3] = robot.location[1];
        t.x[6] = robot.location[2];
        std::fill(t.P, t.P + 81, 0.f);
        for (int i = 0; i < 9; ++i) t.P[i * 9 + i] = 0.1f;             // track.h:55-58
        set_track(robot, t, track_state ? track_state + r : nullptr, track_id ? track_id + r : nullptr);
        tracks_.push_back(std::move(t));
    }
}

std::vector<TrackInfo> Tracker::tracks() const {
    std::vector<TrackInfo> out;
    for (const Track& t : tracks_) {
        TrackInfo i{};
        i.id = t.id;
        i.label = t.label();
        i.state = t.state;
        i.init_count = t.init_count;
        i.miss_count = t.miss_count;
        i.location[0] = t.x[0];
        i.location[1] = t.x[3];
        i.location[2] = t.x[6];
        std::copy(t.x, t.x + 9, i.filter_state);
        out.push_back(i);
    }
    return out;
}

}  // namespace rmr
