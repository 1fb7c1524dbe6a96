// radar::Tracker -- the step after the hot path (see track.cu).  Host code: <= max_cars robots per frame.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/rm_radar_b200.h"

namespace rmr {

enum TrackState : int { kTentative = 0, kConfirmed = 1, kDeleted = 2 };

struct TrackInfo {          // inspection record (tests, visualisation)
    int id, label, state, init_count, miss_count;
    float location[3];
    float filter_state[9];
};

class Tracker {
public:
    Tracker(const float observation_noise[3], int class_num, int init_thresh, int miss_thresh, float max_acceleration,
            float acceleration_correlation_time, float distance_weight, float feature_weight, int max_iter,
            float distance_thresh);
    // Tracker::update(robots, timestamp): robots are updated in place the way Robot::setTrack does it;
    // track_state[i] = -1 (not tracked) / 0 tentative / 1 confirmed, track_id[i] = id or -1
    void update(rmr_robot_t* robots, int n, int64_t timestamp_ns, int32_t* track_state, int32_t* track_id);
    std::vector<TrackInfo> tracks() const;

private:
    struct Track {
        std::vector<float> feature_sum;     // row sums of the reference's feature matrix (features.h:173-197 only reads those)
        int64_t timestamp_ns;
        int id, init_count = 0, miss_count = 0, state = kTentative;
        float x[9];
        float P[81];
        int label() const;
        void feature(std::vector<float>& out) const;
    };
    void predict(Track& t, int64_t timestamp_ns) const;
    void correct(Track& t, const float z[3]) const;
    void robot_feature(const rmr_robot_t& r, std::vector<float>& out) const;
    float cost(const Track& t, const rmr_robot_t& r, const std::vector<float>& robot_feature) const;
    static void set_track(rmr_robot_t& r, const Track& t, int32_t* state, int32_t* id);

    float noise_[3];
    int class_num_, init_thresh_, miss_thresh_;
    float max_acc_, tau_, wd_, wf_;
    int max_iter_;
    float dthr_;
    std::vector<Track> tracks_;
    int latest_id_ = 0;
};

// auction.h:33-126: values[agents][tasks] row-major -> task per agent (-1 = unmatched)
std::vector<int> auction(const std::vector<float>& values, int n_agents, int n_tasks, int max_iter);

}  // namespace rmr
