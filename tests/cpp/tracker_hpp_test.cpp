// C++ host test of radar::Tracker (include/radar.hpp): reads one observation per line
//   frame_ns n | per robot: detected label conf located x y z
// feeds every frame to Tracker::update the way SampleRadar::runOnce does after Locator::search, and prints the robots
// after the update plus the live tracks as JSON lines.  tests/test_track.py compares them with the oracle.  CPU only.
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "radar.hpp"

int main() {
    radar::Tracker tracker(radar::Point3f{0.2f, 0.2f, 0.2f}, 12, 3, 2);
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream in(line);
        long long t;
        int n;
        if (!(in >> t >> n)) break;
        std::vector<rmr_robot_t> recs(static_cast<size_t>(n));
        std::vector<radar::Robot> robots;
        for (int i = 0; i < n; ++i) {
            int detected, label, located;
            float conf, x, y, z;
            in >> detected >> label >> conf >> located >> x >> y >> z;
            rmr_robot_t r{};
            r.label = detected ? label : -1;
            r.is_detected = detected;
            r.confidence = conf;
            if (detected) {
                r.n_armors = 1;
                r.armors[0] = rmr_detection_t{0, 0, 8, 8, static_cast<float>(label), conf};
            }
            r.is_located = located;
            r.location[0] = x; r.location[1] = y; r.location[2] = z;
            robots.push_back(radar::Robot::fromRecord(r));
        }
        tracker.update(robots, t);
        std::printf("{\"robots\": [");
        for (size_t i = 0; i < robots.size(); ++i) {
            const radar::Robot& r = robots[i];
            std::printf("%s{\"state\": %d, \"id\": %d, \"label\": %d, \"located\": %s, \"location\": [%.9g, %.9g, %.9g]}", i ? ", " : "",
                        r.isTracked() ? static_cast<int>(*r.track_state()) : -1, r.track_id().value_or(-1), r.label().value_or(-1),
                        r.isLocated() ? "true" : "false", r.isLocated() ? r.location()->x : 0.f,
                        r.isLocated() ? r.location()->y : 0.f, r.isLocated() ? r.location()->z : 0.f);
        }
        std::printf("], \"tracks\": [");
        const auto tracks = tracker.tracks();
        for (size_t i = 0; i < tracks.size(); ++i)
            std::printf("%s{\"id\": %d, \"label\": %d, \"state\": %d, \"init\": %d, \"miss\": %d}", i ? ", " : "", tracks[i].id,
                        tracks[i].label, tracks[i].state, tracks[i].init_count, tracks[i].miss_count);
        std::printf("]}\n");
    }
    return 0;
}
