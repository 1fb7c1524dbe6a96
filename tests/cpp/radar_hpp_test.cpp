// C++ host test of include/radar.hpp: drives the reference-shaped classes the way
// SampleRadar::runOnce does (/root/reference/samples/sample_radar.h:106-127: Locator::update + cluster,
// RobotDetector::detect, Locator::search) and prints one JSON document that tests/test_gpu_cpp_host.py
// compares with the oracle.  Inputs are raw files written by the test (no OpenCV / PCL needed):
//   radar_hpp_test car.rmeng armor.rmeng frame.bgr W H background.f32 cloud.f32
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "radar.hpp"

namespace {
std::vector<char> slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) {
        std::fprintf(stderr, "cannot open %s\n", path);
        std::exit(2);
    }
    return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
}  // namespace

int main(int argc, char** argv) {
    if (argc != 8 && argc != 9) {
        std::fprintf(stderr, "usage: %s car.rmeng armor.rmeng frame.bgr W H background.f32 cloud.f32 [frame.jpg]\n", argv[0]);
        return 2;
    }
    const int W = std::atoi(argv[4]), H = std::atoi(argv[5]);
    const std::vector<char> frame = slurp(argv[3]), bg = slurp(argv[6]), cloud = slurp(argv[7]);
    if (frame.size() != static_cast<size_t>(W) * H * 3) {
        std::fprintf(stderr, "frame size mismatch\n");
        return 2;
    }
    // calibration of the sample (samples/main.cpp:12-21)
    const radar::Matx33f K{1685.51538398561f, 0, 1278.99324114319f, 0, 1685.26471848220f, 1037.21273138299f, 0, 0, 1};
    const radar::Matx44f L2C{0, -1, 0, 0.85443f, 0, 0, -1, -37.6845f, 1, 0, 0, 12.2631f, 0, 0, 0, 1};
    const radar::Matx44f W2C{0.05975021f, 0.99807031f, 0.01689906f, -7179.65399136f, 0.28962566f, -0.00113262f,
                             -0.95713933f, -4671.34956587f, -0.9552732f, 0.06208368f, -0.28913445f, 28286.8920291f,
                             0, 0, 0, 1};
    // constructor failures throw (detector.cpp:80): a missing engine must surface as invalid_argument
    bool threw = false;
    try {
        radar::Detector bad("/nonexistent/model.rmeng", 1, radar::Size{W, H}, 1);
    } catch (const std::invalid_argument&) {
        threw = true;
    }
    radar::RobotDetector detector(argv[1], argv[2], radar::Size{W, H}, 12, 20, 4);   // sample_radar.h:32-34
    radar::Locator locator(W, H, K, L2C, W2C);
    locator.update(radar::CloudView{reinterpret_cast<const float*>(bg.data()), static_cast<int>(bg.size() / 12), 12});
    locator.update(radar::CloudView{reinterpret_cast<const float*>(cloud.data()), static_cast<int>(cloud.size() / 12), 12});
    locator.cluster();
    std::vector<radar::Robot> robots =
        detector.detect(radar::ImageView{reinterpret_cast<const unsigned char*>(frame.data()), W, H, W * 3});
    locator.search(robots);

    // the fused entry point must give the same robots on the same inputs (fresh objects: same locator history)
    bool run_once_same = true;
    {
        radar::RobotDetector det2(argv[1], argv[2], radar::Size{W, H}, 12, 20, 4);
        radar::Locator loc2(W, H, K, L2C, W2C);
        loc2.update(radar::CloudView{reinterpret_cast<const float*>(bg.data()), static_cast<int>(bg.size() / 12), 12});
        std::vector<radar::Robot> r2 = radar::runOnce(
            det2, loc2, radar::ImageView{reinterpret_cast<const unsigned char*>(frame.data()), W, H, W * 3},
            radar::CloudView{reinterpret_cast<const float*>(cloud.data()), static_cast<int>(cloud.size() / 12), 12});
        run_once_same = r2.size() == robots.size();
        for (size_t i = 0; run_once_same && i < robots.size(); ++i) {
            run_once_same = r2[i].label() == robots[i].label() && r2[i].isLocated() == robots[i].isLocated() &&
                            r2[i].rectf()->x == robots[i].rectf()->x && r2[i].rectf()->width == robots[i].rectf()->width;
            if (run_once_same && robots[i].isLocated())
                run_once_same = r2[i].location()->x == robots[i].location()->x && r2[i].location()->z == robots[i].location()->z;
        }
    }
    // throughput mode: two streams fed the same frame and cloud must both give the single-stream robots; then the
    // exchange with a world of one returns this rank's record block
    bool run_batch_same = true, exchange_ok = true;
    {
        radar::RobotDetector det2(argv[1], argv[2], radar::Size{W, H}, 12, 20, 4, 0.75f, 0.65f, 0.25f, 0.65f, 0.50f, 640, 640,
                                  "images", 3, 5, true, 0, 2);
        radar::Locator la(W, H, K, L2C, W2C), lb(W, H, K, L2C, W2C);
        for (radar::Locator* l : {&la, &lb})
            l->update(radar::CloudView{reinterpret_cast<const float*>(bg.data()), static_cast<int>(bg.size() / 12), 12});
        std::vector<char> frames2(frame), clouds2(cloud);
        frames2.insert(frames2.end(), frame.begin(), frame.end());
        clouds2.insert(clouds2.end(), cloud.begin(), cloud.end());
        const auto per_stream = radar::runBatch(
            det2, {&la, &lb}, radar::ImageView{reinterpret_cast<const unsigned char*>(frames2.data()), W, H, W * 3},
            radar::CloudView{reinterpret_cast<const float*>(clouds2.data()), static_cast<int>(cloud.size() / 12), 12});
        run_batch_same = per_stream.size() == 2;
        for (size_t f = 0; run_batch_same && f < 2; ++f) {
            run_batch_same = per_stream[f].size() == robots.size();
            for (size_t i = 0; run_batch_same && i < robots.size(); ++i) {
                const radar::Robot& a = per_stream[f][i];
                run_batch_same = a.label() == robots[i].label() && a.isLocated() == robots[i].isLocated() &&
                                 a.rectf()->x == robots[i].rectf()->x && a.rectf()->height == robots[i].rectf()->height;
                if (run_batch_same && robots[i].isLocated())
                    run_batch_same = a.location()->x == robots[i].location()->x && a.location()->y == robots[i].location()->y;
            }
        }
        radar::Exchange ex(radar::Exchange::uniqueId(), 0, 1, 0, 20);
        ex.publish(robots);
        const std::vector<float> all = ex.collect();
        ex.close();
        exchange_ok = all.size() == 20u * RMR_RECORD_FLOATS;
        for (size_t i = 0; exchange_ok && i < 20; ++i) {
            const float* r = all.data() + i * RMR_RECORD_FLOATS;
            if (i >= robots.size()) { exchange_ok = r[0] == 0.f; continue; }
            exchange_ok = r[0] == 1.f && r[1] == static_cast<float>(robots[i].label().value_or(-1)) &&
                          (r[3] != 0.f) == robots[i].isLocated();
            if (exchange_ok && robots[i].isLocated()) exchange_ok = r[4] == robots[i].location()->x && r[6] == robots[i].location()->z;
        }
    }
    // cv::imread stand-in: frame.jpg decoded on the device must be the bytes of frame.bgr (= cv2.imread of that file),
    // and detect(decoder, file) must give the robots of detect(frame)
    bool jpeg_same = false, jpeg_detect_same = false, jpeg_rejects = false;
    if (argc == 9) {
        const std::vector<char> jpg = slurp(argv[8]);
        radar::JpegDecoder decoder;
        const radar::HostImage img = decoder.imdecode(jpg.data(), jpg.size());
        jpeg_same = img.width == W && img.height == H && std::memcmp(img.data.data(), frame.data(), frame.size()) == 0;
        const std::vector<radar::Robot> r3 = detector.detect(decoder, jpg.data(), jpg.size());
        jpeg_detect_same = r3.size() == robots.size();
        for (size_t i = 0; jpeg_detect_same && i < robots.size(); ++i)
            jpeg_detect_same = r3[i].label() == robots[i].label() && r3[i].rectf()->x == robots[i].rectf()->x &&
                               r3[i].rectf()->height == robots[i].rectf()->height;
        try {
            decoder.imdecode("not a jpeg", 10);
        } catch (const std::invalid_argument&) {
            jpeg_rejects = true;
        }
    }
    std::printf("{\"run_batch_same\": %s, \"exchange_ok\": %s, ", run_batch_same ? "true" : "false", exchange_ok ? "true" : "false");
    std::printf("\"ctor_throws\": %s, \"run_once_same\": %s, \"jpeg_same\": %s, \"jpeg_detect_same\": %s, \"jpeg_rejects\": %s, \"robots\": [",
                threw ? "true" : "false", run_once_same ? "true" : "false", jpeg_same ? "true" : "false",
                jpeg_detect_same ? "true" : "false", jpeg_rejects ? "true" : "false");
    for (size_t i = 0; i < robots.size(); ++i) {
        const radar::Robot& r = robots[i];
        const auto rf = r.rectf().value();
        const auto ri = r.rect().value();
        std::printf("%s{\"rect\": [%.9g, %.9g, %.9g, %.9g], \"rect_int\": [%d, %d, %d, %d], \"label\": %d, "
                    "\"confidence\": %.9g, \"n_armors\": %d, \"located\": %s, \"location\": [%.9g, %.9g, %.9g]}",
                    i ? ", " : "", rf.x, rf.y, rf.width, rf.height, ri.x, ri.y, ri.width, ri.height,
                    r.label().value_or(-1), r.confidence().value_or(0.f),
                    r.isDetected() ? static_cast<int>(r.armors()->size()) : -1, r.isLocated() ? "true" : "false",
                    r.isLocated() ? r.location()->x : 0.f, r.isLocated() ? r.location()->y : 0.f,
                    r.isLocated() ? r.location()->z : 0.f);
    }
    std::printf("]}\n");
    if (!robots.empty()) std::cerr << robots[0] << std::endl;   // operator<< compiles and runs
    return 0;
}
