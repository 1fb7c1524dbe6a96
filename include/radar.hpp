// radar.hpp — C++ host mirror of the reference's public API for the detect + locate hot path.
//
// The reference surfaces this path through src/radar.h:15-18 as four C++ classes
// (radar::Detector, radar::RobotDetector, radar::Locator, radar::Robot).  This header rebuilds that
// class surface — same names, constructor argument order, defaults, optional-valued getters and
// error behaviour — as a thin, header-only layer over the C ABI in rm_radar_b200.h.  Every method
// forwards to exactly one extern "C" entry point; nothing is computed here.
//
// OpenCV and PCL are not required.  Images and clouds are passed as light views
// (radar::ImageView, radar::CloudView).  Where <opencv2/core.hpp> / <pcl/point_cloud.h> are
// available the reference's own overloads (cv::Mat, cv::Size, cv::Matx, pcl::PointCloud::Ptr)
// are enabled as well, which makes `#include "radar.hpp"` a drop-in for `#include "radar.h"`.
//
// Error behaviour (reference: src/detect/common.h:31-62, detector.cpp:80,181):
//   constructors throw std::invalid_argument / std::runtime_error; detect/update/cluster/search are
//   noexcept and abort on a CUDA failure, printing the message first.
#pragma once

#include <array>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <chrono>
#include <optional>
#include <ostream>
#include <stdexcept>
#include <string>
#include <string_view>
#include <type_traits>
#include <vector>

#include "rm_radar_b200.h"

#if __has_include(<opencv2/core.hpp>) && !defined(RADAR_HPP_NO_OPENCV)
#include <opencv2/core.hpp>
#define RADAR_HPP_HAS_OPENCV 1
#endif
#if __has_include(<pcl/point_cloud.h>) && __has_include(<pcl/point_types.h>) && !defined(RADAR_HPP_NO_PCL)
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#define RADAR_HPP_HAS_PCL 1
#endif

namespace radar {

// ---- light value types standing where cv:: types stand in the reference -------------------------
struct Size {
    int width = 0, height = 0;
};
struct Rect2f {
    float x = 0, y = 0, width = 0, height = 0;
};
struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
};
struct Point3f {
    float x = 0, y = 0, z = 0;
};
using Matx33f = std::array<float, 9>;    // row-major, like cv::Matx33f::val
using Matx44f = std::array<float, 16>;

// BGR u8 HWC image (cv::Mat CV_8UC3): borrowed for the duration of a call
struct ImageView {
    const unsigned char* data = nullptr;
    int width = 0, height = 0;
    int stride_bytes = 0;   // 0 = tightly packed
};
// xyz float points `stride_bytes` apart (pcl::PointXYZ: 16)
struct CloudView {
    const float* xyz = nullptr;
    int size = 0;
    int stride_bytes = 16;
};

// enum Label — src/robot/robot.h:32-45
enum Label {
    BlueHero = 0,
    BlueEngineer = 1,
    BlueInfantryThree = 2,
    BlueInfantryFour = 3,
    BlueInfantryFive = 4,
    RedHero = 5,
    RedEngineer = 6,
    RedInfantryThree = 7,
    RedInfantryFour = 8,
    RedInfantryFive = 9,
    BlueSentry = 10,
    RedSentry = 11
};

// radar::Detection — src/detect/detection.h:25-68: six floats, standard layout, same field order
struct Detection {
    Detection() = default;
    Detection(float x, float y, float width, float height, float label, float confidence)
        : x{x}, y{y}, width{width}, height{height}, label{label}, confidence{confidence} {}
    friend std::ostream& operator<<(std::ostream& os, const Detection& d) {
        return os << "{ x: " << d.x << ", y: " << d.y << ", width: " << d.width << ", height: " << d.height
                  << ", label: " << d.label << ", confidence: " << d.confidence << " }";
    }
    float x = 0, y = 0, width = 0, height = 0, label = 0, confidence = 0;
};
static_assert(std::is_standard_layout_v<Detection> && sizeof(Detection) == sizeof(rmr_detection_t),
              "Detection must stay layout-compatible with rmr_detection_t");

namespace detail {
[[noreturn]] inline void fatal(const char* where) noexcept {
    // CUDA_CHECK_NOEXCEPT — src/detect/common.h:54-62: message on stderr, then abort
    std::fprintf(stderr, "radar: %s failed: %s\n", where, rmr_last_error());
    std::abort();
}
inline void throw_status(int status) {
    if (status == RMR_OK) return;
    const std::string msg = rmr_last_error();
    if (status == RMR_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}
inline int lround_half_even(float v) noexcept {   // cvRound
    return static_cast<int>(__builtin_nearbyintf(v));
}
}  // namespace detail

// radar::Robot — src/robot/robot.h:53-164 (tracking members are out of scope of the hot path)
enum class TrackState { Tentative, Confirmed, Deleted };   // src/track/track.h:25

class Robot {
   public:
    Robot() = default;
    inline bool isDetected() const noexcept { return armors_.has_value(); }
    inline bool isLocated() const noexcept { return location_.has_value(); }
    inline std::optional<int> label() const noexcept { return label_; }
    // Rect2f -> Rect by rounding, like cv::Rect_<int>(cv::Rect2f) (robot.h:111)
    inline std::optional<Rect> rect() const noexcept {
        if (!rect_) return std::nullopt;
        return Rect{detail::lround_half_even(rect_->x), detail::lround_half_even(rect_->y),
                    detail::lround_half_even(rect_->width), detail::lround_half_even(rect_->height)};
    }
    inline std::optional<Rect2f> rectf() const noexcept { return rect_; }
    inline std::optional<float> confidence() const noexcept { return confidence_; }
    inline std::optional<std::vector<Detection>> armors() const noexcept { return armors_; }
    inline std::optional<Point3f> location() const noexcept { return location_; }
    // robot.h:81,139-141: set by Tracker::update (Robot::setTrack, robot.cpp:81-94)
    inline bool isTracked() const noexcept { return track_state_.has_value(); }
    inline std::optional<TrackState> track_state() const noexcept { return track_state_; }
    inline std::optional<int> track_id() const noexcept { return track_id_; }
    // metres, world frame: the C ABI already applied setLocation's mm -> m (robot.h:93-95)
    inline void setLocationMetres(const Point3f& p) noexcept { location_ = p; }

    friend std::ostream& operator<<(std::ostream& os, const Robot& r) {
        os << "Robot: {\n  Label: ";
        if (r.label_) os << *r.label_; else os << "None";
        os << "\n  Rect: ";
        if (r.rect_) os << "[" << r.rect_->x << ", " << r.rect_->y << ", " << r.rect_->width << ", " << r.rect_->height << "]";
        else os << "None";
        os << "\n  Confidence: ";
        if (r.confidence_) os << *r.confidence_; else os << "None";
        os << "\n  Location: ";
        if (r.location_) os << "[" << r.location_->x << ", " << r.location_->y << ", " << r.location_->z << "]";
        else os << "None";
        return os << "\n}";
    }

    // C-ABI record <-> Robot (used by RobotDetector / Locator below)
    static Robot fromRecord(const rmr_robot_t& rec) {
        Robot r;
        if (rec.has_rect) r.rect_ = Rect2f{rec.rect[0], rec.rect[1], rec.rect[2], rec.rect[3]};
        if (rec.is_detected) {
            r.label_ = rec.label;
            r.confidence_ = rec.confidence;
            std::vector<Detection> a(static_cast<size_t>(rec.n_armors));
            for (int i = 0; i < rec.n_armors; ++i) {
                const rmr_detection_t& d = rec.armors[i];
                a[static_cast<size_t>(i)] = Detection(d.x, d.y, d.width, d.height, d.label, d.confidence);
            }
            r.armors_ = std::move(a);
        }
        if (rec.is_located) r.location_ = Point3f{rec.location[0], rec.location[1], rec.location[2]};
        return r;
    }
    void toRecord(rmr_robot_t& rec) const noexcept {
        rec = rmr_robot_t{};
        if (rect_) {
            rec.has_rect = 1;
            rec.rect[0] = rect_->x; rec.rect[1] = rect_->y; rec.rect[2] = rect_->width; rec.rect[3] = rect_->height;
        }
        rec.label = label_.value_or(-1);
        rec.cluster = -2;
    }
    // every field the tracker reads (Robot::feature, robot.cpp:102-122, needs the armours)
    void toFullRecord(rmr_robot_t& rec) const noexcept {
        toRecord(rec);
        if (armors_) {
            rec.is_detected = 1;
            rec.confidence = confidence_.value_or(0.f);
            rec.n_armors = static_cast<int32_t>(std::min<size_t>(armors_->size(), RMR_MAX_ARMORS));
            for (int i = 0; i < rec.n_armors; ++i) {
                const Detection& d = (*armors_)[static_cast<size_t>(i)];
                rec.armors[i] = rmr_detection_t{d.x, d.y, d.width, d.height, d.label, d.confidence};
            }
        }
        if (location_) {
            rec.is_located = 1;
            rec.location[0] = location_->x; rec.location[1] = location_->y; rec.location[2] = location_->z;
        }
    }
    // Robot::setTrack as the C ABI reports it: state / id of the matched track, label and location after the update
    void applyTrack(const rmr_robot_t& rec, int state, int id) noexcept {
        if (state < 0) return;
        track_state_ = static_cast<TrackState>(state);
        track_id_ = id;
        if (rec.label >= 0) label_ = rec.label;
        if (rec.is_located) location_ = Point3f{rec.location[0], rec.location[1], rec.location[2]};
    }

   private:
    std::optional<std::vector<Detection>> armors_ = std::nullopt;
    std::optional<Point3f> location_ = std::nullopt;
    std::optional<Rect2f> rect_ = std::nullopt;
    std::optional<int> label_ = std::nullopt;
    std::optional<float> confidence_ = std::nullopt;
    std::optional<TrackState> track_state_ = std::nullopt;
    std::optional<int> track_id_ = std::nullopt;
};

// radar::Detector — src/detect/detector.h:84-134
class Detector {
   public:
    Detector() = delete;
    Detector(const Detector&) = delete;
    Detector& operator=(const Detector&) = delete;
    // `engine_path`: `<x>.engine`, `<x>.onnx` or `<x>.rmeng`; the library loads the `<x>.rmeng` plan and, like the
    // reference with its TensorRT cache (detector.cpp:74-99), builds it from the sibling `<x>.onnx` on first use
    // (rmr_engine_resolve); neither file -> std::invalid_argument.  opt_batch_size, input_name and
    // opt_level are TensorRT builder knobs: accepted for signature compatibility, unused.
    explicit Detector(std::string_view engine_path, int classes, Size image_size, int max_batch_size,
                      std::optional<int> opt_batch_size = std::nullopt, float nms_thresh = 0.65f,
                      float conf_thresh = 0.25f, int input_width = 640, int input_height = 640,
                      std::string_view input_name = "images", int input_channels = 3, int opt_level = 3,
                      bool compat_letterbox = true, int device = 0) {
        (void)opt_batch_size; (void)input_name; (void)opt_level;
        if (input_channels != 3) throw std::invalid_argument("input_channels must be 3");
        const std::string path(engine_path);
        detail::throw_status(rmr_detector_create(&handle_, path.c_str(), classes, image_size.width,
                                                 image_size.height, max_batch_size, nms_thresh, conf_thresh,
                                                 input_width, input_height, compat_letterbox ? 1 : 0, device));
        max_batch_ = max_batch_size;
    }
    ~Detector() { rmr_detector_destroy(handle_); }

    // detect(const cv::Mat&) -> std::vector<Detection>
    std::vector<Detection> detect(const ImageView& image) noexcept {
        std::vector<Detection> out(kCapacity);
        int n = 0;
        const int stride = image.stride_bytes ? image.stride_bytes : image.width * 3;
        if (rmr_detector_detect(handle_, image.data, image.width, image.height, stride,
                                reinterpret_cast<rmr_detection_t*>(out.data()), kCapacity, &n) != RMR_OK)
            detail::fatal("Detector::detect");
        out.resize(static_cast<size_t>(n < kCapacity ? n : kCapacity));
        return out;
    }
    // detect(container of cv::Mat) -> std::vector<std::vector<Detection>>
    std::vector<std::vector<Detection>> detect(const std::vector<ImageView>& images) noexcept {
        const int k = static_cast<int>(images.size());
        std::vector<std::vector<Detection>> result(images.size());
        if (k == 0) return result;
        std::vector<const uint8_t*> ptr(images.size());
        std::vector<int> w(images.size()), h(images.size()), s(images.size()), counts(images.size());
        for (int i = 0; i < k; ++i) {
            ptr[i] = images[i].data; w[i] = images[i].width; h[i] = images[i].height;
            s[i] = images[i].stride_bytes ? images[i].stride_bytes : images[i].width * 3;
        }
        std::vector<Detection> flat(static_cast<size_t>(k) * kCapacity);
        if (rmr_detector_detect_batch(handle_, ptr.data(), w.data(), h.data(), s.data(), k,
                                      reinterpret_cast<rmr_detection_t*>(flat.data()), kCapacity,
                                      counts.data()) != RMR_OK)
            detail::fatal("Detector::detect(batch)");
        for (int i = 0; i < k; ++i) {
            const int n = counts[i] < kCapacity ? counts[i] : kCapacity;
            result[i].assign(flat.begin() + static_cast<long>(i) * kCapacity,
                             flat.begin() + static_cast<long>(i) * kCapacity + n);
        }
        return result;
    }
#ifdef RADAR_HPP_HAS_OPENCV
    explicit Detector(std::string_view engine_path, int classes, cv::Size image_size, int max_batch_size,
                      std::optional<int> opt_batch_size = std::nullopt, float nms_thresh = 0.65f,
                      float conf_thresh = 0.25f, int input_width = 640, int input_height = 640,
                      std::string_view input_name = "images", int input_channels = 3, int opt_level = 3)
        : Detector(engine_path, classes, Size{image_size.width, image_size.height}, max_batch_size, opt_batch_size,
                   nms_thresh, conf_thresh, input_width, input_height, input_name, input_channels, opt_level) {}
    std::vector<Detection> detect(const cv::Mat& image) noexcept {
        return detect(ImageView{image.data, image.cols, image.rows, static_cast<int>(image.step)});
    }
    std::vector<std::vector<Detection>> detect(const std::vector<cv::Mat>& images) noexcept {
        std::vector<ImageView> v;
        for (const cv::Mat& m : images) v.push_back(ImageView{m.data, m.cols, m.rows, static_cast<int>(m.step)});
        return detect(v);
    }
#endif
    rmr_detector_t* handle() const noexcept { return handle_; }

   private:
    static constexpr int kCapacity = 256;
    rmr_detector_t* handle_ = nullptr;
    int max_batch_ = 0;
};

// radar::RobotDetector — src/detect/detector.h:171-190, detector.cpp:377-455
// cv::imread stand-in for baseline JPEG (samples/main.cpp:24-40): the file image is decoded on the device; the frame
// either comes back to the host (imdecode) or stays in HBM for RobotDetector::detect(JpegDecoder&, ...).
struct HostImage {
    std::vector<unsigned char> data;   // BGR, packed rows
    int width = 0, height = 0;
    ImageView view() const { return ImageView{data.data(), width, height, width * 3}; }
};
class JpegDecoder {
   public:
    JpegDecoder(const JpegDecoder&) = delete;
    JpegDecoder& operator=(const JpegDecoder&) = delete;
    explicit JpegDecoder(int device = 0) { detail::throw_status(rmr_jpeg_decoder_create(&handle_, device)); }
    ~JpegDecoder() { rmr_jpeg_decoder_destroy(handle_); }
    // cv::imdecode(bytes, cv::IMREAD_COLOR); throws std::invalid_argument for files outside the supported subset
    HostImage imdecode(const void* file_bytes, size_t size) {
        HostImage out;
        detail::throw_status(rmr_jpeg_info(file_bytes, size, &out.width, &out.height, nullptr, nullptr, nullptr, nullptr));
        out.data.resize(static_cast<size_t>(out.width) * out.height * 3);
        detail::throw_status(rmr_jpeg_decode(handle_, file_bytes, size, out.data.data(), out.data.size(), &out.width, &out.height));
        return out;
    }
#ifdef RADAR_HPP_HAS_OPENCV
    cv::Mat imdecode(const std::vector<unsigned char>& file) {
        int w = 0, h = 0;
        detail::throw_status(rmr_jpeg_info(file.data(), file.size(), &w, &h, nullptr, nullptr, nullptr, nullptr));
        cv::Mat out(h, w, CV_8UC3);
        detail::throw_status(rmr_jpeg_decode(handle_, file.data(), file.size(), out.data, out.total() * 3, &w, &h));
        return out;
    }
#endif
    rmr_jpeg_decoder_t* handle() const noexcept { return handle_; }

   private:
    rmr_jpeg_decoder_t* handle_ = nullptr;
};

class RobotDetector {
   public:
    RobotDetector() = delete;
    RobotDetector(const RobotDetector&) = delete;
    RobotDetector& operator=(const RobotDetector&) = delete;
    explicit RobotDetector(std::string_view car_engine_path, std::string_view armor_engine_path, Size image_size,
                           int armor_classes, int max_cars, int opt_cars, float iou_thresh = 0.75f,
                           float car_nms_thresh = 0.65f, float car_conf_thresh = 0.25f,
                           float armor_nms_thresh = 0.65f, float armor_conf_thresh = 0.50f, float input_width = 640,
                           float input_height = 640, std::string_view input_name = "images", int input_channels = 3,
                           int opt_level = 5, bool compat_letterbox = true, int device = 0, int frames = 1)
        : max_cars_(max_cars), frames_(frames) {
        (void)opt_cars; (void)input_name; (void)opt_level;
        if (input_channels != 3) throw std::invalid_argument("input_channels must be 3");
        const std::string car(car_engine_path), armor(armor_engine_path);
        // frames > 1: throughput mode, the detector takes that many images per call (detect(std::vector<ImageView>) /
        // runBatch) — the batched Detector::detect of the reference (detector.cu:439-502) carried through the cascade
        detail::throw_status(rmr_robot_detector_create_batched(
            &handle_, car.c_str(), armor.c_str(), image_size.width, image_size.height, armor_classes, max_cars,
            iou_thresh, car_nms_thresh, car_conf_thresh, armor_nms_thresh, armor_conf_thresh,
            static_cast<int>(input_width), static_cast<int>(input_height), compat_letterbox ? 1 : 0, device, frames));
        records_.resize(static_cast<size_t>(max_cars) * static_cast<size_t>(frames));
    }
    ~RobotDetector() { rmr_robot_detector_destroy(handle_); }

    // std::vector<Robot> detect(const cv::Mat&)
    std::vector<Robot> detect(const ImageView& image) {
        int n = 0;
        const int stride = image.stride_bytes ? image.stride_bytes : image.width * 3;
        if (rmr_robot_detector_detect(handle_, image.data, image.width, image.height, stride, records_.data(),
                                      max_cars_, &n) != RMR_OK)
            detail::fatal("RobotDetector::detect");
        std::vector<Robot> robots;
        robots.reserve(static_cast<size_t>(n));
        for (int i = 0; i < n && i < max_cars_; ++i) robots.push_back(Robot::fromRecord(records_[static_cast<size_t>(i)]));
        return robots;
    }
    // `n` frames of one size, back to back in memory (frame i at data + i * height * stride): robots per frame
    std::vector<std::vector<Robot>> detect(const ImageView& first, int n) {
        if (n < 1 || n > frames_) throw std::invalid_argument("RobotDetector::detect: frame count outside [1, frames]");
        std::vector<int> counts(static_cast<size_t>(n));
        const int stride = first.stride_bytes ? first.stride_bytes : first.width * 3;
        if (rmr_robot_detector_detect_frames(handle_, first.data, 0, n, first.width, first.height, stride, records_.data(),
                                             max_cars_, counts.data()) != RMR_OK)
            detail::fatal("RobotDetector::detect(frames)");
        std::vector<std::vector<Robot>> out(static_cast<size_t>(n));
        for (int f = 0; f < n; ++f)
            for (int i = 0; i < counts[static_cast<size_t>(f)] && i < max_cars_; ++i)
                out[static_cast<size_t>(f)].push_back(Robot::fromRecord(records_[static_cast<size_t>(f) * max_cars_ + i]));
        return out;
    }
    int frames() const noexcept { return frames_; }
    int maxCars() const noexcept { return max_cars_; }
    // cv::imread + detect: the JPEG is decoded on the device and the frame never crosses PCIe
    std::vector<Robot> detect(JpegDecoder& decoder, const void* file_bytes, size_t size) {
        int n = 0;
        if (rmr_robot_detector_detect_jpeg(handle_, decoder.handle(), file_bytes, size, records_.data(), max_cars_, &n) != RMR_OK)
            detail::fatal("RobotDetector::detect(jpeg)");
        std::vector<Robot> robots;
        robots.reserve(static_cast<size_t>(n));
        for (int i = 0; i < n && i < max_cars_; ++i) robots.push_back(Robot::fromRecord(records_[static_cast<size_t>(i)]));
        return robots;
    }
#ifdef RADAR_HPP_HAS_OPENCV
    explicit RobotDetector(std::string_view car_engine_path, std::string_view armor_engine_path, cv::Size image_size,
                           int armor_classes, int max_cars, int opt_cars, float iou_thresh = 0.75f,
                           float car_nms_thresh = 0.65f, float car_conf_thresh = 0.25f,
                           float armor_nms_thresh = 0.65f, float armor_conf_thresh = 0.50f, float input_width = 640,
                           float input_height = 640, std::string_view input_name = "images", int input_channels = 3,
                           int opt_level = 5)
        : RobotDetector(car_engine_path, armor_engine_path, Size{image_size.width, image_size.height}, armor_classes,
                        max_cars, opt_cars, iou_thresh, car_nms_thresh, car_conf_thresh, armor_nms_thresh,
                        armor_conf_thresh, input_width, input_height, input_name, input_channels, opt_level) {}
    std::vector<Robot> detect(const cv::Mat& image) {
        return detect(ImageView{image.data, image.cols, image.rows, static_cast<int>(image.step)});
    }
#endif
    rmr_robot_detector_t* handle() const noexcept { return handle_; }

   private:
    rmr_robot_detector_t* handle_ = nullptr;
    int max_cars_ = 0, frames_ = 1;
    std::vector<rmr_robot_t> records_;
};

// radar::Locator — src/locate/locator.h:53-71
class Locator {
   public:
    Locator() = delete;
    Locator(const Locator&) = delete;
    Locator& operator=(const Locator&) = delete;
    Locator(int image_width, int image_height, const Matx33f& intrinsic, const Matx44f& lidar_to_camera,
            const Matx44f& world_to_camera, float zoom_factor = 0.5f, size_t queue_size = 3,
            float min_depth_diff = 500, float max_depth_diff = 4000, float cluster_tolerance = 400,
            int min_cluster_size = 8, int max_cluster_size = 1000, float max_distance = 29300, int device = 0) {
        detail::throw_status(rmr_locator_create(&handle_, image_width, image_height, intrinsic.data(),
                                                lidar_to_camera.data(), world_to_camera.data(), zoom_factor,
                                                static_cast<int>(queue_size), min_depth_diff, max_depth_diff,
                                                cluster_tolerance, min_cluster_size, max_cluster_size, max_distance,
                                                device));
    }
    ~Locator() { rmr_locator_destroy(handle_); }

    // update(const pcl::PointCloud<pcl::PointXYZ>::Ptr&): a null / empty cloud is logged and skipped
    void update(const CloudView& cloud) noexcept {
        if (cloud.xyz == nullptr) std::fprintf(stderr, "radar::Locator::update: cloud is null\n");
        else if (cloud.size == 0) std::fprintf(stderr, "radar::Locator::update: cloud is empty\n");
        if (rmr_locator_update(handle_, cloud.xyz, cloud.size, cloud.stride_bytes) != RMR_OK)
            detail::fatal("Locator::update");
    }
    void cluster() noexcept {
        if (rmr_locator_cluster(handle_) != RMR_OK) detail::fatal("Locator::cluster");
    }
    // search(std::vector<Robot>&): sets location() of every robot whose box holds foreground points
    void search(std::vector<Robot>& robots) const noexcept {
        if (robots.empty()) return;
        std::vector<rmr_robot_t> recs(robots.size());
        for (size_t i = 0; i < robots.size(); ++i) robots[i].toRecord(recs[i]);
        if (rmr_locator_search(handle_, recs.data(), static_cast<int>(recs.size())) != RMR_OK)
            detail::fatal("Locator::search");
        for (size_t i = 0; i < robots.size(); ++i)
            if (recs[i].is_located)
                robots[i].setLocationMetres(Point3f{recs[i].location[0], recs[i].location[1], recs[i].location[2]});
    }
#ifdef RADAR_HPP_HAS_OPENCV
    Locator(int image_width, int image_height, const cv::Matx33f& intrinsic, const cv::Matx44f& lidar_to_camera,
            const cv::Matx44f& world_to_camera, float zoom_factor = 0.5f, size_t queue_size = 3,
            float min_depth_diff = 500, float max_depth_diff = 4000, float cluster_tolerance = 400,
            int min_cluster_size = 8, int max_cluster_size = 1000, float max_distance = 29300)
        : Locator(image_width, image_height, toArray<9>(intrinsic.val), toArray<16>(lidar_to_camera.val),
                  toArray<16>(world_to_camera.val), zoom_factor, queue_size, min_depth_diff, max_depth_diff,
                  cluster_tolerance, min_cluster_size, max_cluster_size, max_distance) {}
#endif
#ifdef RADAR_HPP_HAS_PCL
    void update(const pcl::PointCloud<pcl::PointXYZ>::Ptr& cloud) noexcept {
        if (!cloud) { update(CloudView{nullptr, 0, 16}); return; }
        update(CloudView{cloud->empty() ? nullptr : &cloud->points[0].x, static_cast<int>(cloud->size()),
                         static_cast<int>(sizeof(pcl::PointXYZ))});
    }
#endif
    rmr_locator_t* handle() const noexcept { return handle_; }

   private:
    template <size_t N>
    static std::array<float, N> toArray(const float* v) {
        std::array<float, N> a;
        for (size_t i = 0; i < N; ++i) a[i] = v[i];
        return a;
    }
    rmr_locator_t* handle_ = nullptr;
};

// One frame of the whole path — the body of SampleRadar::runOnce (samples/sample_radar.h:106-127) without the
// tracker and the GUI: Locator::update + cluster overlap with RobotDetector::detect, then Locator::search.
// radar::Tracker -- src/track/tracker.h:23-53.  Same constructor arguments and defaults; update() takes the robots of
// one frame and their time stamp, and marks / completes them the way Robot::setTrack does.
class Tracker {
   public:
    Tracker(const Tracker&) = delete;
    Tracker& operator=(const Tracker&) = delete;
    Tracker(const Point3f& observation_noise, int class_num, int init_thresh = 4, int miss_thresh = 10,
            float max_acceleration = 2.0f, float acceleration_correlation_time = 1.0f, float distance_weight = 0.40f,
            float feature_weight = 0.60f, int max_iter = 100, float distance_thresh = 0.8f) {
        const float noise[3] = {observation_noise.x, observation_noise.y, observation_noise.z};
        detail::throw_status(rmr_tracker_create(&handle_, noise, class_num, init_thresh, miss_thresh, max_acceleration,
                                                acceleration_correlation_time, distance_weight, feature_weight, max_iter,
                                                distance_thresh));
    }
    ~Tracker() { rmr_tracker_destroy(handle_); }

    void update(std::vector<Robot>& robots, const std::chrono::high_resolution_clock::time_point& timestamp) {
        update(robots, std::chrono::duration_cast<std::chrono::nanoseconds>(timestamp.time_since_epoch()).count());
    }
    void update(std::vector<Robot>& robots, long long timestamp_ns) {
        const int n = static_cast<int>(robots.size());
        records_.resize(robots.size());
        state_.assign(robots.size(), -1);
        id_.assign(robots.size(), -1);
        for (int i = 0; i < n; ++i) robots[static_cast<size_t>(i)].toFullRecord(records_[static_cast<size_t>(i)]);
        if (rmr_tracker_update(handle_, records_.data(), n, timestamp_ns, state_.data(), id_.data()) != RMR_OK)
            detail::fatal("Tracker::update");
        for (int i = 0; i < n; ++i)
            robots[static_cast<size_t>(i)].applyTrack(records_[static_cast<size_t>(i)], state_[static_cast<size_t>(i)],
                                                      id_[static_cast<size_t>(i)]);
    }
    std::vector<rmr_track_t> tracks() const {
        int n = 0;
        std::vector<rmr_track_t> out(64);
        detail::throw_status(rmr_tracker_tracks(handle_, out.data(), 64, &n));
        out.resize(static_cast<size_t>(std::min(n, 64)));
        return out;
    }
    rmr_tracker_t* handle() const noexcept { return handle_; }

   private:
    rmr_tracker_t* handle_ = nullptr;
    std::vector<rmr_robot_t> records_;
    std::vector<int32_t> state_, id_;
};

inline std::vector<Robot> runOnce(RobotDetector& detector, Locator& locator, const ImageView& image,
                                  const CloudView& cloud, int max_robots = 64) {
    std::vector<rmr_robot_t> recs(static_cast<size_t>(max_robots));
    int n = 0;
    const int stride = image.stride_bytes ? image.stride_bytes : image.width * 3;
    if (rmr_run_once(detector.handle(), locator.handle(), image.data, 0, image.width, image.height, stride, cloud.xyz, 0,
                     cloud.size, cloud.stride_bytes, recs.data(), max_robots, &n) != RMR_OK)
        detail::fatal("runOnce");
    std::vector<Robot> robots;
    for (int i = 0; i < n && i < max_robots; ++i) robots.push_back(Robot::fromRecord(recs[static_cast<size_t>(i)]));
    return robots;
}

// SampleRadar::runOnce as the sample writes it, tracker included (sample_radar.h:106-127)
inline std::vector<Robot> runOnce(RobotDetector& detector, Locator& locator, Tracker& tracker, const ImageView& image,
                                  const CloudView& cloud, const std::chrono::high_resolution_clock::time_point& timestamp,
                                  int max_robots = 64) {
    std::vector<Robot> robots = runOnce(detector, locator, image, cloud, max_robots);
    tracker.update(robots, timestamp);
    return robots;
}

// SampleRadar::runOnce for several camera + LiDAR streams in one call (throughput mode): stream i has its own Locator
// (background, depth queue); frames of one size back to back (frame i at first.data + i * height * stride), clouds of
// one size back to back (cloud i at first_cloud.xyz + i * size * stride floats).  Robots per stream.
inline std::vector<std::vector<Robot>> runBatch(RobotDetector& detector, const std::vector<Locator*>& locators,
                                                const ImageView& first, const CloudView& first_cloud) {
    const int n = static_cast<int>(locators.size());
    if (n < 1 || n > detector.frames()) throw std::invalid_argument("runBatch: stream count outside [1, frames]");
    std::vector<rmr_locator_t*> handles(locators.size());
    for (size_t i = 0; i < locators.size(); ++i) {
        if (!locators[i]) throw std::invalid_argument("runBatch: null locator");
        handles[i] = locators[i]->handle();
    }
    const int cap = detector.maxCars();
    std::vector<rmr_robot_t> recs(static_cast<size_t>(n) * static_cast<size_t>(cap));
    std::vector<int> counts(static_cast<size_t>(n));
    const int stride = first.stride_bytes ? first.stride_bytes : first.width * 3;
    if (rmr_run_batch(detector.handle(), handles.data(), n, first.data, 0, first.width, first.height, stride,
                      first_cloud.xyz, 0, first_cloud.size, first_cloud.stride_bytes, recs.data(), cap,
                      counts.data()) != RMR_OK)
        detail::fatal("runBatch");
    std::vector<std::vector<Robot>> out(static_cast<size_t>(n));
    for (int f = 0; f < n; ++f)
        for (int i = 0; i < counts[static_cast<size_t>(f)] && i < cap; ++i)
            out[static_cast<size_t>(f)].push_back(Robot::fromRecord(recs[static_cast<size_t>(f) * cap + i]));
    return out;
}

// The multi-GPU exchange (one process per GPU, one camera + LiDAR stream per rank): what SampleRadar would call after
// runOnce on every GPU.  The reference has no distributed code; the record block is rmr_comm_* 's
// [valid, label, confidence, is_located, x, y, z, rect area] per robot.  Rank 0 makes the id (Exchange::uniqueId) and
// the host program hands it to the other ranks (MPI, a socket, torch.distributed in bench.py).
class Exchange {
   public:
    using Id = std::array<uint8_t, RMR_COMM_ID_BYTES>;
    static Id uniqueId() {
        Id id{};
        detail::throw_status(rmr_comm_unique_id(id.data()));
        return id;
    }
    Exchange(const Exchange&) = delete;
    Exchange& operator=(const Exchange&) = delete;
    Exchange(const Id& id, int rank, int world, int device = 0, int max_robots = 20)
        : world_(world), max_robots_(max_robots) {
        detail::throw_status(rmr_comm_create(&handle_, id.data(), rank, world, device, max_robots));
    }
    ~Exchange() { rmr_comm_destroy(handle_); }
    // enqueue this rank's robots for the all-gather and return at once (it overlaps the next frame)
    void publish(const std::vector<Robot>& robots) {
        records_.resize(robots.size());
        for (size_t i = 0; i < robots.size(); ++i) robots[i].toFullRecord(records_[i]);
        if (rmr_comm_publish(handle_, records_.data(), static_cast<int>(robots.size()), nullptr) != RMR_OK)
            detail::fatal("Exchange::publish");
    }
    // wait for the last publish: [world][max_robots][RMR_RECORD_FLOATS]
    std::vector<float> collect() {
        std::vector<float> all(static_cast<size_t>(world_) * static_cast<size_t>(max_robots_) * RMR_RECORD_FLOATS);
        if (rmr_comm_collect(handle_, all.data()) != RMR_OK) detail::fatal("Exchange::collect");
        return all;
    }
    // orderly collective shutdown: every rank, at the same point of the program
    void close() { detail::throw_status(rmr_comm_close(handle_)); }

   private:
    rmr_comm_t* handle_ = nullptr;
    int world_ = 1, max_robots_ = 20;
    std::vector<rmr_robot_t> records_;
};

}  // namespace radar
