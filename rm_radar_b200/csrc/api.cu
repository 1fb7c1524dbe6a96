// extern "C" boundary (include/rm_radar_b200.h).  Exceptions become status codes + a thread-local
// message; nothing here computes on the CPU.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "../../include/rm_radar_b200.h"
#include "detector.h"
#include "engine.h"
#include "locate.h"
#include "jpeg.h"
#include "pcd.h"
#include "track.h"
#include "comm.h"

using namespace rmr;

struct rmr_detector {
    std::unique_ptr<Detector> owned;
    Detector* impl = nullptr;
};
struct rmr_robot_detector {
    std::unique_ptr<RobotDetector> impl;
    rmr_detector car_view, armor_view;
    int last_cars = 0;
};
struct rmr_tracker {
    std::unique_ptr<Tracker> impl;
};
struct rmr_jpeg_decoder {
    std::unique_ptr<JpegDecoder> impl;
    cudaEvent_t decoded = nullptr;
};
struct rmr_comm {
    std::unique_ptr<Comm> impl;
};
struct rmr_locator {
    std::unique_ptr<Locator> impl;
    std::unique_ptr<PcdParser> pcd;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    int device = 0;
};

namespace {
thread_local std::string g_last_error;

template <class F>
int guarded(F&& f) {
    try {
        f();
        return RMR_OK;
    } catch (const std::invalid_argument& e) {
        g_last_error = e.what();
        return RMR_ERR_INVALID_ARGUMENT;
    } catch (const CudaError& e) {
        g_last_error = e.what();
        return RMR_ERR_CUDA;
    } catch (const CapacityError& e) {
        g_last_error = e.what();
        return RMR_ERR_CAPACITY;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return RMR_ERR_RUNTIME;
    }
}

void fill_robot(const RobotRecord& r, rmr_robot_t* o) {
    std::memset(o, 0, sizeof(*o));
    std::memcpy(o->rect, r.rect, sizeof(o->rect));
    o->has_rect = r.has_rect;
    o->is_detected = r.detected;
    o->label = r.detected ? r.label : -1;
    o->confidence = r.confidence;
    if (static_cast<int>(r.armors.size()) > RMR_MAX_ARMORS)
        throw CapacityError("a robot carries " + std::to_string(r.armors.size()) + " armour detections, rmr_robot_t holds " +
                            std::to_string(RMR_MAX_ARMORS));
    o->n_armors = static_cast<int>(r.armors.size());
    for (int i = 0; i < o->n_armors; ++i) std::memcpy(&o->armors[i], &r.armors[i], sizeof(rmr_detection_t));
    o->cluster = -2;
}
}  // namespace

extern "C" {

const char* rmr_last_error(void) { return g_last_error.c_str(); }

int rmr_device_count(int* count) {
    return guarded([&] { RMR_CUDA(cudaGetDeviceCount(count)); });
}

// ---------------------------------------------------------------- engine build (host only)
int rmr_engine_build(const char* onnx_path, const char* engine_path, int input_width, int input_height) {
    return guarded([&] {
        if (!onnx_path || !engine_path) throw std::invalid_argument("null argument");
        if (input_width <= 0 || input_height <= 0) throw std::invalid_argument("input size must be positive");
        build_engine(onnx_path, engine_path, input_height, input_width);
    });
}

int rmr_engine_resolve(const char* path, int input_width, int input_height, char* out_path, int capacity) {
    return guarded([&] {
        if (!path || !out_path || capacity <= 0) throw std::invalid_argument("null argument");
        const std::string r = resolve_engine(path, input_height, input_width);
        if (static_cast<int>(r.size()) + 1 > capacity) throw CapacityError("resolved engine path does not fit the buffer");
        std::memcpy(out_path, r.c_str(), r.size() + 1);
    });
}

// ---------------------------------------------------------------- Detector
int rmr_detector_create(rmr_detector_t** out, const char* engine_path, int classes, int image_width,
                        int image_height, int max_batch_size, float nms_thresh, float conf_thresh, int input_width,
                        int input_height, int compat, int device) {
    return guarded([&] {
        if (!out || !engine_path) throw std::invalid_argument("null argument");
        auto h = std::make_unique<rmr_detector>();
        h->owned = std::make_unique<Detector>(engine_path, classes, image_width, image_height, max_batch_size,
                                              nms_thresh, conf_thresh, input_width, input_height, compat != 0, device);
        h->impl = h->owned.get();
        *out = h.release();
    });
}

void rmr_detector_destroy(rmr_detector_t* d) { delete d; }

int rmr_detector_detect(rmr_detector_t* d, const uint8_t* bgr, int width, int height, int stride_bytes,
                        rmr_detection_t* out, int capacity, int* count) {
    return guarded([&] {
        if (!d || !out || !count) throw std::invalid_argument("null argument");
        auto dets = d->impl->detect_host(bgr, width, height, stride_bytes);
        *count = static_cast<int>(dets.size());
        const int n = std::min<int>(*count, capacity);
        std::memcpy(out, dets.data(), sizeof(rmr_detection_t) * n);
    });
}

int rmr_detector_detect_batch(rmr_detector_t* d, const uint8_t* const* bgr, const int* widths, const int* heights,
                              const int* strides_bytes, int n_images, rmr_detection_t* out, int capacity,
                              int* counts) {
    return guarded([&] {
        if (!d || !out || !counts) throw std::invalid_argument("null argument");
        auto res = d->impl->detect_host_batch(bgr, widths, heights, strides_bytes, n_images);
        for (int i = 0; i < n_images; ++i) {
            counts[i] = static_cast<int>(res[i].size());
            const int n = std::min<int>(counts[i], capacity);
            std::memcpy(out + static_cast<size_t>(i) * capacity, res[i].data(), sizeof(rmr_detection_t) * n);
        }
    });
}

int rmr_detector_last_input(rmr_detector_t* d, float* out, int n_images) {
    return guarded([&] { d->impl->last_input(out, n_images); });
}
int rmr_detector_last_output(rmr_detector_t* d, float* out, int n_images) {
    return guarded([&] { d->impl->last_output(out, n_images); });
}
int rmr_detector_info(rmr_detector_t* d, int* anchors, int* classes, int* kernel_launches, double* flops_per_image) {
    return guarded([&] {
        if (anchors) *anchors = d->impl->net().anchors();
        if (classes) *classes = d->impl->classes();
        if (kernel_launches) *kernel_launches = d->impl->last_launches();
        if (flops_per_image) *flops_per_image = d->impl->net().flops_per_image();
    });
}
int rmr_detector_plan_stats(rmr_detector_t* d, int batch, int* launches, int* umma_convs, int* graph_lanes) {
    return guarded([&] {
        if (!d) throw std::invalid_argument("null argument");
        RMR_CUDA(cudaSetDevice(d->impl->device()));
        d->impl->net().plan_stats(batch, launches, umma_convs, graph_lanes);
    });
}
int rmr_detector_set_stream(rmr_detector_t* d, void* cuda_stream) {
    return guarded([&] { d->impl->set_stream(static_cast<cudaStream_t>(cuda_stream)); });
}
int rmr_detector_time_forward(rmr_detector_t* d, int batch, int iters, float* ms) {
    return guarded([&] {
        if (!d || !ms || iters <= 0) throw std::invalid_argument("bad argument");
        RMR_CUDA(cudaSetDevice(d->impl->device()));
        cudaStream_t s = d->impl->stream();
        d->impl->net().forward(batch, s);   // warm-up + graph capture
        d->impl->net().forward(batch, s);
        cudaEvent_t e0, e1;
        RMR_CUDA(cudaEventCreate(&e0));
        RMR_CUDA(cudaEventCreate(&e1));
        RMR_CUDA(cudaEventRecord(e0, s));
        for (int i = 0; i < iters; ++i) d->impl->net().forward(batch, s);
        RMR_CUDA(cudaEventRecord(e1, s));
        RMR_CUDA(cudaEventSynchronize(e1));
        float t = 0.f;
        RMR_CUDA(cudaEventElapsedTime(&t, e0, e1));
        *ms = t / iters;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    });
}

int rmr_detector_profile_ops(rmr_detector_t* d, int batch, int iters, double* rows, int capacity, int* n_ops) {
    return guarded([&] {
        if (!d || !rows || !n_ops) throw std::invalid_argument("null argument");
        RMR_CUDA(cudaSetDevice(d->impl->device()));
        Net& net = d->impl->net();
        const std::vector<float> ms = net.profile_ops(batch, iters, d->impl->stream());
        const auto& ops = net.ops();
        *n_ops = static_cast<int>(ops.size());
        for (int i = 0; i < *n_ops && i < capacity; ++i) {
            const EngineOp& op = ops[i];
            double* r = rows + static_cast<size_t>(i) * 12;
            r[0] = op.type; r[1] = net.op_uses_umma(batch, i) ? 1 : 0;
            r[2] = op.src_h; r[3] = op.src_w; r[4] = op.src_c; r[5] = op.dst_h; r[6] = op.dst_w; r[7] = op.dst_c;
            r[8] = op.k; r[9] = op.stride;
            r[10] = op.type == 0 ? 2.0 * batch * op.dst_h * op.dst_w * static_cast<double>(op.dst_c) * op.k * op.k * op.src_c : 0.0;
            r[11] = ms[i];
        }
    });
}

// ---------------------------------------------------------------- RobotDetector
int rmr_robot_detector_create(rmr_robot_detector_t** out, const char* car_engine, const char* armor_engine,
                              int image_width, int image_height, int armor_classes, int max_cars, float iou_thresh,
                              float car_nms_thresh, float car_conf_thresh, float armor_nms_thresh,
                              float armor_conf_thresh, int input_width, int input_height, int compat, int device) {
    return guarded([&] {
        if (!out || !car_engine || !armor_engine) throw std::invalid_argument("null argument");
        auto h = std::make_unique<rmr_robot_detector>();
        h->impl = std::make_unique<RobotDetector>(car_engine, armor_engine, image_width, image_height, armor_classes,
                                                  max_cars, iou_thresh, car_nms_thresh, car_conf_thresh,
                                                  armor_nms_thresh, armor_conf_thresh, input_width, input_height,
                                                  compat != 0, device);
        h->car_view.impl = &h->impl->car();
        h->armor_view.impl = &h->impl->armor();
        *out = h.release();
    });
}

int rmr_robot_detector_create_batched(rmr_robot_detector_t** out, const char* car_engine, const char* armor_engine,
                                      int image_width, int image_height, int armor_classes, int max_cars, float iou_thresh,
                                      float car_nms_thresh, float car_conf_thresh, float armor_nms_thresh,
                                      float armor_conf_thresh, int input_width, int input_height, int compat, int device,
                                      int frames) {
    return guarded([&] {
        if (!out || !car_engine || !armor_engine) throw std::invalid_argument("null argument");
        auto h = std::make_unique<rmr_robot_detector>();
        h->impl = std::make_unique<RobotDetector>(car_engine, armor_engine, image_width, image_height, armor_classes,
                                                  max_cars, iou_thresh, car_nms_thresh, car_conf_thresh,
                                                  armor_nms_thresh, armor_conf_thresh, input_width, input_height,
                                                  compat != 0, device, frames);
        h->car_view.impl = &h->impl->car();
        h->armor_view.impl = &h->impl->armor();
        *out = h.release();
    });
}

int rmr_robot_detector_detect_frames(rmr_robot_detector_t* d, const void* frames, int frames_on_device, int n_frames,
                                     int width, int height, int stride_bytes, rmr_robot_t* out, int capacity,
                                     int* counts) {
    return guarded([&] {
        if (!d || !out || !counts) throw std::invalid_argument("null argument");
        d->impl->begin_batch(static_cast<const uint8_t*>(frames), frames_on_device != 0, n_frames, width, height, stride_bytes);
        auto robots = d->impl->finish_batch();
        for (int f = 0; f < n_frames; ++f) {
            counts[f] = static_cast<int>(robots[f].size());
            const int n = std::min<int>(counts[f], capacity);
            for (int i = 0; i < n; ++i) fill_robot(robots[f][i], out + static_cast<size_t>(f) * capacity + i);
        }
    });
}

void rmr_robot_detector_destroy(rmr_robot_detector_t* d) { delete d; }

static int robot_detect_common(rmr_robot_detector_t* d, const void* ptr, bool device_ptr, int width, int height,
                               int stride_bytes, rmr_robot_t* out, int capacity, int* count) {
    return guarded([&] {
        if (!d || !out || !count) throw std::invalid_argument("null argument");
        auto robots = device_ptr ? d->impl->detect_device(static_cast<const uint8_t*>(ptr), width, height, stride_bytes)
                                 : d->impl->detect_host(static_cast<const uint8_t*>(ptr), width, height, stride_bytes);
        *count = static_cast<int>(robots.size());
        const int n = std::min<int>(*count, capacity);
        for (int i = 0; i < n; ++i) fill_robot(robots[i], out + i);
    });
}

int rmr_robot_detector_detect(rmr_robot_detector_t* d, const uint8_t* bgr, int width, int height, int stride_bytes,
                              rmr_robot_t* out, int capacity, int* count) {
    return robot_detect_common(d, bgr, false, width, height, stride_bytes, out, capacity, count);
}
int rmr_robot_detector_detect_device(rmr_robot_detector_t* d, const void* dev_bgr, int width, int height,
                                     int stride_bytes, rmr_robot_t* out, int capacity, int* count) {
    return robot_detect_common(d, dev_bgr, true, width, height, stride_bytes, out, capacity, count);
}

int rmr_robot_detector_last_cars(rmr_robot_detector_t* d, rmr_detection_t* out, int capacity, int* count) {
    return guarded([&] {
        const auto& c = d->impl->last_cars();
        *count = static_cast<int>(c.size());
        std::memcpy(out, c.data(), sizeof(rmr_detection_t) * std::min<int>(*count, capacity));
    });
}
int rmr_robot_detector_last_armors(rmr_robot_detector_t* d, int car_index, rmr_detection_t* out, int capacity,
                                   int* count) {
    return guarded([&] {
        const auto& a = d->impl->last_armors();
        if (car_index < 0 || car_index >= static_cast<int>(a.size())) throw std::invalid_argument("car index");
        *count = static_cast<int>(a[car_index].size());
        std::memcpy(out, a[car_index].data(), sizeof(rmr_detection_t) * std::min<int>(*count, capacity));
    });
}
int rmr_robot_detector_set_stream(rmr_robot_detector_t* d, void* cuda_stream) {
    return guarded([&] { d->impl->set_stream(static_cast<cudaStream_t>(cuda_stream)); });
}
int rmr_robot_detector_last_stats(rmr_robot_detector_t* d, int* kernel_launches, double* conv_flops, int* n_cars) {
    return guarded([&] {
        if (kernel_launches) *kernel_launches = d->impl->last_launches();
        if (conv_flops) *conv_flops = d->impl->last_flops();
        if (n_cars) *n_cars = static_cast<int>(d->impl->last_cars().size());
    });
}
int rmr_robot_detector_last_timing(rmr_robot_detector_t* d, float* car_forward_ms, float* armor_forward_ms) {
    return guarded([&] {
        if (!d) throw std::invalid_argument("null argument");
        if (car_forward_ms) *car_forward_ms = d->impl->last_car_ms();
        if (armor_forward_ms) *armor_forward_ms = d->impl->last_armor_ms();
    });
}
rmr_detector_t* rmr_robot_detector_car(rmr_robot_detector_t* d) { return &d->car_view; }
rmr_detector_t* rmr_robot_detector_armor(rmr_robot_detector_t* d) { return &d->armor_view; }

// ---------------------------------------------------------------- Locator
int rmr_locator_create(rmr_locator_t** out, int image_width, int image_height, const float intrinsic[9],
                       const float lidar_to_camera[16], const float world_to_camera[16], float zoom_factor,
                       int queue_size, float min_depth_diff, float max_depth_diff, float cluster_tolerance,
                       int min_cluster_size, int max_cluster_size, float max_distance, int device) {
    return guarded([&] {
        if (!out || !intrinsic || !lidar_to_camera || !world_to_camera) throw std::invalid_argument("null argument");
        RMR_CUDA(cudaSetDevice(device));
        LocatorConfig cfg;
        cfg.image_width = image_width; cfg.image_height = image_height;
        std::memcpy(cfg.intrinsic, intrinsic, sizeof(cfg.intrinsic));
        std::memcpy(cfg.lidar_to_camera, lidar_to_camera, sizeof(cfg.lidar_to_camera));
        std::memcpy(cfg.world_to_camera, world_to_camera, sizeof(cfg.world_to_camera));
        cfg.zoom_factor = zoom_factor; cfg.queue_size = queue_size;
        cfg.min_depth_diff = min_depth_diff; cfg.max_depth_diff = max_depth_diff;
        cfg.cluster_tolerance = cluster_tolerance;
        cfg.min_cluster_size = min_cluster_size; cfg.max_cluster_size = max_cluster_size;
        cfg.max_distance = max_distance;
        auto h = std::make_unique<rmr_locator>();
        h->device = device;
        h->impl = std::make_unique<Locator>(cfg);
        RMR_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
        h->stream = h->own_stream;
        *out = h.release();
    });
}

void rmr_locator_destroy(rmr_locator_t* l) {
    if (!l) return;
    cudaSetDevice(l->device);
    if (l->own_stream) {
        cudaStreamSynchronize(l->own_stream);
        l->pcd.reset();
        l->impl.reset();
        cudaStreamDestroy(l->own_stream);
    }
    delete l;
}

int rmr_locator_update(rmr_locator_t* l, const float* xyz, int n_points, int stride_bytes) {
    return guarded([&] {
        RMR_CUDA(cudaSetDevice(l->device));
        if (xyz && (stride_bytes % 4 != 0 || stride_bytes < 12)) throw std::invalid_argument("bad point stride");
        l->impl->update_host(xyz, n_points, stride_bytes / 4, l->stream);
    });
}
int rmr_locator_update_device(rmr_locator_t* l, const void* dev_xyz, int n_points, int stride_bytes) {
    return guarded([&] {
        RMR_CUDA(cudaSetDevice(l->device));
        if (dev_xyz && (stride_bytes % 4 != 0 || stride_bytes < 12)) throw std::invalid_argument("bad point stride");
        l->impl->update_device(static_cast<const float*>(dev_xyz), n_points, stride_bytes / 4, l->stream);
    });
}
int rmr_locator_cluster(rmr_locator_t* l) {
    return guarded([&] {
        RMR_CUDA(cudaSetDevice(l->device));
        l->impl->cluster(l->stream);
    });
}
int rmr_locator_search(rmr_locator_t* l, rmr_robot_t* robots, int n_robots) {
    return guarded([&] {
        RMR_CUDA(cudaSetDevice(l->device));
        if (n_robots <= 0) return;
        std::vector<RectF> rects(n_robots);
        std::vector<LocResult> res(n_robots);
        for (int i = 0; i < n_robots; ++i)
            rects[i] = RectF{robots[i].rect[0], robots[i].rect[1], robots[i].rect[2], robots[i].rect[3],
                             robots[i].has_rect};
        l->impl->search(rects.data(), res.data(), n_robots, l->stream);
        for (int i = 0; i < n_robots; ++i) {
            if (!res[i].located) continue;   // reference leaves location_ untouched (locate.cpp:300-302)
            robots[i].is_located = 1;
            robots[i].location[0] = res[i].x; robots[i].location[1] = res[i].y; robots[i].location[2] = res[i].z;
            robots[i].cluster = res[i].cluster;
            robots[i].cluster_points = res[i].npoints;
        }
    });
}
int rmr_locator_update_pcd(rmr_locator_t* l, const void* file_bytes, size_t size, int* n_points) {
    return guarded([&] {
        if (!l) throw std::invalid_argument("null argument");
        RMR_CUDA(cudaSetDevice(l->device));
        if (!l->pcd) l->pcd = std::make_unique<PcdParser>();
        const int n = l->pcd->parse(file_bytes, size, l->impl->cloud_buffer(), l->impl->max_points(), l->stream);
        if (n_points) *n_points = n;
        l->impl->update_device(n > 0 ? l->impl->cloud_buffer() : nullptr, n, 3, l->stream);
    });
}
int rmr_pcd_parse(const void* file_bytes, size_t size, float* xyz, int capacity, int* n_points, int device) {
    return guarded([&] {
        if (!xyz || !n_points) throw std::invalid_argument("null argument");
        RMR_CUDA(cudaSetDevice(device));
        const PcdHeader h = pcd_parse_header(file_bytes, size);
        if (h.n_points > capacity) throw std::invalid_argument("PCD: more points than `capacity`");
        PcdParser parser;
        float* dev = nullptr;
        RMR_CUDA(cudaMalloc(&dev, sizeof(float) * 3 * std::max<long>(h.n_points, 1)));
        cudaStream_t s;
        RMR_CUDA(cudaStreamCreate(&s));
        try {
            const int n = parser.parse(file_bytes, size, dev, capacity, s);
            RMR_CUDA(cudaMemcpyAsync(xyz, dev, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, s));
            RMR_CUDA(cudaStreamSynchronize(s));
            *n_points = n;
        } catch (...) {
            cudaStreamDestroy(s);
            cudaFree(dev);
            throw;
        }
        cudaStreamDestroy(s);
        cudaFree(dev);
    });
}
int rmr_tracker_create(rmr_tracker_t** out, const float observation_noise[3], int class_num, int init_thresh,
                       int miss_thresh, float max_acceleration, float acceleration_correlation_time,
                       float distance_weight, float feature_weight, int max_iter, float distance_thresh) {
    return guarded([&] {
        if (!out) throw std::invalid_argument("null argument");
        auto h = std::make_unique<rmr_tracker>();
        h->impl = std::make_unique<Tracker>(observation_noise, class_num, init_thresh, miss_thresh, max_acceleration,
                                            acceleration_correlation_time, distance_weight, feature_weight, max_iter,
                                            distance_thresh);
        *out = h.release();
    });
}
void rmr_tracker_destroy(rmr_tracker_t* t) { delete t; }
int rmr_tracker_update(rmr_tracker_t* t, rmr_robot_t* robots, int n, int64_t timestamp_ns, int32_t* track_state,
                       int32_t* track_id) {
    return guarded([&] {
        if (!t) throw std::invalid_argument("null argument");
        t->impl->update(robots, n, timestamp_ns, track_state, track_id);
    });
}
int rmr_tracker_tracks(rmr_tracker_t* t, rmr_track_t* out, int capacity, int* count) {
    return guarded([&] {
        if (!t || !count) throw std::invalid_argument("null argument");
        const std::vector<TrackInfo> tr = t->impl->tracks();
        *count = static_cast<int>(tr.size());
        for (int i = 0; i < *count && i < capacity && out; ++i) {
            out[i].id = tr[i].id;
            out[i].label = tr[i].label;
            out[i].state = tr[i].state;
            out[i].init_count = tr[i].init_count;
            out[i].miss_count = tr[i].miss_count;
            std::memcpy(out[i].location, tr[i].location, sizeof(out[i].location));
            std::memcpy(out[i].filter_state, tr[i].filter_state, sizeof(out[i].filter_state));
        }
    });
}
int rmr_auction(const float* values, int n_agents, int n_tasks, int max_iter, int32_t* assignment) {
    return guarded([&] {
        if (n_agents < 0 || n_tasks < 0 || (n_agents > 0 && !assignment) || (n_agents * n_tasks > 0 && !values))
            throw std::invalid_argument("bad argument");
        const std::vector<float> v(values, values + static_cast<size_t>(n_agents) * n_tasks);
        const std::vector<int> a = auction(v, n_agents, n_tasks, max_iter);
        for (int i = 0; i < n_agents; ++i) assignment[i] = a[i];
    });
}
int rmr_jpeg_decoder_create(rmr_jpeg_decoder_t** out, int device) {
    return guarded([&] {
        if (!out) throw std::invalid_argument("null argument");
        auto h = std::make_unique<rmr_jpeg_decoder>();
        h->impl = std::make_unique<JpegDecoder>(device);
        RMR_CUDA(cudaEventCreateWithFlags(&h->decoded, cudaEventDisableTiming));
        *out = h.release();
    });
}
void rmr_jpeg_decoder_destroy(rmr_jpeg_decoder_t* dec) {
    if (!dec) return;
    dec->impl.reset();
    if (dec->decoded) cudaEventDestroy(dec->decoded);
    delete dec;
}
int rmr_jpeg_decoder_set_stream(rmr_jpeg_decoder_t* dec, void* cuda_stream) {
    return guarded([&] {
        if (!dec) throw std::invalid_argument("null argument");
        dec->impl->set_stream(static_cast<cudaStream_t>(cuda_stream));
    });
}
int rmr_jpeg_info(const void* file_bytes, size_t size, int* width, int* height, int* components, int* h_samp,
                  int* v_samp, int* restart_interval) {
    return guarded([&] {
        const JpegHeader h = jpeg_parse_header(file_bytes, size);
        if (width) *width = h.width;
        if (height) *height = h.height;
        if (components) *components = h.components;
        if (h_samp) *h_samp = h.h_samp;
        if (v_samp) *v_samp = h.v_samp;
        if (restart_interval) *restart_interval = h.restart_interval;
    });
}
int rmr_jpeg_decode(rmr_jpeg_decoder_t* dec, const void* file_bytes, size_t size, uint8_t* bgr, size_t capacity,
                    int* width, int* height) {
    return guarded([&] {
        if (!dec || !file_bytes) throw std::invalid_argument("null argument");
        dec->impl->decode_to_host(file_bytes, size, bgr, capacity, width, height);
    });
}
int rmr_jpeg_decode_device(rmr_jpeg_decoder_t* dec, const void* file_bytes, size_t size, void* dev_bgr,
                           int stride_bytes, const void** frame, int* width, int* height) {
    return guarded([&] {
        if (!dec || !file_bytes) throw std::invalid_argument("null argument");
        const uint8_t* f = dec->impl->decode(file_bytes, size, static_cast<uint8_t*>(dev_bgr), stride_bytes, width, height);
        if (frame) *frame = f;
    });
}
int rmr_jpeg_decoder_status(rmr_jpeg_decoder_t* dec, int* status, int* sync_rounds, int* loop_decodes,
                            int* kernel_launches, size_t* upload_bytes) {
    return guarded([&] {
        if (!dec) throw std::invalid_argument("null argument");
        const int st = dec->impl->status();
        if (status) *status = st;
        if (sync_rounds) *sync_rounds = dec->impl->last_rounds();
        if (loop_decodes) *loop_decodes = dec->impl->last_loop_decodes();
        if (kernel_launches) *kernel_launches = dec->impl->last_launches();
        if (upload_bytes) *upload_bytes = dec->impl->last_upload_bytes();
    });
}
int rmr_jpeg_decoder_profile(rmr_jpeg_decoder_t* dec, const void* file_bytes, size_t size, float* stage_ms13) {
    return guarded([&] {
        if (!dec || !file_bytes || !stage_ms13) throw std::invalid_argument("null argument");
        dec->impl->profile(file_bytes, size, stage_ms13);
    });
}
int rmr_jpeg_decoder_read_coefficients(rmr_jpeg_decoder_t* dec, int16_t* out, long capacity_blocks, long* n_blocks) {
    return guarded([&] {
        if (!dec) throw std::invalid_argument("null argument");
        const long n = dec->impl->read_coefficients(out, capacity_blocks);
        if (n_blocks) *n_blocks = n;
    });
}
int rmr_robot_detector_detect_jpeg(rmr_robot_detector_t* d, rmr_jpeg_decoder_t* dec, const void* file_bytes,
                                   size_t size, rmr_robot_t* out, int capacity, int* count) {
    if (!d || !dec || !file_bytes) return guarded([] { throw std::invalid_argument("null argument"); });
    int w = 0, h = 0;
    const uint8_t* frame = nullptr;
    const int st = guarded([&] {
        frame = dec->impl->decode(file_bytes, size, nullptr, 0, &w, &h);
        cudaStream_t det_stream = d->impl->car().stream();
        if (det_stream != dec->impl->stream()) {
            RMR_CUDA(cudaEventRecord(dec->decoded, dec->impl->stream()));
            RMR_CUDA(cudaStreamWaitEvent(det_stream, dec->decoded, 0));
        }
    });
    if (st) return st;
    const int st2 = robot_detect_common(d, frame, true, w, h, w * 3, out, capacity, count);
    if (st2) return st2;
    return guarded([&] {
        const int js = dec->impl->status();
        if (js) throw std::runtime_error("jpeg: corrupt entropy-coded data (status " + std::to_string(js) + ")");
    });
}
int rmr_locator_load_background(rmr_locator_t* l, const float* image, int width, int height) {
    return guarded([&] {
        if (!l || !image) throw std::invalid_argument("null argument");
        if (width != l->impl->wz() || height != l->impl->hz())
            throw std::invalid_argument("background image size does not match the zoomed depth image");
        RMR_CUDA(cudaSetDevice(l->device));
        RMR_CUDA(cudaStreamSynchronize(l->stream));
        RMR_CUDA(cudaMemcpy(l->impl->background_image_mut(), image, sizeof(float) * width * height, cudaMemcpyHostToDevice));
    });
}
int rmr_locator_set_stream(rmr_locator_t* l, void* cuda_stream) {
    return guarded([&] { l->stream = static_cast<cudaStream_t>(cuda_stream); });
}
int rmr_locator_image_size(rmr_locator_t* l, int* width, int* height) {
    return guarded([&] { *width = l->impl->wz(); *height = l->impl->hz(); });
}
int rmr_locator_read_image(rmr_locator_t* l, int which, void* out) {
    return guarded([&] {
        RMR_CUDA(cudaSetDevice(l->device));
        const void* src = nullptr;
        switch (which) {
            case 0: src = l->impl->depth_image(); break;
            case 1: src = l->impl->background_image(); break;
            case 2: src = l->impl->diff_image(); break;
            case 3: src = l->impl->label_image(); break;
            default: throw std::invalid_argument("which");
        }
        RMR_CUDA(cudaStreamSynchronize(l->stream));
        RMR_CUDA(cudaMemcpy(out, src, sizeof(float) * l->impl->wz() * l->impl->hz(), cudaMemcpyDeviceToHost));
    });
}
int rmr_locator_stats(rmr_locator_t* l, int* n_foreground, int* n_clusters) {
    return guarded([&] {
        RMR_CUDA(cudaSetDevice(l->device));
        if (n_foreground) *n_foreground = l->impl->fg_count_sync(l->stream);
        if (n_clusters) *n_clusters = l->impl->num_clusters_sync(l->stream);
    });
}
int rmr_locator_read_foreground(rmr_locator_t* l, float* xyz_pix, int capacity) {
    return guarded([&] {
        RMR_CUDA(cudaSetDevice(l->device));
        const int n = std::min(l->impl->fg_count_sync(l->stream), capacity);
        RMR_CUDA(cudaMemcpy(xyz_pix, l->impl->fg_points(), sizeof(float) * 4 * n, cudaMemcpyDeviceToHost));
    });
}

// ---------------------------------------------------------------- one frame (SampleRadar::runOnce)
int rmr_run_once(rmr_robot_detector_t* d, rmr_locator_t* l, const void* frame, int frame_on_device, int width, int height,
                 int stride_bytes, const void* xyz, int cloud_on_device, int n_points, int point_stride_bytes,
                 rmr_robot_t* out, int capacity, int* count) {
    return guarded([&] {
        if (!d || !l || !out || !count) throw std::invalid_argument("null argument");
        if (xyz && (point_stride_bytes % 4 != 0 || point_stride_bytes < 12)) throw std::invalid_argument("bad point stride");
        // RMR_TRACE=1: host wall-clock of the stages of this call on stderr (us since entry), for tuning only
        static const bool trace = [] { const char* e = std::getenv("RMR_TRACE"); return e && (e[0] == '1' || e[0] == '2'); }();
        const auto t0 = std::chrono::steady_clock::now();
        double t_us[5] = {0, 0, 0, 0, 0};
        auto stamp = [&](int i) { if (trace) t_us[i] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); };
        // 1. car stage goes out first (upload + letterbox + network + decode/NMS + D2H), nothing waits
        d->impl->begin(static_cast<const uint8_t*>(frame), frame_on_device != 0, width, height, stride_bytes);
        // 2. the locator's launches are issued while the car network runs (its own stream).  A host cloud queues behind a
        //    host frame: the frame upload is the one copy on the critical path and should not share the link
        //    (RMR_UPLOAD_CONCURRENT=1 lets the two copies overlap, for A/B timing)
        RMR_CUDA(cudaSetDevice(l->device));
        static const bool upload_concurrent = [] { const char* e = std::getenv("RMR_UPLOAD_CONCURRENT"); return e && e[0] == '1'; }();
        if (!frame_on_device && !cloud_on_device && xyz && !upload_concurrent)
            RMR_CUDA(cudaStreamWaitEvent(l->stream, d->impl->frame_uploaded(), 0));
        if (cloud_on_device) l->impl->update_device(static_cast<const float*>(xyz), n_points, point_stride_bytes / 4, l->stream);
        else l->impl->update_host(static_cast<const float*>(xyz), n_points, point_stride_bytes / 4, l->stream);
        l->impl->cluster(l->stream);
        stamp(0);
        // 3. car boxes -> armor stage enqueued.  A robot's rectangle is the box of the car it is made from
        //    (Robot::setDetection, robot.cpp:41-74) and Locator::search treats every rectangle on its own
        //    (locate.cpp:276-326), so the search of all car boxes runs beside the armor network instead of after it
        const std::vector<Detection>& cars = d->impl->cars();
        stamp(1);
        const int n_cars = static_cast<int>(cars.size());
        std::vector<LocResult> res(n_cars);
        const bool early = n_cars > 0 && n_cars <= l->impl->max_robots();
        if (early) {
            std::vector<RectF> rects(n_cars);
            for (int i = 0; i < n_cars; ++i) rects[i] = RectF{cars[i].x, cars[i].y, cars[i].width, cars[i].height, 1};
            l->impl->search_begin(rects.data(), n_cars, l->stream);
        }
        // 4. armors -> robots (label vote, de-duplication)
        stamp(2);
        auto robots = d->impl->finish();
        stamp(3);
        *count = static_cast<int>(robots.size());
        const int n = std::min<int>(*count, capacity);
        for (int i = 0; i < n; ++i) fill_robot(robots[i], out + i);
        // 5. Locator::search results onto the records
        if (early) {
            l->impl->search_end(res.data(), n_cars, l->stream);
        } else if (n > 0) {     // more cars than the locator's search capacity: search the robots that are left
            std::vector<RectF> rects(n);
            for (int i = 0; i < n; ++i)
                rects[i] = RectF{out[i].rect[0], out[i].rect[1], out[i].rect[2], out[i].rect[3], out[i].has_rect};
            res.resize(n);
            l->impl->search(rects.data(), res.data(), n, l->stream);
        } else {
            RMR_CUDA(cudaStreamSynchronize(l->stream));
        }
        stamp(4);
        if (trace)
            std::fprintf(stderr, "rmr_run_once us: enqueued %.1f | cars+armor enqueued %.1f | search enqueued %.1f | armor done %.1f | "
                         "search done %.1f | device car %.1f armor %.1f\n", t_us[0], t_us[1], t_us[2], t_us[3], t_us[4],
                         d->impl->last_car_ms() * 1e3, d->impl->last_armor_ms() * 1e3);
        for (int i = 0; i < n; ++i) {
            const LocResult& r = res[early ? robots[i].car : i];
            if (!r.located) continue;
            out[i].is_located = 1;
            out[i].location[0] = r.x; out[i].location[1] = r.y; out[i].location[2] = r.z;
            out[i].cluster = r.cluster;
            out[i].cluster_points = r.npoints;
        }
    });
}

// ---------------------------------------------------------------- n frames at once (throughput mode)
int rmr_run_batch(rmr_robot_detector_t* d, rmr_locator_t* const* locators, int n_frames, const void* frames,
                  int frames_on_device, int width, int height, int stride_bytes, const void* xyz, int clouds_on_device,
                  int n_points, int point_stride_bytes, rmr_robot_t* out, int capacity, int* counts) {
    return guarded([&] {
        if (!d || !locators || !out || !counts || n_frames <= 0) throw std::invalid_argument("null argument");
        if (xyz && (point_stride_bytes % 4 != 0 || point_stride_bytes < 12)) throw std::invalid_argument("bad point stride");
        for (int f = 0; f < n_frames; ++f)
            if (!locators[f]) throw std::invalid_argument("null locator");
        // 1. the car stage of every frame goes out first, nothing waits
        d->impl->begin_batch(static_cast<const uint8_t*>(frames), frames_on_device != 0, n_frames, width, height, stride_bytes);
        // 2. every stream's locator updates and clusters while the car network runs
        const size_t cloud_floats = static_cast<size_t>(n_points) * (point_stride_bytes / 4);
        for (int f = 0; f < n_frames; ++f) {
            rmr_locator_t* l = locators[f];
            RMR_CUDA(cudaSetDevice(l->device));
            if (!frames_on_device && !clouds_on_device && xyz) RMR_CUDA(cudaStreamWaitEvent(l->stream, d->impl->frame_uploaded(), 0));
            const float* cloud = static_cast<const float*>(xyz) + f * cloud_floats;
            if (clouds_on_device) l->impl->update_device(cloud, n_points, point_stride_bytes / 4, l->stream);
            else l->impl->update_host(cloud, n_points, point_stride_bytes / 4, l->stream);
            l->impl->cluster(l->stream);
        }
        // 3. car boxes of every frame; each stream's search of them runs beside the armor stage (see rmr_run_once)
        const std::vector<std::vector<Detection>>& cars = d->impl->batch_cars();
        std::vector<char> early(n_frames, 0);
        for (int f = 0; f < n_frames; ++f) {
            const int nc = static_cast<int>(cars[f].size());
            if (nc == 0 || nc > locators[f]->impl->max_robots()) continue;
            std::vector<RectF> rects(nc);
            for (int i = 0; i < nc; ++i) rects[i] = RectF{cars[f][i].x, cars[f][i].y, cars[f][i].width, cars[f][i].height, 1};
            locators[f]->impl->search_begin(rects.data(), nc, locators[f]->stream);
            early[f] = 1;
        }
        // 4. armor stage over all ROIs -> robots per frame
        auto robots = d->impl->finish_batch();
        const std::vector<std::vector<Detection>>& cars_done = d->impl->last_batch_cars();
        // 5. search results onto the records (late search only where a frame has more cars than the locator takes)
        std::vector<LocResult> res;
        for (int f = 0; f < n_frames; ++f) {
            counts[f] = static_cast<int>(robots[f].size());
            const int n = std::min<int>(counts[f], capacity);
            rmr_robot_t* o = out + static_cast<size_t>(f) * capacity;
            for (int i = 0; i < n; ++i) fill_robot(robots[f][i], o + i);
            rmr_locator_t* l = locators[f];
            const int nc = static_cast<int>(cars_done[f].size());
            if (early[f]) {
                res.resize(nc);
                l->impl->search_end(res.data(), nc, l->stream);
            } else if (n > 0) {
                std::vector<RectF> rects(n);
                for (int i = 0; i < n; ++i) rects[i] = RectF{o[i].rect[0], o[i].rect[1], o[i].rect[2], o[i].rect[3], o[i].has_rect};
                res.resize(n);
                l->impl->search(rects.data(), res.data(), n, l->stream);
            } else {
                RMR_CUDA(cudaStreamSynchronize(l->stream));
                continue;
            }
            for (int i = 0; i < n; ++i) {
                const LocResult& r = res[early[f] ? robots[f][i].car : i];
                if (!r.located) continue;
                o[i].is_located = 1;
                o[i].location[0] = r.x; o[i].location[1] = r.y; o[i].location[2] = r.z;
                o[i].cluster = r.cluster;
                o[i].cluster_points = r.npoints;
            }
        }
    });
}

// ---------------------------------------------------------------- conv self-test
int rmr_conv_selftest(int n, int h_in, int w_in, int cin, int cout, int k, int stride, int act, int residual,
                      int out_f32, unsigned seed, int iters, float* max_abs_diff, float* max_ref, float* ms) {
    return guarded([&] {
        const int pad = k / 2;
        const int h_out = (h_in + 2 * pad - k) / stride + 1, w_out = (w_in + 2 * pad - k) / stride + 1;
        const int cout_pad = (cout + 15) / 16 * 16;
        // views with channel offsets inside wider buffers, like Concat/Split produce
        const int in_pitch = cin + 16, in_coff = 8, out_pitch = cout_pad + 16, out_coff = out_f32 ? 4 : 8;
        const size_t in_elems = static_cast<size_t>(n) * h_in * w_in * in_pitch;
        const size_t out_elems = static_cast<size_t>(n) * h_out * w_out * out_pitch;
        const size_t w_elems = static_cast<size_t>(cout_pad) * k * k * cin;
        std::mt19937 rng(seed);
        std::uniform_real_distribution<float> dist(-1.f, 1.f);
        std::vector<__half> h_in_buf(in_elems), h_w(w_elems), h_res(out_elems);
        std::vector<float> h_bias(cout_pad);
        for (auto& v : h_in_buf) v = __float2half(dist(rng));
        const float wscale = 1.f / std::sqrt(static_cast<float>(k * k * cin));
        for (size_t i = 0; i < w_elems; ++i) h_w[i] = __float2half(i / (static_cast<size_t>(k) * k * cin) < static_cast<size_t>(cout) ? dist(rng) * wscale * 2.f : 0.f);
        for (auto& v : h_res) v = __float2half(dist(rng));
        for (auto& v : h_bias) v = dist(rng) * 0.1f;
        __half *d_in, *d_w, *d_res;
        float* d_bias;
        void *d_out_a, *d_out_b;
        const size_t out_bytes = out_elems * (out_f32 ? 4 : 2);
        RMR_CUDA(cudaMalloc(&d_in, in_elems * 2));
        RMR_CUDA(cudaMalloc(&d_w, w_elems * 2));
        RMR_CUDA(cudaMalloc(&d_res, out_elems * 2));
        RMR_CUDA(cudaMalloc(&d_bias, cout_pad * 4));
        RMR_CUDA(cudaMalloc(&d_out_a, out_bytes));
        RMR_CUDA(cudaMalloc(&d_out_b, out_bytes));
        RMR_CUDA(cudaMemcpy(d_in, h_in_buf.data(), in_elems * 2, cudaMemcpyHostToDevice));
        RMR_CUDA(cudaMemcpy(d_w, h_w.data(), w_elems * 2, cudaMemcpyHostToDevice));
        RMR_CUDA(cudaMemcpy(d_res, h_res.data(), out_elems * 2, cudaMemcpyHostToDevice));
        RMR_CUDA(cudaMemcpy(d_bias, h_bias.data(), cout_pad * 4, cudaMemcpyHostToDevice));
        RMR_CUDA(cudaMemset(d_out_a, 0, out_bytes));
        RMR_CUDA(cudaMemset(d_out_b, 0, out_bytes));
        ConvDesc d;
        d.in = d_in; d.in_pitch = in_pitch; d.in_coff = in_coff; d.cin = cin; d.h_in = h_in; d.w_in = w_in;
        d.out_pitch = out_pitch; d.out_coff = out_coff; d.cout = cout; d.out_f32 = out_f32;
        d.h_out = h_out; d.w_out = w_out; d.k = k; d.stride = stride; d.act = act;
        if (residual) { d.res = d_res; d.res_pitch = out_pitch; d.res_coff = 8; }
        d.w = d_w; d.bias = d_bias; d.cout_pad = cout_pad; d.cin_pad = cin; d.n = n;
        cudaStream_t s;
        RMR_CUDA(cudaStreamCreate(&s));
        d.out = d_out_a;
        ConvLaunch l = make_conv_launch(d);
        void* d_scratch = nullptr;
        if (conv_scratch_bytes(l)) {
            RMR_CUDA(cudaMalloc(&d_scratch, conv_scratch_bytes(l)));
            RMR_CUDA(cudaMemset(d_scratch, 0, conv_scratch_bytes(l)));
            conv_bind_scratch(l, d_scratch);
        }
        launch_conv_umma(l, s);
        ConvDesc dr = d;
        dr.out = d_out_b;
        launch_conv_simt(dr, s);
        RMR_CUDA(cudaStreamSynchronize(s));
        std::vector<uint8_t> ha(out_bytes), hb(out_bytes);
        RMR_CUDA(cudaMemcpy(ha.data(), d_out_a, out_bytes, cudaMemcpyDeviceToHost));
        RMR_CUDA(cudaMemcpy(hb.data(), d_out_b, out_bytes, cudaMemcpyDeviceToHost));
        float md = 0.f, mr = 0.f;
        for (size_t i = 0; i < out_elems; ++i) {
            const float a = out_f32 ? reinterpret_cast<float*>(ha.data())[i] : __half2float(reinterpret_cast<__half*>(ha.data())[i]);
            const float b = out_f32 ? reinterpret_cast<float*>(hb.data())[i] : __half2float(reinterpret_cast<__half*>(hb.data())[i]);
            const float df = std::fabs(a - b);
            if (!(df <= md)) md = df;   // NaN propagates into md
            mr = std::max(mr, std::fabs(b));
        }
        *max_abs_diff = md;
        *max_ref = mr;
        if (ms) {
            *ms = 0.f;
            if (iters > 0) {
                cudaEvent_t e0, e1;
                RMR_CUDA(cudaEventCreate(&e0));
                RMR_CUDA(cudaEventCreate(&e1));
                RMR_CUDA(cudaEventRecord(e0, s));
                for (int i = 0; i < iters; ++i) launch_conv_umma(l, s);
                RMR_CUDA(cudaEventRecord(e1, s));
                RMR_CUDA(cudaEventSynchronize(e1));
                float t;
                RMR_CUDA(cudaEventElapsedTime(&t, e0, e1));
                *ms = t / iters;
                cudaEventDestroy(e0); cudaEventDestroy(e1);
            }
        }
        cudaStreamDestroy(s);
        cudaFree(d_in); cudaFree(d_w); cudaFree(d_res); cudaFree(d_bias); cudaFree(d_out_a); cudaFree(d_out_b);
        cudaFree(d_scratch);
    });
}

int rmr_conv_timeline(int n, int h_in, int w_in, int cin, int cout, int k, int stride, long long* out,
                      int capacity_ctas, int* n_ctas) {
    return guarded([&] {
        const int pad = k / 2;
        const int h_out = (h_in + 2 * pad - k) / stride + 1, w_out = (w_in + 2 * pad - k) / stride + 1;
        const int cout_pad = (cout + 15) / 16 * 16;
        const size_t in_elems = static_cast<size_t>(n) * h_in * w_in * cin;
        const size_t out_elems = static_cast<size_t>(n) * h_out * w_out * cout_pad;
        const size_t w_elems = static_cast<size_t>(cout_pad) * k * k * cin;
        __half *d_in, *d_w, *d_out;
        float* d_bias;
        RMR_CUDA(cudaMalloc(&d_in, in_elems * 2));
        RMR_CUDA(cudaMalloc(&d_w, w_elems * 2));
        RMR_CUDA(cudaMalloc(&d_out, out_elems * 2));
        RMR_CUDA(cudaMalloc(&d_bias, cout_pad * 4));
        RMR_CUDA(cudaMemset(d_in, 0, in_elems * 2));
        RMR_CUDA(cudaMemset(d_w, 0, w_elems * 2));
        RMR_CUDA(cudaMemset(d_bias, 0, cout_pad * 4));
        ConvDesc d;
        d.in = d_in; d.in_pitch = cin; d.in_coff = 0; d.cin = cin; d.h_in = h_in; d.w_in = w_in;
        d.out = d_out; d.out_pitch = cout_pad; d.out_coff = 0; d.cout = cout; d.out_f32 = 0;
        d.h_out = h_out; d.w_out = w_out; d.k = k; d.stride = stride; d.act = 1;
        d.w = d_w; d.bias = d_bias; d.cout_pad = cout_pad; d.cin_pad = cin; d.n = n;
        ConvLaunch l = make_conv_launch(d);
        void* d_scratch = nullptr;
        if (conv_scratch_bytes(l)) {
            RMR_CUDA(cudaMalloc(&d_scratch, conv_scratch_bytes(l)));
            RMR_CUDA(cudaMemset(d_scratch, 0, conv_scratch_bytes(l)));
            conv_bind_scratch(l, d_scratch);
        }
        const int ctas = static_cast<int>(l.grid.x * l.grid.y * l.grid.z);
        if (l.v2) *n_ctas = -ctas;   // sign tells the caller which slot map applies (conv2.cu)
        long long* d_dbg;
        RMR_CUDA(cudaMalloc(&d_dbg, sizeof(long long) * 64 * ctas));
        RMR_CUDA(cudaMemset(d_dbg, 0, sizeof(long long) * 64 * ctas));
        cudaStream_t s;
        RMR_CUDA(cudaStreamCreate(&s));
        for (int i = 0; i < 3; ++i) launch_conv_umma(l, s);
        l.p.dbg = d_dbg;
        l.q.dbg = d_dbg;
        if (const char* f = std::getenv("RMR_DBG_MODE")) l.q.dbg_mode = std::atoi(f);
        if (const char* f = std::getenv("RMR_DBG_FLAGS")) l.p.dbg_flags = std::atoi(f);
        launch_conv_umma(l, s);
        RMR_CUDA(cudaStreamSynchronize(s));
        *n_ctas = l.v2 ? -ctas : ctas;
        RMR_CUDA(cudaMemcpy(out, d_dbg, sizeof(long long) * 64 * std::min(ctas, capacity_ctas), cudaMemcpyDeviceToHost));
        cudaStreamDestroy(s);
        cudaFree(d_in); cudaFree(d_w); cudaFree(d_out); cudaFree(d_bias); cudaFree(d_dbg); cudaFree(d_scratch);
    });
}

// ---------------------------------------------------------------- multi-GPU exchange
int rmr_comm_unique_id(uint8_t* id) {
    return guarded([&] {
        if (!id) throw std::invalid_argument("null argument");
        comm_unique_id(id);
    });
}

int rmr_comm_create(rmr_comm_t** out, const uint8_t* id, int rank, int world, int device, int max_robots) {
    return guarded([&] {
        if (!out || !id) throw std::invalid_argument("null argument");
        auto h = std::make_unique<rmr_comm>();
        h->impl = std::make_unique<Comm>(id, rank, world, device, max_robots);
        *out = h.release();
    });
}

int rmr_comm_close(rmr_comm_t* c) {
    return guarded([&] {
        if (!c) throw std::invalid_argument("null argument");
        c->impl->close();
    });
}

void rmr_comm_destroy(rmr_comm_t* c) { delete c; }

int rmr_comm_publish(rmr_comm_t* c, const rmr_robot_t* robots, int n, void* after_stream) {
    return guarded([&] {
        if (!c || (!robots && n > 0)) throw std::invalid_argument("null argument");
        c->impl->publish(robots, n, static_cast<cudaStream_t>(after_stream));
    });
}

int rmr_comm_collect(rmr_comm_t* c, float* out) {
    return guarded([&] {
        if (!c || !out) throw std::invalid_argument("null argument");
        c->impl->collect(out);
    });
}

int rmr_comm_pack(const rmr_robot_t* robots, int n, int max_robots, float* block) {
    return guarded([&] {
        if ((!robots && n > 0) || !block || max_robots < 1) throw std::invalid_argument("null argument");
        Comm::pack(robots, n, max_robots, block);
    });
}

int rmr_postprocess_selftest(const float* candidates, int n, float nms_thresh, rmr_detection_t* out, int capacity, int* count) {
    return guarded([&] {
        if ((!candidates && n > 0) || !out || !count) throw std::invalid_argument("null argument");
        auto dets = postprocess_selftest(candidates, n, nms_thresh);
        *count = static_cast<int>(dets.size());
        std::memcpy(out, dets.data(), sizeof(rmr_detection_t) * std::min<int>(*count, capacity));
    });
}

int rmr_conv_plan(int n, int h_in, int w_in, int cin, int cout, int k, int stride, int* out) {
    return guarded([&] {
        if (!out) throw std::invalid_argument("null argument");
        if (n < 1 || h_in < 1 || w_in < 1 || cin < 1 || cout < 1 || (k != 1 && k != 3) || (stride != 1 && stride != 2))
            throw std::invalid_argument("conv plan: batch / extents must be positive, k in {1, 3}, stride in {1, 2}");
        const int pad = k / 2;
        ConvDesc d;
        d.n = n; d.h_in = h_in; d.w_in = w_in; d.cin = cin; d.cin_pad = cin; d.cout = cout; d.cout_pad = (cout + 15) / 16 * 16;
        d.k = k; d.stride = stride; d.act = 1;
        d.h_out = (h_in + 2 * pad - k) / stride + 1; d.w_out = (w_in + 2 * pad - k) / stride + 1;
        d.in_pitch = cin; d.out_pitch = d.cout_pad;
        if (!conv_umma_supported(d)) throw std::invalid_argument("conv shape not supported by the tcgen05 path");
        for (int i = 0; i < 16; ++i) out[i] = 0;
        if (conv2_enabled() && conv2_supported(d)) {
            ConvLaunch l;
            std::memset(&l, 0, sizeof(l));
            plan_conv2(d, l);
            const Conv2Params& q = l.q;
            const int kb = q.units_per_split * (q.halo ? 9 : 1);
            const int v[16] = {2, q.block_n, q.splits, q.halo, q.m_tiles, static_cast<int>(l.grid.x), (q.m_tiles + q.gm - 1) / q.gm, kb,
                               q.sa, q.halo ? q.sb : q.g, q.b_resident, l.smem_bytes, q.tw, q.th, q.tn, q.bk};
            for (int i = 0; i < 16; ++i) out[i] = v[i];
        } else {
            out[0] = 1;   // the round-1 planner encodes tensor maps while planning: needs a device
        }
    });
}

}  // extern "C"
