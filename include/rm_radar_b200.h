/* rm_radar_b200 — C ABI of the B200-native detect + locate hot path.
 *
 * The reference (zmsbruce/rm_radar) exposes this path as C++ classes in four shared libraries
 * (src/radar.h:15-18); it has no FFI layer, so this C ABI is the boundary a binding would target.
 * Each entry point names the reference interface it replaces.  include/radar.hpp rebuilds the
 * reference's C++ class surface (radar::Detector / RobotDetector / Locator / Robot) on top of it.
 *
 * Conventions: every function returns 0 on success or a negative rmr_status; rmr_last_error()
 * returns the message of the last failure on the calling thread.  Handles own all device memory,
 * streams and pinned buffers (reference: detector.cpp:151-161) and are NOT thread-safe, exactly
 * like the reference objects (detector.h: single buffer set).  Inputs are borrowed for the duration
 * of the call.  Images are BGR u8 HWC (cv::Mat CV_8UC3), clouds are float xyz with a byte stride
 * (pcl::PointXYZ = 16).  No CPU fallback exists: without a CUDA device creation fails.
 */
#ifndef RM_RADAR_B200_H
#define RM_RADAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMR_MAX_ARMORS 16

typedef enum rmr_status {
    RMR_OK = 0,
    RMR_ERR_INVALID_ARGUMENT = -1, /* std::invalid_argument in the reference ctors (detector.cpp:80,181) */
    RMR_ERR_RUNTIME = -2,          /* std::runtime_error (common.h:31-39, detector.cpp:184,199,205) */
    RMR_ERR_CUDA = -3,             /* CUDA_CHECK / CUDA_CHECK_NOEXCEPT failures (common.h:31-62) */
    RMR_ERR_CAPACITY = -4          /* a fixed internal capacity was exceeded (candidates / detections per image, armours per
                                      robot, foreground points, clusters): the reference has no such caps and returns every
                                      survivor (detector.cu:561-579), so the call fails instead of returning a truncated result */
} rmr_status;

/* radar::Detection — src/detect/detection.h:25-68 (six floats, standard layout) */
typedef struct rmr_detection {
    float x, y, width, height, label, confidence;
} rmr_detection_t;

/* radar::Robot — src/robot/robot.h:53-164, the fields the hot path fills */
typedef struct rmr_robot {
    float rect[4];             /* rect_  (x, y, w, h) in frame pixels */
    int32_t has_rect;
    int32_t is_detected;       /* armors_.has_value()  (robot.h:65) */
    int32_t label;             /* enum Label, robot.h:32-45 */
    float confidence;
    int32_t n_armors;
    rmr_detection_t armors[RMR_MAX_ARMORS]; /* frame coordinates (robot.cpp:68-73) */
    int32_t is_located;        /* location_.has_value() (robot.h:73) */
    float location[3];         /* metres, world (robot.h:93-95) */
    int32_t cluster;           /* diagnostic: chosen cluster id (-1 = unclustered group) */
    int32_t cluster_points;    /* diagnostic: points averaged */
} rmr_robot_t;

typedef struct rmr_detector rmr_detector_t;
typedef struct rmr_robot_detector rmr_robot_detector_t;
typedef struct rmr_locator rmr_locator_t;
typedef struct rmr_comm rmr_comm_t;

const char* rmr_last_error(void);
int rmr_device_count(int* count);

/* ---- radar::Detector — src/detect/detector.h:84-134 ---------------------------------------- */
/* Detector::Detector(engine_path, classes, image_size, max_batch_size, opt_batch_size, nms_thresh,
 *   conf_thresh, input_width, input_height, input_name, input_channels, opt_level) detector.h:87-93.
 * engine_path: a `.rmeng` file built by `python -m rm_radar_b200.engine model.onnx model.rmeng`
 * (stands where the TensorRT `.engine` cache stands, detector.cpp:74-99).
 * compat != 0 reproduces the reference's int-truncated letterbox geometry. */
int rmr_detector_create(rmr_detector_t** out, const char* engine_path, int classes, int image_width,
                        int image_height, int max_batch_size, float nms_thresh, float conf_thresh,
                        int input_width, int input_height, int compat, int device);
void rmr_detector_destroy(rmr_detector_t* d);
/* Detector::detect(const cv::Mat&) -> std::vector<Detection>  (detector.h:117-134, single image) */
int rmr_detector_detect(rmr_detector_t* d, const uint8_t* bgr, int width, int height, int stride_bytes,
                        rmr_detection_t* out, int capacity, int* count);
/* Detector::detect(container of cv::Mat) -> vector<vector<Detection>>  (batch);
 * out is [n_images][capacity], counts is [n_images] */
int rmr_detector_detect_batch(rmr_detector_t* d, const uint8_t* const* bgr, const int* widths, const int* heights,
                              const int* strides_bytes, int n_images, rmr_detection_t* out, int capacity,
                              int* counts);
/* inspection for parity tests: network input of the last call as float [n][3][H][W] (blobKernel
 * layout, detector.cu:151-171) and the head output [n][4+classes][anchors] (TensorRT output layout) */
int rmr_detector_last_input(rmr_detector_t* d, float* out, int n_images);
int rmr_detector_last_output(rmr_detector_t* d, float* out, int n_images);
int rmr_detector_info(rmr_detector_t* d, int* anchors, int* classes, int* kernel_launches, double* flops_per_image);
/* plan of the network at `batch`: launches per forward, tcgen05 conv launches among them, parallel graph lanes */
int rmr_detector_plan_stats(rmr_detector_t* d, int batch, int* launches, int* umma_convs, int* graph_lanes);
int rmr_detector_set_stream(rmr_detector_t* d, void* cuda_stream);
/* bench: replays the network (the captured conv-stack graph) `iters` times at `batch` on the
 * detector's stream, timed with CUDA events on that stream; ms = average per replay */
int rmr_detector_time_forward(rmr_detector_t* d, int batch, int iters, float* ms);
/* bench / profiling: per-op table of the engine plan at `batch`.  Each row of `rows` is 12 doubles:
 * type (0 conv, 1 maxpool5, 2 upsample2, 3 copy), tcgen05 path used (0/1), h_in, w_in, cin, h_out, w_out,
 * cout, k, stride, flops (2*MAC, conv only), ms per launch (CUDA events, `iters` back-to-back launches). */
int rmr_detector_profile_ops(rmr_detector_t* d, int batch, int iters, double* rows, int capacity, int* n_ops);

/* ---- radar::RobotDetector — src/detect/detector.h:171-190 ---------------------------------- */
/* RobotDetector::RobotDetector(car_path, armor_path, image_size, armor_classes, max_cars, opt_cars,
 *   iou_thresh, car_nms, car_conf, armor_nms, armor_conf, input_width, input_height, ...) */
int rmr_robot_detector_create(rmr_robot_detector_t** out, const char* car_engine, const char* armor_engine,
                              int image_width, int image_height, int armor_classes, int max_cars,
                              float iou_thresh, float car_nms_thresh, float car_conf_thresh,
                              float armor_nms_thresh, float armor_conf_thresh, int input_width, int input_height,
                              int compat, int device);
/* throughput mode (BASELINE config[2]): the same detector for `frames` images per call — the car network runs them as
 * one batch (Detector::detect(container of cv::Mat), detector.cu:439-502), the armor network every ROI of all of them */
int rmr_robot_detector_create_batched(rmr_robot_detector_t** out, const char* car_engine, const char* armor_engine,
                                      int image_width, int image_height, int armor_classes, int max_cars,
                                      float iou_thresh, float car_nms_thresh, float car_conf_thresh,
                                      float armor_nms_thresh, float armor_conf_thresh, int input_width,
                                      int input_height, int compat, int device, int frames);
/* frames: n_frames images of one size back to back (frame i at frames + i * height * stride_bytes), host or device;
 * out is [n_frames][capacity], counts is [n_frames] */
int rmr_robot_detector_detect_frames(rmr_robot_detector_t* d, const void* frames, int frames_on_device, int n_frames,
                                     int width, int height, int stride_bytes, rmr_robot_t* out, int capacity,
                                     int* counts);
void rmr_robot_detector_destroy(rmr_robot_detector_t* d);
/* RobotDetector::detect(const cv::Mat&) -> std::vector<Robot>   (detector.cpp:413-455) */
int rmr_robot_detector_detect(rmr_robot_detector_t* d, const uint8_t* bgr, int width, int height, int stride_bytes,
                              rmr_robot_t* out, int capacity, int* count);
/* same, frame already resident in device memory (HBM): no host->device copy */
int rmr_robot_detector_detect_device(rmr_robot_detector_t* d, const void* dev_bgr, int width, int height,
                                     int stride_bytes, rmr_robot_t* out, int capacity, int* count);
/* diagnostics of the last call: car detections (frame coords) and per-car armour detections (ROI coords) */
int rmr_robot_detector_last_cars(rmr_robot_detector_t* d, rmr_detection_t* out, int capacity, int* count);
int rmr_robot_detector_last_armors(rmr_robot_detector_t* d, int car_index, rmr_detection_t* out, int capacity,
                                   int* count);
int rmr_robot_detector_set_stream(rmr_robot_detector_t* d, void* cuda_stream);
/* kernels launched and conv FLOPs executed by the last detect call (bench accounting) */
int rmr_robot_detector_last_stats(rmr_robot_detector_t* d, int* kernel_launches, double* conv_flops, int* n_cars);
/* device time of the two network replays of the last detect call (CUDA events on the detector's stream) */
int rmr_robot_detector_last_timing(rmr_robot_detector_t* d, float* car_forward_ms, float* armor_forward_ms);
rmr_detector_t* rmr_robot_detector_car(rmr_robot_detector_t* d);
rmr_detector_t* rmr_robot_detector_armor(rmr_robot_detector_t* d);

/* ---- radar::Locator — src/locate/locator.h:53-71 -------------------------------------------- */
/* Locator::Locator(image_width, image_height, intrinsic, lidar_to_camera, world_to_camera, zoom_factor,
 *   queue_size, min_depth_diff, max_depth_diff, cluster_tolerance, min_cluster_size, max_cluster_size,
 *   max_distance)   locator.h:59-65.  Matrices are row-major (cv::Matx33f / Matx44f). */
int rmr_locator_create(rmr_locator_t** out, int image_width, int image_height, const float intrinsic[9],
                       const float lidar_to_camera[16], const float world_to_camera[16], float zoom_factor,
                       int queue_size, float min_depth_diff, float max_depth_diff, float cluster_tolerance,
                       int min_cluster_size, int max_cluster_size, float max_distance, int device);
void rmr_locator_destroy(rmr_locator_t* l);
/* Locator::update(const PointCloud<PointXYZ>::Ptr&)  locate.cpp:158-220; NULL / n == 0 = null / empty cloud */
int rmr_locator_update(rmr_locator_t* l, const float* xyz, int n_points, int stride_bytes);
int rmr_locator_update_device(rmr_locator_t* l, const void* dev_xyz, int n_points, int stride_bytes);
/* Locator::cluster()  locate.cpp:231-264 */
int rmr_locator_cluster(rmr_locator_t* l);
/* Locator::search(std::vector<Robot>&)  locate.cpp:276-326: fills is_located / location */
int rmr_locator_search(rmr_locator_t* l, rmr_robot_t* robots, int n_robots);
/* PCD ingestion: the reference reads its clouds with pcl::io::loadPCDFile
 * (samples/main.cpp:42-72).  `file_bytes` is the whole file image (PCD v0.7, DATA ascii or binary, FIELDS
 * containing x y z); the body is parsed on the device and fed to Locator::update without a host-side cloud. */
int rmr_locator_update_pcd(rmr_locator_t* l, const void* file_bytes, size_t size, int* n_points);
/* the same parser with the points copied back (tests, tools): xyz = float [capacity][3] on the host */
int rmr_pcd_parse(const void* file_bytes, size_t size, float* xyz, int capacity, int* n_points, int device);
/* ---- radar::Tracker — src/track/tracker.h:23-53 (the step after the path) ----
 * Host code (a <= 20 x 20 assignment and a 9-state Singer EKF per track); it consumes the robot records detect +
 * locate have just produced.  Tracker::Tracker(observation_noise, class_num, init_thresh = 4, miss_thresh = 10,
 * max_acceleration = 2, acceleration_correlation_time = 1, distance_weight = 0.4, feature_weight = 0.6,
 * max_iter = 100, distance_thresh = 0.8) — tracker.cpp:47-60. */
typedef struct rmr_tracker rmr_tracker_t;
typedef struct rmr_track {      /* radar::Track, src/track/track.h:27-196 (inspection) */
    int32_t id, label, state;   /* state: 0 tentative, 1 confirmed */
    int32_t init_count, miss_count;
    float location[3];
    float filter_state[9];      /* x vx ax  y vy ay  z vz az */
} rmr_track_t;
int rmr_tracker_create(rmr_tracker_t** out, const float observation_noise[3], int class_num, int init_thresh,
                       int miss_thresh, float max_acceleration, float acceleration_correlation_time,
                       float distance_weight, float feature_weight, int max_iter, float distance_thresh);
void rmr_tracker_destroy(rmr_tracker_t* t);
/* Tracker::update(std::vector<Robot>&, time_point) — tracker.cpp:126-220.  `robots` is updated in place the way
 * Robot::setTrack does it (robot.cpp:81-94: a confirmed track overrides label and location, a tentative one fills
 * what is missing); track_state[i] = -1 not tracked / 0 tentative / 1 confirmed, track_id[i] = id or -1 (both may
 * be NULL).  timestamp_ns stands for the high_resolution_clock time_point. */
int rmr_tracker_update(rmr_tracker_t* t, rmr_robot_t* robots, int n, int64_t timestamp_ns, int32_t* track_state,
                       int32_t* track_id);
int rmr_tracker_tracks(rmr_tracker_t* t, rmr_track_t* out, int capacity, int* count);
/* radar::track::auction — src/track/auction.h:33-126 (values [n_agents][n_tasks] row-major -> task per agent, -1) */
int rmr_auction(const float* values, int n_agents, int n_tasks, int max_iter, int32_t* assignment);

/* ---- JPEG frames ----
 * Replaces cv::imread in front of RobotDetector::detect (samples/main.cpp:24-40): the file image is uploaded as it
 * is (~1/16 of the raw frame) and decoded on the device into the BGR frame the detector reads.  Baseline / extended
 * sequential Huffman JPEG, 8 bit, grayscale or YCbCr 4:4:4 / 4:2:2 / 4:2:0, with or without restart markers;
 * anything else -- and files whose EXIF orientation or RGB coding would make cv::imread return something else than
 * the plain decode -- fails with RMR_ERR_INVALID_ARGUMENT.  Output is bit-identical to libjpeg-turbo's default decode
 * (JDCT_ISLOW, fancy upsampling), i.e. to what cv::imread returns. */
typedef struct rmr_jpeg_decoder rmr_jpeg_decoder_t;
int rmr_jpeg_decoder_create(rmr_jpeg_decoder_t** out, int device);
void rmr_jpeg_decoder_destroy(rmr_jpeg_decoder_t* dec);
int rmr_jpeg_decoder_set_stream(rmr_jpeg_decoder_t* dec, void* cuda_stream);
/* header fields without touching the device */
int rmr_jpeg_info(const void* file_bytes, size_t size, int* width, int* height, int* components, int* h_samp,
                  int* v_samp, int* restart_interval);
/* cv::imread: decode into a host buffer of `capacity` >= width * height * 3 bytes (BGR, packed rows) */
int rmr_jpeg_decode(rmr_jpeg_decoder_t* dec, const void* file_bytes, size_t size, uint8_t* bgr, size_t capacity,
                    int* width, int* height);
/* decode into device memory, asynchronously on the decoder's stream: `dev_bgr` (row pitch `stride_bytes`) or, when
 * dev_bgr is NULL, the decoder's own frame buffer (pitch width * 3), returned through `frame`. */
int rmr_jpeg_decode_device(rmr_jpeg_decoder_t* dec, const void* file_bytes, size_t size, void* dev_bgr,
                           int stride_bytes, const void** frame, int* width, int* height);
/* waits for the last decode; *status = 0 when the entropy-coded data decoded cleanly */
int rmr_jpeg_decoder_status(rmr_jpeg_decoder_t* dec, int* status, int* sync_rounds, int* loop_decodes,
                            int* kernel_launches, size_t* upload_bytes);
/* profiling aid: device ms of the 7 stages of one decode (upload, clear, unstuff, entropy, DC scan, IDCT, colour)
 * followed by the 6 phases of the entropy kernel (pass 0, pass 1, chase, verify loop, scan, write): float[13] */
int rmr_jpeg_decoder_profile(rmr_jpeg_decoder_t* dec, const void* file_bytes, size_t size, float* stage_ms13);
/* test hook: quantised coefficient blocks of the last decode, int16 [n_blocks][64], scan order x natural order */
int rmr_jpeg_decoder_read_coefficients(rmr_jpeg_decoder_t* dec, int16_t* out, long capacity_blocks, long* n_blocks);
/* cv::imread + RobotDetector::detect without the raw frame ever crossing PCIe: decode on the decoder's stream,
 * detect on the detector's stream behind an event */
int rmr_robot_detector_detect_jpeg(rmr_robot_detector_t* d, rmr_jpeg_decoder_t* dec, const void* file_bytes,
                                   size_t size, rmr_robot_t* out, int capacity, int* count);
/* background persistence: the running-max background depth image (locate.cpp:188-191) is the
 * only long-lived Locator state; the reference re-derives it from background.pcd at every start
 * (samples/README.md:3).  save = rmr_locator_read_image(l, 1, out); load replaces it (float [Hz][Wz], zoomed size). */
int rmr_locator_load_background(rmr_locator_t* l, const float* image, int width, int height);
int rmr_locator_set_stream(rmr_locator_t* l, void* cuda_stream);
/* inspection for parity tests.  which: 0 depth, 1 background, 2 diff (float), 3 cluster-label image (int32) */
int rmr_locator_image_size(rmr_locator_t* l, int* width, int* height);
int rmr_locator_read_image(rmr_locator_t* l, int which, void* out);
int rmr_locator_stats(rmr_locator_t* l, int* n_foreground, int* n_clusters);
int rmr_locator_read_foreground(rmr_locator_t* l, float* xyz_pix, int capacity); /* [n][4]: x,y,z,pixel */

/* ---- one frame of the whole path — SampleRadar::runOnce, samples/sample_radar.h:106-127 (without the
 * tracker and the GUI).  The reference runs Locator::update + cluster and RobotDetector::detect on two
 * std::async threads and joins before Locator::search; here one host thread enqueues the car stage, then the
 * locator's work on the locator's own stream, and only then waits — same overlap, no thread hand-off.
 * `frame` / `xyz` are host pointers, or device pointers when the *_on_device flag is set.
 * The detector and the locator must live on the same device and use different streams (the default). */
int rmr_run_once(rmr_robot_detector_t* d, rmr_locator_t* l, const void* frame, int frame_on_device, int width, int height,
                 int stride_bytes, const void* xyz, int cloud_on_device, int n_points, int point_stride_bytes,
                 rmr_robot_t* out, int capacity, int* count);

/* ---- conv layer self-test (tests/bench only): tcgen05 path vs the CUDA-core checker --------- */
/* runs one conv of the given shape on random data through both kernels on the current device and
 * returns the max abs difference; also times the tcgen05 kernel (ms per launch over `iters`). */
int rmr_conv_selftest(int n, int h_in, int w_in, int cin, int cout, int k, int stride, int act, int residual,
                      int out_f32, unsigned seed, int iters, float* max_abs_diff, float* max_ref, float* ms);

/* profiling aid (tests/bench only): per-CTA clock64 timeline of one tcgen05 conv launch, 64 slots per CTA
 * (slot map in csrc/conv.cu); `out` holds capacity_ctas * 64 int64 */
int rmr_conv_timeline(int n, int h_in, int w_in, int cin, int cout, int k, int stride, long long* out,
                      int capacity_ctas, int* n_ctas);

/* SampleRadar::runOnce for n_frames camera + LiDAR streams at once (throughput mode): locators[i] is the Locator of
 * stream i (its own background / depth queue), frames and clouds are back to back (cloud i at
 * xyz + i * n_points * point_stride_bytes).  The car network runs all frames as one batch while the locators update
 * and cluster, the armor network all ROIs, then every locator searches its frame's robots.
 * out is [n_frames][capacity], counts is [n_frames]. */
int rmr_run_batch(rmr_robot_detector_t* d, rmr_locator_t* const* locators, int n_frames, const void* frames,
                  int frames_on_device, int width, int height, int stride_bytes, const void* xyz, int clouds_on_device,
                  int n_points, int point_stride_bytes, rmr_robot_t* out, int capacity, int* counts);

/* post-processing self-test (tests only): the NMS + restore kernel on caller-supplied candidates [n][6] =
 * (x, y, w, h, label, conf) in network coordinates, row index = anchor index (NMSKernel + the NaN filter,
 * detector.cu:315-360, 561-579).  Survivors come back in anchor order.  More candidates than the internal capacity:
 * RMR_ERR_CAPACITY. */
int rmr_postprocess_selftest(const float* candidates, int n, float nms_thresh, rmr_detection_t* out, int capacity, int* count);

/* ---- multi-GPU exchange (one process per GPU, one camera + LiDAR stream per rank) ------------------
 * The path shards by stream with no data-path collective; the only exchange is one NCCL all-gather per step of the
 * fixed-size block of world-frame robot records: 8 floats per robot, max_robots rows per rank,
 *   [valid, label (-1 = undetected), confidence, is_located, x, y, z (metres, world), rect area].
 * The reference has no distributed code; this is what SampleRadar would call after runOnce on every GPU.
 * NCCL is loaded at run time (libnccl.so.2); rank 0 makes the id and the host program hands it to the other ranks. */
#define RMR_COMM_ID_BYTES 128
#define RMR_RECORD_FLOATS 8
int rmr_comm_unique_id(uint8_t* id /* [RMR_COMM_ID_BYTES] */);
int rmr_comm_create(rmr_comm_t** out, const uint8_t* id, int rank, int world, int device, int max_robots);
/* orderly collective shutdown: every rank calls it at the same point of the program (waits for the exchange in flight,
 * then ncclCommDestroy).  rmr_comm_destroy on a communicator that was not closed aborts it (never waits for a peer). */
int rmr_comm_close(rmr_comm_t* c);
void rmr_comm_destroy(rmr_comm_t* c);
/* pack this rank's records and enqueue upload + all-gather + download on the communicator's stream; returns at once.
 * after_stream (may be NULL): a CUDA stream whose work so far the exchange is ordered behind */
int rmr_comm_publish(rmr_comm_t* c, const rmr_robot_t* robots, int n, void* after_stream);
/* wait for the last publish; out is [world][max_robots][RMR_RECORD_FLOATS] */
int rmr_comm_collect(rmr_comm_t* c, float* out);
/* the packing alone (host, no GPU): one rank's block [max_robots][RMR_RECORD_FLOATS] */
int rmr_comm_pack(const rmr_robot_t* robots, int n, int max_robots, float* block);

/* planning aid (tests/tools only, no GPU needed): the launch plan the tcgen05 conv would use for one layer.
 * out[16] = version (1 = conv.cu, 2 = conv2.cu), block_n, splits, halo, m_tiles, ctas, tiles per CTA,
 * k-blocks per tile and CTA, activation slots, weight slots, weights resident (0/1), shared memory bytes,
 * tile w, tile h, tile n, bk */
int rmr_conv_plan(int n, int h_in, int w_in, int cin, int cout, int k, int stride, int* out);

/* ---- engine build (host only, no GPU needed) -------------------------------------------------------------------
 * Replaces the reference's TensorRT build-and-cache step (/root/reference/src/detect/detector.cpp:74-99 engine path
 * resolution, :177-243 buildEngineFromONNX, :281-311 serialise).  rmr_engine_build writes the `.rmeng` plan of
 * `onnx_path` for a network input of input_width x input_height.  rmr_engine_resolve maps the path a caller hands
 * to Detector (`<x>.engine`, `<x>.onnx` or `<x>.rmeng`) to `<x>.rmeng`, building it from `<x>.onnx` when absent
 * (RMR_ERR_INVALID_ARGUMENT when neither exists, like the reference's std::invalid_argument), and copies the NUL-terminated
 * result into out_path (RMR_ERR_CAPACITY if it does not fit).  The detector constructors call it themselves. */
int rmr_engine_build(const char* onnx_path, const char* engine_path, int input_width, int input_height);
int rmr_engine_resolve(const char* path, int input_width, int input_height, char* out_path, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* RM_RADAR_B200_H */
