// Device-resident Locator: LiDAR -> camera projection into a zoomed depth image, running-max
// background, queue-deep foreground differencing, Euclidean clustering of the foreground and per-box
// dominant-cluster centroid -> world.  Replaces /root/reference/src/locate/locate.cpp (OpenCV Matx/Mat,
// PCL EuclideanClusterExtraction + FLANN kd-tree, TBB par_unseq, unordered_map pixel index).
#pragma once
#include <vector>

#include "common.cuh"

namespace rmr {

struct LocatorConfig {
    int image_width = 0, image_height = 0;
    float intrinsic[9];
    float lidar_to_camera[16];
    float world_to_camera[16];
    // defaults: /root/reference/src/locate/locator.h:59-65
    float zoom_factor = 0.5f;
    int queue_size = 3;
    float min_depth_diff = 500.f, max_depth_diff = 4000.f;
    float cluster_tolerance = 400.f;
    int min_cluster_size = 8, max_cluster_size = 1000;
    float max_distance = 29300.f;
};

struct LocateCalib {      // kernel-side constants
    float K[9], L[12], Kinv[9], R[9], t[3];
    double M[12];         // (W2C)^-1 * L2C, rows 0..2
    float zoom, min_diff, max_diff, max_distance, tol;
    int wz, hz, min_size, max_size;
};

struct RectF {
    float x, y, w, h;
    int valid;            // robot.rect().has_value()
};
struct LocResult {
    float x, y, z;        // metres, world (Robot::setLocation, robot.h:93-95)
    int located;          // 0 = no foreground pixel in the box
    int cluster, npoints; // chosen group (-1 = unclustered) and its size
};

class Locator {
public:
    explicit Locator(const LocatorConfig& cfg, int max_points = 1 << 21, int max_foreground = 1 << 17,
                     int max_robots = 64);
    ~Locator();
    Locator(const Locator&) = delete;
    Locator& operator=(const Locator&) = delete;

    // Locator::update (locate.cpp:158-220).  `points`: xyz floats, `stride_floats` apart (PointXYZ = 4).
    void update_host(const float* points, int n, int stride_floats, cudaStream_t s);
    void update_device(const float* dev_points, int n, int stride_floats, cudaStream_t s);
    // Locator::cluster (locate.cpp:231-264)
    void cluster(cudaStream_t s);
    // Locator::search (locate.cpp:276-326); rects/results are host arrays of `n`
    void search(const RectF* rects, LocResult* results, int n, cudaStream_t s);
    // the two halves of search(): everything enqueued / the one wait (several locators overlap their searches)
    void search_begin(const RectF* rects, int n, cudaStream_t s);
    void search_end(LocResult* results, int n, cudaStream_t s);
    int max_robots() const { return max_robots_; }
    void search_device(const RectF* dev_rects, LocResult* dev_results, int n, cudaStream_t s);

    int wz() const { return calib_.wz; }
    int hz() const { return calib_.hz; }
    // inspection (tests): device pointers
    const float* depth_image() const;
    const float* background_image() const { return bg_; }
    float* background_image_mut() { return bg_; }
    float* cloud_buffer() { return cloud_; }   // device staging for host / file clouds: [max_points][4] floats
    int max_points() const { return max_points_; }
    const float* diff_image() const { return diff_; }
    const int* label_image() const { return label_img_; }
    const float* fg_points() const { return fg_pts_; }
    int fg_count_sync(cudaStream_t s);
    int num_clusters_sync(cudaStream_t s);
    const LocateCalib& calib() const { return calib_; }
    void reset();

private:
    LocateCalib calib_{};
    int npix_ = 0, queue_size_ = 3, ring_head_ = 0, ring_count_ = 0;
    int max_points_, max_fg_, max_robots_, hash_size_;
    int nblocks_ = 0;
    float* cloud_ = nullptr;               // device copy of the host cloud
    float* pinned_cloud_ = nullptr;
    cudaEvent_t cloud_uploaded_ = nullptr;   // the pinned staging buffer may be rewritten once this has fired
    unsigned long long* packed_ = nullptr; // (point index + 1) << 32 | depth bits : last writer wins
    float *bg_ = nullptr, *diff_ = nullptr, *ring_ = nullptr;
    int* label_img_ = nullptr;
    int *block_counts_ = nullptr, *block_offsets_ = nullptr;
    int* counters_ = nullptr;              // [0]=fg count, [1]=num clusters, [2]=valid-root count
    float* fg_pts_ = nullptr;              // [max_fg][4]: x,y,z,pixel index bits
    unsigned long long* cell_keys_ = nullptr;   // open-addressing hash: packed (ix, iy, iz) of occupied cells
    int *parent_ = nullptr, *next_ = nullptr, *heads_ = nullptr, *comp_size_ = nullptr, *cluster_id_ = nullptr,
        *root_list_ = nullptr;
    int* hist_ = nullptr;                  // [max_robots][kMaxClusters + 1]
    RectF* dev_rects_ = nullptr;
    LocResult* dev_results_ = nullptr;
    RectF* pinned_rects_ = nullptr;
    LocResult* pinned_results_ = nullptr;
    int* pinned_counters_ = nullptr;       // counters of the last cluster(): [0] foreground kept, [1] clusters kept, [2] clusters found, [3] foreground found
};

constexpr int kMaxClusters = 8191;

}  // namespace rmr
