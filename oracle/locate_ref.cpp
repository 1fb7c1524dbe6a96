// ORACLE / CPU BASELINE (test + bench infrastructure only — never linked into the product library).
//
// C++ restatement ("port") of the reference's CPU locate path, /root/reference/src/locate/locate.cpp,
// keeping its structure and data structures so that timing it says something about the reference:
//   * update():  per-point projection loop (locate.cpp:173-193) + one full-image diff pass per queued
//                frame (locate.cpp:200-219); std::thread chunks stand in for TBB par_unseq
//   * cluster(): serial row-major scan + unordered_map<Point2i,int> with the reference's hx^hy hash
//                (locator.h:38-44, locate.cpp:237-250), kd-tree radius search + BFS region growing,
//                size filter and size-descending sort = PCL EuclideanClusterExtraction's published
//                algorithm (PCL itself is an un-vendored dependency, absent from this image)
//   * search():  per-box std::map<int, vector<Point3f>> grouping, first-max group, mean, lidarToWorld
//                (locate.cpp:276-311)
// Semantics fixed where the reference races: sequential cloud order (last writer wins), frames
// oldest -> newest; `>=` bounds (B#12).  Arithmetic order matches
// oracle/locate_oracle.py (no FMA contraction: built with -ffp-contract=off).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <numeric>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct P3 { float x, y, z; };
struct Pix { int x, y; bool operator==(const Pix& o) const { return x == o.x && y == o.y; } };
struct PixHash {   // locator.h:38-44: hash<int>(x) ^ hash<int>(y)
    size_t operator()(const Pix& p) const { return std::hash<int>()(p.x) ^ std::hash<int>()(p.y); }
};

struct KdTree {
    struct Node { int lo, hi, axis; float split; int left, right; };
    const std::vector<P3>* pts = nullptr;
    std::vector<int> idx;
    std::vector<Node> nodes;
    static float coord(const P3& p, int a) { return a == 0 ? p.x : (a == 1 ? p.y : p.z); }
    int build(int lo, int hi) {
        Node n{lo, hi, -1, 0.f, -1, -1};
        const int id = static_cast<int>(nodes.size());
        nodes.push_back(n);
        if (hi - lo > 15) {   // FLANN KDTreeSingleIndex default leaf size
            float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
            for (int i = lo; i < hi; ++i)
                for (int a = 0; a < 3; ++a) {
                    const float c = coord((*pts)[idx[i]], a);
                    mn[a] = std::min(mn[a], c); mx[a] = std::max(mx[a], c);
                }
            int axis = 0;
            for (int a = 1; a < 3; ++a) if (mx[a] - mn[a] > mx[axis] - mn[axis]) axis = a;
            const int mid = (lo + hi) / 2;
            std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi,
                             [&](int a, int b) { return coord((*pts)[a], axis) < coord((*pts)[b], axis); });
            nodes[id].axis = axis;
            nodes[id].split = coord((*pts)[idx[mid]], axis);
            const int l = build(lo, mid);
            const int r = build(mid, hi);
            nodes[id].left = l; nodes[id].right = r;
        }
        return id;
    }
    void set(const std::vector<P3>& p) {
        pts = &p; idx.resize(p.size()); std::iota(idx.begin(), idx.end(), 0); nodes.clear();
        if (!p.empty()) build(0, static_cast<int>(p.size()));
    }
    void radius(int node, const P3& q, float r, float r2, std::vector<int>& out) const {
        const Node& n = nodes[node];
        if (n.axis < 0) {
            for (int i = n.lo; i < n.hi; ++i) {
                const P3& p = (*pts)[idx[i]];
                const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
                const float d2 = (dx * dx + dy * dy) + dz * dz;
                if (d2 < r2) out.push_back(idx[i]);
            }
            return;
        }
        const float d = coord(q, n.axis) - n.split;
        if (d <= r) radius(n.left, q, r, r2, out);
        if (d >= -r) radius(n.right, q, r, r2, out);
    }
};

struct Locator {
    int wz, hz, queue_size, min_size, max_size, threads;
    float zoom, min_diff, max_diff, max_distance, tol;
    float K[9], L[16], Kinv[9], R[9], t[3];
    double M[12];
    std::vector<float> depth, background, diff;
    std::deque<std::vector<float>> ring;
    std::vector<P3> fg;
    std::unordered_map<Pix, int, PixHash> point_index;
    std::unordered_map<int, int> index_cluster;
    std::vector<std::vector<int>> clusters;
    KdTree tree;
    std::vector<int> label_img;

    void lidar_to_camera(const P3& p, float& u, float& v, float& d) const {
        float cam[3], pix[3];
        for (int i = 0; i < 3; ++i) cam[i] = ((L[i * 4] * p.x + L[i * 4 + 1] * p.y) + L[i * 4 + 2] * p.z) + L[i * 4 + 3];
        for (int i = 0; i < 3; ++i) pix[i] = (K[i * 3] * cam[0] + K[i * 3 + 1] * cam[1]) + K[i * 3 + 2] * cam[2];
        u = (pix[0] * zoom) / pix[2]; v = (pix[1] * zoom) / pix[2]; d = pix[2];
    }
    P3 camera_to_lidar(float u, float v, float d) const {
        const float ccx = u / zoom, ccy = v / zoom;
        float in[3];
        for (int i = 0; i < 3; ++i)
            in[i] = ((((Kinv[i * 3] * d) * ccx) + ((Kinv[i * 3 + 1] * d) * ccy)) + (Kinv[i * 3 + 2] * d)) + t[i];
        P3 o;
        o.x = (R[0] * in[0] + R[1] * in[1]) + R[2] * in[2];
        o.y = (R[3] * in[0] + R[4] * in[1]) + R[5] * in[2];
        o.z = (R[6] * in[0] + R[7] * in[1]) + R[8] * in[2];
        return o;
    }

    void update(const float* pts, int n, int stride) {
        std::fill(depth.begin(), depth.end(), 0.f);
        std::fill(diff.begin(), diff.end(), 0.f);
        if (!pts || n <= 0) return;
        // projection is embarrassingly parallel; the scatter stays sequential to keep cloud order
        std::vector<int> pix(n);
        std::vector<float> dep(n);
        auto work = [&](int lo, int hi) {
            for (int i = lo; i < hi; ++i) {
                const P3 p{pts[(size_t)i * stride], pts[(size_t)i * stride + 1], pts[(size_t)i * stride + 2]};
                pix[i] = -1;
                if (p.x == 0.f && p.y == 0.f && p.z == 0.f) continue;
                if (p.x > max_distance) continue;
                float u, v, d;
                lidar_to_camera(p, u, v, d);
                if (!(u >= 0.f && u < (float)wz && v >= 0.f && v < (float)hz)) continue;
                pix[i] = (int)v * wz + (int)u;
                dep[i] = d;
            }
        };
        run_parallel(n, work);
        for (int i = 0; i < n; ++i) {
            if (pix[i] < 0) continue;
            if (dep[i] > background[pix[i]]) background[pix[i]] = dep[i];
            depth[pix[i]] = dep[i];
        }
        ring.push_back(depth);
        if ((int)ring.size() > queue_size) ring.pop_front();
        for (const auto& img : ring) {
            auto pass = [&](int lo, int hi) {
                for (int j = lo; j < hi; ++j) {
                    const float value = img[j];
                    if (value == 0.f) continue;
                    const float df = background[j] - value;
                    if (df >= min_diff && df <= max_diff) diff[j] = value;
                }
            };
            run_parallel(wz * hz, pass);
        }
    }

    template <class F>
    void run_parallel(int n, F&& f) {
        const int nt = std::max(1, std::min(threads, n / 4096 + 1));
        if (nt == 1) { f(0, n); return; }
        std::vector<std::thread> th;
        const int chunk = (n + nt - 1) / nt;
        for (int t0 = 0; t0 < nt; ++t0) th.emplace_back(f, std::min(n, t0 * chunk), std::min(n, (t0 + 1) * chunk));
        for (auto& x : th) x.join();
    }

    void cluster() {
        point_index.clear(); index_cluster.clear(); clusters.clear(); fg.clear();
        std::fill(label_img.begin(), label_img.end(), -2);
        for (int i = 0; i < hz; ++i)
            for (int j = 0; j < wz; ++j) {
                const float value = diff[(size_t)i * wz + j];
                if (value == 0.f) continue;
                fg.push_back(camera_to_lidar((float)j, (float)i, value));
                point_index.emplace(Pix{j, i}, (int)fg.size() - 1);
            }
        if (fg.empty()) return;
        tree.set(fg);
        const float r2 = tol * tol;
        std::vector<char> processed(fg.size(), 0);
        std::vector<int> nn;
        for (int i = 0; i < (int)fg.size(); ++i) {
            if (processed[i]) continue;
            std::vector<int> q{i};
            processed[i] = 1;
            for (size_t s = 0; s < q.size(); ++s) {
                nn.clear();
                tree.radius(0, fg[q[s]], tol, r2, nn);
                for (int j : nn) if (!processed[j]) { processed[j] = 1; q.push_back(j); }
            }
            if ((int)q.size() >= min_size && (int)q.size() <= max_size) {
                std::sort(q.begin(), q.end());
                clusters.push_back(std::move(q));
            }
        }
        // size descending, ties by smallest member index (clusters are discovered in that order)
        std::stable_sort(clusters.begin(), clusters.end(),
                         [](const std::vector<int>& a, const std::vector<int>& b) { return a.size() > b.size(); });
        for (size_t c = 0; c < clusters.size(); ++c)
            for (int index : clusters[c]) index_cluster.emplace(index, (int)c);
        for (const auto& kv : point_index) {
            auto it = index_cluster.find(kv.second);
            label_img[(size_t)kv.first.y * wz + kv.first.x] = it == index_cluster.end() ? -1 : it->second;
        }
    }

    // rect: rounded integer rect (Robot::rect()); out: xyz metres; returns 1 if located
    int search(const float* rect_f, float* out, int* info) const {
        const int rx = (int)std::lrintf(rect_f[0]), ry = (int)std::lrintf(rect_f[1]);
        const int rw = (int)std::lrintf(rect_f[2]), rh = (int)std::lrintf(rect_f[3]);
        const float cx = rx * zoom + rw * zoom * 0.5f, cy = ry * zoom + rh * zoom * 0.5f;
        const int zw = (int)(rw * zoom), zh = (int)(rh * zoom);
        const int zx = (int)(cx - zw * 0.5f), zy = (int)(cy - zh * 0.5f);
        const int x1 = std::max(zx, 0), y1 = std::max(zy, 0), x2 = std::min(zx + zw, wz), y2 = std::min(zy + zh, hz);
        std::map<int, std::vector<P3>> cand;
        for (int v = y1; v < y2; ++v)
            for (int u = x1; u < x2; ++u) {
                const float d = diff[(size_t)v * wz + u];
                if (d == 0.f) continue;
                const int index = point_index.at(Pix{u, v});
                auto it = index_cluster.find(index);
                cand[it == index_cluster.end() ? -1 : it->second].push_back(camera_to_lidar((float)u, (float)v, d));
            }
        if (cand.empty()) return 0;
        auto best = cand.begin();
        for (auto it = cand.begin(); it != cand.end(); ++it) if (best->second.size() < it->second.size()) best = it;
        double sx = 0, sy = 0, sz = 0;
        for (const P3& p : best->second) { sx += p.x; sy += p.y; sz += p.z; }
        const double n = (double)best->second.size();
        sx /= n; sy /= n; sz /= n;
        for (int i = 0; i < 3; ++i) out[i] = (float)((M[i * 4] * sx + M[i * 4 + 1] * sy + M[i * 4 + 2] * sz + M[i * 4 + 3]) * 1e-3);
        if (info) { info[0] = best->first; info[1] = (int)best->second.size(); }
        return 1;
    }
};

bool invert(const double* a, double* out, int n) {
    double m[4][8];
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { m[i][j] = a[i * n + j]; m[i][n + j] = i == j; }
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r) if (std::fabs(m[r][c]) > std::fabs(m[p][c])) p = r;
        if (std::fabs(m[p][c]) < 1e-300) return false;
        if (p != c) for (int j = 0; j < 2 * n; ++j) std::swap(m[p][j], m[c][j]);
        const double d = m[c][c];
        for (int j = 0; j < 2 * n; ++j) m[c][j] /= d;
        for (int r = 0; r < n; ++r) {
            if (r == c || m[r][c] == 0.0) continue;
            const double f = m[r][c];
            for (int j = 0; j < 2 * n; ++j) m[r][j] -= f * m[c][j];
        }
    }
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) out[i * n + j] = m[i][n + j];
    return true;
}

}  // namespace

extern "C" {

void* locref_create(int image_w, int image_h, const float* K, const float* L2C, const float* W2C, float zoom,
                    int queue, float min_diff, float max_diff, float tol, int min_size, int max_size,
                    float max_distance, int threads) {
    auto* l = new Locator();
    l->zoom = zoom; l->wz = (int)(image_w * zoom); l->hz = (int)(image_h * zoom);
    l->queue_size = queue; l->min_diff = min_diff; l->max_diff = max_diff; l->tol = tol;
    l->min_size = min_size; l->max_size = max_size; l->max_distance = max_distance;
    l->threads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    std::memcpy(l->K, K, sizeof(l->K)); std::memcpy(l->L, L2C, sizeof(l->L));
    double Kd[9], Ki[9], Ld[16], Li[16], Wd[16], Wi[16];
    for (int i = 0; i < 9; ++i) Kd[i] = K[i];
    for (int i = 0; i < 16; ++i) { Ld[i] = L2C[i]; Wd[i] = W2C[i]; }
    if (!invert(Kd, Ki, 3) || !invert(Ld, Li, 4) || !invert(Wd, Wi, 4)) { delete l; return nullptr; }
    for (int i = 0; i < 9; ++i) l->Kinv[i] = (float)Ki[i];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) l->R[i * 3 + j] = (float)Li[i * 4 + j]; l->t[i] = (float)Li[i * 4 + 3]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) {
        double acc = 0; for (int k = 0; k < 4; ++k) acc += (double)(float)Wi[i * 4 + k] * Ld[k * 4 + j];
        l->M[i * 4 + j] = acc;
    }
    const size_t n = (size_t)l->wz * l->hz;
    l->depth.assign(n, 0.f); l->background.assign(n, 0.f); l->diff.assign(n, 0.f); l->label_img.assign(n, -2);
    return l;
}
void locref_destroy(void* h) { delete static_cast<Locator*>(h); }
void locref_update(void* h, const float* pts, int n, int stride_floats) { static_cast<Locator*>(h)->update(pts, n, stride_floats); }
void locref_cluster(void* h) { static_cast<Locator*>(h)->cluster(); }
int locref_search(void* h, const float* rect, float* xyz, int* info) { return static_cast<Locator*>(h)->search(rect, xyz, info); }
void locref_search_many(void* h, const float* rects, int n, float* xyz, int* located) {
    auto* l = static_cast<Locator*>(h);
    for (int i = 0; i < n; ++i) located[i] = l->search(rects + 4 * i, xyz + 3 * i, nullptr);
}
void locref_size(void* h, int* wz, int* hz) { *wz = static_cast<Locator*>(h)->wz; *hz = static_cast<Locator*>(h)->hz; }
int locref_counts(void* h, int* nfg, int* nclusters) {
    auto* l = static_cast<Locator*>(h);
    *nfg = (int)l->fg.size(); *nclusters = (int)l->clusters.size();
    return 0;
}
// which: 0 depth 1 background 2 diff (float) 3 labels (int)
void locref_read(void* h, int which, void* out) {
    auto* l = static_cast<Locator*>(h);
    const size_t n = (size_t)l->wz * l->hz;
    if (which == 0) std::memcpy(out, l->depth.data(), n * 4);
    else if (which == 1) std::memcpy(out, l->background.data(), n * 4);
    else if (which == 2) std::memcpy(out, l->diff.data(), n * 4);
    else std::memcpy(out, l->label_img.data(), n * 4);
}

}  // extern "C"
