#include "locate.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <utility>

namespace rmr {

namespace {

// ---- transforms: fixed evaluation order, one IEEE rounding per operation (see oracle) ----
// lidarToCamera — locate.cpp:73-81
__device__ __forceinline__ void lidar_to_camera(const LocateCalib& c, float x, float y, float z, float& u, float& v,
                                                float& d) {
    float cam[3], pix[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        cam[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c.L[i * 4 + 0], x), __fmul_rn(c.L[i * 4 + 1], y)),
                                     __fmul_rn(c.L[i * 4 + 2], z)),
                           c.L[i * 4 + 3]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
        pix[i] = __fadd_rn(__fadd_rn(__fmul_rn(c.K[i * 3 + 0], cam[0]), __fmul_rn(c.K[i * 3 + 1], cam[1])),
                           __fmul_rn(c.K[i * 3 + 2], cam[2]));
    u = __fdiv_rn(__fmul_rn(pix[0], c.zoom), pix[2]);
    v = __fdiv_rn(__fmul_rn(pix[1], c.zoom), pix[2]);
    d = pix[2];
}

// cameraToLidar — locate.cpp:54-61: R * (Kinv * z * [u/zoom, v/zoom, 1] + t)   (Appendix B#11: literal)
__device__ __forceinline__ float3 camera_to_lidar(const LocateCalib& c, float u, float v, float d) {
    const float ccx = __fdiv_rn(u, c.zoom), ccy = __fdiv_rn(v, c.zoom);
    float inner[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a = __fmul_rn(__fmul_rn(c.Kinv[i * 3 + 0], d), ccx);
        const float b = __fmul_rn(__fmul_rn(c.Kinv[i * 3 + 1], d), ccy);
        const float e = __fmul_rn(c.Kinv[i * 3 + 2], d);
        inner[i] = __fadd_rn(__fadd_rn(__fadd_rn(a, b), e), c.t[i]);
    }
    float o[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        o[i] = __fadd_rn(__fadd_rn(__fmul_rn(c.R[i * 3 + 0], inner[0]), __fmul_rn(c.R[i * 3 + 1], inner[1])),
                         __fmul_rn(c.R[i * 3 + 2], inner[2]));
    return make_float3(o[0], o[1], o[2]);
}

// ---- update ----
// One thread per LiDAR point (locate.cpp:173-193).  The reference's par_unseq loop races on the
// depth pixel (last writer wins) and on the background max; here the winner is the point with the
// highest cloud index (= sequential semantics) via a 64-bit atomicMax on (index+1)<<32 | depth bits,
// and the background is an exact atomic max (positive floats order like their bit patterns).
__global__ void __launch_bounds__(256) project_kernel(const __grid_constant__ LocateCalib c,
                                                      const float* __restrict__ pts, int n, int stride,
                                                      unsigned long long* __restrict__ packed,
                                                      float* __restrict__ bg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = pts[static_cast<size_t>(i) * stride + 0];
    const float y = pts[static_cast<size_t>(i) * stride + 1];
    const float z = pts[static_cast<size_t>(i) * stride + 2];
    if (x == 0.f && y == 0.f && z == 0.f) return;
    if (x > c.max_distance) return;
    float u, v, d;
    lidar_to_camera(c, x, y, z, u, v, d);
    // Appendix B#12: `>=` (the reference's `>` would index one past the row); NaN fails every test
    if (!(u >= 0.f && u < static_cast<float>(c.wz) && v >= 0.f && v < static_cast<float>(c.hz))) return;
    const int pix = static_cast<int>(v) * c.wz + static_cast<int>(u);
    if (d > 0.f) atomicMax(reinterpret_cast<int*>(bg) + pix, __float_as_int(d));
    const unsigned long long key =
        (static_cast<unsigned long long>(static_cast<unsigned>(i) + 1u) << 32) | static_cast<unsigned>(__float_as_int(d));
    atomicMax(packed + pix, key);
}

// Per pixel: resolve the winning depth into the newest ring slot, clear the staging word, apply the
// queue oldest -> newest (locate.cpp:200-219, Appendix B#10), reset the label image, count foreground.
__global__ void __launch_bounds__(256) resolve_diff_kernel(const __grid_constant__ LocateCalib c, int npix,
                                                           unsigned long long* __restrict__ packed,
                                                           const float* __restrict__ bg, float* __restrict__ ring,
                                                           int ring_head, int ring_count, int queue_size,
                                                           float* __restrict__ diff, int* __restrict__ label_img,
                                                           int* __restrict__ block_counts) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    bool fg = false;
    if (pix < npix) {
        const unsigned long long key = packed[pix];
        const float depth = key ? __int_as_float(static_cast<int>(key & 0xffffffffull)) : 0.f;
        if (key) packed[pix] = 0ull;
        ring[static_cast<size_t>(ring_head) * npix + pix] = depth;
        const float b = bg[pix];
        float out = 0.f;
        // ring_head is the newest; oldest is ring_head - (ring_count-1)
        for (int q = ring_count - 1; q >= 0; --q) {
            int slot = ring_head - q;
            if (slot < 0) slot += queue_size;
            const float value = (q == 0) ? depth : ring[static_cast<size_t>(slot) * npix + pix];
            if (value == 0.f) continue;
            const float df = __fsub_rn(b, value);
            if (df >= c.min_diff && df <= c.max_diff) out = value;
        }
        diff[pix] = out;
        label_img[pix] = -2;
        fg = out != 0.f;
    }
    const int cnt = __syncthreads_count(fg);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(1024) scan_blocks_kernel(const int* __restrict__ counts, int* __restrict__ offsets,
                                                           int nblocks, int* __restrict__ counters, int max_fg) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < nblocks ? counts[i] : 0;
        int incl = v;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const int prefix = carry + (warp ? warp_sums[warp - 1] : 0) + incl - v;
        if (i < nblocks) offsets[i] = prefix;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        counters[0] = min(carry, max_fg);
        counters[1] = 0;
        counters[2] = 0;
        counters[3] = carry;   // untruncated foreground count (overflow diagnostics)
    }
}

// ---- radius graph on a fine cell grid ----
// Cell edge = kCellFrac * tolerance with kCellFrac < 1/sqrt(3): any two points of one cell are closer
// than the tolerance (diagonal = 0.987 tol, a margin far above float rounding), so a whole cell is one
// union without a single distance test, and a point has neighbours only in the 5x5x5 cells around it.
constexpr float kCellFrac = 0.57f;
constexpr unsigned long long kEmptyKey = ~0ull;

__device__ __forceinline__ unsigned long long cell_key(int ix, int iy, int iz) {
    return (static_cast<unsigned long long>(static_cast<unsigned>(ix + (1 << 20)) & 0x1FFFFFu) << 42) |
           (static_cast<unsigned long long>(static_cast<unsigned>(iy + (1 << 20)) & 0x1FFFFFu) << 21) |
           static_cast<unsigned long long>(static_cast<unsigned>(iz + (1 << 20)) & 0x1FFFFFu);
}
__device__ __forceinline__ unsigned key_hash(unsigned long long k, unsigned mask) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33;
    return static_cast<unsigned>(k) & mask;
}
__device__ __forceinline__ int3 cell_of(const LocateCalib& c, float x, float y, float z) {
    const float inv = 1.f / (kCellFrac * c.tol);
    return make_int3(static_cast<int>(floorf(x * inv)), static_cast<int>(floorf(y * inv)),
                     static_cast<int>(floorf(z * inv)));
}

// Row-major stable compaction of foreground pixels + back-projection (locate.cpp:237-250): the point
// index equals the reference's cloud_foreground_ index.  Also initialises union-find and threads the
// point onto its cell's list (exact cell match: open addressing on the packed cell coordinates).
__global__ void __launch_bounds__(256) compact_kernel(const __grid_constant__ LocateCalib c, int npix,
                                                      const float* __restrict__ diff,
                                                      const int* __restrict__ block_offsets, int max_fg,
                                                      float* __restrict__ fg_pts, int* __restrict__ parent,
                                                      int* __restrict__ next, int* __restrict__ heads,
                                                      unsigned long long* __restrict__ keys, unsigned hash_mask,
                                                      int* __restrict__ comp_size) {
    __shared__ int warp_tot[8];
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    const float depth = pix < npix ? diff[pix] : 0.f;
    const bool fg = depth != 0.f;
    const unsigned ballot = __ballot_sync(0xffffffffu, fg);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_tot[warp] = __popc(ballot);
    __syncthreads();
    if (!fg) return;
    int idx = block_offsets[blockIdx.x] + __popc(ballot & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) idx += warp_tot[w];
    if (idx >= max_fg) return;
    const int u = pix % c.wz, v = pix / c.wz;
    const float3 p = camera_to_lidar(c, static_cast<float>(u), static_cast<float>(v), depth);
    reinterpret_cast<float4*>(fg_pts)[idx] = make_float4(p.x, p.y, p.z, __int_as_float(pix));
    parent[idx] = idx;
    comp_size[idx] = 0;
    const int3 ci = cell_of(c, p.x, p.y, p.z);
    const unsigned long long key = cell_key(ci.x, ci.y, ci.z);
    unsigned h = key_hash(key, hash_mask);
    while (true) {
        const unsigned long long prev = atomicCAS(keys + h, kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) break;
        h = (h + 1) & hash_mask;
    }
    next[idx] = atomicExch(heads + h, idx);
}

// Reads bypass L1 (other SMs link concurrently).  A stale read can only return an *older* ancestor,
// which uf_union's atomicMin detects (old != a) and retries from.
__device__ __forceinline__ int uf_find(const int* parent, int x) {
    int p = __ldcg(parent + x);
    while (p != x) {
        x = p;
        p = __ldcg(parent + x);
    }
    return x;
}

__global__ void fill_int_kernel(int* p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }   // a > b: hang the larger root under the smaller
        const int old = atomicMin(parent + a, b);
        if (old == a) return;
        a = old;   // a was no longer a root: retry from its new parent
    }
}

// One thread per (point, neighbour cell).  Slot 0 chains the point to the next one on its own cell's
// list (same cell => within tolerance).  Slots 1..62 cover the half space of the 124 surrounding
// cells, so that each unordered pair of cells is examined from exactly one side: the thread walks the
// neighbour cell's list up to the first point within the tolerance and unions with it — the rest of
// that cell is already one component.  Every union is backed by a genuine edge of the radius graph and
// every edge ends up inside one component, so the partition equals the exact connected components
// (PCL EuclideanClusterExtraction, locate.cpp:255-257); roots are minimal member indices.
constexpr int kLinkSlots = 63;
__global__ void __launch_bounds__(256) link_kernel(const __grid_constant__ LocateCalib c,
                                                   const int* __restrict__ counters,
                                                   const float* __restrict__ fg_pts, int* __restrict__ parent,
                                                   const int* __restrict__ next, const int* __restrict__ heads,
                                                   const unsigned long long* __restrict__ keys, unsigned hash_mask) {
    const int n = counters[0];
    const long total = static_cast<long>(n) * kLinkSlots;
    const float tol2 = __fmul_rn(c.tol, c.tol);
    // grid-stride: the grid is sized for the machine, not for max_foreground
    for (long t = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<long>(gridDim.x) * blockDim.x) {
        const int i = static_cast<int>(t / kLinkSlots);
        const int slot = static_cast<int>(t - static_cast<long>(i) * kLinkSlots);
        if (slot == 0) {
            const int j = next[i];
            if (j >= 0) uf_union(parent, i, j);
            continue;
        }
        // half-space enumeration of (dx, dy, dz) in [-2, 2]^3 \ {0}: slot s -> linear index 62 + s
        const int lin = 62 + slot;
        const int dz = lin / 25 - 2, dy = (lin / 5) % 5 - 2, dx = lin % 5 - 2;
        const float4 p = reinterpret_cast<const float4*>(fg_pts)[i];
        const int3 ci = cell_of(c, p.x, p.y, p.z);
        const unsigned long long key = cell_key(ci.x + dx, ci.y + dy, ci.z + dz);
        unsigned h = key_hash(key, hash_mask);
        bool found = false;
        while (true) {
            const unsigned long long k = keys[h];
            if (k == key) { found = true; break; }
            if (k == kEmptyKey) break;
            h = (h + 1) & hash_mask;
        }
        if (!found) continue;
        for (int j = heads[h]; j >= 0; j = next[j]) {
            const float4 q = reinterpret_cast<const float4*>(fg_pts)[j];
            const float ex = __fsub_rn(p.x, q.x), ey = __fsub_rn(p.y, q.y), ez = __fsub_rn(p.z, q.z);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
            if (d2 < tol2) {
                uf_union(parent, i, j);
                break;
            }
        }
    }
}

__global__ void __launch_bounds__(256) flatten_kernel(const int* __restrict__ counters, int* __restrict__ parent,
                                                      int* __restrict__ comp_size) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counters[0]) return;
    const int r = uf_find(parent, i);
    parent[i] = r;
    atomicAdd(comp_size + r, 1);
}

// PCL keeps a component iff min <= size <= max (oversize ones are dropped whole)
__global__ void __launch_bounds__(256) collect_roots_kernel(const __grid_constant__ LocateCalib c,
                                                            int* __restrict__ counters,
                                                            const int* __restrict__ parent,
                                                            const int* __restrict__ comp_size,
                                                            int* __restrict__ root_list) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counters[0]) return;
    if (parent[i] != i) return;
    const int s = comp_size[i];
    if (s < c.min_size || s > c.max_size) return;
    const int slot = atomicAdd(counters + 2, 1);
    if (slot < kMaxClusters) root_list[slot] = i;
}

// cluster id = rank by (size descending, smallest member index ascending)
__global__ void __launch_bounds__(256) rank_roots_kernel(int* __restrict__ counters,
                                                         const int* __restrict__ comp_size,
                                                         const int* __restrict__ root_list,
                                                         int* __restrict__ cluster_id) {
    const int nr = min(counters[2], kMaxClusters);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) counters[1] = nr;
    if (k >= nr) return;
    const int r = root_list[k];
    const int s = comp_size[r];
    int rank = 0;
    for (int j = 0; j < nr; ++j) {
        const int r2 = root_list[j];
        const int s2 = comp_size[r2];
        rank += (s2 > s || (s2 == s && r2 < r)) ? 1 : 0;
    }
    cluster_id[r] = rank;
}

__global__ void __launch_bounds__(256) label_kernel(const __grid_constant__ LocateCalib c,
                                                    const int* __restrict__ counters,
                                                    const float* __restrict__ fg_pts, const int* __restrict__ parent,
                                                    const int* __restrict__ comp_size,
                                                    const int* __restrict__ cluster_id, int* __restrict__ label_img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counters[0]) return;
    const int r = parent[i];
    const int s = comp_size[r];
    int id = -1;
    if (s >= c.min_size && s <= c.max_size) id = cluster_id[r];
    const int pix = __float_as_int(reinterpret_cast<const float4*>(fg_pts)[i].w);
    label_img[pix] = id;
}

// ---- search: one block per robot (locate.cpp:276-311, zoom :337-350) ----
__device__ __forceinline__ int cv_round(float v) { return __float2int_rn(v); }   // cvRound: half to even

__global__ void __launch_bounds__(256) search_kernel(const __grid_constant__ LocateCalib c,
                                                     const RectF* __restrict__ rects, LocResult* __restrict__ results,
                                                     const float* __restrict__ diff, const int* __restrict__ label_img,
                                                     const int* __restrict__ counters, int* __restrict__ hist_all) {
    const int rb = blockIdx.x;
    const RectF rf = rects[rb];
    LocResult res{0.f, 0.f, 0.f, 0, -2, 0};
    __shared__ int s_best_cnt[256];
    __shared__ int s_best_id[256];
    __shared__ double s_sum[3][256];
    if (!rf.valid) {
        if (threadIdx.x == 0) results[rb] = res;
        return;
    }
    // Robot::rect(): Rect2f -> Rect by cvRound (robot.h:111)
    const int rx = cv_round(rf.x), ry = cv_round(rf.y), rw = cv_round(rf.w), rh = cv_round(rf.h);
    const float z = c.zoom;
    const float cxf = __fadd_rn(__fmul_rn(static_cast<float>(rx), z), __fmul_rn(__fmul_rn(static_cast<float>(rw), z), 0.5f));
    const float cyf = __fadd_rn(__fmul_rn(static_cast<float>(ry), z), __fmul_rn(__fmul_rn(static_cast<float>(rh), z), 0.5f));
    const int zw = static_cast<int>(__fmul_rn(static_cast<float>(rw), z));
    const int zh = static_cast<int>(__fmul_rn(static_cast<float>(rh), z));
    const int zx = static_cast<int>(__fsub_rn(cxf, __fmul_rn(static_cast<float>(zw), 0.5f)));
    const int zy = static_cast<int>(__fsub_rn(cyf, __fmul_rn(static_cast<float>(zh), 0.5f)));
    const int x1 = max(zx, 0), y1 = max(zy, 0), x2 = min(zx + zw, c.wz), y2 = min(zy + zh, c.hz);
    const int w = x2 - x1, h = y2 - y1;
    if (w <= 0 || h <= 0) {
        if (threadIdx.x == 0) results[rb] = res;
        return;
    }
    const int nclusters = counters[1];
    int* hist = hist_all + static_cast<size_t>(rb) * (kMaxClusters + 1);
    const int area = w * h;
    // pass 1: group sizes per cluster id (-1 -> slot 0)
    for (int k = threadIdx.x; k < area; k += blockDim.x) {
        const int pix = (y1 + k / w) * c.wz + x1 + k % w;
        const int lbl = label_img[pix];
        if (lbl != -2) atomicAdd(hist + lbl + 1, 1);
    }
    __syncthreads();
    // pass 2: first maximum in ascending id order (std::map iteration + max_element)
    int bc = 0, bi = 0x7fffffff;
    for (int k = threadIdx.x; k <= nclusters; k += blockDim.x) {
        const int cnt = hist[k];
        if (cnt > bc || (cnt == bc && cnt > 0 && k < bi)) { bc = cnt; bi = k; }
    }
    s_best_cnt[threadIdx.x] = bc;
    s_best_id[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const int oc = s_best_cnt[threadIdx.x + o], oi = s_best_id[threadIdx.x + o];
            if (oc > s_best_cnt[threadIdx.x] || (oc == s_best_cnt[threadIdx.x] && oi < s_best_id[threadIdx.x])) {
                s_best_cnt[threadIdx.x] = oc;
                s_best_id[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    const int best_cnt = s_best_cnt[0];
    const int best = s_best_id[0] - 1;
    if (best_cnt == 0) {
        if (threadIdx.x == 0) results[rb] = res;
        return;
    }
    // pass 3: centroid of that group (fixed-shape tree reduction: deterministic)
    double sx = 0, sy = 0, sz = 0;
    for (int k = threadIdx.x; k < area; k += blockDim.x) {
        const int u = x1 + k % w, v = y1 + k / w;
        const int pix = v * c.wz + u;
        if (label_img[pix] == best) {
            const float3 p = camera_to_lidar(c, static_cast<float>(u), static_cast<float>(v), diff[pix]);
            sx += p.x; sy += p.y; sz += p.z;
        }
    }
    s_sum[0][threadIdx.x] = sx; s_sum[1][threadIdx.x] = sy; s_sum[2][threadIdx.x] = sz;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int a = 0; a < 3; ++a) s_sum[a][threadIdx.x] += s_sum[a][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double mx = s_sum[0][0] / best_cnt, my = s_sum[1][0] / best_cnt, mz = s_sum[2][0] / best_cnt;
        // lidarToWorld (locate.cpp:37-42) then Robot::setLocation mm -> m (robot.h:93-95)
        res.x = static_cast<float>((c.M[0] * mx + c.M[1] * my + c.M[2] * mz + c.M[3]) * 1e-3);
        res.y = static_cast<float>((c.M[4] * mx + c.M[5] * my + c.M[6] * mz + c.M[7]) * 1e-3);
        res.z = static_cast<float>((c.M[8] * mx + c.M[9] * my + c.M[10] * mz + c.M[11]) * 1e-3);
        res.located = 1;
        res.cluster = best;
        res.npoints = best_cnt;
        results[rb] = res;
    }
}

// ---- small dense inverses in double (cv::Matx::inv stand-in, locate.cpp:132-136) ----
bool invert(const double* a, double* out, int n) {
    double m[4][8];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            m[i][j] = a[i * n + j];
            m[i][n + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int col = 0; col < n; ++col) {
        int piv = col;
        for (int r = col + 1; r < n; ++r)
            if (std::fabs(m[r][col]) > std::fabs(m[piv][col])) piv = r;
        if (std::fabs(m[piv][col]) < 1e-300) return false;
        if (piv != col)
            for (int j = 0; j < 2 * n; ++j) std::swap(m[piv][j], m[col][j]);
        const double d = m[col][col];
        for (int j = 0; j < 2 * n; ++j) m[col][j] /= d;
        for (int r = 0; r < n; ++r) {
            if (r == col) continue;
            const double f = m[r][col];
            if (f == 0.0) continue;
            for (int j = 0; j < 2 * n; ++j) m[r][j] -= f * m[col][j];
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) out[i * n + j] = m[i][n + j];
    return true;
}

}  // namespace

Locator::Locator(const LocatorConfig& cfg, int max_points, int max_foreground, int max_robots)
    : max_points_(max_points), max_fg_(max_foreground), max_robots_(max_robots) {
    LocateCalib& c = calib_;
    c.zoom = cfg.zoom_factor;
    c.wz = static_cast<int>(cfg.image_width * cfg.zoom_factor);    // locate.cpp:121-122
    c.hz = static_cast<int>(cfg.image_height * cfg.zoom_factor);
    if (c.wz <= 0 || c.hz <= 0) throw std::invalid_argument("Locator: empty zoomed image");
    c.min_diff = cfg.min_depth_diff; c.max_diff = cfg.max_depth_diff;
    c.max_distance = cfg.max_distance; c.tol = cfg.cluster_tolerance;
    c.min_size = cfg.min_cluster_size; c.max_size = cfg.max_cluster_size;
    queue_size_ = std::max(cfg.queue_size, 1);
    std::memcpy(c.K, cfg.intrinsic, sizeof(c.K));
    std::memcpy(c.L, cfg.lidar_to_camera, sizeof(c.L));   // rows 0..2
    double Kd[9], Ki[9], Ld[16], Li[16], Wd[16], Wi[16];
    for (int i = 0; i < 9; ++i) Kd[i] = cfg.intrinsic[i];
    for (int i = 0; i < 16; ++i) { Ld[i] = cfg.lidar_to_camera[i]; Wd[i] = cfg.world_to_camera[i]; }
    if (!invert(Kd, Ki, 3) || !invert(Ld, Li, 4) || !invert(Wd, Wi, 4))
        throw std::invalid_argument("Locator: singular calibration matrix");
    for (int i = 0; i < 9; ++i) c.Kinv[i] = static_cast<float>(Ki[i]);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) c.R[i * 3 + j] = static_cast<float>(Li[i * 4 + j]);
        c.t[i] = static_cast<float>(Li[i * 4 + 3]);
    }
    // M = float(W2C^-1) * L2C in double
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
            double acc = 0;
            for (int k = 0; k < 4; ++k) acc += static_cast<double>(static_cast<float>(Wi[i * 4 + k])) * Ld[k * 4 + j];
            c.M[i * 4 + j] = acc;
        }
    npix_ = c.wz * c.hz;
    nblocks_ = (npix_ + 255) / 256;
    hash_size_ = 1;
    while (hash_size_ < 2 * max_fg_) hash_size_ <<= 1;

    RMR_CUDA(cudaMalloc(&cloud_, sizeof(float) * 4 * max_points_));
    RMR_CUDA(cudaMallocHost(&pinned_cloud_, sizeof(float) * 4 * max_points_));
    RMR_CUDA(cudaEventCreateWithFlags(&cloud_uploaded_, cudaEventDisableTiming));
    RMR_CUDA(cudaMalloc(&packed_, sizeof(unsigned long long) * npix_));
    RMR_CUDA(cudaMalloc(&bg_, sizeof(float) * npix_));
    RMR_CUDA(cudaMalloc(&diff_, sizeof(float) * npix_));
    RMR_CUDA(cudaMalloc(&ring_, sizeof(float) * npix_ * queue_size_));
    RMR_CUDA(cudaMalloc(&label_img_, sizeof(int) * npix_));
    RMR_CUDA(cudaMalloc(&block_counts_, sizeof(int) * nblocks_));
    RMR_CUDA(cudaMalloc(&block_offsets_, sizeof(int) * nblocks_));
    RMR_CUDA(cudaMalloc(&counters_, sizeof(int) * 8));
    RMR_CUDA(cudaMalloc(&fg_pts_, sizeof(float) * 4 * max_fg_));
    RMR_CUDA(cudaMalloc(&parent_, sizeof(int) * max_fg_));
    RMR_CUDA(cudaMalloc(&next_, sizeof(int) * max_fg_));
    RMR_CUDA(cudaMalloc(&heads_, sizeof(int) * hash_size_));
    RMR_CUDA(cudaMalloc(&cell_keys_, sizeof(unsigned long long) * hash_size_));
    RMR_CUDA(cudaMalloc(&comp_size_, sizeof(int) * max_fg_));
    RMR_CUDA(cudaMalloc(&cluster_id_, sizeof(int) * max_fg_));
    RMR_CUDA(cudaMalloc(&root_list_, sizeof(int) * (kMaxClusters + 1)));
    RMR_CUDA(cudaMalloc(&hist_, sizeof(int) * static_cast<size_t>(max_robots_) * (kMaxClusters + 1)));
    RMR_CUDA(cudaMallocHost(&pinned_counters_, sizeof(int) * 4));
    RMR_CUDA(cudaMalloc(&dev_rects_, sizeof(RectF) * max_robots_));
    RMR_CUDA(cudaMalloc(&dev_results_, sizeof(LocResult) * max_robots_));
    RMR_CUDA(cudaMallocHost(&pinned_rects_, sizeof(RectF) * max_robots_));
    RMR_CUDA(cudaMallocHost(&pinned_results_, sizeof(LocResult) * max_robots_));
    reset();
}

Locator::~Locator() {
    cudaFree(cloud_); cudaFreeHost(pinned_cloud_);
    if (cloud_uploaded_) cudaEventDestroy(cloud_uploaded_); cudaFree(packed_); cudaFree(bg_); cudaFree(diff_);
    cudaFree(ring_); cudaFree(label_img_); cudaFree(block_counts_); cudaFree(block_offsets_); cudaFree(counters_);
    cudaFree(fg_pts_); cudaFree(parent_); cudaFree(next_); cudaFree(heads_); cudaFree(cell_keys_); cudaFree(comp_size_);
    cudaFree(cluster_id_); cudaFree(root_list_); cudaFree(hist_); cudaFree(dev_rects_); cudaFree(dev_results_);
    cudaFreeHost(pinned_rects_); cudaFreeHost(pinned_results_); cudaFreeHost(pinned_counters_);
}

// Appendix B#13: the reference never initialises its images and relies on fresh zero pages
void Locator::reset() {
    RMR_CUDA(cudaMemset(packed_, 0, sizeof(unsigned long long) * npix_));
    RMR_CUDA(cudaMemset(bg_, 0, sizeof(float) * npix_));
    RMR_CUDA(cudaMemset(diff_, 0, sizeof(float) * npix_));
    RMR_CUDA(cudaMemset(ring_, 0, sizeof(float) * npix_ * queue_size_));
    fill_int_kernel<<<(npix_ + 255) / 256, 256>>>(label_img_, npix_, -2);
    RMR_CUDA(cudaDeviceSynchronize());
    RMR_CUDA(cudaMemset(counters_, 0, sizeof(int) * 8));
    RMR_CUDA(cudaMemset(block_counts_, 0, sizeof(int) * nblocks_));
    ring_head_ = 0;
    ring_count_ = 0;
}

const float* Locator::depth_image() const { return ring_ + static_cast<size_t>(ring_head_) * npix_; }

void Locator::update_host(const float* points, int n, int stride_floats, cudaStream_t s) {
    if (points == nullptr || n <= 0) {
        update_device(nullptr, 0, stride_floats, s);
        return;
    }
    if (n > max_points_) throw std::invalid_argument("Locator::update: cloud larger than max_points");
    // pack xyz to 3 floats while staging through pinned memory (PointXYZ has a padding float); the
    // previous upload must have left the staging buffer before it is rewritten
    RMR_CUDA(cudaEventSynchronize(cloud_uploaded_));
    if (stride_floats == 3) {
        std::memcpy(pinned_cloud_, points, sizeof(float) * 3 * n);
    } else {
        for (int i = 0; i < n; ++i) {
            pinned_cloud_[3 * i + 0] = points[static_cast<size_t>(i) * stride_floats + 0];
            pinned_cloud_[3 * i + 1] = points[static_cast<size_t>(i) * stride_floats + 1];
            pinned_cloud_[3 * i + 2] = points[static_cast<size_t>(i) * stride_floats + 2];
        }
    }
    RMR_CUDA(cudaMemcpyAsync(cloud_, pinned_cloud_, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, s));
    RMR_CUDA(cudaEventRecord(cloud_uploaded_, s));
    update_device(cloud_, n, 3, s);
}

void Locator::update_device(const float* dev_points, int n, int stride_floats, cudaStream_t s) {
    if (dev_points == nullptr || n <= 0) {
        // locate.cpp:160-171: both images are cleared, then the call returns (nothing is queued)
        RMR_CUDA(cudaMemsetAsync(diff_, 0, sizeof(float) * npix_, s));
        RMR_CUDA(cudaMemsetAsync(block_counts_, 0, sizeof(int) * nblocks_, s));
        fill_int_kernel<<<(npix_ + 255) / 256, 256, 0, s>>>(label_img_, npix_, -2);
        return;
    }
    project_kernel<<<(n + 255) / 256, 256, 0, s>>>(calib_, dev_points, n, stride_floats, packed_, bg_);
    // push_back + pop_front (locate.cpp:195-198)
    if (ring_count_ == 0) ring_head_ = 0;
    else ring_head_ = (ring_head_ + 1) % queue_size_;
    ring_count_ = std::min(ring_count_ + 1, queue_size_);
    resolve_diff_kernel<<<nblocks_, 256, 0, s>>>(calib_, npix_, packed_, bg_, ring_, ring_head_, ring_count_,
                                                 queue_size_, diff_, label_img_, block_counts_);
    RMR_CUDA(cudaGetLastError());
}

void Locator::cluster(cudaStream_t s) {
    scan_blocks_kernel<<<1, 1024, 0, s>>>(block_counts_, block_offsets_, nblocks_, counters_, max_fg_);
    RMR_CUDA(cudaMemsetAsync(heads_, 0xFF, sizeof(int) * hash_size_, s));
    RMR_CUDA(cudaMemsetAsync(cell_keys_, 0xFF, sizeof(unsigned long long) * hash_size_, s));
    compact_kernel<<<nblocks_, 256, 0, s>>>(calib_, npix_, diff_, block_offsets_, max_fg_, fg_pts_, parent_, next_,
                                            heads_, cell_keys_, static_cast<unsigned>(hash_size_ - 1), comp_size_);
    const int fg_blocks256 = (max_fg_ + 255) / 256;
    link_kernel<<<148 * 8, 256, 0, s>>>(
        calib_, counters_, fg_pts_, parent_, next_, heads_, cell_keys_, static_cast<unsigned>(hash_size_ - 1));
    flatten_kernel<<<fg_blocks256, 256, 0, s>>>(counters_, parent_, comp_size_);
    collect_roots_kernel<<<fg_blocks256, 256, 0, s>>>(calib_, counters_, parent_, comp_size_, root_list_);
    rank_roots_kernel<<<(kMaxClusters + 255) / 256, 256, 0, s>>>(counters_, comp_size_, root_list_, cluster_id_);
    label_kernel<<<fg_blocks256, 256, 0, s>>>(calib_, counters_, fg_pts_, parent_, comp_size_, cluster_id_,
                                              label_img_);
    RMR_CUDA(cudaGetLastError());
}

void Locator::search_device(const RectF* dev_rects, LocResult* dev_results, int n, cudaStream_t s) {
    if (n <= 0) return;
    if (n > max_robots_) throw std::invalid_argument("Locator::search: too many robots");
    RMR_CUDA(cudaMemsetAsync(hist_, 0, sizeof(int) * static_cast<size_t>(n) * (kMaxClusters + 1), s));
    search_kernel<<<n, 256, 0, s>>>(calib_, dev_rects, dev_results, diff_, label_img_, counters_, hist_);
    RMR_CUDA(cudaGetLastError());
}

void Locator::search(const RectF* rects, LocResult* results, int n, cudaStream_t s) {
    if (n <= 0) return;
    search_begin(rects, n, s);
    search_end(results, n, s);
}

// asynchronous half: rectangles up, search kernel, results + the counters of the last cluster() down
void Locator::search_begin(const RectF* rects, int n, cudaStream_t s) {
    if (n <= 0) return;
    if (n > max_robots_) throw std::invalid_argument("Locator::search: too many robots");
    std::memcpy(pinned_rects_, rects, sizeof(RectF) * n);
    RMR_CUDA(cudaMemcpyAsync(dev_rects_, pinned_rects_, sizeof(RectF) * n, cudaMemcpyHostToDevice, s));
    search_device(dev_rects_, dev_results_, n, s);
    RMR_CUDA(cudaMemcpyAsync(pinned_results_, dev_results_, sizeof(LocResult) * n, cudaMemcpyDeviceToHost, s));
    RMR_CUDA(cudaMemcpyAsync(pinned_counters_, counters_, sizeof(int) * 4, cudaMemcpyDeviceToHost, s));
}

void Locator::search_end(LocResult* results, int n, cudaStream_t s) {
    if (n <= 0) return;
    RMR_CUDA(cudaStreamSynchronize(s));
    // the reference clusters every foreground point (locate.cpp:237-263); a truncated foreground or cluster list would
    // move robots, so running out of room is an error, not a smaller answer
    if (pinned_counters_[3] > max_fg_)
        throw CapacityError(std::to_string(pinned_counters_[3]) + " foreground pixels, Locator capacity " + std::to_string(max_fg_));
    if (pinned_counters_[2] > kMaxClusters)
        throw CapacityError(std::to_string(pinned_counters_[2]) + " clusters, Locator capacity " + std::to_string(kMaxClusters));
    std::memcpy(results, pinned_results_, sizeof(LocResult) * n);
}

int Locator::fg_count_sync(cudaStream_t s) {
    int v[4];
    RMR_CUDA(cudaMemcpyAsync(v, counters_, sizeof(v), cudaMemcpyDeviceToHost, s));
    RMR_CUDA(cudaStreamSynchronize(s));
    return v[0];
}

int Locator::num_clusters_sync(cudaStream_t s) {
    int v[4];
    RMR_CUDA(cudaMemcpyAsync(v, counters_, sizeof(v), cudaMemcpyDeviceToHost, s));
    RMR_CUDA(cudaStreamSynchronize(s));
    return v[1];
}

}  // namespace rmr
