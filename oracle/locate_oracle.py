"""ORACLE (test infrastructure only — never imported by the product path).

numpy restatement of the reference Locator (`/root/reference/src/locate/locate.cpp`), with the
sequential-order semantics SURVEY.md Appendix B fixes for the reference's racy loops:
  * depth image: last point in cloud order wins a pixel; background = exact running max (B#9)
  * diff image: queued frames applied oldest → newest (B#10)
  * cameraToLidar follows the reference formula R·(K⁻¹·z·p + t) literally (B#11)
  * u >= Wz or v >= Hz is rejected (the reference's `>` would index one past the row, B#12)

Third-party arithmetic not vendored in /root/reference: PCL `EuclideanClusterExtraction` +
`search::KdTree`/FLANN (version unpinned: `src/locate/CMakeLists.txt:2`), call sites
`locate.cpp:142-145,255-257`.  Published algorithm restated here: BFS region growing over radius
neighbours (squared float distance < tolerance²), i.e. connected components of the radius graph;
a component is kept iff min_size <= n <= max_size (oversize components are dropped whole);
clusters are ordered by size descending, ties by ascending smallest member index (PCL's final
std::sort is unstable; for <= 16 clusters it degenerates to insertion sort over the reversed
range, which gives exactly this order).  The reference's own tests at this boundary only check
`clusters_.size() == 2` and `location().has_value()` (`test/locate/locator_test.cpp:118,167`)
→ numerically "parity unpinned"; those two properties are reproduced in tests/.
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components
from scipy.spatial import cKDTree

f32 = np.float32


def rect_round(rect):
    """Robot::rect(): Rect2f → cv::Rect via cvRound (half to even). robot.h:111"""
    return tuple(int(np.rint(f32(v))) for v in rect)


class LocatorOracle:
    # Locator::Locator — locate.cpp:112-146 ; defaults locator.h:59-65
    def __init__(self, image_width, image_height, intrinsic, lidar_to_camera, world_to_camera,
                 zoom_factor=0.5, queue_size=3, min_depth_diff=500, max_depth_diff=4000,
                 cluster_tolerance=400, min_cluster_size=8, max_cluster_size=1000, max_distance=29300):
        self.zoom = f32(zoom_factor)
        self.Wz = int(f32(image_width) * self.zoom)
        self.Hz = int(f32(image_height) * self.zoom)
        self.queue_size = int(queue_size)
        self.K = np.asarray(intrinsic, f32).reshape(3, 3)
        self.L2C = np.asarray(lidar_to_camera, f32).reshape(4, 4)
        self.W2C = np.asarray(world_to_camera, f32).reshape(4, 4)
        self.Kinv = np.linalg.inv(self.K.astype(np.float64)).astype(f32)
        c2l = np.linalg.inv(self.L2C.astype(np.float64)).astype(f32)
        self.R = c2l[:3, :3].copy()
        self.t = c2l[:3, 3].copy()
        self.C2W = np.linalg.inv(self.W2C.astype(np.float64)).astype(f32)
        self.min_diff = f32(min_depth_diff)
        self.max_diff = f32(max_depth_diff)
        self.max_distance = f32(max_distance)
        self.tol = f32(cluster_tolerance)
        self.min_size = int(min_cluster_size)
        self.max_size = int(max_cluster_size)
        self.depth = np.zeros((self.Hz, self.Wz), f32)
        self.background = np.zeros((self.Hz, self.Wz), f32)   # B#13: zero-initialised
        self.diff = np.zeros((self.Hz, self.Wz), f32)
        self.ring = []
        self.fg_points = np.zeros((0, 3), f32)
        self.fg_pixels = np.zeros((0, 2), np.int32)
        self.labels = np.zeros(0, np.int32)
        self.label_image = np.full((self.Hz, self.Wz), -2, np.int32)
        self.num_clusters = 0
        self.stats = {}

    # Arithmetic order is fixed (sequential left-to-right sums, one IEEE rounding per operation,
    # no FMA contraction) so that the CUDA kernels can reproduce it bit for bit with
    # __fmul_rn/__fadd_rn/__fdiv_rn.  cv::Matx products in the reference accumulate in the same
    # k-order; whether its -Ofast build contracts them into FMAs is not knowable from source.

    # lidarToCamera — locate.cpp:73-81
    def lidar_to_camera(self, pts: np.ndarray):
        x = pts[:, 0].astype(f32); y = pts[:, 1].astype(f32); z = pts[:, 2].astype(f32)
        L, K = self.L2C, self.K
        cam = [((L[i, 0] * x + L[i, 1] * y) + L[i, 2] * z) + L[i, 3] for i in range(3)]
        pix = [(K[i, 0] * cam[0] + K[i, 1] * cam[1]) + K[i, 2] * cam[2] for i in range(3)]
        with np.errstate(divide="ignore", invalid="ignore"):
            u = (pix[0] * self.zoom) / pix[2]
            v = (pix[1] * self.zoom) / pix[2]
        return u.astype(f32), v.astype(f32), pix[2].astype(f32)

    # cameraToLidar — locate.cpp:54-61 (B#11: literal)
    def camera_to_lidar(self, u, v, d):
        u = np.asarray(u, f32); v = np.asarray(v, f32); d = np.asarray(d, f32)
        ccx = u / self.zoom
        ccy = v / self.zoom
        Ki, R, t = self.Kinv, self.R, self.t
        # (intrinsic_inv_ * point.z) is a scaled matrix, then times [ccx, ccy, 1]
        inner = [(((Ki[i, 0] * d) * ccx + (Ki[i, 1] * d) * ccy) + (Ki[i, 2] * d)) + t[i] for i in range(3)]
        out = [(R[i, 0] * inner[0] + R[i, 1] * inner[1]) + R[i, 2] * inner[2] for i in range(3)]
        return np.stack(out, axis=-1).astype(f32)

    # lidarToWorld — locate.cpp:37-42
    def lidar_to_world(self, p):
        M = (self.C2W.astype(np.float64) @ self.L2C.astype(np.float64))
        p4 = np.array([p[0], p[1], p[2], 1.0], np.float64)
        return (M @ p4)[:3]

    # Locator::update — locate.cpp:158-220
    def update(self, cloud: np.ndarray | None):
        self.depth[:] = 0
        self.diff[:] = 0
        if cloud is None or len(cloud) == 0:
            return
        pts = np.asarray(cloud, f32)[:, :3]
        keep = ~((pts[:, 0] == 0) & (pts[:, 1] == 0) & (pts[:, 2] == 0))
        keep &= ~(pts[:, 0] > self.max_distance)
        u, v, d = self.lidar_to_camera(pts)
        with np.errstate(invalid="ignore"):
            inb = ~((u < 0) | (u >= f32(self.Wz)) | (v < 0) | (v >= f32(self.Hz)))
        inb &= np.isfinite(u) & np.isfinite(v)
        keep &= inb
        idx = np.nonzero(keep)[0]
        ui = u[idx].astype(np.int32)   # Mat::at<float>(int,int): truncation
        vi = v[idx].astype(np.int32)
        di = d[idx]
        flat = vi * self.Wz + ui
        np.maximum.at(self.background.reshape(-1), flat, di)
        self.depth.reshape(-1)[flat] = di   # numpy fancy assignment: last index wins
        # (numpy guarantees the last value for repeated indices in a single assignment)
        uniq = np.unique(flat).size
        self.stats = dict(valid=int(idx.size), collisions=int(idx.size - uniq))
        self.ring.append(self.depth.copy())
        if len(self.ring) > self.queue_size:
            self.ring.pop(0)
        for img in self.ring:   # oldest → newest
            nz = img != 0
            df = self.background - img
            ok = nz & (df >= self.min_diff) & (df <= self.max_diff)
            self.diff[ok] = img[ok]

    # Locator::cluster — locate.cpp:231-264
    def cluster(self):
        vs, us = np.nonzero(self.diff)      # row-major scan order
        self.fg_pixels = np.stack([us, vs], axis=1).astype(np.int32)
        depth = self.diff[vs, us]
        self.fg_points = self.camera_to_lidar(us.astype(f32), vs.astype(f32), depth)
        n = len(us)
        self.labels = np.full(n, -1, np.int32)
        self.label_image = np.full((self.Hz, self.Wz), -2, np.int32)
        self.num_clusters = 0
        self.cluster_sizes = []
        if n == 0:
            return
        comp = radius_components(self.fg_points, self.tol)
        sizes = np.bincount(comp)
        first = np.full(sizes.size, n, np.int64)
        np.minimum.at(first, comp, np.arange(n))
        valid = [c for c in range(sizes.size) if self.min_size <= sizes[c] <= self.max_size]
        valid.sort(key=lambda c: (-int(sizes[c]), int(first[c])))
        remap = np.full(sizes.size, -1, np.int32)
        for rank, c in enumerate(valid):
            remap[c] = rank
        self.labels = remap[comp]
        self.num_clusters = len(valid)
        self.cluster_sizes = [int(sizes[c]) for c in valid]
        self.label_image[vs, us] = self.labels

    # Locator::zoom — locate.cpp:337-350
    def zoom_rect(self, rect):
        x, y, w, h = rect
        z = self.zoom
        cx = f32(f32(x) * z + f32(f32(w) * z) * f32(0.5))
        cy = f32(f32(y) * z + f32(f32(h) * z) * f32(0.5))
        rw = int(f32(w) * z)
        rh = int(f32(h) * z)
        rx = int(f32(cx - f32(rw) * f32(0.5)))
        ry = int(f32(cy - f32(rh) * f32(0.5)))
        # ret &= image_rect
        x1 = max(rx, 0); y1 = max(ry, 0)
        x2 = min(rx + rw, self.Wz); y2 = min(ry + rh, self.Hz)
        if x2 <= x1 or y2 <= y1:
            return (0, 0, 0, 0)
        return (x1, y1, x2 - x1, y2 - y1)

    # Locator::search(Robot&) — locate.cpp:276-311 ; returns (xyz metres | None, info)
    def search_rect(self, rect_f):
        rect = self.zoom_rect(rect_round(rect_f))
        x, y, w, h = rect
        sub = self.diff[y:y + h, x:x + w]
        vs, us = np.nonzero(sub)
        if len(vs) == 0:
            return None, dict(rect=rect, cluster=None, n=0)
        ids = self.label_image[y:y + h, x:x + w][vs, us]
        uniq, counts = np.unique(ids, return_counts=True)   # ascending id, -1 first
        best = uniq[np.argmax(counts)]                       # first maximum
        sel = ids == best
        pts = self.camera_to_lidar((us[sel] + x).astype(f32), (vs[sel] + y).astype(f32), sub[vs, us][sel])
        mean = pts.astype(np.float64).mean(axis=0)           # B#17: float64, tolerance 1e-3 m
        world = self.lidar_to_world(mean) * 1e-3             # Robot::setLocation robot.h:93-95
        return world, dict(rect=rect, cluster=int(best), n=int(sel.sum()))

    def search(self, rects):
        return [self.search_rect(r)[0] if r is not None else None for r in rects]


def radius_components(points: np.ndarray, tol) -> np.ndarray:
    """Connected components of the graph ‖pi−pj‖² < tol² (float32 squared distance, as FLANN's
    L2_Simple accumulates it).  Returns a component id per point."""
    n = len(points)
    p = points.astype(f32)
    tree = cKDTree(p.astype(np.float64))
    pairs = tree.query_pairs(float(tol) * (1 + 1e-6), output_type="ndarray")
    if len(pairs):
        d = p[pairs[:, 0]] - p[pairs[:, 1]]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(f32) + d[:, 2] * d[:, 2]
        pairs = pairs[d2.astype(f32) < f32(tol) * f32(tol)]
    g = coo_matrix((np.ones(len(pairs), np.int8), (pairs[:, 0], pairs[:, 1])), shape=(n, n))
    _, comp = connected_components(g, directed=False)
    return comp


def radius_components_bruteforce(points: np.ndarray, tol) -> np.ndarray:
    """O(n²) BFS statement of PCL's extractEuclideanClusters region growing; small n only."""
    n = len(points)
    p = points.astype(f32)
    comp = np.full(n, -1, np.int64)
    c = 0
    t2 = f32(tol) * f32(tol)
    for i in range(n):
        if comp[i] >= 0:
            continue
        queue = [i]
        comp[i] = c
        while queue:
            j = queue.pop()
            d = p - p[j]
            d2 = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(f32) + d[:, 2] * d[:, 2]).astype(f32)
            nb = np.nonzero((d2 < t2) & (comp < 0))[0]
            comp[nb] = c
            queue.extend(nb.tolist())
        c += 1
    return comp


def read_pcd(path: str) -> np.ndarray:
    """PCD v0.7 reader (ASCII + binary xyz float32) — stands in for pcl::io::loadPCDFile
    (`/root/reference/samples/main.cpp:42-72`); SURVEY.md Appendix C.3."""
    with open(path, "rb") as fh:
        data = fh.read()
    pos = 0
    npts = 0
    kind = None
    while True:
        end = data.index(b"\n", pos)
        line = data[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if line.startswith("POINTS"):
            npts = int(line.split()[1])
        elif line.startswith("DATA"):
            kind = line.split()[1]
            break
    if kind == "ascii":
        arr = np.array(data[pos:].split(), dtype=np.float32).reshape(-1, 3)
    else:
        arr = np.frombuffer(data, dtype="<f4", count=npts * 3, offset=pos).reshape(-1, 3).copy()
    assert arr.shape[0] == npts
    return arr
