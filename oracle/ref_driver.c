/*
 * ref_driver.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Thin batch loops around the *reference's own* functions so that Python
 * (ctypes) can run and time them without one FFI call per query.  This file
 * is compiled TOGETHER WITH the reference sources where they lie under
 * /root/reference (see oracle/Makefile) into oracle/_ref/libnavref_<RxC>.so.
 * It contains no SLAM arithmetic of its own: every number it returns comes
 * out of a reference function (extract_feature src/slam.c:11, flattenPoints
 * src/slam.c:64, buildKDTree utils/kdtree.c:65, nearestNeighborSearch
 * utils/kdtree.c:110, convertToPointCloud utils/pointcloud.c:8).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the resulting library.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "slam.h" /* reference header: brings pointcloud.h + kdtree.h */

/* un-headered externals of src/slam.c (external linkage) */
void extract_feature(PointCloud *lidarPointCloud, int feature[MAX_ROWS][MAX_COLS]);
void flattenPoints(Point rowPoints[MAX_COLS], int rowFeature[MAX_COLS],
                   Point flattenedPoints[MAX_COLS], size_t *numPoints);

int refdrv_rows(void) { return MAX_ROWS; }
int refdrv_cols(void) { return MAX_COLS; }
size_t refdrv_sizeof_pointcloud(void) { return sizeof(PointCloud); }
size_t refdrv_sizeof_slam_attr(void) { return sizeof(SLAM_attr); }
size_t refdrv_sizeof_kdnode(void) { return sizeof(KDNode); }
size_t refdrv_sizeof_neighbor_result(void) { return sizeof(NeighborResult); }
size_t refdrv_offsetof_frame_count(void) { return offsetof(SLAM_attr, frameCount); }
size_t refdrv_offsetof_trees(void) { return offsetof(SLAM_attr, kdtree_lastframe); }
size_t refdrv_offsetof_error(void) { return offsetof(SLAM_attr, error); }

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* one query loop over a single tree; out_dist[i] stays INFINITY for a NULL root */
double refdrv_nn_batch(KDNode *root, const Point *q, size_t nq, Point *out_pt, double *out_dist) {
    double t0 = now_s();
    for (size_t i = 0; i < nq; ++i) {
        Point target = q[i];
        double best = INFINITY;
        nearestNeighborSearch(root, &target, &out_pt[i], &best, 0);
        out_dist[i] = best;
    }
    return now_s() - t0;
}

/* per-row flatten + build, exactly the calls src/slam.c:167-172 makes */
double refdrv_build_rows(PointCloud *global_cloud, int *feature, KDNode **trees, size_t *counts) {
    double t0 = now_s();
    for (int row = 0; row < MAX_ROWS; ++row) {
        Point flat[MAX_COLS];
        size_t n = 0;
        flattenPoints(global_cloud->ToF_position[row], feature + (size_t)row * MAX_COLS, flat, &n);
        trees[row] = buildKDTree(flat, n, 0);
        counts[row] = n;
    }
    return now_s() - t0;
}

/* per-row queries for every labelled point (call pattern of src/slam.c:236-244);
 * results are written at the flat pixel position row*MAX_COLS+col; unlabelled
 * pixels keep dist = -1. Returns seconds; *nq_out = number of queries issued. */
double refdrv_nn_rows(KDNode **trees, PointCloud *queries, int *feature,
                      Point *out_pt, double *out_dist, size_t *nq_out) {
    size_t nq = 0;
    double t0 = now_s();
    for (int row = 0; row < MAX_ROWS; ++row) {
        for (int col = 0; col < MAX_COLS; ++col) {
            size_t p = (size_t)row * MAX_COLS + col;
            if (feature[p] == 1) {
                Point target = queries->ToF_position[row][col];
                double best = INFINITY;
                nearestNeighborSearch(trees[row], &target, &out_pt[p], &best, 0);
                out_dist[p] = best;
                ++nq;
            } else {
                out_dist[p] = -1.0;
            }
        }
    }
    *nq_out = nq;
    return now_s() - t0;
}

void refdrv_free_rows(KDNode **trees) {
    for (int row = 0; row < MAX_ROWS; ++row) {
        freeKDTree(trees[row]);
        trees[row] = NULL;
    }
}

double refdrv_extract_feature_timed(PointCloud *cloud, int *feature, int reps) {
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        memset(feature, 0, sizeof(int) * MAX_ROWS * MAX_COLS);
        double t0 = now_s();
        extract_feature(cloud, (int(*)[MAX_COLS])feature);
        double dt = now_s() - t0;
        if (dt < best) best = dt;
    }
    return best;
}

/* timed large-map build (permutes pts in place, like the reference) */
double refdrv_build_timed(Point *pts, size_t n, KDNode **root_out) {
    double t0 = now_s();
    *root_out = buildKDTree(pts, n, 0);
    return now_s() - t0;
}

/* pre-order dump of a reference tree: point + depth, for structure checks */
static size_t dump_rec(KDNode *node, int depth, Point *out_pt, int *out_depth, size_t pos, size_t cap) {
    if (!node || pos >= cap) return pos;
    out_pt[pos] = node->point;
    out_depth[pos] = depth;
    pos = dump_rec(node->left, depth + 1, out_pt, out_depth, pos + 1, cap);
    return dump_rec(node->right, depth + 1, out_pt, out_depth, pos, cap);
}
size_t refdrv_tree_preorder(KDNode *root, Point *out_pt, int *out_depth, size_t cap) {
    return dump_rec(root, 0, out_pt, out_depth, 0, cap);
}

/* ---- the caller-side data formats: the reference's own L9 CSV reader --------------
 * L9_LidarProcessData (src/main.c:77-128) lives in main.c, which build_ref.sh links into this library
 * with main renamed.  Returns the seconds the call took. */
void L9_LidarProcessData(const char *filename, PointCloud *lidarData, size_t *lidarCount);

double refdrv_l9_read(const char *path, PointCloud *frames, size_t *count) {
    double t0 = now_s();
    L9_LidarProcessData(path, frames, count);
    return now_s() - t0;
}

/* the reference's own JSON readers (src/main.c:12-75,130-178), linked from main.c like the CSV reader;
 * both need jansson's json_loadf at run time (libjansson.so.4).  Counts start at zero as in main.c. */
void LidarProcessData(const char *filename, L5_LidarDataFrame *lidarData, size_t *lidarCount);
void IMUProcessData(const char *filename, IMUDataFrame *imuData, size_t *imuCount);
size_t refdrv_sizeof_l5_frame(void) { return sizeof(L5_LidarDataFrame); }
size_t refdrv_sizeof_imu_frame(void) { return sizeof(IMUDataFrame); }

size_t refdrv_l5_json_read(const char *path, L5_LidarDataFrame *frames) {
    size_t count = 0;
    LidarProcessData(path, frames, &count);
    return count;
}
size_t refdrv_imu_json_read(const char *path, IMUDataFrame *frames) {
    size_t count = 0;
    IMUProcessData(path, frames, &count);
    return count;
}
