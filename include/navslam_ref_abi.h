/*
 * navslam_ref_abi.h -- the reference's own C interface for the front-end path, restated so
 * that a per-shape shim (nav-slam_b200/shim/navslam_shim.c) can export byte-identical symbols
 * on top of libnavslam_b200.so.  A maintainer of the reference does NOT need this file: their
 * main.c keeps including their own pointcloud.h / kdtree.h / slam.h and simply links against
 * libnavslam_shim_<ROWS>x<COLS>.so instead of slam.c + kdtree.c + pointcloud.c.
 *
 * Shape: the reference hard-codes MAX_ROWS / MAX_COLS (utils/pointcloud.h:9-10) and bakes them
 * into struct layouts and array-typed parameters, so one shim is compiled per shape with
 * -DMAX_ROWS=.. -DMAX_COLS=..  Layouts below must match the reference's headers exactly;
 * tests/test_shim_abi.py checks sizeof/offsetof against the reference's compiled code.
 */
#ifndef NAVSLAM_REF_ABI_H
#define NAVSLAM_REF_ABI_H

#include <stddef.h>

#ifndef MAX_ROWS
#define MAX_ROWS 8 /* utils/pointcloud.h:5,9 (L5 default) */
#endif
#ifndef MAX_COLS
#define MAX_COLS 8 /* utils/pointcloud.h:6,10 */
#endif

/* utils/pointcloud.h:13-17 */
typedef struct {
    int ToF_timestamps;
    int ToF_distances[MAX_ROWS][MAX_COLS]; /* mm */
} L5_LidarDataFrame;

/* utils/pointcloud.h:20-29 */
typedef struct {
    int IMU_timestamps;
    double roll, pitch, yaw; /* degrees */
    double x, y, z;          /* metres  */
} IMUDataFrame;

/* utils/pointcloud.h:32-35: mm, degrees */
typedef struct {
    double x, y, z, roll, pitch, yaw;
} Pos;

/* utils/pointcloud.h:39-44 */
typedef struct {
    double x, y, z;
} Point;

/* utils/pointcloud.h:47-51: points start at byte offset 8 */
typedef struct {
    int ToF_timestamps;
    Point ToF_position[MAX_ROWS][MAX_COLS];
} PointCloud;

/* utils/kdtree.h:7-11.  In the shim a KDNode* is an OPAQUE handle: only kdtree.c ever
 * dereferences nodes in the reference (main.c:262,387 just forward one to printKDTree). */
typedef struct KDNode {
    Point point;
    struct KDNode *left;
    struct KDNode *right;
} KDNode;

/* utils/kdtree.h:14-18 */
typedef struct {
    Point oriPoint;
    Point nearestPoint;
    double distance;
} NeighborResult;

/* headers/slam.h:10-18 */
typedef struct {
    PointCloud globalPointCloud[100];
    int frameCount;
    KDNode *kdtree_lastframe[MAX_ROWS];
    double error;
} SLAM_attr;

/* utils/pointcloud.h:55,57 */
void convertToPointCloud(int distances[MAX_ROWS][MAX_COLS], Point pointCloud[MAX_ROWS][MAX_COLS]);
void printPointCloud(PointCloud pointcloud);

/* utils/kdtree.h:21,24,27,30 */
KDNode *buildKDTree(Point *points, size_t numPoints, int depth);
void freeKDTree(KDNode *root);
void nearestNeighborSearch(KDNode *root, Point *target, Point *result, double *bestDist, int depth);
void printKDTree(KDNode *root, int depth);

/* headers/slam.h:22,25,28 */
void init_slam(SLAM_attr *attr, Pos pos, PointCloud *lidarPointCloud);
Pos slam_localization(SLAM_attr *attr, PointCloud *lidarPointCloud, Pos pos_predict, Pos pos_last);
void slam_mapping(SLAM_attr *attr, Pos pos, PointCloud *lidarPointCloud);

/* un-headered externals of src/slam.c:11,64,84,95,118 and utils/kdtree.c:8,14,20 that have external
 * linkage in the reference and are natural function-level entry points */
void extract_feature(PointCloud *lidarPointCloud, int feature[MAX_ROWS][MAX_COLS]);
void flattenPoints(Point rowPoints[MAX_COLS], int rowFeature[MAX_COLS], Point flattenedPoints[MAX_COLS],
                   size_t *numPoints);
void compute_posdiff(Pos *pos_now, Pos *pos_last, double pos_diff[6]);
void getRotationMatrix(double roll, double pitch, double yaw, double R[3][3]);
void mapCoordinatesToLastFrame(PointCloud *globalPointCloudData, double transform[3],
                               PointCloud *positionInLastFrame);                       /* src/slam.c:118 */
int getAxis(int depth);
double euclideanDistance(Point p1, Point p2);
void nth_element(Point *points, size_t first, size_t last, size_t nth, int axis);   /* utils/kdtree.c:20 */

/* additions of the shim (not in the reference): batched sibling of nearestNeighborSearch and
 * access to the underlying library objects */
struct nav_ctx;
struct nav_kdtree;
int navslam_nn_batch(KDNode *root, const Point *targets, size_t n, int *index_out, double *dist_out,
                     Point *nearest_out);
struct nav_ctx *navslam_context_of(SLAM_attr *attr);
/* give the device context behind a SLAM_attr back (the reference has no teardown call; anything still
 * held is released at exit) */
void navslam_release(SLAM_attr *attr);
struct nav_kdtree *navslam_tree_of(KDNode *root);

#endif
