/*
 * navslam_oracle.h -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, runtime-shaped restatement of the NAV-SLAM front end (reference:
 * wuHakureReimu/NAV-SLAM, src/slam.c, utils/kdtree.c, utils/pointcloud.c).
 * Every function cites the reference file:line whose arithmetic it follows.
 * Parity status: PINNED -- tests/test_oracle_vs_ref.py checks every function
 * here bit-for-bit against the reference's own code compiled into
 * oracle/_ref/libnavref_<RxC>.so (the reference ships no golden vectors,
 * SURVEY section 4), and tests/golden/ holds vectors generated from that library.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may include, link or execute this code.  The product
 * (nav-slam_b200/) never does and fails loudly without its CUDA library.
 */
#ifndef NAVSLAM_ORACLE_H
#define NAVSLAM_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double x, y, z; } nso_point;                       /* utils/pointcloud.h:39-44 */
typedef struct { double x, y, z, roll, pitch, yaw; } nso_pos;       /* utils/pointcloud.h:32-35 */
typedef struct { nso_point ori, nearest; double distance; } nso_corr; /* utils/kdtree.h:14-18 */

/* a2: depth matrix (mm) -> xyz, utils/pointcloud.c:8-48 */
void nso_convert_to_pointcloud(int rows, int cols, const int *dist, nso_point *out);

/* a3: curvature (0 where the reference does not evaluate) and edge labels, src/slam.c:11-61.
 * nso_extract_feature only ever writes 1s, like the reference. */
void nso_curvature(int rows, int cols, const nso_point *cloud, double *curv);
void nso_extract_feature(int rows, int cols, const nso_point *cloud, int *feature);

/* a4: stable row compaction, src/slam.c:64-72 */
size_t nso_flatten(int cols, const nso_point *row, const int *row_feature, nso_point *out);

/* a7: rotation (radians in) and per-point rigid transform, src/slam.c:95-115,145-160,118-131 */
void nso_rotation(double roll, double pitch, double yaw, double R[9]);
void nso_deg_rotation(const nso_pos *pos, double R[9]); /* DEG2RAD of src/slam.c:8 applied first */
void nso_transform(size_t n, const nso_point *in, const double R[9], const double t[3], nso_point *out);
void nso_shift(size_t n, const nso_point *in, const double d[3], nso_point *out);

/* a5/a6 in the reference's own shape: Lomuto quick-select median tree + near-first DFS,
 * utils/kdtree.c:20-82,110-152.  Permutes pts in place like the reference. */
typedef struct nso_node nso_node;
nso_node *nso_tree_build(nso_point *pts, size_t n, int depth);
void nso_tree_free(nso_node *root);
void nso_tree_nn(const nso_node *root, const nso_point *q, nso_point *best, double *best_dist, int depth);
size_t nso_tree_preorder(const nso_node *root, nso_point *out, int *depth_out, size_t cap);

/* a6 canonical form used by the CUDA path (north_star: exact NN, ties -> lowest index):
 * idx = lowest index attaining the minimum of dsq=(dx*dx+dy*dy)+dz*dz, dist = sqrt(dsq).
 * n == 0 gives idx -1 and dist +inf (the reference leaves its outputs untouched, kdtree.c:112). */
void nso_nn_brute(const nso_point *pts, size_t n, const nso_point *q, size_t nq,
                  int32_t *idx, double *dist);
/* number of points whose sqrt-distance equals the minimum (tie census for the parity tests) */
void nso_nn_tie_count(const nso_point *pts, size_t n, const nso_point *q, size_t nq, int32_t *count);

/* whole SLAM step, src/slam.c:134-431, runtime shape.  tie_mode 0 = reference tree
 * (first visited wins), 1 = canonical lowest index. */
typedef struct nso_slam nso_slam;
nso_slam *nso_slam_create(int rows, int cols, int tie_mode);
void nso_slam_destroy(nso_slam *s);
void nso_slam_init(nso_slam *s, const nso_pos *pos, const nso_point *cloud, nso_point *global_out);
/* returns the number of correspondences after the per-row dedupe (src/slam.c:247-283);
 * corr_out may be NULL; iterations_out receives how many Adam iterations ran. */
size_t nso_slam_localize(nso_slam *s, const nso_point *cloud, const nso_pos *pos_predict,
                         const nso_pos *pos_last, nso_pos *pos_out, nso_corr *corr_out,
                         size_t corr_cap, int *iterations_out);
void nso_slam_map(nso_slam *s, const nso_pos *pos, const nso_point *cloud, nso_point *global_out);
double nso_slam_error(const nso_slam *s);
int nso_slam_frame_count(const nso_slam *s);
/* the front end of one frame without dedupe/Adam (SURVEY 8d "one frame of work"):
 * features + query transform + per-row NN against the previous frame, then mapping with `pos`.
 * nn_idx (flat column of the matched map point, -1 if none / unlabelled) and nn_dist (-1 unlabelled). */
void nso_frontend_frame(nso_slam *s, const nso_point *cloud, const nso_pos *pos_predict,
                        const nso_pos *pos_last, const nso_pos *pos_final,
                        int *feature_out, int32_t *nn_idx, double *nn_dist, nso_point *global_out);

/* ---- caller-side data formats ------------------------------------------
 * nso_l9_csv_read restates L9_LidarProcessData (src/main.c:77-128) with the C library's own fscanf;
 * nso_csv_format_frame restates the fprintf loop of src/main.c:320-352 with snprintf.  Both are the
 * checkers for nav_l9_csv_read / nav_csv_format_frame. */
int nso_l9_csv_read(const char *path, int rows, int cols, size_t max_frames, nso_point *frames,
                    int32_t *timestamps, size_t *n_frames);
size_t nso_csv_format_frame(char *buf, size_t cap, unsigned long long timestamp, int rows, int cols,
                            const nso_point *global_cloud, const int32_t *distances, const double *imu6,
                            const nso_pos *lidar_pos, const nso_pos *ekf_pos);

#ifdef __cplusplus
}
#endif
#endif
