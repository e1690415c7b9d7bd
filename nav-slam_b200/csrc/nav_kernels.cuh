// nav_kernels.cuh -- launch wrappers shared between the .cu files of libnavslam_b200.
#pragma once
#include "nav_common.cuh"

namespace nav {

// ---- stencil.cu
void launch_labels(const double *cloud, int *labels, double *curv_or_null, long long n_rows, int cols,
                   int sm_count, cudaStream_t stream);
void launch_convert(const int *dist, const double *tan_col, const double *tan_row, double *out, int rows,
                    int cols, int sm_count, cudaStream_t stream);
void launch_transform(const double *in, double *out, long long n, const PoseXf &pose, int sm_count,
                      cudaStream_t stream);

// ---- rowmap.cu : the per-row "tree" of the SLAM step
// Device-resident state of one context: the previous frame's labelled global points of every
// (sequence,row), compacted in column order, with 16-point leaf boxes and 256-point super boxes.
struct RowMap {
    double *pts;      // [n_seq*rows][cols][3]  compacted points of each row (first map_n valid)
    int *col;         // [n_seq*rows][cols]     source column of each compacted point
    int *rank;        // [n_seq*rows][cols]     #labelled columns strictly left of column c
    int *count;       // [n_seq*rows]
    double *box;      // [n_seq*rows][n_chunks][6]   lo.xyz, hi.xyz
    double *sbox;     // [n_seq*rows][n_super][6]
    int n_chunks, n_super;
};

struct MatchOut {
    int *nn_idx;       // [n_seq*rows][cols]
    double *nn_dist;   // [n_seq*rows][cols]
    nav_corr *corr_rows;  // [n_seq*rows][cols]  per-row deduped correspondences (may be null)
    int *corr_row_count;  // [n_seq*rows]
};

size_t match_smem_bytes(int cols, bool dedupe);
size_t map_smem_bytes(int cols);
int configure_row_kernels(int cols);  // opt in to large dynamic shared memory; 0 on success

void launch_map_build(const double *cloud, const int *labels, double *global_out, const RowMap &map,
                      const PoseBatch &poses, int n_seq, int rows, int cols, cudaStream_t stream);
void launch_match(const double *cloud, const int *labels, const RowMap &map, const MatchOut &out,
                  const PoseBatch &poses, int n_seq, int rows, int cols, bool dedupe, cudaStream_t stream);
void launch_gather_corr(const nav_corr *corr_rows, const int *corr_row_count, nav_corr *corr_out,
                        int *corr_total, int n_seq, int rows, int cols, cudaStream_t stream);
void launch_flatten_row(const double *row_pts, const int *row_feature, double *out, int *count, int cols,
                        cudaStream_t stream);

}  // namespace nav
