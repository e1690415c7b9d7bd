// nav_kernels.cuh -- launch wrappers shared between the .cu files of libnavslam_b200.
#pragma once
#include "nav_common.cuh"

namespace nav {

// ---- stencil.cu
void launch_labels(const double *cloud, int *labels, double *curv_or_null, long long n_rows, int cols,
                   int sm_count, unsigned *n_exact, cudaStream_t stream);
void launch_convert(const int *dist, const double *tan_col, const double *tan_row, double *out, int rows,
                    int cols, int sm_count, cudaStream_t stream);
void launch_transform(const double *in, double *out, long long n, const PoseXf &pose, int sm_count,
                      cudaStream_t stream);

// ---- rowmap.cu : the per-row "tree" of the SLAM step
// Device-resident state of one context: the mapped (global-frame) cloud of the previous frame, left
// in image order, plus per 16-column block a label mask and a bounding box of its labelled points,
// and per 256-column block a super box.
struct RowMap {
    double *pts;      // [n_seq*rows][cols][3]   global cloud of the frame mapped last
    unsigned *mask;   // [n_seq*rows][n_chunks]  bit i = column 16*b+i is an edge point
    float4 *box;      // [n_seq*rows][n_chunks][2]   {lo.xyz,-}, {hi.xyz,-} rounded outward (+inf/-inf when empty)
    float4 *sbox;     // [n_seq*rows][n_super][2]
    int n_chunks, n_super;
};

struct MatchOut {
    int *nn_idx;          // [n_seq*rows][cols]
    double *nn_dist;      // [n_seq*rows][cols]
    nav_corr *corr_rows;  // [n_seq*rows][cols]  per-row deduped correspondences
    int *corr_row_count;  // [n_seq*rows]
};

// where k_dedupe_rows posts the fit statistics of a whole frame for the host to poll (closed loop):
// host[0..4] = {N, sum rx, sum ry, sum rz, sum |r|^2}, host[5] = seq (written last, as a 64-bit word)
struct FitMailbox {
    double *host;       // host-mapped pinned memory, 6 doubles
    unsigned *ticket;   // device counter of finished row CTAs (zero between launches)
    unsigned long long seq;
};

// a whole sequence of frames with known poses in one launch (one thread-block cluster per image row walks
// through the frames); map0 is searched by frame 0, frame f builds the map for frame f+1 in the other buffer.
// d_pose_loc / d_pose_fin: device arrays [n_frames][n_seq].  Needs frame_seq_supported(cols).
bool frame_seq_supported(int cols);
int frame_seq_inline_frames();  // sequences up to this length (one sequence) carry their poses as kernel parameters
int launch_frame_seq(const double *frames, long long frame_stride, int n_frames, int *labels, const RowMap &map0,
                     const RowMap &map1, const MatchOut &out, const PoseXf *d_pose_loc, const PoseXf *d_pose_fin,
                     int n_seq, int rows, int cols, unsigned *n_exact, cudaStream_t stream,
                     const PoseXf *h_pose_loc = nullptr, const PoseXf *h_pose_fin = nullptr);

size_t dedupe_smem_bytes(int cols);
int configure_row_kernels(int cols);  // opt in to large dynamic shared memory; 0 on success

void launch_frame_map(const double *cloud, const int *labels, const RowMap &map, const PoseBatch &poses,
                      int n_seq, int rows, int cols, cudaStream_t stream);
// map_next/final_poses non-null: the kernel also builds the next map (frame transformed with its final
// pose) into *map_next, which must be a different buffer than `map`
void launch_frame_match(const double *cloud, int *labels, bool fused_labels, const RowMap &map,
                        const MatchOut &out, const PoseBatch &poses, int n_seq, int rows, int cols,
                        unsigned *n_exact, cudaStream_t stream, const RowMap *map_next = nullptr,
                        const PoseBatch *final_poses = nullptr, bool pdl = false);
// row_stats (optional): [n_seq*rows][5] per-row {n, sum r (3), sum |r|^2} of the deduped correspondences;
// write_corr = false skips the correspondence entries themselves
void launch_dedupe(const double *cloud, const int *labels, const RowMap &map, const MatchOut &out,
                   const PoseBatch &poses, int n_seq, int rows, int cols, cudaStream_t stream,
                   double *row_stats = nullptr, bool write_corr = true, const FitMailbox *mail = nullptr);
// statistics-only dedupe spread over the tiles of a row (thread-block cluster per row); part: [n_seq*rows*tiles][5]
bool dedupe_stats_supported(int cols);
int launch_dedupe_stats(const double *cloud, const int *labels, const RowMap &map, const MatchOut &out,
                        const PoseBatch &poses, int n_seq, int rows, int cols, cudaStream_t stream, double *part,
                        const FitMailbox &mail);
// the closed-loop step in one launch: (do_map) map of the previous frame from prev_cloud / prev_labels /
// prev_poses, match of this frame against it, statistics dedupe, mailbox post.  Needs dedupe_stats_supported(cols).
int launch_loop_step(const double *cloud, const int *labels, const RowMap &map, const MatchOut &out,
                     const PoseBatch &poses, int n_seq, int rows, int cols, cudaStream_t stream,
                     const double *prev_cloud, const int *prev_labels, const PoseBatch &prev_poses, bool do_map,
                     double *part, const FitMailbox &mail);
void launch_gather_corr(const nav_corr *corr_rows, const int *corr_row_count, nav_corr *corr_out,
                        int *corr_total, int n_seq, int rows, int cols, cudaStream_t stream);
void launch_corr_stats(const nav_corr *corr, const int *corr_total, double *stats_out, int n_seq, int rows, int cols,
                       int sm_count, cudaStream_t stream);
void launch_flatten_row(const double *row_pts, const int *row_feature, const unsigned *row_mask, double *out,
                        int *col_out, int *count, int cols, cudaStream_t stream);

// ---- csvfmt.cu : CSV rows of one frame formatted on the device (src/main.c:320-352)
struct CsvJob {
    const double *cloud;  // [n][3] global cloud
    const int *dist;      // [n] or null (prints 0)
    unsigned long long ts;
    long long n;
    int cols, ts_len, tail_len;
    char tail[480];  // ",%.2f" x 18 + "\n", the same on every line (formatted on the host)
};
size_t csv_scratch_bytes(long long n);
// writes the text to d_text (needs n * (124 + tail_len) bytes); d_scratch[0..8) = total bytes,
// d_scratch[8..12) = 1 if some value needs the host formatter (inf, nan, |v| >= 2^57)
int launch_csv_format(const CsvJob &job, char *d_text, void *d_scratch, cudaStream_t stream);
// io.cu: the 18 pose columns of a line

}  // namespace nav
