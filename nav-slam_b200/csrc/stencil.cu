// stencil.cu -- range-image kernels: curvature / edge labels (a3), depth->xyz (a2), rigid
// transform (a7).  Reference arithmetic: src/slam.c:11-61, utils/pointcloud.c:8-48,
// src/slam.c:145-160.
//
// Curvature stencil layout.  An image row is an AoS run of C points (24 B each).  One CTA owns a
// tile of kTile consecutive columns of one row: it stages the tile plus a 2-point halo on each
// side in shared memory with coalesced 8-byte loads, computes for every staged point the two
// *forward* distances f1(j)=|p_j - p_{j+1}| and f2(j)=|p_j - p_{j+2}| once (the backward taps of
// the reference are the same numbers: p_j - p_{j-1} = -(p_{j-1} - p_j) exactly, and the sign is
// squared away), shares them through shared memory, and then evaluates
//     S = ((d-2 + d-1) + d+1) + d+2, avg = S/4, V = sum (d-avg)^2 in tap order,
//     curv = (V/4) / (avg*avg + (double)1e-6f), label = curv > 0.1
// exactly as src/slam.c:37-58 does.  That halves the fp64 square roots (2 per point, not 4; the
// reference evaluates 8).  Algorithmic traffic: 24 B read + 4 B written per point.
#include "nav_kernels.cuh"

namespace nav {

constexpr int kTile = 256;
constexpr int kHalo = 2;

template <bool kWriteCurv>
__global__ void __launch_bounds__(kTile)
k_labels_exact(const double *__restrict__ cloud, int *__restrict__ labels,
               double *__restrict__ curv_out, long long n_rows, int cols, int tiles_per_row) {
    __shared__ double s_pts[(kTile + 2 * kHalo) * 3];
    __shared__ double s_f1[kTile + kHalo];
    __shared__ double s_f2[kTile + kHalo];

    const long long n_tiles = n_rows * tiles_per_row;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row = tile / tiles_per_row;
        const int c0 = (int)(tile % tiles_per_row) * kTile;
        const double *row_ptr = cloud + row * (long long)cols * 3;

        // stage columns [c0-2, c0+kTile+2) clipped to the row; out-of-row slots are zero and only
        // feed border columns, which the reference never evaluates (src/slam.c:16)
        const int first = c0 - kHalo;
        for (int i = threadIdx.x; i < (kTile + 2 * kHalo) * 3; i += kTile) {
            int col = first + i / 3;
            double v = 0.0;
            if (col >= 0 && col < cols) v = __ldg(row_ptr + (long long)first * 3 + i);
            s_pts[i] = v;
        }
        __syncthreads();

        for (int i = threadIdx.x; i < kTile + kHalo; i += kTile) {
            const double *p = s_pts + i * 3;
            s_f1[i] = __dsqrt_rn(dsq3(dsub(p[0], p[3]), dsub(p[1], p[4]), dsub(p[2], p[5])));
            s_f2[i] = __dsqrt_rn(dsq3(dsub(p[0], p[6]), dsub(p[1], p[7]), dsub(p[2], p[8])));
        }
        __syncthreads();

        const int col = c0 + threadIdx.x;
        if (col < cols) {
            const int li = threadIdx.x + kHalo;
            double curv = 0.0;
            if (col >= kHalo && col < cols - kHalo) {
                const double dm2 = s_f2[li - 2], dm1 = s_f1[li - 1], dp1 = s_f1[li], dp2 = s_f2[li];
                const double sum = dadd(dadd(dadd(dm2, dm1), dp1), dp2);
                const double avg = dmul(sum, 0.25);  // sum / 4 is exact scaling
                if (avg > 0.0) {
                    double e = dsub(dm2, avg);
                    double var = dmul(e, e);
                    e = dsub(dm1, avg);
                    var = dadd(var, dmul(e, e));
                    e = dsub(dp1, avg);
                    var = dadd(var, dmul(e, e));
                    e = dsub(dp2, avg);
                    var = dadd(var, dmul(e, e));
                    curv = __ddiv_rn(dmul(var, 0.25), dadd(dmul(avg, avg), (double)1e-6f));
                }
            }
            const long long o = row * (long long)cols + col;
            labels[o] = curv > 0.1 ? 1 : 0;
            if (kWriteCurv) curv_out[o] = curv;
        }
        __syncthreads();
    }
}

void launch_labels(const double *cloud, int *labels, double *curv_or_null, long long n_rows, int cols,
                   int sm_count, cudaStream_t stream) {
    if (n_rows <= 0 || cols <= 0) return;
    const int tiles_per_row = div_up(cols, kTile);
    const long long n_tiles = n_rows * tiles_per_row;
    long long grid = (long long)sm_count * 8;
    if (grid > n_tiles) grid = n_tiles;
    if (curv_or_null)
        k_labels_exact<true><<<(unsigned)grid, kTile, 0, stream>>>(cloud, labels, curv_or_null, n_rows, cols,
                                                                  tiles_per_row);
    else
        k_labels_exact<false><<<(unsigned)grid, kTile, 0, stream>>>(cloud, labels, nullptr, n_rows, cols,
                                                                   tiles_per_row);
}

// ---------------------------------------------------------------- a2 ------------------------
// utils/pointcloud.c:18-47.  tan(theta_col) / tan(phi_row) come from the host's libm (the
// reference's own tan), uploaded once per context; y = (-d) * tan is one rounded multiply.
__global__ void k_convert(const int *__restrict__ dist, const double *__restrict__ tan_col,
                          const double *__restrict__ tan_row, double *__restrict__ out, int rows, int cols) {
    const long long n = (long long)rows * cols;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols), c = (int)(i % cols);
        const double d = (double)dist[i];
        double x = 0.0, y = 0.0, z = 0.0;
        if (!(d <= 0.0)) {
            x = d;
            y = dmul(-d, tan_col[c]);
            z = dmul(-d, tan_row[r]);
        }
        out[i * 3 + 0] = x;
        out[i * 3 + 1] = y;
        out[i * 3 + 2] = z;
    }
}

void launch_convert(const int *dist, const double *tan_col, const double *tan_row, double *out, int rows,
                    int cols, int sm_count, cudaStream_t stream) {
    const long long n = (long long)rows * cols;
    if (n <= 0) return;
    long long grid = (n + 255) / 256;
    if (grid > sm_count * 8LL) grid = sm_count * 8LL;
    k_convert<<<(unsigned)grid, 256, 0, stream>>>(dist, tan_col, tan_row, out, rows, cols);
}

// ---------------------------------------------------------------- a7 ------------------------
__global__ void k_transform(const double *__restrict__ in, double *__restrict__ out, long long n, PoseXf pose) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        P3 p = {in[i * 3], in[i * 3 + 1], in[i * 3 + 2]};
        P3 g = xf_point(pose, p);
        out[i * 3] = g.x;
        out[i * 3 + 1] = g.y;
        out[i * 3 + 2] = g.z;
    }
}

void launch_transform(const double *in, double *out, long long n, const PoseXf &pose, int sm_count,
                      cudaStream_t stream) {
    if (n <= 0) return;
    long long grid = (n + 255) / 256;
    if (grid > sm_count * 8LL) grid = sm_count * 8LL;
    k_transform<<<(unsigned)grid, 256, 0, stream>>>(in, out, n, pose);
}

}  // namespace nav
