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
#include "stencil_tile.cuh"

namespace nav {

// labels for n_rows image rows (any number of images back to back): persistent grid-stride loop
// over (row, 256-column tile) pairs, fp32-filtered evaluation with exact fallback (stencil_tile.cuh)
__global__ void __launch_bounds__(kTile)
k_labels(const double *__restrict__ cloud, int *__restrict__ labels, long long n_rows, int cols,
         int tiles_per_row, unsigned *__restrict__ n_exact) {
    __shared__ StencilSmem s;
    const long long n_tiles = n_rows * tiles_per_row;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row = tile / tiles_per_row;
        const int c0 = (int)(tile % tiles_per_row) * kTile;
        tile_stage(s, cloud + row * (long long)cols * 3, c0, cols);
        __syncthreads();
        const int label = tile_labels_filtered(s, c0, cols, n_exact);
        const int col = c0 + threadIdx.x;
        if (col < cols) labels[row * (long long)cols + col] = label;
    }
}

// exact binary64 curvature (test hook of the C ABI) and the labels derived from it
__global__ void __launch_bounds__(kTile)
k_curvature(const double *__restrict__ cloud, int *__restrict__ labels, double *__restrict__ curv_out,
            long long n_rows, int cols, int tiles_per_row) {
    __shared__ StencilSmem s;
    const long long n_tiles = n_rows * tiles_per_row;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row = tile / tiles_per_row;
        const int c0 = (int)(tile % tiles_per_row) * kTile;
        tile_stage(s, cloud + row * (long long)cols * 3, c0, cols);
        __syncthreads();
        const double curv = tile_curvature_exact(s, c0, cols);
        const int col = c0 + threadIdx.x;
        if (col < cols) {
            const long long o = row * (long long)cols + col;
            labels[o] = curv > 0.1 ? 1 : 0;
            curv_out[o] = curv;
        }
    }
}

void launch_labels(const double *cloud, int *labels, double *curv_or_null, long long n_rows, int cols,
                   int sm_count, unsigned *n_exact, cudaStream_t stream) {
    if (n_rows <= 0 || cols <= 0) return;
    const int tiles_per_row = div_up(cols, kTile);
    const long long n_tiles = n_rows * tiles_per_row;
    long long grid = (long long)sm_count * 8;
    if (grid > n_tiles) grid = n_tiles;
    if (curv_or_null)
        k_curvature<<<(unsigned)grid, kTile, 0, stream>>>(cloud, labels, curv_or_null, n_rows, cols, tiles_per_row);
    else
        k_labels<<<(unsigned)grid, kTile, 0, stream>>>(cloud, labels, n_rows, cols, tiles_per_row, n_exact);
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
