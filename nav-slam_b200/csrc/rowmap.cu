// rowmap.cu -- the per-row map ("kdtree_lastframe[row]", headers/slam.h:14) of the SLAM step and
// the per-row exact nearest-neighbour match against it.
//
// What the reference does (src/slam.c:167-172, 236-284, 422-427; utils/kdtree.c): for every image
// row it compacts the previous frame's edge points (global frame) into an array (flattenPoints),
// builds a pointer kd-tree over those <= MAX_COLS points, and answers one exact 1-NN query per
// labelled point of the current frame against the tree of the *same row*; results are then
// de-duplicated per row.
//
// What this file does instead (same answers, B200-shaped).  A lidar ring is a polyline: points of
// neighbouring columns are neighbours in space.  So the "tree" of a row is simply the mapped
// global cloud of that row, left in place, plus
//     * a 16-bit label mask per block of 16 columns (which of the 16 points are map points),
//     * the bounding box of the labelled points of every 16-column block (leaf box),
//     * the bounding box of every 256-column block (super box).
// Everything is local to a (row, 256-column tile), so both kernels run one independent CTA per
// tile -- 512 CTAs for a 64x2048 image -- with no cross-CTA prefix sums and no compaction (a4) on
// the matching path at all.  Correctness never depends on the polyline assumption: a query visits
// every block whose box lower bound is <= its current best, and the lower bound is a true lower
// bound of the rounded distance the reference computes (nav_common.cuh).
//
//   k_frame_map   : rigid transform of the tile (a7) -> global cloud, label masks, leaf/super boxes
//   k_frame_match : curvature/labels of the tile (a3, fp32-filtered + exact fallback), then one
//                   thread per labelled column: query transform (a7) and exact search (a5/a6),
//                   seeded with the leaf block of its own column.  Ties -> lowest column.
//   k_dedupe_rows : per-row dedupe of src/slam.c:247-283 (a8) with shared-memory atomics
//   k_export_row / k_flatten_row : stable row compaction (a4) for the API and the shim
#include <limits.h>
#include <math.h>

#include "nav_kernels.cuh"
#include "stencil_tile.cuh"

namespace nav {

constexpr unsigned kFull = 0xffffffffu;
static_assert(kTile == kChunk * kChunksPerSuper, "one CTA tile = one super block of 16 leaf blocks");

__device__ __forceinline__ P3 load_p3(const double *__restrict__ p) {
    P3 v = {p[0], p[1], p[2]};
    return v;
}
__device__ __forceinline__ P3 ldg_p3(const double *__restrict__ p) {
    P3 v = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
    return v;
}
__device__ __forceinline__ void store_p3(double *__restrict__ p, const P3 &v) {
    p[0] = v.x;
    p[1] = v.y;
    p[2] = v.z;
}

// min/max over the 16 lanes of a half warp
__device__ __forceinline__ void half_minmax(double &lo, double &hi) {
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(kFull, lo, d, 16));
        hi = fmax(hi, __shfl_xor_sync(kFull, hi, d, 16));
    }
}

// ---------------------------------------------------------------------------------------------
// grid = n_rows * tiles_per_row CTAs of kTile threads; thread = one column
__global__ void __launch_bounds__(kTile)
k_frame_map(const double *__restrict__ cloud, const int *__restrict__ labels, RowMap map,
            const __grid_constant__ PoseBatch poses, int rows, int cols, int tiles_per_row) {
    __shared__ double s_box[kChunksPerSuper][6];
    const int rid = blockIdx.x / tiles_per_row;  // sequence * rows + row
    const int tile = blockIdx.x % tiles_per_row;
    const int seq = rid / rows;
    const long long base = (long long)rid * cols;
    const PoseXf &pose = poses.p[seq];
    const int c = tile * kTile + threadIdx.x;
    const bool valid = c < cols;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, l16 = lane & 15;

    bool lab = false;
    P3 g = {0, 0, 0};
    if (valid) {
        lab = labels[base + c] == 1;
        g = xf_point(pose, ldg_p3(cloud + (base + c) * 3));
        store_p3(map.pts + (base + c) * 3, g);
    }
    const unsigned ballot = __ballot_sync(kFull, lab);
    const unsigned mask16 = (ballot >> (half * 16)) & 0xffffu;
    double lo[3], hi[3];
    lo[0] = lab ? g.x : INFINITY;
    lo[1] = lab ? g.y : INFINITY;
    lo[2] = lab ? g.z : INFINITY;
    hi[0] = lab ? g.x : -INFINITY;
    hi[1] = lab ? g.y : -INFINITY;
    hi[2] = lab ? g.z : -INFINITY;
#pragma unroll
    for (int a = 0; a < 3; ++a) half_minmax(lo[a], hi[a]);
    const int leaf_in_tile = warp * 2 + half;
    const int leaf = tile * kChunksPerSuper + leaf_in_tile;
    if (l16 == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            s_box[leaf_in_tile][a] = lo[a];
            s_box[leaf_in_tile][3 + a] = hi[a];
        }
        if (leaf < map.n_chunks) {
            map.mask[(long long)rid * map.n_chunks + leaf] = mask16;
            double *b = map.box + ((long long)rid * map.n_chunks + leaf) * 6;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                b[a] = lo[a];
                b[3 + a] = hi[a];
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        double slo[3], shi[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            slo[a] = s_box[l16][a];
            shi[a] = s_box[l16][3 + a];
            half_minmax(slo[a], shi[a]);
        }
        if (lane == 0) {
            double *b = map.sbox + ((long long)rid * map.n_super + tile) * 6;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                b[a] = slo[a];
                b[3 + a] = shi[a];
            }
        }
    }
}

// scan the labelled points of one 16-column block of the map row
__device__ __forceinline__ void scan_leaf(const double *__restrict__ row_pts, unsigned mask, int col0, const P3 &q,
                                          double &best, int &bcol) {
    while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const int col = col0 + b;
        const double *p = row_pts + (long long)col * 3;
        // operand order of euclideanDistance(root->point, *target), utils/kdtree.c:116
        const double d = dsq3(dsub(__ldg(p), q.x), dsub(__ldg(p + 1), q.y), dsub(__ldg(p + 2), q.z));
        if (d < best || (d == best && col < bcol)) {
            best = d;
            bcol = col;
        }
    }
}

// grid as k_frame_map.  kFusedLabels: compute the labels of the tile here (and store them);
// otherwise read them from `labels`.
template <bool kFusedLabels>
__global__ void __launch_bounds__(kTile)
k_frame_match(const double *__restrict__ cloud, int *__restrict__ labels, RowMap map, MatchOut out,
              const __grid_constant__ PoseBatch poses, int rows, int cols, int tiles_per_row,
              unsigned *__restrict__ n_exact) {
    __shared__ StencilSmem s;
    const int rid = blockIdx.x / tiles_per_row;
    const int tile = blockIdx.x % tiles_per_row;
    const int seq = rid / rows, row = rid % rows;
    const long long base = (long long)rid * cols;
    const int c0 = tile * kTile;
    const int c = c0 + threadIdx.x;

    int label;
    P3 p;
    if (kFusedLabels) {
        tile_stage(s, cloud + base * 3, c0, cols);
        __syncthreads();
        label = tile_labels_filtered(s, c0, cols, n_exact);
        const double *sp = s.pts + (threadIdx.x + kHalo) * 3;
        p.x = sp[0];
        p.y = sp[1];
        p.z = sp[2];
        if (c < cols) labels[base + c] = label;
    } else {
        label = c < cols ? labels[base + c] : 0;
        if (c < cols) p = ldg_p3(cloud + (base + c) * 3);
    }
    if (c >= cols) return;
    if (label != 1) {
        out.nn_idx[base + c] = -1;
        out.nn_dist[base + c] = -1.0;
        return;
    }
    const PoseXf &pose = poses.p[seq];
    const P3 q = shift_point(pose, xf_point(pose, p));

    const double *m_pts = map.pts + base * 3;
    const unsigned *m_mask = map.mask + (long long)rid * map.n_chunks;
    const double *m_box = map.box + (long long)rid * map.n_chunks * 6;
    const double *m_sbox = map.sbox + (long long)rid * map.n_super * 6;
    const int n_leaf = map.n_chunks, n_sup = map.n_super;

    double best = INFINITY;
    int bcol = -1;
    const int seed = c / kChunk;
    scan_leaf(m_pts, __ldg(m_mask + seed), seed * kChunk, q, best, bcol);
    for (int sc = 0; sc < n_sup; ++sc) {
        double bb[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) bb[a] = __ldg(m_sbox + sc * 6 + a);
        if (!(box_lower_bound(bb, q) <= best)) continue;
        const int l1 = min(n_leaf, (sc + 1) * kChunksPerSuper);
        for (int lf = sc * kChunksPerSuper; lf < l1; ++lf) {
            if (lf == seed) continue;
#pragma unroll
            for (int a = 0; a < 6; ++a) bb[a] = __ldg(m_box + lf * 6 + a);
            if (!(box_lower_bound(bb, q) <= best)) continue;
            scan_leaf(m_pts, __ldg(m_mask + lf), lf * kChunk, q, best, bcol);
        }
    }
    out.nn_idx[base + c] = bcol >= 0 ? row * cols + bcol : -1;
    out.nn_dist[base + c] = bcol >= 0 ? __dsqrt_rn(best) : INFINITY;
}

void launch_frame_map(const double *cloud, const int *labels, const RowMap &map, const PoseBatch &poses,
                      int n_seq, int rows, int cols, cudaStream_t stream) {
    const int tiles = div_up(cols, kTile);
    k_frame_map<<<n_seq * rows * tiles, kTile, 0, stream>>>(cloud, labels, map, poses, rows, cols, tiles);
}

void launch_frame_match(const double *cloud, int *labels, bool fused_labels, const RowMap &map,
                        const MatchOut &out, const PoseBatch &poses, int n_seq, int rows, int cols,
                        unsigned *n_exact, cudaStream_t stream) {
    const int tiles = div_up(cols, kTile);
    const int grid = n_seq * rows * tiles;
    if (fused_labels)
        k_frame_match<true><<<grid, kTile, 0, stream>>>(cloud, labels, map, out, poses, rows, cols, tiles, n_exact);
    else
        k_frame_match<false><<<grid, kTile, 0, stream>>>(cloud, labels, map, out, poses, rows, cols, tiles, n_exact);
}

// ---------------------------------------------------------------------------------------------
// exclusive prefix of `pred` over the block (thread order), plus the block total
__device__ __forceinline__ int block_excl_count(bool pred, int *s_warp, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_warps = blockDim.x >> 5;
    const unsigned m = __ballot_sync(kFull, pred);
    const int in_warp = __popc(m & ((1u << lane) - 1u));
    __syncthreads();  // protects s_warp reuse across rounds
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
        int v = lane < n_warps ? s_warp[lane] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += t;
        }
        s_warp[32 + lane] = incl - v;
        if (lane == 31) s_warp[64] = incl;
    }
    __syncthreads();
    total = s_warp[64];
    return s_warp[32 + warp] + in_warp;
}

constexpr int kRowThreads = 512;

// per-row dedupe, src/slam.c:247-283: one entry per matched map point; the query with the smallest
// distance wins (earliest column on equal distance, strict '>' at slam.c:264); entries in order of
// the first query that matched the point.  One CTA per (sequence,row); dynamic smem = 16 B * cols.
__global__ void __launch_bounds__(kRowThreads)
k_dedupe_rows(const double *__restrict__ cloud, const int *__restrict__ labels, RowMap map, MatchOut out,
              const __grid_constant__ PoseBatch poses, int rows, int cols) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_warp[65];
    unsigned long long *s_best = (unsigned long long *)s_raw;
    int *s_first = (int *)(s_best + cols);
    int *s_win = s_first + cols;
    const int rid = blockIdx.x;
    const int seq = rid / rows, row = rid % rows;
    const long long base = (long long)rid * cols;
    const PoseXf &pose = poses.p[seq];
    for (int j = threadIdx.x; j < cols; j += kRowThreads) {
        s_best[j] = ~0ull;
        s_first[j] = INT_MAX;
        s_win[j] = INT_MAX;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cols; c += kRowThreads) {
        const int idx = out.nn_idx[base + c];
        if (labels[base + c] != 1 || idx < 0) continue;
        const int key = idx - row * cols;
        atomicMin(&s_best[key], (unsigned long long)__double_as_longlong(out.nn_dist[base + c]));
        atomicMin(&s_first[key], c);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cols; c += kRowThreads) {
        const int idx = out.nn_idx[base + c];
        if (labels[base + c] != 1 || idx < 0) continue;
        const int key = idx - row * cols;
        if ((unsigned long long)__double_as_longlong(out.nn_dist[base + c]) == s_best[key]) atomicMin(&s_win[key], c);
    }
    __syncthreads();
    int n_out = 0;
    nav_corr *rows_out = out.corr_rows + base;
    for (int c0 = 0; c0 < cols; c0 += kRowThreads) {
        const int c = c0 + threadIdx.x;
        int key = -1;
        bool is_first = false;
        if (c < cols && labels[base + c] == 1) {
            const int idx = out.nn_idx[base + c];
            if (idx >= 0) {
                key = idx - row * cols;
                is_first = s_first[key] == c;
            }
        }
        int total;
        const int pos = n_out + block_excl_count(is_first, s_warp, total);
        if (is_first) {
            const P3 ori = xf_point(pose, load_p3(cloud + (base + s_win[key]) * 3));
            const double *np = map.pts + (base + key) * 3;
            nav_corr e;
            e.ori.x = ori.x;
            e.ori.y = ori.y;
            e.ori.z = ori.z;
            e.nearest.x = np[0];
            e.nearest.y = np[1];
            e.nearest.z = np[2];
            e.distance = __longlong_as_double((long long)s_best[key]);
            rows_out[pos] = e;
        }
        n_out += total;
    }
    if (threadIdx.x == 0) out.corr_row_count[rid] = n_out;
}

size_t dedupe_smem_bytes(int cols) { return (size_t)cols * 16; }

int configure_row_kernels(int cols) {
    static int configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    const int need = (int)dedupe_smem_bytes(cols);
    if (need <= 48 * 1024 || (dev >= 0 && dev < 64 && configured[dev] >= need)) return 0;
    cudaError_t e = cudaFuncSetAttribute(k_dedupe_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
    if (e == cudaSuccess && dev >= 0 && dev < 64) configured[dev] = need;
    return (int)e;
}

void launch_dedupe(const double *cloud, const int *labels, const RowMap &map, const MatchOut &out,
                   const PoseBatch &poses, int n_seq, int rows, int cols, cudaStream_t stream) {
    k_dedupe_rows<<<n_seq * rows, kRowThreads, dedupe_smem_bytes(cols), stream>>>(cloud, labels, map, out, poses,
                                                                               rows, cols);
}

// rows of one sequence back to back: corr_out[seq][offset(row) + i]
__global__ void k_gather_corr(const nav_corr *__restrict__ corr_rows, const int *__restrict__ row_count,
                              nav_corr *__restrict__ corr_out, int *__restrict__ corr_total, int rows, int cols) {
    __shared__ int s_off;
    const int rid = blockIdx.x;
    const int seq = rid / rows, row = rid % rows;
    if (threadIdx.x == 0) {
        int off = 0;
        for (int r = 0; r < row; ++r) off += row_count[seq * rows + r];
        s_off = off;
        if (row == rows - 1) corr_total[seq] = off + row_count[rid];
    }
    __syncthreads();
    const int n = row_count[rid];
    const double *src = (const double *)(corr_rows + (long long)rid * cols);
    double *dst = (double *)(corr_out + (long long)seq * rows * cols + s_off);
    for (int i = threadIdx.x; i < n * 7; i += blockDim.x) dst[i] = src[i];
}

void launch_gather_corr(const nav_corr *corr_rows, const int *corr_row_count, nav_corr *corr_out,
                        int *corr_total, int n_seq, int rows, int cols, cudaStream_t stream) {
    k_gather_corr<<<n_seq * rows, 256, 0, stream>>>(corr_rows, corr_row_count, corr_out, corr_total, rows, cols);
}

// ---------------------------------------------------------------------------------------------
// flattenPoints (src/slam.c:64-72): stable compaction of one row where feature == 1 (function-level
// mirror), or where the map's label mask is set (k_export_row: the flattenedPoints array the
// reference would hand to buildKDTree, src/slam.c:170-171)
__global__ void __launch_bounds__(kRowThreads)
k_flatten_row(const double *__restrict__ row_pts, const int *__restrict__ row_feature,
              const unsigned *__restrict__ row_mask, double *__restrict__ out, int *__restrict__ col_out,
              int *__restrict__ count, int cols) {
    __shared__ int s_warp[65];
    int carry = 0;
    for (int c0 = 0; c0 < cols; c0 += kRowThreads) {
        const int c = c0 + threadIdx.x;
        bool lab = false;
        if (c < cols) lab = row_feature ? row_feature[c] == 1 : ((row_mask[c / kChunk] >> (c % kChunk)) & 1u) != 0;
        int total;
        const int pos = carry + block_excl_count(lab, s_warp, total);
        if (lab) {
            store_p3(out + (long long)pos * 3, load_p3(row_pts + (long long)c * 3));
            if (col_out) col_out[pos] = c;
        }
        carry += total;
    }
    if (threadIdx.x == 0) *count = carry;
}

void launch_flatten_row(const double *row_pts, const int *row_feature, const unsigned *row_mask, double *out,
                        int *col_out, int *count, int cols, cudaStream_t stream) {
    k_flatten_row<<<1, kRowThreads, 0, stream>>>(row_pts, row_feature, row_mask, out, col_out, count, cols);
}

}  // namespace nav
