// rowmap.cu -- the per-row map ("kdtree_lastframe[row]", headers/slam.h:14) of the SLAM step and
// the per-row exact nearest-neighbour match against it.
//
// What the reference does (src/slam.c:167-172, 236-284, 422-427; utils/kdtree.c): for every image
// row it compacts the previous frame's edge points (global frame) into an array, builds a pointer
// kd-tree over those <= MAX_COLS points, and answers one exact 1-NN query per labelled point of the
// current frame against the tree of the *same row*; results are then de-duplicated per row.
//
// What this file does instead (same answers, B200-shaped): one CTA per (sequence,row).
//   * k_map_build: rigid transform of the whole row (a7) + stable ballot/prefix compaction of the
//     labelled points (a4) + 16-point leaf boxes and 256-point super boxes over the compacted
//     run (a5).  A lidar ring is a polyline, so consecutive points are spatially coherent and the
//     boxes are tight; nothing depends on that for correctness.
//   * k_match: the row's map (<= cols points, 24 B each) and boxes are staged in shared memory
//     once; every labelled point of the current frame is transformed to its query (a7) by its own
//     thread and searched exactly: seed with the leaf at the query's own column rank, then visit
//     only super boxes / leaf boxes whose rounded lower bound is <= the current best.  Ties go to
//     the lowest map index (lexicographic (dsq, index) compare).  Optional per-row dedupe (a8) with
//     shared-memory atomics reproduces the reference's "closest query per matched point, order of
//     first appearance" list.
#include <limits.h>
#include <math.h>

#include "nav_kernels.cuh"

namespace nav {

constexpr int kRowThreads = 512;
constexpr unsigned kFull = 0xffffffffu;

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
        s_warp[32 + lane] = incl - v;  // exclusive warp offsets
        if (lane == 31) s_warp[64] = incl;
    }
    __syncthreads();
    total = s_warp[64];
    return s_warp[32 + warp] + in_warp;
}

__device__ __forceinline__ P3 load_p3(const double *__restrict__ p) {
    P3 v = {p[0], p[1], p[2]};
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
__global__ void __launch_bounds__(kRowThreads)
k_map_build(const double *__restrict__ cloud, const int *__restrict__ labels, double *__restrict__ global_out,
            RowMap map, const __grid_constant__ PoseBatch poses, int rows, int cols) {
    __shared__ int s_warp[65];
    const int rid = blockIdx.x;
    const int seq = rid / rows;
    const long long base = (long long)rid * cols;
    const PoseXf &pose = poses.p[seq];
    double *m_pts = map.pts + base * 3;
    int *m_col = map.col + base;
    int *m_rank = map.rank + base;

    int carry = 0;
    for (int c0 = 0; c0 < cols; c0 += kRowThreads) {
        const int c = c0 + threadIdx.x;
        const bool valid = c < cols;
        bool lab = false;
        P3 g = {0, 0, 0};
        if (valid) {
            lab = labels[base + c] == 1;
            g = xf_point(pose, load_p3(cloud + (base + c) * 3));
            if (global_out) store_p3(global_out + (base + c) * 3, g);
        }
        int total;
        const int pos = carry + block_excl_count(lab, s_warp, total);
        if (valid) m_rank[c] = pos;
        if (lab) {
            store_p3(m_pts + (long long)pos * 3, g);
            m_col[pos] = c;
        }
        carry += total;
    }
    const int n = carry;
    if (threadIdx.x == 0) map.count[rid] = n;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, l16 = lane & 15;
    const int n_warps = kRowThreads >> 5;
    const int n_ch = div_up(n, kChunk);
    double *m_box = map.box + (long long)rid * map.n_chunks * 6;
    for (int chb = warp * 2; chb < n_ch; chb += n_warps * 2) {
        const int ch = chb + half;
        const int j = ch * kChunk + l16;
        const bool have = ch < n_ch && j < n;
        double lo[3], hi[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double v = have ? m_pts[(long long)j * 3 + a] : 0.0;
            lo[a] = have ? v : INFINITY;
            hi[a] = have ? v : -INFINITY;
            half_minmax(lo[a], hi[a]);
        }
        if (ch < n_ch && l16 == 0) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                m_box[ch * 6 + a] = lo[a];
                m_box[ch * 6 + 3 + a] = hi[a];
            }
        }
    }
    __syncthreads();
    const int n_sc = div_up(n_ch, kChunksPerSuper);
    double *m_sbox = map.sbox + (long long)rid * map.n_super * 6;
    for (int scb = warp * 2; scb < n_sc; scb += n_warps * 2) {
        const int sc = scb + half;
        const int ch = sc * kChunksPerSuper + l16;
        const bool have = sc < n_sc && ch < n_ch;
        double lo[3], hi[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            lo[a] = have ? m_box[ch * 6 + a] : INFINITY;
            hi[a] = have ? m_box[ch * 6 + 3 + a] : -INFINITY;
            half_minmax(lo[a], hi[a]);
        }
        if (sc < n_sc && l16 == 0) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                m_sbox[sc * 6 + a] = lo[a];
                m_sbox[sc * 6 + 3 + a] = hi[a];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
struct MatchSmem {
    double *map;   // [cols*3]
    double *box;   // [n_chunks*6]
    double *sbox;  // [n_super*6]
    unsigned long long *bestd;  // [cols]   (dedupe)
    int *qcol;     // [cols]
    int *qj;       // [cols]
    int *wincol;   // [cols]   (dedupe)
    int *first;    // [cols]   (dedupe)
};

static __host__ __device__ size_t match_layout(int cols, bool dedupe, MatchSmem *s, unsigned char *basep) {
    const int n_ch = div_up(cols, kChunk), n_sc = div_up(n_ch, kChunksPerSuper);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 15) & ~(size_t)15;
        return o;
    };
    size_t o_map = take((size_t)cols * 3 * 8), o_box = take((size_t)n_ch * 6 * 8), o_sbox = take((size_t)n_sc * 6 * 8);
    size_t o_bestd = dedupe ? take((size_t)cols * 8) : 0;
    size_t o_qcol = take((size_t)cols * 4), o_qj = take((size_t)cols * 4);
    size_t o_win = dedupe ? take((size_t)cols * 4) : 0, o_first = dedupe ? take((size_t)cols * 4) : 0;
    if (s) {
        s->map = (double *)(basep + o_map);
        s->box = (double *)(basep + o_box);
        s->sbox = (double *)(basep + o_sbox);
        s->bestd = (unsigned long long *)(basep + o_bestd);
        s->qcol = (int *)(basep + o_qcol);
        s->qj = (int *)(basep + o_qj);
        s->wincol = (int *)(basep + o_win);
        s->first = (int *)(basep + o_first);
    }
    return off;
}

size_t match_smem_bytes(int cols, bool dedupe) { return match_layout(cols, dedupe, nullptr, nullptr); }
size_t map_smem_bytes(int) { return 0; }

__device__ __forceinline__ void scan_leaf(const double *__restrict__ s_map, int ch, int n, const P3 &q,
                                          double &best, int &bj) {
    const int j0 = ch * kChunk;
    const int j1 = min(n, j0 + kChunk);
    for (int j = j0; j < j1; ++j) {
        const double *p = s_map + j * 3;
        // operand order of euclideanDistance(root->point, *target), utils/kdtree.c:116
        const double d = dsq3(dsub(p[0], q.x), dsub(p[1], q.y), dsub(p[2], q.z));
        if (d < best || (d == best && j < bj)) {
            best = d;
            bj = j;
        }
    }
}

template <bool kDedupe>
__global__ void __launch_bounds__(kRowThreads)
k_match(const double *__restrict__ cloud, const int *__restrict__ labels, RowMap map, MatchOut out,
        const __grid_constant__ PoseBatch poses, int rows, int cols) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ int s_warp[65];
    MatchSmem s;
    match_layout(cols, kDedupe, &s, s_raw);

    const int rid = blockIdx.x;
    const int seq = rid / rows, row = rid % rows;
    const long long base = (long long)rid * cols;
    const PoseXf &pose = poses.p[seq];
    const int n = map.count[rid];
    const int n_ch = div_up(n, kChunk), n_sc = div_up(n_ch, kChunksPerSuper);

    // stage the row's map and boxes
    {
        const double *g_pts = map.pts + base * 3;
        for (int i = threadIdx.x; i < n * 3; i += kRowThreads) s.map[i] = g_pts[i];
        const double *g_box = map.box + (long long)rid * map.n_chunks * 6;
        for (int i = threadIdx.x; i < n_ch * 6; i += kRowThreads) s.box[i] = g_box[i];
        const double *g_sbox = map.sbox + (long long)rid * map.n_super * 6;
        for (int i = threadIdx.x; i < n_sc * 6; i += kRowThreads) s.sbox[i] = g_sbox[i];
        if (kDedupe) {
            for (int j = threadIdx.x; j < n; j += kRowThreads) {
                s.bestd[j] = ~0ull;
                s.wincol[j] = INT_MAX;
                s.first[j] = INT_MAX;
            }
        }
    }

    // compact the labelled columns of the current frame (query order = ascending column)
    int nq = 0;
    for (int c0 = 0; c0 < cols; c0 += kRowThreads) {
        const int c = c0 + threadIdx.x;
        const bool lab = c < cols && labels[base + c] == 1;
        int total;
        const int pos = nq + block_excl_count(lab, s_warp, total);
        if (lab) s.qcol[pos] = c;
        if (c < cols && !lab) {
            out.nn_idx[base + c] = -1;
            out.nn_dist[base + c] = -1.0;
        }
        nq += total;
    }
    __syncthreads();

    const int *m_col = map.col + base;
    const int *m_rank = map.rank + base;
    for (int k = threadIdx.x; k < nq; k += kRowThreads) {
        const int c = s.qcol[k];
        const P3 q = shift_point(pose, xf_point(pose, load_p3(cloud + (base + c) * 3)));
        double best = INFINITY;
        int bj = -1;
        if (n > 0) {
            const int seed = min(m_rank[c], n - 1) / kChunk;
            scan_leaf(s.map, seed, n, q, best, bj);
            for (int sc = 0; sc < n_sc; ++sc) {
                if (!(box_lower_bound(s.sbox + sc * 6, q) <= best)) continue;
                const int ch1 = min(n_ch, (sc + 1) * kChunksPerSuper);
                for (int ch = sc * kChunksPerSuper; ch < ch1; ++ch) {
                    if (ch == seed) continue;
                    if (!(box_lower_bound(s.box + ch * 6, q) <= best)) continue;
                    scan_leaf(s.map, ch, n, q, best, bj);
                }
            }
        }
        const double dist = bj >= 0 ? __dsqrt_rn(best) : INFINITY;
        out.nn_idx[base + c] = bj >= 0 ? row * cols + m_col[bj] : -1;
        out.nn_dist[base + c] = dist;
        s.qj[k] = bj;
        if (kDedupe && bj >= 0) {
            atomicMin(&s.bestd[bj], (unsigned long long)__double_as_longlong(dist));
            atomicMin(&s.first[bj], c);
        }
    }
    if (!kDedupe) return;

    // ---- per-row dedupe, src/slam.c:247-283: one entry per matched point, the query with the
    // smallest distance wins (earliest column on equal distance), entries in order of the first
    // query that matched the point
    __syncthreads();
    for (int k = threadIdx.x; k < nq; k += kRowThreads) {
        const int bj = s.qj[k];
        if (bj < 0) continue;
        const int c = s.qcol[k];
        const unsigned long long d = (unsigned long long)__double_as_longlong(out.nn_dist[base + c]);
        if (d == s.bestd[bj]) atomicMin(&s.wincol[bj], c);
    }
    __syncthreads();
    int n_out = 0;
    nav_corr *rows_out = out.corr_rows + base;
    for (int k0 = 0; k0 < nq; k0 += kRowThreads) {
        const int k = k0 + threadIdx.x;
        int bj = -1;
        bool is_first = false;
        if (k < nq) {
            bj = s.qj[k];
            is_first = bj >= 0 && s.first[bj] == s.qcol[k];
        }
        int total;
        const int pos = n_out + block_excl_count(is_first, s_warp, total);
        if (is_first) {
            const int wc = s.wincol[bj];
            const P3 ori = xf_point(pose, load_p3(cloud + (base + wc) * 3));
            nav_corr e;
            e.ori.x = ori.x;
            e.ori.y = ori.y;
            e.ori.z = ori.z;
            e.nearest.x = s.map[bj * 3];
            e.nearest.y = s.map[bj * 3 + 1];
            e.nearest.z = s.map[bj * 3 + 2];
            e.distance = __longlong_as_double((long long)s.bestd[bj]);
            rows_out[pos] = e;
        }
        n_out += total;
    }
    if (threadIdx.x == 0) out.corr_row_count[rid] = n_out;
}

// The attribute is per function and per device, not per context: only ever raise it.
int configure_row_kernels(int cols) {
    static int configured[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    const int need = (int)match_smem_bytes(cols, true);
    if (dev >= 0 && dev < 64 && configured[dev] >= need) return 0;
    cudaError_t e;
    e = cudaFuncSetAttribute(k_match<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_match<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, need);
    if (e == cudaSuccess && dev >= 0 && dev < 64) configured[dev] = need;
    return (int)e;
}

void launch_map_build(const double *cloud, const int *labels, double *global_out, const RowMap &map,
                      const PoseBatch &poses, int n_seq, int rows, int cols, cudaStream_t stream) {
    k_map_build<<<n_seq * rows, kRowThreads, 0, stream>>>(cloud, labels, global_out, map, poses, rows, cols);
}

void launch_match(const double *cloud, const int *labels, const RowMap &map, const MatchOut &out,
                  const PoseBatch &poses, int n_seq, int rows, int cols, bool dedupe, cudaStream_t stream) {
    const size_t smem = match_smem_bytes(cols, dedupe);
    if (dedupe)
        k_match<true><<<n_seq * rows, kRowThreads, smem, stream>>>(cloud, labels, map, out, poses, rows, cols);
    else
        k_match<false><<<n_seq * rows, kRowThreads, smem, stream>>>(cloud, labels, map, out, poses, rows, cols);
}

// ---------------------------------------------------------------------------------------------
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
// flattenPoints (src/slam.c:64-72) for one row, function-level mirror
__global__ void __launch_bounds__(kRowThreads)
k_flatten_row(const double *__restrict__ row_pts, const int *__restrict__ row_feature, double *__restrict__ out,
              int *__restrict__ count, int cols) {
    __shared__ int s_warp[65];
    int carry = 0;
    for (int c0 = 0; c0 < cols; c0 += kRowThreads) {
        const int c = c0 + threadIdx.x;
        const bool lab = c < cols && row_feature[c] == 1;
        int total;
        const int pos = carry + block_excl_count(lab, s_warp, total);
        if (lab) store_p3(out + (long long)pos * 3, load_p3(row_pts + (long long)c * 3));
        carry += total;
    }
    if (threadIdx.x == 0) *count = carry;
}

void launch_flatten_row(const double *row_pts, const int *row_feature, double *out, int *count, int cols,
                        cudaStream_t stream) {
    k_flatten_row<<<1, kRowThreads, 0, stream>>>(row_pts, row_feature, out, count, cols);
}

}  // namespace nav
