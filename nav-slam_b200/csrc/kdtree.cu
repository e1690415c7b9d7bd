// kdtree.cu -- flat implicit kd-tree for large maps (a5) and batched exact 1-NN (a6).
//
// Reference: utils/kdtree.c:65-82 builds a pointer tree by recursive median split (axis =
// depth % 3, median index n/2, children on [0,m) and [m+1,n)) with one malloc per node and an
// in-place Lomuto quick-select; utils/kdtree.c:110-152 answers one query per call by recursion.
//
// Here the tree is the *in-order array* that recursion leaves behind: the node of range [lo,hi)
// sits at mid = lo + (hi-lo)/2, its children own [lo,mid) and [mid+1,hi) -- the same shape rule
// as the reference, so with distinct keys it is the same tree.  Nodes are 32-byte records
// {x,y,z,orig_index} (one DRAM sector each, subtrees contiguous in memory).
//
// Build, level by level, all segments of a level at once (no recursion, no per-node allocation):
//   1. three index lists, each sorted along one axis (radix sort of order-preserving 64-bit keys);
//   2. every segment picks its split axis -- kSplitCyclic: level % 3 like the reference (the exported
//      tree then is the reference's tree); kSplitWidest (default): the axis of largest extent, read off
//      the two ends of the segment in each sorted list.  Maps made of surfaces (walls, floors) have
//      axes along which a segment has no extent; cycling through them doubles the search at every such
//      level, which the widest-extent rule avoids (4x fewer node visits on the accumulated room map);
//   3. the list of that axis holds the segment sorted, so its median is the middle element: mark each
//      point left / median / right;
//   4. all three lists are stably partitioned inside every segment (one prefix sum of packed
//      left/median counts over the 3n positions + one scatter), which keeps them sorted for the levels
//      below (for the list of the split axis this is the identity).
// After ceil(log2 n) levels the three lists coincide and are the in-order layout.  The split axis is
// stored in each node, so the search does not care which rule built the tree.
//
// Query: one thread per query.  Child ranges are pure arithmetic on (lo,hi).  Three interchangeable
// kernels (identical answers, chosen by measurement in kd_nn()): k_kd_nn_stack keeps pending far
// subtrees on a 32-entry thread-local stack (the default); k_kd_nn / k_kd_nn_conv are stackless -- the
// way back up is recovered from two bit masks (which side was taken, parity of each ancestor's size)
// and only the split coordinate of an ancestor is re-read (L1/L2 hits).  Subtrees of <= 8 nodes are
// scanned as a contiguous run.  Far subtrees are visited iff the rounded plane distance^2 is <= the
// current best dsq; candidates compare lexicographically on (dsq, original index): exact NN, lowest
// index on ties.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "nav_kdtree.cuh"

namespace nav {

// ------------------------------------------------------------------------------ build ------
__device__ __forceinline__ unsigned long long order_key(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void k_make_keys(const double *__restrict__ pts, long long n, unsigned long long *__restrict__ kx,
                            unsigned long long *__restrict__ ky, unsigned long long *__restrict__ kz,
                            int *__restrict__ idx) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        kx[i] = order_key(pts[i * 3]);
        ky[i] = order_key(pts[i * 3 + 1]);
        kz[i] = order_key(pts[i * 3 + 2]);
        idx[i] = (int)i;
    }
}

// segment of position p at `level`: descend from the root range by arithmetic.
// returns false if p became a node at a shallower level.
__device__ __forceinline__ bool segment_of(int p, int n, int level, int &lo, int &hi) {
    lo = 0;
    hi = n;
    for (int l = 0; l < level; ++l) {
        const int mid = lo + ((hi - lo) >> 1);
        if (p < mid)
            hi = mid;
        else if (p == mid)
            return false;
        else
            lo = mid + 1;
    }
    return true;
}

// range of segment number s (bits of s, most significant first, = sides taken from the root) at `level`
__device__ __forceinline__ bool segment_range(int s, int n, int level, int &lo, int &hi) {
    lo = 0;
    hi = n;
    for (int l = level - 1; l >= 0; --l) {
        const int mid = lo + ((hi - lo) >> 1);
        if ((s >> l) & 1)
            lo = mid + 1;
        else
            hi = mid;
        if (lo >= hi) return false;
    }
    return true;
}

// split axis of every segment of this level, stored at the position its node will take (mid)
__global__ void k_choose_axis(const double *__restrict__ pts, const int *__restrict__ lists, int n, int level,
                              int rule, unsigned char *__restrict__ axis_at) {
    const long long n_seg = 1ll << level;
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < n_seg;
         s += (long long)gridDim.x * blockDim.x) {
        int lo, hi;
        if (!segment_range((int)s, n, level, lo, hi)) continue;
        int axis = level % 3;
        if (rule == kSplitWidest) {
            double ext[3];
            for (int d = 0; d < 3; ++d) {
                const double first = pts[(long long)lists[(long long)d * n + lo] * 3 + d];
                const double last = pts[(long long)lists[(long long)d * n + hi - 1] * 3 + d];
                ext[d] = last - first;
            }
            axis = 0;  // ties and NaN extents keep the lowest axis
            if (ext[1] > ext[axis]) axis = 1;
            if (ext[2] > ext[axis]) axis = 2;
        }
        axis_at[lo + ((hi - lo) >> 1)] = (unsigned char)axis;
    }
}

// side codes: 0 left, 1 median (becomes the node), 2 right
__global__ void k_mark(const int *__restrict__ lists, int n, int level, const unsigned char *__restrict__ axis_at,
                       int *__restrict__ seg_lo, int *__restrict__ seg_mid, unsigned char *__restrict__ side) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        int lo, hi;
        if (!segment_of(p, n, level, lo, hi)) {
            seg_lo[p] = -1;
            seg_mid[p] = -1;
            continue;
        }
        const int mid = lo + ((hi - lo) >> 1);
        seg_lo[p] = lo;
        seg_mid[p] = mid;
        side[lists[(long long)axis_at[mid] * n + p]] = p < mid ? 0 : (p == mid ? 1 : 2);
    }
}

// the three lists are handled as one array of 3n positions (list a = positions [a*n, (a+1)*n)).
// Input of the prefix sum, evaluated on the fly by the scan kernel (no flag array is materialised):
// packed counters, low word = "goes left", high word = "is the median", of the point at position g.
struct SideFlag {
    const int *lists;
    const int *seg_lo;
    const unsigned char *side;
    long long n;
    __device__ __forceinline__ unsigned long long operator()(long long g) const {
        const long long p = g < n ? g : (g < 2 * n ? g - n : g - 2 * n);
        if (seg_lo[p] < 0) return 0ull;
        const unsigned char sd = side[lists[g]];
        return sd == 0 ? 1ull : (sd == 1 ? (1ull << 32) : 0ull);
    }
};

__global__ void k_scatter(const int *__restrict__ lists, int *__restrict__ lists_out, int n,
                          const int *__restrict__ seg_lo, const int *__restrict__ seg_mid,
                          const unsigned char *__restrict__ side, const unsigned long long *__restrict__ scan) {
    const long long base = (long long)blockIdx.y * n;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const long long g = base + p;
        const int idx = lists[g];
        const int lo = seg_lo[p];
        if (lo < 0) {
            lists_out[g] = idx;
            continue;
        }
        const int mid = seg_mid[p];
        const unsigned long long rel = scan[g] - scan[base + lo];
        const int lefts = (int)(rel & 0xffffffffull), meds = (int)(rel >> 32);
        const unsigned char s = side[idx];
        int dst;
        if (s == 0)
            dst = lo + lefts;
        else if (s == 1)
            dst = mid;
        else
            dst = mid + 1 + ((p - lo) - lefts - meds);
        lists_out[base + dst] = idx;
    }
}

__global__ void k_emit_nodes(const double *__restrict__ pts, const int *__restrict__ order, int n,
                             const unsigned char *__restrict__ axis_at, KdNode *__restrict__ nodes) {
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int i = order[p];
        KdNode nd;
        nd.x = pts[(long long)i * 3];
        nd.y = pts[(long long)i * 3 + 1];
        nd.z = pts[(long long)i * 3 + 2];
        nd.idx = i;
        nd.axis = axis_at[p];
        nodes[p] = nd;
    }
}

// ---- finisher: once every segment of a level holds at most kFinSeg points, one CTA takes one segment
// and builds the whole subtree below it in shared memory -- the same choose / mark / stable partition
// steps as the global levels, with __syncthreads() where those have kernel boundaries -- and writes the
// finished nodes.  Replaces the last ~11 levels (5 launches and several passes over all 3n list entries
// each) by one launch.
constexpr int kFinSeg = 2048;
constexpr int kFinThreads = 512;
constexpr int kFinItems = kFinSeg / kFinThreads;
// lists (2 x 3 x int) + prefix sums + global ids + split axes + side codes
constexpr size_t kFinSmemBytes = sizeof(int) * 6 * kFinSeg + sizeof(unsigned) * kFinSeg + sizeof(int) * kFinSeg + 2 * kFinSeg;

// nodes created by the global levels (depth < level0): thread j in [1, 2^level0) is segment j - 2^L of level L
__global__ void k_emit_top(const double *__restrict__ pts, const int *__restrict__ list0, int n, int level0,
                           const unsigned char *__restrict__ axis_at, KdNode *__restrict__ nodes) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j < 1 || j >= (1ll << level0)) return;
    const int L = 63 - __clzll(j);
    int lo, hi;
    if (!segment_range((int)(j - (1ll << L)), n, L, lo, hi)) return;
    const int mid = lo + ((hi - lo) >> 1);
    const int i = list0[mid];
    KdNode nd;
    nd.x = pts[(long long)i * 3];
    nd.y = pts[(long long)i * 3 + 1];
    nd.z = pts[(long long)i * 3 + 2];
    nd.idx = i;
    nd.axis = axis_at[mid];
    nodes[mid] = nd;
}

__global__ void __launch_bounds__(kFinThreads, 3)
k_kd_finish(const double *__restrict__ pts, const int *__restrict__ lists_in, int n, int level0, int split_rule,
            int *__restrict__ loc, KdNode *__restrict__ nodes) {
    extern __shared__ __align__(16) unsigned char fin_smem[];
    int *cur = reinterpret_cast<int *>(fin_smem);        // [3][kFinSeg]
    int *alt = cur + 3 * kFinSeg;                         // [3][kFinSeg]
    unsigned *sc = reinterpret_cast<unsigned *>(alt + 3 * kFinSeg);  // [kFinSeg] exclusive prefix, lefts | meds << 16
    int *gid = reinterpret_cast<int *>(sc + kFinSeg);    // [kFinSeg] local id -> index of the point in the build input
    unsigned char *ax = reinterpret_cast<unsigned char *>(gid + kFinSeg);  // [kFinSeg] split axis of the node at a position
    unsigned char *side = ax + kFinSeg;                   // [kFinSeg] left / median / right code per local id
    __shared__ unsigned s_wsum[kFinThreads / 32];
    int LO, HI;
    if (!segment_range((int)blockIdx.x, n, level0, LO, HI)) return;
    const int m = HI - LO;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // local ids: the i-th entry of the x-sorted list is point i of this CTA; `loc` (n ints of global
    // scratch, each CTA touches only its own points) translates the two other lists once, after which
    // every level works on shared memory only
    for (int i = tid; i < m; i += kFinThreads) {
        const int g = lists_in[LO + i];
        gid[i] = g;
        loc[g] = i;
        cur[i] = i;
    }
    for (int i = tid; i < kFinSeg; i += kFinThreads) ax[i] = 0;
    __syncthreads();
    for (int d = 1; d < 3; ++d)
        for (int i = tid; i < m; i += kFinThreads) cur[d * kFinSeg + i] = loc[lists_in[(long long)d * n + LO + i]];
    // this thread owns positions 4*tid .. 4*tid+3 (relative to LO); lo_r/hi_r: their current segment,
    // hi_r < 0 once the position has become a node
    int lo_r[kFinItems], hi_r[kFinItems];
#pragma unroll
    for (int k = 0; k < kFinItems; ++k) {
        lo_r[k] = 0;
        hi_r[k] = (kFinItems * tid + k < m) ? m : -1;
    }
    __syncthreads();
    for (int l = 0; (m >> l) >= 2; ++l) {
        // choose: one thread per segment of this level
        for (int sgm = tid; sgm < (1 << l); sgm += kFinThreads) {
            int lo, hi;
            if (!segment_range(sgm, m, l, lo, hi)) continue;
            int axis = (level0 + l) % 3;
            if (split_rule == kSplitWidest) {
                double ext[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const double first = pts[(long long)gid[cur[d * kFinSeg + lo]] * 3 + d];
                    const double last = pts[(long long)gid[cur[d * kFinSeg + hi - 1]] * 3 + d];
                    ext[d] = last - first;
                }
                axis = 0;
                if (ext[1] > ext[axis]) axis = 1;
                if (ext[2] > ext[axis]) axis = 2;
            }
            ax[lo + ((hi - lo) >> 1)] = (unsigned char)axis;
        }
        __syncthreads();
        // mark: left / median / right of every point, read off the list of its segment's split axis
#pragma unroll
        for (int k = 0; k < kFinItems; ++k) {
            if (hi_r[k] < 0) continue;
            const int p = kFinItems * tid + k, mid = lo_r[k] + ((hi_r[k] - lo_r[k]) >> 1);
            side[cur[ax[mid] * kFinSeg + p]] = p < mid ? 0 : (p == mid ? 1 : 2);
        }
        __syncthreads();
        // stable partition of the three lists inside every segment
        for (int d = 0; d < 3; ++d) {
            unsigned f[kFinItems], run = 0;
            int idv[kFinItems];
            unsigned char sdv[kFinItems];
#pragma unroll
            for (int k = 0; k < kFinItems; ++k) {
                const int p = kFinItems * tid + k;
                idv[k] = p < m ? cur[d * kFinSeg + p] : 0;
                sdv[k] = 3;
                f[k] = 0;
                if (hi_r[k] >= 0) {
                    sdv[k] = side[idv[k]];
                    f[k] = sdv[k] == 0 ? 1u : (sdv[k] == 1 ? (1u << 16) : 0u);
                }
                run += f[k];
            }
            unsigned inc = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) s_wsum[warp] = inc;
            __syncthreads();
            unsigned base = 0;
            for (int w = 0; w < warp; ++w) base += s_wsum[w];
            unsigned ex = base + inc - run;
#pragma unroll
            for (int k = 0; k < kFinItems; ++k) {
                sc[kFinItems * tid + k] = ex;
                ex += f[k];
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kFinItems; ++k) {
                const int p = kFinItems * tid + k;
                if (p >= m) continue;
                if (hi_r[k] < 0) {
                    alt[d * kFinSeg + p] = idv[k];
                    continue;
                }
                const int lo = lo_r[k], mid = lo + ((hi_r[k] - lo) >> 1);
                const unsigned rel = sc[p] - sc[lo];
                const int lefts = (int)(rel & 0xffffu), meds = (int)(rel >> 16);
                const int dst = sdv[k] == 0 ? lo + lefts : (sdv[k] == 1 ? mid : mid + 1 + ((p - lo) - lefts - meds));
                alt[d * kFinSeg + dst] = idv[k];
            }
            __syncthreads();
        }
        // descend: every position moves into the child segment that contains it
#pragma unroll
        for (int k = 0; k < kFinItems; ++k) {
            if (hi_r[k] < 0) continue;
            const int p = kFinItems * tid + k, mid = lo_r[k] + ((hi_r[k] - lo_r[k]) >> 1);
            if (p < mid)
                hi_r[k] = mid;
            else if (p == mid)
                hi_r[k] = -1;
            else
                lo_r[k] = mid + 1;
        }
        int *t2 = cur;
        cur = alt;
        alt = t2;
    }
    for (int p = tid; p < m; p += kFinThreads) {
        const int i = gid[cur[p]];
        KdNode nd;
        nd.x = pts[(long long)i * 3];
        nd.y = pts[(long long)i * 3 + 1];
        nd.z = pts[(long long)i * 3 + 2];
        nd.idx = i;
        nd.axis = ax[p];
        nodes[LO + p] = nd;
    }
}

#define KD_CHECK(call)                     \
    do {                                   \
        cudaError_t e_ = (call);           \
        if (e_ != cudaSuccess) {           \
            status = e_;                   \
            goto done;                     \
        }                                  \
    } while (0)

// bounding box of the points from the ends of the three sorted lists: bbox = {lo.xyz, hi.xyz}
__global__ void k_bbox_from_lists(const double *__restrict__ pts, const int *__restrict__ lists, int n,
                                  double *__restrict__ bbox) {
    const int a = threadIdx.x;
    if (a < 3) {
        bbox[a] = pts[(long long)lists[(long long)a * n] * 3 + a];
        bbox[3 + a] = pts[(long long)lists[(long long)a * n + n - 1] * 3 + a];
    }
}

cudaError_t kd_build(const double *d_pts, size_t n_sz, KdNode *d_nodes, double *d_bbox, int sm_count,
                     cudaStream_t stream, uint64_t *launches, int split_rule) {
    if (n_sz == 0) return cudaSuccess;
    if (n_sz > (size_t)0x7fffffff) return cudaErrorInvalidValue;
    const int n = (int)n_sz;
    cudaError_t status = cudaSuccess;
    unsigned long long *keys = nullptr, *keys_alt = nullptr, *flags = nullptr;
    int *idx0 = nullptr, *lists = nullptr, *lists_alt = nullptr, *seg_lo = nullptr, *seg_mid = nullptr;
    unsigned char *side = nullptr, *axis_at = nullptr;
    void *tmp = nullptr;
    size_t tmp_sort = 0, tmp_scan = 0, tmp_bytes = 0;
    uint64_t nl = 0;
    const int threads = 256;
    int grid = (int)((n_sz + threads - 1) / threads);
    if (grid > sm_count * 16) grid = sm_count * 16;

    KD_CHECK(cudaMallocAsync(&keys, sizeof(unsigned long long) * n_sz * 3, stream));
    KD_CHECK(cudaMallocAsync(&keys_alt, sizeof(unsigned long long) * n_sz, stream));
    KD_CHECK(cudaMallocAsync(&flags, sizeof(unsigned long long) * n_sz * 3, stream));
    KD_CHECK(cudaMallocAsync(&idx0, sizeof(int) * n_sz, stream));
    KD_CHECK(cudaMallocAsync(&lists, sizeof(int) * n_sz * 3, stream));
    KD_CHECK(cudaMallocAsync(&lists_alt, sizeof(int) * n_sz * 3, stream));
    KD_CHECK(cudaMallocAsync(&seg_lo, sizeof(int) * n_sz, stream));
    KD_CHECK(cudaMallocAsync(&seg_mid, sizeof(int) * n_sz, stream));
    KD_CHECK(cudaMallocAsync(&side, n_sz, stream));
    KD_CHECK(cudaMallocAsync(&axis_at, n_sz, stream));
    KD_CHECK(cudaMemsetAsync(axis_at, 0, n_sz, stream));
    KD_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, keys, keys_alt, idx0, lists, n, 0, 64, stream));
    {
        auto flag_in = thrust::make_transform_iterator(thrust::counting_iterator<long long>(0),
                                                       SideFlag{lists, seg_lo, side, (long long)n});
        KD_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, flag_in, flags, 3ll * n, stream));
    }
    tmp_bytes = tmp_sort > tmp_scan ? tmp_sort : tmp_scan;
    KD_CHECK(cudaMallocAsync(&tmp, tmp_bytes, stream));

    k_make_keys<<<grid, threads, 0, stream>>>(d_pts, n, keys, keys + n_sz, keys + 2 * n_sz, idx0);
    ++nl;
    for (int a = 0; a < 3; ++a) {
        KD_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tmp_sort, keys + a * n_sz, keys_alt, idx0,
                                                 lists + a * n_sz, n, 0, 64, stream));
        nl += 8;
    }
    if (d_bbox) {
        k_bbox_from_lists<<<1, 32, 0, stream>>>(d_pts, lists, n, d_bbox);
        ++nl;
    }
    {
        int *cur = lists, *alt = lists_alt;
        const dim3 grid3((unsigned)grid, 3u);  // blockIdx.y = list
        // global levels until every segment fits one CTA of the finisher
        int level0 = 0;
        while ((n >> level0) > kFinSeg) ++level0;
        for (int level = 0; level < level0; ++level) {
            long long sgrid = ((1ll << level) + threads - 1) / threads;
            if (sgrid > grid) sgrid = grid;
            k_choose_axis<<<(int)sgrid, threads, 0, stream>>>(d_pts, cur, n, level, split_rule, axis_at);
            k_mark<<<grid, threads, 0, stream>>>(cur, n, level, axis_at, seg_lo, seg_mid, side);
            auto flag_in = thrust::make_transform_iterator(thrust::counting_iterator<long long>(0),
                                                           SideFlag{cur, seg_lo, side, (long long)n});
            KD_CHECK(cub::DeviceScan::ExclusiveSum(tmp, tmp_scan, flag_in, flags, 3ll * n, stream));
            k_scatter<<<grid3, threads, 0, stream>>>(cur, alt, n, seg_lo, seg_mid, side, flags);
            nl += 5;
            int *t2 = cur;
            cur = alt;
            alt = t2;
        }
        if (level0 > 0) {
            const long long tn = 1ll << level0;
            k_emit_top<<<(unsigned)((tn + threads - 1) / threads), threads, 0, stream>>>(d_pts, cur, n, level0, axis_at,
                                                                                         d_nodes);
            ++nl;
        }
        // opt in to > 48 KB of dynamic shared memory (per device; the call is cheap enough to repeat)
        KD_CHECK(cudaFuncSetAttribute(k_kd_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFinSmemBytes));
        k_kd_finish<<<1u << level0, kFinThreads, kFinSmemBytes, stream>>>(d_pts, cur, n, level0, split_rule, seg_lo,
                                                                           d_nodes);
        ++nl;
    }
    KD_CHECK(cudaGetLastError());
done:
    cudaFreeAsync(keys, stream);
    cudaFreeAsync(keys_alt, stream);
    cudaFreeAsync(flags, stream);
    cudaFreeAsync(idx0, stream);
    cudaFreeAsync(lists, stream);
    cudaFreeAsync(lists_alt, stream);
    cudaFreeAsync(seg_lo, stream);
    cudaFreeAsync(seg_mid, stream);
    cudaFreeAsync(side, stream);
    cudaFreeAsync(axis_at, stream);
    cudaFreeAsync(tmp, stream);
    if (launches) *launches += nl;
    return status;
}

// ------------------------------------------------------------------------------ query ------
__device__ __forceinline__ void load_node(const KdNode *__restrict__ nodes, int i, double &x, double &y,
                                          double &z, int &idx, int &axis) {
    const double2 *p = reinterpret_cast<const double2 *>(nodes + i);
    const double2 a = __ldg(p);
    const double2 b = __ldg(p + 1);
    x = a.x;
    y = a.y;
    z = b.x;
    const long long w = __double_as_longlong(b.y);
    idx = (int)(w & 0xffffffffll);
    axis = (int)(w >> 32);
}

__device__ __forceinline__ double node_axis(const KdNode *__restrict__ nodes, int i, int axis) {
    return __ldg(reinterpret_cast<const double *>(nodes + i) + axis);
}

// Subtrees of at most kKdBucket nodes are a contiguous run of the in-order array: all three search
// kernels evaluate such a run point by point (independent loads, converged lanes) instead of walking it.
// Measured on B200 with the short-stack kernel, 131 072 queries: bucket 1 / 4 / 8 / 16 -> 107 / 98 / 88 / 98 us
// on 1 M uniform points, 169 / 148 / 138 / 162 us on the accumulated room map.
constexpr int kKdBucket = 8;
__device__ __forceinline__ void scan_run(const KdNode *__restrict__ nodes, int lo, int hi, double qx, double qy,
                                         double qz, double &best, int &bidx) {
    for (int i = lo; i < hi; ++i) {
        double x, y, z;
        int idx, axis;
        load_node(nodes, i, x, y, z, idx, axis);
        const double d = dsq3(dsub(x, qx), dsub(y, qy), dsub(z, qz));
        if (d < best || (d == best && idx < bidx)) {
            best = d;
            bidx = idx;
        }
    }
}

// Morton (Z-order) keys of the queries, 10 bits per axis inside the tree's bounding box: sorting
// the queries by this key makes the 32 lanes of a warp walk nearly the same root-to-leaf paths
// (coherent branches, shared cache lines).  Ordering only affects speed, never the answers.
__device__ __forceinline__ unsigned spread10(unsigned v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void k_morton_keys(const double *__restrict__ queries, long long nq, const double *__restrict__ bbox,
                              unsigned *__restrict__ keys, int *__restrict__ vals) {
    const double lo0 = bbox[0], lo1 = bbox[1], lo2 = bbox[2];
    const double s0 = 1023.0 / fmax(bbox[3] - lo0, 1e-300), s1 = 1023.0 / fmax(bbox[4] - lo1, 1e-300),
                 s2 = 1023.0 / fmax(bbox[5] - lo2, 1e-300);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nq;
         i += (long long)gridDim.x * blockDim.x) {
        const double x = (queries[i * 3] - lo0) * s0, y = (queries[i * 3 + 1] - lo1) * s1,
                     z = (queries[i * 3 + 2] - lo2) * s2;
        const unsigned ix = (unsigned)fmin(fmax(x, 0.0), 1023.0), iy = (unsigned)fmin(fmax(y, 0.0), 1023.0),
                       iz = (unsigned)fmin(fmax(z, 0.0), 1023.0);  // NaN -> 0
        keys[i] = spread10(ix) | (spread10(iy) << 1) | (spread10(iz) << 2);
        vals[i] = (int)i;
    }
}

__global__ void __launch_bounds__(128)
k_kd_nn(const KdNode *__restrict__ nodes, int n, const double *__restrict__ queries, long long nq,
        const int *__restrict__ perm, int *__restrict__ idx_out, double *__restrict__ dist_out) {
    const long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (slot >= nq) return;
    const long long qi = perm ? perm[slot] : slot;
    if (n <= 0) {  // utils/kdtree.c:112: NULL root leaves the outputs untouched; we report "none"
        idx_out[qi] = -1;
        dist_out[qi] = INFINITY;
        return;
    }
    const double qx = queries[qi * 3], qy = queries[qi * 3 + 1], qz = queries[qi * 3 + 2];
    double best = INFINITY;
    int bidx = -1;

    int lo = 0, hi = n, depth = 0;
    unsigned path = 0, par = 0;  // bit d: side taken below the depth-d ancestor (1 = right) / its size parity
    unsigned long long axes = 0;  // two bits per depth: split axis of the ancestor at that depth
    bool arriving_down = true;
    while (true) {
        const int mid = lo + ((hi - lo) >> 1);
        bool go_far = false;
        double diff;
        if (arriving_down && hi - lo <= kKdBucket) {
            scan_run(nodes, lo, hi, qx, qy, qz, best, bidx);  // then climb
        } else if (arriving_down) {
            double x, y, z;
            int idx, axis;
            load_node(nodes, mid, x, y, z, idx, axis);
            axes = (axes & ~(3ull << (2 * depth))) | ((unsigned long long)axis << (2 * depth));
            // operand order root - target (utils/kdtree.c:16); squared, so the sign is immaterial
            const double d = dsq3(dsub(x, qx), dsub(y, qy), dsub(z, qz));
            if (d < best || (d == best && idx < bidx)) {
                best = d;
                bidx = idx;
            }
            diff = dsub(axis == 0 ? qx : (axis == 1 ? qy : qz), axis == 0 ? x : (axis == 1 ? y : z));
            const bool near_right = !(diff < 0.0);  // target < node -> left, else right (kdtree.c:130-141)
            par = (par & ~(1u << depth)) | ((unsigned)((hi - lo) & 1) << depth);
            // near child
            const int clo = near_right ? mid + 1 : lo, chi = near_right ? hi : mid;
            if (clo < chi) {
                path = (path & ~(1u << depth)) | ((unsigned)near_right << depth);
                lo = clo;
                hi = chi;
                ++depth;
                continue;
            }
            // empty near child: fall through as if we had just come back from it
            go_far = true;
            path = (path & ~(1u << depth)) | ((unsigned)near_right << depth);
        } else {
            const int axis = (int)((axes >> (2 * depth)) & 3ull);
            const double key = node_axis(nodes, mid, axis);
            diff = dsub(axis == 0 ? qx : (axis == 1 ? qy : qz), key);
            const bool near_right = !(diff < 0.0);
            const bool from_right = (path >> depth) & 1u;
            go_far = from_right == near_right;  // came back from the near side
        }
        if (go_far) {
            const bool near_right = (path >> depth) & 1u;
            const int clo = near_right ? lo : mid + 1, chi = near_right ? mid : hi;
            const double plane = dmul(diff, diff);
            // the reference prunes with |delta| < best (kdtree.c:147); '<=' on squares also keeps
            // exact ties reachable so the lowest index wins; NaN planes are never pruned
            if (clo < chi && !(plane > best)) {
                path ^= (1u << depth);
                lo = clo;
                hi = chi;
                ++depth;
                arriving_down = true;
                continue;
            }
        }
        // this node is finished: climb
        if (depth == 0) break;
        --depth;
        const int cs = hi - lo;
        const unsigned parity = (par >> depth) & 1u;
        if ((path >> depth) & 1u) {  // we are the right child
            const int s = 2 * cs + 1 + (parity ? 0 : 1);
            lo = hi - s;
        } else {
            const int s = 2 * cs + (int)parity;
            hi = lo + s;
        }
        arriving_down = false;
    }
    idx_out[qi] = bidx;
    dist_out[qi] = bidx >= 0 ? __dsqrt_rn(best) : INFINITY;
}

// Converged variant of the stackless search (the default).  Every lane performs exactly one "node
// step" per loop iteration -- first visit (distance + descend) or return visit (far-side test) share
// one instruction stream -- so the warp does not alternate between a descent path and a climb path;
// lanes that finish pull the next query from a global counter instead of idling until the slowest
// lane of the warp is done; and the climb skips, with integer arithmetic only, every ancestor whose
// far side is already known to be out of reach (a per-level "pending" bit, cleared at the first
// visit when plane^2 > best).  Same nodes compared, same lexicographic rule: identical answers.
__global__ void __launch_bounds__(128)
k_kd_nn_conv(const KdNode *__restrict__ nodes, int n, const double *__restrict__ queries, long long nq,
             int *__restrict__ idx_out, double *__restrict__ dist_out, unsigned long long *__restrict__ counter) {
    const unsigned lane = threadIdx.x & 31u;
    long long qi = -1;
    bool done = false, fresh = true;
    double qx = 0, qy = 0, qz = 0, best = INFINITY;
    int bidx = -1, lo = 0, hi = 0, depth = 0;
    unsigned path = 0, par = 0, pend = 0;
    while (true) {
        // ---- refill: lanes without a query take the next ones (one atomic per warp)
        const unsigned need = __ballot_sync(0xffffffffu, qi < 0 && !done);
        if (need) {
            unsigned long long base = 0;
            const int leader = __ffs(need) - 1;
            if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(need));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (qi < 0 && !done) {
                const long long mine = (long long)base + __popc(need & ((1u << lane) - 1u));
                if (mine < nq) {
                    qi = mine;
                    qx = queries[qi * 3];
                    qy = queries[qi * 3 + 1];
                    qz = queries[qi * 3 + 2];
                    best = INFINITY;
                    bidx = -1;
                    lo = 0;
                    hi = n;
                    depth = 0;
                    path = par = pend = 0;
                    fresh = true;
                } else {
                    done = true;
                }
            }
        }
        if (__all_sync(0xffffffffu, done)) break;
        if (done) continue;
        // ---- one node step (or one small contiguous subtree evaluated point by point)
        const int mid = lo + ((hi - lo) >> 1);
        const unsigned bit = 1u << depth;
        if (fresh && hi - lo <= kKdBucket) {
            scan_run(nodes, lo, hi, qx, qy, qz, best, bidx);
            fresh = false;
        } else {
            double x, y, z;
            int idx, axis;
            load_node(nodes, mid, x, y, z, idx, axis);
            const double diff = dsub(axis == 0 ? qx : (axis == 1 ? qy : qz), axis == 0 ? x : (axis == 1 ? y : z));
            const double plane = dmul(diff, diff);
            if (fresh) {
                // operand order root - target (utils/kdtree.c:16); squared, so the sign is immaterial
                const double d = dsq3(dsub(x, qx), dsub(y, qy), dsub(z, qz));
                if (d < best || (d == best && idx < bidx)) {
                    best = d;
                    bidx = idx;
                }
                const bool near_right = !(diff < 0.0);  // target < node -> left, else right (kdtree.c:130-141)
                par = (par & ~bit) | ((unsigned)((hi - lo) & 1) << depth);
                path = (path & ~bit) | ((unsigned)near_right << depth);
                const bool far_nonempty = near_right ? (lo < mid) : (mid + 1 < hi);
                // '<=' on squares keeps exact ties reachable (lowest index wins); NaN planes are never pruned
                pend = (far_nonempty && !(plane > best)) ? (pend | bit) : (pend & ~bit);
                const int clo = near_right ? mid + 1 : lo, chi = near_right ? hi : mid;
                if (clo < chi) {
                    lo = clo;
                    hi = chi;
                    ++depth;
                    continue;
                }
                fresh = false;  // empty near side: handle this node as a return visit right away
            }
            if ((pend & bit) && !(plane > best)) {  // far side still within reach: go there
                const bool near_right = (path >> depth) & 1u;
                pend &= ~bit;
                path ^= bit;
                if (near_right)
                    hi = mid;
                else
                    lo = mid + 1;
                ++depth;
                fresh = true;
                continue;
            }
        }
        pend &= ~bit;
        const unsigned below = pend & (bit - 1u);
        if (!below) {  // nothing pending anywhere above: this query is finished
            idx_out[qi] = bidx;
            dist_out[qi] = bidx >= 0 ? __dsqrt_rn(best) : INFINITY;
            qi = -1;
            continue;
        }
        const int target = 31 - __clz(below);
        while (depth > target) {  // integer-only climb; ranges follow from (side taken, size parity)
            --depth;
            const int cs = hi - lo;
            const unsigned parity = (par >> depth) & 1u;
            if ((path >> depth) & 1u)
                lo = hi - (2 * cs + 1 + (parity ? 0 : 1));
            else
                hi = lo + (2 * cs + (int)parity);
        }
    }
}

// The same search with a short explicit stack (thread-local memory, L1 resident): far children are
// pushed with their plane distance on the way down and popped (or discarded) later, so no ancestor
// is ever re-read and nothing climbs level by level.  Visits exactly the nodes the stackless kernel
// visits (same pruning rule, same lexicographic compare) -> identical answers.
// (Tried on top of this: the incremental per-axis cell bound of Arya & Mount, offsets stacked as
// floats rounded toward zero.  Exact, but 10-15 % slower on every map measured -- the plane test
// already rejects nearly everything the stronger bound would; profiles/README.md.)
__global__ void __launch_bounds__(128)
k_kd_nn_stack(const KdNode *__restrict__ nodes, int n, const double *__restrict__ queries, long long nq,
              const int *__restrict__ perm, int *__restrict__ idx_out, double *__restrict__ dist_out) {
    const long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (slot >= nq) return;
    const long long qi = perm ? perm[slot] : slot;
    if (n <= 0) {
        idx_out[qi] = -1;
        dist_out[qi] = INFINITY;
        return;
    }
    const double qx = queries[qi * 3], qy = queries[qi * 3 + 1], qz = queries[qi * 3 + 2];
    double best = INFINITY;
    int bidx = -1;
    int st_lo[32], st_hi[32];
    double st_plane[32];
    int sp = 0;
    int lo = 0, hi = n;
    while (true) {
        while (lo < hi) {
            if (hi - lo <= kKdBucket) {
                scan_run(nodes, lo, hi, qx, qy, qz, best, bidx);
                break;
            }
            const int mid = lo + ((hi - lo) >> 1);
            double x, y, z;
            int idx, axis;
            load_node(nodes, mid, x, y, z, idx, axis);
            const double d = dsq3(dsub(x, qx), dsub(y, qy), dsub(z, qz));
            if (d < best || (d == best && idx < bidx)) {
                best = d;
                bidx = idx;
            }
            const double diff = dsub(axis == 0 ? qx : (axis == 1 ? qy : qz), axis == 0 ? x : (axis == 1 ? y : z));
            const bool near_right = !(diff < 0.0);
            const int flo = near_right ? lo : mid + 1, fhi = near_right ? mid : hi;
            if (flo < fhi) {
                st_lo[sp] = flo;
                st_hi[sp] = fhi;
                st_plane[sp] = dmul(diff, diff);
                ++sp;
            }
            if (near_right)
                lo = mid + 1;
            else
                hi = mid;
        }
        // next pending far subtree whose plane is still within reach (NaN planes are never pruned)
        bool found = false;
        while (sp > 0) {
            --sp;
            if (!(st_plane[sp] > best)) {
                lo = st_lo[sp];
                hi = st_hi[sp];
                found = true;
                break;
            }
        }
        if (!found) break;
    }
    idx_out[qi] = bidx;
    dist_out[qi] = bidx >= 0 ? __dsqrt_rn(best) : INFINITY;
}

cudaError_t kd_nn(const KdNode *d_nodes, size_t n, const double *d_bbox, const double *d_queries, size_t nq,
                  int *d_idx, double *d_dist, int sm_count, cudaStream_t stream, uint64_t *launches,
                  unsigned long long *d_counter) {
    if (nq == 0) return cudaSuccess;
    const int threads = 128;
    const unsigned grid = (unsigned)((nq + threads - 1) / threads);
    // Measured on B200 (profiles/README.md): Morton-ordering the queries speeds the traversal itself up
    // by only 13 % (1 M points, 131 072 queries: 127.6 -> 111.2 us) while key generation + radix sort
    // cost ~75 us, so it is off unless NAV_KD_SORT=1 (e.g. for much larger query sets).
    static const int sort_mode = getenv("NAV_KD_SORT") ? atoi(getenv("NAV_KD_SORT")) : 0;
    const bool sort_queries = sort_mode && d_bbox && n >= 4096 && nq >= 8192 && nq < (size_t)0x7fffffff;
    // Kernel choice by measurement on B200 (131 072 queries, widest-extent trees, 8-node runs scanned;
    // profiles/README.md), short stack / plain stackless / converged stackless:
    //   1 M uniform points 80 / 96 / 108 us, 4 M 124 / 126 / 144 us, 10 M 164 / 178 / 171 us,
    //   accumulated room map (1 M points on surfaces) 122 / 162 / 185 us.
    // The short stack wins or ties everywhere, so it is the default; NAV_KD_KERNEL = plain|conv|stack
    // selects the others (same answers).
    static const char *kk = getenv("NAV_KD_KERNEL");  // read once
    const bool use_stack = kk ? !strcmp(kk, "stack") : true;
    const bool use_conv = kk && !strcmp(kk, "conv");
    if (!sort_queries && use_conv && d_counter && n > 0) {
        cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream);
        unsigned pgrid = (unsigned)sm_count * 12u;  // 12 x 128 threads = 48 warps per SM (40 registers)
        if ((unsigned long long)pgrid * threads > nq) pgrid = (unsigned)((nq + threads - 1) / threads);
        k_kd_nn_conv<<<pgrid, threads, 0, stream>>>(d_nodes, (int)n, d_queries, (long long)nq, d_idx, d_dist,
                                                    d_counter);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    if (!sort_queries) {
        if (use_stack)
            k_kd_nn_stack<<<grid, threads, 0, stream>>>(d_nodes, (int)n, d_queries, (long long)nq, nullptr, d_idx,
                                                        d_dist);
        else
            k_kd_nn<<<grid, threads, 0, stream>>>(d_nodes, (int)n, d_queries, (long long)nq, nullptr, d_idx, d_dist);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    cudaError_t status = cudaSuccess;
    unsigned *keys = nullptr, *keys_out = nullptr;
    int *vals = nullptr, *perm = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    KD_CHECK(cudaMallocAsync(&keys, nq * 4, stream));
    KD_CHECK(cudaMallocAsync(&keys_out, nq * 4, stream));
    KD_CHECK(cudaMallocAsync(&vals, nq * 4, stream));
    KD_CHECK(cudaMallocAsync(&perm, nq * 4, stream));
    KD_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, vals, perm, (int)nq, 0, 30, stream));
    KD_CHECK(cudaMallocAsync(&tmp, tmp_bytes, stream));
    {
        int kgrid = (int)((nq + 255) / 256);
        if (kgrid > sm_count * 8) kgrid = sm_count * 8;
        k_morton_keys<<<kgrid, 256, 0, stream>>>(d_queries, (long long)nq, d_bbox, keys, vals);
    }
    KD_CHECK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, vals, perm, (int)nq, 0, 30, stream));
    k_kd_nn<<<grid, threads, 0, stream>>>(d_nodes, (int)n, d_queries, (long long)nq, perm, d_idx, d_dist);
    if (launches) *launches += 6;
    KD_CHECK(cudaGetLastError());
done:
    cudaFreeAsync(keys, stream);
    cudaFreeAsync(keys_out, stream);
    cudaFreeAsync(vals, stream);
    cudaFreeAsync(perm, stream);
    cudaFreeAsync(tmp, stream);
    return status;
}

// --------------------------------------------------------------- exact fp64 brute force ------
constexpr int kBfTile = 1024;
__global__ void __launch_bounds__(256)
k_bf_nn(const double *__restrict__ pts, long long n, const double *__restrict__ queries, long long nq,
        int *__restrict__ idx_out, double *__restrict__ dist_out) {
    __shared__ double s_p[kBfTile * 3];
    const long long qi = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool live = qi < nq;
    double qx = 0, qy = 0, qz = 0;
    if (live) {
        qx = queries[qi * 3];
        qy = queries[qi * 3 + 1];
        qz = queries[qi * 3 + 2];
    }
    double best = INFINITY;
    int bidx = -1;
    for (long long t0 = 0; t0 < n; t0 += kBfTile) {
        const int cnt = (int)min((long long)kBfTile, n - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt * 3; i += blockDim.x) s_p[i] = pts[t0 * 3 + i];
        __syncthreads();
        if (live) {
            for (int j = 0; j < cnt; ++j) {
                const double d = dsq3(dsub(s_p[j * 3], qx), dsub(s_p[j * 3 + 1], qy), dsub(s_p[j * 3 + 2], qz));
                if (d < best) {  // ascending index scan: strict '<' keeps the lowest index
                    best = d;
                    bidx = (int)(t0 + j);
                }
            }
        }
    }
    if (live) {
        idx_out[qi] = bidx;
        dist_out[qi] = bidx >= 0 ? __dsqrt_rn(best) : INFINITY;
    }
}

cudaError_t bf_nn(const double *d_pts, size_t n, const double *d_queries, size_t nq, int *d_idx, double *d_dist,
                  cudaStream_t stream) {
    if (nq == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((nq + 255) / 256);
    k_bf_nn<<<grid, 256, 0, stream>>>(d_pts, (long long)n, d_queries, (long long)nq, d_idx, d_dist);
    return cudaGetLastError();
}

}  // namespace nav
