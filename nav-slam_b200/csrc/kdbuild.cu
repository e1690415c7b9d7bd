// kdbuild.cu -- level-by-level build of the flat kd-tree for large maps (a5), by radix SELECT of the
// median and one partition pass per level.  No sort of the whole input, no library kernels.
//
// Reference: utils/kdtree.c:65-82 builds a pointer tree by recursive median split -- axis = depth % 3,
// median index m = n/2 found by an in-place Lomuto quick-select (utils/kdtree.c:20-62), children on
// [0,m) and [m+1,n) -- one malloc per node.  The tree here is the in-order array that recursion leaves
// behind: the node of range [lo,hi) sits at mid = lo + (hi-lo)/2, its children own [lo,mid) and
// [mid+1,hi).  Which point becomes the node is defined by a total order, (key on the split axis, index in
// the build input): the median is the element of rank (hi-lo)/2 in that order.  With distinct keys and the
// cyclic axis rule this is the reference's tree node for node; with equal keys the reference's choice is an
// accident of its quick-select and ours is the lowest-index rule.  Because the node of every
// range is a function of the SET of points in the range, the order in which a partition pass leaves the
// points inside the two halves does not matter: the result is deterministic although the passes place
// elements with atomic counters.
//
// Working set: structure of arrays, per point three order-preserving 64-bit keys (one per axis; the
// coordinate's bits are recoverable from the key, so the node array is written from the keys alone) and
// the 32-bit index: 28 B/point, two buffers (ping-pong).  Per level, all segments at once:
//
//   large segments (> kMidCap points), three launches per level:
//     k_kd_hist      every CTA takes <= 4096 points of ONE segment: split axis from the segment's bounding box
//                    (widest extent, or depth % 3), 2048-bin histogram of  (key - box.lo) >> shift  in shared
//                    memory, added to the segment's global histogram               (reads  8 B/point)
//     k_kd_partition the bin holding the median rank follows from the histogram; points in lower bins go to
//                    the left part, in higher bins to the right part, points of that one bin ("candidates",
//                    about size/2048 of them) in between; places are reserved with one atomic per CTA and
//                    category; the bounding boxes of the two halves are reduced on the way
//                                                                                  (reads 28 + writes 28 B/point)
//     k_kd_resolve   one CTA per segment orders the candidates exactly (shared-memory bitonic sort on
//                    (key, index)), which puts the median at mid, and writes the node
//   middle segments (<= kMidCap): k_kd_level_cta, ONE launch per level, one CTA per segment doing the
//                    same three steps with shared-memory histogram and counters
//   small segments (<= 2048): k_kd_finish, ONE launch: a CTA loads its segment into shared memory, sorts
//                    three index lists (one per axis) once, and builds the whole subtree below it by
//                    stable partitions of the lists.
// A bin that holds more candidates than fit in shared memory (thousands of equal or nearly equal keys:
// walls in integer-millimetre data) is narrowed by further histogram rounds over the candidates -- first on
// the key bits below `shift`, then on the index -- inside the resolving CTA; slower, same result.
//
// Algorithmic traffic: lower bound 24 B read + 32 B written per point; this scheme moves
// 24 + 28 (keys) + levels x 64 + 28 + 32 B per point, all of it coalesced.
#include <stdlib.h>

#include "nav_kdtree.cuh"

namespace nav {

typedef unsigned long long u64;

constexpr int kBins = 2048;          // histogram bins of one selection round (11 bits)
constexpr int kBinBits = 11;
constexpr int kCap = 2048;           // candidates that are ordered in shared memory
constexpr int kFinSeg = 2048;        // segments of at most this many points are finished by one CTA
constexpr int kChunkElems = 4096;    // points per CTA of the large-segment kernels
constexpr int kPartThreads = 256;
constexpr int kPartItems = kChunkElems / kPartThreads;
constexpr int kCtaThreads = 512;     // resolve / per-segment kernels
constexpr unsigned kFullMask = 0xffffffffu;

struct KdSoA {
    u64 *k[3];
    int *id;
};
struct SegBox {   // bounding box of a segment in key space
    u64 lo[3], hi[3];
};
struct SegPick {  // outcome of the selection rounds of a large segment: the candidates' place and key range
    int below, count, pad0, pad1;
    u64 klo, khi;
};
struct SegSplit {
    int axis, shift;
    u64 base;
};

__device__ __forceinline__ u64 order_key(double v) {
    const u64 b = (u64)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_value(u64 k) {
    const u64 b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
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

// smallest shift with (range >> shift) < kBins
__device__ __forceinline__ int bin_shift(u64 range) {
    const int bits = 64 - __clzll((long long)range);
    return bits > kBinBits ? bits - kBinBits : 0;
}

__device__ __forceinline__ SegSplit split_of(const SegBox &b, int level, int rule) {
    int axis = level % 3;  // utils/kdtree.c:72
    if (rule == kSplitWidest) {
        double e[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) e[d] = key_value(b.hi[d]) - key_value(b.lo[d]);
        axis = 0;  // ties and NaN extents keep the lowest axis
        if (e[1] > e[axis]) axis = 1;
        if (e[2] > e[axis]) axis = 2;
    }
    SegSplit s;
    s.axis = axis;
    s.base = b.lo[axis];
    s.shift = bin_shift(b.hi[axis] - b.lo[axis]);
    return s;
}

// key range [klo, khi] of bin b of a split (khi clipped to the box: klo + 2^shift - 1 may not fit in 64 bits)
__device__ __forceinline__ void bin_range(const SegSplit &sp, u64 box_hi, int b, u64 &klo, u64 &khi) {
    klo = sp.base + ((u64)b << sp.shift);
    const u64 span = (1ull << sp.shift) - 1ull;
    khi = (box_hi - klo > span) ? klo + span : box_hi;
}

__device__ __forceinline__ u64 umin64(u64 a, u64 b) { return a < b ? a : b; }
__device__ __forceinline__ u64 umax64(u64 a, u64 b) { return a > b ? a : b; }

// running bounding box of one thread
struct BoxAcc {
    u64 lo[3], hi[3];
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            lo[d] = ~0ull;
            hi[d] = 0ull;
        }
    }
    __device__ __forceinline__ void add(u64 a, u64 b, u64 c) {
        lo[0] = umin64(lo[0], a);
        hi[0] = umax64(hi[0], a);
        lo[1] = umin64(lo[1], b);
        hi[1] = umax64(hi[1], b);
        lo[2] = umin64(lo[2], c);
        hi[2] = umax64(hi[2], c);
    }
    __device__ __forceinline__ void warp_reduce() {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                lo[d] = umin64(lo[d], __shfl_xor_sync(kFullMask, lo[d], o));
                hi[d] = umax64(hi[d], __shfl_xor_sync(kFullMask, hi[d], o));
            }
        }
    }
};

// Merges the running boxes of all threads of the CTA into s_box[0] (from a) and s_box[1] (from b): warp
// reduction, one row of twelve values per warp in shared memory, then twelve threads fold the rows.  (64-bit
// min / max atomics on shared memory are compare-and-swap loops: sixteen warps contending for the same
// twelve words cost microseconds.)  Called by all T threads; contains two barriers.
template <int T>
__device__ __forceinline__ void commit_boxes(BoxAcc &a, BoxAcc &b, SegBox *s_box, u64 (*s_part)[12]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a.warp_reduce();
    b.warp_reduce();
    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            s_part[warp][d] = a.lo[d];
            s_part[warp][3 + d] = a.hi[d];
            s_part[warp][6 + d] = b.lo[d];
            s_part[warp][9 + d] = b.hi[d];
        }
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        const int v = threadIdx.x, d = v % 3;
        const bool is_min = (v % 6) < 3;
        SegBox &dst = s_box[v / 6];
        u64 acc = is_min ? dst.lo[d] : dst.hi[d];
        for (int w = 0; w < T / 32; ++w) acc = is_min ? umin64(acc, s_part[w][v]) : umax64(acc, s_part[w][v]);
        if (is_min)
            dst.lo[d] = acc;
        else
            dst.hi[d] = acc;
    }
    __syncthreads();
}

__device__ __forceinline__ void box_reset(SegBox *b) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        b->lo[d] = ~0ull;
        b->hi[d] = 0ull;
    }
}

// ------------------------------------------------------------------------------------------------
// keys + root bounding box
__global__ void __launch_bounds__(256)
k_kd_keys(const double *__restrict__ pts, int n, KdSoA out, SegBox *__restrict__ box0) {
    __shared__ SegBox s_box[2];
    __shared__ u64 s_part[256 / 32][12];
    if (threadIdx.x < 2) box_reset(&s_box[threadIdx.x]);
    __syncthreads();
    BoxAcc acc, none;
    acc.reset();
    none.reset();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u64 a = order_key(pts[(long long)i * 3]), b = order_key(pts[(long long)i * 3 + 1]),
                  c = order_key(pts[(long long)i * 3 + 2]);
        out.k[0][i] = a;
        out.k[1][i] = b;
        out.k[2][i] = c;
        out.id[i] = i;
        acc.add(a, b, c);
    }
    commit_boxes<256>(acc, none, s_box, s_part);
    if (threadIdx.x < 3) {
        atomicMin(&box0->lo[threadIdx.x], s_box[0].lo[threadIdx.x]);
        atomicMax(&box0->hi[threadIdx.x], s_box[0].hi[threadIdx.x]);
    }
}

// box[0] = empty box; histogram row 0 and the counters of segment 0 = 0
__global__ void k_kd_init(SegBox *box0, unsigned *ghist, unsigned *ghist2, unsigned *fill) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) box_reset(box0);
    if (t < kBins) {
        ghist[t] = 0u;
        ghist2[t] = 0u;
    }
    if (t < 4) fill[t] = 0u;
}

// ------------------------------------------------------------------------------------------------
// block-wide search of the bin that holds rank `target` in hist[kBins] (shared or global memory):
// bin, number of elements in lower bins, number in the bin.  All threads get the result.  T = blockDim.x.
template <int T>
__device__ __forceinline__ void block_pick(const unsigned *hist, int target, int *s_res, unsigned *s_wsum, int &bin,
                                           int &below, int &count) {
    constexpr int kPer = kBins / T;
    static_assert(kBins % T == 0 && T % 32 == 0, "bins are split evenly over the threads");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned loc[kPer], sum = 0;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        loc[q] = hist[tid * kPer + q];
        sum += loc[q];
    }
    unsigned inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(kFullMask, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    unsigned basew = 0;
    for (int w = 0; w < warp; ++w) basew += s_wsum[w];
    const unsigned ex = basew + inc - sum;
    if (sum != 0u && ex <= (unsigned)target && (unsigned)target < ex + sum) {
        unsigned run = ex;
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            if ((unsigned)target < run + loc[q]) {
                s_res[0] = tid * kPer + q;
                s_res[1] = (int)run;
                s_res[2] = (int)loc[q];
                break;
            }
            run += loc[q];
        }
    }
    __syncthreads();
    bin = s_res[0];
    below = s_res[1];
    count = s_res[2];
    __syncthreads();  // s_res / s_wsum may be reused right away
}

// ------------------------------------------------------------------------------------------------
// shared memory of the per-segment kernels
struct KdSmem {
    u64 K[3][kCap];
    int ID[kCap];
    unsigned hist[kBins];
    unsigned hist2[kBins];
    unsigned short perm[kCap];
    unsigned short tmp[kCap];
    SegBox box[2];        // accumulated bounding boxes of the left / right half
    u64 part[kCtaThreads / 32][12];  // per-warp rows of commit_boxes
    unsigned wsum[32];
    int res[4];
    unsigned cnt[4];
    u64 piv_key[3];
    int piv_id;
};

// composite order (key, index); list entries >= c are padding and sort last
__device__ __forceinline__ bool perm_greater(unsigned a, unsigned b, int c, const u64 *key, const int *id) {
    if (a >= (unsigned)c) return b < (unsigned)c;
    if (b >= (unsigned)c) return false;
    const u64 ka = key[a], kb = key[b];
    return ka > kb || (ka == kb && id[a] > id[b]);
}

// bitonic argsort of perm[0..P) (P a power of two >= c), T threads, ends with a barrier
template <int T>
__device__ __forceinline__ void bitonic_argsort(unsigned short *perm, int P, int c, const u64 *key, const int *id) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += T) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i + j;
                const unsigned a = perm[i], b = perm[l];
                const bool up = (i & k) == 0;
                if (perm_greater(a, b, c, key, id) == up) {
                    perm[i] = (unsigned short)b;
                    perm[l] = (unsigned short)a;
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int pow2_at_least(int c) {
    int P = 2;
    while (P < c) P <<= 1;
    return P;
}

// Argsort of c <= kBins elements by (key, index) in O(c): counting sort over kBins bins of the key range
// [klo, khi] -- with about as many bins as elements a bin holds one or two of them on ordinary data -- and a
// brute-force ranking inside each bin.  Lidar maps are not ordinary at every scale: a thin slab of wall points
// plus a few far outliers puts hundreds of points into a dozen bins.  Then a SECOND level subdivides every bin
// of the first histogram into as many equal key intervals as it holds points (so there are again about c
// buckets in total, and the slab is resolved at its own scale) and the points are counted once more.
// perm[0..c) receives the order; tmp is a scratch list, hist / hist2 kBins words each, wsum 1 + T/32 words.
// Returns false, leaving perm undefined, when a bucket still holds more than kBucketMax elements (equal or
// nearly equal keys): the caller then takes the bitonic network.  Called by all T threads; the result is
// CTA-uniform; ends with a barrier.
constexpr int kBucketMax = 48;

// exclusive scan of h[0..kBins) in place (thread t owns kBins / T consecutive entries) and the largest entry.
// wsum: 1 + T/32 words, wsum[0] receives the maximum.  Called by all T threads, ends with a barrier.
template <int T>
__device__ __forceinline__ void scan_bins(unsigned *h, unsigned *wsum) {
    constexpr int kPer = kBins / T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) wsum[0] = 0u;
    unsigned cnt[kPer], run = 0, big = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        cnt[k] = h[kPer * tid + k];
        run += cnt[k];
        big = max(big, cnt[k]);
    }
    unsigned inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(kFullMask, inc, o);
        if (lane >= o) inc += t;
    }
    big = __reduce_max_sync(kFullMask, big);
    __syncthreads();  // wsum[0] is reset
    if (lane == 31) wsum[1 + warp] = inc;
    if (lane == 0) atomicMax(&wsum[0], big);
    __syncthreads();
    const unsigned before = lane < warp ? wsum[1 + lane] : 0u;  // T / 32 <= 32 warps
    unsigned ex = __reduce_add_sync(kFullMask, before) + inc - run;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        h[kPer * tid + k] = ex;
        ex += cnt[k];
    }
    __syncthreads();
}

template <int T>
__device__ __forceinline__ bool counting_argsort(unsigned short *perm, unsigned short *tmp, unsigned *hist, unsigned *hist2,
                                                 unsigned *wsum, int c, const u64 *key, const int *id, u64 klo, u64 khi) {
    constexpr int kPer = kBins / T;
    const int tid = threadIdx.x;
    const int shift = bin_shift(khi - klo);
    for (int i = tid; i < kBins; i += T) hist[i] = 0u;
    __syncthreads();
    unsigned bin[kPer], slot[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int i = tid + k * T;
        bin[k] = 0u;
        slot[k] = 0u;
        if (i < c) {
            bin[k] = (unsigned)((key[i] - klo) >> shift);
            slot[k] = atomicAdd(&hist[bin[k]], 1u);
        }
    }
    __syncthreads();
    scan_bins<T>(hist, wsum);
    const unsigned *start = hist;
    if (wsum[0] > (unsigned)kBucketMax) {  // CTA-uniform
        // second level: bin b with n_b points becomes n_b buckets start[b] .. start[b] + n_b - 1
        for (int i = tid; i < kBins; i += T) hist2[i] = 0u;
        __syncthreads();
        const u64 low_mask = shift ? ((1ull << shift) - 1ull) : 0ull;
        const int pre = shift > 52 ? shift - 52 : 0;  // keeps (low >> pre) * n_b below 2^63
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int i = tid + k * T;
            if (i < c) {
                const unsigned b = bin[k];
                const unsigned bs = hist[b], nb = (b + 1 < (unsigned)kBins ? hist[b + 1] : (unsigned)c) - bs;
                const u64 low = (key[i] - klo) & low_mask;
                unsigned sub = (unsigned)(((low >> pre) * (u64)nb) >> (shift - pre));
                sub = min(sub, nb - 1u);
                bin[k] = bs + sub;
                slot[k] = atomicAdd(&hist2[bin[k]], 1u);
            }
        }
        __syncthreads();
        scan_bins<T>(hist2, wsum);
        start = hist2;
        if (wsum[0] > (unsigned)kBucketMax) {
            __syncthreads();  // everybody has read the flag before a following call can reset it
            return false;
        }
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int i = tid + k * T;
        if (i < c) tmp[start[bin[k]] + slot[k]] = (unsigned short)i;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int i = tid + k * T;
        if (i >= c) continue;
        const int bs = (int)start[bin[k]], be = bin[k] + 1 < (unsigned)kBins ? (int)start[bin[k] + 1] : c;
        const u64 mine = key[i];
        const int my_id = id[i];
        int rank = 0;
        for (int j = bs; j < be; ++j) {
            const int o = tmp[j];
            const u64 ko = key[o];
            rank += (ko < mine || (ko == mine && id[o] < my_id)) ? 1 : 0;
        }
        perm[bs + rank] = (unsigned short)i;
    }
    __syncthreads();
    return true;
}

// Orders the candidate region out[a, a+c) so that the element of rank r (composite order on `axis`) sits at
// a + r, everything smaller in front of it and everything larger behind it; accumulates the keys of the
// elements in front / behind into sm.box[0] / sm.box[1]; leaves the pivot in sm.piv_*.  klo..khi bound the
// candidates' keys on `axis`.  `scratch` is a buffer whose range [a, a+c) is free.  Called by all T threads.
template <int T>
__device__ void resolve_region(KdSmem &sm, const KdSoA &out, const KdSoA &scratch, int a, int c, int r, int axis,
                               u64 klo, u64 khi) {
    const int tid = threadIdx.x, lane = tid & 31;
    BoxAcc accL, accR;
    accL.reset();
    accR.reset();
    if (c <= kCap) {
        for (int t = tid; t < c; t += T) {
            sm.K[0][t] = out.k[0][a + t];
            sm.K[1][t] = out.k[1][a + t];
            sm.K[2][t] = out.k[2][a + t];
            sm.ID[t] = out.id[a + t];
        }
        __syncthreads();
        if (!counting_argsort<T>(sm.perm, sm.tmp, sm.hist, sm.hist2, sm.wsum, c, sm.K[axis], sm.ID, klo, khi)) {
            const int P = pow2_at_least(c);
            for (int t = tid; t < P; t += T) sm.perm[t] = t < c ? (unsigned short)t : (unsigned short)0xffffu;
            __syncthreads();
            bitonic_argsort<T>(sm.perm, P, c, sm.K[axis], sm.ID);
        }
        for (int t = tid; t < c; t += T) {
            const int src = sm.perm[t];
            const u64 k0 = sm.K[0][src], k1 = sm.K[1][src], k2 = sm.K[2][src];
            out.k[0][a + t] = k0;
            out.k[1][a + t] = k1;
            out.k[2][a + t] = k2;
            out.id[a + t] = sm.ID[src];
            if (t < r)
                accL.add(k0, k1, k2);
            else if (t > r)
                accR.add(k0, k1, k2);
            else {
                sm.piv_key[0] = k0;
                sm.piv_key[1] = k1;
                sm.piv_key[2] = k2;
                sm.piv_id = sm.ID[src];
            }
        }
        commit_boxes<T>(accL, accR, sm.box, sm.part);
        return;
    }
    // ---- more candidates than fit: narrow the key range (then the index range) with further histogram
    // rounds over the region until the members of the range fit, order those, and take the pivot
    const u64 *ka = out.k[axis];
    int rr = r, cnt = c, ilo = 0, ihi = 0x7fffffff;
    bool by_id = false;
    while (cnt > kCap) {
        for (int t = tid; t < kBins; t += T) sm.hist[t] = 0u;
        __syncthreads();
        if (!by_id && klo == khi) by_id = true;  // all remaining candidates share one key: go on with the index
        int sh;
        if (!by_id) {
            sh = bin_shift(khi - klo);
            for (int t = tid; t < c; t += T) {
                const u64 k = ka[a + t];
                if (k >= klo && k <= khi) atomicAdd(&sm.hist[(unsigned)((k - klo) >> sh)], 1u);
            }
        } else {
            sh = bin_shift((u64)(unsigned)(ihi - ilo));
            for (int t = tid; t < c; t += T) {
                const int id = out.id[a + t];
                if (ka[a + t] == klo && id >= ilo && id <= ihi) atomicAdd(&sm.hist[(unsigned)(id - ilo) >> sh], 1u);
            }
        }
        __syncthreads();
        int b2, below2, cnt2;
        block_pick<T>(sm.hist, rr, sm.res, sm.wsum, b2, below2, cnt2);
        rr -= below2;
        cnt = cnt2;
        if (!by_id) {
            klo += (u64)b2 << sh;
            const u64 span = (1ull << sh) - 1ull;
            if (khi - klo > span) khi = klo + span;
        } else {
            ilo += b2 << sh;
            const int span = (1 << sh) - 1;
            if (ihi - ilo > span) ihi = ilo + span;
        }
    }
    if (tid == 0) sm.cnt[0] = 0u;
    __syncthreads();
    for (int t = tid; t < c; t += T) {
        const u64 k = ka[a + t];
        const int id = out.id[a + t];
        const bool in = by_id ? (k == klo && id >= ilo && id <= ihi) : (k >= klo && k <= khi);
        if (in) {
            const unsigned slot = atomicAdd(&sm.cnt[0], 1u);
            sm.K[0][slot] = k;
            sm.ID[slot] = id;
        }
    }
    __syncthreads();
    {
        const int m2 = (int)sm.cnt[0];  // == cnt
        const int P = pow2_at_least(m2);
        for (int t = tid; t < P; t += T) sm.perm[t] = t < m2 ? (unsigned short)t : (unsigned short)0xffffu;
        __syncthreads();
        bitonic_argsort<T>(sm.perm, P, m2, sm.K[0], sm.ID);
    }
    const u64 pk = sm.K[0][sm.perm[rr]];
    const int pid = sm.ID[sm.perm[rr]];
    __syncthreads();
    // ---- partition the region around the pivot into the scratch buffer, then bring it back
    if (tid < 4) sm.cnt[tid] = 0u;
    __syncthreads();
    const int mid = a + r;
    for (int t0 = 0; t0 < c; t0 += T) {
        const int t = t0 + tid;
        int cat = 3;
        u64 k0 = 0, k1 = 0, k2 = 0;
        int id = 0;
        if (t < c) {
            k0 = out.k[0][a + t];
            k1 = out.k[1][a + t];
            k2 = out.k[2][a + t];
            id = out.id[a + t];
            const u64 k = axis == 0 ? k0 : (axis == 1 ? k1 : k2);
            cat = (k < pk || (k == pk && id < pid)) ? 0 : ((k == pk && id == pid) ? 1 : 2);
        }
        const unsigned bl = __ballot_sync(kFullMask, cat == 0), bg = __ballot_sync(kFullMask, cat == 2);
        unsigned basel = 0, baseg = 0;
        if (lane == 0) {
            basel = atomicAdd(&sm.cnt[0], (unsigned)__popc(bl));
            baseg = atomicAdd(&sm.cnt[2], (unsigned)__popc(bg));
        }
        basel = __shfl_sync(kFullMask, basel, 0);
        baseg = __shfl_sync(kFullMask, baseg, 0);
        const unsigned lt = (1u << lane) - 1u;
        int pos = -1;
        if (cat == 0)
            pos = a + (int)basel + __popc(bl & lt);
        else if (cat == 1)
            pos = mid;
        else if (cat == 2)
            pos = mid + 1 + (int)baseg + __popc(bg & lt);
        if (pos >= 0) {
            scratch.k[0][pos] = k0;
            scratch.k[1][pos] = k1;
            scratch.k[2][pos] = k2;
            scratch.id[pos] = id;
        }
    }
    __syncthreads();
    for (int t = tid; t < c; t += T) {
        const u64 k0 = scratch.k[0][a + t], k1 = scratch.k[1][a + t], k2 = scratch.k[2][a + t];
        const int id = scratch.id[a + t];
        out.k[0][a + t] = k0;
        out.k[1][a + t] = k1;
        out.k[2][a + t] = k2;
        out.id[a + t] = id;
        if (t < r)
            accL.add(k0, k1, k2);
        else if (t > r)
            accR.add(k0, k1, k2);
        else {
            sm.piv_key[0] = k0;
            sm.piv_key[1] = k1;
            sm.piv_key[2] = k2;
            sm.piv_id = id;
        }
    }
    commit_boxes<T>(accL, accR, sm.box, sm.part);
}

__device__ __forceinline__ void emit_node(KdNode *nodes, int pos, const u64 key[3], int id, int axis) {
    KdNode nd;
    nd.x = key_value(key[0]);
    nd.y = key_value(key[1]);
    nd.z = key_value(key[2]);
    nd.idx = id;
    nd.axis = axis;
    nodes[pos] = nd;
}

// ------------------------------------------------------------------------------------------------
// Second selection round.  When the bin of the first round that holds the median is crowded (more than kCap
// points: a thin slab of a wall or the floor, many points within a millimetre) the points of THAT bin are
// binned again, by the next kBinBits key bits below `shift` -- or, if the bin is a single key value
// (shift == 0), by the leading bits of the point index, the tie-breaker of the order.  Both are monotone in
// (key, index), so "lower sub-bin" still means "smaller".
struct Refine {
    int on;        // 0: the first round's bin is the candidate set
    int shift2;    // key refinement: sub-bin = (key - klo1) >> shift2
    int by_id;     // index refinement: sub-bin = id >> id_shift
    int id_shift;
    u64 klo1;      // lowest key of the first round's bin
};
__device__ __forceinline__ Refine refine_of(const SegSplit &sp, int b1, int cnt1, int n) {
    Refine r;
    r.on = cnt1 > kCap;
    r.klo1 = sp.base + ((u64)b1 << sp.shift);
    r.by_id = sp.shift == 0;
    r.shift2 = sp.shift > kBinBits ? sp.shift - kBinBits : 0;
    const int id_bits = 32 - __clz(max(n - 1, 1));
    r.id_shift = id_bits > kBinBits ? id_bits - kBinBits : 0;
    return r;
}
__device__ __forceinline__ unsigned sub_bin(const Refine &r, u64 key, int id) {
    return r.by_id ? (unsigned)id >> r.id_shift : (unsigned)((key - r.klo1) >> r.shift2);
}

// large segments, step 1: histogram of the split-axis keys.  grid = n_seg * cps CTAs.
__global__ void __launch_bounds__(kPartThreads)
k_kd_hist(KdSoA cur, int n, int level, int rule, int cps, const SegBox *__restrict__ box_cur,
          SegBox *__restrict__ box_next, unsigned *__restrict__ ghist) {
    __shared__ unsigned s_hist[kBins];
    const int s = blockIdx.x / cps, j = blockIdx.x % cps;
    int lo, hi;
    if (!segment_range(s, n, level, lo, hi)) return;
    const int start = lo + j * kChunkElems;
    if (start >= hi) return;
    const int end = min(hi, start + kChunkElems);
    const SegSplit sp = split_of(box_cur[s], level, rule);
    if (j == 0 && threadIdx.x == 0) {  // the halves' boxes are accumulated by the next two kernels
        box_reset(&box_next[2 * s]);
        box_reset(&box_next[2 * s + 1]);
    }
    for (int t = threadIdx.x; t < kBins; t += kPartThreads) s_hist[t] = 0u;
    __syncthreads();
    const u64 *ka = cur.k[sp.axis];
    {   // all of the thread's keys are loaded before the first shared-memory atomic (the loop with the atomic
        // inside was bound by one load latency per iteration)
        u64 kv[kPartItems];
#pragma unroll
        for (int it = 0; it < kPartItems; ++it) {
            const int e = start + it * kPartThreads + threadIdx.x;
            kv[it] = e < end ? ka[e] : 0ull;
        }
#pragma unroll
        for (int it = 0; it < kPartItems; ++it) {
            const int e = start + it * kPartThreads + threadIdx.x;
            if (e < end) atomicAdd(&s_hist[(unsigned)((kv[it] - sp.base) >> sp.shift)], 1u);
        }
    }
    __syncthreads();
    unsigned *g = ghist + (size_t)s * kBins;
    for (int t = threadIdx.x; t < kBins; t += kPartThreads)
        if (s_hist[t]) atomicAdd(&g[t], s_hist[t]);
}

// large segments, step 1b: histogram of the second round over the points of a crowded first-round bin.
// Same grid as k_kd_hist; CTAs of segments whose bin is not crowded leave at once.
__global__ void __launch_bounds__(kPartThreads)
k_kd_hist2(KdSoA cur, int n, int level, int rule, int cps, const SegBox *__restrict__ box_cur,
           const unsigned *__restrict__ ghist, unsigned *__restrict__ ghist2) {
    __shared__ unsigned s_hist[kBins];
    __shared__ int s_res[4];
    __shared__ unsigned s_wsum[32];
    const int s = blockIdx.x / cps, j = blockIdx.x % cps;
    int lo, hi;
    if (!segment_range(s, n, level, lo, hi)) return;
    const int start = lo + j * kChunkElems;
    if (start >= hi) return;
    const int end = min(hi, start + kChunkElems);
    const SegSplit sp = split_of(box_cur[s], level, rule);
    int b1, below1, cnt1;
    block_pick<kPartThreads>(ghist + (size_t)s * kBins, (hi - lo) >> 1, s_res, s_wsum, b1, below1, cnt1);
    const Refine rf = refine_of(sp, b1, cnt1, n);
    if (!rf.on) return;
    for (int t = threadIdx.x; t < kBins; t += kPartThreads) s_hist[t] = 0u;
    __syncthreads();
    const u64 *ka = cur.k[sp.axis];
    u64 kv[kPartItems];
#pragma unroll
    for (int it = 0; it < kPartItems; ++it) {
        const int e = start + it * kPartThreads + threadIdx.x;
        kv[it] = e < end ? ka[e] : 0ull;
    }
#pragma unroll
    for (int it = 0; it < kPartItems; ++it) {
        const int e = start + it * kPartThreads + threadIdx.x;
        if (e < end && (int)((kv[it] - sp.base) >> sp.shift) == b1)
            atomicAdd(&s_hist[sub_bin(rf, kv[it], rf.by_id ? cur.id[e] : 0)], 1u);
    }
    __syncthreads();
    unsigned *g = ghist2 + (size_t)s * kBins;
    for (int t = threadIdx.x; t < kBins; t += kPartThreads)
        if (s_hist[t]) atomicAdd(&g[t], s_hist[t]);
}

// large segments, step 2: lower bins | candidates | higher bins.  Same grid as k_kd_hist.
__global__ void __launch_bounds__(kPartThreads)
k_kd_partition(KdSoA cur, KdSoA nxt, int n, int level, int rule, int cps, const SegBox *__restrict__ box_cur,
               SegBox *__restrict__ box_next, const unsigned *__restrict__ ghist, const unsigned *__restrict__ ghist2,
               unsigned *__restrict__ fill, SegPick *__restrict__ pick) {
    __shared__ int s_res[4];
    __shared__ unsigned s_wsum[32];
    __shared__ unsigned s_w[kPartThreads / 32][3];
    __shared__ unsigned s_base[3];
    __shared__ SegBox s_box[2];
    __shared__ u64 s_part[kPartThreads / 32][12];
    const int s = blockIdx.x / cps, j = blockIdx.x % cps;
    int lo, hi;
    if (!segment_range(s, n, level, lo, hi)) return;
    const int start = lo + j * kChunkElems;
    if (start >= hi) return;
    const int end = min(hi, start + kChunkElems);
    const SegBox my_box = box_cur[s];
    const SegSplit sp = split_of(my_box, level, rule);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 2) box_reset(&s_box[tid]);
    int b, below, cnt;
    block_pick<kPartThreads>(ghist + (size_t)s * kBins, (hi - lo) >> 1, s_res, s_wsum, b, below, cnt);
    u64 klo, khi;
    bin_range(sp, my_box.hi[sp.axis], b, klo, khi);
    // crowded bin: the second round (k_kd_hist2) has binned its points again
    const Refine rf = refine_of(sp, b, cnt, n);
    int b2 = 0;
    if (rf.on) {
        int below2, cnt2;
        block_pick<kPartThreads>(ghist2 + (size_t)s * kBins, ((hi - lo) >> 1) - below, s_res, s_wsum, b2, below2, cnt2);
        below += below2;
        cnt = cnt2;
        if (!rf.by_id) {   // key range of the sub-bin (clipped to the first round's bin)
            const u64 lo2 = rf.klo1 + ((u64)b2 << rf.shift2), span = (1ull << rf.shift2) - 1ull;
            khi = (khi - lo2 > span) ? lo2 + span : khi;
            klo = lo2;
        }
    }
    if (j == 0 && tid == 0) {
        SegPick p = {below, cnt, 0, 0, klo, khi};
        pick[s] = p;
    }
    // classify this CTA's points (2 bits each, kept in a register) and count the three categories
    const u64 *ka = cur.k[sp.axis];
    unsigned cats = 0;
    int nl = 0, nc = 0, nr = 0;
#pragma unroll
    for (int it = 0; it < kPartItems; ++it) {
        const int e = start + it * kPartThreads + tid;
        unsigned cat = 3u;
        if (e < end) {
            const u64 key = ka[e];
            const int bin = (int)((key - sp.base) >> sp.shift);
            cat = bin < b ? 0u : (bin > b ? 2u : 1u);
            if (rf.on && cat == 1u) {
                const int sb = (int)sub_bin(rf, key, rf.by_id ? cur.id[e] : 0);
                cat = sb < b2 ? 0u : (sb > b2 ? 2u : 1u);
            }
        }
        cats |= cat << (2 * it);
        nl += cat == 0u;
        nc += cat == 1u;
        nr += cat == 2u;
    }
    nl = __reduce_add_sync(kFullMask, nl);
    nc = __reduce_add_sync(kFullMask, nc);
    nr = __reduce_add_sync(kFullMask, nr);
    if (lane == 0) {
        s_w[warp][0] = (unsigned)nl;
        s_w[warp][1] = (unsigned)nc;
        s_w[warp][2] = (unsigned)nr;
    }
    __syncthreads();
    if (tid < 3) {  // CTA totals -> one reservation per category; warp totals -> exclusive prefixes
        unsigned run = 0;
        for (int w = 0; w < kPartThreads / 32; ++w) {
            const unsigned v = s_w[w][tid];
            s_w[w][tid] = run;
            run += v;
        }
        s_base[tid] = run ? atomicAdd(&fill[(size_t)s * 4 + tid], run) : 0u;
    }
    __syncthreads();
    int run_l = lo + (int)(s_base[0] + s_w[warp][0]);
    int run_c = lo + below + (int)(s_base[1] + s_w[warp][1]);
    int run_r = lo + below + cnt + (int)(s_base[2] + s_w[warp][2]);
    BoxAcc accL, accR;
    accL.reset();
    accR.reset();
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < kPartItems; ++it) {
        const int e = start + it * kPartThreads + tid;
        const unsigned cat = (cats >> (2 * it)) & 3u;
        const unsigned bl = __ballot_sync(kFullMask, cat == 0u), bc = __ballot_sync(kFullMask, cat == 1u),
                       br = __ballot_sync(kFullMask, cat == 2u);
        if (cat != 3u) {
            const u64 k0 = cur.k[0][e], k1 = cur.k[1][e], k2 = cur.k[2][e];
            const int id = cur.id[e];
            int pos;
            if (cat == 0u) {
                pos = run_l + __popc(bl & lt);
                accL.add(k0, k1, k2);
            } else if (cat == 1u) {
                pos = run_c + __popc(bc & lt);
            } else {
                pos = run_r + __popc(br & lt);
                accR.add(k0, k1, k2);
            }
            nxt.k[0][pos] = k0;
            nxt.k[1][pos] = k1;
            nxt.k[2][pos] = k2;
            nxt.id[pos] = id;
        }
        run_l += __popc(bl);
        run_c += __popc(bc);
        run_r += __popc(br);
    }
    commit_boxes<kPartThreads>(accL, accR, s_box, s_part);
    if (tid < 6) {
        const int side = tid / 3, d = tid % 3;
        SegBox *dst = &box_next[2 * s + side];
        if (s_box[side].lo[d] != ~0ull) atomicMin(&dst->lo[d], s_box[side].lo[d]);
        if (s_box[side].hi[d] != 0ull) atomicMax(&dst->hi[d], s_box[side].hi[d]);
    }
}

// large segments, step 3: one CTA per segment orders the candidates, writes the node, and clears the
// histogram rows / counters the two halves will use on the next level
__global__ void __launch_bounds__(kCtaThreads)
k_kd_resolve(KdSoA out, KdSoA scratch, int n, int level, int rule, const SegBox *__restrict__ box_cur,
             SegBox *__restrict__ box_next, const SegPick *__restrict__ pick, unsigned *__restrict__ ghist,
             unsigned *__restrict__ ghist2, unsigned *__restrict__ fill, int clear_next, KdNode *__restrict__ nodes,
             double *__restrict__ bbox_out) {
    extern __shared__ __align__(16) unsigned char kd_smem_raw[];
    KdSmem &sm = *reinterpret_cast<KdSmem *>(kd_smem_raw);
    const int s = blockIdx.x, tid = threadIdx.x;
    if (clear_next) {
        unsigned *g = ghist + (size_t)(2 * s) * kBins, *g2 = ghist2 + (size_t)(2 * s) * kBins;
        for (int t = tid; t < 2 * kBins; t += kCtaThreads) {
            g[t] = 0u;
            g2[t] = 0u;
        }
    }
    if (tid < 8) fill[(size_t)(2 * s) * 4 + tid] = 0u;
    int lo, hi;
    if (!segment_range(s, n, level, lo, hi)) return;
    const SegSplit sp = split_of(box_cur[s], level, rule);
    if (level == 0 && bbox_out && tid < 3) {
        bbox_out[tid] = key_value(box_cur[0].lo[tid]);
        bbox_out[3 + tid] = key_value(box_cur[0].hi[tid]);
    }
    const SegPick p = pick[s];
    if (tid < 2) box_reset(&sm.box[tid]);
    __syncthreads();
    const int m = (hi - lo) >> 1;
    resolve_region<kCtaThreads>(sm, out, scratch, lo + p.below, p.count, m - p.below, sp.axis, p.klo, p.khi);
    if (tid == 0) emit_node(nodes, lo + m, sm.piv_key, sm.piv_id, sp.axis);
    if (tid < 6) {
        const int side = tid / 3, d = tid % 3;
        SegBox *dst = &box_next[2 * s + side];
        if (sm.box[side].lo[d] != ~0ull) atomicMin(&dst->lo[d], sm.box[side].lo[d]);
        if (sm.box[side].hi[d] != 0ull) atomicMax(&dst->hi[d], sm.box[side].hi[d]);
    }
}

// middle segments: the whole level of one segment in one CTA
__global__ void __launch_bounds__(kCtaThreads)
k_kd_level_cta(KdSoA cur, KdSoA nxt, int n, int level, int rule, const SegBox *__restrict__ box_cur,
               SegBox *__restrict__ box_next, KdNode *__restrict__ nodes, double *__restrict__ bbox_out) {
    extern __shared__ __align__(16) unsigned char kd_smem_raw[];
    KdSmem &sm = *reinterpret_cast<KdSmem *>(kd_smem_raw);
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    int lo, hi;
    if (!segment_range(s, n, level, lo, hi)) return;
    const SegSplit sp = split_of(box_cur[s], level, rule);
    if (level == 0 && bbox_out && tid < 3) {
        bbox_out[tid] = key_value(box_cur[0].lo[tid]);
        bbox_out[3 + tid] = key_value(box_cur[0].hi[tid]);
    }
    for (int t = tid; t < kBins; t += kCtaThreads) sm.hist[t] = 0u;
    if (tid < 2) box_reset(&sm.box[tid]);
    if (tid < 4) sm.cnt[tid] = 0u;
    __syncthreads();
    const u64 *ka = cur.k[sp.axis];
    for (int e0 = lo; e0 < hi; e0 += 8 * kCtaThreads) {  // eight loads in flight per thread in front of the atomics
        u64 kv[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int e = e0 + it * kCtaThreads + tid;
            kv[it] = e < hi ? ka[e] : 0ull;
        }
#pragma unroll
        for (int it = 0; it < 8; ++it)
            if (e0 + it * kCtaThreads + tid < hi) atomicAdd(&sm.hist[(unsigned)((kv[it] - sp.base) >> sp.shift)], 1u);
    }
    __syncthreads();
    const int m = (hi - lo) >> 1;
    int b, below, cnt;
    block_pick<kCtaThreads>(sm.hist, m, sm.res, sm.wsum, b, below, cnt);
    BoxAcc accL, accR;
    accL.reset();
    accR.reset();
    const unsigned lt = (1u << lane) - 1u;
    for (int e0 = lo; e0 < hi; e0 += kCtaThreads) {
        const int e = e0 + tid;
        unsigned cat = 3u;
        u64 k0 = 0, k1 = 0, k2 = 0;
        int id = 0;
        if (e < hi) {
            k0 = cur.k[0][e];
            k1 = cur.k[1][e];
            k2 = cur.k[2][e];
            id = cur.id[e];
            const u64 k = sp.axis == 0 ? k0 : (sp.axis == 1 ? k1 : k2);
            const int bin = (int)((k - sp.base) >> sp.shift);
            cat = bin < b ? 0u : (bin > b ? 2u : 1u);
        }
        const unsigned bl = __ballot_sync(kFullMask, cat == 0u), bc = __ballot_sync(kFullMask, cat == 1u),
                       br = __ballot_sync(kFullMask, cat == 2u);
        unsigned o0 = 0, o1 = 0, o2 = 0;
        if (lane == 0) {
            if (bl) o0 = atomicAdd(&sm.cnt[0], (unsigned)__popc(bl));
            if (bc) o1 = atomicAdd(&sm.cnt[1], (unsigned)__popc(bc));
            if (br) o2 = atomicAdd(&sm.cnt[2], (unsigned)__popc(br));
        }
        o0 = __shfl_sync(kFullMask, o0, 0);
        o1 = __shfl_sync(kFullMask, o1, 0);
        o2 = __shfl_sync(kFullMask, o2, 0);
        if (cat != 3u) {
            int pos;
            if (cat == 0u) {
                pos = lo + (int)o0 + __popc(bl & lt);
                accL.add(k0, k1, k2);
            } else if (cat == 1u) {
                pos = lo + below + (int)o1 + __popc(bc & lt);
            } else {
                pos = lo + below + cnt + (int)o2 + __popc(br & lt);
                accR.add(k0, k1, k2);
            }
            nxt.k[0][pos] = k0;
            nxt.k[1][pos] = k1;
            nxt.k[2][pos] = k2;
            nxt.id[pos] = id;
        }
    }
    commit_boxes<kCtaThreads>(accL, accR, sm.box, sm.part);  // its barriers also publish the candidates written above
    u64 klo, khi;
    bin_range(sp, box_cur[s].hi[sp.axis], b, klo, khi);
    resolve_region<kCtaThreads>(sm, nxt, cur, lo + below, cnt, m - below, sp.axis, klo, khi);
    if (tid == 0) emit_node(nodes, lo + m, sm.piv_key, sm.piv_id, sp.axis);
    if (tid < 2) box_next[2 * s + tid] = sm.box[tid];
}

// ------------------------------------------------------------------------------------------------
// finisher: one CTA builds the whole subtree of a segment of <= kFinSeg points in shared memory.
// Three index lists, each sorted once along one axis (bitonic, all three in the same stage loop), then per
// level: every sub-segment picks its axis (widest extent read off the ends of its three lists, or
// depth % 3), the middle element of that axis' list is the median, and all three lists are stably
// partitioned inside every sub-segment (prefix sum of packed left / median counts + scatter), which keeps
// them sorted for the levels below.
constexpr int kFinThreads = 512;
static_assert(kFinSeg == kBins, "the finisher's counting sort uses one bin per slot of the segment");
constexpr int kFinItems = kFinSeg / kFinThreads;
struct FinSmem {
    u64 K[3][kFinSeg];
    int gid[kFinSeg];
    unsigned short lst[2][3][kFinSeg];
    unsigned sc[kFinSeg];           // exclusive prefix, lefts | medians << 16
    unsigned sc2[kFinSeg];          // second-level bins of the counting sort
    unsigned char ax[kFinSeg];      // split axis of the node at a position
    unsigned char side[kFinSeg];    // left / median / right code per local id
    unsigned wsum[1 + kFinThreads / 32];
};

__global__ void __launch_bounds__(kFinThreads)
k_kd_finish(KdSoA in, int n, int level0, int rule, KdNode *__restrict__ nodes, double *__restrict__ bbox_out,
            const SegBox *__restrict__ box0) {
    extern __shared__ __align__(16) unsigned char kd_smem_raw[];
    FinSmem &sm = *reinterpret_cast<FinSmem *>(kd_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (level0 == 0 && blockIdx.x == 0 && bbox_out && tid < 3) {
        bbox_out[tid] = key_value(box0->lo[tid]);
        bbox_out[3 + tid] = key_value(box0->hi[tid]);
    }
    int LO, HI;
    if (!segment_range((int)blockIdx.x, n, level0, LO, HI)) return;
    const int m = HI - LO;
    const int P = pow2_at_least(m);
    for (int i = tid; i < kFinSeg; i += kFinThreads) {
        if (i < m) {
            sm.K[0][i] = in.k[0][LO + i];
            sm.K[1][i] = in.k[1][LO + i];
            sm.K[2][i] = in.k[2][LO + i];
            sm.gid[i] = in.id[LO + i];
        }
        const unsigned short v = i < m ? (unsigned short)i : (unsigned short)0xffffu;
        sm.lst[0][0][i] = v;
        sm.lst[0][1][i] = v;
        sm.lst[0][2][i] = v;
        sm.ax[i] = 0;
    }
    __syncthreads();
    // Sort the three index lists, one per axis, by (key, index).  Counting sort: with as many bins as there
    // are points (kFinSeg bins over the segment's key range on that axis) a bin holds one or two points on
    // ordinary data, so one histogram, one scan and a brute-force ranking inside each bin order the list in
    // O(m) -- the bitonic network (66 stages of 3 x 1024 compare-exchanges for 2048 points) was 60 % of this
    // kernel.  An axis with a crowded bin (equal or nearly equal keys: walls in integer data) takes the
    // bitonic network instead.
    const SegBox sbox = box0[blockIdx.x];
    for (int d = 0; d < 3; ++d) {
        unsigned short *perm = sm.lst[0][d];
        if (!counting_argsort<kFinThreads>(perm, sm.lst[1][d], sm.sc, sm.sc2, sm.wsum, m, sm.K[d], sm.gid, sbox.lo[d], sbox.hi[d])) {
            for (int i = tid; i < kFinSeg; i += kFinThreads) perm[i] = i < m ? (unsigned short)i : (unsigned short)0xffffu;
            __syncthreads();
            bitonic_argsort<kFinThreads>(perm, P, m, sm.K[d], sm.gid);
        }
    }
    // this thread owns positions kFinItems*tid .. +kFinItems-1 (relative to LO); lo_r/hi_r: their current
    // sub-segment, hi_r < 0 once the position has become a node
    int lo_r[kFinItems], hi_r[kFinItems];
#pragma unroll
    for (int k = 0; k < kFinItems; ++k) {
        lo_r[k] = 0;
        hi_r[k] = (kFinItems * tid + k < m) ? m : -1;
    }
    int cur = 0;
    for (int l = 0; (m >> l) >= 2; ++l) {
        unsigned short(*L)[kFinSeg] = sm.lst[cur];
        unsigned short(*A)[kFinSeg] = sm.lst[cur ^ 1];
        // choose: one thread per sub-segment of this level
        for (int sgm = tid; sgm < (1 << l); sgm += kFinThreads) {
            int lo, hi;
            if (!segment_range(sgm, m, l, lo, hi)) continue;
            int axis = (level0 + l) % 3;
            if (rule == kSplitWidest) {
                double ext[3];
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    ext[d] = key_value(sm.K[d][L[d][hi - 1]]) - key_value(sm.K[d][L[d][lo]]);
                axis = 0;
                if (ext[1] > ext[axis]) axis = 1;
                if (ext[2] > ext[axis]) axis = 2;
            }
            sm.ax[lo + ((hi - lo) >> 1)] = (unsigned char)axis;
        }
        __syncthreads();
        // mark: left / median / right of every point, read off the list of its sub-segment's split axis
#pragma unroll
        for (int k = 0; k < kFinItems; ++k) {
            if (hi_r[k] < 0) continue;
            const int p = kFinItems * tid + k, mid = lo_r[k] + ((hi_r[k] - lo_r[k]) >> 1);
            sm.side[L[sm.ax[mid]][p]] = p < mid ? 0 : (p == mid ? 1 : 2);
        }
        __syncthreads();
        // stable partition of the three lists inside every sub-segment
        for (int d = 0; d < 3; ++d) {
            unsigned f[kFinItems], run = 0;
            int idv[kFinItems];
            unsigned char sdv[kFinItems];
#pragma unroll
            for (int k = 0; k < kFinItems; ++k) {
                const int p = kFinItems * tid + k;
                idv[k] = p < m ? L[d][p] : 0;
                sdv[k] = 3;
                f[k] = 0;
                if (hi_r[k] >= 0) {
                    sdv[k] = sm.side[idv[k]];
                    f[k] = sdv[k] == 0 ? 1u : (sdv[k] == 1 ? (1u << 16) : 0u);
                }
                run += f[k];
            }
            unsigned inc = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(kFullMask, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) sm.wsum[warp] = inc;
            __syncthreads();
            const unsigned before = lane < warp ? sm.wsum[lane] : 0u;
            unsigned ex = __reduce_add_sync(kFullMask, before) + inc - run;
#pragma unroll
            for (int k = 0; k < kFinItems; ++k) {
                sm.sc[kFinItems * tid + k] = ex;
                ex += f[k];
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kFinItems; ++k) {
                const int p = kFinItems * tid + k;
                if (p >= m) continue;
                if (hi_r[k] < 0) {
                    A[d][p] = (unsigned short)idv[k];
                    continue;
                }
                const int lo = lo_r[k], mid = lo + ((hi_r[k] - lo) >> 1);
                const unsigned rel = sm.sc[p] - sm.sc[lo];
                const int lefts = (int)(rel & 0xffffu), meds = (int)(rel >> 16);
                const int dst = sdv[k] == 0 ? lo + lefts : (sdv[k] == 1 ? mid : mid + 1 + ((p - lo) - lefts - meds));
                A[d][dst] = (unsigned short)idv[k];
            }
            __syncthreads();
        }
        // descend: every position moves into the child sub-segment that contains it
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
        cur ^= 1;
    }
    for (int p = tid; p < m; p += kFinThreads) {
        const int i = sm.lst[cur][0][p];
        KdNode nd;
        nd.x = key_value(sm.K[0][i]);
        nd.y = key_value(sm.K[1][i]);
        nd.z = key_value(sm.K[2][i]);
        nd.idx = sm.gid[i];
        nd.axis = sm.ax[p];
        nodes[LO + p] = nd;
    }
}

// ------------------------------------------------------------------------------------------------
// Workspace comes from a pool of this library's own (the device's default pool and its release threshold
// are left alone): freed blocks stay cached, so a build per frame does not map and unmap memory.
cudaError_t kd_pool_alloc(void **p, size_t bytes, int device, cudaStream_t stream) {
    static cudaMemPool_t pools[64] = {};
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    if (!pools[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaError_t e = cudaMemPoolCreate(&pools[device], &props);
        if (e != cudaSuccess) return e;
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pools[device], cudaMemPoolAttrReleaseThreshold, &thr);
    }
    return cudaMallocFromPoolAsync(p, bytes ? bytes : 1, pools[device], stream);
}

#define KD_CHECK(call)                     \
    do {                                   \
        cudaError_t e_ = (call);           \
        if (e_ != cudaSuccess) {           \
            status = e_;                   \
            goto done;                     \
        }                                  \
    } while (0)

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

cudaError_t kd_build(const double *d_pts, size_t n_sz, KdNode *d_nodes, double *d_bbox, int sm_count,
                     cudaStream_t stream, uint64_t *launches, int split_rule) {
    if (n_sz == 0) return cudaSuccess;
    if (n_sz > (size_t)0x7fffffff) return cudaErrorInvalidValue;
    const int n = (int)n_sz;
    cudaError_t status = cudaSuccess;
    int device = 0;
    cudaGetDevice(&device);
    // segments above this size take the three multi-CTA kernels, at or below it the one-CTA-per-segment
    // kernel (measured on B200, 1 M / 10 M points: see profiles/README.md)
    static const int mid_cap = env_int("NAV_KD_MIDCAP", 16384);
    static bool configured[64] = {};
    if (device >= 0 && device < 64 && !configured[device]) {
        cudaError_t e = cudaFuncSetAttribute(k_kd_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KdSmem));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(k_kd_level_cta, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KdSmem));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(k_kd_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FinSmem));
        if (e != cudaSuccess) return e;
        configured[device] = true;
    }
    // levels handled before the finisher; the largest segment of level L holds floor(n / 2^L) points
    int level0 = 0;
    while ((n >> level0) > kFinSeg) ++level0;
    int top_levels = 0;
    while (top_levels < level0 && (n >> top_levels) > mid_cap) ++top_levels;
    const size_t seg_cap = (size_t)1 << level0;        // boxes of the finisher's level are written, never read
    const size_t hist_rows = (size_t)1 << top_levels;  // rows 2s, 2s+1 are cleared for the level below the last top level
    u64 *keys = nullptr;
    int *ids = nullptr;
    SegBox *box = nullptr;
    unsigned *ghist = nullptr, *fill = nullptr;
    SegPick *pick = nullptr;
    uint64_t nl = 0;
    KD_CHECK(kd_pool_alloc((void **)&keys, sizeof(u64) * n_sz * 6, device, stream));
    KD_CHECK(kd_pool_alloc((void **)&ids, sizeof(int) * n_sz * 2, device, stream));
    KD_CHECK(kd_pool_alloc((void **)&box, sizeof(SegBox) * seg_cap * 2 * 2, device, stream));
    KD_CHECK(kd_pool_alloc((void **)&ghist, sizeof(unsigned) * kBins * hist_rows * 2 * 2, device, stream));  // both rounds
    KD_CHECK(kd_pool_alloc((void **)&fill, sizeof(unsigned) * 4 * hist_rows * 2 + 64, device, stream));
    KD_CHECK(kd_pool_alloc((void **)&pick, sizeof(SegPick) * hist_rows, device, stream));
    {
        KdSoA buf[2];
        for (int b = 0; b < 2; ++b) {
            for (int d = 0; d < 3; ++d) buf[b].k[d] = keys + ((size_t)b * 3 + d) * n_sz;
            buf[b].id = ids + (size_t)b * n_sz;
        }
        SegBox *boxes[2] = {box, box + seg_cap * 2};
        int cur = 0;
        unsigned *ghist2 = ghist + (size_t)kBins * hist_rows * 2;
        k_kd_init<<<(kBins + 255) / 256, 256, 0, stream>>>(boxes[0], ghist, ghist2, fill);
        int grid = (n + 255) / 256;
        if (grid > sm_count * 8) grid = sm_count * 8;
        k_kd_keys<<<grid, 256, 0, stream>>>(d_pts, n, buf[0], boxes[0]);
        nl += 2;
        for (int level = 0; level < level0; ++level) {
            const int n_seg = 1 << level;
            const int max_size = n >> level;
            if (level < top_levels) {
                const int cps = (max_size + kChunkElems - 1) / kChunkElems;
                const unsigned g = (unsigned)n_seg * (unsigned)cps;
                k_kd_hist<<<g, kPartThreads, 0, stream>>>(buf[cur], n, level, split_rule, cps, boxes[cur], boxes[cur ^ 1],
                                                        ghist);
                k_kd_hist2<<<g, kPartThreads, 0, stream>>>(buf[cur], n, level, split_rule, cps, boxes[cur], ghist, ghist2);
                k_kd_partition<<<g, kPartThreads, 0, stream>>>(buf[cur], buf[cur ^ 1], n, level, split_rule, cps,
                                                             boxes[cur], boxes[cur ^ 1], ghist, ghist2, fill, pick);
                k_kd_resolve<<<n_seg, kCtaThreads, sizeof(KdSmem), stream>>>(
                    buf[cur ^ 1], buf[cur], n, level, split_rule, boxes[cur], boxes[cur ^ 1], pick, ghist, ghist2, fill,
                    level + 1 < top_levels ? 1 : 0, d_nodes, d_bbox);
                nl += 4;
            } else {
                k_kd_level_cta<<<n_seg, kCtaThreads, sizeof(KdSmem), stream>>>(buf[cur], buf[cur ^ 1], n, level, split_rule,
                                                                              boxes[cur], boxes[cur ^ 1], d_nodes, d_bbox);
                nl += 1;
            }
            cur ^= 1;
        }
        k_kd_finish<<<1u << level0, kFinThreads, sizeof(FinSmem), stream>>>(buf[cur], n, level0, split_rule, d_nodes, d_bbox,
                                                                          boxes[cur]);
        nl += 1;
    }
    KD_CHECK(cudaGetLastError());
done:
    if (keys) cudaFreeAsync(keys, stream);
    if (ids) cudaFreeAsync(ids, stream);
    if (box) cudaFreeAsync(box, stream);
    if (ghist) cudaFreeAsync(ghist, stream);
    if (fill) cudaFreeAsync(fill, stream);
    if (pick) cudaFreeAsync(pick, stream);
    if (launches) *launches += nl;
    return status;
}

}  // namespace nav
