// nav_kdtree.cuh -- flat kd-tree interfaces (kdtree.cu)
#pragma once
#include "nav_common.cuh"

namespace nav {

struct __align__(32) KdNode {
    double x, y, z;
    int idx;   // index of the point in the build input
    int axis;  // split axis of this node (0, 1, 2)
};
static_assert(sizeof(KdNode) == 32, "one DRAM sector per node");

enum { kSplitCyclic = 0, kSplitWidest = 1 };  // = NAV_KD_SPLIT_* of include/navslam_b200.h
cudaError_t kd_build(const double *d_pts, size_t n, KdNode *d_nodes, double *d_bbox, int sm_count,
                     cudaStream_t stream, uint64_t *launches, int split_rule);
// stream-ordered allocation from the library's own memory pool of `device` (freed with cudaFreeAsync)
cudaError_t kd_pool_alloc(void **p, size_t bytes, int device, cudaStream_t stream);
// Delivery of a query shard's answers into the result buffers of all ranks of a node (peer memory).
constexpr int kMaxPeers = 8;
struct KdFanOut {
    int *idx[kMaxPeers];                  // rank r's idx array   [nq_total]
    double *dist[kMaxPeers];              // rank r's dist array  [nq_total]
    unsigned long long *flags[kMaxPeers]; // rank r's flag slots  [world]
    unsigned *ticket;                     // local counter of finished CTAs (zero between launches)
    long long q_lo;                       // first query of this rank's shard in the full arrays
    unsigned long long seq;               // call number posted to the flag slots
    int world, rank;
    // point-sharded map (mode 2): every rank searches ITS part of the map for ALL queries and sends each partial
    // answer (squared distance, global point index) to the rank that owns the query, slot [sender][query - owner.lo]
    double *pdsq[kMaxPeers];              // rank r's partial squared distances [world][shard_cap]
    int *pidx[kMaxPeers];                 // rank r's partial indices           [world][shard_cap]
    long long nq_total;
    int shard_cap, idx_offset;
};
// the contiguous shard [lo, hi) of rank r of n items (first n % world ranks get one more), nav-slam_b200/sharding.py
__host__ __device__ inline void shard_range(long long n, int world, int r, long long &lo, long long &hi) {
    const long long base = n / world, extra = n % world;
    lo = r * base + (r < extra ? r : extra);
    hi = lo + base + (r < extra ? 1 : 0);
}
__host__ __device__ inline int shard_owner(long long n, int world, long long i) {
    const long long base = n / world, extra = n % world;
    if (i < (base + 1) * extra) return (int)(i / (base + 1));
    return (int)(extra + (i - (base + 1) * extra) / (base > 0 ? base : 1));
}
cudaError_t kd_nn_partial(const KdNode *d_nodes, size_t n, const double *d_queries, size_t nq, const KdFanOut &fan,
                          cudaStream_t stream, uint64_t *launches);
cudaError_t peer_merge(const KdFanOut &fan, cudaStream_t stream);
cudaError_t kd_nn_fanout(const KdNode *d_nodes, size_t n, const double *d_queries, size_t nq, const KdFanOut &fan,
                         cudaStream_t stream, uint64_t *launches);
cudaError_t peer_wait(const unsigned long long *flags, int world, unsigned long long seq, unsigned *err, cudaStream_t stream);
cudaError_t kd_nn(const KdNode *d_nodes, size_t n, const double *d_bbox, const double *d_queries, size_t nq,
                  int *d_idx, double *d_dist, int sm_count, cudaStream_t stream, uint64_t *launches,
                  unsigned long long *d_counter);  // d_counter: 8 bytes of device scratch (work queue head)
cudaError_t bf_nn(const double *d_pts, size_t n, const double *d_queries, size_t nq, int *d_idx, double *d_dist,
                  cudaStream_t stream);

// tensor-core candidate tiles + exact re-rank (bf_tc.cu); *h_evals_out (optional, forces a stream sync)
// receives the number of exact distance evaluations the re-rank needed
cudaError_t bf_nn_tc(const double *d_pts, size_t n, const double *d_queries, size_t nq, int *d_idx, double *d_dist,
                     int sm_count, cudaStream_t stream, unsigned long long *h_evals_out);

}  // namespace nav
