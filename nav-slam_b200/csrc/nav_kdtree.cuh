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
};
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
