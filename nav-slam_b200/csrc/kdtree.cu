// kdtree.cu -- batched exact 1-NN (a6) on the flat implicit kd-tree for large maps, and the exact
// binary64 brute force.
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
// The build lives in kdbuild.cu (radix select of the median + one partition pass per level).  The split
// axis is stored in each node, so the search does not care which rule built the tree.
//
// Query: one thread per query.  Child ranges are pure arithmetic on (lo,hi).  Three interchangeable
// kernels (identical answers, chosen by measurement in kd_nn()): k_kd_nn_stack keeps pending far
// subtrees on a 32-entry thread-local stack (the default); k_kd_nn / k_kd_nn_conv are stackless -- the
// way back up is recovered from two bit masks (which side was taken, parity of each ancestor's size)
// and only the split coordinate of an ancestor is re-read (L1/L2 hits).  Subtrees of <= 8 nodes are
// scanned as a contiguous run.  Far subtrees are visited iff the rounded plane distance^2 is <= the
// current best dsq; candidates compare lexicographically on (dsq, original index): exact NN, lowest
// index on ties.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "nav_kdtree.cuh"

namespace nav {

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
// the same run with all (<= kKdBucket) node loads issued before the first distance is formed: one memory
// latency per run instead of one per node (the loop below waits for every node in turn)
__device__ __forceinline__ void scan_run_batched(const KdNode *__restrict__ nodes, int lo, int hi, double qx, double qy,
                                                 double qz, double &best, int &bidx) {
    double2 a[kKdBucket], b[kKdBucket];
#pragma unroll
    for (int k = 0; k < kKdBucket; ++k) {
        const int i = min(lo + k, hi - 1);  // a short run repeats its last node, which the compare ignores
        const double2 *p = reinterpret_cast<const double2 *>(nodes + i);
        a[k] = __ldg(p);
        b[k] = __ldg(p + 1);
    }
#pragma unroll
    for (int k = 0; k < kKdBucket; ++k) {
        const int idx = (int)(__double_as_longlong(b[k].y) & 0xffffffffll);
        const double d = dsq3(dsub(a[k].x, qx), dsub(a[k].y, qy), dsub(b[k].x, qz));
        if (d < best || (d == best && idx < bidx)) {
            best = d;
            bidx = idx;
        }
    }
}

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
// kFanOut: the answers of this rank's query shard are written straight into the result buffers of EVERY rank
// (peer memory over NVLink; nav_kdtree_nn_allgather_dev) -- the all-gather is the kernel's epilogue --, and
// the CTA that finishes last posts this rank's arrival in every rank's flag slot.
template <int kFanOut>   // 0: plain, 1: replicated tree + query shard delivered to all ranks, 2: map shard, partials to the owners
__global__ void __launch_bounds__(128)
k_kd_nn_stack_t(const KdNode *__restrict__ nodes, int n, const double *__restrict__ queries, long long nq,
                const int *__restrict__ perm, int *__restrict__ idx_out, double *__restrict__ dist_out,
                const __grid_constant__ KdFanOut fan) {
    const long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool live = slot < nq;
    if (!kFanOut && !live) return;
    const long long qi = live ? (perm ? perm[slot] : slot) : 0;
    int bidx = -1;
    double best = INFINITY;
    if (live && n > 0) {
    const double qx = queries[qi * 3], qy = queries[qi * 3 + 1], qz = queries[qi * 3 + 2];
    int st_lo[32], st_hi[32];
    double st_plane[32];
    int sp = 0;
    int lo = 0, hi = n;
    while (true) {
        while (lo < hi) {
            if (hi - lo <= kKdBucket) {
                scan_run_batched(nodes, lo, hi, qx, qy, qz, best, bidx);
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
    }
    const double dist = bidx >= 0 ? __dsqrt_rn(best) : INFINITY;
    if (!kFanOut) {
        idx_out[qi] = bidx;
        dist_out[qi] = dist;
        return;
    }
    if (live) {
        if (kFanOut == 1) {
            const long long at = fan.q_lo + qi;
#pragma unroll 1
            for (int r = 0; r < fan.world; ++r) {
                fan.idx[r][at] = bidx;
                fan.dist[r][at] = dist;
            }
        } else {
            const int owner = shard_owner(fan.nq_total, fan.world, qi);
            long long lo, hi;
            shard_range(fan.nq_total, fan.world, owner, lo, hi);
            const long long at = (long long)fan.rank * fan.shard_cap + (qi - lo);
            fan.pdsq[owner][at] = best;   // the squared distance decides the merge, as it decides the search
            fan.pidx[owner][at] = bidx >= 0 ? bidx + fan.idx_offset : -1;
        }
        __threadfence_system();  // the stores to peer memory are performed before this thread's part in the ticket
    }
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(fan.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x < fan.world) {
        __threadfence_system();
        if (threadIdx.x == 0) *fan.ticket = 0u;
        *(volatile unsigned long long *)(fan.flags[threadIdx.x] + fan.rank) = fan.seq;  // "rank has delivered call seq"
    }
}

// waits until every rank has posted `seq` in this rank's flag slots (local memory; the peers write it);
// err[0] is set if that takes longer than timeout_ns
__global__ void k_peer_wait(const unsigned long long *flags, int world, unsigned long long seq, unsigned long long timeout_ns,
                            unsigned *err) {
    if ((int)threadIdx.x >= world) return;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*(volatile const unsigned long long *)(flags + threadIdx.x) < seq) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) {
            atomicExch(err, 1u + threadIdx.x);
            return;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

cudaError_t kd_nn(const KdNode *d_nodes, size_t n, const double *d_bbox, const double *d_queries, size_t nq,
                  int *d_idx, double *d_dist, int sm_count, cudaStream_t stream, uint64_t *launches,
                  unsigned long long *d_counter) {
    if (nq == 0) return cudaSuccess;
    const int threads = 128;
    const unsigned grid = (unsigned)((nq + threads - 1) / threads);
    // (Ordering the queries along a Morton curve was measured in round 1: 13 % off the traversal, 75 us of
    // sorting in front of it -- dropped, together with the library sort it needed; profiles/README.md.)
    // Kernel choice by measurement on B200 (131 072 queries, widest-extent trees, 8-node runs scanned;
    // profiles/README.md), short stack / plain stackless / converged stackless:
    //   1 M uniform points 80 / 96 / 108 us, 4 M 124 / 126 / 144 us, 10 M 164 / 178 / 171 us,
    //   accumulated room map (1 M points on surfaces) 122 / 162 / 185 us.
    // The short stack wins or ties everywhere, so it is the default; NAV_KD_KERNEL = plain|conv|stack
    // selects the others (same answers).
    static const char *kk = getenv("NAV_KD_KERNEL");  // read once
    const bool use_stack = kk ? !strcmp(kk, "stack") : true;
    const bool use_conv = kk && !strcmp(kk, "conv");
    if (use_conv && d_counter && n > 0) {
        cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream);
        unsigned pgrid = (unsigned)sm_count * 12u;  // 12 x 128 threads = 48 warps per SM (40 registers)
        if ((unsigned long long)pgrid * threads > nq) pgrid = (unsigned)((nq + threads - 1) / threads);
        k_kd_nn_conv<<<pgrid, threads, 0, stream>>>(d_nodes, (int)n, d_queries, (long long)nq, d_idx, d_dist,
                                                    d_counter);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    if (use_stack)
        k_kd_nn_stack_t<0><<<grid, threads, 0, stream>>>(d_nodes, (int)n, d_queries, (long long)nq, nullptr, d_idx, d_dist,
                                                          KdFanOut{});
    else
        k_kd_nn<<<grid, threads, 0, stream>>>(d_nodes, (int)n, d_queries, (long long)nq, nullptr, d_idx, d_dist);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// the query shard [fan.q_lo, fan.q_lo + nq) answered and delivered to every rank's buffer (see KdFanOut)
cudaError_t kd_nn_fanout(const KdNode *d_nodes, size_t n, const double *d_queries, size_t nq, const KdFanOut &fan,
                         cudaStream_t stream, uint64_t *launches) {
    const int threads = 128;
    // an empty shard still has to post its arrival: one CTA
    const unsigned grid = (unsigned)((nq + threads - 1) / threads) ? (unsigned)((nq + threads - 1) / threads) : 1u;
    k_kd_nn_stack_t<1><<<grid, threads, 0, stream>>>(d_nodes, (int)n, d_queries, (long long)nq, nullptr, nullptr, nullptr, fan);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// Point-sharded map, second step: this rank owns the query shard [q_lo, q_lo + n_mine); for each of its queries it
// takes the smallest (squared distance, global index) of the `world` partial answers -- exactly the pair a search of
// the whole map would keep --, and delivers index and distance to every rank's result arrays; the CTA that finishes
// last posts the arrival flags of the second round.
__global__ void __launch_bounds__(128)
k_peer_merge(const __grid_constant__ KdFanOut fan, long long n_mine) {
    const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (q < n_mine) {
        double best = INFINITY;
        int bidx = -1;
        for (int r = 0; r < fan.world; ++r) {
            const long long at = (long long)r * fan.shard_cap + q;
            const double d = __ldcg(fan.pdsq[fan.rank] + at);
            const int i = __ldcg(fan.pidx[fan.rank] + at);
            if (i >= 0 && (d < best || (d == best && i < bidx) || bidx < 0)) {
                best = d;
                bidx = i;
            }
        }
        const double dist = bidx >= 0 ? __dsqrt_rn(best) : INFINITY;
#pragma unroll 1
        for (int r = 0; r < fan.world; ++r) {
            fan.idx[r][fan.q_lo + q] = bidx;
            fan.dist[r][fan.q_lo + q] = dist;
        }
        __threadfence_system();
    }
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(fan.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x < fan.world) {
        __threadfence_system();
        if (threadIdx.x == 0) *fan.ticket = 0u;
        *(volatile unsigned long long *)(fan.flags[threadIdx.x] + fan.rank) = fan.seq;
    }
}

cudaError_t kd_nn_partial(const KdNode *d_nodes, size_t n, const double *d_queries, size_t nq, const KdFanOut &fan,
                          cudaStream_t stream, uint64_t *launches) {
    const int threads = 128;
    const unsigned grid = (unsigned)((nq + threads - 1) / threads) ? (unsigned)((nq + threads - 1) / threads) : 1u;
    k_kd_nn_stack_t<2><<<grid, threads, 0, stream>>>(d_nodes, (int)n, d_queries, (long long)nq, nullptr, nullptr, nullptr, fan);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t peer_merge(const KdFanOut &fan, cudaStream_t stream) {
    long long lo, hi;
    shard_range(fan.nq_total, fan.world, fan.rank, lo, hi);
    const long long n_mine = hi - lo;
    const unsigned grid = (unsigned)((n_mine + 127) / 128) ? (unsigned)((n_mine + 127) / 128) : 1u;
    k_peer_merge<<<grid, 128, 0, stream>>>(fan, n_mine);
    return cudaGetLastError();
}

cudaError_t peer_wait(const unsigned long long *flags, int world, unsigned long long seq, unsigned *err, cudaStream_t stream) {
    k_peer_wait<<<1, 32, 0, stream>>>(flags, world, seq, 30000000000ull /* 30 s */, err);
    return cudaGetLastError();
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
