// csvfmt.cu -- the CSV row writer of src/main.c:320-352 on the device.
// One line per pixel: "ts,row,col,x,y,z,distance" + the 18 pose columns, "%.2f" exact (csv_fixed2.cuh).
// Lines have different lengths, so the text is laid out by a two-level prefix sum:
//   k_csv_len    per-pixel line length -> per-block length            (reads 24 B + 4 B per pixel)
//   k_csv_scan   exclusive scan of the block lengths, total length    (one CTA)
//   k_csv_write  block scan -> every line formatted into shared memory at its offset, pose columns
//                appended warp-cooperatively, block text stored to HBM with consecutive bytes per lane
// Algorithmic bytes per pixel: 28 read (twice) + about 170 written.  Values that the integer path
// cannot print (inf, nan, |v| >= 2^57) raise a flag; the host API then formats that frame on the host.
#include <cuda_runtime.h>

#include "csv_fixed2.cuh"
#include "nav_kernels.cuh"

namespace nav {

namespace {

constexpr int kCsvBlock = 256;

struct LineHead {
    unsigned long long qx, qy, qz;
    int len;  // characters in front of the pose columns; 0 when the pixel cannot be printed here
    int dist;
    bool nx, ny, nz;
};

__device__ __forceinline__ LineHead line_head(const CsvJob &job, long long p, int row, int col) {
    LineHead h;
    const double *src = job.cloud + 3 * p;
    const double x = src[0], y = src[1], z = src[2];
    const bool ok = fixed2_scaled(x, h.nx, h.qx) & fixed2_scaled(y, h.ny, h.qy) & fixed2_scaled(z, h.nz, h.qz);
    h.dist = job.dist ? job.dist[p] : 0;
    h.len = ok ? job.ts_len + dec_len((unsigned)row) + dec_len((unsigned)col) + fixed2_len(h.nx, h.qx) +
                     fixed2_len(h.ny, h.qy) + fixed2_len(h.nz, h.qz) + int_len(h.dist) + 6
               : 0;
    return h;
}

template <int BLOCK>
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned *warp_sums, unsigned &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    unsigned base = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w) {
        const unsigned s = warp_sums[w];
        if (w < warp) base += s;
        sum += s;
    }
    total = sum;
    return base + inc - v;
}

__global__ void __launch_bounds__(kCsvBlock) k_csv_len(CsvJob job, unsigned *block_len, unsigned *flags) {
    __shared__ unsigned warp_sums[kCsvBlock / 32];
    const long long p = (long long)blockIdx.x * kCsvBlock + threadIdx.x;
    unsigned len = 0;
    if (p < job.n) {
        const int row = (int)(p / job.cols), col = (int)(p - (long long)row * job.cols);
        const LineHead h = line_head(job, p, row, col);
        if (h.len == 0) atomicOr(flags, 1u);
        len = (unsigned)(h.len + job.tail_len);
    }
    unsigned total;
    block_exclusive_scan<kCsvBlock>(len, warp_sums, total);
    if (threadIdx.x == 0) block_len[blockIdx.x] = total;
}

// one CTA walks the block lengths in chunks of its size, carrying the running sum
__global__ void __launch_bounds__(1024) k_csv_scan(const unsigned *block_len, int n_blocks,
                                                    unsigned long long *block_off, unsigned long long *total_out) {
    __shared__ unsigned warp_sums[32];
    unsigned long long carry = 0;
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < n_blocks ? block_len[i] : 0;
        unsigned total;
        const unsigned ex = block_exclusive_scan<1024>(v, warp_sums, total);
        if (i < n_blocks) block_off[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(kCsvBlock) k_csv_write(CsvJob job, const unsigned long long *block_off,
                                                          char *out) {
    extern __shared__ char text[];  // kCsvBlock * (kCsvHeadMax + tail_len)
    __shared__ unsigned warp_sums[kCsvBlock / 32];
    __shared__ unsigned line_end[kCsvBlock];  // offset just behind each line's head
    __shared__ char s_tail[kCsvTailMax];      // lanes index it with different offsets: not from the param bank
    for (int j = threadIdx.x; j < job.tail_len; j += kCsvBlock) s_tail[j] = job.tail[j];
    const long long p = (long long)blockIdx.x * kCsvBlock + threadIdx.x;
    LineHead h;
    h.len = 0;
    int row = 0, col = 0;
    unsigned len = 0;
    if (p < job.n) {
        row = (int)(p / job.cols);
        col = (int)(p - (long long)row * job.cols);
        h = line_head(job, p, row, col);
        len = (unsigned)(h.len + job.tail_len);
    }
    unsigned total;
    const unsigned off = block_exclusive_scan<kCsvBlock>(len, warp_sums, total);
    if (p < job.n && h.len) {
        char *o = text + off;
        o = put_uint(o, job.ts, job.ts_len);
        *o++ = ',';
        o = put_uint(o, (unsigned)row, dec_len((unsigned)row));
        *o++ = ',';
        o = put_uint(o, (unsigned)col, dec_len((unsigned)col));
        *o++ = ',';
        o = put_fixed2(o, h.nx, h.qx);
        *o++ = ',';
        o = put_fixed2(o, h.ny, h.qy);
        *o++ = ',';
        o = put_fixed2(o, h.nz, h.qz);
        *o++ = ',';
        o = put_int(o, h.dist);
    }
    line_end[threadIdx.x] = (p < job.n) ? off + (unsigned)h.len : 0xffffffffu;
    __syncthreads();
    // pose columns: the lanes of a warp copy the shared tail behind each of the warp's 32 heads
    const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
    for (int l = 0; l < 32; ++l) {
        const unsigned at = line_end[wbase + l];
        if (at == 0xffffffffu) break;
        for (int j = lane; j < job.tail_len; j += 32) text[at + j] = s_tail[j];
    }
    __syncthreads();
    char *dst = out + block_off[blockIdx.x];
    for (unsigned j = threadIdx.x; j < total; j += kCsvBlock) dst[j] = text[j];
}

}  // namespace

size_t csv_scratch_bytes(long long n) {
    const size_t nb = (size_t)((n + kCsvBlock - 1) / kCsvBlock);
    return nb * (sizeof(unsigned) + sizeof(unsigned long long)) + 64;
}

// scratch layout: [total u64][flags u32][pad][block_off u64 x nb][block_len u32 x nb]
int launch_csv_format(const CsvJob &job, char *d_text, void *d_scratch, cudaStream_t stream) {
    const int nb = (int)((job.n + kCsvBlock - 1) / kCsvBlock);
    unsigned long long *total = (unsigned long long *)d_scratch;
    unsigned *flags = (unsigned *)((char *)d_scratch + 8);
    unsigned long long *block_off = (unsigned long long *)((char *)d_scratch + 16);
    unsigned *block_len = (unsigned *)(block_off + nb);
    const size_t smem = (size_t)kCsvBlock * (size_t)(kCsvHeadMax + job.tail_len);
    if (cudaFuncSetAttribute(k_csv_write, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kCsvBlock * (kCsvHeadMax + kCsvTailMax)) != cudaSuccess)
        return 1;
    if (cudaMemsetAsync(d_scratch, 0, 16, stream) != cudaSuccess) return 1;
    k_csv_len<<<nb, kCsvBlock, 0, stream>>>(job, block_len, flags);
    k_csv_scan<<<1, 1024, 0, stream>>>(block_len, nb, block_off, total);
    k_csv_write<<<nb, kCsvBlock, smem, stream>>>(job, block_off, d_text);
    return cudaGetLastError() != cudaSuccess;
}

}  // namespace nav
