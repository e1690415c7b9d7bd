// io.cu -- the data formats either side of the path, host code inside the C-ABI
// library: the L9 CSV reader of src/main.c:77-128 and the 25-column CSV row writer of
// src/main.c:243,320-352, both byte/bit compatible with the reference and several times faster
// (no fscanf / fprintf in the loop).  These are the callers' formats, not kernels: nothing here touches
// the GPU, and frames are parsed straight into caller memory (pinned memory from nav_host_alloc
// makes them a DMA source for the frame calls).
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/navslam_b200.h"
#include "csv_fixed2.cuh"

extern "C" const char *nav_last_error(void);
int nav_io_fail(const char *fmt, ...);  // defined in capi.cu (sets the thread-local error)

// ---------------------------------------------------------------------------------- reader -----
// src/main.c:86-118: skip one header line (at most 255 characters are consumed by the reference's
// fgets), then records "%d,%d,%d,%lf,%lf,%lf,%d" until the first one that does not parse.  A record
// whose row/col is out of range is skipped without touching the frame bookkeeping; a change of the
// frame number starts the next frame slot.  Unlike the reference (col > MAX_COLS, main.c:100) a
// record with col == cols is rejected instead of written out of bounds.
namespace {

inline const char *skip_ws(const char *p, const char *end) {  // fscanf directives skip leading white space
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r' || *p == '\v' || *p == '\f')) ++p;
    return p;
}

inline bool parse_int(const char *&p, const char *end, int &out) {
    p = skip_ws(p, end);
    const char *q = p;
    bool neg = false;
    if (q < end && (*q == '-' || *q == '+')) neg = *q++ == '-';
    if (q >= end || *q < '0' || *q > '9') return false;
    long long v = 0;
    while (q < end && *q >= '0' && *q <= '9') {
        v = v * 10 + (*q++ - '0');
        if (v > 0x7fffffffLL + 1) v = 0x7fffffffLL + 1;
    }
    out = (int)(neg ? -v : v);
    p = q;
    return true;
}

inline bool parse_double(const char *&p, const char *end, double &out) {
    p = skip_ws(p, end);
    if (p >= end) return false;
    // fast path: optional sign, up to 15 significant digits with an optional fraction, no exponent:
    // the value is an exactly representable integer divided by an exactly representable power of ten,
    // which one IEEE division rounds correctly -- the same result strtod / fscanf("%lf") give.
    const char *q = p;
    bool neg = false;
    if (*q == '-' || *q == '+') neg = *q++ == '-';
    unsigned long long mant = 0;
    int digits = 0, frac = 0;
    bool any = false;
    while (q < end && *q >= '0' && *q <= '9') {
        mant = mant * 10 + (unsigned)(*q++ - '0');
        if (mant) ++digits;
        any = true;
    }
    if (q < end && *q == '.') {
        ++q;
        while (q < end && *q >= '0' && *q <= '9') {
            mant = mant * 10 + (unsigned)(*q++ - '0');
            if (mant) ++digits;
            ++frac;
            any = true;
        }
    }
    static const double p10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const bool simple = any && digits <= 15 && frac <= 22 && (q >= end || (*q != 'e' && *q != 'E' && *q != 'x' && *q != 'X' &&
                                                                          *q != 'n' && *q != 'N' && *q != 'i' && *q != 'I'));
    if (simple) {
        const double v = (double)mant / p10[frac];
        out = neg ? -v : v;
        p = q;
        return true;
    }
    // anything else (exponents, inf/nan, long mantissas): the C library, on a bounded copy
    char tmp[512];
    size_t len = (size_t)(end - p) < sizeof(tmp) - 1 ? (size_t)(end - p) : sizeof(tmp) - 1;
    memcpy(tmp, p, len);
    tmp[len] = 0;
    char *stop = nullptr;
    const double v = strtod(tmp, &stop);
    if (stop == tmp) return false;
    out = v;
    p += stop - tmp;
    return true;
}

inline bool expect(const char *&p, const char *end, char c) {  // a literal in the format: no white-space skipping
    if (p < end && *p == c) {
        ++p;
        return true;
    }
    return false;
}

}  // namespace

extern "C" int nav_l9_csv_read(const char *path, int rows, int cols, size_t max_frames, nav_point *frames_out,
                               int *timestamps_out, size_t *n_frames_out) {
    if (!path || !frames_out || !n_frames_out || rows < 1 || cols < 1)
        return nav_io_fail("nav_l9_csv_read: bad argument");
    *n_frames_out = 0;
    FILE *fp = fopen(path, "rb");
    if (!fp) return nav_io_fail("nav_l9_csv_read: cannot open %s: %s", path, strerror(errno));
    std::string text;
    {
        char chunk[1 << 16];
        size_t got;
        while ((got = fread(chunk, 1, sizeof(chunk), fp)) > 0) text.append(chunk, got);
        fclose(fp);
    }
    const char *p = text.data(), *end = p + text.size();
    {   // fgets(header, 256): at most 255 characters, stopping after the first newline
        size_t n = 0;
        while (p < end && n < 255) {
            ++n;
            if (*p++ == '\n') break;
        }
        if (n == 0) return 0;  // empty file: zero frames, like the reference
    }
    const size_t npx = (size_t)rows * cols;
    long long current = -1;
    bool have_frame = false;
    size_t count = 0;  // index of the frame being filled
    while (true) {
        int frame, row, col, conf;
        double x, y, z;
        const char *q = p;
        if (!(parse_int(q, end, frame) && expect(q, end, ',') && parse_int(q, end, row) && expect(q, end, ',') &&
              parse_int(q, end, col) && expect(q, end, ',') && parse_double(q, end, x) && expect(q, end, ',') &&
              parse_double(q, end, y) && expect(q, end, ',') && parse_double(q, end, z) && expect(q, end, ',') &&
              parse_int(q, end, conf)))
            break;
        p = q;
        if (row < 0 || row >= rows || col < 0 || col >= cols) continue;
        if (!have_frame || frame != current) {
            if (have_frame) ++count;
            if (count >= max_frames) {  // the reference overruns its lidarData[10] here; stop instead
                *n_frames_out = max_frames;
                return nav_io_fail("nav_l9_csv_read: more than %zu frames in %s", max_frames, path);
            }
            have_frame = true;
            current = frame;
            if (timestamps_out) timestamps_out[count] = frame;
        }
        nav_point &dst = frames_out[count * npx + (size_t)row * cols + col];
        dst.x = x;
        dst.y = y;
        dst.z = z;
    }
    *n_frames_out = have_frame ? count + 1 : 0;
    return 0;
}

// ---------------------------------------------------------------------------------- writer -----
namespace {

// printf("%.2f") of any double: the integer path of csv_fixed2.cuh, the C library for inf / nan / >= 2^57
inline char *fmt_fixed2(char *o, double v) {
    bool neg;
    unsigned long long q;
    if (nav::fixed2_scaled(v, neg, q)) return nav::put_fixed2(o, neg, q);
    char tmp[400];
    const int n = snprintf(tmp, sizeof(tmp), "%.2f", v);
    memcpy(o, tmp, (size_t)n);
    return o + n;
}

inline char *fmt_uint(char *o, unsigned long long v) { return nav::put_uint(o, v, nav::dec_len(v)); }
inline char *fmt_int(char *o, long long v) { return nav::put_int(o, v); }

}  // namespace

namespace nav {
// ",%.2f" x 18 + "\n" (src/main.c:331-348); out needs 18 * 340 + 2 bytes for arbitrary doubles
size_t csv_pose_columns(char *out, const double imu[6], const nav_pos *lidar_pos, const nav_pos *ekf_pos) {
    char *t = out;
    const double zero6[6] = {0, 0, 0, 0, 0, 0};
    const double *im = imu ? imu : zero6;
    for (int i = 0; i < 6; ++i) {
        *t++ = ',';
        t = fmt_fixed2(t, im[i]);
    }
    const double lp[6] = {lidar_pos->x, lidar_pos->y, lidar_pos->z, lidar_pos->roll, lidar_pos->pitch, lidar_pos->yaw};
    for (int i = 0; i < 6; ++i) {
        *t++ = ',';
        t = fmt_fixed2(t, lp[i]);
    }
    const double ep[6] = {ekf_pos ? ekf_pos->x : 0, ekf_pos ? ekf_pos->y : 0, ekf_pos ? ekf_pos->z : 0,
                          ekf_pos ? ekf_pos->roll : 0, ekf_pos ? ekf_pos->pitch : 0, ekf_pos ? ekf_pos->yaw : 0};
    for (int i = 0; i < 6; ++i) {
        *t++ = ',';
        t = fmt_fixed2(t, ep[i]);
    }
    *t++ = '\n';
    return (size_t)(t - out);
}
}  // namespace nav

extern "C" const char *nav_csv_header(void) {
    // src/main.c:243
    return "Timestamp,Row,Col,x,y,z,distance,IMU_x,IMU_y,IMU_z,IMU_roll,IMU_pitch,IMU_yaw,LiDAR_x,LiDAR_y,LiDAR_z,"
           "LiDAR_roll,LiDAR_pitch,LiDAR_yaw,EKF_x,EKF_y,EKF_z,EKF_roll,EKF_pitch,EKF_yaw\n";
}

// One frame = rows*cols lines of src/main.c:324-349 (the L5 handler, which passes a double to every
// %.2f).  distances == NULL prints 0, imu == NULL / ekf_pos == NULL print 0.00 in their six columns: what
// the L9 handler (main.c:437-461) means by its literal 0 arguments -- as written it hands ints to %.2f,
// which is undefined behaviour and not reproduced.
extern "C" size_t nav_csv_format_frame(char *buf, size_t cap, unsigned long long timestamp, int rows, int cols,
                                       const nav_point *global_cloud, const int *distances, const double imu[6],
                                       const nav_pos *lidar_pos, const nav_pos *ekf_pos) {
    if (!buf || !global_cloud || !lidar_pos || rows < 1 || cols < 1) return 0;
    // the 18 pose columns are the same text on every line of the frame
    char tail[18 * 340 + 4];
    const size_t tail_len = nav::csv_pose_columns(tail, imu, lidar_pos, ekf_pos);
    char *o = buf;
    char *const lim = buf + cap;
    char head[3 * 340 + 80];
    for (int r = 0; r < rows; ++r) {
        for (int c = 0; c < cols; ++c) {
            const nav_point &p = global_cloud[(size_t)r * cols + c];
            char *h = fmt_uint(head, timestamp);
            *h++ = ',';
            h = fmt_int(h, r);
            *h++ = ',';
            h = fmt_int(h, c);
            *h++ = ',';
            h = fmt_fixed2(h, p.x);
            *h++ = ',';
            h = fmt_fixed2(h, p.y);
            *h++ = ',';
            h = fmt_fixed2(h, p.z);
            *h++ = ',';
            h = fmt_int(h, distances ? distances[(size_t)r * cols + c] : 0);
            const size_t head_len = (size_t)(h - head);
            if ((size_t)(lim - o) < head_len + tail_len) return 0;  // caller's buffer too small
            memcpy(o, head, head_len);
            memcpy(o + head_len, tail, tail_len);
            o += head_len + tail_len;
        }
    }
    return (size_t)(o - buf);
}
