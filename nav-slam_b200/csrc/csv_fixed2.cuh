// csv_fixed2.cuh -- printf("%.2f") of a binary64 with integers only, shared by the host writer (io.cu)
// and the device writer (csvfmt.cu).  round-half-even of |v| * 100 on the exact binary value, which is
// what glibc's correctly rounded conversion prints (src/main.c:324 uses "%.2f" for 21 of 25 columns).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define NAV_HD __host__ __device__ __forceinline__
#else
#define NAV_HD inline
#endif

namespace nav {

// q = round_half_even(|v| * 100) for finite |v| < 2^57; false for everything else (inf, nan, huge)
NAV_HD bool fixed2_scaled(double v, bool &neg, unsigned long long &q) {
    unsigned long long bits;
#if defined(__CUDA_ARCH__)
    bits = (unsigned long long)__double_as_longlong(v);
#else
    memcpy(&bits, &v, 8);
#endif
    neg = (bits >> 63) != 0;
    const int ex = (int)((bits >> 52) & 0x7ff);
    unsigned long long m = bits & 0xfffffffffffffull;
    if (ex == 0x7ff) return false;
    int e;  // |v| = m * 2^e
    if (ex == 0) {
        e = -1074;
    } else {
        m |= 1ull << 52;
        e = ex - 1075;
    }
    if (e >= 0) {
        if (e > 4) return false;
        q = (m << e) * 100ull;  // < 2^57 * 100 < 2^64
        return true;
    }
    const int k = -e;
    const unsigned long long prod = m * 100ull;  // < 2^60
    if (k >= 62) {  // |v| * 100 < 2^60 / 2^62: rounds to zero
        q = 0;
        return true;
    }
    unsigned long long quo = prod >> k;
    const unsigned long long rem = prod & ((1ull << k) - 1), half = 1ull << (k - 1);
    if (rem > half || (rem == half && (quo & 1))) ++quo;
    q = quo;
    return true;
}

NAV_HD int dec_len(unsigned long long v) {
    int n = 1;
    while (v >= 10000) {
        v /= 10000;
        n += 4;
    }
    if (v >= 1000) return n + 3;
    if (v >= 100) return n + 2;
    if (v >= 10) return n + 1;
    return n;
}

// writes the `len` decimal digits of v ending at o[len-1]; returns o + len
template <typename CharPtr>
NAV_HD CharPtr put_uint(CharPtr o, unsigned long long v, int len) {
    for (int i = len - 1; i >= 0; --i) {
        o[i] = (char)('0' + (int)(v % 10));
        v /= 10;
    }
    return o + len;
}

NAV_HD int fixed2_len(bool neg, unsigned long long q) { return (neg ? 1 : 0) + dec_len(q / 100) + 3; }

template <typename CharPtr>
NAV_HD CharPtr put_fixed2(CharPtr o, bool neg, unsigned long long q) {
    if (neg) *o++ = '-';
    const unsigned long long ip = q / 100;
    const int fr = (int)(q - ip * 100);
    o = put_uint(o, ip, dec_len(ip));
    o[0] = '.';
    o[1] = (char)('0' + fr / 10);
    o[2] = (char)('0' + fr % 10);
    return o + 3;
}

NAV_HD int int_len(long long v) { return v < 0 ? 1 + dec_len((unsigned long long)(-v)) : dec_len((unsigned long long)v); }

template <typename CharPtr>
NAV_HD CharPtr put_int(CharPtr o, long long v) {
    if (v < 0) {
        *o++ = '-';
        return put_uint(o, (unsigned long long)(-v), dec_len((unsigned long long)(-v)));
    }
    return put_uint(o, (unsigned long long)v, dec_len((unsigned long long)v));
}

// the fixed part of a CSV line in front of the 18 pose columns is at most
// 20 (timestamp) + 10 + 10 (row, col) + 3 * 22 (sign, 18 digits, ".dd") + 11 (distance) + 7 commas
constexpr int kCsvHeadMax = 124;
constexpr int kCsvTailMax = 480;  // device path: pose columns of one line, ",%.2f" x 18 + "\n"

}  // namespace nav
