// Decimal text -> int / double for the localmap reader (the fast input path; the reference reads
// the same files with fscanf "%d" / "%lf", LinearSFMImp.cpp:3044-3132).  Host only.
//
// parse_double returns EXACTLY what strtod returns (value bits and end pointer) -- it only takes a shortcut
// when the shortcut is provably the correctly rounded result, and calls strtod otherwise:
//   * at most 15 significant digits and |exponent| <= 22: the digits as an integer w (exact in a double) times
//     or divided by 10^k (exact in a double) is ONE correctly rounded operation (Clinger's fast path);
//   * at most 19 significant digits (w < 2^64) and |exponent| <= 27: the same single operation in the x87
//     extended format (64-bit significand: w and 10^k are exact there), then rounded to double.  Rounding twice
//     can only differ from rounding once when the extended result sits exactly on a double-precision tie
//     (low 11 bits of its significand = 0x400): that case goes to strtod (1 in 2048);
//   * anything else (more digits, big exponents, hex floats, inf / nan, no digits at all) goes to strtod.
// tests/helpers/fast_num_check.cpp compares both routines with strtod / strtol on random and edge-case inputs.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace fastnum {

static inline bool is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// strtol(p, &end, 10) for values that fit 18 digits; longer digit strings go to strtol itself
static inline long parse_long(const char *p, char **end)
{
    const char *s = p;
    while (is_space(*s)) s++;
    bool neg = false;
    if (*s == '-') { neg = true; s++; } else if (*s == '+') s++;
    const char *d0 = s;
    unsigned long long w = 0;
    while (*s >= '0' && *s <= '9') { w = w * 10 + (unsigned)(*s - '0'); s++; }
    if (s == d0 || s - d0 > 18) return strtol(p, end, 10);
    *end = const_cast<char *>(s);
    return neg ? -(long)w : (long)w;
}

static inline double parse_double(const char *p, char **end)
{
    static const double p10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    static const long double p10l[28] = {1e0L,  1e1L,  1e2L,  1e3L,  1e4L,  1e5L,  1e6L,  1e7L,  1e8L,  1e9L,
                                         1e10L, 1e11L, 1e12L, 1e13L, 1e14L, 1e15L, 1e16L, 1e17L, 1e18L, 1e19L,
                                         1e20L, 1e21L, 1e22L, 1e23L, 1e24L, 1e25L, 1e26L, 1e27L};
    const char *s = p;
    while (is_space(*s)) s++;
    bool neg = false;
    if (*s == '-') { neg = true; s++; } else if (*s == '+') s++;
    uint64_t w = 0;
    int nd = 0;              // significant digits held by w (leading zeros do not count)
    int e10 = 0;
    bool any = false, lost = false;
    const char *int0 = s;
    while (*s >= '0' && *s <= '9') {
        const unsigned d = (unsigned)(*s - '0');
        if (nd < 19) { w = w * 10 + d; nd += (w != 0); }
        else { e10++; lost = lost || d != 0; }
        any = true; s++;
    }
    if (s - int0 == 1 && *int0 == '0' && (*s == 'x' || *s == 'X')) return strtod(p, end);      // hex float
    if (*s == '.') {
        s++;
        while (*s >= '0' && *s <= '9') {
            const unsigned d = (unsigned)(*s - '0');
            if (nd < 19) { w = w * 10 + d; nd += (w != 0); e10--; }
            else lost = lost || d != 0;
            any = true; s++;
        }
    }
    if (!any) return strtod(p, end);                    // inf, nan, or not a number at all
    if (*s == 'e' || *s == 'E') {
        const char *t = s + 1;
        bool eneg = false;
        if (*t == '-') { eneg = true; t++; } else if (*t == '+') t++;
        if (*t >= '0' && *t <= '9') {
            int ex = 0;
            while (*t >= '0' && *t <= '9') { if (ex < 100000) ex = ex * 10 + (*t - '0'); t++; }
            e10 += eneg ? -ex : ex;
            s = t;
        }                                               // "1e", "1e+": the exponent marker is not part of the number
    }
    if (lost) return strtod(p, end);
    double v;
    if (w == 0) v = 0.0;
    else if (nd <= 15 && e10 >= -22 && e10 <= 22) {
        v = (double)w;
        v = e10 < 0 ? v / p10[-e10] : v * p10[e10];
    } else if (e10 >= -27 && e10 <= 27) {
#if defined(__x86_64__) && defined(__LDBL_MANT_DIG__) && __LDBL_MANT_DIG__ == 64
        long double x = (long double)w;
        x = e10 < 0 ? x / p10l[-e10] : x * p10l[e10];
        uint64_t sig;
        memcpy(&sig, &x, sizeof(sig));                  // 80-bit extended: the first 8 bytes are the significand
        if ((sig & 0x7ff) == 0x400) return strtod(p, end);
        v = (double)x;
#else
        return strtod(p, end);
#endif
    } else return strtod(p, end);
    *end = const_cast<char *>(s);
    return neg ? -v : v;
}

// ---- output: the bytes of printf("%lf") = "%.6f" and of "%d" (the reference's writers, LinearSFMImp.cpp:2113,
// 7946, 7959) ----
// |v| = m * 2^-s exactly (m < 2^53), so |v| * 10^6 = (m * 10^6) / 2^s is an exact rational with a 73-bit numerator:
// the shift with round-half-to-even on the exact remainder is what glibc's exact printf produces.  Values of 2^43 and
// above, infinities and NaNs go to snprintf.  Returns the number of characters written (no terminator).
static inline int put_uint(char *buf, unsigned long long x)
{
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + x % 10); x /= 10; } while (x);
    for (int i = 0; i < n; i++) buf[i] = tmp[n - 1 - i];
    return n;
}

static inline int put_int(char *buf, int x)
{
    if (x < 0) { buf[0] = '-'; return 1 + put_uint(buf + 1, (unsigned long long)(-(long long)x)); }
    return put_uint(buf, (unsigned long long)x);
}

static inline int put_f6(char *buf, size_t cap, double v)
{
    uint64_t u;
    memcpy(&u, &v, 8);
    const bool neg = (u >> 63) != 0;
    const int be = (int)((u >> 52) & 0x7ff);
    uint64_t m = u & ((1ull << 52) - 1);
    if (be == 0x7ff || be >= 1023 + 43) return snprintf(buf, cap, "%lf", v);       // inf / nan / |v| >= 2^43
    int s;                                           // |v| = m * 2^-s
    if (be == 0) s = 1074; else { m |= 1ull << 52; s = 1075 - be; }
    const unsigned __int128 P = (unsigned __int128)m * 1000000u;                   // < 2^73
    uint64_t q;
    if (s >= 100) q = 0;                             // |v| * 10^6 < 2^-27: rounds to zero
    else {
        q = (uint64_t)(P >> s);
        const unsigned __int128 rem = P & ((((unsigned __int128)1) << s) - 1), half = ((unsigned __int128)1) << (s - 1);
        if (rem > half || (rem == half && (q & 1))) q++;
    }
    int n = 0;
    if (neg) buf[n++] = '-';
    n += put_uint(buf + n, q / 1000000u);
    buf[n++] = '.';
    unsigned f = (unsigned)(q % 1000000u);
    for (int i = 5; i >= 0; i--) { buf[n + i] = (char)('0' + f % 10); f /= 10; }
    return n + 6;
}

} // namespace fastnum
