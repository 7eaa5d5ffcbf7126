// Fuzz / edge-case check of linearsfm_b200/csrc/fast_num.h against strtod / strtol (the value BITS and the end
// pointer must be identical for every input) and of its writers against printf("%lf") / printf("%d") (same bytes).   fast_num_check <seed> <cases>   ->  "mismatches N"
#include "fast_num.h"
#include <cstdio>
#include <cmath>
#include <cstring>
#include <random>
#include <string>
#include <vector>

static long long bad = 0, slow = 0;
static void check(const std::string &txt)
{
    const char *p = txt.c_str();
    char *e1, *e2;
    const double a = strtod(p, &e1), b = fastnum::parse_double(p, &e2);
    uint64_t ua, ub;
    memcpy(&ua, &a, 8); memcpy(&ub, &b, 8);
    const bool nan_ok = a != a && b != b;
    if ((ua != ub && !nan_ok) || e1 != e2) {
        if (bad < 20) printf("MISMATCH double '%s': strtod %.17g (+%ld) fast %.17g (+%ld)\n", p, a, (long)(e1 - p), b, (long)(e2 - p));
        bad++;
    }
    char *e3, *e4;
    const long c = strtol(p, &e3, 10), d = fastnum::parse_long(p, &e4);
    if (c != d || e3 != e4) {
        if (bad < 20) printf("MISMATCH long '%s': strtol %ld (+%ld) fast %ld (+%ld)\n", p, c, (long)(e3 - p), d, (long)(e4 - p));
        bad++;
    }
}

int main(int argc, char **argv)
{
    const unsigned seed = argc > 1 ? (unsigned)atoi(argv[1]) : 1u;
    const long long cases = argc > 2 ? atoll(argv[2]) : 1000000;
    const char *edge[] = {"0", "-0", "+0", "0.0", "-0.0", "1", "-1", "+1", ".5", "-.5", "5.", "5.e3", ".e3", ".", "-", "+", "", " ",
        "1e", "1e+", "1e-", "1e5", "1E5", "1e+5", "1e-5", "1e400", "1e-400", "-1e400", "0e400", "1e22", "1e23", "1e-22", "1e-23",
        "9007199254740992", "9007199254740993", "9007199254740991", "18446744073709551615", "18446744073709551616",
        "99999999999999999999", "0.1", "0.2", "0.3", "123456789012345678", "1234567890123456789", "12345678901234567890",
        "0x10", "0x1p3", "0x", "0xg", "inf", "-inf", "nan", "infinity", "INF", "NAN", "nan(1)", "1.7976931348623157e308",
        "1.7976931348623159e308", "2.2250738585072014e-308", "4.9406564584124654e-324", "2.4703282292062327e-324",
        "  \t\n 12.5", "\v\f\r3", "12abc", "12.5.6", "1e5e6", "1.5e+", "007", "0007.5000", "1,5", "1 5", "--1", "+-1", "-+1",
        "0.000000000000000000000000000000000001", "100000000000000000000000000000", "1.00000000000000000000000001",
        "1.0000000000000002220446049250313", "1.00000000000000011102230246251565404236316680908203125",
        "1.00000000000000011102230246251565404236316680908203124", "1.00000000000000011102230246251565404236316680908203126",
        "9223372036854775807", "-9223372036854775808", "9223372036854775808", "999999999999999999", "1000000000000000000",
        "8.5", "0.5e1", "5e-1", "3.14159265358979323846", "2.718281828459045", "6.02214076e23", "6.62607015e-34"};
    for (const char *e : edge) check(e);
    std::mt19937_64 rng(seed);
    char buf[128];
    for (long long i = 0; i < cases; i++) {
        double v;
        const int kind = (int)(rng() % 8);
        if (kind == 0) { uint64_t u = rng(); memcpy(&v, &u, 8); if (v != v || v - v != 0) v = 1.0; }      // any finite double
        else if (kind == 1) v = (double)(long long)(rng() % 2000001) - 1000000.0;                          // integers
        else {
            const double mant = (double)(rng() >> 11) / 9007199254740992.0;                               // [0,1)
            const int ex = (int)(rng() % 41) - 20;
            v = (mant * 2 - 1) * pow(10.0, ex);
        }
        static const char *fmts[] = {"%.17g", "%.16g", "%.15g", "%.12g", "%.6f", "%f", "%.10e", "%.18e", "%.20g", "%.9f", "%g", "%.16e", "%.15f", "%.19g"};
        snprintf(buf, sizeof buf, fmts[rng() % 14], v);
        std::string s(buf);
        const int deco = (int)(rng() % 16);
        if (deco == 0) s = " " + s; else if (deco == 1) s = "+" + s; else if (deco == 2) s += " 17"; else if (deco == 3) s += "e";
        else if (deco == 4) s += "x"; else if (deco == 5) s = "\n\t" + s + "\n";
        check(s);
        if (kind == 2) {       // random digit strings with a random point and exponent: hits the 16-19 digit and tie paths
            std::string t;
            const int nd = 1 + (int)(rng() % 24);
            const int dot = (int)(rng() % (nd + 1));
            for (int k = 0; k < nd; k++) { if (k == dot) t += '.'; t += (char)('0' + rng() % 10); }
            if (rng() % 2) { t += (rng() % 2) ? "e" : "E"; if (rng() % 2) t += (rng() % 2) ? "-" : "+"; t += std::to_string(rng() % 40); }
            check(t);
        }
    }
    // exact double-precision ties written with 17-19 digits: the extended path must hand them to strtod
    for (long long i = 0; i < cases / 10; i++) {
        const uint64_t m = (1ull << 52) | (rng() & ((1ull << 52) - 1));
        // (2m+1) * 2^k for small k is a tie between two doubles when written out exactly; keep <= 19 digits
        const unsigned __int128 t = (unsigned __int128)(2 * m + 1);
        unsigned long long lo = (unsigned long long)t;
        snprintf(buf, sizeof buf, "%llu", lo);
        check(buf);
        snprintf(buf, sizeof buf, "%llu.5", (unsigned long long)m);      // m + 0.5: tie at the 2^52 scale
        check(buf);
        snprintf(buf, sizeof buf, "%llue-3", lo);
        check(buf);
    }
    // ---- writers: put_f6 / put_int must produce the bytes of printf("%lf") / printf("%d") ----
    auto chk_f6 = [&](double v) {
        char a[400], b[400];
        const int la = snprintf(a, sizeof a, "%lf", v), lb = fastnum::put_f6(b, sizeof b, v);
        if (la != lb || memcmp(a, b, la)) { if (bad < 20) { b[lb] = 0; printf("MISMATCH %%lf of %.17g: '%s' vs '%s'\n", v, a, b); } bad++; }
    };
    const double fedge[] = {0.0, -0.0, 1.0, -1.0, 0.5e-6, 1.5e-6, 2.5e-6, 0.9999995, 0.99999949999999, 1e-7, 4.9e-324, -4.9e-324,
        2.2e-308, 123456.7890125, 123456.7890135, 8796093022207.999, 8796093022208.0, 8796093022208.5, 1e13, 1e15, 1e22, 1e300,
        -1e300, INFINITY, -INFINITY, NAN, 0.1, 0.2, 0.3, 999999.9999995, 999999.9999994999, 1e-6, 5e-7, 4.999999999999999e-7,
        5.000000000000001e-7, -5e-7, -4.9e-7};
    for (double v : fedge) chk_f6(v);
    for (long long i = 0; i < cases; i++) {
        double v;
        const int k = (int)(rng() % 6);
        if (k == 0) { uint64_t u = rng(); memcpy(&v, &u, 8); }
        else if (k == 1) v = ((double)(rng() >> 11) / 9007199254740992.0 * 2 - 1) * pow(10.0, (int)(rng() % 30) - 15);
        else if (k == 2) v = (double)(long long)(rng() % 4000001 - 2000000) / (double)(1ull << (rng() % 30));      // exact binary fractions: true ties
        else if (k == 3) v = (double)(long long)(rng() % 2000001 - 1000000) * 1e-6 + ((int)(rng() % 3) - 1) * 5e-7;
        else if (k == 4) v = ((double)(rng() >> 11) / 9007199254740992.0 * 2 - 1) * 1000.0;
        else v = ldexp((double)(rng() >> 11), -(int)(rng() % 120));
        chk_f6(v);
    }
    for (long long i = 0; i < cases / 10 + 10; i++) {
        static const int iedge[10] = {0, 1, -1, 2147483647, -2147483647 - 1, 10, -10, 99, 100, -100};
        const int x = i < 10 ? iedge[i] : (int)rng();
        char a[32], b[32];
        const int la = snprintf(a, 32, "%d", x), lb = fastnum::put_int(b, x);
        if (la != lb || memcmp(a, b, la)) { if (bad < 20) printf("MISMATCH %%d of %d\n", x); bad++; }
    }
    printf("mismatches %lld\n", bad);
    return bad != 0;
}
