/* Compares gort_fmt_f with snprintf("%f") byte for byte.  usage: fastfmt_check <count> <seed>; exit 0 = identical */
#include <stdlib.h>
#include "../../gort_b200/host/fastfmt.h"

static uint64_t s[2];
static uint64_t next(void) { uint64_t a = s[0], b = s[1]; s[0] = b; a ^= a << 23; s[1] = a ^ b ^ (a >> 17) ^ (b >> 26); return s[1] + b; }

static long bad = 0, checked = 0;
static void check(double x)
{
    char a[400], b[400];
    int n = gort_fmt_f(a, x);
    int m = snprintf(b, sizeof b, "%f", x);
    checked++;
    if (n != m || memcmp(a, b, (size_t) n) != 0) {
        if (bad++ < 10) { a[n] = 0; fprintf(stderr, "MISMATCH %.17g: ours '%s' printf '%s'\n", x, a, b); }
    }
}

int main(int argc, char **argv)
{
    long count = argc > 1 ? atol(argv[1]) : 1000000;
    s[0] = argc > 2 ? (uint64_t) atoll(argv[2]) : 12345; s[1] = 0x9E3779B97F4A7C15ULL;
    /* edge values */
    const double edge[] = { 0.0, -0.0, 1e-7, -1e-7, 4.9999999e-7, 5e-7, 5.0000001e-7, 0.9999995, 0.99999949999999, 1.0, 9.9999995, 123456789.1234565,
                            999999999.9999995, 1e9, 1e15, 1e22, 1e300, -1e300, 2.5e-6, 0.0078125, 0.0234375, 1.5e-6, 0.5, 0.05, 360.0, 89.999999,
                            NAN, -NAN, INFINITY, -INFINITY, 4.9406564584124654e-324, 2.2250738585072014e-308 };
    for (size_t i = 0; i < sizeof edge / sizeof edge[0]; i++) { check(edge[i]); check(-edge[i]); }
    /* exact ties k + 0.5 in the seventh decimal: x = m / 128 * 2^-j with odd m gives x * 1e6 = (odd * 15625) / 2^(j+1) ... */
    for (int m = 1; m < 20000; m += 2) { check(m / 128.0); check(-m / 128.0); check(m / 128.0 / 1e0 + 0.0); }
    for (long i = 0; i < count; i++) {
        uint64_t r = next();
        double u = (double) (r >> 11) * (1.0 / 9007199254740992.0);       /* [0,1) */
        switch (i & 7) {
        case 0: check(u); break;                                           /* reflectances */
        case 1: check(u * 360.0 - 180.0); break;                           /* angles */
        case 2: check(u * 1e-5); break;                                    /* tiny */
        case 3: check((double) (next() % 2000000) / 2e6 + u * 1e-13); break;   /* near ties / grid points */
        case 4: check((double) (next() % 2000001) * 5e-7); break;          /* decimal ties as doubles */
        case 5: check(u * 1e9); break;
        case 6: { double x; uint64_t b = next(); memcpy(&x, &b, 8); check(x); break; }   /* any bit pattern */
        default: check(floor(u * 1e6) / 1e6); break;
        }
    }
    fprintf(stderr, "%ld values checked, %ld mismatches\n", checked, bad);
    return bad ? 1 : 0;
}
