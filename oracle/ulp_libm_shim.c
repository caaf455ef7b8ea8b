/* oracle/ulp_libm_shim.c -- TEST INFRASTRUCTURE, not product code.
 *
 * A "1-ULP libm": every transcendental the GORT path calls is evaluated in long double, rounded to
 * double, and then moved by +1 or -1 ULP pseudo-randomly (probability 1/4 each, keyed on the
 * argument bits and a seed).  Linked with -Wl,-Bsymbolic into libgort_oracle_ulp.so together with
 * the unchanged oracle sources, it shows how far the REFERENCE ALGORITHM ITSELF moves when libm is
 * merely 1-ULP accurate instead of glibc's rounding -- i.e. the conditioning of each output.
 * CUDA's libm is a 1-2 ULP libm, so the GPU path cannot agree with the reference better than this
 * on ill-conditioned entries (1 - exp(-x) for x ~ 1e-7, 1 - Kc - Kz - Kg ~ 0, ...).  The parity
 * tests use it only to EXPLAIN an entry that misses the strict 1e-9 bound: the miss must be within
 * a small multiple of this sensitivity.  sqrt/fabs/ceil/floor are exact in both libms: untouched.
 */
#include <stdint.h>
#include <string.h>

long double expl(long double), logl(long double), sinl(long double), cosl(long double), tanl(long double),
    atanl(long double), acosl(long double), asinl(long double), powl(long double, long double);
double nextafter(double, double);

static uint64_t g_seed = 1;
void gort_oracle_ulp_seed(uint64_t s) { g_seed = s * 2 + 1; }

static double jitter(double r, double x, uint64_t salt)
{
    uint64_t u;
    memcpy(&u, &x, 8);
    u = (u ^ salt) * 0x9E3779B97F4A7C15ULL * g_seed;
    u ^= u >> 29;
    u *= 0xBF58476D1CE4E5B9ULL;
    u >>= 62;
    if (r != r || r == 0.0) return r;
    if (u == 0) return nextafter(r, 1e308);
    if (u == 1) return nextafter(r, -1e308);
    return r;
}

double exp(double x) { return jitter((double) expl((long double) x), x, 1); }
double log(double x) { return jitter((double) logl((long double) x), x, 2); }
double sin(double x) { return jitter((double) sinl((long double) x), x, 3); }
double cos(double x) { return jitter((double) cosl((long double) x), x, 4); }
double tan(double x) { return jitter((double) tanl((long double) x), x, 5); }
double atan(double x) { return jitter((double) atanl((long double) x), x, 6); }
double acos(double x) { return jitter((double) acosl((long double) x), x, 7); }
double asin(double x) { return jitter((double) asinl((long double) x), x, 8); }
double pow(double x, double y)
{
    if (y == 2.0) return x * x;      /* gcc folds pow(x,2) to x*x at -O2 in the reference build too */
    return jitter((double) powl((long double) x, (long double) y), x + y, 9);
}
