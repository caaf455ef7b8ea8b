/* fastfmt.h -- "%f" (six decimals) for the gortt output stream, byte-identical to printf.
 *
 * The reference prints every value with printf("%f ") (gortt.c:310-327); a hemispherical sweep at 1 nm is
 * 24.5 million values, 221 MB of text, and glibc's exact-arithmetic printf makes the text the bottleneck of the
 * command line.  For finite |x| < 1e9 the six-decimal rounding is decided from t = fl(x * 1e6): the exact product
 * lies within half an ulp of t, so unless t sits within one ulp of a rounding tie (k + 0.5) the nearest integer to
 * t is the correctly rounded result.  Near a tie, for huge values, NaN and infinity the function falls back to
 * snprintf("%f").  tests/test_fastfmt_cpu.py checks byte equality with printf on 10^7 values incl. exact ties.
 */
#ifndef GORT_FASTFMT_H
#define GORT_FASTFMT_H
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

/* writes the characters of printf("%f", x) to dst (no terminator); returns their number (dst needs >= 330 bytes) */
static inline int gort_fmt_f(char *dst, double x)
{
    const double ax = fabs(x);
    if (ax < 1e9) {                                   /* also false for NaN */
        const double t = ax * 1e6;                    /* < 1e15: integers and halves are exact here */
        const double fl = floor(t);
        const double fr = t - fl;                     /* exact */
        const double ulp = t < 1.0 ? 2.3e-16 : t * 2.3e-16;
        if (fabs(fr - 0.5) > ulp) {
            uint64_t n = (uint64_t) fl + (fr > 0.5 ? 1u : 0u);
            uint64_t ip = n / 1000000u;
            uint32_t fp = (uint32_t) (n - ip * 1000000u);
            char tmp[24];
            int k = 0, len = 0;
            if (signbit(x)) dst[len++] = '-';
            do { tmp[k++] = (char) ('0' + ip % 10u); ip /= 10u; } while (ip);
            while (k) dst[len++] = tmp[--k];
            dst[len++] = '.';
            for (int d = 5; d >= 0; d--) { dst[len + d] = (char) ('0' + fp % 10u); fp /= 10u; }
            return len + 6;
        }
    }
    return snprintf(dst, 330, "%f", x);
}

/* buffered writer over a FILE*: the formatted text goes out in 1 MB blocks */
typedef struct { FILE *fp; size_t n; char buf[1 << 20]; } gort_out;

static inline void gort_out_flush(gort_out *o) { if (o->n) { fwrite(o->buf, 1, o->n, o->fp); o->n = 0; } }
static inline void gort_out_need(gort_out *o, size_t k) { if (o->n + k > sizeof o->buf) gort_out_flush(o); }
static inline void gort_out_str(gort_out *o, const char *s)
{
    size_t k = strlen(s);
    if (k > sizeof o->buf / 2) { gort_out_flush(o); fwrite(s, 1, k, o->fp); return; }
    gort_out_need(o, k);
    memcpy(o->buf + o->n, s, k); o->n += k;
}
/* "%f " */
static inline void gort_out_f(gort_out *o, double x)
{
    gort_out_need(o, 340);
    o->n += (size_t) gort_fmt_f(o->buf + o->n, x);
    o->buf[o->n++] = ' ';
}
#endif
