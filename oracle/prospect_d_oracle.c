/* oracle/prospect_d_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement, in plain C, of the reference's PROSPECT-D leaf optics:
 *   PROSPECT-D/prospect_DB.f90:72-191  (subroutine prospect_DB, bind(C) "prospect_DB_")
 *   PROSPECT-D/tav_abs.f90:16-60       (subroutine tav_abs)
 *   PROSPECT-D/dataSpec_PDB.f90:27-1179 (tables, via gort_b200/data/gort_tables.h)
 *
 * PARITY UNPINNED: the reference for this path is Fortran 90 and no Fortran
 * compiler exists in the build image or on the GPU box, and the reference
 * ships no golden vectors.  This file restates the published algorithm with
 * gfortran's arithmetic semantics for the reference's build flags
 * (makefile:3,22-23 -> "-Wall -g", no -fdefault-real-8):
 *   - unsuffixed real literals and DATA constants are REAL(4) and are widened
 *     to REAL(8) on use  ->  tables are stored as binary32 and widened here;
 *   - pi = atan(1.)*4. is single-precision pi (tav_abs.f90:30);
 *   - x**2 (integer exponent) is x*x, x**3 is (x*x)*x, x**2. / x**3. (real
 *     exponent) is pow(); operators associate left to right.
 * Independent cross-check available in tests: the exponential-integral
 * polynomial against scipy.special.exp1.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
 * anything under oracle/.
 */
#include <math.h>
#include <string.h>
#include <stdint.h>
#include "../gort_b200/data/gort_tables.h"

#define NWL GORT_PROSPECT_NW

static double f32bits(uint32_t u)
{
    float f;
    memcpy(&f, &u, sizeof f);
    return (double) f;
}

/* tav_abs.f90:16-60 -- average transmissivity of a dielectric interface for
 * isotropic light inside the cone of half-angle theta (degrees). */
static double oracle_tav_abs(double theta, double nr)
{
    /* tav_abs.f90:30  pi = atan(1.)*4. : evaluated in REAL(4) */
    const float pif = atanf(1.0f) * 4.0f;
    const double pi = (double) pif;
    const double rd = pi / 180.0;                       /* :31 */
    const double n2 = pow(nr, 2.0);                     /* :32 nr**2. */
    const double np = n2 + 1.0;                         /* :33 */
    const double nm = n2 - 1.0;                         /* :34 */
    const double a = ((nr + 1.0) * (nr + 1.0)) / 2.0;   /* :35 */
    const double k = -(((n2 - 1.0) * (n2 - 1.0)) / 4.0);/* :36 */
    const double sa = sin(theta * rd);                  /* :37 */
    double b1;
    if (theta == 90.0) {                                /* :39-43 */
        b1 = 0.0;
    } else {
        b1 = sqrt((sa * sa - np / 2.0) * (sa * sa - np / 2.0) + k);
    }
    const double b2 = sa * sa - np / 2.0;               /* :45 */
    const double b = b1 - b2;                           /* :46 */
    const double b3 = (b * b) * b;                      /* :47 b**3 */
    const double a3 = (a * a) * a;                      /* :48 */
    const double ts = ((pow(k, 2.0) / (6.0 * b3) + k / b) - b / 2.0)
                    - ((pow(k, 2.0) / (6.0 * a3) + k / a) - a / 2.0);   /* :49 */
    const double tp1 = -(((2.0 * n2) * (b - a)) / (np * np));           /* :51 */
    const double tp2 = -((((2.0 * n2) * np) * log(b / a)) / (nm * nm)); /* :52 */
    const double tp3 = (n2 * (1.0 / b - 1.0 / a)) / 2.0;                /* :53 */
    const double tp4 = (((16.0 * pow(n2, 2.0)) * (n2 * n2 + 1.0))
                        * log(((2.0 * np) * b - nm * nm) / ((2.0 * np) * a - nm * nm)))
                       / (pow(np, 3.0) * (nm * nm));                    /* :54 */
    const double tp5 = ((16.0 * pow(n2, 3.0))
                        * (1.0 / ((2.0 * np) * b - nm * nm) - 1.0 / ((2.0 * np) * a - nm * nm)))
                       / ((np * np) * np);                              /* :55 */
    const double tp = (((tp1 + tp2) + tp3) + tp4) + tp5;                /* :56 */
    return (ts + tp) / (2.0 * (sa * sa));                               /* :57 */
}

/* prospect_DB.f90:94-141 -- transmissivity of the elementary layer through
 * the exponential integral (NAG S13AAF polynomials). */
static double oracle_plate_tau(double k)
{
    double xx, yy;
    if (k <= 0.0) return 1.0;                            /* :104-106 */
    if (k <= 4.0) {                                      /* :107-123 */
        xx = 0.5 * k - 1.0;
        yy = (((((((((((((((-3.60311230482612224e-13
            * xx + 3.46348526554087424e-12) * xx - 2.99627399604128973e-11)
            * xx + 2.57747807106988589e-10) * xx - 2.09330568435488303e-9)
            * xx + 1.59501329936987818e-8) * xx - 1.13717900285428895e-7)
            * xx + 7.55292885309152956e-7) * xx - 4.64980751480619431e-6)
            * xx + 2.63830365675408129e-5) * xx - 1.37089870978830576e-4)
            * xx + 6.47686503728103400e-4) * xx - 2.76060141343627983e-3)
            * xx + 1.05306034687449505e-2) * xx - 3.57191348753631956e-2)
            * xx + 1.07774527938978692e-1) * xx - 2.96997075145080963e-1;
        yy = (yy * xx + 8.64664716763387311e-1) * xx + 7.42047691268006429e-1;
        yy = yy - log(k);
        return (1.0 - k) * exp(-k) + (k * k) * yy;
    }
    if (k <= 85.0) {                                     /* :124-138 */
        xx = 14.5 / (k + 3.25) - 1.0;
        yy = (((((((((((((((-1.62806570868460749e-12
            * xx - 8.95400579318284288e-13) * xx - 4.08352702838151578e-12)
            * xx - 1.45132988248537498e-11) * xx - 8.35086918940757852e-11)
            * xx - 2.13638678953766289e-10) * xx - 1.10302431467069770e-9)
            * xx - 3.67128915633455484e-9) * xx - 1.66980544304104726e-8)
            * xx - 6.11774386401295125e-8) * xx - 2.70306163610271497e-7)
            * xx - 1.05565006992891261e-6) * xx - 4.72090467203711484e-6)
            * xx - 1.95076375089955937e-5) * xx - 9.16450482931221453e-5)
            * xx - 4.05892130452128677e-4) * xx - 2.14213055000334718e-3;
        yy = ((yy * xx - 1.06374875116569657e-2) * xx - 8.50699154984571871e-2) * xx
             + 9.23755307807784058e-1;
        yy = (exp(-k) * yy) / k;
        return (1.0 - k) * exp(-k) + (k * k) * yy;
    }
    return 0.0;                                          /* :139-141 */
}

/* One wavelength of prospect_DB.f90:94-189. leaf7 = N,Cab,Car,Anth,Cbrown,Cw,Cm */
static void oracle_prospect_one(const double *leaf7, int i, double *refl, double *tran)
{
    const double N = leaf7[0], Cab = leaf7[1], Car = leaf7[2], Anth = leaf7[3],
                 Cbrown = leaf7[4], Cw = leaf7[5], Cm = leaf7[6];
    const double nr = f32bits(gort_tab_refractive_f32[i]);
    /* :94 */
    const double k = (((((Cab * f32bits(gort_tab_k_cab_f32[i]) + Car * f32bits(gort_tab_k_car_f32[i]))
                         + Anth * f32bits(gort_tab_k_anth_f32[i])) + Cbrown * f32bits(gort_tab_k_brown_f32[i]))
                       + Cw * f32bits(gort_tab_k_cw_f32[i])) + Cm * f32bits(gort_tab_k_cm_f32[i])) / N;
    const double tau = oracle_plate_tau(k);
    const double t12 = oracle_tav_abs(90.0, nr);         /* :145-146 */
    const double talf = oracle_tav_abs(40.0, nr);        /* :147-148 */
    const double ralf = 1.0 - talf;                      /* :149 */
    const double r12 = 1.0 - t12;                        /* :150 */
    const double t21 = t12 / (nr * nr);                  /* :151 */
    const double r21 = 1.0 - t21;                        /* :152 */
    double denom = 1.0 - (r21 * r21) * (tau * tau);      /* :154 */
    const double Ta = ((talf * tau) * t21) / denom;      /* :155 */
    const double Ra = ralf + (r21 * tau) * Ta;           /* :156 */
    const double t = ((t12 * tau) * t21) / denom;        /* :158 */
    const double r = r12 + (r21 * tau) * t;              /* :159 */
    /* Stokes N-layer system, :167-178 */
    const double D = sqrt(((((1.0 + r) + t) * ((1.0 + r) - t)) * ((1.0 - r) + t)) * ((1.0 - r) - t));
    const double rq = r * r;
    const double tq = t * t;
    const double a = (((1.0 + rq) - tq) + D) / (2.0 * r);
    const double b = (((1.0 - rq) + tq) + D) / (2.0 * t);
    const double bNm1 = pow(b, N - 1.0);
    const double bN2 = bNm1 * bNm1;
    const double a2 = a * a;
    denom = a2 * bN2 - 1.0;
    double Rsub = (a * (bN2 - 1.0)) / denom;
    double Tsub = (bNm1 * (a2 - 1.0)) / denom;
    if (r + t >= 1.0) {                                  /* :181-184 zero absorption */
        Tsub = t / (t + (1.0 - t) * (N - 1.0));
        Rsub = 1.0 - Tsub;
    }
    denom = 1.0 - Rsub * r;                              /* :187 */
    *tran = (Ta * Tsub) / denom;                         /* :188 */
    *refl = Ra + (((Ta * Rsub) * t) / denom);            /* :189 */
}

/* Same C binding the Fortran exports (prospect_DB.f90:72, include/gortt.h:292):
 * RT is column-major (2101,2): RT[i] reflectance, RT[i+2101] transmittance. */
void prospect_DB_(double *N, double *Cab, double *Car, double *Anth, double *Cbrown,
                  double *Cw, double *Cm, double *RT)
{
    double leaf7[7] = { *N, *Cab, *Car, *Anth, *Cbrown, *Cw, *Cm };
    for (int i = 0; i < NWL; i++)
        oracle_prospect_one(leaf7, i, &RT[i], &RT[i + NWL]);
}

/* Plain entry point for tests. */
void gort_oracle_prospect_full(const double *leaf7, double *refl2101, double *tran2101)
{
    for (int i = 0; i < NWL; i++)
        oracle_prospect_one(leaf7, i, &refl2101[i], &tran2101[i]);
}

double gort_oracle_tav_abs(double theta_deg, double nr) { return oracle_tav_abs(theta_deg, nr); }
double gort_oracle_plate_tau(double k) { return oracle_plate_tau(k); }
