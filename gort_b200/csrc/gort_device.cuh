// gort_device.cuh -- device-side building blocks shared by the GORT kernels (sm_100a, FP64).
//
// Compiled with -fmad=false: every a*b+c written with operators stays an unfused multiply and
// add, in the reference's operation order, so that the geometry / gap-probability code sees the
// same cancellations the reference sees (r*r - h2*h2 must stay exactly 0).
// Where contraction is wanted (the per-wavelength loop) the code calls fma() explicitly.
//
// Reference citations are to /root/reference (tquaife/gort).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/gort_b200.h"

#define GORT_PI 3.14159265358979323846      // M_PI
#define GORT_1_PI 0.31830988618379067154    // M_1_PI
#define GORT_SIN_PI 1.2246467991473532e-16  // sin(M_PI) in FP64, used for the raa = 180 deg pass

namespace gort {

// ---- per-set canopy scalars needed by the BRDF / energy path (gortt.c:641-697) ---------------
struct Canopy {
    double lambda, r, b, h1, h2, favd;
    double rr;        // r*r                       gortt.c:642
    double h;         // 2 r (b/r) + h2 - h1       gortt.c:644
    double ellip;     // b/r                       gortt.c:641
    double elai;      // favd*(1.333333*lambda*pi*ellip*r^3)   gortt.c:657
    double kfavd;     // k*favd with k = 0.5       gortt.c:655,658
    double k_open, k_openep;   // LUT scalars
};

__device__ __forceinline__ Canopy canopy_load(const double* __restrict__ structure, int n_sets, int m,
                                              const double* __restrict__ lut)
{
    Canopy c;
    c.lambda = structure[0 * (size_t) n_sets + m];
    c.r      = structure[1 * (size_t) n_sets + m];
    c.b      = structure[2 * (size_t) n_sets + m];
    c.h1     = structure[3 * (size_t) n_sets + m];
    c.h2     = structure[4 * (size_t) n_sets + m];
    c.favd   = structure[5 * (size_t) n_sets + m];
    c.ellip = c.b / c.r;
    c.rr = c.r * c.r;
    double rrr = c.rr * c.r;
    c.h = 2.0 * c.r * c.ellip + c.h2 - c.h1;
    c.elai = c.favd * ((1.333333) * c.lambda * GORT_PI * c.ellip * rrr);
    c.kfavd = 0.5 * c.favd;
    const double* l = lut + (size_t) m * GORT_LUT_STRIDE;
    c.k_open = l[2 * GORT_NTH];
    c.k_openep = l[2 * GORT_NTH + 1];
    return c;
}

// ---- one input line after the reference's angle preparation (gortt.c:240-291) ------------------
struct Line {
    double vza, vaa, sza, saa, raa;   // radians, normalised
};

__device__ __forceinline__ double dtor(double x) { return x * GORT_PI / 180.0; }

// gortt.c:279 and gortt_albedo.c:97
__device__ __forceinline__ double fold_raa(double raa)
{
    return fabs((raa - 2 * GORT_PI * (int) (0.5 + raa * GORT_1_PI * 0.5)));
}

__device__ __forceinline__ Line line_from_degrees(double vza_d, double vaa_d, double sza_d, double saa_d)
{
    Line g;
    g.vza = dtor(vza_d); g.vaa = dtor(vaa_d); g.sza = dtor(sza_d); g.saa = dtor(saa_d);
    if (g.sza < 0.0) { g.saa += GORT_PI; g.sza *= -1.0; }              // gortt.c:253-260
    if (g.vza < 0.0) { g.vaa += GORT_PI; g.vza *= -1.0; }
    // gortt.c:264-274; bounded so that a non-finite azimuth cannot spin forever
    for (int it = 0; it < 4096 && g.saa > 2 * GORT_PI; it++) g.saa -= 2 * GORT_PI;
    for (int it = 0; it < 4096 && g.vaa > 2 * GORT_PI; it++) g.vaa -= 2 * GORT_PI;
    for (int it = 0; it < 4096 && g.saa < 0; it++) g.saa += 2 * GORT_PI;
    for (int it = 0; it < 4096 && g.vaa < 0; it++) g.vaa += 2 * GORT_PI;
    g.raa = fold_raa(g.saa - g.vaa);                                   // gortt.c:278-279
    return g;
}

// ---- wavelength-independent terms of one (set, line): the "geometry record" --------------------
// Everything gortt_rsurf computes before its wavelength loop (gortt.c:424-449) plus the
// wavelength-independent factors the loop re-derives per band (Kuusk hotspot, t_0, sec terms).
struct GeomRec {
    double Kc, Kg, Kt, Kz, Kpg, Kpz;   // gortt.c:429-449
    double fd;                          // gortt.c:291
    double mus;                         // cos(sza')
    double q;                           // kuusk / (2 cos sza' cos vza')   gortt.c:504-507
    double tp0;                         // p_neq0_heq0_sza + p_ngt0_heq0_sza   gortt_brdf.c:447
    double pe_s;                        // p_ngt0_heq0_sza
    double t0;                          // exp(-k elai / cos sza')   gortt_brdf.c:534
    double pn0_s;                       // p_neq0_heq0_sza (energy balance, gortt_albedo.c:37)
};
#define GORT_REC_FIELDS 13
#define GORT_REC_STRIDE 16      // doubles per packed line record in HBM (128 bytes)

// gortt.c:872-915 for one zenith angle.  The reference indexes p_n0[0][ceil(pos)] with no bound;
// for zenith = 90 deg rounding can give 91 (one past the table): clamp to the last row.
__device__ __forceinline__ void zenith_lerp(const double* __restrict__ lut, double za, double& pn0, double& pe)
{
    const double dth = 1 * GORT_PI / 180.0;       // gortt.c:76
    double pos = fabs(za) / dth;
    int ci = (int) ceil(pos), fi = (int) floor(pos);
    double d = pos - fi;
    ci = min(max(ci, 0), GORT_NTH - 1);
    fi = min(max(fi, 0), GORT_NTH - 1);
    pn0 = d * lut[ci] + (1.0 - d) * lut[fi];
    pe = d * lut[GORT_NTH + ci] + (1.0 - d) * lut[GORT_NTH + fi];
}

#define GORT_MAX(x, y) ((x) > (y) ? (x) : (y))
#define GORT_MIN(x, y) ((x) < (y) ? (x) : (y))

struct Trig {   // trig of the primed zeniths, shared by the three principal-plane passes
    double ts, tv, secs, secv, cs, cv, ss, sv;
};

// gortt_brdf.c:23-100 (live branches only)
__device__ __forceinline__ double overlap_fn(const Canopy& c, const Trig& t, double cr, double sr)
{
    double d = t.ts * t.ts + t.tv * t.tv - 2.0 * t.ts * t.tv * cr;
    double D = sqrt(GORT_MAX(0.0, d));
    double tt = t.ts * t.tv * sr;
    double t2 = sqrt(D * D + tt * tt);
    double t1 = (t.secs + t.secv);
    double cos_t = (c.h / c.b) * t2 / t1;
    cos_t = GORT_MAX(-1.0, cos_t);
    cos_t = GORT_MIN(1.0, cos_t);
    double th = acos(cos_t);
    return GORT_MAX(0.0, (th - sin(th) * cos_t) * (t.secs + t.secv) / GORT_PI);
}

// gortt_brdf.c:7-20
__device__ __forceinline__ double kg_from_overlap(const Canopy& c, const Trig& t, double overlap)
{
    return exp(-(c.lambda * (c.r * c.r) * GORT_PI * (t.secs + t.secv - overlap)));
}

struct CrownShared {   // raa-independent pieces of gortt_kc_fFbeta (gortt_brdf.c:195-207, :223-232)
    double Mv, theta_Mi, Gamma_v, beta;
};

// gortt_brdf.c:171-238 for one relative azimuth
__device__ __forceinline__ void kc_fF(const Canopy& c, const Trig& t, const CrownShared& s,
                                      double sza_p, double vza_p, bool view_gt_sun,
                                      double raa, double cr, double overlap, double Kg,
                                      double& f, double& F)
{
    double phase_prime = t.cv * t.cs + t.sv * t.ss * cr;
    double Gamma = GORT_PI * c.rr * (t.secs + t.secv - overlap);
    double Gamma_c = GORT_PI * c.rr * t.secv * 0.5 * (1.0 + phase_prime);
    F = Gamma_c / Gamma;
    double M = 1.0 - (1.0 - Kg) / (c.lambda * Gamma);
    double PiMi = (1 - cos(s.theta_Mi * (1 - (sza_p - vza_p * cr) / GORT_PI))) / 2.0;
    double PvMv = s.Mv - (1.0 - cos(vza_p * cr - sza_p)) / 2.0;
    double Po;
    if ((raa < dtor(270.)) && (raa > dtor(90.))) Po = PvMv;
    else if (view_gt_sun) Po = PiMi;
    else Po = PvMv;
    f = F * (1.0 - s.Gamma_v * (PvMv + PiMi - Po) / Gamma_c) / (1.0 - M);
}

// Everything wavelength-independent for one line.  vza/sza/raa in radians, already normalised.
__device__ __forceinline__ GeomRec geom_record(const Canopy& c, const double* __restrict__ lut_m,
                                               const gort_options& opt, double vza, double sza, double raa,
                                               double fd)
{
    GeomRec o;
    // gortt.c:283-284 / :424-425, gortt_prime_theta gortt.c:581-588
    double vza_p = atan((c.b / c.r) * tan(vza));
    double sza_p = atan((c.b / c.r) * tan(sza));
    Trig t;
    t.ts = tan(sza_p); t.tv = tan(vza_p);
    sincos(sza_p, &t.ss, &t.cs);
    sincos(vza_p, &t.sv, &t.cv);
    t.secs = 1.0 / t.cs; t.secv = 1.0 / t.cv;

    double pn0_s, pe_s, pn0_v, pe_v;
    zenith_lerp(lut_m, sza, pn0_s, pe_s);
    zenith_lerp(lut_m, vza, pn0_v, pe_v);

    double sr, cr;
    sincos(raa, &sr, &cr);

    // gortt.c:429  Kg at the actual relative azimuth
    double ov = overlap_fn(c, t, cr, sr);
    double Kg = kg_from_overlap(c, t, ov);

    // gortt_kc, gortt_brdf.c:118-169: f,F at raa, then at 0 and 180 degrees
    CrownShared s;
    {
        double xs = c.lambda * GORT_PI * c.rr * t.secs;
        double xv = c.lambda * GORT_PI * c.rr * t.secv;
        double Mi = (1.0 - (1.0 - exp(-xs)) / xs);
        s.Mv = (1.0 - (1.0 - exp(-xv)) / xv);
        s.theta_Mi = acos(1.0 - 2.0 * Mi);
        s.Gamma_v = GORT_PI * c.rr * t.secv;
        if (sza_p < 0.000000001) {
            s.beta = 0.0;
        } else {
            double D = c.r * (1.0 / tan(sza_p / 2.0));
            double lg = c.lambda * s.Gamma_v;
            s.beta = (lg) / (lg + (c.h2 - c.h1) / D) * (1.0 - exp(-lg - (c.h2 - c.h1) / D)) / (1.0 - exp(-lg));
        }
    }
    bool vgs = fabs(vza) > fabs(sza);
    double f, F, f0, F0, f180, F180;
    kc_fF(c, t, s, sza_p, vza_p, vgs, raa, cr, ov, Kg, f, F);
    double ov0 = overlap_fn(c, t, 1.0, 0.0);
    double Kg0 = kg_from_overlap(c, t, ov0);
    kc_fF(c, t, s, sza_p, vza_p, vgs, 0.0, 1.0, ov0, Kg0, f0, F0);
    double ov180 = overlap_fn(c, t, -1.0, GORT_SIN_PI);
    double Kg180 = kg_from_overlap(c, t, ov180);
    kc_fF(c, t, s, sza_p, vza_p, vgs, GORT_PI, -1.0, ov180, Kg180, f180, F180);
    double frac = raa / GORT_PI;
    if (frac > 1.0) frac = 2.0 - frac;
    double beta = opt.use_beta ? opt.beta : s.beta;
    f = (1. - frac) * f0 * F0 + frac * f180 * F180;
    f = beta * f + (1.0 - beta) * F;
    double Kc = f * (1.0 - Kg);

    // gortt.c:439-449
    double e_v = exp(-(c.lambda * GORT_PI * c.rr) / t.cv);
    double e_s = exp(-(c.lambda * GORT_PI * c.rr) / t.cs);
    double Kz = e_v - Kg;
    double Kt = 1.0 - Kc - Kz - Kg;
    Kt = GORT_MAX(0.0, Kt);
    double Kpg = e_s - Kg;
    double Kpz = 1.0 - e_v - Kpg;

    // gortt_kuusk, gortt_brdf.c:638-702 (uses the TRUE zeniths, k = k_vza = 0.5)
    double ssz, csz, svz, cvz;
    sincos(sza, &ssz, &csz);
    sincos(vza, &svz, &cvz);
    double cos_xi = csz * cvz + ssz * svz * cr;
    double lsza = -log(pe_s) / c.kfavd;
    double lvza = -log(pe_v) / c.kfavd;
    double arg = lsza * lsza + lvza * lvza - 2. * lsza * lvza * cos_xi;
    double t1, t2;
    if (arg > 0.0) {
        double lsv = sqrt(arg);
        t2 = (1.0 - exp(-lsv / c.r)) / (lsv / c.r);
    } else {
        t2 = 1.0;
    }
    if ((lsza * lvza) > 0.0) t1 = sqrt(lsza * lvza);
    else t1 = 0.0;
    double H = exp(c.kfavd * t1 * t2);
    double kuusk = pe_s * pe_v * H;

    o.Kc = Kc; o.Kg = Kg; o.Kt = Kt; o.Kz = Kz; o.Kpg = Kpg; o.Kpz = Kpz;
    o.fd = fd;
    o.mus = t.cs;
    o.q = kuusk / (2.0 * t.cs * t.cv);
    o.tp0 = pn0_s + pe_s;
    o.pe_s = pe_s;
    o.t0 = exp(-(0.5 * c.elai * t.secs));
    o.pn0_s = pn0_s;
    return o;
}

// ---- the per-wavelength loop (gortt.c:460-567) split by what each term depends on -------------
// (set, wavelength) only
struct LeafTerms {
    double omega, gam, Tff, Rff, pff, tff, tpff, A, rs, Xf, Zf;
};
// + sun zenith (and fd)
struct SunTerms {
    double G, Z, T, PD, FCf;   // PD = p_df + CdCG ; FCf = (1-fd)*Cf
};

__device__ __forceinline__ LeafTerms leaf_terms(const Canopy& c, double rleaf, double tleaf, double rsoil)
{
    LeafTerms L;
    L.omega = rleaf + tleaf;                                   // gortt.c:469
    L.gam = sqrt(1 - L.omega);                                 // gortt.c:470
    L.rs = rsoil;
    L.Tff = exp(-(2.0 * L.gam * 0.5 * c.elai));                // gortt_brdf.c:492
    L.Rff = (1.0 - L.gam) / (1.0 + L.gam);                     // :574
    double den = 1. - (L.Tff * L.Rff) * (L.Tff * L.Rff);       // :403, :512
    L.pff = L.Rff * (1. - L.Tff * L.Tff) / den;                // :510-512
    L.tff = L.Tff * (1. - L.Rff * L.Rff) / den;                // :401-403
    double K = c.k_open + c.k_openep;                          // :381
    L.tpff = L.tff * (1.0 - K) + K;                            // :382
    double gfunc = -(4.0 / 9.0) * (rleaf - tleaf) / L.omega;   // :591
    L.A = (1.0 - L.omega) * L.omega * (1.0 - gfunc);           // gortt.c:504-506 without kuusk
    L.Xf = (rsoil / (1.0 - rsoil * L.pff)) * (L.tpff - c.k_open);   // gortt.c:520-521
    L.Zf = (L.tpff - c.k_openep) * rsoil;                      // gortt.c:492
    return L;
}

__device__ __forceinline__ SunTerms sun_terms(const Canopy& c, const LeafTerms& L, double fd, double mus,
                                              double t0, double tp0, double pe_s)
{
    SunTerms S;
    double two_mu_g = 2.0 * mus * L.gam;
    double Rdf = (1.0 - L.gam) / (1.0 + two_mu_g);                                   // gortt_brdf.c:552
    double Tdf = (L.omega / 2.0) * ((1. + 2. * mus) / (1. - two_mu_g * two_mu_g)) * (L.Tff - t0);   // :467-471
    double SS = t0 * Rdf + Tdf * L.Rff;
    double tdf = Tdf - L.pff * SS;                                                   // :423-424
    double pdf = Rdf - L.tff * SS;                                                   // :628-630
    double tpdf = tdf * (1 - tp0);                                                   // :361
    double omf = 1 - fd;
    S.G = fd * L.rs + omf * L.rs;                                                    // gortt.c:481-484
    double Zd = (tpdf + pe_s) * L.rs;                                                // gortt.c:491
    S.Z = fd * Zd + omf * L.Zf;                                                      // gortt.c:494
    double K = c.k_open + c.k_openep;
    double CfG = (K * S.G + (1 - K) * S.Z) * c.k_openep;                             // gortt.c:516-517
    double CdCG = (tpdf + tp0) * L.Xf;                                               // gortt.c:519-521
    double CfCG = L.tpff * L.Xf;                                                     // gortt.c:523-525
    double Cf = L.pff + CfG + CfCG;                                                  // gortt.c:512,529
    S.T = fd * CdCG + omf * CfCG;                                                    // gortt.c:541-550
    S.PD = pdf + CdCG;
    S.FCf = omf * Cf;
    return S;
}

// the only part that depends on the view direction: 9 FP64 instructions
__device__ __forceinline__ double view_rsurf(const Canopy& c, const LeafTerms& L, const SunTerms& S,
                                             double fd, double q, double Kc, double Kg, double Kt, double Kz,
                                             double Kpg, double Kpz, double& C_out)
{
    double zg = fma(S.G, Kpg, S.Z * Kpz);                    // Z*K'z + G*K'g          gortt.c:514
    double Cd = fma(L.A, q, S.PD);                           // CdC + CdCG             gortt.c:504-507
    Cd = fma(c.k_openep, zg, Cd);                            // + CdG                  gortt.c:528
    double C = fma(fd, Cd, S.FCf);                           // gortt.c:531
    C_out = C;
    return fma(Kz, S.Z, fma(Kt, S.T, fma(Kg, S.G, Kc * C))); // gortt.c:557
}

}  // namespace gort
