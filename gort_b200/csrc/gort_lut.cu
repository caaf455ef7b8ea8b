// gort_lut.cu -- KOpen / P(n) gap-probability LUT generation (sm_100a, FP64, -fmad=false).
//
// Replaces gortt_init_params (gortt.c:632-868) + gortt_gap_probabilities
// (gortt_pn_kopen.c:7-129 and everything it calls, :134-924, :1083-1140) and
// gortt_gap_probabilities_Q08 (gortt_pn_kopen.c:1144-1200), restricted to what reaches the
// outputs: p_n0[0][t], epgap[0][t], k_open[0], k_openep[0].  The reference's
// vb / fb / t_open / dt_open / dk_open tables are never read by the BRDF or albedo path and are
// not computed.
//
// Mapping: one CTA per canopy parameter set, one thread per zenith index t (91 of 96 lanes):
//   v_g[h][t] -> p_n0[h][t] (15 layers)  ->  p_s0  ->  E[S]  ->  for every entry height sp_i and
//   crown count n: P(n | s'), within-crown path s and its histogram bin  ->  epgap[0][t];
//   thread 0 then runs the trapezoid rule over theta in the reference's order.
// The path-length histogram pd_s[0][t][bin] is not materialised: epgap[0][t] = sum over bins of
// exp(-s_bin tau') pd_s[bin] is accumulated entry by entry with the entry's bin index, which is
// the same sum in a different association (differences ~1e-16 relative).
//
// Code size matters here: with everything inlined and the Simpson loops unrolled the kernel was 18 000 SASS
// instructions and its top stall was instruction fetch (stall_no_instruction ~70 % of samples in the geometry
// code, ncu profiles/r1_lut).  The geometry helpers are therefore __noinline__ and their loops not unrolled.
//
// Parity hazards honoured: no FMA contraction (r*r - h2*h2 must be exactly
// 0 when h2 == r), running-sum loop counters (z += dz_p, h += dh), (int)(s/ds + 0.5) binning,
// the |a3| < 1e-10 clamp, float-typed Simpson factors.
#include "gort_device.cuh"
#include "gort_internal.h"

namespace gort {

#define LUT_THREADS 96
#define LUT_MAXCROWNS 30
#define LUT_NH_ES 20
#define LUT_NOINT 20

struct Crown {
    double r, rr, rrr;
    double h1_p, h2_p, z2_p, dz_p, ds, lv_p, tau_p;
};

struct Ang {            // trig of one theta_p
    double th, s, c, t;  // angle, sin, cos, tan
};

// gortt_pn_kopen.c:285-305
__device__ __forceinline__ double left_circle_area(double r, double x_cut)
{
    double area_tot = GORT_PI * r * r;
    double ang_sector = acos(fabs(x_cut) / r) * 2.0;
    double area_sector = area_tot * ang_sector / (2.0 * GORT_PI);
    double area_triangle = fabs(x_cut) * sqrt(r * r - x_cut * x_cut);
    if (x_cut > 0.0) return area_tot - (area_sector - area_triangle);
    return area_sector - area_triangle;
}

// gortt_pn_kopen.c:309-323
__device__ __forceinline__ double right_ellipse_area(double r, double b, double x_cut)
{
    double x_cut_p = x_cut / (b / r);
    double a_p = GORT_PI * r * r;
    a_p -= left_circle_area(r, x_cut_p);
    return a_p * (b / r);
}

// gortt_pn_kopen.c:170-229 with the "weird" section :233-282 inlined
__device__ __noinline__ double cross_section(const Crown& c, const Ang& a, double h, double z)
{
    if (z < h - c.r) return 0.0;
    double h_low = h - c.r * a.s;
    double h_high = h + c.r * a.s;
    if (z <= h_low) {
        double q = c.rr - (h - z) * (h - z);
        double r_p = (q <= 0) ? 0 : sqrt(q);
        return GORT_PI * r_p * r_p;
    } else if (z > h_low && z < h_high) {
        double zdiff = h - z;
        double r_p = sqrt(c.rr - zdiff * zdiff);
        double x_cc = zdiff * a.t;
        double x_p = x_cc / (1.0 - a.c * a.c);
        double a_cp = left_circle_area(r_p, x_p - x_cc);
        double a_ep = right_ellipse_area(c.r, c.r * (1.0 / a.c), x_p);
        return a_cp + a_ep;
    }
    return GORT_PI * c.rr * (1.0 / a.c);
}

// gortt_pn_kopen.c:149-167
__device__ __noinline__ double proj_volume(const Crown& c, const Ang& a, double h)
{
    double vol = 0.0;
    int guard = 0;
    for (double z = c.h1_p + c.dz_p / 2.0; z <= c.h2_p && guard < 100000; z += c.dz_p, guard++)
        vol += cross_section(c, a, h, z) * (c.dz_p);
    return vol;
}

// gortt_pn_kopen.c:858-872
__device__ __forceinline__ double triang_fcn(double x, double b, double r, double tan_the)
{
    double a1 = tan_the * (x - b);
    double a2 = r * r - x * x;
    double a3 = a2 - a1 * a1;
    if (fabs(a3) < 0.0000000001) a3 = 0.0;
    return 2.0 * a1 * sqrt(a3);
}

// gortt_pn_kopen.c:811-854
__device__ __noinline__ double triang(double b, double r, const Ang& a)
{
    double sint = a.s, cost = a.c;
    double a1 = r * r - b * b * sint * sint;
    double x0 = b * (sint * sint) + sqrt(a1) * cost;
    const int m = LUT_NOINT;
    double h = .50 * (x0 - b) / (double) (float) m;
    double sum1 = 0.0;
#pragma unroll 4
    for (int i = 0; i < m; i++) sum1 += triang_fcn(b + (double) (float) (2 * i + 1) * h, b, r, a.t);
    double volume = 4.0 * sum1;
    double sum2 = 0.0;
#pragma unroll 4
    for (int i = 0; i < m - 1; i++) sum2 += triang_fcn(b + (double) (float) (2 * (i + 1)) * h, b, r, a.t);
    volume += 2.0 * sum2;
    volume += triang_fcn(x0, b, r, a.t);
    volume += triang_fcn(b, b, r, a.t);
    volume *= h / 3.0;
    return volume;
}

// gortt_pn_kopen.c:796-806
__device__ __forceinline__ double sector(double a1, double a2, double r)
{
    double b1 = r * r * a1 - (a1 * a1 * a1) / 3.0;
    double b2 = r * r * a2 - (a2 * a2 * a2) / 3.0;
    return GORT_PI * (b2 - b1) / 2.0;
}

// gortt_pn_kopen.c:771-792
__device__ __noinline__ double trisec(double hh, double hh_b, const Ang& a, double r)
{
    double tmp = (hh - hh_b);
    double x = -1.0 * tmp * a.s + sqrt(r * r - tmp * tmp) * a.c;
    double b = -tmp / a.s;
    return triang(b, r, a) + sector(x, r, r);
}

// gortt_pn_kopen.c:876-886
__device__ __forceinline__ double cylind_fcn(double x, double r)
{
    return .50 * x * sqrt(r * r - x * x) + .50 * r * r * asin(x / r);
}

// gortt_pn_kopen.c:891-924
__device__ __noinline__ double cylind(double r, double h1, double h2, double h)
{
    double slope = h / (h2 - h1);
    double tmp1 = sqrt(r * r - h1 * h1);
    double tmp2 = sqrt(r * r - h2 * h2);
    double volume = tmp1 * tmp1 * tmp1 - tmp2 * tmp2 * tmp2;
    volume /= 3.0;
    volume -= h1 * (cylind_fcn(h2, r) - cylind_fcn(h1, r));
    volume *= 2.0 * slope;
    if (h2 < r) {
        double phi = acos(h2 / r);
        double s1 = r * r * phi;
        double s2 = r * sin(phi) * h2;
        volume += (s1 - s2) * h;
    }
    return volume;
}

// gortt_pn_kopen.c:665-768; hp_h = height_p[h], hp_s = height_p[h_s]
__device__ __noinline__ double tube_vol(const Crown& c, const Ang& a, double hp_h, double hp_s, double h_b)
{
    const double r = c.r;
    double V, V_sp1, V_sp2, V_cyln, h_t, h_tt;
    double tmp_s = (hp_s - hp_h) / a.c;
    double V_0 = GORT_PI * c.rr * tmp_s;
    V_0 += (4.0 / 3.0) * GORT_PI * c.rrr;

    if ((hp_h - r) >= h_b) {
        V = 0.0;
    } else if ((hp_h - r * a.s) >= h_b) {
        h_t = r - (hp_h - h_b);
        V = (GORT_PI / 3.0) * h_t * h_t * (3.0 * r - h_t);
    } else if ((hp_h + r * a.s) >= h_b) {
        V_sp1 = (2.0 / 3.0) * GORT_PI * c.rrr;
        V_sp1 -= trisec(hp_h, h_b, a, r);
        h_tt = (h_b - (hp_h - r * a.s)) / a.c;
        if (hp_s - r * a.s >= h_b) {
            double hh1 = (hp_h - h_b) / a.s;
            V_cyln = cylind(r, hh1, r, h_tt);
            V_sp2 = 0.0;
        } else {
            double hh1 = (hp_h - h_b) / a.s;
            double hh2 = (hp_s - h_b) / a.s;
            double hh = (hp_s - hp_h) / a.c;
            V_cyln = cylind(r, hh1, hh2, hh);
            V_sp2 = trisec(h_b, hp_s, a, r);
        }
        V = V_sp1 + V_cyln + V_sp2;
    } else if (hp_s - r * a.s >= h_b) {
        double tmp_h = (h_b - hp_h) / a.c;
        V_cyln = GORT_PI * r * r * tmp_h;
        V_sp1 = (2.0 / 3.0) * GORT_PI * c.rrr;
        V = V_sp1 + V_cyln;
    } else if (hp_s + r * a.s >= h_b) {
        h_tt = (hp_s + r * a.s - h_b) / a.c;
        double hh1 = (h_b - hp_s) / a.s;
        double tmp_h = (hp_s - hp_h) / a.c;
        V_cyln = GORT_PI * r * r * tmp_h - cylind(r, hh1, r, h_tt);
        V_sp2 = trisec(h_b, hp_s, a, r);
        V_sp1 = (2.0 / 3.0) * GORT_PI * c.rrr;
        V = V_cyln + V_sp2 + V_sp1;
    } else if (hp_s + r >= h_b) {
        h_t = r - (h_b - hp_s);
        V_sp1 = (GORT_PI / 3.0) * h_t * h_t * (3.0 * r - h_t);
        V = V_0 - V_sp1;
    } else {
        V = V_0;
    }
    return V;
}

// gortt_pn_kopen.c:566-645; hz = height_p[z]
__device__ __noinline__ double mean_single_crown_path(const Crown& c, const Ang& a, double hz, double h)
{
    if (hz > h + c.r - 0.0001) return 0.0;
    if (hz < h - c.r + 0.0001) return 4.0 * c.r / 3.0;
    double V_sphere = 4.0 * GORT_PI * c.rrr / 3.0;
    double zdiff = fabs(h - hz);
    double ht = c.r - zdiff;
    double V_slice = GORT_PI * ht * ht / 3.0 * (3.0 * c.r - ht);
    double V_tot = (hz > h) ? V_slice : V_sphere - V_slice;
    V_tot /= a.c;
    double proj_area;
    if (h < hz) proj_area = cross_section(c, a, h, (h - zdiff));
    else proj_area = cross_section(c, a, h, (h + zdiff));
    return V_tot / proj_area;
}

// gortt_pn_kopen.c:534-563
__device__ double expected_single_crown_path(const Crown& c, const Ang& a, double hz)
{
    double ES = 0.0;
    double dh = (c.h2_p - c.h1_p) / (double) LUT_NH_ES;
    int guard = 0;
    for (double h = c.h1_p + dh / 2.0; h <= c.h2_p && guard < 100000; h += dh, guard++)
        ES += mean_single_crown_path(c, a, hz, h) * ((1.0 / (c.h2_p - c.h1_p)) * dh);
    return ES;
}

// lut_full_kernel: one CTA per GROUP of consecutive parameter sets that share (r, b, h1, h2) bit for bit (capped at
// group_cap <= LUT_GROUP_CAP sets, chunk boundaries at multiples of the cap).
//
// Work decomposition (round 2).  The reference's nest is  zenith t (91) x entry height sp_i (13) x crown count n (30)
// (gortt_pn_kopen.c:457-527).  Round 1 gave every zenith ONE thread that walked the 13 entry heights serially and kept
// v_g[15], s'[13] and the 13 tube-volume differences in registers (96 registers, 3 warps per CTA, 22.7 % of the SM's
// warp slots, FP64 pipe 45 % busy, 98 % of the EnKF member update).  Now a CTA is 96 x LUT_NK threads: lane = zenith
// (neighbouring zeniths take the same branches of the sphere / cylinder geometry), and the entry heights -- and the
// 14 + K distinct cross-sections of phase 1 -- are dealt round-robin to the LUT_NK thread rows; the per-zenith arrays
// live in shared memory, so a thread carries one entry height at a time (<= 80 registers, two 384-thread CTAs per SM).
//   phase 1, once per group: everything that depends on crown shape and zenith only -- the projected cross-section
//            volumes v_g[h][t] (gortt_pn_kopen.c:24-32, :149-323), E[S] (:534-563), and for every entry height the
//            tube-volume difference of :496 (Simpson rule, sphere/cylinder sections);
//   phase 2, per sub-group of up to SUB members that also share the stem density: p_n0 = exp(-lv' v_g), then the
//            crown-count loop (:489-527) once per entry height; per member only the within-crown gap sums.
// The bin attenuation exp(-s_bin tau') of gortt_calc_epgap (:1110-1114) depends on (set, bin) only -- not on zenith,
// entry height or crown count -- so it is TABULATED once per member (LUT_TAB bins, one exp each, the literal formula)
// instead of being re-evaluated at every change of bin inside the crown-count loop (~10 exp per (zenith, entry height)
// before: the largest single item of a set with its own crown shape).
// Each thread row accumulates its entry heights in the reference's order and the LUT_NK partial sums are added in a fixed
// order: the same terms as the reference's bin-by-bin sum in a different association (~1e-16 relative).
#define LUT_GROUP_CAP 64
#define LUT_NK 4                        // thread rows per CTA (entry heights / cross-sections dealt round-robin)
#define LUT_CTA (LUT_THREADS * LUT_NK)
#define LUT_TAB 512                     // tabulated bins of exp(-s_bin tau'); bins beyond use the formula directly
#define LUT_NA 32                       // distinct cross-sections per zenith: 14 + K, K < 16
// 1/n!, n = 0..30, each the FP64 quotient 1.0 / n! (the reference tabulates n! in gortt.c:752-754 and divides)
__constant__ double c_inv_fact[LUT_MAXCROWNS + 1] = {
    1.0,
    1.0,
    0.5,
    0.16666666666666666,
    0.041666666666666664,
    0.008333333333333333,
    0.001388888888888889,
    0.0001984126984126984,
    2.48015873015873e-05,
    2.7557319223985893e-06,
    2.755731922398589e-07,
    2.505210838544172e-08,
    2.08767569878681e-09,
    1.6059043836821613e-10,
    1.1470745597729725e-11,
    7.647163731819816e-13,
    4.779477332387385e-14,
    2.8114572543455206e-15,
    1.5619206968586225e-16,
    8.22063524662433e-18,
    4.110317623312165e-19,
    1.9572941063391263e-20,
    8.896791392450574e-22,
    3.8681701706306835e-23,
    1.6117375710961184e-24,
    6.446950284384474e-26,
    2.4795962632247972e-27,
    9.183689863795546e-29,
    3.2798892370698385e-30,
    1.1309962886447718e-31,
    3.769987628815906e-33};
#define LUT_SUB 8                       // members per sub-group (same shape AND same stem density)
#define LUT_NSP (GORT_NLAYERS - 2)     // entry heights sp_i = 1 .. 13 (sp_i = 14 contributes p_s0 = 0)

__device__ __forceinline__ bool same_shape(const double* __restrict__ st, size_t n, int a, int b)
{
    return st[1 * n + a] == st[1 * n + b] && st[2 * n + a] == st[2 * n + b] &&
           st[3 * n + a] == st[3 * n + b] && st[4 * n + a] == st[4 * n + b];
}

template <int SUB>
struct LutSmem {
    double hp[GORT_NLAYERS];                 // height_p
    double zk[16];                           // crown-centre heights of the midpoint rule
    int K;
    double ang[4][LUT_THREADS];              // theta_p, sin, cos, tan per zenith
    double es[LUT_THREADS];                  // E[S] per zenith
    union {                                  // A is dead once v_g is summed (a barrier before the first use of part)
        double A[LUT_NA][LUT_THREADS];           // distinct cross-sections
        double part[LUT_NK][SUB][LUT_THREADS];   // partial within-crown gap sums per thread row
    };
    double vg[GORT_NLAYERS][LUT_THREADS];    // v_g[h][t]
    double pn0[GORT_NLAYERS][LUT_THREADS];   // p_n0[h][t] of the current sub-group
    double tube[LUT_NSP][LUT_THREADS];       // tube-volume difference per entry height
    double tab[SUB][LUT_TAB];                // exp(-s_bin tau') per member of the current sub-group
};

template <int SUB>
__global__ void __launch_bounds__(LUT_CTA, 2)
lut_full_kernel(int n_sets, int group_cap, const double* __restrict__ structure, double* __restrict__ lut)
{
    extern __shared__ __align__(16) unsigned char lut_smem_raw[];
    LutSmem<SUB>& sm = *reinterpret_cast<LutSmem<SUB>*>(lut_smem_raw);
    const int m0 = blockIdx.x;
    const int tid = threadIdx.x;
    const int t = tid % LUT_THREADS;             // zenith index (lanes of a warp: consecutive zeniths)
    const int kk = tid / LUT_THREADS;            // thread row
    const size_t N = (size_t) n_sets;
    // group heads: a set whose crown shape differs from its predecessor's, or that sits on a chunk boundary
    if (m0 > 0 && (m0 % group_cap) != 0 && same_shape(structure, N, m0, m0 - 1)) return;
    int m1 = m0 + 1;
    while (m1 < n_sets && (m1 % group_cap) != 0 && same_shape(structure, N, m1, m1 - 1)) m1++;
    // two instantiations share the work: SUB = 1 takes the groups whose members all differ in stem density -- in
    // particular every single-set group --, SUB = LUT_SUB takes the groups that start with a sub-group (same stem
    // density, favd varying)
    {
        const bool shares = (m1 - m0 >= 2) && structure[0 * N + m0] == structure[0 * N + m0 + 1];
        if ((SUB == 1) == shares) return;
    }

    // ---- gortt_init_params, gortt.c:641-697: the shape-only part ------------------------------------
    const double r      = structure[1 * N + m0];
    const double b      = structure[2 * N + m0];
    const double h1     = structure[3 * N + m0];
    const double h2     = structure[4 * N + m0];
    const double ellip = b / r;
    Crown c;
    c.r = r; c.rr = r * r; c.rrr = c.rr * r;
    const double z1 = h1 - r * ellip;
    const double z2 = h2 + r * ellip;
    c.z2_p = z2 / ellip;
    c.h1_p = h1 / ellip;
    c.h2_p = h2 / ellip;
    const double dz = (double) (z2 - z1) / ((double) GORT_NLAYERS - 1.0);
    c.ds = dz;
    c.dz_p = dz / ellip;
    c.lv_p = 0.0; c.tau_p = 0.0;                 // per member, below
    if (tid < GORT_NLAYERS) {                                                    // gortt.c:778-781
        double height = z2 - dz * (double) (GORT_NLAYERS - 1 - tid);
        sm.hp[tid] = height / ellip;
    }
    if (tid == LUT_THREADS) {
        // crown-centre heights of the midpoint rule, gortt_pn_kopen.c:162: a running sum
        int K = 0;
        for (double z = c.h1_p + c.dz_p / 2.0; z <= c.h2_p && K < 16; z += c.dz_p) sm.zk[K++] = z;
        sm.K = K;
    }
    const double dth = 1 * GORT_PI / 180.0;
    if (kk == LUT_NK - 1 && t < GORT_NTH) {
        double theta = dth * (double) t;                                         // gortt.c:783-797
        if (theta >= GORT_PI / 2.0) theta = GORT_PI / 2.0 - 1.0 * GORT_PI / 180.0;
        double th = atan(tan(theta) * ellip);
        if (th >= GORT_PI / 2.0) th = GORT_PI / 2.0 - 1.0 * GORT_PI / 180.0;
        sm.ang[0][t] = th; sm.ang[1][t] = sin(th); sm.ang[2][t] = cos(th); sm.ang[3][t] = tan(th);
    }
    __syncthreads();
    Ang a;
    a.th = a.s = a.c = a.t = 0.0;
    const bool live = t < GORT_NTH;               // 91 of 96 lanes
    const bool path = t < GORT_NTH - 1;           // epgap only for t < nth - 1, gortt_pn_kopen.c:1099
    if (live) { a.th = sm.ang[0][t]; a.s = sm.ang[1][t]; a.c = sm.ang[2][t]; a.t = sm.ang[3][t]; }
    const int K = sm.K;
    const bool tabulated = K >= 1 && K < 16;
    const double hp0 = sm.hp[0];

    // ---- phase 1 -----------------------------------------------------------------------------------
    if (live) {
        // v_g[h][t], gortt_pn_kopen.c:29, :149-167: midpoint rule over the crown-centre height z of the projected
        // cross-section of a crown centred at z seen from layer height h.  The cross-section depends on h - z only,
        // the layer heights and the midpoints are both dz' apart, so the 15 x K evaluations take only 14 + K distinct
        // values: each is evaluated once (at its first (h, z) pair) and the 15 sums are formed in the reference's
        // order.  (The other pairs differ from it by the rounding of h - z, ~1e-16.)
        if (tabulated) {
#pragma unroll 1
            for (int j = kk; j < GORT_NLAYERS + K - 1; j += LUT_NK) {
                const int i = max(0, j - (K - 1)), k = i - (j - (K - 1));
                sm.A[j][t] = cross_section(c, a, sm.hp[i], sm.zk[k]);
            }
        } else {
#pragma unroll 1
            for (int h = kk; h < GORT_NLAYERS; h += LUT_NK) sm.vg[h][t] = proj_volume(c, a, sm.hp[h]);
        }
        if (path) {
            if (kk == 0) sm.es[t] = expected_single_crown_path(c, a, hp0);      // :445
#pragma unroll 1
            for (int k = kk; k < LUT_NSP; k += LUT_NK) {
                const double hps = sm.hp[GORT_NLAYERS - 2 - k];                  // :457, sp_i = 13 down to 1
                sm.tube[k][t] = tube_vol(c, a, hp0, hps, c.h2_p) - tube_vol(c, a, hp0, hps, c.h1_p);   // :496
            }
        }
    }
    __syncthreads();
    if (live && tabulated) {
#pragma unroll 1
        for (int h = kk; h < GORT_NLAYERS; h += LUT_NK) {
            double vol = 0.0;
            for (int k = 0; k < K; k++) vol += sm.A[h - k + K - 1][t] * (c.dz_p);
            sm.vg[h][t] = vol;
        }
    }
    // (the barrier that publishes vg is the first one of the member loop)

    // ---- members of the group, in sub-groups of up to SUB consecutive members that also share the stem
    //      density: p_n0, P(n) and the bin sequence depend on lambda only, favd enters through the bin attenuation
    //      exp(-s_bin tau') alone (gortt_pn_kopen.c:1110-1114), so the crown-count loop runs once per sub-group
    //      and only the sums are per member ----
    const double inv_ds = 1.0 / c.ds;
    for (int ms = m0; ms < m1;) {
        const double lambda = structure[0 * N + ms];
        int nj = 1;
        while (nj < SUB && ms + nj < m1 && structure[0 * N + ms + nj] == lambda) nj++;
        const double lv = lambda / (h2 - h1);
        const double lv_p = lv * ellip;
        __syncthreads();                                 // vg complete; previous sub-group done with pn0 / tab / part
        // bin attenuations of the sub-group's members: exp(-s_bin tau'), s_bin = bin * ds, tau' = k favd'  (:1110-1114)
        for (int i = tid; i < nj * LUT_TAB; i += LUT_CTA) {
            const int j = i / LUT_TAB, bin = i - j * LUT_TAB;
            const double favd_p = structure[5 * N + ms + j] * ellip;
            const double sbin = (double) bin * c.ds;
            sm.tab[j][bin] = exp(-sbin * (0.5 * favd_p));
        }
        if (live) {
#pragma unroll 1
            for (int h = kk; h < GORT_NLAYERS; h += LUT_NK) sm.pn0[h][t] = exp(-1.0 * lv_p * sm.vg[h][t]);   // gortt_pn_kopen.c:30
        }
        __syncthreads();
        double e_t[SUB], tau[SUB];
#pragma unroll
        for (int j = 0; j < SUB; j++) {
            e_t[j] = 0.0;
            tau[j] = 0.5 * (structure[5 * N + ms + (j < nj ? j : 0)] * ellip);
        }
        if (path) {
            const double es = sm.es[t];
#pragma unroll 1
            for (int k = kk; k < LUT_NSP; k += LUT_NK) {
                const int sp_i = GORT_NLAYERS - 2 - k;
                const double P_s_p = sm.pn0[sp_i + 1][t] - sm.pn0[sp_i][t];      // :43, :482
                const double temp1 = sm.tube[k][t] * lv_p;                       // :497
                const double E = exp(-temp1);
                const double sp = (double) (sm.hp[sp_i] - hp0) / a.c;            // :464
                // crown-count loop, gortt_pn_kopen.c:489-527, with its loop invariants hoisted:
                //   P(n) = temp1^n e^-temp1 / (n! (1 - e^-temp1))             :501-502
                //   s    = s' (1 - exp(-n E[S]/s'))                            :508
                // exp(-n x) is advanced as q^n (q = exp(-x)); because s selects a histogram bin through
                // (int)(s/ds + 0.5) (:134-139, :522) the literal exp is evaluated instead whenever the
                // product form (with s/ds as s * (1/ds)) lands within 1e-9 of a bin boundary, so the bin is
                // always the one the literal formula gives.
                const double c0 = E / (1.0 - E);
                const double x = es / sp;
                const double q = exp(-x);
                double pw = 1.0, qn = 1.0;
#pragma unroll 1
                for (int n = 1; n <= LUT_MAXCROWNS; n++) {                       // :489
                    pw *= temp1;                                                 // temp1^n
                    qn *= q;
                    const double P_n = pw * c0 * c_inv_fact[n];
                    double u = sp * (1.0 - qn) * inv_ds + 0.5;
                    if (fabs(u - rint(u)) < 1e-9)
                        u = sp * (1.0 - exp(-1.0 * (double) n * es / sp)) / c.ds + 0.5;
                    const int idx = (int) u;
                    const double wgt = P_n * P_s_p;
                    // gortt_calc_epgap + gortt_calc_pgap, :1110-1114, :1138
                    if (idx >= 0 && idx < LUT_TAB) {
#pragma unroll
                        for (int j = 0; j < SUB; j++) e_t[j] += sm.tab[j][idx] * wgt;
                    } else {
                        const double sbin = (double) idx * c.ds;
#pragma unroll
                        for (int j = 0; j < SUB; j++) e_t[j] += exp(-sbin * tau[j]) * wgt;
                    }
                }
            }
        }
        if (live) {
#pragma unroll
            for (int j = 0; j < SUB; j++) sm.part[kk][j][t] = e_t[j];
        }
        __syncthreads();
        // the openness factors (trapezoid rule over the 91 zeniths) are formed by kopen_kernel afterwards
        if (live) {
            for (int j = kk; j < nj; j += LUT_NK) {
                double e = sm.part[0][j][t];
#pragma unroll
                for (int q2 = 1; q2 < LUT_NK; q2++) e += sm.part[q2][j][t];
                double* o = lut + (size_t) (ms + j) * GORT_LUT_STRIDE;
                o[t] = sm.pn0[0][t];
                o[GORT_NTH + t] = e;
            }
        }
        ms += nj;
    }
}

// gortt_pn_kopen.c:1144-1200
__global__ void __launch_bounds__(LUT_THREADS)
lut_q08_kernel(int n_sets, const double* __restrict__ structure, double* __restrict__ lut)
{
    __shared__ double s_pn0[GORT_NTH];
    __shared__ double s_epg[GORT_NTH];
    __shared__ double s_sin2[GORT_NTH];
    const int m = blockIdx.x;
    const int t = threadIdx.x;
    const double lambda = structure[0 * (size_t) n_sets + m];
    const double r      = structure[1 * (size_t) n_sets + m];
    const double b      = structure[2 * (size_t) n_sets + m];
    const double favd   = structure[5 * (size_t) n_sets + m];
    const double ellip = b / r;
    const double rr = r * r;
    const double dth = 1 * GORT_PI / 180.0;
    if (t < GORT_NTH) {
        double theta = dth * (double) t;
        if (theta >= GORT_PI / 2.0) theta = GORT_PI / 2.0 - 1.0 * GORT_PI / 180.0;
        double theta_p = atan(tan(theta) * ellip);
        if (theta_p >= GORT_PI / 2.0) theta_p = GORT_PI / 2.0 - 1.0 * GORT_PI / 180.0;
        double cc = GORT_PI * rr * lambda;                               // :1164
        double l = favd * b * 4. / 3. * cc;                              // :1166
        double k2 = 0.348535 * pow(cc, (-1.08069 - 0.0874595 * cc));     // :1168
        double k1 = 0.0014166;
        double a = cc * (exp(k1 * cc * cc) - exp(-k2 * l));              // :1171
        double pn0 = exp(-cc / (cos(theta_p)));                          // :1185
        double epg = exp(-a / (cos(theta_p))) - pn0;                     // :1186
        s_pn0[t] = pn0; s_epg[t] = epg; s_sin2[t] = sin(2.0 * theta);
        double* o = lut + (size_t) m * GORT_LUT_STRIDE;
        o[t] = pn0;
        o[GORT_NTH + t] = epg;
    }
    __syncthreads();
    if (t == 0) {
        double ko = 0.0, ke = 0.0;
        // :1176-1177 read rows that are still calloc'd zeros
        double tmp1_last = 0.0 * s_sin2[0], tmp2_last = 0.0 * s_sin2[0];
        for (int i = 1; i < GORT_NTH; i++) {
            double tmp1 = s_pn0[i] * s_sin2[i];
            ko += (tmp1 + tmp1_last) / 2.0 * dth;
            tmp1_last = tmp1;
            double tmp2 = s_epg[i] * s_sin2[i];
            ke += (tmp2 + tmp2_last) / 2.0 * dth;
            tmp2_last = tmp2;
        }
        double* o = lut + (size_t) m * GORT_LUT_STRIDE;
        o[2 * GORT_NTH] = ko;
        o[2 * GORT_NTH + 1] = ke;
    }
}

// gortt_calc_kopen, gortt_pn_kopen.c:351-375 for h = 0: k_open = trapezoid rule of p_n0 sin(2 theta) over the 91
// zeniths, k_openep the same for epgap.  One warp per parameter set: the panels (f_i + f_{i-1})/2 * dth, i = 1..90,
// three per lane, then a shuffle tree (the reference adds them left to right: same panels, different association).
__global__ void __launch_bounds__(128)
kopen_kernel(int n_sets, double* __restrict__ lut)
{
    const int m = (int) (((long) blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (m >= n_sets) return;
    const int lane = threadIdx.x & 31;
    const double dth = 1 * GORT_PI / 180.0;
    double* o = lut + (size_t) m * GORT_LUT_STRIDE;
    double ko = 0.0, ke = 0.0;
    for (int i = 1 + lane; i < GORT_NTH; i += 32) {
        double th1 = dth * (double) i, th0 = dth * (double) (i - 1);                 // gortt.c:783-787
        if (th1 >= GORT_PI / 2.0) th1 = GORT_PI / 2.0 - 1.0 * GORT_PI / 180.0;
        if (th0 >= GORT_PI / 2.0) th0 = GORT_PI / 2.0 - 1.0 * GORT_PI / 180.0;
        const double s1 = sin(2.0 * th1), s0 = sin(2.0 * th0);
        ko += (o[i] * s1 + o[i - 1] * s0) / 2.0 * dth;
        ke += (o[GORT_NTH + i] * s1 + o[GORT_NTH + i - 1] * s0) / 2.0 * dth;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ko += __shfl_xor_sync(0xffffffffu, ko, off);
        ke += __shfl_xor_sync(0xffffffffu, ke, off);
    }
    if (lane == 0) { o[2 * GORT_NTH] = ko; o[2 * GORT_NTH + 1] = ke; }
}

int launch_lut(gort_ctx *ctx, cudaStream_t s, int n_sets, const double *structure, int method, double *lut)
{
    note_other_work(ctx);
    if (method == GORT_LUT_Q08) lut_q08_kernel<<<n_sets, LUT_THREADS, 0, s>>>(n_sets, structure, lut);
    else {
        // group cap: as large as possible (phase 1 is shared by the whole group) while the batch still yields
        // enough groups to fill the GPU a few times over; results do not depend on it
        int cap = n_sets / (ctx->sm_count * 12);
        if (cap < 1) cap = 1;
        if (cap > LUT_GROUP_CAP) cap = LUT_GROUP_CAP;
        static_assert(sizeof(LutSmem<LUT_SUB>) <= 113 * 1024, "two LUT CTAs per SM");
        if (!ctx->lut_attr_set) {
            cudaError_t e = cudaFuncSetAttribute(lut_full_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LutSmem<1>));
            if (e == cudaSuccess) e = cudaFuncSetAttribute(lut_full_kernel<LUT_SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(LutSmem<LUT_SUB>));
            if (e != cudaSuccess) return check_cuda(ctx, e, "lut_full_kernel shared memory");
            ctx->lut_attr_set = 1;
        }
        lut_full_kernel<1><<<n_sets, LUT_CTA, sizeof(LutSmem<1>), s>>>(n_sets, cap, structure, lut);
        lut_full_kernel<LUT_SUB><<<n_sets, LUT_CTA, sizeof(LutSmem<LUT_SUB>), s>>>(n_sets, cap, structure, lut);
        kopen_kernel<<<(unsigned) (((long) n_sets * 32 + 127) / 128), 128, 0, s>>>(n_sets, lut);
        ctx->launches += 2;
    }
    ctx->launches++;
    return check_cuda(ctx, cudaGetLastError(), "gort_lut launch");
}

}  // namespace gort
