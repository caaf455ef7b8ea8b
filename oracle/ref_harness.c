/* oracle/ref_harness.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Thin in-memory driver over the UNMODIFIED reference C sources, which are
 * compiled where they lie under /root/reference by oracle/Makefile into
 * oracle/_ref/libgortt_ref.so (never copied into this repo).  The harness
 * only fills the reference's own structs (include/gortt.h:41-212), calls the
 * reference's own functions and copies doubles out, so results can be compared
 * in memory instead of through the "%f" text of gortt.c:310-324.
 *
 * The only arithmetic restated here is the per-line angle preparation that the
 * reference performs inside main() (gortt.c:240-291) and cannot be called
 * separately; the CLI text tests pin that part against the real binary.
 *
 * PROSPECT-D is Fortran in the reference; prospect_DB_ is supplied by
 * oracle/prospect_d_oracle.c (no Fortran compiler in this image).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <gortt.h>

extern double default_soil_vector_1[], default_soil_vector_2[], default_soil_vector_3[], default_soil_vector_4[];
void gortt_gap_probabilities_Q08(gortt_parameters *, gortt_geometry *);

#define REF_NTH 91
#define REF_LUT_LEN (2 * REF_NTH + 2)

/* defaults of gortt.c:67-96 that are not structure inputs */
static void ref_defaults(gortt_parameters *p, gortt_geometry *g)
{
    memset(p, 0, sizeof *p);
    memset(g, 0, sizeof *g);
    p->dz = 0.20;
    p->ds = 0.20;
    p->dth = DTOR(1);
    p->nlayers = 15;
    p->lad = LAD_05;
    p->maxcrowns = 30;
    p->nh_es = 20;
    p->npoints = 32;
    p->use_user_fd = FALSE;
    p->read_prob_file = FALSE;
    p->write_prob_file = FALSE;
    p->use_user_beta = FALSE;
}

static void ref_free_params(gortt_parameters *p)
{
    /* mirrors the deallocation list of gortt.c:355-378 */
    gortt_dealloc_1d(p->height); gortt_dealloc_1d(p->height_p);
    gortt_dealloc_1d(p->theta); gortt_dealloc_1d(p->theta_p);
    gortt_dealloc_1d(p->vb); gortt_dealloc_1d(p->k_open); gortt_dealloc_1d(p->dk_open);
    gortt_dealloc_1d(p->lk_up); gortt_dealloc_1d(p->lk_down); gortt_dealloc_1d(p->k_openep);
    gortt_dealloc_1d(p->es); gortt_dealloc_1d(p->factorial);
    gortt_dealloc_2d(p->fb, p->nlayers); gortt_dealloc_2d(p->t_open, p->nlayers);
    gortt_dealloc_2d(p->dt_open, p->nlayers); gortt_dealloc_2d(p->s_p, p->nlayers);
    gortt_dealloc_2d(p->v_g, p->nlayers); gortt_dealloc_2d(p->p_s0, p->nlayers);
    gortt_dealloc_2d(p->p_n0, p->nlayers); gortt_dealloc_2d(p->epgap, p->nlayers);
    gortt_dealloc_3d(p->pd_s, p->nlayers, p->nth);
}

static void ref_set_structure(gortt_parameters *p, const double *st6)
{
    p->lambda = st6[0]; p->r = st6[1]; p->b = st6[2];
    p->h1 = st6[3]; p->h2 = st6[4]; p->favd = st6[5];
}

/* LUT record: p_n0[0][0..90], epgap[0][0..90], k_open[0], k_openep[0]
 * method 0 = gortt_gap_probabilities, 1 = gortt_gap_probabilities_Q08 (gortt.c:116-120). */
int ref_lut(const double *st6, int method, double *lut)
{
    gortt_parameters p; gortt_geometry g;
    ref_defaults(&p, &g);
    ref_set_structure(&p, st6);
    gortt_init_params(&p, &g);
    if (method == 1) gortt_gap_probabilities_Q08(&p, &g);
    else gortt_gap_probabilities(&p, &g);
    for (int t = 0; t < REF_NTH; t++) {
        lut[t] = p.p_n0[0][t];
        lut[REF_NTH + t] = p.epgap[0][t];
    }
    lut[2 * REF_NTH] = p.k_open[0];
    lut[2 * REF_NTH + 1] = p.k_openep[0];
    ref_free_params(&p);
    return 0;
}

/* Intermediates of the gap-probability path for unit-level checks:
 * v_g[15][91], p_n0[15][91], es-at-h0 is not retained by the reference (overwritten per t),
 * derived scalars (elai, dz, ds, ...) in derived[16], theta_p[91], height_p[15]. */
int ref_lut_intermediates(const double *st6, double *v_g, double *p_n0, double *derived,
                          double *theta_p, double *height_p)
{
    gortt_parameters p; gortt_geometry g;
    ref_defaults(&p, &g);
    ref_set_structure(&p, st6);
    gortt_init_params(&p, &g);
    gortt_gap_probabilities(&p, &g);
    for (int h = 0; h < 15; h++)
        for (int t = 0; t < REF_NTH; t++) {
            v_g[h * REF_NTH + t] = p.v_g[h][t];
            p_n0[h * REF_NTH + t] = p.p_n0[h][t];
        }
    derived[0] = p.ellipticity; derived[1] = p.elai; derived[2] = p.tau; derived[3] = p.tau_p;
    derived[4] = p.lv; derived[5] = p.lv_p; derived[6] = p.z1; derived[7] = p.z2;
    derived[8] = p.dz; derived[9] = p.dz_p; derived[10] = p.h1_p; derived[11] = p.h2_p;
    derived[12] = p.z1_p; derived[13] = p.z2_p; derived[14] = p.h; derived[15] = p.favd_p;
    for (int t = 0; t < REF_NTH; t++) theta_p[t] = p.theta_p[t];
    for (int h = 0; h < 15; h++) height_p[h] = p.height_p[h];
    ref_free_params(&p);
    return 0;
}

/* The intermediates that never reach the BRDF, straight from the reference's struct after its own
 * gortt_gap_probabilities: vb[15], fb[15][91], t_open[15][15], dt_open[15][15], dk_open[15], k_open[15]. */
int ref_lut_dead(const double *st6, double *vb, double *fb, double *t_open, double *dt_open,
                 double *dk_open, double *k_open)
{
    gortt_parameters p; gortt_geometry g;
    ref_defaults(&p, &g);
    ref_set_structure(&p, st6);
    gortt_init_params(&p, &g);
    gortt_gap_probabilities(&p, &g);
    for (int h = 0; h < 15; h++) {
        vb[h] = p.vb[h]; dk_open[h] = p.dk_open[h]; k_open[h] = p.k_open[h];
        for (int t = 0; t < REF_NTH; t++) fb[h * REF_NTH + t] = p.fb[h][t];
        for (int z = 0; z < 15; z++) { t_open[h * 15 + z] = p.t_open[h][z]; dt_open[h * 15 + z] = p.dt_open[h][z]; }
    }
    ref_free_params(&p);
    return 0;
}

/* Fill a reference parameter struct from structure + a LUT record (as "-P" does, gortt.c:131-146). */
static void ref_params_with_lut(gortt_parameters *p, gortt_geometry *g, const double *st6,
                                const double *lut, const double *opt)
{
    ref_defaults(p, g);
    ref_set_structure(p, st6);
    if (opt) {
        if (opt[0] != 0.0) { p->use_user_beta = TRUE; p->beta = opt[1]; }
        if (opt[2] != 0.0) { p->use_user_fd = TRUE; p->fd = opt[3]; }
    }
    gortt_init_params(p, g);
    for (int t = 0; t < REF_NTH; t++) {
        p->p_n0[0][t] = lut[t];
        p->epgap[0][t] = lut[REF_NTH + t];
    }
    p->k_open[0] = lut[2 * REF_NTH];
    p->k_openep[0] = lut[2 * REF_NTH + 1];
}

/* per-line angle preparation, gortt.c:240-291 (degrees in) */
static void ref_prepare_line(gortt_parameters *p, gortt_geometry *g, const double *ang4)
{
    g->vza = DTOR(ang4[0]); g->vaa = DTOR(ang4[1]); g->sza = DTOR(ang4[2]); g->saa = DTOR(ang4[3]);
    if (g->sza < 0.0) { g->saa += M_PI; g->sza *= -1.0; }
    if (g->vza < 0.0) { g->vaa += M_PI; g->vza *= -1.0; }
    while (g->saa > 2 * M_PI) g->saa -= 2 * M_PI;
    while (g->vaa > 2 * M_PI) g->vaa -= 2 * M_PI;
    while (g->saa < 0) g->saa += 2 * M_PI;
    while (g->vaa < 0) g->vaa += 2 * M_PI;
    g->raa = g->saa - g->vaa;
    g->raa = fabs((g->raa - 2 * M_PI * (int) (0.5 + g->raa * M_1_PI * 0.5)));
    g->vza_prime = gortt_prime_theta(p, g->vza);
    g->sza_prime = gortt_prime_theta(p, g->sza);
    p->k_vza = gortt_leaf_angle_distribution(p, g->vza);
    if (!p->use_user_fd) p->fd = cos(g->sza) / (cos(g->sza) + 0.09);
}

/* BRDF for ngeom lines (vza vaa sza saa in degrees) x nw wavelengths.
 * opt = {use_beta, beta, use_fd, fd} or NULL.
 * rsurf[ngeom][nw]; scomp[ngeom][nw][4] (C,G,T,Z) or NULL; kprop[ngeom][4] (Kc,Kg,Kt,Kz) or NULL */
int ref_brdf(const double *st6, const double *lut, const double *opt,
             int ngeom, const double *ang, int nw,
             const double *rleaf, const double *tleaf, const double *rsoil,
             double *rsurf, double *scomp, double *kprop)
{
    gortt_parameters p; gortt_geometry g; gortt_spectra s;
    ref_params_with_lut(&p, &g, st6, lut, opt);
    memset(&s, 0, sizeof s);
    s.nw = nw;
    s.rleaf = (double *) rleaf; s.tleaf = (double *) tleaf; s.rsoil = (double *) rsoil;
    s.rsurf = (double *) malloc(sizeof(double) * nw);
    s.scomp = (double *) malloc(sizeof(double) * nw * 4);
    for (int i = 0; i < ngeom; i++) {
        ref_prepare_line(&p, &g, ang + 4 * i);
        gortt_set_zenith_dependant_probabilities(&p, &g);
        gortt_rsurf(&p, &g, &s);
        memcpy(rsurf + (size_t) i * nw, s.rsurf, sizeof(double) * nw);
        if (scomp) memcpy(scomp + (size_t) i * nw * 4, s.scomp, sizeof(double) * nw * 4);
        if (kprop) { kprop[4 * i] = g.Kc; kprop[4 * i + 1] = g.Kg; kprop[4 * i + 2] = g.Kt; kprop[4 * i + 3] = g.Kz; }
    }
    free(s.rsurf); free(s.scomp);
    ref_free_params(&p);
    return 0;
}

/* Albedo / fAPAR (gortt_albedo.c:7-138) for ngeom sun geometries.  The reference sizes its
 * scratch by npoints=32 but indexes it by wavelength (gortt_albedo.c:79-88), so wavelengths
 * are driven in chunks of <= 32. Outputs [ngeom][nw]. */
int ref_energy(const double *st6, const double *lut, const double *opt,
               int ngeom, const double *ang, int nw,
               const double *rleaf, const double *tleaf, const double *rsoil,
               double *albedo, double *favegt, double *fasoil)
{
    gortt_parameters p; gortt_geometry g; gortt_spectra s;
    ref_params_with_lut(&p, &g, st6, lut, opt);
    p.abscissa = (double *) malloc(sizeof(double) * p.npoints);
    p.weights = (double *) malloc(sizeof(double) * p.npoints);
    gauleg(-1., 1., p.abscissa, p.weights, p.npoints);
    memset(&s, 0, sizeof s);
    s.rsurf = (double *) malloc(sizeof(double) * 32);
    s.scomp = (double *) malloc(sizeof(double) * 32 * 4);
    for (int i = 0; i < ngeom; i++) {
        for (int w0 = 0; w0 < nw; w0 += 32) {
            int cw = nw - w0 < 32 ? nw - w0 : 32;
            s.nw = cw;
            s.rleaf = (double *) rleaf + w0; s.tleaf = (double *) tleaf + w0; s.rsoil = (double *) rsoil + w0;
            s.albedo = albedo + (size_t) i * nw + w0;
            s.favegt = favegt + (size_t) i * nw + w0;
            s.fasoil = fasoil + (size_t) i * nw + w0;
            ref_prepare_line(&p, &g, ang + 4 * i);
            gortt_set_zenith_dependant_probabilities(&p, &g);
            gortt_rsurf(&p, &g, &s);       /* as main does before gortt_energy, gortt.c:294-322 */
            gortt_energy(&p, &g, &s);
        }
    }
    free(s.rsurf); free(s.scomp);
    free(p.abscissa); free(p.weights);
    ref_free_params(&p);
    return 0;
}

/* Leaf and soil spectra through the reference's own interface functions
 * (gortt.c:1286-1374).  leaf7 = N,Cab,Car,Anth,Cbrown,Cw,Cm; soil4 = rsl1..4;
 * user_leaf/user_soil < 0 means "not set". */
int ref_spectra(const double *leaf7, const double *soil4, double user_leaf, double user_soil,
                int nw, const double *wl, double *rleaf, double *tleaf, double *rsoil)
{
    gortt_spectra s;
    memset(&s, 0, sizeof s);
    s.nw = nw; s.wavelength = (double *) wl;
    s.rleaf = rleaf; s.tleaf = tleaf; s.rsoil = rsoil;
    s.p_N = leaf7[0]; s.p_Cab = leaf7[1]; s.p_Car = leaf7[2]; s.p_Anth = leaf7[3];
    s.p_Cbrown = leaf7[4]; s.p_Cw = leaf7[5]; s.p_Cm = leaf7[6];
    s.rsl1 = soil4[0]; s.rsl2 = soil4[1]; s.rsl3 = soil4[2]; s.rsl4 = soil4[3];
    s.is_user_leaf = user_leaf >= 0.0; s.user_r_leaf = user_leaf;
    s.is_user_soil = user_soil >= 0.0; s.user_r_soil = user_soil;
    gortt_price_soil(&s, default_soil_vector_1, default_soil_vector_2, default_soil_vector_3, default_soil_vector_4);
    gortt_prospect_interface(&s);
    return 0;
}

void ref_gauleg(int n, double *x, double *w) { gauleg(-1., 1., x, w, n); }

/* Timing leg for bench.py (cpu_baseline / --impl reference): evaluates the BRDF path
 * `reps` times over the given lines and returns the number of (geometry, wavelength)
 * evaluations performed.  Nothing but the reference's own functions is timed. */
long ref_brdf_repeat(const double *st6, const double *lut, int ngeom, const double *ang, int nw,
                     const double *rleaf, const double *tleaf, const double *rsoil,
                     int reps, double *rsurf_last)
{
    long n = 0;
    double *buf = (double *) malloc(sizeof(double) * (size_t) ngeom * nw);
    for (int r = 0; r < reps; r++) {
        ref_brdf(st6, lut, NULL, ngeom, ang, nw, rleaf, tleaf, rsoil, buf, NULL, NULL);
        n += (long) ngeom * nw;
    }
    if (rsurf_last) memcpy(rsurf_last, buf, sizeof(double) * (size_t) ngeom * nw);
    free(buf);
    return n;
}
