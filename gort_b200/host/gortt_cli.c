/* gortt_cli.c -- the `gortt` command line, host side in C over the C ABI (include/gort_b200.h).
 *
 * Drop-in for the reference program's process interface (tquaife/gort):
 *     gortt [options] < angles.dat > output.dat            README.md:30-36
 * Same option set, matched in the same order with the same prefix lengths (gortt.c:1022-1115), same
 * float-typed -HB/-BR/-PCC/-LAI (gortt.c:1014), same stdin format (gortt.c:153-184, :232-237), same
 * stdout bytes (gortt.c:230, :310-327), same "-W" / "-P" LUT text layout (gortt.c:123-146), same error
 * messages and exit codes.  All model arithmetic runs on the GPU through libgort_b200; this file only
 * parses, batches the input lines into one call per kernel, and prints.
 *
 * Known, deliberate differences (DESIGN.md "CLI"):
 *   - input lines are read with getline(): the reference's 1000-byte line buffer (include/gortt.h:28)
 *     that caps a run at ~247 wavelengths is lifted; shorter inputs behave identically;
 *   - the -u usage text documents the same options but is not byte-identical;
 *   - all geometry lines are evaluated in one batch before printing, so on a malformed line k the
 *     k-1 good lines are still printed first, exactly as the reference would have printed them;
 *   - the reference's out-of-bounds reads (nw > 32 with -energy, wavelength 2500) are not inherited.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include "gort_b200.h"
#include "fastfmt.h"

typedef struct {
    /* structure, gortt.c:67-72 */
    double lambda, r, b, h1, h2, favd;
    /* spectra, gortt.c:38-59 */
    double leaf[7], soil[4];
    int is_user_leaf, is_user_soil, is_file_soil, soil_dump;
    double user_r_leaf, user_r_soil;
    char *soil_file;
    /* control, gortt.c:32-36 */
    int prnspec, prnprop, energy, q08, lidar;
    int read_prob, write_prob;
    char *prob_fn;
    gort_options opt;
} cli_t;

static void usage(const char *bin)
{
    fprintf(stderr,
        "usage: %s [options] < angles.dat\n\n"
        "stdin, line 1:   N M W_1 ... W_M   (N geometries, M wavelengths in nm, 400-2500)\n"
        "stdin, N lines:  view_zenith view_azimuth solar_zenith solar_azimuth   (degrees)\n\n"
        "crown geometry:\n"
        "  -beta x      force the mutual-shadowing proportion to x (default: Li & Strahler 1992 model)\n"
        "  old style:   -h1 x  -h2 x  lower/upper bound of crown centres (m);  -b x  -r x  vertical/\n"
        "               horizontal crown radius (m);  -lambda x  stem density (1/m2)\n"
        "  new style:   -HB x  centroid height range / vertical radius;  -BR x  vertical / horizontal\n"
        "               radius;  -PCC x  projected crown cover at nadir  (any of these overrides old style)\n"
        "leaf material: -favd x  foliage volume area density,  or  -LAI x  scene leaf area index\n"
        "PROSPECT-D:    -N x  -Cab x  -Car x  -Canth x  -Cbrown x  -Cw x  -Cm x\n"
        "Price soil:    -rsl1 x  -rsl2 x  -rsl3 x  -rsl4 x\n"
        "overrides:     -alb_leaf x (PROSPECT off)  -alb_soil x (Price off)  -soil_spectra file (Price off)\n"
        "               -soil_dump file  print the 1-nm soil table of file and exit (the reference's unfinished -soil_spectra)\n"
        "diffuse light: -diffuse x  (diffuse fraction; default cos(sza)/(cos(sza)+0.09) direct)\n"
        "gap LUT:       -W  write gap probabilities to stdout and exit;  -P file  read them back\n"
        "               -q08_pn_kopen  closed-form gap probabilities of Quaife et al. (2008)\n"
        "output:        -prnspec  component spectra in {}   -prnprop  viewed proportions in []\n"
        "               -energy   albedo, fAPAR(veg), fA(soil) per wavelength   -u  this message\n\n",
        bin);
}

/* gortt.c:1003-1136 */
static void parse(int argc, char **argv, cli_t *c)
{
    int use_true_p = 0, use_lai = 0;
    float hb = 2.0f, br = 1.0f, pcc = 0.5f, lai = 2.0f;
#define ARG() (i + 1 < argc ? atof(argv[++i]) : (fprintf(stderr, "%s: option %s needs a value\n", argv[0], argv[i]), exit(EXIT_FAILURE), 0.0))
#define SARG() (i + 1 < argc ? argv[++i] : (fprintf(stderr, "%s: option %s needs a value\n", argv[0], argv[i]), exit(EXIT_FAILURE), (char *) 0))
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (*a != '-') {
            fprintf(stderr, "%s: unknown argument on command line: %s\n", argv[0], argv[1]);   /* sic: argv[1] */
            fprintf(stderr, "(use the option -u to see brief usage instructions)\n");
            exit(EXIT_FAILURE);
        }
        /**/ if (!strncasecmp(a, "-favd", 5)) c->favd = ARG();
        else if (!strncasecmp(a, "-h1", 3)) c->h1 = ARG();
        else if (!strncasecmp(a, "-h2", 3)) c->h2 = ARG();
        else if (!strncasecmp(a, "-lambda", 7)) c->lambda = ARG();
        else if (!strncmp(a, "-HB", 3)) { use_true_p = 1; hb = (float) ARG(); }
        else if (!strncmp(a, "-BR", 3)) { use_true_p = 1; br = (float) ARG(); }
        else if (!strncmp(a, "-PCC", 7)) { use_true_p = 1; pcc = (float) ARG(); }
        else if (!strncmp(a, "-LAI", 7)) { use_lai = 1; lai = (float) ARG(); }
        else if (!strncasecmp(a, "-beta", 5)) { c->opt.use_beta = 1; c->opt.beta = ARG(); }
        else if (!strncasecmp(a, "-diffuse", 5)) { c->opt.use_fd = 1; c->opt.fd = 1.0 - ARG(); }
        else if (!strncmp(a, "-alb_leaf", 9)) { c->is_user_leaf = 1; c->user_r_leaf = ARG(); }
        else if (!strncmp(a, "-alb_soil", 9)) { c->is_user_soil = 1; c->is_file_soil = 0; c->user_r_soil = ARG(); }
        else if (!strncmp(a, "-soil_spectra", 10)) { c->is_user_soil = 0; c->is_file_soil = 1; c->soil_file = SARG(); }
        else if (!strncmp(a, "-soil_dump", 10)) { c->is_user_soil = 0; c->is_file_soil = 1; c->soil_dump = 1; c->soil_file = SARG(); }
        else if (!strncmp(a, "-prnspec", 7)) c->prnspec = 1;
        else if (!strncmp(a, "-prnprop", 7)) c->prnprop = 1;
        else if (!strncmp(a, "-energy", 7)) c->energy = 1;
        else if (!strncmp(a, "-q08_pn_kopen", 7)) c->q08 = 1;
        else if (!strncmp(a, "-lidar", 6)) c->lidar = 1;
        else if (!strncmp(a, "-P", 2)) { c->read_prob = 1; c->prob_fn = SARG(); }
        else if (!strncmp(a, "-W", 2)) c->write_prob = 1;
        else if (!strncasecmp(a, "-N", 2)) c->leaf[0] = ARG();
        else if (!strncasecmp(a, "-cab", 4)) c->leaf[1] = ARG();
        else if (!strncasecmp(a, "-car", 4)) c->leaf[2] = ARG();
        else if (!strncasecmp(a, "-canth", 3)) c->leaf[3] = ARG();
        else if (!strncasecmp(a, "-cbrown", 3)) c->leaf[4] = ARG();
        else if (!strncasecmp(a, "-cw", 3)) c->leaf[5] = ARG();
        else if (!strncasecmp(a, "-cm", 3)) c->leaf[6] = ARG();
        else if (!strncasecmp(a, "-rsl1", 5)) c->soil[0] = ARG();
        else if (!strncasecmp(a, "-rsl2", 5)) c->soil[1] = ARG();
        else if (!strncasecmp(a, "-rsl3", 5)) c->soil[2] = ARG();
        else if (!strncasecmp(a, "-rsl4", 5)) c->soil[3] = ARG();
        else if (!strncasecmp(a, "-b", 2)) c->b = ARG();
        else if (!strncasecmp(a, "-r", 2)) c->r = ARG();
        else if (!strncasecmp(a, "-u", 2)) { usage(argv[0]); exit(EXIT_SUCCESS); }
        else {
            fprintf(stderr, "%s: unknown option on command line: %s\n", argv[0], a);
            fprintf(stderr, "(use the option -u to see brief usage instructions)\n");
            exit(EXIT_FAILURE);
        }
    }
    if (use_true_p) {                                   /* gortt.c:1117-1125 */
        c->r = 10.;
        c->b = br * c->r;
        c->h1 = c->b * 2.;
        c->h2 = hb * c->b + c->h1;
        c->lambda = pcc / (c->r * c->r * M_PI);
    }
    if (use_lai)                                        /* gortt.c:1127-1131 */
        c->favd = lai * 3. / (c->lambda * c->r * c->r * M_PI * c->b * 4.0);
}

/* "-soil_spectra file" (gortt.c:1056-1060, :103).  The reference's reader (gortt.c:1388-1451) is a development stub:
 * it builds the 1-nm table, prints it and exits with failure.  Here the option is FINISHED: the table replaces the
 * Price soil spectrum of the run (gort_soil_table_read + gort_soil_from_table, include/gort_b200.h).  "-soil_dump
 * file" keeps what the stub does -- the table as "%d %lf" lines, then exit(EXIT_FAILURE) -- so that the table can
 * still be compared byte for byte with the reference binary's.  Malformed files give the reference's messages. */
static void soil_file_read(const cli_t *c, double *table)
{
    char err[700];
    if (gort_soil_table_read(c->soil_file, table, err, sizeof err) != GORT_OK) {
        fprintf(stderr, "gortt: %s\n", err);
        exit(EXIT_FAILURE);
    }
}

static void soil_file_dump(const cli_t *c)
{
    static double table[GORT_SOIL_TABLE_NW];
    soil_file_read(c, table);
    for (int i = 0; i <= 2100; i++) printf("%d %lf\n", i + 400, table[i]);      /* gortt.c:1441-1442 */
    exit(EXIT_FAILURE);
}

static void die_gort(const char *bin, gort_ctx *ctx, const char *what)
{
    fprintf(stderr, "%s: %s: %s\n", bin, what, gort_last_error(ctx));
    exit(EXIT_FAILURE);
}

/* next whitespace-delimited token of *p (gortt.c:1237-1283), NULL at end of line */
static char *next_token(char **p)
{
    char *s = *p;
    while (*s && isspace((unsigned char) *s)) s++;
    if (!*s) return NULL;
    char *t = s;
    while (*s && !isspace((unsigned char) *s)) s++;
    if (*s) *s++ = '\0';
    *p = s;
    return t;
}

int main(int argc, char **argv)
{
    cli_t c;
    memset(&c, 0, sizeof c);
    c.soil[0] = 0.2; c.soil[1] = 0.1; c.soil[2] = 0.03726; c.soil[3] = -0.002426;        /* gortt.c:38-41 */
    c.leaf[0] = 1.2; c.leaf[1] = 30.; c.leaf[2] = 10.; c.leaf[3] = 1.0; c.leaf[4] = 0.0;  /* gortt.c:53-59 */
    c.leaf[5] = 0.015; c.leaf[6] = 0.009;
    c.lambda = 0.405; c.r = 0.76; c.b = 3.55263 * c.r; c.h1 = 3.0; c.h2 = 8.5; c.favd = 0.858;   /* gortt.c:67-72 */
    parse(argc, argv, &c);

    static double soil_table[GORT_SOIL_TABLE_NW];
    if (c.is_file_soil) {                               /* gortt.c:103 */
        if (c.soil_dump) soil_file_dump(&c);
        soil_file_read(&c, soil_table);
    }

    gort_ctx *ctx = NULL;
    if (gort_create(0, &ctx) != GORT_OK) die_gort(argv[0], NULL, "cannot initialise the GPU");

    /* ---- gap probabilities, gortt.c:116-146 ---- */
    double st[6] = { c.lambda, c.r, c.b, c.h1, c.h2, c.favd };
    double lut[GORT_LUT_STRIDE];
    memset(lut, 0, sizeof lut);
    if (!c.read_prob)
        if (gort_lut_batch(ctx, 1, st, c.q08 ? GORT_LUT_Q08 : GORT_LUT_FULL, lut) != GORT_OK)
            die_gort(argv[0], ctx, "gap probabilities");
    if (c.write_prob) {
        gort_lut_write_text(lut, stdout);
        gort_destroy(ctx);
        return EXIT_SUCCESS;
    }
    if (c.read_prob) {
        if (gort_lut_read_text(c.prob_fn, lut) != GORT_OK) {
            fprintf(stderr, "%s: error opening probability file: %s\n", argv[0], c.prob_fn);
            exit(EXIT_FAILURE);
        }
    }

    /* ---- header line, gortt.c:153-184 ---- */
    char *line = NULL; size_t cap = 0;
    if (getline(&line, &cap, stdin) < 0) {
        fprintf(stderr, "%s: error reading data on stdin\n", argv[0]);
        exit(EXIT_FAILURE);
    }
    char *out_head = strdup(line);
    char *p = line, *tok;
    if (!(tok = next_token(&p))) {
        fprintf(stderr, "%s: error reading number of angles from line 1\n", argv[0]);
        exit(EXIT_FAILURE);
    }
    int na_check = atoi(tok);
    if (!(tok = next_token(&p))) {
        fprintf(stderr, "%s: error reading number of wavebands from line 1\n", argv[0]);
        exit(EXIT_FAILURE);
    }
    int nw_check = atoi(tok);
    int nw = 0, wcap = nw_check > 0 ? nw_check : 1;
    double *wl = (double *) malloc(sizeof(double) * wcap);
    while ((tok = next_token(&p))) {
        if (nw == wcap) { wcap *= 2; wl = (double *) realloc(wl, sizeof(double) * wcap); }
        wl[nw++] = atof(tok);
    }
    if (nw_check != nw) {
        fprintf(stderr, "%s: expected number of wavelengths (%d) does not match with number found (%d)\n", argv[0], nw_check, nw);
        exit(EXIT_FAILURE);
    }

    /* ---- spectra, gortt.c:224-227 ---- */
    double *rleaf = NULL, *tleaf = NULL, *rsoil = NULL;
    if (nw > 0) {
        rleaf = (double *) malloc(sizeof(double) * nw);
        tleaf = (double *) malloc(sizeof(double) * nw);
        rsoil = (double *) malloc(sizeof(double) * nw);
        int rc = gort_spectra_batch(ctx, 1, c.leaf, c.soil, c.is_user_leaf ? c.user_r_leaf : -1.0,
                                    c.is_user_soil ? c.user_r_soil : -1.0, nw, wl, rleaf, tleaf, rsoil);
        if (rc == GORT_ERR_RANGE) {
            fprintf(stderr, "gortt_price_soil: wavlength out of range (400-2500)\n");      /* gortt.c:1300 */
            exit(EXIT_FAILURE);
        }
        if (rc != GORT_OK) die_gort(argv[0], ctx, "spectra");
        if (c.is_file_soil && gort_soil_from_table(ctx, soil_table, 1, nw, wl, rsoil) != GORT_OK) die_gort(argv[0], ctx, "soil spectrum");
        /* -alb_leaf / -alb_soil may be negative in the reference; the ABI uses "< 0" for "not set" */
        if (c.is_user_leaf && c.user_r_leaf < 0.0) for (int i = 0; i < nw; i++) rleaf[i] = tleaf[i] = c.user_r_leaf / 2.0;
        if (c.is_user_soil && c.user_r_soil < 0.0) for (int i = 0; i < nw; i++) rsoil[i] = c.user_r_soil;
    }

    printf("%s", out_head);                             /* gortt.c:230 */

    /* ---- geometry lines, gortt.c:232-238: gather, then one batched call per kernel ---- */
    int na = 0, acap = na_check > 0 ? na_check : 16, bad_line = 0;
    double *ang = (double *) malloc(sizeof(double) * 4 * acap);     /* line-major while reading */
    while (getline(&line, &cap, stdin) >= 0) {
        double v[4];
        if (sscanf(line, "%lf %lf %lf %lf", &v[0], &v[1], &v[2], &v[3]) != 4) { bad_line = 1; break; }
        if (na == acap) { acap *= 2; ang = (double *) realloc(ang, sizeof(double) * 4 * acap); }
        memcpy(ang + 4 * (size_t) na, v, sizeof v);
        na++;
    }

    if (na > 0 && nw > 0) {
        double *soa = (double *) malloc(sizeof(double) * 4 * na);   /* [4][na] for the ABI */
        for (int i = 0; i < na; i++) for (int k = 0; k < 4; k++) soa[(size_t) k * na + i] = ang[4 * (size_t) i + k];
        gort_shape sh;
        memset(&sh, 0, sizeof sh);
        sh.n_sets = 1; sh.n_geom = na; sh.n_wl = nw; sh.opt = c.opt;
        size_t n = (size_t) na * nw;
        double *rsurf = (double *) malloc(sizeof(double) * n);
        double *scomp = c.prnspec ? (double *) malloc(sizeof(double) * 4 * n) : NULL;
        double *kprop = c.prnprop ? (double *) malloc(sizeof(double) * 4 * na) : NULL;
        if (gort_brdf_batch(ctx, &sh, st, lut, soa, rleaf, tleaf, rsoil, rsurf, scomp, kprop) != GORT_OK)
            die_gort(argv[0], ctx, "BRDF");
        double *alb = NULL, *fv = NULL, *fs = NULL;
        if (c.energy) {                                 /* gortt.c:321-325 */
            alb = (double *) malloc(sizeof(double) * n);
            fv = (double *) malloc(sizeof(double) * n);
            fs = (double *) malloc(sizeof(double) * n);
            if (gort_energy_batch(ctx, &sh, st, lut, soa, rleaf, tleaf, rsoil, alb, fv, fs) != GORT_OK)
                die_gort(argv[0], ctx, "energy balance");
        }
        /* gortt.c:310-327; same bytes as the reference's printf("%f ") calls, through a fast formatter */
        gort_out *out = (gort_out *) malloc(sizeof(gort_out));
        if (!out) { fprintf(stderr, "%s: out of memory\n", argv[0]); exit(EXIT_FAILURE); }
        fflush(stdout);
        out->fp = stdout; out->n = 0;
        for (int i = 0; i < na; i++) {
            const double *a4 = ang + 4 * (size_t) i;
            for (int q = 0; q < 4; q++) gort_out_f(out, a4[q]);
            for (int k = 0; k < nw; k++) {
                size_t o = (size_t) i * nw + k;
                gort_out_f(out, rsurf[o]);
                if (c.prnspec) {
                    gort_out_str(out, "{ ");
                    for (int q = 0; q < 4; q++) gort_out_f(out, scomp[4 * o + q]);
                    gort_out_str(out, "} ");
                }
            }
            if (c.prnprop) {
                gort_out_str(out, "[ ");
                for (int q = 0; q < 4; q++) gort_out_f(out, kprop[4 * (size_t) i + q]);
                gort_out_str(out, "] ");
            }
            if (c.energy)
                for (int k = 0; k < nw; k++) {
                    size_t o = (size_t) i * nw + k;
                    gort_out_f(out, alb[o]); gort_out_f(out, fv[o]); gort_out_f(out, fs[o]);
                }
            gort_out_str(out, "\n");
        }
        gort_out_flush(out);
        free(out);
        free(soa); free(rsurf); free(scomp); free(kprop); free(alb); free(fv); free(fs);
    } else if (na > 0) {
        for (int i = 0; i < na; i++) {                  /* no wavelengths: angles only */
            const double *a4 = ang + 4 * (size_t) i;
            printf("%f %f %f %f ", a4[0], a4[1], a4[2], a4[3]);
            if (c.prnprop) {                            /* proportions do not depend on wavelength */
                gort_shape sh; memset(&sh, 0, sizeof sh);
                sh.n_sets = 1; sh.n_geom = 1; sh.n_wl = 1; sh.opt = c.opt;
                double a1[4] = { a4[0], a4[1], a4[2], a4[3] }, one = 0.25, rs1 = 0.1, r1, k4[4];
                if (gort_brdf_batch(ctx, &sh, st, lut, a1, &one, &one, &rs1, &r1, NULL, k4) != GORT_OK)
                    die_gort(argv[0], ctx, "BRDF");
                printf("[ %f %f %f %f ] ", k4[0], k4[1], k4[2], k4[3]);
            }
            printf("\n");
        }
    }
    fflush(stdout);

    if (bad_line) {                                     /* gortt.c:234-237 */
        fprintf(stderr, "%s: error on input, line %d\n", argv[0], na + 1);
        exit(EXIT_FAILURE);
    }
    if (na_check != na) {                               /* gortt.c:331-334 */
        fprintf(stderr, "%s: expected number of angles (%d) does not match with number found (%d)\n", argv[0], na_check, na);
        exit(EXIT_FAILURE);
    }
    free(ang); free(wl); free(rleaf); free(tleaf); free(rsoil); free(out_head); free(line);
    gort_destroy(ctx);
    return EXIT_SUCCESS;
}
