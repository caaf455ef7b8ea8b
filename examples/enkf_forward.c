/* enkf_forward.c -- an ensemble forward operator in plain C over the C ABI (include/gort_b200.h): the batched use
 * INTEGRATION.md section 3 describes.  M ensemble members (canopy structure, leaf biochemistry, soil weights),
 * G view/illumination geometries per member, MODIS-like bands; gap-probability LUTs, PROSPECT-D / Price spectra and
 * BRDF all on the GPU, one call each.
 *
 *   enkf_forward <members.bin> <out.bin>
 *
 * members.bin: int32 M, G, W; then doubles structure[6][M], leaf[7][M], soil[4][M], wavelength[W], angles[4][M][G].
 * out.bin:     doubles rsurf[M][G][W].
 * Built by gort_b200/csrc/Makefile into gort_b200/bin/enkf_forward; tests/test_examples_gpu.py compares its output
 * with the same calls made through the Python binding.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "gort_b200.h"

static void die(const char *what, gort_ctx *ctx)
{
    fprintf(stderr, "enkf_forward: %s: %s\n", what, gort_last_error(ctx));
    exit(EXIT_FAILURE);
}

static double *read_doubles(FILE *f, size_t n)
{
    double *p = (double *) malloc(sizeof(double) * (n ? n : 1));
    if (!p || fread(p, sizeof(double), n, f) != n) { fprintf(stderr, "enkf_forward: short read\n"); exit(EXIT_FAILURE); }
    return p;
}

int main(int argc, char **argv)
{
    if (argc != 3) { fprintf(stderr, "usage: %s members.bin out.bin\n", argv[0]); return EXIT_FAILURE; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return EXIT_FAILURE; }
    int32_t hdr[3];
    if (fread(hdr, sizeof(int32_t), 3, f) != 3) { fprintf(stderr, "enkf_forward: bad header\n"); return EXIT_FAILURE; }
    const size_t M = (size_t) hdr[0], G = (size_t) hdr[1], W = (size_t) hdr[2];
    double *structure = read_doubles(f, 6 * M), *leaf = read_doubles(f, 7 * M), *soil = read_doubles(f, 4 * M);
    double *wl = read_doubles(f, W), *angles = read_doubles(f, 4 * M * G);
    fclose(f);

    gort_ctx *gx = NULL;
    if (gort_create(0, &gx) != GORT_OK) die("gort_create", NULL);

    double *lut = (double *) malloc(sizeof(double) * M * GORT_LUT_STRIDE);
    double *rleaf = (double *) malloc(sizeof(double) * M * W), *tleaf = (double *) malloc(sizeof(double) * M * W);
    double *rsoil = (double *) malloc(sizeof(double) * M * W), *rsurf = (double *) malloc(sizeof(double) * M * G * W);
    if (!lut || !rleaf || !tleaf || !rsoil || !rsurf) { fprintf(stderr, "enkf_forward: out of memory\n"); return EXIT_FAILURE; }

    if (gort_lut_batch(gx, (int) M, structure, GORT_LUT_FULL, lut) != GORT_OK) die("gort_lut_batch", gx);
    if (gort_spectra_batch(gx, (int) M, leaf, soil, -1.0, -1.0, (int) W, wl, rleaf, tleaf, rsoil) != GORT_OK)
        die("gort_spectra_batch", gx);
    gort_shape sh = { 0 };
    sh.n_sets = (int) M; sh.n_geom = (int) G; sh.n_wl = (int) W;
    sh.geom_per_set = 1; sh.spectra_per_set = 1;
    if (gort_brdf_batch(gx, &sh, structure, lut, angles, rleaf, tleaf, rsoil, rsurf, NULL, NULL) != GORT_OK)
        die("gort_brdf_batch", gx);

    f = fopen(argv[2], "wb");
    if (!f || fwrite(rsurf, sizeof(double), M * G * W, f) != M * G * W) { perror(argv[2]); return EXIT_FAILURE; }
    fclose(f);
    printf("enkf_forward: %zu members x %zu geometries x %zu bands, %ld kernel launches\n", M, G, W, gort_launch_count(gx));
    gort_destroy(gx);
    free(structure); free(leaf); free(soil); free(wl); free(angles); free(lut); free(rleaf); free(tleaf); free(rsoil); free(rsurf);
    return EXIT_SUCCESS;
}
