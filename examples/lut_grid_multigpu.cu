/* lut_grid_multigpu.cu -- gap-probability LUTs of a parameter grid (BASELINE.json config 5) assembled on EVERY GPU of one
 * box by a C/C++ host, one process, no torch, no NCCL: the use INTEGRATION.md section 3 describes.
 *
 * GPU g owns a contiguous block of the M parameter sets and calls gort_lut_batch_scatter_dev: the kernels that produce a
 * record store it into GPU g's table and, over NVLink peer mappings, into the same rows of every other GPU's table.
 * The reference computes one record per process (gortt.c:108-120) and writes it with -W (gortt.c:123-128).
 *
 *   lut_grid_multigpu <structure.bin> <out.bin> [n_gpus]
 *
 * structure.bin: int32 M; then doubles structure[6][M] (lambda, r, b, h1, h2, favd).
 * out.bin:       doubles lut[M][184], the table of the LAST GPU used (every table is checked to hold the same bytes).
 * Built by gort_b200/csrc/Makefile into gort_b200/bin/lut_grid_multigpu; tests/test_examples_gpu.py compares out.bin with
 * the records one GPU computes through the Python binding.
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "gort_b200.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "lut_grid_multigpu: %s: %s\n", #call, cudaGetErrorString(e_)); return EXIT_FAILURE; } } while (0)

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s structure.bin out.bin [n_gpus]\n", argv[0]); return EXIT_FAILURE; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return EXIT_FAILURE; }
    int32_t M32 = 0;
    if (fread(&M32, sizeof M32, 1, f) != 1 || M32 <= 0) { fprintf(stderr, "lut_grid_multigpu: bad header\n"); return EXIT_FAILURE; }
    const size_t M = (size_t) M32;
    std::vector<double> structure(6 * M);
    if (fread(structure.data(), sizeof(double), 6 * M, f) != 6 * M) { fprintf(stderr, "lut_grid_multigpu: short read\n"); return EXIT_FAILURE; }
    fclose(f);

    int n_dev = 0;
    CK(cudaGetDeviceCount(&n_dev));
    int n = argc > 3 ? atoi(argv[3]) : n_dev;
    if (n > n_dev) n = n_dev;
    if (n > GORT_LUT_MAX_DST + 1) n = GORT_LUT_MAX_DST + 1;
    if ((size_t) n > M) n = (int) M;
    if (n < 1) { fprintf(stderr, "lut_grid_multigpu: no CUDA device\n"); return EXIT_FAILURE; }

    // contiguous blocks, sizes differing by at most one (larger blocks first)
    std::vector<size_t> lo(n + 1, 0);
    for (int g = 0; g < n; g++) lo[g + 1] = lo[g] + M / n + ((size_t) g < M % n ? 1 : 0);

    std::vector<gort_ctx *> gx(n, nullptr);
    std::vector<double *> table(n, nullptr), d_st(n, nullptr);
    const size_t table_bytes = M * GORT_LUT_STRIDE * sizeof(double);
    for (int g = 0; g < n; g++) {
        CK(cudaSetDevice(g));
        for (int p = 0; p < n; p++) {
            if (p == g) continue;
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, g, p));
            if (!can) { fprintf(stderr, "lut_grid_multigpu: GPU %d cannot map GPU %d's memory\n", g, p); return EXIT_FAILURE; }
            cudaError_t e = cudaDeviceEnablePeerAccess(p, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
            (void) cudaGetLastError();
        }
        CK(cudaMalloc(&table[g], table_bytes));
        CK(cudaMemset(table[g], 0xff, table_bytes));                       // NaN pattern: a row nobody wrote shows
        // this GPU's block of the structure array, [6][m]
        const size_t m = lo[g + 1] - lo[g];
        std::vector<double> blk(6 * m);
        for (int k = 0; k < 6; k++) memcpy(&blk[k * m], &structure[k * M + lo[g]], m * sizeof(double));
        CK(cudaMalloc(&d_st[g], 6 * m * sizeof(double)));
        CK(cudaMemcpy(d_st[g], blk.data(), 6 * m * sizeof(double), cudaMemcpyHostToDevice));
        if (gort_create(g, &gx[g]) != GORT_OK) { fprintf(stderr, "lut_grid_multigpu: gort_create(%d): %s\n", g, gort_last_error(NULL)); return EXIT_FAILURE; }
    }
    for (int g = 0; g < n; g++) CK((cudaSetDevice(g), cudaDeviceSynchronize()));      // every table is cleared before anyone stores

    // enqueue only: the n calls run concurrently, each GPU storing its rows into all n tables
    for (int g = 0; g < n; g++) {
        double *dst[GORT_LUT_MAX_DST];
        int nd = 0;
        for (int p = 0; p < n; p++) if (p != g) dst[nd++] = table[p] + lo[g] * GORT_LUT_STRIDE;
        if (gort_lut_batch_scatter_dev(gx[g], NULL, (int) (lo[g + 1] - lo[g]), d_st[g], GORT_LUT_FULL,
                                       table[g] + lo[g] * GORT_LUT_STRIDE, nd, dst, 0) != GORT_OK) {
            fprintf(stderr, "lut_grid_multigpu: gort_lut_batch_scatter_dev on GPU %d: %s\n", g, gort_last_error(gx[g]));
            return EXIT_FAILURE;
        }
    }
    for (int g = 0; g < n; g++)
        if (gort_synchronize(gx[g]) != GORT_OK) { fprintf(stderr, "lut_grid_multigpu: GPU %d: %s\n", g, gort_last_error(gx[g])); return EXIT_FAILURE; }

    // every table holds the same bytes
    std::vector<double> first(M * GORT_LUT_STRIDE), other(M * GORT_LUT_STRIDE);
    CK(cudaSetDevice(n - 1));
    CK(cudaMemcpy(first.data(), table[n - 1], table_bytes, cudaMemcpyDeviceToHost));
    int same = 1;
    for (int g = 0; g + 1 < n; g++) {
        CK(cudaSetDevice(g));
        CK(cudaMemcpy(other.data(), table[g], table_bytes, cudaMemcpyDeviceToHost));
        if (memcmp(first.data(), other.data(), table_bytes) != 0) { same = 0; fprintf(stderr, "lut_grid_multigpu: table of GPU %d differs\n", g); }
    }
    printf("%zu records on %d GPU(s); tables identical: %s\n", M, n, same ? "yes" : "NO");
    f = fopen(argv[2], "wb");
    if (!f || fwrite(first.data(), sizeof(double), first.size(), f) != first.size()) { perror(argv[2]); return EXIT_FAILURE; }
    fclose(f);
    for (int g = 0; g < n; g++) { gort_destroy(gx[g]); cudaSetDevice(g); cudaFree(table[g]); cudaFree(d_st[g]); }
    return same ? EXIT_SUCCESS : EXIT_FAILURE;
}
