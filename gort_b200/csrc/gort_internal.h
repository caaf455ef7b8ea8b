// gort_internal.h -- host-side internals of libgort_b200 (context, scratch buffers, launch
// prototypes).  Not part of the public ABI (include/gort_b200.h is).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "../../include/gort_b200.h"

#define GORT_NSCRATCH 16

struct gort_ctx {
    int device;
    cudaStream_t stream;
    char err[512];
    long launches;
    int sm_count;
    // grow-only device scratch used by the host-pointer entry points
    void *scratch[GORT_NSCRATCH];
    size_t scratch_cap[GORT_NSCRATCH];
    // grow-only device scratch used internally by _dev entry points (geometry records)
    void *work;
    size_t work_cap;
    // device constants built at creation
    double *d_gauleg;      // [2][32] abscissa, weights          (gauleg, gortt_albedo.c:142-198)
    double *d_prospect;    // [9][2101] refractive,k_Cab,k_Car,k_Anth,k_Brown,k_Cw,k_Cm, tav90, tav40
    double *d_soil;        // [4][421] Price EOF vectors
    // optional per-kernel event timing of the BRDF path (gort_profile_begin/end)
    cudaEvent_t *prof_ev;  // [3 * prof_cap]
    int prof_cap, prof_n;
};

namespace gort {

int set_error(gort_ctx *ctx, int code, const char *fmt, ...);
int check_cuda(gort_ctx *ctx, cudaError_t e, const char *what);
// returns device pointer with at least `bytes` capacity in scratch slot `slot`
void *scratch(gort_ctx *ctx, int slot, size_t bytes);
void *workspace(gort_ctx *ctx, size_t bytes);

// kernels (device pointers, async on `s`)
int launch_brdf(gort_ctx *ctx, cudaStream_t s, const gort_shape &sh, const double *structure,
                const double *lut, const double *angles, const double *rleaf, const double *tleaf,
                const double *rsoil, double *rsurf, double *scomp, double *kprop);
int launch_energy(gort_ctx *ctx, cudaStream_t s, const gort_shape &sh, const double *structure,
                  const double *lut, const double *angles, const double *rleaf, const double *tleaf,
                  const double *rsoil, double *albedo, double *favegt, double *fasoil);
int launch_lut(gort_ctx *ctx, cudaStream_t s, int n_sets, const double *structure, int method, double *lut);
int launch_spectra(gort_ctx *ctx, cudaStream_t s, int n_sets, const double *leaf, const double *soil,
                   double user_leaf, double user_soil, int n_wl, const double *wl,
                   double *rleaf, double *tleaf, double *rsoil);
int launch_prospect_full(gort_ctx *ctx, cudaStream_t s, int n_sets, const double *leaf, double *refl, double *tran);
int launch_gauleg(gort_ctx *ctx, cudaStream_t s, double *d_out /*[2][32]*/);
int launch_tav_tables(gort_ctx *ctx, cudaStream_t s, double *d_prospect);
int upload_soil_tables(gort_ctx *ctx, cudaStream_t s, double *d_soil);
int launch_dfma_peak(gort_ctx *ctx, cudaStream_t s, double *tflops);

}  // namespace gort
