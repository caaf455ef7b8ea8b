// gort_rsurf_wide.cuh -- the per-wavelength loop of gortt_rsurf (gortt.c:460-567) for wide spectra
// (W >= 64), the dominant kernel of the BRDF path.
//
// Work decomposition (grid.x = wavelength chunks, grid.y = contiguous line ranges):
//   * a CTA owns one wavelength chunk (LPT wavelengths per thread, lanes = consecutive wavelengths,
//     so every warp store is a contiguous 256-byte run of one output row) and walks a CONTIGUOUS
//     range of input lines, sized so that the whole grid is one resident wave;
//   * the packed 128-byte line records written by geom_kernel are staged into shared memory
//     WIDE_STAGE_LINES at a time (a straight coalesced copy) and cut into "runs" of lines that share
//     parameter set and sun (flags computed by geom_kernel from the input angles);
//   * the (set, lambda) terms (exp, sqrt, four divides) are computed once per set into shared memory;
//   * the (sun, lambda) terms are re-derived only at the start of a run and then live in registers:
//     A, PD, FCf, G, Z, T;
//   * inside a run only the view-dependent part remains per (line, lambda): 4 broadcast LDS.128 per
//     line, 9 FP64 instructions and one coalesced 8-byte store per evaluation, no branches.
// HBM traffic is 8 B per evaluation (rsurf) -- the binding roofline for this kernel.
// Output rows are `pitch` doubles apart; a pitch that is a multiple of 4 doubles keeps every warp
// store sector-aligned (measured: 36 us vs 54 us per 196 MB for the same store stream, tools/microbench).
#pragma once
#include "gort_device.cuh"

namespace gort {

#define WIDE_STAGE_LINES 256      // lines staged in shared memory per pass (32 KB)
#define WIDE_NLEAF 11             // omega gam Tff Rff pff tff tpff rs Xf Zf A

struct WideArgs {
    int n_sets, n_geom, n_wl, spectra_per_set;
    int chunk;                    // wavelengths per CTA (= LPT * blockDim.x)
    long pitch;                   // output row stride in doubles (>= n_wl)
    long lines_per_cta;
    const double *structure, *lut, *rec, *rleaf, *tleaf, *rsoil;
    double *rsurf, *scomp;
};

template <int LPT, bool SCOMP, int MINB>
__global__ void __launch_bounds__(256, MINB)
rsurf_wide_kernel(const WideArgs a)
{
    extern __shared__ double2 smem2[];
    double2* srec = smem2;                                                    // [WIDE_STAGE_LINES][8] packed records
    int* runend = reinterpret_cast<int*>(srec + 8 * WIDE_STAGE_LINES);        // [WIDE_STAGE_LINES]
    double* leaf = reinterpret_cast<double*>(runend + WIDE_STAGE_LINES);      // [WIDE_NLEAF][chunk]

    const long L = (long) a.n_sets * a.n_geom;
    const long line_begin = (long) blockIdx.y * a.lines_per_cta;
    const long line_end = min(L, line_begin + a.lines_per_cta);
    if (line_begin >= line_end) return;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int wbase = blockIdx.x * a.chunk;
    const int chunk = a.chunk;

    int m = (int) (line_begin / a.n_geom);
    long set_end = (long) (m + 1) * a.n_geom;                     // first line of the next set
    double k_open = 0.0, ke = 0.0;

    double sA[LPT], sPD[LPT], sFCf[LPT], sG[LPT], sZ[LPT], sT[LPT];
#pragma unroll
    for (int j = 0; j < LPT; j++) { sA[j] = sPD[j] = sFCf[j] = sG[j] = sZ[j] = sT[j] = 0.0; }
    // column of slot j; slots past the end of the spectrum are clamped to the last wavelength: they
    // recompute and re-store that value (same address, same bits), which keeps the inner loop free of
    // predicates and branches
    int wq[LPT];
#pragma unroll
    for (int j = 0; j < LPT; j++) wq[j] = min(wbase + tid + j * nthr, a.n_wl - 1);   // chunk == LPT * nthr

    for (long s0 = line_begin; s0 < line_end; s0 += WIDE_STAGE_LINES) {
        const int nl = (int) min((long) WIDE_STAGE_LINES, line_end - s0);
        __syncthreads();                                          // previous stage fully consumed
        {   // ---- stage the packed records of lines [s0, s0+nl): 8 double2 per line, coalesced ----
            const double2* g = reinterpret_cast<const double2*>(a.rec + (size_t) s0 * GORT_REC_STRIDE);
            for (int i = tid; i < nl * 8; i += nthr) srec[i] = __ldg(g + i);
        }
        __syncthreads();
        // a "run" = a maximal stretch of lines sharing set and sun: starts at line 0 of the stage or at a
        // flagged line and ends before the next flagged line
        for (int i = tid; i < nl; i += nthr) {
            const int f = (int) __double_as_longlong(srec[8 * i + 6].x);
            if (i == 0 || f != 0) {
                int e = i + 1;
                while (e < nl && (int) __double_as_longlong(srec[8 * e + 6].x) == 0) e++;
                runend[i] = e;
            }
        }
        __syncthreads();

        int l = 0;
        while (l < nl) {
            int f = (int) __double_as_longlong(srec[8 * l + 6].x);
            if (s0 + l == line_begin) f = 3;                      // a CTA's first line starts everything
            if (f & 2) {
                // ---- new parameter set: (set, lambda) terms of this chunk into shared memory ----
                const long ln = s0 + l;
                if (ln >= set_end) { m++; set_end += a.n_geom; }  // lines advance one run at a time within a set
                const Canopy c = canopy_load(a.structure, a.n_sets, m, a.lut);
                k_open = c.k_open; ke = c.k_openep;
                const size_t sb = a.spectra_per_set ? (size_t) m * a.n_wl : 0;
                __syncthreads();                                  // everyone is done reading the old leaf terms
                for (int k = tid; k < chunk; k += nthr) {
                    const int w = min(wbase + k, a.n_wl - 1);
                    LeafTerms Lf = leaf_terms(c, a.rleaf[sb + w], a.tleaf[sb + w], a.rsoil[sb + w]);
                    leaf[0 * chunk + k] = Lf.omega; leaf[1 * chunk + k] = Lf.gam; leaf[2 * chunk + k] = Lf.Tff;
                    leaf[3 * chunk + k] = Lf.Rff;   leaf[4 * chunk + k] = Lf.pff; leaf[5 * chunk + k] = Lf.tff;
                    leaf[6 * chunk + k] = Lf.tpff;  leaf[7 * chunk + k] = Lf.rs;  leaf[8 * chunk + k] = Lf.Xf;
                    leaf[9 * chunk + k] = Lf.Zf;    leaf[10 * chunk + k] = Lf.A;
                }
                __syncthreads();
            }
            if (f & 1) {
                // ---- new sun: (sun, lambda) terms into registers ----
                const double fd = srec[8 * l + 3].y;
                const double2 s0v = srec[8 * l + 4], s1v = srec[8 * l + 5];       // (mus,t0) (tp0,pe_s)
                Canopy c;
                c.k_open = k_open; c.k_openep = ke;
#pragma unroll
                for (int j = 0; j < LPT; j++) {
                    const int k = tid + j * nthr;                 // < chunk == LPT * nthr
                    LeafTerms Lf;
                    Lf.omega = leaf[0 * chunk + k]; Lf.gam = leaf[1 * chunk + k]; Lf.Tff = leaf[2 * chunk + k];
                    Lf.Rff = leaf[3 * chunk + k];   Lf.pff = leaf[4 * chunk + k]; Lf.tff = leaf[5 * chunk + k];
                    Lf.tpff = leaf[6 * chunk + k];  Lf.rs = leaf[7 * chunk + k];  Lf.Xf = leaf[8 * chunk + k];
                    Lf.Zf = leaf[9 * chunk + k];
                    SunTerms S = sun_terms(c, Lf, fd, s0v.x, s0v.y, s1v.x, s1v.y);
                    sA[j] = leaf[10 * chunk + k];
                    sPD[j] = S.PD; sFCf[j] = S.FCf; sG[j] = S.G; sZ[j] = S.Z; sT[j] = S.T;
                }
            }
            // ---- the run: only the view-dependent part per (line, lambda) ----
            const int e = runend[l];
            double* out[LPT];
#pragma unroll
            for (int j = 0; j < LPT; j++) out[j] = a.rsurf + (size_t) (s0 + l) * a.pitch + wq[j];
            const double2* vr = srec + 8 * l;
#pragma unroll 2
            for (; l < e; l++, vr += 8) {
                const double2 v0 = vr[0], v1 = vr[1], v2 = vr[2], v3 = vr[3];   // (Kc,Kg) (Kt,Kz) (K'g,K'z) (q,fd)
#pragma unroll
                for (int j = 0; j < LPT; j++) {
                    const double zg = fma(sG[j], v2.x, sZ[j] * v2.y);             // Z*K'z + G*K'g   gortt.c:514
                    double Cd = fma(sA[j], v3.x, sPD[j]);                         // CdC + CdCG      gortt.c:504-507
                    Cd = fma(ke, zg, Cd);                                         // + CdG           gortt.c:528
                    const double C = fma(v3.y, Cd, sFCf[j]);                      // gortt.c:531
                    const double r = fma(v1.y, sZ[j], fma(v1.x, sT[j], fma(v0.y, sG[j], v0.x * C)));   // gortt.c:557
                    *out[j] = r;
                    if (SCOMP) {
                        double* sc = a.scomp + 4 * (out[j] - a.rsurf);
                        *reinterpret_cast<double2*>(sc) = make_double2(C, sG[j]);
                        *reinterpret_cast<double2*>(sc + 2) = make_double2(sT[j], sZ[j]);
                    }
                    out[j] += a.pitch;
                }
            }
        }
    }
}

}  // namespace gort
