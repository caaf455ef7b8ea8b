// gort_rsurf_rows.cuh -- the per-wavelength loop of gortt_rsurf (gortt.c:460-567) for FULL-SPECTRUM calls: whole
// output rows per CTA, one persistent CTA per SM.  This is the kernel the C2 sweep (11 664 lines x 2101 bands)
// runs; gort_rsurf_wide.cuh remains the general form (any W >= 64, component signatures, unaligned rows).
//
// Why a second decomposition.  The per-wavelength kernel is bound by HBM writes (8 B per evaluation, DESIGN.md
// 4.1).  Measured with a pure store stream of C2's shape (tools/microbench/tma_bw.cu, profiles/r1_microbench_tma_bw.txt):
//      3 column chunks x 2 CTAs per SM, 5.6 KB bulk stores (the wide kernel's shape)   33.0 us   5.97 TB/s
//      1 CTA per SM writing whole 16.9 KB rows of a CONTIGUOUS block of lines           31.2 us   6.32 TB/s
// 148 long sequential streams are kinder to HBM than 296 strided ones.  And with the whole spectrum in one CTA the
// (set, lambda) table -- exp, sqrt and five divisions per wavelength, which every one of the wide kernel's ~100 CTAs
// per chunk recomputes (2.7 us of FP64 pipe per CTA before its first store) -- is computed ONCE per call by spare
// CTAs of the geometry kernel (leaf_table_tile below) and arrives in shared memory as one TMA bulk copy.
//
// Shared memory (1 CTA per SM): table [9][ncolt] (153 KB for 2101 bands) | row ring [NBUF][ncolt] | line records
// [128][16] | mbarriers.  ncolt = 4 x blockDim.x >= n_col: thread t owns columns kq + 32 j, j < 4, kq =
// (t / 32) * 128 + t % 32, so every warp's j-th shared-memory store is one conflict-free 256-byte segment.
//
// Per line and compute warp: 3 broadcast LDS.128 of the record, 5 FMAs per column (the regrouped form of gortt.c:557,
// see gort_rsurf_wide.cuh), 4 STS into the ring slot, a proxy fence and ONE mbarrier arrival; a dedicated store thread
// turns every full slot into one cp.async.bulk (SASS UBLKCP) of the whole row and frees the slot once the TMA engine
// has read it out (cp.async.bulk.wait_group.read).  No CTA-wide barrier in the line loop.
//
// Pipeline across kernels and calls: identical contract to the wide kernel (per-tile ready flags instead of grid
// completion under programmatic dependent launch; per-CTA epoch gate before the first store; bounded waits).
#pragma once
#include "gort_device.cuh"

namespace gort {

#define ROWS_MAX_THREADS 544      // 17 warps: up to 2176 columns (table + ring + records = 220 KB of shared memory)
#define ROWS_STAGE_LINES 128      // line records staged per pass (16 KB)
#define ROWS_NLEAF 9              // omega gam Tff Rff pff tff rs Xf A  (same order as the wide kernel's table)
#define ROWS_NBUF 3               // ring slots (rows)
#define ROWS_TAB_TILE 128         // columns per table tile (one flag each)

struct RowsArgs {
    int n_sets, n_geom, n_wl, spectra_per_set;
    int n_col;                    // columns stored per row (a multiple of 2: bulk copies move 16-byte units)
    int ncolt;                    // table / ring row length, 4 * blockDim.x
    int pdl;
    int dbg;                      // development only (GORT_ROWS_DBG): 1 = issue no stores, 2 = skip the row arithmetic
    const unsigned long long *flags;    // this call's flag array: [geometry tiles | table tiles]
    long tab_flag_base;                 // index of the first table-tile flag
    unsigned long long call_no;
    unsigned long long *fault, *done;
    unsigned long long wait_target, epoch;
    long pitch, lines_per_cta;
    const double *lut, *rec, *table;    // table [n_sets][ROWS_NLEAF][ncolt]
    double *rsurf;
    unsigned long long *tl;             // optional timeline (development aid, GORT_TIMELINE)
};

// ---- (set, lambda) table tile: called by the spare CTAs of the geometry kernels -----------------------------
// One thread per (set, column) of table [n_sets][9][ncolt]; columns >= n_wl repeat column n_wl - 1 (they are the
// row padding the output kernels fill, and the unused tail of the last warp).  Publishes one flag per tile.
__device__ __forceinline__ void leaf_table_tile(long tile, int n_sets, int n_wl, int spectra_per_set, int ncolt,
                                                const double* __restrict__ structure, const double* __restrict__ lut,
                                                const double* __restrict__ rleaf, const double* __restrict__ tleaf,
                                                const double* __restrict__ rsoil, double* __restrict__ table,
                                                unsigned long long* __restrict__ tab_flags, unsigned long long call_no)
{
    const int tiles_per_set = ncolt / ROWS_TAB_TILE;
    const int m = (int) (tile / tiles_per_set);
    const int k = (int) (tile - (long) m * tiles_per_set) * ROWS_TAB_TILE + threadIdx.x;
    if (threadIdx.x < ROWS_TAB_TILE && m < n_sets) {
        const Canopy c = canopy_load(structure, n_sets, m, lut);
        const size_t sb = (spectra_per_set ? (size_t) m * n_wl : 0) + min(k, n_wl - 1);
        const LeafTerms Lf = leaf_terms(c, rleaf[sb], tleaf[sb], rsoil[sb]);
        double* t = table + (size_t) m * ROWS_NLEAF * ncolt + k;
        t[0 * (size_t) ncolt] = Lf.omega; t[1 * (size_t) ncolt] = Lf.gam; t[2 * (size_t) ncolt] = Lf.Tff;
        t[3 * (size_t) ncolt] = Lf.Rff;   t[4 * (size_t) ncolt] = Lf.pff; t[5 * (size_t) ncolt] = Lf.tff;
        t[6 * (size_t) ncolt] = Lf.rs;    t[7 * (size_t) ncolt] = Lf.Xf;  t[8 * (size_t) ncolt] = Lf.A;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(tab_flags + tile), "l"(call_no) : "memory");
    }
}

__device__ __forceinline__ bool rows_mbar_try_wait(unsigned mbar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    return ok != 0;
}

// bounded acquire-poll of one flag; returns false if the wait expired (2 s)
__device__ __forceinline__ bool rows_wait_flag(const unsigned long long* p, unsigned long long target)
{
    unsigned long long v, t0 = 0, t1;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (v >= target) return true;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        __nanosleep(100);
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        if (v >= target) return true;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 2000000000ull) return false;
    }
}

// mbarrier helpers (shared::cta addresses)
__device__ __forceinline__ void rows_mbar_wait(unsigned mbar, unsigned parity)
{
    while (!rows_mbar_try_wait(mbar, parity)) { }
}
// Warp-wide wait with ONE polling lane.  With every thread of 17 warps spinning on the same shared-memory word the
// polls themselves serialise in the shared-memory pipeline and every arrival queues behind them: the empty kernel
// (no arithmetic, no stores) took 0.39 us per row that way.
__device__ __forceinline__ void rows_mbar_wait_warp(unsigned mbar, unsigned parity)
{
    if ((threadIdx.x & 31) == 0) rows_mbar_wait(mbar, parity);
    __syncwarp();
}
__device__ __forceinline__ void rows_mbar_arrive(unsigned mbar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar) : "memory");
}

// Warp roles: warps 0 .. ncw-1 compute (ncw = ncolt / 128, 128 columns each), the last warp's lane 0 is the STORE
// thread.  The two sides meet only through mbarriers, one pair per ring slot:
//     full[slot]   ncw arrivals (lane 0 of every compute warp, after the warp's STS + proxy fence)  -> store thread
//     empty[slot]  1 arrival (store thread, once the TMA engine has read the slot out)             -> compute warps
// so no warp ever waits for another compute warp: a warp may be up to two rows ahead of the slowest one, and the
// dependent chain of a row (record LDS -> 5 FMAs -> STS -> fence) overlaps across warps instead of being paid once
// per row by the whole CTA (the first version of this kernel, one CTA barrier per row, ran at 0.6 us per row against
// the 0.4 us HBM needs).  Compute warps synchronise among themselves (named barrier 1) only to restage records or to
// change the table -- once per CTA in C2.
__global__ void __launch_bounds__(ROWS_MAX_THREADS + 32, 1)
rsurf_rows_kernel(const RowsArgs a)
{
    constexpr int STAGE = ROWS_STAGE_LINES;
    constexpr int NBUF = ROWS_NBUF;
#define ROWS_TL(k) do { if (a.tl) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.tl[(size_t) blockIdx.x * 8 + (k)] = t_; } } while (0)
    extern __shared__ __align__(128) unsigned char rows_smem[];
    const int ncolt = a.ncolt;
    double* tab = reinterpret_cast<double*>(rows_smem);                          // [9][ncolt]
    double* ring = tab + (size_t) ROWS_NLEAF * ncolt;                            // [NBUF][ncolt]
    double2* srec = reinterpret_cast<double2*>(ring + (size_t) NBUF * ncolt);    // [STAGE][8]
    unsigned long long* mbar_p = reinterpret_cast<unsigned long long*>(srec + 8 * STAGE);   // load, full[NBUF], empty[NBUF]
    __shared__ volatile int s_fault;

    const long L = (long) a.n_sets * a.n_geom;
    const unsigned cta = blockIdx.x;
    const long line_begin = (long) blockIdx.x * a.lines_per_cta;
    const long line_end = min(L, line_begin + a.lines_per_cta);
    const int tid = threadIdx.x;
    const int ncw = ncolt / 128;                                  // compute warps
    const int ncomp = ncw * 32;                                   // compute threads
    const unsigned mbar_load = (unsigned) __cvta_generic_to_shared(mbar_p);
    const unsigned mbar_full = mbar_load + 8, mbar_empty = mbar_load + 8 + 8 * NBUF;
    const unsigned tab_bytes = (unsigned) (sizeof(double) * ROWS_NLEAF * (size_t) ncolt);
    const int tiles_per_set = ncolt / ROWS_TAB_TILE;
    const unsigned row_bytes = 8u * (unsigned) a.n_col;

    if (line_begin >= line_end) {
        if (tid == 0) { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); a.done[cta] = a.epoch; }
        return;
    }
    if (tid == 0) {
        ROWS_TL(0);
        s_fault = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar_load) : "memory");
        for (int b = 0; b < NBUF; b++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar_full + 8 * b), "r"(ncw) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar_empty + 8 * b) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= ncomp) {
        // ================= store thread =================
        if (tid != ncomp) return;
        const long nrows = line_end - line_begin;
        bool gate_open = false;
        for (long i = 0; i < nrows; i++) {
            const int slot = (int) (i % NBUF);
            rows_mbar_wait(mbar_full + 8 * slot, (unsigned) ((i / NBUF) & 1));
            if (!gate_open) {
                // cross-call gate, as in the wide kernel: before the first store wait until the CTA with the same
                // index in the previous launch of the same shape and outputs (it wrote this very region) has
                // finished, then release the dependents (the next call's geometry kernel)
                if (!rows_wait_flag(a.done + cta, a.wait_target)) { s_fault = 1; *a.fault = a.call_no; }
                asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
                gate_open = true;
                ROWS_TL(4);
            }
            if (!s_fault && !(a.dbg & 1)) {                       // a producer never published: store nothing
                const unsigned sa = (unsigned) __cvta_generic_to_shared(ring + (size_t) slot * ncolt);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             :: "l"(a.rsurf + (size_t) (line_begin + i) * a.pitch), "r"(sa), "r"(row_bytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (i >= 1) {
                // all but the newest group have been read out of shared memory: row i-1's slot is free again
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                rows_mbar_arrive(mbar_empty + 8 * (int) ((i - 1) % NBUF));
            }
        }
        // publish: all bulk stores of this CTA complete, then fence + release
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        ROWS_TL(5);
        __threadfence();
        asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(a.done + cta), "l"(a.epoch) : "memory");
        return;
    }

    // ================= compute warps =================
#define ROWS_CBAR() asm volatile("bar.sync 1, %0;" :: "r"(ncomp) : "memory")
    const int kq = (tid >> 5) * 128 + (tid & 31);
    unsigned parity = 0;
    int m = (int) (line_begin / a.n_geom);
    long set_end = (long) (m + 1) * a.n_geom;
    double k_open = 0.0, ke = 0.0;
    bool need_table = true;       // the table of set m is not in shared memory yet
    bool first_sun = true;
    long cnt = 0;                 // rows this CTA has produced

    double sA[4], sP[4], sG[4], sZ[4], sT[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { sA[j] = sP[j] = sG[j] = sZ[j] = sT[j] = 0.0; }

    for (long s0 = line_begin; s0 < line_end; s0 += STAGE) {
        const int nl = (int) min((long) STAGE, line_end - s0);
        // ---- stage the packed records of lines [s0, s0 + nl) (and, the first time, the table of the first set):
        //      acquire the producers' flags, then ONE thread arms the mbarrier and issues the TMA bulk loads ----
        {
            const long tile0 = s0 >> 5, tile1 = (s0 + nl - 1) >> 5;
            bool ok = true;
            for (long t = tile0 + tid; t <= tile1; t += ncomp) ok &= rows_wait_flag(a.flags + t, a.call_no);
            if (need_table)
                for (int t = tid; t < tiles_per_set; t += ncomp)
                    ok &= rows_wait_flag(a.flags + a.tab_flag_base + (long) m * tiles_per_set + t, a.call_no);
            if (!ok) { s_fault = 1; *a.fault = a.call_no; }
            ROWS_CBAR();                                          // flags acquired; previous stage fully consumed
            if (tid == 0) {
                const unsigned rec_bytes = (unsigned) nl * 128u;
                const unsigned total = rec_bytes + (need_table ? tab_bytes : 0u);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar_load), "r"(total) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"((unsigned) __cvta_generic_to_shared(srec)), "l"(a.rec + (size_t) s0 * GORT_REC_STRIDE),
                                "r"(rec_bytes), "r"(mbar_load) : "memory");
                if (need_table)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"((unsigned) __cvta_generic_to_shared(tab)),
                                    "l"(a.table + (size_t) m * ROWS_NLEAF * ncolt), "r"(tab_bytes), "r"(mbar_load) : "memory");
            }
            if (need_table) {
                const double* l = a.lut + (size_t) m * GORT_LUT_STRIDE;
                k_open = l[2 * GORT_NTH]; ke = l[2 * GORT_NTH + 1];
            }
            rows_mbar_wait_warp(mbar_load, parity);
            parity ^= 1u;
            need_table = false;
            if (s0 == line_begin && tid == 0) { ROWS_TL(1); ROWS_TL(2); }      // table + first records in shared memory
        }

        for (int l = 0; l < nl; l++) {
            int f = (int) __double_as_longlong(srec[8 * l + 2].y);
            if (s0 + l == line_begin) f = 3;
            if ((f & 2) && s0 + l >= set_end) {
                // ---- next parameter set: its table replaces the current one ----
                m++; set_end += a.n_geom;
                bool ok = true;
                for (int t = tid; t < tiles_per_set; t += ncomp)
                    ok &= rows_wait_flag(a.flags + a.tab_flag_base + (long) m * tiles_per_set + t, a.call_no);
                if (!ok) { s_fault = 1; *a.fault = a.call_no; }
                ROWS_CBAR();                                      // every compute warp is done with the old table
                if (tid == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar_load), "r"(tab_bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"((unsigned) __cvta_generic_to_shared(tab)),
                                    "l"(a.table + (size_t) m * ROWS_NLEAF * ncolt), "r"(tab_bytes), "r"(mbar_load) : "memory");
                }
                const double* lt = a.lut + (size_t) m * GORT_LUT_STRIDE;
                k_open = lt[2 * GORT_NTH]; ke = lt[2 * GORT_NTH + 1];
                rows_mbar_wait_warp(mbar_load, parity);
                parity ^= 1u;
            }
            if (f & 1) {
                // ---- new sun: (sun, lambda) terms of this thread's four columns into registers ----
                const double fd = srec[8 * l + 3].y;
                const double2 s0v = srec[8 * l + 4], s1v = srec[8 * l + 5];       // (mus,t0) (tp0,pe_s)
                Canopy c;
                c.k_open = k_open; c.k_openep = ke;
                const double K = k_open + ke;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int k = kq + 32 * j;                    // < ncolt
                    LeafTerms Lf;
                    Lf.omega = tab[0 * ncolt + k]; Lf.gam = tab[1 * ncolt + k]; Lf.Tff = tab[2 * ncolt + k];
                    Lf.Rff = tab[3 * ncolt + k];   Lf.pff = tab[4 * ncolt + k]; Lf.tff = tab[5 * ncolt + k];
                    Lf.rs = tab[6 * ncolt + k];    Lf.Xf = tab[7 * ncolt + k];
                    Lf.tpff = Lf.tff * (1.0 - K) + K;             // gortt_brdf.c:381-382, as in leaf_terms
                    Lf.Zf = (Lf.tpff - ke) * Lf.rs;               // gortt.c:492
                    const SunTerms S = sun_terms(c, Lf, fd, s0v.x, s0v.y, s1v.x, s1v.y);
                    sA[j] = tab[8 * ncolt + k];
                    sP[j] = S.PDF; sG[j] = S.G; sZ[j] = S.Z; sT[j] = S.T;
                }
                if (first_sun && tid == 0) ROWS_TL(3);            // first sun terms in registers
                first_sun = false;
            }
            // ---- the line: wait for the ring slot, 5 FMAs per column into it, hand it to the store thread ----
            const int slot = (int) (cnt % NBUF);
            if (cnt >= NBUF) rows_mbar_wait_warp(mbar_empty + 8 * slot, (unsigned) ((cnt / NBUF - 1) & 1));
            if (!(a.dbg & 2)) {
                const double2* vr = srec + 8 * l;
                const double2 v0 = vr[0], v1 = vr[1];              // (cA,Kc) (cG,cZ)
                const double cT = vr[2].x;                         // Kt
                double* sb = ring + (size_t) slot * ncolt + kq;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    sb[32 * j] = fma(v0.x, sA[j], fma(v0.y, sP[j], fma(v1.x, sG[j], fma(v1.y, sZ[j], cT * sT[j]))));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if ((tid & 31) == 0) rows_mbar_arrive(mbar_full + 8 * slot);
            cnt++;
        }
    }
#undef ROWS_CBAR
#undef ROWS_TL
}

}  // namespace gort
