// TMA bulk-store pattern sweep (B200): a CTA assembles `chunk` doubles of one row in shared memory (NBUF-deep
// ring) and one thread issues cp.async.bulk.global.shared::cta for the whole chunk.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bw tma_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NBUF>
__global__ void __launch_bounds__(1024) rows_tma(double* out, long pitch, int chunk, long lines_per_cta, long L, double v)
{
    extern __shared__ __align__(128) double buf[];            // [NBUF][chunk]
    long l0 = (long) blockIdx.y * lines_per_cta, l1 = min(L, l0 + lines_per_cta);
    int c0 = blockIdx.x * chunk;
    unsigned bytes = (unsigned) chunk * 8u;
    int b = 0;
    for (long l = l0; l < l1; l++) {
        double* s = buf + (size_t) b * chunk;
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(NBUF - 1) : "memory");
        __syncthreads();
        for (int k = threadIdx.x; k < chunk; k += blockDim.x) s[k] = v + l;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned sa = (unsigned) __cvta_generic_to_shared(s);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(out + l * pitch + c0), "r"(sa), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        b = (b + 1 == NBUF) ? 0 : b + 1;
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// warp-granular variant: every warp owns 32*LPT consecutive columns (LPT*256 bytes per row) and its own NBUF-deep
// ring in shared memory; lane 0 issues one bulk store per row.  No CTA-wide barrier anywhere.
template <int NBUF, int LPT>
__global__ void __launch_bounds__(384) rows_tma_warp(double* out, long pitch, int ncol, int chunk, long lines_per_cta, long L, double v)
{
    extern __shared__ __align__(128) double buf[];            // [warps][NBUF][32*LPT]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long l0 = (long) blockIdx.y * lines_per_cta, l1 = min(L, l0 + lines_per_cta);
    const int c0 = blockIdx.x * chunk + warp * 32 * LPT;
    const int nvalid = max(0, min(32 * LPT, ncol - c0));
    double* ring = buf + (size_t) warp * NBUF * 32 * LPT;
    double acc[LPT];
#pragma unroll
    for (int j = 0; j < LPT; j++) acc[j] = v + j;
    int b = 0;
    for (long l = l0; l < l1; l++) {
        double* s = ring + b * 32 * LPT;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(NBUF - 1) : "memory");
        __syncwarp();
#pragma unroll
        for (int j = 0; j < LPT; j++) { acc[j] = fma(acc[j], 1.0000001, 1e-9); s[32 * j + lane] = acc[j]; }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && nvalid > 0) {
            unsigned sa = (unsigned) __cvta_generic_to_shared(s);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(out + l * pitch + c0), "r"(sa), "r"(8u * nvalid) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        b = (b + 1 == NBUF) ? 0 : b + 1;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int NBUF, int LPT>
static void run_warp(double* out, long pitch, long L, int n_chunks, int threads, int occ)
{
    const int sms = 148;
    int chunk = LPT * threads;
    long nby = (long) sms * occ / n_chunks; long lpc = (L + nby - 1) / nby; nby = (L + lpc - 1) / lpc;
    size_t smem = sizeof(double) * NBUF * chunk;
    auto k = rows_tma_warp<NBUF, LPT>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    dim3 grid(n_chunks, (unsigned) nby);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        for (int i = 0; i < 3; i++) k<<<grid, threads, smem>>>(out, pitch, (int) pitch, chunk, lpc, L, 1.0);
        cudaEventRecord(e0);
        for (int i = 0; i < 20; i++) k<<<grid, threads, smem>>>(out, pitch, (int) pitch, chunk, lpc, L, 1.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20; if (ms < best) best = ms;
    }
    printf("warp-TMA LPT %d (%4d B/op) chunks %d thr %4d occ %d nbuf %d (%ld CTAs x %ld lines): %.1f us  %.0f GB/s  [%s]\n", LPT, LPT * 256, n_chunks, threads, occ, NBUF,
           n_chunks * nby, lpc, best * 1e3, (double) L * pitch * 8.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

template <int NBUF>
static void run(double* out, long pitch, long L, int n_chunks, int threads, int occ)
{
    const int sms = 148;
    int chunk = (int) (pitch / n_chunks);
    long nby = (long) sms * occ / n_chunks; long lpc = (L + nby - 1) / nby; nby = (L + lpc - 1) / lpc;
    size_t smem = sizeof(double) * NBUF * chunk;
    auto k = rows_tma<NBUF>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    dim3 grid(n_chunks, (unsigned) nby);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9, sum = 0;
    for (int rep = 0; rep < 3; rep++) {
        for (int i = 0; i < 3; i++) k<<<grid, threads, smem>>>(out, pitch, chunk, lpc, L, 1.0);
        cudaEventRecord(e0);
        for (int i = 0; i < 20; i++) k<<<grid, threads, smem>>>(out, pitch, chunk, lpc, L, 1.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20; sum += ms; if (ms < best) best = ms;
    }
    printf("chunks %d (%5d B) thr %4d occ %d nbuf %d (%ld CTAs x %ld lines): best %.1f us avg %.1f us  %.0f GB/s  [%s]\n", n_chunks, chunk * 8, threads, occ, NBUF,
           n_chunks * nby, lpc, best * 1e3, sum / 3 * 1e3, (double) L * pitch * 8.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const long L = 11664, pitch = 2112;
    double* out; cudaMalloc(&out, sizeof(double) * L * pitch);
    for (int n_chunks : {1, 2, 3, 4, 6}) for (int occ : {1, 2, 3, 4}) {
        int threads = (int) (pitch / n_chunks / 4); threads = (threads + 31) / 32 * 32; if (threads > 1024) threads = 1024; if (threads < 64) threads = 64;
        if (occ * threads > 2048) continue;
        run<2>(out, pitch, L, n_chunks, threads, occ);
        run<4>(out, pitch, L, n_chunks, threads, occ);
    }
    // warp-granular bulk stores
    run_warp<2, 4>(out, pitch, L, 3, 192, 2); run_warp<4, 4>(out, pitch, L, 3, 192, 2); run_warp<8, 4>(out, pitch, L, 3, 192, 2);
    run_warp<4, 4>(out, pitch, L, 2, 288, 2); run_warp<2, 4>(out, pitch, L, 2, 288, 2);
    run_warp<4, 4>(out, pitch, L, 3, 192, 3); run_warp<4, 4>(out, pitch, L, 3, 192, 1);
    run_warp<4, 3>(out, pitch, L, 2, 352, 2); run_warp<4, 2>(out, pitch, L, 3, 352, 2); run_warp<8, 2>(out, pitch, L, 3, 352, 2);
    run_warp<4, 8>(out, pitch, L, 2, 160, 2); run_warp<4, 8>(out, pitch, L, 1, 288, 2); run_warp<2, 8>(out, pitch, L, 1, 288, 2);
    return 0;
}
