// Microbenchmark of the rsurf_wide_kernel run loop (B200): per line 3 broadcast LDS.128 of coefficients, LPT
// 5-FMA chains on register-resident (sun, lambda) terms, LPT coalesced stores.  Variants isolate what costs
// store throughput: FP64 work, shared-memory loads, store width, unroll depth, warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o loop_bw loop_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int LPT, int UNROLL, bool WIDE16, bool COMPUTE, int MAP, int MAXT>
__global__ void __launch_bounds__(MAXT) loop(double* out, const double* __restrict__ coef, int W, long pitch, int chunk, long lines_per_cta, long L)
{
    extern __shared__ double2 srec[];                 // [lines_per_cta][3]
    long l0 = (long) blockIdx.y * lines_per_cta, l1 = min(L, l0 + lines_per_cta);
    int nl = (int) (l1 - l0);
    for (int i = threadIdx.x; i < nl * 3; i += blockDim.x) srec[i] = reinterpret_cast<const double2*>(coef)[l0 * 3 + i];
    __syncthreads();
    const int wbase = blockIdx.x * chunk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double sA[LPT], sP[LPT], sG[LPT], sZ[LPT], sT[LPT];
    int col[LPT];
#pragma unroll
    for (int j = 0; j < LPT; j++) {
        // 8-byte mapping: slot j -> column warp*32*LPT + 32*j + lane; 16-byte mapping: slots (2p, 2p+1) are adjacent columns
        // MAP 0: warp-contiguous (slot j of warp w at w*32*LPT + 32*j + lane); MAP 1: CTA-strided (slot j at tid + j*blockDim)
        col[j] = WIDE16 ? (MAP == 0 ? warp * 32 * LPT + (j / 2) * 64 + 2 * lane + (j & 1) : 2 * (int) threadIdx.x + (j / 2) * 2 * (int) blockDim.x + (j & 1))
                        : (MAP == 0 ? warp * 32 * LPT + 32 * j + lane : (int) threadIdx.x + j * (int) blockDim.x);
        double x = 1e-3 * (wbase + col[j]);
        sA[j] = x; sP[j] = x + 1; sG[j] = x + 2; sZ[j] = x + 3; sT[j] = x + 4;
    }
    bool ok[LPT];
#pragma unroll
    for (int j = 0; j < LPT; j++) ok[j] = wbase + col[j] < W;
    double* o = out + l0 * pitch + wbase;
    const double2* vr = srec;
#pragma unroll UNROLL
    for (int l = 0; l < nl; l++, vr += 3, o += pitch) {
        double r[LPT];
        if (COMPUTE) {
            const double2 v0 = vr[0], v1 = vr[1];
            const double cT = vr[2].x;
#pragma unroll
            for (int j = 0; j < LPT; j++) r[j] = fma(v0.x, sA[j], fma(v0.y, sP[j], fma(v1.x, sG[j], fma(v1.y, sZ[j], cT * sT[j]))));
        } else {
#pragma unroll
            for (int j = 0; j < LPT; j++) r[j] = sA[j] + l;
        }
        if (WIDE16) {
#pragma unroll
            for (int j = 0; j < LPT; j += 2) if (ok[j]) *reinterpret_cast<double2*>(o + col[j]) = make_double2(r[j], r[j + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < LPT; j++) if (ok[j]) o[col[j]] = r[j];
        }
    }
}

template <int LPT, int UNROLL, bool WIDE16, bool COMPUTE, int MAP, int MAXT = 256>
static void run(const char* name, double* out, double* coef, int W, long pitch, long L, int occ, int n_chunks_force = 0, int interleave = 0)
{
    const int sms = 148;
    int n_chunks = n_chunks_force ? n_chunks_force : (W + LPT * 256 - 1) / (LPT * 256);
    int threads = (W + n_chunks * LPT - 1) / (n_chunks * LPT); threads = (threads + 31) / 32 * 32;
    int chunk = LPT * threads;
    long nby = (long) sms * occ / n_chunks; long lpc = (L + nby - 1) / nby; nby = (L + lpc - 1) / lpc;
    size_t smem = sizeof(double2) * 3 * lpc;
    auto k = loop<LPT, UNROLL, WIDE16, COMPUTE, MAP, MAXT>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    dim3 grid(n_chunks, (unsigned) nby);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) k<<<grid, threads, smem>>>(out, coef, W, pitch, chunk, lpc, L);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; i++) k<<<grid, threads, smem>>>(out, coef, W, pitch, chunk, lpc, L);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
    printf("%-26s map %d LPT %d unroll %d occ %d (%d thr, %d x %ld CTAs, %ld lines): %.1f us  %.0f GB/s  [%s]\n", name, MAP, LPT, UNROLL, occ, threads,
           n_chunks, nby, lpc, ms * 1e3, (double) L * W * 8.0 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const long L = 11664, pitch = 2112;
    double *out, *coef; cudaMalloc(&out, sizeof(double) * L * pitch); cudaMalloc(&coef, sizeof(double) * 6 * L);
    cudaMemset(coef, 0, sizeof(double) * 6 * L);
    for (int W : {2101, 2112}) {
        printf("---- W = %d of pitch %ld written\n", W, pitch);
        for (int occ : {2, 3}) {
            run<4, 4, false, false, 0>("stores only, 8 B", out, coef, W, pitch, L, occ);
            run<4, 4, false, false, 1>("stores only, 8 B", out, coef, W, pitch, L, occ);
            run<4, 4, false, true, 0>("5-FMA + LDS, 8 B", out, coef, W, pitch, L, occ);
            run<4, 4, false, true, 1>("5-FMA + LDS, 8 B", out, coef, W, pitch, L, occ);
            run<4, 4, true, true, 1>("5-FMA + LDS, 16 B", out, coef, W, pitch, L, occ);
            run<2, 4, false, true, 1>("5-FMA + LDS, 8 B", out, coef, W, pitch, L, occ);
            run<2, 4, true, true, 1>("5-FMA + LDS, 16 B", out, coef, W, pitch, L, occ);
        }
        run<3, 4, false, true, 1, 704>("full row 704 thr", out, coef, W, pitch, L, 1, 1);
        run<3, 4, false, true, 0, 704>("full row 704 thr", out, coef, W, pitch, L, 1, 1);
        run<3, 4, false, true, 1, 352>("half row 352 thr", out, coef, W, pitch, L, 2, 2);
        run<3, 4, false, true, 0, 352>("half row 352 thr", out, coef, W, pitch, L, 2, 2);
    }
    return 0;
}
