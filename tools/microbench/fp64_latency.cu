// FP64 pipe microbenchmark for B200 (sm_100a): dependent-chain latency and throughput versus
// independent chains per warp (ILP) and warps per SM (TLP).  nvcc -arch=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, long long* cyc, int iters, double m, double b)
{
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = fma(a[i], m, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm, int sms)
{
    double* out; long long* cyc;
    int threads = 32 * warps_per_sm, iters = 4096;
    cudaMalloc(&out, sizeof(double) * sms * threads); cudaMalloc(&cyc, sizeof(long long) * sms);
    chain<ILP><<<sms, threads>>>(out, cyc, iters, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    chain<ILP><<<sms, threads>>>(out, cyc, iters, 0.999999, 1e-7);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    double per = (double) c / iters;                      // cycles per loop iteration (ILP dfmas per warp)
    double tf = 2.0 * ILP * iters * (double) sms * threads / (ms * 1e-3) / 1e12;
    printf("ILP %d warps/SM %2d: %.2f cycles per iteration, %.2f cycles per DFMA per warp, %.2f TFLOP/s\n", ILP,
           warps_per_sm, per, per / ILP, tf);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    int ws[] = {1, 4, 8, 16, 32};
    for (int w : ws) { run<1>(w, sms); run<2>(w, sms); run<4>(w, sms); run<8>(w, sms); }
    return 0;
}
