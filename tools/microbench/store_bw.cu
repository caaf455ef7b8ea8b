// HBM write-stream microbenchmark (B200): how fast can N warps/SM stream FP64 results to HBM?
// Shapes mimic rsurf_wide_kernel: rows of W doubles, a warp writes 256 contiguous bytes per store.
#include <cstdio>
#include <cuda_runtime.h>

// each CTA owns a column chunk [c0, c0+chunk) and walks `lines` rows; LPT stores per thread per row
template <int LPT, int MODE>
__global__ void rows(double* out, int W, int chunk, long lines_per_cta, long L, double v)
{
    long l0 = (long) blockIdx.y * lines_per_cta, l1 = min(L, l0 + lines_per_cta);
    int c0 = blockIdx.x * chunk;
    for (long l = l0; l < l1; l++) {
#pragma unroll
        for (int j = 0; j < LPT; j++) {
            int w = min(c0 + (int) threadIdx.x + j * (int) blockDim.x, W - 1);
            if (MODE == 0) out[l * W + w] = v + l;
            else if (MODE == 1) __stcs(out + l * W + w, v + l);
            else __stwt(out + l * W + w, v + l);
        }
    }
}

// TMA variant: the CTA assembles its chunk of one row in shared memory (double-buffered) and one thread
// issues a bulk asynchronous store (cp.async.bulk.global.shared::cta -> SASS UBLKCP) of the whole chunk.
// Requires 16-byte aligned rows (W % 2 == 0) and chunk sizes.
template <int LPT>
__global__ void rows_tma(double* out, int W, int chunk, long lines_per_cta, long L, double v)
{
    extern __shared__ __align__(128) double buf[];            // [2][chunk]
    long l0 = (long) blockIdx.y * lines_per_cta, l1 = min(L, l0 + lines_per_cta);
    int c0 = blockIdx.x * chunk;
    int n = min(chunk, W - c0);                               // doubles of this chunk inside the row
    unsigned bytes = (unsigned) n * 8u;
    int b = 0;
    for (long l = l0; l < l1; l++, b ^= 1) {
        double* s = buf + (size_t) b * chunk;
        // the bulk store issued two rows ago read this buffer: wait until at most 1 group is still reading
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int j = 0; j < LPT; j++) s[threadIdx.x + j * blockDim.x] = v + l;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned sa = (unsigned) __cvta_generic_to_shared(s);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(out + l * W + c0), "r"(sa), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void flat(double* out, long n, double v)
{
    long i = (long) blockIdx.x * blockDim.x + threadIdx.x, st = (long) gridDim.x * blockDim.x;
    for (; i < n; i += st) out[i] = v;
}

__global__ void flat2(double2* out, long n2, double v)
{
    long i = (long) blockIdx.x * blockDim.x + threadIdx.x, st = (long) gridDim.x * blockDim.x;
    for (; i < n2; i += st) out[i] = make_double2(v, v);
}

static float timeit(void (*launch)(void*), void* ctx)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) launch(ctx);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; i++) launch(ctx);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 20;
}

struct Cfg { double* out; int W, chunk, threads, lpt; long L, lines_per_cta; dim3 grid; long n; int blocks; };
static int g_mode = 0;
static void l_rows2(void* p) { Cfg* c = (Cfg*) p; if (g_mode == 1) { rows<2,1><<<c->grid, c->threads>>>(c->out, c->W, c->chunk, c->lines_per_cta, c->L, 1.0); return; } if (g_mode == 2) { rows<2,2><<<c->grid, c->threads>>>(c->out, c->W, c->chunk, c->lines_per_cta, c->L, 1.0); return; } rows<2,0><<<c->grid, c->threads>>>(c->out, c->W, c->chunk, c->lines_per_cta, c->L, 1.0); }
static void l_rows4(void* p) { Cfg* c = (Cfg*) p; if (g_mode == 1) { rows<4,1><<<c->grid, c->threads>>>(c->out, c->W, c->chunk, c->lines_per_cta, c->L, 1.0); return; } if (g_mode == 2) { rows<4,2><<<c->grid, c->threads>>>(c->out, c->W, c->chunk, c->lines_per_cta, c->L, 1.0); return; } rows<4,0><<<c->grid, c->threads>>>(c->out, c->W, c->chunk, c->lines_per_cta, c->L, 1.0); }
static void l_tma2(void* p) { Cfg* c = (Cfg*) p; size_t sm = 2 * sizeof(double) * c->chunk; cudaFuncSetAttribute(rows_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm); rows_tma<2><<<c->grid, c->threads, sm>>>(c->out, c->W, c->chunk, c->lines_per_cta, c->L, 1.0); }
static void l_tma4(void* p) { Cfg* c = (Cfg*) p; size_t sm = 2 * sizeof(double) * c->chunk; cudaFuncSetAttribute(rows_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm); rows_tma<4><<<c->grid, c->threads, sm>>>(c->out, c->W, c->chunk, c->lines_per_cta, c->L, 1.0); }
static void l_flat(void* p) { Cfg* c = (Cfg*) p; flat<<<c->blocks, c->threads>>>(c->out, c->n, 1.0); }
static void l_flat2(void* p) { Cfg* c = (Cfg*) p; flat2<<<c->blocks, c->threads>>>((double2*) c->out, c->n / 2, 1.0); }

int main()
{
    const int W = 2101; const long L = 11664; const long n = L * W;
    double* out; cudaMalloc(&out, sizeof(double) * L * 2112);
    Cfg c; c.out = out; c.W = W; c.L = L; c.n = n;
    int sms = 148;
    for (int threads : {256, 1024}) for (int bps : {1, 2, 4, 8}) {
        if (threads * bps > 2048) continue;
        c.threads = threads; c.blocks = sms * bps;
        float ms = timeit(l_flat, &c);
        printf("flat  8B stores, %4d thr x %d CTA/SM: %.1f us  %.0f GB/s\n", threads, bps, ms * 1e3, n * 8.0 / ms / 1e6);
        ms = timeit(l_flat2, &c);
        printf("flat 16B stores, %4d thr x %d CTA/SM: %.1f us  %.0f GB/s\n", threads, bps, ms * 1e3, n * 8.0 / ms / 1e6);
    }
    for (int Wt : {2101, 2104, 2112}) for (int mode : {0, 1, 2}) for (int lpt : {2, 4}) for (int occ : {2, 4}) {
        const int W = Wt; c.W = W; g_mode = mode;
        int n_chunks = (W + lpt * 256 - 1) / (lpt * 256);
        int threads = (W + n_chunks * lpt - 1) / (n_chunks * lpt); threads = (threads + 31) / 32 * 32;
        c.threads = threads; c.chunk = lpt * threads; c.lpt = lpt;
        long nby = (long) sms * occ / n_chunks; c.lines_per_cta = (L + nby - 1) / nby; nby = (L + c.lines_per_cta - 1) / c.lines_per_cta;
        c.grid = dim3(n_chunks, (unsigned) nby);
        float ms = timeit(lpt == 2 ? l_rows2 : l_rows4, &c);
        printf("W %d mode %d rows LPT %d, %3d thr, %d CTA/SM (%d x %ld CTAs, %ld lines each): %.1f us  %.0f GB/s\n", W, mode, lpt, threads, occ,
               n_chunks, nby, c.lines_per_cta, ms * 1e3, (double) L * W * 8.0 / ms / 1e6);
    }
    // TMA bulk stores from shared memory (aligned pitches only)
    for (int Wt : {2104, 2112}) for (int lpt : {2, 4}) for (int occ : {1, 2, 4}) {
        const int W = Wt; c.W = W;
        int n_chunks = (W + lpt * 256 - 1) / (lpt * 256);
        int threads = (W + n_chunks * lpt - 1) / (n_chunks * lpt); threads = (threads + 31) / 32 * 32;
        c.threads = threads; c.chunk = lpt * threads; c.lpt = lpt;
        long nby = (long) sms * occ / n_chunks; c.lines_per_cta = (L + nby - 1) / nby; nby = (L + c.lines_per_cta - 1) / c.lines_per_cta;
        c.grid = dim3(n_chunks, (unsigned) nby);
        float ms = timeit(lpt == 2 ? l_tma2 : l_tma4, &c);
        cudaError_t e = cudaDeviceSynchronize();
        printf("W %d TMA bulk rows LPT %d, %3d thr, %d CTA/SM (%d x %ld CTAs, %ld lines each): %.1f us  %.0f GB/s  [%s]\n", W, lpt, threads, occ,
               n_chunks, nby, c.lines_per_cta, ms * 1e3, (double) L * W * 8.0 / ms / 1e6, cudaGetErrorString(e));
    }
    return 0;
}
