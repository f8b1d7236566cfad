// FP64 pipe probe for B200: dependent-issue latency and throughput vs warps/SMSP x independent chains per thread.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probes/fp64_probe scripts/probes/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void chains(double* out, long long* cyc, int iters)
{
    double a[C];
#pragma unroll
    for (int i = 0; i < C; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double m = 1.0000001, c = 1e-12;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < C; ++i) a[i] = fma(a[i], m, c);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// mixed: one dependent DFMA chain + K independent integer/ALU ops per DFMA (does the FP64 pipe co-issue with ALU work?)
template <int C>
__global__ void mufu_mix(double* out, long long* cyc, int iters)
{
    double a[C];
#pragma unroll
    for (int i = 0; i < C; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < C; ++i)
            {
                const double d = fma(a[i], a[i], 3.0);
                double r0;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(d));
                const double e = fma(-d, r0, 1.0);
                const double r = fma(r0, e, r0);
                a[i] = a[i] * fma(0.5, r, 0.9);
            }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int C>
void run(const char* name, void (*k)(double*, long long*, int), int opsPerIter, int threads, int ctasPerSm)
{
    double* d; long long* c;
    cudaMalloc(&d, 8); cudaMalloc(&c, 8);
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148 * ctasPerSm, threads>>>(d, c, iters);
    cudaEventRecord(e0);
    k<<<148 * ctasPerSm, threads>>>(d, c, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
    const double ops = (double) 148 * ctasPerSm * threads * (double) iters * opsPerIter;
    printf("%-10s C=%d threads=%4d ctas/SM=%d warps/SMSP=%4.1f : %7.1f cyc/iter/warp  %6.2f cyc per dependent op  %7.2f Tinst/s (%.1f %% of 148*64*1.965G)\n", name, C,
           threads, ctasPerSm, threads * ctasPerSm / 128.0, (double) cy / iters, (double) cy / iters / (opsPerIter / C), ops / ms / 1e9,
           100.0 * ops / (ms * 1e-3) / (148.0 * 64 * 1.965e9));
    cudaFree(d); cudaFree(c);
}

int main()
{
    for (int threads : {32, 128, 256, 512, 1024})
    {
        run<1>("dfma", chains<1>, 8 * 1, threads, 1);
        run<2>("dfma", chains<2>, 8 * 2, threads, 1);
        run<4>("dfma", chains<4>, 8 * 4, threads, 1);
        run<8>("dfma", chains<8>, 8 * 8, threads, 1);
    }
    for (int threads : {128, 256, 512})
    {
        run<1>("satmix", mufu_mix<1>, 4 * 5, threads, 1);
        run<2>("satmix", mufu_mix<2>, 4 * 10, threads, 1);
        run<4>("satmix", mufu_mix<4>, 4 * 20, threads, 1);
    }
    return 0;
}
