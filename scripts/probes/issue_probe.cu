// Does an FP64 instruction hold the warp scheduler's issue port for its two pipe cycles, or can other instruction classes
// issue in between?  Per thread: eight independent DFMA chains, plus M instructions of another class per DFMA (independent
// chains of their own).  If the time of M > 0 equals the time of M = 0 the other class co-issues for free; if it grows like
// (2 + M) / 2 every instruction costs an issue slot and a DFMA costs two.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/probes/issue_probe scripts/probes/issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

// KIND 0 = IMAD, 1 = FFMA, 2 = LDS.64, 3 = SHFL, 4 = LOP3/IADD (ALU), 5 = MUFU
template <int KIND, int MNUM, int MDEN>   // M = MNUM / MDEN other instructions per DFMA
__global__ void __launch_bounds__(256) mix(double* out, int iters, int dummy)
{
    __shared__ double sh[256 * 2];
    sh[threadIdx.x] = threadIdx.x;
    sh[threadIdx.x + 256] = 1.0;
    __syncthreads();
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double m = 1.0000001, c = 1e-12;
    unsigned q[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { q[i] = threadIdx.x * 7 + i + dummy; f[i] = 1.0f + i; }
    double acc = 0.0;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int u = 0; u < 4 * MDEN; ++u)
        {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
#pragma unroll
            for (int j = 0; j < 8 * MNUM / MDEN; ++j)
            {
                const int i = j & 7;
                if (KIND == 0) q[i] = q[i] * 1664525u + (unsigned) dummy;
                if (KIND == 1) f[i] = fmaf(f[i], 1.0000001f, 1e-7f);
                if (KIND == 2) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned) __cvta_generic_to_shared(sh) + ((q[i] & 255u) << 3))); acc += 0; q[i] ^= __double2loint(v) & 0; }
                if (KIND == 3) q[i] = __shfl_xor_sync(0xffffffffu, q[i], 1);
                if (KIND == 4) q[i] = (q[i] ^ (unsigned) dummy) + 0x9e3779b9u;
                if (KIND == 5) f[i] = __frcp_rn(f[i]) ;
            }
        }
    }
    double s = acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + q[i] + f[i];
    if (s == 123.456) out[0] = s;
}

template <int KIND, int MNUM, int MDEN>
float run(const char* name, int ctasPerSm)
{
    double* d;
    cudaMalloc(&d, 8);
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<KIND, MNUM, MDEN><<<148 * ctasPerSm, 256>>>(d, iters, 0);
    cudaEventRecord(e0);
    mix<KIND, MNUM, MDEN><<<148 * ctasPerSm, 256>>>(d, iters, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = (double) 148 * ctasPerSm * 256 * (double) iters * 4 * MDEN * 8;
    printf("%-6s M=%d/%d ctas/SM=%d: %8.3f ms  DFMA rate %5.1f %% of 148*64*1.965G\n", name, MNUM, MDEN, ctasPerSm, ms,
           100.0 * dfma / (ms * 1e-3) / (148.0 * 64 * 1.965e9));
    cudaFree(d);
    return ms;
}

int main()
{
    for (int c = 2; c <= 4; c += 2)
    {
        run<0, 0, 1>("none", c);
        run<0, 1, 2>("IMAD", c); run<0, 1, 1>("IMAD", c); run<0, 2, 1>("IMAD", c);
        run<1, 1, 2>("FFMA", c); run<1, 1, 1>("FFMA", c); run<1, 2, 1>("FFMA", c);
        run<2, 1, 2>("LDS", c);  run<2, 1, 1>("LDS", c);
        run<3, 1, 2>("SHFL", c); run<3, 1, 1>("SHFL", c);
        run<4, 1, 2>("ALU", c);  run<4, 1, 1>("ALU", c);  run<4, 2, 1>("ALU", c);
        run<5, 1, 2>("MUFU", c); run<5, 1, 1>("MUFU", c);
    }
    return 0;
}
