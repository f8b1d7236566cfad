// Upper bound of the EQ inner loops without any inter-warp synchronisation: pass 1 + warp scan + pass 2 on registers,
// 20 bands x reps, constants in shared memory, 2 CTAs/SM (as eq_kernel).  Prints achieved FP64 instruction rate.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++20 -I. -o scripts/probes/eq_inner_probe scripts/probes/eq_inner_probe.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "convopeq_b200/csrc/cpq_eq.cuh"
using namespace cpq;

template <int MODE>   // 0: pass 2 only, 1: pass 1 + scan + matvecs + pass 2, 2: as 1 with the next band's pass 1 fused into pass 2
__global__ void __launch_bounds__(256, 2) probe(const double* cstG, double* out, int reps)
{
    extern __shared__ __align__(16) double sm[];
    for (int i = threadIdx.x; i < CPQ_NUM_BANDS * kEqcStride; i += blockDim.x) sm[i] = cstG[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double x[kEqL];
#pragma unroll
    for (int j = 0; j < kEqL; ++j) x[j] = 0.01 * sin(0.1 * (threadIdx.x * kEqL + j));
    unsigned hiMax = 0;
    const double sat = 0.20000000298023224, alpha = fma(-8.0, sat, 9.0) / 9.0, gamma = 8.0 * sat / 3.0;
    double ic1 = 0.0, ic2 = 0.0;
    EqAcc acc { 0.0, 0.0, 0.0, 0.0 };
    for (int r = 0; r < reps; ++r)
        for (int b = 0; b < CPQ_NUM_BANDS; ++b)
        {
            const double* bc = sm + b * kEqcStride;
            if (MODE >= 1)
            {
                double c1 = acc.c1, c2 = acc.c2, d1 = acc.d1, d2 = acc.d2;
                if (MODE == 1)
#pragma unroll
                for (int j = 0; j < kEqL; j += 2)
                {
                    const double2 wa = reinterpret_cast<const double2*>(bc + kEqcW)[j];
                    const double2 wb = reinterpret_cast<const double2*>(bc + kEqcW)[j + 1];
                    c1 = fma(wa.x, x[j], c1); c2 = fma(wa.y, x[j], c2);
                    d1 = fma(wb.x, x[j + 1], d1); d2 = fma(wb.y, x[j + 1], d2);
                }
                c1 += d1; c2 += d2;
#pragma unroll
                for (int d = 0; d < 5; ++d)
                {
                    const double p1 = __shfl_up_sync(0xffffffffu, c1, 1 << d);
                    const double p2 = __shfl_up_sync(0xffffffffu, c2, 1 << d);
                    if (lane >= (1 << d))
                    {
                        const double2 r0 = reinterpret_cast<const double2*>(bc + kEqcMs + 4 * d)[0];
                        const double2 r1 = reinterpret_cast<const double2*>(bc + kEqcMs + 4 * d)[1];
                        c1 = fma(r0.x, p1, fma(r0.y, p2, c1));
                        c2 = fma(r1.x, p1, fma(r1.y, p2, c2));
                    }
                }
                double e1 = __shfl_up_sync(0xffffffffu, c1, 1), e2 = __shfl_up_sync(0xffffffffu, c2, 1);
                if (lane == 0) { e1 = 0.0; e2 = 0.0; }
                double p1 = ic1 * 1e-3, p2 = ic2 * 1e-3;
                matvec2(bc + kEqcPlo + 4 * (lane & 7), p1, p2, 0.0, 0.0);
                matvec2(bc + kEqcPhi + 4 * (lane >> 3), p1, p2, e1, e2);
                ic1 = p1; ic2 = p2;
            }
            acc = EqAcc { 0.0, 0.0, 0.0, 0.0 };
            if (MODE == 2) eq_pass2<true, 1, true>(x, ic1, ic2, bc, alpha, gamma, hiMax, sm + ((b + 1) % CPQ_NUM_BANDS) * kEqcStride + kEqcW, &acc);
            else eq_pass2<true, 1>(x, ic1, ic2, bc, alpha, gamma, hiMax);
        }
    double s = ic1 + ic2 + hiMax;
#pragma unroll
    for (int j = 0; j < kEqL; ++j) s += x[j];
    if (s == 123.456) out[0] = s;
}

int main()
{
    std::vector<double> c(CPQ_NUM_BANDS * kEqcStride, 0.0);
    for (int b = 0; b < CPQ_NUM_BANDS; ++b)
    {
        double* bc = c.data() + b * kEqcStride;
        const double g = std::tan(3.14159265358979 * (50.0 * std::pow(1.4, b)) / 48000.0), k = 1.0 / 1.5;
        const double a1 = 1.0 / (1.0 + g * (g + k)), a2 = g * a1, a3 = g * a2;
        bc[0] = a1; bc[1] = a2; bc[2] = a3; bc[3] = 1.0; bc[4] = 0.1; bc[5] = 0.0; bc[7] = 1.0; bc[8] = g; bc[9] = 2 * g;
        for (int i = kEqcW; i < kEqcStride; ++i) bc[i] = 1e-3 * std::sin(i * 0.37 + b);
    }
    double *dc, *dout;
    cudaMalloc(&dc, c.size() * 8); cudaMalloc(&dout, 8);
    cudaMemcpy(dc, c.data(), c.size() * 8, cudaMemcpyHostToDevice);
    const size_t smem = c.size() * 8;
    cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 40;
    for (int mode = 0; mode < 3; ++mode)
        for (int threads : {128, 224, 256})
            for (int ctas : {1, 2})
            {
                auto launch = [&] { if (mode == 0) probe<0><<<148 * ctas, threads, smem>>>(dc, dout, reps); else if (mode == 2) probe<2><<<148 * ctas, threads, smem>>>(dc, dout, reps); else probe<1><<<148 * ctas, threads, smem>>>(dc, dout, reps); };
                launch();
                cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                const double fp64PerSample = mode == 0 ? 12.0 : 12.0 + 2.0 + (20.0 + 8.0) / kEqL;
                const double samples = 148.0 * ctas * threads * kEqL * CPQ_NUM_BANDS * reps;
                const double rate = samples * fp64PerSample / (ms * 1e-3);
                printf("mode %d threads %3d ctas/SM %d (%4.1f warps/SMSP): %.3f ms  %.2f T FP64-inst/s = %.1f %% of pipe peak; %.2f G band-samples/s -> 20-band %.2f G ch-samples/s\n",
                       mode, threads, ctas, threads * ctas / 128.0, ms, rate / 1e12, 100.0 * rate / (148.0 * 64 * 1.965e9), samples / (ms * 1e-3) / 1e9,
                       samples / 20 / (ms * 1e-3) / 1e9);
            }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
