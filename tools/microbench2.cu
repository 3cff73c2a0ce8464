// FP64 issue-rate microbenchmarks behind the J/K kernel's occupancy analysis (run on the B200 box):
//   DFMA throughput as a function of resident warps per SM and independent chains per warp (ILP), dependent-chain
//   latency, and whether DMMA (FP64 tensor) work runs beside DFMA work or on the same pipe.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters)
{
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
    const double x = 1.0000001, y = 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// per iteration: NF independent DFMAs + ND DMMAs
template <int NF, int ND>
__global__ void mix_kernel(double* out, int iters)
{
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x + i;
    double c[4][2] = {};
    const double x = 1.0000001, y = 0.5, ma = threadIdx.x * 1e-3, mb = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NF; ++i) a[i & 15] = fma(a[i & 15], x, y);
#pragma unroll
        for (int i = 0; i < ND; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i & 3][0]), "+d"(c[i & 3][1]) : "d"(ma), "d"(mb));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
static float time_kernel(K kern, int grid, int tpb, double* out, int iters)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<grid, tpb>>>(out, 10); cudaDeviceSynchronize();
    cudaEventRecord(e0); kern<<<grid, tpb>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 1 << 24));
    const int iters = 4000;
    printf("DFMA TFLOP/s by resident warps per SM (one CTA per SM) and independent chains per warp:\n");
    for (int warps : {4, 8, 12, 16, 32}) {
        float m1 = time_kernel(dfma_kernel<1>, sms, warps * 32, out, iters);
        float m2 = time_kernel(dfma_kernel<2>, sms, warps * 32, out, iters);
        float m4 = time_kernel(dfma_kernel<4>, sms, warps * 32, out, iters);
        float m8 = time_kernel(dfma_kernel<8>, sms, warps * 32, out, iters);
        auto tf = [&](float ms, int ilp) { return 2.0 * 8 * ilp * iters * (double)sms * warps * 32 / (ms * 1e-3) / 1e12; };
        printf("  warps/SM %2d : ILP1 %6.2f  ILP2 %6.2f  ILP4 %6.2f  ILP8 %6.2f\n", warps, tf(m1, 1), tf(m2, 2), tf(m4, 4), tf(m8, 8));
    }
    {   // dependent-chain latency: 1 warp per SM, ILP 1 -> cycles per DFMA = clock * time / count
        float ms = time_kernel(dfma_kernel<1>, sms, 32, out, iters);
        int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
        printf("dependent DFMA: %.1f ns each (%.1f cycles at the %d MHz nominal clock)\n", ms * 1e6 / (8.0 * iters), ms * 1e-3 * khz * 1e3 / (8.0 * iters), khz / 1000);
    }
    printf("DFMA + DMMA mixed, 8 warps per SM (TFLOP/s of each kind in the same kernel):\n");
    {
        const int warps = 8;
        float a = time_kernel(mix_kernel<16, 0>, sms, warps * 32, out, iters);
        float b = time_kernel(mix_kernel<0, 2>, sms, warps * 32, out, iters);
        float c = time_kernel(mix_kernel<16, 2>, sms, warps * 32, out, iters);
        double nf = 2.0 * 16 * iters * (double)sms * warps * 32, nd = 2.0 * 256 * 2 * iters * (double)sms * warps;
        printf("  16 DFMA alone      : %.3f ms  (%.2f TF)\n", a, nf / (a * 1e-3) / 1e12);
        printf("  2 DMMA alone       : %.3f ms  (%.2f TF)\n", b, nd / (b * 1e-3) / 1e12);
        printf("  16 DFMA + 2 DMMA   : %.3f ms  (DFMA %.2f TF + DMMA %.2f TF; sum of the separate times %.3f ms)\n", c,
               nf / (c * 1e-3) / 1e12, nd / (c * 1e-3) / 1e12, a + b);
    }
    return 0;
}
