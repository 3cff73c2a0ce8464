// Device microbenchmarks that size the J/K kernel design (run on the B200 box):
//   read BW with 128-bit streaming loads, red.global.add.f64 throughput, DFMA and DMMA issue rates.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void read_kernel(const double2* __restrict__ p, size_t n2, double* out)
{
    double acc = 0.0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        double2 a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(a.x), "=d"(a.y) : "l"(p + i));
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(b.x), "=d"(b.y) : "l"(p + i + stride));
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(c.x), "=d"(c.y) : "l"(p + i + 2 * stride));
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(d.x), "=d"(d.y) : "l"(p + i + 3 * stride));
        acc += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
    }
    for (; i < n2; i += stride) acc += p[i].x + p[i].y;
    if (acc == 1.2345e-300) out[0] = acc;
}

// every lane does `iters` REDs to distinct addresses spread over `span` doubles
__global__ void red_kernel(double* dst, size_t span, int iters)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        size_t a = (t + (size_t)it * 7919u * 32u) % span;
        asm volatile("red.global.add.f64 [%0], %1;" :: "l"(dst + a), "d"(1.0) : "memory");
    }
}

__global__ void dfma_kernel(double* out, int iters)
{
    double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7, x = 1.0000001, y = 0.5;
    for (int it = 0; it < iters; ++it) {
        a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
        a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void dmma_kernel(double* out, int iters)
{
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s SMs %d\n", prop.name, sms);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    // ---- read bandwidth
    size_t bytes = (size_t)16 << 30;
    double2* buf; double* out; CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&out, 1 << 24));
    CK(cudaMemset(buf, 0, bytes));
    for (int tpb : {256, 512}) for (int bps : {2, 4, 8}) {
        int grid = sms * bps;
        read_kernel<<<grid, tpb>>>(buf, bytes / 16, out); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int r = 0; r < 5; ++r) read_kernel<<<grid, tpb>>>(buf, bytes / 16, out);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("read   tpb %d blocks/SM %d : %.1f GB/s\n", tpb, bps, 5.0 * bytes / (ms * 1e-3) / 1e9);
    }
    // ---- RED throughput
    for (size_t span : {(size_t)1 << 18, (size_t)1 << 21}) {
        int iters = 200, grid = sms * 8, tpb = 256;
        red_kernel<<<grid, tpb>>>((double*)buf, span, 10); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        red_kernel<<<grid, tpb>>>((double*)buf, span, iters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        double n = (double)grid * tpb * iters;
        printf("red.f64 span %zu doubles: %.2f G red/s (%.1f per SM per us)\n", span, n / (ms * 1e-3) / 1e9, n / ms / 1e3 / sms);
    }
    // ---- DFMA / DMMA
    {
        int iters = 20000, grid = sms * 4, tpb = 256;
        dfma_kernel<<<grid, tpb>>>(out, 10); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); dfma_kernel<<<grid, tpb>>>(out, iters); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("DFMA: %.2f TFLOP/s\n", 2.0 * 8 * iters * (double)grid * tpb / (ms * 1e-3) / 1e12);
        dmma_kernel<<<grid, tpb>>>(out, 10); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); dmma_kernel<<<grid, tpb>>>(out, iters); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("DMMA m8n8k4: %.2f TFLOP/s\n", 2.0 * 256 * 8 * iters * (double)grid * (tpb / 32) / (ms * 1e-3) / 1e12);
    }
    return 0;
}
