// Microbenchmark: shared-memory cost of LDGSTS.128 (cp.async 16 B/lane, 512 B/warp) vs source alignment.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(256) k(const char* __restrict__ src, size_t span, int iters, int misalign, double* out, size_t rstride, int extra_lds)
{
    extern __shared__ __align__(128) unsigned char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm + warp * 4096) + lane * 16;
    size_t gw = (size_t)blockIdx.x * 8 + warp;
    size_t pos = (gw * 7919u * 512u) % span;
    double acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const char* p = src + ((pos + (size_t)r * rstride) % span) + misalign + lane * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + r * 512), "l"(p) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        acc += *(double*)(sm + warp * 4096 + lane * 16);
        for (int e = 0; e < extra_lds; ++e) { double2 v = *(double2*)(sm + warp * 4096 + ((e * 512 + lane * 16) & 4095)); acc += v.x + v.y; }
        pos = (pos + 32768) % span;
    }
    if (acc == 1.234e-300) out[0] = acc;
}

int main(int argc, char** argv)
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    size_t span = (argc > 1 ? (size_t)atol(argv[1]) : 32) << 20;  // MB; 32 = L2 resident, 8192 = DRAM
    char* buf; double* out; CK(cudaMalloc(&buf, span + 8192)); CK(cudaMalloc(&out, 64)); CK(cudaMemset(buf, 0, span + 8192));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    for (int cfg = 0; cfg < 4; ++cfg) { int mis = 16; size_t rstride = (cfg & 1) ? ((size_t)1 << 20) + 4096 : 4096; int extra = (cfg & 2) ? 16 : 0;
        int iters = 2000, grid = prop.multiProcessorCount * 4;
        k<<<grid, 256, 32768>>>(buf, span, 10, mis, out, rstride, extra); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0)); k<<<grid, 256, 32768>>>(buf, span, iters, mis, out, rstride, extra); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double bytes = (double)grid * 8 * iters * 8 * 512;
        printf("span %zu MB rstride %zu extra_lds %d: %.1f GB/s\n", span >> 20, rstride, extra, bytes / (ms * 1e-3) / 1e9);
    }
    return 0;
}
