// hop_bench.cu -- latency of one mbarrier hand-off between two warps of a CTA (arrive -> waiter resumes), as a
// function of how many other warps are spinning in mbarrier.try_wait loops on other barriers, and of the wait flavour.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ bool try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns)
{
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity), "r"(ns) : "memory");
    return done != 0;
}
__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory"); }

// delay > 0: warp 0 idles `delay` cycles before every arrive and publishes the arrive time; warp 1 measures wake latency
__device__ long long g_lat[148];
// warps 0 and 1 ping-pong `iters` times over two barriers; warps 2.. spin on a barrier that completes only at the end.
// flavour 0: try_wait loop, 1: test_wait loop, 2: test_wait + nanosleep(32), 3: try_wait with 1000 ns hint
__global__ void k(int iters, int spinners, int flavour, long long *out, int delay)
{
    __shared__ unsigned long long bars[4];
    __shared__ volatile long long t_arr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bars[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bars[2])));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    auto wait = [&](uint32_t bar, uint32_t par) {
        if (flavour == 0) { while (!try_wait(bar, par)) {} }
        else if (flavour == 1) { while (!test_wait(bar, par)) {} }
        else if (flavour == 2) { while (!test_wait(bar, par)) { __nanosleep(32); } }
        else { while (!try_wait_hint(bar, par, 1000)) {} }
    };
    if (warp == 0) {
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            if (delay) { const long long c = clock64(); while (clock64() - c < delay) {} }
            if (lane == 0) { t_arr = clock64(); arrive(smem_u32(&bars[0])); }
            wait(smem_u32(&bars[1]), i & 1);
            __syncwarp();
        }
        const long long t1 = clock64();
        if (lane == 0) { out[blockIdx.x] = t1 - t0; arrive(smem_u32(&bars[2])); }
    } else if (warp == 1) {
        long long lat = 0;
        for (int i = 0; i < iters; i++) {
            wait(smem_u32(&bars[0]), i & 1);
            lat += clock64() - t_arr;
            __syncwarp();
            if (lane == 0) arrive(smem_u32(&bars[1]));
        }
        if (lane == 0) g_lat[blockIdx.x] = lat;
    } else if (warp < 2 + spinners) {
        wait(smem_u32(&bars[2]), 0);
    }
}
int main()
{
    long long *d, h[148];
    cudaMalloc(&d, sizeof(h));
    const int iters = 500;
    for (int delay : {0, 500, 2000, 8000})
    for (int flavour : {0, 3})
        for (int spinners : {0, 26}) {
            for (int rep = 0; rep < 2; rep++) {
                k<<<148, 896>>>(iters, spinners, flavour, d, delay);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            long long lat[148];
            cudaMemcpyFromSymbol(lat, g_lat, sizeof(lat));
            printf("delay %4d flavour %d spinners %2d: %.0f cycles per round trip, wake latency after arrive %.0f\n", delay, flavour, spinners,
                   (double)h[0] / iters, (double)lat[0] / iters);
        }
    return 0;
}
