// ldtm_bench.cu -- tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps and the shape.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define R16(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
#define R16B(v) "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
__global__ void __launch_bounds__(1024, 1) k(int iters, int mode, long long *out, float *sink)
{
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tm = slot;
    const uint32_t taddr = tm + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(32 * ((warp >> 2) % 8));
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        uint32_t v[32];
        if (mode == 0) {          // x16, wait after each
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n" : R16(v) : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            acc += __uint_as_float(v[3]);
        } else if (mode == 1) {   // x32, wait after each
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n" : R16(v), R16B(v) : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            acc += __uint_as_float(v[3]) + __uint_as_float(v[20]);
        } else {                  // two x16 back to back, one wait
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n" : R16(v) : "r"(taddr) : "memory");
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n" : R16B(v) : "r"(taddr + 16) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            acc += __uint_as_float(v[3]) + __uint_as_float(v[20]);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 12345.f) sink[tid] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512u) : "memory");
}
int main()
{
    long long *d, h[148];
    float *sink;
    cudaMalloc(&d, sizeof(h)); cudaMalloc(&sink, 4096);
    const int iters = 2000;
    for (int mode = 0; mode < 3; mode++)
        for (int warps : {1, 4, 8, 16, 32}) {
            for (int rep = 0; rep < 2; rep++) {
                k<<<148, warps * 32>>>(iters, mode, d, sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            const double bytes = (double)warps * iters * 32 * 4 * (mode == 0 ? 16 : 32);
            printf("mode %d warps %2d: %.1f cycles per load-iteration, %.0f B/cycle/SM\n", mode, warps, (double)h[0] / iters, bytes / h[0]);
        }
    return 0;
}
