// umma_bench.cu -- microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, operands in
// shared memory, K-major no-swizzle) as a function of N and of the A start-address alignment.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/umma_bench scripts/umma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t elect_one_sync()
{
    uint32_t pred = 0;
    asm volatile("{\n.reg .pred px;\nelect.sync _|px, 0xFFFFFFFF;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t it = 0; it < (1u << 28); it++) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}

// mode 0: A windows cycle over 9 "taps" with unaligned 16-byte row shifts; mode 1: aligned shifts (multiples of 128 B)
__global__ void __launch_bounds__(128, 1) k(int N, int iters, int mode, int m64, long long *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tm = slot;
    const int M = m64 ? 64 : 128;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a_s = smem_u32(smem) + 4096, b_s = smem_u32(smem) + 128 * 1024;
    const uint32_t PLANE = 14848;
    long long t0 = 0, t1 = 0;
    if (warp == 0 && elect_one_sync()) {
        t0 = clock64();
        for (int it = 0; it < iters / 18; it++) {
            const uint32_t arow = a_s + (uint32_t)(it % 7) * 2048u, dcol = tm + (uint32_t)(it & 1) * (uint32_t)N;
#pragma unroll
            for (int s = 0; s < 18; s++) {
                const int tap = s >> 1, half = s & 1;
                const int shift = mode == 0 ? ((tap / 3 - 1) * 8 + (tap % 3 - 1)) * 16 : mode == 1 ? (tap - 4) * 128 : 0;
                const uint64_t ad = umma_desc(arow + 2 * half * PLANE + shift, PLANE, 128);
                const uint64_t bd = umma_desc(b_s + (uint32_t)(s * 2 * N * 16 % (24 * 1024)), N * 16, 128);
                umma(dcol, ad, bd, idesc, s ? 1u : 0u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(smem_u32(&bar), 0);
        t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512u) : "memory");
}

int main()
{
    long long *d, h[148];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 1800;
    for (int m64 = 0; m64 < 2; m64++)
        for (int mode = 0; mode < 3; mode++)
            for (int N : {16, 32, 64, 96, 128, 256}) {
                if (m64 && N > 128) continue;
                for (int rep = 0; rep < 2; rep++) {
                    k<<<148, 128, 200 * 1024>>>(N, iters, mode, m64, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                long long mx = 0, mn = 1ll << 60;
                for (int i = 0; i < 148; i++) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
                printf("M=%d mode=%d N=%3d: %.1f cycles/MMA (min %.1f)\n", m64 ? 64 : 128, mode, N, (double)mx / iters, (double)mn / iters);
            }
    return 0;
}
