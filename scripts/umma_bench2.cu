// umma_bench2.cu -- what does the single issuing thread pay per tile?  6 x tcgen05.mma (M=128, N=96, K=16)
// + commit per tile, with optional already-satisfied mbarrier waits and fences in between.
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
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done;
}
// mode bit0: two satisfied try_waits per tile; bit1: tcgen05.fence::after per tile; bit2: wait for the tile's own commit
// (fully serial); bit3: test_wait instead of try_wait
__global__ void __launch_bounds__(896, 1) k(int tiles, int mode, int noise, long long *out, float *sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bars[8];
    __shared__ volatile int stop;
    if (threadIdx.x == 0) stop = 0;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    for (int i = tid; i < 160 * 1024 / 16; i += 896) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int i = 0; i < 8; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bars[i])));
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
    const int N = 96;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_s = smem_u32(smem) + 4096, b_s = smem_u32(smem) + 128 * 1024;
    const uint32_t PLANE = 14848;
    if (tid == 32) {     // complete phase 0 of bars 4,5 so that waits on parity 0 are satisfied immediately
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&bars[4])) : "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&bars[5])) : "memory");
    }
    __syncthreads();
    __shared__ volatile unsigned turn;
    if (tid == 0) turn = 0;
    __syncthreads();
    if ((mode & 16) && warp < 3) {
        if (elect_one_sync()) {
            const long long t0 = clock64();
            long long lat = 0;
            for (int g = warp; g < tiles; g += 3) {
                while (turn != (unsigned)g) {}
                const uint32_t arow = a_s + (uint32_t)(g % 7) * 2048u, dcol = tm + (uint32_t)(g % 3) * 96u;
#pragma unroll
                for (int s = 0; s < 6; s++) {
                    const int dy = s / 2 - 1, half = s & 1;
                    const uint64_t ad = umma_desc(arow + 2 * half * PLANE + dy * 128, PLANE, 128);
                    const uint64_t bd = umma_desc(b_s + (uint32_t)(s * 2 * 1536), 1536, 128);
                    umma(dcol, ad, bd, idesc, s ? 1u : 0u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bars[warp])) : "memory");
                __threadfence_block();
                turn = g + 1;
                if (mode & 32) {   // also wait for my tile to complete (ring of 3 with instantaneous epilogue)
                    const long long c0 = clock64();
                    mbar_wait(smem_u32(&bars[warp]), (uint32_t)((g / 3) & 1));
                    lat += clock64() - c0;
                }
            }
            if (warp == 0) {
                const long long t1 = clock64();
                out[2 * blockIdx.x] = t1 - t0;
                out[2 * blockIdx.x + 1] = lat * 3;
                stop = 1;
            }
        }
    } else
    if (!(mode & 16) && warp == 0 && elect_one_sync()) {
        const long long t0 = clock64();
        uint32_t ph = 0;
        for (int g = 0; g < tiles; g++) {
            if (mode & 1) {
                if (mode & 8) { while (!mbar_test(smem_u32(&bars[4]), 0)) {} while (!mbar_test(smem_u32(&bars[5]), 0)) {} }
                else { mbar_wait(smem_u32(&bars[4]), 0); mbar_wait(smem_u32(&bars[5]), 0); }
            }
            if (mode & 2) asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t arow = a_s + (uint32_t)(g % 7) * 2048u, dcol = tm + (uint32_t)(g % 3) * 96u;
#pragma unroll
            for (int s = 0; s < 6; s++) {
                const int dy = s / 2 - 1, half = s & 1;
                const uint64_t ad = umma_desc(arow + 2 * half * PLANE + dy * 128, PLANE, 128);
                const uint64_t bd = umma_desc(b_s + (uint32_t)(s * 2 * 1536), 1536, 128);
                umma(dcol, ad, bd, idesc, s ? 1u : 0u);
            }
            const uint32_t bar = smem_u32(&bars[(mode & 4) ? 0 : (g % 3)]);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
            if (mode & 4) { mbar_wait(bar, ph); ph ^= 1; }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bars[6])) : "memory");
        const long long t1 = clock64();
        mbar_wait(smem_u32(&bars[6]), 0);
        const long long t2 = clock64();
        out[2 * blockIdx.x] = t1 - t0;
        out[2 * blockIdx.x + 1] = t2 - t0;
        stop = 1;
    } else if (warp >= 4 && noise) {
        // noise warps: 1 = TMEM loads, 2 = shared-memory stores + loads, 3 = shuffles, 4 = TMEM stores
        const int lane = tid & 31;
        const uint32_t taddr = tm + ((uint32_t)((warp & 3) * 32) << 16) + 300u + (uint32_t)(16 * ((warp >> 2) % 3));
        float acc = 0.f;
        uint4 *sm4 = reinterpret_cast<uint4 *>(smem + 64 * 1024 + (warp * 32 + lane) * 16);
        while (!stop) {
            if (noise == 1) {
                uint32_t v[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                               "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                acc += __uint_as_float(v[3]);
            } else if (noise == 2) {
                uint4 x = *sm4; x.x += 1; *sm4 = x; acc += x.y;
            } else if (noise == 3) {
                for (int i = 0; i < 8; i++) acc += __shfl_sync(0xffffffffu, acc, (lane + 1) & 31);
            } else if (noise == 4) {
                uint32_t v[16];
                for (int i = 0; i < 16; i++) v[i] = lane + i;
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};\n" ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                               "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
            }
        }
        if (acc == 12345.f) sink[tid] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512u) : "memory");
}
int main()
{
    long long *d, h[296];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int tiles = 210;
    float *sink; cudaMalloc(&sink, 4096 * 4);
    for (int noise : {0, 1})
    for (int mode : {0, 16, 48}) {
        for (int rep = 0; rep < 2; rep++) {
            k<<<148, 896, 200 * 1024>>>(tiles, mode, noise, d, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("noise %d mode %2d: issue loop %.0f cycles/tile, until all complete %.0f cycles/tile\n", noise, mode, (double)h[0] / tiles, (double)h[1] / tiles);
    }
    return 0;
}
