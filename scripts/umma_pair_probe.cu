// umma_pair_probe.cu -- the mechanics of a CTA pair (tcgen05 cta_group::2) in isolation, before the trunk kernel relies on
// them: collective tensor-memory allocation, one M = 256 MMA issued by the leader over both CTAs' A rows with each CTA
// holding HALF of B's N rows at the same shared-memory offset, the multicast commit that signals both CTAs' mbarriers, a
// remote mbarrier arrive (peer -> leader), and the per-MMA cost against cta_group::1.
// Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I alphazero-general_b200/csrc -o build/umma_pair_probe scripts/umma_pair_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "azb_tc_ptx.cuh"
using namespace azbtc;

constexpr int N = 96, NH = N / 2, FROWS = 160, PAD = 16, PLANE = FROWS * 16;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void remote_arrive(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
__host__ __device__ constexpr uint32_t idesc2(int n, int m) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void umma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar), "h"(mask)
                 : "memory");
}

// a: [2 CTAs][2 chunks][FROWS][8] bf16, b: [2 chunks][N][8] bf16 (CTA r stages rows r*NH .. of every chunk), d: [256][N]
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k(const __nv_bfloat16 *a, const __nv_bfloat16 *b, float *d, int shift, int iters, long long *cycles)
{
    __shared__ __align__(128) unsigned char fa[2 * PLANE];
    __shared__ __align__(128) unsigned char fb[2 * NH * 16];
    __shared__ unsigned long long bars[2];          // 0: MMA done (multicast commit), 1: peer ready (remote arrive, leader only)
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_rank();
    for (int i = tid; i < 2 * FROWS * 8; i += 128) reinterpret_cast<__nv_bfloat16 *>(fa)[i] = a[(size_t)rank * 2 * FROWS * 8 + i];
    for (int i = tid; i < 2 * NH * 8; i += 128) {
        const int ch = i / (NH * 8), rem = i - ch * NH * 8;
        reinterpret_cast<__nv_bfloat16 *>(fb)[i] = b[(size_t)ch * N * 8 + (size_t)rank * NH * 8 + rem];
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                              // both CTAs staged their operands and initialised their barriers
    tc_fence_after();
    const uint32_t tm = slot;
    if (rank == 1 && tid == 0) remote_arrive(map_to_rank(smem_u32(&bars[1]), 0));     // "my operands are in place"
    if (rank == 0 && warp == 0 && elect_one_sync()) {
        mbar_wait(smem_u32(&bars[1]), 0);
        tc_fence_after();
        const uint64_t ad = umma_desc(smem_u32(fa) + (uint32_t)((PAD + shift) * 16), PLANE, 128);
        const uint64_t bd = umma_desc(smem_u32(fb), NH * 16, 128);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) umma2(tm, ad, bd, idesc2(N, 256), i > 0 ? 1u : 0u);
        commit2(smem_u32(&bars[0]), 3);
        mbar_wait(smem_u32(&bars[0]), 0);
        if (cycles) cycles[blockIdx.x / 2] = clock64() - t0;
    }
    __syncwarp();
    mbar_wait(smem_u32(&bars[0]), 0);                // both CTAs: the multicast commit arrived on MY barrier
    tc_fence_after();
    for (int c = 0; c < N / 16; c++) {
        uint32_t v[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + 16 * c, v);
        tmem_ld_wait();
        if (blockIdx.x < 2)
            for (int j = 0; j < 16; j++) d[((size_t)rank * 128 + tid) * N + 16 * c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(128u) : "memory");
}

int main()
{
    std::vector<__nv_bfloat16> ha(2 * 2 * FROWS * 8), hb(2 * N * 8);
    std::vector<float> fa(ha.size()), fb(hb.size());
    srand(2);
    for (size_t i = 0; i < ha.size(); i++) { fa[i] = (float)(rand() % 7 - 3); ha[i] = __float2bfloat16(fa[i]); }
    for (size_t i = 0; i < hb.size(); i++) { fb[i] = (float)(rand() % 5 - 2); hb[i] = __float2bfloat16(fb[i]); }
    __nv_bfloat16 *da, *db;
    float *dd;
    long long *dc;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dd, 256 * N * 4); cudaMalloc(&dc, 148 * 8);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    std::vector<float> hd(256 * N);
    int bad_total = 0;
    for (int s : {0, 8, 1, -9}) {
        cudaMemset(dd, 0, 256 * N * 4);
        k<<<2, 128>>>(da, db, dd, s, 1, nullptr);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("shift %d: CUDA error %s\n", s, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r = 0; r < 256; r++)
            for (int n = 0; n < N; n++) {
                float ref = 0;
                const int cta = r / 128, rr = r % 128;
                for (int kk = 0; kk < 16; kk++)
                    ref += fa[(size_t)cta * 2 * FROWS * 8 + (kk / 8) * FROWS * 8 + (PAD + s + rr) * 8 + kk % 8] * fb[(kk / 8) * N * 8 + n * 8 + kk % 8];
                if (ref != hd[r * N + n]) { if (bad < 4) printf("  r %d n %d: got %g want %g\n", r, n, hd[r * N + n], ref); bad++; }
            }
        printf("pair MMA M=256 N=%d, A shift %+d rows: %s (%d mismatches)\n", N, s, bad ? "WRONG" : "ok", bad);
        bad_total += bad;
    }
    // cost per MMA on all SMs (74 pairs)
    const int iters = 4096;
    for (int rep = 0; rep < 2; rep++) {
        k<<<148, 128>>>(da, db, dd, 0, iters, dc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("timing: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    }
    std::vector<long long> hc(74);
    cudaMemcpy(hc.data(), dc, 74 * 8, cudaMemcpyDeviceToHost);
    double sum = 0;
    for (auto v : hc) sum += (double)v;
    printf("cta_group::2 M=256 N=%d K=16: %.1f cycles per MMA (cta_group::1 M=128: 56.1; math floor %.1f)\n", N, sum / 74 / iters, N / 2.0);
    printf(bad_total ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad_total != 0;
}
