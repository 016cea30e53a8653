// umma_shift_probe.cu -- does a K-major no-swizzle UMMA operand descriptor accept a start address that is only
// 16-byte aligned (a one-row shift inside an 8-row core matrix)?  The 128-channel trunk kernel forms every 3x3 tap
// (dy, dx) as `start address += 16 (8 dy + dx)`; the 32/64-channel kernel only ever used multiples of 128 bytes.
// Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I alphazero-general_b200/csrc -o /tmp/probe scripts/umma_shift_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "azb_tc_ptx.cuh"
using namespace azbtc;

constexpr int FROWS = 160, PAD = 16, PLANE = FROWS * 16, N = 128;

__global__ void __launch_bounds__(128, 1) k(const __nv_bfloat16 *a, const __nv_bfloat16 *b, float *d, int shift)
{
    __shared__ __align__(128) unsigned char fa[2 * PLANE];
    __shared__ __align__(128) unsigned char fb[2 * N * 16];
    __shared__ unsigned long long bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 2 * PLANE / 2; i += 128) reinterpret_cast<__nv_bfloat16 *>(fa)[i] = a[i];
    for (int i = tid; i < 2 * N * 8; i += 128) reinterpret_cast<__nv_bfloat16 *>(fb)[i] = b[i];
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) tmem_alloc<128>(smem_u32(&slot));
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 0 && elect_one_sync()) {
        const uint64_t ad = umma_desc(smem_u32(fa) + (uint32_t)((PAD + shift) * 16), PLANE, 128);
        const uint64_t bd = umma_desc(smem_u32(fb), N * 16, 128);
        umma_f16(tm, ad, bd, umma_idesc(N, false), 0u);
        umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    for (int c = 0; c < N / 16; c++) {
        uint32_t v[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + 16 * c, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; j++) d[tid * N + 16 * c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tm);
}

int main()
{
    // A: frame [2 chunks][FROWS][8] with small integers, B: [2 chunks][N][8]
    std::vector<__nv_bfloat16> ha(2 * FROWS * 8), hb(2 * N * 8);
    std::vector<float> fa(ha.size()), fb(hb.size());
    srand(1);
    for (size_t i = 0; i < ha.size(); i++) { fa[i] = (float)(rand() % 7 - 3); ha[i] = __float2bfloat16(fa[i]); }
    for (size_t i = 0; i < hb.size(); i++) { fb[i] = (float)(rand() % 5 - 2); hb[i] = __float2bfloat16(fb[i]); }
    __nv_bfloat16 *da, *db;
    float *dd;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dd, 128 * N * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    std::vector<float> hd(128 * N);
    int bad_total = 0;
    const int shifts[] = {0, 8, -8, 1, -1, 7, 9, -7, -9, 3, -5};
    for (int s : shifts) {
        k<<<1, 128>>>(da, db, dd, s);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("shift %d: CUDA error %s\n", s, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r = 0; r < 128; r++)
            for (int n = 0; n < N; n++) {
                float ref = 0;
                for (int kk = 0; kk < 16; kk++)
                    ref += fa[(kk / 8) * FROWS * 8 + (PAD + s + r) * 8 + kk % 8] * fb[(kk / 8) * N * 8 + n * 8 + kk % 8];
                if (ref != hd[r * N + n]) bad++;
            }
        printf("shift %+d rows (%+d bytes): %s (%d mismatches)\n", s, s * 16, bad ? "WRONG" : "ok", bad);
        bad_total += bad;
    }
    printf(bad_total ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad_total != 0;
}
