// umma_shift_timing.cu -- cycles per tcgen05.mma (M = 128, K = 16, bf16, operands in shared memory, no swizzle) when the
// A (or B) operand's start address is 128-byte aligned vs. shifted by one 16-byte row: does a core matrix that straddles
// two 128-byte lines cost extra shared-memory wavefronts?  Decides between the per-tap (shifted start) and the
// shared-A-read + rotation forms of the 3x3 convolution for the 128-channel trunk.
// Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I alphazero-general_b200/csrc -o build/umma_shift_timing scripts/umma_shift_timing.cu
#include <cstdio>
#include <vector>
#include "azb_tc_ptx.cuh"
using namespace azbtc;

constexpr int PLANE = 4608;      // 288 frame rows x 16 B, as the 128-channel frame

__host__ __device__ constexpr uint32_t idesc_n(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

// which: 0 = shift A, 1 = shift B.  The A operand is 128 rows, the B operand n rows; both [2 K chunks][rows][16 B].
__global__ void __launch_bounds__(128, 1) k(int n, int shift_rows, int which, int iters, long long *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 0 && elect_one_sync()) {
        const uint32_t a_s = smem_u32(smem) + 1024, b_s = smem_u32(smem) + 48 * 1024;
        const uint32_t sa = which == 0 ? (uint32_t)(shift_rows * 16) : 0u, sb = which == 1 ? (uint32_t)(shift_rows * 16) : 0u;
        const uint32_t id = idesc_n(n);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            // walk over 8 K steps like a trunk layer (different chunk planes), two accumulators
            const uint32_t ks = (uint32_t)(i & 7);
            const uint64_t ad = umma_desc(a_s + sa + 2 * ks * PLANE * 0 + (ks & 3) * 2048, PLANE, 128);
            const uint64_t bd = umma_desc(b_s + sb + (ks & 3) * 256, (uint32_t)(n < 256 ? n * 16 : PLANE), 128);
            umma_f16(tm + (uint32_t)((i & 1) * 256), ad, bd, id, 1u);
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

int main()
{
    long long *d;
    cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    const int iters = 4096;
    struct { int n, shift, which; const char *what; } cases[] = {
        {128, 0, 0, "N=128 aligned"}, {128, 1, 0, "N=128 A shifted 1 row"}, {128, 9, 0, "N=128 A shifted 9 rows"},
        {128, 8, 0, "N=128 A shifted 8 rows"}, {96, 0, 0, "N=96 aligned"}, {96, 1, 0, "N=96 A shifted 1 row"},
        {256, 0, 1, "N=256 aligned"}, {256, 1, 1, "N=256 B shifted 1 row"}, {256, 1, 0, "N=256 A shifted 1 row"},
        {64, 0, 0, "N=64 aligned"}, {32, 0, 0, "N=32 aligned"}, {32, 1, 0, "N=32 A shifted 1 row"},
    };
    for (auto &c : cases) {
        for (int rep = 0; rep < 2; rep++) {
            k<<<148, 128, 96 * 1024>>>(c.n, c.shift, c.which, iters, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.what, cudaGetErrorString(e)); return 1; }
        }
        std::vector<long long> h(148);
        cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
        double s = 0;
        for (auto v : h) s += (double)v;
        const double cyc = s / 148 / iters;
        printf("%-28s %7.1f cycles / MMA   (math floor %5.1f, operand bytes %5d -> %5.1f wavefronts)\n", c.what, cyc,
               2.0 * 128 * c.n * 16 / 8192, (128 + c.n) * 32, (128 + c.n) * 32 / 128.0);
    }
    return 0;
}
