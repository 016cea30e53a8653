// azb_resnet_tc.cu -- fused leaf evaluation on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM): the reference's pre-activation ResNet
// (alphazero/NNetArchitecture.py:69-120) for a tile of 16 Connect4 boards per
// CTA, one kernel launch per batch.  Same inference-time algebra as
// azb_resnet.cu (azb200/fused_nn.py folds the batch norms and the affine heads).
//
// Implicit GEMM.  A board is stored as a frame of (H+1) x 8 positions: one zero
// column shared by consecutive board rows and one zero row shared by consecutive
// boards are all the padding a 3x3 convolution needs, and a tap (dy, dx) becomes a
// constant shift of dy*8 + dx frame rows.  16 boards = 896 frame rows = 7 M-tiles
// of 128 rows; the GEMM computes all frame rows and the epilogue forces the
// padding rows back to zero.
//
// Shared-memory layout of an activation frame: [8-channel chunk][frame row][16 B]
// -- exactly the canonical K-major, no-swizzle UMMA operand with the 8 rows of a
// core matrix contiguous (SBO = 128 B) and the two 16-byte K chunks of one MMA a
// plane apart (LBO = plane size).  Shifting the operand by a tap is then just
// `start address += shift * 16 B`, so the nine taps need no im2col and no copies,
// and the epilogue's stores (one 16-byte chunk per thread, consecutive threads =
// consecutive rows) are bank-conflict free.  The stem (<= 8 input channels) packs
// two taps into one K=16 step by pointing LBO at the second tap's shift.
//
// Tensor memory: the fp32 residual stream x lives in TMEM for the whole network
// (7 tiles x 32 columns); conv2 of every block accumulates straight onto it (the
// residual add is the MMA's accumulate input), conv1 uses a second set of 7 x 32
// columns.  One elected thread issues the MMAs of a layer tile after tile and
// commits each tile to its own mbarrier; eight epilogue warps (two per TMEM lane
// quadrant) drain finished tiles (tcgen05.ld -> bias / BN / ReLU -> bf16 ->
// shared memory) while the tensor core works on the following tiles.  Weights of
// the next layer are streamed by one bulk copy (UBLKCP) during the current one.
// The affine heads are one small mma.sync over the final activation, softmax fp32.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/azb200_nn.h"

namespace {
namespace tc {

constexpr int NB = 16;                 // boards per CTA
constexpr int CH = 32;                 // trunk channels
constexpr int BH = 6, BW = 7;          // board
constexpr int FW = 8;                  // frame row pitch: BW + one shared zero column
constexpr int FB = (BH + 1) * FW;      // 56 frame rows per board: + one shared zero row
constexpr int ROWS = NB * FB;          // 896
constexpr int TILES = ROWS / 128;      // 7
static_assert(ROWS % 128 == 0, "frame rows of a CTA must fill whole M=128 tiles");
constexpr int PADR = 16;               // zero rows in front of / behind the boards (|tap shift| <= 9)
constexpr int FROWS = ROWS + 2 * PADR; // 928
constexpr int PLANE = FROWS * 16;      // bytes of one 8-channel chunk plane (14848)
constexpr int FRAME = 4 * PLANE;       // 59392
constexpr int KCH = 36;                // 16-byte K chunks per layer: 9 taps x 4
constexpr int WL_BYTES = KCH * CH * 16;    // 18432: [k chunk][cout][8 cin] bf16
constexpr int NHEAD = 16;
constexpr int KH = FB * CH;            // head K: frame row * 32 + channel (1792)
constexpr int RS_H = KH + 8;           // bf16 row stride of the head matrix
constexpr int MAXD = 6, MAXL = 1 + 2 * MAXD;
constexpr int PRM_FLOATS = MAXL * CH + 2 * MAXD * CH;
constexpr int WARPS = 9;               // warp 0: control / MMA issue; warps 1-8: epilogue
constexpr int THREADS = WARPS * 32;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_X = 0, COL_T = 256;   // x tiles at 32*t, conv1 accumulators at 256 + 32*t

constexpr size_t SMEM_BYTES = 2 * (size_t)FRAME + 2 * (size_t)WL_BYTES + (size_t)NHEAD * RS_H * 2 +
                              (size_t)PRM_FLOATS * 4 + 16 * 8 + 16 + 128;

// ---- PTX wrappers --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// bounded wait: a barrier that never completes is a programming error -- trap instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t it = 0; it < (1u << 26); it++) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (done) return;
    }
    __trap();
}
// one lane of a converged warp; inside the guarded region the compiler keeps addresses and
// descriptors in uniform registers (a plain `lane == 0` test makes it wrap every tcgen05.mma in a
// divergence "waterfall" loop, measured at +20 cycles per MMA)
__device__ __forceinline__ uint32_t elect_one_sync()
{
    uint32_t pred = 0;
    asm volatile("{\n.reg .pred px;\nelect.sync _|px, 0xFFFFFFFF;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// K-major, no-swizzle shared-memory operand descriptor (cute::UMMA::SmemDescriptor):
// start address [0,14) and the two strides [16,30) / [32,46) in 16-byte units, version 1 at [46,48)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D = f32 [4,6), A = B = bf16 [7,10) [10,13), both K-major, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CH >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
                 "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
#define AZB_R32(v)                                                                                                     \
    "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),        \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),         \
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),        \
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define AZB_I32(v)                                                                                                     \
    "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),      \
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),    \
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),    \
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
// one accumulator row (this thread's TMEM lane), 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
                 "%29,%30,%31}, [%32];\n"
                 : AZB_R32(v)
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
                 "%29,%30,%31};\n" ::AZB_I32(v),
                 "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}

enum { EPI_STEM = 0, EPI_CONV1 = 1, EPI_CONV2 = 2 };

// tap t = (dy+1)*3 + (dx+1) -> frame-row shift dy*8 + dx
__device__ __forceinline__ int tap_shift(int t) { return (t / 3 - 1) * FW + (t % 3 - 1); }

// All MMAs of one layer, issued by one thread: tile after tile, each tile committed to its own mbarrier.
//   stem : 5 K-steps, each covering two taps of the <= 8 input channels (chunk plane 0 of `in`)
//   trunk: 18 K-steps = 9 taps x two 16-channel halves
__device__ __forceinline__ void issue_layer(uint32_t in_s, uint32_t w_s, uint32_t tmem_d, bool stem, bool accumulate,
                                            uint32_t bar_tile0)
{
#pragma unroll 1
    for (int t = 0; t < TILES; t++) {
        const uint32_t row0 = in_s + (uint32_t)((PADR + 128 * t) * 16);
        const uint32_t d = tmem_d + (uint32_t)(32 * t);
        if (stem) {
#pragma unroll
            for (int s = 0; s < 5; s++) {
                const int sh0 = tap_shift(2 * s), sh1 = s < 4 ? tap_shift(2 * s + 1) : tap_shift(8) + 1;
                const uint64_t ad = umma_desc(row0 + (uint32_t)(sh0 * 16), (uint32_t)((sh1 - sh0) * 16), 128u);
                const uint64_t bd = umma_desc(w_s + (uint32_t)(2 * s * CH * 16), (uint32_t)(CH * 16), 128u);
                umma_f16(d, ad, bd, s > 0 ? 1u : 0u);
            }
        } else {
#pragma unroll
            for (int s = 0; s < 18; s++) {
                const int tap = s >> 1, half = s & 1;
                const uint64_t ad = umma_desc(row0 + (uint32_t)(2 * half * PLANE + tap_shift(tap) * 16), (uint32_t)PLANE, 128u);
                const uint64_t bd = umma_desc(w_s + (uint32_t)(2 * s * CH * 16), (uint32_t)(CH * 16), 128u);
                umma_f16(d, ad, bd, (s > 0 || accumulate) ? 1u : 0u);
            }
        }
        umma_commit(bar_tile0 + 8u * (uint32_t)t);
    }
}

// Epilogue of one layer for the tiles of this warp's group (warps 1-4: tiles 0,2,4,6; warps 5-8: 1,3,5).
// Each thread owns one frame row of the tile = one TMEM lane (lane quadrant = warp % 4).
template <int EPI>
__device__ __forceinline__ void epilogue(uint32_t tmem_base, uint32_t col0, unsigned char *outf, const float *bias,
                                         const float *nsc, const float *nsh, uint32_t bar_tile0, uint32_t parity, int warp,
                                         int lane, float *dump)
{
    const int q = warp & 3, grp = (warp - 1) >> 2;
#pragma unroll 1
    for (int t = grp; t < TILES; t += 2) {
        mbar_wait(bar_tile0 + 8u * (uint32_t)t, parity);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + col0 + (uint32_t)(32 * t);
        uint32_t v[32];
        tmem_ld32(taddr, v);
        const int fr = t * 128 + q * 32 + lane;
        const bool live = ((fr & 7) != 7) && (((fr >> 3) % (BH + 1)) != BH);
        float r[32];
#pragma unroll
        for (int c = 0; c < 32; c++) r[c] = __uint_as_float(v[c]);
        if (EPI == EPI_STEM) {                                  // x = relu(conv + bias) goes back to TMEM
#pragma unroll
            for (int c = 0; c < 32; c++) {
                r[c] = fmaxf(r[c] + bias[c], 0.0f);
                v[c] = __float_as_uint(r[c]);
            }
            tmem_st32(taddr, v);
        } else if (EPI == EPI_CONV1) {                          // b = relu(conv1 + bias)   (BN2 folded)
#pragma unroll
            for (int c = 0; c < 32; c++) r[c] = fmaxf(r[c] + bias[c], 0.0f);
        }
        if (EPI != EPI_CONV1 && nsc != nullptr) {               // a = relu(bn1(x)) of the block that reads x next
#pragma unroll
            for (int c = 0; c < 32; c++) r[c] = fmaxf(fmaf(r[c], nsc[c], nsh[c]), 0.0f);
        }
        if (dump != nullptr) {
#pragma unroll
            for (int c = 0; c < 32; c++) dump[(size_t)fr * 32 + c] = live ? r[c] : 0.0f;
        }
        unsigned char *row = outf + (size_t)(PADR + fr) * 16;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint4 o;
            o.x = live ? pack_bf16(r[8 * c + 0], r[8 * c + 1]) : 0u;
            o.y = live ? pack_bf16(r[8 * c + 2], r[8 * c + 3]) : 0u;
            o.z = live ? pack_bf16(r[8 * c + 4], r[8 * c + 5]) : 0u;
            o.w = live ? pack_bf16(r[8 * c + 6], r[8 * c + 7]) : 0u;
            *reinterpret_cast<uint4 *>(row + (size_t)c * PLANE) = o;
        }
    }
    fence_proxy_async();      // the next layer's MMAs read these rows through the async proxy
    tc_fence_before();
}

template <int NOUT>
__global__ void __launch_bounds__(THREADS, 1)
k_resnet_tc(const float *__restrict__ obs, float *__restrict__ policy, float *__restrict__ value, int B, int in_ch, int depth,
            const unsigned char *__restrict__ wconv, const float *__restrict__ cbias, const float *__restrict__ bn_scale,
            const float *__restrict__ bn_shift, const __nv_bfloat16 *__restrict__ whead, const float *__restrict__ bhead,
            float *__restrict__ dump, int dump_layer)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *fa = smem;                                            // activation frame a
    unsigned char *fb = fa + FRAME;                                      // activation frame b (first: the observation)
    unsigned char *w0 = fb + FRAME;                                      // weights of the even / odd layers
    unsigned char *w1 = w0 + WL_BYTES;
    __nv_bfloat16 *wh = reinterpret_cast<__nv_bfloat16 *>(w1 + WL_BYTES);   // [NHEAD][RS_H]
    float *prm = reinterpret_cast<float *>(wh + NHEAD * RS_H);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(prm + PRM_FLOATS);   // tile[7], w[2], head
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 16);
    float *red = reinterpret_cast<float *>(w0);                          // heads: [WARPS][NB][NHEAD] (after the trunk)
    float *fin = reinterpret_cast<float *>(w1);                          // heads: [NB][NHEAD]

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int board0 = blockIdx.x * NB;
    const int layers = 1 + 2 * depth;
    const uint32_t bar_tile0 = smem_u32(bars), bar_w0 = smem_u32(bars + TILES), bar_h = smem_u32(bars + TILES + 2);
    const uint32_t fa_s = smem_u32(fa), fb_s = smem_u32(fb), w_s[2] = {smem_u32(w0), smem_u32(w1)};

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < TILES + 3; i++) mbar_init(bar_tile0 + 8u * (uint32_t)i, 1);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    // zero both frames (padding rows / columns must read as zero), stage the per-channel parameters
    {
        uint4 *z = reinterpret_cast<uint4 *>(fa);
        for (int i = tid; i < 2 * FRAME / 16; i += THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < layers * CH; i += THREADS) prm[i] = cbias[i];
        for (int i = tid; i < depth * CH; i += THREADS) {
            prm[MAXL * CH + i] = bn_scale[i];
            prm[MAXL * CH + MAXD * CH + i] = bn_shift[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const float *s_bias = prm, *s_sc = prm + MAXL * CH, *s_sh = prm + MAXL * CH + MAXD * CH;

    if (tid == 0) {       // weights of layer 0 and the head matrix
        mbar_expect_tx(bar_w0, WL_BYTES);
        bulk_g2s(w_s[0], wconv, WL_BYTES, bar_w0);
        mbar_expect_tx(bar_h, (uint32_t)(NHEAD * RS_H * 2));
        bulk_g2s(smem_u32(wh), whead, (uint32_t)(NHEAD * RS_H * 2), bar_h);
    }
    // observation -> chunk plane 0 of frame b (channels >= in_ch stay zero)
    for (int i = tid; i < NB * BH * BW; i += THREADS) {
        const int bl = i / (BH * BW), pos = i - bl * (BH * BW), y = pos / BW, xx = pos - y * BW;
        const int gb = board0 + bl;
        float c[8];
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] = (gb < B && k < in_ch) ? obs[((size_t)gb * in_ch + k) * (BH * BW) + pos] : 0.0f;
        uint4 o;
        o.x = pack_bf16(c[0], c[1]); o.y = pack_bf16(c[2], c[3]); o.z = pack_bf16(c[4], c[5]); o.w = pack_bf16(c[6], c[7]);
        *reinterpret_cast<uint4 *>(fb + (size_t)(PADR + bl * FB + y * FW + xx) * 16) = o;
    }
    fence_proxy_async();
    __syncthreads();

    // ---- trunk: one pass per layer ------------------------------------------------
#pragma unroll 1
    for (int l = 0; l < layers; l++) {
        const bool stem = l == 0, is_c1 = (l & 1) == 1;
        const uint32_t parity = (uint32_t)(l & 1);
        if (warp == 0) {
            if (elect_one_sync()) {
                if (l + 1 < layers) {      // stream the next layer's weights into the buffer layer l-1 has released
                    const uint32_t bw = bar_w0 + 8u * (uint32_t)((l + 1) & 1);
                    mbar_expect_tx(bw, WL_BYTES);
                    bulk_g2s(w_s[(l + 1) & 1], wconv + (size_t)(l + 1) * WL_BYTES, WL_BYTES, bw);
                }
                mbar_wait(bar_w0 + 8u * (uint32_t)(l & 1), (uint32_t)((l >> 1) & 1));
                tc_fence_after();
                // stem: obs (frame b) -> x;  conv1: a -> T;  conv2: b -> x (+=)
                issue_layer((stem || !is_c1) ? fb_s : fa_s, w_s[l & 1], tmem_base + (is_c1 ? COL_T : COL_X), stem,
                            !stem && !is_c1, bar_tile0);
            }
            __syncwarp();
        } else {
            float *dmp = (dump != nullptr && l == dump_layer) ? dump + (size_t)blockIdx.x * ROWS * 32 : nullptr;
            if (stem) {
                epilogue<EPI_STEM>(tmem_base, COL_X, fa, s_bias, depth > 0 ? s_sc : nullptr, s_sh, bar_tile0, parity, warp, lane, dmp);
            } else if (is_c1) {
                epilogue<EPI_CONV1>(tmem_base, COL_T, fb, s_bias + l * CH, nullptr, nullptr, bar_tile0, parity, warp, lane, dmp);
            } else {
                const int nb = l >> 1;     // the block that consumes x next
                const bool last = l + 1 == layers;
                epilogue<EPI_CONV2>(tmem_base, COL_X, fa, nullptr, last ? nullptr : s_sc + nb * CH, s_sh + nb * CH, bar_tile0,
                                    parity, warp, lane, dmp);
            }
        }
        __syncthreads();
        tc_fence_after();
    }

    // ---- heads: logits[board][j] = sum_k wh[j][k] * a[board][k] + bhead[j], k = frame row * 32 + channel ----------
    mbar_wait(bar_h, 0u);
    {
        const uint32_t h_s = smem_u32(wh);
        float acc[2][4];
#pragma unroll
        for (int i = 0; i < 2; i++) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f; }
        const int brd = lane & 15, khalf = lane >> 4;
        const uint32_t a_lane = fa_s + (uint32_t)((PADR + brd * FB) * 16 + khalf * PLANE);
        const uint32_t h_lane = h_s + (uint32_t)(((((lane >> 4) << 3) + (lane & 7)) * RS_H + ((lane >> 3) & 1) * 8) * 2);
        for (int ks = warp; ks < KH / 16; ks += WARPS) {
            const int fr = ks >> 1, half = ks & 1;
            uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
            ldmatrix_x4(a0, a1, a2, a3, a_lane + (uint32_t)(fr * 16 + 2 * half * PLANE));
            ldmatrix_x4(b0, b1, b2, b3, h_lane + (uint32_t)(ks * 32));
            mma_bf16(acc[0], a0, a1, a2, a3, b0, b1);
            mma_bf16(acc[1], a0, a1, a2, a3, b2, b3);
        }
        const int g = lane >> 2, t = lane & 3;                     // c0,c1: board g; c2,c3: board g+8; cols 2t, 2t+1
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            red[(warp * NB + g) * NHEAD + nt * 8 + 2 * t] = acc[nt][0];
            red[(warp * NB + g) * NHEAD + nt * 8 + 2 * t + 1] = acc[nt][1];
            red[(warp * NB + g + 8) * NHEAD + nt * 8 + 2 * t] = acc[nt][2];
            red[(warp * NB + g + 8) * NHEAD + nt * 8 + 2 * t + 1] = acc[nt][3];
        }
        __syncthreads();
        if (tid < NB * NHEAD) {
            float v = 0.0f;
            for (int wq = 0; wq < WARPS; wq++) v += red[wq * NB * NHEAD + tid];
            fin[tid] = v + ((tid % NHEAD) < NOUT ? bhead[tid % NHEAD] : 0.0f);
        }
        __syncthreads();
        if (tid < NB && board0 + tid < B) {
            const int gb = board0 + tid;
            constexpr int A = NOUT - 3;
            float lg[NOUT];
#pragma unroll
            for (int j = 0; j < NOUT; j++) lg[j] = fin[tid * NHEAD + j];
            float mp = lg[0], mv = lg[A];
#pragma unroll
            for (int j = 1; j < A; j++) mp = fmaxf(mp, lg[j]);
#pragma unroll
            for (int j = A + 1; j < NOUT; j++) mv = fmaxf(mv, lg[j]);
            float sp = 0.0f, sv = 0.0f;
#pragma unroll
            for (int j = 0; j < NOUT; j++) {
                lg[j] = expf(lg[j] - (j < A ? mp : mv));
                if (j < A) sp += lg[j]; else sv += lg[j];
            }
#pragma unroll
            for (int j = 0; j < A; j++) policy[(size_t)gb * A + j] = lg[j] / sp;
#pragma unroll
            for (int j = A; j < NOUT; j++) value[(size_t)gb * 3 + (j - A)] = lg[j] / sv;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace tc
}  // namespace

extern "C" int azb_nn_tc_layer_bytes(void) { return tc::WL_BYTES; }
extern "C" int azb_nn_tc_head_row_stride(void) { return tc::RS_H; }
extern "C" int azb_nn_tc_boards_per_cta(void) { return tc::NB; }
extern "C" int azb_nn_tc_frame_rows_per_board(void) { return tc::FB; }

static int tc_launch(const azb_nn_weights *w, const float *obs, float *policy, float *value, int32_t batch, void *stream,
                     float *dump, int dump_layer)
{
    if (!w || !obs || !policy || !value || batch <= 0) return -7;
    if (w->channels != tc::CH || w->board_h != tc::BH || w->board_w != tc::BW || w->action_size != 7 || w->in_channels > 8 ||
        w->in_channels < 1 || w->depth < 0 || w->depth > tc::MAXD)
        return -1;
    cudaStream_t s = (cudaStream_t)stream;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(tc::k_resnet_tc<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES) != cudaSuccess)
            return -2;
        configured = true;
    }
    const int grid = (batch + tc::NB - 1) / tc::NB;
    tc::k_resnet_tc<10><<<grid, tc::THREADS, tc::SMEM_BYTES, s>>>(
        obs, policy, value, batch, w->in_channels, w->depth, reinterpret_cast<const unsigned char *>(w->wconv), w->cbias,
        w->bn_scale, w->bn_shift, reinterpret_cast<const __nv_bfloat16 *>(w->whead), w->bhead, dump, dump_layer);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

extern "C" int azb_nn_forward_tc(const azb_nn_weights *w, const float *obs, float *policy, float *value, int32_t batch,
                                 void *stream)
{
    return tc_launch(w, obs, policy, value, batch, stream, nullptr, -1);
}

extern "C" int azb_nn_forward_tc_debug(const azb_nn_weights *w, const float *obs, float *policy, float *value, int32_t batch,
                                       void *stream, float *dump, int32_t dump_layer)
{
    return tc_launch(w, obs, policy, value, batch, stream, dump, dump_layer);
}
