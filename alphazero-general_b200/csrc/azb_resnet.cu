// azb_resnet.cu -- fused leaf evaluation for small boards: the reference's
// pre-activation ResNet (alphazero/NNetArchitecture.py:69-120) for one tile of
// boards entirely in shared memory, one kernel launch per batch.
//
// Inference-time algebra (done on the host by azb200/fused_nn.py):
//   * stem   conv3x3 -> BN -> ReLU            : BN folded into the conv (weights, bias)
//   * block  BN1 -> ReLU -> conv1 -> BN2 -> ReLU -> conv2 -> (+x)
//            BN1 stays an elementwise scale/shift, BN2 is folded into conv1
//   * heads  conv1x1 -> BN -> flatten -> Linear x3 (activation = Identity,
//            NNetArchitecture.py:86-102): purely affine, folded into one
//            [A+3] x [H*W*C] matrix; softmax in fp32 (== exp(log_softmax)).
// The 3x3 convolutions run on the tensor cores as implicit GEMMs
// (mma.sync m16n8k16, bf16 operands, fp32 accumulation); the residual stream,
// BN1, biases, heads and softmax are fp32.  This is the bf16 performance mode;
// the strict-fp32 parity path stays PyTorch/cuDNN (azb200/nnet.py).
//
// Tile: NB = 8 boards per CTA -> 8*H*W = 336 output positions = 21 m-tiles of
// 16 rows for a 6x7 board; 11 warps (a pair of m-tiles each).  Shared memory
// (218 KB): fp32 residual x, bf16 activations a / b (zero-padded (H+2)x(W+2)
// frames, so a 3x3 tap is a constant row offset), double-buffered bf16 weights
// of the current / next layer, the folded head matrix.  BN1 + ReLU of the next
// block is applied in the epilogue of the convolution that produces x, so a
// block costs two tensor-core passes and two barriers; the heads are one more
// small MMA over the bf16 copy of the final x.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/azb200_nn.h"

namespace {

constexpr int NB = 8;          // boards per CTA
constexpr int CH = 32;         // trunk channels
constexpr int THREADS = 352;    // 11 warps: warp w owns m-tiles 2w, 2w+1 of the 21 (6x7 boards)
constexpr int RS_A = 40;       // bf16 row stride of a / b (80 B: conflict-free ldmatrix rows)
constexpr int RS_X = 40;       // fp32 row stride of x (float2 epilogue accesses conflict-free per half-warp)
constexpr int KW = 9 * CH;     // 288: K of a full 3x3 conv
constexpr int RS_W = KW + 8;   // 296: bf16 row stride of the weight matrix [cout][k]
constexpr int NHEAD = 16;      // head outputs padded to two n-tiles
constexpr int MAXD = 6;        // residual blocks supported by the shared-memory parameter block
constexpr int MAXL = 1 + 2 * MAXD;
constexpr int PRM_FLOATS = MAXL * CH + 2 * MAXD * CH;

__device__ __forceinline__ void ldmatrix_x4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// ---- TMA bulk copy (global -> shared, completion on an mbarrier) ---------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

template <int H, int W>
struct Geo {
    static constexpr int P = H * W;                  // positions per board
    static constexpr int PW = W + 2;
    static constexpr int PP = (H + 2) * (W + 2);     // padded frame
    static constexpr int M = NB * P;                 // rows of the implicit GEMM
    static constexpr int MT = M / 16;
    static_assert(M % 16 == 0, "NB * H * W must be a multiple of 16");
    __device__ __forceinline__ static int padded(int m)
    {
        const int b = m / P, c = m - b * P, y = c / W, x = c - y * W;
        return b * PP + (y + 1) * PW + (x + 1);
    }
};

enum { EPI_TO_B = 1, EPI_TO_X = 2 };

// one 3x3 convolution layer over the CTA's tile with `kchunks` 16-channel chunks
// of input channels per tap.
//   EPI_TO_B : outb = bf16(relu(acc + bias))                      (conv1, BN2 folded)
//   EPI_TO_X : x = (add ? x : 0) + acc (+ bias, relu for the stem) and, for the next
//              tensor-core pass, outb = bf16(act(x * nsc + nsh))   (BN1 of the next block)
template <int H, int W, int EPI>
__device__ __forceinline__ void conv_layer(const __nv_bfloat16 *in, const __nv_bfloat16 *wsm, const float *bias,
                                           float *x, __nv_bfloat16 *outb, const int *ptab, int kchunks, int warp,
                                           int lane, bool add, const float *nsc, const float *nsh)
{
    using G = Geo<H, W>;
    const uint32_t in_s = (uint32_t)__cvta_generic_to_shared(in);
    const uint32_t w_s = (uint32_t)__cvta_generic_to_shared(wsm);
    const int kper = kchunks * 16;                   // K per tap
    // each warp works on a pair of m-tiles at once: the B fragments are shared and
    // eight independent accumulator chains keep the tensor pipe busy
    for (int mt0 = warp * 2; mt0 < G::MT; mt0 += (THREADS / 32) * 2) {
        const bool two = mt0 + 1 < G::MT;                           // warp-uniform
        const int prow0 = ptab[mt0 * 16 + (lane & 15)];             // A rows this lane addresses
        const int prow1 = ptab[(two ? mt0 + 1 : mt0) * 16 + (lane & 15)];
        const uint32_t a_lane0 = in_s + (uint32_t)((prow0 * RS_A + (lane >> 4) * 8) * 2);
        const uint32_t a_lane1 = in_s + (uint32_t)((prow1 * RS_A + (lane >> 4) * 8) * 2);
        float acc[2][4][4];
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int i = 0; i < 4; i++) { acc[u][i][0] = acc[u][i][1] = acc[u][i][2] = acc[u][i][3] = 0.0f; }
        // B rows: n = nt2*16 + (lane/16)*8 + lane%8, k offset ((lane/8)%2)*8
        const uint32_t b_lane = w_s + (uint32_t)(((((lane >> 4) << 3) + (lane & 7)) * RS_W + ((lane >> 3) & 1) * 8) * 2);
#pragma unroll
        for (int tap = 0; tap < 9; tap++) {
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            const int toff = (dy * G::PW + dx) * RS_A * 2;
            for (int kc = 0; kc < kchunks; kc++) {
                uint32_t a0, a1, a2, a3, e0, e1, e2, e3;
                ldmatrix_x4(a0, a1, a2, a3, a_lane0 + toff + kc * 32);
                ldmatrix_x4(e0, e1, e2, e3, a_lane1 + toff + kc * 32);
                const int kbase = (tap * kper + kc * 16) * 2;
                uint32_t b0, b1, b2, b3, b4, b5, b6, b7;
                ldmatrix_x4(b0, b1, b2, b3, b_lane + kbase);
                ldmatrix_x4(b4, b5, b6, b7, b_lane + 16 * RS_W * 2 + kbase);
                mma_bf16(acc[0][0], a0, a1, a2, a3, b0, b1);
                mma_bf16(acc[0][1], a0, a1, a2, a3, b2, b3);
                mma_bf16(acc[0][2], a0, a1, a2, a3, b4, b5);
                mma_bf16(acc[0][3], a0, a1, a2, a3, b6, b7);
                if (two) {
                    mma_bf16(acc[1][0], e0, e1, e2, e3, b0, b1);
                    mma_bf16(acc[1][1], e0, e1, e2, e3, b2, b3);
                    mma_bf16(acc[1][2], e0, e1, e2, e3, b4, b5);
                    mma_bf16(acc[1][3], e0, e1, e2, e3, b6, b7);
                }
            }
        }
        // epilogue: c0,c1 -> row g, cols 2t,2t+1; c2,c3 -> row g+8
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (u == 1 && !two) break;
            const int mt = mt0 + u;
            const int m0 = mt * 16 + g, m1 = m0 + 8;               // x is indexed by the GEMM row, a / b by the padded frame
            const int p0 = ptab[m0], p1 = ptab[m1];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                const int col = nt * 8 + 2 * t;
                const float bx = bias ? bias[col] : 0.0f, by = bias ? bias[col + 1] : 0.0f;
                float r00 = acc[u][nt][0] + bx, r01 = acc[u][nt][1] + by;
                float r10 = acc[u][nt][2] + bx, r11 = acc[u][nt][3] + by;
                if (EPI == EPI_TO_B) {
                    r00 = fmaxf(r00, 0.0f); r01 = fmaxf(r01, 0.0f); r10 = fmaxf(r10, 0.0f); r11 = fmaxf(r11, 0.0f);
                } else {
                    float2 *q0 = reinterpret_cast<float2 *>(x + m0 * RS_X + col);
                    float2 *q1 = reinterpret_cast<float2 *>(x + m1 * RS_X + col);
                    if (add) {                                      // residual: x += conv2(b)
                        const float2 v0 = *q0, v1 = *q1;
                        r00 += v0.x; r01 += v0.y; r10 += v1.x; r11 += v1.y;
                    } else {                                        // stem: x = relu(conv + bias)
                        r00 = fmaxf(r00, 0.0f); r01 = fmaxf(r01, 0.0f); r10 = fmaxf(r10, 0.0f); r11 = fmaxf(r11, 0.0f);
                    }
                    *q0 = make_float2(r00, r01);
                    *q1 = make_float2(r10, r11);
                    if (nsc != nullptr) {                           // relu(bn1(x)) of the block that consumes x next
                        const float sx = nsc[col], sy = nsc[col + 1], tx = nsh[col], ty = nsh[col + 1];
                        r00 = fmaxf(fmaf(r00, sx, tx), 0.0f); r01 = fmaxf(fmaf(r01, sy, ty), 0.0f);
                        r10 = fmaxf(fmaf(r10, sx, tx), 0.0f); r11 = fmaxf(fmaf(r11, sy, ty), 0.0f);
                    }
                }
                *reinterpret_cast<__nv_bfloat162 *>(outb + p0 * RS_A + col) = __floats2bfloat162_rn(r00, r01);
                *reinterpret_cast<__nv_bfloat162 *>(outb + p1 * RS_A + col) = __floats2bfloat162_rn(r10, r11);
            }
        }
    }
}

template <int H, int W, int NOUT>
__global__ void __launch_bounds__(THREADS, 1)
k_resnet_fused(const float *__restrict__ obs, float *__restrict__ policy, float *__restrict__ value, int B, int in_ch,
               int depth, const __nv_bfloat16 *__restrict__ wconv, const float *__restrict__ cbias,
               const float *__restrict__ bn_scale, const float *__restrict__ bn_shift,
               const __nv_bfloat16 *__restrict__ whead, const float *__restrict__ bhead)
{
    using G = Geo<H, W>;
    constexpr int KH = G::P * CH;                    // head K: pos * CH + ch
    constexpr int RS_H = KH + 8;                     // bf16 row stride of the head matrix
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *x = reinterpret_cast<float *>(smem_raw);                                   // [M][RS_X] residual stream
    __nv_bfloat16 *a = reinterpret_cast<__nv_bfloat16 *>(x + G::M * RS_X);            // [NB*PP][RS_A]
    __nv_bfloat16 *b = a + NB * G::PP * RS_A;
    __nv_bfloat16 *w0 = b + NB * G::PP * RS_A;                                        // [CH][RS_W] x 2
    __nv_bfloat16 *w1 = w0 + CH * RS_W;
    __nv_bfloat16 *wh = w1 + CH * RS_W;                                               // [NHEAD][RS_H]
    int *ptab = reinterpret_cast<int *>(wh + NHEAD * RS_H);                           // [M] GEMM row -> padded frame row
    float *prm = reinterpret_cast<float *>(ptab + G::M);                              // biases [L][CH], bn scale/shift [D][CH] x2
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(prm + PRM_FLOATS);   // mbarriers: w0, w1, heads
    float *red = reinterpret_cast<float *>(w0);                                       // heads: [warps][NB][NHEAD] (aliases w0)
    float *fin = reinterpret_cast<float *>(w1);                                       // heads: [NB][NHEAD] logits (aliases w1)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int board0 = blockIdx.x * NB;
    const int layers = 1 + 2 * depth;
    constexpr uint32_t WBYTES = CH * RS_W * 2;
    const uint32_t bar_w[2] = {(uint32_t)__cvta_generic_to_shared(bars), (uint32_t)__cvta_generic_to_shared(bars + 1)};
    const uint32_t bar_h = (uint32_t)__cvta_generic_to_shared(bars + 2);
    uint32_t phase[2] = {0u, 0u};

    if (tid == 0) {
        mbar_init(bar_w[0], 1); mbar_init(bar_w[1], 1); mbar_init(bar_h, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    // one thread streams a whole layer's weight matrix with a single bulk copy (all readers of
    // the destination buffer have passed a __syncthreads before this is called)
    auto prefetch_w = [&](int layer, __nv_bfloat16 *dst, int buf) {
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar_w[buf], WBYTES);
            bulk_g2s((uint32_t)__cvta_generic_to_shared(dst), reinterpret_cast<const char *>(wconv) + (size_t)layer * WBYTES,
                     WBYTES, bar_w[buf]);
        }
    };
    auto wait_w = [&](int buf) { mbar_wait(bar_w[buf], phase[buf]); phase[buf] ^= 1u; };
    prefetch_w(0, w0, 0);
    if (tid == 0) {   // the folded head matrix, needed only at the very end
        mbar_expect_tx(bar_h, (uint32_t)(NHEAD * RS_H * 2));
        bulk_g2s((uint32_t)__cvta_generic_to_shared(wh), whead, (uint32_t)(NHEAD * RS_H * 2), bar_h);
    }
    // per-channel parameters -> shared memory
    for (int i = tid; i < layers * CH; i += THREADS) prm[i] = cbias[i];
    for (int i = tid; i < depth * CH; i += THREADS) { prm[MAXL * CH + i] = bn_scale[i]; prm[MAXL * CH + MAXD * CH + i] = bn_shift[i]; }
    const float *s_bias = prm, *s_sc = prm + MAXL * CH, *s_sh = prm + MAXL * CH + MAXD * CH;

    // zero the activation frames (borders must read as zero padding), build the row table, load the observation
    {
        uint4 *z = reinterpret_cast<uint4 *>(a);
        const int n16 = 2 * NB * G::PP * RS_A * 2 / 16;
        for (int i = tid; i < n16; i += THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int m = tid; m < G::M; m += THREADS) ptab[m] = G::padded(m);
    }
    __syncthreads();
    for (int i = tid; i < NB * in_ch * G::P; i += THREADS) {
        const int bl = i / (in_ch * G::P), r = i - bl * in_ch * G::P, c = r / G::P, pos = r - c * G::P;
        const int gb = board0 + bl;
        const float v = gb < B ? obs[(size_t)gb * in_ch * G::P + r] : 0.0f;
        b[ptab[bl * G::P + pos] * RS_A + c] = __float2bfloat16(v);
    }

    // stem: x = relu(conv(obs) + bias0), a = relu(bn1_0(x))   (K = 9 taps x 16 channels, channels >= in_ch are zero)
    __syncthreads();
    wait_w(0);
    if (layers > 1) prefetch_w(1, w1, 1);
    conv_layer<H, W, EPI_TO_X>(b, w0, s_bias, x, a, ptab, 1, warp, lane, false, depth > 0 ? s_sc : nullptr, s_sh);
    __syncthreads();

    for (int blk = 0; blk < depth; blk++) {
        const int l1 = 1 + 2 * blk, l2 = l1 + 1;
        __nv_bfloat16 *wl1 = (l1 & 1) ? w1 : w0, *wl2 = (l2 & 1) ? w1 : w0;
        prefetch_w(l2, wl2, l2 & 1);
        wait_w(l1 & 1);
        // b = relu(conv1(a) + bias)   (BN2 folded)
        conv_layer<H, W, EPI_TO_B>(a, wl1, s_bias + l1 * CH, x, b, ptab, 2, warp, lane, false, nullptr, nullptr);
        __syncthreads();
        if (l2 + 1 < layers) prefetch_w(l2 + 1, wl1, l1 & 1);
        wait_w(l2 & 1);
        // x += conv2(b);  a = relu(bn1_{blk+1}(x))  (the last block leaves a = bf16(x) for the heads)
        const bool last = blk + 1 == depth;
        conv_layer<H, W, EPI_TO_X>(b, wl2, nullptr, x, a, ptab, 2, warp, lane, true,
                                   last ? nullptr : s_sc + (blk + 1) * CH, s_sh + (blk + 1) * CH);
        __syncthreads();
    }
    mbar_wait(bar_h, 0u);

    // heads: logits[board][j] = sum_k wh[j][k] * a[board][k] + bhead[j], k = pos*CH + ch, as one small MMA
    // (rows 0-7 = boards, rows 8-15 duplicate them); the 2*P k-steps are split over the warps.
    {
        const uint32_t a_s = (uint32_t)__cvta_generic_to_shared(a);
        const uint32_t h_s = (uint32_t)__cvta_generic_to_shared(wh);
        float acc[2][4];
#pragma unroll
        for (int i = 0; i < 2; i++) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f; }
        const int brd = lane & 7;
        const uint32_t h_lane = h_s + (uint32_t)(((((lane >> 4) << 3) + (lane & 7)) * RS_H + ((lane >> 3) & 1) * 8) * 2);
        for (int ks = warp; ks < 2 * G::P; ks += THREADS / 32) {
            const int pos = ks >> 1, chunk = ks & 1;
            const uint32_t addr = a_s + (uint32_t)((ptab[brd * G::P + pos] * RS_A + chunk * 16 + (lane >> 4) * 8) * 2);
            uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
            ldmatrix_x4(a0, a1, a2, a3, addr);
            ldmatrix_x4(b0, b1, b2, b3, h_lane + ks * 32);
            mma_bf16(acc[0], a0, a1, a2, a3, b0, b1);
            mma_bf16(acc[1], a0, a1, a2, a3, b2, b3);
        }
        const int g = lane >> 2, t = lane & 3;                     // c0,c1: row g (board g), cols 2t, 2t+1
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            red[(warp * NB + g) * NHEAD + nt * 8 + 2 * t] = acc[nt][0];
            red[(warp * NB + g) * NHEAD + nt * 8 + 2 * t + 1] = acc[nt][1];
        }
        __syncthreads();
        if (tid < NB * NHEAD) {
            float v = 0.0f;
            for (int wq = 0; wq < THREADS / 32; wq++) v += red[wq * NB * NHEAD + tid];
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
}

template <int H, int W, int NOUT>
constexpr size_t smem_bytes()
{
    using G = Geo<H, W>;
    return (size_t)G::M * RS_X * 4 + 2 * (size_t)NB * G::PP * RS_A * 2 + 2 * (size_t)CH * RS_W * 2 +
           (size_t)NHEAD * (G::P * CH + 8) * 2 + (size_t)G::M * 4 + (size_t)PRM_FLOATS * 4 + 32;
}

}  // namespace

extern "C" int azb_nn_weight_row_stride(void) { return RS_W; }
extern "C" int azb_nn_head_row_stride(void) { return 6 * 7 * CH + 8; }
extern "C" int azb_nn_boards_per_cta(void) { return NB; }

extern "C" int azb_nn_forward(const azb_nn_weights *w, const float *obs, float *policy, float *value, int32_t batch,
                              void *stream)
{
    if (!w || !obs || !policy || !value || batch <= 0) return -7;
    if (w->channels != CH || w->board_h != 6 || w->board_w != 7 || w->action_size != 7 || w->in_channels > 16 ||
        w->depth < 0 || w->depth > MAXD)
        return -1;
    cudaStream_t s = (cudaStream_t)stream;
    static bool configured = false;
    constexpr size_t SM = smem_bytes<6, 7, 10>();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_resnet_fused<6, 7, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM);
        if (e != cudaSuccess) return -2;
        configured = true;
    }
    const int grid = (batch + NB - 1) / NB;
    k_resnet_fused<6, 7, 10><<<grid, THREADS, SM, s>>>(obs, policy, value, batch, w->in_channels, w->depth,
                                                      reinterpret_cast<const __nv_bfloat16 *>(w->wconv), w->cbias, w->bn_scale,
                                                      w->bn_shift, reinterpret_cast<const __nv_bfloat16 *>(w->whead), w->bhead);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ---- host tensor upload by the SMs --------------------------------------------------------------------------
// The reference-facing surface hands the network HOST tensors every simulation (Coach.py:339, NNetWrapper.py:227:
// `batch.cuda()`).  On this platform a cudaMemcpyAsync host->device of a few MB pays ~190 us of DMA start-up
// (scripts/pcie_probe.py: 2.75 MB in 231 us, 64 MB at 55 GB/s); pinned host memory is mapped into the device's
// address space, so a kernel can stream it over PCIe directly.
namespace {
__global__ void __launch_bounds__(256) k_upload(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16, const unsigned char *src_tail,
                                                unsigned char *dst_tail, int tail)
{
    // eight independent 16-byte loads per thread in flight: PCIe needs ~100 KB outstanding, and it should come from FEW
    // CTAs -- a CTA of this kernel on an SM keeps a trunk CTA (which owns its SM's registers) off it, and with its pair
    // partner a whole TPC, for the ~50 us the upload lasts
    constexpr int U = 8;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n16; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + i + u * stride));
#pragma unroll
        for (int u = 0; u < U; u++) dst[i + u * stride] = v[u];
    }
    for (; i < n16; i += stride) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + i));
        dst[i] = v;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}
}  // namespace

extern "C" int azb_upload_pinned(void *dst_device, const void *src_pinned_host, int64_t bytes, void *stream)
{
    if (!dst_device || !src_pinned_host || bytes < 0) return -7;
    if (bytes == 0) return 0;
    if ((((uintptr_t)dst_device) | ((uintptr_t)src_pinned_host)) & 15) return -7;
    const size_t n16 = (size_t)bytes / 16;
    const int tail = (int)(bytes - (int64_t)n16 * 16);
    // few CTAs: PCIe needs ~100 KB in flight, not SMs -- the upload should leave the SMs to the kernels it overlaps with
    // (e2e leg, M sims/s with 2 / 4 / 8 / 12 / 16 / 32 CTAs of eight loads per thread: 17.5 / 18.5 / 22.8 / 22.4 / 22.7 / 21.6)
    static int cap = 0;
    if (cap == 0) { const char *env = getenv("AZB_UPLOAD_CTAS"); cap = env ? atoi(env) : 8; if (cap < 1) cap = 1; }
    int grid = (int)((n16 + 255) / 256);
    grid = grid < 1 ? 1 : (grid > cap ? cap : grid);
    k_upload<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint4 *>(dst_device), reinterpret_cast<const uint4 *>(src_pinned_host), n16,
                                                     reinterpret_cast<const unsigned char *>(src_pinned_host) + n16 * 16,
                                                     reinterpret_cast<unsigned char *>(dst_device) + n16 * 16, tail);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
