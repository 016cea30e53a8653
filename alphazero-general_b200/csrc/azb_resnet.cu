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
// 16 rows for a 6x7 board; 11 warps (a pair of m-tiles each); shared memory: fp32 residual x, bf16
// activations a / b (zero-padded (H+2)x(W+2) frames, so a 3x3 tap is a constant
// row offset), double-buffered bf16 weights of the current / next layer.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/azb200_nn.h"

namespace {

constexpr int NB = 8;          // boards per CTA
constexpr int CH = 32;         // trunk channels
constexpr int THREADS = 352;    // 11 warps: warp w owns m-tiles 2w, 2w+1 of the 21 (6x7 boards)
constexpr int RS_A = 40;       // bf16 row stride of a / b (80 B: conflict-free ldmatrix rows)
constexpr int RS_X = 36;       // fp32 row stride of x
constexpr int KW = 9 * CH;     // 288: K of a full 3x3 conv
constexpr int RS_W = KW + 8;   // 296: bf16 row stride of the weight matrix [cout][k]

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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

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

enum { EPI_STEM = 0, EPI_TO_B = 1, EPI_ADD_X = 2 };

// one 3x3 convolution layer over the CTA's tile: out = conv(in) with `kchunks`
// 16-channel chunks of input channels per tap.
template <int H, int W, int EPI>
__device__ __forceinline__ void conv_layer(const __nv_bfloat16 *in, const __nv_bfloat16 *wsm, const float *bias,
                                           float *x, __nv_bfloat16 *outb, const int *ptab, int kchunks, int warp, int lane)
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
                if (EPI == EPI_ADD_X) {
                    float2 *q0 = reinterpret_cast<float2 *>(x + m0 * RS_X + col);
                    float2 *q1 = reinterpret_cast<float2 *>(x + m1 * RS_X + col);
                    float2 v0 = *q0, v1 = *q1;
                    v0.x += acc[u][nt][0]; v0.y += acc[u][nt][1]; v1.x += acc[u][nt][2]; v1.y += acc[u][nt][3];
                    *q0 = v0; *q1 = v1;
                } else {
                    const float bx = bias[col], by = bias[col + 1];
                    const float r00 = fmaxf(acc[u][nt][0] + bx, 0.0f), r01 = fmaxf(acc[u][nt][1] + by, 0.0f);
                    const float r10 = fmaxf(acc[u][nt][2] + bx, 0.0f), r11 = fmaxf(acc[u][nt][3] + by, 0.0f);
                    if (EPI == EPI_STEM) {
                        *reinterpret_cast<float2 *>(x + m0 * RS_X + col) = make_float2(r00, r01);
                        *reinterpret_cast<float2 *>(x + m1 * RS_X + col) = make_float2(r10, r11);
                    } else {
                        *reinterpret_cast<__nv_bfloat162 *>(outb + p0 * RS_A + col) = __floats2bfloat162_rn(r00, r01);
                        *reinterpret_cast<__nv_bfloat162 *>(outb + p1 * RS_A + col) = __floats2bfloat162_rn(r10, r11);
                    }
                }
            }
        }
    }
}

template <int H, int W, int NOUT>
__global__ void __launch_bounds__(THREADS, 1)
k_resnet_fused(const float *__restrict__ obs, float *__restrict__ policy, float *__restrict__ value, int B, int in_ch,
               int depth, const __nv_bfloat16 *__restrict__ wconv, const float *__restrict__ cbias,
               const float *__restrict__ bn_scale, const float *__restrict__ bn_shift,
               const float *__restrict__ whead, const float *__restrict__ bhead)
{
    using G = Geo<H, W>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *x = reinterpret_cast<float *>(smem_raw);                                   // [M][RS_X] residual stream
    __nv_bfloat16 *a = reinterpret_cast<__nv_bfloat16 *>(x + G::M * RS_X);            // [NB*PP][RS_A]
    __nv_bfloat16 *b = a + NB * G::PP * RS_A;
    __nv_bfloat16 *w0 = b + NB * G::PP * RS_A;                                        // [CH][RS_W] x 2
    __nv_bfloat16 *w1 = w0 + CH * RS_W;
    int *ptab = reinterpret_cast<int *>(w1 + CH * RS_W);                              // [M] GEMM row -> padded frame row
    float *red = reinterpret_cast<float *>(ptab + G::M);                              // [warps][NB*NOUT] head partial sums
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int board0 = blockIdx.x * NB;
    const int layers = 1 + 2 * depth;
    constexpr int WBYTES = CH * RS_W * 2;

    auto prefetch_w = [&](int layer, __nv_bfloat16 *dst) {
        const char *src = reinterpret_cast<const char *>(wconv) + (size_t)layer * WBYTES;
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
        for (int i = tid; i < WBYTES / 16; i += THREADS) cp_async16(d + i * 16, src + i * 16);
        cp_async_commit();
    };
    prefetch_w(0, w0);

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
        a[ptab[bl * G::P + pos] * RS_A + c] = __float2bfloat16(v);
    }

    // stem: x = relu(conv(a) + bias0)       (K = 9 taps x 16 channels, channels >= in_ch are zero)
    cp_async_wait_all();
    __syncthreads();
    if (layers > 1) prefetch_w(1, w1);
    conv_layer<H, W, EPI_STEM>(a, w0, cbias, x, nullptr, ptab, 1, warp, lane);
    __syncthreads();

    for (int blk = 0; blk < depth; blk++) {
        // a = relu(bn1(x))
        const float *sc = bn_scale + blk * CH, *sh = bn_shift + blk * CH;
        {
            const int c2 = (tid & 15) * 2;                 // THREADS % 16 == 0: a thread keeps its channel pair
            const float s0 = sc[c2], s1 = sc[c2 + 1], t0 = sh[c2], t1 = sh[c2 + 1];
            for (int m = tid >> 4; m < G::M; m += THREADS / 16) {
                const float2 v = *reinterpret_cast<const float2 *>(x + m * RS_X + c2);
                const float r0 = fmaxf(fmaf(v.x, s0, t0), 0.0f), r1 = fmaxf(fmaf(v.y, s1, t1), 0.0f);
                *reinterpret_cast<__nv_bfloat162 *>(a + ptab[m] * RS_A + c2) = __floats2bfloat162_rn(r0, r1);
            }
        }
        const int l1 = 1 + 2 * blk, l2 = l1 + 1;
        __nv_bfloat16 *wl1 = (l1 & 1) ? w1 : w0, *wl2 = (l2 & 1) ? w1 : w0;
        cp_async_wait_all();
        __syncthreads();
        prefetch_w(l2, wl2);
        // b = relu(conv1(a) + bias)   (BN2 folded)
        conv_layer<H, W, EPI_TO_B>(a, wl1, cbias + l1 * CH, x, b, ptab, 2, warp, lane);
        cp_async_wait_all();
        __syncthreads();
        if (l2 + 1 < layers) prefetch_w(l2 + 1, wl1);
        // x += conv2(b)
        conv_layer<H, W, EPI_ADD_X>(b, wl2, nullptr, x, nullptr, ptab, 2, warp, lane);
        __syncthreads();
    }

    // heads: logits[board][j] = sum_f whead[j][f] * x[board][f] + bhead[j],  f = pos*CH + ch.
    // Each thread owns a slice of the features for ALL boards of the tile, so a head weight is
    // read once per CTA; partial sums are reduced by shuffles, then across warps in shared memory.
    {
        float acc[NB][NOUT];
#pragma unroll
        for (int bq = 0; bq < NB; bq++)
#pragma unroll
            for (int j = 0; j < NOUT; j++) acc[bq][j] = 0.0f;
        for (int f = tid; f < G::P * CH; f += THREADS) {
            const int pos = f / CH, ch = f - pos * CH;
            float wv[NOUT];
#pragma unroll
            for (int j = 0; j < NOUT; j++) wv[j] = whead[(size_t)j * G::P * CH + f];
#pragma unroll
            for (int bq = 0; bq < NB; bq++) {
                const float xv = x[(bq * G::P + pos) * RS_X + ch];
#pragma unroll
                for (int j = 0; j < NOUT; j++) acc[bq][j] = fmaf(wv[j], xv, acc[bq][j]);
            }
        }
#pragma unroll
        for (int bq = 0; bq < NB; bq++)
#pragma unroll
            for (int j = 0; j < NOUT; j++) {
                float v = acc[bq][j];
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if (lane == 0) red[warp * (NB * NOUT) + bq * NOUT + j] = v;
            }
        __syncthreads();
        if (tid < NB) {
            const int gb = board0 + tid;
            if (gb < B) {
                float lg[NOUT];
#pragma unroll
                for (int j = 0; j < NOUT; j++) {
                    float v = bhead[j];
                    for (int wq = 0; wq < THREADS / 32; wq++) v += red[wq * (NB * NOUT) + tid * NOUT + j];
                    lg[j] = v;
                }
                constexpr int A = NOUT - 3;
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
}

template <int H, int W, int NOUT>
constexpr size_t smem_bytes()
{
    using G = Geo<H, W>;
    return (size_t)G::M * RS_X * 4 + 2 * (size_t)NB * G::PP * RS_A * 2 + 2 * (size_t)CH * RS_W * 2 + (size_t)G::M * 4 +
           (size_t)(THREADS / 32) * NB * NOUT * 4;
}

}  // namespace

extern "C" int azb_nn_weight_row_stride(void) { return RS_W; }
extern "C" int azb_nn_boards_per_cta(void) { return NB; }

extern "C" int azb_nn_forward(const azb_nn_weights *w, const float *obs, float *policy, float *value, int32_t batch,
                              void *stream)
{
    if (!w || !obs || !policy || !value || batch <= 0) return -7;
    if (w->channels != CH || w->board_h != 6 || w->board_w != 7 || w->action_size != 7 || w->in_channels > 16 ||
        w->depth < 0)
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
                                                  w->bn_shift, w->whead, w->bhead);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
