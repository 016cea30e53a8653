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
// plane apart (LBO = plane size).  Shifting the operand by a vertical tap is just
// `start address += dy * 128 B`: no im2col, no copies; and the epilogue's stores
// (one 16-byte chunk per thread, consecutive threads = consecutive rows) are
// bank-conflict free.
//
// A tcgen05.mma with M = 128, K = 16 costs ~44 cycles at N = 32 but only ~56 at
// N = 96 (measured, scripts/umma_bench.cu: the A operand read from shared memory
// dominates), so the three horizontal taps share one A read: the B operand holds
// [dx][cout] = 96 columns and the MMA produces P_dx[i] = sum_dy A[i + 8 dy] W[dy,dx];
// the convolution is D[r] = P_-1[r-1] + P_0[r] + P_+1[r+1], a one-lane shuffle in the
// epilogue (rows r-1 / r+1 that fall outside the 8-row group are padding, where P is
// exactly zero).  6 MMAs per tile and layer instead of 18.  The stem (<= 8 input
// planes) packs two vertical taps into one K = 16 step by pointing LBO at the
// second tap's rows.
//
// Tensor memory (512 columns): the fp32 residual stream x of the 7 tiles (224
// columns, owned by the epilogue threads for the whole network) + a ring of three
// 96-column accumulators.  Warp roles: warps 0-2 issue the MMAs, one per ring slot
// (one elected thread each, uniform registers; issue is blocking and the thread's
// own latencies would otherwise idle the tensor pipe), tile after tile across
// layers, throttled only by the ring and by the rows of the previous layer a tile
// needs (per-tile progress counters, no CTA-wide barrier inside the network);
// warp 3 streams each layer's weights with one bulk
// copy (UBLKCP) into a ring of three buffers; three groups of eight warps (TMEM lane
// quadrant x 16-channel half) drain the accumulators (tcgen05.ld -> shuffle-add ->
// bias / residual / BN / ReLU in packed f32x2 -> bf16 -> shared memory).  The affine heads are one small mma.sync over the final
// activation, softmax in fp32.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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
constexpr int PADR = 16;               // zero rows in front of / behind the boards
constexpr int FROWS = ROWS + 2 * PADR; // 928
constexpr int PLANE = FROWS * 16;      // bytes of one 8-channel chunk plane (14848)
constexpr int FRAME = 4 * PLANE;       // 59392
constexpr int NACC = 3 * CH;           // 96 accumulator columns: [dx][cout]
constexpr int KCH = 12;                // 16-byte K chunks per layer: 3 vertical taps x 4
constexpr int WCHUNK = NACC * 16;      // 1536 bytes: one K chunk of the B operand
constexpr int WL_BYTES = KCH * WCHUNK; // 18432: [k chunk][dx*32 + cout][8 cin] bf16
constexpr int NWBUF = 3, NPBUF = 3;
constexpr int NOUT = 10;               // 7 policy logits + 3 value logits
constexpr int KH = FB * CH;            // head K: frame row * 32 + channel (1792)
constexpr int RS_H = KH + 8;           // bf16 row stride of the head matrix
constexpr int MAXD = 6, MAXL = 1 + 2 * MAXD;
constexpr int PRM_FLOATS = MAXL * CH + 2 * MAXD * CH;
constexpr int PROD_WARP = NPBUF;       // warps 0..2: MMA issue for ring slot 0..2; warp 3: weight producer
constexpr int EPI_WARP0 = NPBUF + 1;   // then the epilogue warps
constexpr int GRP_WARPS = 8;           // per ring slot: 4 TMEM lane quadrants x 2 column halves
constexpr int WARPS = EPI_WARP0 + GRP_WARPS * NPBUF;   // 28
constexpr int THREADS = WARPS * 32;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_X = 0, COL_P = 224;   // x tiles at 32*t; accumulator ring at 224 + 96*b
static_assert(COL_P + NPBUF * NACC <= TMEM_COLS, "tensor memory budget");

// mbarriers
enum { BAR_PFULL = 0, BAR_PEMPTY = BAR_PFULL + NPBUF, BAR_READY = BAR_PEMPTY + NPBUF /* [TILES]: rows 128t-8 .. 128t+135 written */,
       BAR_WFULL = BAR_READY + TILES, BAR_WEMPTY = BAR_WFULL + NWBUF, BAR_HEAD = BAR_WEMPTY + NWBUF, NBARS };
// Shared-memory bandwidth is the shared resource: a tcgen05.mma at N = 96 reads 7 KB of operands in 56 cycles (the
// full 128 B/clk), so anything that polls shared memory steals from the tensor pipe.  Waits are mbarrier waits (the
// hardware suspends the thread); the only polled word is the turn counter that keeps the MMAs of a tile contiguous.
enum { CNT_TURN = 0 /* next tile whose MMAs may enter the tensor pipe */, NCNT = 1 };
constexpr int HB = FB + 1;             // rows per board of the head-input layout (57: board stride 912 B, conflict-free ldmatrix)

constexpr size_t SMEM_BYTES = 2 * (size_t)FRAME + (size_t)NWBUF * WL_BYTES + (size_t)NOUT * RS_H * 2 +
                              (size_t)PRM_FLOATS * 4 + 32 * 8 + 64 + 128;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// ---- PTX wrappers --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
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
// Bounded wait: a barrier that never completes is a programming error -- trap instead of hanging the GPU.  The
// suspend-time hint lets the hardware park the thread instead of returning to re-poll: the kernel is bound by
// instruction issue in the epilogue warps, and every retry of a waiting warp takes issue slots from a working one.
template <int BACKOFF_NS = 0>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 22); it++) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done)
                     : "r"(bar), "r"(parity), "r"(20000u)
                     : "memory");
        if (done) return;
        if (BACKOFF_NS > 0) __nanosleep(BACKOFF_NS);     // waiters with slack: every retry is a shared-memory transaction
    }
    __trap();
}
// the turn counter orders only the *issue* of tcgen05.mma by different threads (a performance matter): relaxed
// accesses.  (A release store compiles to MEMBAR.ALL.CTA, which stalls the issuing thread until its MMAs retire.)
__device__ __forceinline__ void turn_store(uint32_t addr, uint32_t v)
{
    asm volatile("st.relaxed.cta.shared::cta.u32 [%0], %1;\n" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void turn_wait(uint32_t addr, uint32_t g)
{
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 24); it++) {
        uint32_t v;
        asm volatile("ld.relaxed.cta.shared::cta.u32 %0, [%1];\n" : "=r"(v) : "r"(addr) : "memory");
        if (v >= g) return;                 // three polling threads per CTA: cheap; a nanosleep here costs ~500 cycles of turn latency
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
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NACC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

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
#define AZB_R16(v)                                                                                                     \
    "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),        \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
#define AZB_I16(v)                                                                                                     \
    "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),      \
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
// this thread's TMEM lane, 16 consecutive fp32 columns (no wait: the caller batches loads)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : AZB_R16(v)
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};\n" ::AZB_I16(v),
                 "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
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

// MMAs of one tile (elected thread).  trunk: 3 vertical taps x two 16-channel halves; stem: two K steps over
// chunk plane 0 ((dy=-1, dy=0) and (dy=+1, zero weights)).
__device__ __forceinline__ void issue_tile(uint32_t in_s, uint32_t w_s, uint32_t d_tmem, int t, bool stem)
{
    const uint32_t row0 = in_s + (uint32_t)((PADR + 128 * t) * 16);
    if (stem) {
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const uint64_t ad = umma_desc(row0 + (uint32_t)((2 * s - 1) * FW * 16), (uint32_t)(FW * 16), 128u);
            const uint64_t bd = umma_desc(w_s + (uint32_t)(2 * s * WCHUNK), (uint32_t)WCHUNK, 128u);
            umma_f16(d_tmem, ad, bd, s > 0 ? 1u : 0u);
        }
    } else {
#pragma unroll
        for (int s = 0; s < 6; s++) {
            const int dy = s / 2 - 1, half = s & 1;
            const uint64_t ad = umma_desc(row0 + (uint32_t)(2 * half * PLANE + dy * FW * 16), (uint32_t)PLANE, 128u);
            const uint64_t bd = umma_desc(w_s + (uint32_t)(2 * s * WCHUNK), (uint32_t)WCHUNK, 128u);
            umma_f16(d_tmem, ad, bd, s > 0 ? 1u : 0u);
        }
    }
}

enum { EPI_STEM = 0, EPI_CONV1 = 1, EPI_CONV2 = 2 };

__device__ __forceinline__ float2 f2(uint32_t lo, uint32_t hi) { return make_float2(__uint_as_float(lo), __uint_as_float(hi)); }
__device__ __forceinline__ uint32_t relu_bf16x2(float2 v)      // bf16x2(max(v, 0)): rounding is monotonic, so relu commutes with it
{
    __nv_bfloat162 h = __hmax2(__float22bfloat162_rn(v), __float2bfloat162_rn(0.0f));
    return *reinterpret_cast<uint32_t *>(&h);
}

// Epilogue of one tile for 16 of its 32 channels: this thread owns frame row `fr` = TMEM lane q*32 + lane and the
// channels [16*half, 16*half + 16).  Packed f32x2 arithmetic (FADD2 / FFMA2).
//   stem : x = relu(D + bias) -> TMEM;  a = relu(bn1_0(x))
//   conv1: b = relu(D + bias)                                   (BN2 folded)
//   conv2: x += D -> TMEM;  a = relu(bn1_next(x))  (last block: a = x)
// t_p / t_x / bias / nsc / nsh / out_row / dump_row already point at this thread's column half.
template <int EPI, bool DBG>
__device__ __forceinline__ void epilogue_tile(uint32_t t_p, uint32_t t_x, unsigned char *out_row, int out_plane, const float *bias,
                                              const float *nsc, const float *nsh, bool live, int lane, uint32_t bar_pempty,
                                              float *dump_row)
{
    uint32_t pm[16], p0[16], pp[16], xv[16];
    tmem_ld16(t_p, pm);
    tmem_ld16(t_p + (uint32_t)CH, p0);
    tmem_ld16(t_p + (uint32_t)(2 * CH), pp);
    if (EPI == EPI_CONV2) tmem_ld16(t_x, xv);
    tmem_ld_wait();
    tc_fence_before();                    // the accumulator is in registers: hand the ring slot back to the MMA warp
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_pempty);
    // D[r] = P_-1[r-1] + P_0[r] + P_+1[r+1]: rotate by one lane.  Lane 0 receives lane 31's value, a padding-column
    // row where P is exactly zero -- which is what row r-1 of lane 0 (also a padding column) holds; lane 31 is a
    // padding row itself and is never stored.
    // The two neighbour terms travel as packed fp16 pairs (11-bit mantissa, far below the bf16 rounding of the operands):
    // shuffles go through the shared-memory data pipe, which is the resource this kernel is bound by.
    const int src_up = (lane + 31) & 31, src_dn = (lane + 1) & 31;
    float2 r[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const __half2 hm = __floats2half2_rn(__uint_as_float(pm[2 * c]), __uint_as_float(pm[2 * c + 1]));
        const __half2 hp = __floats2half2_rn(__uint_as_float(pp[2 * c]), __uint_as_float(pp[2 * c + 1]));
        const uint32_t um = __shfl_sync(0xffffffffu, *reinterpret_cast<const uint32_t *>(&hm), src_up);
        const uint32_t up_ = __shfl_sync(0xffffffffu, *reinterpret_cast<const uint32_t *>(&hp), src_dn);
        const float2 up = __half22float2(*reinterpret_cast<const __half2 *>(&um));
        const float2 dn = __half22float2(*reinterpret_cast<const __half2 *>(&up_));
        r[c] = __fadd2_rn(__fadd2_rn(up, f2(p0[2 * c], p0[2 * c + 1])), dn);
    }
    float2 prm2[8];                        // 16 per-channel parameters with four 16-byte loads
    if (EPI != EPI_CONV2) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float4 b4 = reinterpret_cast<const float4 *>(bias)[c];
            prm2[2 * c] = make_float2(b4.x, b4.y); prm2[2 * c + 1] = make_float2(b4.z, b4.w);
        }
    }
    if (EPI == EPI_STEM) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            r[c] = __fadd2_rn(r[c], prm2[c]);
            r[c].x = fmaxf(r[c].x, 0.0f); r[c].y = fmaxf(r[c].y, 0.0f);
            xv[2 * c] = __float_as_uint(r[c].x); xv[2 * c + 1] = __float_as_uint(r[c].y);
        }
        tmem_st16(t_x, xv);
    } else if (EPI == EPI_CONV1) {
#pragma unroll
        for (int c = 0; c < 8; c++) r[c] = __fadd2_rn(r[c], prm2[c]);
    } else {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            r[c] = __fadd2_rn(r[c], f2(xv[2 * c], xv[2 * c + 1]));
            xv[2 * c] = __float_as_uint(r[c].x); xv[2 * c + 1] = __float_as_uint(r[c].y);
        }
        tmem_st16(t_x, xv);
    }
    uint32_t o[8];
    if (EPI != EPI_CONV1 && nsc == nullptr) {               // last block: a = bf16(x), no activation
#pragma unroll
        for (int c = 0; c < 8; c++) o[c] = pack_bf16(r[c].x, r[c].y);
    } else {
        if (EPI != EPI_CONV1) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float4 s4 = reinterpret_cast<const float4 *>(nsc)[c], h4 = reinterpret_cast<const float4 *>(nsh)[c];
                r[2 * c] = __ffma2_rn(r[2 * c], make_float2(s4.x, s4.y), make_float2(h4.x, h4.y));
                r[2 * c + 1] = __ffma2_rn(r[2 * c + 1], make_float2(s4.z, s4.w), make_float2(h4.z, h4.w));
            }
        }
#pragma unroll
        for (int c = 0; c < 8; c++) o[c] = relu_bf16x2(r[c]);
    }
    if (DBG && dump_row != nullptr) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162 *>(&o[c]);
            dump_row[2 * c] = live ? __low2float(h) : 0.0f;
            dump_row[2 * c + 1] = live ? __high2float(h) : 0.0f;
        }
    }
    if (live) {                           // padding rows are zero from the start and stay so
        *reinterpret_cast<uint4 *>(out_row) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4 *>(out_row + out_plane) = make_uint4(o[4], o[5], o[6], o[7]);
    }
    if (EPI != EPI_CONV1) tmem_st_wait();
}

template <bool DBG>
__global__ void __launch_bounds__(THREADS, 1)
k_resnet_tc(const float *__restrict__ obs, float *__restrict__ policy, float *__restrict__ value, int B, int in_ch, int depth,
            const unsigned char *__restrict__ wconv, const float *__restrict__ cbias, const float *__restrict__ bn_scale,
            const float *__restrict__ bn_shift, const __nv_bfloat16 *__restrict__ whead, const float *__restrict__ bhead,
            float *__restrict__ dump, int dump_layer, int n_full, const int *__restrict__ rows, const int *__restrict__ count_ptr,
            int sms)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *fa = smem;                                            // activation frame a
    unsigned char *fb = fa + FRAME;                                      // activation frame b (first: the observation)
    unsigned char *wb = fb + FRAME;                                      // weight ring [NWBUF][WL_BYTES]
    __nv_bfloat16 *wh = reinterpret_cast<__nv_bfloat16 *>(wb + NWBUF * WL_BYTES);   // [NOUT][RS_H]
    float *prm = reinterpret_cast<float *>(wh + NOUT * RS_H);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(prm + PRM_FLOATS);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 32);
    uint32_t *cnts = tmem_slot + 4;                                      // [NCNT] progress counters
    float *red = reinterpret_cast<float *>(wb);                          // heads: [WARPS][NB][16] (after the trunk)
    float *fin = red + WARPS * NB * 16;                                  // heads: [NB][16]

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    // CTAs [0, n_full) evaluate 16 boards (7 M-tiles); the CTAs behind them 8 boards (3.5 -> 4 M-tiles, 4/7 of the
    // time): the host turns the tiles of a partially filled last wave into twice as many half tiles
    // compact mode (rows != nullptr): evaluate boards rows[0 .. *count_ptr) of obs, answers to the same rows; the
    // batch size lives in device memory (the engine's select kernel counts the non-terminal leaves), so the wave
    // shaping the host does for a known batch is redone here
    if (rows != nullptr) {
        B = *count_ptr;
        const int T = (B + NB - 1) / NB, rem = T % sms;
        n_full = T;
        if (rem > 0 && 2 * rem <= sms) n_full = T - rem;
    }
    const bool full = (int)blockIdx.x < n_full;
    const int nb = full ? NB : NB / 2, tiles = full ? TILES : (TILES + 1) / 2;
    const int board0 = full ? (int)blockIdx.x * NB : n_full * NB + ((int)blockIdx.x - n_full) * (NB / 2);
    if (board0 >= B) return;
    // DBG build (azb_nn_forward_tc_debug): dump_layer = 99 writes per-warp phase clocks, else the activations of a layer
    const bool timing = DBG && dump != nullptr && dump_layer == 99;
    long long ts[5];
    if (timing) ts[0] = clock64();
    const int layers = 1 + 2 * depth;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t fa_s = smem_u32(fa), fb_s = smem_u32(fb), wb_s = smem_u32(wb);
#define BAR(i) (bar0 + 8u * (uint32_t)(i))

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < NPBUF; i++) { mbar_init(BAR(BAR_PFULL + i), 1); mbar_init(BAR(BAR_PEMPTY + i), GRP_WARPS); }
            // every epilogue warp of tiles t-1, t, t+1 arrives on READY[t]
            for (int i = 0; i < tiles; i++) mbar_init(BAR(BAR_READY + i), GRP_WARPS * ((i == 0 || i == tiles - 1) ? 2 : 3));
            for (int i = 0; i < NWBUF; i++) { mbar_init(BAR(BAR_WFULL + i), 1); mbar_init(BAR(BAR_WEMPTY + i), NPBUF); }
            mbar_init(BAR(BAR_HEAD), 1);
            for (int i = 0; i < NCNT; i++) cnts[i] = 0u;
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

    // observation -> chunk plane 0 of frame b (channels >= in_ch stay zero)
    for (int i = tid; i < nb * BH * BW; i += THREADS) {
        const int bl = i / (BH * BW), pos = i - bl * (BH * BW), y = pos / BW, xx = pos - y * BW;
        int gb = board0 + bl;
        const bool have = gb < B;
        if (have && rows != nullptr) gb = rows[gb];
        float c[8];
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] = (have && k < in_ch) ? obs[((size_t)gb * in_ch + k) * (BH * BW) + pos] : 0.0f;
        uint4 o;
        o.x = pack_bf16(c[0], c[1]); o.y = pack_bf16(c[2], c[3]); o.z = pack_bf16(c[4], c[5]); o.w = pack_bf16(c[6], c[7]);
        *reinterpret_cast<uint4 *>(fb + (size_t)(PADR + bl * FB + y * FW + xx) * 16) = o;
    }
    fence_proxy_async();
    __syncthreads();
    if (timing) ts[1] = clock64();

    const uint32_t cnt0 = smem_u32(cnts);
#define CNT(i) (cnt0 + 4u * (uint32_t)(i))
    if (warp < NPBUF) {
        // ---- MMA issuers: warp k feeds ring slot k (tiles g = k, k+3, ... across layers) --------------------
        if (elect_one_sync()) {
            const int total = layers * tiles;
            const uint32_t d_tmem = tmem_base + COL_P + (uint32_t)(warp * NACC);
            int l = 0, t = warp, cur_l = -1;
            uint32_t w_s = 0;
#pragma unroll 1
            for (int g = warp, use = 0; g < total; g += NPBUF, use++) {
                if (l != cur_l) {
                    if (cur_l >= 0) umma_commit(BAR(BAR_WEMPTY + cur_l % NWBUF));   // my MMAs of that layer are all issued
                    cur_l = l;
                    mbar_wait(BAR(BAR_WFULL + l % NWBUF), (uint32_t)((l / NWBUF) & 1));
                    w_s = wb_s + (uint32_t)((l % NWBUF) * WL_BYTES);
                }
                // rows 128t-8 .. 128t+135 of the previous layer's output, and my ring slot drained
                const bool trace = DBG && timing && blockIdx.x == 0;
                float *tl = trace ? dump + (size_t)1024 * 256 + (size_t)g * 16 : nullptr;
                if (trace) tl[0] = (float)(clock64() - ts[0]);
                if (l > 0) mbar_wait(BAR(BAR_READY + t), (uint32_t)((l - 1) & 1));
                if (trace) tl[1] = (float)(clock64() - ts[0]);
                if (use > 0) mbar_wait(BAR(BAR_PEMPTY + warp), (uint32_t)((use - 1) & 1));
                if (trace) tl[2] = (float)(clock64() - ts[0]);
                tc_fence_after();
                // the three issuers overlap their waits, but the MMAs of a tile enter the pipe back to back and in tile
                // order: interleaved, all three tiles would complete together and the ring would run in lock-step
                turn_wait(CNT(CNT_TURN), (uint32_t)g);
                if (trace) tl[3] = (float)(clock64() - ts[0]);
                issue_tile((l & 1) ? fa_s : fb_s, w_s, d_tmem, t, l == 0);   // stem: obs (b);  conv1: a;  conv2: b
                umma_commit(BAR(BAR_PFULL + warp));
                turn_store(CNT(CNT_TURN), (uint32_t)(g + 1));
                if (trace) tl[4] = (float)(clock64() - ts[0]);
                t += NPBUF;
                if (t >= tiles) { t -= tiles; l++; }
            }
            umma_commit(BAR(BAR_WEMPTY + cur_l % NWBUF));
        }
        __syncwarp();
    } else if (warp == PROD_WARP) {
        // ---- weight producer --------------------------------------------------------
        if (elect_one_sync()) {
            mbar_expect_tx(BAR(BAR_HEAD), (uint32_t)(NOUT * RS_H * 2));
            bulk_g2s(smem_u32(wh), whead, (uint32_t)(NOUT * RS_H * 2), BAR(BAR_HEAD));
#pragma unroll 1
            for (int l = 0; l < layers; l++) {
                const int wbuf = l % NWBUF, use = l / NWBUF;
                if (use > 0) mbar_wait(BAR(BAR_WEMPTY + wbuf), (uint32_t)((use - 1) & 1));
                mbar_expect_tx(BAR(BAR_WFULL + wbuf), WL_BYTES);
                bulk_g2s(wb_s + (uint32_t)(wbuf * WL_BYTES), wconv + (size_t)l * WL_BYTES, WL_BYTES, BAR(BAR_WFULL + wbuf));
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue: group `grp` (8 warps) drains ring slot `grp`; a warp owns one TMEM lane quadrant and 16 channels
        const int e = warp - EPI_WARP0, grp = e / GRP_WARPS, q = warp & 3, half = (e % GRP_WARPS) >> 2;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t t_p = tmem_base + lane_off + COL_P + (uint32_t)(grp * NACC + 16 * half);
        const uint32_t t_x0 = tmem_base + lane_off + COL_X + (uint32_t)(16 * half);
        const uint32_t bar_full = BAR(BAR_PFULL + grp), bar_empty = BAR(BAR_PEMPTY + grp);
        const int total = layers * tiles, ho = 16 * half, r0 = q * 32 + lane, live_rows = nb * FB;
        unsigned char *const fa_h = fa + (size_t)(2 * half) * PLANE + (size_t)PADR * 16;
        unsigned char *const fb_h = fb + (size_t)(2 * half) * PLANE + (size_t)PADR * 16;
        int l = 0, t = grp;
        uint32_t par = 0;
#pragma unroll 1
        for (int g = grp; g < total; g += NPBUF) {
            const bool is_c1 = (l & 1) == 1, last = l + 1 == layers;
            // one warp of the group polls the mbarrier, the other seven sleep in a hardware barrier: a third of all
            // instructions this kernel issued were wait-loop iterations of epilogue warps, competing for issue slots
            // with the warps that had work
            const bool trace = DBG && timing && blockIdx.x == 0 && lane == 0;
            float *tl = trace ? dump + (size_t)1024 * 256 + (size_t)g * 16 : nullptr;
            // everything that does not need the accumulator is computed before the wait: once the tile is complete, the
            // ring slot is held until every warp has its loads, so the instructions in front of them are on the
            // critical path of the tensor pipe
            const int fr = t * 128 + r0;
            const bool live = ((fr & 7) != 7) && (((fr >> 3) % (BH + 1)) != BH) && fr < live_rows;
            const uint32_t t_x = t_x0 + (uint32_t)(CH * t);
            // the last layer writes the head input: [8-channel chunk][board][HB rows][16 B]
            unsigned char *out_row;
            int out_plane = PLANE;
            if (last) {
                const int brd = fr / FB;
                out_row = fa + ((size_t)(2 * half) * (NB * HB) + (size_t)(fr + brd)) * 16;      // brd*HB + fr - brd*FB
                out_plane = NB * HB * 16;
            } else {
                out_row = (is_c1 ? fb_h : fa_h) + (size_t)fr * 16;
            }
            float *dmp = (DBG && dump != nullptr && l == dump_layer && fr < live_rows) ? dump + ((size_t)board0 * FB + fr) * CH + ho : nullptr;
            if (trace && (e % GRP_WARPS) == 0) tl[5] = (float)(clock64() - ts[0]);
            if ((e % GRP_WARPS) == 0) mbar_wait<32>(bar_full, par);
            if (trace && (e % GRP_WARPS) == 0) tl[6] = (float)(clock64() - ts[0]);
            asm volatile("bar.sync %0, %1;\n" ::"r"(1 + grp), "r"(GRP_WARPS * 32) : "memory");
            par ^= 1u;
            tc_fence_after();
            if (trace && (e % GRP_WARPS) == 7) tl[7] = (float)(clock64() - ts[0]);
            if (l == 0) {
                epilogue_tile<EPI_STEM, DBG>(t_p, t_x, out_row, out_plane, s_bias + ho, depth > 0 ? s_sc + ho : nullptr, s_sh + ho,
                                             live, lane, bar_empty, dmp);
            } else if (is_c1) {
                epilogue_tile<EPI_CONV1, DBG>(t_p, t_x, out_row, out_plane, s_bias + l * CH + ho, nullptr, nullptr, live, lane,
                                              bar_empty, dmp);
            } else {
                const int nb = l >> 1;                        // the block that consumes x next
                epilogue_tile<EPI_CONV2, DBG>(t_p, t_x, out_row, out_plane, nullptr, last ? nullptr : s_sc + nb * CH + ho,
                                              s_sh + nb * CH + ho, live, lane, bar_empty, dmp);
            }
            fence_proxy_async();              // the next layer's MMAs read these rows through the async proxy
            __syncwarp();
            if (lane == 0) {                  // every tile whose MMAs read these rows
                if (t > 0) mbar_arrive(BAR(BAR_READY + t - 1));
                mbar_arrive(BAR(BAR_READY + t));
                if (t + 1 < tiles) mbar_arrive(BAR(BAR_READY + t + 1));
            }
            if (trace) tl[8 + (e % GRP_WARPS)] = (float)(clock64() - ts[0]);
            t += NPBUF;
            if (t >= tiles) { t -= tiles; l++; }
        }
    }
#undef CNT
    if (timing) ts[2] = clock64();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (timing) ts[3] = clock64();

    // ---- heads: logits[board][j] = sum_k wh[j][k] * a[board][k] + bhead[j], k = frame row * 32 + channel ----------
    mbar_wait(BAR(BAR_HEAD), 0u);
    {
        const uint32_t h_s = smem_u32(wh);
        float acc[2][4];
#pragma unroll
        for (int i = 0; i < 2; i++) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f; }
        const int brd = lane & 15, khalf = lane >> 4;
        const uint32_t a_lane = fa_s + (uint32_t)((brd * HB) * 16 + khalf * (NB * HB * 16));
        int hrow = ((lane >> 4) << 3) + (lane & 7);                     // output row this lane addresses for ldmatrix
        hrow = hrow < NOUT ? hrow : NOUT - 1;                           // rows >= NOUT: any valid row, result unused
        const uint32_t h_lane = h_s + (uint32_t)((hrow * RS_H + ((lane >> 3) & 1) * 8) * 2);
        for (int ks = warp; ks < KH / 16; ks += WARPS) {
            const int fr = ks >> 1, half = ks & 1;
            uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
            ldmatrix_x4(a0, a1, a2, a3, a_lane + (uint32_t)(fr * 16 + 2 * half * (NB * HB * 16)));
            ldmatrix_x4(b0, b1, b2, b3, h_lane + (uint32_t)(ks * 32));
            mma_bf16(acc[0], a0, a1, a2, a3, b0, b1);
            mma_bf16(acc[1], a0, a1, a2, a3, b2, b3);
        }
        const int gq = lane >> 2, tq = lane & 3;                   // c0,c1: board gq; c2,c3: board gq+8; cols 2tq, 2tq+1
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            red[(warp * NB + gq) * 16 + nt * 8 + 2 * tq] = acc[nt][0];
            red[(warp * NB + gq) * 16 + nt * 8 + 2 * tq + 1] = acc[nt][1];
            red[(warp * NB + gq + 8) * 16 + nt * 8 + 2 * tq] = acc[nt][2];
            red[(warp * NB + gq + 8) * 16 + nt * 8 + 2 * tq + 1] = acc[nt][3];
        }
        __syncthreads();
        if (tid < NB * 16) {
            float v = 0.0f;
            for (int wq = 0; wq < WARPS; wq++) v += red[wq * NB * 16 + tid];
            fin[tid] = v + ((tid & 15) < NOUT ? bhead[tid & 15] : 0.0f);
        }
        __syncthreads();
        if (tid < nb && board0 + tid < B) {
            const int gb = rows != nullptr ? rows[board0 + tid] : board0 + tid;
            constexpr int A = NOUT - 3;
            float lg[NOUT];
#pragma unroll
            for (int j = 0; j < NOUT; j++) lg[j] = fin[tid * 16 + j];
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
    if (timing && lane == 0) {        // per warp: prologue, own trunk work, wait for the slowest warp, heads; + SM id
        ts[4] = clock64();
        float *o = dump + ((size_t)blockIdx.x * 32 + warp) * 8;
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        o[0] = (float)(ts[1] - ts[0]); o[1] = (float)(ts[2] - ts[1]); o[2] = (float)(ts[3] - ts[2]); o[3] = (float)(ts[4] - ts[3]);
        o[4] = (float)smid;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
#undef BAR
}

}  // namespace tc
}  // namespace

extern "C" int azb_nn_tc_layer_bytes(void) { return tc::WL_BYTES; }
extern "C" int azb_nn_tc_head_row_stride(void) { return tc::RS_H; }
extern "C" int azb_nn_tc_boards_per_cta(void) { return tc::NB; }
extern "C" int azb_nn_tc_frame_rows_per_board(void) { return tc::FB; }

static int tc_launch(const azb_nn_weights *w, const float *obs, float *policy, float *value, int32_t batch, void *stream,
                     float *dump, int dump_layer, const int *rows = nullptr, const int *count = nullptr)
{
    if (!w || !obs || !policy || !value || batch <= 0) return -7;
    if (w->channels != tc::CH || w->board_h != tc::BH || w->board_w != tc::BW || w->action_size != 7 || w->in_channels > 8 ||
        w->in_channels < 1 || w->depth < 0 || w->depth > tc::MAXD)
        return -1;
    cudaStream_t s = (cudaStream_t)stream;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(tc::k_resnet_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(tc::k_resnet_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES) != cudaSuccess)
            return -2;
        configured = true;
    }
    // wave shaping: `rem` = tiles of 16 boards in the partially filled last wave; when twice as many CTAs still fit in
    // one wave, they run as half tiles of 8 boards (4 M-tiles instead of 7 on the critical path of the launch)
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaDeviceProp prop;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return -2;
        sms = prop.multiProcessorCount;
    }
    const int T = (batch + tc::NB - 1) / tc::NB;
    const int rem = T % sms;
    int n_full = T, n_half = 0;
    if (rem > 0 && 2 * rem <= sms) { n_full = T - rem; n_half = 2 * rem; }
    int grid = n_full + n_half;
    if (rows != nullptr) grid = T + sms / 2;       // upper bound over every batch size <= batch; surplus CTAs exit at once
    if (dump != nullptr)
        tc::k_resnet_tc<true><<<grid, tc::THREADS, tc::SMEM_BYTES, s>>>(
            obs, policy, value, batch, w->in_channels, w->depth, reinterpret_cast<const unsigned char *>(w->wconv), w->cbias,
            w->bn_scale, w->bn_shift, reinterpret_cast<const __nv_bfloat16 *>(w->whead), w->bhead, dump, dump_layer, n_full, rows, count, sms);
    else
        tc::k_resnet_tc<false><<<grid, tc::THREADS, tc::SMEM_BYTES, s>>>(
            obs, policy, value, batch, w->in_channels, w->depth, reinterpret_cast<const unsigned char *>(w->wconv), w->cbias,
            w->bn_scale, w->bn_shift, reinterpret_cast<const __nv_bfloat16 *>(w->whead), w->bhead, nullptr, -1, n_full, rows, count, sms);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

extern "C" int azb_nn_forward_tc(const azb_nn_weights *w, const float *obs, float *policy, float *value, int32_t batch,
                                 void *stream)
{
    return tc_launch(w, obs, policy, value, batch, stream, nullptr, -1);
}

extern "C" int azb_nn_forward_tc_rows(const azb_nn_weights *w, const float *obs, float *policy, float *value, const int32_t *rows,
                                      const int32_t *count, int32_t max_batch, void *stream)
{
    if (!rows || !count) return -7;
    return tc_launch(w, obs, policy, value, max_batch, stream, nullptr, -1, rows, count);
}

extern "C" int azb_nn_forward_tc_debug(const azb_nn_weights *w, const float *obs, float *policy, float *value, int32_t batch,
                                       void *stream, float *dump, int32_t dump_layer)
{
    return tc_launch(w, obs, policy, value, batch, stream, dump, dump_layer);
}
