// azb_resnet_g.cu -- the reference's pre-activation ResNet (alphazero/NNetArchitecture.py:69-120, the network
// NNetWrapper.process evaluates, alphazero/NNetWrapper.py:225-232) on the 5th-generation tensor cores for every
// shipped geometry: boards up to 7x7, 32 or 64 trunk channels, any action size -- and at the reference's numerics.
//
// Precision.  PREC_BF16X2 (the default) keeps every convolution / head operand as a pair of bf16 values hi + lo
// (16 significant bits; TF32, what the reference gets from cuDNN, has 11) and evaluates a.w as
// a_hi.w_hi + a_hi.w_lo + a_lo.w_hi with fp32 accumulation in TMEM: three tcgen05.mma per K step, probabilities within
// 1e-5 of the fp32 module (tests/test_nn_tc.py states and checks the bound).  PREC_F16X2 is the same scheme with fp16
// pairs (22 significant bits while the low part stays a normal fp16 number, i.e. for |v| >= 0.03; below that the absolute
// error is the subnormal spacing 6e-8; |v| saturates at 65504): the default of the 128-channel network, whose 17 layers of
// K = 1152 leave bf16x2 at 1.07e-5 with sharpened heads.  PREC_F16 (one pass, 11-bit significand = TF32's) and PREC_BF16
// (one pass, 8 bits) are opt-in performance modes.
//
// Trunk kernel (k_trunk_tc).  A board is an 8 x 8 frame of positions (live H x W corner, the rest zero): one zero
// column and >= one zero row per board are all the padding a 3x3 convolution needs, a tap (dy, dx) is a shift of
// 8 dy + dx frame rows, and an M = 128 tile is exactly two boards, so tiles never exchange data: a tile's
// activation is rewritten IN PLACE by its own epilogue (one frame instead of two: this is what lets the split operands
// fit in shared memory), and tile t of layer l+1 waits only for tile t of layer l.  Frame layout in shared memory:
// [part hi|lo][8-channel chunk][frame row][16 B] = the canonical K-major no-swizzle UMMA operand (core matrix = 8 rows,
// SBO = 128 B, LBO = plane size); a vertical tap is `start address += 128 dy`.  The three horizontal taps share one A
// read: B holds [dx][cout] (N = 3 CH) and the epilogue forms D[r] = P_-1[r-1] + P_0[r] + P_+1[r+1] with one-lane
// rotations (rows that wrap are padding columns, where P is exactly zero).  TMEM holds the fp32 residual stream
// (CH columns per tile) and a ring of N = 3 CH accumulators.  Weights stream through a ring of slabs (1 or 3 vertical
// taps each, cp.async.bulk) so that a 64-channel layer (144 KB of split operands) fits beside the activations.
// Roles: up to three MMA-issuing warps (one elected thread each; a turn counter keeps a tile's MMAs contiguous), one
// weight producer, NGROUPS epilogue groups of 4 x CH/16 warps (TMEM lane quadrant x 16-channel group).
// The last layer's epilogue writes the trunk output straight into the head GEMM's operand layout in global memory.
//
// Head kernel (k_head_tc).  The heads (conv1x1 -> BN -> flatten -> Linear .. Linear with Identity activations,
// NNetArchitecture.py:86-102,110-118) are one affine map of the trunk output, folded on the host in float64
// (azb200/fused_nn.py): logits[B, A+3] = act[B, H W CH] . Wh^T + b as a tcgen05 GEMM with M = 128 boards per CTA,
// bulk-copy staged operands (4-stage ring), same split arithmetic; softmax in fp32 (in the epilogue when all outputs fit
// one N tile, else k_softmax over the logits).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "azb_tc_ptx.cuh"
#include "../../include/azb200_nn.h"

namespace {
namespace g {
using namespace azbtc;

constexpr int MAXD = 6, MAXL = 1 + 2 * MAXD;
__host__ __device__ constexpr bool prec_split(int prec) { return prec == AZB_NN_BF16X2 || prec == AZB_NN_F16X2; }   // operands as hi + lo
__host__ __device__ constexpr bool prec_f16(int prec) { return prec == AZB_NN_F16 || prec == AZB_NN_F16X2; }        // element type fp16

template <int CH_, int TILES_, int PREC_, int NGROUPS_, int DYS_, int NSLOT_, bool PAIR_ = false, int NISSUE_ = 0>
struct TrunkCfg {
    static constexpr int CH = CH_, TILES = TILES_, PREC = PREC_, NGROUPS = NGROUPS_, DYS = DYS_, NSLOT = NSLOT_;
    // PAIR: two CTAs of a cluster run their tiles in lockstep as ONE M = 256 MMA stream (cta_group::2) issued by the
    // leader; each CTA stages only half of the N rows of every weight chunk (NB), which takes the B operand's share of
    // the shared-memory wavefronts per MMA from 3 CH / 4 to 3 CH / 8 -- the pipe both trunk configurations are bound by
    static constexpr bool PAIR = PAIR_;
    static constexpr int PARTS = prec_split(PREC) ? 2 : 1;
    static constexpr bool F16 = prec_f16(PREC);
    static constexpr int C8 = CH / 8, KST = CH / 16;
    static constexpr int ROWS = TILES * 128, PADR = 16, FROWS = ROWS + 2 * PADR;
    static constexpr int PLANE = FROWS * 16, FPART = C8 * PLANE, FRAME = PARTS * FPART;
    static constexpr int NACC = 3 * CH, NB = PAIR ? NACC / 2 : NACC, WCHUNK = NB * 16;       // B rows staged by this CTA
    static constexpr int SLABS = 3 / DYS;                               // weight slabs per trunk layer
    static constexpr int SLAB_PART = DYS * C8 * WCHUNK, SLAB = PARTS * SLAB_PART;
    static constexpr int STEM_PART = 4 * WCHUNK;                        // stem slab: [part][4 K chunks][NACC][8]
    static constexpr int NRING = (512 - TILES * CH) / NACC >= 3 ? 3 : (512 - TILES * CH) / NACC;
    static constexpr int NISSUE = NISSUE_ > 0 ? NISSUE_ : (TILES < 3 ? TILES : 3);      // MMA-issuing threads
    static constexpr int GW = 4 * (CH / 16);                            // warps of one epilogue group
    static constexpr int EPI_WARP0 = 4;                                 // warps 0..2 issue, warp 3 produces
    static constexpr int WARPS = EPI_WARP0 + NGROUPS * GW, THREADS = WARPS * 32;
    static constexpr uint32_t COL_X = 0, COL_P = TILES * CH;
    static constexpr int PRM_FLOATS = MAXL * CH + 2 * MAXD * CH;
    static constexpr size_t SMEM = (size_t)FRAME + (size_t)NSLOT * SLAB + 40 * 8 + 64;
    static constexpr uint32_t IDESC = umma_idesc(NACC, F16, PAIR ? 256 : 128);
    static_assert(3 % DYS == 0 && STEM_PART <= SLAB_PART && NSLOT >= SLABS, "slab geometry");
    static_assert(!PAIR || NACC % 16 == 0, "an M = 256 MMA needs N % 16 == 0");
    static_assert(NRING >= 2 && COL_P + NRING * NACC <= 512, "tensor memory budget");
    // Every barrier wait names a phase by its parity only, so when a waiter reaches a wait, all earlier phases of that barrier
    // must be known to be complete.  The configurations below satisfy it (an issuer always meets the same ring slot, or the
    // in-order MMA stream implies the phases a group skipped); 8 tiles of 32 channels = a ring of two under three issuers
    // does not (issuer 0 would wait for the third drain of slot 0 having seen only the first) and trapped in the bounded
    // wait; with two issuers and two groups it runs, at 183.7 instead of 162.3 us per 6960 boards.
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    static_assert(NACC <= 256 && NACC % 16 == 0 && WARPS <= 32, "shape limits");
};

enum { EPI_STEM = 0, EPI_CONV1 = 1, EPI_CONV2 = 2 };

// DBG kernels only (azb_nng_forward_debug): clock64 of CTA 0's pipeline events per (layer, tile) -- [0] the tile's MMAs start
// to issue, [1] they are committed, [2] its epilogue sees the accumulator, [3] the epilogue hands the tile on
__device__ long long g_trace[4 * 256];
// DBG kernels only: (SM id, %globaltimer at entry, at exit) of every CTA with blockIdx.x < 2048
__device__ long long g_cta_trace[3 * 2048];
__device__ __forceinline__ long long global_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t)); return t; }
__device__ __forceinline__ int sm_id() { int v; asm volatile("mov.u32 %0, %%smid;\n" : "=r"(v)); return v; }

__device__ __forceinline__ float2 f2(uint32_t lo, uint32_t hi) { return make_float2(__uint_as_float(lo), __uint_as_float(hi)); }
__device__ __forceinline__ uint32_t bf2_bits(__nv_bfloat162 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ uint32_t f16x2_sat(float2 v)      // round to nearest, clamp to the largest finite fp16
{
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(v.y), "f"(v.x));
    return r;
}

// 16 fp32 values (8 pairs) -> operand format, stored as two 16-byte K chunks per part
template <class C>
__device__ __forceinline__ void store_operand(const float2 (&r)[8], unsigned char *dst, size_t part_stride, size_t chunk_stride,
                                              bool live, float *dump_row)
{
    uint32_t o[8], o2[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        if (C::F16) {
            o[c] = f16x2_sat(r[c]);
            if (C::PARTS == 2) {
                const float2 hf = __half22float2(*reinterpret_cast<const __half2 *>(&o[c]));
                o2[c] = f16x2_sat(make_float2(__fsub_rn(r[c].x, hf.x), __fsub_rn(r[c].y, hf.y)));
            }
        } else {
            const __nv_bfloat162 h = __float22bfloat162_rn(r[c]);
            o[c] = bf2_bits(h);
            if (C::PARTS == 2) {
                const float2 hf = __bfloat1622float2(h);
                o2[c] = bf2_bits(__float22bfloat162_rn(make_float2(__fsub_rn(r[c].x, hf.x), __fsub_rn(r[c].y, hf.y))));
            }
        }
    }
    if (dump_row != nullptr) {                 // test hook: the value the next layer's MMAs see
#pragma unroll
        for (int c = 0; c < 8; c++) {
            float2 v;
            if (C::F16) {
                v = __half22float2(*reinterpret_cast<__half2 *>(&o[c]));
                if (C::PARTS == 2) {
                    const float2 w = __half22float2(*reinterpret_cast<__half2 *>(&o2[c]));
                    v.x += w.x; v.y += w.y;
                }
            } else {
                v = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&o[c]));
                if (C::PARTS == 2) {
                    const float2 w = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&o2[c]));
                    v.x += w.x; v.y += w.y;
                }
            }
            dump_row[2 * c] = live ? v.x : 0.0f;
            dump_row[2 * c + 1] = live ? v.y : 0.0f;
        }
    }
    if (live) {                                // padding rows are zero from the start and stay so
        *reinterpret_cast<uint4 *>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4 *>(dst + chunk_stride) = make_uint4(o[4], o[5], o[6], o[7]);
        if (C::PARTS == 2) {
            *reinterpret_cast<uint4 *>(dst + part_stride) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
            *reinterpret_cast<uint4 *>(dst + part_stride + chunk_stride) = make_uint4(o2[4], o2[5], o2[6], o2[7]);
        }
    }
}

// Per-channel parameters of the epilogues (folded-BN bias of stem / conv1, BN1 scale and shift of every block).  The
// 32 / 64-channel kernel takes them BY VALUE as a kernel argument (constant bank): ncu's per-instruction accounting showed
// the float4 parameter loads from shared memory (12 LDS.128 per warp, tile and layer, two wavefronts each) to be 22 % of the
// LSU wavefronts on the very pipe the MMA operand reads saturate; LDC goes through the constant cache instead.
template <int CH>
struct NetParams {
    float bias[MAXL * CH], sc[MAXD * CH], sh[MAXD * CH];
};
template <int CH>
struct ConstPrm {                              // view of a __grid_constant__ NetParams
    const NetParams<CH> &p;
    __device__ __forceinline__ float4 bias(int off, int c) const { return make_float4(p.bias[off + 4 * c], p.bias[off + 4 * c + 1], p.bias[off + 4 * c + 2], p.bias[off + 4 * c + 3]); }
    __device__ __forceinline__ float4 scale(int off, int c) const { return make_float4(p.sc[off + 4 * c], p.sc[off + 4 * c + 1], p.sc[off + 4 * c + 2], p.sc[off + 4 * c + 3]); }
    __device__ __forceinline__ float4 shift(int off, int c) const { return make_float4(p.sh[off + 4 * c], p.sh[off + 4 * c + 1], p.sh[off + 4 * c + 2], p.sh[off + 4 * c + 3]); }
};
struct SmemPrm {                               // the three arrays staged in shared memory (128-channel kernel)
    const float *b, *s, *h;
    __device__ __forceinline__ float4 bias(int off, int c) const { return reinterpret_cast<const float4 *>(b + off)[c]; }
    __device__ __forceinline__ float4 scale(int off, int c) const { return reinterpret_cast<const float4 *>(s + off)[c]; }
    __device__ __forceinline__ float4 shift(int off, int c) const { return reinterpret_cast<const float4 *>(h + off)[c]; }
};

// The part of an epilogue that follows the accumulator gather: r = the convolution output of this thread's row for 16
// channels, xv = the residual stream's 16 values (EPI_CONV2 only; rewritten for stem / conv2).
// boff: offset of this thread's 16 bias values (stem / conv1); soff: of its 16 BN1 scale / shift values, < 0: no BN + ReLU
// after the residual update (the last block hands x itself to the heads; depth 0: the stem does)
// KEEP_X = false (k_trunk_tc): the LAST layer (soff < 0: its output goes to the heads) does not write x back to tensor
// memory -- nobody reads it, and in a persistent CTA the next round's stem epilogue of the same tile may already be running
template <class C, int EPI, bool DBG, bool WAIT_ST = true, bool KEEP_X = true, class PV>
__device__ __forceinline__ void epilogue_finish(float2 (&r)[8], uint32_t (&xv)[16], uint32_t t_x, unsigned char *dst, size_t part_stride,
                                                size_t chunk_stride, const PV &pv, int boff, int soff, bool live, float *dump_row)
{
    if (EPI != EPI_CONV2) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float4 b4 = pv.bias(boff, c);
            r[2 * c] = __fadd2_rn(r[2 * c], make_float2(b4.x, b4.y));
            r[2 * c + 1] = __fadd2_rn(r[2 * c + 1], make_float2(b4.z, b4.w));
        }
    }
    if (EPI == EPI_STEM) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            r[c].x = fmaxf(r[c].x, 0.0f); r[c].y = fmaxf(r[c].y, 0.0f);
            xv[2 * c] = __float_as_uint(r[c].x); xv[2 * c + 1] = __float_as_uint(r[c].y);
        }
        if (KEEP_X || soff >= 0) tmem_st16(t_x, xv);
    } else if (EPI == EPI_CONV2) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            r[c] = __fadd2_rn(r[c], f2(xv[2 * c], xv[2 * c + 1]));
            xv[2 * c] = __float_as_uint(r[c].x); xv[2 * c + 1] = __float_as_uint(r[c].y);
        }
        if (KEEP_X || soff >= 0) tmem_st16(t_x, xv);
    }
    if (EPI == EPI_CONV1 || soff >= 0) {
        if (EPI != EPI_CONV1) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float4 s4 = pv.scale(soff, c), h4 = pv.shift(soff, c);
                r[2 * c] = __ffma2_rn(r[2 * c], make_float2(s4.x, s4.y), make_float2(h4.x, h4.y));
                r[2 * c + 1] = __ffma2_rn(r[2 * c + 1], make_float2(s4.z, s4.w), make_float2(h4.z, h4.w));
            }
        }
#pragma unroll
        for (int c = 0; c < 8; c++) { r[c].x = fmaxf(r[c].x, 0.0f); r[c].y = fmaxf(r[c].y, 0.0f); }
    }
    store_operand<C>(r, dst, part_stride, chunk_stride, live, DBG ? dump_row : nullptr);
    if (WAIT_ST && EPI != EPI_CONV1) tmem_st_wait();
}

// Epilogue of one tile for 16 of its channels: this thread owns tile row q*32 + lane (= its TMEM lane) and the channels
// [16 cg, 16 cg + 16).
//   stem : x = relu(D + bias) -> TMEM;  a = relu(bn1_0(x))      (depth 0: a = x)
//   conv1: b = relu(D + bias)                                   (BN2 folded into conv1)
//   conv2: x += D -> TMEM;  a = relu(bn1_next(x))               (last block: a = x, the head input)
template <class C, int EPI, bool DBG, class PV>
__device__ __forceinline__ void epilogue_tile(uint32_t t_p, uint32_t t_x, unsigned char *dst, size_t part_stride, size_t chunk_stride,
                                              const PV &pv, int boff, int soff, bool live, int lane, uint32_t bar_pempty, uint32_t bar_ready,
                                              float *dump_row)
{
    uint32_t pm[16], p0[16], pp[16], xv[16];
    tmem_ld16(t_p, pm);
    tmem_ld16(t_p + (uint32_t)C::CH, p0);
    tmem_ld16(t_p + (uint32_t)(2 * C::CH), pp);
    if (EPI == EPI_CONV2) tmem_ld16(t_x, xv);
    tmem_ld_wait();
    tc_fence_before();                    // the accumulator is in registers: hand the ring slot back to the MMA warps
    // bar_ready != 0 (last layer of a round that is not the last): the tile itself is handed on here, too -- x has been read,
    // the next round's observations are in the frame (staged by the caller; this fence publishes them to the async proxy),
    // and what is left of this epilogue touches neither the frame nor tensor memory
    if (bar_ready != 0u) fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        if (C::PAIR) mbar_arrive_cluster(bar_pempty); else mbar_arrive(bar_pempty);
        if (bar_ready != 0u) { if (C::PAIR) mbar_arrive_cluster(bar_ready); else mbar_arrive(bar_ready); }
    }
    const int src_up = (lane + 31) & 31, src_dn = (lane + 1) & 31;
    float2 r[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        float2 up, dn;
        if (C::PREC == AZB_NN_BF16) {     // neighbour terms as packed fp16 pairs: far below the bf16 operand rounding
            const __half2 hm = __floats2half2_rn(__uint_as_float(pm[2 * c]), __uint_as_float(pm[2 * c + 1]));
            const __half2 hp = __floats2half2_rn(__uint_as_float(pp[2 * c]), __uint_as_float(pp[2 * c + 1]));
            const uint32_t um = __shfl_sync(0xffffffffu, h2_bits(hm), src_up);
            const uint32_t ud = __shfl_sync(0xffffffffu, h2_bits(hp), src_dn);
            up = __half22float2(*reinterpret_cast<const __half2 *>(&um));
            dn = __half22float2(*reinterpret_cast<const __half2 *>(&ud));
        } else {
            up = f2(__shfl_sync(0xffffffffu, pm[2 * c], src_up), __shfl_sync(0xffffffffu, pm[2 * c + 1], src_up));
            dn = f2(__shfl_sync(0xffffffffu, pp[2 * c], src_dn), __shfl_sync(0xffffffffu, pp[2 * c + 1], src_dn));
        }
        r[c] = __fadd2_rn(__fadd2_rn(up, f2(p0[2 * c], p0[2 * c + 1])), dn);
    }
    epilogue_finish<C, EPI, DBG, true, false>(r, xv, t_x, dst, part_stride, chunk_stride, pv, boff, soff, live, dump_row);
}

// MMAs of one (tile, slab).  Trunk slab: [part][dy in slab][K chunk][NACC][8]; per 16-channel K step the passes
// hi.hi (+ hi.lo + lo.hi when the operands are split).  Stem slab: [part][4 K chunks][NACC][8], two K steps over chunk
// plane 0: (dy=-1, dy=0) and (dy=+1, zero weights) -- the second K chunk of a step is the first one 8 rows further.
// turn_bar != 0: the barrier the NEXT tile's issuer waits on; it is arrived on before this slab's last vertical tap, so that
// that thread's wake-up (~300 cycles, measured as the gap between two tiles' MMAs) passes under the MMAs still to issue.
// The two tiles' MMAs may interleave in the pipe for a few instructions; they touch different accumulators.
template <class C>
__device__ __forceinline__ void issue_slab(uint32_t frame_s, uint32_t w_s, uint32_t d_tmem, int t, int layer, int j, uint32_t turn_bar = 0u)
{
    const uint32_t row0 = frame_s + (uint32_t)((C::PADR + 128 * t) * 16);
    if (layer == 0) {
        if (turn_bar != 0u) mbar_arrive_relaxed(turn_bar);
#pragma unroll
        for (int s = 0; s < 2; s++) {
#pragma unroll
            for (int pass = 0; pass < (C::PARTS == 2 ? 3 : 1); pass++) {
                const int pa = pass == 2 ? 1 : 0, pb = pass == 1 ? 1 : 0;
                const uint64_t ad = umma_desc(row0 + (uint32_t)(pa * C::FPART + (2 * s - 1) * 128), 128u, 128u);
                const uint64_t bd = umma_desc(w_s + (uint32_t)(pb * C::STEM_PART + 2 * s * C::WCHUNK), (uint32_t)C::WCHUNK, 128u);
                if (C::PAIR) umma_f16_pair(d_tmem, ad, bd, C::IDESC, (s > 0 || pass > 0) ? 1u : 0u);
                else umma_f16(d_tmem, ad, bd, C::IDESC, (s > 0 || pass > 0) ? 1u : 0u);
            }
        }
    } else {
#pragma unroll
        for (int dl = 0; dl < C::DYS; dl++) {
            const int dy = j * C::DYS + dl - 1;
            if (turn_bar != 0u && dl == C::DYS - 1) mbar_arrive_relaxed(turn_bar);
#pragma unroll
            for (int ks = 0; ks < C::KST; ks++) {
#pragma unroll
                for (int pass = 0; pass < (C::PARTS == 2 ? 3 : 1); pass++) {
                    const int pa = pass == 2 ? 1 : 0, pb = pass == 1 ? 1 : 0;
                    const uint64_t ad = umma_desc(row0 + (uint32_t)(pa * C::FPART + 2 * ks * C::PLANE + dy * 128), (uint32_t)C::PLANE, 128u);
                    const uint64_t bd = umma_desc(w_s + (uint32_t)(pb * C::SLAB_PART + (dl * C::C8 + 2 * ks) * C::WCHUNK),
                                                  (uint32_t)C::WCHUNK, 128u);
                    if (C::PAIR) umma_f16_pair(d_tmem, ad, bd, C::IDESC, (j > 0 || dl > 0 || ks > 0 || pass > 0) ? 1u : 0u);
                    else umma_f16(d_tmem, ad, bd, C::IDESC, (j > 0 || dl > 0 || ks > 0 || pass > 0) ? 1u : 0u);
                }
            }
        }
    }
}

// how the global list of tiles (two boards each) is dealt to CTAs: whole waves of `sms` CTAs with equal shares, so a
// partially filled last wave costs max-tiles-per-CTA, not a full CTA
struct TileShare { int nct, base, extra; };
__host__ __device__ inline TileShare tile_share(int boards, int tiles_per_cta, int sms)
{
    TileShare s;
    const int T2 = (boards + 1) / 2;
    const int waves = (T2 + tiles_per_cta * sms - 1) / (tiles_per_cta * sms);
    s.nct = waves * sms < T2 ? waves * sms : T2;
    s.base = s.nct > 0 ? T2 / s.nct : 0;
    s.extra = s.nct > 0 ? T2 % s.nct : 0;
    return s;
}

// k_trunk_tc: how the tiles (pair mode: pairs of tiles) are dealt to units (CTAs, pair mode: CTA pairs).  persist: ONE wave
// of resident units with equal shares, a unit running its share in `rounds` rounds: the first `big` rounds of `tiles`
// tiles, the others of tiles - 1 -- exactly its share -- as long as the smaller rounds still give every MMA-issuing
// thread a tile (tiles - 1 >= issuers); else all rounds of `tiles` tiles with at most rounds - 1 empty ones.  Otherwise
// whole waves of units with one round each (tile_share).
struct UnitShare { int units, n, first, rounds, tiles, big; };
__host__ __device__ inline UnitShare unit_share(int boards, int tiles_per_cta, int sms, bool pair, bool persist, int unit, int issuers)
{
    UnitShare u;
    const int T2 = (boards + 1) / 2, items = pair ? (T2 + 1) / 2 : T2, resident = pair ? sms / 2 : sms;
    if (persist) u.units = items < resident ? items : resident;
    else {
        const int waves = (items + tiles_per_cta * resident - 1) / (tiles_per_cta * resident);
        u.units = waves * resident < items ? waves * resident : items;
    }
    u.n = 0; u.first = 0; u.rounds = 0; u.tiles = 0; u.big = 0;
    if (unit < 0 || unit >= u.units) return u;
    const int base = items / u.units, extra = items % u.units;
    u.n = base + (unit < extra ? 1 : 0);
    u.first = unit * base + (unit < extra ? unit : extra);
    u.rounds = (u.n + tiles_per_cta - 1) / tiles_per_cta;
    u.tiles = (u.n + u.rounds - 1) / u.rounds;
    u.big = u.tiles - 1 >= issuers ? u.n - u.rounds * (u.tiles - 1) : u.rounds;
    return u;
}

template <class C, bool DBG>
__global__ void __launch_bounds__(C::THREADS, 1)
k_trunk_tc(const float *__restrict__ obs, int B, int in_ch, int H, int W, int depth, const unsigned char *__restrict__ wtrunk,
           const __grid_constant__ NetParams<C::CH> P, unsigned char *__restrict__ gact, int MT, int KC, float *__restrict__ dump, int dump_layer,
           const int *__restrict__ rows, const int *__restrict__ count_ptr, int sms, int persist)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *frame = smem;
    unsigned char *wb = frame + C::FRAME;                                       // weight slab ring
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(wb + (size_t)C::NSLOT * C::SLAB);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 40);
    uint32_t *cnts = tmem_slot + 4;

    // pair mode: PEMPTY / READY / TURN / WFULLP are the LEADER's (rank 0; both CTAs' epilogue warps arrive there, the
    // peer's weight relay arrives on WFULLP); PFULL / WEMPTY exist in both CTAs and are signalled by multicast commits
    enum { BAR_PFULL = 0, BAR_PEMPTY = BAR_PFULL + C::NRING, BAR_READY = BAR_PEMPTY + C::NRING,
           BAR_WFULL = BAR_READY + C::TILES, BAR_WEMPTY = BAR_WFULL + C::NSLOT, BAR_TURN = BAR_WEMPTY + C::NSLOT,
           BAR_WFULLP = BAR_TURN + 3, NBARS = BAR_WFULLP + C::NSLOT };
    static_assert(NBARS <= 32, "one barrier per lane of warp 0");

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (DBG && blockIdx.x == 0 && tid == 0) g_trace[1000] = clock64();          // CTA entry
    if (DBG && tid == 0 && blockIdx.x < 2048) { g_cta_trace[3 * blockIdx.x] = sm_id(); g_cta_trace[3 * blockIdx.x + 1] = global_ns(); }
    int count_now = B;
    if (rows != nullptr) count_now = *count_ptr;         // compact mode: the batch size lives in device memory
    {   // zero the frame (padding rows / columns must read as zero) while that load is in flight
        uint4 *z = reinterpret_cast<uint4 *>(frame);
        for (int i = tid; i < C::FRAME / 16; i += C::THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    B = count_now;
    if (DBG && blockIdx.x == 0 && tid == 0) g_trace[1008] = clock64();
    // A unit (a CTA; pair mode: a CTA pair, the two CTAs of a cluster in lockstep, rank r taking the r-th tile of every pair
    // of tiles) runs its share of the tiles in `rounds` rounds of `tiles` tiles each; round r, local tile t is the unit's
    // tile k = r * tiles + t (beyond the share: an empty tile, computed on stale rows and never stored).  A round is nothing
    // but `layers` more entries of the same (layer, tile) sequence: the last layer's epilogue of round r stages the
    // observations of round r + 1 in the frame rows it has just finished reading.  persist: one wave of units and several
    // rounds, so the set-up of a CTA (6 k cycles), its last-epilogue tail and the gap to the next CTA are paid once instead
    // of once per 5-7 tiles; otherwise whole waves of units with one round each.
    const uint32_t rank = C::PAIR ? cluster_ctarank() : 0u;
    const int unit = C::PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const UnitShare us = unit_share(B, C::TILES, sms, C::PAIR, persist != 0, unit, C::NISSUE);
    if (us.n <= 0) return;                               // the whole cluster leaves
    // round r has `tiles` tiles for r < big, tiles - 1 after (big == rounds: all alike, the last ones possibly empty)
    const int tiles = us.tiles, big = us.big, nrounds = us.rounds;
    const int small = big < nrounds ? tiles - 1 : tiles;      // (>= NISSUE >= 1 when it differs from tiles)
    // (compact) board index of half hb of the unit's tile k: b0 + KSTEP * k + hb; it exists if it is below blim
    constexpr int KSTEP = C::PAIR ? 4 : 2;
    const int b0 = C::PAIR ? 4 * us.first + 2 * (int)rank : 2 * us.first;
    const int blim = B < b0 + KSTEP * us.n ? B : b0 + KSTEP * us.n;
    const int layers = 1 + 2 * depth, total = layers * (big * tiles + (nrounds - big) * small);
    const int nissue = tiles < C::NISSUE ? tiles : C::NISSUE;          // (smaller rounds exist only if tiles - 1 >= NISSUE)
    const uint32_t bar0 = smem_u32(bars), frame_s = smem_u32(frame), wb_s = smem_u32(wb);
#define BAR(i) (bar0 + 8u * (uint32_t)(i))

    // The CTA's set-up was 11 % of its life (scripts/nn_trace.py: 8.1 k of 72 k cycles from entry to the first MMA), so
    // its latencies overlap (6.1 k now): the frame is zeroed under the load of the batch size, warp 0 initialises one barrier
    // per lane, requests the first weight slabs and allocates tensor memory while the other warps fetch the observations
    // into registers; the cluster barrier that publishes the barriers to the peer is arrived on (relaxed: the mbarrier-init
    // fence is the release) before the observations are converted and waited for after.
    const int nslabs = 1 + (layers - 1) * C::SLABS;
    const unsigned char *wsrc = wtrunk + (C::PAIR ? (size_t)rank * C::SLAB : 0);
    const size_t wstride = C::PAIR ? 2 * (size_t)C::SLAB : (size_t)C::SLAB;
    const int vslabs = us.rounds * nslabs;                                      // the slab stream repeats every round
    const int early = vslabs < C::NSLOT ? vslabs : C::NSLOT;                 // weight slabs requested before the roles start
    if (warp == 0) {
        if (lane < NBARS) {
            constexpr uint32_t EW = C::PAIR ? 2 * C::GW : C::GW;      // epilogue warps arriving on a leader barrier
            // the MMAs of a tile enter the tensor pipe back to back and in tile order: issuer w waits for TURN[w], which the
            // issuer of the previous tile arrives on (an mbarrier wait parks the thread; a polled shared-memory word takes
            // wavefronts of the shared-memory pipe away from the MMA operand reads the kernel is bound by)
            const uint32_t cnt = (lane >= BAR_PEMPTY && lane < BAR_WFULL) ? EW : (lane >= BAR_WEMPTY && lane < BAR_TURN) ? (uint32_t)nissue : 1u;
            mbar_init(BAR(lane), cnt);
        }
        if (DBG && blockIdx.x == 0 && lane == 0) g_trace[1009] = clock64();
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        __syncwarp();
        if (DBG && blockIdx.x == 0 && lane == 0) g_trace[1010] = clock64();
        if (lane == 0) {
            cnts[0] = 0u;
            mbar_arrive(BAR(BAR_TURN));                        // tile 0 may go
            // the first weight slabs: their barriers are this CTA's own, nothing else has to be ready
            for (int s = 0; s < early; s++) {
                const int ss = s % nslabs;
                const uint32_t bytes = ss == 0 ? (uint32_t)(C::PARTS * C::STEM_PART) : (uint32_t)C::SLAB;
                mbar_expect_tx(BAR(BAR_WFULL + s), bytes);
                bulk_g2s(wb_s + (uint32_t)(s * C::SLAB), wsrc + (size_t)ss * wstride, bytes, BAR(BAR_WFULL + s));
            }
        }
        __syncwarp();
        if (DBG && blockIdx.x == 0 && lane == 0) g_trace[1004] = clock64();
        if (C::PAIR) tmem_alloc_pair<512>(smem_u32(tmem_slot)); else tmem_alloc<512>(smem_u32(tmem_slot));
        if (DBG && blockIdx.x == 0 && lane == 0) g_trace[1005] = clock64();
    }
    const int HW = H * W, n_obs = tiles * 2 * HW;
    float2 oc[4];
    bool o_have = false;
    int o_row = 0;
    if (tid < n_obs) {
        const int bl = tid / HW, pos = tid - bl * HW, y = pos / W, xx = pos - y * W;
        int gb = b0 + KSTEP * (bl >> 1) + (bl & 1);
        o_have = gb < blim;
        if (o_have && rows != nullptr) gb = rows[gb];
        o_row = C::PADR + bl * 64 + y * 8 + xx;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            oc[k].x = (o_have && 2 * k < in_ch) ? obs[((size_t)gb * in_ch + 2 * k) * HW + pos] : 0.0f;
            oc[k].y = (o_have && 2 * k + 1 < in_ch) ? obs[((size_t)gb * in_ch + 2 * k + 1) * HW + pos] : 0.0f;
        }
    }
    tc_fence_before();
    if (DBG && blockIdx.x == 0 && tid == 64) g_trace[1006] = clock64();         // a warp that only zeroed the frame is done
    __syncthreads();
    if (DBG && blockIdx.x == 0 && tid == 0) g_trace[1007] = clock64();
    // the peer's barriers must exist before anything arrives on them: arrive on the cluster barrier now, wait for it only
    // when the roles start (1.4 k cycles that the observation staging hides)
    if (C::PAIR) asm volatile("barrier.cluster.arrive.relaxed.aligned;\n" ::: "memory");   // fence.mbarrier_init above is the release
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (DBG && blockIdx.x == 0 && tid == 0) g_trace[1001] = clock64();          // barriers, tensor memory, zeroed frame, parameters
    const uint32_t lead0 = C::PAIR ? map_to_rank(bar0, 0u) : bar0;      // the leader's barrier block
#define LBAR(i) (lead0 + 8u * (uint32_t)(i))
    const ConstPrm<C::CH> pv{P};

    // observation -> chunk plane 0 (channels >= in_ch stay zero)
    auto stage_obs = [&](const float2 (&c)[4], int frow) {
        uint32_t o[4], o2[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (C::F16) {
                o[k] = f16x2_sat(c[k]);
                const float2 hf = __half22float2(*reinterpret_cast<const __half2 *>(&o[k]));
                o2[k] = f16x2_sat(make_float2(c[k].x - hf.x, c[k].y - hf.y));
            } else {
                const __nv_bfloat162 h = __float22bfloat162_rn(c[k]);
                o[k] = bf2_bits(h);
                const float2 hf = __bfloat1622float2(h);
                o2[k] = bf2_bits(__float22bfloat162_rn(make_float2(c[k].x - hf.x, c[k].y - hf.y)));
            }
        }
        unsigned char *d = frame + (size_t)frow * 16;
        *reinterpret_cast<uint4 *>(d) = make_uint4(o[0], o[1], o[2], o[3]);
        if (C::PARTS == 2) *reinterpret_cast<uint4 *>(d + C::FPART) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
    };
    if (tid < n_obs) stage_obs(oc, o_row);
    for (int i = tid + C::THREADS; i < n_obs; i += C::THREADS) {          // more positions than threads (7 tiles of 7x7 boards)
        const int bl = i / HW, pos = i - bl * HW, y = pos / W, xx = pos - y * W;
        int gb = b0 + KSTEP * (bl >> 1) + (bl & 1);
        const bool have = gb < blim;
        if (have && rows != nullptr) gb = rows[gb];
        float2 c[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            c[k].x = (have && 2 * k < in_ch) ? obs[((size_t)gb * in_ch + 2 * k) * HW + pos] : 0.0f;
            c[k].y = (have && 2 * k + 1 < in_ch) ? obs[((size_t)gb * in_ch + 2 * k + 1) * HW + pos] : 0.0f;
        }
        stage_obs(c, C::PADR + bl * 64 + y * 8 + xx);
    }
    fence_proxy_async();
    __syncthreads();
    if (C::PAIR) asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (DBG && blockIdx.x == 0 && tid == 0) g_trace[1002] = clock64();          // observations in the frame

    if (warp < 3) {
        // ---- MMA issuers: warp k feeds the tiles g = k, k + nissue, ... (g = layer * tiles + tile) ---------------
        if (warp < nissue && rank == 0 && elect_one_sync()) {
            int v = 0, l = 0, rs = 0, t = warp, seen = 0, turn = 0;     // v = round * layers + l; rs = round * nslabs (warp < tiles)
            int r = 0, tc = tiles;                                     // round, tiles of this round
            const uint32_t my_turn = BAR(BAR_TURN + warp), next_turn = BAR(BAR_TURN + (warp + 1 == nissue ? 0 : warp + 1));
#pragma unroll 1
            for (int g = warp; g < total; g += nissue) {
                const int slot = g % C::NRING, use = g / C::NRING;
                const uint32_t d_tmem = tmem_base + C::COL_P + (uint32_t)(slot * C::NACC);
                if (C::PAIR) {
                    if (v > 0) mbar_wait_cluster(BAR(BAR_READY + t), (uint32_t)((v - 1) & 1));
                    if (use > 0) mbar_wait_cluster(BAR(BAR_PEMPTY + slot), (uint32_t)((use - 1) & 1));
                } else {
                    if (v > 0) mbar_wait(BAR(BAR_READY + t), (uint32_t)((v - 1) & 1));         // my tile's previous epilogue
                    if (use > 0) mbar_wait(BAR(BAR_PEMPTY + slot), (uint32_t)((use - 1) & 1));  // ring slot drained
                }
                tc_fence_after();
                if (nissue > 1) mbar_wait(my_turn, (uint32_t)(turn & 1));          // a single issuer is in order by itself
                turn++;
                if (DBG && blockIdx.x == 0 && g < 256) g_trace[4 * g] = clock64();
                const int s0 = rs + (l == 0 ? 0 : 1 + (l - 1) * C::SLABS), ns = l == 0 ? 1 : C::SLABS;
                const bool my_last = t + nissue >= tc;             // my last tile of this layer: release its slabs
#pragma unroll 1
                for (int j = 0; j < ns; j++) {
                    const int s = s0 + j, ws = s % C::NSLOT;
                    if (s >= seen) {
                        mbar_wait(BAR(BAR_WFULL + ws), (uint32_t)((s / C::NSLOT) & 1));
                        if (C::PAIR) mbar_wait_cluster(BAR(BAR_WFULLP + ws), (uint32_t)((s / C::NSLOT) & 1));   // the peer's half
                        seen = s + 1;
                        tc_fence_after();
                    }
                    issue_slab<C>(frame_s, wb_s + (uint32_t)(ws * C::SLAB), d_tmem, t, l, j, (nissue > 1 && j == ns - 1) ? next_turn : 0u);
                    if (my_last) { if (C::PAIR) umma_commit_pair(BAR(BAR_WEMPTY + ws)); else umma_commit(BAR(BAR_WEMPTY + ws)); }
                }
                if (C::PAIR) umma_commit_pair(BAR(BAR_PFULL + slot)); else umma_commit(BAR(BAR_PFULL + slot));
                if (DBG && blockIdx.x == 0 && g < 256) g_trace[4 * g + 1] = clock64();
                t += nissue;
                while (t >= tc) { t -= tc; v++; if (++l == layers) { l = 0; rs += nslabs; if (++r == big) tc = small; } }
            }
        }
        // ---- peer CTA of a pair: tell the leader when MY half of a weight slab has landed ------------------------------
        if (C::PAIR && warp == 0 && rank == 1 && elect_one_sync()) {
#pragma unroll 1
            for (int s = 0; s < vslabs; s++) {
                const int ws = s % C::NSLOT;
                mbar_wait(BAR(BAR_WFULL + ws), (uint32_t)((s / C::NSLOT) & 1));
                mbar_arrive_cluster(LBAR(BAR_WFULLP + ws));
            }
        }
        __syncwarp();
    } else if (warp == 3) {
        // ---- weight producer (pair mode: each CTA streams its own half of every slab: [slab][rank][...]) --------------
        if (elect_one_sync()) {
#pragma unroll 1
            for (int s = early; s < vslabs; s++) {
                const int ws = s % C::NSLOT, use = s / C::NSLOT, ss = s % nslabs;
                if (use > 0) mbar_wait(BAR(BAR_WEMPTY + ws), (uint32_t)((use - 1) & 1));
                const uint32_t bytes = ss == 0 ? (uint32_t)(C::PARTS * C::STEM_PART) : (uint32_t)C::SLAB;
                mbar_expect_tx(BAR(BAR_WFULL + ws), bytes);
                bulk_g2s(wb_s + (uint32_t)(ws * C::SLAB), wsrc + (size_t)ss * wstride, bytes, BAR(BAR_WFULL + ws));
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue groups --------------------------------------------------------------------------------------
        const int e = warp - C::EPI_WARP0, grp = e / C::GW, within = e % C::GW, q = warp & 3, cg = within >> 2;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int r0 = q * 32 + lane, fy = (r0 & 63) >> 3, fx = r0 & 7;
        const bool live_pos = fx < W && fy < H;
        const int pos = fy * W + fx;
        int l = 0, kb = 0, t = grp, r = 0, tc = tiles;       // layer, first tile of the round, tile of the round, round, its tiles
        while (t >= tc) { t -= tc; if (++l == layers) { l = 0; kb += tc; if (++r == big) tc = small; } }
#pragma unroll 1
        for (int g = grp; g < total; g += C::NGROUPS) {
            const int slot = g % C::NRING, use = g / C::NRING;
            const bool is_c1 = (l & 1) == 1, last = l + 1 == layers;
            const uint32_t t_p = tmem_base + lane_off + C::COL_P + (uint32_t)(slot * C::NACC + 16 * cg);
            const uint32_t t_x = tmem_base + lane_off + C::COL_X + (uint32_t)(C::CH * t + 16 * cg);
            const int brd = b0 + KSTEP * (kb + t) + (r0 >> 6);          // (compact) board index of this row
            const bool live = live_pos && brd < blim;
            unsigned char *dst;
            size_t part_stride, chunk_stride;
            if (last) {     // head GEMM A operand: [part][M tile of 128 boards][K chunk = pos * C8 + c][board][16 B]
                chunk_stride = (size_t)128 * 16;
                part_stride = (size_t)MT * KC * chunk_stride;
                dst = gact + ((size_t)(brd >> 7) * KC + (size_t)(pos * C::C8 + 2 * cg)) * chunk_stride + (size_t)(brd & 127) * 16;
            } else {
                chunk_stride = C::PLANE;
                part_stride = C::FPART;
                dst = frame + (size_t)(2 * cg) * C::PLANE + (size_t)(C::PADR + 128 * t + r0) * 16;
            }
            float *dmp = nullptr;
            if (DBG && dump != nullptr && l == dump_layer && brd < blim) dmp = dump + ((size_t)brd * 64 + (r0 & 63)) * C::CH + 16 * cg;
            // Last layer of a round that is not the last (`handoff`): its epilogue writes to global memory only, and the tile's
            // frame rows are free as soon as its MMAs have completed.  So the next round's observation of this thread's
            // position is requested into L1 now, under the wait for those MMAs (its row index is the one register held across
            // the wait: the kernel sits at its register limit, and values parked in local memory would expose the latency of
            // both loads), converted into the frame right after the wait, and epilogue_tile hands the tile to the next round's
            // stem together with the accumulator -- the stem's MMAs run under this epilogue's stores.
            // (the tile slot exists in the next round unless that round is a smaller one and this is the last slot)
            const bool handoff = last && r + 1 < nrounds && t < (r + 1 < big ? tiles : small);
            int ngb = -1;
            if (handoff && cg == 0 && live_pos) {
                const int nb = brd + KSTEP * tc;
                if (nb < blim) {
                    ngb = rows != nullptr ? rows[nb] : nb;
                    const float *o = obs + (size_t)ngb * in_ch * HW + pos;
                    for (int k = 0; k < in_ch; k++) asm volatile("prefetch.global.L1 [%0];\n" ::"l"(o + (size_t)k * HW));
                }
            }
            if (within == 0) mbar_wait<32>(BAR(BAR_PFULL + slot), (uint32_t)(use & 1));
            asm volatile("bar.sync %0, %1;\n" ::"r"(1 + grp), "r"(C::GW * 32) : "memory");
            tc_fence_after();
            if (DBG && blockIdx.x == 0 && within == 0 && lane == 0 && g < 256) g_trace[4 * g + 2] = clock64();
            if (ngb >= 0) {
                float2 nc[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    nc[k].x = 2 * k < in_ch ? obs[((size_t)ngb * in_ch + 2 * k) * HW + pos] : 0.0f;
                    nc[k].y = 2 * k + 1 < in_ch ? obs[((size_t)ngb * in_ch + 2 * k + 1) * HW + pos] : 0.0f;
                }
                stage_obs(nc, C::PADR + 128 * t + r0);
            }
            const uint32_t rdy = handoff ? LBAR(BAR_READY + t) : 0u;
            const int ho = 16 * cg;
            if (l == 0) {
                epilogue_tile<C, EPI_STEM, DBG>(t_p, t_x, dst, part_stride, chunk_stride, pv, ho, depth > 0 ? ho : -1, live, lane,
                                                LBAR(BAR_PEMPTY + slot), rdy, dmp);
            } else if (is_c1) {
                epilogue_tile<C, EPI_CONV1, DBG>(t_p, t_x, dst, part_stride, chunk_stride, pv, l * C::CH + ho, -1, live, lane,
                                                 LBAR(BAR_PEMPTY + slot), 0u, dmp);
            } else {
                const int nblk = l >> 1;                      // the block that consumes x next
                epilogue_tile<C, EPI_CONV2, DBG>(t_p, t_x, dst, part_stride, chunk_stride, pv, 0, last ? -1 : nblk * C::CH + ho, live, lane,
                                                 LBAR(BAR_PEMPTY + slot), rdy, dmp);
            }
            if (!handoff) {
                fence_proxy_async();          // the next layer's MMAs read these rows through the async proxy
                __syncwarp();
                if (lane == 0) { if (C::PAIR) mbar_arrive_cluster(LBAR(BAR_READY + t)); else mbar_arrive(BAR(BAR_READY + t)); }
            }
            if (DBG && blockIdx.x == 0 && within == 0 && lane == 0 && g < 256) g_trace[4 * g + 3] = clock64();
            t += C::NGROUPS;
            while (t >= tc) { t -= tc; if (++l == layers) { l = 0; kb += tc; if (++r == big) tc = small; } }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (C::PAIR) cluster_sync_all();                     // the leader's MMAs read the peer's shared memory until the very end
    if (warp == 0) { if (C::PAIR) tmem_dealloc_pair<512>(tmem_base); else tmem_dealloc<512>(tmem_base); }
    if (DBG && blockIdx.x == 0 && tid == 0) g_trace[1003] = clock64();          // CTA exit
    if (DBG && tid == 0 && blockIdx.x < 2048) g_cta_trace[3 * blockIdx.x + 2] = global_ns();
#undef BAR
#undef LBAR
}

// ---- 128-channel trunk (connect4/train.py:44-49: 128 channels x 8 blocks) ---------------------------------------------
// One layer of split operands is 590 KB -- it cannot stay in shared memory, and N = 3 CH = 384 exceeds an MMA's N -- so
// this kernel differs from k_trunk_tc in two places.  (1) Every tap is its own MMA: a K-major no-swizzle descriptor may
// start at any 16-byte row (measured on the B200, scripts/umma_shift_probe.cu / umma_shift_timing.cu: exact, and an
// M = 128, N = 128, K = 16 MMA takes 64 cycles = its math floor whatever the shift), so tap (dy, dx) is `start address +=
// 16 (8 dy + dx)`, N = CH = 128, and the epilogue needs no rotations.  (2) Weights stream through a ring of 16 KB slabs
// ((tap, 32 input channels): [part][4 K chunks][128 cout][8 cin]) and the MMAs run slab-major: the CTA's two tiles (four
// boards; tensor memory = 2 x (128 residual + 128 accumulator columns)) consume a slab one after the other, so a layer's
// weights cross L2 -> shared memory once per four boards.  One issuing thread, one producer, one epilogue group of eight
// warps per tile (TMEM lane quadrant x half of the channels, four 16-channel groups each).
template <int PREC_>
struct WideCfg {
    static constexpr int CH = 128, TILES = 2, PREC = PREC_, MAXDEPTH = 8, MAXLAYERS = 1 + 2 * MAXDEPTH;
    static constexpr int PARTS = prec_split(PREC) ? 2 : 1;
    static constexpr bool F16 = prec_f16(PREC);
    static constexpr int C8 = CH / 8;
    static constexpr int ROWS = TILES * 128, PADR = 16, FROWS = ROWS + 2 * PADR;
    static constexpr int PLANE = FROWS * 16, FPART = C8 * PLANE, FRAME = PARTS * FPART;
    static constexpr int WCHUNK = CH * 16;                              // one 8-channel K chunk of the B operand
    static constexpr int SLAB_PART = 4 * WCHUNK, SLAB = PARTS * SLAB_PART;
    static constexpr int KQ = C8 / 4, LSLABS = 9 * KQ, STEM_SLABS = 3;  // slabs per tap / per trunk layer / of the stem
    static constexpr int NSLOT = PARTS == 2 ? 4 : 8;
    static constexpr int CGW = 2, CGI = CH / 16 / CGW;                  // warps per lane quadrant, channel groups per warp
    static constexpr int GW = 4 * CGW, EPI_WARP0 = 4;
    static constexpr int WARPS = EPI_WARP0 + TILES * GW, THREADS = WARPS * 32;
    static constexpr uint32_t COL_X = 0, COL_P = TILES * CH;
    static constexpr int PRM_FLOATS = MAXLAYERS * CH + 2 * MAXDEPTH * CH;
    static constexpr size_t SMEM = (size_t)FRAME + (size_t)NSLOT * SLAB + (size_t)PRM_FLOATS * 4 + 24 * 8 + 64;
    static constexpr uint32_t IDESC = umma_idesc(CH, F16);
    static_assert(COL_P + TILES * CH <= 512, "tensor memory budget");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    static_assert(PADR >= 9 + 1, "taps reach 9 frame rows beyond a tile");
};

// Epilogue of one tile row for this warp's CGI 16-channel groups, accumulator = the finished convolution (same cases as
// epilogue_tile).  The tensor-memory loads of group i + 1 are in flight while group i is processed, and the residual
// stores are waited for once at the end: between two layers the tensor pipe idles for exactly this function.
template <class C, int EPI, bool DBG>
__device__ __forceinline__ void epilogue_wide_row(uint32_t t_p, uint32_t t_x, unsigned char *dst, size_t part_stride, size_t chunk_stride,
                                                  const SmemPrm &pv, int boff, int soff, bool live, float *dump_row)
{
    uint32_t acc[2][16], xv[2][16];
    tmem_ld16(t_p, acc[0]);
    if (EPI == EPI_CONV2) tmem_ld16(t_x, xv[0]);
#pragma unroll
    for (int i = 0; i < C::CGI; i++) {
        tmem_ld_wait();
        if (i + 1 < C::CGI) {
            tmem_ld16(t_p + (uint32_t)(16 * (i + 1)), acc[(i + 1) & 1]);
            if (EPI == EPI_CONV2) tmem_ld16(t_x + (uint32_t)(16 * (i + 1)), xv[(i + 1) & 1]);
        }
        float2 r[8];
#pragma unroll
        for (int c = 0; c < 8; c++) r[c] = f2(acc[i & 1][2 * c], acc[i & 1][2 * c + 1]);
        epilogue_finish<C, EPI, DBG, false>(r, xv[i & 1], t_x + (uint32_t)(16 * i), dst + (size_t)(2 * i) * chunk_stride, part_stride, chunk_stride,
                                            pv, boff + 16 * i, soff >= 0 ? soff + 16 * i : -1, live,
                                            DBG && dump_row != nullptr ? dump_row + 16 * i : nullptr);
    }
    if (EPI != EPI_CONV1) tmem_st_wait();
}

template <class C, bool DBG>
__global__ void __launch_bounds__(C::THREADS, 1)
k_trunk_wide(const float *__restrict__ obs, int B, int in_ch, int H, int W, int depth, const unsigned char *__restrict__ wtrunk,
             const float *__restrict__ cbias, const float *__restrict__ bn_scale, const float *__restrict__ bn_shift,
             unsigned char *__restrict__ gact, int MT, int KC, float *__restrict__ dump, int dump_layer,
             const int *__restrict__ rows, const int *__restrict__ count_ptr, int sms)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *frame = smem;
    unsigned char *wb = frame + C::FRAME;                                       // weight slab ring
    float *prm = reinterpret_cast<float *>(wb + (size_t)C::NSLOT * C::SLAB);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(prm + C::PRM_FLOATS);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 24);

    enum { BAR_PFULL = 0, BAR_READY = BAR_PFULL + C::TILES, BAR_WFULL = BAR_READY + C::TILES, BAR_WEMPTY = BAR_WFULL + C::NSLOT,
           NBARS = BAR_WEMPTY + C::NSLOT };
    static_assert(NBARS <= 24, "barrier storage");

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (rows != nullptr) B = *count_ptr;                 // compact mode: the batch size lives in device memory
    const TileShare sh = tile_share(B, C::TILES, sms);
    if ((int)blockIdx.x >= sh.nct) return;
    const int tiles = sh.base + ((int)blockIdx.x < sh.extra ? 1 : 0);
    const int tile0 = (int)blockIdx.x * sh.base + ((int)blockIdx.x < sh.extra ? (int)blockIdx.x : sh.extra);
    const int board0 = 2 * tile0;
    const int layers = 1 + 2 * depth;
    const uint32_t bar0 = smem_u32(bars), frame_s = smem_u32(frame), wb_s = smem_u32(wb);
#define BAR(i) (bar0 + 8u * (uint32_t)(i))

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < C::TILES; i++) { mbar_init(BAR(BAR_PFULL + i), 1); mbar_init(BAR(BAR_READY + i), C::GW); }
            for (int i = 0; i < C::NSLOT; i++) { mbar_init(BAR(BAR_WFULL + i), 1); mbar_init(BAR(BAR_WEMPTY + i), 1); }
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncwarp();
        tmem_alloc<512>(smem_u32(tmem_slot));
    }
    {   // zero the frame (padding rows / columns must read as zero), stage the per-channel parameters
        uint4 *z = reinterpret_cast<uint4 *>(frame);
        for (int i = tid; i < C::FRAME / 16; i += C::THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < layers * C::CH; i += C::THREADS) prm[i] = cbias[i];
        for (int i = tid; i < depth * C::CH; i += C::THREADS) {
            prm[C::MAXLAYERS * C::CH + i] = bn_scale[i];
            prm[C::MAXLAYERS * C::CH + C::MAXDEPTH * C::CH + i] = bn_shift[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const float *s_bias = prm, *s_sc = prm + C::MAXLAYERS * C::CH, *s_sh = prm + C::MAXLAYERS * C::CH + C::MAXDEPTH * C::CH;

    // observation -> chunk plane 0 (channels >= in_ch stay zero)
    const int HW = H * W;
    for (int i = tid; i < tiles * 2 * HW; i += C::THREADS) {
        const int bl = i / HW, pos = i - bl * HW, y = pos / W, xx = pos - y * W;
        int gb = board0 + bl;
        const bool have = gb < B;
        if (have && rows != nullptr) gb = rows[gb];
        float2 c[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            c[k].x = (have && 2 * k < in_ch) ? obs[((size_t)gb * in_ch + 2 * k) * HW + pos] : 0.0f;
            c[k].y = (have && 2 * k + 1 < in_ch) ? obs[((size_t)gb * in_ch + 2 * k + 1) * HW + pos] : 0.0f;
        }
        uint32_t o[4], o2[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (C::F16) {
                o[k] = f16x2_sat(c[k]);
                const float2 hf = __half22float2(*reinterpret_cast<const __half2 *>(&o[k]));
                o2[k] = f16x2_sat(make_float2(c[k].x - hf.x, c[k].y - hf.y));
            } else {
                const __nv_bfloat162 h = __float22bfloat162_rn(c[k]);
                o[k] = bf2_bits(h);
                const float2 hf = __bfloat1622float2(h);
                o2[k] = bf2_bits(__float22bfloat162_rn(make_float2(c[k].x - hf.x, c[k].y - hf.y)));
            }
        }
        unsigned char *d = frame + (size_t)(C::PADR + bl * 64 + y * 8 + xx) * 16;
        *reinterpret_cast<uint4 *>(d) = make_uint4(o[0], o[1], o[2], o[3]);
        if (C::PARTS == 2) *reinterpret_cast<uint4 *>(d + C::FPART) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
    }
    fence_proxy_async();
    __syncthreads();

    const int nslabs = C::STEM_SLABS + (layers - 1) * C::LSLABS;
    constexpr int NPASS = C::PARTS == 2 ? 3 : 1;
    if (warp == 0) {
        // ---- MMA issuer: layer by layer, slab by slab, tile by tile ------------------------------------------------------
        if (elect_one_sync()) {
            int s = 0;
#pragma unroll 1
            for (int l = 0; l < layers; l++) {
                const int ns = l == 0 ? C::STEM_SLABS : C::LSLABS;
#pragma unroll 1
                for (int j = 0; j < ns; j++, s++) {
                    const int ws = s % C::NSLOT;
                    mbar_wait(BAR(BAR_WFULL + ws), (uint32_t)((s / C::NSLOT) & 1));
                    tc_fence_after();
                    const uint32_t w_s = wb_s + (uint32_t)(ws * C::SLAB);
                    // stem slab j: dx = j - 1, K chunks (dy=-1, dy=0, zero, dy=+1) of chunk plane 0: two K steps whose second
                    // chunk is the first one 8 rows further.  Trunk slab j: tap j / KQ, input channels 32 (j % KQ) ..
                    const int tap = j / C::KQ, kq = j - tap * C::KQ;
                    const int shift = l == 0 ? (j - 1) : (8 * (tap / 3 - 1) + (tap % 3 - 1));          // frame rows
#pragma unroll 1
                    for (int t = 0; t < tiles; t++) {
                        if (j == 0 && l > 0) { mbar_wait(BAR(BAR_READY + t), (uint32_t)((l - 1) & 1)); tc_fence_after(); }
                        const uint32_t d_tmem = tmem_base + C::COL_P + (uint32_t)(C::CH * t);
                        const uint32_t row0 = frame_s + (uint32_t)((C::PADR + 128 * t + shift) * 16);
#pragma unroll
                        for (int kk = 0; kk < 2; kk++) {
#pragma unroll
                            for (int pass = 0; pass < NPASS; pass++) {
                                const int pa = pass == 2 ? 1 : 0, pb = pass == 1 ? 1 : 0;
                                uint64_t ad;
                                if (l == 0) ad = umma_desc(row0 + (uint32_t)(pa * C::FPART) + (uint32_t)((kk == 0 ? -8 : 0) * 16), 128u, 128u);
                                else ad = umma_desc(row0 + (uint32_t)(pa * C::FPART + (4 * kq + 2 * kk) * C::PLANE), (uint32_t)C::PLANE, 128u);
                                const uint64_t bd = umma_desc(w_s + (uint32_t)(pb * C::SLAB_PART + 2 * kk * C::WCHUNK), (uint32_t)C::WCHUNK, 128u);
                                umma_f16(d_tmem, ad, bd, C::IDESC, (j > 0 || kk > 0 || pass > 0) ? 1u : 0u);
                            }
                        }
                    }
                    umma_commit(BAR(BAR_WEMPTY + ws));
                }
                for (int t = 0; t < tiles; t++) umma_commit(BAR(BAR_PFULL + t));
            }
        }
        __syncwarp();
    } else if (warp == 3) {
        // ---- weight producer ------------------------------------------------------------------------------------
        if (elect_one_sync()) {
#pragma unroll 1
            for (int s = 0; s < nslabs; s++) {
                const int ws = s % C::NSLOT, use = s / C::NSLOT;
                if (use > 0) mbar_wait(BAR(BAR_WEMPTY + ws), (uint32_t)((use - 1) & 1));
                mbar_expect_tx(BAR(BAR_WFULL + ws), (uint32_t)C::SLAB);
                bulk_g2s(wb_s + (uint32_t)(ws * C::SLAB), wtrunk + (size_t)s * C::SLAB, (uint32_t)C::SLAB, BAR(BAR_WFULL + ws));
            }
        }
        __syncwarp();
    } else if (warp >= C::EPI_WARP0) {
        // ---- epilogue group t: tile t of every layer --------------------------------------------------------------------
        const int e = warp - C::EPI_WARP0, t = e / C::GW, within = e % C::GW, q = warp & 3, cgw = within >> 2;
        if (t < tiles) {
            const uint32_t lane_off = (uint32_t)(q * 32) << 16;
            const int r0 = q * 32 + lane, fy = (r0 & 63) >> 3, fx = r0 & 7;
            const int brd = board0 + 2 * t + (r0 >> 6);                     // (compact) board index of this row
            const bool live = fx < W && fy < H && brd < B;
            const int pos = fy * W + fx;
#pragma unroll 1
            for (int l = 0; l < layers; l++) {
                const bool is_c1 = (l & 1) == 1, last = l + 1 == layers;
                if (within == 0) mbar_wait<32>(BAR(BAR_PFULL + t), (uint32_t)(l & 1));
                asm volatile("bar.sync %0, %1;\n" ::"r"(1 + t), "r"(C::GW * 32) : "memory");
                tc_fence_after();
                {
                    const int cg = cgw * C::CGI, ho = 16 * cg;       // this warp's first channel group
                    const uint32_t t_p = tmem_base + lane_off + C::COL_P + (uint32_t)(C::CH * t + ho);
                    const uint32_t t_x = tmem_base + lane_off + C::COL_X + (uint32_t)(C::CH * t + ho);
                    unsigned char *dst;
                    size_t part_stride, chunk_stride;
                    if (last) {     // head GEMM A operand: [part][M tile of 128 boards][K chunk = pos * C8 + c][board][16 B]
                        chunk_stride = (size_t)128 * 16;
                        part_stride = (size_t)MT * KC * chunk_stride;
                        dst = gact + ((size_t)(brd >> 7) * KC + (size_t)(pos * C::C8 + 2 * cg)) * chunk_stride + (size_t)(brd & 127) * 16;
                    } else {
                        chunk_stride = C::PLANE;
                        part_stride = C::FPART;
                        dst = frame + (size_t)(2 * cg) * C::PLANE + (size_t)(C::PADR + 128 * t + r0) * 16;
                    }
                    float *dmp = nullptr;
                    if (DBG && dump != nullptr && l == dump_layer && brd < B) dmp = dump + ((size_t)brd * 64 + (r0 & 63)) * C::CH + ho;
                    const SmemPrm pv{s_bias, s_sc, s_sh};
                    if (l == 0) {
                        epilogue_wide_row<C, EPI_STEM, DBG>(t_p, t_x, dst, part_stride, chunk_stride, pv, ho, depth > 0 ? ho : -1, live, dmp);
                    } else if (is_c1) {
                        epilogue_wide_row<C, EPI_CONV1, DBG>(t_p, t_x, dst, part_stride, chunk_stride, pv, l * C::CH + ho, -1, live, dmp);
                    } else {
                        const int nblk = l >> 1;                      // the block that consumes x next
                        epilogue_wide_row<C, EPI_CONV2, DBG>(t_p, t_x, dst, part_stride, chunk_stride, pv, 0, last ? -1 : nblk * C::CH + ho, live,
                                                             dmp);
                    }
                }
                tc_fence_before();                // accumulator and residual accesses are complete (wait::ld / wait::st inside)
                fence_proxy_async();              // the next layer's MMAs read these rows through the async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(BAR_READY + t));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
#undef BAR
}

// ---- heads ---------------------------------------------------------------------------------------------------------
constexpr int HMAXST = 12, HKC = 4;                        // deepest operand ring; K chunks (of 8) per pipeline stage
constexpr int HTHREADS = 192;

template <int PREC>
__global__ void __launch_bounds__(HTHREADS, 1)
k_head_tc(const unsigned char *__restrict__ gact, const unsigned char *__restrict__ whead, const float *__restrict__ bhead,
          float *__restrict__ logits, float *__restrict__ policy, float *__restrict__ value, int B, const int *__restrict__ rows,
          const int *__restrict__ count_ptr, int MT, int KC, int NT, int ntiles, int nout_pad, int A, uint32_t tmem_cols, int HSTAGES)
{
    constexpr int PARTS = prec_split(PREC) ? 2 : 1;
    constexpr bool F16 = prec_f16(PREC);
    extern __shared__ __align__(128) unsigned char smem[];
    const int a_part = HKC * 128 * 16, b_part = HKC * NT * 16, stage_bytes = PARTS * (a_part + b_part);
    // HSTAGES stages in flight: with 55 CTAs for 6960 boards the kernel is bound by the latency of its own operand stream
    // (4 stages = 66 KB in flight per SM: 18 us; 12 stages: the HBM time of the 37 MB it reads)
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + (size_t)HSTAGES * stage_bytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * HMAXST + 1);
    enum { BAR_FULL = 0, BAR_EMPTY = HMAXST, BAR_DONE = 2 * HMAXST };
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int mt = blockIdx.x, nt = blockIdx.y;
    if (rows != nullptr) B = *count_ptr;
    if (mt * 128 >= B) return;
    const uint32_t bar0 = smem_u32(bars), st_s = smem_u32(smem);
#define BAR(i) (bar0 + 8u * (uint32_t)(i))
    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < HSTAGES; i++) { mbar_init(BAR(BAR_FULL + i), 1); mbar_init(BAR(BAR_EMPTY + i), 1); }
            mbar_init(BAR(BAR_DONE), 1);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nk = KC / HKC;
    if (warp == 0) {
        if (elect_one_sync()) {
#pragma unroll 1
            for (int i = 0; i < nk; i++) {
                const int st = i % HSTAGES, use = i / HSTAGES;
                if (use > 0) mbar_wait(BAR(BAR_EMPTY + st), (uint32_t)((use - 1) & 1));
                mbar_expect_tx(BAR(BAR_FULL + st), (uint32_t)stage_bytes);
                const uint32_t dst = st_s + (uint32_t)(st * stage_bytes);
#pragma unroll
                for (int p = 0; p < PARTS; p++) {
                    bulk_g2s(dst + (uint32_t)(p * a_part), gact + (((size_t)p * MT + mt) * KC + (size_t)i * HKC) * 2048, (uint32_t)a_part,
                             BAR(BAR_FULL + st));
                    bulk_g2s(dst + (uint32_t)(PARTS * a_part + p * b_part),
                             whead + (((size_t)p * ntiles + nt) * KC + (size_t)i * HKC) * (size_t)NT * 16, (uint32_t)b_part, BAR(BAR_FULL + st));
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (elect_one_sync()) {
            const uint32_t idesc = umma_idesc(NT, F16);
#pragma unroll 1
            for (int i = 0; i < nk; i++) {
                const int st = i % HSTAGES, use = i / HSTAGES;
                mbar_wait(BAR(BAR_FULL + st), (uint32_t)(use & 1));
                tc_fence_after();
                const uint32_t a_s = st_s + (uint32_t)(st * stage_bytes), b_s = a_s + (uint32_t)(PARTS * a_part);
#pragma unroll
                for (int ks = 0; ks < HKC / 2; ks++) {
#pragma unroll
                    for (int pass = 0; pass < (PARTS == 2 ? 3 : 1); pass++) {
                        const int pa = pass == 2 ? 1 : 0, pb = pass == 1 ? 1 : 0;
                        const uint64_t ad = umma_desc(a_s + (uint32_t)(pa * a_part + 2 * ks * 2048), 2048u, 128u);
                        const uint64_t bd = umma_desc(b_s + (uint32_t)(pb * b_part + 2 * ks * NT * 16), (uint32_t)(NT * 16), 128u);
                        umma_f16(tmem_base, ad, bd, idesc, (i > 0 || ks > 0 || pass > 0) ? 1u : 0u);
                    }
                }
                umma_commit(BAR(BAR_EMPTY + st));
            }
            umma_commit(BAR(BAR_DONE));
        }
        __syncwarp();
    } else {
        // epilogue: warp w reads TMEM lanes 32 (w % 4) ..; lane = board within the M tile
        const int q = warp & 3, row = q * 32 + lane, i = mt * 128 + row;
        mbar_wait<64>(BAR(BAR_DONE), 0u);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool have = i < B;
        if (ntiles == 1 && NT == 16) {             // all outputs in one thread: softmax here
            uint32_t v[16];
            tmem_ld16(t_row, v);
            tmem_ld_wait();
            if (have) {
                const int gb = rows != nullptr ? rows[i] : i;
                float lg[16];
#pragma unroll
                for (int j = 0; j < 16; j++) lg[j] = __uint_as_float(v[j]) + bhead[j];
                float mp = -INFINITY, mv = -INFINITY;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    if (j < A) mp = fmaxf(mp, lg[j]);
                    else if (j < A + 3) mv = fmaxf(mv, lg[j]);
                }
                float sp = 0.0f, sv = 0.0f;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    if (j < A) { lg[j] = expf(lg[j] - mp); sp += lg[j]; }
                    else if (j < A + 3) { lg[j] = expf(lg[j] - mv); sv += lg[j]; }
                }
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    if (j < A) policy[(size_t)gb * A + j] = lg[j] / sp;
                    else if (j < A + 3) value[(size_t)gb * 3 + (j - A)] = lg[j] / sv;
                }
            }
        } else {
#pragma unroll 1
            for (int c = 0; c < NT / 16; c++) {
                uint32_t v[16];
                tmem_ld16(t_row + (uint32_t)(16 * c), v);
                tmem_ld_wait();
                if (have) {
                    float *o = logits + (size_t)i * nout_pad + (size_t)nt * NT + 16 * c;
                    const float *bb = bhead + nt * NT + 16 * c;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 b4 = *reinterpret_cast<const float4 *>(bb + j);
                        *reinterpret_cast<float4 *>(o + j) = make_float4(__uint_as_float(v[j]) + b4.x, __uint_as_float(v[j + 1]) + b4.y,
                                                                         __uint_as_float(v[j + 2]) + b4.z, __uint_as_float(v[j + 3]) + b4.w);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
#undef BAR
}

// exp(log_softmax) of the policy logits [0, A) and the value logits [A, A+3): one warp per board
__global__ void __launch_bounds__(256)
k_softmax(const float *__restrict__ logits, float *__restrict__ policy, float *__restrict__ value, int B, const int *__restrict__ rows,
          const int *__restrict__ count_ptr, int nout_pad, int A)
{
    if (rows != nullptr) B = *count_ptr;
    const int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (i >= B) return;
    const int gb = rows != nullptr ? rows[i] : i;
    const float *lg = logits + (size_t)i * nout_pad;
    float m = -INFINITY;
    for (int j = lane; j < A; j += 32) m = fmaxf(m, lg[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.0f;
    for (int j = lane; j < A; j += 32) s += expf(lg[j] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.0f / s;
    for (int j = lane; j < A; j += 32) policy[(size_t)gb * A + j] = expf(lg[j] - m) * inv;
    if (lane == 0) {
        const float v0 = lg[A], v1 = lg[A + 1], v2 = lg[A + 2], mv = fmaxf(v0, fmaxf(v1, v2));
        const float e0 = expf(v0 - mv), e1 = expf(v1 - mv), e2 = expf(v2 - mv), sv = e0 + e1 + e2;
        value[(size_t)gb * 3] = e0 / sv; value[(size_t)gb * 3 + 1] = e1 / sv; value[(size_t)gb * 3 + 2] = e2 / sv;
    }
}

// ---- launch ----------------------------------------------------------------------------------------------------------
//                          CH  TILES PREC NGROUPS DYS NSLOT
// 32 channels: two epilogue groups keep up with the MMAs of a tile (measured 191.9 us with two, 195.6 us with three for
// 6960 boards in split mode; 121.3 vs 124.8 us in bf16) and leave the registers of eight warps unused
template <int PREC> using Cfg32 = TrunkCfg<32, 7, PREC, 2, 3, prec_split(PREC) ? 2 : 3>;
template <int PREC> using Cfg64 = TrunkCfg<64, 2, PREC, 1, 1, prec_split(PREC) ? 3 : 6>;
// CTA pairs (AZB_NNG_PAIR): half of every weight slab per CTA, so the ring holds more of them
template <int PREC> using Cfg32P = TrunkCfg<32, 7, PREC, 3, 3, prec_split(PREC) ? 3 : 4, true>;
template <int PREC> using Cfg64P = TrunkCfg<64, 2, PREC, 1, 1, 6, true>;


int sm_count()
{
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaDeviceProp prop;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return -1;
        sms = prop.multiProcessorCount;
    }
    return sms;
}

template <class K, class... Args>
int launch_maybe_pair(K kernel, bool pair, int grid, int threads, size_t smem, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pair ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...) == cudaSuccess ? 0 : -2;
}

template <class C>
int launch_trunk(const azb_nng_net *n, const float *obs, int batch, const int *rows, const int *count, cudaStream_t s, float *dump,
                 int dump_layer)
{
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_trunk_tc<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(k_trunk_tc<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess)
            return -2;
        configured = true;
    }
    int sms = sm_count();
    if (sms <= 0) return -2;
    if (C::PAIR) {
        // tiles are dealt in whole waves of resident CTA pairs: ask how many clusters of two fit (a GPC with an odd number of
        // usable SMs leaves one unpaired)
        static int pairs = 0;
        if (pairs == 0) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(sms & ~1));
            cfg.blockDim = dim3((unsigned)C::THREADS);
            cfg.dynamicSmemBytes = C::SMEM;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, k_trunk_tc<C, false>, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = sms / 2; }
            pairs = n < sms / 2 ? n : sms / 2;
            if (getenv("AZB_NN_VERBOSE")) fprintf(stderr, "azb_nng: %d resident CTA pairs (cudaOccupancyMaxActiveClusters %d, %d SMs)\n", pairs, n, sms);
        }
        sms = 2 * pairs;
    }
    // compact mode: upper bound, surplus CTAs exit at once.  Pair mode: pairs of tiles over pairs of CTAs
    // one wave of persistent units, or whole waves of one-round units: the net's flags, else AZB_NNG_PERSIST=0 / 1, else the default
    static int persist_default = -1;
    if (persist_default < 0) {
        const char *e = getenv("AZB_NNG_PERSIST");
        persist_default = e != nullptr ? (atoi(e) != 0 ? 1 : 0) : 1;
    }
    const int persist = (n->flags & AZB_NNG_PERSIST) ? 1 : (n->flags & AZB_NNG_ONE_ROUND) ? 0 : persist_default;
    const int grid = (C::PAIR ? 2 : 1) * unit_share(batch, C::TILES, sms, C::PAIR, persist != 0, 0, C::NISSUE).units;
    const int MT = (n->max_boards + 127) / 128;
    const unsigned char *wt = reinterpret_cast<const unsigned char *>(n->wtrunk);
    unsigned char *gact = reinterpret_cast<unsigned char *>(n->gact);
    // per-channel parameters by value (constant bank): h_params = host copy of [cbias | bn_scale | bn_shift]
    NetParams<C::CH> P = {};
    const int L = 1 + 2 * n->depth, D1 = n->depth > 0 ? n->depth : 1;
    memcpy(P.bias, n->h_params, sizeof(float) * (size_t)L * C::CH);
    memcpy(P.sc, n->h_params + (size_t)L * C::CH, sizeof(float) * (size_t)D1 * C::CH);
    memcpy(P.sh, n->h_params + (size_t)(L + D1) * C::CH, sizeof(float) * (size_t)D1 * C::CH);
    if (dump != nullptr)
        return launch_maybe_pair(k_trunk_tc<C, true>, C::PAIR, grid, C::THREADS, C::SMEM, s, obs, batch, (int)n->in_channels, (int)n->board_h,
                                 (int)n->board_w, (int)n->depth, wt, P, gact, MT, (int)n->head_kc, dump, dump_layer, rows, count, sms, persist);
    return launch_maybe_pair(k_trunk_tc<C, false>, C::PAIR, grid, C::THREADS, C::SMEM, s, obs, batch, (int)n->in_channels, (int)n->board_h,
                             (int)n->board_w, (int)n->depth, wt, P, gact, MT, (int)n->head_kc, (float *)nullptr, -1, rows, count, sms, persist);
}

template <class C>
int launch_wide(const azb_nng_net *n, const float *obs, int batch, const int *rows, const int *count, cudaStream_t s, float *dump,
                int dump_layer)
{
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k_trunk_wide<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(k_trunk_wide<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess)
            return -2;
        configured = true;
    }
    const int sms = sm_count();
    if (sms <= 0) return -2;
    const int grid = tile_share(batch, C::TILES, sms).nct;      // compact mode: upper bound, surplus CTAs exit at once
    const int MT = (n->max_boards + 127) / 128;
    if (dump != nullptr)
        k_trunk_wide<C, true><<<grid, C::THREADS, C::SMEM, s>>>(obs, batch, n->in_channels, n->board_h, n->board_w, n->depth,
                                                               reinterpret_cast<const unsigned char *>(n->wtrunk), n->cbias, n->bn_scale,
                                                               n->bn_shift, reinterpret_cast<unsigned char *>(n->gact), MT, n->head_kc, dump,
                                                               dump_layer, rows, count, sms);
    else
        k_trunk_wide<C, false><<<grid, C::THREADS, C::SMEM, s>>>(obs, batch, n->in_channels, n->board_h, n->board_w, n->depth,
                                                                reinterpret_cast<const unsigned char *>(n->wtrunk), n->cbias, n->bn_scale,
                                                                n->bn_shift, reinterpret_cast<unsigned char *>(n->gact), MT, n->head_kc,
                                                                nullptr, -1, rows, count, sms);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

template <int PREC>
int launch_head(const azb_nng_net *n, float *policy, float *value, int batch, const int *rows, const int *count, cudaStream_t s)
{
    constexpr int PARTS = prec_split(PREC) ? 2 : 1;
    const int NT = n->head_nt, ntiles = n->head_ntiles, nout_pad = NT * ntiles;
    const size_t stage = (size_t)PARTS * (HKC * 128 * 16 + HKC * NT * 16), fixed = (2 * HMAXST + 1) * 8 + 64;
    int stages = (int)((220 * 1024 - fixed) / stage);
    stages = stages > HMAXST ? HMAXST : stages;
    if (stages > n->head_kc / HKC) stages = n->head_kc / HKC;
    if (stages < 2) return -1;
    const size_t smem = (size_t)stages * stage + fixed;
    static size_t configured = 0;
    if (smem > configured) {
        if (cudaFuncSetAttribute(k_head_tc<PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
        configured = smem;
    }
    uint32_t cols = 32;
    while ((int)cols < NT) cols <<= 1;
    const int MT = (n->max_boards + 127) / 128, mtiles = (batch + 127) / 128;
    k_head_tc<PREC><<<dim3(mtiles, ntiles), HTHREADS, smem, s>>>(reinterpret_cast<const unsigned char *>(n->gact),
                                                                 reinterpret_cast<const unsigned char *>(n->whead), n->bhead, n->logits, policy,
                                                                 value, batch, rows, count, MT, n->head_kc, NT, ntiles, nout_pad, n->action_size,
                                                                 cols, stages);
    if (cudaGetLastError() != cudaSuccess) return -2;
    if (!(ntiles == 1 && NT == 16)) {
        k_softmax<<<(batch + 7) / 8, 256, 0, s>>>(n->logits, policy, value, batch, rows, count, nout_pad, n->action_size);
        if (cudaGetLastError() != cudaSuccess) return -2;
    }
    return 0;
}

template <int PREC>
int forward_prec(const azb_nng_net *n, const float *obs, float *policy, float *value, int batch, const int *rows, const int *count,
                 cudaStream_t s, float *dump, int dump_layer)
{
    int rc;
    const bool pair = (n->flags & AZB_NNG_PAIR) != 0;
    if (n->channels == 32) rc = pair ? launch_trunk<Cfg32P<PREC>>(n, obs, batch, rows, count, s, dump, dump_layer)
                                     : launch_trunk<Cfg32<PREC>>(n, obs, batch, rows, count, s, dump, dump_layer);
    else if (n->channels == 64) rc = pair ? launch_trunk<Cfg64P<PREC>>(n, obs, batch, rows, count, s, dump, dump_layer)
                                          : launch_trunk<Cfg64<PREC>>(n, obs, batch, rows, count, s, dump, dump_layer);
    else rc = launch_wide<WideCfg<PREC>>(n, obs, batch, rows, count, s, dump, dump_layer);
    if (rc != 0) return rc;
    return launch_head<PREC>(n, policy, value, batch, rows, count, s);
}

int forward(const azb_nng_net *n, const float *obs, float *policy, float *value, int batch, const int *rows, const int *count, void *stream,
            float *dump, int dump_layer)
{
    if (!n || !obs || !policy || !value || batch <= 0 || (rows != nullptr) != (count != nullptr)) return -7;
    if ((n->channels != 32 && n->channels != 64 && n->channels != 128) || n->board_h < 1 || n->board_h > 7 || n->board_w < 1 ||
        n->board_w > 7 || n->in_channels < 1 || n->in_channels > 8 || n->depth < 0 ||
        n->depth > (n->channels == 128 ? WideCfg<AZB_NN_BF16>::MAXDEPTH : MAXD) || n->action_size < 1 || batch > n->max_boards ||
        n->head_nt % 16 != 0 || n->head_nt < 16 || n->head_nt > 256 || n->head_ntiles < 1 || n->head_kc % HKC != 0 ||
        n->head_kc < n->board_h * n->board_w * (n->channels / 8) || n->head_nt * n->head_ntiles < n->action_size + 3 ||
        !n->wtrunk || !n->cbias || !n->bn_scale || !n->bn_shift || !n->whead || !n->bhead || !n->gact ||
        (n->channels != 128 && !n->h_params) ||
        (!(n->head_ntiles == 1 && n->head_nt == 16) && !n->logits))
        return -1;
    cudaStream_t s = (cudaStream_t)stream;
    switch (n->precision) {
    case AZB_NN_BF16: return forward_prec<AZB_NN_BF16>(n, obs, policy, value, batch, rows, count, s, dump, dump_layer);
    case AZB_NN_F16: return forward_prec<AZB_NN_F16>(n, obs, policy, value, batch, rows, count, s, dump, dump_layer);
    case AZB_NN_BF16X2: return forward_prec<AZB_NN_BF16X2>(n, obs, policy, value, batch, rows, count, s, dump, dump_layer);
    case AZB_NN_F16X2: return forward_prec<AZB_NN_F16X2>(n, obs, policy, value, batch, rows, count, s, dump, dump_layer);
    default: return -1;
    }
}

}  // namespace g
}  // namespace

extern "C" int azb_nng_layout(int32_t channels, int32_t precision, int32_t *out)
{
    if (!out || (channels != 32 && channels != 64 && channels != 128) || precision < AZB_NN_BF16 || precision > AZB_NN_F16X2) return -1;
    const int parts = g::prec_split(precision) ? 2 : 1;
    if (channels == 128) {                                  /* k_trunk_wide: one slab = (tap, 32 input channels) */
        using W = g::WideCfg<AZB_NN_BF16>;
        out[0] = parts;
        out[1] = 0;                                         /* per-tap slabs */
        out[2] = parts * W::SLAB_PART;
        out[3] = 2 * W::TILES;
        out[4] = g::HKC;
        out[5] = W::MAXDEPTH;
        out[6] = W::LSLABS;
        out[7] = W::STEM_SLABS;
        return 0;
    }
    const int dys = channels == 32 ? 3 : 1, tiles = channels == 32 ? 7 : 2;
    out[0] = parts;
    out[1] = dys;                                           /* vertical taps per weight slab */
    out[2] = parts * dys * (channels / 8) * 3 * channels * 16;   /* slab bytes */
    out[3] = 2 * tiles;                                     /* boards per CTA */
    out[4] = g::HKC;                                        /* head K chunks per stage: head_kc is a multiple */
    out[5] = g::MAXD;
    out[6] = 3 / dys;                                       /* weight slabs per trunk layer */
    out[7] = 1;                                             /* weight slabs of the stem */
    return 0;
}

/* host-only: how k_trunk_tc deals the tiles of `boards` boards to its units -- see include/azb200_nn.h */
extern "C" int azb_nng_tile_plan(int32_t channels, int32_t boards, int32_t sms, int32_t pair, int32_t persist, int32_t unit, int32_t *out)
{
    if (!out || (channels != 32 && channels != 64) || boards < 0 || sms < 2) return -1;
    const g::UnitShare u = g::unit_share(boards, channels == 32 ? 7 : 2, sms, pair != 0, persist != 0, unit, channels == 32 ? 3 : 2);
    out[0] = u.units; out[1] = u.n; out[2] = u.first; out[3] = u.rounds; out[4] = u.tiles; out[5] = u.big;
    return 0;
}

extern "C" int azb_nng_forward(const azb_nng_net *net, const float *obs, float *policy, float *value, int32_t batch, const int32_t *rows,
                               const int32_t *count, void *stream)
{
    return g::forward(net, obs, policy, value, batch, rows, count, stream, nullptr, -1);
}

/* test / tuning hook: the pipeline trace of the last debug forward (see g_trace), n <= 1024 values */
extern "C" int azb_nng_trace(long long *out, int32_t n)
{
    if (!out || n < 0 || n > 4 * 256) return -7;
    return cudaMemcpyFromSymbol(out, g::g_trace, (size_t)n * sizeof(long long)) == cudaSuccess ? 0 : -2;
}

/* tuning hook: (SM id, globaltimer ns at entry, at exit) of the first n <= 2048 CTAs of the last debug forward */
extern "C" int azb_nng_cta_trace(long long *out, int32_t n)
{
    if (!out || n < 0 || n > 2048) return -7;
    return cudaMemcpyFromSymbol(out, g::g_cta_trace, (size_t)n * 3 * sizeof(long long)) == cudaSuccess ? 0 : -2;
}

extern "C" int azb_nng_forward_debug(const azb_nng_net *net, const float *obs, float *policy, float *value, int32_t batch, void *stream,
                                     float *dump, int32_t dump_layer)
{
    if (!dump) return -7;
    return g::forward(net, obs, policy, value, batch, nullptr, nullptr, stream, dump, dump_layer);
}
