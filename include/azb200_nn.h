/*
 * azb200_nn.h -- C ABI of the fused leaf evaluator in libazb200.so: the
 * reference's pre-activation ResNet (alphazero/NNetArchitecture.py:69-120, the
 * network NNetWrapper.process runs, alphazero/NNetWrapper.py:225-232) for small
 * boards as ONE kernel launch per batch (bf16 tensor-core convolutions, fp32
 * residual stream / heads / softmax).  Weights are prepared on the host with
 * batch-norm folded (azb200/fused_nn.py documents the algebra).  Supported
 * geometry: 6x7 boards, 32 trunk channels, 7 actions (Connect4); everything
 * else keeps the PyTorch/cuDNN evaluator.
 */
#ifndef AZB200_NN_H
#define AZB200_NN_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct azb_nn_weights {
    int32_t channels;      /* args.num_channels (32)                                          */
    int32_t depth;         /* args.depth: residual blocks                                     */
    int32_t in_channels;   /* Game.observation_size()[0]                                      */
    int32_t board_h, board_w;
    int32_t action_size;   /* Game.action_size()                                              */
    const void *wconv;     /* device bf16 [1+2*depth][channels][azb_nn_weight_row_stride()]:  */
                           /* row = output channel, k = tap*16+cin (stem) / tap*32+cin        */
    const float *cbias;    /* device f32 [1+2*depth][channels] (folded BN shift; 0 for conv2) */
    const float *bn_scale; /* device f32 [depth][channels]: BN1 of every block                */
    const float *bn_shift; /* device f32 [depth][channels]                                    */
    const void *whead;     /* device bf16 [16][azb_nn_head_row_stride()]: folded heads, row j =  */
                           /* output j (policy logits, then value logits), k = pos*channels+ch  */
    const float *bhead;    /* device f32 [16] (action_size+3 used)                               */
} azb_nn_weights;

/* obs: device f32 [batch, C, H, W]; policy: device f32 [batch, A] (probabilities);
 * value: device f32 [batch, 3].  Asynchronous on `stream`.  0 on success,
 * -1 unsupported geometry, -2 CUDA error, -7 bad argument. */
int azb_nn_forward(const azb_nn_weights *w, const float *obs, float *policy, float *value, int32_t batch, void *stream);

/* The same network on the 5th-generation tensor cores (tcgen05.mma, accumulators and the fp32
 * residual stream in TMEM; csrc/azb_resnet_tc.cu), 16 boards per CTA, in_channels <= 8.
 * Weight layouts differ from azb_nn_forward:
 *   wconv  bf16 [1+2*depth][azb_nn_tc_layer_bytes()/(3*channels*16)][3*channels][8]: K-major no-swizzle
 *          UMMA B operand with the three horizontal taps side by side, row n = (dx+1)*channels + cout,
 *          16-byte K chunk = (dy+1)*4 + cin/8 (trunk) or dy+1 with cin < 8 (stem)
 *   whead  bf16 [action_size+3][azb_nn_tc_head_row_stride()], k = frame_row*channels + ch with
 *          frame_row = y*8 + x over the azb_nn_tc_frame_rows_per_board() rows of a board frame
 * Same status codes. */
int azb_nn_forward_tc(const azb_nn_weights *w, const float *obs, float *policy, float *value, int32_t batch, void *stream);
/* Compact evaluation: boards rows[0 .. *count) of obs (device int32 row indices, device int32 count <= max_batch);
 * policy / value are written to the same rows, all other rows are left untouched.  The engine's select kernel lists
 * the leaves that need the network (azb_nn_rows_ptr / azb_nn_count_ptr in azb200.h): a terminal leaf's value is its
 * win state and the reference discards the network's answer for it (MCTS.pyx:234-235). */
int azb_nn_forward_tc_rows(const azb_nn_weights *w, const float *obs, float *policy, float *value, const int32_t *rows,
                           const int32_t *count, int32_t max_batch, void *stream);
/* test hook: also writes the fp32 activation produced by layer `dump_layer`'s epilogue to
 * dump[ceil(batch/16)*16][frame_rows][channels] (padding rows zero) */
int azb_nn_forward_tc_debug(const azb_nn_weights *w, const float *obs, float *policy, float *value, int32_t batch,
                            void *stream, float *dump, int32_t dump_layer);
int azb_nn_tc_layer_bytes(void);
int azb_nn_tc_head_row_stride(void);
int azb_nn_tc_boards_per_cta(void);
int azb_nn_tc_frame_rows_per_board(void);
/* ---- generic tcgen05 evaluator (csrc/azb_resnet_g.cu): boards up to 7x7, 32 / 64 / 128 trunk channels, any action
 * size, at the reference's numerics.  Replaces NNetWrapper.process (alphazero/NNetWrapper.py:225-232) for the ResNet of
 * alphazero/NNetArchitecture.py:69-120 with the presets of Coach.py:103-116 (32 ch), envs/hnefatafl/train_brandubh.py:50-55
 * (64 ch, 7x7, 588 actions), envs/connect4/train.py:44-49 (128 ch x 8 blocks).  Operand precision of the convolutions and the head GEMM (accumulation is always fp32):
 *   AZB_NN_BF16X2  every operand as hi + lo bf16 (16 significant bits; TF32 -- what cuDNN gives the reference -- has
 *                  11), a.w = a_hi.w_hi + a_hi.w_lo + a_lo.w_hi: probabilities within 1e-5 of the fp32 module.  Default.
 *   AZB_NN_F16X2   every operand as hi + lo fp16, same three products: 22 significant bits for |v| >= 0.03 (below, the
 *                  low part is a subnormal fp16: absolute error 6e-8), |v| saturates at 65504.  Default of the
 *                  128-channel network (envs/connect4/train.py:44-49), where 17 layers of K = 1152 leave BF16X2 at the
 *                  edge of 1e-5.
 *   AZB_NN_F16     fp16 operands (11 significant bits = TF32's), one pass; activations saturate at 65504.
 *   AZB_NN_BF16    bf16 operands (8 bits), one pass.                                                              */
enum { AZB_NN_BF16 = 0, AZB_NN_F16 = 1, AZB_NN_BF16X2 = 2, AZB_NN_F16X2 = 3 };

/* azb_nng_net.flags.  AZB_NNG_PAIR: the 32 / 64-channel trunk runs as clusters of two CTAs whose tiles advance in
 * lockstep as one M = 256 MMA stream (tcgen05 cta_group::2), each CTA staging half of the N rows of every weight chunk.
 * AZB_NNG_PERSIST / AZB_NNG_ONE_ROUND: how the 32 / 64-channel trunk deals its tiles -- ONE wave of resident CTAs (pairs)
 * that run their share in several rounds (set-up and tail of a CTA paid once), or whole waves of CTAs with one round
 * each.  Neither flag: the kernel's default (environment AZB_NNG_PERSIST=0 / 1, else persistent).  Same arithmetic:
 * outputs are bit-identical. */
enum { AZB_NNG_PAIR = 1, AZB_NNG_PERSIST = 2, AZB_NNG_ONE_ROUND = 4 };

typedef struct azb_nng_net {
    int32_t channels, depth, in_channels, board_h, board_w, action_size;
    int32_t precision;      /* AZB_NN_*                                                                       */
    int32_t max_boards;     /* capacity of gact / logits                                                      */
    int32_t head_nt;        /* head GEMM N tile (multiple of 16, <= 256)                                      */
    int32_t head_ntiles;    /* head_nt * head_ntiles >= action_size + 3                                       */
    int32_t head_kc;        /* head K in 8-element chunks, >= H*W*channels/8, multiple of layout[4]           */
    int32_t flags;          /* AZB_NNG_PAIR: wtrunk is laid out for CTA pairs (32 / 64 channels), see wtrunk;          */
                            /*   AZB_NNG_PERSIST / AZB_NNG_ONE_ROUND                                                   */
    const void *wtrunk;     /* device, slab stream: slab s at s * layout[2] bytes.  32 / 64 channels: slab 0 = stem    */
                            /*   [part][4 K chunks: dy=-1,0,+1,zero][3*channels][8 cin], the others, layer by     */
                            /*   layer, [part][dy in slab][cin/8][dx*channels + cout][8 cin]; part = hi, lo.       */
                            /*   128 channels: stem slabs dx = -1, 0, +1 as [part][4 K chunks: dy=-1,0,zero,+1]    */
                            /*   [cout][8 cin], then per layer 36 slabs (tap = 3(dy+1)+(dx+1), kq):                */
                            /*   [part][4 K chunks = cin/8 in 4kq..4kq+3][cout][8 cin]                             */
                            /*   AZB_NNG_PAIR: slab s of CTA rank r at (2 s + r) * layout[2] / 2 bytes, same layouts   */
                            /*   with the N rows [r * 3 channels / 2, (r + 1) * 3 channels / 2) of every K chunk      */
    const float *cbias;     /* device f32 [1+2*depth][channels] (folded BN shift; 0 for conv2)                 */
    const float *bn_scale;  /* device f32 [max(depth,1)][channels]: BN1 of every block                         */
    const float *bn_shift;
    const void *whead;      /* device [part][n tile][head_kc][head_nt][8]: folded heads, output j = tile*nt + row, */
                            /*   K chunk = (y*W + x)*channels/8 + ch/8                                           */
    const float *bhead;     /* device f32 [head_nt * head_ntiles]                                              */
    void *gact;             /* scratch, device: parts * ceil(max_boards/128) * head_kc * 2048 bytes             */
    float *logits;          /* scratch, device f32 [max_boards][head_nt*head_ntiles]; may be NULL when all outputs */
                            /*   fit one 16-wide tile                                                            */
    const float *h_params;  /* HOST copy of [cbias | bn_scale | bn_shift] (same shapes, concatenated): the 32 / 64-   */
                            /*   channel kernel takes them by value as a kernel argument (constant bank) -- read at  */
                            /*   every azb_nng_forward call; required for 32 / 64 channels, unused for 128           */
} azb_nng_net;

/* out[0..7] = operand parts, vertical taps per weight slab (0: the 128-channel kernel, one slab per tap and 32 input
 * channels), slab bytes, boards per CTA, head K-chunk granularity, max depth, weight slabs per trunk layer, weight slabs
 * of the stem.  -1: unsupported channels / precision. */
int azb_nng_layout(int32_t channels, int32_t precision, int32_t *out);
/* Host only, no device needed: how the 32 / 64-channel trunk deals the tiles (two boards each; pair: pairs of tiles) of
 * `boards` boards to its units (CTAs; pair: CTA pairs) on `sms` SMs.  out[0..5] = number of units, then for `unit`: its
 * number of tiles, its first tile, its rounds, the tiles of a big round, the number of big rounds (the other rounds have
 * one tile less: big * tiles + (rounds - big) * (tiles - 1) == number of tiles; when that would leave an MMA-issuing
 * thread without a tile, big == rounds and rounds * tiles >= number of tiles, the surplus being empty tiles).  All 0 for
 * a unit beyond the grid.  -1: bad argument. */
int azb_nng_tile_plan(int32_t channels, int32_t boards, int32_t sms, int32_t pair, int32_t persist, int32_t unit, int32_t *out);
/* obs f32 [batch, C, H, W] -> policy f32 [batch, A], value f32 [batch, 3] (probabilities).  rows / count (both or
 * neither): compact evaluation of boards rows[0 .. *count) as azb_nn_forward_tc_rows, batch = upper bound.
 * Asynchronous on `stream` (two or three kernel launches).  Status codes as azb_nn_forward. */
int azb_nng_forward(const azb_nng_net *net, const float *obs, float *policy, float *value, int32_t batch, const int32_t *rows,
                    const int32_t *count, void *stream);
/* test hook: also writes the activation layer `dump_layer`'s epilogue hands to the next layer (operand rounding
 * included) to dump f32 [batch][64 frame rows][channels] (padding rows zero) */
int azb_nng_forward_debug(const azb_nng_net *net, const float *obs, float *policy, float *value, int32_t batch, void *stream,
                          float *dump, int32_t dump_layer);

/* tuning hook: clock64 stamps of CTA 0's pipeline in the last azb_nng_forward_debug call, four per (layer, tile) in issue
 * order: MMAs start to issue, MMAs committed, epilogue sees the accumulator, epilogue hands the tile on.  n <= 1024. */
int azb_nng_trace(long long *out, int32_t n);
/* tuning hook: (SM id, %globaltimer at entry, at exit) of the first n <= 2048 CTAs of the last azb_nng_forward_debug call */
int azb_nng_cta_trace(long long *out, int32_t n);

/* Upload of a caller-owned PINNED host tensor (the batch_tensor / policy_tensor / value_tensor of the reference's
 * SelfPlayAgent protocol, SelfPlayAgent.pyx:14-16; NNetWrapper.process does `batch.cuda()`, NNetWrapper.py:227) by a
 * kernel that reads the mapped host memory over PCIe -- a small cudaMemcpyAsync host->device pays ~190 us of DMA
 * start-up on this platform.  The same kernel serves the other direction (dst = the pinned host tensor's device mapping,
 * src = device memory: posted PCIe writes), which azb200.nnet.download uses for the batch / policy / value tensors the
 * protocol returns to the host.  Both pointers 16-byte aligned.  Asynchronous on `stream`; status codes as above. */
int azb_upload_pinned(void *dst_device, const void *src_pinned_host, int64_t bytes, void *stream);
int azb_nn_weight_row_stride(void);
int azb_nn_head_row_stride(void);
int azb_nn_boards_per_cta(void);

#ifdef __cplusplus
}
#endif
#endif
