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
/* Upload of a caller-owned PINNED host tensor (the batch_tensor / policy_tensor / value_tensor of the reference's
 * SelfPlayAgent protocol, SelfPlayAgent.pyx:14-16; NNetWrapper.process does `batch.cuda()`, NNetWrapper.py:227) by a
 * kernel that reads the mapped host memory over PCIe -- a small cudaMemcpyAsync host->device pays ~190 us of DMA
 * start-up on this platform.  Both pointers 16-byte aligned.  Asynchronous on `stream`; status codes as above. */
int azb_upload_pinned(void *dst_device, const void *src_pinned_host, int64_t bytes, void *stream);
int azb_nn_weight_row_stride(void);
int azb_nn_head_row_stride(void);
int azb_nn_boards_per_cta(void);

#ifdef __cplusplus
}
#endif
#endif
