/* monorun_head.h -- C ABI of libmonorun_head.so: the dense correspondence head of MonoRUn on B200 (sm_100a).
 *
 * Replaces, for inference, FCNNOCDecoder.forward up to slice_pred
 *   (monorun/models/roi_heads/bbox_3d_heads/dense_decoders/fcn_noc_decoder.py:189-235):
 *   3 x [conv3x3 256->256 + ReLU] @14x14  ->  x += Linear(16->256)(latent)  ->  CARAFEPack(256, x2, k_up 5,
 *   encoder 3x3, compressed 64)  ->  [conv3x3 256->256 + ReLU] @28x28  ->  conv1x1 256->(3+2) C 2
 * which the reference runs as cuDNN / mmcv launches.  Here every convolution is an implicit GEMM on the 5th-gen
 * tensor cores (tcgen05.mma, bf16 operands, fp32 accumulators in TMEM, operands staged by TMA), and the CARAFE
 * pixel-shuffle + softmax + reassembly is one bandwidth-bound kernel.  The output is the head's unsliced
 * all_pred [N, cout, 28, 28] fp32, the tensor mrpnp_solve_dense (monorun_pnp.h) consumes with num_classes > 0.
 *
 * All pointers are DEVICE pointers unless stated otherwise; the caller owns every buffer; calls are asynchronous
 * on the given stream.  No torch types cross this boundary.  Return 0 on success, negative on error
 * (mrhead_last_error() has the message).
 */
#ifndef MONORUN_HEAD_H
#define MONORUN_HEAD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRHEAD_VERSION 1
#define MRHEAD_OK 0
#define MRHEAD_ERR_ARG (-1)
#define MRHEAD_ERR_CUDA (-2)

#define MRHEAD_OUT_BF16_ROWS 0   /* bf16 [n (h+2)(w+2), cout]      padded-flat NHWC, halo rows zero             */
#define MRHEAD_OUT_F32_ROWS 1    /* fp32 [n (h+2)(w+2), cout_pad]  padded-flat                                  */
#define MRHEAD_OUT_F32_PLANAR 2  /* fp32 [n, cout, h, w]           NCHW                                         */

typedef struct mrhead_ctx mrhead_ctx;

/* One convolution layer.  `weight` is the layer's nn.Conv2d weight [cout, cin, kh, kw] re-packed ONCE by the caller
 * to bf16 [taps, cout_pad, cin] (tap = ky * kw + kx; rows cout..cout_pad-1 zero): each tap is a K-major GEMM B tile.
 * cin: multiple of 64.  cout_pad: multiple of 16, <= 256.  taps: 1 or 9.  bias: fp32 [cout] or NULL.            */
typedef struct mrhead_layer {
    const void* weight;
    const float* bias;
    int32_t cin, cout, cout_pad, taps, relu;
} mrhead_layer;

/* The whole head (fcn_noc_decoder.py:93-157).  num_convs <= 4, num_convs_up <= 2.                               */
typedef struct mrhead_weights {
    mrhead_layer convs[4];
    int32_t num_convs;
    const float* latent_w;      /* fp32 [256, latent_channels] (nn.Linear weight) or NULL = no latent vector         */
    const float* latent_b;      /* fp32 [256] or NULL                                                                */
    int32_t latent_channels;
    int32_t latent_activation;  /* 0 none, 1 ReLU, 2 LeakyReLU(0.01)  (fcn_noc_decoder.py:84-91)                     */
    mrhead_layer compressor;    /* CARAFEPack.channel_compressor, 1x1 256 -> 64                                      */
    mrhead_layer encoder;       /* CARAFEPack.content_encoder, 3x3 64 -> 25 * 4                                      */
    mrhead_layer convs_up[2];
    int32_t num_convs_up;
    mrhead_layer final;         /* conv_final, 1x1 256 -> (3+2) C 2                                                  */
} mrhead_weights;

int mrhead_create(mrhead_ctx** out, int device);
void mrhead_destroy(mrhead_ctx* ctx);
int mrhead_version(void);
const char* mrhead_last_error(void);
int64_t mrhead_launch_count(const mrhead_ctx* ctx);   /* kernels of this library launched so far */

/* x fp32 [n, c, h, w] (RoI features, NCHW) -> bf16 padded-flat [n (h+2)(w+2), c] with a zero halo. */
int mrhead_pack_input(mrhead_ctx* ctx, const float* x, int n, int c, int h, int w, void* out, void* stream);

/* One convolution on a padded-flat bf16 activation `in` [n (h+2)(w+2), cin].  row_bias: fp32 [n, cout] added after the
 * activation, or NULL.  out layout per out_mode (MRHEAD_OUT_*).  3x3 = padding 1, stride 1; 1x1 = padding 0.        */
int mrhead_conv(mrhead_ctx* ctx, const mrhead_layer* layer, const void* in, int n, int h, int w, const float* row_bias,
                int out_mode, void* out, void* stream);

/* row_bias [n, cout] = b + act(latent [n, k]) W^T */
int mrhead_latent_bias(mrhead_ctx* ctx, const float* latent, const float* w, const float* b, int n, int k, int cout,
                       int activation, float* out, void* stream);

/* CARAFE pixel-shuffle + softmax + reassembly (k_up = 5, scale 2, group 1, 256 channels):
 * feat bf16 padded-flat [n (h+2)(w+2), 256], logits fp32 padded-flat [n (h+2)(w+2), ld_logits] (channels 0..99 used)
 * -> out bf16 padded-flat [n (2h+2)(2w+2), 256].                                                                   */
int mrhead_carafe(mrhead_ctx* ctx, const void* feat, const float* logits, int ld_logits, int n, int h, int w, void* out,
                  void* stream);

/* Bytes of device scratch mrhead_forward needs for n RoIs of h x w (14 x 14) features. */
size_t mrhead_workspace_bytes(const mrhead_weights* wts, int n, int h, int w);

/* The whole head: x fp32 [n, 256, h, w], latent fp32 [n, latent_channels] or NULL  ->  all_pred fp32
 * [n, final.cout, 2h, 2w].  workspace: device scratch of at least mrhead_workspace_bytes(...), 1024-byte aligned.   */
int mrhead_forward(mrhead_ctx* ctx, const mrhead_weights* wts, const float* x, const float* latent, int n, int h, int w,
                   void* workspace, size_t workspace_bytes, float* all_pred, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MONORUN_HEAD_H */
