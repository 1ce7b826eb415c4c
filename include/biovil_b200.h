/*
 * biovil_b200.h - C ABI of the B200-native BioViL image-encoder + prompt-scorer hot path.
 *
 * The reference (marcomistretta/incremental_multimodal_medical_learning_II) is pure Python and has no FFI:
 * its "operator interface" for this path is the Python class surface of health_multimodal.image.  This
 * library sits one level below a Python mirror of that surface (incremental_multimodal_medical_learning_ii_b200/
 * image/*.py, loaded with ctypes) and replaces, per entry point:
 *
 *   bv_forward        ImageModel.forward                         health_multimodal/image/model/model.py:141-154
 *                     ImageEncoder.forward                       health_multimodal/image/model/model.py:197-205
 *                     ResNetHIML.forward (+ torchvision Bottleneck) health_multimodal/image/model/resnet.py:25-47
 *                     MLP.forward (projector)                    health_multimodal/image/model/modules.py:43-55
 *                     get_patchwise_projected_embeddings         health_multimodal/image/model/model.py:161-175
 *                     patch x prompt similarity map              health_multimodal/vlp/inference_engine.py:93-108
 *   bv_forward_graph  the same forward, captured once per argument set as a CUDA graph and replayed (the reference's
 *                     extraction loop calls the model with batch size 1: chexpert-get-embedding.py:47-49, 68-80)
 *   bv_quantize_frames_f32  undoes ToTensor + ExpandChannels on 8-bit data   health_multimodal/image/data/transforms.py:12-38
 *   bv_set_prompts    Trainer.bert_forward_mean (prompt side)    Trainer.py:1657-1680
 *   bv_score          Trainer.myCosineSimilarity + label loop    Trainer.py:1682-1704, 805-837, 1019-1047
 *   bv_pairwise_cosine  Trainer.myCosineSimilarity alone (one call, no state)   Trainer.py:1682-1704
 *   bv_jpeg_info / bv_jpeg_decode_gray_u8   read_image of a grey JPEG (nvJPEG)   DataRetrieval.py:70-96
 *   bv_resize_center_crop_u8  transforms.Resize + CenterCrop on 8-bit frames (Pillow 8bpc bilinear, bit-exact)
 *                     DataRetrieval.py:175-180; health_multimodal/image/data/transforms.py:30-41
 *   bv_smooth_heatmaps  gaussian_filter(sigma) of the similarity maps health_multimodal/vlp/inference_engine.py:107-109
 *   bv_heatmaps_to_image_size  convert_similarity_to_image_size (nearest upsample + NaN pad)  health_multimodal/vlp/inference_engine.py:113-155
 *   bv_set_profile / bv_get_profile   (measurement only; no reference counterpart)
 *   bv_conv2d_nhwc    one Conv2d+BatchNorm2d(+ReLU)(+residual)   (unit-test entry for the tcgen05 kernel)
 *   bv_conv_chain_nhwc  Bottleneck tail (conv3+bn3+identity/downsample+ReLU) chained with the next Bottleneck's
 *                     conv1+bn1+ReLU in one kernel               (unit-test entry; torchvision Bottleneck.forward under
 *                                                                 health_multimodal/image/model/resnet.py:38-42)
 *
 * Conventions: plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 * the caller owns all buffers and the stream; calls are asynchronous on `stream`; functions return 0 on
 * success or a negative bv_status and never throw; bv_last_error() gives the message of the calling
 * thread's last failure.  One handle per device; a handle is not thread-safe.  There is no CPU fallback:
 * on a machine without an sm_100 GPU every compute entry point fails with BV_ERR_NO_DEVICE.
 */
#ifndef BIOVIL_B200_H_
#define BIOVIL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bv_handle bv_handle;
typedef void* bv_stream; /* cudaStream_t */

typedef enum {
    BV_OK = 0,
    BV_ERR_INVALID = -1,   /* bad argument / unsupported shape */
    BV_ERR_NO_DEVICE = -2, /* no CUDA device of compute capability 10.x */
    BV_ERR_CUDA = -3,      /* a CUDA runtime / driver call failed */
    BV_ERR_WORKSPACE = -4  /* workspace too small */
} bv_status;

typedef enum { BV_DTYPE_U8 = 0, BV_DTYPE_F32 = 1 } bv_dtype;

/* One convolution with eval-mode BatchNorm folded in: w is bf16 [cout][r][s][cin] (K-major), bias fp32 [cout]. */
typedef struct {
    const void* w;
    const float* bias;
    int32_t cin, cout, r, s, stride, pad;
} bv_conv;

#define BV_NUM_BLOCKS 16 /* Bottleneck blocks, layers [3,4,6,3] (resnet.py:80) */

typedef struct {
    /* Stem conv7x7/2 as a [64][K] matrix over gathered patches (k = c*49 + r*7 + s, zero padded):
     *   stem_u8 : 1 input channel (the 3 identical ExpandChannels copies summed), 1/255 folded in, K = 64
     *   stem_f1 : 1 input channel, float frames,                                              K = 64
     *   stem_f3 : 3 input channels, float frames (general ImageModel.forward input),          K = 192    */
    bv_conv stem_u8, stem_f1, stem_f3;
    bv_conv conv1[BV_NUM_BLOCKS], conv2[BV_NUM_BLOCKS], conv3[BV_NUM_BLOCKS];
    bv_conv downsample[BV_NUM_BLOCKS]; /* w == NULL when the block has no downsample branch */
    bv_conv proj0;                     /* projector conv 2048->128 (+BN folded), ReLU */
    const float* proj3_wt;             /* projector conv 128->128 TRANSPOSED: fp32 [k][d] */
    const float* proj3_b;              /* fp32 [128] */
    bv_conv stem_u8_k8;                /* stem for the fused 8-bit kernel: [64][64], k = r*8 + s (s = 7 and r = 7 zero),
                                          1/255 folded in; w == NULL -> bv_forward uses gather + GEMM + max-pool */
} bv_weights;

/* Optional outputs of bv_forward; any pointer may be NULL. */
typedef struct {
    float* global_emb;      /* [B,128]    un-normalised projected global embedding (what the fork's forward returns) */
    float* patch_emb;       /* [B,H/32,W/32,128] projected patch embeddings, channel-last */
    int32_t normalize_patch;/* L2-normalise patch_emb over the last dim (F.normalize, eps 1e-12) */
    float* pooled;          /* [B,2048]   img_embedding: global average pool of the trunk output */
    void* trunk_nhwc_bf16;  /* [B,H/32,W/32,2048] bf16 copy of the trunk output (patch_embedding) */
    float* sim;             /* [B,L,2]    (pos,neg) cosine vs the prompts installed with bv_set_prompts */
    float* prob;            /* [B,L]      sigmoid(pos-neg) */
    uint8_t* pred;          /* [B,L]      1 iff pos > neg */
    float* score;           /* [B,L]      (pos+1)/2 */
    float* heat;            /* [B,H/32,W/32,L] normalised-patch . unit positive prompt */
} bv_outputs;

const char* bv_last_error(void);
const char* bv_version(void);

/* Host-side shape helpers (no GPU needed). */
size_t bv_workspace_bytes(int32_t batch, int32_t channels, int32_t height, int32_t width);
int32_t bv_patch_grid(int32_t size); /* size / 32 */

/* `host_weights` is a host struct of DEVICE pointers (packed bf16 weights, fp32 biases).  The pointers must stay valid and
 * their contents unchanged for the life of the handle: bv_create synchronises the device once and keeps host copies of
 * the bias vectors, which the convolution kernels receive by value (constant bank) - create a new handle after changing
 * parameters (what ImageModel does when a parameter's version counter moves). */
int32_t bv_create(bv_handle** out, const bv_weights* host_weights, int32_t device);
void bv_destroy(bv_handle* h);

/* prompts: fp32 [L][2][P][128], index 0 = positive, 1 = negative, un-normalised (mean over prompts already applied
 * by the caller when the reduction is "mean": then P == 1; P > 1 = max over per-prompt cosines).
 * heat_text: optional fp32 [L][128] text vectors for the patch heat-maps (NULL -> positive prompt 0 of each label). */
int32_t bv_set_prompts(bv_handle* h, const float* prompts, int32_t num_labels, int32_t prompts_per_polarity,
                       const float* heat_text, bv_stream stream);

int32_t bv_forward(bv_handle* h, const void* frames, int32_t dtype, int32_t batch, int32_t channels, int32_t height,
                   int32_t width, void* workspace, size_t workspace_bytes, const bv_outputs* host_out,
                   bv_stream stream);

/* bv_forward recorded once into a CUDA graph per distinct argument set (frame / workspace / output pointers, shape,
 * installed prompt buffers) and replayed with ONE cudaGraphLaunch afterwards: same arguments, same results.  For the
 * launch-bound regime (the reference's batch-size-1 loop); callers must pass the SAME buffers again to hit the cache
 * (8 entries, oldest dropped).  Falls back to direct launches while profiling is on. */
int32_t bv_forward_graph(bv_handle* h, const void* frames, int32_t dtype, int32_t batch, int32_t channels, int32_t height,
                         int32_t width, void* workspace, size_t workspace_bytes, const bv_outputs* host_out,
                         bv_stream stream);

/* Float frames x [B,C,H,W] (C = 1 or 3) -> out u8 [B,1,H,W] = round(255 x) of channel 0; *bad (device int32, zeroed
 * here) becomes non-zero when some pixel is NOT 8-bit data (|255 x - k| > 1e-3, k outside 0..255, NaN) or the channels
 * differ: the caller then feeds the float frames to bv_forward instead.  H*W multiple of 4, x 16-byte aligned. */
int32_t bv_quantize_frames_f32(const float* x, int32_t batch, int32_t channels, int32_t height, int32_t width,
                               uint8_t* out, int32_t* bad, bv_stream stream);

/* Score cached embeddings emb [B,128] (un-normalised) against the installed prompts. */
int32_t bv_score(bv_handle* h, const float* emb, int32_t batch, float* sim, float* prob, uint8_t* pred, float* score,
                 bv_stream stream);

/* out[b][p] = cosine(x[b], y[p]) for x [B,128], y [P,128] (both un-normalised; rows divided by their L2 norm, no eps:
 * torchmetrics.pairwise_cosine_similarity as Trainer.myCosineSimilarity calls it); reduce_max != 0 -> out [B] = max over p
 * (the MAX_EMB branch, Trainer.py:1691-1694).  Stateless: no handle, no allocation, asynchronous on `stream`. */
int32_t bv_pairwise_cosine(const float* x, const float* y, int32_t batch, int32_t num_prompts, int32_t reduce_max,
                           float* out, bv_stream stream);

/* JPEG decode in front of the resize kernel (nvJPEG, resolved with dlopen at first use).  bv_jpeg_info parses the
 * header of a HOST buffer; bv_jpeg_decode_gray_u8 decodes it into the DEVICE buffer out [height][pitch] (luma plane =
 * the grey values of a single-component JPEG).  Not bit-exact against libjpeg (IDCT rounding): within +-2 grey levels. */
int32_t bv_jpeg_info(const uint8_t* host_data, size_t length, int32_t* width, int32_t* height, int32_t* components);
int32_t bv_jpeg_decode_gray_u8(const uint8_t* host_data, size_t length, uint8_t* out, int32_t width, int32_t height,
                               int32_t pitch, bv_stream stream);
/* n streams in ONE call through nvjpegDecodeBatched: backend 3 = hardware JPEG engines, 2 = GPU-assisted Huffman decode,
 * 0 = try 3 then 2.  outs[i] / pitches[i]: device buffers sized from bv_jpeg_info.  Returns the backend used (> 0) or a
 * negative bv_status when no batched backend takes the batch (callers then use bv_jpeg_decode_gray_u8 per image). */
int32_t bv_jpeg_decode_batch_gray_u8(const uint8_t* const* host_datas, const size_t* lengths, int32_t n, uint8_t* const* outs,
                                     const int32_t* pitches, int32_t backend, bv_stream stream);

/* Resize(size) -> CenterCrop(crop) of n same-sized 8-bit grayscale frames src [n,h,w] -> out [n,crop,crop] with the
 * integer arithmetic of Pillow's 8bpc bilinear resampler (short side -> size, long side int(size*long/short), centre
 * crop offsets int(round((dim-crop)/2))).  workspace: bv_resize_workspace_bytes(...) bytes of device memory. */
size_t bv_resize_workspace_bytes(int32_t n, int32_t height, int32_t width, int32_t size, int32_t crop);
int32_t bv_resize_center_crop_u8(const uint8_t* src, int32_t n, int32_t height, int32_t width, int32_t size,
                                 int32_t crop, uint8_t* out, void* workspace, size_t workspace_bytes, bv_stream stream);

/* Smooth patch-similarity maps heat [B,gh,gw,L] -> out [B,gh,gw,L] with scipy.ndimage.gaussian_filter semantics
 * (order 0, mode 'reflect', truncate 4.0, separable, same sigma on both axes); gh*gw <= 1024, radius <= 16. */
int32_t bv_smooth_heatmaps(const float* heat, int32_t batch, int32_t grid_h, int32_t grid_w, int32_t num_labels,
                           float sigma, float* out, bv_stream stream);

/* Patch-grid similarity maps heat [B,gh,gw,L] -> out [B,L,height,width] in the ORIGINAL image's pixels, as
 * ImageTextInferenceEngine.convert_similarity_to_image_size does with interpolation="nearest" (its default;
 * health_multimodal/vlp/inference_engine.py:113-155): the grid is stretched (F.interpolate nearest) over the centre-crop
 * square of int(crop_size * min(height, width) / resize_size) pixels (crop_size pixels when resize_size == 0; the whole
 * image when crop_size == 0) and everything outside it is NaN (F.pad).  B * L <= 65535 maps per call. */
int32_t bv_heatmaps_to_image_size(const float* heat, int32_t batch, int32_t grid_h, int32_t grid_w, int32_t num_labels,
                                  int32_t height, int32_t width, int32_t resize_size, int32_t crop_size, float* out,
                                  bv_stream stream);

/* Number of kernels the last bv_forward launched (for launch accounting in the benchmark). */
int32_t bv_last_forward_launches(const bv_handle* h);

/* Per-launch timing of bv_forward with cudaEvents recorded on the launch stream (measurement only; adds one
 * event record per launch).  bv_get_profile synchronises on the last event and fills up to `capacity` records of
 * the most recent bv_forward; returns the number of launches or a negative bv_status. */
typedef struct {
    char name[64];
    double flops; /* 2*M*N*K actually issued to the tensor cores (0 for non-GEMM kernels) */
    double bytes; /* algorithmic HBM bytes of the launch */
    float ms;
} bv_launch_info;
int32_t bv_set_profile(bv_handle* h, int32_t enable);
int32_t bv_get_profile(bv_handle* h, bv_launch_info* host_out, int32_t capacity);

/* One fused convolution on NHWC bf16 through the tcgen05 implicit-GEMM kernel:
 *   out[B,Ho,Wo,cout] = act(conv(x, c) + c.bias (+ conv(x2, c2) + c2.bias) (+ residual))
 * x2/c2 (optional, may be NULL) is a second operand pair accumulated into the same tile (the fused
 * downsample branch); residual is bf16 [B,Ho,Wo,cout] or NULL; out is bf16, or fp32 when out_fp32 != 0. */
int32_t bv_conv2d_nhwc(const void* x, int32_t batch, int32_t height, int32_t width, const bv_conv* host_c,
                       const void* x2, int32_t height2, int32_t width2, const bv_conv* host_c2, const void* residual,
                       int32_t relu, void* out, int32_t out_fp32, bv_stream stream);

/* Two chained convolutions through one kernel (chain_gemm.cuh):
 *   out1[B,Ho,Wo,N1] = relu(conv(x, c) + c.bias (+ conv(x2, c2) + c2.bias | + residual))      bf16
 *   out2[B,Ho,Wo,N2] = relu(conv1x1(out1, next) + next.bias)                                   bf16
 * `next` must be a 1x1 stride-1 convolution with cin == c.cout; c.cout a multiple of 128 (<= 512); next.cout in
 * {64, 128, 256}.  x2/c2 and residual are mutually exclusive; either may be NULL. */
int32_t bv_conv_chain_nhwc(const void* x, int32_t batch, int32_t height, int32_t width, const bv_conv* host_c,
                           const void* x2, int32_t height2, int32_t width2, const bv_conv* host_c2,
                           const void* residual, void* out1, const bv_conv* host_next, void* out2, bv_stream stream);

/* The same chain on CTA pairs (pair_chain.cuh; tcgen05 cta_group::2) for the deep layers' identity blocks:
 *   out1[B,H,W,N1] = relu(conv1x1(x, c) + c.bias + residual),  out2[B,H,W,N2] = relu(conv1x1(out1, next) + next.bias)
 * c.cin in {128, 256}, c.cout a multiple of 128, next.cout in {128, 256}; residual is required. */
int32_t bv_pair_chain_nhwc(const void* x, int32_t batch, int32_t height, int32_t width, const bv_conv* host_c,
                           const void* residual, void* out1, const bv_conv* host_next, void* out2, bv_stream stream);

/* Fused layer1 Bottleneck tail on CTA pairs (l1_block.cuh), unit-test entry:
 *   t2   = relu(conv3x3(t1, c2) + c2.bias)                       (64 -> 64 channels, stride 1, pad 1; never stored)
 *   out1 = relu(conv1x1(t2, c3) + c3.bias + residual)  [B,H,W,256] bf16
 *   out2 = relu(conv1x1(out1, next) + next.bias)       [B,H,W,64]  bf16
 * Replaces Bottleneck.conv2/bn2/relu, conv3/bn3/+identity/relu and the next Bottleneck's conv1/bn1/relu
 * (torchvision Bottleneck.forward under health_multimodal/image/model/resnet.py:39). */
int32_t bv_l1_block_nhwc(const void* t1, int32_t batch, int32_t height, int32_t width, const bv_conv* host_c2,
                         const bv_conv* host_c3, const void* residual, void* out1, const bv_conv* host_next, void* out2,
                         bv_stream stream);

/* The same kernel for the FIRST block of the layer (downsample branch instead of an identity, resnet.py:39 with
 * Bottleneck.downsample): out1 = relu(conv1x1(t2, c3) + c3.bias + conv1x1(x0, ds) + ds.bias), x0 [B,H,W,64] the block
 * input, ds a 64 -> 256 1x1 stride-1 bv_conv; everything else as bv_l1_block_nhwc. */
int32_t bv_l1_block_ds_nhwc(const void* t1, int32_t batch, int32_t height, int32_t width, const bv_conv* host_c2,
                            const bv_conv* host_c3, const void* x0, const bv_conv* host_ds, void* out1,
                            const bv_conv* host_next, void* out2, bv_stream stream);

/* Validation vehicle for the CTA-pair (tcgen05 cta_group::2) building blocks in csrc/pair_gemm.cuh:
 *   out[M,N] (fp32) = a[M,K] (bf16) x w[N,K]^T (bf16);  N multiple of 32 in [32,256], K multiple of 64. */
int32_t bv_pair_gemm_test(const void* a, const void* w, int32_t m, int32_t n, int32_t k, float* out, bv_stream stream);

/* The stem alone on 8-bit frames (conv1 7x7/2 -> bn1 -> relu -> maxpool, health_multimodal/image/model/resnet.py:34-37):
 *   frames u8 [B][H][W] (H, W multiples of 32), w8 = bv_weights.stem_u8_k8, out bf16 NHWC [B][H/4][W/4][64].
 *   variant 0 = row-streaming kernel (csrc/stem_rows.cuh, what bv_forward uses), 1 = tile kernel (csrc/stem_fused.cuh).
 * Kernel-level parity tests call this; bv_forward runs the same launches. */
int32_t bv_stem_u8_nhwc(const void* frames, int32_t batch, int32_t height, int32_t width, const bv_conv* host_w8,
                        void* out, int32_t variant, bv_stream stream);

/* The row-streaming stem with layer1.0's conv1 fused in (torchvision Bottleneck.forward conv1 -> bn1 -> relu on the
 * max-pool output, resnet.py:38-39): out = max-pool output as above, out1 = relu(conv1x1(out) + bias) bf16 NHWC
 * [B][H/4][W/4][64]; host_c1 is a 64 -> 64 1x1 bv_conv.  Experiment entry: bv_forward launches this form only with
 * BV_STEM_C1=1 (measured slower than stem + separate conv1, DESIGN.md switches table); the default is bv_stem_u8_nhwc. */
int32_t bv_stem_conv1_u8_nhwc(const void* frames, int32_t batch, int32_t height, int32_t width, const bv_conv* host_w8,
                              const bv_conv* host_c1, void* out, void* out1, bv_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* BIOVIL_B200_H_ */
