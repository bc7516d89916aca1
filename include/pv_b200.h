/*
 * pv_b200.h - C ABI of libpvb200.so, the B200 (sm_100a) forward-path engine for the
 * PyTorchVideo hot path.
 *
 * Every entry point takes plain device pointers + sizes + an explicit cudaStream_t (passed as
 * void* so that the header needs no CUDA include), allocates nothing, never throws, and returns
 * 0 on success or a negative pv_status.  pv_last_error() returns a thread-local message.
 *
 * Activations are "NDHWC" (channels-last-3d): element (n,t,h,w,c) of a tensor lives at
 *   base + (((n*T + t)*H + h)*W + w) * row_stride + c
 * where row_stride >= C is the distance (in elements) between consecutive positions; a tensor
 * may therefore be a channel slice of a wider buffer (this is how torch.cat(dim=1) is fused
 * away).  Channel counts of internal activations are padded to a multiple of 8 and the pad
 * lanes are kept at exactly zero.
 *
 * Each function cites the reference call site(s) (facebookresearch/pytorchvideo @ f3142bb,
 * paths relative to the reference root) whose arithmetic it replaces.
 */
#ifndef PV_B200_H_
#define PV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PV_ABI_VERSION 1

typedef enum pv_status {
  PV_OK = 0,
  PV_ERR_INVALID = -1,     /* bad argument / unsupported shape                       */
  PV_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed                    */
  PV_ERR_UNSUPPORTED = -3, /* no kernel for this configuration (never falls back)    */
  PV_ERR_NO_DEVICE = -4    /* no sm_100 device visible                               */
} pv_status;

typedef enum pv_dtype { PV_F16 = 0, PV_F32 = 1, PV_U8 = 2 } pv_dtype;

typedef enum pv_act {
  PV_ACT_NONE = 0,
  PV_ACT_RELU = 1,
  PV_ACT_SWISH = 2,   /* x*sigmoid(x)        reference layers/swish.py:25-28     */
  PV_ACT_GELU = 3,    /* exact erf GELU      reference layers/attention.py:74   */
  PV_ACT_SIGMOID = 4
} pv_act;

typedef enum pv_conv_algo {
  PV_ALGO_AUTO = 0,
  PV_ALGO_DIRECT = 1,   /* CUDA-core tiled direct convolution (any shape, f16 or f32 storage) */
  PV_ALGO_TCGEN05 = 2   /* TMA + tcgen05.mma implicit GEMM (f16 storage, fp32 TMEM accum)    */
} pv_conv_algo;

typedef enum pv_pool_mode { PV_POOL_MAX = 0, PV_POOL_AVG = 1 } pv_pool_mode;

/* ---------------------------------------------------------------------------------------------
 * Library / device
 * ------------------------------------------------------------------------------------------- */
int pv_abi_version(void);
const char* pv_last_error(void);
/* Fills sm_count / cc (major*10+minor); returns PV_ERR_NO_DEVICE when no GPU is visible. */
int pv_device_info(int* sm_count, int* cc);
/* Number of kernels launched by this library since load (bench.py "gpu_launches"). */
long long pv_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Clip transform chain, fused (one kernel):
 *   UniformTemporalSubsample -> Div255 -> Normalize -> ShortSideScale(bilinear) -> crop
 * Replaces: transforms/functional.py:19-41 (uniform_temporal_subsample: idx_t is computed on the
 * host with torch.linspace exactly as functional.py:39-40 and passed in), functional.py:604-615
 * (div_255), transforms/transforms.py:177-195 (Normalize), functional.py:92-131
 * (short_side_scale -> F.interpolate bilinear, align_corners=False), torchvision CenterCrop /
 * RandomCrop / functional.py:302-347 (uniform_crop) window selection.
 *
 * src is a uint8 (or f32/f16) clip addressed as src[c*sc + t*st + h*sh + w*sw], so both the
 * CTHW-contiguous layout and the decoder's THWC-interleaved layout (data/utils.py:26-31) work.
 * The bilinear taps are host-computed tables (bit-exact w.r.t. ATen's index/lambda math):
 *   for output row y:  rows y0[y], y1[y], weight ly[y] of row y1   (already offset by the crop)
 *   for output col x:  cols x0[x], x1[x], weight lx[x] of col x1
 * out[c][j][y][x] = ly0*(lx0*v00 + lx1*v01) + ly1*(lx0*v10 + lx1*v11),
 *   v = (float(src)/255 - mean[c]) / std[c], taken from frame idx_t[j];  dst is [C,n_t,oh,ow].
 * With identity tables / mean 0 / std 1 / div255 0 the same kernel is each single transform
 * (UniformTemporalSubsample, Div255, Normalize, ShortSideScale, crop) on its own.
 * ------------------------------------------------------------------------------------------- */
typedef struct pv_clip_transform_desc {
  int C, n_t, out_h, out_w;
  long long sc, st, sh, sw;   /* source strides in ELEMENTS of the source dtype    */
  float mean[4], stdv[4];     /* per channel (C <= 4)                             */
  int src_dtype;              /* PV_U8 (decoder frames), PV_F32 or PV_F16         */
  int dst_dtype;              /* PV_F16 or PV_F32                                 */
  int div255;                 /* 1: v = src/255 first (Div255); 0: v = src        */
} pv_clip_transform_desc;

int pv_clip_transform_fwd(const pv_clip_transform_desc* d, const void* src,
                          const int32_t* idx_t,
                          const int32_t* y0, const int32_t* y1, const float* ly,
                          const int32_t* x0, const int32_t* x1, const float* lx,
                          void* dst, void* stream);

/* Batched chain: n_clips clips per launch (same source geometry), taps computed in the kernel with ATen's
 * arithmetic (no tables), optional per-clip geometry for the train chain, optional SECOND output = the SlowFast
 * slow pathway (pytorchvideo_trainer/datamodule/transforms.py:99-138 SlowFastPackPathway), uint8 pass-through
 * for pure frame selection (transforms/functional.py:19-41 keeps the dtype; :134-160 _repeated).
 *   src[clip*s_clip + c*sc + idx_t[j]*st + h*sh + w*sw]  ->  dst[clip*d_clip + ((c*n_t + j)*out_h + y)*out_w + x]
 *   slow_pos[j] >= 0: the same pixel is also written to dst_slow[clip*d_slow_clip + ((c*n_slow + slow_pos[j])*out_h + y)*out_w + x]
 *   geom (device, optional): per clip {new_h, new_w, top, left, hflip, first_frame} overriding the descriptor's
 *   values; first_frame is added to every idx_t[j] (temporal views of one video: s_clip = 0)
 * idx_t / slow_pos / geom are DEVICE int32 arrays; src/dst device pointers (src may be pinned host memory
 * mapped into the device address space: decoder frames are read exactly once).                            */
typedef struct pv_clip_batch_desc {
  int C, n_clips, n_t, n_slow;
  int in_h, in_w, new_h, new_w;     /* source frame, resize target (== source for no resize)      */
  int top, left, out_h, out_w;      /* crop window inside the resized frame                       */
  int hflip;                        /* mirror the cropped output along W                          */
  long long sc, st, sh, sw, s_clip; /* source strides in ELEMENTS (channel, frame, row, col, clip) */
  long long d_clip, d_slow_clip;    /* destination strides between clips, in elements             */
  float mean[4], stdv[4];
  int div255, normalize;            /* 1: x/255 first; 1: (x-mean)/std                             */
  int src_dtype, dst_dtype;         /* src: PV_U8|PV_F32|PV_F16, dst: PV_F16|PV_F32|PV_U8 (pass-through) */
} pv_clip_batch_desc;

int pv_clip_transform_batch(const pv_clip_batch_desc* d, const void* src, const int32_t* idx_t,
                            const int32_t* slow_pos, const int32_t* geom, void* dst, void* dst_slow,
                            void* stream);

/* Test-time multi-view ensembling (pytorchvideo_trainer/module/video_classification.py:290-311, docs model_zoo.md:63
 * "3 spatial x 10 temporal views"): out[v][k] = reduce over the n_views consecutive rows of video v;
 * mode 0 = sum, 1 = mean (sum / clip count), 2 = max.  preds: [n_videos * n_views][K] f32.                    */
int pv_view_reduce(const float* preds, float* out, int n_videos, int n_views, int K, int mode, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Layout / dtype conversion at the API boundary.
 * NCDHW (reference model input, models/net.py:41-44) -> NDHWC with C padded to c_pad (zeros).
 * src dtype f32 or f16; dst dtype f16 or f32.
 * ------------------------------------------------------------------------------------------- */
int pv_ncdhw_to_ndhwc(const void* src, int src_dtype, void* dst, int dst_dtype,
                      int N, int C, int T, int H, int W, int c_pad, long long dst_row_stride,
                      void* stream);
/* Same, but every output row is x_w_phys pixels wide with w_pad zero pixels on the left and zeros
 * on the right (dst holds N*T*H*w_phys*c_pad + c_pad*8 elements): the stem layout of window mode. */
int pv_ncdhw_to_ndhwc_padw(const void* src, int src_dtype, void* dst, int dst_dtype,
                           int N, int C, int T, int H, int W, int c_pad, int w_pad, int w_phys,
                           void* stream);
/* Clears n floats (SE accumulators) with a stream-ordered memset. */
int pv_zero_f32(float* dst, long long n, void* stream);
/* NDHWC -> NCDHW f32 (for handing feature maps back to torch when a model has no head). */
int pv_ndhwc_to_ncdhw(const void* src, int src_dtype, long long src_row_stride, float* dst,
                      int N, int C, int T, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------
 * 3-D convolution + folded BatchNorm(eval) scale/bias + optional residual + activation.
 * Replaces nn.Conv3d -> nn.BatchNorm3d(eval) -> activation chains built at
 * models/resnet.py:98-132 (conv_a/b/c), :422-438 (branch1), models/stem.py:80-87,
 * models/slowfast.py:672-679 (FuseFastToSlow), models/x3d.py:66-88,160-217, models/csn.py:169,
 * layers/convolutions.py:191-237 (Conv2plus1d), models/stem.py:330-337 (PatchEmbed) and the
 * residual add + ReLU of models/resnet.py:1179-1189.
 *   y = act( conv(x, w) * scale[co] + bias[co] (+ residual) )
 * groups must be 1 (dense) or == Ci == Co (depthwise).
 *
 * Packed weights (host-side, see pytorchvideo_b200/engine/packing.py):
 *   PV_ALGO_DIRECT, dense : w[tap][ci][co]        (storage dtype, Ci/Co padded)
 *   depthwise (any algo)  : w[tap][c]             (storage dtype)
 *   PV_ALGO_TCGEN05       : w[co][tap][ci_pad64]  (f16, K-major rows of length taps*ci_pad64;
 *                           C_in < 64: ci_pad64 = Ci, row padded to a multiple of 64;
 *                           window mode: w[co][kt*kh][win] with win = 16|32|64 >= kw*Ci)
 * ------------------------------------------------------------------------------------------- */
typedef struct pv_conv3d_desc {
  int dtype;                 /* storage dtype of x, y, residual, w: PV_F16 | PV_F32          */
  int N, Ti, Hi, Wi, Ci;     /* Ci, Co: padded channel counts                               */
  int To, Ho, Wo, Co;
  int kt, kh, kw;
  int st, sh, sw;
  int pt, ph, pw;
  int dt, dh, dw;            /* dilation                                                    */
  int groups;
  int act;                   /* pv_act applied last                                         */
  int has_residual;
  long long x_row_stride, y_row_stride, res_row_stride;
  int ci_pad64;              /* PV_ALGO_TCGEN05 only: per-tap K extent of the packed weights */
  /* "window mode" for 3/4-channel stems on the tensor cores: the input rows physically carry
   * x_w_pad zero pixels on the left (>= pw) and are x_w_phys pixels wide in memory (written that
   * way by pv_ncdhw_to_ndhwc_padw); Wi stays the logical width.  0 = ordinary layout.           */
  int x_w_pad, x_w_phys;
  /* Depthwise path only: element distance between consecutive samples of x / y when a sample is
   * not T*H*W*row_stride apart (MViT token tensors carry a cls row in front of every sample's
   * T*H*W patch tokens).  0 = densely packed.                                                   */
  long long x_batch_stride, y_batch_stride;
} pv_conv3d_desc;

/* Temporal tap reduction used to factor a (kt,kh,kw) stem convolution with few output channels into
 * a (1,kh,kw) convolution producing kt*Co channels (one group of Co per temporal tap, all taps in ONE
 * tensor-core pass over the input) followed by this sum:
 *   y[n][t][p][co] = act(scale[co] * sum_dt Yk[n][t*st + dt*dil - pt][p][dt*Co + co] + bias[co])
 * (frames outside [0,Ti) contribute zero = the temporal zero padding).  Same arithmetic as the
 * direct convolution up to one f16 rounding of the per-tap partial sums.                          */
int pv_temporal_tap_sum(const void* yk, void* y, int dtype, int N, int Ti, int To, long long hw, int Co,
                        int kt, int st, int pt, int dil, const float* scale, const float* bias, int act,
                        long long in_row_stride, long long out_row_stride, void* stream);

int pv_conv3d_fwd(const pv_conv3d_desc* d, int algo, const void* x, const void* w,
                  const float* scale, const float* bias, const void* residual, void* y,
                  void* stream);
/* Depthwise convolution (groups == Ci == Co) + folded BN + activation with optional fused
 * Squeeze-Excitation statistics: when se_sums != NULL, se_sums[n][c] += sum over output positions
 * of the PRE-activation value (caller zeroes it; feeds pv_se_gate).  f16 storage runs as a TMA-fed
 * shared-memory stencil; other cases take the generic CUDA-core stencil (+ pv_channel_sum).
 * Replaces conv_b -> norm_b -> SE-pool of models/x3d.py:180-198, the X3D stem conv_xy
 * (models/x3d.py:74-82), CSN's conv_b (models/csn.py:169) and MViT's pooling convs
 * (layers/attention.py:364-403).  w: [tap][C]. */
int pv_dwconv3d_fwd(const pv_conv3d_desc* d, const void* x, const void* w, const float* scale,
                    const float* bias, void* y, float* se_sums, void* stream);

/* 1 if PV_ALGO_TCGEN05 supports this descriptor (pure host-side check, no GPU needed). */
int pv_conv3d_tcgen05_supported(const pv_conv3d_desc* d);

/* Stem convolutions (ResNetBasicStem.forward models/stem.py:252-260 conv; X3D stem conv_t models/x3d.py:83-88) on the
 * 4-channel, W-padded network input, stride 2 along W: zero-copy im2col - the A operand of a filter row is the RAW
 * input row in shared memory, addressed by a no-swizzle UMMA descriptor with a 16-byte K-chunk stride (csrc/pv_stem.cu).
 * Descriptor: window-mode conventions of pv_conv3d_desc (x_w_pad, x_w_phys, ci_pad64 = window length 16|32|64).
 * w: f16 [K / 8][pad16(Co)][8], K = (dt * kh + dh) * window + (lead + dw) * 4 + c;  zero_row: >= 4 KiB of zeros. */
int pv_conv3d_stem_rows_supported(const pv_conv3d_desc* d);
int pv_conv3d_stem_rows_fwd(const pv_conv3d_desc* d, const void* x, const void* w, const float* scale,
                            const float* bias, const void* zero_row, void* y, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused bottleneck block for narrow pathways (SlowFast Fast pathway), ONE launch:
 *   a = relu(bn_a(conv_a(x)))  (kt,1,1) C_in -> C_mid;   b = relu(bn_b(conv_b(a)))  (1,3,3) stride (1,sb,sb), pad 1
 *   y = act(bn_c(conv_c(b)) + shortcut),  shortcut = x (identity) or bn_1(conv_1(x)), 1x1x1 stride (1,sb,sb)
 * Replaces BottleneckBlock.forward (models/resnet.py:1345-1365) + ResBlock.forward (resnet.py:1179-1189) for
 * C_mid in {8, 16, 32}: a and b stay in shared memory, x is read once, the residual comes from the resident x tile.
 * x, y: NDHWC f16; weights f16 packed [n][k], k = tap * C + ci, K zero-padded to a multiple of 16:
 *   wa [Cmid][pad16(kt*Cin)], wb [Cmid][pad16(9*Cmid)], wc [Cout][pad16(Cmid)], wsc [Cout][pad16(Cin)] (or NULL);
 * folded BatchNorm (scale, bias) fp32 per output channel for each of the four convolutions.
 * ------------------------------------------------------------------------------------------- */
typedef struct pv_bottleneck_desc {
  int N, T, H, W;            /* input extents; output (T, (H-1)/sb+1, (W-1)/sb+1)                 */
  int Cin, Cmid, Cout;       /* padded channel counts (multiples of 8; Cin a power of two)        */
  int kt, sb;                /* temporal taps of conv_a (1|3, padding kt/2); spatial stride of conv_b (1|2) */
  int has_shortcut;          /* 1: projection shortcut conv_1 + bn_1; 0: identity (Cin == Cout, sb == 1)    */
  int act;                   /* PV_ACT_RELU | PV_ACT_NONE after the residual add                  */
  long long x_row_stride, y_row_stride;
} pv_bottleneck_desc;
int pv_bottleneck_fused_supported(const pv_bottleneck_desc* d);
int pv_bottleneck_fused_fwd(const pv_bottleneck_desc* d, const void* x, const void* wa, const void* wb,
                            const void* wc, const void* wsc, const float* sa, const float* ba,
                            const float* sb, const float* bb, const float* sc, const float* bc,
                            const float* ssc, const float* bsc, void* y, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Pooling. nn.MaxPool3d stem pool (models/stem.py:94-100), nn.AvgPool3d head pools
 * (models/head.py:116-118, models/slowfast.py:333-341, models/x3d.py:490), MViT skip-path
 * MaxPool3d (layers/attention.py:677-679).  AvgPool divides by the full kernel volume
 * (count_include_pad=True, the torch default); MaxPool pads with -inf.
 * ------------------------------------------------------------------------------------------- */
typedef struct pv_pool3d_desc {
  int dtype, mode;
  int N, Ti, Hi, Wi, C;
  int To, Ho, Wo;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
  long long x_row_stride, y_row_stride;
  long long x_batch_stride, y_batch_stride;   /* 0 = densely packed (see pv_conv3d_desc) */
} pv_pool3d_desc;
int pv_pool3d_fwd(const pv_pool3d_desc* d, const void* x, void* y, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Squeeze-Excitation (fvcore SqueezeExcitation as used at models/x3d.py:190-198):
 *   pv_channel_sum : sums[n][c] += sum over positions of x        (sums must be zeroed)
 *   pv_se_gate     : gate[n][c] = sigmoid(W2 relu(W1 (sums/npos) + b1) + b2)   (f32 weights)
 *   pv_scale_act   : y = act(x * gate[n][c])   in place allowed
 * ------------------------------------------------------------------------------------------- */
int pv_channel_sum(const void* x, int dtype, long long row_stride, int N, long long npos, int C,
                   float* sums, void* stream);
int pv_se_gate(const float* sums, long long npos, int N, int C, int Cr,
               const float* w1, const float* b1, const float* w2, const float* b2,
               int c_stride_w, float* gate, void* stream);
int pv_scale_act(const void* x, void* y, int dtype, long long x_row_stride,
                 long long y_row_stride, int N, long long npos, int C, const float* gate,
                 int act, void* stream);

/* RoIAlign for the detection heads (pytorchvideo/models/head.py:441-482 ResNetRoIHead.forward: pool -> squeeze T ->
 * roi_layer(x, bboxes) -> pool_spatial ...; roi_layer = torchvision.ops.RoIAlign(output_size, spatial_scale,
 * sampling_ratio), aligned=False, head.py:209-227).  x: NDHWC features with T == 1 ([N][H][W][C], rows of
 * x_row_stride elements, C % 8 == 0); rois: DEVICE fp32 [K][5] = (batch index, x1, y1, x2, y2) in input pixels;
 * y: [K][pooled_h][pooled_w][C].  Sampling grid, bilinear weights and boundary rules are torchvision's.          */
int pv_roi_align_fwd(const void* x, int dtype, long long x_row_stride, int N, int H, int W, int C,
                     const float* rois, int K, int pooled_h, int pooled_w, float spatial_scale,
                     int sampling_ratio, void* y, long long y_row_stride, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Head tail (models/head.py:371-391): optional softmax over channels per position
 * (activation applied BEFORE the global average, head.py:383-390), then mean over positions,
 * un-pad and cast to f32:  out[n][c] = mean_p act(x[n][p][c]),  c < C_valid.
 * ------------------------------------------------------------------------------------------- */
int pv_head_reduce(const void* x, int dtype, long long row_stride, int N, long long npos,
                   int C_valid, int softmax, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MViT pieces (layers/attention.py).
 * pv_layernorm : y = (x-mean)/sqrt(var+eps)*gamma+beta over C channels (f32 statistics); every row
 *                holds `groups` consecutive groups of C channels normalised independently with the
 *                same gamma/beta (groups = heads for the per-head norm_q/k/v, attention.py:200-205)
 *                nn.LayerNorm(eps=1e-6) at attention.py:655,703 / vision_transformers.py:333.
 * nn.Linear    : (attention.py:93-95,315-320,541,716) has NO entry point of its own: y = act(x W^T + b) (+ residual) on a
 *                token tensor [B*N, C] is pv_conv3d_fwd with a 1x1x1 filter (N = B, T = H = 1, W = tokens).
 * pv_attention : o = softmax((q*scale) k^T) v (+ q)    attention.py:531-539, flash-style, the
 *                N_q x N_k matrix is never materialised.  q/k/v/o are [B][N][H][D] with
 *                explicit row strides (elements between consecutive tokens), D = head dim.
 *                f16: tcgen05 kernel (S and O accumulators in TMEM, Q/K/V tiles by TMA, V consumed MN-major as it
 *                lies in memory, two-sweep softmax: csrc/pv_attention_tc.cu), mma.sync kernel for head dims /
 *                strides the TMA path rejects; f32: CUDA-core flash kernel.
 * ------------------------------------------------------------------------------------------- */
int pv_layernorm(const void* x, void* y, int dtype, long long rows, int groups, int C,
                 long long x_row_stride, long long y_row_stride, const float* gamma,
                 const float* beta, float eps, void* stream);
/* pv_layernorm with (a) several (gamma, beta) sets: group g of a row uses set g / groups_per_set (gamma / beta hold
 * groups / groups_per_set sets of C floats) - the pooled K and V of one block, adjacent channel slices of one buffer, are
 * normalised by ONE launch with norm_k | norm_v (attention.py:200-205); (b) an optional second source for every
 * npos-th row (row % npos == 0, read from cls_src + (row / npos) * cls_batch_stride): the cls token by-passes the
 * pooling conv (attention.py:184-186, 196-197) and is normalised with the pooled rows without a copy launch.     */
int pv_layernorm_sets(const void* x, void* y, int dtype, long long rows, int groups, int C,
                      long long x_row_stride, long long y_row_stride, const float* gamma, const float* beta,
                      int groups_per_set, const void* cls_src, long long cls_batch_stride, long long npos,
                      float eps, void* stream);
/* Strided row copy dst[r][0:C] = src[r][0:C] (cls-token rows around the pooling ops). */
int pv_copy_rows(const void* src, void* dst, int dtype, long long rows, int C,
                 long long src_row_stride, long long dst_row_stride, void* stream);
/* MViT cls token + positional encoding (layers/positional_encoding.py:112-136):
 *   y[b][0][:]   = pos[0][:]                      (pos[0] = cls_token + pos_embed_class, f32)
 *   y[b][1+i][:] = x[b][i][:] + pos[1+i][:]       (pos[1+i] = spatial[i %% HW] + temporal[i / HW])
 * x: [B][n_patch][C] (row stride x_row_stride), y: [B][1+n_patch][C] dense.                     */
int pv_add_pos_cls(const void* x, void* y, int dtype, int B, long long n_patch, int C,
                   long long x_row_stride, const float* pos, int has_cls, void* stream);
/* Same with separate dtypes: x f16 -> y f32 starts the fp32 residual trunk of the f16 engine (see pv_add_layernorm). */
int pv_add_pos_cls_to(const void* x, int x_dtype, void* y, int y_dtype, int B, long long n_patch, int C,
                      long long x_row_stride, const float* pos, int has_cls, void* stream);
/* Residual add + LayerNorm of MultiScaleBlock.forward (layers/attention.py:746-757: x = x_res + x_block; norm2(x); ...
 * x = x + x_mlp; the next block's norm1 at :730 / the model's norm_embed, models/vision_transformers.py:177) on an
 * fp32 trunk:  s = a + b in fp32 (a: a_dtype f16|f32, row stride a_row_stride; b: f16 branch output or NULL),
 *   sum[r][:] = s (fp32, optional - the residual stream never takes an f16 rounding),
 *   y[r][:]   = LayerNorm(s) * gamma + beta as f16 (optional - the A operand of the next GEMM), fp32 statistics.
 * C % 8 == 0, C <= 768; gamma / beta fp32, 16-byte aligned.                                                    */
int pv_add_layernorm(const void* a, int a_dtype, long long a_row_stride, const void* b, long long b_row_stride,
                     float* sum, long long sum_row_stride, void* y, long long y_row_stride, long long rows, int C,
                     const float* gamma, const float* beta, float eps, void* stream);
typedef struct pv_attention_desc {
  int dtype;
  int B, H, Nq, Nk, D;
  long long q_row_stride, k_row_stride, v_row_stride, o_row_stride; /* per token           */
  long long q_batch_stride, k_batch_stride, v_batch_stride, o_batch_stride;
  float scale;
  int add_q_residual;   /* residual_pool=True: o += q (cls row included, attention.py:535-536) */
} pv_attention_desc;
int pv_attention_fwd(const pv_attention_desc* d, const void* q, const void* k, const void* v,
                     void* o, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PV_B200_H_ */
