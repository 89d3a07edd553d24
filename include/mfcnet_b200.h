/*
 * mfcnet_b200.h -- C ABI of libmfcnet_b200.so, the B200 (sm_100a) implementation of
 * mfcnet-tracker's multi-frame inference hot path.
 *
 * Boundary contract
 * -----------------
 *  - plain C: raw device pointers, ints and a cudaStream_t passed as void*; no torch types.
 *  - every entry point returns 0 on success or a negative MFC_E* code; the message of the
 *    last error on the calling thread is returned by mfc_last_error().
 *  - no allocation, no synchronisation and no global mutable state inside the hot calls:
 *    the caller owns every buffer (inputs, outputs, workspaces) and the stream.
 *  - there is no CPU fallback: on a device that is not sm_100 every launch returns
 *    MFC_EARCH.
 *
 * The reference is pure Python/PyTorch; its only FFI on this path is the CuPy
 * RawKernel launch in models/unflow_correlation.py:296-329 (raw data_ptr()s + stream).
 * Each entry point below names the reference code it replaces (paths relative to the
 * reference checkout).  The reference-side binding a maintainer would add (ctypes) is
 * shown in INTEGRATION.md.
 *
 * Activation layout ("C8"): [B][ceil(C/8)][H][W][8] of fp16 (default) or bf16 -- channel
 * chunks of 8 are planes, so a pixel's 8 channels are one 16-byte vector, channel concat
 * is a list of plane pointers (zero-copy) and a plane is directly the K-major, no-swizzle
 * shared-memory image a tcgen05.mma A-operand descriptor walks.
 */
#ifndef MFCNET_B200_H_
#define MFCNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFC_ABI_VERSION 5

/* error codes */
#define MFC_OK 0
#define MFC_EINVAL (-1)   /* bad argument / unsupported shape            */
#define MFC_EARCH (-2)    /* device is not sm_100 (B200)                 */
#define MFC_ECUDA (-3)    /* CUDA runtime error (see mfc_last_error())   */
#define MFC_ENOMEM (-4)   /* caller-provided workspace too small         */

/* element types of C8 activations / packed weights */
#define MFC_F16 0
#define MFC_BF16 1

#define MFC_MAX_SRC 8

int mfc_abi_version(void);
const char* mfc_last_error(void);
/* 0 if `device` can run the kernels (compute capability 10.x), else MFC_EARCH. */
int mfc_device_check(int device);

/* ------------------------------------------------------------------------------------
 * Layout conversion.  Replaces the implicit NCHW tensors + torch.cat of
 * models/multiframe_model.py:424-436 (frames, flows and depths enter as separate NCHW
 * fp32 tensors and are concatenated on channels).
 * ---------------------------------------------------------------------------------- */

/* Gather up to 8 fp32 NCHW channel planes into ONE C8 plane (chunk) of `dst`.
 * plane[j] points at channel j's H*W plane of sample 0 (NULL -> zeros);
 * plane_bstride[j] = elements between consecutive samples of that source. */
typedef struct MfcGather {
  const float* plane[8];
  long long plane_bstride[8];
} MfcGather;
int mfc_gather_nchw_to_c8(const MfcGather* g, void* dst_chunk, long long dst_bstride_bytes,
                          int B, int H, int W, int dtype, void* stream);

/* C8 (first `C` channels) -> fp32 NCHW [B][C][H][W]. */
int mfc_c8_to_nchw(const void* src, long long src_bstride_bytes, float* dst, int B, int C, int H, int W,
                   int dtype, void* stream);

/* ------------------------------------------------------------------------------------
 * Weight preparation (run once per checkpoint load, results cached by the caller).
 * ---------------------------------------------------------------------------------- */

/* Weight standardisation of models/resunet.py:56-62: per output channel
 * (w - mean) * rsqrt(biased_var + eps) over (Cin,kh,kw).  w, out: fp32 [Cout][fan_in]. */
int mfc_weight_standardize(const float* w, float* out, int Cout, int fan_in, float eps, void* stream);

/* Eval-mode BatchNorm as a per-channel affine (the BN after every conv in
 * models/multiframe_model.py:62-73 and models/hrnet.py): scale = g*rsqrt(var+eps),
 * shift = b - mean*scale, optionally composed with a conv bias (shift += bias*scale). */
int mfc_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var,
                const float* conv_bias /*or NULL*/, float eps, float* scale, float* shift, int C, void* stream);

/* Geometry the conv kernel derives from a descriptor (needed to size buffers). */
typedef struct MfcConvInfo {
  int nb;              /* output channels per N-block (multiple of 16, <= 256)        */
  int nblk;            /* number of N-blocks; padded Cout = nb*nblk                   */
  int cin_chunks;      /* total 8-channel input planes (sum of src[].nchunks)         */
  int ksteps;          /* ceil(cin_chunks/2): 16-channel MMA K-steps per tap          */
  int tile_h, tile_w;  /* output tile                                                 */
  int tiles_per_image; /* output tiles in one sample                                  */
  int stats_per_image; /* GroupNorm partial records per sample (one per persistent CTA):
                          the stats buffer is [B][stats_per_image][nb*nblk][2]         */
  int runs;            /* 128-row MMA runs per tile                                   */
  int kstages;         /* channel stages of the K loop                                */
  int nstages;         /* depth of the shared-memory stage ring                       */
  int grid;            /* persistent CTAs launched                                    */
  int smem_bytes;      /* dynamic shared memory per CTA                               */
  int tmem_cols;       /* TMEM columns allocated per CTA                              */
  long long packed_weight_bytes;
  int weight_layout;   /* 0 = one B block per filter tap, 1 = sliding-accumulate layout (vertical taps stacked along N);
                          packed weights are only valid for descriptors that report the same layout and nb/nblk/ksteps */
  int flags;           /* bit 0: the plan runs the affine-addressed fast epilogue (required by io->head_w) */
} MfcConvInfo;

typedef struct MfcSrc {
  const void* ptr;         /* C8 planes [B][nchunks][Hin][Win][8] (device)                      */
  const float* affine;     /* NULL, or [B][nchunks*8][2] (scale,shift): x -> silu(x*scale+shift) */
  long long batch_stride;  /* bytes between samples                                             */
  int nchunks;             /* 8-channel planes taken from this source                           */
  int reserved;
} MfcSrc;

typedef struct MfcConvDesc {
  int B, Hin, Win;         /* source size (before the optional nearest x2)        */
  int Hout, Wout;
  int Cout;                /* real output channels                                */
  int kh, kw, stride, pad; /* stride 1 or 2                                       */
  int upsample;            /* 1, or 2 = nearest-neighbour x2 fused into the loader */
  int act;                 /* 0 none, 1 ReLU, 2 LeakyReLU(0.1) (after scale/shift and residual) */
  int dtype;               /* MFC_F16 / MFC_BF16 (inputs, weights, C8 outputs)    */
  int nsrc;                /* channel-concat sources, in order                    */
  /* Output-parity mode for ConvTranspose2d(k=4, s=2, p=1) (models/ternausnet.py:35): the transposed conv is
   * four 2x2 convs, one per output parity (py,px).  in_off_{y,x} (0 or 1) shift the input window
   * (input row = oy*stride - pad + in_off_y + ky); out_stride 2 makes the epilogue store pixel (oy,ox) at
   * (oy*2 + out_off_y, ox*2 + out_off_x) of a [2*Hout][2*Wout] output tensor.  out_stride 1 (default, 0
   * is read as 1) and zero offsets = an ordinary conv. */
  int in_off_y, in_off_x;  /* ordinary convs (out_stride 0 / 1): REDUCE the padding of that axis, 0 <= in_off <= pad -- rectangular
                              kernels such as RAFT's ConvGRU (1x5 with padding (0, 2): kh 1, kw 5, pad 2, in_off_y 2);
                              Hout = (Hup + 2 (pad - in_off_y) + pad_br - kh)/stride + 1                                 */
  int out_stride, out_off_y, out_off_x;
  int pad_br;              /* extra zero padding on the bottom and right edges only (nn.ZeroPad2d([l, l+e, t, t+e]) in front of
                              a pad-0 conv, models/unflow_model.py:85-130: pad = l, pad_br = e); Hout = (Hup + 2 pad + pad_br - kh)/stride + 1 */
  int reserved;            /* flags: MFC_CONV_HAS_RESIDUAL when mfc_conv2d_fwd will be given io->residual (the plan then
                              reserves the shared-memory prefetch rings for it)                                        */
  MfcSrc src[MFC_MAX_SRC];
} MfcConvDesc;
#define MFC_CONV_HAS_RESIDUAL 1
/* Walk the work items from the last sample to the first.  Activations of a batch are larger than the 126 MB L2: a conv that
 * runs in the opposite direction of its producer starts with the part of its input that is still cache-resident ("snake"
 * order through the layers).  Results are identical up to the grouping of the GroupNorm partial sums. */
#define MFC_CONV_REVERSE_ORDER 2
/* mfc_conv2d_fwd will be given io->stats: the plan must keep the padded Cout within the 256 channels the epilogue's
 * shared-memory scratch holds (part of the plan key, like MFC_CONV_HAS_RESIDUAL). */
#define MFC_CONV_WANT_STATS 4
/* mfc_conv2d_fwd will be given io->head_w: the plan is restricted to tilings whose epilogue holds all output channels of a
 * pixel in one thread with affine pixel addressing (Cout <= 16; part of the plan key). */
#define MFC_CONV_WANT_HEAD 8

int mfc_conv2d_query(const MfcConvDesc* d, MfcConvInfo* info);

/* OIHW fp32 (device) -> the packed tcgen05 B-operand image.
 * chan_map[k] (device int32, length cin_chunks*8, or NULL = identity) gives, for padded
 * input channel k of the concat (8 per plane), the index into the weight's Cin axis or -1.
 * scale[Cout] (device fp32 or NULL): per-output-channel factor multiplied into the weights before they are rounded
 * (a folded BatchNorm scale; mfc_conv2d_fwd is then called with io->scale = NULL). */
int mfc_conv2d_pack_weights(const MfcConvDesc* d, const float* w_oihw, int Cin_w, const int* chan_map,
                            const float* scale, void* packed, void* stream);

/* Fused convolution.  Replaces every nn.Conv2d / WeightStandardizedConv2d +
 * following BatchNorm/ReLU, the torch.cat skip-concats (models/resunet.py:168,174), the
 * nearest Upsample (:39-43) and the pixel-unshuffle Downsample (:45-49, expressed as a
 * k2 s2 conv) on the path:
 *   acc  = sum_taps  W * T(src)                 T = optional silu(GroupNorm-affine) per source
 *   v    = acc*scale[co] + shift[co]            (bias, folded BN)
 *   v   += R(residual)                          R = optional silu(affine) of a C8 tensor
 *   v    = act(v)
 *   stats[b][tile][co] += (v, v*v)              optional per-channel partial sums for GroupNorm
 *   y_c8 / y_nchw = v                           either or both
 */
typedef struct MfcConvIO {
  const void* w_packed;
  const float* scale;        /* [nb*nblk] or NULL (=1) */
  const float* shift;        /* [nb*nblk] or NULL (=0) */
  const void* residual;      /* C8 [B][ceil(Cout/8)][Hout][Wout][8] or NULL */
  const float* res_affine;   /* NULL or [B][ceil(Cout/8)*8][2]               */
  long long res_batch_stride;
  void* y_c8;                /* C8 output or NULL                            */
  long long y_batch_stride;  /* bytes                                        */
  float* y_nchw;             /* fp32 [B][Cout][Hout][Wout] or NULL           */
  float* stats;              /* [B][stats_per_image][nb*nblk][2] or NULL     */
  /* Optional fused 1x1 "head": a second linear layer applied to the epilogue values v (after shift / residual / act),
   *   y_nchw[b][n][pixel] = head_b[n] + sum_co head_w[n*16 + co] * v[co],   n < head_n <= 8,
   * evaluated in fp32 on the values still in registers: the conv's own output need not be stored (y_c8 may be NULL) and the
   * last activation / weights of a network are never rounded to fp16 (models/multiframe_model.py:72-73: the fusion head's
   * final 1x1 conv).  Needs the descriptor flag MFC_CONV_WANT_HEAD, Cout <= 16, y_nchw, no stats / residual. */
  void* y_lo;                /* NULL, or a second C8 output (geometry / stride of y_c8) receiving the rounding residue
                                v - fp(v) of every stored value: [y_c8, y_lo] read as two concat sources with the same weights
                                carry the activation to ~22 bits (a network's last hidden activation)                    */
  const float* head_w;       /* [head_n][16] fp32 (device; columns >= Cout zero) or NULL */
  const float* head_b;       /* [head_n] or NULL                                         */
  int head_n;
  int reserved;
  int* overflow;             /* NULL, or a device counter: fp16 range guard.  Every epilogue warp that stored a value
                                beyond +-65504 into y_c8 (it became +-inf) adds 1.  The caller zeroes and reads it
                                (mfc_run_list users: one counter for the whole program).                              */
} MfcConvIO;
int mfc_conv2d_fwd(const MfcConvDesc* d, const MfcConvIO* io, void* stream);

/* Plans.  The tiling of a geometry (B, sizes, channels, kernel, stride, pad, upsample, chunk count, affine-on-load,
 * flags, dtype, out_stride) is fixed the first time mfc_conv2d_query / _pack_weights / _fwd sees it and never changes
 * afterwards in that process: packed weights, the statistics buffer and mfc_gn_finalize's record count depend on it.
 * It is taken from (1) mfc_conv2d_autotune, if that ran for the geometry first, (2) the imported tuning table,
 * (3) the cost model.  (2) and (3) are deterministic: processes that do not tune live produce identical bits.
 *
 * mfc_conv2d_autotune measures the planner's shortlisted tilings (tile shape, K staging, weight layout) of `d` on the
 * device with the caller's real buffers and keeps the fastest.  It is a no-op (returns MFC_OK) for a geometry that already
 * has a plan or a table entry.  io->w_packed is ignored; the raw OIHW weights are packed into `scratch_packed`
 * (`scratch_bytes`: twice the packed_weight_bytes of mfc_conv2d_query covers every candidate; candidates that do not fit are
 * skipped) once per weight image.  io->stats, when given, must hold [B][148][nb*nblk][2] floats.  Synchronises `stream`.
 *
 * Weights, scale and shift are read by the conv kernel BEFORE it waits for its predecessor on the stream (programmatic
 * dependent launch): they must be complete -- the producing work synchronised -- before a conv that uses them is launched. */
int mfc_conv2d_autotune(const MfcConvDesc* d, const MfcConvIO* io, const float* w_oihw, int Cin_w, const int* chan_map,
                        void* scratch_packed, long long scratch_bytes, int reps, void* stream);
/* Packed-weight scratch bytes that cover every candidate mfc_conv2d_autotune measures for `d` (does not plan `d`). */
long long mfc_conv2d_autotune_scratch_bytes(const MfcConvDesc* d);
/* Diagnostic: the candidates mfc_conv2d_autotune would measure, one per line
 * "TH TW slide CBc NB nstages kstages R nacc tiles_x grid", in the cost model's order.  Returns the bytes needed. */
long long mfc_conv2d_shortlist(const MfcConvDesc* d, int per_bucket, char* buf, long long cap);
/* Tuning table as text, one geometry per line ("<16 key ints> : TH TW slide CBc NB nstages", '#' = comment).
 * export returns the bytes needed including the terminating 0 and writes at most `cap`; import returns the number of
 * entries taken (geometries that already have a plan keep it) or a negative error code. */
long long mfc_conv2d_plan_export(char* buf, long long cap);
int mfc_conv2d_plan_import(const char* text);

/* GroupNorm statistics -> per-(sample,channel) affine.  Replaces nn.GroupNorm
 * (models/resunet.py:72,77) split in two: partial sums come from the producing conv's
 * epilogue, this finalises them (fp64) into scale = g*rstd, shift = b - mean*g*rstd. */
int mfc_gn_finalize(const float* stats, int B, int stats_per_image, int cpad, int C, int groups,
                    long long pixels, const float* gamma, const float* beta, float eps,
                    float* affine /*[B][ceil(C/8)*8][2]*/, void* stream);

/* out = silu(a*scale+shift) + r   (ResnetBlock tail with identity res_conv,
 * models/resunet.py:90-95).  All C8 [B][chunks][H][W][8]. */
int mfc_affine_silu_add(const void* a, const float* affine, const void* r, void* out,
                        int B, int chunks, long long pixels, int dtype, int* overflow /*or NULL*/, void* stream);

/* nn.MaxPool2d(2, 2) (models/ternausnet.py:56,107) on a C8 tensor [B][chunks][H][W][8] -> [H/2][W/2]. */
int mfc_maxpool2(const void* src, long long src_bstride_bytes, void* dst, long long dst_bstride_bytes,
                 int B, int chunks, int H, int W, int dtype, void* stream);

/* ------------------------------------------------------------------------------------
 * HRNet resampling (models/hrnet.py).
 * ---------------------------------------------------------------------------------- */

/* Fuse step of HighResolutionModule.forward (models/hrnet.py:237-260) and the head's
 * upsample+concat (:464-471): out = act(scale * sum_j T_j + shift), where term j is a C8
 * tensor either at the output size or at a lower resolution that is bilinearly upsampled
 * (F.interpolate, mode='bilinear', align_corners=False) while being read.  Terms are added
 * in order, in fp32.  scale/shift: per output channel [chunks*8] or NULL. */
typedef struct MfcFuseTerm {
  const void* ptr;          /* C8 [B][chunks][H][W][8] */
  long long batch_stride;   /* bytes */
  int H, W;
} MfcFuseTerm;
typedef struct MfcFuseArgs {
  int B, chunks, H, W;      /* output geometry */
  int nterms;               /* 1..MFC_MAX_SRC */
  int act;                  /* 0 none, 1 ReLU */
  int dtype;
  int reserved;
  MfcFuseTerm term[MFC_MAX_SRC];
  const float* scale;
  const float* shift;
  void* out;
  long long out_batch_stride;
  int* overflow;            /* fp16 range guard counter or NULL (see MfcConvIO.overflow) */
  void* out_lo;             /* NULL, or a second C8 tensor (same geometry / stride as `out`) that receives the rounding
                               residue v - fp(v): a consumer that reads [out, out_lo] as two concat sources with the same
                               weights sees the value to ~22 bits (used for a network's last activation) */
} MfcFuseArgs;
int mfc_fuse_sum(const MfcFuseArgs* a, void* stream);

/* F.interpolate(x, size=(Hout,Wout), mode='bilinear', align_corners=False) of fp32 NCHW maps
 * (models/hrnet.py:473-474).  Writes fp32 NCHW and / or the first ceil(C/8) C8 planes. */
typedef struct MfcResizeArgs {
  const float* src;         /* [B][C][Hin][Win] */
  float* dst_nchw;          /* [B][C][Hout][Wout] or NULL */
  void* dst_c8;             /* C8 [B][ceil(C/8)][Hout][Wout][8] or NULL */
  long long c8_batch_stride;
  int B, C, Hin, Win, Hout, Wout, dtype, reserved;
} MfcResizeArgs;
int mfc_bilinear_resize(const MfcResizeArgs* a, void* stream);

/* ------------------------------------------------------------------------------------
 * Temporal fusion pieces.
 * ---------------------------------------------------------------------------------- */

/* Flow warp of MultiFrameNetBasic (models/multiframe_model.py:89-170): bilinear
 * grid_sample(align_corners=True, zeros) of frame i>=1 class maps and depth at
 *   grid[:,:,:H,:W] + flow/((W-1)/2,(H-1)/2)   (the stored 576x720 grid, cropped).
 * seg: C8 plane per frame (nchunks_seg planes each); flow: fp32 NCHW (B,2,H,W) per frame;
 * depth: fp32 NCHW (B,1,H,W) per frame or NULL.  Outputs: warped seg planes (frame 0 is
 * passed through by the caller, zero-copy) and one C8 depth plane (d0, warped d1..dK-1). */
typedef struct MfcWarpArgs {
  int B, H, W, K;
  int seg_chunks;                   /* planes per frame (ceil(N/8))              */
  int grid_h, grid_w;               /* 576, 720                                  */
  const float* grid;                /* (1,2,grid_h,grid_w) fp32                  */
  const void* seg[MFC_MAX_SRC];     /* per frame i (i=0 unused), C8             */
  long long seg_bstride[MFC_MAX_SRC];
  const float* flow[MFC_MAX_SRC];   /* flow[i-1] for frame i: (B,2,H,W) fp32    */
  long long flow_bstride[MFC_MAX_SRC];   /* elements between samples of flow[i]   */
  const float* depth[MFC_MAX_SRC];  /* per frame (B,1,H,W) fp32 or NULL          */
  long long depth_bstride[MFC_MAX_SRC];  /* elements between samples of depth[i]  */
  void* seg_out[MFC_MAX_SRC];       /* per frame i>=1, C8                        */
  long long seg_out_bstride[MFC_MAX_SRC];
  void* depth_out;                  /* one C8 plane or NULL                      */
  long long depth_out_bstride;
  int dtype;
} MfcWarpArgs;
int mfc_flow_warp(const MfcWarpArgs* a, void* stream);

/* Heat-map head (F.log_softmax call sites src/engine.py:65,141;
 * scripts/test_multiframe_segmentation_on_videos_v3.py:281,289): one pass over fp32 NCHW
 * logits -> log-probs and/or probs (exp of log-probs) and the first-max argmax (uint8). */
int mfc_heatmap_head(const float* logits, int B, int N, long long pixels,
                     float* logp /*or NULL*/, float* prob /*or NULL*/, uint8_t* argmax /*or NULL*/,
                     void* stream);

/* numpy.argmax(axis=1) of a (B,N,pixels) fp32 map as uint8: the first maximum wins
 * (scripts/test_multiframe_segmentation_on_videos_v3.py:289, utils/localization_utils_v2.py:201). */
int mfc_argmax_u8(const float* x, int B, int N, long long pixels, uint8_t* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Training loss, forward (src/engine.py:65-66, src/loss.py:6-63): log_softmax of the model output,
 * weighted NLL (nn.NLLLoss(weight=class_weights), mean reduction) and the soft-Jaccard term over the
 * foreground classes; total = w_nll*nll + w_jaccard*jaccard.  logits: fp32 [B][N][pixels], N <= 16;
 * target: int64 [B][pixels] in [0,N); class_weights: [N] or NULL.  workspace: fp64
 * [mfc_segmentation_loss_workspace(B,pixels)] (bytes).  out: float[3] = {total, nll, jaccard} (device).
 * Deterministic (fixed-order two-level reduction).
 * ---------------------------------------------------------------------------------- */
long long mfc_segmentation_loss_workspace(int B, int N, long long pixels);
int mfc_segmentation_loss(const float* logits, const long long* target, const float* class_weights,
                          int B, int N, long long pixels, float w_nll, float w_jaccard,
                          void* workspace, float* out, void* stream);

/* The additive statistics of the loss over one rank's shard: sums[2 + 3(N-1)] doubles = (sum w[t](-logp[t]), sum w[t], then
 * per class c >= 1: I_c, S_c, T_c of src/loss.py:45-63).  Data-parallel ranks add their records (one 14-double all-reduce)
 * and every rank evaluates the loss of the GLOBAL batch -- what nn.DataParallel's gather-to-GPU0 computes
 * (src/engine.py:56-66).  `workspace`: mfc_segmentation_loss_workspace bytes. */
int mfc_segmentation_loss_sums(const float* logits, const long long* target, const float* class_weights, int B, int N, long long pixels,
                               void* workspace, double* sums, void* stream);
int mfc_segmentation_loss_from_sums(const double* sums, int N, float w_nll, float w_jaccard, float* out /*[3]*/, void* stream);

/* Backward of the loss (the `loss.backward()` of src/engine.py:70 down to the model output):
 * dlogits[B][N][pixels] = scale * d total / d logits of this rank's shard.  `sums` holds n_records records of loss
 * statistics that are added first (1 = the record of mfc_segmentation_loss_sums, possibly all-reduced over ranks; 0 = `sums`
 * is the workspace mfc_segmentation_loss just filled for the same logits).  `coef`: 64 floats of device scratch. */
int mfc_segmentation_loss_bwd(const float* logits, const long long* target, const float* class_weights, int B, int N, long long pixels,
                              float w_nll, float w_jaccard, float scale, const void* sums, int n_records, float* coef, float* dlogits,
                              void* stream);

/* One torch.optim.Adam step (amsgrad off; scripts/train_multiframe_detection.py:128-151 builds two parameter groups =
 * two calls with different lr) on a flat fp32 bucket.  `step` is the 1-based step count; grad_scale multiplies the
 * gradient first (the 1/world_size of the data-parallel average after the NCCL sum). */
int mfc_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, float grad_scale, void* stream);


/* Frame ingest of the video loop (scripts/test_multiframe_segmentation_on_videos_v3.py:234-258), bit-exact with its numpy /
 * torchvision arithmetic.  `bgr`: B uint8 frames [H][W][3] as cv2 delivers them (device memory, frame_stride_bytes apart),
 * already at the network's input size.
 *   mfc_ingest_rgb  : BGR2RGB -> /255 -> (t - mean[c]) / std[c]  -> fp32 [B][3][H][W]   (mean / std: 3 HOST floats each)
 *   mfc_ingest_depth: OpenCV's fixed-point BGR2GRAY -> /255       -> fp32 [B][1][H][W] */
int mfc_ingest_rgb(const uint8_t* bgr, long long frame_stride_bytes, float* out, int B, int H, int W, const float* mean3_host,
                   const float* std3_host, void* stream);
int mfc_ingest_depth(const uint8_t* bgr, long long frame_stride_bytes, float* out, int B, int H, int W, void* stream);
/* cv2.resize(frame, (W, H)) (default INTER_LINEAR, 8-bit; scripts/test_multiframe_segmentation_on_videos_v3.py:253,257) on the
 * device, bit-exact with OpenCV's fixed-point scheme: src uint8 [B][h][w][C] (C = 1 or 3, `frame_stride_bytes` between
 * frames) -> dst uint8 [B][H][W][C], dense.  mfc_bgr2gray_u8 = cvtColor(BGR2GRAY) (:244, at the source size, before the
 * resize); mfc_ingest_gray = astype(float32)/255 of a gray frame (:258): uint8 [n] -> fp32 [n]. */
int mfc_resize_u8(const uint8_t* src, long long frame_stride_bytes, int h, int w, int C, uint8_t* dst, int B, int H, int W, void* stream);
int mfc_bgr2gray_u8(const uint8_t* bgr, long long frame_stride_bytes, uint8_t* gray, int B, int H, int W, void* stream);
int mfc_ingest_gray(const uint8_t* gray, float* out, long long n, void* stream);

/* ------------------------------------------------------------------------------------
 * UnFlow network around the correlation (models/unflow_model.py); the convolutions / transposed convolutions /
 * LeakyReLUs of FlowNetC + 2 x FlowNetS run through mfc_conv2d_fwd (act 2, pad_br, parity mode).
 *   mfc_unflow_preprocess : UnFlow.forward :253-262 -- RGB -> BGR, minus the per-channel means; fp32 [B][3][H][W] in / out
 *   mfc_nchw_to_c8        : fp32 NCHW -> C8 planes (the 441-channel cost volume entering moduleCombined, :120-123,165)
 *   mfc_unflow_warp       : backward(second, flow) (:6-17): grid_sample(bilinear, border, align_corners=False) at
 *                           linspace(-1,1) + flow/((size-1)/2); optionally also |first - warped| (Simple.forward :224-226)
 *   mfc_unflow_upscale    : moduleUpscale (:58-61,77): ConvTranspose2d(2,2,k3,s2,p1,bias=False) + ReplicationPad2d([0,1,0,1]),
 *                           times `scale`; x [B][2][h][w] -> out [B][2][2h][2w], w = the transposed conv's [2][2][3][3] weight
 * ---------------------------------------------------------------------------------- */
int mfc_unflow_preprocess(const float* rgb, float* out, int B, int H, int W, void* stream);
int mfc_nchw_to_c8(const float* src, void* dst, long long dst_bstride_bytes, int B, int C, int H, int W, int dtype, void* stream);
int mfc_unflow_warp(const float* second, const float* flow, const float* first /*or NULL*/, float* warped, float* absdiff /*or NULL*/,
                    int B, int C, int H, int W, void* stream);
int mfc_unflow_upscale(const float* x, const float* w, float* out, int B, int h, int w_in, float scale, void* stream);

/* ------------------------------------------------------------------------------------
 * RAFT-large, the online optical-flow provider of the video loop (scripts/test_multiframe_segmentation_on_videos_v3.py:264-271,
 * 342-350 and src/engine.py:39-53 call torchvision.models.optical_flow.raft_large -- a third-party dependency of the reference,
 * torchvision 0.26 `models/optical_flow/raft.py`).  Its convolutions run through mfc_conv2d_fwd; these are the other pieces.
 *   mfc_pointwise : element-wise glue on dense C8 tensors [B][chunks][pixels][8] (kind below)
 *   mfc_raft_op   : correlation volume / pyramid pooling / pyramid lookup / flow update / convex upsampling (kind below)
 * ---------------------------------------------------------------------------------- */
#define MFC_PW_AFFINE_ADD 0  /* out = [relu_out]( [relu_a](a*s_a + t_a) + (r*s_r + t_r) ); a_aff / r / r_aff may be NULL
                                (ResidualBlock.forward with InstanceNorm: relu(x + relu(norm(conv(.))))); affines [B][chunks*8][2] */
#define MFC_PW_CTX_SPLIT 1   /* a has 2*chunks planes: out = tanh(a[:, :chunks]), out2 = relu(a[:, chunks:]) (RAFT.forward)  */
#define MFC_PW_GRU_RH 2      /* a = fused z|r pre-activations (2*chunks planes), r = h: out = sigmoid(a[:, chunks:]) * h      */
#define MFC_PW_GRU_UPDATE 3  /* a = z|r, r = q pre-activation, out = h IN PLACE: h = (1-sigmoid(z))*h + sigmoid(z)*tanh(q)   */
typedef struct MfcPointwiseArgs {
  const void* a;
  const float* a_aff;
  const void* r;
  const float* r_aff;
  void* out;
  void* out2;
  long long pixels;
  int kind, B, chunks, dtype, relu_a, relu_out;
} MfcPointwiseArgs;
int mfc_pointwise(const MfcPointwiseArgs* a, void* stream);

#define MFC_RAFT_CORR_VOLUME 0 /* p0 = fmap1, p1 = fmap2 (fp32 [B][C][h*w]), p2 = out fp32 [B][h*w][h*w] = <f1_i, f2_j> * scale
                                  (CorrBlock._compute_corr_volume: scale = 1/sqrt(C))                                           */
#define MFC_RAFT_POOL 1        /* p0 = in fp32 [B][h][w] (B = all leading dimensions), p1 = out [B][h/2][w/2]: avg_pool2d(2, 2)  */
#define MFC_RAFT_LOOKUP 2      /* p0..p3 = pyramid levels ([B*h*w][h>>l][w>>l]), p4 = flow fp32 [B][2][h][w] (coords1 - coords0),
                                  p5 = out C8 [B][ceil(levels*(2r+1)^2 / 8)][h][w][8]  (CorrBlock.index_pyramid)                 */
#define MFC_RAFT_FLOW_ADD 3    /* p0 = flow fp32 [B][2][h][w] += p1 = delta_flow                                                */
#define MFC_RAFT_UPSAMPLE 4    /* p0 = flow fp32 [B][2][h][w], p1 = mask fp32 [B][576][h][w] (before the multiplier `scale`),
                                  p2 = out fp32 [B][2][8h][8w]  (upsample_flow with up_mask)                                     */
#define MFC_RAFT_RESIZE_AC 5   /* p0 = in fp32 [B][C][h][w], p2 = out fp32 [B][C][levels][radius] (levels = Hout, radius = Wout):
                                  F.interpolate(in * scale, size, mode='bilinear', align_corners=True) -- the video script's resize
                                  of flow / 0.5 to the frame size (scripts/test_multiframe_segmentation_on_videos_v3.py:269)      */
typedef struct MfcRaftArgs {
  const void* p0;
  const void* p1;
  const void* p2;
  const void* p3;
  const void* p4;
  void* p5;
  int kind, B, C, h, w, levels, radius, dtype;
  float scale;
  int reserved;
} MfcRaftArgs;
int mfc_raft_op(const MfcRaftArgs* a, void* stream);

/* ------------------------------------------------------------------------------------
 * UnFlow correlation cost volume (models/unflow_correlation.py:10-105,282-337).
 * first/second: fp32 NCHW contiguous; out: fp32 [B][D*D][H][W], D = 2*(max_disp/stride2)+1,
 * out[b,(iy*D+ix),y,x] = mean_c first[b,c,y,x]*second[b,c,y+dy,x+dx], zero outside.
 * exact_order != 0 reproduces the reference's summation order bit-for-bit
 * (32 strided partial FMA chains, serial lane add, one divide).
 * ---------------------------------------------------------------------------------- */
int mfc_correlation_fwd(const float* first, const float* second, float* out,
                        int B, int C, int H, int W, int max_disp, int stride2, int exact_order,
                        void* stream);
/* Backward of the cost volume (`_FunctionCorrelation.backward`, models/unflow_correlation.py:339-391; kernels :107-235):
 *   grad_first [b,c,y,x] = (1/C) sum_d grad_out[b,d,y,x]           * second[b,c,y+dy,x+dx]
 *   grad_second[b,c,y,x] = (1/C) sum_d grad_out[b,d,y-dy,x-dx]     * first [b,c,y-dy,x-dx]
 * in the reference's summation order (bit-identical results).  Either gradient pointer may be NULL. */
int mfc_correlation_bwd(const float* first, const float* second, const float* grad_out, float* grad_first /*or NULL*/,
                        float* grad_second /*or NULL*/, int B, int C, int H, int W, int max_disp, int stride2, void* stream);

/* ------------------------------------------------------------------------------------
 * Key-point extraction (utils/localization_utils_v2.py:5-40).
 * ---------------------------------------------------------------------------------- */

/* scipy.ndimage.gaussian_filter(heat, sigma) semantics: radius int(4*sigma+0.5), reflect
 * boundary, axis 0 then axis 1, fp64 accumulation, fp32 store after each axis.
 * weights: the 2*radius+1 fp64 taps (host computes them exactly as scipy does).
 * tmp: fp32 scratch of the same size. */
int mfc_gaussian_blur(const float* heat, float* tmp, float* out, int B, int H, int W,
                      const double* weights_dev, int radius, void* stream);

/* localmax = (maximum_filter(sm, footprint) == sm) & blob, as 0/255 uint8.
 * footprint: fh x fw uint8 (device), scipy origin convention (centre = size//2), reflect. */
int mfc_localmax_mask(const float* sm, const uint8_t* cls, int cls_id, const uint8_t* footprint,
                      int fh, int fw, uint8_t* mask, int B, int H, int W, void* stream);

/* (argmax == cls_id) as 0/255 uint8. */
int mfc_class_mask(const uint8_t* cls, int cls_id, uint8_t* mask, long long n, void* stream);

/* cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) + contourArea + moments on ONE 0/255
 * mask (H x W), restricted to what calc_centroids (utils/localization_utils_v2.py:15-33) needs.
 * For every external contour one record of 6 doubles is appended (in no particular order):
 *   {a00, a10, a01, first_x, first_y, 0}
 * a00/a10/a01 are the exact integer Green's-theorem sums of the closed border polygon (positive
 * orientation), from which the caller derives, with OpenCV's own formulas,
 *   contourArea = |a00| * 0.5,  m00 = a00 * 0.5,  m10 = a10 * (1/6),  m01 = a01 * (1/6);
 * (first_x, first_y) is the contour's first point = the component's raster-first
 * pixel, so OpenCV's output order is "descending first_y*W+first_x".  *n_out (device int)
 * receives the TOTAL number of external contours; records beyond max_contours are dropped.
 * labels: int32 scratch [8*H*W], 16-byte aligned (labels, per-pixel contour codes, 3 int64 sums per
 * pixel); mfc_refine_tip_mask reads it afterwards. */
int mfc_trace_contours(const uint8_t* mask, int H, int W, int* labels, double* out, int max_contours,
                       int* n_out, void* stream);

/* The two largest contours of a mfc_trace_contours record list, in the order
 * `sorted(contours, key=cv2.contourArea, reverse=True)[:2]` gives them (area descending, ties in
 * findContours order) -- all calc_centroids / calc_base_centroid ever use
 * (utils/localization_utils_v2.py:17, ...videos_v3.py:47).  top: 2 records of 6 doubles,
 * {a00, a10, a01, first_x, first_y, present (1 / 0)}. */
int mfc_top_contours(const double* rec, const int* n_contours, int max_contours, int W, double* top,
                     void* stream);

/* The class map of the video script when --score_detection_threshold > 0
 * (scripts/test_multiframe_segmentation_on_videos_v3.py:282-287): 0, then classes 1..N-1 painted in
 * ascending order where prob > thr (float32 comparison, as numpy does against a Python float). */
int mfc_threshold_classes(const float* prob, int B, int N, long long pixels, float thr, uint8_t* out,
                          void* stream);

/* out = (cls == cls_id) ? heat : 0   (`left_tip_heatmap[left_tip==0] = 0`, ...videos_v3.py:88). */
int mfc_mask_heat(const float* heat, const uint8_t* cls, int cls_id, float* out, long long n,
                  void* stream);

/* refine_tip_segmentation (...videos_v3.py:32-42): out = mask inside the FILLED two largest external
 * contours (area descending, ties in findContours order) whose contourArea >= area_threshold -- i.e.
 * the selected components plus whatever is nested in their holes.  Must follow mfc_trace_contours on
 * the same mask: labels / rec / n_contours are that call's scratch, records and count, untouched.
 * sel: 2 device ints, receives the raster index of each selected contour's first point or -1. */
int mfc_refine_tip_mask(const uint8_t* mask, int H, int W, const int* labels, const double* rec,
                        int max_contours, const int* n_contours, double area_threshold, int* sel,
                        uint8_t* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Command list: one C call issues a whole pre-built forward (the per-frame SFC pass is
 * ~60 launches).  Replaces the Python-level op-by-op dispatch of nn.Module.forward
 * (models/resunet.py:153-180, models/multiframe_model.py:424-438).  The structs the
 * commands point at are owned by the caller and must stay alive during the call only.
 * ---------------------------------------------------------------------------------- */
#define MFC_OP_CONV 1            /* a = MfcConvDesc*, b = MfcConvIO*  */
#define MFC_OP_GN_FINALIZE 2     /* a = MfcGnArgs*                    */
#define MFC_OP_AFFINE_SILU_ADD 3 /* a = MfcAddArgs*                   */
#define MFC_OP_GATHER 4          /* a = MfcGatherArgs*                */
#define MFC_OP_WARP 5            /* a = MfcWarpArgs*                  */
#define MFC_OP_FUSE_SUM 6        /* a = MfcFuseArgs*                  */
#define MFC_OP_RESIZE 7          /* a = MfcResizeArgs*                */
#define MFC_OP_MAXPOOL2 8        /* a = MfcPoolArgs*                  */
#define MFC_OP_HEATMAP 9         /* a = MfcHeatmapArgs*               */
#define MFC_OP_POINTWISE 10      /* a = MfcPointwiseArgs*             */
#define MFC_OP_RAFT 11           /* a = MfcRaftArgs*                  */

typedef struct MfcGnArgs {
  const float* stats;
  const float* gamma;
  const float* beta;
  float* affine;
  long long pixels;
  int B, stats_per_image, cpad, C, groups;
  float eps;
} MfcGnArgs;

typedef struct MfcAddArgs {
  const void* a;
  const float* affine;
  const void* r;
  void* out;
  long long pixels;
  int B, chunks, dtype;
  int reserved;
  int* overflow;   /* fp16 range guard counter or NULL (see MfcConvIO.overflow) */
} MfcAddArgs;

typedef struct MfcGatherArgs {
  MfcGather g;
  void* dst;
  long long dst_bstride_bytes;
  int B, H, W, dtype;
} MfcGatherArgs;

typedef struct MfcPoolArgs {
  const void* src;
  void* dst;
  long long src_bstride_bytes, dst_bstride_bytes;
  int B, chunks, H, W, dtype, reserved;
} MfcPoolArgs;

typedef struct MfcHeatmapArgs {   /* mfc_heatmap_head as a list command (TernausNet's log_softmax head) */
  const float* logits;
  float* logp;
  float* prob;
  uint8_t* argmax;
  long long pixels;
  int B, N;
} MfcHeatmapArgs;

/* `lane`: 0 = the caller's stream; 1..3 = side streams owned by the library, for commands that are independent of the
 * other lanes (the parallel branches of an HRNet module: each conv there fills only part of the SMs).  MFC_OP_FORK makes
 * the side lanes wait for everything issued so far on lane 0, MFC_OP_JOIN makes lane 0 wait for the side lanes; a list must
 * JOIN before it ends.  Under mfc_graph_capture the lanes become parallel branches of the graph. */
#define MFC_OP_FORK 100
#define MFC_OP_JOIN 101
#define MFC_MAX_LANES 4
typedef struct MfcCmd {
  int op;
  int lane;
  const void* a;
  const void* b;
} MfcCmd;

int mfc_run_list(const MfcCmd* cmds, int n, void* stream);

/* Measurement variant (bench.py's live roofline): the same list with a CUDA event recorded on
 * `stream` around every command; synchronises the stream and writes each command's device time
 * in milliseconds to ms_out[n].  This is the one entry point that creates events and blocks. */
int mfc_run_list_timed(const MfcCmd* cmds, int n, void* stream, float* ms_out);

/* CUDA-graph replay of a command list whose pointers are all static: capture once (on a private stream; kernels are not
 * executed), then one mfc_graph_launch per step on the caller's stream.  The graph holds copies of all kernel arguments;
 * the buffers they point to must stay alive and in place. */
int mfc_graph_capture(const MfcCmd* cmds, int n, void** graph_out);
int mfc_graph_launch(void* graph, void* stream);
int mfc_graph_destroy(void* graph);


#ifdef __cplusplus
}
#endif
#endif /* MFCNET_B200_H_ */
