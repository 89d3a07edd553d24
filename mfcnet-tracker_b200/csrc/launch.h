// launch.h -- internal launcher prototypes shared between the kernel translation units and api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mfcnet_b200.h"
#include "conv_tc.cuh"

namespace mfc {

// conv_tc.cu
cudaError_t launch_conv(const ConvParams& p, bool bf16, cudaStream_t st);
cudaError_t launch_pack_weights(const float* w, int Cout, int Cin_w, int taps, const int* chan_map, int cin_chunks,
                                int ksteps, int NB, int nblk, int pair_kw, int taps_w, int slide_kh, int kw, const float* scale,
                                void* out, bool bf16, cudaStream_t st);

// pointwise.cu
cudaError_t launch_gather(const MfcGather& g, void* dst, long long dst_bs, int B, int H, int W, bool bf16, cudaStream_t st);
cudaError_t launch_c8_to_nchw(const void* src, long long src_bs, float* dst, int B, int C, int H, int W, bool bf16, cudaStream_t st);
cudaError_t launch_weight_standardize(const float* w, float* out, int Cout, int fan_in, float eps, cudaStream_t st);
cudaError_t launch_bn_fold(const float* g, const float* b, const float* m, const float* v, const float* cb, float eps,
                           float* scale, float* shift, int C, cudaStream_t st);
cudaError_t launch_gn_finalize(const float* stats, int B, int tiles, int cpad, int C, int groups, long long pixels,
                               const float* gamma, const float* beta, float eps, float* affine, cudaStream_t st);
cudaError_t launch_affine_silu_add(const void* a, const float* affine, const void* r, void* out, int B, int chunks,
                                   long long pixels, bool bf16, int* ovf, cudaStream_t st);
bool silu_accurate();  // built with -DMFC_SILU_ACCURATE: two-MUFU SiLU (ex2 + rcp) instead of tanh.approx in every kernel

// fusion_ops.cu
cudaError_t launch_flow_warp(const MfcWarpArgs& a, cudaStream_t st);
cudaError_t launch_heatmap_head(const float* logits, int B, int N, long long pixels, float* logp, float* prob, uint8_t* amax,
                                cudaStream_t st);

cudaError_t launch_argmax_u8(const float* x, int B, int N, long long pixels, uint8_t* out, cudaStream_t st);

// resample.cu
cudaError_t launch_maxpool2(const void* src, long long src_bs, void* dst, long long dst_bs, int B, int chunks, int H, int W,
                            bool bf16, cudaStream_t st);
cudaError_t launch_fuse_sum(const MfcFuseArgs& a, cudaStream_t st);
cudaError_t launch_bilinear_resize(const float* src, int B, int C, int Hin, int Win, int Hout, int Wout, float* dst_nchw,
                                   void* dst_c8, long long c8_bs, bool bf16, cudaStream_t st);

// loss.cu
int loss_blocks(int B, long long pixels);
cudaError_t launch_segmentation_loss(const float* logits, const long long* target, const float* cw, int B, int N, long long pixels,
                                     float w_nll, float w_jacc, double* partials, float* out, cudaStream_t st);

// train_ops.cu
cudaError_t launch_segmentation_loss_bwd(const float* logits, const long long* target, const float* cw, int B, int N, long long pixels,
                                         float w_nll, float w_jacc, float scale, const double* partials, int n_partials, float* coef,
                                         float* dlogits, cudaStream_t st);
cudaError_t launch_segmentation_loss_sums(const float* logits, const long long* target, const float* cw, int B, int N, long long pixels,
                                          double* partials, double* sums, cudaStream_t st);
cudaError_t launch_segmentation_loss_from_sums(const double* sums, int N, float w_nll, float w_jacc, float* out, cudaStream_t st);
cudaError_t launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd,
                        float bc1, float bc2_sqrt, float grad_scale, cudaStream_t st);

// correlation.cu
cudaError_t launch_correlation(const float* first, const float* second, float* out, int B, int C, int H, int W, int max_disp,
                               int stride2, int exact_order, cudaStream_t st);

cudaError_t launch_correlation_bwd(const float* first, const float* second, const float* grad_out, float* grad_first,
                                   float* grad_second, int B, int C, int H, int W, int max_disp, int stride2, cudaStream_t st);

// correlation_tma.cu (stride-1, max displacement 4: TMA-fed register-tiled kernel)
bool correlation_tma_supported(int C, int H, int W, int max_disp, int stride2);
void correlation_tma_boxes(int H, int W, int stride2, unsigned* box1, unsigned* box2, unsigned* estride);
cudaError_t launch_correlation_tma(const CUtensorMap& map1, const CUtensorMap& map2, float* out, int B, int C, int H, int W,
                                   int stride2, cudaStream_t st);

// unflow_ops.cu
cudaError_t launch_unflow_prep(const float* rgb, float* out, int B, long long pixels, cudaStream_t st);
cudaError_t launch_nchw_to_c8(const float* src, void* dst, long long dst_bs, int B, int C, long long pixels, bool bf16, cudaStream_t st);
cudaError_t launch_unflow_warp(const float* second, const float* flow, const float* first, float* warped, float* absdiff, int B, int C,
                               int H, int W, cudaStream_t st);
cudaError_t launch_unflow_upscale(const float* x, const float* w, float* out, int B, int h, int wd, float scale, cudaStream_t st);

cudaError_t launch_resize_u8(const uint8_t* src, long long frame_stride, int h, int w, int C, uint8_t* dst, int B, int H, int W,
                             cudaStream_t st);
cudaError_t launch_bgr2gray_u8(const uint8_t* bgr, long long frame_stride, uint8_t* out, int B, long long pixels, cudaStream_t st);
cudaError_t launch_ingest_gray(const uint8_t* gray, float* out, long long n, cudaStream_t st);

// raft_ops.cu
cudaError_t launch_pointwise(int kind, const void* a, const float* a_aff, const void* r, const float* r_aff, void* out, void* out2, int B,
                             int chunks, long long pixels, int relu_a, int relu_out, bool bf16, cudaStream_t st);
cudaError_t launch_raft_corr_volume(const float* f1, const float* f2, float* out, int B, int C, int HW, float scale, cudaStream_t st);
cudaError_t launch_raft_corr_pool(const float* in, float* out, long long N, int h, int w, cudaStream_t st);
cudaError_t launch_raft_lookup(const float* const* lvl, const float* flow, void* out, int B, int h, int w, int levels, int radius, int chunks,
                               bool bf16, cudaStream_t st);
cudaError_t launch_raft_flow_add(float* flow, const float* delta, long long n, cudaStream_t st);
cudaError_t launch_raft_resize_ac(const float* in, float* out, int BC, int h, int w, int H, int W, float mult, cudaStream_t st);
cudaError_t launch_raft_upsample(const float* flow, const float* mask, float* out, int B, int h, int w, float mult, cudaStream_t st);

// ingest.cu
cudaError_t launch_ingest_rgb(const uint8_t* bgr, long long frame_stride, float* out, int B, long long pixels, const float* mean,
                              const float* stdv, cudaStream_t st);
cudaError_t launch_ingest_depth(const uint8_t* bgr, long long frame_stride, float* out, int B, long long pixels, cudaStream_t st);

// localize.cu
cudaError_t launch_gaussian_blur(const float* heat, float* tmp, float* out, int B, int H, int W, const double* w, int radius,
                                 cudaStream_t st);
cudaError_t launch_localmax_mask(const float* sm, const uint8_t* cls, int cls_id, const uint8_t* fp, int fh, int fw,
                                 uint8_t* mask, int B, int H, int W, cudaStream_t st);
cudaError_t launch_class_mask(const uint8_t* cls, int cls_id, uint8_t* mask, long long n, cudaStream_t st);
cudaError_t launch_trace_contours(const uint8_t* mask, int H, int W, int* labels, double* out, int max_contours, int* n_out,
                                  cudaStream_t st);

cudaError_t launch_threshold_classes(const float* prob, int B, int N, long long pixels, float thr, uint8_t* out, cudaStream_t st);
cudaError_t launch_mask_heat(const float* heat, const uint8_t* cls, int cls_id, float* out, long long n, cudaStream_t st);
cudaError_t launch_refine_tip_mask(const uint8_t* mask, int H, int W, const int* labels, const double* rec, int max_contours,
                                   const int* n_contours, double area_threshold, int* sel, uint8_t* out, cudaStream_t st);

cudaError_t launch_top_contours(const double* rec, const int* n_contours, int max_contours, int W, double* top, cudaStream_t st);

}  // namespace mfc
