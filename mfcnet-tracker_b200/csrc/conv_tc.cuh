// conv_tc.cuh -- parameters shared by the tcgen05 implicit-GEMM convolution kernel and its host planner.
#pragma once
#include "common.cuh"
#include "../../include/mfcnet_b200.h"

namespace mfc {

struct ConvTiling {
  int TH, TW;        // output tile
  int P;             // smem row pitch in pixels = TW + (kw-1)/stride
  int R;             // 128-pixel MMA runs covering the flattened (row, P) tile
  int rows_sub;      // rows loaded per parity sub-plane = TH + (kh-1)/stride
  int slots_sub;     // 16-byte pixel slots per parity sub-plane
  int CBc;           // 8-channel planes per K stage (even)
  int kstages;
  int nbuf;          // A/B stage buffers (1 or 2)
  int tiles_x, tiles_y;
  int NB, nblk;
  int ksteps;        // total 16-channel K steps = ceil(cin_chunks/2)
  int cin_chunks;
  uint32_t plane_bytes;    // bytes of one 8-channel plane (all parity sub-planes)
  uint32_t a_stage_bytes;  // CBc * plane_bytes
  uint32_t b_stage_bytes;  // (CBc/2) * taps * 2*NB*16
  uint32_t smem_bytes;
  uint32_t tmem_cols;
  uint32_t off_scale, off_stats, off_a, off_b;  // smem carve-up (bytes from the 128B-aligned base)
  int ctas_per_sm;
};

struct ConvParams {
  int B, Hin, Win, Hout, Wout, Cout;
  int kh, kw, stride, pad, upsample, act;
  int nsrc;
  const uint8_t* src_ptr[MFC_MAX_SRC];
  const float* src_aff[MFC_MAX_SRC];
  long long src_bs[MFC_MAX_SRC];
  int src_end[MFC_MAX_SRC];  // exclusive prefix end (in chunks) of each source
  ConvTiling t;
  FastDiv divP;
  uint32_t idesc;
  const uint8_t* w;
  const float* scale;
  const float* shift;
  const uint8_t* res;
  const float* res_aff;
  long long res_bs;
  uint8_t* y;
  long long y_bs;
  float* y_nchw;
  float* stats;
};

// host planner (conv_plan.cpp part of api.cu)
int conv_nb(int cout, int* nblk);
bool conv_choose_tiling(const MfcConvDesc& d, ConvTiling& out);

}  // namespace mfc
