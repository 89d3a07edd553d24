// conv_tc.cuh -- parameters shared by the tcgen05 implicit-GEMM convolution kernel and its host planner.
#pragma once
#include <vector>

#include "common.cuh"
#include "../../include/mfcnet_b200.h"

namespace mfc {

// warp roles of conv_tc_kernel (one persistent CTA per SM)
constexpr int kOutStageBytes = 2 * 2 * 2 * 2 * 128 * 16;   // [warp group][buffer][row of the pair][plane][128 pixels] x 16 bytes
constexpr int kEpiWarps = 8;                       // warps 0..7 : TMEM -> registers -> global (lane quarter = warp % 4)
constexpr int kMmaWarp0 = kEpiWarps;               // warps 8..9  : tcgen05.mma issue (one elected lane each; warp i
constexpr int kMmaWarps = 2;                       //               issues the runs r = i, i+2, ...: a single thread
                                                   //               cannot issue N=16 MMAs at the tensor pipe's rate)
constexpr int kProdWarp0 = kMmaWarp0 + kMmaWarps;  // warps 10..15: tile staging (global -> transform -> smem)
constexpr int kProdWarps = 6;                      // 16 warps = 512 threads: 128 registers each
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kConvThreads = (kEpiWarps + kMmaWarps + kProdWarps) * 32;
constexpr int kMaxStages = 8;
constexpr int kResDepth = 4;                       // residual prefetch ring depth (steps) of an epilogue warp

struct ConvTiling {
  int TH, TW;        // output tile
  int P;             // smem row pitch in pixels = TW + (kw-1)/stride
  int R;             // 128-pixel MMA runs covering the flattened (row, P) tile
  int rows_sub;      // rows loaded per parity sub-plane = TH + (kh-1)/stride
  int slots_sub;     // 16-byte pixel slots per parity sub-plane
  int CBc;           // 8-channel planes per K stage (even)
  int kstages;       // K stages per work item
  int nstages;       // depth of the shared-memory stage ring
  int nacc;          // TMEM accumulator buffers (2 = the epilogue of item i overlaps the MMAs of item i+1)
  int pair;          // 1: single 8-channel input plane, stride 1: one K=16 MMA covers the taps (ky,kx) and (ky,kx+1)
                     //    (second K half = the same plane one pixel to the right, LBO = 16 bytes) -> ceil(kw/2) entries per row
  int entries;       // MMA entries per K step: kh*kw, or kh*ceil(kw/2) with tap pairing
  int kacc;          // K-split accumulator sets per buffer (1, 2 or 4): consecutive taps rotate over them so that
                     // back-to-back MMAs are independent even when the tile has a single run; the epilogue sums them
  int slide;         // 1: "sliding accumulate" mode for stride-1 kh>1 convs with few output channels.  P = 128, so an MMA run
                     //    is exactly one INPUT row of the tile; the kh vertical taps become N-groups of one wide MMA
                     //    (N = window * NB): the MMA of input row r and horizontal tap kx adds W[ky][kx] * in[r] into the
                     //    accumulator columns of the output rows r-ky, which are adjacent column blocks [row][NB] in TMEM.
                     //    kh x fewer MMAs and A-operand reads than one MMA per tap; accumulators are zeroed by the epilogue.
  int nrows_b;       // rows (N) of one packed B block: NB, or kh*NB in slide mode (N-group g' = kh-1-ky)
  int P_lo, rows_lo;          // nearest-x2 convs with TMA: extent of the LOW-resolution box staged per plane ...
  uint32_t lo_plane_bytes;    // ... its size, and where the boxes sit inside a stage (after the planes and the streamed weights);
  uint32_t off_lo;            //     the producer warps expand them x2 into the planes
  int tma;           // 1: stride-1, no-upsample conv: the halo tile of every plane is ONE TMA box load (zero fill = padding)
  int b_resident;    // 1: the whole packed weight blob is loaded once per CTA; 0: streamed with each stage
  int tiles_x, tiles_y;
  int NB, nblk;
  int ksteps;        // total 16-channel K steps = ceil(cin_chunks/2)
  int cin_chunks;
  uint32_t plane_bytes;    // bytes of one 8-channel plane (all parity sub-planes)
  uint32_t a_stage_bytes;  // CBc * plane_bytes
  uint32_t b_stage_bytes;  // (CBc/2) * taps * 2*NB*16
  uint32_t stage_bytes;    // a_stage_bytes (+ b_stage_bytes when streamed), 128-byte aligned
  uint32_t smem_bytes;
  uint32_t tmem_cols;      // allocated (power of two >= 32)
  uint32_t acc_cols;       // kacc * R * NB, columns of one accumulator buffer
  uint32_t off_resring;    // residual prefetch rings of the epilogue warps (only with MFC_CONV_HAS_RESIDUAL)
  uint32_t off_scale, off_stats, off_bres, off_stage;  // smem carve-up (bytes from the 128B-aligned base)
  uint32_t off_ostage;     // != 0: output staging ring of the slide16 statistics epilogue (kOutStageBytes; bulk-copy stores)
  int grid;                // persistent CTAs
};

struct ConvParams {
  int B, Hin, Win, Hout, Wout, Cout;
  int kh, kw, stride, pad, upsample, act;
  int in_off_y, in_off_x, out_stride, out_off_y, out_off_x;  // output-parity mode of the k4 s2 transposed conv
  int nsrc;
  const uint8_t* src_ptr[MFC_MAX_SRC];
  const float* src_aff[MFC_MAX_SRC];
  long long src_bs[MFC_MAX_SRC];
  int src_end[MFC_MAX_SRC];  // exclusive prefix end (in chunks) of each source
  ConvTiling t;
  FastDiv divP, div_nblk, div_tx, div_ty;
  uint32_t idesc;
  const uint8_t* w;
  const float* scale;
  const float* shift;
  const uint8_t* res;
  const float* res_aff;
  long long res_bs;
  uint8_t* y;
  uint8_t* y_lo;   // optional rounding-residue output (same geometry as y)
  long long y_bs;
  float* y_nchw;
  float* stats;
  const float* head_w;   // fused 1x1 head (fast NCHW epilogue only): [head_n][16] weights, [head_n] bias (or NULL), see MfcConvIO
  const float* head_b;
  int head_n;
  int* ovf;       // fp16 outputs only: incremented by every epilogue warp that stored a value beyond +-65504 (or NULL)
  alignas(64) CUtensorMap tmap[MFC_MAX_SRC];  // t.tma: source i as the 5-D tensor (8 ch, W, H, chunk, sample)
  int direct;     // no work for the producer warps between the TMA landing and the MMAs (no GroupNorm-on-load, no x2 expansion):
                  // the MMA warps wait on the TMA barrier themselves and ONE thread keeps the loads issued -- one barrier hop
                  // and six warps' per-stage hand-off less per work item
  int tma_wide;   // stride-1 sources are described to TMA as 4-D tensors of 8-byte elements (2 per pixel): the innermost box
                  // dimension is a whole tile row (P*16 bytes) instead of one 16-byte pixel, i.e. one L2 request stream per
                  // row instead of one per pixel
  int acc_init;   // slide mode without an epilogue scale: the accumulators are initialised with the per-channel shift (instead of
                  // zeros) at kernel start and whenever the epilogue drains them, so values leave TMEM finished
  int epi_fast;   // NB == 16, one N-block, out_stride 1, and the pixel of (run, lane) is affine in the run index: sliding mode
                  // (a run = one output row) or full-width tiles of a halo-free conv (a run = 128 consecutive pixels)
  int chunk;             // work items are dealt to the CTAs in chunks of `chunk` consecutive items (same sample, adjacent tiles):
  int full_items;        // fewer GroupNorm-record flushes and sample switches per CTA; items >= full_items (the tail) go round-robin
  FastDiv div_grid, div_chunk;
  int reverse, total_items;  // reverse: walk the work items from the last sample to the first (see MFC_CONV_REVERSE_ORDER)
  int debug;  // measurement only (MFC_CONV_DEBUG): bit0 skip producer copies, bit1 skip epilogue body, bit2 skip MMAs,
              // bit3 role timing
};

// host planner
int conv_nb(int cout, int* nblk);
void conv_candidates(const MfcConvDesc& d, std::vector<ConvTiling>& out, bool sorted_by_cost = false);
void conv_shortlist(const MfcConvDesc& d, int per_bucket, std::vector<ConvTiling>& out);

}  // namespace mfc
