// api.cu -- the C ABI of libmfcnet_b200.so (include/mfcnet_b200.h): argument validation, the conv
// planner cache, and dispatch into the kernel launchers.  No allocation, no synchronisation.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <mutex>
#include <tuple>
#include <vector>

#include "launch.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(MFC_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

// 0 when the CURRENT device is compute capability 10.x.  Cached per device.
int arch_ok() {
  static int cache[64];  // 0 unknown, 1 ok, 2 bad
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev < 0 || dev >= 64) return fail(MFC_EINVAL, "device index %d out of range", dev);
  if (cache[dev] == 0) {
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
    cache[dev] = (major == 10) ? 1 : 2;
  }
  if (cache[dev] != 1)
    return fail(MFC_EARCH, "device %d is not sm_100 (B200); libmfcnet_b200 has no fallback path", dev);
  return MFC_OK;
}

#define MFC_REQUIRE_ARCH()          \
  do {                              \
    int rc_ = arch_ok();            \
    if (rc_ != MFC_OK) return rc_;  \
  } while (0)

#define MFC_LAUNCH(expr, what)                       \
  do {                                               \
    cudaError_t e_ = (expr);                         \
    if (e_ != cudaSuccess) return cuda_fail(e_, what); \
    return MFC_OK;                                   \
  } while (0)

inline bool dtype_ok(int dt) { return dt == MFC_F16 || dt == MFC_BF16; }

// ---- conv planner cache ---------------------------------------------------------------------
// A plan (tiling) is a pure function of the geometry key.  It comes from, in this order: (1) an entry this process
// measured with mfc_conv2d_autotune BEFORE the geometry was first planned, (2) the imported tuning table
// (mfc_conv2d_plan_import: choices measured once on a B200 and committed with the package), (3) the cost model.
// Once a plan has been handed out (query / pack / fwd) it is frozen for the life of the process: packed weights,
// statistics buffers and gn_finalize's record count all depend on it.  (2) and (3) are deterministic, so two processes
// that do not tune live compute bit-identical results.
using PlanKey = std::tuple<int, int, int, int, int, int, int, int, int, int, int, int, int, int, int, int>;
struct PlanEntry {
  mfc::ConvTiling t;
  bool handed_out = false;
  bool tuned = false;   // measured in this process or resolved from the table (exported by mfc_conv2d_plan_export)
};
struct TableChoice {   // what identifies one candidate of conv_enumerate for a key
  int TH, TW, slide, CBc, NB, nstages;
};
std::mutex g_plan_mu;
std::map<PlanKey, PlanEntry> g_plans;
std::map<PlanKey, TableChoice> g_table;

int validate_desc(const MfcConvDesc* d) {
  if (!d) return fail(MFC_EINVAL, "conv: null descriptor");
  if (d->B < 1 || d->Hin < 1 || d->Win < 1 || d->Hout < 1 || d->Wout < 1 || d->Cout < 1)
    return fail(MFC_EINVAL, "conv: non-positive dimension");
  if (d->kh < 1 || d->kw < 1 || d->kh > 11 || d->kw > 11) return fail(MFC_EINVAL, "conv: kernel %dx%d unsupported", d->kh, d->kw);
  if (d->stride != 1 && d->stride != 2) return fail(MFC_EINVAL, "conv: stride %d unsupported", d->stride);
  if (d->upsample != 1 && d->upsample != 2) return fail(MFC_EINVAL, "conv: upsample %d unsupported", d->upsample);
  if (d->act < 0 || d->act > 2) return fail(MFC_EINVAL, "conv: act %d unsupported", d->act);
  if (d->pad_br < 0 || d->pad_br > 5) return fail(MFC_EINVAL, "conv: pad_br %d unsupported", d->pad_br);
  if (!dtype_ok(d->dtype)) return fail(MFC_EINVAL, "conv: dtype %d unsupported", d->dtype);
  if (d->nsrc < 1 || d->nsrc > MFC_MAX_SRC) return fail(MFC_EINVAL, "conv: nsrc %d out of range", d->nsrc);
  if (d->pad < 0 || d->pad > 5) return fail(MFC_EINVAL, "conv: pad %d unsupported", d->pad);
  const int Hup = d->Hin * d->upsample, Wup = d->Win * d->upsample;
  // ordinary convs: in_off_{y,x} REDUCE the padding of one axis (rectangular kernels: a 1x5 conv with padding (0, 2) is
  // pad = 2, in_off_y = 2); the input window of output row oy starts at oy*stride - pad + in_off_y either way
  const bool parity_mode = d->out_stride == 2;
  const int pv = d->pad - (parity_mode ? 0 : d->in_off_y), ph = d->pad - (parity_mode ? 0 : d->in_off_x);
  int ho = (Hup + 2 * pv + d->pad_br - d->kh) / d->stride + 1, wo = (Wup + 2 * ph + d->pad_br - d->kw) / d->stride + 1;
  if (d->out_stride != 0 && d->out_stride != 1 && d->out_stride != 2) return fail(MFC_EINVAL, "conv: out_stride %d unsupported", d->out_stride);
  if (parity_mode && ((unsigned)d->in_off_y > 1u || (unsigned)d->in_off_x > 1u || (unsigned)d->out_off_y > 1u || (unsigned)d->out_off_x > 1u))
    return fail(MFC_EINVAL, "conv: parity offsets must be 0 or 1");
  if (!parity_mode && (d->in_off_y < 0 || d->in_off_x < 0 || d->in_off_y > d->pad || d->in_off_x > d->pad))
    return fail(MFC_EINVAL, "conv: in_off (%d, %d) must lie in [0, pad]", d->in_off_y, d->in_off_x);
  if (d->out_stride == 2) {  // one output parity of ConvTranspose2d(4,2,1): a 2x2 pad-1 conv restricted to Hin x Win outputs
    if (d->kh != 2 || d->kw != 2 || d->stride != 1 || d->pad != 1 || d->upsample != 1)
      return fail(MFC_EINVAL, "conv: out_stride 2 is the k4 s2 p1 transposed-conv parity mode (k=2, s=1, p=1)");
    ho = d->Hin;
    wo = d->Win;
  } else if (d->out_off_y || d->out_off_x) {
    return fail(MFC_EINVAL, "conv: output offsets need out_stride 2");
  } else if ((d->in_off_y || d->in_off_x) && (d->stride != 1 || d->upsample != 1)) {
    return fail(MFC_EINVAL, "conv: per-axis padding (in_off) needs stride 1 and no upsampling");
  }
  if (ho != d->Hout || wo != d->Wout)
    return fail(MFC_EINVAL, "conv: Hout/Wout %dx%d inconsistent with input %dx%d (expected %dx%d)", d->Hout, d->Wout, Hup, Wup, ho, wo);
  for (int i = 0; i < d->nsrc; ++i)
    if (d->src[i].nchunks < 1) return fail(MFC_EINVAL, "conv: source %d has no channel planes", i);
  return MFC_OK;
}

constexpr int kPlanFlags = MFC_CONV_HAS_RESIDUAL | MFC_CONV_WANT_STATS | MFC_CONV_WANT_HEAD;

PlanKey plan_key(const MfcConvDesc* d) {
  int chunks = 0, aff = 0;
  for (int i = 0; i < d->nsrc; ++i) {
    chunks += d->src[i].nchunks;
    aff |= d->src[i].affine != nullptr;
  }
  return PlanKey{d->B, d->Hin, d->Win, d->Hout, d->Wout, d->Cout, d->kh, d->kw, d->stride, d->pad, d->upsample, chunks, aff,
                 d->reserved & kPlanFlags, d->dtype, d->out_stride == 2 ? 2 : 1};
}

bool same_choice(const mfc::ConvTiling& t, const TableChoice& c) {
  return t.TH == c.TH && t.TW == c.TW && t.slide == c.slide && t.CBc == c.CBc && t.NB == c.NB && t.nstages == c.nstages;
}

// pixel of (run, lane) affine in the run index: sliding mode, or full-width tiles of a halo-free conv
bool tiling_fast_epilogue(const MfcConvDesc* d, const mfc::ConvTiling& t) {
  const bool affine_rows = t.slide || (t.P == t.TW && t.TW == d->Wout && t.tiles_x == 1);
  return t.NB == 16 && t.nblk == 1 && d->out_stride != 2 && t.kacc == 1 && affine_rows;
}

// statistics need the padded Cout to fit the per-warp shared-memory scratch (see conv_fwd_tiled); a fused head needs the
// fast epilogue
bool tiling_serves(const MfcConvDesc* d, const mfc::ConvTiling& t) {
  if ((d->reserved & MFC_CONV_WANT_STATS) && t.NB * t.nblk > 256) return false;
  if ((d->reserved & MFC_CONV_WANT_HEAD) && !tiling_fast_epilogue(d, t)) return false;
  return true;
}

// g_plan_mu held.  Resolves the plan of `d` (table choice, else the cost model) without marking it handed out.
int resolve_locked(const MfcConvDesc* d, const PlanKey& key, PlanEntry** out) {
  auto it = g_plans.find(key);
  if (it == g_plans.end()) {
    PlanEntry e;
    memset(&e.t, 0, sizeof(e.t));
    bool found = false;
    auto tb = g_table.find(key);
    if (tb != g_table.end()) {
      std::vector<mfc::ConvTiling> all;
      mfc::conv_candidates(*d, all);
      for (const auto& t : all)
        if (same_choice(t, tb->second) && tiling_serves(d, t)) {
          e.t = t;
          e.tuned = true;
          found = true;
          break;
        }
    }
    if (!found) {
      std::vector<mfc::ConvTiling> ranked;
      mfc::conv_candidates(*d, ranked, /*sorted_by_cost=*/true);
      for (const auto& t : ranked)
        if (tiling_serves(d, t)) {
          e.t = t;
          found = true;
          break;
        }
    }
    if (!found) return fail(MFC_EINVAL, "conv: no tiling fits shared memory / TMEM for this shape");
    it = g_plans.emplace(key, e).first;
  }
  *out = &it->second;
  return MFC_OK;
}

int get_tiling(const MfcConvDesc* d, mfc::ConvTiling* out) {
  const PlanKey key = plan_key(d);
  std::lock_guard<std::mutex> g(g_plan_mu);
  PlanEntry* e = nullptr;
  int rc = resolve_locked(d, key, &e);
  if (rc != MFC_OK) return rc;
  e->handed_out = true;
  *out = e->t;
  return MFC_OK;
}

// ---- TMA descriptors ---------------------------------------------------------------------------
// cuTensorMapEncodeTiled lives in libcuda; it is resolved through the runtime (no link-time dependency on the
// driver library, so the .so still loads -- and plans -- on a box without a GPU).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

// Source i of a stride-1 conv as the 5-D tensor (8 channels = one 16-byte pixel slot, W, H, chunk, sample); the box is
// one halo tile of one 8-channel plane: P x rows_sub slots, dense in shared memory = the tcgen05 operand layout.
int encode_source_maps(const MfcConvDesc* d, mfc::ConvParams* p) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(MFC_ECUDA, "conv: cuTensorMapEncodeTiled is not available from this driver");
  static const int no_wide = getenv("MFC_CONV_TMA_WIDE") ? (atoi(getenv("MFC_CONV_TMA_WIDE")) == 0) : 0;  // measurement switch
  const int box_w = d->upsample == 2 ? p->t.P_lo : p->t.P, box_h = d->upsample == 2 ? p->t.rows_lo : p->t.rows_sub;
  // Stride 1: a tile row is contiguous in memory (P pixels x 16 bytes).  Described with 16-byte pixels as the innermost
  // dimension the copy engine walks the box pixel by pixel (measured: ~10 bytes/clk per SM, 2.9 TB/s over the chip); with
  // 8-byte elements the innermost box dimension is the whole row (up to 256 elements = 128 pixels).
  p->tma_wide = (!no_wide && d->stride == 1 && box_w * 2 <= 256) ? 1 : 0;
  for (int i = 0; i < d->nsrc; ++i) {
    const cuuint64_t plane = (cuuint64_t)d->Hin * d->Win * 16;
    const cuuint64_t bstride = d->B > 1 ? (cuuint64_t)d->src[i].batch_stride : plane * (cuuint64_t)d->src[i].nchunks;
    CUresult r;
    if (p->tma_wide) {
      cuuint64_t dims[4] = {(cuuint64_t)d->Win * 2, (cuuint64_t)d->Hin, (cuuint64_t)d->src[i].nchunks, (cuuint64_t)d->B};
      cuuint64_t strides[3] = {(cuuint64_t)d->Win * 16, plane, bstride};
      cuuint32_t box[4] = {(cuuint32_t)box_w * 2, (cuuint32_t)box_h, 1, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      r = enc(&p->tmap[i], CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(d->src[i].ptr), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dims[5] = {8, (cuuint64_t)d->Win, (cuuint64_t)d->Hin, (cuuint64_t)d->src[i].nchunks, (cuuint64_t)d->B};
      cuuint64_t strides[4] = {16, (cuuint64_t)d->Win * 16, plane, bstride};
      const cuuint32_t st = (cuuint32_t)d->stride;  // stride 2: box extents in tensor elements, every second one is taken
      cuuint32_t box[5] = {8, (cuuint32_t)box_w * st, (cuuint32_t)box_h * st, 1, 1};
      cuuint32_t estr[5] = {1, st, st, 1, 1};
      r = enc(&p->tmap[i], CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, const_cast<void*>(d->src[i].ptr), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS)
      return fail(MFC_ECUDA, "conv: cuTensorMapEncodeTiled failed for source %d (CUresult %d; %dx%d, %d chunks, box %dx%d, wide %d)", i, (int)r,
                  d->Hin, d->Win, d->src[i].nchunks, box_w, box_h, p->tma_wide);
  }
  return MFC_OK;
}

int conv_fwd_tiled(const MfcConvDesc* d, const MfcConvIO* io, const mfc::ConvTiling& tiling, void* stream) {
  int rc = MFC_OK;
  if (!io || !io->w_packed) return fail(MFC_EINVAL, "conv: null io / weights");
  if (!io->y_c8 && !io->y_nchw) return fail(MFC_EINVAL, "conv: no output buffer");
  mfc::ConvParams p;
  memset(&p, 0, sizeof(p));
  p.t = tiling;
  p.B = d->B; p.Hin = d->Hin; p.Win = d->Win; p.Hout = d->Hout; p.Wout = d->Wout; p.Cout = d->Cout;
  p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad = d->pad; p.upsample = d->upsample; p.act = d->act;
  p.in_off_y = d->in_off_y; p.in_off_x = d->in_off_x; p.out_stride = d->out_stride == 2 ? 2 : 1;
  p.out_off_y = d->out_off_y; p.out_off_x = d->out_off_x;
  if ((io->residual != nullptr) != ((d->reserved & MFC_CONV_HAS_RESIDUAL) != 0))
    return fail(MFC_EINVAL, "conv: io->residual and the descriptor's MFC_CONV_HAS_RESIDUAL flag disagree");
  if (p.out_stride == 2 && (io->residual || io->stats)) return fail(MFC_EINVAL, "conv: transposed-conv parity mode has no residual / statistics");
  p.nsrc = d->nsrc;
  int end = 0;
  for (int i = 0; i < MFC_MAX_SRC; ++i) {
    if (i < d->nsrc) {
      if (!d->src[i].ptr) return fail(MFC_EINVAL, "conv: source %d is null", i);
      if (((uintptr_t)d->src[i].ptr & 15) || (d->src[i].batch_stride & 15)) return fail(MFC_EINVAL, "conv: source %d not 16-byte aligned", i);
      p.src_ptr[i] = (const uint8_t*)d->src[i].ptr;
      p.src_aff[i] = d->src[i].affine;
      p.src_bs[i] = d->src[i].batch_stride;
      end += d->src[i].nchunks;
    }
    p.src_end[i] = end;
  }
  p.divP = mfc::make_fastdiv((uint32_t)p.t.P);
  p.div_nblk = mfc::make_fastdiv((uint32_t)p.t.nblk);
  p.div_tx = mfc::make_fastdiv((uint32_t)p.t.tiles_x);
  p.div_ty = mfc::make_fastdiv((uint32_t)p.t.tiles_y);
  p.idesc = mfc::make_idesc_f16(p.t.NB, d->dtype == MFC_BF16);
  p.reverse = (d->reserved & MFC_CONV_REVERSE_ORDER) ? 1 : 0;
  p.total_items = p.B * p.t.tiles_x * p.t.tiles_y * p.t.nblk;
  {
    static const int chunk_env = getenv("MFC_CONV_CHUNK") ? atoi(getenv("MFC_CONV_CHUNK")) : 4;
    const int c = chunk_env < 1 ? 1 : (chunk_env > 64 ? 64 : chunk_env);
    const int per_group = p.t.grid * c;
    p.chunk = (p.t.nblk == 1 && c > 1 && p.total_items >= 2 * per_group) ? c : 1;
    p.full_items = p.chunk > 1 ? (p.total_items / per_group) * per_group : 0;
    p.div_grid = mfc::make_fastdiv((uint32_t)p.t.grid);
    p.div_chunk = mfc::make_fastdiv((uint32_t)p.chunk);
  }
  p.w = (const uint8_t*)io->w_packed;
  p.scale = io->scale;
  p.shift = io->shift;
  p.res = (const uint8_t*)io->residual;
  p.res_aff = io->res_affine;
  p.res_bs = io->res_batch_stride;
  p.y = (uint8_t*)io->y_c8;
  p.y_lo = (uint8_t*)io->y_lo;
  if (io->y_lo && (!io->y_c8 || ((uintptr_t)io->y_lo & 15))) return fail(MFC_EINVAL, "conv: y_lo needs y_c8 and 16-byte alignment");
  p.y_bs = io->y_batch_stride;
  p.y_nchw = io->y_nchw;
  p.stats = io->stats;
  p.ovf = io->overflow;
  {
    static const int no_fast = getenv("MFC_CONV_EPI_FAST") ? (atoi(getenv("MFC_CONV_EPI_FAST")) == 0) : 0;  // measurement switch
    // the FAST kernels exist for the statistics mode and the fp32-output (NCHW / fused head) mode (launch_conv_mode)
    const bool mode_ok = !io->residual && ((io->stats != nullptr) != (io->y_nchw != nullptr));
    p.epi_fast = (tiling_fast_epilogue(d, p.t) && mode_ok && (!no_fast || io->head_w)) ? 1 : 0;
    // shift-initialised accumulators: sliding mode, FAST kernels only (the general epilogue keeps its scale/shift pass)
    p.acc_init = (p.epi_fast && p.t.slide && io->scale == nullptr && io->shift != nullptr) ? 1 : 0;
  }
  if (io->head_w) {
    if (!(d->reserved & MFC_CONV_WANT_HEAD) || !p.epi_fast || !io->y_nchw || io->stats || io->residual || io->head_n < 1 || io->head_n > 8 ||
        d->Cout > 16)
      return fail(MFC_EINVAL, "conv: fused head needs MFC_CONV_WANT_HEAD, Cout <= 16, y_nchw, 1 <= head_n <= 8, no stats / residual");
    p.head_w = io->head_w;
    p.head_b = io->head_b;
    p.head_n = io->head_n;
  }
  {
    static const int dbg = (getenv("MFC_CONV_DEBUG") ? atoi(getenv("MFC_CONV_DEBUG")) : 0) | (mfc::silu_accurate() ? 16 : 0);
    p.debug = dbg;
  }
  if (p.t.tma) {
    rc = encode_source_maps(d, &p);
    if (rc != MFC_OK) return rc;
    static const int no_direct = getenv("MFC_CONV_DIRECT") ? (atoi(getenv("MFC_CONV_DIRECT")) == 0) : 0;  // measurement switch
    bool any_aff = false;
    for (int i = 0; i < d->nsrc; ++i) any_aff = any_aff || d->src[i].affine != nullptr;
    p.direct = (!no_direct && !any_aff && d->upsample == 1 && !(p.debug & 1)) ? 1 : 0;
  }
  if (p.stats && p.t.NB * p.t.nblk > 256) return fail(MFC_EINVAL, "conv: GroupNorm statistics need Cout <= 256");
  if (((uintptr_t)p.w & 15) || ((uintptr_t)p.y & 15) || ((uintptr_t)p.res & 15) || (p.y_bs & 15) || (p.res_bs & 15))
    return fail(MFC_EINVAL, "conv: weights / output / residual not 16-byte aligned");
  MFC_LAUNCH(mfc::launch_conv(p, d->dtype == MFC_BF16, (cudaStream_t)stream), "conv2d_fwd");
}


}  // namespace

extern "C" {

int mfc_abi_version(void) { return MFC_ABI_VERSION; }
const char* mfc_last_error(void) { return g_err; }

int mfc_device_check(int device) {
  int major = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
  if (major != 10) return fail(MFC_EARCH, "device %d has compute capability major %d, need 10 (B200)", device, major);
  return MFC_OK;
}

// ---- layout ------------------------------------------------------------------------------------
int mfc_gather_nchw_to_c8(const MfcGather* g, void* dst_chunk, long long dst_bstride_bytes, int B, int H, int W, int dtype,
                          void* stream) {
  MFC_REQUIRE_ARCH();
  if (!g || !dst_chunk || B < 1 || H < 1 || W < 1 || !dtype_ok(dtype)) return fail(MFC_EINVAL, "gather: bad argument");
  MFC_LAUNCH(mfc::launch_gather(*g, dst_chunk, dst_bstride_bytes, B, H, W, dtype == MFC_BF16, (cudaStream_t)stream), "gather");
}

int mfc_c8_to_nchw(const void* src, long long src_bstride_bytes, float* dst, int B, int C, int H, int W, int dtype, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!src || !dst || B < 1 || C < 1 || H < 1 || W < 1 || !dtype_ok(dtype)) return fail(MFC_EINVAL, "c8_to_nchw: bad argument");
  MFC_LAUNCH(mfc::launch_c8_to_nchw(src, src_bstride_bytes, dst, B, C, H, W, dtype == MFC_BF16, (cudaStream_t)stream), "c8_to_nchw");
}

// ---- weights -----------------------------------------------------------------------------------
int mfc_weight_standardize(const float* w, float* out, int Cout, int fan_in, float eps, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!w || !out || Cout < 1 || fan_in < 1) return fail(MFC_EINVAL, "weight_standardize: bad argument");
  MFC_LAUNCH(mfc::launch_weight_standardize(w, out, Cout, fan_in, eps, (cudaStream_t)stream), "weight_standardize");
}

int mfc_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, const float* conv_bias, float eps,
                float* scale, float* shift, int C, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!gamma || !beta || !mean || !var || !scale || !shift || C < 1) return fail(MFC_EINVAL, "bn_fold: bad argument");
  MFC_LAUNCH(mfc::launch_bn_fold(gamma, beta, mean, var, conv_bias, eps, scale, shift, C, (cudaStream_t)stream), "bn_fold");
}

int mfc_conv2d_query(const MfcConvDesc* d, MfcConvInfo* info) {
  int rc = validate_desc(d);
  if (rc != MFC_OK) return rc;
  if (!info) return fail(MFC_EINVAL, "conv_query: null info");
  mfc::ConvTiling t;
  rc = get_tiling(d, &t);
  if (rc != MFC_OK) return rc;
  info->nb = t.NB;
  info->nblk = t.nblk;
  info->cin_chunks = t.cin_chunks;
  info->ksteps = t.ksteps;
  info->tile_h = t.TH;
  info->tile_w = t.TW;
  info->tiles_per_image = t.tiles_x * t.tiles_y;
  info->stats_per_image = t.grid;
  info->runs = t.R;
  info->kstages = t.kstages;
  info->nstages = t.nstages;
  info->grid = t.grid;
  info->smem_bytes = (int)t.smem_bytes;
  info->tmem_cols = (int)t.tmem_cols;
  info->packed_weight_bytes = (long long)t.nblk * t.ksteps * t.entries * 2 * t.nrows_b * 16;
  info->weight_layout = t.slide;
  info->flags = tiling_fast_epilogue(d, t) ? 1 : 0;
  return MFC_OK;
}

int mfc_conv2d_pack_weights(const MfcConvDesc* d, const float* w_oihw, int Cin_w, const int* chan_map, const float* scale, void* packed,
                            void* stream) {
  MFC_REQUIRE_ARCH();
  int rc = validate_desc(d);
  if (rc != MFC_OK) return rc;
  if (!w_oihw || !packed || Cin_w < 1) return fail(MFC_EINVAL, "conv_pack: bad argument");
  mfc::ConvTiling t;
  rc = get_tiling(d, &t);
  if (rc != MFC_OK) return rc;
  MFC_LAUNCH(mfc::launch_pack_weights(w_oihw, d->Cout, Cin_w, t.entries, chan_map, t.cin_chunks, t.ksteps, t.NB, t.nblk,
                                      t.pair ? d->kw : 0, d->kh * d->kw, t.slide ? d->kh : 0, d->kw, scale, packed,
                                      d->dtype == MFC_BF16, (cudaStream_t)stream),
             "conv_pack");
}

int mfc_conv2d_fwd(const MfcConvDesc* d, const MfcConvIO* io, void* stream) {
  MFC_REQUIRE_ARCH();
  int rc = validate_desc(d);
  if (rc != MFC_OK) return rc;
  mfc::ConvTiling t;
  rc = get_tiling(d, &t);
  if (rc != MFC_OK) return rc;
  return conv_fwd_tiled(d, io, t, stream);
}

/* Bytes of packed-weight scratch that cover every candidate mfc_conv2d_autotune would measure for `d` (0 on a bad
 * descriptor).  Does not plan the geometry (mfc_conv2d_query would freeze its plan). */
long long mfc_conv2d_autotune_scratch_bytes(const MfcConvDesc* d) {
  if (validate_desc(d) != MFC_OK) return 0;
  std::vector<mfc::ConvTiling> cands;
  mfc::conv_shortlist(*d, getenv("MFC_CONV_TUNE_WIDTH") ? atoi(getenv("MFC_CONV_TUNE_WIDTH")) : 3, cands);
  long long need = 0;
  for (const auto& t : cands) need = std::max(need, (long long)t.nblk * t.ksteps * t.entries * 2 * t.nrows_b * 16);
  return need;
}

/* Diagnostic: the candidates mfc_conv2d_autotune would measure for `d`, one per line
 * "TH TW slide CBc NB nstages kstages R nacc tiles_x grid", in the cost model's order.  Returns the bytes needed. */
long long mfc_conv2d_shortlist(const MfcConvDesc* d, int per_bucket, char* buf, long long cap) {
  if (validate_desc(d) != MFC_OK) return 0;
  std::vector<mfc::ConvTiling> cands;
  mfc::conv_shortlist(*d, per_bucket, cands);
  std::string out;
  char line[160];
  for (const auto& t : cands) {
    snprintf(line, sizeof(line), "%d %d %d %d %d %d %d %d %d %d %d\n", t.TH, t.TW, t.slide, t.CBc, t.NB, t.nstages, t.kstages, t.R, t.nacc,
             t.tiles_x, t.grid);
    out += line;
  }
  const long long need = (long long)out.size() + 1;
  if (buf && cap > 0) {
    const long long n = need <= cap ? need - 1 : cap - 1;
    memcpy(buf, out.data(), (size_t)n);
    buf[n] = 0;
  }
  return need;
}

/* Measures the planner's shortlisted tilings of `d` on the device with the caller's real buffers and keeps the fastest for
 * every later query / pack / fwd of the same geometry.  io->w_packed is ignored: the raw weights are packed into
 * `scratch_packed` (packed_weight_bytes of mfc_conv2d_query) once per weight layout.  io->stats, when given, must hold
 * [B][148][nb*nblk][2] floats (any candidate's grid fits).  Synchronises the stream.  Idempotent per geometry. */
int mfc_conv2d_autotune(const MfcConvDesc* d, const MfcConvIO* io, const float* w_oihw, int Cin_w, const int* chan_map,
                        void* scratch_packed, long long scratch_bytes, int reps, void* stream) {
  MFC_REQUIRE_ARCH();
  int rc = validate_desc(d);
  if (rc != MFC_OK) return rc;
  if (!io || !w_oihw || !scratch_packed || reps < 1) return fail(MFC_EINVAL, "conv_autotune: bad argument");
  const PlanKey key = plan_key(d);
  {
    // a plan that was already handed out (or measured, or taken from the table) is final: weights may have been packed
    // and statistics buffers sized for it
    std::lock_guard<std::mutex> g(g_plan_mu);
    auto it = g_plans.find(key);
    if (it != g_plans.end() && (it->second.handed_out || it->second.tuned)) return MFC_OK;
    if (g_table.count(key)) return MFC_OK;
  }
  std::vector<mfc::ConvTiling> cands;
  mfc::conv_shortlist(*d, getenv("MFC_CONV_TUNE_WIDTH") ? atoi(getenv("MFC_CONV_TUNE_WIDTH")) : 3, cands);
  if (cands.empty()) return fail(MFC_EINVAL, "conv: no tiling fits shared memory / TMEM for this shape");
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return fail(MFC_ECUDA, "conv_autotune: cudaEventCreate failed");
  MfcConvIO tio = *io;
  tio.w_packed = scratch_packed;
  float best_ms = 1e30f;
  int best_i = -1;
  // candidates are measured grouped by weight image (layout, N-block width): one packing per group
  std::vector<char> done(cands.size(), 0);
  for (size_t g0 = 0; g0 < cands.size() && rc == MFC_OK; ++g0) {
    if (done[g0]) continue;
    bool packed = false;
    for (size_t i = g0; i < cands.size() && rc == MFC_OK; ++i) {
      const mfc::ConvTiling& t = cands[i];
      if (done[i] || t.slide != cands[g0].slide || t.NB != cands[g0].NB) continue;
      done[i] = 1;
      if (!tiling_serves(d, t)) continue;
      if ((long long)t.nblk * t.ksteps * t.entries * 2 * t.nrows_b * 16 > scratch_bytes) continue;
      if (!packed) {
        cudaError_t e = mfc::launch_pack_weights(w_oihw, d->Cout, Cin_w, t.entries, chan_map, t.cin_chunks, t.ksteps, t.NB, t.nblk,
                                                 t.pair ? d->kw : 0, d->kh * d->kw, t.slide ? d->kh : 0, d->kw, nullptr, scratch_packed,
                                                 d->dtype == MFC_BF16, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the conv reads its weights before griddepcontrol.wait
        if (e != cudaSuccess) rc = cuda_fail(e, "conv_autotune: pack");
        packed = true;
      }
      if (rc == MFC_OK) rc = conv_fwd_tiled(d, &tio, t, stream);  // warm-up (instruction cache, L2 state)
      if (rc == MFC_EINVAL) {  // this candidate does not support the requested epilogue (e.g. statistics with a padded Cout > 256)
        rc = MFC_OK;
        continue;
      }
      cudaEventRecord(e0, st);
      for (int r = 0; r < reps && rc == MFC_OK; ++r) rc = conv_fwd_tiled(d, &tio, t, stream);
      cudaEventRecord(e1, st);
      if (rc != MFC_OK) break;
      cudaError_t e = cudaEventSynchronize(e1);
      if (e != cudaSuccess) {
        rc = cuda_fail(e, "conv_autotune: sync");
        break;
      }
      float ms = 0.0f;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best_ms) {
        best_ms = ms;
        best_i = (int)i;
      }
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc != MFC_OK) return rc;
  if (best_i < 0) return fail(MFC_EINVAL, "conv_autotune: no candidate ran");
  {
    std::lock_guard<std::mutex> g(g_plan_mu);
    auto it = g_plans.find(key);
    if (it == g_plans.end() || !it->second.handed_out) {  // another thread may have planned this geometry meanwhile
      PlanEntry e;
      e.t = cands[best_i];
      e.tuned = true;
      g_plans[key] = e;
    }
  }
  return MFC_OK;
}

/* Tuning table.  One line per geometry:
 *   B Hin Win Hout Wout Cout kh kw stride pad upsample chunks affine flags dtype out_stride : TH TW slide CBc NB nstages
 * export: every plan of this process that was measured (mfc_conv2d_autotune) or came from an imported table.  Returns the
 * number of bytes the text needs (incl. the terminating 0); writes at most `cap` bytes. */
long long mfc_conv2d_plan_export(char* buf, long long cap) {
  std::lock_guard<std::mutex> g(g_plan_mu);
  std::string out;
  char line[256];
  for (const auto& kv : g_plans) {
    if (!kv.second.tuned) continue;
    const PlanKey& k = kv.first;
    const mfc::ConvTiling& t = kv.second.t;
    snprintf(line, sizeof(line), "%d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d : %d %d %d %d %d %d\n", std::get<0>(k), std::get<1>(k),
             std::get<2>(k), std::get<3>(k), std::get<4>(k), std::get<5>(k), std::get<6>(k), std::get<7>(k), std::get<8>(k),
             std::get<9>(k), std::get<10>(k), std::get<11>(k), std::get<12>(k), std::get<13>(k), std::get<14>(k), std::get<15>(k),
             t.TH, t.TW, t.slide, t.CBc, t.NB, t.nstages);
    out += line;
  }
  for (const auto& kv : g_table) {  // imported entries that were not needed by this process are passed through
    if (g_plans.count(kv.first)) continue;
    const PlanKey& k = kv.first;
    const TableChoice& c = kv.second;
    snprintf(line, sizeof(line), "%d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d : %d %d %d %d %d %d\n", std::get<0>(k), std::get<1>(k),
             std::get<2>(k), std::get<3>(k), std::get<4>(k), std::get<5>(k), std::get<6>(k), std::get<7>(k), std::get<8>(k),
             std::get<9>(k), std::get<10>(k), std::get<11>(k), std::get<12>(k), std::get<13>(k), std::get<14>(k), std::get<15>(k),
             c.TH, c.TW, c.slide, c.CBc, c.NB, c.nstages);
    out += line;
  }
  const long long need = (long long)out.size() + 1;
  if (buf && cap > 0) {
    const long long n = need <= cap ? need - 1 : cap - 1;
    memcpy(buf, out.data(), (size_t)n);
    buf[n] = 0;
  }
  return need;
}

/* import: lines of the format above ('#' starts a comment).  Geometries already planned in this process keep their plan.
 * Returns the number of entries taken, or a negative error code for a malformed line. */
int mfc_conv2d_plan_import(const char* text) {
  if (!text) return fail(MFC_EINVAL, "plan_import: null text");
  std::lock_guard<std::mutex> g(g_plan_mu);
  int taken = 0, lineno = 0;
  const char* p = text;
  while (*p) {
    const char* e = strchr(p, '\n');
    const size_t len = e ? (size_t)(e - p) : strlen(p);
    ++lineno;
    if (len > 0 && len < 240 && p[0] != '#') {
      char line[256];
      memcpy(line, p, len);
      line[len] = 0;
      int k[16];
      TableChoice c;
      const int n = sscanf(line, "%d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d : %d %d %d %d %d %d", &k[0], &k[1], &k[2], &k[3], &k[4],
                           &k[5], &k[6], &k[7], &k[8], &k[9], &k[10], &k[11], &k[12], &k[13], &k[14], &k[15], &c.TH, &c.TW, &c.slide,
                           &c.CBc, &c.NB, &c.nstages);
      bool blank = true;
      for (size_t i = 0; i < len; ++i) blank = blank && (line[i] == ' ' || line[i] == '\t' || line[i] == '\r');
      if (!blank) {
        if (n != 22) return fail(MFC_EINVAL, "plan_import: malformed line %d", lineno);
        const PlanKey key{k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7], k[8], k[9], k[10], k[11], k[12], k[13], k[14], k[15]};
        if (!g_plans.count(key)) {
          g_table[key] = c;
          ++taken;
        }
      }
    }
    if (!e) break;
    p = e + 1;
  }
  return taken;
}

int mfc_gn_finalize(const float* stats, int B, int stats_per_image, int cpad, int C, int groups, long long pixels, const float* gamma,
                    const float* beta, float eps, float* affine, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!stats || !gamma || !beta || !affine || B < 1 || stats_per_image < 1 || C < 1 || groups < 1 || C % groups || cpad < C || pixels < 1)
    return fail(MFC_EINVAL, "gn_finalize: bad argument");
  MFC_LAUNCH(mfc::launch_gn_finalize(stats, B, stats_per_image, cpad, C, groups, pixels, gamma, beta, eps, affine, (cudaStream_t)stream),
             "gn_finalize");
}

int mfc_affine_silu_add(const void* a, const float* affine, const void* r, void* out, int B, int chunks, long long pixels, int dtype,
                        int* overflow, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!a || !affine || !r || !out || B < 1 || chunks < 1 || pixels < 1 || !dtype_ok(dtype)) return fail(MFC_EINVAL, "affine_silu_add: bad argument");
  MFC_LAUNCH(mfc::launch_affine_silu_add(a, affine, r, out, B, chunks, pixels, dtype == MFC_BF16, overflow, (cudaStream_t)stream), "affine_silu_add");
}

// ---- fusion ------------------------------------------------------------------------------------
int mfc_flow_warp(const MfcWarpArgs* a, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!a || a->B < 1 || a->H < 2 || a->W < 2 || a->K < 2 || a->K > MFC_MAX_SRC || !a->grid || a->seg_chunks < 0 || !dtype_ok(a->dtype))
    return fail(MFC_EINVAL, "flow_warp: bad argument");
  if (a->H > a->grid_h || a->W > a->grid_w) return fail(MFC_EINVAL, "flow_warp: %dx%d exceeds the stored %dx%d grid", a->H, a->W, a->grid_h, a->grid_w);
  for (int f = 1; f < a->K; ++f) {
    if (!a->flow[f - 1]) return fail(MFC_EINVAL, "flow_warp: flow %d is null", f - 1);
    if (a->seg_chunks > 0 && (!a->seg[f] || !a->seg_out[f])) return fail(MFC_EINVAL, "flow_warp: seg/seg_out %d is null", f);
  }
  MFC_LAUNCH(mfc::launch_flow_warp(*a, (cudaStream_t)stream), "flow_warp");
}

int mfc_heatmap_head(const float* logits, int B, int N, long long pixels, float* logp, float* prob, uint8_t* argmax, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!logits || B < 1 || N < 1 || N > 255 || pixels < 1) return fail(MFC_EINVAL, "heatmap_head: bad argument");
  MFC_LAUNCH(mfc::launch_heatmap_head(logits, B, N, pixels, logp, prob, argmax, (cudaStream_t)stream), "heatmap_head");
}

int mfc_argmax_u8(const float* x, int B, int N, long long pixels, uint8_t* out, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!x || !out || B < 1 || N < 1 || N > 255 || pixels < 1) return fail(MFC_EINVAL, "argmax_u8: bad argument");
  MFC_LAUNCH(mfc::launch_argmax_u8(x, B, N, pixels, out, (cudaStream_t)stream), "argmax_u8");
}

int mfc_maxpool2(const void* src, long long src_bs, void* dst, long long dst_bs, int B, int chunks, int H, int W, int dtype, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!src || !dst || B < 1 || chunks < 1 || H < 2 || W < 2 || !dtype_ok(dtype)) return fail(MFC_EINVAL, "maxpool2: bad argument");
  if (((uintptr_t)src & 15) || ((uintptr_t)dst & 15) || (src_bs & 15) || (dst_bs & 15)) return fail(MFC_EINVAL, "maxpool2: not 16-byte aligned");
  MFC_LAUNCH(mfc::launch_maxpool2(src, src_bs, dst, dst_bs, B, chunks, H, W, dtype == MFC_BF16, (cudaStream_t)stream), "maxpool2");
}

// ---- HRNet resampling ---------------------------------------------------------------------------
int mfc_fuse_sum(const MfcFuseArgs* a, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!a || a->B < 1 || a->chunks < 1 || a->H < 1 || a->W < 1 || a->nterms < 1 || a->nterms > MFC_MAX_SRC || !a->out || !dtype_ok(a->dtype))
    return fail(MFC_EINVAL, "fuse_sum: bad argument");
  if ((a->scale == nullptr) != (a->shift == nullptr)) return fail(MFC_EINVAL, "fuse_sum: scale and shift come together");
  for (int j = 0; j < a->nterms; ++j) {
    const MfcFuseTerm& t = a->term[j];
    if (!t.ptr || t.H < 1 || t.W < 1 || t.H > a->H || t.W > a->W) return fail(MFC_EINVAL, "fuse_sum: term %d invalid (only upsampling)", j);
    if (((uintptr_t)t.ptr & 15) || (t.batch_stride & 15)) return fail(MFC_EINVAL, "fuse_sum: term %d not 16-byte aligned", j);
  }
  MFC_LAUNCH(mfc::launch_fuse_sum(*a, (cudaStream_t)stream), "fuse_sum");
}

int mfc_bilinear_resize(const MfcResizeArgs* a, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!a || !a->src || (!a->dst_nchw && !a->dst_c8) || a->B < 1 || a->C < 1 || a->Hin < 1 || a->Win < 1 || a->Hout < 1 || a->Wout < 1 ||
      !dtype_ok(a->dtype))
    return fail(MFC_EINVAL, "bilinear_resize: bad argument");
  MFC_LAUNCH(mfc::launch_bilinear_resize(a->src, a->B, a->C, a->Hin, a->Win, a->Hout, a->Wout, a->dst_nchw, a->dst_c8,
                                         a->c8_batch_stride, a->dtype == MFC_BF16, (cudaStream_t)stream),
             "bilinear_resize");
}

// ---- loss ----------------------------------------------------------------------------------------
long long mfc_segmentation_loss_workspace(int B, int N, long long pixels) {
  if (B < 1 || N < 2 || N > 16 || pixels < 1) return 0;
  return (long long)mfc::loss_blocks(B, pixels) * (2 + 3 * (N - 1)) * (long long)sizeof(double);
}

int mfc_segmentation_loss(const float* logits, const long long* target, const float* class_weights, int B, int N, long long pixels,
                          float w_nll, float w_jaccard, void* workspace, float* out, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!logits || !target || !workspace || !out || B < 1 || N < 2 || N > 16 || pixels < 1) return fail(MFC_EINVAL, "segmentation_loss: bad argument");
  MFC_LAUNCH(mfc::launch_segmentation_loss(logits, target, class_weights, B, N, pixels, w_nll, w_jaccard, (double*)workspace, out,
                                           (cudaStream_t)stream),
             "segmentation_loss");
}

/* The additive statistics of the loss over this rank's shard: sums[2 + 3(N-1)] doubles = (sum w[t](-logp[t]), sum w[t],
 * then per class c>=1: I_c, S_c, T_c).  Data-parallel ranks add their records (one tiny all-reduce) and every rank then
 * evaluates the loss of the GLOBAL batch -- what nn.DataParallel's gather-to-GPU0 computes (src/engine.py:56-66). */
int mfc_segmentation_loss_sums(const float* logits, const long long* target, const float* class_weights, int B, int N, long long pixels,
                               void* workspace, double* sums, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!logits || !target || !workspace || !sums || B < 1 || N < 2 || N > 16 || pixels < 1)
    return fail(MFC_EINVAL, "segmentation_loss_sums: bad argument");
  MFC_LAUNCH(mfc::launch_segmentation_loss_sums(logits, target, class_weights, B, N, pixels, (double*)workspace, sums, (cudaStream_t)stream),
             "segmentation_loss_sums");
}

int mfc_segmentation_loss_from_sums(const double* sums, int N, float w_nll, float w_jaccard, float* out, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!sums || !out || N < 2 || N > 16) return fail(MFC_EINVAL, "segmentation_loss_from_sums: bad argument");
  MFC_LAUNCH(mfc::launch_segmentation_loss_from_sums(sums, N, w_nll, w_jaccard, out, (cudaStream_t)stream), "segmentation_loss_from_sums");
}

/* Backward: dlogits = scale * d total / d logits of this rank's shard, where `sums` holds n_records records of loss
 * statistics that are added first (1 = the reduced record of mfc_segmentation_loss_sums, possibly all-reduced over ranks;
 * 0 = `sums` is the workspace mfc_segmentation_loss just filled for the same logits).  `coef` is 64 floats of scratch. */
int mfc_segmentation_loss_bwd(const float* logits, const long long* target, const float* class_weights, int B, int N, long long pixels,
                              float w_nll, float w_jaccard, float scale, const void* sums, int n_records, float* coef, float* dlogits,
                              void* stream) {
  MFC_REQUIRE_ARCH();
  if (!logits || !target || !sums || !coef || !dlogits || B < 1 || N < 2 || N > 16 || pixels < 1 || n_records < 0)
    return fail(MFC_EINVAL, "segmentation_loss_bwd: bad argument");
  MFC_LAUNCH(mfc::launch_segmentation_loss_bwd(logits, target, class_weights, B, N, pixels, w_nll, w_jaccard, scale, (const double*)sums,
                                               n_records, coef, dlogits, (cudaStream_t)stream),
             "segmentation_loss_bwd");
}

/* One torch.optim.Adam step (amsgrad off) on a flat fp32 bucket; `step` is the 1-based step count, grad_scale multiplies the
 * gradient first (the 1/world_size of the data-parallel average). */
int mfc_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, float grad_scale, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!param || !grad || !exp_avg || !exp_avg_sq || n < 1 || step < 1) return fail(MFC_EINVAL, "adam_step: bad argument");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  MFC_LAUNCH(mfc::launch_adam(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2),
                              grad_scale, (cudaStream_t)stream),
             "adam_step");
}

// ---- frame ingest --------------------------------------------------------------------------------
int mfc_ingest_rgb(const uint8_t* bgr, long long frame_stride_bytes, float* out, int B, int H, int W, const float* mean3_host,
                   const float* std3_host, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!bgr || !out || !mean3_host || !std3_host || B < 1 || H < 1 || W < 1 || frame_stride_bytes < (long long)H * W * 3)
    return fail(MFC_EINVAL, "ingest_rgb: bad argument");
  MFC_LAUNCH(mfc::launch_ingest_rgb(bgr, frame_stride_bytes, out, B, (long long)H * W, mean3_host, std3_host, (cudaStream_t)stream), "ingest_rgb");
}

int mfc_ingest_depth(const uint8_t* bgr, long long frame_stride_bytes, float* out, int B, int H, int W, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!bgr || !out || B < 1 || H < 1 || W < 1 || frame_stride_bytes < (long long)H * W * 3) return fail(MFC_EINVAL, "ingest_depth: bad argument");
  MFC_LAUNCH(mfc::launch_ingest_depth(bgr, frame_stride_bytes, out, B, (long long)H * W, (cudaStream_t)stream), "ingest_depth");
}

// ---- correlation -------------------------------------------------------------------------------
int mfc_correlation_fwd(const float* first, const float* second, float* out, int B, int C, int H, int W, int max_disp, int stride2,
                        int exact_order, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!first || !second || !out || B < 1 || C < 1 || H < 1 || W < 1) return fail(MFC_EINVAL, "correlation: bad argument");
  if (stride2 != 1 && stride2 != 2) return fail(MFC_EINVAL, "correlation: stride2 %d unsupported (1 or 2)", stride2);
  if (max_disp < 0 || max_disp > 32 || max_disp % stride2) return fail(MFC_EINVAL, "correlation: max_disp %d unsupported", max_disp);
  if (!exact_order && mfc::correlation_tma_supported(C, H, W, max_disp, stride2) && !(((uintptr_t)first | (uintptr_t)second | (uintptr_t)out) & 15) &&
      encode_tiled_fn()) {
    // both operands as (W, H, C, B) fp32 tensors; boxes = the CTA tile (+ halo for `second`), zero fill outside the image
    CUtensorMap m1, m2;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
    cuuint32_t b1[4], b2[4], es[4];
    mfc::correlation_tma_boxes(H, W, stride2, b1, b2, es);
    EncodeTiledFn enc = encode_tiled_fn();
    CUresult r1 = enc(&m1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(first), dims, strides, b1, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(second), dims, strides, b2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS)
      MFC_LAUNCH(mfc::launch_correlation_tma(m1, m2, out, B, C, H, W, stride2, (cudaStream_t)stream), "correlation (tma)");
  }
  MFC_LAUNCH(mfc::launch_correlation(first, second, out, B, C, H, W, max_disp, stride2, exact_order, (cudaStream_t)stream), "correlation");
}

int mfc_correlation_bwd(const float* first, const float* second, const float* grad_out, float* grad_first, float* grad_second, int B,
                        int C, int H, int W, int max_disp, int stride2, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!first || !second || !grad_out || (!grad_first && !grad_second) || B < 1 || C < 1 || H < 1 || W < 1)
    return fail(MFC_EINVAL, "correlation_bwd: bad argument");
  if (stride2 != 1 && stride2 != 2) return fail(MFC_EINVAL, "correlation_bwd: stride2 %d unsupported (1 or 2)", stride2);
  if (max_disp < 0 || max_disp > 32 || max_disp % stride2) return fail(MFC_EINVAL, "correlation_bwd: max_disp %d unsupported", max_disp);
  MFC_LAUNCH(mfc::launch_correlation_bwd(first, second, grad_out, grad_first, grad_second, B, C, H, W, max_disp, stride2, (cudaStream_t)stream),
             "correlation_bwd");
}

// ---- UnFlow network pieces -----------------------------------------------------------------------
int mfc_unflow_preprocess(const float* rgb, float* out, int B, int H, int W, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!rgb || !out || B < 1 || H < 1 || W < 1) return fail(MFC_EINVAL, "unflow_preprocess: bad argument");
  MFC_LAUNCH(mfc::launch_unflow_prep(rgb, out, B, (long long)H * W, (cudaStream_t)stream), "unflow_preprocess");
}

int mfc_nchw_to_c8(const float* src, void* dst, long long dst_bstride_bytes, int B, int C, int H, int W, int dtype, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!src || !dst || B < 1 || C < 1 || H < 1 || W < 1 || !dtype_ok(dtype) || ((uintptr_t)dst & 15) || (dst_bstride_bytes & 15))
    return fail(MFC_EINVAL, "nchw_to_c8: bad argument");
  MFC_LAUNCH(mfc::launch_nchw_to_c8(src, dst, dst_bstride_bytes, B, C, (long long)H * W, dtype == MFC_BF16, (cudaStream_t)stream), "nchw_to_c8");
}

int mfc_unflow_warp(const float* second, const float* flow, const float* first, float* warped, float* absdiff, int B, int C, int H, int W,
                    void* stream) {
  MFC_REQUIRE_ARCH();
  if (!second || !flow || !warped || (absdiff && !first) || B < 1 || C < 1 || H < 2 || W < 2) return fail(MFC_EINVAL, "unflow_warp: bad argument");
  MFC_LAUNCH(mfc::launch_unflow_warp(second, flow, first, warped, absdiff, B, C, H, W, (cudaStream_t)stream), "unflow_warp");
}

int mfc_unflow_upscale(const float* x, const float* w, float* out, int B, int h, int w_in, float scale, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!x || !w || !out || B < 1 || h < 1 || w_in < 1) return fail(MFC_EINVAL, "unflow_upscale: bad argument");
  MFC_LAUNCH(mfc::launch_unflow_upscale(x, w, out, B, h, w_in, scale, (cudaStream_t)stream), "unflow_upscale");
}

int mfc_resize_u8(const uint8_t* src, long long frame_stride_bytes, int h, int w, int C, uint8_t* dst, int B, int H, int W, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!src || !dst || B < 1 || h < 1 || w < 1 || H < 1 || W < 1 || (C != 1 && C != 3) || frame_stride_bytes < (long long)h * w * C)
    return fail(MFC_EINVAL, "resize_u8: bad argument");
  MFC_LAUNCH(mfc::launch_resize_u8(src, frame_stride_bytes, h, w, C, dst, B, H, W, (cudaStream_t)stream), "resize_u8");
}
int mfc_bgr2gray_u8(const uint8_t* bgr, long long frame_stride_bytes, uint8_t* gray, int B, int H, int W, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!bgr || !gray || B < 1 || H < 1 || W < 1 || frame_stride_bytes < (long long)H * W * 3) return fail(MFC_EINVAL, "bgr2gray_u8: bad argument");
  MFC_LAUNCH(mfc::launch_bgr2gray_u8(bgr, frame_stride_bytes, gray, B, (long long)H * W, (cudaStream_t)stream), "bgr2gray_u8");
}
int mfc_ingest_gray(const uint8_t* gray, float* out, long long n, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!gray || !out || n < 1) return fail(MFC_EINVAL, "ingest_gray: bad argument");
  MFC_LAUNCH(mfc::launch_ingest_gray(gray, out, n, (cudaStream_t)stream), "ingest_gray");
}

// ---- RAFT pieces ---------------------------------------------------------------------------------
int mfc_pointwise(const MfcPointwiseArgs* a, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!a || !a->a || !a->out || a->B < 1 || a->chunks < 1 || a->pixels < 1 || !dtype_ok(a->dtype) || a->kind < 0 || a->kind > 3)
    return fail(MFC_EINVAL, "pointwise: bad argument");
  if ((a->kind == MFC_PW_CTX_SPLIT && !a->out2) || ((a->kind == MFC_PW_GRU_RH || a->kind == MFC_PW_GRU_UPDATE) && !a->r))
    return fail(MFC_EINVAL, "pointwise: kind %d misses an operand", a->kind);
  MFC_LAUNCH(mfc::launch_pointwise(a->kind, a->a, a->a_aff, a->r, a->r_aff, a->out, a->out2, a->B, a->chunks, a->pixels, a->relu_a,
                                   a->relu_out, a->dtype == MFC_BF16, (cudaStream_t)stream), "pointwise");
}

int mfc_raft_op(const MfcRaftArgs* a, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!a || a->B < 1 || a->h < 1 || a->w < 1) return fail(MFC_EINVAL, "raft_op: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  switch (a->kind) {
    case MFC_RAFT_CORR_VOLUME:
      if (!a->p0 || !a->p1 || !a->p2 || a->C < 1) return fail(MFC_EINVAL, "raft_op: corr_volume needs p0, p1, p2, C");
      MFC_LAUNCH(mfc::launch_raft_corr_volume((const float*)a->p0, (const float*)a->p1, (float*)a->p2, a->B, a->C, a->h * a->w, a->scale, st),
                 "raft corr_volume");
    case MFC_RAFT_POOL:
      if (!a->p0 || !a->p1 || a->h < 2 || a->w < 2) return fail(MFC_EINVAL, "raft_op: pool needs p0, p1 and h, w >= 2");
      MFC_LAUNCH(mfc::launch_raft_corr_pool((const float*)a->p0, (float*)a->p1, (long long)a->B, a->h, a->w, st), "raft pool");
    case MFC_RAFT_LOOKUP: {
      if (!a->p0 || !a->p4 || !a->p5 || a->levels < 1 || a->levels > 4 || a->radius < 0 || a->radius > 8 || !dtype_ok(a->dtype))
        return fail(MFC_EINVAL, "raft_op: lookup needs 1..4 levels, flow, out");
      if ((a->h >> (a->levels - 1)) < 2 || (a->w >> (a->levels - 1)) < 2) return fail(MFC_EINVAL, "raft_op: feature map too small for %d levels", a->levels);
      const float* lvl[4] = {(const float*)a->p0, (const float*)a->p1, (const float*)a->p2, (const float*)a->p3};
      for (int i = 0; i < a->levels; ++i)
        if (!lvl[i]) return fail(MFC_EINVAL, "raft_op: lookup level %d missing", i);
      const int side = 2 * a->radius + 1, chunks = (a->levels * side * side + 7) / 8;
      MFC_LAUNCH(mfc::launch_raft_lookup(lvl, (const float*)a->p4, a->p5, a->B, a->h, a->w, a->levels, a->radius, chunks, a->dtype == MFC_BF16, st),
                 "raft lookup");
    }
    case MFC_RAFT_FLOW_ADD:
      if (!a->p0 || !a->p1) return fail(MFC_EINVAL, "raft_op: flow_add needs p0, p1");
      MFC_LAUNCH(mfc::launch_raft_flow_add((float*)a->p0, (const float*)a->p1, (long long)a->B * 2 * a->h * a->w, st), "raft flow_add");
    case MFC_RAFT_UPSAMPLE:
      if (!a->p0 || !a->p1 || !a->p2) return fail(MFC_EINVAL, "raft_op: upsample needs p0, p1, p2");
      MFC_LAUNCH(mfc::launch_raft_upsample((const float*)a->p0, (const float*)a->p1, (float*)a->p2, a->B, a->h, a->w, a->scale, st), "raft upsample");
    case MFC_RAFT_RESIZE_AC:
      if (!a->p0 || !a->p2 || a->C < 1 || a->levels < 1 || a->radius < 1) return fail(MFC_EINVAL, "raft_op: resize needs p0, p2, C, Hout, Wout");
      MFC_LAUNCH(mfc::launch_raft_resize_ac((const float*)a->p0, (float*)a->p2, a->B * a->C, a->h, a->w, a->levels, a->radius, a->scale, st),
                 "raft resize");
    default:
      return fail(MFC_EINVAL, "raft_op: kind %d unknown", a->kind);
  }
}

// ---- key points --------------------------------------------------------------------------------
int mfc_gaussian_blur(const float* heat, float* tmp, float* out, int B, int H, int W, const double* weights_dev, int radius, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!heat || !tmp || !out || !weights_dev || B < 1 || H < 1 || W < 1 || radius < 0 || radius > 64) return fail(MFC_EINVAL, "gaussian_blur: bad argument");
  MFC_LAUNCH(mfc::launch_gaussian_blur(heat, tmp, out, B, H, W, weights_dev, radius, (cudaStream_t)stream), "gaussian_blur");
}

int mfc_localmax_mask(const float* sm, const uint8_t* cls, int cls_id, const uint8_t* footprint, int fh, int fw, uint8_t* mask, int B,
                      int H, int W, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!sm || !cls || !footprint || !mask || fh < 1 || fw < 1 || fh > 31 || fw > 31 || B < 1 || H < 1 || W < 1)
    return fail(MFC_EINVAL, "localmax_mask: bad argument");
  MFC_LAUNCH(mfc::launch_localmax_mask(sm, cls, cls_id, footprint, fh, fw, mask, B, H, W, (cudaStream_t)stream), "localmax_mask");
}

int mfc_class_mask(const uint8_t* cls, int cls_id, uint8_t* mask, long long n, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!cls || !mask || n < 1) return fail(MFC_EINVAL, "class_mask: bad argument");
  MFC_LAUNCH(mfc::launch_class_mask(cls, cls_id, mask, n, (cudaStream_t)stream), "class_mask");
}

int mfc_trace_contours(const uint8_t* mask, int H, int W, int* labels, double* out, int max_contours, int* n_out, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!mask || !labels || !out || !n_out || H < 1 || W < 1 || max_contours < 1) return fail(MFC_EINVAL, "trace_contours: bad argument");
  MFC_LAUNCH(mfc::launch_trace_contours(mask, H, W, labels, out, max_contours, n_out, (cudaStream_t)stream), "trace_contours");
}

int mfc_threshold_classes(const float* prob, int B, int N, long long pixels, float thr, uint8_t* out, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!prob || !out || B < 1 || N < 2 || N > 255 || pixels < 1) return fail(MFC_EINVAL, "threshold_classes: bad argument");
  MFC_LAUNCH(mfc::launch_threshold_classes(prob, B, N, pixels, thr, out, (cudaStream_t)stream), "threshold_classes");
}

int mfc_mask_heat(const float* heat, const uint8_t* cls, int cls_id, float* out, long long n, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!heat || !cls || !out || n < 1) return fail(MFC_EINVAL, "mask_heat: bad argument");
  MFC_LAUNCH(mfc::launch_mask_heat(heat, cls, cls_id, out, n, (cudaStream_t)stream), "mask_heat");
}

int mfc_refine_tip_mask(const uint8_t* mask, int H, int W, const int* labels, const double* rec, int max_contours, const int* n_contours,
                        double area_threshold, int* sel, uint8_t* out, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!mask || !labels || !rec || !n_contours || !sel || !out || H < 1 || W < 1 || max_contours < 1)
    return fail(MFC_EINVAL, "refine_tip_mask: bad argument");
  MFC_LAUNCH(mfc::launch_refine_tip_mask(mask, H, W, labels, rec, max_contours, n_contours, area_threshold, sel, out, (cudaStream_t)stream),
             "refine_tip_mask");
}

int mfc_top_contours(const double* rec, const int* n_contours, int max_contours, int W, double* top, void* stream) {
  MFC_REQUIRE_ARCH();
  if (!rec || !n_contours || !top || max_contours < 1 || W < 1) return fail(MFC_EINVAL, "top_contours: bad argument");
  MFC_LAUNCH(mfc::launch_top_contours(rec, n_contours, max_contours, W, top, (cudaStream_t)stream), "top_contours");
}

// ---- command list ------------------------------------------------------------------------------
namespace {
// side streams / events of the calling thread, per device (created on first use, never destroyed)
struct LanePool {
  bool ready = false;
  cudaStream_t s[MFC_MAX_LANES - 1];
  cudaEvent_t ev_main, ev[MFC_MAX_LANES - 1];
};
LanePool* lane_pool() {
  thread_local LanePool pools[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  LanePool& p = pools[dev];
  if (!p.ready) {
    for (int i = 0; i < MFC_MAX_LANES - 1; ++i) {
      if (cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&p.ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    if (cudaEventCreateWithFlags(&p.ev_main, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    p.ready = true;
  }
  return &p;
}
int run_list_impl(const MfcCmd* cmds, int n, void* stream, bool use_lanes);
}  // namespace

int mfc_run_list(const MfcCmd* cmds, int n, void* stream) { return run_list_impl(cmds, n, stream, true); }

namespace {
int run_list_impl(const MfcCmd* cmds, int n, void* main_stream, bool use_lanes) {
  if (!cmds || n < 0) return fail(MFC_EINVAL, "run_list: bad argument");
  LanePool* pool = nullptr;
  for (int i = 0; i < n; ++i) {
    const MfcCmd& c = cmds[i];
    int rc = MFC_OK;
    void* stream = main_stream;
    if (use_lanes && (c.lane != 0 || c.op == MFC_OP_FORK || c.op == MFC_OP_JOIN)) {
      if (c.lane < 0 || c.lane >= MFC_MAX_LANES) return fail(MFC_EINVAL, "run_list: lane %d out of range at %d", c.lane, i);
      if (!pool) pool = lane_pool();
      if (!pool) return fail(MFC_ECUDA, "run_list: cannot create the side streams");
      if (c.lane > 0) stream = (void*)pool->s[c.lane - 1];
    }
    switch (c.op) {
      case MFC_OP_FORK:
        if (use_lanes) {
          cudaError_t e = cudaEventRecord(pool->ev_main, (cudaStream_t)main_stream);
          for (int l = 0; l < MFC_MAX_LANES - 1 && e == cudaSuccess; ++l) e = cudaStreamWaitEvent(pool->s[l], pool->ev_main, 0);
          if (e != cudaSuccess) rc = cuda_fail(e, "run_list: fork");
        }
        break;
      case MFC_OP_JOIN:
        if (use_lanes) {
          cudaError_t e = cudaSuccess;
          for (int l = 0; l < MFC_MAX_LANES - 1 && e == cudaSuccess; ++l) {
            e = cudaEventRecord(pool->ev[l], pool->s[l]);
            if (e == cudaSuccess) e = cudaStreamWaitEvent((cudaStream_t)main_stream, pool->ev[l], 0);
          }
          if (e != cudaSuccess) rc = cuda_fail(e, "run_list: join");
        }
        break;
      case MFC_OP_CONV:
        rc = mfc_conv2d_fwd((const MfcConvDesc*)c.a, (const MfcConvIO*)c.b, stream);
        break;
      case MFC_OP_GN_FINALIZE: {
        const MfcGnArgs* g = (const MfcGnArgs*)c.a;
        rc = mfc_gn_finalize(g->stats, g->B, g->stats_per_image, g->cpad, g->C, g->groups, g->pixels, g->gamma, g->beta, g->eps, g->affine, stream);
        break;
      }
      case MFC_OP_AFFINE_SILU_ADD: {
        const MfcAddArgs* g = (const MfcAddArgs*)c.a;
        rc = mfc_affine_silu_add(g->a, g->affine, g->r, g->out, g->B, g->chunks, g->pixels, g->dtype, g->overflow, stream);
        break;
      }
      case MFC_OP_GATHER: {
        const MfcGatherArgs* g = (const MfcGatherArgs*)c.a;
        rc = mfc_gather_nchw_to_c8(&g->g, g->dst, g->dst_bstride_bytes, g->B, g->H, g->W, g->dtype, stream);
        break;
      }
      case MFC_OP_WARP:
        rc = mfc_flow_warp((const MfcWarpArgs*)c.a, stream);
        break;
      case MFC_OP_POINTWISE:
        rc = mfc_pointwise((const MfcPointwiseArgs*)c.a, stream);
        break;
      case MFC_OP_RAFT:
        rc = mfc_raft_op((const MfcRaftArgs*)c.a, stream);
        break;
      case MFC_OP_FUSE_SUM:
        rc = mfc_fuse_sum((const MfcFuseArgs*)c.a, stream);
        break;
      case MFC_OP_RESIZE:
        rc = mfc_bilinear_resize((const MfcResizeArgs*)c.a, stream);
        break;
      case MFC_OP_MAXPOOL2: {
        const MfcPoolArgs* g = (const MfcPoolArgs*)c.a;
        rc = mfc_maxpool2(g->src, g->src_bstride_bytes, g->dst, g->dst_bstride_bytes, g->B, g->chunks, g->H, g->W, g->dtype, stream);
        break;
      }
      case MFC_OP_HEATMAP: {
        const MfcHeatmapArgs* g = (const MfcHeatmapArgs*)c.a;
        rc = mfc_heatmap_head(g->logits, g->B, g->N, g->pixels, g->logp, g->prob, g->argmax, stream);
        break;
      }
      default:
        rc = fail(MFC_EINVAL, "run_list: unknown op %d at %d", c.op, i);
    }
    if (rc != MFC_OK) {
      char tmp[400];
      strncpy(tmp, g_err, sizeof(tmp) - 1);
      tmp[sizeof(tmp) - 1] = 0;
      return fail(rc, "run_list[%d] op %d: %s", i, c.op, tmp);
    }
  }
  return MFC_OK;
}
}  // namespace

// ---- CUDA graphs ---------------------------------------------------------------------------------
// A command list whose pointers are all static (the streaming runner's per-slot programs) is captured once and replayed
// with ONE cudaGraphLaunch per frame: the ~350 launches of an HRNet frame are latency-bound when issued one by one from
// the host.  Capture happens on a private stream (the caller's stream may be the legacy default stream, which cannot be
// captured); programmatic-dependent-launch attributes become programmatic edges of the graph.
int mfc_graph_capture(const MfcCmd* cmds, int n, void** graph_out) {
  MFC_REQUIRE_ARCH();
  if (!cmds || n < 1 || !graph_out) return fail(MFC_EINVAL, "graph_capture: bad argument");
  *graph_out = nullptr;
  cudaStream_t cs;
  cudaError_t e = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
  if (e != cudaSuccess) return cuda_fail(e, "graph_capture: cudaStreamCreate");
  e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) {
    cudaStreamDestroy(cs);
    return cuda_fail(e, "graph_capture: cudaStreamBeginCapture");
  }
  const int rc = mfc_run_list(cmds, n, cs);
  cudaGraph_t g = nullptr;
  e = cudaStreamEndCapture(cs, &g);
  cudaStreamDestroy(cs);
  if (rc != MFC_OK) {
    if (g) cudaGraphDestroy(g);
    return rc;
  }
  if (e != cudaSuccess || !g) return cuda_fail(e, "graph_capture: cudaStreamEndCapture");
  cudaGraphExec_t ex = nullptr;
  e = cudaGraphInstantiate(&ex, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return cuda_fail(e, "graph_capture: cudaGraphInstantiate");
  *graph_out = (void*)ex;
  return MFC_OK;
}

int mfc_graph_launch(void* graph, void* stream) {
  if (!graph) return fail(MFC_EINVAL, "graph_launch: null graph");
  MFC_LAUNCH(cudaGraphLaunch((cudaGraphExec_t)graph, (cudaStream_t)stream), "graph_launch");
}

int mfc_graph_destroy(void* graph) {
  if (!graph) return MFC_OK;
  MFC_LAUNCH(cudaGraphExecDestroy((cudaGraphExec_t)graph), "graph_destroy");
}

int mfc_run_list_timed(const MfcCmd* cmds, int n, void* stream, float* ms_out) {
  if (!cmds || n < 1 || !ms_out) return fail(MFC_EINVAL, "run_list_timed: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<cudaEvent_t> ev((size_t)n + 1);
  for (auto& e : ev)
    if (cudaEventCreate(&e) != cudaSuccess) return fail(MFC_ECUDA, "run_list_timed: cudaEventCreate failed");
  int rc = MFC_OK;
  cudaEventRecord(ev[0], st);
  for (int i = 0; i < n && rc == MFC_OK; ++i) {
    rc = run_list_impl(cmds + i, 1, stream, false);  // serial on the caller's stream: per-command times
    cudaEventRecord(ev[i + 1], st);
  }
  cudaError_t e = cudaEventSynchronize(ev[n]);
  if (rc == MFC_OK && e != cudaSuccess) rc = cuda_fail(e, "run_list_timed: sync");
  if (rc == MFC_OK)
    for (int i = 0; i < n; ++i) cudaEventElapsedTime(&ms_out[i], ev[i], ev[i + 1]);
  for (auto& x : ev) cudaEventDestroy(x);
  return rc;
}

}  // extern "C"
