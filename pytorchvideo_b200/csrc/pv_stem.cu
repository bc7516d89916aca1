// Stem convolutions (3 input channels, stride 2 along W) on tcgen05 with ZERO-COPY im2col.
//
// The network input is stored NDHWC with 4 channels (f16) and the W padding written physically, so one input pixel
// is 8 bytes and an output pixel of a stride-2 convolution advances 16 bytes along the row.  A canonical NO-SWIZZLE
// K-major UMMA operand is made of core matrices of 8 rows x 16 bytes whose rows are 16 bytes apart; with
//     SBO (next 8-row group) = 128 B   and   LBO (next 16-byte K chunk) = 16 B
// the address of (row m, chunk j) is  start + 16 * (m + j):  row m of the A operand is the window of the RAW input row
// that starts at output pixel m - the im2col matrix of a (.., kw) filter row exists without being built
// (tools/probe/umma_window.cu verifies the addressing on B200).  So the A operand of a filter row (dt, dh) is just
// the input row (t + dt, 2 h + dh) copied once into shared memory by ONE bulk copy (cp.async.bulk, ~1.9 KB), instead
// of a 128-window tensor-map box per filter row (16 KB of overlapping 64-byte rows, one TMA request per window:
// the r01 "window mode", 330 us for the SlowFast Fast stem).
//
// One tile = one output row (up to 128 pixels along W) x all output channels; one pipeline stage = the kt * kh input
// rows of that tile; the packed weights stay resident in shared memory for the whole (persistent) CTA.  Warp roles,
// TMEM accumulator ring and the fused BN / activation epilogue are those of pv_igemm.cu.
#include "pv_common.cuh"
#include "pv_sm100.cuh"
#include "pv_epilogue.cuh"

#include <stdlib.h>
#include <string.h>

namespace pv {

using namespace sm100;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();   // pv_igemm.cu

constexpr int ST_PROD_WARPS = 4;
constexpr int ST_MMA_WARP = ST_PROD_WARPS;
constexpr int ST_EPI_WARP0 = ST_PROD_WARPS + 1;
constexpr int ST_THREADS = (ST_PROD_WARPS + 1 + EPI_WARPS) * 32;   // 416
constexpr int ST_MAX_ROWS = 64;                                      // kt * kh filter rows per tile

struct StemParams {
  int N, Ti, Hi, To, Ho, Wo;
  int kt, kh, st, sh, pt, ph, dt, dh;
  int rows;                 // kt * kh input rows per tile
  int win;                  // window elements per filter row (16 | 32 | 64) = K of one filter row
  int wtiles;               // tiles along W (128 output pixels each)
  int block_n;              // output channels padded to 16
  int stages, nacc, acc_stride, tmem_cols;
  unsigned seg_bytes;       // shared-memory bytes reserved per input row segment
  unsigned w_bytes;         // packed weights
  long long row_pitch;      // bytes between input rows (Wphys * 8)
  long long base_off;       // byte offset of the first window of a row (physical padding - conv padding - lead pixel)
  EpiParams epi;
};

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// no-swizzle K-major descriptor: LBO = bytes between 16-byte K chunks, SBO = bytes between 8-row groups
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__global__ void __launch_bounds__(ST_THREADS, 1)
conv3d_stem_rows_kernel(const __grid_constant__ StemParams P, const unsigned char* __restrict__ x,
                        const unsigned char* __restrict__ w, const unsigned char* __restrict__ zero_row,
                        const float* __restrict__ scale, const float* __restrict__ bias) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stages = P.stages, nacc = P.nacc;
  const uint32_t w_off = 0;                                                  // resident weights
  const uint32_t ring_off = (P.w_bytes + 1023u) & ~1023u;
  const uint32_t stage_bytes = (uint32_t)P.rows * P.seg_bytes + 2048u;       // + slack: the last windows read past a row
  const uint32_t staging_off = (ring_off + (uint32_t)stages * stage_bytes + 1023u) & ~1023u;
  const uint32_t staging = smem_base + staging_off;
  const uint32_t bar_base = staging + (uint32_t)EPI_SMEM_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * stages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * stages + nacc + s); };
  const uint32_t res_bar = bar_base + 8u * (2 * stages + 2 * nacc);
  const uint32_t w_bar = res_bar + 16u;
  const uint32_t tmem_slot = w_bar + 8u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < nacc; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), epi_narrow(P.block_n) ? 4 : EPI_WARPS); }
    mbar_init(res_bar, 1);
    mbar_init(res_bar + 8u, 1);
    mbar_init(w_bar, 1);
    prefetch_tmap(&P.epi.y_map);
    fence_mbar_init();
  }
  if (warp == ST_MMA_WARP) {
    tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int total_tiles = P.N * P.To * P.Ho * P.wtiles;
  auto tile_coords = [&](int tile, int& wt, int& ho, int& to, int& n) {
    wt = tile % P.wtiles; tile /= P.wtiles;
    ho = tile % P.Ho; tile /= P.Ho;
    to = tile % P.To; n = tile / P.To;
  };

  if (warp < ST_PROD_WARPS) {
    // ================================ producers: one bulk copy per input row ================================
    if (warp == 0 && elect_one()) {          // the packed weights, once
      mbar_arrive_expect_tx(w_bar, P.w_bytes);
      for (uint32_t o = 0; o < P.w_bytes; o += 32768u)
        bulk_g2s(smem_base + w_off + o, w + o, min(32768u, P.w_bytes - o), w_bar);
    }
    int stage = 0;
    uint32_t phase = 0;
    const long long frame_pitch = P.row_pitch * P.Hi;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int wt, ho, to, n;
      tile_coords(tile, wt, ho, to, n);
      mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t st_base = smem_base + ring_off + (uint32_t)stage * stage_bytes;
      // bytes of a row this tile reads: 128 windows 16 B apart + the window itself, clipped to the physical row
      const long long col0 = P.base_off + (long long)wt * 128 * 16;
      long long want = 127ll * 16 + (long long)P.win * 2;          // a multiple of 16, like col0 and the row pitch
      if (col0 + want > P.row_pitch) want = P.row_pitch - col0;      // (windows of pixels >= Wo may then see stale bytes: their rows are never stored)
      const uint32_t nbytes = (uint32_t)want;
      if (elect_one()) {
        if (warp == 0) mbar_arrive_expect_tx(full_bar(stage), (uint32_t)P.rows * nbytes);
        for (int r = warp; r < P.rows; r += ST_PROD_WARPS) {
          const int fdt = r / P.kh, fdh = r - fdt * P.kh;
          const int ti = to * P.st - P.pt + fdt * P.dt, hi = ho * P.sh - P.ph + fdh * P.dh;
          const bool ok = (unsigned)ti < (unsigned)P.Ti && (unsigned)hi < (unsigned)P.Hi;
          const unsigned char* src = ok ? x + ((long long)n * P.Ti + ti) * frame_pitch + (long long)hi * P.row_pitch + col0
                                        : zero_row;                       // H / T padding: a row of zeros
          bulk_g2s(st_base + (uint32_t)r * P.seg_bytes, src, nbytes, full_bar(stage));
        }
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == ST_MMA_WARP) {
    // ================================ MMA issuer ============================================================
    const uint32_t idesc = make_idesc_f16(128, P.block_n);
    const int ksteps = P.win >> 4;
    const uint32_t b_lbo = (uint32_t)P.block_n * 16u;          // weights: [K / 8][N][8] - all N rows of a K chunk contiguous
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    mbar_wait(w_bar, 0);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * P.acc_stride);
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      const uint32_t st_base = smem_base + ring_off + (uint32_t)stage * stage_bytes;
      if (elect_one()) {
        for (int r = 0; r < P.rows; ++r) {
          for (int ks = 0; ks < ksteps; ++ks) {
            // A: sliding windows over the raw row; B: K chunks (r * win + 16 ks) / 8 onwards
            const uint64_t a_desc = desc_noswz(st_base + (uint32_t)r * P.seg_bytes + (uint32_t)ks * 32u, 16u, 128u);
            const uint64_t b_desc = desc_noswz(smem_base + w_off + (uint32_t)((r * P.win + ks * 16) >> 3) * b_lbo, b_lbo, 128u);
            umma_f16(d_tmem, a_desc, b_desc, idesc, (r | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(empty_bar(stage));
        umma_commit(tfull_bar(acc));
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1u; }
      if (++acc == nacc) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ================================ epilogue warps ========================================================
    const int quarter = warp & 3;
    const int ewarp = warp - ST_EPI_WARP0;
    int tile_seq = 0;
    uint32_t res_phase = 0;
    const bool narrow = epi_narrow(P.block_n);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_seq) {
      if (narrow && (tile_seq & 1) != (ewarp >> 2)) continue;     // the other group's tile
      const int acc = tile_seq % nacc;
      const uint32_t acc_phase = (uint32_t)((tile_seq / nacc) & 1);
      int wt, ho, to, n;
      tile_coords(tile, wt, ho, to, n);
      if (epi_direct(P.epi)) {
        epilogue_tile_direct(P.epi, scale, bias, tmem_base + (uint32_t)(acc * P.acc_stride), quarter, lane, 0, wt * 128, ho, to, n,
                             tfull_bar(acc), acc_phase, tempty_bar(acc));
        continue;
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      epilogue_tile(P.epi, scale, bias, tmem_base + (uint32_t)(acc * P.acc_stride), staging, smem_gen + staging_off, res_bar,
                    res_phase, ewarp, quarter, lane, 0, wt * 128, ho, to, n, tempty_bar(acc), tile_seq);
    }
    if ((ewarp & 3) == 0 && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == ST_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  }
}

static int stem_window_lead(const pv_conv3d_desc* d) { return (((d->x_w_pad - d->pw) * d->Ci * 2) % 16) ? 1 : 0; }

}  // namespace pv

using namespace pv;

// Eligible: dense conv on the 4-channel, W-padded network input with stride 2 along W (16 bytes per output pixel),
// no residual, output channels <= 256.  The descriptor uses the window-mode conventions of pv_conv3d_desc
// (x_w_pad, x_w_phys, ci_pad64 = window length) but the weights are packed [K / 8][N16][8] (engine/packing.py).
extern "C" int pv_conv3d_stem_rows_supported(const pv_conv3d_desc* d) {
  if (!d || d->dtype != PV_F16 || d->groups != 1 || d->has_residual) return 0;
  if (d->Ci != 4 || d->sw != 2 || d->dw != 1 || d->x_w_pad <= 0 || d->x_row_stride != 4) return 0;
  if (d->x_w_pad < d->pw || (d->x_w_phys * 8) % 16) return 0;
  const int lead = stem_window_lead(d);
  const int run = (d->kw + lead) * 4;
  const int win = run <= 16 ? 16 : (run <= 32 ? 32 : 64);
  if (run > 64 || d->ci_pad64 != win) return 0;
  if (d->Co % 8 || d->Co > 256 || d->y_row_stride % 8) return 0;
  if (d->kt * d->kh > ST_MAX_ROWS) return 0;
  const long long base_off = (long long)(d->x_w_pad - d->pw - lead) * 8;
  if (base_off < 0 || base_off % 16) return 0;
  // every window of the last output pixel must lie inside the physical row
  if (base_off + (long long)(d->Wo - 1) * 16 + (long long)win * 2 > (long long)d->x_w_phys * 8) return 0;
  const int bn = (d->Co + 15) / 16 * 16;
  const long long wbytes = (long long)d->kt * d->kh * win * bn * 2;
  if (wbytes > 120 * 1024) return 0;
  return 1;
}

extern "C" int pv_conv3d_stem_rows_fwd(const pv_conv3d_desc* d, const void* x, const void* w, const float* scale,
                                       const float* bias, const void* zero_row, void* y, void* stream) {
  PV_CHECK_ARG(d && x && w && scale && bias && zero_row && y, "null argument");
  if (!pv_conv3d_stem_rows_supported(d)) { set_error("stem rows kernel: unsupported convolution"); return PV_ERR_UNSUPPORTED; }
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return PV_ERR_CUDA; }
  const int sm_count = current_sm_count();
  if (sm_count <= 0) { set_error("cannot query the SM count"); return PV_ERR_CUDA; }
  StemParams P;
  memset(&P, 0, sizeof(P));
  P.N = d->N; P.Ti = d->Ti; P.Hi = d->Hi; P.To = d->To; P.Ho = d->Ho; P.Wo = d->Wo;
  P.kt = d->kt; P.kh = d->kh; P.st = d->st; P.sh = d->sh; P.pt = d->pt; P.ph = d->ph; P.dt = d->dt; P.dh = d->dh;
  P.rows = d->kt * d->kh;
  P.win = d->ci_pad64;
  P.wtiles = (int)cdiv(d->Wo, 128);
  P.block_n = (d->Co + 15) / 16 * 16;
  P.row_pitch = (long long)d->x_w_phys * 8;
  P.base_off = (long long)(d->x_w_pad - d->pw - stem_window_lead(d)) * 8;
  P.w_bytes = (unsigned)((long long)P.rows * P.win * P.block_n * 2);
  P.seg_bytes = (unsigned)((127 * 16 + P.win * 2 + 127) & ~127);
  P.acc_stride = (P.block_n + 31) / 32 * 32;
  P.nacc = 512 / P.acc_stride;
  if (P.nacc > 8) P.nacc = 8;
  if (P.nacc < 2) P.nacc = 2;
  {
    int cols = P.nacc * P.acc_stride, p2 = 32;
    while (p2 < cols) p2 <<= 1;
    P.tmem_cols = p2;
  }
  const unsigned stage_bytes = (unsigned)P.rows * P.seg_bytes + 2048u;
  {
    const long long budget = 227 * 1024 - 2048 - ((P.w_bytes + 1023) & ~1023u) - EPI_SMEM_BYTES - 512;
    long long st = budget / stage_bytes;
    if (st > 8) st = 8;
    if (st < 2) { set_error("stem rows kernel: not enough shared memory for two stages"); return PV_ERR_UNSUPPORTED; }
    P.stages = (int)st;
  }
  const size_t smem_bytes = 2048 + ((P.w_bytes + 1023) & ~1023u) + (size_t)P.stages * stage_bytes + 1024 + EPI_SMEM_BYTES +
                            8 * (2 * P.stages + 2 * 8 + 8) + 16;
  // ---- epilogue: output tile = box [64 ch, 128 px, 1, 1, 1] of y [Co, Wo, Ho, To, N]
  P.epi.block_n = P.block_n;
  P.epi.Co = d->Co;
  P.epi.rows = d->Wo < 128 ? d->Wo : 128;
  P.epi.act = d->act;
  P.epi.has_residual = 0;
  { const char* e = getenv("PVB200_DEBUG"); P.epi.dbg = e ? atoi(e) : 0; }
  P.epi.y_ptr = (__half*)y;
  P.epi.r_ptr = nullptr;
  const long long ostr[4] = {1, d->Wo, (long long)d->Wo * d->Ho, (long long)d->Wo * d->Ho * d->To};
  const int O[4] = {d->Wo, d->Ho, d->To, d->N};
  const int box[4] = {P.epi.rows, 1, 1, 1};
  for (int m = 0; m < 4; ++m) {
    P.epi.O[m] = O[m];
    P.epi.box[m] = box[m];
    P.epi.y_str[m] = ostr[m] * d->y_row_stride;
    P.epi.r_str[m] = 0;
  }
  {
    cuuint64_t gdim[5] = {(cuuint64_t)d->Co, (cuuint64_t)O[0], (cuuint64_t)O[1], (cuuint64_t)O[2], (cuuint64_t)O[3]};
    cuuint64_t gstr[4];
    cuuint32_t bx[5] = {64, (cuuint32_t)box[0], 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
    for (int m = 0; m < 4; ++m) gstr[m] = (cuuint64_t)(ostr[m] * d->y_row_stride * 2);
    CUresult cr = encode(&P.epi.y_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, y, gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(stem Y) failed: %d", (int)cr); return PV_ERR_CUDA; }
  }
  PV_OPT_IN_SMEM(conv3d_stem_rows_kernel, 227 * 1024);
  const long long total_tiles = (long long)d->N * d->To * d->Ho * P.wtiles;
  if (total_tiles == 0) return PV_OK;
  PV_CHECK_ARG(total_tiles < (1ll << 31), "too many tiles");
  const int grid = (int)(total_tiles < sm_count ? total_tiles : sm_count);
  {
    static const bool use_pdl = getenv("PVB200_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(ST_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    PV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv3d_stem_rows_kernel, P, (const unsigned char*)x, (const unsigned char*)w,
                                  (const unsigned char*)zero_row, scale, bias));
  }
  PV_LAUNCH_OK("conv3d_stem_rows_kernel");
  return PV_OK;
}
