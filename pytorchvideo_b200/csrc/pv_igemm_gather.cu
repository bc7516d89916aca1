// Implicit-GEMM convolution for NARROW inputs (C_in < 64: the 3-channel stems, the SlowFast Fast
// pathway C=8..32, X3D's 24/48/56-wide tensors) on tcgen05 tensor cores.
//
// With few input channels one filter tap is only 8..64 bytes of K, so a 64-channel TMA box per tap
// would be mostly zero fill.  Here the GEMM K axis is the flattened (tap, ci) index and the A tile
// (128 output positions x 64 K-elements, 128B-swizzled K-major) is assembled by 7 producer warps
// with coalesced 8/16-byte global loads (im2col gather, zero fill for padding) written straight to
// the swizzled shared-memory layout the UMMA descriptor expects; weights still arrive by TMA.
// MMA issue, TMEM double buffering and the fused epilogue are the same as in pv_igemm.cu.
#include "pv_common.cuh"
#include "pv_sm100.cuh"
#include "pv_epilogue.cuh"

#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace pv {

using namespace sm100;

constexpr int GG_BM = 128;
constexpr int GG_BK = 64;
constexpr int GG_A_BYTES = GG_BM * GG_BK * 2;
constexpr int GG_MAX_UNITS = 256;
constexpr int GG_EPI_WARPS = EPI_WARPS;   // 8
constexpr int GG_PROD_WARPS = 7;          // 16 warps in all: 4 per SM sub-partition -> 128 registers per thread
constexpr int GG_THREADS = (GG_EPI_WARPS + 1 + GG_PROD_WARPS) * 32;   // 512

struct GatherParams {
  CUtensorMap b_map;
  int N, Ti, Hi, Wi, To, Ho, Wo;
  int st, sh, sw, pt, ph, pw;
  int gbytes;        // gather unit: 8 or 16 bytes
  int units_total;   // taps * (Ci*2/gbytes)
  int upk;           // units per 64-element k-block: 128 / gbytes
  int upt, taps;     // units per tap, taps (<= 64: one validity bit per tap)
  int num_kb;
  long long x_row_stride;
  long long M;
  int m_tiles, n_tiles, block_n, Co, stages, tmem_cols, acc_stride, nacc;
  int nprod;   // active producer warps; stages is a multiple of nprod so every smem slot has ONE owner warp
  int depth;   // k-blocks of cp.async a producer warp keeps in flight: 2 when it owns >= 2 slots, else 1
  int epi_bytes;   // shared staging of the TMA-store epilogue (0 when the direct epilogue is used)
  EpiParams epi;
  int unit_off[GG_MAX_UNITS];        // element offset of the unit relative to the row's (t0,h0,w0) corner
  unsigned int unit_d[GG_MAX_UNITS];  // packed (dt | dh<<8 | dw<<16) tap displacement (dilation applied)
};

__global__ void __launch_bounds__(GG_THREADS, 1)
conv3d_igemm_gather_kernel(const __grid_constant__ GatherParams P, const __half* __restrict__ x,
                           const float* __restrict__ scale, const float* __restrict__ bias) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stages = P.stages;
  const uint32_t b_bytes = (uint32_t)P.block_n * GG_BK * 2;
  const uint32_t stage_bytes = GG_A_BYTES + b_bytes;
  const uint32_t staging_off = (uint32_t)((stages * stage_bytes + 1023u) & ~1023u);
  const uint32_t staging = smem_base + staging_off;
  const uint32_t bar_base = staging + (uint32_t)P.epi_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * stages + s); };
  const int nacc = P.nacc;                   // accumulator stages in TMEM (2..8)
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * stages + nacc + s); };
  const uint32_t res_bar = bar_base + 8u * (2 * stages + 2 * nacc);
  const uint32_t tmem_slot = bar_base + 8u * (2 * stages + 2 * nacc + 2);

  __shared__ int s_off[GG_MAX_UNITS];
  __shared__ unsigned s_d[GG_MAX_UNITS];
  for (int i = threadIdx.x; i < GG_MAX_UNITS; i += blockDim.x) { s_off[i] = P.unit_off[i]; s_d[i] = P.unit_d[i]; }
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int MMA_WARP = GG_EPI_WARPS;          // warp 8
  constexpr int PROD_WARP0 = GG_EPI_WARPS + 1;    // warps 9..15

  if (warp == MMA_WARP && lane == 0) {
    prefetch_tmap(&P.b_map);
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 2);   // the owning producer warp's arrive + its expect_tx arrive
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < nacc; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), epi_narrow(P.block_n) ? 4 : GG_EPI_WARPS);
    }
    mbar_init(res_bar, 1);
    mbar_init(res_bar + 8u, 1);
    prefetch_tmap(&P.epi.y_map);
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch)
  // overlaps the tail of the previous kernel in the stream / graph; its results are only touched below.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int total_tiles = P.n_tiles * P.m_tiles;

  if (warp >= PROD_WARP0) {
    // ================================ gather producers =======================================
    // Warp-per-k-block im2col gather.  Producer warp w owns k-blocks g = w, w+nprod, ... of this CTA's
    // (tile, k-block) sequence and fills the whole 128 x 64 A tile of that k-block alone: lane l
    // covers rows l, l+32, l+64, l+96, each unit is a zero-filling cp.async straight into the
    // swizzled layout.  All producer warps are in flight on different k-blocks, so the serial issue
    // latency of one thread (measured ~1000 clk per 4 copies when every warp worked on the SAME
    // k-block) no longer bounds the pipeline; a warp publishes its k-block (proxy fence + one
    // mbarrier arrive) when its own copies have landed.
    const int wprod = warp - PROD_WARP0;
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_g = my_tiles * P.num_kb;
    int cur_seq = -1;
    const __half* xrow[4];
    unsigned long long vmask[4];     // bit `tap` set <=> row q's input position for that tap is inside the tensor
    int n_tile = 0;
    const unsigned Ti = (unsigned)P.Ti, Hi = (unsigned)P.Hi, Wi = (unsigned)P.Wi;
    const uint32_t rsw = (uint32_t)(lane & 7);           // (lane + 32q) & 7
    const uint32_t row_off = (uint32_t)lane * 128u;
    // Slot ownership: stages % nprod == 0, so slot s is only ever filled by warp s % nprod, in order;
    // a parity wait can then never alias a completion two phases back (it could with free-running
    // warps sharing slots).
    // ncu source view: the producers wait on their own cp.async data and on the empty barrier about equally.
    // Optional depth 2 (PVB200_GATHER_DEPTH2): with >= 2 slots per warp a k-block is published one iteration
    // later (cp.async groups), i.e. two k-blocks of copies in flight per warp - measured no faster, off by default.
    const bool deep = P.depth >= 2;
    int pending = -1;                // stage whose copies were issued last iteration and are not yet published
    for (int g = wprod; wprod < P.nprod && g < total_g; g += P.nprod) {
      const int tile_seq = g / P.num_kb;
      const int kb = g - tile_seq * P.num_kb;
      const int stage = g % stages;
      const uint32_t phase = (uint32_t)((g / stages) & 1);
      if (tile_seq != cur_seq) {
        // Per tile: row corners + one validity bit per (row, tap).  Everything in the copy loop below is
        // then branch-free (the bounds tests used to compile into divergent short-circuit branches with a
        // constant-bank load each: ~4000 clk of exposed latency per k-block).
        cur_seq = tile_seq;
        const int tile = (int)blockIdx.x + tile_seq * (int)gridDim.x;
        n_tile = tile % P.n_tiles;
        const int m_tile = tile / P.n_tiles;
        int t0[4], h0[4], w0[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const long long m = (long long)m_tile * GG_BM + lane + 32 * q;
          const bool rok = m < P.M;
          const uint32_t mu = rok ? (uint32_t)m : 0u;
          uint32_t r = mu / (uint32_t)P.Wo;
          const int wo = (int)(mu - r * (uint32_t)P.Wo);
          uint32_t r2 = r / (uint32_t)P.Ho;
          const int ho = (int)(r - r2 * (uint32_t)P.Ho);
          const uint32_t n_u = r2 / (uint32_t)P.To;
          const int to = (int)(r2 - n_u * (uint32_t)P.To);
          t0[q] = to * P.st - P.pt; h0[q] = ho * P.sh - P.ph; w0[q] = wo * P.sw - P.pw;
          if (!rok) t0[q] = -100000;        // every tap out of range -> all-zero row
          xrow[q] = x + ((((long long)n_u * P.Ti + t0[q]) * P.Hi + h0[q]) * P.Wi + w0[q]) * P.x_row_stride;
          vmask[q] = 0ull;
        }
        for (int tap = 0; tap < P.taps; ++tap) {
          const unsigned dd = s_d[tap * P.upt];
          const int dt_ = (int)(dd & 0xffu), dh_ = (int)((dd >> 8) & 0xffu), dw_ = (int)((dd >> 16) & 0xffu);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const unsigned ok = (unsigned)((unsigned)(t0[q] + dt_) < Ti) & (unsigned)((unsigned)(h0[q] + dh_) < Hi) &
                                (unsigned)((unsigned)(w0[q] + dw_) < Wi);
            vmask[q] |= (unsigned long long)ok << tap;
          }
        }
      }
      mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t a_tile = smem_base + stage * stage_bytes;
      if (elect_one()) {
        mbar_arrive_expect_tx(full_bar(stage), b_bytes);
        tma_load_2d(a_tile + GG_A_BYTES, &P.b_map, full_bar(stage), kb * GG_BK, n_tile * P.block_n);
      }
      __syncwarp();
      const int u_base = kb * P.upk;
      int units_here = P.units_total - u_base;
      if (units_here > P.upk) units_here = P.upk;
      // zero-fill only up to the end of the last 16-element MMA step that holds data (the MMA warp issues
      // exactly those steps); smem beyond it is never read
      const int need = (((units_here * P.gbytes + 31) >> 5) << 5) / P.gbytes;
      const bool live = !(P.epi.dbg & 4);
      const uint32_t rbase = a_tile + row_off;
      if (P.gbytes == 16) {
        for (int ui = 0; ui < need; ++ui) {
          const bool uok = live && ui < units_here;
          const int uu = uok ? u_base + ui : 0;
          const int off = s_off[uu];
          const unsigned tap = s_d[uu] >> 24;
          const uint32_t dst = rbase + (((uint32_t)ui ^ rsw) << 4);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool ok = uok && ((vmask[q] >> tap) & 1ull);
            const __half* src = ok ? xrow[q] + off : x;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)q * 4096u), "l"(src),
                         "r"(ok ? 16u : 0u) : "memory");
          }
        }
      } else {
        for (int ui = 0; ui < need; ++ui) {
          const bool uok = live && ui < units_here;
          const int uu = uok ? u_base + ui : 0;
          const int off = s_off[uu];
          const unsigned tap = s_d[uu] >> 24;
          const uint32_t dst = rbase + ((((uint32_t)ui >> 1) ^ rsw) << 4) + (((uint32_t)ui & 1u) << 3);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool ok = uok && ((vmask[q] >> tap) & 1ull);
            const __half* src = ok ? xrow[q] + off : x;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + (uint32_t)q * 4096u), "l"(src),
                         "r"(ok ? 8u : 0u) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (deep) {
        if (pending >= 0) {
          asm volatile("cp.async.wait_group 1;" ::: "memory");     // everything but the group just committed has landed
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) mbar_arrive(full_bar(pending));
          __syncwarp();
        }
        pending = stage;
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one()) mbar_arrive(full_bar(stage));
      }
    }
    if (pending >= 0) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (elect_one()) mbar_arrive(full_bar(pending));
    }
  } else if (warp == MMA_WARP) {
    // ================================ MMA issuer ============================================
    // The WHOLE warp walks the (tile, k-block) sequence with warp-uniform control flow and waits on the
    // barriers; one elected lane issues the tcgen05 instructions.  (Running the loop inside
    // `if (lane == 0)` makes every tcgen05.mma / commit a divergent-code instruction: the compiler then
    // moves each operand to a uniform register with R2UR inside an ELECT / BRA.U.ANY waterfall.)
    const uint32_t idesc = make_idesc_f16(GG_BM, P.block_n);
    const int num_kb = P.num_kb;
    const int k16_full = (P.upk * P.gbytes) >> 5;
    const int k16_last = (((P.units_total - (num_kb - 1) * P.upk) * P.gbytes) + 31) >> 5;   // steps that hold data
    const bool skip_mma = (P.epi.dbg & 32) != 0;
    const int acc_stride = P.acc_stride;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_stride);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a_addr = smem_base + stage * stage_bytes;
        const uint64_t a_desc = make_kmajor_desc(a_addr, 128);
        const uint64_t b_desc = make_kmajor_desc(a_addr + GG_A_BYTES, 128);
        const bool last = kb == num_kb - 1;
        const int k16 = skip_mma ? 0 : (last ? k16_last : k16_full);
        if (elect_one()) {
          for (int k = 0; k < k16; ++k)
            umma_f16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (last) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
      if (++acc == nacc) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ================================ epilogue warps ========================================
    const int quarter = warp & 3;
    int tile_seq = 0;
    uint32_t res_phase = 0;
    const bool narrow = epi_narrow(P.block_n);
    const bool direct = epi_direct(P.epi);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_seq) {
      if (narrow && (tile_seq & 1) != (warp >> 2)) continue;      // the other group's tile
      const int acc = tile_seq % nacc;
      const uint32_t acc_phase = (uint32_t)((tile_seq / nacc) & 1);
      const int n_tile = tile % P.n_tiles;
      const int m_tile = tile / P.n_tiles;
      if (direct) {
        epilogue_tile_direct(P.epi, scale, bias, tmem_base + (uint32_t)(acc * P.acc_stride), quarter, lane,
                             n_tile * P.block_n, m_tile * GG_BM, 0, 0, 0, tfull_bar(acc), acc_phase, tempty_bar(acc));
        continue;
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      epilogue_tile(P.epi, scale, bias, tmem_base + (uint32_t)(acc * P.acc_stride), staging, smem_gen + staging_off,
                    res_bar, res_phase, warp, quarter, lane, n_tile * P.block_n, m_tile * GG_BM, 0, 0, 0,
                    tempty_bar(acc), tile_seq);
    }
    if ((warp & 3) == 0 && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  }
}

// =============================================================================================
// Host side
// =============================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();   // pv_igemm.cu

int conv3d_gather_supported(const pv_conv3d_desc* d) {
  if (d->dtype != PV_F16 || d->groups != 1) return 0;
  if (!(d->Ci == 4 || (d->Ci % 8 == 0 && d->Ci < 64))) return 0;
  if (d->Co % 8) return 0;
  if (d->ci_pad64 != d->Ci) return 0;        // weights packed with the un-padded per-tap K extent
  const int gbytes = d->Ci == 4 ? 8 : 16;
  const int units = d->kt * d->kh * d->kw * (d->Ci * 2 / gbytes);
  if (units > GG_MAX_UNITS) return 0;
  if (d->x_row_stride % (gbytes / 2) || d->y_row_stride % 8 || (d->has_residual && d->res_row_stride % 8)) return 0;
  if (d->dt * (d->kt - 1) > 255 || d->dh * (d->kh - 1) > 255 || d->dw * (d->kw - 1) > 255) return 0;
  if (d->kt * d->kh * d->kw > 64) return 0;   // one validity bit per tap
  const long long M = (long long)d->N * d->To * d->Ho * d->Wo;
  if (M >= (1ll << 31)) return 0;
  return 1;
}

int conv3d_gather_launch(const pv_conv3d_desc* d, const void* x, const void* w, const float* scale,
                         const float* bias, const void* residual, void* y, cudaStream_t stream) {
  if (!conv3d_gather_supported(d)) {
    set_error("gather tcgen05 path does not support this convolution");
    return PV_ERR_UNSUPPORTED;
  }
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return PV_ERR_CUDA; }
  const int sm_count = current_sm_count();
  if (sm_count <= 0) { set_error("cannot query the SM count of the current device"); return PV_ERR_CUDA; }
  GatherParams P;
  memset(&P, 0, sizeof(P));
  P.N = d->N; P.Ti = d->Ti; P.Hi = d->Hi; P.Wi = d->Wi; P.To = d->To; P.Ho = d->Ho; P.Wo = d->Wo;
  P.st = d->st; P.sh = d->sh; P.sw = d->sw; P.pt = d->pt; P.ph = d->ph; P.pw = d->pw;
  P.gbytes = d->Ci == 4 ? 8 : 16;
  const int upt = d->Ci * 2 / P.gbytes;
  const int taps = d->kt * d->kh * d->kw;
  P.units_total = taps * upt;
  P.upt = upt; P.taps = taps;
  P.upk = 128 / P.gbytes;
  P.num_kb = (P.units_total + P.upk - 1) / P.upk;
  P.x_row_stride = d->x_row_stride;
  P.M = (long long)d->N * d->To * d->Ho * d->Wo;
  P.m_tiles = (int)cdiv(P.M, GG_BM);
  P.Co = d->Co;
  for (int it = 0; it < d->kt; ++it)
    for (int ih = 0; ih < d->kh; ++ih)
      for (int iw = 0; iw < d->kw; ++iw) {
        const int tap = (it * d->kh + ih) * d->kw + iw;
        const int dt = it * d->dt, dh = ih * d->dh, dw = iw * d->dw;
        const long long off = (((long long)dt * d->Hi + dh) * d->Wi + dw) * d->x_row_stride;
        for (int q = 0; q < upt; ++q) {
          const int u = tap * upt + q;
          const long long o = off + (long long)q * (P.gbytes / 2);
          if (o > 0x7fffffffll) { set_error("gather offset overflow"); return PV_ERR_UNSUPPORTED; }
          P.unit_off[u] = (int)o;
          P.unit_d[u] = (unsigned)dt | ((unsigned)dh << 8) | ((unsigned)dw << 16) | ((unsigned)tap << 24);
        }
      }
  {
    const int co16 = (int)cdiv(d->Co, 16) * 16;
    int bn = co16 <= 256 ? co16 : 256;
    // keep a few hundred tiles in flight for small layers (several N tiles must be multiples of 64)
    while (bn > 64 && (long long)P.m_tiles * cdiv(d->Co, bn) < 2 * sm_count && (bn / 2) % 64 == 0) bn /= 2;
    P.block_n = bn;
    P.n_tiles = (int)cdiv(d->Co, bn);
  }
  {
    P.acc_stride = (P.block_n + 31) / 32 * 32;
    P.nacc = 512 / P.acc_stride;
    if (P.nacc > 8) P.nacc = 8;
    if (P.nacc < 2) P.nacc = 2;
    int cols = P.nacc * P.acc_stride, p2 = 32;
    while (p2 < cols) p2 <<= 1;
    P.tmem_cols = p2;
  }
  P.epi.block_n = P.block_n;
  P.epi.Co = d->Co;
  P.epi.rows = GG_BM;
  P.epi.act = d->act;
  P.epi.has_residual = d->has_residual;
  { const char* e = getenv("PVB200_DEBUG"); P.epi.dbg = e ? atoi(e) : 0; }
  P.epi_bytes = epi_direct(P.epi) ? 0 : EPI_SMEM_BYTES;     // the direct epilogue needs no staging: more stages
  const int stage_bytes = GG_A_BYTES + P.block_n * GG_BK * 2;
  {
    int st = (227 * 1024 - 2048 - 2048 /*static tables*/ - P.epi_bytes - 512) / stage_bytes;
    if (st > 16) st = 16;
    if (st < 2) { set_error("gather: not enough smem stages"); return PV_ERR_UNSUPPORTED; }
    static const bool shallow = getenv("PVB200_GATHER_DEPTH2") == nullptr;   // measured: depth 2 is not faster (SlowFast 3.45 vs 3.42 ms), opt-in only
    // prefer two slots per producer warp (two k-blocks of copies in flight) over more warps with one slot
    P.nprod = st < GG_PROD_WARPS ? st : GG_PROD_WARPS;
    if (!shallow && st >= 8 && st / 2 < P.nprod) P.nprod = st / 2;
    P.stages = (st / P.nprod) * P.nprod;
    P.depth = (!shallow && P.stages >= 2 * P.nprod) ? 2 : 1;
  }
  const size_t smem_bytes = (size_t)P.stages * stage_bytes + 2048 + P.epi_bytes + 8 * (2 * P.stages + 2 * 8 + 4) + 16;
  P.epi.y_ptr = (__half*)y;
  P.epi.r_ptr = (const __half*)residual;
  for (int m = 0; m < 4; ++m) { P.epi.O[m] = 1; P.epi.box[m] = 1; P.epi.y_str[m] = 0; P.epi.r_str[m] = 0; }
  P.epi.O[0] = (int)P.M; P.epi.box[0] = GG_BM;
  P.epi.y_str[0] = d->y_row_stride; P.epi.r_str[0] = d->res_row_stride;
  for (int pass = 0; pass < 2; ++pass) {     // output / residual as [Co, M, 1, 1, 1]
    if (pass == 1 && !d->has_residual) break;
    const long long rs = pass == 0 ? d->y_row_stride : d->res_row_stride;
    void* base = pass == 0 ? y : const_cast<void*>(residual);
    cuuint64_t gdim[5] = {(cuuint64_t)d->Co, (cuuint64_t)P.M, 1, 1, 1};
    cuuint64_t gstr[4] = {(cuuint64_t)rs * 2, (cuuint64_t)rs * 2 * (cuuint64_t)P.M, (cuuint64_t)rs * 2 * (cuuint64_t)P.M,
                          (cuuint64_t)rs * 2 * (cuuint64_t)P.M};
    cuuint32_t box[5] = {64, GG_BM, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
    CUresult cr = encode(pass == 0 ? &P.epi.y_map : &P.epi.r_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, base, gdim, gstr,
                         box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(gather %s) failed: %d", pass ? "R" : "Y", (int)cr); return PV_ERR_CUDA; }
  }
  {
    const long long kpad = (long long)cdiv((long long)taps * d->Ci, 64) * 64;
    cuuint64_t gdim[2] = {(cuuint64_t)kpad, (cuuint64_t)d->Co};
    cuuint64_t gstr[1] = {(cuuint64_t)kpad * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)P.block_n}, estr[2] = {1, 1};
    CUresult cr = encode(&P.b_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)w, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(B, gather) failed: %d", (int)cr); return PV_ERR_CUDA; }
  }
  PV_OPT_IN_SMEM(conv3d_igemm_gather_kernel, 225 * 1024);
  const long long total_tiles = (long long)P.m_tiles * P.n_tiles;
  if (total_tiles == 0) return PV_OK;
  const int grid = (int)(total_tiles < sm_count ? total_tiles : sm_count);
  {
    // launched with the programmatic-stream-serialization attribute (PDL); PVB200_NO_PDL=1 falls back to a plain launch
    static const bool use_pdl = getenv("PVB200_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(GG_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    PV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv3d_igemm_gather_kernel, P, (const __half*)x, scale, bias));
  }
  PV_LAUNCH_OK("conv3d_igemm_gather_kernel");
  return PV_OK;
}

}  // namespace pv
