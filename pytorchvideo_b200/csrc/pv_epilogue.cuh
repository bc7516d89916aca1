// Fused epilogue shared by the TMA-fed and the gather-fed implicit-GEMM kernels.
//
// TMEM accumulator (128 rows x BLOCK_N fp32) -> registers (tcgen05.ld) -> y = act(acc*scale+bias
// (+residual)) -> f16 -> 128B-swizzled shared staging -> TMA tensor store.  The residual tile is
// brought in by a TMA tensor load into the SAME staging buffer and updated in place, so both the
// residual read and the output write are full-line bulk transfers issued by one thread, with the
// tile-edge and channel-tail clipping done by the TMA unit (no per-thread masks or address math).
//
// 8 epilogue warps: warp e owns TMEM lane quarter (e & 3) and the 64-column half (e >> 2) of every
// 128-column group; folded-BN scale/bias for the tile are staged once in shared memory.  Staging is
// 32 KiB = one 128-column group, or two alternating 64-column slots when BLOCK_N <= 64 so that the
// previous tile's bulk store can still be draining while the next tile is staged.
#pragma once
#include "pv_common.cuh"
#include "pv_sm100.cuh"

namespace pv {
namespace sm100 {

constexpr int EPI_GROUP_COLS = 128;
constexpr int EPI_STAGING_BYTES = 2 * 128 * 128;   // two [128 rows x 64 f16] swizzled sub-tiles
constexpr int EPI_SB_BYTES = 2 * 2 * 256 * 4;      // scale[256] + bias[256] (fp32), one copy per epilogue group
constexpr int EPI_SMEM_BYTES = EPI_STAGING_BYTES + EPI_SB_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;

struct EpiParams {
  CUtensorMap y_map;    // output  [Co, d1, d2, d3, d4], box [64, b1, b2, b3, b4], SWIZZLE_128B
  CUtensorMap r_map;    // residual, same geometry
  int block_n, Co, rows, act, has_residual;
  int teams;   // wide residual tiles: 1 = eight warps on one tile, 2 = two teams of four warps on alternate tiles (epilogue_tile_ring_teams)
  int dbg;   // debug bit mask (PVB200_DEBUG env): 1 = skip stores, 2 = skip epilogue math, 4 = producers skip loads,
             // 32 = MMA warp skips the MMAs, 64 = flip the direct / TMA-staged epilogue choice (see epi_direct)
  // direct (register -> global) epilogue for BLOCK_N <= 64: row r of a tile decodes into box coordinates
  // (dim 0 fastest, the order the A-operand TMA box lands in shared memory) -> element offsets
  __half* y_ptr;
  const __half* r_ptr;
  long long y_str[4], r_str[4];   // element stride of +1 in merged output dim i
  int O[4], box[4];               // merged output extents / tile box
};

__device__ __forceinline__ void tma_store_5d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(tmap)),
      "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// BLOCK_N <= 64 ("narrow"): the 8 epilogue warps split into two independent groups of 4 that take
// alternate tiles, each with its own staging slot, scale/bias copy, residual barrier and named
// barrier - the per-tile epilogue is a latency chain, so two tiles in flight double its throughput.
__host__ __device__ inline bool epi_narrow(int block_n) { return block_n <= 64; }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

template <int ACT>
__device__ __forceinline__ float act_t(float x) {
  if (ACT == PV_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == PV_ACT_NONE) return x;
  return apply_act(x, ACT);
}

// One 64-column sub-tile of one row: two x32 TMEM loads, math, swizzled 16-byte cells.
template <int ACT, bool RES>
__device__ __forceinline__ void epi_subtile(uint32_t t_addr, uint8_t* srow, uint32_t rsw, const float* sc,
                                            const float* bi, int ncols) {
#pragma unroll 1
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(t_addr + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int cc = c0 + h * 8;
      uint4* cell = reinterpret_cast<uint4*>(srow + ((((uint32_t)cc >> 3) ^ rsw) << 4));
      const float4 s0 = *reinterpret_cast<const float4*>(sc + cc), s1 = *reinterpret_cast<const float4*>(sc + cc + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(bi + cc), b1 = *reinterpret_cast<const float4*>(bi + cc + 4);
      float f[8];
      f[0] = fmaf(__uint_as_float(v[h * 8 + 0]), s0.x, b0.x);
      f[1] = fmaf(__uint_as_float(v[h * 8 + 1]), s0.y, b0.y);
      f[2] = fmaf(__uint_as_float(v[h * 8 + 2]), s0.z, b0.z);
      f[3] = fmaf(__uint_as_float(v[h * 8 + 3]), s0.w, b0.w);
      f[4] = fmaf(__uint_as_float(v[h * 8 + 4]), s1.x, b1.x);
      f[5] = fmaf(__uint_as_float(v[h * 8 + 5]), s1.y, b1.y);
      f[6] = fmaf(__uint_as_float(v[h * 8 + 6]), s1.z, b1.z);
      f[7] = fmaf(__uint_as_float(v[h * 8 + 7]), s1.w, b1.w);
      if (RES) {
        const uint4 rv = *cell;
        const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 r2 = __half22float2(rh[q]);
          f[2 * q] += r2.x;
          f[2 * q + 1] += r2.y;
        }
      }
      uint4 ov;
      __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
      for (int q = 0; q < 4; ++q) oh[q] = __floats2half2_rn(act_t<ACT>(f[2 * q]), act_t<ACT>(f[2 * q + 1]));
      *cell = ov;
    }
  }
}

template <int ACT>
__device__ __forceinline__ void epi_subtile_res(bool res, uint32_t t_addr, uint8_t* srow, uint32_t rsw,
                                                const float* sc, const float* bi, int ncols) {
  if (res) epi_subtile<ACT, true>(t_addr, srow, rsw, sc, bi, ncols);
  else epi_subtile<ACT, false>(t_addr, srow, rsw, sc, bi, ncols);
}

// Called by the 8 epilogue warps (256 threads) for one output tile.
//   ewarp     : 0..7 index of this warp among the epilogue warps; `quarter` = hardware warp id & 3
//   t_acc     : TMEM address of the accumulator (column base), lane field NOT yet applied
//   epi_smem  : shared address (1024-aligned) / generic pointer of the EPI_SMEM_BYTES region
//   tile_seq  : running tile counter of this CTA (selects the staging slot when BLOCK_N <= 64)
__device__ __forceinline__ void epilogue_tile(const EpiParams& E, const float* __restrict__ scale,
                                              const float* __restrict__ bias, uint32_t t_acc,
                                              uint32_t epi_smem, uint8_t* epi_gen, uint32_t res_bar,
                                              uint32_t& res_phase, int ewarp, int quarter, int lane, int n0,
                                              int c1, int c2, int c3, int c4, uint32_t tempty_bar,
                                              int tile_seq) {
  const int row = quarter * 32 + lane;
  const bool narrow = epi_narrow(E.block_n);
  const int group = narrow ? (ewarp >> 2) : 0;   // narrow: two groups of 4 warps on alternate tiles
  const int chalf = narrow ? 0 : (ewarp >> 2);
  const int etid = narrow ? ((ewarp & 3) * 32 + lane) : (ewarp * 32 + lane);
  const int nthr = narrow ? 128 : 256;
  const int bar_id = 1 + group;
  const bool leader = (etid == 0);
  const uint32_t t_row = t_acc + ((uint32_t)(quarter * 32) << 16);
  const uint32_t rsw = (uint32_t)(row & 7);
  float* sb = reinterpret_cast<float*>(epi_gen + EPI_STAGING_BYTES) + group * 512;
  res_bar += 8u * (uint32_t)group;
  (void)tile_seq;
  // stage scale / bias of this N tile (zeros beyond Co keep pad lanes at exactly zero)
  for (int i = etid; i < E.block_n; i += nthr) {
    const int c = n0 + i;
    const bool ok = c < E.Co;
    sb[i] = ok ? __ldg(scale + c) : 0.f;
    sb[256 + i] = ok ? __ldg(bias + c) : 0.f;
  }
  for (int g0 = 0; g0 < E.block_n; g0 += EPI_GROUP_COLS) {
    const int gcols = min(EPI_GROUP_COLS, E.block_n - g0);
    const int nsub = (gcols + 63) >> 6;
    const uint32_t slot = (uint32_t)group * 16384u;
    // (a) the staging slot is free once the bulk store that last used it has read it out
    if (leader) {
      tma_store_wait_read0();
      if (E.has_residual) {
        mbar_arrive_expect_tx(res_bar, (uint32_t)(nsub * E.rows * 128));
        for (int s = 0; s < nsub; ++s)
          tma_load_5d(epi_smem + slot + (uint32_t)s * 16384u, &E.r_map, res_bar, n0 + g0 + s * 64, c1, c2, c3, c4);
      }
    }
    epi_bar_sync(bar_id, nthr);     // slot free + scale/bias visible
    if (E.has_residual) {
      mbar_wait(res_bar, res_phase);
      res_phase ^= 1u;
    }
    // (b) drain this warp's 64-column half of the group
    if (chalf < nsub && !(E.dbg & 2)) {
      const int cbase = g0 + chalf * 64;
      const int ncols = min(64, gcols - chalf * 64);
      uint8_t* srow = epi_gen + slot + chalf * 16384 + row * 128;
      const float* sc = sb + cbase;
      const float* bi = sb + 256 + cbase;
      switch (E.act) {
        case PV_ACT_RELU: epi_subtile_res<PV_ACT_RELU>(E.has_residual, t_row + (uint32_t)cbase, srow, rsw, sc, bi, ncols); break;
        case PV_ACT_NONE: epi_subtile_res<PV_ACT_NONE>(E.has_residual, t_row + (uint32_t)cbase, srow, rsw, sc, bi, ncols); break;
        case PV_ACT_SWISH: epi_subtile_res<PV_ACT_SWISH>(E.has_residual, t_row + (uint32_t)cbase, srow, rsw, sc, bi, ncols); break;
        case PV_ACT_GELU: epi_subtile_res<PV_ACT_GELU>(E.has_residual, t_row + (uint32_t)cbase, srow, rsw, sc, bi, ncols); break;
        default: epi_subtile_res<PV_ACT_SIGMOID>(E.has_residual, t_row + (uint32_t)cbase, srow, rsw, sc, bi, ncols); break;
      }
    }
    if (g0 + EPI_GROUP_COLS >= E.block_n) {   // accumulator fully read: hand TMEM back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar);   // shared::cluster address: own CTA, or the pair's leader
    }
    // (c) publish the staged tile to the async proxy and store it
    if (!(E.dbg & 16)) fence_proxy_async_smem();
    epi_bar_sync(bar_id, nthr);
    if (leader && !(E.dbg & 1)) {
      for (int s = 0; s < nsub; ++s)
        tma_store_5d(&E.y_map, epi_smem + slot + (uint32_t)s * 16384u, n0 + g0 + s * 64, c1, c2, c3, c4);
      tma_store_commit();
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Direct epilogue for narrow tiles (BLOCK_N <= 64).  Measured on B200 (tools/narrow_probe.py): the
// TMA-staged path is a per-tile latency chain of two TMA round trips (residual load, then the
// store's shared-memory read that frees the staging slot; ~1500 clk each) which bounds a narrow
// layer at ~2400 clk per 128-row tile.  A thread owns one output row; with <= 64 channels that row
// is <= 128 contiguous bytes, so plain 16-byte global loads/stores are already full-sector
// transfers.  The residual row is requested BEFORE the accumulator barrier is waited on (its
// latency hides behind the MMA), the accumulator is handed back right after the last tcgen05.ld,
// and there is no shared staging and no CTA-level barrier at all.  Called per warp; the two groups
// of 4 epilogue warps take alternate tiles.
// ---------------------------------------------------------------------------------------------
template <int ACT, bool RES, int MAXC>     // MAXC: 16-byte chunks per row this instantiation covers (4: BLOCK_N <= 32, 8: <= 64)
__device__ __forceinline__ void epi_direct_row(const EpiParams& E, const float* __restrict__ scale,
                                               const float* __restrict__ bias, uint32_t t_row, __half* yrow,
                                               const __half* rrow, bool row_ok, int n0, uint32_t tfull_bar,
                                               uint32_t tfull_phase, uint32_t tempty_bar, int lane) {
  uint4 rv[MAXC];
  if (RES) {
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
      rv[j] = make_uint4(0u, 0u, 0u, 0u);
      if (j * 8 < E.block_n && row_ok && n0 + j * 8 < E.Co) rv[j] = *reinterpret_cast<const uint4*>(rrow + j * 8);
    }
  }
  mbar_wait(tfull_bar, tfull_phase);
  tc_fence_after();
#pragma unroll
  for (int c0 = 0; c0 < MAXC * 8; c0 += 32) {
    if (c0 < E.block_n) {
      uint32_t v[32];
      tmem_ld32(t_row + (uint32_t)c0, v);
      tmem_ld_wait();
      if (c0 + 32 >= E.block_n) {        // accumulator fully read: hand TMEM back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar);   // shared::cluster address: own CTA, or the pair's leader
      }
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int cc = c0 + h * 8;
        const int c = n0 + cc;
        if (cc < E.block_n && c < E.Co && row_ok && !(E.dbg & 2)) {
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c + 4));
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c)), b1 = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
          float f[8];
          f[0] = fmaf(__uint_as_float(v[h * 8 + 0]), s0.x, b0.x);
          f[1] = fmaf(__uint_as_float(v[h * 8 + 1]), s0.y, b0.y);
          f[2] = fmaf(__uint_as_float(v[h * 8 + 2]), s0.z, b0.z);
          f[3] = fmaf(__uint_as_float(v[h * 8 + 3]), s0.w, b0.w);
          f[4] = fmaf(__uint_as_float(v[h * 8 + 4]), s1.x, b1.x);
          f[5] = fmaf(__uint_as_float(v[h * 8 + 5]), s1.y, b1.y);
          f[6] = fmaf(__uint_as_float(v[h * 8 + 6]), s1.z, b1.z);
          f[7] = fmaf(__uint_as_float(v[h * 8 + 7]), s1.w, b1.w);
          if (RES) {
            const __half2* rh = reinterpret_cast<const __half2*>(&rv[(c0 >> 3) + h]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 r2 = __half22float2(rh[q]);
              f[2 * q] += r2.x;
              f[2 * q + 1] += r2.y;
            }
          }
          uint4 ov;
          __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
          for (int q = 0; q < 4; ++q) oh[q] = __floats2half2_rn(act_t<ACT>(f[2 * q]), act_t<ACT>(f[2 * q + 1]));
          if (!(E.dbg & 1)) *reinterpret_cast<uint4*>(yrow + cc) = ov;
        }
      }
    }
  }
}

template <int ACT>
__device__ __forceinline__ void epi_direct_act(const EpiParams& E, const float* scale, const float* bias, uint32_t t_row,
                                               __half* yrow, const __half* rrow, bool row_ok, int n0, uint32_t tfull_bar,
                                               uint32_t tfull_phase, uint32_t tempty_bar, int lane) {
  if (E.block_n <= 32) {
    if (E.has_residual) epi_direct_row<ACT, true, 4>(E, scale, bias, t_row, yrow, rrow, row_ok, n0, tfull_bar, tfull_phase, tempty_bar, lane);
    else epi_direct_row<ACT, false, 4>(E, scale, bias, t_row, yrow, rrow, row_ok, n0, tfull_bar, tfull_phase, tempty_bar, lane);
  } else {
    if (E.has_residual) epi_direct_row<ACT, true, 8>(E, scale, bias, t_row, yrow, rrow, row_ok, n0, tfull_bar, tfull_phase, tempty_bar, lane);
    else epi_direct_row<ACT, false, 8>(E, scale, bias, t_row, yrow, rrow, row_ok, n0, tfull_bar, tfull_phase, tempty_bar, lane);
  }
}

// o0..o3: tile origin in the merged output dims.  Waits on tfull itself.
__device__ __forceinline__ void epilogue_tile_direct(const EpiParams& E, const float* __restrict__ scale,
                                                     const float* __restrict__ bias, uint32_t t_acc, int quarter,
                                                     int lane, int n0, int o0, int o1, int o2, int o3,
                                                     uint32_t tfull_bar, uint32_t tfull_phase, uint32_t tempty_bar) {
  const int row = quarter * 32 + lane;
  bool ok = row < E.rows;
  long long yoff = n0, roff = n0;
  {
    int rr = row;
    const int og[4] = {o0, o1, o2, o3};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int bi = E.box[i];
      const int q = bi > 1 ? rr / bi : rr;
      const int ci = bi > 1 ? rr - q * bi : 0;
      rr = q;
      const int gi = og[i] + ci;
      ok = ok && gi < E.O[i];
      yoff += (long long)gi * E.y_str[i];
      roff += (long long)gi * E.r_str[i];
    }
  }
  const uint32_t t_row = t_acc + ((uint32_t)(quarter * 32) << 16);
  __half* yrow = E.y_ptr + (ok ? yoff : 0);
  const __half* rrow = E.has_residual ? E.r_ptr + (ok ? roff : 0) : nullptr;
  switch (E.act) {
    case PV_ACT_RELU: epi_direct_act<PV_ACT_RELU>(E, scale, bias, t_row, yrow, rrow, ok, n0, tfull_bar, tfull_phase, tempty_bar, lane); break;
    case PV_ACT_NONE: epi_direct_act<PV_ACT_NONE>(E, scale, bias, t_row, yrow, rrow, ok, n0, tfull_bar, tfull_phase, tempty_bar, lane); break;
    case PV_ACT_SWISH: epi_direct_act<PV_ACT_SWISH>(E, scale, bias, t_row, yrow, rrow, ok, n0, tfull_bar, tfull_phase, tempty_bar, lane); break;
    case PV_ACT_GELU: epi_direct_act<PV_ACT_GELU>(E, scale, bias, t_row, yrow, rrow, ok, n0, tfull_bar, tfull_phase, tempty_bar, lane); break;
    default: epi_direct_act<PV_ACT_SIGMOID>(E, scale, bias, t_row, yrow, rrow, ok, n0, tfull_bar, tfull_phase, tempty_bar, lane); break;
  }
}

// Rows of <= 64 bytes (BLOCK_N <= 32) go out directly; wider rows keep the TMA-staged path (a warp-wide
// 16-byte store then touches 32 different 128-byte lines - measured slower than the bulk store at
// BLOCK_N = 64).  PVB200_DEBUG bit 64 flips the choice for A/B measurements.
__host__ __device__ inline bool epi_direct(const EpiParams& E) {
  return epi_narrow(E.block_n) && ((E.block_n <= 32) != ((E.dbg & 64) != 0));
}


// ---------------------------------------------------------------------------------------------
// WIP (round 2, NOT validated on hardware): wide tiles with a residual, double-buffered staging.
// profiles/r01_source_counters.md: on res2 conv_c the MMA warp waits for a free accumulator 68x per
// tile - the epilogue is the bound, and per 128-column group it is a chain of: wait until the previous
// bulk store has read the staging buffer -> residual TMA load (~1.5 us) -> in-place update -> store.
// Here the residual of group q+1 is requested right after the store of group q is issued, into the
// OTHER staging buffer (free once the store of group q-1 has been read: wait_group.read 1), so its
// latency overlaps the store of q, the accumulator wait and the math of the next group.
//   q            : running group counter of this CTA (buffer = q & 1, one residual barrier per buffer)
//   nxt_*        : first group of the NEXT tile of this CTA (nxt_valid = 0 at the last tile)
// The caller issues the very first residual load (group 0 of its first tile) before the tile loop with
// epi_prefetch_residual(...).
// ---------------------------------------------------------------------------------------------
// One 128-column group of this CTA's epilogue sequence: q-th group = group (q % groups_per_tile) of its (q / gpt)-th tile.
struct EpiGroup { int valid, n0, ncols, c1, c2, c3, c4; };
constexpr int EPI_RING = 3;          // staging buffers of the residual ring; residual loads run 2 groups ahead

__device__ __forceinline__ void epi_prefetch_residual(const EpiParams& E, uint32_t epi_smem, uint32_t res_bar, int buf,
                                                      const EpiGroup& G) {
  const int nsub = (G.ncols + 63) >> 6;
  const uint32_t bar = res_bar + 8u * (uint32_t)buf;
  mbar_arrive_expect_tx(bar, (uint32_t)(nsub * E.rows * 128));
  for (int s = 0; s < nsub; ++s)
    tma_load_5d(epi_smem + (uint32_t)buf * EPI_STAGING_BYTES + (uint32_t)s * 16384u, &E.r_map, bar, G.n0 + s * 64, G.c1, G.c2, G.c3, G.c4);
}

// Wide residual tiles (conv_c + shortcut add + ReLU of every bottleneck): these layers have a short K loop and are
// bound by their epilogue - residual tile in, output tile out.  profiles/r01_source_counters.md: the MMA warp waits
// for a free accumulator 68x per tile.  The staging area is a ring of EPI_RING buffers of one 128-column group each:
//   residual(q + 2) is requested right after store(q) is issued, into the buffer group q - 1 used (free once the
//   store of q - 1 has been READ, i.e. all bulk groups but the newest: cp.async.bulk.wait_group.read 1),
// so a residual load has a whole group period (accumulator wait + tcgen05.ld + math of group q + 1) to land and a
// store drains while the next group is computed.  group_at(q) maps the running group counter to its coordinates.
template <typename GroupAt>
__device__ __forceinline__ void epilogue_tile_ring(const EpiParams& E, const float* __restrict__ scale,
                                                   const float* __restrict__ bias, uint32_t t_acc, uint32_t epi_smem,
                                                   uint8_t* epi_gen, uint32_t res_bar, uint32_t (&res_phase)[EPI_RING],
                                                   int& q, int ewarp, int quarter, int lane, int n_tile0,
                                                   uint32_t tempty_bar, GroupAt group_at) {
  constexpr int R = EPI_RING;
  constexpr bool RES = true;
  const int row = quarter * 32 + lane;
  const int chalf = ewarp >> 2;
  const int etid = ewarp * 32 + lane;
  const bool leader = (etid == 0);
  const uint32_t t_row = t_acc + ((uint32_t)(quarter * 32) << 16);
  const uint32_t rsw = (uint32_t)(row & 7);
  float* sb = reinterpret_cast<float*>(epi_gen + R * EPI_STAGING_BYTES);     // scale/bias live after the ring
  for (int i = etid; i < E.block_n; i += EPI_THREADS) {
    const int c = n_tile0 + i;
    const bool ok = c < E.Co;
    sb[i] = ok ? __ldg(scale + c) : 0.f;
    sb[256 + i] = ok ? __ldg(bias + c) : 0.f;
  }
  for (int g0 = 0; g0 < E.block_n; g0 += EPI_GROUP_COLS, ++q) {
    const int buf = q % R;
    const EpiGroup G = group_at(q);
    const int nsub = (G.ncols + 63) >> 6;
    const uint32_t slot = (uint32_t)buf * EPI_STAGING_BYTES;
    epi_bar_sync(1, EPI_THREADS);                       // scale/bias visible; previous group fully staged; buffer free (leader waited)
    mbar_wait(res_bar + 8u * (uint32_t)buf, res_phase[buf]);   // residual of THIS group (requested two groups ago)
    res_phase[buf] ^= 1u;
    if (chalf < nsub && !(E.dbg & 2)) {
      const int cbase = g0 + chalf * 64;
      const int ncols = min(64, G.ncols - chalf * 64);
      uint8_t* srow = epi_gen + slot + chalf * 16384 + row * 128;
      const float* sc = sb + cbase;
      const float* bi = sb + 256 + cbase;
      switch (E.act) {
        case PV_ACT_RELU: epi_subtile<PV_ACT_RELU, RES>(t_row + (uint32_t)cbase, srow, rsw, sc, bi, ncols); break;
        case PV_ACT_NONE: epi_subtile<PV_ACT_NONE, RES>(t_row + (uint32_t)cbase, srow, rsw, sc, bi, ncols); break;
        case PV_ACT_SWISH: epi_subtile<PV_ACT_SWISH, RES>(t_row + (uint32_t)cbase, srow, rsw, sc, bi, ncols); break;
        case PV_ACT_GELU: epi_subtile<PV_ACT_GELU, RES>(t_row + (uint32_t)cbase, srow, rsw, sc, bi, ncols); break;
        default: epi_subtile<PV_ACT_SIGMOID, RES>(t_row + (uint32_t)cbase, srow, rsw, sc, bi, ncols); break;
      }
    }
    if (g0 + EPI_GROUP_COLS >= E.block_n) {             // accumulator fully read: hand TMEM back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar);   // shared::cluster address: own CTA, or the pair's leader
    }
    fence_proxy_async_smem();
    epi_bar_sync(1, EPI_THREADS);
    if (leader) {
      if (!(E.dbg & 1)) {
        for (int s = 0; s < nsub; ++s)
          tma_store_5d(&E.y_map, epi_smem + slot + (uint32_t)s * 16384u, G.n0 + s * 64, G.c1, G.c2, G.c3, G.c4);
      }
      tma_store_commit();
      const EpiGroup N2 = group_at(q + 2);              // its buffer was last used by group q - 1
      if (N2.valid) {
        tma_store_wait_read1();                         // every store but the one just issued has read its buffer
        epi_prefetch_residual(E, epi_smem, res_bar, (q + 2) % EPI_RING, N2);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Two epilogue teams for wide residual tiles.  tools/epi_sweep.py (profiles/r02_epilogue_sweep.md): a conv_c + residual layer is a
// latency chain per staging group - residual landed -> TMEM -> registers -> staging -> proxy fence -> barrier -> bulk store -> wait
// for the previous store's read -> next residual request - and with eight warps on ONE tile only one such chain is in flight per SM.
// TMEM holds two accumulators (BLOCK_N <= 256), so here the eight warps split into two teams of four (each team owns all four TMEM
// lane quarters) that drain ALTERNATE tiles concurrently: a group is one 64-column sub-tile (16 KiB), each team has its own ring of
// EPI_RING 16-KiB buffers (together the same 96 KiB as the single ring), its own scale / bias copy, residual barriers, named barrier
// and issuing thread (bulk-group queues are per thread).  Same shape as the narrow-tile path of epilogue_tile, plus the residual ring.
//   team_q      : running group counter of this TEAM (buffer = team_q % EPI_RING)
//   group_at(q) : the team's q-th group -> (tile coordinates, first column, columns); groups per tile = ceil(BLOCK_N / 64)
// ---------------------------------------------------------------------------------------------
constexpr int EPI_TEAM_BUF_BYTES = 16384;                          // one [128 rows x 64 ch] swizzled sub-tile
constexpr int EPI_TEAM_RING_BYTES = EPI_RING * EPI_TEAM_BUF_BYTES;

__device__ __forceinline__ void epi_team_prefetch_residual(const EpiParams& E, uint32_t team_smem, uint32_t team_res_bar, int buf,
                                                           const EpiGroup& G) {
  const uint32_t bar = team_res_bar + 8u * (uint32_t)buf;
  mbar_arrive_expect_tx(bar, (uint32_t)(E.rows * 128));
  tma_load_5d(team_smem + (uint32_t)buf * EPI_TEAM_BUF_BYTES, &E.r_map, bar, G.n0, G.c1, G.c2, G.c3, G.c4);
}

template <typename GroupAt>
__device__ __forceinline__ void epilogue_tile_ring_teams(const EpiParams& E, const float* __restrict__ scale,
                                                         const float* __restrict__ bias, uint32_t t_acc, uint32_t epi_smem,
                                                         uint8_t* epi_gen, uint32_t res_bar, uint32_t (&res_phase)[EPI_RING],
                                                         int& team_q, int ewarp, int quarter, int lane, int n_tile0,
                                                         uint32_t tempty_bar, GroupAt group_at) {
  const int team = ewarp >> 2;
  const int row = quarter * 32 + lane;
  const int etid = (ewarp & 3) * 32 + lane;             // 0..127 inside the team
  const bool leader = (etid == 0);
  const int bar_id = 1 + team;
  const uint32_t t_row = t_acc + ((uint32_t)(quarter * 32) << 16);
  const uint32_t rsw = (uint32_t)(row & 7);
  const uint32_t team_smem = epi_smem + (uint32_t)team * EPI_TEAM_RING_BYTES;
  uint8_t* team_gen = epi_gen + team * EPI_TEAM_RING_BYTES;
  const uint32_t team_res_bar = res_bar + 8u * (uint32_t)(team * EPI_RING);
  float* sb = reinterpret_cast<float*>(epi_gen + 2 * EPI_TEAM_RING_BYTES) + team * 512;   // scale / bias live after both rings
  for (int i = etid; i < E.block_n; i += 128) {
    const int c = n_tile0 + i;
    const bool ok = c < E.Co;
    sb[i] = ok ? __ldg(scale + c) : 0.f;
    sb[256 + i] = ok ? __ldg(bias + c) : 0.f;
  }
  for (int g0 = 0; g0 < E.block_n; g0 += 64, ++team_q) {
    const int buf = team_q % EPI_RING;
    const EpiGroup G = group_at(team_q);
    epi_bar_sync(bar_id, 128);                          // scale/bias visible; previous group of this team fully staged
    mbar_wait(team_res_bar + 8u * (uint32_t)buf, res_phase[buf]);      // residual of THIS group (requested two groups ago)
    res_phase[buf] ^= 1u;
    if (!(E.dbg & 2)) {
      uint8_t* srow = team_gen + buf * EPI_TEAM_BUF_BYTES + row * 128;
      const float* sc = sb + g0;
      const float* bi = sb + 256 + g0;
      switch (E.act) {
        case PV_ACT_RELU: epi_subtile<PV_ACT_RELU, true>(t_row + (uint32_t)g0, srow, rsw, sc, bi, G.ncols); break;
        case PV_ACT_NONE: epi_subtile<PV_ACT_NONE, true>(t_row + (uint32_t)g0, srow, rsw, sc, bi, G.ncols); break;
        case PV_ACT_SWISH: epi_subtile<PV_ACT_SWISH, true>(t_row + (uint32_t)g0, srow, rsw, sc, bi, G.ncols); break;
        case PV_ACT_GELU: epi_subtile<PV_ACT_GELU, true>(t_row + (uint32_t)g0, srow, rsw, sc, bi, G.ncols); break;
        default: epi_subtile<PV_ACT_SIGMOID, true>(t_row + (uint32_t)g0, srow, rsw, sc, bi, G.ncols); break;
      }
    }
    if (g0 + 64 >= E.block_n) {                         // accumulator fully read: hand TMEM back to the MMA warp (4 arrivals)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar);
    }
    fence_proxy_async_smem();
    epi_bar_sync(bar_id, 128);
    if (leader) {
      if (!(E.dbg & 1)) tma_store_5d(&E.y_map, team_smem + (uint32_t)buf * EPI_TEAM_BUF_BYTES, G.n0, G.c1, G.c2, G.c3, G.c4);
      tma_store_commit();
      const EpiGroup N2 = group_at(team_q + 2);         // its buffer was last used by group team_q - 1
      if (N2.valid) {
        tma_store_wait_read1();                         // every store of this thread but the newest has read its buffer
        epi_team_prefetch_residual(E, team_smem, team_res_bar, (team_q + 2) % EPI_RING, N2);
      }
    }
  }
}

__host__ __device__ inline bool epi_wide_teams(const EpiParams& E) {
  return !epi_narrow(E.block_n) && E.has_residual && E.teams == 2 && !(E.dbg & 512);
}

__host__ __device__ inline bool epi_wide_prefetch(const EpiParams& E) {
  return !epi_narrow(E.block_n) && E.has_residual && !(E.dbg & 512);     // PVB200_DEBUG=512: fall back to the single-buffer epilogue
}

}  // namespace sm100
}  // namespace pv
