// Fused epilogue shared by the TMA-fed and the gather-fed implicit-GEMM kernels.
//
// TMEM accumulator (128 rows x BLOCK_N fp32) -> registers (tcgen05.ld) -> y = act(acc*scale+bias
// (+residual)) -> f16 -> 128B-swizzled shared staging -> TMA tensor store.  The residual tile is
// brought in by a TMA tensor load into the SAME staging buffer and updated in place, so both the
// residual read and the output write are full-line bulk transfers issued by one thread, with the
// tile-edge and channel-tail clipping done by the TMA unit (no per-thread masks or address math).
// Columns are processed in groups of <=128 (two 64-column sub-tiles, 32 KiB of staging).
#pragma once
#include "pv_common.cuh"
#include "pv_sm100.cuh"

namespace pv {
namespace sm100 {

constexpr int EPI_GROUP_COLS = 128;
constexpr int EPI_STAGING_BYTES = 2 * 128 * 128;   // two [128 rows x 64 f16] swizzled sub-tiles
constexpr int EPI_THREADS = 128;

struct EpiParams {
  CUtensorMap y_map;    // output  [Co, d1, d2, d3, d4], box [64, b1, b2, b3, b4], SWIZZLE_128B
  CUtensorMap r_map;    // residual, same geometry
  int block_n, Co, rows, act, has_residual;
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
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Called by the 4 epilogue warps (128 threads) for one output tile.
//   t_acc     : TMEM address of the accumulator (column base), lane field NOT yet applied
//   staging   : shared address (1024-aligned) / generic pointer of the 32 KiB staging buffer
//   res_bar   : mbarrier for the residual TMA loads, res_phase its running parity
//   c1..c4    : tile origin in the output tensor map's spatial dims;  n0: first output channel
//   tempty_bar: arrived (one lane per warp) as soon as the accumulator has been drained
__device__ __forceinline__ void epilogue_tile(const EpiParams& E, const float* __restrict__ scale,
                                              const float* __restrict__ bias, uint32_t t_acc,
                                              uint32_t staging, uint8_t* staging_gen, uint32_t res_bar,
                                              uint32_t& res_phase, int quarter, int lane, int n0, int c1,
                                              int c2, int c3, int c4, uint32_t tempty_bar) {
  const int row = quarter * 32 + lane;
  const bool leader = (quarter == 0) && (lane == 0);
  const uint32_t t_row = t_acc + ((uint32_t)(quarter * 32) << 16);
  const uint32_t rsw = (uint32_t)(row & 7);
  for (int g0 = 0; g0 < E.block_n; g0 += EPI_GROUP_COLS) {
    const int gcols = min(EPI_GROUP_COLS, E.block_n - g0);
    const int nsub = (gcols + 63) >> 6;
    // (a) staging is free once the previous group's stores have been read out of shared memory
    if (leader) {
      tma_store_wait_read0();
      if (E.has_residual) {
        mbar_arrive_expect_tx(res_bar, (uint32_t)(nsub * E.rows * 128));
        for (int s = 0; s < nsub; ++s)
          tma_load_5d(staging + (uint32_t)s * 16384u, &E.r_map, res_bar, n0 + g0 + s * 64, c1, c2, c3, c4);
      }
    }
    if (E.has_residual) {
      mbar_wait(res_bar, res_phase);
      res_phase ^= 1u;
    } else {
      epi_bar_sync();
    }
    // (b) drain the accumulator, 16 columns at a time
    for (int c0 = 0; c0 < gcols; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(t_row + (uint32_t)(g0 + c0), v);
      tmem_ld_wait();
      const int sub = c0 >> 6;
      uint8_t* srow = staging_gen + sub * 16384 + row * 128;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col = n0 + g0 + c0 + h * 8;          // global output channel of the first of 8 lanes
        const uint32_t j = (uint32_t)(((c0 & 63) >> 3) + h);   // 16B chunk inside the 128B row
        uint4* cell = reinterpret_cast<uint4*>(srow + ((j ^ rsw) << 4));
        float f[8];
        if (col < E.Co) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            f[q] = __uint_as_float(v[h * 8 + q]) * __ldg(scale + col + q) + __ldg(bias + col + q);
          if (E.has_residual) {
            const uint4 rv = *cell;
            const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 r2 = __half22float2(rh[q]);
              f[2 * q] += r2.x;
              f[2 * q + 1] += r2.y;
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) f[q] = apply_act(f[q], E.act);
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) f[q] = 0.f;
        }
        uint4 ov;
        __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
        for (int q = 0; q < 4; ++q) oh[q] = __floats2half2_rn(f[2 * q], f[2 * q + 1]);
        *cell = ov;
      }
    }
    if (g0 + EPI_GROUP_COLS >= E.block_n) {   // accumulator fully read: hand TMEM back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
    }
    // (c) publish the staged tile to the async proxy and store it
    fence_proxy_async_smem();
    epi_bar_sync();
    if (leader) {
      for (int s = 0; s < nsub; ++s)
        tma_store_5d(&E.y_map, staging + (uint32_t)s * 16384u, n0 + g0 + s * 64, c1, c2, c3, c4);
      tma_store_commit();
    }
  }
}

}  // namespace sm100
}  // namespace pv
