// Implicit-GEMM 3-D convolution on tcgen05 tensor cores (sm_100a).
//
//   D[m][co] = sum_{tap,ci} X[pos(m) + tap][ci] * Wt[co][tap][ci]        (fp32 accumulate in TMEM)
//   y        = act(D * scale[co] + bias[co] (+ residual))               (fused epilogue)
//
// * Activations are NDHWC f16.  The GEMM-M tile (128 rows) is a BOX of output positions
//   (b0 x b1 x b2 x b3 over the W/H/T/N-like dims, merged where the kernel is trivial), so for
//   every filter tap the A operand is ONE TMA tiled load of box [64 ch, b0, b1, b2, b3] at
//   shifted coordinates; TMA's out-of-bounds zero fill implements the convolution padding and
//   the channel tail (Ci % 64 != 0).  Strided convolutions use one tensor map per stride-parity
//   class (base pointer offset + multiplied global strides), so only documented tiled-mode
//   features are relied on.
// * B operand (packed weights, K-major rows of taps*ci_pad64) is a 2-D TMA load [64, BLOCK_N].
// * Both land in 128B-swizzled shared memory and feed tcgen05.mma (M=128, N=BLOCK_N, K=16) via
//   shared-memory descriptors; accumulators live in TMEM (double buffered) and are drained by
//   four epilogue warps with tcgen05.ld.
// * Warp-specialised persistent kernel: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc),
//   warps2-5 = epilogue; smem ring of `stages` {A,B} slots with full/empty mbarriers.
#include "pv_common.cuh"
#include "pv_sm100.cuh"
#include "pv_epilogue.cuh"

#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace pv {

using namespace sm100;

constexpr int IG_BM = 128;        // UMMA M
constexpr int IG_MAX_TAPS = 64;
constexpr int IG_MAX_MAPS = 8;
constexpr int IG_PROD_WARPS = 4;   // TMA producer warps (one elected lane each, k-blocks round-robin)
constexpr int IG_MMA_WARP = IG_PROD_WARPS;
constexpr int IG_EPI_WARP0 = IG_PROD_WARPS + 1;
constexpr int IG_THREADS = (IG_PROD_WARPS + 1 + EPI_WARPS) * 32;   // 416

struct IgemmParams {
  CUtensorMap a_maps[IG_MAX_MAPS];
  CUtensorMap b_map;
  // tiling over the 4 merged output dims (0 = innermost)
  int O[4];        // output extents
  int box[4];      // tile box
  int nt[4];       // tiles per dim
  int rows;        // box[0]*box[1]*box[2]*box[3] (<= 128)
  int n_tiles, m_tiles;
  int block_n;
  int Co;
  int taps, num_kc;
  int kbytes;      // bytes of K per smem row and pipeline stage: 128 (64 ch, SW128) | 64 | 32 (window mode)
  int stages;
  int tmem_cols;
  int acc_stride;   // TMEM columns between accumulator stages (block_n rounded up to 32)
  int nacc;         // number of accumulator stages
  int G, cpt;       // k-blocks per pipeline stage (chunk), chunks per tile
  int split_ab;     // producer warps that share the loads of a chunk (1, 2 or 4; PVB200_SPLIT_AB)
  int epi_bytes;    // epilogue shared memory (one or - residual prefetch - two staging buffers + scale/bias)
  int alias_epi;    // 1: every CTA runs exactly ONE tile, so the epilogue staging may reuse the (then idle) tile ring:
                    //    one more pipeline stage for the layers with fewer tiles than SMs (res4 / res5)
  int pair;         // 1: CTA pair (cluster of 2, tcgen05 cta_group::2, M = 256 per MMA), see pv_sm100.cuh
  int pair_tiles;   // pair mode: n_tiles * ceil(m_tiles / 2) units of work (one per cluster and iteration)
  EpiParams epi;
  signed char tap_q[IG_MAX_TAPS][4];
  unsigned char tap_map[IG_MAX_TAPS];
};

// PAIR is a template parameter, not a run-time flag: a cubin that contains cta_group::2 instructions can only be
// launched as a cluster of two ("cluster misconfiguration" otherwise), so the single-CTA kernel must not carry them.
template <bool PAIR>
__global__ void __launch_bounds__(IG_THREADS, 1)
conv3d_igemm_kernel(const __grid_constant__ IgemmParams P, const float* __restrict__ scale,
                    const float* __restrict__ bias) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stages = P.stages;
  constexpr bool pair = PAIR;
  const uint32_t cta_rank = pair ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int b_rows = pair ? (P.block_n >> 1) : P.block_n;       // B rows THIS CTA keeps in shared memory
  const uint32_t a_bytes = (uint32_t)IG_BM * P.kbytes;
  const uint32_t b_bytes = (uint32_t)b_rows * P.kbytes;
  const int G = P.G;
  const uint32_t stage_bytes = (uint32_t)G * (a_bytes + b_bytes);   // [G x A k-block][G x B k-block]
  const int k_elems = P.kbytes >> 1;
  // epilogue staging (1024-aligned) and the barriers live after the tile ring
  const uint32_t ring_end = (uint32_t)((stages * stage_bytes + 1023u) & ~1023u);
  const uint32_t staging_off = P.alias_epi ? 0u : ring_end;
  const uint32_t staging = smem_base + staging_off;
  const uint32_t bar_base = smem_base + max(ring_end, staging_off + (uint32_t)P.epi_bytes);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * stages + s); };
  const int nacc = P.nacc;                   // accumulator stages in TMEM (2..8)
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * stages + nacc + s); };
  const uint32_t res_bar = bar_base + 8u * (2 * stages + 2 * nacc);
  const uint32_t tmem_slot = bar_base + 8u * (2 * stages + 2 * nacc + 2 * EPI_RING);     // res_bar: one barrier per ring buffer (and team)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&P.b_map);
    prefetch_tmap(&P.a_maps[0]);
    for (int s = 0; s < stages; ++s) {
      // one expect_tx arrive per issuing producer warp; pair mode: the leader's barrier counts the leader's
      // expect_tx arrive plus the peer's plain (remote) arrive, and the bytes of BOTH CTAs' loads
      mbar_init(full_bar(s), pair ? 2 : 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < nacc; ++s) {
      mbar_init(tfull_bar(s), 1);
      // one arrive per epilogue warp of the tile's group (pair mode: from both CTAs, on the leader's barrier)
      mbar_init(tempty_bar(s), ((epi_narrow(P.block_n) || epi_wide_teams(P.epi)) ? 4 : EPI_WARPS) * (pair ? 2 : 1));
    }
    for (int s = 0; s < 2 * EPI_RING; ++s) mbar_init(res_bar + 8u * s, 1);
    prefetch_tmap(&P.epi.y_map);
    fence_mbar_init();
  }
  if (warp == IG_MMA_WARP) {
    if (pair) {       // both CTAs, same warp id, same shared-memory slot
      tmem_alloc_pair(tmem_slot, (uint32_t)P.tmem_cols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (pair) cluster_sync_all();     // the peer's barriers must be initialised before anything arrives on them
  else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch)
  // overlaps the tail of the previous kernel in the stream / graph; its results are only touched below.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // Work units: a tile (n_tile, m_tile) per CTA, or - pair mode - two M-adjacent tiles (2*mp + rank) per cluster.
  const int total_tiles = pair ? P.pair_tiles : P.n_tiles * P.m_tiles;
  const int unit0 = pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_step = pair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_kb = P.taps * P.num_kc;
  // unit -> (n0 tile index, merged output coordinates of this CTA's 128-row tile); a missing second tile of the
  // last pair gets coordinates outside the tensor: its loads are zero-filled and its stores clipped by TMA
  auto tile_coords = [&](int unit, int& n_tile, int (&o)[4]) {
    n_tile = unit % P.n_tiles;
    int mt = unit / P.n_tiles;
    if (pair) mt = 2 * mt + (int)cta_rank;
    const bool valid = mt < P.m_tiles;
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[i] = (mt % P.nt[i]) * P.box[i]; mt /= P.nt[i]; }
    if (!valid) o[3] = P.O[3] + 128;
  };

  if (warp < IG_PROD_WARPS) {
    // ================================ TMA producers =========================================
    // One producer warp: warp-uniform loop, one elected lane issues the TMA loads of a whole chunk
    // (up to G k-blocks = 2G bulk-tensor loads on ONE mbarrier).  Measured on B200: extra producer warps
    // do not help; what bounds narrow-N layers is the number of barrier rounds, hence the chunking.
    // The loads of a chunk (2 per k-block: A and B) are dealt round-robin to `nsplit` producer warps - one issuing
    // thread sustains only ~one bulk-tensor load per 330-520 clk (tools/probe/tma_rate.cu), four warps one per ~130.
    // Warp 0 alone arrives on the full barrier with the expected bytes of the WHOLE chunk; the other warps' loads may
    // complete before that arrive (the transaction count goes negative, the phase cannot complete without the arrive).
    // (the load-skipping probe completes the full barrier from warp 0 alone: the other producer warps must not run, or they can fall
    //  whole laps behind and wait for a phase that never comes once the tiles are exhausted)
    const int nsplit = (pair || (P.epi.dbg & 4)) ? 1 : P.split_ab;
    if (warp < nsplit) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = (uint32_t)P.rows * P.kbytes + b_bytes;
      const int num_kc = P.num_kc, cpt = P.cpt;
      const bool skip_loads = (P.epi.dbg & 4) != 0;
      for (int tile = unit0; tile < total_tiles; tile += unit_step) {
        int n_tile, o[4];
        tile_coords(tile, n_tile, o);
        const int n0 = n_tile * P.block_n + (int)cta_rank * b_rows;     // pair: this CTA's half of the B rows
        int tap = 0, kc = 0;                  // (tap, channel chunk) of the next k-block, stepped without a divide
        for (int ch = 0; ch < cpt; ++ch) {
          const int kb0 = ch * G;
          const int nsub = min(G, num_kb - kb0);
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t st_base = smem_base + (uint32_t)stage * stage_bytes;
          if (elect_one()) {
            if (pair) {
              // data -> own shared memory, transaction bytes -> the LEADER's full barrier
              const uint32_t fb = mapa_shared(full_bar(stage), 0);
              if (leader) mbar_arrive_expect_tx(full_bar(stage), 2u * (uint32_t)nsub * tx_bytes);
              else mbar_arrive_cluster(fb);
              int tp = tap, kk = kc;
              for (int j = 0; j < nsub; ++j) {
                const void* amap = &P.a_maps[P.tap_map[tp]];
                tma_load_5d_pair(st_base + (uint32_t)j * a_bytes, amap, fb, kk * k_elems, o[0] + P.tap_q[tp][0],
                                 o[1] + P.tap_q[tp][1], o[2] + P.tap_q[tp][2], o[3] + P.tap_q[tp][3]);
                tma_load_2d_pair(st_base + (uint32_t)G * a_bytes + (uint32_t)j * b_bytes, &P.b_map, fb,
                                 (kb0 + j) * k_elems, n0);
                if (++kk == num_kc) { kk = 0; ++tp; }
              }
            } else if (skip_loads) {               // probe: pipeline skeleton without the loads
              if (warp == 0) mbar_arrive(full_bar(stage));
            } else {
              if (warp == 0) mbar_arrive_expect_tx(full_bar(stage), (uint32_t)nsub * tx_bytes);
              int tp = tap, kk = kc;
              for (int j = 0; j < nsub; ++j) {
                if ((2 * j) % nsplit == warp) {
                  const void* amap = &P.a_maps[P.tap_map[tp]];
                  tma_load_5d(st_base + (uint32_t)j * a_bytes, amap, full_bar(stage), kk * k_elems, o[0] + P.tap_q[tp][0],
                              o[1] + P.tap_q[tp][1], o[2] + P.tap_q[tp][2], o[3] + P.tap_q[tp][3]);
                }
                if ((2 * j + 1) % nsplit == warp)
                  tma_load_2d(st_base + (uint32_t)G * a_bytes + (uint32_t)j * b_bytes, &P.b_map, full_bar(stage),
                              (kb0 + j) * k_elems, n0);
                if (++kk == num_kc) { kk = 0; ++tp; }
              }
            }
          }
          __syncwarp();
          for (int j = 0; j < nsub; ++j) { if (++kc == num_kc) { kc = 0; ++tap; } }
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == IG_MMA_WARP) {
    // ================================ MMA issuer ============================================
    // The whole warp runs the loop with warp-uniform control flow and waits on the barriers; one
    // elected lane issues the tcgen05 instructions, so their operands stay in uniform registers
    // (inside `if (lane == 0)` every tcgen05.mma / commit became an R2UR + ELECT/BRA.U.ANY waterfall
    // and the issuing thread cost ~800 clk per k-block - the bound of every narrow layer).
    // Pair mode: only the leader CTA issues (M = 256 over both CTAs' A rows / accumulator lanes); its
    // commits arrive on the barriers of BOTH CTAs.
    if (leader) {
      const uint32_t idesc = make_idesc_f16(pair ? 2 * IG_BM : IG_BM, P.block_n);
      const int k16 = (P.epi.dbg & 32) ? 0 : (P.kbytes >> 5);
      const int kbytes = P.kbytes, acc_stride = P.acc_stride;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const int cpt = P.cpt;
      for (int tile = unit0; tile < total_tiles; tile += unit_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_stride);
        for (int ch = 0; ch < cpt; ++ch) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t st_base = smem_base + (uint32_t)stage * stage_bytes;
          const int kb0 = ch * G;
          const int nsub = min(G, num_kb - kb0);
          if (elect_one()) {
            for (int j = 0; j < nsub; ++j) {
              const uint64_t a_desc = make_kmajor_desc(st_base + (uint32_t)j * a_bytes, kbytes);
              const uint64_t b_desc = make_kmajor_desc(st_base + (uint32_t)G * a_bytes + (uint32_t)j * b_bytes, kbytes);
              for (int k = 0; k < k16; ++k) {
                // advance 16 elements (32 B) along K inside the swizzle row: +2 in (addr>>4)
                const uint32_t accum = ((kb0 + j) | k) != 0 ? 1u : 0u;
                if (pair) umma_f16_pair(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, accum);
                else umma_f16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, accum);
              }
            }
            if (pair) {
              umma_commit_pair(empty_bar(stage));
              if (ch == cpt - 1) umma_commit_pair(tfull_bar(acc));
            } else {
              umma_commit(empty_bar(stage));                    // frees the chunk's smem when its MMAs retire
              if (ch == cpt - 1) umma_commit(tfull_bar(acc));   // accumulator complete
            }
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == nacc) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ================================ epilogue warps ========================================
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int ewarp = warp - IG_EPI_WARP0;
    int tile_seq = 0;
    uint32_t res_phase = 0;
    const bool narrow = epi_narrow(P.block_n);
    // the accumulator is handed back on the MMA issuer's barrier: the leader CTA's (a shared::cluster address)
    auto tempty_addr = [&](int a) { return pair ? mapa_shared(tempty_bar(a), 0) : tempty_bar(a); };
    // wide residual tiles: ring of staging buffers, residual loads two groups ahead (epilogue_tile_ring)
    const bool wide_prefetch = epi_wide_prefetch(P.epi);
    uint32_t res_phase3[EPI_RING];
#pragma unroll
    for (int s = 0; s < EPI_RING; ++s) res_phase3[s] = 0u;
    int q = 0;
    const int gpt = (P.block_n + EPI_GROUP_COLS - 1) / EPI_GROUP_COLS;      // groups per tile
    auto group_at = [&](int qq) {
      EpiGroup G;
      const int ts = qq / gpt, gi = qq - ts * gpt;
      const int unit = unit0 + ts * unit_step;
      G.valid = unit < total_tiles ? 1 : 0;
      int nt = 0, oc[4] = {0, 0, 0, 0};
      if (G.valid) tile_coords(unit, nt, oc);
      G.n0 = nt * P.block_n + gi * EPI_GROUP_COLS;
      G.ncols = min(EPI_GROUP_COLS, P.block_n - gi * EPI_GROUP_COLS);
      G.c1 = oc[0]; G.c2 = oc[1]; G.c3 = oc[2]; G.c4 = oc[3];
      return G;
    };
    const bool wide_teams = epi_wide_teams(P.epi);
    const int team = ewarp >> 2;
    const int gpt_team = (P.block_n + 63) >> 6;                             // 64-column groups per tile (team mode)
    auto team_group_at = [&](int qq) {                                      // the team's qq-th group: tile 2 * (qq / gpt) + team of this CTA
      EpiGroup G;
      const int ts = qq / gpt_team, gi = qq - ts * gpt_team;
      const int unit = unit0 + (2 * ts + team) * unit_step;
      G.valid = unit < total_tiles ? 1 : 0;
      int nt = 0, oc[4] = {0, 0, 0, 0};
      if (G.valid) tile_coords(unit, nt, oc);
      G.n0 = nt * P.block_n + gi * 64;
      G.ncols = min(64, P.block_n - gi * 64);
      G.c1 = oc[0]; G.c2 = oc[1]; G.c3 = oc[2]; G.c4 = oc[3];
      return G;
    };
    if (wide_teams && (ewarp & 3) == 0 && lane == 0) {                     // each team leader: residuals of its first two groups
      for (int qq = 0; qq < 2; ++qq) {
        const EpiGroup G0 = team_group_at(qq);
        if (G0.valid)
          epi_team_prefetch_residual(P.epi, staging + (uint32_t)team * EPI_TEAM_RING_BYTES, res_bar + 8u * (uint32_t)(team * EPI_RING),
                                     qq % EPI_RING, G0);
      }
    }
    if (!wide_teams && wide_prefetch && ewarp == 0 && lane == 0) {
      for (int qq = 0; qq < 2; ++qq) {
        const EpiGroup G0 = group_at(qq);
        if (G0.valid) epi_prefetch_residual(P.epi, staging, res_bar, qq % EPI_RING, G0);
      }
    }
    for (int tile = unit0; tile < total_tiles; tile += unit_step, ++tile_seq) {
      if ((narrow || wide_teams) && (tile_seq & 1) != (ewarp >> 2)) continue;     // the other group's / team's tile
      const int acc = tile_seq % nacc;
      const uint32_t acc_phase = (uint32_t)((tile_seq / nacc) & 1);
      int n_tile, o[4];
      tile_coords(tile, n_tile, o);
      if (epi_direct(P.epi)) {
        epilogue_tile_direct(P.epi, scale, bias, tmem_base + (uint32_t)(acc * P.acc_stride), quarter, lane,
                             n_tile * P.block_n, o[0], o[1], o[2], o[3], tfull_bar(acc), acc_phase, tempty_addr(acc));
        continue;
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (wide_teams) {
        epilogue_tile_ring_teams(P.epi, scale, bias, tmem_base + (uint32_t)(acc * P.acc_stride), staging, smem_gen + staging_off,
                                 res_bar, res_phase3, q, ewarp, quarter, lane, n_tile * P.block_n, tempty_addr(acc), team_group_at);
        continue;
      }
      if (wide_prefetch) {
        epilogue_tile_ring(P.epi, scale, bias, tmem_base + (uint32_t)(acc * P.acc_stride), staging, smem_gen + staging_off,
                           res_bar, res_phase3, q, ewarp, quarter, lane, n_tile * P.block_n, tempty_addr(acc), group_at);
        continue;
      }
      epilogue_tile(P.epi, scale, bias, tmem_base + (uint32_t)(acc * P.acc_stride), staging, smem_gen + staging_off,
                    res_bar, res_phase, ewarp, quarter, lane, n_tile * P.block_n, o[0], o[1], o[2], o[3],
                    tempty_addr(acc), tile_seq);
    }
    if ((ewarp & 3) == 0 && lane == 0) tma_store_wait_all();   // smem must outlive the bulk stores (both group leaders)
  }

  tc_fence_before();
  if (pair) cluster_sync_all();      // the peer may still be reading this CTA's shared memory / signalling its barriers
  else __syncthreads();
  if (warp == IG_MMA_WARP) {
    tc_fence_after();
    if (pair) tmem_dealloc_pair(tmem_base, (uint32_t)P.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  }
}

// =============================================================================================
// Host side: problem reduction, tile search, tensor maps, launch
// =============================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

struct Dim4 {      // one merged spatial dim
  long long I, O;  // input / output extent
  int k, s, p, dil;
  long long stride_bytes;   // byte distance of +1 input index in this dim
  bool nomerge;
};

struct IgemmPlan {
  Dim4 dim[4];
  int ndims;
  int orig2m[4];   // original dim (0=W,1=H,2=T,3=N) -> merged dim index
};

// Window mode (narrow C_in stems): the K axis of one pipeline stage is the contiguous run of
// kw*Ci input elements that one output pixel reads in one (dt,dh) filter row.  The TMA tensor map
// uses OVERLAPPING strides - dim0 = the run (padded to 16/32/64 elements), dim1 = output column
// with byte stride sw*Ci*2 - so one tiled load delivers 128 sliding windows.  The input rows carry
// x_w_pad physical zero pixels on the left (and enough on the right), written by the layout
// conversion, which implements the W padding; H/T padding is TMA out-of-bounds fill as usual.
static bool window_mode(const pv_conv3d_desc* d) { return d->x_w_pad > 0 && d->Ci <= 8; }
// Narrow TMA mode: C_in = 16 / 32 with the weights packed at the un-padded per-tap K extent.  One tap is a
// 32 / 64-byte box row (SWIZZLE_32B / 64B), one k-block per tap, several taps per pipeline stage - fewer
// barrier rounds and no SM instructions for the A operand, which beats the cp.async gather for these
// widths (the gather kernel keeps C_in = 4 / 8 / 24 / 40 / 48 / 56).
bool conv3d_tma_narrow(const pv_conv3d_desc* d) {
  static const bool off = getenv("PVB200_GATHER_ALL") != nullptr;
  return !off && !window_mode(d) && (d->Ci == 16 || d->Ci == 32) && d->ci_pad64 == d->Ci;
}
// TMA needs a 16-byte aligned base: if the first tap of output column 0 sits at an odd pixel of a
// 4-channel row, the window starts one pixel earlier (the packed weights carry a zero pixel there).
static int window_lead(const pv_conv3d_desc* d) { return (((d->x_w_pad - d->pw) * d->Ci * 2) % 16) ? 1 : 0; }
static int window_elems(const pv_conv3d_desc* d) {
  const int run = (d->kw + window_lead(d)) * d->Ci;
  return run <= 16 ? 16 : (run <= 32 ? 32 : 64);
}

// Reduce (W,H,T,N) to <=4 merged dims: runs of adjacent trivial dims (k=1,s=1,p=0) collapse.
static void reduce_dims(const pv_conv3d_desc* d, IgemmPlan* pl) {
  const long long rs = d->x_row_stride * 2;   // bytes per position
  Dim4 raw[4];
  if (window_mode(d)) {
    const long long wp = d->x_w_phys;
    raw[0] = {d->Wo, d->Wo, 1, 1, 0, 1, (long long)d->sw * rs, true};
    raw[1] = {d->Hi, d->Ho, d->kh, d->sh, d->ph, d->dh, wp * rs, false};
    raw[2] = {d->Ti, d->To, d->kt, d->st, d->pt, d->dt, wp * d->Hi * rs, false};
    raw[3] = {d->N, d->N, 1, 1, 0, 1, wp * d->Hi * d->Ti * rs, false};
  } else {
    raw[0] = {d->Wi, d->Wo, d->kw, d->sw, d->pw, d->dw, rs, false};
    raw[1] = {d->Hi, d->Ho, d->kh, d->sh, d->ph, d->dh, (long long)d->Wi * rs, false};
    raw[2] = {d->Ti, d->To, d->kt, d->st, d->pt, d->dt, (long long)d->Wi * d->Hi * rs, false};
    raw[3] = {d->N, d->N, 1, 1, 0, 1, (long long)d->Wi * d->Hi * d->Ti * rs, false};
  }
  int n = 0;
  for (int i = 0; i < 4; ++i) {
    const bool triv = raw[i].k == 1 && raw[i].s == 1 && raw[i].p == 0 && !raw[i].nomerge;
    if (n > 0) {
      Dim4& prev = pl->dim[n - 1];
      const bool ptriv = prev.k == 1 && prev.s == 1 && prev.p == 0 && !prev.nomerge;
      if (triv && ptriv && prev.stride_bytes * prev.I == raw[i].stride_bytes) {
        prev.I *= raw[i].I;
        prev.O *= raw[i].O;
        pl->orig2m[i] = n - 1;
        continue;
      }
    }
    pl->orig2m[i] = n;
    pl->dim[n++] = raw[i];
  }
  pl->ndims = n;
  for (int i = n; i < 4; ++i)
    pl->dim[i] = {1, 1, 1, 1, 0, 1, pl->dim[n - 1].stride_bytes * pl->dim[n - 1].I, false};
}

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

int conv3d_tcgen05_supported(const pv_conv3d_desc* d, char* why, size_t why_len) {
#define NOPE(...)                          \
  do {                                     \
    if (why) snprintf(why, why_len, __VA_ARGS__); \
    return 0;                              \
  } while (0)
  if (d->dtype != PV_F16) NOPE("tcgen05 path needs f16 storage");
  if (d->groups != 1) NOPE("tcgen05 path is dense only (groups=%d)", d->groups);
  if (window_mode(d)) {
    if (d->Ci != 4 && d->Ci != 8) NOPE("window mode needs Ci in {4,8}");
    if (d->Co % 8) NOPE("Co must be a multiple of 8");
    if (d->dw != 1) NOPE("window mode needs dilation_w == 1");
    if (d->x_row_stride != d->Ci) NOPE("window mode needs densely packed pixels");
    if ((d->sw * d->Ci * 2) % 16) NOPE("window stride must be a multiple of 16 bytes");
    if ((d->kw + window_lead(d)) * d->Ci > 64) NOPE("window run longer than 64 elements");
    if (d->x_w_pad < d->pw) NOPE("physical W padding smaller than the conv padding");
    if (d->x_w_phys < d->x_w_pad + d->Wi + (d->pw > 0 ? d->pw : 0) || (d->x_w_phys * d->Ci * 2) % 16)
      NOPE("bad physical row width");
    if (d->kt * d->kh > IG_MAX_TAPS) NOPE("too many (dt,dh) taps");
    if (d->st * d->sh > IG_MAX_MAPS) NOPE("stride product too large");
    if (d->ci_pad64 != window_elems(d)) NOPE("window mode: ci_pad64 must equal the window length %d", window_elems(d));
    if (d->y_row_stride % 8 || (d->has_residual && d->res_row_stride % 8)) NOPE("row strides %% 8");
    return 1;
  }
  if (d->Ci % 8 || d->Co % 8) NOPE("Ci/Co must be multiples of 8");
  if (d->x_row_stride % 8 || d->y_row_stride % 8 || (d->has_residual && d->res_row_stride % 8))
    NOPE("row strides must be multiples of 8 elements (16 B)");
  if (d->kt * d->kh * d->kw > IG_MAX_TAPS) NOPE("too many taps");
  if (d->st * d->sh * d->sw > IG_MAX_MAPS) NOPE("stride product > %d", IG_MAX_MAPS);
  if (!conv3d_tma_narrow(d) && (d->ci_pad64 < d->Ci || d->ci_pad64 % 64)) NOPE("ci_pad64 must be a multiple of 64 >= Ci");
  const long long M = (long long)d->N * d->To * d->Ho * d->Wo;
  if (M >= (1ll << 31)) NOPE("too many output positions");
  for (int off : {d->pt, d->ph, d->pw, d->dt * (d->kt - 1), d->dh * (d->kh - 1), d->dw * (d->kw - 1)})
    if (off > 100) NOPE("tap offset too large");
  return 1;
#undef NOPE
}

static double tile_cycles(int n) {   // crude per-k-block cost model
  double mma = 2.0 * n, smem = 128.0 + n;
  return mma > smem ? mma : smem;
}

int conv3d_tcgen05_launch(const pv_conv3d_desc* d, const void* x, const void* w, const float* scale,
                          const float* bias, const void* residual, void* y, cudaStream_t stream) {
  char why[160];
  if (!conv3d_tcgen05_supported(d, why, sizeof(why))) {
    set_error("PV_ALGO_TCGEN05 unsupported: %s", why);
    return PV_ERR_UNSUPPORTED;
  }
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return PV_ERR_CUDA; }

  const int sm_count = current_sm_count();
  if (sm_count <= 0) { set_error("cannot query the SM count of the current device"); return PV_ERR_CUDA; }

  IgemmPlan pl;
  reduce_dims(d, &pl);
  IgemmParams P;
  memset(&P, 0, sizeof(P));

  // ---- tile box search: maximise useful rows per 128-row tile
  {
    long long O[4];
    for (int i = 0; i < 4; ++i) O[i] = pl.dim[i].O;
    double best = -1;
    int bb[4] = {1, 1, 1, 1};
    for (int b0 = 1; b0 <= 128 && b0 <= O[0]; ++b0) {
      for (int b1 = 1; b0 * b1 <= 128 && b1 <= O[1]; ++b1) {
        for (int b2 = 1; b0 * b1 * b2 <= 128 && b2 <= O[2]; ++b2) {
          int b3 = 128 / (b0 * b1 * b2);
          if (b3 > O[3]) b3 = (int)O[3];
          if (b3 < 1) b3 = 1;
          const double tiles = (double)cdiv(O[0], b0) * cdiv(O[1], b1) * cdiv(O[2], b2) * cdiv(O[3], b3);
          const double eff = (double)O[0] * O[1] * O[2] * O[3] / (tiles * 128.0);
          // prefer efficiency, then longer inner runs
          const double score = eff + 1e-6 * b0;
          if (score > best) { best = score; bb[0] = b0; bb[1] = b1; bb[2] = b2; bb[3] = b3; }
        }
      }
    }
    long long mt = 1;
    for (int i = 0; i < 4; ++i) {
      P.O[i] = (int)O[i];
      P.box[i] = bb[i];
      P.nt[i] = (int)cdiv(O[i], bb[i]);
      mt *= P.nt[i];
    }
    P.rows = bb[0] * bb[1] * bb[2] * bb[3];
    P.m_tiles = (int)mt;
  }

  // ---- BLOCK_N: minimise waves * per-tile cost
  {
    const int co16 = (int)cdiv(d->Co, 16) * 16;
    int cands[6] = {256, 128, 64, 32, 16, co16 <= 256 ? co16 : 256};
    double best = 1e30;
    int bn = 16;
    for (int c : cands) {
      if (c > 256 || c > co16) continue;
      if (c < co16 && (c % 64)) continue;   // several N tiles: TMA-store sub-tiles are 64 channels wide
      const long long tiles = (long long)P.m_tiles * cdiv(d->Co, c);
      const double waves = (double)cdiv(tiles, sm_count);
      const double cost = waves * tile_cycles(c) + 1e-3 * cdiv(d->Co, c);
      if (cost < best) { best = cost; bn = c; }
    }
    // PVB200_BN=64|128|256: A/B override of the tile width (tools/epi_sweep.py); read per launch on purpose
    { const char* e = getenv("PVB200_BN"); const int v = e ? atoi(e) : 0; if ((v == 64 || v == 128 || v == 256) && v < co16) bn = v; }
    P.block_n = bn;
    P.n_tiles = (int)cdiv(d->Co, bn);
  }
  const bool wmode = window_mode(d);
  const bool narrow = conv3d_tma_narrow(d);
  const int win = wmode ? window_elems(d) : (narrow ? d->Ci : 64);
  P.kbytes = 2 * win;
  // CTA pair (cta_group::2): full 64-channel k-blocks, N a multiple of 16 and at least two M tiles.  Each CTA then
  // stages A (16 KiB) + HALF of B per k-block.  PVB200_PAIR=0 disables, =2 forces it for every eligible layer.
  {
    static const int pair_env = getenv("PVB200_PAIR") ? atoi(getenv("PVB200_PAIR")) : 0;
    const long long k_total = (long long)d->kt * d->kh * d->kw * d->ci_pad64;
    const bool eligible = !wmode && !narrow && P.kbytes == 128 && P.block_n % 16 == 0 && P.block_n >= 32 && P.m_tiles >= 2 &&
                          sm_count >= 2;
    // worth it where the mainloop dominates: deep K with a wide tile (shallow pointwise layers are epilogue-bound)
    const bool wanted = pair_env == 2 || (pair_env == 1 && P.block_n >= 128 && k_total >= 512);
    P.pair = (eligible && wanted) ? 1 : 0;
    P.pair_tiles = P.n_tiles * (int)cdiv(P.m_tiles, 2);
  }
  P.Co = d->Co;
  P.taps = wmode ? d->kt * d->kh : d->kt * d->kh * d->kw;
  P.num_kc = (wmode || narrow) ? 1 : d->ci_pad64 / 64;
  P.epi.block_n = P.block_n;
  P.epi.Co = d->Co;
  P.epi.rows = P.rows;
  P.epi.act = d->act;
  P.epi.has_residual = d->has_residual;
  { const char* e = getenv("PVB200_DEBUG"); P.epi.dbg = e ? atoi(e) : 0; }
  // wide residual tiles: two epilogue teams on alternate tiles (needs two accumulators in TMEM); PVB200_EPI_TEAMS=1|2, read per launch
  {
    // Measured (profiles/r02_epilogue_sweep.md, one box): res2 conv_c 69.8 -> 62.1 us; SlowFast 3.070 -> 3.014 ms, CSN-R101 5.37 -> 5.19 ms,
    // R(2+1)D 3.166 -> 3.070 ms, Slow-R50 1.913 -> 1.866 ms; X3D-M (96 / 192-wide tiles) 7.03 -> 7.14 ms - hence full-width tiles only.
    // Per layer (same file, run 4): 10.6 tiles per CTA (res2 conv_c) 69.5 -> 61.6 us, 5.3 tiles (res3) 45.2 -> 45.3 us, 2.6 tiles (res4)
    // 23.4 -> 25.4 us, 1.4 tiles (res5) 19.0 -> 19.5 us: a tile drained by four warps instead of eight only pays when each team has several
    // tiles to pipeline - hence at least four tiles per CTA.
    const char* e = getenv("PVB200_EPI_TEAMS");
    const int want = e ? atoi(e) : 2;
    P.epi.teams = (want == 2 && !P.pair && d->has_residual && P.block_n == 256 &&
                   (long long)P.m_tiles * P.n_tiles >= 4ll * sm_count) ? 2 : 1;
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
  // k-blocks per pipeline stage: a barrier round + commit costs the issuing thread ~500+ clk, one k-block
  // of MMAs only k16 * N/2 clk - group k-blocks until a stage carries ~1000 clk of tensor work.
  const int kb_bytes = (IG_BM + (P.pair ? P.block_n / 2 : P.block_n)) * P.kbytes;
  {
    const int num_kb = P.taps * P.num_kc;
    const int mma_clk = (P.kbytes >> 5) * (P.block_n < 16 ? 16 : P.block_n) / 2;
    int G = (1000 + mma_clk - 1) / mma_clk;
    if (G > 4) G = 4;
    if (G > num_kb) G = num_kb;
    { const char* e = getenv("PVB200_G"); if (e && atoi(e) >= 1 && atoi(e) <= 8) G = atoi(e) < num_kb ? atoi(e) : num_kb; }
    P.epi_bytes = EPI_SMEM_BYTES + (epi_wide_prefetch(P.epi) ? (EPI_RING - 1) * EPI_STAGING_BYTES : 0);
    {
      static const bool no_alias = getenv("PVB200_NO_ALIAS") != nullptr;
      const long long units = P.pair ? 2ll * P.pair_tiles : (long long)P.m_tiles * P.n_tiles;
      P.alias_epi = (!no_alias && units <= sm_count && !epi_wide_prefetch(P.epi) && !epi_direct(P.epi)) ? 1 : 0;
    }
    const int budget = 227 * 1024 - 2048 - (P.alias_epi ? 0 : P.epi_bytes) - 256;
    while (G > 1 && budget / (G * kb_bytes) < 3) --G;
    int st = budget / (G * kb_bytes);
    if (st > 24 / G) st = 24 / G > 2 ? 24 / G : 2;
    if (st < 2) st = 2;
    P.G = G;
    {
      // default: 2 warps (A / B) for one k-block per stage, 4 for chunked stages (window / narrow modes: up to 8 loads)
      static const int split_env = getenv("PVB200_SPLIT_AB") ? atoi(getenv("PVB200_SPLIT_AB")) : -1;
      int ns = split_env >= 1 ? split_env : (G >= 2 ? 4 : 2);
      if (ns > IG_PROD_WARPS) ns = IG_PROD_WARPS;
      if (ns == 3) ns = 2;
      P.split_ab = P.pair ? 1 : ns;
    }
    P.cpt = (num_kb + G - 1) / G;
    P.stages = st;
  }
  const int stage_bytes = P.G * kb_bytes;
  if (P.alias_epi && (long long)P.stages * stage_bytes < P.epi_bytes) P.alias_epi = 0;     // ring smaller than the staging area
  const size_t smem_bytes = (size_t)P.stages * stage_bytes + 2048 /*align*/ + (P.alias_epi ? 0 : P.epi_bytes) +
                            8 * (2 * P.stages + 2 * 8 + 2 * EPI_RING + 1) + 16;

  // ---- taps -> (parity map, coordinate shift); original dims order: tap index = (kt, kh, kw)
  // merged dims never merge a dim that has taps, so each non-trivial original dim maps to one
  // merged dim.  Build per-merged-dim lists.
  int ks[4], ss[4], ps[4], dls[4];
  for (int i = 0; i < 4; ++i) { ks[i] = pl.dim[i].k; ss[i] = pl.dim[i].s; ps[i] = pl.dim[i].p; dls[i] = pl.dim[i].dil; }
  const int* orig2m = pl.orig2m;
  unsigned used_maps = 0;
  const int kw_loop = wmode ? 1 : d->kw;
  for (int it = 0; it < d->kt; ++it)
    for (int ih = 0; ih < d->kh; ++ih)
      for (int iw = 0; iw < kw_loop; ++iw) {
        const int tap = (it * d->kh + ih) * kw_loop + iw;
        int q[4] = {0, 0, 0, 0}, r[4] = {0, 0, 0, 0};
        // window mode: the W taps live inside the window and the W padding is physical
        const int offs[3] = {wmode ? 0 : iw * d->dw - d->pw, ih * d->dh - d->ph, it * d->dt - d->pt};
        const int strd[3] = {wmode ? 1 : d->sw, d->sh, d->st};
        for (int o = 0; o < 3; ++o) {
          const int m = orig2m[o];
          if (strd[o] == 1 && offs[o] == 0) continue;
          const int qq = floordiv(offs[o], strd[o]);
          q[m] = qq;
          r[m] = offs[o] - qq * strd[o];
        }
        int cls = 0, mul = 1;
        for (int m = 0; m < 4; ++m) { cls += r[m] * mul; mul *= ss[m]; }
        if (cls >= IG_MAX_MAPS) { set_error("internal: parity class %d", cls); return PV_ERR_INVALID; }
        P.tap_map[tap] = (unsigned char)cls;
        for (int m = 0; m < 4; ++m) P.tap_q[tap][m] = (signed char)q[m];
        used_maps |= 1u << cls;
      }

  // ---- tensor maps
  for (int cls = 0; cls < IG_MAX_MAPS; ++cls) {
    if (!(used_maps & (1u << cls))) continue;
    int r[4], c = cls;
    for (int m = 0; m < 4; ++m) { r[m] = c % ss[m]; c /= ss[m]; }
    long long base_off = wmode ? (long long)(d->x_w_pad - d->pw - window_lead(d)) * d->Ci * 2 : 0;   // bytes
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
    gdim[0] = (cuuint64_t)(wmode ? win : d->Ci);
    box[0] = (cuuint32_t)win;
    for (int m = 0; m < 4; ++m) {
      const long long cnt = (pl.dim[m].I - r[m] + ss[m] - 1) / ss[m];
      gdim[m + 1] = (cuuint64_t)(cnt > 0 ? cnt : 1);
      gstr[m] = (cuuint64_t)(pl.dim[m].stride_bytes * ss[m]);
      box[m + 1] = (cuuint32_t)P.box[m];
      base_off += (long long)r[m] * pl.dim[m].stride_bytes;
    }
    void* gptr = (void*)((const char*)x + base_off);
    const CUtensorMapSwizzle swz = P.kbytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : (P.kbytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult cr = encode(&P.a_maps[cls], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, gptr, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(A, class %d) failed: %d (Ci=%d dims=%llu,%llu,%llu,%llu box=%u,%u,%u,%u)",
                cls, (int)cr, d->Ci, (unsigned long long)gdim[1], (unsigned long long)gdim[2],
                (unsigned long long)gdim[3], (unsigned long long)gdim[4], box[1], box[2], box[3], box[4]);
      return PV_ERR_CUDA;
    }
  }
  {
    const long long krow = (long long)P.taps * d->ci_pad64;     // window mode: ci_pad64 == window length
    const long long kpitch = narrow ? (krow + 63) / 64 * 64 : krow;   // narrow mode shares the gather packing (row end padded to 64)
    cuuint64_t gdim[2] = {(cuuint64_t)krow, (cuuint64_t)d->Co};
    cuuint64_t gstr[1] = {(cuuint64_t)kpitch * 2};
    cuuint32_t box[2] = {(cuuint32_t)win, (cuuint32_t)(P.pair ? P.block_n / 2 : P.block_n)}, estr[2] = {1, 1};
    const CUtensorMapSwizzle swz = P.kbytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : (P.kbytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult cr = encode(&P.b_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)w, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)cr); return PV_ERR_CUDA; }
  }

  // ---- output / residual tensor maps (merged OUTPUT dims follow the input merge pattern)
  {
    const long long ostr_orig[4] = {1, d->Wo, (long long)d->Wo * d->Ho, (long long)d->Wo * d->Ho * d->To};
    long long ostr[4] = {0, 0, 0, 0};
    bool seen[4] = {false, false, false, false};
    for (int o = 0; o < 4; ++o) {
      const int m = pl.orig2m[o];
      if (!seen[m]) { seen[m] = true; ostr[m] = ostr_orig[o]; }
    }
    P.epi.y_ptr = (__half*)y;
    P.epi.r_ptr = (const __half*)residual;
    for (int m = 0; m < 4; ++m) {
      P.epi.O[m] = P.O[m];
      P.epi.box[m] = P.box[m];
      P.epi.y_str[m] = seen[m] ? ostr[m] * d->y_row_stride : 0;      // filler dims have extent 1
      P.epi.r_str[m] = seen[m] ? ostr[m] * d->res_row_stride : 0;
    }
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1 && !d->has_residual) break;
      const long long rs = pass == 0 ? d->y_row_stride : d->res_row_stride;
      void* base = pass == 0 ? y : const_cast<void*>(residual);
      cuuint64_t gdim[5], gstr[4];
      cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
      gdim[0] = (cuuint64_t)d->Co;
      box[0] = 64;
      long long prev = rs * 2;      // byte extent covered so far (for filler dims)
      for (int m = 0; m < 4; ++m) {
        gdim[m + 1] = (cuuint64_t)P.O[m];
        const long long sb = seen[m] ? ostr[m] * rs * 2 : prev;
        gstr[m] = (cuuint64_t)sb;
        prev = sb * (long long)P.O[m];
        box[m + 1] = (cuuint32_t)P.box[m];
      }
      CUresult cr = encode(pass == 0 ? &P.epi.y_map : &P.epi.r_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, base, gdim,
                           gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(%s) failed: %d", pass == 0 ? "Y" : "R", (int)cr);
        return PV_ERR_CUDA;
      }
    }
  }

  PV_OPT_IN_SMEM(conv3d_igemm_kernel<false>, 227 * 1024);
  PV_OPT_IN_SMEM(conv3d_igemm_kernel<true>, 227 * 1024);
  const long long total_tiles = (long long)P.m_tiles * P.n_tiles;
  if (total_tiles == 0) return PV_OK;
  int grid = (int)(total_tiles < sm_count ? total_tiles : sm_count);
  if (P.pair) {          // one cluster of two CTAs per TPC
    const int clusters = P.pair_tiles < sm_count / 2 ? P.pair_tiles : sm_count / 2;
    grid = 2 * clusters;
  }
  {
    // launched with the programmatic-stream-serialization attribute (PDL); PVB200_NO_PDL=1 falls back to a plain launch
    static const bool use_pdl = getenv("PVB200_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(IG_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (use_pdl) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    if (P.pair) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = 2;
      attr[na].val.clusterDim.y = 1;
      attr[na].val.clusterDim.z = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    if (P.pair) PV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv3d_igemm_kernel<true>, P, scale, bias));
    else PV_CUDA_OK(cudaLaunchKernelEx(&cfg, conv3d_igemm_kernel<false>, P, scale, bias));
  }
  PV_LAUNCH_OK("conv3d_igemm_kernel");
  return PV_OK;
}

}  // namespace pv
