// Pooled multi-head attention of MViT (layers/attention.py:531-539) on tcgen05 tensor cores (sm_100a):
//
//   O = softmax(scale * Q K^T) V (+ Q),   Q [B, Nq, H*D], K / V [B, Nk, H*D] token-major f16, head h = channels h*D .. h*D+D-1
//
// One CTA = 128 query rows of one (batch, head).  Both GEMMs run on tcgen05.mma with accumulators in TMEM:
//   S = Q K^T : M = 128 queries, N = 128 keys, K = D.   Q and K tiles arrive by TMA as D/32 chunks of [rows x 32] f16
//               (64-byte rows, SWIZZLE_64B, K-major) - the K tile is used exactly as it lies in memory.
//   O += P V  : M = 128 queries, N = 32 (one chunk of D) x D/32, K = 128 keys.  P (f16) is written by the softmax warps
//               into shared memory in the 128B-swizzled K-major layout; V is the B operand in MN-MAJOR form (rows = keys,
//               contiguous along D), i.e. again the TMA tile as it lies in memory - no transposed copy
//               (tools/probe/umma_mnmajor.cu pins the descriptor convention).
// Softmax without rescaling: the key tiles are swept TWICE.  Sweep 1 computes S tile by tile and reduces the row maxima;
// sweep 2 recomputes S, writes P = exp2((S - max) * scale * log2 e) and accumulates O in TMEM with plain accumulation.
// The second Q K^T costs 50 % more tensor work (the tensor pipe is far from the limiter here) and removes the
// TMEM read-modify-write of O per key tile that an online softmax needs.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2-5 = softmax / epilogue (one thread per
// query row; warp w may touch TMEM lanes 32 (w % 4) ..).  All hand-offs are mbarriers; waits are bounded (trap).
#include "pv_common.cuh"
#include "pv_sm100.cuh"

#include <stdlib.h>
#include <string.h>

namespace pv {

using namespace sm100;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();   // pv_igemm.cu

constexpr int AT_BQ = 128, AT_BK = 128;
constexpr int AT_THREADS = 6 * 32;
constexpr int AT_CHUNK_BYTES = 128 * 64;            // [128 rows x 32 f16] swizzle-64B chunk

struct AttnTcParams {
  CUtensorMap q_map, k_map, v_map;                  // [H*D, N, B] f16, box [32, 128, 1], SWIZZLE_64B
  int B, H, Nq, Nk, nkt;
  float scale_log2e;
  int add_q_residual;
  long long q_row_stride, q_batch_stride, o_row_stride, o_batch_stride;
  unsigned v_lbo, v_sbo;                            // MN-major descriptor strides of the V operand (bytes)
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ uint64_t at_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint64_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}
__device__ __forceinline__ void at_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ float at_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ AttnTcParams P, const __half* __restrict__ q, __half* __restrict__ o) {
  constexpr int NCH = D / 32;                                   // 32-channel chunks of the head dim
  constexpr uint32_t TILE_BYTES = NCH * AT_CHUNK_BYTES;         // one Q / K / V tile
  constexpr uint32_t P_BYTES = 2 * 128 * 128;                   // [128 x 128] f16 as two 128B-swizzled chunks of 64 keys
  extern __shared__ uint8_t at_smem_raw[];
  const uint32_t sbase = (smem_u32(at_smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = at_smem_raw + (sbase - smem_u32(at_smem_raw));
  const uint32_t q_s = sbase, k_s = q_s + TILE_BYTES, v_s = k_s + 2 * TILE_BYTES, p_s = v_s + 2 * TILE_BYTES;
  const uint32_t bar0 = p_s + 2 * P_BYTES;
  // barriers
  const uint32_t qfull = bar0;
  auto kfull = [&](int i) { return bar0 + 8u * (1 + i); };
  auto kempty = [&](int i) { return bar0 + 8u * (3 + i); };
  auto vfull = [&](int i) { return bar0 + 8u * (5 + i); };
  auto vempty = [&](int i) { return bar0 + 8u * (7 + i); };
  auto sfull = [&](int i) { return bar0 + 8u * (9 + i); };
  auto sempty = [&](int i) { return bar0 + 8u * (11 + i); };
  auto pfull = [&](int i) { return bar0 + 8u * (13 + i); };
  auto pempty = [&](int i) { return bar0 + 8u * (15 + i); };
  const uint32_t ofull = bar0 + 8u * 17;
  const uint32_t tmem_slot = bar0 + 8u * 18;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ;
  const int b = blockIdx.y / P.H, h = blockIdx.y - b * P.H;
  const int nkt = P.nkt;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&P.q_map); prefetch_tmap(&P.k_map); prefetch_tmap(&P.v_map);
    mbar_init(qfull, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(kfull(i), 1); mbar_init(kempty(i), 1);
      mbar_init(vfull(i), 1); mbar_init(vempty(i), 1);
      mbar_init(sfull(i), 1); mbar_init(sempty(i), 4);       // one arrive per softmax warp
      mbar_init(pfull(i), 4); mbar_init(pempty(i), 1);
    }
    mbar_init(ofull, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t s_tmem = tmem_base;              // S buffers: columns [0,128) and [128,256)
  const uint32_t o_tmem = tmem_base + 256u;       // O: columns [256, 256 + D)

  if (warp == 0) {
    // ================================ TMA producer ==========================================================
    if (elect_one()) {
      mbar_arrive_expect_tx(qfull, TILE_BYTES);
      for (int c = 0; c < NCH; ++c) tma_load_3d(q_s + c * AT_CHUNK_BYTES, &P.q_map, qfull, h * D + c * 32, q0, b);
    }
    __syncwarp();
    int vi = 0;
    for (int s = 0; s < 2 * nkt; ++s) {
      const int j = s < nkt ? s : s - nkt;
      const int buf = s & 1;
      mbar_wait(kempty(buf), ((uint32_t)(s >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(kfull(buf), TILE_BYTES);
        for (int c = 0; c < NCH; ++c) tma_load_3d(k_s + buf * TILE_BYTES + c * AT_CHUNK_BYTES, &P.k_map, kfull(buf), h * D + c * 32, j * AT_BK, b);
      }
      __syncwarp();
      if (s >= nkt) {
        const int vb = vi & 1;
        mbar_wait(vempty(vb), ((uint32_t)(vi >> 1) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(vfull(vb), TILE_BYTES);
          for (int c = 0; c < NCH; ++c) tma_load_3d(v_s + vb * TILE_BYTES + c * AT_CHUNK_BYTES, &P.v_map, vfull(vb), h * D + c * 32, j * AT_BK, b);
        }
        __syncwarp();
        ++vi;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ============================================================
    const uint32_t idesc_s = make_idesc_f16(128, 128);
    const uint32_t idesc_pv = make_idesc_f16(128, 32) | (1u << 16);         // B operand MN-major
    auto issue_s = [&](int s) {
      const int buf = s & 1;
      mbar_wait(kfull(buf), (uint32_t)(s >> 1) & 1u);
      mbar_wait(sempty(buf), ((uint32_t)(s >> 1) & 1u) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const uint64_t a = make_kmajor_desc(q_s + c * AT_CHUNK_BYTES, 64);
          const uint64_t bd = make_kmajor_desc(k_s + buf * TILE_BYTES + c * AT_CHUNK_BYTES, 64);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            umma_f16(s_tmem + (uint32_t)buf * 128u, a + (uint64_t)(2 * ks), bd + (uint64_t)(2 * ks), idesc_s, (c | ks) != 0 ? 1u : 0u);
        }
        umma_commit(kempty(buf));
        umma_commit(sfull(buf));
      }
      __syncwarp();
    };
    mbar_wait(qfull, 0);
    tc_fence_after();
    for (int s = 0; s < nkt; ++s) issue_s(s);                 // sweep 1: row maxima
    issue_s(nkt);                                              // sweep 2, first tile
    for (int j = 0; j < nkt; ++j) {
      if (j + 1 < nkt) issue_s(nkt + j + 1);                   // keep the tensor pipe busy while the softmax of tile j runs
      const int pb = j & 1;
      mbar_wait(pfull(pb), (uint32_t)(j >> 1) & 1u);
      mbar_wait(vfull(pb), (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {                       // 16 keys per MMA
          const uint64_t a = make_kmajor_desc(p_s + pb * P_BYTES + (uint32_t)(ks >> 2) * 16384u, 128) + (uint64_t)(2 * (ks & 3));
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            const uint64_t bd = at_desc(v_s + pb * TILE_BYTES + c * AT_CHUNK_BYTES + (uint32_t)ks * 16u * 64u, P.v_lbo, P.v_sbo, 4);
            umma_f16(o_tmem + (uint32_t)c * 32u, a, bd, idesc_pv, (j | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(pempty(pb));
        umma_commit(vempty(pb));
        if (j == nkt - 1) umma_commit(ofull);
      }
      __syncwarp();
    }
  } else {
    // ================================ softmax / epilogue warps ==============================================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    float m = -INFINITY;
    // ---- sweep 1: row maximum of the raw scores
    for (int s = 0; s < nkt; ++s) {
      const int buf = s & 1;
      mbar_wait(sfull(buf), (uint32_t)(s >> 1) & 1u);
      tc_fence_after();
      const int valid = min(AT_BK, P.Nk - s * AT_BK);
#pragma unroll 1
      for (int c0 = 0; c0 < AT_BK; c0 += 32) {
        uint32_t v[32];
        at_ld32(s_tmem + lane_off + (uint32_t)buf * 128u + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < valid) m = fmaxf(m, __uint_as_float(v[i]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sempty(buf));
    }
    const float mc = m * P.scale_log2e;
    float l = 0.f;
    // ---- sweep 2: P = exp2(s * c - m * c), row sums, P -> shared memory (128B-swizzled K-major, f16)
    for (int j = 0; j < nkt; ++j) {
      const int s = nkt + j, buf = s & 1, pb = j & 1;
      mbar_wait(sfull(buf), (uint32_t)(s >> 1) & 1u);
      mbar_wait(pempty(pb), ((uint32_t)(j >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const int valid = min(AT_BK, P.Nk - j * AT_BK);
      uint8_t* prow = sgen + (p_s - sbase) + pb * P_BYTES + row * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < AT_BK; c0 += 32) {
        uint32_t v[32];
        at_ld32(s_tmem + lane_off + (uint32_t)buf * 128u + (uint32_t)c0, v);
        tmem_ld_wait();
        uint8_t* pchunk = prow + (c0 >> 6) * 16384;
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          uint4 pk;
          __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i0 = g8 * 8 + 2 * e;
            float p0 = (c0 + i0 < valid) ? at_exp2(fmaf(__uint_as_float(v[i0]), P.scale_log2e, -mc)) : 0.f;
            float p1 = (c0 + i0 + 1 < valid) ? at_exp2(fmaf(__uint_as_float(v[i0 + 1]), P.scale_log2e, -mc)) : 0.f;
            const __half2 hp = __floats2half2_rn(p0, p1);
            const float2 back = __half22float2(hp);
            l += back.x + back.y;                        // the row sum of the values the P.V GEMM really uses
            ph[e] = hp;
          }
          const uint32_t cc = (uint32_t)((c0 & 63) >> 3) + (uint32_t)g8;         // 16-byte chunk inside the 128-byte row
          *reinterpret_cast<uint4*>(pchunk + ((cc ^ (uint32_t)(row & 7)) << 4)) = pk;
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // P was written by the generic proxy, UMMA reads it
      __syncwarp();
      if (lane == 0) { mbar_arrive(sempty(buf)); mbar_arrive(pfull(pb)); }
    }
    // ---- epilogue: O / l (+ q), f16, one row of D channels per thread
    mbar_wait(ofull, 0);
    tc_fence_after();
    const int qi = q0 + row;
    const float inv = 1.f / l;
    __half* orow = o + (long long)b * P.o_batch_stride + (long long)qi * P.o_row_stride + h * D;
    const __half* qrow = q + (long long)b * P.q_batch_stride + (long long)qi * P.q_row_stride + h * D;
#pragma unroll 1
    for (int c0 = 0; c0 < D; c0 += 32) {
      uint32_t v[32];
      at_ld32(o_tmem + lane_off + (uint32_t)c0, v);
      tmem_ld_wait();
      if (qi < P.Nq) {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[g8 * 8 + e]) * inv;
          if (P.add_q_residual) {
            const uint4 qv = *reinterpret_cast<const uint4*>(qrow + c0 + g8 * 8);
            const __half2* qh = reinterpret_cast<const __half2*>(&qv);
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 t = __half22float2(qh[e]); f[2 * e] += t.x; f[2 * e + 1] += t.y; }
          }
          uint4 ov;
          __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
          for (int e = 0; e < 4; ++e) oh[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
          *reinterpret_cast<uint4*>(orow + c0 + g8 * 8) = ov;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

template <int D>
static int launch_attention_tc(const pv_attention_desc* d, const void* q, const void* k, const void* v, void* o, cudaStream_t s) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return PV_ERR_UNSUPPORTED;
  AttnTcParams P;
  memset(&P, 0, sizeof(P));
  P.B = d->B; P.H = d->H; P.Nq = d->Nq; P.Nk = d->Nk; P.nkt = (int)cdiv(d->Nk, AT_BK);
  P.scale_log2e = d->scale * 1.4426950408889634f;
  P.add_q_residual = d->add_q_residual;
  P.q_row_stride = d->q_row_stride; P.q_batch_stride = d->q_batch_stride;
  P.o_row_stride = d->o_row_stride; P.o_batch_stride = d->o_batch_stride;
  {
    // MN-major, SWIZZLE_64B: PVB200_ATTN_VDESC=<lbo>,<sbo> overrides (tools/probe/umma_mnmajor.cu)
    P.v_lbo = 8 * 64; P.v_sbo = 8 * 64;
    const char* e = getenv("PVB200_ATTN_VDESC");
    if (e) { unsigned a = 0, b2 = 0; if (sscanf(e, "%u,%u", &a, &b2) == 2) { P.v_lbo = a; P.v_sbo = b2; } }
  }
  struct { CUtensorMap* m; const void* p; long long rs, bs; int n; } maps[3] = {
      {&P.q_map, q, d->q_row_stride, d->q_batch_stride, d->Nq},
      {&P.k_map, k, d->k_row_stride, d->k_batch_stride, d->Nk},
      {&P.v_map, v, d->v_row_stride, d->v_batch_stride, d->Nk}};
  for (auto& t : maps) {
    cuuint64_t gdim[3] = {(cuuint64_t)(d->H * D), (cuuint64_t)t.n, (cuuint64_t)d->B};
    cuuint64_t gstr[2] = {(cuuint64_t)t.rs * 2, (cuuint64_t)t.bs * 2};
    cuuint32_t box[3] = {32, 128, 1}, estr[3] = {1, 1, 1};
    CUresult cr = encode(t.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(t.p), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return PV_ERR_UNSUPPORTED;
  }
  constexpr size_t smem = 2048 + (size_t)(D / 32) * AT_CHUNK_BYTES * 5 + 2 * 2 * 128 * 128 + 8 * 20 + 16;
  PV_OPT_IN_SMEM(attention_tc_kernel<D>, smem);
  dim3 grid((unsigned)cdiv(d->Nq, AT_BQ), (unsigned)(d->B * d->H)), block(AT_THREADS);
  attention_tc_kernel<D><<<grid, block, smem, s>>>(P, (const __half*)q, (__half*)o);
  PV_LAUNCH_OK("attention_tc_kernel");
  return PV_OK;
}

// tcgen05 path; returns PV_ERR_UNSUPPORTED when the shape does not qualify (the caller falls back to the mma.sync kernel)
int attention_tc_dispatch(const pv_attention_desc* d, const void* q, const void* k, const void* v, void* o, cudaStream_t s) {
  static const bool off = getenv("PVB200_ATTN_NO_TC") != nullptr;
  if (off || d->dtype != PV_F16) return PV_ERR_UNSUPPORTED;
  if (d->q_row_stride % 8 || d->k_row_stride % 8 || d->v_row_stride % 8 || d->o_row_stride % 8) return PV_ERR_UNSUPPORTED;
  if (d->q_batch_stride % 8 || d->k_batch_stride % 8 || d->v_batch_stride % 8 || d->o_batch_stride % 8) return PV_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(o)) & 15)
    return PV_ERR_UNSUPPORTED;
  if (d->B * d->H > 65535 || d->Nk < 1 || d->Nq < 1) return PV_ERR_UNSUPPORTED;
  switch (d->D) {
    case 32: return launch_attention_tc<32>(d, q, k, v, o, s);
    case 64: return launch_attention_tc<64>(d, q, k, v, o, s);
    case 96: return launch_attention_tc<96>(d, q, k, v, o, s);
    default: return PV_ERR_UNSUPPORTED;
  }
}

}  // namespace pv
