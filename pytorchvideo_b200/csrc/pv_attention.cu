// MViT pooled attention, flash style: o = softmax((q*scale) k^T) v (+ q); the N_q x N_k score
// matrix never reaches HBM (reference layers/attention.py:531-539 materialises it).
//
// Generic CUDA-core kernel (f16 or f32 storage, fp32 maths, any head dim D <= 128 with D%32==0).
// A CTA owns ATT_BQ query rows of one (batch, head); K/V stream through shared memory in tiles
// of 32 keys; online softmax state (running max / sum / output) stays in registers.
#include "pv_common.cuh"
#include <stdlib.h>

namespace pv {

constexpr int ATT_WARPS = 8;
constexpr int ATT_QPW = 4;                     // queries per warp
constexpr int ATT_BQ = ATT_WARPS * ATT_QPW;    // 32 queries per CTA
constexpr int ATT_BK = 32;                     // keys per tile

template <typename T, int D>
__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_kernel(pv_attention_desc d, const T* __restrict__ q, const T* __restrict__ k,
                 const T* __restrict__ v, T* __restrict__ o) {
  constexpr int DS = D + 1;            // padded smem row stride (floats)
  constexpr int NC = D / 32;           // output columns per lane
  extern __shared__ float sh[];
  float* Qs = sh;                      // [ATT_BQ][DS]
  float* Ks = Qs + ATT_BQ * DS;        // [ATT_BK][DS]
  float* Vs = Ks + ATT_BK * DS;        // [ATT_BK][DS]
  float* Ps = Vs + ATT_BK * DS;        // [ATT_WARPS][ATT_QPW][32]

  const int bh = blockIdx.y;
  const int b = bh / d.H, h = bh - b * d.H;
  const int q0 = blockIdx.x * ATT_BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const T* qb = q + (long long)b * d.q_batch_stride + (long long)h * D;
  const T* kb = k + (long long)b * d.k_batch_stride + (long long)h * D;
  const T* vb = v + (long long)b * d.v_batch_stride + (long long)h * D;
  T* ob = o + (long long)b * d.o_batch_stride + (long long)h * D;

  // stage the query tile (unscaled copy kept for the residual; scaled used for the scores)
  for (int e = threadIdx.x; e < ATT_BQ * D; e += blockDim.x) {
    const int r = e / D, c = e - r * D;
    const int qi = q0 + r;
    Qs[r * DS + c] = qi < d.Nq ? Elem<T>::ld(qb + (long long)qi * d.q_row_stride + c) : 0.f;
  }

  float m[ATT_QPW], l[ATT_QPW], acc[ATT_QPW][NC];
#pragma unroll
  for (int i = 0; i < ATT_QPW; ++i) {
    m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[i][c] = 0.f;
  }

  for (int k0 = 0; k0 < d.Nk; k0 += ATT_BK) {
    __syncthreads();   // previous tile fully consumed (also covers the Q staging above)
    for (int e = threadIdx.x; e < ATT_BK * D; e += blockDim.x) {
      const int r = e / D, c = e - r * D;
      const int ki = k0 + r;
      float kv = 0.f, vv = 0.f;
      if (ki < d.Nk) {
        kv = Elem<T>::ld(kb + (long long)ki * d.k_row_stride + c);
        vv = Elem<T>::ld(vb + (long long)ki * d.v_row_stride + c);
      }
      Ks[r * DS + c] = kv;
      Vs[r * DS + c] = vv;
    }
    __syncthreads();
    const bool key_ok = (k0 + lane) < d.Nk;
#pragma unroll
    for (int i = 0; i < ATT_QPW; ++i) {
      const float* qrow = Qs + (warp * ATT_QPW + i) * DS;
      const float* krow = Ks + lane * DS;
      float s = 0.f;
#pragma unroll 8
      for (int c = 0; c < D; ++c) s = fmaf(qrow[c] * d.scale, krow[c], s);
      s = key_ok ? s : -INFINITY;
      float mx = s;
      for (int off = 16; off; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m[i], mx);
      const float p = key_ok ? __expf(s - m_new) : 0.f;
      float ps = p;
      for (int off = 16; off; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      const float corr = (m[i] == -INFINITY) ? 0.f : __expf(m[i] - m_new);
      l[i] = l[i] * corr + ps;
      m[i] = m_new;
      float* prow = Ps + (warp * ATT_QPW + i) * 32;
      prow[lane] = p;
      __syncwarp();
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[i][c] *= corr;
#pragma unroll 8
      for (int j = 0; j < ATT_BK; ++j) {
        const float pj = prow[j];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[i][c] = fmaf(pj, Vs[j * DS + lane + 32 * c], acc[i][c]);
      }
      __syncwarp();
    }
  }

#pragma unroll
  for (int i = 0; i < ATT_QPW; ++i) {
    const int qi = q0 + warp * ATT_QPW + i;
    if (qi >= d.Nq) continue;
    const float inv = 1.f / l[i];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float val = acc[i][c] * inv;
      if (d.add_q_residual) val += Qs[(warp * ATT_QPW + i) * DS + lane + 32 * c];
      Elem<T>::st(ob + (long long)qi * d.o_row_stride + lane + 32 * c, val);
    }
  }
}

template <typename T, int D>
static int launch_attention(const pv_attention_desc* d, const void* q, const void* k, const void* v,
                            void* o, cudaStream_t s) {
  const size_t smem = (size_t)((ATT_BQ + 2 * ATT_BK) * (D + 1) + ATT_WARPS * ATT_QPW * 32) * sizeof(float);
  PV_OPT_IN_SMEM((attention_kernel<T, D>), smem);
  dim3 grid((unsigned)cdiv(d->Nq, ATT_BQ), (unsigned)(d->B * d->H)), block(ATT_WARPS * 32);
  attention_kernel<T, D><<<grid, block, smem, s>>>(*d, (const T*)q, (const T*)k, (const T*)v, (T*)o);
  PV_LAUNCH_OK("attention_kernel");
  return PV_OK;
}

int attention_tc_dispatch(const pv_attention_desc* d, const void* q, const void* k, const void* v, void* o, cudaStream_t s);
int attention_mma_dispatch(const pv_attention_desc* d, const void* q, const void* k, const void* v, void* o,
                           cudaStream_t s);   // pv_attention_mma.cu

}  // namespace pv

extern "C" int pv_attention_fwd(const pv_attention_desc* d, const void* q, const void* k,
                                const void* v, void* o, void* stream) {
  PV_CHECK_ARG(d && q && k && v && o, "null argument");
  PV_CHECK_ARG(d->dtype == PV_F16 || d->dtype == PV_F32, "attention dtype must be f16|f32");
  PV_CHECK_ARG(d->B > 0 && d->H > 0 && d->Nq > 0 && d->Nk > 0, "empty attention problem");
  PV_CHECK_ARG((long long)d->B * d->H <= 65535, "B*H too large");
  cudaStream_t s = (cudaStream_t)stream;
  if (d->dtype == PV_F16 && !getenv("PVB200_ATTN_SIMT")) {     // tensor-core paths (f16 storage)
    int rc = pv::attention_tc_dispatch(d, q, k, v, o, s);      // tcgen05 / TMEM / TMA (pv_attention_tc.cu)
    if (rc != PV_ERR_UNSUPPORTED) return rc;
    rc = pv::attention_mma_dispatch(d, q, k, v, o, s);         // mma.sync fallback (head dims / strides the TMA path rejects)
    if (rc != PV_ERR_UNSUPPORTED) return rc;
  }
#define PV_ATT(DD)                                                                              \
  if (d->D == DD)                                                                               \
    return d->dtype == PV_F16 ? pv::launch_attention<__half, DD>(d, q, k, v, o, s)              \
                              : pv::launch_attention<float, DD>(d, q, k, v, o, s);
  PV_ATT(32)
  PV_ATT(64)
  PV_ATT(96)
  PV_ATT(128)
#undef PV_ATT
  pv::set_error("attention head dim %d unsupported (32/64/96/128)", d->D);
  return PV_ERR_UNSUPPORTED;
}
